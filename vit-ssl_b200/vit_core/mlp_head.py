"""Classification head: LayerNorm + Linear on the CLS token (reference: vit_core/mlp_head.py:6-15)."""
import torch
from torch import nn

from ._backend import functional as Fb
from ._backend import eager


class MLPHead(nn.Module):
    def __init__(self, d_model: int, num_classes: int):
        super().__init__()
        self.norm = nn.LayerNorm(d_model)
        self.linear = nn.Linear(d_model, num_classes)

    @eager
    def forward(self, x):
        h = Fb.layer_norm(x, self.norm)
        out_dtype = torch.bfloat16 if torch.is_autocast_enabled() else torch.float32
        return Fb.mlp(h, [self.linear], [False], out_dtype=out_dtype)
