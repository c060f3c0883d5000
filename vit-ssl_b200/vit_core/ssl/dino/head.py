"""DINO projection head (reference: vit_core/ssl/dino/head.py:7-23): 3-layer GELU MLP
(D -> 2048 -> 2048 -> D), L2 normalisation, weight-normalised Linear(D -> K) with bias and
trainable magnitude g. State-dict keys match the reference's parametrization layout
(`fully_connected.parametrizations.weight.original0/1`)."""
import torch
from torch import nn
from torch.nn.utils.parametrizations import weight_norm

from .._backend_access import Fb
from ..._backend import eager


class DINOHead(nn.Module):
    def __init__(self, embed_dim, output_dim, hidden_dim=2048):
        super().__init__()
        self.mlp = nn.Sequential(
            nn.Linear(embed_dim, hidden_dim), nn.GELU(),
            nn.Linear(hidden_dim, hidden_dim), nn.GELU(),
            nn.Linear(hidden_dim, embed_dim),
        )
        # parameter container only: forward never materialises the fp32 normalised weight
        self.fully_connected = weight_norm(nn.Linear(embed_dim, output_dim), name="weight")

    @eager
    def forward(self, x):
        z = Fb.mlp(x, [self.mlp[0], self.mlp[2], self.mlp[4]], [True, True, False])
        wn = self.fully_connected.parametrizations.weight
        out_dtype = torch.bfloat16 if torch.is_autocast_enabled() else torch.float32
        return Fb.normalize_wn_linear(z, wn.original0, wn.original1, self.fully_connected.bias, out_dtype)
