"""Pre-LN transformer encoder block (reference: vit_core/encoder_block.py:9-53).

x = x + drop(MHA(LN1(x)));  x = x + drop(FFN(LN2(x)));  returns (x, attn_probs | None).
The block (and any stack of blocks, see `_backend.functional.encoder_stack`) runs as one fused
sequence of kernels: add+LayerNorm, QKV GEMM, attention, out-proj GEMM, add+LayerNorm, FFN GEMMs.
"""
from torch import nn

from ._backend import functional as Fb
from .attention import MultiHeadedAttention
from .feed_forward import FeedForwardBlock
from ._backend import eager


class EncoderBlock(nn.Module):
    def __init__(self, d_model: int = 512, num_heads: int = 8, mlp_dim: int = 3072, dropout: float = 0.1):
        super().__init__()
        self.self_attention = MultiHeadedAttention(d_model, num_heads)
        self.feed_forward = FeedForwardBlock(d_model, mlp_dim, dropout)
        self.layer_norm1 = nn.LayerNorm(d_model)
        self.layer_norm2 = nn.LayerNorm(d_model)
        self.drop1 = nn.Dropout(dropout)
        self.drop2 = nn.Dropout(dropout)

    @eager
    def forward(self, x, return_attn=False):
        return Fb.encoder_stack([self], x, return_attn)
