"""Scaled dot-product and multi-head attention (reference: vit_core/attention.py:5-106).

Behaviour kept from the reference: tuple returns `(out, probs | None)`, bias-free projections, the
1/sqrt(d_k) scale applied to the scores, softmax over the key axis, no attention dropout, arbitrary
query / key / value lengths. The math runs in the tcgen05 attention kernel (d_k = 64, S <= 256) or
the generic CUDA kernel otherwise.
"""
import torch
from torch import nn

from ._backend import functional as Fb
from ._backend import eager


@eager
def ScaledDotProductAttention(query, key, value, return_attn: bool = False):
    """softmax(Q K^T / sqrt(d_k)) V on tensors shaped (..., seq, d). Returns (context, probs|None)."""
    return Fb.scaled_dot_product_attention(query, key, value, return_attn)


class MultiHeadedAttention(nn.Module):
    def __init__(self, d_model: int, num_heads: int):
        super().__init__()
        assert d_model % num_heads == 0, (
            f"d_model({d_model}) must be cleanly divisible by num_heads({num_heads})!"
        )
        self.d_model = d_model
        self.num_heads = num_heads
        self.d_k = self.d_v = d_model // num_heads
        # parameter containers, created in the reference's order (attention.py:54-58)
        self.w_query = nn.Linear(d_model, d_model, bias=False)
        self.w_key = nn.Linear(d_model, d_model, bias=False)
        self.w_value = nn.Linear(d_model, d_model, bias=False)
        self.final_linear = nn.Linear(d_model, d_model, bias=False)

    @eager
    def forward(self, query, key, value, return_attn: bool = False):
        out, probs = Fb.multi_head_attention(self, query, key, value, return_attn)
        return Fb.autocast_out(out), probs
