"""vit_core — B200-native drop-in for kristi700/ViT-SSL's `vit_core` package.

Same module paths, class names, constructor signatures, state_dict keys and return conventions as
the reference (vit_core/__init__.py:1-5); every forward/backward runs hand-written sm_100a CUDA
kernels through the C-ABI library `libvitssl_b200.so` (see include/vitssl_b200.h). There is no CPU
fallback: calling a module on CPU tensors raises `VitsslError`.
"""
from .vit import ViT
from .encoder_block import EncoderBlock
from .feed_forward import FeedForwardBlock
from .attention import MultiHeadedAttention, ScaledDotProductAttention
from .patch_embedding import ConvolutionalPatchEmbedding, ManualPatchEmbedding, DynamicPatchEmbedding

from . import optim as _optim

_optim.register()  # torch.optim.VitsslAdamW: fused AdamW selectable from the reference's config (optim.py)

__all__ = [
    "ViT", "EncoderBlock", "FeedForwardBlock", "MultiHeadedAttention", "ScaledDotProductAttention",
    "ConvolutionalPatchEmbedding", "ManualPatchEmbedding", "DynamicPatchEmbedding",
]
