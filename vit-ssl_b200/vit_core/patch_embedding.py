"""Patch embedders (reference: vit_core/patch_embedding.py:11-128).

All three produce `[B, N+1, D]` fp32 tokens = [CLS, patches...] + positional embedding. The patch
projection is an im2col GEMM (patch features ordered (c, ph, pw), patches row-major) with the bias
in the GEMM epilogue; CLS / positional embedding are added by the token-assembly kernel.
"""
import torch
import torch.nn as nn

from ._backend import functional as Fb
from ._backend import eager


def _check_divisible(h, w, p):
    if h % p != 0 or w % p != 0:
        raise ValueError(f"Image dimensions H={h}, W={w} must be divisible by patch_size={p}")


class DynamicPatchEmbedding(nn.Module):
    """Conv patchify + CLS + positional embedding, bicubically interpolated when the input grid
    differs from the construction-time grid (patch_embedding.py:26-48)."""

    def __init__(self, input_shape, embed_dim, patch_size):
        super().__init__()
        self.patch_size = patch_size
        self.grid_size = (input_shape[1] // patch_size, input_shape[2] // patch_size)
        self.num_patches = self.grid_size[0] * self.grid_size[1]
        self.proj = nn.Conv2d(input_shape[0], embed_dim, kernel_size=patch_size, stride=patch_size)
        self.cls_token = nn.Parameter(torch.rand(1, 1, embed_dim))
        self.positional_embedding = nn.Parameter(torch.rand(1, self.num_patches + 1, embed_dim))

    @eager
    def interpolate_pos_encoding(self, x, w, h):
        """Positional embedding for a (w x h) patch grid: the trained one, or its patch part resized
        bicubically (align_corners=False) with the CLS row passed through (patch_embedding.py:26-48).
        The resize runs as a 16-tap row-interpolation kernel with precomputed tables, forward and
        backward, instead of F.interpolate on a permuted copy."""
        npatch = x.shape[1] if torch.is_tensor(x) else int(x)
        if npatch == self.num_patches and w == h:
            return self.positional_embedding
        return Fb.interpolate_pos_embedding(self.positional_embedding, self.grid_size, (w, h))

    @eager
    def forward(self, x):
        """x: [B,C,H,W] images, or a list of equally-shaped batches (crops of one resolution), which
        are embedded as if concatenated along the batch — without materialising the concatenation."""
        _, _, height, width = (x[0] if isinstance(x, (list, tuple)) else x).shape
        if height % self.patch_size != 0 or width % self.patch_size != 0:
            raise ValueError(
                f"Input image dimensions ({height}x{width}) must be divisible by patch size ({self.patch_size})."
            )
        gh, gw = height // self.patch_size, width // self.patch_size
        pos = self.interpolate_pos_encoding(gh * gw, gh, gw)
        return Fb.embed_patches(x, self, self.proj.weight, self.proj.bias, self.cls_token, pos, self.patch_size)


class ConvolutionalPatchEmbedding(nn.Module):
    def __init__(self, input_shape, embedding_dimension, patch_size):
        super().__init__()
        _check_divisible(input_shape[1], input_shape[2], patch_size)
        self.patch_size = patch_size
        self.conv = nn.Conv2d(input_shape[0], embedding_dimension, kernel_size=patch_size, stride=patch_size)
        self.cls_token = nn.Parameter(torch.rand(1, 1, embedding_dimension))
        self.positional_embedding = nn.Parameter(
            torch.rand(1, (input_shape[1] // patch_size) ** 2 + 1, embedding_dimension)
        )

    @eager
    def forward(self, x):
        return Fb.embed_patches(x, self, self.conv.weight, self.conv.bias, self.cls_token,
                                self.positional_embedding, self.patch_size)


class ManualPatchEmbedding(nn.Module):
    def __init__(self, input_shape, embedding_dimension, patch_size):
        super().__init__()
        _check_divisible(input_shape[1], input_shape[2], patch_size)
        self.patch_size = patch_size
        self.unfold = nn.Unfold(kernel_size=(patch_size, patch_size), stride=patch_size)  # stateless; kept for API parity
        self.linear = nn.Linear(input_shape[0] * patch_size * patch_size, embedding_dimension)
        self.cls_token = nn.Parameter(torch.rand(1, 1, embedding_dimension))
        self.positional_embedding = nn.Parameter(
            torch.rand(1, (input_shape[1] // patch_size) ** 2 + 1, embedding_dimension)
        )

    @eager
    def forward(self, x):
        return Fb.embed_patches(x, self, self.linear.weight, self.linear.bias, self.cls_token,
                                self.positional_embedding, self.patch_size)
