"""Supervised ViT (reference: vit_core/vit.py:9-45): ConvolutionalPatchEmbedding -> L encoder
blocks -> CLS token -> MLPHead. Construction order (blocks, patch embedding, head) matches the
reference so the same seed gives the same initial weights."""
from torch import nn

from ._backend import functional as Fb
from .encoder_block import EncoderBlock
from .mlp_head import MLPHead
from .patch_embedding import ConvolutionalPatchEmbedding
from ._backend import dp, eager


class ViT(nn.Module):
    def __init__(self, num_classes: int, num_blocks: int, input_shape, embed_dim: int, patch_size: int,
                 num_heads: int = 8, mlp_dim: int = 3072, dropout: float = 0.1):
        super().__init__()
        self.encoder_blocks = nn.ModuleList(
            [EncoderBlock(embed_dim, num_heads, mlp_dim, dropout) for _ in range(num_blocks)]
        )
        self.patch_embedding = ConvolutionalPatchEmbedding(input_shape, embed_dim, patch_size)
        self.classification_head = MLPHead(embed_dim, num_classes)

    @eager
    def forward(self, x, return_attn=False):
        dp.maybe_attach(self)  # data parallel under torchrun without touching the trainer
        x = self.patch_embedding(x)
        x, attn_probs = Fb.encoder_stack(self.encoder_blocks, x, return_attn)
        logits = self.classification_head(x[:, 0])
        if return_attn:
            return logits, attn_probs
        return logits
