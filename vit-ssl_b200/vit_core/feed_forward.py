"""Position-wise feed-forward block (reference: vit_core/feed_forward.py:7-28):
linear_out(dropout(gelu_erf(linear_in(x)))). Bias + GELU (+dropout) are fused into the epilogue of
the first tcgen05 GEMM; its backward fuses gelu' into the dgrad GEMM."""
from torch import nn

from ._backend import functional as Fb
from ._backend import eager


class FeedForwardBlock(nn.Module):
    def __init__(self, d_model: int = 512, d_ff: int = 2048, dropout: float = 0.1):
        super().__init__()
        self.linear_in = nn.Linear(d_model, d_ff)
        self.linear_out = nn.Linear(d_ff, d_model)
        self.dropout = nn.Dropout(dropout)

    @eager
    def forward(self, x):
        y = Fb.mlp(x, [self.linear_in, self.linear_out], [True, False], dropout_p=self.dropout.p,
                   training=self.training)
        return Fb.autocast_out(y)
