"""SimMIM masked-image-modelling ViT (reference: vit_core/ssl/simmim/model.py:8-93).

forward(x) -> (predicted pixels [B*n_m, C*p*p], target pixels [B*n_m, C*p*p][, bool_mask [B,N,1]])
with rows in ascending (b, n) order of the mask. No CLS token. Patches are projected before the
mask-token substitution, so masked patches send no gradient into `projection` (model.py:45-48).
"""
import torch
from torch import nn

from .._backend_access import Fb, ops
from ...encoder_block import EncoderBlock
from .masking import draw_mask
from ..._backend import dp, eager
from ..._backend.scalar import prefetch_scalar


class MaskedPrediction(torch.Tensor):
    """The prediction tensor `SimMIMViT.forward` returns. It is an ordinary tensor in every respect
    but one: `nn.L1Loss(reduction="mean")(pred, targets)` — the criterion the reference trainer
    builds from configs/simmim/training.yaml:2-5 and calls at simmim_trainer.py:67 — dispatches to
    the fused masked-L1 kernel (one pass forward, one pass backward with the GradScaler factor read
    on the device) instead of torch's elementwise chain. Anything else strips the subclass."""

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        if func is torch.nn.functional.l1_loss:
            out = _fused_l1(*args, **kwargs)
            if out is not NotImplemented:
                return out
        with torch._C.DisableTorchFunctionSubclass():
            return func(*args, **kwargs)


def _fused_l1(input, target, size_average=None, reduce=None, reduction="mean", weight=None):
    if (size_average is not None or reduce is not None or reduction != "mean" or weight is not None
            or not (torch.is_tensor(input) and torch.is_tensor(target))
            or input.shape != target.shape or not (input.is_cuda and target.is_cuda)
            or input.dtype not in (torch.bfloat16, torch.float32)
            or target.dtype not in (torch.bfloat16, torch.float32) or target.requires_grad):
        return NotImplemented
    with torch._C.DisableTorchFunctionSubclass():
        pred = input.as_subclass(torch.Tensor)
        tgt = target.as_subclass(torch.Tensor)
        return prefetch_scalar(Fb.l1_loss(pred, tgt))


class SimMIMViT(nn.Module):
    def __init__(self, num_blocks: int, input_shape, embed_dim: int, patch_size: int, num_heads: int = 8,
                 mlp_dim: int = 3072, dropout: float = 0.1, mask_ratio: float = 0.6):
        super().__init__()
        self.encoder_blocks = nn.ModuleList(
            [EncoderBlock(embed_dim, num_heads, mlp_dim, dropout) for _ in range(num_blocks)]
        )
        self.unfold = nn.Unfold(kernel_size=(patch_size, patch_size), stride=patch_size)  # stateless, API parity
        self.projection = nn.Linear(input_shape[0] * patch_size * patch_size, embed_dim)
        self.mask_token = nn.Parameter(torch.randn(1, 1, embed_dim))
        self.positional_embedding = nn.Parameter(torch.rand(1, (input_shape[1] // patch_size) ** 2, embed_dim))
        self.simmim_head = nn.Linear(embed_dim, input_shape[0] * patch_size * patch_size)
        self.mask_ratio = mask_ratio
        self.input_shape = input_shape
        self.patch_size = patch_size

    # -- shared pieces --------------------------------------------------------------------
    def _encode_masked(self, x):
        B, C, H, W = x.shape
        p = self.patch_size
        N = (H // p) * (W // p)
        _, bool_mask, rows, inv = draw_mask(B, N, self.mask_ratio, x.device, want_indices=False)
        x = Fb._as_image(x)  # fp32 (ToTensor output) or raw uint8 bytes, scaled by 1/255 in the kernels
        targets = ops.gather_patches_f32(x, rows, p)
        tokens = Fb.embed_patches(x, self, self.projection.weight, self.projection.bias, None,
                                  self.positional_embedding, p, mask_u8=bool_mask.view(torch.uint8).reshape(-1),
                                  mask_token=self.mask_token)
        tokens, _ = Fb.encoder_stack(self.encoder_blocks, tokens)
        masked = Fb.gather_rows(tokens, rows, inv)
        return masked, targets, bool_mask

    @eager
    def forward(self, x, return_bool_mask=False):
        dp.maybe_attach(self)  # data parallel under torchrun without touching the trainer
        masked, targets, bool_mask = self._encode_masked(x)
        pred = Fb.autocast_out(Fb.mlp(masked, [self.simmim_head], [False])).as_subclass(MaskedPrediction)
        if return_bool_mask:
            return pred, targets, bool_mask.unsqueeze(-1)
        return pred, targets

    @eager
    def reconstruction_loss(self, x):
        """Fused objective: mean |pred - target| over the masked patches (what the reference trainer
        computes with nn.L1Loss on forward()'s outputs, simmim_trainer.py:66-67)."""
        dp.maybe_attach(self)  # data parallel under torchrun without touching the trainer
        masked, targets, _ = self._encode_masked(x)
        pred = Fb.mlp(masked, [self.simmim_head], [False])
        return prefetch_scalar(Fb.l1_loss(pred, targets))

    @torch.no_grad()
    @eager
    def inference_forward(self, x, return_patch_features=False):
        self.eval()
        tokens = Fb.embed_patches(x, self, self.projection.weight, self.projection.bias, None,
                                  self.positional_embedding, self.patch_size)
        tokens, _ = Fb.encoder_stack(self.encoder_blocks, tokens)
        return tokens if return_patch_features else ops.mean_tokens(tokens.contiguous())
