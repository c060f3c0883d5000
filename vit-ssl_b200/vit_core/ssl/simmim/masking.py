"""SimMIM random patch masking (reference: vit_core/ssl/simmim/masking.py:6-37).

Bit-exactness: the reference's mask is defined by B sequential `torch.randperm(N, device=...)[:n_m]`
draws from the device's default generator (masking.py:22-25). `draw_mask` replays that integer
algorithm (Philox keys -> stable sort -> duplicate-key Fisher-Yates) for all B samples in ONE
kernel launch and advances the torch generator by exactly what the B calls would consume, so for
the same generator state the mask — and every later random draw — equals the reference's bit for
bit. The kernel also emits the bool mask, the masked-row ids in `x[bool_mask]` order and the
inverse map, so nothing here synchronises with the host (the reference's `patches[bool_mask]`
forces a `nonzero` sync). N > 1024 replays the reference's call sequence with torch itself.
"""
from typing import Tuple

import torch

from .._backend_access import ops
from ..._backend import eager


def _draw_with_torch(batch_size, num_patches, num_masked, device):
    idx = torch.stack([torch.randperm(num_patches, device=device)[:num_masked] for _ in range(batch_size)], dim=0)
    bool_mask, rows, inv = mask_tables(idx, num_patches)
    return idx, bool_mask, rows, inv


@eager
def draw_mask(batch_size: int, num_patches: int, mask_ratio: float, device, want_indices: bool = True):
    """-> (indices int64 [B,n_m] in draw order | None, bool_mask [B,N], rows int32 [B*n_m], inv int32 [B*N])."""
    num_masked = int(num_patches * mask_ratio)
    if num_patches > ops.RANDPERM_MAX_N:
        return _draw_with_torch(batch_size, num_patches, num_masked, device)
    return ops.simmim_mask(batch_size, num_patches, num_masked, device, want_perm=want_indices)


def draw_mask_indices(batch_size: int, num_patches: int, mask_ratio: float, device) -> torch.Tensor:
    return draw_mask(batch_size, num_patches, mask_ratio, device)[0]


def mask_tables(mask_indices: torch.Tensor, num_patches: int):
    """From drawn indices [B, n_m] build: bool mask [B,N], flat masked row ids (ascending (b, n)
    order == order of `x[bool_mask]`), and the inverse map row -> position or -1 (torch ops; the
    fused kernel produces the same tables directly)."""
    B, n_m = mask_indices.shape
    device = mask_indices.device
    bool_mask = torch.zeros((B, num_patches), dtype=torch.bool, device=device)
    bool_mask.scatter_(1, mask_indices, True)
    sorted_idx, _ = torch.sort(mask_indices, dim=1)
    rows = (sorted_idx + torch.arange(B, device=device).unsqueeze(1) * num_patches).reshape(-1)
    inv = torch.full((B * num_patches,), -1, dtype=torch.int32, device=device)
    inv[rows] = torch.arange(B * n_m, dtype=torch.int32, device=device)
    return bool_mask, rows.to(torch.int32), inv


def simple_masking(patches: torch.Tensor, mask_ratio: float) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Same contract as the reference: returns (patches, bool_mask [B,N], targets [B*n_m, P])."""
    B, N, P = patches.shape
    _, bool_mask, rows, _ = draw_mask(B, N, mask_ratio, patches.device, want_indices=False)
    targets = patches.reshape(B * N, P).index_select(0, rows.long())
    return patches, bool_mask, targets
