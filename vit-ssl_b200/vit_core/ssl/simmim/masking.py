"""SimMIM random patch masking (reference: vit_core/ssl/simmim/masking.py:6-37).

Bit-exactness: the mask is defined by B sequential `torch.randperm(N, device=...)[:n_m]` draws
from the device's default generator (masking.py:22-25). We issue exactly that call sequence, so
for the same generator state the mask equals the reference's bit for bit. Everything after the
draws is integer work done without host synchronisation (the reference's `patches[bool_mask]`
forces a `nonzero` sync; we derive the same row order from a sort of the drawn indices).
"""
from typing import Tuple

import torch


def draw_mask_indices(batch_size: int, num_patches: int, mask_ratio: float, device) -> torch.Tensor:
    num_masked = int(num_patches * mask_ratio)
    idx = [torch.randperm(num_patches, device=device)[:num_masked] for _ in range(batch_size)]
    return torch.stack(idx, dim=0)


def mask_tables(mask_indices: torch.Tensor, num_patches: int):
    """From drawn indices [B, n_m] build: bool mask [B,N], flat masked row ids (ascending (b, n)
    order == order of `x[bool_mask]`), and the inverse map row -> position or -1."""
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
    idx = draw_mask_indices(B, N, mask_ratio, patches.device)
    bool_mask, rows, _ = mask_tables(idx, N)
    targets = patches.reshape(B * N, P).index_select(0, rows.long())
    return patches, bool_mask, targets
