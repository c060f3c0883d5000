"""DINO student/teacher model (reference: vit_core/ssl/dino/model.py:12-155).

Kept from the reference: construction order (teacher backbone, deepcopy -> student, teacher head,
student head overwritten with the teacher head's weights), teacher parameters frozen, the `center`
buffer [1, K] updated INSIDE the teacher forward (also in eval) and rebound rather than mutated,
view-major outputs (teacher [G*B, K], student [V*B, K]), EMA over parameters only.

B200-native: the two student passes (global / local crops) share one head call; the center update
is a column-sum kernel + EMA kernel (with a cross-rank all-reduce of the 4*K-byte sums under data
parallelism); the EMA teacher update is one multi-tensor launch instead of ~470 tiny ones.
"""
import copy
from typing import List, Tuple

import torch
from torch import nn

from .._backend_access import Fb, dp, ops
from ...encoder_block import EncoderBlock
from ...patch_embedding import DynamicPatchEmbedding
from .head import DINOHead
from ..._backend import eager


class ViTBackbone(nn.Module):
    def __init__(self, num_blocks: int, input_shape: Tuple[int, int], embed_dim: int, patch_size: int,
                 num_heads: int = 8, mlp_dim: int = 3072, dropout: float = 0.1):
        super().__init__()
        self.encoder_blocks = nn.ModuleList(
            [EncoderBlock(embed_dim, num_heads, mlp_dim, dropout) for _ in range(num_blocks)]
        )
        self.patch_embedding = DynamicPatchEmbedding(input_shape, embed_dim, patch_size)

    @eager
    def forward(self, x, return_attn=False):
        x = self.patch_embedding(x)
        x, attn_probs = Fb.encoder_stack(self.encoder_blocks, x, return_attn)
        cls = x[:, 0]
        if return_attn:
            return cls, attn_probs
        return cls

    @eager
    def forward_views(self, crops):
        """CLS features of several crop batches of different resolution (model.py:117-118 runs the
        backbone once per resolution): one autograd node over all of them, rows concatenated in order."""
        tokens = [self.patch_embedding(c) for c in crops]
        outs = Fb.encoder_stack_multi(self.encoder_blocks, tokens)
        return torch.cat([o[:, 0] for o in outs], dim=0)


class DINOViT(nn.Module):
    def __init__(self, num_blocks: int, input_shape, embed_dim: int, patch_size: int, num_heads: int = 8,
                 mlp_dim: int = 3072, dropout: float = 0.1, output_dim: int = 65536,
                 center_momentum: float = 0.9):
        super().__init__()
        self.center_momentum = center_momentum
        self.teacher_backbone = ViTBackbone(num_blocks, input_shape, embed_dim, patch_size, num_heads,
                                            mlp_dim, dropout)
        self.student_backbone = copy.deepcopy(self.teacher_backbone)
        self.teacher_head = DINOHead(embed_dim, output_dim)
        self.student_head = DINOHead(embed_dim, output_dim)
        self.student_head.load_state_dict(self.teacher_head.state_dict())
        for p in self.teacher_backbone.parameters():
            p.requires_grad = False
        for p in self.teacher_head.parameters():
            p.requires_grad = False
        self.register_buffer("center", torch.zeros(1, output_dim))

    def _student_forward(self, x):
        return self.student_head(self.student_backbone(x))

    @torch.no_grad()
    @eager
    def _update_center(self, teacher_output):
        """center <- m * center + (1 - m) * mean over ALL ranks' teacher rows (model.py:91-99)."""
        t = teacher_output.detach()
        tb = t if t.dtype == torch.bfloat16 else Fb._as_bf16(t)
        colsum = ops.colsum_bf16(tb)
        dp.all_reduce_sum_(colsum)
        inv_rows = 1.0 / (tb.shape[0] * dp.world_size())
        new_center = ops.center_ema(self.center.reshape(-1).float().contiguous(), colsum,
                                    self.center_momentum, inv_rows)
        self.center = new_center.view(1, -1)

    def _teacher_forward(self, x):
        out = self.teacher_head(self.teacher_backbone(x))
        self._update_center(out.detach())
        return out

    @eager
    def forward(self, multi_crop_views: List[torch.Tensor], num_global_views: int):
        dp.maybe_attach(self)
        # view packing: the crops of one resolution go to the patch kernels as a list and
        # are unfolded into one patch matrix; the reference concatenates the images first (model.py:114-115)
        views = list(multi_crop_views)
        same = lambda vs: all(v.shape == vs[0].shape and v.dtype == vs[0].dtype for v in vs)
        global_views, local_views = views[:num_global_views], views[num_global_views:]
        global_crops = global_views if same(global_views) else torch.cat(global_views, dim=0)
        crops = [global_crops]
        if local_views:
            crops.append(local_views if same(local_views) else torch.cat(local_views, dim=0))
        # one backbone node and one head call for all views: rows stay view-major (globals first),
        # as in model.py:117-119
        student_output = self.student_head(self.student_backbone.forward_views(crops))
        with torch.no_grad():
            teacher_output = self._teacher_forward(global_crops)
        return teacher_output, student_output

    @torch.no_grad()
    @eager
    def momentum_update_teacher(self, teacher_momentum):
        """theta_t <- m theta_t + (1 - m) theta_s over backbone then head parameters, paired by
        registration order (model.py:126-139) — a single multi-tensor kernel."""
        student = list(self.student_backbone.parameters()) + list(self.student_head.parameters())
        teacher = list(self.teacher_backbone.parameters()) + list(self.teacher_head.parameters())
        shadows = [Fb.shadow_destination(t) for t in teacher]
        ops.multi_ema_shadow([t.data for t in teacher], [s.detach().data for s in student], shadows,
                             float(teacher_momentum))
        torch.autograd.graph.increment_version(teacher)  # the kernel wrote through raw pointers ...
        Fb.mark_shadows_fresh([t for t, sh in zip(teacher, shadows) if sh is not None])  # ... shadows included

    @torch.no_grad()
    @eager
    def inference_forward(self, x, return_features=False):
        self.eval()
        features = self.teacher_backbone(x)
        if return_features:
            return features
        return self.teacher_head(features)
