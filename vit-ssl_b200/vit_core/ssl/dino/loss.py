"""DINO loss (reference: vit_core/ssl/dino/loss.py:7-29).

For the trainer's shapes teacher [G,B,K], student [V,B,K], center [1,K] the reference value is
  -(1/(G B K)) sum_b sum_k (sum_g softmax((T_g - c)/tau_t))_k (sum_v log_softmax(S_v/tau_s))_k
(all teacher/student view pairs including same-view pairs; the mean also divides by K). The
kernels evaluate this factorised form in a single pass over the logits; gradient flows to the
student only (teacher is detached, loss.py:22)."""
from torch import nn

from .._backend_access import Fb
from ..._backend import eager
from ..._backend.scalar import prefetch_scalar


class DINOLoss(nn.Module):
    def __init__(self, teacher_temp: float, student_temp: float):
        super().__init__()
        self.teacher_temp = teacher_temp
        self.student_temp = student_temp

    @eager
    def forward(self, teacher_output, student_output, center):
        """teacher [G, *rest, K], student [V, *rest, K] (the trainer passes [G,B,K] / [V,B,K],
        dino_trainer.py:89-98; 2-D inputs are the reference's all-pairs-across-rows case, i.e. rest
        = ()), center broadcastable to [K]. Any number of student views: the factorised loss is
        additive over groups of views, so more than 12 are processed as several kernel calls."""
        t, s = teacher_output, student_output
        if t.dim() < 2 or s.dim() != t.dim() or t.shape[1:] != s.shape[1:]:
            raise ValueError(
                "DINOLoss expects teacher [G,...,K] and student [V,...,K] with equal trailing dims; "
                f"got {tuple(t.shape)} / {tuple(s.shape)}")
        if center.numel() != t.shape[-1]:
            raise ValueError(f"center {tuple(center.shape)} does not match the head width {t.shape[-1]}")
        if t.shape[0] > 4:
            raise ValueError(f"DINOLoss kernels support up to 4 teacher views, got {t.shape[0]}")
        K = t.shape[-1]
        t3, s3 = t.reshape(t.shape[0], -1, K), s.reshape(s.shape[0], -1, K)
        loss = None
        for v0 in range(0, s3.shape[0], 12):
            part = Fb.dino_loss(t3, s3[v0:v0 + 12], center, self.teacher_temp, self.student_temp)
            loss = part if loss is None else loss + part
        return prefetch_scalar(loss)
