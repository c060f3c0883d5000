"""DINO loss (reference: vit_core/ssl/dino/loss.py:7-29).

For the trainer's shapes teacher [G,B,K], student [V,B,K], center [1,K] the reference value is
  -(1/(G B K)) sum_b sum_k (sum_g softmax((T_g - c)/tau_t))_k (sum_v log_softmax(S_v/tau_s))_k
(all teacher/student view pairs including same-view pairs; the mean also divides by K). The
kernels evaluate this factorised form in a single pass over the logits; gradient flows to the
student only (teacher is detached, loss.py:22)."""
from torch import nn

from .._backend_access import Fb
from ..._backend import eager


class DINOLoss(nn.Module):
    def __init__(self, teacher_temp: float, student_temp: float):
        super().__init__()
        self.teacher_temp = teacher_temp
        self.student_temp = student_temp

    @eager
    def forward(self, teacher_output, student_output, center):
        if teacher_output.dim() != 3 or student_output.dim() != 3:
            raise ValueError(
                "DINOLoss expects teacher [G,B,K] and student [V,B,K] (as produced by the trainer, "
                f"dino_trainer.py:89-98); got {tuple(teacher_output.shape)} / {tuple(student_output.shape)}"
            )
        return Fb.dino_loss(teacher_output, student_output, center, self.teacher_temp, self.student_temp)
