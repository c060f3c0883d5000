"""Host-side DINO schedules (reference: vit_core/ssl/dino/dino_utils.py:4-36). Pure scalar math
evaluated once per epoch by the trainer (dino_trainer.py:46,80); no kernel involved."""
import math


def _half_cosine(start: float, end: float, frac: float) -> float:
    return end - (end - start) * 0.5 * (1.0 + math.cos(math.pi * frac))


class DINOMomentumScheduler:
    def __init__(self, m_start: float, m_end: float, total_iters: int):
        self.m_start, self.m_end, self.total_iters = m_start, m_end, total_iters

    def get_momentum(self, current_step: int) -> float:
        if current_step >= self.total_iters:
            return self.m_end
        return _half_cosine(self.m_start, self.m_end, current_step / self.total_iters)


class DINOTeacherTempScheduler:
    def __init__(self, temp_start: float, temp_end: float, total_iters: int, schedule_type: str = "cosine"):
        self.t_start, self.t_end = temp_start, temp_end
        self.total_iters, self.schedule_type = total_iters, schedule_type

    def get_temp(self, current_step: int) -> float:
        if current_step >= self.total_iters:
            return self.t_end
        frac = current_step / self.total_iters
        if self.schedule_type == "linear":
            return self.t_start + (self.t_end - self.t_start) * frac
        return _half_cosine(self.t_start, self.t_end, frac)
