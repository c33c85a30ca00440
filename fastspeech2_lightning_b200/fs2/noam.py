"""Noam learning-rate schedule — mirror of reference fs2/noam.py:5-26 (host-side scalar math)."""
from torch.optim.lr_scheduler import LRScheduler


class NoamLR(LRScheduler):
    """lr = base · w^0.5 · min(s^-0.5, s · w^-1.5), with step s clamped to ≥ 1."""

    def __init__(self, optimizer, warmup_steps):
        self.warmup_steps = warmup_steps
        super().__init__(optimizer)

    def get_lr(self):
        s = max(1, self.last_epoch)
        w = self.warmup_steps
        scale = w**0.5 * min(s ** (-0.5), s * w ** (-1.5))
        return [base_lr * scale for base_lr in self.base_lrs]
