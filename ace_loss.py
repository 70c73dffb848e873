"""Drop-in for the reference's `ace_loss` (reference ace_loss.py). On the CUDA training path the loss is evaluated inside
the fused tail kernel (csrc/repro_loss.cuh); this class keeps the public API (and serves PyTorch-side callers)."""
import numpy as np
import torch


def weighted_tanh(repro_errs, weight):
    return weight * torch.tanh(repro_errs / weight).sum()


class ReproLoss:
    """Per-pixel reprojection loss: tanh / dyntanh / l1 / l1+sqrt / l1+logl1 (reference ace_loss.py:11-90)."""

    def __init__(self, total_iterations, soft_clamp, soft_clamp_min, type='dyntanh', circle_schedule=True):
        self.total_iterations = total_iterations
        self.soft_clamp = soft_clamp
        self.soft_clamp_min = soft_clamp_min
        self.type = type
        self.circle_schedule = circle_schedule

    def loss_weight(self, iteration):
        """The tanh weight the CUDA kernel receives for `iteration` (ace_loss.py:53-69)."""
        if self.type == "dyntanh":
            schedule_weight = iteration / self.total_iterations
            if self.circle_schedule:
                schedule_weight = 1 - np.sqrt(1 - schedule_weight ** 2)
            return float((1 - schedule_weight) * self.soft_clamp + self.soft_clamp_min)
        return float(self.soft_clamp)

    def compute(self, repro_errs_b1N, iteration):
        if repro_errs_b1N.nelement() == 0:
            return 0
        if self.type in ("tanh", "dyntanh"):
            return weighted_tanh(repro_errs_b1N, self.loss_weight(iteration))
        mask = repro_errs_b1N > self.soft_clamp
        loss_l1 = repro_errs_b1N[~mask].sum()
        if self.type == "l1":
            return loss_l1
        if self.type == "l1+sqrt":
            return loss_l1 + torch.sqrt(self.soft_clamp * repro_errs_b1N[mask]).sum()
        return loss_l1 + torch.log(1 + (self.soft_clamp * repro_errs_b1N[mask])).sum()
