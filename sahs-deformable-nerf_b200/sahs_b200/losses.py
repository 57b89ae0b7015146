"""Per-class masked losses of the Stage-I training script (host-side torch ops).

ref: nerf/nerf_helpers.py:14-62 (MaskCrossEntropyLoss, MaskMSELoss), train_stage_rays_auto.py:455-465 (assembly).
"""
from __future__ import annotations

import torch


def _per_class(per_pixel: torch.Tensor, mask: torch.Tensor):
    count = torch.count_nonzero(mask, dim=0).clamp(min=1)
    return torch.sum(per_pixel * mask, dim=0) / count


class MaskMSELoss(torch.nn.Module):
    """forward(mask[.,12], input[.,3], target[.,3]) -> (mean squared error, per-class masked mse, weighted per-class)."""

    def __init__(self, weights=None):
        super().__init__()
        self.weights = weights

    def forward(self, mask, input, target):
        mask = mask.reshape(-1, mask.shape[-1])
        diff = torch.sum(torch.square(input.reshape(-1, 3) - target.reshape(-1, 3)), dim=-1, keepdim=True)
        masked = _per_class(diff, mask)
        w = self.weights if self.weights is not None else torch.ones(mask.shape[-1], device=mask.device)
        return torch.mean(diff), masked, w * masked


class MaskCrossEntropyLoss(torch.nn.Module):
    """forward(mask[.,12], input[.,12] (probabilities), target[.,12]) -> (mean CE, per-class CE, weighted per-class)."""

    def __init__(self, weights=None):
        super().__init__()
        self.weights = weights

    def forward(self, mask, input, target):
        mask = mask.reshape(-1, mask.shape[-1])
        input = input.reshape(-1, input.shape[-1])
        target = target.reshape(-1, target.shape[-1])
        ce = -torch.sum(target * torch.log(input + 1e-10), dim=-1, keepdim=True)
        masked = _per_class(ce, mask)
        w = self.weights if self.weights is not None else torch.ones(mask.shape[-1], device=mask.device)
        return torch.mean(ce), masked, w * masked


def stage1_loss(rgb_coarse, rgb_fine, target_rgb, mask, mse=None, ce=None):
    """coarse + fine: l2 + 0.02 * CE + 0.005 * (mouth classes 7..8), ref: train_stage_rays_auto.py:455-465.
    Returns (loss, sample_prob) where sample_prob is the dynamic per-class sampling weight (:466-468)."""
    mse = mse or MaskMSELoss()
    ce = ce or MaskCrossEntropyLoss()
    maskf = mask.to(rgb_coarse.dtype)
    total, parts = 0.0, []
    for rgb in (rgb_coarse, rgb_fine):
        l2, m_l2, w_l2 = mse(maskf, rgb[..., :3], target_rgb[..., :3])
        c, m_c, w_c = ce(maskf, rgb[..., 3:], maskf)
        total = total + l2 + 0.02 * c + 0.005 * torch.sum(m_l2[7:9] + m_c[7:9])
        parts += [w_l2, w_c]
    s = sum(parts)
    return total, (s / s.sum()).detach()
