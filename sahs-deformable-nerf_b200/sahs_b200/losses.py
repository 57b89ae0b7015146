"""Per-class masked losses of the Stage-I training script.

ref: nerf/nerf_helpers.py:14-62 (MaskCrossEntropyLoss, MaskMSELoss), train_stage_rays_auto.py:455-468 (assembly).

`MaskMSELoss` / `MaskCrossEntropyLoss` mirror the reference's modules (same constructor, same 3-tuple result) for
callers that use them one at a time.  `stage1_loss` is what the training step calls: the whole assembly -- both
levels, the mouth term, the dynamic `sample_prob` and the gradient w.r.t. both maps -- in two small launches
(csrc/loss.cu, `sahs_stage1_loss`) behind a torch.autograd.Function.
"""
from __future__ import annotations

import torch

from . import lib as L

CE_WEIGHT, MOUTH_WEIGHT, MOUTH_CLASSES = 0.02, 0.005, (7, 9)      # train_stage_rays_auto.py:457-458


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


def stage1_loss_modules(rgb_coarse, rgb_fine, target_rgb, mask, mse=None, ce=None):
    """The assembly written with the two modules, as the reference's script does it (any device, ~75 launches):
    coarse + fine: l2 + 0.02 * CE + 0.005 * (mouth classes 7..8), ref: train_stage_rays_auto.py:455-465.
    Returns (loss, sample_prob) where sample_prob is the dynamic per-class sampling weight (:466-468)."""
    mse = mse or MaskMSELoss()
    ce = ce or MaskCrossEntropyLoss()
    maskf = mask.to(rgb_coarse.dtype)
    total, parts = 0.0, []
    for rgb in (rgb_coarse, rgb_fine):
        l2, m_l2, w_l2 = mse(maskf, rgb[..., :3], target_rgb[..., :3])
        c, m_c, w_c = ce(maskf, rgb[..., 3:], maskf)
        total = total + l2 + CE_WEIGHT * c + MOUTH_WEIGHT * torch.sum(m_l2[7:9] + m_c[7:9])
        parts += [w_l2, w_c]
    s = sum(parts)
    return total, (s / s.sum()).detach()


class _Stage1LossFn(torch.autograd.Function):
    """The forward call computes the loss, its statistics, sample_prob and d loss / d map; backward only scales."""

    @staticmethod
    def forward(ctx, map_c, map_f, target, mask):
        lib = L.load()
        dev = map_c.device
        mc, mf = L.f32c(map_c.reshape(-1, 15)), L.f32c(map_f.reshape(-1, 15))
        tg = L.f32c(target.reshape(-1, target.shape[-1])[:, :3])
        mk = L.f32c(mask.reshape(-1, 12))
        R = mc.shape[0]
        if mf.shape[0] != R or tg.shape[0] != R or mk.shape[0] != R:
            raise RuntimeError("stage1_loss: maps, target and mask must cover the same rays")
        # one allocation: statistics [53], sample_prob [12], pad to 16 B, per-CTA partial sums [4096]
        buf = torch.empty(68 + 4096, dtype=torch.float32, device=dev)
        stats, prob, work = buf[:53], buf[53:65], buf[68:]
        d_c, d_f = torch.empty_like(mc), torch.empty_like(mf)
        L.check(lib.sahs_stage1_loss(L.ptr(mc), L.ptr(mf), L.ptr(tg), L.ptr(mk), R, 12, CE_WEIGHT, MOUTH_WEIGHT,
                                     MOUTH_CLASSES[0], MOUTH_CLASSES[1], L.ptr(stats), L.ptr(prob), L.ptr(d_c),
                                     L.ptr(d_f), L.ptr(work), L.stream_ptr(dev)), "stage1_loss")
        ctx.save_for_backward(d_c, d_f)
        ctx.shapes = (map_c.shape, map_f.shape)
        ctx.mark_non_differentiable(prob, stats)
        return stats[0], prob, stats

    @staticmethod
    def backward(ctx, g, _gp, _gs):
        d_c, d_f = ctx.saved_tensors
        return (d_c * g).view(ctx.shapes[0]), (d_f * g).view(ctx.shapes[1]), None, None


def stage1_loss(rgb_coarse, rgb_fine, target_rgb, mask, return_stats: bool = False):
    """coarse + fine: l2 + 0.02 * CE + 0.005 * (mouth classes 7..8), ref: train_stage_rays_auto.py:455-465, and the
    dynamic per-class sampling weight `sample_prob` (:466-468), fused into two small kernels (sums, then gradients).
    rgb_* [R,15] maps from run_one_iter_of_nerf, target_rgb [R,>=3], mask [R,12].  Returns (loss, sample_prob)
    (+ the 53 statistics of include/sahs_b200.h when return_stats).  CUDA tensors only: there is no CPU path."""
    if not rgb_coarse.is_cuda:
        raise RuntimeError("stage1_loss needs CUDA tensors (there is no CPU fallback); "
                           "stage1_loss_modules is the module-by-module form")
    loss, prob, stats = _Stage1LossFn.apply(rgb_coarse, rgb_fine, target_rgb, mask)
    return (loss, prob, stats) if return_stats else (loss, prob)
