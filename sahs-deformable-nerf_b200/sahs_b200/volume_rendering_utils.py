"""ref: nerf/volume_rendering_utils.py:7-78."""
from __future__ import annotations

import torch

from . import ops


class _CompositeFn(torch.autograd.Function):
    """Alpha compositing with a hand-written backward (csrc/composite.cu).  Density noise is either a tensor or drawn
    in the kernels from (noise_std, rng): forward and backward regenerate the same values, nothing is stored."""

    @staticmethod
    def forward(ctx, raw, z, rd, noise, bg, apply_bg, white, noise_std, rng):
        T = torch.ops.sahs_b200
        if noise is None and rng is not None and noise_std > 0.0:
            rgb, disp, acc, w, depth = T.composite_fwd_rng(raw, z, rd, float(noise_std), int(rng[0]), rng[1], int(rng[2]), bg,
                                                           apply_bg, white)
        else:
            rng = None
            rgb, disp, acc, w, depth = T.composite_fwd(raw, z, rd, noise, bg, apply_bg, white)
        ctx.save_for_backward(raw, z, rd, noise if noise is not None else torch.empty(0),
                              bg if bg is not None else torch.empty(0))
        ctx.flags = (noise is not None, bg is not None, apply_bg, white)
        ctx.rng, ctx.noise_std = rng, float(noise_std)
        ctx.w_last = w[:, -1].detach() if (bg is not None and ctx.needs_input_grad[4]) else None
        return rgb, disp, acc, w, depth

    @staticmethod
    def backward(ctx, d_rgb, d_disp, d_acc, d_w, d_depth):
        raw, z, rd, noise, bg = ctx.saved_tensors
        has_noise, has_bg, apply_bg, white = ctx.flags
        T = torch.ops.sahs_b200
        if ctx.rng is not None:
            d_raw = T.composite_bwd_rng(raw, z, rd, ctx.noise_std, int(ctx.rng[0]), ctx.rng[1], int(ctx.rng[2]),
                                        bg if has_bg else None, apply_bg, white, d_rgb, d_disp, d_acc, d_w, d_depth)
        else:
            d_raw = T.composite_bwd(raw, z, rd, noise if has_noise else None, bg if has_bg else None, apply_bg, white,
                                    d_rgb, d_disp, d_acc, d_w, d_depth)
        # the background prior is an input of the blend rgb_map += w_last * bg (last sample = raw background values,
        # ref: nerf/volume_rendering_utils.py:28-33): a trainable background (`train_background`,
        # train_stage_rays_auto.py:171-176, :245) gets d bg = w_last * d rgb_map
        d_bg = None
        if has_bg and ctx.needs_input_grad[4]:
            if not apply_bg:
                raise RuntimeError("gradient w.r.t. background_prior needs the fused background overwrite")
            d_bg = ctx.w_last[:, None] * d_rgb if d_rgb is not None else torch.zeros_like(bg)
        return d_raw, None, None, None, d_bg, None, None, None, None


def composite(raw, z, rd, noise=None, bg=None, apply_bg_overwrite=False, white_background=False, noise_std=0.0, rng=None):
    """noise: explicit [R,S] tensor (already scaled), or (noise_std, rng = (seed, counter tensor or None, stream)) for
    in-kernel draws, or neither for the deterministic path."""
    if not raw.is_cuda:
        raise RuntimeError("sahs_b200 ops need CUDA tensors (there is no CPU path)")
    if torch.is_grad_enabled() and (raw.requires_grad or (bg is not None and bg.requires_grad)):
        return _CompositeFn.apply(raw, z, rd, noise, bg, apply_bg_overwrite, white_background, float(noise_std), rng)
    if noise is None and rng is not None and noise_std > 0.0:
        return torch.ops.sahs_b200.composite_fwd_rng(raw, z, rd, float(noise_std), int(rng[0]), rng[1], int(rng[2]), bg,
                                                     apply_bg_overwrite, white_background)
    return torch.ops.sahs_b200.composite_fwd(raw, z, rd, noise, bg, apply_bg_overwrite, white_background)


def volume_render_radiance_field(radiance_field, depth_values, ray_directions, radiance_field_noise_std=0.0,
                                 white_background=False, background_prior=None):
    """Reference signature and return order (rgb_map, disp_map, acc_map, weights, depth_map).
    The caller has already written the background into radiance_field[:, -1, :-1] (ref: nerf/train_utils.py:
    135-136), so no overwrite happens here; gradients reach those entries exactly as in the reference."""
    rng = None
    if radiance_field_noise_std > 0.0:                      # noise drawn inside the kernels (stream 1)
        from .train_utils import next_rng
        seed, counter = next_rng()
        rng = (seed, counter, 1)
    return composite(radiance_field, depth_values, ray_directions, None, background_prior, False, white_background,
                     float(radiance_field_noise_std), rng)
