"""ref: nerf/volume_rendering_utils.py:7-78."""
from __future__ import annotations

import torch

from . import ops


class _CompositeFn(torch.autograd.Function):
    """Alpha compositing with a hand-written backward (csrc/composite.cu)."""

    @staticmethod
    def forward(ctx, raw, z, rd, noise, bg, apply_bg, white):
        rgb, disp, acc, w, depth = torch.ops.sahs_b200.composite_fwd(raw, z, rd, noise, bg, apply_bg, white)
        ctx.save_for_backward(raw, z, rd, noise if noise is not None else torch.empty(0),
                              bg if bg is not None else torch.empty(0))
        ctx.flags = (noise is not None, bg is not None, apply_bg, white)
        ctx.w_last = w[:, -1].detach() if (bg is not None and ctx.needs_input_grad[4]) else None
        return rgb, disp, acc, w, depth

    @staticmethod
    def backward(ctx, d_rgb, d_disp, d_acc, d_w, d_depth):
        raw, z, rd, noise, bg = ctx.saved_tensors
        has_noise, has_bg, apply_bg, white = ctx.flags
        d_raw = torch.ops.sahs_b200.composite_bwd(raw, z, rd, noise if has_noise else None, bg if has_bg else None, apply_bg,
                                                  white, d_rgb, d_disp, d_acc, d_w, d_depth)
        # the background prior is an input of the blend rgb_map += w_last * bg (last sample = raw background values,
        # ref: nerf/volume_rendering_utils.py:28-33): a trainable background (`train_background`,
        # train_stage_rays_auto.py:171-176, :245) gets d bg = w_last * d rgb_map
        d_bg = None
        if has_bg and ctx.needs_input_grad[4]:
            if not apply_bg:
                raise RuntimeError("gradient w.r.t. background_prior needs the fused background overwrite")
            d_bg = ctx.w_last[:, None] * d_rgb if d_rgb is not None else torch.zeros_like(bg)
        return d_raw, None, None, None, d_bg, None, None


def composite(raw, z, rd, noise=None, bg=None, apply_bg_overwrite=False, white_background=False):
    if not raw.is_cuda:
        raise RuntimeError("sahs_b200 ops need CUDA tensors (there is no CPU path)")
    if torch.is_grad_enabled() and (raw.requires_grad or (bg is not None and bg.requires_grad)):
        return _CompositeFn.apply(raw, z, rd, noise, bg, apply_bg_overwrite, white_background)
    return torch.ops.sahs_b200.composite_fwd(raw, z, rd, noise, bg, apply_bg_overwrite, white_background)


def volume_render_radiance_field(radiance_field, depth_values, ray_directions, radiance_field_noise_std=0.0,
                                 white_background=False, background_prior=None):
    """Reference signature and return order (rgb_map, disp_map, acc_map, weights, depth_map).
    The caller has already written the background into radiance_field[:, -1, :-1] (ref: nerf/train_utils.py:
    135-136), so no overwrite happens here; gradients reach those entries exactly as in the reference."""
    noise = None
    if radiance_field_noise_std > 0.0:
        noise = torch.randn(radiance_field.shape[:-1], dtype=torch.float32,
                            device=radiance_field.device) * radiance_field_noise_std
    return composite(radiance_field, depth_values, ray_directions, noise, background_prior, False, white_background)
