"""The per-ray pipeline with the reference's call signatures.

ref: nerf/train_utils.py:9-50 (run_network), :72-206 (predict_and_render_radiance), :209-321
(run_one_iter_of_nerf).  One call = a fixed sequence of CUDA kernels per ray chunk:
  coarse_z -> fold_frame -> field(coarse) -> composite -> sample_pdf+merge -> field(fine) -> composite
with no Python loop over point chunks (`chunksize` only bounds the ray chunk, as a memory hint).
"""
from __future__ import annotations

from typing import Optional

import torch

from . import ops
from .volume_rendering_utils import composite

# rays per internal chunk: raw_fine is [rays, 128, 16] fp32 = 8 KB/ray -> 4 GB at 512k rays
MAX_RAYS_PER_CHUNK = 1 << 19

# Measurement hook (bench.py): when set, called as hook(stage_name, fn) around every kernel stage of a ray chunk and
# must return fn().  None in production: a stage is then a plain call.
STAGE_HOOK = None


# In-kernel random draws (stochastic mode).  Every ray chunk takes a fresh Philox key derived from torch's seed and a
# host-side call counter, so `torch.manual_seed` makes runs reproducible.  RNG_COUNTER (optional CUDA int64 scalar) is
# added to the key ON THE DEVICE: a training step recorded in a CUDA graph (sahs_b200.train.GraphedStep) has its host-side
# key frozen, and advancing this counter once per step (ops.counter_add) is what gives every replay fresh draws.
RNG_COUNTER = None
_RNG_CALLS = [0]


def next_rng():
    """(seed, device counter or None) for the draws of one ray chunk."""
    _RNG_CALLS[0] += 1
    seed = (torch.initial_seed() * 0x9E3779B97F4A7C15 + _RNG_CALLS[0] * 0xD1B54A32D192ED03) & 0x7FFFFFFFFFFFFFFF
    return seed, RNG_COUNTER


def _stage(name, fn):
    return fn() if STAGE_HOOK is None else STAGE_HOOK(name, fn)


def run_network(level, network_fn, pts, ray_batch, chunksize, use_viewdirs, driving=None, pose=None, pose_c=None,
                latent_code=None, spatial_embeddings=None):
    """ref: nerf/train_utils.py:9-50.  Evaluates the field at explicit points (all chunks in one launch)."""
    dirs = ray_batch[..., None, 3:6].expand(pts.shape)
    x = torch.cat((pts.reshape(-1, 3), dirs.reshape(-1, 3)), dim=-1)
    raw = network_fn(level, x, driving, pose, pose_c, latent_code=latent_code)
    return raw.reshape(*pts.shape[:-1], raw.shape[-1])


def _render_chunk(model, nerf_opts, ro, rd, near, far, driving_vec, pose_code, bg, draws):
    R = ro.shape[0]
    dev = ro.device
    nc, nf = int(nerf_opts.num_coarse), int(nerf_opts.num_fine)
    perturb = bool(nerf_opts.perturb)
    noise_std = float(nerf_opts.radiance_field_noise_std)
    white = bool(nerf_opts.white_background)
    draws = draws or {}
    # Stochastic mode: the reference draws t_rand (train_utils.py:112), the density noise (volume_rendering_utils.py:47)
    # and u (nerf_helpers.py:473) from torch's generator.  Here they are drawn inside the kernels (Philox streams 0-3
    # of one key per chunk): no [R,S] tensor is materialised.  Tests inject the reference's own draws through `draws`.
    rng = next_rng() if (perturb or noise_std > 0.0) else None
    if not ro.is_cuda:
        raise RuntimeError("sahs_b200 ops need CUDA tensors (there is no CPU path)")
    T = torch.ops.sahs_b200
    t_vals = ops.linspace_dev(nc, dev)
    t_rand = draws.get("t_rand") if perturb else None
    if perturb and t_rand is None:
        z_c = T.coarse_z_rng(R, nc, near, far, bool(nerf_opts.lindisp), t_vals, rng[0], rng[1], 0)
    else:
        z_c = T.coarse_z(R, nc, near, far, bool(nerf_opts.lindisp), t_vals, t_rand)

    def comp(name, raw, z, key, stream):
        n = draws.get(key) if noise_std > 0.0 else None
        r = (rng[0], rng[1], stream) if (noise_std > 0.0 and n is None) else None
        return _stage(name, lambda: composite(raw, z, rd, n, bg, bg is not None, white, noise_std, r))

    raw_c = _stage("field_coarse", lambda: model.field("coarse", ro, rd, z_c, driving_vec, pose_code))
    rgb_c, disp_c, acc_c, w_c, depth_c = comp("composite_coarse", raw_c, z_c, "noise_c", 1)
    if nf <= 0:
        raise RuntimeError("num_fine == 0 is a dead branch in the reference (depth_fine undefined, "
                           "ref: nerf/train_utils.py:205-206); not supported")
    u = draws.get("u") if perturb else None                     # det = (perturb == 0.0), ref: nerf/train_utils.py:162
    if perturb and u is None:
        z_s, z_f = _stage("sample_pdf_merge",
                          lambda: T.sample_pdf_merge_rng(z_c, w_c.detach(), nf, rng[0], rng[1], 2))   # .detach(), ref: :164
    else:
        z_s, z_f = _stage("sample_pdf_merge", lambda: T.sample_pdf_merge(z_c, w_c.detach(), nf, u))
    raw_f = _stage("field_fine", lambda: model.field("fine", ro, rd, z_f, driving_vec, pose_code))
    rgb_f, disp_f, acc_f, w_f, depth_f = comp("composite_fine", raw_f, z_f, "noise_f", 3)
    return rgb_c, disp_c, acc_c, rgb_f, disp_f, acc_f, w_f[:, -1], depth_f


def predict_and_render_radiance(ray_batch, model, options, mode="train", driving=None, pose=None, pose_c=None,
                                background_prior=None, latent_code=None, spatial_embeddings=None, ray_dirs_fake=None,
                                _draws=None):
    """ref: nerf/train_utils.py:72-206.  ray_batch[r] = (ro 3 | rd 3 | near | far | mask...)."""
    ro, rd = ray_batch[..., :3], ray_batch[..., 3:6]
    # near / far ride in columns 6-7 of every ray (ref: :255-256) but are the config's two scalars: read them from the
    # options (no device sync); callers that really carry other bounds per batch pass them in the rays and set
    # options.dataset.near / far = None
    near, far = getattr(options.dataset, "near", None), getattr(options.dataset, "far", None)
    if near is None or far is None:
        near, far = float(ray_batch[0, 6]), float(ray_batch[0, 7])
    near, far = float(near), float(far)
    dvec = model.driving_vector(driving)
    pcode = model.pose_code(pose)
    return _render_chunk(model, getattr(options.nerf, mode), ro, rd, near, far, dvec, pcode, background_prior, _draws)


def run_one_iter_of_nerf(height, width, focal_length, model, ray_origins, ray_directions, options, mode="train",
                         driving=None, pose=None, pose_c=None, background_prior=None, latent_code=None,
                         ray_directions_ablation=None, spatial_embeddings=None, inHead=None, _draws=None):
    """ref: nerf/train_utils.py:209-321.  Returns the 8-tuple (rgb_coarse, disp_coarse, acc_coarse, rgb_fine,
    disp_fine, acc_fine, weights_fine[:, -1], depth_fine); "validation" mode restores image shapes."""
    if options.dataset.no_ndc is False:
        raise RuntimeError("NDC rays are a dead branch in the reference (every config sets no_ndc: True and "
                           "rd_ablations would be undefined, ref: nerf/train_utils.py:243-254); not supported")
    ro = ray_origins.reshape(-1, 3)
    rd = ray_directions.reshape(-1, 3)
    near, far = float(options.dataset.near), float(options.dataset.far)
    nerf_opts = getattr(options.nerf, mode)
    dvec = model.driving_vector(driving)            # once per call (the reference re-runs AudioNet per chunk)
    pcode = model.pose_code(pose)
    R = ro.shape[0]
    chunk = min(MAX_RAYS_PER_CHUNK, max(int(nerf_opts.chunksize), 1) * 4)
    outs = []
    for i in range(0, R, chunk):
        bg = background_prior[i:i + chunk] if background_prior is not None else None
        dr = None
        if _draws:
            dr = {k: (v[i:i + chunk] if v is not None else None) for k, v in _draws.items()}
        outs.append(_render_chunk(model, nerf_opts, ro[i:i + chunk], rd[i:i + chunk], near, far, dvec, pcode, bg, dr))
    images = [torch.cat(parts, dim=0) if len(parts) > 1 else parts[0] for parts in zip(*outs)]
    if mode == "validation":
        shp = tuple(ray_directions.shape[:-1])
        images = [im.view(*shp, im.shape[-1]) if im.dim() == 2 else im.view(shp) for im in images]
    return tuple(images)


class GaussianSmoothing(torch.nn.Module):
    """Depthwise Gaussian filter over 1-3 trailing dims, the module the training script builds for its (optional) mask
    smoothing (ref: nerf/train_utils.py:409-473): per-dimension kernel
    exp(-((i - mean) / (2 sigma))^2) / (sigma sqrt(2 pi)) -- the reference's own (non-standard) exponent --, product
    over dimensions, normalised to sum 1, applied with `padding=5`."""

    def __init__(self, channels, kernel_size, sigma, dim=2):
        super().__init__()
        import math
        import numbers
        sizes = [kernel_size] * dim if isinstance(kernel_size, numbers.Number) else list(kernel_size)
        sigmas = [sigma] * dim if isinstance(sigma, numbers.Number) else list(sigma)
        if dim not in (1, 2, 3):
            raise RuntimeError("Only 1, 2 and 3 dimensions are supported. Received {}.".format(dim))
        kernel = torch.ones([int(s) for s in sizes], dtype=torch.float32)
        for axis, (size, std) in enumerate(zip(sizes, sigmas)):
            i = torch.arange(int(size), dtype=torch.float32)
            g = torch.exp(-((i - (size - 1) / 2) / (2 * std)) ** 2) / (std * math.sqrt(2 * math.pi))
            shape = [1] * dim
            shape[axis] = int(size)
            kernel = kernel * g.view(shape)
        kernel = kernel / kernel.sum()
        self.register_buffer("weight", kernel.expand(channels, 1, *kernel.shape).clone())
        self.groups = channels
        self.conv = (torch.nn.functional.conv1d, torch.nn.functional.conv2d, torch.nn.functional.conv3d)[dim - 1]

    def forward(self, input):
        return self.conv(input, weight=self.weight, groups=self.groups, padding=5)
