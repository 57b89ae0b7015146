"""Tensor-level wrappers over the C ABI (one function per entry point of include/sahs_b200.h)."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import lib as L

MAP_CH = 15
RAW_CH = 16


_LINSPACE = {}


def linspace_dev(steps: int, device) -> torch.Tensor:
    """torch.linspace(0, 1, steps) made on the HOST (bit-identical to the reference's CPU linspace) and
    cached per device: one upload per process instead of one pageable H2D copy per call."""
    key = (int(steps), str(torch.device(device)))
    t = _LINSPACE.get(key)
    if t is None:
        t = torch.linspace(0.0, 1.0, steps, dtype=torch.float32).to(device)
        _LINSPACE[key] = t
    return t


_linspace_dev = linspace_dev


def operand_format() -> str:
    """16-bit operand format of trunk and heads on the render path of the loaded library: "f16" (default build) or
    "bf16" (`make BF16=1`, selected with SAHS_B200_LIB=.../libsahs_b200_bf16.so)."""
    return "bf16" if L.load().sahs_operand_format() == 1 else "f16"


def make_rng(seed: int, counter: Optional[torch.Tensor], stream: int) -> L.RngC:
    """struct sahs_rng: Philox key = seed (+ the device counter when given), `stream` selects the draw."""
    if counter is not None and (counter.dtype != torch.int64 or not counter.is_cuda or counter.numel() != 1):
        raise RuntimeError("rng counter must be a CUDA int64 scalar tensor")
    return L.RngC(int(seed) & 0xFFFFFFFFFFFFFFFF, L.ptr(counter) if counter is not None else None, int(stream), 0)


def rng_fill(n: int, seed: int, counter: Optional[torch.Tensor], stream: int, normal: bool, scale: float, device):
    """The values the kernels draw for (seed, counter, stream), materialised (tests / diagnostics)."""
    lib = L.load()
    out = torch.empty(int(n), dtype=torch.float32, device=device)
    r = make_rng(seed, counter, stream)
    L.check(lib.sahs_rng_fill(L.ptr(out), int(n), C.byref(r), int(bool(normal)), float(scale), L.stream_ptr(device)),
            "rng_fill")
    return out


def get_ray_bundle(height: int, width: int, intrinsics, c2w: torch.Tensor):
    """ref: nerf/nerf_helpers.py:178-233.  Returns (ro, rd) of shape (H, W, 3)."""
    lib = L.load()
    fx, fy, cx, cy = [float(v) for v in intrinsics]
    pose = L.f32c(c2w[:3, :4])
    ro = torch.empty(height, width, 3, dtype=torch.float32, device=pose.device)
    rd = torch.empty_like(ro)
    L.check(lib.sahs_get_ray_bundle(height, width, fx, fy, cx, cy, L.ptr(pose), L.ptr(ro), L.ptr(rd),
                                    L.stream_ptr(pose.device)), "get_ray_bundle")
    return ro, rd


def coarse_z(num_rays: int, num_samples: int, near: float, far: float, lindisp: bool, device,
             t_rand: Optional[torch.Tensor] = None, t_vals: Optional[torch.Tensor] = None, rng=None) -> torch.Tensor:
    """ref: nerf/train_utils.py:93-113.  rng = (seed, counter, stream): perturbed depths with t_rand drawn in the kernel."""
    lib = L.load()
    if t_vals is None:
        t_vals = _linspace_dev(num_samples, device)
    z = torch.empty(num_rays, num_samples, dtype=torch.float32, device=device)
    if rng is not None and t_rand is None:
        r = make_rng(*rng)
        L.check(lib.sahs_coarse_z_rng(num_rays, num_samples, float(near), float(far), int(bool(lindisp)), L.ptr(t_vals),
                                      C.byref(r), L.ptr(z), L.stream_ptr(device)), "coarse_z_rng")
        return z
    tr = L.f32c(t_rand) if t_rand is not None else None
    L.check(lib.sahs_coarse_z(num_rays, num_samples, float(near), float(far), int(bool(lindisp)), L.ptr(t_vals),
                              L.ptr(tr), L.ptr(z), L.stream_ptr(device)), "coarse_z")
    return z


def positional_encoding(x: torch.Tensor, num_freqs: int, include_input: bool = True) -> torch.Tensor:
    """ref: nerf/nerf_helpers.py:305-349 (log sampling)."""
    lib = L.load()
    xc = L.f32c(x)
    d = xc.shape[-1]
    n = xc.numel() // d
    width = d * ((1 if include_input else 0) + 2 * num_freqs)
    if width == d and include_input:
        return x
    out = torch.empty(*xc.shape[:-1], width, dtype=torch.float32, device=xc.device)
    L.check(lib.sahs_positional_encoding(L.ptr(xc), n, d, num_freqs, int(include_input), L.ptr(out),
                                         L.stream_ptr(xc.device)), "positional_encoding")
    return out


def field_fwd(cspec, level: int, packed, frame_const, grid, ro, rd, z, debug=None, debug_pass: int = -1):
    """raw[R,S,16] of the points ro + rd * z (the fused field kernel).  ref: nerf/train_utils.py:9-50 and everything
    below it (nerf/models.py:367-380, :514-528)."""
    lib = L.load()
    ro, rd, z = L.f32c(ro), L.f32c(rd), L.f32c(z)
    R, S = z.shape
    raw = torch.empty(R, S, RAW_CH, dtype=torch.float32, device=z.device)
    L.check(lib.sahs_field_fwd(C.byref(cspec), int(level), L.ptr(packed), L.ptr(frame_const), L.ptr(grid), L.ptr(ro),
                               L.ptr(rd), L.ptr(z), R, S, L.ptr(raw), L.ptr(debug), int(debug_pass),
                               L.stream_ptr(z.device)), "field_fwd")
    return raw


def tape_rows(num_points: int) -> int:
    """Tapes are tile-major chunk images: whole 128-point tiles are stored."""
    return (num_points + 127) // 128 * 128


def field_fwd_train(cspec, level: int, packed_train, frame_const, grid, ro, rd, z, tx_total: int, n_mask_layers: int):
    """Training forward: (raw, activation tape, sign masks, saves), see sahs_b200/train.py."""
    lib = L.load()
    ro, rd, z = L.f32c(ro), L.f32c(rd), L.f32c(z)
    R, S = z.shape
    P, dev = R * S, z.device
    raw = torch.empty(R, S, RAW_CH, dtype=torch.float32, device=dev)
    tape_x = torch.empty(tape_rows(P), tx_total, dtype=torch.float16, device=dev)     # [tiles][slots][128x64]
    masks = torch.empty(n_mask_layers, P, 2, 4, dtype=torch.int32, device=dev)
    saves = torch.empty(P, 8, dtype=torch.float32, device=dev)
    L.check(lib.sahs_field_fwd_train(C.byref(cspec), int(level), L.ptr(packed_train), L.ptr(frame_const), L.ptr(grid),
                                     L.ptr(ro), L.ptr(rd), L.ptr(z), R, S, L.ptr(raw), L.ptr(tape_x), L.ptr(masks),
                                     L.ptr(saves), L.stream_ptr(dev)), "field_fwd_train")
    return raw, tape_x, masks, saves


def field_bwd(cspec, level: int, packed_t, frame_const, grid, ro, rd, z, d_raw, scale, masks, saves, td_total: int):
    """Activation-gradient chain: (gradient tape, channel-last grid gradient), both multiplied by `scale`."""
    lib = L.load()
    ro, rd, z, d_raw = L.f32c(ro), L.f32c(rd), L.f32c(z), L.f32c(d_raw)
    R, S = z.shape
    P, dev = R * S, z.device
    tape_d = torch.empty(tape_rows(P), td_total, dtype=torch.float16, device=dev)      # [tiles][slots][128x64]
    grid_grad = torch.zeros(32, 32, 32, 32, dtype=torch.float32, device=dev)
    L.check(lib.sahs_field_bwd(C.byref(cspec), int(level), L.ptr(packed_t), L.ptr(frame_const), L.ptr(grid), L.ptr(ro),
                               L.ptr(rd), L.ptr(z), R, S, L.ptr(d_raw), L.ptr(scale), L.ptr(masks), L.ptr(saves),
                               L.ptr(tape_d), L.ptr(grid_grad), L.stream_ptr(dev)), "field_bwd")
    return tape_d, grid_grad


def composite_fwd(raw, z, rd, noise=None, bg=None, apply_bg_overwrite=False, white_background=False, noise_std=0.0,
                  rng=None):
    """ref: nerf/volume_rendering_utils.py:7-78.  Returns (rgb_map, disp, acc, weights, depth).  Density noise: the
    `noise` tensor, or (noise_std, rng = (seed, counter, stream)) drawn in the kernel."""
    lib = L.load()
    raw, z, rd = L.f32c(raw), L.f32c(z), L.f32c(rd)
    R, S = z.shape
    if raw.shape != (R, S, RAW_CH):
        raise RuntimeError(f"radiance_field must be [R,S,16], got {tuple(raw.shape)}")
    noise = L.f32c(noise) if noise is not None else None
    bg = L.f32c(bg) if bg is not None else None
    if bg is not None and bg.shape[-1] != MAP_CH:
        raise RuntimeError("background_prior must have 15 channels (rgb 3 + semantic 12)")
    dev = raw.device
    rgb = torch.empty(R, MAP_CH, dtype=torch.float32, device=dev)
    disp = torch.empty(R, dtype=torch.float32, device=dev)
    acc = torch.empty_like(disp)
    depth = torch.empty_like(disp)
    w = torch.empty(R, S, dtype=torch.float32, device=dev)
    if noise is None and rng is not None and noise_std > 0.0:
        r = make_rng(*rng)
        L.check(lib.sahs_composite_fwd_rng(L.ptr(raw), L.ptr(z), L.ptr(rd), float(noise_std), C.byref(r), L.ptr(bg), MAP_CH,
                                           int(apply_bg_overwrite), R, S, int(white_background), L.ptr(rgb), L.ptr(disp),
                                           L.ptr(acc), L.ptr(w), L.ptr(depth), L.stream_ptr(dev)), "composite_fwd_rng")
        return rgb, disp, acc, w, depth
    L.check(lib.sahs_composite_fwd(L.ptr(raw), L.ptr(z), L.ptr(rd), L.ptr(noise), L.ptr(bg), MAP_CH,
                                   int(apply_bg_overwrite), R, S, int(white_background), L.ptr(rgb), L.ptr(disp),
                                   L.ptr(acc), L.ptr(w), L.ptr(depth), L.stream_ptr(dev)), "composite_fwd")
    return rgb, disp, acc, w, depth


def composite_bwd(raw, z, rd, noise, bg, apply_bg_overwrite, white_background, d_rgb, d_disp, d_acc, d_w, d_depth,
                  noise_std=0.0, rng=None):
    lib = L.load()
    raw, z, rd = L.f32c(raw), L.f32c(z), L.f32c(rd)
    R, S = z.shape
    c = lambda t: L.f32c(t) if t is not None else None
    noise, bg, d_rgb, d_disp, d_acc, d_w, d_depth = map(c, (noise, bg, d_rgb, d_disp, d_acc, d_w, d_depth))
    d_raw = torch.empty_like(raw)
    if noise is None and rng is not None and noise_std > 0.0:
        r = make_rng(*rng)
        L.check(lib.sahs_composite_bwd_rng(L.ptr(raw), L.ptr(z), L.ptr(rd), float(noise_std), C.byref(r), L.ptr(bg), MAP_CH,
                                           int(apply_bg_overwrite), R, S, int(white_background), L.ptr(d_rgb),
                                           L.ptr(d_disp), L.ptr(d_acc), L.ptr(d_w), L.ptr(d_depth), L.ptr(d_raw),
                                           L.stream_ptr(raw.device)), "composite_bwd_rng")
        return d_raw
    L.check(lib.sahs_composite_bwd(L.ptr(raw), L.ptr(z), L.ptr(rd), L.ptr(noise), L.ptr(bg), MAP_CH,
                                   int(apply_bg_overwrite), R, S, int(white_background), L.ptr(d_rgb), L.ptr(d_disp),
                                   L.ptr(d_acc), L.ptr(d_w), L.ptr(d_depth), L.ptr(d_raw), L.stream_ptr(raw.device)),
            "composite_bwd")
    return d_raw


def sample_pdf_merge(z, weights, num_fine: int, u: Optional[torch.Tensor] = None, return_inds: bool = False, rng=None):
    """sample_pdf_2(mid(z), weights[...,1:-1], num_fine) + sort(cat(z, samples)).
    ref: nerf/nerf_helpers.py:454-497, nerf/train_utils.py:157-166.  u=None -> deterministic linspace, unless
    rng = (seed, counter, stream): u ~ U[0,1) drawn in the kernel."""
    lib = L.load()
    z, weights = L.f32c(z), L.f32c(weights)
    R, S = z.shape
    dev = z.device
    if u is None and rng is not None:
        zs = torch.empty(R, num_fine, dtype=torch.float32, device=dev)
        zm = torch.empty(R, S + num_fine, dtype=torch.float32, device=dev)
        inds = torch.empty(R, num_fine, dtype=torch.int64, device=dev) if return_inds else None
        r = make_rng(*rng)
        L.check(lib.sahs_sample_pdf_merge_rng(L.ptr(z), L.ptr(weights), C.byref(r), R, S, num_fine, L.ptr(zs), L.ptr(zm),
                                              L.ptr(inds), L.stream_ptr(dev)), "sample_pdf_merge_rng")
        return (zs, zm, inds) if return_inds else (zs, zm)
    if u is None:
        uu, per_ray = _linspace_dev(num_fine, dev), 0
    else:
        uu, per_ray = L.f32c(u), 1
        if uu.shape != (R, num_fine):
            raise RuntimeError("u must be [R, num_fine]")
    zs = torch.empty(R, num_fine, dtype=torch.float32, device=dev)
    zm = torch.empty(R, S + num_fine, dtype=torch.float32, device=dev)
    inds = torch.empty(R, num_fine, dtype=torch.int64, device=dev) if return_inds else None
    L.check(lib.sahs_sample_pdf_merge(L.ptr(z), L.ptr(weights), L.ptr(uu), per_ray, R, S, num_fine, L.ptr(zs),
                                      L.ptr(zm), L.ptr(inds), L.stream_ptr(dev)), "sample_pdf_merge")
    return (zs, zm, inds) if return_inds else (zs, zm)


def sample_pdf_bins(bins, weights, num_fine: int, u: Optional[torch.Tensor] = None, return_inds: bool = False):
    """sample_pdf_2(bins, weights, num_fine, det=(u is None)), ref: nerf/nerf_helpers.py:454-497."""
    lib = L.load()
    bins, weights = L.f32c(bins), L.f32c(weights)
    R, nb = bins.shape
    if weights.shape != (R, nb - 1):
        raise RuntimeError("weights must be [R, bins-1]")
    dev = bins.device
    if u is None:
        uu, per_ray = _linspace_dev(num_fine, dev), 0
    else:
        uu, per_ray = L.f32c(u), 1
    out = torch.empty(R, num_fine, dtype=torch.float32, device=dev)
    inds = torch.empty(R, num_fine, dtype=torch.int64, device=dev) if return_inds else None
    L.check(lib.sahs_sample_pdf(L.ptr(bins), L.ptr(weights), L.ptr(uu), per_ray, R, nb, num_fine, L.ptr(out),
                                L.ptr(inds), L.stream_ptr(dev)), "sample_pdf")
    return (out, inds) if return_inds else out


def frame_postprocess(rgb_map: torch.Tensor):
    """[..., 15] composited map -> (rgb uint8 [...,3], label uint8 [...], palette colour uint8 [...,3]).
    ref: eval_stage_rays.py:221-227 (cast_to_image), nerf/utils.py:112-140 (label2color)."""
    lib = L.load()
    m = L.f32c(rgb_map)
    shp = m.shape[:-1]
    n = m.numel() // MAP_CH
    rgb = torch.empty(*shp, 3, dtype=torch.uint8, device=m.device)
    lab = torch.empty(*shp, dtype=torch.uint8, device=m.device)
    col = torch.empty(*shp, 3, dtype=torch.uint8, device=m.device)
    L.check(lib.sahs_frame_postprocess(L.ptr(m), n, L.ptr(rgb), L.ptr(lab), L.ptr(col), L.stream_ptr(m.device)),
            "frame_postprocess")
    return rgb, lab, col


_SAMPLER_WS = {}
_SEL_POSITIVE_OFFSET = 4 * (256 + 6)      # byte offset of SelState::positive (csrc/sampler.cu)


def weighted_sample(mask: torch.Tensor, class_prob: torch.Tensor, num_select: int, seed: int,
                    validate: bool = False, seed_counter: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Semantic-weighted ray batch: `num_select` distinct pixel indices, drawn without replacement with probability
    proportional to sum_c class_prob[c] * mask[i, c] -- the device-side replacement of
    np.random.choice(H*W, n, replace=False, p=probs) (ref: train_stage_rays_auto.py:390-420).
    mask: int32 [..., C] (one-hot in the reference); returns int64 [num_select] (a set: order unspecified).
    With fewer positive-weight pixels than `num_select`, np.random.choice raises ("Fewer non-zero entries in p than
    size"); the kernel instead fills the remainder with zero-weight pixels.  `validate=True` reads the kernel's count of
    positive-weight pixels back (one device sync) and raises ValueError like numpy; the training loop leaves it off.
    `seed_counter` (int64 device scalar): the kernel uses seed + counter, read on the device -- the form a CUDA graph can
    replay (advance the counter with `counter_add` once per step)."""
    lib = L.load()
    if not mask.is_cuda:
        raise RuntimeError("weighted_sample needs CUDA tensors (there is no CPU fallback)")
    C_ = mask.shape[-1]
    m = mask.reshape(-1, C_)
    if m.dtype != torch.int32 or not m.is_contiguous():
        m = m.to(torch.int32).contiguous()
    n = m.shape[0]
    prob = L.f32c(class_prob.to(m.device))
    key = (m.device, n)
    ws = _SAMPLER_WS.get(key)
    if ws is None:
        ws = torch.empty(2048 + 4 * n, dtype=torch.uint8, device=m.device)
        _SAMPLER_WS[key] = ws
    out = torch.empty(num_select, dtype=torch.int64, device=m.device)
    if seed_counter is not None:
        if seed_counter.dtype != torch.int64 or not seed_counter.is_cuda or seed_counter.numel() != 1:
            raise RuntimeError("seed_counter must be a CUDA int64 scalar tensor")
        L.check(lib.sahs_weighted_sample_dev(L.ptr(m), L.ptr(prob), n, C_, int(num_select), int(seed) & 0xFFFFFFFFFFFFFFFF,
                                             L.ptr(seed_counter), L.ptr(out), L.ptr(ws), ws.numel(),
                                             L.stream_ptr(m.device)), "weighted_sample_dev")
    else:
        L.check(lib.sahs_weighted_sample(L.ptr(m), L.ptr(prob), n, C_, int(num_select), int(seed) & 0xFFFFFFFFFFFFFFFF,
                                         L.ptr(out), L.ptr(ws), ws.numel(), L.stream_ptr(m.device)), "weighted_sample")
    if validate:
        positive = int(ws[_SEL_POSITIVE_OFFSET:_SEL_POSITIVE_OFFSET + 4].view(torch.int32).item())
        if positive < num_select:
            raise ValueError(f"Fewer non-zero entries in p than size ({positive} < {num_select})")
    return out


def field_status():
    lib = L.load()
    out = (C.c_int * 4)()
    L.check(lib.sahs_field_status(out), "field_status")
    return list(out)


def counter_add(counter: torch.Tensor, inc: int = 1) -> None:
    """counter (CUDA int64 scalar) += inc, as one tiny launch on the current stream (capturable)."""
    lib = L.load()
    L.check(lib.sahs_counter_add(L.ptr(counter), int(inc), L.stream_ptr(counter.device)), "counter_add")


def normal_map(depthmap: torch.Tensor, focal, weights: Optional[torch.Tensor] = None, central_difference: bool = False):
    """`torch_normal_map` of the eval script (ref: eval_stage_rays.py:116-151): [N,N] depth -> [N-k,N-k,3] fp32 normals
    in [0,255] (k = 2 with central differences, else 1); `weights` = the fine pass's background weight for the clean-up."""
    lib = L.load()
    d = L.f32c(depthmap)
    if d.dim() != 2 or d.shape[0] != d.shape[1]:
        raise RuntimeError("torch_normal_map needs a square [N, N] depth map (the reference's meshgrid only broadcasts then)")
    n = d.shape[0]
    k = 2 if central_difference else 1
    w = L.f32c(weights) if weights is not None else None
    if w is not None and w.shape != d.shape:
        raise RuntimeError("weights must have the depth map's shape")
    fx, fy, cx, cy = [float(v) for v in focal]
    out = torch.empty(n - k, n - k, 3, dtype=torch.float32, device=d.device)
    L.check(lib.sahs_normal_map(L.ptr(d), n, fx, fy, cx, cy, L.ptr(w), int(bool(central_difference)), L.ptr(out),
                                L.stream_ptr(d.device)), "normal_map")
    return out


# ---- Stage II (SPADE generator): csrc/spade_conv.cu -----------------------------------------------------------------
def spade_conv(x, packed, bias, cin: int, cout: int, out_h: int, out_w: int, mode: int, up: int = 0, down: int = 0,
               epilogue: int = 0, aux=None, aux_shift: int = 0, mean=None, rstd=None):
    """One 3x3 convolution of the Stage-II network on NHWC fp16 (include/sahs_b200.h `sahs_conv_desc`).
    x [in_h, in_w, cs]; packed [ntiles, chunks, ntile, 64] fp16 swizzled blocks; returns [out_h, out_w, cout]
    (fp16, or fp32 with epilogue flag 8).  ref: nerf/_init_spade.py:114-139, :235-282."""
    lib = L.load()
    dev = x.device
    f32 = bool(epilogue & 8)
    out = torch.empty(out_h, out_w, cout, dtype=torch.float32 if f32 else torch.float16, device=dev)
    d = L.ConvDescC()
    d.in_, d.in_h, d.in_w, d.in_cs, d.cin = L.ptr(x), x.shape[0], x.shape[1], x.stride(1), int(cin)
    d.out_h, d.out_w, d.mode, d.up_shift, d.down_shift = int(out_h), int(out_w), int(mode), int(up), int(down)
    d.packed_w, d.bias, d.ntile, d.ntiles, d.epilogue = L.ptr(packed), L.ptr(bias), packed.shape[2], packed.shape[0], int(epilogue)
    d.aux, d.aux_cs, d.aux_shift = (L.ptr(aux), aux.stride(1), int(aux_shift)) if aux is not None else (None, 0, 0)
    d.mean, d.rstd = L.ptr(mean), L.ptr(rstd)
    d.out, d.out_cs, d.cout = L.ptr(out), int(cout), int(cout)
    L.check(lib.sahs_spade_conv(C.byref(d), L.stream_ptr(dev)), "spade_conv")
    return out


def spade_conv_t2(x, packed4, bias4, cin: int, cout: int):
    """nn.ConvTranspose2d(3x3, stride 2, padding 1, output_padding 1) on NHWC fp16: four parity-class launches (mode 4)
    that together write every pixel of the [2 in_h, 2 in_w, cout] output.  ref: nerf/_init_spade.py:255."""
    lib = L.load()
    dev = x.device
    out = torch.empty(2 * x.shape[0], 2 * x.shape[1], cout, dtype=torch.float16, device=dev)
    for cls in range(4):
        packed, bias = packed4[cls], bias4[cls]
        d = L.ConvDescC()
        d.in_, d.in_h, d.in_w, d.in_cs, d.cin = L.ptr(x), x.shape[0], x.shape[1], x.stride(1), int(cin)
        d.out_h, d.out_w, d.mode, d.up_shift, d.down_shift = out.shape[0], out.shape[1], 4, 0, 0
        d.packed_w, d.bias, d.ntile, d.ntiles, d.epilogue = L.ptr(packed), L.ptr(bias), packed.shape[2], packed.shape[0], 0
        d.aux, d.aux_cs, d.aux_shift, d.mean, d.rstd = None, 0, 0, None, None
        d.out, d.out_cs, d.cout, d.t2_class = L.ptr(out), int(cout), int(cout), cls
        L.check(lib.sahs_spade_conv(C.byref(d), L.stream_ptr(dev)), "spade_conv (transposed, class %d)" % cls)
    return out


def instnorm_stats(x, eps: float = 1e-5):
    """Per-channel (mean, 1 / sqrt(biased var + eps)) of an NHWC fp16 tensor [H, W, C] (nn.InstanceNorm2d, affine=False)."""
    lib = L.load()
    c = x.shape[2]
    ws = torch.empty(512 * c, dtype=torch.float64, device=x.device)          # one fp64 partial per block (<= 256 blocks)
    mean = torch.empty(c, dtype=torch.float32, device=x.device)
    rstd = torch.empty(c, dtype=torch.float32, device=x.device)
    L.check(lib.sahs_instnorm_stats(L.ptr(x), x.shape[0] * x.shape[1], c, x.stride(1), float(eps), L.ptr(ws), L.ptr(mean),
                                    L.ptr(rstd), L.stream_ptr(x.device)), "instnorm_stats")
    return mean, rstd


def avgpool2(x):
    """nn.AvgPool2d(2, stride=2) on NHWC fp16 [H, W, C]."""
    lib = L.load()
    y = torch.empty(x.shape[0] // 2, x.shape[1] // 2, x.shape[2], dtype=torch.float16, device=x.device)
    L.check(lib.sahs_avgpool2(L.ptr(x), x.shape[0], x.shape[1], x.shape[2], L.ptr(y), L.stream_ptr(x.device)), "avgpool2")
    return y


def spade_conv_status():
    lib = L.load()
    out = (C.c_int * 4)()
    L.check(lib.sahs_spade_conv_status(out), "spade_conv_status")
    return list(out)
