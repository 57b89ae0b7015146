"""PyTorch custom operators over the C ABI: `torch.ops.sahs_b200.*`.

The reference has no operator layer of its own (its boundary is the Python surface of `nerf`);
the north star asks for the four stages "exposed as PyTorch custom ops over a thin C-ABI extension".  Every hot-path
entry point of include/sahs_b200.h is registered here with a schema, a CUDA implementation (the ctypes call in
sahs_b200/ops.py -- raw device pointers, current stream) and a Meta implementation (shapes only, so the ops trace
under FakeTensor / torch.export).  There is deliberately NO CPU implementation: dispatching a CPU tensor raises
NotImplementedError from the dispatcher -- the product path has no fallback.

  stage (1)  sahs_b200::get_ray_bundle, ::coarse_z, ::positional_encoding
  stage (2)  sahs_b200::field_fwd, ::field_fwd_train, ::field_bwd
  stage (3)  sahs_b200::composite_fwd, ::composite_bwd
  stage (4)  sahs_b200::sample_pdf, ::sample_pdf_merge
  next rows  sahs_b200::frame_postprocess, ::normal_map, ::weighted_sample; Stage II: ::spade_conv, ::spade_conv_t2, ::instnorm_stats, ::avgpool2

The reference-named Python functions (train_utils / nerf_helpers / volume_rendering_utils / models) call these ops;
autograd is wired by torch.autograd.Function classes on top (volume_rendering_utils._CompositeFn, train.FieldTrainFn).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import torch

from . import lib as L
from . import ops

_lib = torch.library.Library("sahs_b200", "DEF")

_SPEC_FIELDS = [n for n, _ in L.ModelSpecC._fields_]
_SPEC_CACHE = {}


def spec_ints(cspec: L.ModelSpecC) -> List[int]:
    return [int(getattr(cspec, n)) for n in _SPEC_FIELDS]


def _cspec(ints) -> L.ModelSpecC:
    key = tuple(int(v) for v in ints)
    c = _SPEC_CACHE.get(key)
    if c is None:
        if len(key) != len(_SPEC_FIELDS):
            raise RuntimeError("model spec must have %d integers (struct sahs_model_spec)" % len(_SPEC_FIELDS))
        c = L.ModelSpecC(*key)
        _SPEC_CACHE[key] = c
    return c


def _define(schema: str, cuda_impl, meta_impl):
    _lib.define(schema)
    name = schema.split("(", 1)[0]
    _lib.impl(name, cuda_impl, "CUDA")
    _lib.impl(name, meta_impl, "Meta")


def _e(shape, like, dtype=torch.float32):
    return torch.empty(shape, dtype=dtype, device=like.device)


# ---- stage (1) ------------------------------------------------------------------------------------------------------
_define("get_ray_bundle(int height, int width, float fx, float fy, float cx, float cy, Tensor pose) -> (Tensor, Tensor)",
        lambda h, w, fx, fy, cx, cy, pose: ops.get_ray_bundle(h, w, (fx, fy, cx, cy), pose),
        lambda h, w, fx, fy, cx, cy, pose: (_e((h, w, 3), pose), _e((h, w, 3), pose)))

_define("coarse_z(int num_rays, int num_samples, float near, float far, bool lindisp, Tensor t_vals, Tensor? t_rand) -> Tensor",
        lambda r, s, near, far, lindisp, t_vals, t_rand: ops.coarse_z(r, s, near, far, lindisp, t_vals.device, t_rand, t_vals),
        lambda r, s, near, far, lindisp, t_vals, t_rand: _e((r, s), t_vals))

_define("positional_encoding(Tensor x, int num_freqs, bool include_input) -> Tensor",
        lambda x, n, inc: ops.positional_encoding(x, n, inc),
        lambda x, n, inc: _e(tuple(x.shape[:-1]) + (x.shape[-1] * ((1 if inc else 0) + 2 * n),), x))


_define("coarse_z_rng(int num_rays, int num_samples, float near, float far, bool lindisp, Tensor t_vals, int seed, "
        "Tensor? counter, int stream) -> Tensor",
        lambda r, s, near, far, lindisp, t_vals, seed, ctr, st:
            ops.coarse_z(r, s, near, far, lindisp, t_vals.device, None, t_vals, rng=(seed, ctr, st)),
        lambda r, s, near, far, lindisp, t_vals, seed, ctr, st: _e((r, s), t_vals))


# ---- stage (2) ------------------------------------------------------------------------------------------------------
def _field_fwd_cuda(spec, level, packed, fc, grid, ro, rd, z):
    return ops.field_fwd(_cspec(spec), level, packed, fc, grid, ro, rd, z)


_define("field_fwd(int[] spec, int level, Tensor packed, Tensor frame_const, Tensor grid, Tensor ro, Tensor rd, Tensor z) -> Tensor",
        _field_fwd_cuda, lambda spec, level, packed, fc, grid, ro, rd, z: _e(tuple(z.shape) + (16,), z))


def _field_fwd_train_cuda(spec, level, packed_train, fc, grid, ro, rd, z, tx_total, n_mask_layers):
    return ops.field_fwd_train(_cspec(spec), level, packed_train, fc, grid, ro, rd, z, tx_total, n_mask_layers)


def _field_fwd_train_meta(spec, level, packed_train, fc, grid, ro, rd, z, tx_total, n_mask_layers):
    P = z.shape[0] * z.shape[1]
    rows = (P + 127) // 128 * 128
    return (_e(tuple(z.shape) + (16,), z), _e((rows, tx_total), z, torch.float16),
            _e((n_mask_layers, P, 2, 4), z, torch.int32), _e((P, 8), z))


_define("field_fwd_train(int[] spec, int level, Tensor packed_train, Tensor frame_const, Tensor grid, Tensor ro, Tensor rd, "
        "Tensor z, int tx_total, int n_mask_layers) -> (Tensor, Tensor, Tensor, Tensor)", _field_fwd_train_cuda,
        _field_fwd_train_meta)


def _field_bwd_cuda(spec, level, packed_t, fc, grid, ro, rd, z, d_raw, scale, masks, saves, td_total):
    return ops.field_bwd(_cspec(spec), level, packed_t, fc, grid, ro, rd, z, d_raw, scale, masks, saves, td_total)


def _field_bwd_meta(spec, level, packed_t, fc, grid, ro, rd, z, d_raw, scale, masks, saves, td_total):
    rows = (z.shape[0] * z.shape[1] + 127) // 128 * 128
    return _e((rows, td_total), z, torch.float16), _e((32, 32, 32, 32), z)


_define("field_bwd(int[] spec, int level, Tensor packed_t, Tensor frame_const, Tensor grid, Tensor ro, Tensor rd, Tensor z, "
        "Tensor d_raw, Tensor scale, Tensor masks, Tensor saves, int td_total) -> (Tensor, Tensor)", _field_bwd_cuda,
        _field_bwd_meta)


# ---- stage (3) ------------------------------------------------------------------------------------------------------
def _composite_meta(raw, z, rd, noise, bg, apply_bg, white):
    R, S = z.shape
    return _e((R, 15), z), _e((R,), z), _e((R,), z), _e((R, S), z), _e((R,), z)


_define("composite_fwd(Tensor raw, Tensor z, Tensor rd, Tensor? noise, Tensor? bg, bool apply_bg_overwrite, bool white_background)"
        " -> (Tensor, Tensor, Tensor, Tensor, Tensor)",
        lambda raw, z, rd, noise, bg, apply_bg, white: ops.composite_fwd(raw, z, rd, noise, bg, apply_bg, white),
        _composite_meta)

_define("composite_bwd(Tensor raw, Tensor z, Tensor rd, Tensor? noise, Tensor? bg, bool apply_bg_overwrite, bool white_background, "
        "Tensor? d_rgb, Tensor? d_disp, Tensor? d_acc, Tensor? d_w, Tensor? d_depth) -> Tensor",
        lambda raw, z, rd, noise, bg, apply_bg, white, d_rgb, d_disp, d_acc, d_w, d_depth:
            ops.composite_bwd(raw, z, rd, noise, bg, apply_bg, white, d_rgb, d_disp, d_acc, d_w, d_depth),
        lambda raw, *a: torch.empty_like(raw))


_define("composite_fwd_rng(Tensor raw, Tensor z, Tensor rd, float noise_std, int seed, Tensor? counter, int stream, Tensor? bg, "
        "bool apply_bg_overwrite, bool white_background) -> (Tensor, Tensor, Tensor, Tensor, Tensor)",
        lambda raw, z, rd, std, seed, ctr, st, bg, apply_bg, white:
            ops.composite_fwd(raw, z, rd, None, bg, apply_bg, white, noise_std=std, rng=(seed, ctr, st)),
        lambda raw, z, rd, std, seed, ctr, st, bg, apply_bg, white: _composite_meta(raw, z, rd, None, bg, apply_bg, white))

_define("composite_bwd_rng(Tensor raw, Tensor z, Tensor rd, float noise_std, int seed, Tensor? counter, int stream, Tensor? bg, "
        "bool apply_bg_overwrite, bool white_background, Tensor? d_rgb, Tensor? d_disp, Tensor? d_acc, Tensor? d_w, "
        "Tensor? d_depth) -> Tensor",
        lambda raw, z, rd, std, seed, ctr, st, bg, apply_bg, white, d_rgb, d_disp, d_acc, d_w, d_depth:
            ops.composite_bwd(raw, z, rd, None, bg, apply_bg, white, d_rgb, d_disp, d_acc, d_w, d_depth, noise_std=std,
                              rng=(seed, ctr, st)),
        lambda raw, *a: torch.empty_like(raw))


# ---- stage (4) ------------------------------------------------------------------------------------------------------
_define("sample_pdf_merge(Tensor z, Tensor weights, int num_fine, Tensor? u) -> (Tensor, Tensor)",
        lambda z, w, nf, u: ops.sample_pdf_merge(z, w, nf, u),
        lambda z, w, nf, u: (_e((z.shape[0], nf), z), _e((z.shape[0], z.shape[1] + nf), z)))

_define("sample_pdf_merge_rng(Tensor z, Tensor weights, int num_fine, int seed, Tensor? counter, int stream) -> (Tensor, Tensor)",
        lambda z, w, nf, seed, ctr, st: ops.sample_pdf_merge(z, w, nf, None, rng=(seed, ctr, st)),
        lambda z, w, nf, seed, ctr, st: (_e((z.shape[0], nf), z), _e((z.shape[0], z.shape[1] + nf), z)))

_define("sample_pdf(Tensor bins, Tensor weights, int num_samples, Tensor? u) -> Tensor",
        lambda bins, w, n, u: ops.sample_pdf_bins(bins, w, n, u),
        lambda bins, w, n, u: _e((bins.shape[0], n), bins))


# ---- section 8(f) rows ----------------------------------------------------------------------------------------------
_define("frame_postprocess(Tensor rgb_map) -> (Tensor, Tensor, Tensor)",
        lambda m: ops.frame_postprocess(m),
        lambda m: (_e(tuple(m.shape[:-1]) + (3,), m, torch.uint8), _e(tuple(m.shape[:-1]), m, torch.uint8),
                   _e(tuple(m.shape[:-1]) + (3,), m, torch.uint8)))

_define("weighted_sample(Tensor mask, Tensor class_prob, int num_select, int seed) -> Tensor",
        lambda mask, prob, n, seed: ops.weighted_sample(mask, prob, n, seed),
        lambda mask, prob, n, seed: _e((n,), mask, torch.int64))

_define("normal_map(Tensor depthmap, float fx, float fy, float cx, float cy, Tensor? weights, bool central_difference) -> Tensor",
        lambda d, fx, fy, cx, cy, w, central: ops.normal_map(d, (fx, fy, cx, cy), w, central),
        lambda d, fx, fy, cx, cy, w, central: _e((d.shape[0] - (2 if central else 1), d.shape[1] - (2 if central else 1), 3), d))

# ---- Stage II (SPADE generator) -----------------------------------------------------------------------------------------
_define("spade_conv(Tensor x, Tensor packed, Tensor bias, int cin, int cout, int out_h, int out_w, int mode, int up, int down, "
        "int epilogue, Tensor? aux, int aux_shift, Tensor? mean, Tensor? rstd) -> Tensor",
        lambda x, packed, bias, cin, cout, oh, ow, mode, up, down, epi, aux, ash, mean, rstd:
            ops.spade_conv(x, packed, bias, cin, cout, oh, ow, mode, up, down, epi, aux, ash, mean, rstd),
        lambda x, packed, bias, cin, cout, oh, ow, mode, up, down, epi, aux, ash, mean, rstd:
            _e((oh, ow, cout), x, torch.float32 if (epi & 8) else torch.float16))

_define("spade_conv_t2(Tensor x, Tensor[] packed, Tensor[] bias, int cin, int cout) -> Tensor",
        lambda x, packed, bias, cin, cout: ops.spade_conv_t2(x, packed, bias, cin, cout),
        lambda x, packed, bias, cin, cout: _e((2 * x.shape[0], 2 * x.shape[1], cout), x, torch.float16))

_define("instnorm_stats(Tensor x, float eps) -> (Tensor, Tensor)",
        lambda x, eps: ops.instnorm_stats(x, eps),
        lambda x, eps: (_e((x.shape[2],), x), _e((x.shape[2],), x)))

_define("avgpool2(Tensor x) -> Tensor",
        lambda x: ops.avgpool2(x),
        lambda x: _e((x.shape[0] // 2, x.shape[1] // 2, x.shape[2]), x, torch.float16))

OP_NAMES = ("get_ray_bundle", "coarse_z", "coarse_z_rng", "positional_encoding", "field_fwd", "field_fwd_train", "field_bwd",
            "composite_fwd", "composite_bwd", "composite_fwd_rng", "composite_bwd_rng", "sample_pdf_merge",
            "sample_pdf_merge_rng", "sample_pdf", "frame_postprocess", "weighted_sample",
            "normal_map", "spade_conv", "spade_conv_t2", "instnorm_stats", "avgpool2")
