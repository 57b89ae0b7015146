"""Training path of the fused field: forward with tapes + hand-written activation-gradient chain (CUDA), weight
gradients as GEMMs over the tapes.

What the reference gets from autograd through WarpFieldMLP / HyperSheetMLP / NeRFMLP / grid_sample / the positional
encodings (ref: nerf/modules.py:254-295, :371-390, :444-462, nerf/models.py:301-365) is produced here by
  * `sahs_field_fwd_train`  -- the fused forward, additionally writing every layer's activated output (the fp16
                               "activation tape": tile-major images of the kernel's own operand chunks, TMA bulk
                               stores), the activation sign masks and the warped point;
  * `sahs_field_bwd`        -- the fused activation-gradient chain (tcgen05, transposed fp16 weights, scaled), writing every
                               layer's dY to the "gradient tape" (same layout) and scattering into the embedding-grid
                               gradient;
  * `sahs_field_wgrad`      -- dW_l = dY_l^T X_{l-1} and db_l for every layer: tcgen05 GEMMs fed by bulk loads of the
                               tape chunks, accumulators in TMEM across a tile range, red.global.add into fp32 grads;
  * the frame-constant input columns (driving 76 | pose code 36) were folded into biases in the forward, so their
    weight gradient is the rank-1 product db x cvec and d(driving) = W_const^T db.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional

import torch

from . import lib as L

_LAYOUT_KEYS = ["tx_e0", "e0_k", "tx_wh", "whh", "w_layers", "tx_e1", "e1_k", "tx_th", "th", "t_layers", "tx_feat",
                "tx_xtra", "tx_hh", "tx_total", "td_wh", "td_final", "td_th", "td_feat", "td_hh", "td_out", "td_total",
                "n_mask_layers", "e0_dim", "e1_dim", "xtra_dim", "wh", "hh", "w_skip", "t_skip", "ct_off", "ct_len",
                "use_w", "hd", "packed_t_bytes", "bwd_stages", "fc_total", "packed_train_bytes"]


def train_layout(cspec) -> Dict[str, int]:
    lib = L.load()
    out = (C.c_int32 * 40)()
    L.check(lib.sahs_train_layout(C.byref(cspec), out, 40), "train_layout")
    return dict(zip(_LAYOUT_KEYS, list(out)))


from .ops import tape_rows  # noqa: E402,F401  (tapes are tile-major chunk images: whole 128-point tiles)


class TrainState:
    """Per-level packed images for training, rebuilt when parameters change (same keying as packed_level)."""

    def __init__(self, model, level: str):
        lib = L.load()
        lvl = 0 if level == "coarse" else 1
        st = model.packed_level(level)                      # inference image also gives the pointer table + grid
        self.cspec, self.arr, self.params, self.grid = st["cspec"], st["arr"], st["params"], st["grid"]
        self.spec_ints = st["spec_ints"]
        self.key = st["key"]
        self.lay = train_layout(self.cspec)
        dev = st["packed"].device
        self.packed_train = torch.empty(self.lay["packed_train_bytes"], dtype=torch.uint8, device=dev)
        self.packed_t = torch.empty(self.lay["packed_t_bytes"], dtype=torch.uint8, device=dev)
        s = L.stream_ptr(dev)
        L.check(lib.sahs_pack_params_train(C.byref(self.cspec), lvl, self.arr, L.ptr(self.packed_train), s), "pack_train")
        L.check(lib.sahs_pack_params_bwd(C.byref(self.cspec), lvl, self.arr, L.ptr(self.packed_t), s), "pack_bwd")


def train_state(model, level: str) -> TrainState:
    cache = model.__dict__.setdefault("_train_states", {})
    st = model.packed_level(level)
    ts = cache.get(level)
    if ts is None or ts.key != st["key"]:
        ts = TrainState(model, level)
        cache[level] = ts
    return ts


def field_forward_tapes(model, level, ro, rd, z, driving_vec, pose_code):
    """Training forward of one level: raw[R,S,16] plus everything the backward needs (activation tape, sign masks,
    warped points).  Returns (raw, saved) -- `saved` is what FieldTrainFn keeps for backward; tests decode the tapes."""
    ts = train_state(model, level)
    lay = ts.lay
    lvl = 0 if level == "coarse" else 1
    ro, rd, z = L.f32c(ro.detach()), L.f32c(rd.detach()), L.f32c(z.detach())
    fc = model.frame_constants(level, driving_vec.detach(), pose_code.detach())
    raw, tape_x, masks, saves = torch.ops.sahs_b200.field_fwd_train(ts.spec_ints, lvl, ts.packed_train, fc, ts.grid, ro, rd,
                                                                    z, lay["tx_total"], lay["n_mask_layers"])
    return raw, dict(ts=ts, level=level, ro=ro, rd=rd, z=z, fc=fc, tape_x=tape_x, masks=masks, saves=saves)


def field_backward_tapes(saved, d_raw):
    """Activation-gradient chain of one level.  Returns (tape_d, grid_grad, scale): every layer's dY (fp16, multiplied
    by `scale`) as tile-major chunk images, the channel-last embedding-grid gradient (scaled) and the device scalar."""
    ts = saved["ts"]
    lvl = 0 if saved["level"] == "coarse" else 1
    d_raw = L.f32c(d_raw)
    # fp16 range management without a host sync: scale d_raw so that its largest entry is 16
    scale = (16.0 / d_raw.abs().amax().clamp_min(1e-30)).reshape(1).float()
    tape_d, grid_grad = torch.ops.sahs_b200.field_bwd(ts.spec_ints, lvl, ts.packed_t, saved["fc"], ts.grid, saved["ro"],
                                                      saved["rd"], saved["z"], d_raw, scale, saved["masks"],
                                                      saved["saves"], ts.lay["td_total"])
    return tape_d, grid_grad, scale


def decode_tape(tape: torch.Tensor, num_points: int) -> torch.Tensor:
    """Tile-major chunk images ([tiles][slots][128 rows x 64 columns, 128B-swizzled]) -> plain [num_points, columns]
    (diagnostics and tests; the kernels never need this form)."""
    cols = tape.shape[1]
    tiles, slots = tape.shape[0] // 128, cols // 64
    t = tape.reshape(tiles, slots, 16, 8, 8, 8)             # [tile][slot][row>>3][row&7][16-byte unit][8 halves]
    unit = torch.arange(8, device=tape.device)
    src = (unit[None, :] ^ unit[:, None])                    # logical unit u of row r sits at physical unit u ^ (r & 7)
    t = torch.gather(t, 4, src[None, None, None, :, :, None].expand(tiles, slots, 16, 8, 8, 8))
    return t.reshape(tiles, slots, 128, 64).permute(0, 2, 1, 3).reshape(tiles * 128, cols)[:num_points]


F16_MAX = 65504.0


def audit_fp16_range(model, level, ro, rd, z, driving_vec, pose_code) -> dict:
    """Range audit of the fp16 operands for a checkpoint.  The kernels convert every layer's activated output to fp16 with
    `cvt.rn.satfinite`: a value beyond +-65504 is clamped, not turned into inf, so a trained model whose hidden
    activations outgrow fp16 would render clipped values without any error.  This runs the tape-writing forward -- it
    stores exactly the 16-bit operands the tensor cores consume, layer by layer -- on the given rays and counts, per
    layer, the entries sitting at +-65504 and the largest magnitude.  `saturated == 0` and a comfortable `headroom`
    (65504 / max_abs) on a few representative frames is what qualifies a checkpoint for the fp16 build; otherwise use the
    bf16 build (`make BF16=1`, ops.operand_format() == 1).  Diagnostic: torch ops on the decoded tape, not a hot path."""
    with torch.no_grad():
        raw, saved = field_forward_tapes(model, level, ro, rd, z, driving_vec, pose_code)
        lay = saved["ts"].lay
        X = decode_tape(saved["tape_x"], z.numel()).abs()
        items = []
        if lay["use_w"]:
            for i in range(lay["w_layers"]):
                c = lay["tx_wh"] + i * lay["whh"]
                items.append((f"deform{i}", c, lay["wh"] + lay["hh"]))
        for i in range(lay["t_layers"]):
            items.append((f"trunk{i}", lay["tx_th"] + i * lay["th"], lay["th"]))
        items.append(("feat", lay["tx_feat"], lay["th"]))
        for i in range(4):
            items.append((f"head{i}", lay["tx_hh"] + i * 2 * lay["hd"], 2 * lay["hd"]))
        layers, total, worst = {}, 0, 0.0
        for name, c, n in items:
            blk = X[:, c:c + n]
            sat = int((blk >= F16_MAX).sum())                 # satfinite: the clamp value itself (never inf / nan)
            mx = float(blk.max()) if blk.numel() else 0.0
            layers[name] = {"max_abs": mx, "saturated": sat}
            total += sat
            worst = max(worst, mx)
    return {"saturated": total, "max_abs": worst, "headroom": F16_MAX / max(worst, 1e-30), "points": int(z.numel()),
            "raw_finite": bool(torch.isfinite(raw).all()), "layers": layers}


class FieldTrainFn(torch.autograd.Function):
    """raw = field(level, ro + rd z, rd; driving_vec, params).  Gradients: params of that level (incl. the shared
    deformation nets and the embedding grid) and driving_vec."""

    @staticmethod
    def forward(ctx, model, level, ro, rd, z, driving_vec, pose_code, *params):
        raw, saved = field_forward_tapes(model, level, ro, rd, z, driving_vec, pose_code)
        ctx.model, ctx.saved = model, saved
        ctx.cond = (driving_vec.detach(), pose_code.detach())
        return raw

    @staticmethod
    def backward(ctx, d_raw):
        model, saved = ctx.model, ctx.saved
        ts, level = saved["ts"], saved["level"]
        lay = ts.lay
        drv, pcode = ctx.cond
        R, S = saved["z"].shape
        P = R * S
        tape_d, grid_grad, scale = field_backward_tapes(saved, d_raw)
        cvec = torch.cat((drv.reshape(-1), pcode.reshape(-1)))
        views, d_cvec, flat, sizes = _weight_grads_kernel(model, level, ts, lay, saved["tape_x"], tape_d, cvec, P)
        inv = 1.0 / scale
        out = flat * inv                     # un-scale every parameter gradient of the level in one launch; a fresh
        d_cvec = d_cvec * inv                # tensor, because autograd may keep what it is handed
        grads, o = [], 0
        for g, n in zip(views, sizes):
            grads.append(None if g is None else out[o:o + n].view(g.shape))
            o += n
        grads[0] = (grid_grad * inv).permute(3, 0, 1, 2).unsqueeze(0).contiguous() if model.spec.use_grid else None
        d_driving = d_cvec[:76].reshape(drv.shape)
        ctx.saved = None                     # release the tapes
        return (None, None, None, None, None, d_driving, None) + tuple(grads)


def _weight_grads_kernel(model, level, ts, lay, tx, td, cvec, P):
    """Hand-written wgrad: one kernel accumulates dW/db of every layer; the folded frame-constant columns are the
    rank-1 products db x cvec (and d cvec = W_const^T db)."""
    lib = L.load()
    params = model._level_params(level)
    dev = tx.device
    # persistent flat fp32 gradient buffer of this level (one memset per step, stable pointers -> the wgrad kernel's
    # unit list is uploaded once) with one view per parameter
    cache = model.__dict__.setdefault("_wgrad_buffers", {})
    ent = cache.get(level)
    sig = tuple((None if p is None else (tuple(p.shape), p.device)) for p in params)
    if ent is None or ent["sig"] != sig:
        sizes = [0 if p is None else p.numel() for p in params]
        flat = torch.zeros(sum(sizes), dtype=torch.float32, device=dev)
        views, o = [], 0
        for p, n in zip(params, sizes):
            views.append(None if p is None else flat[o:o + n].view(p.shape))
            o += n
        arr = (C.c_void_p * len(views))()
        for i, g in enumerate(views):
            arr[i] = g.data_ptr() if g is not None else None
        ent = {"sig": sig, "flat": flat, "views": views, "arr": arr, "sizes": sizes,
               "ws": torch.empty(256 * 1024, dtype=torch.uint8, device=dev),
               "ws_token": C.c_uint64(0)}      # fingerprint of the unit list this workspace holds (0: unknown)
        cache[level] = ent
    flat, grads, arr, ws = ent["flat"], list(ent["views"]), ent["arr"], ent["ws"]
    flat.zero_()
    L.check(lib.sahs_field_wgrad(C.byref(ts.cspec), 0 if level == "coarse" else 1, arr, L.ptr(tx), L.ptr(td), P,
                                 L.ptr(ws), ws.numel(), C.byref(ent["ws_token"]), L.stream_ptr(dev)), "field_wgrad")
    d_cvec = torch.zeros(112, dtype=torch.float32, device=dev)
    s = model.spec
    e0d, e1d, wh, hh, th = lay["e0_dim"], lay["e1_dim"], lay["wh"], lay["hh"], lay["th"]

    def fold(w_param, w_grad, b_grad, col0, c_off, c_len):
        if c_len <= 0:
            return
        w_grad[:, col0:col0 + c_len] = torch.outer(b_grad, cvec[c_off:c_off + c_len])
        d_cvec[c_off:c_off + c_len] += w_param[:, col0:col0 + c_len].detach().float().t() @ b_grad

    k = 1
    if s.use_warp:
        for n in (wh, hh):
            for i in range(lay["w_layers"]):
                if i == 0:
                    fold(params[k], grads[k], grads[k + 1], e0d, 0, 112)
                elif i == lay["w_skip"]:
                    fold(params[k], grads[k], grads[k + 1], n + e0d, 0, 112)
                k += 2
            k += 2                                   # fc_final / fc_ambient
    for i in range(lay["t_layers"]):
        if i == 0:
            fold(params[k], grads[k], grads[k + 1], e1d, lay["ct_off"], lay["ct_len"])
        elif i == lay["t_skip"]:
            fold(params[k], grads[k], grads[k + 1], th + e1d, lay["ct_off"], lay["ct_len"])
        k += 2
    return grads, d_cvec, flat, ent["sizes"]


def field_train(model, level, ro, rd, z, driving_vec, pose_code):
    params = model._level_params(level)
    live = [p if p is not None else torch.empty(0, device=z.device) for p in params]
    return FieldTrainFn.apply(model, level, ro, rd, z, driving_vec, pose_code, *live)


class GraphedStep:
    """A whole training step as ONE CUDA graph launch.

    `step_fn()` must run a complete step from device-resident state only -- ray batch (`weighted_sample(...,
    seed_counter=...)`), `run_one_iter_of_nerf(mode="train")`, loss, `backward()`, `FlatAdam(capturable=True).step()`,
    `ops.counter_add(seed_counter)` -- and return the tensors to read back (e.g. the loss).  The step is run `warmup`
    times eagerly on a side stream (allocations, cuDNN plans, the wgrad unit-list upload), recorded once, and every
    call replays the recording: the ~1.3 ms of launch gaps of the eager step (36 of our launches plus ~100 small torch
    ops per step, measured in round 1) disappear.  Replays reuse the recording's memory, so outputs are overwritten by
    the next call."""

    def __init__(self, step_fn, warmup: int = 3):
        self._fn = step_fn
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(int(warmup), 1)):
                step_fn()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.outputs = step_fn()

    def __call__(self):
        self.graph.replay()
        return self.outputs
