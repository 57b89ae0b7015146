"""CPU tests of the host logic of the fused field path (no GPU, no compute calls into the library):
the stage plan, pack descriptors and frame-constant folding produced by libsahs_b200's host code are
interpreted in numpy (fp32) following the kernel's pass protocol and must reproduce the oracle's field
forward.  This pins every row/column offset of the weight packing before any GPU time is spent."""
import ctypes as C

import numpy as np
import pytest
import torch

import sahs_fixtures as FX
from oracle import sahs_oracle as O

ST_WAIT_A, ST_COMMIT, ST_FRESH, ST_F16, ST_WIDE = 1, 2, 4, 8, 16


def _ordered_param_names(spec, level):
    names = ["spatial_embeddings" if spec.use_grid else None]
    if spec.use_warp:
        for i in range(spec.warp_layers):
            names += [f"warp_field_mlp.layers_xyz.{i}.weight", f"warp_field_mlp.layers_xyz.{i}.bias"]
        names += ["warp_field_mlp.fc_final.weight", "warp_field_mlp.fc_final.bias"]
    if spec.use_ambient:
        for i in range(spec.hyper_layers):
            names += [f"hyper_sheep_mlp.layers_ambient.{i}.weight", f"hyper_sheep_mlp.layers_ambient.{i}.bias"]
        names += ["hyper_sheep_mlp.fc_ambient.weight", "hyper_sheep_mlp.fc_ambient.bias"]
    p = f"nerf_mlps.{level}."
    for i in range(spec.trunk_layers):
        names += [p + f"layers_xyz.{i}.weight", p + f"layers_xyz.{i}.bias"]
    names += [p + "fc_feat.weight", p + "fc_feat.bias", p + "fc_alpha.weight", p + "fc_alpha.bias"]
    for i in range(4):
        names += [p + f"layers_dir.{i}.weight", p + f"layers_dir.{i}.bias"]
    names += [p + "fc_rgb.weight", p + "fc_rgb.bias"]
    for i in range(4):
        names += [p + f"layers_seg.{i}.weight", p + f"layers_seg.{i}.bias"]
    names += [p + "fc_seg.weight", p + "fc_seg.bias"]
    return names


def _plan(spec_model):
    from sahs_b200 import lib as L
    lib = L.load()
    cs = spec_model.to_c()
    n = lib.sahs_param_count(C.byref(cs))
    stages = (C.c_int32 * (14 * 160))()
    folds = (C.c_int32 * (8 * 64))()
    copies = (C.c_int32 * (3 * 8))()
    dims = (C.c_int32 * 32)()
    L.check(lib.sahs_debug_plan(C.byref(cs), n, stages, 160, folds, 64, copies, 8, dims), "debug_plan")
    dims = list(dims)
    ns, nf, nc = dims[0], dims[1], dims[2]
    keys = ["num_stages", "num_fold", "num_copy", "total_bytes", "fc_total", "e0_dim", "e0_k", "e1_dim", "e1_k",
            "e0_resident", "e0_chunk_base", "whh", "off_wbias", "off_wfinal", "off_tbias", "off_featb", "off_alpha",
            "off_hbias", "off_outb", "xtra_dim", "w_split"]
    return (np.array(stages[:14 * ns]).reshape(ns, 14), np.array(folds[:8 * nf]).reshape(nf, 8),
            np.array(copies[:3 * nc]).reshape(nc, 3), dict(zip(keys, dims)), n)


def _emulate(spec_model, ospec, sd, level, xyz, dirs, driving_vec, pose):
    stages, folds, copies, dm, nparams = _plan(spec_model)
    names = _ordered_param_names(ospec, level)
    assert len(names) == nparams
    P = [None if n is None else sd[n].numpy() for n in names]
    get = lambda pid: P[pid - 1]
    cvec = np.concatenate([driving_vec.numpy(), O.pose_code(pose).numpy()]).astype(np.float32)
    fc = np.zeros(dm["fc_total"], np.float32)
    for b_id, w_id, ld, col0, ncols, c_off, n, dst in folds:
        v = get(b_id).reshape(-1)[:n].copy()
        if w_id:
            W = get(w_id)
            assert W.shape[1] == ld and W.shape[0] == n
            v = v + W[:, col0:col0 + ncols] @ cvec[c_off:c_off + ncols]
        fc[dst:dst + n] = v
    for pid, count, dst in copies:
        fc[dst:dst + count] = get(pid).reshape(-1)[:count]
    # split into passes
    passes, cur = [], []
    for st in stages:
        if st[2] & ST_WAIT_A and cur:
            raise AssertionError("WAIT_A inside an open pass")
        cur.append(st)
        if st[2] & ST_COMMIT:
            passes.append(cur)
            cur = []
    assert not cur
    it = iter(passes)
    npts = xyz.shape[0]
    X = np.zeros((npts, 256), np.float32)
    D = np.full((npts, 256), np.nan, np.float32)

    def run_pass():
        written = set()
        for n, ks, flags, a_chunk, d_col, dst_off, pid, r0, c0, dr0, nr, ncol, a_chunk2, lo in next(it):
            W = get(pid)
            img = np.zeros((n, 64), np.float32)
            img[dr0:dr0 + nr, :ncol] = W[r0:r0 + nr, c0:c0 + ncol]
            if dm["w_split"] and (flags & ST_F16):       # emulate the fp16 hi / lo weight planes exactly
                hi = img.astype(np.float16).astype(np.float32)
                img = (img - hi).astype(np.float16).astype(np.float32) if lo else hi
            k = 16 * ks
            k0 = 16 * (a_chunk >> 8)                 # ST_WIDE stages multiply half a chunk (a_k16 rides in bits 8..)
            a_chunk &= 0xFF
            contrib = X[:, a_chunk * 64 + k0:a_chunk * 64 + k0 + k] @ img[:, :k].T
            if a_chunk2 != 255:
                contrib = contrib + X[:, a_chunk2 * 64 + k0:a_chunk2 * 64 + k0 + k] @ img[:, :k].T
            if flags & ST_FRESH:
                D[:, d_col:d_col + n] = contrib
            else:
                D[:, d_col:d_col + n] += contrib
        return D

    pe = lambda t, L, inc: O.positional_encoding(torch.from_numpy(np.ascontiguousarray(t)), L, inc).numpy()
    s = ospec
    pts = xyz.numpy()
    mapped, amb = pts.copy(), None
    f16 = lambda a: a.astype(np.float16).astype(np.float32)
    if s.use_warp and dm["w_split"]:
        e0 = pe(pts, s.xyz_L, True)
        o = dm["off_wfinal"]
        wh, hh = s.warp_hidden, s.hyper_hidden
        wf = fc[o:o + 3 * wh].reshape(3, wh); bf = fc[o + 3 * wh:o + 3 * wh + 3]
        o2 = o + 3 * wh + 4
        wa = fc[o2:o2 + s.amb_dim * hh].reshape(s.amb_dim, hh); ba = fc[o2 + s.amb_dim * hh:o2 + s.amb_dim * hh + s.amb_dim]

        def write_split(vals, hi0, lo0, width):
            X[:, hi0 * 64:hi0 * 64 + width] = 0
            X[:, lo0 * 64:lo0 * 64 + width] = 0
            hi = f16(vals)
            X[:, hi0 * 64:hi0 * 64 + vals.shape[1]] = hi
            X[:, lo0 * 64:lo0 * 64 + vals.shape[1]] = f16(vals - hi)

        for net, n, boff in ((0, wh, 0), (1, hh, wh)):
            write_split(e0, 0, 2, 128)
            for i in range(s.warp_layers):
                if i == s.warp_skip:
                    run_pass()
                    write_split(e0, 0, 2, 128)
                run_pass()
                b0 = dm["off_wbias"] + i * dm["whh"] + boff
                h = np.maximum(D[:, :n] + fc[b0:b0 + n], 0)
                if i < s.warp_layers - 1:
                    write_split(h, 0, n // 64, n)
            if net == 0:
                mapped = pts + np.tanh(h @ wf.T + bf)
            else:
                amb = h @ wa.T + ba
    elif s.use_warp:
        e0 = pe(pts, s.xyz_L, True)
        cb = dm["e0_chunk_base"]
        X[:, cb * 64:cb * 64 + e0.shape[1]] = e0
        X[:, cb * 64 + e0.shape[1]:cb * 64 + dm["e0_k"]] = 0
        for i in range(s.warp_layers):
            if i == s.warp_skip and not dm["e0_resident"]:
                run_pass()
                X[:, :dm["e0_k"]] = 0
                X[:, :e0.shape[1]] = e0
            run_pass()
            h = np.maximum(D[:, :dm["whh"]] + fc[dm["off_wbias"] + i * dm["whh"]: dm["off_wbias"] + (i + 1) * dm["whh"]], 0)
            if i < s.warp_layers - 1:
                X[:, :dm["whh"]] = h
        o = dm["off_wfinal"]
        wh, hh = s.warp_hidden, s.hyper_hidden
        wf = fc[o:o + 3 * wh].reshape(3, wh); bf = fc[o + 3 * wh:o + 3 * wh + 3]
        o2 = o + 3 * wh + 4
        wa = fc[o2:o2 + s.amb_dim * hh].reshape(s.amb_dim, hh); ba = fc[o2 + s.amb_dim * hh:o2 + s.amb_dim * hh + s.amb_dim]
        mapped = pts + np.tanh(h[:, :wh] @ wf.T + bf)
        amb = h[:, wh:] @ wa.T + ba
    emb = O.grid_sample_trilinear(sd["spatial_embeddings"], torch.from_numpy(mapped)).numpy()
    e1 = pe(mapped, s.xyz_L, True)
    if amb is not None:
        e1 = np.concatenate([e1, pe(amb.astype(np.float32), s.amb_L, s.amb_inc)], 1)
    assert e1.shape[1] == dm["e1_dim"]

    def write_e1():
        X[:, :128] = 0
        X[:, :e1.shape[1]] = e1

    write_e1()
    lrelu = lambda v: np.maximum(v, 0.01 * v)
    for i in range(s.trunk_layers):
        if i == s.trunk_skip:
            run_pass()
            write_e1()
        run_pass()
        X[:, :256] = lrelu(D + fc[dm["off_tbias"] + i * 256: dm["off_tbias"] + (i + 1) * 256])
    run_pass()
    feat = D + fc[dm["off_featb"]:dm["off_featb"] + 256]
    sigma = feat @ fc[dm["off_alpha"]:dm["off_alpha"] + 256] + fc[dm["off_alpha"] + 256]
    X[:, :256] = feat
    run_pass()
    X[:, :64] = 0
    ed = pe(dirs.numpy(), s.dir_L, True)
    X[:, :ed.shape[1]] = ed
    X[:, ed.shape[1]:ed.shape[1] + 32] = emb
    for i in range(4):
        run_pass()
        X[:, :256] = lrelu(D + fc[dm["off_hbias"] + i * 256: dm["off_hbias"] + (i + 1) * 256])
    run_pass()
    out = D[:, :15] + fc[dm["off_outb"]:dm["off_outb"] + 15]
    assert next(it, None) is None, "plan has more passes than the worker protocol consumes"
    return np.concatenate([out, sigma[:, None]], 1), dm, stages


@pytest.mark.parametrize("cfg_name", ["audio/person_2_auto", "expression/person_2", "expression/person_1"])
def test_plan_interpreter_matches_oracle(cfg_name):
    from sahs_b200.models import ModelSpec
    cfg = FX.load_cfg(cfg_name)
    ospec = O.spec_from_cfg(cfg)
    mspec = ModelSpec.from_cfg(cfg)
    sd = FX.make_state_dict(ospec, seed=7, dense=True)
    g = torch.Generator().manual_seed(3)
    n = 96
    xyz = (torch.rand(n, 3, generator=g) * 2 - 1) * 0.4
    dirs = torch.randn(n, 3, generator=g) * 0.3 + torch.tensor([0.0, 0.0, -1.0])
    fr = FX.make_frame_inputs(ospec, 4, 4, seed=5)
    drv = O.driving_vector(sd, ospec, fr["driving"])
    for level in ("coarse", "fine"):
        ref = O.field_forward(sd, ospec, level, xyz, dirs, drv, fr["pose"]).numpy()
        got, dm, stages = _emulate(mspec, ospec, sd, level, xyz, dirs, drv, fr["pose"])
        scale = max(1.0, np.abs(ref).max())
        # colour / semantic logits to fp32 noise level; the density logit is x400 in the dense fixture and, with 15
        # octaves, magnifies the ~1e-7 difference of the warped point between two fp32 evaluation orders
        assert np.abs(got[:, :15] - ref[:, :15]).max() < 2e-4, (cfg_name, level)
        assert np.abs(got[:, 15] - ref[:, 15]).max() < (2e-4 if ospec.xyz_L <= 10 else 1e-3) * scale, (cfg_name, level)
        # stage images are laid out back to back in consumption order
        offs = stages[:, 5]
        wide = (stages[:, 2] & ST_WIDE) != 0                          # ST_WIDE: [256 x 32] images, 64-byte rows
        nbytes = stages[:, 0] * np.where(wide, 64, 128)
        assert offs[0] == 0 and np.all(np.diff(offs) == nbytes[:-1])
        assert dm["total_bytes"] == offs[-1] + nbytes[-1]
        assert np.all(stages[:, 0] % 16 == 0) and np.all(stages[~wide, 0] <= 128)      # UMMA M=128 needs N%16==0
        assert np.any(wide) and np.all(stages[wide, 0] == 256) and np.all(stages[wide, 1] <= 2)
        assert np.all(nbytes <= 16384)                                  # a stage fits one ring slot
        assert np.all((stages[:, 1] >= 1) & (stages[:, 1] <= 4))


def test_library_exports_every_declared_symbol():
    """The C-ABI library loads and exports each function include/sahs_b200.h declares."""
    import os
    import re
    from sahs_b200 import lib as L
    hdr = open(os.path.join(FX.REPO, "include", "sahs_b200.h")).read()
    declared = set(re.findall(r"\b(sahs_[a-z0-9_]+)\s*\(", hdr))
    lib = L.load()
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert declared == set(L.EXPORTS), declared ^ set(L.EXPORTS)
    assert lib.sahs_abi_version() == 2


def test_exp_lr_matches_training_script_schedule():
    """lr_new = lr * decay_factor ** (i / (lr_decay * 1000)) (ref: train_stage_rays_auto.py:503-507), shipped values."""
    from sahs_b200 import exp_lr
    cfg = FX.load_cfg("audio/person_2_auto")
    lr0, f, steps = float(cfg.optimizer.lr), float(cfg.scheduler.lr_decay_factor), float(cfg.scheduler.lr_decay) * 1000
    assert exp_lr(lr0, f, steps, 0) == lr0
    assert abs(exp_lr(lr0, f, steps, int(steps)) - lr0 * f) < 1e-15
    assert abs(exp_lr(lr0, f, steps, 125000) - lr0 * f ** 0.5) < 1e-15
    vals = [exp_lr(lr0, f, steps, i) for i in range(0, 400000, 50000)]
    assert all(a > b for a, b in zip(vals, vals[1:]))
