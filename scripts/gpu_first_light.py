"""Diagnostic bring-up script for the GPU box: prints per-kernel / per-pass errors against the oracle instead of
asserting, so that one gpurun call yields as much information as possible.  Not part of the product path.

    python scripts/gpu_first_light.py basics
    python scripts/gpu_first_light.py field audio/person_2_auto
"""
import os
import sys
import time

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "tests"), os.path.join(REPO, "sahs-deformable-nerf_b200")):
    sys.path.insert(0, p)

import sahs_fixtures as FX  # noqa: E402
from oracle import sahs_oracle as O  # noqa: E402
import sahs_b200  # noqa: E402
from sahs_b200 import ops  # noqa: E402

GOLD = os.path.join(REPO, "tests", "golden")
dev = torch.device("cuda:0")
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def mx(a, b):
    return float((a.double().cpu() - b.double().cpu()).abs().max())


def basics():
    print("device:", torch.cuda.get_device_name(0), torch.cuda.get_device_capability(0))
    g = np.load(os.path.join(GOLD, "helpers.npz"))
    ro, rd = ops.get_ray_bundle(12, 20, list(g["intr"]), T(g["pose"]))
    print("ray_bundle: rd bit mismatches", int((rd.cpu() != torch.from_numpy(g["rd"])).sum()), "ro",
          int((ro.cpu() != torch.from_numpy(g["ro"])).sum()), "max", mx(rd, torch.from_numpy(g["rd"])))
    for L, inc in ((10, 1), (15, 1), (4, 1), (3, 0)):
        out = ops.positional_encoding(T(g["x"]), L, bool(inc))
        print(f"PE L={L} inc={inc}: max abs err", mx(out, torch.from_numpy(g[f"pe_L{L}_inc{inc}"])))
    opts = O.RenderOpts(num_coarse=64, near=0.483771014213562, far=1.083771014213562)
    z = ops.coarse_z(33, 64, opts.near, opts.far, False, dev)
    print("coarse_z det bit mismatches", int((z.cpu() != O.coarse_z(opts, 33)).sum()))
    tr = torch.rand(33, 64)
    opts.perturb = True
    z = ops.coarse_z(33, 64, opts.near, opts.far, False, dev, tr.to(dev))
    print("coarse_z perturb bit mismatches", int((z.cpu() != O.coarse_z(opts, 33, tr)).sum()))
    # sample_pdf
    g = np.load(os.path.join(GOLD, "sample_pdf_2048.npz"))
    R = g["weights"].shape[0]
    zrow = torch.from_numpy(g["z"]).expand(R, 64).contiguous()
    bins = 0.5 * (zrow[:, 1:] + zrow[:, :-1])
    s, inds = ops.sample_pdf_bins(bins.to(dev), T(g["weights"]), 64, None, return_inds=True)
    print("sample_pdf(bins): index mismatches", int((inds.cpu() != torch.from_numpy(g["ref_inds"].astype(np.int64))).sum()),
          "sample bit mismatches", int((s.cpu() != torch.from_numpy(g["ref_samples"])).sum()),
          "max abs", mx(s, torch.from_numpy(g["ref_samples"])))
    wfull = torch.zeros(R, 64)
    wfull[:, 1:-1] = torch.from_numpy(g["weights"])
    zs, zm, inds = ops.sample_pdf_merge(zrow.to(dev), wfull.to(dev), 64, None, return_inds=True)
    print("sample_pdf_merge: index mismatches", int((inds.cpu() != torch.from_numpy(g["ref_inds"].astype(np.int64))).sum()),
          "merged bit mismatches", int((zm.cpu() != torch.from_numpy(g["ref_z_merged"])).sum()))
    ss = ops.sample_pdf_bins(bins.to(dev), T(g["weights"]), 64, T(g["u_s"]))
    print("sample_pdf stochastic: bit mismatches", int((ss.cpu() != torch.from_numpy(g["ref_samples_s"])).sum()))
    # composite
    for name in ("composite_bg", "composite_nobg_white"):
        g = np.load(os.path.join(GOLD, name + ".npz"))
        bg = T(g["bg"]) if int(g["with_bg"]) else None
        out = ops.composite_fwd(T(g["raw"]), T(g["z"]), T(g["rd"]), None, bg, bool(int(g["with_bg"])), bool(int(g["white"])))
        errs = {n: mx(o, torch.from_numpy(g["ref_" + n])) for n, o in zip(["rgb", "disp", "acc", "weights", "depth"], out)}
        print(name, "fwd max abs:", {k: f"{v:.2e}" for k, v in errs.items()})
        # backward vs autograd through the oracle
        raw = torch.from_numpy(g["raw"]).clone().requires_grad_(True)
        zt, rdt = torch.from_numpy(g["z"]), torch.from_numpy(g["rd"])
        bgt = torch.from_numpy(g["bg"]) if int(g["with_bg"]) else None
        rin = raw
        if bgt is not None:
            rin = torch.cat((raw[:, :-1], torch.cat((bgt, raw[:, -1, -1:]), -1)[:, None]), 1)
        o = O.composite(rin, zt, rdt, None, bool(int(g["white"])), bgt)
        gen = torch.Generator().manual_seed(1)
        gs = [torch.randn(t.shape, generator=gen) for t in o]
        loss = sum((a * b).sum() for a, b in zip(o, gs))
        loss.backward()
        d_raw = ops.composite_bwd(T(g["raw"]), T(g["z"]), T(g["rd"]), None, bg, bool(int(g["with_bg"])),
                                  bool(int(g["white"])), gs[0].to(dev), gs[1].to(dev), gs[2].to(dev), gs[3].to(dev),
                                  gs[4].to(dev))
        ref = raw.grad
        den = ref.abs().max()
        print(name, "bwd max abs err", mx(d_raw, ref), "ref max", float(den), "rel", mx(d_raw, ref) / float(den))


def field(cfg_name):
    cfg = FX.load_cfg(cfg_name)
    ospec = O.spec_from_cfg(cfg)
    sd = FX.make_state_dict(ospec, seed=42, dense=True)
    gname = {"audio/person_2_auto": "field_audio", "expression/person_2": "field_expr2"}.get(cfg_name)
    model = getattr(sahs_b200.models, cfg.models.mask.type)(cfg)
    model.load_state_dict(sd, strict=True)
    model = model.to(dev)
    if gname:
        g = np.load(os.path.join(GOLD, gname + ".npz"))
        assert abs(FX.state_checksum(sd) - float(g["state_checksum"])) < 1e-6 * float(g["state_checksum"])
        xyz, dirs = torch.from_numpy(g["xyz"]), torch.from_numpy(g["dirs"])
        pose, drv = torch.from_numpy(g["pose"]), torch.from_numpy(g["driving_vec"])
    else:
        gen = torch.Generator().manual_seed(9)
        xyz = (torch.rand(512, 3, generator=gen) * 2 - 1) * 0.35
        dirs = torch.randn(512, 3, generator=gen) * 0.3 + torch.tensor([0, 0, -1.0])
        fr = FX.make_frame_inputs(ospec, 8, 8, seed=1)
        pose, drv = fr["pose"], O.driving_vector(sd, ospec, fr["driving"])
    n = xyz.shape[0]
    pcode = O.pose_code(pose)
    print("pose code (ours vs oracle):", mx(model.pose_code(pose.to(dev)), pcode))
    z0 = torch.zeros(n, 1, device=dev)
    for level in ("coarse", "fine"):
        ref, inter = O.field_forward(sd, ospec, level, xyz, dirs, drv, pose, return_intermediates=True)
        fc = model.frame_constants(level, drv.to(dev), pcode.to(dev))
        torch.cuda.synchronize()
        print(f"[{level}] packed bytes", model.packed_level(level)["packed"].numel(), "fc floats", fc.numel())
        dbg = torch.zeros(128, 256, device=dev)

        def run(dbg_pass):
            dbg.zero_()
            raw = model.field(level, xyz.to(dev), dirs.to(dev), z0, drv.to(dev), pcode.to(dev), frame_const=fc,
                              debug=dbg, debug_pass=dbg_pass)
            torch.cuda.synchronize()
            return raw.reshape(n, 16).cpu(), dbg.cpu().clone()

        t0 = time.time()
        raw, _ = run(-1)
        print(f"[{level}] first launch ok in {time.time() - t0:.3f}s; status", ops.field_status())
        err = (raw - ref).abs()
        print(f"[{level}] raw max abs err: rgb {float(err[:, :3].max()):.3e} seg {float(err[:, 3:15].max()):.3e} "
              f"sigma {float(err[:, 15].max()):.3e} (|ref| max rgb {float(ref[:, :3].abs().max()):.2f} "
              f"seg {float(ref[:, 3:15].abs().max()):.2f} sigma {float(ref[:, 15].abs().max()):.2f}) "
              f"nan {int(torch.isnan(raw).sum())}")
        if level == "coarse":
            s = ospec
            if s.use_warp:
                for i in range(s.warp_layers):
                    _, d = run(i)
                    want = torch.cat((inter[f"warp{i}"][:128], inter[f"hyper{i}"][:128]), 1)
                    got = d[:, :want.shape[1]]
                    print(f"   pass WARP({i}): max err {mx(got, want):.3e} (max |ref| {float(want.abs().max()):.2f})")
            _, d = run(16)
            print(f"   MAPPED: xyz err {mx(d[:, :3], inter['mapped'][:128]):.3e}",
                  (f"amb err {mx(d[:, 3:3 + s.amb_dim], inter['amb'][:128]):.3e}" if s.use_ambient else ""),
                  f"emb err {mx(d[:, 8:40], inter['emb'][:128]):.3e} (max |emb| {float(inter['emb'].abs().max()):.2f})")
            for i in range(s.trunk_layers):
                _, d = run(32 + i)
                print(f"   pass TRUNK({i}): max err {mx(d, inter[f'trunk{i}'][:128]):.3e} (max |ref| {float(inter[f'trunk{i}'].abs().max()):.2f})")
            _, d = run(32 + s.trunk_layers)
            print(f"   pass FEAT: max err {mx(d, inter['feat'][:128]):.3e} (max |ref| {float(inter['feat'].abs().max()):.2f})")
            for i in range(4):
                _, d = run(64 + i)
                want = torch.cat((inter[f"dir{i}"][:128], inter[f"seg{i}"][:128]), 1)
                print(f"   pass HEAD({i}): max err {mx(d, want):.3e} (max |ref| {float(want.abs().max()):.2f})")
    # timing at scale
    R, S = 8192, 128
    ro = torch.zeros(R, 3, device=dev); ro[:, 2] = 0.78
    rd = torch.randn(R, 3, device=dev) * 0.1; rd[:, 2] = -1
    z = torch.linspace(0.48, 1.08, S, device=dev).expand(R, S).contiguous()
    fc = model.frame_constants("fine", drv.to(dev), pcode.to(dev))
    for _ in range(2):
        model.field("fine", ro, rd, z, drv.to(dev), pcode.to(dev), frame_const=fc)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        model.field("fine", ro, rd, z, drv.to(dev), pcode.to(dev), frame_const=fc)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"field fine {R}x{S} points: {ms:.3f} ms -> {R * S / ms / 1e3:.1f} Mpts/s, {R / ms * 1e3:.0f} rays/s-equivalent(fine only)")


if __name__ == "__main__":
    what = sys.argv[1]
    if what == "basics":
        basics()
    else:
        field(sys.argv[2])
