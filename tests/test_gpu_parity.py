"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI (ctypes) behind the reference's Python
signatures, against the golden vectors frozen from the live reference and against the CPU oracle on the same
seeded inputs.  Bars (BASELINE.json north_star): bit-exact for index/integer work (sample_pdf indices, merged
depths, ray bundle, coarse depths); rendered rgb/depth max-abs <= 1e-2 and PSNR >= 50 dB (bf16 tensor-core MLP)."""
import math
import os

import numpy as np
import pytest
import torch

import sahs_fixtures as FX
from oracle import sahs_oracle as O

pytestmark = pytest.mark.gpu

GOLD = os.path.join(FX.REPO, "tests", "golden")
load = lambda n: np.load(os.path.join(GOLD, n + ".npz"))
DEV = "cuda:0"
G = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
C = lambda a: torch.from_numpy(np.ascontiguousarray(a))


def _report(line: str) -> None:
    """Measured numbers go to gpurun_out/test_report.txt (when run through gpurun)."""
    print(line)
    d = os.path.join(FX.REPO, "gpurun_out")
    if os.path.isdir(d):
        with open(os.path.join(d, "test_report.txt"), "a") as f:
            f.write(line + "\n")


def maxabs(a, b):
    return float((a.detach().double().cpu() - b.detach().double().cpu()).abs().max())


def psnr(a, b):
    mse = float(((a.detach().double().cpu() - b.detach().double().cpu()) ** 2).mean())
    return 99.0 if mse == 0 else -10.0 * math.log10(mse)


@pytest.fixture(scope="module")
def sahs():
    import sahs_b200
    from sahs_b200 import lib
    lib.load()                       # fails loudly if the CUDA library is missing: there is no fallback
    return sahs_b200


def _model(sahs, cfg_name, trained_like=False):
    cfg = FX.load_cfg(cfg_name)
    spec = O.spec_from_cfg(cfg)
    sd = FX.make_state_dict(spec, seed=42, dense=True, trained_like=trained_like)
    model = getattr(sahs.models, cfg.models.mask.type)(cfg)
    model.load_state_dict(sd, strict=True)
    return cfg, spec, sd, model.to(DEV)


# ------------------------------------------------------------------------------------------------------
# (1) ray generation, depths, positional encoding
# ------------------------------------------------------------------------------------------------------
def test_ray_bundle_bit_exact(sahs):
    g = load("helpers")
    ro, rd = sahs.get_ray_bundle(12, 20, g["intr"], G(g["pose"]))
    assert torch.equal(rd.cpu(), C(g["rd"])) and torch.equal(ro.cpu(), C(g["ro"]))
    # scalar focal form + a frame-sized bundle against the oracle
    pose = FX.make_pose(4, 0.78, 10.0)
    ro, rd = sahs.get_ray_bundle(512, 512, [1200.0, 1200.0, 0.5, 0.5], pose.to(DEV))
    ro_o, rd_o = O.get_ray_bundle(512, 512, [1200.0, 1200.0, 0.5, 0.5], pose)
    assert torch.equal(rd.cpu(), rd_o) and torch.equal(ro.cpu(), ro_o.contiguous())


def test_positional_encoding(sahs):
    g = load("helpers")
    for L, inc in ((10, 1), (15, 1), (4, 1), (3, 0)):
        out = sahs.positional_encoding(G(g["x"]), L, bool(inc))
        assert maxabs(out, C(g[f"pe_L{L}_inc{inc}"])) <= 2e-6        # sincosf (<= 2 ulp) vs glibc
    assert sahs.positional_encoding(G(g["x"]), 0, True).shape == (257, 3)
    empty = sahs.positional_encoding(torch.zeros(0, 3, device=DEV), 4, True)
    assert empty.shape == (0, 27)


@pytest.mark.parametrize("lindisp", [False, True])
def test_coarse_depths_bit_exact(sahs, lindisp):
    from sahs_b200 import ops
    opts = O.RenderOpts(num_coarse=64, near=0.483771014213562, far=1.083771014213562, lindisp=lindisp)
    z = ops.coarse_z(37, 64, opts.near, opts.far, lindisp, DEV)
    assert torch.equal(z.cpu(), O.coarse_z(opts, 37))
    tr = torch.rand(37, 64, generator=torch.Generator().manual_seed(3))
    opts.perturb = True
    z = ops.coarse_z(37, 64, opts.near, opts.far, lindisp, DEV, tr.to(DEV))
    assert torch.equal(z.cpu(), O.coarse_z(opts, 37, tr))


def test_frame_postprocess_bit_exact(sahs):
    g = load("postprocess")
    rgb, label, col = sahs.frame_postprocess(G(g["map"]))
    assert torch.equal(rgb.cpu(), C(g["ref_rgb"])) and torch.equal(label.cpu(), C(g["ref_label"]))
    assert torch.equal(col.cpu(), C(g["ref_color"]))
    big = torch.rand(512, 512, 15, device=DEV)
    r2, l2, c2 = sahs.frame_postprocess(big)
    ro, lo, co = O.frame_postprocess(big.cpu())
    assert torch.equal(r2.cpu(), ro) and torch.equal(l2.cpu(), lo) and torch.equal(c2.cpu(), co)


# ------------------------------------------------------------------------------------------------------
# (4) sample_pdf + merge: bit-exact indices when fed the reference's weights
# ------------------------------------------------------------------------------------------------------
def test_sample_pdf_bit_exact_vs_reference_golden(sahs):
    from sahs_b200 import ops
    g = load("sample_pdf_2048")
    R = g["weights"].shape[0]
    z = C(g["z"]).expand(R, 64).contiguous()
    bins = 0.5 * (z[:, 1:] + z[:, :-1])
    s, inds = ops.sample_pdf_bins(bins.to(DEV), G(g["weights"]), 64, None, return_inds=True)
    assert torch.equal(inds.cpu(), C(g["ref_inds"].astype(np.int64)))
    assert torch.equal(s.cpu(), C(g["ref_samples"]))
    # public reference signature
    assert torch.equal(sahs.sample_pdf_2(bins.to(DEV), G(g["weights"]), 64, det=True).cpu(), C(g["ref_samples"]))
    # pipeline form: depths + full compositing weights -> samples + merged (sorted) depths
    wfull = torch.zeros(R, 64)
    wfull[:, 1:-1] = C(g["weights"])
    zs, zm, inds2 = ops.sample_pdf_merge(z.to(DEV), wfull.to(DEV), 64, None, return_inds=True)
    assert torch.equal(inds2.cpu(), C(g["ref_inds"].astype(np.int64)))
    assert torch.equal(zs.cpu(), C(g["ref_samples"])) and torch.equal(zm.cpu(), C(g["ref_z_merged"]))
    # stochastic u (unsorted samples exercise the bitonic merge)
    ss = ops.sample_pdf_bins(bins.to(DEV), G(g["weights"]), 64, G(g["u_s"]))
    assert torch.equal(ss.cpu(), C(g["ref_samples_s"]))
    zs, zm = ops.sample_pdf_merge(z.to(DEV), wfull.to(DEV), 64, G(g["u_s"]))
    want, _ = torch.sort(torch.cat((z, C(g["ref_samples_s"])), -1), -1)
    assert torch.equal(zm.cpu(), want)


def test_sample_pdf_properties_full_frame(sahs):
    """Size-independent properties at the BASELINE ray count (262,144 rays): merged depths are sorted, are a
    permutation of (z, samples), samples stay inside [bin_0, bin_last]; oracle equality on a strided subset."""
    from sahs_b200 import ops
    R = 262144
    gen = torch.Generator().manual_seed(5)
    opts = O.RenderOpts(num_coarse=64, near=0.4838, far=1.0838)
    z = ops.coarse_z(R, 64, opts.near, opts.far, False, DEV)
    w = (torch.rand(R, 64, generator=gen) ** 8).to(DEV)
    zs, zm, inds = ops.sample_pdf_merge(z, w, 64, None, return_inds=True)
    assert bool((zm[:, 1:] >= zm[:, :-1]).all())
    both, _ = torch.sort(torch.cat((z, zs), -1), -1)
    assert torch.equal(both, zm)
    mids = 0.5 * (z[:, 1:] + z[:, :-1])
    assert bool((zs >= mids[:, :1]).all()) and bool((zs <= mids[:, -1:]).all())
    assert int(inds.min()) >= 1 and int(inds.max()) <= 63
    sub = slice(0, R, 4099)
    s_o, i_o = O.sample_pdf(mids[sub].cpu(), w[sub, 1:-1].cpu(), 64, det=True, return_inds=True)
    assert torch.equal(inds[sub].cpu(), i_o) and torch.equal(zs[sub].cpu(), s_o)


@pytest.mark.parametrize("S,NF,shuffle", [(64, 48, False), (64, 64, True), (40, 17, False), (128, 64, False)])
def test_sample_pdf_merge_paths(sahs, S, NF, shuffle):
    """The merge has two code paths: 64 new samples into ascending coarse depths (register sort + rank merge) and
    everything else (bitonic network in shared memory, also taken when the coarse depths are not ascending).
    Both must equal sort(cat(z, samples)) bitwise, and the samples must equal the oracle's."""
    from sahs_b200 import ops
    R = 1531
    gen = torch.Generator().manual_seed(S * 1000 + NF)
    z, _ = torch.sort(0.48 + 0.6 * torch.rand(R, S, generator=gen), -1)
    if shuffle:
        z = z[:, torch.randperm(S, generator=gen)].contiguous()
    w = torch.rand(R, S, generator=gen) ** 6
    u = torch.rand(R, NF, generator=gen)
    zs, zm, inds = ops.sample_pdf_merge(z.to(DEV), w.to(DEV), NF, u.to(DEV), return_inds=True)
    mids = 0.5 * (z[:, 1:] + z[:, :-1])
    if not shuffle:     # searchsorted semantics need ascending bins; the shuffled case checks the merge only
        s_o, i_o = O.sample_pdf(mids, w[:, 1:-1], NF, det=False, u=u, return_inds=True)
        assert torch.equal(inds.cpu(), i_o) and torch.equal(zs.cpu(), s_o)
    want, _ = torch.sort(torch.cat((z, zs.cpu()), -1), -1)
    assert torch.equal(zm.cpu(), want)


@pytest.mark.parametrize("mode", ["det", "stochastic", "unsorted_coarse", "ties", "flat"])
def test_sample_pdf_64x64_kernel_equals_general_kernel(sahs, mode, monkeypatch):
    """The renderer's shape (64 coarse + 64 new samples) has its own kernel (search-free merge).  It must return the
    general kernel's bits -- samples, indices, merged depths -- for deterministic and stochastic u, for coarse depths
    that are not ascending (fallback sort), for repeated depths / samples, and for flat weights; plus the oracle."""
    from sahs_b200 import ops
    R = 4097
    gen = torch.Generator().manual_seed(11)
    z, _ = torch.sort(0.48 + 0.6 * torch.rand(R, 64, generator=gen), -1)
    w = torch.rand(R, 64, generator=gen) ** 8
    u = None
    if mode == "stochastic":
        u = torch.rand(R, 64, generator=gen).to(DEV)
    elif mode == "unsorted_coarse":
        z = z[:, torch.randperm(64, generator=gen)].contiguous()
    elif mode == "ties":
        z = (z * 16).round() / 16                      # many equal depths
        w = (w > 0.5).float()                          # zero-probability bins: runs of identical samples
    elif mode == "flat":
        w = torch.zeros(R, 64)
    z, w = z.to(DEV), w.to(DEV)
    fast = ops.sample_pdf_merge(z, w, 64, u, return_inds=True)
    monkeypatch.setenv("SAHS_SAMPLE_PDF_GENERIC", "1")
    general = ops.sample_pdf_merge(z, w, 64, u, return_inds=True)
    monkeypatch.delenv("SAHS_SAMPLE_PDF_GENERIC")
    for a, b in zip(fast, general):
        assert torch.equal(a, b)
    want, _ = torch.sort(torch.cat((z, fast[0]), -1), -1)
    assert torch.equal(fast[1], want)
    if mode in ("det", "stochastic", "flat"):
        mids = 0.5 * (z[:, 1:] + z[:, :-1]).cpu()
        s_o, i_o = O.sample_pdf(mids, w[:, 1:-1].cpu(), 64, det=u is None, u=None if u is None else u.cpu(), return_inds=True)
        assert torch.equal(fast[2].cpu(), i_o) and torch.equal(fast[0].cpu(), s_o)
    # a view that is only 4-byte aligned takes the general kernel (the fast one loads float2): same bits
    buf = torch.zeros(R * 64 + 1, device=DEV)
    buf[1:] = z.reshape(-1)
    odd = ops.sample_pdf_merge(buf[1:].view(R, 64), w, 64, u, return_inds=True)
    for a, b in zip(fast, odd):
        assert torch.equal(a, b)


# ------------------------------------------------------------------------------------------------------
# (3) compositing
# ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["composite_bg", "composite_nobg_white"])
def test_composite_forward_golden(sahs, name):
    from sahs_b200 import ops
    g = load(name)
    with_bg, white = bool(int(g["with_bg"])), bool(int(g["white"]))
    bg = G(g["bg"]) if with_bg else None
    out = ops.composite_fwd(G(g["raw"]), G(g["z"]), G(g["rd"]), None, bg, with_bg, white)
    for n, o in zip(["rgb", "disp", "acc", "weights", "depth"], out):
        ref = C(g["ref_" + n])
        if n == "disp":
            rel = float(((o.cpu() - ref).abs() / ref.abs()).max())
            assert rel <= 1e-4, rel
        else:
            assert maxabs(o, ref) <= 2e-5, (n, maxabs(o, ref))        # fp32, different exp/scan order only
    # reference signature: the caller has already overwritten the last sample with the background
    raw_in = G(g["raw"]).clone()
    if with_bg:
        raw_in[:, -1, :-1] = bg
    rgb, disp, acc, w, depth = sahs.volume_render_radiance_field(raw_in, G(g["z"]), G(g["rd"]), 0.0, white, bg)
    assert maxabs(rgb, C(g["ref_rgb"])) <= 2e-5 and maxabs(w, C(g["ref_weights"])) <= 2e-5


@pytest.mark.parametrize("name", ["composite_bg", "composite_nobg_white"])
def test_composite_backward_vs_oracle_autograd(sahs, name):
    from sahs_b200 import ops
    g = load(name)
    with_bg, white = bool(int(g["with_bg"])), bool(int(g["white"]))
    raw = C(g["raw"]).clone().requires_grad_(True)
    z, rd = C(g["z"]), C(g["rd"])
    bg = C(g["bg"]) if with_bg else None
    rin = raw
    if with_bg:      # out-of-place version of raw[:, -1, :-1] = bg
        rin = torch.cat((raw[:, :-1], torch.cat((bg, raw[:, -1, -1:]), -1)[:, None]), 1)
    outs = O.composite(rin, z, rd, None, white, bg)
    gen = torch.Generator().manual_seed(1)
    gs = [torch.randn(t.shape, generator=gen) for t in outs]
    sum((a * b).sum() for a, b in zip(outs, gs)).backward()
    d_raw = ops.composite_bwd(G(g["raw"]), z.to(DEV), rd.to(DEV), None, bg.to(DEV) if with_bg else None, with_bg, white,
                              *[t.to(DEV) for t in gs])
    ref = raw.grad
    assert maxabs(d_raw, ref) <= 2e-4 * float(ref.abs().max())
    # autograd.Function wiring
    raw_g = G(g["raw"]).clone().requires_grad_(True)
    from sahs_b200.volume_rendering_utils import composite
    o2 = composite(raw_g, z.to(DEV), rd.to(DEV), None, bg.to(DEV) if with_bg else None, with_bg, white)
    sum((a * b.to(DEV)).sum() for a, b in zip(o2, gs)).backward()
    assert maxabs(raw_g.grad, ref) <= 2e-4 * float(ref.abs().max())


@pytest.mark.parametrize("S", [1, 7, 40, 96, 200])
def test_composite_odd_sample_counts_vs_oracle(sahs, S):
    """Sample counts that are not a power-of-two number of 32-lane chunks (the kernel rounds the chunk count up to
    1 / 2 / 4 / 8: the last sample -- the background one -- then sits in an inner chunk), forward and backward."""
    from sahs_b200 import ops
    gen = torch.Generator().manual_seed(100 + S)
    R = 37
    raw = torch.randn(R, S, 16, generator=gen) * 2.0
    raw[..., -1] = torch.randn(R, S, generator=gen) * 20.0
    z, _ = torch.sort(torch.rand(R, S, generator=gen) * 0.6 + 0.48, dim=-1)
    rd = torch.randn(R, 3, generator=gen) * 0.2 + torch.tensor([0, 0, -1.0])
    bg = torch.cat((torch.rand(R, 3, generator=gen), torch.ones(R, 1), torch.zeros(R, 11)), -1)
    rin = raw.clone().requires_grad_(True)
    rr = torch.cat((rin[:, :-1], torch.cat((bg, rin[:, -1, -1:]), -1)[:, None]), 1)
    ref = O.composite(rr, z, rd, None, False, bg)
    out = ops.composite_fwd(raw.to(DEV), z.to(DEV), rd.to(DEV), None, bg.to(DEV), True, False)
    for n, o, r_ in zip(["rgb", "disp", "acc", "weights", "depth"], out, ref):
        if n == "disp":
            assert float(((o.cpu() - r_.detach()).abs() / r_.detach().abs()).max()) <= 1e-4
        else:
            assert maxabs(o, r_) <= 2e-5, (S, n, maxabs(o, r_))
    gs = [torch.randn(t.shape, generator=gen) for t in ref]
    sum((a * b).sum() for a, b in zip(ref, gs)).backward()
    d_raw = ops.composite_bwd(raw.to(DEV), z.to(DEV), rd.to(DEV), None, bg.to(DEV), True, False, *[t.to(DEV) for t in gs])
    assert maxabs(d_raw, rin.grad) <= 2e-4 * float(rin.grad.abs().max())


def test_composite_linearity_full_frame(sahs):
    """262,144 rays x 128 samples: acc == 1 with a background prior (last alpha is 1), rgb_map is linear in the
    background colour, weights are non-negative and sum to acc."""
    from sahs_b200 import ops
    R, S = 262144, 128
    gen = torch.Generator(device=DEV).manual_seed(7)
    raw = torch.randn(R, S, 16, device=DEV, generator=gen)
    raw[..., -1] = torch.randn(R, S, device=DEV, generator=gen) * 20
    z = torch.sort(torch.rand(R, S, device=DEV, generator=gen) * 0.6 + 0.48, -1)[0]
    rd = torch.randn(R, 3, device=DEV, generator=gen)
    bg0 = torch.zeros(R, 15, device=DEV)
    bg1 = torch.rand(R, 15, device=DEV, generator=gen)
    rgb0, _, acc, w, _ = ops.composite_fwd(raw, z, rd, None, bg0, True, False)
    rgb1, _, _, w1, _ = ops.composite_fwd(raw, z, rd, None, bg1, True, False)
    assert float((acc - 1).abs().max()) <= 2e-5
    assert float((w.sum(-1) - acc).abs().max()) <= 2e-5 and float(w.min()) >= 0
    assert torch.equal(w, w1)
    assert float((rgb1 - rgb0 - w[:, -1:] * bg1).abs().max()) <= 2e-5


# ------------------------------------------------------------------------------------------------------
# (2) fused field kernel
# ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["field_audio", "field_expr2"])
def test_field_forward_vs_reference_golden(sahs, name):
    g = load(name)
    cfg, spec, sd, model = _model(sahs, str(g["cfg_name"]))
    assert abs(FX.state_checksum(sd) - float(g["state_checksum"])) <= 1e-9 * float(g["state_checksum"])
    n = g["xyz"].shape[0]
    x = torch.cat((G(g["xyz"]), G(g["dirs"]), torch.zeros(n, 12, device=DEV)), -1)
    fr = FX.make_frame_inputs(spec, 8, 8, seed=int(g["seed"]), pose_z=0.78)
    with torch.no_grad():
        for level in ("coarse", "fine"):
            raw = model(level, x, fr["driving"].to(DEV), G(g["pose"]), None)       # reference call signature
            ref = C(g["ref_raw_" + level])
            assert raw.shape == (n, 16) and not bool(torch.isnan(raw).any())
            # 16-bit operands, fp32 accumulate: relative to the logit scale of each head.  The density logit is the
            # sensitive one: the dense fixture scales fc_alpha by 400 and the encoding of the warped point
            # amplifies the deformation net's rounding error by 2^(L-1) (L = 15 for the expression config).
            for sl, tol in ((slice(0, 3), 3e-2), (slice(3, 15), 3e-2), (slice(15, 16), 3e-2 if spec.xyz_L <= 10 else 0.2)):
                scale = max(1.0, float(ref[:, sl].abs().max()))
                assert maxabs(raw[:, sl], ref[:, sl]) <= tol * scale, (level, sl, maxabs(raw[:, sl], ref[:, sl]), scale)
    from sahs_b200 import ops
    assert ops.field_status()[0] == 0
    # the reference's own intermediates stored with the golden (first tile): warped point, ambient coordinates,
    # trilinear embedding.  fp32 tails of the kernel: the only fp16 content is the deformation nets' hidden layers
    # (split precision for 15 octaves).
    dbg = torch.zeros(128, 256, device=DEV)
    with torch.no_grad():
        dvec, pcode = model.driving_vector(fr["driving"].to(DEV)), model.pose_code(G(g["pose"]))
        z0 = torch.zeros(n, 1, device=DEV)
        model.field("coarse", G(g["xyz"]), G(g["dirs"]), z0, dvec, pcode, debug=dbg, debug_pass=16)   # SAHS_DBG_MAPPED
    d = dbg.cpu()
    tol_map = 2e-4 if spec.xyz_L <= 10 else 2e-6        # merged fp16 deformation phase | split precision
    assert maxabs(d[:, :3], C(g["mapped"])[:128]) <= tol_map, maxabs(d[:, :3], C(g["mapped"])[:128])
    assert maxabs(d[:, 3:3 + spec.amb_dim], C(g["amb"])[:128]) <= 50 * tol_map
    emb_ref = C(g["emb"])[:128]
    assert maxabs(d[:, 8:40], emb_ref) <= 2e-3 * float(emb_ref.abs().max()) + 10 * tol_map * 30


@pytest.mark.parametrize("cfg_name", ["audio/person_2_auto", "expression/person_2", "expression/person_1"])
def test_field_intermediates_vs_oracle(sahs, cfg_name):
    """Every pass of the render kernel against the oracle's intermediates on the trained-like fixture (first tile):
    deformation layers, warped point, trunk layers, fc_feat, both heads.  16-bit operands: 4e-3 of the layer's range."""
    cfg, spec, sd, model = _model(sahs, cfg_name, trained_like=True)
    fr = FX.make_frame_inputs(spec, 8, 8, seed=2, pose_z=FX.probe_pose_z(spec))
    gen = torch.Generator().manual_seed(17)
    n = 128
    xyz = (torch.rand(n, 3, generator=gen) * 2 - 1) * 0.3
    dirs = torch.randn(n, 3, generator=gen) * 0.3 + torch.tensor([0.0, 0.0, -1.0])
    with torch.no_grad():
        drv = O.driving_vector(sd, spec, fr["driving"])
        ref, inter = O.field_forward(sd, spec, "fine", xyz, dirs, drv, fr["pose"], return_intermediates=True)
        dvec, pcode = model.driving_vector(fr["driving"].to(DEV)), model.pose_code(fr["pose"].to(DEV))
        z0 = torch.zeros(n, 1, device=DEV)
        dbg = torch.zeros(128, 256, device=DEV)

        def run(dbg_pass):
            dbg.zero_()
            model.field("fine", xyz.to(DEV), dirs.to(DEV), z0, dvec, pcode, debug=dbg, debug_pass=dbg_pass)
            torch.cuda.synchronize()
            return dbg.cpu().clone()

        checks = []
        if spec.use_warp:
            for i in range(spec.warp_layers):
                want = torch.cat((inter[f"warp{i}"], inter[f"hyper{i}"]), 1)
                checks.append((f"deform{i}", run(i)[:, :want.shape[1]], want))
            d = run(16)
            checks.append(("mapped", d[:, :3], inter["mapped"]))
            checks.append(("amb", d[:, 3:3 + spec.amb_dim], inter["amb"]))
        checks.append(("emb", run(16)[:, 8:40], inter["emb"]))
        for i in range(spec.trunk_layers):
            checks.append((f"trunk{i}", run(32 + i), inter[f"trunk{i}"]))
        checks.append(("feat", run(32 + spec.trunk_layers), inter["feat"]))
        for i in range(4):
            checks.append((f"head{i}", run(64 + i), torch.cat((inter[f"dir{i}"], inter[f"seg{i}"]), 1)))
    bad = []
    for name, got, want in checks:
        err, rng = maxabs(got, want), max(float(want.abs().max()), 1e-6)
        if err > 4e-3 * rng + 2e-4:
            bad.append((name, err, rng))
    assert not bad, bad
    assert sahs.ops.field_status()[0] == 0


def test_field_forward_no_deformation_config(sahs):
    cfg, spec, sd, model = _model(sahs, "expression/person_1")
    gen = torch.Generator().manual_seed(9)
    n = 300                                                     # ragged: not a multiple of the 128-row tile
    xyz = (torch.rand(n, 3, generator=gen) * 2 - 1) * 0.35
    dirs = torch.randn(n, 3, generator=gen) * 0.3 + torch.tensor([0, 0, -1.0])
    fr = FX.make_frame_inputs(spec, 8, 8, seed=1)
    ref = O.field_forward(sd, spec, "fine", xyz, dirs, fr["driving"], fr["pose"])
    x = torch.cat((xyz, dirs), -1).to(DEV)
    with torch.no_grad():
        raw = model("fine", x, fr["driving"].to(DEV), fr["pose"].to(DEV), None)
    for sl in (slice(0, 3), slice(3, 15), slice(15, 16)):
        scale = max(1.0, float(ref[:, sl].abs().max()))
        assert maxabs(raw[:, sl], ref[:, sl]) <= 3e-2 * scale


def test_field_empty_and_ragged(sahs):
    cfg, spec, sd, model = _model(sahs, "audio/person_2_auto")
    fr = FX.make_frame_inputs(spec, 8, 8, seed=1)
    with torch.no_grad():
        out = model("coarse", torch.zeros(0, 18, device=DEV), fr["driving"].to(DEV), fr["pose"].to(DEV), None)
        assert out.shape == (0, 16)
        one = model("coarse", torch.tensor([[0.1, 0.0, -0.1, 0, 0, -1.0]], device=DEV), fr["driving"].to(DEV),
                    fr["pose"].to(DEV), None)
        many = model("coarse", torch.tensor([[0.1, 0.0, -0.1, 0, 0, -1.0]], device=DEV).expand(129, 6).contiguous(),
                     fr["driving"].to(DEV), fr["pose"].to(DEV), None)
    assert torch.equal(one[0], many[0]) and torch.equal(many[0], many[128])   # tile position does not matter


@pytest.mark.parametrize("cfg_name", ["audio/person_2_auto", "expression/person_2", "expression/person_1"])
def test_field_pair_kernel_matches_single_cta_kernel(sahs, cfg_name, monkeypatch):
    """The three render kernels run the same arithmetic on every row -- the CTA-pair kernel (default), the two-tile
    "duo" kernel (SAHS_FIELD_DUO=1: one cluster per SM pair, two tile pairs in flight, one in-order issuer) and the
    single-CTA kernel (SAHS_FIELD_PAIR=0): bit-identical outputs for odd tile counts, masked peer tiles (one tile),
    a set without work (two tiles), ragged tails and many tiles per cluster."""
    cfg, spec, sd, model = _model(sahs, cfg_name)
    fr = FX.make_frame_inputs(spec, 8, 8, seed=1)
    drv, pose = fr["driving"].to(DEV), fr["pose"].to(DEV)
    gen = torch.Generator().manual_seed(11)
    for n in (1, 129, 257, 3 * 128 + 1, 5 * 128 + 7, 300 * 128 + 5, 1200 * 128):
        xyz = (torch.rand(n, 3, generator=gen) * 2 - 1) * 0.35
        dirs = torch.randn(n, 3, generator=gen) * 0.3 + torch.tensor([0, 0, -1.0])
        x = torch.cat((xyz, dirs), -1).to(DEV)
        outs = []
        for pair, duo in (("0", "0"), ("1", "0"), ("1", "1")):
            monkeypatch.setenv("SAHS_FIELD_PAIR", pair)
            monkeypatch.setenv("SAHS_FIELD_DUO", duo)
            with torch.no_grad():
                outs.append(model("fine", x, drv, pose, None))
            torch.cuda.synchronize()
            assert sahs.ops.field_status()[0] == 0, (cfg_name, n, pair, duo, sahs.ops.field_status())
        assert torch.equal(outs[0], outs[1]), (cfg_name, n, "pair vs single")
        assert torch.equal(outs[0], outs[2]), (cfg_name, n, "duo vs single")
    assert sahs.ops.field_status()[0] == 0


# ------------------------------------------------------------------------------------------------------
# end to end: run_one_iter_of_nerf
# ------------------------------------------------------------------------------------------------------
E2E_CASES = ["e2e_audio_val", "e2e_expr2_val", "e2e_expr2_trained_val", "e2e_expr1_val", "e2e_audio_train_stoch",
             "e2e_audio_nobg_val", "e2e_audio_white_val"]


@pytest.mark.parametrize("name", E2E_CASES)
def test_run_one_iter_vs_reference_golden(sahs, name):
    """run_one_iter_of_nerf, free running (coarse pass -> sample_pdf -> fine pass, nothing fed from the reference),
    against the live reference's outputs: the three config families, validation and stochastic train mode (replayed
    draws), with a background prior, without one, and with white_background."""
    g = load(name)
    trained_like = bool(int(g["trained_like"]))
    cfg, spec, sd, model = _model(sahs, str(g["cfg_name"]), trained_like)
    assert abs(FX.state_checksum(sd) - float(g["state_checksum"])) <= 1e-9 * float(g["state_checksum"])
    H, W, mode = int(g["H"]), int(g["W"]), str(g["mode"])
    fr = FX.make_frame_inputs(spec, H, W, seed=int(g["seed"]), pose_z=float(g["pose_z"]))
    node = getattr(cfg.nerf, mode)
    draws = None
    if int(g["stochastic"]):
        node.perturb, node.radiance_field_noise_std = True, 0.1
        draws = {k: G(g["draw_" + k]) for k in ("t_rand", "noise_c", "u", "noise_f")}
    else:
        node.perturb, node.radiance_field_noise_std = False, 0.0
    bg_mode = str(g["bg_mode"])
    node.white_background = bg_mode == "white"
    bg = fr["background"].view(-1, 15).to(DEV) if bg_mode == "prior" else None
    pose = fr["pose"].to(DEV)
    with torch.no_grad():
        ro, rd = sahs.get_ray_bundle(H, W, fr["intrinsics"], pose)
        out = sahs.run_one_iter_of_nerf(H, W, fr["intrinsics"][0], model, ro, rd, cfg, mode=mode,
                                        driving=fr["driving"].to(DEV), pose=pose, pose_c=None,
                                        background_prior=bg, inHead=fr["mask"].to(DEV), _draws=draws)
    assert len(out) == 8
    if mode == "validation":
        assert out[0].shape == (H, W, 15) and out[3].shape == (H, W, 15) and out[7].shape == (H, W)
    names = ["rgb_c", "disp_c", "acc_c", "rgb_f", "disp_f", "acc_f", "w_last_f", "depth_f"]
    flat = {n: (o.reshape(-1, 15) if o.shape[-1] == 15 and o.dim() > 1 else o.reshape(-1)) for n, o in zip(names, out)}
    # north-star tolerance: max-abs 1e-2 on rgb (and the semantic channels) and depth, PSNR >= 50 dB.
    stress = spec.xyz_L > 10 and not trained_like
    for n in ("rgb_c", "rgb_f"):
        assert psnr(flat[n][:, :3], C(g["ref_" + n])[:, :3]) >= 50.0
    assert maxabs(flat["rgb_c"], C(g["ref_rgb_c"])) <= 1e-2
    assert maxabs(flat["acc_c"], C(g["ref_acc_c"])) <= 1e-2
    assert maxabs(flat["acc_f"], C(g["ref_acc_f"])) <= (1e-4 if bg_mode == "prior" else 1e-2)
    errs = {n: maxabs(flat[n], C(g["ref_" + n])) for n in ("rgb_c", "rgb_f", "depth_f", "w_last_f")}
    _report(f"[e2e] {name}: " + " ".join(f"{k} {v:.2e}" for k, v in errs.items())
            + f" psnr_f {psnr(flat['rgb_f'][:, :3], C(g['ref_rgb_f'])[:, :3]):.1f} dB")
    if not stress:
        assert errs["rgb_f"] <= 1e-2 and errs["depth_f"] <= 1e-2 and errs["w_last_f"] <= 1e-2, errs
        rel = float(((flat["disp_f"].cpu() - C(g["ref_disp_f"])).abs() / C(g["ref_disp_f"]).abs()).max())
        assert rel <= 2e-2
        rel = float(((flat["disp_c"].cpu() - C(g["ref_disp_c"])).abs() / C(g["ref_disp_c"]).abs()).max())
        assert rel <= 2e-2
    else:
        # STRESS CASE -- white-noise weights with 15 encoding octaves (expression/person_2,3): the field is spiky far
        # below the sample spacing and the *reference algorithm itself* moves by ~8e-3 (rgb) when the fine depths are
        # jittered by one fp32 ulp (tests/test_oracle_golden.py::test_fine_pass_conditioning), so a free-running
        # comparison cannot meet 1e-2 unless the coarse pass is bit-identical (the trained-like case of the same config,
        # e2e_expr2_trained_val, is held to the free-running bar above).  Here the fine pass is checked the way the
        # sample_pdf criterion is worded: fed the reference's own fine depths.
        assert errs["rgb_f"] <= 6e-2 and errs["depth_f"] <= 6e-2
        from sahs_b200 import ops
        with torch.no_grad():
            ro_f, rd_f = ro.reshape(-1, 3), rd.reshape(-1, 3)
            z_f = G(g["z_f"])
            raw_f = model.field("fine", ro_f, rd_f, z_f, model.driving_vector(fr["driving"].to(DEV)), model.pose_code(pose))
            rgb_f, disp_f, acc_f, w_f, depth_f = ops.composite_fwd(raw_f, z_f, rd_f, None, bg, True, False)
        assert maxabs(rgb_f, C(g["ref_rgb_f"])) <= 1e-2, maxabs(rgb_f, C(g["ref_rgb_f"]))
        assert psnr(rgb_f[:, :3], C(g["ref_rgb_f"])[:, :3]) >= 50.0
        assert maxabs(depth_f, C(g["ref_depth_f"])) <= 1e-2, maxabs(depth_f, C(g["ref_depth_f"]))
        assert maxabs(w_f[:, -1], C(g["ref_w_last_f"])) <= 1e-2
        k = g["raw_f"].shape[0]
        ref_raw = C(g["raw_f"])
        assert maxabs(raw_f[:k, :-1, :15], ref_raw[:, :-1, :15]) <= 5e-3      # colour / semantic logits of the MLP


def test_render_is_chunking_invariant(sahs):
    """Rays are independent: rendering a ray set in one call or in two halves gives identical bits."""
    cfg, spec, sd, model = _model(sahs, "audio/person_2_auto", trained_like=True)
    cfg.nerf.validation.perturb = False
    H, W = 8, 24
    fr = FX.make_frame_inputs(spec, H, W, seed=3)
    pose = fr["pose"].to(DEV)
    bg = fr["background"].view(-1, 15).to(DEV)
    with torch.no_grad():
        ro, rd = sahs.get_ray_bundle(H, W, fr["intrinsics"], pose)
        ro, rd = ro.reshape(-1, 3), rd.reshape(-1, 3)
        kw = dict(mode="train", driving=fr["driving"].to(DEV), pose=pose)
        cfg.nerf.train.perturb, cfg.nerf.train.radiance_field_noise_std = False, 0.0
        full = sahs.run_one_iter_of_nerf(H, W, 1.0, model, ro, rd, cfg, background_prior=bg, **kw)
        a = sahs.run_one_iter_of_nerf(H, W, 1.0, model, ro[:100], rd[:100], cfg, background_prior=bg[:100], **kw)
        b = sahs.run_one_iter_of_nerf(H, W, 1.0, model, ro[100:], rd[100:], cfg, background_prior=bg[100:], **kw)
    for f, x, y in zip(full, a, b):
        assert torch.equal(f, torch.cat((x, y), 0))


# ------------------------------------------------------------------------------------------------------
# semantic-weighted ray sampler
# ------------------------------------------------------------------------------------------------------
def test_weighted_sampler_structure_and_reproducibility(sahs):
    gen = torch.Generator().manual_seed(5)
    H = W = 512
    labels = torch.randint(0, 12, (H * W,), generator=gen)
    mask = torch.nn.functional.one_hot(labels, 12).to(torch.int32)
    prob = torch.rand(12, generator=gen) + 0.1
    prob[3] = 0.0                                              # one class is never sampled
    a = sahs.weighted_sample(mask.to(DEV), prob.to(DEV), 2048, seed=123).cpu()
    b = sahs.weighted_sample(mask.to(DEV), prob.to(DEV), 2048, seed=123).cpu()
    c = sahs.weighted_sample(mask.to(DEV), prob.to(DEV), 2048, seed=124).cpu()
    assert a.shape == (2048,) and a.dtype == torch.int64
    assert int(a.min()) >= 0 and int(a.max()) < H * W
    assert a.unique().numel() == 2048                           # without replacement
    assert set(a.tolist()) == set(b.tolist())                   # same seed -> same set
    assert len(set(a.tolist()) & set(c.tolist())) < 200         # another seed -> another draw
    assert not bool((labels[a] == 3).any())                     # zero-weight pixels are never drawn
    # class frequencies follow the class weights (each class has ~1/12 of the pixels): 4 sigma
    counts = torch.bincount(labels[a], minlength=12).double()
    expect = 2048 * (prob.double() * torch.bincount(labels, minlength=12).double())
    expect = expect / expect.sum() * 2048
    sigma = expect.clamp_min(1.0).sqrt()
    assert bool(((counts - expect).abs() <= 4.5 * sigma + 1).all()), (counts, expect)


def test_weighted_sampler_matches_reference_distribution(sahs):
    """Inclusion probabilities of every item against the reference's sequential np.random.choice(replace=False, p=...)
    on a case where sampling without replacement differs visibly from independent draws (n is 40 % of N)."""
    N, n, trials = 20, 8, 4000
    w = torch.tensor([8.0, 4, 4, 2, 2, 2, 1, 1, 1, 1, 1, 1, 0.5, 0.5, 0.5, 0.5, 0.25, 0.25, 0.25, 0.0])
    mask = torch.eye(N, dtype=torch.int32)                     # item i belongs to "class" i
    hits = torch.zeros(N, dtype=torch.float64)
    m_dev, w_dev = mask.to(DEV), w.to(DEV)
    for t in range(trials):
        idx = sahs.weighted_sample(m_dev, w_dev, n, seed=1000 + t)
        hits[idx.cpu()] += 1
    ours = hits / trials
    rng = np.random.default_rng(0)
    ref_hits = np.zeros(N)
    for t in range(trials):
        ref_hits[O.weighted_sample(mask, w, n, rng)] += 1
    ref = torch.from_numpy(ref_hits / trials)
    assert ours[-1] == 0 and abs(float(ours.sum()) - n) < 1e-9
    sigma = (ref * (1 - ref) / trials).sqrt() * (2 ** 0.5)      # both sides are Monte-Carlo estimates
    assert bool(((ours - ref).abs() <= 4.5 * sigma + 2e-3).all()), (ours, ref)


# ------------------------------------------------------------------------------------------------------
# boundary: torch custom operators, remaining reference-named helpers
# ------------------------------------------------------------------------------------------------------
def test_custom_ops_registered_and_consistent(sahs):
    """torch.ops.sahs_b200.* (sahs_b200/custom_ops.py): every hot-path entry point is a registered PyTorch operator with
    a CUDA kernel and a Meta (shape) kernel, no CPU kernel; schema / fake-tensor consistency via torch.library.opcheck."""
    from sahs_b200 import custom_ops
    T = torch.ops.sahs_b200
    for n in custom_ops.OP_NAMES:
        assert hasattr(T, n), n
    g = load("composite_bg")
    raw, z, rd, bg = G(g["raw"]), G(g["z"]), G(g["rd"]), G(g["bg"])
    checks = ("test_schema", "test_faketensor")
    torch.library.opcheck(T.composite_fwd, (raw, z, rd, None, bg, True, False), test_utils=checks)
    w = torch.rand(64, 64, device=DEV)
    zc = T.coarse_z(64, 64, 0.4838, 1.0838, False, sahs.ops.linspace_dev(64, DEV), None)
    torch.library.opcheck(T.sample_pdf_merge, (zc, w, 64, None), test_utils=checks)
    torch.library.opcheck(T.positional_encoding, (torch.randn(33, 3, device=DEV), 10, True), test_utils=checks)
    torch.library.opcheck(T.get_ray_bundle, (8, 8, 1200.0, 1200.0, 0.5, 0.5, FX.make_pose(0).to(DEV)), test_utils=checks)
    cfg, spec, sd, model = _model(sahs, "audio/person_2_auto", trained_like=True)
    fr = FX.make_frame_inputs(spec, 8, 8, seed=1)
    st = model.packed_level("fine")
    with torch.no_grad():
        fc = model.frame_constants("fine", model.driving_vector(fr["driving"].to(DEV)), model.pose_code(fr["pose"].to(DEV)))
    ro = torch.zeros(16, 3, device=DEV)
    ro[:, 2] = 0.78
    rdir = torch.tensor([[0.0, 0.0, -1.0]], device=DEV).expand(16, 3).contiguous()
    z16 = zc[:16].contiguous()
    torch.library.opcheck(T.field_fwd, (st["spec_ints"], 1, st["packed"], fc, st["grid"], ro, rdir, z16), test_utils=checks)
    # the public functions are these operators
    raw16 = model.field("fine", ro, rdir, z16, None, None, frame_const=fc)
    assert torch.equal(raw16, T.field_fwd(st["spec_ints"], 1, st["packed"], fc, st["grid"], ro, rdir, z16))
    # no CPU kernel: the dispatcher refuses CPU tensors (no fallback)
    with pytest.raises((NotImplementedError, RuntimeError)):
        T.composite_fwd(raw.cpu(), z.cpu(), rd.cpu(), None, bg.cpu(), True, False)
    with pytest.raises(RuntimeError):
        sahs.volume_render_radiance_field(raw.cpu(), z.cpu(), rd.cpu())


def test_ray_bundle_by_mask_vs_oracle(sahs):
    pose = FX.make_pose(3, 0.78, 10.0)
    intr = [37.5, 40.0, 0.48, 0.53]
    mask = (torch.rand(16, 16, generator=torch.Generator().manual_seed(4)) > 0.4).float()
    ro, rd = sahs.get_ray_bundle_by_mask(16, 16, np.array(intr), pose.to(DEV), mask.to(DEV))
    ro_o, rd_o = O.get_ray_bundle_by_mask(16, 16, intr, pose, mask)     # pinned against the reference on CPU
    assert torch.equal(ro.cpu(), ro_o) and torch.equal(rd.cpu(), rd_o)


def test_trainable_background_gets_its_gradient(sahs):
    """`train_background` (ref: train_stage_rays_auto.py:171-176, :245): the background prior is an nn.Parameter; its
    gradient through the fused overwrite + compositing must equal autograd through the oracle."""
    from sahs_b200.volume_rendering_utils import composite
    g = load("composite_bg")
    raw, z, rd = C(g["raw"]), C(g["z"]), C(g["rd"])
    bg = C(g["bg"]).clone().requires_grad_(True)
    raw_r = raw.clone().requires_grad_(True)
    rin = torch.cat((raw_r[:, :-1], torch.cat((bg, raw_r[:, -1, -1:]), -1)[:, None]), 1)
    outs = O.composite(rin, z, rd, None, False, bg)
    gen = torch.Generator().manual_seed(3)
    gs = [torch.randn(t.shape, generator=gen) for t in outs]
    sum((a * b).sum() for a, b in zip(outs, gs)).backward()
    bg_g = G(g["bg"]).clone().requires_grad_(True)
    raw_g = G(g["raw"]).clone().requires_grad_(True)
    o2 = composite(raw_g, z.to(DEV), rd.to(DEV), None, bg_g, True, False)
    sum((a * b.to(DEV)).sum() for a, b in zip(o2, gs)).backward()
    assert bg_g.grad is not None
    assert maxabs(bg_g.grad, bg.grad) <= 2e-5 * float(bg.grad.abs().max())
    assert maxabs(raw_g.grad, raw_r.grad) <= 2e-4 * float(raw_r.grad.abs().max())
    # background only (frozen field): still differentiable
    bg_h = G(g["bg"]).clone().requires_grad_(True)
    o3 = composite(G(g["raw"]), z.to(DEV), rd.to(DEV), None, bg_h, True, False)
    o3[0].sum().backward()
    assert maxabs(bg_h.grad, o3[3][:, -1:].expand(-1, 15)) == 0.0


@pytest.mark.parametrize("central", [False, True])
def test_normal_map_vs_oracle(sahs, central):
    """torch_normal_map (ref: eval_stage_rays.py:116-151) on a rendered-looking depth map with and without the
    background-weight clean-up; the oracle is pinned against the reference's function on CPU."""
    gen = torch.Generator().manual_seed(8)
    n = 512
    yy, xx = torch.meshgrid(torch.linspace(-1, 1, n), torch.linspace(-1, 1, n), indexing="ij")
    depth = 0.6 + 0.1 * torch.sin(3 * xx) * torch.cos(2 * yy) + 0.002 * torch.rand(n, n, generator=gen)
    weights = torch.rand(n, n, generator=gen) * 0.5
    focal = [1200.0, 1200.0, 0.5, 0.5]
    for w in (None, weights):
        want = O.torch_normal_map(depth, focal, w, True, central)
        got = sahs.torch_normal_map(depth.to(DEV), np.array(focal), None if w is None else w.to(DEV), clean=True,
                                    central_difference=central)
        assert got.shape == want.shape and got.dtype == torch.float32
        assert maxabs(got, want) <= 0.05, maxabs(got, want)          # values in [0, 255]; fp32 rounding of the normalisation
    assert sahs.torch_normal_map(depth.to(DEV), focal, weights.to(DEV), clean=False).shape == (n - 1, n - 1, 3)


def test_in_kernel_random_draws(sahs):
    """Stochastic mode without materialised random tensors: t_rand, the density noise and u are drawn inside the kernels
    (Philox-4x32-10, include/sahs_b200.h `sahs_rng`).  (1) the generator: range, moments, reproducibility, independent
    streams, the device counter; (2) every kernel that draws equals, BIT FOR BIT, the same kernel fed the materialised
    values (`sahs_rng_fill`) -- forward and backward of the compositing regenerate identical noise; (3) the pipeline."""
    from sahs_b200 import ops
    seed, n = 123456789, 1 << 20
    u = ops.rng_fill(n, seed, None, 0, False, 1.0, DEV)
    g = ops.rng_fill(n, seed, None, 1, True, 1.0, DEV)
    assert float(u.min()) >= 0.0 and float(u.max()) < 1.0
    assert abs(float(u.mean()) - 0.5) < 2e-3 and abs(float(u.var()) - 1.0 / 12) < 1e-3
    assert abs(float(g.mean())) < 4e-3 and abs(float(g.var()) - 1.0) < 6e-3 and abs(float((g ** 4).mean()) - 3.0) < 0.05
    assert torch.equal(u, ops.rng_fill(n, seed, None, 0, False, 1.0, DEV))
    assert not torch.equal(u, ops.rng_fill(n, seed, None, 2, False, 1.0, DEV))          # another stream
    ctr = torch.full((), 5, dtype=torch.int64, device=DEV)
    assert torch.equal(ops.rng_fill(n, seed, ctr, 0, False, 1.0, DEV), ops.rng_fill(n, seed + 5, None, 0, False, 1.0, DEV))
    assert abs(float(torch.corrcoef(torch.stack((u[:-1], u[1:])))[0, 1])) < 4e-3         # neighbours uncorrelated
    # (2) kernels
    R, S, NF = 300, 64, 64
    opts = O.RenderOpts(num_coarse=S, near=0.4838, far=1.0838, perturb=True)
    z_rng = ops.coarse_z(R, S, opts.near, opts.far, False, DEV, rng=(seed, ctr, 0))
    t_rand = ops.rng_fill(R * S, seed, ctr, 0, False, 1.0, DEV).view(R, S)
    assert torch.equal(z_rng, ops.coarse_z(R, S, opts.near, opts.far, False, DEV, t_rand))
    assert torch.equal(z_rng.cpu(), O.coarse_z(opts, R, t_rand.cpu()))
    gen = torch.Generator(device=DEV).manual_seed(3)
    raw = torch.randn(R, S, 16, device=DEV, generator=gen)
    raw[..., -1] = torch.randn(R, S, device=DEV, generator=gen) * 10
    rd = torch.randn(R, 3, device=DEV, generator=gen) * 0.1
    rd[:, 2] = -1
    bg = torch.rand(R, 15, device=DEV, generator=gen)
    std = 0.1
    noise = ops.rng_fill(R * S, seed, ctr, 1, True, std, DEV).view(R, S)
    a = ops.composite_fwd(raw, z_rng, rd, None, bg, True, False, noise_std=std, rng=(seed, ctr, 1))
    b = ops.composite_fwd(raw, z_rng, rd, noise, bg, True, False)
    for x, y in zip(a, b):
        assert torch.equal(x, y)
    assert not torch.equal(a[3], ops.composite_fwd(raw, z_rng, rd, None, bg, True, False)[3])   # the noise matters
    gs = [torch.randn(t.shape, device=DEV, generator=gen) for t in a]
    da = ops.composite_bwd(raw, z_rng, rd, None, bg, True, False, *gs, noise_std=std, rng=(seed, ctr, 1))
    db = ops.composite_bwd(raw, z_rng, rd, noise, bg, True, False, *gs)
    assert torch.equal(da, db)
    w = a[3]
    zs, zm = ops.sample_pdf_merge(z_rng, w, NF, None, rng=(seed, ctr, 2))
    uu = ops.rng_fill(R * NF, seed, ctr, 2, False, 1.0, DEV).view(R, NF)
    zs2, zm2 = ops.sample_pdf_merge(z_rng, w, NF, uu)
    assert torch.equal(zs, zs2) and torch.equal(zm, zm2)
    # (3) pipeline in the shipped stochastic training settings: finite, reproducible under torch.manual_seed, fresh
    # draws from call to call, and the autograd path runs
    from sahs_b200 import train_utils as TU
    cfg, spec, sd, model = _model(sahs, "audio/person_2_auto", trained_like=True)
    cfg.nerf.train.perturb, cfg.nerf.train.radiance_field_noise_std = True, 0.1
    fr = FX.make_frame_inputs(spec, 8, 8, seed=2)
    pose = fr["pose"].to(DEV)
    with torch.no_grad():
        ro, rdd = sahs.get_ray_bundle(8, 8, fr["intrinsics"], pose)
    kw = dict(mode="train", driving=fr["driving"].to(DEV), pose=pose, background_prior=fr["background"].view(-1, 15).to(DEV))

    def run():
        with torch.no_grad():
            return sahs.run_one_iter_of_nerf(8, 8, 1.0, model, ro, rdd, cfg, **kw)

    torch.manual_seed(11)
    TU._RNG_CALLS[0] = 0
    o1, o2 = run(), run()
    torch.manual_seed(11)
    TU._RNG_CALLS[0] = 0
    o3 = run()
    assert all(bool(torch.isfinite(t).all()) for t in o1)
    assert torch.equal(o1[3], o3[3]) and not torch.equal(o1[3], o2[3])
    out = sahs.run_one_iter_of_nerf(8, 8, 1.0, model, ro, rdd, cfg, **kw)
    (out[3].sum() + out[0].sum()).backward()
    assert all(p.grad is not None and bool(torch.isfinite(p.grad).all()) for p in model.parameters())


def test_weighted_sampler_validate_raises_like_numpy(sahs):
    from sahs_b200 import ops
    mask = torch.eye(20, dtype=torch.int32, device=DEV)
    w = torch.zeros(20, device=DEV)
    w[:5] = 1.0
    assert ops.weighted_sample(mask, w, 5, seed=1, validate=True).sort().values.tolist() == [0, 1, 2, 3, 4]
    with pytest.raises(ValueError, match="Fewer non-zero entries"):
        ops.weighted_sample(mask, w, 8, seed=1, validate=True)
