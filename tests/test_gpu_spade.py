"""GPU tests (-m gpu) of the Stage-II SPADE path: the implicit-GEMM conv kernel in every mode / epilogue against the CPU
interpretation of the same packed operands (tests/spade_emulator.py, itself checked against torch on the CPU), the two
helper kernels against torch, and both generators end to end against the reference's golden outputs."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import sahs_fixtures as FX
import spade_fixtures as SF
from oracle import spade_oracle as SO

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GOLD = os.path.join(SF.REPO, "tests", "golden")


def _report(line):
    print(line)
    d = os.path.join(FX.REPO, "gpurun_out")
    if os.path.isdir(d):
        with open(os.path.join(d, "test_report.txt"), "a") as f:
            f.write(line + "\n")


@pytest.fixture(scope="module")
def gen():
    from sahs_b200 import spade as SP
    return SP.Generator()


def _status():
    from sahs_b200 import ops
    return ops.spade_conv_status()


CASES = [
    # name, cin, cout, in (h, w), out (h, w), mode, up, down, epilogue
    ("s1_64_64", 64, 64, (16, 24), (16, 24), 0, 0, 0, ""),
    ("s1_tail", 64, 128, (9, 7), (9, 7), 0, 0, 0, "relu"),
    ("s1_256_256_add", 256, 256, (12, 20), (12, 20), 0, 0, 0, "add"),
    ("s1_up", 128, 128, (8, 12), (16, 24), 0, 1, 0, "relu"),
    ("s1_down", 64, 128, (16, 24), (8, 12), 0, 0, 1, "relu"),
    ("s2", 128, 128, (16, 24), (8, 12), 1, 0, 0, "add"),
    ("t2", 256, 256, (6, 10), (12, 20), 2, 0, 0, ""),
    ("t2_wide", 128, 128, (8, 128), (16, 256), 2, 0, 0, ""),
    ("first", 3, 64, (20, 28), (20, 28), 3, 0, 0, ""),
    ("last_f32", 64, 3, (20, 28), (20, 28), 0, 0, 0, "f32"),
    ("many_tiles", 64, 64, (96, 128), (96, 128), 0, 0, 0, "relu"),
]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_conv_kernel_vs_cpu_interpretation(gen, case):
    import spade_emulator as EM
    from sahs_b200 import spade as SP
    name, cin, cout, (ih, iw), (oh, ow), mode, up, down, epi = case
    g = torch.Generator().manual_seed(len(name) * 131 + cin)
    w = torch.randn((cin, cout, 3, 3) if mode == 2 else (cout, cin, 3, 3), generator=g) * (1.0 / (9 * cin) ** 0.5)
    b = torch.randn(cout, generator=g) * 0.2
    p = SP._pack_conv(w, b, transposed=(mode == 2))
    pd = SP._pack_conv(w.to(DEV), b.to(DEV), transposed=(mode == 2))
    if mode == 3:
        x = EM.EmulatedMixin._image(torch.rand(1, 3, ih, iw, generator=g)).half()
    else:
        x = torch.randn(ih, iw, cin, generator=g).half()
    add = torch.randn(oh, ow, cout, generator=g).half() if epi == "add" else None
    kw = dict(up=up, down=down, relu=(epi == "relu"), add=add, f32=(epi == "f32"))
    want = EM.conv(p, x.float(), oh, ow, mode, **{**kw, "add": None if add is None else add.float()})
    kw["add"] = None if add is None else add.to(DEV)
    got = gen._conv(pd, x.to(DEV), oh, ow, mode, **kw)
    if mode == 2:      # the transposed conv has two implementations (one nine-tap launch / four parity-class launches)
        cls = torch.ops.sahs_b200.spade_conv_t2(x.to(DEV), [c.packed for c in pd.classes], [c.bias for c in pd.classes], cin, cout)
        one = gen._conv(pd.full, x.to(DEV), oh, ow, 2)
        assert float((cls.float() - one.float()).abs().max()) <= 2e-3 * max(1.0, float(one.float().abs().max()))
    torch.cuda.synchronize()
    assert _status()[0] == 0, _status()
    assert got.dtype == (torch.float32 if epi == "f32" else torch.float16) and tuple(got.shape) == (oh, ow, cout)
    err = float((got.float().cpu() - want).abs().max())
    scale = max(1.0, float(want.abs().max()))
    _report(f"[spade conv] {name}: max-abs {err:.2e} (scale {scale:.2f})")
    assert err <= (2e-4 if epi == "f32" else 1.5e-3) * scale


@pytest.mark.parametrize("c,shift", [(64, 0), (128, 1), (256, 0)])
def test_spade_epilogue_vs_cpu_interpretation(gen, c, shift):
    """[gamma|beta] tiles + instance-norm modulation + LeakyReLU(0.2) in the conv epilogue; aux read through aux_shift"""
    import spade_emulator as EM
    from sahs_b200 import spade as SP
    g = torch.Generator().manual_seed(c + shift)
    oh, ow = 10, 14
    wg, wb = torch.randn(c, 128, 3, 3, generator=g) * 0.02, torch.randn(c, 128, 3, 3, generator=g) * 0.02
    bg, bb = torch.randn(c, generator=g) * 0.1, torch.randn(c, generator=g) * 0.1
    p = SP._pack_gamma_beta(wg, bg, wb, bb)
    pd = SP._pack_gamma_beta(wg.to(DEV), bg.to(DEV), wb.to(DEV), bb.to(DEV))
    actv = torch.randn(oh, ow, 128, generator=g).clamp_min(0).half()
    aux = (torch.randn(oh >> shift, ow >> shift, c, generator=g) * 3 + 1).half()
    mean, rstd = EM.EmulatedMixin()._stats(aux)
    want = EM.conv(p, actv.float(), oh, ow, 0, spade=(aux.float(), mean, rstd), aux_shift=shift)
    mg, rg = gen._stats(aux.to(DEV))
    assert float((mg.cpu() - mean).abs().max()) <= 1e-4 and float((rg.cpu() / rstd - 1).abs().max()) <= 1e-4
    got = gen._conv(pd, actv.to(DEV), oh, ow, 0, spade=(aux.to(DEV), mg, rg), aux_shift=shift)
    torch.cuda.synchronize()
    assert _status()[0] == 0, _status()
    err = float((got.float().cpu() - want).abs().max())
    _report(f"[spade epilogue] c={c} shift={shift}: max-abs {err:.2e} (scale {float(want.abs().max()):.2f})")
    assert err <= 2e-3 * max(1.0, float(want.abs().max()))


def test_instnorm_stats_and_avgpool_vs_torch(gen):
    g = torch.Generator().manual_seed(9)
    for h, w, c in ((256, 256, 64), (37, 53, 128), (8, 8, 256)):
        x = (torch.randn(h, w, c, generator=g) * 2 + 0.5).half().to(DEV)
        mean, rstd = gen._stats(x)
        xf = x.float().reshape(-1, c).double()
        assert float((mean.double() - xf.mean(0)).abs().max()) <= 1e-5
        want = 1.0 / torch.sqrt(xf.var(0, unbiased=False) + 1e-5)
        assert float((rstd.double() / want - 1).abs().max()) <= 1e-5
    x = torch.randn(64, 96, 128, generator=g).half().to(DEV)
    y = gen._avgpool(x)
    want = F.avg_pool2d(x.float().permute(2, 0, 1).unsqueeze(0), 2, stride=2)[0].permute(1, 2, 0)
    assert float((y.float() - want).abs().max()) <= 2e-3


@pytest.mark.parametrize("tag", ["64", "96x128"])
def test_generator_vs_reference_golden(tag):
    from sahs_b200 import spade as SP
    g = np.load(os.path.join(GOLD, "spade_gen.npz"))
    m = SP.Generator()
    m.load_state_dict(SF.make_state_dict("generator", seed=0), strict=True)
    m = m.to(DEV)
    taps = {}
    out = m(torch.from_numpy(g[f"i_src_{tag}"]).to(DEV), torch.from_numpy(g[f"i_raw_{tag}"]).to(DEV), taps)
    torch.cuda.synchronize()
    assert _status()[0] == 0, _status()
    ref = torch.from_numpy(g[f"ref_out_{tag}"])
    lines = []
    if tag == "64":
        for k in ("layer2", "layer4", "layer5", "layer6"):
            r = torch.from_numpy(g[f"ref_{k}_{tag}"])
            e = float((taps[k].float().cpu().permute(2, 0, 1).unsqueeze(0) - r).abs().max()) / float(r.abs().max())
            lines.append(f"{k} {e:.1e}")
            assert e <= 2e-2, (k, e)
    err = float((out.cpu() - ref).abs().max())
    rng = float(ref.abs().max())
    mse = float(((out.cpu() - ref) ** 2).mean())
    psnr = 10 * np.log10(rng * rng / max(mse, 1e-30))
    _report(f"[spade e2e] Generator {tag}: max-abs {err:.2e} of range {rng:.2f}, PSNR {psnr:.1f} dB; " + " ".join(lines))
    assert tuple(out.shape) == tuple(ref.shape) and out.dtype == torch.float32
    assert err <= 2e-2 * rng and psnr >= 45.0


def test_generator_audio_vs_reference_golden():
    from sahs_b200 import spade as SP
    g = np.load(os.path.join(GOLD, "spade_audio.npz"))
    m = SP.Generator_audio()
    m.load_state_dict(SF.make_state_dict("generator_audio", seed=1), strict=True)
    m = m.to(DEV)
    out = m(torch.from_numpy(g["i_src"]).to(DEV), torch.from_numpy(g["i_raw"]).to(DEV), torch.from_numpy(g["audio"]).to(DEV))
    torch.cuda.synchronize()
    assert _status()[0] == 0, _status()
    ref = torch.from_numpy(g["ref_out"])
    err, rng = float((out.cpu() - ref).abs().max()), float(ref.abs().max())
    _report(f"[spade e2e] Generator_audio 64: max-abs {err:.2e} of range {rng:.2f}")
    assert err <= 2e-2 * rng


def test_generator_full_frame_vs_oracle_and_timing():
    """512x512 (the Stage-I frame size): size-independent check against the oracle on the host + a first timing"""
    from sahs_b200 import spade as SP
    sd = SF.make_state_dict("generator", seed=0)
    inp = SF.make_inputs(512, 512, seed=5)
    m = SP.Generator()
    m.load_state_dict(sd, strict=True)
    m = m.to(DEV)
    a, b = inp["i_src"].to(DEV), inp["i_raw"].to(DEV)
    out = m(a, b)
    torch.cuda.synchronize()
    assert _status()[0] == 0, _status()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        m(a, b)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    want = SO.generator(sd, inp["i_src"], inp["i_raw"])
    err, rng = float((out.cpu() - want).abs().max()), float(want.abs().max())
    mse = float(((out.cpu() - want) ** 2).mean())
    _report(f"[spade e2e] Generator 512x512: {ms:.2f} ms per frame; max-abs {err:.2e} of range {rng:.2f}, "
            f"PSNR {10 * np.log10(rng * rng / max(mse, 1e-30)):.1f} dB")
    assert err <= 3e-2 * rng


@pytest.mark.parametrize("kind", ["generator", "generator_audio"])
def test_graphed_generator_equals_eager(kind):
    """one CUDA graph per frame (GraphedGenerator) = the eager forward, bit for bit, also after the inputs change"""
    from sahs_b200 import spade as SP
    m = SP.Generator() if kind == "generator" else SP.Generator_audio()
    m.load_state_dict(SF.make_state_dict(kind, seed=2), strict=True)
    m = m.to(DEV)
    i0, i1 = SF.make_inputs(64, 96, seed=1), SF.make_inputs(64, 96, seed=2)
    args = lambda i: tuple(i[k].to(DEV) for k in (("i_src", "i_raw") if kind == "generator" else ("i_src", "i_raw", "audio")))
    g = SP.GraphedGenerator(m, *args(i0))
    for i in (i0, i1, i0):
        want = m(*args(i)).clone()
        got = g(*args(i))
        torch.cuda.synchronize()
        assert torch.equal(got, want)
    assert _status()[0] == 0
    g.graph.reset()


def test_custom_ops_cover_stage2():
    from sahs_b200 import custom_ops
    for name in ("spade_conv", "instnorm_stats", "avgpool2"):
        assert name in custom_ops.OP_NAMES and hasattr(torch.ops.sahs_b200, name)
    x = torch.randn(8, 8, 64, device=DEV).half()
    with pytest.raises((NotImplementedError, RuntimeError)):
        torch.ops.sahs_b200.avgpool2(x.cpu())                      # no CPU kernel: the dispatcher raises


@pytest.mark.parametrize("kind", ["generator", "generator_audio"])
def test_identity_cache_equals_full_forward_on_gpu(kind):
    """clip refinement: the identity photo's part (IdEncoder + the conditioning activations of its maps) once per clip;
    every frame then equals the full forward bit for bit, eagerly and as a CUDA graph of the per-frame work only"""
    from sahs_b200 import spade as SP
    m = SP.Generator() if kind == "generator" else SP.Generator_audio()
    m.load_state_dict(SF.make_state_dict(kind, seed=2), strict=True)
    m = m.to(DEV)
    i0, i1 = SF.make_inputs(64, 96, seed=1), SF.make_inputs(64, 96, seed=2)
    frame = (lambda i: (i["i_raw"].to(DEV), i["audio"].to(DEV))) if kind == "generator_audio" else (lambda i: (i["i_raw"].to(DEV),))
    src = i0["i_src"].to(DEV)
    ident = m.encode_identity(src)
    m.tally = {}
    m(src, *frame(i0))
    full, m.tally = m.tally["conv_launches"], {}
    first = m.refine(ident, *frame(i0)).clone()
    m.tally = {}
    second = m.refine(ident, *frame(i1)).clone()
    cached, m.tally = m.tally["conv_launches"], None
    assert full == 70 and cached == 70 - 9 - (18 if kind == "generator" else 12)
    assert torch.equal(first, m(src, *frame(i0))) and torch.equal(second, m(src, *frame(i1)))
    g = SP.GraphedGenerator(m, *frame(i0), identity=ident)
    for i in (i1, i0):
        assert torch.equal(g(*frame(i)), m(src, *frame(i)))
    assert _status()[0] == 0
    g.graph.reset()
