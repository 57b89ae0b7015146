"""CPU tests of the Stage-II SPADE path: oracle vs the reference's golden outputs, drop-in
state_dict compatibility, and the host-side logic of sahs_b200/spade.py (weight packing, gather rules, network wiring)
interpreted on the CPU by tests/spade_emulator.py."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import sahs_fixtures as FX  # noqa: F401  (puts the package on sys.path)
import spade_fixtures as SF
from oracle import spade_oracle as SO

GOLD = os.path.join(SF.REPO, "tests", "golden")


def _t(a):
    return torch.from_numpy(np.asarray(a))


def test_oracle_matches_reference_golden():
    g = np.load(os.path.join(GOLD, "spade_gen.npz"))
    sd = SF.make_state_dict("generator", seed=0)
    for tag in ("64", "96x128"):
        out, inter = SO.generator(sd, _t(g[f"i_src_{tag}"]), _t(g[f"i_raw_{tag}"]), return_intermediates=True)
        ref = _t(g[f"ref_out_{tag}"])
        assert float((out - ref).abs().max()) <= 2e-4 * float(ref.abs().max())
        if tag == "64":
            for k in ("layer2", "layer4", "layer5", "layer6"):
                r = _t(g[f"ref_{k}_{tag}"])
                assert float((inter[k] - r).abs().max()) <= 2e-4 * float(r.abs().max())
    ga = np.load(os.path.join(GOLD, "spade_audio.npz"))
    sda = SF.make_state_dict("generator_audio", seed=1)
    out = SO.generator_audio(sda, _t(ga["i_src"]), _t(ga["i_raw"]), _t(ga["audio"]))
    assert float((out - _t(ga["ref_out"])).abs().max()) <= 2e-4 * float(np.abs(ga["ref_out"]).max())


@pytest.mark.parametrize("kind", ["generator", "generator_audio"])
def test_state_dict_is_the_references(kind):
    """same keys and shapes as the reference module (tests/golden/spade_keys.json, written from the live reference), and a
    reference checkpoint loads with strict=True"""
    from sahs_b200 import spade as SP
    m = SP.Generator() if kind == "generator" else SP.Generator_audio()
    want = SF.load_keys(kind)
    got = {k: list(v.shape) for k, v in m.state_dict().items()}
    assert got == want
    assert list(got) == list(want)                       # same order, too
    m.load_state_dict(SF.make_state_dict(kind, seed=3), strict=True)


def _packed_conv(cin, cout, transposed=False, seed=0):
    from sahs_b200 import spade as SP
    gen = torch.Generator().manual_seed(seed)
    w = torch.randn((cin, cout, 3, 3) if transposed else (cout, cin, 3, 3), generator=gen) * 0.1
    b = torch.randn(cout, generator=gen)
    return SP._pack_conv(w, b, transposed), w.half().float(), b


@pytest.mark.parametrize("case", ["s1", "s1_up", "s1_down", "s2", "t2", "first", "tiny_cout"])
def test_packed_conv_and_gather_rules_vs_torch(case):
    import spade_emulator as EM
    from sahs_b200 import spade as SP
    gen = torch.Generator().manual_seed(5)
    if case == "first":
        p, w, b = _packed_conv(3, 64)
        img = torch.rand(1, 3, 10, 12, generator=gen)
        x = EM.EmulatedMixin._image(img)
        got = EM.conv(p, x, 10, 12, SP.MODE_FIRST)
        want = F.conv2d(img, w, b, padding=1)
    elif case == "tiny_cout":
        p, w, b = _packed_conv(64, 3)
        assert (p.ntile, p.ntiles) == (16, 1)
        xin = torch.randn(1, 64, 9, 7, generator=gen)
        got = EM.conv(p, xin[0].permute(1, 2, 0), 9, 7, SP.MODE_S1)
        want = F.conv2d(xin, w, b, padding=1)
    else:
        cin, cout = 128, 256
        xin = torch.randn(1, cin, 6, 10, generator=gen)
        x = xin[0].permute(1, 2, 0).contiguous()
        if case == "t2":
            p, w, b = _packed_conv(cin, cout, transposed=True)
            got = EM.conv(p, x, 12, 20, SP.MODE_T2)                     # small input: one nine-tap launch (mode 2)
            want = F.conv_transpose2d(xin, w, b, stride=2, padding=1, output_padding=1)
            assert float((EM.conv_t2(p, x) - got).abs().max()) <= 1e-5     # = the four parity-class convs (mode 4)
        else:
            p, w, b = _packed_conv(cin, cout)
            assert (p.ntile, p.ntiles) == (128, 2)
            if case == "s1":
                got, want = EM.conv(p, x, 6, 10, SP.MODE_S1), F.conv2d(xin, w, b, padding=1)
            elif case == "s1_up":
                got = EM.conv(p, x, 12, 20, SP.MODE_S1, up=1)
                want = F.conv2d(F.interpolate(xin, size=(12, 20), mode="nearest"), w, b, padding=1)
            elif case == "s1_down":
                got = EM.conv(p, x, 3, 5, SP.MODE_S1, down=1)
                want = F.conv2d(F.interpolate(xin, size=(3, 5), mode="nearest"), w, b, padding=1)
            else:
                got, want = EM.conv(p, x, 3, 5, SP.MODE_S2), F.conv2d(xin, w, b, stride=2, padding=1)
    want = want[0].permute(1, 2, 0)
    assert got.shape == want.shape
    assert float((got - want).abs().max()) <= 2e-4 * max(1.0, float(want.abs().max()))


def test_generator_host_logic_vs_oracle():
    """the product's orchestration (folded BatchNorm / spectral norm, [gamma|beta] tiles, resize shifts, folded upsample,
    residual wiring) interpreted on the CPU = the oracle, up to the fp16 rounding of the packed weights"""
    import spade_emulator as EM
    sd = SF.make_state_dict("generator", seed=0)
    inp = SF.make_inputs(32, 40, seed=2)
    m = EM.EmulatedGenerator()
    m.load_state_dict(sd, strict=True)
    taps = {}
    got = m(inp["i_src"], inp["i_raw"], taps)
    want, inter = SO.generator(sd, inp["i_src"], inp["i_raw"], return_intermediates=True)
    for k, v in inter.items():
        if k not in taps:
            continue
        t = taps[k].permute(2, 0, 1).unsqueeze(0)
        assert float((t - v).abs().max()) <= 5e-3 * float(v.abs().max()), k
    assert got.shape == want.shape
    assert float((got - want).abs().max()) <= 5e-3 * float(want.abs().max())


def test_generator_audio_host_logic_vs_oracle():
    import spade_emulator as EM
    sd = SF.make_state_dict("generator_audio", seed=1)
    inp = SF.make_inputs(32, 24, seed=4)
    m = EM.EmulatedGeneratorAudio()
    m.load_state_dict(sd, strict=True)
    got = m(inp["i_src"], inp["i_raw"], inp["audio"])
    want = SO.generator_audio(sd, inp["i_src"], inp["i_raw"], inp["audio"])
    assert float((got - want).abs().max()) <= 5e-3 * float(want.abs().max())


@pytest.mark.parametrize("kind", ["generator", "generator_audio"])
def test_identity_cache_equals_full_forward(kind):
    """G.refine(G.encode_identity(I_src), frame) = G(I_src, frame) exactly, with the IdEncoder and the identity-conditioned
    mlp_shared convs computed once per clip (18 of them for Generator, 12 for Generator_audio)"""
    import spade_emulator as EM
    m = EM.EmulatedGenerator() if kind == "generator" else EM.EmulatedGeneratorAudio()
    m.load_state_dict(SF.make_state_dict(kind, seed=0), strict=True)
    i0, i1 = SF.make_inputs(16, 24, seed=1), SF.make_inputs(16, 24, seed=2)
    extra = (lambda i: (i["audio"],)) if kind == "generator_audio" else (lambda i: ())
    ident = m.encode_identity(i0["i_src"])
    first = m.refine(ident, i0["i_raw"], *extra(i0))
    assert len(ident.actv) == (18 if kind == "generator" else 12)
    second = m.refine(ident, i1["i_raw"], *extra(i1))
    assert torch.equal(first, m(i0["i_src"], i0["i_raw"], *extra(i0)))
    assert torch.equal(second, m(i0["i_src"], i1["i_raw"], *extra(i1)))
    with pytest.raises(RuntimeError):
        m.refine(ident, torch.zeros(1, 3, 32, 24), *extra(i0))


@pytest.mark.parametrize("out_w,out_h,up,down", [(128, 3, 0, 0), (24, 16, 0, 0), (7, 9, 0, 0), (20, 12, 1, 0), (12, 8, 0, 1), (256, 2, 0, 0)])
def test_one_load_three_taps_rule_equals_tap_by_tap_gather(out_w, out_h, up, down):
    """Model of the stride-1 gather of csrc/spade_conv.cu: per (ky, chunk) a thread loads ITS pixel once and stores it into
    the three kx operand tiles at rows r+1, r, r-1 (same image row only); its own row gets zeros where tap kx falls outside
    the image; the two pixels beyond the tile's ends come from the halo lanes.  The three tiles must equal the tap-by-tap
    gather (src_pixel) for every tile, width (tiles spanning several image rows, ragged last tile) and resize mode."""
    import spade_emulator as EM
    from sahs_b200 import spade as SP
    in_h, in_w = (out_h << down) >> up, (out_w << down) >> up
    img = torch.arange(1, in_h * in_w + 1, dtype=torch.float32).reshape(in_h, in_w)      # pixel id; 0 = padding
    P = out_h * out_w

    def px(oy, ox, ky, kx):          # tap-by-tap reference value
        ok, iy, ix = EM.src_pixel(SP.MODE_S1, torch.tensor(oy), torch.tensor(ox), ky, kx, in_h, in_w, out_h, out_w, up, down)
        return float(img[int(iy), int(ix)]) if bool(ok) else 0.0

    for tile in range((P + 127) // 128):
        for ky in range(3):
            tiles3 = torch.full((3, 128), float("nan"))
            for r in range(128):                                  # the 128 gather threads
                p = tile * 128 + r
                live = p < P
                oy, ox = (p // out_w, p % out_w) if live else (0, 0)
                v = px(oy, ox, ky, 1) if live else 0.0
                for kx in range(3):
                    dr, dx = r - kx + 1, ox - kx + 1
                    if kx == 1 or (live and 0 <= dr < 128 and 0 <= dx < out_w):
                        assert torch.isnan(tiles3[kx, dr]), "two writers for one operand row"
                        tiles3[kx, dr] = v
                    if kx != 1 and (not live or ox + kx - 1 < 0 or ox + kx - 1 >= out_w):
                        assert torch.isnan(tiles3[kx, r]), "two writers for one operand row"
                        tiles3[kx, r] = 0.0
            for kx, drow in ((0, 0), (2, 127)):                   # the two halo lanes
                p = tile * 128 + drow
                if p < P:
                    oy, ox = p // out_w, p % out_w
                    if (ox >= 1) if kx == 0 else (ox + 1 < out_w):
                        assert torch.isnan(tiles3[kx, drow]), "two writers for one operand row"
                        tiles3[kx, drow] = px(oy, ox, ky, kx)
            for kx in range(3):
                for r in range(128):
                    p = tile * 128 + r
                    if p < P:                                     # rows beyond the last pixel are never stored
                        assert not torch.isnan(tiles3[kx, r]), (tile, ky, kx, r, "operand row never written")
                        assert float(tiles3[kx, r]) == px(p // out_w, p % out_w, ky, kx), (tile, ky, kx, r)
