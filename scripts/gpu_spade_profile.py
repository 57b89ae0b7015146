"""Stage-II SPADE generator at 512x512 on one GPU (diagnostic): eager and CUDA-graph ms per frame, algorithmic FLOPs,
per-conv CUDA-event times, and the library baselines (the oracle port as cuDNN fp32 / fp16 channels_last on the same GPU).
python scripts/gpu_spade_profile.py [--size 512]"""
import argparse
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "tests"), os.path.join(REPO, "sahs-deformable-nerf_b200")):
    sys.path.insert(0, p)
import spade_fixtures as SF  # noqa: E402
from oracle import spade_oracle as SO  # noqa: E402
from sahs_b200 import spade as SP  # noqa: E402


def timed(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--no-baselines", action="store_true")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    H = W = args.size
    sd = SF.make_state_dict("generator", seed=0)
    inp = SF.make_inputs(H, W, seed=5)
    m = SP.Generator()
    m.load_state_dict(sd, strict=True)
    m = m.to(dev)
    a, b = inp["i_src"].to(dev), inp["i_raw"].to(dev)
    m.tally = {}
    m(a, b)
    tally, m.tally = m.tally, None
    ms_eager = timed(lambda: m(a, b))
    g = SP.GraphedGenerator(m, a, b)
    ms_graph = timed(lambda: g(a, b))
    tf = tally["flop"] / 1e12
    print(f"Generator {H}x{W}: {tally['conv_launches']} conv + {tally['other_launches']} helper launches, "
          f"{tf:.3f} TFLOP per frame")
    print(f"  eager {ms_eager:.3f} ms ({tf / ms_eager * 1e3:.0f} TFLOP/s)   one CUDA graph {ms_graph:.3f} ms "
          f"({tf / ms_graph * 1e3:.0f} TFLOP/s)")
    # ---- per-conv times (CUDA events around each launch, eager) ----
    rows = []
    orig = m._conv

    def wrapped(p, x, oh, ow, mode, **kw):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = orig(p, x, oh, ow, mode, **kw)
        e1.record()
        if isinstance(p, SP._PackedT2):             # four parity-class launches
            for cls, c in enumerate(p.classes):
                nt = c.taps[0] * c.taps[1]
                rows.append((e0, e1, f"{p.cin:>3}->{p.cout:<3} {oh:>3}x{ow:<3} transposed, parity classes (1/2/2/4 taps) "
                                     f"ntile {c.ntile}x{c.ntiles}", 2.0 * x.shape[0] * x.shape[1] * p.cout * nt * p.cin))
            return out
        n_out = 2 * p.cout if kw.get("spade") is not None else p.cout
        taps = 2.25 if mode == SP.MODE_T2 else 9.0
        rows.append((e0, e1, f"{p.cin:>3}->{n_out:<3} {oh:>3}x{ow:<3} mode {SP.MODE_FIRST if p.first else mode} "
                             f"up{kw.get('up', 0)} dn{kw.get('down', 0)} ntile {p.ntile}x{p.ntiles}"
                             f"{' spade' if kw.get('spade') is not None else ''}", 2.0 * oh * ow * n_out * taps * p.cin))
        return out

    m._conv = wrapped
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        m(a, b)
        torch.cuda.synchronize()
    m._conv = orig
    kern = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    kern.sort(key=lambda e: e.time_range.start)
    convs = [e for e in kern if "spade_conv_kernel" in e.name]
    others = {}
    for e in kern:
        if "spade_conv_kernel" not in e.name:
            t, n = others.get(e.name[:60], (0.0, 0))
            others[e.name[:60]] = (t + e.time_range.elapsed_us() / 1e3, n + 1)
    assert len(convs) == len(rows), (len(convs), len(rows))
    agg = {}
    for ev, (_, _, desc, fl) in zip(convs, rows):
        t, f, n = agg.get(desc, (0.0, 0.0, 0))
        agg[desc] = (t + ev.time_range.elapsed_us() / 1e3, f + fl, n + 1)
    tot = sum(v[0] for v in agg.values())
    print(f"  conv kernel durations (CUPTI, one forward; sum {tot:.3f} ms = {tf / tot * 1e3:.0f} TFLOP/s):")
    for desc, (t, f, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f"    {t:7.3f} ms  x{n:<2} {f / t / 1e9:7.0f} TFLOP/s  {desc}")
    print("  other kernels:")
    for name, (t, n) in sorted(others.items(), key=lambda kv: -kv[1][0]):
        print(f"    {t:7.3f} ms  x{n:<3} {name}")
    if args.no_baselines:
        return
    # ---- library baselines: the reference algorithm through cuDNN on this GPU ----
    sdd = {k: v.to(dev) for k, v in sd.items()}
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    with torch.no_grad():
        ref = SO.generator(sdd, a, b)
        ms_fp32 = timed(lambda: SO.generator(sdd, a, b), n=5, warm=2)
        torch.backends.cudnn.allow_tf32 = True
        ms_tf32 = timed(lambda: SO.generator(sdd, a, b), n=5, warm=2)
        sdh = {k: (v.half() if v.is_floating_point() else v) for k, v in sdd.items()}
        ah, bh = a.half().contiguous(memory_format=torch.channels_last), b.half().contiguous(memory_format=torch.channels_last)
        torch.backends.cudnn.benchmark = True
        ms_fp16 = timed(lambda: SO.generator(sdh, ah, bh), n=5, warm=3)
        out = m(a, b)
    err, rng = float((out - ref).abs().max()), float(ref.abs().max())
    print(f"  oracle port on this GPU (cuDNN): fp32 {ms_fp32:.2f} ms, TF32 {ms_tf32:.2f} ms, fp16 channels_last {ms_fp16:.2f} ms")
    print(f"  ours vs cuDNN fp32: max-abs {err:.2e} of range {rng:.2f}")


if __name__ == "__main__":
    main()
