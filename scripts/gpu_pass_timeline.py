"""Per-pass timeline of one tile of the field kernel (debug build, SAHS_DBG_PROF): where a CTA's cycles go.

    python scripts/gpu_pass_timeline.py [pair: 0|1] [R] [light: 0|1|2]

light = 1 records per-pass events only (SAHS_DBG_PROF_LIGHT): the per-stage events of the full mode cost about 250
cycles per stage in the issuer's loop and inflate a tile by ~50 %.
light = 2 (SAHS_DBG_PROF_PROD) records the workers' per-pass events from the PRODUCTION instantiation: true speed
(needs a library built with -DSAHS_PROF_PROD=1; the default build records nothing in this mode).

Block 0 records clock64 events of its third tile for worker threads 0 and 255, the MMA issuer and the TMA
producer.  Diagnostic only.
"""
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "tests"), os.path.join(REPO, "sahs-deformable-nerf_b200")):
    sys.path.insert(0, p)

import sahs_fixtures as FX  # noqa: E402
from oracle import sahs_oracle as O  # noqa: E402
import sahs_b200  # noqa: E402

dev = torch.device("cuda:0")


def events(buf, role):
    a = buf[role * 4096:(role + 1) * 4096].reshape(-1, 2)
    n = 0
    while n < len(a) and a[n, 0] != 0:
        n += 1
    return a[:n]


def main(pair, R, light=0):
    os.environ["SAHS_FIELD_PAIR"] = "1" if pair else "0"
    cfg = FX.load_cfg("audio/person_2_auto")
    ospec = O.spec_from_cfg(cfg)
    sd = FX.make_state_dict(ospec, seed=42, dense=True)
    model = getattr(sahs_b200.models, cfg.models.mask.type)(cfg)
    model.load_state_dict(sd, strict=True)
    model = model.to(dev)
    fr = FX.make_frame_inputs(ospec, 8, 8, seed=1)
    drv = O.driving_vector(sd, ospec, fr["driving"]).to(dev)
    pcode = O.pose_code(fr["pose"]).to(dev)
    fc = model.frame_constants("fine", drv, pcode)
    S = 128
    ro = torch.zeros(R, 3, device=dev); ro[:, 2] = 0.78
    rd = torch.randn(R, 3, device=dev) * 0.1; rd[:, 2] = -1
    z = torch.linspace(0.48, 1.08, S, device=dev).expand(R, S).contiguous()
    dbg = torch.zeros(128, 256, device=dev)
    for _ in range(2):
        dbg.zero_()
        model.field("fine", ro, rd, z, drv, pcode, frame_const=fc, debug=dbg, debug_pass={0: 99, 1: 98, 2: 97}[light])
        torch.cuda.synchronize()
    buf = dbg.cpu().numpy().view(np.int64).reshape(-1)
    w0, w255, mma, tma = (events(buf, r) for r in range(4))
    print(f"pair={pair} R={R}: events worker0 {len(w0)} worker255 {len(w255)} mma {len(mma)} tma {len(tma)}")
    if len(w0) == 0:
        return
    t0 = int(w0[0, 1])
    rel = lambda c: int(c) - t0
    print(f"tile duration (worker 0): {rel(w0[-1, 1])} cycles")
    # worker passes: sequence of signal(1) / acc wake (100000+tag)
    sig0 = [rel(c) for t, c in w0 if t == 1]
    acc0 = [(int(t) - 100000, rel(c)) for t, c in w0 if t >= 100000]
    sig255 = [rel(c) for t, c in w255 if t == 1]
    acc255 = [rel(c) for t, c in w255 if t >= 100000]
    wake = [(int(t) - 10000, rel(c)) for t, c in mma if 10000 <= t < 20000]
    full = {int(t) - 20000: rel(c) for t, c in mma if 20000 <= t < 25000}
    pfull = {int(t) - 25000: rel(c) for t, c in mma if 25000 <= t < 30000}
    issued = {int(t) - 30000: rel(c) for t, c in mma if 30000 <= t < 40000}
    free = {int(t) - 40000: rel(c) for t, c in tma if 40000 <= t < 50000}
    nst = max(issued) + 1 if issued else 0
    marks = [(int(t), rel(c)) for t, c in w0 if 200 < t < 300]
    if marks:
        print("mid-tile worker phase (worker 0): " + ", ".join(f"{t}@{c}" for t, c in marks) +
              "  [201 acc wake, 202 fp32 last deformation layer, 203 tanh/exchange, 204 embedding gather, 205 E1 written]")
    print(f"passes: signals {len(sig0)}, acc wakes {len(acc0)}, mma wakes {len(wake)}, stages {nst}")
    if not wake:          # production instantiation: worker-side events only
        tot_m = tot_e = 0
        print(" pass  tag |    sig0  sig255 |    acc0  acc255 | mma_phase  worker phase (next sig - acc)")
        for i in range(min(len(sig0), len(acc0))):
            s0, s255 = sig0[i], sig255[i] if i < len(sig255) else -1
            tag, a0 = acc0[i]
            a255 = acc255[i] if i < len(acc255) else -1
            nxt = sig0[i + 1] if i + 1 < len(sig0) else rel(w0[-1, 1])
            tot_m += a0 - max(s0, s255)
            tot_e += nxt - a0
            print(f"{i:5d} {tag:5d} | {s0:7d} {s255:7d} | {a0:7d} {a255:7d} | {a0 - max(s0, s255):6d}  {nxt - a0:6d}")
        print(f"sum: waiting for MMA (signal -> acc wake) {tot_m}; worker phases {tot_e}; first signal at {sig0[0]}")
        return 0
    tot_epi = tot_mma = tot_wake = 0
    print(" pass  tag  st0 nst | sig0 sig255 | mmawake (+lat) | issue_end | acc0 acc255 | mma_phase  epilogue(next sig - acc)")
    for i in range(min(len(sig0), len(acc0), len(wake))):
        st0 = wake[i][0]
        st1 = wake[i + 1][0] if i + 1 < len(wake) else nst
        s0, s255 = sig0[i], sig255[i] if i < len(sig255) else -1
        mw = wake[i][1]
        ie = issued.get(st1 - 1, -1)
        tag, a0 = acc0[i]
        a255 = acc255[i] if i < len(acc255) else -1
        nxt = sig0[i + 1] if i + 1 < len(sig0) else rel(w0[-1, 1])
        lat = mw - max(s0, s255)
        tot_wake += max(lat, 0)
        tot_mma += a0 - max(s0, s255)
        tot_epi += nxt - a0
        print(f"{i:5d} {tag:5d} {st0:4d} {st1 - st0:3d} | {s0:7d} {s255:7d} | {mw:7d} ({lat:5d}) | {ie:7d} | {a0:7d} {a255:7d} |"
              f" {a0 - max(s0, s255):6d}  {nxt - a0:6d}")
    print(f"sum: waiting for MMA (signal -> acc wake) {tot_mma}, of which issuer wake-up latency {tot_wake}; "
          f"worker phases {tot_epi}")
    # stage table: wait for the weights
    print(" st | slot free (tma) | full seen by mma | issued | wait_for_full = full - prev issued")
    prev = None
    stall = 0
    for st in range(nst):
        f = full.get(st, -1)
        w = (f - prev) if prev is not None else 0
        pf = pfull.get(st)
        extra = f" pfull {pf}" if pf is not None else ""
        if st < 40 or st % 8 == 0:
            print(f"{st:4d} | {free.get(st, -1):8d} | {f:8d} | {issued.get(st, -1):8d} | {w:6d}{extra}")
        prev = issued.get(st, f)
    return 0


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 0, int(sys.argv[2]) if len(sys.argv) > 2 else 16384,
         int(sys.argv[3]) if len(sys.argv) > 3 else 0)
