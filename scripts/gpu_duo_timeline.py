"""Timeline of the two-tile ("duo") render kernel (SAHS_DBG_PROF_DUO): for the 4th tile of both sets of cluster 0,
per pass: when the workers signalled their operand, when the issuer picked the pass and finished issuing it, when the
accumulator woke the workers, and how long the worker phase took.  Production instantiation, true speed."""
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


def main(R, stages=0):
    cfg = FX.load_cfg("audio/person_2_auto")
    ospec = O.spec_from_cfg(cfg)
    sd = FX.make_state_dict(ospec, seed=42, dense=True, trained_like=True)
    model = getattr(sahs_b200.models, cfg.models.mask.type)(cfg)
    model.load_state_dict(sd, strict=True)
    model = model.to(dev)
    fr = FX.make_frame_inputs(ospec, 8, 8, seed=1)
    drv = O.driving_vector(sd, ospec, fr["driving"]).to(dev)
    pcode = O.pose_code(fr["pose"]).to(dev)
    fc = model.frame_constants("fine", drv, pcode)
    S = 128
    if stages:
        R = R - 1          # odd point count asks the kernel for per-stage issuer events
        S = 127
    ro = torch.zeros(R, 3, device=dev); ro[:, 2] = 0.78
    rd = torch.randn(R, 3, device=dev) * 0.1; rd[:, 2] = -1
    z = torch.linspace(0.48, 1.08, S, device=dev).expand(R, S).contiguous()
    dbg = torch.zeros(128, 256, device=dev)
    for _ in range(2):
        dbg.zero_()
        model.field("fine", ro, rd, z, drv, pcode, frame_const=fc, debug=dbg, debug_pass=96)
        torch.cuda.synchronize()
    buf = dbg.cpu().numpy().view(np.int64).reshape(-1)
    w = [events(buf, 0), events(buf, 1)]
    iss = events(buf, 2)
    if len(w[0]) == 0:
        print("no events (fewer than 4 tiles per set?)")
        return
    t0 = int(min(x[0, 1] for x in w if len(x)))
    if stages:
        # per-stage issuer events of set 0: wait for the weights
        pre = {int(t) - 25000: int(c) - t0 for t, c in iss if 25000 <= t < 26000}
        post = {int(t) - 20000: int(c) - t0 for t, c in iss if 20000 <= t < 21000}
        print(" st | before wait | after wait | waited | since previous stage's wait end")
        prev = None
        for st in sorted(post):
            print(f"{st:4d} | {pre.get(st, -1):8d} | {post[st]:8d} | {post[st] - pre.get(st, post[st]):6d} | {(post[st] - prev) if prev is not None else 0:6d}")
            prev = post[st]
    for k in (0, 1):
        ev = w[k]
        if len(ev) == 0:
            print(f"\nset {k}: no events (set without work)")
            continue
        rel = lambda c: int(c) - t0
        sig = [rel(c) for t, c in ev if t == 1]
        acc = [(int(t) - 100000, rel(c)) for t, c in ev if t >= 100000]
        start, end = rel(ev[0, 1]), rel(ev[-1, 1])
        pick = [(int(t) - 10000 - 1000 * k, rel(c)) for t, c in iss if 10000 + 1000 * k <= t < 11000 + 1000 * k]
        done = [rel(c) for t, c in iss if 30000 + 1000 * k <= t < 31000 + 1000 * k]
        print(f"\nset {k}: tile {start} .. {end} = {end - start} cycles; {len(sig)} signals, {len(acc)} wakes, {len(pick)} passes seen by the issuer")
        print(" pass  tag st0 |   signal |  picked (+wait) | issued (+dur) | acc wake (+drain) | mma phase | worker phase")
        tw = ti = td = te = 0
        for i in range(min(len(sig), len(acc), len(pick), len(done))):
            s_, (tag, a_), (st0, p_), d_ = sig[i], acc[i], pick[i], done[i]
            nxt = sig[i + 1] if i + 1 < len(sig) else end
            tw += p_ - s_; ti += d_ - p_; td += a_ - d_; te += nxt - a_
            print(f"{i:5d} {tag:5d} {st0:3d} | {s_:8d} | {p_:8d} ({p_ - s_:5d}) | {d_:8d} ({d_ - p_:5d}) | {a_:8d} ({a_ - d_:5d}) | {a_ - s_:6d} | {nxt - a_:6d}")
        print(f"sums: signal->picked {tw}, issue {ti}, issue end->acc wake {td}, worker phases {te}, first signal at {sig[0] - start} after tile start")


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 65536, int(sys.argv[2]) if len(sys.argv) > 2 else 0)
