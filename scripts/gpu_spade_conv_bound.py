"""What bounds spade_conv_kernel?  Times one large layer (cin -> cout at HxW, stride 1) with parts of the kernel switched
off through SAHS_CONV_DBG (results are then wrong on purpose; only the time matters):
  1 no gather loads, 2 no operand stores / proxy fence, 4 no MMAs, 8 16-byte weight copies, 16 no output stores,
  32 tap-by-tap gather in the stride-1 modes (the first version of the kernel), 64 no zero-tap skipping (transposed conv).
python scripts/gpu_spade_conv_bound.py"""
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "tests"), os.path.join(REPO, "sahs-deformable-nerf_b200")):
    sys.path.insert(0, p)
from sahs_b200 import spade as SP  # noqa: E402

dev = torch.device("cuda:0")
gen = SP.Generator()


def run(cin, cout, h, w, flags, mode=0):
    g = torch.Generator().manual_seed(1)
    wt = torch.randn(cout, cin, 3, 3, generator=g) * 0.02
    p = SP._pack_conv(wt.to(dev), torch.zeros(cout, device=dev), transposed=(mode == 2))
    x = torch.randn(h, w, cin, generator=g).half().to(dev)
    if mode == 2:
        h, w = 2 * h, 2 * w
    out = {}
    for f in flags:
        os.environ["SAHS_CONV_DBG"] = str(f)
        for _ in range(3):
            gen._conv(p, x, h, w, mode)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(10):
            gen._conv(p, x, h, w, mode)
        e1.record()
        torch.cuda.synchronize()
        out[f] = e0.elapsed_time(e1) / 10
    os.environ["SAHS_CONV_DBG"] = "0"
    fl = 2.0 * h * w * cout * (2.25 if mode == 2 else 9) * cin
    print(f"{cin}->{cout} {h}x{w} mode {mode}: " + "  ".join(f"[{f}] {t * 1e3:.0f} us ({fl / t / 1e9:.0f} TF/s)" for f, t in out.items()))


FLAGS = [0, 32, 1, 2, 3, 4, 8, 16, 1 | 2 | 8 | 16, 1 | 2 | 4 | 8 | 16]
run(128, 128, 512, 512, FLAGS)
run(64, 128, 512, 512, FLAGS)
run(256, 256, 128, 128, FLAGS)
run(128, 128, 256, 256, [0, 64, 1, 4], mode=2)       # transposed conv: 64 = no zero-tap skipping
