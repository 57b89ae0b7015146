"""What bounds spade_conv_kernel?  Times one large layer (cin -> cout at HxW, stride 1) with parts of the kernel switched
off through SAHS_CONV_DBG (results are then wrong on purpose; only the time matters):
  1 no gather loads, 2 no operand stores / proxy fence, 4 no MMAs, 8 16-byte weight copies, 16 no output stores.
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


def run(cin, cout, h, w, flags):
    g = torch.Generator().manual_seed(1)
    wt = torch.randn(cout, cin, 3, 3, generator=g) * 0.02
    p = SP._pack_conv(wt.to(dev), torch.zeros(cout, device=dev))
    x = torch.randn(h, w, cin, generator=g).half().to(dev)
    out = {}
    for f in flags:
        os.environ["SAHS_CONV_DBG"] = str(f)
        for _ in range(3):
            gen._conv(p, x, h, w, 0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(10):
            gen._conv(p, x, h, w, 0)
        e1.record()
        torch.cuda.synchronize()
        out[f] = e0.elapsed_time(e1) / 10
    os.environ["SAHS_CONV_DBG"] = "0"
    fl = 2.0 * h * w * cout * 9 * cin
    print(f"{cin}->{cout} {h}x{w}: " + "  ".join(f"[{f}] {t * 1e3:.0f} us ({fl / t / 1e9:.0f} TF/s)" for f, t in out.items()))


FLAGS = [0, 1, 2, 3, 4, 8, 16, 1 | 2 | 8 | 16, 1 | 2 | 4 | 8 | 16]
run(128, 128, 512, 512, FLAGS)
run(64, 128, 512, 512, FLAGS)
run(256, 256, 128, 128, FLAGS)
