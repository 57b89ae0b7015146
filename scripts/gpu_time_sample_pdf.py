"""Times sahs_sample_pdf_merge at the BASELINE ray count (diagnostic).  python scripts/gpu_time_sample_pdf.py"""
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "sahs-deformable-nerf_b200")):
    sys.path.insert(0, p)
from sahs_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
R = 262144
z = ops.coarse_z(R, 64, 0.4838, 1.0838, False, dev)
w = torch.rand(R, 64, device=dev) ** 8
u = torch.rand(R, 64, device=dev)
for name, uu in (("det", None), ("stochastic", u), ("det, general kernel", None), ("stochastic, general kernel", u)):
    os.environ.pop("SAHS_SAMPLE_PDF_GENERIC", None)
    if "general" in name:
        os.environ["SAHS_SAMPLE_PDF_GENERIC"] = "1"
    for _ in range(3):
        ops.sample_pdf_merge(z, w, 64, uu)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(20):
        ops.sample_pdf_merge(z, w, 64, uu)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    byt = R * (64 * 4 * 2 + 64 * 4 + 128 * 4)
    print(f"sample_pdf_merge {name}: {ms:.3f} ms  ({byt / ms / 1e6:.0f} GB/s algorithmic)")
