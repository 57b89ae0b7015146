"""Compositing forward alone at the frame's shapes (262,144 rays x 64 and x 128 samples), for an `ncu --set full`
capture of composite_fwd_kernel<4> (inside the bench its counters come back as nan: it is the last launch of the frame)."""
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "tests"), os.path.join(REPO, "sahs-deformable-nerf_b200")):
    sys.path.insert(0, p)
import sahs_b200  # noqa: E402,F401
from sahs_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
R = 262144
gen = torch.Generator(device=dev).manual_seed(1)
for S in (64, 128):
    raw = torch.randn(R, S, 16, device=dev, generator=gen)
    raw[..., -1] = torch.randn(R, S, device=dev, generator=gen) * 10 + 3
    z = torch.sort(torch.rand(R, S, device=dev, generator=gen) * 0.6 + 0.48, -1)[0]
    rd = torch.randn(R, 3, device=dev, generator=gen) * 0.1
    rd[:, 2] = -1
    bg = torch.rand(R, 15, device=dev, generator=gen)
    for _ in range(3):
        out = ops.composite_fwd(raw, z, rd, None, bg, True, False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        out = ops.composite_fwd(raw, z, rd, None, bg, True, False)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"S={S}: {ms:.4f} ms per launch, {R * (72 * S + 144) / ms / 1e6:.0f} GB/s algorithmic")
    del raw, z, out
