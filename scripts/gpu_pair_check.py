"""A/B check of the CTA-pair field kernel against the single-CTA kernel (same process, SAHS_FIELD_PAIR toggled).

    python scripts/gpu_pair_check.py [audio/person_2_auto]

Prints the status word if a launch fails, bitwise differences between the two kernels on ragged and full sizes,
and the timing of both on one fine-level launch of a 512x512 frame.  Diagnostic only.
"""
import os
import sys
import time

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "tests"), os.path.join(REPO, "sahs-deformable-nerf_b200")):
    sys.path.insert(0, p)

import sahs_fixtures as FX  # noqa: E402
from oracle import sahs_oracle as O  # noqa: E402
import sahs_b200  # noqa: E402
from sahs_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")


def main(cfg_name):
    cfg = FX.load_cfg(cfg_name)
    ospec = O.spec_from_cfg(cfg)
    sd = FX.make_state_dict(ospec, seed=42, dense=True)
    model = getattr(sahs_b200.models, cfg.models.mask.type)(cfg)
    model.load_state_dict(sd, strict=True)
    model = model.to(dev)
    fr = FX.make_frame_inputs(ospec, 8, 8, seed=1)
    pose, drv = fr["pose"], O.driving_vector(sd, ospec, fr["driving"])
    pcode = O.pose_code(pose).to(dev)
    drv = drv.to(dev)
    fc = model.frame_constants("fine", drv, pcode)

    def run(pair, ro, rd, z):
        os.environ["SAHS_FIELD_PAIR"] = "1" if pair else "0"
        out = model.field("fine", ro, rd, z, drv, pcode, frame_const=fc)
        torch.cuda.synchronize()
        return out

    gen = torch.Generator().manual_seed(3)
    for R, S in ((1, 64), (3, 64), (5, 128), (64, 128), (2048, 128), (4099, 64)):
        ro = torch.zeros(R, 3); ro[:, 2] = 0.78
        rd = torch.randn(R, 3, generator=gen) * 0.2; rd[:, 2] = -1
        z = torch.linspace(0.48, 1.08, S).expand(R, S).contiguous() + torch.rand(R, 1, generator=gen) * 0.01
        ro, rd, z = ro.to(dev), rd.to(dev), z.to(dev)
        try:
            a = run(False, ro, rd, z)
            t0 = time.time()
            b = run(True, ro, rd, z)
            dt = time.time() - t0
        except Exception as e:  # noqa: BLE001
            print(f"R={R} S={S}: FAILED {e}; status {ops.field_status() if False else ''}")
            import ctypes as C
            out = (C.c_int * 4)()
            sahs_b200.lib._lib.sahs_field_status(out)
            print("status word:", list(out))
            return 1
        nd = int((a != b).sum())
        print(f"R={R} S={S} ({R * S} points, {(R * S + 127) // 128} tiles): differing values {nd} of {a.numel()}, "
              f"max abs {float((a - b).abs().max()):.3e}, nan {int(torch.isnan(b).sum())}, pair launch {dt * 1e3:.1f} ms")
    # timing: one fine-level launch of a 512^2 frame
    R, S = 262144, 128
    ro = torch.zeros(R, 3, device=dev); ro[:, 2] = 0.78
    rd = torch.randn(R, 3, device=dev) * 0.1; rd[:, 2] = -1
    z = torch.linspace(0.48, 1.08, S, device=dev).expand(R, S).contiguous()
    for pair in (False, True, False, True):
        run(pair, ro, rd, z)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(2):
            model.field("fine", ro, rd, z, drv, pcode, frame_const=fc)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 2
        tf = R * S * 1.732864e6 / (ms * 1e-3) / 1e12
        print(f"pair={int(pair)}: fine launch {ms:.2f} ms  {tf:.1f} TFLOP/s algorithmic ({tf / 1350:.3f} of sustained peak)")
    print("status", ops.field_status())
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1] if len(sys.argv) > 1 else "audio/person_2_auto"))
