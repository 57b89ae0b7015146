"""The bf16 build (`make BF16=1` -> lib/libsahs_b200_bf16.so: trunk and heads take bf16 operands on the render path) is a
second library selected with SAHS_B200_LIB; a process binds one library, so the check runs in a child process."""
import json
import os
import subprocess
import sys

import pytest

import sahs_fixtures as FX

pytestmark = pytest.mark.gpu
LIB = os.path.join(FX.REPO, "sahs-deformable-nerf_b200", "lib", "libsahs_b200_bf16.so")

CHILD = r"""
import json, sys, torch
sys.path[:0] = [%(repo)r, %(repo)r + "/tests", %(repo)r + "/sahs-deformable-nerf_b200"]
import sahs_b200, sahs_fixtures as FX
from sahs_b200 import ops
from oracle import sahs_oracle as O
dev = torch.device("cuda:0")
out = {"format": ops.operand_format()}
for name in ("audio/person_2_auto", "expression/person_2"):
    cfg = FX.load_cfg(name)
    cfg.nerf.validation.perturb = False
    spec = O.spec_from_cfg(cfg)
    sd = FX.make_state_dict(spec, seed=42, dense=True, trained_like=True)
    H = W = 12
    fr = FX.make_frame_inputs(spec, H, W, seed=0, pose_z=FX.probe_pose_z(spec))
    model = getattr(sahs_b200.models, cfg.models.mask.type)(cfg)
    model.load_state_dict(sd)
    model = model.to(dev)
    pose = fr["pose"].to(dev)
    with torch.no_grad():
        ro, rd = sahs_b200.get_ray_bundle(H, W, fr["intrinsics"], pose)
        got = sahs_b200.run_one_iter_of_nerf(H, W, fr["intrinsics"][0], model, ro, rd, cfg, mode="validation",
                                             driving=fr["driving"].to(dev), pose=pose,
                                             background_prior=fr["background"].view(-1, 15).to(dev))
        ro_c, rd_c = O.get_ray_bundle(H, W, fr["intrinsics"], fr["pose"])
        ref = O.run_one_iter(sd, spec, O.opts_from_cfg(cfg, "validation"), ro_c, rd_c, fr["driving"], fr["pose"],
                             fr["background"].view(-1, 15))
    rgb, rgb_ref = got[3].reshape(-1, 15).cpu(), ref[3]
    mse = float(((rgb[:, :3] - rgb_ref[:, :3]) ** 2).mean())
    out[name] = {"rgb_maxabs": float((rgb - rgb_ref).abs().max()),
                 "coarse_maxabs": float((got[0].reshape(-1, 15).cpu() - ref[0]).abs().max()),
                 "depth_maxabs": float((got[7].reshape(-1).cpu() - ref[7]).abs().max()),
                 "psnr": 99.0 if mse == 0 else -10.0 * __import__("math").log10(mse),
                 "status": ops.field_status()[0]}
print("RESULT " + json.dumps(out))
"""


def test_bf16_build_renders_within_bf16_tolerance():
    if not os.path.exists(LIB):
        pytest.skip("lib/libsahs_b200_bf16.so not built (make -C sahs-deformable-nerf_b200 BF16=1)")
    env = dict(os.environ, SAHS_B200_LIB=LIB)
    p = subprocess.run([sys.executable, "-c", CHILD % {"repo": FX.REPO}], env=env, capture_output=True, text=True,
                       timeout=600)
    assert p.returncode == 0, p.stderr[-3000:]
    res = json.loads([ln for ln in p.stdout.splitlines() if ln.startswith("RESULT ")][-1][7:])
    d = os.path.join(FX.REPO, "gpurun_out")
    if os.path.isdir(d):
        with open(os.path.join(d, "test_report.txt"), "a") as f:
            f.write("[bf16 build] " + json.dumps(res) + "\n")
    assert res["format"] == "bf16"
    # bf16 operands carry 8 significand bits (fp16: 11).  Measured on B200: rgb/seg 4.9e-3 / 3.1e-3, depth 4.4e-3 / 1.5e-3,
    # PSNR 65 / 71 dB (audio / expression config); the fp16 build's bars are 1e-2 and 50 dB at ~1e-3 measured.
    for name in ("audio/person_2_auto", "expression/person_2"):
        r = res[name]
        assert r["status"] == 0 and r["rgb_maxabs"] < 2e-2 and r["depth_maxabs"] < 2e-2 and r["psnr"] >= 55.0, (name, r)
