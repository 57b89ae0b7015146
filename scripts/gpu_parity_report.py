"""Prints the end-to-end parity numbers (CUDA path vs golden vectors frozen from the live reference)."""
import math
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

DEV = "cuda:0"
for name in ("e2e_audio_val", "e2e_expr2_val", "e2e_audio_train_stoch"):
    g = np.load(os.path.join(REPO, "tests", "golden", name + ".npz"))
    cfg = FX.load_cfg(str(g["cfg_name"]))
    spec = O.spec_from_cfg(cfg)
    sd = FX.make_state_dict(spec, seed=42, dense=True)
    model = getattr(sahs_b200.models, cfg.models.mask.type)(cfg)
    model.load_state_dict(sd)
    model = model.to(DEV)
    H, W, mode = int(g["H"]), int(g["W"]), str(g["mode"])
    fr = FX.make_frame_inputs(spec, H, W, seed=int(g["seed"]), pose_z=float(g["pose_z"]))
    node = getattr(cfg.nerf, mode)
    draws = None
    if int(g["stochastic"]):
        node.perturb, node.radiance_field_noise_std = True, 0.1
        draws = {k: torch.from_numpy(g["draw_" + k]).to(DEV) for k in ("t_rand", "noise_c", "u", "noise_f")}
    else:
        node.perturb, node.radiance_field_noise_std = False, 0.0
    pose = fr["pose"].to(DEV)
    with torch.no_grad():
        ro, rd = sahs_b200.get_ray_bundle(H, W, fr["intrinsics"], pose)
        out = sahs_b200.run_one_iter_of_nerf(H, W, fr["intrinsics"][0], model, ro, rd, cfg, mode=mode,
                                             driving=fr["driving"].to(DEV), pose=pose,
                                             background_prior=fr["background"].view(-1, 15).to(DEV),
                                             inHead=fr["mask"].to(DEV), _draws=draws)
    names = ["rgb_c", "disp_c", "acc_c", "rgb_f", "disp_f", "acc_f", "w_last_f", "depth_f"]
    line = [name]
    for n, o in zip(names, out):
        ref = torch.from_numpy(g["ref_" + n])
        o = o.reshape(ref.shape).cpu()
        err = float((o - ref).abs().max())
        s = f"{n} {err:.2e}"
        if n.startswith("rgb"):
            mse = float(((o[:, :3] - ref[:, :3]) ** 2).mean())
            s += f" (psnr {(-10 * math.log10(mse)) if mse > 0 else 99:.1f} dB)"
        line.append(s)
    print(" | ".join(line))
