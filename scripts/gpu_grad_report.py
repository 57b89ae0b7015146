"""Per-parameter gradient parity report: CUDA training step vs torch autograd through the CPU oracle."""
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "tests"), os.path.join(REPO, "sahs-deformable-nerf_b200")):
    sys.path.insert(0, p)
import sahs_fixtures as FX  # noqa: E402
from oracle import sahs_oracle as O  # noqa: E402
import sahs_b200 as sahs  # noqa: E402

DEV = "cuda:0"
cfg_name = sys.argv[1] if len(sys.argv) > 1 else "audio/person_2_auto"
H, W = 6, 8
cfg = FX.load_cfg(cfg_name)
cfg.nerf.train.perturb, cfg.nerf.train.radiance_field_noise_std = False, 0.0
spec = O.spec_from_cfg(cfg)
sd = FX.make_state_dict(spec, seed=42, dense=True)
fr = FX.make_frame_inputs(spec, H, W, seed=4)
target = torch.rand(H * W, 3, generator=torch.Generator().manual_seed(77))
mask = fr["mask"].view(-1, 12).float()
sd_ref = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
opts = O.opts_from_cfg(cfg, "train")
ro, rd = O.get_ray_bundle(H, W, fr["intrinsics"], fr["pose"])
out_ref = O.run_one_iter(sd_ref, spec, opts, ro, rd, fr["driving"], fr["pose"], fr["background"].view(-1, 15))
loss_ref, _ = sahs.stage1_loss(out_ref[0], out_ref[3], target, mask)
loss_ref.backward()
model = getattr(sahs.models, cfg.models.mask.type)(cfg)
model.load_state_dict(sd)
model = model.to(DEV)
pose = fr["pose"].to(DEV)
with torch.no_grad():
    ro_g, rd_g = sahs.get_ray_bundle(H, W, fr["intrinsics"], pose)
out = sahs.run_one_iter_of_nerf(H, W, fr["intrinsics"][0], model, ro_g, rd_g, cfg, mode="train",
                                driving=fr["driving"].to(DEV), pose=pose,
                                background_prior=fr["background"].view(-1, 15).to(DEV), inHead=fr["mask"].to(DEV))
loss, _ = sahs.stage1_loss(out[0], out[3], target.to(DEV), mask.to(DEV))
loss.backward()
torch.cuda.synchronize()
print("loss", float(loss), "ref", float(loss_ref))
for name, p in model.named_parameters():
    g = p.grad.detach().cpu().double().reshape(-1)
    r = sd_ref[name].grad.double().reshape(-1)
    rel = float((g - r).abs().max()) / max(float(r.abs().max()), 1e-30)
    cos = float(torch.dot(g, r) / (g.norm() * r.norm() + 1e-30))
    ratio = float(g.norm() / (r.norm() + 1e-30))
    print(f"{name:45s} max|ref| {float(r.abs().max()):.3e}  rel {rel:.3e}  cos {cos:.5f}  norm ratio {ratio:.4f}")
