"""Tiny end-to-end case (render, sampler and one training step on 4x8 rays): a quick liveness check on a GPU box."""
import os, sys
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "tests"), os.path.join(REPO, "sahs-deformable-nerf_b200")):
    sys.path.insert(0, p)
import sahs_fixtures as FX
from oracle import sahs_oracle as O
import sahs_b200
DEV = "cuda:0"
cfg = FX.load_cfg("audio/person_2_auto")
spec = O.spec_from_cfg(cfg)
sd = FX.make_state_dict(spec, seed=42, dense=True)
H, W = 4, 8
fr = FX.make_frame_inputs(spec, H, W, seed=0)
model = sahs_b200.AudioFaceModel(cfg); model.load_state_dict(sd); model = model.to(DEV)
pose = fr["pose"].to(DEV)
with torch.no_grad():
    ro, rd = sahs_b200.get_ray_bundle(H, W, fr["intrinsics"], pose)
    cfg.nerf.validation.perturb = False
    out = sahs_b200.run_one_iter_of_nerf(H, W, fr["intrinsics"][0], model, ro, rd, cfg, mode="validation",
                                         driving=fr["driving"].to(DEV), pose=pose,
                                         background_prior=fr["background"].view(-1, 15).to(DEV), inHead=fr["mask"].to(DEV))
    torch.cuda.synchronize()
    print("render ok", float(out[3].abs().sum()))
    idx = sahs_b200.weighted_sample(fr["mask"].view(-1, 12).to(DEV), torch.ones(12, device=DEV), 8, seed=1)
    print("sampler ok", sorted(idx.tolist()))
out = sahs_b200.run_one_iter_of_nerf(H, W, fr["intrinsics"][0], model, ro, rd, cfg, mode="train", driving=fr["driving"].to(DEV),
                                     pose=pose, background_prior=fr["background"].view(-1, 15).to(DEV))
loss, _ = sahs_b200.stage1_loss(out[0], out[3], torch.rand(H * W, 3, device=DEV), fr["mask"].view(-1, 12).float().to(DEV))
loss.backward()
torch.cuda.synchronize()
print("train step ok", float(loss))
