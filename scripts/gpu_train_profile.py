"""Where does a training step spend its GPU time?  (torch.profiler kernel table; diagnostic only)"""
import os, sys, time
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
sd = FX.make_state_dict(spec, seed=42, dense=True, trained_like=os.environ.get("SAHS_FIXTURE", "trained") != "dense")
H = W = 512
fr = FX.make_frame_inputs(spec, H, W, seed=0)
model = sahs_b200.AudioFaceModel(cfg); model.load_state_dict(sd); model = model.to(DEV)
opt = sahs_b200.FlatAdam(model.parameters(), lr=5e-4)
pose = fr["pose"].to(DEV)
with torch.no_grad():
    ro, rd = sahs_b200.get_ray_bundle(H, W, fr["intrinsics"], pose)
ro, rd = ro.reshape(-1, 3), rd.reshape(-1, 3)
bg = fr["background"].view(-1, 15).to(DEV)
maskf = fr["mask"].view(-1, 12).float().to(DEV)
target = torch.rand(H * W, 3, device=DEV)
drv = fr["driving"].to(DEV)
mask_i32 = fr["mask"].view(-1, 12).to(torch.int32).to(DEV).contiguous()
prob = torch.ones(12, device=DEV)
it = [0]
def step():
    it[0] += 1
    sel = sahs_b200.weighted_sample(mask_i32, prob, 2048, seed=it[0])
    out = sahs_b200.run_one_iter_of_nerf(H, W, 1200.0, model, ro[sel], rd[sel], cfg, mode="train", driving=drv, pose=pose,
                                         background_prior=bg[sel])
    loss, _ = sahs_b200.stage1_loss(out[0], out[3], target[sel], maskf[sel])
    opt.zero_grad(set_to_none=True); loss.backward(); opt.step()
for _ in range(3): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5): step()
torch.cuda.synchronize()
print("wall ms/step", (time.perf_counter() - t0) / 5 * 1e3)
if os.environ.get("SAHS_NO_TORCH_PROFILER") == "1":   # under ncu: CUPTI cannot serve both
    sys.exit(0)
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3): step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=28, max_name_column_width=70))
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=30, max_name_column_width=60))
