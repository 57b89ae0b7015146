"""Debug: per-point gradient w.r.t. the warped point (ours, from the gradient tape) vs oracle autograd."""
import os, sys
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "tests"), os.path.join(REPO, "sahs-deformable-nerf_b200")):
    sys.path.insert(0, p)
import sahs_fixtures as FX
from oracle import sahs_oracle as O
import sahs_b200 as sahs
from sahs_b200 import train as TR

DEV = "cuda:0"
cfg = FX.load_cfg("audio/person_2_auto")
spec = O.spec_from_cfg(cfg)
sd = FX.make_state_dict(spec, seed=42, dense=True)
fr = FX.make_frame_inputs(spec, 8, 8, seed=4)
gen = torch.Generator().manual_seed(3)
n = 1024
xyz = (torch.rand(n, 3, generator=gen) * 2 - 1) * 0.35
dirs = torch.randn(n, 3, generator=gen) * 0.3 + torch.tensor([0, 0, -1.0])
drv = O.driving_vector(sd, spec, fr["driving"]).detach()
gout = torch.randn(n, 16, generator=gen) * 1e-3
# oracle
sd_ref = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
raw, inter = O.field_forward(sd_ref, spec, "fine", xyz, dirs, drv, fr["pose"], return_intermediates=True)
inter["mapped"].retain_grad(); inter["amb"].retain_grad(); inter["emb"].retain_grad(); inter["dx"].retain_grad()
for k, v in inter.items():
    if v.requires_grad and not v.is_leaf: v.retain_grad()
(raw * gout).sum().backward()
dmap_ref = inter["mapped"].grad
# ours
model = sahs.AudioFaceModel(cfg); model.load_state_dict(sd); model = model.to(DEV)
stash = {}
orig = TR._weight_grads
def spy(model_, level, lay, tx, td, cvec):
    stash.update(lay=lay, td=td.clone(), tx=tx.clone())
    return orig(model_, level, lay, tx, td, cvec)
TR._weight_grads = spy
z0 = torch.zeros(n, 1, device=DEV)
pcode = model.pose_code(fr["pose"].to(DEV))
raw_g = model.field("fine", xyz.to(DEV), dirs.to(DEV), z0, drv.to(DEV).requires_grad_(True), pcode)
(raw_g.reshape(n, 16) * gout.to(DEV)).sum().backward()
torch.cuda.synchronize()
scale = 16.0 / float(gout.abs().max())
lay, td = stash["lay"], stash["td"].float().cpu() / scale
def cmp(name, ours, ref):
    a, b = ours.double().reshape(-1), ref.double().reshape(-1)
    print(f"  {name:10s} cos {float((a*b).sum()/(a.norm()*b.norm()+1e-30)):.5f} ratio {float(a.norm()/(b.norm()+1e-30)):.4f}")
lk = lambda h: torch.where(h > 0, torch.ones_like(h), torch.full_like(h, 0.01))
rl = lambda h: (h > 0).float()
print("per-layer d(pre-activation): ours (tape) vs oracle autograd")
for i in range(3, -1, -1):
    ref = torch.cat((inter[f"dir{i}"].grad * lk(inter[f"dir{i}"]), inter[f"seg{i}"].grad * lk(inter[f"seg{i}"])), 1).detach()
    cmp(f"head{i}", td[:, lay["td_hh"] + i*256: lay["td_hh"] + (i+1)*256], ref)
cmp("feat", td[:, lay["td_feat"]:lay["td_feat"]+256], inter["feat"].grad)
for i in range(7, -1, -1):
    h = inter[f"trunk{i}"]
    cmp(f"trunk{i}", td[:, lay["td_th"] + i*256: lay["td_th"] + (i+1)*256], (h.grad * lk(h)).detach())
for i in range(5, -1, -1):
    hw, hh = inter[f"warp{i}"], inter[f"hyper{i}"]
    ref = torch.cat(((hw.grad * rl(hw)), (hh.grad * rl(hh))), 1).detach()
    cmp(f"warp|hyp{i}", td[:, lay["td_wh"] + i*192: lay["td_wh"] + (i+1)*192], ref)
dpre = td[:, lay["td_final"]:lay["td_final"] + 3]
damb = td[:, lay["td_final"] + 3:lay["td_final"] + 5]
t = inter["dx"].detach()
dpre_ref = inter["dx"].grad * 1.0   # grad wrt tanh output
dpre_ref = dmap_ref * (1 - t * t)
print("raw fwd err", float((raw_g.reshape(n,16).cpu() - raw.detach()).abs().max()))
for k in range(3):
    a, b = dpre[:, k].double(), dpre_ref[:, k].double()
    print(f"dpre[{k}]: cos {float((a*b).sum()/(a.norm()*b.norm())):.5f} ratio {float(a.norm()/b.norm()):.4f} max|ref| {float(b.abs().max()):.3e} max err {float((a-b).abs().max()):.3e}")
for k in range(2):
    a, b = damb[:, k].double(), inter["amb"].grad[:, k].double()
    print(f"damb[{k}]: cos {float((a*b).sum()/(a.norm()*b.norm())):.5f} ratio {float(a.norm()/b.norm()):.4f}")
# oracle decomposition: grid part vs PE part of d mapped
sd2 = {k: v.clone() for k, v in sd.items()}
m = inter["mapped"].detach().clone().requires_grad_(True)
emb = O.grid_sample_trilinear(sd2["spatial_embeddings"], m)
(emb * inter["emb"].grad).sum().backward()
dgrid_ref = m.grad
print("oracle: |d mapped| grid part", float(dgrid_ref.norm()), "total", float(dmap_ref.norm()), "PE part", float((dmap_ref - dgrid_ref).norm()))
ours_dmap = dpre / (1 - t * t)
e = ours_dmap - dmap_ref
print("ours-ref: err norm", float(e.norm()), " corr(err, grid part)", float((e*dgrid_ref).sum()/(e.norm()*dgrid_ref.norm()+1e-30)), " corr(err, PE part)", float((e*(dmap_ref-dgrid_ref)).sum()/(e.norm()*(dmap_ref-dgrid_ref).norm()+1e-30)))
worst = e.abs().max(dim=1)[0].topk(5)[1]
for i in worst:
    print(int(i), "ours", ours_dmap[i].tolist(), "ref", dmap_ref[i].tolist(), "grid part", dgrid_ref[i].tolist(), "mapped", inter["mapped"][i].tolist())
