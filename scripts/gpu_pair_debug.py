"""Debug helper: where do the pair and single-CTA kernels differ? (diagnostic only)"""
import os, sys
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "tests"), os.path.join(REPO, "sahs-deformable-nerf_b200")):
    sys.path.insert(0, p)
import sahs_fixtures as FX
from oracle import sahs_oracle as O
import sahs_b200
DEV = "cuda:0"
cfg_name = sys.argv[1] if len(sys.argv) > 1 else "expression/person_2"
cfg = FX.load_cfg(cfg_name)
spec = O.spec_from_cfg(cfg)
sd = FX.make_state_dict(spec, seed=42, dense=True)
model = getattr(sahs_b200.models, cfg.models.mask.type)(cfg)
model.load_state_dict(sd, strict=True)
model = model.to(DEV)
fr = FX.make_frame_inputs(spec, 8, 8, seed=1)
drv, pose = fr["driving"].to(DEV), fr["pose"].to(DEV)
gen = torch.Generator().manual_seed(11)
for n in (129, 257, 384, 647, 1024, 5000):
    xyz = (torch.rand(n, 3, generator=gen) * 2 - 1) * 0.35
    dirs = torch.randn(n, 3, generator=gen) * 0.3 + torch.tensor([0, 0, -1.0])
    x = torch.cat((xyz, dirs), -1).to(DEV)
    outs = {}
    for pair in ("0", "1", "1"):
        os.environ["SAHS_FIELD_PAIR"] = pair
        with torch.no_grad():
            o = model("fine", x, drv, pose, None)
        torch.cuda.synchronize()
        outs.setdefault(pair, []).append(o)
    a, b, b2 = outs["0"][0], outs["1"][0], outs["1"][1]
    d = (a != b)
    rows = d.any(1).nonzero().flatten().tolist()
    print(f"n={n}: differing rows {len(rows)} of {n}; pair run-to-run equal: {torch.equal(b, b2)}")
    if rows:
        print("   rows (tile:row):", [(r // 128, r % 128) for r in rows[:24]])
        print("   cols differing per row:", [int(d[r].sum()) for r in rows[:24]])
        r = rows[0]
        print("   first row single:", a[r].tolist())
        print("   first row pair  :", b[r].tolist())
        print("   max abs diff", float((a - b).abs().max()))
