"""Seeded synthetic fixtures shared by oracle/make_golden.py (build container, with the reference) and
the tests / bench (GPU box, without it).  Everything is generated from torch CPU generators so the same
seed gives the same tensors on both machines; golden files carry checksums to prove it.

Shapes/names follow SURVEY.md Appendix A (the reference's state_dict layout)."""
from __future__ import annotations

import math
import os
import sys
from typing import Dict

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(_HERE)
for p in (REPO, os.path.join(REPO, "sahs-deformable-nerf_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

from oracle import sahs_oracle as O  # noqa: E402

def load_cfg(name: str):
    """Built-in equivalent of one of the reference's shipped YAMLs, e.g. 'audio/person_2_auto'."""
    from sahs_b200.configs import builtin_config
    return builtin_config(name)


def linear_shapes(spec: O.ModelSpec) -> Dict[str, tuple]:
    """name -> shape for every parameter of the live model."""
    s: Dict[str, tuple] = {}
    if spec.use_grid:
        s["spatial_embeddings"] = (1, O.GRID_CH, 32, 32, 32)
    in0 = spec.xyz_dim + O.DRIVING_DIM + O.POSE_PE_DIM

    def skip_stack(prefix, n, hid, skip, in_dim):
        for i in range(n):
            k = in_dim if i == 0 else (hid + in_dim if i == skip else hid)
            s[f"{prefix}.{i}.weight"] = (hid, k)
            s[f"{prefix}.{i}.bias"] = (hid,)

    if spec.use_warp:
        skip_stack("warp_field_mlp.layers_xyz", spec.warp_layers, spec.warp_hidden, spec.warp_skip, in0)
        s["warp_field_mlp.fc_final.weight"] = (3, spec.warp_hidden)
        s["warp_field_mlp.fc_final.bias"] = (3,)
    if spec.use_ambient:
        skip_stack("hyper_sheep_mlp.layers_ambient", spec.hyper_layers, spec.hyper_hidden, spec.hyper_skip, in0)
        s["hyper_sheep_mlp.fc_ambient.weight"] = (spec.amb_dim, spec.hyper_hidden)
        s["hyper_sheep_mlp.fc_ambient.bias"] = (spec.amb_dim,)
    tin = spec.xyz_dim + spec.amb_pe_dim + (O.DRIVING_DIM if spec.trunk_driving else 0) + \
        (O.POSE_PE_DIM if spec.trunk_pose else 0)
    H = spec.trunk_hidden
    for lvl in ("coarse", "fine"):
        p = f"nerf_mlps.{lvl}."
        skip_stack(p + "layers_xyz", spec.trunk_layers, H, spec.trunk_skip, tin)
        s[p + "fc_feat.weight"], s[p + "fc_feat.bias"] = (H, H), (H,)
        s[p + "fc_alpha.weight"], s[p + "fc_alpha.bias"] = (1, H), (1,)
        d0 = H + (spec.dir_dim if spec.use_viewdirs else 0) + (O.GRID_CH if spec.use_grid and spec.use_viewdirs else 0)
        for i in range(4):
            s[p + f"layers_dir.{i}.weight"] = (H // 2, d0 if i == 0 else H // 2)
            s[p + f"layers_dir.{i}.bias"] = (H // 2,)
        s[p + "fc_rgb.weight"], s[p + "fc_rgb.bias"] = (3, H // 2), (3,)
        for i in range(4):
            s[p + f"layers_seg.{i}.weight"] = (H // 2, H if i == 0 else H // 2)
            s[p + f"layers_seg.{i}.bias"] = (H // 2,)
        s[p + "fc_seg.weight"], s[p + "fc_seg.bias"] = (O.NUM_SEG, H // 2), (O.NUM_SEG,)
    if spec.model_type == "AudioFaceModel":
        for i, (co, ci) in zip((0, 2, 4, 6), ((32, 29), (32, 32), (64, 32), (64, 64))):
            s[f"audNet_head.encoder_conv.{i}.weight"] = (co, ci, 3)
            s[f"audNet_head.encoder_conv.{i}.bias"] = (co,)
        s["audNet_head.encoder_fc1.0.weight"], s["audNet_head.encoder_fc1.0.bias"] = (64, 64), (64,)
        s["audNet_head.encoder_fc1.2.weight"], s["audNet_head.encoder_fc1.2.bias"] = (76, 64), (76,)
    return s


def make_state_dict(spec: O.ModelSpec, seed: int = 42, dense: bool = True) -> Dict[str, torch.Tensor]:
    """nn.Linear-like uniform(+-1/sqrt(fan_in)) init from a private generator; `dense` applies the
    SURVEY.md section 7.1 tweak so that density is not ~0 everywhere (otherwise the background sample takes
    all the weight and parity is vacuous)."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    shapes = linear_shapes(spec)
    for name, shp in shapes.items():
        if name == "spatial_embeddings":
            sd[name] = torch.randn(shp, generator=g) * 0.01
            continue
        wname = name.rsplit(".", 1)[0] + ".weight"
        fan_in = math.prod(shapes[wname][1:])
        bound = 1.0 / math.sqrt(fan_in)
        sd[name] = (torch.rand(shp, generator=g) * 2 - 1) * bound
    if dense:
        for lvl in ("coarse", "fine"):
            sd[f"nerf_mlps.{lvl}.fc_alpha.weight"] *= 400.0
            sd[f"nerf_mlps.{lvl}.fc_alpha.bias"].fill_(5.0)
        if "spatial_embeddings" in sd:
            sd["spatial_embeddings"] *= 30.0
    return sd


def state_checksum(sd: Dict[str, torch.Tensor]) -> float:
    return float(sum(v.double().abs().sum() for v in sd.values()))


def make_pose(seed: int = 0, z: float = 0.78, max_deg: float = 0.0) -> torch.Tensor:
    """[3,4] camera-to-world: identity (or a small seeded rotation) | (0,0,z)  (SURVEY.md B.5)."""
    g = torch.Generator().manual_seed(1000 + seed)
    R = torch.eye(3)
    if max_deg > 0:
        ang = (torch.rand(3, generator=g) * 2 - 1) * math.radians(max_deg)
        cx, cy, cz = torch.cos(ang)
        sx, sy, sz = torch.sin(ang)
        Rx = torch.tensor([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
        Ry = torch.tensor([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
        Rz = torch.tensor([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
        R = (Rz @ Ry @ Rx).float()
    t = torch.tensor([[0.0], [0.0], [z]])
    return torch.cat((R, t), dim=1).float()


def make_frame_inputs(spec: O.ModelSpec, H: int, W: int, seed: int = 0, pose_z: float = 0.78,
                      max_deg: float = 5.0, focal: float = 1200.0):
    """Synthetic per-frame inputs: pose, intrinsics, driving input, background, mask."""
    g = torch.Generator().manual_seed(2000 + seed)
    pose = make_pose(seed, pose_z, max_deg)
    intr = [focal * H / 512.0, focal * H / 512.0, 0.5, 0.5]
    if spec.model_type == "AudioFaceModel":
        driving = torch.randn(16, 29, generator=g)
    else:
        driving = torch.randn(76, generator=g) * 0.5
    bg = torch.cat((torch.rand(H, W, 3, generator=g), torch.ones(H, W, 1), torch.zeros(H, W, 11)), -1)
    cls = torch.randint(0, 12, (H, W), generator=g)
    mask = torch.nn.functional.one_hot(cls, 12).to(torch.int32)
    return dict(pose=pose, intrinsics=intr, driving=driving, background=bg, mask=mask)
