"""Field models with the reference's constructor, call signature and state_dict layout.

ref: nerf/models.py:189-380 (NeRFaceModel), :482-528 (pose_to_euler_trans, AudioFaceModel),
     nerf/modules.py:43-73 (AudioNet), :168-295 (NeRFMLP), :323-390 (WarpFieldMLP), :401-462 (HyperSheetMLP).

The torch modules below are *parameter containers only*: same names and shapes as the reference so that
`checkpoint["model_state_dict"]` loads unchanged.  The arithmetic of WarpFieldMLP / HyperSheetMLP / NeRFMLP /
the embedding-grid gather runs in the fused sm_100a kernel (csrc/field_fwd.cu); the per-frame AudioNet
(0.1 MFLOP, once per frame) stays a handful of PyTorch library calls as SURVEY.md section 8a (a12) prescribes.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, List, Optional

import torch

from . import lib as L
from . import ops

DRIVING_DIM = 76
POSE_CODE_DIM = 36
GRID_CH = 32


# Packed weight images are cached per model and must be rebuilt after every weight update.  In-place updates through
# ordinary torch ops bump `Tensor._version`; fused optimizers (Adam(fused=True), ...) do not, so every optimizer step in
# the process also advances this epoch.  Code that rewrites weights behind torch's back calls model.invalidate_packed().
_OPT_EPOCH = [0]


def _on_optimizer_step(*_args, **_kwargs):
    _OPT_EPOCH[0] += 1


try:
    from torch.optim.optimizer import register_optimizer_step_post_hook as _reg_hook
    _reg_hook(_on_optimizer_step)
except Exception:  # noqa: BLE001  (very old torch: fall back to the version counters only)
    pass


@dataclass(frozen=True)
class ModelSpec:
    """Dimensions the reference derives in NeRFaceModel.__init__ (ref: nerf/models.py:189-299)."""
    model_type: str
    xyz_L: int
    xyz_inc: bool
    dir_L: int
    dir_inc: bool
    use_viewdirs: bool
    use_ambient: bool
    amb_dim: int
    amb_L: int
    amb_inc: bool
    use_warp: bool
    warp_layers: int
    warp_hidden: int
    warp_skip: int
    hyper_layers: int
    hyper_hidden: int
    hyper_skip: int
    trunk_layers: int
    trunk_hidden: int
    trunk_skip: int
    trunk_driving: bool
    trunk_pose: bool
    use_grid: bool

    @staticmethod
    def from_cfg(cfg) -> "ModelSpec":
        m = cfg.models
        # Quirks preserved: the fine level is built from models.coarse.num_layers /
        # hidden_size / use_pose, and NeRFMLP's skip_connect_every stays at its default 3 because the
        # constructor call never forwards the YAML value (ref: nerf/models.py:259-296, nerf/modules.py:176).
        return ModelSpec(
            model_type=m.mask.type,
            xyz_L=int(m.coarse.num_encoding_fn_xyz), xyz_inc=bool(m.coarse.include_input_xyz),
            dir_L=int(m.coarse.num_encoding_fn_dir), dir_inc=bool(m.coarse.include_input_dir),
            use_viewdirs=bool(m.coarse.use_viewdirs),
            use_ambient=bool(m.hyper.use_ambient), amb_dim=int(m.hyper.ambient_coord_dim),
            amb_L=int(m.hyper.num_encoding_fn_ambient), amb_inc=bool(m.hyper.include_input_ambient),
            use_warp=bool(m.warp.use_warp), warp_layers=int(m.warp.num_layers),
            warp_hidden=int(m.warp.hidden_size), warp_skip=int(m.warp.skip_connect_every),
            hyper_layers=int(m.hyper.num_layers), hyper_hidden=int(m.hyper.hidden_size),
            hyper_skip=int(m.hyper.skip_connect_every),
            trunk_layers=int(m.coarse.num_layers), trunk_hidden=int(m.coarse.hidden_size), trunk_skip=3,
            trunk_driving=bool(m.coarse.include_driving), trunk_pose=bool(m.coarse.use_pose),
            use_grid=bool(m.coarse.use_spatial_embeddings))

    @property
    def xyz_dim(self) -> int:
        return (3 if self.xyz_inc else 0) + 6 * self.xyz_L

    @property
    def dir_dim(self) -> int:
        return (3 if self.dir_inc else 0) + 6 * self.dir_L

    @property
    def amb_pe_dim(self) -> int:
        if not self.use_ambient:
            return 0
        return (self.amb_dim if self.amb_inc else 0) + 2 * self.amb_dim * self.amb_L

    def to_c(self) -> L.ModelSpecC:
        c = L.ModelSpecC()
        for name, _ in L.ModelSpecC._fields_:
            setattr(c, name, int(getattr(self, name)))
        if not self.use_ambient:
            c.amb_dim, c.amb_L, c.amb_inc = 0, 0, 0
        return c


class _Holder(torch.nn.Module):
    """nn.Module without a forward: carries named Linear layers for state_dict compatibility."""

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("parameter container; the arithmetic runs in the fused CUDA field kernel")


def _skip_stack(n_layers, hidden, skip, in_dim) -> torch.nn.ModuleList:
    layers = torch.nn.ModuleList()
    for i in range(n_layers):
        k = in_dim if i == 0 else (hidden + in_dim if i == skip else hidden)
        layers.append(torch.nn.Linear(k, hidden))
    return layers


class AudioNet(torch.nn.Module):
    """ref: nerf/modules.py:43-73 (per-frame audio window [16,29] -> 76-d code)."""

    def __init__(self, dim_aud=76, win_size=16):
        super().__init__()
        self.win_size, self.dim_aud = win_size, dim_aud
        conv = lambda ci, co: torch.nn.Conv1d(ci, co, kernel_size=3, stride=2, padding=1, bias=True)
        act = lambda: torch.nn.LeakyReLU(0.02, True)
        self.encoder_conv = torch.nn.Sequential(conv(29, 32), act(), conv(32, 32), act(), conv(32, 64), act(),
                                                conv(64, 64), act())
        self.encoder_fc1 = torch.nn.Sequential(torch.nn.Linear(64, 64), act(), torch.nn.Linear(64, dim_aud))

    def forward(self, x):
        half = self.win_size // 2
        x = x[:, 8 - half:8 + half, :].permute(0, 2, 1)
        x = self.encoder_conv(x).squeeze(-1)
        return self.encoder_fc1(x).squeeze()


def rot_to_euler(R: torch.Tensor) -> torch.Tensor:
    """ref: nerf/models.py:482-498."""
    e2 = torch.atan2(R[:, 0, 0], -R[:, 0, 1])
    e1 = torch.asin(-R[:, 0, 2])
    e0 = torch.atan2(R[:, 2, 2], R[:, 1, 2])
    return torch.stack((e0, e1, e2), dim=1)


def pose_to_euler_trans(poses: torch.Tensor, device=None) -> torch.Tensor:
    """ref: nerf/models.py:501-504."""
    return torch.cat((rot_to_euler(poses), poses[:, :3, 3]), dim=1)


class NeRFaceModel(torch.nn.Module):
    """ref: nerf/models.py:189-380.  model(level, x[P,>=6], driving, pose, pose_c, latent_code=None) -> raw[P,16]."""

    def __init__(self, cfg):
        super().__init__()
        self.spec = ModelSpec.from_cfg(cfg)
        s = self.spec
        self.num_coarse = cfg.nerf.train.num_coarse
        self.num_fine = cfg.nerf.train.num_fine
        self.spatial_embeddings = None
        if s.use_grid:
            self.spatial_embeddings = torch.nn.Parameter(torch.randn(1, GRID_CH, 32, 32, 32) * 0.01)
        in0 = s.xyz_dim + DRIVING_DIM + POSE_CODE_DIM          # dim_pose = 0 + 36 even with include_pose False
        if s.use_warp:
            self.warp_field_mlp = _Holder()
            self.warp_field_mlp.layers_xyz = _skip_stack(s.warp_layers, s.warp_hidden, s.warp_skip, in0)
            self.warp_field_mlp.fc_final = torch.nn.Linear(s.warp_hidden, 3)
        if s.use_ambient:
            self.hyper_sheep_mlp = _Holder()
            self.hyper_sheep_mlp.layers_ambient = _skip_stack(s.hyper_layers, s.hyper_hidden, s.hyper_skip, in0)
            self.hyper_sheep_mlp.fc_ambient = torch.nn.Linear(s.hyper_hidden, s.amb_dim)
        tin = s.xyz_dim + s.amb_pe_dim + (DRIVING_DIM if s.trunk_driving else 0) + (POSE_CODE_DIM if s.trunk_pose else 0)
        H = s.trunk_hidden
        mlps = {}
        levels = ["coarse"] + (["fine"] if hasattr(cfg.models, "fine") else [])
        for lvl in levels:
            m = _Holder()
            m.layers_xyz = _skip_stack(s.trunk_layers, H, s.trunk_skip, tin)
            m.fc_feat = torch.nn.Linear(H, H)
            m.fc_alpha = torch.nn.Linear(H, 1)
            d0 = H + (s.dir_dim if s.use_viewdirs else 0) + (GRID_CH if s.use_grid and s.use_viewdirs else 0)
            m.layers_dir = torch.nn.ModuleList([torch.nn.Linear(d0 if i == 0 else H // 2, H // 2) for i in range(4)])
            m.fc_rgb = torch.nn.Linear(H // 2, 3)
            m.layers_seg = torch.nn.ModuleList([torch.nn.Linear(H if i == 0 else H // 2, H // 2) for i in range(4)])
            m.fc_seg = torch.nn.Linear(H // 2, 12)
            mlps[lvl] = m
        self.nerf_mlps = torch.nn.ModuleDict(mlps)
        self._packed: Dict[int, dict] = {}
        self._invalidations = 0

    # ---- packing: fp32 master parameters -> fp16 stage images (re-done whenever a parameter changed) ----
    def _level_params(self, level: str) -> List[Optional[torch.Tensor]]:
        s = self.spec
        out: List[Optional[torch.Tensor]] = [self.spatial_embeddings if s.use_grid else None]
        if s.use_warp:
            for lin in self.warp_field_mlp.layers_xyz:
                out += [lin.weight, lin.bias]
            out += [self.warp_field_mlp.fc_final.weight, self.warp_field_mlp.fc_final.bias]
        if s.use_ambient:
            for lin in self.hyper_sheep_mlp.layers_ambient:
                out += [lin.weight, lin.bias]
            out += [self.hyper_sheep_mlp.fc_ambient.weight, self.hyper_sheep_mlp.fc_ambient.bias]
        m = self.nerf_mlps[level]
        for lin in m.layers_xyz:
            out += [lin.weight, lin.bias]
        out += [m.fc_feat.weight, m.fc_feat.bias, m.fc_alpha.weight, m.fc_alpha.bias]
        for lin in m.layers_dir:
            out += [lin.weight, lin.bias]
        out += [m.fc_rgb.weight, m.fc_rgb.bias]
        for lin in m.layers_seg:
            out += [lin.weight, lin.bias]
        out += [m.fc_seg.weight, m.fc_seg.bias]
        return out

    def invalidate_packed(self) -> None:
        """Force a repack on the next call (for weight updates torch cannot see, e.g. custom kernels)."""
        self._invalidations += 1

    def packed_level(self, level: str) -> dict:
        """Packed weight image + parameter pointer table of one level, rebuilt when parameters changed."""
        lib = L.load()
        lvl = 0 if level == "coarse" else 1
        params = [p.detach() if p is not None else None for p in self._level_params(level)]
        # staleness key: storage + autograd version of every parameter, plus the process-wide optimizer-step epoch
        # (fused / foreach optimizers update parameters without bumping `_version`)
        key = (_OPT_EPOCH[0], self._invalidations) + tuple((p.data_ptr(), p._version) for p in params if p is not None)
        st = self._packed.get(lvl)
        if st is not None and st["key"] == key:
            return st
        for p in params:
            if p is not None and (not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous()):
                raise RuntimeError("model parameters must be contiguous fp32 CUDA tensors (model.to('cuda'))")
        cspec = self.spec.to_c()
        if lib.sahs_param_count(C.byref(cspec)) != len(params):
            raise RuntimeError("internal error: parameter list does not match the C ABI's canonical order")
        pb, fb, gb = C.c_size_t(), C.c_size_t(), C.c_size_t()
        L.check(lib.sahs_field_sizes(C.byref(cspec), C.byref(pb), C.byref(fb), C.byref(gb)), "field_sizes")
        dev = params[-1].device
        packed = torch.empty(pb.value, dtype=torch.uint8, device=dev)
        grid = torch.empty(gb.value // 4, dtype=torch.float32, device=dev) if gb.value else None
        arr = L.param_ptr_array(params)
        L.check(lib.sahs_pack_params(C.byref(cspec), lvl, arr, L.ptr(packed), L.ptr(grid), L.stream_ptr(dev)),
                "pack_params")
        from .custom_ops import spec_ints
        st = dict(key=key, cspec=cspec, spec_ints=spec_ints(cspec), params=params, arr=arr, packed=packed, grid=grid,
                  fc_floats=fb.value // 4)
        self._packed[lvl] = st
        return st

    def frame_constants(self, level: str, driving_vec: torch.Tensor, pose_code: torch.Tensor) -> torch.Tensor:
        """Biases with the frame-constant driving/pose input columns folded in (+ small fp32 head weights)."""
        lib = L.load()
        st = self.packed_level(level)
        dev = st["packed"].device
        drv, pc = L.f32c(driving_vec.detach()).reshape(-1), L.f32c(pose_code.detach()).reshape(-1)
        if drv.numel() != DRIVING_DIM or pc.numel() != POSE_CODE_DIM:
            raise RuntimeError("driving vector must have 76 entries and the pose code 36")
        fc = torch.empty(st["fc_floats"], dtype=torch.float32, device=dev)
        L.check(lib.sahs_fold_frame(C.byref(st["cspec"]), 0 if level == "coarse" else 1, st["arr"], L.ptr(drv),
                                    L.ptr(pc), L.ptr(fc), L.stream_ptr(dev)), "fold_frame")
        return fc

    # ---- per-frame conditioning (tiny; PyTorch) ----
    def driving_vector(self, driving: torch.Tensor) -> torch.Tensor:
        return driving.reshape(-1)                       # raw 76-d expression (ref: nerf/models.py:370)

    def pose_code(self, pose: torch.Tensor) -> torch.Tensor:
        """36-d PE_{L=3,no input}(euler|trans), ref: nerf/models.py:203-207, :371-372."""
        e = pose_to_euler_trans(pose[None, :3, :4].float())
        return ops.positional_encoding(e, 3, include_input=False).reshape(-1)

    def field(self, level: str, ro, rd, z, driving_vec, pose_code, frame_const=None, debug=None, debug_pass=-1):
        """raw[R,S,16] for points ro + rd*z (the fused kernel).  ref: nerf/train_utils.py:9-50."""
        lib = L.load()
        if debug is None and frame_const is None and torch.is_grad_enabled() and \
                any(p.requires_grad for p in self.parameters()):
            from .train import field_train          # training: tapes + hand-written backward
            return field_train(self, level, ro, rd, z, driving_vec, pose_code)
        st = self.packed_level(level)
        fc = frame_const if frame_const is not None else self.frame_constants(level, driving_vec, pose_code)
        lvl = 0 if level == "coarse" else 1
        if debug is not None:          # diagnostics (per-pass outputs, timelines): straight through the C ABI
            return ops.field_fwd(st["cspec"], lvl, st["packed"], fc, st["grid"], ro, rd, z, debug, debug_pass)
        if not z.is_cuda:
            raise RuntimeError("sahs_b200 ops need CUDA tensors (there is no CPU path)")
        return torch.ops.sahs_b200.field_fwd(st["spec_ints"], lvl, st["packed"], fc, st["grid"], ro, rd, z)

    def forward(self, level, x, driving=None, pose=None, pose_c=None, latent_code=None, **kwargs):
        """Reference call signature (ref: nerf/models.py:367-380): x[...,:3] points, x[...,3:6] view directions;
        the mask columns x[...,6:] and pose_c ride along unused exactly as in the reference's live code."""
        pts = x[..., :3].reshape(-1, 3)
        dirs = x[..., 3:6].reshape(-1, 3)
        z0 = torch.zeros(pts.shape[0], 1, dtype=torch.float32, device=pts.device)   # point = ro + rd*0
        raw = self.field(level, pts, dirs, z0, self.driving_vector(driving), self.pose_code(pose))
        return raw.reshape(*x.shape[:-1], 16)


class AudioFaceModel(NeRFaceModel):
    """ref: nerf/models.py:507-528."""

    def __init__(self, cfg):
        super().__init__(cfg)
        self.audNet_head = AudioNet(76, 16)

    def driving_vector(self, audio: torch.Tensor) -> torch.Tensor:
        return self.audNet_head(audio.unsqueeze(0)).reshape(-1)       # ref: nerf/models.py:517
