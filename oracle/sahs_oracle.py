"""CPU oracle for the SAHS per-ray render path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A from-scratch restatement (torch-CPU fp32 + numpy) of the reference algorithm
`run_one_iter_of_nerf` -> `predict_and_render_radiance` and everything below it.  Only `tests/`,
`__graft_entry__.smoke()` and bench.py's cpu_baseline / `--impl reference` legs may import this
module, and only as the checker / CPU baseline; the product package (sahs_b200) never does.

Pinning: the reference ships no tests or golden vectors, so this oracle is pinned
against the *live reference* imported in the build container by oracle/make_golden.py, which also
freezes seeded input/output vectors into tests/golden/*.npz ("parity unpinned by reference tests;
pinned by live reference run + frozen goldens").

All `ref:` citations are relative to /root/reference/nerf-pytorch/.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F

NUM_SEG = 12          # semantic classes, ref: nerf/modules.py:247 (fc_seg -> 12)
RAW_CH = 16           # rgb3 | seg12 | sigma1, ref: nerf/modules.py:295
DRIVING_DIM = 76      # ref: nerf/modules.py:200,341 (include_driving -> 76)
POSE_PE_DIM = 36      # 2*6*3, ref: nerf/modules.py:347, nerf/models.py:203-207
GRID_CH = 32          # ref: nerf/models.py:201


# --------------------------------------------------------------------------------------------
# architecture description derived from the YAML exactly like the reference constructors do
# --------------------------------------------------------------------------------------------
@dataclass
class ModelSpec:
    """Dimensions the reference derives in NeRFaceModel.__init__ (ref: nerf/models.py:189-299)."""
    model_type: str = "AudioFaceModel"
    xyz_L: int = 10
    xyz_inc: bool = True
    dir_L: int = 4
    dir_inc: bool = True
    use_viewdirs: bool = True
    use_ambient: bool = True
    amb_dim: int = 2
    amb_L: int = 4
    amb_inc: bool = True
    use_warp: bool = True
    warp_layers: int = 6
    warp_hidden: int = 128
    warp_skip: int = 4
    hyper_layers: int = 6
    hyper_hidden: int = 64
    hyper_skip: int = 4
    trunk_layers: int = 8
    trunk_hidden: int = 256
    trunk_skip: int = 3          # NeRFMLP default, never overridden (ref: nerf/modules.py:176, models.py:259-296)
    trunk_driving: bool = False  # include_driving (ref: nerf/modules.py:200)
    trunk_pose: bool = True      # use_pose (ref: nerf/modules.py:212-214)
    use_grid: bool = True
    num_coarse_train: int = 64
    num_fine_train: int = 64

    @property
    def xyz_dim(self):
        return (3 if self.xyz_inc else 0) + 6 * self.xyz_L

    @property
    def dir_dim(self):
        return (3 if self.dir_inc else 0) + 6 * self.dir_L

    @property
    def amb_pe_dim(self):
        if not self.use_ambient:
            return 0
        return (self.amb_dim if self.amb_inc else 0) + 2 * self.amb_dim * self.amb_L


def spec_from_cfg(cfg) -> ModelSpec:
    m = cfg.models
    return ModelSpec(
        model_type=m.mask.type,
        xyz_L=m.coarse.num_encoding_fn_xyz, xyz_inc=m.coarse.include_input_xyz,
        dir_L=m.coarse.num_encoding_fn_dir, dir_inc=m.coarse.include_input_dir,
        use_viewdirs=m.coarse.use_viewdirs,
        use_ambient=m.hyper.use_ambient, amb_dim=m.hyper.ambient_coord_dim,
        amb_L=m.hyper.num_encoding_fn_ambient, amb_inc=m.hyper.include_input_ambient,
        use_warp=m.warp.use_warp, warp_layers=m.warp.num_layers, warp_hidden=m.warp.hidden_size,
        warp_skip=m.warp.skip_connect_every,
        hyper_layers=m.hyper.num_layers, hyper_hidden=m.hyper.hidden_size,
        hyper_skip=m.hyper.skip_connect_every,
        trunk_layers=m.coarse.num_layers, trunk_hidden=m.coarse.hidden_size, trunk_skip=3,
        trunk_driving=bool(m.coarse.include_driving), trunk_pose=bool(m.coarse.use_pose),
        use_grid=bool(m.coarse.use_spatial_embeddings),
        num_coarse_train=cfg.nerf.train.num_coarse, num_fine_train=cfg.nerf.train.num_fine,
    )


# --------------------------------------------------------------------------------------------
# L1 math helpers
# --------------------------------------------------------------------------------------------
def get_ray_bundle(height: int, width: int, intrinsics, c2w: torch.Tensor):
    """Pinhole rays, ref: nerf/nerf_helpers.py:178-233 (+ meshgrid_xy :84-96).

    Pixel (row j, col i): dir = [(i - W*cx)/fx, -(j - H*cy)/fy, -1]; rd = R @ dir (un-normalised);
    ro = translation column.  Output shape (H, W, 3)."""
    fx, fy, cx, cy = [float(v) for v in intrinsics]
    dt = c2w.dtype
    cols = torch.arange(width, dtype=dt).view(1, width).expand(height, width)
    rows = torch.arange(height, dtype=dt).view(height, 1).expand(height, width)
    d = torch.stack(((cols - width * cx) / fx, -(rows - height * cy) / fy, -torch.ones_like(cols)), dim=-1)
    rd = (d[..., None, :] * c2w[:3, :3]).sum(dim=-1)
    ro = c2w[:3, -1].expand(rd.shape)
    return ro, rd


def get_ray_bundle_by_mask(height: int, width: int, intrinsics, c2w: torch.Tensor, mask: torch.Tensor):
    """ref: nerf/nerf_helpers.py:122-176: world-space rays where mask == 1, camera-frame direction and zero origin
    where mask == 0 (blended as mask * a + (1 - mask) * b)."""
    fx, fy, cx, cy = [float(v) for v in intrinsics]
    dt = c2w.dtype
    cols = torch.arange(width, dtype=dt).view(1, width).expand(height, width)
    rows = torch.arange(height, dtype=dt).view(height, 1).expand(height, width)
    d = torch.stack(((cols - width * cx) / fx, -(rows - height * cy) / fy, -torch.ones_like(cols)), dim=-1)
    m = mask.to(dt)[..., None].expand(height, width, 3)
    rd = (1 - m) * d + m * (d[..., None, :] * c2w[:3, :3]).sum(dim=-1)
    ro = (1 - m) * torch.zeros(1, 1, 3, dtype=dt) + m * c2w[:3, -1].expand(rd.shape)
    return ro, rd


def positional_encoding(x: torch.Tensor, num_freqs: int, include_input: bool = True) -> torch.Tensor:
    """[x, sin(2^0 x), cos(2^0 x), ..., sin(2^(L-1) x), cos(2^(L-1) x)], whole-vector blocks.
    ref: nerf/nerf_helpers.py:305-349 (log_sampling=True is the only mode the configs use)."""
    parts = [x] if include_input else []
    for k in range(num_freqs):
        f = float(2.0 ** k)
        parts.append(torch.sin(x * f))
        parts.append(torch.cos(x * f))
    return parts[0] if len(parts) == 1 else torch.cat(parts, dim=-1)


def pose_to_euler_trans(pose: torch.Tensor) -> torch.Tensor:
    """(e0,e1,e2,tx,ty,tz) from a [3|4, 4] pose, ref: nerf/models.py:482-504."""
    R = pose[:3, :3]
    e2 = torch.atan2(R[0, 0], -R[0, 1])
    e1 = torch.asin(-R[0, 2])
    e0 = torch.atan2(R[2, 2], R[1, 2])
    return torch.stack((e0, e1, e2, pose[0, 3], pose[1, 3], pose[2, 3]))


def pose_code(pose: torch.Tensor) -> torch.Tensor:
    """36-d pose code = PE_{L=3, no input}(euler|trans), ref: nerf/models.py:203-207, :519-520."""
    return positional_encoding(pose_to_euler_trans(pose)[None], 3, include_input=False)[0]


def audio_net(sd: Dict[str, torch.Tensor], audio: torch.Tensor, prefix="audNet_head.") -> torch.Tensor:
    """AudioNet: [16,29] window -> 76-d code, ref: nerf/modules.py:43-73."""
    x = audio[None, 0:16, :].permute(0, 2, 1)                      # [1,29,16]
    for i in (0, 2, 4, 6):
        x = F.conv1d(x, sd[f"{prefix}encoder_conv.{i}.weight"], sd[f"{prefix}encoder_conv.{i}.bias"],
                     stride=2, padding=1)
        x = F.leaky_relu(x, 0.02)
    x = x.squeeze(-1)
    x = F.leaky_relu(F.linear(x, sd[f"{prefix}encoder_fc1.0.weight"], sd[f"{prefix}encoder_fc1.0.bias"]), 0.02)
    x = F.linear(x, sd[f"{prefix}encoder_fc1.2.weight"], sd[f"{prefix}encoder_fc1.2.bias"])
    return x.reshape(-1)


def _skip_mlp(sd, prefix, n_layers, skip, x_in, act, taps=None, tap_name=None):
    """Shared body of WarpFieldMLP / HyperSheetMLP / NeRFMLP trunk: at layer `skip` the input is
    cat(x, initial) (ref: nerf/modules.py:254-262, :371-388, :444-460)."""
    x = x_in
    for i in range(n_layers):
        if i == skip:
            x = torch.cat((x, x_in), dim=-1)
        x = act(F.linear(x, sd[f"{prefix}.{i}.weight"], sd[f"{prefix}.{i}.bias"]))
        if taps is not None:
            taps[f"{tap_name}{i}"] = x
    return x


def grid_sample_trilinear(grid: torch.Tensor, pts: torch.Tensor) -> torch.Tensor:
    """Explicit form of sample_from_3dgrid (ref: nerf/models.py:346-365): grid [1,C,D,H,W],
    x indexes W (last), y indexes H, z indexes D; align_corners=True; zero padding; raw coords."""
    C, D, H, W = grid.shape[1:]
    g = grid[0].permute(1, 2, 3, 0)                               # [D,H,W,C]
    ix = (pts[:, 0] + 1.0) * (0.5 * (W - 1))
    iy = (pts[:, 1] + 1.0) * (0.5 * (H - 1))
    iz = (pts[:, 2] + 1.0) * (0.5 * (D - 1))
    x0, y0, z0 = torch.floor(ix), torch.floor(iy), torch.floor(iz)
    out = torch.zeros(pts.shape[0], C, dtype=pts.dtype, device=pts.device)
    for dz in (0, 1):
        for dy in (0, 1):
            for dx in (0, 1):
                xi, yi, zi = x0 + dx, y0 + dy, z0 + dz
                w = (1 - (ix - xi).abs()) * (1 - (iy - yi).abs()) * (1 - (iz - zi).abs())
                ok = (xi >= 0) & (xi <= W - 1) & (yi >= 0) & (yi <= H - 1) & (zi >= 0) & (zi <= D - 1)
                xl, yl, zl = xi.clamp(0, W - 1).long(), yi.clamp(0, H - 1).long(), zi.clamp(0, D - 1).long()
                out += (w * ok)[:, None] * g[zl, yl, xl]
    return out


def field_forward(sd: Dict[str, torch.Tensor], spec: ModelSpec, level: str, xyz: torch.Tensor,
                  viewdirs: torch.Tensor, driving: torch.Tensor, pose: torch.Tensor,
                  return_intermediates: bool = False):
    """raw[P,16] for sample points, ref: nerf/models.py:367-380 / :514-528 (+ :301-365).

    `driving` is the 76-d vector already produced by AudioNet (audio configs) or the raw expression
    vector (NeRFaceModel).  viewdirs are the un-normalised ray directions (ref: nerf/train_utils.py:15)."""
    P = xyz.shape[0]
    drv = driving.reshape(1, -1).expand(P, -1)
    pcode = pose_code(pose).reshape(1, -1).expand(P, -1)
    inter = {}
    e0 = positional_encoding(xyz, spec.xyz_L, spec.xyz_inc)
    mapped = xyz
    if spec.use_warp:
        h = _skip_mlp(sd, "warp_field_mlp.layers_xyz", spec.warp_layers, spec.warp_skip,
                      torch.cat((e0, drv, pcode), -1), F.relu, inter, "warp")
        dx = torch.tanh(F.linear(h, sd["warp_field_mlp.fc_final.weight"], sd["warp_field_mlp.fc_final.bias"]))
        mapped = xyz + dx
        inter["dx"] = dx
    amb = None
    if spec.use_ambient:
        h = _skip_mlp(sd, "hyper_sheep_mlp.layers_ambient", spec.hyper_layers, spec.hyper_skip,
                      torch.cat((e0, drv, pcode), -1), F.relu, inter, "hyper")
        amb = F.linear(h, sd["hyper_sheep_mlp.fc_ambient.weight"], sd["hyper_sheep_mlp.fc_ambient.bias"])
        inter["amb"] = amb
    emb = None
    if spec.use_grid:
        emb = grid_sample_trilinear(sd["spatial_embeddings"], mapped)
        inter["emb"] = emb
    # query_template, ref: nerf/models.py:331-344
    e1 = positional_encoding(mapped, spec.xyz_L, spec.xyz_inc)
    if amb is not None:
        e1 = torch.cat((e1, positional_encoding(amb, spec.amb_L, spec.amb_inc)), -1)
    ed = positional_encoding(viewdirs, spec.dir_L, spec.dir_inc) if spec.use_viewdirs else None
    p = f"nerf_mlps.{level}."
    initial = e1
    if spec.trunk_driving:
        initial = torch.cat((initial, drv), -1)
    # NeRFaceModel.forward does not hand `pose` to query_template (ref: nerf/models.py:378-379);
    # its configs have use_pose False so the branch is consistent.
    if spec.trunk_pose:
        initial = torch.cat((initial, pcode), -1)
    lrelu = lambda t: F.leaky_relu(t, 0.01)
    h = _skip_mlp(sd, p + "layers_xyz", spec.trunk_layers, spec.trunk_skip, initial, lrelu, inter, "trunk")
    feat = F.linear(h, sd[p + "fc_feat.weight"], sd[p + "fc_feat.bias"])
    sigma = F.linear(feat, sd[p + "fc_alpha.weight"], sd[p + "fc_alpha.bias"])
    hd = feat
    if spec.use_viewdirs:
        hd = torch.cat((feat, ed), -1)
        if emb is not None:
            hd = torch.cat((hd, emb), -1)
    for i in range(4):
        hd = lrelu(F.linear(hd, sd[p + f"layers_dir.{i}.weight"], sd[p + f"layers_dir.{i}.bias"]))
        inter[f"dir{i}"] = hd
    rgb = F.linear(hd, sd[p + "fc_rgb.weight"], sd[p + "fc_rgb.bias"])
    hs = feat
    for i in range(4):
        hs = lrelu(F.linear(hs, sd[p + f"layers_seg.{i}.weight"], sd[p + f"layers_seg.{i}.bias"]))
        inter[f"seg{i}"] = hs
    seg = F.linear(hs, sd[p + "fc_seg.weight"], sd[p + "fc_seg.bias"])
    raw = torch.cat((rgb, seg, sigma), -1)
    if return_intermediates:
        inter.update(mapped=mapped, feat=feat)
        return raw, inter
    return raw


def composite(raw: torch.Tensor, z: torch.Tensor, rd: torch.Tensor, noise: Optional[torch.Tensor] = None,
              white_background: bool = False, background_prior: Optional[torch.Tensor] = None):
    """Alpha compositing, ref: nerf/volume_rendering_utils.py:7-78 + cumprod_exclusive
    (nerf/nerf_helpers.py:99-120).  `raw` must already carry the background overwrite of the last
    sample (ref: nerf/train_utils.py:135-136).  `noise` (already scaled by noise_std) replaces the
    in-function randn draw.  Returns (rgb_map, disp, acc, weights, depth)."""
    dists = torch.cat((z[:, 1:] - z[:, :-1], torch.full_like(z[:, :1], 1e10)), -1)
    dists = dists * rd.norm(p=2, dim=-1, keepdim=True)
    if background_prior is not None:
        col = torch.sigmoid(raw[:, :-1, :3])
        if background_prior.shape[1] > 4:
            col = torch.cat((col, torch.softmax(raw[:, :-1, 3:-1], dim=-1)), -1)
        col = torch.cat((col, raw[:, -1:, :-1]), dim=1)          # last sample: raw bg values
    else:
        col = torch.sigmoid(raw[..., :-1])
    s = raw[..., -1] if noise is None else raw[..., -1] + noise
    sigma = torch.relu(s).clone()
    sigma[:, -1] += 1e-6
    alpha = 1.0 - torch.exp(-sigma * dists)
    t = torch.cumprod(1.0 - alpha + 1e-10, dim=-1)
    trans = torch.cat((torch.ones_like(t[:, :1]), t[:, :-1]), -1)
    w = alpha * trans
    rgb_map = (w[..., None] * col).sum(dim=-2)
    depth = (w * z).sum(dim=-1)
    acc = w.sum(dim=-1)
    disp = 1.0 / torch.max(1e-10 * torch.ones_like(depth), depth / acc)
    if white_background:
        rgb_map = rgb_map + (1.0 - acc[..., None])
    return rgb_map, disp, acc, w, depth


# ---- sample_pdf with ATen-CPU summation orders emulated in numpy (host independent) ---------
def _aten_inner_sum_f32(w: np.ndarray) -> np.ndarray:
    """Row sums of a contiguous fp32 [r,n] array in the order ATen's CPU `sum` uses
    (vectorized_inner_sum, 8 lanes x ilp 4).  Valid for n < 512."""
    r, n = w.shape
    assert w.dtype == np.float32 and n < 512
    nv = n // 8
    size_ilp = nv // 4
    vec = lambda k: w[:, 8 * k:8 * k + 8]
    zero = np.zeros((r, 8), np.float32)
    p = [zero.copy() for _ in range(4)]
    for i in range(size_ilp):
        for k in range(4):
            p[k] = (p[k] + vec(4 * i + k)).astype(np.float32)
    for i in range(size_ilp * 4, nv):
        p[0] = (p[0] + vec(i)).astype(np.float32)
    for k in range(1, 4):
        p[0] = (p[0] + p[k]).astype(np.float32)
    acc = np.zeros(r, np.float32)
    for k in range(nv * 8, n):
        acc = (acc + w[:, k]).astype(np.float32)
    for lane in range(8):
        acc = (acc + p[0][:, lane]).astype(np.float32)
    return acc


def _aten_cumsum_f32(x: np.ndarray) -> np.ndarray:
    """ATen CPU cumsum over the last dim: sequential, double accumulator, stored as fp32."""
    return np.cumsum(x.astype(np.float64), axis=-1).astype(np.float32)


def linspace_f32(steps: int) -> np.ndarray:
    """torch.linspace(0,1,steps) fp32 values (symmetric evaluation)."""
    return torch.linspace(0.0, 1.0, steps, dtype=torch.float32).numpy()


def sample_pdf(bins: torch.Tensor, weights: torch.Tensor, num_samples: int, det: bool = True,
               u: Optional[torch.Tensor] = None, return_inds: bool = False):
    """Inverse-CDF importance sampling, ref: nerf/nerf_helpers.py:454-497 (sample_pdf_2).
    CPU tensors: ATen's CPU summation orders emulated in numpy (host independent, bit-exact vs the reference on CPU).
    CUDA tensors (bench.py's torch-on-GPU baseline leg only): the reference's own torch ops on the device."""
    if bins.is_cuda:
        return _sample_pdf_torch(bins, weights, num_samples, det, u, return_inds)
    w = (weights.detach().numpy().astype(np.float32) + np.float32(1e-5)).astype(np.float32)
    total = _aten_inner_sum_f32(np.ascontiguousarray(w))
    pdf = (w / total[:, None]).astype(np.float32)
    cdf = np.concatenate((np.zeros((w.shape[0], 1), np.float32), _aten_cumsum_f32(pdf)), -1)
    if u is None:
        assert det, "stochastic mode needs the caller's uniform draws"
        uu = np.broadcast_to(linspace_f32(num_samples), (w.shape[0], num_samples)).copy()
    else:
        uu = u.numpy().astype(np.float32)
    nb = cdf.shape[-1]
    # searchsorted(right=True): number of cdf entries <= u
    inds = (cdf[:, None, :] <= uu[:, :, None]).sum(-1).astype(np.int64)
    below = np.maximum(inds - 1, 0)
    above = np.minimum(inds, nb - 1)
    b = bins.detach().numpy().astype(np.float32)
    cdf_b, cdf_a = np.take_along_axis(cdf, below, 1), np.take_along_axis(cdf, above, 1)
    bin_b, bin_a = np.take_along_axis(b, below, 1), np.take_along_axis(b, above, 1)
    denom = (cdf_a - cdf_b).astype(np.float32)
    denom = np.where(denom < np.float32(1e-5), np.float32(1.0), denom)
    t = ((uu - cdf_b).astype(np.float32) / denom).astype(np.float32)
    samples = (bin_b + (t * (bin_a - bin_b).astype(np.float32)).astype(np.float32)).astype(np.float32)
    out = torch.from_numpy(samples)
    if return_inds:
        return out, torch.from_numpy(inds)
    return out


def _sample_pdf_torch(bins, weights, num_samples, det, u, return_inds):
    """sample_pdf_2 with the reference's torch ops (ref: nerf/nerf_helpers.py:454-497), any device."""
    w = weights + 1e-5
    pdf = w / torch.sum(w, dim=-1, keepdim=True)
    cdf = torch.cat((torch.zeros_like(pdf[..., :1]), torch.cumsum(pdf, dim=-1)), dim=-1)
    if u is None:
        assert det, "stochastic mode needs the caller's uniform draws"
        u = torch.linspace(0.0, 1.0, steps=num_samples, dtype=w.dtype, device=w.device).expand(w.shape[0], num_samples)
    u = u.contiguous()
    inds = torch.searchsorted(cdf.contiguous(), u, right=True)
    below = torch.clamp(inds - 1, min=0)
    above = torch.clamp(inds, max=cdf.shape[-1] - 1)
    cdf_b, cdf_a = torch.gather(cdf, 1, below), torch.gather(cdf, 1, above)
    bin_b, bin_a = torch.gather(bins, 1, below), torch.gather(bins, 1, above)
    denom = cdf_a - cdf_b
    denom = torch.where(denom < 1e-5, torch.ones_like(denom), denom)
    samples = bin_b + (u - cdf_b) / denom * (bin_a - bin_b)
    return (samples, inds) if return_inds else samples


# --------------------------------------------------------------------------------------------
# L3 pipeline
# --------------------------------------------------------------------------------------------
@dataclass
class RenderOpts:
    """The config keys predict_and_render_radiance reads (ref: nerf/train_utils.py:93-162)."""
    num_coarse: int = 64
    num_fine: int = 64
    perturb: bool = False
    lindisp: bool = False
    noise_std: float = 0.0
    white_background: bool = False
    near: float = 0.0
    far: float = 1.0


def opts_from_cfg(cfg, mode: str) -> RenderOpts:
    n = getattr(cfg.nerf, mode)
    return RenderOpts(num_coarse=n.num_coarse, num_fine=n.num_fine, perturb=bool(n.perturb),
                      lindisp=bool(n.lindisp), noise_std=float(n.radiance_field_noise_std),
                      white_background=bool(n.white_background), near=float(cfg.dataset.near),
                      far=float(cfg.dataset.far))


def coarse_z(opts: RenderOpts, num_rays: int, t_rand: Optional[torch.Tensor] = None, device=None) -> torch.Tensor:
    """ref: nerf/train_utils.py:93-113."""
    near = torch.full((num_rays, 1), opts.near, dtype=torch.float32, device=device)
    far = torch.full((num_rays, 1), opts.far, dtype=torch.float32, device=device)
    t = torch.linspace(0.0, 1.0, opts.num_coarse, dtype=torch.float32).to(near.device)
    if not opts.lindisp:
        z = near * (1.0 - t) + far * t
    else:
        z = 1.0 / (1.0 / near * (1.0 - t) + 1.0 / far * t)
    if opts.perturb:
        assert t_rand is not None, "stochastic mode needs the caller's uniform draws"
        mids = 0.5 * (z[:, 1:] + z[:, :-1])
        upper = torch.cat((mids, z[:, -1:]), -1)
        lower = torch.cat((z[:, :1], mids), -1)
        z = lower + (upper - lower) * t_rand
    return z


def render_rays(sd, spec: ModelSpec, opts: RenderOpts, ro, rd, driving, pose, background_prior=None,
                t_rand=None, u=None, noise_c=None, noise_f=None, chunk_points: int = 1 << 17,
                return_aux: bool = False):
    """predict_and_render_radiance restated (ref: nerf/train_utils.py:72-206).  Returns the 8-tuple
    (rgb_c, disp_c, acc_c, rgb_f, disp_f, acc_f, w_last_f, depth_f).  Random draws are inputs."""
    R = ro.shape[0]

    def run_field(level, z):
        S = z.shape[1]
        pts = (ro[:, None, :] + rd[:, None, :] * z[:, :, None]).reshape(-1, 3)
        dirs = rd[:, None, :].expand(R, S, 3).reshape(-1, 3)
        outs = [field_forward(sd, spec, level, pts[i:i + chunk_points], dirs[i:i + chunk_points], driving, pose)
                for i in range(0, pts.shape[0], chunk_points)]
        raw = torch.cat(outs, 0).reshape(R, S, RAW_CH)
        if background_prior is not None:
            raw[:, -1, :-1] = background_prior
        return raw

    z_c = coarse_z(opts, R, t_rand, device=ro.device)
    raw_c = run_field("coarse", z_c)
    rgb_c, disp_c, acc_c, w_c, depth_c = composite(raw_c, z_c, rd, noise_c, opts.white_background, background_prior)
    z_mid = 0.5 * (z_c[:, 1:] + z_c[:, :-1])
    z_s = sample_pdf(z_mid, w_c[:, 1:-1], opts.num_fine, det=not opts.perturb, u=u)
    z_f, _ = torch.sort(torch.cat((z_c, z_s), -1), dim=-1)
    raw_f = run_field("fine", z_f)
    rgb_f, disp_f, acc_f, w_f, depth_f = composite(raw_f, z_f, rd, noise_f, opts.white_background, background_prior)
    out = (rgb_c, disp_c, acc_c, rgb_f, disp_f, acc_f, w_f[:, -1], depth_f)
    if return_aux:
        return out, dict(z_c=z_c, raw_c=raw_c, w_c=w_c, depth_c=depth_c, z_s=z_s, z_f=z_f, raw_f=raw_f, w_f=w_f)
    return out


def driving_vector(sd, spec: ModelSpec, driving_in: torch.Tensor) -> torch.Tensor:
    """AudioFaceModel runs AudioNet on the [16,29] window (ref: nerf/models.py:517); NeRFaceModel
    uses the 76-d expression vector as is (ref: nerf/models.py:370)."""
    if spec.model_type == "AudioFaceModel":
        return audio_net(sd, driving_in)
    return driving_in.reshape(-1)


def run_one_iter(sd, spec: ModelSpec, opts: RenderOpts, ray_origins, ray_directions, driving_in, pose,
                 background_prior=None, chunk_rays: int = 1 << 17, **draws):
    """run_one_iter_of_nerf restated for the live no_ndc=True branch (ref: nerf/train_utils.py:209-321);
    flat [R,...] outputs (train-mode shape)."""
    ro = ray_origins.reshape(-1, 3)
    rd = ray_directions.reshape(-1, 3)
    drv = driving_vector(sd, spec, driving_in)
    outs = []
    for i in range(0, ro.shape[0], chunk_rays):
        bg = background_prior[i:i + chunk_rays] if background_prior is not None else None
        outs.append(render_rays(sd, spec, opts, ro[i:i + chunk_rays], rd[i:i + chunk_rays], drv, pose, bg,
                                **{k: (v[i:i + chunk_rays] if v is not None else None) for k, v in draws.items()}))
    return tuple(torch.cat(parts, 0) for parts in zip(*outs))


# ---- frame post-processing (what the eval script writes) ------------------------------------------------------
_SEG_PALETTE = [[0, 0, 0], [204, 0, 0], [76, 153, 0], [204, 204, 0], [51, 51, 255], [0, 255, 255], [102, 51, 0],
                [102, 204, 0], [255, 255, 0], [0, 0, 204], [255, 153, 51], [0, 204, 0]]


def frame_postprocess(rgb_map: torch.Tensor):
    """uint8 rgb (clamp, *255, truncate: torchvision ToPILImage on a float tensor, ref: eval_stage_rays.py:221-227),
    argmax label and reversed-channel palette colour (ref: nerf/utils.py:112-140)."""
    rgb = (rgb_map[..., :3].clamp(0.0, 1.0) * 255).to(torch.uint8)
    label = torch.argmax(rgb_map[..., 3:], dim=-1)
    pal = torch.tensor([c[::-1] for c in _SEG_PALETTE], dtype=torch.uint8)
    return rgb, label.to(torch.uint8), pal[label]


def torch_normal_map(depthmap: torch.Tensor, focal, weights: Optional[torch.Tensor] = None, clean: bool = True,
                     central_difference: bool = False) -> torch.Tensor:
    """Depth map -> normal map, ref: eval_stage_rays.py:116-151 (square maps, as the reference's broadcasting requires)."""
    n = depthmap.shape[0]
    cx, cy, fx, fy = focal[2] * n, focal[3] * n, focal[0], focal[1]
    cols = torch.arange(n).view(1, n).expand(n, n)
    rows = torch.arange(n).view(n, 1).expand(n, n)
    points = torch.stack((((cols - cx) * depthmap) / fx, -((rows - cy) * depthmap) / fy, depthmap), dim=-1)
    k = 2 if central_difference else 1
    dx = points[k:, :, :] - points[:-k, :, :]
    dy = points[:, k:, :] - points[:, :-k, :]
    normals = torch.cross(dy[:-k, :, :], dx[:, :-k, :], dim=2)
    normals = normals / torch.sqrt(torch.sum(normals * normals, 2, keepdim=True))
    normals = normals * 0.5 + 0.5
    if clean and weights is not None:
        mask = weights[:-k, :-k, None].expand(-1, -1, 3)
        normals = torch.where(mask > 0.22, torch.ones_like(normals), normals)
        normals = (1 - mask) * normals + mask * torch.ones_like(normals)
    return normals * 255


def weighted_sample_probs(mask: torch.Tensor, class_prob: torch.Tensor) -> np.ndarray:
    """probs of the semantic-weighted ray sampler (ref: train_stage_rays_auto.py:390-394):
    sum_c sample_prob[c] * mask[..., c], normalised."""
    p = (class_prob.reshape(1, -1) * mask.reshape(-1, mask.shape[-1]).float()).sum(-1).double().numpy()
    return p / p.sum()


def weighted_sample(mask: torch.Tensor, class_prob: torch.Tensor, num_select: int, rng: np.random.Generator) -> np.ndarray:
    """The reference's draw (ref: train_stage_rays_auto.py:416-418): np.random.choice(N, n, replace=False, p=probs).
    (A Generator instead of the legacy global RandomState: same sequential without-replacement distribution.)"""
    p = weighted_sample_probs(mask, class_prob)
    return rng.choice(p.shape[0], size=num_select, replace=False, p=p)


def masked_mse(mask: torch.Tensor, inp: torch.Tensor, target: torch.Tensor):
    """MaskMSELoss.forward (ref: nerf/nerf_helpers.py:40-62) with weights=None -> (mean, per-class)."""
    mask = mask.reshape(-1, mask.shape[-1])
    count = torch.count_nonzero(mask, dim=0)
    count = torch.where(count == 0, torch.ones_like(count), count)
    diff = torch.sum(torch.square(inp.reshape(-1, 3) - target.reshape(-1, 3)), dim=-1).unsqueeze(-1)
    return torch.mean(diff), torch.sum(diff * mask, dim=0) / count


def masked_cross_entropy(mask: torch.Tensor, inp: torch.Tensor, target: torch.Tensor):
    """MaskCrossEntropyLoss.forward (ref: nerf/nerf_helpers.py:14-37) with weights=None -> (mean, per-class)."""
    mask = mask.reshape(-1, mask.shape[-1])
    count = torch.count_nonzero(mask, dim=0)
    count = torch.where(count == 0, torch.ones_like(count), count)
    ce = -torch.sum(target.reshape(-1, target.shape[-1]) * torch.log(inp.reshape(-1, inp.shape[-1]) + 1e-10),
                    dim=-1).unsqueeze(-1)
    return torch.mean(ce), torch.sum(ce * mask, dim=0) / count


def stage1_loss(rgb_coarse: torch.Tensor, rgb_fine: torch.Tensor, target_rgb: torch.Tensor, mask: torch.Tensor):
    """Loss assembly of the training script (ref: train_stage_rays_auto.py:455-468): per level
    l2 + 0.02 * CE + 0.005 * sum(masked_l2[7:9] + masked_CE[7:9]); sample_prob = normalised sum of the four per-class
    vectors.  Differentiable torch ops (the gradient tests run autograd through it).  Returns (loss, sample_prob)."""
    mask = mask.to(rgb_coarse.dtype)
    total, parts = 0.0, []
    for rgb in (rgb_coarse, rgb_fine):
        l2, m_l2 = masked_mse(mask, rgb[..., :3], target_rgb[..., :3])
        ce, m_ce = masked_cross_entropy(mask, rgb[..., 3:], mask)
        total = total + (l2 + 0.02 * ce + 0.005 * torch.sum(m_l2[7:9] + m_ce[7:9]))
        parts += [m_l2, m_ce]
    s = parts[0] + parts[1] + parts[2] + parts[3]
    return total, (s / (parts[0].sum() + parts[1].sum() + parts[2].sum() + parts[3].sum())).detach()
