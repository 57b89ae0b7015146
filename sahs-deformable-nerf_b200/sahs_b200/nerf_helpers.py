"""Reference-named helpers of nerf/nerf_helpers.py backed by the CUDA library.

ref: nerf/nerf_helpers.py:76-81 (get_minibatches), :84-96 (meshgrid_xy), :99-120 (cumprod_exclusive),
     :178-233 (get_ray_bundle), :305-359 (positional_encoding, get_embedding_function), :454-497 (sample_pdf_2).
"""
from __future__ import annotations

import torch

from . import ops


def get_minibatches(inputs: torch.Tensor, chunksize: int = 1024 * 8):
    """List of row chunks (ref: nerf/nerf_helpers.py:76-81)."""
    return [inputs[i:i + chunksize] for i in range(0, inputs.shape[0], chunksize)]


def meshgrid_xy(tensor1: torch.Tensor, tensor2: torch.Tensor):
    """numpy 'xy' meshgrid (ref: nerf/nerf_helpers.py:84-96)."""
    ii, jj = torch.meshgrid(tensor1, tensor2, indexing="ij")
    return ii.transpose(-1, -2), jj.transpose(-1, -2)


def cumprod_exclusive(tensor: torch.Tensor) -> torch.Tensor:
    """Exclusive cumulative product along the last dim (ref: nerf/nerf_helpers.py:99-120).  The render path
    does not call this (the scan lives inside the compositing kernel); kept for API compatibility."""
    c = torch.cumprod(tensor, -1)
    return torch.cat((torch.ones_like(c[..., :1]), c[..., :-1]), -1)


def get_ray_bundle(height: int, width: int, intrinsics, tform_cam2world: torch.Tensor, center=[0.5, 0.5]):
    """ref: nerf/nerf_helpers.py:178-233.  `intrinsics` = [fx, fy, cx, cy] (cx, cy relative) or a scalar focal."""
    try:
        n = len(intrinsics)
    except TypeError:
        n = 0
    if n < 4:
        f = float(intrinsics if n == 0 else intrinsics[0])
        intrinsics = [f, f, 0.5, 0.5]
    fx, fy, cx, cy = [float(v) for v in intrinsics]
    return torch.ops.sahs_b200.get_ray_bundle(int(height), int(width), fx, fy, cx, cy, tform_cam2world)


def get_ray_bundle_by_mask(height: int, width: int, intrinsics, tform_cam2world: torch.Tensor, mask, center=[0.5, 0.5]):
    """ref: nerf/nerf_helpers.py:122-176.  Pixels with mask == 1 get the world-space ray of `get_ray_bundle`, pixels with
    mask == 0 keep the camera-frame direction and a zero origin (the reference blends with mask / 1 - mask)."""
    ro_w, rd_w = get_ray_bundle(height, width, intrinsics, tform_cam2world)
    eye = torch.eye(3, 4, dtype=torch.float32, device=rd_w.device)
    _, rd_cam = get_ray_bundle(height, width, intrinsics, eye)
    m = mask.to(rd_w)[..., None].expand_as(rd_w)
    return m * ro_w, (1 - m) * rd_cam + m * rd_w


def dump_rays(origins, points, radiance_field, path="rays_small.ply"):
    """Debug point-cloud dump (ref: nerf/nerf_helpers.py:499-540): samples whose sigmoid(relu(raw[..., 3])) exceeds
    0.9999996, every 100th of the first tenth of them, as an ASCII PLY with the sample's first three channels as colour."""
    dens = torch.sigmoid(torch.relu(radiance_field[:, :, 3]))
    ray_idx, depth_idx = torch.where(dens > 0.9999996)
    total = int(ray_idx.shape[0] // 10)
    print("point cloud with %d points" % total)
    keep = torch.arange(0, total, 100, device=ray_idx.device)
    pts = points[ray_idx[keep], depth_idx[keep]].detach().cpu()
    col = (radiance_field[ray_idx[keep], depth_idx[keep], :3] * 255).detach().cpu()
    with open(path, "w") as fid:
        fid.write("ply\nformat ascii 1.0\nelement vertex %d\n" % total)
        fid.write("property float x\nproperty float y\nproperty float z\n")
        fid.write("property uchar red\nproperty uchar green\nproperty uchar blue\nend_header\n")
        for p, c in zip(pts.tolist(), col.tolist()):
            fid.write("%f %f %f %d  %d %d\n" % (p[0], p[1], p[2], c[0], c[1], c[2]))


def positional_encoding(tensor, num_encoding_functions=6, include_input=True, log_sampling=True) -> torch.Tensor:
    """ref: nerf/nerf_helpers.py:305-349."""
    if not log_sampling:
        raise RuntimeError("only log_sampling=True is implemented (every shipped config uses it)")
    if num_encoding_functions == 0 and include_input:
        return tensor
    return torch.ops.sahs_b200.positional_encoding(tensor, int(num_encoding_functions), bool(include_input))


def get_embedding_function(num_encoding_functions=6, include_input=True, log_sampling=True):
    """ref: nerf/nerf_helpers.py:352-359."""
    return lambda x: positional_encoding(x, num_encoding_functions, include_input, log_sampling)


def sample_pdf_2(bins, weights, num_samples, det=False):
    """ref: nerf/nerf_helpers.py:454-497.  bins [R,nb], weights [R,nb-1] -> samples [R,num_samples]."""
    R = bins.shape[0]
    u = None if det else torch.rand(R, num_samples, dtype=torch.float32, device=bins.device)
    return torch.ops.sahs_b200.sample_pdf(bins, weights, int(num_samples), u)


sample_pdf = sample_pdf_2


def img2mse(img_src, img_tgt):
    return torch.nn.functional.mse_loss(img_src, img_tgt)


def mse2psnr(mse):
    import math
    if mse == 0:
        mse = 1e-5
    return -10.0 * math.log10(mse)
