"""sahs_b200 -- B200-native (sm_100a) per-ray render/train hot path of SAHS-Deformable-Nerf.

Exports the names the reference's Stage-I scripts import from `nerf` (ref: nerf/__init__.py:1-10,
eval_stage_rays.py:28-39, train_stage_rays_auto.py:21-23) for the hot path."""
from . import custom_ops  # noqa: F401  (registers torch.ops.sahs_b200.*)
from . import models  # noqa: F401
from .cfgnode import CfgNode  # noqa: F401
from .configs import builtin_config  # noqa: F401
from .losses import MaskCrossEntropyLoss, MaskMSELoss, stage1_loss, stage1_loss_modules  # noqa: F401
from .models import AudioFaceModel, NeRFaceModel  # noqa: F401
from . import utils  # noqa: F401
from ._loaders import load_blender_data, load_flame_data, load_llff_data  # noqa: F401
from .nerf_helpers import (cumprod_exclusive, dump_rays, get_embedding_function, get_minibatches, get_ray_bundle,  # noqa: F401
                           get_ray_bundle_by_mask, img2mse, meshgrid_xy, mse2psnr, positional_encoding, sample_pdf,
                           sample_pdf_2)
from .train_utils import GaussianSmoothing, predict_and_render_radiance, run_network, run_one_iter_of_nerf  # noqa: F401
from .volume_rendering_utils import volume_render_radiance_field  # noqa: F401


def frame_postprocess(rgb_map):
    """[..., 15] composited map -> (rgb uint8 [...,3], label uint8 [...], palette colour uint8 [...,3]);
    ref: eval_stage_rays.py:221-227 (cast_to_image), nerf/utils.py:112-140 (label2color)."""
    import torch
    return torch.ops.sahs_b200.frame_postprocess(rgb_map)


def weighted_sample(mask, class_prob, num_select, seed):
    """Semantic-weighted ray batch without replacement on the device (ref: train_stage_rays_auto.py:390-420)."""
    import torch
    return torch.ops.sahs_b200.weighted_sample(mask, class_prob, int(num_select), int(seed) & 0x7FFFFFFFFFFFFFFF)
from .optim import FlatAdam, exp_lr  # noqa: F401,E402
from .train import audit_fp16_range  # noqa: F401,E402



def torch_normal_map(depthmap, focal, weights=None, clean=True, central_difference=False):
    """ref: eval_stage_rays.py:116-151 (same signature and result: [N-k, N-k, 3] fp32 in [0, 255])."""
    import torch
    fx, fy, cx, cy = [float(v) for v in focal]
    w = weights if (clean and weights is not None) else None
    return torch.ops.sahs_b200.normal_map(depthmap, fx, fy, cx, cy, w, bool(central_difference))
