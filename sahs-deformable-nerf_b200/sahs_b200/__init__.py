"""sahs_b200 -- B200-native (sm_100a) per-ray render/train hot path of SAHS-Deformable-Nerf.

Exports the names the reference's Stage-I scripts import from `nerf` (ref: nerf/__init__.py:1-10,
eval_stage_rays.py:28-39, train_stage_rays_auto.py:21-23) for the hot path."""
from . import models  # noqa: F401
from .cfgnode import CfgNode  # noqa: F401
from .configs import builtin_config  # noqa: F401
from .losses import MaskCrossEntropyLoss, MaskMSELoss, stage1_loss, stage1_loss_modules  # noqa: F401
from .models import AudioFaceModel, NeRFaceModel  # noqa: F401
from .nerf_helpers import (cumprod_exclusive, get_embedding_function, get_minibatches, get_ray_bundle,  # noqa: F401
                           img2mse, meshgrid_xy, mse2psnr, positional_encoding, sample_pdf, sample_pdf_2)
from .train_utils import predict_and_render_radiance, run_network, run_one_iter_of_nerf  # noqa: F401
from .volume_rendering_utils import volume_render_radiance_field  # noqa: F401
from .ops import frame_postprocess, weighted_sample  # noqa: F401,E402
from .optim import FlatAdam, exp_lr  # noqa: F401,E402
