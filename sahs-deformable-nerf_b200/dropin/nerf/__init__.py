"""Drop-in alias: put `sahs-deformable-nerf_b200/dropin` and `sahs-deformable-nerf_b200` on sys.path ahead of the
reference tree and the `from nerf import (...)` / `from nerf.load_flame import load_flame_data` lines of
eval_stage_rays.py:28-39 and train_stage_rays_auto.py:19-23 resolve to the B200-native path.
Mirrors the reference's nerf/__init__.py:1-10 (star re-exports of models, modules, nerf_helpers, train_utils,
volume_rendering_utils plus CfgNode and the three loaders)."""
from sahs_b200 import *  # noqa: F401,F403
from sahs_b200 import (CfgNode, load_blender_data, load_flame_data, load_llff_data, models, nerf_helpers,  # noqa: F401
                       train_utils, utils, volume_rendering_utils)
from sahs_b200.models import AudioNet, pose_to_euler_trans, rot_to_euler  # noqa: F401
from . import cfgnode, load_blender, load_flame, load_llff  # noqa: F401,E402
