"""Drop-in alias: put `sahs-deformable-nerf_b200/dropin` and `sahs-deformable-nerf_b200` on sys.path ahead of the
reference tree and `from nerf import (...)` in eval_stage_rays.py / train_stage_rays_auto.py resolves to the
B200-native path."""
from sahs_b200 import *  # noqa: F401,F403
from sahs_b200 import models, nerf_helpers, train_utils, volume_rendering_utils  # noqa: F401
