"""sahs_b200 -- B200-native (sm_100a) per-ray render/train hot path of SAHS-Deformable-Nerf."""
from .cfgnode import CfgNode  # noqa: F401
from .configs import builtin_config  # noqa: F401
