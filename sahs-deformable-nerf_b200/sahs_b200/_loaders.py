"""Dataset loaders are outside the hot path (out of scope) and are NOT re-implemented.  The
reference's Stage-I scripts import them from `nerf`, so the drop-in package must provide the names: each shim forwards
to the reference's own loader module when the reference tree is available (environment variable
SAHS_REFERENCE_NERF = path of the reference's `nerf` directory) and raises a clear error otherwise."""
from __future__ import annotations

import importlib.util
import os
import sys


def _reference_loader(module: str, func: str):
    root = os.environ.get("SAHS_REFERENCE_NERF", "")
    path = os.path.join(root, module + ".py")
    if not root or not os.path.isfile(path):
        raise RuntimeError(
            f"{func} is a dataset loader of the reference (nerf/{module}.py), outside the B200 hot path; set "
            "SAHS_REFERENCE_NERF=/path/to/reference/nerf-pytorch/nerf to forward to the reference's implementation")
    name = "_sahs_ref_" + module
    mod = sys.modules.get(name)
    if mod is None:
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
    return getattr(mod, func)


def load_flame_data(*args, **kwargs):
    """ref: nerf/load_flame.py (frames, poses, expressions, backgrounds of a FLAME-tracked capture)."""
    return _reference_loader("load_flame", "load_flame_data")(*args, **kwargs)


def load_llff_data(*args, **kwargs):
    """ref: nerf/load_llff.py."""
    return _reference_loader("load_llff", "load_llff_data")(*args, **kwargs)


def load_blender_data(*args, **kwargs):
    """ref: nerf/load_blender.py."""
    return _reference_loader("load_blender", "load_blender_data")(*args, **kwargs)
