"""Reference harness -- TEST INFRASTRUCTURE ONLY (never imported by the product path).

Imports the unmodified reference package from /root/reference/nerf-pytorch so that
(a) oracle/sahs_oracle.py can be pinned against it and (b) golden vectors can be generated
(oracle/make_golden.py).  /root/reference only exists in the build container, never on the
GPU box, so nothing in `-m gpu` tests, smoke() or bench.py may import this module.

The reference imports two third-party modules that are absent offline and unused on the hot path
(`imageio` in nerf/load_blender.py:5, `pytorch3d.transforms` in nerf/nerf_helpers.py:4); empty stub
modules are registered for them.  The reference tree itself is untouched.
"""
import contextlib
import os
import sys
import types

REF_ROOT = os.environ.get("SAHS_REFERENCE_ROOT", "/root/reference/nerf-pytorch")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "nerf"))


def import_reference():
    """Return the reference `nerf` package (imported under its own name `nerf`)."""
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    for name in ("imageio", "pytorch3d", "pytorch3d.transforms"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["pytorch3d"].transforms = sys.modules["pytorch3d.transforms"]
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import nerf  # noqa: the reference package
    assert os.path.realpath(nerf.__file__).startswith(os.path.realpath(REF_ROOT)), nerf.__file__
    return nerf


def load_reference_cfg(rel_yaml: str):
    """YAML -> reference CfgNode (eval_stage_rays.py:265-267)."""
    import yaml
    nerf = import_reference()
    with open(os.path.join(REF_ROOT, rel_yaml), "r") as f:
        return nerf.CfgNode(yaml.load(f, Loader=yaml.FullLoader))


@contextlib.contextmanager
def relu_clone_patch():
    """Neutral patch needed only for *backward* runs of the reference under torch>=2:
    volume_rendering_utils.py:56-57 mutates a ReLU output in place."""
    import torch
    orig = torch.nn.functional.relu
    torch.nn.functional.relu = lambda x, inplace=False: orig(x).clone()
    try:
        yield
    finally:
        torch.nn.functional.relu = orig
