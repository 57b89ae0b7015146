"""CPU tests (-m "not gpu") of the drop-in boundary: the `nerf` alias package satisfies the
reference scripts' import statements verbatim, and the host-side helpers it adds behave like the reference's."""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest
import torch

import sahs_fixtures as FX
from oracle import ref_harness as RH
from oracle import sahs_oracle as O

PKG = os.path.join(FX.REPO, "sahs-deformable-nerf_b200")

# the import statements of the Stage-I scripts, verbatim (ref: eval_stage_rays.py:28-39, train_stage_rays_auto.py:19-23)
# and the package-level names of nerf/__init__.py:1-10
SCRIPT_IMPORTS = '''
from nerf import (
    CfgNode,
    get_ray_bundle,
    get_ray_bundle_by_mask,
    load_flame_data,
    load_llff_data,
    models,
    get_embedding_function,
    run_one_iter_of_nerf,
    meshgrid_xy,
    utils
)
from nerf.load_flame import load_flame_data

from nerf import (CfgNode, get_embedding_function, get_ray_bundle, get_ray_bundle_by_mask, img2mse,
                  load_llff_data, meshgrid_xy, models, utils, MaskMSELoss,
                  mse2psnr, run_one_iter_of_nerf, dump_rays, GaussianSmoothing, MaskCrossEntropyLoss)
from nerf.cfgnode import CfgNode
from nerf.load_blender import load_blender_data
from nerf.load_llff import load_llff_data
import nerf
for name in ("predict_and_render_radiance", "run_network", "volume_render_radiance_field", "sample_pdf", "sample_pdf_2",
             "positional_encoding", "cumprod_exclusive", "get_minibatches", "AudioFaceModel", "NeRFaceModel", "AudioNet",
             "pose_to_euler_trans", "rot_to_euler"):
    assert hasattr(nerf, name), name
assert callable(utils.shrink) and callable(utils.label2color) and callable(utils.color2label_np)
model = getattr(models, "AudioFaceModel")
# Stage II inference scripts (ref: eval_get_texture_photo_audio.py:22,36,154-159, eval_get_texture_photo_3dmm.py:36,125)
from nerf._init_spade import *
from nerf._init_spade import Generator as Generator
G = Generator_audio()
assert hasattr(G, "load_state_dict") and hasattr(G, "refine_network") and hasattr(G, "idencoder") and hasattr(G, "AudioNet")
try:
    Discriminator()
    raise SystemExit("Discriminator must raise")
except NotImplementedError:
    pass
print("OK")
'''


def test_reference_scripts_import_lines_resolve_to_the_dropin():
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(PKG, "dropin"), PKG])
    env.pop("SAHS_REFERENCE_NERF", None)
    r = subprocess.run([sys.executable, "-c", textwrap.dedent(SCRIPT_IMPORTS)], env=env, capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), r.stderr[-2000:]


def test_loader_shims_raise_without_the_reference_tree(monkeypatch):
    import sahs_b200
    monkeypatch.delenv("SAHS_REFERENCE_NERF", raising=False)
    for fn in (sahs_b200.load_flame_data, sahs_b200.load_llff_data, sahs_b200.load_blender_data):
        with pytest.raises(RuntimeError, match="outside the B200 hot path"):
            fn("/nonexistent")


def _ref():
    if not RH.reference_available():
        pytest.skip("reference tree not present on this machine")
    return RH.import_reference()


def test_utils_match_reference():
    nerf = _ref()
    from nerf import utils as RU
    from sahs_b200 import utils as U
    rng = np.random.default_rng(0)
    scores = rng.random((9, 7, 12)).astype(np.float32)
    assert np.array_equal(U.shrink(scores), RU.shrink(scores))
    labels = rng.integers(0, 12, (9, 7))
    img = U.PALETTE[labels].copy()
    img[0, 0] = [1, 2, 3]                                   # not a palette colour: stays all-zero
    assert np.array_equal(U.color2label_np(img), RU.color2label_np(img))
    t = torch.from_numpy(scores)
    assert torch.equal(U.label2color(t), RU.label2color(t))


def test_gaussian_smoothing_matches_reference():
    nerf = _ref()
    import sahs_b200
    for ch, k, sig, dim, shape in ((3, 11, 2.0, 2, (2, 3, 20, 17)), (1, 11, 1.5, 2, (1, 1, 16, 16)), (2, 11, 3.0, 1, (1, 2, 40))):
        a, b = sahs_b200.GaussianSmoothing(ch, k, sig, dim), nerf.GaussianSmoothing(ch, k, sig, dim)
        assert torch.allclose(a.weight, b.weight, rtol=1e-6, atol=1e-9)
        x = torch.randn(shape, generator=torch.Generator().manual_seed(1))
        assert torch.allclose(a(x), b(x), rtol=1e-5, atol=1e-7)


def test_dump_rays_matches_reference(tmp_path, monkeypatch):
    nerf = _ref()
    import sahs_b200
    gen = torch.Generator().manual_seed(2)
    raw = torch.rand(64, 80, 16, generator=gen)
    raw[..., 3] = torch.randn(64, 80, generator=gen) * 12 + 12
    pts = torch.randn(64, 80, 3, generator=gen)
    monkeypatch.chdir(tmp_path)
    nerf.dump_rays(None, pts, raw)                            # writes rays_small.ply into the cwd
    want = open(tmp_path / "rays_small.ply").read()
    sahs_b200.dump_rays(None, pts, raw, path=str(tmp_path / "ours.ply"))
    assert open(tmp_path / "ours.ply").read() == want and want.count("\n") > 12


def test_oracle_ray_bundle_by_mask_matches_reference():
    nerf = _ref()
    pose = FX.make_pose(3, 0.78, 10.0)
    intr = np.array([37.5, 40.0, 0.48, 0.53])
    mask = (torch.rand(16, 16, generator=torch.Generator().manual_seed(4)) > 0.4).float()
    ro, rd = nerf.get_ray_bundle_by_mask(16, 16, intr, pose, mask)
    ro_o, rd_o = O.get_ray_bundle_by_mask(16, 16, list(intr), pose, mask)
    assert torch.equal(ro, ro_o) and torch.equal(rd, rd_o)


def _reference_script_function(script, name):
    """A function of one of the reference's CLI scripts, which cannot be imported here (matplotlib / imageio at module
    level): its source segment is compiled on its own, with the names it needs taken from the reference's `nerf` package."""
    import ast
    nerf = _ref()
    path = os.path.join(RH.REF_ROOT, script)
    src = open(path).read()
    tree = ast.parse(src)
    node = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == name)
    ns = {"torch": torch, "np": np, "meshgrid_xy": nerf.meshgrid_xy}
    exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), ns)
    return ns[name]


def test_oracle_normal_map_matches_reference():
    ref_fn = _reference_script_function("eval_stage_rays.py", "torch_normal_map")
    gen = torch.Generator().manual_seed(6)
    n = 24
    depth = 0.5 + 0.2 * torch.rand(n, n, generator=gen)
    weights = torch.rand(n, n, generator=gen) * 0.5
    focal = np.array([1200.0 * n / 512, 1150.0 * n / 512, 0.5, 0.48])
    for central in (False, True):
        for w in (None, weights):
            want = ref_fn(depth.clone(), focal, weights=None if w is None else w.clone(), clean=True, central_difference=central)
            got = O.torch_normal_map(depth, list(focal), w, True, central)
            assert got.shape == want.shape
            assert float((got - want).abs().max()) <= 2e-3, (central, w is None)      # values in [0, 255]
