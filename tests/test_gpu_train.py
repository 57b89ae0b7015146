"""GPU gradient-parity tests (-m gpu): one Stage-I training step (forward with tapes, hand-written compositing and
field backward kernels, tape GEMMs) against torch autograd through the CPU oracle on the same rays and loss.

Tolerances (per parameter tensor, cosine similarity with the fp32 autograd gradient and norm ratio).  The CUDA path
returns the exact gradient of *its own* forward (fp16 operands); against the fp32 reference the difference is
dominated by ReLU / LeakyReLU units whose pre-activation sign flips under the forward's ~1e-3 rounding, an error
that grows layer by layer going down the backward chain (measured with scripts/gpu_dmap_debug.py: cos 0.9999 at the
last trunk layer -> 0.9963 at the first -> 0.997 in the deformation nets; switching the gradient chain from bf16
to fp16 operands did not change it).
  audio config (10 octaves): heads / trunk / hyper / grid >= 0.999, everything >= 0.99, norm ratio within 10 %.
  expression/person_2 (15 octaves): the encoding of the warped point is ill-conditioned (2^14 gain, see
  tests/test_oracle_golden.py::test_fine_pass_conditioning) and training uses the merged fp16 deformation phase, so
  only the radiance MLPs are held to a bar (heads >= 0.98, trunk >= 0.9); the deformation nets must stay positively
  correlated (>= 0.5) -- a documented limitation."""
import os

import numpy as np
import pytest
import torch

import sahs_fixtures as FX
from oracle import sahs_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _setup(cfg_name, H, W, seed):
    import sahs_b200
    cfg = FX.load_cfg(cfg_name)
    cfg.nerf.train.perturb, cfg.nerf.train.radiance_field_noise_std = False, 0.0
    spec = O.spec_from_cfg(cfg)
    sd = FX.make_state_dict(spec, seed=42, dense=True)
    fr = FX.make_frame_inputs(spec, H, W, seed=seed)
    gen = torch.Generator().manual_seed(77)
    target = torch.rand(H * W, 3, generator=gen)
    mask = fr["mask"].view(-1, 12).float()
    return sahs_b200, cfg, spec, sd, fr, target, mask


@pytest.mark.parametrize("cfg_name", ["audio/person_2_auto", "expression/person_2"])
def test_train_step_gradients_vs_oracle_autograd(cfg_name):
    H, W = 6, 8
    sahs, cfg, spec, sd, fr, target, mask = _setup(cfg_name, H, W, seed=4)
    # ---- oracle: autograd through the CPU restatement ----
    sd_ref = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    opts = O.opts_from_cfg(cfg, "train")
    ro, rd = O.get_ray_bundle(H, W, fr["intrinsics"], fr["pose"])
    drv_in = fr["driving"]
    out_ref = O.run_one_iter(sd_ref, spec, opts, ro, rd, drv_in, fr["pose"], fr["background"].view(-1, 15))
    loss_ref, _ = O.stage1_loss(out_ref[0], out_ref[3], target, mask)
    loss_ref.backward()
    # ---- ours ----
    model = getattr(sahs.models, cfg.models.mask.type)(cfg)
    model.load_state_dict(sd)
    model = model.to(DEV)
    pose = fr["pose"].to(DEV)
    with torch.no_grad():
        ro_g, rd_g = sahs.get_ray_bundle(H, W, fr["intrinsics"], pose)
    out = sahs.run_one_iter_of_nerf(H, W, fr["intrinsics"][0], model, ro_g, rd_g, cfg, mode="train",
                                    driving=drv_in.to(DEV), pose=pose,
                                    background_prior=fr["background"].view(-1, 15).to(DEV), inHead=fr["mask"].to(DEV))
    loss, sample_prob = sahs.stage1_loss(out[0], out[3], target.to(DEV), mask.to(DEV))
    loss.backward()
    torch.cuda.synchronize()
    from sahs_b200 import ops
    assert ops.field_status()[0] == 0
    assert abs(float(loss.detach()) - float(loss_ref.detach())) <= 2e-3 * max(1.0, abs(float(loss_ref.detach())))
    assert sample_prob.shape == (12,)
    bad = {}
    ill = spec.xyz_L > 10
    for name, p in model.named_parameters():
        g_ref = sd_ref[name].grad
        assert p.grad is not None and g_ref is not None, name
        g = p.grad.detach().cpu().double().reshape(-1)
        r = g_ref.double().reshape(-1)
        if float(r.abs().max()) < 1e-12:
            assert float(g.abs().max()) < 1e-9, name       # e.g. a coarse net that only sees empty space
            continue
        cos = float(torch.dot(g, r) / (g.norm() * r.norm() + 1e-30))
        ratio = float(g.norm() / r.norm())
        head = any(t in name for t in ("layers_dir", "layers_seg", "fc_rgb", "fc_seg"))
        if not ill:
            need = 0.999 if ("nerf_mlps" in name or "hyper" in name or "spatial" in name) else 0.99
            if p.numel() <= 16 and need < 0.999:
                # a 3-element deformation-head bias summed over this test's 48 rays: a handful of ReLU sign flips
                # (forward rounding) moves its direction by ~1e-2; measured 0.989-0.996 across builds
                need = 0.98
            ok = cos >= need and 0.9 <= ratio <= 1.12
        else:
            need = 0.98 if head else (0.9 if "nerf_mlps" in name else 0.5)
            ok = cos >= need
        if not ok:
            bad[name] = (cos, ratio, need)
    assert not bad, bad


@pytest.mark.parametrize("fused", [False, True])
def test_optimizer_step_repacks_weights(fused):
    """After an in-place parameter update the packed images are rebuilt (stale-weight guard) and the loss moves.  Fused
    optimizers do not bump Tensor._version: the optimizer-step hook in models.py covers them."""
    sahs, cfg, spec, sd, fr, target, mask = _setup("audio/person_2_auto", 4, 8, seed=6)
    model = sahs.AudioFaceModel(cfg)
    model.load_state_dict(sd)
    model = model.to(DEV)
    opt = torch.optim.Adam(model.parameters(), lr=5e-4, fused=fused)
    pose = fr["pose"].to(DEV)
    with torch.no_grad():
        ro, rd = sahs.get_ray_bundle(4, 8, fr["intrinsics"], pose)
    losses = []
    for _ in range(3):
        out = sahs.run_one_iter_of_nerf(4, 8, 1.0, model, ro, rd, cfg, mode="train", driving=fr["driving"].to(DEV),
                                        pose=pose, background_prior=fr["background"].view(-1, 15).to(DEV))
        loss, _ = sahs.stage1_loss(out[0], out[3], target.to(DEV), mask.to(DEV))
        opt.zero_grad()
        loss.backward()
        opt.step()
        losses.append(float(loss.detach()))
    assert losses[2] < losses[0], losses


def test_stage1_loss_kernel_vs_reference_golden():
    """sahs_stage1_loss (one launch: loss, statistics, sample_prob, gradients) against the reference's own values
    (tests/golden/stage1_loss.npz, made by oracle/make_golden.py from the unmodified modules).  fp32 sums in a
    different order: 2e-6 relative on the scalars, 1e-6 * max|g| on the gradients."""
    import numpy as np
    import sahs_b200
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "stage1_loss.npz"))
    mc = torch.from_numpy(g["map_c"]).to(DEV).requires_grad_(True)
    mf = torch.from_numpy(g["map_f"]).to(DEV).requires_grad_(True)
    tg, mk = torch.from_numpy(g["target"]).to(DEV), torch.from_numpy(g["mask"]).to(DEV)
    loss, prob, stats = sahs_b200.stage1_loss(mc, mf, tg, mk, return_stats=True)
    (3.0 * loss).backward()                                  # upstream gradient != 1
    torch.cuda.synchronize()
    assert abs(float(loss.detach()) - float(g["ref_loss"])) <= 2e-6 * abs(float(g["ref_loss"]))
    assert np.allclose(prob.cpu().numpy(), g["ref_prob"], rtol=2e-5, atol=1e-8)
    assert np.allclose(stats.cpu().numpy(), g["ref_stats"], rtol=2e-5, atol=1e-8)
    for got, want in ((mc.grad, g["ref_d_coarse"]), (mf.grad, g["ref_d_fine"])):
        want = 3.0 * torch.from_numpy(want)
        assert float((got.cpu() - want).abs().max()) <= 1e-6 * float(want.abs().max())
    # the module-by-module form (what the reference's script writes) agrees, on the device too
    l2, p2 = sahs_b200.stage1_loss_modules(mc.detach(), mf.detach(), tg, mk)
    assert abs(float(l2) - float(loss)) <= 2e-6 * abs(float(loss)) and torch.allclose(p2, prob, rtol=2e-5, atol=1e-8)
    # deterministic: a second launch gives the same bits
    loss_b, prob_b = sahs_b200.stage1_loss(mc.detach(), mf.detach(), tg, mk)
    assert float(loss_b) == float(loss) and torch.equal(prob_b, prob)


def test_stage1_loss_kernel_full_batch_vs_oracle():
    """BASELINE batch size (2048 rays) with random one-hot labels, against the oracle's autograd."""
    import sahs_b200
    gen = torch.Generator().manual_seed(8)
    R = 2048
    mc = torch.rand(R, 15, generator=gen)
    mf = torch.rand(R, 15, generator=gen)
    mc[:, 3:] = torch.softmax(torch.randn(R, 12, generator=gen) * 3, -1)
    mf[:, 3:] = torch.softmax(torch.randn(R, 12, generator=gen) * 3, -1)
    tg = torch.rand(R, 3, generator=gen)
    mk = torch.nn.functional.one_hot(torch.randint(0, 12, (R,), generator=gen), 12).float()
    oc, of = mc.clone().requires_grad_(True), mf.clone().requires_grad_(True)
    lo, po = O.stage1_loss(oc, of, tg, mk)
    lo.backward()
    gc, gf = mc.to(DEV).requires_grad_(True), mf.to(DEV).requires_grad_(True)
    lg, pg = sahs_b200.stage1_loss(gc, gf, tg.to(DEV), mk.to(DEV))
    lg.backward()
    assert abs(float(lg.detach()) - float(lo.detach())) <= 2e-6 * abs(float(lo.detach()))
    assert torch.allclose(pg.cpu(), po, rtol=2e-5, atol=1e-8)
    for got, want in ((gc.grad, oc.grad), (gf.grad, of.grad)):
        assert float((got.cpu() - want).abs().max()) <= 1e-6 * float(want.abs().max())
    with pytest.raises(RuntimeError):
        sahs_b200.stage1_loss(mc, mf, tg, mk)                # CPU tensors: no fallback


def test_flat_adam_matches_torch_adam():
    """FlatAdam (one sahs_adam_step launch on a flat buffer) against torch.optim.Adam's single-tensor CPU path -- the
    optimizer the reference constructs (train_stage_rays_auto.py:200-210) -- over several steps with odd tensor
    sizes, a learning rate rewritten between steps (:503-509) and a parameter that gets no gradient at first.
    fp32 with the same operation order: 2e-6 relative on the parameters after 6 steps."""
    import sahs_b200
    gen = torch.Generator().manual_seed(12)
    shapes = [(128, 175), (3,), (1, 32, 7, 5, 3), (1,), (64, 239), (13,)]
    ref = [torch.randn(s, generator=gen).requires_grad_(True) for s in shapes]
    ours = [r.detach().clone().to(DEV).requires_grad_(True) for r in ref]
    o_ref = torch.optim.Adam(ref, lr=5e-4, foreach=False, fused=False)
    o_new = sahs_b200.FlatAdam(ours, lr=5e-4)
    assert all(p.data_ptr() >= o_new.flat_param.data_ptr() for p in ours)          # parameters re-homed into the buffer
    for it in range(6):
        o_ref.zero_grad(set_to_none=True)
        o_new.zero_grad(set_to_none=True)
        for k, (r, p) in enumerate(zip(ref, ours)):
            if k == 3 and it < 2:
                continue                                             # no gradient for this one yet
            g = torch.randn(r.shape, generator=gen) * (10.0 ** (k - 3))
            r.grad = g.clone()
            p.grad = g.to(DEV)
        o_ref.step()
        o_new.step()
        lr = sahs_b200.exp_lr(5e-4, 0.1, 3.0, it + 1)
        o_ref.param_groups[0]["lr"] = lr
        o_new.param_groups[0]["lr"] = lr
    torch.cuda.synchronize()
    for k, (r, p) in enumerate(zip(ref, ours)):
        if k == 3:
            continue     # torch skips a parameter without gradient (its step count lags); FlatAdam treats it as g = 0
        err = float((p.detach().cpu() - r.detach()).abs().max())
        assert err <= 2e-6 * float(r.detach().abs().max()) + 1e-7, (k, err)
        st = o_new.state[p]
        for name in ("exp_avg", "exp_avg_sq"):      # moments: 1e-6 of the tensor's largest entry (sums of signed terms)
            want = o_ref.state[r][name]
            assert float((st[name].cpu() - want).abs().max()) <= 1e-6 * float(want.abs().max()), (k, name)
    assert float(o_new.state[ours[0]]["step"]) == 6.0
    with pytest.raises(RuntimeError):
        sahs_b200.FlatAdam([torch.zeros(4, requires_grad=True)], lr=1e-3)          # CPU parameters: no fallback


def test_flat_adam_training_loop_repacks_and_descends():
    """End to end with the model: parameters re-homed into the flat buffer still drive the packed fp16 weight images
    (the cache keys on data_ptr and on the optimizer-step epoch), and the loss goes down."""
    sahs, cfg, spec, sd, fr, target, mask = _setup("audio/person_2_auto", 4, 8, seed=6)
    model = sahs.AudioFaceModel(cfg)
    model.load_state_dict(sd)
    model = model.to(DEV)
    opt = sahs.FlatAdam(model.parameters(), lr=5e-4)
    pose = fr["pose"].to(DEV)
    with torch.no_grad():
        ro, rd = sahs.get_ray_bundle(4, 8, fr["intrinsics"], pose)
    losses = []
    for _ in range(4):
        out = sahs.run_one_iter_of_nerf(4, 8, 1.0, model, ro, rd, cfg, mode="train", driving=fr["driving"].to(DEV),
                                        pose=pose, background_prior=fr["background"].view(-1, 15).to(DEV))
        loss, _ = sahs.stage1_loss(out[0], out[3], target.to(DEV), mask.to(DEV))
        opt.zero_grad()
        loss.backward()
        opt.step()
        losses.append(float(loss.detach()))
    assert losses[-1] < losses[0], losses
    sd2 = model.state_dict()
    assert all(torch.isfinite(v).all() for v in sd2.values())
