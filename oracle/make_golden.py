"""Pin the oracle against the live reference and freeze golden vectors -- TEST INFRASTRUCTURE.

Run in the build container only (needs /root/reference):

    python -m oracle.make_golden            # writes tests/golden/*.npz, prints oracle-vs-reference errors

For every case the *unmodified reference* (oracle/ref_harness.py) is executed on seeded synthetic inputs
(tests/sahs_fixtures.py); the oracle restatement must reproduce it (bit-exact for sample_pdf indices and
z-values, <=2e-5 abs for float maps) before the vectors are written.  The GPU box has no reference tree:
its tests regenerate the same seeded inputs and compare the CUDA path with these files and with the oracle.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(_HERE)
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))

from oracle import ref_harness as RH  # noqa: E402
from oracle import sahs_oracle as O  # noqa: E402
import sahs_fixtures as FX  # noqa: E402

GOLD = os.path.join(REPO, "tests", "golden")


def _np(t):
    return t.detach().cpu().numpy()


def _maxabs(a, b):
    return float((a.double() - b.double()).abs().max())


def build_reference_model(nerf, rcfg, sd):
    model = getattr(nerf.models, rcfg.models.mask.type)(rcfg)
    missing = model.load_state_dict(sd, strict=True)
    model.eval()
    return model


def e2e_case(nerf, name, cfg_name, H, W, seed, pose_z, mode, stochastic=False, keep_rays=24, trained_like=True,
             bg_mode="prior"):
    """End-to-end run_one_iter_of_nerf (ref: nerf/train_utils.py:209-321) vs oracle.run_one_iter.
    bg_mode: "prior" (background_prior given, the shipped eval/train scripts), "none" (background_prior=None: every
    channel of every sample goes through the sigmoid, ref: nerf/volume_rendering_utils.py:35) or "white" (none +
    white_background=True, ref: :75-76)."""
    rcfg = RH.load_reference_cfg(f"config/{cfg_name}.yml")
    spec = O.spec_from_cfg(rcfg)
    sd = FX.make_state_dict(spec, seed=42, dense=True, trained_like=trained_like)
    fr = FX.make_frame_inputs(spec, H, W, seed=seed, pose_z=pose_z)
    node = getattr(rcfg.nerf, mode)
    node.white_background = (bg_mode == "white")
    if not stochastic:
        node.perturb = False
        node.radiance_field_noise_std = 0.0
    else:
        node.perturb = True
        node.radiance_field_noise_std = 0.1
    model = build_reference_model(nerf, rcfg, sd)
    pose = fr["pose"]
    ro, rd = nerf.get_ray_bundle(H, W, np.array(fr["intrinsics"]), pose)
    bg = fr["background"].view(-1, 15) if bg_mode == "prior" else None
    R = H * W
    draws = {}
    with torch.no_grad():
        torch.manual_seed(777 + seed)
        ref = nerf.run_one_iter_of_nerf(H, W, fr["intrinsics"][0], model, ro, rd, rcfg, mode=mode,
                                        driving=fr["driving"], pose=pose, pose_c=None, background_prior=bg,
                                        inHead=fr["mask"])
        if stochastic:
            # replay the reference's draw order: t_rand (train_utils.py:112), coarse noise
            # (volume_rendering_utils.py:47), u (nerf_helpers.py:473), fine noise.
            torch.manual_seed(777 + seed)
            draws["t_rand"] = torch.rand(R, node.num_coarse)
            draws["noise_c"] = torch.randn(R, node.num_coarse) * node.radiance_field_noise_std
            draws["u"] = torch.rand(R, node.num_fine)
            draws["noise_f"] = torch.randn(R, node.num_coarse + node.num_fine) * node.radiance_field_noise_std
    ref = [r.reshape(R, 15) if r.shape[-1] == 15 and r.numel() == R * 15 else r.reshape(R) for r in ref]
    opts = O.opts_from_cfg(rcfg, mode)
    ro_o, rd_o = O.get_ray_bundle(H, W, fr["intrinsics"], pose)
    assert torch.equal(ro_o.contiguous(), ro.contiguous()) and _maxabs(rd_o, rd) == 0.0, "ray bundle mismatch"
    with torch.no_grad():
        drv = O.driving_vector(sd, spec, fr["driving"])
        out, aux = O.render_rays(sd, spec, opts, ro.reshape(-1, 3), rd.reshape(-1, 3), drv, pose, bg,
                                 return_aux=True, **draws)
    names = ["rgb_c", "disp_c", "acc_c", "rgb_f", "disp_f", "acc_f", "w_last_f", "depth_f"]
    errs = {n: _maxabs(a, b) for n, a, b in zip(names, out, ref)}
    print(f"[{name}] oracle vs reference max-abs:", {k: f"{v:.2e}" for k, v in errs.items()})
    sig_c = aux["raw_c"][:, :-1, -1]
    frac_pos, w_last_c = float((sig_c > 0).float().mean()), float(aux["w_c"][:, -1].mean())
    print(f"[{name}] coarse pass: {100 * frac_pos:.0f} % of samples with sigma > 0, mean background weight "
          f"{w_last_c:.3f}, depth_c in [{float(aux['depth_c'].min()):.3f}, {float(aux['depth_c'].max()):.3f}], "
          f"rgb_f in [{float(out[3][:, :3].min()):.2f}, {float(out[3][:, :3].max()):.2f}]")
    if trained_like and bg_mode == "prior":
        # the fixture must exercise the coarse pass (the old audio fixture rendered pure background)
        assert frac_pos > 0.5 and w_last_c < 0.9, (name, frac_pos, w_last_c)
    tol = dict(rgb_c=3e-5, rgb_f=3e-5, acc_c=1e-5, acc_f=1e-5, w_last_f=3e-5, depth_c=3e-5, depth_f=3e-5)
    for k, v in errs.items():
        if k.startswith("disp"):
            continue                      # 1/depth amplifies rounding; checked relatively below
        assert v <= tol.get(k, 3e-5), (name, k, v)
    for i in (1, 4):
        rel = float(((out[i] - ref[i]).abs() / ref[i].abs()).max())
        assert rel < 1e-4, (name, names[i], rel)
    k = keep_rays
    save = dict(
        cfg_name=cfg_name, H=H, W=W, seed=seed, pose_z=pose_z, mode=mode, stochastic=int(stochastic),
        trained_like=int(trained_like), bg_mode=bg_mode, coarse_frac_pos=frac_pos, coarse_w_last=w_last_c,
        state_checksum=FX.state_checksum(sd), pose=_np(pose), driving_vec=_np(drv),
        **{"ref_" + n: _np(r) for n, r in zip(names, ref)},
        z_c=_np(aux["z_c"][:k]), w_c=_np(aux["w_c"]), z_s=_np(aux["z_s"]), z_f=_np(aux["z_f"]),
        raw_c=_np(aux["raw_c"][:k]), raw_f=_np(aux["raw_f"][:k]), depth_c=_np(aux["depth_c"]),
    )
    for kk, v in draws.items():
        save["draw_" + kk] = _np(v)
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **save)


def field_case(nerf, name, cfg_name, n_pts, seed):
    """model(level, x, driving, pose, pose_c) (ref: nerf/models.py:367-380, :514-528) vs oracle.field_forward."""
    rcfg = RH.load_reference_cfg(f"config/{cfg_name}.yml")
    spec = O.spec_from_cfg(rcfg)
    sd = FX.make_state_dict(spec, seed=42, dense=True)
    fr = FX.make_frame_inputs(spec, 8, 8, seed=seed, pose_z=0.78)
    model = build_reference_model(nerf, rcfg, sd)
    g = torch.Generator().manual_seed(300 + seed)
    xyz = (torch.rand(n_pts, 3, generator=g) * 2 - 1) * torch.tensor([0.35, 0.35, 0.35])
    xyz[:8] *= 4.0                                        # a few points outside the grid cube
    dirs = torch.randn(n_pts, 3, generator=g) * 0.3 + torch.tensor([0.0, 0.0, -1.0])
    x = torch.cat((xyz, dirs, torch.zeros(n_pts, 12)), -1)
    # sample_from_3dgrid reshapes by num_coarse(+num_fine): n_pts must be a multiple of 128
    assert n_pts % 128 == 0
    out = {}
    with torch.no_grad():
        drv = O.driving_vector(sd, spec, fr["driving"])
        for level in ("coarse", "fine"):
            ref = model(level, x, fr["driving"], fr["pose"], None)
            mine, inter = O.field_forward(sd, spec, level, xyz, dirs, drv, fr["pose"], return_intermediates=True)
            err = _maxabs(ref, mine)
            print(f"[{name}/{level}] oracle vs reference raw max-abs {err:.2e} (|raw| max {float(ref.abs().max()):.2f})")
            assert err < 2e-3 * max(1.0, float(ref.abs().max()) / 50), err
            out["ref_raw_" + level] = _np(ref)
            if level == "coarse":
                out["mapped"] = _np(inter["mapped"])
                if "amb" in inter:
                    out["amb"] = _np(inter["amb"])
                if "emb" in inter:
                    out["emb"] = _np(inter["emb"])
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), cfg_name=cfg_name, seed=seed, xyz=_np(xyz), dirs=_np(dirs),
                        state_checksum=FX.state_checksum(sd), pose=_np(fr["pose"]), driving_vec=_np(drv), **out)


def sample_pdf_case(nerf, name, rows, seed):
    """sample_pdf_2 (ref: nerf/nerf_helpers.py:454-497): bit-exact indices and samples."""
    g = torch.Generator().manual_seed(seed)
    z = torch.linspace(0.0, 1.0, 64) * 0.6 + 0.4838
    z = z.expand(rows, 64).contiguous()
    bins = 0.5 * (z[:, 1:] + z[:, :-1])
    # a mix of peaky (surface-like), flat and near-zero weight rows
    w = torch.rand(rows, 62, generator=g) ** 6
    peak = torch.randint(0, 62, (rows,), generator=g)
    w[torch.arange(rows), peak] += torch.rand(rows, generator=g) * 3
    w[: rows // 8] *= 1e-6
    w[rows // 8: rows // 4] = torch.rand(rows // 4 - rows // 8, 62, generator=g)
    ref = nerf.nerf_helpers.sample_pdf_2(bins, w, 64, det=True)
    # reference indices, recomputed with the reference's own ops (nerf_helpers.py:459-482)
    ww = w + 1e-5
    pdf = ww / torch.sum(ww, dim=-1, keepdim=True)
    cdf = torch.cat([torch.zeros_like(pdf[..., :1]), torch.cumsum(pdf, dim=-1)], dim=-1)
    u = torch.linspace(0.0, 1.0, steps=64).expand(rows, 64).contiguous()
    inds_ref = torch.searchsorted(cdf.contiguous(), u, right=True)
    mine, inds = O.sample_pdf(bins, w, 64, det=True, return_inds=True)
    print(f"[{name}] sample_pdf: index mismatches {(inds != inds_ref).sum().item()} / {inds.numel()}, "
          f"sample bit mismatches {(mine != ref).sum().item()}")
    assert torch.equal(inds, inds_ref), "oracle sample_pdf indices are not bit-exact vs reference"
    assert torch.equal(mine, ref), "oracle sample_pdf samples are not bit-exact vs reference"
    # stochastic variant with injected u
    torch.manual_seed(seed + 1)
    ref_s = nerf.nerf_helpers.sample_pdf_2(bins, w, 64, det=False)
    torch.manual_seed(seed + 1)
    u_s = torch.rand(rows, 64)
    mine_s = O.sample_pdf(bins, w, 64, det=False, u=u_s)
    assert torch.equal(mine_s, ref_s), "stochastic sample_pdf mismatch"
    zf_ref, _ = torch.sort(torch.cat((z, ref), -1), dim=-1)
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), z=_np(z[:1]), weights=_np(w),
                        ref_samples=_np(ref), ref_inds=_np(inds_ref).astype(np.uint8), ref_z_merged=_np(zf_ref),
                        u_s=_np(u_s), ref_samples_s=_np(ref_s))


def composite_case(nerf, name, rows, S, seed, with_bg, white):
    """volume_render_radiance_field (ref: nerf/volume_rendering_utils.py:7-78)."""
    g = torch.Generator().manual_seed(seed)
    raw = torch.randn(rows, S, 16, generator=g) * 2.0
    raw[..., -1] = torch.randn(rows, S, generator=g) * 30.0
    raw[: rows // 4, :, -1] -= 40.0                            # mostly-empty rays
    z, _ = torch.sort(torch.rand(rows, S, generator=g) * 0.6 + 0.4838, dim=-1)
    rd = torch.randn(rows, 3, generator=g) * 0.2 + torch.tensor([0, 0, -1.0])
    bg = torch.cat((torch.rand(rows, 3, generator=g), torch.ones(rows, 1), torch.zeros(rows, 11)), -1) if with_bg else None
    raw_in = raw.clone()
    if with_bg:
        raw_in[:, -1, :-1] = bg
    ref = nerf.volume_render_radiance_field(raw_in.clone(), z, rd, radiance_field_noise_std=0.0,
                                            white_background=white, background_prior=bg)
    mine = O.composite(raw_in.clone(), z, rd, None, white, bg)
    names = ["rgb", "disp", "acc", "weights", "depth"]
    errs = {n: _maxabs(a, b) for n, a, b in zip(names, mine, ref)}
    print(f"[{name}] composite oracle vs reference:", {k: f"{v:.1e}" for k, v in errs.items()})
    assert all(v == 0.0 for v in errs.values()), errs
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), raw=_np(raw), z=_np(z), rd=_np(rd),
                        bg=_np(bg) if with_bg else np.zeros(0, np.float32), with_bg=int(with_bg), white=int(white),
                        **{"ref_" + n: _np(r) for n, r in zip(names, ref)})


def helpers_case(nerf, name):
    """get_ray_bundle / positional_encoding / pose code (ref: nerf/nerf_helpers.py:178-233, :305-349;
    nerf/models.py:482-504)."""
    pose = FX.make_pose(3, 0.78, 10.0)
    intr = [37.5, 40.0, 0.48, 0.53]
    ro, rd = nerf.get_ray_bundle(12, 20, np.array(intr), pose)
    ro_o, rd_o = O.get_ray_bundle(12, 20, intr, pose)
    assert torch.equal(rd, rd_o) and torch.equal(ro.contiguous(), ro_o.contiguous())
    g = torch.Generator().manual_seed(5)
    x = torch.randn(257, 3, generator=g)
    out = {}
    for L, inc in ((10, True), (15, True), (4, True), (3, False)):
        ref = nerf.positional_encoding(x, L, inc)
        assert torch.equal(ref, O.positional_encoding(x, L, inc))
        out[f"pe_L{L}_inc{int(inc)}"] = _np(ref)
    code_ref = nerf.positional_encoding(nerf.models.pose_to_euler_trans(pose.unsqueeze(0), "cpu"), 3, False)[0]
    assert torch.equal(code_ref, O.pose_code(pose))
    print(f"[{name}] ray bundle / PE / pose code: exact")
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), pose=_np(pose), intr=np.array(intr), ro=_np(ro.contiguous()),
                        rd=_np(rd), x=_np(x), pose_code=_np(code_ref), **out)


def postprocess_case(nerf, name):
    """label2color (ref: nerf/utils.py:112-140) and the ToPILImage conversion used by cast_to_image
    (ref: eval_stage_rays.py:221-227)."""
    import torchvision
    from nerf import utils as ref_utils
    g = torch.Generator().manual_seed(9)
    m = torch.rand(37, 23, 15, generator=g) * 1.4 - 0.2
    m[..., 3:] = torch.softmax(torch.randn(37, 23, 12, generator=g) * 3, -1)
    col_ref = (ref_utils.label2color(m[..., 3:]) * 255).round().to(torch.uint8)
    img_ref = torch.from_numpy(np.array(torchvision.transforms.ToPILImage()(m[..., :3].permute(2, 0, 1).clamp(0.0, 1.0))))
    rgb, label, col = O.frame_postprocess(m)
    assert torch.equal(col, col_ref) and torch.equal(rgb, img_ref)
    assert torch.equal(label.long(), torch.argmax(m[..., 3:], -1))
    print(f"[{name}] frame post-processing: exact")
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), map=_np(m), ref_rgb=_np(img_ref), ref_color=_np(col_ref),
                        ref_label=_np(label))


def loss_case(nerf, name, rows, seed):
    """Stage-I loss assembly (ref: nerf/nerf_helpers.py:14-62 modules, train_stage_rays_auto.py:455-468 assembly),
    with one class absent from the batch (count 0 -> 1) and autograd gradients w.r.t. both maps."""
    g = torch.Generator().manual_seed(seed)

    def make_map():
        m = torch.rand(rows, 15, generator=g)
        m[:, 3:] = torch.softmax(torch.randn(rows, 12, generator=g) * 2.0, -1)
        return m.requires_grad_(True)

    map_c, map_f = make_map(), make_map()
    target = torch.rand(rows, 3, generator=g)
    cls = torch.randint(0, 11, (rows,), generator=g)            # class 11 never drawn
    mask = torch.nn.functional.one_hot(cls, 12).float()
    mse_loss, ce_loss = nerf.MaskMSELoss(), nerf.MaskCrossEntropyLoss()
    c_l2, m_c_l2, w_c_l2 = mse_loss(mask, map_c[..., :3], target[..., :3])
    c_ce, m_c_ce, w_c_ce = ce_loss(mask, map_c[..., 3:], mask)
    coarse = c_l2 + 0.02 * c_ce + 0.005 * torch.sum(m_c_l2[7:9] + m_c_ce[7:9])
    f_l2, m_f_l2, w_f_l2 = mse_loss(mask, map_f[..., :3], target[..., :3])
    f_ce, m_f_ce, w_f_ce = ce_loss(mask, map_f[..., 3:], mask)
    fine = f_l2 + 0.02 * f_ce + 0.005 * torch.sum(m_f_l2[7:9] + m_f_ce[7:9])
    prob = (w_c_l2 + w_c_ce + w_f_l2 + w_f_ce) / (w_c_l2.sum() + w_c_ce.sum() + w_f_l2.sum() + w_f_ce.sum())
    loss = coarse + fine
    loss.backward()
    oc, of = map_c.detach().clone().requires_grad_(True), map_f.detach().clone().requires_grad_(True)
    loss_o, prob_o = O.stage1_loss(oc, of, target, mask)
    loss_o.backward()
    errs = {"loss": _maxabs(loss.detach(), loss_o.detach()), "prob": _maxabs(prob.detach(), prob_o),
            "d_coarse": _maxabs(map_c.grad, oc.grad), "d_fine": _maxabs(map_f.grad, of.grad)}
    print(f"[{name}] oracle vs reference:", errs)
    assert all(v == 0.0 for v in errs.values()), errs
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), map_c=_np(map_c), map_f=_np(map_f), target=_np(target),
                        mask=_np(mask), ref_loss=_np(loss), ref_prob=_np(prob), ref_d_coarse=_np(map_c.grad),
                        ref_d_fine=_np(map_f.grad),
                        ref_stats=np.concatenate([_np(x).reshape(-1) for x in
                                                  (loss, c_l2, c_ce, f_l2, f_ce, m_c_l2, m_c_ce, m_f_l2, m_f_ce)]))


def main():
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(os.cpu_count() or 1)
    nerf = RH.import_reference()
    if len(sys.argv) > 1 and sys.argv[1] == "loss":          # add this one case without regenerating the others
        loss_case(nerf, "stage1_loss", 777, 31)
        return
    helpers_case(nerf, "helpers")
    postprocess_case(nerf, "postprocess")
    sample_pdf_case(nerf, "sample_pdf_2048", 2048, 11)
    composite_case(nerf, "composite_bg", 256, 64, 21, True, False)
    composite_case(nerf, "composite_nobg_white", 100, 128, 22, False, True)
    field_case(nerf, "field_audio", "audio/person_2_auto", 512, 1)
    field_case(nerf, "field_expr2", "expression/person_2", 512, 2)
    e2e_case(nerf, "e2e_audio_val", "audio/person_2_auto", 16, 16, 0, 0.78, "validation")
    # white-noise weights + 15 octaves: the reference itself is ill-conditioned here (stress case, fine pass is checked
    # fed the reference's fine depths); the trained-like fixture below is held to the free-running bar
    e2e_case(nerf, "e2e_expr2_val", "expression/person_2", 12, 12, 1, 0.5, "validation", trained_like=False)
    e2e_case(nerf, "e2e_expr2_trained_val", "expression/person_2", 12, 12, 1, 0.5, "validation")
    e2e_case(nerf, "e2e_expr1_val", "expression/person_1", 12, 12, 3, 0.5, "validation")
    e2e_case(nerf, "e2e_audio_train_stoch", "audio/person_2_auto", 12, 12, 2, 0.78, "train", stochastic=True)
    e2e_case(nerf, "e2e_audio_nobg_val", "audio/person_2_auto", 8, 8, 5, 0.78, "validation", bg_mode="none")
    e2e_case(nerf, "e2e_audio_white_val", "audio/person_2_auto", 8, 8, 6, 0.78, "validation", bg_mode="white")
    loss_case(nerf, "stage1_loss", 777, 31)
    print("golden vectors written to", GOLD)


if __name__ == "__main__":
    main()
