"""Multi-GPU tests (-m gpu, skipped with fewer than 2 devices; run with `gpurun --gpus 2 -- python -m pytest
tests/test_gpu_dist.py -m gpu`): one process per GPU over NCCL.

  * a frame rendered in tile-aligned ray ranges on N ranks and all-gathered (parallel.render_sharded) is BITWISE equal
    to the single-GPU frame (rays are independent end to end);
  * data-parallel training: the all-reduced, averaged gradients of N ray shards equal the gradients of the concatenated
    batch (mean loss over equal shards), and FlatAdam's built-in all-reduce applies the same update on every rank."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import sahs_fixtures as FX

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _setup(rank, world, port):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    return dev


def _model_and_frame(dev, H, W, cfg_name="audio/person_2_auto"):
    import sahs_b200
    from oracle import sahs_oracle as O
    cfg = FX.load_cfg(cfg_name)
    cfg.nerf.train.perturb, cfg.nerf.train.radiance_field_noise_std = False, 0.0
    spec = O.spec_from_cfg(cfg)
    sd = FX.make_state_dict(spec, seed=42, dense=True, trained_like=True)
    model = getattr(sahs_b200.models, cfg.models.mask.type)(cfg)
    model.load_state_dict(sd)
    model = model.to(dev)
    fr = FX.make_frame_inputs(spec, H, W, seed=3)
    pose = fr["pose"].to(dev)
    with torch.no_grad():
        ro, rd = sahs_b200.get_ray_bundle(H, W, fr["intrinsics"], pose)
    return sahs_b200, cfg, model, fr, pose, ro.reshape(-1, 3), rd.reshape(-1, 3)


def _render_worker(rank, world, port):
    from sahs_b200 import parallel as PL
    dev = _setup(rank, world, port)
    try:
        H, W = 24, 40                                  # 960 rays: 7.5 tiles -> uneven, tile-aligned shards
        sahs, cfg, model, fr, pose, ro, rd = _model_and_frame(dev, H, W)
        bg = fr["background"].view(-1, 15).to(dev)
        drv = fr["driving"].to(dev)

        def render(ro_p, rd_p, bg_p):
            with torch.no_grad():
                return sahs.run_one_iter_of_nerf(H, W, 1.0, model, ro_p, rd_p, cfg, mode="train", driving=drv, pose=pose,
                                                 background_prior=bg_p)

        got = PL.render_sharded(render, ro, rd, bg)
        whole = render(ro, rd, bg)
        b, e = PL.shard_range(ro.shape[0], rank, world)
        assert b % 128 == 0 and 0 < e - b < ro.shape[0]
        for a, w_ in zip(got, whole):
            assert a.shape == w_.shape and torch.equal(a, w_)      # bitwise
    finally:
        dist.destroy_process_group()


def _train_worker(rank, world, port):
    dev = _setup(rank, world, port)
    try:
        H, W = 16, 32                                  # 512 rays, 256 per rank
        sahs, cfg, model, fr, pose, ro, rd = _model_and_frame(dev, H, W)
        bg = fr["background"].view(-1, 15).to(dev)
        drv = fr["driving"].to(dev)
        R = ro.shape[0]
        target = torch.rand(R, 15, generator=torch.Generator().manual_seed(9)).to(dev)   # colour + semantic channels

        def loss_of(sl):
            out = sahs.run_one_iter_of_nerf(H, W, 1.0, model, ro[sl], rd[sl], cfg, mode="train", driving=drv, pose=pose,
                                            background_prior=bg[sl])
            return ((out[3] - target[sl]) ** 2).mean() + ((out[0] - target[sl]) ** 2).mean()

        # reference: the whole batch on this rank
        model.zero_grad(set_to_none=True)
        loss_of(slice(0, R)).backward()
        full = {n: p.grad.detach().clone() for n, p in model.named_parameters()}
        # data parallel: my half, then one all-reduce of the flat gradient buffer, averaged
        model.zero_grad(set_to_none=True)
        half = R // world
        loss_of(slice(rank * half, (rank + 1) * half)).backward()
        from sahs_b200 import parallel as PL
        PL.allreduce_gradients(list(model.parameters()), average=True)
        torch.cuda.synchronize()
        worst = 1.0
        for n, p in model.named_parameters():
            g, r = p.grad.double().reshape(-1), full[n].double().reshape(-1)
            cos = float(torch.dot(g, r) / (g.norm() * r.norm() + 1e-30))
            ratio = float(g.norm() / (r.norm() + 1e-30))
            worst = min(worst, cos)
            # same fp16 kernels on both sides; differences: atomics order, per-shard gradient scale (16 / max|d_raw|)
            assert cos >= 0.9999 and abs(ratio - 1) <= 2e-3, (n, cos, ratio)
        # every rank holds the same averaged gradients
        flat = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
        gathered = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(gathered, flat)
        assert all(torch.equal(gathered[0], t) for t in gathered[1:])
        # FlatAdam averages the gradients itself: parameters stay identical across ranks after a step
        opt = sahs.FlatAdam(model.parameters(), lr=1e-3)
        model.zero_grad(set_to_none=True)
        loss_of(slice(rank * half, (rank + 1) * half)).backward()
        opt.step()
        flat = opt.flat_param.detach().clone()
        gathered = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(gathered, flat)
        assert all(torch.equal(gathered[0], t) for t in gathered[1:])
    finally:
        dist.destroy_process_group()


def _need_two():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 CUDA devices (gpurun --gpus 2)")


def test_sharded_render_is_bitwise_equal_to_single_gpu():
    _need_two()
    mp.spawn(_render_worker, args=(2, _free_port()), nprocs=2, join=True)


def test_data_parallel_gradients_equal_the_concatenated_batch():
    _need_two()
    mp.spawn(_train_worker, args=(2, _free_port()), nprocs=2, join=True)
