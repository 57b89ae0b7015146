"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: ray sharding, output gather, gradient all-reduce."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import sahs_fixtures  # noqa: F401  (sets sys.path)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, R):
    from sahs_b200 import parallel as PL
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        ro, rd = torch.randn(R, 3, generator=g), torch.randn(R, 3, generator=g)
        bg = torch.rand(R, 15, generator=g)

        def fake_render(ro_, rd_, bg_):      # stands in for the CUDA pipeline: any per-ray function
            return (ro_ * 2 + rd_).sum(-1, keepdim=True) * bg_, (ro_ - rd_).norm(dim=-1)

        full = fake_render(ro, rd, bg)
        got = PL.render_sharded(fake_render, ro, rd, bg)
        for a, b in zip(got, full):
            assert torch.equal(a, b)
        b0, e0 = PL.shard_range(R, rank, world)
        assert (b0 % PL.TILE == 0 or b0 == R) and (e0 % PL.TILE == 0 or e0 == R)
        # gradient all-reduce == gradient of the concatenated batch
        w = torch.nn.Parameter(torch.ones(5, 3))
        x = torch.arange(15, dtype=torch.float32).view(5, 3) * (rank + 1)
        (w * x).sum().backward()
        PL.allreduce_gradients([w], average=False)
        want = sum(torch.arange(15, dtype=torch.float32).view(5, 3) * (r + 1) for r in range(world))
        assert torch.equal(w.grad, want)
    finally:
        dist.destroy_process_group()


def _clip_worker(rank, world, port, num_frames):
    from sahs_b200 import parallel as PL
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rendered = []

        def fake_frame(f):                   # stands in for render + frame_postprocess: uint8 [pixels, 4]
            rendered.append(f)
            return (torch.arange(24, dtype=torch.int32).view(6, 4) * (f + 1) % 251).to(torch.uint8)

        clip = PL.render_clip(fake_frame, num_frames)
        assert rendered == [f for f in range(num_frames) if f % world == rank]       # whole frames, round robin
        if rank == 0:
            assert clip.shape == (num_frames, 6, 4) and clip.dtype == torch.uint8
            for f in range(num_frames):
                assert torch.equal(clip[f], (torch.arange(24, dtype=torch.int32).view(6, 4) * (f + 1) % 251).to(torch.uint8))
        else:
            assert clip is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("num_frames", [5, 4, 1])
def test_clip_frames_round_robin_and_gather_world2(num_frames):
    mp.spawn(_clip_worker, args=(2, _free_port(), num_frames), nprocs=2, join=True)


@pytest.mark.parametrize("R", [1000, 128, 5])
def test_sharded_render_and_allreduce_world2(R):
    mp.spawn(_worker, args=(2, _free_port(), R), nprocs=2, join=True)


def test_shard_ranges_cover_exactly():
    from sahs_b200 import parallel as PL
    for R in (0, 1, 127, 128, 129, 262144, 2048, 1000):
        for world in (1, 2, 4, 8):
            spans = [PL.shard_range(R, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == R
            for (b0, e0), (b1, e1) in zip(spans, spans[1:]):
                assert e0 == b1 and b0 <= e0
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= PL.TILE
    assert [PL.frame_owner(f, 8) for f in range(10)] == [0, 1, 2, 3, 4, 5, 6, 7, 0, 1]


def _stage2_clip_worker(rank, world, port, num_frames):
    """Stage II over ranks: frame f is refined entirely by rank f mod world (same sharding as the Stage-I clip), with the
    identity photo's part computed once per rank; the product's host logic runs on the CPU through tests/spade_emulator."""
    import spade_emulator as EM
    import spade_fixtures as SF
    from sahs_b200 import parallel as PL
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(2)
        m = EM.EmulatedGenerator()
        m.load_state_dict(SF.make_state_dict("generator", seed=0), strict=True)
        src = SF.make_inputs(16, 16, seed=0)["i_src"]
        ident = m.encode_identity(src)

        def refine_frame(f):                  # uint8 frame like eval_get_texture_photo_audio.py:187-193 (clip, * 255)
            out = m.refine(ident, SF.make_inputs(16, 16, seed=10 + f)["i_raw"])
            return (out[0].permute(1, 2, 0).clamp(0, 1) * 255).to(torch.uint8)

        clip = PL.render_clip(refine_frame, num_frames)
        if rank == 0:
            assert clip.shape == (num_frames, 16, 16, 3)
            for f in range(num_frames):       # = the full forward of that frame, whichever rank refined it
                want = m(src, SF.make_inputs(16, 16, seed=10 + f)["i_raw"])
                assert torch.equal(clip[f], (want[0].permute(1, 2, 0).clamp(0, 1) * 255).to(torch.uint8))
        else:
            assert clip is None
    finally:
        dist.destroy_process_group()


def test_stage2_clip_refinement_over_ranks_world2():
    mp.spawn(_stage2_clip_worker, args=(2, _free_port(), 3), nprocs=2, join=True)
