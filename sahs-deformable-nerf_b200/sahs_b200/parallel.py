"""One process per GPU; rays shard across ranks.

The path has no cross-ray operation, so rendering needs no data-path collective: a frame's rays are cut into
contiguous, tile-aligned ranges, every rank renders its own range with the same replicated weights, and the
per-ray outputs are gathered once per frame (render) -- or whole frames go to different ranks (clip rendering)
and nothing is exchanged at all.  Training is data parallel over rays/images: one all-reduce of the flat fp32
gradient buffer per step.  `torch.distributed` (NCCL on GPUs, gloo in the CPU tests) is plumbing only.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist

TILE = 128  # rows of one field-kernel tile; shards are multiples of it so no tile straddles two ranks


def shard_range(num_rays: int, rank: int, world: int, align: int = TILE) -> Tuple[int, int]:
    """[begin, end) of the rays rank `rank` owns: contiguous, `align`-aligned, sizes differing by <= align."""
    if world <= 1:
        return 0, num_rays
    tiles = (num_rays + align - 1) // align
    base, extra = divmod(tiles, world)
    t0 = rank * base + min(rank, extra)
    t1 = t0 + base + (1 if rank < extra else 0)
    return min(t0 * align, num_rays), min(t1 * align, num_rays)


def frame_owner(frame_index: int, world: int) -> int:
    """Clip rendering (BASELINE config 5): frame f is rendered entirely by rank f mod world."""
    return frame_index % world


def gather_rays(parts: Sequence[torch.Tensor], num_rays: int, group=None) -> List[torch.Tensor]:
    """All-gather per-ray outputs of a sharded render back into frame order.  `parts` are this rank's tensors
    ([n_local, ...]); returns tensors of leading size num_rays on every rank."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return list(parts)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = [shard_range(num_rays, r, world) for r in range(world)]
    out = []
    for t in parts:
        pad = max(e - b for b, e in sizes)
        buf = t.new_zeros((pad,) + tuple(t.shape[1:]))
        buf[: t.shape[0]] = t
        bufs = [torch.empty_like(buf) for _ in range(world)]
        dist.all_gather(bufs, buf, group=group)
        out.append(torch.cat([bufs[r][: e - b] for r, (b, e) in enumerate(sizes)], dim=0))
    return out


def render_sharded(render_fn, ro: torch.Tensor, rd: torch.Tensor, background=None, group=None):
    """Strong-scaling render of one frame: rank k renders rays shard_range(R, k, n) with `render_fn(ro, rd, bg)`
    (a closure over run_one_iter_of_nerf in train mode, which returns flat per-ray tensors) and the outputs are
    gathered.  Bitwise identical to the single-GPU render because rays are independent."""
    R = ro.shape[0]
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    b, e = shard_range(R, rank, world)
    outs = render_fn(ro[b:e], rd[b:e], background[b:e] if background is not None else None)
    return gather_rays(list(outs), R, group)


def render_clip(render_frame_fn, num_frames: int, group=None):
    """Clip rendering (BASELINE config 5): frame f is rendered entirely by rank `frame_owner(f, world)` with
    `render_frame_fn(f)`, which returns the frame as ONE uint8 tensor (e.g. `cat(frame_postprocess(rgb_map)[0:2])`:
    uint8 rgb + argmax label, 4 B/pixel instead of the 76 B/pixel fp32 maps).  No collective runs while frames are being
    rendered; at the end the frames are gathered once.  Returns the clip `[num_frames, ...]` in frame order on rank 0
    and None on the other ranks (single process: the clip)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    mine = [render_frame_fn(f) for f in range(num_frames) if frame_owner(f, world) == rank]
    if world == 1:
        return torch.stack(mine) if mine else None
    per_rank = (num_frames + world - 1) // world
    shape = [None]
    if mine:
        shape[0] = (tuple(mine[0].shape), mine[0].dtype)
    shapes = [None] * world
    dist.all_gather_object(shapes, shape[0], group=group)
    ref = next((sh for sh in shapes if sh is not None), None)
    if ref is None:
        return None
    fshape, dtype = ref
    if mine:
        dev = mine[0].device
    else:       # a rank without a frame (fewer frames than ranks) still takes part in the gather
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    local = torch.zeros((per_rank,) + fshape, dtype=dtype, device=dev)      # padded to the largest share
    for i, fr in enumerate(mine):
        local[i] = fr
    bufs = [torch.empty_like(local) for _ in range(world)]
    dist.all_gather(bufs, local, group=group)
    if rank != 0:
        return None
    return torch.stack([bufs[frame_owner(f, world)][f // world] for f in range(num_frames)])


def allreduce_gradients(params, group=None, average: bool = True) -> None:
    """Data-parallel training: one all-reduce of a flat fp32 gradient buffer per step."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, group=group)
    if average:
        flat /= dist.get_world_size(group)
    views, off = [], 0
    for g in grads:
        n = g.numel()
        views.append(flat[off:off + n].view_as(g))
        off += n
    if hasattr(torch, "_foreach_copy_"):
        torch._foreach_copy_(grads, views)          # one multi-tensor launch instead of one copy per parameter
    else:
        for g, v in zip(grads, views):
            g.copy_(v)
