"""Data-parallel sharding of the hot path over the GPUs of one box (one process per GPU).

The reference has no distributed code; the path shards because patches are
independent (per-patch zero padding, no halo): a rank owns a z-slab of a tiling (full-volume
segmentation, BASELINE configs[3]) or a contiguous slice of a patch list (latent extraction,
configs[2]).  There is no data-path collective; the only exchange steps are the all-reduce of the
PCA moments and the all-gather of the 2-D projections, done here on tensors of either backend
(NCCL for CUDA tensors on the GPU box, gloo for CPU tensors in the host-logic tests).
"""
import numpy as np
import torch
import torch.distributed as dist


def world(group=None):
    """(rank, world_size); (0, 1) when torch.distributed is not initialised."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def shard_range(n, rank, world_size):
    """Contiguous, balanced slice [a, b) of n work units for `rank` (sizes differ by at most 1)."""
    if not 0 <= rank < world_size:
        raise ValueError("rank %d outside world of %d" % (rank, world_size))
    base, rem = divmod(int(n), int(world_size))
    a = rank * base + min(rank, rem)
    return a, a + base + (1 if rank < rem else 0)


def slab_bounds(vol_shape, width, rank, world_size):
    """z-range [z0, z1) of the slab owned by `rank`: whole tile rows, balanced over the ranks."""
    nz_tiles = int(vol_shape[0]) // int(width)
    a, b = shard_range(nz_tiles, rank, world_size)
    return a * int(width), b * int(width)


def slab_of(patches, rank, world_size):
    """Sub-table (same class) of the patches whose corner lies in this rank's z-slab.

    Works for Grid (scalar width) and fixed-width Patches tilings; the union over the ranks is the
    whole table and the slabs are disjoint, so stitching the per-rank outputs needs no exchange."""
    width = patches.wd if hasattr(patches, "wd") else int(np.asarray(patches.widths)[0, 0])
    z0, z1 = slab_bounds(patches.vol_shape, width, rank, world_size)
    z = np.asarray(patches.points)[:, 0].astype(np.int64)
    idx = np.nonzero((z >= z0) & (z < z1))[0]
    return patches.select_by_indices(idx), (z0, z1)


def allreduce_sum_(t, group=None):
    """In-place SUM all-reduce of a tensor (moments of LatentPCA); no-op for a single process."""
    if world(group)[1] > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def allgather_rows(t, group=None):
    """Concatenates equally sized per-rank row blocks (n_local, k) in rank order."""
    ws = world(group)[1]
    if ws == 1:
        return t
    outs = [torch.empty_like(t) for _ in range(ws)]
    dist.all_gather(outs, t.contiguous(), group=group)
    return torch.cat(outs, dim=0)


def max_over_ranks(value, device=None, group=None):
    """MAX of a python float over the ranks (timings are reported as the slowest rank's)."""
    if world(group)[1] == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
