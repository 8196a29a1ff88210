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


def local_slab(patches, rank, world_size):
    """The tiles of this rank's z-slab re-based to the slab: returns (sub-table whose vol_shape is the slab's and whose
    z corners start at 0, (z0, z1)).  The rank then only ever holds / uploads vol[z0:z1] (BASELINE configs[3]: "slab-
    partitioned ... with no halo, because patches are independent"); stitching the per-rank masks back is the
    concatenation of the slabs in rank order."""
    sub, (z0, z1) = slab_of(patches, rank, world_size)
    vs = tuple(int(v) for v in patches.vol_shape)
    pts = np.asarray(sub.points).astype(np.int64).copy()
    pts[:, 0] -= z0
    from .structures.grid import Grid
    from .structures.patches import Patches
    shape = (z1 - z0,) + vs[1:]
    if len(pts) == 0:
        return None, (z0, z1)
    if isinstance(patches, Grid):
        return Grid(shape, initialize_by="data", points=pts.astype(np.uint32), width=patches.wd), (z0, z1)
    return Patches(shape, "data", points=pts.astype(np.uint32), widths=np.asarray(sub.widths)), (z0, z1)


def z_sorted_shard(points, widths, vol_shape, rank, world_size):
    """Shard of an arbitrary patch list for latent extraction (BASELINE configs[2]): the list is ordered by corner z and
    cut into `world_size` equal runs, so rank r needs only the planes [z_lo, z_hi) its run touches (its slab plus a
    halo of width - 1 planes) instead of the whole volume.

    Returns (idx, local_points, (z_lo, z_hi)): idx = indices of this rank's patches in the ORIGINAL list (ascending z,
    stable), local_points = their corners with z re-based to z_lo.  Concatenating the per-rank results in rank order
    gives the rows in ``np.concatenate(idx_of_every_rank)`` order; ``unpermute`` undoes it."""
    points = np.asarray(points)
    widths = np.asarray(widths)
    order = np.argsort(points[:, 0].astype(np.int64), kind="stable")
    a, b = shard_range(len(order), rank, world_size)
    idx = order[a:b]
    if len(idx) == 0:
        return idx, points[:0].copy(), (0, 0)
    z = points[idx, 0].astype(np.int64)
    z_lo = int(z.min())
    z_hi = int((z + widths[idx, 0].astype(np.int64)).max())
    assert z_hi <= int(vol_shape[0])
    loc = points[idx].astype(np.int64)
    loc[:, 0] -= z_lo
    return idx, loc.astype(points.dtype), (z_lo, z_hi)


def z_sorted_order(points):
    """The row order in which ``z_sorted_shard`` results concatenate over the ranks (independent of world_size)."""
    return np.argsort(np.asarray(points)[:, 0].astype(np.int64), kind="stable")


def unpermute(rows_sorted, order):
    """rows_sorted[k] belongs to original row order[k] -> array in the original row order."""
    out = np.empty_like(rows_sorted)
    out[order] = rows_sorted
    return out


def allreduce_minmax(lo, hi, device=None, group=None):
    """(min over ranks of lo, max over ranks of hi): the global ``calc_voxel_min_max`` of a slab-partitioned volume."""
    if world(group)[1] == 1:
        return float(lo), float(hi)
    t = torch.tensor([-float(lo), float(hi)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return -float(t[0].item()), float(t[1].item())


class PeerReducer:
    """One-shot SUM all-reduce of small float64 CUDA vectors over NVLink peer memory (csrc/peer_comm.cu): every rank maps
    every peer's slot buffer through CUDA IPC; a reduction is one single-CTA kernel per rank, no NCCL launch.  Used for the
    SyncBN reductions of the training step (34 latency-bound all-reduces of <= 4 KB per step of M_a01).  All ranks must
    make the same sequence of calls on their current stream."""
    SLOT_DOUBLES = 2048
    NSLOTS = 64

    def __init__(self, group=None):
        """Every rank runs the same collectives whatever happens locally (a rank that cannot allocate or map must not leave
        the others waiting in a different collective): all-gather of (status, handle), then a MIN all-reduce of the import
        status; raises RuntimeError on EVERY rank if any rank failed."""
        import ctypes
        from . import _lib
        self._L = _lib.lib()
        self._h = None
        rank, ws = world(group)
        dev = torch.device("cuda", torch.cuda.current_device())
        h = ctypes.c_void_p()
        mine = (ctypes.c_ubyte * 64)()
        ok = self._L.te_peer_create(rank, ws, self.SLOT_DOUBLES, self.NSLOTS, ctypes.byref(h)) == 0
        ok = ok and self._L.te_peer_export(h, mine) == 0
        t = torch.tensor([1 if ok else 0] + list(mine), dtype=torch.uint8, device=dev)
        outs = [torch.empty_like(t) for _ in range(ws)]
        dist.all_gather(outs, t, group=group)
        table = torch.stack(outs).cpu().numpy()
        all_ok = bool(table[:, 0].all())
        if all_ok:
            all_ok = self._L.te_peer_import(h, table[:, 1:].tobytes()) == 0
        flag = torch.tensor([1 if all_ok else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        if int(flag.item()) == 0:
            if h.value:
                self._L.te_peer_destroy(h)
            raise RuntimeError("peer-memory buffers could not be created / mapped on every rank: %s"
                               % self._L.te_last_error().decode(errors="replace"))
        self._h = h

    def allreduce_(self, t):
        from . import _device, _lib
        _lib.check(self._L.te_peer_allreduce_f64(self._h, t.data_ptr(), t.numel(), _device.stream_ptr()), "te_peer_allreduce_f64")
        return t

    def __del__(self):
        try:
            if self._h is not None:
                self._L.te_peer_destroy(self._h)
        except Exception:  # noqa: BLE001
            pass


_PEER = {}


def peer_reducer(group=None):
    """The process's PeerReducer for `group` (created on first use), or None when the peer path does not apply: a single
    process, a backend other than NCCL (CPU / gloo rigs), several ranks sharing one GPU -- or when it was not asked for:
    the path is OPT-IN (TOMOENC_PEER_ALLREDUCE=1).  It is validated at 2 ranks (tests/test_zz_late_additions_gpu.py: 200 reductions
    against NCCL, CUDA-graph replay, a data-parallel fit step); a run at 8 ranks together with the bucketed gradient
    all-reduce on a side stream hung, so NCCL stays the default for every reduction."""
    import os
    key = id(group)
    if key in _PEER:
        return _PEER[key]
    red = None
    ws = world(group)[1]
    if (ws > 1 and os.environ.get("TOMOENC_PEER_ALLREDUCE", "0") == "1" and dist.get_backend(group) == "nccl"
            and torch.cuda.is_available() and ws <= 16):
        try:
            red = PeerReducer(group)                # raises on every rank or on none: all ranks take the same path
        except RuntimeError as e:                   # no IPC / no peer access: NCCL does the job
            import warnings
            warnings.warn("peer-memory all-reduce unavailable (%s): using NCCL for the small reductions" % (e,))
            red = None
    _PEER[key] = red
    return red


def allreduce_sum_(t, group=None):
    """In-place SUM all-reduce of a tensor (moments of LatentPCA, SyncBN sums, gradients); no-op for a single process.
    Small contiguous float64 CUDA vectors go through the peer-memory reducer when it is available."""
    if world(group)[1] > 1:
        if t.is_cuda and t.dtype == torch.float64 and t.is_contiguous() and 0 < t.numel() <= PeerReducer.SLOT_DOUBLES:
            red = peer_reducer(group)
            if red is not None:
                return red.allreduce_(t)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def allgather_rows(t, group=None):
    """Concatenates per-rank row blocks (n_local, k) in rank order.  Row counts may differ between ranks (a balanced
    shard of n rows differs by at most one row): the counts are gathered first and short blocks are padded for the
    collective, then trimmed."""
    ws = world(group)[1]
    if ws == 1:
        return t
    t = t.contiguous()
    if t.is_cuda and dist.get_backend(group) == "gloo":
        # gloo has no all_gather for CUDA tensors (single-GPU test rigs): stage through the host
        return allgather_rows(t.cpu(), group).to(t.device)
    counts = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
    all_counts = [torch.empty_like(counts) for _ in range(ws)]
    dist.all_gather(all_counts, counts, group=group)
    all_counts = [int(c.item()) for c in all_counts]
    m = max(all_counts)
    if t.shape[0] < m:
        t = torch.cat([t, t.new_zeros((m - t.shape[0],) + tuple(t.shape[1:]))], dim=0)
    outs = [torch.empty_like(t) for _ in range(ws)]
    dist.all_gather(outs, t, group=group)
    return torch.cat([o[:c] for o, c in zip(outs, all_counts)], dim=0)


def max_over_ranks(value, device=None, group=None):
    """MAX of a python float over the ranks (timings are reported as the slowest rank's)."""
    if world(group)[1] == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
