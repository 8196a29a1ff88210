"""CPU, world_size 2, gloo: the host logic of the multi-GPU path (tomoencoders_b200/sharding.py) --
slab / list partitioning, the PCA-moment all-reduce and the projection all-gather -- checked against
the single-process oracle.  No CUDA calls: the moments of each rank's shard are computed with numpy
exactly as te_pca_accumulate defines them ([count, sum h, sum h h^T] in float64)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import pca_np
from tomoencoders_b200 import Grid, Patches, sharding


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _moments(h):
    h = h.astype(np.float64)
    return np.concatenate([[len(h)], h.sum(0), (h.T @ h).reshape(-1)])


def _worker(rank, world_size, port, h_all, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        assert sharding.world() == (rank, world_size)
        a, b = sharding.shard_range(len(h_all), rank, world_size)
        h = h_all[a:b]
        m = sharding.allreduce_sum_(torch.from_numpy(_moments(h)))
        np.testing.assert_allclose(m.numpy(), _moments(h_all), rtol=1e-12)
        # every rank solves the same eigenproblem from the reduced moments (te_pca_solve on the GPU box)
        n, d = int(m[0]), h_all.shape[1]
        mean = m[1:1 + d].numpy() / n
        cov = (m[1 + d:].numpy().reshape(d, d) - n * np.outer(mean, mean)) / (n - 1)
        w, v = np.linalg.eigh(cov)
        comps = v[:, ::-1][:, :2].T
        comps *= np.sign(comps[np.arange(2), np.abs(comps).argmax(1)])[:, None]
        z_local = torch.from_numpy(((h - mean) @ comps.T).astype(np.float32))
        # equal shard sizes are required by allgather_rows; pad the short shard like the bench never needs to
        assert len(h_all) % world_size == 0
        z_all = sharding.allgather_rows(z_local)
        assert z_all.shape == (len(h_all), 2)
        t = sharding.max_over_ranks(10.0 + rank)
        assert t == 10.0 + world_size - 1
        # the peer-memory reducer is opt-in AND needs NCCL + CUDA: on a gloo / CPU rig it never engages, whatever the switch
        # says, and small float64 vectors (the SyncBN sums it would take) go through the backend's all-reduce
        os.environ["TOMOENC_PEER_ALLREDUCE"] = "1"
        assert sharding.peer_reducer() is None
        small = sharding.allreduce_sum_(torch.full((8,), float(rank + 1), dtype=torch.float64))
        assert torch.equal(small, torch.full((8,), 3.0, dtype=torch.float64))
        if rank == 0:
            np.save(os.path.join(out_dir, "z.npy"), z_all.numpy())
            np.save(os.path.join(out_dir, "comps.npy"), comps)
    finally:
        dist.destroy_process_group()


def test_pca_moments_allreduce_and_allgather_world2(tmp_path):
    g = np.random.default_rng(0)
    d = 16
    basis = np.linalg.qr(g.standard_normal((d, d)))[0]
    h_all = ((g.standard_normal((600, d)) * np.array([9, 4] + [0.3] * (d - 2))) @ basis.T + 2.0).astype(np.float32)
    mp.spawn(_worker, args=(2, _free_port(), h_all, str(tmp_path)), nprocs=2, join=True)
    _, comps, _, z = pca_np.exact_pca(h_all, 2)
    got_c, got_z = np.load(tmp_path / "comps.npy"), np.load(tmp_path / "z.npy")
    assert np.abs(np.abs(got_c @ comps.T) - np.eye(2)).max() < 1e-8
    np.testing.assert_allclose(got_z, z, atol=2e-4)


def test_shard_range_is_a_balanced_partition():
    for n in (0, 1, 7, 512, 1000003):
        for ws in (1, 2, 3, 8):
            spans = [sharding.shard_range(n, r, ws) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(ws - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_slabs_partition_a_tiling_without_overlap():
    g = Grid((256, 128, 192), width=64)
    p = Patches((256, 128, 192), "regular-grid", patch_size=(64,) * 3)
    for table in (g, p):
        for ws in (1, 2, 4):
            seen = []
            for r in range(ws):
                sub, (z0, z1) = sharding.slab_of(table, r, ws)
                assert type(sub) is type(table)
                assert z0 % 64 == 0 and z1 % 64 == 0
                assert np.all((sub.points[:, 0] >= z0) & (sub.points[:, 0] < z1))
                seen.append(sub.points)
            allp = np.concatenate(seen)
            assert len(allp) == len(table)
            np.testing.assert_array_equal(np.unique(allp, axis=0), np.unique(np.asarray(table.points), axis=0))


def test_local_slab_rebases_tiles_and_z_sorted_shard_covers_the_list():
    g = Grid((256, 128, 192), width=64)
    p = Patches((256, 128, 192), "regular-grid", patch_size=(64,) * 3)
    for table in (g, p):
        for ws in (1, 2, 4):
            total = 0
            for r in range(ws):
                loc, (z0, z1) = sharding.local_slab(table, r, ws)
                assert type(loc) is type(table) and loc.vol_shape == (z1 - z0, 128, 192)
                sub, _ = sharding.slab_of(table, r, ws)
                back = np.asarray(loc.points).astype(np.int64)
                back[:, 0] += z0
                np.testing.assert_array_equal(back, np.asarray(sub.points).astype(np.int64))
                total += len(loc)
            assert total == len(table)
    # more ranks than tile rows: the surplus ranks own nothing
    loc, (z0, z1) = sharding.local_slab(Grid((64, 64, 64), width=64), 1, 2)
    assert loc is None and z0 == z1
    # arbitrary patch list: equal runs in z order, each touching only [z_lo, z_hi)
    rng = np.random.default_rng(3)
    vs = (300, 100, 100)
    pts = np.stack([rng.integers(0, vs[a] - 32, 1001) for a in range(3)], axis=1).astype(np.uint32)
    wds = np.full_like(pts, 32)
    order = sharding.z_sorted_order(pts)
    for ws in (1, 3, 8):
        got = []
        for r in range(ws):
            idx, loc, (z_lo, z_hi) = sharding.z_sorted_shard(pts, wds, vs, r, ws)
            assert abs(len(idx) - 1001 / ws) < 1
            assert loc[:, 0].min() == 0 and int(loc[:, 0].max()) + 32 == z_hi - z_lo
            np.testing.assert_array_equal(loc[:, 1:], pts[idx, 1:])
            np.testing.assert_array_equal(loc[:, 0].astype(np.int64) + z_lo, pts[idx, 0].astype(np.int64))
            got.append(idx)
        np.testing.assert_array_equal(np.concatenate(got), order)
    rows = rng.standard_normal((1001, 2)).astype(np.float32)
    np.testing.assert_array_equal(sharding.unpermute(rows[order], order), rows)


def _worker_ragged(rank, world_size, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        a, b = sharding.shard_range(7, rank, world_size)              # 4 + 3 rows
        z = torch.arange(a, b, dtype=torch.float32)[:, None].repeat(1, 2)
        allz = sharding.allgather_rows(z)
        assert allz.shape == (7, 2) and torch.equal(allz[:, 0], torch.arange(7, dtype=torch.float32))
        lo, hi = sharding.allreduce_minmax(-1.0 - rank, 5.0 + rank)
        assert (lo, hi) == (-1.0 - (world_size - 1), 5.0 + world_size - 1)
    finally:
        dist.destroy_process_group()


def test_ragged_allgather_and_minmax_allreduce_world2(tmp_path):
    mp.spawn(_worker_ragged, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
