"""GPU, world_size 2: the sharded hot path against the single-rank result.

Two processes, one per GPU over NCCL when the box has two GPUs, otherwise both on GPU 0 over gloo (same product code:
sharding.py only calls torch.distributed).  Checks:
  * segmentation: every rank segments its z-slab of ONE volume (sharding.local_slab, host slab in -> host mask out);
    the concatenated slabs equal the single-rank mask bit for bit;
  * latent PCA: every rank encodes its run of the z-sorted patch list from slab + halo planes only, the moments are
    all-reduced, every rank solves the same eigenproblem: components equal the single-rank fit (||cos| - 1| < 1e-5) and the
    all-gathered, un-permuted projections equal the single-rank projections."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

SMALL = dict(n_filters=[16, 32], n_blocks=2, activation="lrelu", batch_norm=True, isconcat=[True] * 2, pool_size=[2, 2])
ENC = dict(n_filters=[16, 32], n_blocks=2, activation="lrelu", batch_norm=True, pool_size=[2, 2], hidden_units=[64, 16],
           POOL_FLAG=True)
VS = (128, 96, 96)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _setup(seed_u, seed_e):
    from oracle import nets_torch as O, phantom, weights as OW
    from tomoencoders_b200.neural_nets import SelfSupervisedCAE, SurfaceSegmenter
    vol, _ = phantom.volume_f16(VS)
    fe = SurfaceSegmenter(model_initialization="define-new", descriptor_tag="mr", mode="fp16", **SMALL)
    fe.models["segmenter"].set_weights(OW.draw(O.unet_spec(**SMALL), seed_u))
    ae = SelfSupervisedCAE(model_initialization="define-new", model_size=(32,) * 3, descriptor_tag="mr", mode="fp16", **ENC)
    ae.models["encoder"].set_weights(OW.draw(O.encoder_spec((32,) * 3, **ENC)[0], seed_e))
    rng = np.random.default_rng(5)
    pts = np.stack([rng.integers(0, VS[a] - 32, 301) for a in range(3)], axis=1).astype(np.uint32)
    return vol, fe, ae, pts


def _worker(rank, world_size, port, out_dir):
    ngpu = torch.cuda.device_count()
    torch.cuda.set_device(rank % ngpu)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    backend = "nccl" if ngpu >= world_size else "gloo"
    dist.init_process_group(backend, rank=rank, world_size=world_size)
    try:
        from tomoencoders_b200 import Grid, Patches, sharding
        from tomoencoders_b200.neural_nets import LatentPCA
        vol, fe, ae, pts = _setup(3, 4)
        dev = torch.device("cuda", rank % ngpu)
        # ---- segmentation of this rank's slab
        grid = Grid(VS, width=32)
        loc, (z0, z1) = sharding.local_slab(grid, rank, world_size)
        slab = np.ascontiguousarray(vol[z0:z1])
        lo, hi = fe.calc_voxel_min_max(slab, 4)
        mm = sharding.allreduce_minmax(lo, hi, device=dev)
        assert mm == (float(vol[::4, ::4, ::4].min()), float(vol[::4, ::4, ::4].max()))
        mask = fe.segment_volume(slab, loc, 5, min_max=mm)
        np.save(os.path.join(out_dir, "mask_%d.npy" % rank), mask)
        # ---- latents of this rank's run of the z-sorted list, from slab + halo planes only
        wds = np.full_like(pts, 32)
        idx, loc_pts, (zl, zh) = sharding.z_sorted_shard(pts, wds, VS, rank, world_size)
        dvol = torch.from_numpy(np.ascontiguousarray(vol[zl:zh])).to(dev)
        p = Patches(tuple(dvol.shape), "data", points=loc_pts, widths=np.full_like(loc_pts, 32))
        lat = ae.embed_volume(dvol, p, 16, min_max=mm)
        pca = LatentPCA(2)
        pca.partial_fit(lat)
        pca.finalize()
        z = pca.transform_all_gather(lat).cpu().numpy()
        z = sharding.unpermute(z, sharding.z_sorted_order(pts))
        if rank == 0:
            np.save(os.path.join(out_dir, "comps.npy"), pca.components_)
            np.save(os.path.join(out_dir, "z.npy"), z)
            np.save(os.path.join(out_dir, "mm.npy"), np.asarray(mm))
        else:
            np.save(os.path.join(out_dir, "comps_1.npy"), pca.components_)
    finally:
        dist.destroy_process_group()


def test_two_ranks_reproduce_the_single_rank_masks_and_pca(tmp_path):
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    from tomoencoders_b200 import Grid, Patches
    from tomoencoders_b200.neural_nets import LatentPCA
    vol, fe, ae, pts = _setup(3, 4)
    mm = tuple(float(v) for v in np.load(tmp_path / "mm.npy"))
    want = fe.segment_volume(vol, Grid(VS, width=32), 5, min_max=mm)
    got = np.concatenate([np.load(tmp_path / "mask_0.npy"), np.load(tmp_path / "mask_1.npy")], axis=0)
    assert 0.02 < want.mean() < 0.98
    np.testing.assert_array_equal(got, want)
    p = Patches(VS, "data", points=pts, widths=np.full_like(pts, 32))
    lat = ae.embed_volume(vol, p, 16, min_max=mm)
    pca = LatentPCA(2).fit(lat)
    c2 = np.load(tmp_path / "comps.npy")
    np.testing.assert_array_equal(c2, np.load(tmp_path / "comps_1.npy"))        # every rank holds the same solution
    assert np.abs(np.abs(np.sum(c2 * pca.components_, axis=1)) - 1.0).max() < 1e-5
    z2 = np.load(tmp_path / "z.npy")
    assert z2.shape == (301, 2)
    np.testing.assert_allclose(z2, pca.transform(lat), atol=2e-3 * np.abs(z2).max())

