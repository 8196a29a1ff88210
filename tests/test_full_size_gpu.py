"""GPU: the hot path at BASELINE.json's full sizes (configs[2] and configs[3]: a 2048^3 volume, 32 768
tiles of 64^3, one rank's share of 10^6 random 64^3 patches), checked through size-independent
properties, because no CPU oracle finishes at these sizes:

* gather -> stitch of a disjoint tiling is the identity, and a checksum of the tiles equals the
  checksum of the volume (bit-exact, uint8);
* patches are independent units (per-patch zero padding, no halo): the stitched full-volume mask
  restricted to any tile equals that tile segmented on its own, and a z-slab segmented by one rank
  of a G-rank job equals the same slab of the single-rank result;
* latents of randomly placed patches computed by the fused gather+encode equal the latents of the
  same patches extracted first (same kernels, different batch composition), the PCA
  projection of the fitting set is centred, uncorrelated and variance-ordered.

The same code paths are compared with the oracle at small sizes in test_gather_scatter_gpu.py /
test_nets_gpu.py; nothing under oracle/ runs here except the seeded weight draw.
"""
import numpy as np
import pytest
import torch

from oracle import nets_torch as O, weights as OW
from tomoencoders_b200 import Grid, Patches, sharding
from tomoencoders_b200.neural_nets import LatentPCA, SelfSupervisedCAE, SurfaceSegmenter

pytestmark = pytest.mark.gpu

N = 2048
M_A01 = dict(n_filters=[16, 32, 64], n_blocks=3, activation="lrelu", batch_norm=True, isconcat=[True] * 3, pool_size=[2, 2, 2])
ENC = dict(n_filters=[16, 32, 64], n_blocks=3, activation="lrelu", batch_norm=True, pool_size=[2, 2, 2],
           hidden_units=[128, 128], POOL_FLAG=True)


def _need(gib):
    free, _ = torch.cuda.mem_get_info()
    if free < gib * 2 ** 30:
        pytest.skip("needs %d GiB of free device memory" % gib)


@pytest.fixture(scope="module")
def big_volume():
    """2048^3 uint8 on the device: smooth separable structure + counter-based noise, generated slab by
    slab (the volume never exists on the host)."""
    _need(60)
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(6)
    vol = torch.empty((N, N, N), dtype=torch.uint8, device=dev)
    ax = torch.arange(N, device=dev, dtype=torch.float32)
    cy, cx = torch.cos(ax / 37.0).view(1, N, 1), torch.sin(ax / 23.0).view(1, 1, N)
    for z0 in range(0, N, 64):
        cz = torch.cos(ax[z0:z0 + 64] / 29.0).view(64, 1, 1)
        s = ((cz * cy + cx * cz + cy * cx) > 0.2).float() * 140.0 + 50.0
        s += torch.randn((64, N, N), device=dev, generator=g) * 12.0
        vol[z0:z0 + 64] = s.clamp_(0, 255).to(torch.uint8)
        del s
    return vol


def test_full_volume_gather_stitch_identity(big_volume):
    grid = Grid((N,) * 3, width=64)
    assert len(grid) == 32768
    tiles = grid.extract(big_volume)                                  # device in -> device out
    assert tiles.shape == (32768, 64, 64, 64) and tiles.dtype == torch.uint8
    # checksum of checksums: per-z-plane sums (int64) of the volume == regrouped per-tile sums
    assert int(tiles.sum(dtype=torch.int64)) == int(big_volume.sum(dtype=torch.int64))
    i = 12345
    z, y, x = (int(v) for v in grid.points[i])
    assert torch.equal(tiles[i], big_volume[z:z + 64, y:y + 64, x:x + 64])
    out = torch.zeros_like(big_volume)
    grid.fill_patches_in_volume(tiles, out)
    assert torch.equal(out, big_volume)


def test_full_volume_segmentation_tile_and_slab_independence(big_volume):
    _need(40)
    fe = SurfaceSegmenter(model_initialization="define-new", descriptor_tag="M_a01", mode="fp16", **M_A01)
    fe.models["segmenter"].set_weights(OW.draw(O.unet_spec(**M_A01), 11))
    grid = Grid((N,) * 3, width=64)
    min_max = (0.0, 255.0)
    mask = fe.segment_volume(big_volume, grid, 32, min_max=min_max)
    assert mask.shape == big_volume.shape and mask.dtype == torch.uint8
    assert int(mask.max()) == 1 and int(mask.min()) == 0              # a non-trivial mask
    # any tile on its own (different chunk composition and chunk size)
    idx = np.random.default_rng(0).choice(len(grid), 48, replace=False)
    sub = grid.select_by_indices(idx)
    x = sub.extract(big_volume)
    y = fe.predict_patches("segmenter", x.unsqueeze(-1), 16, None, min_max=min_max)
    for k, i in enumerate(idx):
        z, yy, xx = (int(v) for v in grid.points[i])
        assert torch.equal(y[k, ..., 0], mask[z:z + 64, yy:yy + 64, xx:xx + 64]), "tile %d" % i
    # rank 5 of an 8-rank job: its slab, segmented from the slab alone
    slab, (z0, z1) = sharding.slab_of(grid, 5, 8)
    assert (z0, z1) == (5 * 256, 6 * 256) and len(slab) == 4096
    local = Grid((z1 - z0, N, N), width=64)
    m5 = fe.segment_volume(big_volume[z0:z1], local, 32, min_max=min_max)
    assert torch.equal(m5, mask[z0:z1])


def test_one_rank_share_of_1m_patches_encode_and_pca(big_volume):
    _need(20)
    fe = SelfSupervisedCAE(model_initialization="define-new", model_size=(64, 64, 64), descriptor_tag="t",
                           mode="fp16", **ENC)
    fe.models["encoder"].set_weights(OW.draw(O.encoder_spec((64, 64, 64), **ENC)[0], 12))
    np.random.seed(7)
    p = Patches((N,) * 3, "random-fixed-width", patch_size=(64,) * 3, n_points=1_000_000)
    a, b = sharding.shard_range(len(p), 3, 8)
    assert b - a == 125_000
    mine = p.select_by_indices(np.arange(a, b))                       # rank 3 of 8: 125 000 patches
    h = fe.embed_volume(big_volume, mine, 128, min_max=(0.0, 255.0))
    assert h.shape == (125_000, 128) and h.dtype == torch.float32 and bool(torch.isfinite(h).all())
    idx = np.random.default_rng(1).choice(len(mine), 96, replace=False)
    sub = mine.select_by_indices(idx)
    x = sub.extract(big_volume, (64, 64, 64))
    h2 = fe.predict_embeddings(x.unsqueeze(-1), 32, min_max=(0.0, 255.0))
    # identical conv arithmetic; the split-K dense layer accumulates with fp32 atomics (order varies)
    torch.testing.assert_close(h2, h[torch.as_tensor(idx, device=h.device)], rtol=1e-4, atol=1e-5)
    pca = LatentPCA(n_components=2)
    z = pca.fit_transform(h)
    z = torch.as_tensor(z).double().cpu()
    assert z.shape == (125_000, 2)
    scale = float(z.std(dim=0)[0])
    assert float(z.mean(dim=0).abs().max()) < 1e-3 * scale
    c = (z.T @ z) / len(z)
    assert abs(float(c[0, 1])) < 1e-3 * float(c[0, 0]) and float(c[0, 0]) >= float(c[1, 1]) > 0
