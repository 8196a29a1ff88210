#!/usr/bin/env python
"""Benchmark of the TomoEncoders hot path on B200 (contract: see the task statement / DESIGN.md).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload at N=1 = BASELINE.json configs[1]: 3-D U-net segmenter (M_a01: n_filters [16,32,64], pools
[2,2,2], skip concatenation on) over a synthetic 512^3 float16 volume tiled into 512 patches of 64^3,
chunked 32 patches at a time: fused gather+clip+rescale -> U-net (tcgen05) -> labels -> stitch.
One "step" = one full pass over the volume.  With N ranks (torchrun) every rank segments its own
512^3 slab (patches are independent: no data-path collective; weak scaling) -- that is `value`.
The same JSON line carries, at every N, the two multi-GPU workloads BASELINE.json names at their stated sizes:
  "cfg4": ONE 2048^3 fp16 volume, z-slab partitioned over the N ranks (sharding.local_slab), host -> host through
          calc_voxel_min_max (+ a MIN/MAX all-reduce of two scalars) and SurfaceSegmenter.segment_volume: STRONG scaling;
  "cfg3": 1 M random 64^3 patches of that volume encoded to 128-d latents (rank r takes the r-th run of the z-sorted list
          and uploads only the planes it touches = slab + 63-plane halo), all-reduce of the 16 513 float64 PCA moments,
          redundant eigensolve, projection, all-gather of the (n, 2) rows, host coordinates in -> host rows out.
Rank 0 prints ONE JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

# stdout must carry the JSON line only, but NCCL writes "NCCL version ..." to file descriptor 1 from C: keep a private
# duplicate of the real stdout for the JSON line and point descriptor 1 at stderr for everything else
_JSON_OUT = os.fdopen(os.dup(1), "w")
sys.stdout.flush()
os.dup2(2, 1)


def emit(line):
    _JSON_OUT.write(json.dumps(line) + "\n")
    _JSON_OUT.flush()


ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

M_A01 = dict(n_filters=[16, 32, 64], n_blocks=3, activation="lrelu", batch_norm=True, isconcat=[True] * 3,
             pool_size=[2, 2, 2])
ENC = dict(n_filters=[16, 32, 64], n_blocks=3, activation="lrelu", batch_norm=True, pool_size=[2, 2, 2],
           hidden_units=[128, 128], POOL_FLAG=True)
VOL = (512, 512, 512)
PATCH = 64
CHUNK = 32
ENC_PATCHES = 16384         # 64^3 patches encoded per rank per step in the "encode" leg, ONE PCA fit per step (configs[2]: 125 000 patches per rank and fit)
BIG = (2048, 2048, 2048)    # configs[2] / configs[3] volume
BIG_PATCHES = 1_000_000     # configs[2]
BIG_BLOCK = 128             # z planes per generated block of the synthetic 2048^3 volume (block b is seeded by b: any rank regenerates any plane)
CPU_TILES = 64              # bounded CPU sample of the same workload: two FULL chunks (a partial chunk is padded to 32 by the reference)
ENC_CHUNK = 128             # patches per encoder launch (the encoder keeps only pooled tensors: 128 patches = 1.3 GB)
METRIC = "voxels/sec segmented (3-D U-net M_a01, 64^3 tiles of a 512^3 fp16 volume per GPU)"


def synth_volume_torch(shape, seed, device):
    """Smooth blobs + noise in [0,1], float16, generated on the device (structure does not affect
    the arithmetic; BASELINE configs only fix shape and dtype)."""
    import torch
    g = torch.Generator(device=device).manual_seed(seed)
    coarse = torch.randn((1, 1) + tuple(s // 16 for s in shape), generator=g, device=device)
    up = torch.nn.functional.interpolate(coarse, size=shape, mode="trilinear", align_corners=False)[0, 0]
    y = (up > up.median()).float()
    del up
    x = y + 0.1 * torch.randn(shape, generator=g, device=device)
    x = (x - x.min()) / (x.max() - x.min())
    return x.to(torch.float16).contiguous()


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons for one GPU while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self._stop = threading.Event()
        self._t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 7:
                    self.rows.append(parts)
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm = sorted(float(r[0]) for r in self.rows)
        reasons = []
        for name, col in (("hw_slowdown", 3), ("hw_thermal_slowdown", 4), ("sw_thermal_slowdown", 5), ("sw_power_cap", 6)):
            if any(r[col].lower().startswith("active") for r in self.rows):
                reasons.append(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]),
                "power_w_max": max(float(r[2]) for r in self.rows), "reasons": reasons, "samples": len(self.rows)}


def cpu_reference_leg(n_patches, threads=None):
    """The reference's CPU path for the same workload, restated by the oracle (TensorFlow is not
    installable here): extract -> clip/rescale -> U-net (torch CPU fp32, all host
    threads) -> round -> fill, on `n_patches` 64^3 tiles of a float16 volume.  Returns voxels/s."""
    import torch
    from oracle import nets_torch as O, structures_np as S, weights as OW
    # all host cores, explicitly: torchrun exports OMP_NUM_THREADS=1, which would silently run this arm on one core
    torch.set_num_threads(threads or os.cpu_count() or 1)
    g = np.random.default_rng(5)
    nz = max(1, (n_patches + 3) // 4)
    vol = g.uniform(0, 1, (PATCH * nz, PATCH * 2, PATCH * 2)).astype(np.float16)
    pts, wds = S.regular_grid_points(vol.shape, (PATCH,) * 3)
    pts, wds = pts[:n_patches], wds[:n_patches]
    w = OW.draw(O.unet_spec(**M_A01), 11)
    t0 = time.perf_counter()
    x = S.extract(vol, pts, wds, (PATCH,) * 3)
    y = O.segmenter_predict_patches(x[..., np.newaxis], CHUNK, w, min_max=(0.0, 1.0), **M_A01)
    out = np.zeros(vol.shape, np.uint8)
    S.fill_patches_in_volume(y[..., 0], pts, wds, out)
    dt = time.perf_counter() - t0
    return n_patches * PATCH ** 3 / dt, dt, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # one full chunk per step: the reference pads a partial last chunk to chunk_size (surface_segmenter.py:277-281),
    # so a smaller sample would charge the CPU arm for 32 patches' worth of work while crediting it with fewer
    n_patches = CHUNK
    times = []
    for i in range(args.warmup + args.steps):
        v, dt, threads = cpu_reference_leg(n_patches)
        if i >= args.warmup:
            times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    value = n_patches * PATCH ** 3 / (ms / 1e3)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "voxels/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[1] model and patch size on the host CPU: %d tiles of 64^3 per step (bounded sample)" % n_patches,
                   "model": "M_a01", "patch": PATCH, "chunk": CHUNK},
        "cpu_baseline": {"value": value, "unit": "voxels/s", "cores": threads, "kind": "port",
                         "sample": "%d x 64^3 tiles per step: oracle extract + torch-CPU fp32 U-net + fill" % n_patches},
        "e2e": {"value": value, "unit": "voxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def fill_big_slab(host, z_lo, z_hi, dev):
    """Planes [z_lo, z_hi) of the synthetic 2048^3 float16 volume (blobs + noise in [0, 1]) into the pinned host tensor
    `host`; generated on the device in 128-plane blocks seeded by the block index, so every rank produces identical
    planes whatever its slab is."""
    import torch
    b0, b1 = z_lo // BIG_BLOCK, -(-z_hi // BIG_BLOCK)
    for b in range(b0, b1):
        g = torch.Generator(device=dev).manual_seed(9000 + b)
        coarse = torch.randn((1, 1, BIG_BLOCK // 16 + 1, BIG[1] // 16 + 1, BIG[2] // 16 + 1), generator=g, device=dev)
        x = torch.nn.functional.interpolate(coarse, size=(BIG_BLOCK, BIG[1], BIG[2]), mode="trilinear", align_corners=True)[0, 0]
        x = (x > 0).to(torch.float16)
        x.add_(torch.randn(x.shape, generator=g, device=dev, dtype=torch.float16), alpha=0.1)
        x.clamp_(-0.5, 1.5).add_(0.5).mul_(0.5)
        a, c = max(z_lo, b * BIG_BLOCK), min(z_hi, (b + 1) * BIG_BLOCK)
        host[a - z_lo:c - z_lo].copy_(x[a - b * BIG_BLOCK:c - b * BIG_BLOCK])
        del x, coarse
    torch.cuda.synchronize()


def leg_cfg4(fe, rank, world, dev, barrier, max_over_ranks, host_slab, z0, z1, reps):
    """BASELINE configs[3]: the rank's z-slab of ONE 2048^3 volume, host -> host."""
    import torch
    from tomoencoders_b200 import Grid, sharding
    grid = Grid(BIG, width=PATCH)
    loc, (a, b) = sharding.local_slab(grid, rank, world)
    assert (a, b) == (z0, z1)
    host_mask = torch.empty((z1 - z0,) + BIG[1:], dtype=torch.uint8, pin_memory=True)
    mm_box = [None]

    def step():
        lo, hi = fe.calc_voxel_min_max(host_slab, 4)             # uploads every 4th row of every 4th plane of the slab
        mm_box[0] = sharding.allreduce_minmax(lo, hi, device=dev)   # two scalars: the only exchange of this workload
        fe.segment_volume(host_slab, loc, CHUNK, min_max=mm_box[0], out=host_mask)
        torch.cuda.synchronize()

    step()                                                       # plans, staging buffers, chunk graphs
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        step()
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1)) / reps
    # checksum of the stitched mask (all ranks): identical at every N if the slabs reproduce the single-GPU result
    csum = torch.tensor([int(fe._staging_bufs["out"].sum(dtype=torch.int64).item())], dtype=torch.int64, device=dev)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(csum)
    nvox = BIG[0] * BIG[1] * BIG[2]
    res = {"metric": "voxels/sec segmented: ONE 2048^3 float16 volume, z-slab per rank, host volume in -> host uint8 mask out",
           "value": nvox / (ms / 1e3), "unit": "voxels/s", "ms_per_step": ms, "steps": reps, "scaling": "strong",
           "tiles_total": len(grid), "tiles_this_rank": len(loc), "slab_rank0": [z0, z1], "min_max": list(mm_box[0]),
           "h2d_bytes_per_step": int(nvox * 2 + nvox * 2 // 16), "d2h_bytes_per_step": int(nvox),
           "mask_checksum": int(csum.item()), "mask_fraction_ones": float(csum.item()) / nvox,
           "exchange": "all-reduce(MAX) of (-min, max): 2 float64" if world > 1 else "none (1 rank)"}
    del host_mask
    return res


def leg_cfg3(ae, rank, world, dev, barrier, max_over_ranks, host_planes, z_base, reps=1):
    """BASELINE configs[2]: 1 M random 64^3 patches of the 2048^3 volume -> latents -> PCA -> (n, 2), host -> host.
    host_planes holds the volume planes from z_base upwards that this rank needs."""
    import torch
    import torch.distributed as dist
    from tomoencoders_b200 import Patches, sharding
    from tomoencoders_b200.neural_nets import LatentPCA
    rng = np.random.default_rng(77)                               # the same list on every rank
    pts = np.stack([rng.integers(0, BIG[a] - PATCH, BIG_PATCHES) for a in range(3)], axis=1).astype(np.uint32)
    wds = np.full_like(pts, PATCH)
    idx, loc_pts, (z_lo, z_hi) = sharding.z_sorted_shard(pts, wds, BIG, rank, world)
    order = sharding.z_sorted_order(pts)
    assert z_lo >= z_base and z_hi - z_base <= host_planes.shape[0]
    src = host_planes[z_lo - z_base:z_hi - z_base]
    timings = {}

    def step(n_use=None):
        lp = loc_pts if n_use is None else loc_pts[:n_use]
        t = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        t[0].record()
        dvol = src.to(dev, non_blocking=True)                     # slab + halo of the host volume
        p = Patches(tuple(dvol.shape), "data", points=lp, widths=np.full_like(lp, PATCH))
        t[1].record()
        lat = ae.embed_volume(dvol, p, ENC_CHUNK, min_max=(0.0, 1.0))
        t[2].record()
        pca = LatentPCA(2)
        pca.partial_fit(lat)
        pca.finalize()                                            # all-reduce of 16 513 float64 moments + eigensolve
        t[3].record()
        z = pca.transform_all_gather(lat)                         # projection + all-gather of the (n_local, 2) rows
        z_host = z.cpu().numpy()
        t[4].record()
        torch.cuda.synchronize()
        out = sharding.unpermute(z_host, order) if n_use is None else z_host
        timings.update(upload_ms=t[0].elapsed_time(t[1]), encode_ms=t[1].elapsed_time(t[2]),
                       fit_ms=t[2].elapsed_time(t[3]), project_gather_ms=t[3].elapsed_time(t[4]))
        return out, pca

    step(n_use=min(len(loc_pts), 8 * ENC_CHUNK))                  # warm-up: plans, graphs (a short list)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        z_all, pca = step()
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1)) / reps
    assert z_all.shape == (BIG_PATCHES, 2)
    # the collectives on their own (device-timed, 10 repetitions each)
    coll = {"allreduce_moments_us": 0.0, "allgather_rows_us": 0.0}
    if world > 1:
        m = torch.zeros(1 + 128 + 128 * 128, dtype=torch.float64, device=dev)
        rows = torch.zeros((len(loc_pts), 2), dtype=torch.float32, device=dev)
        for name, fn in (("allreduce_moments_us", lambda: dist.all_reduce(m)),
                         ("allgather_rows_us", lambda: sharding.allgather_rows(rows))):
            fn()
            torch.cuda.synchronize()
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0.record()
            for _ in range(10):
                fn()
            c1.record()
            torch.cuda.synchronize()
            coll[name] = 1e3 * c0.elapsed_time(c1) / 10
    return {"metric": "64^3 patches/sec encoded + PCA-projected: 1 M random patches of ONE 2048^3 float16 volume, host volume + host coordinates in -> host (n, 2) rows out",
            "value": BIG_PATCHES / (ms / 1e3), "unit": "patches/s", "ms_per_step": ms, "steps": reps, "scaling": "strong",
            "patches_total": BIG_PATCHES, "patches_this_rank": int(len(loc_pts)), "planes_this_rank": [int(z_lo), int(z_hi)],
            "volume_placement": "slab + halo per rank: the z-sorted patch list is cut into N equal runs, rank r uploads only planes [z_lo, z_hi) (no broadcast of the volume)",
            "h2d_bytes_per_step_this_rank": int((z_hi - z_lo) * BIG[1] * BIG[2] * 2 + len(loc_pts) * 24),
            "d2h_bytes_per_step": int(BIG_PATCHES * 8),
            "phase_ms_rank0": {k: round(v, 3) for k, v in timings.items()},
            "collectives": dict(coll, what="all_reduce(SUM) of 16 513 float64 moments; all_gather of (n_local, 2) float32 rows (+ one int64 count each)"),
            "components_row0_head": [float(v) for v in pca.components_[0][:4]],
            "explained_variance": [float(v) for v in pca.explained_variance_],
            "z_checksum": float(np.abs(z_all).sum())}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from tomoencoders_b200 import Grid, Patches, _device, _lib
    from tomoencoders_b200.neural_nets import LatentPCA, SelfSupervisedCAE, SurfaceSegmenter

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    mode = os.environ.get("TOMOENC_MODE", "fp16")
    fe = SurfaceSegmenter(model_initialization="define-new", descriptor_tag="M_a01", mode=mode, **M_A01)
    # random-init weights: Keras-default glorot-uniform draws of the product's own initialiser (seeded)
    from tomoencoders_b200.neural_nets import _spec
    fe.models["segmenter"].set_weights(_spec.keras_default_init(fe.models["segmenter"].spec, seed=11))
    ae = SelfSupervisedCAE(model_initialization="define-new", model_size=(PATCH,) * 3, descriptor_tag="enc", mode=mode, **ENC)
    ae.models["encoder"].set_weights(_spec.keras_default_init(ae.models["encoder"].spec, seed=12))

    vol = synth_volume_torch(VOL, 4 + rank, dev)
    grid = Grid(VOL, width=PATCH)
    n_tiles = len(grid)
    min_max = fe.calc_voxel_min_max(vol, 4)
    seg = torch.zeros(VOL, dtype=torch.uint8, device=dev)
    rng = np.random.default_rng(7 + rank)
    enc_pts = np.stack([rng.integers(0, VOL[a] - PATCH, ENC_PATCHES) for a in range(3)], axis=1)
    enc_patches = Patches(VOL, "data", points=enc_pts, widths=np.full_like(enc_pts, PATCH))
    lat = torch.empty((ENC_PATCHES, 128), dtype=torch.float32, device=dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def seg_step():
        fe.segment_volume(vol, grid, CHUNK, min_max=min_max, out=seg)

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ------------------------------------------------------------------ segmentation, device resident
    for _ in range(args.warmup):
        seg_step()
    net = fe.models["segmenter"].net_for((PATCH,) * 3, CHUNK)
    barrier()
    launches0 = _lib.lib().te_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        ev0.record()
        for _ in range(args.steps):
            seg_step()
        ev1.record()
        barrier()
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    launches = int(_lib.lib().te_launch_count() - launches0)
    ms_per_step = ms_total / args.steps
    vox_per_step = VOL[0] * VOL[1] * VOL[2]
    value = world * vox_per_step / (ms_per_step / 1e3)

    # ------------------------------------------------------------------ per-layer roofline leg (separate pass, events per launch)
    net.set_profiling(True)            # (the chunk loop runs eagerly while profiling: the events sit between the launches)
    seg_step()
    torch.cuda.synchronize()
    lt = net.layer_times()
    net.set_profiling(False)

    # ------------------------------------------------------------------ the same step with bf16 activations (north_star's dtype; fp16 is the
    # default because bf16's 8-bit mantissa flips near-tie labels: tests/test_nets_gpu.py, profiles/r02_margin_hist.txt)
    other = "bf16" if mode != "bf16" else "fp16"
    fe_o = SurfaceSegmenter(model_initialization="define-new", descriptor_tag="M_a01", mode=other, **M_A01)
    fe_o.models["segmenter"].set_weights(fe.models["segmenter"].get_weights())
    for _ in range(2):
        fe_o.segment_volume(vol, grid, CHUNK, min_max=min_max, out=seg)
    barrier()
    o0, o1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    o0.record()
    n_o = max(1, min(args.steps, 3))
    for _ in range(n_o):
        fe_o.segment_volume(vol, grid, CHUNK, min_max=min_max, out=seg)
    o1.record()
    barrier()
    other_ms = max_over_ranks(o0.elapsed_time(o1)) / n_o
    other_line = {"dtype": other, "value": world * VOL[0] * VOL[1] * VOL[2] / (other_ms / 1e3), "unit": "voxels/s", "ms_per_step": other_ms}
    del fe_o
    fe.segment_volume(vol, grid, CHUNK, min_max=min_max, out=seg)
    by_kernel = {}
    for name, kernel, flops, ms, cnt in lt:
        if cnt == 0:
            continue
        k = by_kernel.setdefault(kernel.replace("+head", ""), {"ms": 0.0, "launches": 0, "flops": 0.0, "layers": []})
        k["ms"] += ms
        k["launches"] += cnt
        k["flops"] += flops * n_tiles          # algorithmic FLOPs over the whole step
        k["layers"].append(name)
    tot_ms = sum(k["ms"] for k in by_kernel.values())
    dom_name, dom = max(by_kernel.items(), key=lambda kv: kv[1]["ms"])
    n_chunks = -(-n_tiles // CHUNK)
    layers_tbl = {name: {"kernel": kernel, "us_per_chunk": round(1e3 * ms / cnt, 2),
                         "tflops": round(flops * CHUNK / (ms / cnt / 1e3) / 1e12, 1) if flops else None}
                  for name, kernel, flops, ms, cnt in lt if cnt}
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(peaks_path):
        peaks = json.load(open(peaks_path))
        tf_peak, peak_src = float(peaks.get("bf16_tflops_sustained", 1390.1)), "MEASURED_PEAKS.json bf16_tflops_sustained"
        hbm_peak = float(peaks.get("hbm_gbs", 6545.6))
    else:
        tf_peak, peak_src, hbm_peak = 1400.0, "fallback (B200_PROFILING.md sustained)", 6650.0
    achieved = dom["flops"] / (dom["ms"] / 1e3) / 1e12
    # DRAM traffic per launch of the dominant kernel: ncu --set full capture committed under profiles/ (one launch of
    # every layer at this very shape: 32 x 64^3 patches), averaged over the layers this kernel instantiation runs
    traffic, traffic_note = None, None
    tpath = os.path.join(ROOT, "profiles", "traffic_r02.json")
    if os.path.isfile(tpath) and CHUNK == 32:
        tj = json.load(open(tpath))["layers"]
        # only captures of THIS kernel configuration count (the file records the kernel each layer ran at capture time)
        got = [tj[l]["dram_read_bytes"] + tj[l]["dram_write_bytes"] for l in dom["layers"]
               if l in tj and tj[l].get("kernel", "").replace(" ", "") == dom_name.replace(" ", "")]
        if got and len(got) == len(dom["layers"]):
            traffic = sum(got) / len(got)
            traffic_note = "mean dram read+write bytes per launch over %s (profiles/traffic_r02.json, ncu --set full)" % \
                ", ".join(dom["layers"])
        else:
            traffic_note = "profiles/traffic_r02.json holds no capture of %s for every layer it runs: not reported" % dom_name
    roofline = {"bound": "tensor", "kernel": dom_name, "layers": dom["layers"], "achieved": achieved, "peak": tf_peak,
                "unit": "TFLOP/s", "frac": achieved / tf_peak, "traffic": traffic, "traffic_note": traffic_note,
                "peak_source": peak_src,
                "launches_per_step": dom["launches"], "avg_launch_ms": dom["ms"] / dom["launches"],
                # share of the timed step (not of the summed kernel time); the per-launch events of the profiling pass and
                # the timed loop are separate passes over the same work
                "share_of_step": dom["ms"] / ms_per_step, "share_of_kernel_time": dom["ms"] / tot_ms,
                "kernel_time_over_step": tot_ms / ms_per_step}
    all_tc_flops = sum(k["flops"] for n_, k in by_kernel.items() if n_.startswith("conv_umma"))
    all_tc_ms = sum(k["ms"] for n_, k in by_kernel.items() if n_.startswith("conv_umma"))
    kernels = {n_: {"ms_per_step": round(k["ms"], 3), "launches": k["launches"],
                    "tflops": (k["flops"] / (k["ms"] / 1e3) / 1e12) if k["flops"] else None} for n_, k in by_kernel.items()}

    # ------------------------------------------------------------------ gather / stitch bandwidth (HBM roofline of the data-movement kernels)
    # kernel-only: C-ABI calls with device-resident coordinates, CUDA events on the launching stream,
    # L2 flushed between repetitions; "api" = the same through Grid.extract / fill_patches_in_volume
    L = _lib.lib()
    d_pts = _device.points_to_device(grid.points)
    d_wds = _device.points_to_device(np.full((n_tiles, 3), PATCH))
    x_tiles = torch.empty((n_tiles,) + (PATCH,) * 3, dtype=vol.dtype, device=dev)
    lab = torch.zeros((n_tiles,) + (PATCH,) * 3, dtype=torch.uint8, device=dev)
    act = torch.empty((n_tiles,) + (PATCH,) * 3, dtype=net.act_dtype, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    flush_r = torch.zeros(256 << 18, dtype=torch.int32, device=dev)      # 256 MB read back after the write
    st = _device.stream_ptr()
    p3, vs3 = _lib.i32x3((PATCH,) * 3), _lib.i64x3(VOL)

    def timed_us(fn, reps=10):
        fn()
        tot = 0.0
        for _ in range(reps):
            # write 256 MB (evicts everything), then READ another 256 MB: the dirty lines of the write are written back
            # before the timed kernel starts instead of being charged to it (L2 is a write-back cache)
            flush.zero_()
            flush_r.sum()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            tot += a.elapsed_time(b)
        return 1e3 * tot / reps

    t_gather = timed_us(lambda: _lib.check(L.te_gather(vol.data_ptr(), 2, vs3, d_pts.data_ptr(), d_wds.data_ptr(), n_tiles, p3,
                                                       x_tiles.data_ptr(), _lib.TE_HINT_ALIGNED, st)))
    t_gnorm = timed_us(lambda: _lib.check(L.te_gather_norm(vol.data_ptr(), _lib.TE_F16, vs3, d_pts.data_ptr(), d_wds.data_ptr(),
                                                           n_tiles, p3, 1, 1, float(min_max[0]), float(min_max[1]),
                                                           act.data_ptr(), net.act_te, st)))
    t_scatter = timed_us(lambda: _lib.check(L.te_scatter(lab.data_ptr(), 1, d_pts.data_ptr(), n_tiles, p3, seg.data_ptr(), vs3,
                                                         None, _lib.TE_HINT_ALIGNED, st)))
    t_gather_api = timed_us(lambda: grid.extract(vol))
    t_scatter_api = timed_us(lambda: grid.fill_patches_in_volume(lab, seg))
    nvox = vol.numel()
    gather_gbs = 2 * nvox * 2 / t_gather / 1e3                     # read f16 + write f16
    gnorm_gbs = (nvox * 2 + nvox * act.element_size()) / t_gnorm / 1e3
    scatter_gbs = 2 * nvox / t_scatter / 1e3                       # read u8 + write u8
    del x_tiles, lab, act

    # random, unaligned uint8 corners: the access pattern of configs[0] (4096 x 32^3 of a 256^3 volume) and of configs[2]
    # (64^3 patches at random corners; here 16 384 of a 512^3 volume).  Algorithmic bytes: n P^3 written +
    # min(n P^3, |V|) read + 24 n of coordinates.
    def random_gather_case(vshape, P, n, seed):
        v8 = torch.randint(0, 256, vshape, dtype=torch.uint8, device=dev)
        rg = np.random.default_rng(seed)
        rp = np.stack([rg.integers(0, vshape[a] - P, n) for a in range(3)], axis=1).astype(np.uint32)
        dp, dw = _device.points_to_device(rp), _device.points_to_device(np.full_like(rp, P))
        o8 = torch.empty((n, P, P, P), dtype=torch.uint8, device=dev)
        vs_ = _lib.i64x3(vshape)
        p_ = _lib.i32x3((P,) * 3)
        us_plain = timed_us(lambda: _lib.check(L.te_gather(v8.data_ptr(), 1, vs_, dp.data_ptr(), dw.data_ptr(), n, p_, o8.data_ptr(),
                                                          _lib.TE_HINT_UNBINNED, st)))
        # the product path (structures/_table.py) hands the kernel a (z, y)-sorted processing order: output order unchanged
        perm = torch.argsort(dp[:, 0].to(torch.int64) * vshape[1] + dp[:, 1].to(torch.int64)).to(torch.int32)
        us = timed_us(lambda: _lib.check(L.te_gather_perm(v8.data_ptr(), 1, vs_, dp.data_ptr(), dw.data_ptr(), n, p_, o8.data_ptr(),
                                                          _lib.TE_HINT_UNBINNED, perm.data_ptr(), st)))
        us_api = timed_us(lambda: Patches(vshape, "data", points=rp, widths=np.full_like(rp, P)).extract(v8, (P,) * 3))
        alg = n * P ** 3 + min(n * P ** 3, v8.numel()) + 24 * n
        return {"us": us, "us_index_order": us_plain, "us_api": us_api, "algorithmic_bytes": alg, "gbs": alg / us / 1e3,
                "frac": alg / us / 1e3 / hbm_peak}
    rand_cfg1 = random_gather_case((256,) * 3, 32, 4096, 21)
    rand_cfg3 = random_gather_case((512,) * 3, 64, 16384, 22)
    del flush, flush_r

    # ------------------------------------------------------------------ end to end: host volume -> host mask through the public API
    host_vol = torch.empty(VOL, dtype=torch.float16).pin_memory()
    host_vol.copy_(vol)
    host_seg = torch.empty(VOL, dtype=torch.uint8).pin_memory()

    def e2e_step():
        # host arrays in, host array out: the package uploads the planes the min/max needs, then streams the volume up
        # and the finished mask slabs down on a side stream while the network runs
        mm = fe.calc_voxel_min_max(host_vol, 4)
        fe.segment_volume(host_vol, grid, CHUNK, min_max=mm, out=host_seg)
        torch.cuda.synchronize()

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        e2e_step()
    e1.record()
    barrier()
    e2e_ms = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    e2e_value = world * vox_per_step / (e2e_ms / 1e3)

    # ------------------------------------------------------------------ encode + PCA leg (patches/s)
    def enc_step():
        pca = LatentPCA(2)
        ae.embed_volume(vol, enc_patches, ENC_CHUNK, min_max=min_max, out=lat)
        pca.partial_fit(lat)
        pca.finalize()                      # all-reduce of the moments when world > 1
        return pca.transform_all_gather(lat)

    enc_step()
    barrier()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record()
    n_enc_steps = max(1, min(args.steps, 3))
    for _ in range(n_enc_steps):
        z2 = enc_step()
    c1.record()
    barrier()
    enc_ms = max_over_ranks(c0.elapsed_time(c1)) / n_enc_steps
    enc_value = world * ENC_PATCHES / (enc_ms / 1e3)
    enc_flops = ae.models["encoder"].net_for((PATCH,) * 3, ENC_CHUNK).flops_per_patch

    # ------------------------------------------------------------------ training leg (BASELINE configs[4]): fwd + bwd + Adam,
    # 64 patch pairs of 64^3 split over the ranks (8 per rank at N = 8; never fewer than 8 per rank here), SyncBN,
    # one flat gradient all-reduce per step
    train = None
    if os.environ.get("TOMOENC_BENCH_TRAIN", "1") != "0":
        from tomoencoders_b200.neural_nets._train import UnetTrainer
        del vol, seg
        torch.cuda.empty_cache()
        tb = max(8, 64 // world)
        tr = UnetTrainer(M_A01, (PATCH,) * 3, tb, mode="bf16", seed=13)
        gx = torch.Generator(device=dev).manual_seed(8 + rank)
        tx = torch.rand((tb,) + (PATCH,) * 3, generator=gx, device=dev)
        ty = (torch.rand((tb,) + (PATCH,) * 3, generator=gx, device=dev) < 0.5).to(torch.uint8)
        for _ in range(4):                  # two eager steps, the CUDA-graph capture, one replay
            tr.train_step(tx, ty)
        barrier()
        t0e, t1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches_t0 = _lib.lib().te_launch_count()
        t0e.record()
        n_tr = max(1, min(args.steps, 3))
        for _ in range(n_tr):
            tloss = tr.train_step(tx, ty)
        t1e.record()
        barrier()
        tr_ms = max_over_ranks(t0e.elapsed_time(t1e)) / n_tr
        train = {"metric": "training patch pairs/sec (U-net M_a01 fwd+bwd+Adam, 64^3 pairs, bf16 activations, fp32 master weights)",
                 "value": world * tb / (tr_ms / 1e3), "unit": "patch pairs/s", "ms_per_step": tr_ms, "global_batch": world * tb,
                 "per_rank_batch": tb, "loss": tloss, "tflops_effective_3x_fwd": 3 * world * tb * net.flops_per_patch / (tr_ms / 1e3) / 1e12,
                 "gpu_launches_per_step": int((_lib.lib().te_launch_count() - launches_t0) / n_tr) + tr.graph_launches,
                 "cuda_graph": tr.graph_launches > 0,
                 "collectives": "all-reduce(flat fp32 gradient, 4.77 M values) + SyncBN all-reduces of per-channel sums" if world > 1 else "none (1 rank)"}
        del tr, tx, ty

    # reconstruction leg, rank 0 only: one z chunk of a 2048-wide scan, 1500 angles -- ramp filter, then the
    # fused back-projection of every 32^3 patch of the chunk straight into clipped / rescaled fp16 network input
    recon = None
    if rank == 0 and not args.no_recon:
        from tomoencoders_b200 import reconstruction as RC
        r_n, r_nz, r_nt = 2048, 32, 1500
        rd = torch.randn((r_nt, r_nz, r_n), device=dev, generator=torch.Generator(device=dev).manual_seed(9))
        rth = torch.linspace(0, float(np.pi), r_nt + 1, device=dev)[:-1].contiguous()
        rpts = torch.from_numpy(Grid((r_nz, r_n, r_n), width=32).points.astype(np.int32)).to(dev)
        rout = torch.empty((len(rpts), 32, 32, 32), dtype=torch.float16, device=dev)

        def recon_step():
            RC.fbp_filter(rd)
            RC.BackProjector(rd, rth).patches(r_n / 2 + 0.5, rpts, 32, min_max=(-50.0, 50.0), out=rout)
        recon_step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            recon_step()
        e1.record()
        torch.cuda.synchronize()
        r_ms = e0.elapsed_time(e1) / 3
        updates = r_nt * r_nz * r_n * r_n
        sm_hz = clocks.summary()["sm_mhz"] * 1e6
        recon = {"metric": "voxels/sec reconstructed (FBP: ramp filter + back-projection of all 32^3 patches of a 1500 x 32 x 2048 chunk into fp16 network input)",
                 "value": r_nz * r_n * r_n / (r_ms / 1e3), "unit": "voxels/s", "ms_per_step": r_ms,
                 "updates_per_s": updates / (r_ms / 1e3),
                 "bound": "L1 data pipe: 8 bytes (two fp32 detector samples) per voxel-angle update through a 128 B/clk/SM load path",
                 "frac_of_l1_bound": updates / (r_ms / 1e3) / (148 * 16 * sm_hz),
                 "note": "ncu of the kernel: profiles/ncu_r01_recon.txt (l1tex data-pipe wavefronts 83 % of peak)"}
        del rd, rout

    # ------------------------------------------------------------------ BASELINE configs[3] and configs[2] at their stated sizes (all ranks)
    cfg4 = cfg3 = None
    if not args.no_big:
        from tomoencoders_b200 import sharding
        # single process: a failure of these two secondary, memory-hungry legs (17 GB of pinned host memory, 2048^3 of work)
        # must not cost the headline line -- it is reported inside the leg's entry instead.  With several ranks an exception
        # propagates (a rank that skips a collective would hang the others).
        try:
            torch.cuda.empty_cache()
            z0, z1 = sharding.slab_bounds(BIG, PATCH, rank, world)
            rng77 = np.random.default_rng(77)
            zz = np.sort(rng77.integers(0, BIG[0] - PATCH, BIG_PATCHES))          # the z column of leg_cfg3's list (first draw)
            a3, b3 = sharding.shard_range(BIG_PATCHES, rank, world)
            zl3, zh3 = int(zz[a3]), int(zz[b3 - 1]) + PATCH
            z_base, z_top = min(z0, zl3), max(z1, zh3)
            host_planes = torch.empty((z_top - z_base,) + BIG[1:], dtype=torch.float16, pin_memory=True)
            fill_big_slab(host_planes, z_base, z_top, dev)
            cfg4 = leg_cfg4(fe, rank, world, dev, barrier, max_over_ranks, host_planes[z0 - z_base:z1 - z_base], z0, z1,
                            reps=1 if world == 1 else 2)
            fe.__dict__.pop("_staging_bufs", None)
            torch.cuda.empty_cache()
            cfg3 = leg_cfg3(ae, rank, world, dev, barrier, max_over_ranks, host_planes, z_base)
            del host_planes
        except Exception as e:  # noqa: BLE001
            if world > 1:
                raise
            err = {"error": "%s: %s" % (type(e).__name__, e)}
            print("bench: 2048^3 legs failed: %s" % err["error"], file=sys.stderr)
            cfg4 = cfg4 if cfg4 is not None else err
            cfg3 = cfg3 if cfg3 is not None else err

    if rank == 0:
        cpu_baseline = None                  # timed on rank 0 at N = 1 only (the other ranks would idle on it)
        if world == 1:
            cpu_v, cpu_dt, cpu_threads = cpu_reference_leg(CPU_TILES)
            cpu_baseline = {"value": cpu_v, "unit": "voxels/s", "cores": cpu_threads, "kind": "port",
                            "sample": "%d x 64^3 tiles: oracle extract + torch-CPU fp32 U-net M_a01 + fill (%.1f s)" % (CPU_TILES, cpu_dt)}
        flops_patch = net.flops_per_patch
        line = {
            "metric": METRIC, "value": value, "unit": "voxels/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": mode, "data": "synthetic",
            "config": {"workload": "BASELINE configs[1]: U-net M_a01 segmenter, 512^3 float16 volume per GPU, 512 tiles of 64^3, chunk %d" % CHUNK,
                       "model": "M_a01 n_filters=[16,32,64] pool=[2,2,2] isconcat=True", "patch": PATCH, "chunk": CHUNK,
                       "parallelism": "independent slabs (1 per rank), no data-path collective" if world > 1 else "single GPU",
                       "l2": "inputs_exceed_l2 (268 MB volume + %.1f GB activations per chunk vs 126 MB L2)" % (net.workspace_bytes / 1e9),
                       "gflop_per_patch": flops_patch / 1e9},
            "tflops_effective": world * n_tiles * flops_patch / (ms_per_step / 1e3) / 1e12,
            "clocks": clocks.summary(),
            "e2e": {"value": e2e_value, "unit": "voxels/s", "ms_per_step": e2e_ms,
                    # the volume once + every 4th row of every 4th plane once more (one strided DMA) for the min/max that has
                    # to precede the first chunk
                    "h2d_bytes_per_step": int(host_vol.numel() * 2 + (-(-VOL[0] // 4)) * (-(-VOL[1] // 4)) * VOL[2] * 2),
                    "d2h_bytes_per_step": int(host_seg.numel()),
                    "api": "calc_voxel_min_max(host volume, 4) + SurfaceSegmenter.segment_volume(host volume, grid, chunk, min_max, out=host mask): pinned host tensors, slab-wise H2D / D2H overlapped with the network"},
            "gpu_launches": launches,
            "roofline": roofline,
            "roofline_all_tensor_core_layers": {"achieved": all_tc_flops / (all_tc_ms / 1e3) / 1e12, "unit": "TFLOP/s",
                                                "frac": all_tc_flops / (all_tc_ms / 1e3) / 1e12 / tf_peak,
                                                "share_of_step": all_tc_ms / ms_per_step},
            "kernels": kernels,
            "layers": layers_tbl,
            "other_dtype": other_line,
            "cfg4": cfg4,
            "cfg3": cfg3,
            "hbm": {"gather_gbs": gather_gbs, "gather_frac": gather_gbs / hbm_peak,
                    "gather_norm_gbs": gnorm_gbs, "gather_norm_frac": gnorm_gbs / hbm_peak,
                    "scatter_gbs": scatter_gbs, "scatter_frac": scatter_gbs / hbm_peak, "peak": hbm_peak, "unit": "GB/s",
                    "random_u8_cfg1_4096x32^3_of_256^3": rand_cfg1, "random_u8_cfg3like_16384x64^3_of_512^3": rand_cfg3,
                    "us": {"gather": t_gather, "gather_norm": t_gnorm, "scatter": t_scatter,
                           "gather_api": t_gather_api, "scatter_api": t_scatter_api},
                    "note": "kernel-only (C ABI, CUDA events, L2 flushed by a 256 MB write followed by a 256 MB read so that no dirty lines remain): te_gather of the 512^3 fp16 volume into 512 tiles "
                            "(read+write 537 MB), te_gather_norm (fused clip/rescale, the product path), te_scatter of the "
                            "uint8 mask (read+write 268 MB); *_api = the same through Grid.extract / fill_patches_in_volume"},
            "encode": {"metric": "64^3 patches/sec encoded + PCA-projected (encoder [16,32,64], latent 128)",
                       "value": enc_value, "unit": "patches/s", "ms_per_step": enc_ms, "patches_per_rank_per_step": ENC_PATCHES, "chunk": ENC_CHUNK,
                       "tflops_effective": enc_value * enc_flops / 1e12,
                       "collectives": "allreduce(16513 f64 PCA moments) + allgather((n,2) projections)" if world > 1 else "none (1 rank)"},
            "train": train,
            "recon": recon,
            "cpu_baseline": cpu_baseline,
        }
        if cfg3 is not None and "value" in cfg3:
            # "patches/s encoded" end to end = the configs[2] leg above: host volume + host coordinates in, host (n, 2) rows out
            line["encode"]["e2e"] = {"value": cfg3["value"], "unit": cfg3["unit"], "ms_per_step": cfg3["ms_per_step"],
                                     "h2d_bytes_per_step": cfg3["h2d_bytes_per_step_this_rank"],
                                     "d2h_bytes_per_step": cfg3["d2h_bytes_per_step"],
                                     "api": "Patches + SelfSupervisedCAE.embed_volume(host slab, ...) + LatentPCA.partial_fit / finalize / "
                                            "transform_all_gather -> numpy rows (the `cfg3` leg: 1 M patches of one 2048^3 volume)"}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-recon", action="store_true", help="skip the reconstruction leg")
    ap.add_argument("--no-big", action="store_true", help="skip the 2048^3 legs (BASELINE configs[2] / configs[3])")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
