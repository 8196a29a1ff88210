"""Profiling target: ONE chunk of the BASELINE configs[1] path through the C ABI -- te_gather_norm (32 tiles of a 512^3
float16 volume) -> te_unet_forward (M_a01, 17 fused layers) -> te_scatter (uint8 mask) -- executed twice (warm-up +
the pass to capture).  Writes the kernel order of one pass to gpurun_out/r02_layer_names.json.

    python tools/prof_chunk.py                          # plain run (exit 0 before any ncu run)
    ncu --set full --clock-control none --import-source on -k regex:'conv|gather|scatter' --launch-skip 19 --launch-count 19 \
        -o gpurun_out/r02_layers python tools/prof_chunk.py
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tomoencoders_b200 import Grid, _device, _lib  # noqa: E402
from tomoencoders_b200.neural_nets import _spec  # noqa: E402
from tomoencoders_b200.neural_nets._net import Net  # noqa: E402

M_A01 = dict(n_filters=[16, 32, 64], n_blocks=3, pool_size=[2, 2, 2], isconcat=[True] * 3)
VOL, P, CHUNK = (512, 512, 512), 64, 32


def main():
    mode = os.environ.get("TOMOENC_MODE", "fp16")
    net = Net("unet", (P,) * 3, CHUNK, mode=mode, **M_A01)
    net.set_weights(_spec.keras_default_init(_spec.unet_spec(**M_A01), seed=11))
    vol = torch.rand(VOL, device="cuda").to(torch.float16)
    seg = torch.zeros(VOL, dtype=torch.uint8, device="cuda")
    grid = Grid(VOL, width=P)
    pts = _device.points_to_device(grid.points[64:64 + CHUNK])
    wds = _device.points_to_device(np.full((CHUNK, 3), P))
    xa = torch.empty((CHUNK,) + (P,) * 3, dtype=net.act_dtype, device="cuda")
    y = torch.empty((CHUNK,) + (P,) * 3, dtype=torch.uint8, device="cuda")
    L, st = _lib.lib(), _device.stream_ptr()
    vs, p3 = _lib.i64x3(VOL), _lib.i32x3((P,) * 3)
    for _ in range(2):
        _lib.check(L.te_gather_norm(vol.data_ptr(), _lib.TE_F16, vs, pts.data_ptr(), wds.data_ptr(), CHUNK, p3, 1, 1, 0.0, 1.0,
                                    xa.data_ptr(), net.act_te, st))
        net.unet_forward_into(xa, y)
        _lib.check(L.te_scatter(y.data_ptr(), 1, pts.data_ptr(), CHUNK, p3, seg.data_ptr(), vs, None, _lib.TE_HINT_ALIGNED, st))
    torch.cuda.synchronize()
    names = ["gather_norm"]
    net.set_profiling(True)
    net.unet_forward_into(xa, y)
    torch.cuda.synchronize()
    for name, kernel, flops, ms, cnt in net.layer_times():
        if cnt:
            names.append({"layer": name, "kernel": kernel, "gflop_per_chunk": flops * CHUNK / 1e9})
    names.append("scatter")
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    if not os.environ.get("NV_COMPUTE_PROFILER_PERFWORKS_DIR") and "--names" in sys.argv:
        json.dump(names, open(os.path.join(ROOT, "gpurun_out", "r02_layer_names.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
