"""Short profiling target (used under ncu): a few forward passes of the BASELINE configs[1] U-net
(M_a01, 64^3 patches, chunk of 32) with synthetic inputs and random-init weights."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from tomoencoders_b200.neural_nets import _spec  # noqa: E402
from tomoencoders_b200.neural_nets._net import Net  # noqa: E402

M_A01 = dict(n_filters=[16, 32, 64], n_blocks=3, pool_size=[2, 2, 2], isconcat=[True] * 3)


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    mode = os.environ.get("TOMOENC_MODE", "fp16")
    net = Net("unet", (64, 64, 64), chunk, mode=mode, **M_A01)
    net.set_weights(_spec.keras_default_init(_spec.unet_spec(**M_A01), seed=11))
    x = torch.rand((chunk, 64, 64, 64), device="cuda").to(net.act_dtype)
    out = torch.empty((chunk, 64, 64, 64), dtype=torch.uint8, device="cuda")
    for _ in range(reps):
        net.unet_forward_into(x, out)
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    net.unet_forward_into(x, out)
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    print("forward of %d patches: %.3f ms  (%.1f TFLOP/s effective, %.1f Mvox/s)" %
          (chunk, ms, chunk * net.flops_per_patch / ms / 1e9, chunk * 64 ** 3 / ms / 1e3))
    print("label mean", float(out.float().mean()))
    if os.environ.get("PROF_LAYERS"):
        net.set_profiling(True)
        for _ in range(3):
            net.unet_forward_into(x, out)
        torch.cuda.synchronize()
        for name, kernel, flops, ms_l, cnt in net.layer_times():
            if cnt:
                print("  %-18s %-34s %8.1f us  %7.1f TFLOP/s" % (name, kernel, 1e3 * ms_l / cnt, chunk * flops / (ms_l / cnt) / 1e9))
        net.set_profiling(False)


if __name__ == "__main__":
    main()
