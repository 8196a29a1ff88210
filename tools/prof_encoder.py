"""Per-layer timing of the BASELINE configs[2] encoder (n_filters [16,32,64], hidden [128,128], 64^3 patches)."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from tomoencoders_b200.neural_nets import _spec  # noqa: E402
from tomoencoders_b200.neural_nets._net import Net  # noqa: E402

ENC = dict(n_filters=[16, 32, 64], n_blocks=3, pool_size=[2, 2, 2], hidden_units=[128, 128], POOL_FLAG=True)


def main():
    chunk = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    mode = os.environ.get("TOMOENC_MODE", "fp16")
    net = Net("encoder", (64, 64, 64), chunk, mode=mode, **ENC)
    net.set_weights(_spec.keras_default_init(_spec.encoder_spec((64, 64, 64), **ENC), seed=12))
    x = torch.rand((chunk, 64, 64, 64), device="cuda").to(net.act_dtype)
    for _ in range(3):
        z = net.encoder_forward(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        z = net.encoder_forward(x)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print("encoder forward of %d patches: %.3f ms (%.0f patches/s, %.1f TFLOP/s effective)" %
          (chunk, ms, chunk / ms * 1e3, chunk * net.flops_per_patch / ms / 1e9))
    net.set_profiling(True)
    for _ in range(3):
        net.encoder_forward(x)
    torch.cuda.synchronize()
    for name, kernel, flops, ms_l, cnt in net.layer_times():
        if cnt:
            print("  %-18s %-34s %8.1f us" % (name, kernel, 1e3 * ms_l / cnt))


if __name__ == "__main__":
    main()
