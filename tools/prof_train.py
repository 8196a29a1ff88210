"""Times one training step of the BASELINE configs[4] model (U-net M_a01, 64^3 patch pairs) and prints
a per-operator breakdown (CUDA events around every C-ABI call of the graph walker)."""
import os
import sys
import time
from collections import OrderedDict

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from tomoencoders_b200.neural_nets import _train  # noqa: E402

M_A01 = dict(n_filters=[16, 32, 64], n_blocks=3, pool_size=[2, 2, 2], isconcat=[True] * 3)


def main():
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    size = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    tr = _train.UnetTrainer(M_A01, (size,) * 3, batch, mode="bf16", seed=1)
    g = np.random.default_rng(0)
    x = torch.from_numpy(g.uniform(0, 1, (batch,) + (size,) * 3).astype(np.float32)).cuda()
    y = (x > 0.5).to(torch.uint8)
    for _ in range(4):                      # two eager steps, the graph capture, one replay
        loss = tr.train_step(x, y)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = 3
    for _ in range(n):
        loss = tr.train_step(x, y)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    flops = 3 * 122.23e9 * batch * (size / 64) ** 3
    print(("train step [%s], batch" % ("graph replay" if getattr(tr, "_graph", None) is not None else "eager")) + " %d x %d^3: %.1f ms (host wall %.1f ms)  loss %.4f  ~%.0f TFLOP/s (3 x fwd FLOPs)  workspace %.2f GB" %
          (batch, size, ms, 1e3 * (time.perf_counter() - t0) / n, loss, flops / ms / 1e9, tr.workspace_bytes() / 1e9))
    # per-operator breakdown
    times = OrderedDict()
    orig = tr._call

    def timed(name, *args):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        orig(name, *args)
        b.record()
        times.setdefault(name, []).append((a, b))
    tr._call = timed
    os.environ["TOMOENC_TRAIN_GRAPH"] = "0"        # the per-operator breakdown needs the eager path
    tr.train_step(x, y)
    torch.cuda.synchronize()
    tot = 0.0
    rows = []
    for name, evs in times.items():
        t = sum(a.elapsed_time(b) for a, b in evs)
        rows.append((t, name, len(evs)))
        tot += t
    for t, name, cnt in sorted(rows, reverse=True):
        print("  %-28s %4d calls %9.2f ms  %5.1f%%" % (name, cnt, t, 100 * t / tot))
    print("  sum of operator times %.1f ms" % tot)
    for name in ("te_tr_conv3_wgrad", "te_tr_conv3", "te_tr_upconv_bwd"):
        print("  %s per call (ms, call order):" % name, " ".join("%.2f" % a.elapsed_time(b) for a, b in times[name]))


if __name__ == "__main__":
    main()
