"""Developer diagnostic (GPU): per-layer comparison of the tcgen05 modes against the fp32 check mode,
and of the check mode against the torch-CPU oracle.  Not a test; prints a table."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from oracle import nets_torch as O, weights as OW  # noqa: E402
from tomoencoders_b200.neural_nets._net import Net  # noqa: E402


def rel(a, b):
    a = a.double().flatten(); b = b.double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30)), float((a - b).abs().max())


def run(kind, patch, params, nb=2, modes=("bf16", "fp16"), seed=11):
    dev = torch.device("cuda")
    g = np.random.default_rng(seed)
    x = g.uniform(0, 1, (nb,) + patch + (1,)).astype(np.float32)
    if kind == "unet":
        spec = O.unet_spec(**params)
    else:
        spec, _, _ = O.encoder_spec(patch, **params)
    w = OW.draw(spec, seed)
    t0 = time.time()
    if kind == "unet":
        ref = O.unet_forward(x, w, return_logits=True, **params)[..., 0]
    else:
        ref = O.encoder_forward(x, w, **params)
    print("oracle (torch CPU fp32) %.2f s; output std %.4f" % (time.time() - t0, ref.std()))
    nets = {}
    outs = {}
    for mode in ("fp32",) + tuple(modes):
        net = Net(kind, patch, nb, mode=mode, **params)
        net.set_weights(w)
        xt = torch.from_numpy(x[..., 0]).to(dev).to(net.act_dtype).contiguous()
        torch.cuda.synchronize()
        t0 = time.time()
        if kind == "unet":
            _, logits = net.unet_forward(xt, want_logits=True)
            out = logits
        else:
            out = net.encoder_forward(xt)
        torch.cuda.synchronize()
        print("mode %-5s forward %.1f ms (first call)  workspace %.1f MB" % (mode, (time.time() - t0) * 1e3, net.workspace_bytes / 1e6))
        nets[mode] = net
        outs[mode] = out.float().cpu()
    r = rel(outs["fp32"], torch.from_numpy(ref))
    print("fp32 check vs oracle: rel %.3e  maxabs %.3e" % r)
    for mode in modes:
        print("%s vs oracle: rel %.3e maxabs %.3e" % ((mode,) + rel(outs[mode], torch.from_numpy(ref))))
        if kind == "unet":
            agree = float(((outs[mode] > 0) == (torch.from_numpy(ref) > 0)).float().mean())
            print("   label agreement vs oracle: %.5f" % agree)
    # per layer
    layers = nets["fp32"].layers()
    for mode in modes:
        print("--- per-layer %s vs fp32 check" % mode)
        for i, (name, fl, tc) in enumerate(nets[mode].layers()):
            try:
                n32 = nets["fp32"]._L.te_net_layer_output
                # element count is unknown here: ask with a generous capacity
                cap = 1 << 28
                buf32 = torch.empty(cap // 8, dtype=torch.float32, device=dev)
                c32 = nets["fp32"]._L.te_net_layer_output(nets["fp32"]._h, i, buf32.data_ptr(), buf32.numel(), None)
                buf = torch.empty(cap // 8, dtype=torch.float32, device=dev)
                c = nets[mode]._L.te_net_layer_output(nets[mode]._h, i, buf.data_ptr(), buf.numel(), None)
                torch.cuda.synchronize()
                if c32 <= 0 or c != c32:
                    print("%2d %-18s (no stored tensor)" % (i, name))
                    continue
                e = rel(buf[:c], buf32[:c])
                print("%2d %-18s tc=%d n=%9d rel %.3e maxabs %.3e" % (i, name, tc, c, e[0], e[1]))
            except Exception as ex:  # noqa: BLE001
                print("%2d %-18s error %s" % (i, name, ex))


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "small"
    if which == "small":
        run("unet", (32, 32, 32), dict(n_filters=[16, 32], n_blocks=2, pool_size=[2, 2], isconcat=[True, True]))
        run("encoder", (32, 32, 32), dict(n_filters=[16, 32, 64], n_blocks=3, pool_size=[2, 2, 2], hidden_units=[128, 128], POOL_FLAG=True))
    elif which == "ma01":
        run("unet", (64, 64, 64), dict(n_filters=[16, 32, 64], n_blocks=3, pool_size=[2, 2, 2], isconcat=[True] * 3), nb=2)
