"""Component timing of bench.py's encode leg (gather_norm -> encoder -> PCA moments -> eigensolve -> projection)
on one GPU: where the time of a 4096-patch step goes."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from tomoencoders_b200 import Patches, _device  # noqa: E402
from tomoencoders_b200.neural_nets import LatentPCA, SelfSupervisedCAE, _spec  # noqa: E402
from tomoencoders_b200.neural_nets.keras_processor import gather_normalized  # noqa: E402

ENC = dict(n_filters=[16, 32, 64], n_blocks=3, activation="lrelu", batch_norm=True, pool_size=[2, 2, 2],
           hidden_units=[128, 128], POOL_FLAG=True)
VOL, P, N, CHUNK = (512,) * 3, 64, 4096, 128


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    dev = torch.device("cuda", 0)
    mode = os.environ.get("TOMOENC_MODE", "fp16")
    ae = SelfSupervisedCAE(model_initialization="define-new", model_size=(P,) * 3, descriptor_tag="enc", mode=mode, **ENC)
    ae.models["encoder"].set_weights(_spec.keras_default_init(ae.models["encoder"].spec, seed=12))
    vol = torch.rand(VOL, device=dev).to(torch.float16)
    rng = np.random.default_rng(7)
    pts = np.stack([rng.integers(0, VOL[a] - P, N) for a in range(3)], axis=1)
    patches = Patches(VOL, "data", points=pts, widths=np.full_like(pts, P))
    lat = torch.empty((N, 128), dtype=torch.float32, device=dev)
    net = ae.models["encoder"].net_for((P,) * 3, CHUNK)
    d_pts = _device.points_to_device(patches.points)
    d_wds = _device.points_to_device(patches.widths)
    mm = (0.0, 1.0)

    t_embed = timed(lambda: ae.embed_volume(vol, patches, CHUNK, min_max=mm, out=lat))
    xa = gather_normalized(vol, VOL, d_pts[:CHUNK], d_wds[:CHUNK], CHUNK, (P,) * 3, mm, False, net)
    t_g = timed(lambda: gather_normalized(vol, VOL, d_pts[:CHUNK], d_wds[:CHUNK], CHUNK, (P,) * 3, mm, False, net), 10)
    t_e = timed(lambda: net.encoder_forward(xa, out=lat[:CHUNK]), 10)
    pca = LatentPCA(2)
    t_acc = timed(lambda: pca.partial_fit(lat), 5)
    t_fin = timed(lambda: pca.finalize(), 3)
    t_tr = timed(lambda: pca.transform_all_gather(lat), 5)
    print("embed_volume %d patches: %.2f ms  (%.0f patches/s)" % (N, t_embed, N / t_embed * 1e3))
    print("  per %d-chunk: gather_norm %.3f ms, encoder %.3f ms  -> x%d = %.2f ms" %
          (CHUNK, t_g, t_e, N // CHUNK, (t_g + t_e) * (N // CHUNK)))
    print("  PCA: accumulate %.3f ms, finalize (solve + readback) %.3f ms, transform %.3f ms" % (t_acc, t_fin, t_tr))
    print("  sum = %.2f ms -> %.0f patches/s" % (t_embed + t_acc + t_fin + t_tr, N / (t_embed + t_acc + t_fin + t_tr) * 1e3))


if __name__ == "__main__":
    main()
