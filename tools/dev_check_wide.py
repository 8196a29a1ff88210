"""Dev check: U-net with 512 bottleneck channels (per-tile bias staging path of conv_umma) against the torch-CPU oracle."""
import numpy as np, sys
import os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle import nets_torch as O, weights as OW
from tomoencoders_b200.neural_nets import SurfaceSegmenter
P = dict(n_filters=[32, 64, 128], n_blocks=3, activation="lrelu", batch_norm=True, isconcat=[True] * 3, pool_size=[2, 2, 2])
w = OW.draw(O.unet_spec(**P), 31)
fe = SurfaceSegmenter(model_initialization="define-new", descriptor_tag="M_a03", mode="fp16", **P)
fe.models["segmenter"].set_weights(w)
x = np.random.default_rng(0).uniform(0, 1, (3, 32, 32, 32, 1)).astype(np.float32)
y = fe.predict_patches("segmenter", x, 2, None, min_max=(0.0, 1.0))
ref = O.segmenter_predict_patches(x, 2, w, min_max=(0.0, 1.0), **P)
print("agreement", float((y == ref).mean()), "ref mean", float(ref.mean()))
assert (y == ref).mean() > 0.995
print("wide-channel (512) path ok")
