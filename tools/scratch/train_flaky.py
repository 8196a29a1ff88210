import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from scipy.ndimage import gaussian_filter
from tomoencoders_b200 import Patches
from tomoencoders_b200.neural_nets import SurfaceSegmenter
SMALL = dict(n_filters=[16, 32], n_blocks=2, activation="lrelu", batch_norm=True, isconcat=[True, True], pool_size=[2, 2])
g = np.random.default_rng(1)
b = gaussian_filter(g.standard_normal((96, 96, 96)), 3)
Y = (b > np.quantile(b, 0.6)).astype(np.uint8)
X = (Y + 0.1 * g.standard_normal(Y.shape)).astype(np.float32)
for ne in (3, 6):
    for seed in range(4):
        np.random.seed(seed)
        fe = SurfaceSegmenter(model_initialization="define-new", descriptor_tag="t", mode="bf16", **SMALL)
        hist = fe.train([X], [Y], (32, 32, 32), 3, 8, 0.1, True, 0.05, ne, steps_per_epoch=10, verbose=0)
        p = Patches(X.shape, "regular-grid", patch_size=(32,) * 3)
        yp = fe.predict_patches("segmenter", p.extract(X, (32,) * 3)[..., np.newaxis], 8, None, min_max=(0.0, 1.0))
        acc = float((yp[..., 0] == p.extract(Y, (32,) * 3)).mean())
        print(ne, seed, ["%.3f" % h for h in hist], "acc %.4f" % acc, flush=True)
