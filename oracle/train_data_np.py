"""numpy restatement of the segmenter's training data generator.  TEST INFRASTRUCTURE.

Follows /root/reference/tomo_encoders/neural_nets/keras_processor.py:213-232 (extract_training_patch_pairs) and
neural_nets/surface_segmenter.py:75-82 (_find_blanks) line by line, with the random draws made by the caller so that the
deterministic part (binned gather + np.rot90) can be compared bit for bit.  Pinned for the gather by oracle/structures_np
(itself pinned by reference-generated golden vectors); np.rot90 / np.std are numpy's own."""
import numpy as np

from . import structures_np as S


def find_blanks_std(Y_gt, points, widths, input_size):
    """surface_segmenter.py:78-79: np.std of every extracted label patch (float64)."""
    y_tmp = S.extract(Y_gt, points, widths, input_size)[..., np.newaxis]
    return np.std(y_tmp, axis=(1, 2, 3))[:, 0]


def training_pairs(X, Y, points, widths, input_size, noise, nrots, axes):
    """keras_processor.py:218-230 with the random quantities given: noise (n,P,P,P,1) or None, nrots (n,) or None,
    axes (n,2).  Returns x (float64 like the reference: x + float64 noise) and y."""
    y = S.extract(Y, points, widths, input_size)[..., np.newaxis]
    x = S.extract(X, points, widths, input_size)[..., np.newaxis]
    if noise is not None:
        x = x + noise
    else:
        x = x.copy()
    y = y.copy()
    if nrots is not None:
        for ii in range(len(points)):
            ax = tuple(int(a) for a in axes[ii])
            x[ii, ..., 0] = np.rot90(x[ii, ..., 0], k=int(nrots[ii]), axes=ax)
            y[ii, ..., 0] = np.rot90(y[ii, ..., 0], k=int(nrots[ii]), axes=ax)
    return x, y
