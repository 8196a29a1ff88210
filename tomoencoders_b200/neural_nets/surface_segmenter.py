"""``SurfaceSegmenter``: drop-in for /root/reference/tomo_encoders/neural_nets/surface_segmenter.py:36-350
(3-D U-net; inference: clip + rescale + predict + round -> uint8 labels; training: random multi-scale
patch pairs with blank rejection, noise and rot90 augmentation -> Adam / binary cross-entropy fit steps)."""
import os
import time

import numpy as np
import torch

from .. import _device, _lib
from ..structures.patches import Patches
from .keras_processor import Vox2VoxProcessor_fCNN
from .Unet3D import build_Unet_3D

MAX_ITERS = 2000                    # surface_segmenter.py:27


def _patch_std(vol, patches, input_size):
    """(n,) float32 CUDA tensor: np.std of every binned patch of vol (te_patch_std)."""
    t = _device.to_supported(_device.to_device(vol))
    n = len(patches)
    out = torch.empty(max(n, 1), dtype=torch.float32, device=t.device)
    d_pts = _device.points_to_device(patches.points)
    d_wds = _device.points_to_device(patches.widths)
    _lib.check(_lib.lib().te_patch_std(t.data_ptr(), _device.te_dtype(t), _lib.i64x3(patches.vol_shape), d_pts.data_ptr(),
                                       d_wds.data_ptr(), n, _lib.i32x3(input_size), out.data_ptr(), _device.stream_ptr()),
               "te_patch_std")
    return out[:n]


def _training_pairs(X, Y, patches, input_size, sigma, nrots, axes, seed):
    """x (n,P,P,P,1) float32 and y (n,P,P,P,1) uint8 CUDA tensors (te_training_pairs)."""
    tx = _device.to_supported(_device.to_device(X))
    ty = _device.to_supported(_device.to_device(Y))
    assert tuple(tx.shape) == tuple(patches.vol_shape) == tuple(ty.shape), "volume shapes do not match vol_shape"
    n = len(patches)
    dev = tx.device
    x = torch.empty((n,) + tuple(input_size), dtype=torch.float32, device=dev)
    y = torch.empty((n,) + tuple(input_size), dtype=torch.uint8, device=dev)
    d_pts = _device.points_to_device(patches.points)
    d_wds = _device.points_to_device(patches.widths)
    d_sig = None if sigma is None else torch.from_numpy(np.ascontiguousarray(sigma, dtype=np.float32)).to(dev)
    d_rot = None if nrots is None else torch.from_numpy(np.ascontiguousarray(nrots, dtype=np.int32)).to(dev)
    d_ax = None if axes is None else torch.from_numpy(np.ascontiguousarray(axes, dtype=np.int32)).to(dev)
    _lib.check(_lib.lib().te_training_pairs(tx.data_ptr(), _device.te_dtype(tx), ty.data_ptr(), _device.te_dtype(ty),
                                            _lib.i64x3(patches.vol_shape), d_pts.data_ptr(), d_wds.data_ptr(), n,
                                            _lib.i32x3(input_size), None if d_sig is None else d_sig.data_ptr(),
                                            None if d_rot is None else d_rot.data_ptr(),
                                            None if d_ax is None else d_ax.data_ptr(), int(seed) & (2 ** 64 - 1),
                                            x.data_ptr(), y.data_ptr(), _device.stream_ptr()), "te_training_pairs")
    return x.unsqueeze(-1), y.unsqueeze(-1)


class SurfaceSegmenter(Vox2VoxProcessor_fCNN):
    output_type = "labels"          # predict_patches always rounds to uint8 (surface_segmenter.py:291)
    _clip_input = True              # np.clip(x, *min_max) before the rescale (surface_segmenter.py:272-275)

    def _build_models(self, descriptor_tag="misc", **model_params):
        """surface_segmenter.py:297-324."""
        if model_params is None:
            raise ValueError("Need model hyperparameters or instance of model. Neither were provided")
        self.models = {}
        self.model_tag = "Unet_%s" % (descriptor_tag)
        self.models["segmenter"] = build_Unet_3D(mode=self.mode, **model_params)

    def _load_models(self, model_names=None, model_path="some/path", **model_params):
        """surface_segmenter.py:134-154, with ``.npz`` weight files written by ``save_models`` (the
        Keras HDF5 container needs TensorFlow / h5py); the architecture comes from ``model_params``."""
        self.models = {}
        for model_key, model_name in model_names.items():
            m = build_Unet_3D(mode=self.mode, **model_params)
            m.load_weights(os.path.join(model_path, model_name + ".npz"))
            self.models[model_key] = m
        self.model_tag = "_".join(model_names["segmenter"].split("_")[1:])

    def predict_patches(self, model_key, x, chunk_size, out_arr, min_max=None):
        """surface_segmenter.py:246-294: returns uint8 labels with the shape of x."""
        return super().predict_patches(model_key, x, chunk_size, out_arr, min_max=min_max, TIMEIT=False)

    def segment_volume(self, vol, p_grid, chunk_size, min_max=None, out=None):
        """Fused extract -> predict -> fill of a whole volume (see process_volume)."""
        return self.process_volume(vol, p_grid, chunk_size, min_max=min_max, model_key="segmenter", out=out)

    # ------------------------------------------------------------------ training (surface_segmenter.py:43-124, 202-244)
    def train(self, Xs, Ys, input_size, max_stride, batch_size, cutoff, random_rotate, add_noise, n_epochs,
              steps_per_epoch=None, mode="bf16", verbose=1):
        """surface_segmenter.py:43-66.  Xs / Ys: lists of volumes (numpy or torch, host or CUDA) and
        their ground-truth label volumes.  Keras' ``fit(x=generator, epochs, steps_per_epoch)`` is
        replaced by UnetTrainer fit steps (Adam defaults, BinaryCrossentropy, BN in training mode,
        :322-323); the reference's validation_steps are a no-op there (no validation_data) and here.
        The trained weights (and BN moving statistics) are written back to models["segmenter"]."""
        from ._train import UnetTrainer
        model = self.models["segmenter"]
        dg = self._data_generator(Xs, Ys, batch_size, input_size, max_stride, cutoff, random_rotate, add_noise)
        tot_steps, val_split = 1000, 0.2
        if steps_per_epoch is None:
            steps_per_epoch = int((1 - val_split) * tot_steps // batch_size)
        trainer = getattr(self, "_trainer", None)
        if trainer is None or trainer.patch != tuple(input_size) or trainer.batch != batch_size:
            trainer = UnetTrainer(model.model_params, input_size, batch_size, mode=mode, weights=model.get_weights())
            self._trainer = trainer
        else:
            trainer.set_weights(model.get_weights())
        t0 = time.time()
        history = []
        for epoch in range(n_epochs):
            losses = []
            for _ in range(steps_per_epoch):
                x, y = next(dg)
                losses.append(trainer.train_step(x, y))
            history.append(float(np.mean(losses)))
            if verbose:
                print("Epoch %d/%d - %d steps - loss: %.4f" % (epoch + 1, n_epochs, steps_per_epoch, history[-1]))
        model.set_weights(trainer.get_weights())
        if verbose:
            print("training time = %.2f seconds" % (time.time() - t0))
        return history

    def _find_blanks(self, patches, cutoff, Y_gt, input_size):
        """surface_segmenter.py:75-82: keep patches whose label std exceeds cutoff x the largest std.  One reduction
        kernel (te_patch_std) over the binned patches; the patch array itself is never materialised."""
        assert tuple(Y_gt.shape) == tuple(patches.vol_shape), "volume of Y_gt does not match vol_shape"
        ystd = _patch_std(Y_gt, patches, input_size)
        return (ystd > ystd.max() * cutoff).cpu().numpy().astype(bool)

    def get_training_patches(self, vol_shape, batch_size, input_size, max_stride, cutoff, Y_gt):
        """surface_segmenter.py:85-124: random multi-scale patches ('random': width = stride x input
        size, stride in [1, max_stride)), rejected when blank, until batch_size survive."""
        ip, tot_len, patches = 0, 0, None
        while tot_len < batch_size:
            assert ip <= MAX_ITERS, "stuck in loop while finding patches to train on"
            ip += 1
            p_tmp = Patches(vol_shape, initialize_by="random", min_patch_size=input_size, max_stride=max_stride,
                            n_points=batch_size)
            if cutoff > 0.0:
                cond_list = self._find_blanks(p_tmp, cutoff, Y_gt, input_size)
            else:
                cond_list = np.ones(len(p_tmp), dtype=bool)
            if np.sum(cond_list) > 0:
                p_tmp = p_tmp.filter_by_condition(cond_list)
                if patches is None:
                    patches = p_tmp.copy()
                else:
                    patches.append(p_tmp)
                tot_len = len(patches)        # (the reference only counts after the second round: one extra loop)
        assert patches is not None, "get_patches() failed to return any patches with selected conditions"
        return patches.select_random_sample(batch_size)

    def extract_training_patch_pairs(self, X, Y, patches, add_noise, random_rotate, input_size, seed=None):
        """keras_processor.py:213-232 in one kernel (te_training_pairs): binned gather of x and y, per-patch Gaussian
        noise with sigma ~ U(0, add_noise), random rot90 about a random ordered axis pair.  The per-patch scalars
        (sigma, k, axes) are drawn on the host with the reference's own np.random calls, in its order; the noise
        field comes from the kernel's counter-based generator (seed: drawn from np.random when not given)."""
        batch_size = len(patches)
        patches._calc_binning(tuple(input_size))
        std_batch = np.random.uniform(0, add_noise, batch_size).astype(np.float32)
        nrots = axes = None
        if random_rotate:
            nrots = np.random.randint(0, 4, batch_size).astype(np.int32)
            axes = np.stack([np.random.choice([0, 1, 2], size=2, replace=False) for _ in range(batch_size)]).astype(np.int32)
        if seed is None:
            seed = int(np.random.randint(0, 2 ** 31 - 1))
        return _training_pairs(X, Y, patches, tuple(input_size), std_batch, nrots, axes, seed)

    def _data_generator(self, Xs, Ys, batch_size, input_size, max_stride, cutoff, random_rotate, add_noise):
        """surface_segmenter.py:202-244: yields (x float32 (B,P,P,P,1), y uint8) CUDA tensors."""
        Xs = [_device.to_device(X) for X in Xs]
        Ys = [_device.to_device(Y) for Y in Ys]
        while True:
            n_vols = len(Xs)
            idx_vols = np.repeat(np.arange(0, n_vols), int(np.ceil(batch_size / n_vols)))[:batch_size]
            xy = []
            for ivol in range(n_vols):
                Y_gt = Ys[ivol] if len(Ys) > 1 else Ys[0]
                sub_batch_size = int(np.sum(idx_vols == ivol))
                if sub_batch_size < 1:
                    continue
                patches = self.get_training_patches(tuple(Xs[ivol].shape), sub_batch_size, input_size, max_stride, cutoff, Y_gt)
                xy.append(self.extract_training_patch_pairs(Xs[ivol], Y_gt, patches, add_noise, random_rotate, input_size))
            yield torch.cat([a for a, _ in xy], dim=0), torch.cat([b for _, b in xy], dim=0)

    data_generator = _data_generator          # the reference's `train` calls data_generator (quirk Q3)

    def random_data_generator(self, batch_size, input_size=(64, 64, 64)):
        """surface_segmenter.py:326-334."""
        while True:
            x_shape = tuple([batch_size] + list(input_size) + [1])
            x = np.random.uniform(0, 1, x_shape)
            y = np.random.randint(0, 2, x_shape)
            x[x == 0] = 1.0e-12
            yield x, y
