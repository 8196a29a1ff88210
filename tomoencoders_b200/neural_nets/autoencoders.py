"""``SelfSupervisedCAE``: drop-in for /root/reference/tomo_encoders/neural_nets/autoencoders.py:529-777
on the inference path: conv encoder -> latent vectors (``predict_embeddings``) and
encoder + decoder reconstruction (``predict_patches("autoencoder", ...)``)."""
import os
import time

import numpy as np
import torch

from .. import _device, _lib
from ..structures.patches import Patches
from .keras_processor import _ChunkGraph, EmbeddingLearner, Model, gather_normalized, normalize_chunk


def build_encoder_r(input_shape, **model_params):
    """autoencoders.py:246-347.  Returns (encoder, flatten_shape, preflatten_shape)."""
    from . import _spec
    mode = model_params.pop("mode", None)
    model_size = tuple(input_shape[:3])
    m = Model("encoder", model_params, model_size=model_size, mode=mode, name="encoder")
    p = m.model_params
    flat, pre = _spec.encoder_shapes(model_size, p["n_filters"], p["n_blocks"], p.get("pool_size", 2),
                                     p.get("POOL_FLAG", True))
    return m, flat, pre


def build_decoder_r(flatten_shape, preflatten_shape, model_size=None, **model_params):
    """autoencoders.py:350-447 (model_size is needed here because the plan is sized per input)."""
    mode = model_params.pop("mode", None)
    return Model("decoder", model_params, model_size=model_size, mode=mode, name="decoder")


class SelfSupervisedCAE(EmbeddingLearner):
    output_type = "data"

    def __init__(self, **kwargs):
        self.model_keys = ["encoder", "decoder", "autoencoder"]
        super().__init__(**kwargs)

    def _build_models(self, model_size=(64, 64, 64), descriptor_tag="misc", **model_params):
        """autoencoders.py:551-577.  ``stdinput=True`` (build_CAE_3D of the ICIP porosity encoder,
        scratchpad/IEEE_ICIP_paper/porosity_encoders.py:246-302; its encoder is build_encoder_r with POOL_FLAG=False)
        standardises every input patch to its own [0, 1] range in front of the encoder, fused into the gather."""
        if model_params is None:
            raise ValueError("Need model hyperparameters or instance of model. Neither were provided")
        self.stdinput = bool(model_params.pop("stdinput", False))
        self.models = {}
        self.model_size = tuple(int(s) for s in model_size)
        self.model_tag = "%s" % descriptor_tag
        enc, flat, pre = build_encoder_r(self.model_size + (1,), mode=self.mode, **model_params)
        self.models["encoder"] = enc
        self.models["decoder"] = build_decoder_r(flat, pre, model_size=self.model_size, mode=self.mode, **model_params)
        self.models["autoencoder"] = None

    def _load_models(self, model_tag=None, model_size=(64, 64, 64), model_path="some-path", **model_params):
        """autoencoders.py:579-602 with ``.npz`` weight files (see SurfaceSegmenter._load_models)."""
        assert model_tag is not None, "need model_tag"
        self._build_models(model_size=model_size, descriptor_tag=model_tag, **model_params)
        for model_key in ("encoder", "decoder"):
            self.models[model_key].load_weights(os.path.join(model_path, "%s_%s.npz" % (model_key, model_tag)))

    def get_patches(self, vol_shape, sampling_method, batch_size, max_stride=None):
        """autoencoders.py:657-672."""
        if sampling_method in ["grid", "regular-grid", "random-fixed-width"]:
            return Patches(vol_shape, initialize_by=sampling_method, patch_size=self.model_size, n_points=batch_size)
        if sampling_method in ["random"]:
            return Patches(vol_shape, initialize_by=sampling_method, min_patch_size=self.model_size,
                           max_stride=max_stride, n_points=batch_size)
        raise ValueError("sampling method not supported")

    # ------------------------------------------------------------------ training (autoencoders.py:604-718)
    def extract_training_sub_volumes(self, X, patches, add_noise, random_rotate):
        """autoencoders.py:640-655 on the device (add_noise is accepted and, as in the reference, unused)."""
        batch_size = len(patches)
        x = _device.to_device(patches.extract(_device.to_device(X), self.model_size)).float()
        if random_rotate:
            nrots = np.random.randint(0, 4, batch_size)
            for ii in range(batch_size):
                axes = tuple(int(a) for a in np.random.choice([0, 1, 2], size=2, replace=False))
                if nrots[ii]:
                    x[ii] = torch.rot90(x[ii], int(nrots[ii]), axes)
        return x.unsqueeze(-1)

    def data_generator(self, Xs, batch_size, sampling_method, max_stride=1, random_rotate=False, add_noise=0.1):
        """autoencoders.py:604-638: yields (B,P,P,P,1) float32 CUDA tensors."""
        Xs = [_device.to_device(X) for X in Xs]
        while True:
            n_vols = len(Xs)
            idx_vols = np.repeat(np.arange(0, n_vols), int(np.ceil(batch_size / n_vols)))[:batch_size]
            x = []
            for ivol in range(n_vols):
                nsub = int(np.sum(idx_vols == ivol))
                if nsub < 1:
                    continue
                patches = self.get_patches(tuple(Xs[ivol].shape), sampling_method, nsub, max_stride=max_stride)
                x.append(self.extract_training_sub_volumes(Xs[ivol], patches, add_noise, random_rotate))
            yield torch.cat(x, dim=0)

    def train(self, vols, batch_size=10, sampling_method="random-fixed-width", n_epochs=10, random_rotate=True,
              add_noise=0.1, max_stride=1, normalize_sampling_factor=2, steps_per_epoch=None, mode="bf16", verbose=1):
        """autoencoders.py:674-718: RegularizedAutoencoder(encoder, decoder, weight 1/250, 'kl') compiled with Adam and
        fitted on generated batches; here CAETrainer fit steps.  The volumes are expected in the network's input range
        [0, 1] (the reference leaves normalisation to its data loader, :683).  The trained weights are written back to
        models["encoder"] / models["decoder"]; returns the per-epoch mean of the total loss."""
        from ._train import CAETrainer
        enc, dec = self.models["encoder"], self.models["decoder"]
        dg = self.data_generator(vols, batch_size, sampling_method, max_stride=max_stride, random_rotate=random_rotate,
                                 add_noise=add_noise)
        tot_steps, val_split = 500, 0.2
        if steps_per_epoch is None:
            steps_per_epoch = int((1 - val_split) * tot_steps // batch_size)
        trainer = getattr(self, "_trainer", None)
        if trainer is None or trainer.batch != batch_size:
            trainer = CAETrainer(enc.model_params, self.model_size, batch_size, mode=mode,
                                 weights=enc.get_weights() + dec.get_weights())
            self._trainer = trainer
        else:
            trainer.set_weights(enc.get_weights() + dec.get_weights())
        t0 = time.time()
        history = []
        for epoch in range(n_epochs):
            losses = [trainer.train_step(next(dg)) for _ in range(steps_per_epoch)]
            history.append(float(np.mean(losses)))
            if verbose:
                print("Epoch %d/%d - %d steps - loss: %.5f" % (epoch + 1, n_epochs, steps_per_epoch, history[-1]))
        we, wd = trainer.get_model_weights()
        enc.set_weights(we)
        dec.set_weights(wd)
        self.models["autoencoder"] = "trained"       # the reference stores the fitted RegularizedAutoencoder here
        if verbose:
            print("training time = %.2f seconds" % (time.time() - t0))
        return history

    def predict_embeddings(self, x, chunk_size, min_max=None, TIMEIT=False):
        """autoencoders.py:729-777: x (nb,pz,py,px,1) -> (nb, latent) float32 (rescale, no clip)."""
        assert x.ndim == 5, "x must be 5-dimensional (batch_size, nz, ny, nx, 1)."
        t0 = time.time()
        nb = len(x)
        on_device = _device.is_device_array(x)
        enc = self.models["encoder"]
        patch = tuple(int(s) for s in x.shape[1:4])
        out = torch.empty((nb, enc.latent), dtype=torch.float32, device=_device.device())
        for a in range(0, nb, chunk_size):
            b = min(a + chunk_size, nb)
            xc = x[a:b, ..., 0]
            xc = _device.to_device(xc if on_device else np.ascontiguousarray(xc))
            net = enc.net_for(patch, max(chunk_size, b - a))
            net.encoder_forward(normalize_chunk(xc, min_max, False, net, standardize=self.stdinput), out=out[a:b])
        res = out if on_device else out.cpu().numpy()
        t_unit = (time.time() - t0) * 1000.0 / max(nb, 1)
        if TIMEIT:
            print("inf. time p. input patch size %s = %.2f ms, nb = %i" % (str(patch), t_unit, nb))
            print("\n")
            return res, t_unit
        return res

    def embed_volume(self, vol, patches, chunk_size, min_max=None, out=None):
        """Fused patches.extract(vol, model_size) -> predict_embeddings: the (n,P,P,P) patch array is
        never materialised; each chunk is gathered (with binning) and rescaled straight into the
        encoder's input layout.  Returns (n, latent) float32 where vol lives (host / device)."""
        return self._guard_fp16("encoder", lambda: self._embed_volume(vol, patches, chunk_size, min_max, out))

    def _embed_volume(self, vol, patches, chunk_size, min_max, out):
        vs = tuple(int(v) for v in patches.vol_shape)
        assert tuple(vol.shape) == vs, "Shape of big volume does not match vol_shape attribute of patches data"
        widths = patches.widths if hasattr(patches, "widths") else patches._widths()
        if hasattr(patches, "_calc_binning"):
            patches._calc_binning(self.model_size)
        n = len(patches)
        host = not _device.is_device_array(vol)
        t_vol = _device.to_supported(_device.to_device(vol))
        enc = self.models["encoder"]
        z = out if (out is not None and _device.is_device_array(out)) else \
            torch.empty((n, enc.latent), dtype=torch.float32, device=t_vol.device)
        if n:
            net = enc.net_for(self.model_size, min(chunk_size, n))
            d_pts = _device.points_to_device(patches.points)
            d_wds = _device.points_to_device(widths)
            L = _lib.lib()
            lo, hi = (0.0, 1.0) if min_max is None else (float(min_max[0]), float(min_max[1]))
            vs64, p32 = _lib.i64x3(vs), _lib.i32x3(self.model_size)
            msize = tuple(int(v) for v in self.model_size)

            def chunk_body(pts_t, wds_t, nn, xa, zc):
                if self.stdinput:
                    xa = gather_normalized(t_vol, vs, pts_t, wds_t, nn, msize, min_max, False, net, standardize=True)
                else:
                    _lib.check(L.te_gather_norm(t_vol.data_ptr(), _device.te_dtype(t_vol), vs64, pts_t.data_ptr(), wds_t.data_ptr(),
                                                nn, p32, 0 if min_max is None else 1, 0, lo, hi, xa.data_ptr(), net.act_te,
                                                _device.stream_ptr()), "te_gather_norm")
                net.encoder_forward(xa, out=zc)

            starts = list(range(0, n, chunk_size))
            # one captured graph per chunk length (gather + rescale -> encoder), replayed with the chunk's corner table
            graphs = self.use_graphs and not self.stdinput and not net.profiling and len(starts) > 1
            for a in starts:
                b = min(a + chunk_size, n)
                if graphs:
                    key = ("enc", b - a, t_vol.data_ptr(), str(t_vol.dtype), vs, msize, min_max is None, lo, hi)
                    g = self._chunk_graph(net, key, lambda: _ChunkGraph(b - a, msize, net.act_dtype, torch.float32, t_vol.device,
                                                                       out_shape=(b - a, enc.latent)))
                    g.varying_widths = True
                    g.run(d_pts[a:b], d_wds[a:b], chunk_body)
                    z[a:b].copy_(g.y)
                else:
                    xa = torch.empty((b - a,) + msize, dtype=net.act_dtype, device=t_vol.device)
                    chunk_body(d_pts[a:b], d_wds[a:b], b - a, xa, z[a:b])
        if out is not None and not _device.is_device_array(out):
            out[...] = z.cpu().numpy()
            return out
        return z.cpu().numpy() if (host and out is None) else z
