"""``SurfaceSegmenter``: drop-in for /root/reference/tomo_encoders/neural_nets/surface_segmenter.py:36-350
on the inference path (3-D U-net, clip + rescale + predict + round -> uint8 labels)."""
import os

import numpy as np

from .keras_processor import Vox2VoxProcessor_fCNN
from .Unet3D import build_Unet_3D


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

    def random_data_generator(self, batch_size, input_size=(64, 64, 64)):
        """surface_segmenter.py:326-334."""
        while True:
            x_shape = tuple([batch_size] + list(input_size) + [1])
            x = np.random.uniform(0, 1, x_shape)
            y = np.random.randint(0, 2, x_shape)
            x[x == 0] = 1.0e-12
            yield x, y
