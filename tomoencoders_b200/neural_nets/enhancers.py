"""``Enhancer_fCNN``: drop-in for /root/reference/tomo_encoders/neural_nets/enhancers.py:30-111 on
the inference path: the same U-net as the segmenter with a float ("data") output -- the denoiser."""
import os

from .keras_processor import Vox2VoxProcessor_fCNN
from .Unet3D import build_Unet_3D


class Enhancer_fCNN(Vox2VoxProcessor_fCNN):
    output_type = "data"

    def _build_models(self, descriptor_tag="misc", **model_params):
        """enhancers.py:84-111."""
        if model_params is None:
            raise ValueError("Need model hyperparameters or instance of model. Neither were provided")
        self.models = {}
        self.model_tag = "Unet_%s" % (descriptor_tag)
        self.models["enhancer"] = build_Unet_3D(mode=self.mode, **model_params)

    def _load_models(self, model_names=None, model_path="some/path", **model_params):
        self.models = {}
        for model_key, model_name in model_names.items():
            m = build_Unet_3D(mode=self.mode, **model_params)
            m.load_weights(os.path.join(model_path, model_name + ".npz"))
            self.models[model_key] = m
        self.model_tag = "_".join(model_names["enhancer"].split("_")[1:])

    def test_speeds(self, chunk_size, n_reps=3, input_size=None, model_key="enhancer"):
        return super().test_speeds(chunk_size, n_reps=n_reps, input_size=input_size, model_key=model_key)

    def enhance_volume(self, vol, patches, chunk_size, min_max=None, out=None):
        return self.process_volume(vol, patches, chunk_size, min_max=min_max, model_key="enhancer", out=out)
