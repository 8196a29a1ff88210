from .keras_processor import GenericKerasProcessor, Vox2VoxProcessor_fCNN, EmbeddingLearner, Model
from .surface_segmenter import SurfaceSegmenter
from .enhancers import Enhancer_fCNN
from .autoencoders import SelfSupervisedCAE, build_encoder_r, build_decoder_r
from .Unet3D import build_Unet_3D
from .latent_pca import LatentPCA, fit_PCA, transform_PCA, rescale_z

__all__ = ["GenericKerasProcessor", "Vox2VoxProcessor_fCNN", "EmbeddingLearner", "Model", "SurfaceSegmenter",
           "Enhancer_fCNN", "SelfSupervisedCAE", "build_encoder_r", "build_decoder_r", "build_Unet_3D",
           "LatentPCA", "fit_PCA", "transform_PCA", "rescale_z"]
