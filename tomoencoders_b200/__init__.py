"""tomoencoders_b200 -- B200-native (sm_100a) implementation of the TomoEncoders hot path:
patch gather -> 3-D conv encoder / U-net segmenter -> latent PCA -> patch scatter / stitch,
behind the reference's Python API (tomo_encoders/__init__.py:8-11 exports Patches, Grid and DataFile).
"""
from .structures.patches import Patches
from .structures.grid import Grid
from .structures.datafile import DataFile
from .misc import voxel_processing

__all__ = ["Patches", "Grid", "DataFile", "voxel_processing"]
__version__ = "0.1.0"
