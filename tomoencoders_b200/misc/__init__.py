from . import voxel_processing

__all__ = ["voxel_processing"]
