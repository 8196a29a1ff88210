"""Filtered back-projection next to the hot path: drop-in for
/root/reference/tomo_encoders/reconstruction/{cuda_kernels,prep,recon}.py over csrc/recon_ops.cu."""
from . import cuda_kernels, prep, recon  # noqa: F401
from .cuda_kernels import BackProjector, rec_all, rec_mask, rec_patch, rec_pts, rec_pts_xy  # noqa: F401
from .prep import calc_padding, fbp_filter, preprocess  # noqa: F401
from .recon import (extract_from_mask, extract_segmented, make_mask, recon_all, recon_all_gpu,  # noqa: F401
                    recon_patches_3d)
