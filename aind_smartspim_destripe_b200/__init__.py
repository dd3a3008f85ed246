"""B200-native engine for the plane-filter hot path of aind-smartspim-destripe.

``filtering`` and ``zarr_destriper`` mirror the reference modules of the same names
(/root/reference/code/aind_smartspim_destripe/); ``engine`` is the ctypes binding of the
CUDA library (C-ABI: include/dstr_b200.h).  No CPU fallback exists.
"""

__version__ = "0.2.0"

from . import engine, filtering, synthetic, zarr_destriper  # noqa: F401
from .filtering import (  # noqa: F401
    filter_planes,
    filter_streaks,
    filter_stripes,
    flatfield_correction,
    get_foreground_background_mean,
    log_space_fft_filtering,
)
from .zarr_destriper import destripe_volume, execute_worker, release_volume_resources, z_slab  # noqa: F401
