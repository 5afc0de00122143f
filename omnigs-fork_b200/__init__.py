"""omnigs-fork_b200 — B200-native (sm_100a) equirectangular Gaussian-splatting rasterizer.

Host-side mirror of the reference's operator interface for the lonlat hot path
(raikuma/OmniGS-fork: include/rasterize_points.h:29-80, src/gaussian_rasterizer.cpp:34-170) on top of
the C ABI in include/omnigs_b200.h (libomnigs_b200.so, hand-written CUDA).  PyTorch is used only for
device memory, streams and torch.distributed.

The directory name contains a hyphen (it follows the reference repo's name), so import it with
    importlib.import_module("omnigs-fork_b200")
tests/conftest.py and bench.py also register it as ``omnigs_fork_b200``.
"""
from ._lib import load_library, library_path, OgsError  # noqa: F401
from .rasterize_points import (  # noqa: F401
    RasterizeGaussiansCUDA,
    RasterizeGaussiansBackwardCUDA,
    RasterizeGaussiansBackwardView,
    RasterizeGaussiansGeometry,
    RasterizeGaussiansBlend,
    markVisible,
    export_forward_state,
    set_seam_wrap,
    LONLAT,
    PINHOLE,
)
from .rasterizer import GaussianRasterizationSettings, GaussianRasterizer, rasterize_gaussians  # noqa: F401

__all__ = [
    "RasterizeGaussiansCUDA", "RasterizeGaussiansBackwardCUDA", "RasterizeGaussiansBackwardView",
    "RasterizeGaussiansGeometry", "RasterizeGaussiansBlend", "markVisible", "export_forward_state", "set_seam_wrap",
    "GaussianRasterizationSettings", "GaussianRasterizer", "rasterize_gaussians",
    "load_library", "library_path", "OgsError", "LONLAT", "PINHOLE",
]
