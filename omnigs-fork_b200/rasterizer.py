"""The caller side of the boundary: autograd wrapper with the reference's semantics.

Mirrors ``GaussianRasterizerFunction`` / ``GaussianRasterizer`` / ``GaussianRasterizationSettings``
(reference src/gaussian_rasterizer.cpp:34-221, include/gaussian_rasterizer.h:41-141): which tensors
are saved, how absent arguments are passed (empty tensors) and the order in which the backward
8-tuple is mapped onto the forward inputs.  In the reference this layer is C++ and stays unchanged
(it links against the drop-in symbols of csrc/rasterize_points.cpp); this Python twin exists so the
tests and the data-parallel trainer can drive the path through autograd.
"""
from dataclasses import dataclass

import torch

from .rasterize_points import RasterizeGaussiansCUDA, RasterizeGaussiansBackwardCUDA, markVisible, LONLAT


@dataclass
class GaussianRasterizationSettings:
    """reference include/gaussian_rasterizer.h:41-68"""
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool = False
    camera_type: int = LONLAT
    render_depth: bool = False


class _RasterizeGaussians(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, rs):
        # reference gaussian_rasterizer.cpp:34-102
        num_rendered, color, radii, geomBuffer, binningBuffer, imgBuffer = RasterizeGaussiansCUDA(
            rs.bg, means3D, colors_precomp, opacities, scales, rotations, rs.scale_modifier, cov3Ds_precomp,
            rs.viewmatrix, rs.projmatrix, rs.tanfovx, rs.tanfovy, rs.image_height, rs.image_width,
            sh, rs.sh_degree, rs.campos, rs.prefiltered, rs.camera_type, rs.render_depth)
        ctx.rs = rs
        ctx.num_rendered = num_rendered
        ctx.save_for_backward(colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, sh,
                              geomBuffer, binningBuffer, imgBuffer)
        ctx.mark_non_differentiable(radii)
        return color, radii

    @staticmethod
    def backward(ctx, grad_out_color, _grad_radii):
        # reference gaussian_rasterizer.cpp:104-170
        rs = ctx.rs
        colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, sh, geomBuffer, binningBuffer, imgBuffer = ctx.saved_tensors
        (dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscales,
         dL_drotations) = RasterizeGaussiansBackwardCUDA(
            rs.bg, means3D, radii, colors_precomp, scales, rotations, rs.scale_modifier, cov3Ds_precomp,
            rs.viewmatrix, rs.projmatrix, rs.tanfovx, rs.tanfovy, grad_out_color.contiguous(), sh, rs.sh_degree,
            rs.campos, geomBuffer, ctx.num_rendered, binningBuffer, imgBuffer, rs.camera_type)

        def opt(g, ref):  # absent inputs (empty tensors) get no gradient
            return g if ref.numel() else None
        return (dL_dmeans3D, dL_dmeans2D, opt(dL_dsh, sh), opt(dL_dcolors, colors_precomp), dL_dopacity,
                opt(dL_dscales, scales), opt(dL_drotations, rotations), opt(dL_dcov3D, cov3Ds_precomp), None)


def rasterize_gaussians(means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                        raster_settings):
    """reference include/gaussian_rasterizer.h:91-112"""
    return _RasterizeGaussians.apply(means3D, means2D, sh, colors_precomp, opacities, scales, rotations,
                                     cov3Ds_precomp, raster_settings)


class GaussianRasterizer(torch.nn.Module):
    """reference include/gaussian_rasterizer.h:114-141, src/gaussian_rasterizer.cpp:173-221"""

    def __init__(self, raster_settings):
        super().__init__()
        self.raster_settings = raster_settings

    def markVisibleGaussians(self, positions):
        rs = self.raster_settings
        with torch.no_grad():
            return markVisible(positions, rs.viewmatrix, rs.projmatrix, rs.camera_type)

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None):
        def has(t):
            return t is not None and t.numel() > 0
        # reference gaussian_rasterizer.cpp:190-196
        if has(shs) == has(colors_precomp):
            raise RuntimeError("Please provide excatly one of either SHs or precomputed colors!")
        if (has(scales) or has(rotations)) == has(cov3D_precomp) or has(scales) != has(rotations):
            raise RuntimeError("Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!")
        empty = torch.empty((0,), dtype=torch.float32, device=means3D.device)  # the reference's "None" (:198-208)
        shs = shs if has(shs) else empty
        colors_precomp = colors_precomp if has(colors_precomp) else empty
        scales = scales if has(scales) else empty
        rotations = rotations if has(rotations) else empty
        cov3D_precomp = cov3D_precomp if has(cov3D_precomp) else empty
        return rasterize_gaussians(means3D, means2D, shs, colors_precomp, opacities, scales, rotations,
                                   cov3D_precomp, self.raster_settings)
