// rasterize_points.h — the LibTorch boundary of the B200 rasterizer.
//
// Declares the three functions the reference declares in include/rasterize_points.h:29-80 with the same names,
// argument order, argument types, defaults and result tuples (the mangled symbols are identical), so that the
// reference's autograd function (src/gaussian_rasterizer.cpp:48,135) and src/operate_points.cu:114 link against this
// implementation unchanged.  Implemented in rasterize_points.cpp on top of the C ABI (include/omnigs_b200.h).
//
// Conventions: "None" is an empty CUDA tensor; exactly one of {sh, colors} and one of
// {(scales, rotations), cov3D_precomp} is non-empty; float32, contiguous; viewmatrix = Tcw^T row-major, campos = the
// camera centre; rotations are (w, x, y, z), already normalised; opacity is post-sigmoid, scales post-exp.
// camera_type 3 = equirectangular ("lonlat"), 1 = pinhole; anything else throws std::runtime_error like the reference.
#pragma once
#include <torch/torch.h>

#include <tuple>

namespace omnigs_b200 {
using Tensor = torch::Tensor;
// (num_rendered, out_color [3,H,W], radii [P] int32, geomBuffer, binningBuffer, imageBuffer — three opaque byte tensors
// whose layout is private to this library and only has to live until the backward of the same frame)
using RasterizeForwardResult = std::tuple<int, Tensor, Tensor, Tensor, Tensor, Tensor>;
// (dL_dmeans2D [P,3], dL_dcolors [P,3], dL_dopacity [P,1], dL_dmeans3D [P,3], dL_dcov3D [P,6], dL_dsh [P,M,3],
//  dL_dscales [P,3], dL_drotations [P,4]) — every element written, zeros for culled Gaussians
using RasterizeBackwardResult = std::tuple<Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor>;
} // namespace omnigs_b200

// Forward.  Ignored for camera_type 3, as in the reference: projmatrix, tan_fovx, tan_fovy, prefiltered, render_depth.
omnigs_b200::RasterizeForwardResult RasterizeGaussiansCUDA(
	const torch::Tensor& background /*[3]*/, const torch::Tensor& means3D /*[P,3]*/,
	const torch::Tensor& colors /*[P,3] or empty*/, const torch::Tensor& opacity /*[P,1]*/,
	const torch::Tensor& scales /*[P,3] or empty*/, const torch::Tensor& rotations /*[P,4] or empty*/,
	const float scale_modifier, const torch::Tensor& cov3D_precomp /*[P,6] or empty*/,
	const torch::Tensor& viewmatrix /*[4,4]*/, const torch::Tensor& projmatrix /*[4,4]*/,
	const float tan_fovx, const float tan_fovy, const int image_height, const int image_width,
	const torch::Tensor& sh /*[P,M,3] or empty*/, const int degree, const torch::Tensor& campos /*[3]*/,
	const bool prefiltered, const int camera_type = 1, const bool render_depth = false);

// Backward of the frame whose forward returned (R, radii, geomBuffer, binningBuffer, imageBuffer).
omnigs_b200::RasterizeBackwardResult RasterizeGaussiansBackwardCUDA(
	const torch::Tensor& background, const torch::Tensor& means3D, const torch::Tensor& radii,
	const torch::Tensor& colors, const torch::Tensor& scales, const torch::Tensor& rotations,
	const float scale_modifier, const torch::Tensor& cov3D_precomp,
	const torch::Tensor& viewmatrix, const torch::Tensor& projmatrix, const float tan_fovx, const float tan_fovy,
	const torch::Tensor& dL_dout_color /*[3,H,W]*/, const torch::Tensor& sh, const int degree,
	const torch::Tensor& campos, const torch::Tensor& geomBuffer, const int R,
	const torch::Tensor& binningBuffer, const torch::Tensor& imageBuffer, const int camera_type = 1);

// bool [P]: everything for camera_type 3 (rasterizer_impl.cu:185-192), checkFrustum for camera_type 1.
torch::Tensor markVisible(torch::Tensor& means3D, torch::Tensor& viewmatrix, torch::Tensor& projmatrix,
                          const int camera_type = 1);
