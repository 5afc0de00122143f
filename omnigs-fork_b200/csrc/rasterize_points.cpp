// rasterize_points.cpp — drop-in replacement for the reference's src/rasterize_points.cu.
//
// Exports the same three C++ symbols with the same LibTorch signatures, return tuples and error
// behaviour (reference src/rasterize_points.cu:49-319); everything below the tensors goes through
// the C ABI of libomnigs_b200.so.  Differences that are invisible to the caller:
//   * the three byte buffers are sized with ogs_*_bytes and allocated once (the reference grows
//     them through resize_ callbacks, :41-47); their layout is ours;
//   * work is launched on torch's current CUDA stream (the reference uses the legacy default one);
//   * gradients are allocated with empty(): the library writes every element, so the reference's
//     324 B/Gaussian of zero-fill (:200-208,246-247) disappears;
//   * camera_type == 1 (pinhole, with render_depth) and 3 (lonlat) both run through the same library kernels.
#include "rasterize_points.h"

#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>

#include <stdexcept>
#include <string>

#include "../../include/omnigs_b200.h"

namespace {

constexpr int kNumChannels = 3; // reference cuda_rasterizer/config.h:25

// An empty tensor is the reference's "None": its data pointer is null (gaussian_rasterizer.cpp:198-208).
const float* fptr(const torch::Tensor& t) { return t.numel() == 0 ? nullptr : t.data_ptr<float>(); }
char* bptr(const torch::Tensor& t) { return t.numel() == 0 ? nullptr : reinterpret_cast<char*>(t.data_ptr()); }

void check(int status)
{
	if (status != OGS_OK)
		throw std::runtime_error(std::string("[omnigs_b200] error ") + std::to_string(status) + ": " + ogs_last_error());
}

void check_camera(int camera_type)
{
	if (camera_type != 1 && camera_type != 3) throw std::runtime_error("[CudaRasterizer]Invalid camera_type");
}

} // namespace

std::tuple<int, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor>
RasterizeGaussiansCUDA(
	const torch::Tensor& background, const torch::Tensor& means3D, const torch::Tensor& colors,
	const torch::Tensor& opacity, const torch::Tensor& scales, const torch::Tensor& rotations,
	const float scale_modifier, const torch::Tensor& cov3D_precomp, const torch::Tensor& viewmatrix,
	const torch::Tensor& projmatrix, const float tan_fovx, const float tan_fovy,
	const int image_height, const int image_width, const torch::Tensor& sh, const int degree,
	const torch::Tensor& campos, const bool prefiltered, const int camera_type, const bool render_depth)
{
	(void)prefiltered;
	if (means3D.ndimension() != 2 || means3D.size(1) != 3) {
		AT_ERROR("means3D must have dimensions (num_points, 3)");
	}
	const int P = means3D.size(0);
	const int H = image_height;
	const int W = image_width;

	auto float_opts = means3D.options().dtype(torch::kFloat32);
	auto byte_opts = means3D.options().dtype(torch::kUInt8);
	torch::Tensor radii = torch::empty({P}, means3D.options().dtype(torch::kInt32));
	torch::Tensor geomBuffer = torch::empty({0}, byte_opts);
	torch::Tensor binningBuffer = torch::empty({0}, byte_opts);
	torch::Tensor imgBuffer = torch::empty({0}, byte_opts);

	int rendered = 0;
	if (P == 0) // the reference launches nothing and returns its zero-filled outputs (:84-85,97)
		return std::make_tuple(rendered, torch::zeros({kNumChannels, H, W}, float_opts), radii, geomBuffer, binningBuffer, imgBuffer);
	check_camera(camera_type);

	c10::cuda::CUDAGuard guard(means3D.device());
	cudaStream_t stream = at::cuda::getCurrentCUDAStream();
	int M = 0;
	if (sh.size(0) != 0) M = sh.size(1);

	// keep the contiguous views alive for the duration of the launches
	const torch::Tensor bg = background.contiguous(), m3 = means3D.contiguous(), col = colors.contiguous(),
	                    op = opacity.contiguous(), sc = scales.contiguous(), rot = rotations.contiguous(),
	                    cov = cov3D_precomp.contiguous(), vm = viewmatrix.contiguous(), shc = sh.contiguous(),
	                    cam = campos.contiguous(), pm = projmatrix.contiguous();

	torch::Tensor out_color = torch::empty({kNumChannels, H, W}, float_opts);
	geomBuffer = torch::empty({(int64_t)ogs_geom_bytes(P)}, byte_opts);
	imgBuffer = torch::empty({(int64_t)ogs_img_bytes(W, H)}, byte_opts);
	int64_t num_rendered = 0;
	if (camera_type == 1) // PINHOLE (rasterize_points.cu:105-132)
		check(ogs_pinhole_forward_stage1(P, degree, M, W, H, fptr(m3), fptr(shc), fptr(col), fptr(op), fptr(sc),
		                                 scale_modifier, fptr(rot), fptr(cov), fptr(vm), fptr(pm), fptr(cam), tan_fovx,
		                                 tan_fovy, render_depth ? 1 : 0, radii.data_ptr<int>(), bptr(geomBuffer),
		                                 bptr(imgBuffer), &num_rendered, stream));
	else
		check(ogs_lonlat_forward_stage1(P, degree, M, W, H, fptr(m3), fptr(shc), fptr(col), fptr(op), fptr(sc),
		                                scale_modifier, fptr(rot), fptr(cov), fptr(vm), fptr(cam),
		                                radii.data_ptr<int>(), bptr(geomBuffer), bptr(imgBuffer), &num_rendered, stream));
	binningBuffer = torch::empty({(int64_t)ogs_binning_bytes(num_rendered, W, H)}, byte_opts);
	check(ogs_lonlat_forward_stage2(P, W, H, num_rendered, fptr(bg), bptr(geomBuffer), bptr(binningBuffer),
	                                bptr(imgBuffer), out_color.data_ptr<float>(), stream));
	rendered = (int)num_rendered;
	return std::make_tuple(rendered, out_color, radii, geomBuffer, binningBuffer, imgBuffer);
}

std::tuple<torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor>
RasterizeGaussiansBackwardCUDA(
	const torch::Tensor& background, const torch::Tensor& means3D, const torch::Tensor& radii,
	const torch::Tensor& colors, const torch::Tensor& scales, const torch::Tensor& rotations,
	const float scale_modifier, const torch::Tensor& cov3D_precomp, const torch::Tensor& viewmatrix,
	const torch::Tensor& projmatrix, const float tan_fovx, const float tan_fovy,
	const torch::Tensor& dL_dout_color, const torch::Tensor& sh, const int degree, const torch::Tensor& campos,
	const torch::Tensor& geomBuffer, const int R, const torch::Tensor& binningBuffer,
	const torch::Tensor& imageBuffer, const int camera_type)
{
	const int P = means3D.size(0);
	const int H = dL_dout_color.size(1);
	const int W = dL_dout_color.size(2);
	int M = 0;
	if (sh.size(0) != 0) M = sh.size(1);

	auto opts = means3D.options();
	auto alloc = [&](std::initializer_list<int64_t> shape) {
		return P == 0 ? torch::zeros(shape, opts) : torch::empty(shape, opts);
	};
	torch::Tensor dL_dmeans3D = alloc({P, 3});
	torch::Tensor dL_dmeans2D = alloc({P, 3});
	torch::Tensor dL_dcolors = alloc({P, kNumChannels});
	torch::Tensor dL_dopacity = alloc({P, 1});
	torch::Tensor dL_dcov3D = alloc({P, 6});
	torch::Tensor dL_dsh = alloc({P, M, 3});
	torch::Tensor dL_dscales = alloc({P, 3});
	torch::Tensor dL_drotations = alloc({P, 4});

	if (P != 0) {
		check_camera(camera_type);
		c10::cuda::CUDAGuard guard(means3D.device());
		cudaStream_t stream = at::cuda::getCurrentCUDAStream();
		const torch::Tensor pm = projmatrix.contiguous();
		const torch::Tensor bg = background.contiguous(), m3 = means3D.contiguous(), col = colors.contiguous(),
		                    sc = scales.contiguous(), rot = rotations.contiguous(), cov = cov3D_precomp.contiguous(),
		                    vm = viewmatrix.contiguous(), shc = sh.contiguous(), cam = campos.contiguous(),
		                    dL = dL_dout_color.contiguous(), rad = radii.contiguous(), gb = geomBuffer.contiguous(),
		                    bb = binningBuffer.contiguous(), ib = imageBuffer.contiguous();
		if (camera_type == 1) // PINHOLE (rasterize_points.cu:212-243)
			check(ogs_pinhole_backward(P, degree, M, R, W, H, fptr(bg), fptr(m3), fptr(shc), fptr(col), fptr(sc),
			                           scale_modifier, fptr(rot), fptr(cov), fptr(vm), fptr(pm), fptr(cam), tan_fovx, tan_fovy,
			                           rad.data_ptr<int>(), bptr(gb), bptr(bb), bptr(ib), fptr(dL),
			                           dL_dmeans2D.data_ptr<float>(), nullptr, dL_dopacity.data_ptr<float>(),
			                           dL_dcolors.data_ptr<float>(), dL_dmeans3D.data_ptr<float>(), dL_dcov3D.data_ptr<float>(),
			                           M ? dL_dsh.data_ptr<float>() : nullptr, dL_dscales.data_ptr<float>(),
			                           dL_drotations.data_ptr<float>(), stream));
		else
			check(ogs_lonlat_backward(P, degree, M, R, W, H, fptr(bg), fptr(m3), fptr(shc), fptr(col), fptr(sc),
			                          scale_modifier, fptr(rot), fptr(cov), fptr(vm), fptr(cam), rad.data_ptr<int>(),
			                          bptr(gb), bptr(bb), bptr(ib), fptr(dL),
			                          dL_dmeans2D.data_ptr<float>(), nullptr, dL_dopacity.data_ptr<float>(),
			                          dL_dcolors.data_ptr<float>(), dL_dmeans3D.data_ptr<float>(), dL_dcov3D.data_ptr<float>(),
			                          M ? dL_dsh.data_ptr<float>() : nullptr, dL_dscales.data_ptr<float>(),
			                          dL_drotations.data_ptr<float>(), stream));
	}
	return std::make_tuple(dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscales, dL_drotations);
}

torch::Tensor markVisible(torch::Tensor& means3D, torch::Tensor& viewmatrix, torch::Tensor& projmatrix, const int camera_type)
{
	const int P = means3D.size(0);
	torch::Tensor present = torch::full({P}, false, means3D.options().dtype(at::kBool));
	if (P != 0) {
		check_camera(camera_type);
		c10::cuda::CUDAGuard guard(means3D.device());
		uint8_t* out = reinterpret_cast<uint8_t*>(present.data_ptr<bool>());
		if (camera_type == 1) { // checkFrustum (rasterize_points.cu:299-306)
			const torch::Tensor m3 = means3D.contiguous(), vm = viewmatrix.contiguous(), pm = projmatrix.contiguous();
			check(ogs_mark_visible_pinhole(P, fptr(m3), fptr(vm), fptr(pm), out, at::cuda::getCurrentCUDAStream()));
		} else {
			check(ogs_mark_all_visible(P, out, at::cuda::getCurrentCUDAStream()));
		}
	}
	return present;
}
