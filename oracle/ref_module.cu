// Python binding around the UNMODIFIED reference rasterizer (oracle/_ref build).
//
// Test/bench infrastructure only.  This translation unit is ours; it is linked with the
// reference's own sources compiled where they lie under /root/reference
// (cuda_rasterizer/{forward,backward,rasterizer_impl}.cu and src/rasterize_points.cu) by
// oracle/build_ref.sh.  It exposes the reference's three LibTorch entry points
// (include/rasterize_points.h:29-80) and re-uses the reference's own fromChunk carving
// (cuda_rasterizer/rasterizer_impl.cu:198-245) to expose the opaque buffers' contents.
#include <torch/extension.h>
#include "include/rasterize_points.h"
#include "cuda_rasterizer/rasterizer_impl.h"
#include "cuda_rasterizer/rasterizer.h"
#include <vector>

namespace {

torch::Tensor view_bytes(void* p, int64_t nbytes, const torch::Tensor& owner)
{
	// copy out of the opaque buffer into a fresh u8 tensor (caller re-interprets)
	auto out = torch::empty({nbytes}, owner.options().dtype(torch::kUInt8));
	if (nbytes > 0)
		cudaMemcpy(out.data_ptr(), p, (size_t)nbytes, cudaMemcpyDeviceToDevice);
	return out;
}

py::dict unpack_geom(torch::Tensor geomBuffer, int64_t P)
{
	char* chunk = reinterpret_cast<char*>(geomBuffer.data_ptr());
	auto g = CudaRasterizer::GeometryState::fromChunk(chunk, (size_t)P);
	py::dict d;
	d["depths"] = view_bytes(g.depths, 4 * P, geomBuffer).view(torch::kFloat32);
	d["clamped"] = view_bytes(g.clamped, 3 * P, geomBuffer).view({P, 3});
	d["internal_radii"] = view_bytes(g.internal_radii, 4 * P, geomBuffer).view(torch::kInt32);
	d["means2D"] = view_bytes(g.means2D, 8 * P, geomBuffer).view(torch::kFloat32).view({P, 2});
	d["cov3D"] = view_bytes(g.cov3D, 24 * P, geomBuffer).view(torch::kFloat32).view({P, 6});
	d["conic_opacity"] = view_bytes(g.conic_opacity, 16 * P, geomBuffer).view(torch::kFloat32).view({P, 4});
	d["rgb"] = view_bytes(g.rgb, 12 * P, geomBuffer).view(torch::kFloat32).view({P, 3});
	d["tiles_touched"] = view_bytes(g.tiles_touched, 4 * P, geomBuffer).view(torch::kInt32);
	d["point_offsets"] = view_bytes(g.point_offsets, 4 * P, geomBuffer).view(torch::kInt32);
	return d;
}

py::dict unpack_binning(torch::Tensor binningBuffer, int64_t R)
{
	char* chunk = reinterpret_cast<char*>(binningBuffer.data_ptr());
	auto b = CudaRasterizer::BinningState::fromChunk(chunk, (size_t)R);
	py::dict d;
	d["point_list"] = view_bytes(b.point_list, 4 * R, binningBuffer).view(torch::kInt32);
	d["point_list_unsorted"] = view_bytes(b.point_list_unsorted, 4 * R, binningBuffer).view(torch::kInt32);
	d["point_list_keys"] = view_bytes(b.point_list_keys, 8 * R, binningBuffer).view(torch::kInt64);
	d["point_list_keys_unsorted"] = view_bytes(b.point_list_keys_unsorted, 8 * R, binningBuffer).view(torch::kInt64);
	return d;
}

py::dict unpack_img(torch::Tensor imgBuffer, int64_t N, int64_t T)
{
	char* chunk = reinterpret_cast<char*>(imgBuffer.data_ptr());
	auto s = CudaRasterizer::ImageState::fromChunk(chunk, (size_t)N);
	py::dict d;
	d["accum_alpha"] = view_bytes(s.accum_alpha, 4 * N, imgBuffer).view(torch::kFloat32);
	d["n_contrib"] = view_bytes(s.n_contrib, 4 * N, imgBuffer).view(torch::kInt32);
	d["ranges"] = view_bytes(s.ranges, 8 * T, imgBuffer).view(torch::kInt32).view({T, 2});
	return d;
}

// The reference's {Lonlat,}Rasterizer::backward (cuda_rasterizer/rasterizer.h:66-92,126-154) called with our own output tensors, so
// that its intermediate dL_dconic — which RasterizeGaussiansBackwardCUDA allocates and drops (src/rasterize_points.cu:203) —
// can be inspected: the input of its per-Gaussian chain, needed to evaluate that chain in double (tests/test_parity_gpu.py).
// Same argument preparation as src/rasterize_points.cu:246-276.  Returns the 8-tuple plus dL_dconic [P,4].
// camera_type 3 -> LonlatRasterizer::backward, camera_type 1 -> Rasterizer::backward (rasterizer.h:66-92; projmatrix and
// tan(fov/2) are used by that camera only).
std::vector<torch::Tensor> backward_with_conic(
	const torch::Tensor& background, const torch::Tensor& means3D, const torch::Tensor& radii, const torch::Tensor& colors,
	const torch::Tensor& scales, const torch::Tensor& rotations, float scale_modifier, const torch::Tensor& cov3D_precomp,
	const torch::Tensor& viewmatrix, const torch::Tensor& projmatrix, float tan_fovx, float tan_fovy,
	const torch::Tensor& dL_dout_color, const torch::Tensor& sh, int degree,
	const torch::Tensor& campos, const torch::Tensor& geomBuffer, int R, const torch::Tensor& binningBuffer,
	const torch::Tensor& imageBuffer, int camera_type)
{
	const int P = means3D.size(0), H = dL_dout_color.size(1), W = dL_dout_color.size(2);
	const int M = sh.size(0) != 0 ? sh.size(1) : 0;
	auto o = means3D.options();
	auto z = [&](std::vector<int64_t> s) { return torch::zeros(s, o); };
	torch::Tensor dL_dmeans3D = z({P, 3}), dL_dmeans2D = z({P, 3}), dL_dcolors = z({P, 3}), dL_dconic = z({P, 2, 2}),
	              dL_dopacity = z({P, 1}), dL_dcov3D = z({P, 6}), dL_dsh = z({P, M, 3}), dL_dscales = z({P, 3}),
	              dL_drotations = z({P, 4}), dpx_dt = z({P, 3}), dpy_dt = z({P, 3});
	if (P != 0 && camera_type == 3)
		CudaRasterizer::LonlatRasterizer::backward(P, degree, M, R, background.contiguous().data_ptr<float>(), W, H,
			means3D.contiguous().data_ptr<float>(), sh.contiguous().data_ptr<float>(), colors.contiguous().data_ptr<float>(),
			scales.data_ptr<float>(), scale_modifier, rotations.data_ptr<float>(), cov3D_precomp.contiguous().data_ptr<float>(),
			viewmatrix.contiguous().data_ptr<float>(), campos.contiguous().data_ptr<float>(), radii.contiguous().data_ptr<int>(),
			reinterpret_cast<char*>(geomBuffer.contiguous().data_ptr()), reinterpret_cast<char*>(binningBuffer.contiguous().data_ptr()),
			reinterpret_cast<char*>(imageBuffer.contiguous().data_ptr()), dL_dout_color.contiguous().data_ptr<float>(),
			dL_dmeans2D.data_ptr<float>(), dL_dconic.data_ptr<float>(), dL_dopacity.data_ptr<float>(), dL_dcolors.data_ptr<float>(),
			dL_dmeans3D.data_ptr<float>(), dL_dcov3D.data_ptr<float>(), dL_dsh.data_ptr<float>(), dL_dscales.data_ptr<float>(),
			dL_drotations.data_ptr<float>(), dpx_dt.data_ptr<float>(), dpy_dt.data_ptr<float>());
	else if (P != 0)
		CudaRasterizer::Rasterizer::backward(P, degree, M, R, background.contiguous().data_ptr<float>(), W, H,
			means3D.contiguous().data_ptr<float>(), sh.contiguous().data_ptr<float>(), colors.contiguous().data_ptr<float>(),
			scales.data_ptr<float>(), scale_modifier, rotations.data_ptr<float>(), cov3D_precomp.contiguous().data_ptr<float>(),
			viewmatrix.contiguous().data_ptr<float>(), projmatrix.contiguous().data_ptr<float>(), campos.contiguous().data_ptr<float>(),
			tan_fovx, tan_fovy, radii.contiguous().data_ptr<int>(),
			reinterpret_cast<char*>(geomBuffer.contiguous().data_ptr()), reinterpret_cast<char*>(binningBuffer.contiguous().data_ptr()),
			reinterpret_cast<char*>(imageBuffer.contiguous().data_ptr()), dL_dout_color.contiguous().data_ptr<float>(),
			dL_dmeans2D.data_ptr<float>(), dL_dconic.data_ptr<float>(), dL_dopacity.data_ptr<float>(), dL_dcolors.data_ptr<float>(),
			dL_dmeans3D.data_ptr<float>(), dL_dcov3D.data_ptr<float>(), dL_dsh.data_ptr<float>(), dL_dscales.data_ptr<float>(),
			dL_drotations.data_ptr<float>());
	return { dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscales, dL_drotations, dL_dconic.view({P, 4}) };
}

} // namespace

PYBIND11_MODULE(omnigs_ref, m)
{
	m.doc() = "UNMODIFIED raikuma/OmniGS-fork rasterizer (oracle build, test infrastructure)";
	m.def("RasterizeGaussiansCUDA", &RasterizeGaussiansCUDA);
	m.def("RasterizeGaussiansBackwardCUDA", &RasterizeGaussiansBackwardCUDA);
	m.def("markVisible", &markVisible);
	m.def("unpack_geom", &unpack_geom);
	m.def("backward_with_conic", &backward_with_conic);
	m.def("unpack_binning", &unpack_binning);
	m.def("unpack_img", &unpack_img);
}
