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

} // namespace

PYBIND11_MODULE(omnigs_ref, m)
{
	m.doc() = "UNMODIFIED raikuma/OmniGS-fork rasterizer (oracle build, test infrastructure)";
	m.def("RasterizeGaussiansCUDA", &RasterizeGaussiansCUDA);
	m.def("RasterizeGaussiansBackwardCUDA", &RasterizeGaussiansBackwardCUDA);
	m.def("markVisible", &markVisible);
	m.def("unpack_geom", &unpack_geom);
	m.def("unpack_binning", &unpack_binning);
	m.def("unpack_img", &unpack_img);
}
