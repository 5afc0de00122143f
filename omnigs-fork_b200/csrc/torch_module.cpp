// torch_module.cpp — Python binding of the C++ drop-in (rasterize_points.cpp) so the parity tests can
// drive the very symbols the reference's C++ caller would link against.
#include <torch/extension.h>

#include "rasterize_points.h"

PYBIND11_MODULE(omnigs_b200_torch, m)
{
	m.doc() = "libomnigs_b200 LibTorch drop-in: RasterizeGaussiansCUDA / RasterizeGaussiansBackwardCUDA / markVisible";
	m.def("RasterizeGaussiansCUDA", &RasterizeGaussiansCUDA);
	m.def("RasterizeGaussiansBackwardCUDA", &RasterizeGaussiansBackwardCUDA);
	m.def("markVisible", &markVisible);
}
