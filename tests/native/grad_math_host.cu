// grad_math_host.cu — host build of omnigs-fork_b200/csrc/gaussian_grad.cuh for tests/test_grad_math_cpu.py.
// Compiled by the test with `nvcc -x cu` as a host-only shared library (no device code is launched): the product's own
// per-Gaussian backward chains evaluated in float (what the kernel computes) and in double (ground truth), with the
// argument list of the oracle's ogs_oracle_preprocess_bwd / ogs_oracle_pinhole_preprocess_bwd.
#include "../../omnigs-fork_b200/csrc/gaussian_grad.cuh"
#include <stdint.h>
#include <vector>

using namespace ogs::grad;

template <typename F>
static void run(int P, int D, int M, const float* means3D, const int32_t* radii, const float* shs, const uint8_t* clamped,
                const float* scales, const float* rotations, float mod, const float* cov3D, const float* viewmatrix,
                const float* projmatrix, int W, int H, float tan_fovx, float tan_fovy, const float* campos,
                const float* dL_dmean2D, const float* dL_dconic, const float* dL_dcolor,
                F* dL_dmeans3D, F* dL_dcov3D, F* dL_dsh, F* dL_dscale, F* dL_drot)
{
	F V[16], Pm[16];
	for (int i = 0; i < 16; i++) { V[i] = viewmatrix[i]; Pm[i] = projmatrix ? projmatrix[i] : 0.f; }
	const Vec3<F> cam = { (F)campos[0], (F)campos[1], (F)campos[2] };
	// rasterizer_impl.cu:476-477 evaluates the focal lengths in float
	const float fxf = projmatrix ? W / (2.0f * tan_fovx) : 0.f, fyf = projmatrix ? H / (2.0f * tan_fovy) : 0.f;
#pragma omp parallel for schedule(static)
	for (int i = 0; i < P; i++) {
		if (!(radii[i] > 0)) continue;
		const Vec3<F> mean = { (F)means3D[3 * i], (F)means3D[3 * i + 1], (F)means3D[3 * i + 2] };
		F c6[6], d6[6];
		for (int k = 0; k < 6; k++) c6[k] = cov3D[6 * (size_t)i + k];
		const F gmx = dL_dmean2D[3 * (size_t)i], gmy = dL_dmean2D[3 * (size_t)i + 1];
		const F gA = dL_dconic[4 * (size_t)i], gB = dL_dconic[4 * (size_t)i + 1], gC = dL_dconic[4 * (size_t)i + 3];
		Vec3<F> dm = projmatrix
			? projection_backward<F, true>(mean, c6, V, Pm, W, H, (F)fxf, (F)fyf, (F)tan_fovx, (F)tan_fovy, gmx, gmy, gA, gB, gC, d6)
			: projection_backward<F, false>(mean, c6, V, Pm, W, H, (F)0, (F)0, (F)0, (F)0, gmx, gmy, gA, gB, gC, d6);
		for (int k = 0; k < 6; k++) dL_dcov3D[6 * (size_t)i + k] = d6[k];
		if (shs) {
			F dRGB[3];
			for (int c = 0; c < 3; c++) dRGB[c] = clamped[3 * (size_t)i + c] ? (F)0 : (F)dL_dcolor[3 * (size_t)i + c];
			const float* row = shs + (size_t)i * M * 3;
			F* drow = dL_dsh + (size_t)i * M * 3;
			auto sh = [row](int k, int c) { return (F)row[3 * k + c]; };
			auto dsh = [drow](int k, int c, F v) { drow[3 * k + c] = v; };
			const Vec3<F> d3 = colour_backward<F>(D, mean, cam, sh, dRGB, dsh);
			dm.x += d3.x; dm.y += d3.y; dm.z += d3.z;
		}
		dL_dmeans3D[3 * (size_t)i] = dm.x; dL_dmeans3D[3 * (size_t)i + 1] = dm.y; dL_dmeans3D[3 * (size_t)i + 2] = dm.z;
		if (scales) {
			const F s[3] = { (F)mod * scales[3 * i], (F)mod * scales[3 * i + 1], (F)mod * scales[3 * i + 2] };
			const F q[4] = { rotations[4 * i], rotations[4 * i + 1], rotations[4 * i + 2], rotations[4 * i + 3] };
			scale_rotation_grad<F>(s, q, d6, dL_dscale + 3 * (size_t)i, dL_drot + 4 * (size_t)i);
		}
	}
}

#define ARGS int P, int D, int M, const float* means3D, const int32_t* radii, const float* shs, const uint8_t* clamped,  \
	const float* scales, const float* rotations, float mod, const float* cov3D, const float* viewmatrix,                \
	const float* projmatrix, int W, int H, float tan_fovx, float tan_fovy, const float* campos,                         \
	const float* dL_dmean2D, const float* dL_dconic, const float* dL_dcolor
#define PASS P, D, M, means3D, radii, shs, clamped, scales, rotations, mod, cov3D, viewmatrix, projmatrix, W, H, tan_fovx,  \
	tan_fovy, campos, dL_dmean2D, dL_dconic, dL_dcolor

extern "C" void ogs_grad_host_f32(ARGS, float* dL_dmeans3D, float* dL_dcov3D, float* dL_dsh, float* dL_dscale, float* dL_drot)
{
	run<float>(PASS, dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscale, dL_drot);
}
extern "C" void ogs_grad_host_f64(ARGS, double* dL_dmeans3D, double* dL_dcov3D, double* dL_dsh, double* dL_dscale, double* dL_drot)
{
	run<double>(PASS, dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscale, dL_drot);
}
