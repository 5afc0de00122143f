// pinhole_math.cuh — per-Gaussian math of the perspective camera (camera_type = 1).
//
// Same contract as lonlat_math.cuh: everything that feeds an integer or bit-compared output of the forward
// (depth key, pixel position, conic, radius, tile rect) is written with explicit round-to-nearest intrinsics in
// the operation order of the reference's compiled preprocessCUDA (cuda_rasterizer/forward.cu:232-340, read off
// its sm_100 SASS): sums of three products are fma(a2,b2, fma(a0,b0, mul(a1,b1))), `/` is an IEEE division,
// 1/x an IEEE reciprocal.  The backward (backward.cu:156-292, :558-608) is tolerance-bound: gaussian_grad.cuh.
#pragma once
#include "lonlat_math.cuh"

namespace ogs {

// transformPoint4x4 (auxiliary.h:95-105): p_hom = M * (p, 1), M column-major
OGS_D float4 proj_point_p(const float* M, float3 p)
{
	float4 h;
	h.x = __fadd_rn(dot3p(p.x, M[0], p.y, M[4], p.z, M[8]), M[12]);
	h.y = __fadd_rn(dot3p(p.x, M[1], p.y, M[5], p.z, M[9]), M[13]);
	h.z = __fadd_rn(dot3p(p.x, M[2], p.y, M[6], p.z, M[10]), M[14]);
	h.w = __fadd_rn(dot3p(p.x, M[3], p.y, M[7], p.z, M[11]), M[15]);
	return h;
}

// The perspective Jacobian with the reference's frustum clamp (forward.cu:94-108 / backward.cu:179-196).
// t is the camera-space mean; on return t.x, t.y hold the clamped values the Jacobian was evaluated at.
struct PinholeJac {   // the four non-zero entries of d(pixel)/d(t)
	float j00, j02, j11, j12;
};
OGS_D PinholeJac pinhole_jacobian_p(float3& t, float focal_x, float focal_y, float tan_fovx, float tan_fovy,
                                   float& x_grad_mul, float& y_grad_mul)
{
	const float limx = __fmul_rn(1.3f, tan_fovx);
	const float limy = __fmul_rn(1.3f, tan_fovy);
	const float txtz = __fdiv_rn(t.x, t.z);
	const float tytz = __fdiv_rn(t.y, t.z);
	t.x = __fmul_rn(fminf(limx, fmaxf(-limx, txtz)), t.z);
	t.y = __fmul_rn(fminf(limy, fmaxf(-limy, tytz)), t.z);
	x_grad_mul = (txtz < -limx || txtz > limx) ? 0.f : 1.f;
	y_grad_mul = (tytz < -limy || tytz > limy) ? 0.f : 1.f;
	const float tzz = __fmul_rn(t.z, t.z);
	PinholeJac J;
	J.j00 = __fdiv_rn(focal_x, t.z);
	J.j02 = __fdiv_rn(-__fmul_rn(focal_x, t.x), tzz);
	J.j11 = __fdiv_rn(focal_y, t.z);
	J.j12 = __fdiv_rn(-__fmul_rn(focal_y, t.y), tzz);
	return J;
}

// computeCov2D (forward.cu:86-128): 2-D covariance with the 0.3 px blur; same T = W J, T^T Vrk^T T products
// as the lonlat camera (zero entries multiplied through like glm's generic product does).
OGS_D float3 cov2d_pinhole_p(float3 t, const float* V, const float* c6, float focal_x, float focal_y,
                             float tan_fovx, float tan_fovy)
{
	float gx, gy;
	const PinholeJac J = pinhole_jacobian_p(t, focal_x, focal_y, tan_fovx, tan_fovy, gx, gy);
	float T0[3], T1[3];
#pragma unroll
	for (int k = 0; k < 3; k++) {
		T0[k] = dot3p(V[4 * k + 0], J.j00, V[4 * k + 1], 0.0f, V[4 * k + 2], J.j02);
		T1[k] = dot3p(V[4 * k + 0], 0.0f, V[4 * k + 1], J.j11, V[4 * k + 2], J.j12);
	}
	const float v0[3] = { c6[0], c6[1], c6[2] }, v1[3] = { c6[1], c6[3], c6[4] }, v2[3] = { c6[2], c6[4], c6[5] };
	const float p00 = dot3p(T0[0], v0[0], T0[1], v0[1], T0[2], v0[2]);
	const float p01 = dot3p(T1[0], v0[0], T1[1], v0[1], T1[2], v0[2]);
	const float p10 = dot3p(T0[0], v1[0], T0[1], v1[1], T0[2], v1[2]);
	const float p11 = dot3p(T1[0], v1[0], T1[1], v1[1], T1[2], v1[2]);
	const float p20 = dot3p(T0[0], v2[0], T0[1], v2[1], T0[2], v2[2]);
	const float p21 = dot3p(T1[0], v2[0], T1[1], v2[1], T1[2], v2[2]);
	float3 cov;
	cov.x = __fadd_rn(dot3p(p00, T0[0], p10, T0[1], p20, T0[2]), 0.3f);
	cov.y = dot3p(p01, T0[0], p11, T0[1], p21, T0[2]);
	cov.z = __fadd_rn(dot3p(p01, T1[0], p11, T1[1], p21, T1[2]), 0.3f);
	return cov;
}

} // namespace ogs
