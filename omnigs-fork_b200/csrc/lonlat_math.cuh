// lonlat_math.cuh — per-Gaussian equirectangular projection math, forward and backward.
//
// Bit-exactness contract: radii, tile rects and depth keys must equal the reference's bit for
// bit, so the float expression SHAPES below (operand order, association, the glm mat3 product
// formula, the double-precision ndc->pixel step) follow the reference on purpose; nvcc's default
// -fmad contraction then makes the same fusing decisions.  Reference lines are cited per function
// (raikuma/OmniGS-fork, cuda_rasterizer/).  Everything else about the kernels (fusion, layout,
// scheduling) is ours.
#pragma once
#include "ogs_common.cuh"

namespace ogs {

// Column-major 3x3, c[col][row]; product formula of glm (the reference's matrix library):
// (A*B)[c][r] = A[0][r]*B[c][0] + A[1][r]*B[c][1] + A[2][r]*B[c][2]
struct M3 {
	float c[3][3];
};
OGS_HD M3 m3_cols(float a0, float a1, float a2, float b0, float b1, float b2, float c0, float c1, float c2)
{
	M3 m;
	m.c[0][0] = a0; m.c[0][1] = a1; m.c[0][2] = a2;
	m.c[1][0] = b0; m.c[1][1] = b1; m.c[1][2] = b2;
	m.c[2][0] = c0; m.c[2][1] = c1; m.c[2][2] = c2;
	return m;
}
OGS_HD M3 m3_mul(const M3& A, const M3& B)
{
	M3 R;
#pragma unroll
	for (int c = 0; c < 3; c++)
#pragma unroll
		for (int r = 0; r < 3; r++)
			R.c[c][r] = A.c[0][r] * B.c[c][0] + A.c[1][r] * B.c[c][1] + A.c[2][r] * B.c[c][2];
	return R;
}
OGS_HD M3 m3_t(const M3& A)
{
	return m3_cols(A.c[0][0], A.c[1][0], A.c[2][0],
	               A.c[0][1], A.c[1][1], A.c[2][1],
	               A.c[0][2], A.c[1][2], A.c[2][2]);
}

// t = Tcw * p  (auxiliary.h:85-93); V is Tcw stored column-major
OGS_HD float3 view_point(const float* V, float3 p)
{
	float3 t = {
		V[0] * p.x + V[4] * p.y + V[8] * p.z + V[12],
		V[1] * p.x + V[5] * p.y + V[9] * p.z + V[13],
		V[2] * p.x + V[6] * p.y + V[10] * p.z + V[14],
	};
	return t;
}
// R^T * v  (auxiliary.h:116-124)
OGS_HD float3 view_vec_t(const float* V, float3 p)
{
	float3 o = {
		V[0] * p.x + V[1] * p.y + V[2] * p.z,
		V[4] * p.x + V[5] * p.y + V[6] * p.z,
		V[8] * p.x + V[9] * p.y + V[10] * p.z,
	};
	return o;
}

// NDC -> pixel, evaluated in double like the reference (auxiliary.h:51-54)
OGS_HD float ndc_to_pix(float v, int S)
{
	return ((v + 1.0) * S - 1.0) * 0.5;
}

// clamped tile rect (auxiliary.h:56-66 getRect)
OGS_HD void tile_rect(float2 p, int max_radius, int gx, int gy, int& x0, int& y0, int& x1, int& y1)
{
	x0 = min(gx, max((int)0, (int)((p.x - max_radius) / kTile)));
	y0 = min(gy, max((int)0, (int)((p.y - max_radius) / kTile)));
	x1 = min(gx, max((int)0, (int)((p.x + max_radius + kTile - 1) / kTile)));
	y1 = min(gy, max((int)0, (int)((p.y + max_radius + kTile - 1) / kTile)));
}

// quaternion (w,x,y,z) -> rotation, NOT normalised (forward.cu:203-214)
OGS_HD M3 quat_matrix(float4 q)
{
	float r = q.x, x = q.y, y = q.z, z = q.w;
	return m3_cols(
		1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y),
		2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x),
		2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y));
}

// Sigma = (S R)^T (S R), upper triangle (forward.cu:194-228)
OGS_HD void cov3d_from_scale_rot(float3 scale, float mod, float4 q, float* cov6)
{
	M3 S = m3_cols(1.0f, 0.f, 0.f, 0.f, 1.0f, 0.f, 0.f, 0.f, 1.0f);
	S.c[0][0] = mod * scale.x;
	S.c[1][1] = mod * scale.y;
	S.c[2][2] = mod * scale.z;
	M3 R = quat_matrix(q);
	M3 Mm = m3_mul(S, R);
	M3 Sigma = m3_mul(m3_t(Mm), Mm);
	cov6[0] = Sigma.c[0][0];
	cov6[1] = Sigma.c[0][1];
	cov6[2] = Sigma.c[0][2];
	cov6[3] = Sigma.c[1][1];
	cov6[4] = Sigma.c[1][2];
	cov6[5] = Sigma.c[2][2];
}

// The five non-zero entries of d(pixel)/d(t) for the equirect projection (forward.cu:147-162)
struct LonlatJac {
	float j00, j02, j10, j11, j12;
};
OGS_HD LonlatJac lonlat_jacobian(float3 t, int W, int H)
{
	float trxztrxz = t.x * t.x + t.z * t.z;
	float trxztrxz_inv = 1.0f / (trxztrxz + kEps7);
	float trxz = sqrtf(trxztrxz);
	float trxz_inv = 1.0f / (trxz + kEps7);
	float trtr = trxztrxz + t.y * t.y;
	float trtr_inv = 1.0f / (trtr + kEps7);

	float W_div_2pi = W * 0.5f * kPiInv;
	float H_div_pi = H * kPiInv;

	LonlatJac J;
	J.j00 = W_div_2pi * t.z * trxztrxz_inv;
	J.j02 = -W_div_2pi * t.x * trxztrxz_inv;
	J.j10 = -H_div_pi * t.x * t.y * trxz_inv * trtr_inv;
	J.j11 = H_div_pi * trxz * trtr_inv;
	J.j12 = -H_div_pi * t.z * t.y * trxz_inv * trtr_inv;
	return J;
}

// T = W*J, cov = T^T Vrk^T T  (forward.cu:164-181); the zero entries are multiplied through,
// exactly as the generic glm product does.
OGS_HD void lonlat_T_cov(const LonlatJac& Jv, const float* V, const float* cov6, M3& T, M3& Vrk, M3& cov)
{
	M3 J = m3_cols(Jv.j00, 0.0f, Jv.j02, Jv.j10, Jv.j11, Jv.j12, 0.0f, 0.0f, 0.0f);
	M3 Wm = m3_cols(V[0], V[4], V[8], V[1], V[5], V[9], V[2], V[6], V[10]);
	T = m3_mul(Wm, J);
	Vrk = m3_cols(cov6[0], cov6[1], cov6[2], cov6[1], cov6[3], cov6[4], cov6[2], cov6[4], cov6[5]);
	cov = m3_mul(m3_mul(m3_t(T), m3_t(Vrk)), T);
}

// ------------------------------------------------------------------ pinned forward chain
// Everything that feeds an integer or bit-compared output of the forward (depth key, pixel
// position, conic, radius, tile rect) is written with explicit round-to-nearest intrinsics in the
// exact operation order of the reference's compiled kernel (read off its sm_100 SASS, see
// DESIGN.md "Numeric pinning"), so the result cannot drift with compiler context:
//   * a sum of three products is  fma(a2,b2, fma(a0,b0, mul(a1,b1)))   (NVVM's contraction of
//     (a0*b0 + a1*b1) + a2*b2), including the products with literal zeros glm's generic mat3
//     multiply carries along;
//   * det = fma(cx, cz, -(cy*cy)) and mid*mid - det = fma(mid, mid, -det)  (fused by ptxas);
//   * ndc->pixel runs in double: ((v + 1.0) * S - 1.0) * 0.5 with the middle step a DFMA.
OGS_D float dot3p(float a0, float b0, float a1, float b1, float a2, float b2)
{
	return __fmaf_rn(a2, b2, __fmaf_rn(a0, b0, __fmul_rn(a1, b1)));
}
OGS_D M3 m3_mul_p(const M3& A, const M3& B)
{
	M3 R;
#pragma unroll
	for (int c = 0; c < 3; c++)
#pragma unroll
		for (int r = 0; r < 3; r++)
			R.c[c][r] = dot3p(A.c[0][r], B.c[c][0], A.c[1][r], B.c[c][1], A.c[2][r], B.c[c][2]);
	return R;
}
OGS_D float3 view_point_p(const float* V, float3 p)
{
	float3 t;
	t.x = __fadd_rn(dot3p(p.x, V[0], p.y, V[4], p.z, V[8]), V[12]);
	t.y = __fadd_rn(dot3p(p.x, V[1], p.y, V[5], p.z, V[9]), V[13]);
	t.z = __fadd_rn(dot3p(p.x, V[2], p.y, V[6], p.z, V[10]), V[14]);
	return t;
}
OGS_D float ndc_to_pix_p(float v, int S)
{
	return (float)__dmul_rn(__fma_rn(__dadd_rn((double)v, 1.0), (double)S, -1.0), 0.5);
}
OGS_D void tile_rect_p(float2 p, int max_radius, int gx, int gy, int& x0, int& y0, int& x1, int& y1)
{
	const float R = (float)max_radius, inv = 1.0f / kTile; // division by 16 is an exact scaling
	x0 = min(gx, max(0, (int)__fmul_rn(__fsub_rn(p.x, R), inv)));
	y0 = min(gy, max(0, (int)__fmul_rn(__fsub_rn(p.y, R), inv)));
	// (p + R + BLOCK) - 1, left to right as written in auxiliary.h:63-64
	x1 = min(gx, max(0, (int)__fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(p.x, R), (float)kTile), -1.0f), inv)));
	y1 = min(gy, max(0, (int)__fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(p.y, R), (float)kTile), -1.0f), inv)));
}
// The model's activations (gaussian_model.cpp:54-77) as LibTorch evaluates them in float32:
// sigmoid(x) = 1 / (1 + exp(-x)); normalize(q) = q / max(||q||_2, 1e-12).
OGS_D float sigmoid_act(float x) { return 1.0f / (1.0f + expf(-x)); }
OGS_D float quat_norm_clamped(float4 q)
{
	return fmaxf(sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w), 1e-12f);
}
OGS_D float4 normalize_quat(float4 q)
{
	const float n = quat_norm_clamped(q);
	return make_float4(q.x / n, q.y / n, q.z / n, q.w / n);
}
OGS_D void cov3d_from_scale_rot_p(float3 scale, float mod, float4 q, float* cov6)
{
	const float r = q.x, x = q.y, y = q.z, z = q.w;
	// products kept as plain multiplies / fused into an fma exactly as ptxas scheduled them
	const float yy = __fmul_rn(y, y), zz = __fmul_rn(z, z), xz = __fmul_rn(x, z), rz = __fmul_rn(r, z), rx = __fmul_rn(r, x);
	M3 R = m3_cols(
		__fsub_rn(1.f, __fmul_rn(2.f, __fadd_rn(yy, zz))), __fmul_rn(2.f, __fmaf_rn(x, y, -rz)), __fmul_rn(2.f, __fmaf_rn(r, y, xz)),
		__fmul_rn(2.f, __fmaf_rn(x, y, rz)), __fsub_rn(1.f, __fmul_rn(2.f, __fmaf_rn(x, x, zz))), __fmul_rn(2.f, __fmaf_rn(y, z, -rx)),
		__fmul_rn(2.f, __fmaf_rn(-r, y, xz)), __fmul_rn(2.f, __fmaf_rn(y, z, rx)), __fsub_rn(1.f, __fmul_rn(2.f, __fmaf_rn(x, x, yy))));
	M3 S = m3_cols(__fmul_rn(mod, scale.x), 0.f, 0.f, 0.f, __fmul_rn(mod, scale.y), 0.f, 0.f, 0.f, __fmul_rn(mod, scale.z));
	M3 Mm = m3_mul_p(S, R);
	M3 Sigma = m3_mul_p(m3_t(Mm), Mm);
	cov6[0] = Sigma.c[0][0];
	cov6[1] = Sigma.c[0][1];
	cov6[2] = Sigma.c[0][2];
	cov6[3] = Sigma.c[1][1];
	cov6[4] = Sigma.c[1][2];
	cov6[5] = Sigma.c[2][2];
}
// 2-D covariance (with the 0.3 blur) of a Gaussian whose camera-space mean is t.
OGS_D float3 cov2d_lonlat_p(float3 t, const float* V, const float* c6, int W, int H)
{
	const float a = __fmaf_rn(t.x, t.x, __fmul_rn(t.z, t.z));
	const float a_inv = __frcp_rn(__fadd_rn(a, kEps7));
	const float rho = __fsqrt_rn(a);
	const float rho_inv = __frcp_rn(__fadd_rn(rho, kEps7));
	const float rr = __fmaf_rn(t.y, t.y, a);
	const float rr_inv = __frcp_rn(__fadd_rn(rr, kEps7));
	const float Wd = __fmul_rn(__fmul_rn((float)W, 0.5f), kPiInv);
	const float Hd = __fmul_rn((float)H, kPiInv);
	const float j00 = __fmul_rn(__fmul_rn(Wd, t.z), a_inv);
	const float j02 = __fmul_rn(__fmul_rn(t.x, -Wd), a_inv);
	const float j10 = __fmul_rn(__fmul_rn(__fmul_rn(__fmul_rn(t.x, -Hd), t.y), rho_inv), rr_inv);
	const float j11 = __fmul_rn(__fmul_rn(Hd, rho), rr_inv);
	const float j12 = __fmul_rn(__fmul_rn(__fmul_rn(t.y, __fmul_rn(t.z, -Hd)), rho_inv), rr_inv);
	// T = W*J: rows 0 and 1 of (J R); the third is identically unused
	float T0[3], T1[3];
#pragma unroll
	for (int k = 0; k < 3; k++) {
		T0[k] = dot3p(V[4 * k + 0], j00, V[4 * k + 1], 0.0f, V[4 * k + 2], j02);
		T1[k] = dot3p(V[4 * k + 0], j10, V[4 * k + 1], j11, V[4 * k + 2], j12);
	}
	// (T^T Vrk^T)[k][r] = T_r . Vrk[:,k]
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

// ------------------------------------------------------------------ spherical harmonics
__device__ const float kSH_C0 = 0.28209479177387814f;
__device__ const float kSH_C1 = 0.4886025119029199f;
__device__ const float kSH_C2[] = { 1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                                    -1.0925484305920792f, 0.5462742152960396f };
__device__ const float kSH_C3[] = { -0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                                    0.3731763325901154f, -0.4570457994644658f, 1.445305721320277f,
                                    -0.5900435899266435f };

struct V3 {
	float x, y, z;
};
OGS_D V3 operator+(V3 a, V3 b) { return { a.x + b.x, a.y + b.y, a.z + b.z }; }
OGS_D V3 operator-(V3 a, V3 b) { return { a.x - b.x, a.y - b.y, a.z - b.z }; }
OGS_D V3 operator*(float s, V3 a) { return { s * a.x, s * a.y, s * a.z }; }
OGS_D V3 operator*(V3 a, float s) { return { a.x * s, a.y * s, a.z * s }; }
OGS_D V3& operator+=(V3& a, V3 b) { a.x += b.x; a.y += b.y; a.z += b.z; return a; }
OGS_D float dot3(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

// SH -> RGB for one Gaussian; `sh` points at its M x 3 coefficients (forward.cu:30-83).
// Returns the clamp mask in bits 0..2.
template <typename ShPtr>
OGS_D V3 sh_to_rgb(int deg, float3 mean, float3 campos, ShPtr sh, unsigned& clamp_mask)
{
	V3 dir = { mean.x - campos.x, mean.y - campos.y, mean.z - campos.z };
	float len = sqrtf(dot3(dir, dir));
	dir = { dir.x / len, dir.y / len, dir.z / len };

	V3 result = kSH_C0 * sh(0);
	if (deg > 0) {
		float x = dir.x, y = dir.y, z = dir.z;
		result = result - kSH_C1 * y * sh(1) + kSH_C1 * z * sh(2) - kSH_C1 * x * sh(3);
		if (deg > 1) {
			float xx = x * x, yy = y * y, zz = z * z;
			float xy = x * y, yz = y * z, xz = x * z;
			result = result +
				kSH_C2[0] * xy * sh(4) +
				kSH_C2[1] * yz * sh(5) +
				kSH_C2[2] * (2.0f * zz - xx - yy) * sh(6) +
				kSH_C2[3] * xz * sh(7) +
				kSH_C2[4] * (xx - yy) * sh(8);
			if (deg > 2) {
				result = result +
					kSH_C3[0] * y * (3.0f * xx - yy) * sh(9) +
					kSH_C3[1] * xy * z * sh(10) +
					kSH_C3[2] * y * (4.0f * zz - xx - yy) * sh(11) +
					kSH_C3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy) * sh(12) +
					kSH_C3[4] * x * (4.0f * zz - xx - yy) * sh(13) +
					kSH_C3[5] * z * (xx - yy) * sh(14) +
					kSH_C3[6] * x * (xx - 3.0f * yy) * sh(15);
			}
		}
	}
	result.x += 0.5f; result.y += 0.5f; result.z += 0.5f;
	clamp_mask = (result.x < 0 ? 1u : 0u) | (result.y < 0 ? 2u : 0u) | (result.z < 0 ? 4u : 0u);
	return { result.x < 0.0f ? 0.0f : result.x, result.y < 0.0f ? 0.0f : result.y, result.z < 0.0f ? 0.0f : result.z };
}

// The same evaluation with every operation pinned to the order of the reference's compiled computeColorFromSH
// (its sm_100 SASS): dir = (p - c) / sqrt(dot), every term is  result = fma(coef, sh_k, result)  with
// coef = (C * a) * b  built from plain multiplies; 2zz = zz + zz; 3xx - yy = fma(xx, 3, -yy);
// 4zz - xx - yy = fma(zz, 4, -xx) - yy; 2zz - 3xx - 3yy = fma(yy, -3, fma(xx, -3, zz + zz)); xx - 3yy = fma(yy, -3, xx).
// The colours enter the image directly, so this makes the image bit-identical to the reference's, not merely <= 1e-5.
template <typename ShPtr>
OGS_D V3 sh_to_rgb_p(int deg, float3 mean, float3 campos, ShPtr sh, unsigned& clamp_mask)
{
	const float dx = __fsub_rn(mean.x, campos.x), dy = __fsub_rn(mean.y, campos.y), dz = __fsub_rn(mean.z, campos.z);
	const float len = __fsqrt_rn(dot3p(dx, dx, dy, dy, dz, dz));
	const float x = __fdiv_rn(dx, len), y = __fdiv_rn(dy, len), z = __fdiv_rn(dz, len);
	V3 c = sh(0);
	V3 r = { __fmul_rn(c.x, kSH_C0), __fmul_rn(c.y, kSH_C0), __fmul_rn(c.z, kSH_C0) };
	auto acc = [&r, &sh](float t, int k) {
		const V3 s = sh(k);
		r.x = __fmaf_rn(t, s.x, r.x); r.y = __fmaf_rn(t, s.y, r.y); r.z = __fmaf_rn(t, s.z, r.z);
	};
	if (deg > 0) {
		acc(-__fmul_rn(kSH_C1, y), 1);
		acc(__fmul_rn(kSH_C1, z), 2);
		acc(-__fmul_rn(kSH_C1, x), 3);
		if (deg > 1) {
			const float xx = __fmul_rn(x, x), yy = __fmul_rn(y, y), zz = __fmul_rn(z, z);
			const float xy = __fmul_rn(x, y), yz = __fmul_rn(y, z), xz = __fmul_rn(x, z);
			const float zz2 = __fadd_rn(zz, zz), xx_yy = __fsub_rn(xx, yy);
			acc(__fmul_rn(kSH_C2[0], xy), 4);
			acc(__fmul_rn(kSH_C2[1], yz), 5);
			acc(__fmul_rn(kSH_C2[2], __fsub_rn(__fsub_rn(zz2, xx), yy)), 6);
			acc(__fmul_rn(kSH_C2[3], xz), 7);
			acc(__fmul_rn(kSH_C2[4], xx_yy), 8);
			if (deg > 2) {
				const float u = __fsub_rn(__fmaf_rn(zz, 4.0f, -xx), yy);   // 4zz - xx - yy
				acc(__fmul_rn(__fmul_rn(kSH_C3[0], y), __fmaf_rn(xx, 3.0f, -yy)), 9);
				acc(__fmul_rn(__fmul_rn(kSH_C3[1], xy), z), 10);
				acc(__fmul_rn(__fmul_rn(kSH_C3[2], y), u), 11);
				acc(__fmul_rn(__fmul_rn(kSH_C3[3], z), __fmaf_rn(yy, -3.0f, __fmaf_rn(xx, -3.0f, zz2))), 12);
				acc(__fmul_rn(__fmul_rn(kSH_C3[4], x), u), 13);
				acc(__fmul_rn(__fmul_rn(kSH_C3[5], z), xx_yy), 14);
				acc(__fmul_rn(__fmul_rn(kSH_C3[6], x), __fmaf_rn(yy, -3.0f, xx)), 15);
			}
		}
	}
	r.x = __fadd_rn(r.x, 0.5f); r.y = __fadd_rn(r.y, 0.5f); r.z = __fadd_rn(r.z, 0.5f);
	clamp_mask = (r.x < 0 ? 1u : 0u) | (r.y < 0 ? 2u : 0u) | (r.z < 0 ? 4u : 0u);
	return { r.x < 0.0f ? 0.0f : r.x, r.y < 0.0f ? 0.0f : r.y, r.z < 0.0f ? 0.0f : r.z };
}

// d|v|^-1 v / dv applied to dv (auxiliary.h:134-144)
OGS_D float3 dnormvdv(float3 v, float3 dv)
{
	float sum2 = v.x * v.x + v.y * v.y + v.z * v.z;
	float invsum32 = 1.0f / sqrtf(sum2 * sum2 * sum2);
	float3 o;
	o.x = ((+sum2 - v.x * v.x) * dv.x - v.y * v.x * dv.y - v.z * v.x * dv.z) * invsum32;
	o.y = (-v.x * v.y * dv.x + (sum2 - v.y * v.y) * dv.y - v.z * v.y * dv.z) * invsum32;
	o.z = (-v.x * v.z * dv.x - v.y * v.z * dv.y + (sum2 - v.z * v.z) * dv.z) * invsum32;
	return o;
}

// SH backward (backward.cu:30-151).  sh(k) reads coefficient k, dsh(k, v) stores its gradient.
// Returns the mean-gradient contribution through the view direction.
template <typename ShPtr, typename DshStore>
OGS_D float3 sh_backward(int deg, float3 mean, float3 campos, ShPtr sh, V3 dL_dRGB, DshStore dsh)
{
	V3 dir_orig = { mean.x - campos.x, mean.y - campos.y, mean.z - campos.z };
	float len = sqrtf(dot3(dir_orig, dir_orig));
	float x = dir_orig.x / len, y = dir_orig.y / len, z = dir_orig.z / len;

	V3 dRGBdx = { 0, 0, 0 }, dRGBdy = { 0, 0, 0 }, dRGBdz = { 0, 0, 0 };

	dsh(0, kSH_C0 * dL_dRGB);
	if (deg > 0) {
		dsh(1, (-kSH_C1 * y) * dL_dRGB);
		dsh(2, (kSH_C1 * z) * dL_dRGB);
		dsh(3, (-kSH_C1 * x) * dL_dRGB);
		dRGBdx = -kSH_C1 * sh(3);
		dRGBdy = -kSH_C1 * sh(1);
		dRGBdz = kSH_C1 * sh(2);
		if (deg > 1) {
			float xx = x * x, yy = y * y, zz = z * z;
			float xy = x * y, yz = y * z, xz = x * z;
			dsh(4, (kSH_C2[0] * xy) * dL_dRGB);
			dsh(5, (kSH_C2[1] * yz) * dL_dRGB);
			dsh(6, (kSH_C2[2] * (2.f * zz - xx - yy)) * dL_dRGB);
			dsh(7, (kSH_C2[3] * xz) * dL_dRGB);
			dsh(8, (kSH_C2[4] * (xx - yy)) * dL_dRGB);
			dRGBdx += kSH_C2[0] * y * sh(4) + kSH_C2[2] * 2.f * -x * sh(6) + kSH_C2[3] * z * sh(7) + kSH_C2[4] * 2.f * x * sh(8);
			dRGBdy += kSH_C2[0] * x * sh(4) + kSH_C2[1] * z * sh(5) + kSH_C2[2] * 2.f * -y * sh(6) + kSH_C2[4] * 2.f * -y * sh(8);
			dRGBdz += kSH_C2[1] * y * sh(5) + kSH_C2[2] * 2.f * 2.f * z * sh(6) + kSH_C2[3] * x * sh(7);
			if (deg > 2) {
				dsh(9, (kSH_C3[0] * y * (3.f * xx - yy)) * dL_dRGB);
				dsh(10, (kSH_C3[1] * xy * z) * dL_dRGB);
				dsh(11, (kSH_C3[2] * y * (4.f * zz - xx - yy)) * dL_dRGB);
				dsh(12, (kSH_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy)) * dL_dRGB);
				dsh(13, (kSH_C3[4] * x * (4.f * zz - xx - yy)) * dL_dRGB);
				dsh(14, (kSH_C3[5] * z * (xx - yy)) * dL_dRGB);
				dsh(15, (kSH_C3[6] * x * (xx - 3.f * yy)) * dL_dRGB);

				dRGBdx += (
					kSH_C3[0] * sh(9) * 3.f * 2.f * xy +
					kSH_C3[1] * sh(10) * yz +
					kSH_C3[2] * sh(11) * -2.f * xy +
					kSH_C3[3] * sh(12) * -3.f * 2.f * xz +
					kSH_C3[4] * sh(13) * (-3.f * xx + 4.f * zz - yy) +
					kSH_C3[5] * sh(14) * 2.f * xz +
					kSH_C3[6] * sh(15) * 3.f * (xx - yy));
				dRGBdy += (
					kSH_C3[0] * sh(9) * 3.f * (xx - yy) +
					kSH_C3[1] * sh(10) * xz +
					kSH_C3[2] * sh(11) * (-3.f * yy + 4.f * zz - xx) +
					kSH_C3[3] * sh(12) * -3.f * 2.f * yz +
					kSH_C3[4] * sh(13) * -2.f * xy +
					kSH_C3[5] * sh(14) * -2.f * yz +
					kSH_C3[6] * sh(15) * -3.f * 2.f * xy);
				dRGBdz += (
					kSH_C3[1] * sh(10) * xy +
					kSH_C3[2] * sh(11) * 4.f * 2.f * yz +
					kSH_C3[3] * sh(12) * 3.f * (2.f * zz - xx - yy) +
					kSH_C3[4] * sh(13) * 4.f * 2.f * xz +
					kSH_C3[5] * sh(14) * (xx - yy));
			}
		}
	}
	float3 dL_ddir = { dot3(dRGBdx, dL_dRGB), dot3(dRGBdy, dL_dRGB), dot3(dRGBdz, dL_dRGB) };
	return dnormvdv(float3{ dir_orig.x, dir_orig.y, dir_orig.z }, dL_ddir);
}

// cov3D backward: dL/dSigma (6) -> dL/dscale, dL/dquaternion (backward.cu:489-552)
OGS_D void cov3d_backward(float3 scale, float mod, float4 q, const float* dL_dcov6, float3& dL_dscale, float4& dL_dq)
{
	float r = q.x, x = q.y, y = q.z, z = q.w;
	M3 R = quat_matrix(q);
	M3 S = m3_cols(1.0f, 0.f, 0.f, 0.f, 1.0f, 0.f, 0.f, 0.f, 1.0f);
	float3 s = { mod * scale.x, mod * scale.y, mod * scale.z };
	S.c[0][0] = s.x;
	S.c[1][1] = s.y;
	S.c[2][2] = s.z;
	M3 Mm = m3_mul(S, R);

	M3 dL_dSigma = m3_cols(
		dL_dcov6[0], 0.5f * dL_dcov6[1], 0.5f * dL_dcov6[2],
		0.5f * dL_dcov6[1], dL_dcov6[3], 0.5f * dL_dcov6[4],
		0.5f * dL_dcov6[2], 0.5f * dL_dcov6[4], dL_dcov6[5]);

	M3 M2;
#pragma unroll
	for (int c = 0; c < 3; c++)
#pragma unroll
		for (int rr = 0; rr < 3; rr++) M2.c[c][rr] = Mm.c[c][rr] * 2.0f;
	M3 dL_dM = m3_mul(M2, dL_dSigma);

	M3 Rt = m3_t(R);
	M3 dL_dMt = m3_t(dL_dM);

	dL_dscale.x = Rt.c[0][0] * dL_dMt.c[0][0] + Rt.c[0][1] * dL_dMt.c[0][1] + Rt.c[0][2] * dL_dMt.c[0][2];
	dL_dscale.y = Rt.c[1][0] * dL_dMt.c[1][0] + Rt.c[1][1] * dL_dMt.c[1][1] + Rt.c[1][2] * dL_dMt.c[1][2];
	dL_dscale.z = Rt.c[2][0] * dL_dMt.c[2][0] + Rt.c[2][1] * dL_dMt.c[2][1] + Rt.c[2][2] * dL_dMt.c[2][2];

#pragma unroll
	for (int k = 0; k < 3; k++) {
		dL_dMt.c[0][k] *= s.x;
		dL_dMt.c[1][k] *= s.y;
		dL_dMt.c[2][k] *= s.z;
	}
#define OGS_D_(a, b) dL_dMt.c[a][b]
	dL_dq.x = 2 * z * (OGS_D_(0, 1) - OGS_D_(1, 0)) + 2 * y * (OGS_D_(2, 0) - OGS_D_(0, 2)) + 2 * x * (OGS_D_(1, 2) - OGS_D_(2, 1));
	dL_dq.y = 2 * y * (OGS_D_(1, 0) + OGS_D_(0, 1)) + 2 * z * (OGS_D_(2, 0) + OGS_D_(0, 2)) + 2 * r * (OGS_D_(1, 2) - OGS_D_(2, 1)) - 4 * x * (OGS_D_(2, 2) + OGS_D_(1, 1));
	dL_dq.z = 2 * x * (OGS_D_(1, 0) + OGS_D_(0, 1)) + 2 * r * (OGS_D_(2, 0) - OGS_D_(0, 2)) + 2 * z * (OGS_D_(1, 2) + OGS_D_(2, 1)) - 4 * y * (OGS_D_(2, 2) + OGS_D_(0, 0));
	dL_dq.w = 2 * r * (OGS_D_(0, 1) - OGS_D_(1, 0)) + 2 * x * (OGS_D_(2, 0) + OGS_D_(0, 2)) + 2 * y * (OGS_D_(1, 2) + OGS_D_(2, 1)) - 4 * z * (OGS_D_(1, 1) + OGS_D_(0, 0));
#undef OGS_D_
}

// Shared middle of the two cameras' covariance backward (backward.cu:193-268 pinhole, :370-452 lonlat):
// conic = inverse(cov2D), cov2D = T^T Vrk T + 0.3 I, T = W J.  From dL/dconic: dL/dcov3D (6) and the gradient
// w.r.t. the five possibly non-zero Jacobian entries (J rows 0 and 1; j01 is structurally zero).
OGS_D void conic_to_cov3d_and_jacobian_backward(const LonlatJac& Jv, const float* V, const float* cov6, float3 dL_dconic,
                                                 float* dL_dcov6, float& dL_dJ00, float& dL_dJ02, float& dL_dJ10,
                                                 float& dL_dJ11, float& dL_dJ12)
{
	M3 T, Vrk, cov2D;
	lonlat_T_cov(Jv, V, cov6, T, Vrk, cov2D);
	M3 Wm = m3_cols(V[0], V[4], V[8], V[1], V[5], V[9], V[2], V[6], V[10]);

	float a = cov2D.c[0][0] += 0.3f;
	float b = cov2D.c[0][1];
	float c = cov2D.c[1][1] += 0.3f;

	float denom = a * c - b * b;
	float dL_da = 0, dL_db = 0, dL_dc = 0;
	float denom2inv = 1.0f / ((denom * denom) + kEps7);

#define T_(i, j) T.c[i][j]
	if (denom2inv != 0) {
		dL_da = denom2inv * (-c * c * dL_dconic.x + 2 * b * c * dL_dconic.y + (denom - a * c) * dL_dconic.z);
		dL_dc = denom2inv * (-a * a * dL_dconic.z + 2 * a * b * dL_dconic.y + (denom - a * c) * dL_dconic.x);
		dL_db = denom2inv * 2 * (b * c * dL_dconic.x - (denom + 2 * b * b) * dL_dconic.y + a * b * dL_dconic.z);

		dL_dcov6[0] = (T_(0, 0) * T_(0, 0) * dL_da + T_(0, 0) * T_(1, 0) * dL_db + T_(1, 0) * T_(1, 0) * dL_dc);
		dL_dcov6[3] = (T_(0, 1) * T_(0, 1) * dL_da + T_(0, 1) * T_(1, 1) * dL_db + T_(1, 1) * T_(1, 1) * dL_dc);
		dL_dcov6[5] = (T_(0, 2) * T_(0, 2) * dL_da + T_(0, 2) * T_(1, 2) * dL_db + T_(1, 2) * T_(1, 2) * dL_dc);
		dL_dcov6[1] = 2 * T_(0, 0) * T_(0, 1) * dL_da + (T_(0, 0) * T_(1, 1) + T_(0, 1) * T_(1, 0)) * dL_db + 2 * T_(1, 0) * T_(1, 1) * dL_dc;
		dL_dcov6[2] = 2 * T_(0, 0) * T_(0, 2) * dL_da + (T_(0, 0) * T_(1, 2) + T_(0, 2) * T_(1, 0)) * dL_db + 2 * T_(1, 0) * T_(1, 2) * dL_dc;
		dL_dcov6[4] = 2 * T_(0, 2) * T_(0, 1) * dL_da + (T_(0, 1) * T_(1, 2) + T_(0, 2) * T_(1, 1)) * dL_db + 2 * T_(1, 1) * T_(1, 2) * dL_dc;
	} else {
#pragma unroll
		for (int i = 0; i < 6; i++) dL_dcov6[i] = 0;
	}

#define K_(i, j) Vrk.c[i][j]
	float dL_dT00 = 2 * (T_(0, 0) * K_(0, 0) + T_(0, 1) * K_(0, 1) + T_(0, 2) * K_(0, 2)) * dL_da +
		(T_(1, 0) * K_(0, 0) + T_(1, 1) * K_(0, 1) + T_(1, 2) * K_(0, 2)) * dL_db;
	float dL_dT01 = 2 * (T_(0, 0) * K_(1, 0) + T_(0, 1) * K_(1, 1) + T_(0, 2) * K_(1, 2)) * dL_da +
		(T_(1, 0) * K_(1, 0) + T_(1, 1) * K_(1, 1) + T_(1, 2) * K_(1, 2)) * dL_db;
	float dL_dT02 = 2 * (T_(0, 0) * K_(2, 0) + T_(0, 1) * K_(2, 1) + T_(0, 2) * K_(2, 2)) * dL_da +
		(T_(1, 0) * K_(2, 0) + T_(1, 1) * K_(2, 1) + T_(1, 2) * K_(2, 2)) * dL_db;
	float dL_dT10 = 2 * (T_(1, 0) * K_(0, 0) + T_(1, 1) * K_(0, 1) + T_(1, 2) * K_(0, 2)) * dL_dc +
		(T_(0, 0) * K_(0, 0) + T_(0, 1) * K_(0, 1) + T_(0, 2) * K_(0, 2)) * dL_db;
	float dL_dT11 = 2 * (T_(1, 0) * K_(1, 0) + T_(1, 1) * K_(1, 1) + T_(1, 2) * K_(1, 2)) * dL_dc +
		(T_(0, 0) * K_(1, 0) + T_(0, 1) * K_(1, 1) + T_(0, 2) * K_(1, 2)) * dL_db;
	float dL_dT12 = 2 * (T_(1, 0) * K_(2, 0) + T_(1, 1) * K_(2, 1) + T_(1, 2) * K_(2, 2)) * dL_dc +
		(T_(0, 0) * K_(2, 0) + T_(0, 1) * K_(2, 1) + T_(0, 2) * K_(2, 2)) * dL_db;
#undef K_
#undef T_

#define W_(i, j) Wm.c[i][j]
	dL_dJ00 = W_(0, 0) * dL_dT00 + W_(0, 1) * dL_dT01 + W_(0, 2) * dL_dT02;
	dL_dJ02 = W_(2, 0) * dL_dT00 + W_(2, 1) * dL_dT01 + W_(2, 2) * dL_dT02;
	dL_dJ10 = W_(0, 0) * dL_dT10 + W_(0, 1) * dL_dT11 + W_(0, 2) * dL_dT12;
	dL_dJ11 = W_(1, 0) * dL_dT10 + W_(1, 1) * dL_dT11 + W_(1, 2) * dL_dT12;
	dL_dJ12 = W_(2, 0) * dL_dT10 + W_(2, 1) * dL_dT11 + W_(2, 2) * dL_dT12;
#undef W_
}

// Backward through conic = inverse(cov2D), cov2D = T^T Vrk T and the lonlat Jacobian, including
// the second-derivative terms of the projection (backward.cu:297-485).
// Outputs: dL/dcov3D (6), the covariance branch of dL/dmean (world), and the Jacobian rows.
OGS_D void cov2d_lonlat_backward(float3 mean, const float* cov6, const float* V, int W, int H,
                                 float3 dL_dconic /*(.x,.y,.w of the accumulator)*/,
                                 float* dL_dcov6, float3& dL_dmean, float3& dpx_dt, float3& dpy_dt)
{
	float3 t = view_point(V, mean);

	float txtx = t.x * t.x;
	float tyty = t.y * t.y;
	float tztz = t.z * t.z;
	float txtytz = t.x * t.y * t.z;

	float trxztrxz = txtx + tztz;
	float trxztrxz_inv = 1.0f / (trxztrxz + kEps7);
	float trxztrxztrxztrxz_inv = trxztrxz_inv * trxztrxz_inv;
	float trxz = sqrtf(trxztrxz);
	float trxz_inv = 1.0f / (trxz + kEps7);
	float trtr = trxztrxz + tyty;
	float trtr_inv = 1.0f / (trtr + kEps7);
	float trtrtrtr_inv = trtr_inv * trtr_inv;
	float trxz_trtrtrtr_inv = trxz_inv * trtrtrtr_inv;
	float trxztrxztrxz_trtrtrtr_inv = trxztrxz_inv * trxz_trtrtrtr_inv;
	float tyty_minus_trxztrxz = tyty - trxztrxz;

	float W_div_2pi = W * 0.5f * kPiInv;
	float H_div_pi = H * kPiInv;

	LonlatJac Jv;
	Jv.j00 = W_div_2pi * t.z * trxztrxz_inv;
	Jv.j02 = -W_div_2pi * t.x * trxztrxz_inv;
	Jv.j10 = -H_div_pi * t.x * t.y * trxz_inv * trtr_inv;
	Jv.j11 = H_div_pi * trxz * trtr_inv;
	Jv.j12 = -H_div_pi * t.z * t.y * trxz_inv * trtr_inv;

	dpx_dt = { Jv.j00, 0.0f, Jv.j02 };
	dpy_dt = { Jv.j10, Jv.j11, Jv.j12 };

	float dL_dJ00, dL_dJ02, dL_dJ10, dL_dJ11, dL_dJ12;
	conic_to_cov3d_and_jacobian_backward(Jv, V, cov6, dL_dconic, dL_dcov6, dL_dJ00, dL_dJ02, dL_dJ10, dL_dJ11, dL_dJ12);

	float temp1 = H_div_pi * tyty_minus_trxztrxz * trxz_trtrtrtr_inv;
	float temp2 = H_div_pi * txtytz * (trtr + 2.0f * trxztrxz) * trxztrxztrxz_trtrtrtr_inv;
	float temp3 = W_div_2pi * (txtx - tztz) * trxztrxztrxztrxz_inv;
	float temp4 = W_div_2pi * 2.0f * t.x * t.z * trxztrxztrxztrxz_inv;
	float temp5 = H_div_pi * t.y * trxztrxztrxz_trtrtrtr_inv;

	float dL_dtx = -dL_dJ00 * temp4
		+ dL_dJ02 * temp3
		+ dL_dJ10 * temp5 * (2.0f * txtx * trxztrxz - tztz * trtr)
		+ dL_dJ11 * t.x * temp1
		+ dL_dJ12 * temp2;

	float dL_dty = dL_dJ10 * t.x * temp1
		- dL_dJ11 * H_div_pi * 2.0f * trxz * t.y * trtrtrtr_inv
		+ dL_dJ12 * t.z * temp1;

	float dL_dtz = dL_dJ00 * temp3
		+ dL_dJ02 * temp4
		+ dL_dJ10 * temp2
		+ dL_dJ11 * t.z * temp1
		+ dL_dJ12 * temp5 * (2.0f * tztz * trxztrxz - txtx * trtr);

	dL_dmean = view_vec_t(V, float3{ dL_dtx, dL_dty, dL_dtz });
}

} // namespace ogs
