// lonlat_math.cuh — per-Gaussian equirectangular projection math of the FORWARD pass.
//
// Bit-exactness contract: radii, tile rects, depth keys, conics and colours must equal the reference's bit for
// bit, so every operation below is an explicit round-to-nearest intrinsic in the order the reference's compiled
// kernel executes it (see "pinned forward chain").  Reference lines are cited per function (raikuma/OmniGS-fork,
// cuda_rasterizer/).  The backward pass is tolerance-bound and lives in gaussian_grad.cuh in our own form.
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
OGS_HD M3 m3_t(const M3& A)
{
	return m3_cols(A.c[0][0], A.c[1][0], A.c[2][0],
	               A.c[0][1], A.c[1][1], A.c[2][1],
	               A.c[0][2], A.c[1][2], A.c[2][2]);
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

// SH -> RGB for one Gaussian (forward.cu:30-83); `sh(k)` returns coefficient k, the clamp mask comes back in bits 0..2.
// Every operation pinned to the order of the reference's compiled computeColorFromSH
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

} // namespace ogs
