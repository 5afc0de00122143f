// gaussian_grad.cuh — per-Gaussian backward math (vector-Jacobian products), written from the chain rule in matrix
// form, not from the reference's statement list.  Host + device, templated on the scalar type: the kernels instantiate
// float, tests/test_grad_math_cpu.py instantiates double as ground truth and float beside the reference restatement
// (oracle/lonlat_oracle.c) to show both float evaluations sit at the same distance from it.
//
// What it must equal (tolerance-bound, gradients only): computeCov2DLonLatCUDA (reference cuda_rasterizer/backward.cu:
// 297-485), computeCov2DCUDA (:156-292), computeColorFromSH backward (:30-151), computeCov3D backward (:489-552),
// dnormvdv (auxiliary.h:134-144), the screen-position branches of preprocessLonLatCUDA / preprocessCUDA (:642-660, :583-597).
//
// Notation.  Rcw(i,k) = V[4k+i] (viewmatrix is Tcw transposed).  t = Rcw p + tcw.  J (2x3) = d(pixel)/d(t).
// A = J Rcw (2x3) = d(pixel)/d(world).  cov2D = A Sigma A^T + 0.3 I.  conic = cov2D^-1.
// With S = dL/dcov2D (symmetric 2x2):   dL/dSigma = A^T S A,   dL/dA = 2 S (A Sigma),   dL/dJ = dL/dA Rcw^T.
// The structural zeros (J01 = 0 for both cameras, J10 = 0 for the pinhole) are never multiplied.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#ifndef OGS_HD
#define OGS_HD __host__ __device__ __forceinline__
#endif

namespace ogs {
namespace grad {

template <typename F>
struct Vec3 {
	F x, y, z;
};
template <typename F>
struct Rows23 {   // a 2x3 matrix by rows
	F a[3], b[3];
};
template <typename F>
struct Sym2 {     // symmetric 2x2: [[xx, xy], [xy, yy]]
	F xx, xy, yy;
};

template <typename F>
OGS_HD F eps7() { return (F)0.0000001f; }

// t = Rcw p + tcw
template <typename F>
OGS_HD Vec3<F> to_camera(const F* V, Vec3<F> p)
{
	return { V[0] * p.x + V[4] * p.y + V[8] * p.z + V[12],
	         V[1] * p.x + V[5] * p.y + V[9] * p.z + V[13],
	         V[2] * p.x + V[6] * p.y + V[10] * p.z + V[14] };
}
// Rcw^T d: a camera-space gradient carried back to world space
template <typename F>
OGS_HD Vec3<F> to_world_grad(const F* V, Vec3<F> d)
{
	return { V[0] * d.x + V[1] * d.y + V[2] * d.z,
	         V[4] * d.x + V[5] * d.y + V[6] * d.z,
	         V[8] * d.x + V[9] * d.y + V[10] * d.z };
}

// ------------------------------------------------------------------ equirectangular projection, local quantities
// px = (W / 2pi) atan2(tx, tz) + ..., py = (H / pi) asin(ty / r) + ...; the reference guards every denominator with
// +1e-7 and differentiates with the guarded reciprocals as constants (forward.cu:147-162, backward.cu:331-368): kept.
template <typename F>
struct LonlatLocal {
	F tx, ty, tz;
	F a, rho, rr;        // tx^2 + tz^2, its root, a + ty^2
	F ia, irho, irr;     // 1/(a + eps), 1/(rho + eps), 1/(rr + eps)
	F Wd, Hd;            // W / 2pi, H / pi
};
template <typename F>
OGS_HD LonlatLocal<F> lonlat_local(Vec3<F> t, int W, int H)
{
	LonlatLocal<F> L;
	L.tx = t.x; L.ty = t.y; L.tz = t.z;
	L.a = t.x * t.x + t.z * t.z;
	L.rho = sqrt(L.a);
	L.rr = L.a + t.y * t.y;
	L.ia = (F)1 / (L.a + eps7<F>());
	L.irho = (F)1 / (L.rho + eps7<F>());
	L.irr = (F)1 / (L.rr + eps7<F>());
	L.Wd = (F)W * (F)0.5f * (F)0.318309886183790671537767526745028724f;
	L.Hd = (F)H * (F)0.318309886183790671537767526745028724f;
	return L;
}
// rows of J: a = d(px)/dt = Wd (tz, 0, -tx) / a,  b = d(py)/dt = Hd (-tx ty / rho, rho, -tz ty / rho) / rr
template <typename F>
OGS_HD Rows23<F> lonlat_jacobian_rows(const LonlatLocal<F>& L)
{
	Rows23<F> J;
	const F wa = L.Wd * L.ia, hb = L.Hd * L.irr, hc = hb * L.irho * L.ty;
	J.a[0] = wa * L.tz; J.a[1] = (F)0; J.a[2] = -wa * L.tx;
	J.b[0] = -hc * L.tx; J.b[1] = hb * L.rho; J.b[2] = -hc * L.tz;
	return J;
}
// dL/dt through the t-dependence of J (the projection's second derivatives), given g = dL/dJ (g.a[1] is unused: J01 = 0).
//   d(a-row)/dt:  Wd/a^2 * [ (-2 tx tz, tx^2 - tz^2) ; (tx^2 - tz^2, 2 tx tz) ]      (rows: J00, J02; columns: tx, tz)
//   d(b-row)/dt:  with k = Hd / (rho rr^2), e = ty^2 - a, m = tx ty tz (rr + 2a) / a:
//       dJ10 = k (ty (2 tx^2 a - tz^2 rr) / a,  tx e,  m)
//       dJ11 = (k tx e,  -2 Hd rho ty / rr^2,  k tz e)
//       dJ12 = k (m,  tz e,  ty (2 tz^2 a - tx^2 rr) / a)
template <typename F>
OGS_HD Vec3<F> lonlat_jacobian_vjp(const LonlatLocal<F>& L, const Rows23<F>& g)
{
	const F xx = L.tx * L.tx, zz = L.tz * L.tz, xz2 = (F)2 * L.tx * L.tz;
	const F wq = L.Wd * L.ia * L.ia;
	const F k = L.Hd * L.irho * L.irr * L.irr;
	const F e = L.ty * L.ty - L.a;
	const F m = L.tx * L.ty * L.tz * (L.rr + (F)2 * L.a) * L.ia;
	const F ty_ia = L.ty * L.ia;
	const F d10x = ty_ia * ((F)2 * xx * L.a - zz * L.rr);
	const F d12z = ty_ia * ((F)2 * zz * L.a - xx * L.rr);
	Vec3<F> d;
	d.x = wq * ((xx - zz) * g.a[2] - xz2 * g.a[0]) + k * (d10x * g.b[0] + L.tx * e * g.b[1] + m * g.b[2]);
	d.y = k * e * (L.tx * g.b[0] + L.tz * g.b[2]) - (F)2 * L.Hd * L.rho * L.ty * L.irr * L.irr * g.b[1];
	d.z = wq * ((xx - zz) * g.a[0] + xz2 * g.a[2]) + k * (m * g.b[0] + L.tz * e * g.b[1] + d12z * g.b[2]);
	return d;
}

// ------------------------------------------------------------------ perspective projection (camera_type 1)
// J rows at the frustum-clamped point: a = (fx / tz, 0, -fx tx / tz^2), b = (0, fy / tz, -fy ty / tz^2); a clamped
// coordinate passes no gradient to tx / ty (forward.cu:94-108, backward.cu:179-196).
template <typename F>
struct PinholeLocal {
	F tx, ty, tz;        // clamped x, y
	F fx, fy;
	F pass_x, pass_y;    // 0 where the clamp was active
};
template <typename F>
OGS_HD PinholeLocal<F> pinhole_local(Vec3<F> t, F fx, F fy, F tan_fovx, F tan_fovy)
{
	PinholeLocal<F> L;
	const F limx = (F)1.3f * tan_fovx, limy = (F)1.3f * tan_fovy;
	const F rx = t.x / t.z, ry = t.y / t.z;
	L.pass_x = (rx < -limx || rx > limx) ? (F)0 : (F)1;
	L.pass_y = (ry < -limy || ry > limy) ? (F)0 : (F)1;
	L.tx = fmin(limx, fmax(-limx, rx)) * t.z;
	L.ty = fmin(limy, fmax(-limy, ry)) * t.z;
	L.tz = t.z; L.fx = fx; L.fy = fy;
	return L;
}
template <typename F>
OGS_HD Rows23<F> pinhole_jacobian_rows(const PinholeLocal<F>& L)
{
	const F iz = (F)1 / L.tz, iz2 = iz * iz;
	Rows23<F> J;
	J.a[0] = L.fx * iz; J.a[1] = (F)0; J.a[2] = -L.fx * L.tx * iz2;
	J.b[0] = (F)0; J.b[1] = L.fy * iz; J.b[2] = -L.fy * L.ty * iz2;
	return J;
}
template <typename F>
OGS_HD Vec3<F> pinhole_jacobian_vjp(const PinholeLocal<F>& L, const Rows23<F>& g)
{
	const F iz = (F)1 / L.tz, iz2 = iz * iz, iz3 = iz2 * iz;
	Vec3<F> d;
	d.x = -L.pass_x * L.fx * iz2 * g.a[2];
	d.y = -L.pass_y * L.fy * iz2 * g.b[2];
	d.z = -iz2 * (L.fx * g.a[0] + L.fy * g.b[1]) + (F)2 * iz3 * (L.fx * L.tx * g.a[2] + L.fy * L.ty * g.b[2]);
	return d;
}
// Screen position through the full projection: ndc = (h.x, h.y) / (h.w + eps), h = Pm (m, 1), Pm column-major.
// Returns Pm^T-projected gradient: d_j = ((Pm[4j] - ndc.x Pm[4j+3]) gx + (Pm[4j+1] - ndc.y Pm[4j+3]) gy) / (h.w + eps).
template <typename F>
OGS_HD Vec3<F> projection_vjp(const F* Pm, Vec3<F> m, F gx, F gy)
{
	const F hx = Pm[0] * m.x + Pm[4] * m.y + Pm[8] * m.z + Pm[12];
	const F hy = Pm[1] * m.x + Pm[5] * m.y + Pm[9] * m.z + Pm[13];
	const F hw = Pm[3] * m.x + Pm[7] * m.y + Pm[11] * m.z + Pm[15];
	const F iw = (F)1 / (hw + eps7<F>());
	const F nx = hx * iw, ny = hy * iw;
	const F ax = gx * iw, ay = gy * iw, aw = -(nx * ax + ny * ay);
	return { Pm[0] * ax + Pm[1] * ay + Pm[3] * aw, Pm[4] * ax + Pm[5] * ay + Pm[7] * aw, Pm[8] * ax + Pm[9] * ay + Pm[11] * aw };
}

// ------------------------------------------------------------------ covariance chain
// A = J Rcw
template <typename F>
OGS_HD Rows23<F> world_rows(const Rows23<F>& J, const F* V, bool a1_zero = true, bool b0_zero = false)
{
	Rows23<F> A;
#pragma unroll
	for (int k = 0; k < 3; k++) {
		A.a[k] = J.a[0] * V[4 * k] + J.a[2] * V[4 * k + 2] + (a1_zero ? (F)0 : J.a[1] * V[4 * k + 1]);
		A.b[k] = J.b[1] * V[4 * k + 1] + J.b[2] * V[4 * k + 2] + (b0_zero ? (F)0 : J.b[0] * V[4 * k]);
	}
	return A;
}
// U = A Sigma, Sigma symmetric with upper triangle c6 = (00, 01, 02, 11, 12, 22)
template <typename F>
OGS_HD Rows23<F> times_sigma(const Rows23<F>& A, const F* c6)
{
	Rows23<F> U;
	U.a[0] = A.a[0] * c6[0] + A.a[1] * c6[1] + A.a[2] * c6[2];
	U.a[1] = A.a[0] * c6[1] + A.a[1] * c6[3] + A.a[2] * c6[4];
	U.a[2] = A.a[0] * c6[2] + A.a[1] * c6[4] + A.a[2] * c6[5];
	U.b[0] = A.b[0] * c6[0] + A.b[1] * c6[1] + A.b[2] * c6[2];
	U.b[1] = A.b[0] * c6[1] + A.b[1] * c6[3] + A.b[2] * c6[4];
	U.b[2] = A.b[0] * c6[2] + A.b[1] * c6[4] + A.b[2] * c6[5];
	return U;
}
// cov2D = U A^T + 0.3 I
template <typename F>
OGS_HD Sym2<F> cov2d_of(const Rows23<F>& U, const Rows23<F>& A)
{
	Sym2<F> c;
	c.xx = U.a[0] * A.a[0] + U.a[1] * A.a[1] + U.a[2] * A.a[2] + (F)0.3f;
	c.xy = U.a[0] * A.b[0] + U.a[1] * A.b[1] + U.a[2] * A.b[2];
	c.yy = U.b[0] * A.b[0] + U.b[1] * A.b[1] + U.b[2] * A.b[2] + (F)0.3f;
	return c;
}
// conic = cov2D^-1  =>  dL/dcov2D = -conic G conic with G = [[gA, gB], [gB, gC]] in the reference's convention (its
// dL/dconic.y is the derivative w.r.t. ONE of the two off-diagonal entries: half the derivative w.r.t. the parameter B).
// Written on the adjugate: conic = adj / det, so S = -(adj G adj) / det^2, with the reference's guard det^2 + 1e-7 and its
// "no gradient when that reciprocal underflows to 0" rule (backward.cu:395-407).
template <typename F>
OGS_HD Sym2<F> cov2d_grad_from_conic_grad(const Sym2<F>& c, F gA, F gB, F gC)
{
	const F det = c.xx * c.yy - c.xy * c.xy;
	const F w = (F)1 / (det * det + eps7<F>());
	Sym2<F> S = { (F)0, (F)0, (F)0 };
	if (w != (F)0) {
		// rows of adj G, adj = [[yy, -xy], [-xy, xx]]
		const F p0 = c.yy * gA - c.xy * gB, p1 = c.yy * gB - c.xy * gC;
		const F q0 = c.xx * gB - c.xy * gA, q1 = c.xx * gC - c.xy * gB;
		S.xx = -w * (p0 * c.yy - p1 * c.xy);
		S.xy = -w * (p1 * c.xx - p0 * c.xy);
		S.yy = -w * (q1 * c.xx - q0 * c.xy);
	}
	return S;
}
// dL/dSigma = A^T S A in the 6-vector convention (off-diagonal entries count twice)
template <typename F>
OGS_HD void sigma_grad(const Rows23<F>& A, const Sym2<F>& S, F* d6)
{
	F v[3], w[3];   // S A by columns
#pragma unroll
	for (int k = 0; k < 3; k++) {
		v[k] = S.xx * A.a[k] + S.xy * A.b[k];
		w[k] = S.xy * A.a[k] + S.yy * A.b[k];
	}
	d6[0] = A.a[0] * v[0] + A.b[0] * w[0];
	d6[3] = A.a[1] * v[1] + A.b[1] * w[1];
	d6[5] = A.a[2] * v[2] + A.b[2] * w[2];
	d6[1] = (F)2 * (A.a[0] * v[1] + A.b[0] * w[1]);
	d6[2] = (F)2 * (A.a[0] * v[2] + A.b[0] * w[2]);
	d6[4] = (F)2 * (A.a[1] * v[2] + A.b[1] * w[2]);
}
// dL/dJ = 2 S U Rcw^T
template <typename F>
OGS_HD Rows23<F> jacobian_grad(const Rows23<F>& U, const Sym2<F>& S, const F* V)
{
	F da[3], db[3];
#pragma unroll
	for (int k = 0; k < 3; k++) {
		da[k] = (F)2 * (S.xx * U.a[k] + S.xy * U.b[k]);
		db[k] = (F)2 * (S.xy * U.a[k] + S.yy * U.b[k]);
	}
	Rows23<F> g;
#pragma unroll
	for (int i = 0; i < 3; i++) {
		g.a[i] = da[0] * V[i] + da[1] * V[4 + i] + da[2] * V[8 + i];
		g.b[i] = db[0] * V[i] + db[1] * V[4 + i] + db[2] * V[8 + i];
	}
	return g;
}

// ------------------------------------------------------------------ Sigma = Rm diag(s)^2 Rm^T
// Rm = rotation of the (un-normalised) quaternion (r, x, y, z); s = scale_modifier * scale.
// With D = dL/dSigma as a symmetric matrix (off-diagonals halved) and E = D Rm:
//   dL/ds_k = 2 s_k (Rm^T E)_kk        (the reference returns this derivative w.r.t. s, not w.r.t. scale: kept)
//   dL/dRm  = 2 E diag(s)^2 =: Q,  dL/dq = <Q, dRm/dq>.
template <typename F>
OGS_HD void scale_rotation_grad(const F* s, const F* q, const F* d6, F* ds, F* dq)
{
	const F r = q[0], x = q[1], y = q[2], z = q[3];
	const F Rm[3][3] = {
		{ (F)1 - (F)2 * (y * y + z * z), (F)2 * (x * y - r * z), (F)2 * (x * z + r * y) },
		{ (F)2 * (x * y + r * z), (F)1 - (F)2 * (x * x + z * z), (F)2 * (y * z - r * x) },
		{ (F)2 * (x * z - r * y), (F)2 * (y * z + r * x), (F)1 - (F)2 * (x * x + y * y) } };
	const F D[3][3] = { { d6[0], (F)0.5f * d6[1], (F)0.5f * d6[2] },
	                    { (F)0.5f * d6[1], d6[3], (F)0.5f * d6[4] },
	                    { (F)0.5f * d6[2], (F)0.5f * d6[4], d6[5] } };
	F Q[3][3];
#pragma unroll
	for (int k = 0; k < 3; k++) {
		F e[3];
#pragma unroll
		for (int j = 0; j < 3; j++) e[j] = D[j][0] * Rm[0][k] + D[j][1] * Rm[1][k] + D[j][2] * Rm[2][k];
		ds[k] = (F)2 * s[k] * (Rm[0][k] * e[0] + Rm[1][k] * e[1] + Rm[2][k] * e[2]);
		const F w = (F)2 * s[k] * s[k];
#pragma unroll
		for (int j = 0; j < 3; j++) Q[j][k] = w * e[j];
	}
	// dRm/dr = 2 [[0,-z,y],[z,0,-x],[-y,x,0]]        dRm/dx = 2 [[0,y,z],[y,-2x,-r],[z,r,-2x]]
	// dRm/dy = 2 [[-2y,x,r],[x,0,z],[-r,z,-2y]]      dRm/dz = 2 [[-2z,-r,x],[r,-2z,y],[x,y,0]]
	const F s01 = Q[0][1] + Q[1][0], s02 = Q[0][2] + Q[2][0], s12 = Q[1][2] + Q[2][1];
	const F a01 = Q[1][0] - Q[0][1], a02 = Q[0][2] - Q[2][0], a12 = Q[2][1] - Q[1][2];
	dq[0] = (F)2 * (z * a01 + y * a02 + x * a12);
	dq[1] = (F)2 * (y * s01 + z * s02 + r * a12) - (F)4 * x * (Q[1][1] + Q[2][2]);
	dq[2] = (F)2 * (x * s01 + r * a02 + z * s12) - (F)4 * y * (Q[0][0] + Q[2][2]);
	dq[3] = (F)2 * (r * a01 + x * s02 + y * s12) - (F)4 * z * (Q[0][0] + Q[1][1]);
}

// ------------------------------------------------------------------ spherical harmonics
// colour = 0.5 + sum_k b_k(dir) sh_k (clamped at 0), dir = (p - c) / |p - c|.  The 16 real SH weights b_k with the
// reference's constants and signs (forward.cu:30-83), and the gradient of sum_k w_k b_k w.r.t. dir.
template <typename F>
struct ShConst {
	static OGS_HD F c0() { return (F)0.28209479177387814f; }
	static OGS_HD F c1() { return (F)0.4886025119029199f; }
	static OGS_HD F c2(int i)
	{
		return i == 2 ? (F)0.31539156525252005f : (i == 4 ? (F)0.5462742152960396f : (i == 0 ? (F)1.0925484305920792f : (F)-1.0925484305920792f));
	}
	static OGS_HD F c3(int i)
	{
		return (i == 0 || i == 6) ? (F)-0.5900435899266435f
		     : (i == 1) ? (F)2.890611442640554f
		     : (i == 2 || i == 4) ? (F)-0.4570457994644658f
		     : (i == 3) ? (F)0.3731763325901154f : (F)1.445305721320277f;
	}
};
// Every product below is a single rounding (no contraction across statements): the data-parallel path rebuilds
// dL/dsh = b_k * dL/dRGB per view in another kernel and must get the same bits (parallel.py, sh_gradient_from_views).
template <typename F>
OGS_HD F mul1(F a, F b)
{
#if defined(__CUDA_ARCH__)
	if (sizeof(F) == 4) return (F)__fmul_rn((float)a, (float)b);
#endif
	return a * b;
}
template <typename F>
OGS_HD F sub1(F a, F b)
{
#if defined(__CUDA_ARCH__)
	if (sizeof(F) == 4) return (F)__fsub_rn((float)a, (float)b);
#endif
	return a - b;
}
template <typename F>
OGS_HD void sh_weights(int deg, F x, F y, F z, F* b /*16*/)
{
	typedef ShConst<F> C;
#pragma unroll
	for (int k = 0; k < 16; k++) b[k] = (F)0;
	b[0] = C::c0();
	if (deg < 1) return;
	b[1] = -mul1(C::c1(), y); b[2] = mul1(C::c1(), z); b[3] = -mul1(C::c1(), x);
	if (deg < 2) return;
	const F xx = mul1(x, x), yy = mul1(y, y), zz = mul1(z, z), xy = mul1(x, y), yz = mul1(y, z), xz = mul1(x, z);
	const F zz2 = mul1((F)2, zz), xmy = sub1(xx, yy);
	b[4] = mul1(C::c2(0), xy);
	b[5] = mul1(C::c2(1), yz);
	b[6] = mul1(C::c2(2), sub1(sub1(zz2, xx), yy));
	b[7] = mul1(C::c2(3), xz);
	b[8] = mul1(C::c2(4), xmy);
	if (deg < 3) return;
	const F u = sub1(sub1(mul1((F)4, zz), xx), yy);                      // 4zz - xx - yy
	b[9] = mul1(mul1(C::c3(0), y), sub1(mul1((F)3, xx), yy));
	b[10] = mul1(mul1(C::c3(1), xy), z);
	b[11] = mul1(mul1(C::c3(2), y), u);
	b[12] = mul1(mul1(C::c3(3), z), sub1(sub1(zz2, mul1((F)3, xx)), mul1((F)3, yy)));
	b[13] = mul1(mul1(C::c3(4), x), u);
	b[14] = mul1(mul1(C::c3(5), z), xmy);
	b[15] = mul1(mul1(C::c3(6), x), sub1(xx, mul1((F)3, yy)));
}
// grad_dir sum_k w_k b_k(dir)
template <typename F>
OGS_HD Vec3<F> sh_weights_vjp(int deg, F x, F y, F z, const F* w /*16*/)
{
	typedef ShConst<F> C;
	Vec3<F> g = { (F)0, (F)0, (F)0 };
	if (deg < 1) return g;
	g.x = -C::c1() * w[3]; g.y = -C::c1() * w[1]; g.z = C::c1() * w[2];
	if (deg < 2) return g;
	const F q4 = C::c2(0) * w[4], q5 = C::c2(1) * w[5], q6 = C::c2(2) * w[6], q7 = C::c2(3) * w[7], q8 = C::c2(4) * w[8];
	g.x += y * q4 + z * q7 + (F)2 * x * (q8 - q6);
	g.y += x * q4 + z * q5 - (F)2 * y * (q6 + q8);
	g.z += y * q5 + x * q7 + (F)4 * z * q6;
	if (deg < 3) return g;
	const F xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
	const F q9 = C::c3(0) * w[9], q10 = C::c3(1) * w[10], q11 = C::c3(2) * w[11], q12 = C::c3(3) * w[12];
	const F q13 = C::c3(4) * w[13], q14 = C::c3(5) * w[14], q15 = C::c3(6) * w[15];
	g.x += xy * ((F)6 * q9 - (F)2 * q11) + yz * q10 + xz * ((F)2 * q14 - (F)6 * q12) + ((F)4 * zz - (F)3 * xx - yy) * q13
	       + (F)3 * (xx - yy) * q15;
	g.y += (F)3 * (xx - yy) * q9 + xz * q10 + ((F)4 * zz - xx - (F)3 * yy) * q11 - yz * ((F)6 * q12 + (F)2 * q14)
	       - xy * ((F)2 * q13 + (F)6 * q15);
	g.z += xy * q10 + (F)8 * (yz * q11 + xz * q13) + (F)3 * ((F)2 * zz - xx - yy) * q12 + (xx - yy) * q14;
	return g;
}
// d(v / |v|)/dv applied to g:  (g |v|^2 - v (v . g)) / |v|^3
template <typename F>
OGS_HD Vec3<F> normalize_vjp(Vec3<F> v, Vec3<F> g)
{
	const F n2 = v.x * v.x + v.y * v.y + v.z * v.z;
	const F i3 = (F)1 / sqrt(n2 * n2 * n2);
	const F vg = v.x * g.x + v.y * g.y + v.z * g.z;
	return { (g.x * n2 - v.x * vg) * i3, (g.y * n2 - v.y * vg) * i3, (g.z * n2 - v.z * vg) * i3 };
}

// ------------------------------------------------------------------ the per-Gaussian chains the kernel (and the CPU test) call
// Covariance / conic branch and screen-position branch.  (g_mx, g_my) = dL/dmean2D in the reference's convention
// (pixel derivative times W/2, H/2: backward.cu:821-826), (gA, gB, gC) = dL/dconic (.x, .y, .w).
// Fills d6 = dL/dcov3D and returns dL/dmean (world).  Pm, fx .. tan_fovy are used by the perspective camera only.
template <typename F, bool kPinhole>
OGS_HD Vec3<F> projection_backward(Vec3<F> mean, const F* c6, const F* V, const F* Pm, int W, int H, F fx, F fy, F tan_fovx,
                                   F tan_fovy, F g_mx, F g_my, F gA, F gB, F gC, F* d6)
{
	const Vec3<F> t = to_camera(V, mean);
	if (kPinhole) {
		const PinholeLocal<F> L = pinhole_local(t, fx, fy, tan_fovx, tan_fovy);
		const Rows23<F> J = pinhole_jacobian_rows(L);
		const Rows23<F> A = world_rows(J, V, true, true);
		const Rows23<F> U = times_sigma(A, c6);
		const Sym2<F> S = cov2d_grad_from_conic_grad(cov2d_of(U, A), gA, gB, gC);
		sigma_grad(A, S, d6);
		const Vec3<F> dm = to_world_grad(V, pinhole_jacobian_vjp(L, jacobian_grad(U, S, V)));
		const Vec3<F> ds = projection_vjp(Pm, mean, g_mx, g_my);
		return { dm.x + ds.x, dm.y + ds.y, dm.z + ds.z };
	} else {
		const LonlatLocal<F> L = lonlat_local(t, W, H);
		const Rows23<F> J = lonlat_jacobian_rows(L);
		const Rows23<F> A = world_rows(J, V);
		const Rows23<F> U = times_sigma(A, c6);
		const Sym2<F> S = cov2d_grad_from_conic_grad(cov2d_of(U, A), gA, gB, gC);
		sigma_grad(A, S, d6);
		Vec3<F> dt = lonlat_jacobian_vjp(L, jacobian_grad(U, S, V));
		// screen position: dL/dpixel = dL/dmean2D * (2/W, 2/H), carried through the rows of J
		const F px = g_mx * ((F)2 / (F)W), py = g_my * ((F)2 / (F)H);
		dt.x += px * J.a[0] + py * J.b[0];
		dt.y += py * J.b[1];
		dt.z += px * J.a[2] + py * J.b[2];
		return to_world_grad(V, dt);
	}
}

// Colour branch.  sh(k, c) reads coefficient k, channel c; dsh(k, c, v) receives dL/dsh (only k < (deg+1)^2 are produced;
// pass a no-op to skip them).  dRGB is dL/dcolour with clamped channels already zeroed.  Returns dL/dmean (world).
template <typename F, typename Load, typename Store>
OGS_HD Vec3<F> colour_backward(int deg, Vec3<F> mean, Vec3<F> cam, Load sh, const F* dRGB, Store dsh)
{
	const Vec3<F> v = { mean.x - cam.x, mean.y - cam.y, mean.z - cam.z };
	const F len = sqrt(v.x * v.x + v.y * v.y + v.z * v.z);
	const F x = v.x / len, y = v.y / len, z = v.z / len;
	F b[16], w[16];
	sh_weights(deg, x, y, z, b);
	const int n = (deg + 1) * (deg + 1);
#pragma unroll
	for (int k = 0; k < 16; k++) {
		w[k] = (F)0;
		if (k < n) {
			w[k] = sh(k, 0) * dRGB[0] + sh(k, 1) * dRGB[1] + sh(k, 2) * dRGB[2];
			dsh(k, 0, mul1(b[k], dRGB[0]));
			dsh(k, 1, mul1(b[k], dRGB[1]));
			dsh(k, 2, mul1(b[k], dRGB[2]));
		}
	}
	return normalize_vjp(v, sh_weights_vjp(deg, x, y, z, w));
}

} // namespace grad
} // namespace ogs
