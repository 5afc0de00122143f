/*
 * lonlat_oracle.c — CPU restatement (plain C99, float32, optional OpenMP) of the reference's
 * lonlat rasterizer path.  TEST INFRASTRUCTURE — see lonlat_oracle.h for who may use it.
 *
 * Conventions carried over from the reference:
 *   - viewmatrix V is Tcw stored column-major: t = (V0 x+V4 y+V8 z+V12, ...)   auxiliary.h:85-93
 *   - glm matrices are column-major, m[col][row]; mat3(a,b,c,...) fills columns;
 *     (A*B)[c][r] = A[0][r]*B[c][0] + A[1][r]*B[c][1] + A[2][r]*B[c][2]
 *   - everything is float32 except ndc2Pix, which the reference evaluates in double
 *     (auxiliary.h:51-54).
 * Parity: pinned against tests/golden/*.npz (outputs of the unmodified reference on a B200).
 * CPU libm (atan2f/asinf/expf) may differ from CUDA libdevice in the last ulp, so integer
 * outputs can flip on rare ceil/truncation boundaries; the golden test bounds that.
 */
#include "lonlat_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define TILE 16          /* config.h:26-27  BLOCK_X = BLOCK_Y = 16 */
#define EPS7 0.0000001f  /* the reference's 1e-7 guards */

static const float PI_INV = 0.318309886183790671537767526745028724f;     /* M_1_PIf32 */
static const float TWO_PI_INV = 0.636619772367581343075535053490057448f; /* M_2_PIf32 */

/* auxiliary.h:32-49 */
static const float SH_C0 = 0.28209479177387814f;
static const float SH_C1 = 0.4886025119029199f;
static const float SH_C2[5] = { 1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                                -1.0925484305920792f, 0.5462742152960396f };
static const float SH_C3[7] = { -0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                                0.3731763325901154f, -0.4570457994644658f, 1.445305721320277f,
                                -0.5900435899266435f };

static int g_wrap = 0;
void ogs_oracle_set_seam_wrap(int on) { g_wrap = on ? 1 : 0; }

/* threads of the following calls (OpenMP builds; a no-op otherwise): bench.py's single-thread CPU row */
void ogs_oracle_set_num_threads(int n)
{
#ifdef _OPENMP
	if (n > 0) omp_set_num_threads(n);
#else
	(void)n;
#endif
}

int ogs_oracle_num_threads(void)
{
#ifdef _OPENMP
	return omp_get_max_threads();
#else
	return 1;
#endif
}

/* ---------------------------------------------------------------- glm-style helpers */
typedef struct { float c[3][3]; } m3; /* c[col][row] */

static m3 m3_cols(float a0, float a1, float a2, float b0, float b1, float b2, float c0, float c1, float c2)
{
	m3 m;
	m.c[0][0] = a0; m.c[0][1] = a1; m.c[0][2] = a2;
	m.c[1][0] = b0; m.c[1][1] = b1; m.c[1][2] = b2;
	m.c[2][0] = c0; m.c[2][1] = c1; m.c[2][2] = c2;
	return m;
}
static m3 m3_mul(const m3* A, const m3* B)
{
	m3 R;
	for (int c = 0; c < 3; c++)
		for (int r = 0; r < 3; r++)
			R.c[c][r] = A->c[0][r] * B->c[c][0] + A->c[1][r] * B->c[c][1] + A->c[2][r] * B->c[c][2];
	return R;
}
static m3 m3_t(const m3* A)
{
	return m3_cols(A->c[0][0], A->c[1][0], A->c[2][0],
	               A->c[0][1], A->c[1][1], A->c[2][1],
	               A->c[0][2], A->c[1][2], A->c[2][2]);
}

/* auxiliary.h:85-93 transformPoint4x3 */
static void view_point(const float* V, const float* p, float* t)
{
	t[0] = V[0] * p[0] + V[4] * p[1] + V[8] * p[2] + V[12];
	t[1] = V[1] * p[0] + V[5] * p[1] + V[9] * p[2] + V[13];
	t[2] = V[2] * p[0] + V[6] * p[1] + V[10] * p[2] + V[14];
}
/* auxiliary.h:116-124 transformVec4x3Transpose */
static void view_vec_transpose(const float* V, const float* p, float* o)
{
	o[0] = V[0] * p[0] + V[1] * p[1] + V[2] * p[2];
	o[1] = V[4] * p[0] + V[5] * p[1] + V[6] * p[2];
	o[2] = V[8] * p[0] + V[9] * p[1] + V[10] * p[2];
}

/* auxiliary.h:51-54 ndc2Pix — double arithmetic, narrowed on return */
static float ndc_to_pix(float v, int S)
{
	return (float)(((v + 1.0) * S - 1.0) * 0.5);
}

static int imin(int a, int b) { return a < b ? a : b; }
static int imax(int a, int b) { return a > b ? a : b; }

/* auxiliary.h:56-66 getRect (clamping variant — the live one) */
static void tile_rect(float px, float py, int max_radius, int gx, int gy, int* x0, int* y0, int* x1, int* y1)
{
	*x0 = imin(gx, imax(0, (int)((px - max_radius) / TILE)));
	*y0 = imin(gy, imax(0, (int)((py - max_radius) / TILE)));
	*x1 = imin(gx, imax(0, (int)((px + max_radius + TILE - 1) / TILE)));
	*y1 = imin(gy, imax(0, (int)((py + max_radius + TILE - 1) / TILE)));
}

/* wrap-around variant of the x range (extension, see ogs_oracle_set_seam_wrap): x0 in [0,gx), x1 = x0 + width */
static void tile_rect_any(float px, float py, int max_radius, int gx, int gy, int* x0, int* y0, int* x1, int* y1)
{
	tile_rect(px, py, max_radius, gx, gy, x0, y0, x1, y1);
	if (g_wrap) {
		const float R = (float)max_radius;
		/* getRect's expressions (auxiliary.h:59,63) with floor instead of clamp-to-zero truncation */
		int xa = (int)floorf((px - R) / TILE);
		int xb = (int)floorf((px + R + TILE - 1) / TILE);
		int w = imin(xb - xa, gx);
		*x0 = ((xa % gx) + gx) % gx;
		*x1 = *x0 + w;
	}
}
static float nearest_copy_x(float mx, float tile_cx, float Wf)
{
	const float d = mx - tile_cx;
	if (!g_wrap) return mx;
	return d > 0.5f * Wf ? mx - Wf : (d < -0.5f * Wf ? mx + Wf : mx);
}

/* rasterizer_impl.cu:47-62 */
uint32_t ogs_oracle_higher_msb(uint32_t n)
{
	uint32_t msb = sizeof(n) * 4;
	uint32_t step = msb;
	while (step > 1) {
		step /= 2;
		if (n >> msb) msb += step; else msb -= step;
	}
	if (n >> msb) msb++;
	return msb;
}

/* forward.cu:194-228 computeCov3D (quaternion NOT normalised, :203) */
static void rot_matrix(const float* q, m3* R)
{
	float r = q[0], x = q[1], y = q[2], z = q[3];
	*R = m3_cols(1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y),
	             2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x),
	             2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y));
}
static void cov3d_from_scale_rot(const float* scale, float mod, const float* q, float* cov6)
{
	m3 S = m3_cols(1.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 1.f);
	S.c[0][0] = mod * scale[0];
	S.c[1][1] = mod * scale[1];
	S.c[2][2] = mod * scale[2];
	m3 R; rot_matrix(q, &R);
	m3 M = m3_mul(&S, &R);
	m3 Mt = m3_t(&M);
	m3 Sigma = m3_mul(&Mt, &M);
	cov6[0] = Sigma.c[0][0]; cov6[1] = Sigma.c[0][1]; cov6[2] = Sigma.c[0][2];
	cov6[3] = Sigma.c[1][1]; cov6[4] = Sigma.c[1][2]; cov6[5] = Sigma.c[2][2];
}

/* The lonlat Jacobian entries — forward.cu:147-162 (and backward.cu:340-360) */
typedef struct { float j00, j02, j10, j11, j12; } lonlat_jac;
static lonlat_jac lonlat_jacobian(const float* t, int W, int H)
{
	lonlat_jac J;
	float a = t[0] * t[0] + t[2] * t[2];
	float a_inv = 1.0f / (a + EPS7);
	float rho = sqrtf(a);
	float rho_inv = 1.0f / (rho + EPS7);
	float rr = a + t[1] * t[1];
	float rr_inv = 1.0f / (rr + EPS7);
	float Wd = W * 0.5f * PI_INV;
	float Hd = H * PI_INV;
	J.j00 = Wd * t[2] * a_inv;
	J.j02 = -Wd * t[0] * a_inv;
	J.j10 = -Hd * t[0] * t[1] * rho_inv * rr_inv;
	J.j11 = Hd * rho * rr_inv;
	J.j12 = -Hd * t[2] * t[1] * rho_inv * rr_inv;
	return J;
}

/* T = W*J and cov = T^T Vrk^T T — forward.cu:164-181 */
static void lonlat_T_and_cov(const lonlat_jac* Jv, const float* V, const float* cov6, m3* T, m3* Vrk, m3* cov)
{
	m3 J = m3_cols(Jv->j00, 0.0f, Jv->j02, Jv->j10, Jv->j11, Jv->j12, 0.0f, 0.0f, 0.0f);
	m3 Wm = m3_cols(V[0], V[4], V[8], V[1], V[5], V[9], V[2], V[6], V[10]);
	*T = m3_mul(&Wm, &J);
	*Vrk = m3_cols(cov6[0], cov6[1], cov6[2], cov6[1], cov6[3], cov6[4], cov6[2], cov6[4], cov6[5]);
	m3 Tt = m3_t(T), Vt = m3_t(Vrk);
	m3 tmp = m3_mul(&Tt, &Vt);
	*cov = m3_mul(&tmp, T);
}

/* forward.cu:30-83 computeColorFromSH */
static void sh_to_rgb(int idx, int deg, int M, const float* means, const float* campos, const float* shs,
                      uint8_t* clamped, float* out)
{
	float dir[3] = { means[3 * idx] - campos[0], means[3 * idx + 1] - campos[1], means[3 * idx + 2] - campos[2] };
	float len = sqrtf(dir[0] * dir[0] + dir[1] * dir[1] + dir[2] * dir[2]);
	dir[0] = dir[0] / len; dir[1] = dir[1] / len; dir[2] = dir[2] / len;
	const float* sh = shs + (size_t)idx * M * 3;
	float x = dir[0], y = dir[1], z = dir[2];
	for (int ch = 0; ch < 3; ch++) {
#define SHK(k) sh[3 * (k) + ch]
		float res = SH_C0 * SHK(0);
		if (deg > 0) {
			res = res - SH_C1 * y * SHK(1) + SH_C1 * z * SHK(2) - SH_C1 * x * SHK(3);
			if (deg > 1) {
				float xx = x * x, yy = y * y, zz = z * z;
				float xy = x * y, yz = y * z, xz = x * z;
				res = res +
					SH_C2[0] * xy * SHK(4) +
					SH_C2[1] * yz * SHK(5) +
					SH_C2[2] * (2.0f * zz - xx - yy) * SHK(6) +
					SH_C2[3] * xz * SHK(7) +
					SH_C2[4] * (xx - yy) * SHK(8);
				if (deg > 2) {
					res = res +
						SH_C3[0] * y * (3.0f * xx - yy) * SHK(9) +
						SH_C3[1] * xy * z * SHK(10) +
						SH_C3[2] * y * (4.0f * zz - xx - yy) * SHK(11) +
						SH_C3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy) * SHK(12) +
						SH_C3[4] * x * (4.0f * zz - xx - yy) * SHK(13) +
						SH_C3[5] * z * (xx - yy) * SHK(14) +
						SH_C3[6] * x * (xx - 3.0f * yy) * SHK(15);
				}
			}
		}
#undef SHK
		res += 0.5f;
		clamped[3 * idx + ch] = (res < 0);
		out[ch] = res < 0.0f ? 0.0f : res;
	}
}

/* ---------------------------------------------------------------- preprocess forward */
int64_t ogs_oracle_preprocess_fwd(
	int P, int D, int M,
	const float* means3D, const float* scales, float scale_modifier, const float* rotations,
	const float* opacities, const float* shs, const float* cov3D_precomp, const float* colors_precomp,
	const float* viewmatrix, const float* campos, int W, int H,
	int32_t* radii, float* means2D, float* depths, float* cov3D, float* rgb,
	float* conic_opacity, uint32_t* tiles_touched, uint32_t* point_offsets, uint8_t* clamped)
{
	const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE; /* rasterizer_impl.cu:575 */
	const float* V = viewmatrix;

#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
	for (int idx = 0; idx < P; idx++) {
		radii[idx] = 0;          /* forward.cu:626-627 */
		tiles_touched[idx] = 0;

		/* near cull: auxiliary.h:198-220 too_close */
		float t[3];
		view_point(V, means3D + 3 * idx, t);
		float rr = t[0] * t[0] + t[1] * t[1] + t[2] * t[2];
		if (rr <= 0.04f) continue;
		float r = sqrtf(rr);

		/* auxiliary.h:236-248 point3ToLonlatScreen */
		float inv_r = 1.0f / (r + EPS7);
		float lon = atan2f(t[0], t[2]);
		float lat = asinf(t[1] * inv_r);
		float sx = lon * PI_INV, sy = lat * TWO_PI_INV;

		/* forward.cu:643-652 */
		const float* c6;
		if (cov3D_precomp) c6 = cov3D_precomp + 6 * (size_t)idx;
		else {
			cov3d_from_scale_rot(scales + 3 * idx, scale_modifier, rotations + 4 * idx, cov3D + 6 * (size_t)idx);
			c6 = cov3D + 6 * (size_t)idx;
		}

		/* forward.cu:130-189 computeCov2DLonlat (recomputes t from the world mean) */
		float t2[3];
		view_point(V, means3D + 3 * idx, t2);
		lonlat_jac J = lonlat_jacobian(t2, W, H);
		m3 T, Vrk, cov;
		lonlat_T_and_cov(&J, V, c6, &T, &Vrk, &cov);
		cov.c[0][0] += 0.3f;
		cov.c[1][1] += 0.3f;
		float cx = cov.c[0][0], cy = cov.c[0][1], cz = cov.c[1][1];

		/* forward.cu:660-664 */
		float det = (cx * cz - cy * cy);
		if (det == 0.0f) continue;
		float det_inv = 1.f / det;
		float conic[3] = { cz * det_inv, -cy * det_inv, cx * det_inv };

		/* forward.cu:671-674 */
		float mid = 0.5f * (cx + cz);
		float lambda1 = mid + sqrtf(fmaxf(0.1f, mid * mid - det));
		float lambda2 = mid - sqrtf(fmaxf(0.1f, mid * mid - det));
		float my_radius = ceilf(3.f * sqrtf(fmaxf(lambda1, lambda2)));

		/* forward.cu:677-683 */
		float px = ndc_to_pix(sx, W), py = ndc_to_pix(sy, H);
		int x0, y0, x1, y1;
		tile_rect_any(px, py, (int)my_radius, gx, gy, &x0, &y0, &x1, &y1);
		if ((x1 - x0) * (y1 - y0) == 0) continue;

		/* forward.cu:688-694 */
		if (!colors_precomp)
			sh_to_rgb(idx, D, M, means3D, campos, shs, clamped, rgb + 3 * (size_t)idx);

		/* forward.cu:697-702 */
		depths[idx] = r;
		radii[idx] = (int32_t)my_radius;
		means2D[2 * idx] = px; means2D[2 * idx + 1] = py;
		conic_opacity[4 * idx + 0] = conic[0];
		conic_opacity[4 * idx + 1] = conic[1];
		conic_opacity[4 * idx + 2] = conic[2];
		conic_opacity[4 * idx + 3] = opacities[idx];
		tiles_touched[idx] = (uint32_t)((y1 - y0) * (x1 - x0));
	}

	/* rasterizer_impl.cu:622 InclusiveSum (uint32 wrap-around like the reference) */
	uint32_t acc = 0;
	for (int i = 0; i < P; i++) { acc += tiles_touched[i]; point_offsets[i] = acc; }
	return P > 0 ? (int64_t)(int32_t)point_offsets[P - 1] : 0; /* :627-628 reads it as int */
}

/* ---------------------------------------------------------------- binning */
void ogs_oracle_bin(
	int P, int W, int H,
	const float* means2D, const float* depths, const int32_t* radii, const uint32_t* point_offsets,
	int64_t R, uint64_t* keys_unsorted, uint32_t* values_unsorted,
	uint64_t* keys_sorted, uint32_t* point_list, uint32_t* ranges)
{
	const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
	const int64_t Tn = (int64_t)gx * gy;

	/* rasterizer_impl.cu:94-140 duplicateWithKeys */
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1024)
#endif
	for (int idx = 0; idx < P; idx++) {
		if (radii[idx] > 0) {
			uint32_t off = (idx == 0) ? 0 : point_offsets[idx - 1];
			int x0, y0, x1, y1;
			tile_rect_any(means2D[2 * idx], means2D[2 * idx + 1], radii[idx], gx, gy, &x0, &y0, &x1, &y1);
			uint32_t dbits;
			memcpy(&dbits, &depths[idx], 4);
			for (int y = y0; y < y1; y++)
				for (int x = x0; x < x1; x++) {
					uint64_t key = (uint64_t)(y * gx + (x >= gx ? x - gx : x));
					key <<= 32;
					key |= dbits;
					keys_unsorted[off] = key;
					values_unsorted[off] = (uint32_t)idx;
					off++;
				}
		}
	}

	/* rasterizer_impl.cu:651-661: stable LSD radix sort over bits [0, 32+bit), 8-bit digits */
	const int end_bit = 32 + (int)ogs_oracle_higher_msb((uint32_t)Tn);
	if (R > 0) {
		uint64_t* kb = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)R);
		uint32_t* vb = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)R);
		uint64_t* ka = keys_sorted; uint32_t* va = point_list;
		memcpy(ka, keys_unsorted, sizeof(uint64_t) * (size_t)R);
		memcpy(va, values_unsorted, sizeof(uint32_t) * (size_t)R);
		for (int shift = 0; shift < end_bit; shift += 8) {
			int nb = end_bit - shift < 8 ? end_bit - shift : 8;
			uint64_t mask = ((uint64_t)1 << nb) - 1;
			size_t count[257];
			memset(count, 0, sizeof(count));
			for (int64_t i = 0; i < R; i++) count[((ka[i] >> shift) & mask) + 1]++;
			for (int b = 0; b < 256; b++) count[b + 1] += count[b];
			for (int64_t i = 0; i < R; i++) {
				size_t d = (size_t)((ka[i] >> shift) & mask);
				kb[count[d]] = ka[i];
				vb[count[d]] = va[i];
				count[d]++;
			}
			uint64_t* tk = ka; ka = kb; kb = tk;
			uint32_t* tv = va; va = vb; vb = tv;
		}
		if (ka != keys_sorted) {
			memcpy(keys_sorted, ka, sizeof(uint64_t) * (size_t)R);
			memcpy(point_list, va, sizeof(uint32_t) * (size_t)R);
			free(ka); free(va);
		} else {
			free(kb); free(vb);
		}
	}

	/* rasterizer_impl.cu:664 + :145-167 */
	memset(ranges, 0, sizeof(uint32_t) * 2 * (size_t)Tn);
	for (int64_t i = 0; i < R; i++) {
		uint32_t cur = (uint32_t)(keys_sorted[i] >> 32);
		if (i == 0) ranges[2 * cur] = 0;
		else {
			uint32_t prev = (uint32_t)(keys_sorted[i - 1] >> 32);
			if (cur != prev) {
				ranges[2 * prev + 1] = (uint32_t)i;
				ranges[2 * cur] = (uint32_t)i;
			}
		}
		if (i == R - 1) ranges[2 * cur + 1] = (uint32_t)R;
	}
}

/* ---------------------------------------------------------------- render forward */
void ogs_oracle_render_fwd(
	int W, int H, const uint32_t* ranges, const uint32_t* point_list,
	const float* means2D, const float* colors, const float* conic_opacity,
	const float* background, float* final_T, uint32_t* n_contrib, float* out_color)
{
	const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
	const size_t HW = (size_t)H * W;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) collapse(2)
#endif
	for (int ty = 0; ty < gy; ty++)
		for (int tx = 0; tx < gx; tx++) {
			const uint32_t r0 = ranges[2 * (ty * gx + tx)], r1 = ranges[2 * (ty * gx + tx) + 1];
			for (int ly = 0; ly < TILE; ly++)
				for (int lx = 0; lx < TILE; lx++) {
					int pxi = tx * TILE + lx, pyi = ty * TILE + ly;
					if (!(pxi < W && pyi < H)) continue; /* forward.cu:371-373 */
					float pixf[2] = { (float)pxi, (float)pyi };
					float T = 1.0f, C[3] = { 0, 0, 0 };
					uint32_t contributor = 0, last_contributor = 0;
					/* forward.cu:417-455; the 256-entry batching does not change per-pixel results */
					for (uint32_t k = r0; k < r1; k++) {
						contributor++;
						uint32_t id = point_list[k];
						float dx = nearest_copy_x(means2D[2 * id], tx * TILE + 7.5f, (float)W) - pixf[0], dy = means2D[2 * id + 1] - pixf[1];
						const float* co = conic_opacity + 4 * (size_t)id;
						float power = -0.5f * (co[0] * dx * dx + co[2] * dy * dy) - co[1] * dx * dy;
						if (power > 0.0f) continue;
						float alpha = fminf(0.99f, co[3] * expf(power));
						if (alpha < 1.0f / 255.0f) continue;
						float test_T = T * (1 - alpha);
						if (test_T < 0.0001f) break; /* done = true */
						for (int ch = 0; ch < 3; ch++) C[ch] += colors[3 * (size_t)id + ch] * alpha * T;
						T = test_T;
						last_contributor = contributor;
					}
					size_t pix = (size_t)pyi * W + pxi; /* forward.cu:460-466 */
					final_T[pix] = T;
					n_contrib[pix] = last_contributor;
					for (int ch = 0; ch < 3; ch++) out_color[ch * HW + pix] = C[ch] + T * background[ch];
				}
		}
}

/* ---------------------------------------------------------------- render backward */
static void atomic_addf(float* p, float v)
{
#ifdef _OPENMP
#pragma omp atomic
#endif
	*p += v;
}

void ogs_oracle_render_bwd(
	int W, int H, const uint32_t* ranges, const uint32_t* point_list, const float* background,
	const float* means2D, const float* conic_opacity, const float* colors,
	const float* final_T, const uint32_t* n_contrib, const float* dL_dpixels,
	float* dL_dmean2D, float* dL_dconic, float* dL_dopacity, float* dL_dcolors)
{
	const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
	const size_t HW = (size_t)H * W;
	const float ddelx_dx = (float)(0.5 * W); /* backward.cu:739-740 */
	const float ddely_dy = (float)(0.5 * H);
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) collapse(2)
#endif
	for (int ty = 0; ty < gy; ty++)
		for (int tx = 0; tx < gx; tx++) {
			const uint32_t r0 = ranges[2 * (ty * gx + tx)], r1 = ranges[2 * (ty * gx + tx) + 1];
			for (int ly = 0; ly < TILE; ly++)
				for (int lx = 0; lx < TILE; lx++) {
					int pxi = tx * TILE + lx, pyi = ty * TILE + ly;
					if (!(pxi < W && pyi < H)) continue;
					size_t pix = (size_t)pyi * W + pxi;
					float pixf[2] = { (float)pxi, (float)pyi };
					const float T_final = final_T[pix];
					float T = T_final;
					uint32_t contributor = r1 - r0;                 /* backward.cu:723 */
					const int last_contributor = (int)n_contrib[pix];
					float accum_rec[3] = { 0, 0, 0 }, dL_dpixel[3], last_color[3] = { 0, 0, 0 };
					for (int ch = 0; ch < 3; ch++) dL_dpixel[ch] = dL_dpixels[ch * HW + pix];
					float last_alpha = 0;
					/* back to front: backward.cu:752, :762-841 */
					for (uint32_t k = r1; k-- > r0;) {
						contributor--;
						if (contributor >= (uint32_t)last_contributor) continue; /* int promoted to unsigned, :767 */
						uint32_t id = point_list[k];
						float dx = nearest_copy_x(means2D[2 * id], tx * TILE + 7.5f, (float)W) - pixf[0], dy = means2D[2 * id + 1] - pixf[1];
						const float* co = conic_opacity + 4 * (size_t)id;
						float power = -0.5f * (co[0] * dx * dx + co[2] * dy * dy) - co[1] * dx * dy;
						if (power > 0.0f) continue;
						float G = expf(power);
						float alpha = fminf(0.99f, co[3] * G);
						if (alpha < 1.0f / 255.0f) continue;

						T = T / (1.f - alpha);
						float dchannel_dcolor = alpha * T;
						float dL_dalpha = 0.0f;
						for (int ch = 0; ch < 3; ch++) {
							float c = colors[3 * (size_t)id + ch];
							accum_rec[ch] = last_alpha * last_color[ch] + (1.f - last_alpha) * accum_rec[ch];
							last_color[ch] = c;
							float dL_dchannel = dL_dpixel[ch];
							dL_dalpha += (c - accum_rec[ch]) * dL_dchannel;
							atomic_addf(&dL_dcolors[3 * (size_t)id + ch], dchannel_dcolor * dL_dchannel);
						}
						dL_dalpha *= T;
						last_alpha = alpha;

						float bg_dot_dpixel = 0;
						for (int ch = 0; ch < 3; ch++) bg_dot_dpixel += background[ch] * dL_dpixel[ch];
						dL_dalpha += (-T_final / (1.f - alpha)) * bg_dot_dpixel;

						float dL_dG = co[3] * dL_dalpha;
						float gdx = G * dx, gdy = G * dy;
						float dG_ddelx = -gdx * co[0] - gdy * co[1];
						float dG_ddely = -gdy * co[2] - gdx * co[1];

						atomic_addf(&dL_dmean2D[3 * (size_t)id + 0], dL_dG * dG_ddelx * ddelx_dx);
						atomic_addf(&dL_dmean2D[3 * (size_t)id + 1], dL_dG * dG_ddely * ddely_dy);
						atomic_addf(&dL_dconic[4 * (size_t)id + 0], -0.5f * gdx * dx * dL_dG);
						atomic_addf(&dL_dconic[4 * (size_t)id + 1], -0.5f * gdx * dy * dL_dG);
						atomic_addf(&dL_dconic[4 * (size_t)id + 3], -0.5f * gdy * dy * dL_dG);
						atomic_addf(&dL_dopacity[id], G * dL_dalpha);
					}
				}
		}
}

/* ---------------------------------------------------------------- per-Gaussian backward */

/* auxiliary.h:134-144 dnormvdv (float3) */
static void dnormvdv3(const float* v, const float* dv, float* o)
{
	float sum2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
	float invsum32 = 1.0f / sqrtf(sum2 * sum2 * sum2);
	o[0] = ((+sum2 - v[0] * v[0]) * dv[0] - v[1] * v[0] * dv[1] - v[2] * v[0] * dv[2]) * invsum32;
	o[1] = (-v[0] * v[1] * dv[0] + (sum2 - v[1] * v[1]) * dv[1] - v[2] * v[1] * dv[2]) * invsum32;
	o[2] = (-v[0] * v[2] * dv[0] - v[1] * v[2] * dv[1] + (sum2 - v[2] * v[2]) * dv[2]) * invsum32;
}

/* backward.cu:30-151 computeColorFromSH (backward) */
static void sh_bwd(int idx, int deg, int M, const float* means, const float* campos, const float* shs,
                   const uint8_t* clamped, const float* dL_dcolor, float* dL_dmeans, float* dL_dshs)
{
	float dir_orig[3] = { means[3 * idx] - campos[0], means[3 * idx + 1] - campos[1], means[3 * idx + 2] - campos[2] };
	float len = sqrtf(dir_orig[0] * dir_orig[0] + dir_orig[1] * dir_orig[1] + dir_orig[2] * dir_orig[2]);
	float x = dir_orig[0] / len, y = dir_orig[1] / len, z = dir_orig[2] / len;
	const float* sh = shs + (size_t)idx * M * 3;
	float* dsh = dL_dshs + (size_t)idx * M * 3;

	float dRGB[3];
	for (int ch = 0; ch < 3; ch++) dRGB[ch] = dL_dcolor[3 * (size_t)idx + ch] * (clamped[3 * idx + ch] ? 0.f : 1.f);

	float ddir[3] = { 0, 0, 0 }; /* dL_ddir accumulated as dot(dRGBd{x,y,z}, dL_dRGB) */
	float dRGBdx[3] = { 0, 0, 0 }, dRGBdy[3] = { 0, 0, 0 }, dRGBdz[3] = { 0, 0, 0 };

#define SHK(k) sh[3 * (k) + ch]
#define DSH(k, w) dsh[3 * (k) + ch] = (w) * dRGB[ch]
	for (int ch = 0; ch < 3; ch++) {
		DSH(0, SH_C0);
		if (deg > 0) {
			DSH(1, -SH_C1 * y);
			DSH(2, SH_C1 * z);
			DSH(3, -SH_C1 * x);
			dRGBdx[ch] = -SH_C1 * SHK(3);
			dRGBdy[ch] = -SH_C1 * SHK(1);
			dRGBdz[ch] = SH_C1 * SHK(2);
			if (deg > 1) {
				float xx = x * x, yy = y * y, zz = z * z;
				float xy = x * y, yz = y * z, xz = x * z;
				DSH(4, SH_C2[0] * xy);
				DSH(5, SH_C2[1] * yz);
				DSH(6, SH_C2[2] * (2.f * zz - xx - yy));
				DSH(7, SH_C2[3] * xz);
				DSH(8, SH_C2[4] * (xx - yy));
				dRGBdx[ch] += SH_C2[0] * y * SHK(4) + SH_C2[2] * 2.f * -x * SHK(6) + SH_C2[3] * z * SHK(7) + SH_C2[4] * 2.f * x * SHK(8);
				dRGBdy[ch] += SH_C2[0] * x * SHK(4) + SH_C2[1] * z * SHK(5) + SH_C2[2] * 2.f * -y * SHK(6) + SH_C2[4] * 2.f * -y * SHK(8);
				dRGBdz[ch] += SH_C2[1] * y * SHK(5) + SH_C2[2] * 2.f * 2.f * z * SHK(6) + SH_C2[3] * x * SHK(7);
				if (deg > 2) {
					DSH(9, SH_C3[0] * y * (3.f * xx - yy));
					DSH(10, SH_C3[1] * xy * z);
					DSH(11, SH_C3[2] * y * (4.f * zz - xx - yy));
					DSH(12, SH_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy));
					DSH(13, SH_C3[4] * x * (4.f * zz - xx - yy));
					DSH(14, SH_C3[5] * z * (xx - yy));
					DSH(15, SH_C3[6] * x * (xx - 3.f * yy));
					dRGBdx[ch] += (
						SH_C3[0] * SHK(9) * 3.f * 2.f * xy +
						SH_C3[1] * SHK(10) * yz +
						SH_C3[2] * SHK(11) * -2.f * xy +
						SH_C3[3] * SHK(12) * -3.f * 2.f * xz +
						SH_C3[4] * SHK(13) * (-3.f * xx + 4.f * zz - yy) +
						SH_C3[5] * SHK(14) * 2.f * xz +
						SH_C3[6] * SHK(15) * 3.f * (xx - yy));
					dRGBdy[ch] += (
						SH_C3[0] * SHK(9) * 3.f * (xx - yy) +
						SH_C3[1] * SHK(10) * xz +
						SH_C3[2] * SHK(11) * (-3.f * yy + 4.f * zz - xx) +
						SH_C3[3] * SHK(12) * -3.f * 2.f * yz +
						SH_C3[4] * SHK(13) * -2.f * xy +
						SH_C3[5] * SHK(14) * -2.f * yz +
						SH_C3[6] * SHK(15) * -3.f * 2.f * xy);
					dRGBdz[ch] += (
						SH_C3[1] * SHK(10) * xy +
						SH_C3[2] * SHK(11) * 4.f * 2.f * yz +
						SH_C3[3] * SHK(12) * 3.f * (2.f * zz - xx - yy) +
						SH_C3[4] * SHK(13) * 4.f * 2.f * xz +
						SH_C3[5] * SHK(14) * (xx - yy));
				}
			}
		}
	}
#undef SHK
#undef DSH
	ddir[0] = dRGBdx[0] * dRGB[0] + dRGBdx[1] * dRGB[1] + dRGBdx[2] * dRGB[2];
	ddir[1] = dRGBdy[0] * dRGB[0] + dRGBdy[1] * dRGB[1] + dRGBdy[2] * dRGB[2];
	ddir[2] = dRGBdz[0] * dRGB[0] + dRGBdz[1] * dRGB[1] + dRGBdz[2] * dRGB[2];
	float dmean[3];
	dnormvdv3(dir_orig, ddir, dmean);
	dL_dmeans[3 * (size_t)idx + 0] += dmean[0];
	dL_dmeans[3 * (size_t)idx + 1] += dmean[1];
	dL_dmeans[3 * (size_t)idx + 2] += dmean[2];
}

/* backward.cu:489-552 computeCov3D (backward) */
static void cov3d_bwd(int idx, const float* scale, float mod, const float* q, const float* dL_dcov3Ds,
                      float* dL_dscales, float* dL_drots)
{
	float r = q[0], x = q[1], y = q[2], z = q[3];
	m3 R; rot_matrix(q, &R);
	m3 S = m3_cols(1.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 1.f);
	float s[3] = { mod * scale[0], mod * scale[1], mod * scale[2] };
	S.c[0][0] = s[0]; S.c[1][1] = s[1]; S.c[2][2] = s[2];
	m3 M = m3_mul(&S, &R);
	const float* d = dL_dcov3Ds + 6 * (size_t)idx;
	m3 dSigma = m3_cols(d[0], 0.5f * d[1], 0.5f * d[2],
	                    0.5f * d[1], d[3], 0.5f * d[4],
	                    0.5f * d[2], 0.5f * d[4], d[5]);
	m3 M2; /* 2.0f * M */
	for (int c = 0; c < 3; c++) for (int rr = 0; rr < 3; rr++) M2.c[c][rr] = M.c[c][rr] * 2.0f;
	m3 dM = m3_mul(&M2, &dSigma);
	m3 Rt = m3_t(&R), dMt = m3_t(&dM);
	float* ds = dL_dscales + 3 * (size_t)idx;
	for (int k = 0; k < 3; k++)
		ds[k] = Rt.c[k][0] * dMt.c[k][0] + Rt.c[k][1] * dMt.c[k][1] + Rt.c[k][2] * dMt.c[k][2];
	for (int k = 0; k < 3; k++) { dMt.c[k][0] *= s[k]; dMt.c[k][1] *= s[k]; dMt.c[k][2] *= s[k]; }
#define D(a, b) dMt.c[a][b]
	float* dq = dL_drots + 4 * (size_t)idx; /* written un-normalised, backward.cu:551 */
	dq[0] = 2 * z * (D(0, 1) - D(1, 0)) + 2 * y * (D(2, 0) - D(0, 2)) + 2 * x * (D(1, 2) - D(2, 1));
	dq[1] = 2 * y * (D(1, 0) + D(0, 1)) + 2 * z * (D(2, 0) + D(0, 2)) + 2 * r * (D(1, 2) - D(2, 1)) - 4 * x * (D(2, 2) + D(1, 1));
	dq[2] = 2 * x * (D(1, 0) + D(0, 1)) + 2 * r * (D(2, 0) - D(0, 2)) + 2 * z * (D(1, 2) + D(2, 1)) - 4 * y * (D(2, 2) + D(0, 0));
	dq[3] = 2 * r * (D(0, 1) - D(1, 0)) + 2 * x * (D(2, 0) + D(0, 2)) + 2 * y * (D(1, 2) + D(2, 1)) - 4 * z * (D(1, 1) + D(0, 0));
#undef D
}

void ogs_oracle_preprocess_bwd(
	int P, int D, int M,
	const float* means3D, const int32_t* radii, const float* shs, const uint8_t* clamped,
	const float* scales, const float* rotations, float scale_modifier, const float* cov3D,
	const float* viewmatrix, int W, int H, const float* campos,
	const float* dL_dmean2D, const float* dL_dconic, float* dL_dcolor,
	float* dL_dmeans3D, float* dL_dcov3D, float* dL_dsh,
	float* dL_dscale, float* dL_drot, float* dpx_dt, float* dpy_dt)
{
	const float* V = viewmatrix;
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
	for (int idx = 0; idx < P; idx++) {
		if (!(radii[idx] > 0)) continue;

		/* ---- backward.cu:297-485 computeCov2DLonLatCUDA ---- */
		const float* c6 = cov3D + 6 * (size_t)idx;
		float gx_ = dL_dconic[4 * (size_t)idx], gy_ = dL_dconic[4 * (size_t)idx + 1], gz_ = dL_dconic[4 * (size_t)idx + 3];
		float t[3];
		view_point(V, means3D + 3 * idx, t);

		float txtx = t[0] * t[0], tyty = t[1] * t[1], tztz = t[2] * t[2];
		float txtytz = t[0] * t[1] * t[2];
		float a2 = txtx + tztz;                        /* trxztrxz */
		float a2_inv = 1.0f / (a2 + EPS7);
		float a4_inv = a2_inv * a2_inv;                /* trxztrxztrxztrxz_inv */
		float rho = sqrtf(a2);
		float rho_inv = 1.0f / (rho + EPS7);
		float rr = a2 + tyty;                          /* trtr */
		float rr_inv = 1.0f / (rr + EPS7);
		float rr2_inv = rr_inv * rr_inv;               /* trtrtrtr_inv */
		float rho_rr2_inv = rho_inv * rr2_inv;         /* trxz_trtrtrtr_inv */
		float rho3_rr2_inv = a2_inv * rho_rr2_inv;     /* trxztrxztrxz_trtrtrtr_inv */
		float tyty_minus_a2 = tyty - a2;

		float Wd = W * 0.5f * PI_INV;
		float Hd = H * PI_INV;
		float j00 = Wd * t[2] * a2_inv;
		float j02 = -Wd * t[0] * a2_inv;
		float j10 = -Hd * t[0] * t[1] * rho_inv * rr_inv;
		float j11 = Hd * rho * rr_inv;
		float j12 = -Hd * t[2] * t[1] * rho_inv * rr_inv;

		dpx_dt[3 * (size_t)idx + 0] = j00; dpx_dt[3 * (size_t)idx + 1] = 0.0f; dpx_dt[3 * (size_t)idx + 2] = j02;
		dpy_dt[3 * (size_t)idx + 0] = j10; dpy_dt[3 * (size_t)idx + 1] = j11;  dpy_dt[3 * (size_t)idx + 2] = j12;

		lonlat_jac Jv = { j00, j02, j10, j11, j12 };
		m3 T, Vrk, cov2D;
		lonlat_T_and_cov(&Jv, V, c6, &T, &Vrk, &cov2D);
		m3 Wm = m3_cols(V[0], V[4], V[8], V[1], V[5], V[9], V[2], V[6], V[10]);

		float a = cov2D.c[0][0] += 0.3f;
		float b = cov2D.c[0][1];
		float c = cov2D.c[1][1] += 0.3f;

		float denom = a * c - b * b;
		float dL_da = 0, dL_db = 0, dL_dc = 0;
		float denom2inv = 1.0f / ((denom * denom) + EPS7);
		float* dcov = dL_dcov3D + 6 * (size_t)idx;
#define Tm(i, j) T.c[i][j]
		if (denom2inv != 0) {
			dL_da = denom2inv * (-c * c * gx_ + 2 * b * c * gy_ + (denom - a * c) * gz_);
			dL_dc = denom2inv * (-a * a * gz_ + 2 * a * b * gy_ + (denom - a * c) * gx_);
			dL_db = denom2inv * 2 * (b * c * gx_ - (denom + 2 * b * b) * gy_ + a * b * gz_);

			dcov[0] = (Tm(0, 0) * Tm(0, 0) * dL_da + Tm(0, 0) * Tm(1, 0) * dL_db + Tm(1, 0) * Tm(1, 0) * dL_dc);
			dcov[3] = (Tm(0, 1) * Tm(0, 1) * dL_da + Tm(0, 1) * Tm(1, 1) * dL_db + Tm(1, 1) * Tm(1, 1) * dL_dc);
			dcov[5] = (Tm(0, 2) * Tm(0, 2) * dL_da + Tm(0, 2) * Tm(1, 2) * dL_db + Tm(1, 2) * Tm(1, 2) * dL_dc);
			dcov[1] = 2 * Tm(0, 0) * Tm(0, 1) * dL_da + (Tm(0, 0) * Tm(1, 1) + Tm(0, 1) * Tm(1, 0)) * dL_db + 2 * Tm(1, 0) * Tm(1, 1) * dL_dc;
			dcov[2] = 2 * Tm(0, 0) * Tm(0, 2) * dL_da + (Tm(0, 0) * Tm(1, 2) + Tm(0, 2) * Tm(1, 0)) * dL_db + 2 * Tm(1, 0) * Tm(1, 2) * dL_dc;
			dcov[4] = 2 * Tm(0, 2) * Tm(0, 1) * dL_da + (Tm(0, 1) * Tm(1, 2) + Tm(0, 2) * Tm(1, 1)) * dL_db + 2 * Tm(1, 1) * Tm(1, 2) * dL_dc;
		} else {
			for (int i = 0; i < 6; i++) dcov[i] = 0;
		}
#define Vk(i, j) Vrk.c[i][j]
		float dL_dT00 = 2 * (Tm(0, 0) * Vk(0, 0) + Tm(0, 1) * Vk(0, 1) + Tm(0, 2) * Vk(0, 2)) * dL_da +
			(Tm(1, 0) * Vk(0, 0) + Tm(1, 1) * Vk(0, 1) + Tm(1, 2) * Vk(0, 2)) * dL_db;
		float dL_dT01 = 2 * (Tm(0, 0) * Vk(1, 0) + Tm(0, 1) * Vk(1, 1) + Tm(0, 2) * Vk(1, 2)) * dL_da +
			(Tm(1, 0) * Vk(1, 0) + Tm(1, 1) * Vk(1, 1) + Tm(1, 2) * Vk(1, 2)) * dL_db;
		float dL_dT02 = 2 * (Tm(0, 0) * Vk(2, 0) + Tm(0, 1) * Vk(2, 1) + Tm(0, 2) * Vk(2, 2)) * dL_da +
			(Tm(1, 0) * Vk(2, 0) + Tm(1, 1) * Vk(2, 1) + Tm(1, 2) * Vk(2, 2)) * dL_db;
		float dL_dT10 = 2 * (Tm(1, 0) * Vk(0, 0) + Tm(1, 1) * Vk(0, 1) + Tm(1, 2) * Vk(0, 2)) * dL_dc +
			(Tm(0, 0) * Vk(0, 0) + Tm(0, 1) * Vk(0, 1) + Tm(0, 2) * Vk(0, 2)) * dL_db;
		float dL_dT11 = 2 * (Tm(1, 0) * Vk(1, 0) + Tm(1, 1) * Vk(1, 1) + Tm(1, 2) * Vk(1, 2)) * dL_dc +
			(Tm(0, 0) * Vk(1, 0) + Tm(0, 1) * Vk(1, 1) + Tm(0, 2) * Vk(1, 2)) * dL_db;
		float dL_dT12 = 2 * (Tm(1, 0) * Vk(2, 0) + Tm(1, 1) * Vk(2, 1) + Tm(1, 2) * Vk(2, 2)) * dL_dc +
			(Tm(0, 0) * Vk(2, 0) + Tm(0, 1) * Vk(2, 1) + Tm(0, 2) * Vk(2, 2)) * dL_db;
#undef Vk
#undef Tm
#define Wk(i, j) Wm.c[i][j]
		float dL_dJ00 = Wk(0, 0) * dL_dT00 + Wk(0, 1) * dL_dT01 + Wk(0, 2) * dL_dT02;
		float dL_dJ02 = Wk(2, 0) * dL_dT00 + Wk(2, 1) * dL_dT01 + Wk(2, 2) * dL_dT02;
		float dL_dJ10 = Wk(0, 0) * dL_dT10 + Wk(0, 1) * dL_dT11 + Wk(0, 2) * dL_dT12;
		float dL_dJ11 = Wk(1, 0) * dL_dT10 + Wk(1, 1) * dL_dT11 + Wk(1, 2) * dL_dT12;
		float dL_dJ12 = Wk(2, 0) * dL_dT10 + Wk(2, 1) * dL_dT11 + Wk(2, 2) * dL_dT12;
#undef Wk
		/* second derivatives of the lonlat projection, backward.cu:455-475 */
		float temp1 = Hd * tyty_minus_a2 * rho_rr2_inv;
		float temp2 = Hd * txtytz * (rr + 2.0f * a2) * rho3_rr2_inv;
		float temp3 = Wd * (txtx - tztz) * a4_inv;
		float temp4 = Wd * 2.0f * t[0] * t[2] * a4_inv;
		float temp5 = Hd * t[1] * rho3_rr2_inv;

		float dL_dtx = -dL_dJ00 * temp4
			+ dL_dJ02 * temp3
			+ dL_dJ10 * temp5 * (2.0f * txtx * a2 - tztz * rr)
			+ dL_dJ11 * t[0] * temp1
			+ dL_dJ12 * temp2;
		float dL_dty = dL_dJ10 * t[0] * temp1
			- dL_dJ11 * Hd * 2.0f * rho * t[1] * rr2_inv
			+ dL_dJ12 * t[2] * temp1;
		float dL_dtz = dL_dJ00 * temp3
			+ dL_dJ02 * temp4
			+ dL_dJ10 * temp2
			+ dL_dJ11 * t[2] * temp1
			+ dL_dJ12 * temp5 * (2.0f * tztz * a2 - txtx * rr);

		float dLdt[3] = { dL_dtx, dL_dty, dL_dtz }, dmean[3];
		view_vec_transpose(V, dLdt, dmean);
		float* dm = dL_dmeans3D + 3 * (size_t)idx;
		dm[0] = dmean[0]; dm[1] = dmean[1]; dm[2] = dmean[2]; /* assignment, backward.cu:484 */

		/* ---- backward.cu:613-669 preprocessLonLatCUDA ---- */
		float dsx_dpx = 2.0f / (float)W;
		float dsy_dpy = 2.0f / (float)H;
		float dL_dpx = dL_dmean2D[3 * (size_t)idx + 0] * dsx_dpx;
		float dL_dpy = dL_dmean2D[3 * (size_t)idx + 1] * dsy_dpy;
		float dLdt2[3] = {
			dL_dpx * j00 + dL_dpy * j10,
			dL_dpx * 0.0f + dL_dpy * j11,
			dL_dpx * j02 + dL_dpy * j12 };
		float dmean2[3];
		view_vec_transpose(V, dLdt2, dmean2);
		dm[0] += dmean2[0]; dm[1] += dmean2[1]; dm[2] += dmean2[2];

		if (shs)
			sh_bwd(idx, D, M, means3D, campos, shs, clamped, dL_dcolor, dL_dmeans3D, dL_dsh);
		if (scales)
			cov3d_bwd(idx, scales + 3 * idx, scale_modifier, rotations + 4 * idx, dL_dcov3D, dL_dscale, dL_drot);
	}
}

/* ================================================================ perspective camera (camera_type 1)
 * The binning and blend restatements above are camera independent (the reference uses the same kernels);
 * only the two per-Gaussian steps differ. */

/* auxiliary.h:95-105 transformPoint4x4 */
static void proj_point(const float* M, const float* p, float* h)
{
	h[0] = M[0] * p[0] + M[4] * p[1] + M[8] * p[2] + M[12];
	h[1] = M[1] * p[0] + M[5] * p[1] + M[9] * p[2] + M[13];
	h[2] = M[2] * p[0] + M[6] * p[1] + M[10] * p[2] + M[14];
	h[3] = M[3] * p[0] + M[7] * p[1] + M[11] * p[2] + M[15];
}

/* forward.cu:94-108 / backward.cu:179-196: frustum clamp of t.x, t.y and the perspective Jacobian */
static lonlat_jac pinhole_jacobian(float* t, float focal_x, float focal_y, float tan_fovx, float tan_fovy,
                                   float* x_grad_mul, float* y_grad_mul)
{
	const float limx = 1.3f * tan_fovx;
	const float limy = 1.3f * tan_fovy;
	const float txtz = t[0] / t[2];
	const float tytz = t[1] / t[2];
	t[0] = fminf(limx, fmaxf(-limx, txtz)) * t[2];
	t[1] = fminf(limy, fmaxf(-limy, tytz)) * t[2];
	*x_grad_mul = (txtz < -limx || txtz > limx) ? 0.f : 1.f;
	*y_grad_mul = (tytz < -limy || tytz > limy) ? 0.f : 1.f;
	lonlat_jac J;
	J.j00 = focal_x / t[2];
	J.j02 = -(focal_x * t[0]) / (t[2] * t[2]);
	J.j10 = 0.0f;
	J.j11 = focal_y / t[2];
	J.j12 = -(focal_y * t[1]) / (t[2] * t[2]);
	return J;
}

/* forward.cu:232-340 preprocessCUDA (with in_frustum, auxiliary.h:166-196, and computeCov2D, forward.cu:86-128).
 * render_depth != 0: rgb = depth in all channels, what renderDepthCUDA blends (forward.cu:566-567). */
int64_t ogs_oracle_pinhole_preprocess_fwd(
	int P, int D, int M,
	const float* means3D, const float* scales, float scale_modifier, const float* rotations,
	const float* opacities, const float* shs, const float* cov3D_precomp, const float* colors_precomp,
	const float* viewmatrix, const float* projmatrix, const float* campos, int W, int H,
	float tan_fovx, float tan_fovy, int render_depth,
	int32_t* radii, float* means2D, float* depths, float* cov3D, float* rgb,
	float* conic_opacity, uint32_t* tiles_touched, uint32_t* point_offsets, uint8_t* clamped)
{
	const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
	const float* V = viewmatrix;
	const float focal_y = H / (2.0f * tan_fovy);  /* rasterizer_impl.cu:274-276 */
	const float focal_x = W / (2.0f * tan_fovx);

#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
	for (int idx = 0; idx < P; idx++) {
		radii[idx] = 0;
		tiles_touched[idx] = 0;

		float p_view[3];
		view_point(V, means3D + 3 * idx, p_view);
		if (p_view[2] <= 0.2f) continue;   /* in_frustum */

		float p_hom[4];
		proj_point(projmatrix, means3D + 3 * idx, p_hom);
		float p_w = 1.0f / (p_hom[3] + EPS7);
		float sx = p_hom[0] * p_w, sy = p_hom[1] * p_w;

		const float* c6;
		if (cov3D_precomp) c6 = cov3D_precomp + 6 * (size_t)idx;
		else {
			cov3d_from_scale_rot(scales + 3 * idx, scale_modifier, rotations + 4 * idx, cov3D + 6 * (size_t)idx);
			c6 = cov3D + 6 * (size_t)idx;
		}

		float t[3];
		view_point(V, means3D + 3 * idx, t);
		float gmx, gmy;
		lonlat_jac J = pinhole_jacobian(t, focal_x, focal_y, tan_fovx, tan_fovy, &gmx, &gmy);
		m3 T, Vrk, cov;
		lonlat_T_and_cov(&J, V, c6, &T, &Vrk, &cov);
		cov.c[0][0] += 0.3f;
		cov.c[1][1] += 0.3f;
		float cx = cov.c[0][0], cy = cov.c[0][1], cz = cov.c[1][1];

		float det = (cx * cz - cy * cy);
		if (det == 0.0f) continue;
		float det_inv = 1.f / det;
		float conic[3] = { cz * det_inv, -cy * det_inv, cx * det_inv };

		float mid = 0.5f * (cx + cz);
		float lambda1 = mid + sqrtf(fmaxf(0.1f, mid * mid - det));
		float lambda2 = mid - sqrtf(fmaxf(0.1f, mid * mid - det));
		float my_radius = ceilf(3.f * sqrtf(fmaxf(lambda1, lambda2)));

		float px = ndc_to_pix(sx, W), py = ndc_to_pix(sy, H);
		int x0, y0, x1, y1;
		tile_rect(px, py, (int)my_radius, gx, gy, &x0, &y0, &x1, &y1);
		if ((x1 - x0) * (y1 - y0) == 0) continue;

		if (!colors_precomp)
			sh_to_rgb(idx, D, M, means3D, campos, shs, clamped, rgb + 3 * (size_t)idx);
		if (render_depth)
			rgb[3 * (size_t)idx] = rgb[3 * (size_t)idx + 1] = rgb[3 * (size_t)idx + 2] = p_view[2];

		depths[idx] = p_view[2];
		radii[idx] = (int32_t)my_radius;
		means2D[2 * idx] = px; means2D[2 * idx + 1] = py;
		conic_opacity[4 * idx + 0] = conic[0];
		conic_opacity[4 * idx + 1] = conic[1];
		conic_opacity[4 * idx + 2] = conic[2];
		conic_opacity[4 * idx + 3] = opacities[idx];
		tiles_touched[idx] = (uint32_t)((y1 - y0) * (x1 - x0));
	}
	uint32_t acc = 0;
	for (int i = 0; i < P; i++) { acc += tiles_touched[i]; point_offsets[i] = acc; }
	return P > 0 ? (int64_t)(int32_t)point_offsets[P - 1] : 0;
}

/* backward.cu:156-292 computeCov2DCUDA + :558-608 preprocessCUDA */
void ogs_oracle_pinhole_preprocess_bwd(
	int P, int D, int M,
	const float* means3D, const int32_t* radii, const float* shs, const uint8_t* clamped,
	const float* scales, const float* rotations, float scale_modifier, const float* cov3D,
	const float* viewmatrix, const float* projmatrix, int W, int H, float tan_fovx, float tan_fovy, const float* campos,
	const float* dL_dmean2D, const float* dL_dconic, float* dL_dcolor,
	float* dL_dmeans3D, float* dL_dcov3D, float* dL_dsh, float* dL_dscale, float* dL_drot)
{
	const float* V = viewmatrix;
	const float* proj = projmatrix;
	const float h_y = H / (2.0f * tan_fovy);  /* rasterizer_impl.cu:476-477 */
	const float h_x = W / (2.0f * tan_fovx);
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
	for (int idx = 0; idx < P; idx++) {
		if (!(radii[idx] > 0)) continue;
		const float* c6 = cov3D + 6 * (size_t)idx;
		float gx_ = dL_dconic[4 * (size_t)idx], gy_ = dL_dconic[4 * (size_t)idx + 1], gz_ = dL_dconic[4 * (size_t)idx + 3];
		float t[3];
		view_point(V, means3D + 3 * idx, t);
		float x_grad_mul, y_grad_mul;
		lonlat_jac Jv = pinhole_jacobian(t, h_x, h_y, tan_fovx, tan_fovy, &x_grad_mul, &y_grad_mul);
		m3 T, Vrk, cov2D;
		lonlat_T_and_cov(&Jv, V, c6, &T, &Vrk, &cov2D);
		m3 Wm = m3_cols(V[0], V[4], V[8], V[1], V[5], V[9], V[2], V[6], V[10]);

		float a = cov2D.c[0][0] += 0.3f;
		float b = cov2D.c[0][1];
		float c = cov2D.c[1][1] += 0.3f;
		float denom = a * c - b * b;
		float dL_da = 0, dL_db = 0, dL_dc = 0;
		float denom2inv = 1.0f / ((denom * denom) + EPS7);
		float* dcov = dL_dcov3D + 6 * (size_t)idx;
#define Tm(i, j) T.c[i][j]
		if (denom2inv != 0) {
			dL_da = denom2inv * (-c * c * gx_ + 2 * b * c * gy_ + (denom - a * c) * gz_);
			dL_dc = denom2inv * (-a * a * gz_ + 2 * a * b * gy_ + (denom - a * c) * gx_);
			dL_db = denom2inv * 2 * (b * c * gx_ - (denom + 2 * b * b) * gy_ + a * b * gz_);
			dcov[0] = (Tm(0, 0) * Tm(0, 0) * dL_da + Tm(0, 0) * Tm(1, 0) * dL_db + Tm(1, 0) * Tm(1, 0) * dL_dc);
			dcov[3] = (Tm(0, 1) * Tm(0, 1) * dL_da + Tm(0, 1) * Tm(1, 1) * dL_db + Tm(1, 1) * Tm(1, 1) * dL_dc);
			dcov[5] = (Tm(0, 2) * Tm(0, 2) * dL_da + Tm(0, 2) * Tm(1, 2) * dL_db + Tm(1, 2) * Tm(1, 2) * dL_dc);
			dcov[1] = 2 * Tm(0, 0) * Tm(0, 1) * dL_da + (Tm(0, 0) * Tm(1, 1) + Tm(0, 1) * Tm(1, 0)) * dL_db + 2 * Tm(1, 0) * Tm(1, 1) * dL_dc;
			dcov[2] = 2 * Tm(0, 0) * Tm(0, 2) * dL_da + (Tm(0, 0) * Tm(1, 2) + Tm(0, 2) * Tm(1, 0)) * dL_db + 2 * Tm(1, 0) * Tm(1, 2) * dL_dc;
			dcov[4] = 2 * Tm(0, 2) * Tm(0, 1) * dL_da + (Tm(0, 1) * Tm(1, 2) + Tm(0, 2) * Tm(1, 1)) * dL_db + 2 * Tm(1, 1) * Tm(1, 2) * dL_dc;
		} else {
			for (int i = 0; i < 6; i++) dcov[i] = 0;
		}
#define Vk(i, j) Vrk.c[i][j]
		float dL_dT00 = 2 * (Tm(0, 0) * Vk(0, 0) + Tm(0, 1) * Vk(0, 1) + Tm(0, 2) * Vk(0, 2)) * dL_da +
			(Tm(1, 0) * Vk(0, 0) + Tm(1, 1) * Vk(0, 1) + Tm(1, 2) * Vk(0, 2)) * dL_db;
		float dL_dT01 = 2 * (Tm(0, 0) * Vk(1, 0) + Tm(0, 1) * Vk(1, 1) + Tm(0, 2) * Vk(1, 2)) * dL_da +
			(Tm(1, 0) * Vk(1, 0) + Tm(1, 1) * Vk(1, 1) + Tm(1, 2) * Vk(1, 2)) * dL_db;
		float dL_dT02 = 2 * (Tm(0, 0) * Vk(2, 0) + Tm(0, 1) * Vk(2, 1) + Tm(0, 2) * Vk(2, 2)) * dL_da +
			(Tm(1, 0) * Vk(2, 0) + Tm(1, 1) * Vk(2, 1) + Tm(1, 2) * Vk(2, 2)) * dL_db;
		float dL_dT10 = 2 * (Tm(1, 0) * Vk(0, 0) + Tm(1, 1) * Vk(0, 1) + Tm(1, 2) * Vk(0, 2)) * dL_dc +
			(Tm(0, 0) * Vk(0, 0) + Tm(0, 1) * Vk(0, 1) + Tm(0, 2) * Vk(0, 2)) * dL_db;
		float dL_dT11 = 2 * (Tm(1, 0) * Vk(1, 0) + Tm(1, 1) * Vk(1, 1) + Tm(1, 2) * Vk(1, 2)) * dL_dc +
			(Tm(0, 0) * Vk(1, 0) + Tm(0, 1) * Vk(1, 1) + Tm(0, 2) * Vk(1, 2)) * dL_db;
		float dL_dT12 = 2 * (Tm(1, 0) * Vk(2, 0) + Tm(1, 1) * Vk(2, 1) + Tm(1, 2) * Vk(2, 2)) * dL_dc +
			(Tm(0, 0) * Vk(2, 0) + Tm(0, 1) * Vk(2, 1) + Tm(0, 2) * Vk(2, 2)) * dL_db;
#undef Vk
#undef Tm
#define Wk(i, j) Wm.c[i][j]
		float dL_dJ00 = Wk(0, 0) * dL_dT00 + Wk(0, 1) * dL_dT01 + Wk(0, 2) * dL_dT02;
		float dL_dJ02 = Wk(2, 0) * dL_dT00 + Wk(2, 1) * dL_dT01 + Wk(2, 2) * dL_dT02;
		float dL_dJ11 = Wk(1, 0) * dL_dT10 + Wk(1, 1) * dL_dT11 + Wk(1, 2) * dL_dT12;
		float dL_dJ12 = Wk(2, 0) * dL_dT10 + Wk(2, 1) * dL_dT11 + Wk(2, 2) * dL_dT12;
#undef Wk
		/* backward.cu:270-283 */
		float tz = 1.f / t[2];
		float tz2 = tz * tz;
		float tz3 = tz2 * tz;
		float dL_dtx = x_grad_mul * -h_x * tz2 * dL_dJ02;
		float dL_dty = y_grad_mul * -h_y * tz2 * dL_dJ12;
		float dL_dtz = -h_x * tz2 * dL_dJ00 - h_y * tz2 * dL_dJ11 + (2 * h_x * t[0]) * tz3 * dL_dJ02 + (2 * h_y * t[1]) * tz3 * dL_dJ12;
		float dLdt[3] = { dL_dtx, dL_dty, dL_dtz }, dmean[3];
		view_vec_transpose(V, dLdt, dmean);
		float* dm = dL_dmeans3D + 3 * (size_t)idx;
		dm[0] = dmean[0]; dm[1] = dmean[1]; dm[2] = dmean[2];   /* assignment, backward.cu:291 */

		/* backward.cu:579-597: screen position through the full projection */
		const float* m = means3D + 3 * (size_t)idx;
		float m_hom[4];
		proj_point(proj, m, m_hom);
		float m_w = 1.0f / (m_hom[3] + EPS7);
		float mul1 = (proj[0] * m[0] + proj[4] * m[1] + proj[8] * m[2] + proj[12]) * m_w * m_w;
		float mul2 = (proj[1] * m[0] + proj[5] * m[1] + proj[9] * m[2] + proj[13]) * m_w * m_w;
		float g2x = dL_dmean2D[3 * (size_t)idx + 0], g2y = dL_dmean2D[3 * (size_t)idx + 1];
		dm[0] += (proj[0] * m_w - proj[3] * mul1) * g2x + (proj[1] * m_w - proj[3] * mul2) * g2y;
		dm[1] += (proj[4] * m_w - proj[7] * mul1) * g2x + (proj[5] * m_w - proj[7] * mul2) * g2y;
		dm[2] += (proj[8] * m_w - proj[11] * mul1) * g2x + (proj[9] * m_w - proj[11] * mul2) * g2y;

		if (shs)
			sh_bwd(idx, D, M, means3D, campos, shs, clamped, dL_dcolor, dL_dmeans3D, dL_dsh);
		if (scales)
			cov3d_bwd(idx, scales + 3 * idx, scale_modifier, rotations + 4 * idx, dL_dcov3D, dL_dscale, dL_drot);
	}
}

/* rasterizer_impl.cu:64-77 checkFrustum */
void ogs_oracle_pinhole_mark_visible(int P, const float* means3D, const float* viewmatrix, uint8_t* present)
{
	for (int idx = 0; idx < P; idx++) {
		float p_view[3];
		view_point(viewmatrix, means3D + 3 * idx, p_view);
		present[idx] = (p_view[2] <= 0.2f) ? 0 : 1;
	}
}
