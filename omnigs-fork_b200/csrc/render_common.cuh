// render_common.cuh — pieces shared by the forward and backward tile kernels.
#pragma once
#include "ogs_common.cuh"

namespace ogs {

constexpr int kRenderThreads = 256;   // one 16x16 tile per CTA, one pixel per thread
constexpr int kBatch = 256;           // list entries staged per round (reference BLOCK_SIZE)
// Each warp owns an 8x4 pixel sub-tile (instead of the reference's 16x2 strip): a compact
// footprint makes the per-warp sub-tile cull below reject far more (warp, Gaussian) pairs.
constexpr int kSubW = 8, kSubH = 4;

// Safety margins for skipping work that provably cannot change the result.
// A pair (pixel, Gaussian) contributes only if alpha = min(0.99, o*exp(power)) >= 1/255, i.e.
// power >= tau := -ln(255*o).  `power` is a float32 quadratic form; its evaluation error is
// bounded by a few ulps of its largest term, so conservative tests subtract
//   kCullAbs + kCullRel * (|A|dx^2 + |C|dy^2 + 2|B||dx||dy|)   (evaluated at the farthest corner)
// before declaring "cannot contribute".  The skipped pairs are exactly pairs the reference also
// skips (forward.cu:428-438), so images, n_contrib and gradients are unchanged.
constexpr float kCullAbs = 0.01f;
constexpr float kCullRel = 8e-6f;

// The blend decision in `power` space.  A pair contributes iff alpha = min(0.99, o * expf(power)) >= 1/255
// (forward.cu:434-438).  o * expf(p) is non-decreasing in p, so per Gaussian there is a smallest float p_cut with
// o * expf(p_cut) >= 1/255 and the decision is exactly `power >= p_cut` — evaluated with the SAME expf and the same
// multiply the blend uses.  It is found once per Gaussian (preprocess) by walking a few ulps around -ln(255 o);
// the blend kernels then skip pairs with one compare, and the backward (which must repeat the forward's decisions:
// a pixel that un-multiplies T by a Gaussian the forward skipped is off by 0.4 % for the rest of its list) needs
// no accurate expf at all.  The forward still applies the reference's own alpha test after expf.
OGS_D float alpha_cutoff_power(float opacity)
{
	if (!(opacity > 0.f)) return (opacity <= 0.f) ? INFINITY : -INFINITY;   // o <= 0 never contributes; NaN: no skip
	float p = -logf(255.0f * opacity);
	int it = 0;
	for (; it < 16 && !(__fmul_rn(opacity, expf(p)) < kAlphaMin); it++) p = nextafterf(p, -INFINITY);
	for (; it < 48 && (__fmul_rn(opacity, expf(p)) < kAlphaMin); it++) p = nextafterf(p, INFINITY);
	// not converged (cannot happen for finite o; guards against a non-monotone expf): fall back to a safe bound
	return (it < 48) ? p : (-logf(255.0f * opacity) - 1e-3f);
}

// Can the Gaussian (mean m, conic A,B,C, threshold tau) reach alpha >= 1/255 anywhere on the
// pixel-centre box [x0,x1]x[y0,y1]?  Conservative: returns true when unsure.
// The maximum of power(d) = -0.5*(A dx^2 + 2B dx dy + C dy^2) over the box of offsets
// d = m - pix is 0 when the mean lies inside; otherwise (convex form, minimiser outside the box)
// it is attained on a face that faces the mean, where the 1-D minimiser has a closed form.
// The minimiser's division is rcp.approx (1 ulp) instead of an IEEE division (8 instructions and a slow path): the form's
// derivative along the face vanishes at the minimiser, so an ulp of error there moves qmin by far less than the margin.
OGS_D bool gaussian_touches_box(float mx, float my, float A, float B, float C, float tau,
                                float x0, float y0, float x1, float y1)
{
	const float dx_lo = mx - x1, dx_hi = mx - x0, dy_lo = my - y1, dy_hi = my - y0;
	const bool in_x = (dx_lo <= 0.f) && (dx_hi >= 0.f);
	const bool in_y = (dy_lo <= 0.f) && (dy_hi >= 0.f);
	if (in_x && in_y) return true;
	// not provably positive definite -> keep.  The lower bound also keeps rcp.approx.ftz below in its normal range (a
	// denormal conic entry would flush to 0 and the clamped point would no longer be the face minimiser); conic entries
	// never exceed 1 / 0.3 (the 0.3 px blur bounds the eigenvalues of cov2D from below), so no upper guard is needed.
	if (!(A > 1e-30f && C > 1e-30f && A * C - B * B > 0.f)) return true;
	float qmin = INFINITY;
	if (!in_x) {
		const float dx = (dx_lo > 0.f) ? dx_lo : dx_hi;
		const float dy = fminf(fmaxf(-B * dx * rcp_approx(C), dy_lo), dy_hi);   // C > 0; see the note on the divisions above
		qmin = fminf(qmin, A * dx * dx + 2.f * B * dx * dy + C * dy * dy);
	}
	if (!in_y) {
		const float dy = (dy_lo > 0.f) ? dy_lo : dy_hi;
		const float dx = fminf(fmaxf(-B * dy * rcp_approx(A), dx_lo), dx_hi);
		qmin = fminf(qmin, A * dx * dx + 2.f * B * dx * dy + C * dy * dy);
	}
	const float mdx = fmaxf(fabsf(dx_lo), fabsf(dx_hi)), mdy = fmaxf(fabsf(dy_lo), fabsf(dy_hi));
	const float mag = A * mdx * mdx + C * mdy * mdy + 2.f * fabsf(B) * mdx * mdy;
	return -0.5f * qmin >= tau - kCullAbs - kCullRel * mag;
}

// power(d) of a (pixel, Gaussian) pair in the reference's compiled operation order
// (forward.cu:424-427 / backward.cu:772-775 as scheduled in its sm_100 SASS):
//   q = fma(A*dx, dx, (C*dy)*dy);  power = fma(q, -0.5, -((B*dx)*dy))
// Pinned with .rn intrinsics so the skip decisions (power > 0, alpha < 1/255, T < 1e-4) are
// bit-identical to the reference's.
OGS_D float pair_power(float mx, float my, float A, float B, float C, float2 pixf, float& dx, float& dy)
{
	dx = __fsub_rn(mx, pixf.x);
	dy = __fsub_rn(my, pixf.y);
	const float q = __fmaf_rn(__fmul_rn(A, dx), dx, __fmul_rn(__fmul_rn(C, dy), dy));
	return __fmaf_rn(q, -0.5f, -__fmul_rn(__fmul_rn(B, dx), dy));
}

// Seam wrap-around (opt-in): a tile uses the copy of the Gaussian (x, x - W, x + W) nearest to its centre.
OGS_D float nearest_copy_x(float mx, float tile_cx, float Wf)
{
	const float d = mx - tile_cx;
	return d > 0.5f * Wf ? mx - Wf : (d < -0.5f * Wf ? mx + Wf : mx);
}

// Which tile a CTA works on (OGS_TILE_ORDER, default 0).  order 0: row-major (blockIdx.x).  order 1: tile rows alternately
// from the top and the bottom of the frame, equator last.  order 2: from the equator outwards, poles last.  Row-major inside
// a row in all of them, so horizontally neighbouring tiles (which share most of their Gaussians) still run together.
// Measured per-tile blend times (profiles/r02_tile_timeline.json): on uniformly filled scenes the EQUATOR tiles are the
// slow ones (C2: 71 us against 33 us at the poles in the forward), so order 2 is the longest-first order that keeps locality.
OGS_D int tile_of_block(int block, int gx, int gy, int order)
{
	if (order == 0) return block;
	const int r = block / gx, x = block - r * gx;
	int row;
	if (order == 1) {
		row = (r & 1) ? (gy - 1 - (r >> 1)) : (r >> 1);
	} else {
		// mid, mid-1, mid+1, mid-2, ... and, once the shorter side has run out, the rest of the longer side
		const int mid = gy >> 1, k = (r + 1) >> 1;
		const int below = mid, above = gy - 1 - mid, paired = min(below, above);
		if (r == 0) row = mid;
		else if (k <= paired) row = (r & 1) ? mid - k : mid + k;
		else row = (below > above) ? mid - (paired + (r - 2 * paired)) : mid + (paired + (r - 2 * paired));
	}
	return row * gx + x;
}

// Per-tile start / end timestamps (OGS_TILE_TIMELINE builds only; tools/tile_timeline.py): what the tail of a blend
// kernel costs is measured, not guessed.
#ifdef OGS_TILE_TIMELINE
OGS_D unsigned long long global_ns()
{
	unsigned long long t;
	asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
	return t;
}
#define OGS_TILE_CLOCK(ptr, tile, which)                                                   \
	do {                                                                                   \
		if (threadIdx.x == 0 && (ptr) != nullptr) (ptr)[2 * (size_t)(tile) + (which)] = global_ns(); \
	} while (0)
#else
#define OGS_TILE_CLOCK(ptr, tile, which) do { } while (0)
#endif

// One staged list entry in shared memory (48 bytes, a single base address per inner-loop iteration):
//   a = (mean.x, mean.y, conic.x, conic.y)
//   b = (conic.z, tau_safe, opacity, list position as bits)
//   c = (colour.r, colour.g, colour.b, Gaussian id as bits)
struct __align__(16) StagedEntry {
	float4 a, b, c;
};

OGS_D float4 lds_f4(uint32_t shared_addr)
{
	float4 r;
	asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(shared_addr) : "memory");
	return r;
}

// Stable block-wide compaction slot for `keep` flags (list order must be preserved: blending is
// order dependent).  Returns this thread's slot (valid when keep) and the block total.
// Uses one __syncthreads; s_warp_cnt must hold kRenderThreads/32 words.
OGS_D int block_compact_slot(bool keep, uint32_t* s_warp_cnt, int& total)
{
	const unsigned ballot = __ballot_sync(0xffffffffu, keep);
	const int warp = threadIdx.x >> 5;
	if ((threadIdx.x & 31) == 0) s_warp_cnt[warp] = __popc(ballot);
	__syncthreads();
	int off = 0, tot = 0;
#pragma unroll
	for (int w = 0; w < kRenderThreads / 32; w++) {
		int c = (int)s_warp_cnt[w];
		if (w < warp) off += c;
		tot += c;
	}
	total = tot;
	return off + __popc(ballot & lanemask_lt());
}

} // namespace ogs
