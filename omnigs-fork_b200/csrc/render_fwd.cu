// render_fwd.cu — 16x16-tile front-to-back alpha blending (forward).
//
// Same per-pixel arithmetic, guards and outputs as the reference's renderCUDA
// (cuda_rasterizer/forward.cu:346-467): power, alpha = min(0.99, o*exp(power)), the 1/255 and
// T < 1e-4 cut-offs, final_T, n_contrib (1-based list position of the last blended entry) and the
// planar [3,H,W] image with background.  What differs is how the work is organised:
//   * each 256-entry batch of the tile's list is gathered once per CTA, tested against the tile
//     (exact conservative ellipse/box test) and stably compacted into shared memory together with
//     its colour, list position and alpha threshold — the reference stages every entry and reads
//     colours from global memory per contributing pixel (forward.cu:448);
//   * each warp owns an 8x4 sub-tile and re-tests 32 staged entries at a time (one per lane),
//     iterating only over the ballot of entries that can reach its 32 pixels;
//   * pairs whose power is below the per-Gaussian threshold skip expf;
//   * the sub-block decisions are kept: one byte per list entry (bit w = warp w's sub-block can be reached) goes to
//     hit_bytes, so that the backward blend does not gather or test what the forward already ruled out.
// Skipped pairs are pairs the reference skips too, so results are unchanged.
#include "render_common.cuh"
#include "launchers.cuh"
#include "async_copy.cuh"
#include <cstdlib>

namespace ogs {

#ifdef OGS_TILE_TIMELINE
static __device__ unsigned long long* g_fwd_tile_clock = nullptr;
#define OGS_FWD_CLOCK g_fwd_tile_clock
#else
#define OGS_FWD_CLOCK ((unsigned long long*)nullptr)
#endif
static int tile_order_env()
{
	static const int v = [] { const char* e = getenv("OGS_TILE_ORDER"); return e ? atoi(e) : 0; }();
	return v;
}

// resident CTAs per SM: with the 52-instruction blend loop 8 (32 registers, full occupancy) measures best at C2
// (4: 0.547, 5: 0.504, 6: 0.495, 8: 0.486 ms); with the earlier 60-instruction loop it was 5 (4: 0.598, 5: 0.558,
// 6: 0.562, 8: 0.577)
#ifndef OGS_FWD_MINBLOCKS
#define OGS_FWD_MINBLOCKS 8
#endif
__global__ void __launch_bounds__(kRenderThreads, OGS_FWD_MINBLOCKS) render_fwd_kernel(
	const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list, int W, int H, int gx, int gy, int order,
	const float4* __restrict__ g0, const float4* __restrict__ g1, const float2* __restrict__ gb,
	const unsigned long long* __restrict__ scalars, const float* __restrict__ bg_color,
	float* __restrict__ final_T, uint32_t* __restrict__ n_contrib, float* __restrict__ out_color,
	uint8_t* __restrict__ hit_bytes)
{
	__shared__ StagedEntry s_e[kBatch];
	__shared__ uint32_t s_warp_cnt[kRenderThreads / 32];
	// bit w of s_hit[i]: entry i of the round can reach sub-block w (this kernel's warp w).  Flushed as one byte per list
	// entry into hit_bytes: the backward walks the same lists against the same eight sub-blocks and takes the forward's
	// decisions instead of repeating the tile-level and sub-block tests (render_bwd.cu).
	__shared__ uint32_t s_hit[kBatch];

	const int tile = tile_of_block(blockIdx.x, gx, gy, order);
	OGS_TILE_CLOCK(OGS_FWD_CLOCK, tile, 0);
	const int tile_x = tile % gx, tile_y = tile / gx;
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int sub_x0 = tile_x * kTile + (warp & 1) * kSubW;
	const int sub_y0 = tile_y * kTile + (warp >> 1) * kSubH;
	const int px = sub_x0 + (lane & (kSubW - 1));
	const int py = sub_y0 + (lane / kSubW);
	const bool inside = px < W && py < H;
	const float2 pixf = { (float)px, (float)py };

	const float tx0 = (float)(tile_x * kTile), ty0 = (float)(tile_y * kTile);
	const float tx1 = tx0 + (kTile - 1), ty1 = ty0 + (kTile - 1);
	const float sx0 = (float)sub_x0, sy0 = (float)sub_y0;
	const float sx1 = sx0 + (kSubW - 1), sy1 = sy0 + (kSubH - 1);

	const uint2 range = ranges[tile];
	const int n = (int)(range.y - range.x);
	const int rounds = (n + kBatch - 1) / kBatch;
	const float wrap_W = (n > 0 && scalars[7] != 0ull) ? (float)W : 0.f;   // > 0: seam wrap-around mode of this frame

	// a pixel's own cut-off in `power`: -inf while it blends, +inf once it is finished (forward.cu:441-445) or outside
	float lane_cut = inside ? -INFINITY : INFINITY;
	const uint32_t s_base = (uint32_t)__cvta_generic_to_shared(s_e);
	float T = 1.0f;
	uint32_t last_contributor = 0;
	float C[3] = { 0.f, 0.f, 0.f };

	int pending_round = -1;   // round whose hit bits still sit in shared memory (block-uniform)
	for (int round = 0; round < rounds; round++) {
		// all pixels of the tile saturated -> stop (forward.cu:399-401); also guards smem reuse
		const int finished = __syncthreads_count(lane_cut > 0.f);
		if (pending_round >= 0) {
			const int pi = pending_round * kBatch + threadIdx.x;
			if (pi < n) hit_bytes[range.x + pi] = (uint8_t)s_hit[threadIdx.x];
		}
		s_hit[threadIdx.x] = 0u;   // ordered against this round's atomicOr by the barrier after the compaction
		pending_round = -1;
		if (finished == kRenderThreads) break;

		// ---- gather one entry per thread, tile-level cull ----
		const int i = round * kBatch + threadIdx.x;
		bool keep = false;
		float4 a = make_float4(0, 0, 0, 0), b = a;
		float cb = 0.f, tau = 0.f;
		if (i < n) {
			const uint32_t id = point_list[range.x + i];
			a = g0[id];
			b = g1[id];
			const float2 bt = gb[id];
			cb = bt.x;
			if (wrap_W > 0.f) a.x = nearest_copy_x(a.x, tx0 + 0.5f * (kTile - 1), wrap_W);
			tau = bt.y;
			keep = gaussian_touches_box(a.x, a.y, a.z, a.w, b.x, tau, tx0, ty0, tx1, ty1);
		}
		int total;
		const int slot = block_compact_slot(keep, s_warp_cnt, total);
		if (keep) {
			s_e[slot].a = a;
			s_e[slot].b = make_float4(b.x, tau, b.y, __uint_as_float((uint32_t)(i + 1))); // 1-based list position
			s_e[slot].c = make_float4(b.z, b.w, cb, 0.f);
		}
		__syncthreads();
		pending_round = round;

		// ---- per-warp: sub-tile cull of 32 staged entries at a time, blend the survivors ----
		for (int base = 0; base < total; base += 32) {
			const int s = base + lane;
			bool hit = false;
			if (s < total) {
				const float4 ea = s_e[s].a;
				const float4 eb = s_e[s].b;
				hit = gaussian_touches_box(ea.x, ea.y, ea.z, ea.w, eb.x, eb.y, sx0, sy0, sx1, sy1);
				// list position within the round = (1-based position - 1) mod kBatch
				if (hit) atomicOr(&s_hit[(__float_as_uint(eb.w) - 1u) & (uint32_t)(kBatch - 1)], 1u << warp);
			}
			unsigned m = __ballot_sync(0xffffffffu, hit);
			if (__all_sync(0xffffffffu, lane_cut > 0.f)) break;
			while (m) {
				// one 32-bit shared address per entry (a generic pointer costs an S2R + LEA per iteration)
				const uint32_t e = s_base + (uint32_t)(base + __ffs(m) - 1) * (uint32_t)sizeof(StagedEntry);
				m &= m - 1;
				const float4 ea = lds_f4(e);
				const float4 eb = lds_f4(e + 16);
				// forward.cu:424-455, arithmetic pinned to the reference's compiled order
				float dx, dy;
				const float power = pair_power(ea.x, ea.y, ea.z, ea.w, eb.x, pixf, dx, dy);
				if (power > 0.0f) continue;
				// below the Gaussian's alpha cut-off alpha would be < 1/255 (skips expf); a finished pixel's cut-off is
				// +inf, so it drops out here without a test of its own
				if (power < fmaxf(eb.y, lane_cut)) continue;
				const float alpha = fminf(0.99f, __fmul_rn(eb.z, expf(power)));
				if (alpha < kAlphaMin) continue;
				const float test_T = __fmul_rn(T, __fsub_rn(1.0f, alpha));
				if (test_T < 0.0001f) {   // finished (forward.cu:441-445); the asm keeps this one predicated move
					asm volatile("mov.f32 %0, 0f7F800000;" : "=f"(lane_cut));
					continue;
				}
				const float4 ec = lds_f4(e + 32);
				C[0] = __fmaf_rn(T, __fmul_rn(alpha, ec.x), C[0]);
				C[1] = __fmaf_rn(T, __fmul_rn(alpha, ec.y), C[1]);
				C[2] = __fmaf_rn(T, __fmul_rn(alpha, ec.z), C[2]);
				T = test_T;
				last_contributor = __float_as_uint(eb.w);
			}
		}
	}

	if (pending_round >= 0) {   // the last round that ran: its bits are complete once every warp has left its blend loop
		__syncthreads();
		const int pi = pending_round * kBatch + threadIdx.x;
		if (pi < n) hit_bytes[range.x + pi] = (uint8_t)s_hit[threadIdx.x];
	}
	if (inside) {
		const size_t pix_id = (size_t)W * py + px;
		const size_t HW = (size_t)H * W;
		final_T[pix_id] = T;
		n_contrib[pix_id] = last_contributor;
		out_color[0 * HW + pix_id] = __fmaf_rn(bg_color[0], T, C[0]);
		out_color[1 * HW + pix_id] = __fmaf_rn(bg_color[1], T, C[1]);
		out_color[2 * HW + pix_id] = __fmaf_rn(bg_color[2], T, C[2]);
	}
	OGS_TILE_CLOCK(OGS_FWD_CLOCK, tile, 1);
}

// ------------------------------------------------------------------ experiment: asynchronous staging of the Gaussian batches
// BASELINE.json's north_star asks for "TMA/shared-memory staging of Gaussian batches" in the blend kernels.  This variant
// (OGS_FWD_STAGING=ldgsts | bulk; the product default stays the register-staged kernel above unless this one measures
// faster, see profiles/r02_optimisation_log.md) stages every round's records straight from global to shared memory with
// no register hop and one round AHEAD of the blend (double buffer):
//   kStaging 1: cp.async (LDGSTS) 16 + 16 + 8 bytes per entry, completion by cp.async.wait_group;
//   kStaging 2: cp.async.bulk (UBLKCP, the TMA unit's 1-D path) for the two 16-byte records with mbarrier complete_tx,
//               the 8-byte record by LDGSTS (bulk copies move multiples of 16 bytes).
// Records land UNcompacted (entry i of the round at slot i, its list position is implicit); the tile-level cull reads
// them from shared memory and compacts slot indices, the blend loop goes through the index.
#ifndef OGS_FWD_ASYNC_MINBLOCKS
#define OGS_FWD_ASYNC_MINBLOCKS 6
#endif
template <int kStaging>
__global__ void __launch_bounds__(kRenderThreads, OGS_FWD_ASYNC_MINBLOCKS) render_fwd_async_kernel(
	const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list, int W, int H, int gx,
	const float4* __restrict__ g0, const float4* __restrict__ g1, const float2* __restrict__ gb,
	const unsigned long long* __restrict__ scalars, const float* __restrict__ bg_color,
	float* __restrict__ final_T, uint32_t* __restrict__ n_contrib, float* __restrict__ out_color)
{
	__shared__ __align__(16) float4 s_a[2][kBatch];
	__shared__ __align__(16) float4 s_b[2][kBatch];
	__shared__ __align__(8) float2 s_c[2][kBatch];
	__shared__ uint16_t s_idx[kBatch];
	__shared__ uint32_t s_warp_cnt[kRenderThreads / 32];
	__shared__ __align__(8) uint64_t s_bar[2];

	const int tile = blockIdx.x;
	const int tile_x = tile % gx, tile_y = tile / gx;
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const int sub_x0 = tile_x * kTile + (warp & 1) * kSubW;
	const int sub_y0 = tile_y * kTile + (warp >> 1) * kSubH;
	const int px = sub_x0 + (lane & (kSubW - 1));
	const int py = sub_y0 + (lane / kSubW);
	const bool inside = px < W && py < H;
	const float2 pixf = { (float)px, (float)py };
	const float tx0 = (float)(tile_x * kTile), ty0 = (float)(tile_y * kTile);
	const float tx1 = tx0 + (kTile - 1), ty1 = ty0 + (kTile - 1);
	const float sx0 = (float)sub_x0, sy0 = (float)sub_y0;
	const float sx1 = sx0 + (kSubW - 1), sy1 = sy0 + (kSubH - 1);

	const uint2 range = ranges[tile];
	const int n = (int)(range.y - range.x);
	const int rounds = (n + kBatch - 1) / kBatch;
	const float wrap_W = (n > 0 && scalars[7] != 0ull) ? (float)W : 0.f;

	if (kStaging == 2 && tid == 0) {
		mbar_init(&s_bar[0], 1);
		mbar_init(&s_bar[1], 1);
	}
	__syncthreads();

	auto issue = [&](int round, uint32_t id_valid, uint32_t id) {
		const int buf = round & 1;
		if (kStaging == 2) {
			// one arrival with the round's byte count, then every thread's two bulk copies
			if (tid == 0) {
				const int cnt = min(kBatch, n - round * kBatch);
				mbar_arrive_expect_tx(&s_bar[buf], (uint32_t)cnt * 32u);
			}
			__syncthreads();   // the expectation is armed before any copy can complete
			if (id_valid) {
				bulk_load(&s_a[buf][tid], g0 + id, 16u, &s_bar[buf]);
				bulk_load(&s_b[buf][tid], g1 + id, 16u, &s_bar[buf]);
			}
		} else if (id_valid) {
			asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" :: "r"(smem_addr(&s_a[buf][tid])), "l"(g0 + id) : "memory");
			asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" :: "r"(smem_addr(&s_b[buf][tid])), "l"(g1 + id) : "memory");
		}
		if (id_valid)
			asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(smem_addr(&s_c[buf][tid])), "l"(gb + id) : "memory");
		asm volatile("cp.async.commit_group;" ::: "memory");
	};

	float lane_cut = inside ? -INFINITY : INFINITY;
	float T = 1.0f;
	uint32_t last_contributor = 0;
	float C[3] = { 0.f, 0.f, 0.f };

	// prologue: round 0 in flight, ids of round 1 loaded
	uint32_t id_next = 0;
	if (rounds > 0) {
		const bool v0 = tid < n;
		const uint32_t id0 = v0 ? point_list[range.x + tid] : 0u;
		issue(0, v0, id0);
		if (kBatch + tid < n) id_next = point_list[range.x + kBatch + tid];
	}
	for (int round = 0; round < rounds; round++) {
		const int buf = round & 1;
		const bool done_all = __syncthreads_count(lane_cut > 0.f) == kRenderThreads;   // also: everybody has left s_idx / buffer buf^1
		// next round's copies go out before this round is consumed
		const bool have_next = (round + 1 < rounds) && !done_all;
		if (have_next) {
			issue(round + 1, (round + 1) * kBatch + tid < n, id_next);
			if ((round + 2) * kBatch + tid < n) id_next = point_list[range.x + (round + 2) * kBatch + tid];
		}
		// wait for THIS round's records
		if (have_next) asm volatile("cp.async.wait_group 1;" ::: "memory");
		else asm volatile("cp.async.wait_group 0;" ::: "memory");
		if (kStaging == 2) mbar_wait(&s_bar[buf], (uint32_t)((round >> 1) & 1));
		__syncthreads();
		if (done_all) break;

		// ---- tile-level cull from shared memory, compact slot indices ----
		const int i = round * kBatch + tid;
		bool keep = false;
		if (i < n) {
			float4 a = s_a[buf][tid];
			const float4 b = s_b[buf][tid];
			const float2 c = s_c[buf][tid];
			if (wrap_W > 0.f) {
				a.x = nearest_copy_x(a.x, tx0 + 0.5f * (kTile - 1), wrap_W);
				s_a[buf][tid].x = a.x;
			}
			keep = gaussian_touches_box(a.x, a.y, a.z, a.w, b.x, c.y, tx0, ty0, tx1, ty1);
		}
		int total;
		const int slot = block_compact_slot(keep, s_warp_cnt, total);
		if (keep) s_idx[slot] = (uint16_t)tid;
		__syncthreads();

		const uint32_t a_base = smem_addr(&s_a[buf][0]), b_base = smem_addr(&s_b[buf][0]), c_base = smem_addr(&s_c[buf][0]);
		for (int base = 0; base < total; base += 32) {
			const int s = base + lane;
			bool hit = false;
			if (s < total) {
				const int e = s_idx[s];
				const float4 ea = s_a[buf][e];
				hit = gaussian_touches_box(ea.x, ea.y, ea.z, ea.w, s_b[buf][e].x, s_c[buf][e].y, sx0, sy0, sx1, sy1);
			}
			unsigned m = __ballot_sync(0xffffffffu, hit);
			if (__all_sync(0xffffffffu, lane_cut > 0.f)) break;
			while (m) {
				const uint32_t e = s_idx[base + __ffs(m) - 1];
				m &= m - 1;
				const float4 ea = lds_f4(a_base + e * 16u);
				const float4 eb = lds_f4(b_base + e * 16u);     // (conic.z, opacity, r, g)
				float2 ec;                                       // (b, cut-off)
				asm("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(ec.x), "=f"(ec.y) : "r"(c_base + e * 8u) : "memory");
				float dx, dy;
				const float power = pair_power(ea.x, ea.y, ea.z, ea.w, eb.x, pixf, dx, dy);
				if (power > 0.0f) continue;
				if (power < fmaxf(ec.y, lane_cut)) continue;
				const float alpha = fminf(0.99f, __fmul_rn(eb.y, expf(power)));
				if (alpha < kAlphaMin) continue;
				const float test_T = __fmul_rn(T, __fsub_rn(1.0f, alpha));
				if (test_T < 0.0001f) {
					asm volatile("mov.f32 %0, 0f7F800000;" : "=f"(lane_cut));
					continue;
				}
				C[0] = __fmaf_rn(T, __fmul_rn(alpha, eb.z), C[0]);
				C[1] = __fmaf_rn(T, __fmul_rn(alpha, eb.w), C[1]);
				C[2] = __fmaf_rn(T, __fmul_rn(alpha, ec.x), C[2]);
				T = test_T;
				last_contributor = (uint32_t)(round * kBatch) + e + 1u;   // 1-based list position
			}
		}
	}
	asm volatile("cp.async.wait_group 0;" ::: "memory");

	if (inside) {
		const size_t pix_id = (size_t)W * py + px;
		const size_t HW = (size_t)H * W;
		final_T[pix_id] = T;
		n_contrib[pix_id] = last_contributor;
		out_color[0 * HW + pix_id] = __fmaf_rn(bg_color[0], T, C[0]);
		out_color[1 * HW + pix_id] = __fmaf_rn(bg_color[1], T, C[1]);
		out_color[2 * HW + pix_id] = __fmaf_rn(bg_color[2], T, C[2]);
	}
}

// Measurement only (ogs_export_pair_counts): the amount of blending WORK in a frame, independent of how a kernel organises
// it.  One thread per pixel walks its tile's list the way the reference's renderCUDA does (forward.cu:403-455) up to the
// pixel's last contributor and counts  [0] list entries visited (= sum of n_contrib),  [1] pairs that blend (power <= 0 and
// alpha >= 1/255),  [2] list entries a tile-synchronous kernel walks (max n_contrib of the tile, per pixel),  [3] pixels.
__global__ void __launch_bounds__(kRenderThreads) pair_count_kernel(
	const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list, int W, int H, int gx,
	const float4* __restrict__ g0, const float4* __restrict__ g1, const uint32_t* __restrict__ n_contrib,
	unsigned long long* __restrict__ counts)
{
	__shared__ unsigned long long s_cnt[4];
	__shared__ int s_max;
	const int tile = blockIdx.x;
	const int px = (tile % gx) * kTile + (threadIdx.x % kTile), py = (tile / gx) * kTile + (threadIdx.x / kTile);
	const bool inside = px < W && py < H;
	if (threadIdx.x < 4) s_cnt[threadIdx.x] = 0ull;
	if (threadIdx.x == 0) s_max = 0;
	__syncthreads();
	const uint2 range = ranges[tile];
	const int last = inside ? (int)n_contrib[(size_t)W * py + px] : 0;
	atomicMax(&s_max, last);
	const float2 pixf = { (float)px, (float)py };
	unsigned long long blended = 0;
	for (int i = 0; i < last; i++) {
		const uint32_t id = point_list[range.x + i];
		const float4 a = g0[id];
		const float4 b = g1[id];
		float dx, dy;
		const float power = pair_power(a.x, a.y, a.z, a.w, b.x, pixf, dx, dy);
		if (power > 0.0f) continue;
		if (fminf(0.99f, __fmul_rn(b.y, expf(power))) < kAlphaMin) continue;
		blended++;
	}
	atomicAdd(&s_cnt[0], (unsigned long long)last);
	atomicAdd(&s_cnt[1], blended);
	if (inside) atomicAdd(&s_cnt[3], 1ull);
	__syncthreads();
	if (threadIdx.x == 0) {
		atomicAdd(&counts[0], s_cnt[0]);
		atomicAdd(&counts[1], s_cnt[1]);
		atomicAdd(&counts[2], (unsigned long long)s_max * s_cnt[3]);
		atomicAdd(&counts[3], s_cnt[3]);
	}
}

int launch_pair_count(const uint2* ranges, const uint32_t* point_list, int W, int H, const float4* g0, const float4* g1,
                      const uint32_t* n_contrib, unsigned long long* counts, cudaStream_t st)
{
	const int gx = ceil_div(W, kTile), gy = ceil_div(H, kTile);
	OGS_CUDA_TRY(cudaMemsetAsync(counts, 0, 4 * sizeof(unsigned long long), st));
	pair_count_kernel<<<gx * gy, kRenderThreads, 0, st>>>(ranges, point_list, W, H, gx, g0, g1, n_contrib, counts);
	OGS_CUDA_TRY(cudaGetLastError());
	return OGS_OK;
}

#ifdef OGS_TILE_TIMELINE
extern "C" __attribute__((visibility("default"))) int ogs_debug_set_fwd_tile_clock(unsigned long long* p)
{
	return cudaMemcpyToSymbol(g_fwd_tile_clock, &p, sizeof(p)) == cudaSuccess ? 0 : -2;
}
#endif

int launch_render_fwd(const uint2* ranges, const uint32_t* point_list, int W, int H,
                      const float4* g0, const float4* g1, const float2* gb, const unsigned long long* scalars, const float* bg,
                      float* final_T, uint32_t* n_contrib, float* out_color, uint8_t* hit_bytes, cudaStream_t st)
{
	const int gx = ceil_div(W, kTile), gy = ceil_div(H, kTile);
	// OGS_FWD_STAGING=ldgsts|bulk selects the asynchronous-staging experiment (A/B measurements only; it leaves no hit
	// bytes, so the backward then runs its own tests: render_hit_bytes_enabled())
	const int staging = fwd_staging_mode();
	if (staging == 1)
		render_fwd_async_kernel<1><<<gx * gy, kRenderThreads, 0, st>>>(ranges, point_list, W, H, gx, g0, g1, gb, scalars, bg,
		                                                              final_T, n_contrib, out_color);
	else if (staging == 2)
		render_fwd_async_kernel<2><<<gx * gy, kRenderThreads, 0, st>>>(ranges, point_list, W, H, gx, g0, g1, gb, scalars, bg,
		                                                              final_T, n_contrib, out_color);
	else
		render_fwd_kernel<<<gx * gy, kRenderThreads, 0, st>>>(ranges, point_list, W, H, gx, gy, tile_order_env(), g0, g1, gb,
		                                                     scalars, bg, final_T, n_contrib, out_color, hit_bytes);
	OGS_CUDA_TRY(cudaGetLastError());
	return OGS_OK;
}

} // namespace ogs
