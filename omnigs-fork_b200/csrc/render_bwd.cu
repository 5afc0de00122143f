// render_bwd.cu — 16x16-tile back-to-front replay of the blend (backward).
//
// Per-pixel arithmetic follows the reference's backward renderCUDA
// (cuda_rasterizer/backward.cu:672-843): T is un-multiplied step by step, accum_rec / last_alpha /
// last_color carry the colour behind the current Gaussian, and the derivatives of the pixel colour
// w.r.t. the Gaussian's 2-D mean, conic, opacity and colour are the same expressions.  The reference
// then issues 9 global float atomics per (pixel, Gaussian) pair (backward.cu:805,829-840).  Here:
//   * the tile list is cut at the tile's largest n_contrib (nothing behind it was blended);
//   * the forward left one byte per list entry saying which of the tile's eight 8x4 sub-blocks the entry can reach
//     (render_fwd.cu, hit_bytes): entries with a zero byte are neither gathered nor tested, the others are stably
//     compacted into shared memory with their byte (without the bytes — OGS_BWD_HITS=0, or after the experiment forward
//     kernels — the kernel runs the tile-level and sub-block tests itself: kHits = false);
//   * a CTA is two warps; each warp owns a 16x8 half-tile and each LANE owns four pixels, one in each
//     8x4 sub-block of that half.  The warp takes 32 staged entries at a time (one per lane, four ballots: a bit of
//     the entry's byte per sub-block) and visits only entries that can reach it;
//   * per visited entry every lane adds up, over its (up to four) contributing pixels, nine raw sums:
//       u*dx, u*dy, u*dx^2, u*dx*dy, u*dy^2 (u = dL/dG * G), G*dL/dalpha and the three colour terms.
//     They are reduced over the warp with ONE transposing butterfly (14 shuffles for 9 values) —
//     the fixed cost per entry is paid once per 128 pixels, not once per 32 — and go to L2 as they are:
//     one red.global.add.f32 instruction per (half-tile, Gaussian) whose nine active lanes cover 36
//     contiguous bytes of the Gaussian's packed 12-float accumulator row.  The raw sums are linear in the
//     pixels, so the conic factors of the reference's quantities
//       dL/dmean2D.x = -(W/2) (A Su_dx + B Su_dy),  dL/dconic.x = -1/2 Su_dx2, ...
//     are applied once per Gaussian by the per-Gaussian backward (preprocess_bwd.cu), not here.
// Summation order differs from the reference's (unordered) atomics; parity is within the stated tolerance.
#include "render_common.cuh"
#include "launchers.cuh"
#include <cstdlib>

namespace ogs {

#ifdef OGS_TILE_TIMELINE
static __device__ unsigned long long* g_bwd_tile_clock = nullptr;
#define OGS_BWD_CLOCK g_bwd_tile_clock
#else
#define OGS_BWD_CLOCK ((unsigned long long*)nullptr)
#endif

constexpr int kBwdThreads = 64;                 // two warps per tile
constexpr int kBwdSlots = 4;                    // pixels per lane (one per 8x4 sub-block)
#ifndef OGS_BWD_BATCH
#define OGS_BWD_BATCH 128
#endif
constexpr int kBwdBatch = OGS_BWD_BATCH;        // list entries staged per round
constexpr int kBwdPerThread = kBwdBatch / kBwdThreads;
#ifndef OGS_BWD_CHUNK
#define OGS_BWD_CHUNK 256
#endif
constexpr int kBwdChunk = OGS_BWD_CHUNK;        // list positions whose hit bytes are scanned together (kHits)

// Sum v[0..7] and v8 over the 32 lanes.  On return lane L holds in `z` the total of value (L>>2)
// (replicated over the 4 lanes of a quad) and every lane holds the total of v8 in `z8`.
OGS_D void warp_transpose_reduce9(const float (&v)[8], float v8, int lane, float& z, float& z8)
{
	const unsigned full = 0xffffffffu;
	const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
	float w[4], u[2];
#pragma unroll
	for (int i = 0; i < 4; i++) {
		const float send = b4 ? v[i] : v[i + 4];
		const float keep = b4 ? v[i + 4] : v[i];
		w[i] = keep + __shfl_xor_sync(full, send, 16);
	}
#pragma unroll
	for (int i = 0; i < 2; i++) {
		const float send = b3 ? w[i] : w[i + 2];
		const float keep = b3 ? w[i + 2] : w[i];
		u[i] = keep + __shfl_xor_sync(full, send, 8);
	}
	{
		const float send = b2 ? u[0] : u[1];
		const float keep = b2 ? u[1] : u[0];
		z = keep + __shfl_xor_sync(full, send, 4);
	}
	z += __shfl_xor_sync(full, z, 2);
	z += __shfl_xor_sync(full, z, 1);
	z8 = v8;
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) z8 += __shfl_xor_sync(full, z8, o);
}

OGS_D float ex2_approx(float x)
{
	float r;
	asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
	return r;
}

// kHits: the forward blend left one byte per list entry whose bit w says "can reach sub-block w of the tile" (the same
// eight 8x4 sub-blocks: this kernel's warp W, slot s is the forward's warp 4W + s).  The tile-level and sub-block tests
// below are then not repeated: entries with a zero byte are not even gathered, the others bring their ballots with them.
// Bits are only ever missing for entries a sub-block's pixels had all finished before (no pixel blended them).
template <int kMinBlocks, bool kHits>
__global__ void __launch_bounds__(kBwdThreads, kMinBlocks) render_bwd_kernel(
	const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list, int W, int H, int gx, int gy, int order,
	const float* __restrict__ bg_color,
	const float4* __restrict__ g0, const float4* __restrict__ g1, const float2* __restrict__ gb,
	const unsigned long long* __restrict__ scalars,
	const float* __restrict__ final_Ts, const uint32_t* __restrict__ n_contrib,
	const float* __restrict__ dL_dpixels, float* __restrict__ grad_acc /*[P,12]*/, const uint8_t* __restrict__ hit_bytes)
{
	__shared__ StagedEntry s_e[kBwdBatch];
	__shared__ uint8_t s_mask[kHits ? kBwdBatch : 4];
	__shared__ int s_pos[kHits ? kBwdChunk + kBwdBatch : 1];        // reachable list positions waiting to be staged, descending
	__shared__ uint8_t s_posb[kHits ? kBwdChunk + kBwdBatch : 4];   // and their hit bytes
	// per (slot, thread): dL/dpixel (r, g, b) and -T_final * (bg . dL/dpixel) — read once per contributing pair with one
	// conflict-free 128-bit load instead of living in 16 registers
	__shared__ float4 s_pix[kBwdSlots][kBwdThreads];
	__shared__ int s_warp_cnt[kBwdThreads / 32];
	__shared__ int s_max_contrib;

	const int tile = tile_of_block(blockIdx.x, gx, gy, order);
	OGS_TILE_CLOCK(OGS_BWD_CLOCK, tile, 0);
	const int tile_x = tile % gx, tile_y = tile / gx;
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const size_t HW = (size_t)H * W;

	const float tx0 = (float)(tile_x * kTile), ty0 = (float)(tile_y * kTile);
	const float tx1 = tx0 + (kTile - 1), ty1 = ty0 + (kTile - 1);
	// this warp's half-tile starts at row 8*warp; slot s is the 8x4 sub-block (s & 1, s >> 1) of it
	const int half_x0 = tile_x * kTile, half_y0 = tile_y * kTile + 8 * warp;

	const uint2 range = ranges[tile];
	// this lane's element of a Gaussian's accumulator row (see the reduction at the end of a visit); opaque so that the
	// pointer stays in registers instead of being rebuilt from %tid for every visit
	const bool ninth = (lane == 1), red_lane = (lane & 3) == 0 || ninth;
	float* lane_acc = grad_acc + (ninth ? 8 : (lane >> 2));
	asm volatile("" : "+l"(lane_acc));
	int lane_bits = lane;   // same for the butterfly's lane-bit selects
	asm volatile("" : "+r"(lane_bits));
	uint32_t pix_addr = (uint32_t)__cvta_generic_to_shared(&s_pix[0][tid]);
	asm volatile("" : "+r"(pix_addr));   // opaque: one register for the whole kernel instead of S2R + LEA per use

	// per-pixel state (backward.cu:717-740), one set per slot
	// accum_rec holds the colour accumulated behind the NEXT Gaussian to be visited: the reference's
	// update accum_rec = last_alpha*last_color + (1-last_alpha)*accum_rec (backward.cu:797) is applied
	// right after a Gaussian is processed instead of right before the next one (same expression, same
	// values, no last_alpha / last_color registers).
	float T[kBwdSlots];
	float accum_rec[kBwdSlots][3];
	int last_contributor[kBwdSlots];
	const float px0f = (float)(half_x0 + (lane & (kSubW - 1))), py0f = (float)(half_y0 + (lane / kSubW));
	// the second column / row of sub-blocks (exact integers); opaque so that they stay in two registers instead of being
	// re-added for every slot of every visit
	float px1f = px0f + (float)kSubW, py1f = py0f + (float)kSubH;
	asm volatile("" : "+f"(px1f), "+f"(py1f));
	const float bg[3] = { bg_color[0], bg_color[1], bg_color[2] };
	int my_max = 0;
#pragma unroll
	for (int s = 0; s < kBwdSlots; s++) {
		const int px = half_x0 + kSubW * (s & 1) + (lane & (kSubW - 1));
		const int py = half_y0 + kSubH * (s >> 1) + (lane / kSubW);
		const bool inside = px < W && py < H;
		const size_t pix_id = (size_t)W * py + px;
		const float T_final = inside ? final_Ts[pix_id] : 0.f;
		T[s] = T_final;
		last_contributor[s] = inside ? (int)n_contrib[pix_id] : 0;   // 0 => no list position passes the test
		my_max = max(my_max, last_contributor[s]);
		float bg_dot = 0.f, dLp[3];
#pragma unroll
		for (int ch = 0; ch < 3; ch++) {
			accum_rec[s][ch] = 0.f;
			dLp[ch] = inside ? dL_dpixels[ch * HW + pix_id] : 0.f;
			bg_dot += bg[ch] * dLp[ch];
		}
		// (-T_final / (1 - alpha)) * bg_dot = .w * 1/(1 - alpha)
		s_pix[s][tid] = make_float4(dLp[0], dLp[1], dLp[2], -T_final * bg_dot);
	}

	// entries at list positions >= max(n_contrib) of the tile were blended by no pixel
	if (tid == 0) s_max_contrib = 0;
	__syncthreads();
	if (my_max > 0) atomicMax(&s_max_contrib, my_max);
	__syncthreads();
	const int n = min((int)(range.y - range.x), s_max_contrib);
	const int rounds = (n + kBwdBatch - 1) / kBwdBatch;
	const float wrap_W = (n > 0 && scalars[7] != 0ull) ? (float)W : 0.f;   // > 0: seam wrap-around mode of this frame

	// ---- per-warp replay over the staged entries that can reach this half-tile (entries [0, total) of s_e) ----
	auto replay = [&](const int total) {
		for (int base = 0; base < total; base += 32) {
			const int e_idx = base + lane;
			unsigned m[kBwdSlots];
			if constexpr (kHits) {
				const uint32_t mb = (e_idx < total) ? ((uint32_t)s_mask[e_idx] >> (4 * warp)) : 0u;
#pragma unroll
				for (int s = 0; s < kBwdSlots; s++) m[s] = __ballot_sync(0xffffffffu, (mb >> s) & 1u);
			} else {
				bool hit[kBwdSlots] = { false, false, false, false };
				if (e_idx < total) {
					const float4 ea = s_e[e_idx].a;
					const float4 eb = s_e[e_idx].b;
#pragma unroll
					for (int s = 0; s < kBwdSlots; s++) {
						const float bx0 = (float)(half_x0 + kSubW * (s & 1)), by0 = (float)(half_y0 + kSubH * (s >> 1));
						hit[s] = gaussian_touches_box(ea.x, ea.y, ea.z, ea.w, eb.x, eb.y, bx0, by0, bx0 + (kSubW - 1), by0 + (kSubH - 1));
					}
				}
#pragma unroll
				for (int s = 0; s < kBwdSlots; s++) m[s] = __ballot_sync(0xffffffffu, hit[s]);
			}
			unsigned m_any = m[0] | m[1] | m[2] | m[3];
			while (m_any) {
				const int bit = __ffs(m_any) - 1;
				m_any &= m_any - 1;
				const int j = base + bit;
				const StagedEntry* e = &s_e[j];
				const float4 ea = e->a;
				const float4 eb = e->b;
				const float4 ec = e->c;
				const int list_pos = __float_as_int(eb.w);

				float r[8] = { 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f };
				float r8 = 0.f;
				int any_valid = 0;
#pragma unroll
				for (int s = 0; s < kBwdSlots; s++) {
					if (!((m[s] >> bit) & 1u)) continue;   // warp-uniform
					// backward.cu:766-782: same guards (and the same pinned arithmetic) as the forward
					float dx, dy;
					const float2 pixf = { (s & 1) ? px1f : px0f, (s >> 1) ? py1f : py0f };
					const float power = pair_power(ea.x, ea.y, ea.z, ea.w, eb.x, pixf, dx, dy);
					bool valid = (list_pos < last_contributor[s]) && !(power > 0.0f) && !(power < eb.y);
					float G = 0.f, alpha = 0.f;
					if (valid) {
						// power >= the Gaussian's cut-off (eb.y) IS the forward's alpha >= 1/255 decision, so the gradient
						// arithmetic itself is free to use ex2.approx (2 instructions instead of expf's 9): gradients are
						// tolerance-bound, not bit-compared
	#ifdef OGS_BWD_EXACT_MATH   // A/B build (tools/grad_noise.py): the reference's own expf / IEEE division
						G = expf(power);
	#else
						G = ex2_approx(power * 1.4426950408889634f);   // ex2.approx.ftz: no denormal rescaling (power >= cut-off > -6)
	#endif
						alpha = fminf(0.99f, __fmul_rn(eb.z, G));
					}
					if (valid) {
						any_valid = 1;
						// backward.cu:784-840; one reciprocal (rcp.approx, 1 ulp; 1 - alpha >= 0.01) replaces the two
						// divisions by (1 - alpha)
						const float4 pix = lds_f4(pix_addr + (uint32_t)(s * kBwdThreads * sizeof(float4)));
						const float dL_dpixel[3] = { pix.x, pix.y, pix.z };
	#ifdef OGS_BWD_EXACT_MATH
						const float inv = __fdiv_rn(1.f, 1.f - alpha);
	#else
						const float inv = rcp_approx(1.f - alpha);
	#endif
						T[s] = T[s] * inv;
						const float dchannel_dcolor = alpha * T[s];
						float dL_dalpha = 0.0f;
						const float col[3] = { ec.x, ec.y, ec.z };
#pragma unroll
						for (int ch = 0; ch < 3; ch++) {
							// accum_rec' = alpha c + (1 - alpha) accum_rec (backward.cu:797) = accum_rec + alpha (c - accum_rec):
							// the difference is needed anyway, so the update is one FMA
							const float d = col[ch] - accum_rec[s][ch];
							dL_dalpha = fmaf(d, dL_dpixel[ch], dL_dalpha);
							accum_rec[s][ch] = fmaf(alpha, d, accum_rec[s][ch]);
						}
						r[6] += dchannel_dcolor * dL_dpixel[0];
						r[7] += dchannel_dcolor * dL_dpixel[1];
						r8 += dchannel_dcolor * dL_dpixel[2];
						dL_dalpha *= T[s];
						dL_dalpha += pix.w * inv;

						const float u = (eb.z * dL_dalpha) * G;   // dL/dG * G
						const float udx = u * dx, udy = u * dy;
						r[0] += udx;
						r[1] += udy;
						r[2] += udx * dx;
						r[3] += udx * dy;
						r[4] += udy * dy;
						r[5] += G * dL_dalpha;
					}
				}
				if (!__any_sync(0xffffffffu, any_valid != 0)) continue;
				float z, z8;
				warp_transpose_reduce9(r, r8, lane_bits, z, z8);
				// lanes 0,4,..,28 hold sums 0..7, lane 1 takes the ninth: one 9-lane L2 reduction
				if (red_lane)
					red_add(lane_acc + (size_t)__float_as_uint(ec.w) * 12, (lane_bits == 1) ? z8 : z);
			}
		}
	};
	if constexpr (kHits) {
		// The forward's hit bytes say which list positions matter.  A chunk of kBwdChunk positions is scanned first (bytes
		// only) and its reachable positions are appended, in descending list order, to s_pos; whenever kBwdBatch of them
		// wait there (or the list is exhausted) a batch is gathered and replayed — the gather latency and the two barriers
		// of a batch are paid per kBwdBatch STAGED entries, where the loop below (own tests) pays them per kBwdBatch list
		// positions, a quarter of which survive at C2.
		const int chunks = (n + kBwdChunk - 1) / kBwdChunk;
		// have: reachable positions waiting in s_pos[off, off + have); all three are block-uniform
		int have = 0, off = 0, next_chunk = 0;
		for (;;) {
			if (have < kBwdBatch && next_chunk < chunks) {
				// ---- scan the next chunk's hit bytes, append its reachable positions (a short remainder moves to the front) ----
				if (off > 0) {
					int mv[kBwdPerThread];
					uint8_t mb[kBwdPerThread];
#pragma unroll
					for (int q = 0; q < kBwdPerThread; q++) {
						const int j = q * kBwdThreads + tid;
						mv[q] = (j < have) ? s_pos[off + j] : 0;
						mb[q] = (j < have) ? s_posb[off + j] : (uint8_t)0;
					}
					__syncthreads();
#pragma unroll
					for (int q = 0; q < kBwdPerThread; q++) {
						const int j = q * kBwdThreads + tid;
						if (j < have) { s_pos[j] = mv[q]; s_posb[j] = mb[q]; }
					}
					off = 0;
				}
				constexpr int kPer = kBwdChunk / kBwdThreads;
				uint32_t hb[kPer];
				int my_keep = 0;
				const int first_pos = n - 1 - (next_chunk * kBwdChunk + tid * kPer);   // this thread's highest list position
#pragma unroll
				for (int q = 0; q < kPer; q++) {
					hb[q] = (first_pos - q >= 0) ? (uint32_t)hit_bytes[range.x + first_pos - q] : 0u;
					my_keep += hb[q] != 0u ? 1 : 0;
				}
				int incl = my_keep;
#pragma unroll
				for (int o = 1; o < 32; o <<= 1) {
					const int u = __shfl_up_sync(0xffffffffu, incl, o);
					if (lane >= o) incl += u;
				}
				__syncthreads();   // s_warp_cnt of the previous scan has been read by everybody
				if (lane == 31) s_warp_cnt[warp] = incl;
				__syncthreads();
				int slot = have + incl - my_keep + (warp == 1 ? s_warp_cnt[0] : 0);
				const int kept = s_warp_cnt[0] + s_warp_cnt[1];
#pragma unroll
				for (int q = 0; q < kPer; q++) {
					if (hb[q] != 0u) {
						s_pos[slot] = first_pos - q;
						s_posb[slot] = (uint8_t)hb[q];
						slot++;
					}
				}
				have += kept;
				next_chunk++;
				__syncthreads();
				continue;
			}
			if (have == 0) break;
			// ---- gather and replay one batch ----
			const int total = min(kBwdBatch, have);
#pragma unroll
			for (int q = 0; q < kBwdPerThread; q++) {
				const int j = q * kBwdThreads + tid;   // s_pos is in descending list order, so is s_e
				if (j < total) {
					const int pos = s_pos[off + j];
					const uint32_t id = point_list[range.x + pos];
					float4 a = g0[id];
					const float4 b = g1[id];
					const float2 bt = gb[id];
					if (wrap_W > 0.f) a.x = nearest_copy_x(a.x, tx0 + 0.5f * (kTile - 1), wrap_W);
					s_e[j].a = a;
					s_e[j].b = make_float4(b.x, bt.y, b.y, __int_as_float(pos));
					s_e[j].c = make_float4(b.z, b.w, bt.x, __uint_as_float(id));
					s_mask[j] = s_posb[off + j];
				}
			}
			__syncthreads();
			replay(total);
			__syncthreads();   // s_e is rewritten by the next batch
			off += total;
			have -= total;
		}
	} else {
		for (int round = 0; round < rounds; round++) {
			// ---- gather (reverse list order, 4 consecutive entries per thread), tile-level cull ----
			float4 a[kBwdPerThread], b[kBwdPerThread];
			float cb[kBwdPerThread], tau[kBwdPerThread];
			uint32_t id[kBwdPerThread];
			int pos[kBwdPerThread];
			bool keep[kBwdPerThread];
			int my_keep = 0;
			uint32_t hb[kBwdPerThread];
#pragma unroll
			for (int q = 0; q < kBwdPerThread; q++) {
				pos[q] = n - 1 - (round * kBwdBatch + tid * kBwdPerThread + q);   // 0-based list position
				hb[q] = (kHits && pos[q] >= 0) ? (uint32_t)hit_bytes[range.x + pos[q]] : 0u;
			}
#pragma unroll
			for (int q = 0; q < kBwdPerThread; q++)
				id[q] = (pos[q] >= 0 && (!kHits || hb[q] != 0u)) ? point_list[range.x + pos[q]] : 0u;
#pragma unroll
			for (int q = 0; q < kBwdPerThread; q++) {
				keep[q] = false;
				if (pos[q] >= 0 && (!kHits || hb[q] != 0u)) {
					a[q] = g0[id[q]];
					b[q] = g1[id[q]];
					const float2 bt = gb[id[q]];
					cb[q] = bt.x;
					if (wrap_W > 0.f) a[q].x = nearest_copy_x(a[q].x, tx0 + 0.5f * (kTile - 1), wrap_W);
					tau[q] = bt.y;
					keep[q] = kHits ? true : gaussian_touches_box(a[q].x, a[q].y, a[q].z, a[q].w, b[q].x, tau[q], tx0, ty0, tx1, ty1);
				}
				my_keep += keep[q] ? 1 : 0;
			}
			// stable compaction: thread-major order == descending list position
			int incl = my_keep;
#pragma unroll
			for (int o = 1; o < 32; o <<= 1) {
				const int u = __shfl_up_sync(0xffffffffu, incl, o);
				if (lane >= o) incl += u;
			}
			if (lane == 31) s_warp_cnt[warp] = incl;
			__syncthreads();   // every warp has finished replaying the previous round's s_e
			int slot = incl - my_keep + (warp == 1 ? s_warp_cnt[0] : 0);
			const int total = s_warp_cnt[0] + s_warp_cnt[1];
#pragma unroll
			for (int q = 0; q < kBwdPerThread; q++) {
				if (keep[q]) {
					s_e[slot].a = a[q];
					s_e[slot].b = make_float4(b[q].x, tau[q], b[q].y, __int_as_float(pos[q]));
					s_e[slot].c = make_float4(b[q].z, b[q].w, cb[q], __uint_as_float(id[q]));
					if (kHits) s_mask[slot] = (uint8_t)hb[q];
					slot++;
				}
			}
			__syncthreads();

			replay(total);
		}
	}
	OGS_TILE_CLOCK(OGS_BWD_CLOCK, tile, 1);
}

#ifdef OGS_TILE_TIMELINE
extern "C" __attribute__((visibility("default"))) int ogs_debug_set_bwd_tile_clock(unsigned long long* p)
{
	return cudaMemcpyToSymbol(g_bwd_tile_clock, &p, sizeof(p)) == cudaSuccess ? 0 : -2;
}
#endif

int launch_render_bwd(const uint2* ranges, const uint32_t* point_list, int W, int H, const float* bg,
                      const float4* g0, const float4* g1, const float2* gb, const unsigned long long* scalars,
                      const float* final_T, const uint32_t* n_contrib, const float* dL_dpix,
                      float* grad_acc, const uint8_t* hit_bytes, cudaStream_t st)
{
	const int gx = ceil_div(W, kTile), gy = ceil_div(H, kTile);
	// resident CTAs per SM the kernel is compiled for (register budget); OGS_BWD_MINBLOCKS is a tuning knob
	static const int variant = [] { const char* e = getenv("OGS_BWD_MINBLOCKS"); return e ? atoi(e) : 16; }();
	static const int order = [] { const char* e = getenv("OGS_TILE_ORDER"); return e ? atoi(e) : 0; }();
	const bool hits = render_hit_bytes_enabled() && hit_bytes != nullptr;
#define OGS_BWD_LAUNCH(MB)                                                                                                         \
	do {                                                                                                                           \
		if (hits) render_bwd_kernel<MB, true><<<gx * gy, kBwdThreads, 0, st>>>(ranges, point_list, W, H, gx, gy, order, bg, g0, g1, gb, \
		                                                                       scalars, final_T, n_contrib, dL_dpix, grad_acc, hit_bytes); \
		else render_bwd_kernel<MB, false><<<gx * gy, kBwdThreads, 0, st>>>(ranges, point_list, W, H, gx, gy, order, bg, g0, g1, gb,   \
		                                                                    scalars, final_T, n_contrib, dL_dpix, grad_acc, hit_bytes); \
	} while (0)
	switch (variant) {
	case 8: OGS_BWD_LAUNCH(8); break;
	case 10: OGS_BWD_LAUNCH(10); break;
	case 12: OGS_BWD_LAUNCH(12); break;
	case 14: OGS_BWD_LAUNCH(14); break;
	default: OGS_BWD_LAUNCH(16); break;
	}
#undef OGS_BWD_LAUNCH
	OGS_CUDA_TRY(cudaGetLastError());
	return OGS_OK;
}

} // namespace ogs
