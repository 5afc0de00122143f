// render_bwd.cu — 16x16-tile back-to-front replay of the blend (backward).
//
// Per-pixel arithmetic follows the reference's backward renderCUDA
// (cuda_rasterizer/backward.cu:672-843): T is un-multiplied by division, accum_rec / last_alpha /
// last_color carry the colour behind the current Gaussian, and the nine partial derivatives
// (dL/dmean2D.xy in NDC units, dL/dconic .x .y .w, dL/dopacity, dL/dcolour rgb) use the same
// expressions.  The reference then issues 9 global float atomics per (pixel, Gaussian) pair
// (backward.cu:805,829-840).  Here instead:
//   * the tile list is cut at the tile's largest n_contrib (nothing behind it was blended),
//   * entries are tile-culled and compacted like in the forward, each warp handles an 8x4 sub-tile
//     and iterates only over entries that can reach it,
//   * the nine partials are summed over the warp's 32 pixels with a transposing butterfly
//     (14 shuffles for 9 values instead of 45), accumulated across the CTA's 8 warps in shared
//     memory, and flushed with ONE vectorised L2 reduction set per (tile, Gaussian):
//     2 x red.global.add.v4.f32 + 1 x red.global.add.f32 into a packed 12-float accumulator row.
// Summation order differs from the reference's atomics (which are unordered anyway); parity is
// within the stated 1e-4 relative tolerance.
#include "render_common.cuh"
#include "launchers.cuh"

namespace ogs {

constexpr int kAccStride = 9;

// Sum v[0..7] and v8 over the 32 lanes.  On return lane L holds in `z` the total of value (L>>2)
// (replicated over the 4 lanes of a quad) and every lane holds the total of v8 in `z8`.
OGS_D void warp_transpose_reduce9(const float (&v)[8], float v8, float& z, float& z8)
{
	const unsigned full = 0xffffffffu;
	const int lane = threadIdx.x & 31;
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

__global__ void __launch_bounds__(kRenderThreads) render_bwd_kernel(
	const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list, int W, int H, int gx,
	const float* __restrict__ bg_color,
	const float4* __restrict__ g0, const float4* __restrict__ g1, const float* __restrict__ gb,
	const float* __restrict__ final_Ts, const uint32_t* __restrict__ n_contrib,
	const float* __restrict__ dL_dpixels, float* __restrict__ grad_acc /*[P,12]*/)
{
	__shared__ StagedEntry s_e[kBatch];
	__shared__ float s_acc[kBatch * kAccStride];
	__shared__ uint32_t s_warp_cnt[kRenderThreads / 32];
	__shared__ int s_max_contrib;

	const int tile = blockIdx.x;
	const int tile_x = tile % gx, tile_y = tile / gx;
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int sub_x0 = tile_x * kTile + (warp & 1) * kSubW;
	const int sub_y0 = tile_y * kTile + (warp >> 1) * kSubH;
	const int px = sub_x0 + (lane & (kSubW - 1));
	const int py = sub_y0 + (lane / kSubW);
	const bool inside = px < W && py < H;
	const size_t pix_id = (size_t)W * py + px;
	const size_t HW = (size_t)H * W;
	const float2 pixf = { (float)px, (float)py };

	const float tx0 = (float)(tile_x * kTile), ty0 = (float)(tile_y * kTile);
	const float tx1 = tx0 + (kTile - 1), ty1 = ty0 + (kTile - 1);
	const float sx0 = (float)sub_x0, sy0 = (float)sub_y0;
	const float sx1 = sx0 + (kSubW - 1), sy1 = sy0 + (kSubH - 1);

	const uint2 range = ranges[tile];

	// per-pixel state (backward.cu:717-740)
	const float T_final = inside ? final_Ts[pix_id] : 0;
	float T = T_final;
	const int last_contributor = inside ? (int)n_contrib[pix_id] : 0;
	float accum_rec[3] = { 0.f, 0.f, 0.f };
	float dL_dpixel[3] = { 0.f, 0.f, 0.f };
	if (inside) {
#pragma unroll
		for (int ch = 0; ch < 3; ch++) dL_dpixel[ch] = dL_dpixels[ch * HW + pix_id];
	}
	float last_alpha = 0.f;
	float last_color[3] = { 0.f, 0.f, 0.f };
	const float ddelx_dx = 0.5 * W;
	const float ddely_dy = 0.5 * H;
	float bg_dot_dpixel = 0.f;
#pragma unroll
	for (int ch = 0; ch < 3; ch++) bg_dot_dpixel += bg_color[ch] * dL_dpixel[ch];
	const float neg_Tfinal_bg = -T_final * bg_dot_dpixel;   // (-T_final / (1 - alpha)) * bg_dot = this * 1/(1 - alpha)

	// entries at list positions >= max(n_contrib) of the tile were blended by no pixel
	if (threadIdx.x == 0) s_max_contrib = 0;
	__syncthreads();
	if (last_contributor > 0) atomicMax(&s_max_contrib, last_contributor);
	__syncthreads();
	const int n = min((int)(range.y - range.x), s_max_contrib);
	const int rounds = (n + kBatch - 1) / kBatch;

	for (int round = 0; round < rounds; round++) {
		// ---- gather (reverse list order), tile-level cull, stable compaction ----
		const int i = n - 1 - (round * kBatch + (int)threadIdx.x); // 0-based list position
		bool keep = false;
		float4 a = make_float4(0, 0, 0, 0), b = a;
		float cb = 0.f, tau = 0.f;
		uint32_t id = 0;
		if (i >= 0) {
			id = point_list[range.x + i];
			a = g0[id];
			b = g1[id];
			cb = gb[id];
			tau = alpha_power_threshold(b.y);
			keep = gaussian_touches_box(a.x, a.y, a.z, a.w, b.x, tau, tx0, ty0, tx1, ty1);
		}
		int total;
		const int slot = block_compact_slot(keep, s_warp_cnt, total); // contains a __syncthreads
		if (keep) {
			s_e[slot].a = a;
			s_e[slot].b = make_float4(b.x, tau, b.y, __int_as_float(i));   // 0-based list position
			s_e[slot].c = make_float4(b.z, b.w, cb, __uint_as_float(id));
		}
		for (int k = threadIdx.x; k < total * kAccStride; k += kRenderThreads) s_acc[k] = 0.f;
		__syncthreads();

		// ---- per-warp replay over the entries that can reach this sub-tile ----
		for (int base = 0; base < total; base += 32) {
			const int s = base + lane;
			bool hit = false;
			if (s < total) {
				const float4 ea = s_e[s].a;
				const float4 eb = s_e[s].b;
				hit = gaussian_touches_box(ea.x, ea.y, ea.z, ea.w, eb.x, eb.y, sx0, sy0, sx1, sy1);
			}
			unsigned m = __ballot_sync(0xffffffffu, hit);
			while (m) {
				const int j = base + __ffs(m) - 1;
				m &= m - 1;
				const StagedEntry* e = &s_e[j];
				const float4 ea = e->a;
				const float4 eb = e->b;

				// backward.cu:766-782: same guards (and the same pinned arithmetic) as the forward
				float2 d;
				const float power = pair_power(ea.x, ea.y, ea.z, ea.w, eb.x, pixf, d.x, d.y);
				bool valid = inside && (__float_as_int(eb.w) < last_contributor) && !(power > 0.0f) && !(power < eb.y);
				float G = 0.f, alpha = 0.f;
				if (valid) {
					G = expf(power);
					alpha = fminf(0.99f, __fmul_rn(eb.z, G));
					valid = !(alpha < kAlphaMin);
				}
				if (!__any_sync(0xffffffffu, valid)) continue;

				float v[8] = { 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f };
				float v8 = 0.f;
				if (valid) {
					// backward.cu:784-840.  One correctly rounded reciprocal replaces the reference's two
					// divisions by (1 - alpha): T/(1-a) and -T_final/(1-a) (difference <= 1 ulp, inside the
					// 1e-4 gradient tolerance).
					const float4 ec = e->c;
					const float inv = __frcp_rn(1.f - alpha);
					T = T * inv;
					const float dchannel_dcolor = alpha * T;
					float dL_dalpha = 0.0f;
					const float col[3] = { ec.x, ec.y, ec.z };
#pragma unroll
					for (int ch = 0; ch < 3; ch++) {
						const float c = col[ch];
						accum_rec[ch] = last_alpha * last_color[ch] + (1.f - last_alpha) * accum_rec[ch];
						last_color[ch] = c;
						dL_dalpha += (c - accum_rec[ch]) * dL_dpixel[ch];
					}
					v[6] = dchannel_dcolor * dL_dpixel[0];
					v[7] = dchannel_dcolor * dL_dpixel[1];
					v8 = dchannel_dcolor * dL_dpixel[2];
					dL_dalpha *= T;
					last_alpha = alpha;
					dL_dalpha += neg_Tfinal_bg * inv;

					const float dL_dG = eb.z * dL_dalpha;
					const float gdx = G * d.x;
					const float gdy = G * d.y;
					const float dG_ddelx = -gdx * ea.z - gdy * ea.w;
					const float dG_ddely = -gdy * eb.x - gdx * ea.w;
					v[0] = dL_dG * dG_ddelx * ddelx_dx;
					v[1] = dL_dG * dG_ddely * ddely_dy;
					v[2] = -0.5f * gdx * d.x * dL_dG;
					v[3] = -0.5f * gdx * d.y * dL_dG;
					v[4] = -0.5f * gdy * d.y * dL_dG;
					v[5] = G * dL_dalpha;
				}
				float z, z8;
				warp_transpose_reduce9(v, v8, z, z8);
				// lanes 0,4,..,28 hold sums 0..7, lane 1 takes the ninth: one shared-memory atomic pass
				const bool ninth = (lane == 1);
				if ((lane & 3) == 0 || ninth)
					atomicAdd(&s_acc[j * kAccStride + (ninth ? 8 : (lane >> 2))], ninth ? z8 : z);
			}
		}
		__syncthreads();

		// ---- flush: one packed reduction set per (tile, Gaussian) ----
		if ((int)threadIdx.x < total) {
			const float* acc = &s_acc[threadIdx.x * kAccStride];
			float r[9];
			bool nz = false;
#pragma unroll
			for (int k = 0; k < 9; k++) {
				r[k] = acc[k];
				nz |= (r[k] != 0.f);
			}
			if (nz) {
				float* dst = grad_acc + (size_t)__float_as_uint(s_e[threadIdx.x].c.w) * 12;
				red_add_v4(dst, r[0], r[1], r[2], r[3]);
				red_add_v4(dst + 4, r[4], r[5], r[6], r[7]);
				red_add(dst + 8, r[8]);
			}
		}
		__syncthreads();
	}
}

int launch_render_bwd(const uint2* ranges, const uint32_t* point_list, int W, int H, const float* bg,
                      const float4* g0, const float4* g1, const float* gb,
                      const float* final_T, const uint32_t* n_contrib, const float* dL_dpix,
                      float* grad_acc, cudaStream_t st)
{
	const int gx = ceil_div(W, kTile), gy = ceil_div(H, kTile);
	render_bwd_kernel<<<gx * gy, kRenderThreads, 0, st>>>(ranges, point_list, W, H, gx, bg, g0, g1, gb,
	                                                     final_T, n_contrib, dL_dpix, grad_acc);
	OGS_CUDA_TRY(cudaGetLastError());
	return OGS_OK;
}

} // namespace ogs
