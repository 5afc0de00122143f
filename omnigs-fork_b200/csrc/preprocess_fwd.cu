// preprocess_fwd.cu — per-Gaussian forward preprocessing for the lonlat camera.
//
// One thread per Gaussian.  Follows the reference's preprocessLonlatCUDA step for step
// (cuda_rasterizer/forward.cu:593-703) so that radii / tile rects / depth keys are bit-identical,
// but writes our own packed, gather-friendly records instead of the reference's GeometryState and
// fuses the bookkeeping the reference does in later kernels:
//   * sum of tiles_touched (= num_rendered) via one 64-bit atomic per block (replaces the
//     InclusiveSum + device->host read of its last element, rasterizer_impl.cu:622-628),
//   * the 2-D tile-coverage difference array (4 atomics per visible Gaussian) from which the tile
//     ranges and the radix digit histograms are derived without touching the R-sized key list,
//   * the depth-sort key (float bits of r; 0xFFFFFFFF for Gaussians that emit nothing).
// The 192-byte SH rows of the CTA are fetched by the bulk-copy engine (cp.async.bulk + mbarrier) while
// the threads do the projection math; see async_copy.cuh.
//
// Raw-parameter mode: the kernel reads the model's stored tensors and applies the
// activations the reference applies with separate LibTorch ops before every render
// (gaussian_model.cpp:54-77 via gaussian_renderer.cpp:212-258): opacity = sigmoid(opacity_),
// scales = exp(scaling_), rotations = normalize(rotation_), SH row = cat(features_dc_, features_rest_).
// The CTA's features_rest_ rows (128 x 180 B, contiguous) and features_dc_ rows (128 x 12 B) come in as
// two bulk copies; a thread reads its row with a 45-word stride (conflict-free).
#include "pinhole_math.cuh"
#include "render_common.cuh"
#include "launchers.cuh"
#include "async_copy.cuh"

namespace ogs {

constexpr int kPreThreads = 128;

// Out of line: keeps the expf/logf bodies of the ulp walk out of the (long, register-hungry) projection code.
__device__ __noinline__ float alpha_cutoff_power_call(float opacity) { return alpha_cutoff_power(opacity); }

// kMode: 0 SH rows by plain loads, 1 SH rows [P,16,3] by per-row bulk copies,
//        2 raw parameters with split SH by per-CTA bulk copies, 3 raw parameters by plain loads.
// kPinhole: the perspective camera (camera_type 1; forward.cu:232-340) instead of the equirectangular one.
template <int kMode, bool kPinhole = false>
__global__ void __launch_bounds__(kPreThreads) preprocess_lonlat_fwd_kernel(const PreprocessFwdArgs a)
{
	constexpr bool kBulkSH = (kMode == 1 || kMode == 2);
	constexpr bool kRaw = (kMode >= 2);
	__shared__ float sV[16];
	__shared__ float sP[kPinhole ? 16 : 1];
	__shared__ float sCam[3];
	__shared__ unsigned long long s_block_tiles;
	__shared__ unsigned int s_block_cols;     // sum of the rects' widths: the number of column segments (binning.cu)
	__shared__ __align__(16) float s_sh[kBulkSH ? kPreThreads * kShPitchFloats : 4];
	__shared__ __align__(8) uint64_t s_bar;

	const int tid = threadIdx.x;
	const int idx = blockIdx.x * kPreThreads + tid;
	if (tid < 16) sV[tid] = a.viewmatrix[tid];
	if (kPinhole && tid < 16) sP[tid] = a.projmatrix[tid];
	if (tid < 3) sCam[tid] = a.campos[tid];
	if (tid == 0) {
		s_block_tiles = 0ull;
		s_block_cols = 0u;
		if (kBulkSH) {
			const int rows = min(kPreThreads, a.P - (int)blockIdx.x * kPreThreads);
			mbar_init(&s_bar, 1);
			if constexpr (kMode == 2) {
				// whole multiples of 4 rows keep both chunk sizes multiples of 16 bytes; <= 3 tail rows load plainly
				const uint32_t rows4 = (uint32_t)rows & ~3u;
				mbar_arrive_expect_tx(&s_bar, rows4 * kShRowFloats * 4u);
				if (rows4) {
					const size_t first = (size_t)blockIdx.x * kPreThreads;
					bulk_load(&s_sh[0], a.features_rest + first * kRawRestFloats, rows4 * kRawRestFloats * 4u, &s_bar);
					bulk_load(&s_sh[kRawDcOffset], a.features_dc + first * 3, rows4 * 12u, &s_bar);
				}
			} else {
				mbar_arrive_expect_tx(&s_bar, (uint32_t)rows * kShRowFloats * 4u);
			}
		}
	}
	__syncthreads();
	if constexpr (kMode == 1) if (idx < a.P)
		bulk_load(&s_sh[tid * kShPitchFloats], a.shs + (size_t)idx * kShRowFloats, kShRowFloats * 4u, &s_bar);
	if constexpr (kMode == 2) if (idx < a.P) {
		const int rows = min(kPreThreads, a.P - (int)blockIdx.x * kPreThreads);
		if (tid >= (rows & ~3)) {
			for (int k = 0; k < kRawRestFloats; k++) s_sh[tid * kRawRestFloats + k] = a.features_rest[(size_t)idx * kRawRestFloats + k];
			for (int k = 0; k < 3; k++) s_sh[kRawDcOffset + tid * 3 + k] = a.features_dc[(size_t)idx * 3 + k];
		}
	}

	uint32_t my_tiles = 0;
	int out_radius = 0;
	uint32_t key = 0xFFFFFFFFu;
	bool emits = false;           // visible inside this band: needs tile instances
	bool visible = false;         // visible in the full frame (radii > 0): needs its SH clamp mask
	float3 p_orig = { 0.f, 0.f, 0.f }, conic = { 0.f, 0.f, 0.f };
	float2 point_image = { 0.f, 0.f };
	float r = 0.f;
	int x0 = 0, x1 = 0, by0 = 0, by1 = 0;

	if (idx < a.P) {
		float V[16];
#pragma unroll
		for (int i = 0; i < 16; i++) V[i] = sV[i];

		p_orig = { a.means3D[3 * idx], a.means3D[3 * idx + 1], a.means3D[3 * idx + 2] };
		const float3 t = view_point_p(V, p_orig);
		bool in_view;
		float2 p_proj;
		if constexpr (kPinhole) {
			// in_frustum (auxiliary.h:166-196): only the near plane culls; depth is the camera-space z
			in_view = !(t.z <= 0.2f);
			r = t.z;
			// projection through the full transform (forward.cu:273-277)
			const float4 p_hom = proj_point_p(sP, p_orig);
			const float p_w = __frcp_rn(__fadd_rn(p_hom.w, kEps7));
			p_proj = { __fmul_rn(p_hom.x, p_w), __fmul_rn(p_hom.y, p_w) };
		} else {
			// near cull (auxiliary.h:198-220): r^2 <= 0.04 drops the Gaussian
			const float rr = dot3p(t.x, t.x, t.y, t.y, t.z, t.z);
			in_view = !(rr <= 0.04f);
			r = __fsqrt_rn(rr);
			// lon/lat screen coordinates (auxiliary.h:236-248)
			const float inv_r = __frcp_rn(__fadd_rn(r, kEps7));
			const float lon = atan2f(t.x, t.z);
			const float lat = asinf(__fmul_rn(t.y, inv_r));
			p_proj = { __fmul_rn(lon, kPiInv), __fmul_rn(lat, kTwoPiInv) };
		}
		if (in_view) {
			// 3-D covariance (forward.cu:643-652)
			float cov6[6];
			if (a.cov3D_precomp != nullptr) {
#pragma unroll
				for (int i = 0; i < 6; i++) cov6[i] = a.cov3D_precomp[6 * (size_t)idx + i];
			} else {
				float3 sc = { a.scales[3 * idx], a.scales[3 * idx + 1], a.scales[3 * idx + 2] };
				float4 q = reinterpret_cast<const float4*>(a.rotations)[idx];
				if (kRaw) {
					sc = { expf(sc.x), expf(sc.y), expf(sc.z) };   // getScalingActivation, gaussian_model.cpp:54-57
					q = normalize_quat(q);                          // getRotationActivation, :59-62
				}
				cov3d_from_scale_rot_p(sc, a.scale_modifier, q, cov6);
#pragma unroll
				for (int i = 0; i < 6; i++) a.cov3D[6 * (size_t)idx + i] = cov6[i];
			}

			// 2-D covariance through the camera's Jacobian + 0.3 px blur (forward.cu:130-189; pinhole :86-128)
			const float3 cov = kPinhole ? cov2d_pinhole_p(t, V, cov6, a.focal_x, a.focal_y, a.tan_fovx, a.tan_fovy)
			                            : cov2d_lonlat_p(t, V, cov6, a.W, a.H);

			// conic (forward.cu:660-664)
			const float det = __fmaf_rn(cov.x, cov.z, -__fmul_rn(cov.y, cov.y));
			if (det != 0.0f) {
				const float det_inv = __frcp_rn(det);
				conic = { __fmul_rn(cov.z, det_inv), __fmul_rn(cov.y, -det_inv), __fmul_rn(cov.x, det_inv) };

				// screen-space extent (forward.cu:671-683)
				const float mid = __fmul_rn(__fadd_rn(cov.x, cov.z), 0.5f);
				const float sq = __fsqrt_rn(fmaxf(__fmaf_rn(mid, mid, -det), 0.1f));
				const float lambda1 = __fadd_rn(mid, sq);
				const float lambda2 = __fsub_rn(mid, sq);
				const float my_radius = ceilf(__fmul_rn(__fsqrt_rn(fmaxf(lambda1, lambda2)), 3.f));
				point_image = { ndc_to_pix_p(p_proj.x, a.W), ndc_to_pix_p(p_proj.y, a.H) };
				int y0, y1;
				tile_rect_p(point_image, (int)my_radius, a.gx, a.gy, x0, y0, x1, y1);
				if (a.seam_wrap) {
					// opt-in extension: x range modulo the tile grid (cf. the reference's unused getRectCyclic,
					// auxiliary.h:68-83).  x0 in [0, gx), x1 = x0 + width <= x0 + gx.
					const float R = (float)(int)my_radius;
					// same expressions as getRect (auxiliary.h:59,63) with floor instead of clamp-to-zero truncation,
					// so rects that do not touch the seam are identical in both modes
					const int xa = (int)floorf(__fmul_rn(__fsub_rn(point_image.x, R), 1.0f / kTile));
					const int xb = (int)floorf(__fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(point_image.x, R), (float)kTile), -1.0f), 1.0f / kTile));
					const int w = min(xb - xa, a.gx);
					x0 = ((xa % a.gx) + a.gx) % a.gx;
					x1 = x0 + w;
				}
				if ((x1 - x0) * (y1 - y0) != 0) {
					out_radius = (int)my_radius;
					visible = true;
					// latitude-band clip (identity for the full image)
					by0 = max(y0, a.band_y0);
					by1 = min(y1, a.band_y1);
					emits = by1 > by0;
				}
			}
		}
	}

	// every thread waits for the CTA's SH rows (a CTA must not retire with bulk copies in flight)
	if (kBulkSH) mbar_wait(&s_bar, 0);

	// The clamp mask is per-Gaussian forward state the per-Gaussian backward needs for EVERY visible
	// Gaussian, also when this rank's band does not contain it (accumulators may arrive from other ranks).
	if (visible) {
		float3 rgb;
		unsigned cmask = 0;
		if (a.defer_colors) {
			rgb = { 0.f, 0.f, 0.f };   // launch_sh_colors fills colour and clamp mask before the blend
		} else if (a.colors_precomp == nullptr) {
			V3 c;
			if constexpr (kMode == 2) {
				float shr[kShRowFloats];
#pragma unroll
				for (int k = 0; k < 3; k++) shr[k] = s_sh[kRawDcOffset + tid * 3 + k];
#pragma unroll
				for (int k = 0; k < kRawRestFloats; k++) shr[3 + k] = s_sh[tid * kRawRestFloats + k];
				auto sh = [&shr](int k) { return V3{ shr[3 * k], shr[3 * k + 1], shr[3 * k + 2] }; };
				c = sh_to_rgb_p(a.D, p_orig, float3{ sCam[0], sCam[1], sCam[2] }, sh, cmask);
			} else if constexpr (kMode == 3) {
				const float* dc = a.features_dc + (size_t)idx * 3;
				const float* rest = a.features_rest + (size_t)idx * (a.M - 1) * 3;
				auto sh = [dc, rest](int k) {
					const float* p = k == 0 ? dc : rest + 3 * (k - 1);
					return V3{ p[0], p[1], p[2] };
				};
				c = sh_to_rgb_p(a.D, p_orig, float3{ sCam[0], sCam[1], sCam[2] }, sh, cmask);
			} else if constexpr (kBulkSH) {
				float shr[kShRowFloats];
				const float4* row = reinterpret_cast<const float4*>(&s_sh[tid * kShPitchFloats]);
#pragma unroll
				for (int k = 0; k < kShRowFloats / 4; k++) {
					const float4 q = row[k];
					shr[4 * k] = q.x; shr[4 * k + 1] = q.y; shr[4 * k + 2] = q.z; shr[4 * k + 3] = q.w;
				}
				auto sh = [&shr](int k) { return V3{ shr[3 * k], shr[3 * k + 1], shr[3 * k + 2] }; };
				c = sh_to_rgb_p(a.D, p_orig, float3{ sCam[0], sCam[1], sCam[2] }, sh, cmask);
			} else {
				const float* shp = a.shs + (size_t)idx * a.M * 3;
				auto sh = [shp](int k) { return V3{ shp[3 * k], shp[3 * k + 1], shp[3 * k + 2] }; };
				c = sh_to_rgb_p(a.D, p_orig, float3{ sCam[0], sCam[1], sCam[2] }, sh, cmask);
			}
			rgb = { c.x, c.y, c.z };
		} else {
			rgb = { a.colors_precomp[3 * (size_t)idx], a.colors_precomp[3 * (size_t)idx + 1],
			        a.colors_precomp[3 * (size_t)idx + 2] };
		}
		if (kPinhole && a.render_depth) rgb = { r, r, r };   // renderDepthCUDA blends the depth in all channels (forward.cu:566-567)
		a.clamped[idx] = (uint8_t)cmask;
		// so is the packed record: its conic turns the summed raw accumulators into gradients (preprocess_bwd.cu)
		a.g0[idx] = make_float4(point_image.x, point_image.y, conic.x, conic.y);
		float opacity = a.opacities[idx];
		if (kRaw) opacity = sigmoid_act(opacity);   // getOpacityActivation, gaussian_model.cpp:74-77
		a.g1[idx] = make_float4(conic.z, opacity, rgb.x, rgb.y);
		a.gb[idx] = make_float2(rgb.z, alpha_cutoff_power_call(opacity));
		if (emits) {
			a.depth[idx] = r;
			a.rect[idx] = make_uint2((uint32_t)x0 | ((uint32_t)x1 << 16), (uint32_t)by0 | ((uint32_t)by1 << 16));
			my_tiles = (uint32_t)((by1 - by0) * (x1 - x0));
			key = __float_as_uint(r);
			const int pitch = a.gx + 1;
			const int xe = min(x1, a.gx);   // x1 > gx only for a rect that wraps around the seam
			atomicAdd(&a.tile_diff[by0 * pitch + x0], 1);
			atomicAdd(&a.tile_diff[by0 * pitch + xe], -1);
			atomicAdd(&a.tile_diff[by1 * pitch + x0], -1);
			atomicAdd(&a.tile_diff[by1 * pitch + xe], 1);
			if (x1 > a.gx) {               // the wrapped part [0, x1 - gx)
				atomicAdd(&a.tile_diff[by0 * pitch], 1);
				atomicAdd(&a.tile_diff[by0 * pitch + (x1 - a.gx)], -1);
				atomicAdd(&a.tile_diff[by1 * pitch], -1);
				atomicAdd(&a.tile_diff[by1 * pitch + (x1 - a.gx)], 1);
			}
		}
	}
	if (idx < a.P) {
		a.radii[idx] = out_radius;
		a.tiles_touched[idx] = my_tiles;
		a.sort_key[idx] = key;
	}

	// block total of tiles_touched -> one 64-bit atomic
	uint32_t wsum = my_tiles, csum = emits ? (uint32_t)(x1 - x0) : 0u;
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) {
		wsum += __shfl_xor_sync(0xffffffffu, wsum, o);
		csum += __shfl_xor_sync(0xffffffffu, csum, o);
	}
	if ((tid & 31) == 0 && wsum) {
		atomicAdd(&s_block_tiles, (unsigned long long)wsum);
		atomicAdd(&s_block_cols, csum);
	}
	__syncthreads();
	if (tid == 0 && s_block_tiles) {
		atomicAdd(a.total_tiles, s_block_tiles);
		atomicAdd(a.total_tiles + 1, (unsigned long long)s_block_cols);
	}
	if (tid == 0 && blockIdx.x == 0) a.total_tiles[7] = (unsigned long long)a.seam_wrap;
}

__global__ void mark_all_visible_kernel(int P, uint8_t* present)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < P) present[i] = 1;
}

// Deferred colour evaluation (data-parallel trainer): SH -> RGB and the clamp mask of every Gaussian that is visible in
// the full frame, written into the packed records the blend kernels gather (computeColorFromSH at its call site,
// forward.cu:688-692, pinned like the fused path: sh_to_rgb_p).  SH rows by per-row bulk copies when M == 16.
template <bool kBulk>
__global__ void __launch_bounds__(kPreThreads) sh_colors_kernel(
	int P, int D, int M, const float* __restrict__ means3D, const float* __restrict__ shs, const float* __restrict__ campos,
	const int* __restrict__ radii, float4* __restrict__ g1, float2* __restrict__ gb, uint8_t* __restrict__ clamped)
{
	__shared__ __align__(16) float s_sh[kBulk ? kPreThreads * kShPitchFloats : 4];
	__shared__ __align__(8) uint64_t s_bar;
	__shared__ int s_rows;
	const int tid = threadIdx.x;
	const int idx = blockIdx.x * kPreThreads + tid;
	const bool visible = idx < P && radii[idx] > 0;
	if (kBulk) {
		if (tid == 0) { s_rows = 0; mbar_init(&s_bar, 1); }
		__syncthreads();
		if (visible) atomicAdd(&s_rows, 1);
		__syncthreads();
		if (tid == 0) mbar_arrive_expect_tx(&s_bar, (uint32_t)s_rows * kShRowFloats * 4u);
		__syncthreads();
		if (visible) bulk_load(&s_sh[tid * kShPitchFloats], shs + (size_t)idx * kShRowFloats, kShRowFloats * 4u, &s_bar);
		mbar_wait(&s_bar, 0);   // every thread: a CTA must not retire with copies in flight
	}
	if (!visible) return;
	const float3 p = { means3D[3 * (size_t)idx], means3D[3 * (size_t)idx + 1], means3D[3 * (size_t)idx + 2] };
	const float3 cam = { campos[0], campos[1], campos[2] };
	unsigned cmask = 0;
	V3 c;
	if constexpr (kBulk) {
		float shr[kShRowFloats];
		const float4* row = reinterpret_cast<const float4*>(&s_sh[tid * kShPitchFloats]);
#pragma unroll
		for (int k = 0; k < kShRowFloats / 4; k++) {
			const float4 q = row[k];
			shr[4 * k] = q.x; shr[4 * k + 1] = q.y; shr[4 * k + 2] = q.z; shr[4 * k + 3] = q.w;
		}
		auto sh = [&shr](int k) { return V3{ shr[3 * k], shr[3 * k + 1], shr[3 * k + 2] }; };
		c = sh_to_rgb_p(D, p, cam, sh, cmask);
	} else {
		const float* shp = shs + (size_t)idx * M * 3;
		auto sh = [shp](int k) { return V3{ shp[3 * k], shp[3 * k + 1], shp[3 * k + 2] }; };
		c = sh_to_rgb_p(D, p, cam, sh, cmask);
	}
	clamped[idx] = (uint8_t)cmask;
	float4 r1 = g1[idx];
	r1.z = c.x; r1.w = c.y;
	g1[idx] = r1;
	float2 rb = gb[idx];
	rb.x = c.z;
	gb[idx] = rb;
}

int launch_sh_colors(int P, int D, int M, const float* means3D, const float* shs, const float* campos, const int* radii,
                     float4* g1, float2* gb, uint8_t* clamped, cudaStream_t st)
{
	if (P <= 0) return OGS_OK;
	const int blocks = ceil_div(P, kPreThreads);
	if (sh_rows_bulk_capable(shs, M))
		sh_colors_kernel<true><<<blocks, kPreThreads, 0, st>>>(P, D, M, means3D, shs, campos, radii, g1, gb, clamped);
	else
		sh_colors_kernel<false><<<blocks, kPreThreads, 0, st>>>(P, D, M, means3D, shs, campos, radii, g1, gb, clamped);
	OGS_CUDA_TRY(cudaGetLastError());
	return OGS_OK;
}

int launch_preprocess_fwd(const PreprocessFwdArgs& a, cudaStream_t st)
{
	const int blocks = ceil_div(a.P, kPreThreads);
	if (a.pinhole) {
		if (a.shs != nullptr && sh_rows_bulk_capable(a.shs, a.M))
			preprocess_lonlat_fwd_kernel<1, true><<<blocks, kPreThreads, 0, st>>>(a);
		else
			preprocess_lonlat_fwd_kernel<0, true><<<blocks, kPreThreads, 0, st>>>(a);
	} else if (a.raw) {
		if (sh_rows_bulk_capable(a.features_rest, a.M) && sh_rows_bulk_capable(a.features_dc, a.M))
			preprocess_lonlat_fwd_kernel<2><<<blocks, kPreThreads, 0, st>>>(a);
		else
			preprocess_lonlat_fwd_kernel<3><<<blocks, kPreThreads, 0, st>>>(a);
	} else if (!a.defer_colors && a.shs != nullptr && sh_rows_bulk_capable(a.shs, a.M))
		preprocess_lonlat_fwd_kernel<1><<<blocks, kPreThreads, 0, st>>>(a);
	else
		preprocess_lonlat_fwd_kernel<0><<<blocks, kPreThreads, 0, st>>>(a);
	OGS_CUDA_TRY(cudaGetLastError());
	return OGS_OK;
}

// checkFrustum (rasterizer_impl.cu:64-77 with in_frustum, auxiliary.h:166-196): only z > 0.2 in camera space
__global__ void check_frustum_kernel(int P, const float* __restrict__ means3D, const float* __restrict__ viewmatrix,
                                     uint8_t* present)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= P) return;
	float V[16];
#pragma unroll
	for (int k = 0; k < 16; k++) V[k] = viewmatrix[k];
	const float3 t = view_point_p(V, float3{ means3D[3 * (size_t)i], means3D[3 * (size_t)i + 1], means3D[3 * (size_t)i + 2] });
	present[i] = (t.z <= 0.2f) ? 0 : 1;
}

int launch_check_frustum(int P, const float* means3D, const float* viewmatrix, uint8_t* present, cudaStream_t st)
{
	check_frustum_kernel<<<ceil_div(P, 256), 256, 0, st>>>(P, means3D, viewmatrix, present);
	OGS_CUDA_TRY(cudaGetLastError());
	return OGS_OK;
}

int launch_mark_all_visible(int P, uint8_t* present, cudaStream_t st)
{
	mark_all_visible_kernel<<<ceil_div(P, 256), 256, 0, st>>>(P, present);
	OGS_CUDA_TRY(cudaGetLastError());
	return OGS_OK;
}

} // namespace ogs
