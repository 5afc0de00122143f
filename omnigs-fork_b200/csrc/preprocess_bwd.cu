// preprocess_bwd.cu — fused per-Gaussian backward.
//
// One kernel does what the reference spreads over computeCov2DLonLatCUDA + preprocessLonLatCUDA
// (cuda_rasterizer/backward.cu:297-485, :613-669) plus the zero-fills of its LibTorch shim
// (src/rasterize_points.cu:200-208,246-247):
//   packed render-backward accumulators (raw per-Gaussian sums over pixels, render_bwd.cu) x conic
//   -> dL/dmean2D, dL/dconic, dL/dopacity, dL/dcolour (backward.cu:821-840),
//   conic/covariance branch (incl. the projection's second derivatives) -> dL/dcov3D, dL/dmean,
//   screen-position branch through the Jacobian rows (kept in registers; the reference round-trips
//   them through the dpx_dt / dpy_dt tensors) -> dL/dmean,
//   SH backward -> dL/dsh, dL/dmean;  scale/rotation backward -> dL/dscale, dL/drot.
// Every output element is written, exact zeros for culled Gaussians, so callers may pass
// uninitialised memory.
// With M == 16 the SH rows come in and the dL/dsh rows go out through the bulk-copy engine
// (cp.async.bulk, 192 bytes per Gaussian, 208-byte shared-memory pitch; async_copy.cuh): the thread
// loads its row with conflict-free 128-bit shared loads, overwrites it in place with the gradient row and
// hands it back to the copy engine, so neither direction issues 192-byte-strided global accesses.
//
// Raw-parameter mode: gradients are taken through the model's activations as well
// (sigmoid / exp / normalize / cat, gaussian_model.cpp:54-77), i.e. the kernel returns what LibTorch's
// autograd would deliver to the optimiser's six parameter tensors (xyz_, features_dc_, features_rest_,
// opacity_, scaling_, rotation_), and the dL/dfeatures rows of the CTA leave as two bulk stores.
// Data-parallel / multi-view mode (PreprocessBwdArgs::accumulate, dL_drgb_view, stat_*): the four
// geometry gradients are ADDED to what the step's bucket already holds, the densification statistics of the view
// (gaussian_mapper.cpp:427-434) are folded in by the same thread, and instead of the 192-byte dL/dsh row the kernel
// leaves the view's clamp-masked dL/dRGB (12 bytes): dL/dsh = sum over views of b(dir_view) (x) dL/dRGB_view is rebuilt
// once per step by sh_gradient_from_views_kernel below, bit-identical to the sum of the per-view rows in view order.
#include "pinhole_math.cuh"
#include "gaussian_grad.cuh"
#include "launchers.cuh"
#include "async_copy.cuh"
#include <cstdlib>

namespace ogs {

constexpr int kPreBwdThreads = 128;

// kMode as in preprocess_fwd.cu: 0 plain SH rows, 1 bulk SH rows, 2 raw parameters (bulk), 3 raw parameters (plain)
// kPinhole: the perspective camera's covariance and screen-position branches (backward.cu:156-292, :583-597).
// resident CTAs per SM the kernel is compiled for: 6 (<= 85 registers) measured 0.119 ms at C2 against 0.138 ms unbounded
// (96-110 registers, 4-5 CTAs per SM) — profiles/r02_optimisation_log.md
#ifndef OGS_PREBWD_MINBLOCKS
#define OGS_PREBWD_MINBLOCKS 6
#endif
template <int kMode, bool kPinhole = false>
__global__ void __launch_bounds__(kPreBwdThreads, OGS_PREBWD_MINBLOCKS) preprocess_lonlat_bwd_kernel(const PreprocessBwdArgs a)
{
	constexpr bool kBulkSH = (kMode == 1 || kMode == 2);
	constexpr bool kRaw = (kMode >= 2);
	__shared__ float sV[16];
	__shared__ float sCam[3];
	__shared__ float sP[kPinhole ? 16 : 1];
	__shared__ __align__(16) float s_sh[kBulkSH ? kPreBwdThreads * kShPitchFloats : 4];
	__shared__ __align__(8) uint64_t s_bar;
	__shared__ __align__(8) uint64_t s_rows_done;   // raw mode: counts the CTA's gradient rows written to shared memory
	const int tid = threadIdx.x;
	// a.first_block > 0: a launch over a sub-range of the Gaussians (chunks of a pipelined accumulator exchange)
	const int blk = (int)blockIdx.x + a.first_block;
	const int idx = blk * kPreBwdThreads + tid;
	if (tid < 16) sV[tid] = a.viewmatrix[tid];
	if (tid < 3) sCam[tid] = a.campos[tid];
	if (kPinhole && tid < 16) sP[tid] = a.projmatrix[tid];
	const int rows = min(kPreBwdThreads, a.P - blk * kPreBwdThreads);
	const int rows4 = rows & ~3;   // raw mode: rows covered by the per-CTA bulk copies (sizes stay multiples of 16 B)
	if (kBulkSH && tid == 0) {
		mbar_init(&s_bar, 1);
		if constexpr (kMode == 2) {
			mbar_init(&s_rows_done, (uint32_t)rows);
			mbar_arrive_expect_tx(&s_bar, (uint32_t)rows4 * kShRowFloats * 4u);
			if (rows4) {
				const size_t first = (size_t)blk * kPreBwdThreads;
				bulk_load(&s_sh[0], a.features_rest + first * kRawRestFloats, (uint32_t)rows4 * kRawRestFloats * 4u, &s_bar);
				bulk_load(&s_sh[kRawDcOffset], a.features_dc + first * 3, (uint32_t)rows4 * 12u, &s_bar);
			}
		} else {
			mbar_arrive_expect_tx(&s_bar, (uint32_t)rows * kShRowFloats * 4u);
		}
	}
	__syncthreads();
	if constexpr (kMode == 1) {
		if (idx < a.P)
			bulk_load(&s_sh[tid * kShPitchFloats], a.shs + (size_t)idx * kShRowFloats, kShRowFloats * 4u, &s_bar);
	}
	if constexpr (kMode == 2) {
		if (idx < a.P && tid >= rows4) {
			for (int k = 0; k < kRawRestFloats; k++) s_sh[tid * kRawRestFloats + k] = a.features_rest[(size_t)idx * kRawRestFloats + k];
			for (int k = 0; k < 3; k++) s_sh[kRawDcOffset + tid * 3 + k] = a.features_dc[(size_t)idx * 3 + k];
		}
	}
	if (idx >= a.P) {
		if (kBulkSH) mbar_wait(&s_bar, 0);   // do not retire while the CTA's copies are in flight
		return;
	}

	const int radius = a.radii[idx];
	const bool visible = radius > 0;
	const bool write_sh = a.dL_drgb_view == nullptr;   // data-parallel mode leaves dL/dRGB instead of the dL/dsh row

	float g[9];
	if (visible) {
		// raw sums  Su_dx, Su_dy, Su_dx2, Su_dxdy, Su_dy2 (u = dL/dG * G), S G*dL/dalpha, S colour terms
		const float4* row = reinterpret_cast<const float4*>(a.grad_acc + (size_t)idx * 12);
		const float4 r0 = row[0], r1 = row[1];
		const float4 c0 = a.g0[idx];
		const float A = c0.z, B = c0.w, C = a.g1[idx].x;
		const float ddelx_dx = 0.5 * a.W, ddely_dy = 0.5 * a.H;
		g[0] = -ddelx_dx * (A * r0.x + B * r0.y);
		g[1] = -ddely_dy * (C * r0.y + B * r0.x);
		g[2] = -0.5f * r0.z; g[3] = -0.5f * r0.w;
		g[4] = -0.5f * r1.x; g[5] = r1.y; g[6] = r1.z; g[7] = r1.w;
		g[8] = a.grad_acc[(size_t)idx * 12 + 8];
	} else {
#pragma unroll
		for (int k = 0; k < 9; k++) g[k] = 0.f;
	}

	// the render-backward outputs in the reference's layouts (all optional)
	if (a.dL_dmean2D) {
		a.dL_dmean2D[3 * (size_t)idx + 0] = g[0];
		a.dL_dmean2D[3 * (size_t)idx + 1] = g[1];
		a.dL_dmean2D[3 * (size_t)idx + 2] = 0.f;
	}
	if (a.dL_dconic) reinterpret_cast<float4*>(a.dL_dconic)[idx] = make_float4(g[2], g[3], 0.f, g[4]);
	float dopacity = g[5];
	if constexpr (kRaw) {
		// d sigmoid: the activated opacity is in the packed record (gaussian_model.cpp:74-77)
		const float o = visible ? a.g1[idx].y : 0.f;
		dopacity = g[5] * o * (1.f - o);
	} else if (a.dL_dcolor) {
		a.dL_dcolor[3 * (size_t)idx + 0] = g[6];
		a.dL_dcolor[3 * (size_t)idx + 1] = g[7];
		a.dL_dcolor[3 * (size_t)idx + 2] = g[8];
	}
	// densification statistics of this view (gaussian_mapper.cpp:427-434, gaussian_model.cpp:839-853)
	if (a.stat_grad_norm) {
		const float gn = visible ? sqrtf(g[0] * g[0] + g[1] * g[1]) : 0.f;
		const float one = visible ? 1.f : 0.f, rad = (float)radius;
		if (a.accumulate) {
			a.stat_grad_norm[idx] += gn;
			a.stat_visible[idx] += one;
			a.stat_max_radius[idx] = fmaxf(a.stat_max_radius[idx], rad);
		} else {
			a.stat_grad_norm[idx] = gn;
			a.stat_visible[idx] = one;
			a.stat_max_radius[idx] = rad;
		}
	}

	float dcov6[6] = { 0.f, 0.f, 0.f, 0.f, 0.f, 0.f };
	float3 dmean = { 0.f, 0.f, 0.f };
	float dscale[3] = { 0.f, 0.f, 0.f };
	float drot[4] = { 0.f, 0.f, 0.f, 0.f };
	float drgb[3] = { 0.f, 0.f, 0.f };
	float* dsh_row = (!kRaw && a.dL_dsh && write_sh) ? a.dL_dsh + (size_t)idx * a.M * 3 : nullptr;
	bool sh_waited = false, sh_row_ready = false;

	if (visible) {
		float V[16];
#pragma unroll
		for (int i = 0; i < 16; i++) V[i] = sV[i];
		const grad::Vec3<float> mean = { a.means3D[3 * idx], a.means3D[3 * idx + 1], a.means3D[3 * idx + 2] };
		float cov6[6];
#pragma unroll
		for (int i = 0; i < 6; i++) cov6[i] = a.cov3D[6 * (size_t)idx + i];

		// covariance / conic branch and screen-position branch (backward.cu:297-485 + :642-660; pinhole :156-292 + :583-597)
		// in our own chain-rule form: gaussian_grad.cuh
		float Pm[16];
		if constexpr (kPinhole) {
#pragma unroll
			for (int i = 0; i < 16; i++) Pm[i] = sP[i];
		}
		const grad::Vec3<float> dmp = grad::projection_backward<float, kPinhole>(
			mean, cov6, V, Pm, a.W, a.H, a.focal_x, a.focal_y, a.tan_fovx, a.tan_fovy, g[0], g[1], g[2], g[3], g[4], dcov6);
		dmean = { dmp.x, dmp.y, dmp.z };

		// colour branch (backward.cu:30-151)
		if (kRaw || a.shs != nullptr) {
			const unsigned cm = a.clamped[idx];
			drgb[0] = (cm & 1u) ? 0.f : g[6];
			drgb[1] = (cm & 2u) ? 0.f : g[7];
			drgb[2] = (cm & 4u) ? 0.f : g[8];
			const grad::Vec3<float> cam = { sCam[0], sCam[1], sCam[2] };
			grad::Vec3<float> dm3;
			if constexpr (kBulkSH) {
				mbar_wait(&s_bar, 0);
				sh_waited = true;
				// one register row: colour_backward reads coefficient k before it writes gradient k, so dL/dsh overwrites the
				// coefficients in place (48 registers less than separate rows)
				float shr[kShRowFloats];
				if constexpr (kMode == 2) {
#pragma unroll
					for (int k = 0; k < 3; k++) shr[k] = s_sh[kRawDcOffset + tid * 3 + k];
#pragma unroll
					for (int k = 0; k < kRawRestFloats; k++) shr[3 + k] = s_sh[tid * kRawRestFloats + k];
				} else {
					const float4* row = reinterpret_cast<const float4*>(&s_sh[tid * kShPitchFloats]);
#pragma unroll
					for (int k = 0; k < kShRowFloats / 4; k++) {
						const float4 q = row[k];
						shr[4 * k] = q.x; shr[4 * k + 1] = q.y; shr[4 * k + 2] = q.z; shr[4 * k + 3] = q.w;
					}
				}
				auto sh = [&shr](int k, int c) { return shr[3 * k + c]; };
				auto dsh = [&shr](int k, int c, float v) { shr[3 * k + c] = v; };
				dm3 = grad::colour_backward<float>(a.D, mean, cam, sh, drgb, dsh);
				if (write_sh) {
					const int n3 = 3 * (a.D + 1) * (a.D + 1);
#pragma unroll
					for (int k = 0; k < kShRowFloats; k++)
						if (k >= n3) shr[k] = 0.f;   // rows beyond (D+1)^2 get zero gradient
					if constexpr (kMode == 2) {
#pragma unroll
						for (int k = 0; k < 3; k++) s_sh[kRawDcOffset + tid * 3 + k] = shr[k];
#pragma unroll
						for (int k = 0; k < kRawRestFloats; k++) s_sh[tid * kRawRestFloats + k] = shr[3 + k];
					} else {
						float4* row = reinterpret_cast<float4*>(&s_sh[tid * kShPitchFloats]);
#pragma unroll
						for (int k = 0; k < kShRowFloats / 4; k++)
							row[k] = make_float4(shr[4 * k], shr[4 * k + 1], shr[4 * k + 2], shr[4 * k + 3]);
					}
					sh_row_ready = true;
				}
			} else if constexpr (kMode == 3) {
				const float* dc = a.features_dc + (size_t)idx * 3;
				const float* rest = a.features_rest + (size_t)idx * (a.M - 1) * 3;
				float* ddc = a.dL_dfeatures_dc + (size_t)idx * 3;
				float* drest = a.dL_dfeatures_rest + (size_t)idx * (a.M - 1) * 3;
				auto sh = [dc, rest](int k, int c) { return k == 0 ? dc[c] : rest[3 * (k - 1) + c]; };
				auto dsh = [ddc, drest, write_sh](int k, int c, float v) {
					if (write_sh) { if (k == 0) ddc[c] = v; else drest[3 * (k - 1) + c] = v; }
				};
				dm3 = grad::colour_backward<float>(a.D, mean, cam, sh, drgb, dsh);
				if (write_sh) {
					for (int k = (a.D + 1) * (a.D + 1); k < a.M; k++)
						for (int c = 0; c < 3; c++) dsh(k, c, 0.f);
					sh_row_ready = true;
				}
			} else {
				const float* shp = a.shs + (size_t)idx * a.M * 3;
				auto sh = [shp](int k, int c) { return shp[3 * k + c]; };
				auto dsh = [dsh_row](int k, int c, float v) { if (dsh_row) dsh_row[3 * k + c] = v; };
				dm3 = grad::colour_backward<float>(a.D, mean, cam, sh, drgb, dsh);
				if (dsh_row)
					for (int k = (a.D + 1) * (a.D + 1) * 3; k < a.M * 3; k++) dsh_row[k] = 0.f;
			}
			dmean.x += dm3.x; dmean.y += dm3.y; dmean.z += dm3.z;
		}

		// scale / rotation branch (backward.cu:489-552)
		if (a.scales != nullptr) {
			float sc[3] = { a.scales[3 * idx], a.scales[3 * idx + 1], a.scales[3 * idx + 2] };
			const float4 q4 = reinterpret_cast<const float4*>(a.rotations)[idx];
			float q[4] = { q4.x, q4.y, q4.z, q4.w };
			if constexpr (kRaw) {
				// through exp (gaussian_model.cpp:54-57) and normalize (:59-62):
				// dL/ds_raw = dL/ds * exp(s_raw);  dL/dq_raw = (dL/dqn - qn <qn, dL/dqn>) / ||q_raw||
				const float act[3] = { expf(sc[0]), expf(sc[1]), expf(sc[2]) };
				const float n = quat_norm_clamped(q4);
				const float qn[4] = { q[0] / n, q[1] / n, q[2] / n, q[3] / n };
#pragma unroll
				for (int k = 0; k < 3; k++) sc[k] = a.scale_modifier * act[k];
				grad::scale_rotation_grad<float>(sc, qn, dcov6, dscale, drot);
				const float d = qn[0] * drot[0] + qn[1] * drot[1] + qn[2] * drot[2] + qn[3] * drot[3];
#pragma unroll
				for (int k = 0; k < 3; k++) dscale[k] *= act[k];
#pragma unroll
				for (int k = 0; k < 4; k++) drot[k] = (drot[k] - qn[k] * d) / n;
			} else {
#pragma unroll
				for (int k = 0; k < 3; k++) sc[k] *= a.scale_modifier;
				grad::scale_rotation_grad<float>(sc, q, dcov6, dscale, drot);
			}
		}
	} else if (dsh_row && !kBulkSH) {
		for (int k = 0; k < a.M * 3; k++) dsh_row[k] = 0.f;
	}
	if constexpr (kMode == 3) {
		if (write_sh && !sh_row_ready) {
			for (int k = 0; k < 3; k++) a.dL_dfeatures_dc[(size_t)idx * 3 + k] = 0.f;
			for (int k = 0; k < (a.M - 1) * 3; k++) a.dL_dfeatures_rest[(size_t)idx * (a.M - 1) * 3 + k] = 0.f;
		}
	}
	if constexpr (kMode == 2) {
		// gradient rows (zeros for culled Gaussians) leave as two bulk stores per CTA; <= 3 tail rows plainly
		if (!sh_waited) mbar_wait(&s_bar, 0);
		if (write_sh) {
			if (!sh_row_ready) {
				for (int k = 0; k < 3; k++) s_sh[kRawDcOffset + tid * 3 + k] = 0.f;
				for (int k = 0; k < kRawRestFloats; k++) s_sh[tid * kRawRestFloats + k] = 0.f;
			}
			if (tid >= rows4) {
				for (int k = 0; k < kRawRestFloats; k++) a.dL_dfeatures_rest[(size_t)idx * kRawRestFloats + k] = s_sh[tid * kRawRestFloats + k];
				for (int k = 0; k < 3; k++) a.dL_dfeatures_dc[(size_t)idx * 3 + k] = s_sh[kRawDcOffset + tid * 3 + k];
			}
			// no block-wide barrier here: threads beyond P have left, and a partly active warp must not meet
			// bar.sync divergently.  Every active thread arrives on s_rows_done; thread 0 waits for all of them.
			fence_async_smem();
			mbar_arrive(&s_rows_done);
			if (tid == 0 && rows4) {
				mbar_wait(&s_rows_done, 0);
				const size_t first = (size_t)blk * kPreBwdThreads;
				bulk_store(a.dL_dfeatures_rest + first * kRawRestFloats, &s_sh[0], (uint32_t)rows4 * kRawRestFloats * 4u);
				bulk_store(a.dL_dfeatures_dc + first * 3, &s_sh[kRawDcOffset], (uint32_t)rows4 * 12u);
				bulk_commit();
			}
		}
	}
	if constexpr (kMode == 1) {
		// hand the gradient row (zeros for culled Gaussians) to the copy engine
		if (!sh_waited) mbar_wait(&s_bar, 0);
		if (write_sh) {
			float4* row = reinterpret_cast<float4*>(&s_sh[tid * kShPitchFloats]);
			if (!sh_row_ready) {
#pragma unroll
				for (int k = 0; k < kShRowFloats / 4; k++) row[k] = make_float4(0.f, 0.f, 0.f, 0.f);
			}
			fence_async_smem();
			bulk_store(dsh_row, row, kShRowFloats * 4u);
			bulk_commit();
		}
	}
	if (a.dL_drgb_view) {
		a.dL_drgb_view[3 * (size_t)idx + 0] = drgb[0];
		a.dL_drgb_view[3 * (size_t)idx + 1] = drgb[1];
		a.dL_drgb_view[3 * (size_t)idx + 2] = drgb[2];
	}

	float* m3 = a.dL_dmean3D + 3 * (size_t)idx;
	float* ds = a.dL_dscale + 3 * (size_t)idx;
	float4* dr = reinterpret_cast<float4*>(a.dL_drot) + idx;
	if (a.accumulate) {
		m3[0] += dmean.x; m3[1] += dmean.y; m3[2] += dmean.z;
		a.dL_dopacity[idx] += dopacity;
		ds[0] += dscale[0]; ds[1] += dscale[1]; ds[2] += dscale[2];
		const float4 o = *dr;
		*dr = make_float4(o.x + drot[0], o.y + drot[1], o.z + drot[2], o.w + drot[3]);
	} else {
		m3[0] = dmean.x; m3[1] = dmean.y; m3[2] = dmean.z;
		a.dL_dopacity[idx] = dopacity;
		ds[0] = dscale[0]; ds[1] = dscale[1]; ds[2] = dscale[2];
		*dr = make_float4(drot[0], drot[1], drot[2], drot[3]);
	}
	if (a.dL_dcov3D) {
#pragma unroll
		for (int i = 0; i < 6; i++) a.dL_dcov3D[6 * (size_t)idx + i] = dcov6[i];
	}
	if (kBulkSH) bulk_wait_read_all();   // shared memory must outlive the outgoing copy
}

// dL/dsh of a whole training step from the per-view dL/dRGB factors: row k of a Gaussian's gradient is rank one per view,
// b_k(dir_view) * dL/dRGB_view, so ranks exchange 12 bytes per Gaussian and view instead of 192 and every rank rebuilds the
// same sum here (views in a fixed order: all replicas get identical bits, equal to adding up the per-view dL/dsh tensors
// in that order).  drgb[v] may be a peer GPU's buffer (NVLink loads: the transfer IS this kernel's input stream).
// One warp per 32 Gaussians: the 48-float rows are staged in shared memory and leave as fully coalesced 128-bit stores.
constexpr int kShViewsThreads = 128;
constexpr int kShViewsBatch = 4;   // views whose factors are loaded together (independent NVLink loads in flight per thread)
__global__ void __launch_bounds__(kShViewsThreads) sh_gradient_from_views_kernel(const ShFromViewsArgs a)
{
	__shared__ float s_rows[kShViewsThreads / 32][kShRowFloats * 33];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int n = (a.D + 1) * (a.D + 1);
	float* s = s_rows[warp];
	// grid-stride over groups of 32 Gaussians
	const int groups = (a.P + 31) / 32;
	for (int grp = blockIdx.x * (kShViewsThreads / 32) + warp; grp < groups; grp += gridDim.x * (kShViewsThreads / 32)) {
		const int first = grp * 32;
		const int idx = first + lane;
		float acc[kShRowFloats];
#pragma unroll
		for (int k = 0; k < kShRowFloats; k++) acc[k] = 0.f;
		if (idx < a.P) {
			const float mx = a.means3D[3 * (size_t)idx], my = a.means3D[3 * (size_t)idx + 1], mz = a.means3D[3 * (size_t)idx + 2];
			for (int v0 = 0; v0 < a.n_views; v0 += kShViewsBatch) {
				float dd[kShViewsBatch][3];
#pragma unroll
				for (int j = 0; j < kShViewsBatch; j++) {
					const bool have = v0 + j < a.n_views;
					const float* dv = a.drgb[have ? v0 + j : 0] + 3 * (size_t)idx;
					dd[j][0] = have ? dv[0] : 0.f;
					dd[j][1] = have ? dv[1] : 0.f;
					dd[j][2] = have ? dv[2] : 0.f;
				}
#pragma unroll
				for (int j = 0; j < kShViewsBatch; j++) {
					const float d0 = dd[j][0], d1 = dd[j][1], d2 = dd[j][2];
					if (d0 == 0.f && d1 == 0.f && d2 == 0.f) continue;   // not visible in this view (or fully clamped): adds exact zeros
					const int v = v0 + j;
					const float vx = mx - a.campos[3 * v], vy = my - a.campos[3 * v + 1], vz = mz - a.campos[3 * v + 2];
					const float len = sqrtf(vx * vx + vy * vy + vz * vz);
					float b[16];
					grad::sh_weights<float>(a.D, vx / len, vy / len, vz / len, b);
#pragma unroll
					for (int k = 0; k < 16; k++) {
						if (k < n) {
							acc[3 * k + 0] = __fadd_rn(acc[3 * k + 0], grad::mul1(b[k], d0));
							acc[3 * k + 1] = __fadd_rn(acc[3 * k + 1], grad::mul1(b[k], d1));
							acc[3 * k + 2] = __fadd_rn(acc[3 * k + 2], grad::mul1(b[k], d2));
						}
					}
				}
			}
		}
		// transposed staging at a 33-word pitch: element (row r, float k) at s[k * 33 + r] — lane-per-row writes and
		// lane-per-output-float reads are both conflict-free
		const int rows = min(32, a.P - first);
#pragma unroll
		for (int k = 0; k < kShRowFloats; k++) s[k * 33 + lane] = acc[k];
		__syncwarp();
		if (a.dL_dsh) {
			// [P,16,3]: the warp's rows are rows * 48 contiguous floats, stored as 128-bit words
			float4* out = reinterpret_cast<float4*>(a.dL_dsh + (size_t)first * kShRowFloats);
			for (int i = lane; i < rows * (kShRowFloats / 4); i += 32) {
				const int f = 4 * i, r = f / kShRowFloats, k = f - r * kShRowFloats;   // 48 is a multiple of 4: one row per word
				out[i] = make_float4(s[k * 33 + r], s[(k + 1) * 33 + r], s[(k + 2) * 33 + r], s[(k + 3) * 33 + r]);
			}
		} else {
			// split layout (raw-parameter trainer): dL/dfeatures_rest rows of 45 floats, dL/dfeatures_dc rows of 3
			float* rest = a.dL_dfeatures_rest + (size_t)first * kRawRestFloats;
			float* dc = a.dL_dfeatures_dc + (size_t)first * 3;
			for (int i = lane; i < rows * kRawRestFloats; i += 32) {
				const int r = i / kRawRestFloats, k = i - r * kRawRestFloats;
				rest[i] = s[(3 + k) * 33 + r];
			}
			for (int i = lane; i < rows * 3; i += 32) dc[i] = s[(i % 3) * 33 + i / 3];
		}
		__syncwarp();   // the staging rows are reused by the next group
	}
}

int launch_sh_gradient_from_views(const ShFromViewsArgs& a, cudaStream_t st)
{
	if (a.P <= 0) return OGS_OK;
	const int per_block = kShViewsThreads;   // Gaussians per block and trip
	// one trip per block: a capped, slower grid was measured on four B200s (profiles/r02_dp_rebuild_sweep_n4.log) — it
	// disturbs the latency-bound sort passes it overlaps for longer and the step gets slower
	const int blocks = ceil_div(a.P, per_block);
	sh_gradient_from_views_kernel<<<blocks, kShViewsThreads, 0, st>>>(a);
	OGS_CUDA_TRY(cudaGetLastError());
	return OGS_OK;
}

int launch_preprocess_bwd(const PreprocessBwdArgs& a, cudaStream_t st)
{
	const int all_blocks = ceil_div(a.P, kPreBwdThreads);
	if (a.first_block < 0 || a.first_block > all_blocks || a.num_blocks < 0) return fail(OGS_ERR_INVALID_ARG, "bad Gaussian block range");
	const int blocks = a.num_blocks > 0 ? min(a.num_blocks, all_blocks - a.first_block) : all_blocks - a.first_block;
	if (blocks <= 0) return OGS_OK;
	if (a.pinhole) {
		if (a.shs != nullptr && a.dL_dsh != nullptr && sh_rows_bulk_capable(a.shs, a.M) && sh_rows_bulk_capable(a.dL_dsh, a.M))
			preprocess_lonlat_bwd_kernel<1, true><<<blocks, kPreBwdThreads, 0, st>>>(a);
		else
			preprocess_lonlat_bwd_kernel<0, true><<<blocks, kPreBwdThreads, 0, st>>>(a);
	} else if (a.raw) {
		if (sh_rows_bulk_capable(a.features_rest, a.M) && sh_rows_bulk_capable(a.features_dc, a.M) &&
		    sh_rows_bulk_capable(a.dL_dfeatures_rest, a.M) && sh_rows_bulk_capable(a.dL_dfeatures_dc, a.M))
			preprocess_lonlat_bwd_kernel<2><<<blocks, kPreBwdThreads, 0, st>>>(a);
		else
			preprocess_lonlat_bwd_kernel<3><<<blocks, kPreBwdThreads, 0, st>>>(a);
	} else if (a.shs != nullptr && sh_rows_bulk_capable(a.shs, a.M) &&
	           ((a.dL_dsh != nullptr && sh_rows_bulk_capable(a.dL_dsh, a.M)) || (a.dL_dsh == nullptr && a.dL_drgb_view != nullptr)))
		// (multi-view mode leaves dL/dRGB instead of the dL/dsh row: the SH rows still come in through the bulk path)
		preprocess_lonlat_bwd_kernel<1><<<blocks, kPreBwdThreads, 0, st>>>(a);
	else
		preprocess_lonlat_bwd_kernel<0><<<blocks, kPreBwdThreads, 0, st>>>(a);
	OGS_CUDA_TRY(cudaGetLastError());
	return OGS_OK;
}

} // namespace ogs
