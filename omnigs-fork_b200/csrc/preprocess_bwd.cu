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
#include "pinhole_math.cuh"
#include "launchers.cuh"
#include "async_copy.cuh"

namespace ogs {

constexpr int kPreBwdThreads = 128;

// kMode as in preprocess_fwd.cu: 0 plain SH rows, 1 bulk SH rows, 2 raw parameters (bulk), 3 raw parameters (plain)
// kPinhole: the perspective camera's covariance and screen-position branches (backward.cu:156-292, :583-597).
template <int kMode, bool kPinhole = false>
__global__ void __launch_bounds__(kPreBwdThreads) preprocess_lonlat_bwd_kernel(const PreprocessBwdArgs a)
{
	constexpr bool kBulkSH = (kMode == 1 || kMode == 2);
	constexpr bool kRaw = (kMode >= 2);
	__shared__ float sV[16];
	__shared__ float sCam[3];
	__shared__ __align__(16) float s_sh[kBulkSH ? kPreBwdThreads * kShPitchFloats : 4];
	__shared__ __align__(8) uint64_t s_bar;
	__shared__ __align__(8) uint64_t s_rows_done;   // raw mode: counts the CTA's gradient rows written to shared memory
	const int tid = threadIdx.x;
	const int idx = blockIdx.x * kPreBwdThreads + tid;
	if (tid < 16) sV[tid] = a.viewmatrix[tid];
	if (tid < 3) sCam[tid] = a.campos[tid];
	const int rows = min(kPreBwdThreads, a.P - (int)blockIdx.x * kPreBwdThreads);
	const int rows4 = rows & ~3;   // raw mode: rows covered by the per-CTA bulk copies (sizes stay multiples of 16 B)
	if (kBulkSH && tid == 0) {
		mbar_init(&s_bar, 1);
		if constexpr (kMode == 2) {
			mbar_init(&s_rows_done, (uint32_t)rows);
			mbar_arrive_expect_tx(&s_bar, (uint32_t)rows4 * kShRowFloats * 4u);
			if (rows4) {
				const size_t first = (size_t)blockIdx.x * kPreBwdThreads;
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

	const bool visible = a.radii[idx] > 0;

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

	// the render-backward outputs in the reference's layouts
	if (!kRaw || a.dL_dmean2D) {
		a.dL_dmean2D[3 * (size_t)idx + 0] = g[0];
		a.dL_dmean2D[3 * (size_t)idx + 1] = g[1];
		a.dL_dmean2D[3 * (size_t)idx + 2] = 0.f;
	}
	if (a.dL_dconic) reinterpret_cast<float4*>(a.dL_dconic)[idx] = make_float4(g[2], g[3], 0.f, g[4]);
	if constexpr (kRaw) {
		// d sigmoid: the activated opacity is in the packed record (gaussian_model.cpp:74-77)
		const float o = visible ? a.g1[idx].y : 0.f;
		a.dL_dopacity[idx] = g[5] * o * (1.f - o);
	} else {
		a.dL_dopacity[idx] = g[5];
		a.dL_dcolor[3 * (size_t)idx + 0] = g[6];
		a.dL_dcolor[3 * (size_t)idx + 1] = g[7];
		a.dL_dcolor[3 * (size_t)idx + 2] = g[8];
	}

	float dcov6[6] = { 0.f, 0.f, 0.f, 0.f, 0.f, 0.f };
	float3 dmean = { 0.f, 0.f, 0.f };
	float3 dscale = { 0.f, 0.f, 0.f };
	float4 drot = { 0.f, 0.f, 0.f, 0.f };
	float* dsh_row = (!kRaw && a.dL_dsh) ? a.dL_dsh + (size_t)idx * a.M * 3 : nullptr;
	bool sh_waited = false, sh_row_ready = false;

	if (visible) {
		float V[16];
#pragma unroll
		for (int i = 0; i < 16; i++) V[i] = sV[i];
		const float3 mean = { a.means3D[3 * idx], a.means3D[3 * idx + 1], a.means3D[3 * idx + 2] };
		float cov6[6];
#pragma unroll
		for (int i = 0; i < 6; i++) cov6[i] = a.cov3D[6 * (size_t)idx + i];

		if constexpr (kPinhole) {
			// covariance / conic branch (backward.cu:156-292), then the screen position through the full
			// projection (backward.cu:583-597)
			cov2d_pinhole_backward(mean, cov6, V, a.focal_x, a.focal_y, a.tan_fovx, a.tan_fovy,
			                       float3{ g[2], g[3], g[4] }, dcov6, dmean);
			const float3 dm2 = proj_point_backward(a.projmatrix, mean, g[0], g[1]);
			dmean.x += dm2.x; dmean.y += dm2.y; dmean.z += dm2.z;
		} else {
			// covariance / conic branch (backward.cu:297-485); assigns dL/dmean
			float3 dpx_dt, dpy_dt;
			cov2d_lonlat_backward(mean, cov6, V, a.W, a.H, float3{ g[2], g[3], g[4] }, dcov6, dmean, dpx_dt, dpy_dt);

			// screen-position branch (backward.cu:642-660)
			const float dsx_dpx = 2.0f / (float)a.W;
			const float dsy_dpy = 2.0f / (float)a.H;
			const float dL_dpx = g[0] * dsx_dpx;
			const float dL_dpy = g[1] * dsy_dpy;
			const float dL_dtx = dL_dpx * dpx_dt.x + dL_dpy * dpy_dt.x;
			const float dL_dty = dL_dpx * dpx_dt.y + dL_dpy * dpy_dt.y;
			const float dL_dtz = dL_dpx * dpx_dt.z + dL_dpy * dpy_dt.z;
			const float3 dm2 = view_vec_t(V, float3{ dL_dtx, dL_dty, dL_dtz });
			dmean.x += dm2.x; dmean.y += dm2.y; dmean.z += dm2.z;
		}

		// SH backward (backward.cu:30-151)
		if (kRaw || a.shs != nullptr) {
			const unsigned cm = a.clamped[idx];
			V3 dRGB = { g[6], g[7], g[8] };
			dRGB.x *= (cm & 1u) ? 0 : 1;
			dRGB.y *= (cm & 2u) ? 0 : 1;
			dRGB.z *= (cm & 4u) ? 0 : 1;
			float3 dm3;
			if constexpr (kMode == 2) {
				mbar_wait(&s_bar, 0);
				sh_waited = true;
				float shr[kShRowFloats], dshr[kShRowFloats];
#pragma unroll
				for (int k = 0; k < 3; k++) shr[k] = s_sh[kRawDcOffset + tid * 3 + k];
#pragma unroll
				for (int k = 0; k < kRawRestFloats; k++) shr[3 + k] = s_sh[tid * kRawRestFloats + k];
#pragma unroll
				for (int k = 0; k < kShRowFloats; k++) dshr[k] = 0.f;
				auto sh = [&shr](int k) { return V3{ shr[3 * k], shr[3 * k + 1], shr[3 * k + 2] }; };
				auto dsh = [&dshr](int k, V3 v) { dshr[3 * k] = v.x; dshr[3 * k + 1] = v.y; dshr[3 * k + 2] = v.z; };
				dm3 = sh_backward(a.D, mean, float3{ sCam[0], sCam[1], sCam[2] }, sh, dRGB, dsh);
#pragma unroll
				for (int k = 0; k < 3; k++) s_sh[kRawDcOffset + tid * 3 + k] = dshr[k];
#pragma unroll
				for (int k = 0; k < kRawRestFloats; k++) s_sh[tid * kRawRestFloats + k] = dshr[3 + k];
				sh_row_ready = true;
			} else if constexpr (kMode == 3) {
				const float* dc = a.features_dc + (size_t)idx * 3;
				const float* rest = a.features_rest + (size_t)idx * (a.M - 1) * 3;
				float* ddc = a.dL_dfeatures_dc + (size_t)idx * 3;
				float* drest = a.dL_dfeatures_rest + (size_t)idx * (a.M - 1) * 3;
				auto sh = [dc, rest](int k) {
					const float* p = k == 0 ? dc : rest + 3 * (k - 1);
					return V3{ p[0], p[1], p[2] };
				};
				auto dsh = [ddc, drest](int k, V3 v) {
					float* p = k == 0 ? ddc : drest + 3 * (k - 1);
					p[0] = v.x; p[1] = v.y; p[2] = v.z;
				};
				dm3 = sh_backward(a.D, mean, float3{ sCam[0], sCam[1], sCam[2] }, sh, dRGB, dsh);
				for (int k = (a.D + 1) * (a.D + 1); k < a.M; k++) dsh(k, V3{ 0.f, 0.f, 0.f });
				sh_row_ready = true;
			} else if constexpr (kBulkSH) {
				mbar_wait(&s_bar, 0);
				sh_waited = true;
				float shr[kShRowFloats], dshr[kShRowFloats];
				float4* row = reinterpret_cast<float4*>(&s_sh[tid * kShPitchFloats]);
#pragma unroll
				for (int k = 0; k < kShRowFloats / 4; k++) {
					const float4 q = row[k];
					shr[4 * k] = q.x; shr[4 * k + 1] = q.y; shr[4 * k + 2] = q.z; shr[4 * k + 3] = q.w;
				}
#pragma unroll
				for (int k = 0; k < kShRowFloats; k++) dshr[k] = 0.f;   // rows beyond (D+1)^2 stay zero
				auto sh = [&shr](int k) { return V3{ shr[3 * k], shr[3 * k + 1], shr[3 * k + 2] }; };
				auto dsh = [&dshr](int k, V3 v) { dshr[3 * k] = v.x; dshr[3 * k + 1] = v.y; dshr[3 * k + 2] = v.z; };
				dm3 = sh_backward(a.D, mean, float3{ sCam[0], sCam[1], sCam[2] }, sh, dRGB, dsh);
#pragma unroll
				for (int k = 0; k < kShRowFloats / 4; k++)
					row[k] = make_float4(dshr[4 * k], dshr[4 * k + 1], dshr[4 * k + 2], dshr[4 * k + 3]);
				sh_row_ready = true;
			} else {
				const float* shp = a.shs + (size_t)idx * a.M * 3;
				auto sh = [shp](int k) { return V3{ shp[3 * k], shp[3 * k + 1], shp[3 * k + 2] }; };
				auto dsh = [dsh_row](int k, V3 v) {
					dsh_row[3 * k] = v.x; dsh_row[3 * k + 1] = v.y; dsh_row[3 * k + 2] = v.z;
				};
				dm3 = sh_backward(a.D, mean, float3{ sCam[0], sCam[1], sCam[2] }, sh, dRGB, dsh);
				for (int k = (a.D + 1) * (a.D + 1) * 3; k < a.M * 3; k++) dsh_row[k] = 0.f;
			}
			dmean.x += dm3.x; dmean.y += dm3.y; dmean.z += dm3.z;
		}

		// scale / rotation backward (backward.cu:489-552)
		if (a.scales != nullptr) {
			float3 sc = { a.scales[3 * idx], a.scales[3 * idx + 1], a.scales[3 * idx + 2] };
			float4 q = reinterpret_cast<const float4*>(a.rotations)[idx];
			if constexpr (kRaw) {
				// through exp (gaussian_model.cpp:54-57) and normalize (:59-62):
				// dL/ds_raw = dL/ds * exp(s_raw);  dL/dq_raw = (dL/dqn - qn <qn, dL/dqn>) / ||q_raw||
				sc = { expf(sc.x), expf(sc.y), expf(sc.z) };
				const float n = quat_norm_clamped(q);
				const float4 qn = make_float4(q.x / n, q.y / n, q.z / n, q.w / n);
				cov3d_backward(sc, a.scale_modifier, qn, dcov6, dscale, drot);
				dscale = { dscale.x * sc.x, dscale.y * sc.y, dscale.z * sc.z };
				const float d = qn.x * drot.x + qn.y * drot.y + qn.z * drot.z + qn.w * drot.w;
				drot = make_float4((drot.x - qn.x * d) / n, (drot.y - qn.y * d) / n, (drot.z - qn.z * d) / n, (drot.w - qn.w * d) / n);
			} else {
				cov3d_backward(sc, a.scale_modifier, q, dcov6, dscale, drot);
			}
		}
	} else if (dsh_row && !kBulkSH) {
		for (int k = 0; k < a.M * 3; k++) dsh_row[k] = 0.f;
	}
	if constexpr (kMode == 3) {
		if (!sh_row_ready) {
			for (int k = 0; k < 3; k++) a.dL_dfeatures_dc[(size_t)idx * 3 + k] = 0.f;
			for (int k = 0; k < (a.M - 1) * 3; k++) a.dL_dfeatures_rest[(size_t)idx * (a.M - 1) * 3 + k] = 0.f;
		}
	}
	if constexpr (kMode == 2) {
		// gradient rows (zeros for culled Gaussians) leave as two bulk stores per CTA; <= 3 tail rows plainly
		if (!sh_waited) mbar_wait(&s_bar, 0);
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
			const size_t first = (size_t)blockIdx.x * kPreBwdThreads;
			bulk_store(a.dL_dfeatures_rest + first * kRawRestFloats, &s_sh[0], (uint32_t)rows4 * kRawRestFloats * 4u);
			bulk_store(a.dL_dfeatures_dc + first * 3, &s_sh[kRawDcOffset], (uint32_t)rows4 * 12u);
			bulk_commit();
		}
	}
	if constexpr (kMode == 1) {
		// hand the gradient row (zeros for culled Gaussians) to the copy engine
		if (!sh_waited) mbar_wait(&s_bar, 0);
		float4* row = reinterpret_cast<float4*>(&s_sh[tid * kShPitchFloats]);
		if (!sh_row_ready) {
#pragma unroll
			for (int k = 0; k < kShRowFloats / 4; k++) row[k] = make_float4(0.f, 0.f, 0.f, 0.f);
		}
		fence_async_smem();
		bulk_store(dsh_row, row, kShRowFloats * 4u);
		bulk_commit();
	}

	a.dL_dmean3D[3 * (size_t)idx + 0] = dmean.x;
	a.dL_dmean3D[3 * (size_t)idx + 1] = dmean.y;
	a.dL_dmean3D[3 * (size_t)idx + 2] = dmean.z;
	if (!kRaw || a.dL_dcov3D) {
#pragma unroll
		for (int i = 0; i < 6; i++) a.dL_dcov3D[6 * (size_t)idx + i] = dcov6[i];
	}
	a.dL_dscale[3 * (size_t)idx + 0] = dscale.x;
	a.dL_dscale[3 * (size_t)idx + 1] = dscale.y;
	a.dL_dscale[3 * (size_t)idx + 2] = dscale.z;
	reinterpret_cast<float4*>(a.dL_drot)[idx] = drot;
	if (kBulkSH) bulk_wait_read_all();   // shared memory must outlive the outgoing copy
}

int launch_preprocess_bwd(const PreprocessBwdArgs& a, cudaStream_t st)
{
	const int blocks = ceil_div(a.P, kPreBwdThreads);
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
	} else if (a.shs != nullptr && a.dL_dsh != nullptr && sh_rows_bulk_capable(a.shs, a.M) && sh_rows_bulk_capable(a.dL_dsh, a.M))
		preprocess_lonlat_bwd_kernel<1><<<blocks, kPreBwdThreads, 0, st>>>(a);
	else
		preprocess_lonlat_bwd_kernel<0><<<blocks, kPreBwdThreads, 0, st>>>(a);
	OGS_CUDA_TRY(cudaGetLastError());
	return OGS_OK;
}

} // namespace ogs
