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
#include "lonlat_math.cuh"
#include "launchers.cuh"
#include "async_copy.cuh"

namespace ogs {

constexpr int kPreBwdThreads = 128;

template <bool kBulkSH>
__global__ void __launch_bounds__(kPreBwdThreads) preprocess_lonlat_bwd_kernel(const PreprocessBwdArgs a)
{
	__shared__ float sV[16];
	__shared__ float sCam[3];
	__shared__ __align__(16) float s_sh[kBulkSH ? kPreBwdThreads * kShPitchFloats : 4];
	__shared__ __align__(8) uint64_t s_bar;
	const int tid = threadIdx.x;
	const int idx = blockIdx.x * kPreBwdThreads + tid;
	if (tid < 16) sV[tid] = a.viewmatrix[tid];
	if (tid < 3) sCam[tid] = a.campos[tid];
	if (kBulkSH && tid == 0) {
		const int rows = min(kPreBwdThreads, a.P - (int)blockIdx.x * kPreBwdThreads);
		mbar_init(&s_bar, 1);
		mbar_arrive_expect_tx(&s_bar, (uint32_t)rows * kShRowFloats * 4u);
	}
	__syncthreads();
	if (kBulkSH && idx < a.P)
		bulk_load(&s_sh[tid * kShPitchFloats], a.shs + (size_t)idx * kShRowFloats, kShRowFloats * 4u, &s_bar);
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
	a.dL_dmean2D[3 * (size_t)idx + 0] = g[0];
	a.dL_dmean2D[3 * (size_t)idx + 1] = g[1];
	a.dL_dmean2D[3 * (size_t)idx + 2] = 0.f;
	if (a.dL_dconic) reinterpret_cast<float4*>(a.dL_dconic)[idx] = make_float4(g[2], g[3], 0.f, g[4]);
	a.dL_dopacity[idx] = g[5];
	a.dL_dcolor[3 * (size_t)idx + 0] = g[6];
	a.dL_dcolor[3 * (size_t)idx + 1] = g[7];
	a.dL_dcolor[3 * (size_t)idx + 2] = g[8];

	float dcov6[6] = { 0.f, 0.f, 0.f, 0.f, 0.f, 0.f };
	float3 dmean = { 0.f, 0.f, 0.f };
	float3 dscale = { 0.f, 0.f, 0.f };
	float4 drot = { 0.f, 0.f, 0.f, 0.f };
	float* dsh_row = a.dL_dsh ? a.dL_dsh + (size_t)idx * a.M * 3 : nullptr;
	bool sh_waited = false, sh_row_ready = false;

	if (visible) {
		float V[16];
#pragma unroll
		for (int i = 0; i < 16; i++) V[i] = sV[i];
		const float3 mean = { a.means3D[3 * idx], a.means3D[3 * idx + 1], a.means3D[3 * idx + 2] };
		float cov6[6];
#pragma unroll
		for (int i = 0; i < 6; i++) cov6[i] = a.cov3D[6 * (size_t)idx + i];

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

		// SH backward (backward.cu:30-151)
		if (a.shs != nullptr) {
			const unsigned cm = a.clamped[idx];
			V3 dRGB = { g[6], g[7], g[8] };
			dRGB.x *= (cm & 1u) ? 0 : 1;
			dRGB.y *= (cm & 2u) ? 0 : 1;
			dRGB.z *= (cm & 4u) ? 0 : 1;
			float3 dm3;
			if (kBulkSH) {
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
			const float3 sc = { a.scales[3 * idx], a.scales[3 * idx + 1], a.scales[3 * idx + 2] };
			const float4 q = reinterpret_cast<const float4*>(a.rotations)[idx];
			cov3d_backward(sc, a.scale_modifier, q, dcov6, dscale, drot);
		}
	} else if (dsh_row && !kBulkSH) {
		for (int k = 0; k < a.M * 3; k++) dsh_row[k] = 0.f;
	}
	if (kBulkSH) {
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
#pragma unroll
	for (int i = 0; i < 6; i++) a.dL_dcov3D[6 * (size_t)idx + i] = dcov6[i];
	a.dL_dscale[3 * (size_t)idx + 0] = dscale.x;
	a.dL_dscale[3 * (size_t)idx + 1] = dscale.y;
	a.dL_dscale[3 * (size_t)idx + 2] = dscale.z;
	reinterpret_cast<float4*>(a.dL_drot)[idx] = drot;
	if (kBulkSH) bulk_wait_read_all();   // shared memory must outlive the outgoing copy
}

int launch_preprocess_bwd(const PreprocessBwdArgs& a, cudaStream_t st)
{
	const int blocks = ceil_div(a.P, kPreBwdThreads);
	if (a.shs != nullptr && a.dL_dsh != nullptr && sh_rows_bulk_capable(a.shs, a.M) && sh_rows_bulk_capable(a.dL_dsh, a.M))
		preprocess_lonlat_bwd_kernel<true><<<blocks, kPreBwdThreads, 0, st>>>(a);
	else
		preprocess_lonlat_bwd_kernel<false><<<blocks, kPreBwdThreads, 0, st>>>(a);
	OGS_CUDA_TRY(cudaGetLastError());
	return OGS_OK;
}

} // namespace ogs
