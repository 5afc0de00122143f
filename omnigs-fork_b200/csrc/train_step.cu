// train_step.cu — the per-iteration work either side of the rasterizer:
//
//   * photometric loss, forward and backward in two kernels:
//       loss = (1 - lambda) * mean|I - gt| + lambda * (1 - SSIM(I, gt))
//     with I = rendered * mask, optionally cropped at the bottom (gaussian_mapper.cpp:391-413).
//     SSIM is the reference's: 11x11 Gaussian window (sigma 1.5), zero padding, per channel, mean over all
//     elements (include/loss_utils.h:58-131, built there from five grouped conv2d calls and autograd).
//     Forward kernel: one CTA per 32x32 pixel block and channel stages the 42x42 halo of both images in shared
//     memory, runs the separable window (register-tiled, four outputs per thread) over the five moment maps, evaluates the SSIM map and its three
//     partial derivatives (w.r.t. mu1, sigma1^2, sigma12) and accumulates the two loss sums.
//     Backward kernel: the window is self-adjoint, so dL/dI is the same separable filter applied to the three
//     derivative maps:  dSSIM/dI(p) = (w * dm_dmu1)(p) + 2 I(p) (w * dm_dsigma1sq)(p) + gt(p) (w * dm_dsigma12)(p).
//   * Adam over the optimiser's parameter groups in ONE launch (torch::optim::Adam as configured at
//     gaussian_model.cpp:485-518: betas (0.9, 0.999), eps 1e-15, no weight decay, per-group learning rate);
//   * densification statistics (gaussian_mapper.cpp:427-434, gaussian_model.cpp:839-853).
#include "launchers.cuh"
#include <cmath>

namespace ogs {

constexpr int kLossTile = 32;                  // output pixels per CTA edge
constexpr int kLossThreads = 256;
constexpr int kWin = 11, kHalf = 5;
constexpr int kHalo = kLossTile + 2 * kHalf;   // 42
constexpr int kHaloPitch = 44;                 // 16-byte aligned rows; a 4-output group reads columns 4g .. 4g+15 <= 43
constexpr int kGroups = kLossTile / 4;         // 4 horizontally (H pass) / vertically (V pass) adjacent outputs per thread

struct LossArgs {
	int W, H, H_used;          // rows >= H_used are cropped away (skip_bottom_ratio)
	int mask_channels;         // 0: no mask, 1: [1,H,W], 3: [3,H,W]
	float lambda_dssim;
	float w[kWin];             // normalised 1-D window
	const float* rendered;     // [3,H,W]
	const float* gt;           // [3,H,W]
	const float* mask;
	float* dmaps;              // [3 maps][3,H,W] workspace: dm_dmu1, dm_dsigma1sq, dm_dsigma12
	double* sums;              // [2]: sum |I - gt|, sum SSIM map
	float* loss_out;           // [3]: loss, L1, SSIM
	float* dL_dpix;            // [3,H,W]
};

OGS_D float masked_pixel(const LossArgs& a, int ch, int x, int y)
{
	const size_t i = (size_t)y * a.W + x, HW = (size_t)a.H * a.W;
	float v = a.rendered[ch * HW + i];
	if (a.mask_channels == 1) v *= a.mask[i];
	else if (a.mask_channels == 3) v *= a.mask[ch * HW + i];
	return v;
}

// Both kernels run the separable 11-tap window register-tiled: a thread produces FOUR adjacent outputs per pass
// from 14 inputs (128-bit shared loads in the horizontal pass), so an output costs 3.5 shared loads per map instead
// of 11 (the first version was bound by shared-memory bandwidth: ncu L1/shared 88-91 %).
__global__ void __launch_bounds__(kLossThreads) ssim_l1_fwd_kernel(const LossArgs a)
{
	__shared__ __align__(16) float s_x[kHalo][kHaloPitch], s_y[kHalo][kHaloPitch];
	__shared__ __align__(16) float s_h[5][kHalo][kLossTile];
	__shared__ float s_red[2][kLossThreads / 32];
	const int tid = threadIdx.x;
	const int ch = blockIdx.z;
	const int x0 = blockIdx.x * kLossTile - kHalf, y0 = blockIdx.y * kLossTile - kHalf;
	const size_t HW = (size_t)a.H * a.W;

	for (int i = tid; i < kHalo * kHaloPitch; i += kLossThreads) {
		const int r = i / kHaloPitch, c = i % kHaloPitch;
		const int x = x0 + c, y = y0 + r;
		float vx = 0.f, vy = 0.f;   // zero padding outside the (cropped) image
		if (c < kHalo && x >= 0 && x < a.W && y >= 0 && y < a.H_used) {
			vx = masked_pixel(a, ch, x, y);
			vy = a.gt[ch * HW + (size_t)y * a.W + x];
		}
		s_x[r][c] = vx;
		s_y[r][c] = vy;
	}
	__syncthreads();
	// horizontal pass: 42 rows x 8 groups of 4 columns, five moment maps
	for (int item = tid; item < kHalo * kGroups; item += kLossThreads) {
		const int r = item / kGroups, g = item % kGroups;
		float vx[16], vy[16];
#pragma unroll
		for (int q = 0; q < 4; q++) {
			const float4 fx = *reinterpret_cast<const float4*>(&s_x[r][4 * g + 4 * q]);
			const float4 fy = *reinterpret_cast<const float4*>(&s_y[r][4 * g + 4 * q]);
			vx[4 * q] = fx.x; vx[4 * q + 1] = fx.y; vx[4 * q + 2] = fx.z; vx[4 * q + 3] = fx.w;
			vy[4 * q] = fy.x; vy[4 * q + 1] = fy.y; vy[4 * q + 2] = fy.z; vy[4 * q + 3] = fy.w;
		}
		float o[5][4];
#pragma unroll
		for (int m = 0; m < 5; m++)
#pragma unroll
			for (int j = 0; j < 4; j++) o[m][j] = 0.f;
#pragma unroll
		for (int i = 0; i < 14; i++) {
			const float xx = vx[i] * vx[i], yy = vy[i] * vy[i], xy = vx[i] * vy[i];
#pragma unroll
			for (int j = 0; j < 4; j++) {
				const int k = i - j;   // tap index of input i for output j
				if (k >= 0 && k < kWin) {
					const float w = a.w[k];
					o[0][j] += w * vx[i]; o[1][j] += w * vy[i];
					o[2][j] += w * xx; o[3][j] += w * yy; o[4][j] += w * xy;
				}
			}
		}
#pragma unroll
		for (int m = 0; m < 5; m++)
			*reinterpret_cast<float4*>(&s_h[m][r][4 * g]) = make_float4(o[m][0], o[m][1], o[m][2], o[m][3]);
	}
	__syncthreads();
	// vertical pass: column c, output rows 4 rg .. 4 rg + 3
	const int c = tid % kLossTile, rg = tid / kLossTile;
	float o[5][4];
#pragma unroll
	for (int m = 0; m < 5; m++)
#pragma unroll
		for (int j = 0; j < 4; j++) o[m][j] = 0.f;
#pragma unroll
	for (int i = 0; i < 14; i++) {
		float v[5];
#pragma unroll
		for (int m = 0; m < 5; m++) v[m] = s_h[m][4 * rg + i][c];
#pragma unroll
		for (int j = 0; j < 4; j++) {
			const int k = i - j;
			if (k >= 0 && k < kWin) {
				const float w = a.w[k];
#pragma unroll
				for (int m = 0; m < 5; m++) o[m][j] += w * v[m];
			}
		}
	}
	float l1 = 0.f, ssim = 0.f;
	const int x = blockIdx.x * kLossTile + c;
#pragma unroll
	for (int j = 0; j < 4; j++) {
		const int ly = 4 * rg + j, y = blockIdx.y * kLossTile + ly;
		if (x < a.W && y < a.H_used) {
			const float mu1 = o[0][j], mu2 = o[1][j], e11 = o[2][j], e22 = o[3][j], e12 = o[4][j];
			// loss_utils.h:93-108
			const float mu1_sq = mu1 * mu1, mu2_sq = mu2 * mu2, mu1_mu2 = mu1 * mu2;
			const float sigma1_sq = e11 - mu1_sq, sigma2_sq = e22 - mu2_sq, sigma12 = e12 - mu1_mu2;
			const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
			const float A = mu1_sq + mu2_sq + C1, B = sigma1_sq + sigma2_sq + C2;
			const float C = 2.f * mu1_mu2 + C1, D = 2.f * sigma12 + C2;
			const float invAB = 1.f / (A * B);
			ssim += C * D * invAB;
			// partial derivatives of the map (sigma's dependence on mu1 folded into dm_dmu1)
			const float dm_dmu1 = 2.f * invAB * (mu2 * (D - C) - mu1 * C * D / A + mu1 * C * D / B);
			const float dm_ds1 = -C * D * invAB / B;
			const float dm_ds12 = 2.f * C * invAB;
			const size_t i = (size_t)y * a.W + x;
			a.dmaps[(0 * 3 + ch) * HW + i] = dm_dmu1;
			a.dmaps[(1 * 3 + ch) * HW + i] = dm_ds1;
			a.dmaps[(2 * 3 + ch) * HW + i] = dm_ds12;
			l1 += fabsf(s_x[ly + kHalf][c + kHalf] - s_y[ly + kHalf][c + kHalf]);
		}
	}
	l1 = warp_sum(l1);
	ssim = warp_sum(ssim);
	if ((tid & 31) == 0) { s_red[0][tid >> 5] = l1; s_red[1][tid >> 5] = ssim; }
	__syncthreads();
	if (tid < 2) {
		double t = 0.0;
		for (int w = 0; w < kLossThreads / 32; w++) t += (double)s_red[tid][w];
		atomicAdd(&a.sums[tid], t);
	}
}

__global__ void __launch_bounds__(kLossThreads) ssim_l1_bwd_kernel(const LossArgs a)
{
	__shared__ __align__(16) float s_d[3][kHalo][kHaloPitch];
	__shared__ __align__(16) float s_h[3][kHalo][kLossTile];
	const int tid = threadIdx.x;
	const int ch = blockIdx.z;
	const int x0 = blockIdx.x * kLossTile - kHalf, y0 = blockIdx.y * kLossTile - kHalf;
	const size_t HW = (size_t)a.H * a.W;
	const double count = 3.0 * (double)a.W * (double)a.H_used;

	if (blockIdx.x == 0 && blockIdx.y == 0 && ch == 0 && tid == 0) {
		const double l1 = a.sums[0] / count, ssim = a.sums[1] / count;
		a.loss_out[0] = (float)((1.0 - (double)a.lambda_dssim) * l1 + (double)a.lambda_dssim * (1.0 - ssim));
		a.loss_out[1] = (float)l1;
		a.loss_out[2] = (float)ssim;
	}
	for (int i = tid; i < kHalo * kHaloPitch; i += kLossThreads) {
		const int r = i / kHaloPitch, c = i % kHaloPitch;
		const int x = x0 + c, y = y0 + r;
		const bool in = c < kHalo && x >= 0 && x < a.W && y >= 0 && y < a.H_used;
		const size_t p = (size_t)y * a.W + x;
#pragma unroll
		for (int m = 0; m < 3; m++) s_d[m][r][c] = in ? a.dmaps[(m * 3 + ch) * HW + p] : 0.f;
	}
	__syncthreads();
	for (int item = tid; item < kHalo * kGroups; item += kLossThreads) {
		const int r = item / kGroups, g = item % kGroups;
		float o[3][4];
#pragma unroll
		for (int m = 0; m < 3; m++) {
			float v[16];
#pragma unroll
			for (int q = 0; q < 4; q++) {
				const float4 f = *reinterpret_cast<const float4*>(&s_d[m][r][4 * g + 4 * q]);
				v[4 * q] = f.x; v[4 * q + 1] = f.y; v[4 * q + 2] = f.z; v[4 * q + 3] = f.w;
			}
#pragma unroll
			for (int j = 0; j < 4; j++) {
				float t = 0.f;
#pragma unroll
				for (int k = 0; k < kWin; k++) t += a.w[k] * v[j + k];
				o[m][j] = t;
			}
			*reinterpret_cast<float4*>(&s_h[m][r][4 * g]) = make_float4(o[m][0], o[m][1], o[m][2], o[m][3]);
		}
	}
	__syncthreads();
	const int c = tid % kLossTile, rg = tid / kLossTile;
	float o[3][4];
#pragma unroll
	for (int m = 0; m < 3; m++)
#pragma unroll
		for (int j = 0; j < 4; j++) o[m][j] = 0.f;
#pragma unroll
	for (int i = 0; i < 14; i++) {
		float v[3];
#pragma unroll
		for (int m = 0; m < 3; m++) v[m] = s_h[m][4 * rg + i][c];
#pragma unroll
		for (int j = 0; j < 4; j++) {
			const int k = i - j;
			if (k >= 0 && k < kWin) {
#pragma unroll
				for (int m = 0; m < 3; m++) o[m][j] += a.w[k] * v[m];
			}
		}
	}
	const int x = blockIdx.x * kLossTile + c;
	if (x >= a.W) return;
#pragma unroll
	for (int j = 0; j < 4; j++) {
		const int y = blockIdx.y * kLossTile + 4 * rg + j;
		if (y >= a.H) continue;
		const size_t i = (size_t)y * a.W + x;
		float grad = 0.f;
		if (y < a.H_used) {
			const float I = masked_pixel(a, ch, x, y), g = a.gt[ch * HW + i];
			const float dssim = o[0][j] + 2.f * I * o[1][j] + g * o[2][j];
			const float diff = I - g;
			const float sgn = (diff > 0.f) ? 1.f : ((diff < 0.f) ? -1.f : 0.f);
			const float inv = (float)(1.0 / count);
			grad = (1.f - a.lambda_dssim) * sgn * inv - a.lambda_dssim * dssim * inv;
			if (a.mask_channels == 1) grad *= a.mask[i];
			else if (a.mask_channels == 3) grad *= a.mask[ch * HW + i];
		}
		a.dL_dpix[ch * HW + i] = grad;
	}
}

// lambda_dssim == 0: the SSIM term has no weight, so the loss is the L1 term alone — one streaming pass that writes
// dL/dI = sign(I - gt) / count (times the mask) and adds up |I - gt| (a double per CTA into sums[0]).
constexpr int kL1Threads = 256;
__global__ void __launch_bounds__(kL1Threads) l1_only_kernel(const LossArgs a)
{
	__shared__ double s_red[kL1Threads / 32];
	const size_t HW = (size_t)a.H * a.W, used = (size_t)a.H_used * a.W;
	const float inv = (float)(1.0 / (3.0 * (double)a.W * (double)a.H_used));
	// 128-bit accesses when the channel planes, the crop line and the four base pointers are 16-byte aligned
	const bool vec = (HW % 4 == 0) && (used % 4 == 0) &&
	                 ((reinterpret_cast<uintptr_t>(a.rendered) | reinterpret_cast<uintptr_t>(a.gt) |
	                   reinterpret_cast<uintptr_t>(a.mask) | reinterpret_cast<uintptr_t>(a.dL_dpix)) % 16 == 0);
	float acc = 0.f;   // per-thread partial (a few hundred terms); doubles from the warp reduction on
	auto element = [&](float I, float g, float m, float& grad) {
		const float diff = I * m - g;
		acc += fabsf(diff);
		grad = ((diff > 0.f) ? inv : ((diff < 0.f) ? -inv : 0.f)) * m;
	};
	for (int ch = 0; ch < 3; ch++) {
		const float* I = a.rendered + ch * HW;
		const float* G = a.gt + ch * HW;
		const float* M = a.mask_channels == 0 ? nullptr : (a.mask_channels == 1 ? a.mask : a.mask + ch * HW);
		float* D = a.dL_dpix + ch * HW;
		const size_t stride = (size_t)gridDim.x * kL1Threads, first = (size_t)blockIdx.x * kL1Threads + threadIdx.x;
		if (vec) {
			for (size_t q = first; q < HW / 4; q += stride) {
				float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
				if (4 * q < used) {
					const float4 x = reinterpret_cast<const float4*>(I)[q], g = reinterpret_cast<const float4*>(G)[q];
					const float4 m = M ? reinterpret_cast<const float4*>(M)[q] : make_float4(1.f, 1.f, 1.f, 1.f);
					element(x.x, g.x, m.x, d.x); element(x.y, g.y, m.y, d.y);
					element(x.z, g.z, m.z, d.z); element(x.w, g.w, m.w, d.w);
				}
				reinterpret_cast<float4*>(D)[q] = d;
			}
		} else {
			for (size_t i = first; i < HW; i += stride) {
				float d = 0.f;
				if (i < used) element(I[i], G[i], M ? M[i] : 1.f, d);
				D[i] = d;
			}
		}
	}
	double t = (double)acc;
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
	if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = t;
	__syncthreads();
	if (threadIdx.x == 0) {
		double sum = 0.0;
		for (int w = 0; w < kL1Threads / 32; w++) sum += s_red[w];
		atomicAdd(&a.sums[0], sum);
	}
}
__global__ void l1_only_finish_kernel(const LossArgs a)
{
	const float l1 = (float)(a.sums[0] / (3.0 * (double)a.W * (double)a.H_used));
	a.loss_out[0] = l1;
	a.loss_out[1] = l1;
	a.loss_out[2] = 0.f;   // SSIM is not evaluated when it has no weight
}

int launch_photometric_loss(int W, int H, int H_used, float lambda_dssim, const float* rendered, const float* gt,
                            const float* mask, int mask_channels, float* workspace, float* loss_out, float* dL_dpix,
                            cudaStream_t st)
{
	LossArgs a{};
	a.W = W; a.H = H; a.H_used = H_used; a.mask_channels = mask ? mask_channels : 0; a.lambda_dssim = lambda_dssim;
	// loss_utils.h:58-68: exp(-(x - 5)^2 / (2 sigma^2)) in float32, normalised by its float32 sum
	float wsum = 0.f;
	for (int k = 0; k < kWin; k++) {
		const int t = k - kWin / 2;
		a.w[k] = std::exp(-(float)(t * t) / (2.0f * 1.5f * 1.5f));
		wsum += a.w[k];
	}
	for (int k = 0; k < kWin; k++) a.w[k] /= wsum;
	a.rendered = rendered; a.gt = gt; a.mask = mask;
	a.sums = reinterpret_cast<double*>(workspace);
	a.dmaps = workspace + 4;   // behind the two 8-byte sums
	a.loss_out = loss_out; a.dL_dpix = dL_dpix;
	OGS_CUDA_TRY(cudaMemsetAsync(a.sums, 0, 2 * sizeof(double), st));
	if (lambda_dssim == 0.f) {
		l1_only_kernel<<<kNumSMs * 8, kL1Threads, 0, st>>>(a);
		l1_only_finish_kernel<<<1, 1, 0, st>>>(a);
		OGS_CUDA_TRY(cudaGetLastError());
		return OGS_OK;
	}
	const dim3 grid(ceil_div(W, kLossTile), ceil_div(H, kLossTile), 3);
	ssim_l1_fwd_kernel<<<grid, kLossThreads, 0, st>>>(a);
	ssim_l1_bwd_kernel<<<grid, kLossThreads, 0, st>>>(a);
	OGS_CUDA_TRY(cudaGetLastError());
	return OGS_OK;
}

size_t photometric_loss_workspace_bytes(int W, int H) { return (4 + 9 * (size_t)W * H) * sizeof(float); }

// ------------------------------------------------------------------ Adam, all parameter groups in one launch
// torch/csrc/api/src/optim/adam.cpp step(): exp_avg, exp_avg_sq, denom, addcdiv_ for one element
OGS_D void adam_element(const AdamLaunch& a, float step_size, float grad, float& m, float& v, float& p)
{
	m = m * a.beta1 + grad * a.one_minus_beta1;
	v = v * a.beta2 + (a.one_minus_beta2 * grad) * grad;
	const float denom = sqrtf(v) / a.sqrt_bias_correction2 + a.eps;
	p = p - step_size * (m / denom);
}

__global__ void __launch_bounds__(256) adam_groups_kernel(const AdamLaunch a)
{
	// the block's group: a.first_block[g] <= blockIdx.x < a.first_block[g + 1]
	int g = 0;
#pragma unroll
	for (int k = 1; k < kAdamMaxGroups; k++)
		if (k < a.groups && (int)blockIdx.x >= a.first_block[k]) g = k;
	const AdamGroup grp = a.group[g];
	// four consecutive elements per thread: 128-bit accesses when the group's tensors are 16-byte aligned
	const size_t i = ((size_t)blockIdx.x - a.first_block[g]) * kAdamPerBlock + (size_t)threadIdx.x * 4;
	if (i >= grp.n) return;
	if (grp.vec4 && i + 4 <= grp.n) {
		const float4 gr = *reinterpret_cast<const float4*>(grp.grad + i);
		float4 m = *reinterpret_cast<float4*>(grp.exp_avg + i);
		float4 v = *reinterpret_cast<float4*>(grp.exp_avg_sq + i);
		float4 p = *reinterpret_cast<float4*>(grp.param + i);
		adam_element(a, grp.step_size, gr.x, m.x, v.x, p.x);
		adam_element(a, grp.step_size, gr.y, m.y, v.y, p.y);
		adam_element(a, grp.step_size, gr.z, m.z, v.z, p.z);
		adam_element(a, grp.step_size, gr.w, m.w, v.w, p.w);
		*reinterpret_cast<float4*>(grp.exp_avg + i) = m;
		*reinterpret_cast<float4*>(grp.exp_avg_sq + i) = v;
		*reinterpret_cast<float4*>(grp.param + i) = p;
	} else {
		for (size_t k = i; k < min(i + 4, grp.n); k++) {
			float m = grp.exp_avg[k], v = grp.exp_avg_sq[k], p = grp.param[k];
			adam_element(a, grp.step_size, grp.grad[k], m, v, p);
			grp.exp_avg[k] = m; grp.exp_avg_sq[k] = v; grp.param[k] = p;
		}
	}
}

int launch_adam(AdamLaunch a, cudaStream_t st)
{
	int blocks = 0;
	for (int g = 0; g < a.groups; g++) {
		AdamGroup& grp = a.group[g];
		grp.vec4 = ((reinterpret_cast<uintptr_t>(grp.param) | reinterpret_cast<uintptr_t>(grp.grad) |
		             reinterpret_cast<uintptr_t>(grp.exp_avg) | reinterpret_cast<uintptr_t>(grp.exp_avg_sq)) & 15u) == 0;
		a.first_block[g] = blocks;
		blocks += (int)((a.group[g].n + kAdamPerBlock - 1) / kAdamPerBlock);
	}
	a.first_block[a.groups] = blocks;
	if (blocks == 0) return OGS_OK;
	adam_groups_kernel<<<blocks, 256, 0, st>>>(a);
	OGS_CUDA_TRY(cudaGetLastError());
	return OGS_OK;
}

// ------------------------------------------------------------------ densification statistics
// gaussian_mapper.cpp:427-434 + gaussian_model.cpp:839-853, for the Gaussians with radii > 0:
//   max_radii2D = max(max_radii2D, radii);  xyz_gradient_accum += ||dL/dmean2D.xy||;  denom += 1
__global__ void densify_stats_kernel(int P, const int* __restrict__ radii, const float* __restrict__ dL_dmean2D,
                                     float* __restrict__ max_radii2D, float* __restrict__ xyz_gradient_accum,
                                     float* __restrict__ denom)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= P) return;
	const int r = radii[i];
	if (r <= 0) return;
	const float gx = dL_dmean2D[3 * (size_t)i], gy = dL_dmean2D[3 * (size_t)i + 1];
	max_radii2D[i] = fmaxf(max_radii2D[i], (float)r);
	xyz_gradient_accum[i] += sqrtf(gx * gx + gy * gy);
	denom[i] += 1.f;
}

// Per-view increments of the same statistics as three plain arrays (what data-parallel ranks exchange before
// applying them: sum, sum, max).
__global__ void view_stats_kernel(int P, const int* __restrict__ radii, const float* __restrict__ dL_dmean2D,
                                  float* __restrict__ grad_norm, float* __restrict__ visible, float* __restrict__ radius)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= P) return;
	const int r = radii[i];
	const float gx = dL_dmean2D[3 * (size_t)i], gy = dL_dmean2D[3 * (size_t)i + 1];
	grad_norm[i] = r > 0 ? sqrtf(gx * gx + gy * gy) : 0.f;
	visible[i] = r > 0 ? 1.f : 0.f;
	radius[i] = (float)r;
}

int launch_view_stats(int P, const int* radii, const float* dL_dmean2D, float* grad_norm, float* visible, float* radius,
                      cudaStream_t st)
{
	if (P == 0) return OGS_OK;
	view_stats_kernel<<<ceil_div(P, 256), 256, 0, st>>>(P, radii, dL_dmean2D, grad_norm, visible, radius);
	OGS_CUDA_TRY(cudaGetLastError());
	return OGS_OK;
}

int launch_densify_stats(int P, const int* radii, const float* dL_dmean2D, float* max_radii2D,
                         float* xyz_gradient_accum, float* denom, cudaStream_t st)
{
	if (P == 0) return OGS_OK;
	densify_stats_kernel<<<ceil_div(P, 256), 256, 0, st>>>(P, radii, dL_dmean2D, max_radii2D, xyz_gradient_accum, denom);
	OGS_CUDA_TRY(cudaGetLastError());
	return OGS_OK;
}

} // namespace ogs
