// launchers.cuh — kernel argument blocks and host-side launchers shared between the
// translation units of libomnigs_b200.so.
#pragma once
#include <cstdlib>
#include "ogs_common.cuh"

namespace ogs {

struct PreprocessFwdArgs {
	int P, D, M, W, H, gx, gy, band_y0, band_y1, seam_wrap;
	float scale_modifier;
	const float* means3D;
	const float* shs;
	const float* colors_precomp;
	const float* opacities;
	const float* scales;
	const float* rotations;
	const float* cov3D_precomp;
	// raw-parameter mode (raw != 0): opacities / scales / rotations hold the stored (pre-activation)
	// tensors and the SH row is split into features_dc [P,1,3] and features_rest [P,M-1,3]
	int raw;
	const float* features_dc;
	const float* features_rest;
	// defer_colors != 0: geometry only — colours and clamp masks are filled in later by launch_sh_colors (the data-parallel
	// trainer overlaps the exchange of the SH gradients with this kernel, the depth order and the tile sort)
	int defer_colors;
	// perspective camera (pinhole != 0): full projection matrix, focal lengths in pixels, tan(fov/2)
	int pinhole, render_depth;
	const float* projmatrix;
	float focal_x, focal_y, tan_fovx, tan_fovy;
	const float* viewmatrix;
	const float* campos;
	int* radii;
	float4* g0;
	float4* g1;
	float2* gb;
	float* depth;
	uint2* rect;
	uint32_t* tiles_touched;
	float* cov3D;
	uint8_t* clamped;
	uint32_t* sort_key;
	int* tile_diff;
	unsigned long long* total_tiles;   // scalars[0] = sum tiles_touched; scalars[7] = seam-wrap flag of this frame
};
struct PreprocessBwdArgs {
	int P, D, M, W, H;
	float scale_modifier;
	const float* means3D;
	const float* shs;
	const float* scales;
	const float* rotations;
	const float* cov3D;
	const float* viewmatrix;
	const float* campos;
	const int* radii;
	const uint8_t* clamped;
	const float* grad_acc;   // [P,12] raw render-backward sums
	int first_block, num_blocks;   // sub-range of the 128-Gaussian blocks to process (num_blocks 0 = all from first_block)
	const float4* g0;        // conic.xy in .zw
	const float4* g1;        // conic.z in .x, activated opacity in .y
	// raw-parameter mode (raw != 0): scales / rotations are the stored tensors, the SH row is split, and the
	// outputs are gradients w.r.t. the stored tensors (dL_dopacity, dL_dscale, dL_drot, dL_dfeatures_*);
	// dL_dmean2D and dL_dcov3D are optional, dL_dcolor / dL_dsh unused
	int raw;
	const float* features_dc;
	const float* features_rest;
	float* dL_dfeatures_dc;
	float* dL_dfeatures_rest;
	// perspective camera (pinhole != 0)
	int pinhole;
	const float* projmatrix;
	float focal_x, focal_y, tan_fovx, tan_fovy;
	float* dL_dmean2D;
	float* dL_dconic;
	float* dL_dopacity;
	float* dL_dcolor;
	float* dL_dmean3D;
	float* dL_dcov3D;
	float* dL_dsh;
	float* dL_dscale;
	float* dL_drot;
	// multi-view / data-parallel mode: accumulate != 0 ADDS this view to dL_dmean3D / dL_dopacity /
	// dL_dscale / dL_drot and to the statistics; dL_drgb_view != NULL replaces the dL/dsh (dL/dfeatures) row by the
	// view's clamp-masked dL/dRGB [P,3] (see launch_sh_gradient_from_views); stat_* (all three or none): the view's
	// densification statistics (sum of |dL_dmean2D.xy|, visibility count, max radius as float)
	int accumulate;
	float* dL_drgb_view;
	float* stat_grad_norm;
	float* stat_visible;
	float* stat_max_radius;
};
// dL/dsh [P,M,3] (or split dL/dfeatures_dc [P,1,3] + dL/dfeatures_rest [P,M-1,3]) = sum over views, in view order, of
// b_k(direction from the view's camera centre) * dL/dRGB_view.  drgb[v] may point into a peer GPU's memory.
constexpr int kMaxStepViews = 16;
struct ShFromViewsArgs {
	int P, D, M, n_views;
	const float* means3D;
	const float* campos;                 // n_views x 3 (device)
	const float* drgb[kMaxStepViews];    // each [P,3]
	float* dL_dsh;                       // [P,M,3] or NULL
	float* dL_dfeatures_dc;              // split layout (raw-parameter trainer) when dL_dsh is NULL
	float* dL_dfeatures_rest;
};
int launch_sh_gradient_from_views(const ShFromViewsArgs& a, cudaStream_t st);
int launch_preprocess_fwd(const PreprocessFwdArgs& a, cudaStream_t st);
// SH -> RGB + clamp mask of every Gaussian with radii > 0 into the packed records (forward.cu:688-692 with :30-83)
int launch_sh_colors(int P, int D, int M, const float* means3D, const float* shs, const float* campos, const int* radii,
                     float4* g1, float2* gb, uint8_t* clamped, cudaStream_t st);
int launch_mark_all_visible(int P, uint8_t* present, cudaStream_t st);
int launch_check_frustum(int P, const float* means3D, const float* viewmatrix, uint8_t* present, cudaStream_t st);
int launch_depth_order(const GeomState& g, int P, cudaStream_t st);
int launch_tile_ranges(const ImageState& img, int W, int H, cudaStream_t st);
int launch_emit_and_tile_sort(const GeomState& g, const ImageState& img, const BinningState& b,
                              int P, int64_t R, int W, int H, cudaStream_t st);
int launch_rebuild_keys(const ImageState& img, const BinningState& b, const GeomState& g, int W, int H,
                        uint64_t* keys, cudaStream_t st);
// hit_bytes: one byte per list entry (R bytes; the dead tile-key array of the binning buffer): bit w = the entry can reach
// the 8x4 sub-block w of its tile, as decided by the forward blend; the backward blend consumes it
int launch_render_fwd(const uint2* ranges, const uint32_t* point_list, int W, int H,
                      const float4* g0, const float4* g1, const float2* gb, const unsigned long long* scalars, const float* bg,
                      float* final_T, uint32_t* n_contrib, float* out_color, uint8_t* hit_bytes, cudaStream_t st);
// OGS_FWD_STAGING=ldgsts|bulk: the asynchronous-staging experiment kernels (1 / 2), which do not write hit bytes
inline int fwd_staging_mode()
{
	static const int v = [] {
		const char* e = getenv("OGS_FWD_STAGING");
		return !e ? 0 : (e[0] == 'l' ? 1 : (e[0] == 'b' ? 2 : 0));
	}();
	return v;
}
// the backward takes the forward's sub-block decisions unless the experiment forward ran or OGS_BWD_HITS=0 (A/B runs)
inline bool render_hit_bytes_enabled()
{
	static const bool v = [] { const char* e = getenv("OGS_BWD_HITS"); return !(e && e[0] == '0'); }();
	return v && fwd_staging_mode() == 0;
}
int launch_pair_count(const uint2* ranges, const uint32_t* point_list, int W, int H, const float4* g0, const float4* g1,
                      const uint32_t* n_contrib, unsigned long long* counts, cudaStream_t st);
int launch_render_bwd(const uint2* ranges, const uint32_t* point_list, int W, int H, const float* bg,
                      const float4* g0, const float4* g1, const float2* gb, const unsigned long long* scalars,
                      const float* final_T, const uint32_t* n_contrib, const float* dL_dpix,
                      float* grad_acc, const uint8_t* hit_bytes, cudaStream_t st);
int launch_preprocess_bwd(const PreprocessBwdArgs& a, cudaStream_t st);

// ---- train_step.cu: the per-iteration work either side of the rasterizer
size_t photometric_loss_workspace_bytes(int W, int H);
int launch_photometric_loss(int W, int H, int H_used, float lambda_dssim, const float* rendered, const float* gt,
                            const float* mask, int mask_channels, float* workspace, float* loss_out, float* dL_dpix,
                            cudaStream_t st);
constexpr int kAdamMaxGroups = 8;
constexpr int kAdamPerBlock = 1024;
struct AdamGroup {
	float* param;
	const float* grad;
	float* exp_avg;
	float* exp_avg_sq;
	size_t n;
	float step_size;             // lr / (1 - beta1^t), evaluated in double on the host like torch::optim::Adam
	int vec4;                    // all four tensors 16-byte aligned (set by launch_adam)
};
struct AdamLaunch {
	int groups;
	int first_block[kAdamMaxGroups + 1];
	float beta1, beta2, eps, sqrt_bias_correction2;
	float one_minus_beta1, one_minus_beta2;   // 1 - beta evaluated in double (1 - 0.999f is off by 1.3e-5)
	AdamGroup group[kAdamMaxGroups];
};
int launch_adam(AdamLaunch a, cudaStream_t st);
int launch_view_stats(int P, const int* radii, const float* dL_dmean2D, float* grad_norm, float* visible, float* radius,
                      cudaStream_t st);
// ---- peer_collective.cu
constexpr int kMaxPeers = 8;
// the first count_sum floats are summed, the count_max floats behind them (non-negative: ordered like their bit patterns)
// are max-reduced
int launch_peer_allreduce_sum(float* const* bufs, int world, int rank, size_t count, cudaStream_t st);
int launch_multimem_allreduce_sum(float* multicast, int world, int rank, size_t count, cudaStream_t st);
int launch_peer_allreduce(float* const* bufs, int world, int rank, size_t count_sum, size_t count_max, cudaStream_t st);
int launch_multimem_allreduce(float* multicast, int world, int rank, size_t count_sum, size_t count_max, cudaStream_t st);
int launch_band_rows_allgather(float* const* images, int world, const float* src, int W, int H, int y0, int y1, cudaStream_t st);
int launch_densify_stats(int P, const int* radii, const float* dL_dmean2D, float* max_radii2D,
                         float* xyz_gradient_accum, float* denom, cudaStream_t st);


} // namespace ogs
