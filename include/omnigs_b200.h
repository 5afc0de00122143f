/*
 * omnigs_b200.h — C ABI of libomnigs_b200.so: the B200-native (sm_100a) equirectangular
 * ("lonlat", camera_type = 3) Gaussian-splatting rasterizer.
 *
 * This is the drop-in boundary underneath the reference's LibTorch entry points
 * (raikuma/OmniGS-fork, paths relative to its root):
 *
 *   reference interface                                       replaced by
 *   --------------------------------------------------------  ------------------------------------
 *   CudaRasterizer::LonlatRasterizer::forward                  ogs_lonlat_forward_stage1 +
 *     cuda_rasterizer/rasterizer.h:102-122,                      ogs_lonlat_forward_stage2
 *     rasterizer_impl.cu:540-697
 *   CudaRasterizer::LonlatRasterizer::backward                 ogs_lonlat_backward
 *     rasterizer.h:125-154, rasterizer_impl.cu:701-795
 *   CudaRasterizer::LonlatRasterizer::markVisible              ogs_mark_all_visible
 *     rasterizer.h:98-100, rasterizer_impl.cu:185-192
 *   required<GeometryState/ImageState/BinningState>(n)         ogs_geom_bytes / ogs_img_bytes /
 *     rasterizer_impl.h:96-102, rasterizer_impl.cu:198-245       ogs_binning_bytes
 *   CudaRasterizer::Rasterizer::{forward,backward,markVisible}  ogs_pinhole_forward_stage1 / ogs_pinhole_backward /
 *     rasterizer.h:39-92, rasterizer_impl.cu:170-183,250-530     ogs_mark_visible_pinhole
 *   model activations + rasterizer + their autograd            ogs_lonlat_forward_raw_stage1 /
 *     gaussian_model.cpp:54-77, gaussian_renderer.cpp:212-258    ogs_lonlat_backward_raw
 *   loss_utils::l1_loss / ssim + autograd                      ogs_photometric_loss
 *     loss_utils.h:31-34,58-131, gaussian_mapper.cpp:391-413
 *   torch::optim::Adam::step over the six groups               ogs_adam_step
 *     gaussian_model.cpp:485-511
 *   max_radii2D / addDensificationStats                        ogs_densify_stats
 *     gaussian_mapper.cpp:427-434, gaussian_model.cpp:839-853
 *
 * The reference passes std::function<char*(size_t)> allocators (rasterize_points.cu:41-47,92-94);
 * a C ABI cannot, so the forward is split in two at the one point where a buffer size depends on a
 * device result (num_rendered, rasterizer_impl.cu:627-632): stage 1 returns num_rendered, the caller
 * sizes and allocates the binning buffer, stage 2 finishes the frame.
 *
 * Conventions (identical to the reference unless stated):
 *   - every pointer is a DEVICE pointer (float32 / int32 / uint8) unless the name says "host";
 *   - optional inputs are NULL when absent: exactly one of {shs, colors_precomp} and one of
 *     {(scales, rotations), cov3D_precomp} is non-NULL (gaussian_rasterizer.cpp:190-208);
 *   - viewmatrix is Tcw stored column-major (16 floats), campos the camera centre (3 floats),
 *     rotations are (w,x,y,z) and are NOT normalised by the kernels (forward.cu:203);
 *   - the three byte buffers are opaque: their layout is private to this library (it is NOT the
 *     reference's layout) and only has to survive from forward to backward of the same frame;
 *   - `stream` is a cudaStream_t (NULL = legacy default stream).  All work is ordered on it: stage 1 forks one small
 *     kernel (the tile ranges) onto a library-owned non-blocking stream beside the depth-order passes and joins it
 *     back with events before returning, so the caller sees ordinary stream semantics (OGS_SIDE_STREAM=0 disables it);
 *   - all calls return 0 on success or a negative OGS_ERR_* code; ogs_last_error() gives text.
 *     Like the reference, kernels are launched asynchronously; only stage 1 blocks (for the
 *     num_rendered read-back, as the reference does at rasterizer_impl.cu:627-628).
 */
#ifndef OMNIGS_B200_H
#define OMNIGS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OGS_OK 0
#define OGS_ERR_INVALID_ARG (-1)   /* NULL/negative/inconsistent argument                      */
#define OGS_ERR_CUDA (-2)          /* a CUDA runtime call failed (see ogs_last_error)            */
#define OGS_ERR_TOO_MANY (-3)      /* num_rendered >= 2^31 (the reference's int) or image larger than 2^16 tiles/axis  */
#define OGS_ERR_NO_DEVICE (-4)     /* no usable sm_100 device                                    */

#define OGS_ABI_VERSION 2

#if defined(__GNUC__)
#define OGS_API __attribute__((visibility("default")))
#else
#define OGS_API
#endif

OGS_API int ogs_abi_version(void);
/* Text of the last error raised on the calling thread ("" if none). */
OGS_API const char* ogs_last_error(void);

/* ---- workspace sizes (replace required<...State>(n), rasterizer_impl.h:96-102) ---- */
OGS_API size_t ogs_geom_bytes(int P);
OGS_API size_t ogs_img_bytes(int W, int H);
OGS_API size_t ogs_binning_bytes(int64_t num_rendered, int W, int H);

/*
 * Forward, stage 1: per-Gaussian preprocessing (forward.cu:593-703), the tile-count prefix
 * structures, the depth ordering of the Gaussians and the num_rendered read-back
 * (rasterizer_impl.cu:592-628).  Writes radii[P] (0 for culled Gaussians) and *num_rendered_host.
 * geom_buffer >= ogs_geom_bytes(P), img_buffer >= ogs_img_bytes(W,H).
 */
OGS_API int ogs_lonlat_forward_stage1(
	int P, int D, int M, int W, int H,
	const float* means3D, const float* shs, const float* colors_precomp, const float* opacities,
	const float* scales, float scale_modifier, const float* rotations, const float* cov3D_precomp,
	const float* viewmatrix, const float* campos,
	int* radii, char* geom_buffer, char* img_buffer,
	int64_t* num_rendered_host, void* stream);

/*
 * Forward, stage 2: tile-instance emission (duplicateWithKeys, rasterizer_impl.cu:94-140), the
 * radix sort (:651-661), tile ranges (:664-679) and the alpha-blend (forward.cu:346-467).
 * binning_buffer >= ogs_binning_bytes(num_rendered, W, H).  out_color is planar [3,H,W] and is
 * fully written (background where nothing is rendered).
 */
OGS_API int ogs_lonlat_forward_stage2(
	int P, int W, int H, int64_t num_rendered, const float* background,
	char* geom_buffer, char* binning_buffer, char* img_buffer,
	float* out_color, void* stream);
/* The two halves of stage 2: ..._bin = emission + sort (rasterizer_impl.cu:632-679, needs no colours),
 * ..._blend = renderCUDA (forward.cu:346-467).  stage2 == bin followed by blend; bin may be repeated on the same buffers. */
OGS_API int ogs_lonlat_forward_bin(
	int P, int W, int H, int64_t num_rendered, char* geom_buffer, char* binning_buffer, char* img_buffer, void* stream);
OGS_API int ogs_lonlat_forward_blend(
	int P, int W, int H, int64_t num_rendered, const float* background,
	char* geom_buffer, char* binning_buffer, char* img_buffer, float* out_color, void* stream);
/*
 * Stage 1 without colours, and the colours on their own (data-parallel training): the SH
 * coefficients are the only parameters the geometry, the depth order and the tile sort do not read, so a trainer can
 * start a step with ..._stage1_geometry + ..._bin while the previous step's SH gradients are still being exchanged and
 * applied, then call ..._forward_colors (computeColorFromSH at its call site, forward.cu:688-692, for every Gaussian
 * with radii > 0; bit-identical colours and clamp masks) and ..._blend.
 */
OGS_API int ogs_lonlat_forward_stage1_geometry(
	int P, int W, int H,
	const float* means3D, const float* opacities, const float* scales, float scale_modifier, const float* rotations,
	const float* cov3D_precomp, const float* viewmatrix, const float* campos,
	int* radii, char* geom_buffer, char* img_buffer, int64_t* num_rendered_host, void* stream);
OGS_API int ogs_lonlat_forward_colors(
	int P, int D, int M, const float* means3D, const float* shs, const float* campos, const int* radii,
	char* geom_buffer, void* stream);

/*
 * Backward (rasterizer_impl.cu:701-795): render backward (backward.cu:672-843) then the fused
 * per-Gaussian backward (backward.cu:297-485 and :613-669).
 * Every element of every output is written (culled rows get exact zeros), so outputs need NOT be
 * zero-filled by the caller (the reference's shim zero-fills them, rasterize_points.cu:200-208).
 * Shapes: dL_dmean2D [P,3] (z = 0), dL_dcolor [P,3], dL_dopacity [P], dL_dmean3D [P,3],
 * dL_dcov3D [P,6], dL_dsh [P,M,3] (may be NULL when M == 0), dL_dscale [P,3], dL_drot [P,4].
 * dL_dconic [P,4] is optional (NULL to skip): the reference's intermediate (.x,.y,.w used).
 */
OGS_API int ogs_lonlat_backward(
	int P, int D, int M, int64_t num_rendered, int W, int H,
	const float* background,
	const float* means3D, const float* shs, const float* colors_precomp,
	const float* scales, float scale_modifier, const float* rotations, const float* cov3D_precomp,
	const float* viewmatrix, const float* campos, const int* radii,
	char* geom_buffer, char* binning_buffer, char* img_buffer,
	const float* dL_dpix,
	float* dL_dmean2D, float* dL_dconic, float* dL_dopacity, float* dL_dcolor,
	float* dL_dmean3D, float* dL_dcov3D, float* dL_dsh, float* dL_dscale, float* dL_drot,
	void* stream);

/*
 * The two halves of ogs_lonlat_backward, for callers that exchange data between them (latitude bands,
 * SURVEY.md §8(e-b)): ..._render zero-fills the packed accumulators ([P,12] float at byte
 * ogs_grad_acc_offset(P) of the 256-byte-aligned geometry buffer: per-Gaussian raw sums over pixels
 * u*dx, u*dy, u*dx^2, u*dx*dy, u*dy^2 (u = dL/dG * G), G*dL/dalpha, colour.rgb terms, 3 pad — linear in
 * the pixels, so band partials add up) and replays the blend of this rank's tiles into them
 * (backward.cu:672-843); the caller may sum them over ranks (48 B/Gaussian instead of 324);
 * ..._finish applies the conic factors (backward.cu:821-836) and runs the fused per-Gaussian backward
 * (backward.cu:297-485, :613-669) from them.
 */
OGS_API int ogs_lonlat_backward_render(
	int P, int64_t num_rendered, int W, int H, const float* background,
	char* geom_buffer, char* binning_buffer, char* img_buffer, const float* dL_dpix, void* stream);
OGS_API int ogs_lonlat_backward_finish(
	int P, int D, int M, int W, int H,
	const float* means3D, const float* shs, const float* scales, float scale_modifier, const float* rotations,
	const float* cov3D_precomp, const float* viewmatrix, const float* campos, const int* radii, char* geom_buffer,
	float* dL_dmean2D, float* dL_dconic, float* dL_dopacity, float* dL_dcolor,
	float* dL_dmean3D, float* dL_dcov3D, float* dL_dsh, float* dL_dscale, float* dL_drot, void* stream);
OGS_API size_t ogs_grad_acc_offset(int P);
/* The same two halves on caller-owned accumulators (grad_acc: [P,12] float, 16-byte aligned, e.g. in symmetric memory
 * so that the band ranks can sum it with ogs_peer_allreduce / ogs_multimem_allreduce instead of NCCL). */
OGS_API int ogs_lonlat_backward_render_into(
	int P, int64_t num_rendered, int W, int H, const float* background,
	char* geom_buffer, char* binning_buffer, char* img_buffer, const float* dL_dpix, float* grad_acc, void* stream);
OGS_API int ogs_lonlat_backward_finish_from(
	int P, int D, int M, int W, int H,
	const float* means3D, const float* shs, const float* scales, float scale_modifier, const float* rotations,
	const float* cov3D_precomp, const float* viewmatrix, const float* campos, const int* radii, char* geom_buffer,
	const float* grad_acc,
	float* dL_dmean2D, float* dL_dconic, float* dL_dopacity, float* dL_dcolor,
	float* dL_dmean3D, float* dL_dcov3D, float* dL_dsh, float* dL_dscale, float* dL_drot, void* stream);
/* ..._finish_from over Gaussians [first, first + count) only (first a multiple of 128; count 0 = up to P): lets the band
 * ranks pipeline the accumulator exchange — chunk k+1 is being summed over NVLink while chunk k is differentiated.
 * All pointers are the full arrays' bases. */
OGS_API int ogs_lonlat_backward_finish_range(
	int P, int D, int M, int W, int H,
	const float* means3D, const float* shs, const float* scales, float scale_modifier, const float* rotations,
	const float* cov3D_precomp, const float* viewmatrix, const float* campos, const int* radii, char* geom_buffer,
	const float* grad_acc, int first, int count,
	float* dL_dmean2D, float* dL_dconic, float* dL_dopacity, float* dL_dcolor,
	float* dL_dmean3D, float* dL_dcov3D, float* dL_dsh, float* dL_dscale, float* dL_drot, void* stream);

/*
 * One view of a multi-view / data-parallel training step (the reference trains one view per step
 * and has no multi-GPU code, so this is an extension whose oracle is "the sum of the reference's per-view gradients").
 * Like ogs_lonlat_backward, except that
 *   - the four geometry gradients the optimiser consumes (dL_dmean3D [P,3], dL_dopacity [P], dL_dscale [P,3],
 *     dL_drot [P,4]) are WRITTEN when accumulate == 0 (first view of the step) and ADDED TO otherwise;
 *   - the view's densification statistics (gaussian_mapper.cpp:427-434, gaussian_model.cpp:839-853) go the same way
 *     into stat_grad_norm (sum of |dL_dmean2D.xy|), stat_visible (count) and stat_max_radius (max radius, as float);
 *     pass all three or none;
 *   - instead of dL_dsh [P,M,3] the call leaves dL_drgb_view [P,3]: dL/dcolour with the clamped channels zeroed
 *     (zeros for culled Gaussians).  Row k of dL/dsh is b_k(direction) * dL/dRGB per view (backward.cu:60-150), so
 *     a step's dL/dsh is rebuilt ONCE from all its views' factors by ogs_sh_gradient_from_views: 12 bytes per
 *     Gaussian and view cross NVLink instead of 192.
 * dL_dmean2D [P,3] is optional.
 */
OGS_API int ogs_lonlat_backward_view(
	int P, int D, int M, int64_t num_rendered, int W, int H, const float* background,
	const float* means3D, const float* shs, const float* scales, float scale_modifier, const float* rotations,
	const float* viewmatrix, const float* campos, const int* radii,
	char* geom_buffer, char* binning_buffer, char* img_buffer, const float* dL_dpix,
	int accumulate, float* dL_dmean3D, float* dL_dopacity, float* dL_dscale, float* dL_drot,
	float* dL_drgb_view, float* stat_grad_norm, float* stat_visible, float* stat_max_radius,
	float* dL_dmean2D, void* stream);
/*
 * dL_dsh [P,16,3] (or, for the raw-parameter trainer, dL_dfeatures_dc [P,1,3] + dL_dfeatures_rest [P,15,3]; give one
 * of the two forms) = sum over the step's views, in view order, of b(direction from campos_views[v]) (x) dL_drgb_views[v].
 * Bit-identical to adding up the per-view dL_dsh tensors of ogs_lonlat_backward in that order.  campos_views is a
 * device array [n_views,3]; dL_drgb_views a HOST array of n_views <= 16 device pointers, which may point into peer
 * GPUs' memory (the kernel then reads the factors over NVLink: transfer and rebuild are one pass).
 */
OGS_API int ogs_sh_gradient_from_views(
	int P, int D, int M, int n_views, const float* means3D, const float* campos_views, const float* const* dL_drgb_views,
	float* dL_dsh, float* dL_dfeatures_dc, float* dL_dfeatures_rest, void* stream);

/* markAllVisible (rasterizer_impl.cu:82-90): present[i] = true for i < P (1 byte per flag). */
OGS_API int ogs_mark_all_visible(int P, uint8_t* present, void* stream);

/*
 * Latitude-band variant of stage 1: identical to stage 1 except that every
 * tile rect is clipped to tile rows [band_ty0, band_ty1) before counting, so this rank emits,
 * sorts and renders only its band.  radii are the unclipped reference radii.  With
 * band = [0, ceil(H/16)) it is exactly stage 1.
 */
OGS_API int ogs_lonlat_forward_stage1_band(
	int P, int D, int M, int W, int H, int band_ty0, int band_ty1,
	const float* means3D, const float* shs, const float* colors_precomp, const float* opacities,
	const float* scales, float scale_modifier, const float* rotations, const float* cov3D_precomp,
	const float* viewmatrix, const float* campos,
	int* radii, char* geom_buffer, char* img_buffer,
	int64_t* num_rendered_host, void* stream);

/*
 * Introspection for the parity tests: re-express our private state in the reference's terms
 * (the arrays GeometryState / BinningState / ImageState hold, rasterizer_impl.cu:198-245).
 * Any output pointer may be NULL.  Rows of culled Gaussians are zero-filled.
 */
OGS_API int ogs_export_geometry(
	int P, const char* geom_buffer,
	float* means2D /*[P,2]*/, float* depths /*[P]*/, float* conic_opacity /*[P,4]*/, float* rgb /*[P,3]*/,
	uint32_t* tiles_touched /*[P]*/, uint8_t* clamped /*[P,3]*/, float* cov3D /*[P,6]*/, void* stream);
OGS_API int ogs_export_binning(
	int P, int W, int H, int64_t num_rendered,
	const char* geom_buffer, const char* binning_buffer, const char* img_buffer,
	uint32_t* point_list /*[R]*/, uint64_t* point_list_keys /*[R]*/, uint32_t* ranges /*[T,2]*/,
	float* final_T /*[H*W]*/, uint32_t* n_contrib /*[H*W]*/, void* stream);

/* Measurement: the blending work of a rendered frame, independent of kernel organisation — counts (DEVICE, 4 x uint64):
 * [0] sum over pixels of n_contrib (list entries the reference's renderCUDA visits up to each pixel's last contributor,
 * forward.cu:403-455), [1] (pixel, Gaussian) pairs that blend, [2] entries a tile-synchronous kernel walks (the tile's
 * largest n_contrib, per pixel), [3] pixels.  bench.py turns them into the compute roofline of the blend kernels. */
OGS_API int ogs_export_pair_counts(int P, int W, int H, int64_t num_rendered, const char* geom_buffer,
                                   const char* binning_buffer, const char* img_buffer, uint64_t* counts, void* stream);

/*
 * Perspective camera (camera_type = 1): CudaRasterizer::Rasterizer::forward / backward /
 * markVisible (cuda_rasterizer/rasterizer.h:39-92, rasterizer_impl.cu:170-183,250-530) with preprocessCUDA
 * (forward.cu:232-340), computeCov2D (:86-128), renderDepthCUDA (:472-590) and the backward kernels
 * computeCov2DCUDA / preprocessCUDA (backward.cu:156-292, :558-608).  Same conventions and buffers as the lonlat
 * calls; projmatrix is the full (projection x view) transform stored column-major, tan_fovx/tan_fovy are
 * tan(fov/2).  Stage 2 is ogs_lonlat_forward_stage2 (binning and blending do not depend on the camera).
 * render_depth != 0 blends the camera-space depth into all three channels instead of the colour; the reference's
 * backward has no depth path, so ogs_pinhole_backward after a render_depth forward is undefined.
 */
OGS_API int ogs_pinhole_forward_stage1(
	int P, int D, int M, int W, int H,
	const float* means3D, const float* shs, const float* colors_precomp, const float* opacities,
	const float* scales, float scale_modifier, const float* rotations, const float* cov3D_precomp,
	const float* viewmatrix, const float* projmatrix, const float* campos, float tan_fovx, float tan_fovy,
	int render_depth, int* radii, char* geom_buffer, char* img_buffer, int64_t* num_rendered_host, void* stream);
OGS_API int ogs_pinhole_backward(
	int P, int D, int M, int64_t num_rendered, int W, int H,
	const float* background,
	const float* means3D, const float* shs, const float* colors_precomp,
	const float* scales, float scale_modifier, const float* rotations, const float* cov3D_precomp,
	const float* viewmatrix, const float* projmatrix, const float* campos, float tan_fovx, float tan_fovy,
	const int* radii, char* geom_buffer, char* binning_buffer, char* img_buffer,
	const float* dL_dpix,
	float* dL_dmean2D, float* dL_dconic, float* dL_dopacity, float* dL_dcolor,
	float* dL_dmean3D, float* dL_dcov3D, float* dL_dsh, float* dL_dscale, float* dL_drot,
	void* stream);
/* checkFrustum (rasterizer_impl.cu:64-77): present[i] = camera-space z > 0.2 */
OGS_API int ogs_mark_visible_pinhole(int P, const float* means3D, const float* viewmatrix, const float* projmatrix,
                                     uint8_t* present, void* stream);

/*
 * Raw-parameter entry points.  Before every render the reference turns the model's
 * stored tensors into the rasterizer's arguments with separate LibTorch ops — opacity = sigmoid(opacity_),
 * scales = exp(scaling_), rotations = normalize(rotation_), shs = cat(features_dc_, features_rest_)
 * (src/gaussian_model.cpp:54-77, called from src/gaussian_renderer.cpp:212-258) — and autograd walks the
 * same ops backwards.  These calls take the stored tensors themselves:
 *   xyz [P,3], features_dc [P,1,3], features_rest [P,M-1,3], opacity_raw [P,1] (logit),
 *   scaling_raw [P,3] (log), rotation_raw [P,4] (unnormalised w,x,y,z)
 * and return gradients with respect to them (what the optimiser's six parameter groups receive,
 * gaussian_model.cpp:485-511).  Stage 2 is ogs_lonlat_forward_stage2, unchanged.  dL_dmean2D [P,3]
 * (optional, NULL to skip) is the screen-space gradient the densification statistics use.
 * Every output element is written.
 */
OGS_API int ogs_lonlat_forward_raw_stage1(
	int P, int D, int M, int W, int H,
	const float* xyz, const float* features_dc, const float* features_rest, const float* opacity_raw,
	const float* scaling_raw, float scale_modifier, const float* rotation_raw,
	const float* viewmatrix, const float* campos,
	int* radii, char* geom_buffer, char* img_buffer, int64_t* num_rendered_host, void* stream);
OGS_API int ogs_lonlat_backward_raw(
	int P, int D, int M, int64_t num_rendered, int W, int H, const float* background,
	const float* xyz, const float* features_dc, const float* features_rest,
	const float* scaling_raw, float scale_modifier, const float* rotation_raw,
	const float* viewmatrix, const float* campos, const int* radii,
	char* geom_buffer, char* binning_buffer, char* img_buffer, const float* dL_dpix,
	float* dL_dmean2D, float* dL_dxyz, float* dL_dfeatures_dc, float* dL_dfeatures_rest,
	float* dL_dopacity_raw, float* dL_dscaling_raw, float* dL_drotation_raw, void* stream);

/*
 * The training iteration either side of the rasterizer (SURVEY.md §8 f-3;
 * GaussianMapper::trainForOneIteration, src/gaussian_mapper.cpp:391-434).
 *
 * ogs_photometric_loss: loss = (1 - lambda) * L1 + lambda * (1 - SSIM) between rendered * mask and gt
 * (gaussian_mapper.cpp:391-413; l1_loss and ssim of include/loss_utils.h:31-34,58-131: 11x11 Gaussian
 * window, sigma 1.5, zero padding, mean over all elements), forward AND backward: writes
 * loss_out[3] = {loss, L1, SSIM} (device) and dL_dpix [3,H,W] = dloss/drendered.  mask is NULL, [1,H,W]
 * (mask_channels 1) or [3,H,W] (3).  rows_used < H drops the bottom rows from the loss as
 * skip_bottom_ratio does (gaussian_mapper.cpp:395-407); their dL_dpix is 0.  With lambda == 0 the SSIM term has no
 * weight and is not evaluated (one streaming L1 pass; loss_out[2] = 0).  workspace: 8-byte aligned,
 * ogs_photometric_loss_workspace_bytes(W, H) bytes.
 *
 * ogs_adam_step: one launch of torch::optim::Adam::step over up to 8 parameter groups (the reference's
 * six: xyz, f_dc, f_rest, opacity, scaling, rotation with their own learning rates,
 * gaussian_model.cpp:485-511; eps 1e-15, no weight decay, no amsgrad).  params/grads/exp_avg/exp_avg_sq
 * are HOST arrays of device pointers, counts[g] elements each; `step` counts from 1; betas and eps are
 * doubles as in torch::optim::AdamOptions (1 - beta and the bias corrections are evaluated in double).
 *
 * ogs_densify_stats: for every Gaussian with radii > 0: max_radii2D = max(max_radii2D, radii),
 * xyz_gradient_accum += ||dL_dmean2D.xy||, denom += 1 (gaussian_mapper.cpp:427-434,
 * gaussian_model.cpp:839-853).
 */
OGS_API size_t ogs_photometric_loss_workspace_bytes(int W, int H);
OGS_API int ogs_photometric_loss(
	int W, int H, int rows_used, float lambda_dssim, const float* rendered, const float* gt,
	const float* mask, int mask_channels, char* workspace, float* loss_out, float* dL_dpix, void* stream);
OGS_API int ogs_adam_step(
	int groups, float* const* params, const float* const* grads, float* const* exp_avg, float* const* exp_avg_sq,
	const size_t* counts, const float* lrs, int64_t step, double beta1, double beta2, double eps, void* stream);
OGS_API int ogs_densify_stats(
	int P, const int* radii, const float* dL_dmean2D, float* max_radii2D, float* xyz_gradient_accum, float* denom,
	void* stream);

/*
 * Data-parallel gradient exchange over NVLink peer memory (the reference has no multi-GPU code).
 * bufs[r] (host array, `world` <= 8 entries) is rank r's bucket (count_sum + count_max floats, both multiples of 4) as
 * mapped into THIS process (CUDA IPC / symmetric memory), bufs[rank] the local one.  The call reduces slice `rank` of all
 * buckets in rank order and stores the result into every bucket (reduce-scatter + all-gather in one kernel, peer loads
 * and stores): the first count_sum floats are summed, the count_max floats behind them (non-negative values such as radii,
 * which order like their bit patterns) are max-reduced — one call replaces the all-reduce(SUM) plus the separate
 * all-reduce(MAX) of the densification statistics.  The caller orders it against the other ranks: all buckets written
 * before, all slices stored after (a cross-rank barrier on the same stream on both sides; omnigs-fork_b200/parallel.py
 * uses torch.distributed's symmetric memory).
 * ogs_multimem_allreduce: the same contract through the bucket's NVLink multicast address (NVSwitch in-switch reduction:
 * multimem.ld_reduce / multimem.st on slice `rank`); `multicast` is the multicast mapping of the symmetric buffer.
 */
OGS_API int ogs_peer_allreduce(float* const* bufs, int world, int rank, size_t count_sum, size_t count_max, void* stream);
OGS_API int ogs_multimem_allreduce(float* multicast, int world, int rank, size_t count_sum, size_t count_max, void* stream);
/* Latitude bands: store pixel rows [y0, y1) of the three planes of `src` ([3,H,W], this rank's
 * render) into every rank's image images[r] ([3,H,W] in symmetric memory): an all-gather of band rows by peer stores. */
OGS_API int ogs_band_rows_allgather(float* const* images, int world, int rank, const float* src, int W, int H,
                                    int y0, int y1, void* stream);
/* One view's increments of the densification statistics as plain [P] arrays (gaussian_mapper.cpp:427-434,
 * gaussian_model.cpp:839-853): ||dL_dmean2D.xy|| and 1 where radii > 0 (else 0), and the radius as float — what
 * data-parallel ranks sum / sum / max over views before applying them. */
OGS_API int ogs_view_stats(int P, const int* radii, const float* dL_dmean2D, float* grad_norm, float* visible,
                           float* radius, void* stream);

/*
 * Longitude-seam wrap-around (opt-in extension; off by default = reference parity).
 * The reference's live code clamps tile rects at the left/right image edge (auxiliary.h:56-66,
 * forward.cu:678-681), so a Gaussian straddling lon = +-pi is cut at the seam; its dead code
 * (getRectCyclic, auxiliary.h:68-83; renderLonlat, forward.h:118-129) sketches the fix implemented
 * here: the x tile range is taken modulo the tile grid and every tile uses the copy of the mean
 * (x, x - W or x + W) nearest to it.  Requires W % 16 == 0.  Process-global; also enabled by the
 * environment variable OMNIGS_B200_SEAM_WRAP=1.  The mode a forward ran in is recorded in its geometry
 * buffer, so the matching backward follows it regardless of later changes.
 */
OGS_API int ogs_set_seam_wrap(int on);
OGS_API int ogs_get_seam_wrap(void);

/*
 * Per-stage device timing (CUDA events on the launching stream), for bench.py's roofline block.
 * Off by default; when enabled on the calling thread every stage of the next forward/backward is
 * bracketed by events.  ogs_profile_read synchronises those events and writes milliseconds per stage
 * (0 for stages that did not run); returns the number of stages (OGS_PROF_COUNT).
 */
#define OGS_PROF_PREPROCESS_FWD 0
#define OGS_PROF_DEPTH_ORDER 1    /* depth-key histogram + 4 onesweep passes + emit-offset scan */
#define OGS_PROF_TILE_RANGES 2
#define OGS_PROF_EMIT 3           /* duplicateWithKeys equivalent */
#define OGS_PROF_TILE_SORT 4      /* onesweep passes over the tile ids */
#define OGS_PROF_RENDER_FWD 5
#define OGS_PROF_RENDER_BWD 6
#define OGS_PROF_PREPROCESS_BWD 7
#define OGS_PROF_COUNT 8
OGS_API int ogs_profile_enable(int on);
OGS_API int ogs_profile_read(float* ms, int count);

#ifdef __cplusplus
}
#endif
#endif /* OMNIGS_B200_H */
