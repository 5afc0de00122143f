/*
 * lonlat_oracle.h — CPU restatement (plain C, float32) of the reference's equirectangular
 * ("lonlat", camera_type = 3) rasterizer path.
 *
 * TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` fallback leg may load this.  The product (omnigs-fork_b200/) never does.
 *
 * Every function cites the reference file:line it restates (paths relative to the
 * reference root, raikuma/OmniGS-fork).  Parity pinning: tests/golden/ holds outputs of the
 * UNMODIFIED reference rasterizer (oracle/_ref build, run on a B200 by
 * tests/golden/make_golden.py); tests/test_oracle_golden.py checks this restatement against them.
 *
 * All arrays are caller-allocated, row-major, float32/int32/uint32 unless stated.
 */
#ifndef OGS_LONLAT_ORACLE_H
#define OGS_LONLAT_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/*
 * Seam wrap-around switch (process-global, default 0 = the reference's live behaviour).  With 1 the
 * oracle follows the NON-parity extension of omnigs_b200.h (ogs_set_seam_wrap): x tile range modulo the
 * grid as sketched by the reference's unused getRectCyclic (auxiliary.h:68-83), each tile blends the
 * nearest copy (x, x-W, x+W) of a Gaussian.  No reference output exists for this mode; it is checked by
 * the yaw-invariance property (tests/test_seam_wrap.py).
 */
void ogs_oracle_set_seam_wrap(int on);

/* threads the OpenMP build will use (1 when built without OpenMP) */
void ogs_oracle_set_num_threads(int n);
int ogs_oracle_num_threads(void);

/* cuda_rasterizer/rasterizer_impl.cu:47-62 (getHigherMsb) */
uint32_t ogs_oracle_higher_msb(uint32_t n);

/*
 * Per-Gaussian forward preprocessing — cuda_rasterizer/forward.cu:593-703 (preprocessLonlatCUDA)
 * with auxiliary.h:51-66,85-93,198-220,236-248 and forward.cu:30-83,130-228.
 * NULL for an absent optional input (shs | colors_precomp; scales+rotations | cov3D_precomp).
 * Outputs for culled Gaussians: radii = 0, tiles_touched = 0, everything else untouched
 * (callers pre-fill; the reference leaves them uninitialised).
 * Also fills point_offsets = inclusive prefix sum of tiles_touched (rasterizer_impl.cu:622).
 * Returns num_rendered (= point_offsets[P-1]; 0 when P == 0).
 */
int64_t ogs_oracle_preprocess_fwd(
	int P, int D, int M,
	const float* means3D, const float* scales, float scale_modifier, const float* rotations,
	const float* opacities, const float* shs, const float* cov3D_precomp, const float* colors_precomp,
	const float* viewmatrix, const float* campos, int W, int H,
	int32_t* radii, float* means2D /*P*2*/, float* depths, float* cov3D /*P*6*/, float* rgb /*P*3*/,
	float* conic_opacity /*P*4*/, uint32_t* tiles_touched, uint32_t* point_offsets, uint8_t* clamped /*P*3*/);

/*
 * Binning — rasterizer_impl.cu:94-140 (duplicateWithKeys), :651-661 (stable radix sort on the low
 * 32+bit key bits), :664 + :145-167 (memset + identifyTileRanges).
 * R-sized outputs; ranges is T*2 (x = first, y = one past last; empty tiles stay (0,0)).
 */
void ogs_oracle_bin(
	int P, int W, int H,
	const float* means2D, const float* depths, const int32_t* radii, const uint32_t* point_offsets,
	int64_t R, uint64_t* keys_unsorted, uint32_t* values_unsorted,
	uint64_t* keys_sorted, uint32_t* point_list, uint32_t* ranges);

/* Tile alpha-blend forward — forward.cu:346-467 (renderCUDA).  out_color is planar [3,H,W]. */
void ogs_oracle_render_fwd(
	int W, int H, const uint32_t* ranges, const uint32_t* point_list,
	const float* means2D, const float* colors /*P*3*/, const float* conic_opacity,
	const float* background, float* final_T /*H*W*/, uint32_t* n_contrib /*H*W*/, float* out_color);

/*
 * Tile alpha-blend backward — backward.cu:672-843 (renderCUDA).  The four gradient arrays are
 * ACCUMULATED into (caller zero-fills, as rasterize_points.cu:200-208 does).
 * dL_dmean2D is P*3 (x,y used), dL_dconic is P*4 (.x,.y,.w used).
 */
void ogs_oracle_render_bwd(
	int W, int H, const uint32_t* ranges, const uint32_t* point_list, const float* background,
	const float* means2D, const float* conic_opacity, const float* colors,
	const float* final_T, const uint32_t* n_contrib, const float* dL_dpixels /*[3,H,W]*/,
	float* dL_dmean2D, float* dL_dconic, float* dL_dopacity, float* dL_dcolors);

/*
 * Per-Gaussian backward — backward.cu:297-485 (computeCov2DLonLatCUDA) followed by
 * backward.cu:613-669 (preprocessLonLatCUDA) with :30-151 (SH bwd), :489-552 (cov3D bwd).
 * cov3D is the array the forward used (computed or precomputed).  Outputs must be zero-filled
 * by the caller (rasterize_points.cu:200-208,246-247); culled rows stay zero.
 */
void ogs_oracle_preprocess_bwd(
	int P, int D, int M,
	const float* means3D, const int32_t* radii, const float* shs, const uint8_t* clamped,
	const float* scales, const float* rotations, float scale_modifier, const float* cov3D,
	const float* viewmatrix, int W, int H, const float* campos,
	const float* dL_dmean2D /*P*3*/, const float* dL_dconic /*P*4*/, float* dL_dcolor /*P*3 (read)*/,
	float* dL_dmeans3D /*P*3*/, float* dL_dcov3D /*P*6*/, float* dL_dsh /*P*M*3*/,
	float* dL_dscale /*P*3*/, float* dL_drot /*P*4*/, float* dpx_dt /*P*3*/, float* dpy_dt /*P*3*/);

/*
 * Perspective camera (camera_type 1).  Binning and blending are the functions above (the
 * reference shares those kernels between cameras); these restate the two per-Gaussian steps:
 * forward.cu:232-340 (preprocessCUDA, with in_frustum auxiliary.h:166-196 and computeCov2D forward.cu:86-128;
 * render_depth: what renderDepthCUDA blends, forward.cu:472-590), backward.cu:156-292 + :558-608, and
 * checkFrustum (rasterizer_impl.cu:64-77).  projmatrix = full transform, column-major; tan_fov = tan(fov/2).
 */
int64_t ogs_oracle_pinhole_preprocess_fwd(
	int P, int D, int M,
	const float* means3D, const float* scales, float scale_modifier, const float* rotations,
	const float* opacities, const float* shs, const float* cov3D_precomp, const float* colors_precomp,
	const float* viewmatrix, const float* projmatrix, const float* campos, int W, int H,
	float tan_fovx, float tan_fovy, int render_depth,
	int32_t* radii, float* means2D, float* depths, float* cov3D, float* rgb,
	float* conic_opacity, uint32_t* tiles_touched, uint32_t* point_offsets, uint8_t* clamped);
void ogs_oracle_pinhole_preprocess_bwd(
	int P, int D, int M,
	const float* means3D, const int32_t* radii, const float* shs, const uint8_t* clamped,
	const float* scales, const float* rotations, float scale_modifier, const float* cov3D,
	const float* viewmatrix, const float* projmatrix, int W, int H, float tan_fovx, float tan_fovy, const float* campos,
	const float* dL_dmean2D, const float* dL_dconic, float* dL_dcolor,
	float* dL_dmeans3D, float* dL_dcov3D, float* dL_dsh, float* dL_dscale, float* dL_drot);
void ogs_oracle_pinhole_mark_visible(int P, const float* means3D, const float* viewmatrix, uint8_t* present);

#ifdef __cplusplus
}
#endif
#endif
