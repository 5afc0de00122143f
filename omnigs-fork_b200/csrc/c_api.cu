// c_api.cu — the C ABI of libomnigs_b200.so (include/omnigs_b200.h): buffer layouts,
// orchestration of the forward (two stages) and backward, test-only exports.
//
// Orchestration mirrors CudaRasterizer::LonlatRasterizer::{forward,backward,markVisible}
// (reference cuda_rasterizer/rasterizer_impl.cu:540-697, :701-795, :185-192).
#include "ogs_common.cuh"
#include "launchers.cuh"
#include <cmath>
#include <algorithm>

#include <cstdlib>
#include <mutex>
#include <string>

namespace ogs {

// ------------------------------------------------------------------ errors
static thread_local std::string g_last_error;

int fail(int code, const char* msg)
{
	g_last_error = msg ? msg : "";
	return code;
}
int fail_cuda(cudaError_t e)
{
	g_last_error = std::string("CUDA error: ") + cudaGetErrorName(e) + ": " + cudaGetErrorString(e);
	return OGS_ERR_CUDA;
}

// ------------------------------------------------------------------ options
namespace {
int env_seam_wrap()
{
	const char* e = getenv("OMNIGS_B200_SEAM_WRAP");
	return (e && e[0] && e[0] != '0') ? 1 : 0;
}
int g_seam_wrap = env_seam_wrap();
} // namespace

// ------------------------------------------------------------------ per-stage profiling
namespace {
struct Profiler {
	bool on = false;
	bool created = false;
	cudaEvent_t ev[2 * OGS_PROF_COUNT];
	bool used[OGS_PROF_COUNT] = {};
};
thread_local Profiler g_prof;
} // namespace

void prof_begin(int stage, cudaStream_t st)
{
	if (!g_prof.on) return;
	g_prof.used[stage] = true;
	cudaEventRecord(g_prof.ev[2 * stage], st);
}
void prof_end(int stage, cudaStream_t st)
{
	if (!g_prof.on) return;
	cudaEventRecord(g_prof.ev[2 * stage + 1], st);
}

// ------------------------------------------------------------------ layouts
namespace {
constexpr size_t kAlign = 256;

struct Carver {
	char* base;
	size_t off = 0;
	explicit Carver(char* b) : base(b) {}
	template <typename T>
	T* take(size_t count)
	{
		off = align_up(off, kAlign);
		T* p = reinterpret_cast<T*>(base + off);
		off += count * sizeof(T);
		return p;
	}
};

int sort_tiles(int64_t n) { return (int)((n + kSortItemsPerBlock - 1) / kSortItemsPerBlock); }

GeomState carve_geom(char* base, int P, size_t* total)
{
	Carver c(base);
	GeomState g{};
	const size_t p = (size_t)P;
	g.g0 = c.take<float4>(p);
	g.g1 = c.take<float4>(p);
	g.gb = c.take<float2>(p);
	g.depth = c.take<float>(p);
	g.rect = c.take<uint2>(p);
	g.tiles_touched = c.take<uint32_t>(p);
	g.cov3D = c.take<float>(6 * p);
	g.clamped = c.take<uint8_t>(p);
	g.sort_key[0] = c.take<uint32_t>(p);
	g.sort_key[1] = c.take<uint32_t>(p);
	g.sort_val[0] = c.take<uint32_t>(p);
	g.sort_val[1] = c.take<uint32_t>(p);
	g.emit_offset = c.take<uint32_t>(p + 1);
	g.grad_acc = c.take<float>(12 * p);
	// ---- zero-filled every forward ----
	const int tiles = sort_tiles(P);
	g.depth_hist = c.take<uint32_t>(4 * 256);
	g.zero_begin = reinterpret_cast<char*>(g.depth_hist);
	g.depth_status = c.take<uint32_t>((size_t)4 * tiles * 256);
	g.scan_status = c.take<unsigned long long>((size_t)tiles + 2);   // [tiles + 1] holds the scan's ticket counter
	g.scalars = c.take<unsigned long long>(8);
	g.scalars_bytes = 64;
	c.off = align_up(c.off, kAlign);
	g.zero_bytes = (size_t)((base + c.off) - g.zero_begin);
	if (total) *total = c.off + kAlign;
	return g;
}

ImageState carve_img(char* base, int W, int H, size_t* total)
{
	Carver c(base);
	ImageState s{};
	const int gx = ceil_div(W, kTile), gy = ceil_div(H, kTile);
	const size_t N = (size_t)W * H, T = (size_t)gx * gy;
	s.final_T = c.take<float>(N);
	s.n_contrib = c.take<uint32_t>(N);
	s.ranges = c.take<uint2>(T);
	s.tile_diff = c.take<int>((size_t)(gx + 1) * (gy + 1));
	s.tile_count = c.take<uint32_t>(T);
	s.tile_hist = c.take<uint32_t>((size_t)kMaxTilePasses * kMaxBins);
	if (total) *total = align_up(c.off, kAlign) + kAlign;
	return s;
}

BinningState carve_binning(char* base, int64_t R, int W, int H, size_t* total)
{
	Carver c(base);
	BinningState b{};
	const size_t r = (size_t)(R > 0 ? R : 0);
	const TileSortPlan plan = make_tile_sort_plan(W, H);
	b.key[0] = c.take<uint32_t>(r + 4);    // + 4: the segment path keeps S + 1 <= R + 1 emission offsets in key[1]
	b.key[1] = c.take<uint32_t>(r + 4);
	b.val[0] = c.take<uint32_t>(r);
	b.val[1] = c.take<uint32_t>(r);
	b.point_list = b.val[plan.passes & 1];
	b.first_src = c.take<uint32_t>(r / 2048 + 3);
	size_t status_words = 0;
	const int tiles = sort_tiles(R);
	for (int p = 0; p < plan.passes; p++) status_words += (size_t)tiles << plan.bits[p];
	if (plan.segments) status_words = std::max(status_words, ((size_t)tiles << plan.bits_x) + ((size_t)tiles << plan.bits_y));
	b.status = c.take<uint32_t>(status_words);
	b.zero_begin = reinterpret_cast<char*>(b.status);
	b.tickets = c.take<unsigned int>(kMaxTilePasses);
	b.seg_scan_status = c.take<unsigned long long>((size_t)tiles + 2);
	b.col_diff = c.take<int>(kMaxBins + 1);
	c.off = align_up(c.off, kAlign);
	b.zero_bytes = (size_t)((base + c.off) - b.zero_begin);
	if (total) *total = c.off + kAlign;
	return b;
}

char* aligned_base(char* p)
{
	return reinterpret_cast<char*>(align_up(reinterpret_cast<size_t>(p), kAlign));
}
} // namespace

size_t GeomState::bytes(int P) { size_t t; carve_geom(nullptr, P, &t); return t; }
GeomState GeomState::carve(char* base, int P) { return carve_geom(aligned_base(base), P, nullptr); }
size_t ImageState::bytes(int W, int H) { size_t t; carve_img(nullptr, W, H, &t); return t; }
ImageState ImageState::carve(char* base, int W, int H) { return carve_img(aligned_base(base), W, H, nullptr); }
size_t BinningState::bytes(int64_t R, int W, int H) { size_t t; carve_binning(nullptr, R, W, H, &t); return t; }
BinningState BinningState::carve(char* base, int64_t R, int W, int H) { return carve_binning(aligned_base(base), R, W, H, nullptr); }

// ------------------------------------------------------------------ num_rendered read-back slot
namespace {
struct Readback {
	unsigned long long* pinned = nullptr;
	cudaEvent_t event = nullptr;
	// tile_ranges (one CTA, latency bound) depends only on the per-Gaussian forward, like the depth order: it runs on
	// this side stream beside the depth-order passes (fork / join with the two events)
	cudaStream_t side = nullptr;
	cudaEvent_t fork = nullptr, join = nullptr;
};
constexpr int kMaxDevices = 64;
std::mutex g_rb_mutex;
thread_local Readback g_rb[kMaxDevices];

int get_readback(Readback** out)
{
	int dev = 0;
	OGS_CUDA_TRY(cudaGetDevice(&dev));
	if (dev < 0 || dev >= kMaxDevices) return fail(OGS_ERR_NO_DEVICE, "device ordinal out of range");
	Readback& rb = g_rb[dev];
	if (!rb.pinned) {
		// all-or-nothing: a slot is published only when every resource exists (a failed call is retried from scratch)
		std::lock_guard<std::mutex> lock(g_rb_mutex);
		Readback fresh;
		cudaError_t e = cudaMallocHost(reinterpret_cast<void**>(&fresh.pinned), 64);
		if (e == cudaSuccess) e = cudaEventCreateWithFlags(&fresh.event, cudaEventDisableTiming);
		if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&fresh.side, cudaStreamNonBlocking);
		if (e == cudaSuccess) e = cudaEventCreateWithFlags(&fresh.fork, cudaEventDisableTiming);
		if (e == cudaSuccess) e = cudaEventCreateWithFlags(&fresh.join, cudaEventDisableTiming);
		if (e != cudaSuccess) {
			if (fresh.join) cudaEventDestroy(fresh.join);
			if (fresh.fork) cudaEventDestroy(fresh.fork);
			if (fresh.side) cudaStreamDestroy(fresh.side);
			if (fresh.event) cudaEventDestroy(fresh.event);
			if (fresh.pinned) cudaFreeHost(fresh.pinned);
			return fail_cuda(e);
		}
		rb = fresh;
	}
	*out = &rb;
	return OGS_OK;
}

__global__ void export_geometry_kernel(int P, const float4* g0, const float4* g1, const float2* gb, const float* depth,
                                       const uint32_t* tiles, const uint8_t* clamped, const float* cov3D_in,
                                       float* means2D, float* depths, float* conic_opacity, float* rgb,
                                       uint32_t* tiles_touched, uint8_t* clamped_out, float* cov3D)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= P) return;
	const bool vis = tiles[i] > 0;
	float4 a = vis ? g0[i] : make_float4(0, 0, 0, 0);
	float4 b = vis ? g1[i] : make_float4(0, 0, 0, 0);
	float cb = vis ? gb[i].x : 0.f;
	if (means2D) { means2D[2 * i] = a.x; means2D[2 * i + 1] = a.y; }
	if (depths) depths[i] = vis ? depth[i] : 0.f;
	if (conic_opacity) { conic_opacity[4 * i] = a.z; conic_opacity[4 * i + 1] = a.w; conic_opacity[4 * i + 2] = b.x; conic_opacity[4 * i + 3] = b.y; }
	if (rgb) { rgb[3 * i] = b.z; rgb[3 * i + 1] = b.w; rgb[3 * i + 2] = cb; }
	if (tiles_touched) tiles_touched[i] = tiles[i];
	if (clamped_out) {
		unsigned m = vis ? clamped[i] : 0u;
		clamped_out[3 * i] = m & 1u; clamped_out[3 * i + 1] = (m >> 1) & 1u; clamped_out[3 * i + 2] = (m >> 2) & 1u;
	}
	if (cov3D)
		for (int k = 0; k < 6; k++) cov3D[6 * (size_t)i + k] = vis ? cov3D_in[6 * (size_t)i + k] : 0.f;
}

int check_image(int W, int H)
{
	if (W <= 0 || H <= 0) return fail(OGS_ERR_INVALID_ARG, "image size must be positive");
	if (ceil_div(W, kTile) > 65535 || ceil_div(H, kTile) > 65535)
		return fail(OGS_ERR_TOO_MANY, "more than 65535 tiles along an axis");
	if ((int64_t)ceil_div(W, kTile) * ceil_div(H, kTile) >= (1ll << 30))
		return fail(OGS_ERR_TOO_MANY, "too many tiles");
	return OGS_OK;
}

struct PinholeParams {   // camera_type 1: full projection, tan(fov/2), depth instead of colour
	const float* projmatrix;
	float tan_fovx, tan_fovy;
	int render_depth;
};

int forward_stage1_impl(
	int P, int D, int M, int W, int H, int band_y0, int band_y1,
	const float* means3D, const float* shs, const float* colors_precomp, const float* opacities,
	const float* scales, float scale_modifier, const float* rotations, const float* cov3D_precomp,
	const float* viewmatrix, const float* campos,
	int* radii, char* geom_buffer, char* img_buffer, int64_t* num_rendered_host, cudaStream_t st,
	const float* features_dc = nullptr, const float* features_rest = nullptr, const PinholeParams* pin = nullptr,
	bool defer_colors = false)
{
	// raw-parameter mode: features_dc / features_rest given instead of shs; opacities, scales and rotations
	// are then the stored (pre-activation) tensors
	const bool raw = features_dc != nullptr;
	if (P < 0 || !num_rendered_host) return fail(OGS_ERR_INVALID_ARG, "bad P / num_rendered_host");
	if (int rc = check_image(W, H)) return rc;
	*num_rendered_host = 0;
	if (!img_buffer) return fail(OGS_ERR_INVALID_ARG, "img_buffer is NULL");
	const int gx = ceil_div(W, kTile), gy = ceil_div(H, kTile);
	ImageState img = ImageState::carve(img_buffer, W, H);
	OGS_CUDA_TRY(cudaMemsetAsync(img.tile_diff, 0, sizeof(int) * (size_t)(gx + 1) * (gy + 1), st));
	if (P == 0) return launch_tile_ranges(img, W, H, st);

	if (!means3D || !opacities || !viewmatrix || !campos || !radii || !geom_buffer)
		return fail(OGS_ERR_INVALID_ARG, "a required pointer is NULL");
	if (raw) {
		// SH degree 0 models have an empty features_rest_ (M == 1): allowed, as the reference's cat() path allows it
		if (shs || colors_precomp || cov3D_precomp || (M > 1 && !features_rest) || !scales || !rotations || M < 1)
			return fail(OGS_ERR_INVALID_ARG, "raw-parameter mode takes features_dc, features_rest, scaling and rotation only");
	} else if (defer_colors) {
		if (shs || colors_precomp) return fail(OGS_ERR_INVALID_ARG, "the geometry-only stage takes neither shs nor colors_precomp");
	} else if ((shs == nullptr) == (colors_precomp == nullptr))
		return fail(OGS_ERR_INVALID_ARG, "exactly one of shs / colors_precomp must be given");
	if (((scales == nullptr) || (rotations == nullptr)) == (cov3D_precomp == nullptr))
		return fail(OGS_ERR_INVALID_ARG, "exactly one of (scales, rotations) / cov3D_precomp must be given");
	if ((shs || raw) && (M <= 0 || (D + 1) * (D + 1) > M || D < 0 || D > 3))
		return fail(OGS_ERR_INVALID_ARG, "SH degree / coefficient count mismatch");
	band_y0 = max(0, band_y0);
	band_y1 = min(gy, band_y1);
	const int seam_wrap = g_seam_wrap;
	if (seam_wrap && (W % kTile != 0 || 2 * gx > 65535))
		return fail(OGS_ERR_INVALID_ARG, "seam wrap-around needs an image width that is a multiple of 16");

	GeomState g = GeomState::carve(geom_buffer, P);
	OGS_CUDA_TRY(cudaMemsetAsync(g.zero_begin, 0, g.zero_bytes, st));

	PreprocessFwdArgs a{};
	a.P = P; a.D = D; a.M = M; a.W = W; a.H = H; a.gx = gx; a.gy = gy; a.band_y0 = band_y0; a.band_y1 = band_y1;
	a.seam_wrap = seam_wrap;
	a.scale_modifier = scale_modifier;
	a.means3D = means3D; a.shs = shs; a.colors_precomp = colors_precomp; a.opacities = opacities;
	a.scales = scales; a.rotations = rotations; a.cov3D_precomp = cov3D_precomp;
	a.raw = raw ? 1 : 0; a.features_dc = features_dc; a.features_rest = features_rest;
	a.defer_colors = defer_colors ? 1 : 0;
	if (pin) {
		if (!pin->projmatrix) return fail(OGS_ERR_INVALID_ARG, "projmatrix is NULL");
		if (raw || seam_wrap) return fail(OGS_ERR_INVALID_ARG, "raw-parameter mode and seam wrap-around are lonlat-only");
		a.pinhole = 1; a.render_depth = pin->render_depth; a.projmatrix = pin->projmatrix;
		a.tan_fovx = pin->tan_fovx; a.tan_fovy = pin->tan_fovy;
		// rasterizer_impl.cu:274-276: tan(fov/2) -> focal length in pixels, float arithmetic
		a.focal_y = H / (2.0f * pin->tan_fovy);
		a.focal_x = W / (2.0f * pin->tan_fovx);
	}
	a.viewmatrix = viewmatrix; a.campos = campos; a.radii = radii;
	a.g0 = g.g0; a.g1 = g.g1; a.gb = g.gb; a.depth = g.depth; a.rect = g.rect;
	a.tiles_touched = g.tiles_touched; a.cov3D = g.cov3D; a.clamped = g.clamped;
	a.sort_key = g.sort_key[0]; a.tile_diff = img.tile_diff; a.total_tiles = g.scalars;
	prof_begin(OGS_PROF_PREPROCESS_FWD, st);
	if (int rc = launch_preprocess_fwd(a, st)) return rc;
	prof_end(OGS_PROF_PREPROCESS_FWD, st);

	// num_rendered read-back (the reference blocks here too, rasterizer_impl.cu:627-628); the
	// R-independent kernels are queued behind the copy so they overlap the host round trip.
	Readback* rb = nullptr;
	if (int rc = get_readback(&rb)) return rc;
	OGS_CUDA_TRY(cudaMemcpyAsync(rb->pinned, g.scalars, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
	OGS_CUDA_TRY(cudaEventRecord(rb->event, st));
	// OGS_SIDE_STREAM=0 keeps tile_ranges on the caller's stream behind the depth order (A/B knob)
	static const bool side = [] { const char* e = getenv("OGS_SIDE_STREAM"); return e ? atoi(e) != 0 : true; }();
	const cudaStream_t st_ranges = side ? rb->side : st;
	// Work forked onto the side stream must be joined back on EVERY path out of here: the caller may free or reuse the
	// buffers on `st` as soon as this call returns, error or not.
	int rc_side = OGS_OK, rc_main = OGS_OK;
	if (side) {
		OGS_CUDA_TRY(cudaEventRecord(rb->fork, st));
		OGS_CUDA_TRY(cudaStreamWaitEvent(rb->side, rb->fork, 0));
		prof_begin(OGS_PROF_TILE_RANGES, st_ranges);
		rc_side = launch_tile_ranges(img, W, H, st_ranges);
		prof_end(OGS_PROF_TILE_RANGES, st_ranges);
		const cudaError_t e = cudaEventRecord(rb->join, rb->side);
		if (e != cudaSuccess) { cudaStreamSynchronize(rb->side); return fail_cuda(e); }
	}
	prof_begin(OGS_PROF_DEPTH_ORDER, st);
	rc_main = launch_depth_order(g, P, st);
	prof_end(OGS_PROF_DEPTH_ORDER, st);
	if (side) {
		const cudaError_t e = cudaStreamWaitEvent(st, rb->join, 0);
		if (e != cudaSuccess) { cudaStreamSynchronize(rb->side); return fail_cuda(e); }
	} else if (rc_main == OGS_OK) {
		prof_begin(OGS_PROF_TILE_RANGES, st);
		rc_side = launch_tile_ranges(img, W, H, st);
		prof_end(OGS_PROF_TILE_RANGES, st);
	}
	if (rc_main) return rc_main;
	if (rc_side) return rc_side;
	OGS_CUDA_TRY(cudaEventSynchronize(rb->event));
	const unsigned long long total = *rb->pinned;
	if (total >= (1ull << 31)) {   // the reference's `int num_rendered` overflows here (rasterizer_impl.cu:627-632)
		*num_rendered_host = (int64_t)total;
		return fail(OGS_ERR_TOO_MANY, "num_rendered >= 2^31 tile instances");
	}
	*num_rendered_host = (int64_t)total;
	return OGS_OK;
}
} // namespace

} // namespace ogs

using namespace ogs;

extern "C" {

OGS_API int ogs_abi_version(void) { return OGS_ABI_VERSION; }

OGS_API int ogs_set_seam_wrap(int on) { g_seam_wrap = on ? 1 : 0; return OGS_OK; }
OGS_API int ogs_get_seam_wrap(void) { return g_seam_wrap; }

OGS_API int ogs_profile_enable(int on)
{
	if (on && !g_prof.created) {
		for (int i = 0; i < 2 * OGS_PROF_COUNT; i++) OGS_CUDA_TRY(cudaEventCreate(&g_prof.ev[i]));
		g_prof.created = true;
	}
	g_prof.on = on != 0;
	for (int i = 0; i < OGS_PROF_COUNT; i++) g_prof.used[i] = false;
	return OGS_OK;
}

OGS_API int ogs_profile_read(float* ms, int count)
{
	if (!ms || count < OGS_PROF_COUNT) return fail(OGS_ERR_INVALID_ARG, "ogs_profile_read needs OGS_PROF_COUNT floats");
	for (int i = 0; i < OGS_PROF_COUNT; i++) {
		ms[i] = 0.f;
		if (g_prof.created && g_prof.used[i]) {
			OGS_CUDA_TRY(cudaEventSynchronize(g_prof.ev[2 * i + 1]));
			OGS_CUDA_TRY(cudaEventElapsedTime(&ms[i], g_prof.ev[2 * i], g_prof.ev[2 * i + 1]));
		}
		g_prof.used[i] = false;
	}
	return OGS_PROF_COUNT;
}
OGS_API const char* ogs_last_error(void) { return g_last_error.c_str(); }

OGS_API size_t ogs_geom_bytes(int P) { return GeomState::bytes(P < 0 ? 0 : P); }
OGS_API size_t ogs_img_bytes(int W, int H) { return (W <= 0 || H <= 0) ? 0 : ImageState::bytes(W, H); }
OGS_API size_t ogs_binning_bytes(int64_t R, int W, int H) { return (W <= 0 || H <= 0) ? 0 : BinningState::bytes(R, W, H); }

OGS_API int ogs_lonlat_forward_stage1(
	int P, int D, int M, int W, int H,
	const float* means3D, const float* shs, const float* colors_precomp, const float* opacities,
	const float* scales, float scale_modifier, const float* rotations, const float* cov3D_precomp,
	const float* viewmatrix, const float* campos,
	int* radii, char* geom_buffer, char* img_buffer, int64_t* num_rendered_host, void* stream)
{
	return forward_stage1_impl(P, D, M, W, H, 0, 1 << 30, means3D, shs, colors_precomp, opacities, scales,
	                           scale_modifier, rotations, cov3D_precomp, viewmatrix, campos, radii,
	                           geom_buffer, img_buffer, num_rendered_host, (cudaStream_t)stream);
}

OGS_API int ogs_lonlat_forward_stage1_band(
	int P, int D, int M, int W, int H, int band_ty0, int band_ty1,
	const float* means3D, const float* shs, const float* colors_precomp, const float* opacities,
	const float* scales, float scale_modifier, const float* rotations, const float* cov3D_precomp,
	const float* viewmatrix, const float* campos,
	int* radii, char* geom_buffer, char* img_buffer, int64_t* num_rendered_host, void* stream)
{
	if (band_ty0 < 0 || band_ty1 < band_ty0) return fail(OGS_ERR_INVALID_ARG, "bad latitude band");
	return forward_stage1_impl(P, D, M, W, H, band_ty0, band_ty1, means3D, shs, colors_precomp, opacities, scales,
	                           scale_modifier, rotations, cov3D_precomp, viewmatrix, campos, radii,
	                           geom_buffer, img_buffer, num_rendered_host, (cudaStream_t)stream);
}

// Stage 2, first half: emission + tile-id sort (no colours needed).
OGS_API int ogs_lonlat_forward_bin(
	int P, int W, int H, int64_t num_rendered, char* geom_buffer, char* binning_buffer, char* img_buffer, void* stream)
{
	cudaStream_t st = (cudaStream_t)stream;
	if (P < 0 || num_rendered < 0 || !img_buffer) return fail(OGS_ERR_INVALID_ARG, "bad argument to forward bin");
	if (int rc = check_image(W, H)) return rc;
	if (num_rendered >= (1ll << 31)) return fail(OGS_ERR_TOO_MANY, "num_rendered >= 2^31 tile instances");
	if (num_rendered == 0 || P == 0) return OGS_OK;
	if (!geom_buffer || !binning_buffer) return fail(OGS_ERR_INVALID_ARG, "geom_buffer / binning_buffer is NULL");
	ImageState img = ImageState::carve(img_buffer, W, H);
	GeomState g = GeomState::carve(geom_buffer, P);
	BinningState b = BinningState::carve(binning_buffer, num_rendered, W, H);
	OGS_CUDA_TRY(cudaMemsetAsync(b.zero_begin, 0, b.zero_bytes, st));
	return launch_emit_and_tile_sort(g, img, b, P, num_rendered, W, H, st);
}

// Stage 2, second half: the alpha blend.
OGS_API int ogs_lonlat_forward_blend(
	int P, int W, int H, int64_t num_rendered, const float* background,
	char* geom_buffer, char* binning_buffer, char* img_buffer, float* out_color, void* stream)
{
	cudaStream_t st = (cudaStream_t)stream;
	if (P < 0 || num_rendered < 0 || !background || !img_buffer || !out_color)
		return fail(OGS_ERR_INVALID_ARG, "bad argument to forward blend");
	if (int rc = check_image(W, H)) return rc;
	ImageState img = ImageState::carve(img_buffer, W, H);
	GeomState g{};
	BinningState b{};
	if (P > 0) {
		if (!geom_buffer) return fail(OGS_ERR_INVALID_ARG, "geom_buffer is NULL");
		g = GeomState::carve(geom_buffer, P);
	}
	if (num_rendered > 0) {
		if (!binning_buffer) return fail(OGS_ERR_INVALID_ARG, "binning_buffer is NULL");
		b = BinningState::carve(binning_buffer, num_rendered, W, H);
	}
	prof_begin(OGS_PROF_RENDER_FWD, st);
	// the tile-key array of the binning buffer is dead once the list is sorted: it takes the blend's hit bytes
	const int rc = launch_render_fwd(img.ranges, b.point_list, W, H, g.g0, g.g1, g.gb, g.scalars, background,
	                                 img.final_T, img.n_contrib, out_color, reinterpret_cast<uint8_t*>(b.key[0]), st);
	prof_end(OGS_PROF_RENDER_FWD, st);
	return rc;
}

OGS_API int ogs_lonlat_forward_stage2(
	int P, int W, int H, int64_t num_rendered, const float* background,
	char* geom_buffer, char* binning_buffer, char* img_buffer, float* out_color, void* stream)
{
	if (P < 0 || num_rendered < 0 || !background || !img_buffer || !out_color)
		return fail(OGS_ERR_INVALID_ARG, "bad argument to forward stage 2");
	if (int rc = ogs_lonlat_forward_bin(P, W, H, num_rendered, geom_buffer, binning_buffer, img_buffer, stream)) return rc;
	return ogs_lonlat_forward_blend(P, W, H, num_rendered, background, geom_buffer, binning_buffer, img_buffer, out_color, stream);
}

// Geometry-only stage 1 + deferred colours (data-parallel trainer: the SH gradients of the previous step are still being
// exchanged while this step's geometry, depth order and tile sort run; colours are evaluated right before the blend).
OGS_API int ogs_lonlat_forward_stage1_geometry(
	int P, int W, int H,
	const float* means3D, const float* opacities, const float* scales, float scale_modifier, const float* rotations,
	const float* cov3D_precomp, const float* viewmatrix, const float* campos,
	int* radii, char* geom_buffer, char* img_buffer, int64_t* num_rendered_host, void* stream)
{
	return forward_stage1_impl(P, 0, 0, W, H, 0, 1 << 30, means3D, nullptr, nullptr, opacities, scales, scale_modifier,
	                           rotations, cov3D_precomp, viewmatrix, campos, radii, geom_buffer, img_buffer,
	                           num_rendered_host, (cudaStream_t)stream, nullptr, nullptr, nullptr, true);
}

OGS_API int ogs_lonlat_forward_colors(
	int P, int D, int M, const float* means3D, const float* shs, const float* campos, const int* radii,
	char* geom_buffer, void* stream)
{
	if (P < 0) return fail(OGS_ERR_INVALID_ARG, "bad P");
	if (P == 0) return OGS_OK;
	if (!means3D || !shs || !campos || !radii || !geom_buffer) return fail(OGS_ERR_INVALID_ARG, "a required pointer is NULL");
	if (M <= 0 || (D + 1) * (D + 1) > M || D < 0 || D > 3) return fail(OGS_ERR_INVALID_ARG, "SH degree / coefficient count mismatch");
	GeomState g = GeomState::carve(geom_buffer, P);
	return launch_sh_colors(P, D, M, means3D, shs, campos, radii, g.g1, g.gb, g.clamped, (cudaStream_t)stream);
}

// Backward, part 1: zero the packed accumulators and replay the blend (render backward).
OGS_API int ogs_lonlat_backward_render_into(
	int P, int64_t num_rendered, int W, int H, const float* background,
	char* geom_buffer, char* binning_buffer, char* img_buffer, const float* dL_dpix, float* grad_acc, void* stream);

OGS_API int ogs_lonlat_backward_render(
	int P, int64_t num_rendered, int W, int H, const float* background,
	char* geom_buffer, char* binning_buffer, char* img_buffer, const float* dL_dpix, void* stream)
{
	return ogs_lonlat_backward_render_into(P, num_rendered, W, H, background, geom_buffer, binning_buffer, img_buffer,
	                                       dL_dpix, nullptr, stream);
}

// grad_acc: NULL = the accumulators inside geom_buffer; else a caller-owned [P,12] float array (e.g. in symmetric memory,
// so that latitude-band ranks can sum it with ogs_peer_allreduce / ogs_multimem_allreduce)
OGS_API int ogs_lonlat_backward_render_into(
	int P, int64_t num_rendered, int W, int H, const float* background,
	char* geom_buffer, char* binning_buffer, char* img_buffer, const float* dL_dpix, float* grad_acc, void* stream)
{
	cudaStream_t st = (cudaStream_t)stream;
	if (P < 0 || num_rendered < 0) return fail(OGS_ERR_INVALID_ARG, "bad P / num_rendered");
	if (P == 0) return OGS_OK;
	if (int rc = check_image(W, H)) return rc;
	if (!background || !geom_buffer || !img_buffer || !dL_dpix) return fail(OGS_ERR_INVALID_ARG, "a required pointer is NULL");
	if (num_rendered > 0 && !binning_buffer) return fail(OGS_ERR_INVALID_ARG, "binning_buffer is NULL");
	GeomState g = GeomState::carve(geom_buffer, P);
	ImageState img = ImageState::carve(img_buffer, W, H);
	BinningState b{};
	if (num_rendered > 0) b = BinningState::carve(binning_buffer, num_rendered, W, H);
	if (grad_acc && (reinterpret_cast<uintptr_t>(grad_acc) & 15u)) return fail(OGS_ERR_INVALID_ARG, "grad_acc must be 16-byte aligned");
	float* acc = grad_acc ? grad_acc : g.grad_acc;
	OGS_CUDA_TRY(cudaMemsetAsync(acc, 0, sizeof(float) * 12 * (size_t)P, st));
	prof_begin(OGS_PROF_RENDER_BWD, st);
	if (int rc = launch_render_bwd(img.ranges, b.point_list, W, H, background, g.g0, g.g1, g.gb, g.scalars,
	                               img.final_T, img.n_contrib, dL_dpix, acc, reinterpret_cast<const uint8_t*>(b.key[0]), st)) return rc;
	prof_end(OGS_PROF_RENDER_BWD, st);
	return OGS_OK;
}

// Backward, part 2: the fused per-Gaussian backward from the (possibly all-reduced) accumulators.
struct FinishExtras {            // multi-view / data-parallel mode and external accumulators (all optional)
	const float* grad_acc = nullptr;
	int first = 0, count = 0;    // Gaussian sub-range (count 0 = to the end)
	int accumulate = 0;
	float* dL_drgb_view = nullptr;
	float* stat_grad_norm = nullptr;
	float* stat_visible = nullptr;
	float* stat_max_radius = nullptr;
};
static int backward_finish_impl(
	int P, int D, int M, int W, int H,
	const float* means3D, const float* shs, const float* scales, float scale_modifier, const float* rotations,
	const float* cov3D_precomp, const float* viewmatrix, const float* campos, const int* radii, char* geom_buffer,
	float* dL_dmean2D, float* dL_dconic, float* dL_dopacity, float* dL_dcolor,
	float* dL_dmean3D, float* dL_dcov3D, float* dL_dsh, float* dL_dscale, float* dL_drot, const FinishExtras& x, cudaStream_t st)
{
	if (P < 0) return fail(OGS_ERR_INVALID_ARG, "bad P");
	if (P == 0) return OGS_OK;
	if (int rc = check_image(W, H)) return rc;
	if (!means3D || !viewmatrix || !campos || !radii || !geom_buffer || !dL_dopacity || !dL_dmean3D || !dL_dscale || !dL_drot)
		return fail(OGS_ERR_INVALID_ARG, "a required pointer is NULL");
	if (shs && M > 0 && !dL_dsh && !x.dL_drgb_view) return fail(OGS_ERR_INVALID_ARG, "dL_dsh is NULL");
	if (x.accumulate && shs && !x.dL_drgb_view)
		return fail(OGS_ERR_INVALID_ARG, "accumulating views needs dL_drgb_view (dL/dsh is rebuilt by ogs_sh_gradient_from_views)");
	if ((x.stat_grad_norm != nullptr) != (x.stat_visible != nullptr) || (x.stat_grad_norm != nullptr) != (x.stat_max_radius != nullptr))
		return fail(OGS_ERR_INVALID_ARG, "give all three statistics arrays or none");
	if (((scales == nullptr) || (rotations == nullptr)) == (cov3D_precomp == nullptr))
		return fail(OGS_ERR_INVALID_ARG, "exactly one of (scales, rotations) / cov3D_precomp must be given");
	GeomState g = GeomState::carve(geom_buffer, P);
	PreprocessBwdArgs a{};
	a.P = P; a.D = D; a.M = shs ? M : 0; a.W = W; a.H = H; a.scale_modifier = scale_modifier;
	a.means3D = means3D; a.shs = shs; a.scales = scales; a.rotations = rotations;
	a.cov3D = cov3D_precomp ? cov3D_precomp : g.cov3D;
	a.viewmatrix = viewmatrix; a.campos = campos; a.radii = radii; a.clamped = g.clamped;
	a.grad_acc = x.grad_acc ? x.grad_acc : g.grad_acc;
	if (x.first < 0 || x.count < 0 || (x.first % 128) != 0) return fail(OGS_ERR_INVALID_ARG, "Gaussian range: first must be a non-negative multiple of 128");
	a.first_block = x.first / 128;
	a.num_blocks = x.count > 0 ? (x.count + 127) / 128 : 0;
	a.g0 = g.g0; a.g1 = g.g1;
	a.dL_dmean2D = dL_dmean2D; a.dL_dconic = dL_dconic; a.dL_dopacity = dL_dopacity; a.dL_dcolor = dL_dcolor;
	a.dL_dmean3D = dL_dmean3D; a.dL_dcov3D = dL_dcov3D; a.dL_dsh = shs ? dL_dsh : nullptr;
	a.dL_dscale = dL_dscale; a.dL_drot = dL_drot;
	a.accumulate = x.accumulate; a.dL_drgb_view = shs ? x.dL_drgb_view : nullptr;
	a.stat_grad_norm = x.stat_grad_norm; a.stat_visible = x.stat_visible; a.stat_max_radius = x.stat_max_radius;
	prof_begin(OGS_PROF_PREPROCESS_BWD, st);
	const int rc = launch_preprocess_bwd(a, st);
	prof_end(OGS_PROF_PREPROCESS_BWD, st);
	return rc;
}

OGS_API int ogs_lonlat_backward_finish(
	int P, int D, int M, int W, int H,
	const float* means3D, const float* shs, const float* scales, float scale_modifier, const float* rotations,
	const float* cov3D_precomp, const float* viewmatrix, const float* campos, const int* radii, char* geom_buffer,
	float* dL_dmean2D, float* dL_dconic, float* dL_dopacity, float* dL_dcolor,
	float* dL_dmean3D, float* dL_dcov3D, float* dL_dsh, float* dL_dscale, float* dL_drot, void* stream)
{
	if (P > 0 && (!dL_dmean2D || !dL_dcolor || !dL_dcov3D)) return fail(OGS_ERR_INVALID_ARG, "a required pointer is NULL");
	return backward_finish_impl(P, D, M, W, H, means3D, shs, scales, scale_modifier, rotations, cov3D_precomp, viewmatrix,
	                            campos, radii, geom_buffer, dL_dmean2D, dL_dconic, dL_dopacity, dL_dcolor, dL_dmean3D,
	                            dL_dcov3D, dL_dsh, dL_dscale, dL_drot, FinishExtras{}, (cudaStream_t)stream);
}

// ..._finish from caller-owned accumulators (see ogs_lonlat_backward_render_into)
OGS_API int ogs_lonlat_backward_finish_from(
	int P, int D, int M, int W, int H,
	const float* means3D, const float* shs, const float* scales, float scale_modifier, const float* rotations,
	const float* cov3D_precomp, const float* viewmatrix, const float* campos, const int* radii, char* geom_buffer,
	const float* grad_acc,
	float* dL_dmean2D, float* dL_dconic, float* dL_dopacity, float* dL_dcolor,
	float* dL_dmean3D, float* dL_dcov3D, float* dL_dsh, float* dL_dscale, float* dL_drot, void* stream)
{
	if (P > 0 && (!dL_dmean2D || !dL_dcolor || !dL_dcov3D || !grad_acc)) return fail(OGS_ERR_INVALID_ARG, "a required pointer is NULL");
	FinishExtras x;
	x.grad_acc = grad_acc;
	return backward_finish_impl(P, D, M, W, H, means3D, shs, scales, scale_modifier, rotations, cov3D_precomp, viewmatrix,
	                            campos, radii, geom_buffer, dL_dmean2D, dL_dconic, dL_dopacity, dL_dcolor, dL_dmean3D,
	                            dL_dcov3D, dL_dsh, dL_dscale, dL_drot, x, (cudaStream_t)stream);
}

OGS_API int ogs_lonlat_backward_finish_range(
	int P, int D, int M, int W, int H,
	const float* means3D, const float* shs, const float* scales, float scale_modifier, const float* rotations,
	const float* cov3D_precomp, const float* viewmatrix, const float* campos, const int* radii, char* geom_buffer,
	const float* grad_acc, int first, int count,
	float* dL_dmean2D, float* dL_dconic, float* dL_dopacity, float* dL_dcolor,
	float* dL_dmean3D, float* dL_dcov3D, float* dL_dsh, float* dL_dscale, float* dL_drot, void* stream)
{
	if (P > 0 && (!dL_dmean2D || !dL_dcolor || !dL_dcov3D || !grad_acc)) return fail(OGS_ERR_INVALID_ARG, "a required pointer is NULL");
	FinishExtras x;
	x.grad_acc = grad_acc; x.first = first; x.count = count;
	return backward_finish_impl(P, D, M, W, H, means3D, shs, scales, scale_modifier, rotations, cov3D_precomp, viewmatrix,
	                            campos, radii, geom_buffer, dL_dmean2D, dL_dconic, dL_dopacity, dL_dcolor, dL_dmean3D,
	                            dL_dcov3D, dL_dsh, dL_dscale, dL_drot, x, (cudaStream_t)stream);
}

// One view of a multi-view / data-parallel training step: render backward, then the per-Gaussian
// backward ADDS (accumulate != 0) or writes (== 0, the step's first view) the four geometry gradients and the view's
// densification statistics into the step's bucket and leaves the view's clamp-masked dL/dRGB factor [P,3] instead of a
// 192-byte dL/dsh row; ogs_sh_gradient_from_views rebuilds dL/dsh once per step from all views' factors.
OGS_API int ogs_lonlat_backward_view(
	int P, int D, int M, int64_t num_rendered, int W, int H, const float* background,
	const float* means3D, const float* shs, const float* scales, float scale_modifier, const float* rotations,
	const float* viewmatrix, const float* campos, const int* radii,
	char* geom_buffer, char* binning_buffer, char* img_buffer, const float* dL_dpix,
	int accumulate, float* dL_dmean3D, float* dL_dopacity, float* dL_dscale, float* dL_drot,
	float* dL_drgb_view, float* stat_grad_norm, float* stat_visible, float* stat_max_radius,
	float* dL_dmean2D, void* stream)
{
	if (P > 0 && (!shs || !dL_drgb_view || !scales || !rotations))
		return fail(OGS_ERR_INVALID_ARG, "the view backward needs shs, scales, rotations and dL_drgb_view");
	if (int rc = ogs_lonlat_backward_render(P, num_rendered, W, H, background, geom_buffer, binning_buffer, img_buffer,
	                                        dL_dpix, stream)) return rc;
	FinishExtras x;
	x.accumulate = accumulate ? 1 : 0; x.dL_drgb_view = dL_drgb_view;
	x.stat_grad_norm = stat_grad_norm; x.stat_visible = stat_visible; x.stat_max_radius = stat_max_radius;
	return backward_finish_impl(P, D, M, W, H, means3D, shs, scales, scale_modifier, rotations, nullptr, viewmatrix, campos,
	                            radii, geom_buffer, dL_dmean2D, nullptr, dL_dopacity, nullptr, dL_dmean3D, nullptr, nullptr,
	                            dL_dscale, dL_drot, x, (cudaStream_t)stream);
}

// dL/dsh [P,16,3] (or the split dL/dfeatures_dc + dL/dfeatures_rest) of a whole step from the views' dL/dRGB factors.
// campos_views: device, n_views x 3; dL_drgb_views: HOST array of n_views device pointers (each [P,3]; may be peer memory).
OGS_API int ogs_sh_gradient_from_views(
	int P, int D, int M, int n_views, const float* means3D, const float* campos_views, const float* const* dL_drgb_views,
	float* dL_dsh, float* dL_dfeatures_dc, float* dL_dfeatures_rest, void* stream)
{
	if (P < 0 || n_views < 1 || n_views > kMaxStepViews) return fail(OGS_ERR_INVALID_ARG, "1 <= n_views <= 16");
	if (P == 0) return OGS_OK;
	if (M != 16 || D < 0 || D > 3) return fail(OGS_ERR_INVALID_ARG, "ogs_sh_gradient_from_views needs M == 16, 0 <= D <= 3");
	if (!means3D || !campos_views || !dL_drgb_views) return fail(OGS_ERR_INVALID_ARG, "a required pointer is NULL");
	if ((dL_dsh != nullptr) == (dL_dfeatures_dc != nullptr || dL_dfeatures_rest != nullptr) || (!dL_dsh && (!dL_dfeatures_dc || !dL_dfeatures_rest)))
		return fail(OGS_ERR_INVALID_ARG, "give dL_dsh or (dL_dfeatures_dc, dL_dfeatures_rest)");
	if (dL_dsh && (reinterpret_cast<uintptr_t>(dL_dsh) & 15u)) return fail(OGS_ERR_INVALID_ARG, "dL_dsh must be 16-byte aligned");
	ShFromViewsArgs a{};
	a.P = P; a.D = D; a.M = M; a.n_views = n_views; a.means3D = means3D; a.campos = campos_views;
	for (int v = 0; v < n_views; v++) {
		if (!dL_drgb_views[v]) return fail(OGS_ERR_INVALID_ARG, "a view's dL_drgb pointer is NULL");
		a.drgb[v] = dL_drgb_views[v];
	}
	a.dL_dsh = dL_dsh; a.dL_dfeatures_dc = dL_dfeatures_dc; a.dL_dfeatures_rest = dL_dfeatures_rest;
	return launch_sh_gradient_from_views(a, (cudaStream_t)stream);
}

// Byte offset of the packed render-backward accumulators ([P,12] float, raw sums over pixels: u*dx, u*dy,
// u*dx^2, u*dx*dy, u*dy^2, G*dL/dalpha, colour.rgb terms, 3 pad; linear, so band partials add up) inside a 256-byte aligned geometry buffer.
OGS_API size_t ogs_grad_acc_offset(int P)
{
	if (P <= 0) return 0;
	GeomState g = carve_geom(nullptr, P, nullptr);
	return (size_t)reinterpret_cast<char*>(g.grad_acc);
}

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
	void* stream)
{
	(void)colors_precomp;
	if (int rc = ogs_lonlat_backward_render(P, num_rendered, W, H, background, geom_buffer, binning_buffer,
	                                        img_buffer, dL_dpix, stream)) return rc;
	return ogs_lonlat_backward_finish(P, D, M, W, H, means3D, shs, scales, scale_modifier, rotations, cov3D_precomp,
	                                  viewmatrix, campos, radii, geom_buffer, dL_dmean2D, dL_dconic, dL_dopacity,
	                                  dL_dcolor, dL_dmean3D, dL_dcov3D, dL_dsh, dL_dscale, dL_drot, stream);
}

// ---- perspective camera (camera_type 1): CudaRasterizer::Rasterizer::{forward,backward,markVisible}
OGS_API int ogs_pinhole_forward_stage1(
	int P, int D, int M, int W, int H,
	const float* means3D, const float* shs, const float* colors_precomp, const float* opacities,
	const float* scales, float scale_modifier, const float* rotations, const float* cov3D_precomp,
	const float* viewmatrix, const float* projmatrix, const float* campos, float tan_fovx, float tan_fovy,
	int render_depth, int* radii, char* geom_buffer, char* img_buffer, int64_t* num_rendered_host, void* stream)
{
	const PinholeParams pin{ projmatrix, tan_fovx, tan_fovy, render_depth ? 1 : 0 };
	const int gy = H > 0 ? ceil_div(H, kTile) : 0;
	return forward_stage1_impl(P, D, M, W, H, 0, gy, means3D, shs, colors_precomp, opacities, scales, scale_modifier,
	                           rotations, cov3D_precomp, viewmatrix, campos, radii, geom_buffer, img_buffer,
	                           num_rendered_host, (cudaStream_t)stream, nullptr, nullptr, &pin);
}

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
	void* stream)
{
	(void)colors_precomp;
	cudaStream_t st = (cudaStream_t)stream;
	if (int rc = ogs_lonlat_backward_render(P, num_rendered, W, H, background, geom_buffer, binning_buffer,
	                                        img_buffer, dL_dpix, stream)) return rc;
	if (P == 0) return OGS_OK;
	if (!means3D || !viewmatrix || !projmatrix || !campos || !radii || !geom_buffer ||
	    !dL_dmean2D || !dL_dopacity || !dL_dcolor || !dL_dmean3D || !dL_dcov3D || !dL_dscale || !dL_drot)
		return fail(OGS_ERR_INVALID_ARG, "a required pointer is NULL");
	if (shs && M > 0 && !dL_dsh) return fail(OGS_ERR_INVALID_ARG, "dL_dsh is NULL");
	if (((scales == nullptr) || (rotations == nullptr)) == (cov3D_precomp == nullptr))
		return fail(OGS_ERR_INVALID_ARG, "exactly one of (scales, rotations) / cov3D_precomp must be given");
	GeomState g = GeomState::carve(geom_buffer, P);
	PreprocessBwdArgs a{};
	a.P = P; a.D = D; a.M = shs ? M : 0; a.W = W; a.H = H; a.scale_modifier = scale_modifier;
	a.means3D = means3D; a.shs = shs; a.scales = scales; a.rotations = rotations;
	a.cov3D = cov3D_precomp ? cov3D_precomp : g.cov3D;
	a.viewmatrix = viewmatrix; a.campos = campos; a.radii = radii; a.clamped = g.clamped; a.grad_acc = g.grad_acc;
	a.g0 = g.g0; a.g1 = g.g1;
	a.pinhole = 1; a.projmatrix = projmatrix; a.tan_fovx = tan_fovx; a.tan_fovy = tan_fovy;
	a.focal_y = H / (2.0f * tan_fovy);   // rasterizer_impl.cu:476-477
	a.focal_x = W / (2.0f * tan_fovx);
	a.dL_dmean2D = dL_dmean2D; a.dL_dconic = dL_dconic; a.dL_dopacity = dL_dopacity; a.dL_dcolor = dL_dcolor;
	a.dL_dmean3D = dL_dmean3D; a.dL_dcov3D = dL_dcov3D; a.dL_dsh = shs ? dL_dsh : nullptr;
	a.dL_dscale = dL_dscale; a.dL_drot = dL_drot;
	prof_begin(OGS_PROF_PREPROCESS_BWD, st);
	const int rc = launch_preprocess_bwd(a, st);
	prof_end(OGS_PROF_PREPROCESS_BWD, st);
	return rc;
}

OGS_API int ogs_mark_visible_pinhole(int P, const float* means3D, const float* viewmatrix, const float* projmatrix,
                                     uint8_t* present, void* stream)
{
	(void)projmatrix;   // in_frustum's projected test is commented out in the reference (auxiliary.h:180)
	if (P < 0 || (P > 0 && (!means3D || !viewmatrix || !present))) return fail(OGS_ERR_INVALID_ARG, "bad argument to mark_visible_pinhole");
	if (P == 0) return OGS_OK;
	return launch_check_frustum(P, means3D, viewmatrix, present, (cudaStream_t)stream);
}

// ---- raw-parameter entry points: activations and their backward inside the kernels
OGS_API int ogs_lonlat_forward_raw_stage1(
	int P, int D, int M, int W, int H,
	const float* xyz, const float* features_dc, const float* features_rest, const float* opacity_raw,
	const float* scaling_raw, float scale_modifier, const float* rotation_raw,
	const float* viewmatrix, const float* campos,
	int* radii, char* geom_buffer, char* img_buffer, int64_t* num_rendered_host, void* stream)
{
	if (P > 0 && (!features_dc || (M > 1 && !features_rest))) return fail(OGS_ERR_INVALID_ARG, "features_dc / features_rest are NULL");
	const int gy = H > 0 ? ceil_div(H, kTile) : 0;
	return forward_stage1_impl(P, D, M, W, H, 0, gy, xyz, nullptr, nullptr, opacity_raw, scaling_raw, scale_modifier,
	                           rotation_raw, nullptr, viewmatrix, campos, radii, geom_buffer, img_buffer,
	                           num_rendered_host, (cudaStream_t)stream, features_dc, features_rest);
}

OGS_API int ogs_lonlat_backward_raw(
	int P, int D, int M, int64_t num_rendered, int W, int H, const float* background,
	const float* xyz, const float* features_dc, const float* features_rest,
	const float* scaling_raw, float scale_modifier, const float* rotation_raw,
	const float* viewmatrix, const float* campos, const int* radii,
	char* geom_buffer, char* binning_buffer, char* img_buffer, const float* dL_dpix,
	float* dL_dmean2D, float* dL_dxyz, float* dL_dfeatures_dc, float* dL_dfeatures_rest,
	float* dL_dopacity_raw, float* dL_dscaling_raw, float* dL_drotation_raw, void* stream)
{
	cudaStream_t st = (cudaStream_t)stream;
	if (int rc = ogs_lonlat_backward_render(P, num_rendered, W, H, background, geom_buffer, binning_buffer,
	                                        img_buffer, dL_dpix, stream)) return rc;
	if (P == 0) return OGS_OK;
	if (!xyz || !features_dc || (M > 1 && !features_rest) || !scaling_raw || !rotation_raw || !viewmatrix || !campos || !radii ||
	    !dL_dxyz || !dL_dfeatures_dc || (M > 1 && !dL_dfeatures_rest) || !dL_dopacity_raw || !dL_dscaling_raw || !dL_drotation_raw)
		return fail(OGS_ERR_INVALID_ARG, "a required pointer is NULL");
	if (M < 1 || (D + 1) * (D + 1) > M || D < 0 || D > 3) return fail(OGS_ERR_INVALID_ARG, "SH degree / coefficient count mismatch");
	GeomState g = GeomState::carve(geom_buffer, P);
	PreprocessBwdArgs a{};
	a.P = P; a.D = D; a.M = M; a.W = W; a.H = H; a.scale_modifier = scale_modifier;
	a.means3D = xyz; a.scales = scaling_raw; a.rotations = rotation_raw; a.cov3D = g.cov3D;
	a.viewmatrix = viewmatrix; a.campos = campos; a.radii = radii; a.clamped = g.clamped; a.grad_acc = g.grad_acc;
	a.g0 = g.g0; a.g1 = g.g1;
	a.raw = 1; a.features_dc = features_dc; a.features_rest = features_rest;
	a.dL_dfeatures_dc = dL_dfeatures_dc; a.dL_dfeatures_rest = dL_dfeatures_rest;
	a.dL_dmean2D = dL_dmean2D; a.dL_dopacity = dL_dopacity_raw; a.dL_dmean3D = dL_dxyz;
	a.dL_dscale = dL_dscaling_raw; a.dL_drot = dL_drotation_raw;
	prof_begin(OGS_PROF_PREPROCESS_BWD, st);
	const int rc = launch_preprocess_bwd(a, st);
	prof_end(OGS_PROF_PREPROCESS_BWD, st);
	return rc;
}

// ---- the training step either side of the rasterizer
OGS_API size_t ogs_photometric_loss_workspace_bytes(int W, int H)
{
	return (W > 0 && H > 0) ? photometric_loss_workspace_bytes(W, H) : 0;
}

OGS_API int ogs_photometric_loss(
	int W, int H, int rows_used, float lambda_dssim, const float* rendered, const float* gt,
	const float* mask, int mask_channels, char* workspace, float* loss_out, float* dL_dpix, void* stream)
{
	if (int rc = check_image(W, H)) return rc;
	if (rows_used <= 0 || rows_used > H) return fail(OGS_ERR_INVALID_ARG, "rows_used must be in (0, H]");
	if (!rendered || !gt || !workspace || !loss_out || !dL_dpix) return fail(OGS_ERR_INVALID_ARG, "a required pointer is NULL");
	if (mask && mask_channels != 1 && mask_channels != 3) return fail(OGS_ERR_INVALID_ARG, "mask_channels must be 1 or 3");
	if (reinterpret_cast<uintptr_t>(workspace) & 7u) return fail(OGS_ERR_INVALID_ARG, "workspace must be 8-byte aligned");
	return launch_photometric_loss(W, H, rows_used, lambda_dssim, rendered, gt, mask, mask_channels,
	                               reinterpret_cast<float*>(workspace), loss_out, dL_dpix, (cudaStream_t)stream);
}

OGS_API int ogs_adam_step(
	int groups, float* const* params, const float* const* grads, float* const* exp_avg, float* const* exp_avg_sq,
	const size_t* counts, const float* lrs, int64_t step, double beta1, double beta2, double eps, void* stream)
{
	if (groups < 0 || groups > kAdamMaxGroups) return fail(OGS_ERR_INVALID_ARG, "between 0 and 8 parameter groups");
	if (step < 1) return fail(OGS_ERR_INVALID_ARG, "step counts from 1");
	if (groups && (!params || !grads || !exp_avg || !exp_avg_sq || !counts || !lrs)) return fail(OGS_ERR_INVALID_ARG, "a required pointer is NULL");
	AdamLaunch a{};
	a.groups = groups; a.beta1 = (float)beta1; a.beta2 = (float)beta2; a.eps = (float)eps;
	// torch/csrc/api/src/optim/adam.cpp: bias corrections and step size in double
	const double bc1 = 1.0 - std::pow(beta1, (double)step);
	const double bc2 = 1.0 - std::pow(beta2, (double)step);
	a.sqrt_bias_correction2 = (float)std::sqrt(bc2);
	a.one_minus_beta1 = (float)(1.0 - beta1);
	a.one_minus_beta2 = (float)(1.0 - beta2);
	for (int g = 0; g < groups; g++) {
		if (counts[g] && (!params[g] || !grads[g] || !exp_avg[g] || !exp_avg_sq[g]))
			return fail(OGS_ERR_INVALID_ARG, "a parameter group has a NULL tensor");
		a.group[g] = AdamGroup{ params[g], grads[g], exp_avg[g], exp_avg_sq[g], counts[g], (float)((double)lrs[g] / bc1), 0 };
	}
	return launch_adam(a, (cudaStream_t)stream);
}

OGS_API int ogs_densify_stats(
	int P, const int* radii, const float* dL_dmean2D, float* max_radii2D, float* xyz_gradient_accum, float* denom,
	void* stream)
{
	if (P < 0) return fail(OGS_ERR_INVALID_ARG, "bad P");
	if (P > 0 && (!radii || !dL_dmean2D || !max_radii2D || !xyz_gradient_accum || !denom))
		return fail(OGS_ERR_INVALID_ARG, "a required pointer is NULL");
	return launch_densify_stats(P, radii, dL_dmean2D, max_radii2D, xyz_gradient_accum, denom, (cudaStream_t)stream);
}

OGS_API int ogs_peer_allreduce(float* const* bufs, int world, int rank, size_t count_sum, size_t count_max, void* stream)
{
	if (world < 1 || world > kMaxPeers || rank < 0 || rank >= world) return fail(OGS_ERR_INVALID_ARG, "1 <= world <= 8, 0 <= rank < world");
	if (!bufs || (count_sum & 3u) || (count_max & 3u)) return fail(OGS_ERR_INVALID_ARG, "counts must be multiples of 4 floats");
	for (int r = 0; r < world; r++)
		if (!bufs[r] || (reinterpret_cast<uintptr_t>(bufs[r]) & 15u)) return fail(OGS_ERR_INVALID_ARG, "peer buffers must be 16-byte aligned");
	return launch_peer_allreduce(bufs, world, rank, count_sum, count_max, (cudaStream_t)stream);
}

OGS_API int ogs_multimem_allreduce(float* multicast, int world, int rank, size_t count_sum, size_t count_max, void* stream)
{
	if (world < 1 || rank < 0 || rank >= world) return fail(OGS_ERR_INVALID_ARG, "0 <= rank < world");
	if (!multicast || (reinterpret_cast<uintptr_t>(multicast) & 15u) || (count_sum & 3u))
		return fail(OGS_ERR_INVALID_ARG, "multicast pointer must be 16-byte aligned, count_sum a multiple of 4 floats");
	return launch_multimem_allreduce(multicast, world, rank, count_sum, count_max, (cudaStream_t)stream);
}

OGS_API int ogs_band_rows_allgather(float* const* images, int world, int rank, const float* src, int W, int H,
                                    int y0, int y1, void* stream)
{
	if (world < 1 || world > kMaxPeers || rank < 0 || rank >= world) return fail(OGS_ERR_INVALID_ARG, "1 <= world <= 8, 0 <= rank < world");
	if (int rc = check_image(W, H)) return rc;
	if (!images || !src || y0 < 0 || y1 > H || y1 < y0) return fail(OGS_ERR_INVALID_ARG, "bad band rows");
	for (int r = 0; r < world; r++)
		if (!images[r] || (reinterpret_cast<uintptr_t>(images[r]) & 15u)) return fail(OGS_ERR_INVALID_ARG, "images must be 16-byte aligned");
	if (reinterpret_cast<uintptr_t>(src) & 15u) return fail(OGS_ERR_INVALID_ARG, "src must be 16-byte aligned");
	return launch_band_rows_allgather(images, world, src, W, H, y0, y1, (cudaStream_t)stream);
}

OGS_API int ogs_view_stats(int P, const int* radii, const float* dL_dmean2D, float* grad_norm, float* visible,
                           float* radius, void* stream)
{
	if (P < 0 || (P > 0 && (!radii || !dL_dmean2D || !grad_norm || !visible || !radius)))
		return fail(OGS_ERR_INVALID_ARG, "bad argument to view_stats");
	return launch_view_stats(P, radii, dL_dmean2D, grad_norm, visible, radius, (cudaStream_t)stream);
}

OGS_API int ogs_mark_all_visible(int P, uint8_t* present, void* stream)
{
	if (P < 0 || (P > 0 && !present)) return fail(OGS_ERR_INVALID_ARG, "bad argument to mark_all_visible");
	if (P == 0) return OGS_OK;
	return launch_mark_all_visible(P, present, (cudaStream_t)stream);
}

OGS_API int ogs_export_geometry(
	int P, const char* geom_buffer,
	float* means2D, float* depths, float* conic_opacity, float* rgb,
	uint32_t* tiles_touched, uint8_t* clamped, float* cov3D, void* stream)
{
	if (P <= 0) return OGS_OK;
	if (!geom_buffer) return fail(OGS_ERR_INVALID_ARG, "geom_buffer is NULL");
	GeomState g = GeomState::carve(const_cast<char*>(geom_buffer), P);
	export_geometry_kernel<<<ceil_div(P, 256), 256, 0, (cudaStream_t)stream>>>(
		P, g.g0, g.g1, g.gb, g.depth, g.tiles_touched, g.clamped, g.cov3D,
		means2D, depths, conic_opacity, rgb, tiles_touched, clamped, cov3D);
	OGS_CUDA_TRY(cudaGetLastError());
	return OGS_OK;
}

OGS_API int ogs_export_binning(
	int P, int W, int H, int64_t num_rendered,
	const char* geom_buffer, const char* binning_buffer, const char* img_buffer,
	uint32_t* point_list, uint64_t* point_list_keys, uint32_t* ranges,
	float* final_T, uint32_t* n_contrib, void* stream)
{
	cudaStream_t st = (cudaStream_t)stream;
	if (int rc = check_image(W, H)) return rc;
	if (!img_buffer) return fail(OGS_ERR_INVALID_ARG, "img_buffer is NULL");
	ImageState img = ImageState::carve(const_cast<char*>(img_buffer), W, H);
	const size_t N = (size_t)W * H, T = (size_t)ceil_div(W, kTile) * ceil_div(H, kTile);
	if (ranges) OGS_CUDA_TRY(cudaMemcpyAsync(ranges, img.ranges, sizeof(uint2) * T, cudaMemcpyDeviceToDevice, st));
	if (final_T) OGS_CUDA_TRY(cudaMemcpyAsync(final_T, img.final_T, sizeof(float) * N, cudaMemcpyDeviceToDevice, st));
	if (n_contrib) OGS_CUDA_TRY(cudaMemcpyAsync(n_contrib, img.n_contrib, sizeof(uint32_t) * N, cudaMemcpyDeviceToDevice, st));
	if (num_rendered > 0 && (point_list || point_list_keys)) {
		if (!binning_buffer || !geom_buffer || P <= 0) return fail(OGS_ERR_INVALID_ARG, "buffers missing");
		BinningState b = BinningState::carve(const_cast<char*>(binning_buffer), num_rendered, W, H);
		GeomState g = GeomState::carve(const_cast<char*>(geom_buffer), P);
		if (point_list)
			OGS_CUDA_TRY(cudaMemcpyAsync(point_list, b.point_list, sizeof(uint32_t) * (size_t)num_rendered,
			                             cudaMemcpyDeviceToDevice, st));
		if (point_list_keys)
			if (int rc = launch_rebuild_keys(img, b, g, W, H, point_list_keys, st)) return rc;
	}
	return OGS_OK;
}

// Measurement: how much blending work a rendered frame holds (see pair_count_kernel).  counts: DEVICE array of 4 uint64.
OGS_API int ogs_export_pair_counts(int P, int W, int H, int64_t num_rendered, const char* geom_buffer,
                                   const char* binning_buffer, const char* img_buffer, uint64_t* counts, void* stream)
{
	if (int rc = check_image(W, H)) return rc;
	if (!img_buffer || !counts) return fail(OGS_ERR_INVALID_ARG, "img_buffer / counts is NULL");
	if (P <= 0 || num_rendered <= 0) {
		OGS_CUDA_TRY(cudaMemsetAsync(counts, 0, 4 * sizeof(uint64_t), (cudaStream_t)stream));
		return OGS_OK;
	}
	if (!geom_buffer || !binning_buffer) return fail(OGS_ERR_INVALID_ARG, "buffers missing");
	ImageState img = ImageState::carve(const_cast<char*>(img_buffer), W, H);
	GeomState g = GeomState::carve(const_cast<char*>(geom_buffer), P);
	BinningState b = BinningState::carve(const_cast<char*>(binning_buffer), num_rendered, W, H);
	return launch_pair_count(img.ranges, b.point_list, W, H, g.g0, g.g1, img.n_contrib,
	                         reinterpret_cast<unsigned long long*>(counts), (cudaStream_t)stream);
}

} // extern "C"
