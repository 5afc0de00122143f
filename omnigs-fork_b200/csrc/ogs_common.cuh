// ogs_common.cuh — shared definitions for the B200 (sm_100a) lonlat rasterizer kernels.
//
// Private to the shared library (libomnigs_b200.so).  The public surface is include/omnigs_b200.h.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/omnigs_b200.h"

namespace ogs {

constexpr int kTile = 16;              // reference config.h:26-27 (BLOCK_X = BLOCK_Y = 16)
constexpr int kTilePixels = kTile * kTile;
constexpr int kNumSMs = 148;           // B200: 2 dies x 74 SMs
constexpr float kEps7 = 0.0000001f;    // the reference's 1e-7 guards
constexpr float kPiInv = 0.318309886183790671537767526745028724f;    // M_1_PIf32
constexpr float kTwoPiInv = 0.636619772367581343075535053490057448f; // M_2_PIf32
constexpr float kAlphaMin = 1.0f / 255.0f;

#define OGS_HD __host__ __device__ __forceinline__
#define OGS_D __device__ __forceinline__

#define OGS_CUDA_TRY(expr)                                  \
	do {                                                    \
		cudaError_t _e = (expr);                            \
		if (_e != cudaSuccess) return ogs::fail_cuda(_e);   \
	} while (0)

void prof_begin(int stage, cudaStream_t st);   // no-ops unless ogs_profile_enable(1)
void prof_end(int stage, cudaStream_t st);
int fail_cuda(cudaError_t e);   // records the message, returns OGS_ERR_CUDA
int fail(int code, const char* msg);

OGS_HD size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
OGS_HD int ceil_div(int a, int b) { return (a + b - 1) / b; }

// Number of key bits the tile id occupies — reference rasterizer_impl.cu:47-62 (getHigherMsb)
inline uint32_t higher_msb(uint32_t n)
{
	uint32_t msb = sizeof(n) * 4, step = msb;
	while (step > 1) {
		step /= 2;
		if (n >> msb) msb += step; else msb -= step;
	}
	if (n >> msb) msb++;
	return msb;
}

// ------------------------------------------------------------------ private buffer layouts
// The caller (LibTorch shim / Python host) owns three opaque byte buffers, exactly like the
// reference (rasterize_points.cu:87-94).  Their contents are private between our forward and
// our backward.  All sub-arrays are 256-byte aligned.

struct GeomState {               // per-Gaussian, P-sized
	float4* g0;                  // (mean2D.x, mean2D.y, conic.x, conic.y)
	float4* g1;                  // (conic.z, opacity, colour.r, colour.g)
	float2* gb;                  // (colour.b, alpha cut-off in `power`: see alpha_cutoff_power, render_common.cuh)
	float* depth;                // r = |t|  (sort key, reference forward.cu:697)
	uint2* rect;                 // x0 | x1<<16 , y0 | y1<<16   (clamped tile rect, auxiliary.h:56-66)
	uint32_t* tiles_touched;
	float* cov3D;                // 6 per Gaussian (only written when not precomputed)
	uint8_t* clamped;            // bit c set <=> SH colour channel c was clamped (forward.cu:79-81)
	uint32_t* sort_key[2];       // depth-sort ping-pong (float bits; culled = 0xFFFFFFFF)
	uint32_t* sort_val[2];       // Gaussian index ping-pong; sort_val[0] ends as depth order
	uint32_t* emit_offset;       // P+1: exclusive prefix of tiles_touched in depth order
	uint32_t* depth_hist;        // 4 x 256 digit counts for the depth sort
	uint32_t* depth_status;      // decoupled look-back status words for the 4 depth passes
	unsigned long long* scan_status; // look-back status for the emit-offset scan (64-bit words: sums up to 2^31-1)
	float* grad_acc;             // 12 per Gaussian: render-backward accumulators (zeroed by backward)
	unsigned long long* scalars; // [0] = sum tiles_touched (num_rendered), [1..] tickets/counters
	size_t scalars_bytes;
	char* zero_begin;            // region that stage 1 must zero-fill every call
	size_t zero_bytes;
	static size_t bytes(int P);
	static GeomState carve(char* base, int P);
};

struct ImageState {              // per pixel / per tile
	float* final_T;              // N
	uint32_t* n_contrib;         // N
	uint2* ranges;               // T
	int* tile_diff;              // (gy+1)*(gx+1) 2-D difference array of tile coverage counts
	uint32_t* tile_count;        // T
	uint32_t* tile_hist;         // kMaxTilePasses x kMaxBins digit STARTS (exclusive prefix of the counts) for the tile sort
	static size_t bytes(int W, int H);
	static ImageState carve(char* base, int W, int H);
};

constexpr int kSortItemsPerBlock = 2048; // onesweep tile size (256 threads x 8 keys)
constexpr int kMaxBins = 512;            // <= 9-bit digits
constexpr int kMaxTilePasses = 4;

struct TileSortPlan {            // how the tile-id bits are split into radix passes
	int bit;                     // getHigherMsb(T)
	int passes;
	int shift[kMaxTilePasses];
	int bits[kMaxTilePasses];
	// column-segment path (binning.cu): one pass over the Gaussians' column segments on x, one over the instances on y
	int segments;                // 1 when the grid fits (gx, gy <= 512) and the classic plan needs two passes
	int bits_x, bits_y;
};
constexpr int kColStartsRow = 2;         // rows of ImageState::tile_hist that hold the digit starts of the segment path
constexpr int kRowStartsRow = 3;
TileSortPlan make_tile_sort_plan(int W, int H);

struct BinningState {            // per tile instance, R-sized
	uint32_t* key[2];            // tile id ping-pong
	uint32_t* val[2];            // Gaussian idx ping-pong
	uint32_t* point_list;        // alias of val[passes & 1]: the sorted list
	uint32_t* first_src;         // R/2048 + 2: depth-ordered source owning the first slot of each emit block
	uint32_t* status;            // look-back status, passes x tiles x bins
	unsigned int* tickets;       // one dynamic-tile-id counter per pass
	// column-segment path: key[0] holds the x-sorted segments' Gaussian ids, key[1] their emission offsets
	unsigned long long* seg_scan_status;   // look-back words of the scan over the sorted segments
	int* col_diff;               // kMaxBins + 1: difference array of the segments' x coverage
	char* zero_begin;
	size_t zero_bytes;
	static size_t bytes(int64_t R, int W, int H);
	static BinningState carve(char* base, int64_t R, int W, int H);
};

// ------------------------------------------------------------------ small device helpers
OGS_D float warp_sum(float v)
{
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
	return v;
}

OGS_D unsigned lane_id() { return threadIdx.x & 31; }
OGS_D unsigned lanemask_lt()
{
	unsigned m;
	asm volatile("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
	return m;
}

// streaming (evict-first) loads / stores for data touched exactly once
OGS_D float4 ld_stream_f4(const float4* p)
{
	float4 r;
	asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
	             : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
	return r;
}

// vector reduction to global memory: one L2 RED for four floats (sm_90+)
OGS_D void red_add_v4(float* addr, float a, float b, float c, float d)
{
	asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};"
	             :: "l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
OGS_D float rcp_approx(float x)
{
	float r;
	asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
	return r;
}
OGS_D void red_add(float* addr, float a)
{
	asm volatile("red.global.add.f32 [%0], %1;" :: "l"(addr), "f"(a) : "memory");
}

// Decoupled look-back status words carry flag and value in ONE 32-bit word, so no ordering with any
// other location is needed: relaxed gpu-scope accesses (L2-coherent, no fence, no L1 invalidate) suffice.
// (acquire/release here costs a MEMBAR.GPU + CCTL.IVALL per access and serialises the key loads.)
OGS_D uint32_t ld_acquire(const uint32_t* p)
{
	uint32_t v;
	asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}
OGS_D void st_release(uint32_t* p, uint32_t v)
{
	asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}

} // namespace ogs
