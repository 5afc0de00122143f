// binning.cu — from per-Gaussian tile rects to per-tile, depth-ordered Gaussian lists.
//
// Replaces the reference's InclusiveSum + duplicateWithKeys + 64-bit DeviceRadixSort::SortPairs +
// memset + identifyTileRanges (cuda_rasterizer/rasterizer_impl.cu:622-679) with a B200-first
// pipeline that yields bit-identical ranges / point lists:
//
//   The reference sorts R = num_rendered (tile<<32 | depth_bits) keys with a stable LSD radix sort
//   over 32+bit bits.  An LSD sort processes the low 32 (depth) bits first, and all tile instances
//   of a Gaussian share them, so those passes commute with the duplication step:
//     1. stable-sort the P Gaussians by depth bits            (P-sized, 4 x 8-bit onesweep passes)
//     2. emit tile instances in that order, row-major inside each rect (load-balanced expansion)
//     3. stable-sort the R instances by tile id only           (ceil(bit/9) <= 2 onesweep passes; the first one
//        generates its input with the emission logic instead of loading it)
//   Equal (tile, depth) keys end in ascending Gaussian index, as with the reference (stability).
//   Tile ranges never need the sorted keys: the per-tile instance counts come from a 2-D
//   difference array that preprocess fills with 4 atomics per Gaussian; their prefix sum IS
//   `ranges`, and the radix digit histograms are its marginals.
//
// HBM traffic at R instances: 8R (first pass, which emits its own input) + 12R (second pass) = 20R bytes, against
// 12R + (8 + 24*6)R = 164R for the reference's data flow.
//
// Column-segment path (default whenever the tile grid is at most 512 x 512; "classic" above otherwise and with
// OGS_SEGMENT_SORT=0).  A Gaussian's rect is w columns of h tiles.  Sorting by (y, x) = sorting stably by x, then by y, and
// all h tiles of one column of one Gaussian share x — so the x pass runs over the S = sum(w) ~ R / 5 column SEGMENTS
// (emitted on the fly from the depth order), and only the y pass touches all R instances (expanding every sorted segment
// along y on the fly, slot = y0 + local: no division) and writes the final list: ONE R-sized pass that stores 4 bytes per
// instance.  Extra work: a second emission-offset scan over the S sorted segments and the x digit starts (a difference
// array filled by the first scan).  Same output bits: the order is the stable (y, x, depth, index) order either way.
#include "ogs_common.cuh"
#include "launchers.cuh"
#include <cstdlib>

namespace ogs {

constexpr int kSortThreads = 256;
constexpr int kSortItems = kSortItemsPerBlock / kSortThreads;   // 8 keys per thread
// Look-back status words.  The reference accepts any int num_rendered (rasterizer_impl.cu:627-632), so an inclusive
// digit prefix needs 31 bits.  32-bit words (tile sort, depth sort; partial counts <= 2048):
//   0 = not published, 1 .. 2^31-1 = partial count + 1, bit 31 set = inclusive prefix in the low 31 bits.
// 64-bit words (emission-offset scan, whose per-tile partial sums are unbounded below 2^31): flag in bits 62-63.
struct Status32 {
	typedef uint32_t word;
	static OGS_D word partial(uint32_t v) { return v + 1u; }
	static OGS_D word inclusive(uint32_t v) { return 0x80000000u | v; }
	static OGS_D bool empty(word w) { return w == 0u; }
	static OGS_D bool is_inclusive(word w) { return (w & 0x80000000u) != 0u; }
	static OGS_D uint32_t raw(word w) { return w; }
	static constexpr uint32_t kPartialBias = 1u, kInclusiveBias = 0x80000000u;
	static OGS_D word load(const word* p) { return ld_acquire(p); }
	static OGS_D void store(word* p, word w) { st_release(p, w); }
};
struct Status64 {
	typedef unsigned long long word;
	static OGS_D word partial(uint32_t v) { return (1ull << 62) | v; }
	static OGS_D word inclusive(uint32_t v) { return (2ull << 62) | v; }
	static OGS_D bool empty(word w) { return (w >> 62) == 0ull; }
	static OGS_D bool is_inclusive(word w) { return (w >> 62) == 2ull; }
	static OGS_D uint32_t raw(word w) { return (uint32_t)w; }
	static constexpr uint32_t kPartialBias = 0u, kInclusiveBias = 0u;
	static OGS_D word load(const word* p)
	{
		word v;
		asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
		return v;
	}
	static OGS_D void store(word* p, word w) { asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(w) : "memory"); }
};

// ------------------------------------------------------------------ block-wide exclusive scan
// Exclusive prefix sum over `n` (<= 2*kSortThreads) smem words, in place; returns nothing.
// All kSortThreads threads must call.
__device__ void block_exclusive_scan_512(uint32_t* data, int n, uint32_t* warp_tmp /*>= 8*/)
{
	const int tid = threadIdx.x;
	uint32_t a = (2 * tid < n) ? data[2 * tid] : 0u;
	uint32_t b = (2 * tid + 1 < n) ? data[2 * tid + 1] : 0u;
	uint32_t sum = a + b;
	uint32_t incl = sum;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) {
		uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
		if ((tid & 31) >= o) incl += v;
	}
	if ((tid & 31) == 31) warp_tmp[tid >> 5] = incl;
	__syncthreads();
	uint32_t warp_off = 0;
#pragma unroll
	for (int w = 0; w < kSortThreads / 32; w++)
		if (w < (tid >> 5)) warp_off += warp_tmp[w];
	uint32_t excl = warp_off + incl - sum;
	if (2 * tid < n) data[2 * tid] = excl;
	if (2 * tid + 1 < n) data[2 * tid + 1] = excl + a;
	__syncthreads();
}

// ------------------------------------------------------------------ depth-key digit histogram
// 4 x 256 counts of the 8-bit digits of n 32-bit keys (the upfront histogram of onesweep).
__global__ void __launch_bounds__(256) depth_histogram_kernel(const uint32_t* __restrict__ keys, uint32_t n,
                                                              uint32_t* __restrict__ hist /*4*256*/)
{
	__shared__ uint32_t s[4 * 256];
	for (int i = threadIdx.x; i < 4 * 256; i += blockDim.x) s[i] = 0;
	__syncthreads();
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		uint32_t k = keys[i];
		atomicAdd(&s[k & 255u], 1u);
		atomicAdd(&s[256 + ((k >> 8) & 255u)], 1u);
		atomicAdd(&s[512 + ((k >> 16) & 255u)], 1u);
		atomicAdd(&s[768 + (k >> 24)], 1u);
	}
	__syncthreads();
	for (int i = threadIdx.x; i < 4 * 256; i += blockDim.x)
		if (s[i]) atomicAdd(&hist[i], s[i]);
}

// ------------------------------------------------------------------ onesweep radix pass
// One stable LSD pass over (key,value) pairs on digit (key >> shift) & (2^bits - 1), bits <= 9.
// Chained-scan ("onesweep") formulation: every block takes a dynamic tile id, ranks its 2048
// keys locally, publishes its per-digit counts and resolves its global offsets by decoupled
// look-back over the predecessors' status words; digit totals come from `digit_counts`.
// vals_in == nullptr means value i = i (first depth pass); keys_out == nullptr skips the key
// write (last tile pass).  Sized for occupancy: 256 threads x 8 keys, values are only loaded when
// in flight while the keys are ranked.  Ranking uses one ballot per digit bit (peers = AND of the
// matching ballots): MATCH.ANY costs ~46 cycles per SM per warp-instruction on B200 and dominated the
// first version of this kernel (ncu stall sampling, profiles/), 7-9 VOTEs are several times cheaper.
struct OnesweepSmem {
	uint32_t warp_hist[kSortThreads / 32][kMaxBins]; // per-warp digit counters -> warp offsets
	uint32_t bin_start[kMaxBins];                    // exclusive prefix of the tile's digit totals
	uint32_t global_base[kMaxBins];                  // global offset of digit run minus bin_start
	uint32_t keys[kSortItemsPerBlock];
	uint32_t vals[kSortItemsPerBlock];
	uint32_t tile_hist[kMaxBins];                    // digit counts of this tile (published early)
	uint32_t warp_tmp[8];
};

constexpr int kLookbackBatch = 8;

// Load-balanced emission of tile instances (emit_prepare / emit_slot below): staging area for one block of
// kEmitPerBlock consecutive output slots.  The fused first tile-sort pass overlays it on OnesweepSmem.
constexpr int kEmitPerBlock = 2048;
static_assert(kEmitPerBlock == kSortItemsPerBlock, "the fused emit + sort pass produces one sort tile per block");
struct EmitSmem {
	uint32_t off[kEmitPerBlock + 2];
	uint32_t gid[kEmitPerBlock + 1];
	uint2 rect[kEmitPerBlock + 1];
	uint32_t src[kEmitPerBlock];      // slot -> local source index (after the max-scan)
	uint32_t warp_max[kSortThreads / 32];
	uint32_t first, last;
};
constexpr size_t kOnesweepSmemBytes = sizeof(OnesweepSmem) > sizeof(EmitSmem) ? sizeof(OnesweepSmem) : sizeof(EmitSmem);

// Decoupled look-back for one counter: sum the predecessors' published values back to the nearest
// inclusive one.  Status words of kLookbackBatch predecessors are fetched together, so the walk
// costs one L2 round trip per batch instead of one per tile (with several hundred CTAs in flight
// the not-yet-inclusive window is hundreds of tiles long).
template <typename S>
OGS_D uint32_t lookback_sum(const typename S::word* __restrict__ status, int tile, size_t stride, size_t offset)
{
	// Sum of the published values back to the nearest inclusive word.  The words are added up RAW and the encoding's
	// biases (Status32: +1 per partial word, bit 31 on the inclusive one) are taken off once at the end, which keeps the
	// per-word work at one add + two tests.
	uint32_t raw = 0, partials = 0;
	int t = tile - 1;
	while (true) {
		typename S::word s[kLookbackBatch];
#pragma unroll
		for (int i = 0; i < kLookbackBatch; i++)
			s[i] = (t - i >= 0) ? S::load(&status[(size_t)(t - i) * stride + offset]) : S::inclusive(0u);
		bool done = false;
		int consumed = kLookbackBatch;
#pragma unroll
		for (int i = 0; i < kLookbackBatch; i++) {
			if (!done && consumed == kLookbackBatch) {
				if (S::empty(s[i])) {
					consumed = i;            // not published yet: poll again from this tile
				} else {
					raw += S::raw(s[i]);
					if (S::is_inclusive(s[i])) done = true;
					else partials++;
				}
			}
		}
		if (done) break;
		t -= consumed;
	}
	return raw - partials * S::kPartialBias - S::kInclusiveBias;
}


// The same walk by a whole warp: 32 predecessors per L2 round trip instead of kLookbackBatch.  The scan kernels have ONE
// counter per tile, and the thread that walked it alone kept the other 255 of its CTA at a barrier for 31-44 % of those
// kernels' stall samples (profiles/r02_ncu_summary.md): several hundred CTAs start together, so the window of
// not-yet-inclusive tiles is several hundred tiles long.  All 32 lanes must call; every lane returns the sum.
template <typename S>
OGS_D uint32_t lookback_sum_warp(const typename S::word* __restrict__ status, int tile, int lane)
{
	uint32_t total = 0;
	int t = tile - 1;                 // nearest predecessor not yet accounted for
	while (true) {
		const int idx = t - lane;
		const typename S::word w = (idx >= 0) ? S::load(&status[idx]) : S::inclusive(0u);
		const bool empty = S::empty(w);
		const bool incl = !empty && S::is_inclusive(w);
		const unsigned m_empty = __ballot_sync(0xffffffffu, empty);
		const unsigned m_incl = __ballot_sync(0xffffffffu, incl);
		// lanes usable this round: those before the first unpublished word, up to and including the first inclusive one
		const int first_empty = m_empty ? (__ffs(m_empty) - 1) : 32;
		const int first_incl = m_incl ? (__ffs(m_incl) - 1) : 32;
		const bool done = first_incl < first_empty;
		const int take = done ? first_incl + 1 : first_empty;      // lanes [0, take)
		uint32_t v = 0;
		if (lane < take) {
			v = S::raw(w);
			// take the encoding's bias off per word (Status32: +1 on partial words, bit 31 on the inclusive one)
			v -= (lane == first_incl && done) ? S::kInclusiveBias : S::kPartialBias;
		}
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
		total += v;
		if (done) break;
		t -= take;                    // take == 0: nothing new was published, poll again
	}
	return total;
}

// Slot -> (tile id, Gaussian id) for the kEmitPerBlock output slots [o0, o1) of one block.  Output slot o belongs to
// the depth-ordered Gaussian i with emit_offset[i] <= o < emit_offset[i+1]; inside a Gaussian, slots walk its tile
// rect row-major (the order of duplicateWithKeys, rasterizer_impl.cu:127-138).  After this call em.src / em.off /
// em.gid / em.rect are ready and emit_slot() evaluates any slot of the block.
OGS_D void emit_prepare(EmitSmem& em, uint32_t block, uint32_t o0, uint32_t o1,
                        const uint32_t* __restrict__ emit_offset, const uint32_t* __restrict__ order,
                        const uint2* __restrict__ rect, const uint32_t* __restrict__ first_src)
{
	constexpr int kPerThread = kEmitPerBlock / kSortThreads;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	// sources owning the first and the last slot of this block (table written by gather_scan_kernel)
	if (tid == 0) em.first = first_src[block];
	if (tid == 1) em.last = first_src[block + 1];   // owner of slot min(o1, R-1): may start exactly at o1
#pragma unroll
	for (int k = 0; k < kPerThread; k++) em.src[tid + k * kSortThreads] = 0u;
	__syncthreads();
	const uint32_t first = em.first, last = em.last;
	// every source in [first, last) owns >= 1 slot of this block (zero-count Gaussians sort to the end
	// of the depth order), so ns <= kEmitPerBlock + 1 and the head marks below never collide
	const uint32_t ns = last - first + 1;
	for (uint32_t i = tid; i < ns; i += kSortThreads) {
		const uint32_t g = order[first + i];
		const uint32_t off = emit_offset[first + i];
		em.off[i] = off;
		em.gid[i] = g;
		em.rect[i] = rect[g];
		if (i > 0 && off < o1) em.src[off - o0] = i;   // head of source i (source 0 starts at or before slot 0)
	}
	__syncthreads();
	// inclusive max-scan of the head marks: slot -> source index
	uint32_t v[kPerThread];
	uint32_t run = 0;
#pragma unroll
	for (int k = 0; k < kPerThread; k++) {
		run = max(run, em.src[tid * kPerThread + k]);
		v[k] = run;
	}
	uint32_t incl = run;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) {
		const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o);
		if (lane >= o) incl = max(incl, u);
	}
	if (lane == 31) em.warp_max[warp] = incl;
	const uint32_t prev_lane = __shfl_up_sync(0xffffffffu, incl, 1);
	__syncthreads();
	uint32_t carry = (lane > 0) ? prev_lane : 0u;
	for (int w = 0; w < warp; w++) carry = max(carry, em.warp_max[w]);
#pragma unroll
	for (int k = 0; k < kPerThread; k++) em.src[tid * kPerThread + k] = max(v[k], carry);
	__syncthreads();
}
OGS_D void emit_slot(const EmitSmem& em, uint32_t o, uint32_t o0, int gx, uint32_t& tile_key, uint32_t& gid)
{
	const uint32_t src = em.src[o - o0];
	const uint2 rc = em.rect[src];
	const uint32_t x0 = rc.x & 0xFFFFu, x1 = rc.x >> 16, y0 = rc.y & 0xFFFFu;
	const uint32_t w = x1 - x0;
	const uint32_t local = o - em.off[src];
	// q = local / w without the ~35-instruction integer division: local < 2^31 and q < 2^16, so the float quotient is
	// off by at most one and a two-sided correction makes it exact
	uint32_t q = __float2uint_rz(__uint2float_rz(local) * rcp_approx(__uint2float_rz(w)));
	int rem = (int)(local - q * w);
	if (rem < 0) { q--; rem += (int)w; }
	else if (rem >= (int)w) { q++; rem -= (int)w; }
	uint32_t x = x0 + (uint32_t)rem;
	if (x >= (uint32_t)gx) x -= (uint32_t)gx;   // only rects that wrap around the longitude seam (opt-in mode)
	tile_key = (y0 + q) * (uint32_t)gx + x;
	gid = em.gid[src];
}

// Column-segment path, x pass: slot -> (x, Gaussian id) of the local-th column of the owning Gaussian's rect.
OGS_D void emit_slot_column(const EmitSmem& em, uint32_t o, uint32_t o0, int gx, uint32_t& key, uint32_t& gid)
{
	const uint32_t src = em.src[o - o0];
	uint32_t x = (em.rect[src].x & 0xFFFFu) + (o - em.off[src]);
	if (x >= (uint32_t)gx) x -= (uint32_t)gx;   // rects that wrap around the longitude seam (opt-in mode)
	key = x;
	gid = em.gid[src];
}
// Column-segment path, y pass: slot -> (y, Gaussian id) of the local-th tile of the owning column segment.
OGS_D void emit_slot_row(const EmitSmem& em, uint32_t o, uint32_t o0, uint32_t& key, uint32_t& gid)
{
	const uint32_t src = em.src[o - o0];
	key = (em.rect[src].y & 0xFFFFu) + (o - em.off[src]);
	gid = em.gid[src];
}

// kEmit: the pass generates its (tile key, Gaussian id) pairs itself (emit_prepare / emit_slot) instead of loading
// them: the first tile-sort pass then needs no materialised unsorted list (saves writing and re-reading 8 R bytes).
struct EmitSource {
	const uint32_t* emit_offset;
	const uint32_t* order;
	const uint2* rect;
	const uint32_t* first_src;
	int gx;
};

// kEmit: 0 = keys / values are loaded, 1 = rect emission (tile id keys), 2 = column segments (x keys), 3 = segments
// expanded along y (y keys).  n_dev != NULL: the item count is a device value (number of column segments) and the grid an
// upper bound; CTAs beyond it leave at once.
template <int BITS, int kEmit = 0>
#ifndef OGS_SORT_MINBLOCKS
#define OGS_SORT_MINBLOCKS 5
#endif
__global__ void __launch_bounds__(kSortThreads, OGS_SORT_MINBLOCKS) onesweep_pass_kernel(
	const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
	uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out,
	uint32_t n, int shift,
	const uint32_t* __restrict__ digit_counts, uint32_t* __restrict__ status, unsigned int* __restrict__ ticket,
	const EmitSource es, const bool counts_are_starts, const uint32_t* __restrict__ n_dev)
{
	extern __shared__ __align__(16) unsigned char smem_raw[];
	OnesweepSmem& sm = *reinterpret_cast<OnesweepSmem*>(smem_raw);
	__shared__ uint32_t s_tile;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	constexpr int nbins = 1 << BITS;
	constexpr uint32_t mask = (uint32_t)nbins - 1u;

	if (tid == 0) s_tile = atomicAdd(ticket, 1u);
	if (n_dev) {
		n = *n_dev;
		__syncthreads();
		if ((uint64_t)s_tile * kSortItemsPerBlock >= n) return;   // uniform: the grid was sized for an upper bound of n
	}
	uint32_t key[kSortItems];
	uint32_t val[kSortItems];
	if constexpr (kEmit != 0) {
		// the emission staging area overlays the sort's shared memory; it is dead once the pairs are in registers
		__syncthreads();
		EmitSmem& em = *reinterpret_cast<EmitSmem*>(smem_raw);
		const uint32_t o0 = s_tile * (uint32_t)kSortItemsPerBlock, o1 = min(n, o0 + (uint32_t)kSortItemsPerBlock);
		emit_prepare(em, s_tile, o0, o1, es.emit_offset, es.order, es.rect, es.first_src);
#pragma unroll
		for (int k = 0; k < kSortItems; k++) {
			const uint32_t idx = o0 + warp * (32 * kSortItems) + k * 32 + lane;
			key[k] = 0xFFFFFFFFu;
			val[k] = 0u;
			if (idx < n) {
				if constexpr (kEmit == 1) emit_slot(em, idx, o0, es.gx, key[k], val[k]);
				else if constexpr (kEmit == 2) emit_slot_column(em, idx, o0, es.gx, key[k], val[k]);
				else emit_slot_row(em, idx, o0, key[k], val[k]);
			}
		}
		__syncthreads();
	}
	for (int w = 0; w < kSortThreads / 32; w++)
		for (int b = tid; b < nbins; b += kSortThreads) sm.warp_hist[w][b] = 0;
	// global digit starts: the exclusive scan of the pass histogram (already scanned when counts_are_starts)
	for (int b = tid; b < nbins; b += kSortThreads) {
		sm.global_base[b] = digit_counts[b];
		sm.tile_hist[b] = 0;
	}
	__syncthreads();
	if (!counts_are_starts) block_exclusive_scan_512(sm.global_base, nbins, sm.warp_tmp);

	const uint32_t tile = s_tile;
	const uint32_t tile_base = tile * (uint32_t)kSortItemsPerBlock;
	const uint32_t tile_count = min((uint32_t)kSortItemsPerBlock, n - tile_base);

	// ---- load (warp-striped); count digits and publish the tile's counts before the (slow) ranking ----
	uint32_t rank[kSortItems];
	const uint32_t warp_base = tile_base + warp * (32 * kSortItems);
	if constexpr (kEmit == 0) {
#pragma unroll
		for (int k = 0; k < kSortItems; k++) {
			uint32_t idx = warp_base + k * 32 + lane;
			key[k] = (idx < n) ? keys_in[idx] : 0xFFFFFFFFu;
		}
	}
#pragma unroll
	for (int k = 0; k < kSortItems; k++) {
		uint32_t idx = warp_base + k * 32 + lane;
		if (idx < n) atomicAdd(&sm.tile_hist[(key[k] >> shift) & mask], 1u);
	}
	__syncthreads();
	for (int b = tid; b < nbins; b += kSortThreads)
		Status32::store(&status[(size_t)tile * nbins + b], tile == 0 ? Status32::inclusive(sm.tile_hist[b]) : Status32::partial(sm.tile_hist[b]));
	// values travel with the keys: issue their loads now, they are consumed after the look-back
	if constexpr (kEmit == 0) {
#pragma unroll
		for (int k = 0; k < kSortItems; k++) {
			uint32_t idx = warp_base + k * 32 + lane;
			val[k] = (idx < n) ? (vals_in ? vals_in[idx] : idx) : 0u;
		}
	}
#pragma unroll
	for (int k = 0; k < kSortItems; k++) {
		const uint32_t idx = warp_base + k * 32 + lane;
		const bool valid = idx < n;
		const uint32_t d = (key[k] >> shift) & mask;
		unsigned peers = __ballot_sync(0xffffffffu, valid);
#pragma unroll
		for (int bit = 0; bit < BITS; bit++) {
			const unsigned one = (d >> bit) & 1u;
			const unsigned b = __ballot_sync(0xffffffffu, one != 0u);
			peers &= b ^ (one - 1u);   // lanes whose bit equals mine: b if set, ~b if clear
		}
		const int rank_in = __popc(peers & lanemask_lt());
		uint32_t prev = 0;
		if (valid && rank_in == 0) {
			prev = sm.warp_hist[warp][d];
			sm.warp_hist[warp][d] = prev + __popc(peers);
		}
		__syncwarp();
		prev = __shfl_sync(0xffffffffu, prev, __ffs(peers) - 1);
		rank[k] = prev + rank_in;
	}
	__syncthreads();

	// ---- per-digit totals, warp offsets, tile-local digit starts ----
	for (int b = tid; b < nbins; b += kSortThreads) {
		uint32_t run = 0;
#pragma unroll
		for (int w = 0; w < kSortThreads / 32; w++) {
			uint32_t c = sm.warp_hist[w][b];
			sm.warp_hist[w][b] = run;
			run += c;
		}
		sm.bin_start[b] = run; // total for now
	}
	__syncthreads();

	// this thread's digit totals (needed for the inclusive status after bin_start has become the local starts)
	constexpr int kDigitsPerThread = (nbins + kSortThreads - 1) / kSortThreads;
	uint32_t digit_total[kDigitsPerThread];
#pragma unroll
	for (int q = 0; q < kDigitsPerThread; q++) {
		const int b = tid + q * kSortThreads;
		digit_total[q] = (b < nbins) ? sm.bin_start[b] : 0u;
	}
	__syncthreads();
	block_exclusive_scan_512(sm.bin_start, nbins, sm.warp_tmp); // totals -> tile-local starts

	// ---- stage into digit order: needs nothing from other tiles, so it runs BEFORE the look-back — the predecessors'
	// status words have that much longer to arrive and the threads without a digit wait that much less ----
#pragma unroll
	for (int k = 0; k < kSortItems; k++) {
		uint32_t idx = warp_base + k * 32 + lane;
		if (idx < n) {
			uint32_t d = (key[k] >> shift) & mask;
			uint32_t pos = sm.bin_start[d] + sm.warp_hist[warp][d] + rank[k];
			sm.keys[pos] = key[k];
			sm.vals[pos] = val[k];
		}
	}

	// ---- decoupled look-back (one thread per digit, batched status loads) ----
#pragma unroll
	for (int q = 0; q < kDigitsPerThread; q++) {
		const int b = tid + q * kSortThreads;
		if (b < nbins) {
			uint32_t excl = 0;
			if (tile != 0) {
				excl = lookback_sum<Status32>(status, (int)tile, (size_t)nbins, (size_t)b);
				Status32::store(&status[(size_t)tile * nbins + b], Status32::inclusive(excl + digit_total[q]));
			}
			// digit start + keys of this digit in earlier tiles - tile-local start: slot j of the staged tile goes to base + j
			sm.global_base[b] += excl - sm.bin_start[b];
		}
	}
	__syncthreads();
	for (uint32_t j = tid; j < tile_count; j += kSortThreads) {
		uint32_t kk = sm.keys[j];
		uint32_t d = (kk >> shift) & mask;
		uint32_t out = sm.global_base[d] + j;
		if (keys_out) keys_out[out] = kk;
		vals_out[out] = sm.vals[j];
	}
}

// ------------------------------------------------------------------ exclusive scan (look-back)
// out[i] = sum_{j<i} counts[order[j]] for i in [0, n]; out has n+1 entries.
// Also fills first_src[b] = index i of the source that owns output slot min(b*kEmitPerBlock, R-1)
// for b in [0, ceil(R/kEmitPerBlock)], so the emission needs no search (R = *total).
struct ScanSmem {
	uint32_t warp_tmp[8];
	uint32_t tile;
	uint32_t tile_excl;
};
// kCount: what a source contributes — 0: counts[order[i]] (tiles of a Gaussian's rect), 1: the rect's width (column segments
// of a Gaussian; also fills the difference array of the segments' x coverage), 2: the rect's height (tiles of one column
// segment; `order` = Gaussian id of the sorted segments, n = *n_dev of them, the grid is an upper bound).
template <int kCount>
__global__ void __launch_bounds__(kSortThreads) gather_scan_kernel(
	const uint32_t* __restrict__ counts, const uint32_t* __restrict__ order, uint32_t n, const uint32_t* __restrict__ n_dev,
	const uint2* __restrict__ rect, int gx, int* __restrict__ col_diff,
	uint32_t* __restrict__ out, unsigned long long* __restrict__ status, unsigned int* __restrict__ ticket,
	uint32_t* __restrict__ first_src, const unsigned long long* __restrict__ total)
{
	__shared__ ScanSmem sm;
	__shared__ int s_col[kCount == 1 ? kMaxBins + 1 : 1];
	const int tid = threadIdx.x;
	if (tid == 0) sm.tile = atomicAdd(ticket, 1u);
	if (kCount == 1)
		for (int b = tid; b <= gx; b += kSortThreads) s_col[b] = 0;
	if (n_dev) n = *n_dev;
	__syncthreads();
	const uint32_t tile = sm.tile;
	if ((uint64_t)tile * kSortItemsPerBlock >= n && !(tile == 0 && n == 0)) return;   // uniform
	const uint32_t base = tile * (uint32_t)kSortItemsPerBlock + tid * kSortItems; // blocked arrangement

	uint32_t v[kSortItems];
	uint32_t sum = 0;
#pragma unroll
	for (int k = 0; k < kSortItems; k++) {
		uint32_t i = base + k;
		v[k] = 0u;
		if (i < n) {
			const uint32_t g = order[i];
			if (kCount == 0) {
				v[k] = counts[g];
			} else if (kCount == 1) {
				if (counts[g] > 0u) {
					const uint2 rc = rect[g];
					const int x0 = (int)(rc.x & 0xFFFFu), x1 = (int)(rc.x >> 16);
					v[k] = (uint32_t)(x1 - x0);
					atomicAdd(&s_col[x0], 1);
					atomicAdd(&s_col[min(x1, gx)], -1);
					if (x1 > gx) {               // the part of a seam-wrapping rect that starts again at column 0
						atomicAdd(&s_col[0], 1);
						atomicAdd(&s_col[x1 - gx], -1);
					}
				}
			} else {
				const uint2 rc = rect[g];
				v[k] = (rc.y >> 16) - (rc.y & 0xFFFFu);
			}
		}
		sum += v[k];
	}
	uint32_t incl = sum;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) {
		uint32_t u = __shfl_up_sync(0xffffffffu, incl, o);
		if ((tid & 31) >= o) incl += u;
	}
	if ((tid & 31) == 31) sm.warp_tmp[tid >> 5] = incl;
	__syncthreads();
	if (kCount == 1)
		for (int b = tid; b <= gx; b += kSortThreads)
			if (s_col[b]) atomicAdd(&col_diff[b], s_col[b]);
	uint32_t warp_off = 0, tile_total = 0;
#pragma unroll
	for (int w = 0; w < kSortThreads / 32; w++) {
		uint32_t t = sm.warp_tmp[w];
		if (w < (tid >> 5)) warp_off += t;
		tile_total += t;
	}
	if (tid < 32) {   // warp 0 walks the predecessors together
		uint32_t excl = 0;
		if (tile == 0) {
			if (tid == 0) Status64::store(&status[0], Status64::inclusive(tile_total));
		} else {
			if (tid == 0) Status64::store(&status[tile], Status64::partial(tile_total));
			excl = lookback_sum_warp<Status64>(status, (int)tile, tid);
			if (tid == 0) Status64::store(&status[tile], Status64::inclusive(excl + tile_total));
		}
		if (tid == 0) sm.tile_excl = excl;
	}
	__syncthreads();
	const uint32_t R = (uint32_t)*total;
	uint32_t run = sm.tile_excl + warp_off + incl - sum;
#pragma unroll
	for (int k = 0; k < kSortItems; k++) {
		uint32_t i = base + k;
		if (i < n) {
			out[i] = run;
			if (v[k]) {
				// block boundaries b*kEmitPerBlock (and the last slot R-1) that fall inside [run, run + v[k])
				const uint32_t lo = run, hi = run + v[k];
				for (uint32_t b = (lo + kEmitPerBlock - 1) / kEmitPerBlock; (uint64_t)b * kEmitPerBlock < hi; b++)
					first_src[b] = i;
				if (hi == R) first_src[(R + kEmitPerBlock - 1) / kEmitPerBlock] = i;
			}
		}
		run += v[k];
		if (i == n - 1) out[n] = run;
	}
}

// x digit starts of the column-segment pass: difference array -> coverage counts -> exclusive prefix.  One block.
__global__ void __launch_bounds__(kMaxBins) col_starts_kernel(const int* __restrict__ col_diff, int gx, uint32_t* __restrict__ starts)
{
	__shared__ uint32_t s_w[kMaxBins / 32];
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	auto block_inclusive = [&](uint32_t v) {
		uint32_t incl = v;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) {
			const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o);
			if (lane >= o) incl += u;
		}
		__syncthreads();
		if (lane == 31) s_w[warp] = incl;
		__syncthreads();
		uint32_t off = 0;
		for (int w = 0; w < warp; w++) off += s_w[w];
		return incl + off;
	};
	const uint32_t cover = block_inclusive(tid < gx ? (uint32_t)col_diff[tid] : 0u);   // segments that cover column tid
	const uint32_t c = tid < gx ? cover : 0u;
	starts[tid] = block_inclusive(c) - c;
}

// ------------------------------------------------------------------ tile counts -> ranges + digit histograms
// Single block.  tile_diff is the (gy+1) x (gx+1) 2-D difference array (+1 at (y0,x0) and (y1,x1),
// -1 at (y0,x1) and (y1,x0) per Gaussian).  Produces per-tile counts (2-D prefix sum), `ranges`
// (their exclusive prefix in row-major tile order; empty tiles stay (0,0) like the reference's
// memset + identifyTileRanges) and the per-pass digit histograms of the tile-id sort.
__global__ void __launch_bounds__(1024) tile_ranges_kernel(
	int* __restrict__ tile_diff, int gx, int gy, uint32_t* __restrict__ tile_count, uint2* __restrict__ ranges,
	uint32_t* __restrict__ tile_hist, TileSortPlan plan)
{
	__shared__ uint32_t s_warp[32];
	__shared__ uint32_t s_hist[kMaxTilePasses * kMaxBins];
	const int tid = threadIdx.x, nthreads = blockDim.x;
	const int T = gx * gy, pitch = gx + 1;

	for (int i = tid; i < kMaxTilePasses * kMaxBins; i += nthreads) s_hist[i] = 0;
	// column prefix (down the rows)
	// (eight independent loads in flight per thread: the chain is L2-latency bound, not the adds)
	for (int x = tid; x < pitch; x += nthreads) {
		int run = 0;
		for (int y0 = 0; y0 <= gy; y0 += 8) {
			int v[8];
#pragma unroll
			for (int k = 0; k < 8; k++) v[k] = (y0 + k <= gy) ? __ldcg(&tile_diff[(y0 + k) * pitch + x]) : 0;
#pragma unroll
			for (int k = 0; k < 8; k++) {
				run += v[k];
				if (y0 + k <= gy) tile_diff[(y0 + k) * pitch + x] = run;
			}
		}
	}
	__syncthreads();
	// row prefix (along x), one warp per row at a time -> tile_count
	const int lane = tid & 31, warp = tid >> 5, nwarps = nthreads >> 5;
	for (int y = warp; y < gy; y += nwarps) {
		int carry = 0;
		for (int x0 = 0; x0 < gx; x0 += 32) {
			int x = x0 + lane;
			int v = (x < gx) ? tile_diff[y * pitch + x] : 0;
#pragma unroll
			for (int o = 1; o < 32; o <<= 1) {
				int u = __shfl_up_sync(0xffffffffu, v, o);
				if (lane >= o) v += u;
			}
			v += carry;
			if (x < gx) tile_count[y * gx + x] = (uint32_t)v;
			carry = __shfl_sync(0xffffffffu, v, 31);
		}
	}
	__syncthreads();
	// column-segment path: y digit starts = exclusive prefix of the rows' instance counts (row sums of tile_count)
	if (plan.segments) {
		__shared__ uint32_t s_row[kMaxBins];
		for (int y = warp; y < kMaxBins; y += nwarps) {
			uint32_t part = 0;
			if (y < gy)
				for (int x = lane; x < gx; x += 32) part += tile_count[y * gx + x];
#pragma unroll
			for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
			if (lane == 0) s_row[y] = part;
		}
		__syncthreads();
		if (warp == 0) {
			constexpr int kPerLane = kMaxBins / 32;
			uint32_t v[kPerLane], sum = 0;
#pragma unroll
			for (int k = 0; k < kPerLane; k++) {
				v[k] = s_row[lane * kPerLane + k];
				sum += v[k];
			}
			uint32_t incl = sum;
#pragma unroll
			for (int o = 1; o < 32; o <<= 1) {
				const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o);
				if (lane >= o) incl += u;
			}
			uint32_t run = incl - sum;
#pragma unroll
			for (int k = 0; k < kPerLane; k++) {
				tile_hist[kRowStartsRow * kMaxBins + lane * kPerLane + k] = run;
				run += v[k];
			}
		}
		__syncthreads();
	}
	// exclusive scan over T tiles (row-major): contiguous chunk per thread
	const int chunk = (T + nthreads - 1) / nthreads;
	const int begin = min(T, tid * chunk), end = min(T, begin + chunk);
	uint32_t sum = 0;
	for (int t = begin; t < end; t++) sum += tile_count[t];
	uint32_t incl = sum;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) {
		uint32_t u = __shfl_up_sync(0xffffffffu, incl, o);
		if (lane >= o) incl += u;
	}
	if (lane == 31) s_warp[warp] = incl;
	__syncthreads();
	uint32_t warp_off = 0;
	for (int w = 0; w < warp; w++) warp_off += s_warp[w];
	uint32_t run = warp_off + incl - sum;
	for (int t = begin; t < end; t++) {
		uint32_t c = tile_count[t];
		ranges[t] = c ? make_uint2(run, run + c) : make_uint2(0u, 0u);
		if (c) {
			for (int p = 0; p < plan.passes; p++)
				atomicAdd(&s_hist[p * kMaxBins + (((uint32_t)t >> plan.shift[p]) & ((1u << plan.bits[p]) - 1u))], c);
		}
		run += c;
	}
	__syncthreads();
	// the sort passes want digit STARTS (exclusive prefix of the counts): one warp per pass scans its kMaxBins counts
	// here once instead of every one of the passes' thousands of CTAs doing it
	if (warp < plan.passes) {   // (rows kColStartsRow / kRowStartsRow belong to the column-segment path: passes <= 3 there)
		constexpr int kPerLane = kMaxBins / 32;
		uint32_t* hp = s_hist + warp * kMaxBins;
		uint32_t v[kPerLane], sum = 0;
#pragma unroll
		for (int k = 0; k < kPerLane; k++) {
			v[k] = hp[lane * kPerLane + k];
			sum += v[k];
		}
		uint32_t incl = sum;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) {
			const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o);
			if (lane >= o) incl += u;
		}
		uint32_t run = incl - sum;
#pragma unroll
		for (int k = 0; k < kPerLane; k++) {
			tile_hist[warp * kMaxBins + lane * kPerLane + k] = run;
			run += v[k];
		}
	}
}

// ------------------------------------------------------------------ test-only: rebuild the 64-bit keys
__global__ void rebuild_keys_kernel(const uint2* __restrict__ ranges, int T, const uint32_t* __restrict__ point_list,
                                    const float* __restrict__ depth, unsigned long long* __restrict__ keys)
{
	int t = blockIdx.x;
	if (t >= T) return;
	uint2 r = ranges[t];
	for (uint32_t i = r.x + threadIdx.x; i < r.y; i += blockDim.x)
		keys[i] = ((unsigned long long)(uint32_t)t << 32) | (unsigned long long)__float_as_uint(depth[point_list[i]]);
}

// ------------------------------------------------------------------ host-side launch helpers
TileSortPlan make_tile_sort_plan(int W, int H)
{
	TileSortPlan p{};
	int gx = ceil_div(W, kTile), gy = ceil_div(H, kTile);
	p.bit = (int)higher_msb((uint32_t)(gx * gy));
	p.passes = (p.bit + 8) / 9;
	if (p.passes < 1) p.passes = 1;
	int per = (p.bit + p.passes - 1) / p.passes, s = 0;
	for (int i = 0; i < p.passes; i++) {
		p.shift[i] = s;
		p.bits[i] = (p.bit - s < per) ? (p.bit - s) : per;
		if (p.bits[i] < 1) p.bits[i] = 1;
		s += p.bits[i];
	}
	// column-segment path: needs x and y to be one digit each; pays off when the classic plan has two or more passes
	static const bool want = [] { const char* e = getenv("OGS_SEGMENT_SORT"); return e ? atoi(e) != 0 : true; }();
	p.bits_x = (int)higher_msb((uint32_t)(gx > 1 ? gx - 1 : 1));
	p.bits_y = (int)higher_msb((uint32_t)(gy > 1 ? gy - 1 : 1));
	p.segments = (want && p.passes >= 2 && gx <= kMaxBins && gy <= kMaxBins) ? 1 : 0;
	return p;
}

static cudaError_t ensure_onesweep_smem()
{
	// per device/context attribute, cheap host call: set on every use (several devices per process)
	const int bytes = (int)kOnesweepSmemBytes;
	cudaError_t e = cudaSuccess;
#define OGS_SET(B)                                                                                                          \
	if (e == cudaSuccess) e = cudaFuncSetAttribute(onesweep_pass_kernel<B, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes); \
	if (e == cudaSuccess) e = cudaFuncSetAttribute(onesweep_pass_kernel<B, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes); \
	if (e == cudaSuccess) e = cudaFuncSetAttribute(onesweep_pass_kernel<B, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes); \
	if (e == cudaSuccess) e = cudaFuncSetAttribute(onesweep_pass_kernel<B, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
	OGS_SET(1) OGS_SET(2) OGS_SET(3) OGS_SET(4) OGS_SET(5) OGS_SET(6) OGS_SET(7) OGS_SET(8) OGS_SET(9)
#undef OGS_SET
	return e;
}

// Depth ordering of the P Gaussians (stage 1).  Keys: g.sort_key[0]; result order in g.sort_val[0];
// then emit_offset (P+1) in that order.
int launch_depth_order(const GeomState& g, int P, cudaStream_t st)
{
	OGS_CUDA_TRY(ensure_onesweep_smem());
	const uint32_t n = (uint32_t)P;
	const int tiles = ceil_div(P, kSortItemsPerBlock);
	int hist_blocks = min(ceil_div(P, 256 * 8), kNumSMs * 4);
	if (hist_blocks < 1) hist_blocks = 1;
	depth_histogram_kernel<<<hist_blocks, 256, 0, st>>>(g.sort_key[0], n, g.depth_hist);
	unsigned int* tickets = reinterpret_cast<unsigned int*>(g.scalars + 2);
	for (int p = 0; p < 4; p++) {
		const uint32_t* kin = g.sort_key[p & 1];
		const uint32_t* vin = p == 0 ? nullptr : g.sort_val[p & 1];
		uint32_t* kout = g.sort_key[(p + 1) & 1];
		uint32_t* vout = g.sort_val[(p + 1) & 1];
		onesweep_pass_kernel<8><<<tiles, kSortThreads, kOnesweepSmemBytes, st>>>(
			kin, vin, kout, vout, n, 8 * p, g.depth_hist + 256 * p,
			g.depth_status + (size_t)p * tiles * 256, tickets + p, EmitSource{}, false, nullptr);
	}
	// 4 passes: result back in buffer 0
	OGS_CUDA_TRY(cudaGetLastError());
	return OGS_OK;
}

int launch_tile_ranges(const ImageState& img, int W, int H, cudaStream_t st)
{
	int gx = ceil_div(W, kTile), gy = ceil_div(H, kTile);
	tile_ranges_kernel<<<1, 1024, 0, st>>>(img.tile_diff, gx, gy, img.tile_count, img.ranges, img.tile_hist,
	                                       make_tile_sort_plan(W, H));
	OGS_CUDA_TRY(cudaGetLastError());
	return OGS_OK;
}

// Emission + tile-id sort (stage 2).  Leaves the sorted Gaussian list in b.point_list.
#define OGS_BITS_SWITCH(bits, CALL)                                                                     \
	switch (bits) {                                                                                     \
	case 1: CALL(1); break; case 2: CALL(2); break; case 3: CALL(3); break; case 4: CALL(4); break;     \
	case 5: CALL(5); break; case 6: CALL(6); break; case 7: CALL(7); break; case 8: CALL(8); break;     \
	case 9: CALL(9); break;                                                                             \
	default: return fail(OGS_ERR_INVALID_ARG, "bad radix digit width");                                 \
	}

int launch_emit_and_tile_sort(const GeomState& g, const ImageState& img, const BinningState& b,
                              int P, int64_t R, int W, int H, cudaStream_t st)
{
	if (R <= 0) return OGS_OK;
	OGS_CUDA_TRY(ensure_onesweep_smem());
	const int gx = ceil_div(W, kTile);
	const TileSortPlan plan = make_tile_sort_plan(W, H);
	const uint32_t n = (uint32_t)R;
	const int tiles = (int)((R + kSortItemsPerBlock - 1) / kSortItemsPerBlock);
	// ticket and look-back words of the emission-offset scan are zeroed HERE (not only in stage 1), so stage 2 may run
	// again on the same buffers
	const size_t scan_tiles = (size_t)ceil_div(P, kSortItemsPerBlock);
	unsigned int* scan_ticket = reinterpret_cast<unsigned int*>(g.scan_status + scan_tiles + 1);
	OGS_CUDA_TRY(cudaMemsetAsync(g.scan_status, 0, sizeof(unsigned long long) * (scan_tiles + 2), st));

	if (plan.segments) {
		// ---- column-segment path: scan widths -> x pass over segments -> scan heights -> y pass over instances ----
		const uint32_t* n_segments = g.emit_offset + P;            // device value S, written by the first scan
		uint32_t* seg_gid = b.key[0];                              // x-sorted segments: Gaussian ids
		uint32_t* seg_offset = b.key[1];                           // their emission offsets (S + 1)
		unsigned int* seg_scan_ticket = reinterpret_cast<unsigned int*>(b.seg_scan_status + tiles + 1);
		prof_begin(OGS_PROF_EMIT, st);
		gather_scan_kernel<1><<<(int)scan_tiles, kSortThreads, 0, st>>>(
			g.tiles_touched, g.sort_val[0], (uint32_t)P, nullptr, g.rect, gx, b.col_diff, g.emit_offset, g.scan_status,
			scan_ticket, b.first_src, g.scalars + 1);
		col_starts_kernel<<<1, kMaxBins, 0, st>>>(b.col_diff, gx, img.tile_hist + kColStartsRow * kMaxBins);
		prof_end(OGS_PROF_EMIT, st);
		prof_begin(OGS_PROF_TILE_SORT, st);
		const EmitSource ex{ g.emit_offset, g.sort_val[0], g.rect, b.first_src, gx };
#define OGS_X_PASS(B) onesweep_pass_kernel<B, 2><<<tiles, kSortThreads, kOnesweepSmemBytes, st>>>(                 \
		nullptr, nullptr, nullptr, seg_gid, n, 0, img.tile_hist + kColStartsRow * kMaxBins, b.status, b.tickets + 0, ex, true, n_segments)
		OGS_BITS_SWITCH(plan.bits_x, OGS_X_PASS)
#undef OGS_X_PASS
		gather_scan_kernel<2><<<tiles, kSortThreads, 0, st>>>(
			nullptr, seg_gid, 0u, n_segments, g.rect, gx, nullptr, seg_offset, b.seg_scan_status, seg_scan_ticket, b.first_src,
			g.scalars);
		const EmitSource ey{ seg_offset, seg_gid, g.rect, b.first_src, gx };
		uint32_t* status_y = b.status + ((size_t)tiles << plan.bits_x);
#define OGS_Y_PASS(B) onesweep_pass_kernel<B, 3><<<tiles, kSortThreads, kOnesweepSmemBytes, st>>>(                 \
		nullptr, nullptr, nullptr, b.point_list, n, 0, img.tile_hist + kRowStartsRow * kMaxBins, status_y, b.tickets + 1, ey, true, nullptr)
		OGS_BITS_SWITCH(plan.bits_y, OGS_Y_PASS)
#undef OGS_Y_PASS
		prof_end(OGS_PROF_TILE_SORT, st);
		OGS_CUDA_TRY(cudaGetLastError());
		return OGS_OK;
	}

	prof_begin(OGS_PROF_EMIT, st);
	gather_scan_kernel<0><<<(int)scan_tiles, kSortThreads, 0, st>>>(
		g.tiles_touched, g.sort_val[0], (uint32_t)P, nullptr, nullptr, gx, nullptr, g.emit_offset, g.scan_status, scan_ticket,
		b.first_src, g.scalars);
	prof_end(OGS_PROF_EMIT, st);
	prof_begin(OGS_PROF_TILE_SORT, st);
	const EmitSource es{ g.emit_offset, g.sort_val[0], g.rect, b.first_src, gx };
	size_t status_off = 0;
	for (int p = 0; p < plan.passes; p++) {
		const bool last = (p == plan.passes - 1);
#define OGS_SORT_ARGS                                                                                          \
	b.key[p & 1], b.val[p & 1], last ? nullptr : b.key[(p + 1) & 1], b.val[(p + 1) & 1], n, plan.shift[p],     \
		img.tile_hist + p * kMaxBins, b.status + status_off, b.tickets + p, es, true, nullptr
#define OGS_SORT_CASE(B)                                                                                      \
	if (p == 0) onesweep_pass_kernel<B, 1><<<tiles, kSortThreads, kOnesweepSmemBytes, st>>>(OGS_SORT_ARGS);   \
	else onesweep_pass_kernel<B, 0><<<tiles, kSortThreads, kOnesweepSmemBytes, st>>>(OGS_SORT_ARGS)
		OGS_BITS_SWITCH(plan.bits[p], OGS_SORT_CASE)
#undef OGS_SORT_CASE
#undef OGS_SORT_ARGS
		status_off += (size_t)tiles << plan.bits[p];
	}
	prof_end(OGS_PROF_TILE_SORT, st);
	OGS_CUDA_TRY(cudaGetLastError());
	return OGS_OK;
}

int launch_rebuild_keys(const ImageState& img, const BinningState& b, const GeomState& g, int W, int H,
                        uint64_t* keys, cudaStream_t st)
{
	int T = ceil_div(W, kTile) * ceil_div(H, kTile);
	rebuild_keys_kernel<<<T, 128, 0, st>>>(img.ranges, T, b.point_list, g.depth,
	                                       reinterpret_cast<unsigned long long*>(keys));
	OGS_CUDA_TRY(cudaGetLastError());
	return OGS_OK;
}

} // namespace ogs
