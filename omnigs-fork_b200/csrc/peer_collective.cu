// peer_collective.cu — all-reduce(SUM) of the data-parallel gradient bucket over NVLink peer memory.
//
// Every rank owns one contiguous slice of the bucket.  One kernel per rank reads that slice from ALL ranks'
// buffers (its own and the peers', mapped into this process: peer loads over NVLink / NVSwitch), adds them in
// rank order and stores the sum into ALL ranks' buffers (peer stores).  Per GPU that is (N-1)/N of the bucket in
// and (N-1)/N out, both directions of the links busy at once — a reduce-scatter and an all-gather in one pass —
// and because the owner adds in a fixed order, every replica ends with bit-identical sums.
// Ordering across ranks (everybody's gradients written before anyone reads; everybody's sums written before
// anyone consumes) is the caller's: parallel.py brackets the launch with the symmetric-memory barrier of
// torch.distributed on the same stream.  No spin-waits in this kernel.
#include "launchers.cuh"
#include <cstdlib>

namespace ogs {

struct PeerBuffers {
	float* buf[kMaxPeers];
};

constexpr int kPeerUnroll = 4;   // independent 16-byte peer loads in flight per thread

template <bool kMax>
__global__ void __launch_bounds__(256) peer_allreduce_kernel(const PeerBuffers p, int world, size_t begin4, size_t end4)
{
	// float4 units; each block walks chunks of 256 * kPeerUnroll consecutive float4 of this rank's slice [begin4, end4)
	const size_t stride = (size_t)gridDim.x * blockDim.x * kPeerUnroll;
	for (size_t base = begin4 + (size_t)blockIdx.x * blockDim.x * kPeerUnroll + threadIdx.x; base < end4; base += stride) {
		float4 acc[kPeerUnroll];
#pragma unroll
		for (int u = 0; u < kPeerUnroll; u++) {
			const size_t i = base + (size_t)u * blockDim.x;
			acc[u] = (i < end4) ? reinterpret_cast<const float4*>(p.buf[0])[i] : make_float4(0.f, 0.f, 0.f, 0.f);
		}
#pragma unroll
		for (int r = 1; r < kMaxPeers; r++) {
			if (r < world) {
				float4 v[kPeerUnroll];
#pragma unroll
				for (int u = 0; u < kPeerUnroll; u++) {
					const size_t i = base + (size_t)u * blockDim.x;
					v[u] = (i < end4) ? reinterpret_cast<const float4*>(p.buf[r])[i] : make_float4(0.f, 0.f, 0.f, 0.f);
				}
#pragma unroll
				for (int u = 0; u < kPeerUnroll; u++) {
					if (kMax) {
						acc[u].x = fmaxf(acc[u].x, v[u].x); acc[u].y = fmaxf(acc[u].y, v[u].y);
						acc[u].z = fmaxf(acc[u].z, v[u].z); acc[u].w = fmaxf(acc[u].w, v[u].w);
					} else {
						acc[u].x += v[u].x; acc[u].y += v[u].y; acc[u].z += v[u].z; acc[u].w += v[u].w;
					}
				}
			}
		}
#pragma unroll
		for (int r = 0; r < kMaxPeers; r++) {
			if (r < world) {
#pragma unroll
				for (int u = 0; u < kPeerUnroll; u++) {
					const size_t i = base + (size_t)u * blockDim.x;
					if (i < end4) reinterpret_cast<float4*>(p.buf[r])[i] = acc[u];
				}
			}
		}
	}
}

// NVLink SHARP variant: the bucket's multicast address reaches all ranks' copies at once.  multimem.ld_reduce has the
// switch add the N copies of an element and return the sum (one inbound copy per element instead of N - 1),
// multimem.st broadcasts the result to every rank (one outbound copy).  Each rank handles its own slice.
__global__ void __launch_bounds__(256) multimem_allreduce_sum_kernel(float* mc, size_t begin4, size_t end4)
{
	const size_t stride = (size_t)gridDim.x * blockDim.x * kPeerUnroll;
	for (size_t base = begin4 + (size_t)blockIdx.x * blockDim.x * kPeerUnroll + threadIdx.x; base < end4; base += stride) {
		float4 acc[kPeerUnroll];
#pragma unroll
		for (int u = 0; u < kPeerUnroll; u++) {
			const size_t i = base + (size_t)u * blockDim.x;
			if (i < end4)
				asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
				             : "=f"(acc[u].x), "=f"(acc[u].y), "=f"(acc[u].z), "=f"(acc[u].w)
				             : "l"(reinterpret_cast<float4*>(mc) + i) : "memory");
		}
#pragma unroll
		for (int u = 0; u < kPeerUnroll; u++) {
			const size_t i = base + (size_t)u * blockDim.x;
			if (i < end4)
				asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};"
				             :: "l"(reinterpret_cast<float4*>(mc) + i), "f"(acc[u].x), "f"(acc[u].y), "f"(acc[u].z), "f"(acc[u].w) : "memory");
		}
	}
}

// max over ranks of non-negative floats (the per-view radii): they order like their bit patterns, and the switch reduces
// unsigned integers (multimem.ld_reduce has no f32 max), one element per thread
__global__ void __launch_bounds__(256) multimem_allreduce_max_kernel(float* mc, size_t begin, size_t end)
{
	for (size_t i = begin + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < end; i += (size_t)gridDim.x * blockDim.x) {
		unsigned v;
		asm volatile("multimem.ld_reduce.relaxed.sys.global.max.u32 %0, [%1];" : "=r"(v) : "l"(mc + i) : "memory");
		asm volatile("multimem.st.relaxed.sys.global.u32 [%0], %1;" :: "l"(mc + i), "r"(v) : "memory");
	}
}

// Latitude bands: every rank owns the pixel rows [y0, y1) of the frame; one kernel per rank stores its
// rows of the three colour planes into every rank's image (peer stores over NVLink) — an all-gather of band rows that
// moves each pixel once per peer, where summing zero-padded full frames moved every pixel of every frame.
__global__ void __launch_bounds__(256) band_rows_allgather_kernel(const PeerBuffers dst, int world, const float* __restrict__ src,
                                                                  size_t plane, size_t begin, size_t count)
{
	// elements [begin, begin + count) of each of the 3 planes; vectorised when the range is 16-byte aligned
	const bool vec = ((begin | count | plane) & 3u) == 0;
	const size_t n = vec ? count / 4 : count;
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < 3 * n; i += (size_t)gridDim.x * blockDim.x) {
		const size_t c = i / n, k = i - c * n;
		if (vec) {
			const size_t off = (c * plane + begin) / 4 + k;
			const float4 v = reinterpret_cast<const float4*>(src)[off];
#pragma unroll
			for (int r = 0; r < kMaxPeers; r++)
				if (r < world) reinterpret_cast<float4*>(dst.buf[r])[off] = v;
		} else {
			const size_t off = c * plane + begin + k;
			const float v = src[off];
#pragma unroll
			for (int r = 0; r < kMaxPeers; r++)
				if (r < world) dst.buf[r][off] = v;
		}
	}
}

int launch_band_rows_allgather(float* const* images, int world, const float* src, int W, int H, int y0, int y1, cudaStream_t st)
{
	if (y1 <= y0) return OGS_OK;
	PeerBuffers p{};
	for (int r = 0; r < world; r++) p.buf[r] = images[r];
	const size_t plane = (size_t)W * H, begin = (size_t)y0 * W, count = (size_t)(y1 - y0) * W;
	const size_t want = (3 * count / 4 + 255) / 256 + 1;
	band_rows_allgather_kernel<<<(int)min(want, (size_t)kNumSMs * 8), 256, 0, st>>>(p, world, src, plane, begin, count);
	OGS_CUDA_TRY(cudaGetLastError());
	return OGS_OK;
}

int launch_multimem_allreduce_sum(float* multicast, int world, int rank, size_t count, cudaStream_t st)
{
	const size_t n4 = count / 4;
	const size_t per = (n4 + world - 1) / world;
	const size_t begin4 = min(n4, per * rank), end4 = min(n4, begin4 + per);
	if (end4 <= begin4) return OGS_OK;
	const size_t want = (end4 - begin4 + 256 * kPeerUnroll - 1) / (256 * kPeerUnroll);
	const int blocks = (int)min(want, (size_t)kNumSMs * 8);
	multimem_allreduce_sum_kernel<<<blocks, 256, 0, st>>>(multicast, begin4, end4);
	OGS_CUDA_TRY(cudaGetLastError());
	return OGS_OK;
}

int launch_peer_allreduce_sum(float* const* bufs, int world, int rank, size_t count, cudaStream_t st)
{
	PeerBuffers p{};
	for (int r = 0; r < world; r++) p.buf[r] = bufs[r];
	const size_t n4 = count / 4;                       // the bucket is padded to a multiple of 4 floats
	const size_t per = (n4 + world - 1) / world;
	const size_t begin4 = min(n4, per * rank), end4 = min(n4, begin4 + per);
	if (end4 <= begin4) return OGS_OK;
	const size_t want = (end4 - begin4 + 256 * kPeerUnroll - 1) / (256 * kPeerUnroll);
	static const int per_sm = [] { const char* e = getenv("OGS_PEER_BLOCKS_PER_SM"); return e ? atoi(e) : 8; }();
	const int blocks = (int)min(want, (size_t)kNumSMs * per_sm);
	peer_allreduce_kernel<false><<<blocks, 256, 0, st>>>(p, world, begin4, end4);
	OGS_CUDA_TRY(cudaGetLastError());
	return OGS_OK;
}

int launch_peer_allreduce(float* const* bufs, int world, int rank, size_t count_sum, size_t count_max, cudaStream_t st)
{
	if (int rc = launch_peer_allreduce_sum(bufs, world, rank, count_sum, st)) return rc;
	if (count_max == 0) return OGS_OK;
	PeerBuffers p{};
	for (int r = 0; r < world; r++) p.buf[r] = bufs[r] + count_sum;
	const size_t n4 = count_max / 4;
	const size_t per = (n4 + world - 1) / world;
	const size_t begin4 = min(n4, per * rank), end4 = min(n4, begin4 + per);
	if (end4 <= begin4) return OGS_OK;
	const size_t want = (end4 - begin4 + 256 * kPeerUnroll - 1) / (256 * kPeerUnroll);
	peer_allreduce_kernel<true><<<(int)min(want, (size_t)kNumSMs * 4), 256, 0, st>>>(p, world, begin4, end4);
	OGS_CUDA_TRY(cudaGetLastError());
	return OGS_OK;
}

int launch_multimem_allreduce(float* multicast, int world, int rank, size_t count_sum, size_t count_max, cudaStream_t st)
{
	if (int rc = launch_multimem_allreduce_sum(multicast, world, rank, count_sum, st)) return rc;
	if (count_max == 0) return OGS_OK;
	const size_t per = (count_max + world - 1) / world;
	const size_t begin = min(count_max, per * rank), end = min(count_max, begin + per);
	if (end <= begin) return OGS_OK;
	const size_t want = (end - begin + 255) / 256;
	multimem_allreduce_max_kernel<<<(int)min(want, (size_t)kNumSMs * 8), 256, 0, st>>>(multicast + count_sum, begin, end);
	OGS_CUDA_TRY(cudaGetLastError());
	return OGS_OK;
}

} // namespace ogs
