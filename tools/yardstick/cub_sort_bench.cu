// Yardstick only (never linked into the product): how fast is CUB's sm_100-tuned onesweep
// (CCCL 2.8.2 in CUDA 12.9) on the two sorts of this pipeline?
//   (a) the reference's sort: R (u64 key, u32 value) pairs, bits [0, 32+bit)   (rasterizer_impl.cu:656-661)
//   (b) our tile-id sort:     R (u32 key, u32 value) pairs, bits [0, bit)
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o cub_sort_bench cub_sort_bench.cu
#include <cub/cub.cuh>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <random>

template <typename K>
float run(size_t n, int end_bit, int tiles)
{
	std::vector<K> hk(n);
	std::vector<unsigned> hv(n);
	std::mt19937_64 rng(1);
	for (size_t i = 0; i < n; i++) {
		unsigned long long t = rng() % tiles;
		hk[i] = sizeof(K) == 8 ? (K)((t << 32) | (unsigned)(rng() & 0x7fffffff)) : (K)t;
		hv[i] = (unsigned)i;
	}
	K *k0, *k1; unsigned *v0, *v1;
	cudaMalloc(&k0, n * sizeof(K)); cudaMalloc(&k1, n * sizeof(K)); cudaMalloc(&v0, n * 4); cudaMalloc(&v1, n * 4);
	cudaMemcpy(k0, hk.data(), n * sizeof(K), cudaMemcpyHostToDevice);
	cudaMemcpy(v0, hv.data(), n * 4, cudaMemcpyHostToDevice);
	size_t tmp = 0; void* d = nullptr;
	cub::DeviceRadixSort::SortPairs(d, tmp, k0, k1, v0, v1, n, 0, end_bit);
	cudaMalloc(&d, tmp);
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	float best = 1e9f;
	for (int it = 0; it < 8; it++) {
		cudaEventRecord(e0);
		cub::DeviceRadixSort::SortPairs(d, tmp, k0, k1, v0, v1, n, 0, end_bit);
		cudaEventRecord(e1); cudaEventSynchronize(e1);
		float ms; cudaEventElapsedTime(&ms, e0, e1);
		if (it >= 2 && ms < best) best = ms;
	}
	cudaFree(k0); cudaFree(k1); cudaFree(v0); cudaFree(v1); cudaFree(d);
	return best;
}

int main(int argc, char** argv)
{
	size_t n = argc > 1 ? strtoull(argv[1], 0, 10) : 23262464ull;
	int bit = argc > 2 ? atoi(argv[2]) : 14;
	int tiles = argc > 3 ? atoi(argv[3]) : 8192;
	printf("R=%zu bit=%d\n", n, bit);
	printf("cub u64/u32 pairs, bits [0,%d): %.3f ms\n", 32 + bit, run<unsigned long long>(n, 32 + bit, tiles));
	printf("cub u32/u32 pairs, bits [0,%d): %.3f ms\n", bit, run<unsigned>(n, bit, tiles));
	return 0;
}
