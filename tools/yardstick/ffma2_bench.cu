// Yardstick only (never linked into the product): issue rate of scalar FFMA against packed FFMA2 (fma.rn.f32x2) on
// sm_100a, alone and mixed with ALU work, to read the packed render_bwd variant's result against.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o ffma2_bench ffma2_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ float fma1(float a, float b, float c) { float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }

constexpr int kIters = 4096;

// 16 independent chains of scalar FMAs per thread (16 flop-pairs per iteration)
__global__ void scalar_kernel(float* out, float a, float b)
{
	float v[16];
	for (int i = 0; i < 16; i++) v[i] = threadIdx.x + i;
	for (int it = 0; it < kIters; it++)
#pragma unroll
		for (int i = 0; i < 16; i++) v[i] = fma1(v[i], a, b);
	float s = 0;
	for (int i = 0; i < 16; i++) s += v[i];
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// 8 independent chains of packed FMAs per thread (the same 16 flop-pairs per iteration)
__global__ void packed_kernel(float* out, float a, float b)
{
	f32x2 v[8];
	for (int i = 0; i < 8; i++) v[i] = pack2(threadIdx.x + i, threadIdx.x - i);
	const f32x2 a2 = pack2(a, a), b2 = pack2(b, b);
	for (int it = 0; it < kIters; it++)
#pragma unroll
		for (int i = 0; i < 8; i++) v[i] = fma2(v[i], a2, b2);
	float s = 0;
	for (int i = 0; i < 8; i++) { float lo, hi; unpack2(v[i], lo, hi); s += lo + hi; }
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// the same FMA work plus 16 integer ALU instructions per iteration (an issue-bound mix)
__global__ void scalar_mix_kernel(float* out, float a, float b, unsigned k)
{
	float v[16]; unsigned u[16];
	for (int i = 0; i < 16; i++) { v[i] = threadIdx.x + i; u[i] = threadIdx.x * 3 + i; }
	for (int it = 0; it < kIters; it++)
#pragma unroll
		for (int i = 0; i < 16; i++) { v[i] = fma1(v[i], a, b); u[i] = (u[i] ^ k) + (u[i] >> 3); }
	float s = 0; unsigned t = 0;
	for (int i = 0; i < 16; i++) { s += v[i]; t += u[i]; }
	out[blockIdx.x * blockDim.x + threadIdx.x] = s + (float)t;
}
__global__ void packed_mix_kernel(float* out, float a, float b, unsigned k)
{
	f32x2 v[8]; unsigned u[16];
	for (int i = 0; i < 8; i++) v[i] = pack2(threadIdx.x + i, threadIdx.x - i);
	for (int i = 0; i < 16; i++) u[i] = threadIdx.x * 3 + i;
	const f32x2 a2 = pack2(a, a), b2 = pack2(b, b);
	for (int it = 0; it < kIters; it++) {
#pragma unroll
		for (int i = 0; i < 8; i++) v[i] = fma2(v[i], a2, b2);
#pragma unroll
		for (int i = 0; i < 16; i++) u[i] = (u[i] ^ k) + (u[i] >> 3);
	}
	float s = 0; unsigned t = 0;
	for (int i = 0; i < 8; i++) { float lo, hi; unpack2(v[i], lo, hi); s += lo + hi; }
	for (int i = 0; i < 16; i++) t += u[i];
	out[blockIdx.x * blockDim.x + threadIdx.x] = s + (float)t;
}

template <typename F>
float time_ms(F launch)
{
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	launch(); cudaDeviceSynchronize();
	cudaEventRecord(e0);
	for (int r = 0; r < 5; r++) launch();
	cudaEventRecord(e1); cudaEventSynchronize(e1);
	float ms; cudaEventElapsedTime(&ms, e0, e1);
	return ms / 5;
}

int main()
{
	const int blocks = 148 * 8, threads = 256;
	float* out; cudaMalloc(&out, sizeof(float) * blocks * threads);
	const double fma_pairs = (double)blocks * threads * kIters * 16;   // scalar-equivalent FMAs per launch
	float t;
	t = time_ms([&] { scalar_kernel<<<blocks, threads>>>(out, 1.0001f, 0.5f); });
	printf("scalar FFMA        : %.3f ms  %.2f TFMA/s\n", t, fma_pairs / t * 1e-9);
	t = time_ms([&] { packed_kernel<<<blocks, threads>>>(out, 1.0001f, 0.5f); });
	printf("packed FFMA2       : %.3f ms  %.2f TFMA/s\n", t, fma_pairs / t * 1e-9);
	t = time_ms([&] { scalar_mix_kernel<<<blocks, threads>>>(out, 1.0001f, 0.5f, 12345u); });
	printf("scalar FFMA + ALU  : %.3f ms  %.2f TFMA/s\n", t, fma_pairs / t * 1e-9);
	t = time_ms([&] { packed_mix_kernel<<<blocks, threads>>>(out, 1.0001f, 0.5f, 12345u); });
	printf("packed FFMA2 + ALU : %.3f ms  %.2f TFMA/s\n", t, fma_pairs / t * 1e-9);
	return 0;
}
