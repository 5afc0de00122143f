// async_copy.cuh — bulk asynchronous copies (the 1-D TMA path: cp.async.bulk / UBLKCP) with
// mbarrier completion, used to stage the 192-byte-per-Gaussian SH rows.
//
// A thread-per-Gaussian kernel that reads its own SH row walks global memory with a 192-byte stride:
// every warp-level request touches 32 different lines and the LSU, not HBM, limits throughput
// (ncu: ~20 % of DRAM peak).  The rows of a CTA are contiguous in memory, so each thread instead asks the
// copy engine for its row; rows land in shared memory at a 208-byte pitch (13 x 16 B), which makes the
// per-thread 128-bit shared loads conflict-free (8 consecutive threads -> 8 different 16-byte banks).
#pragma once
#include "ogs_common.cuh"

namespace ogs {

OGS_D uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

OGS_D void mbar_init(uint64_t* bar, uint32_t arrivals)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(arrivals));
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); // make the init visible to the async proxy
}
OGS_D void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
OGS_D void mbar_arrive(uint64_t* bar)
{
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
OGS_D void mbar_wait(uint64_t* bar, uint32_t parity)
{
	asm volatile(
		"{\n"
		".reg .pred p;\n"
		"WAIT_%=:\n"
		"mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
		"@p bra DONE_%=;\n"
		"bra WAIT_%=;\n"
		"DONE_%=:\n"
		"}\n" ::"r"(smem_addr(bar)), "r"(parity) : "memory");
}
// global -> shared, completion counted in bytes on `bar`.  dst, src and bytes must be multiples of 16.
OGS_D void bulk_load(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
	             ::"r"(smem_addr(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}
// shared -> global (bulk-group completion).  Call fence_async_smem() after the generic-proxy writes.
OGS_D void bulk_store(void* gmem_dst, const void* smem_src, uint32_t bytes)
{
	asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
	             ::"l"(gmem_dst), "r"(smem_addr(smem_src)), "r"(bytes) : "memory");
}
OGS_D void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
OGS_D void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
OGS_D void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

constexpr int kShRowFloats = 48;     // 16 coefficients x RGB
constexpr int kShPitchFloats = 52;   // 208-byte pitch
// raw-parameter mode: the CTA's features_rest_ rows (45 floats each, contiguous) at the start of the
// staging area, its features_dc_ rows (3 floats each) behind them
constexpr int kRawRestFloats = 45;
constexpr int kRawDcOffset = 128 * kRawRestFloats;   // 23040 B: 16-byte aligned

// Host-side check: can the SH rows of this call go through the bulk path?
inline bool sh_rows_bulk_capable(const void* base, int M)
{
	return M == 16 && (reinterpret_cast<uintptr_t>(base) & 15u) == 0;
}

} // namespace ogs
