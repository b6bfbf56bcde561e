// mps_async.cuh — thin inline-PTX wrappers for the sm_100a asynchronous-copy machinery used by the CG kernel:
// mbarrier (shared-memory transaction barriers), 1-D bulk async copies global -> shared (the TMA engine's linear mode,
// SASS UBLKCP), L2 cache policies and the generic<->async proxy fence.  No library dependency.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace mps {
namespace async {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
// make the initialised barriers visible to the async proxy before the first bulk copy signals them
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity)
{
	uint32_t ok;
	asm volatile(
		"{\n\t.reg .pred p;\n\t"
		"mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
		"selp.u32 %0, 1, 0, p;\n\t}"
		: "=r"(ok)
		: "r"(smem_u32(bar)), "r"(parity)
		: "memory");
	return ok != 0;
}
// Bounded wait: a protocol bug must end in a trapped launch (an error the host sees), never in a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
	for (uint32_t spin = 0; !mbar_try_wait(bar, parity); spin++)
	{
		if (spin > (1u << 26)) __trap();
	}
}

__device__ __forceinline__ uint64_t policy_evict_first()
{
	uint64_t p;
	asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
	return p;
}
__device__ __forceinline__ uint64_t policy_evict_last()
{
	uint64_t p;
	asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
	return p;
}
__device__ __forceinline__ uint64_t policy_evict_normal()
{
	uint64_t p;
	asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
	return p;
}

// dst (shared), src (global) 16-byte aligned, bytes a multiple of 16; completion is signalled on `bar` as `bytes` tx
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t policy)
{
	asm volatile(
		"cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
		::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
		: "memory");
}

// orders generic-proxy accesses (ordinary loads/stores, here: other CTAs' global stores made visible by a grid barrier)
// before subsequent async-proxy accesses (bulk copies) of the executing thread
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }

__device__ __forceinline__ unsigned long long ld_acquire_gpu(const unsigned long long* p)
{
	unsigned long long v;
	asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ void red_release_gpu_add(unsigned long long* p, unsigned long long v)
{
	asm volatile("red.release.gpu.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

} // namespace async
} // namespace mps
