// tma_bench.cu — development microbenchmark: sustained global -> shared ingest rate of 1-D bulk async copies
// (cp.async.bulk, UBLKCP) per SM as a function of copy size, copies in flight and working-set size (L2 vs HBM),
// against plain 128-bit loads.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_bench tma_bench.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../openmps_b200/csrc/mps_async.cuh"
using namespace mps::async;

__global__ void __launch_bounds__(128, 1) k_bulk(const unsigned char* src, size_t span, uint32_t bytes, uint32_t depth, uint32_t iters, unsigned long long* out)
{
	extern __shared__ __align__(128) unsigned char smem[];
	uint64_t* bar = reinterpret_cast<uint64_t*>(smem + (size_t)depth * bytes);
	if (threadIdx.x == 0) { for (uint32_t s = 0; s < depth; s++) mbar_init(&bar[s], 1); mbar_init_fence(); }
	__syncthreads();
	if (threadIdx.x == 0)
	{
		const uint64_t pol = policy_evict_normal();
		size_t off = ((size_t)blockIdx.x * 7919u * bytes) % (span - bytes);
		off &= ~(size_t)127;
		const long long t0 = clock64();
		for (uint32_t i = 0; i < iters + depth; i++)
		{
			const uint32_t s = i % depth;
			if (i >= depth) mbar_wait(&bar[s], ((i / depth) - 1) & 1);
			if (i < iters)
			{
				mbar_arrive_expect_tx(&bar[s], bytes);
				bulk_g2s(smem + (size_t)s * bytes, src + off, bytes, &bar[s], pol);
				off += (size_t)gridDim.x * bytes; if (off + bytes > span) off = (off + bytes) % (span - bytes) & ~(size_t)127;
			}
		}
		out[blockIdx.x] = clock64() - t0;
	}
}

__global__ void __launch_bounds__(512, 1) k_ldg(const uint4* src, size_t n16, uint32_t iters, unsigned long long* out, uint4* sink)
{
	uint4 acc = make_uint4(0, 0, 0, 0);
	size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	const size_t stride = (size_t)gridDim.x * blockDim.x;
	const long long t0 = clock64();
	for (uint32_t k = 0; k < iters; k++)
	{
#pragma unroll
		for (int u = 0; u < 8; u++) { const uint4 v = __ldcg(src + (i % n16)); acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w; i += stride; }
	}
	if (threadIdx.x == 0) out[blockIdx.x] = clock64() - t0;
	if (acc.x == 0x12345678) sink[0] = acc;
}

// several issuing threads per SM (lane 0 of each of `nw` warps, own stages and barriers each): is the ~420 cycles per copy a limit of
// the issuing thread or of the SM's copy engine?
__global__ void __launch_bounds__(512, 1) k_bulk_multi(const unsigned char* src, size_t span, uint32_t bytes, uint32_t depth, uint32_t iters, uint32_t nw, unsigned long long* out)
{
	extern __shared__ __align__(128) unsigned char smem[];
	const uint32_t w = threadIdx.x / 32;
	uint64_t* bar = reinterpret_cast<uint64_t*>(smem + (size_t)nw * depth * bytes) + w * depth;
	if (threadIdx.x % 32 == 0 && w < nw) { for (uint32_t s = 0; s < depth; s++) mbar_init(&bar[s], 1); mbar_init_fence(); }
	__syncthreads();
	const long long t0 = clock64();
	if (threadIdx.x % 32 == 0 && w < nw)
	{
		const uint64_t pol = policy_evict_normal();
		unsigned char* base = smem + (size_t)w * depth * bytes;
		size_t off = (((size_t)blockIdx.x * nw + w) * 7919u * bytes) % (span - bytes);
		off &= ~(size_t)127;
		for (uint32_t i = 0; i < iters + depth; i++)
		{
			const uint32_t s = i % depth;
			if (i >= depth) mbar_wait(&bar[s], ((i / depth) - 1) & 1);
			if (i < iters)
			{
				mbar_arrive_expect_tx(&bar[s], bytes);
				bulk_g2s(base + (size_t)s * bytes, src + off, bytes, &bar[s], pol);
				off += (size_t)gridDim.x * nw * bytes; if (off + bytes > span) off = (off + bytes) % (span - bytes) & ~(size_t)127;
			}
		}
	}
	__syncthreads();
	if (threadIdx.x == 0) out[blockIdx.x] = clock64() - t0;
}

int main()
{
	const size_t big = 1ull << 30;
	unsigned char* d; cudaMalloc(&d, big); cudaMemset(d, 1, big);
	unsigned long long* out; cudaMalloc(&out, 148 * 8); unsigned long long h[148];
	uint4* sink; cudaMalloc(&sink, 64);
	int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
	printf("clock %d kHz\n", clk);
	cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
	const size_t spans[2] = { 48ull << 20, big };
	for (int sp = 0; sp < 2; sp++)
		for (uint32_t bytes : { 2048u, 8192u, 16384u, 32768u, 65536u })
			for (uint32_t depth : { 1u, 2u, 3u, 6u, 12u })
			{
				if ((size_t)bytes * depth > 200 * 1024) continue;
				const uint32_t iters = (8u << 20) / bytes;
				cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
				k_bulk<<<148, 128, (size_t)bytes * depth + 256>>>(d, spans[sp], bytes, depth, iters, out); // warm (L2)
				cudaEventRecord(a);
				k_bulk<<<148, 128, (size_t)bytes * depth + 256>>>(d, spans[sp], bytes, depth, iters, out);
				cudaEventRecord(b); cudaEventSynchronize(b);
				float ms; cudaEventElapsedTime(&ms, a, b);
				cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
				double cyc = 0; for (int i = 0; i < 148; i++) cyc += h[i]; cyc /= 148;
				printf("bulk span %5zu MB  copy %6u B  depth %2u : %6.1f B/cycle/SM  %7.1f GB/s total (%s)\n", spans[sp] >> 20, bytes, depth,
					(double)bytes * iters / cyc, 148.0 * bytes * iters / (ms * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
			}
	cudaFuncSetAttribute(k_bulk_multi, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
	for (uint32_t bytes : { 256u, 1024u, 2048u, 8192u })
		for (uint32_t nw : { 1u, 2u, 4u, 8u })
			for (uint32_t depth : { 4u, 8u })
			{
				if ((size_t)bytes * depth * nw > 200 * 1024) continue;
				const uint32_t iters = 4096;
				const size_t sm = (size_t)bytes * depth * nw + 8 * depth * nw + 256;
				k_bulk_multi<<<148, 512, sm>>>(d, spans[0], bytes, depth, iters, nw, out);
				k_bulk_multi<<<148, 512, sm>>>(d, spans[0], bytes, depth, iters, nw, out);
				cudaDeviceSynchronize();
				cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
				double cyc = 0; for (int i = 0; i < 148; i++) cyc += h[i]; cyc /= 148;
				printf("multi span 48 MB  copy %6u B  issuers %u  depth %2u : %7.1f cycles/copy/SM  %6.1f B/cycle/SM (%s)\n", bytes, nw, depth,
					cyc / ((double)iters * nw), (double)bytes * iters * nw / cyc, cudaGetErrorString(cudaGetLastError()));
			}
	for (int sp = 0; sp < 2; sp++)
	{
		const uint32_t iters = 256;
		cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
		k_ldg<<<148, 512>>>((const uint4*)d, spans[sp] / 16, iters, out, sink);
		cudaEventRecord(a);
		k_ldg<<<148, 512>>>((const uint4*)d, spans[sp] / 16, iters, out, sink);
		cudaEventRecord(b); cudaEventSynchronize(b);
		float ms; cudaEventElapsedTime(&ms, a, b);
		printf("ldg.128 span %5zu MB : %7.1f GB/s total\n", spans[sp] >> 20, 148.0 * 512 * iters * 8 * 16 / (ms * 1e-3) / 1e9);
	}
	return 0;
}
