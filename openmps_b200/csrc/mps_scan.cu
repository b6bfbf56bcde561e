// mps_scan.cu — device-wide exclusive prefix sum (u32 counts -> u64 offsets, n + 1 outputs), hand-written.
//
// Used per step for: particles-per-cell -> cell_start (the reference's per-cell bucket, Grid.hpp:76,276-331, becomes a
// start/end table), neighbour counts -> neighbour row pointers (Computer.hpp:594-612 keeps a fixed-stride table instead), PPE
// row lengths -> CSR row pointers (replaces the serial uBLAS insertion, Computer.hpp:1337-1349), the chunk tables (mps_chunk.cu)
// and the occupied-cell ranks of every level of the preconditioner's hierarchy (mps_mg.cu).
// Long inputs: three passes (tile sums, scan of tile sums, tile scans): 2 reads + 1 write of the input, HBM-bound and tiny next
// to CG.  Short inputs (<= 32 768 items): one block, one launch (k_scan_small).
#include "mps_solver.h"

namespace mps {
namespace {

constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

__device__ __forceinline__ uint64_t warp_inclusive(uint64_t v)
{
	const unsigned lane = threadIdx.x & 31;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1)
	{
		const uint64_t u = __shfl_up_sync(0xffffffffu, v, o);
		if (lane >= o) v += u;
	}
	return v;
}

// exclusive scan of one value per thread over the block; returns the exclusive prefix, *total gets the block sum
template<int THREADS>
__device__ __forceinline__ uint64_t block_exclusive(uint64_t v, uint64_t* total)
{
	__shared__ uint64_t warp_sums[THREADS / 32];
	__shared__ uint64_t block_total;
	const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	const uint64_t inc = warp_inclusive(v);
	if (lane == 31) warp_sums[wid] = inc;
	__syncthreads();
	if (wid == 0)
	{
		uint64_t w = (lane < THREADS / 32) ? warp_sums[lane] : 0;
		const uint64_t winc = warp_inclusive(w);
		if (lane < THREADS / 32) warp_sums[lane] = winc - w;
		if (lane == THREADS / 32 - 1) block_total = winc;
	}
	__syncthreads();
	const uint64_t r = warp_sums[wid] + inc - v;
	if (total) *total = block_total;
	__syncthreads();
	return r;
}

__global__ void __launch_bounds__(kScanThreads) k_tile_sums(const uint32_t* __restrict__ in, uint64_t n, uint64_t* __restrict__ sums)
{
	const uint64_t base = static_cast<uint64_t>(blockIdx.x) * kScanTile;
	uint64_t v = 0;
#pragma unroll
	for (int k = 0; k < kScanItems; k++)
	{
		const uint64_t i = base + static_cast<uint64_t>(k) * kScanThreads + threadIdx.x; // coalesced
		if (i < n) v += in[i];
	}
	uint64_t total;
	block_exclusive<kScanThreads>(v, &total);
	if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

// one block: exclusive scan of the tile sums in place; sums[nb] = grand total
__global__ void __launch_bounds__(1024) k_scan_sums(uint64_t* __restrict__ sums, uint64_t nb)
{
	__shared__ uint64_t carry_s;
	if (threadIdx.x == 0) carry_s = 0;
	__syncthreads();
	for (uint64_t base = 0; base < nb; base += 1024)
	{
		const uint64_t i = base + threadIdx.x;
		const uint64_t v = (i < nb) ? sums[i] : 0;
		uint64_t total;
		const uint64_t ex = block_exclusive<1024>(v, &total);
		const uint64_t carry = carry_s;
		if (i < nb) sums[i] = carry + ex;
		__syncthreads();
		if (threadIdx.x == 0) carry_s = carry + total;
		__syncthreads();
	}
	if (threadIdx.x == 0) sums[nb] = carry_s;
}

__global__ void __launch_bounds__(kScanThreads) k_tile_scan(const uint32_t* __restrict__ in, uint64_t n, const uint64_t* __restrict__ sums,
	uint64_t* __restrict__ out, uint64_t nb)
{
	// thread t owns items [t*8, t*8+8) of the tile so that its 8 outputs are consecutive
	const uint64_t base = static_cast<uint64_t>(blockIdx.x) * kScanTile + static_cast<uint64_t>(threadIdx.x) * kScanItems;
	uint32_t item[kScanItems];
	uint64_t v = 0;
#pragma unroll
	for (int k = 0; k < kScanItems; k++)
	{
		const uint64_t i = base + k;
		item[k] = (i < n) ? in[i] : 0u;
		v += item[k];
	}
	uint64_t run = block_exclusive<kScanThreads>(v, nullptr) + sums[blockIdx.x];
#pragma unroll
	for (int k = 0; k < kScanItems; k++)
	{
		const uint64_t i = base + k;
		if (i < n) out[i] = run;
		run += item[k];
	}
	if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = sums[nb];
}

// short inputs (the chunk tables, the upper levels of the preconditioner's hierarchy, every scan of a small scene): ONE block walks
// the input in tiles of 1024 x 8 items with a running carry — one launch instead of three
constexpr uint64_t kScanSmall = 32768;
__global__ void __launch_bounds__(1024) k_scan_small(const uint32_t* __restrict__ in, uint64_t n, uint64_t* __restrict__ out)
{
	__shared__ uint64_t carry_s;
	if (threadIdx.x == 0) carry_s = 0;
	__syncthreads();
	for (uint64_t tile = 0; tile < n; tile += 1024 * kScanItems)
	{
		const uint64_t base = tile + static_cast<uint64_t>(threadIdx.x) * kScanItems;
		uint32_t item[kScanItems];
		uint64_t v = 0;
#pragma unroll
		for (int k = 0; k < kScanItems; k++)
		{
			const uint64_t i = base + k;
			item[k] = (i < n) ? in[i] : 0u;
			v += item[k];
		}
		uint64_t total;
		uint64_t run = block_exclusive<1024>(v, &total) + carry_s;
#pragma unroll
		for (int k = 0; k < kScanItems; k++)
		{
			const uint64_t i = base + k;
			if (i < n) out[i] = run;
			run += item[k];
		}
		__syncthreads();
		if (threadIdx.x == 0) carry_s += total;
		__syncthreads();
	}
	if (threadIdx.x == 0) out[n] = carry_s;
}

} // namespace

cudaError_t launch_exclusive_scan_u32_to_u64(const uint32_t* in, uint64_t* out, uint64_t n, DevBuf<uint64_t>& tmp, cudaStream_t st,
	uint64_t* launches)
{
	const uint64_t nb = (n + kScanTile - 1) / kScanTile;
	cudaError_t e = tmp.ensure(nb + 2, st);
	if (e != cudaSuccess) return e;
	if (n == 0)
	{
		return cudaMemsetAsync(out, 0, sizeof(uint64_t), st);
	}
	if (n <= kScanSmall)
	{
		k_scan_small<<<1, 1024, 0, st>>>(in, n, out);
		if (launches) *launches += 1;
		return cudaGetLastError();
	}
	k_tile_sums<<<static_cast<unsigned>(nb), kScanThreads, 0, st>>>(in, n, tmp.p);
	k_scan_sums<<<1, 1024, 0, st>>>(tmp.p, nb);
	k_tile_scan<<<static_cast<unsigned>(nb), kScanThreads, 0, st>>>(in, n, tmp.p, out, nb);
	if (launches) *launches += 3;
	return cudaGetLastError();
}

} // namespace mps
