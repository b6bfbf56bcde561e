// mps_cg.cu — the pressure-Poisson solve: one persistent cooperative kernel runs the WHOLE conjugate-gradient solve.
// Replaces Computer::SolvePressurePoissonEquation (Computer.hpp:1359-1429) and the ViennaCL host kernels under it
// (viennacl/linalg/host_based/sparse_matrix_operations.hpp:146-183 prod_impl, vector_operations.hpp:127 avbv, :540 inner_prod).
//
// Same algorithm and stopping rule as the reference (plain CG, no preconditioner, warm start from the previous
// pressure, converged when r'.r' < (r0.r0) * eps^2, at most n iterations, failure otherwise), restructured so that one
// iteration is TWO grid-wide phases instead of the textbook three kernels:
//
//   phase 1 (sub-warp per row, CSR-vector):   p_i  = r_i + beta * pprev_i                  (own row, written to pcur)
//                                              Ap_i = sum_j a_ij * (r_j + beta * pprev_j)   (neighbour p_j recomputed on the fly
//                                              pAp += p_i * Ap_i                             from r and pprev: bit-identical fma)
//   grid.sync  -> alpha = rr / pAp
//   phase 2 (thread per row):                  x_i += alpha * p_i ; r_i -= alpha * Ap_i ; rr' += r_i^2
//   grid.sync  -> converged? ; beta = rr' / rr
//
// i.e. the "p = r + beta p" pass and its grid-wide dependency are folded into the SpMV by double-buffering p.
// alpha, beta, the residual test and the iteration counter never leave the device; dot products are reduced
// deterministically (fixed per-block partials summed in a fixed order by every block).
//
// Roofline: HBM-bound.  Per iteration and row: CSR 12 B/nnz + row pointer 8 + (r, pprev own) 16 + (pcur, Ap) write 16 +
// phase 2 read 32 + write 16 = 12k + 88 B (gathers of r_j, pprev_j are served by L1/L2: slots are cell-sorted, so the
// columns of neighbouring rows overlap).  bench.py reports achieved = K * (12 nnz + 92 rows) / kernel time, the
// figure SURVEY.md §8d defines.
#include <cstdlib>
#include <vector>

#include <cooperative_groups.h>

#include "mps_async.cuh"
#include "mps_solver.h"

namespace cg = cooperative_groups;

namespace mps {
namespace {

constexpr int kCgThreads = 256;

struct CgArgs
{
	uint64_t n;
	const uint64_t* rowptr;
	const uint32_t* col;
	const double* val;
	const double* b;
	double* x;
	double* r;
	double* pbuf0;
	double* pbuf1;
	double* ap;
	double* partials; // [2][gridDim.x]
	DevScalars* sc;
	double eps;
};

__device__ __forceinline__ double block_sum(double v, double* smem /* kCgThreads / 32 + 1 */)
{
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
	const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	__syncthreads(); // smem may still be read from the previous call
	if (lane == 0) smem[wid] = v;
	__syncthreads();
	if (wid == 0)
	{
		double w = (lane < kCgThreads / 32) ? smem[lane] : 0.0;
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
		if (lane == 0) smem[kCgThreads / 32] = w;
	}
	__syncthreads();
	return smem[kCgThreads / 32];
}

// every block sums the per-block partials in the same fixed order -> identical value everywhere, run to run
__device__ __forceinline__ double grid_sum(const double* partials, unsigned nblocks, double* smem)
{
	double v = 0.0;
	for (unsigned k = threadIdx.x; k < nblocks; k += kCgThreads) v += __ldcg(partials + k);
	return block_sum(v, smem);
}

template<int LPR>
__global__ void __launch_bounds__(kCgThreads) k_cg_solve(CgArgs a)
{
	cg::grid_group grid = cg::this_grid();
	__shared__ double smem[kCgThreads / 32 + 1];

	const uint64_t n = a.n;
	const unsigned nblocks = gridDim.x;
	const uint64_t gtid = static_cast<uint64_t>(blockIdx.x) * kCgThreads + threadIdx.x;
	const uint64_t nthreads = static_cast<uint64_t>(nblocks) * kCgThreads;
	constexpr unsigned kRowsPerWarp = 32 / LPR;
	const uint64_t nsw = nthreads / LPR;                                   // sub-warps in the grid = rows per sweep
	const uint64_t warp_row0 = (gtid / 32) * kRowsPerWarp;                // first row of this warp in sweep 0
	const unsigned sub = (threadIdx.x & 31) / LPR;                         // sub-warp inside the warp
	const unsigned sl = threadIdx.x % LPR;                                 // lane inside the sub-warp
	double* part0 = a.partials;
	double* part1 = a.partials + nblocks;

	// ---- r0 = b - A x ; pprev = 0 ; rr = r0.r0 (Computer.hpp:1382-1386) ----
	double local = 0.0;
	// the sweep loop is warp-uniform (all 32 lanes take the same trips) so that the shuffles are always convergent
	for (uint64_t base = warp_row0; base < n; base += nsw)
	{
		const uint64_t row = base + sub;
		const bool valid = row < n;
		const uint64_t kb = valid ? a.rowptr[row] : 0, ke = valid ? a.rowptr[row + 1] : 0;
		double s = 0.0;
		for (uint64_t k = kb + sl; k < ke; k += LPR) s = fma(a.val[k], a.x[a.col[k]], s);
#pragma unroll
		for (int o = LPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o, LPR);
		if (valid && sl == 0)
		{
			const double ri = a.b[row] - s;
			a.r[row] = ri;
			a.pbuf0[row] = 0.0;
			local = fma(ri, ri, local);
		}
	}
	local = block_sum(local, smem);
	if (threadIdx.x == 0) part0[blockIdx.x] = local;
	grid.sync();
	double rr = grid_sum(part0, nblocks, smem);
	const double rr0 = rr;
	const double tol = rr * a.eps * a.eps;     // residual0, Computer.hpp:1386
	bool converged = (tol == 0);                // Computer.hpp:1389
	double beta = 0.0;
	double* pprev = a.pbuf0;
	double* pcur = a.pbuf1;
	uint64_t it = 0;

	while (it < n && !converged)
	{
		// ---- phase 1: p = r + beta p ; Ap = A p ; pAp ----
		local = 0.0;
		for (uint64_t base = warp_row0; base < n; base += nsw)
		{
			const uint64_t row = base + sub;
			const bool valid = row < n;
			const uint64_t kb = valid ? a.rowptr[row] : 0, ke = valid ? a.rowptr[row + 1] : 0;
			double s = 0.0;
			for (uint64_t k = kb + sl; k < ke; k += LPR)
			{
				const uint32_t c = a.col[k];
				s = fma(a.val[k], fma(beta, pprev[c], a.r[c]), s);
			}
#pragma unroll
			for (int o = LPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o, LPR);
			if (valid && sl == 0)
			{
				const double pi = fma(beta, pprev[row], a.r[row]);
				pcur[row] = pi;
				a.ap[row] = s;
				local = fma(pi, s, local);
			}
		}
		local = block_sum(local, smem);
		if (threadIdx.x == 0) part1[blockIdx.x] = local;
		grid.sync();
		const double pAp = grid_sum(part1, nblocks, smem);
		const double alpha = rr / pAp;

		// ---- phase 2: x += alpha p ; r -= alpha Ap ; rr' ----
		local = 0.0;
		for (uint64_t i = gtid; i < n; i += nthreads)
		{
			const double pi = pcur[i];
			a.x[i] = fma(alpha, pi, a.x[i]);
			const double ri = fma(-alpha, a.ap[i], a.r[i]);
			a.r[i] = ri;
			local = fma(ri, ri, local);
		}
		local = block_sum(local, smem);
		if (threadIdx.x == 0) part0[blockIdx.x] = local;
		grid.sync();
		const double rr_new = grid_sum(part0, nblocks, smem);
		it++;
		converged = (rr_new < tol);            // Computer.hpp:1407-1408
		if (!converged)
		{
			beta = rr_new / rr;                 // Computer.hpp:1417
			double* t = pprev; pprev = pcur; pcur = t;
		}
		rr = rr_new;
	}

	if (gtid == 0)
	{
		a.sc->cg_iterations = it;
		a.sc->rr0 = rr0;
		a.sc->rr = rr;
		a.sc->cg_converged = converged ? 1 : 0;
		if (!converged) atomicMax(&a.sc->error, static_cast<int>(MPS_CG_NOT_CONVERGED)); // Computer.hpp:1424-1428
	}
}


// =====================================================================================================================
// k_cg_stream — the solve of a particle-assembled system (chunk-blob form, mps_device.cuh / mps_chunk.cu).
//
// Same CG, same two phases per iteration, same fixed-order reductions as k_cg_solve above; what changes is how the
// bytes move.  The generic kernel is latency-bound (ncu, profiles/r01a: top stall long_scoreboard on the dependent chain
// row pointer -> (column, value) -> gathered p_j).  Here nothing in the inner loop touches global memory:
//   * each CTA (one per SM) owns a fixed, byte-balanced run of chunks.  A producer warp streams every chunk's blob
//     (values, 16-bit window-local columns, 16-bit row offsets: ONE contiguous 16-byte aligned segment) and the chunk's
//     window of the gathered vectors (r and p_prev: <= 3 / 9 contiguous slot ranges) into a ring of shared-memory stages
//     with 1-D bulk async copies (TMA engine, completion on an mbarrier), several stages ahead of the consumer warps;
//   * the consumers first turn the staged window into p = r + beta p_prev in place (once per window entry instead of
//     once per matrix entry), then run one row per thread (2-D) / per 4 lanes (3-D) entirely out of shared memory;
//   * matrix bytes are 10 B per entry instead of 12 (u16 columns) and the row pointer shrinks from 8 B to 2 B per row;
//   * when the blobs do not fit L2 they are streamed evict_first so that the vectors (5 x 8 B per row) stay L2-resident;
//   * chunks without entries (runs of Dummy / Disabled rows) are skipped: their rows are zeroed once per solve;
//   * the grid barrier is a single release-add / acquire-spin on one counter.
// HBM floor per iteration ~ 10 nnz + 2 rows (blobs); vector traffic is L2-resident up to ~2.5 M rows.
struct CgStreamArgs
{
	uint64_t n;
	const ChunkDesc* desc;     // the chunks that have entries (CgBuffers::live), a.sc->n_live of them
	const unsigned char* blobs;
	const double* b;
	double* x;
	double2* z0;               // {r_i, p_i} interleaved, ping-pong pair: ONE window copy per range brings both gathered vectors
	double2* z1;
	double* ap;
	double* r;                 // preconditioned solve only (k_pcg_stream): the residual; z0 / z1 then hold {z, p} with z = M^-1 r
	double* partials;          // [2][gridDim.x]
	DevScalars* sc;
	double eps;
	uint32_t blob_stage_bytes; // bytes reserved per stage for the blob (multiple of 128)
	uint32_t window_stage;     // doubles reserved per stage and vector for the window
	uint32_t stages;
	int l2_stream;             // blobs do not fit L2: stream them evict_first
	uint32_t producers;        // producer warps per CTA (the last warps of the block)
	unsigned long long* prof;  // optional [gridDim.x][8] cycle counters (mps_get_cg_profile), nullptr = off
	const double* cta_frac;    // [gridDim.x + 1] cumulative cost share of the CTAs (load balance, k_cg_rebalance)
	unsigned long long* cta_meas; // [2][gridDim.x] {cost taken, SpMV cycles} of this solve, or nullptr
	uint64_t own0, own1;       // rows this rank updates ([0, n) on one GPU)
	PeerLink peer;             // multi-GPU persistent solve (k_cg_stream<LPR, true>): neighbours' buffers and mailboxes over NVLink
};

constexpr int kProfStages = 64;     // per-stage counters of the preconditioned solve behind the per-CTA ones (CgBuffers::prof)
// Warps per CTA = consumer warps + producer warps.  Registers are allocated per SM sub-partition (4 per SM, 16384 each): 17..20
// warps put 5 on one of them (<= 96 registers per thread), 9..12 warps 3 (<= 168) — so up to 4 producer warps come for free
// next to 16 (3-D) / 8 (2-D) consumer warps.
constexpr int kMaxProducers = 4;
constexpr int kMaxStreamWarps = 16 + kMaxProducers;
constexpr int kStreamWarps2d = 8 + kMaxProducers;   // 2-D: at most 8 consumer warps
constexpr unsigned kGroups = 2; // consumer groups working on alternate chunks
constexpr uint32_t kDescBatch = 16; // descriptors per half of the producer's descriptor ring

__device__ __forceinline__ double block_sum_n(double v, double* red /* nwarps + 1 */)
{
	const unsigned nwarps = blockDim.x >> 5;
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
	const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	__syncthreads();
	if (lane == 0) red[wid] = v;
	__syncthreads();
	if (wid == 0)
	{
		double w = (lane < nwarps) ? red[lane] : 0.0;
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
		if (lane == 0) red[nwarps] = w;
	}
	__syncthreads();
	return red[nwarps];
}

__device__ __forceinline__ double grid_sum_n(const double* partials, const unsigned nblocks, double* red)
{
	double v = 0.0;
	for (unsigned k = threadIdx.x; k < nblocks; k += blockDim.x) v += __ldcg(partials + k);
	return block_sum_n(v, red);
}

// All CTAs of the (cooperatively launched, hence co-resident) grid; `target` is the running arrival count.
// Release/acquire on the counter is all the ordering needed: every cross-CTA read after the barrier goes to L2 (bulk / 16-byte
// async copies, __ldcg of the partial sums), everything else a CTA reads it wrote itself — so no L1 invalidation and no
// full fence (a __threadfence here costs ~1.5 us per barrier: MEMBAR.SC + CCTL.IVALL).
__device__ __forceinline__ void grid_barrier(unsigned long long* ctr, unsigned long long& target, const unsigned nblocks)
{
	target += nblocks;
	__syncthreads();
	if (threadIdx.x == 0)
	{
		async::red_release_gpu_add(ctr, 1ull);
		for (unsigned spin = 0; async::ld_acquire_gpu(ctr) < target; spin++)
		{
			if (spin > (1u << 27)) __trap(); // a lost CTA must end as a failed launch, never as a hung GPU
		}
	}
	__syncthreads();
}

// ---- multi-GPU: the same persistent kernel on every rank, coupled through peer memory over NVLink (no NCCL and no host in
//      the iteration).  Two things cross GPUs:
//   * the rim of {r, p}: the part of a chunk's window that lies in a neighbour rank's slab is fetched by the producer warp
//     straight from that rank's buffer (the same bulk async copy, peer address as its source: the transfer rides the ring
//     of stages like any other window copy, NVLink latency hidden by the look-ahead); nobody keeps copies of foreign rows;
//   * the two dot products per iteration: CTA 0 of every rank sums its grid's partials, stores {sum, sequence flag} into the
//     mailbox of every rank, its own included (one 16-byte store of self-validating words), waits for the flags of all ranks
//     in its own mailbox, adds the values in rank order (the same order everywhere: every rank gets the same bits and takes the
//     same convergence decision) and publishes the total to the other CTAs of its grid, which poll a local flag.
//   The exchange doubles as the grid barrier and as the cross-GPU barrier that orders the rim reads after the owner's writes:
//   writer -> CTA arrival (release.gpu) -> CTA 0 (acquire.gpu) -> mailbox word over NVLink -> reader; the rows themselves
//   never leave the owner's L2, which is where the neighbour's bulk copies read them.
__device__ __forceinline__ unsigned long long globaltimer_ns()
{
	unsigned long long t;
	asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
	return t;
}
__device__ __forceinline__ void st_relaxed_sys_v2(void* p, unsigned long long a, unsigned long long b)
{
	asm volatile("st.relaxed.sys.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(b) : "memory");
}
__device__ __forceinline__ void ld_relaxed_sys_v2(const void* p, unsigned long long& a, unsigned long long& b)
{
	asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}
__device__ __forceinline__ void st_release_gpu(unsigned long long* p, unsigned long long v)
{
	asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_gpu(const unsigned long long* p)
{
	unsigned long long v;
	asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
// A peer that never arrives (a rank that died) must end as a trapped launch, never as a hung GPU: waits are bounded in time.
constexpr unsigned long long kPeerWaitNs = 30ull * 1000ull * 1000ull * 1000ull;

// Sum of `local` over every thread of every CTA of every rank; `seq` counts the reductions of this solve.  One GPU: per-CTA
// partials, grid barrier, every CTA adds the partials in the same fixed order.  Multi-GPU: see above.
template<bool MG>
__device__ __forceinline__ double all_sum(const CgStreamArgs& a, double local, double* part, double* red, unsigned long long& bar_target,
	unsigned long long& seq, const unsigned nblocks)
{
	local = block_sum_n(local, red);
	if (!MG)
	{
		if (threadIdx.x == 0) part[blockIdx.x] = local;
		grid_barrier(&a.sc->grid_barrier, bar_target, nblocks);
		return grid_sum_n(part, nblocks, red);
	}
	const PeerLink& pl = a.peer;
	seq++;
	bar_target += nblocks - 1; // every thread keeps the same count: plain grid barriers on the same counter may follow (k_pcg_stream)
	const unsigned long long want = pl.tag | seq;
	const unsigned slot = static_cast<unsigned>(seq & 3ull);
	PeerMail* bc = pl.mail[pl.rank] + 4 * kMaxPeerRanks + slot; // the total, published by CTA 0 to the other CTAs of this grid
	if (threadIdx.x < 32)
	{
		const unsigned lane = threadIdx.x;
		double total = 0.0;
		if (blockIdx.x != 0)
		{
			if (lane == 0)
			{
				part[blockIdx.x] = local;
				async::red_release_gpu_add(&a.sc->grid_barrier, 1ull);
				// cheap polls (relaxed), one acquire fence once the flag is there
				const unsigned long long t0 = globaltimer_ns();
				for (unsigned spin = 0; ld_relaxed_gpu(&bc->flag) != want; spin++)
					if ((spin & 1023u) == 1023u && globaltimer_ns() - t0 > kPeerWaitNs) __trap();
				fence_acq_rel_gpu();
				total = __ldcg(&bc->value);
			}
		}
		else
		{
			// CTA 0: wait for the other CTAs of this grid, add the partials in a fixed order, exchange the sum with the other ranks
			if (lane == 0)
			{
				part[0] = local;
				for (unsigned spin = 0; ld_relaxed_gpu(&a.sc->grid_barrier) < bar_target; spin++)
					if (spin > (1u << 27)) __trap();
				fence_acq_rel_gpu();
			}
			__syncwarp();
			double v = 0.0;
			for (unsigned k = lane; k < nblocks; k += 32) v += __ldcg(part + k);
#pragma unroll
			for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
			double mine = 0.0;
			if (lane < static_cast<unsigned>(pl.nranks))
			{
				// My sum into everybody's mailbox (my own included) as ONE 16-byte store of two self-validating words, each
				// {32 bits of the value, 32-bit sequence flag} — the flag-in-data scheme of NCCL's LL protocol: an aligned 8-byte
				// word is never torn, so a reader that sees both flags has the whole value, and no system-scope fence sits on the
				// critical path (a fence.sys costs more than the NVLink flight).  Ordering of the rim rows needs none either:
				// they were in this GPU's L2 before their writers' arrivals were counted above, and the neighbours read them
				// from that L2 (peer addresses are not cached on the reading side) after they have seen this flag.
				const unsigned long long f32 = ((pl.tag >> 32) & 0xffull) << 24 | (seq & 0xffffffull);
				const unsigned long long bits = static_cast<unsigned long long>(__double_as_longlong(v));
				PeerMail* out = pl.mail[lane] + slot * kMaxPeerRanks + pl.rank;
				st_relaxed_sys_v2(out, (bits & 0xffffffffull) | (f32 << 32), (bits >> 32) | (f32 << 32));
				const PeerMail* in = pl.mail[pl.rank] + slot * kMaxPeerRanks + lane;
				const unsigned long long t0 = globaltimer_ns();
				unsigned long long w0 = 0, w1 = 0;
				for (unsigned spin = 0;; spin++)
				{
					ld_relaxed_sys_v2(in, w0, w1);
					if ((w0 >> 32) == f32 && (w1 >> 32) == f32) break;
					if ((spin & 1023u) == 1023u && globaltimer_ns() - t0 > kPeerWaitNs) __trap();
				}
				mine = __longlong_as_double(static_cast<long long>((w0 & 0xffffffffull) | (w1 << 32)));
			}
			for (int r = 0; r < pl.nranks; r++) total += __shfl_sync(0xffffffffu, mine, r); // rank order: identical bits on every rank
			if (lane == 0)
			{
				bc->value = total;
				st_release_gpu(&bc->flag, want);
			}
		}
		if (lane == 0) red[0] = total;
	}
	__syncthreads();
	const double total = red[0];
	__syncthreads(); // red is reused by the next block_sum_n
	return total;
}

struct StreamSmem
{
	unsigned char* base;
	uint32_t stage_bytes, blob_stage_bytes, window_stage;
	uint64_t* full;
	uint64_t* empty;
	uint64_t* dfull;   // [2] descriptor ring halves
	ChunkDesc* dring;  // [2][kDescBatch]
	// one stage = [blob = descriptor copy 128 B + values + columns + row offsets][windows]
	__device__ __forceinline__ unsigned char* stage(uint32_t s) const { return base + static_cast<size_t>(s) * stage_bytes; }
	__device__ __forceinline__ const ChunkDesc* desc(uint32_t s) const { return reinterpret_cast<const ChunkDesc*>(stage(s)); }
	__device__ __forceinline__ unsigned char* blob(uint32_t s) const { return stage(s) + kBlobHeader; }
	// [window of {r, p_prev} : window_stage x 16 B][dense window of the search direction p (or of x) : window_stage x 8 B]
	__device__ __forceinline__ double2* zwin(uint32_t s) const { return reinterpret_cast<double2*>(stage(s) + blob_stage_bytes); }
	__device__ __forceinline__ double* pwin(uint32_t s) const { return reinterpret_cast<double*>(zwin(s) + window_stage); }
};

// One SpMV phase over this CTA's chunks.  ITER = false: r = b - A x, p_prev = 0 (window of x; result into zcur).
// ITER = true: p = r + beta p_prev, Ap = A p (window of zprev = {r, p_prev}; p into zcur[].y).
// Returns this thread's share of r.r / p.Ap.  `it` counts the ring uses so far.
template<int LPR, bool ITER, bool MG, bool PRE = false>
__device__ __forceinline__ double spmv_phase(const CgStreamArgs& a, const StreamSmem& sm, const uint32_t c0, const uint32_t c1, const uint32_t live,
	uint32_t& it, uint32_t& dseq, const double beta, const double2* __restrict__ zprev, double2* __restrict__ zcur,
	const uint64_t pol_matrix, const uint64_t pol_vector)
{
	const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const unsigned np = a.producers;
	const unsigned cwarps = (blockDim.x >> 5) - np; // consumer warps; the last np warps are the producers
	const uint32_t S = a.stages;
	double local = 0.0;
	if (warp >= cwarps)
	{
		// ---- producer warps: run up to S - 1 chunks ahead of the consumers.  A bulk copy holds the issuing WARP for ~560 cycles
		//      whatever its size, and copies issued by several lanes of one warp go one after the other — but copies issued by
		//      different warps overlap (tools/tma_bench.cu, "multi": 564 cycles per copy with one issuing warp, 74 with eight).
		//      So the copies of one chunk (the blob + one per window range: 4 in 2-D, 10 in 3-D) are dealt round-robin to the np
		//      producer warps: copy c is issued by lane c / np of producer warp c % np.  Every producer warp walks the same chunk
		//      sequence and waits for the stage to be free itself; warp 0 alone posts the expected bytes (the stage's barrier
		//      cannot complete before that: its one pending arrival is that post). ----
		const unsigned pw = warp - cwarps;
		if (lane == 0) async::fence_proxy_async(); // other CTAs' stores (made visible by the grid barrier) before our async-proxy reads
		__syncwarp();
		// The descriptors of this CTA's chunks are themselves streamed through a small double-buffered shared-memory ring
		// (kDescBatch per half): under full memory load an ordinary global load takes thousands of cycles, and the producers
		// must never wait for one.
		uint32_t k = it;
		const uint32_t nbatch = (c1 - c0 + kDescBatch - 1) / kDescBatch;
		auto fetch_batch = [&](const uint32_t bidx, const uint32_t seq)
		{
			// producer warp 0, lane 0 only
			const uint32_t first = c0 + bidx * kDescBatch;
			const uint32_t count = (c1 - first < kDescBatch) ? (c1 - first) : kDescBatch;
			const uint32_t bytes = count * static_cast<uint32_t>(sizeof(ChunkDesc));
			async::mbar_arrive_expect_tx(&sm.dfull[seq & 1u], bytes);
			async::bulk_g2s(sm.dring + (seq & 1u) * kDescBatch, a.desc + first, bytes, &sm.dfull[seq & 1u], pol_vector);
		};
		if (pw == 0 && lane == 0)
		{
			if (nbatch > 0) fetch_batch(0, dseq);
			if (nbatch > 1) fetch_batch(1, dseq + 1);
		}
		const uint32_t cidx = pw + lane * np; // this lane's copy of every chunk: 0 the blob, 1.. window range cidx - 1
		for (uint32_t bidx = 0; bidx < nbatch; bidx++)
		{
			const uint32_t seq = dseq + bidx;
			const ChunkDesc* half = sm.dring + (seq & 1u) * kDescBatch;
			async::mbar_wait(&sm.dfull[seq & 1u], (seq >> 1) & 1u);
			const uint32_t first = c0 + bidx * kDescBatch;
			const uint32_t count = (c1 - first < kDescBatch) ? (c1 - first) : kDescBatch;
			for (uint32_t j = 0; j < count; j++)
			{
				const ChunkDesc* d = half + j;
				if (d->nnz == 0) continue; // warp-uniform
				const uint32_t s = k % S, use = k / S;
				uint64_t* const fullbar = &sm.full[s * kGroups + k % kGroups]; // chunk k belongs to consumer group k % kGroups
				k++;
				// {r, p_prev} travel interleaved: one copy per window range instead of two
				const void* src = nullptr; void* dst = nullptr; uint32_t bytes = 0;
				if (cidx == 0) { src = a.blobs + d->blob_off; dst = sm.stage(s); bytes = d->blob_bytes; }
				else if (cidx <= kMaxRanges)
				{
					const uint32_t q = cidx - 1;
					const uint32_t len = d->range_len[q]; // 0 for unused ranges
					if (ITER) { bytes = len * 16u; src = zprev + d->range_start[q]; dst = sm.zwin(s) + d->range_off[q]; }
					else { bytes = len * 8u; src = a.x + d->range_start[q]; dst = sm.pwin(s) + d->range_off[q]; }
				}
				if (lane == 0)
				{
					const long long tw0 = (a.prof && pw == 0) ? clock64() : 0;
					async::mbar_wait(&sm.empty[s], (use & 1u) ^ 1u);
					if (a.prof && pw == 0) a.prof[blockIdx.x * 8 + 4] += static_cast<unsigned long long>(clock64() - tw0);
					if (pw == 0) async::mbar_arrive_expect_tx(fullbar, d->blob_bytes + d->window * (ITER ? 16u : 8u));
				}
				__syncwarp();
				if (MG && ITER && cidx != 0 && bytes)
				{
					// multi-GPU: slots below own0 / from own1 on live in the left / right neighbour's buffer (same index, its memory)
					const PeerLink& pl = a.peer;
					const double2* nbz[2] = { (zprev == a.z0) ? pl.nb_z0[0] : pl.nb_z1[0], (zprev == a.z0) ? pl.nb_z0[1] : pl.nb_z1[1] };
					const uint64_t lo = nbz[0] ? a.own0 : 0ull, hi = nbz[1] ? a.own1 : ~0ull;
					const uint64_t rb = d->range_start[cidx - 1], re = rb + d->range_len[cidx - 1];
					double2* wdst = static_cast<double2*>(dst);
					const uint64_t m0 = rb < lo ? (re < lo ? re : lo) : rb; // [rb, m0) left neighbour
					const uint64_t m1 = re > hi ? (rb > hi ? rb : hi) : re; // [m1, re) right neighbour
					if (m0 > rb) async::bulk_g2s(wdst, nbz[0] + rb, static_cast<uint32_t>(m0 - rb) * 16u, fullbar, pol_vector);
					if (m1 > m0) async::bulk_g2s(wdst + (m0 - rb), zprev + m0, static_cast<uint32_t>(m1 - m0) * 16u, fullbar, pol_vector);
					if (re > m1) async::bulk_g2s(wdst + (m1 - rb), nbz[1] + m1, static_cast<uint32_t>(re - m1) * 16u, fullbar, pol_vector);
				}
				else if (bytes) async::bulk_g2s(dst, src, bytes, fullbar, cidx == 0 ? pol_matrix : pol_vector);
			}
			// every lane of every producer warp is done reading this half
			if (np > 1) asm volatile("bar.sync %0, %1;" ::"r"(1u + kGroups), "r"(np * 32u) : "memory");
			else __syncwarp();
			if (pw == 0 && lane == 0 && bidx + 2 < nbatch) fetch_batch(bidx + 2, seq + 2);
		}
		dseq += nbatch;
	}
	else
	{
		// ---- consumers: kGroups groups of warps take alternate chunks, so that one group's fixed per-chunk work (barrier
		//      waits, window pass, epilogue) overlaps the other's shared-memory traffic.  One row per LPR lanes; everything,
		//      descriptor included, comes from shared memory ----
		const unsigned gwarps = cwarps / kGroups;      // warps per group
		const unsigned group = warp / gwarps;
		const unsigned ctid = threadIdx.x - group * gwarps * 32, nct = gwarps * 32;
		const uint32_t lr = ctid / LPR, sl = ctid % LPR;
		// Chunk k (counted over all phases of the solve) sits in stage k % S and belongs to group k % kGroups.  A group must meet the
		// phases of a barrier it waits on ONE BY ONE (a parity wait cannot tell phase n from phase n + 2), and with an odd S
		// consecutive uses of a stage alternate between the groups — so every (stage, group) pair has a "full" barrier of its own:
		// the pair recurs every P = lcm(S, kGroups) chunks, its use number is k / P.
		const uint32_t P = (S % kGroups == 0) ? S : S * kGroups;
		for (uint32_t q = (group + kGroups - it % kGroups) % kGroups; q < live; q += kGroups)
		{
			const uint32_t k = it + q;
			const uint32_t s = k % S;
			const long long tw0 = (a.prof && threadIdx.x == 0) ? clock64() : 0;
			async::mbar_wait(&sm.full[s * kGroups + group], (k / P) & 1u);
			if (a.prof && threadIdx.x == 0) { a.prof[blockIdx.x * 8 + 1] += static_cast<unsigned long long>(clock64() - tw0); a.prof[blockIdx.x * 8 + 5] += 1; }
			const ChunkDesc* d = sm.desc(s);
			const uint32_t row_begin = d->row_begin, rows = d->rows, nnz_pad = round_up8(d->nnz), window = d->window;
			const int32_t self_off = d->self_off;
			const double* __restrict__ val = reinterpret_cast<const double*>(sm.blob(s));
			const uint16_t* __restrict__ lcol = reinterpret_cast<const uint16_t*>(sm.blob(s) + static_cast<size_t>(nnz_pad) * 8u);
			const uint16_t* __restrict__ rowoff = lcol + nnz_pad;
			double* w0 = sm.pwin(s);
			const bool valid = lr < rows;
			const uint32_t kb = valid ? rowoff[lr] : 0u, ke = valid ? rowoff[lr + 1] : 0u;
			uint32_t e = kb + sl;
			if (ITER)
			{
				// the search direction of the whole window, once per entry: p = r + beta p_prev
				const double2* zw = sm.zwin(s);
				for (uint32_t l = ctid; l < window; l += nct) { const double2 z = zw[l]; w0[l] = fma(beta, z.y, z.x); }
				asm volatile("bar.sync %0, %1;" ::"r"(1u + group), "r"(nct) : "memory"); // this group's warps only
			}
			double acc0 = 0.0, acc1 = 0.0;
			for (; e + 7 * LPR < ke; e += 8 * LPR)
			{
				uint32_t l[8]; double av[8], pv[8];
#pragma unroll
				for (int u = 0; u < 8; u++) { l[u] = lcol[e + u * LPR]; av[u] = val[e + u * LPR]; }
#pragma unroll
				for (int u = 0; u < 8; u++) pv[u] = w0[l[u]];
#pragma unroll
				for (int u = 0; u < 8; u += 2) { acc0 = fma(av[u], pv[u], acc0); acc1 = fma(av[u + 1], pv[u + 1], acc1); }
			}
			if (e + 3 * LPR < ke)
			{
				uint32_t l[4]; double av[4], pv[4];
#pragma unroll
				for (int u = 0; u < 4; u++) { l[u] = lcol[e + u * LPR]; av[u] = val[e + u * LPR]; }
#pragma unroll
				for (int u = 0; u < 4; u++) pv[u] = w0[l[u]];
#pragma unroll
				for (int u = 0; u < 4; u += 2) { acc0 = fma(av[u], pv[u], acc0); acc1 = fma(av[u + 1], pv[u + 1], acc1); }
				e += 4 * LPR;
			}
			for (; e < ke; e += LPR) acc0 = fma(val[e], w0[lcol[e]], acc0);
			double acc = acc0 + acc1;
			if (LPR > 1)
			{
#pragma unroll
				for (int o = LPR / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o, LPR);
			}
			if (valid && sl == 0)
			{
				const uint64_t row = static_cast<uint64_t>(row_begin) + lr;
				if (ITER)
				{
					const double pi = (ke > kb) ? w0[self_off + static_cast<int32_t>(lr)] : 0.0; // rows without entries have r = p = 0
					zcur[row].y = pi;
					a.ap[row] = acc;
					local = fma(pi, acc, local);
				}
				else
				{
					const double ri = a.b[row] - acc;
					if (PRE) { a.r[row] = ri; zcur[row] = make_double2(0.0, 0.0); } // z = M^-1 r is formed by the preconditioner
					else zcur[row] = make_double2(ri, 0.0);
					local = fma(ri, ri, local);
				}
			}
			__syncwarp();
			if (lane == 0) async::mbar_arrive(&sm.empty[s]); // this warp is done reading stage s
		}
	}
	it += live;
	return local;
}

// Per-CTA set-up shared by the persistent kernel (one GPU) and the stepwise kernels (one launch per phase, multi-GPU):
// shared-memory carve-up, barrier initialisation, this CTA's cost-balanced run of chunks and its rows, clamped to the rows
// the rank owns.
struct StreamCta
{
	StreamSmem sm;
	double* red;
	uint32_t c0, c1, live;
	uint64_t row0, row1;
	uint64_t pol_matrix, pol_vector;
};

template<bool ZERO_ROWS, bool MG = false>
__device__ __forceinline__ StreamCta stream_setup(const CgStreamArgs& a, unsigned char* smem_raw)
{
	const uint32_t S = a.stages;
	const unsigned nthreads = blockDim.x, cwarps = (blockDim.x >> 5) - a.producers;
	StreamCta c;
	StreamSmem& sm = c.sm;
	sm.base = smem_raw;
	sm.blob_stage_bytes = a.blob_stage_bytes;
	sm.window_stage = a.window_stage;
	sm.stage_bytes = a.blob_stage_bytes + 24u * a.window_stage;
	sm.dring = reinterpret_cast<ChunkDesc*>(smem_raw + static_cast<size_t>(S) * sm.stage_bytes);
	sm.full = reinterpret_cast<uint64_t*>(sm.dring + 2 * kDescBatch); // [S][kGroups]
	sm.empty = sm.full + S * kGroups;
	sm.dfull = sm.empty + S;
	c.red = reinterpret_cast<double*>(sm.dfull + 2);                          // kMaxStreamWarps + 1
	uint64_t* ctl64 = reinterpret_cast<uint64_t*>(c.red + kMaxStreamWarps + 1); // row0, row1, zero0, zero1
	uint32_t* ctl = reinterpret_cast<uint32_t*>(ctl64 + 4);                    // c0, c1

	const unsigned nblocks = gridDim.x;
	if (threadIdx.x == 0)
	{
		for (uint32_t s = 0; s < S * kGroups; s++) async::mbar_init(&sm.full[s], 1u);
		for (uint32_t s = 0; s < S; s++) async::mbar_init(&sm.empty[s], cwarps / kGroups);
		async::mbar_init(&sm.dfull[0], 1u); async::mbar_init(&sm.dfull[1], 1u);
		async::mbar_init_fence();
		// this CTA's run of the chunks that have entries: split the modelled cost evenly (descriptors hold its exclusive prefix).
		// Chunks without entries (runs of Dummy / Disabled rows; on several GPUs all the other ranks' rows) are not in the list.
		const uint64_t nlive = a.sc->n_live, total = a.sc->cost_total;
		uint32_t bound[2];
		for (int w = 0; w < 2; w++)
		{
			const uint64_t b = blockIdx.x + w;
			if (b >= nblocks) { bound[w] = static_cast<uint32_t>(nlive); continue; }
			const uint64_t target = static_cast<uint64_t>(static_cast<double>(total) * a.cta_frac[b]);
			uint64_t lo = 0, hi = nlive; // first chunk with cost_off >= target
			while (lo < hi)
			{
				const uint64_t mid = (lo + hi) >> 1;
				if (a.desc[mid].cost_off < target) lo = mid + 1; else hi = mid;
			}
			bound[w] = static_cast<uint32_t>(lo);
		}
		ctl[0] = bound[0]; ctl[1] = bound[1];
		// rows this CTA updates in phase 2: an even share of the rows this rank owns — phase 2 is row-local, so its split is
		// independent of who multiplied which chunk (its loads go to L2: another CTA wrote p and Ap)
		const uint64_t own = a.own1 - a.own0;
		ctl64[0] = a.own0 + (own / nblocks) * blockIdx.x + (own % nblocks) * blockIdx.x / nblocks;
		ctl64[1] = a.own0 + (own / nblocks) * (blockIdx.x + 1ull) + (own % nblocks) * (blockIdx.x + 1ull) / nblocks;
		// rows this CTA zeroes before the first phase: from its first chunk to the next CTA's first chunk (the same CTA then writes
		// the residual of the rows that have entries; together the CTAs cover every row this rank owns)
		uint64_t z0r = (blockIdx.x == 0) ? a.own0 : ((bound[0] < nlive) ? a.desc[bound[0]].row_begin : a.own1);
		uint64_t z1r = (blockIdx.x + 1 == nblocks) ? a.own1 : ((bound[1] < nlive) ? a.desc[bound[1]].row_begin : a.own1);
		ctl64[2] = z0r < a.own0 ? a.own0 : (z0r > a.own1 ? a.own1 : z0r);
		ctl64[3] = z1r < a.own0 ? a.own0 : (z1r > a.own1 ? a.own1 : z1r);
		if (a.cta_meas)
		{
			const uint64_t cb = (bound[0] < nlive) ? a.desc[bound[0]].cost_off : total, ce = (bound[1] < nlive) ? a.desc[bound[1]].cost_off : total;
			a.cta_meas[blockIdx.x] = ce - cb;
		}
	}
	__syncthreads();
	c.c0 = ctl[0]; c.c1 = ctl[1];
	c.row0 = ctl64[0]; c.row1 = ctl64[1];
	c.live = c.c1 - c.c0;
	{
		if (ZERO_ROWS)
		{
			for (uint64_t i = ctl64[2] + threadIdx.x; i < ctl64[3]; i += nthreads)
			{
				a.z0[i] = make_double2(0.0, 0.0); a.z1[i] = make_double2(0.0, 0.0); a.ap[i] = 0.0;
				if (a.r) a.r[i] = 0.0;
			}
		}
	}
	__syncthreads();
	c.pol_matrix = a.l2_stream ? async::policy_evict_first() : async::policy_evict_normal();
	c.pol_vector = a.l2_stream ? async::policy_evict_last() : async::policy_evict_normal();
	return c;
}

// phase 2 over rows [row0, row1): x += alpha p ; r -= alpha Ap ; returns this thread's share of r.r (4 independent rows in flight)
template<bool MG = false>
__device__ __forceinline__ double update_rows(const CgStreamArgs& a, const uint64_t row0, const uint64_t row1, const double alpha,
	const double2* __restrict__ zprev, double2* __restrict__ zcur)
{
	const unsigned nthreads = blockDim.x;
	double local = 0.0;
	for (uint64_t base = row0 + threadIdx.x; base < row1; base += 4ull * nthreads)
	{
		double pv[4], xv[4], av[4], rv[4];
#pragma unroll
		for (int q = 0; q < 4; q++)
		{
			const uint64_t i = base + static_cast<uint64_t>(q) * nthreads;
			const bool on = i < row1;
			// L2 loads: p and Ap of these rows were written by whichever CTA multiplied their chunk
			pv[q] = on ? __ldcg(&zcur[i].y) : 0.0; xv[q] = on ? __ldcg(a.x + i) : 0.0; av[q] = on ? __ldcg(a.ap + i) : 0.0; rv[q] = on ? __ldcg(&zprev[i].x) : 0.0;
		}
#pragma unroll
		for (int q = 0; q < 4; q++)
		{
			const uint64_t i = base + static_cast<uint64_t>(q) * nthreads;
			if (i < row1)
			{
				a.x[i] = fma(alpha, pv[q], xv[q]);
				const double ri = fma(-alpha, av[q], rv[q]);
				zcur[i].x = ri;
				local = fma(ri, ri, local);
			}
		}
	}
	return local;
}

template<int LPR, bool MG>
__global__ void __launch_bounds__(kMaxStreamWarps * 32, 1) k_cg_stream(CgStreamArgs a)
{
	extern __shared__ __align__(128) unsigned char smem_raw[];
	// An earlier stage of this step failed (cell overflow, a chunk that does not fit the stages): the descriptors may not be
	// trustworthy, so no CTA touches them.  Every CTA reads the flag before its first barrier, and the only writer inside this
	// kernel is the very end of the solve: the decision is the same everywhere.
	if (*reinterpret_cast<volatile int*>(&a.sc->error) != 0) return;
	const StreamCta cta = stream_setup<true, MG>(a, smem_raw);
	const StreamSmem& sm = cta.sm;
	double* red = cta.red;
	const uint32_t c0 = cta.c0, c1 = cta.c1, live = cta.live;
	const uint64_t row0 = cta.row0, row1 = cta.row1;
	const uint64_t pol_matrix = cta.pol_matrix, pol_vector = cta.pol_vector;
	const uint64_t n = a.n;
	const unsigned nblocks = gridDim.x;
	double* part0 = a.partials;
	double* part1 = a.partials + nblocks;
	unsigned long long bar_target = 0, seq = 0;
	uint32_t it = 0, dseq = 0;
	const bool prof_on = (a.prof != nullptr) && (threadIdx.x == 0);
	const bool meas_on = (a.cta_meas != nullptr) && (threadIdx.x == 0);
	unsigned long long spmv_cycles = 0;

	// ---- r0 = b - A x ; p_prev = 0 ; rr = r0.r0 (Computer.hpp:1382-1386) ----
	double local = spmv_phase<LPR, false, MG>(a, sm, c0, c1, live, it, dseq, 0.0, nullptr, a.z0, pol_matrix, pol_vector);
	double rr = all_sum<MG>(a, local, part0, red, bar_target, seq, nblocks);
	const double rr0 = rr;
	const double tol = rr * a.eps * a.eps;     // Computer.hpp:1386
	bool converged = (tol == 0);                // Computer.hpp:1389
	double beta = 0.0;
	double2* zprev = a.z0; // {r, p_prev}
	double2* zcur = a.z1;  // receives {r', p}
	uint64_t iter = 0;

	while (iter < n && !converged)
	{
		// ---- phase 1: p = r + beta p_prev ; Ap = A p ; p.Ap ----
		const long long t0 = (prof_on || meas_on) ? clock64() : 0;
		local = spmv_phase<LPR, true, MG>(a, sm, c0, c1, live, it, dseq, beta, zprev, zcur, pol_matrix, pol_vector);
		const long long t1 = (prof_on || meas_on) ? clock64() : 0;
		spmv_cycles += static_cast<unsigned long long>(t1 - t0);
		const double pAp = all_sum<MG>(a, local, part1, red, bar_target, seq, nblocks);
		const long long t2 = prof_on ? clock64() : 0;
		const double alpha = rr / pAp;

		// ---- phase 2 over this CTA's own rows: x += alpha p ; r -= alpha Ap ; r.r ----
		local = update_rows<MG>(a, row0, row1, alpha, zprev, zcur);
		const long long t3 = prof_on ? clock64() : 0;
		const double rr_new = all_sum<MG>(a, local, part0, red, bar_target, seq, nblocks);
		if (prof_on)
		{
			const long long t4 = clock64();
			unsigned long long* pr = a.prof + blockIdx.x * 8;
			pr[0] += static_cast<unsigned long long>(t1 - t0);  // phase 1 (SpMV)
			pr[2] += static_cast<unsigned long long>(t3 - t2);  // phase 2
			pr[3] += static_cast<unsigned long long>((t2 - t1) + (t4 - t3)); // both reductions (block sums, grid / peer barrier, sum)
			pr[6] += static_cast<unsigned long long>(t4 - t0);
		}
		iter++;
		converged = (rr_new < tol);            // Computer.hpp:1407-1408
		if (!converged)
		{
			beta = rr_new / rr;                 // Computer.hpp:1417
		}
		{ double2* t = zprev; zprev = zcur; zcur = t; } // zprev now holds {r, p} of this iteration
		rr = rr_new;
	}

	if (meas_on) a.cta_meas[nblocks + blockIdx.x] = spmv_cycles; // what k_cg_rebalance turns into the next solve's split
	if (blockIdx.x == 0 && threadIdx.x == 0)
	{
		a.sc->z_final = (zprev == a.z1) ? 1 : 0;
		a.sc->cg_iterations = iter;
		a.sc->rr0 = rr0;
		a.sc->rr = rr;
		a.sc->cg_converged = converged ? 1 : 0;
		if (!converged) atomicMax(&a.sc->error, static_cast<int>(MPS_CG_NOT_CONVERGED)); // Computer.hpp:1424-1428
	}
}

// =====================================================================================================================
// k_pcg_stream — the same solve, preconditioned:  M^-1 = D^-1 + P1 V P1^T  (mps_mg.cu describes M and builds its operators).
//
// Same system, same warm start, same stopping rule on the same (unpreconditioned) residual norm as the reference
// (||r||^2 < eps^2 ||r0||^2, at most n iterations, Computer.hpp:1382-1428); what changes is the search direction:
//   p = z + beta p_prev,  z = M^-1 r,  alpha = r.z / p.Ap,  beta = r'.z' / r.z
// The SpMV phase is k_cg_stream's, untouched: the interleaved pair it gathers is {z, p_prev} instead of {r, p_prev}.
// Phase 2 grows two row passes and the V-cycle between them (everything on the cell hierarchy is L2-resident and tiny next
// to the fine matrix; its cost is the grid barriers, one per level and direction):
//   2a  x += alpha p ; r -= alpha Ap ; r.r ; z0 = r / diag ; level-0 residual = per-cell sums of r (warp-segmented sums over
//       cell-aligned row ranges: fixed order, no atomics) and its first Jacobi sweep
//       -> barrier (r.r: the reference's convergence test)
//   V   down: residual of level l restricted to level l+1 (thread per coarse cell, gather over its 2^D children) + Jacobi
//       from zero; top: a few more sweeps; up: over-corrected prolongation fused with the post-smoothing sweep
//   2b  z = z0 + e0[cell(row)] ; r.z  -> barrier
constexpr int kRowTiles = 4; // 32-row tiles whose loads a warp keeps in flight together in the row passes

template<bool FIRST>
__device__ __forceinline__ double pcg_rows_a(const CgStreamArgs& a, const MgArgs& m, const uint64_t rb, const uint64_t re, const double alpha, double2* __restrict__ zt)
{
	const unsigned lane = threadIdx.x & 31;
	const double* __restrict__ dinv_c = m.lv[0].dinv;
	double* __restrict__ rc = m.lv[0].r;
	double* __restrict__ ec = m.lv[0].e0;
	double local = 0.0, carry = 0.0, carry_dc = 0.0;
	uint32_t carry_id = kMgNone;
	for (uint64_t base = rb; base < re; base += 32ull * kRowTiles) // warp-uniform trip count
	{
		// all loads of the batch first: the pass is bound by L2 latency, not by arithmetic
		double rv[kRowTiles], pv[kRowTiles], av[kRowTiles], xv[kRowTiles], dv[kRowTiles], dc[kRowTiles];
		uint32_t id[kRowTiles];
#pragma unroll
		for (int q = 0; q < kRowTiles; q++)
		{
			const uint64_t i = base + 32ull * q + lane;
			const bool on = i < re;
			id[q] = on ? __ldg(m.crow + i) : kMgNone;
			dv[q] = on ? __ldg(m.dinv0 + i) : 0.0;
			rv[q] = on ? __ldcg(m.r + i) : 0.0;
			if (!FIRST)
			{
				pv[q] = on ? __ldcg(&zt[i].y) : 0.0; av[q] = on ? __ldcg(a.ap + i) : 0.0; xv[q] = on ? __ldcg(a.x + i) : 0.0;
			}
		}
		// damped inverse diagonal of the row's cell (a broadcast load for the lanes of one cell): fetched here, not where a
		// cell's sum is finished, so that the sums below never wait on memory
#pragma unroll
		for (int q = 0; q < kRowTiles; q++) dc[q] = (id[q] != kMgNone) ? __ldg(dinv_c + id[q]) : 0.0;
#pragma unroll
		for (int q = 0; q < kRowTiles; q++)
		{
			const uint64_t i = base + 32ull * q + lane;
			if (i < re)
			{
				if (!FIRST)
				{
					a.x[i] = fma(alpha, pv[q], xv[q]);
					rv[q] = fma(-alpha, av[q], rv[q]);
					m.r[i] = rv[q];
				}
				zt[i].x = rv[q] * dv[q];
				local = fma(rv[q], rv[q], local);
			}
		}
		// per-cell sums of r: inclusive sums inside every run of equal cell ids (rows of one cell are contiguous), tile by tile,
		// a cell that crosses a tile boundary carried in (carry, carry_id)
#pragma unroll
		for (int q = 0; q < kRowTiles; q++)
		{
			const uint64_t t0 = base + 32ull * q;
			if (t0 >= re) break; // warp-uniform
			const uint64_t i = t0 + lane;
			const bool on = i < re;
			double v = rv[q];
#pragma unroll
			for (int o = 1; o < 32; o <<= 1)
			{
				const double vu = __shfl_up_sync(0xffffffffu, v, o);
				const uint32_t iu = __shfl_up_sync(0xffffffffu, id[q], o);
				if (lane >= static_cast<unsigned>(o) && iu == id[q]) v += vu;
			}
			if (carry_id != kMgNone)
			{
				if (id[q] == carry_id) v += carry;                               // the cell continues from the previous tile
				else if (lane == 0) { rc[carry_id] = carry; ec[carry_id] = carry_dc * carry; } // it ended exactly there
			}
			const uint32_t idn = __shfl_down_sync(0xffffffffu, id[q], 1);
			const bool tail = on && ((lane == 31) ? (i + 1 == re) : (idn != id[q]));
			if (tail) { rc[id[q]] = v; ec[id[q]] = dc[q] * v; }
			const double v31 = __shfl_sync(0xffffffffu, v, 31);
			const uint32_t id31 = __shfl_sync(0xffffffffu, id[q], 31);
			const int open31 = __shfl_sync(0xffffffffu, (on && !tail) ? 1 : 0, 31);
			carry_dc = __shfl_sync(0xffffffffu, dc[q], 31);
			carry = v31;
			carry_id = open31 ? id31 : kMgNone;
		}
	}
	return local;
}

__device__ __forceinline__ double pcg_rows_b(const MgArgs& m, const uint64_t rb, const uint64_t re, const double* __restrict__ ef, double2* __restrict__ zt)
{
	const unsigned lane = threadIdx.x & 31;
	double local = 0.0;
	for (uint64_t base = rb; base < re; base += 32ull * kRowTiles)
	{
		double zv[kRowTiles], rv[kRowTiles], ev[kRowTiles];
		uint32_t id[kRowTiles];
#pragma unroll
		for (int q = 0; q < kRowTiles; q++)
		{
			const uint64_t i = base + 32ull * q + lane;
			const bool on = i < re;
			id[q] = on ? __ldg(m.crow + i) : kMgNone;
			zv[q] = on ? __ldcg(&zt[i].x) : 0.0;
			rv[q] = on ? __ldcg(m.r + i) : 0.0;
		}
#pragma unroll
		for (int q = 0; q < kRowTiles; q++) ev[q] = (id[q] != kMgNone) ? __ldcg(ef + id[q]) : 0.0;
#pragma unroll
		for (int q = 0; q < kRowTiles; q++)
		{
			const uint64_t i = base + 32ull * q + lane;
			if (i < re)
			{
				const double z = zv[q] + ev[q];
				zt[i].x = z;
				local = fma(rv[q], z, local);
			}
		}
	}
	return local;
}

__device__ __forceinline__ double ld_relaxed_sys_f64(const double* p)
{
	double v;
	asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
	return v;
}
// One level vector as a rank sees it.  One GPU / replicated level: the whole vector is local.  Distributed level: cells [lo, hi)
// are this rank's (local, L2); a stencil or a prolongation may also touch the first cells behind either end, which belong to the
// adjacent rank and are read straight from its arena over NVLink (strong system-scope loads: never from a stale local cache line).
template<bool HALO>
struct LevelView
{
	const double* loc; const double* left; const double* right;
	uint32_t lo, hi;
	__device__ __forceinline__ double ld(const uint32_t j) const
	{
		if (!HALO || (j >= lo && j < hi)) return __ldcg(loc + j);
		return ld_relaxed_sys_f64(((j < lo) ? left : right) + j);
	}
};
// sum_s S[s] * e[nbr[s]] over the 3^D stencil of one cell, loads issued nine at a time
template<int K, bool HALO>
__device__ __forceinline__ double stencil_dot(const uint32_t* __restrict__ nb, const double* __restrict__ sv, const LevelView<HALO>& e)
{
	double acc = 0.0;
#pragma unroll
	for (int g = 0; g < K; g += 9)
	{
		uint32_t j[9]; double sc[9], v[9];
#pragma unroll
		for (int u = 0; u < 9; u++) { j[u] = __ldg(nb + g + u); sc[u] = __ldg(sv + g + u); }
#pragma unroll
		for (int u = 0; u < 9; u++) v[u] = (j[u] != kMgNone) ? e.ld(j[u]) : 0.0;
#pragma unroll
		for (int u = 0; u < 9; u++) acc = fma(sc[u], v[u], acc);
	}
	return acc;
}
// the same with the over-corrected prolongation formed on the fly: e[j] + gamma * ehi[parent[j]]
template<int K, bool HALO, bool HALO_HI>
__device__ __forceinline__ double stencil_dot_prolong(const uint32_t* __restrict__ nb, const double* __restrict__ sv, const LevelView<HALO>& e,
	const LevelView<HALO_HI>& ehi, const uint32_t* __restrict__ parent, const double gamma)
{
	double acc = 0.0;
#pragma unroll
	for (int g = 0; g < K; g += 9)
	{
		uint32_t j[9], pj[9]; double sc[9], v[9], vh[9];
#pragma unroll
		for (int u = 0; u < 9; u++) { j[u] = __ldg(nb + g + u); sc[u] = __ldg(sv + g + u); }
#pragma unroll
		for (int u = 0; u < 9; u++) { const bool ok = j[u] != kMgNone; v[u] = ok ? e.ld(j[u]) : 0.0; pj[u] = ok ? __ldg(parent + j[u]) : kMgNone; }
#pragma unroll
		for (int u = 0; u < 9; u++) vh[u] = (pj[u] != kMgNone) ? ehi.ld(pj[u]) : 0.0;
#pragma unroll
		for (int u = 0; u < 9; u++) acc = fma(sc[u], fma(gamma, vh[u], v[u]), acc);
	}
	return acc;
}


// several ranks: which compact cells of level l lie left of cell column `col` (asked for level 0 only: any column is a cell boundary there)
__device__ __forceinline__ uint32_t level_cell_begin(const MgLevelPtrs& lv, const int l, const uint32_t col)
{
	uint64_t idx = static_cast<uint64_t>(col >> l) * lv.colstride;
	if (idx > lv.dense) idx = lv.dense;
	return static_cast<uint32_t>(__ldg(lv.ranktab + idx));
}

// what the V-cycle needs to know about the other ranks (all in registers / shared memory of the CTA)
struct DistCtx
{
	int k;                                 // first replicated level (0 or 1)
	uint32_t lo[1];                        // this rank's level-0 cells
	uint32_t hi[1];
	const uint32_t* gather;                // [nranks + 1] (shared memory) level-0 cell ranges of every rank
	const uint32_t* part_lo;               // [nranks] (shared memory) level-1 cells that have children in rank r's slab: [part_lo, part_hi)
	const uint32_t* part_hi;
};

template<bool MG>
__device__ __forceinline__ LevelView<MG> level_view(const MgArgs& m, const DistCtx& dc, const int l, const int which /* 1: e0, 2: e1 */, const double* loc)
{
	LevelView<MG> v;
	v.loc = loc; v.left = nullptr; v.right = nullptr; v.lo = 0; v.hi = 0xffffffffu;
	if (MG && l < dc.k)
	{
		const uint64_t off = (which == 1) ? m.lv[l].off_e0 : m.lv[l].off_e1;
		const MgDist& d = m.dist;
		v.lo = dc.lo[l]; v.hi = dc.hi[l];
		v.left = (d.rank > 0) ? d.peer_vec[d.rank - 1] + off : loc;
		v.right = (d.rank + 1 < d.nranks) ? d.peer_vec[d.rank + 1] + off : loc;
	}
	return v;
}

// One V(1,1) cycle on the cell hierarchy; on entry lv[0].r and lv[0].e0 = dinv r are complete and visible (the r.r barrier),
// on return (after a grid barrier) the result is in the returned buffer of level 0.
// Levels [0, Ls) are "wide": every thread of the grid takes cells, one grid barrier per level and direction.  Levels [Ls, L) are
// small (a few hundred cells): there a grid barrier (~1.5 us + the L2 latency of the stage behind it) would cost far more than the
// work, so CTA 0 runs them alone between block barriers while the other CTAs wait at the grid barrier that hands the result back.
// Several ranks (MG), k = m.dist.k:
//   k = 1  level 0 is DISTRIBUTED: a rank smooths the cells of its own slab (any cut between cell columns) and reads the halo cells
//          of the adjacent ranks from their arenas; the restriction to level 1 is two steps — every rank sums its own children
//          into the level-1 cells they belong to and leaves these parts in its arena (cross-rank barrier), then EVERY rank adds up
//          the parts of all level-1 cells, its own from L2 and the others straight from their owners over NVLink — so level 1
//          and everything above is replicated (computed whole, on identical data, with identical results, exactly as on one GPU)
//          without any alignment of the slabs;
//   k = 0  small problems: level 0 is gathered from all ranks at the start and the whole cycle is replicated.
// On entry the barrier the caller went through must have been cross-rank.
template<int D, bool MG>
__device__ __forceinline__ const double* mg_vcycle(const CgStreamArgs& a, const MgArgs& m, const DistCtx& dc, const int L, const int Ls,
	unsigned long long& bar_target, unsigned long long& seq, unsigned& part_sel, double* red, const unsigned nblocks, unsigned long long* vprof)
{
	constexpr int K = (D == 3) ? 27 : 9, CH = 1 << D;
	const bool cta0 = blockIdx.x == 0;
	long long tp = vprof ? clock64() : 0;
	auto lap = [&](const int slot) { if (vprof) { const long long t = clock64(); vprof[slot] += static_cast<unsigned long long>(t - tp); tp = t; } };
	const uint64_t gt = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x, gs = static_cast<uint64_t>(nblocks) * blockDim.x;
	const double gamma = m.gamma;
	auto gbar = [&]() { grid_barrier(&a.sc->grid_barrier, bar_target, nblocks); };
	auto pbar = [&]() { double* q = a.partials + (part_sel & 1u) * nblocks; part_sel++; all_sum<true>(a, 0.0, q, red, bar_target, seq, nblocks); };
	// level k of every rank -> here (r and e0: what the replicated part starts from)
	auto gather_level = [&](const int k)
	{
		const MgDist& d = m.dist;
		const MgLevelPtrs& lv = m.lv[k];
		for (int r = 0; r < d.nranks; r++)
		{
			if (r == d.rank) continue;
			const double* __restrict__ pr = d.peer_vec[r] + lv.off_r;
			const double* __restrict__ pe = d.peer_vec[r] + lv.off_e0;
			const uint64_t b = dc.gather[r], e = dc.gather[r + 1];
			for (uint64_t c = b + gt; c < e; c += 4 * gs)
			{
				double vr[4], ve[4];
#pragma unroll
				for (int u = 0; u < 4; u++) { const uint64_t j = c + u * gs; vr[u] = (j < e) ? ld_relaxed_sys_f64(pr + j) : 0.0; ve[u] = (j < e) ? ld_relaxed_sys_f64(pe + j) : 0.0; }
#pragma unroll
				for (int u = 0; u < 4; u++) { const uint64_t j = c + u * gs; if (j < e) { lv.r[j] = vr[u]; lv.e0[j] = ve[u]; } }
			}
		}
		gbar();
	};
	if (MG && dc.k == 0) { gather_level(0); lap(18); }
	// ---- down: r_{l+1} = R (r_l - A_l e_l), e_{l+1} = dinv r_{l+1} ----
	for (int l = 0; l + 1 < L; l++)
	{
		const bool wide = (l + 1 < Ls);
		const bool split = MG && dc.k == 1 && l == 0; // distributed level 0 -> replicated level 1
		if (split)
		{
			// step 1: this rank's share of the restriction — the level-1 cells that have children in its slab, own children only
			// (a level-1 cell spans two cell columns, so at most two adjacent ranks share it) — into the level's `part` vector
			const MgLevelPtrs& lo = m.lv[0];
			const MgLevelPtrs& hi = m.lv[1];
			const LevelView<MG> ev = level_view<MG>(m, dc, 0, 1, lo.e0);
			const uint64_t c_b = dc.part_lo[m.dist.rank], c_e = dc.part_hi[m.dist.rank];
			const uint64_t first = gt / CH, step = gs / CH;
			const unsigned q = threadIdx.x % CH;
			for (uint64_t C0 = c_b; C0 < c_e; C0 += step) // warp-uniform trip count
			{
				const uint64_t C = C0 + first;
				const bool on = C < c_e;
				const uint32_t ch = on ? __ldg(hi.child + C * CH + q) : kMgNone;
				double res = 0.0;
				if (ch != kMgNone && ch >= dc.lo[0] && ch < dc.hi[0])
				{
					const uint64_t c = ch;
					res = __ldcg(lo.r + c) - stencil_dot<K, MG>(lo.nbr + c * K, lo.S + c * K, ev);
				}
#pragma unroll
				for (int o = 1; o < CH; o <<= 1) res += __shfl_xor_sync(0xffffffffu, res, o);
				if (on && q == 0) hi.part[C] = res;
			}
			pbar();
			// step 2: every rank assembles the whole level 1 — the parts of the (one or two) ranks that share a cell, in rank order,
			// own part from L2, the others straight from their owners over NVLink (consecutive cells: coalesced)
			const uint64_t n1 = *hi.count;
			for (uint64_t C = gt; C < n1; C += gs)
			{
				double sum = 0.0;
				for (int r = 0; r < m.dist.nranks; r++)
					if (C >= dc.part_lo[r] && C < dc.part_hi[r])
						sum += (r == m.dist.rank) ? __ldcg(hi.part + C) : ld_relaxed_sys_f64(m.dist.peer_vec[r] + hi.off_part + C);
				hi.r[C] = sum;
				hi.e0[C] = __ldg(hi.dinv + C) * sum;
			}
		}
		else if (wide || cta0)
		{
			const MgLevelPtrs& lo = m.lv[l];
			const MgLevelPtrs& hi = m.lv[l + 1];
			const uint64_t c_e = *hi.count;
			const LevelView<MG> ev = level_view<MG>(m, dc, l, 1, lo.e0);
			// CH lanes per coarse cell, one child each (the children's stencil products are independent chains of L2 round trips:
			// side by side instead of one after the other), summed over the lanes in a fixed order
			const uint64_t first = (wide ? gt : threadIdx.x) / CH, step = (wide ? gs : blockDim.x) / CH; // gs, blockDim.x are multiples of 32
			const unsigned q = threadIdx.x % CH;
			for (uint64_t C0 = 0; C0 < c_e; C0 += step) // warp-uniform trip count
			{
				const uint64_t C = C0 + first;
				const bool on = C < c_e;
				const uint32_t ch = on ? __ldg(hi.child + C * CH + q) : kMgNone;
				double res = 0.0;
				if (ch != kMgNone)
				{
					const uint64_t c = ch;
					res = __ldcg(lo.r + c) - stencil_dot<K, MG>(lo.nbr + c * K, lo.S + c * K, ev);
				}
#pragma unroll
				for (int o = 1; o < CH; o <<= 1) res += __shfl_xor_sync(0xffffffffu, res, o);
				if (on && q == 0)
				{
					hi.r[C] = res;
					hi.e0[C] = __ldg(hi.dinv + C) * res;
				}
			}
		}
		if (wide) gbar();
		else if (cta0) __syncthreads();
		lap(l);
	}
	// ---- top level: more damped-Jacobi sweeps, ping-pong ----
	const MgLevelPtrs& top = m.lv[L - 1];
	int cur = 0;
	{
		const bool wide = (L - 1 < Ls);
		const uint64_t ntop = *top.count;
		for (int t = 0; t < m.top_sweeps; t++)
		{
			if (wide || cta0)
			{
				const double* __restrict__ src = cur ? top.e1 : top.e0;
				double* __restrict__ dst = cur ? top.e0 : top.e1;
				LevelView<false> sv; sv.loc = src;
				for (uint64_t c = wide ? gt : threadIdx.x; c < ntop; c += wide ? gs : blockDim.x)
				{
					const double acc = __ldcg(top.r + c) - stencil_dot<K, false>(top.nbr + c * K, top.S + c * K, sv);
					dst[c] = fma(__ldg(top.dinv + c), acc, __ldcg(src + c));
				}
			}
			if (wide) gbar();
			else if (cta0) __syncthreads();
			cur ^= 1;
		}
		lap(16);
	}
	const double* ehi = cur ? top.e1 : top.e0;
	// ---- up: e_l <- e_l + gamma P e_{l+1}, then one Jacobi sweep; the prolongated neighbours are formed on the fly ----
	bool handed = (L - 1 < Ls); // no small level: nothing to hand back
	for (int l = L - 2; l >= 0; l--)
	{
		const bool wide = (l < Ls);
		const bool owned = MG && (l < dc.k);
		if (wide && !handed) { gbar(); handed = true; lap(17); } // CTA 0's small levels -> everybody
		if (wide || cta0)
		{
			const MgLevelPtrs& lv = m.lv[l];
			const uint64_t c_b = owned ? dc.lo[l] : 0, c_e = owned ? dc.hi[l] : *lv.count;
			const LevelView<MG> ev = level_view<MG>(m, dc, l, 1, lv.e0);
			const LevelView<MG> hv = level_view<MG>(m, dc, l + 1, 2, ehi); // levels above 0 are never distributed: local
			for (uint64_t c = c_b + (wide ? gt : threadIdx.x); c < c_e; c += wide ? gs : blockDim.x)
			{
				const double acc = __ldcg(lv.r + c) - stencil_dot_prolong<K, MG, MG>(lv.nbr + c * K, lv.S + c * K, ev, hv, lv.parent, gamma);
				const double ecur = fma(gamma, hv.ld(__ldg(lv.parent + c)), __ldcg(lv.e0 + c));
				lv.e1[c] = fma(__ldg(lv.dinv + c), acc, ecur);
			}
		}
		if (wide) gbar();                // (a distributed level 0's result is only read by its owner's rows)
		else if (cta0) __syncthreads();
		ehi = m.lv[l].e1;
		lap(20 + l);
	}
	if (!handed) { gbar(); lap(17); }
	return ehi;
}

__device__ __forceinline__ double2 ld_relaxed_sys_f64x2(const double2* p)
{
	double2 v;
	asm volatile("ld.relaxed.sys.global.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p) : "memory");
	return v;
}
// Several ranks, preconditioned solve: before an SpMV phase every rank copies the {z, p} rows of its halo — one cell column of
// each adjacent rank, which is all its matrix windows can reach — from the owners' buffers into its own (same index).  All window
// copies of the phase are then local: one coalesced pull of a few MB per iteration instead of thousands of small bulk copies from
// peer memory whose NVLink latency a two-deep ring cannot hide (101M particles on 8 GPUs: 1.3 ms of a 3.8 ms iteration).
// The caller's barrier in front must have been cross-rank (the owners have finished writing), a grid barrier must follow.
__device__ __forceinline__ void halo_pull(const CgStreamArgs& a, const MgDist& d, double2* z, const uint64_t gt, const uint64_t gs)
{
	const PeerLink& pl = a.peer;
	const double2* nb[2] = { (z == a.z0) ? pl.nb_z0[0] : pl.nb_z1[0], (z == a.z0) ? pl.nb_z0[1] : pl.nb_z1[1] };
	const uint64_t b[2] = { d.halo_lo, a.own1 }, e[2] = { a.own0, d.halo_hi };
#pragma unroll
	for (int side = 0; side < 2; side++)
	{
		if (!nb[side]) continue;
		for (uint64_t i = b[side] + gt; i < e[side]; i += 4 * gs)
		{
			double2 v[4];
#pragma unroll
			for (int u = 0; u < 4; u++) { const uint64_t j = i + u * gs; if (j < e[side]) v[u] = ld_relaxed_sys_f64x2(nb[side] + j); }
#pragma unroll
			for (int u = 0; u < 4; u++) { const uint64_t j = i + u * gs; if (j < e[side]) z[j] = v[u]; }
		}
	}
}

// 2-D runs 8 consumer warps + the producer (cg_configure caps it there): the smaller bound leaves the row passes their registers
template<int LPR, int D, bool MG>
__global__ void __launch_bounds__((D == 3 ? kMaxStreamWarps : kStreamWarps2d) * 32, 1) k_pcg_stream(CgStreamArgs a, MgArgs m)
{
	extern __shared__ __align__(128) unsigned char smem_raw[];
	__shared__ uint32_t s_gather[kMaxPeerRanks + 1], s_part_lo[kMaxPeerRanks], s_part_hi[kMaxPeerRanks];
	if (*reinterpret_cast<volatile int*>(&a.sc->error) != 0) return; // see k_cg_stream
	const StreamCta cta = stream_setup<true, MG>(a, smem_raw);
	const StreamSmem& sm = cta.sm;
	double* red = cta.red;
	const uint32_t c0 = cta.c0, c1 = cta.c1, live = cta.live;
	const uint64_t pol_matrix = cta.pol_matrix, pol_vector = cta.pol_vector;
	const uint64_t n = a.n;
	const unsigned nblocks = gridDim.x;
	// consecutive reductions must use different partial-sum buffers (a fast CTA may write its next partial while a slow one still
	// adds up the previous ones); the plain barriers of the V-cycle do not change that, so the buffer simply alternates per reduction
	unsigned part_sel = 0;
	auto next_part = [&]() { double* q = a.partials + (part_sel & 1u) * nblocks; part_sel++; return q; };
	unsigned long long bar_target = 0, seq = 0;
	uint32_t it = 0, dseq = 0;
	const bool prof_on = (a.prof != nullptr) && (threadIdx.x == 0);
	const bool meas_on = (a.cta_meas != nullptr) && (threadIdx.x == 0);
	unsigned long long spmv_cycles = 0;
	// per-stage cycle counters of the preconditioner as CTA 0 sees them (behind the per-CTA counters; mps_get_cg_profile_stages)
	unsigned long long* vprof = (prof_on && blockIdx.x == 0) ? a.prof + static_cast<size_t>(nblocks) * 8 : nullptr;

	// several ranks: this rank's cells of the distributed levels, every rank's cells of the gathered one
	DistCtx dc;
	dc.k = MG ? m.dist.k : 0;
	dc.gather = s_gather; dc.part_lo = s_part_lo; dc.part_hi = s_part_hi;
	dc.lo[0] = 0; dc.hi[0] = 0;
	if (MG)
	{
		dc.lo[0] = level_cell_begin(m.lv[0], 0, m.dist.col_b[m.dist.rank]);
		dc.hi[0] = level_cell_begin(m.lv[0], 0, m.dist.col_b[m.dist.rank + 1]);
		if (threadIdx.x <= static_cast<unsigned>(m.dist.nranks)) s_gather[threadIdx.x] = level_cell_begin(m.lv[0], 0, m.dist.col_b[threadIdx.x]);
		if (threadIdx.x < static_cast<unsigned>(m.dist.nranks) && m.levels > 1)
		{
			// level-1 cells (pairs of columns) touched by the columns [col_b[r], col_b[r + 1]) of rank r
			s_part_lo[threadIdx.x] = level_cell_begin(m.lv[1], 1, m.dist.col_b[threadIdx.x]);
			s_part_hi[threadIdx.x] = level_cell_begin(m.lv[1], 1, m.dist.col_b[threadIdx.x + 1] + 1);
		}
		__syncthreads();
	}
	// levels in use: up to the first one with few enough cells (device-side counts; the same decision in every CTA); the gathered
	// level of a multi-rank run is always one of them
	int L = 1;
	while (L < m.levels && *m.lv[L - 1].count > m.top_cells) L++;
	if (MG && L < dc.k + 1) L = dc.k + 1;
	int Ls = 0; // first level small enough for CTA 0 alone
	while (Ls < L && *m.lv[Ls].count > m.small_cells) Ls++;
	if (MG && Ls < dc.k + 1) Ls = dc.k + 1; // distributed and gathered levels are always run by the whole grid
	// this warp's rows in phase 2: a cell-aligned range, this rank's level-0 cells split evenly over all warps of the grid
	uint64_t rb, re;
	{
		const uint64_t cell_b = MG ? level_cell_begin(m.lv[0], 0, m.dist.col_b[m.dist.rank]) : 0;
		const uint64_t cells = (MG ? level_cell_begin(m.lv[0], 0, m.dist.col_b[m.dist.rank + 1]) : *m.lv[0].count) - cell_b;
		const uint64_t wpb = blockDim.x >> 5, W = static_cast<uint64_t>(nblocks) * wpb, gw = static_cast<uint64_t>(blockIdx.x) * wpb + (threadIdx.x >> 5);
		const uint64_t cb = cells / W * gw + cells % W * gw / W, ce = cells / W * (gw + 1) + cells % W * (gw + 1) / W;
		rb = m.cstart[cell_b + cb]; re = m.cstart[cell_b + ce];
	}

	// ---- r0 = b - A x ; p_prev = 0 ; rr = r0.r0 (Computer.hpp:1382-1386) ----
	double local = spmv_phase<LPR, false, MG, true>(a, sm, c0, c1, live, it, dseq, 0.0, nullptr, a.z0, pol_matrix, pol_vector);
	double rr = all_sum<MG>(a, local, next_part(), red, bar_target, seq, nblocks);
	const double rr0 = rr;
	const double tol = rr * a.eps * a.eps;     // Computer.hpp:1386
	bool converged = (tol == 0);                // Computer.hpp:1389
	double2* zprev = a.z0; // {z, p_prev}
	double2* zcur = a.z1;  // receives {z', p}
	double rz = 0.0, beta = 0.0;
	uint64_t iter = 0;
	if (!converged)
	{
		// z0 = M^-1 r0
		pcg_rows_a<true>(a, m, rb, re, 0.0, zprev);
		if (MG) all_sum<true>(a, 0.0, next_part(), red, bar_target, seq, nblocks);
		else grid_barrier(&a.sc->grid_barrier, bar_target, nblocks);
		const double* ef = mg_vcycle<D, MG>(a, m, dc, L, Ls, bar_target, seq, part_sel, red, nblocks, vprof);
		local = pcg_rows_b(m, rb, re, ef, zprev);
		rz = all_sum<MG>(a, local, next_part(), red, bar_target, seq, nblocks);
	}

	while (iter < n && !converged)
	{
		// ---- phase 1: p = z + beta p_prev ; Ap = A p ; p.Ap ----
		const long long t0 = (prof_on || meas_on) ? clock64() : 0;
		if (MG)
		{
			// the halo of {z, p_prev} from the adjacent ranks (complete: the r.z reduction in front was cross-rank), then local windows only
			halo_pull(a, m.dist, zprev, static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x, static_cast<uint64_t>(nblocks) * blockDim.x);
			grid_barrier(&a.sc->grid_barrier, bar_target, nblocks);
		}
		local = spmv_phase<LPR, true, false>(a, sm, c0, c1, live, it, dseq, beta, zprev, zcur, pol_matrix, pol_vector);
		const long long t1 = (prof_on || meas_on) ? clock64() : 0;
		spmv_cycles += static_cast<unsigned long long>(t1 - t0);
		const double pAp = all_sum<MG>(a, local, next_part(), red, bar_target, seq, nblocks);
		const double alpha = rz / pAp;

		// ---- phase 2a: x, r, r.r, Jacobi part of z, level-0 residual ----
		const long long u0 = vprof ? clock64() : 0;
		local = pcg_rows_a<false>(a, m, rb, re, alpha, zcur);
		const long long u1 = vprof ? clock64() : 0;
		const double rr_new = all_sum<MG>(a, local, next_part(), red, bar_target, seq, nblocks);
		if (vprof) { const long long u2 = clock64(); vprof[40] += static_cast<unsigned long long>(u1 - u0); vprof[42] += static_cast<unsigned long long>(u2 - u1); vprof[43] += static_cast<unsigned long long>(u0 - t1); }
		iter++;
		rr = rr_new;
		converged = (rr_new < tol);            // Computer.hpp:1407-1408
		if (converged) break;

		// ---- coarse part of z and r.z ----
		const long long t2 = prof_on ? clock64() : 0;
		const double* ef = mg_vcycle<D, MG>(a, m, dc, L, Ls, bar_target, seq, part_sel, red, nblocks, vprof);
		const long long t3 = prof_on ? clock64() : 0;
		local = pcg_rows_b(m, rb, re, ef, zcur);
		const long long u3 = vprof ? clock64() : 0;
		const double rz_new = all_sum<MG>(a, local, next_part(), red, bar_target, seq, nblocks);
		if (vprof) { vprof[41] += static_cast<unsigned long long>(u3 - t3); vprof[44] += static_cast<unsigned long long>(clock64() - u3); }
		if (prof_on)
		{
			const long long t4 = clock64();
			unsigned long long* pr = a.prof + blockIdx.x * 8;
			pr[0] += static_cast<unsigned long long>(t1 - t0);  // phase 1 (SpMV)
			pr[2] += static_cast<unsigned long long>((t2 - t1) + (t4 - t3)); // p.Ap reduction, row passes 2a / 2b, their reductions
			pr[7] += static_cast<unsigned long long>(t3 - t2);  // V-cycle (barriers included)
			pr[6] += static_cast<unsigned long long>(t4 - t0);
		}
		beta = rz_new / rz;
		rz = rz_new;
		{ double2* t = zprev; zprev = zcur; zcur = t; } // zprev now holds {z, p} of this iteration
	}

	if (meas_on) a.cta_meas[nblocks + blockIdx.x] = spmv_cycles;
	if (blockIdx.x == 0 && threadIdx.x == 0)
	{
		a.sc->z_final = (zprev == a.z1) ? 1 : 0;
		a.sc->cg_iterations = iter;
		a.sc->rr0 = rr0;
		a.sc->rr = rr;
		a.sc->cg_converged = converged ? 1 : 0;
		a.sc->mg_levels = static_cast<unsigned int>(L);
		if (!converged) atomicMax(&a.sc->error, static_cast<int>(MPS_CG_NOT_CONVERGED)); // Computer.hpp:1424-1428
	}
}

// Load balance of the next solve from this solve's own counters.  The cost model (ChunkLimits::cost_*) cannot know everything
// that makes a chunk slow (row-length spread, where its window lives, which SM runs the CTA), so every CTA reports the modelled
// cost it took and the cycles its SpMV phases needed; the smoothed cost-per-cycle of CTA b decides how much modelled cost it
// gets next time.  One block; meas = [cost taken x grid][cycles x grid].
__global__ void __launch_bounds__(256) k_cg_rebalance(const unsigned long long* __restrict__ meas, double* __restrict__ speed, double* __restrict__ frac,
	const unsigned nblocks)
{
	__shared__ double sp[256];
	__shared__ double tot[2];
	const unsigned b = threadIdx.x;
	const double cost = (b < nblocks) ? static_cast<double>(meas[b]) : 0.0;
	const double cyc = (b < nblocks) ? static_cast<double>(meas[nblocks + b]) : 0.0;
	const bool ok = (cost > 0.0) && (cyc > 0.0);
	sp[b] = ok ? cost : 0.0;
	__syncthreads();
	if (b == 0) { double t = 0.0; for (unsigned k = 0; k < nblocks; k++) t += sp[k]; tot[0] = t; }
	__syncthreads();
	sp[b] = ok ? cyc : 0.0;
	__syncthreads();
	if (b == 0) { double t = 0.0; for (unsigned k = 0; k < nblocks; k++) t += sp[k]; tot[1] = t; }
	__syncthreads();
	double v = (b < nblocks) ? speed[b] : 0.0;
	if (ok && tot[0] > 0.0 && tot[1] > 0.0)
	{
		double rel = (cost / cyc) / (tot[0] / tot[1]);
		rel = rel < 0.5 ? 0.5 : (rel > 2.0 ? 2.0 : rel);
		v = 0.5 * v + 0.5 * rel;
	}
	if (b < nblocks) speed[b] = v;
	sp[b] = v;
	__syncthreads();
	if (b == 0)
	{
		double t = 0.0;
		for (unsigned k = 0; k < nblocks; k++) t += sp[k];
		double run = 0.0;
		for (unsigned k = 0; k < nblocks; k++) { frac[k] = run / t; run += sp[k]; }
		frac[nblocks] = 1.0;
	}
}

// ---- stepwise form for multi-GPU runs: one launch per phase, the scalars of the recurrence live in device memory so that
//      NCCL all-reduces can run between the launches (mps_comm.cu drives the loop) ----
template<int LPR, int PHASE> // 0: r0 = b - A x ; 1: p, Ap, p.Ap ; 2: x, r, r.r
__global__ void __launch_bounds__(kMaxStreamWarps * 32, 1) k_cg_step(CgStreamArgs a, CgStepScalars* st)
{
	extern __shared__ __align__(128) unsigned char smem_raw[];
	const StreamCta cta = (PHASE == 0) ? stream_setup<true>(a, smem_raw) : stream_setup<false>(a, smem_raw);
	uint32_t it = 0, dseq = 0;
	double2* zprev = st->zcur_is_1 ? a.z0 : a.z1;
	double2* zcur = st->zcur_is_1 ? a.z1 : a.z0;
	double local = 0.0;
	if (PHASE == 0) local = spmv_phase<LPR, false, false>(a, cta.sm, cta.c0, cta.c1, cta.live, it, dseq, 0.0, nullptr, a.z0, cta.pol_matrix, cta.pol_vector);
	if (PHASE == 1) local = spmv_phase<LPR, true, false>(a, cta.sm, cta.c0, cta.c1, cta.live, it, dseq, st->beta, zprev, zcur, cta.pol_matrix, cta.pol_vector);
	if (PHASE == 2) local = update_rows(a, cta.row0, cta.row1, st->rr / st->pAp, zprev, zcur);
	local = block_sum_n(local, cta.red);
	if (threadIdx.x == 0) a.partials[blockIdx.x] = local;
}

// fixed-order sum of the per-CTA partials -> *dst (one block)
__global__ void __launch_bounds__(256) k_cg_reduce(const double* __restrict__ partials, unsigned nblocks, double* dst)
{
	__shared__ double red[256 / 32 + 1];
	double v = 0.0;
	for (unsigned k = threadIdx.x; k < nblocks; k += 256) v += partials[k];
	v = block_sum_n(v, red);
	if (threadIdx.x == 0) *dst = v;
}

// scalar recurrences between the phases (one thread) — Computer.hpp:1386-1417
__global__ void k_cg_after_init(CgStepScalars* st, double eps)
{
	st->rr0 = st->rr;
	st->tol = st->rr * eps * eps;
	st->converged = (st->tol == 0) ? 1 : 0;
	st->beta = 0.0;
	st->iter = 0;
	st->zcur_is_1 = 1; // z0 = {r0, 0} is "previous"
}
__global__ void k_cg_after_iter(CgStepScalars* st)
{
	st->iter += 1;
	st->converged = (st->rr_new < st->tol) ? 1 : 0;
	if (!st->converged) st->beta = st->rr_new / st->rr;
	st->zcur_is_1 ^= 1;
	st->rr = st->rr_new;
}
__global__ void k_cg_finish(const CgStepScalars* st, DevScalars* sc)
{
	sc->z_final = st->zcur_is_1 ? 0 : 1; // the buffer written last is the current "previous"
	sc->cg_iterations = st->iter;
	sc->rr0 = st->rr0;
	sc->rr = st->rr;
	sc->cg_converged = st->converged;
	if (!st->converged) atomicMax(&sc->error, static_cast<int>(MPS_CG_NOT_CONVERGED));
}

inline uint32_t round_up(uint32_t v, uint32_t m) { return (v + m - 1) / m * m; }

struct StreamGeometry
{
	uint32_t blob_stage_bytes, window_stage, stage_bytes;
	size_t fixed_bytes;
};

StreamGeometry stream_geometry(const ChunkLimits& lim)
{
	StreamGeometry g;
	g.blob_stage_bytes = round_up(chunk_blob_bytes(lim.max_rows, lim.max_nnz), 128);
	g.window_stage = round_up(lim.max_window, 16);
	g.stage_bytes = g.blob_stage_bytes + 24u * g.window_stage;
	g.fixed_bytes = 2u * kDescBatch * sizeof(ChunkDesc) + (kGroups + 1u) * 8u * 8u /* barriers, <= 8 stages */ + 16u + (kMaxStreamWarps + 1) * 8u + 32u + 16u;
	return g;
}

struct StreamLaunch
{
	CgStreamArgs a;
	size_t smem_bytes;
	unsigned grid;
	int threads;
};

cudaError_t prepare_stream(mps_solver* s, StreamLaunch& L)
{
	CgBuffers& c = s->cg;
	const StreamGeometry g = stream_geometry(c.limits);
	L.smem_bytes = static_cast<size_t>(c.stages) * g.stage_bytes + g.fixed_bytes;
	L.threads = (c.consumer_warps + c.producers) * 32;
	L.grid = static_cast<unsigned>(s->sm_count); // one CTA per SM: the ring wants the whole shared memory
	cudaError_t e = c.partials.ensure(2ull * L.grid + 8, s->stream);
	if (e != cudaSuccess) return e;
	if (!c.cta_frac.p)
	{
		// uniform split until the kernel has measured itself
		if ((e = c.cta_frac.ensure(L.grid + 1, s->stream)) != cudaSuccess) return e;
		if ((e = c.cta_speed.ensure(L.grid, s->stream)) != cudaSuccess) return e;
		if ((e = c.cta_meas.ensure(2ull * L.grid, s->stream)) != cudaSuccess) return e;
		std::vector<double> f(L.grid + 1), one(L.grid, 1.0);
		for (unsigned b = 0; b <= L.grid; b++) f[b] = static_cast<double>(b) / L.grid;
		if ((e = cudaMemcpyAsync(c.cta_frac.p, f.data(), f.size() * sizeof(double), cudaMemcpyHostToDevice, s->stream)) != cudaSuccess) return e;
		if ((e = cudaMemcpyAsync(c.cta_speed.p, one.data(), one.size() * sizeof(double), cudaMemcpyHostToDevice, s->stream)) != cudaSuccess) return e;
		if ((e = cudaStreamSynchronize(s->stream)) != cudaSuccess) return e; // the host vectors go out of scope
		if (const char* v = std::getenv("MPS_CG_ADAPTIVE")) c.adaptive = std::atoi(v) != 0;
		if (L.grid > 256) c.adaptive = false; // k_cg_rebalance is one block of 256 threads, one per CTA
	}
	CgStreamArgs& a = L.a;
	a.n = c.n; a.desc = c.live.p; a.blobs = c.blobs.p; a.b = c.b.p; a.x = c.x.p; a.z0 = reinterpret_cast<double2*>(c.z0.p); a.z1 = reinterpret_cast<double2*>(c.z1.p);
	a.ap = c.ap.p; a.r = nullptr; a.partials = c.partials.p; a.sc = s->d_sc; a.eps = s->env.eps;
	a.blob_stage_bytes = g.blob_stage_bytes; a.window_stage = g.window_stage; a.stages = static_cast<uint32_t>(c.stages);
	// entries <= neighbour entries + rows: stream the matrix past L2 when it cannot stay resident next to the vectors
	a.l2_stream = ((s->nbr_total + (s->own1() - s->own0())) * 10ull + c.n * 48ull > (96ull << 20)) ? 1 : 0;
	a.producers = static_cast<uint32_t>(c.producers);
	a.prof = nullptr;
	a.cta_frac = c.cta_frac.p; a.cta_meas = nullptr;
	a.own0 = s->own0(); a.own1 = s->own1();
	a.peer = PeerLink{};
	return cudaSuccess;
}

template<int LPR, bool MG>
cudaError_t launch_stream(mps_solver* s)
{
	CgBuffers& c = s->cg;
	StreamLaunch L;
	cudaError_t e = prepare_stream(s, L);
	if (e != cudaSuccess) return e;
	if (MG)
	{
		e = comm_prepare_link(s); // halo extents of this assembly -> rows to push, neighbours' buffers, mailboxes
		if (e != cudaSuccess) return e;
		L.a.peer = s->comm.link;
	}
	e = cudaFuncSetAttribute(k_cg_stream<LPR, MG>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(L.smem_bytes));
	if (e != cudaSuccess) return e;
	int per_sm = 0;
	e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_cg_stream<LPR, MG>, L.threads, L.smem_bytes);
	if (e != cudaSuccess) return e;
	if (per_sm < 1) return cudaErrorLaunchOutOfResources;
	e = cudaMemsetAsync(&s->d_sc->grid_barrier, 0, sizeof(unsigned long long), s->stream);
	if (e != cudaSuccess) return e;
	if (s->cg_profile)
	{
		e = c.prof.ensure(8ull * L.grid, s->stream);
		if (e != cudaSuccess) return e;
		e = cudaMemsetAsync(c.prof.p, 0, 8ull * L.grid * sizeof(unsigned long long), s->stream);
		if (e != cudaSuccess) return e;
		L.a.prof = c.prof.p;
		c.prof_blocks = L.grid;
		c.prof_stages = false;
	}
	if (c.adaptive)
	{
		e = cudaMemsetAsync(c.cta_meas.p, 0, 2ull * L.grid * sizeof(unsigned long long), s->stream);
		if (e != cudaSuccess) return e;
		L.a.cta_meas = c.cta_meas.p;
	}
	void* params[] = { &L.a };
	s->stats.kernel_launches += 1;
	e = cudaLaunchCooperativeKernel(reinterpret_cast<void*>(k_cg_stream<LPR, MG>), dim3(L.grid), dim3(L.threads), params, L.smem_bytes, s->stream);
	if (e != cudaSuccess) return e;
	if (c.adaptive)
	{
		k_cg_rebalance<<<1, 256, 0, s->stream>>>(c.cta_meas.p, c.cta_speed.p, c.cta_frac.p, L.grid);
		s->stats.kernel_launches += 1;
	}
	return cudaGetLastError();
}

// the preconditioned solve: same launch geometry as k_cg_stream + the level tables of mps_mg.cu.  MG (several ranks over peer
// memory): the level vectors live in the "mg" section of every rank's arena (MgDist, mps_device.cuh)
template<int LPR, int D, bool MG>
cudaError_t launch_pcg(mps_solver* s)
{
	CgBuffers& c = s->cg;
	MgBuffers& g = s->mg;
	StreamLaunch L;
	cudaError_t e = prepare_stream(s, L);
	if (e != cudaSuccess) return e;
	if (MG)
	{
		e = comm_prepare_link(s);
		if (e != cudaSuccess) return e;
		L.a.peer = s->comm.link;
	}
	L.a.r = c.r.p;
	MgArgs m{};
	m.levels = g.levels; m.top_sweeps = g.top_sweeps; m.top_cells = g.top_cells; m.small_cells = g.small_cells; m.gamma = g.gamma;
	m.crow = g.crow.p; m.cstart = g.cstart.p; m.dinv0 = g.dinv0.p; m.r = c.r.p;
	double* arena_vec = MG ? comm_mg_section(s, s->comm.rank) : nullptr;
	if (MG && (!arena_vec || !g.in_arena)) return cudaErrorInvalidValue;
	for (int l = 0; l < g.levels; l++)
	{
		MgLevelBufs& b = g.lv[l];
		MgLevelPtrs& q = m.lv[l];
		q.count = b.rank.p + b.dense; q.S = b.S.p; q.nbr = b.nbr.p; q.dinv = b.dinv.p; q.child = b.child.p; q.parent = b.parent.p;
		q.ranktab = b.rank.p; q.colstride = b.dense / static_cast<uint64_t>(b.dims[0]); q.dense = b.dense;
		q.off_r = g.vec_off[l][0]; q.off_e0 = g.vec_off[l][1]; q.off_e1 = g.vec_off[l][2]; q.off_part = g.vec_off[l][3];
		if (MG) { q.r = arena_vec + q.off_r; q.e0 = arena_vec + q.off_e0; q.e1 = arena_vec + q.off_e1; q.part = arena_vec + q.off_part; }
		else { q.r = b.r.p; q.e0 = b.e0.p; q.e1 = b.e1.p; q.part = nullptr; }
	}
	if (MG)
	{
		MgDist& d = m.dist;
		d.on = 1; d.k = g.k_dist; d.rank = s->comm.rank; d.nranks = s->comm.nranks;
		if (!s->slabs_set()) return cudaErrorInvalidValue;
		for (int r = 0; r <= d.nranks; r++) d.col_b[r] = s->col_b[r];
		for (int r = 0; r < d.nranks; r++) { d.peer_vec[r] = comm_mg_section(s, r); if (!d.peer_vec[r]) return cudaErrorInvalidValue; }
		// window ranges start and end on even slots: the halo is widened to even bounds (the vectors have slack behind the last row)
		if (s->halo_lo.size() != static_cast<size_t>(d.nranks)) return cudaErrorInvalidValue;
		d.halo_lo = s->halo_lo[d.rank] & ~1ull;
		d.halo_hi = (s->halo_hi[d.rank] + 1ull) & ~1ull;
		if (d.halo_lo > L.a.own0) d.halo_lo = L.a.own0;
		if (d.halo_hi < L.a.own1) d.halo_hi = L.a.own1;
	}
	e = cudaFuncSetAttribute(k_pcg_stream<LPR, D, MG>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(L.smem_bytes));
	if (e != cudaSuccess) return e;
	int per_sm = 0;
	e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_pcg_stream<LPR, D, MG>, L.threads, L.smem_bytes);
	if (e != cudaSuccess) return e;
	if (per_sm < 1) return cudaErrorLaunchOutOfResources;
	e = cudaMemsetAsync(&s->d_sc->grid_barrier, 0, sizeof(unsigned long long), s->stream);
	if (e != cudaSuccess) return e;
	if (s->cg_profile)
	{
		e = c.prof.ensure(8ull * L.grid + kProfStages, s->stream);
		if (e != cudaSuccess) return e;
		e = cudaMemsetAsync(c.prof.p, 0, (8ull * L.grid + kProfStages) * sizeof(unsigned long long), s->stream);
		if (e != cudaSuccess) return e;
		L.a.prof = c.prof.p;
		c.prof_blocks = L.grid;
		c.prof_stages = true;
	}
	if (c.adaptive)
	{
		e = cudaMemsetAsync(c.cta_meas.p, 0, 2ull * L.grid * sizeof(unsigned long long), s->stream);
		if (e != cudaSuccess) return e;
		L.a.cta_meas = c.cta_meas.p;
	}
	void* params[] = { &L.a, &m };
	s->stats.kernel_launches += 1;
	e = cudaLaunchCooperativeKernel(reinterpret_cast<void*>(k_pcg_stream<LPR, D, MG>), dim3(L.grid), dim3(L.threads), params, L.smem_bytes, s->stream);
	if (e != cudaSuccess) return e;
	if (c.adaptive)
	{
		k_cg_rebalance<<<1, 256, 0, s->stream>>>(c.cta_meas.p, c.cta_speed.p, c.cta_frac.p, L.grid);
		s->stats.kernel_launches += 1;
	}
	return cudaGetLastError();
}

template<int D, bool MG>
cudaError_t launch_pcg_dim(mps_solver* s)
{
	switch (s->cg.lanes_per_row)
	{
	case 1: return launch_pcg<1, D, MG>(s);
	case 2: return launch_pcg<2, D, MG>(s);
	case 4: return launch_pcg<4, D, MG>(s);
	default: return launch_pcg<8, D, MG>(s);
	}
}
template<bool MG>
cudaError_t launch_pcg_lpr(mps_solver* s) { return s->env.dim == 3 ? launch_pcg_dim<3, MG>(s) : launch_pcg_dim<2, MG>(s); }

template<bool MG>
cudaError_t launch_stream_lpr(mps_solver* s)
{
	switch (s->cg.lanes_per_row)
	{
	case 1: return launch_stream<1, MG>(s);
	case 2: return launch_stream<2, MG>(s);
	case 4: return launch_stream<4, MG>(s);
	default: return launch_stream<8, MG>(s);
	}
}

template<int LPR, int PHASE>
cudaError_t launch_step_phase(mps_solver* s, CgStepScalars* st)
{
	StreamLaunch L;
	cudaError_t e = prepare_stream(s, L);
	if (e != cudaSuccess) return e;
	e = cudaFuncSetAttribute(k_cg_step<LPR, PHASE>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(L.smem_bytes));
	if (e != cudaSuccess) return e;
	k_cg_step<LPR, PHASE><<<L.grid, L.threads, L.smem_bytes, s->stream>>>(L.a, st);
	s->stats.kernel_launches += 1;
	return cudaGetLastError();
}

template<int LPR>
cudaError_t launch_step_lpr(mps_solver* s, int phase, CgStepScalars* st)
{
	switch (phase)
	{
	case 0: return launch_step_phase<LPR, 0>(s, st);
	case 1: return launch_step_phase<LPR, 1>(s, st);
	default: return launch_step_phase<LPR, 2>(s, st);
	}
}

template<int LPR>
cudaError_t launch_lpr(mps_solver* s, CgArgs& args, unsigned want_blocks)
{
	int per_sm = 0;
	cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_cg_solve<LPR>, kCgThreads, 0);
	if (e != cudaSuccess) return e;
	if (per_sm < 1) return cudaErrorLaunchOutOfResources;
	unsigned grid = static_cast<unsigned>(per_sm) * static_cast<unsigned>(s->sm_count);
	if (want_blocks < grid) grid = want_blocks;
	if (grid < 1) grid = 1;
	e = s->cg.partials.ensure(2ull * grid, s->stream);
	if (e != cudaSuccess) return e;
	args.partials = s->cg.partials.p;
	void* params[] = { &args };
	s->stats.kernel_launches += 1;
	return cudaLaunchCooperativeKernel(reinterpret_cast<void*>(k_cg_solve<LPR>), dim3(grid), dim3(kCgThreads), params, 0, s->stream);
}

} // namespace

// the preconditioner needs the cell hierarchy of an assembled system; several ranks need each other's arenas mapped (peer memory)
bool mg_wanted(const mps_solver* s) { return s->mg.on && s->cg.chunked && (!s->comm.on || s->comm.peer_mode == 1); }
bool mg_active(const mps_solver* s) { return mg_wanted(s) && !s->cg.external; }

cudaError_t launch_cg(mps_solver* s)
{
	CgBuffers& c = s->cg;
	if (c.n == 0)
	{
		// an empty system is converged by definition (residual0 == 0)
		return cudaSuccess;
	}
	// multi-GPU: the persistent kernel coupled through peer memory when the ranks could map each other's arenas, else NCCL stepwise
	if (mg_active(s)) return s->comm.on ? launch_pcg_lpr<true>(s) : launch_pcg_lpr<false>(s);
	if (c.chunked && !c.external && s->comm.on) return (s->comm.peer_mode == 1) ? launch_stream_lpr<true>(s) : comm_cg_solve(s);
	if (c.chunked && !c.external) return launch_stream_lpr<false>(s);
	CgArgs args;
	args.n = c.n; args.rowptr = c.rowptr.p; args.col = c.col.p; args.val = c.val.p; args.b = c.b.p;
	args.x = c.x.p; args.r = c.r.p; args.pbuf0 = c.p0.p; args.pbuf1 = c.p1.p; args.ap = c.ap.p;
	args.partials = nullptr; args.sc = s->d_sc; args.eps = s->env.eps;
	// lanes per row: ~21 non-zeros per row in 2-D (r_e = 2.4 l0), ~57 in 3-D
	const int lpr = (s->env.dim == 3 && !c.external) ? 16 : 8;
	const unsigned want = blocks_for(c.n * static_cast<uint64_t>(lpr), kCgThreads);
	return lpr == 16 ? launch_lpr<16>(s, args, want) : launch_lpr<8>(s, args, want);
}

// ---- stepwise interface used by the multi-GPU driver loop (mps_comm.cu) -------------------------------------------------
cudaError_t launch_cg_step(mps_solver* s, int phase, CgStepScalars* st)
{
	switch (s->cg.lanes_per_row)
	{
	case 1: return launch_step_lpr<1>(s, phase, st);
	case 2: return launch_step_lpr<2>(s, phase, st);
	case 4: return launch_step_lpr<4>(s, phase, st);
	default: return launch_step_lpr<8>(s, phase, st);
	}
}
cudaError_t launch_cg_reduce(mps_solver* s, double* dst)
{
	k_cg_reduce<<<1, 256, 0, s->stream>>>(s->cg.partials.p, static_cast<unsigned>(s->sm_count), dst);
	s->stats.kernel_launches += 1;
	return cudaGetLastError();
}
cudaError_t launch_cg_scalars(mps_solver* s, int which, CgStepScalars* st)
{
	if (which == 0) k_cg_after_init<<<1, 1, 0, s->stream>>>(st, s->env.eps);
	else if (which == 1) k_cg_after_iter<<<1, 1, 0, s->stream>>>(st);
	else k_cg_finish<<<1, 1, 0, s->stream>>>(st, s->d_sc);
	s->stats.kernel_launches += 1;
	return cudaGetLastError();
}

// Chunk limits for this environment.  Rows per chunk = consumer threads / lanes per row (every consumer sub-warp owns one row
// of the chunk); entries and window are sized for interior rows (~21 entries in 2-D at r_e = 2.4 l0, ~57 in 3-D) with some
// head-room, and a single row must always fit (<= 3^D cells x capacity neighbours); the ring takes as many stages as fit
// the SM's shared memory.  MPS_CG_WARPS / MPS_CG_STAGES override the defaults (tuning).
cudaError_t cg_configure(mps_solver* s)
{
	CgBuffers& c = s->cg;
	const int D = s->env.dim;
	const uint32_t stencil = (D == 3) ? 27u : 9u;
	const uint32_t row_max = stencil * s->env.cell_cap;
	int warps = (D == 3) ? 16 : 8; // measured on B200 (scripts/gpu_tune.sh): 2-D 128-row chunks x 4 stages, 3-D 64-row chunks
	if (const char* v = std::getenv("MPS_CG_WARPS")) warps = std::atoi(v);
	warps = warps / static_cast<int>(kGroups) * static_cast<int>(kGroups);
	if (warps < static_cast<int>(kGroups)) warps = kGroups;
	if (warps > kMaxStreamWarps - kMaxProducers) warps = kMaxStreamWarps - kMaxProducers;
	if (D == 2 && warps > kStreamWarps2d - kMaxProducers) warps = kStreamWarps2d - kMaxProducers;
	c.consumer_warps = warps;
	// producer warps: one bulk copy occupies its issuing warp for ~560 cycles (tools/tma_bench.cu) and a chunk needs 4 (2-D) / 10 (3-D)
	c.producers = kMaxProducers;
	if (const char* v = std::getenv("MPS_CG_PRODUCERS")) { const int w = std::atoi(v); if (w >= 1 && w <= kMaxProducers) c.producers = w; }
	c.lanes_per_row = (D == 3) ? 4 : 1;
	if (const char* v = std::getenv("MPS_CG_LPR")) { const int w = std::atoi(v); if (w == 1 || w == 2 || w == 4 || w == 8) c.lanes_per_row = w; }
	ChunkLimits lim;
	lim.max_rows = static_cast<uint32_t>(warps) / kGroups * 32u / static_cast<uint32_t>(c.lanes_per_row);
	if (lim.max_rows > 256) lim.max_rows = 256;
	// mean neighbours within r_e on the lattice: 2-D pi (r_e/l0)^2, 3-D 4/3 pi (r_e/l0)^3 ; + diagonal + 5 % head-room
	const double q = s->env.r_e / s->env.l0;
	const double k_mean = (D == 3 ? 4.18879 * q * q * q : 3.14159 * q * q) * 1.05 + 2.0;
	lim.max_nnz = round_up(static_cast<uint32_t>(lim.max_rows * k_mean), 8);
	if (lim.max_nnz < row_max + 1) lim.max_nnz = round_up(row_max + 1, 8);
	// window ~ (2 D - 1 ... 3^(D-1)) columns x (rows + 2 cells of mean occupancy)
	const double cell = s->env.neighbor_length / s->env.l0;
	const double occ = (D == 3) ? cell * cell * cell : cell * cell;
	const double cols = (D == 3) ? 9.0 : 3.0;
	lim.max_window = round_up(static_cast<uint32_t>(cols * (lim.max_rows + 2.0 * occ) * 1.15) + 32u, 16);
	if (lim.max_window < row_max + 2 * kMaxRanges) lim.max_window = round_up(row_max + 2 * kMaxRanges, 16);
	lim.cost_fixed = 12000; lim.cost_per_nnz = 1;
	if (const char* v = std::getenv("MPS_CG_COST_FIXED")) lim.cost_fixed = static_cast<uint32_t>(std::atoi(v));
	c.limits = lim;
	const StreamGeometry g = stream_geometry(lim);
	const size_t budget = 225u * 1024u;
	int stages = static_cast<int>((budget - g.fixed_bytes) / g.stage_bytes);
	if (stages > 8) stages = 8;
	if (D == 2 && stages > 4) stages = 4; // measured on B200, 1 M particles: 5 stages 7.42 ms per solve, 4 stages 7.05 ms (more look-ahead, more L2 thrash)
	if (const char* v = std::getenv("MPS_CG_STAGES")) { const int w = std::atoi(v); if (w >= 2 && w <= stages) stages = w; }
	// any depth >= 2 works, odd ones included: every (stage, consumer group) pair has its own "full" barrier (spmv_phase)
	c.stages = stages;
	// u16 row offsets / columns and the shared memory of one SM bound what can be streamed; otherwise the generic kernel
	c.chunked = (lim.max_nnz < 65536u) && (lim.max_window < 65536u) && (stages >= 2);
	if (const char* v = std::getenv("MPS_CG_GENERIC")) { if (std::atoi(v) != 0) c.chunked = false; }
	return cudaSuccess;
}

} // namespace mps
