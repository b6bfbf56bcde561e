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
#include <cooperative_groups.h>

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

cudaError_t launch_cg(mps_solver* s)
{
	CgBuffers& c = s->cg;
	if (c.n == 0)
	{
		// an empty system is converged by definition (residual0 == 0)
		return cudaSuccess;
	}
	CgArgs args;
	args.n = c.n; args.rowptr = c.rowptr.p; args.col = c.col.p; args.val = c.val.p; args.b = c.b.p;
	args.x = c.x.p; args.r = c.r.p; args.pbuf0 = c.p0.p; args.pbuf1 = c.p1.p; args.ap = c.ap.p;
	args.partials = nullptr; args.sc = s->d_sc; args.eps = s->env.eps;
	// lanes per row: ~21 non-zeros per row in 2-D (r_e = 2.4 l0), ~57 in 3-D
	const int lpr = (s->env.dim == 3 && !c.external) ? 16 : 8;
	const unsigned want = blocks_for(c.n * static_cast<uint64_t>(lpr), kCgThreads);
	return lpr == 16 ? launch_lpr<16>(s, args, want) : launch_lpr<8>(s, args, want);
}

} // namespace mps
