// mps_mg.cu — set-up of the multigrid preconditioner of the pressure-Poisson solve (DESIGN.md "PCG").
//
// The reference solves the PPE with plain CG (Computer.hpp:1359-1429); its iteration count grows with resolution (820 at
// 68 k rows, 1 480 at 270 k, 1 250-2 900 per step at 1 M).  Here the same system, with the same stopping rule, is solved by CG
// preconditioned with
//        M^-1 = D^-1 + P1 V P1^T
// D = diag(A); P1 = piecewise-constant prolongation from the CELLS of the neighbour grid (particles are cell-sorted, so the
// rows of one cell are one contiguous slot range); V = one V(1,1) cycle (damped Jacobi, omega; over-corrected coarse-grid
// corrections, gamma) on the hierarchy obtained by merging 2^D cells per level, with Galerkin operators R A P throughout.
// tests/studies/precond_study2.py: 43-46 iterations where plain CG needs 821 / 1 477 (2-D) and 376 (3-D, 150 k rows).
//
// Because aggregates are grid cells, every coarse operator is a 3^D-point stencil on a sparse regular grid:
//   level 0  = occupied cells of the neighbour grid (compact ids in cell-key order), level l+1 = occupied 2^D blocks of level l
//   S[c][s]  = sum of a_ij over i in cell c, j in cell c + offset(s)        (s = 0 .. 3^D - 1, x-major, centre at 3^D / 2)
//   nbr[c][s]= compact id of cell c + offset(s) or kMgNone;  child / parent = the 2^D-to-1 maps between levels
// This file builds all of that on the device, every step, from the cell table of the sort (mps_grid.cu) and the per-row stencil
// sums k_ppe_fill leaves behind (mps_gather.cu).  The solve itself is part of the persistent kernel in mps_cg.cu.
// Integer / streaming work: dense per-level flag + rank arrays (4 + 8 B per cell of the bounding grid), compact per-level
// tables (3^D x 12 B per occupied cell); negligible next to one sweep over the fine matrix.
#include <cstdlib>

#include "mps_solver.h"

namespace mps {
namespace {

constexpr int kThreads = 256;

struct Dims { long long n[3]; };

template<int D> __device__ __forceinline__ void decode(uint64_t key, const Dims& d, long long* c)
{
#pragma unroll
	for (int a = D - 1; a >= 0; a--) { c[a] = static_cast<long long>(key % static_cast<uint64_t>(d.n[a])); key /= static_cast<uint64_t>(d.n[a]); }
}
template<int D> __device__ __forceinline__ uint64_t encode(const long long* c, const Dims& d)
{
	uint64_t k = 0;
#pragma unroll
	for (int a = 0; a < D; a++) k = k * static_cast<uint64_t>(d.n[a]) + static_cast<uint64_t>(c[a]);
	return k;
}
template<int D> __device__ __forceinline__ bool inside(const long long* c, const Dims& d)
{
	bool ok = true;
#pragma unroll
	for (int a = 0; a < D; a++) ok = ok && (c[a] >= 0) && (c[a] < d.n[a]);
	return ok;
}

// level 0: a cell of the neighbour grid is occupied when the sort put at least one particle into it
__global__ void __launch_bounds__(kThreads) k_mg_occ0(uint64_t ncells, const uint32_t* __restrict__ cell_count, uint32_t* __restrict__ flag)
{
	const uint64_t c = static_cast<uint64_t>(blockIdx.x) * kThreads + threadIdx.x;
	if (c < ncells) flag[c] = cell_count[c] ? 1u : 0u;
}

// level l + 1: a block is occupied when one of its 2^D children is (gather form: no atomics, no ordering issues)
template<int D>
__global__ void __launch_bounds__(kThreads) k_mg_occ_up(uint64_t dense_hi, Dims lo, Dims hi, const uint32_t* __restrict__ flag_lo, uint32_t* __restrict__ flag_hi)
{
	const uint64_t C = static_cast<uint64_t>(blockIdx.x) * kThreads + threadIdx.x;
	if (C >= dense_hi) return;
	long long cc[3]; decode<D>(C, hi, cc);
	uint32_t any = 0;
#pragma unroll
	for (int q = 0; q < (1 << D); q++)
	{
		long long c[3];
#pragma unroll
		for (int a = 0; a < D; a++) c[a] = 2 * cc[a] + ((q >> (D - 1 - a)) & 1);
		if (inside<D>(c, lo)) any |= flag_lo[encode<D>(c, lo)];
	}
	flag_hi[C] = any;
}

// ---- compact ids, neighbour / parent tables and child tables of ALL levels in one launch each (the upper levels are a few hundred
//      cells: a launch per level and table was ~25 launches of a few microseconds per step).  Block b works on level l with
//      bb[l] <= b < bb[l + 1].
//      compact:  dense -> compact, key[rank] = dense index of the occupied cell
//      topology: per occupied cell its 3^D neighbours (compact ids) and its parent block at the next level
//      children: per occupied block of level l + 1 its 2^D children at level l ----
struct MgLevelDev
{
	uint64_t dense, bound;
	Dims dims;
	const uint32_t* flag; const uint64_t* rank;
	uint32_t* key; uint32_t* nbr; uint32_t* parent; uint32_t* child;
};
struct MgBatch
{
	int levels;
	MgLevelDev lv[kMgMaxLevels];
};
struct MgBlocks { uint32_t bb[kMgMaxLevels + 1]; };

__device__ __forceinline__ int level_of_block(const MgBlocks& b, const int levels, uint32_t& local)
{
	int l = 0;
	while (l + 1 < levels && blockIdx.x >= b.bb[l + 1]) l++;
	local = blockIdx.x - b.bb[l];
	return l;
}

__global__ void __launch_bounds__(kThreads) k_mg_compact_all(MgBatch m, MgBlocks b)
{
	uint32_t lb; const int l = level_of_block(b, m.levels, lb);
	const MgLevelDev& v = m.lv[l];
	const uint64_t c = static_cast<uint64_t>(lb) * kThreads + threadIdx.x;
	if (c < v.dense && v.flag[c]) v.key[v.rank[c]] = static_cast<uint32_t>(c);
}

template<int D>
__global__ void __launch_bounds__(kThreads) k_mg_topology_all(MgBatch m, MgBlocks b)
{
	constexpr int K = (D == 3) ? 27 : 9;
	uint32_t lb; const int l = level_of_block(b, m.levels, lb);
	const MgLevelDev& v = m.lv[l];
	const bool last = (l + 1 == m.levels);
	const uint64_t cc = static_cast<uint64_t>(lb) * kThreads + threadIdx.x;
	if (cc >= v.bound || cc >= v.rank[v.dense]) return; // rank[dense] = occupied cells of the level
	const Dims lo = v.dims;
	long long c[3]; decode<D>(v.key[cc], lo, c);
	int s = 0;
	for (int ox = -1; ox <= 1; ox++)
		for (int oy = (D == 3 ? -1 : 0); oy <= (D == 3 ? 1 : 0); oy++)
			for (int oz = -1; oz <= 1; oz++, s++)
			{
				long long q[3];
				q[0] = c[0] + ox;
				if (D == 3) q[1] = c[1] + oy;
				q[D - 1] = c[D - 1] + oz;
				uint32_t id = kMgNone;
				if (inside<D>(q, lo))
				{
					const uint64_t k = encode<D>(q, lo);
					if (v.flag[k]) id = static_cast<uint32_t>(v.rank[k]);
				}
				v.nbr[cc * K + s] = id;
			}
	if (!last)
	{
		const MgLevelDev& h = m.lv[l + 1];
		long long p[3];
#pragma unroll
		for (int a = 0; a < D; a++) p[a] = c[a] >> 1;
		v.parent[cc] = static_cast<uint32_t>(h.rank[encode<D>(p, h.dims)]);
	}
}

// blocks of level l >= 1 (bb counts levels 1 .. levels - 1 as 0 .. levels - 2)
template<int D>
__global__ void __launch_bounds__(kThreads) k_mg_children_all(MgBatch m, MgBlocks b)
{
	constexpr int CH = 1 << D;
	uint32_t lb; const int l = level_of_block(b, m.levels - 1, lb) + 1;
	const MgLevelDev& hi = m.lv[l];
	const MgLevelDev& lo = m.lv[l - 1];
	const uint64_t C = static_cast<uint64_t>(lb) * kThreads + threadIdx.x;
	if (C >= hi.bound || C >= hi.rank[hi.dense]) return;
	long long cc[3]; decode<D>(hi.key[C], hi.dims, cc);
#pragma unroll
	for (int q = 0; q < CH; q++)
	{
		long long c[3];
#pragma unroll
		for (int a = 0; a < D; a++) c[a] = 2 * cc[a] + ((q >> (D - 1 - a)) & 1);
		uint32_t id = kMgNone;
		if (inside<D>(c, lo.dims))
		{
			const uint64_t k = encode<D>(c, lo.dims);
			if (lo.flag[k]) id = static_cast<uint32_t>(lo.rank[k]);
		}
		hi.child[C * CH + q] = id;
	}
}

// rows -> compact level-0 cell, and the first row of every occupied cell (+ one past the last row that lies in a cell)
__global__ void __launch_bounds__(kThreads) k_mg_rows(uint64_t n, uint32_t ncells, const uint32_t* __restrict__ skey, const uint64_t* __restrict__ rank0,
	uint32_t* __restrict__ crow)
{
	const uint64_t i = static_cast<uint64_t>(blockIdx.x) * kThreads + threadIdx.x;
	if (i >= n) return;
	const uint32_t k = skey[i];
	crow[i] = (k < ncells) ? static_cast<uint32_t>(rank0[k]) : kMgNone;
}
__global__ void __launch_bounds__(kThreads) k_mg_cstart(uint64_t bound, const uint64_t* __restrict__ count, uint64_t ncells, const uint32_t* __restrict__ key0,
	const uint64_t* __restrict__ cell_start, uint64_t* __restrict__ cstart)
{
	const uint64_t cc = static_cast<uint64_t>(blockIdx.x) * kThreads + threadIdx.x;
	const uint64_t cnt = *count;
	if (cc > bound || cc > cnt) return;
	cstart[cc] = (cc < cnt) ? cell_start[key0[cc]] : cell_start[ncells];
}

// level-0 operator: S[c][s] = sum over the ACTIVE rows of cell c of the per-row stencil sums written by k_ppe_fill
// (row_s[s * stride + (row - row0)], rows [row0, row0 + stride) = this rank's); fixed order => run-to-run identical.  Cells of
// other ranks get 0 (several ranks: the all-reduce of the first replicated level sums the ranks' parts, setup()).
template<int D>
__global__ void __launch_bounds__(kThreads) k_mg_s0(uint64_t bound, const uint64_t* __restrict__ count, const uint64_t* __restrict__ cstart,
	const uint32_t* __restrict__ row_len, const double* __restrict__ row_s, uint64_t stride, uint64_t row0, double* __restrict__ S, double* __restrict__ dinv, double omega)
{
	constexpr int K = (D == 3) ? 27 : 9;
	const uint64_t t = static_cast<uint64_t>(blockIdx.x) * kThreads + threadIdx.x;
	const uint64_t cc = t / K;
	const int s = static_cast<int>(t % K);
	if (cc >= bound || cc >= *count) return;
	double sum = 0.0;
	for (uint64_t r = cstart[cc]; r < cstart[cc + 1]; r++)
		if (r >= row0 && r - row0 < stride && row_len[r]) sum += row_s[static_cast<uint64_t>(s) * stride + (r - row0)];
	S[cc * K + s] = sum;
	if (s == K / 2) dinv[cc] = (sum != 0.0) ? omega / sum : 0.0;
}

// Galerkin operator of level l + 1 from level l: every stencil entry of every child lands in the slot of the parent of the
// cell it points to
template<int D>
__global__ void __launch_bounds__(kThreads) k_mg_galerkin(uint64_t bound, const uint64_t* __restrict__ count_hi, Dims lo, const uint32_t* __restrict__ key_lo,
	const uint32_t* __restrict__ child, const uint32_t* __restrict__ nbr_lo, const double* __restrict__ S_lo, double* __restrict__ S_hi,
	double* __restrict__ dinv_hi, double omega)
{
	constexpr int K = (D == 3) ? 27 : 9;
	constexpr int CH = 1 << D;
	const uint64_t C = static_cast<uint64_t>(blockIdx.x) * kThreads + threadIdx.x;
	if (C >= bound || C >= *count_hi) return;
	double acc[K]; // indexed dynamically below: lives in (L1-cached) local memory, 3^D x 2^D updates per thread
	for (int s = 0; s < K; s++) acc[s] = 0.0;
	for (int q = 0; q < CH; q++)
	{
		const uint32_t c = child[C * CH + q];
		if (c == kMgNone) continue;
		long long cc[3]; decode<D>(key_lo[c], lo, cc);
		int s = 0;
		for (int ox = -1; ox <= 1; ox++)
			for (int oy = (D == 3 ? -1 : 0); oy <= (D == 3 ? 1 : 0); oy++)
				for (int oz = -1; oz <= 1; oz++, s++)
				{
					if (nbr_lo[static_cast<uint64_t>(c) * K + s] == kMgNone) continue;
					// offset of the target's parent relative to C (cc >> 1 == C's coordinates)
					const int tx = static_cast<int>(((cc[0] + ox) >> 1) - (cc[0] >> 1));
					const int ty = (D == 3) ? static_cast<int>(((cc[1] + oy) >> 1) - (cc[1] >> 1)) : 0;
					const int tz = static_cast<int>(((cc[D - 1] + oz) >> 1) - (cc[D - 1] >> 1));
					const int t = (D == 3) ? ((tx + 1) * 3 + (ty + 1)) * 3 + (tz + 1) : (tx + 1) * 3 + (tz + 1);
					acc[t] += S_lo[static_cast<uint64_t>(c) * K + s];
				}
	}
	for (int s = 0; s < K; s++) S_hi[C * K + s] = acc[s];
	dinv_hi[C] = (acc[K / 2] != 0.0) ? omega / acc[K / 2] : 0.0;
}

// several ranks: the damped inverse diagonal of a level whose operator has just been summed over the ranks
__global__ void __launch_bounds__(kThreads) k_mg_dinv(uint64_t bound, int K, const double* __restrict__ S, double* __restrict__ dinv, double omega)
{
	const uint64_t c = static_cast<uint64_t>(blockIdx.x) * kThreads + threadIdx.x;
	if (c >= bound) return;
	const double d = S[c * K + K / 2];
	dinv[c] = (d != 0.0) ? omega / d : 0.0;
}

#define MPS_TRY(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) return e_; } while (0)

Dims dims_of(const MgLevelBufs& l) { Dims d; for (int a = 0; a < 3; a++) d.n[a] = l.dims[a]; return d; }

template<int D>
cudaError_t setup(mps_solver* s)
{
	constexpr int K = (D == 3) ? 27 : 9;
	constexpr int CH = 1 << D;
	MgBuffers& mg = s->mg;
	cudaStream_t st = s->stream;
	const uint64_t n = s->n;
	uint64_t& L = s->stats.kernel_launches;

	// ---- occupancy and compact ids of the upper levels (level 0 was ranked by launch_mg_rank0 during the sort) ----
	for (int l = 0; l + 1 < mg.levels; l++)
	{
		MgLevelBufs& lo = mg.lv[l]; MgLevelBufs& hi = mg.lv[l + 1];
		k_mg_occ_up<D><<<blocks_for(hi.dense, kThreads), kThreads, 0, st>>>(hi.dense, dims_of(lo), dims_of(hi), lo.flag.p, hi.flag.p);
		L += 1;
		MPS_TRY(launch_exclusive_scan_u32_to_u64(hi.flag.p, hi.rank.p, hi.dense, s->scan_tmp, st, &L));
	}
	{
		MgBatch m{};
		m.levels = mg.levels;
		MgBlocks bc{}, bt{}, bh{};
		for (int l = 0; l < mg.levels; l++)
		{
			MgLevelBufs& lv = mg.lv[l];
			MgLevelDev& d = m.lv[l];
			d.dense = lv.dense; d.bound = lv.bound; d.dims = dims_of(lv);
			d.flag = lv.flag.p; d.rank = lv.rank.p; d.key = lv.key.p; d.nbr = lv.nbr.p; d.parent = lv.parent.p; d.child = (l > 0) ? lv.child.p : nullptr;
			bc.bb[l + 1] = bc.bb[l] + blocks_for(lv.dense, kThreads);
			bt.bb[l + 1] = bt.bb[l] + blocks_for(lv.bound, kThreads);
			if (l > 0) bh.bb[l] = bh.bb[l - 1] + blocks_for(lv.bound, kThreads);
		}
		k_mg_compact_all<<<bc.bb[mg.levels], kThreads, 0, st>>>(m, bc);
		k_mg_topology_all<D><<<bt.bb[mg.levels], kThreads, 0, st>>>(m, bt);
		L += 2;
		if (mg.levels > 1) { k_mg_children_all<D><<<bh.bb[mg.levels - 1], kThreads, 0, st>>>(m, bh); L += 1; }
	}
	// ---- rows <-> level-0 cells ----
	MgLevelBufs& l0 = mg.lv[0];
	k_mg_rows<<<blocks_for(n, kThreads), kThreads, 0, st>>>(n, static_cast<uint32_t>(s->env.ncells), s->skey.p, l0.rank.p, mg.crow.p);
	k_mg_cstart<<<blocks_for(l0.bound + 1, kThreads), kThreads, 0, st>>>(l0.bound, l0.rank.p + l0.dense, s->env.ncells, l0.key.p, s->cell_start.p, mg.cstart.p);
	L += 2;
	// ---- operators ----
	k_mg_s0<D><<<blocks_for(l0.bound * K, kThreads), kThreads, 0, st>>>(l0.bound, l0.rank.p + l0.dense, mg.cstart.p, s->row_len.p, mg.row_s.p, s->own1() - s->own0(), s->own0(), l0.S.p,
		l0.dinv.p, mg.omega);
	L += 1;
	for (int l = 0; l < mg.levels; l++)
	{
		MgLevelBufs& lo = mg.lv[l];
		if (s->comm.on && l == mg.k_dist)
		{
			// Several ranks: up to here every rank holds the operators of its own cells only — k_mg_s0 leaves zeros for the rows of
			// other ranks, and the Galerkin sums of zeros are zeros — so the sum over the ranks is the whole operator of level k
			// (a level-1 cell whose children sit on two ranks gets its two partial sums: a + b in either order is the same double).
			// Everything above is rebuilt from it, whole and identical on every rank.
			MPS_TRY(comm_allreduce_sum(s, lo.S.p, lo.bound * K));
			k_mg_dinv<<<blocks_for(lo.bound, kThreads), kThreads, 0, st>>>(lo.bound, K, lo.S.p, lo.dinv.p, mg.omega);
			L += 1;
		}
		if (l + 1 == mg.levels) break;
		MgLevelBufs& hi = mg.lv[l + 1];
		k_mg_galerkin<D><<<blocks_for(hi.bound, kThreads), kThreads, 0, st>>>(hi.bound, hi.rank.p + hi.dense, dims_of(lo), lo.key.p, hi.child.p, lo.nbr.p, lo.S.p,
			hi.S.p, hi.dinv.p, mg.omega);
		L += 1;
	}
	(void)CH;
	return cudaGetLastError();
}

} // namespace

// Level geometry from the grid extents (host only): level l + 1 halves every axis (rounding up) until one cell is left.
void mg_configure(mps_solver* s)
{
	MgBuffers& mg = s->mg;
	const int D = s->env.dim;
	mg.on = true;
	if (const char* v = std::getenv("MPS_CG_PRECOND")) mg.on = std::atoi(v) != 0;
	mg.omega = 0.8; mg.gamma = 1.8; mg.top_sweeps = 4; mg.top_cells = 64;
	if (const char* v = std::getenv("MPS_MG_OMEGA")) mg.omega = std::atof(v);
	if (const char* v = std::getenv("MPS_MG_GAMMA")) mg.gamma = std::atof(v);
	if (const char* v = std::getenv("MPS_MG_TOP_SWEEPS")) { const int k = std::atoi(v); if (k >= 0 && k <= 64) mg.top_sweeps = k; }
	if (const char* v = std::getenv("MPS_MG_TOP_CELLS")) { const int k = std::atoi(v); if (k >= 1) mg.top_cells = static_cast<uint32_t>(k); }
	if (const char* v = std::getenv("MPS_MG_SMALL_CELLS")) { const int k = std::atoi(v); if (k >= 0) mg.small_cells = static_cast<uint32_t>(k); }
	if (const char* v = std::getenv("MPS_MG_DIST_CELLS")) { const long long k = std::atoll(v); if (k >= 0) mg.dist_cells = static_cast<uint64_t>(k); }
	long long d[3] = { 1, 1, 1 };
	for (int a = 0; a < D; a++) d[a] = s->env.grid_n[a];
	int l = 0;
	for (; l < kMgMaxLevels; l++)
	{
		MgLevelBufs& lv = mg.lv[l];
		lv.dense = 1;
		for (int a = 0; a < 3; a++) { lv.dims[a] = d[a]; lv.dense *= static_cast<uint64_t>(d[a]); }
		bool one = true;
		for (int a = 0; a < D; a++) one = one && (d[a] == 1);
		// the V-cycle stops at the first level with <= top_cells occupied cells (k_pcg_stream); a level whose whole bounding grid is
		// that small is certainly the last one it can use: nothing above it needs to be built
		if (one || lv.dense <= mg.top_cells) { l++; break; }
		for (int a = 0; a < D; a++) d[a] = (d[a] + 1) / 2;
	}
	mg.levels = l;
}

// during the sort, right after the cell table: occupied cells of the neighbour grid -> compact ids; the count is read back
// with the neighbour-list size (the step's one host round trip) and sizes every level
cudaError_t launch_mg_rank0(mps_solver* s)
{
	MgBuffers& mg = s->mg;
	if (!mg.on || s->n == 0) return cudaSuccess;
	MgLevelBufs& l0 = mg.lv[0];
	cudaStream_t st = s->stream;
	MPS_TRY(l0.flag.ensure(l0.dense + 1, st)); MPS_TRY(l0.rank.ensure(l0.dense + 2, st));
	k_mg_occ0<<<blocks_for(l0.dense, kThreads), kThreads, 0, st>>>(l0.dense, s->cell_count.p, l0.flag.p);
	s->stats.kernel_launches += 1;
	return launch_exclusive_scan_u32_to_u64(l0.flag.p, l0.rank.p, l0.dense, s->scan_tmp, st, &s->stats.kernel_launches);
}

// buffers of every level for `cells0` occupied cells at level 0 (exact, from the read-back); upper levels are bounded by
// the level below and by their dense grid
cudaError_t mg_ensure(mps_solver* s, uint64_t cells0)
{
	MgBuffers& mg = s->mg;
	if (!mg.on) return cudaSuccess;
	const int K = (s->env.dim == 3) ? 27 : 9, CH = 1 << s->env.dim;
	cudaStream_t st = s->stream;
	mg.cells0 = cells0;
	mg.in_arena = s->comm.on;
	uint64_t bound = cells0;
	uint64_t off = 0;
	for (int l = 0; l < mg.levels; l++)
	{
		MgLevelBufs& lv = mg.lv[l];
		if (l > 0 && lv.dense < bound) bound = lv.dense;
		if (bound < 1) bound = 1;
		lv.bound = bound;
		MPS_TRY(lv.flag.ensure(lv.dense + 1, st)); MPS_TRY(lv.rank.ensure(lv.dense + 2, st));
		MPS_TRY(lv.key.ensure(bound, st)); MPS_TRY(lv.nbr.ensure(bound * K, st)); MPS_TRY(lv.parent.ensure(bound, st));
		if (l > 0) MPS_TRY(lv.child.ensure(bound * CH, st));
		MPS_TRY(lv.S.ensure(bound * K, st)); MPS_TRY(lv.dinv.ensure(bound, st));
		if (!mg.in_arena) { MPS_TRY(lv.r.ensure(bound, st)); MPS_TRY(lv.e0.ensure(bound, st)); MPS_TRY(lv.e1.ensure(bound, st)); }
		const uint64_t pad = (bound + 1) & ~1ull;
		for (int q = 0; q < 4; q++) { mg.vec_off[l][q] = off; off += pad; }
	}
	mg.vec_total = off;
	if (mg.in_arena)
	{
		// several ranks: the level vectors live in the peer arena, at the same offsets on every rank (the bounds follow from cells0,
		// which every rank computes from the same replicated state).  Level 0 is distributed when it holds more than dist_cells cells.
		MPS_TRY(comm_ensure_arena(s, s->n + 64, mg.vec_total));
		mg.k_dist = (cells0 > mg.dist_cells && mg.levels > 1) ? 1 : 0;
	}
	MPS_TRY(mg.crow.ensure(s->n + 64, st)); MPS_TRY(mg.cstart.ensure(cells0 + 2, st));
	MPS_TRY(mg.dinv0.ensure(s->n + 64, st));
	MPS_TRY(mg.row_s.ensure((s->own1() - s->own0()) * K, st)); // per-row stencil sums of the rows this rank assembles
	return cudaSuccess;
}

cudaError_t launch_mg_setup(mps_solver* s)
{
	if (!s->mg.on || s->n == 0) return cudaSuccess;
	return s->env.dim == 2 ? setup<2>(s) : setup<3>(s);
}

} // namespace mps
