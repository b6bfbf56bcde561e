// mps_grid.cu — the cell-linked neighbour grid, rebuilt every step as a counting sort of particles by cell key with
// cell start tables, and the per-step neighbour list.  Replaces Grid::Clear/Store/Iterator (Grid.hpp:222-559) and
// Computer::SearchNeighbor (Computer.hpp:698-756).
//
// Bit-exactness contract (north star: "cell IDs and neighbour sets must be bit-exact"):
//   cell index  = floor((x - MinX) / NeighborLength) with a true IEEE division        (Grid.hpp:89-92,250-254)
//   in-range    : sqrt(sum_k (x_i - x_j)_k^2) < NeighborLength, summed left to right   (Computer.hpp:568-572,744-745)
// This translation unit is compiled with -fmad=false so that no multiply-add is contracted; CUDA's double-precision
// '/' and sqrt() are IEEE round-to-nearest.  The list ORDER is also the reference's: stencil cells x-major ... z-minor
// (Grid.hpp:363-403), inside a cell ascending original id (the serial Store loop, Computer.hpp:705-717) — because slots
// are sorted by (cell key, original id), the three z-neighbour cells of one (x[,y]) column are one contiguous slot range.
//
// Kernels (all HBM/latency-bound integer work, N = particles, C = cells):
//   k_cell_key    N x (read pos 8D, type 1; write key 4, rank 4, atomics on cell_count)
//   scan          C x 12 B
//   k_scatter     N x 16 B ; k_rank_fix N x (4 + occupancy x 4) ; k_reorder N x 2 x (16D + 16 + 1 + 4 + 4)
//   k_search<false/true>  N x (pos of ~3^D cells from L1/L2) -> counts / list
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "mps_solver.h"

namespace mps {
namespace {

constexpr int kThreads = 256;

template<int D>
__device__ __forceinline__ unsigned long long linear_cell(const long long* c, const EnvConst& env)
{
	unsigned long long idx = 0;
#pragma unroll
	for (int a = 0; a < D; a++) idx = idx * static_cast<unsigned long long>(env.grid_n[a]) + static_cast<unsigned long long>(c[a]);
	return idx;
}

// Grid.hpp:250-254 + :284-290.  Returns false when the position is outside the grid.
template<int D>
__device__ __forceinline__ bool cell_of(const Vec<D>& x, const EnvConst& env, long long* c)
{
	bool in = true;
#pragma unroll
	for (int a = 0; a < D; a++)
	{
		const double f = floor((x.v[a] - env.min_x[a]) / env.neighbor_length);
		// evaluated in floating point so that NaN / huge values fall outside instead of hitting an undefined cast
		const bool ok = (f >= 0.0) && (f < static_cast<double>(env.grid_n[a]));
		c[a] = ok ? static_cast<long long>(f) : -1;
		in = in && ok;
	}
	return in;
}

template<int D>
__global__ void __launch_bounds__(kThreads) k_cell_key(uint64_t n, const Vec<D>* __restrict__ pos, uint8_t* __restrict__ type,
	uint32_t* __restrict__ key, uint32_t* __restrict__ rank, uint32_t* __restrict__ cell_count, EnvConst env, DevScalars* sc)
{
	const uint64_t s = static_cast<uint64_t>(blockIdx.x) * kThreads + threadIdx.x;
	if (s >= n) return;
	uint32_t k = static_cast<uint32_t>(env.ncells);
	if (type[s] != kDisabled)
	{
		long long c[3];
		if (cell_of<D>(pos[s], env, c))
		{
			k = static_cast<uint32_t>(linear_cell<D>(c, env));
		}
		else
		{
			type[s] = kDisabled; // Computer.hpp:711-715: a particle that cannot be stored is disabled
			atomicAdd(&sc->disabled_now, 1u);
		}
	}
	key[s] = k;
	const uint32_t r = atomicAdd(&cell_count[k], 1u);
	rank[s] = r;
	// Grid.hpp:311-318: storing into a full bucket throws "Too many particle in a block"
	if (k != static_cast<uint32_t>(env.ncells) && r >= env.cell_cap) atomicMax(&sc->error, static_cast<int>(MPS_CELL_OVERFLOW));
}

__global__ void __launch_bounds__(kThreads) k_scatter(uint64_t n, const uint32_t* __restrict__ key, const uint32_t* __restrict__ rank,
	const uint64_t* __restrict__ cell_start, const uint32_t* __restrict__ orig, uint32_t* __restrict__ perm, uint32_t* __restrict__ perm_orig)
{
	const uint64_t s = static_cast<uint64_t>(blockIdx.x) * kThreads + threadIdx.x;
	if (s >= n) return;
	const uint64_t dst = cell_start[key[s]] + rank[s];
	perm[dst] = static_cast<uint32_t>(s);
	perm_orig[dst] = orig[s];
}

// The atomic rank inside a cell is arbitrary; the reference order is ascending original id.  Each entry counts the
// members of its cell with a smaller original id (a cell holds at most cell_cap <= 16 / 64 particles).
__global__ void __launch_bounds__(kThreads) k_rank_fix(uint64_t n, uint32_t tail_key, const uint32_t* __restrict__ key,
	const uint64_t* __restrict__ cell_start, const uint32_t* __restrict__ perm, const uint32_t* __restrict__ perm_orig, uint32_t* __restrict__ perm2)
{
	const uint64_t d = static_cast<uint64_t>(blockIdx.x) * kThreads + threadIdx.x;
	if (d >= n) return;
	const uint32_t src = perm[d];
	const uint32_t k = key[src];
	if (k == tail_key) { perm2[d] = src; return; } // Disabled tail (key == ncells): order is irrelevant
	const uint64_t b = cell_start[k], e = cell_start[k + 1];
	const uint32_t mine = perm_orig[d];
	uint64_t smaller = 0;
	for (uint64_t m = b; m < e; m++) smaller += (perm_orig[m] < mine) ? 1u : 0u;
	perm2[b + smaller] = src;
}

template<int D>
__global__ void __launch_bounds__(kThreads) k_reorder(uint64_t n, const uint32_t* __restrict__ perm2, const uint32_t* __restrict__ key,
	Particles<D> src, Particles<D> dst, uint32_t* __restrict__ skey, uint32_t* __restrict__ inv)
{
	const uint64_t d = static_cast<uint64_t>(blockIdx.x) * kThreads + threadIdx.x;
	if (d >= n) return;
	const uint32_t s = perm2[d];
	dst.pos[d] = src.pos[s];
	dst.vel[d] = src.vel[s];
	dst.prs[d] = src.prs[s];
	dst.nden[d] = src.nden[s];
	dst.type[d] = src.type[s];
	const uint32_t o = src.orig[s];
	dst.orig[d] = o;
	skey[d] = key[s];
	inv[o] = static_cast<uint32_t>(d);
}

// Computer.hpp:720-755.  FILL = false counts, FILL = true writes the list at nbr_ptr[i].
template<int D, bool FILL>
__global__ void __launch_bounds__(kThreads) k_search(uint64_t first, uint64_t n, const Vec<D>* __restrict__ pos, const uint32_t* __restrict__ skey,
	const uint64_t* __restrict__ cell_start, uint32_t* __restrict__ nbr_cnt, const uint64_t* __restrict__ nbr_ptr,
	uint32_t* __restrict__ nbr, EnvConst env)
{
	const uint64_t i = first + static_cast<uint64_t>(blockIdx.x) * kThreads + threadIdx.x;
	if (i >= n) return;
	const uint32_t k = skey[i];
	if (k == static_cast<uint32_t>(env.ncells))
	{
		if (!FILL) nbr_cnt[i] = 0; // Disabled particles have no list (Computer.hpp:730)
		return;
	}
	// decode the cell
	long long c[3];
	{
		unsigned long long rest = k;
#pragma unroll
		for (int a = D - 1; a >= 0; a--)
		{
			c[a] = static_cast<long long>(rest % static_cast<unsigned long long>(env.grid_n[a]));
			rest /= static_cast<unsigned long long>(env.grid_n[a]);
		}
	}
	const Vec<D> xi = pos[i];
	const long long nz = env.grid_n[D - 1];
	const long long zlo = (c[D - 1] > 0) ? c[D - 1] - 1 : 0;
	const long long zhi = (c[D - 1] + 1 < nz) ? c[D - 1] + 1 : nz - 1;
	uint32_t cnt = 0;
	const uint64_t base = FILL ? nbr_ptr[i] : 0;
	const double nl2_lim = env.nl2_lim;

	for (int ox = -1; ox <= 1; ox++)
	{
		const long long cx = c[0] + ox;
		if (cx < 0 || cx >= env.grid_n[0]) continue;
		for (int oy = (D == 3 ? -1 : 0); oy <= (D == 3 ? 1 : 0); oy++)
		{
			long long b[3];
			b[0] = cx;
			if (D == 3)
			{
				b[1] = c[1] + oy;
				if (b[1] < 0 || b[1] >= env.grid_n[1]) continue;
			}
			b[D - 1] = zlo;
			const unsigned long long lin_lo = linear_cell<D>(b, env);
			const unsigned long long lin_hi = lin_lo + static_cast<unsigned long long>(zhi - zlo);
			const uint64_t jb = cell_start[lin_lo], je = cell_start[lin_hi + 1];
			// four candidates at a time: their positions are loaded side by side (independent requests), then tested in slot order
			for (uint64_t j0 = jb; j0 < je; j0 += 4)
			{
				Vec<D> xj[4];
#pragma unroll
				for (int u = 0; u < 4; u++) { const uint64_t j = j0 + u; xj[u] = pos[j < je ? j : je - 1]; }
#pragma unroll
				for (int u = 0; u < 4; u++)
				{
					const uint64_t j = j0 + u;
					double r2 = 0.0;
#pragma unroll
					for (int a = 0; a < D; a++)
					{
						const double d = xi.v[a] - xj[u].v[a];
						r2 += d * d; // -fmad=false: rounded product, then rounded sum, as uBLAS inner_prod
					}
					if ((j < je) && (j != i) && (r2 < nl2_lim)) // <=> sqrt(r2) < neighbor_length, bit for bit (EnvConst::nl2_lim)
					{
						if (FILL) nbr[base + cnt] = static_cast<uint32_t>(j);
						cnt++;
					}
				}
			}
		}
	}
	if (!FILL) nbr_cnt[i] = cnt;
}

// Multi-GPU: slab boundaries of this step.  The nominal split is equal shares of the modelled WORK of the slots that lie in the
// grid (a Fluid particle carries a full PPE row and every gather stage, a Wall particle a short row, a Dummy particle only shows
// up in other particles' lists: weights w_type, prefix sums over the sorted slots in `wprefix`); every boundary then moves up to
// the next boundary between cell COLUMNS (all cells of one x index) whose index is a multiple of 2^a, so that cells — and the
// 2^a-column blocks of the preconditioner's first a levels — are never shared between ranks.  a is the largest value <= a_max for
// which every rank still gets at least one block of columns; out = [R + 1 slots | R + 1 columns | a | ok].
// The Disabled tail (slots behind the grid) goes to the last rank.  One thread: R binary searches over the cell table.
__global__ void __launch_bounds__(kThreads) k_slot_weight(const uint64_t n, const uint8_t* __restrict__ type, const uint32_t w_fluid, const uint32_t w_wall,
	const uint32_t w_dummy, uint32_t* __restrict__ w)
{
	const uint64_t i = static_cast<uint64_t>(blockIdx.x) * kThreads + threadIdx.x;
	if (i >= n) return;
	const uint8_t t = type[i];
	w[i] = (t == kFluid) ? w_fluid : (t == kWall) ? w_wall : (t == kDummy) ? w_dummy : 0u;
}
__global__ void k_slab_bounds(const int R, const uint64_t n, const uint64_t ncells, const uint32_t ncols, const uint64_t colstride, const int a_max,
	const uint64_t* __restrict__ cell_start, const uint64_t* __restrict__ wprefix, unsigned long long* __restrict__ out)
{
	if (blockIdx.x != 0 || threadIdx.x != 0) return;
	const uint64_t in_grid = cell_start[ncells];
	const uint64_t total = wprefix[in_grid];
	int a = a_max;
	bool ok = false;
	for (; a >= 0 && !ok; a--)
	{
		const uint32_t unit = 1u << a, units = (ncols + unit - 1) / unit;
		ok = true;
		uint32_t prev = 0;
		for (int r = 0; r <= R; r++)
		{
			uint32_t col; uint64_t slot;
			if (r == 0) { col = 0; slot = 0; }
			else if (r == R) { col = units * unit; slot = n; }
			else
			{
				const uint64_t target = total / R * r + total % R * r / R;
				uint32_t lo = 0, hi = units; // smallest block of columns whose first slot has at least `target` work before it
				while (lo < hi)
				{
					const uint32_t mid = (lo + hi) >> 1;
					const uint64_t c = static_cast<uint64_t>(mid) * unit;
					if (wprefix[cell_start[(c < ncols ? c : ncols) * colstride]] < target) lo = mid + 1; else hi = mid;
				}
				// nearest boundary, not the next one: the block below may be closer to the target
				if (lo > 0 && lo > prev / unit + 1)
				{
					const uint64_t cb = static_cast<uint64_t>(lo - 1) * unit, ca = static_cast<uint64_t>(lo) * unit;
					const uint64_t wb = wprefix[cell_start[(cb < ncols ? cb : ncols) * colstride]], wa = wprefix[cell_start[(ca < ncols ? ca : ncols) * colstride]];
					if (target - wb < wa - target) lo -= 1;
				}
				col = lo * unit;
				slot = cell_start[(col < ncols ? col : ncols) * colstride];
			}
			if (r > 0 && col <= prev) ok = false;
			prev = col;
			out[r] = slot; out[R + 1 + r] = col;
		}
		if (ok) { out[2 * R + 2] = static_cast<unsigned long long>(a); break; }
	}
	out[2 * R + 3] = ok ? 1ull : 0ull;
	// halo of every slab: one cell column on either side (the neighbour stencil is 3 cells wide)
	for (int r = 0; r < R; r++)
	{
		const uint64_t c0 = out[R + 1 + r], c1 = out[R + 1 + r + 1];
		const uint64_t cl = (c0 > 0) ? c0 - 1 : 0, ch = (c1 + 1 < ncols) ? c1 + 1 : ncols;
		out[2 * R + 4 + r] = cell_start[(cl < ncols ? cl : ncols) * colstride];
		out[3 * R + 4 + r] = cell_start[(ch < ncols ? ch : ncols) * colstride];
	}
}

template<int D>
__global__ void __launch_bounds__(kThreads) k_get_cells(uint64_t n, const Vec<D>* __restrict__ pos, const uint32_t* __restrict__ orig,
	long long* __restrict__ out, EnvConst env)
{
	const uint64_t s = static_cast<uint64_t>(blockIdx.x) * kThreads + threadIdx.x;
	if (s >= n) return;
	const Vec<D> x = pos[s];
	const uint64_t o = orig[s];
#pragma unroll
	for (int a = 0; a < D; a++)
	{
		// same arithmetic as cell_of, but reported unclipped like Grid::Block (Grid.hpp:250-254)
		out[o * D + a] = static_cast<long long>(floor((x.v[a] - env.min_x[a]) / env.neighbor_length));
	}
}

template<int D>
Particles<D> view(mps_solver* s, int which)
{
	Particles<D> p;
	p.pos = reinterpret_cast<Vec<D>*>(s->pos[which].p);
	p.vel = reinterpret_cast<Vec<D>*>(s->vel[which].p);
	p.prs = s->prs[which].p;
	p.nden = s->nden[which].p;
	p.type = s->type[which].p;
	p.orig = s->orig[which].p;
	return p;
}

#define MPS_TRY(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) return e_; } while (0)

template<int D>
cudaError_t sort_and_search(mps_solver* s)
{
	const uint64_t n = s->n;
	cudaStream_t st = s->stream;
	const EnvConst& env = s->env;
	const unsigned nb = blocks_for(n, kThreads);
	if (n == 0) { s->nbr_total = 0; s->searched = true; return cudaSuccess; }

	MPS_TRY(s->key.ensure(n, st)); MPS_TRY(s->skey.ensure(n, st)); MPS_TRY(s->rank.ensure(n, st));
	MPS_TRY(s->perm.ensure(n, st)); MPS_TRY(s->perm_orig.ensure(n, st)); MPS_TRY(s->perm2.ensure(n, st));
	MPS_TRY(s->cell_count.ensure(env.ncells + 1, st)); MPS_TRY(s->cell_start.ensure(env.ncells + 2, st));
	MPS_TRY(s->nbr_cnt.ensure(n, st)); MPS_TRY(s->nbr_ptr.ensure(n + 1, st));
	const int nxt = s->cur ^ 1;
	const int vs = s->vec_stride();
	MPS_TRY(s->pos[nxt].ensure(n * vs, st)); MPS_TRY(s->vel[nxt].ensure(n * vs, st));
	MPS_TRY(s->prs[nxt].ensure(n, st)); MPS_TRY(s->nden[nxt].ensure(n, st));
	MPS_TRY(s->type[nxt].ensure(n, st)); MPS_TRY(s->orig[nxt].ensure(n, st));

	Particles<D> cur = view<D>(s, s->cur), next = view<D>(s, nxt);

	// 1. keys + per-cell histogram (the atomic's return value is a provisional rank inside the cell)
	MPS_TRY(cudaMemsetAsync(s->cell_count.p, 0, (env.ncells + 1) * sizeof(uint32_t), st));
	MPS_TRY(cudaMemsetAsync(&s->d_sc->disabled_now, 0, sizeof(unsigned int), st));
	k_cell_key<D><<<nb, kThreads, 0, st>>>(n, cur.pos, cur.type, s->key.p, s->rank.p, s->cell_count.p, env, s->d_sc);
	// 2. cell start table (entry ncells = start of the Disabled tail, entry ncells + 1 = n)
	MPS_TRY(launch_exclusive_scan_u32_to_u64(s->cell_count.p, s->cell_start.p, env.ncells + 1, s->scan_tmp, st, &s->stats.kernel_launches));
	// 2b. occupied cells -> compact ids: level 0 of the multigrid preconditioner's cell hierarchy (mps_mg.cu)
	MPS_TRY(launch_mg_rank0(s));
	// 3. scatter, deterministic order inside each cell, permute the state
	k_scatter<<<nb, kThreads, 0, st>>>(n, s->key.p, s->rank.p, s->cell_start.p, cur.orig, s->perm.p, s->perm_orig.p);
	k_rank_fix<<<nb, kThreads, 0, st>>>(n, static_cast<uint32_t>(env.ncells), s->key.p, s->cell_start.p, s->perm.p, s->perm_orig.p, s->perm2.p);
	k_reorder<D><<<nb, kThreads, 0, st>>>(n, s->perm2.p, s->key.p, cur, next, s->skey.p, s->inv.p);
	s->cur = nxt;
	s->stats.kernel_launches += 4;

	// 3b. multi-GPU: this step's slab boundaries, on cell columns, equal shares of the modelled work (one extra, tiny host round trip per step)
	if (s->comm.on)
	{
		const int R = s->comm.nranks;
		MPS_TRY(s->d_bounds.ensure(4ull * R + 4, st));
		const uint64_t colstride = env.ncells / static_cast<uint64_t>(env.grid_n[0]);
		const int a_max = 0; // cuts may fall between any two cell columns (the preconditioner's coarse levels are replicated, MgDist)
		// work per sorted slot by particle type (scratch: rank[] is free after the scatter, nbr_ptr[] is rebuilt by the search below)
		uint32_t w[3] = { 8, 4, 1 };
		if (const char* v = std::getenv("MPS_SLAB_WEIGHTS")) { unsigned a = 0, b = 0, c = 0; if (std::sscanf(v, "%u,%u,%u", &a, &b, &c) == 3 && a + b + c > 0) { w[0] = a; w[1] = b; w[2] = c; } }
		k_slot_weight<<<nb, kThreads, 0, st>>>(n, next.type, w[0], w[1], w[2], s->rank.p);
		MPS_TRY(launch_exclusive_scan_u32_to_u64(s->rank.p, s->nbr_ptr.p, n, s->scan_tmp, st, &s->stats.kernel_launches));
		k_slab_bounds<<<1, 32, 0, st>>>(R, n, env.ncells, static_cast<uint32_t>(env.grid_n[0]), colstride, a_max, s->cell_start.p, s->nbr_ptr.p, s->d_bounds.p);
		s->stats.kernel_launches += 2;
		std::vector<unsigned long long> hb(4ull * R + 4);
		MPS_TRY(cudaMemcpyAsync(hb.data(), s->d_bounds.p, hb.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
		MPS_TRY(cudaStreamSynchronize(st));
		if (!hb[2 * R + 3]) { s->comm_error = "fewer cell columns than GPUs: use fewer GPUs for this problem"; return cudaErrorUnknown; }
		s->own_b.assign(hb.begin(), hb.begin() + R + 1);
		s->col_b.resize(R + 1);
		for (int r = 0; r <= R; r++) s->col_b[r] = static_cast<uint32_t>(hb[R + 1 + r]);
		s->halo_lo.assign(hb.begin() + 2 * R + 4, hb.begin() + 3 * R + 4);
		s->halo_hi.assign(hb.begin() + 3 * R + 4, hb.begin() + 4 * R + 4);
		s->own_n = n;
	}
	// 4. neighbour list: count -> row pointers -> fill
	// lists are built for the rows this rank owns (all of them on one GPU)
	const Vec<D>* pos = next.pos;
	const uint64_t r0 = s->own0(), r1 = s->own1();
	const unsigned nbo = blocks_for(r1 - r0, kThreads);
	if (s->comm.on) MPS_TRY(cudaMemsetAsync(s->nbr_cnt.p, 0, n * sizeof(uint32_t), st));
	if (nbo) k_search<D, false><<<nbo, kThreads, 0, st>>>(r0, r1, pos, s->skey.p, s->cell_start.p, s->nbr_cnt.p, nullptr, nullptr, env);
	s->stats.kernel_launches += 1;
	MPS_TRY(launch_exclusive_scan_u32_to_u64(s->nbr_cnt.p, s->nbr_ptr.p, n, s->scan_tmp, st, &s->stats.kernel_launches));
	// the list length is needed on the host to size the buffer: the step's one host round trip.  The same read brings the
	// number of occupied cells (sizes the preconditioner's levels) and the device error flag: a step whose sort overflowed a
	// cell stops HERE, like the reference's Grid::Exception thrown from Store (Grid.hpp:311-318), before any later stage could
	// run on over-full cells.
	uint64_t total = 0, cells0 = 0;
	int dev_error = 0;
	MPS_TRY(cudaMemcpyAsync(&total, s->nbr_ptr.p + n, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
	if (s->mg.on) MPS_TRY(cudaMemcpyAsync(&cells0, s->mg.lv[0].rank.p + s->mg.lv[0].dense, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
	MPS_TRY(cudaMemcpyAsync(&dev_error, &s->d_sc->error, sizeof(int), cudaMemcpyDeviceToHost, st));
	MPS_TRY(cudaStreamSynchronize(st));
	s->sort_error = dev_error;
	if (dev_error == MPS_CELL_OVERFLOW) { s->nbr_total = 0; s->searched = false; return cudaSuccess; }
	MPS_TRY(mg_ensure(s, cells0));
	MPS_TRY(s->nbr.ensure(total + 1, st));
	if (nbo) k_search<D, true><<<nbo, kThreads, 0, st>>>(r0, r1, pos, s->skey.p, s->cell_start.p, nullptr, s->nbr_ptr.p, s->nbr.p, env);
	s->stats.kernel_launches += 1;
	s->nbr_total = total;
	s->searched = true;
	return cudaGetLastError();
}

} // namespace

cudaError_t launch_sort_and_search(mps_solver* s)
{
	return s->env.dim == 2 ? sort_and_search<2>(s) : sort_and_search<3>(s);
}

cudaError_t launch_get_cells(mps_solver* s, long long* d_cells)
{
	const unsigned nb = blocks_for(s->n, kThreads);
	if (s->n == 0) return cudaSuccess;
	if (s->env.dim == 2)
		k_get_cells<2><<<nb, kThreads, 0, s->stream>>>(s->n, reinterpret_cast<Vec<2>*>(s->pos[s->cur].p), s->orig[s->cur].p, d_cells, s->env);
	else
		k_get_cells<3><<<nb, kThreads, 0, s->stream>>>(s->n, reinterpret_cast<Vec<3>*>(s->pos[s->cur].p), s->orig[s->cur].p, d_cells, s->env);
	s->stats.kernel_launches += 1;
	return cudaGetLastError();
}

} // namespace mps
