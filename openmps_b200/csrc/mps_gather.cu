// mps_gather.cu — the neighbour-gather stages of the MPS step and the streaming updates between them.
//
// One skeleton, accumulate<D>(), mirrors Computer::AccumulateNeighbor<FIELDS...> (Computer.hpp:618-695) including the
// SPP free-surface virtual particle (Computer.hpp:636-693); every stage is a functor on top of it:
//   density            Computer.hpp:780-833        ecs / Dn/Dt        Computer.hpp:838-910
//   explicit forces    Computer.hpp:914-1021       PPE rows + rhs     Computer.hpp:1145-1328
//   pressure gradient  Computer.hpp:1433-1564      dynamic stabiliser Computer.hpp:1568-1656
//
// Parity: this translation unit is compiled with -fmad=false and every expression is written in the reference's
// evaluation order (uBLAS evaluates element-wise, left to right), one thread per particle walking its neighbour list
// in the reference's order.  With identical inputs the stage outputs are therefore expected to be bit-identical to
// the reference built with -ffp-contract=off; the tests allow a few ulp.
//
// Roofline: these kernels read ~8D+9 B per particle of compulsory data but evaluate a sqrt and a division per
// in-range pair in FP64, so they are bounded by the FP64 pipe, not by HBM (DESIGN.md "gather kernels").
#include <cfloat>

#include "mps_solver.h"

namespace mps {
namespace {

constexpr int kThreads = 128;

template<int D> struct Acc { double v[D]; };

template<int D> __device__ __forceinline__ Acc<D> azero() { Acc<D> r; for (int k = 0; k < D; k++) r.v[k] = 0.0; return r; }
__device__ __forceinline__ void add_to(double& s, const double c) { s += c; }
template<int D> __device__ __forceinline__ void add_to(Acc<D>& s, const Acc<D>& c) { for (int k = 0; k < D; k++) s.v[k] += c.v[k]; }

// Particle::W, Particle.hpp:71-75
__device__ __forceinline__ double weight(const double r, const double r_e) { return ((0 < r) && (r < r_e)) ? (r_e / r - 1) : 0; }

template<int D> __device__ __forceinline__ double inner(const Acc<D>& a, const Acc<D>& b)
{
	double t = 0.0;
#pragma unroll
	for (int k = 0; k < D; k++) t += a.v[k] * b.v[k];
	return t;
}
template<int D> __device__ __forceinline__ Acc<D> sub(const Vec<D>& a, const Vec<D>& b)
{
	Acc<D> r;
#pragma unroll
	for (int k = 0; k < D; k++) r.v[k] = a.v[k] - b.v[k];
	return r;
}
template<int D> __device__ __forceinline__ Acc<D> scale(const double s, const Acc<D>& a)
{
	Acc<D> r;
#pragma unroll
	for (int k = 0; k < D; k++) r.v[k] = s * a.v[k];
	return r;
}

struct Lists
{
	const uint64_t* ptr;
	const uint32_t* idx;
};

// Walks the neighbour list of slot i in list order and calls visit(j, x_j, dx = x_j - x_i, r2 = dx.dx) for the entries with
// r2 < r_e2.  kBatch entries at a time: the indices of a batch, then their positions, are loaded side by side — independent
// L2 round trips instead of a chain of two per entry (ncu, profiles/r02a_ncu_full_gather_kernels_3d1m.txt: these kernels stall
// on exactly that chain) — before the batch is visited in order.  Past the end of the list the last entry is loaded again
// and not visited.
constexpr int kBatch = 4;
template<int D, typename VISIT>
__device__ __forceinline__ void walk_in_range(const uint64_t i, const Vec<D>& xi, const Particles<D>& P, const Lists& L, const double r_e2, VISIT visit)
{
	const uint64_t eb = L.ptr[i], ee = L.ptr[i + 1];
	for (uint64_t e0 = eb; e0 < ee; e0 += kBatch)
	{
		uint32_t j[kBatch];
		Vec<D> xj[kBatch];
#pragma unroll
		for (int u = 0; u < kBatch; u++) { const uint64_t e = e0 + u; j[u] = L.idx[e < ee ? e : ee - 1]; } // the list never contains i itself nor Disabled particles
#pragma unroll
		for (int u = 0; u < kBatch; u++) xj[u] = P.pos[j[u]];
#pragma unroll
		for (int u = 0; u < kBatch; u++)
		{
			const Acc<D> dx = sub<D>(xj[u], xi);
			const double r2 = inner<D>(dx, dx);
			if ((e0 + u < ee) && (r2 < r_e2)) visit(j[u], xj[u], dx, r2);
		}
	}
}

// Computer.hpp:618-695.  func(j, x_j, u_j, p_j, type_j, dx = x_j - x_i, r = |dx|) -> contribution; j = -1 denotes the SPP
// virtual particle (type Fluid, position x_spp, velocity u_i, pressure 0).  Fields a stage does not use are not loaded
// (NEED_* flags).  r is evaluated once per pair (the reference's R(), its W() and its norm_2(dx) all round the same
// sqrt(dx.dx)); the weighted centroid dx_g only feeds the SPP branch, so it is only accumulated for particles below n0.
template<int D, bool NEED_U, bool NEED_P, bool NEED_T, typename SUM, typename FUNC>
__device__ __forceinline__ SUM accumulate(const uint64_t i, const Particles<D>& P, const Lists& L, const double* __restrict__ nws,
	const EnvConst& env, SUM sum, FUNC func)
{
	const double r_e = env.r_e;
	const double n0 = env.n0;
	const double this_n = nws[i];
	const bool spp = this_n < n0;
	Acc<D> dx_g = azero<D>();
	const Vec<D> xi = P.pos[i];
	walk_in_range<D>(i, xi, P, L, env.r_e2, [&](const uint32_t j, const Vec<D>& xj, const Acc<D>& dx, const double r2)
		{
			const Vec<D> uj = NEED_U ? P.vel[j] : vzero<D>();
			const double pj = NEED_P ? P.prs[j] : 0.0;
			const uint8_t tj = NEED_T ? P.type[j] : static_cast<uint8_t>(kFluid);
			const double r = sqrt(r2);
			add_to(sum, func(static_cast<long long>(j), xj, uj, pj, tj, dx, r));
			if (spp) add_to(dx_g, scale<D>(weight(r, r_e), dx));
		});
	// SPP virtual particle, Computer.hpp:664-692
	if (spp)
	{
#pragma unroll
		for (int k = 0; k < D; k++) dx_g.v[k] = dx_g.v[k] / n0;
		const double r_g = sqrt(inner<D>(dx_g, dx_g));
		if (r_g > DBL_EPSILON)
		{
			const double w_spp = n0 - this_n;
			const double r_spp = r_e / (w_spp + 1);
			const double f = r_spp / r_g;
			Vec<D> x_spp = vzero<D>();
#pragma unroll
			for (int k = 0; k < D; k++) x_spp.v[k] = xi.v[k] - f * dx_g.v[k];
			const Acc<D> dxs = sub<D>(x_spp, xi);
			add_to(sum, func(-1LL, x_spp, P.vel[i], 0.0, static_cast<uint8_t>(kFluid), dxs, sqrt(inner<D>(dxs, dxs))));
		}
	}
	return sum;
}

// Computer::R, Computer.hpp:568-572
template<int D> __device__ __forceinline__ double dist(const Vec<D>& a, const Vec<D>& b)
{
	const Acc<D> r = sub<D>(a, b);
	return sqrt(inner<D>(r, r));
}

// ---------------------------------------------------------------------------------------------------------------------
// Computer::ComputeNeighborDensities, Computer.hpp:780-833.  COUNT_ROWS additionally produces the PPE row lengths
// (in-range, non-Dummy neighbours + the diagonal) for the second call of the step, and saves originalX (SaveX,
// Computer.hpp:1025-1039) — both read exactly the data this pass has in registers anyway.
template<int D, bool COUNT_ROWS>
__global__ void __launch_bounds__(kThreads) k_density(uint64_t first, uint64_t n, Particles<D> P, Lists L, double* __restrict__ nws,
	uint32_t* __restrict__ row_len, Vec<D>* __restrict__ x0, EnvConst env)
{
	const uint64_t i = first + static_cast<uint64_t>(blockIdx.x) * kThreads + threadIdx.x;
	if (i >= n) return;
	const uint8_t t = P.type[i];
	const double n0 = env.n0;
	if ((t != kDummy) && (t != kDisabled))
	{
		// nWithoutSpp[i] = n0 while accumulating => the SPP branch is inactive (Computer.hpp:800); done inline here
		const double r_e = env.r_e;
		const Vec<D> xi = P.pos[i];
		double sum = 0.0;
		uint32_t len = 1;
		walk_in_range<D>(i, xi, P, L, env.r_e2, [&](const uint32_t j, const Vec<D>&, const Acc<D>&, const double r2)
			{
				sum += weight(sqrt(r2), r_e);
				if (COUNT_ROWS) len += (P.type[j] != kDummy) ? 1u : 0u;
			});
		nws[i] = sum;
		P.nden[i] = (sum < n0) ? n0 : sum; // std::max(thisN, n0)
		if (COUNT_ROWS) { row_len[i] = len; x0[i] = xi; }
	}
	else
	{
		nws[i] = n0;
		if (COUNT_ROWS) row_len[i] = 0;
	}
}

// Computer::NeighborDensityVariationSpeed, Computer.hpp:838-872 (MPS_HS)
template<int D>
__device__ __forceinline__ double dndt(const uint64_t i, const Particles<D>& P, const Lists& L, const double* __restrict__ nws, const EnvConst& env)
{
	if (nws[i] < env.n0) return 0.0;
	const Vec<D> ui = P.vel[i];
	const double s = accumulate<D, true, false, false>(i, P, L, nws, env, 0.0,
		[&](long long, const Vec<D>&, const Vec<D>& u, double, uint8_t, const Acc<D>& dx, const double r) -> double
		{
			const Acc<D> duv = sub<D>(u, ui);
			return inner<D>(dx, duv) / (r * r * r); // r = norm_2(dx): |dx_k| * |dx_k| == dx_k * dx_k
		});
	return -env.r_e * s;
}

// Computer::ComputeErrorCorrection, Computer.hpp:877-910
template<int D>
__global__ void __launch_bounds__(kThreads) k_ecs(uint64_t first, uint64_t n, Particles<D> P, Lists L, const double* __restrict__ nws,
	double* __restrict__ ecs, EnvConst env)
{
	const uint64_t i = first + static_cast<uint64_t>(blockIdx.x) * kThreads + threadIdx.x;
	if (i >= n) return;
	const uint8_t t = P.type[i];
	if ((t != kDummy) && (t != kDisabled))
	{
		const double n0 = env.n0;
		const double this_n = P.nden[i];
		const double speed = dndt<D>(i, P, L, nws, env);
		const double error = (this_n - n0) / n0;
		ecs[i] = fabs(error) * speed + fabs(speed) * error;
	}
}

template<int D>
__global__ void __launch_bounds__(1) k_dndt_one(uint64_t orig_id, const uint32_t* __restrict__ inv, Particles<D> P, Lists L,
	const double* __restrict__ nws, double* __restrict__ out, EnvConst env)
{
	out[0] = dndt<D>(inv[orig_id], P, L, nws, env);
}

// Computer::ComputeExplicitForces, first loop, Computer.hpp:937-990: acceleration of fluid particles into `a` (= du)
template<int D>
__global__ void __launch_bounds__(kThreads) k_explicit_accel(uint64_t first, uint64_t n, Particles<D> P, Lists L, const double* __restrict__ nws,
	Vec<D>* __restrict__ a, EnvConst env)
{
	const uint64_t i = first + static_cast<uint64_t>(blockIdx.x) * kThreads + threadIdx.x;
	if (i >= n) return;
	if (P.type[i] != kFluid) return;
	const Vec<D> xi = P.pos[i], ui = P.vel[i];
	Acc<D> vis = accumulate<D, true, false, true>(i, P, L, nws, env, azero<D>(),
		[&](long long, const Vec<D>&, const Vec<D>& u, double, uint8_t type, const Acc<D>&, const double r) -> Acc<D>
		{
			if (type != kDummy) return scale<D>(env.visc_coef / (r * r * r), sub<D>(u, ui)); // r = R(x_i, x_j)
			return azero<D>();
		});
	if (env.central_gravity)
	{
		// Computer.hpp:981-985
		const double g = env.g[D - 1];
		double r2 = 0.0;
#pragma unroll
		for (int k = 0; k < D; k++) { const double av = fabs(xi.v[k]); r2 += av * av; }
		const double r = sqrt(r2);
		const bool centre = (r < env.l0 * 0.01);
#pragma unroll
		for (int k = 0; k < D; k++) vis.v[k] += g * (centre ? 0.0 : (xi.v[k] / r));
	}
	else
	{
#pragma unroll
		for (int k = 0; k < D; k++) vis.v[k] += env.g[k];
	}
	Vec<D> out = vzero<D>();
#pragma unroll
	for (int k = 0; k < D; k++) out.v[k] = vis.v[k];
	a[i] = out;
}

// Computer::ComputeExplicitForces, second loop, Computer.hpp:996-1020
template<int D>
__global__ void __launch_bounds__(kThreads) k_explicit_move(uint64_t first, uint64_t n, Particles<D> P, const Vec<D>* __restrict__ a,
	const Vec<D>* __restrict__ wall /* original order */, const uint8_t* __restrict__ wall_group, const WallMotions wm, const DevScalars* __restrict__ sc)
{
	const uint64_t i = first + static_cast<uint64_t>(blockIdx.x) * kThreads + threadIdx.x;
	if (i >= n) return;
	const double dt = sc->dt;
	Vec<D> x = P.pos[i], u = P.vel[i];
	if (P.type[i] == kFluid)
	{
		const Vec<D> ai = a[i];
#pragma unroll
		for (int k = 0; k < D; k++) { u.v[k] += ai.v[k] * dt; x.v[k] += u.v[k] * dt; }
	}
	else
	{
		// Wall, Dummy and Disabled particles follow positionWall (Computer.hpp:1011-1019)
		const uint32_t o = P.orig[i];
		Vec<D> xw = wall[o];
		const int grp = wm.count ? wall_group[o] : 0;
		if (grp)
		{
			// positionWall(i, t, dt) for an analytic motion, t = Environment::T() after SetNextT (Computer.hpp:921,1703-1706)
			const WallMotion& m = wm.m[grp - 1];
			const double tau = fmin(fmax(sc->t - m.t0, 0.0), m.t1 - m.t0);
			const double sw = sin(m.omega * tau + m.phase) - sin(m.phase);
#pragma unroll
			for (int k = 0; k < D; k++) xw.v[k] = xw.v[k] + (m.vel[k] * tau + m.amp[k] * sw);
		}
#pragma unroll
		for (int k = 0; k < D; k++) { u.v[k] = (xw.v[k] - x.v[k]) / dt; x.v[k] = xw.v[k]; }
	}
	P.pos[i] = x;
	P.vel[i] = u;
}

// Computer::SaveX, Computer.hpp:1025-1039 (stand-alone form for the stage-level API; the step fuses it into k_density)
template<int D>
__global__ void __launch_bounds__(kThreads) k_save_x(uint64_t n, Particles<D> P, Vec<D>* __restrict__ x0)
{
	const uint64_t i = static_cast<uint64_t>(blockIdx.x) * kThreads + threadIdx.x;
	if (i >= n) return;
	const uint8_t t = P.type[i];
	if ((t != kDummy) && (t != kDisabled)) x0[i] = P.pos[i];
}

// PPE row lengths for the stage-level API (the step gets them from k_density<COUNT_ROWS>)
template<int D>
__global__ void __launch_bounds__(kThreads) k_row_len(uint64_t first, uint64_t n, Particles<D> P, Lists L, uint32_t* __restrict__ row_len, EnvConst env)
{
	const uint64_t i = first + static_cast<uint64_t>(blockIdx.x) * kThreads + threadIdx.x;
	if (i >= n) return;
	const uint8_t t = P.type[i];
	uint32_t len = 0;
	if ((t != kDummy) && (t != kDisabled))
	{
		len = 1;
		const Vec<D> xi = P.pos[i];
		for (uint64_t e = L.ptr[i]; e < L.ptr[i + 1]; e++)
		{
			const uint32_t j = L.idx[e];
			const Acc<D> dx = sub<D>(P.pos[j], xi);
			if (inner<D>(dx, dx) < env.r_e2) len += (P.type[j] != kDummy) ? 1u : 0u;
		}
	}
	row_len[i] = len;
}

// Computer::SetPressurePoissonEquation, Computer.hpp:1145-1328, written straight into device memory:
// row i = [off-diagonals in neighbour-list order ..., diagonal].  Inactive rows (Dummy / Disabled) have length 0 on the
// device (the reference's identity rows never couple to anything: b = x = 0, Computer.hpp:1195-1202).
// CHUNKED = false: plain CSR (row_ptr, col, val).  CHUNKED = true: the chunk-blob form the streaming CG kernel reads
// (mps_device.cuh): values + 16-bit window-local columns at the row's 16-bit offset inside its chunk's blob.
struct PpeOut
{
	const uint64_t* row_ptr; uint32_t* col; double* val;                                  // CSR
	const ChunkDesc* desc; const uint32_t* chunk_of_row; unsigned char* blobs;           // chunk blobs
};

// PRE = true additionally leaves behind what the multigrid preconditioner is built from (mps_mg.cu): 1 / a_ii, and the row's
// entries summed per cell of the 3^D stencil around the row's own cell (the cell of a slot is the one the sort put it in, so
// every neighbour of the list lies in that stencil): row_s[s * stride + (i - row0)] for the rows [row0, row0 + stride) this rank owns.
struct PpePre
{
	const uint32_t* skey; double* row_s; double* dinv0; uint64_t stride, row0;
};

template<int D, bool CHUNKED, bool PRE>
__global__ void __launch_bounds__(kThreads) k_ppe_fill(uint64_t first, uint64_t n, Particles<D> P, Lists L, const double* __restrict__ nws,
	const double* __restrict__ ecs, PpeOut out, PpePre pre, double* __restrict__ b, double* __restrict__ x, EnvConst env, DevScalars* sc)
{
	constexpr int K = (D == 3) ? 27 : 9;
	__shared__ double sacc[PRE ? K : 1][kThreads]; // [slot][thread]: conflict-free, dynamically indexed by slot
	const uint64_t i = first + static_cast<uint64_t>(blockIdx.x) * kThreads + threadIdx.x;
	if (i >= n) return;
	const uint8_t t = P.type[i];
	if ((t == kDummy) || (t == kDisabled))
	{
		b[i] = 0; x[i] = 0;
		if (PRE) pre.dinv0[i] = 0;
		return;
	}
	// slot of cell(j) in the stencil of cell(i): keys are x-major ... z-minor, so key_j - key_i + (all offsets + 1) is the
	// mixed-radix number (dx + 1, [dy + 1,] dz + 1) in radices (., [ny,] nz); every axis has >= 3 cells (Grid.hpp:140-150: +2 spare)
	const int nzc = static_cast<int>(env.grid_n[D - 1]), nyc = (D == 3) ? static_cast<int>(env.grid_n[1]) : 1;
	const long long key_i = PRE ? static_cast<long long>(pre.skey[i]) : 0;
	const int key_shift = (D == 3) ? (nyc * nzc + nzc + 1) : (nzc + 1);
	if (PRE)
	{
#pragma unroll
		for (int s = 0; s < K; s++) sacc[s][threadIdx.x] = 0.0;
	}
	atomicAdd(&sc->active_rows, 1ull); // same address for the whole warp: aggregated by the compiler into one atomic
	const double dt = sc->dt;
	const double n0 = env.n0;
	// where this row's entries go
	uint32_t* col = nullptr; double* val = nullptr; uint16_t* lcol = nullptr;
	uint64_t w = 0;
	constexpr int NR = (D == 3) ? kMaxRanges : 3; // window ranges of a chunk: one per x[,y] column offset of the stencil (mps_chunk.cu)
	uint32_t rstart[NR], rlen[NR], roff[NR];
	if (CHUNKED)
	{
		const ChunkDesc* d = out.desc + out.chunk_of_row[i];
		const uint32_t nnz_pad = round_up8(d->nnz);
		unsigned char* blob = out.blobs + d->blob_off + kBlobHeader;
		val = reinterpret_cast<double*>(blob);
		lcol = reinterpret_cast<uint16_t*>(blob + static_cast<uint64_t>(nnz_pad) * 8u);
		const uint16_t* rowoff = lcol + nnz_pad;
		w = rowoff[static_cast<uint32_t>(i) - d->row_begin];
#pragma unroll
		for (int k = 0; k < NR; k++) { rstart[k] = d->range_start[k]; rlen[k] = d->range_len[k]; roff[k] = d->range_off[k]; }
	}
	else
	{
		col = out.col; val = out.val; w = out.row_ptr[i];
	}
	auto put = [&](const uint32_t j, const double a)
	{
		if (CHUNKED)
		{
			uint32_t l = 0xffffu;
#pragma unroll
			for (int k = NR - 1; k >= 0; k--) { const uint32_t o = j - rstart[k]; if (o < rlen[k]) l = roff[k] + o; }
			lcol[w] = static_cast<uint16_t>(l);
		}
		else col[w] = j;
		val[w] = a;
		w++;
		if (PRE)
		{
			const int m = static_cast<int>(static_cast<long long>(pre.skey[j]) - key_i + key_shift);
			const int dz1 = m % nzc, rest = m / nzc;
			const int slot = (D == 3) ? ((rest / nyc) * 3 + (rest % nyc)) * 3 + dz1 : rest * 3 + dz1;
			sacc[slot][threadIdx.x] += a;
		}
	};

	// ONE walk of the list for the right-hand side (Computer.hpp:1204-1214: Dn/Dt, Computer.hpp:838-872) and the matrix row
	// (Computer.hpp:1246-1327): both are sums over the same in-range pairs, each kept in its own reference order.  Dn/Dt is zero
	// for particles below n0 (Computer.hpp:845) and the SPP virtual particle (Computer.hpp:664-692) exists only for those, so
	// a particle needs either the velocity differences or the weighted centroid, never both.
	const Vec<D> xi = P.pos[i], ui = P.vel[i];
	const double r_e = env.r_e;
	const double this_n = nws[i];
	const bool spp = this_n < n0;
	double s_dn = 0.0, a_ii = 0.0;
	Acc<D> dx_g = azero<D>();
	walk_in_range<D>(i, xi, P, L, env.r_e2, [&](const uint32_t j, const Vec<D>&, const Acc<D>& dx, const double r2)
		{
			const uint8_t tj = P.type[j];
			const double r = sqrt(r2);
			const double r3 = r * r * r;
			if (!spp)
			{
				const Acc<D> duv = sub<D>(P.vel[j], ui);
				s_dn += inner<D>(dx, duv) / r3;
			}
			if (tj != kDummy)
			{
				const double a_ij = env.ppe_coef / r3;
				put(j, a_ij);
				a_ii += -a_ij;
			}
			if (spp) add_to(dx_g, scale<D>(weight(r, r_e), dx));
		});
	if (spp)
	{
#pragma unroll
		for (int k = 0; k < D; k++) dx_g.v[k] = dx_g.v[k] / n0;
		const double r_g = sqrt(inner<D>(dx_g, dx_g));
		if (r_g > DBL_EPSILON)
		{
			const double w_spp = n0 - this_n;
			const double r_spp = r_e / (w_spp + 1);
			const double f = r_spp / r_g;
			Vec<D> x_spp = vzero<D>();
#pragma unroll
			for (int k = 0; k < D; k++) x_spp.v[k] = xi.v[k] - f * dx_g.v[k];
			const double r = dist<D>(xi, x_spp);
			a_ii += -(env.ppe_coef / (r * r * r)); // the virtual particle is a Fluid neighbour without a column
		}
	}
	const double speed = spp ? 0.0 : -r_e * s_dn;
	b[i] = -env.rho / (n0 * dt) * (speed + ecs[i]);
	x[i] = P.prs[i];
	put(static_cast<uint32_t>(i), a_ii);
	if (PRE)
	{
		pre.dinv0[i] = (a_ii != 0) ? 1.0 / a_ii : 0.0;
#pragma unroll
		for (int s = 0; s < K; s++) pre.row_s[static_cast<uint64_t>(s) * pre.stride + (i - pre.row0)] = sacc[s][threadIdx.x];
	}
}

// multi-rank only: right-hand side 0 and initial guess x = P (Computer.hpp:1195-1220) for EVERY row; the owners then
// overwrite their rows in k_ppe_fill.  Rows of other ranks are only ever read as columns.
template<int D>
__global__ void __launch_bounds__(kThreads) k_ppe_guess(uint64_t n, Particles<D> P, double* __restrict__ b, double* __restrict__ x)
{
	const uint64_t i = static_cast<uint64_t>(blockIdx.x) * kThreads + threadIdx.x;
	if (i >= n) return;
	const uint8_t t = P.type[i];
	b[i] = 0;
	x[i] = ((t == kDummy) || (t == kDisabled)) ? 0.0 : P.prs[i];
}

// pressure write-back, Computer.hpp:1076-1097
template<int D>
__global__ void __launch_bounds__(kThreads) k_assign_pressure(uint64_t first, uint64_t n, Particles<D> P, const double* __restrict__ x)
{
	const uint64_t i = first + static_cast<uint64_t>(blockIdx.x) * kThreads + threadIdx.x;
	if (i >= n) return;
	const uint8_t t = P.type[i];
	if ((t != kDummy) && (t != kDisabled))
	{
		const double p = x[i];
		P.prs[i] = (p < 0) ? 0 : p;
	}
}

// Computer::ModifyByPressureGradient, first loop (midpoint form), Computer.hpp:1447-1541
template<int D>
__global__ void __launch_bounds__(kThreads) k_gradient(uint64_t first, uint64_t n, Particles<D> P, Lists L, const double* __restrict__ nws,
	Vec<D>* __restrict__ du, EnvConst env, const DevScalars* __restrict__ sc)
{
	const uint64_t i = first + static_cast<uint64_t>(blockIdx.x) * kThreads + threadIdx.x;
	if (i >= n) return;
	if (P.type[i] != kFluid) return;
	const double dt = sc->dt;
	const double r_e = env.r_e;
	const double pi = P.prs[i];
	const Vec<D> xi = P.pos[i];
	const Acc<D> s = accumulate<D, false, true, true>(i, P, L, nws, env, azero<D>(),
		[&](long long, const Vec<D>&, const Vec<D>&, double p, uint8_t type, const Acc<D>& dx, const double r) -> Acc<D>
		{
			if (type != kDummy)
			{
				const double r2 = inner<D>(dx, dx);
				return scale<D>((p + pi) / r2 * weight(r, r_e), dx); // r = sqrt(r2)
			}
			return azero<D>();
		});
	const double coef = -dt / env.rho * static_cast<double>(D) / env.n0;
	Vec<D> out = vzero<D>();
#pragma unroll
	for (int k = 0; k < D; k++) out.v[k] = coef * s.v[k];
	du[i] = out;
}

// second loops of ModifyByPressureGradient (Computer.hpp:1544-1563) and DynamicStabilize (Computer.hpp:1638-1655)
template<int D>
__global__ void __launch_bounds__(kThreads) k_apply_du(uint64_t first, uint64_t n, Particles<D> P, const Vec<D>* __restrict__ du, const DevScalars* __restrict__ sc)
{
	const uint64_t i = first + static_cast<uint64_t>(blockIdx.x) * kThreads + threadIdx.x;
	if (i >= n) return;
	if (P.type[i] != kFluid) return;
	const double dt = sc->dt;
	const Vec<D> d = du[i];
	Vec<D> x = P.pos[i], u = P.vel[i];
#pragma unroll
	for (int k = 0; k < D; k++) { u.v[k] += d.v[k]; x.v[k] += d.v[k] * dt; }
	P.pos[i] = x;
	P.vel[i] = u;
}

// Computer::DynamicStabilize, first loop, Computer.hpp:1576-1635
template<int D>
__global__ void __launch_bounds__(kThreads) k_ds(uint64_t first, uint64_t n, Particles<D> P, Lists L, const double* __restrict__ nws,
	const Vec<D>* __restrict__ x0, Vec<D>* __restrict__ du, EnvConst env, const DevScalars* __restrict__ sc)
{
	const uint64_t i = first + static_cast<uint64_t>(blockIdx.x) * kThreads + threadIdx.x;
	if (i >= n) return;
	if (P.type[i] != kFluid) return;
	const double dt = sc->dt;
	const double d2 = env.ds_d2;
	const Vec<D> xi = P.pos[i];
	const Vec<D> x0i = x0[i];
	const Acc<D> s = accumulate<D, false, false, true>(i, P, L, nws, env, azero<D>(),
		[&](long long j, const Vec<D>&, const Vec<D>&, double, uint8_t type, const Acc<D>& dx, const double) -> Acc<D>
		{
			if ((type != kDummy) && (j >= 0))
			{
				const double r2 = inner<D>(dx, dx);
				if (r2 < d2)
				{
					const Acc<D> dx0 = sub<D>(x0[j], x0i);
					double nn = 0.0;
#pragma unroll
					for (int k = 0; k < D; k++) { const double av = fabs(dx0.v[k]); nn += av * av; }
					const double nrm = sqrt(nn);
					Acc<D> e;
#pragma unroll
					for (int k = 0; k < D; k++) e.v[k] = dx0.v[k] / nrm;
					const double r_par = inner<D>(dx, e);
					Acc<D> perp;
#pragma unroll
					for (int k = 0; k < D; k++) perp.v[k] = dx.v[k] - r_par * e.v[k];
					const double r_perp2 = inner<D>(perp, perp);
					return scale<D>(sqrt(d2 - r_perp2) - r_par, e);
				}
				return azero<D>();
			}
			return azero<D>();
		});
	const double coef = -1.0 / (2 * dt * env.n0);
	Vec<D> out = vzero<D>();
#pragma unroll
	for (int k = 0; k < D; k++) out.v[k] = coef * s.v[k];
	du[i] = out;
}

// Computer::DetermineDt, Computer.hpp:759-777: max over ALL particles (Disabled included) of u.u
template<int D>
__global__ void __launch_bounds__(256) k_max_u2(uint64_t n, const Vec<D>* __restrict__ vel, DevScalars* sc)
{
	double m = 0.0;
	for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * 256 + threadIdx.x; i < n; i += static_cast<uint64_t>(gridDim.x) * 256)
	{
		const Vec<D> u = vel[i];
		double t = 0.0;
#pragma unroll
		for (int k = 0; k < D; k++) t += u.v[k] * u.v[k];
		// NaN never wins a '<' comparison in std::max_element either
		if (m < t) m = t;
	}
#pragma unroll
	for (int o = 16; o > 0; o >>= 1)
	{
		const double v = __shfl_xor_sync(0xffffffffu, m, o);
		if (m < v) m = v;
	}
	if ((threadIdx.x & 31) == 0) atomicMax(&sc->max_u2_bits, static_cast<unsigned long long>(__double_as_longlong(m)));
}

// dt = (maxU == 0) ? MaxDt : min(MaxDx / maxU, MaxDt) (Computer.hpp:775); then t += dt (Computer.hpp:1703-1706)
__global__ void k_set_dt(DevScalars* sc, double dt_in, int advance, int from_max_u, EnvConst env)
{
	double dt = dt_in;
	if (from_max_u)
	{
		const double max_u = sqrt(__longlong_as_double(static_cast<long long>(sc->max_u2_bits)));
		dt = (max_u == 0) ? env.max_dt : fmin(env.max_dx / max_u, env.max_dt);
	}
	sc->dt = dt;
	if (advance) sc->t += dt;
}

// ---- original order <-> slot order --------------------------------------------------------------------------------
template<int D>
__global__ void __launch_bounds__(kThreads) k_scatter_from_orig(uint64_t first, uint64_t count, bool append, Particles<D> P,
	uint32_t* __restrict__ inv, Vec<D>* __restrict__ wall, uint8_t* __restrict__ wall_group, const double* __restrict__ x, const double* __restrict__ u,
	const double* __restrict__ p, const double* __restrict__ nd, const int32_t* __restrict__ type)
{
	const uint64_t k = static_cast<uint64_t>(blockIdx.x) * kThreads + threadIdx.x;
	if (k >= count) return;
	const uint64_t o = first + k;
	uint64_t s;
	if (append) { s = o; inv[o] = static_cast<uint32_t>(o); P.orig[s] = static_cast<uint32_t>(o); wall_group[o] = 0; }
	else s = inv[o];
	if (x)
	{
		Vec<D> v = vzero<D>();
		for (int a = 0; a < D; a++) v.v[a] = x[k * D + a];
		P.pos[s] = v;
		if (append) wall[o] = v; // Main.cpp:304-315: non-fluid particles stay where they were added
	}
	if (u)
	{
		Vec<D> v = vzero<D>();
		for (int a = 0; a < D; a++) v.v[a] = u[k * D + a];
		P.vel[s] = v;
	}
	if (p) P.prs[s] = p[k];
	if (nd) P.nden[s] = nd[k];
	if (type) P.type[s] = static_cast<uint8_t>(type[k]);
}

template<int D>
__global__ void __launch_bounds__(kThreads) k_gather_to_orig(uint64_t n, Particles<D> P, double* __restrict__ x, double* __restrict__ u,
	double* __restrict__ p, double* __restrict__ nd, int32_t* __restrict__ type)
{
	const uint64_t s = static_cast<uint64_t>(blockIdx.x) * kThreads + threadIdx.x;
	if (s >= n) return;
	const uint64_t o = P.orig[s];
	if (x) { const Vec<D> v = P.pos[s]; for (int a = 0; a < D; a++) x[o * D + a] = v.v[a]; }
	if (u) { const Vec<D> v = P.vel[s]; for (int a = 0; a < D; a++) u[o * D + a] = v.v[a]; }
	if (p) p[o] = P.prs[s];
	if (nd) nd[o] = P.nden[s];
	if (type) type[o] = P.type[s];
}

template<int D>
__global__ void __launch_bounds__(kThreads) k_set_wall(uint64_t count, const uint64_t* __restrict__ ids, const double* __restrict__ x,
	Vec<D>* __restrict__ wall, uint64_t n)
{
	const uint64_t k = static_cast<uint64_t>(blockIdx.x) * kThreads + threadIdx.x;
	if (k >= count) return;
	const uint64_t o = ids[k];
	if (o >= n) return;
	Vec<D> v = vzero<D>();
	for (int a = 0; a < D; a++) v.v[a] = x[k * D + a];
	wall[o] = v;
}

// mps_set_wall_motion: the listed particles (ids == nullptr: every non-fluid particle) join motion group `group`
template<int D>
__global__ void __launch_bounds__(kThreads) k_set_wall_group(uint64_t count, const uint64_t* __restrict__ ids, int group, Particles<D> P,
	const uint32_t* __restrict__ inv, uint8_t* __restrict__ wall_group, uint64_t n)
{
	const uint64_t k = static_cast<uint64_t>(blockIdx.x) * kThreads + threadIdx.x;
	if (k >= count) return;
	const uint64_t o = ids ? ids[k] : k;
	if (o >= n) return;
	if (!ids && P.type[inv[o]] == kFluid) return;
	wall_group[o] = static_cast<uint8_t>(group);
}

// scalar (width 1) or vector (width D, padded source) per-slot array -> original order
template<int D>
__global__ void __launch_bounds__(kThreads) k_vec_to_orig(uint64_t n, const uint32_t* __restrict__ orig, const double* __restrict__ src,
	int width, int src_stride, double* __restrict__ out)
{
	const uint64_t s = static_cast<uint64_t>(blockIdx.x) * kThreads + threadIdx.x;
	if (s >= n) return;
	const uint64_t o = orig[s];
	for (int a = 0; a < width; a++) out[o * width + a] = src[s * src_stride + a];
}

template<int D>
Particles<D> view(mps_solver* s)
{
	Particles<D> p;
	p.pos = reinterpret_cast<Vec<D>*>(s->pos[s->cur].p);
	p.vel = reinterpret_cast<Vec<D>*>(s->vel[s->cur].p);
	p.prs = s->prs[s->cur].p;
	p.nden = s->nden[s->cur].p;
	p.type = s->type[s->cur].p;
	p.orig = s->orig[s->cur].p;
	return p;
}
inline Lists lists(mps_solver* s) { return Lists{ s->nbr_ptr.p, s->nbr.p }; }

#define MPS_TRY(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) return e_; } while (0)
#define MPS_DISPATCH(fn, ...) (s->env.dim == 2 ? fn<2>(__VA_ARGS__) : fn<3>(__VA_ARGS__))

// Every stage below runs over the rows this rank owns, [own0, own1): all rows on one GPU; one x-slab of the cell-sorted
// slots per rank when a communicator is attached (mps_comm.cu), in which case fields that neighbours read are
// all-gathered after the stages that write them.
template<int D> cudaError_t density(mps_solver* s, bool count_rows)
{
	const uint64_t r0 = s->own0(), r1 = s->own1();
	if (s->n == 0) return cudaSuccess;
	const unsigned nb = blocks_for(r1 - r0, kThreads);
	if (count_rows)
	{
		MPS_TRY(s->row_len.ensure(s->n, s->stream));
		if (s->comm.on)
		{
			// rows of other ranks have no entries here; originalX is needed for every row (DynamicStabilize reads neighbours')
			MPS_TRY(cudaMemsetAsync(s->row_len.p, 0, s->n * sizeof(uint32_t), s->stream));
			k_save_x<D><<<blocks_for(s->n, kThreads), kThreads, 0, s->stream>>>(s->n, view<D>(s), reinterpret_cast<Vec<D>*>(s->x0.p));
			s->stats.kernel_launches += 1;
		}
		if (nb) k_density<D, true><<<nb, kThreads, 0, s->stream>>>(r0, r1, view<D>(s), lists(s), s->nws.p, s->row_len.p, reinterpret_cast<Vec<D>*>(s->x0.p), s->env);
	}
	else if (nb)
		k_density<D, false><<<nb, kThreads, 0, s->stream>>>(r0, r1, view<D>(s), lists(s), s->nws.p, nullptr, nullptr, s->env);
	s->stats.kernel_launches += 1;
	MPS_TRY(comm_allgather_state(s, false, false, false, true)); // N is part of the state callers read back (CSV column n)
	return cudaGetLastError();
}
template<int D> cudaError_t ecs(mps_solver* s)
{
	const uint64_t r0 = s->own0(), r1 = s->own1();
	if (r1 == r0) return cudaSuccess;
	k_ecs<D><<<blocks_for(r1 - r0, kThreads), kThreads, 0, s->stream>>>(r0, r1, view<D>(s), lists(s), s->nws.p, s->ecs.p, s->env);
	s->stats.kernel_launches += 1;
	return cudaGetLastError();
}
template<int D> cudaError_t explicit_forces(mps_solver* s)
{
	const uint64_t r0 = s->own0(), r1 = s->own1();
	if (s->n == 0) return cudaSuccess;
	const unsigned nb = blocks_for(r1 - r0, kThreads);
	if (nb)
	{
		k_explicit_accel<D><<<nb, kThreads, 0, s->stream>>>(r0, r1, view<D>(s), lists(s), s->nws.p, reinterpret_cast<Vec<D>*>(s->du.p), s->env);
		k_explicit_move<D><<<nb, kThreads, 0, s->stream>>>(r0, r1, view<D>(s), reinterpret_cast<Vec<D>*>(s->du.p), reinterpret_cast<Vec<D>*>(s->wall.p), s->wall_group.p, s->motions, s->d_sc);
		s->stats.kernel_launches += 2;
	}
	MPS_TRY(comm_allgather_state(s, true, true, false)); // everybody needs the moved x, u
	return cudaGetLastError();
}
template<int D> cudaError_t save_x(mps_solver* s)
{
	if (s->n == 0) return cudaSuccess;
	k_save_x<D><<<blocks_for(s->n, kThreads), kThreads, 0, s->stream>>>(s->n, view<D>(s), reinterpret_cast<Vec<D>*>(s->x0.p));
	s->stats.kernel_launches += 1;
	return cudaGetLastError();
}
// rows lengths must be in s->row_len (k_density<COUNT_ROWS> in the step); `recount` recomputes them (stage-level API)
template<int D> cudaError_t ppe_fill(mps_solver* s, bool recount)
{
	const uint64_t n = s->n;
	if (n == 0) { s->cg.n = 0; return cudaSuccess; }
	const uint64_t r0 = s->own0(), r1 = s->own1();
	const unsigned nb = blocks_for(r1 - r0, kThreads);
	cudaStream_t st = s->stream;
	MPS_TRY(s->row_len.ensure(n, st));
	if (recount)
	{
		if (s->comm.on) MPS_TRY(cudaMemsetAsync(s->row_len.p, 0, n * sizeof(uint32_t), st));
		if (nb) k_row_len<D><<<nb, kThreads, 0, st>>>(r0, r1, view<D>(s), lists(s), s->row_len.p, s->env);
		s->stats.kernel_launches += 1;
	}
	CgBuffers& cg = s->cg;
	MPS_TRY(cg.rowptr.ensure(n + 1, st));
	MPS_TRY(launch_exclusive_scan_u32_to_u64(s->row_len.p, cg.rowptr.p, n, s->scan_tmp, st, &s->stats.kernel_launches));
	// the streaming kernel stages even-aligned windows: one element of slack behind every vector
	MPS_TRY(cg.b.ensure(n + 64, st)); MPS_TRY(cg.x.ensure(n + 64, st)); if (!cg.chunked) MPS_TRY(cg.r.ensure(n + 64, st));
	MPS_TRY(cg.ap.ensure(n + 64, st));
	const bool pre_on = mg_wanted(s); // == mg_active(s) once cg.external is reset below
	if (pre_on) MPS_TRY(cg.r.ensure(n + 64, st));
	if (cg.chunked && s->comm.on) MPS_TRY(comm_ensure_arena(s, n + 64)); // {r, p} live in the arena the neighbour ranks map
	else if (cg.chunked) { MPS_TRY(cg.z0.ensure(2 * (n + 64), st)); MPS_TRY(cg.z1.ensure(2 * (n + 64), st)); }
	else { MPS_TRY(cg.p0.ensure(n + 64, st)); MPS_TRY(cg.p1.ensure(n + 64, st)); }
	cg.n = n; cg.external = false;
	PpeOut out{};
	if (cg.chunked)
	{
		MPS_TRY(launch_chunk_build(s));
		out.desc = cg.desc.p; out.chunk_of_row = cg.chunk_of_row.p; out.blobs = cg.blobs.p;
	}
	else
	{
		// nnz <= neighbour entries + n (every row adds its diagonal): no host round trip needed to size the CSR
		const uint64_t bound = s->nbr_total + n;
		MPS_TRY(cg.col.ensure(bound, st)); MPS_TRY(cg.val.ensure(bound, st));
		out.row_ptr = cg.rowptr.p; out.col = cg.col.p; out.val = cg.val.p;
	}
	MPS_TRY(cudaMemsetAsync(&s->d_sc->active_rows, 0, sizeof(unsigned long long), st));
	if (s->comm.on)
	{
		// rows of other ranks: b = 0 and the initial guess x = P (their windows are gathered like any other column)
		k_ppe_guess<D><<<blocks_for(n, kThreads), kThreads, 0, st>>>(n, view<D>(s), cg.b.p, cg.x.p);
		s->stats.kernel_launches += 1;
	}
	PpePre pre{};
	if (pre_on) { pre.skey = s->skey.p; pre.row_s = s->mg.row_s.p; pre.dinv0 = s->mg.dinv0.p; pre.stride = r1 - r0; pre.row0 = r0; }
	if (nb)
	{
		if (cg.chunked && pre_on)
			k_ppe_fill<D, true, true><<<nb, kThreads, 0, st>>>(r0, r1, view<D>(s), lists(s), s->nws.p, s->ecs.p, out, pre, cg.b.p, cg.x.p, s->env, s->d_sc);
		else if (cg.chunked)
			k_ppe_fill<D, true, false><<<nb, kThreads, 0, st>>>(r0, r1, view<D>(s), lists(s), s->nws.p, s->ecs.p, out, pre, cg.b.p, cg.x.p, s->env, s->d_sc);
		else
			k_ppe_fill<D, false, false><<<nb, kThreads, 0, st>>>(r0, r1, view<D>(s), lists(s), s->nws.p, s->ecs.p, out, pre, cg.b.p, cg.x.p, s->env, s->d_sc);
		s->stats.kernel_launches += 1;
	}
	MPS_TRY(cudaMemcpyAsync(&s->d_sc->nnz_total, cg.rowptr.p + n, sizeof(uint64_t), cudaMemcpyDeviceToDevice, st));
	if (pre_on) MPS_TRY(launch_mg_setup(s)); // topology + Galerkin operators of the cell hierarchy for this assembly
	return cudaGetLastError();
}
template<int D> cudaError_t assign_pressure(mps_solver* s)
{
	const uint64_t r0 = s->own0(), r1 = s->own1();
	if (s->n == 0) return cudaSuccess;
	if (r1 > r0)
	{
		k_assign_pressure<D><<<blocks_for(r1 - r0, kThreads), kThreads, 0, s->stream>>>(r0, r1, view<D>(s), s->cg.x.p);
		s->stats.kernel_launches += 1;
	}
	MPS_TRY(comm_allgather_state(s, false, false, true)); // the pressure gradient reads the neighbours' pressures
	return cudaGetLastError();
}
template<int D> cudaError_t gradient(mps_solver* s)
{
	const uint64_t r0 = s->own0(), r1 = s->own1();
	if (s->n == 0) return cudaSuccess;
	const unsigned nb = blocks_for(r1 - r0, kThreads);
	if (nb)
	{
		k_gradient<D><<<nb, kThreads, 0, s->stream>>>(r0, r1, view<D>(s), lists(s), s->nws.p, reinterpret_cast<Vec<D>*>(s->du.p), s->env, s->d_sc);
		k_apply_du<D><<<nb, kThreads, 0, s->stream>>>(r0, r1, view<D>(s), reinterpret_cast<Vec<D>*>(s->du.p), s->d_sc);
		s->stats.kernel_launches += 2;
	}
	MPS_TRY(comm_allgather_state(s, true, true, false));
	return cudaGetLastError();
}
template<int D> cudaError_t ds(mps_solver* s)
{
	const uint64_t r0 = s->own0(), r1 = s->own1();
	if (s->n == 0) return cudaSuccess;
	const unsigned nb = blocks_for(r1 - r0, kThreads);
	if (nb)
	{
		k_ds<D><<<nb, kThreads, 0, s->stream>>>(r0, r1, view<D>(s), lists(s), s->nws.p, reinterpret_cast<Vec<D>*>(s->x0.p), reinterpret_cast<Vec<D>*>(s->du.p), s->env, s->d_sc);
		k_apply_du<D><<<nb, kThreads, 0, s->stream>>>(r0, r1, view<D>(s), reinterpret_cast<Vec<D>*>(s->du.p), s->d_sc);
		s->stats.kernel_launches += 2;
	}
	MPS_TRY(comm_allgather_state(s, true, true, false));
	return cudaGetLastError();
}
template<int D> cudaError_t max_u2(mps_solver* s)
{
	MPS_TRY(cudaMemsetAsync(&s->d_sc->max_u2_bits, 0, sizeof(unsigned long long), s->stream));
	if (s->n == 0) return cudaSuccess;
	unsigned nb = blocks_for(s->n, 256);
	const unsigned cap = static_cast<unsigned>(s->sm_count) * 8u;
	if (nb > cap) nb = cap;
	k_max_u2<D><<<nb, 256, 0, s->stream>>>(s->n, reinterpret_cast<Vec<D>*>(s->vel[s->cur].p), s->d_sc);
	s->stats.kernel_launches += 1;
	return cudaGetLastError();
}
template<int D> cudaError_t dndt_one(mps_solver* s, uint64_t orig_id, double* d_out)
{
	k_dndt_one<D><<<1, 1, 0, s->stream>>>(orig_id, s->inv.p, view<D>(s), lists(s), s->nws.p, d_out, s->env);
	s->stats.kernel_launches += 1;
	return cudaGetLastError();
}
template<int D> cudaError_t scatter_from_orig(mps_solver* s, const double* x, const double* u, const double* p, const double* nd,
	const int32_t* type, uint64_t first, uint64_t count, bool append)
{
	if (count == 0) return cudaSuccess;
	k_scatter_from_orig<D><<<blocks_for(count, kThreads), kThreads, 0, s->stream>>>(first, count, append, view<D>(s), s->inv.p,
		reinterpret_cast<Vec<D>*>(s->wall.p), s->wall_group.p, x, u, p, nd, type);
	s->stats.kernel_launches += 1;
	return cudaGetLastError();
}
template<int D> cudaError_t gather_to_orig(mps_solver* s, double* x, double* u, double* p, double* nd, int32_t* type)
{
	if (s->n == 0) return cudaSuccess;
	k_gather_to_orig<D><<<blocks_for(s->n, kThreads), kThreads, 0, s->stream>>>(s->n, view<D>(s), x, u, p, nd, type);
	s->stats.kernel_launches += 1;
	return cudaGetLastError();
}
template<int D> cudaError_t set_wall(mps_solver* s, uint64_t count, const uint64_t* ids, const double* x)
{
	if (count == 0) return cudaSuccess;
	k_set_wall<D><<<blocks_for(count, kThreads), kThreads, 0, s->stream>>>(count, ids, x, reinterpret_cast<Vec<D>*>(s->wall.p), s->n);
	s->stats.kernel_launches += 1;
	return cudaGetLastError();
}
template<int D> cudaError_t set_wall_group(mps_solver* s, uint64_t count, const uint64_t* ids, int group)
{
	if (count == 0) return cudaSuccess;
	k_set_wall_group<D><<<blocks_for(count, kThreads), kThreads, 0, s->stream>>>(count, ids, group, view<D>(s), s->inv.p, s->wall_group.p, s->n);
	s->stats.kernel_launches += 1;
	return cudaGetLastError();
}
template<int D> cudaError_t vec_to_orig(mps_solver* s, int which, double* out)
{
	const uint64_t n = (which <= 4) ? s->cg.n : s->n;
	if (n == 0) return cudaSuccess;
	const double* src = nullptr; int width = 1, stride = 1;
	switch (which)
	{
	case 0: src = s->cg.x.p; break;
	case 1: src = s->cg.b.p; break;
	case 2: src = s->cg.r.p; break;
	case 3: src = s->cg.p0.p; break; // generic kernel: direction of the last completed iteration is in one of p0 / p1
	case 4: src = s->cg.ap.p; break;
	case 5: src = s->ecs.p; break;
	case 6: src = s->nws.p; break;
	case 7: src = s->du.p; width = D; stride = s->vec_stride(); break;
	case 8: src = s->x0.p; width = D; stride = s->vec_stride(); break;
	default: return cudaErrorInvalidValue;
	}
	if (s->cg.chunked && !s->cg.external && (which == 2 || which == 3))
	{
		// streaming kernel: {r, p} interleaved; h_sc->z_final (refreshed by the caller) says which buffer is current
		src = (s->h_sc->z_final ? s->cg.z1.p : s->cg.z0.p) + (which == 3 ? 1 : 0);
		stride = 2;
	}
	if (s->cg.external && which <= 4)
	{
		// an externally loaded system is already in the caller's order
		return cudaMemcpyAsync(out, src, n * sizeof(double), cudaMemcpyDeviceToDevice, s->stream);
	}
	k_vec_to_orig<D><<<blocks_for(n, kThreads), kThreads, 0, s->stream>>>(n, s->orig[s->cur].p, src, width, stride, out);
	s->stats.kernel_launches += 1;
	return cudaGetLastError();
}

} // namespace

cudaError_t launch_density(mps_solver* s, bool count_rows) { return MPS_DISPATCH(density, s, count_rows); }
cudaError_t launch_ecs(mps_solver* s) { return MPS_DISPATCH(ecs, s); }
cudaError_t launch_explicit(mps_solver* s) { return MPS_DISPATCH(explicit_forces, s); }
cudaError_t launch_save_x(mps_solver* s) { return MPS_DISPATCH(save_x, s); }
cudaError_t launch_ppe_fill(mps_solver* s) { return MPS_DISPATCH(ppe_fill, s, true); }
cudaError_t launch_ppe_fill_counted(mps_solver* s) { return MPS_DISPATCH(ppe_fill, s, false); }
cudaError_t launch_assign_pressure(mps_solver* s) { return MPS_DISPATCH(assign_pressure, s); }
cudaError_t launch_gradient(mps_solver* s) { return MPS_DISPATCH(gradient, s); }
cudaError_t launch_ds(mps_solver* s) { return MPS_DISPATCH(ds, s); }
cudaError_t launch_max_u2(mps_solver* s) { return MPS_DISPATCH(max_u2, s); }
cudaError_t launch_set_dt(mps_solver* s, double dt, int advance, bool from_max_u)
{
	k_set_dt<<<1, 1, 0, s->stream>>>(s->d_sc, dt, advance, from_max_u ? 1 : 0, s->env);
	s->stats.kernel_launches += 1;
	return cudaGetLastError();
}
cudaError_t launch_dndt_one(mps_solver* s, uint64_t orig_id, double* d_out) { return MPS_DISPATCH(dndt_one, s, orig_id, d_out); }
cudaError_t launch_scatter_from_orig(mps_solver* s, const double* d_x, const double* d_u, const double* d_p, const double* d_n,
	const int32_t* d_type, uint64_t first, uint64_t count, bool append)
{
	return MPS_DISPATCH(scatter_from_orig, s, d_x, d_u, d_p, d_n, d_type, first, count, append);
}
cudaError_t launch_gather_to_orig(mps_solver* s, double* d_x, double* d_u, double* d_p, double* d_n, int32_t* d_type)
{
	return MPS_DISPATCH(gather_to_orig, s, d_x, d_u, d_p, d_n, d_type);
}
cudaError_t launch_set_wall_group(mps_solver* s, uint64_t count, const uint64_t* d_ids, int group) { return MPS_DISPATCH(set_wall_group, s, count, d_ids, group); }
cudaError_t launch_set_wall(mps_solver* s, uint64_t count, const uint64_t* d_ids, const double* d_x) { return MPS_DISPATCH(set_wall, s, count, d_ids, d_x); }
cudaError_t launch_gather_vec_to_orig(mps_solver* s, int which, double* d_out) { return MPS_DISPATCH(vec_to_orig, s, which, d_out); }

} // namespace mps
