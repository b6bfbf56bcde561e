// mps_observe.cu — the reference's benchmark observables as device reductions over the resident particle state (SURVEY.md 8f
// rank 3: "GPU-side reductions ... so long-run physical parity is a test").  The reference computes them in plotting scripts that
// re-read result/particles_%05d.csv:
//   leading edge of the dam break      Benchmark/DamBreak/koshizukaoka1996_edge.py:14-20   max x over Type 0
//   Zhou et al. probes                 Benchmark/DamBreak/zhouetal1999.py:26-39            water height at two stations, wall pressure
//   roundness / centre pressure        Benchmark/CentralGravity/check_result.py:26-57      r_min / r_max over n < beta n0, p nearest the origin
//   hydrostatic column                 Benchmark/StaticPressure (generator only)           moments of (depth, p) over fluid with p > 0
// Here: ONE pass over the state (k_observe: every block folds its particles into a fixed set of accumulators, per-block partials)
// + ONE block that folds the partials in block order (k_observe_fold): deterministic, no atomics on doubles, 24 doubles back to
// the host instead of the 52 B/particle download.  HBM-bound streaming: 8 D + 8 + 8 + 1 + 4 bytes per particle.
#include <cfloat>

#include "mps_solver.h"

namespace mps {
namespace {

constexpr int kThreads = 256;
constexpr int kSlots = 24;

enum Slot
{
	kEdgeX = 0, kTopZ,                          // max
	kRMaxSurf, kRMinSurf,                       // max, min
	kCenterR2, kCenterId, kCenterP,             // lexicographic min of (r^2, original id) and the pressure that goes with it
	kInner, kSumD, kSumP, kSumDD, kSumDP, kMaxDev, // sums + max
	kH1, kH2, kP2Sum, kP2Count,
	kNFluid, kNWall, kNDummy, kNDisabled,
	kPMax, kUMax2, kSurfCount
};

__device__ __forceinline__ void fold(double* a, const double* b)
{
	a[kEdgeX] = fmax(a[kEdgeX], b[kEdgeX]); a[kTopZ] = fmax(a[kTopZ], b[kTopZ]);
	a[kRMaxSurf] = fmax(a[kRMaxSurf], b[kRMaxSurf]); a[kRMinSurf] = fmin(a[kRMinSurf], b[kRMinSurf]);
	if (b[kCenterR2] < a[kCenterR2] || (b[kCenterR2] == a[kCenterR2] && b[kCenterId] < a[kCenterId]))
	{
		a[kCenterR2] = b[kCenterR2]; a[kCenterId] = b[kCenterId]; a[kCenterP] = b[kCenterP];
	}
	a[kInner] += b[kInner]; a[kSumD] += b[kSumD]; a[kSumP] += b[kSumP]; a[kSumDD] += b[kSumDD]; a[kSumDP] += b[kSumDP];
	a[kMaxDev] = fmax(a[kMaxDev], b[kMaxDev]);
	a[kH1] = fmax(a[kH1], b[kH1]); a[kH2] = fmax(a[kH2], b[kH2]);
	a[kP2Sum] += b[kP2Sum]; a[kP2Count] += b[kP2Count];
	a[kNFluid] += b[kNFluid]; a[kNWall] += b[kNWall]; a[kNDummy] += b[kNDummy]; a[kNDisabled] += b[kNDisabled];
	a[kPMax] = fmax(a[kPMax], b[kPMax]); a[kUMax2] = fmax(a[kUMax2], b[kUMax2]);
	a[kSurfCount] += b[kSurfCount];
}
__device__ __forceinline__ void identity(double* a)
{
#pragma unroll
	for (int k = 0; k < kSlots; k++) a[k] = 0.0;
	a[kEdgeX] = -DBL_MAX; a[kTopZ] = -DBL_MAX; a[kRMaxSurf] = -DBL_MAX; a[kRMinSurf] = DBL_MAX;
	a[kCenterR2] = DBL_MAX; a[kCenterId] = DBL_MAX; a[kPMax] = -DBL_MAX;
}

struct ObsArgs
{
	double surface_n;      // check_result.py:41: surface particles have n < beta n0 (lattice n0)
	double rho_g, h;       // hydrostatic reference p = rho g (h - z)
	double x_h1, x_h2, half_l0, min_n, z_p2, half_d; // zhouetal1999.py:26-39
};

// folds this thread's accumulators over the block (shuffles, then one row per warp in shared memory) — fixed order
__device__ __forceinline__ void block_fold(double* acc, double (*sh)[kSlots])
{
	const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
	for (int o = 16; o > 0; o >>= 1)
	{
		double other[kSlots];
#pragma unroll
		for (int k = 0; k < kSlots; k++) other[k] = __shfl_down_sync(0xffffffffu, acc[k], o);
		fold(acc, other);
	}
	if (lane == 0)
	{
#pragma unroll
		for (int k = 0; k < kSlots; k++) sh[warp][k] = acc[k];
	}
	__syncthreads();
	if (threadIdx.x == 0)
		for (unsigned w = 1; w < blockDim.x / 32; w++) fold(acc, sh[w]);
}

template<int D>
__global__ void __launch_bounds__(kThreads) k_observe(const uint64_t n, const Vec<D>* __restrict__ pos, const Vec<D>* __restrict__ vel,
	const double* __restrict__ prs, const double* __restrict__ nden, const uint8_t* __restrict__ type, const uint32_t* __restrict__ orig,
	const ObsArgs a, double* __restrict__ partials)
{
	__shared__ double sh[kThreads / 32][kSlots];
	double acc[kSlots];
	identity(acc);
	for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * kThreads + threadIdx.x; i < n; i += static_cast<uint64_t>(gridDim.x) * kThreads)
	{
		const Vec<D> x = pos[i];
		const Vec<D> u = vel[i];
		const double p = prs[i], nd = nden[i];
		const int t = type[i];
		const double px = x.v[0], pz = x.v[D - 1];
		double r2 = 0.0, u2 = 0.0;
#pragma unroll
		for (int d = 0; d < D; d++) { r2 += x.v[d] * x.v[d]; u2 += u.v[d] * u.v[d]; }
		double one[kSlots];
		identity(one);
		one[kNFluid] = (t == MPS_FLUID); one[kNWall] = (t == MPS_WALL); one[kNDummy] = (t == MPS_DUMMY); one[kNDisabled] = (t == MPS_DISABLED);
		if (t != MPS_DISABLED) { one[kPMax] = p; one[kUMax2] = u2; }
		if (t == MPS_FLUID)
		{
			one[kEdgeX] = px; one[kTopZ] = pz;
			if (p > 0.0)
			{
				const double depth = a.h - pz;
				one[kInner] = 1.0; one[kSumD] = depth; one[kSumP] = p; one[kSumDD] = depth * depth; one[kSumDP] = depth * p;
				one[kMaxDev] = fabs(p - a.rho_g * depth);
			}
		}
		// check_result.py:41-50 looks at every row of the CSV
		const double r = sqrt(r2);
		if (nd < a.surface_n) { one[kRMaxSurf] = r; one[kRMinSurf] = r; one[kSurfCount] = 1.0; }
		one[kCenterR2] = r2; one[kCenterId] = static_cast<double>(orig[i]); one[kCenterP] = p;
		// zhouetal1999.py:29-36
		if (nd > a.min_n)
		{
			if (fabs(px - a.x_h1) < a.half_l0) one[kH1] = fmax(0.0, pz);
			if (fabs(px - a.x_h2) < a.half_l0) one[kH2] = fmax(0.0, pz);
		}
		if (t == MPS_WALL && px < 0.0 && fabs(pz - a.z_p2) < a.half_d) { one[kP2Sum] = p; one[kP2Count] = 1.0; }
		fold(acc, one);
	}
	block_fold(acc, sh);
	if (threadIdx.x == 0)
	{
#pragma unroll
		for (int k = 0; k < kSlots; k++) partials[static_cast<size_t>(blockIdx.x) * kSlots + k] = acc[k];
	}
}

__global__ void __launch_bounds__(kThreads) k_observe_fold(const unsigned blocks, const double* __restrict__ partials, double* __restrict__ out)
{
	__shared__ double sh[kThreads / 32][kSlots];
	double acc[kSlots];
	identity(acc);
	for (unsigned b = threadIdx.x; b < blocks; b += kThreads) fold(acc, partials + static_cast<size_t>(b) * kSlots);
	block_fold(acc, sh);
	if (threadIdx.x == 0)
	{
#pragma unroll
		for (int k = 0; k < kSlots; k++) out[k] = acc[k];
	}
}

#define MPS_TRY(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) return e_; } while (0)

} // namespace

// out: kSlots doubles on the HOST (pinned or not); synchronises the stream
cudaError_t launch_observe(mps_solver* s, const mps_observe_params* p, double* host_out)
{
	const uint64_t n = s->n;
	cudaStream_t st = s->stream;
	unsigned blocks = static_cast<unsigned>(s->sm_count) * 4u;
	const unsigned need = blocks_for(n ? n : 1, kThreads);
	if (need < blocks) blocks = need;
	MPS_TRY(s->obs.ensure(static_cast<size_t>(blocks + 1) * kSlots, st));
	ObsArgs a;
	a.surface_n = p->surface_n; a.rho_g = p->rho_g; a.h = p->surface_z;
	a.x_h1 = p->x_h1; a.x_h2 = p->x_h2; a.half_l0 = 0.5 * s->env.l0; a.min_n = p->min_n; a.z_p2 = p->z_p2; a.half_d = 0.5 * p->d;
	double* partials = s->obs.p + kSlots;
	if (s->env.dim == 2)
		k_observe<2><<<blocks, kThreads, 0, st>>>(n, reinterpret_cast<const Vec<2>*>(s->pos[s->cur].p), reinterpret_cast<const Vec<2>*>(s->vel[s->cur].p),
			s->prs[s->cur].p, s->nden[s->cur].p, s->type[s->cur].p, s->orig[s->cur].p, a, partials);
	else
		k_observe<3><<<blocks, kThreads, 0, st>>>(n, reinterpret_cast<const Vec<3>*>(s->pos[s->cur].p), reinterpret_cast<const Vec<3>*>(s->vel[s->cur].p),
			s->prs[s->cur].p, s->nden[s->cur].p, s->type[s->cur].p, s->orig[s->cur].p, a, partials);
	k_observe_fold<<<1, kThreads, 0, st>>>(blocks, partials, s->obs.p);
	s->stats.kernel_launches += 2;
	MPS_TRY(cudaGetLastError());
	MPS_TRY(cudaMemcpyAsync(host_out, s->obs.p, kSlots * sizeof(double), cudaMemcpyDeviceToHost, st));
	return cudaStreamSynchronize(st);
}

} // namespace mps
