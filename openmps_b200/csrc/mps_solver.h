// mps_solver.h — host-side state of one solver handle and the launch functions each translation unit provides.
#pragma once

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/mps_capi.h"
#include "mps_device.cuh"

namespace mps {

// stage slots of mps_stats.stage_ms
enum Stage
{
	kStSort = 0, kStSearch, kStDensity, kStEcs, kStExplicit, kStPpeAssemble, kStCg, kStPressure, kStGradient, kStDs,
	kStDt, kStCount
};

// head-room of a growing device buffer in percent (MPS_ALLOC_SLACK, default 25: re-allocations become rare as particles or
// neighbours fluctuate; 0 for blocks that only just fit the GPU, e.g. 100M particles on one B200)
inline size_t alloc_slack_percent()
{
	static const size_t v = [] { const char* e = std::getenv("MPS_ALLOC_SLACK"); const long k = e ? std::atol(e) : 25; return static_cast<size_t>(k < 0 ? 0 : (k > 100 ? 100 : k)); }();
	return v;
}

template<typename T>
struct DevBuf
{
	T* p = nullptr;
	size_t cap = 0;
	// grows geometrically; `keep` preserves the first `keep_n` elements
	cudaError_t ensure(size_t n, cudaStream_t s = nullptr, size_t keep_n = 0)
	{
		if (n <= cap) return cudaSuccess;
		size_t want = n + n / 100 * alloc_slack_percent() + 16;
		T* q = nullptr;
		cudaError_t e = cudaMalloc(&q, want * sizeof(T));
		if (e != cudaSuccess) { want = n; e = cudaMalloc(&q, want * sizeof(T)); }
		if (e != cudaSuccess) return e;
		if (p && keep_n)
		{
			e = cudaMemcpyAsync(q, p, keep_n * sizeof(T), cudaMemcpyDeviceToDevice, s);
			if (e != cudaSuccess) return e;
			e = cudaStreamSynchronize(s);
			if (e != cudaSuccess) return e;
		}
		if (p) cudaFree(p);
		p = q; cap = want;
		return cudaSuccess;
	}
	void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct CgBuffers
{
	uint64_t n = 0;          // rows of the loaded system
	DevBuf<uint64_t> rowptr; // n + 1
	DevBuf<uint32_t> col;
	DevBuf<double> val;
	DevBuf<double> b, x, r, p0, p1, ap;
	DevBuf<double> z0, z1;   // streaming kernel: {r_i, p_i} interleaved (2 doubles per row), ping-pong
	DevBuf<double> partials; // 2 x max grid size
	bool external = false;   // loaded through mps_set_system (not assembled from particles)
	bool z_borrowed = false; // z0 / z1 point into the multi-GPU peer arena (mps_comm.cu) and are not owned by the DevBufs

	// chunk-blob form of the assembled system (mps_device.cuh, mps_chunk.cu); rowptr above is still produced (row lengths
	// for inspection, nnz), col / val are only used by externally loaded systems
	bool chunked = false;    // the assembled system is in chunk-blob form and is solved by the streaming kernel
	ChunkLimits limits{};
	int stages = 0;          // shared-memory pipeline depth of the streaming kernel
	int consumer_warps = 8;  // consumer warps per CTA of the streaming kernel
	int producers = 1;       // + producer warps (issue the bulk copies)
	int lanes_per_row = 1;   // 1 (2-D: ~21 entries per row) or 4 (3-D: ~57)
	DevBuf<uint32_t> blk_chunks, blk_bytes, blk_cost, blk_live, chunk_of_row;
	DevBuf<uint64_t> chunk_base, blob_base, cost_base, live_base;
	DevBuf<ChunkDesc> desc;
	DevBuf<ChunkDesc> live;  // the descriptors of the chunks that have entries, in order (runs of Dummy / Disabled rows and, on several
	                         // GPUs, the other ranks' rows are not in it): what the CG kernel's CTAs split and walk
	DevBuf<unsigned char> blobs;
	uint64_t desc_cap = 0;
	// load balance of the streaming kernel's CTAs, learnt from the kernel's own cycle counters (mps_cg.cu k_cg_rebalance):
	DevBuf<double> cta_frac;         // [grid + 1] cumulative share of the modelled cost that CTAs 0..b-1 take (uniform at first)
	DevBuf<double> cta_speed;        // [grid] smoothed relative speed (modelled cost per cycle) of each CTA
	DevBuf<unsigned long long> cta_meas; // [2 x grid] {modelled cost taken, SpMV cycles} of the last solve
	bool adaptive = true;            // MPS_CG_ADAPTIVE=0 freezes the uniform split (bit-reproducible run to run)
	DevBuf<CgStepScalars> step;      // multi-GPU stepwise solve: scalars of the recurrence on the device ...
	CgStepScalars* h_step = nullptr; // ... and their pinned host mirror (convergence test)
	DevBuf<unsigned long long> prof; // per-CTA cycle counters of the last streaming solve (mps_get_cg_profile)
	unsigned prof_blocks = 0;
	bool prof_stages = false;        // prof holds 64 per-stage counters of the preconditioned solve behind the per-CTA ones
};

// multigrid preconditioner of the PPE solve: per-level device buffers (mps_mg.cu)
struct MgLevelBufs
{
	long long dims[3] = { 1, 1, 1 }; // dense grid of the level
	uint64_t dense = 0;              // product of dims
	uint64_t bound = 0;              // upper bound of the occupied cells (sizes the compact arrays and the launches)
	DevBuf<uint32_t> flag;           // [dense] occupied?
	DevBuf<uint64_t> rank;           // [dense + 1] exclusive scan of flag = compact id; last entry = occupied cells
	DevBuf<uint32_t> key, nbr, child, parent;
	DevBuf<double> S, dinv, r, e0, e1;
};
struct MgBuffers
{
	bool on = false;                 // MPS_CG_PRECOND (default on): assembled systems are solved by k_pcg_stream
	int levels = 0;
	double omega = 0.8, gamma = 1.8;
	int top_sweeps = 4;
	uint32_t top_cells = 64;
	uint32_t small_cells = 256;      // MPS_MG_SMALL_CELLS
	uint64_t cells0 = 0;             // occupied cells of the neighbour grid at the last sort (read back with the list size)
	MgLevelBufs lv[kMgMaxLevels];
	DevBuf<uint32_t> crow;
	DevBuf<uint64_t> cstart;
	DevBuf<double> dinv0, row_s;
	// several ranks: the level vectors r / e0 / e1 live in the "mg" section of the peer arena (same layout on every rank) so that
	// neighbour ranks can read halo cells; MgLevelBufs::r / e0 / e1 are then unused
	bool in_arena = false;
	int k_dist = 0;                  // 1: level 0 distributed over the ranks, the rest replicated; 0: everything replicated (MgDist)
	uint64_t dist_cells = 150000;    // level 0 is distributed when it holds more cells than this (MPS_MG_DIST_CELLS)
	uint64_t vec_off[kMgMaxLevels][4] = {}; // r, e0, e1, part of each level inside the mg section (doubles)
	uint64_t vec_total = 0;
};

} // namespace mps

namespace mps {
// one rank of a multi-GPU run (mps_comm.cu): NCCL communicator + this rank's slab of the cell-sorted slots
struct Comm
{
	bool on = false;
	int rank = 0, nranks = 1;
	void* nccl = nullptr;     // ncclComm_t
	DevBuf<unsigned long long> ext;  // halo extents of all ranks (2 per rank), device

	// peer-memory coupling of the persistent CG kernel (mps_comm.cu comm_peer_*): one arena per rank
	// [mailbox | z0 | z1], exported with cudaIpcGetMemHandle and mapped by every other rank
	int peer_mode = 0;                 // 0 not decided yet, 1 peer memory (persistent kernel), 2 NCCL stepwise (IPC unavailable)
	unsigned char* arena = nullptr;    // this rank's arena
	size_t arena_bytes = 0;
	uint64_t arena_rows = 0;           // rows the z sections are sized for
	uint64_t arena_mg = 0;             // doubles of the mg section behind them (level vectors of the preconditioner)
	unsigned char* peer_arena[8] = {}; // peers' arenas mapped into this process ([rank] = arena)
	void* peer_base[8] = {};           // what cudaIpcOpenMemHandle returned (to close)
	unsigned long long solves = 0;     // persistent solves so far (tag of the mailbox flags)
	// the replicated particle state over peer memory: pos[2], vel[2], prs[2], nden[2] of every rank mapped here, so that the
	// per-stage all-gathers are one copy kernel over NVLink between two flag barriers (no NCCL in the step)
	const void* exported[8] = {};      // the local pointers the current exports refer to (a re-allocation forces a new exchange)
	unsigned char* peer_state[8][8] = {}; // [rank][2 * field + buffer]
	void* peer_state_base[8][8] = {};
	unsigned long long bar_seq = 0;    // stream-level barriers so far (monotonic flag value)
	uint64_t peer_gathers = 0;
	PeerLink link{};                   // what the next solve's kernel gets
	bool halo_step = false;            // inside ForwardTime: per-stage gathers pull the halo only (mps_comm.cu comm_allgather_state)
};
} // namespace mps

struct mps_solver
{
	int device = 0;
	int sm_count = 148;
	cudaStream_t stream = nullptr;
	mps::EnvConst env{};
	std::string last_error;
	std::string comm_error;  // text of the last NCCL failure (reported through last_error)

	uint64_t n = 0;          // particles (all types, including Disabled)
	int cur = 0;             // which of the two state copies is live

	// state (double-buffered for the per-step permutation); void* because Vec<D> depends on dim
	mps::DevBuf<double> pos[2], vel[2];    // n * vec_stride doubles
	mps::DevBuf<double> prs[2], nden[2];
	mps::DevBuf<uint8_t> type[2];
	mps::DevBuf<uint32_t> orig[2];
	mps::DevBuf<uint32_t> inv;             // original id -> slot
	mps::DevBuf<double> wall;              // target position of non-fluid particles, ORIGINAL order, vec_stride doubles
	mps::DevBuf<uint8_t> wall_group;       // ORIGINAL order: 0 = follows wall[], g = wall[] + motions.m[g - 1](t) (mps_set_wall_motion)
	mps::WallMotions motions{};            // analytic wall motions evaluated inside k_explicit_move

	// scratch per particle
	mps::DevBuf<double> nws, ecs, du, x0;  // du, x0: vec_stride doubles

	// grid
	mps::DevBuf<uint32_t> key, skey, rank, perm, perm_orig, perm2;
	mps::DevBuf<uint32_t> cell_count;      // ncells + 1
	mps::DevBuf<uint64_t> cell_start;      // ncells + 2
	mps::DevBuf<uint64_t> scan_tmp;

	// neighbour list
	mps::DevBuf<uint32_t> nbr_cnt;
	mps::DevBuf<uint64_t> nbr_ptr;
	mps::DevBuf<uint32_t> nbr;
	uint64_t nbr_total = 0;
	bool searched = false;
	int sort_error = 0;      // device error flag as read back with the list size (MPS_CELL_OVERFLOW stops the step there)

	// PPE
	mps::DevBuf<uint32_t> row_len;
	mps::CgBuffers cg;
	mps::MgBuffers mg;
	uint64_t nnz_total = 0;

	// staging for upload / download (original order)
	mps::DevBuf<double> stage_d;
	mps::DevBuf<int32_t> stage_i;

	mps::DevScalars* d_sc = nullptr;       // device
	mps::DevScalars* h_sc = nullptr;       // pinned host mirror

	mps::DevBuf<double> obs;               // observables: [result | per-block partials] (mps_observe.cu)

	// L2 flush buffer
	mps::DevBuf<char> flush;

	// stats
	mps_stats stats{};
	bool stage_timing = false;
	cudaEvent_t ev0 = nullptr, ev1 = nullptr;
	cudaEvent_t ev_cg0 = nullptr, ev_cg1 = nullptr; // always-on timing of the CG kernel (resolved at the step's own sync)
	int cg_max_blocks_per_sm = 0;
	bool cg_profile = false;

	int vec_stride() const { return env.dim == 2 ? 2 : 4; }

	// multi-GPU: rows (slots) this rank computes; the state itself is replicated (DESIGN.md "multi-GPU").  Every sort cuts the
	// cell-sorted slots into equal shares of modelled work and moves each cut to the nearest boundary between cell COLUMNS (all
	// cells of one x index), so that every cell belongs to exactly one rank (mps_grid.cu k_slab_bounds).
	mps::Comm comm;
	std::vector<uint64_t> own_b;           // [nranks + 1] first slot of every rank's slab (valid while own_n == n)
	std::vector<uint32_t> col_b;           // [nranks + 1] first cell column of every rank's slab
	std::vector<uint64_t> halo_lo, halo_hi; // [nranks] slots [halo_lo[r], own_b[r]) and [own_b[r + 1], halo_hi[r]): one cell column either side of slab r
	uint64_t own_n = ~0ull;
	mps::DevBuf<unsigned long long> d_bounds; // device scratch of k_slab_bounds: [own_b | col_b | a | ok | halo_lo | halo_hi]
	bool slabs_set() const { return comm.on && own_n == n && own_b.size() == static_cast<size_t>(comm.nranks) + 1; }
	uint64_t nominal(int r) const { const uint64_t m = (n + comm.nranks - 1) / comm.nranks, b = static_cast<uint64_t>(r) * m; return b < n ? b : n; }
	uint64_t slab_begin(int r) const { return slabs_set() ? own_b[r] : nominal(r); }
	uint64_t own0() const { return comm.on ? slab_begin(comm.rank) : 0; }
	uint64_t own1() const { return comm.on ? slab_begin(comm.rank + 1) : n; }
};

namespace mps {

// every launch function returns cudaGetLastError() of its launches (no sync) --------------------------------------
// mps_grid.cu
cudaError_t launch_sort_and_search(mps_solver* s);      // cell keys -> counting sort -> reorder -> neighbour list
cudaError_t launch_get_cells(mps_solver* s, long long* d_cells /* n x dim, original order */);
// mps_gather.cu
cudaError_t launch_density(mps_solver* s, bool count_rows);
cudaError_t launch_ecs(mps_solver* s);
cudaError_t launch_explicit(mps_solver* s);
cudaError_t launch_save_x(mps_solver* s);
cudaError_t launch_ppe_fill(mps_solver* s);          // counts row lengths itself (stage-level API)
cudaError_t launch_ppe_fill_counted(mps_solver* s);  // row lengths already produced by launch_density(s, true)
cudaError_t launch_assign_pressure(mps_solver* s);
cudaError_t launch_gradient(mps_solver* s);
cudaError_t launch_ds(mps_solver* s);
cudaError_t launch_max_u2(mps_solver* s);               // -> d_sc->max_u2_bits
cudaError_t launch_set_dt(mps_solver* s, double dt, int advance, bool from_max_u);
cudaError_t launch_dndt_one(mps_solver* s, uint64_t orig_id, double* d_out);
cudaError_t launch_scatter_from_orig(mps_solver* s, const double* d_x, const double* d_u, const double* d_p, const double* d_n,
	const int32_t* d_type, uint64_t first, uint64_t count, bool append);
cudaError_t launch_gather_to_orig(mps_solver* s, double* d_x, double* d_u, double* d_p, double* d_n, int32_t* d_type);
cudaError_t launch_set_wall_group(mps_solver* s, uint64_t count, const uint64_t* d_ids, int group); // d_ids == nullptr: every non-fluid particle
cudaError_t launch_set_wall(mps_solver* s, uint64_t count, const uint64_t* d_ids, const double* d_x);
cudaError_t launch_gather_vec_to_orig(mps_solver* s, int which, double* d_out);
// mps_comm.cu
cudaError_t comm_allgather_state(mps_solver* s, bool pos, bool vel, bool prs, bool nden = false); // no-op without a communicator
cudaError_t comm_cg_solve(mps_solver* s);                                       // multi-rank CG (stepwise kernels + NCCL)
cudaError_t comm_ensure_arena(mps_solver* s, uint64_t rows, uint64_t mg_doubles = 0); // (re)allocates + exchanges the peer arenas; sets cg.z0 / cg.z1
cudaError_t comm_allreduce_sum(mps_solver* s, double* p, uint64_t count);       // in place, NCCL, on the solver's stream
double* comm_mg_section(mps_solver* s, int rank);                               // the mg section of `rank`'s arena as mapped here
cudaError_t comm_prepare_link(mps_solver* s);                                   // halo extents -> comm.link for the next persistent solve
void comm_release_peers(mps_solver* s);
// mps_observe.cu
cudaError_t launch_observe(mps_solver* s, const mps_observe_params* p, double* host_out); // 24 doubles, synchronises
// mps_scan.cu
cudaError_t launch_exclusive_scan_u32_to_u64(const uint32_t* in, uint64_t* out /* n + 1 */, uint64_t n, DevBuf<uint64_t>& tmp, cudaStream_t st,
	uint64_t* launches);
// mps_chunk.cu
cudaError_t launch_chunk_build(mps_solver* s);          // row_len, skey, cell_start -> chunk descriptors + row offsets
// mps_mg.cu
void mg_configure(mps_solver* s);                       // level geometry from the grid extents, MPS_CG_PRECOND / MPS_MG_* switches
cudaError_t launch_mg_rank0(mps_solver* s);             // during the sort: occupied cells -> compact ids
cudaError_t mg_ensure(mps_solver* s, uint64_t cells0);  // sizes every level for `cells0` occupied cells
cudaError_t launch_mg_setup(mps_solver* s);             // after k_ppe_fill: topology + Galerkin operators of every level
bool mg_wanted(const mps_solver* s);                    // the next assembly should build the hierarchy (as mg_active, before cg.external is reset)
bool mg_active(const mps_solver* s);                    // this solve is preconditioned (chunked, switch on, peer memory if several ranks)
// mps_cg.cu
cudaError_t cg_configure(mps_solver* s);                // picks chunk limits / pipeline depth for this environment
cudaError_t launch_cg(mps_solver* s);
cudaError_t launch_cg_step(mps_solver* s, int phase, CgStepScalars* st);    // one phase of the streaming CG as its own launch
cudaError_t launch_cg_reduce(mps_solver* s, double* dst);                   // per-CTA partials -> *dst, fixed order
cudaError_t launch_cg_scalars(mps_solver* s, int which, CgStepScalars* st); // 0 after r0, 1 after an iteration, 2 finish
cudaError_t cg_time_iteration(mps_solver* s, int reps, double* mean_ms, double* bytes);

inline unsigned int blocks_for(uint64_t n, unsigned int threads) { return static_cast<unsigned int>((n + threads - 1) / threads); }

} // namespace mps
