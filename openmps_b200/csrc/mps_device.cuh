// mps_device.cuh — device-side types shared by the kernels of libopenmps_b200.so (sm_100a, FP64).
//
// Data layout in HBM (DESIGN.md "layout"): particles live in CELL-SORTED slot order, rebuilt at every neighbour search
// (key = linear cell id, x-major ... z-minor like the reference's data[i][j][k], Grid.hpp:76,140-150; ties broken by
// ascending original id, which is the reference's insertion order).  Per-slot SoA arrays:
//   pos, vel : Vec<D>  (2-D: 16 B aligned double2; 3-D: 32 B aligned, padded to 4 doubles so that one gather = one sector)
//   prs, nden, nws (nWithoutSpp), ecs : double        type : u8        orig : u32 (slot -> original id)
// plus inv (original id -> slot), the per-step neighbour list (u64 row pointers + u32 slot indices) and the PPE CSR
// (u64 row pointers, u32 slot columns, f64 values).
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace mps {

enum : uint8_t { kFluid = 0, kWall = 1, kDummy = 2, kDisabled = 3 }; // Particle.hpp:16-29

template<int D> struct Vec;
template<> struct alignas(16) Vec<2> { double v[2]; };
template<> struct alignas(32) Vec<3> { double v[4]; }; // v[3] is padding (always 0)

template<int D> __host__ __device__ inline Vec<D> vzero()
{
	Vec<D> r;
	for (int k = 0; k < (D == 2 ? 2 : 4); k++) r.v[k] = 0.0;
	return r;
}

// Constants of one run: Environment (Environment.hpp:129-216) + Grid extents (Grid.hpp:137-150).  Passed by value.
struct EnvConst
{
	int dim;
	int central_gravity;
	double n0, max_dt, max_dx, l0, r_e, r_e2, neighbor_length, rho, nu, eps;
	double nl2_lim;   // the smallest double whose (correctly rounded) square root is >= neighbor_length:
	                  // sqrt(r2) < neighbor_length  <=>  r2 < nl2_lim  for every r2 (sqrt is monotonic), so the search needs no sqrt
	double g[3];      // Environment::G
	double min_x[3];
	double max_x[3];
	long long grid_n[3];       // cells per axis
	unsigned long long ncells; // product; key == ncells marks "not in the grid" (Disabled)
	unsigned int cell_cap;     // Grid::MaxParticles()
	// pair coefficients whose prefix does not depend on the pair, evaluated on the host in the reference's order:
	double visc_coef;  // nu * (5 - DIM) * r_e / n0        (Computer.hpp:961)
	double ppe_coef;   // (5 - DIM) * r_e / n0             (Computer.hpp:1291)
	double ds_d2;      // (L_0 - MaxDx)^2                  (Computer.hpp:1572,1586)
};

// Analytic wall motions evaluated on the device (mps_set_wall_motion; the reference evaluates a host callback positionWall(i, t, dt)
// per non-fluid particle and step, Computer.hpp:993,1012-1019).  A particle of group g follows
//   x(t) = base + vel * tau + amp * (sin(omega * tau + phase) - sin(phase)),   tau = clamp(t - t0, 0, t1 - t0)
// with base = its wall[] entry (the position it was added with / last given by mps_set_wall_positions).
constexpr int kMaxWallMotions = 8;
struct WallMotion { double amp[3]; double vel[3]; double omega, phase, t0, t1; };
struct WallMotions { int count; int pad_; WallMotion m[kMaxWallMotions]; };

// Scalars that live on the device so that a step needs no host round trip.
struct DevScalars
{
	double t, dt;
	unsigned long long max_u2_bits; // max_i |u_i|^2 as ordered bits (non-negative doubles compare like integers)
	unsigned long long nbr_total;   // entries in the neighbour list of the last search
	unsigned long long nnz_total;   // entries in the CSR of the last assembly
	unsigned long long cg_iterations;
	double rr0, rr;                 // ||r0||^2, final ||r||^2
	int cg_converged;
	int z_final;                    // which of the streaming kernel's {r, p} buffers holds the final residual / direction
	int error;                      // sticky mps_status raised on the device (cell overflow, CG failure)
	unsigned int disabled_now;      // particles disabled by the last search
	unsigned long long active_rows; // PPE rows of the last assembly that are not Dummy / Disabled
	unsigned long long n_chunks;    // chunks of the last assembly
	unsigned long long n_live;      // ... of which have matrix entries (the list the CG kernel walks)
	unsigned long long blob_total;  // bytes of all chunk blobs of the last assembly
	unsigned long long cost_total;  // sum of the chunk cost model (load balance of the CG kernel)
	unsigned long long grid_barrier; // arrival counter of the CG kernel's grid barrier (zeroed before each launch)
	unsigned int mg_levels;         // levels of the cell hierarchy the last preconditioned solve used (0: plain CG)
	unsigned int pad0_;
	unsigned long long mg_cells0;   // occupied cells of the neighbour grid at the last sort
};

// ---- PPE in "chunk blob" form (what the CG kernel streams, DESIGN.md "CG kernel") ---------------------------------------
// Rows (slots) are cut into chunks of consecutive rows.  A chunk owns one contiguous, 16-byte aligned byte blob
//   [ copy of the chunk's descriptor : 128 B ][ val : nnz_pad x f64 ][ lcol : nnz_pad x u16 ][ rowoff : rows_pad x u16 ]
//                                                            nnz_pad = round_up(nnz, 8), rows_pad = round_up(rows + 1, 8)
// so that one bulk async copy brings the whole chunk into shared memory.  Columns are 16-bit indices into the chunk's
// WINDOW: the union of the <= 3 (2-D) / 9 (3-D) contiguous slot ranges that hold every neighbour cell of the chunk's rows
// (slots are cell-sorted, so the three z-neighbour cells of one (x[,y]) column are one contiguous range).  The CG kernel
// stages the window of the gathered vectors in shared memory with bulk copies as well.
constexpr int kMaxRanges = 9;
struct alignas(16) ChunkDesc
{
	uint32_t row_begin;               // first row (slot) of the chunk
	uint32_t rows;                    // consecutive rows
	uint32_t nnz;                     // matrix entries of those rows (un-padded)
	uint32_t nranges;                 // window ranges in use
	uint64_t blob_off;                // byte offset of the blob (multiple of 16)
	uint32_t blob_bytes;              // multiple of 16
	uint32_t window;                  // window entries = sum of range_len (each even)
	uint32_t range_start[kMaxRanges]; // first staged slot of each range (even => 16-byte aligned doubles)
	uint16_t range_len[kMaxRanges];   // staged entries (even)
	uint16_t range_off[kMaxRanges];   // position of the range inside the window
	int32_t self_off;                 // window position of row_begin: an ACTIVE row r of the chunk sits at self_off + (r - row_begin)
	uint64_t cost_off;                // exclusive prefix of the chunk cost model (ChunkLimits::cost_*)
	uint64_t pad_;
};
static_assert(sizeof(ChunkDesc) == 128, "one descriptor = one 128-byte bulk copy into the stage");
static_assert(sizeof(ChunkDesc) % 16 == 0, "descriptors are read with 16-byte loads");

struct ChunkLimits
{
	uint32_t max_rows;   // <= 256 (one builder thread partitions one block of 256 rows); = consumer threads / lanes per row
	uint32_t max_nnz;    // entries per chunk (bounds the blob stage in shared memory; < 65536 for u16 row offsets)
	uint32_t max_window; // window entries per chunk (bounds the vector stages in shared memory)
	uint32_t cost_fixed, cost_per_nnz; // load-balance model of one chunk with entries: cost_fixed + cost_per_nnz x entries
};

__host__ __device__ inline uint32_t round_up8(uint32_t v) { return (v + 7u) & ~7u; }
constexpr uint32_t kBlobHeader = 128; // the descriptor travels with the blob: one bulk copy per chunk (a bulk copy costs ~420 cycles of
                                      // the SM's copy engine whatever its size, tools/tma_bench.cu)
__host__ __device__ inline uint32_t chunk_blob_bytes(uint32_t rows, uint32_t nnz) { return kBlobHeader + round_up8(nnz) * 10u + round_up8(rows + 1u) * 2u; }

// scalars of the CG recurrence when the solve runs as one launch per phase (multi-GPU, mps_comm.cu)
struct CgStepScalars
{
	double rr, pAp, rr_new, beta, tol, rr0;
	unsigned long long iter;
	int converged;
	int zcur_is_1;
};

// ---- multi-GPU persistent CG over peer memory (mps_cg.cu k_cg_stream<LPR, true>, mps_comm.cu) ---------------------------
constexpr int kMaxPeerRanks = 8;
struct alignas(16) PeerMail { double value; unsigned long long flag; };  // one rank's contribution to one reduction
// layout of one rank's shared arena (one cudaMalloc, exported with cudaIpcGetMemHandle):
//   [mail 4 x kMaxPeerRanks][total 4][barrier flags kMaxPeerRanks] ... padded to kPeerHeaderBytes ... [z0][z1]
constexpr size_t kPeerMailBytes = 4 * kMaxPeerRanks * sizeof(PeerMail);
constexpr size_t kPeerBarrierOff = kPeerMailBytes + 4 * sizeof(PeerMail); // arrival counters of the stream-level barrier (mps_comm.cu)
constexpr size_t kPeerHeaderBytes = 1024;
static_assert(kPeerBarrierOff + kMaxPeerRanks * sizeof(unsigned long long) <= kPeerHeaderBytes, "arena header");
struct PeerLink
{
	int rank, nranks;
	unsigned long long tag;          // solve number << 32: flags of earlier solves never match
	double2* nb_z0[2];               // [0] left, [1] right neighbour's z0 (peer-mapped), nullptr at the ends of the chain
	double2* nb_z1[2];
	PeerMail* mail[kMaxPeerRanks];   // every rank's mailbox (peer-mapped; mail[rank] is local)
};

// ---- multigrid preconditioner on the cell hierarchy (mps_mg.cu builds it, mps_cg.cu k_pcg_stream applies it) --------------
constexpr uint32_t kMgNone = 0xffffffffu;
constexpr int kMgMaxLevels = 16;
constexpr int kMgMaxDistLevels = 1;  // multi-GPU: level 0 of the cell hierarchy may be distributed over the ranks (MgDist)
struct MgLevelPtrs
{
	const uint64_t* count;   // occupied cells of this level (device value: the total of the level's rank scan)
	const double* S;         // [cells][3^D] Galerkin stencil, centre at 3^D / 2
	const uint32_t* nbr;     // [cells][3^D] compact id of the neighbour cell or kMgNone
	const double* dinv;      // [cells] omega / centre entry (0 where the centre is 0)
	const uint32_t* child;   // [cells][2^D] compact ids one level down (levels >= 1)
	const uint32_t* parent;  // [cells] compact id one level up (all but the last level)
	double* r;               // [cells] restricted residual
	double* e0;              // [cells] correction, two buffers (Jacobi sweeps ping-pong)
	double* e1;
	double* part;            // several ranks, level 1 only: this rank's share of the restriction from the distributed level 0
	// multi-GPU: dense index -> compact id (the level's rank scan, dense + 1 entries), cells per x column, and where the level's
	// vectors sit inside the "mg" section of every rank's peer arena (doubles)
	const uint64_t* ranktab;
	uint64_t colstride, dense;
	uint64_t off_r, off_e0, off_e1, off_part;
};
// Several ranks (mps_comm.cu, DESIGN.md "multi-GPU"): slabs are cut between cell COLUMNS (all cells of one x index), so every
// level-0 cell belongs to exactly one rank and a rank's cells are one contiguous range of compact ids.  k = 1: level 0 is
// DISTRIBUTED — a rank smooths / restricts / prolongs its own cells and reads the neighbour ranks' halo cells straight from their
// arenas over NVLink — and levels >= 1 are replicated (every rank runs them whole, on identical data, with identical results);
// k = 0 (few cells): level 0 is gathered from all ranks and the whole cycle is replicated.  mps_cg.cu mg_vcycle.
struct MgDist
{
	int on;                                // 0 on one GPU
	int k;
	int rank, nranks;
	uint32_t col_b[kMaxPeerRanks + 1];     // first level-0 cell column of every rank's slab
	double* peer_vec[kMaxPeerRanks];       // the mg section of every rank's arena ([rank] is local)
	uint64_t halo_lo, halo_hi;             // slots [halo_lo, own0) and [own1, halo_hi): one cell column of the adjacent ranks (what this rank's matrix windows reach)
};
struct MgArgs
{
	int levels;              // levels built (the kernel stops at the first one with <= top_cells cells)
	int top_sweeps;          // extra damped-Jacobi sweeps on the top level
	uint32_t top_cells;
	uint32_t small_cells;    // levels with at most this many cells are run by CTA 0 alone between block barriers
	double gamma;            // over-correction of every coarse-grid correction
	const uint32_t* crow;    // [rows] level-0 cell of a row (kMgNone for Disabled rows)
	const uint64_t* cstart;  // [cells0 + 1] first row of every level-0 cell; last entry = rows that lie in cells
	const double* dinv0;     // [rows] 1 / a_ii (0 for rows without entries)
	double* r;               // [rows] residual
	MgLevelPtrs lv[kMgMaxLevels];
	MgDist dist;
};

template<int D>
struct Particles
{
	Vec<D>* pos;
	Vec<D>* vel;
	double* prs;
	double* nden;
	uint8_t* type;
	uint32_t* orig;
};

} // namespace mps
