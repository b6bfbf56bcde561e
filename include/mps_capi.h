/* mps_capi.h — C ABI of libopenmps_b200.so: the B200-native (sm_100a, FP64) implementation of OpenMps's per-timestep
 * Moving-Particle-Semi-implicit hot path, Computer::ForwardTime() (reference: src/OpenMps/Computer.hpp:1700-1751).
 *
 * The reference has no FFI layer: its boundary is the C++ class template OpenMps::Computer<> (Computer.hpp:434-1790).
 * include/openmps/Computer.hpp in this repository is the drop-in for that class; every method of it forwards to one of
 * the entry points below ("replaces" = the reference interface the entry point stands in for).  Plain C types only,
 * caller-allocated buffers, no exceptions across the boundary, no torch types.
 *
 * Conventions
 *   - All particle arrays crossing this boundary are in the caller's ORIGINAL insertion order (Computer::Particles()
 *     order, Computer.hpp:1780-1783); vectors are row-major n x dim doubles.  Internally particles live cell-sorted.
 *   - Every function returns an mps_status; mps_last_error(h) gives the message.  Reference exceptions map to
 *     MPS_CG_NOT_CONVERGED (Computer::Exception, Computer.hpp:1424-1428) and MPS_CELL_OVERFLOW (Grid::Exception,
 *     Grid.hpp:314-318).  There is NO CPU fallback: without a CUDA device mps_create fails with MPS_CUDA_ERROR.
 *   - One host thread per handle (same contract as the reference object); work is queued on one CUDA stream per
 *     handle and the call returns once results the caller asked for are in the caller's buffers.
 */
#ifndef OPENMPS_B200_MPS_CAPI_H
#define OPENMPS_B200_MPS_CAPI_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mps_solver* mps_handle;

typedef enum mps_status
{
	MPS_OK = 0,
	MPS_CG_NOT_CONVERGED = 1, /* Computer::Exception("Conjugate Gradient method couldn't solve Pressure Poison Equation") */
	MPS_CELL_OVERFLOW = 2,    /* Grid::Exception("Too many particle in a block") */
	MPS_CUDA_ERROR = 3,
	MPS_NCCL_ERROR = 4,
	MPS_BAD_ARG = 5
} mps_status;

/* Particle::Type, Particle.hpp:16-29 */
enum { MPS_FLUID = 0, MPS_WALL = 1, MPS_DUMMY = 2, MPS_DISABLED = 3 };

/* Arguments of Environment's constructor (Environment.hpp:101-128) + the compile-time variant switches the reference
 * selects with DIM3 (defines.hpp:11) and CENTRAL_GRAVITY (Computer.hpp:925,981), which are run-time here. */
typedef struct mps_env
{
	int32_t dim;             /* 2 or 3 */
	int32_t central_gravity; /* 0 / 1 */
	double max_dt;           /* outputInterval / minStepCountPerOutput (Main.cpp:234) */
	double courant;
	double g;
	double rho;
	double nu;
	double r_e_by_l0;
	double l0;
	double min_x[3];         /* 2-D: {minX, minZ}; 3-D: {minX, minY, minZ} */
	double max_x[3];
} mps_env;

/* Derived constants, exactly as Environment computes them (Environment.hpp:129-216) + grid extents (Grid.hpp:137-150) */
typedef struct mps_env_info
{
	double t, dt, n0, max_dt, max_dx, r_e, neighbor_length, l0, rho, nu;
	int64_t grid_cells[3];
	uint64_t cell_capacity;  /* Grid::MaxParticles(), Grid.hpp:270-273 */
} mps_env_info;

typedef struct mps_stats
{
	uint64_t steps;            /* ForwardTime calls since creation */
	uint64_t cg_iterations;    /* total CG iterations since creation */
	uint64_t last_cg_iterations;
	double last_rr0, last_rr;  /* ||r0||^2 and final ||r||^2 of the last solve */
	uint64_t particles, neighbors, nnz, active_rows; /* sizes of the last step */
	uint64_t kernel_launches;  /* kernels launched by this library since creation */
	uint64_t disabled_last;    /* particles disabled (left the grid) by the last neighbour search, Computer.hpp:711-715 */
	uint64_t comm_calls;       /* NCCL collectives / send-recv groups issued since the last reset (0 on one GPU) */
	double cg_ms;              /* CUDA-event time of the CG kernel, summed over the steps since the last reset (always on) */
	double cg_bytes;           /* algorithmic bytes of those solves: sum of iterations x (12 nnz + 92 active rows), SURVEY 8d */
	double stage_ms[16];       /* accumulated CUDA-event time per stage when stage timing is on (see mps_stage_name) */
	uint64_t stage_calls[16];
	uint64_t matrix_sweeps;    /* passes over the assembled matrix (one per CG iteration + the initial residual), summed since the last reset */
	uint64_t mg_levels;        /* levels of the cell hierarchy the last solve's preconditioner used (0 = plain CG) */
	uint64_t mg_cells;         /* occupied cells of the neighbour grid = unknowns of the first coarse level, last step */
} mps_stats;

/* ---- life cycle --------------------------------------------------------------------------------------------------- */
/* replaces Computer::Computer(allowableResidual, env, posWall, posWallPre), Computer.hpp:1674-1691, and CreateComputer, :1800-1815 */
int mps_create(const mps_env* env, double eps, int device, mps_handle* out);
int mps_destroy(mps_handle h);
const char* mps_last_error(mps_handle h);      /* h may be NULL: error of the last failed mps_create */
int mps_get_env_info(mps_handle h, mps_env_info* out); /* replaces Computer::GetEnvironment(), Computer.hpp:1786-1789 */

/* ---- particles in / out ------------------------------------------------------------------------------------------- */
/* replaces Computer::AddParticles, Computer.hpp:1754-1777 (appends; may be called repeatedly).  Non-fluid particles are
 * pinned to the position they are added with, like the driver's positionWall (Main.cpp:304-315). */
int mps_add_particles(mps_handle h, uint64_t n, const double* x, const double* u, const double* p, const double* nd, const int32_t* type);
uint64_t mps_count(mps_handle h);
/* replaces Computer::Particles(), Computer.hpp:1780-1783.  Any pointer may be NULL. */
int mps_download(mps_handle h, double* x, double* u, double* p, double* nd, int32_t* type);
/* overwrite fields of the existing particles (what the gtest fixtures do through `computer->particles`); NULL = keep */
int mps_upload(mps_handle h, const double* x, const double* u, const double* p, const double* nd);
/* replaces the positionWall(i, t, dt) callback (Computer.hpp:1012-1019): target positions of the listed particles for
 * the coming steps.  The C++ drop-in evaluates the user's callable on the host and forwards the result here. */
int mps_set_wall_positions(mps_handle h, uint64_t n, const uint64_t* ids, const double* x);

/* positionWall(i, t, dt) evaluated ON THE DEVICE for analytic motions (pistons, shaking tanks): no host callback and no upload per
 * step.  The listed particles (ids == NULL: every non-fluid particle) follow, from the next step on,
 *     x(t) = base + velocity * tau + amplitude * (sin(omega * tau + phase) - sin(phase)),  tau = clamp(t - t_begin, 0, t_end - t_begin)
 * where base is the position they were added with (or last given through mps_set_wall_positions) and t is Environment::T() of the
 * step, exactly the t the reference hands to its callback (Computer.hpp:921,1015).  2-D: components are {x, z}.  Up to 8 motions per
 * handle; motion == NULL removes all of them. */
typedef struct mps_wall_motion
{
	double amplitude[3];
	double velocity[3];
	double omega, phase;
	double t_begin, t_end;
} mps_wall_motion;
int mps_set_wall_motion(mps_handle h, uint64_t n, const uint64_t* ids, const mps_wall_motion* motion);

/* ---- time stepping (the hot path) --------------------------------------------------------------------------------- */
int mps_determine_dt(mps_handle h, double* dt);           /* replaces Computer::DetermineDt, Computer.hpp:759-777 */
int mps_forward_time(mps_handle h, double dt);            /* replaces Computer::ForwardTime(dt), Computer.hpp:1700-1742 */
int mps_forward_time_auto(mps_handle h);                  /* replaces Computer::ForwardTime(),   Computer.hpp:1745-1751 */
/* the driver's inner loop `while (T() < nextOutputT) ForwardTime()` (Main.cpp:370-376) as one call: dt, t and the loop condition
 * are evaluated from device-resident scalars; each step still reads back two words (neighbour-list size, status) */
int mps_run_until(mps_handle h, double t_next, uint64_t* steps);
/* `steps` x ForwardTime() timed with CUDA events on the handle's stream (device time, milliseconds) */
int mps_run_steps(mps_handle h, uint64_t steps, double* device_ms);
int mps_get_time(mps_handle h, double* t, double* dt);    /* replaces Environment::T()/Dt(), Environment.hpp:230-242 */
int mps_set_dt(mps_handle h, double dt, int advance);     /* replaces env.Dt() = dt; env.SetNextT() (test fixtures)   */
/* the Computer constructor COPIES an Environment whose t / dt may already be set (every gtest fixture does) */
int mps_set_time(mps_handle h, double t, double dt);

/* ---- single stages (what the upstream gtests reach through `friend`, Computer.hpp:437-460) ------------------------ */
int mps_search_neighbor(mps_handle h);     /* Computer::SearchNeighbor,              Computer.hpp:698-756 + Grid.hpp */
int mps_compute_density(mps_handle h);     /* Computer::ComputeNeighborDensities,    Computer.hpp:780-833  */
int mps_error_correction(mps_handle h);    /* Computer::ComputeErrorCorrection,      Computer.hpp:877-910  */
int mps_explicit_forces(mps_handle h);     /* Computer::ComputeExplicitForces,       Computer.hpp:914-1021 */
int mps_save_x(mps_handle h);              /* Computer::SaveX,                       Computer.hpp:1025-1039 */
int mps_set_ppe(mps_handle h);             /* Computer::SetPressurePoissonEquation,  Computer.hpp:1145-1356 */
int mps_solve_ppe(mps_handle h);           /* Computer::SolvePressurePoissonEquation,Computer.hpp:1359-1429 */
int mps_assign_pressure(mps_handle h);     /* P = max(x, 0),                         Computer.hpp:1076-1097 */
int mps_implicit_forces(mps_handle h);     /* Computer::ComputeImplicitForces,       Computer.hpp:1043-1102 */
int mps_pressure_gradient(mps_handle h);   /* Computer::ModifyByPressureGradient,    Computer.hpp:1433-1564 */
int mps_dynamic_stabilize(mps_handle h);   /* Computer::DynamicStabilize,            Computer.hpp:1568-1656 */
int mps_dndt(mps_handle h, uint64_t i, double* out); /* Computer::NeighborDensityVariationSpeed(i), Computer.hpp:838-872 */

/* ---- inspection (parity tests; original particle ids) ------------------------------------------------------------- */
int mps_get_cells(mps_handle h, int64_t* cells /* n x dim */);              /* Grid::Block<AXIS>, Grid.hpp:250-254 */
/* Computer::NeighborCount / Neighbor, Computer.hpp:594-612: rowptr has n+1 entries; idx may be NULL to size the call */
int mps_get_neighbors(mps_handle h, uint64_t* rowptr, uint64_t* idx);
int mps_get_csr_nnz(mps_handle h, uint64_t* nnz);
/* ppe.A in the reference's layout: all n rows (identity rows for Dummy/Disabled), columns ascending (Computer.hpp:1337-1354) */
int mps_get_csr(mps_handle h, uint64_t* rowptr, uint32_t* col, double* val);
/* which: 0 ppe.x, 1 ppe.b, 2 cg.r, 3 cg.p, 4 cg.Ap, 5 ecs, 6 nWithoutSpp (n doubles); 7 du, 8 originalX (n x dim) */
int mps_get_vec(mps_handle h, int which, double* out);
/* load an arbitrary CSR system into ppe.{A,b,x} (the CG known-answer tests, test_ComputerConjugateGradient.cpp:93-107) */
int mps_set_system(mps_handle h, uint64_t n, const uint64_t* rowptr, const uint32_t* col, const double* val, const double* b, const double* x0);
int mps_get_solution(mps_handle h, uint64_t n, double* x);
/* The multigrid preconditioner of the PPE solve (no counterpart in the reference, whose CG is unpreconditioned): tables of the
 * cell hierarchy after the last assembly, for the tests that check them against P^T A P.  level >= 0: which = 0 cell key (u32),
 * 1 neighbour ids (u32 x 3^D), 2 children (u32 x 2^D), 3 parent (u32), 4 Galerkin stencil (f64 x 3^D), 5 omega / centre,
 * 6 r, 7 e0, 8 e1 (f64); level = -1: 0 row -> cell (u32 x n), 1 first row of every cell (u64 x cells + 1), 2 1 / a_ii (f64 x n),
 * 3 slot -> original particle id (u32 x n).  *count = elements available; at most capacity_bytes are copied to out. */
int mps_debug_mg(mps_handle h, int level, int which, void* out, uint64_t capacity_bytes, uint64_t* count);

/* ---- benchmark observables as device reductions ------------------------------------------------------------------------
 * The reference's plotting scripts re-read result/particles_%05d.csv to get these; here ONE pass over the resident state on the
 * GPU returns them (24 doubles instead of a 52 B/particle download), so that long-run physical parity is an assertion:
 *   edge_x                      Benchmark/DamBreak/koshizukaoka1996_edge.py:14-20   max x over Type 0 (leading edge of the collapse)
 *   h1, h2, p2_sum / p2_count   Benchmark/DamBreak/zhouetal1999.py:26-39            water height at x_h1 / x_h2 (highest particle with
 *                                                                                  n > min_n within l0/2), mean p of Wall particles with x < 0
 *                                                                                  within d/2 of z_p2
 *   r_min/r_max_surface, center_p  Benchmark/CentralGravity/check_result.py:26-57   extremes of |x| over particles with n < surface_n
 *                                                                                  (= beta n0), p of the particle nearest the origin
 *   inner, sum_d, sum_p, sum_dd, sum_dp, max_dev   hydrostatic column (Benchmark/StaticPressure): moments of (depth = surface_z - z, p)
 *                                                                                  over Fluid with p > 0 and max |p - rho_g depth|
 * Counts are exact integers carried as doubles.  Fields whose set is empty hold -DBL_MAX (maxima) / DBL_MAX (minima). */
typedef struct mps_observe_params
{
	double surface_n;   /* central gravity: n below this marks a surface particle */
	double rho_g;       /* hydrostatic reference gradient rho * g */
	double surface_z;   /* z of the free surface (depth = surface_z - z); use top_z of an earlier call if unknown */
	double x_h1, x_h2, min_n, z_p2, d; /* Zhou et al. probe stations */
} mps_observe_params;
typedef struct mps_observables
{
	double edge_x, top_z;
	double r_max_surface, r_min_surface;
	double center_r2, center_id, center_p;
	double inner, sum_d, sum_p, sum_dd, sum_dp, max_dev;
	double h1, h2, p2_sum, p2_count;
	double fluid, wall, dummy, disabled;
	double p_max, u_max2, surface_count;
} mps_observables;
int mps_observe(mps_handle h, const mps_observe_params* params, mps_observables* out);

/* ---- multi-GPU (no counterpart in the reference, which is single-process OpenMP) ------------------------------------
 * One process per GPU.  Rank 0 obtains an NCCL unique id (128 bytes), the launcher distributes it (torch.distributed, MPI,
 * a file ...), every rank attaches it to its handle BEFORE the first step and then adds the SAME particles in the same
 * order.  The particle state is replicated; each rank computes the x-slab [own_first, own_last) of the cell-sorted slots
 * (neighbour lists, gather stages, PPE rows, CG rows); see openmps_b200/csrc/mps_comm.cu.  NCCL carries the set-up (the unique id, the
 * CUDA IPC handles) and two small collectives per step (coarse operator sum, halo extents); halo gathers, the solve's halo and its
 * dot products run over peer memory (mps_comm_mode). */
int mps_comm_unique_id(void* out128);
int mps_comm_init(mps_handle h, int rank, int nranks, const void* id128);
int mps_comm_info(mps_handle h, int* rank, int* nranks, uint64_t* own_first, uint64_t* own_last);
/* how the ranks' CG solves are coupled: 0 not decided yet (before the first assembly) / single GPU, 1 peer memory over NVLink
 * (ONE persistent kernel per solve on every rank; rim rows and dot products cross GPUs by P2P stores, mps_cg.cu), 2 NCCL between
 * per-phase launches (chosen when the ranks cannot map each other's memory, or with MPS_COMM_NCCL_ONLY=1) */
int mps_comm_mode(mps_handle h, int* mode);
/* the slab arithmetic on its own (host only, needs no GPU): slots [first, last) of `rank` among `nranks` for n particles */
int mps_partition_range(uint64_t n, int nranks, int rank, uint64_t* first, uint64_t* last);

/* ---- measurement ---------------------------------------------------------------------------------------------------- */
int mps_set_stage_timing(mps_handle h, int on);   /* CUDA-event timing per stage (adds a sync per stage; off by default) */
int mps_get_stats(mps_handle h, mps_stats* out);
int mps_reset_stats(mps_handle h);
const char* mps_stage_name(int stage);            /* names of mps_stats.stage_ms slots; NULL past the end */
/* time `reps` launches of one named kernel on the current state without changing it (roofline measurements):
 * "density", "cg_iteration" (one SpMV + both vector updates), ...; returns mean device milliseconds per launch */
int mps_time_kernel(mps_handle h, const char* name, int reps, double* mean_ms, double* algorithmic_bytes);
int mps_flush_l2(mps_handle h);                   /* writes a buffer larger than L2 (timing hygiene) */
/* In-kernel cycle counters of the streaming CG kernel (development aid; off by default).  out[0..7] = mean over CTAs of
 * {phase-1 cycles, consumer wait-for-data cycles, phase-2 cycles, grid-barrier cycles, producer wait-for-free-stage cycles,
 *  chunks processed, iteration cycles, 0} of the LAST solve; out[8..15] = the maxima; out[16] = chunks, out[17] = blob bytes,
 *  out[18] = CTAs. */
int mps_set_cg_profile(mps_handle h, int on);
int mps_get_cg_profile(mps_handle h, double* out /* 19 doubles */);
int mps_get_cg_profile_raw(mps_handle h, uint64_t* out /* 8 per CTA */, uint64_t capacity_ctas, uint64_t* ctas);
/* preconditioned solve, cycles summed over the last solve as CTA 0 sees them: [l] restriction to level l + 1 (barrier included),
 * [16] top-level sweeps, [17] wait for the hand-back barrier, [20 + l] prolongation + smoothing of level l, [40] row pass 2a,
 * [41] row pass 2b, [42] r.r reduction, [43] p.Ap reduction, [44] r.z reduction */
int mps_get_cg_profile_stages(mps_handle h, uint64_t* out /* 64 */);

/* ---- environment variables read by the library (tuning and tests; none is needed in normal use) ---------------------------
 *   MPS_CG_PRECOND=0       solve the PPE with the reference's plain CG (Computer.hpp:1359-1429) instead of the multigrid-preconditioned
 *                          CG (default; same system, same stopping rule, ~30x fewer iterations; csrc/mps_mg.cu)
 *   MPS_MG_OMEGA, MPS_MG_GAMMA, MPS_MG_TOP_SWEEPS, MPS_MG_TOP_CELLS   Jacobi damping (0.8), over-correction (1.8), extra sweeps (4) on
 *                          the top level and its size limit (64 cells) of the preconditioner's V-cycle
 *   MPS_CG_ADAPTIVE=0      freeze the CG kernel's CTA split (uniform): runs become bit-identical; default: re-balanced every solve
 *   MPS_COMM_NCCL_ONLY=1   several GPUs: couple the ranks through NCCL between per-phase launches instead of peer memory
 *   MPS_CG_WARPS, MPS_CG_LPR, MPS_CG_STAGES, MPS_CG_PRODUCERS, MPS_CG_COST_FIXED   consumer warps / lanes per row / ring depth /
 *                          producer warps (1..4) / load-balance model of the streaming CG kernels
 *   MPS_MG_DIST_CELLS      several GPUs: level 0 of the cell hierarchy is distributed when it has more cells than this (150000)
 *   MPS_SLAB_WEIGHTS=f,w,d several GPUs: work per Fluid / Wall / Dummy particle in the slab split (8,4,1)
 *   MPS_ALLOC_SLACK=0      no head-room on growing buffers and exact blob sizing (blocks that only just fit one GPU; default 25 %)
 *   MPS_CG_GENERIC=1       solve assembled systems with the generic CSR kernel (k_cg_solve) instead of the streaming one
 */

#ifdef __cplusplus
}
#endif
#endif
