// mps_capi.cu — the C ABI of libopenmps_b200.so (include/mps_capi.h): handle, host orchestration of one MPS time step
// (Computer::ForwardTime, Computer.hpp:1700-1751), state transfer, inspection for the parity tests, timing.
// No CPU fallback: every entry point needs a CUDA device.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <memory>
#include <new>
#include <utility>

#include "mps_solver.h"

using namespace mps;

namespace {

std::string g_create_error;

const char* kStageNames[kStCount] = { "sort", "search", "density", "ecs", "explicit", "ppe_assemble", "cg", "pressure",
	"gradient", "ds", "dt" };

int fail(mps_solver* s, int code, const std::string& msg)
{
	if (s) s->last_error = msg; else g_create_error = msg;
	return code;
}
extern "C" void mps_comm_release(mps_handle s);

int cuda_fail(mps_solver* s, cudaError_t e, const char* where)
{
	if (s && !s->comm_error.empty())
	{
		const std::string msg = std::string(where) + ": " + s->comm_error;
		s->comm_error.clear();
		return fail(s, MPS_NCCL_ERROR, msg);
	}
	return fail(s, MPS_CUDA_ERROR, std::string(where) + ": " + cudaGetErrorString(e));
}

#define CU(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) return cuda_fail(s, e_, #expr); } while (0)
#define NEED(h) do { if (!(h)) return MPS_BAD_ARG; } while (0)

// sticky device-side errors (cell overflow, CG failure) -> status + the reference's exception text
int device_status(mps_solver* s)
{
	const int e = s->h_sc->error;
	if (e == MPS_OK) return MPS_OK;
	if (e == MPS_CELL_OVERFLOW) return fail(s, e, "Too many particle in a block");                                  // Grid.hpp:316
	if (e == MPS_CG_NOT_CONVERGED) return fail(s, e, "Conjugate Gradient method couldn't solve Pressure Poison Equation"); // Computer.hpp:1427
	return fail(s, e, "device error");
}

int clear_device_error(mps_solver* s);

// The reference's exceptions are per call (a caller may catch one and carry on), so the device flag is cleared once reported.
int report_device_status(mps_solver* s)
{
	const int rc = device_status(s);
	if (rc != MPS_OK) clear_device_error(s);
	return rc;
}

int sync_scalars(mps_solver* s)
{
	CU(cudaMemcpyAsync(s->h_sc, s->d_sc, sizeof(DevScalars), cudaMemcpyDeviceToHost, s->stream));
	CU(cudaStreamSynchronize(s->stream));
	return MPS_OK;
}
int clear_device_error(mps_solver* s)
{
	CU(cudaMemsetAsync(&s->d_sc->error, 0, sizeof(int), s->stream));
	s->h_sc->error = 0;
	s->sort_error = 0;
	return MPS_OK;
}

struct StageTimer
{
	mps_solver* s; int stage;
	StageTimer(mps_solver* s_, int st) : s(s_), stage(st) { if (s->stage_timing) cudaEventRecord(s->ev0, s->stream); }
	~StageTimer()
	{
		if (!s->stage_timing) return;
		cudaEventRecord(s->ev1, s->stream);
		cudaEventSynchronize(s->ev1);
		float ms = 0; cudaEventElapsedTime(&ms, s->ev0, s->ev1);
		s->stats.stage_ms[stage] += ms;
		s->stats.stage_calls[stage] += 1;
	}
};

int ensure_particle_capacity(mps_solver* s, uint64_t n_exact)
{
	// slack: in-place all-gathers move nranks x ceil(n / nranks) elements
	const uint64_t n = n_exact + 64;
	cudaStream_t st = s->stream;
	const uint64_t old = s->n;
	const int vs = s->vec_stride();
	for (int b = 0; b < 2; b++)
	{
		const uint64_t keep = (b == s->cur) ? old : 0;
		CU(s->pos[b].ensure(n * vs, st, keep * vs)); CU(s->vel[b].ensure(n * vs, st, keep * vs));
		CU(s->prs[b].ensure(n, st, keep)); CU(s->nden[b].ensure(n, st, keep));
		CU(s->type[b].ensure(n, st, keep)); CU(s->orig[b].ensure(n, st, keep));
	}
	CU(s->inv.ensure(n, st, old));
	CU(s->wall.ensure(n * vs, st, old * vs));
	CU(s->wall_group.ensure(n, st, old));
	CU(s->nws.ensure(n, st, old)); CU(s->ecs.ensure(n, st, old));
	CU(s->du.ensure(n * vs, st, old * vs)); CU(s->x0.ensure(n * vs, st, old * vs));
	return MPS_OK;
}

// host staging -> device staging
template<typename T>
int to_device(mps_solver* s, DevBuf<T>& buf, size_t offset, const T* host, size_t count)
{
	CU(cudaMemcpyAsync(buf.p + offset, host, count * sizeof(T), cudaMemcpyHostToDevice, s->stream));
	return MPS_OK;
}

int set_dt_device(mps_solver* s, double dt, int advance, bool from_max_u)
{
	CU(launch_set_dt(s, dt, advance, from_max_u));
	return MPS_OK;
}

// One reference time step after dt has been set and t advanced: Computer.hpp:1708-1741
int step_stages(mps_solver* s);
int step_body(mps_solver* s)
{
	// several ranks: the stages exchange halos only; one full gather of the state ends the step (mps_comm.cu comm_allgather_state)
	s->comm.halo_step = s->comm.on;
	const int rc = step_stages(s);
	const bool gather_all = s->comm.halo_step && rc == MPS_OK && s->sort_error != MPS_CELL_OVERFLOW;
	s->comm.halo_step = false;
	if (gather_all) { StageTimer t(s, kStDs); CU(comm_allgather_state(s, true, true, true, true)); }
	return rc;
}
int step_stages(mps_solver* s)
{
	{ StageTimer t(s, kStSearch); CU(launch_sort_and_search(s)); }
	if (s->sort_error == MPS_CELL_OVERFLOW)
	{
		// Grid::Store throws out of SearchNeighbor (Grid.hpp:311-318, Computer.hpp:705-717): nothing after the sort runs
		s->h_sc->error = MPS_CELL_OVERFLOW;
		return report_device_status(s);
	}
	{ StageTimer t(s, kStDensity); CU(launch_density(s, false)); }
	{ StageTimer t(s, kStEcs); CU(launch_ecs(s)); }
	{ StageTimer t(s, kStExplicit); CU(launch_explicit(s)); }
	{ StageTimer t(s, kStDensity); CU(launch_density(s, true)); }   // + SaveX + PPE row lengths
	{ StageTimer t(s, kStPpeAssemble); CU(launch_ppe_fill_counted(s)); }
	{
		StageTimer t(s, kStCg);
		CU(cudaEventRecord(s->ev_cg0, s->stream));
		CU(launch_cg(s));
		CU(cudaEventRecord(s->ev_cg1, s->stream));
	}
	{ StageTimer t(s, kStPressure); CU(launch_assign_pressure(s)); }
	{ StageTimer t(s, kStGradient); CU(launch_gradient(s)); }
	{ StageTimer t(s, kStDs); CU(launch_ds(s)); }
	s->stats.steps += 1;
	return MPS_OK;
}

int finish_step(mps_solver* s)
{
	int rc = sync_scalars(s);
	if (rc != MPS_OK) return rc;
	s->stats.last_cg_iterations = s->h_sc->cg_iterations;
	s->stats.cg_iterations += s->h_sc->cg_iterations;
	s->stats.last_rr0 = s->h_sc->rr0; s->stats.last_rr = s->h_sc->rr;
	s->stats.particles = s->n; s->stats.neighbors = s->nbr_total; s->stats.nnz = s->h_sc->nnz_total;
	s->nnz_total = s->h_sc->nnz_total;
	s->stats.active_rows = s->h_sc->active_rows;
	s->stats.disabled_last = s->h_sc->disabled_now;
	if (s->n)
	{
		// the stream is idle here (sync_scalars), so both events have completed
		float ms = 0;
		if (cudaEventElapsedTime(&ms, s->ev_cg0, s->ev_cg1) == cudaSuccess) s->stats.cg_ms += ms;
		s->stats.cg_bytes += static_cast<double>(s->h_sc->cg_iterations) *
			(12.0 * static_cast<double>(s->h_sc->nnz_total) + 92.0 * static_cast<double>(s->h_sc->active_rows));
		s->stats.matrix_sweeps += s->h_sc->cg_iterations + 1; // one SpMV per iteration + the initial residual
	}
	s->stats.mg_levels = mg_active(s) ? s->h_sc->mg_levels : 0;
	s->stats.mg_cells = mg_active(s) ? s->mg.cells0 : 0;
	return report_device_status(s);
}

} // namespace

extern "C" {

const char* mps_stage_name(int stage) { return (stage >= 0 && stage < kStCount) ? kStageNames[stage] : nullptr; }

const char* mps_last_error(mps_handle h) { return h ? h->last_error.c_str() : g_create_error.c_str(); }

int mps_create(const mps_env* env, double eps, int device, mps_handle* out)
{
	if (!env || !out) return fail(nullptr, MPS_BAD_ARG, "null argument");
	if (env->dim != 2 && env->dim != 3) return fail(nullptr, MPS_BAD_ARG, "dim must be 2 or 3");
	*out = nullptr;
	int count = 0;
	cudaError_t e = cudaGetDeviceCount(&count);
	if (e != cudaSuccess || count == 0)
		return fail(nullptr, MPS_CUDA_ERROR, std::string("no CUDA device (this library has no CPU path): ") + cudaGetErrorString(e));
	if (device < 0 || device >= count) return fail(nullptr, MPS_BAD_ARG, "bad device index");
	// every error return below gives back what was created so far (stream, events, the scalar blocks), not only the struct
	auto discard = [](mps_solver* p)
	{
		if (!p) return;
		if (p->ev0) cudaEventDestroy(p->ev0);
		if (p->ev1) cudaEventDestroy(p->ev1);
		if (p->ev_cg0) cudaEventDestroy(p->ev_cg0);
		if (p->ev_cg1) cudaEventDestroy(p->ev_cg1);
		if (p->stream) cudaStreamDestroy(p->stream);
		if (p->d_sc) cudaFree(p->d_sc);
		if (p->h_sc) cudaFreeHost(p->h_sc);
		delete p;
	};
	std::unique_ptr<mps_solver, decltype(discard)> sp(new (std::nothrow) mps_solver(), discard);
	mps_solver* s = sp.get();
	if (!s) return fail(nullptr, MPS_BAD_ARG, "out of host memory");
	s->device = device;
	if ((e = cudaSetDevice(device)) != cudaSuccess) return cuda_fail(nullptr, e, "cudaSetDevice");
	cudaDeviceProp prop;
	if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return cuda_fail(nullptr, e, "cudaGetDeviceProperties");
	s->sm_count = prop.multiProcessorCount;
	if (!prop.cooperativeLaunch) return fail(nullptr, MPS_CUDA_ERROR, "device lacks cooperative launch");
	if ((e = cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking)) != cudaSuccess) return cuda_fail(nullptr, e, "cudaStreamCreate");
	cudaEventCreate(&s->ev0); cudaEventCreate(&s->ev1); cudaEventCreate(&s->ev_cg0); cudaEventCreate(&s->ev_cg1);

	// ---- Environment, in the reference's evaluation order (Environment.hpp:129-216) ----
	EnvConst& c = s->env;
	const int D = env->dim;
	const double l_0 = env->l0, r_eByl_0 = env->r_e_by_l0, courant = env->courant;
	c.dim = D; c.central_gravity = env->central_gravity ? 1 : 0;
	c.max_dt = std::min(env->max_dt, std::sqrt(2 * (courant * l_0) / env->g));
	c.max_dx = courant * l_0;
	c.l0 = l_0;
	c.r_e = r_eByl_0 * l_0;
	c.r_e2 = c.r_e * c.r_e;
	c.neighbor_length = r_eByl_0 * l_0 * (1 + courant * 2);
	{
		// exact threshold of the search predicate (Computer.hpp:743: R(x_i, x_j) < neighborLength) in terms of the squared distance
		volatile double t = c.neighbor_length * c.neighbor_length;
		auto root = [](double v) { volatile double r = std::sqrt(v); return r; };
		while (t > 0 && root(t) >= c.neighbor_length) t = std::nextafter(t, 0.0);
		while (root(t) < c.neighbor_length) t = std::nextafter(t, HUGE_VAL);
		c.nl2_lim = t;
	}
	c.rho = env->rho; c.nu = env->nu; c.eps = eps;
	for (int k = 0; k < 3; k++) { c.g[k] = 0.0; c.min_x[k] = 0.0; c.max_x[k] = 0.0; c.grid_n[k] = 1; }
	c.g[D - 1] = -env->g;
	for (int k = 0; k < D; k++) { c.min_x[k] = env->min_x[k]; c.max_x[k] = env->max_x[k]; }
	{
		// n0: lattice sum over [-ceil(r_e/l0), ceil(r_e/l0))^D, r < R_e (Environment.hpp:164-209)
		const int range = static_cast<int>(std::ceil(r_eByl_0));
		double n0 = 0;
		const int kLo = (D == 3) ? -range : 0, kHi = (D == 3) ? range : 1;
		for (int i = -range; i < range; i++)
			for (int j = -range; j < range; j++)
				for (int k = kLo; k < kHi; k++)
				{
					if (!((i == 0) && (j == 0) && (k == 0)))
					{
						double x[3] = { i * l_0, j * l_0, k * l_0 };
						double t = 0;
						for (int a = 0; a < D; a++) { const double u = std::fabs(x[a]); t += u * u; }
						const double r = std::sqrt(t);
						if (r < c.r_e) n0 += ((0 < r) && (r < c.r_e)) ? (c.r_e / r - 1) : 0;
					}
				}
		c.n0 = n0;
	}
	// grid (Grid.hpp:137-150)
	c.ncells = 1;
	for (int k = 0; k < D; k++)
	{
		c.grid_n[k] = static_cast<long long>(std::ceil((c.max_x[k] - c.min_x[k]) / c.neighbor_length)) + 2;
		if (c.grid_n[k] < 1) return fail(nullptr, MPS_BAD_ARG, "empty domain");
		c.ncells *= static_cast<unsigned long long>(c.grid_n[k]);
	}
	if (c.ncells >= 0xfffffff0ull) return fail(nullptr, MPS_BAD_ARG, "too many cells for 32-bit keys");
	{
		const unsigned long long c1 = static_cast<unsigned long long>(static_cast<long long>(std::ceil(c.neighbor_length / l_0)) + 1);
		unsigned long long cap = 1; for (int k = 0; k < D; k++) cap *= c1;
		c.cell_cap = static_cast<unsigned int>(cap);
	}
	c.visc_coef = c.nu * static_cast<double>(5 - D) * c.r_e / c.n0;   // Computer.hpp:961
	c.ppe_coef = static_cast<double>(5 - D) * c.r_e / c.n0;           // Computer.hpp:1291
	{ const double d = c.l0 - c.max_dx; c.ds_d2 = d * d; }            // Computer.hpp:1572,1586

	if ((e = cudaMalloc(&s->d_sc, sizeof(DevScalars))) != cudaSuccess) return cuda_fail(nullptr, e, "cudaMalloc");
	if ((e = cudaMallocHost(&s->h_sc, sizeof(DevScalars))) != cudaSuccess) return cuda_fail(nullptr, e, "cudaMallocHost");
	std::memset(s->h_sc, 0, sizeof(DevScalars));
	if ((e = cudaMemsetAsync(s->d_sc, 0, sizeof(DevScalars), s->stream)) != cudaSuccess) return cuda_fail(nullptr, e, "cudaMemset");
	if ((e = cudaStreamSynchronize(s->stream)) != cudaSuccess) return cuda_fail(nullptr, e, "cudaStreamSynchronize");
	if ((e = cg_configure(s)) != cudaSuccess) return cuda_fail(nullptr, e, "cg_configure");
	mg_configure(s);
	*out = sp.release();
	return MPS_OK;
}

int mps_destroy(mps_handle s)
{
	NEED(s);
	cudaSetDevice(s->device);
	cudaStreamSynchronize(s->stream);
	comm_release_peers(s); // un-maps the peers' arenas, frees ours (cg.z0 / z1 point into it)
	mps_comm_release(s);
	s->comm.ext.release(); s->cg.step.release();
	if (s->cg.h_step) cudaFreeHost(s->cg.h_step);
	for (int b = 0; b < 2; b++)
	{
		s->pos[b].release(); s->vel[b].release(); s->prs[b].release(); s->nden[b].release(); s->type[b].release(); s->orig[b].release();
	}
	s->inv.release(); s->wall.release(); s->nws.release(); s->ecs.release(); s->du.release(); s->x0.release();
	s->key.release(); s->skey.release(); s->rank.release(); s->perm.release(); s->perm_orig.release(); s->perm2.release();
	s->cell_count.release(); s->cell_start.release(); s->scan_tmp.release();
	s->nbr_cnt.release(); s->nbr_ptr.release(); s->nbr.release(); s->row_len.release();
	s->cg.rowptr.release(); s->cg.col.release(); s->cg.val.release(); s->cg.b.release(); s->cg.x.release(); s->cg.r.release();
	s->cg.p0.release(); s->cg.p1.release(); s->cg.ap.release(); s->cg.partials.release(); s->cg.z0.release(); s->cg.z1.release();
	s->cg.blk_chunks.release(); s->cg.blk_bytes.release(); s->cg.chunk_of_row.release(); s->cg.chunk_base.release();
	s->cg.blob_base.release(); s->cg.blk_cost.release(); s->cg.cost_base.release(); s->cg.cta_frac.release(); s->cg.cta_speed.release(); s->cg.cta_meas.release(); s->cg.desc.release(); s->cg.live.release(); s->cg.blk_live.release(); s->cg.live_base.release(); s->cg.blobs.release(); s->cg.prof.release();
	s->stage_d.release(); s->stage_i.release(); s->flush.release();
	for (int l = 0; l < kMgMaxLevels; l++)
	{
		MgLevelBufs& b = s->mg.lv[l];
		b.flag.release(); b.rank.release(); b.key.release(); b.nbr.release(); b.child.release(); b.parent.release();
		b.S.release(); b.dinv.release(); b.r.release(); b.e0.release(); b.e1.release();
	}
	s->mg.crow.release(); s->mg.cstart.release(); s->mg.dinv0.release(); s->mg.row_s.release();
	if (s->d_sc) cudaFree(s->d_sc);
	if (s->h_sc) cudaFreeHost(s->h_sc);
	if (s->ev0) cudaEventDestroy(s->ev0);
	if (s->ev1) cudaEventDestroy(s->ev1);
	if (s->ev_cg0) cudaEventDestroy(s->ev_cg0);
	if (s->ev_cg1) cudaEventDestroy(s->ev_cg1);
	if (s->stream) cudaStreamDestroy(s->stream);
	delete s;
	return MPS_OK;
}

int mps_get_env_info(mps_handle s, mps_env_info* out)
{
	NEED(s); NEED(out);
	cudaSetDevice(s->device);
	int rc = sync_scalars(s); if (rc) return rc;
	const EnvConst& c = s->env;
	out->t = s->h_sc->t; out->dt = s->h_sc->dt; out->n0 = c.n0; out->max_dt = c.max_dt; out->max_dx = c.max_dx;
	out->r_e = c.r_e; out->neighbor_length = c.neighbor_length; out->l0 = c.l0; out->rho = c.rho; out->nu = c.nu;
	for (int k = 0; k < 3; k++) out->grid_cells[k] = c.grid_n[k];
	out->cell_capacity = c.cell_cap;
	return MPS_OK;
}

uint64_t mps_count(mps_handle s) { return s ? s->n : 0; }

int mps_add_particles(mps_handle s, uint64_t n, const double* x, const double* u, const double* p, const double* nd, const int32_t* type)
{
	NEED(s);
	if (n == 0) return MPS_OK;
	if (!x || !u || !p || !nd || !type) return fail(s, MPS_BAD_ARG, "null particle array");
	if (s->n + n >= 0xfffffff0ull) return fail(s, MPS_BAD_ARG, "too many particles for 32-bit slot indices");
	cudaSetDevice(s->device);
	const int D = s->env.dim;
	int rc = ensure_particle_capacity(s, s->n + n); if (rc) return rc;
	CU(s->stage_d.ensure(n * (2 * D + 2), s->stream));
	CU(s->stage_i.ensure(n, s->stream));
	double* dx = s->stage_d.p; double* du = dx + n * D; double* dp = du + n * D; double* dn = dp + n;
	CU(cudaMemcpyAsync(dx, x, n * D * sizeof(double), cudaMemcpyHostToDevice, s->stream));
	CU(cudaMemcpyAsync(du, u, n * D * sizeof(double), cudaMemcpyHostToDevice, s->stream));
	CU(cudaMemcpyAsync(dp, p, n * sizeof(double), cudaMemcpyHostToDevice, s->stream));
	CU(cudaMemcpyAsync(dn, nd, n * sizeof(double), cudaMemcpyHostToDevice, s->stream));
	CU(cudaMemcpyAsync(s->stage_i.p, type, n * sizeof(int32_t), cudaMemcpyHostToDevice, s->stream));
	const uint64_t first = s->n;
	s->n += n;
	CU(launch_scatter_from_orig(s, dx, du, dp, dn, s->stage_i.p, first, n, true));
	CU(cudaStreamSynchronize(s->stream)); // the caller's buffers may be reused after return
	s->searched = false;
	// a staging area of gigabytes is given back (it returns with the next upload / download): at 100M particles it is 6 GB of HBM
	if (s->stage_d.cap * sizeof(double) > (1ull << 30)) { s->stage_d.release(); s->stage_i.release(); }
	return MPS_OK;
}

int mps_upload(mps_handle s, const double* x, const double* u, const double* p, const double* nd)
{
	NEED(s);
	const uint64_t n = s->n;
	if (n == 0) return MPS_OK;
	cudaSetDevice(s->device);
	const int D = s->env.dim;
	CU(s->stage_d.ensure(n * (2 * D + 2), s->stream));
	double* dx = s->stage_d.p; double* du = dx + n * D; double* dp = du + n * D; double* dn = dp + n;
	if (x) CU(cudaMemcpyAsync(dx, x, n * D * sizeof(double), cudaMemcpyHostToDevice, s->stream));
	if (u) CU(cudaMemcpyAsync(du, u, n * D * sizeof(double), cudaMemcpyHostToDevice, s->stream));
	if (p) CU(cudaMemcpyAsync(dp, p, n * sizeof(double), cudaMemcpyHostToDevice, s->stream));
	if (nd) CU(cudaMemcpyAsync(dn, nd, n * sizeof(double), cudaMemcpyHostToDevice, s->stream));
	CU(launch_scatter_from_orig(s, x ? dx : nullptr, u ? du : nullptr, p ? dp : nullptr, nd ? dn : nullptr, nullptr, 0, n, false));
	CU(cudaStreamSynchronize(s->stream));
	return MPS_OK;
}

int mps_download(mps_handle s, double* x, double* u, double* p, double* nd, int32_t* type)
{
	NEED(s);
	const uint64_t n = s->n;
	if (n == 0) return MPS_OK;
	cudaSetDevice(s->device);
	const int D = s->env.dim;
	CU(s->stage_d.ensure(n * (2 * D + 2), s->stream));
	CU(s->stage_i.ensure(n, s->stream));
	double* dx = s->stage_d.p; double* du = dx + n * D; double* dp = du + n * D; double* dn = dp + n;
	CU(launch_gather_to_orig(s, x ? dx : nullptr, u ? du : nullptr, p ? dp : nullptr, nd ? dn : nullptr, type ? s->stage_i.p : nullptr));
	if (x) CU(cudaMemcpyAsync(x, dx, n * D * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
	if (u) CU(cudaMemcpyAsync(u, du, n * D * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
	if (p) CU(cudaMemcpyAsync(p, dp, n * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
	if (nd) CU(cudaMemcpyAsync(nd, dn, n * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
	if (type) CU(cudaMemcpyAsync(type, s->stage_i.p, n * sizeof(int32_t), cudaMemcpyDeviceToHost, s->stream));
	CU(cudaStreamSynchronize(s->stream));
	return MPS_OK;
}

int mps_set_wall_positions(mps_handle s, uint64_t n, const uint64_t* ids, const double* x)
{
	NEED(s);
	if (n == 0) return MPS_OK;
	if (!ids || !x) return fail(s, MPS_BAD_ARG, "null argument");
	cudaSetDevice(s->device);
	const int D = s->env.dim;
	for (uint64_t k = 0; k < n; k++) if (ids[k] >= s->n) return fail(s, MPS_BAD_ARG, "particle id out of range");
	CU(s->stage_d.ensure(n * (D + 1), s->stream));
	double* dx = s->stage_d.p;
	uint64_t* dids = reinterpret_cast<uint64_t*>(s->stage_d.p + n * D);
	CU(cudaMemcpyAsync(dx, x, n * D * sizeof(double), cudaMemcpyHostToDevice, s->stream));
	CU(cudaMemcpyAsync(dids, ids, n * sizeof(uint64_t), cudaMemcpyHostToDevice, s->stream));
	CU(launch_set_wall(s, n, dids, dx));
	CU(cudaStreamSynchronize(s->stream));
	return MPS_OK;
}

int mps_set_wall_motion(mps_handle s, uint64_t n, const uint64_t* ids, const mps_wall_motion* motion)
{
	NEED(s);
	cudaSetDevice(s->device);
	if (!motion)
	{
		// back to positionWall = wall[] for everybody
		if (s->n) CU(cudaMemsetAsync(s->wall_group.p, 0, s->n, s->stream));
		s->motions = WallMotions{};
		return MPS_OK;
	}
	if (s->motions.count >= kMaxWallMotions) return fail(s, MPS_BAD_ARG, "too many wall motions on this handle (8)");
	if (!(motion->t_end >= motion->t_begin)) return fail(s, MPS_BAD_ARG, "wall motion: t_end < t_begin");
	const uint64_t count = ids ? n : s->n;
	if (count == 0) return MPS_OK;
	const uint64_t* dids = nullptr;
	if (ids)
	{
		for (uint64_t k = 0; k < n; k++) if (ids[k] >= s->n) return fail(s, MPS_BAD_ARG, "particle id out of range");
		CU(s->stage_d.ensure(n, s->stream));
		CU(cudaMemcpyAsync(s->stage_d.p, ids, n * sizeof(uint64_t), cudaMemcpyHostToDevice, s->stream));
		dids = reinterpret_cast<const uint64_t*>(s->stage_d.p);
	}
	WallMotion& m = s->motions.m[s->motions.count];
	for (int a = 0; a < 3; a++) { m.amp[a] = motion->amplitude[a]; m.vel[a] = motion->velocity[a]; }
	if (s->env.dim == 2) { m.amp[2] = 0; m.vel[2] = 0; }
	m.omega = motion->omega; m.phase = motion->phase; m.t0 = motion->t_begin; m.t1 = motion->t_end;
	s->motions.count += 1;
	CU(launch_set_wall_group(s, count, dids, s->motions.count));
	CU(cudaStreamSynchronize(s->stream));
	return MPS_OK;
}

// ---- time stepping -----------------------------------------------------------------------------------------------
int mps_determine_dt(mps_handle s, double* dt)
{
	NEED(s); NEED(dt);
	cudaSetDevice(s->device);
	{ StageTimer t(s, kStDt); CU(launch_max_u2(s)); }
	int rc = sync_scalars(s); if (rc) return rc;
	double bits_as_double;
	std::memcpy(&bits_as_double, &s->h_sc->max_u2_bits, sizeof(double));
	const double max_u = std::sqrt(bits_as_double);
	*dt = (max_u == 0) ? s->env.max_dt : std::min(s->env.max_dx / max_u, s->env.max_dt); // Computer.hpp:775
	return MPS_OK;
}

int mps_set_dt(mps_handle s, double dt, int advance)
{
	NEED(s);
	cudaSetDevice(s->device);
	return set_dt_device(s, dt, advance, false);
}

int mps_set_time(mps_handle s, double t, double dt)
{
	NEED(s);
	cudaSetDevice(s->device);
	s->h_sc->t = t; s->h_sc->dt = dt; // pinned host mirror: stays valid until the copy has run (we synchronise)
	CU(cudaMemcpyAsync(&s->d_sc->t, &s->h_sc->t, 2 * sizeof(double), cudaMemcpyHostToDevice, s->stream));
	CU(cudaStreamSynchronize(s->stream));
	return MPS_OK;
}

int mps_get_time(mps_handle s, double* t, double* dt)
{
	NEED(s);
	cudaSetDevice(s->device);
	int rc = sync_scalars(s); if (rc) return rc;
	if (t) *t = s->h_sc->t;
	if (dt) *dt = s->h_sc->dt;
	return MPS_OK;
}

int mps_forward_time(mps_handle s, double dt)
{
	NEED(s);
	cudaSetDevice(s->device);
	int rc = set_dt_device(s, dt, 1, false); if (rc) return rc;   // Computer.hpp:1703-1706
	rc = step_body(s); if (rc) return rc;
	return finish_step(s);
}

int mps_forward_time_auto(mps_handle s)
{
	NEED(s);
	cudaSetDevice(s->device);
	{ StageTimer t(s, kStDt); CU(launch_max_u2(s)); }              // Computer.hpp:1748
	int rc = set_dt_device(s, 0.0, 1, true); if (rc) return rc;
	rc = step_body(s); if (rc) return rc;
	return finish_step(s);
}

int mps_run_until(mps_handle s, double t_next, uint64_t* steps)
{
	NEED(s);
	cudaSetDevice(s->device);
	uint64_t k = 0;
	int rc = sync_scalars(s); if (rc) return rc;
	while (s->h_sc->t < t_next) // Main.cpp:370
	{
		rc = mps_forward_time_auto(s);
		if (rc) break; // the failed step is not counted: *steps = steps completed, as the driver's error line reports them
		k++;
	}
	if (steps) *steps = k;
	return rc;
}

int mps_run_steps(mps_handle s, uint64_t steps, double* device_ms)
{
	NEED(s);
	cudaSetDevice(s->device);
	cudaEvent_t a, b;
	CU(cudaEventCreate(&a)); CU(cudaEventCreate(&b));
	CU(cudaEventRecord(a, s->stream));
	int rc = MPS_OK;
	for (uint64_t k = 0; k < steps && rc == MPS_OK; k++) rc = mps_forward_time_auto(s);
	CU(cudaEventRecord(b, s->stream));
	CU(cudaEventSynchronize(b));
	float ms = 0; cudaEventElapsedTime(&ms, a, b);
	cudaEventDestroy(a); cudaEventDestroy(b);
	if (device_ms) *device_ms = ms;
	return rc;
}

// ---- single stages -------------------------------------------------------------------------------------------------
#define STAGE_PROLOGUE NEED(s); cudaSetDevice(s->device)
#define STAGE_NEEDS_SEARCH do { if (!s->searched) return fail(s, MPS_BAD_ARG, "SearchNeighbor has not run since particles were added"); } while (0)

int mps_search_neighbor(mps_handle s)
{
	STAGE_PROLOGUE;
	{ StageTimer t(s, kStSearch); CU(launch_sort_and_search(s)); }
	int rc = sync_scalars(s); if (rc) return rc;
	s->stats.disabled_last = s->h_sc->disabled_now;
	return report_device_status(s);
}
int mps_compute_density(mps_handle s) { STAGE_PROLOGUE; STAGE_NEEDS_SEARCH; StageTimer t(s, kStDensity); CU(launch_density(s, false)); return MPS_OK; }
int mps_error_correction(mps_handle s) { STAGE_PROLOGUE; STAGE_NEEDS_SEARCH; StageTimer t(s, kStEcs); CU(launch_ecs(s)); return MPS_OK; }
int mps_explicit_forces(mps_handle s) { STAGE_PROLOGUE; STAGE_NEEDS_SEARCH; StageTimer t(s, kStExplicit); CU(launch_explicit(s)); return MPS_OK; }
int mps_save_x(mps_handle s) { STAGE_PROLOGUE; CU(launch_save_x(s)); return MPS_OK; }
int mps_set_ppe(mps_handle s)
{
	STAGE_PROLOGUE; STAGE_NEEDS_SEARCH;
	{ StageTimer t(s, kStPpeAssemble); CU(launch_ppe_fill(s)); }
	int rc = sync_scalars(s); if (rc) return rc;
	s->nnz_total = s->h_sc->nnz_total; s->stats.nnz = s->nnz_total;
	return MPS_OK;
}
int mps_solve_ppe(mps_handle s)
{
	STAGE_PROLOGUE;
	{ StageTimer t(s, kStCg); CU(launch_cg(s)); }
	int rc = sync_scalars(s); if (rc) return rc;
	s->stats.last_cg_iterations = s->h_sc->cg_iterations;
	s->stats.cg_iterations += s->h_sc->cg_iterations;
	s->stats.last_rr0 = s->h_sc->rr0; s->stats.last_rr = s->h_sc->rr;
	s->stats.matrix_sweeps += s->h_sc->cg_iterations + 1;
	s->stats.mg_levels = mg_active(s) ? s->h_sc->mg_levels : 0;
	s->stats.mg_cells = mg_active(s) ? s->mg.cells0 : 0;
	return report_device_status(s); // a C++ caller may catch the exception and carry on
}
int mps_assign_pressure(mps_handle s) { STAGE_PROLOGUE; StageTimer t(s, kStPressure); CU(launch_assign_pressure(s)); return MPS_OK; }
int mps_pressure_gradient(mps_handle s) { STAGE_PROLOGUE; STAGE_NEEDS_SEARCH; StageTimer t(s, kStGradient); CU(launch_gradient(s)); return MPS_OK; }
int mps_implicit_forces(mps_handle s)
{
	int rc = mps_set_ppe(s); if (rc) return rc;
	rc = mps_solve_ppe(s); if (rc) return rc;
	rc = mps_assign_pressure(s); if (rc) return rc;
	return mps_pressure_gradient(s);
}
int mps_dynamic_stabilize(mps_handle s) { STAGE_PROLOGUE; STAGE_NEEDS_SEARCH; StageTimer t(s, kStDs); CU(launch_ds(s)); return MPS_OK; }

int mps_dndt(mps_handle s, uint64_t i, double* out)
{
	STAGE_PROLOGUE; STAGE_NEEDS_SEARCH; NEED(out);
	if (i >= s->n) return fail(s, MPS_BAD_ARG, "particle id out of range");
	CU(s->stage_d.ensure(1, s->stream));
	CU(launch_dndt_one(s, i, s->stage_d.p));
	CU(cudaMemcpyAsync(out, s->stage_d.p, sizeof(double), cudaMemcpyDeviceToHost, s->stream));
	CU(cudaStreamSynchronize(s->stream));
	return MPS_OK;
}

// ---- inspection ------------------------------------------------------------------------------------------------------
int mps_get_cells(mps_handle s, int64_t* cells)
{
	STAGE_PROLOGUE; NEED(cells);
	const uint64_t n = s->n; const int D = s->env.dim;
	if (n == 0) return MPS_OK;
	CU(s->stage_d.ensure(n * D, s->stream)); // int64 and double have the same size
	long long* d = reinterpret_cast<long long*>(s->stage_d.p);
	CU(launch_get_cells(s, d));
	CU(cudaMemcpyAsync(cells, d, n * D * sizeof(long long), cudaMemcpyDeviceToHost, s->stream));
	CU(cudaStreamSynchronize(s->stream));
	return MPS_OK;
}

int mps_get_neighbors(mps_handle s, uint64_t* rowptr, uint64_t* idx)
{
	STAGE_PROLOGUE; STAGE_NEEDS_SEARCH; NEED(rowptr);
	const uint64_t n = s->n;
	std::vector<uint64_t> ptr(n + 1);
	std::vector<uint32_t> orig(n), inv(n), list(idx ? s->nbr_total : 0);
	CU(cudaMemcpyAsync(ptr.data(), s->nbr_ptr.p, (n + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, s->stream));
	CU(cudaMemcpyAsync(orig.data(), s->orig[s->cur].p, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, s->stream));
	CU(cudaMemcpyAsync(inv.data(), s->inv.p, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, s->stream));
	if (idx && s->nbr_total) CU(cudaMemcpyAsync(list.data(), s->nbr.p, s->nbr_total * sizeof(uint32_t), cudaMemcpyDeviceToHost, s->stream));
	CU(cudaStreamSynchronize(s->stream));
	uint64_t off = 0;
	for (uint64_t o = 0; o < n; o++)
	{
		const uint64_t slot = inv[o];
		rowptr[o] = off;
		const uint64_t b = ptr[slot], e = ptr[slot + 1];
		if (idx) for (uint64_t k = b; k < e; k++) idx[off + (k - b)] = orig[list[k]];
		off += e - b;
	}
	rowptr[n] = off;
	return MPS_OK;
}

int mps_get_csr_nnz(mps_handle s, uint64_t* nnz)
{
	STAGE_PROLOGUE; NEED(nnz);
	if (s->cg.external) { *nnz = s->nnz_total; return MPS_OK; }
	// reference layout: inactive rows are identity rows (Computer.hpp:1246-1256)
	std::vector<uint64_t> ptr(s->cg.n + 1);
	CU(cudaMemcpyAsync(ptr.data(), s->cg.rowptr.p, (s->cg.n + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, s->stream));
	CU(cudaStreamSynchronize(s->stream));
	uint64_t total = 0;
	for (uint64_t i = 0; i < s->cg.n; i++) { const uint64_t len = ptr[i + 1] - ptr[i]; total += len ? len : 1; }
	*nnz = total;
	return MPS_OK;
}

int mps_get_csr(mps_handle s, uint64_t* rowptr, uint32_t* col, double* val)
{
	STAGE_PROLOGUE; NEED(rowptr); NEED(col); NEED(val);
	const uint64_t n = s->cg.n;
	std::vector<uint64_t> ptr(n + 1);
	CU(cudaMemcpyAsync(ptr.data(), s->cg.rowptr.p, (n + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, s->stream));
	CU(cudaStreamSynchronize(s->stream));
	const uint64_t nnz = ptr[n];
	std::vector<uint32_t> c(nnz); std::vector<double> v(nnz);
	if (s->cg.chunked && !s->cg.external)
	{
		// decode the chunk blobs back into slot-space CSR (columns: window-local -> slot through the chunk's ranges)
		int rc = sync_scalars(s); if (rc) return rc;
		const uint64_t nchunks = s->h_sc->n_chunks, blob_total = s->h_sc->blob_total;
		std::vector<ChunkDesc> desc(nchunks);
		std::vector<unsigned char> blobs(blob_total);
		if (nchunks) CU(cudaMemcpyAsync(desc.data(), s->cg.desc.p, nchunks * sizeof(ChunkDesc), cudaMemcpyDeviceToHost, s->stream));
		if (blob_total) CU(cudaMemcpyAsync(blobs.data(), s->cg.blobs.p, blob_total, cudaMemcpyDeviceToHost, s->stream));
		CU(cudaStreamSynchronize(s->stream));
		uint64_t covered = 0;
		for (const ChunkDesc& d : desc)
		{
			if (d.row_begin != covered) return fail(s, MPS_CUDA_ERROR, "chunk descriptors do not tile the rows");
			const uint32_t nnz_pad = round_up8(d.nnz);
			const unsigned char* blob = blobs.data() + d.blob_off + kBlobHeader;
			const double* val = reinterpret_cast<const double*>(blob);
			const uint16_t* lcol = reinterpret_cast<const uint16_t*>(blob + static_cast<uint64_t>(nnz_pad) * 8u);
			const uint16_t* rowoff = lcol + nnz_pad;
			for (uint32_t lr = 0; lr < d.rows; lr++)
			{
				const uint64_t row = static_cast<uint64_t>(d.row_begin) + lr;
				if (static_cast<uint64_t>(rowoff[lr + 1] - rowoff[lr]) != ptr[row + 1] - ptr[row]) return fail(s, MPS_CUDA_ERROR, "chunk row offsets disagree with the row lengths");
				for (uint32_t e = rowoff[lr]; e < rowoff[lr + 1]; e++)
				{
					const uint32_t l = lcol[e];
					uint32_t slot = 0xffffffffu;
					for (uint32_t q = 0; q < d.nranges; q++)
						if (l >= d.range_off[q] && l < static_cast<uint32_t>(d.range_off[q]) + d.range_len[q]) { slot = d.range_start[q] + (l - d.range_off[q]); break; }
					if (slot == 0xffffffffu) return fail(s, MPS_CUDA_ERROR, "window-local column outside the chunk's window");
					const uint64_t k = ptr[row] + (e - rowoff[lr]);
					c[k] = slot; v[k] = val[e];
				}
			}
			covered += d.rows;
		}
		if (covered != n) return fail(s, MPS_CUDA_ERROR, "chunk descriptors do not cover all rows");
	}
	else if (nnz)
	{
		CU(cudaMemcpyAsync(c.data(), s->cg.col.p, nnz * sizeof(uint32_t), cudaMemcpyDeviceToHost, s->stream));
		CU(cudaMemcpyAsync(v.data(), s->cg.val.p, nnz * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
	}
	if (s->cg.external)
	{
		CU(cudaStreamSynchronize(s->stream));
		std::copy(ptr.begin(), ptr.end(), rowptr); std::copy(c.begin(), c.end(), col); std::copy(v.begin(), v.end(), val);
		return MPS_OK;
	}
	std::vector<uint32_t> orig(n), inv(n);
	CU(cudaMemcpyAsync(orig.data(), s->orig[s->cur].p, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, s->stream));
	CU(cudaMemcpyAsync(inv.data(), s->inv.p, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, s->stream));
	CU(cudaStreamSynchronize(s->stream));
	uint64_t off = 0;
	std::vector<std::pair<uint32_t, double>> row;
	for (uint64_t o = 0; o < n; o++)
	{
		const uint64_t slot = inv[o];
		rowptr[o] = off;
		row.clear();
		for (uint64_t k = ptr[slot]; k < ptr[slot + 1]; k++) row.emplace_back(orig[c[k]], v[k]);
		if (row.empty()) row.emplace_back(static_cast<uint32_t>(o), 1.0);
		std::sort(row.begin(), row.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
		for (const auto& e : row) { col[off] = e.first; val[off] = e.second; off++; }
	}
	rowptr[n] = off;
	return MPS_OK;
}

int mps_get_vec(mps_handle s, int which, double* out)
{
	STAGE_PROLOGUE; NEED(out);
	if (which < 0 || which > 8) return fail(s, MPS_BAD_ARG, "bad vector id");
	const uint64_t n = (which <= 4) ? s->cg.n : s->n;
	const int width = (which >= 7) ? s->env.dim : 1;
	if (n == 0) return MPS_OK;
	if (which <= 4 && !s->cg.b.p) return fail(s, MPS_BAD_ARG, "no PPE assembled yet");
	if (which == 2 || which == 3) { int rc = sync_scalars(s); if (rc) return rc; }
	CU(s->stage_d.ensure(n * width, s->stream));
	CU(launch_gather_vec_to_orig(s, which, s->stage_d.p));
	CU(cudaMemcpyAsync(out, s->stage_d.p, n * width * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
	CU(cudaStreamSynchronize(s->stream));
	return MPS_OK;
}

int mps_set_system(mps_handle s, uint64_t n, const uint64_t* rowptr, const uint32_t* col, const double* val, const double* b, const double* x0)
{
	STAGE_PROLOGUE;
	if (!rowptr || !b || !x0) return fail(s, MPS_BAD_ARG, "null argument");
	const uint64_t nnz = rowptr[n];
	if (nnz && (!col || !val)) return fail(s, MPS_BAD_ARG, "null argument");
	for (uint64_t k = 0; k < nnz; k++) if (col[k] >= n) return fail(s, MPS_BAD_ARG, "column out of range");
	CgBuffers& c = s->cg;
	cudaStream_t st = s->stream;
	CU(c.rowptr.ensure(n + 1, st)); CU(c.col.ensure(nnz + 1, st)); CU(c.val.ensure(nnz + 1, st));
	CU(c.b.ensure(n, st)); CU(c.x.ensure(n, st)); CU(c.r.ensure(n, st)); CU(c.p0.ensure(n, st)); CU(c.p1.ensure(n, st)); CU(c.ap.ensure(n, st));
	CU(cudaMemcpyAsync(c.rowptr.p, rowptr, (n + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
	if (nnz)
	{
		CU(cudaMemcpyAsync(c.col.p, col, nnz * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
		CU(cudaMemcpyAsync(c.val.p, val, nnz * sizeof(double), cudaMemcpyHostToDevice, st));
	}
	if (n)
	{
		CU(cudaMemcpyAsync(c.b.p, b, n * sizeof(double), cudaMemcpyHostToDevice, st));
		CU(cudaMemcpyAsync(c.x.p, x0, n * sizeof(double), cudaMemcpyHostToDevice, st));
	}
	CU(cudaStreamSynchronize(st));
	c.n = n; c.external = true; s->nnz_total = nnz;
	return MPS_OK;
}

int mps_get_solution(mps_handle s, uint64_t n, double* x)
{
	STAGE_PROLOGUE; NEED(x);
	if (n != s->cg.n || !s->cg.external) return fail(s, MPS_BAD_ARG, "no external system of that size loaded");
	if (n) CU(cudaMemcpyAsync(x, s->cg.x.p, n * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
	CU(cudaStreamSynchronize(s->stream));
	return MPS_OK;
}

// Tables of the multigrid preconditioner after the last assembly (tests/test_multigrid.py checks them against P^T A P computed
// on the host).  level >= 0: which = 0 key (u32), 1 nbr (u32 x 3^D), 2 child (u32 x 2^D), 3 parent (u32), 4 S (f64 x 3^D), 5 dinv, 6 r,
// 7 e0, 8 e1 (f64); level = -1: which = 0 row -> cell (u32 x n), 1 first row of every cell (u64 x cells + 1), 2 1 / a_ii (f64 x n),
// 3 slot -> original id (u32 x n).  *count = elements available; at most capacity_bytes are copied.
int mps_debug_mg(mps_handle s, int level, int which, void* out, uint64_t capacity_bytes, uint64_t* count)
{
	STAGE_PROLOGUE; NEED(count);
	if (!mg_active(s)) return fail(s, MPS_BAD_ARG, "no preconditioner on this handle");
	const MgBuffers& g = s->mg;
	const uint64_t K = (s->env.dim == 3) ? 27 : 9, CH = 1ull << s->env.dim;
	const void* src = nullptr; uint64_t elems = 0, esize = 8;
	if (level < 0)
	{
		switch (which)
		{
		case 0: src = g.crow.p; elems = s->n; esize = 4; break;
		case 1: src = g.cstart.p; elems = g.cells0 + 1; esize = 8; break;
		case 2: src = g.dinv0.p; elems = s->n; esize = 8; break;
		case 3: src = s->orig[s->cur].p; elems = s->n; esize = 4; break;
		default: return fail(s, MPS_BAD_ARG, "bad table id");
		}
	}
	else
	{
		if (level >= g.levels) return fail(s, MPS_BAD_ARG, "no such level");
		const MgLevelBufs& b = g.lv[level];
		uint64_t cells = 0;
		CU(cudaMemcpyAsync(&cells, b.rank.p + b.dense, sizeof(uint64_t), cudaMemcpyDeviceToHost, s->stream));
		CU(cudaStreamSynchronize(s->stream));
		switch (which)
		{
		case 0: src = b.key.p; elems = cells; esize = 4; break;
		case 1: src = b.nbr.p; elems = cells * K; esize = 4; break;
		case 2: src = b.child.p; elems = level > 0 ? cells * CH : 0; esize = 4; break;
		case 3: src = b.parent.p; elems = (level + 1 < g.levels) ? cells : 0; esize = 4; break;
		case 4: src = b.S.p; elems = cells * K; break;
		case 5: src = b.dinv.p; elems = cells; break;
		case 6: src = b.r.p; elems = cells; break;
		case 7: src = b.e0.p; elems = cells; break;
		case 8: src = b.e1.p; elems = cells; break;
		default: return fail(s, MPS_BAD_ARG, "bad table id");
		}
	}
	*count = elems;
	const uint64_t bytes = std::min<uint64_t>(elems * esize, capacity_bytes);
	if (out && bytes && src) CU(cudaMemcpyAsync(out, src, bytes, cudaMemcpyDeviceToHost, s->stream));
	CU(cudaStreamSynchronize(s->stream));
	return MPS_OK;
}

// ---- measurement -----------------------------------------------------------------------------------------------------
int mps_observe(mps_handle s, const mps_observe_params* params, mps_observables* out)
{
	NEED(s);
	if (!params || !out) return fail(s, MPS_BAD_ARG, "mps_observe: null argument");
	static_assert(sizeof(mps_observables) == 24 * sizeof(double), "mps_observables is the kernel's slot array");
	CU(cudaSetDevice(s->device));
	CU(launch_observe(s, params, reinterpret_cast<double*>(out)));
	return MPS_OK;
}

int mps_set_stage_timing(mps_handle s, int on) { NEED(s); s->stage_timing = on != 0; return MPS_OK; }
int mps_get_stats(mps_handle s, mps_stats* out)
{
	NEED(s); NEED(out);
	*out = s->stats;
	out->particles = s->n; out->neighbors = s->nbr_total; out->nnz = s->nnz_total;
	return MPS_OK;
}
int mps_reset_stats(mps_handle s) { NEED(s); s->stats = mps_stats{}; return MPS_OK; }

int mps_set_cg_profile(mps_handle s, int on) { NEED(s); s->cg_profile = on != 0; return MPS_OK; }

int mps_get_cg_profile(mps_handle s, double* out)
{
	STAGE_PROLOGUE; NEED(out);
	for (int k = 0; k < 19; k++) out[k] = 0;
	int rc = sync_scalars(s); if (rc) return rc;
	out[16] = static_cast<double>(s->h_sc->n_chunks); out[17] = static_cast<double>(s->h_sc->blob_total);
	const unsigned g = s->cg.prof_blocks;
	out[18] = g;
	if (!g || !s->cg.prof.p) return MPS_OK;
	std::vector<unsigned long long> h(8ull * g);
	CU(cudaMemcpyAsync(h.data(), s->cg.prof.p, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s->stream));
	CU(cudaStreamSynchronize(s->stream));
	for (unsigned b = 0; b < g; b++)
		for (int k = 0; k < 8; k++)
		{
			const double v = static_cast<double>(h[8ull * b + k]);
			out[k] += v / g;
			if (v > out[8 + k]) out[8 + k] = v;
		}
	return MPS_OK;
}

int mps_get_cg_profile_raw(mps_handle s, uint64_t* out, uint64_t capacity_ctas, uint64_t* ctas)
{
	STAGE_PROLOGUE; NEED(out); NEED(ctas);
	const unsigned g = s->cg.prof_blocks;
	*ctas = g;
	if (!g || !s->cg.prof.p || capacity_ctas < g) return MPS_OK;
	CU(cudaMemcpyAsync(out, s->cg.prof.p, 8ull * g * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s->stream));
	CU(cudaStreamSynchronize(s->stream));
	return MPS_OK;
}

int mps_get_cg_profile_stages(mps_handle s, uint64_t* out)
{
	STAGE_PROLOGUE; NEED(out);
	for (int k = 0; k < 64; k++) out[k] = 0;
	if (!s->cg.prof_stages || !s->cg.prof.p || !s->cg.prof_blocks) return MPS_OK;
	CU(cudaMemcpyAsync(out, s->cg.prof.p + 8ull * s->cg.prof_blocks, 64 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s->stream));
	CU(cudaStreamSynchronize(s->stream));
	return MPS_OK;
}

int mps_flush_l2(mps_handle s)
{
	STAGE_PROLOGUE;
	const size_t bytes = 256ull << 20; // > 126 MB L2
	CU(s->flush.ensure(bytes, s->stream));
	CU(cudaMemsetAsync(s->flush.p, 1, bytes, s->stream));
	return MPS_OK;
}

int mps_time_kernel(mps_handle s, const char* name, int reps, double* mean_ms, double* algorithmic_bytes)
{
	STAGE_PROLOGUE; STAGE_NEEDS_SEARCH; NEED(name); NEED(mean_ms);
	if (reps < 1) reps = 1;
	const std::string k(name);
	const uint64_t n = s->n; const int D = s->env.dim;
	cudaEvent_t a, b;
	CU(cudaEventCreate(&a)); CU(cudaEventCreate(&b));
	double total = 0, bytes = 0;
	for (int r = 0; r < reps; r++)
	{
		if (k == "density")
		{
			// Idempotent on the current state.  Algorithmic bytes: read x 8D + type 1, write N 8 + nWithoutSpp 8
			CU(cudaEventRecord(a, s->stream));
			CU(launch_density(s, false));
			CU(cudaEventRecord(b, s->stream));
			bytes = static_cast<double>(n) * (8.0 * D + 1 + 16);
		}
		else if (k == "cg_solve")
		{
			// re-assemble (untimed: restores x0 = P) then time one full solve
			CU(launch_ppe_fill(s));
			CU(cudaEventRecord(a, s->stream));
			CU(launch_cg(s));
			CU(cudaEventRecord(b, s->stream));
		}
		else
		{
			cudaEventDestroy(a); cudaEventDestroy(b);
			return fail(s, MPS_BAD_ARG, "unknown kernel name");
		}
		CU(cudaEventSynchronize(b));
		float ms = 0; cudaEventElapsedTime(&ms, a, b);
		total += ms;
	}
	cudaEventDestroy(a); cudaEventDestroy(b);
	if (k == "cg_solve")
	{
		int rc = sync_scalars(s); if (rc) return rc;
		const uint64_t active = s->h_sc->active_rows;
		s->stats.active_rows = active;
		s->stats.last_cg_iterations = s->h_sc->cg_iterations;
		// B_iter = 12 nnz + 92 rows
		bytes = static_cast<double>(s->h_sc->cg_iterations) * (12.0 * static_cast<double>(s->h_sc->nnz_total) + 92.0 * static_cast<double>(active));
		if (s->h_sc->error == MPS_CG_NOT_CONVERGED) clear_device_error(s);
	}
	*mean_ms = total / reps;
	if (algorithmic_bytes) *algorithmic_bytes = bytes;
	return MPS_OK;
}

} // extern "C"
