// oracle/_ref — the UNMODIFIED reference step engine behind a C ABI.  TEST INFRASTRUCTURE ONLY.
//
// This translation unit #includes the reference's own headers from /root/reference/src/OpenMps at build time
// (Computer.hpp, Grid.hpp, Particle.hpp, Environment.hpp, Vector.hpp, defines.hpp + vendored ViennaCL 1.7.1) against
// the Boost stand-in in oracle/boost_shim/ (Boost is not installed in the image).  No reference source is copied
// into this repository; the resulting .so lives in oracle/_ref/ (git-ignored, travels to the GPU box).
//
// It is used (a) to pin oracle/mps_oracle.cpp (our CPU restatement) bit-for-bit, (b) to generate the golden
// fixtures in tests/golden/, and (c) as the "reference" CPU arm of bench.py.  Nothing in the product path
// (openmps_b200/csrc, include/) may link or call it.
//
// Private stages are reached exactly like the upstream gtests do: Computer.hpp:437-460 declares
// `friend class NumberDensityTest` when TEST_NUMBERDENSITY is defined.
#define TEST_NUMBERDENSITY
#include "Computer.hpp"

#include <chrono>
#include <cstdint>
#include <cstring>
#include <memory>

namespace { namespace OpenMps {

struct WallState
{
	std::vector<Vector> target; // what positionWall(i, t, dt) returns (Main.cpp:312-315: the initial position)
};

struct PositionWall
{
	const WallState* w;
	Vector operator()(const std::size_t i, const double, const double) const { return w->target[i]; }
};
struct PositionWallPre
{
	void operator()(const double, const double) const {}
};

using Comp = Computer<const PositionWall&, const PositionWallPre&>;

struct Handle
{
	WallState wall;
	PositionWall pw;
	PositionWallPre pwp;
	std::unique_ptr<Comp> comp;
	std::string lastError;
};

// the friend declared by Computer.hpp:441
class NumberDensityTest
{
public:
	static std::vector<Particle>& Particles(Comp& c) { return c.particles; }
	static Environment& Env(Comp& c) { return c.environment; }
	static Grid& GetGrid(Comp& c) { return c.grid; }
	static std::size_t NeighborCount(Comp& c, std::size_t i) { return c.NeighborCount(i); }
	static std::size_t Neighbor(Comp& c, std::size_t i, std::size_t k) { return c.Neighbor(i, k); }
	static auto& Ppe(Comp& c) { return c.ppe; }
	static auto& Du(Comp& c) { return c.du; }
	static auto& Ecs(Comp& c) { return c.ecs; }
	static auto& OriginalX(Comp& c) { return c.originalX; }
	static auto& NWithoutSpp(Comp& c) { return c.nWithoutSpp; }
	static double DetermineDt(Comp& c) { return c.DetermineDt(); }
	static double DnDt(Comp& c, std::size_t i) { return c.NeighborDensityVariationSpeed(i); }

	static void Stage(Comp& c, const std::string& s)
	{
		if (s == "search") c.SearchNeighbor();
		else if (s == "density") c.ComputeNeighborDensities();
		else if (s == "ecs") c.ComputeErrorCorrection();
		else if (s == "explicit") c.ComputeExplicitForces();
		else if (s == "savex") c.SaveX();
		else if (s == "setppe") c.SetPressurePoissonEquation();
		else if (s == "solveppe") c.SolvePressurePoissonEquation();
		else if (s == "implicit") c.ComputeImplicitForces();
		else if (s == "gradient") c.ModifyByPressureGradient();
		else if (s == "ds") c.DynamicStabilize();
		else throw std::invalid_argument("unknown stage " + s);
	}
};
using F = NumberDensityTest;

}}

using namespace OpenMps;

namespace {
Vector MakeVec(const double* p)
{
	Vector v;
	for (std::size_t k = 0; k < DIM; k++) v[k] = p[k];
	return v;
}
template<typename FN>
int Guard(Handle* h, FN&& fn)
{
	try { fn(); return 0; }
	catch (const Comp::Exception& e) { h->lastError = e.what(); return 1; }
	catch (const Grid::Exception& e) { h->lastError = e.what(); return 2; }
	catch (const std::exception& e) { h->lastError = e.what(); return 5; }
}
}

extern "C" {

int ref_dim() { return static_cast<int>(DIM); }
int ref_central_gravity()
{
#ifdef CENTRAL_GRAVITY
	return 1;
#else
	return 0;
#endif
}

// Environment ctor argument order: Environment.hpp:101-128
void* ref_create(double maxDt, double courant, double g, double rho, double nu, double r_eByl_0, double l_0,
	const double* minX, const double* maxX, double eps)
{
	auto h = new Handle();
	h->pw.w = &h->wall;
#ifdef DIM3
	Environment env(maxDt, courant, g, rho, nu, r_eByl_0, l_0, minX[0], minX[1], minX[2], maxX[0], maxX[1], maxX[2]);
#else
	Environment env(maxDt, courant, g, rho, nu, r_eByl_0, l_0, minX[0], minX[1], maxX[0], maxX[1]);
#endif
	h->comp.reset(new Comp(eps, env, h->pw, h->pwp));
	return h;
}
void ref_destroy(void* hv) { delete static_cast<Handle*>(hv); }
const char* ref_last_error(void* hv) { return static_cast<Handle*>(hv)->lastError.c_str(); }

// Main.cpp:304-329: walls return their initial position
void ref_add_particles(void* hv, uint64_t n, const double* x, const double* u, const double* p, const double* nd, const int32_t* type)
{
	auto h = static_cast<Handle*>(hv);
	std::vector<Particle> ps;
	ps.reserve(n);
	for (uint64_t i = 0; i < n; i++)
	{
		Particle q(static_cast<Particle::Type>(type[i]));
		q.X() = MakeVec(x + i * DIM);
		q.U() = MakeVec(u + i * DIM);
		q.P() = p[i];
		q.N() = nd[i];
		h->wall.target.push_back(q.X());
		ps.push_back(std::move(q));
	}
	h->comp->AddParticles(std::move(ps));
}
void ref_set_wall_positions(void* hv, uint64_t n, const uint64_t* ids, const double* x)
{
	auto h = static_cast<Handle*>(hv);
	for (uint64_t k = 0; k < n; k++) h->wall.target[ids[k]] = MakeVec(x + k * DIM);
}
uint64_t ref_count(void* hv) { return F::Particles(*static_cast<Handle*>(hv)->comp).size(); }

void ref_get_state(void* hv, double* x, double* u, double* p, double* nd, int32_t* type)
{
	auto& ps = F::Particles(*static_cast<Handle*>(hv)->comp);
	for (std::size_t i = 0; i < ps.size(); i++)
	{
		for (std::size_t k = 0; k < DIM; k++) { if (x) x[i * DIM + k] = ps[i].X()[k]; if (u) u[i * DIM + k] = ps[i].U()[k]; }
		if (p) p[i] = ps[i].P();
		if (nd) nd[i] = ps[i].N();
		if (type) type[i] = static_cast<int32_t>(ps[i].TYPE());
	}
}
// overwrite fields of existing particles (type is immutable upstream except through Disable())
void ref_set_state(void* hv, const double* x, const double* u, const double* p, const double* nd)
{
	auto& ps = F::Particles(*static_cast<Handle*>(hv)->comp);
	for (std::size_t i = 0; i < ps.size(); i++)
	{
		for (std::size_t k = 0; k < DIM; k++) { if (x) ps[i].X()[k] = x[i * DIM + k]; if (u) ps[i].U()[k] = u[i * DIM + k]; }
		if (p) ps[i].P() = p[i];
		if (nd) ps[i].N() = nd[i];
	}
}

// out[0..9] = t, dt, n0, MaxDt, MaxDx, R_e, NeighborLength, L_0, Rho, Nu
void ref_get_env(void* hv, double* out)
{
	const auto& e = static_cast<Handle*>(hv)->comp->GetEnvironment();
	out[0] = e.T(); out[1] = e.Dt(); out[2] = e.N0(); out[3] = e.MaxDt; out[4] = e.MaxDx;
	out[5] = e.R_e; out[6] = e.NeighborLength; out[7] = e.L_0; out[8] = e.Rho; out[9] = e.Nu;
}
// what the upstream fixtures do before calling stages directly (test_ComputerImplicitForces.cpp:85-86)
void ref_set_dt(void* hv, double dt, int advance)
{
	auto& e = F::Env(*static_cast<Handle*>(hv)->comp);
	e.Dt() = dt;
	if (advance) e.SetNextT();
}
double ref_determine_dt(void* hv) { return F::DetermineDt(*static_cast<Handle*>(hv)->comp); }

int ref_stage(void* hv, const char* name)
{
	auto h = static_cast<Handle*>(hv);
	return Guard(h, [&] { F::Stage(*h->comp, name); });
}
// dt < 0: ForwardTime() (DetermineDt inside); else ForwardTime(dt).  Stops at the first exception.
int ref_forward(void* hv, uint64_t steps, double dt, uint64_t* done, double* seconds)
{
	auto h = static_cast<Handle*>(hv);
	uint64_t k = 0;
	const auto t0 = std::chrono::steady_clock::now();
	const int rc = Guard(h, [&] {
		for (; k < steps; k++) { if (dt < 0) h->comp->ForwardTime(); else h->comp->ForwardTime(dt); }
	});
	const auto t1 = std::chrono::steady_clock::now();
	if (done) *done = k;
	if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
	return rc;
}
int ref_run_until(void* hv, double tEnd, uint64_t* done)
{
	auto h = static_cast<Handle*>(hv);
	uint64_t k = 0;
	const int rc = Guard(h, [&] {
		while (h->comp->GetEnvironment().T() < tEnd) { h->comp->ForwardTime(); k++; }
	});
	if (done) *done = k;
	return rc;
}

// cell index per particle and axis (Grid.hpp:250-254) and the grid extents (Grid.hpp:140-150)
void ref_get_cells(void* hv, int64_t* cell /* n x DIM */)
{
	auto& c = *static_cast<Handle*>(hv)->comp;
	auto& ps = F::Particles(c);
	auto& g = F::GetGrid(c);
	for (std::size_t i = 0; i < ps.size(); i++)
	{
		cell[i * DIM + AXIS_X] = g.Block<AXIS_X>(ps[i].X());
#ifdef DIM3
		cell[i * DIM + AXIS_Y] = g.Block<AXIS_Y>(ps[i].X());
#endif
		cell[i * DIM + AXIS_Z] = g.Block<AXIS_Z>(ps[i].X());
	}
}
uint64_t ref_grid_capacity(void* hv) { return F::GetGrid(*static_cast<Handle*>(hv)->comp).MaxParticles(); }

// neighbour table as CSR (rowptr has n+1 entries).  Pass idx == nullptr to get only the row pointers.
// Disabled particles keep a stale row upstream (Computer.hpp:730 skips them); they are reported with 0 entries.
void ref_get_neighbors(void* hv, uint64_t* rowptr, uint64_t* idx)
{
	auto& c = *static_cast<Handle*>(hv)->comp;
	auto& ps = F::Particles(c);
	uint64_t off = 0;
	for (std::size_t i = 0; i < ps.size(); i++)
	{
		rowptr[i] = off;
		if (ps[i].TYPE() == Particle::Type::Disabled) continue;
		const auto cnt = F::NeighborCount(c, i);
		if (idx) for (std::size_t k = 0; k < cnt; k++) idx[off + k] = F::Neighbor(c, i, k);
		off += cnt;
	}
	rowptr[ps.size()] = off;
}

uint64_t ref_csr_nnz(void* hv) { return F::Ppe(*static_cast<Handle*>(hv)->comp).A.nnz(); }
// ViennaCL host CSR: handle1 = row offsets (u32), handle2 = columns (u32), handle = values
void ref_get_csr(void* hv, uint32_t* rowptr, uint32_t* col, double* val)
{
	auto& A = F::Ppe(*static_cast<Handle*>(hv)->comp).A;
	const auto* rp = viennacl::linalg::host_based::detail::extract_raw_pointer<unsigned int>(A.handle1());
	const auto* cp = viennacl::linalg::host_based::detail::extract_raw_pointer<unsigned int>(A.handle2());
	const auto* vp = viennacl::linalg::host_based::detail::extract_raw_pointer<double>(A.handle());
	for (std::size_t i = 0; i <= A.size1(); i++) rowptr[i] = rp[i];
	for (std::size_t k = 0; k < A.nnz(); k++) { col[k] = cp[k]; val[k] = vp[k]; }
}
// direct CSR / vector injection for the CG known-answer cases (test_ComputerConjugateGradient.cpp:93-107)
void ref_set_system(void* hv, uint64_t n, const uint32_t* rowptr, const uint32_t* col, const double* val,
	const double* b, const double* x0)
{
	auto& ppe = F::Ppe(*static_cast<Handle*>(hv)->comp);
	using Ppe = std::remove_reference_t<decltype(ppe)>;
	ppe.A = typename Ppe::Matrix{ n, n };
	ppe.x = typename Ppe::LongVector(n);
	ppe.b = typename Ppe::LongVector(n);
	ppe.cg.r = typename Ppe::LongVector(n);
	ppe.cg.p = typename Ppe::LongVector(n);
	ppe.cg.Ap = typename Ppe::LongVector(n);
	viennacl::copy(rowptr, col, val, n, n, static_cast<std::size_t>(rowptr[n]), ppe.A);
	std::vector<double> bb(b, b + n), xx(x0, x0 + n);
	viennacl::copy(bb, ppe.b);
	viennacl::copy(xx, ppe.x);
}

// which: 0 ppe.x, 1 ppe.b, 2 cg.r, 3 cg.p, 4 cg.Ap, 5 ecs, 6 nWithoutSpp (n doubles); 7 du, 8 originalX (n x DIM doubles)
int ref_get_vec(void* hv, int which, double* out)
{
	auto& c = *static_cast<Handle*>(hv)->comp;
	auto& ppe = F::Ppe(c);
	auto fromVcl = [&](const auto& v) {
		std::vector<double> tmp(v.size());
		if (v.size() > 0) viennacl::copy(v, tmp);
		std::copy(tmp.begin(), tmp.end(), out);
	};
	switch (which)
	{
	case 0: fromVcl(ppe.x); return 0;
	case 1: fromVcl(ppe.b); return 0;
	case 2: fromVcl(ppe.cg.r); return 0;
	case 3: fromVcl(ppe.cg.p); return 0;
	case 4: fromVcl(ppe.cg.Ap); return 0;
	case 5: std::copy(F::Ecs(c).begin(), F::Ecs(c).end(), out); return 0;
	case 6: std::copy(F::NWithoutSpp(c).begin(), F::NWithoutSpp(c).end(), out); return 0;
	case 7: for (std::size_t i = 0; i < F::Du(c).size(); i++) for (std::size_t k = 0; k < DIM; k++) out[i * DIM + k] = F::Du(c)[i][k]; return 0;
	case 8: for (std::size_t i = 0; i < F::OriginalX(c).size(); i++) for (std::size_t k = 0; k < DIM; k++) out[i * DIM + k] = F::OriginalX(c)[i][k]; return 0;
	default: return 5;
	}
}
double ref_dndt(void* hv, uint64_t i) { return F::DnDt(*static_cast<Handle*>(hv)->comp, i); }

}
