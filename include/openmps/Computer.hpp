// Computer.hpp — drop-in for openmps/openmps src/OpenMps/Computer.hpp (reference :434-1815), backed by libopenmps_b200.
//
// Same class template, constructor, ForwardTime / AddParticles / Particles / GetEnvironment, Computer::Exception and
// CreateComputer as the reference, plus the private stage methods and members its gtest suite reaches through
// `friend class ...Test` (reference :437-460).  Nothing is computed here: every method forwards to one entry point of the
// C ABI (include/mps_capi.h) and keeps host mirrors (particles in insertion order, neighbour table, ppe) that the
// reference's callers read.  The MPS step itself runs as CUDA kernels on the B200 (openmps_b200/csrc).
//
//   reference interface                                  ->  C ABI
//   Computer(eps, env, posWall, posWallPre)   :1674-1691 ->  mps_create, mps_set_time
//   ForwardTime(dt) / ForwardTime()           :1700-1751 ->  positionWall* callbacks -> mps_set_wall_positions ; mps_forward_time
//   AddParticles                              :1754-1777 ->  mps_add_particles
//   Particles()                               :1780-1783 ->  mps_download (lazy)
//   SearchNeighbor / Neighbor / NeighborCount :594-756   ->  mps_search_neighbor, mps_get_neighbors
//   ComputeNeighborDensities ... DynamicStabilize        ->  mps_compute_density ... mps_dynamic_stabilize
//   ppe.A / ppe.b / ppe.x, Set/SolvePressurePoissonEquation -> mps_set_ppe, mps_get_csr, mps_get_vec, mps_set_system, mps_solve_ppe
//   Computer::Exception / Grid::Exception                ->  MPS_CG_NOT_CONVERGED / MPS_CELL_OVERFLOW
#ifndef COMPUTER_INCLUDED
#define COMPUTER_INCLUDED

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <iterator>
#include <map>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "defines.hpp"
#include "Vector.hpp"
#include "Particle.hpp"
#include "Environment.hpp"
#include "Grid.hpp"

#include "../mps_capi.h"

namespace { namespace OpenMps
{
	template<typename POSITION_WALL, typename POSITION_WALL_PRE> class Computer;

	namespace Detail
	{
		// owns the device-side solver; movable so that Computer keeps the reference's defaulted move constructor
		struct DeviceHandle final
		{
			mps_handle h = nullptr;
			DeviceHandle() = default;
			DeviceHandle(DeviceHandle&& o) noexcept : h(o.h) { o.h = nullptr; }
			DeviceHandle(const DeviceHandle&) = delete;
			DeviceHandle& operator=(const DeviceHandle&) = delete;
			DeviceHandle& operator=(DeviceHandle&&) = delete;
			~DeviceHandle() { if (h) mps_destroy(h); }
		};

		// element of a host mirror that remembers when it is written (so that a hand-filled system is uploaded before a solve)
		class EntryProxy final
		{
			double* slot;       // existing element, or nullptr for an absent sparse entry
			std::map<std::size_t, double>* row;
			std::size_t col;
			bool* dirty;
		public:
			EntryProxy(double* s, std::map<std::size_t, double>* r, std::size_t c, bool* d) : slot(s), row(r), col(c), dirty(d) {}
			operator double() const { return slot ? *slot : 0.0; }
			EntryProxy& operator=(const double v)
			{
				if (slot) *slot = v; else if (row) (*row)[col] = v;
				*dirty = true;
				return *this;
			}
			EntryProxy& operator=(const EntryProxy& o) { return *this = static_cast<double>(o); }
			EntryProxy& operator+=(const double v) { return *this = static_cast<double>(*this) + v; }
			EntryProxy& operator-=(const double v) { return *this = static_cast<double>(*this) - v; }
		};

		// stands in for viennacl::vector<double> / ublas::vector<double> (reference :491)
		class LongVector final
		{
			std::vector<double> v;
			mutable bool dirty = false;
			template<typename PW, typename PWP> friend class OpenMps::Computer;
		public:
			LongVector() = default;
			explicit LongVector(const std::size_t n) : v(n, 0.0), dirty(true) {}
			std::size_t size() const { return v.size(); }
			void clear() { std::fill(v.begin(), v.end(), 0.0); dirty = true; }
			void resize(const std::size_t n) { v.resize(n, 0.0); dirty = true; }
			EntryProxy operator()(const std::size_t i) { return EntryProxy(&v[i], nullptr, 0, &dirty); }
			double operator()(const std::size_t i) const { return v[i]; }
			EntryProxy operator[](const std::size_t i) { return (*this)(i); }
			double operator[](const std::size_t i) const { return v[i]; }
		};

		// stands in for viennacl::compressed_matrix<double> / ublas::compressed_matrix<double> (reference :485-499):
		// rows keep their columns sorted, absent entries read as 0
		class SparseMatrix final
		{
			std::vector<std::map<std::size_t, double>> rows;
			std::size_t ncols = 0;
			mutable bool dirty = false;
			template<typename PW, typename PWP> friend class OpenMps::Computer;
		public:
			SparseMatrix() = default;
			SparseMatrix(const std::size_t n1, const std::size_t n2) : rows(n1), ncols(n2), dirty(true) {}
			std::size_t size1() const { return rows.size(); }
			std::size_t size2() const { return ncols; }
			void clear() { for (auto& r : rows) r.clear(); dirty = true; }
			EntryProxy operator()(const std::size_t i, const std::size_t j)
			{
				auto it = rows[i].find(j);
				return EntryProxy(it == rows[i].end() ? nullptr : &it->second, &rows[i], j, &dirty);
			}
			double operator()(const std::size_t i, const std::size_t j) const
			{
				const auto it = rows[i].find(j);
				return it == rows[i].end() ? 0.0 : it->second;
			}
		};
	}

	// MPS computational space (reference :434)
	template<typename POSITION_WALL, typename POSITION_WALL_PRE>
	class Computer final
	{
#ifdef TEST_CONSTRUCTOR
	friend class ConstructorTest;
#endif
#ifdef TEST_NUMBERDENSITY
	friend class NumberDensityTest;
#endif
#ifdef TEST_EXPLICITFORCES
	friend class ExplicitForcesTest;
#endif
#ifdef TEST_IMPLICITFORCES
	friend class ImplicitForcesTest;
#endif
#ifdef TEST_CONJUGATEGRADIENT
	friend class ConjugateGradientTest;
#endif
#ifdef TEST_PRESSUREGRADIENT
	friend class PressureGradientTest;
#endif
#ifdef TEST_NEIGHBORDENSITY
	friend class NeighborDensityTest;
#endif
#ifdef TEST_NEIGHBORDENSITYVARIATION
	friend class NeighborDensityVariationTest;
#endif

	public:
		struct Exception : public std::runtime_error
		{
			template<typename... T>
			Exception(T&&... v) : std::runtime_error{ std::forward<T>(v)... } {}
		};

	private:
		// host mirror of the particles in insertion order (reference :463); refreshed from the device when stale
		mutable std::vector<Particle> particles;
		mutable bool particlesStale = false;
		std::size_t lastRunSteps = 0;

		Environment environment;
		Grid grid;

		// host mirror of the neighbour table (reference :472): row i = [count, id_0, id_1, ...]
		std::vector<std::vector<std::size_t>> neighbor;

		// pressure Poisson equation (reference :480-537): host mirrors with the element access the tests use
		struct Ppe
		{
			using Matrix = Detail::SparseMatrix;
			using LongVector = Detail::LongVector;
			using TempMatrix = Detail::SparseMatrix;

			Matrix A;
			LongVector x;
			LongVector b;
			double allowableResidual;
			struct ConjugateGradient
			{
				LongVector r;
				LongVector p;
				LongVector Ap;
			} cg;
			TempMatrix tempA;
		} ppe;

		const POSITION_WALL positionWall;
		const POSITION_WALL_PRE positionWallPre;

		Detail::DeviceHandle device;
		std::vector<std::size_t> wallIds;       // particles that follow positionWall (type != fluid), insertion order
		std::vector<double> wallTargets;        // last positions uploaded for them
		bool typesStale = false;                // the device disabled particles since the mirror was refreshed

		// ---- plumbing ------------------------------------------------------------------------------------------------
		void Check(const int rc) const
		{
			if (rc == MPS_OK) return;
			const std::string msg = mps_last_error(device.h);
			if (rc == MPS_CG_NOT_CONVERGED) throw Exception(msg);   // reference :1424-1428
			if (rc == MPS_CELL_OVERFLOW) throw Grid::Exception(msg); // Grid.hpp:314-318
			throw std::runtime_error("libopenmps_b200: " + msg);
		}

		void Refresh() const
		{
			const auto n = particles.size();
			if (n == 0) { particlesStale = false; return; }
			std::vector<double> x(n * DIM), u(n * DIM), p(n), nd(n);
			std::vector<std::int32_t> type(n);
			Check(mps_download(device.h, x.data(), u.data(), p.data(), nd.data(), type.data()));
			for (std::size_t i = 0; i < n; i++)
			{
				auto& q = particles[i];
				for (std::size_t d = 0; d < DIM; d++) { q.x[d] = x[i * DIM + d]; q.u[d] = u[i * DIM + d]; }
				q.p = p[i];
				q.n = nd[i];
				q.type = static_cast<Particle::Type>(type[i]);
			}
			particlesStale = false;
		}

		// after a stage-level call (tests read `computer->particles` directly): bring the mirror up to date at once
		void AfterStage()
		{
			particlesStale = true;
			Refresh();
			RebuildWallIds();
		}

		void RebuildWallIds()
		{
			wallIds.clear();
			for (std::size_t i = 0; i < particles.size(); i++)
				if (particles[i].TYPE() != Particle::Type::IncompressibleNewton) wallIds.push_back(i);
			wallTargets.assign(wallIds.size() * DIM, std::nan(""));
		}

		// the reference calls positionWallPre(t, dt) once and positionWall(i, t, dt) for every non-fluid particle (Wall, Dummy
		// and Disabled) inside ComputeExplicitForces (:993, :1012-1019); here the callables run on the host and only positions
		// that changed since the last step are sent to the device
		void PushWallPositions()
		{
			if (typesStale) { Refresh(); RebuildWallIds(); typesStale = false; }
			const double t = environment.T(), dt = environment.Dt();
			positionWallPre(t, dt);
			std::vector<std::uint64_t> ids;
			std::vector<double> xs;
			for (std::size_t k = 0; k < wallIds.size(); k++)
			{
				const Vector x = positionWall(wallIds[k], t, dt);
				bool same = true;
				for (std::size_t d = 0; d < DIM; d++) same = same && (x[d] == wallTargets[k * DIM + d]);
				if (!same)
				{
					ids.push_back(wallIds[k]);
					for (std::size_t d = 0; d < DIM; d++) { xs.push_back(x[d]); wallTargets[k * DIM + d] = x[d]; }
				}
			}
			if (!ids.empty()) Check(mps_set_wall_positions(device.h, ids.size(), ids.data(), xs.data()));
		}

		void DownloadNeighbors()
		{
			const auto n = particles.size();
			std::vector<std::uint64_t> rowptr(n + 1, 0);
			Check(mps_get_neighbors(device.h, rowptr.data(), nullptr));
			std::vector<std::uint64_t> idx(rowptr[n] ? rowptr[n] : 1);
			Check(mps_get_neighbors(device.h, rowptr.data(), idx.data()));
			neighbor.assign(n, std::vector<std::size_t>());
			for (std::size_t i = 0; i < n; i++)
			{
				auto& row = neighbor[i];
				row.reserve(1 + (rowptr[i + 1] - rowptr[i]));
				row.push_back(static_cast<std::size_t>(rowptr[i + 1] - rowptr[i]));
				for (auto k = rowptr[i]; k < rowptr[i + 1]; k++) row.push_back(static_cast<std::size_t>(idx[k]));
			}
		}

		void DownloadVector(const int which, Detail::LongVector& dst, const std::size_t n)
		{
			dst.v.assign(n, 0.0);
			if (n) Check(mps_get_vec(device.h, which, dst.v.data()));
			dst.dirty = false;
		}

		void DownloadPpe()
		{
			const auto n = particles.size();
			std::uint64_t nnz = 0;
			Check(mps_get_csr_nnz(device.h, &nnz));
			std::vector<std::uint64_t> rowptr(n + 1, 0);
			std::vector<std::uint32_t> col(nnz ? nnz : 1);
			std::vector<double> val(nnz ? nnz : 1);
			Check(mps_get_csr(device.h, rowptr.data(), col.data(), val.data()));
			ppe.A = typename Ppe::Matrix(n, n);
			for (std::size_t i = 0; i < n; i++)
				for (auto k = rowptr[i]; k < rowptr[i + 1]; k++) ppe.A.rows[i][col[k]] = val[k];
			ppe.A.dirty = false;
			DownloadVector(1, ppe.b, n);
			DownloadVector(0, ppe.x, n);
		}

		bool PpeEditedOnHost() const { return ppe.A.dirty || ppe.b.dirty || ppe.x.dirty; }

		void UploadPpe()
		{
			const auto n = ppe.b.size();
			std::vector<std::uint64_t> rowptr(n + 1, 0);
			std::vector<std::uint32_t> col;
			std::vector<double> val;
			for (std::size_t i = 0; i < n; i++)
			{
				if (i < ppe.A.rows.size())
					for (const auto& e : ppe.A.rows[i]) { col.push_back(static_cast<std::uint32_t>(e.first)); val.push_back(e.second); }
				rowptr[i + 1] = col.size();
			}
			std::vector<double> x0(ppe.x.v);
			x0.resize(n, 0.0);
			Check(mps_set_system(device.h, n, rowptr.data(), col.data(), val.data(), ppe.b.v.data(), x0.data()));
			ppe.A.dirty = ppe.b.dirty = ppe.x.dirty = false;
		}

		// ---- the reference's private surface ---------------------------------------------------------------------------
		// distance between two points / particles (reference :568-579): host arithmetic in the reference's order
		static double R(const Vector& x1, const Vector& x2)
		{
			const auto r = x1 - x2;
			return std::sqrt(inner_prod(r, r));
		}
		static double R(const Particle& p1, const Particle& p2) { return R(p1.X(), p2.X()); }

		// reference :594-612
		auto& NeighborCount(const std::size_t i) { return neighbor[i][0]; }
		auto NeighborCount(const std::size_t i) const { return neighbor[i][0]; }
		auto& Neighbor(const std::size_t i, const std::size_t idx) { return neighbor[i][1 + idx]; }
		auto Neighbor(const std::size_t i, const std::size_t idx) const { return neighbor[i][1 + idx]; }

		// reference :698-756
		void SearchNeighbor()
		{
			Check(mps_search_neighbor(device.h));
			AfterStage();
			DownloadNeighbors();
		}

		// reference :759-777
		double DetermineDt()
		{
			double dt = 0;
			Check(mps_determine_dt(device.h, &dt));
			return dt;
		}

		void ComputeNeighborDensities() { Check(mps_compute_density(device.h)); AfterStage(); }            // :780-833
		double NeighborDensityVariationSpeed(const std::size_t i)                                              // :838-872
		{
			double s = 0;
			Check(mps_dndt(device.h, i, &s));
			return s;
		}
		void ComputeErrorCorrection() { Check(mps_error_correction(device.h)); }                              // :877-910
		void ComputeExplicitForces() { PushWallPositions(); Check(mps_explicit_forces(device.h)); AfterStage(); } // :914-1021
		void SaveX() { Check(mps_save_x(device.h)); }                                                         // :1025-1039

		// reference :1145-1356
		void SetPressurePoissonEquation()
		{
			Check(mps_set_ppe(device.h));
			DownloadPpe();
		}

		// reference :1359-1429
		void SolvePressurePoissonEquation()
		{
			const bool external = PpeEditedOnHost();
			if (external) UploadPpe();
			Check(mps_solve_ppe(device.h));
			const auto n = ppe.b.size();
			if (external)
			{
				ppe.x.v.assign(n, 0.0);
				if (n) Check(mps_get_solution(device.h, n, ppe.x.v.data()));
				ppe.x.dirty = false;
			}
			else DownloadVector(0, ppe.x, n);
			DownloadVector(2, ppe.cg.r, n);
			DownloadVector(3, ppe.cg.p, n);
			DownloadVector(4, ppe.cg.Ap, n);
		}

		// reference :1043-1102
		void ComputeImplicitForces()
		{
			SetPressurePoissonEquation();
			SolvePressurePoissonEquation();
			Check(mps_assign_pressure(device.h));
			ModifyByPressureGradient();
		}

		void ModifyByPressureGradient() { Check(mps_pressure_gradient(device.h)); AfterStage(); }            // :1433-1564
		void DynamicStabilize() { Check(mps_dynamic_stabilize(device.h)); AfterStage(); }                    // :1568-1656

	public:
		// @param allowableResidual convergence criterion of the pressure Poisson equation (relative residual)
		// @param env      constants of the computational space
		// @param posWall  position of a wall particle: posWall(i, t, dt)
		// @param posWallPre  called once per step before posWall: posWallPre(t, dt)
		Computer(
			const double allowableResidual,
			const Environment& env,
			const POSITION_WALL& posWall,
			const POSITION_WALL_PRE& posWallPre)
			: environment(env),
			grid(env.NeighborLength, env.L_0, env.MinX, env.MaxX),
			neighbor(),
			positionWall(posWall),
			positionWallPre(posWallPre)
		{
			ppe.allowableResidual = allowableResidual;

			mps_env e{};
			e.dim = static_cast<std::int32_t>(DIM);
#ifdef CENTRAL_GRAVITY
			e.central_gravity = 1;
#else
			e.central_gravity = 0;
#endif
			e.max_dt = env.ArgMaxDt(); e.courant = env.ArgCourant(); e.g = env.ArgG(); e.rho = env.Rho; e.nu = env.Nu;
			e.r_e_by_l0 = env.ArgR_eByl_0(); e.l0 = env.L_0;
			for (std::size_t d = 0; d < DIM; d++) { e.min_x[d] = env.MinX[d]; e.max_x[d] = env.MaxX[d]; }
			int dev = 0;
			if (const char* v = std::getenv("OPENMPS_B200_DEVICE")) dev = std::atoi(v);
			const int rc = mps_create(&e, allowableResidual, dev, &device.h);
			if (rc != MPS_OK) throw std::runtime_error(std::string("libopenmps_b200: ") + mps_last_error(nullptr));
			// the environment is copied with whatever t / dt the caller already set on it
			Check(mps_set_time(device.h, environment.T(), environment.Dt()));
		}

		Computer(Computer&&) noexcept = default;
		Computer(const Computer&) = delete;
		Computer& operator=(Computer&&) noexcept = delete;
		Computer& operator=(const Computer&) = delete;

		// advance time by dt (reference :1700-1742)
		void ForwardTime(const double dt)
		{
			environment.Dt() = dt;
			environment.SetNextT();
			PushWallPositions();
			particlesStale = true;
			Check(mps_forward_time(device.h, dt));
			mps_stats st;
			if (mps_get_stats(device.h, &st) == MPS_OK && st.disabled_last) typesStale = true;
		}

		// advance time by the CFL time step (reference :1745-1751)
		void ForwardTime()
		{
			const auto dt = DetermineDt();
			ForwardTime(dt);
		}

		// ---- not in the reference ----
		// The driver's inner loop `while (T() < tNext) ForwardTime();` (Main.cpp:370-376) as one call: DetermineDt, the step and
		// the time comparison stay on the device side of the C ABI (mps_run_until).  The wall callables are evaluated once, so
		// this is for walls that do not move during the interval (the reference driver's walls never do, Main.cpp:304-315).
		// Returns the number of ForwardTime() steps taken.
		std::size_t RunUntil(const double tNext)
		{
			if (!(environment.T() < tNext)) return 0;
			PushWallPositions();
			particlesStale = true;
			std::uint64_t steps = 0;
			const int rc = mps_run_until(device.h, tNext, &steps);
			double t = 0, dt = 0;
			if (mps_get_time(device.h, &t, &dt) == MPS_OK) { environment.Dt() = dt; environment.SetT(t); }
			typesStale = true; // particles may have left the grid during the interval
			lastRunSteps = static_cast<std::size_t>(steps);
			Check(rc);
			return static_cast<std::size_t>(steps);
		}
		// steps the last RunUntil() completed — also when it ended in an exception (the failed step is not counted)
		std::size_t LastRunSteps() const { return lastRunSteps; }

		// Wall motion without the host in the loop: the listed non-fluid particles (empty list: all of them) follow
		//   positionWall(i, t, dt) + velocity * tau + amplitude * (sin(omega * tau + phase) - sin(phase)),  tau = clamp(t - t_begin, 0, t_end - t_begin)
		// evaluated on the device inside the explicit stage (mps_set_wall_motion), with the positionWall callable giving the base
		// position as before.  With it RunUntil() also serves pistons and shaking tanks (the reference evaluates the callable for
		// every non-fluid particle in every step, :993, :1012-1019).
		void SetWallMotion(const std::vector<std::uint64_t>& ids, const mps_wall_motion& motion)
		{
			Check(mps_set_wall_motion(device.h, ids.size(), ids.empty() ? nullptr : ids.data(), &motion));
		}
		void ClearWallMotion()
		{
			Check(mps_set_wall_motion(device.h, 0, nullptr, nullptr));
		}

		// append particles (reference :1754-1777)
		template<typename PARTICLES>
		void AddParticles(PARTICLES&& src)
		{
			if (particlesStale) Refresh();
			const auto first = particles.size();
			particles.insert(particles.end(), std::make_move_iterator(src.begin()), std::make_move_iterator(src.end()));
			const auto n = particles.size() - first;
			std::vector<double> x(n * DIM), u(n * DIM), p(n), nd(n);
			std::vector<std::int32_t> type(n);
			for (std::size_t k = 0; k < n; k++)
			{
				const auto& q = particles[first + k];
				for (std::size_t d = 0; d < DIM; d++) { x[k * DIM + d] = q.X()[d]; u[k * DIM + d] = q.U()[d]; }
				p[k] = q.P(); nd[k] = q.N(); type[k] = static_cast<std::int32_t>(q.TYPE());
			}
			Check(mps_add_particles(device.h, n, x.data(), u.data(), p.data(), nd.data(), type.data()));
			RebuildWallIds();
			neighbor.resize(particles.size());
		}

		// particles in insertion order (reference :1780-1783)
		const auto& Particles() const
		{
			if (particlesStale) Refresh();
			return this->particles;
		}

		// constants of the computational space (reference :1786-1789)
		const Environment& GetEnvironment() const
		{
			return environment;
		}
	};

	// reference :1800-1815
	template<typename POSITION_WALL, typename POSITION_WALL_PRE>
	inline decltype(auto) CreateComputer(
		const double allowableResidual,
		const Environment& env,
		const POSITION_WALL& posWall,
		const POSITION_WALL_PRE& posWallPre)
	{
		return Computer<decltype(posWall), decltype(posWallPre)>(allowableResidual, env, posWall, posWallPre);
	}
}}
#endif
