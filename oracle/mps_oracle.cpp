// oracle/mps_oracle.cpp — CPU restatement ("port") of the reference's per-timestep MPS hot path.
//
// TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this
// library; the product (openmps_b200/csrc, include/) never links or calls it and has no CPU fallback.
//
// PINNED: this restatement is checked bit-for-bit (x, u, p, n, neighbour lists, CSR, b, CG solution) against the
// reference's own code compiled here (oracle/_ref, see oracle/ref_capi.cpp) by tests/test_oracle_vs_reference.py, and
// against the committed fixtures generated from that build (tests/golden/, tests/golden/make_golden.py), and against
// the known answers of the upstream gtests (n0 = 6.539696962; CG 4x4 -> 2,4,6,8; a_ij = (5-D) r_e / n0 / r^3).
//
// Every function cites the reference lines it follows (paths relative to /root/reference/src/OpenMps).  Arithmetic is
// written in the reference's evaluation order (uBLAS evaluates element-wise and left to right) and must be compiled
// with -ffp-contract=off so that no multiply-add is fused.  Default variant only: 2-D or 3-D, MPS_HS + MPS_HL + MPS_ECS +
// MPS_DS + MPS_SPP + PRESSURE_GRADIENT_MIDPOINT (defines.hpp:11-56), optional CENTRAL_GRAVITY (Computer.hpp:925,981).
#include "mps_oracle.h"

#include <algorithm>
#include <cfloat>
#include <chrono>
#include <cmath>
#include <cstddef>
#include <cstring>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace {

enum Type : int32_t { Fluid = 0, Wall = 1, Dummy = 2, Disabled = 3 }; // Particle.hpp:16-29

struct CgFailure : std::runtime_error { using std::runtime_error::runtime_error; };   // Computer.hpp:1424-1428
struct GridOverflow : std::runtime_error { using std::runtime_error::runtime_error; }; // Grid.hpp:314-318

template<int D>
struct V
{
	double v[D];
	double& operator[](int k) { return v[k]; }
	double operator[](int k) const { return v[k]; }
};
template<int D> inline V<D> Zero() { V<D> r; for (int k = 0; k < D; k++) r[k] = 0.0; return r; }
template<int D> inline V<D> operator+(const V<D>& a, const V<D>& b) { V<D> r; for (int k = 0; k < D; k++) r[k] = a[k] + b[k]; return r; }
template<int D> inline V<D> operator-(const V<D>& a, const V<D>& b) { V<D> r; for (int k = 0; k < D; k++) r[k] = a[k] - b[k]; return r; }
template<int D> inline V<D> operator*(const double s, const V<D>& a) { V<D> r; for (int k = 0; k < D; k++) r[k] = s * a[k]; return r; }
template<int D> inline V<D> operator*(const V<D>& a, const double s) { V<D> r; for (int k = 0; k < D; k++) r[k] = a[k] * s; return r; }
template<int D> inline V<D> operator/(const V<D>& a, const double s) { V<D> r; for (int k = 0; k < D; k++) r[k] = a[k] / s; return r; }
// uBLAS inner_prod: t = 0; t += a_k * b_k, k ascending
template<int D> inline double Inner(const V<D>& a, const V<D>& b) { double t = 0.0; for (int k = 0; k < D; k++) t += a[k] * b[k]; return t; }
// uBLAS norm_2 (unscaled): sqrt(sum |a_k|^2)
template<int D> inline double Norm2(const V<D>& a) { double t = 0.0; for (int k = 0; k < D; k++) { const double u = std::fabs(a[k]); t += u * u; } return std::sqrt(t); }

inline void Add(double& s, const double c) { s += c; }
template<int D> inline void Add(V<D>& s, const V<D>& c) { for (int k = 0; k < D; k++) s[k] += c[k]; }

// Particle.hpp:71-75
inline double W(const double r, const double r_e) { return ((0 < r) && (r < r_e)) ? (r_e / r - 1) : 0; }

struct Base
{
	virtual ~Base() {}
	std::string lastError;
	virtual int Dim() const = 0;
	virtual void AddParticles(uint64_t n, const double* x, const double* u, const double* p, const double* nd, const int32_t* type) = 0;
	virtual void SetWall(uint64_t n, const uint64_t* ids, const double* x) = 0;
	virtual uint64_t Count() const = 0;
	virtual void GetState(double* x, double* u, double* p, double* nd, int32_t* type) const = 0;
	virtual void SetState(const double* x, const double* u, const double* p, const double* nd) = 0;
	virtual void GetEnv(double* out) const = 0;
	virtual void SetDt(double dt, int advance) = 0;
	virtual double DetermineDt() const = 0;
	virtual void Stage(const std::string& s) = 0;
	virtual void Forward(double dt) = 0;
	virtual double Time() const = 0;
	virtual void GetCells(int64_t* c) const = 0;
	virtual uint64_t Capacity() const = 0;
	virtual void GetNeighbors(uint64_t* rowptr, uint64_t* idx) const = 0;
	virtual uint64_t Nnz() const = 0;
	virtual void GetCsr(uint32_t* rowptr, uint32_t* col, double* val) const = 0;
	virtual void SetSystem(uint64_t n, const uint32_t* rowptr, const uint32_t* col, const double* val, const double* b, const double* x0) = 0;
	virtual int GetVec(int which, double* out) const = 0;
	virtual double DnDtOf(uint64_t i) const = 0;
	virtual uint64_t LastIterations() const = 0;
};

template<int D>
struct Oracle final : Base
{
	using Vec = V<D>;
	static constexpr int AX = 0, AY = 1, AZ = D - 1; // defines.hpp:81-90

	// --- Environment (Environment.hpp:41-74,129-216)
	double t = 0, dt = 0, n0 = 0;
	double MaxDt, MaxDx, L_0, R_e, Rho, Nu, NeighborLength, eps;
	Vec G, MinX, MaxX;
	bool centralGravity;

	// --- Grid (Grid.hpp:57-150)
	std::ptrdiff_t gridN[3] = { 1, 1, 1 };
	std::size_t cap = 0;
	std::vector<std::size_t> cellCount;
	std::vector<std::size_t> cellItems;

	// --- particles (Particle.hpp:32-45) and scratch (Computer.hpp:541-556)
	std::vector<Vec> X, U, wallTarget, du, originalX;
	std::vector<double> P, N, ecs, nWithoutSpp;
	std::vector<int32_t> T;

	// --- neighbour table (Computer.hpp:465, 594-612) as CSR
	std::vector<uint64_t> nbrPtr;
	std::vector<uint32_t> nbr;

	// --- PPE (Computer.hpp:480-537); ViennaCL CSR has u32 indices and ascending columns per row
	std::vector<uint32_t> rowPtr, col;
	std::vector<double> val, px, pb, cr, cp, cAp;
	uint64_t lastIterations = 0;

	Oracle(double maxDt, double courant, double g, double rho, double nu, double r_eByl_0, double l_0,
		const double* minX, const double* maxX, double epsIn, bool cg)
		: MaxDt(std::min(maxDt, std::sqrt(2 * (courant * l_0) / g))), // Environment.hpp:138
		MaxDx(courant * l_0), L_0(l_0), R_e(r_eByl_0 * l_0), Rho(rho), Nu(nu),
		NeighborLength(r_eByl_0 * l_0 * (1 + courant * 2)), // Environment.hpp:161
		eps(epsIn), centralGravity(cg)
	{
		G = Zero<D>(); G[AZ] = -g; // Environment.hpp:146-150
		for (int k = 0; k < D; k++) { MinX[k] = minX[k]; MaxX[k] = maxX[k]; }

		// reference particle number density: Environment.hpp:164-209
		const int range = static_cast<int>(std::ceil(r_eByl_0));
		n0 = 0;
		const int kLo = (D == 3) ? -range : 0, kHi = (D == 3) ? range : 1;
		for (int i = -range; i < range; i++)
			for (int j = -range; j < range; j++)
				for (int k = kLo; k < kHi; k++)
				{
					if (!((i == 0) && (j == 0) && (k == 0)))
					{
						Vec x;
						x[0] = i * l_0; x[1] = j * l_0;
						if (D == 3) x[D - 1] = k * l_0;
						const double r = Norm2(x);
						if (r < R_e) n0 += W(r, R_e);
					}
				}

		// grid extents: Grid.hpp:137-150 (Ceil(a, b) = ceil(a / b))
		for (int a = 0; a < D; a++)
			gridN[a] = static_cast<std::ptrdiff_t>(std::ceil((MaxX[a] - MinX[a]) / NeighborLength)) + 2;
		const std::size_t c1 = static_cast<std::size_t>(static_cast<std::ptrdiff_t>(std::ceil(NeighborLength / l_0)) + 1);
		cap = 1; for (int a = 0; a < D; a++) cap *= c1;
		std::size_t cells = 1; for (int a = 0; a < D; a++) cells *= static_cast<std::size_t>(gridN[a]);
		cellCount.assign(cells, 0);
		cellItems.assign(cells * cap, 0);
	}

	int Dim() const override { return D; }
	uint64_t Count() const override { return T.size(); }
	double Time() const override { return t; }
	uint64_t Capacity() const override { return cap; }
	uint64_t LastIterations() const override { return lastIterations; }

	static Vec Make(const double* p) { Vec v; for (int k = 0; k < D; k++) v[k] = p[k]; return v; }

	// Computer.hpp:1754-1777 (+ Main.cpp:304-315: walls are pinned to their initial position)
	void AddParticles(uint64_t n, const double* x, const double* u, const double* p, const double* nd, const int32_t* type) override
	{
		for (uint64_t i = 0; i < n; i++)
		{
			X.push_back(Make(x + i * D)); U.push_back(Make(u + i * D));
			P.push_back(p[i]); N.push_back(nd[i]); T.push_back(type[i]);
			wallTarget.push_back(X.back());
		}
		const std::size_t m = T.size();
		du.resize(m, Zero<D>()); originalX.resize(m, Zero<D>());
		ecs.resize(m, 0.0); nWithoutSpp.resize(m, 0.0);
		nbrPtr.assign(m + 1, 0);
	}
	void SetWall(uint64_t n, const uint64_t* ids, const double* x) override
	{
		for (uint64_t k = 0; k < n; k++) wallTarget[ids[k]] = Make(x + k * D);
	}
	void GetState(double* x, double* u, double* p, double* nd, int32_t* type) const override
	{
		for (std::size_t i = 0; i < T.size(); i++)
		{
			for (int k = 0; k < D; k++) { if (x) x[i * D + k] = X[i][k]; if (u) u[i * D + k] = U[i][k]; }
			if (p) p[i] = P[i];
			if (nd) nd[i] = N[i];
			if (type) type[i] = T[i];
		}
	}
	void SetState(const double* x, const double* u, const double* p, const double* nd) override
	{
		for (std::size_t i = 0; i < T.size(); i++)
		{
			for (int k = 0; k < D; k++) { if (x) X[i][k] = x[i * D + k]; if (u) U[i][k] = u[i * D + k]; }
			if (p) P[i] = p[i];
			if (nd) N[i] = nd[i];
		}
	}
	void GetEnv(double* out) const override
	{
		out[0] = t; out[1] = dt; out[2] = n0; out[3] = MaxDt; out[4] = MaxDx;
		out[5] = R_e; out[6] = NeighborLength; out[7] = L_0; out[8] = Rho; out[9] = Nu;
	}
	void SetDt(double d, int advance) override { dt = d; if (advance) t += dt; } // Environment.hpp:225-228

	// Grid.hpp:89-92,250-254
	std::ptrdiff_t Block(const Vec& x, int axis) const
	{
		return static_cast<std::ptrdiff_t>(std::floor((x[axis] - MinX[axis]) / NeighborLength));
	}
	std::size_t CellIndex(const std::ptrdiff_t* c) const
	{
		std::size_t idx = 0;
		for (int a = 0; a < D; a++) idx = idx * static_cast<std::size_t>(gridN[a]) + static_cast<std::size_t>(c[a]);
		return idx;
	}
	bool InGrid(const std::ptrdiff_t* c) const
	{
		for (int a = 0; a < D; a++) if (!((0 <= c[a]) && (c[a] < gridN[a]))) return false;
		return true;
	}
	void GetCells(int64_t* out) const override
	{
		for (std::size_t i = 0; i < T.size(); i++) for (int a = 0; a < D; a++) out[i * D + a] = Block(X[i], a);
	}

	// Computer.hpp:568-572
	static double R(const Vec& x1, const Vec& x2) { const Vec r = x1 - x2; return std::sqrt(Inner(r, r)); }

	// Computer.hpp:698-756 + Grid.hpp:222-247 (Clear), :276-331 (Store), :334-559 (stencil iterator)
	void SearchNeighbor()
	{
		std::fill(cellCount.begin(), cellCount.end(), std::size_t{ 0 });
		const std::size_t n = T.size();
		for (std::size_t i = 0; i < n; i++)
		{
			if (T[i] != Disabled)
			{
				std::ptrdiff_t c[3];
				for (int a = 0; a < D; a++) c[a] = Block(X[i], a);
				if (InGrid(c))
				{
					const std::size_t cell = CellIndex(c);
					const std::size_t l = cellCount[cell]++;
					if (l >= cap) throw GridOverflow("Too many particle in a block");
					cellItems[cell * cap + l] = i;
				}
				else
				{
					T[i] = Disabled; // Computer.hpp:711-715
				}
			}
		}

		nbr.clear();
		nbrPtr.assign(n + 1, 0);
		for (std::size_t i = 0; i < n; i++)
		{
			nbrPtr[i] = nbr.size();
			if (T[i] == Disabled) continue;
			std::ptrdiff_t c[3] = { 0, 0, 0 };
			for (int a = 0; a < D; a++) c[a] = Block(X[i], a);
			// stencil blocks in lexicographic order, x-major ... z-minor (Grid.hpp:363-403); out-of-range and empty
			// blocks are skipped (Grid.hpp:433-501, 540-548); within a block: insertion order = ascending id
			const int yLo = (D == 3) ? -1 : 0, yHi = (D == 3) ? 1 : 0;
			for (int ox = -1; ox <= 1; ox++)
				for (int oy = yLo; oy <= yHi; oy++)
					for (int oz = -1; oz <= 1; oz++)
					{
						std::ptrdiff_t b[3];
						b[AX] = c[AX] + ox;
						if (D == 3) b[AY] = c[AY] + oy;
						b[AZ] = c[AZ] + oz;
						if (!InGrid(b)) continue;
						const std::size_t cell = CellIndex(b);
						const std::size_t cnt = cellCount[cell];
						for (std::size_t l = 0; l < cnt; l++)
						{
							const std::size_t j = cellItems[cell * cap + l];
							if ((j != i) && (T[j] != Disabled))
							{
								const double r = R(X[i], X[j]);
								if (r < NeighborLength) nbr.push_back(static_cast<uint32_t>(j));
							}
						}
					}
		}
		nbrPtr[n] = nbr.size();
	}

	// Computer.hpp:618-695.  func(j, x_j, u_j, p_j, type_j) returns the pair contribution; j = -1 is the SPP virtual particle.
	template<typename SUM, typename FUNC>
	SUM Accumulate(const std::size_t i, SUM sum, const FUNC& func) const
	{
		const double r_e = R_e;
		const double r_e2 = r_e * r_e;
		Vec dx_g = Zero<D>();
		const Vec thisX = X[i];
		for (uint64_t e = nbrPtr[i]; e < nbrPtr[i + 1]; e++)
		{
			const std::size_t j = nbr[e];
			if (j != i)
			{
				const Vec dx = X[j] - thisX;
				const double r2 = Inner(dx, dx);
				if (r2 < r_e2)
				{
					Add(sum, func(static_cast<std::ptrdiff_t>(j), X[j], U[j], P[j], T[j]));
					Add(dx_g, W(std::sqrt(r2), r_e) * dx);
				}
			}
		}
		// SPP virtual particle: Computer.hpp:664-692
		{
			const double thisN = nWithoutSpp[i];
			dx_g = dx_g / n0;
			if (thisN < n0)
			{
				const double r_g = std::sqrt(Inner(dx_g, dx_g));
				if (r_g > DBL_EPSILON)
				{
					const double w_spp = n0 - thisN;
					const double r_spp = r_e / (w_spp + 1);
					const Vec x_spp = thisX - (r_spp / r_g) * dx_g;
					Add(sum, func(std::ptrdiff_t{ -1 }, x_spp, U[i], 0.0, static_cast<int32_t>(Fluid)));
				}
			}
		}
		return sum;
	}

	// Computer.hpp:759-777
	double DetermineDt() const override
	{
		std::size_t m = 0;
		for (std::size_t i = 1; i < T.size(); i++) if (Inner(U[m], U[m]) < Inner(U[i], U[i])) m = i; // std::max_element: first maximum
		const double maxU = T.empty() ? 0.0 : Norm2(U[m]);
		return (maxU == 0 ? MaxDt : std::min(MaxDx / maxU, MaxDt));
	}

	// Computer.hpp:780-833
	void ComputeNeighborDensities()
	{
		const double r_e = R_e;
		const std::size_t n = T.size();
		for (std::size_t i = 0; i < n; i++)
		{
			nWithoutSpp[i] = n0;
			if ((T[i] != Dummy) && (T[i] != Disabled))
			{
				const Vec thisX = X[i];
				const double thisN = Accumulate(i, 0.0, [&](std::ptrdiff_t j, const Vec& x, const Vec&, double, int32_t) -> double
				{
					return (j < 0) ? 0 : W(R(thisX, x), r_e);
				});
				nWithoutSpp[i] = thisN;
				N[i] = std::max(thisN, n0);
			}
		}
	}

	// Computer.hpp:838-872 (MPS_HS branch)
	double DnDt(const std::size_t i) const
	{
		const double r_e = R_e;
		const double thisN = nWithoutSpp[i];
		const Vec thisX = X[i], thisU = U[i];
		return (thisN < n0) ? 0.0 :
			-r_e * Accumulate(i, 0.0, [&](std::ptrdiff_t, const Vec& x, const Vec& u, double, int32_t) -> double
			{
				const Vec dx = x - thisX;
				const Vec duv = u - thisU;
				const double r = Norm2(dx);
				return Inner(dx, duv) / (r * r * r);
			});
	}
	double DnDtOf(uint64_t i) const override { return DnDt(i); }

	// Computer.hpp:877-910
	void ComputeErrorCorrection()
	{
		const std::size_t n = T.size();
		for (std::size_t i = 0; i < n; i++)
		{
			if ((T[i] != Dummy) && (T[i] != Disabled))
			{
				const double thisN = N[i];
				const double speed = DnDt(i);
				const double error = (thisN - n0) / n0;
				ecs[i] = std::fabs(error) * speed + std::fabs(speed) * error;
			}
		}
	}

	// Computer.hpp:914-1021
	void ComputeExplicitForces()
	{
		const double r_e = R_e, nu = Nu;
		const std::size_t n = T.size();
		std::vector<Vec>& a = du; // Computer.hpp:933
		for (std::size_t i = 0; i < n; i++)
		{
			if (T[i] == Fluid)
			{
				const Vec thisX = X[i], thisU = U[i];
				const Vec vis = Accumulate(i, Zero<D>(), [&](std::ptrdiff_t, const Vec& x, const Vec& u, double, int32_t type) -> Vec
				{
					if (type != Dummy)
					{
						const double r = R(thisX, x);
						// (5 - DIM) is a size_t in the reference; it promotes to double after nu * ...
						return (nu * static_cast<double>(5 - D) * r_e / n0 / (r * r * r)) * (u - thisU);
					}
					return Zero<D>();
				});
				a[i] = vis;
				if (centralGravity)
				{
					// Computer.hpp:981-985
					const double g = G[AZ];
					const Vec x = X[i];
					const double r = Norm2(x);
					const Vec c = (r < L_0 * 0.01) ? Zero<D>() : (x / r);
					Add(a[i], g * c);
				}
				else
				{
					Add(a[i], G);
				}
			}
		}
		// positionWallPre(t, dt): no-op in the driver (Main.cpp:316-318)
		for (std::size_t i = 0; i < n; i++)
		{
			if (T[i] == Fluid)
			{
				Add(U[i], a[i] * dt);
				Add(X[i], U[i] * dt);
			}
			else
			{
				// Wall, Dummy AND Disabled take this branch (Computer.hpp:1011-1019)
				const Vec x = wallTarget[i];
				const Vec u = (x - X[i]) / dt;
				X[i] = x;
				U[i] = u;
			}
		}
	}

	// Computer.hpp:1025-1039
	void SaveX()
	{
		for (std::size_t i = 0; i < T.size(); i++)
			if ((T[i] != Dummy) && (T[i] != Disabled)) originalX[i] = X[i];
	}

	// Computer.hpp:1145-1356
	void SetPressurePoissonEquation()
	{
		const double r_e = R_e, rho = Rho;
		const std::size_t n = T.size();
		if (n != pb.size())
		{
			px.assign(n, 0.0); pb.assign(n, 0.0); cr.assign(n, 0.0); cp.assign(n, 0.0); cAp.assign(n, 0.0);
		}
		for (std::size_t i = 0; i < n; i++)
		{
			if ((T[i] == Dummy) || (T[i] == Disabled))
			{
				pb[i] = 0; px[i] = 0;
			}
			else
			{
				const double speed = DnDt(i);
				const double e = ecs[i];
				pb[i] = -rho / (n0 * dt) * (speed + e);
				px[i] = P[i];
			}
		}

		rowPtr.assign(n + 1, 0);
		col.clear(); val.clear();
		std::vector<std::pair<uint32_t, double>> row;
		for (std::size_t i = 0; i < n; i++)
		{
			row.clear();
			if ((T[i] == Dummy) || (T[i] == Disabled))
			{
				row.emplace_back(static_cast<uint32_t>(i), 1.0);
			}
			else
			{
				const Vec thisX = X[i];
				const double a_ii = Accumulate(i, 0.0, [&](std::ptrdiff_t j, const Vec& x, const Vec&, double, int32_t type) -> double
				{
					if (type != Dummy)
					{
						const double r = R(thisX, x);
						const double a_ij = static_cast<double>(5 - D) * r_e / n0 / (r * r * r);
						if (j >= 0) row.emplace_back(static_cast<uint32_t>(j), a_ij);
						return -a_ij;
					}
					return 0.0;
				});
				row.emplace_back(static_cast<uint32_t>(i), a_ii);
			}
			// uBLAS compressed_matrix keeps columns ascending; viennacl::copy preserves that order
			std::sort(row.begin(), row.end(), [](const auto& p, const auto& q) { return p.first < q.first; });
			rowPtr[i] = static_cast<uint32_t>(col.size());
			for (const auto& e : row) { col.push_back(e.first); val.push_back(e.second); }
		}
		rowPtr[n] = static_cast<uint32_t>(col.size());
	}

	// viennacl/linalg/host_based/sparse_matrix_operations.hpp:146-183 (row sums in ascending column order)
	void SpMV(const std::vector<double>& v, std::vector<double>& out) const
	{
		const std::ptrdiff_t n = static_cast<std::ptrdiff_t>(out.size());
#pragma omp parallel for
		for (std::ptrdiff_t r = 0; r < n; r++)
		{
			double dot = 0;
			for (uint32_t k = rowPtr[r]; k < rowPtr[r + 1]; k++) dot += val[k] * v[col[k]];
			out[r] = dot;
		}
	}
	// viennacl/linalg/host_based/vector_operations.hpp:540-557 (sequential when run with one thread)
	static double Dot(const std::vector<double>& a, const std::vector<double>& b)
	{
		double temp = 0;
		for (std::size_t i = 0; i < a.size(); i++) temp += a[i] * b[i];
		return temp;
	}

	// Computer.hpp:1359-1429
	void SolvePressurePoissonEquation()
	{
		const std::size_t n = px.size();
		SpMV(px, cAp);
		for (std::size_t i = 0; i < n; i++) cr[i] = pb[i] - cAp[i];
		cp = cr;
		double rr = Dot(cr, cr);
		const double residual0 = rr * eps * eps;
		bool isConverged = (residual0 == 0);
		lastIterations = 0;
		for (std::size_t it = 0; (it < n) && (!isConverged); it++)
		{
			SpMV(cp, cAp);
			const double alpha = rr / Dot(cp, cAp);
			for (std::size_t i = 0; i < n; i++) px[i] += alpha * cp[i];
			for (std::size_t i = 0; i < n; i++) cr[i] -= alpha * cAp[i];
			const double rrNew = Dot(cr, cr);
			lastIterations++;
			isConverged = (rrNew < residual0);
			if (!isConverged)
			{
				const double beta = rrNew / rr;
				for (std::size_t i = 0; i < n; i++) cp[i] = cr[i] + beta * cp[i];
				rr = rrNew;
			}
		}
		if (!isConverged) throw CgFailure("Conjugate Gradient method couldn't solve Pressure Poison Equation");
	}

	// Computer.hpp:1433-1564 (PRESSURE_GRADIENT_MIDPOINT branch :1501-1521)
	void ModifyByPressureGradient()
	{
		const double r_e = R_e, rho = Rho;
		const std::size_t n = T.size();
		for (std::size_t i = 0; i < n; i++)
		{
			if (T[i] == Fluid)
			{
				const double thisP = P[i];
				const Vec thisX = X[i];
				const Vec d = (-dt / rho * static_cast<double>(D) / n0) *
					Accumulate(i, Zero<D>(), [&](std::ptrdiff_t, const Vec& x, const Vec&, double p, int32_t type) -> Vec
				{
					if (type != Dummy)
					{
						const Vec dx = x - thisX;
						const double r2 = Inner(dx, dx);
						return ((p + thisP) / r2 * W(std::sqrt(r2), r_e)) * dx;
					}
					return Zero<D>();
				});
				du[i] = d;
			}
		}
		for (std::size_t i = 0; i < n; i++)
		{
			if (T[i] == Fluid)
			{
				const Vec thisDu = du[i];
				Add(U[i], thisDu);
				Add(X[i], thisDu * dt);
			}
		}
	}

	// Computer.hpp:1043-1102 (implicit branch)
	void ComputeImplicitForces()
	{
		SetPressurePoissonEquation();
		SolvePressurePoissonEquation();
		for (std::size_t i = 0; i < T.size(); i++)
		{
			if ((T[i] != Dummy) && (T[i] != Disabled))
			{
				const double p = px[i];
				P[i] = (p < 0) ? 0 : p;
			}
		}
		ModifyByPressureGradient();
	}

	// Computer.hpp:1568-1656
	void DynamicStabilize()
	{
		const double d = L_0 - MaxDx;
		const double d2 = d * d;
		const std::size_t n = T.size();
		for (std::size_t i = 0; i < n; i++)
		{
			if (T[i] == Fluid)
			{
				const Vec thisX = X[i];
				const Vec x0 = originalX[i];
				const Vec result = (-1.0 / (2 * dt * n0)) *
					Accumulate(i, Zero<D>(), [&](std::ptrdiff_t j, const Vec& x, const Vec&, double, int32_t type) -> Vec
				{
					if ((type != Dummy) && (j >= 0))
					{
						const Vec dx = x - thisX;
						const double r2 = Inner(dx, dx);
						if (r2 < d2)
						{
							const Vec xx0 = originalX[static_cast<std::size_t>(j)];
							const Vec dx0 = xx0 - x0;
							const Vec e = dx0 / Norm2(dx0);
							const double r_parallel = Inner(dx, e);
							const Vec dx_perp = dx - r_parallel * e;
							const double r_perp2 = Inner(dx_perp, dx_perp);
							return (std::sqrt(d2 - r_perp2) - r_parallel) * e;
						}
						return Zero<D>();
					}
					return Zero<D>();
				});
				du[i] = result;
			}
		}
		for (std::size_t i = 0; i < n; i++)
		{
			if (T[i] == Fluid)
			{
				const Vec thisDu = du[i];
				Add(U[i], thisDu);
				Add(X[i], thisDu * dt);
			}
		}
	}

	// Computer.hpp:1700-1742
	void Forward(double step) override
	{
		if (step < 0) step = DetermineDt(); // Computer.hpp:1745-1751
		dt = step;
		t += dt;
		SearchNeighbor();
		ComputeNeighborDensities();
		ComputeErrorCorrection();
		ComputeExplicitForces();
		ComputeNeighborDensities();
		SaveX();
		ComputeImplicitForces();
		DynamicStabilize();
	}

	void Stage(const std::string& s) override
	{
		if (s == "search") SearchNeighbor();
		else if (s == "density") ComputeNeighborDensities();
		else if (s == "ecs") ComputeErrorCorrection();
		else if (s == "explicit") ComputeExplicitForces();
		else if (s == "savex") SaveX();
		else if (s == "setppe") SetPressurePoissonEquation();
		else if (s == "solveppe") SolvePressurePoissonEquation();
		else if (s == "implicit") ComputeImplicitForces();
		else if (s == "gradient") ModifyByPressureGradient();
		else if (s == "ds") DynamicStabilize();
		else throw std::invalid_argument("unknown stage " + s);
	}

	void GetNeighbors(uint64_t* rowptr, uint64_t* idx) const override
	{
		for (std::size_t i = 0; i < nbrPtr.size(); i++) rowptr[i] = nbrPtr[i];
		if (idx) for (std::size_t k = 0; k < nbr.size(); k++) idx[k] = nbr[k];
	}
	uint64_t Nnz() const override { return col.size(); }
	void GetCsr(uint32_t* rp, uint32_t* c, double* v) const override
	{
		std::copy(rowPtr.begin(), rowPtr.end(), rp);
		std::copy(col.begin(), col.end(), c);
		std::copy(val.begin(), val.end(), v);
	}
	void SetSystem(uint64_t n, const uint32_t* rp, const uint32_t* c, const double* v, const double* b, const double* x0) override
	{
		rowPtr.assign(rp, rp + n + 1); col.assign(c, c + rp[n]); val.assign(v, v + rp[n]);
		pb.assign(b, b + n); px.assign(x0, x0 + n);
		cr.assign(n, 0.0); cp.assign(n, 0.0); cAp.assign(n, 0.0);
	}
	int GetVec(int which, double* out) const override
	{
		auto flat = [&](const std::vector<Vec>& v) { for (std::size_t i = 0; i < v.size(); i++) for (int k = 0; k < D; k++) out[i * D + k] = v[i][k]; };
		switch (which)
		{
		case 0: std::copy(px.begin(), px.end(), out); return 0;
		case 1: std::copy(pb.begin(), pb.end(), out); return 0;
		case 2: std::copy(cr.begin(), cr.end(), out); return 0;
		case 3: std::copy(cp.begin(), cp.end(), out); return 0;
		case 4: std::copy(cAp.begin(), cAp.end(), out); return 0;
		case 5: std::copy(ecs.begin(), ecs.end(), out); return 0;
		case 6: std::copy(nWithoutSpp.begin(), nWithoutSpp.end(), out); return 0;
		case 7: flat(du); return 0;
		case 8: flat(originalX); return 0;
		default: return 5;
		}
	}
};

template<typename FN>
int Guard(Base* h, FN&& fn)
{
	try { fn(); return 0; }
	catch (const CgFailure& e) { h->lastError = e.what(); return 1; }
	catch (const GridOverflow& e) { h->lastError = e.what(); return 2; }
	catch (const std::exception& e) { h->lastError = e.what(); return 5; }
}
inline Base* B(void* h) { return static_cast<Base*>(h); }

}

extern "C" {

void* orc_create(int dim, int central_gravity, double maxDt, double courant, double g, double rho, double nu,
	double r_eByl_0, double l_0, const double* minX, const double* maxX, double eps)
{
	if (dim == 2) return static_cast<Base*>(new Oracle<2>(maxDt, courant, g, rho, nu, r_eByl_0, l_0, minX, maxX, eps, central_gravity != 0));
	if (dim == 3) return static_cast<Base*>(new Oracle<3>(maxDt, courant, g, rho, nu, r_eByl_0, l_0, minX, maxX, eps, central_gravity != 0));
	return nullptr;
}
void orc_destroy(void* h) { delete B(h); }
const char* orc_last_error(void* h) { return B(h)->lastError.c_str(); }
void orc_add_particles(void* h, uint64_t n, const double* x, const double* u, const double* p, const double* nd, const int32_t* type) { B(h)->AddParticles(n, x, u, p, nd, type); }
void orc_set_wall_positions(void* h, uint64_t n, const uint64_t* ids, const double* x) { B(h)->SetWall(n, ids, x); }
uint64_t orc_count(void* h) { return B(h)->Count(); }
void orc_get_state(void* h, double* x, double* u, double* p, double* nd, int32_t* type) { B(h)->GetState(x, u, p, nd, type); }
void orc_set_state(void* h, const double* x, const double* u, const double* p, const double* nd) { B(h)->SetState(x, u, p, nd); }
void orc_get_env(void* h, double* out) { B(h)->GetEnv(out); }
void orc_set_dt(void* h, double dt, int advance) { B(h)->SetDt(dt, advance); }
double orc_determine_dt(void* h) { return B(h)->DetermineDt(); }
int orc_stage(void* h, const char* name) { return Guard(B(h), [&] { B(h)->Stage(name); }); }
int orc_forward(void* h, uint64_t steps, double dt, uint64_t* done, double* seconds)
{
	uint64_t k = 0;
	const auto t0 = std::chrono::steady_clock::now();
	const int rc = Guard(B(h), [&] { for (; k < steps; k++) B(h)->Forward(dt); });
	const auto t1 = std::chrono::steady_clock::now();
	if (done) *done = k;
	if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
	return rc;
}
int orc_run_until(void* h, double tEnd, uint64_t* done)
{
	uint64_t k = 0;
	const int rc = Guard(B(h), [&] { while (B(h)->Time() < tEnd) { B(h)->Forward(-1.0); k++; } });
	if (done) *done = k;
	return rc;
}
void orc_get_cells(void* h, int64_t* cell) { B(h)->GetCells(cell); }
uint64_t orc_grid_capacity(void* h) { return B(h)->Capacity(); }
void orc_get_neighbors(void* h, uint64_t* rowptr, uint64_t* idx) { B(h)->GetNeighbors(rowptr, idx); }
uint64_t orc_csr_nnz(void* h) { return B(h)->Nnz(); }
void orc_get_csr(void* h, uint32_t* rowptr, uint32_t* col, double* val) { B(h)->GetCsr(rowptr, col, val); }
void orc_set_system(void* h, uint64_t n, const uint32_t* rowptr, const uint32_t* col, const double* val, const double* b, const double* x0) { B(h)->SetSystem(n, rowptr, col, val, b, x0); }
int orc_get_vec(void* h, int which, double* out) { return B(h)->GetVec(which, out); }
double orc_dndt(void* h, uint64_t i) { return B(h)->DnDtOf(i); }
uint64_t orc_last_iterations(void* h) { return B(h)->LastIterations(); }

}
