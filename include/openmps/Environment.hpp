// Environment.hpp — drop-in for src/OpenMps/Environment.hpp (reference :1-258).
// Same constructor argument order (:101-128), same public constants with the same values — each derived constant is
// evaluated by the same floating-point expression (MaxDt :138, MaxDx :143, R_e :145, NeighborLength :161) and n0 is the
// same lattice sum in the same order (:164-209) — and the same mutable t / dt.  The object also remembers the raw
// constructor arguments the device library needs (mps_env, see Computer.hpp).
#ifndef ENVIRONMENT_INCLUDED
#define ENVIRONMENT_INCLUDED

#include <algorithm>
#include <array>
#include <cmath>

#include "Vector.hpp"
#include "Particle.hpp"

namespace { namespace OpenMps
{
	class Environment final
	{
		// what the caller passed (not part of the reference's surface): mps_create wants the arguments, not the derived values
		struct Arguments { double maxDt, courant, g, r_eByl_0; };
		Arguments arg;

		double t = 0;
		double dt = 0;
		double n0 = 0;

		// time-step cap: the smaller of the argument and the free-fall limit dx < g dt^2 / 2 over one Courant length
		static double CapDt(const double maxDt, const double courantLength, const double g) { return std::min(maxDt, std::sqrt(2 * (courantLength) / g)); }

		static Vector Down(const double g)
		{
			Vector v = VectorZero;
			v[DIM - 1] = -g;
			return v;
		}

		template<typename... T>
		static Vector Corner(const T... coordinate)
		{
			static_assert(sizeof...(T) == DIM, "one coordinate per axis");
			const std::array<double, DIM> c = { coordinate... };
			Vector v;
			for (std::size_t d = 0; d < DIM; d++) v[d] = c[d];
			return v;
		}

		// reference particle number density: sum of W over the lattice points of [-ceil(r_e/l_0), ceil(r_e/l_0))^DIM other than the
		// origin that lie within r_e.  The index tuple counts like nested loops with the LAST axis fastest, which is the order the
		// reference adds the terms in.
		static double LatticeDensity(const double r_eByl_0, const double l_0, const double r_e)
		{
			const int range = static_cast<int>(std::ceil(r_eByl_0));
			if (range <= 0) return 0;
			double sum = 0;
			std::array<int, DIM> index;
			index.fill(-range);
			for (bool more = true; more;)
			{
				bool origin = true;
				Vector x;
				for (std::size_t d = 0; d < DIM; d++) { x[d] = index[d] * l_0; origin = origin && (index[d] == 0); }
				if (!origin)
				{
					const double r = norm_2(x);
					if (r < r_e) sum += Particle::W(r, r_e);
				}
				// next tuple
				more = false;
				for (std::size_t d = DIM; d-- > 0;)
				{
					if (++index[d] < range) { more = true; break; }
					index[d] = -range;
				}
			}
			return sum;
		}

	public:
		const double MaxDt;           // largest time step
		const double MaxDx;           // largest displacement per step = courant l_0
		const double L_0;             // initial particle spacing
		const double R_e;             // influence radius
		const Vector G;               // gravity
		const double Rho;             // density
		const double Nu;              // kinematic viscosity
		const Vector MinX;            // corners of the computational domain
		const Vector MaxX;
		const double NeighborLength;  // radius kept in the neighbour lists: r_e (1 + 2 courant)

#ifdef DIM3
		Environment(const double maxDt, const double courant, const double g, const double rho, const double nu, const double r_eByl_0, const double l_0,
			const double minX, const double minY, const double minZ, const double maxX, const double maxY, const double maxZ)
			: arg{ maxDt, courant, g, r_eByl_0 },
			MaxDt(CapDt(maxDt, courant*l_0, g)), MaxDx(courant*l_0), L_0(l_0), R_e(r_eByl_0 * l_0), G(Down(g)), Rho(rho), Nu(nu),
			MinX(Corner(minX, minY, minZ)), MaxX(Corner(maxX, maxY, maxZ)),
			NeighborLength(r_eByl_0 * l_0 * (1 + courant*2))
		{
			n0 = LatticeDensity(r_eByl_0, l_0, R_e);
		}
#else
		Environment(const double maxDt, const double courant, const double g, const double rho, const double nu, const double r_eByl_0, const double l_0,
			const double minX, const double minZ, const double maxX, const double maxZ)
			: arg{ maxDt, courant, g, r_eByl_0 },
			MaxDt(CapDt(maxDt, courant*l_0, g)), MaxDx(courant*l_0), L_0(l_0), R_e(r_eByl_0 * l_0), G(Down(g)), Rho(rho), Nu(nu),
			MinX(Corner(minX, minZ)), MaxX(Corner(maxX, maxZ)),
			NeighborLength(r_eByl_0 * l_0 * (1 + courant*2))
		{
			n0 = LatticeDensity(r_eByl_0, l_0, R_e);
		}
#endif

		// copyable and movable, never assigned (the constants)
		Environment(const Environment&) = default;
		Environment(Environment&&) noexcept = default;
		Environment& operator=(Environment&&) noexcept = delete;
		Environment& operator=(const Environment&) = delete;

		double T() const { return t; }
		double Dt() const { return dt; }
		double& Dt() { return dt; }
		void SetNextT() { t += dt; }
		double N0() const { return n0; }

		// ---- not in the reference ----
		void SetT(const double value) { t = value; } // after a device-resident run of several steps (Computer::RunUntil)
		double ArgMaxDt() const { return arg.maxDt; }
		double ArgCourant() const { return arg.courant; }
		double ArgG() const { return arg.g; }
		double ArgR_eByl_0() const { return arg.r_eByl_0; }
	};
}}
#endif
