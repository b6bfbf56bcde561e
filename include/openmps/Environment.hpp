// Environment.hpp — drop-in for src/OpenMps/Environment.hpp (reference :1-258).
// Same constructor argument order (:101-128), same public constants, same derived values in the same evaluation order
// (MaxDt :138, MaxDx :143, R_e :145, NeighborLength :161, n0 lattice sum :164-209) and the same mutable t / dt.
// It additionally remembers the raw constructor arguments the device library needs (mps_env), see Computer.hpp.
#ifndef ENVIRONMENT_INCLUDED
#define ENVIRONMENT_INCLUDED

#include <algorithm>
#include <cmath>

#include "Vector.hpp"
#include "Particle.hpp"

namespace { namespace OpenMps
{
	class Environment final
	{
	private:
		double t;
		double dt;
		double n0;

		// raw arguments (not part of the reference's surface)
		double argMaxDt, argCourant, argG, argR_eByl_0;

	public:
		const double MaxDt;
		const double MaxDx;
		const double L_0;
		const double R_e;
		const Vector G;
		const double Rho;
		const double Nu;
		const Vector MinX;
		const Vector MaxX;
		const double NeighborLength;

		Environment(
			const double maxDt,
			const double courant,
			const double g,
			const double rho,
			const double nu,
			const double r_eByl_0,
			const double l_0,
			const double minX,
#ifdef DIM3
			const double minY,
#endif
			const double minZ,
			const double maxX,
#ifdef DIM3
			const double maxY,
#endif
			const double maxZ)
			: t(0), dt(0), n0(0),
			argMaxDt(maxDt), argCourant(courant), argG(g), argR_eByl_0(r_eByl_0),
			MaxDt(std::min(maxDt, std::sqrt(2 * (courant*l_0) / g))),
			MaxDx(courant*l_0),
			L_0(l_0),
			R_e(r_eByl_0 * l_0),
#ifdef DIM3
			G(CreateVector(0, 0, -g)),
#else
			G(CreateVector(0, -g)),
#endif
			Rho(rho),
			Nu(nu),
#ifdef DIM3
			MinX(CreateVector(minX, minY, minZ)), MaxX(CreateVector(maxX, maxY, maxZ)),
#else
			MinX(CreateVector(minX, minZ)), MaxX(CreateVector(maxX, maxZ)),
#endif
			NeighborLength(r_eByl_0 * l_0 * (1 + courant*2))
		{
			// reference particle number density: lattice sum over [-ceil(r_e/l_0), ceil(r_e/l_0))^DIM, r < R_e
			const auto range = static_cast<int>(std::ceil(r_eByl_0));
			for (auto i = -range; i < range; i++)
			{
				for (auto j = -range; j < range; j++)
				{
#ifdef DIM3
					for (auto k = -range; k < range; k++)
					{
						if (!((i == 0) && (j == 0) && (k == 0)))
						{
							const auto x = CreateVector(i*l_0, j*l_0, k*l_0);
#else
						if (!((i == 0) && (j == 0)))
						{
							const auto x = CreateVector(i*l_0, j*l_0);
#endif
							const auto r = norm_2(x);
							if (r < R_e)
							{
								n0 += Particle::W(r, R_e);
							}
						}
#ifdef DIM3
					}
#endif
				}
			}
		}

		Environment(Environment&&) noexcept = default;
		Environment(const Environment&) = default;
		Environment& operator=(const Environment&) = delete;
		Environment& operator=(Environment&&) noexcept = delete;

		void SetNextT() { t += dt; }
		double T() const { return t; }
		double& Dt() { return dt; }
		double Dt() const { return dt; }
		double N0() const { return n0; }

		// ---- not in the reference: what mps_create needs ----
		void SetT(const double value) { t = value; } // after a device-resident run of several steps (Computer::RunUntil)
		double ArgMaxDt() const { return argMaxDt; }
		double ArgCourant() const { return argCourant; }
		double ArgG() const { return argG; }
		double ArgR_eByl_0() const { return argR_eByl_0; }
	};
}}
#endif
