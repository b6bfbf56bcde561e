// Particle.hpp — drop-in for src/OpenMps/Particle.hpp (reference :1-128): the host-side particle record.
// Same fields, accessors, Type enumeration (0 fluid, 1 wall, 2 dummy, 3 disabled) and weight function; on the device the
// particles live as cell-sorted SoA arrays, this type only carries them across the API.
#ifndef PARTICLE_INCLUDED
#define PARTICLE_INCLUDED

#include <numeric>

#include "defines.hpp"
#include "Vector.hpp"

namespace { namespace OpenMps
{
	class Particle final
	{
	public:
		enum class Type
		{
			IncompressibleNewton, // 0
			Wall,                 // 1
			Dummy,                // 2
			Disabled,             // 3
		};

	private:
		Vector x;
		Vector u;
		double p;
		double n;
		Type type;

		template<typename PW, typename PWP> friend class Computer; // refreshes the mirror (type included) from the device

	public:
		Particle(const Type t) : x(VectorZero), u(VectorZero), p(0), n(0), type(t) {}
		Particle(const Particle&) = default;
		Particle(Particle&&) noexcept = default;
		Particle& operator=(const Particle&) = default;
		Particle& operator=(Particle&& src) noexcept = default;

		void Disable() { this->type = Type::Disabled; }

		// Particle.hpp:71-75
		static double W(const double r, const double r_e) { return ((0 < r) && (r < r_e)) ? (r_e / r - 1) : 0; }

		const auto& X() const { return x; }
		auto& X() { return x; }
		const auto& U() const { return u; }
		auto& U() { return u; }
		const auto& P() const { return p; }
		auto& P() { return p; }
		const auto& N() const { return n; }
		auto& N() { return n; }
		const auto& TYPE() const { return type; }
	};
}}
#endif
