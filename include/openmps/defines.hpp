// defines.hpp — drop-in for openmps/openmps src/OpenMps/defines.hpp (reference :1-93).
//
// The reference selects its physics variant with #defines in this file.  libopenmps_b200 implements exactly the default
// variant (MPS_HS + MPS_HL + MPS_ECS + MPS_DS + MPS_SPP, PRESSURE_GRADIENT_MIDPOINT, implicit pressure); the same macros
// are defined here so that code and tests written against the reference (#ifdef MPS_SPP ...) take the same branches.
// DIM3 and CENTRAL_GRAVITY stay compile-time switches of the HOST code (the shape of Vector); the device library takes
// them as run-time parameters of mps_create.
#ifndef DEFINE_INCLUDED
#define DEFINE_INCLUDED

// #define DIM3                          // defines.hpp:11 (pass -DDIM3)
// #define ARTIFICIAL_COLLISION_FORCE    // defines.hpp:16 — not supported (does not compile upstream either)
#define PRESSURE_GRADIENT_MIDPOINT       // defines.hpp:21
// #define PRESSURE_EXPLICIT             // defines.hpp:26 — not supported
#define MPS_HS                           // defines.hpp:31
#define MPS_HL                           // defines.hpp:36
#define MPS_ECS                          // defines.hpp:41
// #define MPS_GC                        // defines.hpp:46 — not supported
#define MPS_DS                           // defines.hpp:51
#define MPS_SPP                          // defines.hpp:56
#define USE_VIENNACL                     // defines.hpp:66 — kept only because tests key `ppe.tempA` on it; no ViennaCL is used

#if defined(ARTIFICIAL_COLLISION_FORCE) || defined(PRESSURE_EXPLICIT) || defined(MPS_GC)
#error "libopenmps_b200 implements the reference's default method switches only"
#endif

#include <cstddef>

namespace { namespace OpenMps
{
	// number of space dimensions and the index of each axis in a Vector (z is always the last: gravity points along -z)
#ifdef DIM3
	static constexpr std::size_t DIM{ 3 }, AXIS_X{ 0 }, AXIS_Y{ 1 }, AXIS_Z{ 2 };
#else
	static constexpr std::size_t DIM{ 2 }, AXIS_X{ 0 }, AXIS_Z{ 1 };
#endif
}}

#endif
