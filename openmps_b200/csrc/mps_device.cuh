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
	int error;                      // sticky mps_status raised on the device (cell overflow, CG failure)
	unsigned int disabled_now;      // particles disabled by the last search
	unsigned long long active_rows; // PPE rows of the last assembly that are not Dummy / Disabled
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
