// Grid.hpp — drop-in for src/OpenMps/Grid.hpp (reference :1-574).
//
// The reference's Grid is a fixed-capacity bucket grid on the host (multi_array<size_t, DIM+1>, Store :276-331, 3^DIM stencil
// iterator :334-559).  Here the grid lives on the device (cell keys -> counting sort -> cell start table, csrc/mps_grid.cu);
// this class keeps the public constants, the extents arithmetic and — above all — Grid::Exception, which the C ABI status
// MPS_CELL_OVERFLOW is turned back into.
#ifndef GRID_INCLUDED
#define GRID_INCLUDED

#include <cmath>
#include <stdexcept>
#include <utility>

#include "Vector.hpp"

namespace { namespace OpenMps
{
	class Grid final
	{
	public:
		struct Exception : public std::runtime_error
		{
			template<typename... T>
			Exception(T&&... v) : std::runtime_error{ std::forward<T>(v)... } {}
		};

		// cells of the 3^DIM stencil (Grid.hpp:62-71)
#ifdef DIM3
		static constexpr std::size_t MAX_NEIGHBOR_BLOCK = 3 * 3 * 3;
#else
		static constexpr std::size_t MAX_NEIGHBOR_BLOCK = 3 * 3;
#endif

	private:
		std::size_t maxParticles;
		std::size_t cells[DIM];

	public:
		// Grid.hpp:137-150: ceil((max - min) / blockSize) + 2 cells per axis, capacity (ceil(blockSize / l_0) + 1)^DIM per cell
		Grid(const double neighborLength, const double l_0, const Vector& minX, const Vector& maxX)
		{
			const auto perAxis = static_cast<std::size_t>(std::ceil(neighborLength / l_0)) + 1;
			maxParticles = 1;
			for (std::size_t d = 0; d < DIM; d++)
			{
				maxParticles *= perAxis;
				cells[d] = static_cast<std::size_t>(std::ceil((maxX[d] - minX[d]) / neighborLength)) + 2;
			}
		}

		std::size_t MaxParticles() const { return maxParticles; }
		std::size_t Cells(const std::size_t axis) const { return cells[axis]; }
	};
}}
#endif
