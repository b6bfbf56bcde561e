// ComputingCondition.hpp — drop-in for src/OpenMps/ComputingCondition.hpp (reference :1-54): what the <condition> element of
// the run description holds.  Public constants with the reference's names; move-constructible only, like the reference.
#ifndef COMPUTING_CONDITION_INCLUDED
#define COMPUTING_CONDITION_INCLUDED

#include "defines.hpp"

namespace { namespace OpenMps
{
	class ComputingCondition final
	{
	public:
		const double Eps = 1e-10;                         // CG stopping tolerance (relative residual)
		const double StartTime, EndTime, OutputInterval;  // [s]

		ComputingCondition(const double eps, const double startTime, const double endTime, const double outputInterval)
			: Eps(eps), StartTime(startTime), EndTime(endTime), OutputInterval(outputInterval) {}

		ComputingCondition(ComputingCondition&&) noexcept = default;
		ComputingCondition(const ComputingCondition&) = delete;
		void operator=(const ComputingCondition&) = delete;
		void operator=(ComputingCondition&&) = delete;
	};
}}
#endif
