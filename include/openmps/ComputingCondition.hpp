// ComputingCondition.hpp — drop-in for src/OpenMps/ComputingCondition.hpp (reference :1-54).
#ifndef COMPUTING_CONDITION_INCLUDED
#define COMPUTING_CONDITION_INCLUDED

#include "defines.hpp"

namespace { namespace OpenMps
{
	class ComputingCondition final
	{
	public:
		const double Eps = 1e-10;     // CG stopping tolerance (relative residual)
		const double StartTime;
		const double EndTime;
		const double OutputInterval;

		ComputingCondition(const double eps, const double startTime, const double endTime, const double outputInterval)
			: Eps(eps), StartTime(startTime), EndTime(endTime), OutputInterval(outputInterval)
		{}

		ComputingCondition(ComputingCondition&&) noexcept = default;
		ComputingCondition(const ComputingCondition&) = delete;
		ComputingCondition& operator =(ComputingCondition&&) = delete;
		ComputingCondition& operator =(const ComputingCondition&) = delete;
	};
}}
#endif
