// Stand-in for <boost/math/constants/constants.hpp> — TEST INFRASTRUCTURE ONLY (upstream gtest build,
// test_ComputerPressureGradient.cpp:7,192 uses pi<double>()).
#ifndef OPENMPS_B200_ORACLE_MATH_CONSTANTS_SHIM
#define OPENMPS_B200_ORACLE_MATH_CONSTANTS_SHIM
namespace boost { namespace math { namespace constants {
template<typename T> constexpr T pi() { return static_cast<T>(3.141592653589793238462643383279502884L); }
}}}
#endif
