// Stand-in for <boost/numeric/ublas/vector.hpp> — TEST INFRASTRUCTURE ONLY (oracle build).
//
// Boost is not installed in this image and there is no network, so the reference headers
// (/root/reference/src/OpenMps/*.hpp, used UNMODIFIED at oracle build time) are compiled against
// this minimal re-statement of the uBLAS surface they touch:
//   Vector.hpp:18          c_vector<double, DIM>
//   Computer.hpp:571,652   inner_prod(r, r)            -> t = 0; t += a[i]*b[i] (left to right)
//   Computer.hpp:771,860   norm_2(v)                   -> sqrt(sum |v_i|^2) (left to right; unscaled,
//                                                         BOOST_UBLAS_SCALED_NORM is not defined upstream)
//   element-wise  +  -  *scalar  /scalar  +=  -=  /=   (uBLAS expression templates evaluate per element,
//                                                         so eager per-element evaluation is bit-identical
//                                                         when built with -ffp-contract=off)
// Nothing here is shipped in the product library.
#ifndef OPENMPS_B200_ORACLE_UBLAS_VECTOR_SHIM
#define OPENMPS_B200_ORACLE_UBLAS_VECTOR_SHIM

#include <algorithm>
#include <array>
#include <cassert>
#include <cmath>
#include <cstddef>
#include <numeric>
#include <stdexcept>
#include <string>
#include <tuple>
#include <vector>

namespace boost { namespace numeric { namespace ublas {

template<typename T, std::size_t N>
class c_vector
{
	T d[N];
public:
	using value_type = T;
	using size_type = std::size_t;

	c_vector() { for (std::size_t i = 0; i < N; i++) d[i] = T(); }
	explicit c_vector(size_type) : c_vector() {}
	c_vector(const c_vector&) = default;
	c_vector& operator=(const c_vector&) = default;

	static constexpr size_type size() { return N; }
	T& operator[](const size_type i) { assert(i < N); return d[i]; }
	const T& operator[](const size_type i) const { assert(i < N); return d[i]; }
	T& operator()(const size_type i) { assert(i < N); return d[i]; }
	const T& operator()(const size_type i) const { assert(i < N); return d[i]; }

	c_vector& operator+=(const c_vector& o) { for (std::size_t i = 0; i < N; i++) d[i] += o.d[i]; return *this; }
	c_vector& operator-=(const c_vector& o) { for (std::size_t i = 0; i < N; i++) d[i] -= o.d[i]; return *this; }
	c_vector& operator*=(const T s) { for (std::size_t i = 0; i < N; i++) d[i] *= s; return *this; }
	c_vector& operator/=(const T s) { for (std::size_t i = 0; i < N; i++) d[i] /= s; return *this; }
};

template<typename T, std::size_t N>
inline c_vector<T, N> operator+(const c_vector<T, N>& a, const c_vector<T, N>& b)
{ c_vector<T, N> r; for (std::size_t i = 0; i < N; i++) r[i] = a[i] + b[i]; return r; }
template<typename T, std::size_t N>
inline c_vector<T, N> operator-(const c_vector<T, N>& a, const c_vector<T, N>& b)
{ c_vector<T, N> r; for (std::size_t i = 0; i < N; i++) r[i] = a[i] - b[i]; return r; }
template<typename T, std::size_t N>
inline c_vector<T, N> operator-(const c_vector<T, N>& a)
{ c_vector<T, N> r; for (std::size_t i = 0; i < N; i++) r[i] = -a[i]; return r; }
template<typename T, std::size_t N, typename S, typename = std::enable_if_t<std::is_arithmetic<S>::value>>
inline c_vector<T, N> operator*(const S s, const c_vector<T, N>& a)
{ c_vector<T, N> r; for (std::size_t i = 0; i < N; i++) r[i] = static_cast<T>(s) * a[i]; return r; }
template<typename T, std::size_t N, typename S, typename = std::enable_if_t<std::is_arithmetic<S>::value>>
inline c_vector<T, N> operator*(const c_vector<T, N>& a, const S s)
{ c_vector<T, N> r; for (std::size_t i = 0; i < N; i++) r[i] = a[i] * static_cast<T>(s); return r; }
template<typename T, std::size_t N, typename S, typename = std::enable_if_t<std::is_arithmetic<S>::value>>
inline c_vector<T, N> operator/(const c_vector<T, N>& a, const S s)
{ c_vector<T, N> r; for (std::size_t i = 0; i < N; i++) r[i] = a[i] / static_cast<T>(s); return r; }

template<typename T, std::size_t N>
inline T inner_prod(const c_vector<T, N>& a, const c_vector<T, N>& b)
{ T t = T(0); for (std::size_t i = 0; i < N; i++) t += a[i] * b[i]; return t; }
template<typename T, std::size_t N>
inline T norm_2(const c_vector<T, N>& a)
{ T t = T(); for (std::size_t i = 0; i < N; i++) { const T u = std::abs(a[i]); t += u * u; } return std::sqrt(t); }

// dense heap vector: only named in the non-ViennaCL typedefs of Computer.hpp:491 (never instantiated upstream)
template<typename T>
class vector : public std::vector<T>
{
public:
	using std::vector<T>::vector;
	T& operator()(const std::size_t i) { return (*this)[i]; }
	const T& operator()(const std::size_t i) const { return (*this)[i]; }
};

}}}
#endif
