// Vector.hpp — drop-in for src/OpenMps/Vector.hpp (reference :1-76) without Boost.
//
// The reference's Vector is boost::numeric::ublas::c_vector<double, DIM>; callers only use element access, the usual
// element-wise arithmetic, inner_prod and norm_2.  This is a plain aggregate with exactly that surface; arithmetic is
// evaluated element by element, left to right, like uBLAS expression templates do.
#ifndef VECTOR_INCLUDED
#define VECTOR_INCLUDED

#include <array>
#include <cmath>
#include <tuple>

#include "defines.hpp"

namespace { namespace OpenMps
{
	class Vector final
	{
	private:
		std::array<double, DIM> v;
	public:
		using value_type = double;
		using size_type = std::size_t;
		Vector() : v{} {}
		double& operator[](const std::size_t i) { return v[i]; }
		const double& operator[](const std::size_t i) const { return v[i]; }
		double& operator()(const std::size_t i) { return v[i]; }
		const double& operator()(const std::size_t i) const { return v[i]; }
		static constexpr std::size_t size() { return DIM; }
		const double* data() const { return v.data(); }
		double* data() { return v.data(); }
		auto begin() { return v.begin(); }
		auto end() { return v.end(); }
		auto begin() const { return v.begin(); }
		auto end() const { return v.end(); }

		Vector& operator+=(const Vector& o) { for (std::size_t i = 0; i < DIM; i++) v[i] += o.v[i]; return *this; }
		Vector& operator-=(const Vector& o) { for (std::size_t i = 0; i < DIM; i++) v[i] -= o.v[i]; return *this; }
		Vector& operator*=(const double s) { for (std::size_t i = 0; i < DIM; i++) v[i] *= s; return *this; }
		Vector& operator/=(const double s) { for (std::size_t i = 0; i < DIM; i++) v[i] /= s; return *this; }
	};

	inline Vector operator+(Vector a, const Vector& b) { return a += b; }
	inline Vector operator-(Vector a, const Vector& b) { return a -= b; }
	inline Vector operator-(Vector a) { for (std::size_t i = 0; i < DIM; i++) a[i] = -a[i]; return a; }
	inline Vector operator*(Vector a, const double s) { return a *= s; }
	inline Vector operator*(const double s, Vector a) { return a *= s; }
	inline Vector operator/(Vector a, const double s) { return a /= s; }
	inline bool operator==(const Vector& a, const Vector& b) { for (std::size_t i = 0; i < DIM; i++) if (a[i] != b[i]) return false; return true; }

	// uBLAS inner_prod / norm_2: accumulated from 0, left to right
	inline double inner_prod(const Vector& a, const Vector& b) { double t = 0; for (std::size_t i = 0; i < DIM; i++) t += a[i] * b[i]; return t; }
	inline double norm_2(const Vector& a) { return std::sqrt(inner_prod(a, a)); }

	namespace Detail
	{
		template<decltype(DIM) D> struct CreateVector;
		template<> struct CreateVector<2>
		{
			static auto Get(const std::tuple<double, double>& val) { Vector vec; vec[0] = std::get<0>(val); vec[1] = std::get<1>(val); return vec; }
			static auto Get(const double val) { return Get(std::make_tuple(val, val)); }
		};
		template<> struct CreateVector<3>
		{
			static auto Get(const std::tuple<double, double, double>& val) { Vector vec; vec[0] = std::get<0>(val); vec[1] = std::get<1>(val); vec[DIM - 1] = std::get<2>(val); return vec; }
			static auto Get(const double val) { return Get(std::make_tuple(val, val, val)); }
		};
	}

	// Vector.hpp:62-72 of the reference
	template<typename T, typename... ARGS>
	inline auto CreateVector(const T val, const ARGS... args) { return Detail::CreateVector<DIM>::Get(std::make_tuple(static_cast<double>(val), static_cast<double>(args)...)); }
	template<typename T>
	inline auto CreateVector(const T val) { return Detail::CreateVector<DIM>::Get(static_cast<double>(val)); }

	static const auto VectorZero = CreateVector(0);
}}
#endif
