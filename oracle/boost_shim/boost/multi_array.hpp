// Stand-in for <boost/multi_array.hpp> — TEST INFRASTRUCTURE ONLY (oracle build).
//
// Surface used by the reference: Grid.hpp:76 multi_array<size_t, DIM+1> built from boost::extents[..][..]..,
// chained operator[] down to a real element reference (data[i][k][0]++, Grid.hpp:306-308), shape(),
// the ::index / ::element typedefs (Grid.hpp:83,292); Computer.hpp:465 multi_array<size_t, 2>,
// resize(extents) preserving old contents (Computer.hpp:1763-1765). Row-major (C order) like Boost's default.
#ifndef OPENMPS_B200_ORACLE_MULTI_ARRAY_SHIM
#define OPENMPS_B200_ORACLE_MULTI_ARRAY_SHIM

#include <algorithm>
#include <array>
#include <cassert>
#include <cstddef>
#include <stdexcept>
#include <vector>

namespace boost {

namespace detail_shim {
	template<std::size_t N>
	struct extent_gen
	{
		std::array<std::size_t, N> e;
		extent_gen<N + 1> operator[](const std::ptrdiff_t n) const
		{
			extent_gen<N + 1> r;
			for (std::size_t i = 0; i < N; i++) r.e[i] = e[i];
			r.e[N] = static_cast<std::size_t>(n);
			return r;
		}
	};
	template<>
	struct extent_gen<0>
	{
		extent_gen<1> operator[](const std::ptrdiff_t n) const
		{
			extent_gen<1> r; r.e[0] = static_cast<std::size_t>(n); return r;
		}
	};

	// view of the trailing R dimensions
	template<typename T, std::size_t R>
	struct sub_array
	{
		T* base; const std::size_t* shape; const std::size_t* stride;
		sub_array<T, R - 1> operator[](const std::ptrdiff_t i) const
		{
			assert(i >= 0 && static_cast<std::size_t>(i) < shape[0]);
			return sub_array<T, R - 1>{ base + static_cast<std::size_t>(i) * stride[0], shape + 1, stride + 1 };
		}
	};
	template<typename T>
	struct sub_array<T, 1>
	{
		T* base; const std::size_t* shape; const std::size_t* stride;
		T& operator[](const std::ptrdiff_t i) const
		{
			assert(i >= 0 && static_cast<std::size_t>(i) < shape[0]);
			return base[i];
		}
	};
}

static const detail_shim::extent_gen<0> extents = detail_shim::extent_gen<0>();

template<typename T, std::size_t N>
class multi_array
{
	static_assert(N >= 2, "shim supports rank >= 2 only");
	std::array<std::size_t, N> shp{};
	std::array<std::size_t, N> str{};
	std::vector<T> buf;

	void set_shape(const std::array<std::size_t, N>& e)
	{
		shp = e;
		std::size_t s = 1;
		for (std::size_t i = N; i-- > 0;) { str[i] = s; s *= shp[i]; }
	}

public:
	using element = T;
	using index = std::ptrdiff_t;
	using size_type = std::size_t;

	multi_array() { std::array<std::size_t, N> z{}; set_shape(z); }
	explicit multi_array(const detail_shim::extent_gen<N>& e)
	{
		set_shape(e.e);
		std::size_t total = 1; for (auto v : shp) total *= v;
		buf.assign(total, T());
	}
	multi_array(multi_array&&) noexcept = default;
	multi_array(const multi_array&) = default;
	multi_array& operator=(const multi_array&) = default;
	multi_array& operator=(multi_array&&) noexcept = default;

	const size_type* shape() const { return shp.data(); }

	detail_shim::sub_array<T, N - 1> operator[](const index i)
	{
		assert(i >= 0 && static_cast<std::size_t>(i) < shp[0]);
		return detail_shim::sub_array<T, N - 1>{ buf.data() + static_cast<std::size_t>(i) * str[0], shp.data() + 1, str.data() + 1 };
	}
	detail_shim::sub_array<const T, N - 1> operator[](const index i) const
	{
		assert(i >= 0 && static_cast<std::size_t>(i) < shp[0]);
		return detail_shim::sub_array<const T, N - 1>{ buf.data() + static_cast<std::size_t>(i) * str[0], shp.data() + 1, str.data() + 1 };
	}

	// Boost semantics: elements whose index is valid in both old and new shape are preserved
	void resize(const detail_shim::extent_gen<N>& e)
	{
		multi_array fresh(e);
		bool overlap = true;
		std::array<std::size_t, N> lim;
		for (std::size_t i = 0; i < N; i++) { lim[i] = std::min(shp[i], fresh.shp[i]); if (lim[i] == 0) overlap = false; }
		if (overlap)
		{
			std::array<std::size_t, N> idx{};
			for (;;)
			{
				std::size_t so = 0, sn = 0;
				for (std::size_t i = 0; i < N; i++) { so += idx[i] * str[i]; sn += idx[i] * fresh.str[i]; }
				// copy a contiguous run along the last axis
				std::copy(buf.begin() + static_cast<std::ptrdiff_t>(so), buf.begin() + static_cast<std::ptrdiff_t>(so + lim[N - 1]),
					fresh.buf.begin() + static_cast<std::ptrdiff_t>(sn));
				std::size_t ax = N - 1;
				for (;;)
				{
					if (ax == 0) goto done;
					ax--;
					if (++idx[ax] < lim[ax]) break;
					idx[ax] = 0;
				}
			}
		}
	done:
		*this = std::move(fresh);
	}
};

}
#endif
