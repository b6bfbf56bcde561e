// Stand-in for <boost/numeric/ublas/matrix_sparse.hpp> — TEST INFRASTRUCTURE ONLY (oracle build).
//
// Surface used by the reference: Computer.hpp:473 (TempMatrix typedef), :1177 ctor (n, n), :1228 clear(),
// :1346 A(i, j) = a_ij, and by viennacl::copy (viennacl/compressed_matrix.hpp:150-191): size1/size2,
// const_iterator1/2 with begin()/end()/index1()/index2()/operator*. Storage is row-major with columns kept
// sorted ascending, as uBLAS compressed_matrix<double, row_major> does (so ViennaCL sums each row in ascending
// column order, viennacl/compressed_matrix.hpp:64-79).
#ifndef OPENMPS_B200_ORACLE_UBLAS_SPARSE_SHIM
#define OPENMPS_B200_ORACLE_UBLAS_SPARSE_SHIM

#include <algorithm>
#include <cstddef>
#include <utility>
#include <vector>

namespace boost { namespace numeric { namespace ublas {

template<typename T>
class compressed_matrix
{
	using Entry = std::pair<std::size_t, T>;
	std::size_t n1 = 0, n2 = 0;
	std::vector<std::vector<Entry>> rows;

public:
	using value_type = T;
	using size_type = std::size_t;

	compressed_matrix() = default;
	compressed_matrix(const size_type s1, const size_type s2, const size_type = 0) : n1(s1), n2(s2), rows(s1) {}

	size_type size1() const { return n1; }
	size_type size2() const { return n2; }
	void clear() { for (auto& r : rows) r.clear(); }

	// element access: inserts a zero if absent (what uBLAS' sparse reference proxy does on assignment)
	T& operator()(const size_type i, const size_type j)
	{
		auto& r = rows[i];
		auto it = std::lower_bound(r.begin(), r.end(), j, [](const Entry& e, const size_type c) { return e.first < c; });
		if (it == r.end() || it->first != j) it = r.insert(it, Entry(j, T()));
		return it->second;
	}
	T operator()(const size_type i, const size_type j) const
	{
		const auto& r = rows[i];
		auto it = std::lower_bound(r.begin(), r.end(), j, [](const Entry& e, const size_type c) { return e.first < c; });
		return (it == r.end() || it->first != j) ? T() : it->second;
	}

	class const_iterator2
	{
		const std::vector<Entry>* r; std::size_t row; std::size_t k;
	public:
		const_iterator2(const std::vector<Entry>* rr, std::size_t i, std::size_t kk) : r(rr), row(i), k(kk) {}
		const_iterator2& operator++() { ++k; return *this; }
		bool operator!=(const const_iterator2& o) const { return k != o.k; }
		bool operator==(const const_iterator2& o) const { return k == o.k; }
		size_type index1() const { return row; }
		size_type index2() const { return (*r)[k].first; }
		const T& operator*() const { return (*r)[k].second; }
	};
	class const_iterator1
	{
		const compressed_matrix* m; std::size_t i;
	public:
		const_iterator1(const compressed_matrix* mm, std::size_t ii) : m(mm), i(ii) {}
		const_iterator1& operator++() { ++i; return *this; }
		bool operator!=(const const_iterator1& o) const { return i != o.i; }
		bool operator==(const const_iterator1& o) const { return i == o.i; }
		size_type index1() const { return i; }
		const_iterator2 begin() const { return const_iterator2(&m->rows[i], i, 0); }
		const_iterator2 end() const { return const_iterator2(&m->rows[i], i, m->rows[i].size()); }
	};
	const_iterator1 begin1() const { return const_iterator1(this, 0); }
	const_iterator1 end1() const { return const_iterator1(this, n1); }
};

}}}
#endif
