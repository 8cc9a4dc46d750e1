// TEST INFRASTRUCTURE ONLY (oracle). Never included by the product path.
//
// A self-written linear-algebra binding for the *unmodified* QuantPDE headers.
// QuantPDE/Core:6-9 skips its own Eigen binding when QUANT_PDE_BOUND is already
// defined, so any header that supplies the names below may stand in for
// QuantPDE/src/Bindings/Eigen.hpp.  Eigen 3 is not installed in this image
// (no network), hence this binding.  It provides exactly the API surface the
// reference headers use:
//
//   Index, Vector, Matrix (row-major sparse), IntegerVector, Entry,
//   LinearSolver / SparseLUSolver / BiCGSTABSolver     (Bindings/Eigen.hpp:33-264)
//   Eigen::Success, Eigen::Triplet, Eigen::FFT<Real>   (BlackScholes.hpp:331, Map.hpp:178)
//
// All stepping / operator / penalty / event logic executed by programs built on
// this file is the reference's own code from /root/reference.  Only the
// arithmetic kernels underneath (sparse expression evaluation, the pivoted LU,
// BiCGSTAB+ILUT, the FFT) are written here, imitating Eigen's documented
// policies: natural column order, row partial pivoting with the diagonal
// preferred on ties, row-major SpMV accumulating in stored column order.
#ifndef QPDE_ORACLE_LINALG_BINDING_HPP
#define QPDE_ORACLE_LINALG_BINDING_HPP

#define QUANT_PDE_BOUND

// Eigen's headers pull in a large part of the standard library; the reference
// headers silently rely on that (e.g. std::bind in Metafunctions.hpp:273).
#include <algorithm>
#include <array>
#include <functional>
#include <iostream>
#include <limits>
#include <memory>
#include <tuple>
#include <type_traits>
#include <cassert>
#include <cmath>
#include <complex>
#include <cstddef>
#include <cstring>
#include <utility>
#include <vector>

namespace Eigen {

enum ComputationInfo { Success = 0, NumericalIssue = 1, NoConvergence = 2 };

template <typename Scalar>
class Triplet {
protected:
	int m_row, m_col;
	Scalar m_value;
public:
	Triplet() : m_row(0), m_col(0), m_value(0) {}
	Triplet(int i, int j, const Scalar &v = Scalar(0))
			: m_row(i), m_col(j), m_value(v) {}
	int row() const { return m_row; }
	int col() const { return m_col; }
	const Scalar &value() const { return m_value; }
};

// Complex FFT with Eigen::FFT's default flags: unscaled forward transform of
// a real sequence returning the full spectrum, inverse scaled by 1/n.
template <typename Scalar>
class FFT {
	typedef std::complex<Scalar> C;

	static void transform(std::vector<C> &a, bool inverse) {
		const size_t n = a.size();
		if(n <= 1) return;
		if((n & (n - 1)) != 0) {
			// Not a power of two: plain DFT
			std::vector<C> out(n);
			const Scalar sgn = inverse ? 1. : -1.;
			for(size_t k = 0; k < n; ++k) {
				C acc(0., 0.);
				for(size_t j = 0; j < n; ++j) {
					const Scalar ang = sgn * 2. * M_PI
						* (Scalar)((k * j) % n) / (Scalar)n;
					acc += a[j] * C(std::cos(ang), std::sin(ang));
				}
				out[k] = acc;
			}
			a.swap(out);
			return;
		}
		// Bit reversal
		for(size_t i = 1, j = 0; i < n; ++i) {
			size_t bit = n >> 1;
			for(; j & bit; bit >>= 1) j ^= bit;
			j ^= bit;
			if(i < j) std::swap(a[i], a[j]);
		}
		// Twiddles
		std::vector<C> w(n / 2);
		const Scalar sgn = inverse ? 1. : -1.;
		for(size_t k = 0; k < n / 2; ++k) {
			const Scalar ang = sgn * 2. * M_PI * (Scalar)k / (Scalar)n;
			w[k] = C(std::cos(ang), std::sin(ang));
		}
		for(size_t len = 2; len <= n; len <<= 1) {
			const size_t half = len >> 1, stride = n / len;
			for(size_t i = 0; i < n; i += len) {
				for(size_t k = 0; k < half; ++k) {
					const C u = a[i + k];
					const C t = a[i + k + half] * w[k * stride];
					a[i + k] = u + t;
					a[i + k + half] = u - t;
				}
			}
		}
	}

public:
	void fwd(std::vector<C> &dst, const std::vector<Scalar> &src) {
		dst.resize(src.size());
		for(size_t i = 0; i < src.size(); ++i) dst[i] = C(src[i], 0.);
		transform(dst, false);
	}
	void fwd(std::vector<C> &dst, const std::vector<C> &src) {
		dst = src;
		transform(dst, false);
	}
	void inv(std::vector<Scalar> &dst, const std::vector<C> &src) {
		std::vector<C> tmp(src);
		transform(tmp, true);
		dst.resize(src.size());
		const Scalar s = 1. / (Scalar)src.size();
		for(size_t i = 0; i < src.size(); ++i) dst[i] = tmp[i].real() * s;
	}
	void inv(std::vector<C> &dst, const std::vector<C> &src) {
		dst = src;
		transform(dst, true);
		const Scalar s = 1. / (Scalar)src.size();
		for(size_t i = 0; i < dst.size(); ++i) dst[i] *= s;
	}
};

} // namespace Eigen

namespace QuantPDE {

typedef int Index;

////////////////////////////////////////////////////////////////////////////////
// Dense vector
////////////////////////////////////////////////////////////////////////////////

class Vector {
	std::vector<Real> v;
public:
	Vector() {}
	explicit Vector(Index n) : v((size_t)n) {}
	Vector(Index n, Real c) : v((size_t)n, c) {}

	static Vector Zero(Index n) { return Vector(n, 0.); }
	static Vector Ones(Index n) { return Vector(n, 1.); }
	static Vector Constant(Index n, Real c) { return Vector(n, c); }

	Index size() const { return (Index)v.size(); }
	Index rows() const { return (Index)v.size(); }
	Real *data() { return v.data(); }
	const Real *data() const { return v.data(); }
	Real &operator()(Index i) { return v[i]; }
	const Real &operator()(Index i) const { return v[i]; }
	Real &operator[](Index i) { return v[i]; }
	const Real &operator[](Index i) const { return v[i]; }
	void resize(Index n) { v.resize((size_t)n); }
	void setZero() { std::fill(v.begin(), v.end(), 0.); }

	Vector cwiseAbs() const {
		Vector r(size());
		for(Index i = 0; i < size(); ++i) r.v[i] = std::abs(v[i]);
		return r;
	}
	Vector cwiseQuotient(const Vector &o) const {
		Vector r(size());
		for(Index i = 0; i < size(); ++i) r.v[i] = v[i] / o.v[i];
		return r;
	}
	Vector cwiseProduct(const Vector &o) const {
		Vector r(size());
		for(Index i = 0; i < size(); ++i) r.v[i] = v[i] * o.v[i];
		return r;
	}
	Vector cwiseMax(const Vector &o) const {
		Vector r(size());
		for(Index i = 0; i < size(); ++i)
			r.v[i] = v[i] < o.v[i] ? o.v[i] : v[i];
		return r;
	}
	Vector cwiseMin(const Vector &o) const {
		Vector r(size());
		for(Index i = 0; i < size(); ++i)
			r.v[i] = o.v[i] < v[i] ? o.v[i] : v[i];
		return r;
	}
	Real maxCoeff() const {
		Real m = v[0];
		for(Index i = 1; i < size(); ++i) if(v[i] > m) m = v[i];
		return m;
	}
	Real minCoeff() const {
		Real m = v[0];
		for(Index i = 1; i < size(); ++i) if(v[i] < m) m = v[i];
		return m;
	}
	Real dot(const Vector &o) const {
		Real s = 0.;
		for(Index i = 0; i < size(); ++i) s += v[i] * o.v[i];
		return s;
	}
	Real squaredNorm() const { return dot(*this); }
	Real norm() const { return std::sqrt(squaredNorm()); }

	Vector &operator+=(const Vector &o) {
		for(Index i = 0; i < size(); ++i) v[i] += o.v[i];
		return *this;
	}
	Vector &operator-=(const Vector &o) {
		for(Index i = 0; i < size(); ++i) v[i] -= o.v[i];
		return *this;
	}
	Vector &operator*=(Real c) {
		for(Index i = 0; i < size(); ++i) v[i] *= c;
		return *this;
	}
	Vector &operator/=(Real c) {
		for(Index i = 0; i < size(); ++i) v[i] /= c;
		return *this;
	}
};

inline Vector operator+(const Vector &a, const Vector &b) {
	Vector r(a.size());
	for(Index i = 0; i < a.size(); ++i) r[i] = a[i] + b[i];
	return r;
}
inline Vector operator-(const Vector &a, const Vector &b) {
	Vector r(a.size());
	for(Index i = 0; i < a.size(); ++i) r[i] = a[i] - b[i];
	return r;
}
inline Vector operator-(const Vector &a) {
	Vector r(a.size());
	for(Index i = 0; i < a.size(); ++i) r[i] = -a[i];
	return r;
}
inline Vector operator*(const Vector &a, Real c) {
	Vector r(a.size());
	for(Index i = 0; i < a.size(); ++i) r[i] = a[i] * c;
	return r;
}
inline Vector operator*(Real c, const Vector &a) {
	Vector r(a.size());
	for(Index i = 0; i < a.size(); ++i) r[i] = c * a[i];
	return r;
}
inline Vector operator/(const Vector &a, Real c) {
	Vector r(a.size());
	for(Index i = 0; i < a.size(); ++i) r[i] = a[i] / c;
	return r;
}

////////////////////////////////////////////////////////////////////////////////
// Integer vector (only used for reserve sizes)
////////////////////////////////////////////////////////////////////////////////

class IntegerVector {
	std::vector<int> v;
public:
	IntegerVector() {}
	explicit IntegerVector(Index n) : v((size_t)n) {}
	static IntegerVector Constant(Index n, int c) {
		IntegerVector r(n);
		std::fill(r.v.begin(), r.v.end(), c);
		return r;
	}
	Index size() const { return (Index)v.size(); }
	int &operator()(Index i) { return v[i]; }
	int operator()(Index i) const { return v[i]; }
	int &operator[](Index i) { return v[i]; }
	int operator[](Index i) const { return v[i]; }
};

////////////////////////////////////////////////////////////////////////////////
// Row-major sparse matrix.  Rows are kept as sorted (column, value) runs in
// one flat array; "uncompressed" mode keeps slack per row for insert().
////////////////////////////////////////////////////////////////////////////////

class Matrix {
	Index nr, nc;
	std::vector<Index> start;  // nr + 1: beginning of each row's slot
	std::vector<Index> count;  // nr: entries used in each row
	std::vector<Index> col;
	std::vector<Real>  val;

	void growRow(Index i, Index extra) {
		// Re-lay the storage giving row i `extra` more slots
		std::vector<Index> nstart(nr + 1);
		Index total = 0;
		for(Index r = 0; r < nr; ++r) {
			nstart[r] = total;
			total += (start[r + 1] - start[r]) + (r == i ? extra : 0);
		}
		nstart[nr] = total;
		std::vector<Index> ncol((size_t)total);
		std::vector<Real>  nval((size_t)total);
		for(Index r = 0; r < nr; ++r) {
			for(Index k = 0; k < count[r]; ++k) {
				ncol[nstart[r] + k] = col[start[r] + k];
				nval[nstart[r] + k] = val[start[r] + k];
			}
		}
		start.swap(nstart); col.swap(ncol); val.swap(nval);
	}

public:
	Matrix() : nr(0), nc(0), start(1, 0) {}
	Matrix(Index r, Index c) : nr(r), nc(c), start((size_t)r + 1, 0),
			count((size_t)r, 0) {}

	Index rows() const { return nr; }
	Index cols() const { return nc; }
	Index outerSize() const { return nr; }
	Index nonZeros() const {
		Index s = 0;
		for(Index r = 0; r < nr; ++r) s += count[r];
		return s;
	}

	// Row access for the solvers in this binding
	Index rowCount(Index r) const { return count[r]; }
	const Index *rowCols(Index r) const { return col.data() + start[r]; }
	const Real  *rowVals(Index r) const { return val.data() + start[r]; }

	void resize(Index r, Index c) { *this = Matrix(r, c); }

	void setZero() {
		std::fill(count.begin(), count.end(), 0);
		std::fill(start.begin(), start.end(), 0);
		col.clear(); val.clear();
	}

	void setIdentity() {
		setZero();
		const Index n = nr < nc ? nr : nc;
		col.resize((size_t)n); val.resize((size_t)n);
		for(Index r = 0; r < nr; ++r) {
			start[r] = r < n ? r : n;
			count[r] = r < n ? 1 : 0;
			if(r < n) { col[r] = r; val[r] = 1.; }
		}
		start[nr] = n;
	}

	template <typename Sizes>
	void reserve(const Sizes &sizes) {
		std::vector<Index> nstart(nr + 1);
		Index total = 0;
		for(Index r = 0; r < nr; ++r) {
			nstart[r] = total;
			total += count[r] + (Index)sizes[r];
		}
		nstart[nr] = total;
		std::vector<Index> ncol((size_t)total);
		std::vector<Real>  nval((size_t)total);
		for(Index r = 0; r < nr; ++r) {
			for(Index k = 0; k < count[r]; ++k) {
				ncol[nstart[r] + k] = col[start[r] + k];
				nval[nstart[r] + k] = val[start[r] + k];
			}
		}
		start.swap(nstart); col.swap(ncol); val.swap(nval);
	}

	// Inserts a structurally new zero at (i, j) and returns a reference.
	Real &insert(Index i, Index j) {
		if(start[i] + count[i] >= start[i + 1]) growRow(i, 4);
		Index *c = col.data() + start[i];
		Real  *v = val.data() + start[i];
		Index k = count[i];
		while(k > 0 && c[k - 1] > j) {
			c[k] = c[k - 1]; v[k] = v[k - 1]; --k;
		}
		assert(k == 0 || c[k - 1] != j);
		c[k] = j; v[k] = 0.;
		++count[i];
		return v[k];
	}

	Real &coeffRef(Index i, Index j) {
		Index *c = col.data() + start[i];
		for(Index k = 0; k < count[i]; ++k) {
			if(c[k] == j) return val[start[i] + k];
		}
		return insert(i, j);
	}

	Real coeff(Index i, Index j) const {
		const Index *c = col.data() + start[i];
		for(Index k = 0; k < count[i]; ++k) {
			if(c[k] == j) return val[start[i] + k];
		}
		return 0.;
	}

	void makeCompressed() {
		Index total = 0;
		for(Index r = 0; r < nr; ++r) {
			const Index s = start[r];
			if(s != total) {
				for(Index k = 0; k < count[r]; ++k) {
					col[total + k] = col[s + k];
					val[total + k] = val[s + k];
				}
			}
			start[r] = total;
			total += count[r];
		}
		start[nr] = total;
		col.resize((size_t)total); val.resize((size_t)total);
	}

	template <typename It>
	void setFromTriplets(It begin, It end) {
		std::vector<std::vector<std::pair<Index, Real>>> rows((size_t)nr);
		for(It it = begin; it != end; ++it) {
			rows[it->row()].push_back(
					std::make_pair((Index)it->col(), (Real)it->value()));
		}
		setZero();
		Index total = 0;
		for(Index r = 0; r < nr; ++r) {
			auto &row = rows[r];
			std::stable_sort(row.begin(), row.end(),
				[] (const std::pair<Index, Real> &a,
				    const std::pair<Index, Real> &b) {
					return a.first < b.first;
				});
			start[r] = total;
			Index used = 0;
			for(size_t k = 0; k < row.size(); ++k) {
				if(used > 0 && col.back() == row[k].first) {
					val.back() += row[k].second; // duplicates are summed
				} else {
					col.push_back(row[k].first);
					val.push_back(row[k].second);
					++used;
				}
			}
			count[r] = used;
			total += used;
		}
		start[nr] = total;
	}

	// Builds a result row by row (already sorted appends)
	void appendRowStart(Index r) { start[r] = (Index)col.size(); count[r] = 0; }
	void append(Index r, Index j, Real x) {
		col.push_back(j); val.push_back(x); ++count[r];
	}
	void appendFinish() { start[nr] = (Index)col.size(); }

	Matrix &operator+=(const Matrix &o);
	Matrix &operator-=(const Matrix &o);
	Matrix &operator*=(Real c) {
		for(Index r = 0; r < nr; ++r)
			for(Index k = 0; k < count[r]; ++k) val[start[r] + k] *= c;
		return *this;
	}
	Matrix &operator/=(Real c) {
		for(Index r = 0; r < nr; ++r)
			for(Index k = 0; k < count[r]; ++k) val[start[r] + k] /= c;
		return *this;
	}
};

// Union-merge of two sparse matrices; entries present on one side only are
// copied (a + 0 and 0 + b are exact), as Eigen's sparse binary ops do.
template <typename Op>
inline Matrix sparseMerge(const Matrix &a, const Matrix &b, Op op) {
	assert(a.rows() == b.rows() && a.cols() == b.cols());
	Matrix r(a.rows(), a.cols());
	for(Index i = 0; i < a.rows(); ++i) {
		r.appendRowStart(i);
		const Index na = a.rowCount(i), nb = b.rowCount(i);
		const Index *ca = a.rowCols(i), *cb = b.rowCols(i);
		const Real  *va = a.rowVals(i), *vb = b.rowVals(i);
		Index p = 0, q = 0;
		while(p < na || q < nb) {
			if(q >= nb || (p < na && ca[p] < cb[q])) {
				r.append(i, ca[p], op(va[p], 0.)); ++p;
			} else if(p >= na || cb[q] < ca[p]) {
				r.append(i, cb[q], op(0., vb[q])); ++q;
			} else {
				r.append(i, ca[p], op(va[p], vb[q])); ++p; ++q;
			}
		}
	}
	r.appendFinish();
	return r;
}

inline Matrix operator+(const Matrix &a, const Matrix &b) {
	return sparseMerge(a, b, [] (Real x, Real y) { return x + y; });
}
inline Matrix operator-(const Matrix &a, const Matrix &b) {
	return sparseMerge(a, b, [] (Real x, Real y) { return x - y; });
}
inline Matrix &Matrix::operator+=(const Matrix &o) {
	*this = *this + o; return *this;
}
inline Matrix &Matrix::operator-=(const Matrix &o) {
	*this = *this - o; return *this;
}
inline Matrix operator-(const Matrix &a) {
	Matrix r(a); r *= -1.; return r;
}
inline Matrix operator*(const Matrix &a, Real c) { Matrix r(a); r *= c; return r; }
inline Matrix operator*(Real c, const Matrix &a) { Matrix r(a); r *= c; return r; }
inline Matrix operator/(const Matrix &a, Real c) { Matrix r(a); r /= c; return r; }

// Row-major sparse * dense vector: per row, accumulate from zero in stored
// column order.
inline Vector operator*(const Matrix &a, const Vector &x) {
	assert(a.cols() == x.size());
	Vector y(a.rows());
	for(Index i = 0; i < a.rows(); ++i) {
		const Index n = a.rowCount(i);
		const Index *c = a.rowCols(i);
		const Real  *v = a.rowVals(i);
		Real acc = 0.;
		for(Index k = 0; k < n; ++k) acc += v[k] * x[c[k]];
		y[i] = acc;
	}
	return y;
}

// Sparse * sparse (row-major Gustavson), result columns sorted.
inline Matrix operator*(const Matrix &a, const Matrix &b) {
	assert(a.cols() == b.rows());
	Matrix r(a.rows(), b.cols());
	std::vector<Real> acc((size_t)b.cols(), 0.);
	std::vector<char> used((size_t)b.cols(), 0);
	std::vector<Index> touched;
	for(Index i = 0; i < a.rows(); ++i) {
		touched.clear();
		const Index na = a.rowCount(i);
		const Index *ca = a.rowCols(i);
		const Real  *va = a.rowVals(i);
		for(Index p = 0; p < na; ++p) {
			const Index k = ca[p];
			const Index nb = b.rowCount(k);
			const Index *cb = b.rowCols(k);
			const Real  *vb = b.rowVals(k);
			for(Index q = 0; q < nb; ++q) {
				const Index j = cb[q];
				if(!used[j]) { used[j] = 1; acc[j] = 0.; touched.push_back(j); }
				acc[j] += va[p] * vb[q];
			}
		}
		std::sort(touched.begin(), touched.end());
		r.appendRowStart(i);
		for(size_t t = 0; t < touched.size(); ++t) {
			r.append(i, touched[t], acc[touched[t]]);
			used[touched[t]] = 0;
		}
	}
	r.appendFinish();
	return r;
}

////////////////////////////////////////////////////////////////////////////////

class Entry : public Eigen::Triplet<Real> {
public:
	Entry(Index i, Index j, Real v) : Triplet(i, j, v) {}
	Real &value() { return this->m_value; }
};

////////////////////////////////////////////////////////////////////////////////
// Direct solver: banded LU with row partial pivoting, natural column order.
// Pivot policy follows Eigen::SparseLU's default (diagpivotthresh = 1): the
// diagonal entry is kept iff it is at least as large in magnitude as every
// candidate below it in the column; otherwise the largest candidate is swapped
// in (first maximum wins).  Storage and update order are LAPACK-dgbtf2-like.
////////////////////////////////////////////////////////////////////////////////

class SparseLU {
	Index n, kl, ku, ldab;
	std::vector<Real> ab;     // (2 kl + ku + 1) x n, column major
	std::vector<Index> piv;
	Eigen::ComputationInfo status;

	Real &at(Index i, Index j) { return ab[(size_t)j * ldab + (kl + ku + i - j)]; }

public:
	SparseLU() : n(0), kl(0), ku(0), ldab(0), status(Eigen::Success) {}

	void analyzePattern(const Matrix &A) {
		n = A.rows();
		kl = ku = 0;
		for(Index i = 0; i < n; ++i) {
			const Index cnt = A.rowCount(i);
			const Index *c = A.rowCols(i);
			for(Index k = 0; k < cnt; ++k) {
				if(i - c[k] > kl) kl = i - c[k];
				if(c[k] - i > ku) ku = c[k] - i;
			}
		}
		ldab = 2 * kl + ku + 1;
	}

	void factorize(const Matrix &A) {
		status = Eigen::Success;
		ab.assign((size_t)ldab * n, 0.);
		piv.assign((size_t)n, 0);
		for(Index i = 0; i < n; ++i) {
			const Index cnt = A.rowCount(i);
			const Index *c = A.rowCols(i);
			const Real *v = A.rowVals(i);
			for(Index k = 0; k < cnt; ++k) at(i, c[k]) = v[k];
		}
		const Index kv = kl + ku; // superdiagonals of U incl. fill
		for(Index j = 0; j < n; ++j) {
			const Index km = std::min(kl, n - 1 - j);
			// Pivot search in column j, rows j..j+km
			Index p = j;
			Real best = std::abs(at(j, j));
			for(Index i = j + 1; i <= j + km; ++i) {
				const Real c = std::abs(at(i, j));
				if(c > best) { best = c; p = i; }
			}
			piv[j] = p;
			if(best == 0.) { status = Eigen::NumericalIssue; continue; }
			const Index jend = std::min(j + kv, n - 1);
			if(p != j) {
				for(Index c = j; c <= jend; ++c) std::swap(at(j, c), at(p, c));
			}
			const Real pivot = at(j, j);
			for(Index i = j + 1; i <= j + km; ++i) at(i, j) /= pivot;
			for(Index c = j + 1; c <= jend; ++c) {
				const Real u = at(j, c);
				if(u == 0.) continue;
				for(Index i = j + 1; i <= j + km; ++i) {
					at(i, c) -= at(i, j) * u;
				}
			}
		}
	}

	void compute(const Matrix &A) { analyzePattern(A); factorize(A); }

	Eigen::ComputationInfo info() const { return status; }

	Vector solve(const Vector &b) {
		Vector x(b);
		// L y = P b
		for(Index j = 0; j < n; ++j) {
			const Index km = std::min(kl, n - 1 - j);
			const Index p = piv[j];
			if(p != j) std::swap(x[j], x[p]);
			const Real xj = x[j];
			for(Index i = j + 1; i <= j + km; ++i) x[i] -= at(i, j) * xj;
		}
		// U x = y
		const Index kv = kl + ku;
		for(Index j = n - 1; j >= 0; --j) {
			x[j] /= at(j, j);
			const Real xj = x[j];
			const Index i0 = std::max(0, j - kv);
			for(Index i = i0; i < j; ++i) x[i] -= at(i, j) * xj;
		}
		return x;
	}
};

////////////////////////////////////////////////////////////////////////////////
// BiCGSTAB with an ILUT preconditioner (Eigen defaults: tolerance = machine
// epsilon, at most 2 n iterations; ILUT drop tolerance ~1e-12, fill factor 10).
////////////////////////////////////////////////////////////////////////////////

class IncompleteLUT {
	Index n;
	// L (unit lower, strict part) and U (incl. diagonal) stored by rows
	std::vector<std::vector<std::pair<Index, Real>>> L, U;
	Real droptol;
	int fillfactor;
public:
	IncompleteLUT() : n(0), droptol(1e-12), fillfactor(10) {}

	void compute(const Matrix &A) {
		n = A.rows();
		L.assign((size_t)n, std::vector<std::pair<Index, Real>>());
		U.assign((size_t)n, std::vector<std::pair<Index, Real>>());
		const Index nnz = A.nonZeros();
		const Index fill = (Index)((nnz * fillfactor) / (n > 0 ? n : 1)) + 1;
		std::vector<Real> w((size_t)n, 0.);
		std::vector<char> in((size_t)n, 0);
		std::vector<Index> idx;
		for(Index i = 0; i < n; ++i) {
			idx.clear();
			const Index cnt = A.rowCount(i);
			const Index *c = A.rowCols(i);
			const Real *v = A.rowVals(i);
			Real rownorm = 0.;
			for(Index k = 0; k < cnt; ++k) {
				w[c[k]] = v[k]; in[c[k]] = 1; idx.push_back(c[k]);
				rownorm += v[k] * v[k];
			}
			if(!in[i]) { w[i] = 0.; in[i] = 1; idx.push_back(i); }
			rownorm = std::sqrt(rownorm);
			const Real tol = droptol * rownorm;
			// Eliminate with previous rows in increasing column order
			std::sort(idx.begin(), idx.end());
			for(size_t p = 0; p < idx.size(); ++p) {
				const Index k = idx[p];
				if(k >= i) break;
				Real fact = w[k] / U[k][0].second; // diagonal first
				if(std::abs(fact) <= tol) { w[k] = 0.; continue; }
				w[k] = fact;
				const auto &urow = U[k];
				for(size_t q = 1; q < urow.size(); ++q) {
					const Index j = urow[q].first;
					if(!in[j]) {
						in[j] = 1; w[j] = 0.;
						// keep idx sorted
						auto pos = std::lower_bound(
							idx.begin(), idx.end(), j);
						const size_t off = pos - idx.begin();
						idx.insert(pos, j);
						(void) off;
					}
					w[j] -= fact * urow[q].second;
				}
			}
			// Gather L and U parts with dropping
			std::vector<std::pair<Index, Real>> lrow, urow;
			Real diag = w[i];
			for(size_t p = 0; p < idx.size(); ++p) {
				const Index j = idx[p];
				const Real x = w[j];
				in[j] = 0; w[j] = 0.;
				if(j == i) continue;
				if(std::abs(x) <= tol) continue;
				if(j < i) lrow.push_back(std::make_pair(j, x));
				else      urow.push_back(std::make_pair(j, x));
			}
			auto keepLargest = [&] (std::vector<std::pair<Index, Real>> &r) {
				if((Index)r.size() <= fill) return;
				std::nth_element(r.begin(), r.begin() + fill, r.end(),
					[] (const std::pair<Index, Real> &a,
					    const std::pair<Index, Real> &b) {
						return std::abs(a.second) > std::abs(b.second);
					});
				r.resize((size_t)fill);
				std::sort(r.begin(), r.end());
			};
			keepLargest(lrow);
			keepLargest(urow);
			if(diag == 0.) diag = std::sqrt(droptol) * rownorm + 1e-300;
			U[i].push_back(std::make_pair(i, diag));
			U[i].insert(U[i].end(), urow.begin(), urow.end());
			L[i].swap(lrow);
		}
	}

	Vector solve(const Vector &b) const {
		Vector x(b);
		for(Index i = 0; i < n; ++i) {
			Real s = x[i];
			for(const auto &e : L[i]) s -= e.second * x[e.first];
			x[i] = s;
		}
		for(Index i = n - 1; i >= 0; --i) {
			Real s = x[i];
			const auto &u = U[i];
			for(size_t q = 1; q < u.size(); ++q) s -= u[q].second * x[u[q].first];
			x[i] = s / u[0].second;
		}
		return x;
	}
};

class BiCGSTAB {
	const Matrix *A;
	Matrix Acopy;
	IncompleteLUT precond;
	Eigen::ComputationInfo status;
	Index its;
	Real tol;
public:
	BiCGSTAB() : A(nullptr), status(Eigen::Success), its(0),
			tol(std::numeric_limits<Real>::epsilon()) {}

	void compute(const Matrix &M) {
		Acopy = M; A = &Acopy;
		precond.compute(Acopy);
		status = Eigen::Success;
	}
	void setTolerance(Real t) { tol = t; }
	Eigen::ComputationInfo info() const { return status; }
	Index iterations() const { return its; }

	Vector solve(const Vector &b) {
		return solveWithGuess(b, Vector::Zero(b.size()));
	}

	// Follows the structure of Eigen's bicgstab(): preconditioned BiCGSTAB
	// with restart when rho becomes tiny relative to |r0|^2.
	Vector solveWithGuess(const Vector &rhs, const Vector &guess) {
		const Matrix &mat = *A;
		const Index n = mat.cols();
		Vector x(guess);
		const Index maxIters = 2 * n;
		Vector r = rhs - mat * x;
		Vector r0 = r;
		Real r0_sqnorm = r0.squaredNorm();
		const Real rhs_sqnorm = rhs.squaredNorm();
		its = 0;
		if(rhs_sqnorm == 0.) {
			x.setZero();
			status = Eigen::Success;
			return x;
		}
		Real rho = 1., alpha = 1., w = 1.;
		Vector v = Vector::Zero(n), p = Vector::Zero(n);
		Vector y(n), z(n), kt(n), ks(n), s(n), t(n);
		const Real tol2 = tol * tol * rhs_sqnorm;
		const Real eps2 = std::numeric_limits<Real>::epsilon()
				* std::numeric_limits<Real>::epsilon();
		Index i = 0, restarts = 0;
		while(r.squaredNorm() > tol2 && i < maxIters) {
			const Real rho_old = rho;
			rho = r0.dot(r);
			if(std::abs(rho) < eps2 * r0_sqnorm) {
				r = rhs - mat * x;
				r0 = r;
				rho = r0_sqnorm = r.squaredNorm();
				if(restarts++ == 0) i = 0;
			}
			const Real beta = (rho / rho_old) * (alpha / w);
			p = r + beta * (p - w * v);
			y = precond.solve(p);
			v = mat * y;
			alpha = rho / r0.dot(v);
			s = r - alpha * v;
			z = precond.solve(s);
			t = mat * z;
			const Real tmp = t.squaredNorm();
			if(tmp > 0.) w = t.dot(s) / tmp;
			else         w = 0.;
			x += alpha * y + w * z;
			r = s - w * t;
			++i;
		}
		its = i;
		status = (r.squaredNorm() <= tol2) ? Eigen::Success
				: Eigen::NoConvergence;
		return x;
	}
};

////////////////////////////////////////////////////////////////////////////////
// Solver classes with the interface of Bindings/Eigen.hpp:71-264
////////////////////////////////////////////////////////////////////////////////

class LinearSolver {
#ifdef QUANT_PDE_PERMISSIVE
public:
#else
protected:
#endif
	Matrix A;
	std::vector<size_t> its;
	virtual void initialize() = 0;
public:
	LinearSolver() noexcept {}
	virtual ~LinearSolver() {}
	LinearSolver(const LinearSolver &) = delete;
	LinearSolver &operator=(const LinearSolver &) & = delete;

	template <typename M>
	void initialize(M &&A) {
		this->A = std::forward<M>(A);
		initialize();
	}
	virtual Vector solve(const Vector &b, const Vector &guess) = 0;
	const std::vector<size_t> &iterations() const { return its; }
	const Matrix &matrix() const { return A; }
};

class SparseLUSolver : public LinearSolver {
	SparseLU solver;
	virtual void initialize() {
		solver.analyzePattern(A);
		solver.factorize(A);
		assert( solver.info() == Eigen::Success );
	}
public:
	SparseLUSolver() noexcept : LinearSolver() {}
	virtual Vector solve(const Vector &b, const Vector &) {
		return solver.solve(b);
	}
};

class BiCGSTABSolver : public LinearSolver {
#ifdef QUANT_PDE_PERMISSIVE
public:
#else
private:
#endif
	BiCGSTAB solver;
	virtual void initialize() {
		solver.compute(A);
		assert( solver.info() == Eigen::Success );
	}
public:
	BiCGSTABSolver() noexcept : LinearSolver() {}
	virtual Vector solve(const Vector &b, const Vector &guess) {
		Vector v = solver.solveWithGuess(b, guess);
		assert( solver.info() == Eigen::Success );
		its.push_back( (size_t) solver.iterations() );
		return v;
	}
};

} // namespace QuantPDE

#endif
