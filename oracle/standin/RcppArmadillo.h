// oracle/standin/RcppArmadillo.h -- TEST INFRASTRUCTURE, not product code.
//
// A small stand-in for <RcppArmadillo.h> so that the reference's own hot-path
// translation units (RNG, ShuffledSet, PLS, PLSSimpls, PLSEvaluator,
// BICEvaluator, LMEvaluator, Chromosome, Logger, SingleThreadPopulation,
// MultiThreadedPopulation) compile UNMODIFIED from /root/reference/src in a
// container that has neither R, Rcpp, Armadillo nor BLAS/LAPACK.  Only the API
// subset those files touch is provided.  Semantics that
// matter for exactness:
//   * column-major dense storage, deep copies on copy construction/assignment;
//   * Mat::unsafe_col / Mat::col / Col::rows(a,b) ALIAS the parent's storage and
//     assignment into an alias writes in place (src/PLSSimpls.cpp:58,134 rely on it);
//   * solve() is a Householder-QR least-squares solve that throws
//     std::runtime_error on a (numerically) rank-deficient system.
// The linear algebra is plain loops (vectorised by the compiler), NOT BLAS: every
// CPU-baseline number produced through this header must say so.
#ifndef GASELECT_B200_ORACLE_STANDIN_RCPPARMADILLO_H
#define GASELECT_B200_ORACLE_STANDIN_RCPPARMADILLO_H

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <functional>
#include <iostream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

// ----------------------------------------------------------------------------
// R C API surface
// ----------------------------------------------------------------------------
typedef void* SEXP;
#define RcppExport extern "C"
typedef enum { FALSE = 0, TRUE } Rboolean;

extern "C" {
void Rprintf(const char* fmt, ...);
void REprintf(const char* fmt, ...);
void R_FlushConsole(void);
void R_CheckUserInterrupt(void);
Rboolean R_ToplevelExec(void (*fun)(void*), void* data);
double R_pow_di(double x, int n);
}

// ----------------------------------------------------------------------------
// Rcpp surface
// ----------------------------------------------------------------------------
namespace Rcpp {

class exception : public std::exception {
public:
	exception(const char* msg, const char* /*file*/ = "", int /*line*/ = -1) : message(msg) {}
	virtual ~exception() throw() {}
	virtual const char* what() const throw() { return message.c_str(); }
private:
	std::string message;
};

class LogicalVector {
public:
	LogicalVector() {}
	LogicalVector(std::size_t n, bool v = false) : data(n, v ? 1 : 0) {}
	void push_back(bool v) { data.push_back(v ? 1 : 0); }
	int& operator[](std::size_t i) { return data[i]; }
	const int& operator[](std::size_t i) const { return data[i]; }
	std::size_t size() const { return data.size(); }
private:
	std::vector<int> data;
};

static std::ostream& Rcout = std::cout;

} // namespace Rcpp

// ----------------------------------------------------------------------------
// Armadillo surface
// ----------------------------------------------------------------------------
namespace arma {

typedef unsigned long long uword;

template <class T> class Mat;
template <class T> class Col;
template <class T> class Row;

// contiguous, aliasing view of `n_elem` elements (a column, or a row range of a column vector)
template <class T>
struct subview_col {
	T* mem;
	uword n_elem;
	uword n_rows;
	subview_col(T* m, uword n) : mem(m), n_elem(n), n_rows(n) {}
	subview_col(const subview_col& o) : mem(o.mem), n_elem(o.n_elem), n_rows(o.n_rows) {}
	const T* memptr() const { return mem; }
	T* memptr() { return mem; }
	T operator[](uword i) const { return mem[i]; }
	// element-wise copy INTO the viewed storage
	subview_col& operator=(const subview_col& o) {
		if (o.n_elem != n_elem) throw std::logic_error("standin: subview size mismatch");
		std::memmove(mem, o.mem, n_elem * sizeof(T));
		return *this;
	}
	subview_col& operator=(const Col<T>& o);
	subview_col& operator+=(T s) { for (uword i = 0; i < n_elem; ++i) mem[i] += s; return *this; }
	subview_col& operator-=(T s) { for (uword i = 0; i < n_elem; ++i) mem[i] -= s; return *this; }
};

template <class T>
struct op_trans {
	const Mat<T>& m;
	explicit op_trans(const Mat<T>& mm) : m(mm) {}
};

template <class T>
struct each_row_proxy {
	Mat<T>& m;
	explicit each_row_proxy(Mat<T>& mm) : m(mm) {}
	void operator-=(const Row<T>& r);
};

template <class T>
class Mat {
public:
	uword n_rows;
	uword n_cols;
	uword n_elem;

protected:
	T* mem;
	bool owns; // false: aliases foreign storage (unsafe_col)

	void alloc(uword r, uword c) {
		n_rows = r; n_cols = c; n_elem = r * c;
		mem = (n_elem > 0) ? static_cast<T*>(std::malloc(n_elem * sizeof(T))) : NULL;
		owns = true;
	}
	void release() {
		if (owns && mem) std::free(mem);
		mem = NULL;
	}
	struct alias_tag {};
	Mat(alias_tag, T* m, uword r, uword c) : n_rows(r), n_cols(c), n_elem(r * c), mem(m), owns(false) {}

public:
	Mat() : n_rows(0), n_cols(0), n_elem(0), mem(NULL), owns(true) {}
	Mat(uword r, uword c) { alloc(r, c); }
	Mat(const Mat& o) {
		alloc(o.n_rows, o.n_cols);
		if (n_elem) std::memcpy(mem, o.mem, n_elem * sizeof(T));
	}
	Mat(Mat&& o) {
		if (o.owns) {
			n_rows = o.n_rows; n_cols = o.n_cols; n_elem = o.n_elem; mem = o.mem; owns = true;
			o.mem = NULL; o.n_rows = o.n_cols = o.n_elem = 0;
		} else {
			alloc(o.n_rows, o.n_cols);
			if (n_elem) std::memcpy(mem, o.mem, n_elem * sizeof(T));
		}
	}
	~Mat() { release(); }

	Mat& operator=(const Mat& o) {
		if (this == &o) return *this;
		if (!owns) {
			if (o.n_elem != n_elem) throw std::logic_error("standin: assignment into alias with different size");
			if (n_elem) std::memmove(mem, o.mem, n_elem * sizeof(T));
			return *this;
		}
		if (o.n_elem != n_elem) { release(); alloc(o.n_rows, o.n_cols); }
		n_rows = o.n_rows; n_cols = o.n_cols;
		if (n_elem) std::memcpy(mem, o.mem, n_elem * sizeof(T));
		return *this;
	}
	Mat& operator=(Mat&& o) {
		if (this == &o) return *this;
		if (owns && o.owns) {
			std::swap(mem, o.mem);
			std::swap(n_rows, o.n_rows); std::swap(n_cols, o.n_cols); std::swap(n_elem, o.n_elem);
			return *this;
		}
		return (*this = static_cast<const Mat&>(o));
	}

	void set_size(uword r, uword c) {
		if (!owns) { if (r * c != n_elem) throw std::logic_error("standin: resize of alias"); n_rows = r; n_cols = c; return; }
		if (r * c != n_elem) { release(); alloc(r, c); } else { n_rows = r; n_cols = c; }
	}
	Mat& zeros(uword r, uword c) { set_size(r, c); if (n_elem) std::memset(mem, 0, n_elem * sizeof(T)); return *this; }
	Mat& zeros() { if (n_elem) std::memset(mem, 0, n_elem * sizeof(T)); return *this; }
	Mat& ones() { for (uword i = 0; i < n_elem; ++i) mem[i] = T(1); return *this; }

	T* memptr() { return mem; }
	const T* memptr() const { return mem; }
	T& operator[](uword i) { return mem[i]; }
	const T& operator[](uword i) const { return mem[i]; }
	T& operator()(uword i) { return mem[i]; }
	const T& operator()(uword i) const { return mem[i]; }
	T& operator()(uword r, uword c) { return mem[r + c * n_rows]; }
	const T& operator()(uword r, uword c) const { return mem[r + c * n_rows]; }
	T* begin() { return mem; }
	T* end() { return mem + n_elem; }
	const T* begin() const { return mem; }
	const T* end() const { return mem + n_elem; }
	uword size() const { return n_elem; }

	subview_col<T> col(uword c) { return subview_col<T>(mem + c * n_rows, n_rows); }
	const subview_col<T> col(uword c) const { return subview_col<T>(const_cast<T*>(mem) + c * n_rows, n_rows); }
	Col<T> unsafe_col(uword c);

	Mat cols(const Col<uword>& idx) const;
	Mat rows(const Col<uword>& idx) const;

	op_trans<T> t() const { return op_trans<T>(*this); }
	each_row_proxy<T> each_row() { return each_row_proxy<T>(*this); }

	void insert_cols(uword pos, const Mat& x) {
		if (n_elem > 0 && x.n_rows != n_rows) throw std::logic_error("standin: insert_cols row mismatch");
		const uword r = x.n_rows;
		Mat out(r, n_cols + x.n_cols);
		if (pos) std::memcpy(out.mem, mem, pos * r * sizeof(T));
		if (x.n_elem) std::memcpy(out.mem + pos * r, x.mem, x.n_elem * sizeof(T));
		if (n_cols > pos) std::memcpy(out.mem + (pos + x.n_cols) * r, mem + pos * r, (n_cols - pos) * r * sizeof(T));
		*this = std::move(out);
	}

	Mat& operator+=(T s) { for (uword i = 0; i < n_elem; ++i) mem[i] += s; return *this; }
	Mat& operator-=(T s) { for (uword i = 0; i < n_elem; ++i) mem[i] -= s; return *this; }
	Mat& operator/=(T s) { for (uword i = 0; i < n_elem; ++i) mem[i] /= s; return *this; }
	Mat& operator*=(T s) { for (uword i = 0; i < n_elem; ++i) mem[i] *= s; return *this; }

	template <class U> friend class Col;
	template <class U> friend class Row;
};

template <class T>
class Col : public Mat<T> {
	typedef Mat<T> base;
public:
	Col() : base() { this->n_cols = 1; }
	explicit Col(uword n) : base(n, 1) {}
	Col(const Col& o) : base(static_cast<const base&>(o)) {}
	Col(Col&& o) : base(static_cast<base&&>(o)) { this->n_cols = 1; }
	Col(const base& m) : base(m) {                       // a one-column matrix used as a vector
		this->n_rows = this->n_elem; this->n_cols = 1;
	}
	Col(const subview_col<T>& s) : base(s.n_elem, 1) { if (s.n_elem) std::memcpy(this->mem, s.mem, s.n_elem * sizeof(T)); }
	// aliasing constructor (unsafe_col)
	Col(T* m, uword n, bool /*alias*/) : base(typename base::alias_tag(), m, n, 1) {}

	Col& operator=(const Col& o) { base::operator=(static_cast<const base&>(o)); this->n_rows = this->n_elem; this->n_cols = 1; return *this; }
	Col& operator=(Col&& o) { base::operator=(static_cast<base&&>(o)); this->n_rows = this->n_elem; this->n_cols = 1; return *this; }
	Col& operator=(const base& o) { base::operator=(o); this->n_rows = this->n_elem; this->n_cols = 1; return *this; }

	void set_size(uword n) { base::set_size(n, 1); }
	Col& zeros(uword n) { base::zeros(n, 1); return *this; }
	Col& zeros() { base::zeros(); return *this; }
	Col& ones() { base::ones(); return *this; }

	// size-preserving resize
	void resize(uword n) {
		if (n == this->n_elem) return;
		Col out(n);
		const uword keep = (n < this->n_elem) ? n : this->n_elem;
		if (keep) std::memcpy(out.mem, this->mem, keep * sizeof(T));
		if (n > keep) std::memset(out.mem + keep, 0, (n - keep) * sizeof(T));
		*this = std::move(out);
	}
	// insert `n` zero rows before row `pos`
	void insert_rows(uword pos, uword n) {
		Col out(this->n_elem + n);
		if (pos) std::memcpy(out.mem, this->mem, pos * sizeof(T));
		std::memset(out.mem + pos, 0, n * sizeof(T));
		if (this->n_elem > pos) std::memcpy(out.mem + pos + n, this->mem + pos, (this->n_elem - pos) * sizeof(T));
		*this = std::move(out);
	}

	subview_col<T> rows(uword a, uword b) { return subview_col<T>(this->mem + a, b - a + 1); }
	const subview_col<T> rows(uword a, uword b) const { return subview_col<T>(const_cast<T*>(this->mem) + a, b - a + 1); }
	Col rows(const Col<uword>& idx) const {
		Col out(idx.n_elem);
		for (uword i = 0; i < idx.n_elem; ++i) out.mem[i] = this->mem[idx[i]];
		return out;
	}

	Col& operator+=(T s) { base::operator+=(s); return *this; }
	Col& operator-=(T s) { base::operator-=(s); return *this; }
	Col& operator/=(T s) { base::operator/=(s); return *this; }
	Col& operator*=(T s) { base::operator*=(s); return *this; }
	Col& operator-=(const Col& o) {
		T* __restrict a = this->mem; const T* __restrict b = o.mem;
		for (uword i = 0; i < this->n_elem; ++i) a[i] -= b[i];
		return *this;
	}
	Col& operator+=(const Col& o) {
		T* __restrict a = this->mem; const T* __restrict b = o.mem;
		for (uword i = 0; i < this->n_elem; ++i) a[i] += b[i];
		return *this;
	}
};

template <class T>
class Row : public Mat<T> {
	typedef Mat<T> base;
public:
	Row() : base() { this->n_rows = 1; }
	explicit Row(uword n) : base(1, n) {}
	Row(const Row& o) : base(static_cast<const base&>(o)) {}
	Row(Row&& o) : base(static_cast<base&&>(o)) { this->n_rows = 1; }
	Row& operator=(const Row& o) { base::operator=(static_cast<const base&>(o)); this->n_cols = this->n_elem; this->n_rows = 1; return *this; }
	Row& operator=(Row&& o) { base::operator=(static_cast<base&&>(o)); this->n_cols = this->n_elem; this->n_rows = 1; return *this; }
};

typedef Mat<double> mat;
typedef Col<double> vec;
typedef Col<double> colvec;
typedef Row<double> rowvec;
typedef Col<uword> uvec;

template <class T>
inline subview_col<T>& subview_col<T>::operator=(const Col<T>& o) {
	if (o.n_elem != n_elem) throw std::logic_error("standin: subview size mismatch");
	std::memmove(mem, o.memptr(), n_elem * sizeof(T));
	return *this;
}

template <class T>
inline Col<T> Mat<T>::unsafe_col(uword c) { return Col<T>(mem + c * n_rows, n_rows, true); }

template <class T>
inline Mat<T> Mat<T>::cols(const Col<uword>& idx) const {
	Mat<T> out(n_rows, idx.n_elem);
	for (uword j = 0; j < idx.n_elem; ++j) {
		std::memcpy(out.mem + j * n_rows, mem + idx[j] * n_rows, n_rows * sizeof(T));
	}
	return out;
}

template <class T>
inline Mat<T> Mat<T>::rows(const Col<uword>& idx) const {
	const uword r = idx.n_elem;
	Mat<T> out(r, n_cols);
	const uword* __restrict ix = idx.memptr();
	for (uword j = 0; j < n_cols; ++j) {
		const T* __restrict src = mem + j * n_rows;
		T* __restrict dst = out.mem + j * r;
		for (uword i = 0; i < r; ++i) dst[i] = src[ix[i]];
	}
	return out;
}

template <class T>
inline void each_row_proxy<T>::operator-=(const Row<T>& r) {
	for (uword j = 0; j < m.n_cols; ++j) {
		T* __restrict c = m.memptr() + j * m.n_rows;
		const T v = r[j];
		for (uword i = 0; i < m.n_rows; ++i) c[i] -= v;
	}
}

// ---- reductions -------------------------------------------------------------
namespace detail {
inline double dot_raw(const double* __restrict a, const double* __restrict b, uword n) {
	double s = 0.0;
#pragma omp simd reduction(+ : s)
	for (uword i = 0; i < n; ++i) s += a[i] * b[i];
	return s;
}
inline double sum_raw(const double* __restrict a, uword n) {
	double s = 0.0;
#pragma omp simd reduction(+ : s)
	for (uword i = 0; i < n; ++i) s += a[i];
	return s;
}
} // namespace detail

template <class A, class B>
inline double dot(const A& a, const B& b) { return detail::dot_raw(a.memptr(), b.memptr(), a.n_elem); }

inline double accu(const Mat<double>& a) { return detail::sum_raw(a.memptr(), a.n_elem); }
inline double sum(const Col<double>& a) { return detail::sum_raw(a.memptr(), a.n_elem); }
inline double mean(const Col<double>& a) { return detail::sum_raw(a.memptr(), a.n_elem) / static_cast<double>(a.n_elem); }
inline Row<double> mean(const Mat<double>& a) {
	Row<double> out(a.n_cols);
	for (uword j = 0; j < a.n_cols; ++j) {
		out[j] = detail::sum_raw(a.memptr() + j * a.n_rows, a.n_rows) / static_cast<double>(a.n_rows);
	}
	return out;
}
inline double norm(const Col<double>& a, int /*p == 2*/) { return std::sqrt(detail::dot_raw(a.memptr(), a.memptr(), a.n_elem)); }
inline Col<double> square(const Col<double>& a) {
	Col<double> out(a.n_elem);
	for (uword i = 0; i < a.n_elem; ++i) out[i] = a[i] * a[i];
	return out;
}
// norm_type 0: divide by N-1; 1: divide by N
inline double var(const Col<double>& a, int norm_type = 0) {
	const double m = mean(a);
	double ss = 0.0;
	for (uword i = 0; i < a.n_elem; ++i) { const double d = a[i] - m; ss += d * d; }
	return ss / static_cast<double>(norm_type == 1 ? a.n_elem : a.n_elem - 1);
}

template <class T>
inline Col<T> sort(const subview_col<T>& s) { Col<T> out(s); std::sort(out.begin(), out.end()); return out; }
template <class T>
inline Col<T> sort(const Col<T>& s) { Col<T> out(s); std::sort(out.begin(), out.end()); return out; }

template <class T>
inline Col<T> join_cols(const subview_col<T>& a, const subview_col<T>& b) {
	Col<T> out(a.n_elem + b.n_elem);
	if (a.n_elem) std::memcpy(out.memptr(), a.mem, a.n_elem * sizeof(T));
	if (b.n_elem) std::memcpy(out.memptr() + a.n_elem, b.mem, b.n_elem * sizeof(T));
	return out;
}

// ---- element-wise vector algebra -------------------------------------------------
inline Col<double> operator-(const Col<double>& a, const Col<double>& b) {
	Col<double> out(a.n_elem);
	const double* __restrict pa = a.memptr(); const double* __restrict pb = b.memptr(); double* __restrict po = out.memptr();
	for (uword i = 0; i < a.n_elem; ++i) po[i] = pa[i] - pb[i];
	return out;
}
inline Col<double> operator-(const Col<double>& a, double s) {
	Col<double> out(a.n_elem);
	for (uword i = 0; i < a.n_elem; ++i) out[i] = a[i] - s;
	return out;
}
inline Col<double> operator*(const Col<double>& a, double s) {
	Col<double> out(a.n_elem);
	for (uword i = 0; i < a.n_elem; ++i) out[i] = a[i] * s;
	return out;
}
inline Col<double> operator*(const subview_col<double>& a, double s) {
	Col<double> out(a.n_elem);
	for (uword i = 0; i < a.n_elem; ++i) out[i] = a.mem[i] * s;
	return out;
}
inline Col<double> operator/(const Col<double>& a, double s) {
	Col<double> out(a.n_elem);
	for (uword i = 0; i < a.n_elem; ++i) out[i] = a[i] / s;
	return out;
}

// ---- matrix-vector products ------------------------------------------------------
namespace detail {
// y = A x, column-major A (r x c): accumulate scaled columns
inline Col<double> gemv_n(const Mat<double>& A, const double* __restrict x) {
	const uword r = A.n_rows, c = A.n_cols;
	Col<double> y(r);
	double* __restrict py = y.memptr();
	for (uword i = 0; i < r; ++i) py[i] = 0.0;
	for (uword j = 0; j < c; ++j) {
		const double* __restrict a = A.memptr() + j * r;
		const double xj = x[j];
		for (uword i = 0; i < r; ++i) py[i] += a[i] * xj;
	}
	return y;
}
// y = A' x: one dot product per column
inline Col<double> gemv_t(const Mat<double>& A, const double* __restrict x) {
	const uword r = A.n_rows, c = A.n_cols;
	Col<double> y(c);
	for (uword j = 0; j < c; ++j) y[j] = dot_raw(A.memptr() + j * r, x, r);
	return y;
}
} // namespace detail

inline Col<double> operator*(const Mat<double>& A, const Col<double>& x) { return detail::gemv_n(A, x.memptr()); }
inline Col<double> operator*(const Mat<double>& A, const subview_col<double>& x) { return detail::gemv_n(A, x.mem); }
inline Col<double> operator*(const op_trans<double>& At, const Col<double>& x) { return detail::gemv_t(At.m, x.memptr()); }

// ---- least squares ---------------------------------------------------------------
// min ||A b - y||_2 by Householder QR (no pivoting).  Throws std::runtime_error when a
// diagonal element of R is negligible relative to the largest one (rank deficiency).
inline Col<double> solve(const Mat<double>& Ain, const Col<double>& yin) {
	const uword n = Ain.n_rows, m = Ain.n_cols;
	if (n < m) throw std::runtime_error("solve(): underdetermined system not supported by the stand-in");
	Mat<double> A(Ain);
	Col<double> y(yin);
	std::vector<double> rdiag(m);
	double maxdiag = 0.0;
	for (uword k = 0; k < m; ++k) {
		double* __restrict ak = A.memptr() + k * n;
		double nrm = 0.0;
		for (uword i = k; i < n; ++i) nrm += ak[i] * ak[i];
		nrm = std::sqrt(nrm);
		if (nrm == 0.0) throw std::runtime_error("solve(): rank deficient");
		const double alpha = (ak[k] > 0.0) ? -nrm : nrm;
		// v = x - alpha e1 (stored in place), beta = 2 / v'v
		ak[k] -= alpha;
		double vtv = 0.0;
		for (uword i = k; i < n; ++i) vtv += ak[i] * ak[i];
		if (vtv > 0.0) {
			for (uword j = k + 1; j < m; ++j) {
				double* __restrict aj = A.memptr() + j * n;
				double s = 0.0;
				for (uword i = k; i < n; ++i) s += ak[i] * aj[i];
				s = 2.0 * s / vtv;
				for (uword i = k; i < n; ++i) aj[i] -= s * ak[i];
			}
			double s = 0.0;
			for (uword i = k; i < n; ++i) s += ak[i] * y[i];
			s = 2.0 * s / vtv;
			for (uword i = k; i < n; ++i) y[i] -= s * ak[i];
		}
		rdiag[k] = alpha;
		if (std::fabs(alpha) > maxdiag) maxdiag = std::fabs(alpha);
	}
	for (uword k = 0; k < m; ++k) {
		if (std::fabs(rdiag[k]) <= maxdiag * 1e-13 * static_cast<double>(n)) throw std::runtime_error("solve(): rank deficient");
	}
	Col<double> b(m);
	for (uword kk = m; kk-- > 0;) {
		double s = y[kk];
		for (uword j = kk + 1; j < m; ++j) s -= A(kk, j) * b[j];
		b[kk] = s / rdiag[kk];
	}
	return b;
}

} // namespace arma

#endif
