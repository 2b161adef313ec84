// TEST INFRASTRUCTURE -- never included by the product path.
//
// Header-only stand-in for <Rcpp.h> + <RcppArmadillo.h>, just wide enough to compile the
// UNMODIFIED reference sources
//     /root/reference/src/{full_gibbs,stickbreaking,collapsed_gibbs,collapsed_gibbs_dp,stephens,
//                          utils,my_lpsolve,RcppExports}.cpp
// where they lie (oracle/build_ref.sh -> oracle/_ref/libbmm_ref.so).  R, Rcpp and Armadillo are not in
// this image; the shim supplies only the types and operations those eight files use:
//   Rcpp : SEXP, RObject, List, Vector/Matrix (Numeric*, Integer*), Range, _, as<>, wrap, Rcout, stop,
//          RNGScope, traits::input_parameter, BEGIN_RCPP/END_RCPP, R_registerRoutines & friends,
//          R::rgamma / R::rbeta, rmultinom, unif_rand (routed to oracle/rrng.h)
//   arma : Mat/Col/Row/Cube (column-major, bounds-checked like a default Armadillo build), sub-views,
//          each_col, log, sum(.,0), accu, %, +, -, scalar *,/ , index_max, sort_index, zeros, fill::zeros
// What stays [memory] (third-party code that is not under /root/reference): the nmath
// generators behind rrng.h and Armadillo's accumulation order -- sum(X,0) and accu() use Armadillo's
// two-accumulator (even/odd) loop, restated from arrayops::accumulate / op_sum::apply_noalias_proxy.
// Matrices created without a fill are zero here (Armadillo leaves them uninitialised); that only shows
// where the reference itself returns uninitialised memory.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>
#include <climits>

#include "../rrng.h"

// =============================================================================================
// R C API surface
// =============================================================================================
#define NILSXP 0
#define LGLSXP 10
#define INTSXP 13
#define REALSXP 14
#define VECSXP 19

struct SEXPREC {
    int type = NILSXP;
    std::vector<int> dim;  // empty = plain vector
    std::vector<double> real;
    std::vector<int> integer;  // INTSXP and LGLSXP
    std::vector<std::string> names;
    std::vector<std::shared_ptr<SEXPREC>> elts;  // VECSXP
    size_t length() const {
        return type == REALSXP ? real.size() : type == VECSXP ? elts.size() : integer.size();
    }
};
typedef SEXPREC *SEXP;

typedef void *(*DL_FUNC)();
struct R_CallMethodDef {
    const char *name;
    DL_FUNC fun;
    int numArgs;
};
struct DllInfo {
    const R_CallMethodDef *call_entries = nullptr;
    bool dynamic_symbols = true;
};
#ifndef TRUE
#define TRUE 1
#define FALSE 0
#endif
inline int R_registerRoutines(DllInfo *dll, const void *, const R_CallMethodDef *call, const void *, const void *) {
    dll->call_entries = call;
    return 1;
}
inline int R_useDynamicSymbols(DllInfo *dll, int v) {
    int old = dll->dynamic_symbols;
    dll->dynamic_symbols = v != 0;
    return old;
}
#define RcppExport extern "C"

namespace bmm_shim {
// R's global RNG (GetRNGstate/PutRNGstate): one per thread so chains can run on every core
inline oracle::RRng &rng() {
    static thread_local oracle::RRng r;
    return r;
}
// objects handed across the SEXP boundary stay alive until the caller clears the arena
inline std::vector<std::shared_ptr<SEXPREC>> &arena() {
    static thread_local std::vector<std::shared_ptr<SEXPREC>> a;
    return a;
}
}  // namespace bmm_shim

inline double unif_rand() { return bmm_shim::rng().unif_rand(); }

// nmath rmultinom(n, prob, K, rN); the path only ever draws n = 1 (full_gibbs.cpp:133)
inline void rmultinom(int n, double *prob, int K, int *rN) {
    if (n != 1) throw std::runtime_error("shim rmultinom: only n = 1 is implemented");
    long double p_tot = 0.0L;
    for (int k = 0; k < K; ++k) {
        double pp = prob[k];
        if (!std::isfinite(pp) || pp < 0.0 || pp > 1.0) {
            rN[k] = INT_MIN;  // NA_INTEGER, then return (ML_ERR_ret_NAN)
            return;
        }
        p_tot += pp;
        rN[k] = 0;
    }
    if (std::fabs((double)(p_tot - 1.0L)) > 1e-7) throw std::runtime_error("rbinom: probability sum should be 1");
    for (int k = 0; k < K - 1; ++k) {
        if (prob[k] != 0.0) {
            double pp = (double)(prob[k] / p_tot);
            rN[k] = (pp < 1.0) ? bmm_shim::rng().rbinom1(pp) : n;
            n -= rN[k];
        } else {
            rN[k] = 0;
        }
        if (n <= 0) return;
        p_tot -= prob[k];
    }
    rN[K - 1] = n;
}

// options() look-ups of r-package/src/host.cpp: the stand-in has no options, every look-up is NULL (defaults apply)
inline SEXP Rf_install(const char *) { return nullptr; }
inline SEXP Rf_GetOption1(SEXP) { return nullptr; }
inline bool Rf_isNull(SEXP s) { return s == nullptr; }
inline double Rf_asReal(SEXP) { return 0.0; }
inline int Rf_asInteger(SEXP) { return 0; }

namespace R {
inline double rgamma(double shape, double scale) { return bmm_shim::rng().rgamma(shape, scale); }
inline double rbeta(double a, double b) { return bmm_shim::rng().rbeta(a, b); }
inline double unif_rand() { return bmm_shim::rng().unif_rand(); }
}  // namespace R

// =============================================================================================
// Armadillo surface
// =============================================================================================
namespace arma {

typedef unsigned long long uword;

namespace fill {
struct fill_zeros {};
static const fill_zeros zeros = fill_zeros();
}  // namespace fill

inline void arma_stop_bounds(const char *what) { throw std::out_of_range(std::string("arma shim: ") + what); }
inline void arma_stop_logic(const char *what) { throw std::logic_error(std::string("arma shim: ") + what); }

template <typename eT> class Mat;
template <typename eT> class Col;
template <typename eT> class Row;
template <typename eT> class Cube;
template <typename eT> class subview;
template <typename eT> class subview_each_col;

// a rectangular window of a matrix
template <typename eT>
class subview {
   public:
    Mat<eT> &m;
    const uword r0, c0, n_rows, n_cols, n_elem;
    subview(Mat<eT> &m_, uword r0_, uword c0_, uword nr, uword nc)
        : m(m_), r0(r0_), c0(c0_), n_rows(nr), n_cols(nc), n_elem(nr * nc) {}
    inline eT &at(uword i, uword j) { return m.mem[(r0 + i) + (c0 + j) * m.n_rows]; }
    inline const eT &at(uword i, uword j) const { return m.mem[(r0 + i) + (c0 + j) * m.n_rows]; }
    inline eT &operator()(uword i) {
        if (i >= n_elem) arma_stop_bounds("subview::operator(): index out of bounds");
        return n_rows == 1 ? at(0, i) : at(i % n_rows, i / n_rows);
    }
    inline eT operator()(uword i) const {
        if (i >= n_elem) arma_stop_bounds("subview::operator(): index out of bounds");
        return n_rows == 1 ? at(0, i) : at(i % n_rows, i / n_rows);
    }
    template <typename T2>
    void assign(const Mat<T2> &x) {
        if (x.n_rows != n_rows || x.n_cols != n_cols) arma_stop_logic("copy into submatrix: incompatible dimensions");
        if ((const void *)x.mem == (const void *)m.mem) {  // aliasing
            Mat<T2> tmp(x);
            for (uword j = 0; j < n_cols; ++j)
                for (uword i = 0; i < n_rows; ++i) at(i, j) = (eT)tmp.mem[i + j * n_rows];
            return;
        }
        for (uword j = 0; j < n_cols; ++j)
            for (uword i = 0; i < n_rows; ++i) at(i, j) = (eT)x.mem[i + j * n_rows];
    }
    subview &operator=(const Mat<eT> &x) {
        assign(x);
        return *this;
    }
    subview &operator=(const subview &x) {
        Mat<eT> tmp(x);
        assign(tmp);
        return *this;
    }
    subview &operator+=(const subview &x) {
        if (x.n_rows != n_rows || x.n_cols != n_cols) arma_stop_logic("addition: incompatible dimensions");
        for (uword j = 0; j < n_cols; ++j)
            for (uword i = 0; i < n_rows; ++i) at(i, j) += x.at(i, j);
        return *this;
    }
    subview &operator+=(const Mat<eT> &x) {
        if (x.n_rows != n_rows || x.n_cols != n_cols) arma_stop_logic("addition: incompatible dimensions");
        for (uword j = 0; j < n_cols; ++j)
            for (uword i = 0; i < n_rows; ++i) at(i, j) += x.mem[i + j * n_rows];
        return *this;
    }
    void fill(eT v) {
        for (uword j = 0; j < n_cols; ++j)
            for (uword i = 0; i < n_rows; ++i) at(i, j) = v;
    }
};

template <typename eT>
class Mat {
   public:
    uword n_rows = 0, n_cols = 0, n_elem = 0;
    eT *mem = nullptr;
    bool owner = true;  // false: a slice of a Cube (fixed size, memory owned by the cube)

    Mat() {}
    Mat(uword r, uword c) { init(r, c); }
    Mat(uword r, uword c, const fill::fill_zeros &) { init(r, c); }
    Mat(eT *aux, uword r, uword c, bool /*alias*/) : n_rows(r), n_cols(c), n_elem(r * c), mem(aux), owner(false) {}
    Mat(const Mat &x) {
        init(x.n_rows, x.n_cols);
        if (n_elem) std::memcpy(mem, x.mem, sizeof(eT) * n_elem);
    }
    Mat(Mat &&x) noexcept {
        if (x.owner) {
            n_rows = x.n_rows; n_cols = x.n_cols; n_elem = x.n_elem; mem = x.mem;
            x.mem = nullptr; x.n_rows = x.n_cols = x.n_elem = 0;
        } else {
            init(x.n_rows, x.n_cols);
            if (n_elem) std::memcpy(mem, x.mem, sizeof(eT) * n_elem);
        }
    }
    Mat(const subview<eT> &v) {
        init(v.n_rows, v.n_cols);
        for (uword j = 0; j < n_cols; ++j)
            for (uword i = 0; i < n_rows; ++i) mem[i + j * n_rows] = v.at(i, j);
    }
    ~Mat() {
        if (owner && mem) std::free(mem);
    }
    void init(uword r, uword c) {
        n_rows = r; n_cols = c; n_elem = r * c;
        mem = n_elem ? (eT *)std::calloc(n_elem, sizeof(eT)) : nullptr;
        if (n_elem && !mem) throw std::bad_alloc();
        owner = true;
    }
    void set_size(uword r, uword c) {
        if (r == n_rows && c == n_cols) return;
        if (!owner) arma_stop_logic("size is locked (cube slice)");
        if (mem) std::free(mem);
        init(r, c);
    }
    Mat &operator=(const Mat &x) {
        if (this == &x) return *this;
        set_size(x.n_rows, x.n_cols);
        if (n_elem) std::memmove(mem, x.mem, sizeof(eT) * n_elem);
        return *this;
    }
    Mat &operator=(Mat &&x) noexcept(false) {
        if (this == &x) return *this;
        if (owner && x.owner) {
            if (mem) std::free(mem);
            n_rows = x.n_rows; n_cols = x.n_cols; n_elem = x.n_elem; mem = x.mem;
            x.mem = nullptr; x.n_rows = x.n_cols = x.n_elem = 0;
        } else {
            set_size(x.n_rows, x.n_cols);
            if (n_elem) std::memmove(mem, x.mem, sizeof(eT) * n_elem);
        }
        return *this;
    }
    Mat &operator=(const subview<eT> &v) {
        Mat tmp(v);
        return (*this = std::move(tmp));
    }

    inline eT &operator()(uword i) {
        if (i >= n_elem) arma_stop_bounds("Mat::operator(): index out of bounds");
        return mem[i];
    }
    inline const eT &operator()(uword i) const {
        if (i >= n_elem) arma_stop_bounds("Mat::operator(): index out of bounds");
        return mem[i];
    }
    inline eT &operator[](uword i) { return mem[i]; }
    inline const eT &operator[](uword i) const { return mem[i]; }
    inline eT &operator()(uword i, uword j) {
        if (i >= n_rows || j >= n_cols) arma_stop_bounds("Mat::operator(): index out of bounds");
        return mem[i + j * n_rows];
    }
    inline const eT &operator()(uword i, uword j) const {
        if (i >= n_rows || j >= n_cols) arma_stop_bounds("Mat::operator(): index out of bounds");
        return mem[i + j * n_rows];
    }
    inline eT &at(uword i, uword j) { return mem[i + j * n_rows]; }
    inline const eT &at(uword i, uword j) const { return mem[i + j * n_rows]; }

    subview<eT> row(uword i) {
        if (i >= n_rows) arma_stop_bounds("Mat::row(): index out of bounds");
        return subview<eT>(*this, i, 0, 1, n_cols);
    }
    subview<eT> col(uword j) {
        if (j >= n_cols) arma_stop_bounds("Mat::col(): index out of bounds");
        return subview<eT>(*this, 0, j, n_rows, 1);
    }
    const subview<eT> row(uword i) const { return const_cast<Mat *>(this)->row(i); }
    const subview<eT> col(uword j) const { return const_cast<Mat *>(this)->col(j); }
    subview<eT> tail_rows(uword n) {
        if (n > n_rows) arma_stop_bounds("Mat::tail_rows(): size out of bounds");
        return subview<eT>(*this, n_rows - n, 0, n, n_cols);
    }
    subview<eT> tail(uword n) {  // vectors only
        if (n_cols == 1) return tail_rows(n);
        if (n_rows != 1 || n > n_cols) arma_stop_bounds("Mat::tail(): size out of bounds");
        return subview<eT>(*this, 0, n_cols - n, 1, n);
    }
    subview_each_col<eT> each_col() { return subview_each_col<eT>(*this); }

    void fill(eT v) {
        for (uword i = 0; i < n_elem; ++i) mem[i] = v;
    }
    void replace(eT old_val, eT new_val) {
        for (uword i = 0; i < n_elem; ++i)
            if (mem[i] == old_val) mem[i] = new_val;
    }
    Mat &operator/=(eT v) {
        for (uword i = 0; i < n_elem; ++i) mem[i] /= v;
        return *this;
    }
    Mat operator-() const {
        Mat out(n_rows, n_cols);
        for (uword i = 0; i < n_elem; ++i) out.mem[i] = (eT)(-mem[i]);
        return out;
    }
    Mat t() const {
        Mat out(n_cols, n_rows);
        for (uword j = 0; j < n_cols; ++j)
            for (uword i = 0; i < n_rows; ++i) out.mem[j + i * n_cols] = mem[i + j * n_rows];
        return out;
    }
};

template <typename eT>
class Col : public Mat<eT> {
   public:
    Col() : Mat<eT>(0, 1) {}
    explicit Col(uword n) : Mat<eT>(n, 1) {}
    Col(uword n, const fill::fill_zeros &) : Mat<eT>(n, 1) {}
    Col(const eT *src, uword n) : Mat<eT>(n, 1) {
        for (uword i = 0; i < n; ++i) this->mem[i] = src[i];
    }
    Col(const Mat<eT> &x) : Mat<eT>(x) {
        if (x.n_cols != 1 && x.n_elem) arma_stop_logic("Col: incompatible dimensions");
    }
    Col(Mat<eT> &&x) : Mat<eT>(std::move(x)) {}
    Col(const subview<eT> &v) : Mat<eT>(v) {}
    Col &operator=(const Mat<eT> &x) {
        Mat<eT>::operator=(x);
        return *this;
    }
    Row<eT> t() const;
};

template <typename eT>
class Row : public Mat<eT> {
   public:
    Row() : Mat<eT>(1, 0) {}
    explicit Row(uword n) : Mat<eT>(1, n) {}
    Row(const Mat<eT> &x) : Mat<eT>(x) {
        if (x.n_rows != 1 && x.n_elem) arma_stop_logic("Row: incompatible dimensions");
    }
    Row(Mat<eT> &&x) : Mat<eT>(std::move(x)) {}
    Row(const subview<eT> &v) : Mat<eT>(v) {}
    Row &operator=(const Mat<eT> &x) {
        Mat<eT>::operator=(x);
        return *this;
    }
    Col<eT> t() const {
        Col<eT> out(this->n_elem);
        for (uword i = 0; i < this->n_elem; ++i) out.mem[i] = this->mem[i];
        return out;
    }
};

template <typename eT>
Row<eT> Col<eT>::t() const {
    Row<eT> out(this->n_elem);
    for (uword i = 0; i < this->n_elem; ++i) out.mem[i] = this->mem[i];
    return out;
}

template <typename eT>
class Cube {
   public:
    uword n_rows = 0, n_cols = 0, n_slices = 0, n_elem = 0, n_elem_slice = 0;
    eT *mem = nullptr;
    std::vector<Mat<eT>> mats;  // aliases of the slices, like Armadillo's mat_ptrs

    Cube() {}
    Cube(uword r, uword c, uword s) { init(r, c, s); }
    Cube(uword r, uword c, uword s, const fill::fill_zeros &) { init(r, c, s); }
    Cube(const Cube &x) {
        init(x.n_rows, x.n_cols, x.n_slices);
        if (n_elem) std::memcpy(mem, x.mem, sizeof(eT) * n_elem);
    }
    Cube(Cube &&x) noexcept { steal(x); }
    ~Cube() {
        if (mem) std::free(mem);
    }
    Cube &operator=(const Cube &x) {
        if (this == &x) return *this;
        if (mem) std::free(mem);
        init(x.n_rows, x.n_cols, x.n_slices);
        if (n_elem) std::memcpy(mem, x.mem, sizeof(eT) * n_elem);
        return *this;
    }
    Cube &operator=(Cube &&x) noexcept {
        if (this == &x) return *this;
        if (mem) std::free(mem);
        steal(x);
        return *this;
    }
    void steal(Cube &x) {
        n_rows = x.n_rows; n_cols = x.n_cols; n_slices = x.n_slices; n_elem = x.n_elem; n_elem_slice = x.n_elem_slice;
        mem = x.mem;
        x.mem = nullptr; x.n_rows = x.n_cols = x.n_slices = x.n_elem = x.n_elem_slice = 0; x.mats.clear();
        bind();
    }
    void init(uword r, uword c, uword s) {
        n_rows = r; n_cols = c; n_slices = s; n_elem_slice = r * c; n_elem = r * c * s;
        mem = n_elem ? (eT *)std::calloc(n_elem, sizeof(eT)) : nullptr;
        if (n_elem && !mem) throw std::bad_alloc();
        bind();
    }
    void bind() {
        mats.clear();
        mats.reserve(n_slices);
        for (uword s = 0; s < n_slices; ++s) mats.emplace_back(mem + s * n_elem_slice, n_rows, n_cols, true);
    }
    inline eT &operator()(uword i, uword j, uword s) {
        if (i >= n_rows || j >= n_cols || s >= n_slices) arma_stop_bounds("Cube::operator(): index out of bounds");
        return mem[i + j * n_rows + s * n_elem_slice];
    }
    inline const eT &operator()(uword i, uword j, uword s) const {
        if (i >= n_rows || j >= n_cols || s >= n_slices) arma_stop_bounds("Cube::operator(): index out of bounds");
        return mem[i + j * n_rows + s * n_elem_slice];
    }
    Mat<eT> &slice(uword s) {
        if (s >= n_slices) arma_stop_bounds("Cube::slice(): index out of bounds");
        return mats[s];
    }
    const Mat<eT> &slice(uword s) const {
        if (s >= n_slices) arma_stop_bounds("Cube::slice(): index out of bounds");
        return mats[s];
    }
    Cube tail_slices(uword n) const {  // Armadillo returns a subview_cube; the reference copies it into a cube
        if (n > n_slices) arma_stop_bounds("Cube::tail_slices(): size out of bounds");
        Cube out(n_rows, n_cols, n);
        if (out.n_elem) std::memcpy(out.mem, mem + (n_slices - n) * n_elem_slice, sizeof(eT) * out.n_elem);
        return out;
    }
    void replace(eT old_val, eT new_val) {
        for (uword i = 0; i < n_elem; ++i)
            if (mem[i] == old_val) mem[i] = new_val;
    }
    void fill(eT v) {
        for (uword i = 0; i < n_elem; ++i) mem[i] = v;
    }
};

typedef Mat<double> mat;
typedef Col<double> vec;
typedef Col<double> colvec;
typedef Row<double> rowvec;
typedef Cube<double> cube;
typedef Col<uword> uvec;
typedef Mat<uword> umat;

inline vec zeros(uword n) { return vec(n); }

template <typename eT>
class subview_each_col {
   public:
    const Mat<eT> &m;
    explicit subview_each_col(const Mat<eT> &m_) : m(m_) {}
};
template <typename eT>
Mat<eT> operator-(const subview_each_col<eT> &e, const Mat<eT> &c) {
    if (c.n_cols != 1 || c.n_rows != e.m.n_rows) arma_stop_logic("each_col(): incompatible size");
    Mat<eT> out(e.m.n_rows, e.m.n_cols);
    for (uword j = 0; j < out.n_cols; ++j)
        for (uword i = 0; i < out.n_rows; ++i) out.mem[i + j * out.n_rows] = e.m.mem[i + j * out.n_rows] - c.mem[i];
    return out;
}

inline Mat<double> log(const Mat<double> &x) {
    Mat<double> out(x.n_rows, x.n_cols);
    for (uword i = 0; i < x.n_elem; ++i) out.mem[i] = std::log(x.mem[i]);
    return out;
}
inline Mat<double> log(const subview<double> &v) { return log(Mat<double>(v)); }

// element-wise (Schur) product; mixed element types promote to the first operand's, as the path needs
template <typename T1, typename T2>
Mat<T1> operator%(const Mat<T1> &a, const Mat<T2> &b) {
    if (a.n_rows != b.n_rows || a.n_cols != b.n_cols) arma_stop_logic("element-wise multiplication: incompatible dimensions");
    Mat<T1> out(a.n_rows, a.n_cols);
    for (uword i = 0; i < a.n_elem; ++i) out.mem[i] = a.mem[i] * (T1)b.mem[i];
    return out;
}
template <typename eT>
Mat<eT> operator+(const Mat<eT> &a, const Mat<eT> &b) {
    if (a.n_rows != b.n_rows || a.n_cols != b.n_cols) arma_stop_logic("addition: incompatible dimensions");
    Mat<eT> out(a.n_rows, a.n_cols);
    for (uword i = 0; i < a.n_elem; ++i) out.mem[i] = a.mem[i] + b.mem[i];
    return out;
}
template <typename eT>
Mat<eT> operator-(const Mat<eT> &a, const Mat<eT> &b) {
    if (a.n_rows != b.n_rows || a.n_cols != b.n_cols) arma_stop_logic("subtraction: incompatible dimensions");
    Mat<eT> out(a.n_rows, a.n_cols);
    for (uword i = 0; i < a.n_elem; ++i) out.mem[i] = a.mem[i] - b.mem[i];
    return out;
}
inline Mat<double> operator*(double k, const Mat<double> &a) {
    Mat<double> out(a.n_rows, a.n_cols);
    for (uword i = 0; i < a.n_elem; ++i) out.mem[i] = a.mem[i] * k;
    return out;
}
inline Mat<double> operator*(const Mat<double> &a, double k) { return k * a; }
inline Mat<double> operator/(const Mat<double> &a, double k) {
    Mat<double> out(a.n_rows, a.n_cols);
    for (uword i = 0; i < a.n_elem; ++i) out.mem[i] = a.mem[i] / k;
    return out;
}

// Armadillo's accumulate: two running sums over even / odd positions
template <typename eT>
inline eT accumulate2(const eT *src, uword n) {
    eT acc1 = eT(0), acc2 = eT(0);
    uword j;
    for (j = 1; j < n; j += 2) {
        acc1 += *src++;
        acc2 += *src++;
    }
    if ((j - 1) < n) acc1 += *src;
    return acc1 + acc2;
}
template <typename eT>
Mat<eT> sum(const Mat<eT> &x, uword dim = 0) {
    if (dim == 0) {
        Mat<eT> out(1, x.n_cols);
        for (uword c = 0; c < x.n_cols; ++c) out.mem[c] = accumulate2(x.mem + c * x.n_rows, x.n_rows);
        return out;
    }
    Mat<eT> out(x.n_rows, 1);
    for (uword c = 0; c < x.n_cols; ++c)
        for (uword r = 0; r < x.n_rows; ++r) out.mem[r] += x.mem[r + c * x.n_rows];
    return out;
}
template <typename eT>
eT accu(const Mat<eT> &x) {
    return accumulate2(x.mem, x.n_elem);
}

template <typename eT>
uword index_max(const subview<eT> &v) {
    if (v.n_elem == 0) arma_stop_logic("index_max(): object has no elements");
    uword best = 0;
    eT bv = v(0);
    for (uword i = 1; i < v.n_elem; ++i)
        if (v(i) > bv) { bv = v(i); best = i; }
    return best;
}
template <typename eT>
uword index_max(const Mat<eT> &v) {
    if (v.n_elem == 0) arma_stop_logic("index_max(): object has no elements");
    uword best = 0;
    for (uword i = 1; i < v.n_elem; ++i)
        if (v.mem[i] > v.mem[best]) best = i;
    return best;
}

// sort_index: std::sort on (value, index) packets like Armadillo (not a stable sort)
template <typename eT>
struct sort_packet {
    eT val;
    uword index;
};
template <typename eT>
uvec sort_index(const Mat<eT> &x, const char *dir = "ascend") {
    std::vector<sort_packet<eT>> pk(x.n_elem);
    for (uword i = 0; i < x.n_elem; ++i) { pk[i].val = x.mem[i]; pk[i].index = i; }
    if (dir[0] == 'd')
        std::sort(pk.begin(), pk.end(), [](const sort_packet<eT> &a, const sort_packet<eT> &b) { return a.val > b.val; });
    else
        std::sort(pk.begin(), pk.end(), [](const sort_packet<eT> &a, const sort_packet<eT> &b) { return a.val < b.val; });
    uvec out(x.n_elem);
    for (uword i = 0; i < x.n_elem; ++i) out.mem[i] = pk[i].index;
    return out;
}
template <typename eT>
uvec sort_index(const subview<eT> &v, const char *dir = "ascend") {
    return sort_index(Mat<eT>(v), dir);
}
template <typename eT>
Mat<eT> sort(const Mat<eT> &x, const char *dir = "ascend") {
    Mat<eT> out(x);
    if (dir[0] == 'd') std::sort(out.mem, out.mem + out.n_elem, std::greater<eT>());
    else std::sort(out.mem, out.mem + out.n_elem);
    return out;
}

// relational operators between objects of different element types (what `a <- f(b)` really parses to:
// `a < -f(b)`, stephens.cpp:56,85); the result is discarded by the reference
template <typename T1, typename T2>
umat operator<(const Mat<T1> &a, const Mat<T2> &b) {
    if (a.n_elem != b.n_elem) arma_stop_logic("relational operator: incompatible dimensions");
    umat out(a.n_rows, a.n_cols);
    for (uword i = 0; i < a.n_elem; ++i) out.mem[i] = ((double)a.mem[i] < (double)b.mem[i]) ? 1 : 0;
    return out;
}
template <typename T1, typename T2>
umat operator<(const subview<T1> &a, const Mat<T2> &b) {
    return Mat<T1>(a) < b;
}

}  // namespace arma

// =============================================================================================
// Rcpp surface
// =============================================================================================
namespace Rcpp {

class exception : public std::runtime_error {
   public:
    explicit exception(const std::string &m) : std::runtime_error(m) {}
};
inline void stop(const std::string &m) { throw Rcpp::exception(m); }

// console output is sunk (the CPU baseline is timed with progress printing off)
struct NullStream {
    template <typename T>
    NullStream &operator<<(const T &) { return *this; }
};
inline NullStream Rcout;
inline NullStream Rcerr;

struct RNGScope {  // GetRNGstate()/PutRNGstate(): the thread's generator is always live here
    RNGScope() {}
};

struct Underscore {};
static const Underscore _ = Underscore();
struct Range {
    int lo, hi;
    Range(int a, int b) : lo(a), hi(b) {}
};

inline std::shared_ptr<SEXPREC> new_sexp(int type) {
    auto p = std::make_shared<SEXPREC>();
    p->type = type;
    return p;
}

// a protected R object
class RObject {
   public:
    std::shared_ptr<SEXPREC> p;
    RObject() {}
    RObject(std::shared_ptr<SEXPREC> q) : p(std::move(q)) {}
    RObject(SEXP raw) {  // borrow: the caller owns it for the duration of the call
        p = std::shared_ptr<SEXPREC>(raw, [](SEXPREC *) {});
    }
    operator SEXP() const {
        bmm_shim::arena().push_back(p);
        return p.get();
    }
};

template <typename T> struct sexp_traits;
template <> struct sexp_traits<double> {
    static const int type = REALSXP;
    static std::vector<double> &vec(SEXPREC &s) { return s.real; }
};
template <> struct sexp_traits<int> {
    static const int type = INTSXP;
    static std::vector<int> &vec(SEXPREC &s) { return s.integer; }
};

template <typename T>
inline std::vector<T> sexp_values(const SEXPREC *s) {  // with R's int <-> double coercion
    std::vector<T> out;
    if (!s) return out;
    if (s->type == REALSXP) out.assign(s->real.begin(), s->real.end());
    else out.assign(s->integer.begin(), s->integer.end());
    return out;
}

// Rcpp::Dimension(K, P, S): the dim attribute of an array-valued Vector (used by r-package/src/host.cpp)
class Dimension {
   public:
    std::vector<int> v;
    Dimension(int a) : v{a} {}
    Dimension(int a, int b) : v{a, b} {}
    Dimension(int a, int b, int c) : v{a, b, c} {}
    size_t prod() const {
        size_t n = 1;
        for (int x : v) n *= (size_t)x;
        return n;
    }
};

template <typename T>
class Vector {
   public:
    std::vector<T> d;
    std::vector<int> dimv;  // dim attribute, empty for a plain vector
    Vector() {}
    Vector(int n) : d((size_t)n, T(0)) {}
    Vector(const Dimension &dm) : d(dm.prod(), T(0)), dimv(dm.v) {}
    Vector(const RObject &o) : d(sexp_values<T>(o.p.get())), dimv(o.p.get() ? o.p->dim : std::vector<int>()) {}
    Vector(SEXP s) : d(sexp_values<T>(s)), dimv(s ? s->dim : std::vector<int>()) {}
    Vector &operator=(const RObject &o) {
        d = sexp_values<T>(o.p.get());
        dimv = o.p.get() ? o.p->dim : std::vector<int>();
        return *this;
    }
    // x.attr("dim") read as an IntegerVector (the only attribute the host reads)
    class AttrProxy {
       public:
        const Vector &v;
        std::string name;
        AttrProxy(const Vector &v_, const std::string &n) : v(v_), name(n) {}
        operator Vector<int>() const {
            if (name != "dim") throw Rcpp::exception("Rcpp shim: only the dim attribute is modelled");
            Vector<int> out((int)v.dimv.size());
            for (size_t i = 0; i < v.dimv.size(); ++i) out.d[i] = v.dimv[i];
            return out;
        }
    };
    AttrProxy attr(const std::string &name) const { return AttrProxy(*this, name); }
    inline T &operator[](int i) { return d[(size_t)i]; }
    inline const T &operator[](int i) const { return d[(size_t)i]; }
    inline T &operator()(int i) {
        if (i < 0 || (size_t)i >= d.size()) throw std::out_of_range("Rcpp shim: Vector index out of bounds");
        return d[(size_t)i];
    }
    inline const T &operator()(int i) const {
        if (i < 0 || (size_t)i >= d.size()) throw std::out_of_range("Rcpp shim: Vector index out of bounds");
        return d[(size_t)i];
    }
    typedef T *iterator;  // Rcpp iterators are raw pointers into the SEXP payload
    typedef const T *const_iterator;
    iterator begin() { return d.data(); }
    iterator end() { return d.data() + d.size(); }
    const_iterator begin() const { return d.data(); }
    const_iterator end() const { return d.data() + d.size(); }
    int size() const { return (int)d.size(); }
    int length() const { return (int)d.size(); }
};
typedef Vector<double> NumericVector;
typedef Vector<int> IntegerVector;

inline double max(const NumericVector &x) {
    double m = x.d.empty() ? -INFINITY : x.d[0];
    for (double v : x.d)
        if (v > m) m = v;
    return m;
}
inline int max(const IntegerVector &x) {
    int m = x.d.empty() ? INT_MIN : x.d[0];
    for (int v : x.d)
        if (v > m) m = v;
    return m;
}

template <typename T> class Matrix;

template <typename T>
class MatrixRow {
   public:
    Matrix<T> &m;
    int r;
    MatrixRow(Matrix<T> &m_, int r_) : m(m_), r(r_) {}
    MatrixRow &operator=(const Vector<T> &v) {
        if (v.size() != m.ncol()) throw std::out_of_range("Rcpp shim: row assignment of the wrong length");
        for (int c = 0; c < m.ncol(); ++c) m.d[(size_t)r + (size_t)m.nr * c] = v.d[(size_t)c];
        return *this;
    }
};

template <typename T>
class SubMatrix {
   public:
    const Matrix<T> &m;
    int r0, nrows;
    SubMatrix(const Matrix<T> &m_, int r0_, int nrows_) : m(m_), r0(r0_), nrows(nrows_) {}
};

template <typename T>
class Matrix {
   public:
    std::vector<T> d;
    int nr = 0, nc = 0;
    Matrix() {}
    Matrix(int r, int c) : d((size_t)r * c, T(0)), nr(r), nc(c) {}
    Matrix(const RObject &o) { from(o.p.get()); }
    Matrix(SEXP s) { from(s); }
    Matrix &operator=(const RObject &o) {
        from(o.p.get());
        return *this;
    }
    void from(const SEXPREC *s) {
        if (!s || s->dim.size() != 2) throw Rcpp::exception("Rcpp shim: not a matrix");
        nr = s->dim[0];
        nc = s->dim[1];
        d = sexp_values<T>(s);
    }
    int nrow() const { return nr; }
    int ncol() const { return nc; }
    typedef T *iterator;  // column-major payload, as in Rcpp
    iterator begin() { return d.data(); }
    iterator end() { return d.data() + d.size(); }
    const T *begin() const { return d.data(); }
    const T *end() const { return d.data() + d.size(); }
    inline T &operator()(int i, int j) { return d[(size_t)i + (size_t)nr * j]; }
    inline const T &operator()(int i, int j) const { return d[(size_t)i + (size_t)nr * j]; }
    MatrixRow<T> operator()(int i, const Underscore &) { return MatrixRow<T>(*this, i); }
    SubMatrix<T> operator()(const Range &r, const Underscore &) const {
        if (r.lo < 0 || r.hi >= nr || r.hi < r.lo - 1) throw std::out_of_range("Rcpp shim: Range out of bounds");
        return SubMatrix<T>(*this, r.lo, r.hi - r.lo + 1);
    }
};
typedef Matrix<double> NumericMatrix;
typedef Matrix<int> IntegerMatrix;

// ---- wrap ---------------------------------------------------------------------------------------
template <typename T>
RObject wrap_array(const T *src, size_t n, std::vector<int> dim) {
    auto p = new_sexp(sexp_traits<T>::type);
    sexp_traits<T>::vec(*p).assign(src, src + n);
    p->dim = std::move(dim);
    return RObject(p);
}
inline RObject wrap(const RObject &o) { return o; }
template <typename T>
RObject wrap(const Vector<T> &v) { return wrap_array(v.d.data(), v.d.size(), v.dimv); }
template <typename T>
RObject wrap(const Matrix<T> &m) { return wrap_array(m.d.data(), m.d.size(), {m.nr, m.nc}); }
template <typename T>
RObject wrap(const SubMatrix<T> &s) {
    std::vector<T> tmp((size_t)s.nrows * s.m.nc);
    for (int c = 0; c < s.m.nc; ++c)
        for (int r = 0; r < s.nrows; ++r) tmp[(size_t)r + (size_t)s.nrows * c] = s.m.d[(size_t)(s.r0 + r) + (size_t)s.m.nr * c];
    return wrap_array(tmp.data(), tmp.size(), {s.nrows, s.m.nc});
}
// RcppArmadillo: every Mat / Col / Row carries a dim attribute (no RCPP_ARMADILLO_RETURN_*_AS_VECTOR)
template <typename T>
RObject wrap(const arma::Mat<T> &m) { return wrap_array(m.mem, m.n_elem, {(int)m.n_rows, (int)m.n_cols}); }
template <typename T>
RObject wrap(const arma::subview<T> &v) { return wrap(arma::Mat<T>(v)); }
template <typename T>
RObject wrap(const arma::Cube<T> &c) {
    return wrap_array(c.mem, c.n_elem, {(int)c.n_rows, (int)c.n_cols, (int)c.n_slices});
}
inline RObject wrap(double v) { return wrap_array(&v, 1, {}); }
inline RObject wrap(int v) { return wrap_array(&v, 1, {}); }

class List {
   public:
    std::shared_ptr<SEXPREC> p;
    List() : p(new_sexp(VECSXP)) {}
    class Proxy {
       public:
        List &l;
        std::string name;
        Proxy(List &l_, const std::string &n) : l(l_), name(n) {}
        template <typename T>
        Proxy &operator=(const T &v) {
            l.set(name, wrap(v));
            return *this;
        }
    };
    Proxy operator[](const std::string &name) { return Proxy(*this, name); }
    // List::create(Named("a") = x, Named("b") = y, ...)
    struct NamedValue {
        std::string name;
        RObject value;
    };
    static void fill(List &) {}
    template <typename... Rest>
    static void fill(List &l, const NamedValue &nv, const Rest &...rest) {
        l.set(nv.name, nv.value);
        fill(l, rest...);
    }
    template <typename... Args>
    static List create(const Args &...args) {
        List l;
        fill(l, args...);
        return l;
    }
    void set(const std::string &name, const RObject &v) {
        for (size_t i = 0; i < p->names.size(); ++i)
            if (p->names[i] == name) {
                p->elts[i] = v.p;
                return;
            }
        p->names.push_back(name);
        p->elts.push_back(v.p);
    }
};
inline RObject wrap(const List &l) { return RObject(l.p); }
class Named {
   public:
    std::string name;
    explicit Named(const std::string &n) : name(n) {}
    template <typename T>
    List::NamedValue operator=(const T &v) const { return List::NamedValue{name, wrap(v)}; }
};

// ---- as -----------------------------------------------------------------------------------------
template <typename T> struct as_tag {};

template <typename T>
inline arma::Mat<T> as_impl(as_tag<arma::Mat<T>>, const Matrix<T> &m) {
    arma::Mat<T> out((arma::uword)m.nr, (arma::uword)m.nc);
    if (out.n_elem) std::memcpy(out.mem, m.d.data(), sizeof(T) * out.n_elem);
    return out;
}
template <typename T, typename U>
inline arma::Col<T> as_impl(as_tag<arma::Col<T>>, const Vector<U> &v) {
    arma::Col<T> out((arma::uword)v.d.size());
    for (size_t i = 0; i < v.d.size(); ++i) out.mem[i] = (T)v.d[i];
    return out;
}
template <typename T, typename U>
inline arma::Row<T> as_impl(as_tag<arma::Row<T>>, const Vector<U> &v) {
    arma::Row<T> out((arma::uword)v.d.size());
    for (size_t i = 0; i < v.d.size(); ++i) out.mem[i] = (T)v.d[i];
    return out;
}
// from a SEXP
inline int as_impl(as_tag<int>, SEXP s) {
    if (!s || s->length() != 1) throw Rcpp::exception("Expecting a single value");
    return s->type == REALSXP ? (int)s->real[0] : s->integer[0];
}
inline double as_impl(as_tag<double>, SEXP s) {
    if (!s || s->length() != 1) throw Rcpp::exception("Expecting a single value");
    return s->type == REALSXP ? s->real[0] : (double)s->integer[0];
}
inline bool as_impl(as_tag<bool>, SEXP s) {
    if (!s || s->length() != 1) throw Rcpp::exception("Expecting a single value");
    return s->type == REALSXP ? s->real[0] != 0 : s->integer[0] != 0;
}
template <typename T>
inline Vector<T> as_impl(as_tag<Vector<T>>, SEXP s) { return Vector<T>(s); }
template <typename T>
inline Matrix<T> as_impl(as_tag<Matrix<T>>, SEXP s) { return Matrix<T>(s); }
template <typename T>
inline arma::Col<T> as_impl(as_tag<arma::Col<T>>, SEXP s) {
    std::vector<T> v = sexp_values<T>(s);
    return arma::Col<T>(v.data(), (arma::uword)v.size());
}
template <typename T>
inline arma::Mat<T> as_impl(as_tag<arma::Mat<T>>, SEXP s) {
    if (!s || s->dim.size() != 2) throw Rcpp::exception("Rcpp shim: not a matrix");
    std::vector<T> v = sexp_values<T>(s);
    arma::Mat<T> out((arma::uword)s->dim[0], (arma::uword)s->dim[1]);
    if (out.n_elem) std::memcpy(out.mem, v.data(), sizeof(T) * out.n_elem);
    return out;
}
template <typename T>
inline arma::Cube<T> as_impl(as_tag<arma::Cube<T>>, SEXP s) {
    if (!s || s->dim.size() != 3) throw Rcpp::exception("Rcpp shim: not a 3-d array");
    std::vector<T> v = sexp_values<T>(s);
    arma::Cube<T> out((arma::uword)s->dim[0], (arma::uword)s->dim[1], (arma::uword)s->dim[2]);
    if (out.n_elem) std::memcpy(out.mem, v.data(), sizeof(T) * out.n_elem);
    return out;
}

template <typename T, typename U>
inline T as(const U &x) { return as_impl(as_tag<T>(), x); }

template <typename T>
class InputParameter {
   public:
    SEXP x;
    InputParameter(SEXP x_) : x(x_) {}
    operator T() { return as<T>(x); }
};
namespace traits {
template <typename T>
struct input_parameter {
    typedef InputParameter<T> type;
};
}  // namespace traits

// the error a .Call raised, kept for the caller (R would turn it into a condition)
inline std::string &last_condition() {
    static thread_local std::string s;
    return s;
}

}  // namespace Rcpp

#define BEGIN_RCPP \
    try {
#define END_RCPP                                                   \
    }                                                              \
    catch (std::exception & ex__) {                                \
        Rcpp::last_condition() = ex__.what();                      \
        return (SEXP) nullptr;                                     \
    }                                                              \
    catch (...) {                                                  \
        Rcpp::last_condition() = "c++ exception (unknown reason)"; \
        return (SEXP) nullptr;                                     \
    }
