// MiniEigen -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A small, eager (no expression templates) stand-in for the part of the Eigen 3 API that the reference
// (2iw31Zhv/AnisotropicElastoplasticity) uses, written from scratch so that the reference's OWN, UNMODIFIED sources under
// /root/reference can be compiled in an image that has no Eigen (oracle/Makefile, target _ref/libaep_ref.so).  It is only
// ever included through -Ioracle/ref_shim when building that library; nothing in the product includes it.
//
// What it has to get right for the reference's arithmetic to be the reference's arithmetic:
//   * column-major dense storage, (r,c) / [i] indexing, blocks that alias their parent (row(), col(), block<>, segment<>);
//   * assignment between a row and a column vector of equal length copies element by element (Eigen allows this);
//   * SparseMatrix is compressed column storage filled by setFromTriplets (duplicates summed), products with dense
//     operands in either orientation, InnerIterator in column order;
//   * JacobiSVD returns U, V orthogonal and singular values sorted in decreasing order, all >= 0 (two-sided Jacobi as in
//     Eigen; the reference only uses combinations that do not depend on the remaining sign freedom);
//   * Matrix::Random() draws from std::rand() in [-1, 1].
// One deliberate accommodation: RegularGrid::minBound()/maxBound() return `const Vector3d&` from a VectorXd member
// (RegularGrid.h:54-55), which with real Eigen binds a reference to a dead temporary.  Here a dynamic column vector converts
// to `const Matrix<T,N,1>&` through a cache inside the vector itself, so the reference reads the values it meant to read.
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <type_traits>
#include <vector>

namespace Eigen {

const int Dynamic = -1;
enum { ComputeThinU = 0x08, ComputeFullU = 0x04, ComputeThinV = 0x20, ComputeFullV = 0x10 };
enum { Lower = 0x1, Upper = 0x2, UnitDiag = 0x4, ZeroDiag = 0x8, StrictlyLower = ZeroDiag | Lower, StrictlyUpper = ZeroDiag | Upper };

template <class T, int R, int C> class Matrix;
template <class T, int R, int C> class Block;
template <class T> class DiagWrap;
template <class T, int R, int C> struct Arr;
template <class T, int R, int C> struct TriView;

template <class D> struct traits;
template <class T, int R, int C> struct traits<Matrix<T, R, C>> { typedef T Scalar; enum { Rows = R, Cols = C }; };
template <class T, int R, int C> struct traits<Block<T, R, C>> { typedef T Scalar; enum { Rows = R, Cols = C }; };

template <int A, int B> struct pick { enum { value = (A != Dynamic) ? A : B }; };

// ------------------------------------------------------------------------------------------------ read-only interface
template <class D> class DenseBase {
public:
    typedef typename traits<D>::Scalar Scalar;
    enum { Rows = traits<D>::Rows, Cols = traits<D>::Cols };
    typedef Matrix<Scalar, Rows, Cols> Plain;

    const D& derived() const { return *static_cast<const D*>(this); }
    D& derived() { return *static_cast<D*>(this); }

    int size() const { return derived().rows() * derived().cols(); }
    Scalar lin(int i) const { return derived().cols() == 1 ? derived().coeff(i, 0) : derived().coeff(0, i); }
    Scalar operator()(int r, int c) const { return derived().coeff(r, c); }
    Scalar operator()(int i) const { return lin(i); }
    Scalar operator[](int i) const { return lin(i); }
    Scalar x() const { return lin(0); }
    Scalar y() const { return lin(1); }
    Scalar z() const { return lin(2); }

    Plain eval() const {
        Plain m; m.resize(derived().rows(), derived().cols());
        for (int c = 0; c < m.cols(); ++c) for (int r = 0; r < m.rows(); ++r) m.coeffRef(r, c) = derived().coeff(r, c);
        return m;
    }
    Matrix<Scalar, Cols, Rows> transpose() const {
        Matrix<Scalar, Cols, Rows> m; m.resize(derived().cols(), derived().rows());
        for (int c = 0; c < derived().cols(); ++c) for (int r = 0; r < derived().rows(); ++r) m.coeffRef(c, r) = derived().coeff(r, c);
        return m;
    }
    Scalar squaredNorm() const {
        Scalar s = 0;
        for (int c = 0; c < derived().cols(); ++c) for (int r = 0; r < derived().rows(); ++r) s += derived().coeff(r, c) * derived().coeff(r, c);
        return s;
    }
    Scalar norm() const { return std::sqrt(squaredNorm()); }
    Scalar sum() const {
        Scalar s = 0;
        for (int c = 0; c < derived().cols(); ++c) for (int r = 0; r < derived().rows(); ++r) s += derived().coeff(r, c);
        return s;
    }
    Scalar prod() const {
        Scalar s = 1;
        for (int c = 0; c < derived().cols(); ++c) for (int r = 0; r < derived().rows(); ++r) s *= derived().coeff(r, c);
        return s;
    }
    Scalar minCoeff() const {
        Scalar s = derived().coeff(0, 0);
        for (int c = 0; c < derived().cols(); ++c) for (int r = 0; r < derived().rows(); ++r) s = std::min(s, derived().coeff(r, c));
        return s;
    }
    Scalar maxCoeff() const {
        Scalar s = derived().coeff(0, 0);
        for (int c = 0; c < derived().cols(); ++c) for (int r = 0; r < derived().rows(); ++r) s = std::max(s, derived().coeff(r, c));
        return s;
    }
    Scalar trace() const { Scalar s = 0; for (int i = 0; i < derived().rows(); ++i) s += derived().coeff(i, i); return s; }
    Plain normalized() const { Plain m = eval(); Scalar n = norm(); if (n > Scalar(0)) m /= n; return m; }

    template <class O> Scalar dot(const DenseBase<O>& o) const {
        assert(size() == o.size()); Scalar s = 0;
        for (int i = 0; i < size(); ++i) s += lin(i) * o.lin(i);
        return s;
    }
    template <class O> Plain cross(const DenseBase<O>& o) const {
        assert(size() == 3 && o.size() == 3);
        Plain m; m.resize(derived().rows(), derived().cols());
        Scalar a0 = lin(0), a1 = lin(1), a2 = lin(2), b0 = o.lin(0), b1 = o.lin(1), b2 = o.lin(2);
        m.linRef(0) = a1 * b2 - a2 * b1; m.linRef(1) = a2 * b0 - a0 * b2; m.linRef(2) = a0 * b1 - a1 * b0;
        return m;
    }
    template <class O> Plain cwiseProduct(const DenseBase<O>& o) const {
        assert(size() == o.size());
        Plain m = eval();
        if (derived().rows() == o.derived().rows()) { for (int c = 0; c < m.cols(); ++c) for (int r = 0; r < m.rows(); ++r) m.coeffRef(r, c) *= o.derived().coeff(r, c); }
        else for (int i = 0; i < size(); ++i) m.linRef(i) *= o.lin(i);
        return m;
    }
    Plain cwiseInverse() const {
        Plain m = eval();
        for (int c = 0; c < m.cols(); ++c) for (int r = 0; r < m.rows(); ++r) m.coeffRef(r, c) = Scalar(1) / m.coeff(r, c);
        return m;
    }
    Plain cwiseAbs() const {
        Plain m = eval();
        for (int c = 0; c < m.cols(); ++c) for (int r = 0; r < m.rows(); ++r) m.coeffRef(r, c) = std::abs(m.coeff(r, c));
        return m;
    }
    Scalar determinant() const {
        const D& a = derived(); assert(a.rows() == a.cols());
        if (a.rows() == 1) return a.coeff(0, 0);
        if (a.rows() == 2) return a.coeff(0, 0) * a.coeff(1, 1) - a.coeff(0, 1) * a.coeff(1, 0);
        assert(a.rows() == 3);
        return a.coeff(0, 0) * (a.coeff(1, 1) * a.coeff(2, 2) - a.coeff(1, 2) * a.coeff(2, 1))
             - a.coeff(0, 1) * (a.coeff(1, 0) * a.coeff(2, 2) - a.coeff(1, 2) * a.coeff(2, 0))
             + a.coeff(0, 2) * (a.coeff(1, 0) * a.coeff(2, 1) - a.coeff(1, 1) * a.coeff(2, 0));
    }
    // inverse by cofactors (2x2, 3x3 -- the only sizes the reference inverts)
    Plain inverse() const {
        const D& a = derived(); assert(a.rows() == a.cols());
        Plain m; m.resize(a.rows(), a.cols());
        Scalar det = determinant();
        if (a.rows() == 1) { m.coeffRef(0, 0) = Scalar(1) / a.coeff(0, 0); return m; }
        if (a.rows() == 2) {
            Scalar id = Scalar(1) / det;
            m.coeffRef(0, 0) = a.coeff(1, 1) * id; m.coeffRef(0, 1) = -a.coeff(0, 1) * id;
            m.coeffRef(1, 0) = -a.coeff(1, 0) * id; m.coeffRef(1, 1) = a.coeff(0, 0) * id;
            return m;
        }
        assert(a.rows() == 3);
        Scalar id = Scalar(1) / det;
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) {
            int r1 = (c + 1) % 3, r2 = (c + 2) % 3, c1 = (r + 1) % 3, c2 = (r + 2) % 3;     // cofactor of (c,r), cyclic form
            m.coeffRef(r, c) = (a.coeff(r1, c1) * a.coeff(r2, c2) - a.coeff(r1, c2) * a.coeff(r2, c1)) * id;
        }
        return m;
    }
    DiagWrap<Scalar> asDiagonal() const { DiagWrap<Scalar> d; d.v.resize(size()); for (int i = 0; i < size(); ++i) d.v[i] = lin(i); return d; }
    Arr<Scalar, Rows, Cols> array() const { Arr<Scalar, Rows, Cols> a; a.m = eval(); return a; }
    Matrix<Scalar, Dynamic, 1> diagonal() const {
        int n = std::min(derived().rows(), derived().cols());
        Matrix<Scalar, Dynamic, 1> d; d.resize(n, 1);
        for (int i = 0; i < n; ++i) d.coeffRef(i, 0) = derived().coeff(i, i);
        return d;
    }
    template <unsigned Mode> TriView<Scalar, Rows, Cols> triangularView() const {
        TriView<Scalar, Rows, Cols> t; t.m = eval();
        for (int c = 0; c < t.m.cols(); ++c) for (int r = 0; r < t.m.rows(); ++r) {
            bool keep = (Mode & Upper) ? (c > r) : (r > c);
            if (r == c) keep = !(Mode & ZeroDiag);
            if (!keep) t.m.coeffRef(r, c) = 0;
            else if (r == c && (Mode & UnitDiag)) t.m.coeffRef(r, c) = 1;
        }
        return t;
    }
};

template <class T, int R, int C> struct TriView {
    Matrix<T, R, C> m;
    Matrix<T, R, C> toDenseMatrix() const { return m; }
    TriView<T, C, R> transpose() const { TriView<T, C, R> t; t.m = m.transpose(); return t; }
};

// ------------------------------------------------------------------------------------------------ writable interface
template <class M> struct CommaInit {
    M& m; int i;
    CommaInit(M& mm, typename M::Scalar first) : m(mm), i(0) { put(first); }
    void put(typename M::Scalar v) { int r = i / m.cols(), c = i % m.cols(); m.coeffRef(r, c) = v; ++i; }
    template <class S> CommaInit& operator,(S v) { put(static_cast<typename M::Scalar>(v)); return *this; }
};

template <class D> class DenseMut : public DenseBase<D> {
public:
    typedef DenseBase<D> Base;
    typedef typename Base::Scalar Scalar;
    enum { Rows = Base::Rows, Cols = Base::Cols };
    using Base::derived;
    using Base::operator();
    using Base::operator[];
    using Base::x; using Base::y; using Base::z;

    Scalar& linRef(int i) { return derived().cols() == 1 ? derived().coeffRef(i, 0) : derived().coeffRef(0, i); }
    Scalar& operator()(int r, int c) { return derived().coeffRef(r, c); }
    Scalar& operator()(int i) { return linRef(i); }
    Scalar& operator[](int i) { return linRef(i); }
    Scalar& x() { return linRef(0); }
    Scalar& y() { return linRef(1); }
    Scalar& z() { return linRef(2); }

    // element-wise copy; a row may be assigned to a column of the same length and vice versa
    template <class O> void assignFrom(const DenseBase<O>& o) {
        D& d = derived(); const O& s = o.derived();
        if (d.rows() == s.rows() && d.cols() == s.cols()) { for (int c = 0; c < d.cols(); ++c) for (int r = 0; r < d.rows(); ++r) d.coeffRef(r, c) = s.coeff(r, c); }
        else {
            assert((d.rows() == 1 || d.cols() == 1) && (s.rows() == 1 || s.cols() == 1) && d.rows() * d.cols() == s.rows() * s.cols());
            for (int i = 0; i < d.rows() * d.cols(); ++i) linRef(i) = o.lin(i);
        }
    }
    template <class O> D& operator+=(const DenseBase<O>& o) {
        D& d = derived(); const O& s = o.derived();
        if (d.rows() == s.rows() && d.cols() == s.cols()) { for (int c = 0; c < d.cols(); ++c) for (int r = 0; r < d.rows(); ++r) d.coeffRef(r, c) += s.coeff(r, c); }
        else { assert(this->size() == o.size()); for (int i = 0; i < this->size(); ++i) linRef(i) += o.lin(i); }
        return d;
    }
    template <class O> D& operator-=(const DenseBase<O>& o) {
        D& d = derived(); const O& s = o.derived();
        if (d.rows() == s.rows() && d.cols() == s.cols()) { for (int c = 0; c < d.cols(); ++c) for (int r = 0; r < d.rows(); ++r) d.coeffRef(r, c) -= s.coeff(r, c); }
        else { assert(this->size() == o.size()); for (int i = 0; i < this->size(); ++i) linRef(i) -= o.lin(i); }
        return d;
    }
    template <class S, class = typename std::enable_if<std::is_arithmetic<S>::value>::type> D& operator*=(S s) {
        D& d = derived();
        for (int c = 0; c < d.cols(); ++c) for (int r = 0; r < d.rows(); ++r) d.coeffRef(r, c) *= static_cast<Scalar>(s);
        return d;
    }
    template <class S, class = typename std::enable_if<std::is_arithmetic<S>::value>::type> D& operator/=(S s) {
        D& d = derived();
        for (int c = 0; c < d.cols(); ++c) for (int r = 0; r < d.rows(); ++r) d.coeffRef(r, c) /= static_cast<Scalar>(s);
        return d;
    }
    D& setConstant(Scalar v) {
        D& d = derived();
        for (int c = 0; c < d.cols(); ++c) for (int r = 0; r < d.rows(); ++r) d.coeffRef(r, c) = v;
        return d;
    }
    D& setZero() { return setConstant(Scalar(0)); }
    D& setOnes() { return setConstant(Scalar(1)); }
    D& setIdentity() {
        D& d = derived();
        for (int c = 0; c < d.cols(); ++c) for (int r = 0; r < d.rows(); ++r) d.coeffRef(r, c) = (r == c) ? Scalar(1) : Scalar(0);
        return d;
    }
    void normalize() { Scalar n = this->norm(); if (n > Scalar(0)) *this /= n; }
    CommaInit<D> operator<<(Scalar v) { return CommaInit<D>(derived(), v); }

    // views that alias this object's storage (const-ness is not tracked: the shim only has to compile correct code)
    Block<Scalar, 1, Cols> row(int r) const { const D& d = derived(); return Block<Scalar, 1, Cols>(d.ptr() + r * d.rstride(), 1, d.cols(), d.rstride(), d.cstride()); }
    Block<Scalar, Rows, 1> col(int c) const { const D& d = derived(); return Block<Scalar, Rows, 1>(d.ptr() + c * d.cstride(), d.rows(), 1, d.rstride(), d.cstride()); }
    template <int BR, int BC> Block<Scalar, BR, BC> block(int r, int c) const {
        const D& d = derived(); return Block<Scalar, BR, BC>(d.ptr() + r * d.rstride() + c * d.cstride(), BR, BC, d.rstride(), d.cstride());
    }
    template <int N> Block<Scalar, (Cols == 1 ? N : 1), (Cols == 1 ? 1 : N)> segment(int i) const {
        const D& d = derived();
        if (Cols == 1) return Block<Scalar, (Cols == 1 ? N : 1), (Cols == 1 ? 1 : N)>(d.ptr() + i * d.rstride(), N, 1, d.rstride(), d.cstride());
        return Block<Scalar, (Cols == 1 ? N : 1), (Cols == 1 ? 1 : N)>(d.ptr() + i * d.cstride(), 1, N, d.rstride(), d.cstride());
    }
};

// ------------------------------------------------------------------------------------------------ storage
template <class T, int R, int C, bool Fixed = (R != Dynamic && C != Dynamic)> struct Storage;
template <class T, int R, int C> struct Storage<T, R, C, true> {
    T a[R * C];
    Storage() { for (int i = 0; i < R * C; ++i) a[i] = T(0); }
    T* data() { return a; } const T* data() const { return a; }
    int rows() const { return R; } int cols() const { return C; }
    void resize(int r, int c) { assert(r == R && c == C); (void)r; (void)c; }
};
template <class T, int R, int C> struct Storage<T, R, C, false> {
    std::vector<T> a; int r_, c_;
    mutable T fixcache_[4];                                      // see the header comment (minBound()/maxBound())
    Storage() : r_(R == Dynamic ? 0 : R), c_(C == Dynamic ? 0 : C) {}
    T* data() { return a.data(); } const T* data() const { return a.data(); }
    int rows() const { return r_; } int cols() const { return c_; }
    void resize(int r, int c) {
        assert((R == Dynamic || r == R) && (C == Dynamic || c == C));
        if (r != r_ || c != c_ || a.size() != static_cast<size_t>(r) * c) { r_ = r; c_ = c; a.assign(static_cast<size_t>(r) * c, T(0)); }
    }
};

template <class T, int R, int C> class Matrix : public DenseMut<Matrix<T, R, C>> {
    Storage<T, R, C> s_;
public:
    typedef T Scalar;
    typedef DenseMut<Matrix<T, R, C>> Mut;
    enum { RowsAtCompileTime = R, ColsAtCompileTime = C };

    Matrix() {}
    Matrix(const Matrix& o) : s_(o.s_) {}
    explicit Matrix(int n) { if (R == Dynamic && C != Dynamic && C != 1) resize(n, C); else if (C == 1) resize(n, 1); else resize(1, n); }
    // (a, b): two coefficients for a fixed 2-vector, otherwise (rows, cols)
    template <class A, class B, class = typename std::enable_if<std::is_arithmetic<A>::value && std::is_arithmetic<B>::value>::type>
    Matrix(A a, B b) {
        if (R != Dynamic && C != Dynamic && R * C == 2) { s_.data()[0] = static_cast<T>(a); s_.data()[1] = static_cast<T>(b); }
        else resize(static_cast<int>(a), static_cast<int>(b));
    }
    Matrix(T a, T b, T c) { resize(C == 1 ? 3 : 1, C == 1 ? 1 : 3); s_.data()[0] = a; s_.data()[1] = b; s_.data()[2] = c; }
    Matrix(T a, T b, T c, T d) { resize(C == 1 ? 4 : 1, C == 1 ? 1 : 4); s_.data()[0] = a; s_.data()[1] = b; s_.data()[2] = c; s_.data()[3] = d; }
    template <class O> Matrix(const DenseBase<O>& o) { sizeLike(o.derived().rows(), o.derived().cols()); this->assignFrom(o); }

    Matrix& operator=(const Matrix& o) { s_ = o.s_; return *this; }
    template <class O> Matrix& operator=(const DenseBase<O>& o) {
        // evaluate first: the right-hand side may alias this matrix through a block
        Matrix<T, traits<O>::Rows, traits<O>::Cols> tmp = o.eval();
        sizeLike(tmp.rows(), tmp.cols()); this->assignFrom(tmp); return *this;
    }

    // const T_N& view of a dynamic column vector (RegularGrid.h:54-55)
    template <int R2, int RR = R, int CC = C, class = typename std::enable_if<RR == Dynamic && CC == 1 && R2 != Dynamic && (R2 <= 4)>::type>
    operator const Matrix<T, R2, 1>&() const {
        assert(rows() == R2);
        for (int i = 0; i < R2; ++i) s_.fixcache_[i] = s_.data()[i];
        return *reinterpret_cast<const Matrix<T, R2, 1>*>(s_.fixcache_);
    }

    int rows() const { return s_.rows(); }
    int cols() const { return s_.cols(); }
    void resize(int r, int c) { s_.resize(r, c); }
    void resize(int n) { if (C == 1) s_.resize(n, 1); else if (R == 1) s_.resize(1, n); else s_.resize(n, C); }
    T coeff(int r, int c) const { return s_.data()[r + static_cast<size_t>(c) * s_.rows()]; }
    T& coeffRef(int r, int c) { return s_.data()[r + static_cast<size_t>(c) * s_.rows()]; }
    T* data() { return s_.data(); } const T* data() const { return s_.data(); }
    T* ptr() const { return const_cast<T*>(s_.data()); }
    int rstride() const { return 1; }
    int cstride() const { return s_.rows(); }

    static Matrix Zero() { Matrix m; m.setZero(); return m; }
    static Matrix Zero(int n) { Matrix m(n); m.setZero(); return m; }
    static Matrix Zero(int r, int c) { Matrix m; m.resize(r, c); m.setZero(); return m; }
    static Matrix Ones() { Matrix m; m.setOnes(); return m; }
    static Matrix Ones(int n) { Matrix m(n); m.setOnes(); return m; }
    static Matrix Constant(T v) { Matrix m; m.setConstant(v); return m; }
    static Matrix Identity() { Matrix m; m.setIdentity(); return m; }
    static Matrix Random() {                                      // uniform in [-1, 1] from std::rand()
        Matrix m;
        for (int c = 0; c < m.cols(); ++c) for (int r = 0; r < m.rows(); ++r)
            m.coeffRef(r, c) = static_cast<T>(2.0 * std::rand() / RAND_MAX - 1.0);
        return m;
    }
private:
    // take the shape of the source; a vector keeps its own orientation when the source is the other kind of vector
    void sizeLike(int r, int c) {
        bool fits = (R == Dynamic || R == r) && (C == Dynamic || C == c);
        if (fits) resize(r, c);
        else { assert((r == 1 || c == 1) && (R == 1 || C == 1)); if (C == 1) resize(r * c, 1); else resize(1, r * c); }
    }
};

template <class T, int R, int C> class Block : public DenseMut<Block<T, R, C>> {
    T* p_; int r_, c_, rs_, cs_;
public:
    typedef T Scalar;
    Block(T* p, int r, int c, int rs, int cs) : p_(p), r_(r), c_(c), rs_(rs), cs_(cs) {}
    Block(const Block& o) : p_(o.p_), r_(o.r_), c_(o.c_), rs_(o.rs_), cs_(o.cs_) {}
    int rows() const { return r_; }
    int cols() const { return c_; }
    T coeff(int r, int c) const { return p_[static_cast<size_t>(r) * rs_ + static_cast<size_t>(c) * cs_]; }
    T& coeffRef(int r, int c) { return p_[static_cast<size_t>(r) * rs_ + static_cast<size_t>(c) * cs_]; }
    T* ptr() const { return p_; }
    int rstride() const { return rs_; }
    int cstride() const { return cs_; }
    // assignment writes THROUGH the view (never rebinds it)
    Block& operator=(const Block& o) { Matrix<T, R, C> tmp = o.eval(); this->assignFrom(tmp); return *this; }
    template <class O> Block& operator=(const DenseBase<O>& o) { Matrix<T, traits<O>::Rows, traits<O>::Cols> tmp = o.eval(); this->assignFrom(tmp); return *this; }
};

// ------------------------------------------------------------------------------------------------ dense arithmetic
template <class A, class B>
Matrix<typename A::Scalar, pick<traits<A>::Rows, traits<B>::Rows>::value, pick<traits<A>::Cols, traits<B>::Cols>::value>
operator+(const DenseBase<A>& a, const DenseBase<B>& b) {
    Matrix<typename A::Scalar, pick<traits<A>::Rows, traits<B>::Rows>::value, pick<traits<A>::Cols, traits<B>::Cols>::value> m;
    const A& x = a.derived(); const B& y = b.derived();
    assert(x.rows() == y.rows() && x.cols() == y.cols());
    m.resize(x.rows(), x.cols());
    for (int c = 0; c < m.cols(); ++c) for (int r = 0; r < m.rows(); ++r) m.coeffRef(r, c) = x.coeff(r, c) + y.coeff(r, c);
    return m;
}
template <class A, class B>
Matrix<typename A::Scalar, pick<traits<A>::Rows, traits<B>::Rows>::value, pick<traits<A>::Cols, traits<B>::Cols>::value>
operator-(const DenseBase<A>& a, const DenseBase<B>& b) {
    Matrix<typename A::Scalar, pick<traits<A>::Rows, traits<B>::Rows>::value, pick<traits<A>::Cols, traits<B>::Cols>::value> m;
    const A& x = a.derived(); const B& y = b.derived();
    assert(x.rows() == y.rows() && x.cols() == y.cols());
    m.resize(x.rows(), x.cols());
    for (int c = 0; c < m.cols(); ++c) for (int r = 0; r < m.rows(); ++r) m.coeffRef(r, c) = x.coeff(r, c) - y.coeff(r, c);
    return m;
}
template <class A> typename DenseBase<A>::Plain operator-(const DenseBase<A>& a) {
    typename DenseBase<A>::Plain m = a.eval();
    for (int c = 0; c < m.cols(); ++c) for (int r = 0; r < m.rows(); ++r) m.coeffRef(r, c) = -m.coeff(r, c);
    return m;
}
template <class A, class S, class = typename std::enable_if<std::is_arithmetic<S>::value>::type>
typename DenseBase<A>::Plain operator*(const DenseBase<A>& a, S s) { typename DenseBase<A>::Plain m = a.eval(); m *= s; return m; }
template <class A, class S, class = typename std::enable_if<std::is_arithmetic<S>::value>::type>
typename DenseBase<A>::Plain operator*(S s, const DenseBase<A>& a) {
    typename DenseBase<A>::Plain m = a.eval();
    for (int c = 0; c < m.cols(); ++c) for (int r = 0; r < m.rows(); ++r) m.coeffRef(r, c) = static_cast<typename A::Scalar>(s) * m.coeff(r, c);
    return m;
}
template <class A, class S, class = typename std::enable_if<std::is_arithmetic<S>::value>::type>
typename DenseBase<A>::Plain operator/(const DenseBase<A>& a, S s) { typename DenseBase<A>::Plain m = a.eval(); m /= s; return m; }

template <class A, class B>
Matrix<typename A::Scalar, traits<A>::Rows, traits<B>::Cols> operator*(const DenseBase<A>& a, const DenseBase<B>& b) {
    const A& x = a.derived(); const B& y = b.derived();
    assert(x.cols() == y.rows());
    Matrix<typename A::Scalar, traits<A>::Rows, traits<B>::Cols> m; m.resize(x.rows(), y.cols());
    for (int c = 0; c < y.cols(); ++c) for (int r = 0; r < x.rows(); ++r) {
        typename A::Scalar s = 0;
        for (int k = 0; k < x.cols(); ++k) s += x.coeff(r, k) * y.coeff(k, c);
        m.coeffRef(r, c) = s;
    }
    return m;
}

template <class T> class DiagWrap { public: std::vector<T> v; int size() const { return static_cast<int>(v.size()); } };
template <class A> typename DenseBase<A>::Plain operator*(const DenseBase<A>& a, const DiagWrap<typename A::Scalar>& d) {
    typename DenseBase<A>::Plain m = a.eval(); assert(m.cols() == d.size());
    for (int c = 0; c < m.cols(); ++c) for (int r = 0; r < m.rows(); ++r) m.coeffRef(r, c) *= d.v[c];
    return m;
}
template <class A> typename DenseBase<A>::Plain operator*(const DiagWrap<typename A::Scalar>& d, const DenseBase<A>& a) {
    typename DenseBase<A>::Plain m = a.eval(); assert(m.rows() == d.size());
    for (int c = 0; c < m.cols(); ++c) for (int r = 0; r < m.rows(); ++r) m.coeffRef(r, c) = d.v[r] * m.coeff(r, c);
    return m;
}

// .array() ... .rowwise().sum().matrix()
template <class T, int R, int C> struct RowwiseOf {
    const Matrix<T, R, C>& m;
    Arr<T, R, 1> sum() const {
        Arr<T, R, 1> a; a.m.resize(m.rows(), 1);
        for (int r = 0; r < m.rows(); ++r) { T s = 0; for (int c = 0; c < m.cols(); ++c) s += m.coeff(r, c); a.m.coeffRef(r, 0) = s; }
        return a;
    }
};
template <class T, int R, int C> struct Arr {
    Matrix<T, R, C> m;
    RowwiseOf<T, R, C> rowwise() const { return RowwiseOf<T, R, C>{m}; }
    const Matrix<T, R, C>& matrix() const { return m; }
    T sum() const { return m.sum(); }
};
template <class T, int R, int C, int R2, int C2>
Arr<T, pick<R, R2>::value, pick<C, C2>::value> operator*(const Arr<T, R, C>& a, const Arr<T, R2, C2>& b) {
    Arr<T, pick<R, R2>::value, pick<C, C2>::value> o; o.m = a.m.cwiseProduct(b.m); return o;
}

template <class D> std::ostream& operator<<(std::ostream& os, const DenseBase<D>& a) {
    for (int r = 0; r < a.derived().rows(); ++r) { for (int c = 0; c < a.derived().cols(); ++c) os << (c ? " " : "") << a.derived().coeff(r, c); if (r + 1 < a.derived().rows()) os << "\n"; }
    return os;
}

typedef Matrix<double, Dynamic, Dynamic> MatrixXd;
typedef Matrix<int, Dynamic, Dynamic> MatrixXi;
typedef Matrix<double, Dynamic, 3> MatrixX3d;
typedef Matrix<int, Dynamic, 3> MatrixX3i;
typedef Matrix<double, Dynamic, 1> VectorXd;
typedef Matrix<int, Dynamic, 1> VectorXi;
typedef Matrix<double, 2, 2> Matrix2d;
typedef Matrix<double, 3, 3> Matrix3d;
typedef Matrix<double, 4, 4> Matrix4d;
typedef Matrix<double, 2, 1> Vector2d;
typedef Matrix<double, 3, 1> Vector3d;
typedef Matrix<double, 4, 1> Vector4d;
typedef Matrix<float, 3, 1> Vector3f;
typedef Matrix<float, 4, 1> Vector4f;
typedef Matrix<int, 2, 1> Vector2i;
typedef Matrix<int, 3, 1> Vector3i;
typedef Matrix<double, 1, 3> RowVector3d;

// ------------------------------------------------------------------------------------------------ JacobiSVD (2x2, 3x3)
// Two-sided (Kogbetliantz) Jacobi: every off-diagonal pair is annihilated by a left and a right plane rotation obtained
// from the 2x2 SVD of the pair's sub-block; sweeps repeat until all off-diagonal entries are negligible.  Then the signs are
// moved into U and the triplets sorted by decreasing singular value.
template <class MT> class JacobiSVD {
    enum { N = MT::RowsAtCompileTime };
    typedef Matrix<double, N, N> Mat;
    typedef Matrix<double, N, 1> Vec;
    Mat U_, V_; Vec s_;

    static void rotSym(double a, double b, double d, double& c, double& s) {
        // rotation J = [c s; -s c] with J^T [a b; b d] J diagonal
        if (b == 0.0) { c = 1.0; s = 0.0; return; }
        double tau = (d - a) / (2.0 * b);
        double t = (tau >= 0.0 ? 1.0 : -1.0) / (std::abs(tau) + std::sqrt(1.0 + tau * tau));
        c = 1.0 / std::sqrt(1.0 + t * t); s = t * c;
    }
public:
    JacobiSVD() {}
    template <class D> JacobiSVD(const DenseBase<D>& A, unsigned = 0) { compute(A); }
    template <class D> JacobiSVD& compute(const DenseBase<D>& Ain, unsigned = 0) {
        Mat W = Ain.eval(); U_.setIdentity(); V_.setIdentity();
        double scale = W.cwiseAbs().maxCoeff();
        if (scale == 0.0) scale = 1.0;
        W /= scale;
        const double eps = 2.220446049250313e-16, tiny = 2.2250738585072014e-308;
        for (int sweep = 0; sweep < 64; ++sweep) {
            bool done = true;
            for (int p = 1; p < N; ++p) for (int q = 0; q < p; ++q) {
                double maxd = 0.0;
                for (int k = 0; k < N; ++k) maxd = std::max(maxd, std::abs(W(k, k)));
                double thr = std::max(tiny, 2.0 * eps * maxd);
                if (std::abs(W(p, q)) <= thr && std::abs(W(q, p)) <= thr) continue;
                done = false;
                // 2x2 block [[wpp wpq][wqp wqq]]: first a left rotation that makes it symmetric, then a symmetric Jacobi rotation
                double wpp = W(p, p), wpq = W(p, q), wqp = W(q, p), wqq = W(q, q);
                double t = wpp + wqq, d = wqp - wpq, c1, s1;
                if (std::abs(d) < tiny) { c1 = 1.0; s1 = 0.0; }
                else { double u = t / d, tmp = std::sqrt(1.0 + u * u); s1 = 1.0 / tmp; c1 = u / tmp; }
                // rows p,q <- R1 * rows, R1 = [c1 s1; -s1 c1]
                double a = c1 * wpp + s1 * wqp, b = c1 * wpq + s1 * wqq, dd = -s1 * wpq + c1 * wqq;
                double c2, s2; rotSym(a, b, dd, c2, s2);
                // left rotation L = J^T R1, right rotation J:   L * block * J = diagonal
                double lc = c2 * c1 + s2 * s1, ls = c2 * s1 - s2 * c1;     // L = [lc ls; -ls lc]
                for (int k = 0; k < N; ++k) {
                    double xp = W(p, k), xq = W(q, k);
                    W(p, k) = lc * xp + ls * xq; W(q, k) = -ls * xp + lc * xq;
                }
                for (int k = 0; k < N; ++k) {
                    double xp = W(k, p), xq = W(k, q);
                    W(k, p) = c2 * xp - s2 * xq; W(k, q) = s2 * xp + c2 * xq;
                }
                // accumulate U <- U * L^T, V <- V * J
                for (int k = 0; k < N; ++k) {
                    double xp = U_(k, p), xq = U_(k, q);
                    U_(k, p) = lc * xp + ls * xq; U_(k, q) = -ls * xp + lc * xq;
                    double yp = V_(k, p), yq = V_(k, q);
                    V_(k, p) = c2 * yp - s2 * yq; V_(k, q) = s2 * yp + c2 * yq;
                }
            }
            if (done) break;
        }
        for (int i = 0; i < N; ++i) {
            double a = W(i, i);
            if (a < 0.0) { a = -a; for (int k = 0; k < N; ++k) U_(k, i) = -U_(k, i); }
            s_[i] = a * scale;
        }
        for (int i = 0; i < N; ++i) {                              // selection sort, decreasing
            int m = i;
            for (int j = i + 1; j < N; ++j) if (s_[j] > s_[m]) m = j;
            if (m != i) {
                std::swap(s_[i], s_[m]);
                for (int k = 0; k < N; ++k) { std::swap(U_(k, i), U_(k, m)); std::swap(V_(k, i), V_(k, m)); }
            }
        }
        return *this;
    }
    const Mat& matrixU() const { return U_; }
    const Mat& matrixV() const { return V_; }
    const Vec& singularValues() const { return s_; }
};

// ------------------------------------------------------------------------------------------------ sparse
template <class T> class Triplet {
    int r_, c_; T v_;
public:
    Triplet() : r_(0), c_(0), v_(0) {}
    Triplet(int r, int c, T v = T(0)) : r_(r), c_(c), v_(v) {}
    int row() const { return r_; } int col() const { return c_; } T value() const { return v_; }
};

template <class T> class SparseMatrix;

// op(A) * scalar * diag(d): what the reference builds before multiplying by a dense operand
template <class T> struct SpExpr {
    const SparseMatrix<T>* m; bool trans; T scale; bool hasd; std::vector<T> d;
    int rows() const { return trans ? m->cols() : m->rows(); }
    int cols() const { return trans ? m->rows() : m->cols(); }
};

template <class T> class SparseMatrix {
    int rows_, cols_;
    std::vector<int> colptr_, rowidx_;
    std::vector<T> val_;
public:
    typedef T Scalar;
    SparseMatrix() : rows_(0), cols_(0), colptr_(1, 0) {}
    SparseMatrix(int r, int c) : rows_(r), cols_(c), colptr_(c + 1, 0) {}
    int rows() const { return rows_; }
    int cols() const { return cols_; }
    int outerSize() const { return cols_; }
    int innerSize() const { return rows_; }
    long nonZeros() const { return static_cast<long>(val_.size()); }
    void resize(int r, int c) { rows_ = r; cols_ = c; colptr_.assign(c + 1, 0); rowidx_.clear(); val_.clear(); }
    void setZero() { colptr_.assign(cols_ + 1, 0); rowidx_.clear(); val_.clear(); }
    const std::vector<int>& colptr() const { return colptr_; }
    const std::vector<int>& rowidx() const { return rowidx_; }
    const std::vector<T>& values() const { return val_; }

    // compressed column storage, rows ascending inside a column, duplicates summed.  Two counting passes (O(nnz), like Eigen's
    // own two-transpose scheme); columns are only sorted / merged when some entries did not arrive in strictly ascending row order.
    template <class It> void setFromTriplets(It b, It e) {
        std::vector<int> cnt(cols_ + 1, 0);
        for (It it = b; it != e; ++it) { assert(it->row() >= 0 && it->row() < rows_ && it->col() >= 0 && it->col() < cols_); ++cnt[it->col() + 1]; }
        for (int c = 0; c < cols_; ++c) cnt[c + 1] += cnt[c];
        rowidx_.assign(cnt[cols_], 0); val_.assign(cnt[cols_], T(0));
        std::vector<int> fill(cnt.begin(), cnt.end() - 1);
        bool ordered = true;
        for (It it = b; it != e; ++it) {
            const int c = it->col(), k = fill[c]++;
            rowidx_[k] = it->row(); val_[k] = it->value();
            if (k > cnt[c] && rowidx_[k - 1] >= rowidx_[k]) ordered = false;
        }
        colptr_ = cnt;
        if (ordered) return;
        std::vector<int> ri; std::vector<T> va; ri.reserve(rowidx_.size()); va.reserve(val_.size());
        std::vector<std::pair<int, T>> tmp;
        for (int c = 0; c < cols_; ++c) {
            tmp.clear();
            for (int k = cnt[c]; k < cnt[c + 1]; ++k) tmp.push_back(std::make_pair(rowidx_[k], val_[k]));
            std::stable_sort(tmp.begin(), tmp.end(), [](const std::pair<int, T>& x, const std::pair<int, T>& y) { return x.first < y.first; });
            const int start = static_cast<int>(ri.size());
            for (size_t k = 0; k < tmp.size(); ++k) {
                if (static_cast<int>(ri.size()) > start && ri.back() == tmp[k].first) va.back() += tmp[k].second;
                else { ri.push_back(tmp[k].first); va.push_back(tmp[k].second); }
            }
            colptr_[c] = start;
        }
        colptr_[cols_] = static_cast<int>(ri.size());
        rowidx_.swap(ri); val_.swap(va);
    }
    T coeff(int r, int c) const {
        for (int k = colptr_[c]; k < colptr_[c + 1]; ++k) if (rowidx_[k] == r) return val_[k];
        return T(0);
    }
    SpExpr<T> transpose() const { return SpExpr<T>{this, true, T(1), false, std::vector<T>()}; }
    SpExpr<T> expr() const { return SpExpr<T>{this, false, T(1), false, std::vector<T>()}; }

    class InnerIterator {
        const SparseMatrix& m_; int outer_, k_, end_;
    public:
        InnerIterator(const SparseMatrix& m, int outer) : m_(m), outer_(outer), k_(m.colptr_[outer]), end_(m.colptr_[outer + 1]) {}
        InnerIterator& operator++() { ++k_; return *this; }
        operator bool() const { return k_ < end_; }
        int row() const { return m_.rowidx_[k_]; }
        int col() const { return outer_; }
        int index() const { return m_.rowidx_[k_]; }
        T value() const { return m_.val_[k_]; }
    };
};

template <class T, class S, class = typename std::enable_if<std::is_arithmetic<S>::value>::type>
SpExpr<T> operator*(S s, SpExpr<T> e) { e.scale *= static_cast<T>(s); return e; }
template <class T, class S, class = typename std::enable_if<std::is_arithmetic<S>::value>::type>
SpExpr<T> operator*(SpExpr<T> e, S s) { e.scale *= static_cast<T>(s); return e; }
template <class T, class S, class = typename std::enable_if<std::is_arithmetic<S>::value>::type>
SpExpr<T> operator*(S s, const SparseMatrix<T>& m) { return s * m.expr(); }
template <class T, class S, class = typename std::enable_if<std::is_arithmetic<S>::value>::type>
SpExpr<T> operator*(const SparseMatrix<T>& m, S s) { return s * m.expr(); }
template <class T> SpExpr<T> operator*(SpExpr<T> e, const DiagWrap<T>& d) {
    assert(e.cols() == d.size());
    if (e.hasd) { for (size_t i = 0; i < e.d.size(); ++i) e.d[i] *= d.v[i]; } else { e.d = d.v; e.hasd = true; }
    return e;
}
template <class T> SpExpr<T> operator*(const SparseMatrix<T>& m, const DiagWrap<T>& d) { return m.expr() * d; }

template <class T, class B>
Matrix<T, Dynamic, traits<B>::Cols> operator*(const SpExpr<T>& e, const DenseBase<B>& b) {
    Matrix<T, traits<B>::Rows, traits<B>::Cols> M = b.eval();
    assert(e.cols() == M.rows());
    const int nc = M.cols();
    Matrix<T, Dynamic, traits<B>::Cols> out; out.resize(e.rows(), nc);
    const std::vector<int>& cp = e.m->colptr(); const std::vector<int>& ri = e.m->rowidx(); const std::vector<T>& va = e.m->values();
    for (int j = 0; j < e.m->cols(); ++j)
        for (int k = cp[j]; k < cp[j + 1]; ++k) {
            int i = ri[k];
            int orow = e.trans ? j : i, inner = e.trans ? i : j;
            T w = va[k];
            if (e.hasd) w *= e.d[inner];
            for (int c = 0; c < nc; ++c) out.coeffRef(orow, c) += w * M.coeff(inner, c);
        }
    if (e.scale != T(1)) out *= e.scale;
    return out;
}
template <class T, class B>
Matrix<T, Dynamic, traits<B>::Cols> operator*(const SparseMatrix<T>& m, const DenseBase<B>& b) { return m.expr() * b; }

}  // namespace Eigen
