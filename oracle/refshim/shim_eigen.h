// shim_eigen.h -- a minimal stand-in for the subset of Eigen that
// /root/reference/include/RegisterPhotoICP.h and Miscellaneous.h use, so that the reference header
// itself can be compiled in this container (no Eigen here, no network) into oracle/_ref/.
//
// *** TEST INFRASTRUCTURE ONLY (see oracle/README in DESIGN.md section 2). ***
//
// Written from scratch; everything is evaluated eagerly (no expression templates).  The arithmetic
// order of the few operations whose rounding matters on the spherical path follows what Eigen 3.2
// (the "current version" of 2014 the reference names, doc/mainpage.h:47) does for small fixed sizes:
//   * fixed-size products: coefficient (i,j) = sum_k a(i,k)*b(k,j), k ascending, left to right
//   * norm()/squaredNorm()/sum(): unrolled redux, halves split (n/2 | n - n/2) recursively
//   * general inverse(): partial-pivot LU (r360_inverse6 of gn_math.h for 6x6, Gauss-Jordan otherwise)
//   * rank(): the MRPT Eigen plugin = ColPivHouseholderQR::rank() (r360_rank6 for 6x6)
// These are the same restatements the oracle uses (rgbd360_b200/csrc/gn_math.h).
#pragma once
#include <cassert>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <cstdio>
#include <string>
#include <iostream>
#include <vector>
#include <memory>
#include "../../rgbd360_b200/csrc/gn_math.h"

#define EIGEN_MAKE_ALIGNED_OPERATOR_NEW
#define EIGEN_WORLD_VERSION 3

namespace Eigen {

enum { Dynamic = -1 };

template <typename T, int N>
struct ShimStore {
    T a[N > 0 ? N : 1];
    void resize(size_t) {}
    T* ptr() { return a; }
    const T* ptr() const { return a; }
};
template <typename T>
struct ShimStore<T, -1> {
    std::vector<T> v;
    void resize(size_t n) { v.resize(n); }
    T* ptr() { return v.data(); }
    const T* ptr() const { return v.data(); }
};

template <typename T>
T shim_redux(const T* p, int n) {   // Eigen's redux_novec_unroller: halves, recursively
    if (n == 1) return p[0];
    const int h = n / 2;
    return shim_redux(p, h) + shim_redux(p + h, n - h);
}

template <typename T, int R, int C>
class Matrix;
template <typename T>
class BlockRef;

template <typename T, int R, int C>
struct CommaInit {
    Matrix<T, R, C>* m;
    int k;
    template <typename U>
    CommaInit& operator,(const U& v) {
        const int r = k / m->cols(), c = k % m->cols();
        (*m)(r, c) = (T)v;
        ++k;
        return *this;
    }
};

template <typename T, int R, int C>
class Matrix {
public:
    typedef T Scalar;
    enum { RowsAtCompileTime = R, ColsAtCompileTime = C, Fixed = (R > 0 && C > 0) };
    ShimStore<T, (R > 0 && C > 0) ? R * C : -1> s;
    int r_, c_;

    Matrix() : r_(R > 0 ? R : 0), c_(C > 0 ? C : 0) { s.resize((size_t)r_ * c_); }
    explicit Matrix(int n) {
        if (Fixed) { r_ = R; c_ = C; }
        else if (C == 1) { r_ = n; c_ = 1; }
        else { r_ = 1; c_ = n; }
        s.resize((size_t)r_ * c_);
    }
    // (rows, cols) for dynamic types, (x, y) for fixed 2-vectors
    Matrix(T x, T y) {
        if (Fixed) { r_ = R; c_ = C; s.ptr()[0] = x; s.ptr()[1] = y; }
        else { r_ = R > 0 ? R : (int)x; c_ = C > 0 ? C : (int)y; s.resize((size_t)r_ * c_); }
    }
    Matrix(T x, T y, T z) : r_(R), c_(C) { s.ptr()[0] = x; s.ptr()[1] = y; s.ptr()[2] = z; }
    Matrix(T x, T y, T z, T w) : r_(R), c_(C) { s.ptr()[0] = x; s.ptr()[1] = y; s.ptr()[2] = z; s.ptr()[3] = w; }
    Matrix(const Matrix& o) : s(o.s), r_(o.r_), c_(o.c_) {}
    template <int R2, int C2>
    Matrix(const Matrix<T, R2, C2>& o) { assign(o); }
    Matrix& operator=(const Matrix& o) { s = o.s; r_ = o.r_; c_ = o.c_; return *this; }
    template <int R2, int C2>
    Matrix& operator=(const Matrix<T, R2, C2>& o) { assign(o); return *this; }

    template <int R2, int C2>
    void assign(const Matrix<T, R2, C2>& o) {
        r_ = o.rows(); c_ = o.cols();
        // Eigen transposes implicitly when a row vector is assigned to a column vector (and back)
        if ((R == 1 && r_ != 1 && c_ == 1) || (C == 1 && c_ != 1 && r_ == 1)) { int t = r_; r_ = c_; c_ = t; }
        assert((R < 0 || R == r_) && (C < 0 || C == c_));
        s.resize((size_t)r_ * c_);
        for (int i = 0; i < r_ * c_; ++i) s.ptr()[i] = o.data()[i];
    }

    int rows() const { return r_; }
    int cols() const { return c_; }
    int size() const { return r_ * c_; }
    T* data() { return s.ptr(); }
    const T* data() const { return s.ptr(); }
    void resize(int n) { if (C == 1) { r_ = n; c_ = 1; } else { r_ = 1; c_ = n; } s.resize((size_t)n); }
    void resize(int r, int c) { r_ = r; c_ = c; s.resize((size_t)r * c); }
    void setZero() { for (int i = 0; i < size(); ++i) data()[i] = T(0); }

    // column-major storage, like Eigen's default
    T& operator()(int i, int j) { return s.ptr()[i + (size_t)j * r_]; }
    const T& operator()(int i, int j) const { return s.ptr()[i + (size_t)j * r_]; }
    T& operator()(int i) { return s.ptr()[i]; }
    const T& operator()(int i) const { return s.ptr()[i]; }
    T& operator[](int i) { return s.ptr()[i]; }
    const T& operator[](int i) const { return s.ptr()[i]; }
    T& coeffRef(int i, int j) { return (*this)(i, j); }
    T coeff(int i, int j) const { return (*this)(i, j); }

    static Matrix Zero() { Matrix m; m.setZero(); return m; }
    static Matrix Zero(int n) { Matrix m(n); m.setZero(); return m; }
    static Matrix Zero(int r, int c) { Matrix m; m.resize(r, c); m.setZero(); return m; }
    static Matrix Ones() { Matrix m; for (int i = 0; i < m.size(); ++i) m.data()[i] = T(1); return m; }
    static Matrix Identity() {
        Matrix m; m.setZero();
        for (int i = 0; i < (m.r_ < m.c_ ? m.r_ : m.c_); ++i) m(i, i) = T(1);
        return m;
    }
    static Matrix Identity(int r, int c) {
        Matrix m = Zero(r, c);
        for (int i = 0; i < (r < c ? r : c); ++i) m(i, i) = T(1);
        return m;
    }
    void setIdentity() { *this = Identity(r_, c_); }
    // MRPT's Eigen plugin: MatrixBase::loadFromTextFile (Calib360.h:128) -- a whitespace-separated text matrix,
    // one row per line; fixed-size matrices must find exactly their size.
    void loadFromTextFile(const std::string& file) {
        FILE* f = fopen(file.c_str(), "r");
        if (!f) { std::cerr << "refshim: loadFromTextFile: cannot open " << file << "\n"; std::abort(); }
        std::vector<double> v;
        double x;
        while (fscanf(f, "%lf", &x) == 1) v.push_back(x);
        fclose(f);
        if (!Fixed || (int)v.size() != r_ * c_) { std::cerr << "refshim: loadFromTextFile: " << file << " does not hold a " << r_ << "x" << c_ << " matrix\n"; std::abort(); }
        for (int i = 0; i < r_; ++i) for (int j = 0; j < c_; ++j) (*this)(i, j) = (T)v[(size_t)i * c_ + j];
    }

    template <typename U>
    CommaInit<T, R, C> operator<<(const U& v) {
        (*this)(0, 0) = (T)v;
        CommaInit<T, R, C> ci = { this, 1 };
        return ci;
    }

    Matrix<T, -1, -1> block(int i, int j, int nr, int nc) const {
        Matrix<T, -1, -1> b; b.resize(nr, nc);
        for (int c = 0; c < nc; ++c) for (int r = 0; r < nr; ++r) b(r, c) = (*this)(i + r, j + c);
        return b;
    }
    BlockRef<T> block(int i, int j, int nr, int nc);
    template <int NR, int NC>
    Matrix<T, NR, NC> block(int i, int j) const {
        Matrix<T, NR, NC> b;
        for (int c = 0; c < NC; ++c) for (int r = 0; r < NR; ++r) b(r, c) = (*this)(i + r, j + c);
        return b;
    }
    Matrix<T, R, 1> col(int j) const { Matrix<T, R, 1> v; v.resize(r_, 1); for (int i = 0; i < r_; ++i) v(i) = (*this)(i, j); return v; }
    Matrix<T, 1, C> row(int i) const { Matrix<T, 1, C> v; v.resize(1, c_); for (int j = 0; j < c_; ++j) v(j) = (*this)(i, j); return v; }

    Matrix<T, C, R> transpose() const {
        Matrix<T, C, R> t; t.resize(c_, r_);
        for (int i = 0; i < r_; ++i) for (int j = 0; j < c_; ++j) t(j, i) = (*this)(i, j);
        return t;
    }
    T squaredNorm() const {
        std::vector<T> sq(size());
        for (int i = 0; i < size(); ++i) sq[i] = data()[i] * data()[i];
        return size() ? shim_redux(sq.data(), size()) : T(0);
    }
    T norm() const { return std::sqrt(squaredNorm()); }
    T sum() const { return size() ? shim_redux(data(), size()) : T(0); }
    template <int R2, int C2>
    T dot(const Matrix<T, R2, C2>& o) const {
        std::vector<T> pr(size());
        for (int i = 0; i < size(); ++i) pr[i] = data()[i] * o.data()[i];
        return shim_redux(pr.data(), size());
    }
    template <typename U>
    Matrix<U, R, C> cast() const {
        Matrix<U, R, C> m; m.resize(r_, c_);
        for (int i = 0; i < size(); ++i) m.data()[i] = (U)data()[i];
        return m;
    }
    Matrix operator-() const { Matrix m(*this); for (int i = 0; i < size(); ++i) m.data()[i] = -data()[i]; return m; }

    T determinant() const {
        assert(r_ == c_);
        // plain LU with partial pivoting (not on the spherical path)
        std::vector<double> a(size());
        for (int i = 0; i < r_; ++i) for (int j = 0; j < c_; ++j) a[i * c_ + j] = (double)(*this)(i, j);
        double det = 1;
        for (int k = 0; k < r_; ++k) {
            int p = k;
            for (int i = k + 1; i < r_; ++i) if (std::fabs(a[i * c_ + k]) > std::fabs(a[p * c_ + k])) p = i;
            if (a[p * c_ + k] == 0) return T(0);
            if (p != k) { for (int j = 0; j < c_; ++j) std::swap(a[k * c_ + j], a[p * c_ + j]); det = -det; }
            det *= a[k * c_ + k];
            for (int i = k + 1; i < r_; ++i) {
                double f = a[i * c_ + k] / a[k * c_ + k];
                for (int j = k; j < c_; ++j) a[i * c_ + j] -= f * a[k * c_ + j];
            }
        }
        return (T)det;
    }
    T det() const { return determinant(); }          // MRPT's Eigen plugin (RegisterRGBD360.h:511: printed only)
    Matrix inverse() const {
        assert(r_ == c_);
        Matrix inv; inv.resize(r_, c_);
        if (r_ == 6 && sizeof(T) == sizeof(float)) {   // the alignFrames360 solve (RPI.h:4693)
            float in[36], out[36];
            for (int i = 0; i < 36; ++i) in[i] = (float)data()[i];
            r360_inverse6(in, out);
            for (int i = 0; i < 36; ++i) inv.data()[i] = (T)out[i];
            return inv;
        }
        if (r_ == 4 && sizeof(T) == sizeof(float)) {   // sensor extrinsics (RPI.h:4922, Calib360.h:129): same restatement as the product's
            float in[16], out[16];
            for (int i = 0; i < 16; ++i) in[i] = (float)data()[i];
            r360_inverse4(in, out);
            for (int i = 0; i < 16; ++i) inv.data()[i] = (T)out[i];
            return inv;
        }
        const int n = r_;
        std::vector<T> a((size_t)n * 2 * n);
        for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) { a[i * 2 * n + j] = (*this)(i, j); a[i * 2 * n + n + j] = (i == j) ? T(1) : T(0); }
        for (int k = 0; k < n; ++k) {
            int p = k;
            for (int i = k + 1; i < n; ++i) if (std::fabs(a[i * 2 * n + k]) > std::fabs(a[p * 2 * n + k])) p = i;
            if (p != k) for (int j = 0; j < 2 * n; ++j) std::swap(a[k * 2 * n + j], a[p * 2 * n + j]);
            T d = T(1) / a[k * 2 * n + k];
            for (int j = 0; j < 2 * n; ++j) a[k * 2 * n + j] *= d;
            for (int i = 0; i < n; ++i) if (i != k) { T f = a[i * 2 * n + k]; for (int j = 0; j < 2 * n; ++j) a[i * 2 * n + j] -= f * a[k * 2 * n + j]; }
        }
        for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) inv(i, j) = a[i * 2 * n + n + j];
        return inv;
    }
    // MRPT's Eigen plugin: MatrixBase::rank() (RPI.h:4682)
    int rank() const {
        assert(r_ == 6 && c_ == 6);
        float in[36];
        for (int i = 0; i < 36; ++i) in[i] = (float)data()[i];
        return r360_rank6(in);
    }
    Matrix<T, 3, 1> cross(const Matrix<T, 3, 1>& o) const {
        const Matrix& a = *this;
        return Matrix<T, 3, 1>(a(1) * o(2) - a(2) * o(1), a(2) * o(0) - a(0) * o(2), a(0) * o(1) - a(1) * o(0));
    }

    template <int R2, int C2> Matrix& operator+=(const Matrix<T, R2, C2>& o) { for (int i = 0; i < size(); ++i) data()[i] += o.data()[i]; return *this; }
    template <int R2, int C2> Matrix& operator-=(const Matrix<T, R2, C2>& o) { for (int i = 0; i < size(); ++i) data()[i] -= o.data()[i]; return *this; }
    Matrix& operator*=(T k) { for (int i = 0; i < size(); ++i) data()[i] *= k; return *this; }
    Matrix& operator/=(T k) { for (int i = 0; i < size(); ++i) data()[i] /= k; return *this; }
};

// A writable view: a copy of the block for reading (so every Matrix operator works on it)
// plus the parent's address for assignment.
template <typename T>
class BlockRef : public Matrix<T, -1, -1> {
public:
    T* base;
    int ld, nr, nc;
    BlockRef(T* b, int ld_, int nr_, int nc_) : base(b), ld(ld_), nr(nr_), nc(nc_) {
        this->resize(nr, nc);
        for (int c = 0; c < nc; ++c) for (int r = 0; r < nr; ++r) (*this)(r, c) = base[r + (size_t)c * ld];
    }
    template <int R2, int C2>
    BlockRef& operator=(const Matrix<T, R2, C2>& o) {
        assert(o.rows() == nr && o.cols() == nc);
        for (int c = 0; c < nc; ++c) for (int r = 0; r < nr; ++r) { base[r + (size_t)c * ld] = o(r, c); (*this)(r, c) = o(r, c); }
        return *this;
    }
    BlockRef& operator=(const BlockRef& o) { return (*this) = static_cast<const Matrix<T, -1, -1>&>(o); }
};
template <typename T, int R, int C>
BlockRef<T> Matrix<T, R, C>::block(int i, int j, int nr, int nc) {
    return BlockRef<T>(&(*this)(i, j), r_, nr, nc);
}

template <int A, int B> struct ShimPick { enum { v = (A > 0) ? A : B }; };

template <typename T, int R1, int C1, int R2, int C2>
Matrix<T, ShimPick<R1, R2>::v, ShimPick<C1, C2>::v> operator+(const Matrix<T, R1, C1>& a, const Matrix<T, R2, C2>& b) {
    assert(a.rows() == b.rows() && a.cols() == b.cols());
    Matrix<T, ShimPick<R1, R2>::v, ShimPick<C1, C2>::v> m; m.resize(a.rows(), a.cols());
    for (int i = 0; i < a.size(); ++i) m.data()[i] = a.data()[i] + b.data()[i];
    return m;
}
template <typename T, int R1, int C1, int R2, int C2>
Matrix<T, ShimPick<R1, R2>::v, ShimPick<C1, C2>::v> operator-(const Matrix<T, R1, C1>& a, const Matrix<T, R2, C2>& b) {
    assert(a.rows() == b.rows() && a.cols() == b.cols());
    Matrix<T, ShimPick<R1, R2>::v, ShimPick<C1, C2>::v> m; m.resize(a.rows(), a.cols());
    for (int i = 0; i < a.size(); ++i) m.data()[i] = a.data()[i] - b.data()[i];
    return m;
}
// coefficient-based product, k ascending, left to right (Eigen 3.2 small fixed-size products)
template <typename T, int R1, int C1, int R2, int C2>
Matrix<T, R1, C2> operator*(const Matrix<T, R1, C1>& a, const Matrix<T, R2, C2>& b) {
    assert(a.cols() == b.rows());
    Matrix<T, R1, C2> m; m.resize(a.rows(), b.cols());
    const int K = a.cols();
    for (int j = 0; j < b.cols(); ++j)
        for (int i = 0; i < a.rows(); ++i) {
            T acc = a(i, 0) * b(0, j);
            for (int k = 1; k < K; ++k) acc = acc + a(i, k) * b(k, j);
            m(i, j) = acc;
        }
    return m;
}
// scalar * matrix: Eigen takes the scalar as `const Scalar&`, so a double factor is first narrowed
// to the matrix's scalar type (RPI.h:4682 `lambda*getDiagonalMatrix(hessian)`).
template <typename T, int R, int C>
Matrix<T, R, C> operator*(const typename Matrix<T, R, C>::Scalar& k, const Matrix<T, R, C>& a) {
    Matrix<T, R, C> m(a);
    for (int i = 0; i < a.size(); ++i) m.data()[i] = k * a.data()[i];
    return m;
}
template <typename T, int R, int C>
Matrix<T, R, C> operator*(const Matrix<T, R, C>& a, const typename Matrix<T, R, C>::Scalar& k) {
    Matrix<T, R, C> m(a);
    for (int i = 0; i < a.size(); ++i) m.data()[i] = a.data()[i] * k;
    return m;
}
template <typename T, int R, int C>
Matrix<T, R, C> operator/(const Matrix<T, R, C>& a, const typename Matrix<T, R, C>::Scalar& k) {
    Matrix<T, R, C> m(a);
    for (int i = 0; i < a.size(); ++i) m.data()[i] = a.data()[i] / k;
    return m;
}
// 1x1 results used as scalars (e.g. row * column products assigned to a float)
template <typename T, int R, int C>
std::ostream& operator<<(std::ostream& os, const Matrix<T, R, C>& m) {
    for (int i = 0; i < m.rows(); ++i) {
        for (int j = 0; j < m.cols(); ++j) os << (j ? " " : "") << m(i, j);
        if (i + 1 < m.rows()) os << "\n";
    }
    return os;
}

typedef Matrix<float, 2, 2> Matrix2f;
typedef Matrix<float, 3, 3> Matrix3f;
typedef Matrix<float, 4, 4> Matrix4f;
typedef Matrix<double, 3, 3> Matrix3d;
typedef Matrix<double, 4, 4> Matrix4d;
typedef Matrix<float, 2, 1> Vector2f;
typedef Matrix<float, 3, 1> Vector3f;
typedef Matrix<float, 4, 1> Vector4f;
typedef Matrix<double, 2, 1> Vector2d;
typedef Matrix<double, 3, 1> Vector3d;
typedef Matrix<double, 4, 1> Vector4d;
typedef Matrix<float, -1, -1> MatrixXf;
typedef Matrix<double, -1, -1> MatrixXd;
typedef Matrix<float, -1, 1> VectorXf;
typedef Matrix<double, -1, 1> VectorXd;
typedef Matrix<int, -1, 1> VectorXi;

template <typename T>
struct aligned_allocator : public std::allocator<T> {
    template <typename U> struct rebind { typedef aligned_allocator<U> other; };
    aligned_allocator() {}
    template <typename U> aligned_allocator(const aligned_allocator<U>&) {}
};

enum { ComputeFullU = 1, ComputeFullV = 2, ComputeThinU = 4, ComputeThinV = 8 };
template <typename M>
struct JacobiSVD {   // not on the spherical path; present so Miscellaneous.h parses
    JacobiSVD() {}
    JacobiSVD(const M&, unsigned = 0) { shim_abort(); }
    static void shim_abort() { std::cerr << "refshim: JacobiSVD is not implemented\n"; std::abort(); }
    M matrixU() const { shim_abort(); return M(); }
    M matrixV() const { shim_abort(); return M(); }
    Matrix<typename M::Scalar, -1, 1> singularValues() const { shim_abort(); return Matrix<typename M::Scalar, -1, 1>(); }
};

}  // namespace Eigen
