/* EigenShim.h -- the handful of Eigen 3 dense types the reference's container classes expose in their public interface
 * (ParticleSystem.h:22-41, RegularGrid.h:36-60, LagrangianMesh.h:38-74, HybridSolver.h:18-19), for hosts that do not have
 * Eigen installed (this image has none: SURVEY.md 8c).  When <Eigen/Core> is available it is used instead and this file
 * defines nothing, so a maintainer of the reference keeps the real library.
 *
 * Only storage + the few operations the host layer and its callers need; memory layouts are Eigen's defaults:
 *   MatrixX3d / MatrixX3i   column-major N x 3 (three contiguous planes, leading dimension N)
 *   Matrix3d / Matrix2d     column-major
 */
#ifndef AEP_EIGEN_SHIM_H
#define AEP_EIGEN_SHIM_H

#if defined(AEP_USE_REAL_EIGEN) || (defined(__has_include) && __has_include(<Eigen/Core>) && !defined(AEP_FORCE_EIGEN_SHIM))
#include <Eigen/Core>
#else

#include <cmath>
#include <cstddef>
#include <vector>

namespace Eigen {

template <typename T>
struct Vec3 {
    T v[3];
    Vec3() : v{T(0), T(0), T(0)} {}
    Vec3(T a, T b, T c) : v{a, b, c} {}
    T& operator[](int i) { return v[i]; }
    const T& operator[](int i) const { return v[i]; }
    T& operator()(int i) { return v[i]; }
    const T& operator()(int i) const { return v[i]; }
    T& x() { return v[0]; } T& y() { return v[1]; } T& z() { return v[2]; }
    const T& x() const { return v[0]; } const T& y() const { return v[1]; } const T& z() const { return v[2]; }
    int size() const { return 3; }
    T* data() { return v; }
    const T* data() const { return v; }
    Vec3 operator+(const Vec3& o) const { return Vec3(v[0] + o.v[0], v[1] + o.v[1], v[2] + o.v[2]); }
    Vec3 operator-(const Vec3& o) const { return Vec3(v[0] - o.v[0], v[1] - o.v[1], v[2] - o.v[2]); }
    Vec3 operator-() const { return Vec3(-v[0], -v[1], -v[2]); }
    Vec3 operator*(T s) const { return Vec3(v[0] * s, v[1] * s, v[2] * s); }
    Vec3 operator/(T s) const { return Vec3(v[0] / s, v[1] / s, v[2] / s); }
    Vec3& operator+=(const Vec3& o) { v[0] += o.v[0]; v[1] += o.v[1]; v[2] += o.v[2]; return *this; }
    T dot(const Vec3& o) const { return v[0] * o.v[0] + v[1] * o.v[1] + v[2] * o.v[2]; }
    Vec3 cross(const Vec3& o) const { return Vec3(v[1] * o.v[2] - v[2] * o.v[1], v[2] * o.v[0] - v[0] * o.v[2], v[0] * o.v[1] - v[1] * o.v[0]); }
    T squaredNorm() const { return dot(*this); }
    double norm() const { return std::sqrt((double)squaredNorm()); }
    Vec3 normalized() const { const double n = norm(); return n > 0 ? (*this) / (T)n : *this; }
    T prod() const { return v[0] * v[1] * v[2]; }
    T minCoeff() const { return v[0] < v[1] ? (v[0] < v[2] ? v[0] : v[2]) : (v[1] < v[2] ? v[1] : v[2]); }
    T maxCoeff() const { return v[0] > v[1] ? (v[0] > v[2] ? v[0] : v[2]) : (v[1] > v[2] ? v[1] : v[2]); }
    static Vec3 Zero() { return Vec3(); }
};
template <typename T> inline Vec3<T> operator*(T s, const Vec3<T>& a) { return a * s; }
using Vector3d = Vec3<double>;
using Vector3i = Vec3<int>;
struct Vector4f {                                           /* viewer.core.background_color in main.cpp:97 */
    float v[4];
    Vector4f() : v{0.f, 0.f, 0.f, 0.f} {}
    Vector4f(float a, float b, float c, float d) : v{a, b, c, d} {}
    float& operator[](int i) { return v[i]; }
    const float& operator[](int i) const { return v[i]; }
};

struct VectorXd {
    std::vector<double> s;
    VectorXd() {}
    explicit VectorXd(std::ptrdiff_t n) : s((size_t)n) {}
    VectorXd(const Vector3d& a) : s{a[0], a[1], a[2]} {}      // Vector3d -> VectorXd, as main.cpp:67-69 relies on
    void resize(std::ptrdiff_t n) { s.resize((size_t)n); }
    std::ptrdiff_t size() const { return (std::ptrdiff_t)s.size(); }
    std::ptrdiff_t rows() const { return size(); }
    double& operator[](std::ptrdiff_t i) { return s[(size_t)i]; }
    const double& operator[](std::ptrdiff_t i) const { return s[(size_t)i]; }
    double& operator()(std::ptrdiff_t i) { return s[(size_t)i]; }
    const double& operator()(std::ptrdiff_t i) const { return s[(size_t)i]; }
    double* data() { return s.data(); }
    const double* data() const { return s.data(); }
    VectorXd& setZero() { for (double& x : s) x = 0.0; return *this; }
    VectorXd& setOnes() { for (double& x : s) x = 1.0; return *this; }
    VectorXd& setConstant(double c) { for (double& x : s) x = c; return *this; }
    VectorXd& operator*=(double c) { for (double& x : s) x *= c; return *this; }
    double sum() const { double t = 0; for (double x : s) t += x; return t; }
};
inline VectorXd operator*(double c, const VectorXd& a) { VectorXd r = a; r *= c; return r; }

template <typename T>
struct MatX3 {
    std::vector<T> s;               // column-major, leading dimension n
    std::ptrdiff_t n = 0;
    MatX3() {}
    MatX3(std::ptrdiff_t rows, std::ptrdiff_t cols) { resize(rows, cols); }
    void resize(std::ptrdiff_t rows, std::ptrdiff_t /*cols == 3*/) { n = rows; s.resize((size_t)(3 * rows)); }
    std::ptrdiff_t rows() const { return n; }
    std::ptrdiff_t cols() const { return 3; }
    std::ptrdiff_t size() const { return 3 * n; }
    T& operator()(std::ptrdiff_t r, std::ptrdiff_t c) { return s[(size_t)(c * n + r)]; }
    const T& operator()(std::ptrdiff_t r, std::ptrdiff_t c) const { return s[(size_t)(c * n + r)]; }
    T* data() { return s.data(); }
    const T* data() const { return s.data(); }
    MatX3& setZero() { for (T& x : s) x = T(0); return *this; }
    MatX3& setOnes() { for (T& x : s) x = T(1); return *this; }
    Vec3<T> row(std::ptrdiff_t r) const { return Vec3<T>((*this)(r, 0), (*this)(r, 1), (*this)(r, 2)); }
    void setRow(std::ptrdiff_t r, const Vec3<T>& a) { (*this)(r, 0) = a[0]; (*this)(r, 1) = a[1]; (*this)(r, 2) = a[2]; }
};
using MatrixX3d = MatX3<double>;
using MatrixX3i = MatX3<int>;

template <int N>
struct MatN {
    double a[N * N];
    MatN() { for (double& x : a) x = 0.0; }
    double& operator()(int r, int c) { return a[c * N + r]; }
    const double& operator()(int r, int c) const { return a[c * N + r]; }
    double* data() { return a; }
    const double* data() const { return a; }
    static MatN Identity() { MatN m; for (int i = 0; i < N; ++i) m(i, i) = 1.0; return m; }
    static MatN Zero() { return MatN(); }
};
using Matrix3d = MatN<3>;
using Matrix2d = MatN<2>;

}  // namespace Eigen
#endif /* real Eigen */

namespace aep_host {
/* row access that works with both the shim and real Eigen */
template <typename M, typename V> inline void set_row(M& m, std::ptrdiff_t r, const V& a) { m(r, 0) = a[0]; m(r, 1) = a[1]; m(r, 2) = a[2]; }
template <typename M> inline Eigen::Vector3d get_row(const M& m, std::ptrdiff_t r) { return Eigen::Vector3d(m(r, 0), m(r, 1), m(r, 2)); }
}  // namespace aep_host

#endif /* AEP_EIGEN_SHIM_H */
