// ============================================================================
// oracle/mpm_oracle.cpp  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// fp64 CPU restatement of the MPM substep of the reference
// (2iw31Zhv/AnisotropicElastoplasticity, AnisotropicElastoplasticity/*.cpp).
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference legs may load this library.  The shipped engine (libaep_b200.so)
// never links, loads or calls anything in oracle/.
//
// PINNED TO THE REFERENCE'S OWN CODE: the reference's unmodified sources are compiled where they lie under
// /root/reference against a from-scratch Eigen/igl stand-in (oracle/ref_shim, oracle/ref_driver.cpp ->
// oracle/_ref/libaep_ref.so; the image has no Eigen, libigl, GLFW or GLEW, and the reference ships no tests or
// golden vectors of its own).  tests/test_reference_pin.py holds this restatement to that library -- live where it
// is built, and everywhere through the fixtures it wrote (tests/golden/ref_*.npz, oracle/make_ref_golden.py) -- at
// fp64 rounding, including one whole frame run by HybridSolver::solve itself.  A second, independent pin is the
// numpy/scipy transcription of the reference's sparse-matrix algebra (oracle/literal_numpy.py -> tests/golden/*.npz).
//
// Every function cites the reference file:line it restates.  "HS" = HybridSolver.cpp,
// "LM" = LagrangianMesh.cpp, "RG" = RegularGrid.cpp, "IP" = interpolation.cpp,
// "GE" = geometry.cpp, "LS" = LevelSet.cpp.
//
// Layout conventions of the C API (identical to include/aep_b200.h):
//   N x 3 matrices  : column-major, leading dimension N (Eigen MatrixX3d)
//   N x 3x3 tensors : 9 doubles per item, each 3x3 column-major (std::vector<Matrix3d>)
//   grid node index : k*nx*ny + j*nx + i                          (RG:164-168)
// ============================================================================
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

// ---------------------------------------------------------------- small 3x3 helpers
struct M3 {           // column-major like Eigen: a[3*c + r]
    double a[9];
    double& operator()(int r, int c) { return a[3 * c + r]; }
    double operator()(int r, int c) const { return a[3 * c + r]; }
};
inline M3 m3_zero() { M3 m; for (double& x : m.a) x = 0.0; return m; }
inline M3 m3_ident() { M3 m = m3_zero(); m(0,0) = m(1,1) = m(2,2) = 1.0; return m; }
inline M3 mul(const M3& A, const M3& B) {
    M3 C = m3_zero();
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) {
        double s = 0.0; for (int k = 0; k < 3; ++k) s += A(r,k) * B(k,c); C(r,c) = s; }
    return C;
}
inline M3 transpose(const M3& A) { M3 T; for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) T(r,c) = A(c,r); return T; }
inline M3 add(const M3& A, const M3& B) { M3 C; for (int i = 0; i < 9; ++i) C.a[i] = A.a[i] + B.a[i]; return C; }
inline M3 sub(const M3& A, const M3& B) { M3 C; for (int i = 0; i < 9; ++i) C.a[i] = A.a[i] - B.a[i]; return C; }
inline M3 scale(const M3& A, double s) { M3 C; for (int i = 0; i < 9; ++i) C.a[i] = A.a[i] * s; return C; }
inline double det(const M3& A) {
    return A(0,0) * (A(1,1) * A(2,2) - A(1,2) * A(2,1))
         - A(0,1) * (A(1,0) * A(2,2) - A(1,2) * A(2,0))
         + A(0,2) * (A(1,0) * A(2,1) - A(1,1) * A(2,0));
}
inline M3 inverse(const M3& A) {
    double d = det(A); M3 I;
    I(0,0) =  (A(1,1) * A(2,2) - A(1,2) * A(2,1)) / d;
    I(0,1) = -(A(0,1) * A(2,2) - A(0,2) * A(2,1)) / d;
    I(0,2) =  (A(0,1) * A(1,2) - A(0,2) * A(1,1)) / d;
    I(1,0) = -(A(1,0) * A(2,2) - A(1,2) * A(2,0)) / d;
    I(1,1) =  (A(0,0) * A(2,2) - A(0,2) * A(2,0)) / d;
    I(1,2) = -(A(0,0) * A(1,2) - A(0,2) * A(1,0)) / d;
    I(2,0) =  (A(1,0) * A(2,1) - A(1,1) * A(2,0)) / d;
    I(2,1) = -(A(0,0) * A(2,1) - A(0,1) * A(2,0)) / d;
    I(2,2) =  (A(0,0) * A(1,1) - A(0,1) * A(1,0)) / d;
    return I;
}
inline M3 diag3(double a, double b, double c) { M3 m = m3_zero(); m(0,0) = a; m(1,1) = b; m(2,2) = c; return m; }

// ---------------------------------------------------------------- SVD (Eigen::JacobiSVD contract)
// The reference calls Eigen::JacobiSVD<Matrix3d>(F, ComputeFullU|ComputeFullV) (HS:308, HS:620).
// Eigen is not vendored; its published contract is: F = U diag(s) V^T, U,V orthogonal,
// s >= 0 sorted descending.  Every use in the reference (U f(S) V^T, V S^-1 U^T, U V^T) is
// invariant to the remaining freedom, so any SVD meeting that contract reproduces the
// reference to rounding.  Implemented as one-sided (Hestenes) Jacobi run to convergence.
void svd3(const M3& F, M3& U, double s[3], M3& V) {
    M3 A = F; V = m3_ident();
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0.0;
        for (int p = 0; p < 2; ++p) for (int q = p + 1; q < 3; ++q) {
            double al = 0, be = 0, ga = 0;
            for (int r = 0; r < 3; ++r) { al += A(r,p) * A(r,p); be += A(r,q) * A(r,q); ga += A(r,p) * A(r,q); }
            if (ga == 0.0) continue;
            double lim = 1e-300 + 1e-32 * al * be;
            if (ga * ga <= lim) continue;
            off = std::max(off, std::fabs(ga) / std::sqrt(al * be));
            double zeta = (be - al) / (2.0 * ga);
            double t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
            double c = 1.0 / std::sqrt(1.0 + t * t), sn = c * t;
            for (int r = 0; r < 3; ++r) {
                double ap = A(r,p), aq = A(r,q);
                A(r,p) = c * ap - sn * aq; A(r,q) = sn * ap + c * aq;
                double vp = V(r,p), vq = V(r,q);
                V(r,p) = c * vp - sn * vq; V(r,q) = sn * vp + c * vq;
            }
        }
        if (off < 1e-15) break;
    }
    double n[3]; int order[3] = {0, 1, 2};
    for (int c = 0; c < 3; ++c) n[c] = std::sqrt(A(0,c) * A(0,c) + A(1,c) * A(1,c) + A(2,c) * A(2,c));
    std::sort(order, order + 3, [&](int a, int b) { return n[a] > n[b]; });
    M3 Vs, Us = m3_zero();
    for (int c = 0; c < 3; ++c) {
        int o = order[c]; s[c] = n[o];
        for (int r = 0; r < 3; ++r) { Vs(r,c) = V(r,o); Us(r,c) = (n[o] > 0.0) ? A(r,o) / n[o] : 0.0; }
    }
    // complete U for exactly-zero singular values (never hit by physical states; keeps U orthogonal)
    for (int c = 0; c < 3; ++c) if (s[c] == 0.0) {
        int c1 = (c + 1) % 3, c2 = (c + 2) % 3;
        double x[3] = { Us(1,c1) * Us(2,c2) - Us(2,c1) * Us(1,c2),
                        Us(2,c1) * Us(0,c2) - Us(0,c1) * Us(2,c2),
                        Us(0,c1) * Us(1,c2) - Us(1,c1) * Us(0,c2) };
        double nn = std::sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
        if (nn > 0) for (int r = 0; r < 3; ++r) Us(r,c) = x[r] / nn;
    }
    U = Us; V = Vs;
}

// 2x2 SVD, same contract, for the cloth polar decomposition (LM:437-440).
void svd2(const double A[4] /*col-major*/, double U[4], double s[2], double V[4]) {
    double a00 = A[0], a10 = A[1], a01 = A[2], a11 = A[3];
    double v00 = 1, v10 = 0, v01 = 0, v11 = 1;
    for (int it = 0; it < 60; ++it) {
        double al = a00 * a00 + a10 * a10, be = a01 * a01 + a11 * a11, ga = a00 * a01 + a10 * a11;
        if (ga == 0.0 || ga * ga <= 1e-300 + 1e-32 * al * be) break;
        double zeta = (be - al) / (2.0 * ga);
        double t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
        double c = 1.0 / std::sqrt(1.0 + t * t), sn = c * t;
        double p, q;
        p = a00; q = a01; a00 = c * p - sn * q; a01 = sn * p + c * q;
        p = a10; q = a11; a10 = c * p - sn * q; a11 = sn * p + c * q;
        p = v00; q = v01; v00 = c * p - sn * q; v01 = sn * p + c * q;
        p = v10; q = v11; v10 = c * p - sn * q; v11 = sn * p + c * q;
    }
    double n0 = std::sqrt(a00 * a00 + a10 * a10), n1 = std::sqrt(a01 * a01 + a11 * a11);
    double u00 = n0 > 0 ? a00 / n0 : 1, u10 = n0 > 0 ? a10 / n0 : 0;
    double u01 = n1 > 0 ? a01 / n1 : -u10, u11 = n1 > 0 ? a11 / n1 : u00;
    if (n0 >= n1) { s[0] = n0; s[1] = n1; U[0] = u00; U[1] = u10; U[2] = u01; U[3] = u11; V[0] = v00; V[1] = v10; V[2] = v01; V[3] = v11; }
    else          { s[0] = n1; s[1] = n0; U[0] = u01; U[1] = u11; U[2] = u00; U[3] = u10; V[0] = v01; V[1] = v11; V[2] = v00; V[3] = v10; }
}

// ---------------------------------------------------------------- IP:9-49
inline double cubic_bspline(double x) {                 // IP:9-16
    double ax = std::fabs(x);
    return (ax >= 2.0) ? 0.0
         : (ax >= 1.0 ? -1.0 / 6.0 * ax * ax * ax + ax * ax - 2.0 * ax + 4.0 / 3.0
                      : 0.5 * ax * ax * ax - ax * ax + 2.0 / 3.0);
}
inline double dcubic_bspline(double x) {                // IP:18-33
    return (x >= 2.0) ? 0.0
         : (x >= 1.0 ? -0.5 * x * x + 2.0 * x - 2.0
         : (x >= 0.0 ? 1.5 * x * x - 2.0 * x
         : (x >= -1.0 ? -1.5 * x * x - 2.0 * x
         : (x >= -2.0 ? 0.5 * x * x + 2.0 * x + 2.0 : 0.0))));
}
inline double clampd(double x, double lo, double hi) { return x > hi ? hi : (x < lo ? lo : x); }   // IP:35-49

// ---------------------------------------------------------------- GE:31-62, GE:67-73
void gram_schmidt(M3& Q, M3& R, const M3& A) {
    double d1[3] = {A(0,0), A(1,0), A(2,0)}, d2[3] = {A(0,1), A(1,1), A(2,1)}, d3[3] = {A(0,2), A(1,2), A(2,2)};
    auto nrm = [](const double* v) { return std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); };
    auto dot = [](const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; };
    double r11 = nrm(d1); double q1[3] = {d1[0] / r11, d1[1] / r11, d1[2] / r11};
    double r12 = dot(d2, q1);
    double q2[3] = {d2[0] - r12 * q1[0], d2[1] - r12 * q1[1], d2[2] - r12 * q1[2]};
    double r22 = nrm(q2); for (double& x : q2) x /= r22;
    double r13 = dot(d3, q1), r23 = dot(d3, q2);
    double q3[3] = {d3[0] - r13 * q1[0] - r23 * q2[0], d3[1] - r13 * q1[1] - r23 * q2[1], d3[2] - r13 * q1[2] - r23 * q2[2]};
    double r33 = nrm(q3); for (double& x : q3) x /= r33;
    for (int r = 0; r < 3; ++r) { Q(r,0) = q1[r]; Q(r,1) = q2[r]; Q(r,2) = q3[r]; }
    R = m3_zero(); R(0,0) = r11; R(0,1) = r12; R(0,2) = r13; R(1,1) = r22; R(1,2) = r23; R(2,2) = r33;
}

// ---------------------------------------------------------------- level sets
enum { LS_NONE = 0, LS_GROUND = 1, LS_WALL2GROUND = 2, LS_SPHERE_GROUND = 3, LS_BOX = 4, LS_SAMPLED = 5 };

// phi and grad phi at a point.  kinds 1,2 restate LS:8-42; 3,4 are new primitives required by
// BASELINE.json configs C2/C3/C5 (the reference has no sphere or box collider).
double ls_phi(int kind, const double* P, const double x[3]) {
    switch (kind) {
    case LS_GROUND: return x[2] - P[0];                                               // LS:8-11
    case LS_WALL2GROUND: return std::min(std::min(x[2] - P[2], P[0] - x[0]), P[1] - x[1]);   // LS:18-21
    case LS_SPHERE_GROUND: {
        double dx = x[0] - P[0], dy = x[1] - P[1], dz = x[2] - P[2];
        return std::min(std::sqrt(dx * dx + dy * dy + dz * dz) - P[3], x[2] - P[4]);
    }
    case LS_BOX: {
        double d = x[0] - P[0];
        d = std::min(d, P[3] - x[0]); d = std::min(d, x[1] - P[1]); d = std::min(d, P[4] - x[1]);
        d = std::min(d, x[2] - P[2]); d = std::min(d, P[5] - x[2]);
        return d;
    }
    default: return 1.0;
    }
}
void ls_normal(int kind, const double* P, const double x[3], double n[3]) {
    n[0] = n[1] = 0.0; n[2] = 1.0;
    switch (kind) {
    case LS_GROUND: return;                                                            // LS:13-16
    case LS_WALL2GROUND: {                                                             // LS:23-42
        double dz = std::fabs(x[2] - P[2]), dx = std::fabs(P[0] - x[0]), dy = std::fabs(P[1] - x[1]);
        if (dz <= dx && dz <= dy) { n[0] = 0; n[1] = 0; n[2] = 1; }
        else if (dy <= dx) { n[0] = 0; n[1] = -1; n[2] = 0; }
        else { n[0] = -1; n[1] = 0; n[2] = 0; }
        return;
    }
    case LS_SPHERE_GROUND: {
        double dx = x[0] - P[0], dy = x[1] - P[1], dz = x[2] - P[2];
        double r = std::sqrt(dx * dx + dy * dy + dz * dz);
        if (r - P[3] <= x[2] - P[4] && r > 0.0) { n[0] = dx / r; n[1] = dy / r; n[2] = dz / r; }
        return;
    }
    case LS_BOX: {
        double d[6] = { x[2] - P[2], P[5] - x[2], x[0] - P[0], P[3] - x[0], x[1] - P[1], P[4] - x[1] };
        static const double N[6][3] = { {0,0,1}, {0,0,-1}, {1,0,0}, {-1,0,0}, {0,1,0}, {0,-1,0} };
        int best = 0; for (int f = 1; f < 6; ++f) if (d[f] < d[best]) best = f;
        n[0] = N[best][0]; n[1] = N[best][1]; n[2] = N[best][2];
        return;
    }
    default: return;
    }
}

// ---------------------------------------------------------------- stencils  (HS:18-97)
// One "weight matrix row" per point: <= 64 (node, w, dw/dx, dw/dy, dw/dz) entries, i.e. the
// rows of omegas_/domegas_{1,2,3}_ without materialising Np x Ng sparse matrices.
struct Stencils {
    long n = 0;
    std::vector<int> cnt;          // entries per point
    std::vector<int> idx;          // 64 per point
    std::vector<double> w, d1, d2, d3;
    void resize(long np) { n = np; cnt.assign(np, 0); idx.resize(64 * np); w.resize(64 * np); d1.resize(64 * np); d2.resize(64 * np); d3.resize(64 * np); }
};

struct Sim {
    // RegularGrid (RG:117-177)
    double mn[3], mx[3], h[3]; int res[3]; long Ng = 0;
    std::vector<double> gm, gv, gf, gvt;           // masses, velocities, forces, velocities-before-friction (Ng x 3 col-major)
    // ParticleSystem (ParticleSystem.h:22-41)
    long Np = 0;
    std::vector<double> x, v, B1, B2, B3, FE, FP, cand, m, vol, dens, q;
    double E = 0, nu = 0, thetaC = 0, thetaS = 0;
    Stencils sp;
    // LagrangianMesh (LagrangianMesh.h:38-74)
    long Nv = 0, Nf = 0;
    std::vector<double> vx, vv, vm, vvol, vB1, vB2, vB3;
    std::vector<double> ex, ev, em, evol, eB1, eB2, eB3;
    std::vector<double> ed1, ed2, ed3, eD1, eD2, eD3;
    std::vector<int> faces; std::vector<double> fixedv;
    double mesh_mu = 0, mesh_lambda = 0, mesh_gamma = 0, mesh_k = 0, mesh_cf = 0;
    Stencils sv, se;
    // level set
    int ls_kind = LS_NONE; double ls_par[8] = {0};
    std::vector<uint8_t> ls_inside; std::vector<double> ls_nrm;
    // solver state (HS:827-895)
    int material = 1; double cfl = 0.3, dt = 0.0, t = 0.0, inner_t = 0.0; int frame_flag = 0, frame_no = 0;
    double gravity = 9.8, friction = 0.2, snow_xi = 10.0, sand_h[4] = {35.0, 9.0, 0.2, 10.0}, rate_floor = 3e2, frame_dt = 1.0 / 60.0;
    int nthreads = 1;
    // timers (seconds accumulated per stage)
    double tm[8] = {0};

    int to_index(int i, int j, int k) const { return k * res[0] * res[1] + j * res[0] + i; }       // RG:164-168
    double hmin() const { return std::min(h[0], std::min(h[1], h[2])); }
};

// HS:18-97  evaluateInterpolationWeights_
void build_stencils(const Sim& S, Stencils& st, const std::vector<double>& pos, long np) {
    st.resize(np);
#pragma omp parallel for schedule(static) num_threads(S.nthreads)
    for (long p = 0; p < np; ++p) {
        double pr[3] = { pos[p] - S.mn[0], pos[np + p] - S.mn[1], pos[2 * np + p] - S.mn[2] };
        int fl[3] = { static_cast<int>(pr[0] / S.h[0]), static_cast<int>(pr[1] / S.h[1]), static_cast<int>(pr[2] / S.h[2]) };   // HS:34-36
        double N[3][5], D[3][5];
        for (int a = 0; a < 3; ++a) for (int o = 0; o < 5; ++o) {
            int i = fl[a] - 2 + o;
            N[a][o] = cubic_bspline(pr[a] / S.h[a] - i);                   // HS:48-50
            D[a][o] = dcubic_bspline(pr[a] / S.h[a] - i) / S.h[a];        // HS:52-57
        }
        int c = 0; long base = 64 * p;
        for (int oi = 0; oi < 5; ++oi) { int i = fl[0] - 2 + oi;          // HS:38-42 loop order i, j, k
            for (int oj = 0; oj < 5; ++oj) { int j = fl[1] - 2 + oj;
                for (int ok = 0; ok < 5; ++ok) { int k = fl[2] - 2 + ok;
                    if (0 <= i && i < S.res[0] && 0 <= j && j < S.res[1] && 0 <= k && k < S.res[2]) {       // HS:44-46
                        double wi = N[0][oi], wj = N[1][oj], wk = N[2][ok];
                        if (wi > 0 && wj > 0 && wk > 0) {                                               // HS:60
                            st.idx[base + c] = S.to_index(i, j, k);
                            st.w[base + c]  = wi * wj * wk;
                            st.d1[base + c] = D[0][oi] * wj * wk;
                            st.d2[base + c] = wi * D[1][oj] * wk;
                            st.d3[base + c] = wi * wj * D[2][ok];
                            ++c;
                        }
                    }
                }
            }
        }
        st.cnt[p] = c;
    }
}

inline void atomic_add(double& dst, double v, bool par) {
    if (par) {
#pragma omp atomic update
        dst += v;
    } else dst += v;
}

// HS:144-204 ADD_AFFINE_MOMENTA in its un-optimised form (HS:154-173): p_i += w m (3/h^2) B (x_i - x_p)
// plus HS:118-141 mass and linear momentum, for one point set.
void scatter_mass_momentum(Sim& S, const Stencils& st, long np, const std::vector<double>& mass, const std::vector<double>& vel,
                           const std::vector<double>& b1, const std::vector<double>& b2, const std::vector<double>& b3,
                           const std::vector<double>& pos, std::vector<double>& mom) {
    const double hm = S.hmin(), ratio = 3.0 / hm / hm;                    // HS:175-177
    const long Ng = S.Ng; const bool par = S.nthreads > 1;
#pragma omp parallel for schedule(static) num_threads(S.nthreads)
    for (long p = 0; p < np; ++p) {
        double mp = mass[p];
        double vp[3] = { vel[p], vel[np + p], vel[2 * np + p] };
        double xp[3] = { pos[p], pos[np + p], pos[2 * np + p] };
        double B[3][3] = { { b1[p], b1[np + p], b1[2 * np + p] }, { b2[p], b2[np + p], b2[2 * np + p] }, { b3[p], b3[np + p], b3[2 * np + p] } };
        for (int e = 0; e < st.cnt[p]; ++e) {
            long s = 64 * p + e; int id = st.idx[s]; double w = st.w[s];
            int k = id / (S.res[0] * S.res[1]); int j = (id % (S.res[0] * S.res[1])) / S.res[0]; int i = id - k * S.res[0] * S.res[1] - j * S.res[0];
            double xi[3] = { S.mn[0] + i * S.h[0], S.mn[1] + j * S.h[1], S.mn[2] + k * S.h[2] };    // RG:152-156
            double dx[3] = { xi[0] - xp[0], xi[1] - xp[1], xi[2] - xp[2] };
            atomic_add(S.gm[id], w * mp, par);
            for (int a = 0; a < 3; ++a) {
                double aff = B[a][0] * dx[0] + B[a][1] * dx[1] + B[a][2] * dx[2];
                atomic_add(mom[a * Ng + id], w * mp * (vp[a] + ratio * aff), par);
            }
        }
    }
}

// HS:113-250 particleToGrid_
void particle_to_grid(Sim& S, bool first) {
    std::fill(S.gm.begin(), S.gm.end(), 0.0);
    std::vector<double> mom(3 * S.Ng, 0.0);
    if (S.Np) scatter_mass_momentum(S, S.sp, S.Np, S.m, S.v, S.B1, S.B2, S.B3, S.x, mom);
    if (S.Nv) {
        scatter_mass_momentum(S, S.sv, S.Nv, S.vm, S.vv, S.vB1, S.vB2, S.vB3, S.vx, mom);
        scatter_mass_momentum(S, S.se, S.Nf, S.em, S.ev, S.eB1, S.eB2, S.eB3, S.ex, mom);
    }
    const long Ng = S.Ng;
#pragma omp parallel for schedule(static) num_threads(S.nthreads)
    for (long c = 0; c < Ng; ++c) {                                        // HS:233-240
        double mc = S.gm[c];
        for (int a = 0; a < 3; ++a) S.gv[a * Ng + c] = (mc > 0.0) ? mom[a * Ng + c] / mc : 0.0;
    }
    if (first && S.Np) {                                                   // HS:242-249
        double gvol = S.h[0] * S.h[1] * S.h[2];
#pragma omp parallel for schedule(static) num_threads(S.nthreads)
        for (long p = 0; p < S.Np; ++p) {
            double d = 0.0;
            for (int e = 0; e < S.sp.cnt[p]; ++e) d += S.sp.w[64 * p + e] * S.gm[S.sp.idx[64 * p + e]];
            S.dens[p] = d / gvol; S.vol[p] = S.m[p] * (1.0 / S.dens[p]);
        }
    }
}

inline M3 load_m3(const std::vector<double>& a, long p) { M3 m; std::memcpy(m.a, &a[9 * p], 72); return m; }
inline void store_m3(std::vector<double>& a, long p, const M3& m) { std::memcpy(&a[9 * p], m.a, 72); }

// M = sum_i vel_i (grad w_i)^T  (un-scaled); column c of the result is row p of domegas_c * vel (HS:269-271,297-301)
inline M3 grad_field(const Sim& S, const Stencils& st, long p, const std::vector<double>& vel) {
    M3 G = m3_zero(); const long Ng = S.Ng;
    for (int e = 0; e < st.cnt[p]; ++e) {
        long s = 64 * p + e; int id = st.idx[s];
        double vi[3] = { vel[id], vel[Ng + id], vel[2 * Ng + id] };
        for (int r = 0; r < 3; ++r) { G(r,0) += st.d1[s] * vi[r]; G(r,1) += st.d2[s] * vi[r]; G(r,2) += st.d3[s] * vi[r]; }
    }
    return G;
}
inline void scatter_stress(Sim& S, const Stencils& st, long p, const M3& A, bool par) {      // HS:356-366 / HS:444-454
    const long Ng = S.Ng;
    for (int e = 0; e < st.cnt[p]; ++e) {
        long s = 64 * p + e; int id = st.idx[s];
        for (int r = 0; r < 3; ++r)
            atomic_add(S.gf[r * Ng + id], -(A(r,0) * st.d1[s] + A(r,1) * st.d2[s] + A(r,2) * st.d3[s]), par);
    }
}

// LM:382-460 computeVertexInPlaneForces
void cloth_in_plane(const Sim& S, std::vector<double>& vf, std::vector<double>& pk /*4 per face col-major*/) {
    vf.assign(3 * S.Nv, 0.0); pk.assign(4 * S.Nf, 0.0);
    const long Nf = S.Nf, Nv = S.Nv;
    for (long f = 0; f < Nf; ++f) {
        M3 Dm, dm;
        for (int r = 0; r < 3; ++r) {
            Dm(r,0) = S.eD1[r * Nf + f]; Dm(r,1) = S.eD2[r * Nf + f]; Dm(r,2) = S.eD3[r * Nf + f];
            dm(r,0) = S.ed1[r * Nf + f]; dm(r,1) = S.ed2[r * Nf + f]; dm(r,2) = S.ed3[r * Nf + f];
        }
        M3 Q, R, Q0, R0; gram_schmidt(Q, R, dm); gram_schmidt(Q0, R0, Dm);
        double i11 = 1.0 / R0(0,0), i12 = -R0(0,1) / R0(0,0) / R0(1,1), i22 = 1.0 / R0(1,1);       // GE:67-73
        // refInPlaneR = invRest * inPlaneR (LM:431), 2x2 upper triangular
        double r00 = i11 * R(0,0), r01 = i11 * R(0,1) + i12 * R(1,1), r11 = i22 * R(1,1);
        double A2[4] = { r00, 0.0, r01, r11 }, U[4], sg[2], V[4];
        svd2(A2, U, sg, V);
        // rotation = U V^T (LM:438)
        double rot00 = U[0] * V[0] + U[2] * V[2], rot01 = U[0] * V[1] + U[2] * V[3];
        double rot10 = U[1] * V[0] + U[3] * V[2], rot11 = U[1] * V[1] + U[3] * V[3];
        double J = r00 * r11;                                                                       // LM:441
        // invRefMulDet = [r11 -r01; 0 r00]; stress uses its transpose (LM:433-444)
        double P00 = 2.0 * S.mesh_mu * (r00 - rot00) + S.mesh_lambda * (J - 1.0) * r11;
        double P01 = 2.0 * S.mesh_mu * (r01 - rot01) + S.mesh_lambda * (J - 1.0) * 0.0;
        double P10 = 2.0 * S.mesh_mu * (0.0 - rot10) + S.mesh_lambda * (J - 1.0) * (-r01);
        double P11 = 2.0 * S.mesh_mu * (r11 - rot11) + S.mesh_lambda * (J - 1.0) * r00;
        // inPlanePiolaKirhoffStresses[f] = invRest * P (LM:446)
        pk[4 * f + 0] = i11 * P00 + i12 * P10; pk[4 * f + 2] = i11 * P01 + i12 * P11;
        pk[4 * f + 1] = i22 * P10;             pk[4 * f + 3] = i22 * P11;
        double f2[3], f3[3];
        for (int r = 0; r < 3; ++r) {
            f2[r] = -(P00 * i11 + P01 * i12) * Q(r,0);                                              // LM:452
            f3[r] = -P01 * i22 * Q(r,0) - P11 * i22 * Q(r,1);                                       // LM:453
        }
        int a = S.faces[f], b = S.faces[Nf + f], c = S.faces[2 * Nf + f];
        for (int r = 0; r < 3; ++r) {
            vf[r * Nv + a] += -(f2[r] + f3[r]); vf[r * Nv + b] += f2[r]; vf[r * Nv + c] += f3[r];   // LM:454-458
        }
    }
}

// HS:277-352 body of the particle loop: stress = V_p * P(Fhat) * FE^T
M3 particle_stress(const Sim& S, const M3& Fh, const M3& FE, const M3& FP, double vol, double lambda0, double mu0) {
    double lambda = lambda0, mu = mu0;
    if (S.material == 0) {                                         // SNOW HS:281-287
        double Jp = det(FP);
        lambda = lambda0 * std::exp(S.snow_xi * (1 - Jp)); mu = mu0 * std::exp(S.snow_xi * (1 - Jp));
    }
    M3 U, V; double sg[3]; svd3(Fh, U, sg, V);                     // HS:308
    if (S.material == 0) {                                         // HS:314-325
        M3 Rm = mul(U, transpose(V));
        double J = det(Fh);
        M3 P = add(scale(sub(Fh, Rm), 2.0 * mu), scale(inverse(transpose(Fh)), lambda * (J - 1.0) * J));
        return scale(mul(P, transpose(FE)), vol);
    }
    double ls[3] = { std::log(sg[0]), std::log(sg[1]), std::log(sg[2]) };   // SAND HS:326-339
    double tr = ls[0] + ls[1] + ls[2];
    M3 Dg = diag3(2 * mu * (1.0 / sg[0]) * ls[0] + lambda * tr * (1.0 / sg[0]),
                  2 * mu * (1.0 / sg[1]) * ls[1] + lambda * tr * (1.0 / sg[1]),
                  2 * mu * (1.0 / sg[2]) * ls[2] + lambda * tr * (1.0 / sg[2]));
    return scale(mul(mul(mul(U, Dg), transpose(V)), transpose(FE)), vol);
}

// HS:252-458 computeGridForces_
void compute_grid_forces(Sim& S, double Dt) {
    std::fill(S.gf.begin(), S.gf.end(), 0.0);
    const long Ng = S.Ng; const bool par = S.nthreads > 1;
    if (S.Np) {
        double lambda0 = S.E * S.nu / (1.0 + S.nu) / (1.0 - 2.0 * S.nu), mu0 = S.E / 2.0 / (1.0 + S.nu);   // HS:264-265
#pragma omp parallel for schedule(static) num_threads(S.nthreads)
        for (long p = 0; p < S.Np; ++p) {
            M3 FE = load_m3(S.FE, p);
            M3 Mod = scale(grad_field(S, S.sp, p, S.gv), Dt);              // HS:269-271
            M3 Fh = add(FE, mul(Mod, FE));                                 // HS:306
            M3 stress = particle_stress(S, Fh, FE, load_m3(S.FP, p), S.vol[p], lambda0, mu0);
            scatter_stress(S, S.sp, p, stress, par);
        }
    }
    if (S.Nv) {                                                            // HS:370-455
        std::vector<double> vf, pk; cloth_in_plane(S, vf, pk);
        const long Nv = S.Nv, Nf = S.Nf;
        for (long q = 0; q < Nv; ++q)                                      // HS:378  forces += vertexOmegas^T * vertexInPlaneForces
            for (int e = 0; e < S.sv.cnt[q]; ++e) {
                long s = 64 * q + e; int id = S.sv.idx[s];
                for (int r = 0; r < 3; ++r) S.gf[r * Ng + id] += S.sv.w[s] * vf[r * Nv + q];
            }
        for (long f = 0; f < Nf; ++f) {
            M3 Dm, dm;
            for (int r = 0; r < 3; ++r) {
                Dm(r,0) = S.eD1[r * Nf + f]; Dm(r,1) = S.eD2[r * Nf + f]; Dm(r,2) = S.eD3[r * Nf + f];
                dm(r,0) = S.ed1[r * Nf + f]; dm(r,1) = S.ed2[r * Nf + f]; dm(r,2) = S.ed3[r * Nf + f];
            }
            M3 Q, R; gram_schmidt(Q, R, dm);                               // HS:401-402
            double dr11 = pk[4 * f + 0], dr12 = pk[4 * f + 2], dr22 = pk[4 * f + 3];                // HS:406-409
            double dr13 = S.mesh_gamma * R(0,2), dr23 = S.mesh_gamma * R(1,2);                      // HS:411-412
            double dr33 = R(2,2) > 1.0 ? 0.0 : -S.mesh_k * (1.0 - R(2,2)) * (1.0 - R(2,2));         // HS:414-415
            M3 dR = m3_zero(); dR(0,0) = dr11; dR(0,1) = dr12; dR(0,2) = dr13; dR(1,1) = dr22; dR(1,2) = dr23; dR(2,2) = dr33;
            M3 K = mul(dR, transpose(R));                                  // HS:423
            M3 Sy = m3_zero();                                             // strictUpper(K) + upper(K)^T (HS:425-426)
            for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) {
                double su = (c > r) ? K(r,c) : 0.0;                        // strictly upper part of K at (r,c)
                double ut = (r >= c) ? K(c,r) : 0.0;                       // (upper(K))^T at (r,c) = upper(K)(c,r), nonzero iff c <= r
                Sy(r,c) = su + ut;
            }
            M3 RinvT = transpose(inverse(R));
            M3 T = mul(mul(Q, Sy), RinvT);
            // (restDirectionMatrix^T).col(2) = row 2 of the rest matrix = (D1.z, D2.z, D3.z)   (HS:427)
            double rc[3] = { Dm(2,0), Dm(2,1), Dm(2,2) };
            double dF3[3]; for (int r = 0; r < 3; ++r) dF3[r] = T(r,0) * rc[0] + T(r,1) * rc[1] + T(r,2) * rc[2];
            M3 A;                                                          // HS:429 volume * dF3 * d3^T
            for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) A(r,c) = S.evol[f] * dF3[r] * dm(c,2);
            scatter_stress(S, S.se, f, A, false);
        }
    }
#pragma omp parallel for schedule(static) num_threads(S.nthreads)
    for (long c = 0; c < Ng; ++c) S.gf[2 * Ng + c] -= S.gm[c] * S.gravity; // HS:457
}

// HS:725-737 updateGridVelocities_
void update_grid_velocities(Sim& S, double Dt) {
    const long Ng = S.Ng;
#pragma omp parallel for schedule(static) num_threads(S.nthreads)
    for (long c = 0; c < Ng; ++c)
        if (S.gm[c] > 0.0) for (int a = 0; a < 3; ++a) S.gv[a * Ng + c] += Dt * S.gf[a * Ng + c] / S.gm[c];
}

// RG:188-200 max_velocity, RegularGrid.h:60 CFL_condition
double cfl_condition(const Sim& S) {
    double mv = 0.0; const long Ng = S.Ng;
#pragma omp parallel for schedule(static) reduction(max:mv) num_threads(S.nthreads)
    for (long c = 0; c < Ng; ++c) {
        double n = std::sqrt(S.gv[c] * S.gv[c] + S.gv[Ng + c] * S.gv[Ng + c] + S.gv[2 * Ng + c] * S.gv[2 * Ng + c]);
        if (n > mv) mv = n;
    }
    return mv / S.hmin();
}

// HS:460-551 gridCollisionHandling_
void grid_collision(Sim& S) {
    S.gvt = S.gv;                                                          // HS:463
    const long Ng = S.Ng;
    if (S.ls_kind != LS_NONE) {
#pragma omp parallel for schedule(static) num_threads(S.nthreads)
        for (int k = 0; k < S.res[2]; ++k) for (int j = 0; j < S.res[1]; ++j) for (int i = 0; i < S.res[0]; ++i) {
            int id = S.to_index(i, j, k);
            double gp[3] = { S.mn[0] + i * S.h[0], S.mn[1] + j * S.h[1], S.mn[2] + k * S.h[2] };    // HS:473-476
            bool inside; double n[3];
            if (S.ls_kind == LS_SAMPLED) { inside = S.ls_inside[id] != 0; n[0] = S.ls_nrm[id]; n[1] = S.ls_nrm[Ng + id]; n[2] = S.ls_nrm[2 * Ng + id]; }
            else { inside = ls_phi(S.ls_kind, S.ls_par, gp) <= 0.0; if (inside) ls_normal(S.ls_kind, S.ls_par, gp, n); }
            if (!inside) continue;                                         // HS:478
            double vr[3] = { S.gv[id], S.gv[Ng + id], S.gv[2 * Ng + id] };
            double vn = vr[0] * n[0] + vr[1] * n[1] + vr[2] * n[2];        // HS:486
            if (vn < 0.0) {                                                // HS:488
                double vt[3] = { vr[0] - vn * n[0], vr[1] - vn * n[1], vr[2] - vn * n[2] };
                for (int a = 0; a < 3; ++a) S.gvt[a * Ng + id] = vt[a];    // HS:492
                double vtn = std::sqrt(vt[0] * vt[0] + vt[1] * vt[1] + vt[2] * vt[2]);
                // HS:494-502: stick if |vt| < -mu vn, else vRef = vt.  The Coulomb reduction on HS:501 is a
                // separate expression statement (line 500 ends with ';') and has no effect -- reproduced.
                bool stick = vtn < -S.friction * vn;
                for (int a = 0; a < 3; ++a) S.gv[a * Ng + id] = stick ? 0.0 : vt[a];                 // HS:504-506
            }
        }
    }
    if (S.Nv && !S.fixedv.empty()) {                                       // HS:513-550
        for (long q = 0; q < S.Nv; ++q) {
            if (S.fixedv[q] == 0.0) continue;                              // LM:462-481 vertexIsFixed
            for (int e = 0; e < S.sv.cnt[q]; ++e) {
                int gid = S.sv.idx[64 * q + e];
                int rk = gid / (S.res[0] * S.res[1]); int rj = (gid % (S.res[0] * S.res[1])) / S.res[0];    // RG:170-177
                int ri = gid - rk * S.res[0] * S.res[1] - rj * S.res[0];
                for (int i = ri - 1; i <= ri + 1; ++i) for (int j = rj - 1; j <= rj + 1; ++j) for (int k = rk - 1; k <= rk + 1; ++k) {
                    // the "<= 12" test of HS:533-536 is always true inside a 3x3x3 block; per-axis range is NOT
                    // checked, only the flat index (HS:538-539) -> wraps across rows at domain faces. Reproduced.
                    long index = static_cast<long>(k) * S.res[0] * S.res[1] + static_cast<long>(j) * S.res[0] + i;
                    if (0 <= index && index < Ng) for (int a = 0; a < 3; ++a) { S.gv[a * Ng + index] = 0.0; S.gvt[a * Ng + index] = 0.0; }
                }
            }
        }
    }
}

// row p of omegas * field (Ng x 3)
inline void gather3(const Sim& S, const Stencils& st, long p, const std::vector<double>& fld, double out[3]) {
    out[0] = out[1] = out[2] = 0.0; const long Ng = S.Ng;
    for (int e = 0; e < st.cnt[p]; ++e) { long s = 64 * p + e; int id = st.idx[s]; double w = st.w[s];
        out[0] += w * fld[id]; out[1] += w * fld[Ng + id]; out[2] += w * fld[2 * Ng + id]; }
}

// HS:739-758 updateParticleVelocities_
void update_particle_velocities(Sim& S) {
#pragma omp parallel for schedule(static) num_threads(S.nthreads)
    for (long p = 0; p < S.Np; ++p) { double o[3]; gather3(S, S.sp, p, S.gv, o); for (int a = 0; a < 3; ++a) S.v[a * S.Np + p] = o[a]; }
    if (S.Nv) {
        for (long q = 0; q < S.Nv; ++q) { double o[3]; gather3(S, S.sv, q, S.gv, o); for (int a = 0; a < 3; ++a) S.vv[a * S.Nv + q] = o[a]; }
        for (long f = 0; f < S.Nf; ++f) {                                  // HS:749-756
            int a = S.faces[f], b = S.faces[S.Nf + f], c = S.faces[2 * S.Nf + f];
            for (int r = 0; r < 3; ++r) S.ev[r * S.Nf + f] = (S.vv[r * S.Nv + a] + S.vv[r * S.Nv + b] + S.vv[r * S.Nv + c]) / 3.0;
        }
    }
}

// HS:760-825 updateAffineMomenta_
void update_affine(Sim& S, const Stencils& st, long np, const std::vector<double>& pos,
                   std::vector<double>& b1, std::vector<double>& b2, std::vector<double>& b3, double damp) {
    const long Ng = S.Ng;
#pragma omp parallel for schedule(static) num_threads(S.nthreads)
    for (long p = 0; p < np; ++p) {
        double C[3][3] = {{0}}; double vt[3] = {0, 0, 0};
        double xp[3] = { pos[p], pos[np + p], pos[2 * np + p] };
        for (int e = 0; e < st.cnt[p]; ++e) {
            long s = 64 * p + e; int id = st.idx[s]; double w = st.w[s];
            int k = id / (S.res[0] * S.res[1]); int j = (id % (S.res[0] * S.res[1])) / S.res[0]; int i = id - k * S.res[0] * S.res[1] - j * S.res[0];
            double xi[3] = { S.mn[0] + i * S.h[0], S.mn[1] + j * S.h[1], S.mn[2] + k * S.h[2] };
            for (int a = 0; a < 3; ++a) { double vi = S.gv[a * Ng + id]; vt[a] += w * vi; for (int b = 0; b < 3; ++b) C[a][b] += w * vi * xi[b]; }   // HS:797-806
        }
        for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) C[a][b] -= vt[a] * xp[b];
        double sym[3][3], out[3][3];                                       // HS:808-824
        for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) sym[a][b] = (a == b) ? C[a][a] : 0.5 * (C[a][b] + C[b][a]);
        for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) out[a][b] = (C[a][b] - sym[a][b]) + (1 - damp) * sym[a][b];
        for (int b = 0; b < 3; ++b) { b1[b * np + p] = out[0][b]; b2[b * np + p] = out[1][b]; b3[b * np + p] = out[2][b]; }
    }
}

// HS:940-951 advection: x = omegas * (x_i + Dt * v~_i)
void advect(Sim& S, const Stencils& st, long np, std::vector<double>& pos, double Dt) {
    const long Ng = S.Ng;
#pragma omp parallel for schedule(static) num_threads(S.nthreads)
    for (long p = 0; p < np; ++p) {
        double o[3] = {0, 0, 0};
        for (int e = 0; e < st.cnt[p]; ++e) {
            long s = 64 * p + e; int id = st.idx[s]; double w = st.w[s];
            int k = id / (S.res[0] * S.res[1]); int j = (id % (S.res[0] * S.res[1])) / S.res[0]; int i = id - k * S.res[0] * S.res[1] - j * S.res[0];
            double xi[3] = { S.mn[0] + i * S.h[0], S.mn[1] + j * S.h[1], S.mn[2] + k * S.h[2] };
            for (int a = 0; a < 3; ++a) o[a] += w * (xi[a] + Dt * S.gvt[a * Ng + id]);
        }
        for (int a = 0; a < 3; ++a) pos[a * np + p] = o[a];
    }
}

void update_element_positions(Sim& S) {                                   // LM:371-380
    for (long f = 0; f < S.Nf; ++f) {
        int a = S.faces[f], b = S.faces[S.Nf + f], c = S.faces[2 * S.Nf + f];
        for (int r = 0; r < 3; ++r) S.ex[r * S.Nf + f] = (S.vx[r * S.Nv + a] + S.vx[r * S.Nv + b] + S.vx[r * S.Nv + c]) / 3.0;
    }
}

// HS:553-609 updateDeformationGradient_
void update_deformation_gradient(Sim& S, double Dt) {
    if (S.Np) {
        S.cand.resize(9 * S.Np);
#pragma omp parallel for schedule(static) num_threads(S.nthreads)
        for (long p = 0; p < S.Np; ++p) {
            M3 Mod = scale(grad_field(S, S.sp, p, S.gvt), Dt);             // HS:560-562
            M3 FE = load_m3(S.FE, p);
            store_m3(S.cand, p, add(FE, mul(Mod, FE)));                    // HS:575-576
        }
    }
    if (S.Nv) {
        const long Nf = S.Nf, Nv = S.Nv;
        for (long f = 0; f < Nf; ++f) {
            M3 G = grad_field(S, S.se, f, S.gvt);                          // HS:584-586 (not yet scaled by Dt)
            int a = S.faces[f], b = S.faces[Nf + f], c = S.faces[2 * Nf + f];
            double d3o[3] = { S.ed3[f], S.ed3[Nf + f], S.ed3[2 * Nf + f] };
            for (int r = 0; r < 3; ++r) {
                S.ed1[r * Nf + f] = S.vx[r * Nv + b] - S.vx[r * Nv + a];   // HS:591-592 (vertices already advected)
                S.ed2[r * Nf + f] = S.vx[r * Nv + c] - S.vx[r * Nv + a];   // HS:593-594
                S.ed3[r * Nf + f] = Dt * (G(r,0) * d3o[0] + G(r,1) * d3o[1] + G(r,2) * d3o[2]) + d3o[r];   // HS:601-602
            }
        }
    }
}

// HS:616-679 body of the particle loop
void particle_return_map(const Sim& S, const M3& Fc, M3& FE_out, M3& FP, double& q) {
    const double PI = 3.14159265358979323846;                              // igl::PI
    M3 Ftot = mul(Fc, FP);                                                 // HS:618-619
    M3 U, V; double sg[3]; svd3(Fc, U, sg, V);                             // HS:620
    if (S.material == 0) {                                                 // HS:626-631
        for (int i = 0; i < 3; ++i) sg[i] = clampd(sg[i], 1.0 - S.thetaC, 1.0 + S.thetaS);
    } else {                                                               // HS:632-674
        double lambda = S.E * S.nu / (1.0 + S.nu) / (1.0 - 2.0 * S.nu), mu = S.E / 2.0 / (1.0 + S.nu);
        double phi = (S.sand_h[0] + (S.sand_h[1] * q - S.sand_h[3]) * std::exp(-S.sand_h[2] * q)) * PI / 180.0;   // HS:646-647
        double alpha = std::sqrt(2.0 / 3.0) * 2.0 * std::sin(phi) / (3.0 - std::sin(phi));                         // HS:649-650
        double ls[3] = { std::log(sg[0]), std::log(sg[1]), std::log(sg[2]) };
        double tr = ls[0] + ls[1] + ls[2];
        double dv[3] = { ls[0] - tr / 3.0 * 1.0, ls[1] - tr / 3.0 * 1.0, ls[2] - tr / 3.0 * 1.0 };
        double dvn = std::sqrt(dv[0] * dv[0] + dv[1] * dv[1] + dv[2] * dv[2]);
        double dg = dvn + (3.0 * lambda + 2.0 * mu) / 2.0 / mu * tr * alpha;                                        // HS:654-656
        if (dg <= 0.0) {
        } else if (dvn == 0.0 || tr > 0.0) {                               // HS:662-666
            q += std::sqrt(ls[0] * ls[0] + ls[1] * ls[1] + ls[2] * ls[2]);
            sg[0] = sg[1] = sg[2] = 1.0;
        } else {                                                           // HS:667-673
            for (int i = 0; i < 3; ++i) sg[i] = std::exp(ls[i] - dg * dv[i] / dvn);
            q += dg;
        }
    }
    FE_out = mul(mul(U, diag3(sg[0], sg[1], sg[2])), transpose(V));                                                  // HS:675
    FP = mul(mul(mul(V, diag3(1.0 / sg[0], 1.0 / sg[1], 1.0 / sg[2])), transpose(U)), Ftot);                         // HS:676-677
}

// HS:612-723 updatePlasticity_
void update_plasticity(Sim& S) {
    if (S.Np) {
#pragma omp parallel for schedule(static) num_threads(S.nthreads)
        for (long p = 0; p < S.Np; ++p) {
            M3 FE, FP = load_m3(S.FP, p); double qp = S.q[p];
            particle_return_map(S, load_m3(S.cand, p), FE, FP, qp);
            store_m3(S.FE, p, FE); store_m3(S.FP, p, FP); S.q[p] = qp;
        }
    }
    if (S.Nv) {                                                            // HS:684-722
        const long Nf = S.Nf;
        for (long f = 0; f < Nf; ++f) {
            M3 dm; for (int r = 0; r < 3; ++r) { dm(r,0) = S.ed1[r * Nf + f]; dm(r,1) = S.ed2[r * Nf + f]; dm(r,2) = S.ed3[r * Nf + f]; }
            M3 Q, R; gram_schmidt(Q, R, dm);
            if (R(2,2) > 1.0) { R(2,2) = 1.0; R(0,2) = R(1,2) = 0.0; }
            else {
                double fn = S.mesh_k * (R(2,2) - 1.0) * (R(2,2) - 1.0);
                double fs = S.mesh_gamma * std::sqrt(R(0,2) * R(0,2) + R(1,2) * R(1,2));
                if (fs > S.mesh_cf * fn) { R(0,2) *= S.mesh_cf * fn / fs; R(1,2) *= S.mesh_cf * fn / fs; }
            }
            for (int r = 0; r < 3; ++r) S.ed3[r * Nf + f] = Q(r,0) * R(0,2) + Q(r,1) * R(1,2) + Q(r,2) * R(2,2);              // HS:718-719
        }
    }
}

void rebuild_weights(Sim& S) {                                             // HS:830-850, HS:963-983
    if (S.Np) build_stencils(S, S.sp, S.x, S.Np);
    if (S.Nv) { build_stencils(S, S.sv, S.vx, S.Nv); build_stencils(S, S.se, S.ex, S.Nf); }
}

double now_s() {
#ifdef _OPENMP
    return omp_get_wtime();
#else
    timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec;
#endif
}

// stages 4-9 of one loop iteration given Dt (HS:899-959)
void g2p_block(Sim& S, double Dt) {
    update_particle_velocities(S);                                         // HS:903
    if (S.Np) update_affine(S, S.sp, S.Np, S.x, S.B1, S.B2, S.B3, 0.0);    // HS:908-917
    if (S.Nv) {                                                            // HS:918-935
        update_affine(S, S.sv, S.Nv, S.vx, S.vB1, S.vB2, S.vB3, 1.0);
        update_affine(S, S.se, S.Nf, S.ex, S.eB1, S.eB2, S.eB3, 1.0);
    }
    if (S.Np) advect(S, S.sp, S.Np, S.x, Dt);                              // HS:942-945
    if (S.Nv) { advect(S, S.sv, S.Nv, S.vx, Dt); update_element_positions(S); }   // HS:946-950
    update_deformation_gradient(S, Dt);                                    // HS:955
    update_plasticity(S);                                                  // HS:959
}

}  // namespace

// ============================================================================ C API (ctypes)
extern "C" {

typedef struct orc_sim orc_sim;   // opaque = Sim

orc_sim* orc_create(const double* mn, const double* mx, const int* res) {
    Sim* S = new Sim();
    for (int a = 0; a < 3; ++a) { S->mn[a] = mn[a]; S->mx[a] = mx[a]; S->res[a] = res[a]; S->h[a] = (mx[a] - mn[a]) / res[a]; }   // RG:137-139
    S->Ng = static_cast<long>(res[0]) * res[1] * res[2];
    S->gm.assign(S->Ng, 0.0); S->gv.assign(3 * S->Ng, 0.0); S->gf.assign(3 * S->Ng, 0.0); S->gvt.assign(3 * S->Ng, 0.0);
    return reinterpret_cast<orc_sim*>(S);
}
void orc_destroy(orc_sim* h) { delete reinterpret_cast<Sim*>(h); }

void orc_set_threads(orc_sim* h, int n) {
    Sim* S = reinterpret_cast<Sim*>(h);
#ifdef _OPENMP
    S->nthreads = n > 0 ? n : omp_get_max_threads();
#else
    (void)n; S->nthreads = 1;
#endif
}
int orc_get_threads(orc_sim* h) { return reinterpret_cast<Sim*>(h)->nthreads; }

// material: 0 SNOW, 1 SAND (HybridSolver.h:21-25)
void orc_set_params(orc_sim* h, int material, double cfl, double gravity, double friction, double snow_xi,
                    const double* sand_h4, double rate_floor, double frame_dt) {
    Sim* S = reinterpret_cast<Sim*>(h);
    S->material = material; S->cfl = cfl; S->gravity = gravity; S->friction = friction; S->snow_xi = snow_xi;
    for (int i = 0; i < 4; ++i) S->sand_h[i] = sand_h4[i];
    S->rate_floor = rate_floor; S->frame_dt = frame_dt;
}

void orc_set_particles(orc_sim* h, long n, const double* x, const double* v, const double* B1, const double* B2, const double* B3,
                       const double* FE, const double* FP, const double* m, const double* vol, const double* q,
                       double E, double nu, double thetaC, double thetaS) {
    Sim* S = reinterpret_cast<Sim*>(h); S->Np = n;
    S->x.assign(x, x + 3 * n); S->v.assign(v, v + 3 * n);
    S->B1.assign(B1, B1 + 3 * n); S->B2.assign(B2, B2 + 3 * n); S->B3.assign(B3, B3 + 3 * n);
    S->FE.assign(FE, FE + 9 * n); S->FP.assign(FP, FP + 9 * n); S->cand = S->FE;
    S->m.assign(m, m + n); S->vol.assign(vol, vol + n); S->dens.assign(n, 1.0); S->q.assign(q, q + n);
    S->E = E; S->nu = nu; S->thetaC = thetaC; S->thetaS = thetaS;
}

void orc_set_mesh(orc_sim* h, long nv, long nf, const double* vx, const double* vv, const double* vm, const double* vvol,
                  const double* vB /*3 blocks of nv x 3*/, const int* faces /*nf x 3 col-major*/,
                  const double* ev, const double* em, const double* evol, const double* eB /*3 blocks of nf x 3*/,
                  const double* ed /*d1,d2,d3: 3 blocks nf x 3*/, const double* eD /*rest D1,D2,D3*/, const double* fixedv /*nv or NULL*/,
                  double mu, double lambda, double shear, double stiff, double fric) {
    Sim* S = reinterpret_cast<Sim*>(h); S->Nv = nv; S->Nf = nf;
    S->vx.assign(vx, vx + 3 * nv); S->vv.assign(vv, vv + 3 * nv); S->vm.assign(vm, vm + nv); S->vvol.assign(vvol, vvol + nv);
    S->vB1.assign(vB, vB + 3 * nv); S->vB2.assign(vB + 3 * nv, vB + 6 * nv); S->vB3.assign(vB + 6 * nv, vB + 9 * nv);
    S->faces.assign(faces, faces + 3 * nf);
    S->ev.assign(ev, ev + 3 * nf); S->em.assign(em, em + nf); S->evol.assign(evol, evol + nf);
    S->eB1.assign(eB, eB + 3 * nf); S->eB2.assign(eB + 3 * nf, eB + 6 * nf); S->eB3.assign(eB + 6 * nf, eB + 9 * nf);
    S->ed1.assign(ed, ed + 3 * nf); S->ed2.assign(ed + 3 * nf, ed + 6 * nf); S->ed3.assign(ed + 6 * nf, ed + 9 * nf);
    S->eD1.assign(eD, eD + 3 * nf); S->eD2.assign(eD + 3 * nf, eD + 6 * nf); S->eD3.assign(eD + 6 * nf, eD + 9 * nf);
    if (fixedv) S->fixedv.assign(fixedv, fixedv + nv); else S->fixedv.clear();
    S->mesh_mu = mu; S->mesh_lambda = lambda; S->mesh_gamma = shear; S->mesh_k = stiff; S->mesh_cf = fric;
    S->ex.assign(3 * nf, 0.0); update_element_positions(*S);                // LM:169-185 ctor: element positions from vertices
}

void orc_set_levelset(orc_sim* h, int kind, const double* params8) {
    Sim* S = reinterpret_cast<Sim*>(h); S->ls_kind = kind; for (int i = 0; i < 8; ++i) S->ls_par[i] = params8[i];
}
void orc_set_levelset_samples(orc_sim* h, const uint8_t* inside, const double* normal) {
    Sim* S = reinterpret_cast<Sim*>(h); S->ls_kind = LS_SAMPLED;
    S->ls_inside.assign(inside, inside + S->Ng); S->ls_nrm.assign(normal, normal + 3 * S->Ng);
}

// ---- stage-level entry points (same split as aep_stage_* in include/aep_b200.h)
void orc_rebuild_weights(orc_sim* h) { rebuild_weights(*reinterpret_cast<Sim*>(h)); }
void orc_p2g(orc_sim* h, int first) { particle_to_grid(*reinterpret_cast<Sim*>(h), first != 0); }
void orc_stage_forces(orc_sim* h, double dt) { compute_grid_forces(*reinterpret_cast<Sim*>(h), dt); }
void orc_stage_grid_update(orc_sim* h, double dt) { update_grid_velocities(*reinterpret_cast<Sim*>(h), dt); }
double orc_cfl_condition(orc_sim* h) { return cfl_condition(*reinterpret_cast<Sim*>(h)); }
void orc_stage_collide(orc_sim* h) { grid_collision(*reinterpret_cast<Sim*>(h)); }
void orc_stage_g2p(orc_sim* h, double dt) { g2p_block(*reinterpret_cast<Sim*>(h), dt); }

// HS:830-860: initial weights, first P2G (volumes), initial Dt
void orc_init(orc_sim* h) {
    Sim* S = reinterpret_cast<Sim*>(h);
    rebuild_weights(*S); particle_to_grid(*S, true);
    S->t = 0.0; S->inner_t = 0.0; S->frame_flag = 0; S->frame_no = 0;
    S->dt = S->cfl / std::max(S->rate_floor, cfl_condition(*S));           // HS:860
}

// One iteration of the while loop, HS:867-1032 (without the OBJ dump).  Returns the Dt used for advection.
double orc_substep(orc_sim* h) {
    Sim* S = reinterpret_cast<Sim*>(h);
    double t0 = now_s();
    compute_grid_forces(*S, S->dt);                                        // HS:873 (Dt of the previous iteration)
    double t1 = now_s();
    update_grid_velocities(*S, S->dt);                                     // HS:877
    S->dt = S->cfl / std::max(S->rate_floor, cfl_condition(*S));           // HS:878
    if (S->inner_t + S->dt >= S->frame_dt) {                               // HS:880-892
        S->dt = S->frame_dt - S->inner_t; S->t += S->frame_dt; S->inner_t = 0.0; S->frame_flag = 1;
    } else S->inner_t += S->dt;
    grid_collision(*S);                                                    // HS:899
    double t2 = now_s();
    g2p_block(*S, S->dt);                                                  // HS:903-959
    double t3 = now_s();
    rebuild_weights(*S);                                                   // HS:963-983
    double t4 = now_s();
    particle_to_grid(*S, false);                                           // HS:987
    double t5 = now_s();
    if (S->frame_flag) { S->frame_no++; S->frame_flag = 0; }               // HS:991-1030 (OBJ dump elided)
    S->tm[0] += t1 - t0; S->tm[1] += t2 - t1; S->tm[2] += t3 - t2; S->tm[3] += t4 - t3; S->tm[4] += t5 - t4;
    return S->dt;
}

void orc_set_dt(orc_sim* h, double dt) { reinterpret_cast<Sim*>(h)->dt = dt; }
double orc_get_dt(orc_sim* h) { return reinterpret_cast<Sim*>(h)->dt; }
double orc_get_time(orc_sim* h) { return reinterpret_cast<Sim*>(h)->t; }
int orc_get_frame(orc_sim* h) { return reinterpret_cast<Sim*>(h)->frame_no; }
void orc_get_timers(orc_sim* h, double* out5) { Sim* S = reinterpret_cast<Sim*>(h); for (int i = 0; i < 5; ++i) out5[i] = S->tm[i]; }

static void cp(double* dst, const std::vector<double>& src) { if (dst && !src.empty()) std::memcpy(dst, src.data(), src.size() * sizeof(double)); }

void orc_get_particles(orc_sim* h, double* x, double* v, double* B1, double* B2, double* B3, double* FE, double* FP,
                       double* vol, double* q) {
    Sim* S = reinterpret_cast<Sim*>(h);
    cp(x, S->x); cp(v, S->v); cp(B1, S->B1); cp(B2, S->B2); cp(B3, S->B3); cp(FE, S->FE); cp(FP, S->FP); cp(vol, S->vol); cp(q, S->q);
}
void orc_get_grid(orc_sim* h, double* m, double* v, double* f, double* vt) {
    Sim* S = reinterpret_cast<Sim*>(h); cp(m, S->gm); cp(v, S->gv); cp(f, S->gf); cp(vt, S->gvt);
}
// test hooks for the slab-decomposition driver test (tests/oracle_slab_backend.py): overwrite grid fields after a halo
// exchange, and the volume initialisation of HS:242-249 on its own (needs the COMPLETE grid mass).
void orc_set_grid(orc_sim* h, const double* m, const double* v, const double* f) {
    Sim* S = reinterpret_cast<Sim*>(h);
    if (m) S->gm.assign(m, m + S->Ng);
    if (v) S->gv.assign(v, v + 3 * S->Ng);
    if (f) S->gf.assign(f, f + 3 * S->Ng);
}
void orc_compute_volumes(orc_sim* h) {
    Sim* S = reinterpret_cast<Sim*>(h);
    double gvol = S->h[0] * S->h[1] * S->h[2];
    for (long p = 0; p < S->Np; ++p) {
        double d = 0.0;
        for (int e = 0; e < S->sp.cnt[p]; ++e) d += S->sp.w[64 * p + e] * S->gm[S->sp.idx[64 * p + e]];
        S->dens[p] = d / gvol; S->vol[p] = S->m[p] * (1.0 / S->dens[p]);
    }
}

void orc_get_mesh(orc_sim* h, double* vx, double* vv, double* vB, double* ex, double* ev, double* eB, double* ed) {
    Sim* S = reinterpret_cast<Sim*>(h);
    cp(vx, S->vx); cp(vv, S->vv);
    if (vB) { cp(vB, S->vB1); cp(vB + 3 * S->Nv, S->vB2); cp(vB + 6 * S->Nv, S->vB3); }
    cp(ex, S->ex); cp(ev, S->ev);
    if (eB) { cp(eB, S->eB1); cp(eB + 3 * S->Nf, S->eB2); cp(eB + 6 * S->Nf, S->eB3); }
    if (ed) { cp(ed, S->ed1); cp(ed + 3 * S->Nf, S->ed2); cp(ed + 6 * S->Nf, S->ed3); }
}

// unit-test hooks for the math kernels
double orc_cubic_bspline(double x) { return cubic_bspline(x); }
double orc_dcubic_bspline(double x) { return dcubic_bspline(x); }
void orc_svd3(const double* F9, double* U9, double* s3, double* V9) {
    M3 F, U, V; std::memcpy(F.a, F9, 72); svd3(F, U, s3, V); std::memcpy(U9, U.a, 72); std::memcpy(V9, V.a, 72);
}
void orc_svd2(const double* A4, double* U4, double* s2, double* V4) { svd2(A4, U4, s2, V4); }
void orc_gram_schmidt(const double* A9, double* Q9, double* R9) {
    M3 A, Q, R; std::memcpy(A.a, A9, 72); gram_schmidt(Q, R, A); std::memcpy(Q9, Q.a, 72); std::memcpy(R9, R.a, 72);
}

// single-particle hooks: same code path as the loops above (used to pin the fp32 device math)
void orc_particle_stress(int material, double E, double nu, double snow_xi, const double* Fh9, const double* FE9, const double* FP9,
                         double vol, double* A9) {
    Sim S; S.material = material; S.snow_xi = snow_xi;
    double lambda0 = E * nu / (1.0 + nu) / (1.0 - 2.0 * nu), mu0 = E / 2.0 / (1.0 + nu);
    M3 Fh, FE, FP; std::memcpy(Fh.a, Fh9, 72); std::memcpy(FE.a, FE9, 72); std::memcpy(FP.a, FP9, 72);
    M3 A = particle_stress(S, Fh, FE, FP, vol, lambda0, mu0); std::memcpy(A9, A.a, 72);
}
void orc_particle_return_map(int material, double E, double nu, double thetaC, double thetaS, const double* Fc9, double* FE9,
                             double* FP9, double* q) {
    Sim S; S.material = material; S.E = E; S.nu = nu; S.thetaC = thetaC; S.thetaS = thetaS;
    M3 Fc, FE, FP; std::memcpy(Fc.a, Fc9, 72); std::memcpy(FP.a, FP9, 72);
    particle_return_map(S, Fc, FE, FP, *q); std::memcpy(FE9, FE.a, 72); std::memcpy(FP9, FP.a, 72);
}
double orc_ls_phi(int kind, const double* P, const double* x) { return ls_phi(kind, P, x); }
void orc_ls_normal(int kind, const double* P, const double* x, double* n) { ls_normal(kind, P, x, n); }

}  // extern "C"
