// aep_math.cuh -- register-resident math for the MPM substep kernels (sm_100a, fp32).
//
// Restates, in a GPU-friendly form, the scalar kernels of the reference:
//   cubic_B_spline / Dcubic_B_spline      interpolation.cpp:9-33
//   Eigen::JacobiSVD<Matrix3d> contract   HybridSolver.cpp:308,620   (s >= 0; uses are order/sign invariant)
//   snow / sand stress                    HybridSolver.cpp:281-339
//   snow / sand return mapping            HybridSolver.cpp:626-677
//   Gram-Schmidt QR                       geometry.cpp:31-62
// All matrices are row-major float[9]: M[3*r + c].
#pragma once
#ifndef AEP_HOST_MATH_TEST   // tests/cpu_math_harness.cpp compiles this header for the host with shims
#include <cuda_runtime.h>
#endif

namespace aep {

// ------------------------------------------------------------------------------------------------ B-splines
// A particle sits in cell c with fractional offset f in [0,1).  Its four stencil nodes per axis are
// c-1+o, o = 0..3, at signed distance u = f + 1 - o (in cells).  With g = 1 - f:
//   N:  g^3/6 | f^3/2 - f^2 + 2/3 | g^3/2 - g^2 + 2/3 | f^3/6          (interpolation.cpp:9-16, |u| form)
//   N': -g^2/2 | 3f^2/2 - 2f | -3g^2/2 + 2g | f^2/2                    (interpolation.cpp:18-33)
// Nodes with weight <= 0 are dropped by the reference (HybridSolver.cpp:60); they contribute 0 here.
__device__ __forceinline__ void bspline4(float f, float (&N)[4], float (&D)[4]) {
    const float g = 1.0f - f;
    const float f2 = f * f, g2 = g * g;
    N[0] = g2 * g * (1.0f / 6.0f);
    N[1] = fmaf(f2, fmaf(0.5f, f, -1.0f), 2.0f / 3.0f);
    N[2] = fmaf(g2, fmaf(0.5f, g, -1.0f), 2.0f / 3.0f);
    N[3] = f2 * f * (1.0f / 6.0f);
    D[0] = -0.5f * g2;
    D[1] = f * fmaf(1.5f, f, -2.0f);
    D[2] = -g * fmaf(1.5f, g, -2.0f);
    D[3] = 0.5f * f2;
}

// One lane's node offset o (0..3): value and derivative of the weight of a particle at fraction f.
__device__ __forceinline__ void bspline_lane(float f, int o, float& N, float& D) {
    const bool lo = (o < 2);                 // o = 0,1 use f ; o = 2,3 use the mirrored variable
    const bool inner = (o == 1) || (o == 2);
    // a = |u| folded to [0,1): inner nodes a = f (o=1) or 1-f (o=2); outer nodes: t = 2-|u| = 1-f (o=0) or f (o=3)
    const float a = (o == 1 || o == 3) ? f : 1.0f - f;
    const float a2 = a * a;
    const float n_in = fmaf(a2, fmaf(0.5f, a, -1.0f), 2.0f / 3.0f);
    const float n_out = a2 * a * (1.0f / 6.0f);
    const float d_in = a * fmaf(1.5f, a, -2.0f);          // derivative w.r.t. a
    const float d_out = 0.5f * a2;                        // derivative w.r.t. t
    N = inner ? n_in : n_out;
    // du/da: o=1: +1, o=2: -1 ; d/du of outer: o=0: u=2-t -> -d_out ; o=3: u=t-2 -> +d_out
    const float d = inner ? d_in : d_out;
    D = (o == 1 || o == 3) ? d : -d;
    (void)lo;
}

// ------------------------------------------------------------------------------------------------ 3x3 helpers
__device__ __forceinline__ void mat_mul(const float (&A)[9], const float (&B)[9], float (&C)[9]) {
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c)
            C[3 * r + c] = fmaf(A[3 * r], B[c], fmaf(A[3 * r + 1], B[3 + c], A[3 * r + 2] * B[6 + c]));
}
// C = A * B^T
__device__ __forceinline__ void mat_mul_nt(const float (&A)[9], const float (&B)[9], float (&C)[9]) {
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c)
            C[3 * r + c] = fmaf(A[3 * r], B[3 * c], fmaf(A[3 * r + 1], B[3 * c + 1], A[3 * r + 2] * B[3 * c + 2]));
}
// C = A^T * B
__device__ __forceinline__ void mat_mul_tn(const float (&A)[9], const float (&B)[9], float (&C)[9]) {
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c)
            C[3 * r + c] = fmaf(A[r], B[c], fmaf(A[3 + r], B[3 + c], A[6 + r] * B[6 + c]));
}
__device__ __forceinline__ float mat_det(const float (&A)[9]) {
    return A[0] * (A[4] * A[8] - A[5] * A[7]) - A[1] * (A[3] * A[8] - A[5] * A[6]) + A[2] * (A[3] * A[7] - A[4] * A[6]);
}
// cofactor matrix: cof(A) = det(A) * A^{-T}
__device__ __forceinline__ void mat_cof(const float (&A)[9], float (&C)[9]) {
    C[0] = A[4] * A[8] - A[5] * A[7]; C[1] = A[5] * A[6] - A[3] * A[8]; C[2] = A[3] * A[7] - A[4] * A[6];
    C[3] = A[2] * A[7] - A[1] * A[8]; C[4] = A[0] * A[8] - A[2] * A[6]; C[5] = A[1] * A[6] - A[0] * A[7];
    C[6] = A[1] * A[5] - A[2] * A[4]; C[7] = A[2] * A[3] - A[0] * A[5]; C[8] = A[0] * A[4] - A[1] * A[3];
}

// ------------------------------------------------------------------------------------------------ SVD
// One-sided (Hestenes) Jacobi: rotate the columns of A = F V until orthogonal; s_c = |a_c|, u_c = a_c / s_c.
// Meets Eigen::JacobiSVD's contract up to ordering (s >= 0, U and V orthogonal, F = U diag(s) V^T); every use on
// the hot path is a symmetric function of (s, U, V) so ordering is irrelevant.  The rotation angle may be
// approximate (fast division / rsqrt): the iteration self-corrects, only c^2 + s^2 = 1 must hold to rounding,
// which one Newton step on rsqrt guarantees.  High relative accuracy of s near 1 (no F^T F squaring).
#ifndef AEP_SVD_TOL
#define AEP_SVD_TOL 1e-12f
#endif
#ifdef AEP_HOST_MATH_TEST
static long g_svd_sweeps = 0;
#endif
struct Svd3 {
    float U[9], S[3], V[9];
};

// returns cos^2 of the angle between columns P and Q before the rotation (0 when no rotation was needed)
template <int P, int Q>
__device__ __forceinline__ float jacobi_rot(float (&A)[3][3], float (&W)[3][3]) {
    const float al = fmaf(A[P][0], A[P][0], fmaf(A[P][1], A[P][1], A[P][2] * A[P][2]));
    const float be = fmaf(A[Q][0], A[Q][0], fmaf(A[Q][1], A[Q][1], A[Q][2] * A[Q][2]));
    const float ga = fmaf(A[P][0], A[Q][0], fmaf(A[P][1], A[Q][1], A[P][2] * A[Q][2]));
    const float ab = al * be;
    const bool act = ga * ga > 1e-14f * ab;
    const float zeta = __fdividef(be - al, act ? 2.0f * ga : 1.0f);
    const float az = fabsf(zeta);
    const float rt = fmaf(zeta, zeta, 1.0f);
    float t = __fdividef(1.0f, az + rt * rsqrtf(rt));
    t = act ? copysignf(t, zeta) : 0.0f;
    t = (t == t) ? t : 0.0f;                            // overflow of zeta^2 -> inf * 0: no rotation needed
    const float tt = fmaf(t, t, 1.0f);
    float c = rsqrtf(tt);
    c = c * fmaf(-0.5f * tt, c * c, 1.5f);              // Newton step: c^2 (1 + t^2) = 1 to rounding
    const float s = c * t;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const float ap = A[P][r], aq = A[Q][r];
        A[P][r] = fmaf(c, ap, -s * aq); A[Q][r] = fmaf(s, ap, c * aq);
        const float wp = W[P][r], wq = W[Q][r];
        W[P][r] = fmaf(c, wp, -s * wq); W[Q][r] = fmaf(s, wp, c * wq);
    }
    return act ? __fdividef(ga * ga, ab) : 0.0f;
}

__device__ __forceinline__ void svd3(const float (&F)[9], Svd3& o) {
    float A[3][3], W[3][3];                             // [column][row]
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int r = 0; r < 3; ++r) { A[c][r] = F[3 * r + c]; W[c][r] = (r == c) ? 1.0f : 0.0f; }
    // Cyclic Jacobi converges quadratically: once a sweep started from |cos| < 1e-6 on every pair, what it leaves behind
    // is below fp32 resolution even for (nearly) equal singular values, where convergence is only linear.  ~3.5 sweeps on average.
#pragma unroll 1
    for (int sweep = 0; sweep < 8; ++sweep) {
#ifdef AEP_HOST_MATH_TEST
        g_svd_sweeps += 1;
#endif
        float mx = jacobi_rot<0, 1>(A, W);
        mx = fmaxf(mx, jacobi_rot<0, 2>(A, W));
        mx = fmaxf(mx, jacobi_rot<1, 2>(A, W));
        if (mx < AEP_SVD_TOL) break;
    }
#ifdef AEP_HOST_MATH_TEST
    g_svd_sweeps += 0;
#endif
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float n2 = fmaf(A[c][0], A[c][0], fmaf(A[c][1], A[c][1], A[c][2] * A[c][2]));
        const float s = sqrtf(n2);
        const float inv = 1.0f / fmaxf(s, 1e-30f);
        o.S[c] = s;
#pragma unroll
        for (int r = 0; r < 3; ++r) { o.U[3 * r + c] = A[c][r] * inv; o.V[3 * r + c] = W[c][r]; }
    }
}

// M = U diag(d) V^T
__device__ __forceinline__ void usvt(const float (&U)[9], const float (&d)[3], const float (&V)[9], float (&M)[9]) {
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c)
            M[3 * r + c] = fmaf(U[3 * r] * d[0], V[3 * c], fmaf(U[3 * r + 1] * d[1], V[3 * c + 1], U[3 * r + 2] * d[2] * V[3 * c + 2]));
}

// ------------------------------------------------------------------------------------------------ materials
struct MatParams {
    float lambda0, mu0;        // Lame from E, nu                        HybridSolver.cpp:264-265
    float xi;                  // snow hardening                         HybridSolver.cpp:267
    float lo, hi;              // 1 - theta_c, 1 + theta_s               HybridSolver.cpp:628-630
    float h0, h1, h2, h3;      // sand friction-angle hardening          HybridSolver.cpp:641-644
    float k_vol;               // (3 lambda + 2 mu) / (2 mu)             HybridSolver.cpp:655
    int material;              // 0 snow, 1 sand
};

// ln s for a singular value s of an elastic deformation gradient (s ~ 1): s - 1 is exact in fp32 for s in [1/2, 2], and log1p keeps
// full relative accuracy of the *strain*, which is what the stress and the yield function see (logf(s) alone rounds s to 6e-8
// relative first, i.e. 1e-4 of a 1e-3 strain).
__device__ __forceinline__ float hencky(float s) { return log1pf(s - 1.0f); }

// ------------------------------------------------------------------------------------------------ sand without the SVD
// The Hencky model and the Drucker-Prager projection are ISOTROPIC functions of Fhat: everything the reference computes from the
// singular values (HybridSolver.cpp:326-339, 646-677) can be written with the symmetric matrix H = ln V_L = 1/2 ln(Fhat Fhat^T):
//     ln s_i are the eigenvalues of H       tr(ln s) = tr H       ||dev ln s|| = ||dev H||_F       ||ln s|| = ||H||_F
//     Kirchhoff stress  U diag(2 mu ln s + lambda tr) U^T = 2 mu H + lambda tr(H) I,   P = tau Fhat^-T
//     projected F_E'  = U exp(ln s - k dev ln s) V^T = exp(-k dev H) Fhat              (k = 1, tr/3 removed too, for the tensile apex)
//     F_P' = V S'^-1 U^T Fhat F_P = F_E'^-1 Fhat F_P
// A stiff granular material never strains by more than a few 1e-3 (E = 3.5e5 Pa against kPa of load), so E = Fhat Fhat^T - I is
// tiny and the matrix logarithm / exponential are 5-term series of 3x3 symmetric products: ~250 FMA, no iteration, no divergence,
// against 3-4 Jacobi sweeps (540 FMA-class instructions, data-dependent trip count) + the assembly from U, S, V.  And no
// ill-conditioned singular VECTORS on the way: for nearly equal singular values (always, here) the SVD path loses ~cos/(2 strain) of
// the stress to the residual non-orthogonality of U (see svd3).  Strains beyond |E|_F = 0.05 take the SVD path below.
struct Sym3 { float xx, yy, zz, xy, xz, yz; };
// product of two COMMUTING symmetric matrices (polynomials of the same matrix): symmetric again, 6 entries
__device__ __forceinline__ Sym3 sym_mul(const Sym3& a, const Sym3& b) {
    Sym3 c;
    c.xx = fmaf(a.xx, b.xx, fmaf(a.xy, b.xy, a.xz * b.xz));
    c.yy = fmaf(a.xy, b.xy, fmaf(a.yy, b.yy, a.yz * b.yz));
    c.zz = fmaf(a.xz, b.xz, fmaf(a.yz, b.yz, a.zz * b.zz));
    c.xy = fmaf(a.xx, b.xy, fmaf(a.xy, b.yy, a.xz * b.yz));
    c.xz = fmaf(a.xx, b.xz, fmaf(a.xy, b.yz, a.xz * b.zz));
    c.yz = fmaf(a.xy, b.xz, fmaf(a.yy, b.yz, a.yz * b.zz));
    return c;
}
// s * a + d * I
__device__ __forceinline__ Sym3 sym_axpi(float s, const Sym3& a, float d) {
    Sym3 c; c.xx = fmaf(s, a.xx, d); c.yy = fmaf(s, a.yy, d); c.zz = fmaf(s, a.zz, d); c.xy = s * a.xy; c.xz = s * a.xz; c.yz = s * a.yz; return c;
}
__device__ __forceinline__ float sym_norm2(const Sym3& a) {
    return fmaf(a.xx, a.xx, fmaf(a.yy, a.yy, a.zz * a.zz)) + 2.0f * fmaf(a.xy, a.xy, fmaf(a.xz, a.xz, a.yz * a.yz));
}
#ifndef AEP_SMALL_E2
#define AEP_SMALL_E2 2.5e-3f            // |E|_F^2 below which the series are used (|E|_F < 0.05: the 7th-order term is < 2e-9 |E|)
#endif
// E = F F^T - I from D = F - I (exact in fp32 for entries in [1/2, 2]): E = D + D^T + D D^T.  Returns |E|_F^2.
__device__ __forceinline__ float left_cauchy_green_minus_one(const float (&F)[9], Sym3& E) {
    const float D[9] = { F[0] - 1.0f, F[1], F[2], F[3], F[4] - 1.0f, F[5], F[6], F[7], F[8] - 1.0f };
    E.xx = fmaf(D[0], D[0], fmaf(D[1], D[1], D[2] * D[2])) + 2.0f * D[0];
    E.yy = fmaf(D[3], D[3], fmaf(D[4], D[4], D[5] * D[5])) + 2.0f * D[4];
    E.zz = fmaf(D[6], D[6], fmaf(D[7], D[7], D[8] * D[8])) + 2.0f * D[8];
    E.xy = fmaf(D[0], D[3], fmaf(D[1], D[4], D[2] * D[5])) + (D[1] + D[3]);
    E.xz = fmaf(D[0], D[6], fmaf(D[1], D[7], D[2] * D[8])) + (D[2] + D[6]);
    E.yz = fmaf(D[3], D[6], fmaf(D[4], D[7], D[5] * D[8])) + (D[5] + D[7]);
    return sym_norm2(E);
}
// H = 1/2 ln(I + E) = 1/2 E (I - E/2 + E^2/3 - E^3/4 + E^4/5 - E^5/6), Horner in the matrix.  e2 = |E|_F^2: below |E|_F = 4e-3 (the
// strains a stiff granular material lives at) three terms are exact to fp32 (|E|^3 / 4 < 2e-8 relative): 2 products instead of 5
#ifndef AEP_SHORT_E2
#define AEP_SHORT_E2 1.6e-5f
#endif
__device__ __forceinline__ Sym3 sym_half_log1p(const Sym3& E, float e2) {
    Sym3 t;
    if (e2 < AEP_SHORT_E2) t = sym_axpi(1.0f / 3.0f, E, -0.5f);
    else {
        t = sym_axpi(-1.0f / 6.0f, E, 0.2f);
        t = sym_axpi(1.0f, sym_mul(E, t), -0.25f);
        t = sym_axpi(1.0f, sym_mul(E, t), 1.0f / 3.0f);
        t = sym_axpi(1.0f, sym_mul(E, t), -0.5f);
    }
    t = sym_axpi(1.0f, sym_mul(E, t), 1.0f);
    return sym_axpi(0.5f, sym_mul(E, t), 0.0f);
}
// exp(X) = I + X (I + X/2 (I + X/3 (I + X/4 (I + X/5)))); |X|_F < 4e-3: I + X (I + X/2 (I + X/3)) (|X|^4 / 24 < 1e-11)
__device__ __forceinline__ Sym3 sym_exp(const Sym3& X, float x2) {
    Sym3 t;
    if (x2 < AEP_SHORT_E2) t = sym_axpi(1.0f / 3.0f, X, 1.0f);
    else {
        t = sym_axpi(0.2f, X, 1.0f);
        t = sym_axpi(0.25f, sym_mul(X, t), 1.0f);
        t = sym_axpi(1.0f / 3.0f, sym_mul(X, t), 1.0f);
    }
    t = sym_axpi(0.5f, sym_mul(X, t), 1.0f);
    return sym_axpi(1.0f, sym_mul(X, t), 1.0f);
}
// C = S * B for symmetric S
__device__ __forceinline__ void sym_mat_mul(const Sym3& S, const float (&B)[9], float (&C)[9]) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        C[c] = fmaf(S.xx, B[c], fmaf(S.xy, B[3 + c], S.xz * B[6 + c]));
        C[3 + c] = fmaf(S.xy, B[c], fmaf(S.yy, B[3 + c], S.yz * B[6 + c]));
        C[6 + c] = fmaf(S.xz, B[c], fmaf(S.yz, B[3 + c], S.zz * B[6 + c]));
    }
}
// A^-1 by cofactors (A near a rotation times a mild stretch: well conditioned)
__device__ __forceinline__ void mat_inv(const float (&A)[9], float (&Ai)[9]) {
    float cof[9]; mat_cof(A, cof);
    const float det = fmaf(A[0], cof[0], fmaf(A[1], cof[1], A[2] * cof[2]));
    const float id = 1.0f / det;
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) Ai[3 * r + c] = cof[3 * c + r] * id;         // inverse = cof^T / det
}
// sand stress, small strain:  A = vol * tau * (FE Fhat^-1)^T,  tau = 2 mu H + lambda tr(H) I            HybridSolver.cpp:326-339
__device__ __forceinline__ void sand_stress_small(const MatParams& mp, const Sym3& E, float e2, const float (&Fh)[9], const float (&FE)[9], float vol, float (&A)[9]) {
    const Sym3 H = sym_half_log1p(E, e2);
    const float tr = H.xx + H.yy + H.zz;
    const Sym3 tau = sym_axpi(2.0f * mp.mu0 * vol, H, mp.lambda0 * tr * vol);
    float Fi[9], T[9], Tt[9];
    mat_inv(Fh, Fi); mat_mul(FE, Fi, T);
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) Tt[3 * r + c] = T[3 * c + r];
    sym_mat_mul(tau, Tt, A);
}
// Drucker-Prager projection, small strain (HybridSolver.cpp:646-673): true when the particle yields; then F_E' = M Fhat
__device__ __forceinline__ bool sand_project_small(const MatParams& mp, const Sym3& E, float e2, float& q, Sym3& M) {
    const float PI_F = 3.14159265358979323846f;
    const Sym3 H = sym_half_log1p(E, e2);
    const float phi = (mp.h0 + (mp.h1 * q - mp.h3) * expf(-mp.h2 * q)) * (PI_F / 180.0f);      // HybridSolver.cpp:646-647
    const float sp = sinf(phi);
    const float alpha = 0.81649658092772603f * 2.0f * sp / (3.0f - sp);                          // sqrt(2/3), :649-650
    const float tr = H.xx + H.yy + H.zz;
    const Sym3 dev = sym_axpi(1.0f, H, -tr * (1.0f / 3.0f));
    const float dn = sqrtf(sym_norm2(dev));
    const float dg = dn + mp.k_vol * tr * alpha;                                                 // :654-656
    if (dg <= 0.0f) return false;
    if (dn == 0.0f || tr > 0.0f) {                                                               // :662-666  ln s' = 0
        const float h2 = sym_norm2(H);
        q += sqrtf(h2);
        M = sym_exp(sym_axpi(-1.0f, H, 0.0f), h2);
    } else {                                                                                     // :667-673  ln s' = ln s - (dg/dn) dev
        q += dg;
        M = sym_exp(sym_axpi(-dg / dn, dev, 0.0f), dg * dg);                                  // |k dev H|_F = dg
    }
    return true;
}
// F_E' = M Fhat,  F_P' = F_E'^-1 Fhat F_P                                                        HybridSolver.cpp:618-619, 675-677
__device__ __forceinline__ void sand_apply_small(const Sym3& M, const float (&Fh)[9], float (&FE)[9], float (&FP)[9]) {
    float Ftot[9], Fi[9];
    mat_mul(Fh, FP, Ftot);
    sym_mat_mul(M, Fh, FE);
    mat_inv(FE, Fi);
    mat_mul(Fi, Ftot, FP);
}

// First Piola stress times F_E^T times volume:  A = V_p * P(Fhat) * FE^T       HybridSolver.cpp:314-339
__device__ __forceinline__ void stress_times_FEt_svd(const MatParams& mp, const float (&Fh)[9], const float (&FE)[9],
                                                     float vol, float Jp, float (&A)[9]) {
    Svd3 sv; svd3(Fh, sv);
    float P[9];
    if (mp.material == 0) {
        // snow, fixed corotated with hardening e^{xi (1 - Jp)}:  P = 2 mu (F - R) + lambda (J - 1) J F^{-T}
        const float hard = expf(mp.xi * (1.0f - Jp));
        const float mu = mp.mu0 * hard, la = mp.lambda0 * hard;
        // F - R = U (S - I) V^T, evaluated in that form (no cancellation between O(1) matrices)
        const float d[3] = { 2.0f * mu * (sv.S[0] - 1.0f), 2.0f * mu * (sv.S[1] - 1.0f), 2.0f * mu * (sv.S[2] - 1.0f) };
        usvt(sv.U, d, sv.V, P);
        float cof[9]; mat_cof(Fh, cof);
        const float J = mat_det(Fh);
        const float k = la * (J - 1.0f);
#pragma unroll
        for (int i = 0; i < 9; ++i) P[i] = fmaf(k, cof[i], P[i]);
    } else {
        // sand, Hencky strain:  P = U diag(2 mu ln s / s + lambda tr(ln s) / s) V^T
        const float l0 = hencky(sv.S[0]), l1 = hencky(sv.S[1]), l2 = hencky(sv.S[2]);
        const float tr = l0 + l1 + l2;
        const float d[3] = { (2.0f * mp.mu0 * l0 + mp.lambda0 * tr) / sv.S[0], (2.0f * mp.mu0 * l1 + mp.lambda0 * tr) / sv.S[1],
                             (2.0f * mp.mu0 * l2 + mp.lambda0 * tr) / sv.S[2] };
        usvt(sv.U, d, sv.V, P);
    }
    float T[9]; mat_mul_nt(P, FE, T);
#pragma unroll
    for (int i = 0; i < 9; ++i) A[i] = vol * T[i];
}
// sand at small strain: no SVD (see above); snow and large strains: the SVD path
__device__ __forceinline__ void stress_times_FEt(const MatParams& mp, const float (&Fh)[9], const float (&FE)[9],
                                                 float vol, float Jp, float (&A)[9]) {
    Sym3 E; float e2 = 1.0f;
    if (mp.material != 0 && (e2 = left_cauchy_green_minus_one(Fh, E)) < AEP_SMALL_E2) sand_stress_small(mp, E, e2, Fh, FE, vol, A);
    else stress_times_FEt_svd(mp, Fh, FE, vol, Jp, A);
}

// Plastic return mapping on the candidate Fhat, in two parts so that the caller touches F_P only for particles that yield:
//   return_map_project   SVD of Fhat, projection of the singular values (snow clamp / Drucker-Prager), hardening state q;
//                        returns whether any singular value changed                                 HybridSolver.cpp:612-673
//   return_map_apply     FE = U S' V^T,  FP = V S'^-1 U^T Fhat FP                                   HybridSolver.cpp:618-619, 675-677
// When the projection leaves the singular values untouched the reference's U S V^T / V S^-1 U^T Ftot round trip
// is the identity in exact arithmetic, so FE = Fhat and FP is left alone (closest to the fp64 reference).
__device__ __forceinline__ bool return_map_project(const MatParams& mp, const float (&Fh)[9], Svd3& sv, float (&sn)[3], float& q) {
    svd3(Fh, sv);
    sn[0] = sv.S[0]; sn[1] = sv.S[1]; sn[2] = sv.S[2];
    bool changed = false;
    if (mp.material == 0) {
#pragma unroll
        for (int i = 0; i < 3; ++i) {                                                // HybridSolver.cpp:626-631
            const float c = fminf(fmaxf(sn[i], mp.lo), mp.hi);
            changed |= (c != sn[i]); sn[i] = c;
        }
    } else {
        const float PI_F = 3.14159265358979323846f;
        const float phi = (mp.h0 + (mp.h1 * q - mp.h3) * expf(-mp.h2 * q)) * (PI_F / 180.0f);      // HybridSolver.cpp:646-647
        const float sp = sinf(phi);
        const float alpha = 0.81649658092772603f * 2.0f * sp / (3.0f - sp);                          // sqrt(2/3), :649-650
        const float l0 = hencky(sn[0]), l1 = hencky(sn[1]), l2 = hencky(sn[2]);
        const float tr = l0 + l1 + l2;
        const float m3 = tr * (1.0f / 3.0f);
        const float d0 = l0 - m3, d1 = l1 - m3, d2 = l2 - m3;
        const float dn = sqrtf(fmaf(d0, d0, fmaf(d1, d1, d2 * d2)));
        const float dg = dn + mp.k_vol * tr * alpha;                                                 // :654-656
        if (dg <= 0.0f) {
        } else if (dn == 0.0f || tr > 0.0f) {                                                        // :662-666
            q += sqrtf(fmaf(l0, l0, fmaf(l1, l1, l2 * l2)));
            sn[0] = sn[1] = sn[2] = 1.0f; changed = true;
        } else {                                                                                     // :667-673
            const float k = dg / dn;
            sn[0] = expf(l0 - k * d0); sn[1] = expf(l1 - k * d1); sn[2] = expf(l2 - k * d2);
            q += dg; changed = true;
        }
    }
    return changed;
}
__device__ __forceinline__ void return_map_apply(const Svd3& sv, const float (&sn)[3], const float (&Fh)[9], float (&FE)[9], float (&FP)[9]) {
    float Ftot[9]; mat_mul(Fh, FP, Ftot);                                            // :618-619
    usvt(sv.U, sn, sv.V, FE);                                                        // :675
    const float inv[3] = { 1.0f / sn[0], 1.0f / sn[1], 1.0f / sn[2] };
    float Mi[9]; usvt(sv.V, inv, sv.U, Mi);                                          // V S^-1 U^T
    mat_mul(Mi, Ftot, FP);                                                           // :676-677
}
// both parts (in: Fh, FP, q; out: FE, FP, q)
__device__ __forceinline__ void return_map(const MatParams& mp, const float (&Fh)[9], float (&FE)[9], float (&FP)[9], float& q) {
    Sym3 E; float e2 = 1.0f;
    if (mp.material != 0 && (e2 = left_cauchy_green_minus_one(Fh, E)) < AEP_SMALL_E2) {
        Sym3 M;
        if (sand_project_small(mp, E, e2, q, M)) sand_apply_small(M, Fh, FE, FP);
        else {
#pragma unroll
            for (int i = 0; i < 9; ++i) FE[i] = Fh[i];
        }
        return;
    }
    Svd3 sv; float sn[3];
    if (!return_map_project(mp, Fh, sv, sn, q)) {
#pragma unroll
        for (int i = 0; i < 9; ++i) FE[i] = Fh[i];
        return;
    }
    return_map_apply(sv, sn, Fh, FE, FP);
}

// ------------------------------------------------------------------------------------------------ QR (cloth)
// classical Gram-Schmidt on columns (d1 d2 d3), geometry.cpp:31-62.  Q, R row-major.
__device__ __forceinline__ void gram_schmidt(const float (&d1)[3], const float (&d2)[3], const float (&d3)[3],
                                             float (&Q)[9], float (&R)[9]) {
    const float r11 = sqrtf(d1[0] * d1[0] + d1[1] * d1[1] + d1[2] * d1[2]);
    const float i11 = 1.0f / r11;
    const float q1[3] = { d1[0] * i11, d1[1] * i11, d1[2] * i11 };
    const float r12 = d2[0] * q1[0] + d2[1] * q1[1] + d2[2] * q1[2];
    float q2[3] = { d2[0] - r12 * q1[0], d2[1] - r12 * q1[1], d2[2] - r12 * q1[2] };
    const float r22 = sqrtf(q2[0] * q2[0] + q2[1] * q2[1] + q2[2] * q2[2]);
    const float i22 = 1.0f / r22; q2[0] *= i22; q2[1] *= i22; q2[2] *= i22;
    const float r13 = d3[0] * q1[0] + d3[1] * q1[1] + d3[2] * q1[2];
    const float r23 = d3[0] * q2[0] + d3[1] * q2[1] + d3[2] * q2[2];
    float q3[3] = { d3[0] - r13 * q1[0] - r23 * q2[0], d3[1] - r13 * q1[1] - r23 * q2[1], d3[2] - r13 * q1[2] - r23 * q2[2] };
    const float r33 = sqrtf(q3[0] * q3[0] + q3[1] * q3[1] + q3[2] * q3[2]);
    const float i33 = 1.0f / r33; q3[0] *= i33; q3[1] *= i33; q3[2] *= i33;
#pragma unroll
    for (int r = 0; r < 3; ++r) { Q[3 * r] = q1[r]; Q[3 * r + 1] = q2[r]; Q[3 * r + 2] = q3[r]; }
    R[0] = r11; R[1] = r12; R[2] = r13; R[3] = 0.f; R[4] = r22; R[5] = r23; R[6] = 0.f; R[7] = 0.f; R[8] = r33;
}

}  // namespace aep
