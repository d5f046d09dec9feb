// tests/cpu_math_harness.cpp -- TEST-ONLY: compiles the device math header (csrc/aep_math.cuh) for the host so that the
// fp32 B-spline / SVD / stress / return-mapping code can be checked against the fp64 oracle without a GPU.
// Not part of the product; nothing in the package loads it.
#include <cmath>
#include <cstring>
#define AEP_HOST_MATH_TEST
#define __device__
#define __host__
#define __forceinline__ inline
static inline float rsqrtf(float x) { return 1.0f / std::sqrt(x); }
static inline float __fdividef(float a, float b) { return a / b; }
// the few CUDA vector types the scatter header uses
#define __restrict__
struct float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(16) ulonglong2 { unsigned long long x, y; };
static inline float4 make_float4(float x, float y, float z, float w) { float4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
using std::fmaf;
#include "../anisotropicelastoplasticity_b200/csrc/aep_math.cuh"
#include "../anisotropicelastoplasticity_b200/csrc/aep_scatter.cuh"
#include "../anisotropicelastoplasticity_b200/csrc/aep_gather.cuh"
#include <vector>

using namespace aep;
extern "C" {
void h_bspline4(float f, float* N, float* D) { float n[4], d[4]; bspline4(f, n, d); std::memcpy(N, n, 16); std::memcpy(D, d, 16); }
void h_bspline_lane(float f, int o, float* N, float* D) { bspline_lane(f, o, *N, *D); }
void h_svd3(const float* F, float* U, float* S, float* V) {
    float f[9]; std::memcpy(f, F, 36); Svd3 sv; svd3(f, sv);
    std::memcpy(U, sv.U, 36); std::memcpy(S, sv.S, 12); std::memcpy(V, sv.V, 36);
}
static MatParams mk(int material, double E, double nu, double thetaC, double thetaS) {
    MatParams M; const double la = E * nu / (1.0 + nu) / (1.0 - 2.0 * nu), mu = E / 2.0 / (1.0 + nu);
    M.lambda0 = (float)la; M.mu0 = (float)mu; M.xi = 10.f; M.lo = (float)(1.0 - thetaC); M.hi = (float)(1.0 + thetaS);
    M.h0 = 35.f; M.h1 = 9.f; M.h2 = 0.2f; M.h3 = 10.f; M.k_vol = (float)((3.0 * la + 2.0 * mu) / 2.0 / mu); M.material = material;
    return M;
}
// row-major float[9] in/out
void h_stress(int material, double E, double nu, const float* Fh, const float* FE, float vol, float Jp, float* A) {
    MatParams M = mk(material, E, nu, 2.5e-2, 7.5e-3);
    float fh[9], fe[9], a[9]; std::memcpy(fh, Fh, 36); std::memcpy(fe, FE, 36);
    stress_times_FEt(M, fh, fe, vol, Jp, a); std::memcpy(A, a, 36);
}
// the SVD path alone (what round 1 shipped): the series path must be at least as close to the fp64 reference
void h_stress_svd(int material, double E, double nu, const float* Fh, const float* FE, float vol, float Jp, float* A) {
    MatParams M = mk(material, E, nu, 2.5e-2, 7.5e-3);
    float fh[9], fe[9], a[9]; std::memcpy(fh, Fh, 36); std::memcpy(fe, FE, 36);
    stress_times_FEt_svd(M, fh, fe, vol, Jp, a); std::memcpy(A, a, 36);
}
// matrix functions of the series path: in/out symmetric 3x3 as (xx, yy, zz, xy, xz, yz)
void h_sym_half_log1p(const float* e, float* out) { Sym3 E{e[0], e[1], e[2], e[3], e[4], e[5]}; const Sym3 H = sym_half_log1p(E, sym_norm2(E)); out[0] = H.xx; out[1] = H.yy; out[2] = H.zz; out[3] = H.xy; out[4] = H.xz; out[5] = H.yz; }
void h_sym_exp(const float* x, float* out) { Sym3 X{x[0], x[1], x[2], x[3], x[4], x[5]}; const Sym3 R = sym_exp(X, sym_norm2(X)); out[0] = R.xx; out[1] = R.yy; out[2] = R.zz; out[3] = R.xy; out[4] = R.xz; out[5] = R.yz; }
void h_return_map_svd(int material, double E, double nu, double thetaC, double thetaS, const float* Fh, float* FE, float* FP, float* q) {
    MatParams M = mk(material, E, nu, thetaC, thetaS);
    float fh[9], fe[9], fp[9]; std::memcpy(fh, Fh, 36); std::memcpy(fp, FP, 36);
    Svd3 sv; float sn[3];
    if (return_map_project(M, fh, sv, sn, *q)) return_map_apply(sv, sn, fh, fe, fp); else std::memcpy(fe, fh, 36);
    std::memcpy(FE, fe, 36); std::memcpy(FP, fp, 36);
}
int h_small_strain(const float* Fh) { float fh[9]; std::memcpy(fh, Fh, 36); Sym3 E; return left_cauchy_green_minus_one(fh, E) < AEP_SMALL_E2 ? 1 : 0; }
void h_return_map(int material, double E, double nu, double thetaC, double thetaS, const float* Fh, float* FE, float* FP, float* q) {
    MatParams M = mk(material, E, nu, thetaC, thetaS);
    float fh[9], fe[9], fp[9]; std::memcpy(fh, Fh, 36); std::memcpy(fp, FP, 36);
    return_map(M, fh, fe, fp, *q); std::memcpy(FE, fe, 36); std::memcpy(FP, fp, 36);
}
long h_svd_sweeps() { return g_svd_sweeps; }
void h_gram_schmidt(const float* d1, const float* d2, const float* d3, float* Q, float* R) {
    float a[3], b[3], c[3], q[9], r[9]; std::memcpy(a, d1, 12); std::memcpy(b, d2, 12); std::memcpy(c, d3, 12);
    gram_schmidt(a, b, c, q, r); std::memcpy(Q, q, 36); std::memcpy(R, r, 36);
}
// ---- the two scatters, exactly as the kernels run them: phase A record, then the 16 (j,k) lanes of phase B
// out[((k*4 + j)*4 + i)*4 + c] = contribution of the particle to node offset (i,j,k), c = (m, px, py, pz) resp. (fx, fy, fz, -)
void h_p2g_scatter(const float* f, const float* v, const float* B, float m, const float* h, float* out) {
    const float hmin = std::fmin(h[0], std::fmin(h[1], h[2]));
    const float apic = 3.0f / hmin / hmin;                                   // HybridSolver.cpp:175-177
    float4 rec[P2G_STRIDE];
    p2g_make_record(rec, make_float4(f[0], f[1], f[2], 0.f), make_float4(v[0], v[1], v[2], m), make_float4(B[0], B[1], B[2], 0.f),
                    make_float4(B[3], B[4], B[5], 0.f), make_float4(B[6], B[7], B[8], 0.f), m, apic, h[0], h[1], h[2], 0.f);
    for (int k = 0; k < 4; ++k)
        for (int j = 0; j < 4; ++j) {
            AccRow acc; acc_zero(acc);
            p2g_row_accumulate(rec, 16 + 4 * j, 32 + 4 * k, pk1((float)j), pk1((float)k), acc);
            for (int i = 0; i < 4; ++i) { const float4 q = quad(acc.lo[i], acc.hi[i]); std::memcpy(out + ((k * 4 + j) * 4 + i) * 4, &q, 16); }
        }
}
void h_frc_scatter(const float* f, const float* A, const float* h, float* out) {
    float N[3][4], D[3][4];
    for (int a = 0; a < 3; ++a) { bspline4(f[a], N[a], D[a]); for (int o = 0; o < 4; ++o) D[a][o] *= 1.0f / h[a]; }    // axis_setup scales D by 1/h
    float a9[9]; std::memcpy(a9, A, 36);
    float4 rec[FRC_STRIDE];
    frc_make_record(rec, N[0], D[0], N[1], D[1], N[2], D[2], a9, 0.f);
    for (int k = 0; k < 4; ++k)
        for (int j = 0; j < 4; ++j) {
            AccRow acc; acc_zero(acc);
            frc_row_accumulate(rec, 32 + 8 * j, 64 + 8 * k, acc);
            for (int i = 0; i < 4; ++i) { const float4 q = quad(acc.lo[i], acc.hi[i]); std::memcpy(out + ((k * 4 + j) * 4 + i) * 4, &q, 16); }
        }
}
// ---- the two gathers, exactly as the kernels run them.  vt: nx*ny*nz nodes x (v~x, v~y, v~z, s), node index (k*ny + j)*nx + i.
// use_tile = 1: the nodes are first staged like the TMA box of hw_tile_issue (zeros outside the grid), MODE 2;
// use_tile = 0: clamped loads from the grid, MODE 0 (stencils cut by a domain face).
static GridP small_grid(int nx, int ny, int nz, const float* vt, const float* h) {
    GridP G{}; G.nx = nx; G.ny = ny; G.nz = nz; G.hx = h[0]; G.hy = h[1]; G.hz = h[2];
    G.ihx = 1.0f / h[0]; G.ihy = 1.0f / h[1]; G.ihz = 1.0f / h[2];
    G.vt = reinterpret_cast<float4*>(const_cast<float*>(vt));
    G.sy = nx; G.sz = (long long)nx * ny; G.a1[0] = G.v1[0] = nx; G.a1[1] = G.v1[1] = ny; G.a1[2] = G.v1[2] = nz;
    return G;
}
static int fill_tile(const GridP& G, const int* cell, std::vector<float4>& tile) {      // returns xoff of the particle's first node
    const int ox0 = cell[0] - 1 - TILE_SLACK, j0 = cell[1] - 1, k0 = cell[2] - 1;
    tile.assign(TILE_F4, make_float4(1e30f, 1e30f, 1e30f, 1.0f));                       // poison: entries outside the box must not be read
    for (int idx = 0; idx < TILE_F4; ++idx) {
        const int r = idx / TILE_W, x = idx - r * TILE_W, gx = ox0 + x;
        const int gy = j0 + (r & 3), gz = k0 + (r >> 2);
        if (gx >= 0 && gx < G.nx && gy >= 0 && gy < G.ny && gz >= 0 && gz < G.nz) tile[idx] = G.vt[nidx(G, gx, gy, gz)];
        else tile[idx] = make_float4(0.f, 0.f, 0.f, 0.f);                                  // TMA zero fill outside the tensor
    }
    return (cell[0] - 1) - ox0;
}
void h_gather_grad(int nx, int ny, int nz, const float* vt, const int* cell, const float* f, const float* h, int use_tile, float* g9) {
    const GridP G = small_grid(nx, ny, nz, vt, h);
    Axis ax, ay, az;
    axis_setup(ax, f[0], cell[0], nx, G.ihx); axis_setup(ay, f[1], cell[1], ny, G.ihy); axis_setup(az, f[2], cell[2], nz, G.ihz);
    float g[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    std::vector<float4> tile;
    if (use_tile) { const int xoff = fill_tile(G, cell, tile); gather_grad<2>(G, ax, ay, az, tile.data(), xoff, g); }
    else gather_grad<0>(G, ax, ay, az, nullptr, 0, g);
    std::memcpy(g9, g, 36);
}
void h_g2p_gather(int nx, int ny, int nz, const float* vt, const int* cell, const float* f, const float* h, int use_tile,
                  float* va, float* vc, float* B, float* g, float* smin) {
    const GridP G = small_grid(nx, ny, nz, vt, h);
    Axis ax, ay, az;
    axis_setup(ax, f[0], cell[0], nx, G.ihx); axis_setup(ay, f[1], cell[1], ny, G.ihy); axis_setup(az, f[2], cell[2], nz, G.ihz);
    float rx[4], ry[4], rz[4], nrx[4];                                                  // as in k_g2p: x_i - x_p per axis
    for (int o = 0; o < 4; ++o) { rx[o] = G.hx * ((float)(o - 1) - f[0]); ry[o] = G.hy * ((float)(o - 1) - f[1]); rz[o] = G.hz * ((float)(o - 1) - f[2]); }
    for (int o = 0; o < 4; ++o) nrx[o] = ax.N[o] * rx[o];
    G2PSums S;
    for (int i = 0; i < 3; ++i) { S.vc[i] = 0.f; S.va[i] = 0.f; }
    for (int i = 0; i < 9; ++i) { S.B[i] = 0.f; S.g[i] = 0.f; }
    S.smin = 1.0f;
    std::vector<float4> tile;
    if (use_tile) {
        const int xoff = fill_tile(G, cell, tile);
        g2p_gather<2>(G, ax, ay, az, nrx, rx, ry, rz, tile.data(), xoff, S);
        if (S.smin < 1.0f) g2p_stick_correction<2>(G, ax, ay, az, rx, ry, rz, tile.data(), xoff, S);
    } else {
        g2p_gather<0>(G, ax, ay, az, nrx, rx, ry, rz, nullptr, 0, S);
        if (S.smin < 1.0f) g2p_stick_correction<0>(G, ax, ay, az, rx, ry, rz, nullptr, 0, S);
    }
    std::memcpy(va, S.va, 12); std::memcpy(vc, S.vc, 12); std::memcpy(B, S.B, 36); std::memcpy(g, S.g, 36); *smin = S.smin;
}
}
