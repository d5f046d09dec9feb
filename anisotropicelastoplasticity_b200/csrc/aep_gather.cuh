// aep_gather.cuh -- the arithmetic of the two 64-node gathers (grad v for the forces; v, B, grad v for G2P), free of thread-index
// and launch details so that tests/cpu_math_harness.cpp can run exactly this code on the host against the direct formulas of the
// reference (HybridSolver.cpp:269-301, 739-825).  MODE 2 reads a warp's shared-memory tile, MODE 0 the grid itself.
#pragma once
#include "aep_math.cuh"
#include "aep_pack.cuh"

// -DAEP_GATHER_PK=1 (development, not the default build): the inner loops of both gathers in packed fp32x2 form.  The float4 of
// a node arrives from LDS.128 / LDG.128 as two aligned register pairs (x,y) and (z,s), so a_xy += (x,y) * N and a_zs += (z,s) * N
// are two FFMA2 instead of three FFMA (the s lane rides along unused).  Every lane performs exactly the scalar sequence of
// fmaf's, so the results are bit-identical to the default form (tests/test_device_math_host.py checks that on the host).
#ifndef AEP_GATHER_PK
#define AEP_GATHER_PK 0
#endif
#ifndef AEP_GATHER_PK_GRAD              // the force gather (gather_grad) and the G2P gather (g2p_gather) separately
#define AEP_GATHER_PK_GRAD AEP_GATHER_PK
#endif
#ifndef AEP_GATHER_PK_G2P
#define AEP_GATHER_PK_G2P AEP_GATHER_PK
#endif

namespace aep {

struct GridP {
    int nx, ny, nz;
    int nbx, nby, nbz;                 // number of 8^3 node blocks per axis
    int nqx, nqy;                      // number of 4^3 cell bricks along x, y (sort order)
    int rb0[3], rbn[3];                // 8^3-block range the grid passes run over (the whole grid, or the slab's reach)
    int bricks;                        // 1: brick-major sort keys, 0: plain cell index
    int strips;                        // scatter kernels: CTAs that run concurrently work on `strips` far-apart parts of the sorted order
    float hx, hy, hz, ihx, ihy, ihz;
    float mnx, mny, mnz;
    float apic;                        // 3 / hmin^2                       HybridSolver.cpp:175-177
    float inv_cell_vol;                // 1 / (hx hy hz)                   HybridSolver.cpp:246
    float gravity, friction;
    float cvx, cvy, cvz;               // velocity of a moving collider (0 for the reference's static ones, HS:484): a sticking node moves with it
    long long sy, sz;                  // element strides of the node index along y and z (x is contiguous): node = k*sz + j*sy + i.
                                       // Whole-grid contexts: sy = nx, sz = nx*ny (RegularGrid.cpp:164-168).  A slab context allocates only
                                       // its own node planes and makes the slab axis the slowest one; the array pointers are then biased
                                       // so that global (i,j,k) still index them.
    int a0[3], a1[3];                  // node range [a0, a1) per axis that this context's grid arrays hold (the whole grid, or the slab's reach)
    int v0[3], v1[3];                  // node range whose sums are complete after the halo exchange: the grid passes compute only there (a slab
                                       // also holds one scratch plane per side that only catches the scatter of particles about to migrate)
    float4* mp;
    float4* f;
    float4* vt;
    unsigned char* flags;
    const unsigned char* ls_code;      // 0 outside, 1..6 axis normals (+x -x +y -y +z -z), 7 general (ls_nrm)
    const float4* ls_nrm;
};

__host__ __device__ __forceinline__ size_t nidx(const GridP& G, int i, int j, int k) { return (size_t)((long long)k * G.sz + (long long)j * G.sy + i); }

#ifdef AEP_HOST_MATH_TEST
__device__ __forceinline__ float4 ldg4(const float4* p) { return *p; }
#else
__device__ __forceinline__ float4 ldg4(const float4* p) { return __ldg(p); }
#endif
// a node's float4 as the two register pairs (x,y), (z,w)
template <int MODE> __device__ __forceinline__ ulonglong2 node_pairs(const float4* p) {
#ifdef AEP_HOST_MATH_TEST
    return ld_pairs(p);
#else
    return MODE == 2 ? ld_pairs(p) : __ldg(reinterpret_cast<const ulonglong2*>(p));
#endif
}
// per-axis stencil of a thread-owned particle: weights (masked to 0 outside the grid, HybridSolver.cpp:44-46)
// and clamped node coordinates
struct Axis {
    float N[4], D[4];
    int n0;            // first node (cell - 1), may be -1
};
__device__ __forceinline__ bool axis_setup(Axis& a, float f, int cell, int nres, float ih) {
    bspline4(f, a.N, a.D);
    a.n0 = cell - 1;
    bool complete = true;
#pragma unroll
    for (int o = 0; o < 4; ++o) {
        const int n = a.n0 + o;
        const bool in = (n >= 0) && (n < nres);
        complete &= in;
        a.N[o] = in ? a.N[o] : 0.0f;
        a.D[o] = in ? a.D[o] * ih : 0.0f;
    }
    return complete;
}
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
__device__ __forceinline__ float sel4(const float (&a)[4], int k) { return k == 0 ? a[0] : (k == 1 ? a[1] : (k == 2 ? a[2] : a[3])); }

// Warps that cannot use the tile gather straight from global memory with clamped indices (MODE 0).  A second fallback with
// unclamped "interior" addressing (MODE 1) was 5 % faster on the ~1 % of warps that take it but made both gather kernels 17 %
// larger; the instruction cache matters more (no_instruction stalls, profiles/README.md v10).
#define AEP_FALLBACK_MODE 0
#define TILE_W 8                        // nodes along x held by a half-warp's tile: 16 cell-sorted particles span up to 5 cells in a row
#define TILE_SLACK 1                    // nodes left of lane 0's stencil (tolerates slightly out-of-order particles)
#define TILE_F4 (16 * TILE_W)           // float4 per half-warp tile (16 (j,k) rows): 2 KB
// g[3r+c] = sum_i v_i[r] d_c w_i over the 4x4x4 stencil, x summed first.
// MODE 0: clamped global loads (fallback: stencil cut by a domain face, warp not in one row of cells), MODE 2: loads from the warp's
// shared tile (xoff = first stencil node relative to the tile).
template <int MODE>
__device__ __forceinline__ void gather_grad(const GridP& G, const Axis& ax, const Axis& ay, const Axis& az, const float4* __restrict__ tile, int xoff,
                                            float (&g)[9]) {
    int ni[4], nj[4];
#pragma unroll
    for (int o = 0; o < 4; ++o) { ni[o] = MODE == 2 ? o : clampi(ax.n0 + o, G.a0[0], G.a1[0] - 1); nj[o] = MODE == 2 ? ay.n0 + o : clampi(ay.n0 + o, G.a0[1], G.a1[1] - 1); }
    const float4* base = MODE == 2 ? tile + xoff : G.vt;
#if AEP_GATHER_PK_GRAD
    f32x2 g03 = pk(g[0], g[3]), g14 = pk(g[1], g[4]), g25 = pk(g[2], g[5]);     // rows 0 and 1 of grad v as pairs over the row index
#endif
#pragma unroll 1
    for (int k = 0; k < 4; ++k) {
        const int nk = MODE == 2 ? az.n0 + k : clampi(az.n0 + k, G.a0[2], G.a1[2] - 1);
        const float nzk = sel4(az.N, k), dzk = sel4(az.D, k);
        const float4* plane = MODE == 2 ? base + k * 4 * TILE_W : base + (size_t)((long long)nk * G.sz);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float4* row = MODE == 2 ? plane + j * TILE_W : plane + (size_t)((long long)nj[j] * G.sy);
#if AEP_GATHER_PK_GRAD
            const float nn = ay.N[j] * nzk, dn = ay.D[j] * nzk, nd = ay.N[j] * dzk;
            f32x2 axy = 0ull, azs = 0ull, bxy = 0ull, bzs = 0ull;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const ulonglong2 q = node_pairs<MODE>(MODE == 2 ? row + i : row + ni[i]);
                const f32x2 nw = pk1(ax.N[i]), dw = pk1(ax.D[i]);
                axy = fma2(q.x, nw, axy); azs = fma2(q.y, nw, azs);
                bxy = fma2(q.x, dw, bxy); bzs = fma2(q.y, dw, bzs);
            }
            g03 = fma2(bxy, pk1(nn), g03); g14 = fma2(axy, pk1(dn), g14); g25 = fma2(axy, pk1(nd), g25);
            const float a2 = lo_of(azs), b2 = lo_of(bzs);
            g[6] = fmaf(b2, nn, g[6]); g[7] = fmaf(a2, dn, g[7]); g[8] = fmaf(a2, nd, g[8]);
#else
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, b0 = 0.f, b1 = 0.f, b2 = 0.f;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float4 v = MODE == 2 ? row[i] : ldg4(row + ni[i]);
                a0 = fmaf(v.x, ax.N[i], a0); a1 = fmaf(v.y, ax.N[i], a1); a2 = fmaf(v.z, ax.N[i], a2);
                b0 = fmaf(v.x, ax.D[i], b0); b1 = fmaf(v.y, ax.D[i], b1); b2 = fmaf(v.z, ax.D[i], b2);
            }
            const float nn = ay.N[j] * nzk, dn = ay.D[j] * nzk, nd = ay.N[j] * dzk;
            g[0] = fmaf(b0, nn, g[0]); g[1] = fmaf(a0, dn, g[1]); g[2] = fmaf(a0, nd, g[2]);
            g[3] = fmaf(b1, nn, g[3]); g[4] = fmaf(a1, dn, g[4]); g[5] = fmaf(a1, nd, g[5]);
            g[6] = fmaf(b2, nn, g[6]); g[7] = fmaf(a2, dn, g[7]); g[8] = fmaf(a2, nd, g[8]);
#endif
        }
    }
#if AEP_GATHER_PK_GRAD
    upk(g03, g[0], g[3]); upk(g14, g[1], g[4]); upk(g25, g[2], g[5]);
#endif
}

// The 64-node gather of G2P.  Per (j,k) row the x direction is summed first over v~ only:
//   a = sum v~ Nx, b = sum v~ Dx, d = sum v~ Nx rx           (9 FMA per node)
// then the rows are combined into  va = sum w v~,  g = sum v~ (grad w)^T,  B~ = sum w v~ (x_i - x_p)^T.
// The post-friction velocity is v = s v~ with s in {0,1} and s = 0 only on sticking collider nodes (k_grid_update), so
//   v_p = va - sum_{s=0} w v~   and   B = B~ - sum_{s=0} w v~ (x_i - x_p)^T;
// the correction branch runs only for rows that contain a sticking node.
struct G2PSums {
    float va[3], vc[3], B[9], g[9];
    float smin;                          // min of the s factors seen: < 1 iff the stencil holds a node slowed by friction (0: sticking)
};
template <int MODE>
__device__ __forceinline__ void g2p_gather(const GridP& G, const Axis& ax, const Axis& ay, const Axis& az, const float (&nrx)[4],
                                           const float (&rx)[4], const float (&ry)[4], const float (&rz)[4], const float4* __restrict__ tile, int xoff,
                                           G2PSums& S) {
    int ni[4], nj[4];
#pragma unroll
    for (int o = 0; o < 4; ++o) { ni[o] = MODE == 2 ? o : clampi(ax.n0 + o, G.a0[0], G.a1[0] - 1); nj[o] = MODE == 2 ? ay.n0 + o : clampi(ay.n0 + o, G.a0[1], G.a1[1] - 1); }
    const float4* base = MODE == 2 ? tile + xoff : G.vt;
#if AEP_GATHER_PK_G2P
    f32x2 va01 = pk(S.va[0], S.va[1]);
    f32x2 g03 = pk(S.g[0], S.g[3]), g14 = pk(S.g[1], S.g[4]), g25 = pk(S.g[2], S.g[5]);
    f32x2 B03 = pk(S.B[0], S.B[3]), B14 = pk(S.B[1], S.B[4]), B25 = pk(S.B[2], S.B[5]);
#endif
#pragma unroll 1
    for (int k = 0; k < 4; ++k) {
        const int nk = MODE == 2 ? az.n0 + k : clampi(az.n0 + k, G.a0[2], G.a1[2] - 1);
        const float nzk = sel4(az.N, k), dzk = sel4(az.D, k), rzk = sel4(rz, k);
        const float4* plane = MODE == 2 ? base + k * 4 * TILE_W : base + (size_t)((long long)nk * G.sz);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float4* row = MODE == 2 ? plane + j * TILE_W : plane + (size_t)((long long)nj[j] * G.sy);
#if AEP_GATHER_PK_G2P
            const float nn = ay.N[j] * nzk, dn = ay.D[j] * nzk, nd = ay.N[j] * dzk;
            ulonglong2 q[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) q[i] = node_pairs<MODE>(MODE == 2 ? row + i : row + ni[i]);
            f32x2 axy = 0ull, azs = 0ull, bxy = 0ull, bzs = 0ull, dxy = 0ull, dzs = 0ull;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const f32x2 nw = pk1(ax.N[i]), dw = pk1(ax.D[i]), rw = pk1(nrx[i]);
                axy = fma2(q[i].x, nw, axy); azs = fma2(q[i].y, nw, azs);
                bxy = fma2(q[i].x, dw, bxy); bzs = fma2(q[i].y, dw, bzs);
                dxy = fma2(q[i].x, rw, dxy); dzs = fma2(q[i].y, rw, dzs);
            }
            const f32x2 nn2 = pk1(nn);
            const f32x2 u01 = mul2(axy, nn2);                                   // sum_i w v~_i over the row, components x and y
            const float a2 = lo_of(azs), b2 = lo_of(bzs), d2 = lo_of(dzs), u2 = a2 * nn;
            va01 = add2(va01, u01); S.va[2] += u2;
            g03 = fma2(bxy, nn2, g03); g14 = fma2(axy, pk1(dn), g14); g25 = fma2(axy, pk1(nd), g25);
            S.g[6] = fmaf(b2, nn, S.g[6]); S.g[7] = fmaf(a2, dn, S.g[7]); S.g[8] = fmaf(a2, nd, S.g[8]);
            B03 = fma2(dxy, nn2, B03); B14 = fma2(u01, pk1(ry[j]), B14); B25 = fma2(u01, pk1(rzk), B25);
            S.B[6] = fmaf(d2, nn, S.B[6]); S.B[7] = fmaf(u2, ry[j], S.B[7]); S.B[8] = fmaf(u2, rzk, S.B[8]);
            S.smin = fminf(fminf(S.smin, fminf(hi_of(q[0].y), hi_of(q[1].y))), fminf(hi_of(q[2].y), hi_of(q[3].y)));
#else
            float4 t[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) t[i] = MODE == 2 ? row[i] : ldg4(row + ni[i]);
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, b0 = 0.f, b1 = 0.f, b2 = 0.f, d0 = 0.f, d1 = 0.f, d2 = 0.f;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                a0 = fmaf(t[i].x, ax.N[i], a0); a1 = fmaf(t[i].y, ax.N[i], a1); a2 = fmaf(t[i].z, ax.N[i], a2);
                b0 = fmaf(t[i].x, ax.D[i], b0); b1 = fmaf(t[i].y, ax.D[i], b1); b2 = fmaf(t[i].z, ax.D[i], b2);
                d0 = fmaf(t[i].x, nrx[i], d0); d1 = fmaf(t[i].y, nrx[i], d1); d2 = fmaf(t[i].z, nrx[i], d2);
            }
            const float nn = ay.N[j] * nzk, dn = ay.D[j] * nzk, nd = ay.N[j] * dzk;
            const float u0 = a0 * nn, u1 = a1 * nn, u2 = a2 * nn;            // sum_i w v~_i over the row
            S.va[0] += u0; S.va[1] += u1; S.va[2] += u2;
            S.g[0] = fmaf(b0, nn, S.g[0]); S.g[1] = fmaf(a0, dn, S.g[1]); S.g[2] = fmaf(a0, nd, S.g[2]);
            S.g[3] = fmaf(b1, nn, S.g[3]); S.g[4] = fmaf(a1, dn, S.g[4]); S.g[5] = fmaf(a1, nd, S.g[5]);
            S.g[6] = fmaf(b2, nn, S.g[6]); S.g[7] = fmaf(a2, dn, S.g[7]); S.g[8] = fmaf(a2, nd, S.g[8]);
            S.B[0] = fmaf(d0, nn, S.B[0]); S.B[1] = fmaf(u0, ry[j], S.B[1]); S.B[2] = fmaf(u0, rzk, S.B[2]);
            S.B[3] = fmaf(d1, nn, S.B[3]); S.B[4] = fmaf(u1, ry[j], S.B[4]); S.B[5] = fmaf(u1, rzk, S.B[5]);
            S.B[6] = fmaf(d2, nn, S.B[6]); S.B[7] = fmaf(u2, ry[j], S.B[7]); S.B[8] = fmaf(u2, rzk, S.B[8]);
            S.smin = fminf(fminf(S.smin, fminf(t[0].w, t[1].w)), fminf(t[2].w, t[3].w));   // 0 iff a sticking node was seen (FMNMX: ALU pipe)
#endif
        }
    }
#if AEP_GATHER_PK_G2P
    upk(va01, S.va[0], S.va[1]);
    upk(g03, S.g[0], S.g[3]); upk(g14, S.g[1], S.g[4]); upk(g25, S.g[2], S.g[5]);
    upk(B03, S.B[0], S.B[3]); upk(B14, S.B[1], S.B[4]); upk(B25, S.B[2], S.B[5]);
#endif
}
// second pass, taken only by particles whose stencil holds a sticking node (next to the collider): subtract w v~ of those nodes
// from v_p and B.  Kept out of the gather loop and rolled up: the hot loop stays branch-free and half as long.
template <int MODE>
__device__ __forceinline__ void g2p_stick_correction(const GridP& G, const Axis& ax, const Axis& ay, const Axis& az, const float (&rx)[4],
                                                  const float (&ry)[4], const float (&rz)[4], const float4* __restrict__ tile, int xoff, G2PSums& S) {
#pragma unroll 1
    for (int n = 0; n < 64; ++n) {
        const int i = n & 3, j = (n >> 2) & 3, k = n >> 4;
        float4 t;
        if (MODE == 2) t = tile[xoff + (k * 4 + j) * TILE_W + i];
        else t = ldg4(G.vt + nidx(G, clampi(ax.n0 + i, G.a0[0], G.a1[0] - 1), clampi(ay.n0 + j, G.a0[1], G.a1[1] - 1), clampi(az.n0 + k, G.a0[2], G.a1[2] - 1)));
        if (t.w == 1.0f) continue;
        // v = c + s (v~ - c), c = the collider's velocity (0: the reference's static colliders): s = 0 on sticking nodes (HS:494-502);
        // 0 < s < 1 only with the opt-in Coulomb friction
        const float w = (t.w - 1.0f) * sel4(ax.N, i) * sel4(ay.N, j) * sel4(az.N, k);
        const float cx = w * (t.x - G.cvx), cy = w * (t.y - G.cvy), cz = w * (t.z - G.cvz);
        const float rxi = sel4(rx, i), ryj = sel4(ry, j), rzk = sel4(rz, k);
        S.vc[0] += cx; S.vc[1] += cy; S.vc[2] += cz;
        S.B[0] = fmaf(cx, rxi, S.B[0]); S.B[1] = fmaf(cx, ryj, S.B[1]); S.B[2] = fmaf(cx, rzk, S.B[2]);
        S.B[3] = fmaf(cy, rxi, S.B[3]); S.B[4] = fmaf(cy, ryj, S.B[4]); S.B[5] = fmaf(cy, rzk, S.B[5]);
        S.B[6] = fmaf(cz, rxi, S.B[6]); S.B[7] = fmaf(cz, ryj, S.B[7]); S.B[8] = fmaf(cz, rzk, S.B[8]);
    }
}

}  // namespace aep
