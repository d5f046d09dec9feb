// aep_mesh.cuh -- LagrangianMesh (cloth) path of the hybrid solver on the GPU.
//
// Mesh vertices and element centroids are transferred to / from the grid exactly like particles (HybridSolver.cpp:
// 121-125,137-141,216-230,748-756,918-935,946-950); the cloth constitutive model works per face:
//   in-plane fixed-corotated stress on the QR factors       LagrangianMesh.cpp:382-460
//   normal / shear stress from R's third column             HybridSolver.cpp:389-455
//   d1,d2 from advected vertices, d3 by grad v~             HybridSolver.cpp:581-608
//   cone return mapping on (r13, r23, r33)                  HybridSolver.cpp:684-722
//   pinned vertices zero 3x3x3 node blocks                  HybridSolver.cpp:513-550
// Mesh points are few (<= ~1e6) and cannot be re-sorted (topology), so they reuse the warp-cooperative scatter of the
// particle kernels with run length 1 and otherwise use one thread per point.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cmath>
#include <vector>

#include "aep_kernels.cuh"

namespace aep {

// Mesh points use the particle layout of aep_kernels.cuh for what the transfers touch: X, V = (v, B00), C0, C1 (rest of B), K = (m, -, -, -).
struct MeshDev {
    // vertices
    float4 *VX, *VV, *VC0, *VC1, *VK, *VKe, *VF;   // VF: in-plane force accumulator; VKe: K with the mass zeroed for points of other ranks
    // elements (one per face)
    float4 *EX, *EV, *EC0, *EC1, *EK, *EKe;
    float4 *ED1, *ED2, *ED3;                       // current directions; ED3.w = element volume
    float4 *RD1, *RD2, *RD3;                       // rest directions
    float4* PK;                                    // (pk00, pk01, pk11, -) = invRest * P   LagrangianMesh.cpp:446
    int4* faces;
    int* fixed_ids;
    float mu, lambda, gamma, kstiff, cf;
};
// Slab decomposition of the cloth (SURVEY 8e "Cloth"): every rank holds the whole mesh state; a rank transfers (scatters, gathers,
// advects) only the points whose cell lies in its slab and receives the others' results from their owners.  In-plane forces are
// computed redundantly by everyone (a pure function of the replicated state).  axis < 0: this context owns every point.
struct MeshOwn {
    int axis, lo, hi;
};
__device__ __forceinline__ bool mesh_owned(const MeshOwn& O, int cell) {
    if (O.axis < 0) return true;
    const int c = O.axis == 0 ? (cell & 1023) : (O.axis == 1 ? ((cell >> 10) & 1023) : ((cell >> 20) & 1023));
    return c >= O.lo && c < O.hi;
}

struct MeshState {
    long long nv = 0, nf = 0;
    int n_fixed = 0;
    MeshDev d{};
    MeshOwn own{-1, 0, 0};
    void* block = nullptr; size_t block_bytes = 0;  // every array above lives in this one allocation (peers map it with one IPC handle)
    size_t sync_v_off = 0, sync_v_bytes = 0;        // byte range of the per-vertex / per-element arrays that G2P rewrites (owner -> peers)
    size_t sync_e_off = 0, sync_e_bytes = 0;
    unsigned char *owner_v = nullptr, *owner_e = nullptr;   // 1 where this rank did the G2P of the point in the current substep
    double mn[3]{}, h[3]{};
};

// ------------------------------------------------------------------------------------------------ device helpers
// position difference b - a in world units from packed (cell, frac) records
__device__ __forceinline__ void pos_diff(const GridP& G, const float4& a, const float4& b, float (&d)[3]) {
    const int ca = __float_as_int(a.w), cb = __float_as_int(b.w);
    d[0] = ((float)(cell_i(cb) - cell_i(ca)) + (b.x - a.x)) * G.hx;
    d[1] = ((float)(cell_j(cb) - cell_j(ca)) + (b.y - a.y)) * G.hy;
    d[2] = ((float)(cell_k(cb) - cell_k(ca)) + (b.z - a.z)) * G.hz;
}
// a + delta (world units) -> packed record, clamped to the grid
__device__ __forceinline__ float4 pos_offset(const GridP& G, const float4& a, float dx, float dy, float dz) {
    const int c = __float_as_int(a.w);
    float fx = fmaf(dx, G.ihx, a.x), fy = fmaf(dy, G.ihy, a.y), fz = fmaf(dz, G.ihz, a.z);
    const float flx = floorf(fx), fly = floorf(fy), flz = floorf(fz);
    fx = fminf(fx - flx, 0.99999994f); fy = fminf(fy - fly, 0.99999994f); fz = fminf(fz - flz, 0.99999994f);
    const int ci = clampi(cell_i(c) + (int)flx, 0, G.nx - 1), cj = clampi(cell_j(c) + (int)fly, 0, G.ny - 1), ck = clampi(cell_k(c) + (int)flz, 0, G.nz - 1);
    return make_float4(fx, fy, fz, __int_as_float(cell_pack(ci, cj, ck)));
}

// common G2P gather for a mesh point: v = sum w v_i, va = sum w v~_i, B = sum w v_i (x_i-x_p)^T, g = sum v~_i grad w^T,
// plus the truncated-stencil position correction (see k_g2p)
struct PointGather {
    float vp[3], va[3], B[9], g[9], corr[3];
};
__device__ __forceinline__ void point_gather(const GridP& G, const float4& X, PointGather& o) {
    const int cell = __float_as_int(X.w);
    const int ci = cell_i(cell), cj = cell_j(cell), ck = cell_k(cell);
    Axis ax, ay, az;
    bool complete = axis_setup(ax, X.x, ci, G.nx, G.ihx);
    complete &= axis_setup(ay, X.y, cj, G.ny, G.ihy);
    complete &= axis_setup(az, X.z, ck, G.nz, G.ihz);
    float rx[4], ry[4], rz[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) { rx[q] = G.hx * ((float)(q - 1) - X.x); ry[q] = G.hy * ((float)(q - 1) - X.y); rz[q] = G.hz * ((float)(q - 1) - X.z); }
#pragma unroll
    for (int i = 0; i < 3; ++i) { o.vp[i] = 0.f; o.va[i] = 0.f; o.corr[i] = 0.f; }
#pragma unroll
    for (int i = 0; i < 9; ++i) { o.B[i] = 0.f; o.g[i] = 0.f; }
    float s0 = 0.f, s1x = 0.f, s1y = 0.f, s1z = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int nk = clampi(az.n0 + k, G.a0[2], G.a1[2] - 1);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int nj = clampi(ay.n0 + j, G.a0[1], G.a1[1] - 1);
            const size_t row = nidx(G, 0, nj, nk);
            const float nn = ay.N[j] * az.N[k], dn = ay.D[j] * az.N[k], nd = ay.N[j] * az.D[k];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int ni = clampi(ax.n0 + i, G.a0[0], G.a1[0] - 1);
                const float w = ax.N[i] * nn;
                const float4 t = ldg4(G.vt + row + ni);
                const float dwx = ax.D[i] * nn, dwy = ax.N[i] * dn, dwz = ax.N[i] * nd;
                // post-friction velocity v = c + s (v~ - c), c = collider velocity (0 for the reference's static colliders: v = s v~)
                const float ux = w * fmaf(t.w, t.x - G.cvx, G.cvx), uy = w * fmaf(t.w, t.y - G.cvy, G.cvy), uz = w * fmaf(t.w, t.z - G.cvz, G.cvz);
                o.vp[0] += ux; o.vp[1] += uy; o.vp[2] += uz;
                o.va[0] = fmaf(w, t.x, o.va[0]); o.va[1] = fmaf(w, t.y, o.va[1]); o.va[2] = fmaf(w, t.z, o.va[2]);
                o.B[0] = fmaf(ux, rx[i], o.B[0]); o.B[1] = fmaf(ux, ry[j], o.B[1]); o.B[2] = fmaf(ux, rz[k], o.B[2]);
                o.B[3] = fmaf(uy, rx[i], o.B[3]); o.B[4] = fmaf(uy, ry[j], o.B[4]); o.B[5] = fmaf(uy, rz[k], o.B[5]);
                o.B[6] = fmaf(uz, rx[i], o.B[6]); o.B[7] = fmaf(uz, ry[j], o.B[7]); o.B[8] = fmaf(uz, rz[k], o.B[8]);
                o.g[0] = fmaf(t.x, dwx, o.g[0]); o.g[1] = fmaf(t.x, dwy, o.g[1]); o.g[2] = fmaf(t.x, dwz, o.g[2]);
                o.g[3] = fmaf(t.y, dwx, o.g[3]); o.g[4] = fmaf(t.y, dwy, o.g[4]); o.g[5] = fmaf(t.y, dwz, o.g[5]);
                o.g[6] = fmaf(t.z, dwx, o.g[6]); o.g[7] = fmaf(t.z, dwy, o.g[7]); o.g[8] = fmaf(t.z, dwz, o.g[8]);
                if (!complete) { s0 += w; s1x = fmaf(w, rx[i], s1x); s1y = fmaf(w, ry[j], s1y); s1z = fmaf(w, rz[k], s1z); }
            }
        }
    }
    if (!complete) {
        const float xw = fmaf((float)ci + X.x, G.hx, G.mnx), yw = fmaf((float)cj + X.y, G.hy, G.mny), zw = fmaf((float)ck + X.z, G.hz, G.mnz);
        o.corr[0] = s1x + (s0 - 1.0f) * xw; o.corr[1] = s1y + (s0 - 1.0f) * yw; o.corr[2] = s1z + (s0 - 1.0f) * zw;
    }
}
// C = skew(B) + (1 - damp) sym(B), damp = 1 for the mesh (HybridSolver.cpp:808-824, 920-933)
__device__ __forceinline__ void skew_part(float (&B)[9]) {
    const float a01 = 0.5f * (B[1] - B[3]), a02 = 0.5f * (B[2] - B[6]), a12 = 0.5f * (B[5] - B[7]);
    B[0] = 0.f; B[1] = a01; B[2] = a02; B[3] = -a01; B[4] = 0.f; B[5] = a12; B[6] = -a02; B[7] = -a12; B[8] = 0.f;
}

// warp-cooperative stress scatter shared with the particle force kernel's phase 2 (run length 1 for mesh points):
// f_i -= A grad w_ie    HybridSolver.cpp:444-454
__device__ __forceinline__ void warp_scatter_stress(const GridP& G, float fx, float fy, float fz, int cell, const float (&A)[9], int cnt, int lane) {
    const unsigned FULL = 0xffffffffu;
    const int oi = lane & 3, oj = (lane >> 2) & 3, ok = lane >> 4;
    for (int p = 0; p < cnt; ++p) {
        const int c = __shfl_sync(FULL, cell, p);
        const float pfx = __shfl_sync(FULL, fx, p), pfy = __shfl_sync(FULL, fy, p), pfz = __shfl_sync(FULL, fz, p);
        float a[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) a[i] = __shfl_sync(FULL, A[i], p);
        float wx, dx, wy, dy, wz0, dz0, wz1, dz1;
        bspline_lane(pfx, oi, wx, dx); bspline_lane(pfy, oj, wy, dy); bspline_lane(pfz, ok, wz0, dz0); bspline_lane(pfz, ok + 2, wz1, dz1);
        dx *= G.ihx; dy *= G.ihy; dz0 *= G.ihz; dz1 *= G.ihz;
        const float gx = dx * wy, gy = wx * dy, gz = wx * wy;
        float4 a0, a1;
        { const float d0 = gx * wz0, d1 = gy * wz0, d2 = gz * dz0;
          a0 = make_float4(-fmaf(a[0], d0, fmaf(a[1], d1, a[2] * d2)), -fmaf(a[3], d0, fmaf(a[4], d1, a[5] * d2)), -fmaf(a[6], d0, fmaf(a[7], d1, a[8] * d2)), 0.f); }
        { const float d0 = gx * wz1, d1 = gy * wz1, d2 = gz * dz1;
          a1 = make_float4(-fmaf(a[0], d0, fmaf(a[1], d1, a[2] * d2)), -fmaf(a[3], d0, fmaf(a[4], d1, a[5] * d2)), -fmaf(a[6], d0, fmaf(a[7], d1, a[8] * d2)), 0.f); }
        flush_nodes(G, G.f, c, oi, oj, ok, a0, a1, false);
    }
}

// ------------------------------------------------------------------------------------------------ kernels
// ownership snapshot of a substep (slab contexts only): flags + the masses the P2G of this substep sees (0 for points of other ranks)
__global__ void __launch_bounds__(256) k_mesh_own(MeshDev M, MeshOwn O, const SimClock* __restrict__ clk, unsigned char* __restrict__ owner_v,
                                                  unsigned char* __restrict__ owner_e, float4* __restrict__ VKe, float4* __restrict__ EKe, int nv, int nf) {
    AEP_HALT_PRE(clk);
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < nv) {
        const bool o = mesh_owned(O, __float_as_int(M.VX[t].w));
        owner_v[t] = o ? 1 : 0; float4 kk = M.VK[t]; if (!o) kk.x = 0.f; VKe[t] = kk;
    } else if (t - nv < nf) {
        const int f = t - nv;
        const bool o = mesh_owned(O, __float_as_int(M.EX[f].w));
        owner_e[f] = o ? 1 : 0; float4 kk = M.EK[f]; if (!o) kk.x = 0.f; EKe[f] = kk;
    }
}

// LagrangianMesh::computeVertexInPlaneForces (LagrangianMesh.cpp:382-460), one thread per face
__global__ void __launch_bounds__(128) k_cloth_inplane(MeshDev M, const SimClock* __restrict__ clk, int nf) {
    AEP_HALT_PRE(clk);
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nf) return;
    const float4 e1 = M.ED1[f], e2 = M.ED2[f], e3 = M.ED3[f], r1 = M.RD1[f], r2 = M.RD2[f], r3 = M.RD3[f];
    const float d1[3] = { e1.x, e1.y, e1.z }, d2[3] = { e2.x, e2.y, e2.z }, d3[3] = { e3.x, e3.y, e3.z };
    const float D1[3] = { r1.x, r1.y, r1.z }, D2[3] = { r2.x, r2.y, r2.z }, D3[3] = { r3.x, r3.y, r3.z };
    float Q[9], R[9], Q0[9], R0[9];
    gram_schmidt(d1, d2, d3, Q, R); gram_schmidt(D1, D2, D3, Q0, R0);
    const float i11 = 1.0f / R0[0], i12 = -R0[1] / R0[0] / R0[4], i22 = 1.0f / R0[4];          // geometry.cpp:67-73
    const float r00 = i11 * R[0], r01 = fmaf(i11, R[1], i12 * R[4]), r11 = i22 * R[4];         // invRest * inPlaneR (:431)
    // polar rotation of the 2x2 upper-triangular [[r00 r01],[0 r11]] (det > 0): (A + cof A) / |.|  == U V^T of its SVD (:437-438)
    const float tr = r00 + r11;
    const float inv = rsqrtf(fmaf(tr, tr, r01 * r01));
    const float rot00 = tr * inv, rot01 = r01 * inv, rot10 = -r01 * inv, rot11 = tr * inv;
    const float J = r00 * r11;                                                                   // :441
    const float lj = M.lambda * (J - 1.0f), m2 = 2.0f * M.mu;
    const float P00 = fmaf(m2, r00 - rot00, lj * r11);                                           // :443-444 with invRefMulDet^T
    const float P01 = m2 * (r01 - rot01);
    const float P10 = fmaf(m2, -rot10, lj * (-r01));
    const float P11 = fmaf(m2, r11 - rot11, lj * r00);
    M.PK[f] = make_float4(fmaf(i11, P00, i12 * P10), fmaf(i11, P01, i12 * P11), i22 * P11, i22 * P10);   // :446
    const float c2 = -(P00 * i11 + P01 * i12);                                                   // :452
    const float c3a = -P01 * i22, c3b = -P11 * i22;                                              // :453
    const int4 fa = M.faces[f];
    float f2[3], f3[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) { f2[r] = c2 * Q[3 * r]; f3[r] = fmaf(c3a, Q[3 * r], c3b * Q[3 * r + 1]); }
    float* va = reinterpret_cast<float*>(M.VF + fa.x); float* vb = reinterpret_cast<float*>(M.VF + fa.y); float* vc = reinterpret_cast<float*>(M.VF + fa.z);
#pragma unroll
    for (int r = 0; r < 3; ++r) { atomicAdd(va + r, -(f2[r] + f3[r])); atomicAdd(vb + r, f2[r]); atomicAdd(vc + r, f3[r]); }   // :454-458
}

// forces += vertexOmegas^T * vertexInPlaneForces  (HybridSolver.cpp:378), warp-cooperative, run length 1
__global__ void __launch_bounds__(256) k_vertex_force_scatter(MeshDev M, GridP G, const SimClock* __restrict__ clk, const unsigned char* __restrict__ owner, int nv) {
    AEP_HALT_PRE(clk);
    const int lane = threadIdx.x & 31;
    const int base = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 32;
    if (base >= nv) return;
    const int cnt = min(32, nv - base);
    const unsigned FULL = 0xffffffffu;
    float fx = 0.f, fy = 0.f, fz = 0.f, Fx = 0.f, Fy = 0.f, Fz = 0.f; int cell = 0;
    if (lane < cnt) {
        const float4 X = M.VX[base + lane], F = M.VF[base + lane];
        fx = X.x; fy = X.y; fz = X.z; cell = __float_as_int(X.w);
        if (!owner || owner[base + lane]) { Fx = F.x; Fy = F.y; Fz = F.z; }                      // points of other ranks add nothing here
    }
    const int oi = lane & 3, oj = (lane >> 2) & 3, ok = lane >> 4;
    for (int p = 0; p < cnt; ++p) {
        const int c = __shfl_sync(FULL, cell, p);
        const float pfx = __shfl_sync(FULL, fx, p), pfy = __shfl_sync(FULL, fy, p), pfz = __shfl_sync(FULL, fz, p);
        const float ax = __shfl_sync(FULL, Fx, p), ay = __shfl_sync(FULL, Fy, p), az = __shfl_sync(FULL, Fz, p);
        if (ax == 0.f && ay == 0.f && az == 0.f) continue;
        float wx, wy, wz0, wz1, d;
        bspline_lane(pfx, oi, wx, d); bspline_lane(pfy, oj, wy, d); bspline_lane(pfz, ok, wz0, d); bspline_lane(pfz, ok + 2, wz1, d);
        const float w0 = wx * wy * wz0, w1 = wx * wy * wz1;
        flush_nodes(G, G.f, c, oi, oj, ok, make_float4(w0 * ax, w0 * ay, w0 * az, 0.f), make_float4(w1 * ax, w1 * ay, w1 * az, 0.f), false);
    }
}

// normal / shear part of computeGridForces_ (HybridSolver.cpp:389-455)
__global__ void __launch_bounds__(128) k_cloth_normal(MeshDev M, GridP G, const SimClock* __restrict__ clk, const unsigned char* __restrict__ owner, int nf) {
    AEP_HALT_PRE(clk);
    const int lane = threadIdx.x & 31;
    const int base = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 32;
    if (base >= nf) return;
    const int cnt = min(32, nf - base);
    float fx = 0.f, fy = 0.f, fz = 0.f; int cell = 0;
    float A[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (lane < cnt && (!owner || owner[base + lane])) {
        const int f = base + lane;
        const float4 X = M.EX[f]; fx = X.x; fy = X.y; fz = X.z; cell = __float_as_int(X.w);
        const float4 e1 = M.ED1[f], e2 = M.ED2[f], e3 = M.ED3[f], pk = M.PK[f];
        const float d1[3] = { e1.x, e1.y, e1.z }, d2[3] = { e2.x, e2.y, e2.z }, d3[3] = { e3.x, e3.y, e3.z };
        float Q[9], R[9]; gram_schmidt(d1, d2, d3, Q, R);                                        // :401-402
        float dR[9] = { pk.x, pk.y, M.gamma * R[2], 0.f, pk.z, M.gamma * R[5], 0.f, 0.f, 0.f };  // :406-420
        const float om = 1.0f - R[8];
        dR[8] = R[8] > 1.0f ? 0.0f : -M.kstiff * om * om;                                        // :414-415
        float K[9]; mat_mul_nt(dR, R, K);                                                        // K = dR R^T  :423
        float Sy[9];                                                                             // strictUpper(K) + upper(K)^T  :425-426
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) Sy[3 * r + c] = (c > r) ? K[3 * r + c] : K[3 * c + r];
        // R^{-T}: R upper triangular -> closed-form inverse
        const float i00 = 1.0f / R[0], i11 = 1.0f / R[4], i22 = 1.0f / R[8];
        const float i01 = -R[1] * i00 * i11, i12 = -R[5] * i11 * i22, i02 = (R[1] * R[5] - R[2] * R[4]) * i00 * i11 * i22;
        const float RinvT[9] = { i00, 0.f, 0.f, i01, i11, 0.f, i02, i12, i22 };
        float QS[9], T[9]; mat_mul(Q, Sy, QS); mat_mul(QS, RinvT, T);
        const float4 r1 = M.RD1[f], r2 = M.RD2[f], r3 = M.RD3[f];
        const float rc[3] = { r1.z, r2.z, r3.z };                                                // (rest^T).col(2)  :427
        const float vol = e3.w;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const float dF3 = fmaf(T[3 * r], rc[0], fmaf(T[3 * r + 1], rc[1], T[3 * r + 2] * rc[2]));
            A[3 * r] = vol * dF3 * d3[0]; A[3 * r + 1] = vol * dF3 * d3[1]; A[3 * r + 2] = vol * dF3 * d3[2];      // :429
        }
    }
    warp_scatter_stress(G, fx, fy, fz, cell, A, cnt, lane);
}

// pinned vertices: zero v and v~ in the 3x3x3 node block around every stencil node (HybridSolver.cpp:513-550).
// Per-axis indices are NOT range checked, only the flat index (:538-539) -- reproduced, including the row wrap.
__global__ void k_mesh_pin(MeshDev M, GridP G, const SimClock* __restrict__ clk, int n_fixed) {
    AEP_HALT_PRE(clk);
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_fixed * 64) return;
    const int v = M.fixed_ids[t >> 6], e = t & 63;
    const float4 X = M.VX[v];
    const int cell = __float_as_int(X.w);
    const int oi = e & 3, oj = (e >> 2) & 3, ok = e >> 4;
    float wx, wy, wz, d;
    bspline_lane(X.x, oi, wx, d); bspline_lane(X.y, oj, wy, d); bspline_lane(X.z, ok, wz, d);
    const int ri = cell_i(cell) - 1 + oi, rj = cell_j(cell) - 1 + oj, rk = cell_k(cell) - 1 + ok;
    if (!(wx > 0.f && wy > 0.f && wz > 0.f)) return;                                             // not a nonzero of vertexOmegas_
    if (ri < 0 || ri >= G.nx || rj < 0 || rj >= G.ny || rk < 0 || rk >= G.nz) return;
    const long long Ng = (long long)G.nx * G.ny * G.nz;
    for (int i = ri - 1; i <= ri + 1; ++i) for (int j = rj - 1; j <= rj + 1; ++j) for (int k = rk - 1; k <= rk + 1; ++k) {
        const long long index = ((long long)k * G.ny + j) * G.nx + i;                            // the reference's flat index
        if (index < 0 || index >= Ng) continue;
        const int wi = (int)(index % G.nx), wj = (int)((index / G.nx) % G.ny), wk = (int)(index / ((long long)G.nx * G.ny));
        if (node_held(G, wi, wj, wk)) G.vt[nidx(G, wi, wj, wk)] = make_float4(0.f, 0.f, 0.f, 1.f);
    }
}

// vertices: velocity, affine (damp 1), advection        HybridSolver.cpp:748, 920-926, 948
__global__ void __launch_bounds__(128) k_vertex_g2p(MeshDev M, GridP G, const SimClock* __restrict__ clk, const unsigned char* __restrict__ owner, int nv) {
    AEP_HALT_POST(clk);
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nv || (owner && !owner[v])) return;
    const float dt = clk->dt;
    const float4 X = M.VX[v];
    PointGather pg; point_gather(G, X, pg);
    skew_part(pg.B);
    M.VV[v] = make_float4(pg.vp[0], pg.vp[1], pg.vp[2], pg.B[0]);
    M.VC0[v] = make_float4(pg.B[1], pg.B[2], pg.B[3], pg.B[4]); M.VC1[v] = make_float4(pg.B[5], pg.B[6], pg.B[7], pg.B[8]);
    M.VX[v] = pos_offset(G, X, fmaf(dt, pg.va[0], pg.corr[0]), fmaf(dt, pg.va[1], pg.corr[1]), fmaf(dt, pg.va[2], pg.corr[2]));
}

// elements: mean vertex velocity (:749-756), affine at the OLD centroid (:927-933), d1/d2 from advected vertices and
// d3 += dt grad v~ d3 (:584-606), cone return mapping (:684-722), new centroid (LagrangianMesh.cpp:371-380)
__global__ void __launch_bounds__(128) k_element_g2p(MeshDev M, GridP G, const SimClock* __restrict__ clk, const unsigned char* __restrict__ owner, int nf) {
    AEP_HALT_POST(clk);
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nf || (owner && !owner[f])) return;
    const float dt = clk->dt;
    const float4 X = M.EX[f];
    PointGather pg; point_gather(G, X, pg);
    skew_part(pg.B);
    const int4 fa = M.faces[f];
    const float4 va = M.VV[fa.x], vb = M.VV[fa.y], vc = M.VV[fa.z];
    M.EV[f] = make_float4((va.x + vb.x + vc.x) / 3.0f, (va.y + vb.y + vc.y) / 3.0f, (va.z + vb.z + vc.z) / 3.0f, pg.B[0]);
    M.EC0[f] = make_float4(pg.B[1], pg.B[2], pg.B[3], pg.B[4]); M.EC1[f] = make_float4(pg.B[5], pg.B[6], pg.B[7], pg.B[8]);
    const float4 xa = M.VX[fa.x], xb = M.VX[fa.y], xc = M.VX[fa.z];
    float d1[3], d2[3]; pos_diff(G, xa, xb, d1); pos_diff(G, xa, xc, d2);                        // :591-594
    const float4 e3 = M.ED3[f];
    float d3[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) d3[r] = fmaf(dt, fmaf(pg.g[3 * r], e3.x, fmaf(pg.g[3 * r + 1], e3.y, pg.g[3 * r + 2] * e3.z)), (r == 0 ? e3.x : (r == 1 ? e3.y : e3.z)));   // :601-602
    float Q[9], R[9]; gram_schmidt(d1, d2, d3, Q, R);                                            // :694-695
    float r13 = R[2], r23 = R[5], r33 = R[8];
    if (r33 > 1.0f) { r33 = 1.0f; r13 = 0.f; r23 = 0.f; }                                        // :699-703
    else {
        const float fn = M.kstiff * (r33 - 1.0f) * (r33 - 1.0f);
        const float fs = M.gamma * sqrtf(r13 * r13 + r23 * r23);
        if (fs > M.cf * fn) { const float sc = M.cf * fn / fs; r13 *= sc; r23 *= sc; }           // :711-715
    }
    M.ED1[f] = make_float4(d1[0], d1[1], d1[2], 0.f); M.ED2[f] = make_float4(d2[0], d2[1], d2[2], 0.f);
    M.ED3[f] = make_float4(fmaf(Q[0], r13, fmaf(Q[1], r23, Q[2] * r33)), fmaf(Q[3], r13, fmaf(Q[4], r23, Q[5] * r33)),
                           fmaf(Q[6], r13, fmaf(Q[7], r23, Q[8] * r33)), e3.w);                  // :718-719
    M.EX[f] = pos_offset(G, xa, (d1[0] + d2[0]) / 3.0f, (d1[1] + d2[1]) / 3.0f, (d1[2] + d2[2]) / 3.0f);
}

// packed -> fp64 staging for download: per point x(3) v(3) B rows(9) [ + d1 d2 d3 (9) for elements ]
__global__ void k_mesh_download(const float4* X, const float4* V, const float4* C0, const float4* C1,
                                const float4* D1, const float4* D2, const float4* D3, double* out, int n,
                                double mnx, double mny, double mnz, double hx, double hy, double hz) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const size_t N = (size_t)n;
    const float4 x = X[i], v = V[i], c0 = C0[i], c1 = C1[i];
    const int cell = __float_as_int(x.w);
    out[0 * N + i] = mnx + ((double)cell_i(cell) + (double)x.x) * hx;
    out[1 * N + i] = mny + ((double)cell_j(cell) + (double)x.y) * hy;
    out[2 * N + i] = mnz + ((double)cell_k(cell) + (double)x.z) * hz;
    out[3 * N + i] = v.x; out[4 * N + i] = v.y; out[5 * N + i] = v.z;
    out[6 * N + i] = v.w; out[7 * N + i] = c0.x; out[8 * N + i] = c0.y;
    out[9 * N + i] = c0.z; out[10 * N + i] = c0.w; out[11 * N + i] = c1.x;
    out[12 * N + i] = c1.y; out[13 * N + i] = c1.z; out[14 * N + i] = c1.w;
    if (D1) {
        const float4 a = D1[i], b = D2[i], c = D3[i];
        out[15 * N + i] = a.x; out[16 * N + i] = a.y; out[17 * N + i] = a.z;
        out[18 * N + i] = b.x; out[19 * N + i] = b.y; out[20 * N + i] = b.z;
        out[21 * N + i] = c.x; out[22 * N + i] = c.y; out[23 * N + i] = c.z;
    }
}

// ------------------------------------------------------------------------------------------------ host side
inline void mesh_free(MeshState& m) {
    if (m.block) cudaFree(m.block);
    m.block = nullptr; m.block_bytes = 0; m.nv = m.nf = 0; m.n_fixed = 0; m.owner_v = m.owner_e = nullptr;
}

inline float4 host_pack_pos(const MeshState& m, const GridP& G, double x, double y, double z) {
    const double p[3] = { x, y, z }; const int nres[3] = { G.nx, G.ny, G.nz };
    float fr[3]; int ce[3];
    for (int a = 0; a < 3; ++a) {
        const double u = (p[a] - m.mn[a]) / m.h[a];
        int c = (int)u; double f = u - (double)c;
        if (f < 0.0) { c -= 1; f += 1.0; }
        if (c < 0) { c = 0; f = 0.0; }
        if (c >= nres[a]) { c = nres[a] - 1; f = 0.99999994; }
        float ff = (float)f; if (ff >= 1.0f) ff = 0.99999994f;
        fr[a] = ff; ce[a] = c;
    }
    const int packed = ce[0] | (ce[1] << 10) | (ce[2] << 20);
    float w; std::memcpy(&w, &packed, 4);
    return make_float4(fr[0], fr[1], fr[2], w);
}

// Everything in ONE device allocation, laid out identically on every rank (the layout depends on nv, nf and the number of pinned
// vertices only), so that a peer maps it with one IPC handle and addresses any array by the same offset.  The arrays that G2P
// rewrites come first and are contiguous per point kind: [VX VV VC0 VC1] and [EX EV EC0 EC1 ED1 ED2 ED3].
inline int mesh_upload(MeshState& m, const GridP& G, const double* mn, const double* h, int64_t nv, int64_t nf, const double* vx,
                       const double* vv, const double* vm, const double* vvol, const double* vB, const int32_t* faces,
                       const double* ev, const double* em, const double* evol, const double* eB, const double* ed, const double* eD,
                       const double* fixedv, double mu, double lambda, double shear, double stiff, double fric, cudaStream_t s) {
    (void)vvol;
    if (nv <= 0 || nf <= 0 || !vx || !vv || !vm || !vB || !faces || !ev || !em || !evol || !eB || !ed || !eD) return -1;
    mesh_free(m);
    for (int a = 0; a < 3; ++a) { m.mn[a] = mn[a]; m.h[a] = h[a]; }
    const size_t NV = (size_t)nv, NF = (size_t)nf;
    std::vector<int> fixed;
    if (fixedv) for (size_t i = 0; i < NV; ++i) if (fixedv[i] != 0.0) fixed.push_back((int)i);      // LagrangianMesh.cpp:462-481
    // layout (float4 units)
    const size_t nVarr = 7, nEarr = 15;                       // VX VV VC0 VC1 | VK VKe VF ;  EX EV EC0 EC1 ED1 ED2 ED3 | EK EKe RD1 RD2 RD3 PK + faces(int4)
    const size_t f4 = nVarr * NV + nEarr * NF;
    const size_t bytes = f4 * 16 + ((fixed.size() * 4 + 15) & ~(size_t)15) + ((NV + NF + 255) & ~(size_t)255);
    if (cudaMalloc(&m.block, bytes) != cudaSuccess) return -3;
    m.block_bytes = bytes;
    std::vector<float4> H(f4, make_float4(0.f, 0.f, 0.f, 0.f));
    float4* base = (float4*)m.block; size_t off = 0;
    auto take = [&](float4*& dev, size_t n) { dev = base + off; float4* hp = H.data() + off; off += n; return hp; };
    MeshDev& d = m.d;
    float4 *VX = take(d.VX, NV), *VV = take(d.VV, NV), *VC0 = take(d.VC0, NV), *VC1 = take(d.VC1, NV);
    m.sync_v_off = 0; m.sync_v_bytes = 4 * NV * 16;
    float4 *EX = take(d.EX, NF), *EV = take(d.EV, NF), *EC0 = take(d.EC0, NF), *EC1 = take(d.EC1, NF), *ED1 = take(d.ED1, NF), *ED2 = take(d.ED2, NF), *ED3 = take(d.ED3, NF);
    m.sync_e_off = 4 * NV * 16; m.sync_e_bytes = 7 * NF * 16;
    float4* VKe_dev; float4* EKe_dev;
    float4 *VK = take(d.VK, NV), *VKe = take(VKe_dev, NV); take(d.VF, NV);
    float4 *EK = take(d.EK, NF), *EKe = take(EKe_dev, NF), *RD1 = take(d.RD1, NF), *RD2 = take(d.RD2, NF), *RD3 = take(d.RD3, NF);
    take(d.PK, NF);
    float4* FAh = take(*reinterpret_cast<float4**>(&d.faces), NF);
    for (size_t i = 0; i < NV; ++i) {
        VX[i] = host_pack_pos(m, G, vx[i], vx[NV + i], vx[2 * NV + i]);
        VV[i] = make_float4((float)vv[i], (float)vv[NV + i], (float)vv[2 * NV + i], (float)vB[i]);
        VC0[i] = make_float4((float)vB[NV + i], (float)vB[2 * NV + i], (float)vB[3 * NV + i], (float)vB[4 * NV + i]);
        VC1[i] = make_float4((float)vB[5 * NV + i], (float)vB[6 * NV + i], (float)vB[7 * NV + i], (float)vB[8 * NV + i]);
        VK[i] = make_float4((float)vm[i], 0.f, 0.f, 0.f); VKe[i] = VK[i];
    }
    for (size_t f = 0; f < NF; ++f) {
        const int a = faces[f], b = faces[NF + f], c = faces[2 * NF + f];
        if (a < 0 || b < 0 || c < 0 || a >= nv || b >= nv || c >= nv) { mesh_free(m); return -1; }
        const int4 fa = make_int4(a, b, c, 0); std::memcpy(&FAh[f], &fa, 16);
        // element centroid = mean of its vertices (LagrangianMesh.cpp:371-380, called from the ctor :185)
        EX[f] = host_pack_pos(m, G, (vx[a] + vx[b] + vx[c]) / 3.0, (vx[NV + a] + vx[NV + b] + vx[NV + c]) / 3.0, (vx[2 * NV + a] + vx[2 * NV + b] + vx[2 * NV + c]) / 3.0);
        EV[f] = make_float4((float)ev[f], (float)ev[NF + f], (float)ev[2 * NF + f], (float)eB[f]);
        EC0[f] = make_float4((float)eB[NF + f], (float)eB[2 * NF + f], (float)eB[3 * NF + f], (float)eB[4 * NF + f]);
        EC1[f] = make_float4((float)eB[5 * NF + f], (float)eB[6 * NF + f], (float)eB[7 * NF + f], (float)eB[8 * NF + f]);
        EK[f] = make_float4((float)em[f], 0.f, 0.f, 0.f); EKe[f] = EK[f];
        ED1[f] = make_float4((float)ed[f], (float)ed[NF + f], (float)ed[2 * NF + f], 0.f);
        ED2[f] = make_float4((float)ed[3 * NF + f], (float)ed[4 * NF + f], (float)ed[5 * NF + f], 0.f);
        ED3[f] = make_float4((float)ed[6 * NF + f], (float)ed[7 * NF + f], (float)ed[8 * NF + f], (float)evol[f]);
        RD1[f] = make_float4((float)eD[f], (float)eD[NF + f], (float)eD[2 * NF + f], 0.f);
        RD2[f] = make_float4((float)eD[3 * NF + f], (float)eD[4 * NF + f], (float)eD[5 * NF + f], 0.f);
        RD3[f] = make_float4((float)eD[6 * NF + f], (float)eD[7 * NF + f], (float)eD[8 * NF + f], 0.f);
    }
    d.VKe = VKe_dev; d.EKe = EKe_dev;
    unsigned char* tail = (unsigned char*)(base + f4);
    d.fixed_ids = (int*)tail; tail += (fixed.size() * 4 + 15) & ~(size_t)15;
    m.owner_v = tail; m.owner_e = tail + NV;
    if (cudaMemcpyAsync(base, H.data(), f4 * 16, cudaMemcpyHostToDevice, s) != cudaSuccess) return -2;
    if (!fixed.empty() && cudaMemcpyAsync(d.fixed_ids, fixed.data(), fixed.size() * 4, cudaMemcpyHostToDevice, s) != cudaSuccess) return -2;
    if (cudaMemsetAsync(m.owner_v, 1, NV + NF, s) != cudaSuccess) return -2;
    d.mu = (float)mu; d.lambda = (float)lambda; d.gamma = (float)shear; d.kstiff = (float)stiff; d.cf = (float)fric;
    if (cudaStreamSynchronize(s) != cudaSuccess) return -2;
    m.nv = nv; m.nf = nf; m.n_fixed = (int)fixed.size();
    return 0;
}

inline bool mesh_sliced(const MeshState& m) { return m.own.axis >= 0; }

inline int mesh_own_snapshot(MeshState& m, const SimClock* clk, cudaStream_t s, long long* launches) {
    if (!mesh_sliced(m)) return 0;
    k_mesh_own<<<(int)((m.nv + m.nf + 255) / 256), 256, 0, s>>>(m.d, m.own, clk, m.owner_v, m.owner_e, m.d.VKe, m.d.EKe, (int)m.nv, (int)m.nf);
    *launches += 1;
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

inline int mesh_p2g(MeshState& m, const GridP& G, const SimClock* clk, cudaStream_t s, long long* launches) {
    PartP pv{}, pe{};
    pv.a[PX] = m.d.VX; pv.a[PV] = m.d.VV; pv.a[PC0] = m.d.VC0; pv.a[PC1] = m.d.VC1; pv.a[PK] = m.d.VKe;
    pe.a[PX] = m.d.EX; pe.a[PV] = m.d.EV; pe.a[PC0] = m.d.EC0; pe.a[PC1] = m.d.EC1; pe.a[PK] = m.d.EKe;
    p2g_launch(s, pv, G, m.nv, clk);
    p2g_launch(s, pe, G, m.nf, clk);
    *launches += 2;
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

inline int mesh_forces(MeshState& m, const GridP& G, const SimClock* clk, cudaStream_t s, long long* launches) {
    const unsigned char* ov = mesh_sliced(m) ? m.owner_v : nullptr; const unsigned char* oe = mesh_sliced(m) ? m.owner_e : nullptr;
    cudaMemsetAsync(m.d.VF, 0, (size_t)m.nv * sizeof(float4), s);
    k_cloth_inplane<<<(int)((m.nf + 127) / 128), 128, 0, s>>>(m.d, clk, (int)m.nf);
    k_vertex_force_scatter<<<(int)((m.nv + 255) / 256), 256, 0, s>>>(m.d, G, clk, ov, (int)m.nv);
    k_cloth_normal<<<(int)((m.nf + 127) / 128), 128, 0, s>>>(m.d, G, clk, oe, (int)m.nf);
    *launches += 3;
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

inline int mesh_pin(MeshState& m, const GridP& G, const SimClock* clk, cudaStream_t s, long long* launches) {
    const int n = m.n_fixed * 64;
    k_mesh_pin<<<(n + 127) / 128, 128, 0, s>>>(m.d, G, clk, m.n_fixed);
    *launches += 1;
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

inline int mesh_g2p_vertices(MeshState& m, const GridP& G, SimClock* clk, cudaStream_t s, long long* launches) {
    k_vertex_g2p<<<(int)((m.nv + 127) / 128), 128, 0, s>>>(m.d, G, clk, mesh_sliced(m) ? m.owner_v : nullptr, (int)m.nv);
    *launches += 1;
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}
inline int mesh_g2p_elements(MeshState& m, const GridP& G, SimClock* clk, cudaStream_t s, long long* launches) {
    k_element_g2p<<<(int)((m.nf + 127) / 128), 128, 0, s>>>(m.d, G, clk, mesh_sliced(m) ? m.owner_e : nullptr, (int)m.nf);
    *launches += 1;
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

inline int mesh_download(MeshState& m, double* vx, double* vv, double* vB, double* ex, double* ev, double* eB, double* ed, cudaStream_t s) {
    if (!m.nv) return 0;
    const size_t NV = (size_t)m.nv, NF = (size_t)m.nf;
    double* st = nullptr;
    if (cudaMalloc((void**)&st, std::max(NV * 15, NF * 24) * sizeof(double)) != cudaSuccess) return -3;
    int rc = 0;
    k_mesh_download<<<(int)((NV + 255) / 256), 256, 0, s>>>(m.d.VX, m.d.VV, m.d.VC0, m.d.VC1, nullptr, nullptr, nullptr, st, (int)NV,
                                                          m.mn[0], m.mn[1], m.mn[2], m.h[0], m.h[1], m.h[2]);
    if (vx) cudaMemcpyAsync(vx, st, 3 * NV * sizeof(double), cudaMemcpyDeviceToHost, s);
    if (vv) cudaMemcpyAsync(vv, st + 3 * NV, 3 * NV * sizeof(double), cudaMemcpyDeviceToHost, s);
    if (vB) cudaMemcpyAsync(vB, st + 6 * NV, 9 * NV * sizeof(double), cudaMemcpyDeviceToHost, s);
    if (cudaStreamSynchronize(s) != cudaSuccess) rc = -2;
    k_mesh_download<<<(int)((NF + 255) / 256), 256, 0, s>>>(m.d.EX, m.d.EV, m.d.EC0, m.d.EC1, m.d.ED1, m.d.ED2, m.d.ED3, st, (int)NF,
                                                          m.mn[0], m.mn[1], m.mn[2], m.h[0], m.h[1], m.h[2]);
    if (ex) cudaMemcpyAsync(ex, st, 3 * NF * sizeof(double), cudaMemcpyDeviceToHost, s);
    if (ev) cudaMemcpyAsync(ev, st + 3 * NF, 3 * NF * sizeof(double), cudaMemcpyDeviceToHost, s);
    if (eB) cudaMemcpyAsync(eB, st + 6 * NF, 9 * NF * sizeof(double), cudaMemcpyDeviceToHost, s);
    if (ed) cudaMemcpyAsync(ed, st + 15 * NF, 9 * NF * sizeof(double), cudaMemcpyDeviceToHost, s);
    if (cudaStreamSynchronize(s) != cudaSuccess) rc = -2;
    cudaFree(st);
    return rc;
}

}  // namespace aep
