// aep_scatter.cuh -- the arithmetic of the two scatters (P2G, grid forces), free of memory-space and thread-index details so that
// tests/cpu_math_harness.cpp can run exactly this code on the host against the direct formulas of the reference
// (HybridSolver.cpp:113-231, 356-366).  The kernels in aep_kernels.cuh wrap it: phase A (thread per particle) writes a record,
// phase B (half-warp per particle, lane = (j,k) row of the 4x4x4 stencil) accumulates the 4 nodes of its row from the record.
#pragma once
#include "aep_math.cuh"
#include "aep_pack.cuh"

namespace aep {

// packed accumulators: node i of a lane's row is the float4 (lo[i] | hi[i]) = (x, y | z, w)
struct AccRow {
    f32x2 lo[4], hi[4];
};
__device__ __forceinline__ void acc_zero(AccRow& a) {
#pragma unroll
    for (int i = 0; i < 4; ++i) { a.lo[i] = 0ull; a.hi[i] = 0ull; }
}

// ------------------------------------------------------------------------------------------------ P2G
// particleToGrid_ (HybridSolver.cpp:113-231): m_i = sum w m ; p_i = sum w m (v + (3/h^2) B (x_i - x_p)).
// Per particle the momentum of node offset (i,j,k) is q0 + Qm (i,j,k)^T with Qm = m (3/h^2) B diag(h), q0 = m v - Qm (1 + f).
// Record: P2G_STRIDE float4, an odd stride so that the 32 STS.128 of phase A are bank-conflict free.
//   r0 Nx[4]    r1 Ny[4] (read as float [j])    r2 Nz[4] (float [k])
//   r3 (m, q0x | q0y, q0z)      (mass, momentum) of node offset (0,0,0)
//   r4 r5 r6  (0, Qm[0][c] | Qm[1][c], Qm[2][c]) for c = 0,1,2: what one step along i, j, k adds to (m, p)
//   r7 (cell, cell of the particle before it, -, -)             (read when a run starts only)
#define P2G_STRIDE 9
__device__ __forceinline__ void p2g_make_record(float4* __restrict__ rec, const float4& X, const float4& VM, const float4& c0, const float4& c1,
                                                const float4& c2, float m, float apic, float hx, float hy, float hz, float prev_bits) {
    float Nx[4], Ny[4], Nz[4], D[4];
    bspline4(X.x, Nx, D); bspline4(X.y, Ny, D); bspline4(X.z, Nz, D);
    const float km = m * apic;
    float Q[9] = { km * c0.x * hx, km * c0.y * hy, km * c0.z * hz, km * c1.x * hx, km * c1.y * hy, km * c1.z * hz,
                   km * c2.x * hx, km * c2.y * hy, km * c2.z * hz };
    const float gx = 1.0f + X.x, gy = 1.0f + X.y, gz = 1.0f + X.z;           // x_i - x_p = h (o - (1 + f))
    const float q0x = fmaf(m, VM.x, -(Q[0] * gx + Q[1] * gy + Q[2] * gz));
    const float q0y = fmaf(m, VM.y, -(Q[3] * gx + Q[4] * gy + Q[5] * gz));
    const float q0z = fmaf(m, VM.z, -(Q[6] * gx + Q[7] * gy + Q[8] * gz));
    rec[0] = make_float4(Nx[0], Nx[1], Nx[2], Nx[3]); rec[1] = make_float4(Ny[0], Ny[1], Ny[2], Ny[3]); rec[2] = make_float4(Nz[0], Nz[1], Nz[2], Nz[3]);
    rec[3] = make_float4(m, q0x, q0y, q0z);
    rec[4] = make_float4(0.f, Q[0], Q[3], Q[6]); rec[5] = make_float4(0.f, Q[1], Q[4], Q[7]); rec[6] = make_float4(0.f, Q[2], Q[5], Q[8]);
    rec[7] = make_float4(X.w, prev_bits, 0.f, 0.f);
}
// one particle into the 4 nodes of row (j,k): yoff / zoff = byte offsets of Ny[j] / Nz[k] in the record, J = (j,j), K = (k,k)
__device__ __forceinline__ void p2g_row_accumulate(const float4* __restrict__ r, int yoff, int zoff, f32x2 J, f32x2 K, AccRow& acc) {
    const float4 nx = r[0];
    const float wyz = *reinterpret_cast<const float*>(reinterpret_cast<const char*>(r) + yoff) * *reinterpret_cast<const float*>(reinterpret_cast<const char*>(r) + zoff);
    const ulonglong2 b0 = ld_pairs(r + 3), si = ld_pairs(r + 4), sj = ld_pairs(r + 5), sk = ld_pairs(r + 6);
    f32x2 Tlo = fma2(sj.x, J, fma2(sk.x, K, b0.x));                            // (m, px) of node (0, j, k)
    f32x2 Thi = fma2(sj.y, J, fma2(sk.y, K, b0.y));                            // (py, pz)
    f32x2 W = pk1(nx.x * wyz);
    acc.lo[0] = fma2(W, Tlo, acc.lo[0]); acc.hi[0] = fma2(W, Thi, acc.hi[0]);
    Tlo = add2(Tlo, si.x); Thi = add2(Thi, si.y); W = pk1(nx.y * wyz);
    acc.lo[1] = fma2(W, Tlo, acc.lo[1]); acc.hi[1] = fma2(W, Thi, acc.hi[1]);
    Tlo = add2(Tlo, si.x); Thi = add2(Thi, si.y); W = pk1(nx.z * wyz);
    acc.lo[2] = fma2(W, Tlo, acc.lo[2]); acc.hi[2] = fma2(W, Thi, acc.hi[2]);
    Tlo = add2(Tlo, si.x); Thi = add2(Thi, si.y); W = pk1(nx.w * wyz);
    acc.lo[3] = fma2(W, Tlo, acc.lo[3]); acc.hi[3] = fma2(W, Thi, acc.hi[3]);
}

// ------------------------------------------------------------------------------------------------ grid forces
// f_i += A grad w_i  with  grad w_i = (Dx_i Ny Nz, Nx_i Dy Nz, Nx_i Ny Dz)   (HybridSolver.cpp:356-366), A = -V_p P FE^T
//      = Dx_i U + Nx_i V,  U = A[:,0] Ny Nz,  V = A[:,1] Dy Nz + A[:,2] Ny Dz  per (j,k) row.
// Record (FRC_STRIDE float4, odd stride: conflict-free STS.128):
//   r0 Nx[4]   r1 Dx[4]   r2 r3 (Ny_j, Dy_j) pairs, read as float2 [j]   r4 r5 (Nz_k, Dz_k), float2 [k]
//   r6 r7 r8  columns of A:  (A[0][c], A[1][c] | A[2][c], 0)            r9 (cell, cell of the particle before it, -, -)   (run starts only)
#define FRC_STRIDE 11
__device__ __forceinline__ void frc_make_record(float4* __restrict__ rec, const float (&Nx)[4], const float (&Dx)[4], const float (&Ny)[4], const float (&Dy)[4],
                                                const float (&Nz)[4], const float (&Dz)[4], const float (&A)[9], float cell_bits, float prev_bits = 0.f) {
    rec[0] = make_float4(Nx[0], Nx[1], Nx[2], Nx[3]); rec[1] = make_float4(Dx[0], Dx[1], Dx[2], Dx[3]);
    rec[2] = make_float4(Ny[0], Dy[0], Ny[1], Dy[1]); rec[3] = make_float4(Ny[2], Dy[2], Ny[3], Dy[3]);
    rec[4] = make_float4(Nz[0], Dz[0], Nz[1], Dz[1]); rec[5] = make_float4(Nz[2], Dz[2], Nz[3], Dz[3]);
    rec[6] = make_float4(A[0], A[3], A[6], 0.f); rec[7] = make_float4(A[1], A[4], A[7], 0.f); rec[8] = make_float4(A[2], A[5], A[8], 0.f);
    rec[9] = make_float4(cell_bits, prev_bits, 0.f, 0.f);
}
// yoff / zoff = byte offsets of (Ny,Dy)[j] / (Nz,Dz)[k] in the record
__device__ __forceinline__ void frc_row_accumulate(const float4* __restrict__ r, int yoff, int zoff, AccRow& acc) {
    const float4 nx = r[0], dx = r[1];
    const float2 yj = *reinterpret_cast<const float2*>(reinterpret_cast<const char*>(r) + yoff), zk = *reinterpret_cast<const float2*>(reinterpret_cast<const char*>(r) + zoff);
    const ulonglong2 a0 = ld_pairs(r + 6), a1 = ld_pairs(r + 7), a2 = ld_pairs(r + 8);
    const f32x2 AA = pk1(yj.x * zk.x), BB = pk1(yj.y * zk.x), CC = pk1(yj.x * zk.y);              // Ny Nz, Dy Nz, Ny Dz
    const f32x2 Ulo = mul2(a0.x, AA), Uhi = mul2(a0.y, AA);
    const f32x2 Vlo = fma2(a1.x, BB, mul2(a2.x, CC)), Vhi = fma2(a1.y, BB, mul2(a2.y, CC));
    acc.lo[0] = fma2(Ulo, pk1(dx.x), fma2(Vlo, pk1(nx.x), acc.lo[0])); acc.hi[0] = fma2(Uhi, pk1(dx.x), fma2(Vhi, pk1(nx.x), acc.hi[0]));
    acc.lo[1] = fma2(Ulo, pk1(dx.y), fma2(Vlo, pk1(nx.y), acc.lo[1])); acc.hi[1] = fma2(Uhi, pk1(dx.y), fma2(Vhi, pk1(nx.y), acc.hi[1]));
    acc.lo[2] = fma2(Ulo, pk1(dx.z), fma2(Vlo, pk1(nx.z), acc.lo[2])); acc.hi[2] = fma2(Uhi, pk1(dx.z), fma2(Vhi, pk1(nx.z), acc.hi[2]));
    acc.lo[3] = fma2(Ulo, pk1(dx.w), fma2(Vlo, pk1(nx.w), acc.lo[3])); acc.hi[3] = fma2(Uhi, pk1(dx.w), fma2(Vhi, pk1(nx.w), acc.hi[3]));
}

}  // namespace aep
