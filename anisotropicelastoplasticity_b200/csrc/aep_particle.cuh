// aep_particle.cuh -- the two particle kernels of a substep (sm_100a):
//   k_forces    computeGridForces_, particle part                                   HybridSolver.cpp:252-368
//   k_g2p2g     G2P / APIC / advection / F update / return mapping, and -- fused -- the P2G of the NEXT substep
//               HybridSolver.cpp:739-825, 940-951, 553-578, 612-681 | 113-231
// Both share one skeleton.  A warp owns 32 x ROUNDS consecutive cell-sorted particles, each of its half-warps a contiguous
// stretch of 16 x ROUNDS.  Per round a lane works on one particle (phase A: gather from the grid, constitutive update, record for
// the scatter) and then the half-warp walks its 16 records with one lane per (j,k) stencil row (phase B: scatter window in
// registers, aep_kernels.cuh).  Everything a round needs arrives while the round before computes:
//   * the (cells+3) x 4 x 4 node box of the half-warp's 16 particles by TMA (cp.async.bulk.tensor, one lane issues, an mbarrier
//     completes) -- no LSU wavefronts, no registers, zero fill outside the grid / the slab's reach;
//   * the particle records (X two rounds ahead, F_E / constants one round ahead) by cp.async into the warp's shared memory.
// SCATTER = true fuses the P2G of the next substep into the G2P kernel (records made from registers: V and B never travel through
// memory, one launch and 80 B / particle less).  Built, measured, NOT the default: the fused kernel needs 128 registers and runs 16
// warps per SM, where the scatter alone (k_p2g, 63 registers) runs 48: 14.8 against 16.4 ms per substep at rest and 19.1 against
// 21.9 in the flowing state for two kernels against one (same box; profiles/README.md, round 2).  AEP_FUSED=1 selects it.
#pragma once
#include <cuda.h>
#include "aep_kernels.cuh"

namespace aep {

// ------------------------------------------------------------------------------------------------ async copies
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16_stream(float4* smem_dst, const float4* gsrc) {      // L2 only: streaming particle data stays out of L1
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// wait for the phase with the given parity; a TMA that never completes is a bug, not a reason to hang the GPU: trap after ~1 s
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    unsigned ok = 0;
    for (unsigned spin = 0; !ok; ++spin) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (spin > (1u << 24)) __trap();
    }
}
// 4-D tiled TMA load: coordinates (c, x, y, z) in elements of the tensor map (c = the 4 floats of a node)
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* tm, unsigned long long* bar, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(smem_u32(smem_dst)), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

// ------------------------------------------------------------------------------------------------ half-warp grid tile
// 16 cell-sorted particles normally sit in one row of cells (same j,k; two cells along x at 8 particles per cell), so their 64-node
// stencils live in a (cells+3) x 4 x 4 node box.  One lane asks the TMA unit for the TILE_W x 4 x 4 box of G.vt around them; it
// lands as tile[(k*4 + j)*TILE_W + x] in the half-warp's shared memory and completes the half-warp's mbarrier.  Nodes outside the
// tensor (the grid, or the planes a slab context holds) come back as zeros and meet zero weights (HS:44-46).
//
// Particles that do NOT fit their half-warp's box -- the row of cells wraps, a particle has changed its cell since the last physical
// sort, it arrived from a neighbouring slab -- are DEFERRED: the lane appends the particle's slot to a list, takes the place of a
// massless, volumeless stand-in at the reference cell (so that every address stays inside the box) and writes nothing.  A second,
// small launch of the same kernel (LIST) then walks that list with the global-memory gather.  A running simulation always has a
// few % of such particles; letting them pull their whole half-warp onto the global-memory path (round 2's first build) cost the
// fused kernel 2.2x and the force kernel 1.5x at 4 % movers.
struct DeferP {
    unsigned int* list;                 // slots of the deferred particles (capacity = particle capacity)
    unsigned int* count;
};
struct HwTile {
    int ox0;                            // x node coordinate of tile column 0
};
__device__ __forceinline__ bool cell_fits(int cell, int cref) {
    const int ox0 = cell_i(cref) - 1 - TILE_SLACK, ci = cell_i(cell);
    return ((cell ^ cref) >> 10) == 0 && (ci - 1) >= ox0 && (ci + 2) < ox0 + TILE_W;
}
// issues the box of the half-warp's 16 particles (always: the reference lane fits its own box); `fits` = this lane's particle lies in it.
// The reference cell is the first lane's -- or the ninth lane's, when that one has more of the half-warp with it (the first lane
// may be the stray one).  `cell` comes back as the cell the lane works on: its own, or -- a stray's stand-in -- the cell of the nearest
// fitting lane before it (after it, if there is none), so that the stand-in neither leaves the box nor breaks the run of same-cell
// particles it sits in.
__device__ __forceinline__ void hw_tile_issue(const CUtensorMap* tm, const GridP& G, float4* tile, unsigned long long* bar, HwTile& T, int& cell, bool& fits) {
    const int lane = threadIdx.x & 31, h0 = lane & 16, s = lane & 15;
    const int c1 = __shfl_sync(0xffffffffu, cell, h0), c2 = __shfl_sync(0xffffffffu, cell, h0 | 8);
    const bool f1 = cell_fits(cell, c1), f2 = cell_fits(cell, c2);
    const unsigned b1 = (__ballot_sync(0xffffffffu, f1) >> h0) & 0xffffu, b2 = (__ballot_sync(0xffffffffu, f2) >> h0) & 0xffffu;
    const bool second = __popc(b2) > __popc(b1);
    const int cref = second ? c2 : c1; fits = second ? f2 : f1;
    const unsigned fm = second ? b2 : b1;                                   // never 0: the reference lane fits
    T.ox0 = cell_i(cref) - 1 - TILE_SLACK;
    const unsigned below = fm & ((1u << s) - 1u);
    const int src = below ? 31 - __clz(below) : __ffs(fm) - 1;
    const int sub = __shfl_sync(0xffffffffu, cell, h0 | src);
    cell = fits ? cell : sub;
    if (s == 0) {
        mbar_expect_tx(bar, TILE_F4 * 16);
        tma_load_4d(tile, tm, bar, 0, T.ox0 - G.a0[0], cell_j(cref) - 1 - G.a0[1], cell_k(cref) - 1 - G.a0[2]);
    }
}
// the lanes of `defer` append their particle slots to the list (one atomic per warp)
__device__ __forceinline__ void defer_append(const DeferP& D, bool defer, int p) {
    const unsigned m = __ballot_sync(0xffffffffu, defer);
    if (!m) return;
    const int lane = threadIdx.x & 31;
    unsigned base = 0;
    if (lane == (__ffs(m) - 1)) base = atomicAdd(D.count, (unsigned)__popc(m));
    base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
    if (defer) D.list[base + __popc(m & ((1u << lane) - 1u))] = (unsigned)p;
}
// LIST kernels: a persistent grid walks the deferred list
#define AEP_LIST_CTAS (148 * 8)

// ================================================================================================ forces
// computeGridForces_, particle part (HybridSolver.cpp:252-368).  Phase A (thread per particle): gather
// grad v = sum_i v_i (grad w_i)^T, Fhat = (I + dt grad v) FE, SVD, stress, A = -V_p P FE^T.  Phase B (half-warp per particle):
// f_i += A grad w_ip, window open over ROUNDS x 16 particles.
// LIST = false: the cell-sorted particles, gather from the TMA boxes, strays deferred.  LIST = true: the deferred particles (D.list),
// gather from global memory, persistent grid.
#define FRC_NT 128
#define FRC_HW_PAD 2
#define FRC_HW_F4 (16 * FRC_STRIDE + FRC_HW_PAD)
#ifndef FRC_MIN_CTAS
#define FRC_MIN_CTAS 4
#endif
#ifndef FRC_SPLIT_MIN_CTAS
#define FRC_SPLIT_MIN_CTAS 8            // the gather / stress kernel alone fits 64 registers without spills: 32 warps per SM (2.86 ms against 3.06 at 5 CTAs,
#endif                                  // C5 at rest); its list pass (global-memory gather) wants the registers: 5 CTAs (0.28 against 0.46 ms)
#ifndef FRC_LIST_MIN_CTAS
#define FRC_LIST_MIN_CTAS 5
#endif
// SPLIT: phase A only -- A = -V_p P F_E^T goes to memory (3 float4 per particle, `Aout`) and k_force_scatter does phase B as its own,
// small-register kernel (the same reason G2P and P2G are two kernels: the scatter is bound by the LSU pipe and wants warps, the gather
// by issue slots and wants registers).
struct ForceA { float4* a[3]; };
template <bool SPLIT>
struct __align__(128) FrcWarpSmemT {
    float4 tile[2][TILE_F4];
    float4 rec[2][SPLIT ? 1 : FRC_HW_F4];
    float4 x[2][32];
    float4 e[3][32];
    float4 bounce[SPLIT ? 1 : 32];      // SPLIT: 6.6 KB per warp, 8 CTAs of 4 warps fit an SM's 228 KB
    unsigned long long bar[2];
};
template <int ROUNDS, bool LIST, bool SPLIT>
__global__ void __launch_bounds__(FRC_NT, SPLIT ? (LIST ? FRC_LIST_MIN_CTAS : FRC_SPLIT_MIN_CTAS) : FRC_MIN_CTAS) k_forces(PartP P, GridP G, const __grid_constant__ CUtensorMap tm, MatParams mpar,
                                                                  const SimClock* __restrict__ clk, DeferP D, ForceA Aout) {
    AEP_HALT_PRE(clk);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    typedef FrcWarpSmemT<SPLIT> FrcWarpSmem;
    FrcWarpSmem& W = reinterpret_cast<FrcWarpSmem*>(smem_raw)[threadIdx.x >> 5];
    const int n = LIST ? (int)*D.count : clk->n_slots;                        // particles this launch walks
    const int lane = threadIdx.x & 31, hw = lane >> 4, s = lane & 15, j = s & 3, k = s >> 2;
    const int cta_particles = FRC_NT * ROUNDS;
    const int nchunks = (n + cta_particles - 1) / cta_particles;
    const float dt = clk->dt;
    if (!LIST) {
        if (lane == 0) { mbar_init(&W.bar[0], 1); mbar_init(&W.bar[1], 1); mbar_fence_init(); }
        __syncwarp();
    }
    float4* slot = W.bounce + lane;
    float4* tile = W.tile[hw];
    unsigned long long* bar = &W.bar[hw];
    unsigned tphase = 0;
    int yoff = 32 + 8 * j, zoff = 64 + 8 * k;                                 // byte offsets of (Ny,Dy)[j], (Nz,Dz)[k] inside a record
    asm volatile("" : "+r"(yoff), "+r"(zoff));                                // lane constants: keep them in registers
    const float4* recs = W.rec[hw];
    auto slot_of = [&](int q) -> int { q = min(q, n - 1); return LIST ? (int)D.list[q] : q; };
    for (int chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {       // LIST: persistent grid; else one trip (consecutive chunks: the gather lives on L2 locality)
        const int wbase = chunk * cta_particles + (threadIdx.x >> 5) * (32 * ROUNDS);
        if (wbase >= n) break;                                                // warp-uniform; no block-level barrier below
        const int hbase = wbase + hw * (16 * ROUNDS);
        HwTile T, Tn; T.ox0 = 0; Tn = T;
        bool fit = true, fit_n = true;
        int cuse, cuse_n = 0;                                                 // the cell a lane works on (a stray's stand-in: see hw_tile_issue)
        {   // prologue: X of round 0 (waited for), then its tile, its F_E and the X of round 1 in flight
            const int p0 = slot_of(hbase + s);
            cp_async16_stream(&W.x[0][lane], P.a[PX] + p0);
            cp_async_commit(); cp_async_wait_all();
            cuse = __float_as_int(W.x[0][lane].w);
            if (!LIST) hw_tile_issue(&tm, G, tile, bar, T, cuse, fit);
#pragma unroll
            for (int a = 0; a < 3; ++a) cp_async16_stream(&W.e[a][lane], P.a[PE0 + a] + p0);
            if (ROUNDS > 1) cp_async16_stream(&W.x[1][lane], P.a[PX] + slot_of(hbase + 16 + s));
            cp_async_commit();
        }
        AccRow acc; acc_zero(acc);
        int wcell = -1;
#pragma unroll 1
        for (int round = 0; round < ROUNDS; ++round) {
            const bool more = round + 1 < ROUNDS;
            unsigned starts;
            {   // ---- phase A
                const float4 X = W.x[round & 1][lane];
                const int q = hbase + round * 16 + s;
                const bool act = q < n && fit;                                 // padding lanes and deferred strays: zero volume, at the reference cell
                if (!LIST) defer_append(D, q < n && !fit, q);
                const int cell = cuse;
                Axis ax, ay, az;
                axis_setup(ax, X.x, cell_i(cell), G.nx, G.ihx); axis_setup(ay, X.y, cell_j(cell), G.ny, G.ihy); axis_setup(az, X.z, cell_k(cell), G.nz, G.ihz);
                float g[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                if (!LIST) { mbar_wait(bar, tphase); tphase ^= 1u; gather_grad<2>(G, ax, ay, az, tile, ax.n0 - T.ox0, g); }
                else gather_grad<0>(G, ax, ay, az, tile, 0, g);
                cp_async_wait_all();                                           // F_E of this round, X of the next
                __syncwarp();                                                  // every lane is done with the tile (and phase B of the round before with the records)
                if (more) { cuse_n = __float_as_int(W.x[(round + 1) & 1][lane].w); if (!LIST) hw_tile_issue(&tm, G, tile, bar, Tn, cuse_n, fit_n); }
                const float4 e0 = W.e[0][lane], e1 = W.e[1][lane], e2 = W.e[2][lane];
                if (more) {                                                    // this lane's slots are consumed: next round's F_E, X of the round after
                    const int pn = slot_of(hbase + (round + 1) * 16 + s);
#pragma unroll
                    for (int a = 0; a < 3; ++a) cp_async16_stream(&W.e[a][lane], P.a[PE0 + a] + pn);
                    if (round + 2 < ROUNDS) cp_async16_stream(&W.x[round & 1][lane], P.a[PX] + slot_of(hbase + (round + 2) * 16 + s));
                    cp_async_commit();
                }
                const float FE[9] = { e0.x, e0.y, e0.z, e1.x, e1.y, e1.z, e2.x, e2.y, e2.z };
                float GF[9], Fh[9], A[9];
                mat_mul(g, FE, GF);
#pragma unroll
                for (int i = 0; i < 9; ++i) Fh[i] = fmaf(dt, GF[i], FE[i]);      // HybridSolver.cpp:306
                stress_times_FEt(mpar, Fh, FE, act ? -e0.w : 0.0f, e2.w, A);     // A := -V_p P FE^T (sign of :356-366 folded in)
                if (SPLIT) {
                    if (act) {
                        const int p = slot_of(q);
                        Aout.a[0][p] = make_float4(A[0], A[1], A[2], 0.f); Aout.a[1][p] = make_float4(A[3], A[4], A[5], 0.f); Aout.a[2][p] = make_float4(A[6], A[7], A[8], 0.f);
                    }
                    starts = 0u;
                } else {
                    int prev;
                    starts = run_starts(cell, wcell, prev);
                    frc_make_record(W.rec[hw] + s * FRC_STRIDE, ax.N, ax.D, ay.N, ay.D, az.N, az.D, A, __int_as_float(cell), __int_as_float(prev));
                }
                T = Tn; fit = fit_n; cuse = cuse_n;
            }
            if (SPLIT) continue;
            __syncwarp();
            // ---- phase B: f_i += A grad w_i  with  grad w_i = (Dx_i Ny Nz, Nx_i Dy Nz, Nx_i Ny Dz)   (HybridSolver.cpp:356-366)
            //      = Dx_i U + Nx_i V,  U = A[:,0] Ny Nz,  V = A[:,1] Dy Nz + A[:,2] Ny Dz  per (j,k) row
#pragma unroll 1
            for (int it = 0; it < 16; ++it) {
                const float4* r = recs + it * FRC_STRIDE;
                if ((starts >> it) & 1u) {                                      // a run of particles sharing a cell starts here
                    const float2 cn = *reinterpret_cast<const float2*>(r + 9);
                    if (singleton_at(r + 9, FRC_STRIDE, starts, it, __float_as_int(cn.y))) {      // a lone mover inside a run (see p2g_phase_b)
                        AccRow one; acc_zero(one);
                        frc_row_accumulate(r, yoff, zoff, one);
                        flush_row_pk(G, G.f, slot, __float_as_int(cn.x), j, k, one, false);
                        starts &= ~(2u << it);
                        continue;
                    }
                    window_move(G, G.f, slot, __float_as_int(cn.y), __float_as_int(cn.x), j, k, acc, false);
                }
                frc_row_accumulate(r, yoff, zoff, acc);
            }
        }
        if (wcell >= 0) flush_row_pk(G, G.f, slot, wcell, j, k, acc, false);
        if (!LIST) break;
        __syncwarp();
    }
}

// rounds per half-warp: 8 for large scenes (fewest reductions per particle), fewer when that would leave SMs without work
inline int rounds_for(long long n) {
    static const int forced = getenv("AEP_FORCE_ROUNDS") ? atoi(getenv("AEP_FORCE_ROUNDS")) : 0;     // development: 1, 2 or 8
    if (forced == 1 || forced == 2 || forced == 8) return forced;
    return n >= (1ll << 22) ? 8 : (n >= (1ll << 19) ? 2 : 1);
}
// ---- phase B alone: f_i += A grad w_ip from the A stored by k_forces<SPLIT> (HybridSolver.cpp:356-366).  Same mapping as k_p2g: 256
// threads, a warp owns 32 x ROUNDS consecutive particles, window open over the rounds.  Strays need no list here: a particle
// out of its cell order is just a run of length one (singleton_at).
#define FRC_SCATTER_SMEM ((8 * 2 * FRC_HW_F4 + 256) * 16)
template <int ROUNDS>
__global__ void __launch_bounds__(256) k_force_scatter(PartP P, GridP G, ForceA Ain, int n, const SimClock* __restrict__ clk) {
    AEP_HALT_PRE(clk);
    extern __shared__ __align__(128) unsigned char smem_raw[];                // 49.6 KB: above the static limit
    float4 (*stage)[2][FRC_HW_F4] = reinterpret_cast<float4 (*)[2][FRC_HW_F4]>(smem_raw);
    float4* bounce = reinterpret_cast<float4*>(smem_raw) + 8 * 2 * FRC_HW_F4;  // lane-private slots of requad()
    if (n < 0) n = clk->n_slots;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    float4* slot = bounce + threadIdx.x;
    const int cta_particles = 256 * ROUNDS;
    const int chunk = strided_chunk(blockIdx.x, (n + cta_particles - 1) / cta_particles, G.strips);
    if (chunk < 0) return;
    const int hw = lane >> 4, s = lane & 15, j = s & 3, k = s >> 2;
    const int hbase = chunk * cta_particles + wib * (32 * ROUNDS) + hw * (16 * ROUNDS);
    if (chunk * cta_particles + wib * (32 * ROUNDS) >= n) return;             // warp-uniform; no block-level barrier below
    int yoff = 32 + 8 * j, zoff = 64 + 8 * k;                                 // byte offsets of (Ny,Dy)[j], (Nz,Dz)[k] inside a record
    yoff = __shfl_sync(0xffffffffu, yoff, lane); zoff = __shfl_sync(0xffffffffu, zoff, lane);
    AccRow acc; acc_zero(acc);
    const float4* recs = &stage[wib][hw][0];
    int wcell = -1;
#pragma unroll 1
    for (int round = 0; round < ROUNDS; ++round) {
        unsigned starts;
        {   // ---- phase A: thread per particle
            const int q = hbase + round * 16 + s;                            // may lie past the end: such lanes repeat the last particle with A = 0
            const int p = min(q, n - 1);
            const float4 X = ldg4(P.a[PX] + p);
            float4 a0 = ldg4(Ain.a[0] + p), a1 = ldg4(Ain.a[1] + p), a2 = ldg4(Ain.a[2] + p);
            if (q >= n) { a0 = a1 = a2 = make_float4(0.f, 0.f, 0.f, 0.f); }
            const int cell = __float_as_int(X.w);
            Axis ax, ay, az;
            axis_setup(ax, X.x, cell_i(cell), G.nx, G.ihx); axis_setup(ay, X.y, cell_j(cell), G.ny, G.ihy); axis_setup(az, X.z, cell_k(cell), G.nz, G.ihz);
            const float A[9] = { a0.x, a0.y, a0.z, a1.x, a1.y, a1.z, a2.x, a2.y, a2.z };
            int prev;
            starts = run_starts(cell, wcell, prev);
            __syncwarp();                                                    // phase B of the round before is done with the records
            frc_make_record(&stage[wib][hw][s * FRC_STRIDE], ax.N, ax.D, ay.N, ay.D, az.N, az.D, A, X.w, __int_as_float(prev));
        }
        __syncwarp();
#pragma unroll 1
        for (int it = 0; it < 16; ++it) {
            const float4* r = recs + it * FRC_STRIDE;
            if ((starts >> it) & 1u) {                                        // a run of particles sharing a cell starts here
                const float2 cn = *reinterpret_cast<const float2*>(r + 9);
                if (singleton_at(r + 9, FRC_STRIDE, starts, it, __float_as_int(cn.y))) {
                    AccRow one; acc_zero(one);
                    frc_row_accumulate(r, yoff, zoff, one);
                    flush_row_pk(G, G.f, slot, __float_as_int(cn.x), j, k, one, false);
                    starts &= ~(2u << it);
                    continue;
                }
                window_move(G, G.f, slot, __float_as_int(cn.y), __float_as_int(cn.x), j, k, acc, false);
            }
            frc_row_accumulate(r, yoff, zoff, acc);
        }
    }
    if (wcell >= 0) flush_row_pk(G, G.f, slot, wcell, j, k, acc, false);
}
template <int ROUNDS>
inline void force_scatter_launch_r(cudaStream_t st, const PartP& P, const GridP& G, const ForceA& A, long long n_hi, const SimClock* clk, bool device_count) {
    const long long per_cta = 256ll * ROUNDS;
    const int chunks = (int)((n_hi + per_cta - 1) / per_cta);
    k_force_scatter<ROUNDS><<<strided_grid(chunks, G.strips), 256, FRC_SCATTER_SMEM, st>>>(P, G, A, device_count ? -(int)n_hi : (int)n_hi, clk);
}

template <int ROUNDS, bool SPLIT>
inline cudaError_t forces_launch_r(cudaStream_t st, const PartP& P, const GridP& G, const CUtensorMap& tm, const MatParams& mat, const SimClock* clk, long long n_hi, const DeferP& D, const ForceA& A) {
    const int smem = (int)sizeof(FrcWarpSmemT<SPLIT>) * (FRC_NT / 32);
    const long long per_cta = (long long)FRC_NT * ROUNDS;
    k_forces<ROUNDS, false, SPLIT><<<(unsigned)((n_hi + per_cta - 1) / per_cta), FRC_NT, smem, st>>>(P, G, tm, mat, clk, D, A);
    return cudaSuccess;
}
// The three launches of the force stage, separately (the engine times them one by one in its profiled pass): main kernel over the
// cell-sorted particles (zeroes the count of the deferred list first, in stream order), list pass over the strays, scatter (SPLIT).
inline cudaError_t forces_main_launch(cudaStream_t st, const PartP& P, const GridP& G, const CUtensorMap& tm, const MatParams& mat, const SimClock* clk, long long n_hi, const DeferP& D,
                                      const ForceA& A, bool split) {
    cudaError_t e = cudaMemsetAsync(D.count, 0, sizeof(unsigned int), st);
    if (e != cudaSuccess) return e;
    const int rounds = rounds_for(n_hi);
    if (split) {
        if (rounds == 8) forces_launch_r<8, true>(st, P, G, tm, mat, clk, n_hi, D, A); else if (rounds == 2) forces_launch_r<2, true>(st, P, G, tm, mat, clk, n_hi, D, A); else forces_launch_r<1, true>(st, P, G, tm, mat, clk, n_hi, D, A);
    } else {
        if (rounds == 8) forces_launch_r<8, false>(st, P, G, tm, mat, clk, n_hi, D, A); else if (rounds == 2) forces_launch_r<2, false>(st, P, G, tm, mat, clk, n_hi, D, A); else forces_launch_r<1, false>(st, P, G, tm, mat, clk, n_hi, D, A);
    }
    return cudaGetLastError();
}
inline cudaError_t forces_list_launch(cudaStream_t st, const PartP& P, const GridP& G, const CUtensorMap& tm, const MatParams& mat, const SimClock* clk, long long n_hi, const DeferP& D,
                                      const ForceA& A, bool split) {
    const long long list_ctas = std::min<long long>((n_hi + FRC_NT - 1) / FRC_NT, AEP_LIST_CTAS);
    if (split) k_forces<1, true, true><<<(unsigned)list_ctas, FRC_NT, (int)sizeof(FrcWarpSmemT<true>) * (FRC_NT / 32), st>>>(P, G, tm, mat, clk, D, A);
    else k_forces<1, true, false><<<(unsigned)list_ctas, FRC_NT, (int)sizeof(FrcWarpSmemT<false>) * (FRC_NT / 32), st>>>(P, G, tm, mat, clk, D, A);
    return cudaGetLastError();
}
inline cudaError_t force_scatter_launch(cudaStream_t st, const PartP& P, const GridP& G, const ForceA& A, long long n_hi, const SimClock* clk, bool device_count) {
    switch (rounds_for(n_hi)) {
    case 8: force_scatter_launch_r<8>(st, P, G, A, n_hi, clk, device_count); break;
    case 2: force_scatter_launch_r<2>(st, P, G, A, n_hi, clk, device_count); break;
    default: force_scatter_launch_r<1>(st, P, G, A, n_hi, clk, device_count); break;
    }
    return cudaGetLastError();
}

// ================================================================================================ G2P (+ P2G)
// updateParticleVelocities_ (HybridSolver.cpp:739-745), updateAffineMomenta_ with damp 0 (:760-825, :908-917),
// advection x = sum w (x_i + dt v~_i) (:942-945), updateDeformationGradient_ (:553-578), updatePlasticity_
// (:612-681), all in registers, one thread per particle; then, SCATTER, particleToGrid_ (:113-231) of the advected particle from
// the same registers.  F_P (48 B) is read and written only by particles whose return mapping changed a singular value: for all
// others the reference's F_P update is the identity (aep_math.cuh, return_map_project).
#define G2G_NT 128
#ifndef G2G_MIN_CTAS
#define G2G_MIN_CTAS 4                  // fused kernel: 128 registers
#endif
#ifndef G2P_MIN_CTAS
#define G2P_MIN_CTAS 5                  // G2P alone: 96 registers (8 B of spills), 20 warps per SM: 5.37 against 5.56 ms (C5, rest)
#endif
template <bool SCATTER>
struct __align__(128) G2GWarpSmemT {
    float4 tile[2][TILE_F4];
    float4 rec[2][SCATTER ? P2G_HW_F4 : 1];   // the scatter's records: only the fused kernel has them
    float4 x[2][32];
    float4 e[4][32];                    // E0 E1 E2 K
    float4 bounce[32];
    unsigned long long bar[2];
};
template <int ROUNDS, bool SCATTER, bool LIST>
__global__ void __launch_bounds__(G2G_NT, SCATTER ? G2G_MIN_CTAS : G2P_MIN_CTAS) k_g2p2g(PartP P, GridP G, const __grid_constant__ CUtensorMap tm, MatParams mpar,
                                                               SimClock* __restrict__ clk, MigList ML, DeferP D) {
    AEP_HALT_POST(clk);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    typedef G2GWarpSmemT<SCATTER> G2GWarpSmem;
    G2GWarpSmem& W = reinterpret_cast<G2GWarpSmem*>(smem_raw)[threadIdx.x >> 5];
    const int n = LIST ? (int)*D.count : clk->n_slots;                         // particles this launch walks
    const int lane = threadIdx.x & 31, hw = lane >> 4, s = lane & 15, j = s & 3, k = s >> 2;
    const int cta_particles = G2G_NT * ROUNDS;
    const int nchunks = (n + cta_particles - 1) / cta_particles;
    const float dt = clk->dt;
    if (!LIST) {
        if (lane == 0) { mbar_init(&W.bar[0], 1); mbar_init(&W.bar[1], 1); mbar_fence_init(); }
        __syncwarp();
    }
    float4* slot = W.bounce + lane;
    float4* tile = W.tile[hw];
    unsigned long long* bar = &W.bar[hw];
    unsigned tphase = 0;
    // phase-B lane constants (see k_p2g)
    float fj = (float)j, fk = (float)k;
    int yoff = 16 + 4 * j, zoff = 32 + 4 * k;
    fj = __shfl_sync(0xffffffffu, fj, lane); fk = __shfl_sync(0xffffffffu, fk, lane);
    yoff = __shfl_sync(0xffffffffu, yoff, lane); zoff = __shfl_sync(0xffffffffu, zoff, lane);
    const f32x2 J = pk1(fj), K = pk1(fk);
    auto slot_of = [&](int q) -> int { q = min(q, n - 1); return LIST ? (int)D.list[q] : q; };
    // LIST: a persistent grid over the deferred list.  Else one trip; scatter kernels spread concurrently running CTAs over far-apart
    // parts of the sorted order (strided_chunk)
    for (int cb = blockIdx.x; LIST ? cb < nchunks : true; cb += gridDim.x) {
        const int chunk = LIST ? cb : (SCATTER ? strided_chunk(cb, nchunks, G.strips) : (cb < nchunks ? cb : -1));
        if (chunk < 0) break;
        const int wbase = chunk * cta_particles + (threadIdx.x >> 5) * (32 * ROUNDS);
        if (wbase >= n) break;                                                 // warp-uniform; no block-level barrier below
        const int hbase = wbase + hw * (16 * ROUNDS);
        HwTile T, Tn; T.ox0 = 0; Tn = T;
        bool fit = true, fit_n = true;
        int cuse, cuse_n = 0;                                                  // the cell a lane works on (a stray's stand-in: see hw_tile_issue)
        {   // prologue: X of round 0 (waited for), then its tile, its F_E / constants and the X of round 1 in flight
            const int p0 = slot_of(hbase + s);
            cp_async16_stream(&W.x[0][lane], P.a[PX] + p0);
            cp_async_commit(); cp_async_wait_all();
            cuse = __float_as_int(W.x[0][lane].w);
            if (!LIST) hw_tile_issue(&tm, G, tile, bar, T, cuse, fit);
#pragma unroll
            for (int a = 0; a < 3; ++a) cp_async16_stream(&W.e[a][lane], P.a[PE0 + a] + p0);
            cp_async16_stream(&W.e[3][lane], P.a[PK] + p0);
            if (ROUNDS > 1) cp_async16_stream(&W.x[1][lane], P.a[PX] + slot_of(hbase + 16 + s));
            cp_async_commit();
        }
        AccRow acc; acc_zero(acc);
        int wcell = -1;
#pragma unroll 1
        for (int round = 0; round < ROUNDS; ++round) {
            const bool more = round + 1 < ROUNDS;
            const int q = hbase + round * 16 + s;
            const bool live = q < n && fit;                                 // tail lanes and deferred strays compute on a stand-in and write nothing:
            const int p = slot_of(q);                                       // the warp stays converged for the votes below
            if (!LIST) defer_append(D, q < n && !fit, q);
            unsigned starts = 0;
            {   // ---- phase A
                const float4 X = W.x[round & 1][lane];
                const int cell = cuse;
                int ci = cell_i(cell), cj = cell_j(cell), ck = cell_k(cell);
                Axis ax, ay, az;
                bool complete = axis_setup(ax, X.x, ci, G.nx, G.ihx);
                complete &= axis_setup(ay, X.y, cj, G.ny, G.ihy);
                complete &= axis_setup(az, X.z, ck, G.nz, G.ihz);
                float rx[4], ry[4], rz[4];                                          // x_i - x_p per axis: h (o - 1 - f)
#pragma unroll
                for (int o = 0; o < 4; ++o) { rx[o] = G.hx * ((float)(o - 1) - X.x); ry[o] = G.hy * ((float)(o - 1) - X.y); rz[o] = G.hz * ((float)(o - 1) - X.z); }
                float nrx[4];
#pragma unroll
                for (int o = 0; o < 4; ++o) nrx[o] = ax.N[o] * rx[o];
                G2PSums S;
#pragma unroll
                for (int i = 0; i < 3; ++i) { S.vc[i] = 0.f; S.va[i] = 0.f; }
#pragma unroll
                for (int i = 0; i < 9; ++i) { S.B[i] = 0.f; S.g[i] = 0.f; }
                S.smin = 1.0f;
                if (!LIST) {
                    mbar_wait(bar, tphase); tphase ^= 1u;
                    g2p_gather<2>(G, ax, ay, az, nrx, rx, ry, rz, tile, ax.n0 - T.ox0, S);
                    if (S.smin < 1.0f) g2p_stick_correction<2>(G, ax, ay, az, rx, ry, rz, tile, ax.n0 - T.ox0, S);
                } else {
                    g2p_gather<0>(G, ax, ay, az, nrx, rx, ry, rz, tile, 0, S);
                    if (S.smin < 1.0f) g2p_stick_correction<0>(G, ax, ay, az, rx, ry, rz, tile, 0, S);
                }
                cp_async_wait_all();                                            // F_E / constants of this round, X of the next
                __syncwarp();                                                   // every lane is done with the tile
                if (more) { cuse_n = __float_as_int(W.x[(round + 1) & 1][lane].w); if (!LIST) hw_tile_issue(&tm, G, tile, bar, Tn, cuse_n, fit_n); }
                const float4 e0 = W.e[0][lane], e1 = W.e[1][lane], e2 = W.e[2][lane], kk = W.e[3][lane];
                float (&va)[3] = S.va; float (&B)[9] = S.B; float (&g)[9] = S.g;
                const float vp[3] = { S.va[0] + S.vc[0], S.va[1] + S.vc[1], S.va[2] + S.vc[2] };       // sum w s v~ = sum w v~ - sum w (1-s) v~
                // ---- advection (HybridSolver.cpp:944): x' = sum w (x_i + dt v~_i) = x + [sum w (x_i - x)] + (sum w - 1) x + dt sum w v~
                // the bracket and (sum w - 1) vanish unless the stencil is truncated by the domain boundary (:44-46); both are separable.
                float dxp = dt * va[0], dyp = dt * va[1], dzp = dt * va[2];
                if (!complete) {
                    const float sx = ax.N[0] + ax.N[1] + ax.N[2] + ax.N[3], sy = ay.N[0] + ay.N[1] + ay.N[2] + ay.N[3], sz = az.N[0] + az.N[1] + az.N[2] + az.N[3];
                    const float mx = nrx[0] + nrx[1] + nrx[2] + nrx[3];
                    const float my = ay.N[0] * ry[0] + ay.N[1] * ry[1] + ay.N[2] * ry[2] + ay.N[3] * ry[3];
                    const float mz = az.N[0] * rz[0] + az.N[1] * rz[1] + az.N[2] * rz[2] + az.N[3] * rz[3];
                    const float s0 = sx * sy * sz;
                    const float xw = fmaf((float)ci + X.x, G.hx, G.mnx), yw = fmaf((float)cj + X.y, G.hy, G.mny), zw = fmaf((float)ck + X.z, G.hz, G.mnz);
                    dxp += mx * sy * sz + (s0 - 1.0f) * xw; dyp += sx * my * sz + (s0 - 1.0f) * yw; dzp += sx * sy * mz + (s0 - 1.0f) * zw;
                }
                float nfx = fmaf(dxp, G.ihx, X.x), nfy = fmaf(dyp, G.ihy, X.y), nfz = fmaf(dzp, G.ihz, X.z);
                {
                    const float flx = floorf(nfx), fly = floorf(nfy), flz = floorf(nfz);
                    nfx -= flx; nfy -= fly; nfz -= flz; ci += (int)flx; cj += (int)fly; ck += (int)flz;
                    nfx = fminf(nfx, 0.99999994f); nfy = fminf(nfy, 0.99999994f); nfz = fminf(nfz, 0.99999994f);
                    const int cci = clampi(ci, 0, G.nx - 1), ccj = clampi(cj, 0, G.ny - 1), cck = clampi(ck, 0, G.nz - 1);
                    const bool nan = !(nfx == nfx) || !(nfy == nfy) || !(nfz == nfz);
                    if (cci != ci || ccj != cj || cck != ck || nan) {             // counted; the host turns a non-zero count into AEP_ERR_STATE
                        if (live && kk.x != 0.0f) atomicAdd(&clk->escaped, 1ull);       // (dead slots of a slab context are massless tracers)
                        if (!(nfx == nfx)) nfx = 0.5f;
                        if (!(nfy == nfy)) nfy = 0.5f;
                        if (!(nfz == nfz)) nfz = 0.5f;
                    }
                    ci = cci; cj = ccj; ck = cck;
                }
                // ---- deformation gradient + plasticity
                if (more) {                                                     // this lane's slots are consumed: next round's F_E / constants, X of the round after
                    const int pn = slot_of(hbase + (round + 1) * 16 + s);
#pragma unroll
                    for (int a = 0; a < 3; ++a) cp_async16_stream(&W.e[a][lane], P.a[PE0 + a] + pn);
                    cp_async16_stream(&W.e[3][lane], P.a[PK] + pn);
                    if (round + 2 < ROUNDS) cp_async16_stream(&W.x[round & 1][lane], P.a[PX] + slot_of(hbase + (round + 2) * 16 + s));
                    cp_async_commit();
                }
                float FE[9] = { e0.x, e0.y, e0.z, e1.x, e1.y, e1.z, e2.x, e2.y, e2.z };
                float GF[9], Fh[9];
                mat_mul(g, FE, GF);
#pragma unroll
                for (int i = 0; i < 9; ++i) Fh[i] = fmaf(dt, GF[i], FE[i]);                  // HybridSolver.cpp:575
                float qq = e1.w, Jp = e2.w;
                {
                    Sym3 E;
                    bool yields = false;
                    float FP[9];
                    auto load_FP = [&]() {
                        const float4 q0 = ldg4(P.a[PQ0] + p), q1 = ldg4(P.a[PQ1] + p), q2 = ldg4(P.a[PQ2] + p);
                        FP[0] = q0.x; FP[1] = q0.y; FP[2] = q0.z; FP[3] = q1.x; FP[4] = q1.y; FP[5] = q1.z; FP[6] = q2.x; FP[7] = q2.y; FP[8] = q2.z;
                    };
                    float e2 = 1.0f;
                    if (mpar.material != 0 && (e2 = left_cauchy_green_minus_one(Fh, E)) < AEP_SMALL_E2) {     // sand, small strain: no SVD (aep_math.cuh)
                        Sym3 M;
                        if (sand_project_small(mpar, E, e2, qq, M) && live) { load_FP(); sand_apply_small(M, Fh, FE, FP); yields = true; }
                    } else {
                        Svd3 sv; float sn[3];
                        if (return_map_project(mpar, Fh, sv, sn, qq) && live) { load_FP(); return_map_apply(sv, sn, Fh, FE, FP); yields = true; }
                    }
                    // F_P is touched by yielding particles only: 48 B read + 48 B written that a particle at rest never pays.  (Fetching the
                    // next round's rows ahead with cp.async while a warp's particles yield was tried: 21.06 against 20.91 ms per substep
                    // in the flowing state, profiles/README.md -- the three dependent loads hide behind the other warps already.)
                    if (yields) {                                                             // yielding particle: F_P changes
                        Jp = mat_det(FP);
                        P.a[PQ0][p] = make_float4(FP[0], FP[1], FP[2], 0.f);
                        P.a[PQ1][p] = make_float4(FP[3], FP[4], FP[5], 0.f);
                        P.a[PQ2][p] = make_float4(FP[6], FP[7], FP[8], 0.f);
                    } else {
#pragma unroll
                        for (int i = 0; i < 9; ++i) FE[i] = Fh[i];
                    }
                }
                // ---- write back
                const int ncell = cell_pack(ci, cj, ck);
                {   // particles that left their cell are out of order until the next physical sort
                    const unsigned mv = __ballot_sync(0xffffffffu, live && ncell != cell);
                    if (lane == 0 && mv) atomicAdd(&clk->moved_since_sort, (unsigned long long)__popc(mv));
                }
                const float4 Xn = make_float4(nfx, nfy, nfz, __int_as_float(ncell));
                if (live) {
                    P.a[PX][p] = Xn;
                    P.a[PV][p] = make_float4(vp[0], vp[1], vp[2], B[0]);
                    P.a[PC0][p] = make_float4(B[1], B[2], B[3], B[4]);
                    P.a[PC1][p] = make_float4(B[5], B[6], B[7], B[8]);
                    P.a[PE0][p] = make_float4(FE[0], FE[1], FE[2], e0.w);
                    P.a[PE1][p] = make_float4(FE[3], FE[4], FE[5], qq);
                    P.a[PE2][p] = make_float4(FE[6], FE[7], FE[8], Jp);
                    if (ML.axis >= 0 && kk.x != 0.0f) {                                     // live particle of a slab context: did it leave the slab?
                        const int ca = ML.axis == 0 ? ci : (ML.axis == 1 ? cj : ck);
                        if (ca < ML.lo || ca >= ML.hi) {
                            const int side = ca < ML.lo ? 0 : 1;
                            const unsigned long long slot_ = atomicAdd(ML.counts + side, 1ull);
                            if (slot_ < (unsigned long long)ML.cap) (side == 0 ? ML.list[0] : ML.list[1])[slot_] = (unsigned)p;
                        }
                    }
                }
                if (SCATTER) {   // record of the advected particle for the P2G of the next substep, straight from registers
                    const float m = live ? kk.x : 0.0f;
                    int prev;
                    starts = run_starts(ncell, wcell, prev);
                    p2g_make_record(W.rec[hw] + s * P2G_STRIDE, Xn, make_float4(vp[0], vp[1], vp[2], m), make_float4(B[0], B[1], B[2], 0.f),
                                    make_float4(B[3], B[4], B[5], 0.f), make_float4(B[6], B[7], B[8], 0.f), m, G.apic, G.hx, G.hy, G.hz, __int_as_float(prev));
                }
                T = Tn; fit = fit_n; cuse = cuse_n;
            }
            if (SCATTER) {
                __syncwarp();
                // ---- phase B: half-warp per particle, lane = (j,k) row of the stencil, 4 nodes along x in packed accumulators
                p2g_phase_b(G, W.rec[hw], starts, slot, j, k, yoff, zoff, J, K, acc);
            }
        }
        if (SCATTER && wcell >= 0) flush_row_pk(G, G.mp, slot, wcell, j, k, acc, true);
        if (!LIST) break;
        __syncwarp();
    }
}

template <int ROUNDS, bool SCATTER>
inline void g2p2g_launch_r(cudaStream_t st, const PartP& P, const GridP& G, const CUtensorMap& tm, const MatParams& mat, SimClock* clk,
                           long long n_hi, const MigList& ML, const DeferP& D) {
    const int smem = (int)sizeof(G2GWarpSmemT<SCATTER>) * (G2G_NT / 32);
    const long long per_cta = (long long)G2G_NT * ROUNDS;
    const int chunks = (int)((n_hi + per_cta - 1) / per_cta);
    k_g2p2g<ROUNDS, SCATTER, false><<<SCATTER ? strided_grid(chunks, G.strips) : chunks, G2G_NT, smem, st>>>(P, G, tm, mat, clk, ML, D);
}
// main kernel (zeroes the count of the deferred list first, in stream order) and list pass, separately (timed one by one when profiling)
template <bool SCATTER>
inline cudaError_t g2p2g_main_launch(cudaStream_t st, const PartP& P, const GridP& G, const CUtensorMap& tm, const MatParams& mat, SimClock* clk,
                                     long long n_hi, const MigList& ML, const DeferP& D) {
    cudaError_t e = cudaMemsetAsync(D.count, 0, sizeof(unsigned int), st);
    if (e != cudaSuccess) return e;
    switch (rounds_for(n_hi)) {
    case 8: g2p2g_launch_r<8, SCATTER>(st, P, G, tm, mat, clk, n_hi, ML, D); break;
    case 2: g2p2g_launch_r<2, SCATTER>(st, P, G, tm, mat, clk, n_hi, ML, D); break;
    default: g2p2g_launch_r<1, SCATTER>(st, P, G, tm, mat, clk, n_hi, ML, D); break;
    }
    return cudaGetLastError();
}
template <bool SCATTER>
inline cudaError_t g2p2g_list_launch(cudaStream_t st, const PartP& P, const GridP& G, const CUtensorMap& tm, const MatParams& mat, SimClock* clk,
                                     long long n_hi, const MigList& ML, const DeferP& D) {
    const long long list_ctas = std::min<long long>((n_hi + G2G_NT - 1) / G2G_NT, AEP_LIST_CTAS);
    k_g2p2g<1, SCATTER, true><<<(unsigned)list_ctas, G2G_NT, (int)sizeof(G2GWarpSmemT<SCATTER>) * (G2G_NT / 32), st>>>(P, G, tm, mat, clk, ML, D);
    return cudaGetLastError();
}

// dynamic shared memory above 48 KB needs an opt-in per kernel and device: aep_create calls this once
inline cudaError_t particle_kernels_configure() {
    const int fs = (int)sizeof(FrcWarpSmemT<false>) * (FRC_NT / 32), gs = (int)sizeof(G2GWarpSmemT<true>) * (G2G_NT / 32);
    cudaError_t e;
#define AEP_CFG(kern, bytes) if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes)) != cudaSuccess) return e
    AEP_CFG((k_forces<8, false, false>), fs); AEP_CFG((k_forces<2, false, false>), fs); AEP_CFG((k_forces<1, false, false>), fs); AEP_CFG((k_forces<1, true, false>), fs);
    AEP_CFG((k_forces<8, false, true>), fs); AEP_CFG((k_forces<2, false, true>), fs); AEP_CFG((k_forces<1, false, true>), fs); AEP_CFG((k_forces<1, true, true>), fs);
    AEP_CFG((k_g2p2g<8, true, false>), gs); AEP_CFG((k_g2p2g<2, true, false>), gs); AEP_CFG((k_g2p2g<1, true, false>), gs); AEP_CFG((k_g2p2g<1, true, true>), gs);
    AEP_CFG((k_g2p2g<8, false, false>), gs); AEP_CFG((k_g2p2g<2, false, false>), gs); AEP_CFG((k_g2p2g<1, false, false>), gs); AEP_CFG((k_g2p2g<1, false, true>), gs);
    AEP_CFG(k_force_scatter<8>, FRC_SCATTER_SMEM); AEP_CFG(k_force_scatter<2>, FRC_SCATTER_SMEM); AEP_CFG(k_force_scatter<1>, FRC_SCATTER_SMEM);
#undef AEP_CFG
    return cudaSuccess;
}

}  // namespace aep
