// aep_kernels.cuh -- sm_100a kernels of the MPM substep (particles: sand / snow).
//
// Device data layout (all fp32, 16-byte records so every particle access is one LDG/STG.128):
//   particle arrays (struct of float4 arrays, index = slot in cell-sorted order)
//     X  = (fx, fy, fz, cell)     fractional position inside the cell, packed cell (i | j<<10 | k<<20)
//     VM = (vx, vy, vz, m)
//     C0..C2 = rows of the APIC matrix B                                   (.w unused)
//     E0 = (FE row 0, vol)  E1 = (FE row 1, q)  E2 = (FE row 2, det F_P)
//     Q0 = (FP row 0, id)   Q1 = (FP row 1, m)  Q2 = (FP row 2, -)
//   grid arrays, node index = (k*ny + j)*nx + i  (RegularGrid.cpp:164-168)
//     mp = (m, px, py, pz)   f = (fx, fy, fz, -)   vt = (v~x, v~y, v~z, s)  with v = s * v~  (s in {0,1})
//   block flags: one byte per 8x8x8 node block, set by P2G; grid passes run over the compact list of flagged blocks.
//
// Scatter strategy: shared-memory float atomics are CAS loops on sm_100 (ATOMS.CAST.SPIN) while global memory has native vector
// reductions (REDG.E.ADD.F32x4).  So nothing is accumulated in shared memory: a half-warp walks a run of cell-sorted particles,
// its 16 lanes own the 16 (j,k) rows of the 4x4x4 stencil and keep the 4 nodes of their row in registers while the particles
// stay in one cell (and slide the 4-node window when the next cell is an x-neighbour); a node goes to L2 with one REDG.F32x4 when
// it leaves the window.  Gather strategy: the (cells+3) x 4 x 4 node box of a warp's 32 particles is staged once in the warp's
// own shared memory and every lane reads its 64 nodes from there.
#pragma once
#include <stdint.h>
#include "aep_math.cuh"
#include "aep_pack.cuh"
#include "aep_scatter.cuh"
#include "aep_gather.cuh"

namespace aep {

enum { PX = 0, PVM, PC0, PC1, PC2, PE0, PE1, PE2, PQ0, PQ1, PQ2, P_NARR };

struct PartP {
    float4* a[P_NARR];
};

// simulation clock + reductions, lives in device memory so that n substeps need no host round trip
struct SimClock {
    double t, inner_t, frame_dt, cfl, rate_floor, hmin;
    float dt;                          // dt used by forces / grid update (previous iteration's value, HybridSolver.cpp:873-878)
    unsigned int vmax_bits;            // max |v_i| of the last grid update, as float bits (non-negative -> ordered as uint)
    int frame_flag, frame_no;
    long long substeps;
    unsigned long long escaped;        // particles that tried to leave the grid (clamped), sticky
    float vmax_last;
    int pad;
    // adaptive re-sort (aep_config.sort_every == 0): particles that changed cell since the last physical sort, and the accumulated
    // per-substep out-of-order fraction.  Reset together (16 bytes) when the particles are re-sorted.
    unsigned long long moved_since_sort;
    float sort_cost;
    int pad2;
};

// slab decomposition: G2P appends the slots of particles whose new cell left [lo, hi) along `axis` to per-side index lists, so
// that the migration step touches only the leavers instead of scanning every particle (axis < 0: not a slab / lists not bound)
struct MigList {
    int axis, lo, hi, cap;
    unsigned int* list[2];
    unsigned long long* counts;        // [2] (low, high), caller-owned device memory
};

__device__ __forceinline__ int cell_i(int c) { return c & 1023; }
__device__ __forceinline__ int cell_j(int c) { return (c >> 10) & 1023; }
__device__ __forceinline__ int cell_k(int c) { return (c >> 20) & 1023; }
__device__ __forceinline__ int cell_pack(int i, int j, int k) { return i | (j << 10) | (k << 20); }

__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// CTA index of a grid pass -> block coordinates inside the run range, and the flat block index (flags)
__device__ __forceinline__ int run_block(const GridP& G, int r, int& bx, int& by, int& bz) {
    bx = G.rb0[0] + r % G.rbn[0]; by = G.rb0[1] + (r / G.rbn[0]) % G.rbn[1]; bz = G.rb0[2] + r / (G.rbn[0] * G.rbn[1]);
    return (bz * G.nby + by) * G.nbx + bx;
}

// ================================================================================================ sort keys
// Particles are ordered by 4x4x4-cell brick (x fastest), then by cell inside the brick (x fastest).  A CTA of 256 consecutive
// particles at 8 per cell is then a 4x4x2 slab of cells whose stencils share a 7x7x5 node box (245 nodes, 3.9 KB) instead of the
// 35x4x4 = 560 nodes of 32 cells in a row: the grid gathers of one CTA hit in L1 instead of going to L2.  The scatters only
// need particles of one cell to be adjacent, which any cell-granular order provides.
__device__ __forceinline__ unsigned sort_key(int ci, int cj, int ck, const GridP& G) {
    if (!G.bricks) return (unsigned)((ck * G.ny + cj) * G.nx + ci);
    const unsigned brick = (unsigned)(((ck >> 2) * G.nqy + (cj >> 2)) * G.nqx + (ci >> 2));
    return (brick << 6) | (unsigned)(((ck & 3) << 4) | ((cj & 3) << 2) | (ci & 3));
}
__global__ void k_build_keys(const float4* __restrict__ X, unsigned int* __restrict__ keys, unsigned int* __restrict__ vals,
                             int n, GridP G) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int c = __float_as_int(X[i].w);
    keys[i] = sort_key(cell_i(c), cell_j(c), cell_k(c), G);
    vals[i] = (unsigned)i;
}

// gather all particle arrays through the sort permutation (dst[i] = src[perm[i]])
__global__ void __launch_bounds__(256) k_reorder(PartP src, PartP dst, const unsigned int* __restrict__ perm, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned s = perm[i];
    float4 r[P_NARR];
#pragma unroll
    for (int a = 0; a < P_NARR; ++a) r[a] = ldg4(src.a[a] + s);
#pragma unroll
    for (int a = 0; a < P_NARR; ++a) dst.a[a][i] = r[a];
}

// ================================================================================================ grid passes
// Compact list of the flagged 8^3 blocks (run indices, see run_block).  The dam break occupies 7 % of the 512^3 grid: launching
// one CTA per block of the whole grid cost each pass ~0.2 ms of CTAs that only read a zero flag (0.77 ms of a 16.4 ms substep for
// the three passes); with the list a pass is a persistent grid over the occupied blocks and runs at memory speed.
__global__ void __launch_bounds__(256) k_list_blocks(GridP G, int nrun, unsigned int* __restrict__ list, unsigned int* __restrict__ count) {
    const int r = blockIdx.x * 256 + threadIdx.x;
    bool on = false;
    if (r < nrun) { int bx, by, bz; on = G.flags[run_block(G, r, bx, by, bz)] != 0; }
    const unsigned m = __ballot_sync(0xffffffffu, on);
    if (!m) return;
    const int lane = threadIdx.x & 31;
    unsigned base = 0;
    if (lane == 0) base = atomicAdd(count, (unsigned)__popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (on) list[base + __popc(m & ((1u << lane) - 1u))] = (unsigned)r;
}

// zero (m,p) and f of every block that the previous P2G touched, and drop its flag
__global__ void __launch_bounds__(256) k_clear_blocks(GridP G, const unsigned int* __restrict__ list, const unsigned int* __restrict__ count) {
    const unsigned nb = *count;
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    for (unsigned q = blockIdx.x; q < nb; q += gridDim.x) {
        int bx, by, bz;
        const int b = run_block(G, (int)list[q], bx, by, bz);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int t = threadIdx.x + 256 * h;
            const int i = bx * 8 + (t & 7), j = by * 8 + ((t >> 3) & 7), k = bz * 8 + (t >> 6);
            if (i < G.nx && j < G.ny && k < G.nz) {
                const size_t n = ((size_t)k * G.ny + j) * G.nx + i;
                G.mp[n] = z; G.f[n] = z;
            }
        }
        if (threadIdx.x == 0) G.flags[b] = 0;
    }
}

__device__ __forceinline__ void block_max_to_clock(float v, SimClock* clk) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    __shared__ float red[8];
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) red[w] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        float m = red[0];
        for (int i = 1; i < (int)(blockDim.x >> 5); ++i) m = fmaxf(m, red[i]);
        if (m > 0.0f) atomicMax(&clk->vmax_bits, __float_as_uint(m));
    }
}

// max |p/m| over active nodes: RegularGrid::CFL_condition for the initial dt (HybridSolver.cpp:860)
__global__ void __launch_bounds__(256) k_vmax_from_mp(GridP G, SimClock* clk) {
    int bx, by, bz;
    const int b = run_block(G, blockIdx.x, bx, by, bz);
    if (!G.flags[b]) return;
    float vm = 0.0f;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int t = threadIdx.x + 256 * h;
        const int i = bx * 8 + (t & 7), j = by * 8 + ((t >> 3) & 7), k = bz * 8 + (t >> 6);
        if (i < G.nx && j < G.ny && k < G.nz) {
            const float4 mp = G.mp[((size_t)k * G.ny + j) * G.nx + i];
            if (mp.x > 0.0f) {
                const float im = 1.0f / mp.x;
                const float vx = mp.y * im, vy = mp.z * im, vz = mp.w * im;
                vm = fmaxf(vm, sqrtf(vx * vx + vy * vy + vz * vz));
            }
        }
    }
    block_max_to_clock(vm, clk);
}

// updateGridVelocities_ (HybridSolver.cpp:725-737) + gravity (:457) + max|v| (RegularGrid.cpp:188-200)
// + gridCollisionHandling_ level-set part (:467-511), one coalesced float4 pass over the active blocks.
__global__ void __launch_bounds__(256) k_grid_update(GridP G, SimClock* clk, const unsigned int* __restrict__ list, const unsigned int* __restrict__ count) {
    const unsigned nb = *count;
    const float dt = clk->dt;
    float vm = 0.0f;
    for (unsigned q = blockIdx.x; q < nb; q += gridDim.x) {
    int bx, by, bz;
    run_block(G, (int)list[q], bx, by, bz);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int t = threadIdx.x + 256 * h;
        const int i = bx * 8 + (t & 7), j = by * 8 + ((t >> 3) & 7), k = bz * 8 + (t >> 6);
        if (i < G.nx && j < G.ny && k < G.nz) {
            const size_t n = ((size_t)k * G.ny + j) * G.nx + i;
            const float4 mp = G.mp[n];
            float vx = 0.f, vy = 0.f, vz = 0.f, s = 1.0f;
            if (mp.x > 0.0f) {
                const float4 f = G.f[n];
                const float im = 1.0f / mp.x;
                vx = fmaf(dt, f.x * im, mp.y * im);
                vy = fmaf(dt, f.y * im, mp.z * im);
                vz = fmaf(dt, fmaf(-G.gravity, mp.x, f.z) * im, mp.w * im);
                vm = fmaxf(vm, sqrtf(vx * vx + vy * vy + vz * vz));
                const int code = G.ls_code ? G.ls_code[n] : 0;
                if (code) {
                    float nx_, ny_, nz_;
                    if (code == 7) { const float4 nn = G.ls_nrm[n]; nx_ = nn.x; ny_ = nn.y; nz_ = nn.z; }
                    else {
                        nx_ = (code == 1) ? 1.f : (code == 2 ? -1.f : 0.f);
                        ny_ = (code == 3) ? 1.f : (code == 4 ? -1.f : 0.f);
                        nz_ = (code == 5) ? 1.f : (code == 6 ? -1.f : 0.f);
                    }
                    const float vn = vx * nx_ + vy * ny_ + vz * nz_;                 // HybridSolver.cpp:486
                    if (vn < 0.0f) {                                                 // :488 approaching
                        vx = fmaf(-vn, nx_, vx); vy = fmaf(-vn, ny_, vy); vz = fmaf(-vn, nz_, vz);   // :490-492
                        const float vt = sqrtf(vx * vx + vy * vy + vz * vz);
                        // :494-502 stick test; the Coulomb reduction at :501 is a no-op expression statement
                        s = (vt < -G.friction * vn) ? 0.0f : 1.0f;
                    }
                }
            }
            G.vt[n] = make_float4(vx, vy, vz, s);
        }
    }
    }
    block_max_to_clock(vm, clk);
}

// dt rule + frame clipping of HybridSolver.cpp:878-892, one thread, all in double like the reference
__global__ void k_advance_clock(SimClock* clk, int fixed_dt, int n) {
    const float vmax = __uint_as_float(clk->vmax_bits);
    clk->vmax_last = vmax; clk->vmax_bits = 0u;
    if (fixed_dt == 2) return;                                   // stage-level API: only latch max|v|
    clk->sort_cost += (float)clk->moved_since_sort / (float)max(n, 1);
    if (fixed_dt == 1) {                                         // pinned dt (aep_set_fixed_dt): plain time accumulation
        clk->inner_t += (double)clk->dt; clk->frame_flag = 0;
        if (clk->inner_t >= clk->frame_dt) { clk->inner_t -= clk->frame_dt; clk->t += clk->frame_dt; clk->frame_flag = 1; clk->frame_no += 1; }
        clk->substeps += 1;
        return;
    }
    double dt = clk->cfl / fmax(clk->rate_floor, (double)vmax / clk->hmin);
    if (clk->inner_t + dt >= clk->frame_dt) {
        dt = clk->frame_dt - clk->inner_t; clk->t += clk->frame_dt; clk->inner_t = 0.0; clk->frame_flag = 1; clk->frame_no += 1;
    } else {
        clk->inner_t += dt; clk->frame_flag = 0;
    }
    clk->dt = (float)dt;
    clk->substeps += 1;
}
__global__ void k_initial_dt(SimClock* clk) {                      // HybridSolver.cpp:860
    const float vmax = __uint_as_float(clk->vmax_bits);
    clk->vmax_last = vmax; clk->vmax_bits = 0u;
    clk->dt = (float)(clk->cfl / fmax(clk->rate_floor, (double)vmax / clk->hmin));
}

// ================================================================================================ scatter machinery
// Both scatters (P2G, force) run in two phases per warp of 32 cell-sorted particles:
//   phase A  thread-per-particle: everything that depends on the particle only (1-D weights, affine / stress matrices)
//            goes to a per-warp shared-memory record;
//   phase B  one HALF-WARP per particle, 16 lanes = the 16 (j,k) rows of the 4x4x4 stencil, each lane owns the 4 nodes
//            along x of its row and accumulates in registers (packed fp32x2, aep_pack.cuh) over the run of particles that share
//            a cell.  No atomics and no shuffles in the inner loop.
// Sliding window along x: when the next particle's cell is d = 1..3 cells further along x in the same (j,k) row of cells, only the
// d nodes that fall out of the 4-node window are reduced to memory and the accumulators shift; x-adjacent cells share 3 of their
// 4 nodes per row, so a sorted row of C cells costs C+3 reductions per lane instead of 4C.  slide_row_pk returns false when the window
// cannot slide (the caller then flushes all four nodes with flush_row_pk).

// CTA -> chunk of the sorted particle order for the scatter kernels.  Consecutive CTAs run concurrently; if they also worked on
// consecutive chunks, neighbouring rows and planes of cells would reduce into the same grid nodes at the same time and the L2
// atomic units would serialise them (thin slabs, where a wave of CTAs spans several k-planes, lost 30% to that).  With S strips,
// CTA b takes chunk (b % S) * L + b / S, L = ceil(nchunks / S): concurrent CTAs sit L chunks apart.  Returns -1 past the end.
__device__ __forceinline__ int strided_chunk(int b, int nchunks, int strips) {
    if (strips <= 1) return b < nchunks ? b : -1;
    const int L = (nchunks + strips - 1) / strips;
    const int chunk = (b % strips) * L + b / strips;
    return (b / strips < L && chunk < nchunks) ? chunk : -1;
}
__host__ __device__ __forceinline__ int strided_grid(int nchunks, int strips) {
    if (strips <= 1) return nchunks;
    return ((nchunks + strips - 1) / strips) * strips;
}

// warp-cooperative v1 mapping (2 nodes per lane), still used by the cloth kernels where runs have length 1
__device__ __forceinline__ void flush_nodes(const GridP& G, float4* __restrict__ dst, int cell, int oi, int oj, int ok,
                                            const float4& a0, const float4& a1, bool mark) {
    const int ni = cell_i(cell) - 1 + oi, nj = cell_j(cell) - 1 + oj, nk0 = cell_k(cell) - 1 + ok, nk1 = nk0 + 2;
    const bool inij = (ni >= 0) && (ni < G.nx) && (nj >= 0) && (nj < G.ny);
    if (inij && nk0 >= 0 && nk0 < G.nz) {
        atomicAdd(dst + (((size_t)nk0 * G.ny + nj) * G.nx + ni), a0);
        if (mark) G.flags[((nk0 >> 3) * G.nby + (nj >> 3)) * G.nbx + (ni >> 3)] = 1;
    }
    if (inij && nk1 >= 0 && nk1 < G.nz) {
        atomicAdd(dst + (((size_t)nk1 * G.ny + nj) * G.nx + ni), a1);
        if (mark) G.flags[((nk1 >> 3) * G.nby + (nj >> 3)) * G.nbx + (ni >> 3)] = 1;
    }
}


// ---- packed accumulators: AccRow (aep_scatter.cuh)
// same, but pinned behind the reductions that consumed the old values (volatile): see slide_row_pk
__device__ __forceinline__ void acc_zero_ordered(AccRow& a) {
    asm volatile("mov.b64 %0, 0;\n\tmov.b64 %1, 0;\n\tmov.b64 %2, 0;\n\tmov.b64 %3, 0;" : "=l"(a.lo[0]), "=l"(a.lo[1]), "=l"(a.lo[2]), "=l"(a.lo[3]));
    asm volatile("mov.b64 %0, 0;\n\tmov.b64 %1, 0;\n\tmov.b64 %2, 0;\n\tmov.b64 %3, 0;" : "=l"(a.hi[0]), "=l"(a.hi[1]), "=l"(a.hi[2]), "=l"(a.hi[3]));
}
// The reduction wants its float4 in an aligned register quad while FFMA2 wants the accumulators in pairs that stay put across
// the loop; asked for both, ptxas keeps the quads and pays ~20 MOVs per particle on the hot path to shuffle the pairs.  So the
// (rare) flush bounces the pairs through a lane-private shared-memory slot: STS.64 x2 (pairs only) + LDS.128 into a fresh quad.
__device__ __forceinline__ float4 requad(float4* slot, f32x2 lo, f32x2 hi) {
    float4 v;
    const unsigned a = (unsigned)__cvta_generic_to_shared(slot);
    asm volatile("st.shared.b64 [%4], %5;\n\tst.shared.b64 [%4+8], %6;\n\tld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a), "l"(lo), "l"(hi) : "memory");
    return v;
}
__device__ __forceinline__ void flush_row_pk(const GridP& G, float4* __restrict__ dst, float4* slot, int cell, int j, int k, const AccRow& a, bool mark) {
    const int ni0 = cell_i(cell) - 1, nj = cell_j(cell) - 1 + j, nk = cell_k(cell) - 1 + k;
    if (nj < 0 || nj >= G.ny || nk < 0 || nk >= G.nz) return;
    float4* row = dst + ((size_t)nk * G.ny + nj) * G.nx;
    unsigned char* frow = G.flags + ((nk >> 3) * G.nby + (nj >> 3)) * G.nbx;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int ni = ni0 + i;
        if (ni >= 0 && ni < G.nx) {
            atomicAdd(row + ni, requad(slot, a.lo[i], a.hi[i]));
            // block flags: the first node of the window in the grid, and the last one when it sits in another 8-block
            if (mark && (i == 0 || ni == 0 || ((i == 3 || ni == G.nx - 1) && (ni >> 3) != (max(ni0, 0) >> 3)))) frow[ni >> 3] = 1;
        }
    }
}
// sliding window along x (see slide_row): reduce the d nodes that leave the window, shift the rest
__device__ __forceinline__ bool slide_row_pk(const GridP& G, float4* __restrict__ dst, float4* slot, int cur, int next, int j, int k, AccRow& a, bool mark) {
    const int d = next - cur;
    if (d <= 0 || d >= 4 || (next & 1023) - (cur & 1023) != d) return false;
    const int nj = cell_j(cur) - 1 + j, nk = cell_k(cur) - 1 + k;
    const bool in_jk = nj >= 0 && nj < G.ny && nk >= 0 && nk < G.nz;
    float4* row = dst + ((size_t)nk * G.ny + nj) * G.nx;
    unsigned char* frow = G.flags + ((nk >> 3) * G.nby + (nj >> 3)) * G.nbx;
    int ni = cell_i(cur) - 1;
    for (int s = 0; s < d; ++s, ++ni) {
        if (in_jk && ni >= 0 && ni < G.nx) {
            atomicAdd(row + ni, requad(slot, a.lo[0], a.hi[0]));
            // a node that slides out marks its block only when it is the block's last one: a block the window has left behind saw
            // its last node slide out, and the blocks under the window are marked by the flush that ends the run
            if (mark && (ni & 7) == 7) frow[ni >> 3] = 1;
        }
        // ordered moves (volatile: they stay behind the reduction that consumed node 0), so that the accumulators keep their
        // registers on the hot path and only this rare path pays for the shift
        asm volatile("mov.b64 %0, %4;\n\tmov.b64 %1, %5;\n\tmov.b64 %2, %6;\n\tmov.b64 %3, 0;"
                     : "=l"(a.lo[0]), "=l"(a.lo[1]), "=l"(a.lo[2]), "=l"(a.lo[3]) : "l"(a.lo[1]), "l"(a.lo[2]), "l"(a.lo[3]));
        asm volatile("mov.b64 %0, %4;\n\tmov.b64 %1, %5;\n\tmov.b64 %2, %6;\n\tmov.b64 %3, 0;"
                     : "=l"(a.hi[0]), "=l"(a.hi[1]), "=l"(a.hi[2]), "=l"(a.hi[3]) : "l"(a.hi[1]), "l"(a.hi[2]), "l"(a.hi[3]));
    }
    return true;
}

// ================================================================================================ P2G
// particleToGrid_ (HybridSolver.cpp:113-231): m_i = sum w m ; p_i = sum w m (v + (3/h^2) B (x_i - x_p)).
// Per particle the momentum of node offset (i,j,k) is q0 + Qm (i,j,k)^T with Qm = m (3/h^2) B diag(h), q0 = m v - Qm (1+f).
// Phase-A record of one particle: see p2g_make_record (aep_scatter.cuh).
// Phase B is bound by shared-memory wavefronts as much as by issue slots (profiles/README.md, v9): every LDS of the two half-warps
// costs one wavefront per half-warp, so the record holds plain scalars (FFMA2 takes a broadcast .F32 operand) and the end of a
// run of same-cell particles comes from a ballot of phase A instead of a per-particle load of the cell index.
#define P2G_HW_PAD 2
#define P2G_HW_F4 (16 * P2G_STRIDE + P2G_HW_PAD)
// bit l of the result: the run of same-cell particles ends with the particle of lane l
__device__ __forceinline__ unsigned run_ends(int cell) {
    const int lane = threadIdx.x & 31;
    const int nxt = __shfl_down_sync(0xffffffffu, cell, 1);
    return __ballot_sync(0xffffffffu, nxt != cell || (lane & 15) == 15);
}
// Every flush costs one LSU wavefront per lane and node (the 16 rows of a half-warp lie in 16 different cache lines), and P2G is
// bound by exactly those wavefronts plus the shared-memory reads (l1tex data-pipe 97 % busy, profiles/r1_v9c).  A half-warp
// therefore keeps its window open over P2G_ROUNDS x 16 consecutive particles instead of 16: with 8 particles per cell the
// reductions per lane drop from 5 per 16 particles to 19 per 128 (measured: 3.77 ms at 1 round, 3.05 at 4, 2.95 at 8, 2.91 at 16).  Round r of a warp loads, per half-warp h, the particles
// base + h*16*ROUNDS + r*16 + (lane & 15), so each half-warp walks one contiguous piece of the sorted order.
#ifndef P2G_ROUNDS
#define P2G_ROUNDS 8
#endif
#define P2G_CTA_PARTICLES (256 * P2G_ROUNDS)
__global__ void __launch_bounds__(256) k_p2g(PartP P, GridP G, int n) {
    __shared__ float4 stage[8][2][P2G_HW_F4];
    __shared__ float4 bounce[256];                                            // lane-private slots of requad()
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    float4* slot = bounce + threadIdx.x;
    const int chunk = strided_chunk(blockIdx.x, (n + P2G_CTA_PARTICLES - 1) / P2G_CTA_PARTICLES, G.strips);
    if (chunk < 0) return;
    const int hw = lane >> 4, s = lane & 15, j = s & 3, k = s >> 2;
    const int hbase = chunk * P2G_CTA_PARTICLES + wib * (32 * P2G_ROUNDS) + hw * (16 * P2G_ROUNDS);      // first particle of this half-warp
    if (chunk * P2G_CTA_PARTICLES + wib * (32 * P2G_ROUNDS) >= n) return;    // warp-uniform; no block-level barrier below
    const int hend = hbase + 16 * P2G_ROUNDS;                                // one past its last particle (may exceed n)
    float fj = (float)j, fk = (float)k;
    int yoff = 16 + 4 * j, zoff = 32 + 4 * k;                                 // byte offsets of Ny[j], Nz[k] inside a record
    // lane constants: ptxas re-derives them from %tid on every trip (8 instructions) unless they come out of something it cannot
    // rematerialise -- an identity shuffle
    fj = __shfl_sync(0xffffffffu, fj, lane); fk = __shfl_sync(0xffffffffu, fk, lane);
    yoff = __shfl_sync(0xffffffffu, yoff, lane); zoff = __shfl_sync(0xffffffffu, zoff, lane);
    const f32x2 J = pk1(fj), K = pk1(fk);
    AccRow acc; acc_zero(acc);
    const float4* recs = &stage[wib][hw][0];
#pragma unroll 1
    for (int round = 0; round < P2G_ROUNDS; ++round) {
        unsigned ends;
        {   // ---- phase A: thread per particle
            const int q = hbase + round * 16 + s;                            // may lie past the end: such lanes repeat the last particle with zero mass
            const int p = min(q, n - 1);
            const float4 X = ldg4(P.a[PX] + p), VM = ldg4(P.a[PVM] + p);
            const float4 c0 = ldg4(P.a[PC0] + p), c1 = ldg4(P.a[PC1] + p), c2 = ldg4(P.a[PC2] + p);
            // cell of the half-warp's next particle (-1 behind its last one): the lines are the ones the neighbouring lanes load
            const int ncell = (q + 1 < hend) ? __float_as_int(__ldg(&P.a[PX][min(q + 1, n - 1)].w)) : -1;
            const float m = (q < n) ? VM.w : 0.0f;
            __syncwarp();                                                    // phase B of the round before is done with the records
            p2g_make_record(&stage[wib][hw][s * P2G_STRIDE], X, VM, c0, c1, c2, m, G.apic, G.hx, G.hy, G.hz, __int_as_float(ncell));
            ends = __ballot_sync(0xffffffffu, ncell != __float_as_int(X.w));
        }
        __syncwarp();
        // ---- phase B: half-warp per particle, lane = (j,k) row of the stencil, 4 nodes along x in packed accumulators
        ends >>= hw * 16;
#pragma unroll 1
        for (int it = 0; it < 16; ++it) {
            const float4* r = recs + it * P2G_STRIDE;
            p2g_row_accumulate(r, yoff, zoff, J, K, acc);
            if ((ends >> it) & 1u) {                                          // the run of particles sharing this cell ends here
                const float2 cn = *reinterpret_cast<const float2*>(r + 7);
                const int cur = __float_as_int(cn.x), nxt = __float_as_int(cn.y);
                if (!slide_row_pk(G, G.mp, slot, cur, nxt, j, k, acc, true)) {
                    flush_row_pk(G, G.mp, slot, cur, j, k, acc, true);
                    acc_zero_ordered(acc);
                }
            }
        }
    }
}

// launch helper shared by the engine, the slab arrivals and the mesh transfers
inline void p2g_launch(cudaStream_t st, const PartP& P, const GridP& G, long long n) {
    const int chunks = (int)((n + P2G_CTA_PARTICLES - 1) / P2G_CTA_PARTICLES);
    k_p2g<<<strided_grid(chunks, G.strips), 256, 0, st>>>(P, G, (int)n);
}

// first P2G only: rho_p = sum_i w m_i / (hx hy hz), V_p = m_p / rho_p           HybridSolver.cpp:242-249
__global__ void __launch_bounds__(256) k_init_volumes(PartP P, GridP G, int n) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const float4 X = P.a[PX][p];
    const int cell = __float_as_int(X.w);
    Axis ax, ay, az;
    axis_setup(ax, X.x, cell_i(cell), G.nx, G.ihx); axis_setup(ay, X.y, cell_j(cell), G.ny, G.ihy); axis_setup(az, X.z, cell_k(cell), G.nz, G.ihz);
    float dens = 0.0f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int nk = clampi(az.n0 + k, 0, G.nz - 1);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int nj = clampi(ay.n0 + j, 0, G.ny - 1);
            const float wjk = ay.N[j] * az.N[k];
            const size_t row = ((size_t)nk * G.ny + nj) * G.nx;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int ni = clampi(ax.n0 + i, 0, G.nx - 1);
                dens = fmaf(ax.N[i] * wjk, ldg4(G.mp + row + ni).x, dens);
            }
        }
    }
    dens *= G.inv_cell_vol;
    const float m = P.a[PVM][p].w;
    float4 e0 = P.a[PE0][p]; e0.w = m * (1.0f / dens); P.a[PE0][p] = e0;
}

// v_i = p_i / m_i where m_i > 0 (HybridSolver.cpp:233-240) for every active node, into the vt array (free between P2G and
// the grid update), so that the force gather below does 64 loads and no divisions per particle.
__global__ void __launch_bounds__(256) k_grid_normalise(GridP G, const unsigned int* __restrict__ list, const unsigned int* __restrict__ count) {
    const unsigned nb = *count;
    for (unsigned q = blockIdx.x; q < nb; q += gridDim.x) {
    int bx, by, bz;
    run_block(G, (int)list[q], bx, by, bz);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int t = threadIdx.x + 256 * h;
        const int i = bx * 8 + (t & 7), j = by * 8 + ((t >> 3) & 7), k = bz * 8 + (t >> 6);
        if (i < G.nx && j < G.ny && k < G.nz) {
            const size_t n = ((size_t)k * G.ny + j) * G.nx + i;
            const float4 mp = G.mp[n];
            const float im = mp.x > 0.0f ? 1.0f / mp.x : 0.0f;
            G.vt[n] = make_float4(mp.y * im, mp.z * im, mp.w * im, 1.0f);
        }
    }
    }
}

// ================================================================================================ shared grid tile for the gathers
// A warp of 32 cell-sorted particles normally sits in one row of cells (same j,k; 4 cells along x at 8 particles per cell), so the
// 64-node stencils of its particles live in a (cells+3) x 4 x 4 node box.  The warp stages that box of G.vt in its own slice of
// shared memory with coalesced loads (all in flight at once) and every lane then gathers with LDS.128 instead of 64 dependent
// L1/L2 round trips.  Warp-private tiles need no block barrier, so the warps of a CTA drift apart and overlap each other's load
// and compute phases.  A warp whose particles do not fit the box (row wrap, sparse or unsorted particles, domain faces) takes
// the global-memory path.
#ifndef AEP_USE_TILE
#define AEP_USE_TILE 1
#endif
struct TileRef {
    int ox0, j0, k0;                    // node coordinates of tile[0]
};
// warp-uniform: decide whether the tile path applies and, if so, fill the warp's tile.  cell/complete describe the lane's particle.
__device__ __forceinline__ bool stage_tile(const GridP& G, float4* __restrict__ tile, TileRef& T, int cell, bool complete) {
    if (!AEP_USE_TILE) return false;
    const int lane = threadIdx.x & 31;
    const int cref = __shfl_sync(0xffffffffu, cell, 0);
    T.ox0 = cell_i(cref) - 1 - TILE_SLACK; T.j0 = cell_j(cref) - 1; T.k0 = cell_k(cref) - 1;
    const int ci = cell_i(cell);
    const bool fits = complete && ((cell ^ cref) >> 10) == 0 && (ci - 1) >= T.ox0 && (ci + 2) < T.ox0 + TILE_W;
    if (!__all_sync(0xffffffffu, fits)) return false;
    // rows j0..j0+3, k0..k0+3 are inside the grid because every particle of the warp is `complete`; x is clipped to the grid
#pragma unroll
    for (int q = 0; q < TILE_F4 / 32; ++q) {
        const int idx = lane + 32 * q;
        const int r = idx / TILE_W, x = idx - r * TILE_W, gx = T.ox0 + x;
        if (gx >= 0 && gx < G.nx) tile[idx] = ldg4(G.vt + ((size_t)(T.k0 + (r >> 2)) * G.ny + (T.j0 + (r & 3))) * G.nx + gx);
    }
    __syncwarp();
    return true;
}

// asynchronous version of stage_tile: the warp's grid tile is fetched with cp.async (no registers, no wait) for a chunk whose X
// records are already in registers, so that the copy flies behind the arithmetic of the chunk before.  Warp-uniform result.
__device__ __forceinline__ void cp_async16(float4* smem_dst, const float4* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async16_stream(float4* smem_dst, const float4* gsrc) {      // L2 only: streaming particle data stays out of L1
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }   // all but the N most recent groups
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ bool tile_issue(const GridP& G, float4* __restrict__ tile, TileRef& T, int cell) {
    if (!AEP_USE_TILE) return false;
    const int lane = threadIdx.x & 31;
    const int ci = cell_i(cell), cj = cell_j(cell), ck = cell_k(cell);
    const bool complete = ci >= 1 && ci + 2 < G.nx && cj >= 1 && cj + 2 < G.ny && ck >= 1 && ck + 2 < G.nz;     // == axis_setup's
    const int cref = __shfl_sync(0xffffffffu, cell, 0);
    T.ox0 = cell_i(cref) - 1 - TILE_SLACK; T.j0 = cell_j(cref) - 1; T.k0 = cell_k(cref) - 1;
    const bool fits = complete && ((cell ^ cref) >> 10) == 0 && (ci - 1) >= T.ox0 && (ci + 2) < T.ox0 + TILE_W;
    if (!__all_sync(0xffffffffu, fits)) return false;
#pragma unroll
    for (int q = 0; q < TILE_F4 / 32; ++q) {
        const int idx = lane + 32 * q;
        const int r = idx / TILE_W, x = idx - r * TILE_W, gx = T.ox0 + x;
        if (gx >= 0 && gx < G.nx) cp_async16(tile + idx, G.vt + ((size_t)(T.k0 + (r >> 2)) * G.ny + (T.j0 + (r & 3))) * G.nx + gx);
    }
    return true;
}

// ================================================================================================ forces
// computeGridForces_, particle part (HybridSolver.cpp:252-368).  Phase A (thread per particle): gather
// grad v = sum_i v_i (grad w_i)^T, Fhat = (I + dt grad v) FE, SVD, stress, A = V_p P FE^T.  Phase B (half-warp per particle):
// f_i -= A grad w_ip.
// Phase-A record of one particle for the force scatter: see frc_make_record (aep_scatter.cuh).
// The warp's grid tile of phase A lives in the same shared memory: it is dead once the gather is done.
#define FRC_NT 128
#define FRC_HW_PAD 2
#define FRC_WARP_F4 (2 * (16 * FRC_STRIDE + FRC_HW_PAD))
static_assert(FRC_WARP_F4 >= TILE_F4, "the gather tile is aliased onto the warp's record area");
#ifndef FRC_MIN_CTAS
#define FRC_MIN_CTAS 6
#endif
// One CTA per 128 particles.  A persistent-warp build with cp.async prefetch of X / F_E and an early tile request (the recipe that
// gained 7 % in k_g2p) was measured at 7.73 ms against 6.53 ms here: the tile shares its memory with the records of phase B, so
// it cannot be fetched ahead, and the loop costs more than the two remaining waits (profiles/README.md, v10d).
__global__ void __launch_bounds__(FRC_NT, FRC_MIN_CTAS) k_forces(PartP P, GridP G, MatParams mpar, const SimClock* __restrict__ clk, int n) {
    __shared__ float4 stage[FRC_NT / 32][FRC_WARP_F4];
    __shared__ float4 bounce[FRC_NT];                                          // lane-private slots of requad()
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    float4* slot = bounce + threadIdx.x;
    float4* tile = stage[wib];
    const int base = blockIdx.x * FRC_NT + wib * 32;        // consecutive chunks: the gather of phase A lives on L1/L2 locality (strided: 2% slower)
    if (base >= n) return;                                                       // warp-uniform; no block-level barrier below
    const int cnt = min(32, n - base);
    const float dt = clk->dt;
    unsigned ends;
    {   // ---- phase A
        const int p = base + min(lane, cnt - 1);
        const float4 X = ldg4(P.a[PX] + p);
        const float4 e0 = ldg4(P.a[PE0] + p), e1 = ldg4(P.a[PE1] + p), e2 = ldg4(P.a[PE2] + p);
        const int cell = __float_as_int(X.w);
        Axis ax, ay, az;
        bool complete = axis_setup(ax, X.x, cell_i(cell), G.nx, G.ihx);
        complete &= axis_setup(ay, X.y, cell_j(cell), G.ny, G.ihy);
        complete &= axis_setup(az, X.z, cell_k(cell), G.nz, G.ihz);
        float g[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        TileRef T;
        if (stage_tile(G, tile, T, cell, complete)) gather_grad<2>(G, ax, ay, az, tile, ax.n0 - T.ox0, g);
        else gather_grad<AEP_FALLBACK_MODE>(G, ax, ay, az, tile, 0, g);           // rare: row ends, freshly moved particles, domain faces
        const float FE[9] = { e0.x, e0.y, e0.z, e1.x, e1.y, e1.z, e2.x, e2.y, e2.z };
        float GF[9], Fh[9], A[9];
        mat_mul(g, FE, GF);
#pragma unroll
        for (int i = 0; i < 9; ++i) Fh[i] = fmaf(dt, GF[i], FE[i]);              // HybridSolver.cpp:306
        stress_times_FEt(mpar, Fh, FE, (lane < cnt) ? -e0.w : 0.0f, e2.w, A);    // A := -V_p P FE^T (sign of :356-366 folded in); padding lanes: zero volume
        __syncwarp();                                                            // every lane is done with the tile: reuse it for the records
        float4* rec = stage[wib] + (lane >> 4) * (16 * FRC_STRIDE + FRC_HW_PAD) + (lane & 15) * FRC_STRIDE;
        frc_make_record(rec, ax.N, ax.D, ay.N, ay.D, az.N, az.D, A, X.w);
        ends = run_ends(cell);
    }
    __syncwarp();
    // ---- phase B: f_i += A grad w_i  with  grad w_i = (Dx_i Ny Nz, Nx_i Dy Nz, Nx_i Ny Dz)   (HybridSolver.cpp:356-366)
    //      = Dx_i U + Nx_i V,  U = A[:,0] Ny Nz,  V = A[:,1] Dy Nz + A[:,2] Ny Dz  per (j,k) row
    const int hw = lane >> 4, s = lane & 15, j = s & 3, k = s >> 2;
    int yoff = 32 + 8 * j, zoff = 64 + 8 * k;                                 // byte offsets of (Ny,Dy)[j], (Nz,Dz)[k] inside a record
    asm volatile("" : "+r"(yoff), "+r"(zoff));                                // lane constants: keep them in registers
    ends >>= hw * 16;
    AccRow acc; acc_zero(acc);
    const float4* recs = stage[wib] + hw * (16 * FRC_STRIDE + FRC_HW_PAD);
#pragma unroll 1
    for (int it = 0; it < 16; ++it) {
        const float4* r = recs + it * FRC_STRIDE;
        frc_row_accumulate(r, yoff, zoff, acc);
        if ((ends >> it) & 1u) {                                              // the run of particles sharing this cell ends here
            const int cur = __float_as_int(r[9].x);
            const int nxt = (it == 15) ? -1 : __float_as_int(r[FRC_STRIDE + 9].x);
            if (!slide_row_pk(G, G.f, slot, cur, nxt, j, k, acc, false)) {
                flush_row_pk(G, G.f, slot, cur, j, k, acc, false);
                acc_zero_ordered(acc);
            }
        }
    }
}

inline void forces_launch(cudaStream_t st, int sm_count, const PartP& P, const GridP& G, const MatParams& mat, const SimClock* clk, long long n) {
    (void)sm_count;
    k_forces<<<(unsigned)((n + FRC_NT - 1) / FRC_NT), FRC_NT, 0, st>>>(P, G, mat, clk, (int)n);
}

// ================================================================================================ G2P
// Two builds of the G2P kernel: the default, persistent warps that software-pipeline every memory round trip of a chunk behind the
// arithmetic of the chunk before, and -DAEP_G2P_PIPE=0, one CTA per 128 particles with a register-staged tile.  While the
// kernel was 3800-4200 SASS instructions the pipelined form lost 5-8 % to instruction-cache misses and loop overhead
// (profiles/README.md, v9b / v9d); on the slimmed kernel (2700-3000 instructions) it wins 7 % (6.20 -> 5.76 ms, v10c).
#ifndef AEP_G2P_PIPE
#define AEP_G2P_PIPE 1
#endif
#if AEP_G2P_PIPE
// updateParticleVelocities_ (HybridSolver.cpp:739-745), updateAffineMomenta_ with damp 0 (:760-825, :908-917),
// advection x = sum w (x_i + dt v~_i) (:942-945), updateDeformationGradient_ (:553-578), updatePlasticity_
// (:612-681), all in registers, one thread per particle.  Writes the new sort key.
//
// Persistent warps: G2P_CTAS_PER_SM CTAs per SM, every warp walks chunks of 32 cell-sorted particles with a grid stride.  A chunk
// used to open with two dependent memory round trips (X from DRAM, then the grid tile whose position X decides) and to wait
// again for F_E / F_P after the gather: 20 % of the kernel's stall samples at 4 warps per scheduler
// (profiles/r1_v8_k_g2p_sass_summary.txt).  Everything a chunk needs now arrives by cp.async while the chunk before it computes --
// X two chunks ahead, F_E / F_P one chunk ahead, the grid tile as soon as the gather of the current chunk is done with the
// buffer -- into the warp's own shared memory, so no prefetched value ever occupies a register (a register-staged variant
// spilled the prefetched X at once and stalled on the spill store: profiles/README.md, v9b).
#define G2P_NT 128
#ifndef G2P_CTAS_PER_SM
#define G2P_CTAS_PER_SM 4
#endif
struct G2PWarpSmem {
    float4 tile[TILE_F4];
    float4 x[2][32];
    float4 eq[6][32];
};
__global__ void __launch_bounds__(G2P_NT, G2P_CTAS_PER_SM) k_g2p(PartP P, GridP G, MatParams mpar, SimClock* __restrict__ clk,
                                             unsigned int* __restrict__ keys, unsigned int* __restrict__ vals, int n, MigList ML) {
    __shared__ G2PWarpSmem wsm[G2P_NT / 32];
    const int lane = threadIdx.x & 31;
    G2PWarpSmem& W = wsm[threadIdx.x >> 5];
    float4* tile = W.tile;
    const int nchunks = (n + 31) >> 5, wstride = gridDim.x * (G2P_NT / 32);
    int chunk = blockIdx.x * (G2P_NT / 32) + (threadIdx.x >> 5);            // the warps of a CTA take neighbouring chunks (shared L1 lines)
    if (chunk >= nchunks) return;
    const float dt = clk->dt;
    TileRef T;
    bool use_tile;
    int buf = 0;
    {   // prologue: X of the first chunk (waited for), then its tile, its F_E / F_P and the X of the second chunk in flight
        const int p = min(chunk * 32 + lane, n - 1);
        cp_async16_stream(&W.x[0][lane], P.a[PX] + p);
        cp_async_wait_all();
        use_tile = tile_issue(G, tile, T, __float_as_int(W.x[0][lane].w));
        cp_async_commit();                                                  // group "tile"
#pragma unroll
        for (int a = 0; a < 6; ++a) cp_async16_stream(&W.eq[a][lane], P.a[PE0 + a] + p);
        if (chunk + wstride < nchunks) cp_async16_stream(&W.x[1][lane], P.a[PX] + min((chunk + wstride) * 32 + lane, n - 1));
        cp_async_commit();                                                  // group "particle data"
    }
    for (;;) {
        const int p_raw = chunk * 32 + lane;
        const bool live = p_raw < n;                                        // tail lanes recompute the last particle and write nothing:
        const int p = live ? p_raw : n - 1;                                 // the warp stays converged for the votes below
        const int next = chunk + wstride;
        const bool more = next < nchunks;                                   // warp-uniform
        cp_async_wait_group<1>();                                           // the tile has landed; this chunk's F_E / F_P and the next X may still fly
        __syncwarp();
        const float4 X = W.x[buf][lane];
        const int cell = __float_as_int(X.w);
        int ci = cell_i(cell), cj = cell_j(cell), ck = cell_k(cell);
        Axis ax, ay, az;
        bool complete = axis_setup(ax, X.x, ci, G.nx, G.ihx);
        complete &= axis_setup(ay, X.y, cj, G.ny, G.ihy);
        complete &= axis_setup(az, X.z, ck, G.nz, G.ihz);
        float rx[4], ry[4], rz[4];                                              // x_i - x_p per axis: h (o - 1 - f)
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
        if (use_tile) { g2p_gather<2>(G, ax, ay, az, nrx, rx, ry, rz, tile, ax.n0 - T.ox0, S); if (S.smin == 0.0f) g2p_stick_correction<2>(G, ax, ay, az, rx, ry, rz, tile, ax.n0 - T.ox0, S); }
        else { g2p_gather<0>(G, ax, ay, az, nrx, rx, ry, rz, tile, 0, S); if (S.smin == 0.0f) g2p_stick_correction<0>(G, ax, ay, az, rx, ry, rz, tile, 0, S); }
        cp_async_wait_group<0>();                                           // F_E / F_P of this chunk, X of the next
        __syncwarp();                                                       // every lane is done with the tile
        if (more) use_tile = tile_issue(G, tile, T, __float_as_int(W.x[buf ^ 1][lane].w));
        cp_async_commit();                                                  // group "tile" of the next chunk
        float (&va)[3] = S.va; float (&B)[9] = S.B; float (&g)[9] = S.g;
        const float vp[3] = { S.va[0] + S.vc[0], S.va[1] + S.vc[1], S.va[2] + S.vc[2] };       // sum w s v~ = sum w v~ - sum_{s=0} w v~
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
            if (cci != ci || ccj != cj || cck != ck || nan) {
                if (live) atomicAdd(&clk->escaped, 1ull);
                if (!(nfx == nfx)) nfx = 0.5f;
                if (!(nfy == nfy)) nfy = 0.5f;
                if (!(nfz == nfz)) nfz = 0.5f;
            }
            ci = cci; cj = ccj; ck = cck;
        }
        // ---- deformation gradient + plasticity
        const float4 e0 = W.eq[0][lane], e1 = W.eq[1][lane], e2 = W.eq[2][lane];
        const float4 q0 = W.eq[3][lane], q1 = W.eq[4][lane], q2 = W.eq[5][lane];
        float FE[9] = { e0.x, e0.y, e0.z, e1.x, e1.y, e1.z, e2.x, e2.y, e2.z };
        float FP[9] = { q0.x, q0.y, q0.z, q1.x, q1.y, q1.z, q2.x, q2.y, q2.z };
        float GF[9], Fh[9];
        mat_mul(g, FE, GF);
#pragma unroll
        for (int i = 0; i < 9; ++i) Fh[i] = fmaf(dt, GF[i], FE[i]);                  // HybridSolver.cpp:575
        float q = e1.w;
        return_map(mpar, Fh, FE, FP, q);
        const float Jp = mat_det(FP);
        // ---- write back
        const int ncell = cell_pack(ci, cj, ck);
        {   // particles that left their cell are out of order until the next physical sort
            const unsigned mv = __ballot_sync(0xffffffffu, live && ncell != cell);
            if ((threadIdx.x & 31) == 0 && mv) atomicAdd(&clk->moved_since_sort, (unsigned long long)__popc(mv));
        }
        if (live) {
            P.a[PX][p] = make_float4(nfx, nfy, nfz, __int_as_float(ncell));
            P.a[PVM][p] = make_float4(vp[0], vp[1], vp[2], q1.w);
            P.a[PC0][p] = make_float4(B[0], B[1], B[2], 0.f);
            P.a[PC1][p] = make_float4(B[3], B[4], B[5], 0.f);
            P.a[PC2][p] = make_float4(B[6], B[7], B[8], 0.f);
            P.a[PE0][p] = make_float4(FE[0], FE[1], FE[2], e0.w);
            P.a[PE1][p] = make_float4(FE[3], FE[4], FE[5], q);
            P.a[PE2][p] = make_float4(FE[6], FE[7], FE[8], Jp);
            P.a[PQ0][p] = make_float4(FP[0], FP[1], FP[2], q0.w);
            P.a[PQ1][p] = make_float4(FP[3], FP[4], FP[5], q1.w);
            P.a[PQ2][p] = make_float4(FP[6], FP[7], FP[8], q2.w);
            keys[p] = sort_key(ci, cj, ck, G);
            vals[p] = (unsigned)p;
            if (ML.axis >= 0 && q1.w != 0.0f) {                                     // live particle of a slab context: did it leave the slab?
                const int ca = ML.axis == 0 ? ci : (ML.axis == 1 ? cj : ck);
                if (ca < ML.lo || ca >= ML.hi) {
                    const int side = ca < ML.lo ? 0 : 1;
                    const unsigned long long slot = atomicAdd(ML.counts + side, 1ull);
                    if (slot < (unsigned long long)ML.cap) (side == 0 ? ML.list[0] : ML.list[1])[slot] = (unsigned)p;
                }
            }
        }
        if (!more) break;
        // F_E / F_P of the next chunk and X of the one after it: this lane's slots were consumed above
        {
            const int pn = min(next * 32 + lane, n - 1);
#pragma unroll
            for (int a = 0; a < 6; ++a) cp_async16_stream(&W.eq[a][lane], P.a[PE0 + a] + pn);
            if (next + wstride < nchunks) cp_async16_stream(&W.x[buf][lane], P.a[PX] + min((next + wstride) * 32 + lane, n - 1));
            cp_async_commit();                                              // group "particle data" of the next chunk
        }
        chunk = next; buf ^= 1;
    }
}

inline void g2p_launch(cudaStream_t st, int sm_count, const PartP& P, const GridP& G, const MatParams& mat, SimClock* clk, unsigned int* keys,
                       unsigned int* vals, long long n, const MigList& ML) {
    const long long chunks = (n + 31) / 32, per_cta = G2P_NT / 32;
    const int grid = (int)std::min<long long>((chunks + per_cta - 1) / per_cta, (long long)sm_count * G2P_CTAS_PER_SM);
    k_g2p<<<grid, G2P_NT, 0, st>>>(P, G, mat, clk, keys, vals, (int)n, ML);
}

#else
// updateParticleVelocities_ (HybridSolver.cpp:739-745), updateAffineMomenta_ with damp 0 (:760-825, :908-917),
// advection x = sum w (x_i + dt v~_i) (:942-945), updateDeformationGradient_ (:553-578), updatePlasticity_
// (:612-681), all in registers, one thread per particle.  The 64-node gather sums along x first (per (j,k) row:
// a = sum v~ Nx, b = sum v~ Dx, c = sum s v~ Nx, d = sum s v~ Nx rx), then combines rows.  Writes the new sort key.
#define G2P_NT 128
__global__ void __launch_bounds__(G2P_NT, 4) k_g2p(PartP P, GridP G, MatParams mpar, SimClock* __restrict__ clk,
                                             unsigned int* __restrict__ keys, unsigned int* __restrict__ vals, int n, MigList ML) {
    __shared__ float4 tiles[G2P_NT / 32][TILE_F4];
    float4* tile = tiles[threadIdx.x >> 5];
    const int p_raw = blockIdx.x * G2P_NT + threadIdx.x;
    if ((p_raw & ~31) >= n) return;                                         // whole warp past the end
    const bool live = p_raw < n;                                            // tail lanes recompute the last particle and write nothing:
    const int p = live ? p_raw : n - 1;                                     // the warp stays converged for the votes below
    // the deformation gradients are needed only after the gather: pull their lines towards the SM now
    prefetch_l1(P.a[PE0] + p); prefetch_l1(P.a[PE1] + p); prefetch_l1(P.a[PE2] + p);
    prefetch_l1(P.a[PQ0] + p); prefetch_l1(P.a[PQ1] + p); prefetch_l1(P.a[PQ2] + p);
    const float dt = clk->dt;
    const float4 X = ldg4(P.a[PX] + p);
    const int cell = __float_as_int(X.w);
    int ci = cell_i(cell), cj = cell_j(cell), ck = cell_k(cell);
    Axis ax, ay, az;
    bool complete = axis_setup(ax, X.x, ci, G.nx, G.ihx);
    complete &= axis_setup(ay, X.y, cj, G.ny, G.ihy);
    complete &= axis_setup(az, X.z, ck, G.nz, G.ihz);
    float rx[4], ry[4], rz[4];                                              // x_i - x_p per axis: h (o - 1 - f)
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
    TileRef T;
    if (stage_tile(G, tile, T, cell, complete)) {
        g2p_gather<2>(G, ax, ay, az, nrx, rx, ry, rz, tile, ax.n0 - T.ox0, S);
        if (S.smin == 0.0f) g2p_stick_correction<2>(G, ax, ay, az, rx, ry, rz, tile, ax.n0 - T.ox0, S);
    } else {                                                                 // rare: row ends, freshly moved particles, domain faces
        g2p_gather<AEP_FALLBACK_MODE>(G, ax, ay, az, nrx, rx, ry, rz, tile, 0, S);
        if (S.smin == 0.0f) g2p_stick_correction<0>(G, ax, ay, az, rx, ry, rz, tile, 0, S);
    }
    float (&va)[3] = S.va; float (&B)[9] = S.B; float (&g)[9] = S.g;
    const float vp[3] = { S.va[0] + S.vc[0], S.va[1] + S.vc[1], S.va[2] + S.vc[2] };       // sum w s v~ = sum w v~ - sum_{s=0} w v~
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
        if (cci != ci || ccj != cj || cck != ck || nan) {
            if (live) atomicAdd(&clk->escaped, 1ull);
            if (!(nfx == nfx)) nfx = 0.5f;
            if (!(nfy == nfy)) nfy = 0.5f;
            if (!(nfz == nfz)) nfz = 0.5f;
        }
        ci = cci; cj = ccj; ck = cck;
    }
    // ---- deformation gradient + plasticity
    const float4 e0 = ldg4(P.a[PE0] + p), e1 = ldg4(P.a[PE1] + p), e2 = ldg4(P.a[PE2] + p);
    const float4 q0 = ldg4(P.a[PQ0] + p), q1 = ldg4(P.a[PQ1] + p), q2 = ldg4(P.a[PQ2] + p);
    float FE[9] = { e0.x, e0.y, e0.z, e1.x, e1.y, e1.z, e2.x, e2.y, e2.z };
    float FP[9] = { q0.x, q0.y, q0.z, q1.x, q1.y, q1.z, q2.x, q2.y, q2.z };
    float GF[9], Fh[9];
    mat_mul(g, FE, GF);
#pragma unroll
    for (int i = 0; i < 9; ++i) Fh[i] = fmaf(dt, GF[i], FE[i]);                  // HybridSolver.cpp:575
    float q = e1.w;
    return_map(mpar, Fh, FE, FP, q);
    const float Jp = mat_det(FP);
    // ---- write back
    const int ncell = cell_pack(ci, cj, ck);
    {   // particles that left their cell are out of order until the next physical sort
        const unsigned mv = __ballot_sync(0xffffffffu, live && ncell != cell);
        if ((threadIdx.x & 31) == 0 && mv) atomicAdd(&clk->moved_since_sort, (unsigned long long)__popc(mv));
    }
    if (!live) return;
    P.a[PX][p] = make_float4(nfx, nfy, nfz, __int_as_float(ncell));
    P.a[PVM][p] = make_float4(vp[0], vp[1], vp[2], q1.w);
    P.a[PC0][p] = make_float4(B[0], B[1], B[2], 0.f);
    P.a[PC1][p] = make_float4(B[3], B[4], B[5], 0.f);
    P.a[PC2][p] = make_float4(B[6], B[7], B[8], 0.f);
    P.a[PE0][p] = make_float4(FE[0], FE[1], FE[2], e0.w);
    P.a[PE1][p] = make_float4(FE[3], FE[4], FE[5], q);
    P.a[PE2][p] = make_float4(FE[6], FE[7], FE[8], Jp);
    P.a[PQ0][p] = make_float4(FP[0], FP[1], FP[2], q0.w);
    P.a[PQ1][p] = make_float4(FP[3], FP[4], FP[5], q1.w);
    P.a[PQ2][p] = make_float4(FP[6], FP[7], FP[8], q2.w);
    keys[p] = sort_key(ci, cj, ck, G);
    vals[p] = (unsigned)p;
    if (ML.axis >= 0 && q1.w != 0.0f) {                                     // live particle of a slab context: did it leave the slab?
        const int ca = ML.axis == 0 ? ci : (ML.axis == 1 ? cj : ck);
        if (ca < ML.lo || ca >= ML.hi) {
            const int side = ca < ML.lo ? 0 : 1;
            const unsigned long long slot = atomicAdd(ML.counts + side, 1ull);
            if (slot < (unsigned long long)ML.cap) (side == 0 ? ML.list[0] : ML.list[1])[slot] = (unsigned)p;
        }
    }
}

inline void g2p_launch(cudaStream_t st, int sm_count, const PartP& P, const GridP& G, const MatParams& mat, SimClock* clk, unsigned int* keys,
                       unsigned int* vals, long long n, const MigList& ML) {
    (void)sm_count;
    k_g2p<<<(unsigned)((n + G2P_NT - 1) / G2P_NT), G2P_NT, 0, st>>>(P, G, mat, clk, keys, vals, (int)n, ML);
}

#endif  // AEP_G2P_PIPE
// ================================================================================================ host <-> device
// fp64 reference layouts -> packed fp32 records.  `st` is a staged chunk: columns of length cnt, in the order
// x(3) v(3) B1(3) B2(3) B3(3) m vol q, then FE (cnt x 9, column-major per particle), FP likewise.
__global__ void k_upload_convert(PartP P, GridP G, const double* __restrict__ st, int cnt, int dst0, long long id0,
                                 double mnx, double mny, double mnz, double hx, double hy, double hz, SimClock* clk) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cnt) return;
    const double* c = st;
    auto col = [&](int k) { return c[(size_t)k * cnt + i]; };
    const double mn[3] = { mnx, mny, mnz }, h[3] = { hx, hy, hz };
    const int nres[3] = { G.nx, G.ny, G.nz };
    float fr[3]; int ce[3]; bool bad = false;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const double u = (col(a) - mn[a]) / h[a];
        int cc = (int)u;                                         // static_cast<int>, HybridSolver.cpp:34-36
        double f = u - (double)cc;
        if (f < 0.0) { cc -= 1; f += 1.0; }                      // negative side of the grid: keep f in [0,1)
        if (cc < 0 || cc >= nres[a] || !(u == u)) { bad = true; cc = cc < 0 ? 0 : nres[a] - 1; f = 0.5; }
        float ff = (float)f; if (ff >= 1.0f) ff = 0.99999994f;
        fr[a] = ff; ce[a] = cc;
    }
    if (bad) atomicAdd(&clk->escaped, 1ull);
    const int d = dst0 + i;
    const float m = (float)col(15);
    P.a[PX][d] = make_float4(fr[0], fr[1], fr[2], __int_as_float(cell_pack(ce[0], ce[1], ce[2])));
    P.a[PVM][d] = make_float4((float)col(3), (float)col(4), (float)col(5), m);
    P.a[PC0][d] = make_float4((float)col(6), (float)col(7), (float)col(8), 0.f);
    P.a[PC1][d] = make_float4((float)col(9), (float)col(10), (float)col(11), 0.f);
    P.a[PC2][d] = make_float4((float)col(12), (float)col(13), (float)col(14), 0.f);
    const double* fe = st + (size_t)18 * cnt + (size_t)9 * i;    // column-major 3x3: (r,c) at 3c + r
    const double* fp = st + (size_t)27 * cnt + (size_t)9 * i;
    float FP[9];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) FP[3 * r + cc] = (float)fp[3 * cc + r];
    P.a[PE0][d] = make_float4((float)fe[0], (float)fe[3], (float)fe[6], (float)col(16));
    P.a[PE1][d] = make_float4((float)fe[1], (float)fe[4], (float)fe[7], (float)col(17));
    P.a[PE2][d] = make_float4((float)fe[2], (float)fe[5], (float)fe[8], mat_det(FP));
    P.a[PQ0][d] = make_float4(FP[0], FP[1], FP[2], __int_as_float((int)(id0 + i)));
    P.a[PQ1][d] = make_float4(FP[3], FP[4], FP[5], m);
    P.a[PQ2][d] = make_float4(FP[6], FP[7], FP[8], 0.f);
}

// packed records [s0, s0+cnt) -> staged fp64 chunk for original ids [id0, id0+cnt_ids): scatter by id.
// out columns: x(3) v(3) B1(3) B2(3) B3(3) vol q, then FE, FP (9 each, column-major per particle)
__global__ void k_download_convert(PartP P, GridP G, double* __restrict__ st, int n, long long id0, int cnt_ids,
                                   double mnx, double mny, double mnz, double hx, double hy, double hz) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const float4 q0 = P.a[PQ0][s];
    const long long id = (long long)__float_as_int(q0.w) - id0;
    if (id < 0 || id >= cnt_ids) return;
    const size_t i = (size_t)id, cnt = (size_t)cnt_ids;
    const float4 X = P.a[PX][s], VM = P.a[PVM][s], c0 = P.a[PC0][s], c1 = P.a[PC1][s], c2 = P.a[PC2][s];
    const float4 e0 = P.a[PE0][s], e1 = P.a[PE1][s], e2 = P.a[PE2][s], q1 = P.a[PQ1][s], q2 = P.a[PQ2][s];
    const int cell = __float_as_int(X.w);
    st[0 * cnt + i] = mnx + ((double)cell_i(cell) + (double)X.x) * hx;
    st[1 * cnt + i] = mny + ((double)cell_j(cell) + (double)X.y) * hy;
    st[2 * cnt + i] = mnz + ((double)cell_k(cell) + (double)X.z) * hz;
    st[3 * cnt + i] = VM.x; st[4 * cnt + i] = VM.y; st[5 * cnt + i] = VM.z;
    st[6 * cnt + i] = c0.x; st[7 * cnt + i] = c0.y; st[8 * cnt + i] = c0.z;
    st[9 * cnt + i] = c1.x; st[10 * cnt + i] = c1.y; st[11 * cnt + i] = c1.z;
    st[12 * cnt + i] = c2.x; st[13 * cnt + i] = c2.y; st[14 * cnt + i] = c2.z;
    st[15 * cnt + i] = e0.w; st[16 * cnt + i] = e1.w;
    double* fe = st + 17 * cnt + 9 * i; double* fp = st + 26 * cnt + 9 * i;
    fe[0] = e0.x; fe[3] = e0.y; fe[6] = e0.z; fe[1] = e1.x; fe[4] = e1.y; fe[7] = e1.z; fe[2] = e2.x; fe[5] = e2.y; fe[8] = e2.z;
    fp[0] = q0.x; fp[3] = q0.y; fp[6] = q0.z; fp[1] = q1.x; fp[4] = q1.y; fp[7] = q1.z; fp[2] = q2.x; fp[5] = q2.y; fp[8] = q2.z;
}

__global__ void k_download_positions_f32(PartP P, GridP G, float* __restrict__ out, int n, int by_slot) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const float4 X = P.a[PX][s];
    const int id = by_slot ? s : __float_as_int(P.a[PQ0][s].w);
    const int cell = __float_as_int(X.w);
    out[3 * (size_t)id + 0] = fmaf((float)cell_i(cell) + X.x, G.hx, G.mnx);
    out[3 * (size_t)id + 1] = fmaf((float)cell_j(cell) + X.y, G.hy, G.mny);
    out[3 * (size_t)id + 2] = fmaf((float)cell_k(cell) + X.z, G.hz, G.mnz);
}

// grid -> fp64 reference layout (Ng x 3 column-major).  mode 0: after P2G (v = p/m); mode 1: after grid update
// (v = s v~, vt = v~).  Forces get the gravity term the reference folds in at HybridSolver.cpp:457.
__global__ void k_download_grid(GridP G, double* __restrict__ m, double* __restrict__ v, double* __restrict__ f,
                                double* __restrict__ vt, long long n0, long long cnt, long long Ng, int mode) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cnt) return;
    const size_t n = (size_t)(n0 + i);
    const float4 mp = G.mp[n], ff = G.f[n];
    if (m) m[i] = mp.x;
    if (f) { f[i] = ff.x; f[cnt + i] = ff.y; f[2 * cnt + i] = (double)ff.z - (double)G.gravity * (double)mp.x; }
    const int b = (((int)(n / ((size_t)G.nx * G.ny)) >> 3) * G.nby + ((int)((n / G.nx) % G.ny) >> 3)) * G.nbx + ((int)(n % G.nx) >> 3);
    const bool act = G.flags[b] != 0;
    if (mode == 0) {
        if (v) {
            const double im = mp.x > 0.0f ? 1.0 / (double)mp.x : 0.0;
            v[i] = mp.y * im; v[cnt + i] = mp.z * im; v[2 * cnt + i] = mp.w * im;
        }
    } else {
        const float4 t = act ? G.vt[n] : make_float4(0.f, 0.f, 0.f, 0.f);
        if (v) { v[i] = t.w * t.x; v[cnt + i] = t.w * t.y; v[2 * cnt + i] = t.w * t.z; }
        if (vt) { vt[i] = t.x; vt[cnt + i] = t.y; vt[2 * cnt + i] = t.z; }
    }
}

// active blocks / nodes with mass (roofline accounting)
__global__ void __launch_bounds__(256) k_count_active(GridP G, unsigned long long* __restrict__ out2) {
    int bx, by, bz;
    const int b = run_block(G, blockIdx.x, bx, by, bz);
    if (!G.flags[b]) return;
    int cnt = 0;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int t = threadIdx.x + 256 * h;
        const int i = bx * 8 + (t & 7), j = by * 8 + ((t >> 3) & 7), k = bz * 8 + (t >> 6);
        if (i < G.nx && j < G.ny && k < G.nz) cnt += G.mp[((size_t)k * G.ny + j) * G.nx + i].x > 0.0f;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(out2 + 1, (unsigned long long)cnt);
    if (threadIdx.x == 0) atomicAdd(out2, 1ull);
}

// bulk statistics (double accumulation): sum m x (3), sum 0.5 m v^2, sum det FP, sum m
__global__ void __launch_bounds__(256) k_stats(PartP P, GridP G, double* __restrict__ out6, int n) {
    double acc[6] = {0, 0, 0, 0, 0, 0};
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x) {
        const float4 X = P.a[PX][s], VM = P.a[PVM][s];
        const int cell = __float_as_int(X.w);
        const double m = VM.w;
        acc[0] += m * ((double)G.mnx + ((double)cell_i(cell) + X.x) * (double)G.hx);
        acc[1] += m * ((double)G.mny + ((double)cell_j(cell) + X.y) * (double)G.hy);
        acc[2] += m * ((double)G.mnz + ((double)cell_k(cell) + X.z) * (double)G.hz);
        acc[3] += 0.5 * m * ((double)VM.x * VM.x + (double)VM.y * VM.y + (double)VM.z * VM.z);
        acc[4] += m > 0.0 ? (double)P.a[PE2][s].w : 0.0;          // dead slots of a slab context carry no mass and are not counted
        acc[5] += m;
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        double v = acc[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0) atomicAdd(out6 + k, v);
    }
}

}  // namespace aep
