// aep_kernels.cuh -- sm_100a kernels of the MPM substep: data layout, clock, grid passes, scatter machinery, stand-alone P2G,
// host <-> device conversion.  The two big particle kernels (forces, fused G2P + P2G) live in aep_particle.cuh.
//
// Device data layout (all fp32, 16-byte records so every particle access is one LDG/STG.128):
//   particle arrays (struct of float4 arrays, index = slot in cell-sorted order)
//     X  = (fx, fy, fz, cell)     fractional position inside the cell, packed cell (i | j<<10 | k<<20)
//     V  = (vx, vy, vz, B00)      velocity and the 9 entries of the APIC matrix B (row-major) in 12 floats:
//     C0 = (B01, B02, B10, B11)   written by G2P, read by P2G only -- in the fused substep they never leave the SM between the two
//     C1 = (B12, B20, B21, B22)
//     E0 = (FE row 0, vol)  E1 = (FE row 1, q)  E2 = (FE row 2, det F_P)   read by the force kernel and G2P
//     Q0..Q2 = rows of F_P (.w unused)     touched only by particles whose return mapping changed a singular value
//     K  = (m, id, -, -)                   constants: never written after the upload
//   grid arrays, node index = k*sz + j*sy + i  (whole grid: (k*ny + j)*nx + i, RegularGrid.cpp:164-168)
//     mp = (m, px, py, pz)   f = (fx, fy, fz, -)   vt = (v~x, v~y, v~z, s)  with v = s * v~  (s = 0 on sticking collider nodes)
//   block flags: one byte per 8x8x8 node block, set by the scatters; grid passes run over the compact list of flagged blocks.
//
// Scatter strategy: shared-memory float atomics are CAS loops on sm_100 (ATOMS.CAST.SPIN) while global memory has native vector
// reductions (REDG.E.ADD.F32x4).  So nothing is accumulated in shared memory: a half-warp walks a run of cell-sorted particles,
// its 16 lanes own the 16 (j,k) rows of the 4x4x4 stencil and keep the 4 nodes of their row in registers while the particles
// stay in one cell (and slide the 4-node window when the next cell is an x-neighbour); a node goes to L2 with one REDG.F32x4 when
// it leaves the window.  Gather strategy: the (cells+3) x 4 x 4 node box of a half-warp's 16 particles is fetched by TMA into the
// half-warp's own shared memory and every lane reads its 64 nodes from there.
#pragma once
#include <stdint.h>
#include "aep_math.cuh"
#include "aep_pack.cuh"
#include "aep_scatter.cuh"
#include "aep_gather.cuh"

namespace aep {

enum { PX = 0, PV, PC0, PC1, PE0, PE1, PE2, PQ0, PQ1, PQ2, PK, P_NARR };

struct PartP {
    float4* a[P_NARR];
};

// rows of B from the packed (V.w, C0, C1)
__device__ __forceinline__ void unpack_B(const float4& V, const float4& C0, const float4& C1, float4& b0, float4& b1, float4& b2) {
    b0 = make_float4(V.w, C0.x, C0.y, 0.f); b1 = make_float4(C0.z, C0.w, C1.x, 0.f); b2 = make_float4(C1.y, C1.z, C1.w, 0.f);
}

// simulation clock + reductions, lives in device memory so that n substeps need no host round trip
struct SimClock {
    double t, inner_t, frame_dt, cfl, rate_floor, hmin;
    float dt;                          // dt used by forces / grid update (previous iteration's value, HybridSolver.cpp:873-878)
    unsigned int vmax_bits;            // max |v_i| of the last grid update, as float bits (non-negative -> ordered as uint)
    int frame_flag, frame_no;
    long long substeps;
    unsigned long long escaped;        // particles that tried to leave the grid (clamped), sticky
    float vmax_last;
    int pad0;
    // adaptive re-sort (aep_config.sort_every == 0): particles that changed cell since the last physical sort, and the accumulated
    // per-substep out-of-order fraction.  moved_since_sort + sort_cost are reset together (16 bytes) when the particles are re-sorted.
    unsigned long long moved_since_sort;
    float sort_cost;
    int n_slots;                       // particle slots in use (live + dead).  Device-resident: migration changes it without the host
    int n_dead;                        // slots whose particle migrated away since the last compaction
    int comm_timeout;                  // a peer's flag did not arrive within the spin limit (sticky; the host fails loudly on it)
    // aep_run_frames: the device stops itself.  halt 0 = running; 1 = the clock kernel has just completed the last requested frame,
    // the rest of this substep (G2P ...) still runs; 2 = halted, every kernel returns at once.  Kernels ahead of the clock kernel in
    // a substep return when halt >= 1, kernels behind it when halt >= 2 (the clock kernel turns 1 into 2 the next time it runs).
    int halt;
    int stop_frame;                    // halt when frame_no reaches this (-1: never)
    long long stop_substep;            // ... or when `substeps` reaches this (-1: never)
    int pad2;
    unsigned long long mig_dropped;    // leavers / arrivals that did not fit the migration buffers (sticky; the host fails loudly on it)
    float vmax_mass_floor;             // 0 = the reference rule.  > 0 (opt-in, NOT the reference): nodes lighter than this do not enter max|v|
    int pad1;
    unsigned long long mig_sent, mig_received;   // particles this context handed to / took from its neighbours (peer exchange), cumulative
};
#define AEP_HALT_PRE(clk)  do { if ((clk)->halt >= 1) return; } while (0)
#define AEP_HALT_POST(clk) do { if ((clk)->halt >= 2) return; } while (0)

// slab decomposition: G2P appends the slots of particles whose new cell left [lo, hi) along `axis` to per-side index lists, so
// that the migration step touches only the leavers instead of scanning every particle (axis < 0: not a slab / lists not bound)
struct MigList {
    int axis, lo, hi, cap;
    unsigned int* list[2];
    unsigned long long* counts;        // [2] (low, high), device memory
};

__device__ __forceinline__ int cell_i(int c) { return c & 1023; }
__device__ __forceinline__ int cell_j(int c) { return (c >> 10) & 1023; }
__device__ __forceinline__ int cell_k(int c) { return (c >> 20) & 1023; }
__device__ __forceinline__ int cell_pack(int i, int j, int k) { return i | (j << 10) | (k << 20); }

// CTA index of a grid pass -> block coordinates inside the run range, and the flat block index (flags)
__device__ __forceinline__ int run_block(const GridP& G, int r, int& bx, int& by, int& bz) {
    bx = G.rb0[0] + r % G.rbn[0]; by = G.rb0[1] + (r / G.rbn[0]) % G.rbn[1]; bz = G.rb0[2] + r / (G.rbn[0] * G.rbn[1]);
    return (bz * G.nby + by) * G.nbx + bx;
}
// does this context hold node (i,j,k)?  (whole-grid contexts: the grid; slab contexts: the slab's reach)
__device__ __forceinline__ bool node_held(const GridP& G, int i, int j, int k) {
    return i >= G.a0[0] && i < G.a1[0] && j >= G.a0[1] && j < G.a1[1] && k >= G.a0[2] && k < G.a1[2];
}

// ================================================================================================ sort keys
// Particles are ordered by cell (x fastest) -- or, with sort_bricks, by 4x4x4-cell brick and then by cell inside the brick.  The
// scatters only need particles of one cell to be adjacent, which any cell-granular order provides; the gathers want the 16
// particles of a half-warp in one row of cells.  Keys are built from the positions when a re-sort is due (not by G2P: that cost
// 8 B per particle and substep for something needed every ~30 substeps).
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
// one CTA per block of the whole grid cost each pass ~0.2 ms of CTAs that only read a zero flag; with the list a pass is a
// persistent grid over the occupied blocks and runs at memory speed.  The caller zeroes the count first.
__global__ void __launch_bounds__(256) k_list_blocks(GridP G, int nrun, unsigned int* __restrict__ list, unsigned int* __restrict__ count, const SimClock* __restrict__ clk) {
    if (clk) AEP_HALT_PRE(clk);
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

__device__ __forceinline__ bool node_valid(const GridP& G, int i, int j, int k) {
    return i >= G.v0[0] && i < G.v1[0] && j >= G.v0[1] && j < G.v1[1] && k >= G.v0[2] && k < G.v1[2];
}
// node (i,j,k) of thread t (two per thread) of block (bx,by,bz); false outside the held range.  valid: inside the range whose sums
// are complete (everything for a whole-grid context)
__device__ __forceinline__ bool block_node(const GridP& G, int bx, int by, int bz, int t, size_t& n, bool& valid) {
    const int i = bx * 8 + (t & 7), j = by * 8 + ((t >> 3) & 7), k = bz * 8 + (t >> 6);
    if (!node_held(G, i, j, k)) return false;
    n = nidx(G, i, j, k);
    valid = node_valid(G, i, j, k);
    return true;
}
__device__ __forceinline__ bool block_node(const GridP& G, int bx, int by, int bz, int t, size_t& n) {
    bool valid; return block_node(G, bx, by, bz, t, n, valid);
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
            size_t n;
            if (block_node(G, bx, by, bz, threadIdx.x + 256 * h, n)) { G.mp[n] = z; G.f[n] = z; }
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
    const float floor_m = clk->vmax_mass_floor;
    float vm = 0.0f;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        size_t n; bool valid;
        if (block_node(G, bx, by, bz, threadIdx.x + 256 * h, n, valid) && valid) {
            const float4 mp = G.mp[n];
            if (mp.x > floor_m) {
                const float im = 1.0f / mp.x;
                const float vx = mp.y * im, vy = mp.z * im, vz = mp.w * im;
                vm = fmaxf(vm, sqrtf(vx * vx + vy * vy + vz * vz));
            }
        }
    }
    block_max_to_clock(vm, clk);
}

// Collider description for k_grid_update.  Static colliders (the reference: HS:484) are sampled once on the host into a byte code
// per node (+ normals); a moving collider (opt-in, aep_set_collider_motion) is an analytic level set evaluated at the node, translated
// by its velocity times the simulated time, and the projection works on the velocity relative to the collider.
struct ColliderP {
    int moving;                 // 0: sampled codes (G.ls_code / G.ls_nrm); 1: analytic, evaluated per node
    int kind;                   // AEP_LS_* of the analytic form
    float par[8];
    float vel[3];               // collider velocity
    float off[3];               // translation at the start of this substep (vel * t), refreshed by the clock kernel
    int coulomb;                // 0: the reference's stick-or-slide (HS:494-502, the reduction at :501 is a no-op); 1: opt-in Coulomb friction
};
// analytic level sets on the device (same formulas as LevelSet.cpp:8-42 and the host sampler in aep_engine.cu): true inside (phi <= 0)
__device__ __forceinline__ bool collider_eval(const ColliderP& C, float x, float y, float z, float& nx_, float& ny_, float& nz_) {
    x -= C.off[0]; y -= C.off[1]; z -= C.off[2];
    nx_ = 0.f; ny_ = 0.f; nz_ = 1.f;
    switch (C.kind) {
    case 1: return z - C.par[0] <= 0.0f;                                                   // ground
    case 2: {                                                                               // wall corner + ground
        const float dz = z - C.par[2], dx = C.par[0] - x, dy = C.par[1] - y;
        if (fminf(fminf(dz, dx), dy) > 0.0f) return false;
        const float az = fabsf(dz), ax = fabsf(dx), ay = fabsf(dy);
        if (az <= ax && az <= ay) { } else if (ay <= ax) { ny_ = -1.f; nz_ = 0.f; } else { nx_ = -1.f; nz_ = 0.f; }
        return true;
    }
    case 3: case 6: {                                                                       // sphere (+ ground for kind 3)
        const float dx = x - C.par[0], dy = y - C.par[1], dz = z - C.par[2];
        const float r = sqrtf(dx * dx + dy * dy + dz * dz);
        const float ps = r - C.par[3], pg = C.kind == 3 ? z - C.par[4] : 1e30f;
        if (fminf(ps, pg) > 0.0f) return false;
        if (ps <= pg && r > 0.0f) { const float ir = 1.0f / r; nx_ = dx * ir; ny_ = dy * ir; nz_ = dz * ir; }
        return true;
    }
    case 4: {                                                                               // box (inside is free)
        const float d[6] = { z - C.par[2], C.par[5] - z, x - C.par[0], C.par[3] - x, y - C.par[1], C.par[4] - y };
        int best = 0;
#pragma unroll
        for (int f = 1; f < 6; ++f) if (d[f] < d[best]) best = f;
        if (d[best] > 0.0f) return false;
        nx_ = best == 2 ? 1.f : (best == 3 ? -1.f : 0.f); ny_ = best == 4 ? 1.f : (best == 5 ? -1.f : 0.f); nz_ = best == 0 ? 1.f : (best == 1 ? -1.f : 0.f);
        return true;
    }
    default: return false;
    }
}

// updateGridVelocities_ (HybridSolver.cpp:725-737) + gravity (:457) + max|v| (RegularGrid.cpp:188-200)
// + gridCollisionHandling_ level-set part (:467-511), one coalesced float4 pass over the active blocks.
// CLEAR (the fused substep): (m,p) and f are dead once v~ is written -- zero them and drop the block flag here, so that the
// fused G2P+P2G kernel scatters into a clean grid without a separate clearing pass.
template <bool CLEAR>
__global__ void __launch_bounds__(256) k_grid_update(GridP G, const ColliderP* __restrict__ Cp, SimClock* clk, const unsigned int* __restrict__ list,
                                                     const unsigned int* __restrict__ count) {
    if (CLEAR) AEP_HALT_PRE(clk);
    __shared__ ColliderP C;
    if (threadIdx.x == 0) C = *Cp;
    __syncthreads();
    const unsigned nb = *count;
    const float dt = clk->dt;
    const float floor_m = clk->vmax_mass_floor;
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    float vm = 0.0f;
    for (unsigned q = blockIdx.x; q < nb; q += gridDim.x) {
        int bx, by, bz;
        const int b = run_block(G, (int)list[q], bx, by, bz);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int t = threadIdx.x + 256 * h;
            size_t n; bool valid;
            if (!block_node(G, bx, by, bz, t, n, valid)) continue;
            const float4 mp = G.mp[n];
            float vx = 0.f, vy = 0.f, vz = 0.f, s = 1.0f;
            if (valid && mp.x > 0.0f) {
                const float4 f = G.f[n];
                const float im = 1.0f / mp.x;
                vx = fmaf(dt, f.x * im, mp.y * im);
                vy = fmaf(dt, f.y * im, mp.z * im);
                vz = fmaf(dt, fmaf(-G.gravity, mp.x, f.z) * im, mp.w * im);
                if (mp.x > floor_m) vm = fmaxf(vm, sqrtf(vx * vx + vy * vy + vz * vz));
                bool inside = false; float nx_ = 0.f, ny_ = 0.f, nz_ = 1.f;
                if (C.moving) {
                    const int i = bx * 8 + (t & 7), j = by * 8 + ((t >> 3) & 7), k = bz * 8 + (t >> 6);
                    inside = collider_eval(C, fmaf((float)i, G.hx, G.mnx), fmaf((float)j, G.hy, G.mny), fmaf((float)k, G.hz, G.mnz), nx_, ny_, nz_);
                } else {
                    const int code = G.ls_code ? G.ls_code[n] : 0;
                    inside = code != 0;
                    if (code == 7) { const float4 nn = G.ls_nrm[n]; nx_ = nn.x; ny_ = nn.y; nz_ = nn.z; }
                    else if (code) {
                        nx_ = (code == 1) ? 1.f : (code == 2 ? -1.f : 0.f);
                        ny_ = (code == 3) ? 1.f : (code == 4 ? -1.f : 0.f);
                        nz_ = (code == 5) ? 1.f : (code == 6 ? -1.f : 0.f);
                    }
                }
                if (inside) {
                    const float cvx = C.moving ? C.vel[0] : 0.f, cvy = C.moving ? C.vel[1] : 0.f, cvz = C.moving ? C.vel[2] : 0.f;   // HS:484: static
                    float rx_ = vx - cvx, ry_ = vy - cvy, rz_ = vz - cvz;
                    const float vn = rx_ * nx_ + ry_ * ny_ + rz_ * nz_;                  // HybridSolver.cpp:486
                    if (vn < 0.0f) {                                                     // :488 approaching
                        rx_ = fmaf(-vn, nx_, rx_); ry_ = fmaf(-vn, ny_, ry_); rz_ = fmaf(-vn, nz_, rz_);   // :490-492
                        const float vtn = sqrtf(rx_ * rx_ + ry_ * ry_ + rz_ * rz_);
                        // :494-502 stick test; the Coulomb reduction at :501 is a no-op expression statement.  Opt-in Coulomb: |v_t|
                        // shrinks by mu |v_n| (what :500-501 set out to do), stored as v = c + s (v~ - c) with 0 <= s <= 1 (c: collider velocity)
                        if (!C.coulomb) s = (vtn < -G.friction * vn) ? 0.0f : 1.0f;
                        else s = (vtn <= -G.friction * vn) ? 0.0f : 1.0f + G.friction * vn / vtn;
                        vx = rx_ + cvx; vy = ry_ + cvy; vz = rz_ + cvz;
                    }
                }
            }
            G.vt[n] = make_float4(vx, vy, vz, s);
            if (CLEAR) { G.mp[n] = z4; G.f[n] = z4; }
        }
        if (CLEAR && threadIdx.x == 0) G.flags[b] = 0;
    }
    block_max_to_clock(vm, clk);
}

// dt rule + frame clipping of HybridSolver.cpp:878-892, one thread, all in double like the reference
__global__ void k_advance_clock(SimClock* clk, int fixed_dt, ColliderP* col) {
    if (clk->halt >= 1) { clk->halt = 2; return; }
    const float vmax = __uint_as_float(clk->vmax_bits);
    clk->vmax_last = vmax; clk->vmax_bits = 0u;
    if (fixed_dt == 2) return;                                   // stage-level API: only latch max|v|
    clk->sort_cost += (float)clk->moved_since_sort / (float)max(clk->n_slots, 1);
    if (fixed_dt == 1) {                                         // pinned dt (aep_set_fixed_dt): plain time accumulation
        clk->inner_t += (double)clk->dt; clk->frame_flag = 0;
        if (clk->inner_t >= clk->frame_dt) { clk->inner_t -= clk->frame_dt; clk->t += clk->frame_dt; clk->frame_flag = 1; clk->frame_no += 1; }
    } else {
        double dt = clk->cfl / fmax(clk->rate_floor, (double)vmax / clk->hmin);
        if (clk->inner_t + dt >= clk->frame_dt) {
            dt = clk->frame_dt - clk->inner_t; clk->t += clk->frame_dt; clk->inner_t = 0.0; clk->frame_flag = 1; clk->frame_no += 1;
        } else {
            clk->inner_t += dt; clk->frame_flag = 0;
        }
        clk->dt = (float)dt;
    }
    clk->substeps += 1;
    if ((clk->stop_frame >= 0 && clk->frame_no >= clk->stop_frame) || (clk->stop_substep >= 0 && clk->substeps >= clk->stop_substep)) clk->halt = 1;   // G2P of this substep still runs
    if (col && col->moving) {                                    // where the collider is when the next grid update runs
        const double tt = clk->t + clk->inner_t;
        for (int a = 0; a < 3; ++a) col->off[a] = (float)(col->vel[a] * tt);
    }
}
__global__ void k_initial_dt(SimClock* clk) {                      // HybridSolver.cpp:860
    const float vmax = __uint_as_float(clk->vmax_bits);
    clk->vmax_last = vmax; clk->vmax_bits = 0u;
    clk->dt = (float)(clk->cfl / fmax(clk->rate_floor, (double)vmax / clk->hmin));
}

// ================================================================================================ scatter machinery
// Both scatters (P2G, force) run in two phases per round of 16 cell-sorted particles per half-warp:
//   phase A  thread-per-particle: everything that depends on the particle only (1-D weights, affine / stress matrices)
//            goes to a per-half-warp shared-memory record;
//   phase B  one HALF-WARP per particle, 16 lanes = the 16 (j,k) rows of the 4x4x4 stencil, each lane owns the 4 nodes
//            along x of its row and accumulates in registers (packed fp32x2, aep_pack.cuh) over the run of particles that share
//            a cell.  No atomics and no shuffles in the inner loop.
// Sliding window along x: when the next particle's cell is d = 1..3 cells further along x in the same (j,k) row of cells, only the
// d nodes that fall out of the 4-node window are reduced to memory and the accumulators shift; x-adjacent cells share 3 of their
// 4 nodes per row, so a sorted row of C cells costs C+3 reductions per lane instead of 4C.  A half-warp keeps its window open over
// ROUNDS x 16 consecutive particles: every flush costs one LSU wavefront per lane and node (the 16 rows of a half-warp lie in 16
// different cache lines), and the scatters are bound by exactly those wavefronts plus the shared-memory record reads.
// Run STARTS (not ends) drive the window: phase A ballots "my cell differs from the cell of the particle before me" (the half-warp's
// last cell of the round before for its first lane), so nothing has to be known about particles that are not processed yet -- the
// fused G2P+P2G kernel only learns a particle's new cell when it has advected it.

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
    if (node_held(G, ni, nj, nk0)) {
        atomicAdd(dst + nidx(G, ni, nj, nk0), a0);
        if (mark) G.flags[((nk0 >> 3) * G.nby + (nj >> 3)) * G.nbx + (ni >> 3)] = 1;
    }
    if (node_held(G, ni, nj, nk1)) {
        atomicAdd(dst + nidx(G, ni, nj, nk1), a1);
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
#ifndef AEP_REQUAD_MOV
#define AEP_REQUAD_MOV 1                // 1: register moves; 0: the shared-memory bounce below (round 1).  The scatters are bound by the LSU pipe (98 % in k_p2g in the
                                        // flowing state), the moves ride on the ALU: 17.77 against 18.12 ms per substep (C5, flowing), 13.96 against 14.12 at rest
#endif
__device__ __forceinline__ float4 requad(float4* slot, f32x2 lo, f32x2 hi) {
    float4 v;
#if AEP_REQUAD_MOV
    asm volatile("mov.b64 {%0, %1}, %4;\n\tmov.b64 {%2, %3}, %5;" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(lo), "l"(hi));
    return v;
#endif
    const unsigned a = (unsigned)__cvta_generic_to_shared(slot);
    asm volatile("st.shared.b64 [%4], %5;\n\tst.shared.b64 [%4+8], %6;\n\tld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a), "l"(lo), "l"(hi) : "memory");
    return v;
}
// rows that this context does not hold never receive anything (outside the grid: the weights are masked, HS:44-46; outside a slab's
// reach: a particle can only get there after it has left the slab, and its new owner scatters it)
__device__ __forceinline__ bool row_held(const GridP& G, int nj, int nk) {
    return nj >= G.a0[1] && nj < G.a1[1] && nk >= G.a0[2] && nk < G.a1[2];
}
__device__ __forceinline__ void flush_row_pk(const GridP& G, float4* __restrict__ dst, float4* slot, int cell, int j, int k, const AccRow& a, bool mark) {
    const int ni0 = cell_i(cell) - 1, nj = cell_j(cell) - 1 + j, nk = cell_k(cell) - 1 + k;
    if (!row_held(G, nj, nk)) return;
    float4* row = dst + nidx(G, 0, nj, nk);
    unsigned char* frow = G.flags + ((nk >> 3) * G.nby + (nj >> 3)) * G.nbx;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int ni = ni0 + i;
        if (ni >= G.a0[0] && ni < G.a1[0]) {
            atomicAdd(row + ni, requad(slot, a.lo[i], a.hi[i]));
            // block flags: the first node of the window in the grid, and the last one when it sits in another 8-block
            if (mark && (i == 0 || ni == G.a0[0] || ((i == 3 || ni == G.a1[0] - 1) && (ni >> 3) != (max(ni0, G.a0[0]) >> 3)))) frow[ni >> 3] = 1;
        }
    }
}
// sliding window along x: reduce the d nodes that leave the window, shift the rest
__device__ __forceinline__ bool slide_row_pk(const GridP& G, float4* __restrict__ dst, float4* slot, int cur, int next, int j, int k, AccRow& a, bool mark) {
    const int d = next - cur;
    if (d <= 0 || d >= 4 || (next & 1023) - (cur & 1023) != d) return false;
    const int nj = cell_j(cur) - 1 + j, nk = cell_k(cur) - 1 + k;
    const bool in_jk = row_held(G, nj, nk);
    float4* row = dst + (in_jk ? nidx(G, 0, nj, nk) : 0);
    unsigned char* frow = G.flags + (in_jk ? ((nk >> 3) * G.nby + (nj >> 3)) * G.nbx : 0);
    int ni = cell_i(cur) - 1;
    for (int s = 0; s < d; ++s, ++ni) {
        if (in_jk && ni >= G.a0[0] && ni < G.a1[0]) {
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
// a run of same-cell particles starts at this record: move the window from cell `prev` (-1: nothing accumulated yet) to cell `cur`
__device__ __forceinline__ void window_move(const GridP& G, float4* __restrict__ dst, float4* slot, int prev, int cur, int j, int k, AccRow& a, bool mark) {
    if (prev < 0) return;
    if (!slide_row_pk(G, dst, slot, prev, cur, j, k, a, mark)) {
        flush_row_pk(G, dst, slot, prev, j, k, a, mark);
        acc_zero_ordered(a);
    }
}
// phase A of a round: which records of the half-warp start a run (bit s of the result, already shifted to the caller's half-warp),
// the cell before each lane's (`prev`), and the half-warp's last cell (`wcell`, carried to the next round; -1 before the first)
__device__ __forceinline__ unsigned run_starts(int cell, int& wcell, int& prev) {
    const int lane = threadIdx.x & 31;
    prev = __shfl_up_sync(0xffffffffu, cell, 1);
    if ((lane & 15) == 0) prev = wcell;
    const unsigned starts = __ballot_sync(0xffffffffu, cell != prev);
    wcell = __shfl_sync(0xffffffffu, cell, (lane & 16) | 15);
    return starts >> (lane & 16);
}

// ================================================================================================ P2G (stand-alone)
// particleToGrid_ (HybridSolver.cpp:113-231): m_i = sum w m ; p_i = sum w m (v + (3/h^2) B (x_i - x_p)).
// Per particle the momentum of node offset (i,j,k) is q0 + Qm (i,j,k)^T with Qm = m (3/h^2) B diag(h), q0 = m v - Qm (1+f).
// Phase-A record of one particle: see p2g_make_record (aep_scatter.cuh).
// Used for the first transfer (HS:854), restarts, mesh points and the stage-level API; inside a substep the same phase B runs at
// the end of the fused kernel (aep_particle.cuh) on records made from registers.
#define P2G_HW_PAD 2
#define P2G_HW_F4 (16 * P2G_STRIDE + P2G_HW_PAD)
// A particle that has left its cell since the last physical sort sits alone between two stretches of its old cell's run:
// ... A A A [B] A A A ...  Moving the window to B and back costs two flushes (up to 8 reductions per lane) and the window code twice;
// instead such a SINGLETON goes to memory by itself (4 reductions per lane from a scratch row) and the window stays where it is.
// Recognised where a run starts at record `it`, another one at `it + 1`, and that one returns to the cell before `it`.
__device__ __forceinline__ bool singleton_at(const float4* __restrict__ cell_rec, int stride, unsigned starts, int it, int prev) {
    if (it >= 15 || prev < 0 || !((starts >> (it + 1)) & 1u)) return false;
    return __float_as_int(cell_rec[stride].x) == prev;                       // the next record's cell
}
// one round of phase B: the half-warp's 16 records into the window
__device__ __forceinline__ void p2g_phase_b(const GridP& G, const float4* __restrict__ recs, unsigned starts, float4* slot, int j, int k, int yoff, int zoff,
                                            f32x2 J, f32x2 K, AccRow& acc) {
#pragma unroll 1
    for (int it = 0; it < 16; ++it) {
        const float4* r = recs + it * P2G_STRIDE;
        if ((starts >> it) & 1u) {                                            // a run of particles sharing a cell starts here
            const float2 cn = *reinterpret_cast<const float2*>(r + 7);
            if (singleton_at(r + 7, P2G_STRIDE, starts, it, __float_as_int(cn.y))) {
                AccRow one; acc_zero(one);
                p2g_row_accumulate(r, yoff, zoff, J, K, one);
                flush_row_pk(G, G.mp, slot, __float_as_int(cn.x), j, k, one, true);
                starts &= ~(2u << it);                                        // the window never left the run that continues at it + 1
                continue;
            }
            window_move(G, G.mp, slot, __float_as_int(cn.y), __float_as_int(cn.x), j, k, acc, true);
        }
        p2g_row_accumulate(r, yoff, zoff, J, K, acc);
    }
}
// ROUNDS x 16 consecutive particles per half-warp: 8 for large scenes (fewest reductions), less when that would leave SMs idle
// n < 0: the particle count lives on the device (clk->n_slots; the launch covers -n slots: slab contexts with peer communication)
template <int ROUNDS>
__global__ void __launch_bounds__(256) k_p2g(PartP P, GridP G, int n, const SimClock* __restrict__ clk) {
    if (clk) AEP_HALT_POST(clk);                                              // inside a substep; nullptr: stage-level call
    if (n < 0) n = clk->n_slots;
    __shared__ float4 stage[8][2][P2G_HW_F4];
    __shared__ float4 bounce[256];                                            // lane-private slots of requad()
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    float4* slot = bounce + threadIdx.x;
    const int cta_particles = 256 * ROUNDS;
    const int chunk = strided_chunk(blockIdx.x, (n + cta_particles - 1) / cta_particles, G.strips);
    if (chunk < 0) return;
    const int hw = lane >> 4, s = lane & 15, j = s & 3, k = s >> 2;
    const int hbase = chunk * cta_particles + wib * (32 * ROUNDS) + hw * (16 * ROUNDS);      // first particle of this half-warp
    if (chunk * cta_particles + wib * (32 * ROUNDS) >= n) return;             // warp-uniform; no block-level barrier below
    float fj = (float)j, fk = (float)k;
    int yoff = 16 + 4 * j, zoff = 32 + 4 * k;                                 // byte offsets of Ny[j], Nz[k] inside a record
    // lane constants: ptxas re-derives them from %tid on every trip (8 instructions) unless they come out of something it cannot
    // rematerialise -- an identity shuffle
    fj = __shfl_sync(0xffffffffu, fj, lane); fk = __shfl_sync(0xffffffffu, fk, lane);
    yoff = __shfl_sync(0xffffffffu, yoff, lane); zoff = __shfl_sync(0xffffffffu, zoff, lane);
    const f32x2 J = pk1(fj), K = pk1(fk);
    AccRow acc; acc_zero(acc);
    const float4* recs = &stage[wib][hw][0];
    int wcell = -1;
#pragma unroll 1
    for (int round = 0; round < ROUNDS; ++round) {
        unsigned starts;
        {   // ---- phase A: thread per particle
            const int q = hbase + round * 16 + s;                            // may lie past the end: such lanes repeat the last particle with zero mass
            const int p = min(q, n - 1);
            const float4 X = ldg4(P.a[PX] + p), V = ldg4(P.a[PV] + p), C0 = ldg4(P.a[PC0] + p), C1 = ldg4(P.a[PC1] + p);
            const float m = (q < n) ? __ldg(&P.a[PK][p].x) : 0.0f;
            float4 b0, b1, b2; unpack_B(V, C0, C1, b0, b1, b2);
            int prev;
            starts = run_starts(__float_as_int(X.w), wcell, prev);
            __syncwarp();                                                    // phase B of the round before is done with the records
            p2g_make_record(&stage[wib][hw][s * P2G_STRIDE], X, make_float4(V.x, V.y, V.z, m), b0, b1, b2, m, G.apic, G.hx, G.hy, G.hz, __int_as_float(prev));
        }
        __syncwarp();
        // ---- phase B: half-warp per particle, lane = (j,k) row of the stencil, 4 nodes along x in packed accumulators
        p2g_phase_b(G, recs, starts, slot, j, k, yoff, zoff, J, K, acc);
    }
    if (wcell >= 0) flush_row_pk(G, G.mp, slot, wcell, j, k, acc, true);
}

// launch helper shared by the engine and the mesh transfers
// device_count: the launch covers n slots, the kernel takes the count in use from the clock
inline void p2g_launch(cudaStream_t st, const PartP& P, const GridP& G, long long n, const SimClock* clk = nullptr, bool device_count = false) {
    if (n <= 0) return;
    const int na = device_count ? -(int)n : (int)n;
    if (n >= (1ll << 22)) { const int chunks = (int)((n + 2047) / 2048); k_p2g<8><<<strided_grid(chunks, G.strips), 256, 0, st>>>(P, G, na, clk); }
    else if (n >= (1ll << 19)) { const int chunks = (int)((n + 511) / 512); k_p2g<2><<<strided_grid(chunks, G.strips), 256, 0, st>>>(P, G, na, clk); }
    else { const int chunks = (int)((n + 255) / 256); k_p2g<1><<<strided_grid(chunks, G.strips), 256, 0, st>>>(P, G, na, clk); }
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
        const int nk = clampi(az.n0 + k, G.a0[2], G.a1[2] - 1);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int nj = clampi(ay.n0 + j, G.a0[1], G.a1[1] - 1);
            const float wjk = ay.N[j] * az.N[k];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int ni = clampi(ax.n0 + i, G.a0[0], G.a1[0] - 1);
                dens = fmaf(ax.N[i] * wjk, ldg4(G.mp + nidx(G, ni, nj, nk)).x, dens);
            }
        }
    }
    dens *= G.inv_cell_vol;
    const float m = P.a[PK][p].x;
    float4 e0 = P.a[PE0][p]; e0.w = m * (1.0f / dens); P.a[PE0][p] = e0;
}

// v_i = p_i / m_i where m_i > 0 (HybridSolver.cpp:233-240) for every active node, into the vt array (free between P2G and
// the grid update), so that the force gather below does 64 loads and no divisions per particle.
__global__ void __launch_bounds__(256) k_grid_normalise(GridP G, const SimClock* __restrict__ clk, const unsigned int* __restrict__ list, const unsigned int* __restrict__ count) {
    AEP_HALT_PRE(clk);
    const unsigned nb = *count;
    for (unsigned q = blockIdx.x; q < nb; q += gridDim.x) {
        int bx, by, bz;
        run_block(G, (int)list[q], bx, by, bz);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            size_t n;
            if (!block_node(G, bx, by, bz, threadIdx.x + 256 * h, n)) continue;
            const float4 mp = G.mp[n];
            const float im = mp.x > 0.0f ? 1.0f / mp.x : 0.0f;
            G.vt[n] = make_float4(mp.y * im, mp.z * im, mp.w * im, 1.0f);
        }
    }
}

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
    P.a[PV][d] = make_float4((float)col(3), (float)col(4), (float)col(5), (float)col(6));
    P.a[PC0][d] = make_float4((float)col(7), (float)col(8), (float)col(9), (float)col(10));
    P.a[PC1][d] = make_float4((float)col(11), (float)col(12), (float)col(13), (float)col(14));
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
    P.a[PQ0][d] = make_float4(FP[0], FP[1], FP[2], 0.f);
    P.a[PQ1][d] = make_float4(FP[3], FP[4], FP[5], 0.f);
    P.a[PQ2][d] = make_float4(FP[6], FP[7], FP[8], 0.f);
    P.a[PK][d] = make_float4(m, __int_as_float((int)(id0 + i)), 0.f, 0.f);
}

// one particle's state as the fp64 staging columns (shared by the id-ordered and the slot-ordered download)
// out columns: x(3) v(3) B1(3) B2(3) B3(3) vol q, then FE, FP (9 each, column-major per particle)
__device__ __forceinline__ void download_one(const PartP& P, int s, double* __restrict__ st, size_t i, size_t cnt,
                                             double mnx, double mny, double mnz, double hx, double hy, double hz) {
    const float4 X = P.a[PX][s], V = P.a[PV][s], C0 = P.a[PC0][s], C1 = P.a[PC1][s];
    const float4 e0 = P.a[PE0][s], e1 = P.a[PE1][s], e2 = P.a[PE2][s], q0 = P.a[PQ0][s], q1 = P.a[PQ1][s], q2 = P.a[PQ2][s];
    const int cell = __float_as_int(X.w);
    st[0 * cnt + i] = mnx + ((double)cell_i(cell) + (double)X.x) * hx;
    st[1 * cnt + i] = mny + ((double)cell_j(cell) + (double)X.y) * hy;
    st[2 * cnt + i] = mnz + ((double)cell_k(cell) + (double)X.z) * hz;
    st[3 * cnt + i] = V.x; st[4 * cnt + i] = V.y; st[5 * cnt + i] = V.z;
    st[6 * cnt + i] = V.w; st[7 * cnt + i] = C0.x; st[8 * cnt + i] = C0.y;
    st[9 * cnt + i] = C0.z; st[10 * cnt + i] = C0.w; st[11 * cnt + i] = C1.x;
    st[12 * cnt + i] = C1.y; st[13 * cnt + i] = C1.z; st[14 * cnt + i] = C1.w;
    st[15 * cnt + i] = e0.w; st[16 * cnt + i] = e1.w;
    double* fe = st + 17 * cnt + 9 * i; double* fp = st + 26 * cnt + 9 * i;
    fe[0] = e0.x; fe[3] = e0.y; fe[6] = e0.z; fe[1] = e1.x; fe[4] = e1.y; fe[7] = e1.z; fe[2] = e2.x; fe[5] = e2.y; fe[8] = e2.z;
    fp[0] = q0.x; fp[3] = q0.y; fp[6] = q0.z; fp[1] = q1.x; fp[4] = q1.y; fp[7] = q1.z; fp[2] = q2.x; fp[5] = q2.y; fp[8] = q2.z;
}
// packed records [0, n) -> staged fp64 chunk for original ids [id0, id0+cnt_ids): scatter by id.
__global__ void k_download_convert(PartP P, double* __restrict__ st, int n, long long id0, int cnt_ids,
                                   double mnx, double mny, double mnz, double hx, double hy, double hz) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const float4 kk = P.a[PK][s];
    const long long id = (long long)__float_as_int(kk.y) - id0;
    if (id < 0 || id >= cnt_ids) return;
    download_one(P, s, st, (size_t)id, (size_t)cnt_ids, mnx, mny, mnz, hx, hy, hz);
}

// positions only.  by_slot: out[3*slot]; else out[3*(id - id_base)] with a bounds check (ids of a context are id_base + 0..n_ids)
__global__ void k_download_positions_f32(PartP P, GridP G, float* __restrict__ out, int n, int by_slot, long long id_base, long long n_ids) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const float4 X = P.a[PX][s];
    long long id = s;
    if (!by_slot) { id = (long long)__float_as_int(P.a[PK][s].y) - id_base; if (id < 0 || id >= n_ids) return; }
    const int cell = __float_as_int(X.w);
    out[3 * (size_t)id + 0] = fmaf((float)cell_i(cell) + X.x, G.hx, G.mnx);
    out[3 * (size_t)id + 1] = fmaf((float)cell_j(cell) + X.y, G.hy, G.mny);
    out[3 * (size_t)id + 2] = fmaf((float)cell_k(cell) + X.z, G.hz, G.mnz);
}

// grid -> fp64 reference layout (Ng x 3 column-major, node index (k*ny + j)*nx + i).  mode 0: after P2G (v = p/m); mode 1: after
// grid update (v = s v~, vt = v~).  Forces get the gravity term the reference folds in at HybridSolver.cpp:457.  Nodes that a slab
// context does not hold come back as zeros.
__global__ void k_download_grid(GridP G, double* __restrict__ m, double* __restrict__ v, double* __restrict__ f,
                                double* __restrict__ vt, long long n0, long long cnt, int mode) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= cnt) return;
    const long long g = n0 + t;
    const int i = (int)(g % G.nx), j = (int)((g / G.nx) % G.ny), k = (int)(g / ((long long)G.nx * G.ny));
    float4 mp = make_float4(0.f, 0.f, 0.f, 0.f), ff = mp, tv = mp;
    const bool held = node_held(G, i, j, k);
    const size_t n = held ? nidx(G, i, j, k) : 0;
    if (held) { mp = G.mp[n]; ff = G.f[n]; }
    if (m) m[t] = mp.x;
    if (f) { f[t] = ff.x; f[cnt + t] = ff.y; f[2 * cnt + t] = (double)ff.z - (double)G.gravity * (double)mp.x; }
    const bool act = held && G.flags[((k >> 3) * G.nby + (j >> 3)) * G.nbx + (i >> 3)] != 0;
    if (mode == 0) {
        if (v) {
            const double im = mp.x > 0.0f ? 1.0 / (double)mp.x : 0.0;
            v[t] = mp.y * im; v[cnt + t] = mp.z * im; v[2 * cnt + t] = mp.w * im;
        }
    } else {
        if (act) tv = G.vt[n];
        if (v) { v[t] = G.cvx + tv.w * (tv.x - G.cvx); v[cnt + t] = G.cvy + tv.w * (tv.y - G.cvy); v[2 * cnt + t] = G.cvz + tv.w * (tv.z - G.cvz); }
        if (vt) { vt[t] = tv.x; vt[cnt + t] = tv.y; vt[2 * cnt + t] = tv.z; }
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
        size_t n; bool valid;
        if (block_node(G, bx, by, bz, threadIdx.x + 256 * h, n, valid) && valid) cnt += G.mp[n].x > 0.0f;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(out2 + 1, (unsigned long long)cnt);
    if (threadIdx.x == 0) atomicAdd(out2, 1ull);
}

// bulk statistics (double accumulation): sum m x (3), sum 0.5 m v^2, sum det FP, sum m, live particles
__global__ void __launch_bounds__(256) k_stats(PartP P, GridP G, double* __restrict__ out7, const SimClock* __restrict__ clk) {
    const int n = clk->n_slots;
    double acc[7] = {0, 0, 0, 0, 0, 0, 0};
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x) {
        const float4 X = P.a[PX][s], V = P.a[PV][s];
        const int cell = __float_as_int(X.w);
        const double m = P.a[PK][s].x;
        acc[0] += m * ((double)G.mnx + ((double)cell_i(cell) + X.x) * (double)G.hx);
        acc[1] += m * ((double)G.mny + ((double)cell_j(cell) + X.y) * (double)G.hy);
        acc[2] += m * ((double)G.mnz + ((double)cell_k(cell) + X.z) * (double)G.hz);
        acc[3] += 0.5 * m * ((double)V.x * V.x + (double)V.y * V.y + (double)V.z * V.z);
        acc[4] += m > 0.0 ? (double)P.a[PE2][s].w : 0.0;          // dead slots of a slab context carry no mass and are not counted
        acc[5] += m;
        acc[6] += m > 0.0 ? 1.0 : 0.0;
    }
#pragma unroll
    for (int k = 0; k < 7; ++k) {
        double v = acc[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0) atomicAdd(out7 + k, v);
    }
}

}  // namespace aep
