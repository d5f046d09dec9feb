// aep_halo.cuh -- kernels of the slab decomposition: halo plane pack / add, particle migration.
#pragma once
#include "aep_kernels.cuh"

namespace aep {

struct PlaneMap {
    int axis, plane0, nplanes;      // planes plane0 .. plane0+nplanes-1 along `axis`
    int nu, nv;                     // remaining axes, u fastest
};
__device__ __forceinline__ bool plane_node(const GridP& G, const PlaneMap& M, long long t, size_t& node) {
    const long long per = (long long)M.nu * M.nv;
    const int p = (int)(t / per); const long long r = t % per;
    const int u = (int)(r % M.nu), v = (int)(r / M.nu);
    const int pl = M.plane0 + p;
    int i, j, k;
    if (M.axis == 0) { i = pl; j = u; k = v; } else if (M.axis == 1) { i = u; j = pl; k = v; } else { i = u; j = v; k = pl; }
    if (i < 0 || i >= G.nx || j < 0 || j >= G.ny || k < 0 || k >= G.nz) return false;
    node = ((size_t)k * G.ny + j) * G.nx + i;
    return true;
}

__global__ void __launch_bounds__(256) k_halo_pack(const float4* __restrict__ arr, float4* __restrict__ buf, GridP G, PlaneMap M) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)M.nplanes * M.nu * M.nv) return;
    size_t node;
    buf[t] = plane_node(G, M, t, node) ? arr[node] : make_float4(0.f, 0.f, 0.f, 0.f);
}

__global__ void __launch_bounds__(256) k_halo_add(float4* __restrict__ arr, const float4* __restrict__ buf, GridP G, PlaneMap M) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)M.nplanes * M.nu * M.nv) return;
    size_t node;
    if (!plane_node(G, M, t, node)) return;
    const float4 r = buf[t];
    if (r.x == 0.f && r.y == 0.f && r.z == 0.f && r.w == 0.f) return;
    float4 a = arr[node];
    a.x += r.x; a.y += r.y; a.z += r.z; a.w += r.w;
    arr[node] = a;
    const int i = (int)(node % G.nx), j = (int)((node / G.nx) % G.ny), k = (int)(node / ((size_t)G.nx * G.ny));
    G.flags[((k >> 3) * G.nby + (j >> 3)) * G.nbx + (i >> 3)] = 1;       // so that the block is updated and cleared
}

__device__ __forceinline__ int cell_axis(int cell, int axis) { return axis == 0 ? cell_i(cell) : (axis == 1 ? cell_j(cell) : cell_k(cell)); }

// keys with dead slots (particles that migrated away: mass 0) and particles outside the slab pushed behind every live key
__global__ void k_build_keys_slab(const float4* __restrict__ X, const float4* __restrict__ Q1, unsigned int* __restrict__ keys,
                                  unsigned int* __restrict__ vals, int n, GridP G, int axis, int lo, int hi, int key_bits) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int c = __float_as_int(X[i].w);
    const int ca = cell_axis(c, axis);
    const unsigned key = sort_key(cell_i(c), cell_j(c), cell_k(c), G);
    // dead slots behind live particles that sit outside the slab (those are extracted by the next migration), both behind the slab's own
    keys[i] = Q1[i].w == 0.0f ? (key | (3u << key_bits)) : ((ca < lo || ca >= hi) ? (key | (1u << key_bits)) : key);
    vals[i] = (unsigned)i;
}

__global__ void __launch_bounds__(256) k_migrate_extract(PartP P, int n, int axis, int lo, int hi, float4* __restrict__ to_low,
                                                         float4* __restrict__ to_high, long long cap, unsigned long long* __restrict__ counts) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const int ca = cell_axis(__float_as_int(P.a[PX][p].w), axis);
    if (ca >= lo && ca < hi) return;
    float4 q1 = P.a[PQ1][p];
    if (q1.w == 0.0f) return;                                             // dead slot: left earlier, waits for the next re-sort
    const int side = ca < lo ? 0 : 1;
    const unsigned long long slot = atomicAdd(counts + side, 1ull);
    if ((long long)slot >= cap) return;                                   // caller sees count > capacity and fails loudly
    float4* dst = (side == 0 ? to_low : to_high) + slot * P_NARR;
#pragma unroll
    for (int a = 0; a < P_NARR; ++a) dst[a] = P.a[a][p];
    // the slot stays in the arrays as a massless, volumeless tracer (scatters add exact zeros) until the next physical sort drops it
    float4 vm = P.a[PVM][p]; vm.w = 0.0f; P.a[PVM][p] = vm;
    float4 e0 = P.a[PE0][p]; e0.w = 0.0f; P.a[PE0][p] = e0;
    q1.w = 0.0f; P.a[PQ1][p] = q1;
}

// listed leavers (MigList filled by k_g2p) -> records in the caller's send buffers; the slots become dead (see k_migrate_extract)
__global__ void __launch_bounds__(256) k_migrate_gather(PartP P, MigList ML, float4* __restrict__ to_low, float4* __restrict__ to_high) {
    const int per_side = (ML.cap + 255) / 256;
    const int side = blockIdx.x / per_side;
    const long long t = (long long)(blockIdx.x - side * per_side) * 256 + threadIdx.x;
    const unsigned long long cnt = ML.counts[side];
    if (t >= (long long)ML.cap || (unsigned long long)t >= cnt) return;
    const unsigned p = (side == 0 ? ML.list[0] : ML.list[1])[t];
    float4* dst = (side == 0 ? to_low : to_high) + (size_t)t * P_NARR;
#pragma unroll
    for (int a = 0; a < P_NARR; ++a) dst[a] = P.a[a][p];
    float4 vm = P.a[PVM][p]; vm.w = 0.0f; P.a[PVM][p] = vm;
    float4 e0 = P.a[PE0][p]; e0.w = 0.0f; P.a[PE0][p] = e0;
    float4 q1 = P.a[PQ1][p]; q1.w = 0.0f; P.a[PQ1][p] = q1;
}

__global__ void __launch_bounds__(256) k_migrate_insert(PartP P, int n_old, const float4* __restrict__ buf, int cnt, SimClock* clk) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r == 0) atomicAdd(&clk->moved_since_sort, (unsigned long long)cnt);   // arrivals sit unsorted at the tail
    if (r >= cnt) return;
#pragma unroll
    for (int a = 0; a < P_NARR; ++a) P.a[a][n_old + r] = buf[(size_t)r * P_NARR + a];
}

// download in slot order with ids (multi-GPU gather happens on the host by id)
__global__ void k_download_local(PartP P, GridP G, double* __restrict__ st, long long* __restrict__ ids, int s0, int cnt,
                                 double mnx, double mny, double mnz, double hx, double hy, double hz) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= cnt) return;
    const int s = s0 + t; const size_t i = (size_t)t, c = (size_t)cnt;
    const float4 X = P.a[PX][s], VM = P.a[PVM][s], c0 = P.a[PC0][s], c1 = P.a[PC1][s], c2 = P.a[PC2][s];
    const float4 e0 = P.a[PE0][s], e1 = P.a[PE1][s], e2 = P.a[PE2][s], q0 = P.a[PQ0][s], q1 = P.a[PQ1][s], q2 = P.a[PQ2][s];
    const int cell = __float_as_int(X.w);
    ids[i] = (long long)__float_as_int(q0.w);
    st[0 * c + i] = mnx + ((double)cell_i(cell) + (double)X.x) * hx;
    st[1 * c + i] = mny + ((double)cell_j(cell) + (double)X.y) * hy;
    st[2 * c + i] = mnz + ((double)cell_k(cell) + (double)X.z) * hz;
    st[3 * c + i] = VM.x; st[4 * c + i] = VM.y; st[5 * c + i] = VM.z;
    st[6 * c + i] = c0.x; st[7 * c + i] = c0.y; st[8 * c + i] = c0.z;
    st[9 * c + i] = c1.x; st[10 * c + i] = c1.y; st[11 * c + i] = c1.z;
    st[12 * c + i] = c2.x; st[13 * c + i] = c2.y; st[14 * c + i] = c2.z;
    st[15 * c + i] = e0.w; st[16 * c + i] = e1.w;
    double* fe = st + 17 * c + 9 * i; double* fp = st + 26 * c + 9 * i;
    fe[0] = e0.x; fe[3] = e0.y; fe[6] = e0.z; fe[1] = e1.x; fe[4] = e1.y; fe[7] = e1.z; fe[2] = e2.x; fe[5] = e2.y; fe[8] = e2.z;
    fp[0] = q0.x; fp[3] = q0.y; fp[6] = q0.z; fp[1] = q1.x; fp[4] = q1.y; fp[7] = q1.z; fp[2] = q2.x; fp[5] = q2.y; fp[8] = q2.z;
}

}  // namespace aep
