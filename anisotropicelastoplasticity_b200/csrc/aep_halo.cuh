// aep_halo.cuh -- kernels of the slab decomposition.
//
// Two generations live here:
//   * the caller-driven exchange (k_halo_pack / k_halo_add / k_migrate_*): buffers owned by the caller, who moves the bytes
//     (torch.distributed, a peer copy ...).  Kept for the stage-level / Python-driven path and its tests.
//   * the peer-memory exchange (k_peer_*): every rank maps its neighbours' communication block (CUDA IPC between processes, plain
//     pointers inside one process) and the kernels of a substep store halo planes, migrating particles and max|v| straight into the
//     neighbour's memory over NVLink, then publish an epoch flag; the receiver's stream waits on the flag with a one-thread kernel.
//     No host synchronisation, no NCCL launch, no count round trip: the substep stays a fixed sequence of launches (CUDA-graph ready).
#pragma once
#include "aep_kernels.cuh"

namespace aep {

// ================================================================================================ caller-driven exchange
struct PlaneMap {
    int axis, plane0, nplanes;      // planes plane0 .. plane0+nplanes-1 along `axis`
    int nu, nv;                     // remaining axes, u fastest
};
__device__ __forceinline__ bool plane_node(const GridP& G, const PlaneMap& M, long long t, size_t& node, int& bi) {
    const long long per = (long long)M.nu * M.nv;
    const int p = (int)(t / per); const long long r = t % per;
    const int u = (int)(r % M.nu), v = (int)(r / M.nu);
    const int pl = M.plane0 + p;
    int i, j, k;
    if (M.axis == 0) { i = pl; j = u; k = v; } else if (M.axis == 1) { i = u; j = pl; k = v; } else { i = u; j = v; k = pl; }
    if (!node_held(G, i, j, k)) return false;
    node = nidx(G, i, j, k);
    bi = ((k >> 3) * G.nby + (j >> 3)) * G.nbx + (i >> 3);
    return true;
}

__global__ void __launch_bounds__(256) k_halo_pack(const float4* __restrict__ arr, float4* __restrict__ buf, GridP G, PlaneMap M) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)M.nplanes * M.nu * M.nv) return;
    size_t node; int bi;
    buf[t] = plane_node(G, M, t, node, bi) ? arr[node] : make_float4(0.f, 0.f, 0.f, 0.f);
}

__global__ void __launch_bounds__(256) k_halo_add(float4* __restrict__ arr, const float4* __restrict__ buf, GridP G, PlaneMap M) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)M.nplanes * M.nu * M.nv) return;
    size_t node; int bi;
    if (!plane_node(G, M, t, node, bi)) return;
    const float4 r = buf[t];
    if (r.x == 0.f && r.y == 0.f && r.z == 0.f && r.w == 0.f) return;
    float4 a = arr[node];
    a.x += r.x; a.y += r.y; a.z += r.z; a.w += r.w;
    arr[node] = a;
    G.flags[bi] = 1;                                                      // so that the block is updated and cleared
}

__device__ __forceinline__ int cell_axis(int cell, int axis) { return axis == 0 ? cell_i(cell) : (axis == 1 ? cell_j(cell) : cell_k(cell)); }

// keys with dead slots (particles that migrated away: mass 0) and particles outside the slab pushed behind every live key
__global__ void k_build_keys_slab(const float4* __restrict__ X, const float4* __restrict__ K, unsigned int* __restrict__ keys,
                                  unsigned int* __restrict__ vals, int n, GridP G, int axis, int lo, int hi, int key_bits) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int c = __float_as_int(X[i].w);
    const int ca = cell_axis(c, axis);
    const unsigned key = sort_key(cell_i(c), cell_j(c), cell_k(c), G);
    // dead slots behind live particles that sit outside the slab (those are extracted by the next migration), both behind the slab's own
    keys[i] = K[i].x == 0.0f ? (key | (3u << key_bits)) : ((ca < lo || ca >= hi) ? (key | (1u << key_bits)) : key);
    vals[i] = (unsigned)i;
}

// a particle leaves: its slot stays in the arrays as a massless, volumeless tracer (scatters add exact zeros) until the next
// physical sort drops it.  id -1 keeps it out of the id-ordered downloads.
__device__ __forceinline__ void slot_kill(const PartP& P, unsigned p) {
    float4 e0 = P.a[PE0][p]; e0.w = 0.0f; P.a[PE0][p] = e0;
    P.a[PK][p] = make_float4(0.0f, __int_as_float(-1), 0.f, 0.f);
}

__global__ void __launch_bounds__(256) k_migrate_extract(PartP P, int n, int axis, int lo, int hi, float4* __restrict__ to_low,
                                                         float4* __restrict__ to_high, long long cap, unsigned long long* __restrict__ counts) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const int ca = cell_axis(__float_as_int(P.a[PX][p].w), axis);
    if (ca >= lo && ca < hi) return;
    if (P.a[PK][p].x == 0.0f) return;                                     // dead slot: left earlier, waits for the next re-sort
    const int side = ca < lo ? 0 : 1;
    const unsigned long long slot = atomicAdd(counts + side, 1ull);
    if ((long long)slot >= cap) return;                                   // caller sees count > capacity and fails loudly
    float4* dst = (side == 0 ? to_low : to_high) + slot * P_NARR;
#pragma unroll
    for (int a = 0; a < P_NARR; ++a) dst[a] = P.a[a][p];
    slot_kill(P, (unsigned)p);
}

// listed leavers (MigList filled by G2P) -> records in the caller's send buffers; the slots become dead
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
    slot_kill(P, p);
}

__global__ void __launch_bounds__(256) k_migrate_insert(PartP P, int n_old, const float4* __restrict__ buf, int cnt, SimClock* clk) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r == 0) { atomicAdd(&clk->moved_since_sort, (unsigned long long)cnt); clk->n_slots = n_old + cnt; }   // arrivals sit unsorted at the tail
    if (r >= cnt) return;
#pragma unroll
    for (int a = 0; a < P_NARR; ++a) P.a[a][n_old + r] = buf[(size_t)r * P_NARR + a];
}

// download in slot order with ids (multi-GPU gather happens on the host by id); dead slots come back with id -1
__global__ void k_download_local(PartP P, double* __restrict__ st, long long* __restrict__ ids, int s0, int cnt,
                                 double mnx, double mny, double mnz, double hx, double hy, double hz) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= cnt) return;
    ids[t] = (long long)__float_as_int(P.a[PK][s0 + t].y);
    download_one(P, s0 + t, st, (size_t)t, (size_t)cnt, mnx, mny, mnz, hx, hy, hz);
}

// ================================================================================================ peer-memory exchange
#define AEP_MAX_WORLD 16
// Head of every rank's communication block.  Everything a PEER writes sits here or in the buffers behind it.
struct CommHead {
    unsigned long long halo_flag[2][2];         // [what][side]: epoch of the last complete halo buffer the neighbour on `side` stored here
    unsigned long long mig_flag[2];             // [side]
    unsigned long long mig_count[2];            // [side]: records in mig_in[side] (written before the flag)
    unsigned long long vmax_slot[2][AEP_MAX_WORLD];   // [epoch parity][rank]: (epoch << 32) | float bits of that rank's max|v|
    unsigned long long mesh_flag[2][AEP_MAX_WORLD];   // [vertices, elements][rank]: epoch of that rank's last push of its mesh points
};
// Per-rank bookkeeping, device memory, written by this rank's kernels only (epochs advance on the device: the launch sequence of a
// substep carries no changing arguments, so it can be replayed from a CUDA graph)
struct CommLocal {
    unsigned long long halo_epoch[2][2];        // [what][side] exchanges done
    unsigned long long mig_epoch;
    unsigned long long vmax_epoch;
    unsigned long long mesh_epoch[2];
    unsigned int done[8];                       // "last CTA" counters of the send kernels
};
struct CommPeers {
    int rank, world;
    CommHead* head[AEP_MAX_WORLD];              // every rank's block (own included)
    float4* halo_in[2][2];                      // neighbour's receive buffer that THIS rank fills: [what][my side]  (nullptr: no neighbour)
    float4* mig_in[2];                          // neighbour's migration receive buffer that this rank fills: [my side]
    unsigned char* mesh[AEP_MAX_WORLD];         // every rank's mesh block (nullptr without a mesh)
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v; asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v;
}
// one thread waits until *flag >= want; a flag that never arrives is reported (comm_timeout), not waited for forever
#define AEP_SPIN_LIMIT (1u << 24)                                     // x (256 ns sleep + one system-scope load): several seconds
__device__ __forceinline__ void spin_until(const unsigned long long* flag, unsigned long long want, SimClock* clk) {
    if (clk->comm_timeout) return;                                          // sticky: after one timeout nothing waits any more
    for (unsigned spin = 0; ld_acquire_sys(flag) < want; ++spin) {
        __nanosleep(256);
        if (spin > AEP_SPIN_LIMIT) { clk->comm_timeout = 1; return; }
    }
}

// my planes [plane0, plane0 + nplanes) of `arr` (contiguous: the slab axis is the slowest one of a slab context's layout) -> the
// neighbour's receive buffer, then the epoch flag.  Only non-zero nodes travel; the receiver zeroes what it consumed.
__global__ void __launch_bounds__(256) k_peer_halo_send(const float4* __restrict__ src, long long n_f4, float4* __restrict__ peer_buf,
                                                        unsigned long long* peer_flag, unsigned long long* my_epoch, unsigned int* done,
                                                        const SimClock* __restrict__ clk, int halt_class) {
    if (clk->halt >= halt_class) return;
    for (long long t = (long long)blockIdx.x * 256 + threadIdx.x; t < n_f4; t += (long long)gridDim.x * 256) {
        const float4 v = src[t];
        if (v.x != 0.f || v.y != 0.f || v.z != 0.f || v.w != 0.f) peer_buf[t] = v;
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0 && atomicAdd(done, 1u) == gridDim.x - 1) {
        *done = 0u;
        const unsigned long long e = *my_epoch + 1ull; *my_epoch = e;
        __threadfence_system();
        st_release_sys(peer_flag, e);
    }
}
// wait for the neighbours' flags of this exchange (one thread; see spin_until)
__global__ void k_peer_wait2(const unsigned long long* flag0, const unsigned long long* epoch0, const unsigned long long* flag1, const unsigned long long* epoch1,
                             SimClock* clk, int halt_class) {
    if (clk->halt >= halt_class) return;
    if (flag0) spin_until(flag0, *epoch0, clk);
    if (flag1) spin_until(flag1, *epoch1, clk);
}
// arr[plane0 ...] += what the neighbour stored; the buffer is zeroed for the next exchange; touched blocks are flagged
__global__ void __launch_bounds__(256) k_peer_halo_add(float4* __restrict__ arr_planes, float4* __restrict__ buf, long long n_f4, GridP G, int axis, int plane0,
                                                       long long plane_nodes, const SimClock* __restrict__ clk, int halt_class) {
    if (clk->halt >= halt_class) return;
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    for (long long t = (long long)blockIdx.x * 256 + threadIdx.x; t < n_f4; t += (long long)gridDim.x * 256) {
        const float4 r = buf[t];
        if (r.x == 0.f && r.y == 0.f && r.z == 0.f && r.w == 0.f) continue;
        buf[t] = z;
        float4 a = arr_planes[t];
        a.x += r.x; a.y += r.y; a.z += r.z; a.w += r.w;
        arr_planes[t] = a;
        // node coordinates: plane index along `axis`, the other two axes in memory order (x fastest)
        const int pl = plane0 + (int)(t / plane_nodes); const long long q = t % plane_nodes;
        int i, j, k;
        if (axis == 1) { i = (int)(q % G.nx); k = (int)(q / G.nx); j = pl; }
        else if (axis == 2) { i = (int)(q % G.nx); j = (int)(q / G.nx); k = pl; }
        else { j = (int)(q % G.ny); k = (int)(q / G.ny); i = pl; }
        G.flags[((k >> 3) * G.nby + (j >> 3)) * G.nbx + (i >> 3)] = 1;
    }
}

// max|v| of this rank -> every rank's slot (own included).  One warp; lane r writes to rank r.
__global__ void k_peer_vmax_share(CommPeers Pr, CommLocal* L, const SimClock* __restrict__ clk, int halt_class) {
    if (clk->halt >= halt_class) return;
    const unsigned long long e = L->vmax_epoch + 1ull;
    const unsigned long long v = (e << 32) | (unsigned long long)clk->vmax_bits;
    if ((int)threadIdx.x < Pr.world) st_release_sys(&Pr.head[threadIdx.x]->vmax_slot[e & 1ull][Pr.rank], v);
    __syncwarp();
    if (threadIdx.x == 0) L->vmax_epoch = e;
}
// global max|v| into the clock (one thread), ahead of the dt rule
__global__ void k_peer_vmax_reduce(CommPeers Pr, const CommLocal* L, SimClock* clk, int halt_class) {
    if (clk->halt >= halt_class) return;
    const unsigned long long e = L->vmax_epoch;
    const unsigned long long* slots = Pr.head[Pr.rank]->vmax_slot[e & 1ull];
    unsigned int m = clk->vmax_bits;
    for (int r = 0; r < Pr.world; ++r) {
        unsigned long long v = 0;
        for (unsigned spin = 0; ((v = ld_acquire_sys(slots + r)) >> 32) != e && !clk->comm_timeout; ++spin) {
            __nanosleep(256);
            if (spin > AEP_SPIN_LIMIT) { clk->comm_timeout = 1; break; }
        }
        m = max(m, (unsigned int)(v & 0xffffffffull));
    }
    clk->vmax_bits = m;
}

// listed leavers (MigList filled by G2P) -> records in the neighbours' receive buffers + their counts, then the flags.
// grid = 2 * ceil(cap / 256) CTAs: the first half serves the low side.
__global__ void __launch_bounds__(256) k_peer_migrate_send(PartP P, MigList ML, CommPeers Pr, CommLocal* L, SimClock* clk, int halt_class) {
    if (clk->halt >= halt_class) return;
    const int per_side = (ML.cap + 255) / 256;
    const int side = blockIdx.x / per_side;
    const long long t = (long long)(blockIdx.x - side * per_side) * 256 + threadIdx.x;
    unsigned long long cnt = ML.counts[side];
    if (cnt > (unsigned long long)ML.cap) cnt = (unsigned long long)ML.cap;     // the overflow was counted by G2P's caller below
    float4* dst_base = Pr.mig_in[side];
    if (t < (long long)cnt) {
        const unsigned p = (side == 0 ? ML.list[0] : ML.list[1])[t];
        if (dst_base) {
            float4* dst = dst_base + (size_t)t * P_NARR;
#pragma unroll
            for (int a = 0; a < P_NARR; ++a) dst[a] = P.a[a][p];
        }
        slot_kill(P, p);
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0 && atomicAdd(&L->done[4], 1u) == gridDim.x - 1) {
        L->done[4] = 0u;
        const unsigned long long e = L->mig_epoch + 1ull; L->mig_epoch = e;
        unsigned long long tot = 0;
        for (int sd = 0; sd < 2; ++sd) {
            const unsigned long long raw = ML.counts[sd];
            const unsigned long long c = raw > (unsigned long long)ML.cap ? (unsigned long long)ML.cap : raw;
            if (raw > c) clk->mig_dropped += raw - c;
            tot += c;
            const int nb = sd == 0 ? Pr.rank - 1 : Pr.rank + 1;
            if (nb >= 0 && nb < Pr.world) {
                CommHead* h = Pr.head[nb];
                h->mig_count[1 - sd] = c;                                    // my low side is the neighbour's high side
                __threadfence_system();
                st_release_sys(&h->mig_flag[1 - sd], e);
            } else if (c) clk->mig_dropped += c;                             // left through a domain face: cannot happen (positions are clamped)
            ML.counts[sd] = 0ull;
        }
        clk->n_dead += (int)tot; clk->mig_sent += tot;
    }
}
// append what the neighbours stored (counts from my block's head)
__global__ void __launch_bounds__(256) k_peer_migrate_insert(PartP P, const float4* __restrict__ in_low, const float4* __restrict__ in_high, CommHead* head,
                                                             int cap_slots, SimClock* clk, unsigned int* done, int halt_class) {
    if (clk->halt >= halt_class) return;
    const int n_old = clk->n_slots;
    const int c0 = (int)head->mig_count[0], c1 = (int)head->mig_count[1];
    const int room = max(cap_slots - n_old, 0);
    const int take = min(c0 + c1, room);
    for (int r = blockIdx.x * 256 + threadIdx.x; r < take; r += gridDim.x * 256) {
        const float4* src = r < c0 ? in_low + (size_t)r * P_NARR : in_high + (size_t)(r - c0) * P_NARR;
#pragma unroll
        for (int a = 0; a < P_NARR; ++a) P.a[a][n_old + r] = src[a];
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0 && atomicAdd(done, 1u) == gridDim.x - 1) {
        *done = 0u;
        clk->n_slots = n_old + take; clk->mig_received += (unsigned long long)take;
        clk->moved_since_sort += (unsigned long long)take;                  // arrivals sit unsorted at the tail
        if (take < c0 + c1) clk->mig_dropped += (unsigned long long)(c0 + c1 - take);
        head->mig_count[0] = 0ull; head->mig_count[1] = 0ull;
    }
}

// cloth: the points this rank advanced -> every other rank's copy of the mesh.  `narr` float4 arrays of `n` points each lie back to
// back at byte offset `off` of the mesh block (the same on every rank).  grid.y = peer rank.
__global__ void __launch_bounds__(256) k_peer_mesh_push(CommPeers Pr, CommLocal* L, const unsigned char* __restrict__ owner, size_t off, int narr, int n,
                                                        int which, const SimClock* __restrict__ clk, int halt_class) {
    if (clk->halt >= halt_class) return;
    const int peer = blockIdx.y;
    if (peer != Pr.rank && Pr.mesh[peer]) {
        const float4* src = reinterpret_cast<const float4*>(Pr.mesh[Pr.rank] + off);
        float4* dst = reinterpret_cast<float4*>(Pr.mesh[peer] + off);
        for (int t = blockIdx.x * 256 + threadIdx.x; t < n; t += gridDim.x * 256)
            if (owner[t])
                for (int a = 0; a < narr; ++a) dst[(size_t)a * n + t] = src[(size_t)a * n + t];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0 && atomicAdd(&L->done[5 + which], 1u) == gridDim.x * gridDim.y - 1) {
        L->done[5 + which] = 0u;
        const unsigned long long e = L->mesh_epoch[which] + 1ull; L->mesh_epoch[which] = e;
        __threadfence_system();
        for (int r = 0; r < Pr.world; ++r) if (r != Pr.rank) st_release_sys(&Pr.head[r]->mesh_flag[which][Pr.rank], e);
    }
}
__global__ void k_peer_mesh_wait(CommPeers Pr, const CommLocal* L, int which, SimClock* clk, int halt_class) {
    if (clk->halt >= halt_class) return;
    const unsigned long long e = L->mesh_epoch[which];
    for (int r = 0; r < Pr.world; ++r) if (r != Pr.rank) spin_until(&Pr.head[Pr.rank]->mesh_flag[which][r], e, clk);
}

}  // namespace aep
