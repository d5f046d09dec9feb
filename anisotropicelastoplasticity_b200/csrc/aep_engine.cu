// aep_engine.cu -- context management, stepping sequence and the C ABI (include/aep_b200.h) of libaep_b200.so.
//
// One aep_ctx = one GPU = one slab of the domain.  All work is enqueued on ctx->stream; a substep is
//   forces -> grid update/collide (+max|v|, clears what it consumed) -> clock (dt rule) -> G2P/advect/F/plasticity -> P2G of the next substep
// (HybridSolver.cpp:867-1032) with no host synchronisation, dt and the particle count living in device memory.  The sequence
// carries no changing launch arguments, so it is captured once into a CUDA graph and replayed: one host launch per substep.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cub/device/device_radix_sort.cuh>
#include <nvtx3/nvToolsExt.h>          // header-only; ranges are emitted when AEP_NVTX=1 (a profiler that injects NVTX picks them up)
#include <unistd.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/aep_b200.h"
#include "aep_kernels.cuh"
#include "aep_particle.cuh"
#include "aep_halo.cuh"
#include "aep_mesh.cuh"

using namespace aep;

namespace {

thread_local std::string g_create_error;

struct Timers {
    cudaEvent_t ev[2];
    double ms[AEP_NUM_STAGES];
    long long calls[AEP_NUM_STAGES];
};

// what the host learns about the device-side state a few substeps late, without blocking (SimClock::sort_cost .. comm_timeout)
struct Telemetry {
    float sort_cost; int n_slots, n_dead, comm_timeout;
};

// peer-memory communication of a slab context (aep_halo.cuh)
struct Comm {
    bool exported = false, connected = false;
    int rank = 0, world = 1;
    unsigned char* block = nullptr; size_t bytes = 0;       // this rank's block: CommHead | halo_in[2][2] | mig_in[2]
    size_t halo_off[2][2]{}, mig_off[2]{};
    long long plane_nodes = 0; int mig_cap = 0;
    CommLocal* d_local = nullptr;
    CommPeers peers{};
    std::vector<void*> opened;                               // cudaIpcOpenMemHandle mappings to close
    // several slab contexts inside ONE process (aep_comm_connect_local): a device-side spin on a flag that another stream of the same
    // GPU must raise can starve (streams may share a hardware queue), so inside a process the phases of a substep are ordered by
    // CUDA events between the members' streams; the flag protocol runs all the same and finds its flags already raised
    std::vector<aep_ctx*> group;
    cudaEvent_t ev_phase[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
};

}  // namespace

struct aep_ctx {
    aep_config cfg;
    int device = 0, sm_count = 148;
    cudaStream_t stream = nullptr, copy_stream = nullptr;
    std::string err;
    long long launches = 0;

    // grid
    GridP G{};
    size_t Ng = 0, Nheld = 0; int nblocks = 0;  // nodes of the whole grid / of the planes this context holds; 8^3 blocks (flags array)
    int nrun = 0;                                // blocks the grid passes run over (G.rb0 / G.rbn)
    unsigned int *d_blist = nullptr, *d_bcount = nullptr;   // compact list of the flagged blocks among them (k_list_blocks)
    int pass_ctas = 1;                           // persistent grid of the grid passes
    double h[3]{}, hmin = 0;
    std::vector<void*> dev_allocs;
    unsigned char* d_ls_code = nullptr; float4* d_ls_nrm = nullptr;      // real allocations (Nheld entries); G.ls_* are biased
    int grid_mode = 0;                          // 0: (m,p) fresh from P2G; 1: vt valid
    CUtensorMap tm_vt{};                        // TMA descriptor of G.vt (held planes) for the gather tiles
    ColliderP col{}; ColliderP* d_col = nullptr;

    // particles
    long long n = 0, cap = 0;                   // slots in use as the host knows them (exact, except with peer communication: see n_launch)
    PartP P[2]{}; int cur = 0;
    unsigned int *d_keys[2] = {nullptr, nullptr}, *d_vals[2] = {nullptr, nullptr};
    DeferP defer{nullptr, nullptr};             // particles that do not fit their half-warp's gather box (aep_particle.cuh)
    ForceA forceA{{nullptr, nullptr, nullptr}}; // A = -V P F_E^T between the two force kernels (split_forces)
    bool split_forces = true;                   // development: AEP_SPLIT_FORCES=0 -> gather and scatter of the forces in one kernel
    void* d_sort_tmp = nullptr; size_t sort_tmp_bytes = 0;
    int key_bits = 0;
    MatParams mat{};
    int steps_since_sort = 0;
    long long pending_leave = 0;                // particles extracted for migration (caller-driven path), dropped at the next re-bin
    MigList mig{};                              // leaver lists filled by G2P; axis < 0 when unbound
    float4* mig_buf[2] = {nullptr, nullptr};    // caller-owned send buffers (caller-driven path)
    long long id_base = 0, n_ids = 0;

    // mesh
    MeshState mesh;

    // clock
    SimClock* d_clk = nullptr;
    double* d_stats = nullptr;
    int fixed_dt = 0;

    // staging
    double* d_stage = nullptr; size_t stage_bytes = 0;
    float* d_frame = nullptr; size_t frame_bytes = 0; cudaEvent_t frame_ev = nullptr; bool frame_pending = false;

    // lagged, non-blocking readback of the telemetry
    static constexpr int RING = 4, LAG = 2;
    Telemetry* h_ring = nullptr; cudaEvent_t ring_ev[RING] = {nullptr, nullptr, nullptr, nullptr};
    SimClock* h_clk = nullptr; cudaEvent_t clk_ev[2] = {nullptr, nullptr};     // aep_run_frames polls the device clock through these
    long long step_counter = 0, last_sort_step = -1;
    long long sorts = 0;

    // CUDA graph of one substep
    cudaGraphExec_t graph = nullptr; bool graph_dirty = true; long long graph_launches = 0;
    bool use_graph = true;
    bool fused = false;                         // G2P and the next P2G as two kernels (default) or in one (k_g2p2g<SCATTER = true>; development:
                                                // AEP_FUSED=1).  Measured on the same box, C5: 14.8 against 16.4 ms per substep at rest, 19.1 against 21.9
                                                // in the flowing state (profiles/README.md, round 2): the scatter alone runs at 3x the occupancy

    Comm comm;

    bool profile = false; Timers tm{};
    bool nvtx = false;                          // AEP_NVTX=1: NVTX ranges per stage (host side of the launches) and per substep
    bool inited = false;
};

namespace {

int fail(aep_ctx* c, int code, const char* fmt, ...) {
    char buf[512]; va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    if (c) c->err = buf; else g_create_error = buf;
    return code;
}
#define CU(call)                                                                                       \
    do {                                                                                               \
        cudaError_t e_ = (call);                                                                       \
        if (e_ != cudaSuccess) return fail(c, e_ == cudaErrorMemoryAllocation ? AEP_ERR_ALLOC : AEP_ERR_CUDA, \
                                           "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

template <typename T>
cudaError_t dalloc(aep_ctx* c, T** p, size_t count) {
    cudaError_t e = cudaMalloc((void**)p, std::max<size_t>(count, 1) * sizeof(T));
    if (e == cudaSuccess) c->dev_allocs.push_back(*p);
    return e;
}

inline int cdiv(long long a, int b) { return (int)((a + b - 1) / b); }

const char* const STAGE_NAMES[AEP_NUM_STAGES] = { "aep:sort", "aep:p2g", "aep:forces(gather)", "aep:grid", "aep:g2p", "aep:mesh", "aep:halo", "aep:g2p2g(fused)",
                                                  "aep:force_scatter", "aep:forces(list)", "aep:g2p(list)" };
struct StageTimer {
    aep_ctx* c; int stage;
    StageTimer(aep_ctx* c_, int s) : c(c_), stage(s) {
        if (c->nvtx) nvtxRangePushA(STAGE_NAMES[s]);
        if (c->profile) cudaEventRecord(c->tm.ev[0], c->stream);
    }
    ~StageTimer() {
        if (c->nvtx) nvtxRangePop();
        if (!c->profile) return;
        cudaEventRecord(c->tm.ev[1], c->stream); cudaEventSynchronize(c->tm.ev[1]);
        float ms = 0; cudaEventElapsedTime(&ms, c->tm.ev[0], c->tm.ev[1]);
        c->tm.ms[stage] += ms; c->tm.calls[stage] += 1;
    }
};

int ensure_stage(aep_ctx* c, size_t bytes) {
    if (c->stage_bytes >= bytes) return AEP_OK;
    if (c->d_stage) { cudaFree(c->d_stage); c->d_stage = nullptr; c->stage_bytes = 0; }
    CU(cudaMalloc((void**)&c->d_stage, bytes));
    c->stage_bytes = bytes;
    return AEP_OK;
}

bool peer_mode(const aep_ctx* c) { return c->comm.connected; }
// particle slots the launches cover: exact, or -- with peer communication, where arrivals change the count on the device -- the capacity
// (CTAs past the device-side count return at once; the launch geometry then never changes and the substep graph stays valid)
long long n_launch(const aep_ctx* c) { return peer_mode(c) ? c->cap : c->n; }

// host index of a held node in the context's own layout (mirror of nidx() minus the bias of the pointers)
size_t held_index(const aep_ctx* c, int i, int j, int k) {
    const GridP& G = c->G;
    return (size_t)((long long)(k - G.a0[2]) * G.sz + (long long)(j - G.a0[1]) * G.sy + (i - G.a0[0]));
}

// ---------------------------------------------------------------------------------- level-set sampling (host)
// HS:473-482 evaluates phi / grad phi at grid nodes only, and colliders are static (HS:484), so the level set is
// sampled once per setLevelSet in fp64 on the host (same formulas as LevelSet.cpp:8-42) and uploaded as a byte code.
double ls_phi(int kind, const double* P, const double x[3]) {
    switch (kind) {
    case AEP_LS_GROUND: return x[2] - P[0];
    case AEP_LS_WALL2GROUND: return std::min(std::min(x[2] - P[2], P[0] - x[0]), P[1] - x[1]);
    case AEP_LS_SPHERE_GROUND: case AEP_LS_SPHERE: {
        double dx = x[0] - P[0], dy = x[1] - P[1], dz = x[2] - P[2];
        const double ps = std::sqrt(dx * dx + dy * dy + dz * dz) - P[3];
        return kind == AEP_LS_SPHERE ? ps : std::min(ps, x[2] - P[4]);
    }
    case AEP_LS_BOX: {
        double d = x[0] - P[0];
        d = std::min(d, P[3] - x[0]); d = std::min(d, x[1] - P[1]); d = std::min(d, P[4] - x[1]);
        d = std::min(d, x[2] - P[2]); d = std::min(d, P[5] - x[2]);
        return d;
    }
    default: return 1.0;
    }
}
// returns code 1..6 for axis normals (+x -x +y -y +z -z) or 7 with n filled
int ls_normal_code(int kind, const double* P, const double x[3], double n[3]) {
    switch (kind) {
    case AEP_LS_GROUND: return 5;
    case AEP_LS_WALL2GROUND: {
        double dz = std::fabs(x[2] - P[2]), dx = std::fabs(P[0] - x[0]), dy = std::fabs(P[1] - x[1]);
        if (dz <= dx && dz <= dy) return 5;
        else if (dy <= dx) return 4;
        else return 2;
    }
    case AEP_LS_SPHERE_GROUND: case AEP_LS_SPHERE: {
        double dx = x[0] - P[0], dy = x[1] - P[1], dz = x[2] - P[2];
        double r = std::sqrt(dx * dx + dy * dy + dz * dz);
        if ((kind == AEP_LS_SPHERE || r - P[3] <= x[2] - P[4]) && r > 0.0) { n[0] = dx / r; n[1] = dy / r; n[2] = dz / r; return 7; }
        return 5;
    }
    case AEP_LS_BOX: {
        double d[6] = { x[2] - P[2], P[5] - x[2], x[0] - P[0], P[3] - x[0], x[1] - P[1], P[4] - x[1] };
        static const int code[6] = { 5, 6, 1, 2, 3, 4 };
        int best = 0; for (int f = 1; f < 6; ++f) if (d[f] < d[best]) best = f;
        return code[best];
    }
    default: return 5;
    }
}

template <typename F>
void parallel_for_planes(int k0, int k1, F&& fn) {
    unsigned nt = std::max(1u, std::min(std::thread::hardware_concurrency(), 32u));
    if (k1 - k0 < 8) nt = 1;
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; ++t)
        th.emplace_back([&, t]() { for (int k = k0 + (int)t; k < k1; k += (int)nt) fn(k); });
    for (auto& x : th) x.join();
}

// code / nrm are in the context's own (held) layout
int upload_levelset(aep_ctx* c, const std::vector<unsigned char>& code, const std::vector<float4>* nrm) {
    cudaSetDevice(c->device);
    if (!c->d_ls_code) CU(dalloc(c, &c->d_ls_code, c->Nheld));
    // on the engine's stream, so that the copy is ordered behind substeps that are already queued (the stream is non-blocking)
    CU(cudaMemcpyAsync(c->d_ls_code, code.data(), c->Nheld, cudaMemcpyHostToDevice, c->stream));
    if (nrm) {
        if (!c->d_ls_nrm) CU(dalloc(c, &c->d_ls_nrm, c->Nheld));
        CU(cudaMemcpyAsync(c->d_ls_nrm, nrm->data(), c->Nheld * sizeof(float4), cudaMemcpyHostToDevice, c->stream));
    }
    CU(cudaStreamSynchronize(c->stream));                                    // the host vectors die with the caller
    const long long bias = (long long)nidx(c->G, c->G.a0[0], c->G.a0[1], c->G.a0[2]);
    c->G.ls_code = c->d_ls_code - bias; c->G.ls_nrm = (nrm || c->d_ls_nrm) ? c->d_ls_nrm - bias : nullptr;
    if (!nrm) c->G.ls_nrm = nullptr;
    c->graph_dirty = true;
    return AEP_OK;
}

// ---------------------------------------------------------------------------------- stepping primitives
int launch_check(aep_ctx* c, const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(c, AEP_ERR_CUDA, "launch of %s failed: %s", what, cudaGetErrorString(e));
    return AEP_OK;
}
#define LAUNCH_OK(name) do { c->launches++; int r_ = launch_check(c, name); if (r_) return r_; } while (0)

int set_device_n(aep_ctx* c, int n_slots, int n_dead) {
    const int v[2] = { n_slots, n_dead };
    CU(cudaMemcpyAsync(&c->d_clk->n_slots, v, sizeof v, cudaMemcpyHostToDevice, c->stream));
    return AEP_OK;
}

// physical re-sort.  slab contexts derive the keys so that dead / out-of-slab slots sort behind the live particles and are dropped.
int do_sort(aep_ctx* c) {
    StageTimer T(c, AEP_STAGE_SORT);
    const bool slab = c->cfg.slab_axis >= 0;
    if (peer_mode(c)) {       // the device owns the count: fetch it (a re-sort is rare; this is its only host synchronisation)
        Telemetry t;
        CU(cudaMemcpyAsync(&t, &c->d_clk->sort_cost, sizeof t, cudaMemcpyDeviceToHost, c->stream)); CU(cudaStreamSynchronize(c->stream));
        c->n = t.n_slots; c->pending_leave = t.n_dead;
    }
    const int n = (int)c->n;
    if (n == 0) return AEP_OK;
    int end_bit = c->key_bits;
    if (slab) {
        k_build_keys_slab<<<cdiv(n, 256), 256, 0, c->stream>>>(c->P[c->cur].a[PX], c->P[c->cur].a[PK], c->d_keys[0], c->d_vals[0], n, c->G,
                                                              c->cfg.slab_axis, c->cfg.slab_lo, c->cfg.slab_hi, c->key_bits);
        LAUNCH_OK("k_build_keys_slab");
        end_bit = c->key_bits + 2;
    } else {
        k_build_keys<<<cdiv(n, 256), 256, 0, c->stream>>>(c->P[c->cur].a[PX], c->d_keys[0], c->d_vals[0], n, c->G);
        LAUNCH_OK("k_build_keys");
    }
    size_t tmp = c->sort_tmp_bytes;
    cudaError_t e = cub::DeviceRadixSort::SortPairs(c->d_sort_tmp, tmp, c->d_keys[0], c->d_keys[1], c->d_vals[0], c->d_vals[1],
                                                    n, 0, end_bit, c->stream);
    if (e != cudaSuccess) return fail(c, AEP_ERR_CUDA, "cub radix sort failed: %s", cudaGetErrorString(e));
    c->launches += 1 + (end_bit + 7) / 8 * 2;     // cub internal kernels (histogram + one onesweep pass per 8 bits), counted approximately
    k_reorder<<<cdiv(n, 256), 256, 0, c->stream>>>(c->P[c->cur], c->P[c->cur ^ 1], c->d_vals[1], n);
    LAUNCH_OK("k_reorder");
    c->cur ^= 1;
    if (slab && c->pending_leave) { c->n -= c->pending_leave; c->pending_leave = 0; }   // leavers were sorted to the tail
    if (int r = set_device_n(c, (int)c->n, 0)) return r;
    cudaMemsetAsync(&c->d_clk->moved_since_sort, 0, 12, c->stream);                     // moved_since_sort, sort_cost
    c->steps_since_sort = 0; c->last_sort_step = c->step_counter; c->sorts += 1;
    c->graph_dirty = true;                                                              // the particle arrays swapped, n may have changed
    return AEP_OK;
}

// the flagged blocks of the run range as a compact list (flags change with every scatter and every halo add)
int list_blocks(aep_ctx* c, bool in_substep) {
    cudaMemsetAsync(c->d_bcount, 0, sizeof(unsigned int), c->stream);
    k_list_blocks<<<cdiv(c->nrun, 256), 256, 0, c->stream>>>(c->G, c->nrun, c->d_blist, c->d_bcount, in_substep ? c->d_clk : nullptr);
    LAUNCH_OK("k_list_blocks");
    return AEP_OK;
}

// stand-alone clear + P2G: first transfer, restarts, stage-level API, caller-driven slab path
int do_p2g(aep_ctx* c, bool first) {
    {
        StageTimer T(c, AEP_STAGE_P2G);
        if (int r = list_blocks(c, false)) return r;
        k_clear_blocks<<<c->pass_ctas, 256, 0, c->stream>>>(c->G, c->d_blist, c->d_bcount);
        LAUNCH_OK("k_clear_blocks");
        if (c->n) {
            p2g_launch(c->stream, c->P[c->cur], c->G, c->n);
            LAUNCH_OK("k_p2g");
        }
    }
    if (c->mesh.nv) {
        StageTimer T(c, AEP_STAGE_MESH);
        if (mesh_own_snapshot(c->mesh, c->d_clk, c->stream, &c->launches)) return fail(c, AEP_ERR_CUDA, "mesh ownership snapshot failed");
        int r = mesh_p2g(c->mesh, c->G, nullptr, c->stream, &c->launches); if (r) return fail(c, AEP_ERR_CUDA, "mesh_p2g failed");
    }
    if (first && c->n) {
        k_init_volumes<<<cdiv(c->n, 256), 256, 0, c->stream>>>(c->P[c->cur], c->G, (int)c->n);
        LAUNCH_OK("k_init_volumes");
    }
    c->grid_mode = 0;
    return AEP_OK;
}

int do_forces(aep_ctx* c, bool in_substep) {
    {   // v_i = p_i / m_i on the active blocks (also feeds the cloth-free case: cheap, <1% of a substep)
        StageTimer T(c, AEP_STAGE_GRID);
        if (int r = list_blocks(c, in_substep)) return r;
        k_grid_normalise<<<c->pass_ctas, 256, 0, c->stream>>>(c->G, c->d_clk, c->d_blist, c->d_bcount);
        LAUNCH_OK("k_grid_normalise");
    }
    if (n_launch(c)) {
        cudaError_t e;
        {   StageTimer T(c, AEP_STAGE_FORCES);
            e = forces_main_launch(c->stream, c->P[c->cur], c->G, c->tm_vt, c->mat, c->d_clk, n_launch(c), c->defer, c->forceA, c->split_forces);
            if (e != cudaSuccess) return fail(c, AEP_ERR_CUDA, "k_forces: %s", cudaGetErrorString(e));
            LAUNCH_OK("k_forces"); }
        {   StageTimer T(c, AEP_STAGE_FORCES_LIST);
            e = forces_list_launch(c->stream, c->P[c->cur], c->G, c->tm_vt, c->mat, c->d_clk, n_launch(c), c->defer, c->forceA, c->split_forces);
            if (e != cudaSuccess) return fail(c, AEP_ERR_CUDA, "k_forces<LIST>: %s", cudaGetErrorString(e));
            LAUNCH_OK("k_forces<LIST>"); }
        if (c->split_forces) {
            StageTimer T(c, AEP_STAGE_FORCE_SCATTER);
            e = force_scatter_launch(c->stream, c->P[c->cur], c->G, c->forceA, n_launch(c), c->d_clk, peer_mode(c));
            if (e != cudaSuccess) return fail(c, AEP_ERR_CUDA, "k_force_scatter: %s", cudaGetErrorString(e));
            LAUNCH_OK("k_force_scatter");
        }
    }
    if (c->mesh.nv) {
        StageTimer T(c, AEP_STAGE_MESH);
        int r = mesh_forces(c->mesh, c->G, c->d_clk, c->stream, &c->launches); if (r) return fail(c, AEP_ERR_CUDA, "mesh_forces failed");
    }
    return AEP_OK;
}

// clear: the fused substep (the list built by do_forces is still valid: nothing scattered mass since)
int do_grid(aep_ctx* c, bool clear) {
    StageTimer T(c, AEP_STAGE_GRID);
    if (!clear) { if (int r = list_blocks(c, false)) return r; }
    if (clear) k_grid_update<true><<<c->pass_ctas, 256, 0, c->stream>>>(c->G, c->d_col, c->d_clk, c->d_blist, c->d_bcount);
    else k_grid_update<false><<<c->pass_ctas, 256, 0, c->stream>>>(c->G, c->d_col, c->d_clk, c->d_blist, c->d_bcount);
    LAUNCH_OK("k_grid_update");
    if (c->mesh.nv && c->mesh.n_fixed) {
        int r = mesh_pin(c->mesh, c->G, c->d_clk, c->stream, &c->launches); if (r) return fail(c, AEP_ERR_CUDA, "mesh_pin failed");
    }
    c->grid_mode = 1;
    return AEP_OK;
}

int do_clock(aep_ctx* c) {
    k_advance_clock<<<1, 1, 0, c->stream>>>(c->d_clk, c->fixed_dt, c->d_col);
    LAUNCH_OK("k_advance_clock");
    return AEP_OK;
}

int peer_mesh_push(aep_ctx* c, int which);
int peer_mesh_wait(aep_ctx* c, int which);

// particles: the fused kernel (G2P + P2G of the next substep), or G2P alone (stage-level API, caller-driven slab path)
int do_g2p_particles(aep_ctx* c, bool scatter) {
    if (!n_launch(c)) return AEP_OK;
    if (scatter && !c->fused) {       // G2P, then the P2G of the next substep as its own kernel (same particles, same order, same result)
        if (int r = do_g2p_particles(c, false)) return r;
        StageTimer T(c, AEP_STAGE_P2G);
        p2g_launch(c->stream, c->P[c->cur], c->G, n_launch(c), c->d_clk, peer_mode(c));
        LAUNCH_OK("k_p2g");
        return AEP_OK;
    }
    cudaError_t e;
    {   StageTimer T(c, scatter ? AEP_STAGE_G2P2G : AEP_STAGE_G2P);
        if (c->mig.axis >= 0 && !peer_mode(c)) cudaMemsetAsync(c->mig.counts, 0, 2 * sizeof(unsigned long long), c->stream);
        e = scatter ? g2p2g_main_launch<true>(c->stream, c->P[c->cur], c->G, c->tm_vt, c->mat, c->d_clk, n_launch(c), c->mig, c->defer)
                    : g2p2g_main_launch<false>(c->stream, c->P[c->cur], c->G, c->tm_vt, c->mat, c->d_clk, n_launch(c), c->mig, c->defer);
        if (e != cudaSuccess) return fail(c, AEP_ERR_CUDA, "k_g2p2g: %s", cudaGetErrorString(e));
        LAUNCH_OK("k_g2p2g"); }
    {   StageTimer T(c, AEP_STAGE_G2P_LIST);
        e = scatter ? g2p2g_list_launch<true>(c->stream, c->P[c->cur], c->G, c->tm_vt, c->mat, c->d_clk, n_launch(c), c->mig, c->defer)
                    : g2p2g_list_launch<false>(c->stream, c->P[c->cur], c->G, c->tm_vt, c->mat, c->d_clk, n_launch(c), c->mig, c->defer);
        if (e != cudaSuccess) return fail(c, AEP_ERR_CUDA, "k_g2p2g<LIST>: %s", cudaGetErrorString(e));
        LAUNCH_OK("k_g2p2g<LIST>"); }
    return AEP_OK;
}
// whole G2P of a context without peers (mesh included)
int do_g2p(aep_ctx* c, bool scatter) {
    if (int r = do_g2p_particles(c, scatter)) return r;
    if (c->mesh.nv) {
        StageTimer T(c, AEP_STAGE_MESH);
        if (mesh_g2p_vertices(c->mesh, c->G, c->d_clk, c->stream, &c->launches)) return fail(c, AEP_ERR_CUDA, "mesh_g2p failed");
        if (mesh_g2p_elements(c->mesh, c->G, c->d_clk, c->stream, &c->launches)) return fail(c, AEP_ERR_CUDA, "mesh_g2p failed");
        if (scatter) { if (mesh_p2g(c->mesh, c->G, c->d_clk, c->stream, &c->launches)) return fail(c, AEP_ERR_CUDA, "mesh_p2g failed"); }
    }
    if (scatter) c->grid_mode = 0;
    return AEP_OK;
}

// Re-sort policy (HS:963-983 rebuilds the weights every substep == re-binning; the physical order is only a performance matter:
// runs of same-cell particles get shorter and fewer half-warps fit their gather tile).
int maybe_sort(aep_ctx* c) {
    c->steps_since_sort += 1; c->step_counter += 1;
    // telemetry of this substep travels to a pinned ring without blocking; decisions use the value of LAG substeps ago (waiting for
    // it keeps the host at most LAG substeps ahead of the device, which never starves the GPU)
    const int slot = (int)(c->step_counter % aep_ctx::RING);
    cudaMemcpyAsync(&c->h_ring[slot], &c->d_clk->sort_cost, sizeof(Telemetry), cudaMemcpyDeviceToHost, c->stream);
    cudaEventRecord(c->ring_ev[slot], c->stream);
    Telemetry seen_t{0.f, (int)c->n, 0, 0};
    const long long seen = c->step_counter - aep_ctx::LAG;
    if (seen > c->last_sort_step && seen > 0) {
        const int ps = (int)(seen % aep_ctx::RING);
        cudaEventSynchronize(c->ring_ev[ps]); seen_t = c->h_ring[ps];
    }
    bool sort_now;
    if (c->cfg.sort_every >= 1) sort_now = c->steps_since_sort >= c->cfg.sort_every;
    else sort_now = seen_t.sort_cost >= (float)c->cfg.sort_cost_threshold || c->steps_since_sort >= 32;
    // a slab context also compacts when dead slots (migrated particles) exceed 1/16 of the array, or the array is nearly full
    if (c->cfg.slab_axis >= 0) {
        const long long dead = peer_mode(c) ? seen_t.n_dead : c->pending_leave, slots = peer_mode(c) ? seen_t.n_slots : c->n;
        if (dead * 16 > slots || (dead > 0 && slots > c->cap - c->cap / 16)) sort_now = true;
    }
    return sort_now ? do_sort(c) : AEP_OK;
}
int compact_slab(aep_ctx* c) {
    if (c->cfg.slab_axis < 0 || !c->inited) return AEP_OK;
    if (!peer_mode(c) && c->pending_leave == 0) return AEP_OK;
    return do_sort(c);
}

int peer_halo_send(aep_ctx* c, int what, int halt_class);
int peer_halo_recv(aep_ctx* c, int what, int halt_class);
int peer_vmax_share(aep_ctx* c, int halt_class);
int peer_vmax_reduce(aep_ctx* c, int halt_class);
int peer_migrate_send(aep_ctx* c);
int peer_migrate_recv(aep_ctx* c);
int phase_begin(aep_ctx* c, int phase);
int phase_end(aep_ctx* c, int phase);

// One substep in phases.  A phase ends where this rank has stored something into its neighbours' memory and begins with waiting for
// what they stored (on the device: flags between processes, stream events inside one process).  Without peers the phases simply run
// back to back: forces | grid update | clock, G2P+P2G (mesh: vertices | elements | P2G) | migration | halo of (m,p).
enum { SUBSTEP_PHASES = 7 };
int enqueue_phase(aep_ctx* c, int ph) {
    int r;
    const bool peer = peer_mode(c), mesh = c->mesh.nv != 0;
    if (peer && (r = phase_begin(c, ph))) return r;
    switch (ph) {
    case 0:
        if (mesh && mesh_own_snapshot(c->mesh, c->d_clk, c->stream, &c->launches)) return fail(c, AEP_ERR_CUDA, "mesh ownership snapshot failed");
        if ((r = do_forces(c, true))) return r;                    // HS:873  (dt of the previous iteration)
        if (peer && (r = peer_halo_send(c, 1, 1))) return r;
        break;
    case 1:
        if (peer && (r = peer_halo_recv(c, 1, 1))) return r;
        if ((r = do_grid(c, true))) return r;                      // HS:877, 899
        if (peer && (r = peer_vmax_share(c, 1))) return r;
        break;
    case 2:
        if (peer && (r = peer_vmax_reduce(c, 1))) return r;
        if ((r = do_clock(c))) return r;                           // HS:878-892
        if ((r = do_g2p_particles(c, true))) return r;             // HS:903-959 and, fused, HS:987 (the order of HS:963-983 never affects results)
        c->grid_mode = 0;
        if (mesh) {
            if (mesh_g2p_vertices(c->mesh, c->G, c->d_clk, c->stream, &c->launches)) return fail(c, AEP_ERR_CUDA, "mesh_g2p failed");
            if (peer && (r = peer_mesh_push(c, 0))) return r;       // elements read the advected vertices of every rank
        }
        break;
    case 3:
        if (mesh) {
            if (peer && (r = peer_mesh_wait(c, 0))) return r;
            if (mesh_g2p_elements(c->mesh, c->G, c->d_clk, c->stream, &c->launches)) return fail(c, AEP_ERR_CUDA, "mesh_g2p failed");
            if (peer && (r = peer_mesh_push(c, 1))) return r;
        }
        break;
    case 4:
        if (mesh) {
            if (peer && (r = peer_mesh_wait(c, 1))) return r;
            if (mesh_p2g(c->mesh, c->G, c->d_clk, c->stream, &c->launches)) return fail(c, AEP_ERR_CUDA, "mesh_p2g failed");
        }
        if (peer && (r = peer_migrate_send(c))) return r;
        break;
    case 5:
        if (peer) { if ((r = peer_migrate_recv(c))) return r; if ((r = peer_halo_send(c, 0, 2))) return r; }
        break;
    case 6:
        if (peer && (r = peer_halo_recv(c, 0, 2))) return r;
        break;
    }
    if (peer && (r = phase_end(c, ph))) return r;
    return AEP_OK;
}
// one substep, launches only (what the graph captures)
int enqueue_substep(aep_ctx* c) {
    for (int ph = 0; ph < SUBSTEP_PHASES; ++ph) if (int r = enqueue_phase(c, ph)) return r;
    return AEP_OK;
}

int do_substep(aep_ctx* c) {
    int r;
    struct Range { bool on; explicit Range(bool o) : on(o) { if (on) nvtxRangePushA("aep:substep"); } ~Range() { if (on) nvtxRangePop(); } } range(c->nvtx);
    if (c->use_graph && !c->profile) {
        if (c->graph_dirty || !c->graph) {
            if (c->graph) { cudaGraphExecDestroy(c->graph); c->graph = nullptr; }
            const long long l0 = c->launches;
            cudaGraph_t g = nullptr;
            CU(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
            r = enqueue_substep(c);
            cudaError_t e = cudaStreamEndCapture(c->stream, &g);
            if (r) { if (g) cudaGraphDestroy(g); return r; }
            if (e != cudaSuccess) return fail(c, AEP_ERR_CUDA, "graph capture of the substep failed: %s", cudaGetErrorString(e));
            e = cudaGraphInstantiate(&c->graph, g, 0);
            cudaGraphDestroy(g);
            if (e != cudaSuccess) return fail(c, AEP_ERR_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(e));
            c->graph_launches = c->launches - l0; c->launches = l0;
            c->graph_dirty = false;
        }
        CU(cudaGraphLaunch(c->graph, c->stream));
        c->launches += c->graph_launches;
    } else if ((r = enqueue_substep(c))) return r;
    return maybe_sort(c);
}

int require_init(aep_ctx* c) {
    if (!c) return AEP_ERR_INVALID;
    if (!c->inited) return fail(c, AEP_ERR_INVALID, "aep_init has not been called");
    cudaSetDevice(c->device);
    return AEP_OK;
}

// sticky device-side error conditions -> return code (checked wherever the host synchronises anyway)
int check_clock(aep_ctx* c, const SimClock& clk) {
    if (clk.comm_timeout) return fail(c, AEP_ERR_STATE, "a neighbouring rank did not deliver its halo / migration / max|v| within the spin limit");
    if (clk.mig_dropped) return fail(c, AEP_ERR_STATE, "%llu migrating particles did not fit the migration buffers or the particle capacity; raise aep_config.particle_capacity / the migration capacity", clk.mig_dropped);
    if (clk.escaped) return fail(c, AEP_ERR_STATE, "%llu particle updates left the grid or became NaN (clamped to the boundary cell); the simulation has blown up", clk.escaped);
    return AEP_OK;
}

// TMA descriptor of G.vt over the planes this context holds: 4-D (component, x, y, z), box 4 x TILE_W x 4 x 4
int make_tensor_map(aep_ctx* c) {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                 CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn encode = nullptr;
    if (!encode) {
        void* fn = nullptr; cudaDriverEntryPointQueryResult q;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
        if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !fn) return fail(c, AEP_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
        encode = (EncodeFn)fn;
    }
    const GridP& G = c->G;
    const cuuint64_t dims[4] = { 4, (cuuint64_t)(G.a1[0] - G.a0[0]), (cuuint64_t)(G.a1[1] - G.a0[1]), (cuuint64_t)(G.a1[2] - G.a0[2]) };
    const cuuint64_t strides[3] = { 16, (cuuint64_t)G.sy * 16, (cuuint64_t)G.sz * 16 };
    const cuuint32_t box[4] = { 4, TILE_W, 4, 4 }, estr[4] = { 1, 1, 1, 1 };
    void* base = (void*)(G.vt + nidx(G, G.a0[0], G.a0[1], G.a0[2]));
    CUresult r = encode(&c->tm_vt, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(c, AEP_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return AEP_OK;
}

}  // namespace

// ======================================================================================================== C ABI
// particle arrays (two copies: the re-sort ping-pongs), sort keys and cub's scratch.  Done by aep_create when the configuration names
// a capacity, so that a solver that is created once and fed many scenes / frames pays for the 23 GB of cudaMalloc once.
static int ensure_particle_capacity(aep_ctx* c, long long cap) {
    if (cap <= c->cap) return AEP_OK;
    if (c->cap) return fail(c, AEP_ERR_INVALID, "particle capacity can only be set once per context");
    for (int b = 0; b < 2; ++b)
        for (int a = 0; a < P_NARR; ++a) CU(dalloc(c, &c->P[b].a[a], (size_t)cap));
    for (int b = 0; b < 2; ++b) { CU(dalloc(c, &c->d_keys[b], (size_t)cap)); CU(dalloc(c, &c->d_vals[b], (size_t)cap)); }
    CU(dalloc(c, &c->defer.list, (size_t)cap)); CU(dalloc(c, &c->defer.count, 1));
    for (int a = 0; a < 3; ++a) CU(dalloc(c, &c->forceA.a[a], (size_t)cap));
    size_t tmp = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp, c->d_keys[0], c->d_keys[1], c->d_vals[0], c->d_vals[1], (int)cap, 0, c->key_bits + 2, c->stream);
    CU(cudaMalloc(&c->d_sort_tmp, tmp)); c->sort_tmp_bytes = tmp;
    c->cap = cap;
    return AEP_OK;
}

extern "C" {

int aep_default_config(aep_config* cfg) {
    if (!cfg) return AEP_ERR_INVALID;
    std::memset(cfg, 0, sizeof *cfg);
    cfg->device = 0; cfg->material = AEP_SAND;
    for (int a = 0; a < 3; ++a) { cfg->grid_min[a] = 0.0; cfg->grid_max[a] = 1.0; cfg->res[a] = 64; }
    cfg->cfl = 0.3; cfg->gravity = 9.8; cfg->collider_friction = 0.2; cfg->snow_hardening = 10.0;
    cfg->sand_h[0] = 35.0; cfg->sand_h[1] = 9.0; cfg->sand_h[2] = 0.2; cfg->sand_h[3] = 10.0;
    cfg->dt_rate_floor = 3e2; cfg->frame_dt = 1.0 / 60.0;
    cfg->particle_capacity = 0; cfg->slab_axis = -1; cfg->slab_lo = 0; cfg->slab_hi = 0; cfg->sort_every = 0; cfg->sort_bricks = 0; cfg->sort_cost_threshold = 0.06; cfg->scatter_strips = 64;
    cfg->vmax_min_mass_fraction = 0.0; cfg->coulomb_friction = 0; cfg->use_graph = 1;
    return AEP_OK;
}

int aep_create(aep_ctx** out, const aep_config* cfg) {
    aep_ctx* c = nullptr;
    if (!out || !cfg) return fail(c, AEP_ERR_INVALID, "null argument");
    *out = nullptr;
    for (int a = 0; a < 3; ++a) {
        if (cfg->res[a] < 4 || cfg->res[a] > 1024) return fail(c, AEP_ERR_INVALID, "grid resolution per axis must be in [4, 1024]");
        if (!(cfg->grid_max[a] > cfg->grid_min[a])) return fail(c, AEP_ERR_INVALID, "maxBound must be bigger than minBound");   // RegularGrid.cpp:125-135
    }
    if (cfg->material != AEP_SNOW && cfg->material != AEP_SAND) return fail(c, AEP_ERR_INVALID, "unknown material");
    if (cfg->slab_axis > 2) return fail(c, AEP_ERR_INVALID, "slab_axis must be -1 (whole grid) or 0..2");
    if (cfg->slab_axis >= 0 && (cfg->slab_lo < 0 || cfg->slab_hi > cfg->res[cfg->slab_axis] || cfg->slab_hi <= cfg->slab_lo))
        return fail(c, AEP_ERR_INVALID, "slab [%d, %d) is not inside the grid", cfg->slab_lo, cfg->slab_hi);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(c, AEP_ERR_CUDA, "no CUDA device available (%s); this engine has no CPU fallback", cudaGetErrorString(e));
    if (cfg->device < 0 || cfg->device >= ndev) return fail(c, AEP_ERR_INVALID, "device %d out of range (%d devices)", cfg->device, ndev);
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, cfg->device);
    if (prop.major < 10) return fail(c, AEP_ERR_CUDA, "device %d is sm_%d%d; libaep_b200 is built for sm_100a only", cfg->device, prop.major, prop.minor);

    aep_ctx* ctx = new aep_ctx();
    ctx->cfg = *cfg; ctx->device = cfg->device; ctx->sm_count = prop.multiProcessorCount; ctx->mig.axis = -1;
    ctx->use_graph = cfg->use_graph != 0 && !getenv("AEP_NO_GRAPH");
    if (const char* f = getenv("AEP_FUSED")) ctx->fused = atoi(f) != 0;
    if (const char* f = getenv("AEP_NVTX")) ctx->nvtx = atoi(f) != 0;
    if (const char* f = getenv("AEP_SPLIT_FORCES")) ctx->split_forces = atoi(f) != 0;
    c = ctx;
    auto bail = [&](int code) { std::string m = ctx->err; aep_destroy(ctx); g_create_error = m; return code; };
#define CUC(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { fail(c, AEP_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); return bail(e_ == cudaErrorMemoryAllocation ? AEP_ERR_ALLOC : AEP_ERR_CUDA); } } while (0)
    CUC(cudaSetDevice(cfg->device));
    CUC(particle_kernels_configure());
    CUC(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    CUC(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    GridP& G = ctx->G;
    G.nx = cfg->res[0]; G.ny = cfg->res[1]; G.nz = cfg->res[2];
    G.nbx = (G.nx + 7) / 8; G.nby = (G.ny + 7) / 8; G.nbz = (G.nz + 7) / 8;
    G.nqx = (G.nx + 3) / 4; G.nqy = (G.ny + 3) / 4; G.bricks = cfg->sort_bricks ? 1 : 0; G.strips = std::max(1, cfg->scatter_strips);
    ctx->nblocks = G.nbx * G.nby * G.nbz;
    {   // Node planes this context holds.  Whole-grid contexts (and x-slabs): everything, the reference's layout.  A y- or z-slab holds
        // its cells' stencil reach [lo-1, hi+1] plus one scratch plane per side (the fused G2P+P2G scatters a particle that has just
        // crossed the slab boundary before it migrates) and lays the slab axis out slowest, so that a node plane is contiguous memory
        // (halo exchange = plain copies) and memory shards with the domain.
        const int nres[3] = { G.nx, G.ny, G.nz };
        for (int a = 0; a < 3; ++a) { G.a0[a] = 0; G.a1[a] = nres[a]; G.v0[a] = 0; G.v1[a] = nres[a]; G.rb0[a] = 0; }
        G.sy = G.nx; G.sz = (long long)G.nx * G.ny;
        const int sa = cfg->slab_axis;
        if (sa == 1 || sa == 2) {
            G.a0[sa] = std::max(0, cfg->slab_lo - 2); G.a1[sa] = std::min(nres[sa], cfg->slab_hi + 3);
            G.v0[sa] = std::max(0, cfg->slab_lo - 1); G.v1[sa] = std::min(nres[sa], cfg->slab_hi + 2);
            if (sa == 1) { G.sy = (long long)G.nx * G.nz; G.sz = G.nx; }
        }
        const int nb[3] = { G.nbx, G.nby, G.nbz };
        for (int a = 0; a < 3; ++a) { G.rb0[a] = 0; G.rbn[a] = nb[a]; }
        if (sa >= 0) {
            const int lo = std::max(0, cfg->slab_lo - 2) >> 3, hi = std::min(nres[sa] - 1, cfg->slab_hi + 2) >> 3;
            G.rb0[sa] = lo; G.rbn[sa] = std::max(1, hi - lo + 1);
        }
        ctx->nrun = G.rbn[0] * G.rbn[1] * G.rbn[2];
    }
    ctx->Ng = (size_t)G.nx * G.ny * G.nz;
    ctx->Nheld = (size_t)(G.a1[0] - G.a0[0]) * (G.a1[1] - G.a0[1]) * (G.a1[2] - G.a0[2]);
    for (int a = 0; a < 3; ++a) ctx->h[a] = (cfg->grid_max[a] - cfg->grid_min[a]) / cfg->res[a];      // RegularGrid.cpp:137-139
    ctx->hmin = std::min(ctx->h[0], std::min(ctx->h[1], ctx->h[2]));
    G.hx = (float)ctx->h[0]; G.hy = (float)ctx->h[1]; G.hz = (float)ctx->h[2];
    G.ihx = (float)(1.0 / ctx->h[0]); G.ihy = (float)(1.0 / ctx->h[1]); G.ihz = (float)(1.0 / ctx->h[2]);
    G.mnx = (float)cfg->grid_min[0]; G.mny = (float)cfg->grid_min[1]; G.mnz = (float)cfg->grid_min[2];
    G.apic = (float)(3.0 / ctx->hmin / ctx->hmin);
    G.inv_cell_vol = (float)(1.0 / (ctx->h[0] * ctx->h[1] * ctx->h[2]));
    G.gravity = (float)cfg->gravity; G.friction = (float)cfg->collider_friction;
    {
        float4 *mp, *f, *vt;
        CUC(dalloc(ctx, &mp, ctx->Nheld)); CUC(dalloc(ctx, &f, ctx->Nheld)); CUC(dalloc(ctx, &vt, ctx->Nheld));
        CUC(cudaMemsetAsync(mp, 0, ctx->Nheld * sizeof(float4), ctx->stream));
        CUC(cudaMemsetAsync(f, 0, ctx->Nheld * sizeof(float4), ctx->stream));
        CUC(cudaMemsetAsync(vt, 0, ctx->Nheld * sizeof(float4), ctx->stream));
        G.mp = mp; G.f = f; G.vt = vt;                       // unbiased for a moment: nidx(a0) below is the bias
        const long long bias = (long long)((long long)G.a0[2] * G.sz + (long long)G.a0[1] * G.sy + G.a0[0]);
        G.mp = mp - bias; G.f = f - bias; G.vt = vt - bias;
    }
    CUC(dalloc(ctx, &G.flags, (size_t)ctx->nblocks));
    CUC(dalloc(ctx, &ctx->d_blist, (size_t)ctx->nrun)); CUC(dalloc(ctx, &ctx->d_bcount, 1));
    ctx->pass_ctas = std::max(1, std::min(ctx->nrun, ctx->sm_count * 8));
    CUC(cudaMemsetAsync(G.flags, 0, (size_t)ctx->nblocks, ctx->stream));
    G.ls_code = nullptr; G.ls_nrm = nullptr;
    CUC(dalloc(ctx, &ctx->d_clk, 1)); CUC(dalloc(ctx, &ctx->d_stats, 8)); CUC(dalloc(ctx, &ctx->d_col, 1));
    SimClock clk{}; clk.frame_dt = cfg->frame_dt; clk.cfl = cfg->cfl; clk.rate_floor = cfg->dt_rate_floor; clk.hmin = ctx->hmin;
    clk.stop_frame = -1; clk.stop_substep = -1;
    CUC(cudaMemcpyAsync(ctx->d_clk, &clk, sizeof clk, cudaMemcpyHostToDevice, ctx->stream));
    ctx->col = ColliderP{}; ctx->col.coulomb = cfg->coulomb_friction ? 1 : 0;
    CUC(cudaMemcpyAsync(ctx->d_col, &ctx->col, sizeof ctx->col, cudaMemcpyHostToDevice, ctx->stream));
    CUC(cudaEventCreate(&ctx->tm.ev[0])); CUC(cudaEventCreate(&ctx->tm.ev[1]));
    CUC(cudaEventCreateWithFlags(&ctx->frame_ev, cudaEventDisableTiming));
    CUC(cudaMallocHost((void**)&ctx->h_ring, aep_ctx::RING * sizeof(Telemetry)));
    for (int i = 0; i < aep_ctx::RING; ++i) { ctx->h_ring[i] = Telemetry{0.f, 0, 0, 0}; CUC(cudaEventCreateWithFlags(&ctx->ring_ev[i], cudaEventDisableTiming)); }
    CUC(cudaMallocHost((void**)&ctx->h_clk, 2 * sizeof(SimClock)));
    for (int i = 0; i < 2; ++i) CUC(cudaEventCreateWithFlags(&ctx->clk_ev[i], cudaEventDisableTiming));
    {   // sort keys: (brick << 6) | cell-in-brick, see sort_key()
        const size_t nkeys = (size_t)G.nqx * G.nqy * ((G.nz + 3) / 4) * 64;
        ctx->key_bits = 1; while (((size_t)1 << ctx->key_bits) < nkeys) ctx->key_bits++;
    }
    if (make_tensor_map(ctx)) return bail(AEP_ERR_CUDA);
    if (cfg->particle_capacity > 0 && ensure_particle_capacity(ctx, cfg->particle_capacity)) return bail(AEP_ERR_ALLOC);
    CUC(cudaStreamSynchronize(ctx->stream));
#undef CUC
    *out = ctx;
    return AEP_OK;
}

int aep_destroy(aep_ctx* c) {
    if (!c) return AEP_ERR_INVALID;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->copy_stream) cudaStreamSynchronize(c->copy_stream);
    if (c->graph) cudaGraphExecDestroy(c->graph);
    for (void* p : c->comm.opened) cudaIpcCloseMemHandle(p);
    for (int i = 0; i < 8; ++i) if (c->comm.ev_phase[i]) cudaEventDestroy(c->comm.ev_phase[i]);
    if (c->comm.block) cudaFree(c->comm.block);
    if (c->comm.d_local) cudaFree(c->comm.d_local);
    for (void* p : c->dev_allocs) cudaFree(p);
    if (c->d_stage) cudaFree(c->d_stage);
    if (c->d_frame) cudaFree(c->d_frame);
    if (c->d_sort_tmp) cudaFree(c->d_sort_tmp);
    mesh_free(c->mesh);
    if (c->h_ring) cudaFreeHost(c->h_ring);
    if (c->h_clk) cudaFreeHost(c->h_clk);
    for (int i = 0; i < 2; ++i) if (c->clk_ev[i]) cudaEventDestroy(c->clk_ev[i]);
    for (int i = 0; i < aep_ctx::RING; ++i) if (c->ring_ev[i]) cudaEventDestroy(c->ring_ev[i]);
    if (c->tm.ev[0]) cudaEventDestroy(c->tm.ev[0]);
    if (c->tm.ev[1]) cudaEventDestroy(c->tm.ev[1]);
    if (c->frame_ev) cudaEventDestroy(c->frame_ev);
    if (c->stream) cudaStreamDestroy(c->stream);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    delete c;
    return AEP_OK;
}

const char* aep_last_error(aep_ctx* c) { return c ? c->err.c_str() : g_create_error.c_str(); }

int aep_sync(aep_ctx* c) {
    if (!c) return AEP_ERR_INVALID;
    cudaSetDevice(c->device);
    CU(cudaStreamSynchronize(c->stream));
    return AEP_OK;
}

int aep_upload_particles(aep_ctx* c, int64_t n, const double* x, const double* v, const double* B1, const double* B2,
                         const double* B3, const double* FE, const double* FP, const double* m, const double* vol,
                         const double* q, double E, double nu, double theta_c, double theta_s) {
    if (!c || n < 0 || (n > 0 && (!x || !v || !B1 || !B2 || !B3 || !FE || !FP || !m || !vol || !q))) return fail(c, AEP_ERR_INVALID, "null particle array");
    if (n >= (1ll << 31) - 64 || c->id_base + n >= (1ll << 31)) return fail(c, AEP_ERR_INVALID, "too many particles for one context / particle ids beyond 2^31");
    cudaSetDevice(c->device);
    if (int r = ensure_particle_capacity(c, std::max<long long>(n, c->cfg.particle_capacity))) return r;
    c->n = n; c->n_ids = n; c->cur = 0; c->pending_leave = 0;
    // material constants (HS:261-265, 634-638)
    const double lambda = E * nu / (1.0 + nu) / (1.0 - 2.0 * nu), mu = E / 2.0 / (1.0 + nu);
    MatParams& M = c->mat;
    M.lambda0 = (float)lambda; M.mu0 = (float)mu; M.xi = (float)c->cfg.snow_hardening;
    M.lo = (float)(1.0 - theta_c); M.hi = (float)(1.0 + theta_s);
    M.h0 = (float)c->cfg.sand_h[0]; M.h1 = (float)c->cfg.sand_h[1]; M.h2 = (float)c->cfg.sand_h[2]; M.h3 = (float)c->cfg.sand_h[3];
    M.k_vol = (float)((3.0 * lambda + 2.0 * mu) / 2.0 / mu); M.material = c->cfg.material;
    // chunked staging: 36 doubles per particle
    const long long CH = 1 << 20;
    int r = ensure_stage(c, (size_t)std::min<long long>(std::max<long long>(n, 1), CH) * 36 * sizeof(double)); if (r) return r;
    for (long long p0 = 0; p0 < n; p0 += CH) {
        const long long cnt = std::min(CH, n - p0);
        double* st = c->d_stage;
        const double* mats[5] = { x, v, B1, B2, B3 };
        for (int k = 0; k < 5; ++k)
            for (int a = 0; a < 3; ++a)
                CU(cudaMemcpyAsync(st + (size_t)(3 * k + a) * cnt, mats[k] + (size_t)a * n + p0, cnt * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemcpyAsync(st + (size_t)15 * cnt, m + p0, cnt * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemcpyAsync(st + (size_t)16 * cnt, vol + p0, cnt * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemcpyAsync(st + (size_t)17 * cnt, q + p0, cnt * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemcpyAsync(st + (size_t)18 * cnt, FE + (size_t)9 * p0, 9 * cnt * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemcpyAsync(st + (size_t)27 * cnt, FP + (size_t)9 * p0, 9 * cnt * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        k_upload_convert<<<cdiv(cnt, 256), 256, 0, c->stream>>>(c->P[0], c->G, st, (int)cnt, (int)p0, c->id_base + p0, c->cfg.grid_min[0], c->cfg.grid_min[1],
                                                               c->cfg.grid_min[2], c->h[0], c->h[1], c->h[2], c->d_clk);
        LAUNCH_OK("k_upload_convert");                                        // the staging buffer is reused in stream order
    }
    if ((r = set_device_n(c, (int)n, 0))) return r;
    {   // opt-in dt rule (NOT the reference): nodes lighter than a fraction of one particle's mass do not enter max|v|
        const float floor_m = (float)(c->cfg.vmax_min_mass_fraction > 0.0 && n > 0 ? c->cfg.vmax_min_mass_fraction * m[0] : 0.0);
        CU(cudaMemcpyAsync(&c->d_clk->vmax_mass_floor, &floor_m, sizeof floor_m, cudaMemcpyHostToDevice, c->stream));
    }
    CU(cudaStreamSynchronize(c->stream));                                    // the caller's arrays are free again
    c->inited = false; c->graph_dirty = true;
    return AEP_OK;
}

int aep_upload_mesh(aep_ctx* c, int64_t nv, int64_t nf, const double* vx, const double* vv, const double* vm, const double* vvol,
                    const double* vB, const int32_t* faces, const double* ev, const double* em, const double* evol, const double* eB,
                    const double* ed, const double* eD, const double* fixedv, double mu, double lambda, double shear_stiffness,
                    double stiffness, double friction_coeff) {
    if (!c) return AEP_ERR_INVALID;
    if (c->comm.exported) return fail(c, AEP_ERR_INVALID, "upload the mesh before aep_comm_export (peers map the mesh block)");
    cudaSetDevice(c->device);
    int r = mesh_upload(c->mesh, c->G, c->cfg.grid_min, c->h, nv, nf, vx, vv, vm, vvol, vB, faces, ev, em, evol, eB, ed, eD, fixedv, mu, lambda,
                        shear_stiffness, stiffness, friction_coeff, c->stream);
    if (r) return fail(c, r == -3 ? AEP_ERR_ALLOC : (r == -1 ? AEP_ERR_INVALID : AEP_ERR_CUDA), "mesh upload failed: %s", cudaGetErrorString(cudaGetLastError()));
    // a slab context owns the mesh points whose cell lies in its slab (every rank holds the whole mesh state)
    c->mesh.own = MeshOwn{ c->cfg.slab_axis, c->cfg.slab_lo, c->cfg.slab_hi };
    c->inited = false; c->graph_dirty = true;
    return AEP_OK;
}

int aep_set_levelset_analytic(aep_ctx* c, int kind, const double* P) {
    if (!c) return AEP_ERR_INVALID;
    if (kind == AEP_LS_NONE) { c->G.ls_code = nullptr; c->G.ls_nrm = nullptr; c->graph_dirty = true; return AEP_OK; }
    if ((kind < AEP_LS_GROUND || kind > AEP_LS_BOX) && kind != AEP_LS_SPHERE) return fail(c, AEP_ERR_INVALID, "unknown analytic level set %d", kind);
    if (!P) return fail(c, AEP_ERR_INVALID, "null level-set parameters");
    const GridP& G = c->G;
    std::vector<unsigned char> code(c->Nheld, 0);
    std::vector<float4> nrm; const bool general = (kind == AEP_LS_SPHERE_GROUND || kind == AEP_LS_SPHERE);
    if (general) nrm.assign(c->Nheld, make_float4(0.f, 0.f, 1.f, 0.f));
    const double* mn = c->cfg.grid_min; const double* h = c->h;
    parallel_for_planes(G.a0[2], G.a1[2], [&](int k) {
        for (int j = G.a0[1]; j < G.a1[1]; ++j) for (int i = G.a0[0]; i < G.a1[0]; ++i) {
            const double gp[3] = { mn[0] + i * h[0], mn[1] + j * h[1], mn[2] + k * h[2] };       // HS:473-476
            if (ls_phi(kind, P, gp) <= 0.0) {                                                     // HS:478
                double n[3] = {0, 0, 1};
                const int cd = ls_normal_code(kind, P, gp, n);
                const size_t id = held_index(c, i, j, k);
                code[id] = (unsigned char)cd;
                if (cd == 7) nrm[id] = make_float4((float)n[0], (float)n[1], (float)n[2], 0.f);
            }
        }
    });
    // remember the analytic form: aep_set_collider_motion switches to evaluating it on the device
    c->col.kind = kind; for (int i = 0; i < 8; ++i) c->col.par[i] = (float)P[i];
    CU(cudaMemcpyAsync(c->d_col, &c->col, sizeof c->col, cudaMemcpyHostToDevice, c->stream));
    return upload_levelset(c, code, general ? &nrm : nullptr);
}

int aep_set_levelset_samples(aep_ctx* c, const uint8_t* inside, const double* normal) {
    if (!c || !inside || !normal) return fail(c, AEP_ERR_INVALID, "null level-set samples");
    const GridP& G = c->G;
    std::vector<unsigned char> code(c->Nheld, 0); std::vector<float4> nrm(c->Nheld, make_float4(0.f, 0.f, 1.f, 0.f));
    const size_t Ng = c->Ng;
    for (int k = G.a0[2]; k < G.a1[2]; ++k) for (int j = G.a0[1]; j < G.a1[1]; ++j) for (int i = G.a0[0]; i < G.a1[0]; ++i) {
        const size_t g = ((size_t)k * G.ny + j) * G.nx + i;                   // the reference's node index (RegularGrid.cpp:164-168)
        if (inside[g]) { const size_t id = held_index(c, i, j, k); code[id] = 7; nrm[id] = make_float4((float)normal[g], (float)normal[Ng + g], (float)normal[2 * Ng + g], 0.f); }
    }
    c->col.kind = AEP_LS_SAMPLED;
    return upload_levelset(c, code, &nrm);
}

int aep_set_collider_motion(aep_ctx* c, const double* velocity3) {
    if (!c) return AEP_ERR_INVALID;
    cudaSetDevice(c->device);
    const bool moving = velocity3 && (velocity3[0] != 0.0 || velocity3[1] != 0.0 || velocity3[2] != 0.0);
    if (moving && (c->col.kind < AEP_LS_GROUND || c->col.kind == AEP_LS_SAMPLED)) return fail(c, AEP_ERR_INVALID, "a moving collider needs an analytic level set (aep_set_levelset_analytic first)");
    c->col.moving = moving ? 1 : 0;
    for (int a = 0; a < 3; ++a) { c->col.vel[a] = moving ? (float)velocity3[a] : 0.f; }
    c->G.cvx = c->col.vel[0]; c->G.cvy = c->col.vel[1]; c->G.cvz = c->col.vel[2]; c->graph_dirty = true;    // GridP travels by value in the launches
    // keep the device's current offset (the clock kernel maintains it); only velocity / mode change
    ColliderP tmp = c->col;
    CU(cudaMemcpyAsync(&c->d_col->moving, &tmp.moving, sizeof(int), cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->d_col->vel, tmp.vel, sizeof tmp.vel, cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return AEP_OK;
}

// ---- peer communication pieces (aep_comm.inl) used by init below
}  // extern "C"
#include "aep_comm.inl"
extern "C" {

int aep_init_begin(aep_ctx* c) {
    if (!c) return AEP_ERR_INVALID;
    if (c->n == 0 && c->mesh.nv == 0 && c->cfg.slab_axis < 0) return fail(c, AEP_ERR_INVALID, "nothing to simulate: upload particles and/or a mesh first");
    cudaSetDevice(c->device);
    c->inited = true; c->pending_leave = 0;
    int r;
    if ((r = do_sort(c))) return r;
    if ((r = do_p2g(c, false))) return r;                                   // HS:854 (mass / momentum part)
    if (peer_mode(c)) { if ((r = peer_halo_send(c, 0, 3))) return r; if ((r = phase_end(c, 0))) return r; }
    return AEP_OK;
}
int aep_init_volumes(aep_ctx* c) {
    int r = require_init(c); if (r) return r;
    if (peer_mode(c)) { if ((r = phase_begin(c, 1))) return r; if ((r = peer_halo_recv(c, 0, 3))) return r; }
    if (c->n) { k_init_volumes<<<cdiv(c->n, 256), 256, 0, c->stream>>>(c->P[c->cur], c->G, (int)c->n); LAUNCH_OK("k_init_volumes"); }   // HS:242-249
    k_vmax_from_mp<<<c->nrun, 256, 0, c->stream>>>(c->G, c->d_clk); LAUNCH_OK("k_vmax_from_mp");
    if (peer_mode(c)) { if ((r = peer_vmax_share(c, 3))) return r; if ((r = phase_end(c, 1))) return r; }
    return AEP_OK;
}
int aep_init_dt_async(aep_ctx* c) {
    int r = require_init(c); if (r) return r;
    if (peer_mode(c)) { if ((r = phase_begin(c, 2))) return r; if ((r = peer_vmax_reduce(c, 3))) return r; }
    k_initial_dt<<<1, 1, 0, c->stream>>>(c->d_clk); LAUNCH_OK("k_initial_dt");  // HS:860
    return AEP_OK;
}
int aep_init_dt(aep_ctx* c) {
    int r = aep_init_dt_async(c); if (r) return r;
    CU(cudaStreamSynchronize(c->stream));
    return AEP_OK;
}
int aep_init(aep_ctx* c) {
    int r;
    if ((r = aep_init_begin(c))) return r;
    if ((r = aep_init_volumes(c))) return r;
    return aep_init_dt(c);
}

int aep_substep(aep_ctx* c) { int r = require_init(c); if (r) return r; return do_substep(c); }

int aep_run(aep_ctx* c, int n_substeps) {
    int r = require_init(c); if (r) return r;
    for (int s = 0; s < n_substeps; ++s) if ((r = do_substep(c))) return r;
    return AEP_OK;
}

// The device stops itself (SimClock::halt): the host only polls.  Substeps that were queued behind the one that completed the last
// requested frame return at once, so the frame is clipped exactly (HS:880-892) although the host runs ahead.
int aep_run_frames(aep_ctx* c, int n_frames, int max_substeps, int64_t* substeps_done) {
    int r = require_init(c); if (r) return r;
    if (n_frames < 0 || max_substeps < 0) return fail(c, AEP_ERR_INVALID, "negative frame / substep count");
    SimClock clk;
    CU(cudaMemcpyAsync(&clk, c->d_clk, sizeof clk, cudaMemcpyDeviceToHost, c->stream)); CU(cudaStreamSynchronize(c->stream));
    const long long s0 = clk.substeps;
    struct Stop { int halt, stop_frame; long long stop_substep; } stop = { 0, clk.frame_no + n_frames, s0 + max_substeps };
    static_assert(offsetof(SimClock, stop_substep) - offsetof(SimClock, halt) == 8, "halt, stop_frame, stop_substep are contiguous");
    CU(cudaMemcpyAsync(&c->d_clk->halt, &stop, sizeof stop, cudaMemcpyHostToDevice, c->stream));
    // a frame takes >= 1/(60 dt_max) = 5 substeps (dt <= cfl/rate_floor): poll the device clock every BURST substeps, one burst late
    // (the clock of burst i is read while burst i+1 is already queued: the device never waits for the host; what was queued behind the
    // halt costs a few empty launches)
    const int BURST = 8;
    bool done = n_frames == 0 || max_substeps == 0;
    long long queued = 0; int pending = -1;
    while (!done) {
        for (int s = 0; s < BURST; ++s) { if ((r = do_substep(c))) return r; ++queued; }
        const int slot = (int)((queued / BURST) & 1);
        CU(cudaMemcpyAsync(&c->h_clk[slot], c->d_clk, sizeof(SimClock), cudaMemcpyDeviceToHost, c->stream));
        CU(cudaEventRecord(c->clk_ev[slot], c->stream));
        if (pending >= 0) { CU(cudaEventSynchronize(c->clk_ev[pending])); done = c->h_clk[pending].halt != 0 || c->h_clk[pending].comm_timeout != 0; }
        pending = slot;
        if (queued > (long long)max_substeps + 2 * BURST) done = true;
    }
    CU(cudaMemcpyAsync(&clk, c->d_clk, sizeof clk, cudaMemcpyDeviceToHost, c->stream)); CU(cudaStreamSynchronize(c->stream));
    stop = Stop{ 0, -1, -1 };
    CU(cudaMemcpyAsync(&c->d_clk->halt, &stop, sizeof stop, cudaMemcpyHostToDevice, c->stream));
    if (substeps_done) *substeps_done = clk.substeps - s0;
    return check_clock(c, clk);
}

int aep_p2g(aep_ctx* c, int first) {
    int r = require_init(c); if (r) return r;
    if ((r = do_sort(c))) return r;
    return do_p2g(c, first != 0);
}
int aep_set_dt(aep_ctx* c, double dt) {
    if (!c) return AEP_ERR_INVALID;
    cudaSetDevice(c->device);
    const float f = (float)dt;
    CU(cudaMemcpyAsync(&c->d_clk->dt, &f, sizeof f, cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return AEP_OK;
}
int aep_set_fixed_dt(aep_ctx* c, double dt) {
    if (!c) return AEP_ERR_INVALID;
    c->graph_dirty = true;
    if (dt > 0.0) { c->fixed_dt = 1; return aep_set_dt(c, dt); }
    c->fixed_dt = 0;
    return AEP_OK;
}
// restart: a saved state is uploaded like a fresh one, then only binned and transferred to the grid -- the particle
// volumes are part of the state (HS:242-249 runs once, on the very first P2G) and the clock is put back by aep_set_clock
int aep_resume(aep_ctx* c) { return aep_init_begin(c); }
int aep_set_clock(aep_ctx* c, double dt, double t, double inner_t, int32_t frame_no, int64_t substeps) {
    if (!c) return AEP_ERR_INVALID;
    if (!(dt > 0.0) || !(t >= 0.0) || !(inner_t >= 0.0) || frame_no < 0 || substeps < 0) return fail(c, AEP_ERR_INVALID, "aep_set_clock: dt must be > 0, times and counters >= 0");
    cudaSetDevice(c->device);
    SimClock clk;
    CU(cudaMemcpyAsync(&clk, c->d_clk, sizeof clk, cudaMemcpyDeviceToHost, c->stream)); CU(cudaStreamSynchronize(c->stream));
    clk.dt = (float)dt; clk.t = t; clk.inner_t = inner_t; clk.frame_no = frame_no; clk.frame_flag = 0; clk.substeps = substeps;
    CU(cudaMemcpyAsync(c->d_clk, &clk, sizeof clk, cudaMemcpyHostToDevice, c->stream)); CU(cudaStreamSynchronize(c->stream));
    return AEP_OK;
}
int aep_set_escaped(aep_ctx* c, int64_t escaped) {
    if (!c || escaped < 0) return AEP_ERR_INVALID;
    cudaSetDevice(c->device);
    const unsigned long long v = (unsigned long long)escaped;
    CU(cudaMemcpyAsync(&c->d_clk->escaped, &v, sizeof v, cudaMemcpyHostToDevice, c->stream)); CU(cudaStreamSynchronize(c->stream));
    return AEP_OK;
}
int aep_stage_forces(aep_ctx* c, double dt) { int r = require_init(c); if (r) return r; if ((r = aep_set_dt(c, dt))) return r; return do_forces(c, false); }
int aep_stage_grid(aep_ctx* c, double dt) {
    int r = require_init(c); if (r) return r; if ((r = aep_set_dt(c, dt))) return r;
    if ((r = do_grid(c, false))) return r;
    k_advance_clock<<<1, 1, 0, c->stream>>>(c->d_clk, 2, nullptr); LAUNCH_OK("k_advance_clock");   // latch vmax, keep dt
    return AEP_OK;
}
int aep_stage_g2p(aep_ctx* c, double dt) { int r = require_init(c); if (r) return r; if ((r = aep_set_dt(c, dt))) return r; return do_g2p(c, false); }

int aep_get_clock(aep_ctx* c, double* dt, double* t, double* inner_t, int32_t* frame_no, int64_t* substeps, double* vmax, int64_t* escaped) {
    if (!c) return AEP_ERR_INVALID;
    cudaSetDevice(c->device);
    SimClock clk;
    CU(cudaMemcpyAsync(&clk, c->d_clk, sizeof clk, cudaMemcpyDeviceToHost, c->stream)); CU(cudaStreamSynchronize(c->stream));
    if (dt) *dt = clk.dt; if (t) *t = clk.t; if (inner_t) *inner_t = clk.inner_t; if (frame_no) *frame_no = clk.frame_no;
    if (substeps) *substeps = clk.substeps; if (vmax) *vmax = clk.vmax_last; if (escaped) *escaped = (int64_t)clk.escaped;
    if (clk.comm_timeout || clk.mig_dropped) return check_clock(c, clk);
    return AEP_OK;
}

int aep_get_counters(aep_ctx* c, int64_t* sorts, int64_t* slots, int64_t* dead, int64_t* moved_since_sort) {
    if (!c) return AEP_ERR_INVALID;
    cudaSetDevice(c->device);
    SimClock clk;
    CU(cudaMemcpyAsync(&clk, c->d_clk, sizeof clk, cudaMemcpyDeviceToHost, c->stream)); CU(cudaStreamSynchronize(c->stream));
    if (sorts) *sorts = c->sorts; if (slots) *slots = clk.n_slots; if (dead) *dead = clk.n_dead; if (moved_since_sort) *moved_since_sort = (int64_t)clk.moved_since_sort;
    return AEP_OK;
}

int aep_get_migration(aep_ctx* c, int64_t* sent, int64_t* received) {
    if (!c) return AEP_ERR_INVALID;
    cudaSetDevice(c->device);
    SimClock clk;
    CU(cudaMemcpyAsync(&clk, c->d_clk, sizeof clk, cudaMemcpyDeviceToHost, c->stream)); CU(cudaStreamSynchronize(c->stream));
    if (sent) *sent = (int64_t)clk.mig_sent; if (received) *received = (int64_t)clk.mig_received;
    return AEP_OK;
}

int64_t aep_num_particles(aep_ctx* c) {      // live particles (dead slots of a slab context excluded)
    if (!c) return -1;
    if (peer_mode(c)) {
        cudaSetDevice(c->device);
        Telemetry t;
        if (cudaMemcpyAsync(&t, &c->d_clk->sort_cost, sizeof t, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess || cudaStreamSynchronize(c->stream) != cudaSuccess) return -1;
        return (int64_t)t.n_slots - t.n_dead;
    }
    return c->n - c->pending_leave;
}

int aep_download_particles(aep_ctx* c, double* x, double* v, double* B1, double* B2, double* B3, double* FE, double* FP, double* vol, double* q) {
    if (!c) return AEP_ERR_INVALID;
    if (c->cfg.slab_axis >= 0) return fail(c, AEP_ERR_INVALID, "aep_download_particles returns the original order of a whole-grid context; a slab context holds a changing subset: use aep_download_particles_local");
    cudaSetDevice(c->device);
    const long long n = c->n; if (n == 0) return AEP_OK;
    const long long CH = 1 << 22;
    int r = ensure_stage(c, (size_t)std::min(n, CH) * 36 * sizeof(double)); if (r) return r;
    for (long long p0 = 0; p0 < n; p0 += CH) {
        const long long cnt = std::min(CH, n - p0);
        double* st = c->d_stage;
        k_download_convert<<<cdiv(n, 256), 256, 0, c->stream>>>(c->P[c->cur], st, (int)n, c->id_base + p0, (int)cnt, c->cfg.grid_min[0], c->cfg.grid_min[1],
                                                               c->cfg.grid_min[2], c->h[0], c->h[1], c->h[2]);
        LAUNCH_OK("k_download_convert");
        double* mats[5] = { x, v, B1, B2, B3 };
        for (int k = 0; k < 5; ++k) if (mats[k])
            for (int a = 0; a < 3; ++a)
                CU(cudaMemcpyAsync(mats[k] + (size_t)a * n + p0, st + (size_t)(3 * k + a) * cnt, cnt * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        if (vol) CU(cudaMemcpyAsync(vol + p0, st + (size_t)15 * cnt, cnt * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        if (q) CU(cudaMemcpyAsync(q + p0, st + (size_t)16 * cnt, cnt * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        if (FE) CU(cudaMemcpyAsync(FE + (size_t)9 * p0, st + (size_t)17 * cnt, 9 * cnt * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        if (FP) CU(cudaMemcpyAsync(FP + (size_t)9 * p0, st + (size_t)26 * cnt, 9 * cnt * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
    }
    return AEP_OK;
}

// positions into the frame staging buffer (device), in id order (whole-grid context) or slot order (slab context)
static int frame_positions_to_device(aep_ctx* c, long long* n_out) {
    int r = compact_slab(c); if (r) return r;
    const long long n = c->n; *n_out = n; if (n == 0) return AEP_OK;
    const size_t bytes = (size_t)n * 3 * sizeof(float);
    if (c->frame_bytes < bytes) {
        if (c->d_frame) { cudaFree(c->d_frame); c->d_frame = nullptr; c->frame_bytes = 0; }
        CU(cudaMalloc((void**)&c->d_frame, bytes)); c->frame_bytes = bytes;
    }
    const bool slab = c->cfg.slab_axis >= 0;
    k_download_positions_f32<<<cdiv(n, 256), 256, 0, c->stream>>>(c->P[c->cur], c->G, c->d_frame, (int)n, slab ? 1 : 0, c->id_base, c->n_ids);
    LAUNCH_OK("k_download_positions_f32");
    return AEP_OK;
}
int aep_download_positions_f32(aep_ctx* c, float* xyz) {
    if (!c || !xyz) return AEP_ERR_INVALID;
    cudaSetDevice(c->device);
    long long n = 0; int r = frame_positions_to_device(c, &n); if (r || n == 0) return r;
    CU(cudaMemcpyAsync(xyz, c->d_frame, (size_t)n * 3 * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return AEP_OK;
}
// asynchronous frame download: the snapshot kernel runs in stream order behind the frame that has just completed, the copy to the
// caller's (pinned) buffer on a second stream while the next frame computes.  aep_frame_positions_wait before reading the buffer.
int aep_frame_positions_begin(aep_ctx* c, float* pinned_xyz) {
    if (!c || !pinned_xyz) return AEP_ERR_INVALID;
    cudaSetDevice(c->device);
    if (c->frame_pending) { CU(cudaStreamSynchronize(c->copy_stream)); c->frame_pending = false; }   // the staging buffer is single
    long long n = 0; int r = frame_positions_to_device(c, &n); if (r || n == 0) return r;
    CU(cudaEventRecord(c->frame_ev, c->stream));
    CU(cudaStreamWaitEvent(c->copy_stream, c->frame_ev, 0));
    CU(cudaMemcpyAsync(pinned_xyz, c->d_frame, (size_t)n * 3 * sizeof(float), cudaMemcpyDeviceToHost, c->copy_stream));
    // the next snapshot must not overwrite the staging buffer before this copy has read it
    CU(cudaEventRecord(c->frame_ev, c->copy_stream));
    CU(cudaStreamWaitEvent(c->stream, c->frame_ev, 0));
    c->frame_pending = true;
    return AEP_OK;
}
// page-locked host memory for the asynchronous downloads (the host layer is plain C++ and has no CUDA runtime of its own)
void* aep_host_alloc(int64_t bytes) { void* p = nullptr; return (bytes > 0 && cudaMallocHost(&p, (size_t)bytes) == cudaSuccess) ? p : nullptr; }
void aep_host_free(void* p) { if (p) cudaFreeHost(p); }
int aep_frame_positions_wait(aep_ctx* c) {
    if (!c) return AEP_ERR_INVALID;
    cudaSetDevice(c->device);
    if (c->frame_pending) { CU(cudaStreamSynchronize(c->copy_stream)); c->frame_pending = false; }
    return AEP_OK;
}

int aep_download_grid(aep_ctx* c, double* m, double* v, double* f, double* vt) {
    if (!c) return AEP_ERR_INVALID;
    cudaSetDevice(c->device);
    const long long Ng = (long long)c->Ng, CH = 1ll << 22;
    int r = ensure_stage(c, (size_t)std::min(Ng, CH) * 10 * sizeof(double)); if (r) return r;
    for (long long n0 = 0; n0 < Ng; n0 += CH) {
        const long long cnt = std::min(CH, Ng - n0);
        double* sm = c->d_stage; double* sv = sm + cnt; double* sf = sv + 3 * cnt; double* svt = sf + 3 * cnt;
        k_download_grid<<<cdiv(cnt, 256), 256, 0, c->stream>>>(c->G, sm, sv, sf, svt, n0, cnt, c->grid_mode);
        LAUNCH_OK("k_download_grid");
        if (c->grid_mode == 0) CU(cudaMemsetAsync(svt, 0, 3 * cnt * sizeof(double), c->stream));
        if (m) CU(cudaMemcpyAsync(m + n0, sm, cnt * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        for (int a = 0; a < 3; ++a) {
            if (v) CU(cudaMemcpyAsync(v + (size_t)a * Ng + n0, sv + (size_t)a * cnt, cnt * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
            if (f) CU(cudaMemcpyAsync(f + (size_t)a * Ng + n0, sf + (size_t)a * cnt, cnt * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
            if (vt) CU(cudaMemcpyAsync(vt + (size_t)a * Ng + n0, svt + (size_t)a * cnt, cnt * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        }
        CU(cudaStreamSynchronize(c->stream));
    }
    return AEP_OK;
}

int aep_download_mesh(aep_ctx* c, double* vx, double* vv, double* vB, double* ex, double* ev, double* eB, double* ed) {
    if (!c) return AEP_ERR_INVALID;
    cudaSetDevice(c->device);
    int r = mesh_download(c->mesh, vx, vv, vB, ex, ev, eB, ed, c->stream);
    if (r) return fail(c, r, "mesh download failed: %s", cudaGetErrorString(cudaGetLastError()));
    return AEP_OK;
}

int aep_stats(aep_ctx* c, double* com3, double* kinetic, double* mean_jp, double* mass) {
    if (!c) return AEP_ERR_INVALID;
    cudaSetDevice(c->device);
    double h[7] = {0, 0, 0, 0, 0, 0, 0};
    if (n_launch(c)) {
        CU(cudaMemsetAsync(c->d_stats, 0, 7 * sizeof(double), c->stream));
        k_stats<<<std::min(cdiv(n_launch(c), 256), 148 * 8), 256, 0, c->stream>>>(c->P[c->cur], c->G, c->d_stats, c->d_clk);
        LAUNCH_OK("k_stats");
        CU(cudaMemcpyAsync(h, c->d_stats, sizeof h, cudaMemcpyDeviceToHost, c->stream)); CU(cudaStreamSynchronize(c->stream));
    }
    if (com3) for (int a = 0; a < 3; ++a) com3[a] = h[5] > 0 ? h[a] / h[5] : 0.0;
    if (kinetic) *kinetic = h[3];
    if (mean_jp) *mean_jp = h[6] > 0 ? h[4] / h[6] : 0.0;
    if (mass) *mass = h[5];
    return AEP_OK;
}

int aep_grid_activity(aep_ctx* c, int64_t* active_blocks, int64_t* active_nodes) {
    if (!c) return AEP_ERR_INVALID;
    cudaSetDevice(c->device);
    unsigned long long h[2] = {0, 0};
    CU(cudaMemsetAsync(c->d_stats, 0, 2 * sizeof(unsigned long long), c->stream));
    k_count_active<<<c->nrun, 256, 0, c->stream>>>(c->G, (unsigned long long*)c->d_stats);
    LAUNCH_OK("k_count_active");
    CU(cudaMemcpyAsync(h, c->d_stats, sizeof h, cudaMemcpyDeviceToHost, c->stream)); CU(cudaStreamSynchronize(c->stream));
    if (active_blocks) *active_blocks = (int64_t)h[0];
    if (active_nodes) *active_nodes = (int64_t)h[1];
    return AEP_OK;
}

int64_t aep_kernel_launches(aep_ctx* c) { return c ? c->launches : -1; }
void* aep_stream(aep_ctx* c) { return c ? (void*)c->stream : nullptr; }
int aep_profile(aep_ctx* c, int enable) {
    if (!c) return AEP_ERR_INVALID;
    c->profile = enable != 0;
    for (int i = 0; i < AEP_NUM_STAGES; ++i) { c->tm.ms[i] = 0; c->tm.calls[i] = 0; }
    return AEP_OK;
}
int aep_get_timers(aep_ctx* c, double* ms, int64_t* calls) {
    if (!c) return AEP_ERR_INVALID;
    for (int i = 0; i < AEP_NUM_STAGES; ++i) { if (ms) ms[i] = c->tm.ms[i]; if (calls) calls[i] = c->tm.calls[i]; }
    return AEP_OK;
}

// ---- split stepping for the caller-driven multi-GPU path (un-fused kernels: P2G runs after the caller has moved the migrants)
int aep_step_forces(aep_ctx* c) { int r = require_init(c); if (r) return r; return do_forces(c, false); }
int aep_step_grid(aep_ctx* c) { int r = require_init(c); if (r) return r; return do_grid(c, false); }
int aep_step_g2p(aep_ctx* c) {
    int r = require_init(c); if (r) return r;
    if ((r = do_clock(c))) return r;
    return do_g2p(c, false);
}
int aep_step_p2g(aep_ctx* c) {
    int r = require_init(c); if (r) return r;
    if ((r = maybe_sort(c))) return r;
    return do_p2g(c, false);
}
// halo exchange / migration entry points of the caller-driven path
#include "aep_halo.inl"

}  // extern "C"
