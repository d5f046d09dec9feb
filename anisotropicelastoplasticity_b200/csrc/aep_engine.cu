// aep_engine.cu -- context management, stepping sequence and the C ABI (include/aep_b200.h) of libaep_b200.so.
//
// One aep_ctx = one GPU = one slab of the domain.  All work is enqueued on ctx->stream; a substep is
//   forces -> grid update/collide (+max|v|) -> clock (dt rule) -> G2P/advect/F/plasticity -> re-bin -> P2G
// (HybridSolver.cpp:867-1032) with no host synchronisation, dt living in device memory.
#include <cuda_runtime.h>
#include <cub/device/device_radix_sort.cuh>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/aep_b200.h"
#include "aep_kernels.cuh"
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

}  // namespace

struct aep_ctx {
    aep_config cfg;
    int device = 0, sm_count = 148;
    cudaStream_t stream = nullptr;
    std::string err;
    long long launches = 0;

    // grid
    GridP G{};
    size_t Ng = 0; int nblocks = 0;             // all 8^3 blocks of the grid (flags array)
    int nrun = 0;                                // blocks the grid passes run over (G.rb0 / G.rbn)
    unsigned int *d_blist = nullptr, *d_bcount = nullptr;   // compact list of the flagged blocks among them (k_list_blocks)
    int pass_ctas = 1;                           // persistent grid of the grid passes
    double h[3]{}, hmin = 0;
    std::vector<void*> dev_allocs;
    unsigned char* d_ls_code = nullptr; float4* d_ls_nrm = nullptr;
    int grid_mode = 0;                          // 0: (m,p) fresh from P2G; 1: vt valid

    // particles
    long long n = 0, cap = 0;
    PartP P[2]{}; int cur = 0;
    unsigned int *d_keys[2] = {nullptr, nullptr}, *d_vals[2] = {nullptr, nullptr};
    void* d_sort_tmp = nullptr; size_t sort_tmp_bytes = 0;
    int key_bits = 0;
    MatParams mat{};
    bool keys_valid = false;
    int steps_since_sort = 0;
    long long pending_leave = 0;                // particles extracted for migration, dropped at the next re-bin
    MigList mig{};                              // leaver lists filled by k_g2p (aep_migrate_bind); axis < 0 when unbound
    float4* mig_buf[2] = {nullptr, nullptr};    // caller-owned send buffers
    long long id_base = 0;

    // mesh
    MeshState mesh;

    // clock
    SimClock* d_clk = nullptr;
    double* d_stats = nullptr;
    int fixed_dt = 0;

    // staging
    double* d_stage = nullptr; size_t stage_bytes = 0;

    // adaptive re-sort: lagged, non-blocking readback of SimClock::sort_cost
    static constexpr int RING = 4, LAG = 2;
    float* h_ring = nullptr; cudaEvent_t ring_ev[RING] = {nullptr, nullptr, nullptr, nullptr};
    long long step_counter = 0, last_sort_step = -1;

    bool profile = false; Timers tm{};
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

struct StageTimer {
    aep_ctx* c; int stage;
    StageTimer(aep_ctx* c_, int s) : c(c_), stage(s) { if (c->profile) cudaEventRecord(c->tm.ev[0], c->stream); }
    ~StageTimer() {
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

// ---------------------------------------------------------------------------------- level-set sampling (host)
// HS:473-482 evaluates phi / grad phi at grid nodes only, and colliders are static (HS:484), so the level set is
// sampled once per setLevelSet in fp64 on the host (same formulas as LevelSet.cpp:8-42) and uploaded as a byte code.
double ls_phi(int kind, const double* P, const double x[3]) {
    switch (kind) {
    case AEP_LS_GROUND: return x[2] - P[0];
    case AEP_LS_WALL2GROUND: return std::min(std::min(x[2] - P[2], P[0] - x[0]), P[1] - x[1]);
    case AEP_LS_SPHERE_GROUND: {
        double dx = x[0] - P[0], dy = x[1] - P[1], dz = x[2] - P[2];
        return std::min(std::sqrt(dx * dx + dy * dy + dz * dz) - P[3], x[2] - P[4]);
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
    case AEP_LS_SPHERE_GROUND: {
        double dx = x[0] - P[0], dy = x[1] - P[1], dz = x[2] - P[2];
        double r = std::sqrt(dx * dx + dy * dy + dz * dz);
        if (r - P[3] <= x[2] - P[4] && r > 0.0) { n[0] = dx / r; n[1] = dy / r; n[2] = dz / r; return 7; }
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
void parallel_for_planes(int nz, F&& fn) {
    unsigned nt = std::max(1u, std::min(std::thread::hardware_concurrency(), 32u));
    if (nz < 8) nt = 1;
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; ++t)
        th.emplace_back([&, t]() { for (int k = (int)t; k < nz; k += (int)nt) fn(k); });
    for (auto& x : th) x.join();
}

int upload_levelset(aep_ctx* c, const std::vector<unsigned char>& code, const std::vector<float4>* nrm) {
    cudaSetDevice(c->device);
    if (!c->d_ls_code) CU(dalloc(c, &c->d_ls_code, c->Ng));
    CU(cudaMemcpy(c->d_ls_code, code.data(), c->Ng, cudaMemcpyHostToDevice));
    if (nrm) {
        if (!c->d_ls_nrm) CU(dalloc(c, &c->d_ls_nrm, c->Ng));
        CU(cudaMemcpy(c->d_ls_nrm, nrm->data(), c->Ng * sizeof(float4), cudaMemcpyHostToDevice));
    }
    c->G.ls_code = c->d_ls_code; c->G.ls_nrm = c->d_ls_nrm;
    return AEP_OK;
}

// ---------------------------------------------------------------------------------- stepping primitives
int launch_check(aep_ctx* c, const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(c, AEP_ERR_CUDA, "launch of %s failed: %s", what, cudaGetErrorString(e));
    return AEP_OK;
}
#define LAUNCH_OK(name) do { c->launches++; int r_ = launch_check(c, name); if (r_) return r_; } while (0)

int do_sort(aep_ctx* c, bool build_keys) {
    StageTimer T(c, AEP_STAGE_SORT);
    const int n = (int)c->n;
    if (n == 0) return AEP_OK;
    const bool slab = c->cfg.slab_axis >= 0;
    int end_bit = c->key_bits;
    if (build_keys && slab) {
        k_build_keys_slab<<<cdiv(n, 256), 256, 0, c->stream>>>(c->P[c->cur].a[PX], c->P[c->cur].a[PQ1], c->d_keys[0], c->d_vals[0], n, c->G,
                                                              c->cfg.slab_axis, c->cfg.slab_lo, c->cfg.slab_hi, c->key_bits);
        LAUNCH_OK("k_build_keys_slab");
        end_bit = c->key_bits + 2;
    } else if (build_keys) {
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
    if (slab && build_keys && c->pending_leave) { c->n -= c->pending_leave; c->pending_leave = 0; }   // leavers were sorted to the tail
    cudaMemsetAsync(&c->d_clk->moved_since_sort, 0, 16, c->stream);                                    // moved_since_sort, sort_cost
    c->steps_since_sort = 0; c->last_sort_step = c->step_counter;
    return AEP_OK;
}

// the flagged blocks of the run range as a compact list (flags change with every P2G and every halo add: rebuilt per pass, ~5 us)
int list_blocks(aep_ctx* c) {
    cudaMemsetAsync(c->d_bcount, 0, sizeof(unsigned int), c->stream);
    k_list_blocks<<<cdiv(c->nrun, 256), 256, 0, c->stream>>>(c->G, c->nrun, c->d_blist, c->d_bcount);
    LAUNCH_OK("k_list_blocks");
    return AEP_OK;
}

int do_p2g(aep_ctx* c, bool first) {
    {
        StageTimer T(c, AEP_STAGE_P2G);
        if (int r = list_blocks(c)) return r;
        k_clear_blocks<<<c->pass_ctas, 256, 0, c->stream>>>(c->G, c->d_blist, c->d_bcount);
        LAUNCH_OK("k_clear_blocks");
        if (c->n) {
            p2g_launch(c->stream, c->P[c->cur], c->G, c->n);
            LAUNCH_OK("k_p2g");
        }
    }
    if (c->mesh.nv) {
        StageTimer T(c, AEP_STAGE_MESH);
        int r = mesh_p2g(c->mesh, c->G, c->stream, &c->launches); if (r) return fail(c, AEP_ERR_CUDA, "mesh_p2g failed");
    }
    if (first && c->n) {
        k_init_volumes<<<cdiv(c->n, 256), 256, 0, c->stream>>>(c->P[c->cur], c->G, (int)c->n);
        LAUNCH_OK("k_init_volumes");
    }
    c->grid_mode = 0;
    return AEP_OK;
}

int do_forces(aep_ctx* c) {
    {   // v_i = p_i / m_i on the active blocks (also feeds the cloth-free case: cheap, <1% of a substep)
        StageTimer T(c, AEP_STAGE_GRID);
        if (int r = list_blocks(c)) return r;
        k_grid_normalise<<<c->pass_ctas, 256, 0, c->stream>>>(c->G, c->d_blist, c->d_bcount);
        LAUNCH_OK("k_grid_normalise");
    }
    if (c->n) {
        StageTimer T(c, AEP_STAGE_FORCES);
        forces_launch(c->stream, c->sm_count, c->P[c->cur], c->G, c->mat, c->d_clk, c->n);
        LAUNCH_OK("k_forces");
    }
    if (c->mesh.nv) {
        StageTimer T(c, AEP_STAGE_MESH);
        int r = mesh_forces(c->mesh, c->G, c->stream, &c->launches); if (r) return fail(c, AEP_ERR_CUDA, "mesh_forces failed");
    }
    return AEP_OK;
}

int do_grid(aep_ctx* c) {
    StageTimer T(c, AEP_STAGE_GRID);
    if (int r = list_blocks(c)) return r;
    k_grid_update<<<c->pass_ctas, 256, 0, c->stream>>>(c->G, c->d_clk, c->d_blist, c->d_bcount);
    LAUNCH_OK("k_grid_update");
    if (c->mesh.nv && c->mesh.n_fixed) {
        int r = mesh_pin(c->mesh, c->G, c->stream, &c->launches); if (r) return fail(c, AEP_ERR_CUDA, "mesh_pin failed");
    }
    c->grid_mode = 1;
    return AEP_OK;
}

int do_clock(aep_ctx* c) {
    k_advance_clock<<<1, 1, 0, c->stream>>>(c->d_clk, c->fixed_dt, (int)c->n);
    LAUNCH_OK("k_advance_clock");
    return AEP_OK;
}

int do_g2p(aep_ctx* c) {
    if (c->n) {
        StageTimer T(c, AEP_STAGE_G2P);
        if (c->mig.axis >= 0) cudaMemsetAsync(c->mig.counts, 0, 2 * sizeof(unsigned long long), c->stream);
        g2p_launch(c->stream, c->sm_count, c->P[c->cur], c->G, c->mat, c->d_clk, c->d_keys[0], c->d_vals[0], c->n, c->mig);
        LAUNCH_OK("k_g2p");
    }
    if (c->mesh.nv) {
        StageTimer T(c, AEP_STAGE_MESH);
        int r = mesh_g2p(c->mesh, c->G, c->d_clk, c->stream, &c->launches); if (r) return fail(c, AEP_ERR_CUDA, "mesh_g2p failed");
    }
    return AEP_OK;
}

// Re-sort policy (HS:963-983 rebuilds the weights every substep == re-binning; the physical order is only a performance matter).
// build_keys: slab contexts derive the keys from the positions so that dead / out-of-slab slots sort behind the live particles.
int maybe_sort(aep_ctx* c, bool build_keys) {
    c->steps_since_sort += 1; c->step_counter += 1;
    bool sort_now;
    if (c->cfg.sort_every >= 1) sort_now = c->steps_since_sort >= c->cfg.sort_every;
    else {
        // adaptive: sort_cost of this substep travels to a pinned ring without blocking; the decision uses the value of LAG substeps
        // ago (waiting for it keeps the host at most LAG substeps ahead of the device, which never starves the GPU).
        const int slot = (int)(c->step_counter % aep_ctx::RING);
        cudaMemcpyAsync(&c->h_ring[slot], &c->d_clk->sort_cost, sizeof(float), cudaMemcpyDeviceToHost, c->stream);
        cudaEventRecord(c->ring_ev[slot], c->stream);
        float cost = 0.0f;
        const long long seen = c->step_counter - aep_ctx::LAG;
        if (seen > c->last_sort_step && seen > 0) {
            const int ps = (int)(seen % aep_ctx::RING);
            cudaEventSynchronize(c->ring_ev[ps]); cost = c->h_ring[ps];
        }
        sort_now = cost >= (float)c->cfg.sort_cost_threshold || c->steps_since_sort >= 32;
    }
    // a slab context also compacts when dead slots (migrated particles) exceed 1/16 of the array
    if (c->cfg.slab_axis >= 0 && c->pending_leave * 16 > c->n) sort_now = true;
    return sort_now ? do_sort(c, build_keys) : AEP_OK;
}
int compact_slab(aep_ctx* c) {
    if (c->cfg.slab_axis < 0 || c->pending_leave == 0 || !c->inited) return AEP_OK;
    return do_sort(c, true);
}

int do_substep(aep_ctx* c) {
    int r;
    if ((r = do_forces(c))) return r;       // HS:873  (dt of the previous iteration)
    if ((r = do_grid(c))) return r;         // HS:877, 899
    if ((r = do_clock(c))) return r;        // HS:878-892
    if ((r = do_g2p(c))) return r;          // HS:903-959
    // HS:963-983: weights at the new positions == re-binning.  Correctness never depends on the order (runs end on a cell
    // change, reductions are atomic); sorting only keeps runs long and gathers local, so it may be done every k-th substep.
    if ((r = maybe_sort(c, false))) return r;
    if ((r = do_p2g(c, false))) return r;   // HS:987
    return AEP_OK;
}

int require_init(aep_ctx* c) {
    if (!c) return AEP_ERR_INVALID;
    if (!c->inited) return fail(c, AEP_ERR_INVALID, "aep_init has not been called");
    cudaSetDevice(c->device);
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
    cfg->particle_capacity = 0; cfg->slab_axis = -1; cfg->slab_lo = 0; cfg->slab_hi = 0; cfg->sort_every = 0; cfg->sort_bricks = 0; cfg->sort_cost_threshold = 0.5; cfg->scatter_strips = 64;
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
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(c, AEP_ERR_CUDA, "no CUDA device available (%s); this engine has no CPU fallback", cudaGetErrorString(e));
    if (cfg->device < 0 || cfg->device >= ndev) return fail(c, AEP_ERR_INVALID, "device %d out of range (%d devices)", cfg->device, ndev);
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, cfg->device);
    if (prop.major < 10) return fail(c, AEP_ERR_CUDA, "device %d is sm_%d%d; libaep_b200 is built for sm_100a only", cfg->device, prop.major, prop.minor);

    aep_ctx* ctx = new aep_ctx();
    ctx->cfg = *cfg; ctx->device = cfg->device; ctx->sm_count = prop.multiProcessorCount; ctx->mig.axis = -1;
    c = ctx;
    auto bail = [&](int code) { std::string m = ctx->err; aep_destroy(ctx); g_create_error = m; return code; };
#define CUC(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { fail(c, AEP_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); return bail(e_ == cudaErrorMemoryAllocation ? AEP_ERR_ALLOC : AEP_ERR_CUDA); } } while (0)
    CUC(cudaSetDevice(cfg->device));
    CUC(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    GridP& G = ctx->G;
    G.nx = cfg->res[0]; G.ny = cfg->res[1]; G.nz = cfg->res[2];
    G.nbx = (G.nx + 7) / 8; G.nby = (G.ny + 7) / 8; G.nbz = (G.nz + 7) / 8;
    G.nqx = (G.nx + 3) / 4; G.nqy = (G.ny + 3) / 4; G.bricks = cfg->sort_bricks ? 1 : 0; G.strips = std::max(1, cfg->scatter_strips);
    ctx->nblocks = G.nbx * G.nby * G.nbz;
    {   // grid passes visit the blocks this context can touch: everything, or the slab's node planes slab_lo-1 .. slab_hi+1
        const int nb[3] = { G.nbx, G.nby, G.nbz };
        for (int a = 0; a < 3; ++a) { G.rb0[a] = 0; G.rbn[a] = nb[a]; }
        if (cfg->slab_axis >= 0 && cfg->slab_axis < 3) {
            const int a = cfg->slab_axis;
            const int lo = std::max(0, cfg->slab_lo - 1) >> 3, hi = std::min(cfg->res[a] - 1, cfg->slab_hi + 1) >> 3;
            G.rb0[a] = lo; G.rbn[a] = std::max(1, hi - lo + 1);
        }
        ctx->nrun = G.rbn[0] * G.rbn[1] * G.rbn[2];
    }
    ctx->Ng = (size_t)G.nx * G.ny * G.nz;
    for (int a = 0; a < 3; ++a) ctx->h[a] = (cfg->grid_max[a] - cfg->grid_min[a]) / cfg->res[a];      // RegularGrid.cpp:137-139
    ctx->hmin = std::min(ctx->h[0], std::min(ctx->h[1], ctx->h[2]));
    G.hx = (float)ctx->h[0]; G.hy = (float)ctx->h[1]; G.hz = (float)ctx->h[2];
    G.ihx = (float)(1.0 / ctx->h[0]); G.ihy = (float)(1.0 / ctx->h[1]); G.ihz = (float)(1.0 / ctx->h[2]);
    G.mnx = (float)cfg->grid_min[0]; G.mny = (float)cfg->grid_min[1]; G.mnz = (float)cfg->grid_min[2];
    G.apic = (float)(3.0 / ctx->hmin / ctx->hmin);
    G.inv_cell_vol = (float)(1.0 / (ctx->h[0] * ctx->h[1] * ctx->h[2]));
    G.gravity = (float)cfg->gravity; G.friction = (float)cfg->collider_friction;
    CUC(dalloc(ctx, &G.mp, ctx->Ng)); CUC(dalloc(ctx, &G.f, ctx->Ng)); CUC(dalloc(ctx, &G.vt, ctx->Ng));
    CUC(dalloc(ctx, &G.flags, (size_t)ctx->nblocks));
    CUC(dalloc(ctx, &ctx->d_blist, (size_t)ctx->nrun)); CUC(dalloc(ctx, &ctx->d_bcount, 1));
    ctx->pass_ctas = std::max(1, std::min(ctx->nrun, ctx->sm_count * 8));
    CUC(cudaMemsetAsync(G.mp, 0, ctx->Ng * sizeof(float4), ctx->stream));
    CUC(cudaMemsetAsync(G.f, 0, ctx->Ng * sizeof(float4), ctx->stream));
    CUC(cudaMemsetAsync(G.vt, 0, ctx->Ng * sizeof(float4), ctx->stream));
    CUC(cudaMemsetAsync(G.flags, 0, (size_t)ctx->nblocks, ctx->stream));
    G.ls_code = nullptr; G.ls_nrm = nullptr;
    CUC(dalloc(ctx, &ctx->d_clk, 1)); CUC(dalloc(ctx, &ctx->d_stats, 8));
    SimClock clk{}; clk.frame_dt = cfg->frame_dt; clk.cfl = cfg->cfl; clk.rate_floor = cfg->dt_rate_floor; clk.hmin = ctx->hmin;
    CUC(cudaMemcpyAsync(ctx->d_clk, &clk, sizeof clk, cudaMemcpyHostToDevice, ctx->stream));
    CUC(cudaEventCreate(&ctx->tm.ev[0])); CUC(cudaEventCreate(&ctx->tm.ev[1]));
    CUC(cudaMallocHost((void**)&ctx->h_ring, aep_ctx::RING * sizeof(float)));
    for (int i = 0; i < aep_ctx::RING; ++i) { ctx->h_ring[i] = 0.0f; CUC(cudaEventCreateWithFlags(&ctx->ring_ev[i], cudaEventDisableTiming)); }
    {   // sort keys: (brick << 6) | cell-in-brick, see sort_key()
        const size_t nkeys = (size_t)G.nqx * G.nqy * ((G.nz + 3) / 4) * 64;
        ctx->key_bits = 1; while (((size_t)1 << ctx->key_bits) < nkeys) ctx->key_bits++;
    }
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
    for (void* p : c->dev_allocs) cudaFree(p);
    if (c->d_stage) cudaFree(c->d_stage);
    if (c->d_sort_tmp) cudaFree(c->d_sort_tmp);
    mesh_free(c->mesh);
    if (c->h_ring) cudaFreeHost(c->h_ring);
    for (int i = 0; i < aep_ctx::RING; ++i) if (c->ring_ev[i]) cudaEventDestroy(c->ring_ev[i]);
    if (c->tm.ev[0]) cudaEventDestroy(c->tm.ev[0]);
    if (c->tm.ev[1]) cudaEventDestroy(c->tm.ev[1]);
    if (c->stream) cudaStreamDestroy(c->stream);
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
    if (n >= (1ll << 31) - 64) return fail(c, AEP_ERR_INVALID, "too many particles for one context");
    cudaSetDevice(c->device);
    if (int r = ensure_particle_capacity(c, std::max<long long>(n, c->cfg.particle_capacity))) return r;
    c->n = n; c->cur = 0;
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
    CU(cudaStreamSynchronize(c->stream));                                    // the caller's arrays are free again
    c->inited = false;
    return AEP_OK;
}

int aep_upload_mesh(aep_ctx* c, int64_t nv, int64_t nf, const double* vx, const double* vv, const double* vm, const double* vvol,
                    const double* vB, const int32_t* faces, const double* ev, const double* em, const double* evol, const double* eB,
                    const double* ed, const double* eD, const double* fixedv, double mu, double lambda, double shear_stiffness,
                    double stiffness, double friction_coeff) {
    if (!c) return AEP_ERR_INVALID;
    cudaSetDevice(c->device);
    int r = mesh_upload(c->mesh, c->G, c->cfg.grid_min, c->h, nv, nf, vx, vv, vm, vvol, vB, faces, ev, em, evol, eB, ed, eD, fixedv, mu, lambda,
                        shear_stiffness, stiffness, friction_coeff, c->stream);
    if (r) return fail(c, r, "mesh upload failed: %s", cudaGetErrorString(cudaGetLastError()));
    c->inited = false;
    return AEP_OK;
}

int aep_set_levelset_analytic(aep_ctx* c, int kind, const double* P) {
    if (!c) return AEP_ERR_INVALID;
    if (kind == AEP_LS_NONE) { c->G.ls_code = nullptr; c->G.ls_nrm = nullptr; return AEP_OK; }
    if (kind < AEP_LS_GROUND || kind > AEP_LS_BOX || !P) return fail(c, AEP_ERR_INVALID, "unknown analytic level set %d", kind);
    const int nx = c->G.nx, ny = c->G.ny, nz = c->G.nz;
    std::vector<unsigned char> code(c->Ng, 0);
    std::vector<float4> nrm; const bool general = (kind == AEP_LS_SPHERE_GROUND);
    if (general) nrm.assign(c->Ng, make_float4(0.f, 0.f, 1.f, 0.f));
    const double* mn = c->cfg.grid_min; const double* h = c->h;
    parallel_for_planes(nz, [&](int k) {
        for (int j = 0; j < ny; ++j) for (int i = 0; i < nx; ++i) {
            const double gp[3] = { mn[0] + i * h[0], mn[1] + j * h[1], mn[2] + k * h[2] };       // HS:473-476
            if (ls_phi(kind, P, gp) <= 0.0) {                                                     // HS:478
                double n[3] = {0, 0, 1};
                const int cd = ls_normal_code(kind, P, gp, n);
                const size_t id = ((size_t)k * ny + j) * nx + i;
                code[id] = (unsigned char)cd;
                if (cd == 7) nrm[id] = make_float4((float)n[0], (float)n[1], (float)n[2], 0.f);
            }
        }
    });
    return upload_levelset(c, code, general ? &nrm : nullptr);
}

int aep_set_levelset_samples(aep_ctx* c, const uint8_t* inside, const double* normal) {
    if (!c || !inside || !normal) return fail(c, AEP_ERR_INVALID, "null level-set samples");
    std::vector<unsigned char> code(c->Ng, 0); std::vector<float4> nrm(c->Ng, make_float4(0.f, 0.f, 1.f, 0.f));
    const size_t Ng = c->Ng;
    for (size_t i = 0; i < Ng; ++i)
        if (inside[i]) { code[i] = 7; nrm[i] = make_float4((float)normal[i], (float)normal[Ng + i], (float)normal[2 * Ng + i], 0.f); }
    return upload_levelset(c, code, &nrm);
}

int aep_init_begin(aep_ctx* c) {
    if (!c) return AEP_ERR_INVALID;
    if (c->n == 0 && c->mesh.nv == 0 && c->cfg.slab_axis < 0) return fail(c, AEP_ERR_INVALID, "nothing to simulate: upload particles and/or a mesh first");
    cudaSetDevice(c->device);
    c->inited = true; c->pending_leave = 0;
    int r;
    if ((r = do_sort(c, true))) return r;
    return do_p2g(c, false);                                                // HS:854 (mass / momentum part)
}
int aep_init_volumes(aep_ctx* c) {
    int r = require_init(c); if (r) return r;
    if (c->n) { k_init_volumes<<<cdiv(c->n, 256), 256, 0, c->stream>>>(c->P[c->cur], c->G, (int)c->n); LAUNCH_OK("k_init_volumes"); }   // HS:242-249
    k_vmax_from_mp<<<c->nrun, 256, 0, c->stream>>>(c->G, c->d_clk); LAUNCH_OK("k_vmax_from_mp");
    return AEP_OK;
}
int aep_init_dt(aep_ctx* c) {
    int r = require_init(c); if (r) return r;
    k_initial_dt<<<1, 1, 0, c->stream>>>(c->d_clk); LAUNCH_OK("k_initial_dt");  // HS:860
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

int aep_run_frames(aep_ctx* c, int n_frames, int max_substeps, int64_t* substeps_done) {
    int r = require_init(c); if (r) return r;
    SimClock clk;
    CU(cudaMemcpyAsync(&clk, c->d_clk, sizeof clk, cudaMemcpyDeviceToHost, c->stream)); CU(cudaStreamSynchronize(c->stream));
    const int target = clk.frame_no + n_frames; int64_t done = 0;
    // a frame takes >= 1/(60 * dt_max) = 5 substeps (dt <= cfl/rate_floor); check the device clock every 4 substeps
    while (clk.frame_no < target && done < max_substeps) {
        const int burst = 1;
        for (int s = 0; s < burst; ++s) { if ((r = do_substep(c))) return r; ++done; }
        CU(cudaMemcpyAsync(&clk, c->d_clk, sizeof clk, cudaMemcpyDeviceToHost, c->stream)); CU(cudaStreamSynchronize(c->stream));
    }
    if (substeps_done) *substeps_done = done;
    return AEP_OK;
}

int aep_p2g(aep_ctx* c, int first) {
    int r = require_init(c); if (r) return r;
    if ((r = do_sort(c, true))) return r;
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
int aep_stage_forces(aep_ctx* c, double dt) { int r = require_init(c); if (r) return r; if ((r = aep_set_dt(c, dt))) return r; return do_forces(c); }
int aep_stage_grid(aep_ctx* c, double dt) {
    int r = require_init(c); if (r) return r; if ((r = aep_set_dt(c, dt))) return r;
    if ((r = do_grid(c))) return r;
    k_advance_clock<<<1, 1, 0, c->stream>>>(c->d_clk, 2, (int)c->n); LAUNCH_OK("k_advance_clock");   // latch vmax, keep dt
    return AEP_OK;
}
int aep_stage_g2p(aep_ctx* c, double dt) { int r = require_init(c); if (r) return r; if ((r = aep_set_dt(c, dt))) return r; return do_g2p(c); }

int aep_get_clock(aep_ctx* c, double* dt, double* t, double* inner_t, int32_t* frame_no, int64_t* substeps, double* vmax, int64_t* escaped) {
    if (!c) return AEP_ERR_INVALID;
    cudaSetDevice(c->device);
    SimClock clk;
    CU(cudaMemcpyAsync(&clk, c->d_clk, sizeof clk, cudaMemcpyDeviceToHost, c->stream)); CU(cudaStreamSynchronize(c->stream));
    if (dt) *dt = clk.dt; if (t) *t = clk.t; if (inner_t) *inner_t = clk.inner_t; if (frame_no) *frame_no = clk.frame_no;
    if (substeps) *substeps = clk.substeps; if (vmax) *vmax = clk.vmax_last; if (escaped) *escaped = (int64_t)clk.escaped;
    return AEP_OK;
}

int64_t aep_num_particles(aep_ctx* c) { return c ? c->n - c->pending_leave : -1; }      // live particles (dead slots of a slab context excluded)

int aep_download_particles(aep_ctx* c, double* x, double* v, double* B1, double* B2, double* B3, double* FE, double* FP, double* vol, double* q) {
    if (!c) return AEP_ERR_INVALID;
    cudaSetDevice(c->device);
    int r = compact_slab(c); if (r) return r;
    const long long n = c->n; if (n == 0) return AEP_OK;
    const long long CH = 1 << 22;
    r = ensure_stage(c, (size_t)std::min(n, CH) * 36 * sizeof(double)); if (r) return r;
    for (long long p0 = 0; p0 < n; p0 += CH) {
        const long long cnt = std::min(CH, n - p0);
        double* st = c->d_stage;
        k_download_convert<<<cdiv(n, 256), 256, 0, c->stream>>>(c->P[c->cur], c->G, st, (int)n, p0, (int)cnt, c->cfg.grid_min[0], c->cfg.grid_min[1],
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

int aep_download_positions_f32(aep_ctx* c, float* xyz) {
    if (!c || !xyz) return AEP_ERR_INVALID;
    cudaSetDevice(c->device);
    int r = compact_slab(c); if (r) return r;
    const long long n = c->n; if (n == 0) return AEP_OK;
    r = ensure_stage(c, (size_t)n * 3 * sizeof(float)); if (r) return r;
    k_download_positions_f32<<<cdiv(n, 256), 256, 0, c->stream>>>(c->P[c->cur], c->G, (float*)c->d_stage, (int)n, c->cfg.slab_axis >= 0 ? 1 : 0);
    LAUNCH_OK("k_download_positions_f32");
    CU(cudaMemcpyAsync(xyz, c->d_stage, (size_t)n * 3 * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
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
        k_download_grid<<<cdiv(cnt, 256), 256, 0, c->stream>>>(c->G, sm, sv, sf, svt, n0, cnt, Ng, c->grid_mode);
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
    double h[6] = {0, 0, 0, 0, 0, 0};
    if (c->n) {
        CU(cudaMemsetAsync(c->d_stats, 0, 6 * sizeof(double), c->stream));
        k_stats<<<std::min(cdiv(c->n, 256), 148 * 8), 256, 0, c->stream>>>(c->P[c->cur], c->G, c->d_stats, (int)c->n);
        LAUNCH_OK("k_stats");
        CU(cudaMemcpyAsync(h, c->d_stats, sizeof h, cudaMemcpyDeviceToHost, c->stream)); CU(cudaStreamSynchronize(c->stream));
    }
    if (com3) for (int a = 0; a < 3; ++a) com3[a] = h[5] > 0 ? h[a] / h[5] : 0.0;
    if (kinetic) *kinetic = h[3];
    const long long live = c->n - c->pending_leave;
    if (mean_jp) *mean_jp = live ? h[4] / (double)live : 0.0;
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

// ---- split stepping for the multi-GPU driver
int aep_step_forces(aep_ctx* c) { int r = require_init(c); if (r) return r; return do_forces(c); }
int aep_step_grid(aep_ctx* c) { int r = require_init(c); if (r) return r; return do_grid(c); }
int aep_step_g2p(aep_ctx* c) {
    int r = require_init(c); if (r) return r;
    if ((r = do_clock(c))) return r;
    return do_g2p(c);
}
int aep_step_p2g(aep_ctx* c) {
    int r = require_init(c); if (r) return r;
    if ((r = maybe_sort(c, true))) return r;
    return do_p2g(c, false);
}
// halo exchange / migration entry points
#include "aep_halo.inl"

}  // extern "C"
