// aep_comm.inl -- peer-memory communication of slab contexts: host side (included by aep_engine.cu; kernels in aep_halo.cuh).
//
// Setup, once:   aep_comm_export on every rank -> the caller gathers the 256-byte blobs (torch.distributed all_gather, or simply an
// array inside one process) -> aep_comm_connect with all of them.  From then on aep_init / aep_substep / aep_run / aep_run_frames of a
// slab context include the exchanges: halo planes, migrating particles, max|v| and the cloth's advected points are STORED INTO THE
// NEIGHBOUR'S MEMORY by the kernels of the substep (NVLink peer stores; CUDA IPC mappings between processes), completion travels
// as an epoch flag, and the receiving stream waits for it with a one-thread kernel.  Nothing returns to the host.
//
// Why a receive buffer is never overwritten before it was consumed (phases of enqueue_phase in aep_engine.cu; A, B neighbours):
//   f planes:   A sends f of substep s+1 in its phase 0.  That follows A's phase 6 of substep s (stream order), which waited for B's
//               (m,p) flag of substep s, raised in B's phase 5 -- after B's phase 1 of substep s, where B consumed A's f of substep s.
//   (m,p):      A sends (m,p) of substep s in phase 5, after its phase 1 of s, which waited for B's f flag of s (B's phase 0 of s),
//               raised after B's phase 6 of s-1, where B consumed A's (m,p) of s-1.
//   migrants:   A's phase 4 of s follows its phase 1 of s -> B's phase 0 of s -> B's phase 5 of s-1, where B appended A's migrants of s-1.
//   max|v|:     slots alternate by epoch parity; A rewrites a slot in phase 1 of s+2, after its phase 6 of s+1 -> B's phase 5 of s+1,
//               after B's phase 2 of s, where B read it.
// The sender fences every thread's peer stores (__threadfence_system) before its CTA counts itself done; the last CTA fences again and
// raises the flag with st.release.sys; the waiting thread reads it with ld.acquire.sys and the consumer is a LATER kernel of the stream.
namespace {

struct CommBlob {                       // what a rank tells the others (AEP_COMM_BLOB_BYTES)
    unsigned int magic; int pid; int device; int has_mesh;
    unsigned long long comm_ptr, comm_bytes, mesh_ptr, mesh_bytes;
    long long plane_nodes; int mig_cap; int slab_axis, slab_lo, slab_hi, res_axis;
    cudaIpcMemHandle_t comm_handle, mesh_handle;
    char pad[256 - 4 * 4 - 4 * 8 - 8 - 6 * 4 - 2 * 64];
};
static_assert(sizeof(CommBlob) == AEP_COMM_BLOB_BYTES, "CommBlob is the 256-byte blob of the ABI");

// planes [p0, p0 + np) of the slab axis as a contiguous range of a grid array (slab layout: the slab axis is the slowest)
inline float4* plane_ptr(aep_ctx* c, float4* arr, int p0) {
    const int a = c->cfg.slab_axis;
    return arr + (a == 1 ? nidx(c->G, 0, p0, 0) : nidx(c->G, 0, 0, p0));
}

struct HaloGeom { int np; int send0[2], recv0[2]; long long n_f4; int ctas; };
// (m,p): 4 planes per side -- what I scattered into the neighbour's reach, including the scratch plane that catches particles about
// to migrate; f: the 3 planes both sides share.  Send range / receive range per side (0 = low neighbour, 1 = high).
inline HaloGeom halo_geom(aep_ctx* c, int what) {
    HaloGeom g; const int lo = c->cfg.slab_lo, hi = c->cfg.slab_hi;
    g.np = what == 0 ? 4 : 3;
    g.send0[0] = what == 0 ? lo - 2 : lo - 1; g.send0[1] = hi - 1;
    g.recv0[0] = lo - 1; g.recv0[1] = what == 0 ? hi - 2 : hi - 1;
    g.n_f4 = (long long)g.np * c->comm.plane_nodes;
    g.ctas = (int)std::min<long long>((g.n_f4 + 255) / 256, c->sm_count * 4);
    return g;
}
int peer_halo_send(aep_ctx* c, int what, int halt_class) {
    StageTimer T(c, AEP_STAGE_HALO);
    Comm& m = c->comm;
    const int a = c->cfg.slab_axis, nres = c->cfg.res[a];
    float4* arr = what == 0 ? c->G.mp : c->G.f;
    const HaloGeom g = halo_geom(c, what);
    for (int side = 0; side < 2; ++side) {
        if (!m.peers.halo_in[what][side]) continue;
        if (g.send0[side] < 0 || g.send0[side] + g.np > nres) return fail(c, AEP_ERR_INVALID, "slab too close to the domain face for the halo planes");
        const int nb = side == 0 ? m.rank - 1 : m.rank + 1;
        k_peer_halo_send<<<g.ctas, 256, 0, c->stream>>>(plane_ptr(c, arr, g.send0[side]), g.n_f4, m.peers.halo_in[what][side],
                                                      &m.peers.head[nb]->halo_flag[what][1 - side], &m.d_local->halo_epoch[what][side],
                                                      &m.d_local->done[what * 2 + side], c->d_clk, halt_class);
        LAUNCH_OK("k_peer_halo_send");
    }
    return AEP_OK;
}
int peer_halo_recv(aep_ctx* c, int what, int halt_class) {
    StageTimer T(c, AEP_STAGE_HALO);
    Comm& m = c->comm;
    const int a = c->cfg.slab_axis;
    float4* arr = what == 0 ? c->G.mp : c->G.f;
    const HaloGeom g = halo_geom(c, what);
    CommHead* me = m.peers.head[m.rank];
    const bool has0 = m.peers.halo_in[what][0] != nullptr, has1 = m.peers.halo_in[what][1] != nullptr;
    if (has0 || has1) {
        k_peer_wait2<<<1, 1, 0, c->stream>>>(has0 ? &me->halo_flag[what][0] : nullptr, &m.d_local->halo_epoch[what][0],
                                             has1 ? &me->halo_flag[what][1] : nullptr, &m.d_local->halo_epoch[what][1], c->d_clk, halt_class);
        LAUNCH_OK("k_peer_wait2");
    }
    for (int side = 0; side < 2; ++side) {
        if (!m.peers.halo_in[what][side]) continue;
        float4* buf = reinterpret_cast<float4*>(m.block + m.halo_off[what][side]);
        k_peer_halo_add<<<g.ctas, 256, 0, c->stream>>>(plane_ptr(c, arr, g.recv0[side]), buf, g.n_f4, c->G, a, g.recv0[side], m.plane_nodes, c->d_clk, halt_class);
        LAUNCH_OK("k_peer_halo_add");
    }
    return AEP_OK;
}

int peer_vmax_share(aep_ctx* c, int halt_class) {
    Comm& m = c->comm;
    k_peer_vmax_share<<<1, 32, 0, c->stream>>>(m.peers, m.d_local, c->d_clk, halt_class); LAUNCH_OK("k_peer_vmax_share");
    return AEP_OK;
}
int peer_vmax_reduce(aep_ctx* c, int halt_class) {
    Comm& m = c->comm;
    k_peer_vmax_reduce<<<1, 1, 0, c->stream>>>(m.peers, m.d_local, c->d_clk, halt_class); LAUNCH_OK("k_peer_vmax_reduce");
    return AEP_OK;
}

int peer_migrate_send(aep_ctx* c) {
    StageTimer T(c, AEP_STAGE_HALO);
    Comm& m = c->comm;
    k_peer_migrate_send<<<2 * cdiv(c->mig.cap, 256), 256, 0, c->stream>>>(c->P[c->cur], c->mig, m.peers, m.d_local, c->d_clk, 2);
    LAUNCH_OK("k_peer_migrate_send");
    return AEP_OK;
}
int peer_migrate_recv(aep_ctx* c) {
    StageTimer T(c, AEP_STAGE_HALO);
    Comm& m = c->comm;
    CommHead* me = m.peers.head[m.rank];
    const bool has0 = m.rank > 0, has1 = m.rank < m.world - 1;
    if (has0 || has1) {
        k_peer_wait2<<<1, 1, 0, c->stream>>>(has0 ? &me->mig_flag[0] : nullptr, &m.d_local->mig_epoch, has1 ? &me->mig_flag[1] : nullptr, &m.d_local->mig_epoch, c->d_clk, 2);
        LAUNCH_OK("k_peer_wait2");
        k_peer_migrate_insert<<<32, 256, 0, c->stream>>>(c->P[c->cur], reinterpret_cast<const float4*>(m.block + m.mig_off[0]),
                                                         reinterpret_cast<const float4*>(m.block + m.mig_off[1]), me, (int)c->cap, c->d_clk, &m.d_local->done[7], 2);
        LAUNCH_OK("k_peer_migrate_insert");
    }
    return AEP_OK;
}

// cloth: what this rank advanced (0: vertices, 1: elements) -> every rank's copy ... then wait for everybody else's
int peer_mesh_push(aep_ctx* c, int which) {
    Comm& m = c->comm;
    if (m.world < 2 || !c->mesh.nv) return AEP_OK;
    const MeshState& ms = c->mesh;
    const int n = which == 0 ? (int)ms.nv : (int)ms.nf, narr = which == 0 ? 4 : 7;
    const size_t off = which == 0 ? ms.sync_v_off : ms.sync_e_off;
    dim3 grid((unsigned)std::min(cdiv(n, 256), c->sm_count * 2), (unsigned)m.world);
    k_peer_mesh_push<<<grid, 256, 0, c->stream>>>(m.peers, m.d_local, which == 0 ? ms.owner_v : ms.owner_e, off, narr, n, which, c->d_clk, 2);
    LAUNCH_OK("k_peer_mesh_push");
    return AEP_OK;
}
int peer_mesh_wait(aep_ctx* c, int which) {
    Comm& m = c->comm;
    if (m.world < 2 || !c->mesh.nv) return AEP_OK;
    k_peer_mesh_wait<<<1, 1, 0, c->stream>>>(m.peers, m.d_local, which, c->d_clk, 2);
    LAUNCH_OK("k_peer_mesh_wait");
    return AEP_OK;
}

// inside one process the members' streams are ordered by events (see Comm::group): a phase waits for the phase before it of every member
int phase_begin(aep_ctx* c, int phase) {
    Comm& m = c->comm;
    if (m.group.empty() || phase == 0) return AEP_OK;
    for (aep_ctx* g : m.group) if (g != c) CU(cudaStreamWaitEvent(c->stream, g->comm.ev_phase[phase - 1], 0));
    return AEP_OK;
}
int phase_end(aep_ctx* c, int phase) {
    Comm& m = c->comm;
    if (m.group.empty()) return AEP_OK;
    CU(cudaEventRecord(m.ev_phase[phase], c->stream));
    return AEP_OK;
}

}  // namespace

extern "C" {

int aep_comm_export(aep_ctx* c, void* blob256, int64_t migrate_capacity) {
    if (!c || !blob256) return AEP_ERR_INVALID;
    const int a = c->cfg.slab_axis;
    if (a != 1 && a != 2) return fail(c, AEP_ERR_INVALID, "peer communication needs a y- or z-slab context (slab_axis 1 or 2)");
    if (c->cap <= 0) return fail(c, AEP_ERR_INVALID, "set aep_config.particle_capacity (room for arriving particles) before aep_comm_export");
    cudaSetDevice(c->device);
    Comm& m = c->comm;
    if (!m.exported) {
        m.plane_nodes = a == 1 ? (long long)c->G.nx * c->G.nz : (long long)c->G.nx * c->G.ny;
        m.mig_cap = (int)std::min<int64_t>(std::max<int64_t>(migrate_capacity, 1024), 1 << 28);
        size_t off = (sizeof(CommHead) + 255) & ~(size_t)255;
        for (int what = 0; what < 2; ++what) for (int side = 0; side < 2; ++side) { m.halo_off[what][side] = off; off += (size_t)(what == 0 ? 4 : 3) * m.plane_nodes * sizeof(float4); }
        for (int side = 0; side < 2; ++side) { m.mig_off[side] = off; off += (size_t)m.mig_cap * P_NARR * sizeof(float4); }
        m.bytes = off;
        CU(cudaMalloc((void**)&m.block, m.bytes));
        CU(cudaMemsetAsync(m.block, 0, m.bytes, c->stream));
        CU(cudaMalloc((void**)&m.d_local, sizeof(CommLocal)));
        CU(cudaMemsetAsync(m.d_local, 0, sizeof(CommLocal), c->stream));
        // leaver lists + counts for the fused kernel
        if (c->mig.axis < 0) {
            for (int s = 0; s < 2; ++s) CU(dalloc(c, &c->mig.list[s], (size_t)m.mig_cap));
            CU(dalloc(c, &c->mig.counts, 2));
            CU(cudaMemsetAsync(c->mig.counts, 0, 2 * sizeof(unsigned long long), c->stream));
            c->mig.cap = m.mig_cap; c->mig.lo = c->cfg.slab_lo; c->mig.hi = c->cfg.slab_hi; c->mig.axis = a;
        }
        CU(cudaStreamSynchronize(c->stream));
        m.exported = true;
    }
    CommBlob b; std::memset(&b, 0, sizeof b);
    b.magic = 0xAE9C0331u; b.pid = (int)getpid(); b.device = c->device; b.has_mesh = c->mesh.block ? 1 : 0;
    b.comm_ptr = (unsigned long long)(uintptr_t)m.block; b.comm_bytes = m.bytes;
    b.mesh_ptr = (unsigned long long)(uintptr_t)c->mesh.block; b.mesh_bytes = c->mesh.block_bytes;
    b.plane_nodes = m.plane_nodes; b.mig_cap = m.mig_cap; b.slab_axis = a; b.slab_lo = c->cfg.slab_lo; b.slab_hi = c->cfg.slab_hi; b.res_axis = c->cfg.res[a];
    CU(cudaIpcGetMemHandle(&b.comm_handle, m.block));
    if (c->mesh.block) CU(cudaIpcGetMemHandle(&b.mesh_handle, c->mesh.block));
    std::memcpy(blob256, &b, sizeof b);
    return AEP_OK;
}

int aep_comm_connect(aep_ctx* c, int rank, int world, const void* blobs) {
    if (!c || !blobs) return AEP_ERR_INVALID;
    Comm& m = c->comm;
    if (!m.exported) return fail(c, AEP_ERR_INVALID, "aep_comm_export first");
    if (world < 1 || world > AEP_MAX_WORLD || rank < 0 || rank >= world) return fail(c, AEP_ERR_INVALID, "bad rank / world (at most %d ranks)", AEP_MAX_WORLD);
    cudaSetDevice(c->device);
    const CommBlob* B = reinterpret_cast<const CommBlob*>(blobs);
    const int mypid = (int)getpid();
    m.rank = rank; m.world = world;
    CommPeers& P = m.peers; std::memset(&P, 0, sizeof P);
    P.rank = rank; P.world = world;
    for (int r = 0; r < world; ++r) {
        const CommBlob& b = B[r];
        if (b.magic != 0xAE9C0331u) return fail(c, AEP_ERR_INVALID, "blob of rank %d is not an aep_comm_export blob", r);
        if (b.slab_axis != c->cfg.slab_axis || b.plane_nodes != m.plane_nodes || b.mig_cap != m.mig_cap || b.has_mesh != (c->mesh.block ? 1 : 0) || b.mesh_bytes != c->mesh.block_bytes)
            return fail(c, AEP_ERR_INVALID, "rank %d was set up differently (slab axis, grid, migration capacity or mesh)", r);
        if (r > 0 && B[r - 1].slab_hi != b.slab_lo) return fail(c, AEP_ERR_INVALID, "slabs of ranks %d and %d do not meet", r - 1, r);
        if (world > 1 && b.slab_hi - b.slab_lo < 4) return fail(c, AEP_ERR_INVALID, "slab of rank %d is %d cells wide; need >= 4", r, b.slab_hi - b.slab_lo);
        void* comm = nullptr; void* mesh = nullptr;
        if (r == rank) { comm = m.block; mesh = c->mesh.block; }
        else if (b.pid == mypid) {                      // same process: plain pointers (+ peer access between devices)
            comm = (void*)(uintptr_t)b.comm_ptr; mesh = (void*)(uintptr_t)b.mesh_ptr;
            if (b.device != c->device) { cudaError_t e = cudaDeviceEnablePeerAccess(b.device, 0); if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(c, AEP_ERR_CUDA, "no peer access to device %d: %s", b.device, cudaGetErrorString(e)); cudaGetLastError(); }
        } else {
            CU(cudaIpcOpenMemHandle(&comm, b.comm_handle, cudaIpcMemLazyEnablePeerAccess)); m.opened.push_back(comm);
            if (b.has_mesh) { CU(cudaIpcOpenMemHandle(&mesh, b.mesh_handle, cudaIpcMemLazyEnablePeerAccess)); m.opened.push_back(mesh); }
        }
        P.head[r] = reinterpret_cast<CommHead*>(comm);
        P.mesh[r] = reinterpret_cast<unsigned char*>(mesh);
    }
    // my low side is the lower neighbour's high side: I fill ITS receive buffers of side 1 (offsets are the same on every rank)
    for (int side = 0; side < 2; ++side) {
        const int nb = side == 0 ? rank - 1 : rank + 1;
        if (nb < 0 || nb >= world) continue;
        unsigned char* base = reinterpret_cast<unsigned char*>(P.head[nb]);
        for (int what = 0; what < 2; ++what) P.halo_in[what][side] = reinterpret_cast<float4*>(base + m.halo_off[what][1 - side]);
        P.mig_in[side] = reinterpret_cast<float4*>(base + m.mig_off[1 - side]);
    }
    m.connected = world > 1;
    c->graph_dirty = true;
    return AEP_OK;
}

// several slab contexts of ONE process (one GPU or several): connect them by plain pointers
int aep_comm_connect_local(aep_ctx** ctxs, int world, int64_t migrate_capacity) {
    if (!ctxs || world < 1 || world > AEP_MAX_WORLD) return AEP_ERR_INVALID;
    std::vector<CommBlob> blobs((size_t)world);
    for (int r = 0; r < world; ++r) { int rc = aep_comm_export(ctxs[r], &blobs[r], migrate_capacity); if (rc) return rc; }
    for (int r = 0; r < world; ++r) { int rc = aep_comm_connect(ctxs[r], r, world, blobs.data()); if (rc) return rc; }
    for (int r = 0; r < world; ++r) {
        aep_ctx* c = ctxs[r]; cudaSetDevice(c->device);
        c->comm.group.assign(ctxs, ctxs + world);
        for (int i = 0; i < 8; ++i) if (!c->comm.ev_phase[i]) CU(cudaEventCreateWithFlags(&c->comm.ev_phase[i], cudaEventDisableTiming));
    }
    return AEP_OK;
}
// ... and step them in lockstep from one host thread: a context's stream waits (on the device) for flags its neighbours raise in the
// same substep, so the host must have queued every context's substep before it blocks on any of them
int aep_group_init(aep_ctx** ctxs, int world) {
    int rc;
    for (int r = 0; r < world; ++r) if ((rc = aep_init_begin(ctxs[r]))) return rc;
    for (int r = 0; r < world; ++r) if ((rc = aep_init_volumes(ctxs[r]))) return rc;
    for (int r = 0; r < world; ++r) if ((rc = aep_init_dt_async(ctxs[r]))) return rc;
    for (int r = 0; r < world; ++r) if ((rc = aep_sync(ctxs[r]))) return rc;
    return AEP_OK;
}
int aep_group_run(aep_ctx** ctxs, int world, int n_substeps) {
    int rc;
    for (int r = 0; r < world; ++r) if ((rc = require_init(ctxs[r]))) return rc;
    for (int s = 0; s < n_substeps; ++s) {
        // phase by phase over the members (a phase waits on events the members record at the end of the phase before), then the
        // (possibly host-synchronising) re-sort decisions
        for (int ph = 0; ph < SUBSTEP_PHASES; ++ph)
            for (int r = 0; r < world; ++r) { cudaSetDevice(ctxs[r]->device); if ((rc = enqueue_phase(ctxs[r], ph))) return rc; }
        for (int r = 0; r < world; ++r) { cudaSetDevice(ctxs[r]->device); if ((rc = maybe_sort(ctxs[r]))) return rc; }
    }
    return AEP_OK;
}

}  // extern "C"
