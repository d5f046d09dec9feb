// aep_halo.inl -- C ABI of the caller-driven slab exchange (included inside extern "C" of aep_engine.cu): the caller owns the
// communication buffers and moves the bytes.  The peer-memory exchange that runs inside aep_substep is in aep_comm.inl.
static int slab_check(aep_ctx* c, int side, bool* has) {
    if (!c) return AEP_ERR_INVALID;
    if (c->cfg.slab_axis < 0 || c->cfg.slab_axis > 2) return fail(c, AEP_ERR_INVALID, "context has no slab (cfg.slab_axis < 0)");
    if (side != 0 && side != 1) return fail(c, AEP_ERR_INVALID, "side must be 0 (low) or 1 (high)");
    const int nres = c->cfg.res[c->cfg.slab_axis];
    *has = side == 0 ? (c->cfg.slab_lo > 0) : (c->cfg.slab_hi < nres);
    cudaSetDevice(c->device);
    return AEP_OK;
}
static PlaneMap plane_map(aep_ctx* c, int side) {
    PlaneMap M; M.axis = c->cfg.slab_axis; M.nplanes = 3;                 // un-fused path: P2G runs after the migration, 3 shared planes
    M.plane0 = (side == 0 ? c->cfg.slab_lo : c->cfg.slab_hi) - 1;
    if (M.axis == 0) { M.nu = c->G.ny; M.nv = c->G.nz; } else if (M.axis == 1) { M.nu = c->G.nx; M.nv = c->G.nz; } else { M.nu = c->G.nx; M.nv = c->G.ny; }
    return M;
}

int aep_halo_info(aep_ctx* c, int what, int side, int64_t* n_floats) {
    bool has; int r = slab_check(c, side, &has); if (r) return r;
    if (what != 0 && what != 1) return fail(c, AEP_ERR_INVALID, "what must be 0 (m,p) or 1 (f)");
    const PlaneMap M = plane_map(c, side);
    if (n_floats) *n_floats = has ? (int64_t)M.nplanes * M.nu * M.nv * 4 : 0;
    return AEP_OK;
}

int aep_halo_pack(aep_ctx* c, int what, int side, void* dev_send) {
    bool has; int r = slab_check(c, side, &has); if (r) return r;
    if (!has) return AEP_OK;
    if (!dev_send) return fail(c, AEP_ERR_INVALID, "null halo buffer");
    StageTimer T(c, AEP_STAGE_HALO);
    const PlaneMap M = plane_map(c, side);
    const long long n = (long long)M.nplanes * M.nu * M.nv;
    k_halo_pack<<<cdiv(n, 256), 256, 0, c->stream>>>(what == 0 ? c->G.mp : c->G.f, (float4*)dev_send, c->G, M);
    LAUNCH_OK("k_halo_pack");
    return AEP_OK;
}

int aep_halo_add(aep_ctx* c, int what, int side, const void* dev_recv) {
    bool has; int r = slab_check(c, side, &has); if (r) return r;
    if (!has) return AEP_OK;
    if (!dev_recv) return fail(c, AEP_ERR_INVALID, "null halo buffer");
    StageTimer T(c, AEP_STAGE_HALO);
    const PlaneMap M = plane_map(c, side);
    const long long n = (long long)M.nplanes * M.nu * M.nv;
    k_halo_add<<<cdiv(n, 256), 256, 0, c->stream>>>(what == 0 ? c->G.mp : c->G.f, (const float4*)dev_recv, c->G, M);
    LAUNCH_OK("k_halo_add");
    return AEP_OK;
}

int aep_vmax_get(aep_ctx* c, void* dev_float) {
    if (!c || !dev_float) return AEP_ERR_INVALID;
    cudaSetDevice(c->device);
    CU(cudaMemcpyAsync(dev_float, &c->d_clk->vmax_bits, 4, cudaMemcpyDeviceToDevice, c->stream));
    return AEP_OK;
}
int aep_vmax_set(aep_ctx* c, const void* dev_float) {
    if (!c || !dev_float) return AEP_ERR_INVALID;
    cudaSetDevice(c->device);
    CU(cudaMemcpyAsync(&c->d_clk->vmax_bits, dev_float, 4, cudaMemcpyDeviceToDevice, c->stream));
    return AEP_OK;
}

int aep_migrate_extract(aep_ctx* c, void* dev_to_low, void* dev_to_high, int64_t capacity, int64_t* n_low, int64_t* n_high) {
    int r = require_init(c); if (r) return r;
    if (c->cfg.slab_axis < 0) return fail(c, AEP_ERR_INVALID, "context has no slab");
    if (n_low) *n_low = 0;
    if (n_high) *n_high = 0;
    if (c->n == 0) return AEP_OK;
    if (!dev_to_low || !dev_to_high || capacity < 0) return fail(c, AEP_ERR_INVALID, "null migration buffer");
    StageTimer T(c, AEP_STAGE_HALO);
    unsigned long long* cnt = (unsigned long long*)c->d_stats;
    CU(cudaMemsetAsync(cnt, 0, 2 * sizeof(unsigned long long), c->stream));
    k_migrate_extract<<<cdiv(c->n, 256), 256, 0, c->stream>>>(c->P[c->cur], (int)c->n, c->cfg.slab_axis, c->cfg.slab_lo, c->cfg.slab_hi,
                                                              (float4*)dev_to_low, (float4*)dev_to_high, capacity, cnt);
    LAUNCH_OK("k_migrate_extract");
    unsigned long long h[2];
    CU(cudaMemcpyAsync(h, cnt, sizeof h, cudaMemcpyDeviceToHost, c->stream)); CU(cudaStreamSynchronize(c->stream));
    if ((int64_t)h[0] > capacity || (int64_t)h[1] > capacity)
        return fail(c, AEP_ERR_STATE, "migration buffer too small: %llu / %llu particles leave, capacity %lld", h[0], h[1], (long long)capacity);
    if (n_low) *n_low = (int64_t)h[0]; if (n_high) *n_high = (int64_t)h[1];
    c->pending_leave += (long long)(h[0] + h[1]);                           // dead slots until the next physical sort
    return AEP_OK;
}

/* sync-free variant: k_g2p lists the leavers, extract_begin gathers them (no scan over all particles, no host round trip);
 * the caller moves the counts (device int64[2]) wherever it wants and reports them back with extract_end. */
int aep_migrate_bind(aep_ctx* c, void* dev_to_low, void* dev_to_high, int64_t capacity, void* dev_counts) {
    if (!c) return AEP_ERR_INVALID;
    if (c->cfg.slab_axis < 0) return fail(c, AEP_ERR_INVALID, "context has no slab");
    if (!dev_to_low || !dev_to_high || !dev_counts || capacity <= 0 || capacity > (1 << 28)) return fail(c, AEP_ERR_INVALID, "bad migration buffers");
    cudaSetDevice(c->device);
    if (c->mig.axis >= 0) return fail(c, AEP_ERR_INVALID, "migration buffers are already bound");
    if (c->comm.exported) return fail(c, AEP_ERR_INVALID, "this context uses the peer-memory exchange (aep_comm_export)");
    for (int s = 0; s < 2; ++s) CU(dalloc(c, &c->mig.list[s], (size_t)capacity));
    c->mig.cap = (int)capacity; c->mig.lo = c->cfg.slab_lo; c->mig.hi = c->cfg.slab_hi;
    c->mig.counts = (unsigned long long*)dev_counts;
    c->mig_buf[0] = (float4*)dev_to_low; c->mig_buf[1] = (float4*)dev_to_high;
    CU(cudaMemsetAsync(dev_counts, 0, 2 * sizeof(unsigned long long), c->stream));
    c->mig.axis = c->cfg.slab_axis;
    return AEP_OK;
}
int aep_migrate_extract_begin(aep_ctx* c) {
    int r = require_init(c); if (r) return r;
    if (c->mig.axis < 0) return fail(c, AEP_ERR_INVALID, "aep_migrate_bind has not been called");
    StageTimer T(c, AEP_STAGE_HALO);
    k_migrate_gather<<<2 * cdiv(c->mig.cap, 256), 256, 0, c->stream>>>(c->P[c->cur], c->mig, c->mig_buf[0], c->mig_buf[1]);
    LAUNCH_OK("k_migrate_gather");
    return AEP_OK;
}
int aep_migrate_extract_end(aep_ctx* c, int64_t n_low, int64_t n_high) {
    int r = require_init(c); if (r) return r;
    if (n_low < 0 || n_high < 0) return fail(c, AEP_ERR_INVALID, "negative count");
    if (n_low > c->mig.cap || n_high > c->mig.cap)
        return fail(c, AEP_ERR_STATE, "migration buffer too small: %lld / %lld particles leave, capacity %d", (long long)n_low, (long long)n_high, c->mig.cap);
    c->pending_leave += n_low + n_high;
    return AEP_OK;
}
/* P2G of the last `count` slots only (particles appended by aep_migrate_insert after aep_step_p2g already ran): P2G is additive */
int aep_step_p2g_arrivals(aep_ctx* c, int64_t count) {
    int r = require_init(c); if (r) return r;
    if (count < 0 || count > c->n) return fail(c, AEP_ERR_INVALID, "bad arrival count");
    if (count == 0) return AEP_OK;
    StageTimer T(c, AEP_STAGE_P2G);
    PartP tail = c->P[c->cur];
    for (int a = 0; a < P_NARR; ++a) tail.a[a] += (c->n - count);
    p2g_launch(c->stream, tail, c->G, count);
    LAUNCH_OK("k_p2g");
    return AEP_OK;
}

int aep_migrate_insert(aep_ctx* c, const void* dev_from_low, int64_t n_from_low, const void* dev_from_high, int64_t n_from_high) {
    int r = require_init(c); if (r) return r;
    if (n_from_low < 0 || n_from_high < 0) return fail(c, AEP_ERR_INVALID, "negative count");
    if (c->n + n_from_low + n_from_high > c->cap)
        return fail(c, AEP_ERR_ALLOC, "particle capacity %lld exceeded by migration (%lld + %lld + %lld); raise aep_config.particle_capacity",
                    c->cap, c->n, (long long)n_from_low, (long long)n_from_high);
    StageTimer T(c, AEP_STAGE_HALO);
    if (n_from_low) {
        k_migrate_insert<<<cdiv(n_from_low, 256), 256, 0, c->stream>>>(c->P[c->cur], (int)c->n, (const float4*)dev_from_low, (int)n_from_low, c->d_clk);
        LAUNCH_OK("k_migrate_insert"); c->n += n_from_low;
    }
    if (n_from_high) {
        k_migrate_insert<<<cdiv(n_from_high, 256), 256, 0, c->stream>>>(c->P[c->cur], (int)c->n, (const float4*)dev_from_high, (int)n_from_high, c->d_clk);
        LAUNCH_OK("k_migrate_insert"); c->n += n_from_high;
    }
    return AEP_OK;
}

int aep_set_particle_id_base(aep_ctx* c, int64_t id_base) {
    if (!c || id_base < 0) return AEP_ERR_INVALID;
    if (id_base >= (1ll << 31)) return fail(c, AEP_ERR_INVALID, "particle ids are 32-bit on the device: id_base must stay below 2^31");
    c->id_base = id_base;
    return AEP_OK;
}

int aep_download_particles_local(aep_ctx* c, int64_t* ids, double* x, double* v, double* B1, double* B2, double* B3, double* FE, double* FP,
                                 double* vol, double* q) {
    if (!c) return AEP_ERR_INVALID;
    cudaSetDevice(c->device);
    int r = compact_slab(c); if (r) return r;                             // drop dead slots so that exactly aep_num_particles entries come back
    const long long n = c->n; if (n == 0) return AEP_OK;
    const long long CH = 1 << 22;
    r = ensure_stage(c, (size_t)std::min(n, CH) * 37 * sizeof(double)); if (r) return r;
    for (long long p0 = 0; p0 < n; p0 += CH) {
        const long long cnt = std::min(CH, n - p0);
        double* st = c->d_stage; long long* dids = (long long*)(st + (size_t)35 * cnt);
        k_download_local<<<cdiv(cnt, 256), 256, 0, c->stream>>>(c->P[c->cur], st, dids, (int)p0, (int)cnt, c->cfg.grid_min[0], c->cfg.grid_min[1],
                                                               c->cfg.grid_min[2], c->h[0], c->h[1], c->h[2]);
        LAUNCH_OK("k_download_local");
        double* mats[5] = { x, v, B1, B2, B3 };
        for (int k = 0; k < 5; ++k) if (mats[k])
            for (int a = 0; a < 3; ++a)
                CU(cudaMemcpyAsync(mats[k] + (size_t)a * n + p0, st + (size_t)(3 * k + a) * cnt, cnt * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        if (vol) CU(cudaMemcpyAsync(vol + p0, st + (size_t)15 * cnt, cnt * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        if (q) CU(cudaMemcpyAsync(q + p0, st + (size_t)16 * cnt, cnt * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        if (FE) CU(cudaMemcpyAsync(FE + (size_t)9 * p0, st + (size_t)17 * cnt, 9 * cnt * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        if (FP) CU(cudaMemcpyAsync(FP + (size_t)9 * p0, st + (size_t)26 * cnt, 9 * cnt * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        if (ids) CU(cudaMemcpyAsync(ids + p0, dids, cnt * sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
    }
    return AEP_OK;
}
