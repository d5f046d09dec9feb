// aep_halo.inl -- slab halo exchange + particle migration entry points.  PLACEHOLDER (multi-GPU milestone).
int aep_halo_info(aep_ctx* c, int, int, int64_t*) { return fail(c, AEP_ERR_INVALID, "halo exchange not built yet"); }
int aep_halo_pack(aep_ctx* c, int, int, void**) { return fail(c, AEP_ERR_INVALID, "halo exchange not built yet"); }
int aep_halo_recv_buffer(aep_ctx* c, int, int, void**) { return fail(c, AEP_ERR_INVALID, "halo exchange not built yet"); }
int aep_halo_add(aep_ctx* c, int, int) { return fail(c, AEP_ERR_INVALID, "halo exchange not built yet"); }
int aep_migrate_extract(aep_ctx* c, int64_t*, int64_t*, void**, void**) { return fail(c, AEP_ERR_INVALID, "migration not built yet"); }
int aep_migrate_recv_buffer(aep_ctx* c, int, int64_t, void**) { return fail(c, AEP_ERR_INVALID, "migration not built yet"); }
int aep_migrate_insert(aep_ctx* c, int64_t, int64_t) { return fail(c, AEP_ERR_INVALID, "migration not built yet"); }
