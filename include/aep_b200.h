/* ============================================================================================
 * aep_b200.h -- C ABI of the B200-native MPM substep engine (libaep_b200.so).
 *
 * Drop-in boundary for the hot path of 2iw31Zhv/AnisotropicElastoplasticity:
 *   HybridSolver::solve's loop body, AnisotropicElastoplasticity/HybridSolver.cpp:867-1032
 *   ("HS:" below), and the functions it calls.  The reference has no FFI of its own; its boundary
 *   is the C++ class surface HybridSolver / ParticleSystem / RegularGrid / LagrangianMesh
 *   (HybridSolver.h:27-95).  include/aep/ *.h re-creates those classes on top of this ABI and
 *   INTEGRATION.md shows the few lines a maintainer changes in the reference to bind to it.
 *
 * Conventions
 *   - plain C: opaque handle, POD config struct, raw pointers + sizes, int return codes
 *     (0 = AEP_OK, negative = error; message via aep_last_error).  No exceptions cross the ABI.
 *   - host arrays use the reference's memory layouts:
 *       N x 3 matrix            column-major, leading dimension N      (Eigen::MatrixX3d)
 *       N 3x3 tensors           9 doubles each, column-major per item  (std::vector<Eigen::Matrix3d>)
 *       grid node index         k*nx*ny + j*nx + i                     (RegularGrid.cpp:164-168)
 *     fp64 at the boundary, fp32 on the device; conversion happens on the GPU at upload/download.
 *   - one aep_ctx is single-threaded (caller serialises).  All stepping calls are asynchronous on the
 *     context's CUDA stream; downloads and aep_sync synchronise.
 *   - there is NO CPU fallback: aep_create fails with AEP_ERR_CUDA when no sm_100 device is usable.
 * ============================================================================================ */
#ifndef AEP_B200_H
#define AEP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define AEP_API __attribute__((visibility("default")))
#else
#define AEP_API
#endif

#define AEP_OK 0
#define AEP_ERR_INVALID (-1)   /* bad argument / call order            */
#define AEP_ERR_CUDA (-2)      /* CUDA runtime error (see last_error)  */
#define AEP_ERR_ALLOC (-3)     /* out of device / host memory          */
#define AEP_ERR_STATE (-4)     /* particles left the grid, NaN, ...    */

/* MaterialType, HybridSolver.h:21-25 */
#define AEP_SNOW 0
#define AEP_SAND 1

/* Level-set primitives.  1,2 = LevelSet.cpp:8-42; 3,4 are new (BASELINE configs C2/C3/C5 need a sphere / a box);
 * 5 = arbitrary std::function sampled at the grid nodes by the host (HS:473-482 only ever evaluates phi there). */
#define AEP_LS_NONE 0
#define AEP_LS_GROUND 1         /* params: groundZ                                  */
#define AEP_LS_WALL2GROUND 2    /* params: wallX, wallY, groundZ                    */
#define AEP_LS_SPHERE_GROUND 3  /* params: cx, cy, cz, radius, groundZ              */
#define AEP_LS_BOX 4            /* params: x0, y0, z0, x1, y1, z1 (inside is free)  */
#define AEP_LS_SAMPLED 5
#define AEP_LS_SPHERE 6         /* params: cx, cy, cz, radius  (no ground: a free, possibly moving, ball) */

typedef struct aep_ctx aep_ctx;

/* Every literal the reference hard-codes on the hot path, with its default (SURVEY.md section 5 "Config / flags"). */
typedef struct aep_config {
    int32_t device;            /* CUDA device ordinal                                                        */
    int32_t material;          /* AEP_SAND: literal at HS:873,955,959                                        */
    double grid_min[3];        /* RegularGrid ctor, RegularGrid.cpp:117-162                                  */
    double grid_max[3];
    int32_t res[3];
    int32_t _pad0;
    double cfl;                /* 0.3, main.cpp:27                                                           */
    double gravity;            /* 9.8 along -z, HS:457                                                       */
    double collider_friction;  /* 0.2, HS:465                                                                */
    double snow_hardening;     /* 10, HS:267                                                                 */
    double sand_h[4];          /* 35, 9, 0.2, 10, HS:641-644                                                 */
    double dt_rate_floor;      /* 300, HS:860,878   dt = cfl / max(rate_floor, vmax/hmin)                    */
    double frame_dt;           /* 1/60, HS:880-884                                                           */
    int64_t particle_capacity; /* 0 = exactly what aep_upload_particles is given, allocated by the first upload; >0: the
                                  particle arrays are allocated by aep_create for this many particles (room for particles
                                  migrating in from neighbouring ranks; keeps cudaMalloc out of the upload)  */
    /* spatial slab owned by this context (multi-GPU).  Cells [slab_lo, slab_hi) along slab_axis.
     * slab_axis < 0: the context owns the whole grid.                                                        */
    int32_t slab_axis;
    int32_t slab_lo;
    int32_t slab_hi;
    int32_t sort_every;        /* physical re-sort of the particle arrays: k >= 1 every k-th substep; 0 (default) adaptive: when the
                                  accumulated out-of-order fraction (particles that changed cell since the last sort, summed over
                                  substeps) reaches sort_cost_threshold, or after 32 substeps.  Results never depend on the order. */
    int32_t sort_bricks;       /* 0 (default): order particles by cell index, x fastest; 1: by 4x4x4-cell brick, then cell (better L1 hit rate
                                  of the gathers, but the scatters of one CTA then contend for the same L2 lines: measured slower overall) */
    int32_t scatter_strips;    /* 64: concurrently running P2G CTAs are spread over this many far-apart parts of the sorted particle
                                  order so that they do not reduce into the same grid nodes at the same time (L2 atomic contention) */
    double sort_cost_threshold; /* 0.06: measured on the C5 dam break in flow (profiles/README.md, round 2): a particle that has left its
                                  cell costs ~1.2 ms per % and substep (list pass + broken scatter runs), a re-sort 5.9 ms              */
    /* --- opt-in departures from the reference (all default to the reference's behaviour) --- */
    double vmax_min_mass_fraction; /* 0: the reference's dt rule, max |v_i| over EVERY node with m_i > 0 (RegularGrid.cpp:188-200), which at
                                  nearly massless stencil-edge nodes is rounding noise of p/m and can cut dt by orders of magnitude.
                                  > 0 (NOT the reference): nodes lighter than this fraction of one particle's mass do not enter max |v_i| */
    int32_t coulomb_friction;  /* 0: the reference's stick-or-slide collider (HS:494-502; the Coulomb reduction at :500-501 is a no-op
                                  expression).  1: |v_t| is reduced by mu |v_n| as those lines set out to do                          */
    int32_t use_graph;         /* 1 (default): a substep is replayed from a CUDA graph -- one host launch.  0: launch kernel by kernel */
} aep_config;

AEP_API int aep_default_config(aep_config* cfg);
AEP_API int aep_create(aep_ctx** out, const aep_config* cfg);
AEP_API int aep_destroy(aep_ctx* ctx);
/* ctx may be NULL: returns the message of the last failed aep_create on this thread. */
AEP_API const char* aep_last_error(aep_ctx* ctx);
AEP_API int aep_sync(aep_ctx* ctx);

/* ---- state upload ---------------------------------------------------------------------------------------
 * ParticleSystem public members (ParticleSystem.h:22-41).  B1,B2,B3 = affineMomenta_{1,2,3} (rows of B).
 * E, nu, theta_c, theta_s = youngsModulus, poissonRatio, criticalCompression, criticalStretch.              */
AEP_API int aep_upload_particles(aep_ctx* ctx, int64_t n, const double* x, const double* v, const double* B1,
                                 const double* B2, const double* B3, const double* FE, const double* FP,
                                 const double* m, const double* vol, const double* q, double E, double nu,
                                 double theta_c, double theta_s);

/* LagrangianMesh public members (LagrangianMesh.h:38-74).  vB/eB: three consecutive N x 3 blocks (rows of B);
 * ed/eD: elementDirections_{1,2,3} / rest directions as three consecutive nf x 3 blocks; faces nf x 3
 * column-major int32; fixedv: nv doubles (bindConstraints, LagrangianMesh.cpp:354-363) or NULL.            */
AEP_API int aep_upload_mesh(aep_ctx* ctx, int64_t nv, int64_t nf, const double* vx, const double* vv,
                            const double* vm, const double* vvol, const double* vB, const int32_t* faces,
                            const double* ev, const double* em, const double* evol, const double* eB,
                            const double* ed, const double* eD, const double* fixedv, double mu, double lambda,
                            double shear_stiffness, double stiffness, double friction_coeff);

/* HybridSolver::setLevelSet (HybridSolver.h:89).  Colliders are static (HS:484) and only sampled at nodes.  */
AEP_API int aep_set_levelset_analytic(aep_ctx* ctx, int kind, const double* params8);
AEP_API int aep_set_levelset_samples(aep_ctx* ctx, const uint8_t* inside /*Ng*/, const double* normal /*Ng x 3*/);
/* Moving collider (not in the reference, whose colliders are static: HS:484 sets the collider velocity to zero).  The analytic level
 * set given to aep_set_levelset_analytic translates rigidly with `velocity3`; it is then evaluated at the grid nodes on the device
 * every substep at its current position, and the projection of HS:486-502 works on the velocity relative to the collider.
 * NULL or a zero vector: back to the static, host-sampled collider.                                                          */
AEP_API int aep_set_collider_motion(aep_ctx* ctx, const double* velocity3);

/* ---- stepping ------------------------------------------------------------------------------------------- */
/* HS:830-860: bin particles, first particleToGrid_ (computes volumes, HS:242-249), initial dt.             */
AEP_API int aep_init(aep_ctx* ctx);
/* the same in three parts for the multi-GPU driver: begin (re-bin + P2G) | halo(m,p) | volumes (+ local max|v|) | all-reduce | dt */
AEP_API int aep_init_begin(aep_ctx* ctx);
AEP_API int aep_init_volumes(aep_ctx* ctx);
AEP_API int aep_init_dt(aep_ctx* ctx);
AEP_API int aep_init_dt_async(aep_ctx* ctx);      /* the same without the final stream synchronisation */
/* One iteration of the while loop HS:867-1032 with the reference's dt rule evaluated on the device.        */
AEP_API int aep_substep(aep_ctx* ctx);
/* n iterations back to back, no host synchronisation in between.                                           */
AEP_API int aep_run(aep_ctx* ctx, int n_substeps);
/* Run until `n_frames` more 1/60 s frames have completed (HS:880-892) or `max_substeps` substeps were taken; returns the substeps
 * taken.  The stop condition is evaluated on the device (the clock kernel halts the context, substeps queued behind the halt return
 * at once), the host only polls every few substeps: the last substep of a frame is clipped exactly as at HS:880-884.
 * Returns AEP_ERR_STATE when particles left the grid / became NaN (the state is then clamped garbage), or a peer timed out.   */
AEP_API int aep_run_frames(aep_ctx* ctx, int n_frames, int max_substeps, int64_t* substeps_done);

/* Stage-level entry points for the parity tests; same split as the reference's private methods.            */
AEP_API int aep_p2g(aep_ctx* ctx, int first);                 /* re-bin + particleToGrid_      HS:113-250, 18-97 */
AEP_API int aep_stage_forces(aep_ctx* ctx, double dt);        /* computeGridForces_            HS:252-458        */
AEP_API int aep_stage_grid(aep_ctx* ctx, double dt);          /* updateGridVelocities_ + max|v| + gridCollisionHandling_  HS:725-737, 460-551 */
AEP_API int aep_stage_g2p(aep_ctx* ctx, double dt);           /* HS:739-825, 940-951, 553-609, 612-723           */

AEP_API int aep_set_dt(aep_ctx* ctx, double dt);
/* dt > 0: pin the time step (the adaptive rule of HS:878-892 is bypassed; frames are still counted every frame_dt).
 * dt <= 0: back to the reference rule.  Not a reference feature: used for well-conditioned long-run parity tests. */
AEP_API int aep_set_fixed_dt(aep_ctx* ctx, double dt);
/* dt, simulated time t, time inside the current frame, frames completed, substeps done, max |v_i| of the last
 * grid update, number of particles that tried to leave the grid (sticky).  Any pointer may be NULL.        */
AEP_API int aep_get_clock(aep_ctx* ctx, double* dt, double* t, double* inner_t, int32_t* frame_no,
                          int64_t* substeps, double* vmax, int64_t* escaped);

/* ---- restart (the reference has none) ---------------------------------------------------
 * A checkpoint is what the downloads below return plus aep_get_clock's dt, t, inner_t, frame_no, substeps.  To continue
 * from one: aep_create, upload the saved state (vol = the saved volumes), the collider, then aep_resume INSTEAD of
 * aep_init (bins the particles and runs the mass/momentum P2G of HS:987; does not recompute volumes, HS:242-249, nor
 * the initial dt, HS:860) and aep_set_clock.  The next aep_substep is then the one the saved run would have done.
 * aep_set_fixed_dt is not part of the clock: re-apply it if the saved run used it.  Tested for whole-grid contexts; a slab
 * decomposition is restarted by re-partitioning the saved global state (distributed.py), not rank by rank.               */
AEP_API int aep_resume(aep_ctx* ctx);
AEP_API int aep_set_clock(aep_ctx* ctx, double dt, double t, double inner_t, int32_t frame_no, int64_t substeps);
/* the sticky "particles left the grid" counter of aep_get_clock, so that a restart carries it on */
AEP_API int aep_set_escaped(aep_ctx* ctx, int64_t escaped);

/* ---- state download (original particle order, reference layouts; any pointer may be NULL) --------------- */
AEP_API int64_t aep_num_particles(aep_ctx* ctx);
AEP_API int aep_download_particles(aep_ctx* ctx, double* x, double* v, double* B1, double* B2, double* B3,
                                   double* FE, double* FP, double* vol, double* q);
/* RegularGrid::masses / velocities / forces (RegularGrid.h:36-41) and gridVelocitiesBeforeFriction (HS:460-463).
 * After aep_p2g: v = p/m.  After aep_stage_grid: v = post-collision, vt = pre-friction.                     */
AEP_API int aep_download_grid(aep_ctx* ctx, double* m, double* v, double* f, double* vt);
AEP_API int aep_download_mesh(aep_ctx* ctx, double* vx, double* vv, double* vB, double* ex, double* ev,
                              double* eB, double* ed);
/* positions only, float32 x,y,z interleaved: the per-frame OBJ payload of HS:991-1007.  Original particle order; for a slab
 * context (ids are global there) the context's current order, aep_num_particles entries.                    */
AEP_API int aep_download_positions_f32(aep_ctx* ctx, float* xyz);
/* Asynchronous variant for a frame loop: begin snapshots the positions in stream order (behind everything queued so far) and copies
 * them to `pinned_xyz` (cudaMallocHost / pinned memory, 3 floats per particle) on a second stream while the following substeps
 * compute; wait blocks until the copy has landed.  One download can be in flight per context.                     */
AEP_API int aep_frame_positions_begin(aep_ctx* ctx, float* pinned_xyz);
AEP_API int aep_frame_positions_wait(aep_ctx* ctx);
/* page-locked host memory for the asynchronous download (NULL on failure); plain host code needs no CUDA runtime of its own      */
AEP_API void* aep_host_alloc(int64_t bytes);
AEP_API void aep_host_free(void* p);
/* bulk statistics on the device: centre of mass (3), kinetic energy, mean det F_P, total mass.             */
AEP_API int aep_stats(aep_ctx* ctx, double* com3, double* kinetic, double* mean_jp, double* mass);

/* ---- instrumentation ------------------------------------------------------------------------------------ */
/* 8x8x8-node blocks touched by the last P2G and nodes with m_i > 0 (the "active nodes" of the roofline model).  */
AEP_API int aep_grid_activity(aep_ctx* ctx, int64_t* active_blocks, int64_t* active_nodes);
/* Number of kernels this library launched since aep_create (its own kernels and the cub sort passes).      */
AEP_API int64_t aep_kernel_launches(aep_ctx* ctx);
/* The CUDA stream all work is enqueued on (cudaStream_t as void*), for event timing by the caller.          */
AEP_API void* aep_stream(aep_ctx* ctx);
/* Per-stage device time accumulated since the last reset when profiling is enabled (costs 2 events/stage).  */
AEP_API int aep_profile(aep_ctx* ctx, int enable);
AEP_API int aep_get_timers(aep_ctx* ctx, double* ms /*AEP_NUM_STAGES*/, int64_t* calls /*AEP_NUM_STAGES*/);
#define AEP_STAGE_SORT 0
#define AEP_STAGE_P2G 1
#define AEP_STAGE_FORCES 2
#define AEP_STAGE_GRID 3
#define AEP_STAGE_G2P 4
#define AEP_STAGE_MESH 5
#define AEP_STAGE_HALO 6
#define AEP_STAGE_G2P2G 7   /* the fused kernel (development, AEP_FUSED=1): G2P of this substep + P2G of the next */
#define AEP_STAGE_FORCE_SCATTER 8   /* k_force_scatter; AEP_STAGE_FORCES is the gather / stress kernel over the cell-sorted particles */
#define AEP_STAGE_FORCES_LIST 9     /* the force kernel's pass over the deferred particles (strays) */
#define AEP_STAGE_G2P_LIST 10       /* the G2P (or fused) kernel's pass over the deferred particles */
#define AEP_NUM_STAGES 11
/* physical re-sorts done so far, particle slots in use / dead (slab contexts), particles that changed cell since the last re-sort */
AEP_API int aep_get_counters(aep_ctx* ctx, int64_t* sorts, int64_t* slots, int64_t* dead, int64_t* moved_since_sort);
/* peer-memory exchange: particles this context has handed to / taken from its neighbours so far (HybridSolver.cpp:940-951 moves
 * particles across slab boundaries; the counts live on the device)                                                               */
AEP_API int aep_get_migration(aep_ctx* ctx, int64_t* sent, int64_t* received);

/* ---- multi-GPU slab decomposition -------------------------------------------------------------
 * The context owns the particles whose cell index along cfg.slab_axis lies in [slab_lo, slab_hi).  Their cubic
 * stencils reach node planes slab_lo-1 .. slab_hi+1, so two neighbouring slabs both scatter into the 3 node planes
 * b-1, b, b+1 around their common boundary b.  One symmetric exchange per scatter completes them: each side packs
 * its partial sums of those planes, the caller swaps the buffers (NCCL send/recv, peer copy, ...), each side adds
 * what it received (a+b == b+a bitwise, so both ranks hold identical totals and update the shared planes redundantly).
 *   side 0 = low neighbour, 1 = high neighbour;   what 0 = (m, px, py, pz) after P2G, 1 = (fx, fy, fz, -) after forces.
 * Communication buffers are DEVICE memory OWNED BY THE CALLER (e.g. torch tensors), float32,
 * layout [plane 0..2][node in plane, fastest remaining axis first][4].                                          */
AEP_API int aep_halo_info(aep_ctx* ctx, int what, int side, int64_t* n_floats);
AEP_API int aep_halo_pack(aep_ctx* ctx, int what, int side, void* dev_send);
AEP_API int aep_halo_add(aep_ctx* ctx, int what, int side, const void* dev_recv);
/* max |v_i| of the last grid update as one float32 in caller-owned device memory: get -> all-reduce(max) -> set,
 * between aep_step_grid and aep_step_g2p (the dt rule needs the global maximum, HS:878).                         */
AEP_API int aep_vmax_get(aep_ctx* ctx, void* dev_float);
AEP_API int aep_vmax_set(aep_ctx* ctx, const void* dev_float);
/* split stepping for the multi-GPU driver:  forces | halo(f) | grid | vmax | g2p | migrate | p2g | halo(m,p)      */
AEP_API int aep_step_forces(aep_ctx* ctx);
AEP_API int aep_step_grid(aep_ctx* ctx);
AEP_API int aep_step_g2p(aep_ctx* ctx);   /* advances the clock (dt rule) then G2P / advection / plasticity          */
AEP_API int aep_step_p2g(aep_ctx* ctx);   /* re-bin (drops particles that left the slab) + P2G                       */
/* particle migration after aep_step_g2p: records of AEP_MIGRATE_FLOATS float32 (the 11 float4 of a particle).
 * extract copies the particles whose new cell left the slab into the caller's device buffers (capacity in records)
 * and returns the counts (synchronises); insert appends received records.  Global particle ids travel with them.  */
#define AEP_MIGRATE_FLOATS 44
AEP_API int aep_migrate_extract(aep_ctx* ctx, void* dev_to_low, void* dev_to_high, int64_t capacity, int64_t* n_low, int64_t* n_high);
AEP_API int aep_migrate_insert(aep_ctx* ctx, const void* dev_from_low, int64_t n_from_low, const void* dev_from_high, int64_t n_from_high);
/* Sync-free migration.  bind once: caller-owned device send buffers (capacity records each) and a device int64[2] for the
 * (low, high) leaver counts.  From then on aep_step_g2p lists the leavers while it advects them; extract_begin gathers the
 * listed particles into the send buffers (their slots stay behind as massless dead slots until the next re-sort) without any
 * host synchronisation; the caller ships the counts and records, then reports the counts it read back with extract_end.
 * aep_step_p2g may run between begin and end (it overlaps the count round trip); particles appended afterwards by
 * aep_migrate_insert are transferred with aep_step_p2g_arrivals (P2G is additive).                                        */
AEP_API int aep_migrate_bind(aep_ctx* ctx, void* dev_to_low, void* dev_to_high, int64_t capacity, void* dev_counts_i64x2);
AEP_API int aep_migrate_extract_begin(aep_ctx* ctx);
AEP_API int aep_migrate_extract_end(aep_ctx* ctx, int64_t n_low, int64_t n_high);
AEP_API int aep_step_p2g_arrivals(aep_ctx* ctx, int64_t count);
/* ids of uploaded particles are id_base + index (default 0); set before aep_upload_particles on each rank.      */
AEP_API int aep_set_particle_id_base(aep_ctx* ctx, int64_t id_base);
/* download in the context's current (cell-sorted) order together with the global ids; arrays sized aep_num_particles */
AEP_API int aep_download_particles_local(aep_ctx* ctx, int64_t* ids, double* x, double* v, double* B1, double* B2, double* B3,
                                         double* FE, double* FP, double* vol, double* q);


/* ---- peer-memory exchange: the multi-GPU substep without the host in the loop --------------------------------------------------
 * For y- or z-slab contexts (slab_axis 1 or 2; such a context allocates only the node planes of its slab, laid out with the slab
 * axis slowest).  Every rank exports a 256-byte blob describing its communication block (a CUDA IPC handle between processes, a
 * plain pointer inside one process); after aep_comm_connect with the blobs of ALL ranks (rank order = slab order along the axis),
 * aep_init / aep_substep / aep_run / aep_run_frames carry out the whole exchange themselves: the kernels of a substep store halo planes
 * (3 for f; 4 for (m,p) because the fused G2P+P2G scatters a particle that has just crossed the boundary before it migrates),
 * migrating particles and max|v| into the neighbour's memory over NVLink and publish epoch flags; the receiving stream waits for the
 * flag on the device.  No host synchronisation per substep, no collective-library call; particle counts live on the device.
 * Cloth: every rank holds the whole mesh (upload it on every rank BEFORE aep_comm_export), transfers only the points in its slab and
 * pushes what it advected to all ranks.  Device-side overflow (migration buffers, particle capacity) and peer timeouts are sticky and
 * surface as AEP_ERR_STATE from aep_get_clock / aep_run_frames.                                                                     */
#define AEP_COMM_BLOB_BYTES 256
AEP_API int aep_comm_export(aep_ctx* ctx, void* blob256, int64_t migrate_capacity /* particles per side and substep */);
AEP_API int aep_comm_connect(aep_ctx* ctx, int rank, int world, const void* blobs /* world x 256 bytes */);
/* several slab contexts inside ONE process (tests; one GPU or several): connect them, initialise and step them in lockstep from one
 * thread (a context's stream waits on the device for its neighbours, so their work must be queued before the host blocks).       */
AEP_API int aep_comm_connect_local(aep_ctx** ctxs, int world, int64_t migrate_capacity);
AEP_API int aep_group_init(aep_ctx** ctxs, int world);
AEP_API int aep_group_run(aep_ctx** ctxs, int world, int n_substeps);

#ifdef __cplusplus
}
#endif
#endif /* AEP_B200_H */
