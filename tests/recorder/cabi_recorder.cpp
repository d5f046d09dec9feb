// tests/recorder/cabi_recorder.cpp -- a CALL RECORDER with the C ABI's symbol names.  TEST INFRASTRUCTURE.
//
// It computes NOTHING: no physics, no oracle, no CPU path of the engine.  Every entry point appends its name to a call log, keeps a
// copy of what it was handed, and "downloads" hand back recognisable patterns derived from the uploads (x + frame * (0.01, 0.02, 0.03),
// v = -x, ...).  tests/test_host_glue_recorder.py builds it as a stand-in libaep_b200.so in a temporary directory and runs the host
// bindings against it (LD_LIBRARY_PATH) to check, WITHOUT a GPU, the part of them that is pure plumbing: which arrays reach the ABI in
// which layout and order, how many frames a given maxt produces, what lands in the containers and in the OBJ files.  What the
// real library computes is the business of the GPU parity tests; nothing here can stand in for it.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/aep_b200.h"

struct aep_ctx {
    aep_config cfg;
    std::vector<std::string> calls;
    int64_t n = 0, nv = 0, nf = 0;
    std::vector<double> x, v, B1, B2, B3, FE, FP, m, vol, q; double mat[4] = {0, 0, 0, 0};
    std::vector<double> vx, vv, vm, vvol, vB, ev, em, evol, eB, ed, eD, fixedv; std::vector<int32_t> faces; double mpar[5] = {0, 0, 0, 0, 0};
    std::vector<uint8_t> inside; std::vector<double> normal;
    int frames = 0; int64_t substeps = 0;
};

namespace {
void put(FILE* f, const char* name, const double* p, int64_t len) {
    char nm[32] = {0}; std::strncpy(nm, name, 31); const int32_t dt = 0;
    std::fwrite(nm, 32, 1, f); std::fwrite(&dt, 4, 1, f); std::fwrite(&len, 8, 1, f); if (len) std::fwrite(p, 8, (size_t)len, f);
}
void putv(FILE* f, const char* name, const std::vector<double>& v) { put(f, name, v.data(), (int64_t)v.size()); }
void dump(aep_ctx* c) {
    const char* path = std::getenv("AEP_RECORDER_OUT");
    if (!path) return;
    FILE* f = std::fopen(path, "wb"); if (!f) return;
    int32_t cnt = 0; std::fwrite(&cnt, 4, 1, f);
    std::string log; for (const std::string& s : c->calls) { log += s; log += '\n'; }
    std::vector<double> logd(log.begin(), log.end());                                  // one char per double: the blob format has no text type
    putv(f, "calls", logd); ++cnt;
    const double cfg[13] = {(double)c->cfg.material, c->cfg.cfl, c->cfg.grid_min[0], c->cfg.grid_min[1], c->cfg.grid_min[2], c->cfg.grid_max[0], c->cfg.grid_max[1],
                            c->cfg.grid_max[2], (double)c->cfg.res[0], (double)c->cfg.res[1], (double)c->cfg.res[2], (double)c->frames, (double)c->substeps};
    put(f, "cfg", cfg, 13); ++cnt;
#define PV(name) putv(f, #name, c->name); ++cnt;
    PV(x) PV(v) PV(B1) PV(B2) PV(B3) PV(FE) PV(FP) PV(m) PV(vol) PV(q) PV(vx) PV(vv) PV(vm) PV(vvol) PV(vB) PV(ev) PV(em) PV(evol) PV(eB) PV(ed) PV(eD) PV(fixedv) PV(normal)
#undef PV
    put(f, "mat", c->mat, 4); ++cnt; put(f, "mpar", c->mpar, 5); ++cnt;
    std::vector<double> fd(c->faces.begin(), c->faces.end()), ind(c->inside.begin(), c->inside.end());
    putv(f, "faces", fd); ++cnt; putv(f, "inside", ind); ++cnt;
    std::fseek(f, 0, SEEK_SET); std::fwrite(&cnt, 4, 1, f); std::fclose(f);
}
void cp(std::vector<double>& dst, const double* src, int64_t len) { if (src) dst.assign(src, src + len); else dst.clear(); }
// pattern of a "download": out = a * uploaded + b   (element-wise; b may depend on the frame counter)
void pat(double* out, const std::vector<double>& up, double a, double b) { if (out) for (size_t i = 0; i < up.size(); ++i) out[i] = a * up[i] + b; }
}  // namespace

extern "C" {
int aep_default_config(aep_config* cfg) {
    std::memset(cfg, 0, sizeof *cfg);
    cfg->material = AEP_SAND; cfg->cfl = 0.3; cfg->gravity = 9.8; cfg->collider_friction = 0.2; cfg->snow_hardening = 10.0;
    cfg->sand_h[0] = 35; cfg->sand_h[1] = 9; cfg->sand_h[2] = 0.2; cfg->sand_h[3] = 10; cfg->dt_rate_floor = 300; cfg->frame_dt = 1.0 / 60.0; cfg->slab_axis = -1;
    return AEP_OK;
}
int aep_create(aep_ctx** out, const aep_config* cfg) { aep_ctx* c = new aep_ctx(); c->cfg = *cfg; c->calls.push_back("aep_create"); *out = c; return AEP_OK; }
int aep_destroy(aep_ctx* c) { c->calls.push_back("aep_destroy"); dump(c); delete c; return AEP_OK; }
const char* aep_last_error(aep_ctx*) { return "recorder: no error"; }
int aep_sync(aep_ctx* c) { c->calls.push_back("aep_sync"); return AEP_OK; }
int aep_upload_particles(aep_ctx* c, int64_t n, const double* x, const double* v, const double* B1, const double* B2, const double* B3, const double* FE,
                         const double* FP, const double* m, const double* vol, const double* q, double E, double nu, double tc, double ts) {
    c->calls.push_back("aep_upload_particles"); c->n = n;
    cp(c->x, x, 3 * n); cp(c->v, v, 3 * n); cp(c->B1, B1, 3 * n); cp(c->B2, B2, 3 * n); cp(c->B3, B3, 3 * n); cp(c->FE, FE, 9 * n); cp(c->FP, FP, 9 * n);
    cp(c->m, m, n); cp(c->vol, vol, n); cp(c->q, q, n); c->mat[0] = E; c->mat[1] = nu; c->mat[2] = tc; c->mat[3] = ts;
    return AEP_OK;
}
int aep_upload_mesh(aep_ctx* c, int64_t nv, int64_t nf, const double* vx, const double* vv, const double* vm, const double* vvol, const double* vB,
                    const int32_t* faces, const double* ev, const double* em, const double* evol, const double* eB, const double* ed, const double* eD,
                    const double* fixedv, double mu, double lambda, double shear, double stiff, double fric) {
    c->calls.push_back("aep_upload_mesh"); c->nv = nv; c->nf = nf;
    cp(c->vx, vx, 3 * nv); cp(c->vv, vv, 3 * nv); cp(c->vm, vm, nv); cp(c->vvol, vvol, nv); cp(c->vB, vB, 9 * nv); c->faces.assign(faces, faces + 3 * nf);
    cp(c->ev, ev, 3 * nf); cp(c->em, em, nf); cp(c->evol, evol, nf); cp(c->eB, eB, 9 * nf); cp(c->ed, ed, 9 * nf); cp(c->eD, eD, 9 * nf); cp(c->fixedv, fixedv, nv);
    c->mpar[0] = mu; c->mpar[1] = lambda; c->mpar[2] = shear; c->mpar[3] = stiff; c->mpar[4] = fric;
    return AEP_OK;
}
int aep_set_levelset_analytic(aep_ctx* c, int, const double*) { c->calls.push_back("aep_set_levelset_analytic"); return AEP_OK; }
int aep_set_levelset_samples(aep_ctx* c, const uint8_t* inside, const double* normal) {
    c->calls.push_back("aep_set_levelset_samples");
    const int64_t Ng = (int64_t)c->cfg.res[0] * c->cfg.res[1] * c->cfg.res[2];
    c->inside.assign(inside, inside + Ng); cp(c->normal, normal, 3 * Ng);
    return AEP_OK;
}
int aep_init(aep_ctx* c) { c->calls.push_back("aep_init"); return AEP_OK; }
int aep_run(aep_ctx* c, int n) { c->calls.push_back("aep_run"); c->substeps += n; return AEP_OK; }
int aep_substep(aep_ctx* c) { c->calls.push_back("aep_substep"); c->substeps += 1; return AEP_OK; }
int aep_run_frames(aep_ctx* c, int n_frames, int, int64_t* done) {
    c->calls.push_back("aep_run_frames"); c->frames += n_frames; c->substeps += 17 * n_frames; if (done) *done = 17 * n_frames; return AEP_OK;
}
int aep_get_clock(aep_ctx* c, double* dt, double* t, double* inner_t, int32_t* frame_no, int64_t* substeps, double* vmax, int64_t* escaped) {
    if (dt) *dt = 1e-3; if (t) *t = c->frames / 60.0; if (inner_t) *inner_t = 0; if (frame_no) *frame_no = c->frames; if (substeps) *substeps = c->substeps;
    if (vmax) *vmax = 0; if (escaped) *escaped = 0; return AEP_OK;
}
int64_t aep_num_particles(aep_ctx* c) { return c->n; }
// x = uploaded + frames * 0.01 (every coordinate), v = -x_up, B_a = a + B_up, FE = 2 FE_up, FP = 3 FP_up, vol = vol_up + 5, q = q_up + 7
int aep_download_particles(aep_ctx* c, double* x, double* v, double* B1, double* B2, double* B3, double* FE, double* FP, double* vol, double* q) {
    c->calls.push_back(v ? "aep_download_particles(all)" : "aep_download_particles(x)");
    pat(x, c->x, 1.0, 0.01 * c->frames); pat(v, c->x, -1.0, 0.0); pat(B1, c->B1, 1.0, 1.0); pat(B2, c->B2, 1.0, 2.0); pat(B3, c->B3, 1.0, 3.0);
    pat(FE, c->FE, 2.0, 0.0); pat(FP, c->FP, 3.0, 0.0); pat(vol, c->vol, 1.0, 5.0); pat(q, c->q, 1.0, 7.0);
    return AEP_OK;
}
// vx = uploaded + frames * 0.02, vv = -vx_up, vB = vB_up + 1, ex = 0.5 (constant), ev = ev_up + 4, eB = eB_up + 2, ed = 2 ed_up
int aep_download_mesh(aep_ctx* c, double* vx, double* vv, double* vB, double* ex, double* ev, double* eB, double* ed) {
    c->calls.push_back(vv ? "aep_download_mesh(all)" : "aep_download_mesh(x)");
    pat(vx, c->vx, 1.0, 0.02 * c->frames); pat(vv, c->vx, -1.0, 0.0); pat(vB, c->vB, 1.0, 1.0);
    if (ex) for (int64_t i = 0; i < 3 * c->nf; ++i) ex[i] = 0.5;
    pat(ev, c->ev, 1.0, 4.0); pat(eB, c->eB, 1.0, 2.0); pat(ed, c->ed, 2.0, 0.0);
    return AEP_OK;
}
int aep_download_grid(aep_ctx* c, double* m, double* v, double* f, double* vt) {
    c->calls.push_back("aep_download_grid");
    const int64_t Ng = (int64_t)c->cfg.res[0] * c->cfg.res[1] * c->cfg.res[2];
    if (m) for (int64_t i = 0; i < Ng; ++i) m[i] = 11.0;
    if (v) for (int64_t i = 0; i < 3 * Ng; ++i) v[i] = 12.0;
    if (f) for (int64_t i = 0; i < 3 * Ng; ++i) f[i] = 13.0;
    if (vt) for (int64_t i = 0; i < 3 * Ng; ++i) vt[i] = 14.0;
    return AEP_OK;
}
// the asynchronous per-frame download: float32 xyz per particle, same pattern as aep_download_particles(x)
int aep_frame_positions_begin(aep_ctx* c, float* xyz) {
    c->calls.push_back("aep_frame_positions_begin");
    for (int64_t p = 0; p < c->n; ++p) for (int a = 0; a < 3; ++a) xyz[3 * p + a] = (float)(c->x[(size_t)(a * c->n + p)] + 0.01 * c->frames);
    return AEP_OK;
}
int aep_frame_positions_wait(aep_ctx* c) { c->calls.push_back("aep_frame_positions_wait"); return AEP_OK; }
void* aep_host_alloc(int64_t bytes) { return bytes > 0 ? std::malloc((size_t)bytes) : nullptr; }
void aep_host_free(void* p) { std::free(p); }
int aep_set_fixed_dt(aep_ctx* c, double) { c->calls.push_back("aep_set_fixed_dt"); return AEP_OK; }
int aep_resume(aep_ctx* c) { c->calls.push_back("aep_resume"); return AEP_OK; }
int aep_set_clock(aep_ctx* c, double, double, double, int32_t fr, int64_t ss) { c->calls.push_back("aep_set_clock"); c->frames = fr; c->substeps = ss; return AEP_OK; }
}  // extern "C"
