// ref_driver.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// C entry points around the reference's OWN classes (HybridSolver, ParticleSystem, RegularGrid, LagrangianMesh, compiled
// unmodified from /root/reference by oracle/Makefile against the MiniEigen stand-in, oracle/ref_shim).  The entry points
// mirror oracle/mpm_oracle.cpp's `orc_*` API one for one (same arguments, same memory layouts), so a test can run the same
// scene through the reference, the oracle restatement and the GPU engine and compare stage by stage.
//
// Two ways in:
//   * ref_solve()    calls HybridSolver::solve (HybridSolver.cpp:827-1034) itself: the reference's own time loop, dt rule,
//                    frame clipping and OBJ frame writer.  Hard-wired by the reference to SAND and a rate floor of 300.
//   * ref_substep() / ref_stage_*()  call the reference's private stage methods (HybridSolver.cpp:113-825) in the order of
//                    the loop body HS:867-988.  The stage methods are the reference's code; only the ~40 lines of glue
//                    between them (HS:878-892 dt rule, HS:908-950 affine/advection calls) are restated here, because
//                    solve() cannot be entered mid-way, cannot run SNOW, and cannot pin dt.
// `#define private public` gives this file (only) access to the private members; it does not change object layout.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <functional>
#include <iostream>
#include <mutex>
#include <new>
#include <sstream>
#include <string>
#include <tuple>
#include <vector>
#include <unistd.h>
#include <Eigen/Sparse>

#define private public
#include "HybridSolver.h"
#include "LagrangianMesh.h"
#include "ParticleSystem.h"
#include "RegularGrid.h"
#undef private
#include "LevelSet.h"
#include "geometry.h"
#include "interpolation.h"

using namespace Eigen;

namespace {

enum { LS_NONE = 0, LS_GROUND = 1, LS_WALL2GROUND = 2, LS_SAMPLED = 5 };

struct RefSim {
    RegularGrid* rg = nullptr;
    ParticleSystem* ps = nullptr;
    LagrangianMesh* mesh = nullptr;
    HybridSolver* solver = nullptr;           // lives in zero-filled storage: the reference's ctor leaves mesh_ unset (HybridSolver.h:83-84)
    void* solver_mem = nullptr;
    VectorXd fixed;                           // borrowed by the mesh through bindConstraints (LagrangianMesh.cpp:354-363)
    MatrixX3d vt;                             // gridVelocityBeforeFriction of the running iteration (HS:897)
    int material = SAND;
    double cfl = 0.3, rate_floor = 3e2, frame_dt = 1.0 / 60.0;
    double t = 0.0, inner_t = 0.0, dt = 0.0; int frame_flag = 0, frame_no = 0;
    int ls_kind = LS_NONE; double ls_par[8] = {0};
    std::vector<uint8_t> ls_inside; std::vector<double> ls_nrm;
    double tm[5] = {0, 0, 0, 0, 0};
    std::string err;
    ~RefSim() { if (solver) solver->~HybridSolver(); std::free(solver_mem); delete mesh; delete ps; delete rg; }
};

MaterialType mat(const RefSim& S) { return S.material == 0 ? SNOW : SAND; }

void fill(MatrixX3d& M, const double* src, long n) { M.resize(static_cast<int>(n), 3); if (n) std::memcpy(M.data(), src, sizeof(double) * 3 * n); }
void fillv(VectorXd& v, const double* src, long n) { v.resize(static_cast<int>(n)); if (n) std::memcpy(v.data(), src, sizeof(double) * n); }
void out3(double* dst, const MatrixX3d& M) { if (dst && M.rows()) std::memcpy(dst, M.data(), sizeof(double) * 3 * M.rows()); }
void outv(double* dst, const VectorXd& v) { if (dst && v.size()) std::memcpy(dst, v.data(), sizeof(double) * v.size()); }

void install_levelset(RefSim& S) {
    using namespace std::placeholders;
    const double* P = S.ls_par;
    if (S.ls_kind == LS_GROUND)                                    // LevelSet.cpp:8-16
        S.solver->setLevelSet(std::bind(groundLevelSet, _1, P[0]), std::bind(DgroundLevelSet, _1, P[0]));
    else if (S.ls_kind == LS_WALL2GROUND)                          // LevelSet.cpp:18-42
        S.solver->setLevelSet(std::bind(wall2groundLevelSet, _1, P[0], P[1], P[2]), std::bind(Dwall2groundLevelSet, _1, P[0], P[1], P[2]));
    else if (S.ls_kind == LS_SAMPLED) {
        // any other collider: phi <= 0 flags and normals sampled at the nodes, the only places HS:473-482 evaluates them
        RefSim* s = &S;
        auto node = [s](const Vector3d& x) {
            const RegularGrid& g = *s->rg;
            int i = static_cast<int>(std::lround((x[0] - g.minBound_[0]) / g.h_[0]));
            int j = static_cast<int>(std::lround((x[1] - g.minBound_[1]) / g.h_[1]));
            int k = static_cast<int>(std::lround((x[2] - g.minBound_[2]) / g.h_[2]));
            return g.toIndex(i, j, k);
        };
        S.solver->setLevelSet([s, node](const Vector3d& x) { return s->ls_inside[node(x)] ? -1.0 : 1.0; },
                              [s, node](const Vector3d& x) { long id = node(x), Ng = s->rg->gridNumber(); return Vector3d(s->ls_nrm[id], s->ls_nrm[Ng + id], s->ls_nrm[2 * Ng + id]); });
    } else                                                          // no collider: phi > 0 everywhere
        S.solver->setLevelSet([](const Vector3d&) { return 1.0; }, [](const Vector3d&) { return Vector3d(0.0, 0.0, 1.0); });
}

void rebuild_weights(RefSim& S) {                                  // HS:830-850 == HS:963-983
    HybridSolver& H = *S.solver;
    if (H.ps_) H.evaluateInterpolationWeights_(H.omegas_, H.domegas_1_, H.domegas_2_, H.domegas_3_, H.ps_->positions);
    if (H.mesh_) {
        H.evaluateInterpolationWeights_(H.vertexOmegas_, H.dvertexOmegas_1_, H.dvertexOmegas_2_, H.dvertexOmegas_3_, H.mesh_->vertexPositions);
        H.evaluateInterpolationWeights_(H.elementOmegas_, H.delementOmegas_1_, H.delementOmegas_2_, H.delementOmegas_3_, H.mesh_->elementPositions);
    }
}

// the calls of HS:903-959 in the reference's order
void g2p_block(RefSim& S, double Dt) {
    HybridSolver& H = *S.solver;
    H.updateParticleVelocities_(0.95, Dt);                          // HS:903 (alpha is unused by the reference)
    if (H.ps_) H.updateAffineMomenta_(H.ps_->affineMomenta_1, H.ps_->affineMomenta_2, H.ps_->affineMomenta_3, H.omegas_, H.ps_->positions, H.ps_->velocities, 0.0);   // HS:908-917
    if (H.mesh_) {                                                  // HS:918-935
        H.updateAffineMomenta_(H.mesh_->vertexAffineMomenta_1, H.mesh_->vertexAffineMomenta_2, H.mesh_->vertexAffineMomenta_3, H.vertexOmegas_, H.mesh_->vertexPositions, H.mesh_->vertexVelocities, 1.0);
        H.updateAffineMomenta_(H.mesh_->elementAffineMomenta_1, H.mesh_->elementAffineMomenta_2, H.mesh_->elementAffineMomenta_3, H.elementOmegas_, H.mesh_->elementPositions, H.mesh_->elementVelocities, 1.0);
    }
    if (H.ps_) H.ps_->positions = H.omegas_ * (H.rg_->positions() + Dt * S.vt);                      // HS:942-945
    if (H.mesh_) { H.mesh_->vertexPositions = H.vertexOmegas_ * (H.rg_->positions() + Dt * S.vt); H.mesh_->updateElementPositions(); }   // HS:946-950
    H.updateDeformationGradient_(Dt, mat(S), S.vt);                 // HS:955
    H.updatePlasticity_(Dt, mat(S));                                // HS:959
}

double now_s() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }

}  // namespace

extern "C" {

typedef struct ref_sim ref_sim;   // opaque = RefSim
#define SIM(h) (reinterpret_cast<RefSim*>(h))

ref_sim* ref_create(const double* mn, const double* mx, const int* res) {
    std::clog.setstate(std::ios::failbit);                          // the reference narrates every stage on clog
    RefSim* S = new RefSim();
    VectorXd a(3), b(3);
    for (int i = 0; i < 3; ++i) { a[i] = mn[i]; b[i] = mx[i]; }
    S->rg = new RegularGrid(a, b, Vector3i(res[0], res[1], res[2]));     // RegularGrid.cpp:117-162
    S->rg->masses.setZero(); S->rg->velocities.setZero(); S->rg->forces.setZero();
    S->solver_mem = std::calloc(1, sizeof(HybridSolver));
    S->solver = new (S->solver_mem) HybridSolver(nullptr, S->rg);
    S->solver->setLagrangianMesh(nullptr);
    S->vt.resize(S->rg->gridNumber(), 3);
    install_levelset(*S);
    return reinterpret_cast<ref_sim*>(S);
}
void ref_destroy(ref_sim* h) { delete SIM(h); }
void ref_set_threads(ref_sim*, int) {}
int ref_get_threads(ref_sim*) { return 1; }                        // the reference is serial

// The reference hard-codes gravity 9.8 (HS:457), friction 0.2 (HS:465), hardening 10 (HS:267), the sand curve 35/9/0.2/10
// (HS:641-644) and the frame length 1/60 (HS:880-883): anything else cannot be run through its code.  Returns 0 / -1.
int ref_set_params(ref_sim* h, int material, double cfl, double gravity, double friction, double snow_xi, const double* sand_h4,
                   double rate_floor, double frame_dt) {
    RefSim* S = SIM(h);
    const double hh[4] = {35.0, 9.0, 0.2, 10.0};
    bool ok = gravity == 9.8 && friction == 0.2 && snow_xi == 10.0 && std::fabs(frame_dt - 1.0 / 60.0) < 1e-18;
    for (int i = 0; i < 4; ++i) ok = ok && sand_h4[i] == hh[i];
    if (!ok) { S->err = "parameters differ from the constants hard-coded in the reference"; return -1; }
    S->material = material; S->cfl = cfl; S->rate_floor = rate_floor; S->frame_dt = 1.0 / 60.0;
    return 0;
}
const char* ref_last_error(ref_sim* h) { return SIM(h)->err.c_str(); }

void ref_set_particles(ref_sim* h, long n, const double* x, const double* v, const double* B1, const double* B2, const double* B3,
                       const double* FE, const double* FP, const double* m, const double* vol, const double* q,
                       double E, double nu, double thetaC, double thetaS) {
    RefSim* S = SIM(h);
    MatrixX3d X, V, colors; VectorXd M, Vol, dens, Q;
    fill(X, x, n); fill(V, v, n); fillv(M, m, n); fillv(Vol, vol, n); fillv(Q, q, n);
    dens.resize(static_cast<int>(n)); dens.setOnes(); colors.resize(static_cast<int>(n), 3); colors.setOnes();
    std::vector<Matrix3d> fe(n), fp(n);
    for (long p = 0; p < n; ++p) { std::memcpy(fe[p].data(), FE + 9 * p, 72); std::memcpy(fp[p].data(), FP + 9 * p, 72); }
    delete S->ps;
    S->ps = new ParticleSystem(V, X, fe, fp, M, Vol, dens, Q, E, nu, thetaC, thetaS, 0.2, colors);      // ParticleSystem.cpp:81-117
    fill(S->ps->affineMomenta_1, B1, n); fill(S->ps->affineMomenta_2, B2, n); fill(S->ps->affineMomenta_3, B3, n);
    S->solver->setParticleSystem(S->ps);
}

void ref_set_mesh(ref_sim* h, long nv, long nf, const double* vx, const double* vv, const double* vm, const double* vvol,
                  const double* vB, const int* faces, const double* ev, const double* em, const double* evol, const double* eB,
                  const double* ed, const double* eD, const double* fixedv, double mu, double lambda, double shear, double stiff, double fric) {
    RefSim* S = SIM(h);
    MatrixX3d VX, VV, EV, d1, d2, d3, D1, D2, D3; VectorXd VM, VVol, EM, EVol; MatrixX3i F;
    fill(VX, vx, nv); fill(VV, vv, nv); fillv(VM, vm, nv); fillv(VVol, vvol, nv);
    fill(EV, ev, nf); fillv(EM, em, nf); fillv(EVol, evol, nf);
    fill(d1, ed, nf); fill(d2, ed + 3 * nf, nf); fill(d3, ed + 6 * nf, nf);
    fill(D1, eD, nf); fill(D2, eD + 3 * nf, nf); fill(D3, eD + 6 * nf, nf);
    F.resize(static_cast<int>(nf), 3); std::memcpy(F.data(), faces, sizeof(int) * 3 * nf);
    delete S->mesh;
    S->mesh = new LagrangianMesh(VX, F, VV, EV, VM, VVol, EM, EVol, d1, d2, d3, D1, D2, D3, mu, lambda, shear, stiff, fric);   // LagrangianMesh.cpp:127-195
    fill(S->mesh->vertexAffineMomenta_1, vB, nv); fill(S->mesh->vertexAffineMomenta_2, vB + 3 * nv, nv); fill(S->mesh->vertexAffineMomenta_3, vB + 6 * nv, nv);
    fill(S->mesh->elementAffineMomenta_1, eB, nf); fill(S->mesh->elementAffineMomenta_2, eB + 3 * nf, nf); fill(S->mesh->elementAffineMomenta_3, eB + 6 * nf, nf);
    S->fixed.resize(static_cast<int>(nv)); S->fixed.setZero();
    if (fixedv) std::memcpy(S->fixed.data(), fixedv, sizeof(double) * nv);
    S->mesh->bindConstraints(&S->fixed);                             // the ctor leaves vertexIsFixed_ unset
    S->solver->setLagrangianMesh(S->mesh);
}

void ref_set_levelset(ref_sim* h, int kind, const double* params8) {
    RefSim* S = SIM(h); S->ls_kind = kind; for (int i = 0; i < 8; ++i) S->ls_par[i] = params8[i];
    install_levelset(*S);
}
void ref_set_levelset_samples(ref_sim* h, const uint8_t* inside, const double* normal) {
    RefSim* S = SIM(h); long Ng = S->rg->gridNumber(); S->ls_kind = LS_SAMPLED;
    S->ls_inside.assign(inside, inside + Ng); S->ls_nrm.assign(normal, normal + 3 * Ng);
    install_levelset(*S);
}

void ref_rebuild_weights(ref_sim* h) { rebuild_weights(*SIM(h)); }
void ref_p2g(ref_sim* h, int first) { SIM(h)->solver->particleToGrid_(1e-20, first != 0); }                 // HS:854, HS:987
void ref_stage_forces(ref_sim* h, double dt) { SIM(h)->solver->computeGridForces_(dt, mat(*SIM(h))); }      // HS:873
void ref_stage_grid_update(ref_sim* h, double dt) { SIM(h)->solver->updateGridVelocities_(dt, 1e-20); }     // HS:877
double ref_cfl_condition(ref_sim* h) { return SIM(h)->rg->CFL_condition(); }                                // RegularGrid.h:60
void ref_stage_collide(ref_sim* h) { SIM(h)->solver->gridCollisionHandling_(SIM(h)->vt); }                  // HS:899
void ref_stage_g2p(ref_sim* h, double dt) { g2p_block(*SIM(h), dt); }

void ref_init(ref_sim* h) {                                          // HS:829-865
    RefSim* S = SIM(h);
    rebuild_weights(*S); S->solver->particleToGrid_(1e-20, true);
    S->t = 0.0; S->inner_t = 0.0; S->frame_flag = 0; S->frame_no = 0;
    S->dt = S->cfl / std::max(S->rate_floor, S->rg->CFL_condition());   // HS:860
}

double ref_substep(ref_sim* h) {                                     // one pass of HS:867-1032 (OBJ dump elided)
    RefSim* S = SIM(h); HybridSolver& H = *S->solver;
    double t0 = now_s();
    H.computeGridForces_(S->dt, mat(*S));                            // HS:873
    double t1 = now_s();
    H.updateGridVelocities_(S->dt, 1e-20);                           // HS:877
    S->dt = S->cfl / std::max(S->rate_floor, S->rg->CFL_condition());   // HS:878
    if (S->inner_t + S->dt >= S->frame_dt) { S->dt = S->frame_dt - S->inner_t; S->t += S->frame_dt; S->inner_t = 0.0; S->frame_flag = 1; }   // HS:880-888
    else S->inner_t += S->dt;                                        // HS:889-892
    H.gridCollisionHandling_(S->vt);                                 // HS:897-899
    double t2 = now_s();
    g2p_block(*S, S->dt);
    double t3 = now_s();
    rebuild_weights(*S);                                             // HS:963-983
    double t4 = now_s();
    H.particleToGrid_(1e-20, false);                                 // HS:987
    double t5 = now_s();
    if (S->frame_flag) { S->frame_no++; S->frame_flag = 0; }
    S->tm[0] += t1 - t0; S->tm[1] += t2 - t1; S->tm[2] += t3 - t2; S->tm[3] += t4 - t3; S->tm[4] += t5 - t4;
    return S->dt;
}

// HybridSolver::solve itself, run inside `workdir` (it writes particle/particle_N.obj and mesh/mesh_N.obj there).
// The loop condition is `t <= maxt` with t advancing by 1/60 per finished frame (HS:867,883).  Returns 0 / -1.
int ref_solve(ref_sim* h, double cfl, double maxt, const char* workdir) {
    RefSim* S = SIM(h);
    char cwd[4096];
    if (!getcwd(cwd, sizeof cwd) || chdir(workdir) != 0) { S->err = "cannot enter workdir"; return -1; }
    S->solver->solve(cfl, maxt, 0.95);                                // main.cpp:27
    if (chdir(cwd) != 0) { S->err = "cannot return from workdir"; return -1; }
    return 0;
}

void ref_set_dt(ref_sim* h, double dt) { SIM(h)->dt = dt; }
double ref_get_dt(ref_sim* h) { return SIM(h)->dt; }
double ref_get_time(ref_sim* h) { return SIM(h)->t; }
int ref_get_frame(ref_sim* h) { return SIM(h)->frame_no; }
void ref_get_timers(ref_sim* h, double* out5) { for (int i = 0; i < 5; ++i) out5[i] = SIM(h)->tm[i]; }

void ref_get_particles(ref_sim* h, double* x, double* v, double* B1, double* B2, double* B3, double* FE, double* FP, double* vol, double* q) {
    RefSim* S = SIM(h); if (!S->ps) return;
    const ParticleSystem& P = *S->ps;
    out3(x, P.positions); out3(v, P.velocities); out3(B1, P.affineMomenta_1); out3(B2, P.affineMomenta_2); out3(B3, P.affineMomenta_3);
    for (size_t p = 0; p < P.elasticDeformationGradients.size(); ++p) {
        if (FE) std::memcpy(FE + 9 * p, P.elasticDeformationGradients[p].data(), 72);
        if (FP) std::memcpy(FP + 9 * p, P.plasticDeformationGradients[p].data(), 72);
    }
    outv(vol, P.volumes); outv(q, P.plasticAmount);
}
void ref_get_grid(ref_sim* h, double* m, double* v, double* f, double* vt) {
    RefSim* S = SIM(h); outv(m, S->rg->masses); out3(v, S->rg->velocities); out3(f, S->rg->forces); out3(vt, S->vt);
}
void ref_set_grid(ref_sim* h, const double* m, const double* v, const double* f) {
    RefSim* S = SIM(h); long Ng = S->rg->gridNumber();
    if (m) fillv(S->rg->masses, m, Ng);
    if (v) fill(S->rg->velocities, v, Ng);
    if (f) fill(S->rg->forces, f, Ng);
}
void ref_get_mesh(ref_sim* h, double* vx, double* vv, double* vB, double* ex, double* ev, double* eB, double* ed) {
    RefSim* S = SIM(h); if (!S->mesh) return;
    const LagrangianMesh& M = *S->mesh; long nv = M.vertexPositions.rows(), nf = M.faces.rows();
    out3(vx, M.vertexPositions); out3(vv, M.vertexVelocities);
    if (vB) { out3(vB, M.vertexAffineMomenta_1); out3(vB + 3 * nv, M.vertexAffineMomenta_2); out3(vB + 6 * nv, M.vertexAffineMomenta_3); }
    out3(ex, M.elementPositions); out3(ev, M.elementVelocities);
    if (eB) { out3(eB, M.elementAffineMomenta_1); out3(eB + 3 * nf, M.elementAffineMomenta_2); out3(eB + 6 * nf, M.elementAffineMomenta_3); }
    if (ed) { out3(ed, M.elementDirections_1); out3(ed + 3 * nf, M.elementDirections_2); out3(ed + 6 * nf, M.elementDirections_3); }
}

// ---- the reference's scalar kernels, for known-answer tests of the oracle's restatements
double ref_cubic_bspline(double x) { return cubic_B_spline(x); }                      // interpolation.cpp:9-16
double ref_dcubic_bspline(double x) { return Dcubic_B_spline(x); }                    // interpolation.cpp:18-33
double ref_clamp(double x, double lo, double hi) { return clamp(x, lo, hi); }         // interpolation.cpp:35-49
void ref_gram_schmidt(const double* A9, double* Q9, double* R9) {                     // geometry.cpp:31-62
    Matrix3d A, Q, R; std::memcpy(A.data(), A9, 72); GramSchmidtOrthonomalization(Q, R, A);
    std::memcpy(Q9, Q.data(), 72); std::memcpy(R9, R.data(), 72);
}
void ref_inverse_r(const double* R4, double* out4) { Matrix2d R; std::memcpy(R.data(), R4, 32); Matrix2d o = inverseR(R); std::memcpy(out4, o.data(), 32); }   // geometry.cpp:67-73
double ref_ls_phi(int kind, const double* P, const double* x) {
    Vector3d p(x[0], x[1], x[2]);
    return kind == LS_GROUND ? groundLevelSet(p, P[0]) : wall2groundLevelSet(p, P[0], P[1], P[2]);
}
void ref_ls_normal(int kind, const double* P, const double* x, double* n) {
    Vector3d p(x[0], x[1], x[2]);
    Vector3d g = kind == LS_GROUND ? DgroundLevelSet(p, P[0]) : Dwall2groundLevelSet(p, P[0], P[1], P[2]);
    n[0] = g[0]; n[1] = g[1]; n[2] = g[2];
}
// the SVD the reference's arithmetic runs on in this build (MiniEigen's JacobiSVD), for its own contract test
void ref_svd3(const double* F9, double* U9, double* s3, double* V9) {
    Matrix3d F; std::memcpy(F.data(), F9, 72);
    JacobiSVD<Matrix3d> svd(F, ComputeFullU | ComputeFullV);
    std::memcpy(U9, svd.matrixU().data(), 72); std::memcpy(V9, svd.matrixV().data(), 72); std::memcpy(s3, svd.singularValues().data(), 24);
}
void ref_svd2(const double* A4, double* U4, double* s2, double* V4) {
    Matrix2d A; std::memcpy(A.data(), A4, 32);
    JacobiSVD<Matrix2d> svd(A, ComputeFullU | ComputeFullV);
    std::memcpy(U4, svd.matrixU().data(), 32); std::memcpy(V4, svd.matrixV().data(), 32); std::memcpy(s2, svd.singularValues().data(), 16);
}
// the reference's scene factories (ParticleSystem.cpp:119-401); positions come from std::rand()/time-seeded engines, so
// only counts, masses and material constants can be compared
long ref_factory(int which, const double* a3, const double* b3, double r, double hgt, int n, double* x_out, double* mass_out, double* consts5) {
    Vector3d a(a3[0], a3[1], a3[2]), b(b3[0], b3[1], b3[2]);
    std::cout.setstate(std::ios::failbit);                            // SandBlock prints bmin.z()
    ParticleSystem ps = which == 0 ? ParticleSystem::SnowBall(a, r, n) : which == 1 ? ParticleSystem::SandBall(a, r, n)
                      : which == 2 ? ParticleSystem::SandBlock(a, b, r, n) : ParticleSystem::SandCylinder(a, r, hgt, n);
    std::cout.clear();
    out3(x_out, ps.positions); outv(mass_out, ps.masses);
    consts5[0] = ps.youngsModulus; consts5[1] = ps.poissonRatio; consts5[2] = ps.criticalCompression; consts5[3] = ps.criticalStretch; consts5[4] = ps.friction;
    return ps.positions.rows();
}
// LagrangianMesh::ObjMesh (LagrangianMesh.cpp:197-352): returns nv, nf and the derived per-element / per-vertex data
int ref_obj_mesh(const char* path, double density, double thickness, double E, double nu, double shear, double stiff, double angle_deg,
                 long* nv_nf, double* vx, int* faces, double* vm, double* vvol, double* em, double* evol, double* eD, double* consts3) {
    LagrangianMesh M = LagrangianMesh::ObjMesh(path, density, thickness, E, nu, shear, stiff, angle_deg);
    long nv = M.vertexPositions.rows(), nf = M.faces.rows();
    nv_nf[0] = nv; nv_nf[1] = nf;
    if (vx) {
        out3(vx, M.vertexPositions); std::memcpy(faces, M.faces.data(), sizeof(int) * 3 * nf);
        outv(vm, M.vertexMasses); outv(vvol, M.vertexVolumes); outv(em, M.elementMasses); outv(evol, M.elementVolumes);
        out3(eD, M.elementRestDirections_1()); out3(eD + 3 * nf, M.elementRestDirections_2()); out3(eD + 6 * nf, M.elementRestDirections_3());
        consts3[0] = M.mu; consts3[1] = M.lambda; consts3[2] = M.frictionCoeff;
    }
    return 0;
}

}  // extern "C"
