// HybridSolver.cpp -- the reference's solver class (HybridSolver.h:27-95) as a thin host driver of libaep_b200.so.
//
// solve() = HybridSolver.cpp:827-1034 with the loop body on the GPU:
//   upload containers -> aep_init (weights, first P2G, volumes, initial dt; :829-860) -> repeat { aep_run_frames(1): all substeps
//   of one 1/60 s frame on the device, dt rule included (:867-987) -> positions back under mtx_ -> particle_N.obj / mesh_N.obj
//   (:991-1030) } while t <= maxt -> full state back into the containers.
// Pure host C++ (g++); everything numerical is behind the C ABI (include/aep_b200.h).
#include "../../../include/aep/HybridSolver.h"

#include <sys/stat.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <iostream>
#include <stdexcept>
#include <thread>
#include <vector>

#include "../../../include/aep/LagrangianMesh.h"
#include "../../../include/aep/ParticleSystem.h"
#include "../../../include/aep/RegularGrid.h"

using namespace Eigen;

namespace {
void ck(int rc, aep_ctx* ctx, const char* what) {
    if (rc == AEP_OK) return;
    const char* msg = aep_last_error(ctx);
    throw std::runtime_error(std::string("libaep_b200: ") + what + " failed (" + std::to_string(rc) + "): " + (msg ? msg : "?"));
}
// std::vector<Matrix3d> -> n x 9 doubles (column-major per item: already Eigen's layout, but copy so the shim and real Eigen
// both work without assuming sizeof(Matrix3d) == 72)
std::vector<double> flat9(const std::vector<Matrix3d>& M) {
    std::vector<double> o(M.size() * 9);
    for (size_t i = 0; i < M.size(); ++i) std::memcpy(&o[9 * i], M[i].data(), 9 * sizeof(double));
    return o;
}
void unflat9(const std::vector<double>& s, std::vector<Matrix3d>& M) {
    for (size_t i = 0; i < M.size(); ++i) std::memcpy(M[i].data(), &s[9 * i], 9 * sizeof(double));
}
std::vector<double> stack3(const MatrixX3d& a, const MatrixX3d& b, const MatrixX3d& c) {
    std::vector<double> o; o.reserve((size_t)(a.size() + b.size() + c.size()));
    o.insert(o.end(), a.data(), a.data() + a.size()); o.insert(o.end(), b.data(), b.data() + b.size()); o.insert(o.end(), c.data(), c.data() + c.size());
    return o;
}
void unstack3(const std::vector<double>& s, MatrixX3d& a, MatrixX3d& b, MatrixX3d& c) {
    const size_t n = (size_t)a.size();
    std::memcpy(a.data(), &s[0], n * 8); std::memcpy(b.data(), &s[n], n * 8); std::memcpy(c.data(), &s[2 * n], n * 8);
}
}  // namespace

HybridSolver::HybridSolver(ParticleSystem* ps, RegularGrid* rg) : ps_(ps), rg_(rg), mesh_(nullptr), viewer_(nullptr) {
    aep_default_config(&cfg_);
}
HybridSolver::~HybridSolver() { if (ctx_) aep_destroy(ctx_); }

void HybridSolver::bindViewer(igl::viewer::Viewer* viewer) {                   // HybridSolver.cpp:1036-1053
    viewer_ = viewer;
    if (ps_) ps_->bindViewer(viewer);
    if (rg_) rg_->bindViewer(viewer);
    if (mesh_) mesh_->bindViewer(viewer);
}
void HybridSolver::updateViewer() {                                             // HybridSolver.cpp:1055-1069: render-side hook, takes mtx_
    std::lock_guard<std::mutex> lk(mtx_);
    if (ps_) ps_->updateViewer();
    if (mesh_) mesh_->updateViewer();
}

void HybridSolver::setAnalyticLevelSet(int kind, const double* params, int nparams) {
    if (kind < AEP_LS_NONE || kind > AEP_LS_BOX) throw std::invalid_argument("setAnalyticLevelSet: unknown kind");
    ls_kind_ = kind;
    for (int i = 0; i < 8; ++i) ls_params_[i] = (params && i < nparams) ? params[i] : 0.0;
}

void HybridSolver::uploadAll_() {
    if (ps_) {
        ParticleSystem& p = *ps_;
        const std::vector<double> FE = flat9(p.elasticDeformationGradients), FP = flat9(p.plasticDeformationGradients);
        ck(aep_upload_particles(ctx_, (int64_t)p.masses.size(), p.positions.data(), p.velocities.data(), p.affineMomenta_1.data(),
                                p.affineMomenta_2.data(), p.affineMomenta_3.data(), FE.data(), FP.data(), p.masses.data(), p.volumes.data(),
                                p.plasticAmount.data(), p.youngsModulus, p.poissonRatio, p.criticalCompression, p.criticalStretch),
           ctx_, "aep_upload_particles");
    }
    if (mesh_) {
        LagrangianMesh& m = *mesh_;
        const std::vector<double> vB = stack3(m.vertexAffineMomenta_1, m.vertexAffineMomenta_2, m.vertexAffineMomenta_3);
        const std::vector<double> eB = stack3(m.elementAffineMomenta_1, m.elementAffineMomenta_2, m.elementAffineMomenta_3);
        const std::vector<double> ed = stack3(m.elementDirections_1, m.elementDirections_2, m.elementDirections_3);
        const std::vector<double> eD = stack3(m.elementRestDirections_1(), m.elementRestDirections_2(), m.elementRestDirections_3());
        std::vector<int32_t> faces((size_t)m.faces.size());
        for (size_t i = 0; i < faces.size(); ++i) faces[i] = (int32_t)m.faces.data()[i];
        const VectorXd* fixed = m.constraints();
        ck(aep_upload_mesh(ctx_, (int64_t)m.vertexPositions.rows(), (int64_t)m.faces.rows(), m.vertexPositions.data(), m.vertexVelocities.data(),
                           m.vertexMasses.data(), m.vertexVolumes.data(), vB.data(), faces.data(), m.elementVelocities.data(), m.elementMasses.data(),
                           m.elementVolumes.data(), eB.data(), ed.data(), eD.data(), fixed ? fixed->data() : nullptr, m.mu, m.lambda,
                           m.shearStiffness, m.stiffness, m.frictionCoeff),
           ctx_, "aep_upload_mesh");
    }
    if (ls_kind_ == AEP_LS_SAMPLED && phi_ && dphi_) {
        // HybridSolver.cpp:473-482 evaluates phi / grad phi at grid nodes only and colliders are static (:484): sample once.
        const Vector3i res = rg_->resolution(); const Vector3d mn = rg_->minBound(), h = rg_->h();
        const size_t Ng = (size_t)rg_->gridNumber();
        std::vector<uint8_t> inside(Ng, 0); std::vector<double> normal(3 * Ng, 0.0);
        const unsigned nt = std::max(1u, std::min(std::thread::hardware_concurrency(), 32u));
        std::vector<std::thread> th;
        for (unsigned t = 0; t < nt; ++t)
            th.emplace_back([&, t]() {
                for (int k = (int)t; k < res[2]; k += (int)nt) for (int j = 0; j < res[1]; ++j) for (int i = 0; i < res[0]; ++i) {
                    const Vector3d x(mn[0] + i * h[0], mn[1] + j * h[1], mn[2] + k * h[2]);
                    if (phi_(x) <= 0.0) {
                        const size_t id = ((size_t)k * res[1] + j) * res[0] + i;
                        const Vector3d n = dphi_(x);
                        inside[id] = 1; normal[id] = n[0]; normal[Ng + id] = n[1]; normal[2 * Ng + id] = n[2];
                    }
                }
            });
        for (auto& x : th) x.join();
        ck(aep_set_levelset_samples(ctx_, inside.data(), normal.data()), ctx_, "aep_set_levelset_samples");
    } else if (ls_kind_ != AEP_LS_NONE && ls_kind_ != AEP_LS_SAMPLED) {
        ck(aep_set_levelset_analytic(ctx_, ls_kind_, ls_params_), ctx_, "aep_set_levelset_analytic");
    }
}

void HybridSolver::createContext_(double CFL) {
    if (!rg_) throw std::invalid_argument("HybridSolver: setRegularGrid first");
    if (!ps_ && !mesh_) throw std::invalid_argument("HybridSolver: nothing to simulate (no ParticleSystem, no LagrangianMesh)");
    if (ctx_) { aep_destroy(ctx_); ctx_ = nullptr; }
    cfg_.material = material_ == SNOW ? AEP_SNOW : AEP_SAND;
    cfg_.cfl = CFL;
    for (int a = 0; a < 3; ++a) { cfg_.grid_min[a] = rg_->minBound()[a]; cfg_.grid_max[a] = rg_->maxBound()[a]; cfg_.res[a] = rg_->resolution()[a]; }
    ck(aep_create(&ctx_, &cfg_), nullptr, "aep_create");
    uploadAll_();
}
void HybridSolver::begin(double CFL) {
    createContext_(CFL);
    ck(aep_init(ctx_), ctx_, "aep_init");
    substeps_ = 0;
}
void HybridSolver::advance(int substeps) {
    if (!ctx_) throw std::logic_error("HybridSolver::advance before begin");
    ck(aep_run(ctx_, substeps), ctx_, "aep_run"); substeps_ += substeps;
}
int HybridSolver::advanceFrames(int frames) {
    if (!ctx_) throw std::logic_error("HybridSolver::advanceFrames before begin");
    int64_t done = 0;
    ck(aep_run_frames(ctx_, frames, 1 << 30, &done), ctx_, "aep_run_frames"); substeps_ += done;
    return (int)done;
}
void HybridSolver::clock(double* dt, double* t, int* frameNo, long long* substeps) const {
    int32_t fr = 0; int64_t ss = 0;
    ck(aep_get_clock(ctx_, dt, t, nullptr, &fr, &ss, nullptr, nullptr), ctx_, "aep_get_clock");
    if (frameNo) *frameNo = fr;
    if (substeps) *substeps = ss;
}

void HybridSolver::fetchPositions() {
    std::lock_guard<std::mutex> lk(mtx_);                                       // the render thread reads these (main.cpp:19-24)
    if (ps_) ck(aep_download_particles(ctx_, ps_->positions.data(), nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr), ctx_, "aep_download_particles");
    if (mesh_) ck(aep_download_mesh(ctx_, mesh_->vertexPositions.data(), nullptr, nullptr, mesh_->elementPositions.data(), nullptr, nullptr, nullptr), ctx_, "aep_download_mesh");
}

void HybridSolver::downloadAll_() {
    std::lock_guard<std::mutex> lk(mtx_);
    if (ps_) {
        ParticleSystem& p = *ps_;
        std::vector<double> FE(9 * p.elasticDeformationGradients.size()), FP(FE.size());
        ck(aep_download_particles(ctx_, p.positions.data(), p.velocities.data(), p.affineMomenta_1.data(), p.affineMomenta_2.data(),
                                  p.affineMomenta_3.data(), FE.data(), FP.data(), p.volumes.data(), p.plasticAmount.data()), ctx_, "aep_download_particles");
        unflat9(FE, p.elasticDeformationGradients); unflat9(FP, p.plasticDeformationGradients);
        for (std::ptrdiff_t i = 0; i < p.masses.size(); ++i) p.densities[i] = p.volumes[i] > 0 ? p.masses[i] / p.volumes[i] : 0.0;   // HybridSolver.cpp:246-248
    }
    if (mesh_) {
        LagrangianMesh& m = *mesh_;
        const size_t nv = (size_t)m.vertexPositions.rows(), nf = (size_t)m.faces.rows();
        std::vector<double> vB(9 * nv), eB(9 * nf), ed(9 * nf);
        ck(aep_download_mesh(ctx_, m.vertexPositions.data(), m.vertexVelocities.data(), vB.data(), m.elementPositions.data(),
                             m.elementVelocities.data(), eB.data(), ed.data()), ctx_, "aep_download_mesh");
        unstack3(vB, m.vertexAffineMomenta_1, m.vertexAffineMomenta_2, m.vertexAffineMomenta_3);
        unstack3(eB, m.elementAffineMomenta_1, m.elementAffineMomenta_2, m.elementAffineMomenta_3);
        unstack3(ed, m.elementDirections_1, m.elementDirections_2, m.elementDirections_3);
    }
    rg_->allocateHostMirrors();
    ck(aep_download_grid(ctx_, rg_->masses.data(), rg_->velocities.data(), rg_->forces.data(), nullptr), ctx_, "aep_download_grid");
}

void HybridSolver::finish() {
    if (!ctx_) return;
    downloadAll_();
    aep_destroy(ctx_); ctx_ = nullptr;
}

// ---- checkpoint / restart ------------------------------------------------------------------------------------------------
namespace {
struct BlobWriter {
    FILE* f; int32_t cnt = 0;
    explicit BlobWriter(const std::string& p) : f(std::fopen(p.c_str(), "wb")) {
        if (!f) throw std::runtime_error("cannot write " + p);
        std::fwrite(&cnt, 4, 1, f);
    }
    void put(const char* name, const double* p, int64_t len) {
        char nm[32] = {0}; std::strncpy(nm, name, 31); const int32_t dt = 0;
        std::fwrite(nm, 32, 1, f); std::fwrite(&dt, 4, 1, f); std::fwrite(&len, 8, 1, f); if (len) std::fwrite(p, 8, (size_t)len, f); ++cnt;
    }
    void put9(const char* name, const std::vector<Matrix3d>& M) { const std::vector<double> v = flat9(M); put(name, v.data(), (int64_t)v.size()); }
    ~BlobWriter() { std::fseek(f, 0, SEEK_SET); std::fwrite(&cnt, 4, 1, f); std::fclose(f); }
};
// reads array `name` of exactly `len` doubles into dst
struct BlobReader {
    std::vector<std::pair<std::string, std::vector<double>>> a;
    explicit BlobReader(const std::string& p) {
        FILE* f = std::fopen(p.c_str(), "rb");
        if (!f) throw std::runtime_error("cannot open checkpoint " + p);
        int32_t cnt = 0; bool ok = std::fread(&cnt, 4, 1, f) == 1 && cnt >= 0 && cnt < 4096;
        for (int i = 0; ok && i < cnt; ++i) {
            char nm[33] = {0}; int32_t dt = 0; int64_t len = 0;
            ok = std::fread(nm, 32, 1, f) == 1 && std::fread(&dt, 4, 1, f) == 1 && std::fread(&len, 8, 1, f) == 1 && dt == 0 && len >= 0;
            if (!ok) break;
            std::vector<double> v((size_t)len);
            ok = len == 0 || std::fread(v.data(), 8, (size_t)len, f) == (size_t)len;
            a.emplace_back(nm, std::move(v));
        }
        std::fclose(f);
        if (!ok) throw std::runtime_error("malformed checkpoint " + p);
    }
    const std::vector<double>& get(const char* name, size_t len) const {
        for (const auto& e : a) if (e.first == name) {
            if (e.second.size() != len) throw std::runtime_error(std::string("checkpoint array ") + name + " has " + std::to_string(e.second.size()) + " values, the bound container needs " + std::to_string(len));
            return e.second;
        }
        throw std::runtime_error(std::string("checkpoint has no array ") + name);
    }
    void into(const char* name, double* dst, size_t len) const { const std::vector<double>& v = get(name, len); if (len) std::memcpy(dst, v.data(), len * 8); }
};
}  // namespace

void HybridSolver::writeStateFile(const std::string& path, const ParticleSystem* ps, const LagrangianMesh* mesh, const double clock5[5]) {
    // written next to the target and renamed over it: a crash in the middle of the write leaves the previous restart file intact
    const std::string tmp = path + ".tmp";
    {
    BlobWriter w(tmp);
    const double version = 1.0;
    w.put("aep_checkpoint", &version, 1); w.put("clock", clock5, 5);                      // dt, t, inner_t, frame_no, substeps
    if (ps) {
        w.put("p_x", ps->positions.data(), ps->positions.size()); w.put("p_v", ps->velocities.data(), ps->velocities.size());
        w.put("p_B1", ps->affineMomenta_1.data(), ps->affineMomenta_1.size()); w.put("p_B2", ps->affineMomenta_2.data(), ps->affineMomenta_2.size());
        w.put("p_B3", ps->affineMomenta_3.data(), ps->affineMomenta_3.size());
        w.put9("p_FE", ps->elasticDeformationGradients); w.put9("p_FP", ps->plasticDeformationGradients);
        w.put("p_m", ps->masses.data(), ps->masses.size()); w.put("p_vol", ps->volumes.data(), ps->volumes.size());
        w.put("p_q", ps->plasticAmount.data(), ps->plasticAmount.size());
    }
    if (mesh) {
        const LagrangianMesh& m = *mesh;
        w.put("m_vx", m.vertexPositions.data(), m.vertexPositions.size()); w.put("m_vv", m.vertexVelocities.data(), m.vertexVelocities.size());
        w.put("m_ex", m.elementPositions.data(), m.elementPositions.size()); w.put("m_ev", m.elementVelocities.data(), m.elementVelocities.size());
        const std::vector<double> vB = stack3(m.vertexAffineMomenta_1, m.vertexAffineMomenta_2, m.vertexAffineMomenta_3);
        const std::vector<double> eB = stack3(m.elementAffineMomenta_1, m.elementAffineMomenta_2, m.elementAffineMomenta_3);
        const std::vector<double> ed = stack3(m.elementDirections_1, m.elementDirections_2, m.elementDirections_3);
        w.put("m_vB", vB.data(), (int64_t)vB.size()); w.put("m_eB", eB.data(), (int64_t)eB.size()); w.put("m_ed", ed.data(), (int64_t)ed.size());
    }
    }   // BlobWriter's destructor finalises and closes the file
    if (std::rename(tmp.c_str(), path.c_str()) != 0) throw std::runtime_error("cannot move " + tmp + " over " + path);
}

void HybridSolver::readStateFile(const std::string& path, ParticleSystem* ps, LagrangianMesh* mesh, double clock5[5]) {
    const BlobReader r(path);
    if (r.get("aep_checkpoint", 1)[0] != 1.0) throw std::runtime_error("unsupported checkpoint version in " + path);
    r.into("clock", clock5, 5);
    if (ps) {
        ParticleSystem& p = *ps; const size_t n = (size_t)p.masses.size();
        r.into("p_x", p.positions.data(), 3 * n); r.into("p_v", p.velocities.data(), 3 * n);
        r.into("p_B1", p.affineMomenta_1.data(), 3 * n); r.into("p_B2", p.affineMomenta_2.data(), 3 * n); r.into("p_B3", p.affineMomenta_3.data(), 3 * n);
        unflat9(r.get("p_FE", 9 * n), p.elasticDeformationGradients); unflat9(r.get("p_FP", 9 * n), p.plasticDeformationGradients);
        r.into("p_m", p.masses.data(), n); r.into("p_vol", p.volumes.data(), n); r.into("p_q", p.plasticAmount.data(), n);
    }
    if (mesh) {
        LagrangianMesh& m = *mesh; const size_t nv = (size_t)m.vertexPositions.rows(), nf = (size_t)m.faces.rows();
        r.into("m_vx", m.vertexPositions.data(), 3 * nv); r.into("m_vv", m.vertexVelocities.data(), 3 * nv);
        r.into("m_ex", m.elementPositions.data(), 3 * nf); r.into("m_ev", m.elementVelocities.data(), 3 * nf);
        unstack3(r.get("m_vB", 9 * nv), m.vertexAffineMomenta_1, m.vertexAffineMomenta_2, m.vertexAffineMomenta_3);
        unstack3(r.get("m_eB", 9 * nf), m.elementAffineMomenta_1, m.elementAffineMomenta_2, m.elementAffineMomenta_3);
        unstack3(r.get("m_ed", 9 * nf), m.elementDirections_1, m.elementDirections_2, m.elementDirections_3);
    }
}

void HybridSolver::saveCheckpoint(const std::string& path) {
    if (!ctx_) throw std::logic_error("HybridSolver::saveCheckpoint outside begin()/finish()");
    downloadAll_();
    double c[5]; int32_t fr = 0; int64_t ss = 0;
    ck(aep_get_clock(ctx_, &c[0], &c[1], &c[2], &fr, &ss, nullptr, nullptr), ctx_, "aep_get_clock");
    c[3] = fr; c[4] = (double)ss;
    writeStateFile(path, ps_, mesh_, c);
}

void HybridSolver::resume(const std::string& path, double CFL) {
    double c[5];
    readStateFile(path, ps_, mesh_, c);                          // throws before any GPU work if the file does not fit the containers
    createContext_(CFL);
    ck(aep_resume(ctx_), ctx_, "aep_resume");
    ck(aep_set_clock(ctx_, c[0], c[1], c[2], (int32_t)c[3], (int64_t)c[4]), ctx_, "aep_set_clock");
    substeps_ = (long long)c[4];
}

void HybridSolver::writeFrame_(int frameNo) {                                   // HybridSolver.cpp:991-1030: "v x y z" lines, faces 1-based
    char name[512];
    if (ps_) {
        std::snprintf(name, sizeof name, "%s/particle/particle_%d.obj", out_dir_.c_str(), frameNo);
        if (FILE* f = std::fopen(name, "w")) {
            for (std::ptrdiff_t p = 0; p < ps_->positions.rows(); ++p)
                std::fprintf(f, "v %g %g %g\n", ps_->positions(p, 0), ps_->positions(p, 1), ps_->positions(p, 2));
            std::fclose(f);
        }
    }
    if (mesh_) {
        std::snprintf(name, sizeof name, "%s/mesh/mesh_%d.obj", out_dir_.c_str(), frameNo);
        if (FILE* f = std::fopen(name, "w")) {
            for (std::ptrdiff_t p = 0; p < mesh_->vertexPositions.rows(); ++p)
                std::fprintf(f, "v %g %g %g\n", mesh_->vertexPositions(p, 0), mesh_->vertexPositions(p, 1), mesh_->vertexPositions(p, 2));
            for (std::ptrdiff_t t = 0; t < mesh_->faces.rows(); ++t)
                std::fprintf(f, "f %d %d %d\n", mesh_->faces(t, 0) + 1, mesh_->faces(t, 1) + 1, mesh_->faces(t, 2) + 1);
            std::fclose(f);
        }
    }
}

// alpha (FLIP blend) is accepted and ignored, exactly as in the reference (HybridSolver.cpp:739: pure APIC/PIC).
// Frame pipeline: the positions of frame N are snapshotted on the device the moment the frame completes and travel to page-locked
// host memory on a second stream while frame N+1 computes; particle_N.obj is written by a helper thread while frame N+2 computes.
// aep_run_frames fails loudly (AEP_ERR_STATE -> exception) when particles left the grid or became NaN: no plausible-looking
// frames of a blown-up simulation are written.
void HybridSolver::solve(double CFL, double maxt, double /*alpha*/) {
    begin(CFL);
    if (write_frames_) {                                                        // HybridSolver.cpp:857-858 (system("mkdir ..."))
        ::mkdir(out_dir_.c_str(), 0777); ::mkdir((out_dir_ + "/particle").c_str(), 0777); ::mkdir((out_dir_ + "/mesh").c_str(), 0777);
    }
    const size_t np = ps_ ? (size_t)ps_->masses.size() : 0;
    float* pin[2] = { nullptr, nullptr };
    if (np) for (int b = 0; b < 2; ++b) {
        pin[b] = static_cast<float*>(aep_host_alloc((int64_t)(np * 3 * sizeof(float))));
        if (!pin[b]) { if (pin[0]) aep_host_free(pin[0]); throw std::runtime_error("HybridSolver::solve: cannot allocate page-locked frame buffers"); }
    }
    std::thread writer; std::exception_ptr werr;
    auto finish_frame = [&](int frame, int buf, double tt, int nsub) {          // frame's positions have landed in pin[buf]
        if (np) {
            std::lock_guard<std::mutex> lk(mtx_);                               // the render thread reads positions (main.cpp:19-24)
            for (size_t p = 0; p < np; ++p) for (int a = 0; a < 3; ++a) ps_->positions((std::ptrdiff_t)p, a) = pin[buf][3 * p + a];
        }
        if (write_frames_) writeFrame_(frame);
        if (verbose_) std::clog << "frame " << frame << ": time " << tt << ", " << nsub << " substeps" << std::endl;
    };
    struct Pending { int frame = -1, buf = 0, n = 0; double t = 0.0; } pend;
    auto flush_pending = [&](bool in_background) {                              // the copy of pend.frame has had a whole frame of compute to land
        if (pend.frame < 0) return;
        ck(aep_frame_positions_wait(ctx_), ctx_, "aep_frame_positions_wait");
        if (writer.joinable()) writer.join();                                   // the frame before it is on disk; its buffer is free again
        if (werr) std::rethrow_exception(werr);
        const Pending p = pend; pend.frame = -1;
        if (in_background) writer = std::thread([&, p]() { try { finish_frame(p.frame, p.buf, p.t, p.n); } catch (...) { werr = std::current_exception(); } });
        else finish_frame(p.frame, p.buf, p.t, p.n);
    };
    double t = 0.0; int frameNo = 0;
    try {
        while (t <= maxt) {                                                     // HybridSolver.cpp:867; t advances by 1/60 per finished frame (:883)
            const int n = advanceFrames(1);
            t += 1.0 / 60.0;
            if (np && !mesh_) {
                flush_pending(true);
                ck(aep_frame_positions_begin(ctx_, pin[frameNo & 1]), ctx_, "aep_frame_positions_begin");
                pend.frame = frameNo; pend.buf = frameNo & 1; pend.n = n; pend.t = t;
            } else {                                                            // with a mesh (small scenes): frame by frame
                if (mesh_) {
                    std::lock_guard<std::mutex> lk(mtx_);
                    ck(aep_download_mesh(ctx_, mesh_->vertexPositions.data(), nullptr, nullptr, mesh_->elementPositions.data(), nullptr, nullptr, nullptr), ctx_, "aep_download_mesh");
                }
                if (np) {
                    ck(aep_frame_positions_begin(ctx_, pin[0]), ctx_, "aep_frame_positions_begin");
                    ck(aep_frame_positions_wait(ctx_), ctx_, "aep_frame_positions_wait");
                }
                finish_frame(frameNo, 0, t, n);
            }
            ++frameNo;
        }
        flush_pending(false);
        if (writer.joinable()) writer.join();
        if (werr) std::rethrow_exception(werr);
    } catch (...) {
        if (writer.joinable()) writer.join();
        for (int b = 0; b < 2; ++b) if (pin[b]) aep_host_free(pin[b]);
        throw;
    }
    for (int b = 0; b < 2; ++b) if (pin[b]) aep_host_free(pin[b]);
    finish();
}
