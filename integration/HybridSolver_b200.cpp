// integration/HybridSolver_b200.cpp -- the reference-side binding (option B), as a complete translation unit.
//
// Drop this file into the reference's source directory IN PLACE OF HybridSolver.cpp and link libaep_b200.so: it defines the three
// public methods of the reference's OWN HybridSolver class (HybridSolver.h:27-95) that have a body in HybridSolver.cpp --
//     solve          HybridSolver.cpp:827-1034   -> the loop body runs on the B200 through the C ABI (include/aep_b200.h)
//     bindViewer     HybridSolver.cpp:1036-1053  -> unchanged in meaning
//     updateViewer   HybridSolver.cpp:1055-1069  -> unchanged in meaning
// and nothing else.  ParticleSystem, RegularGrid, LagrangianMesh, LevelSet, geometry, interpolation and main.cpp stay the
// reference's, byte for byte; the private stage methods (evaluateInterpolationWeights_ ... updateAffineMomenta_) and the twelve
// SparseMatrix members of the class are simply never used.  It compiles against the reference's headers and whatever Eigen they
// resolve to (tests/test_zzz_reference_main.py builds it against the reference's unmodified sources).
//
// Behaviour kept: same arguments (alpha is ignored by the reference too, HS:739), frames particle/particle_N.obj and
// mesh/mesh_N.obj in the working directory with the reference's line format, `while (t <= maxt)` frame counting, containers
// updated in place, positions refreshed under mtx_ once per frame.  The material is SAND as hard-wired at HS:873,955,959
// unless the environment says AEP_MATERIAL=snow.  Errors end the process like the reference's do (cerr + exit).
#include "HybridSolver.h"
#include "ParticleSystem.h"
#include "RegularGrid.h"
#include "LagrangianMesh.h"

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <vector>
#include <sys/stat.h>
#include <igl/viewer/Viewer.h>

#include "aep_b200.h"

namespace {
void ck(int rc, aep_ctx* c, const char* what) {
    if (rc == AEP_OK) return;
    const char* msg = aep_last_error(c);
    std::cerr << "[ERROR] libaep_b200: " << what << " failed (" << rc << "): " << (msg ? msg : "?") << std::endl;
    std::exit(1);
}
// three N x 3 column-major matrices one after the other (how the ABI takes the rows of the APIC matrices and the directions)
std::vector<double> stack3(const Eigen::MatrixX3d& a, const Eigen::MatrixX3d& b, const Eigen::MatrixX3d& c) {
    std::vector<double> o; o.reserve(static_cast<size_t>(a.size() + b.size() + c.size()));
    o.insert(o.end(), a.data(), a.data() + a.size()); o.insert(o.end(), b.data(), b.data() + b.size()); o.insert(o.end(), c.data(), c.data() + c.size());
    return o;
}
void unstack3(const std::vector<double>& s, Eigen::MatrixX3d& a, Eigen::MatrixX3d& b, Eigen::MatrixX3d& c) {
    const size_t n = static_cast<size_t>(a.size());
    std::memcpy(a.data(), &s[0], n * 8); std::memcpy(b.data(), &s[n], n * 8); std::memcpy(c.data(), &s[2 * n], n * 8);
}
static_assert(sizeof(Eigen::Matrix3d) == 9 * sizeof(double), "std::vector<Matrix3d> must be 9 contiguous doubles per item");
}  // namespace

void HybridSolver::solve(double CFL, double maxt, double /*alpha*/)
{
    aep_config cfg; aep_default_config(&cfg);                     // the literals of HS:457,465,267,641-644,860,880
    const char* mat = std::getenv("AEP_MATERIAL");
    cfg.material = (mat && std::strcmp(mat, "snow") == 0) ? AEP_SNOW : AEP_SAND;
    cfg.cfl = CFL;
    for (int a = 0; a < 3; ++a) { cfg.grid_min[a] = rg_->minBound()[a]; cfg.grid_max[a] = rg_->maxBound()[a]; cfg.res[a] = rg_->resolution()[a]; }
    aep_ctx* ctx = nullptr;
    ck(aep_create(&ctx, &cfg), nullptr, "aep_create");

    if (ps_ != nullptr) {                                         // the Eigen members already have the ABI's layouts: pointers pass through
        ParticleSystem& p = *ps_;
        ck(aep_upload_particles(ctx, p.masses.size(), p.positions.data(), p.velocities.data(), p.affineMomenta_1.data(), p.affineMomenta_2.data(),
                                p.affineMomenta_3.data(), p.elasticDeformationGradients[0].data(), p.plasticDeformationGradients[0].data(),
                                p.masses.data(), p.volumes.data(), p.plasticAmount.data(), p.youngsModulus, p.poissonRatio,
                                p.criticalCompression, p.criticalStretch), ctx, "aep_upload_particles");
    }
    if (mesh_ != nullptr) {
        LagrangianMesh& m = *mesh_;
        const int64_t nv = m.vertexPositions.rows(), nf = m.faces.rows();
        const std::vector<double> vB = stack3(m.vertexAffineMomenta_1, m.vertexAffineMomenta_2, m.vertexAffineMomenta_3);
        const std::vector<double> eB = stack3(m.elementAffineMomenta_1, m.elementAffineMomenta_2, m.elementAffineMomenta_3);
        const std::vector<double> ed = stack3(m.elementDirections_1, m.elementDirections_2, m.elementDirections_3);
        const std::vector<double> eD = stack3(m.elementRestDirections_1(), m.elementRestDirections_2(), m.elementRestDirections_3());
        std::vector<int32_t> faces(static_cast<size_t>(3 * nf));
        for (size_t i = 0; i < faces.size(); ++i) faces[i] = m.faces.data()[i];
        std::vector<double> fixed(static_cast<size_t>(nv));
        for (int64_t v = 0; v < nv; ++v) fixed[static_cast<size_t>(v)] = m.vertexIsFixed(static_cast<int>(v)) ? 1.0 : 0.0;   // HS:521
        ck(aep_upload_mesh(ctx, nv, nf, m.vertexPositions.data(), m.vertexVelocities.data(), m.vertexMasses.data(), m.vertexVolumes.data(), vB.data(),
                           faces.data(), m.elementVelocities.data(), m.elementMasses.data(), m.elementVolumes.data(), eB.data(), ed.data(), eD.data(),
                           fixed.data(), m.mu, m.lambda, m.shearStiffness, m.stiffness, m.frictionCoeff), ctx, "aep_upload_mesh");
    }
    if (phi_ && dphi_) {                                          // sampled once at the nodes: the loop only evaluates them there, for static colliders (HS:473-484)
        const int Ng = rg_->gridNumber();
        std::vector<uint8_t> inside(static_cast<size_t>(Ng), 0); std::vector<double> normal(static_cast<size_t>(3) * Ng, 0.0);
        for (int k = 0; k < rg_->resolution()[2]; ++k) for (int j = 0; j < rg_->resolution()[1]; ++j) for (int i = 0; i < rg_->resolution()[0]; ++i) {
            const Eigen::Vector3d x(rg_->minBound()[0] + i * rg_->h()[0], rg_->minBound()[1] + j * rg_->h()[1], rg_->minBound()[2] + k * rg_->h()[2]);
            if (phi_(x) <= 0.0) {
                const int id = rg_->toIndex(i, j, k); const Eigen::Vector3d n = dphi_(x);
                inside[id] = 1; normal[id] = n[0]; normal[static_cast<size_t>(Ng) + id] = n[1]; normal[2 * static_cast<size_t>(Ng) + id] = n[2];
            }
        }
        ck(aep_set_levelset_samples(ctx, inside.data(), normal.data()), ctx, "aep_set_levelset_samples");
    }

    ck(aep_init(ctx), ctx, "aep_init");                           // HS:829-860: weights, first P2G, volumes, initial dt
    ::mkdir("particle", 0777); ::mkdir("mesh", 0777);             // HS:857-858
    double t = 0.0; int frameNo = 0;
    while (t <= maxt) {                                           // HS:867
        int64_t substeps = 0;
        ck(aep_run_frames(ctx, 1, 1 << 30, &substeps), ctx, "aep_run_frames");      // every substep of one frame on the GPU, dt rule and frame clipping included
        t += 1.0 / 60.0;                                          // HS:883
        mtx_.lock();                                              // the render thread reads these (HS:1059-1068)
        if (ps_ != nullptr) ck(aep_download_particles(ctx, ps_->positions.data(), 0, 0, 0, 0, 0, 0, 0, 0), ctx, "aep_download_particles");
        if (mesh_ != nullptr) ck(aep_download_mesh(ctx, mesh_->vertexPositions.data(), 0, 0, mesh_->elementPositions.data(), 0, 0, 0), ctx, "aep_download_mesh");
        mtx_.unlock();
        char name[64];
        if (ps_ != nullptr) {                                     // HS:997-1007
            std::snprintf(name, sizeof name, "particle/particle_%d.obj", frameNo);
            if (FILE* f = std::fopen(name, "w")) {
                for (int p = 0; p < ps_->positions.rows(); ++p) std::fprintf(f, "v %g %g %g\n", ps_->positions(p, 0), ps_->positions(p, 1), ps_->positions(p, 2));
                std::fclose(f);
            }
        }
        if (mesh_ != nullptr) {                                   // HS:1009-1025
            std::snprintf(name, sizeof name, "mesh/mesh_%d.obj", frameNo);
            if (FILE* f = std::fopen(name, "w")) {
                for (int p = 0; p < mesh_->vertexPositions.rows(); ++p)
                    std::fprintf(f, "v %g %g %g\n", mesh_->vertexPositions(p, 0), mesh_->vertexPositions(p, 1), mesh_->vertexPositions(p, 2));
                for (int q = 0; q < mesh_->faces.rows(); ++q) std::fprintf(f, "f %d %d %d\n", mesh_->faces(q, 0) + 1, mesh_->faces(q, 1) + 1, mesh_->faces(q, 2) + 1);
                std::fclose(f);
            }
        }
        std::clog << "frame " << frameNo << ": time " << t << ", " << substeps << " substeps" << std::endl;
        ++frameNo;
    }

    // everything back into the reference's containers
    mtx_.lock();
    if (ps_ != nullptr) {
        ParticleSystem& p = *ps_;
        ck(aep_download_particles(ctx, p.positions.data(), p.velocities.data(), p.affineMomenta_1.data(), p.affineMomenta_2.data(), p.affineMomenta_3.data(),
                                  p.elasticDeformationGradients[0].data(), p.plasticDeformationGradients[0].data(), p.volumes.data(), p.plasticAmount.data()),
           ctx, "aep_download_particles");
    }
    if (mesh_ != nullptr) {
        LagrangianMesh& m = *mesh_;
        const size_t nv = static_cast<size_t>(m.vertexPositions.rows()), nf = static_cast<size_t>(m.faces.rows());
        std::vector<double> vB(9 * nv), eB(9 * nf), ed(9 * nf);
        ck(aep_download_mesh(ctx, m.vertexPositions.data(), m.vertexVelocities.data(), vB.data(), m.elementPositions.data(), m.elementVelocities.data(),
                             eB.data(), ed.data()), ctx, "aep_download_mesh");
        unstack3(vB, m.vertexAffineMomenta_1, m.vertexAffineMomenta_2, m.vertexAffineMomenta_3);
        unstack3(eB, m.elementAffineMomenta_1, m.elementAffineMomenta_2, m.elementAffineMomenta_3);
        unstack3(ed, m.elementDirections_1, m.elementDirections_2, m.elementDirections_3);
    }
    ck(aep_download_grid(ctx, rg_->masses.data(), rg_->velocities.data(), rg_->forces.data(), nullptr), ctx, "aep_download_grid");
    mtx_.unlock();
    aep_destroy(ctx);
}

void HybridSolver::bindViewer(igl::viewer::Viewer* viewer)
{
    viewer_ = viewer;
    if (ps_ != nullptr) ps_->bindViewer(viewer);
    if (rg_ != nullptr) rg_->bindViewer(viewer);
    if (mesh_ != nullptr) mesh_->bindViewer(viewer);
}

void HybridSolver::updateViewer()
{
    viewer_->data.clear();
    mtx_.lock();
    if (ps_ != nullptr) ps_->updateViewer();
    if (mesh_ != nullptr) mesh_->updateViewer();
    mtx_.unlock();
}
