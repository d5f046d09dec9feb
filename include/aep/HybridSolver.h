/* HybridSolver.h -- the reference's solver surface (HybridSolver.h:27-95) on top of the C ABI of libaep_b200.so.
 *
 * Same public methods, same argument meaning: the caller owns the containers, the solver keeps non-owning pointers
 * (HybridSolver.h:30-34), `solve` blocks for the whole simulation and writes particle/particle_N.obj, mesh/mesh_N.obj
 * once per 1/60 s frame (HybridSolver.cpp:857-858, 991-1030).  What differs is where the loop body runs: the state is
 * uploaded once, every substep of HybridSolver.cpp:867-1032 executes on the B200 with the dt rule on the device, and
 * the host containers are refreshed under `mtx_` at frame boundaries (positions) and when solve returns (everything).
 * There is no CPU path: solve throws std::runtime_error if libaep_b200 cannot create a context.
 *
 * Extensions (no reference counterpart, all optional): setMaterialType (the reference hard-codes SAND at
 * HybridSolver.cpp:873,955,959), setAnalyticLevelSet (device-side primitives instead of sampling a std::function),
 * setOutputDirectory / setWriteFrames, config(), begin / advance / finish for callers that want to drive substeps,
 * saveCheckpoint / resume.
 */
#ifndef AEP_HOST_HYBRIDSOLVER_H
#define AEP_HOST_HYBRIDSOLVER_H
#include <functional>
#include <mutex>
#include <string>

#include "../aep_b200.h"
#include "EigenShim.h"
#include "LevelSet.h"

class ParticleSystem;
class RegularGrid;
class LagrangianMesh;
namespace igl { namespace viewer { class Viewer; } }

enum MaterialType { SNOW = 0, SAND };

class HybridSolver {
private:
    ParticleSystem* ps_;
    RegularGrid* rg_;
    LagrangianMesh* mesh_;
    igl::viewer::Viewer* viewer_;
    LevelSet phi_;
    DLevelSet dphi_;

    MaterialType material_ = SAND;
    int ls_kind_ = AEP_LS_NONE;
    double ls_params_[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    std::string out_dir_ = ".";
    bool write_frames_ = true;
    bool verbose_ = true;
    aep_config cfg_;
    aep_ctx* ctx_ = nullptr;
    long long substeps_ = 0;

    void uploadAll_();
    void downloadAll_();
    void createContext_(double CFL);
    void writeFrame_(int frameNo);
public:
    std::mutex mtx_;
    HybridSolver(ParticleSystem* ps = nullptr, RegularGrid* rg = nullptr);
    ~HybridSolver();
    HybridSolver(const HybridSolver&) = delete;
    HybridSolver& operator=(const HybridSolver&) = delete;

    void setParticleSystem(ParticleSystem* ps) { ps_ = ps; }
    void setRegularGrid(RegularGrid* rg) { rg_ = rg; }
    void setLagrangianMesh(LagrangianMesh* mesh) { mesh_ = mesh; }
    void setLevelSet(const LevelSet& phi, const DLevelSet& dphi) { phi_ = phi; dphi_ = dphi; ls_kind_ = AEP_LS_SAMPLED; }
    void solve(double CFL, double maxt, double alpha);

    void bindViewer(igl::viewer::Viewer* viewer);
    void updateViewer();

    /* ---- extensions ---- */
    void setMaterialType(MaterialType type) { material_ = type; }
    void setAnalyticLevelSet(int kind, const double* params, int nparams);
    void setOutputDirectory(const std::string& dir) { out_dir_ = dir; }
    void setWriteFrames(bool on) { write_frames_ = on; }
    void setVerbose(bool on) { verbose_ = on; }
    aep_config& config() { return cfg_; }            /* every literal of the hot path (SURVEY 5 "Config / flags") */
    void begin(double CFL);                          /* HybridSolver.cpp:829-860: upload, first P2G, volumes, initial dt */
    void advance(int substeps);                      /* n iterations of the loop body, no host synchronisation */
    int advanceFrames(int frames);                   /* until `frames` more 1/60 s frames completed; returns substeps */
    void fetchPositions();                           /* positions / vertexPositions / elementPositions -> containers */
    void finish();                                   /* all state -> containers (incl. grid mirrors), destroy the context */
    void clock(double* dt, double* t, int* frameNo, long long* substeps) const;
    aep_ctx* context() { return ctx_; }
    /* checkpoint / restart (the reference has none).  saveCheckpoint: between begin() and finish(), brings the
     * full state into the containers and writes it with the clock.  resume: instead of begin() -- fills the bound containers
     * (same particle / vertex / face counts as when saved) from the file, uploads them and continues with the saved clock.
     * File = int32 count, then per array: char name[32], int32 dtype (0 f64), int64 length, raw data.                      */
    void saveCheckpoint(const std::string& path);
    void resume(const std::string& path, double CFL);
    static void writeStateFile(const std::string& path, const ParticleSystem* ps, const LagrangianMesh* mesh, const double clock5[5]);
    static void readStateFile(const std::string& path, ParticleSystem* ps, LagrangianMesh* mesh, double clock5[5]);
};
#endif
