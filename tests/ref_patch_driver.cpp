// tests/ref_patch_driver.cpp -- drives INTEGRATION.md's binding B: the REFERENCE'S OWN container classes (compiled unmodified from
// /root/reference) with integration/HybridSolver_b200.cpp in place of HybridSolver.cpp.  Only the reference's public API is used.
//   ref_patch_driver <scene.bin> <out.bin> <maxt>      solver.solve(cfl, maxt, 0.95) in the working directory
// Scene / output files: the blob format of tests/host_driver.cpp (int32 count; per array char name[32], int32 dtype, int64 length, data).
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <functional>
#include <iostream>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "HybridSolver.h"
#include "LagrangianMesh.h"
#include "LevelSet.h"
#include "ParticleSystem.h"
#include "RegularGrid.h"

using namespace Eigen;

struct Blob { std::map<std::string, std::vector<double>> d; std::map<std::string, std::vector<int32_t>> i; };
static Blob read_blob(const char* path) {
    Blob b; std::ifstream f(path, std::ios::binary); if (!f) { std::fprintf(stderr, "cannot open %s\n", path); std::exit(2); }
    int32_t cnt = 0; f.read((char*)&cnt, 4);
    for (int a = 0; a < cnt; ++a) {
        char name[33] = {0}; int32_t dt; int64_t len; f.read(name, 32); f.read((char*)&dt, 4); f.read((char*)&len, 8);
        if (dt == 0) { auto& v = b.d[name]; v.resize((size_t)len); f.read((char*)v.data(), len * 8); }
        else { auto& v = b.i[name]; v.resize((size_t)len); f.read((char*)v.data(), len * 4); }
    }
    return b;
}
struct Writer {
    std::ofstream f; int32_t cnt = 0;
    explicit Writer(const char* p) : f(p, std::ios::binary) { f.write((char*)&cnt, 4); }
    void put(const char* name, const double* p, int64_t len) {
        char nm[32] = {0}; std::strncpy(nm, name, 31); int32_t dt = 0; f.write(nm, 32); f.write((char*)&dt, 4); f.write((char*)&len, 8); f.write((const char*)p, len * 8); ++cnt;
    }
    ~Writer() { f.seekp(0); f.write((char*)&cnt, 4); }
};
static MatrixX3d mat3(const std::vector<double>& s) { MatrixX3d m; m.resize((int)(s.size() / 3), 3); std::memcpy(m.data(), s.data(), s.size() * 8); return m; }
static VectorXd vec(const std::vector<double>& s) { VectorXd v((int)s.size()); std::memcpy(v.data(), s.data(), s.size() * 8); return v; }
static std::vector<Matrix3d> mats(const std::vector<double>& s) { std::vector<Matrix3d> M(s.size() / 9); for (size_t i = 0; i < M.size(); ++i) std::memcpy(M[i].data(), &s[9 * i], 72); return M; }

static HybridSolver solver;                 // a global, as in main.cpp:17 (the reference's ctor leaves mesh_ unset: zero-initialised storage)

int main(int argc, char** argv) {
    if (argc < 4) { std::fprintf(stderr, "usage: ref_patch_driver <scene.bin> <out.bin> <maxt>\n"); return 2; }
    const Blob b = read_blob(argv[1]);
    const std::vector<double>& g = b.d.at("grid"); const std::vector<int32_t>& res = b.i.at("res"); const std::vector<double>& sc = b.d.at("scalars");
    VectorXd mn(3), mx(3); for (int a = 0; a < 3; ++a) { mn[a] = g[a]; mx[a] = g[3 + a]; }
    RegularGrid rg(mn, mx, Vector3i(res[0], res[1], res[2]));
    solver.setRegularGrid(&rg);
    std::unique_ptr<ParticleSystem> ps; std::unique_ptr<LagrangianMesh> mesh; VectorXd fixed;
    if (b.d.count("x")) {
        const std::vector<double>& mat = b.d.at("material");
        const VectorXd m = vec(b.d.at("m")); VectorXd rho((int)m.size()); rho.setOnes(); MatrixX3d colors; colors.resize((int)m.size(), 3);
        ps.reset(new ParticleSystem(mat3(b.d.at("v")), mat3(b.d.at("x")), mats(b.d.at("FE")), mats(b.d.at("FP")), m, vec(b.d.at("vol")), rho, vec(b.d.at("q")),
                                    mat[0], mat[1], mat[2], mat[3], 0.2, colors));
        ps->affineMomenta_1 = mat3(b.d.at("B1")); ps->affineMomenta_2 = mat3(b.d.at("B2")); ps->affineMomenta_3 = mat3(b.d.at("B3"));
        solver.setParticleSystem(ps.get());
    }
    if (b.d.count("mesh_vx")) {
        const std::vector<int32_t>& fi = b.i.at("mesh_faces"); MatrixX3i F; F.resize((int)(fi.size() / 3), 3); std::memcpy(F.data(), fi.data(), fi.size() * 4);
        const std::vector<double>& mp = b.d.at("mesh_params");
        mesh.reset(new LagrangianMesh(mat3(b.d.at("mesh_vx")), F, mat3(b.d.at("mesh_vv")), mat3(b.d.at("mesh_ev")), vec(b.d.at("mesh_vm")), vec(b.d.at("mesh_vvol")),
                                      vec(b.d.at("mesh_em")), vec(b.d.at("mesh_evol")), mat3(b.d.at("mesh_d1")), mat3(b.d.at("mesh_d2")), mat3(b.d.at("mesh_d3")),
                                      mat3(b.d.at("mesh_D1")), mat3(b.d.at("mesh_D2")), mat3(b.d.at("mesh_D3")), mp[0], mp[1], mp[2], mp[3], mp[4]));
        fixed = b.d.count("mesh_fixed") ? vec(b.d.at("mesh_fixed")) : VectorXd((int)mesh->vertexPositions.rows());
        if (!b.d.count("mesh_fixed")) fixed.setZero();
        mesh->bindConstraints(&fixed);                                  // the reference's ctor leaves the pointer unset
        solver.setLagrangianMesh(mesh.get());
    }
    const int ls_kind = (int)sc[2]; const std::vector<double>& lp = b.d.at("ls_params");
    using namespace std::placeholders;                                  // the way main.cpp:86-91 binds its colliders
    if (ls_kind == 1) solver.setLevelSet(std::bind(groundLevelSet, _1, lp[0]), std::bind(DgroundLevelSet, _1, lp[0]));
    else if (ls_kind == 2) solver.setLevelSet(std::bind(wall2groundLevelSet, _1, lp[0], lp[1], lp[2]), std::bind(Dwall2groundLevelSet, _1, lp[0], lp[1], lp[2]));
    std::clog.setstate(std::ios::failbit);
    solver.solve(sc[1], std::atof(argv[3]), 0.95);
    Writer w(argv[2]);
    if (ps) {
        w.put("x", ps->positions.data(), ps->positions.size()); w.put("v", ps->velocities.data(), ps->velocities.size());
        w.put("B1", ps->affineMomenta_1.data(), ps->affineMomenta_1.size()); w.put("B2", ps->affineMomenta_2.data(), ps->affineMomenta_2.size());
        w.put("B3", ps->affineMomenta_3.data(), ps->affineMomenta_3.size());
        std::vector<double> FE(9 * ps->elasticDeformationGradients.size()), FP(FE.size());
        for (size_t i = 0; i < ps->elasticDeformationGradients.size(); ++i) { std::memcpy(&FE[9 * i], ps->elasticDeformationGradients[i].data(), 72); std::memcpy(&FP[9 * i], ps->plasticDeformationGradients[i].data(), 72); }
        w.put("FE", FE.data(), (int64_t)FE.size()); w.put("FP", FP.data(), (int64_t)FP.size());
        w.put("vol", ps->volumes.data(), ps->volumes.size()); w.put("q", ps->plasticAmount.data(), ps->plasticAmount.size());
    }
    if (mesh) {
        w.put("mesh_vx", mesh->vertexPositions.data(), mesh->vertexPositions.size()); w.put("mesh_vv", mesh->vertexVelocities.data(), mesh->vertexVelocities.size());
        w.put("mesh_ex", mesh->elementPositions.data(), mesh->elementPositions.size()); w.put("mesh_ev", mesh->elementVelocities.data(), mesh->elementVelocities.size());
        w.put("mesh_vB1", mesh->vertexAffineMomenta_1.data(), mesh->vertexAffineMomenta_1.size()); w.put("mesh_eB3", mesh->elementAffineMomenta_3.data(), mesh->elementAffineMomenta_3.size());
        w.put("mesh_d1", mesh->elementDirections_1.data(), mesh->elementDirections_1.size()); w.put("mesh_d3", mesh->elementDirections_3.data(), mesh->elementDirections_3.size());
    }
    w.put("grid_m", rg.masses.data(), rg.masses.size()); w.put("grid_v", rg.velocities.data(), rg.velocities.size()); w.put("grid_f", rg.forces.data(), rg.forces.size());
    return 0;
}
