// tests/host_driver.cpp -- exercises the reference-shaped C++ host classes (include/aep/*.h, libaep_host.so).
//   host_driver unit <tmpdir>                      host-only checks, no GPU needed (containers, level sets, OBJ loader)
//   host_driver run <scene.bin> <out.bin> <mode> <n> [outdir]
//        mode = substeps : begin(CFL); advance(n); finish()         (parity against the oracle is done by the Python test)
//        mode = solve    : solve(CFL, maxt = n/60 - 1/120, alpha)   -> n frames, particle_N.obj / mesh_N.obj in outdir
//        mode = ckpt_save / ckpt_resume : n substeps + saveCheckpoint(<outdir>/state.ckpt) + n substeps / resume(...) + n substeps
//   host_driver objmesh <in.obj> <out.bin> density thickness E nu shear stiff angle_deg    (loader parity vs the reference's)
// Scene file = what tests/test_host_cpp.py writes: [int32 count] then per array: char name[32], int32 dtype (0 f64, 1 i32),
// int64 length, raw data.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <functional>
#include <iostream>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../include/aep/HybridSolver.h"
#include "../include/aep/LagrangianMesh.h"
#include "../include/aep/LevelSet.h"
#include "../include/aep/ParticleSystem.h"
#include "../include/aep/RegularGrid.h"

using namespace Eigen;

#define REQUIRE(c) do { if (!(c)) { std::fprintf(stderr, "REQUIRE failed: %s (%s:%d)\n", #c, __FILE__, __LINE__); std::exit(1); } } while (0)

struct Blob { std::map<std::string, std::vector<double>> d; std::map<std::string, std::vector<int32_t>> i; };

static Blob read_blob(const char* path) {
    Blob b; std::ifstream f(path, std::ios::binary); if (!f) throw std::runtime_error(std::string("cannot open ") + path);
    int32_t cnt = 0; f.read((char*)&cnt, 4);
    for (int a = 0; a < cnt; ++a) {
        char name[32]; int32_t dt; int64_t len; f.read(name, 32); f.read((char*)&dt, 4); f.read((char*)&len, 8);
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
static MatrixX3d mat3(const std::vector<double>& s) { MatrixX3d m; const std::ptrdiff_t n = (std::ptrdiff_t)s.size() / 3; m.resize(n, 3); std::memcpy(m.data(), s.data(), s.size() * 8); return m; }
static VectorXd vec(const std::vector<double>& s) { VectorXd v((std::ptrdiff_t)s.size()); std::memcpy(v.data(), s.data(), s.size() * 8); return v; }
static std::vector<Matrix3d> mats(const std::vector<double>& s) { std::vector<Matrix3d> M(s.size() / 9); for (size_t i = 0; i < M.size(); ++i) std::memcpy(M[i].data(), &s[9 * i], 72); return M; }

static int unit(const std::string& tmp) {
    // ---- RegularGrid (RegularGrid.cpp:117-177)
    RegularGrid rg(Vector3d(-1.0, 0.0, 0.5), Vector3d(1.0, 3.0, 2.0), Vector3i(8, 12, 6));
    REQUIRE(rg.gridNumber() == 8 * 12 * 6);
    REQUIRE(std::fabs(rg.h()[0] - 0.25) < 1e-15 && std::fabs(rg.h()[1] - 0.25) < 1e-15 && std::fabs(rg.h()[2] - 0.25) < 1e-15);
    REQUIRE(std::fabs(rg.gridVolume() - 0.015625) < 1e-15);
    for (int idx : {0, 7, 8, 95, 96, 575}) { int i, j, k; std::tie(i, j, k) = rg.toCoordinate(idx); REQUIRE(rg.toIndex(i, j, k) == idx); }
    REQUIRE(rg.toIndex(3, 2, 1) == 1 * 96 + 2 * 8 + 3);
    const MatrixX3d& P = rg.positions();
    REQUIRE(P.rows() == rg.gridNumber());
    REQUIRE(std::fabs(P(rg.toIndex(3, 2, 1), 0) - (-1.0 + 3 * 0.25)) < 1e-15 && std::fabs(P(rg.toIndex(3, 2, 1), 2) - 0.75) < 1e-15);
    rg.allocateHostMirrors(); rg.velocities(5, 0) = 3.0; rg.velocities(5, 1) = 4.0;
    REQUIRE(std::fabs(rg.max_velocity() - 5.0) < 1e-15 && std::fabs(rg.CFL_condition() - 20.0) < 1e-12);
    bool threw = false; try { RegularGrid bad(Vector3d(0, 0, 0), Vector3d(1, 1, 0), Vector3i(4, 4, 4)); } catch (const std::invalid_argument&) { threw = true; }
    REQUIRE(threw);
    // ---- level sets (LevelSet.cpp:8-42)
    REQUIRE(groundLevelSet(Vector3d(0, 0, 0.3), 0.1) > 0 && groundLevelSet(Vector3d(0, 0, 0.05), 0.1) < 0);
    REQUIRE(DgroundLevelSet(Vector3d(0, 0, 0), 0.1)[2] == 1.0);
    REQUIRE(std::fabs(wall2groundLevelSet(Vector3d(0.95, 0.2, 0.5), 0.9, 0.9, 0.1) + 0.05) < 1e-15);
    REQUIRE(Dwall2groundLevelSet(Vector3d(0.95, 0.2, 0.5), 0.9, 0.9, 0.1)[0] == -1.0);
    REQUIRE(Dwall2groundLevelSet(Vector3d(0.2, 0.95, 0.5), 0.9, 0.9, 0.1)[1] == -1.0);
    REQUIRE(Dwall2groundLevelSet(Vector3d(0.2, 0.2, 0.05), 0.9, 0.9, 0.1)[2] == 1.0);
    REQUIRE(sphereGroundLevelSet(Vector3d(0.5, 0.5, 0.25), Vector3d(0.5, 0.5, 0.2), 0.12, 0.05) < 0);
    REQUIRE(std::fabs(DsphereGroundLevelSet(Vector3d(0.5, 0.5, 0.25), Vector3d(0.5, 0.5, 0.2), 0.12, 0.05)[2] - 1.0) < 1e-12);
    REQUIRE(boxLevelSet(Vector3d(0.5, 0.5, 0.01), Vector3d(0.1, 0.1, 0.1), Vector3d(0.9, 0.9, 0.9)) < 0);
    REQUIRE(DboxLevelSet(Vector3d(0.5, 0.5, 0.01), Vector3d(0.1, 0.1, 0.1), Vector3d(0.9, 0.9, 0.9))[2] == 1.0);
    // ---- ParticleSystem factories (ParticleSystem.cpp:119-401): geometry, constants, determinism in the seed
    ParticleSystem sb = ParticleSystem::SandBlock(Vector3d(0.3, 0.3, 0.15), Vector3d(0.6, 0.6, 0.65), 0.08, 5000, 7);
    ParticleSystem sb2 = ParticleSystem::SandBlock(Vector3d(0.3, 0.3, 0.15), Vector3d(0.6, 0.6, 0.65), 0.08, 5000, 7);
    REQUIRE(sb.positions.rows() == 5000 && sb.masses.size() == 5000 && sb.elasticDeformationGradients.size() == 5000);
    double msum = 0;
    for (int p = 0; p < 5000; ++p) {
        const Vector3d x = aep_host::get_row(sb.positions, p);
        REQUIRE(x[0] >= 0.3 && x[0] <= 0.6 && x[2] >= 0.15 && x[2] <= 0.65);
        REQUIRE((x - Vector3d(0.3, 0.3, 0.4)).norm() >= 0.08);
        REQUIRE(sb.positions(p, 1) == sb2.positions(p, 1));
        REQUIRE(sb.affineMomenta_2(p, 1) == 0.0 && sb.elasticDeformationGradients[p](1, 1) == 1.0 && sb.plasticDeformationGradients[p](0, 1) == 0.0);
        msum += sb.masses[p];
    }
    REQUIRE(std::fabs(msum - 1300.0 * (0.3 * 0.3 * 0.5 - 0.25 * 3.14159265358979323846 * 0.08 * 0.08 * 0.08)) < 1e-9);
    REQUIRE(sb.youngsModulus == 3.537e5 && sb.poissonRatio == 0.3);
    ParticleSystem snow = ParticleSystem::SnowBall(Vector3d(0.5, 0.5, 0.5), 0.1, 1000, 3);
    REQUIRE(snow.youngsModulus == 1.4e5 && snow.poissonRatio == 0.2 && snow.criticalCompression == 2.5e-2 && snow.criticalStretch == 7.5e-3);
    for (int p = 0; p < 1000; ++p) REQUIRE((aep_host::get_row(snow.positions, p) - Vector3d(0.5, 0.5, 0.5)).norm() <= 0.1 + 1e-12);
    ParticleSystem cyl = ParticleSystem::SandCylinder(Vector3d(-0.5, 0.0, 0.1), 0.25, 0.6, 1000, 3);     // main.cpp:45-49
    for (int p = 0; p < 1000; ++p) { REQUIRE(cyl.velocities(p, 2) == -1.0 && cyl.velocities(p, 0) == 0.0); REQUIRE(cyl.positions(p, 2) >= 0.1 && cyl.positions(p, 2) <= 0.7); }
    ParticleSystem ball = ParticleSystem::SandBall(Vector3d(0, 0, 0), 0.2, 100, 1);
    REQUIRE(std::fabs(ball.masses.sum() - 1300.0 * 3.14 * 0.008) < 1e-9);
    // ---- LagrangianMesh: sheet -> OBJ -> ObjMesh round trip (LagrangianMesh.cpp:197-352), volumes, directions, constraints
    LagrangianMesh sheet = LagrangianMesh::SquareSheet(5, Vector3d(0.1, 0.2, 0.7), 0.8, 2e3, 0.04, 200, 0.3, 0.0, 4e4, 30.0);
    REQUIRE(sheet.vertexPositions.rows() == 25 && sheet.faces.rows() == 32);
    REQUIRE(std::fabs(sheet.elementVolumes.sum() - 0.25 * 0.64 * 0.04) < 1e-12);           // sum of areas = 0.8^2, x thickness / 4
    REQUIRE(std::fabs(sheet.vertexVolumes.sum() - 3.0 * sheet.elementVolumes.sum()) < 1e-12);
    REQUIRE(std::fabs(sheet.frictionCoeff - std::tan(30.0 * 3.14159265358979323846 / 180.0)) < 1e-15);
    REQUIRE(std::fabs(sheet.mu - 200.0 / 2.6) < 1e-12 && std::fabs(sheet.lambda - 200.0 * 0.3 / 1.3 / 0.4) < 1e-12);
    REQUIRE(std::fabs(sheet.elementDirections_3(0, 2) - 1.0) < 1e-12 && std::fabs(sheet.elementRestDirections_1()(0, 0) - 0.2) < 1e-12);
    REQUIRE(std::fabs(sheet.elementPositions(0, 2) - 0.7) < 1e-15);
    const std::string obj = tmp + "/sheet.obj";
    { std::ofstream o(obj); o << "# test sheet\n"; o.precision(9);
      for (int v = 0; v < 25; ++v) o << "v " << sheet.vertexPositions(v, 0) << ' ' << sheet.vertexPositions(v, 1) << ' ' << sheet.vertexPositions(v, 2) << "\nvn 0 0 1\n";
      for (int f = 0; f < 32; ++f) o << "f " << sheet.faces(f, 0) + 1 << "/1/1 " << sheet.faces(f, 1) + 1 << "/1/1 " << sheet.faces(f, 2) + 1 << "/1/1\n"; }
    LagrangianMesh loaded = LagrangianMesh::ObjMesh(obj, 2e3, 0.04, 200, 0.3, 0.0, 4e4, 30.0);
    REQUIRE(loaded.vertexPositions.rows() == 25 && loaded.faces.rows() == 32);
    for (int f = 0; f < 32; ++f) { REQUIRE(loaded.faces(f, 1) == sheet.faces(f, 1)); REQUIRE(std::fabs(loaded.elementMasses[f] - sheet.elementMasses[f]) < 1e-6 * sheet.elementMasses[f]);  /* positions parsed as float, like the reference */ }
    VectorXd fixed(25); fixed.setZero(); fixed[0] = 1.0; fixed[4] = 1.0;
    REQUIRE(!sheet.vertexIsFixed(0)); sheet.bindConstraints(&fixed); REQUIRE(sheet.vertexIsFixed(0) && sheet.vertexIsFixed(4) && !sheet.vertexIsFixed(1));
    VectorXd wrong(3); threw = false; try { sheet.bindConstraints(&wrong); } catch (const std::invalid_argument&) { threw = true; } REQUIRE(threw);
    threw = false; try { LagrangianMesh::ObjMesh(tmp + "/missing.obj", 1, 1, 1, 0.3, 0, 0, 0); } catch (const std::runtime_error&) { threw = true; } REQUIRE(threw);
    // ---- checkpoint file round trip (host only): containers + clock -> file -> scrambled containers -> identical again
    {
        ParticleSystem a = ParticleSystem::SandBlock(Vector3d(0.3, 0.3, 0.15), Vector3d(0.6, 0.6, 0.65), 0.08, 300, 7);
        for (int p = 0; p < 300; ++p) { a.affineMomenta_2(p, 1) = 0.25 * p; a.elasticDeformationGradients[p](0, 2) = 1e-3 * p; a.plasticAmount[p] = 0.01 * p; a.volumes[p] = 1e-6 * (p + 1); }
        LagrangianMesh sa = LagrangianMesh::SquareSheet(4, Vector3d(0.1, 0.2, 0.7), 0.6, 2e3, 0.04, 200, 0.3, 0.0, 4e4, 30.0);
        sa.elementDirections_3(5, 1) = 0.125; sa.vertexAffineMomenta_3(2, 0) = -2.0; sa.vertexVelocities(7, 2) = -0.5;
        const double clk[5] = {2.5e-4, 3.0 / 60.0, 0.004, 3.0, 123.0};
        const std::string ck = tmp + "/state.ckpt";
        HybridSolver::writeStateFile(ck, &a, &sa, clk);
        ParticleSystem b = ParticleSystem::SandBlock(Vector3d(0.3, 0.3, 0.15), Vector3d(0.6, 0.6, 0.65), 0.08, 300, 99);   // other seed: different state
        LagrangianMesh sb3 = LagrangianMesh::SquareSheet(4, Vector3d(0.0, 0.0, 0.5), 0.6, 2e3, 0.04, 200, 0.3, 0.0, 4e4, 30.0);
        double got[5] = {0, 0, 0, 0, 0};
        HybridSolver::readStateFile(ck, &b, &sb3, got);
        for (int i = 0; i < 5; ++i) REQUIRE(got[i] == clk[i]);
        for (int p = 0; p < 300; ++p) {
            for (int c = 0; c < 3; ++c) REQUIRE(b.positions(p, c) == a.positions(p, c) && b.affineMomenta_2(p, c) == a.affineMomenta_2(p, c));
            REQUIRE(b.elasticDeformationGradients[p](0, 2) == a.elasticDeformationGradients[p](0, 2) && b.plasticAmount[p] == a.plasticAmount[p] && b.volumes[p] == a.volumes[p]);
        }
        REQUIRE(sb3.elementDirections_3(5, 1) == 0.125 && sb3.vertexAffineMomenta_3(2, 0) == -2.0 && sb3.vertexVelocities(7, 2) == -0.5);
        REQUIRE(sb3.vertexPositions(3, 0) == sa.vertexPositions(3, 0) && sb3.elementPositions(1, 2) == sa.elementPositions(1, 2));
        ParticleSystem small = ParticleSystem::SandBall(Vector3d(0, 0, 0), 0.2, 10, 1);
        threw = false; try { HybridSolver::readStateFile(ck, &small, nullptr, got); } catch (const std::runtime_error&) { threw = true; } REQUIRE(threw);   // wrong particle count
        threw = false; try { HybridSolver::readStateFile(tmp + "/nope.ckpt", &b, nullptr, got); } catch (const std::runtime_error&) { threw = true; } REQUIRE(threw);
        { std::ofstream bad(tmp + "/bad.ckpt", std::ios::binary); bad << "not a checkpoint"; }
        threw = false; try { HybridSolver::readStateFile(tmp + "/bad.ckpt", &b, nullptr, got); } catch (const std::runtime_error&) { threw = true; } REQUIRE(threw);
        HybridSolver idle; threw = false; try { idle.saveCheckpoint(ck); } catch (const std::logic_error&) { threw = true; } REQUIRE(threw);              // no running context
    }
    // ---- HybridSolver argument checks that need no GPU
    HybridSolver hs; threw = false; try { hs.begin(0.3); } catch (const std::invalid_argument&) { threw = true; } REQUIRE(threw);
    std::printf("host unit OK\n");
    return 0;
}

static int run(int argc, char** argv) {
    const Blob b = read_blob(argv[2]);
    const std::string mode = argv[4]; const int n = std::atoi(argv[5]); const std::string outdir = argc > 6 ? argv[6] : ".";
    const std::vector<double>& g = b.d.at("grid");            // min3 max3
    const std::vector<int32_t>& res = b.i.at("res");
    const std::vector<double>& sc = b.d.at("scalars");        // material, cfl, ls_kind, ls_mode (0 analytic, 1 std::function), rate_floor
    RegularGrid rg(Vector3d(g[0], g[1], g[2]), Vector3d(g[3], g[4], g[5]), Vector3i(res[0], res[1], res[2]));
    HybridSolver solver; solver.setRegularGrid(&rg);
    std::unique_ptr<ParticleSystem> ps; std::unique_ptr<LagrangianMesh> mesh; VectorXd fixed;
    if (b.d.count("x")) {
        const std::vector<double>& mat = b.d.at("material");  // E nu thetaC thetaS
        const VectorXd m = vec(b.d.at("m")); VectorXd rho(m.size()); rho.setOnes(); MatrixX3d colors;
        ps.reset(new ParticleSystem(mat3(b.d.at("v")), mat3(b.d.at("x")), mats(b.d.at("FE")), mats(b.d.at("FP")), m, vec(b.d.at("vol")), rho, vec(b.d.at("q")),
                                    mat[0], mat[1], mat[2], mat[3], 0.2, colors));
        ps->affineMomenta_1 = mat3(b.d.at("B1")); ps->affineMomenta_2 = mat3(b.d.at("B2")); ps->affineMomenta_3 = mat3(b.d.at("B3"));
        solver.setParticleSystem(ps.get());
    }
    if (b.d.count("mesh_vx")) {
        const std::vector<int32_t>& fi = b.i.at("mesh_faces"); MatrixX3i F; F.resize((std::ptrdiff_t)fi.size() / 3, 3); std::memcpy(F.data(), fi.data(), fi.size() * 4);
        const std::vector<double>& mp = b.d.at("mesh_params");   // mu lambda shear stiff fric
        mesh.reset(new LagrangianMesh(mat3(b.d.at("mesh_vx")), F, mat3(b.d.at("mesh_vv")), mat3(b.d.at("mesh_ev")), vec(b.d.at("mesh_vm")), vec(b.d.at("mesh_vvol")),
                                      vec(b.d.at("mesh_em")), vec(b.d.at("mesh_evol")), mat3(b.d.at("mesh_d1")), mat3(b.d.at("mesh_d2")), mat3(b.d.at("mesh_d3")),
                                      mat3(b.d.at("mesh_D1")), mat3(b.d.at("mesh_D2")), mat3(b.d.at("mesh_D3")), mp[0], mp[1], mp[2], mp[3], mp[4]));
        if (b.d.count("mesh_fixed")) { fixed = vec(b.d.at("mesh_fixed")); mesh->bindConstraints(&fixed); }
        solver.setLagrangianMesh(mesh.get());
    }
    solver.setMaterialType(sc[0] == 0.0 ? SNOW : SAND);
    const int ls_kind = (int)sc[2]; const std::vector<double>& lp = b.d.at("ls_params");
    if (ls_kind != 0) {
        if (sc[3] == 0.0) solver.setAnalyticLevelSet(ls_kind, lp.data(), (int)lp.size());
        else {
            using namespace std::placeholders;                // the way main.cpp:86-91 binds its colliders
            if (ls_kind == AEP_LS_GROUND) solver.setLevelSet(std::bind(groundLevelSet, _1, lp[0]), std::bind(DgroundLevelSet, _1, lp[0]));
            else if (ls_kind == AEP_LS_WALL2GROUND) solver.setLevelSet(std::bind(wall2groundLevelSet, _1, lp[0], lp[1], lp[2]), std::bind(Dwall2groundLevelSet, _1, lp[0], lp[1], lp[2]));
            else if (ls_kind == AEP_LS_SPHERE_GROUND) { const Vector3d c(lp[0], lp[1], lp[2]); solver.setLevelSet(std::bind(sphereGroundLevelSet, _1, c, lp[3], lp[4]), std::bind(DsphereGroundLevelSet, _1, c, lp[3], lp[4])); }
            else { const Vector3d a(lp[0], lp[1], lp[2]), bb(lp[3], lp[4], lp[5]); solver.setLevelSet(std::bind(boxLevelSet, _1, a, bb), std::bind(DboxLevelSet, _1, a, bb)); }
        }
    }
    if (sc.size() > 4 && sc[4] > 0) solver.config().dt_rate_floor = sc[4];
    solver.setOutputDirectory(outdir); solver.setVerbose(false);
    double info[4] = {0, 0, 0, 0};
    if (mode == "ckpt_save" || mode == "ckpt_resume") {
        // ckpt_save: n substeps, checkpoint to <outdir>/state.ckpt, n more.  ckpt_resume: resume from that file, n substeps.
        solver.setWriteFrames(false);
        if (mode == "ckpt_save") { solver.begin(sc[1]); if (b.d.count("fixed_dt")) aep_set_fixed_dt(solver.context(), b.d.at("fixed_dt")[0]); solver.advance(n); solver.saveCheckpoint(outdir + "/state.ckpt"); }
        else { solver.resume(outdir + "/state.ckpt", sc[1]); if (b.d.count("fixed_dt")) aep_set_fixed_dt(solver.context(), b.d.at("fixed_dt")[0]); }
        solver.advance(n);
        int fr; long long ss; solver.clock(&info[0], &info[1], &fr, &ss); info[2] = fr; info[3] = (double)ss;
        solver.finish();
    } else if (mode == "substeps") {
        solver.setWriteFrames(false);
        solver.begin(sc[1]); solver.advance(n);
        int fr; long long ss; solver.clock(&info[0], &info[1], &fr, &ss); info[2] = fr; info[3] = (double)ss;
        solver.finish();
    } else {
        solver.solve(sc[1], n / 60.0 - 1.0 / 120.0, 0.95);
    }
    Writer w(argv[3]);
    w.put("info", info, 4);
    if (ps) {
        w.put("x", ps->positions.data(), ps->positions.size()); w.put("v", ps->velocities.data(), ps->velocities.size());
        w.put("B1", ps->affineMomenta_1.data(), ps->affineMomenta_1.size());
        std::vector<double> FE(9 * ps->elasticDeformationGradients.size()), FP(FE.size());
        for (size_t i = 0; i < ps->elasticDeformationGradients.size(); ++i) { std::memcpy(&FE[9 * i], ps->elasticDeformationGradients[i].data(), 72); std::memcpy(&FP[9 * i], ps->plasticDeformationGradients[i].data(), 72); }
        w.put("FE", FE.data(), (int64_t)FE.size()); w.put("FP", FP.data(), (int64_t)FP.size());
        w.put("vol", ps->volumes.data(), ps->volumes.size()); w.put("q", ps->plasticAmount.data(), ps->plasticAmount.size());
    }
    if (mesh) {
        w.put("mesh_vx", mesh->vertexPositions.data(), mesh->vertexPositions.size()); w.put("mesh_vv", mesh->vertexVelocities.data(), mesh->vertexVelocities.size());
        w.put("mesh_ex", mesh->elementPositions.data(), mesh->elementPositions.size()); w.put("mesh_d3", mesh->elementDirections_3.data(), mesh->elementDirections_3.size());
    }
    w.put("grid_m", rg.masses.data(), rg.masses.size()); w.put("grid_v", rg.velocities.data(), rg.velocities.size());
    return 0;
}

// host_driver objmesh <in.obj> <out.bin> density thickness E nu shear stiff angle_deg : what LagrangianMesh::ObjMesh derives
static int objmesh(char** argv) {
    LagrangianMesh M = LagrangianMesh::ObjMesh(argv[2], std::atof(argv[4]), std::atof(argv[5]), std::atof(argv[6]), std::atof(argv[7]),
                                               std::atof(argv[8]), std::atof(argv[9]), std::atof(argv[10]));
    Writer w(argv[3]);
    std::vector<double> F((size_t)M.faces.size()); for (std::ptrdiff_t i = 0; i < M.faces.size(); ++i) F[(size_t)i] = M.faces.data()[i];
    const double c[3] = {M.mu, M.lambda, M.frictionCoeff};
    w.put("vx", M.vertexPositions.data(), M.vertexPositions.size()); w.put("faces", F.data(), (int64_t)F.size());
    w.put("vm", M.vertexMasses.data(), M.vertexMasses.size()); w.put("vvol", M.vertexVolumes.data(), M.vertexVolumes.size());
    w.put("em", M.elementMasses.data(), M.elementMasses.size()); w.put("evol", M.elementVolumes.data(), M.elementVolumes.size());
    w.put("D1", M.elementRestDirections_1().data(), M.elementRestDirections_1().size());
    w.put("D2", M.elementRestDirections_2().data(), M.elementRestDirections_2().size());
    w.put("D3", M.elementRestDirections_3().data(), M.elementRestDirections_3().size());
    w.put("ex", M.elementPositions.data(), M.elementPositions.size()); w.put("consts", c, 3);
    return 0;
}

int main(int argc, char** argv) {
    try {
        if (argc >= 3 && std::string(argv[1]) == "unit") return unit(argv[2]);
        if (argc >= 11 && std::string(argv[1]) == "objmesh") return objmesh(argv);
        if (argc >= 6 && std::string(argv[1]) == "run") return run(argc, argv);
        std::fprintf(stderr, "usage: host_driver unit <tmpdir> | run <scene.bin> <out.bin> substeps|solve <n> [outdir]\n");
        return 2;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "host_driver: %s\n", e.what());
        return 3;
    }
}
