// containers.cpp -- host-side state containers with the reference's class surface: RegularGrid, ParticleSystem,
// LagrangianMesh, and the level-set primitives.  Pure host code (g++), no CUDA: these hold what the caller fills in and
// what HybridSolver downloads; all stepping happens in libaep_b200.so.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <random>
#include <sstream>
#include <stdexcept>
#include <string>

#include "../../../include/aep/LagrangianMesh.h"
#include "../../../include/aep/LevelSet.h"
#include "../../../include/aep/ParticleSystem.h"
#include "../../../include/aep/RegularGrid.h"

using namespace Eigen;
using aep_host::get_row;
using aep_host::set_row;

// =============================================================================================== level sets
// LevelSet.cpp:8-42: ground plane and wall+wall+ground corner, axis normals chosen by the closest face.
double groundLevelSet(const Vector3d& x, double groundZ) { return x[2] - groundZ; }
Vector3d DgroundLevelSet(const Vector3d&, double) { return Vector3d(0.0, 0.0, 1.0); }

double wall2groundLevelSet(const Vector3d& x, double wallX, double wallY, double groundZ) {
    return std::min(std::min(x[2] - groundZ, wallX - x[0]), wallY - x[1]);
}
Vector3d Dwall2groundLevelSet(const Vector3d& x, double wallX, double wallY, double groundZ) {
    const double dz = std::fabs(x[2] - groundZ), dx = std::fabs(wallX - x[0]), dy = std::fabs(wallY - x[1]);
    if (dz <= dx && dz <= dy) return Vector3d(0.0, 0.0, 1.0);
    if (dy <= dx) return Vector3d(0.0, -1.0, 0.0);
    return Vector3d(-1.0, 0.0, 0.0);
}
// extensions: same formulas as AEP_LS_SPHERE_GROUND / AEP_LS_BOX in the engine (aep_engine.cu, ls_phi / ls_normal_code)
double sphereGroundLevelSet(const Vector3d& x, const Vector3d& c, double radius, double groundZ) {
    return std::min((x - c).norm() - radius, x[2] - groundZ);
}
Vector3d DsphereGroundLevelSet(const Vector3d& x, const Vector3d& c, double radius, double groundZ) {
    const Vector3d d = x - c; const double r = d.norm();
    if (r - radius <= x[2] - groundZ && r > 0.0) return d / r;
    return Vector3d(0.0, 0.0, 1.0);
}
double boxLevelSet(const Vector3d& x, const Vector3d& a, const Vector3d& b) {
    double d = x[0] - a[0];
    d = std::min(d, b[0] - x[0]); d = std::min(d, x[1] - a[1]); d = std::min(d, b[1] - x[1]);
    d = std::min(d, x[2] - a[2]); d = std::min(d, b[2] - x[2]);
    return d;
}
Vector3d DboxLevelSet(const Vector3d& x, const Vector3d& a, const Vector3d& b) {
    const double d[6] = { x[2] - a[2], b[2] - x[2], x[0] - a[0], b[0] - x[0], x[1] - a[1], b[1] - x[1] };
    static const double n[6][3] = { {0, 0, 1}, {0, 0, -1}, {1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0} };
    int best = 0; for (int f = 1; f < 6; ++f) if (d[f] < d[best]) best = f;
    return Vector3d(n[best][0], n[best][1], n[best][2]);
}

// =============================================================================================== RegularGrid
// RegularGrid.cpp:117-162.  The reference spins forever on bad bounds (and its z test compares maxBound with itself,
// RegularGrid.cpp:127); here a bad argument throws.
RegularGrid::RegularGrid(const VectorXd& minBound, const VectorXd& maxBound, const Vector3i& resolution) : resolution_(resolution) {
    if (minBound.size() < 3 || maxBound.size() < 3) throw std::invalid_argument("RegularGrid: bounds need 3 components");
    for (int a = 0; a < 3; ++a) {
        minBound_[a] = minBound[a]; maxBound_[a] = maxBound[a];
        if (!(maxBound_[a] > minBound_[a])) throw std::invalid_argument("RegularGrid: maxBound must be bigger than minBound");
        if (resolution_[a] < 1) throw std::invalid_argument("RegularGrid: resolution must be positive");
        h_[a] = (maxBound_[a] - minBound_[a]) / resolution_[a];
    }
}
void RegularGrid::allocateHostMirrors() {
    const int n = gridNumber();
    if (masses.size() != n) { masses.resize(n); masses.setZero(); }
    if (forces.rows() != n) { forces.resize(n, 3); forces.setZero(); }
    if (velocities.rows() != n) { velocities.resize(n, 3); velocities.setZero(); }
}
int RegularGrid::toIndex(int i, int j, int k) const { return (k * resolution_[1] + j) * resolution_[0] + i; }   // RegularGrid.cpp:164-168
std::tuple<int, int, int> RegularGrid::toCoordinate(int index) const {
    const int plane = resolution_[0] * resolution_[1];
    const int k = index / plane, r = index - k * plane;
    return std::make_tuple(r % resolution_[0], r / resolution_[0], k);
}
const MatrixX3d& RegularGrid::positions() const {
    if (positions_.rows() != gridNumber()) {
        positions_.resize(gridNumber(), 3);
        for (int k = 0; k < resolution_[2]; ++k) for (int j = 0; j < resolution_[1]; ++j) for (int i = 0; i < resolution_[0]; ++i)
            set_row(positions_, toIndex(i, j, k), Vector3d(minBound_[0] + i * h_[0], minBound_[1] + j * h_[1], minBound_[2] + k * h_[2]));
    }
    return positions_;
}
double RegularGrid::max_velocity() const {                                      // RegularGrid.cpp:188-200
    double best = 0.0;
    for (std::ptrdiff_t i = 0; i < velocities.rows(); ++i) {
        const double s = velocities(i, 0) * velocities(i, 0) + velocities(i, 1) * velocities(i, 1) + velocities(i, 2) * velocities(i, 2);
        if (s > best) best = s;
    }
    return std::sqrt(best);
}

// =============================================================================================== ParticleSystem
ParticleSystem::ParticleSystem(const MatrixX3d& velocities_, const MatrixX3d& positions_, const std::vector<Matrix3d>& FE,
                               const std::vector<Matrix3d>& FP, const VectorXd& masses_, const VectorXd& volumes_, const VectorXd& densities_,
                               const VectorXd& plasticAmount_, double youngsModulus_, double poissonRatio_, double criticalCompression_,
                               double criticalStretch_, double friction_, const MatrixX3d& colors)
    : colors_(colors), masses(masses_), volumes(volumes_), densities(densities_), plasticAmount(plasticAmount_), velocities(velocities_),
      positions(positions_), elasticDeformationGradients(FE), plasticDeformationGradients(FP), youngsModulus(youngsModulus_),
      poissonRatio(poissonRatio_), criticalCompression(criticalCompression_), criticalStretch(criticalStretch_), friction(friction_) {
    const std::ptrdiff_t n = masses.size();
    if (positions.rows() != n || velocities.rows() != n || (std::ptrdiff_t)FE.size() != n || (std::ptrdiff_t)FP.size() != n)
        throw std::invalid_argument("ParticleSystem: array sizes disagree");
    // ParticleSystem.cpp:111-116 only resizes these (uninitialised memory in the reference); zero is the APIC start state
    affineMomenta_1.resize(n, 3); affineMomenta_1.setZero();
    affineMomenta_2.resize(n, 3); affineMomenta_2.setZero();
    affineMomenta_3.resize(n, 3); affineMomenta_3.setZero();
}

namespace {
// common tail of the four factories: F = I, q = 0, equal masses, unit volumes/densities (overwritten by the first P2G)
ParticleSystem make_system(const MatrixX3d& x, const MatrixX3d& v, double totalMass, double E, double nu) {
    const std::ptrdiff_t n = x.rows();
    std::vector<Matrix3d> FE((size_t)n, Matrix3d::Identity()), FP((size_t)n, Matrix3d::Identity());
    VectorXd m(n), vol(n), rho(n), q(n);
    m.setConstant(totalMass / (double)n); vol.setOnes(); rho.setOnes(); q.setZero();
    MatrixX3d colors;
    return ParticleSystem(v, x, FE, FP, m, vol, rho, q, E, nu, 2.5e-2, 7.5e-3, 0.2, colors);
}
struct Sampler {
    std::mt19937_64 gen; std::uniform_real_distribution<double> u{-1.0, 1.0};
    explicit Sampler(unsigned seed) : gen(seed) {}
    double operator()() { return u(gen); }
    Vector3d in_unit_ball() { for (;;) { Vector3d p((*this)(), (*this)(), (*this)()); if (p.norm() <= 1.0) return p; } }
};
MatrixX3d zeros(std::ptrdiff_t n) { MatrixX3d z; z.resize(n, 3); z.setZero(); return z; }
const double kPi = 3.14159265358979323846;
}  // namespace

// ParticleSystem.cpp:119-179: rejection-sampled ball, total mass 100 * 3.14 r^3 (sic), snow constants
ParticleSystem ParticleSystem::SnowBall(const Vector3d& center, double radius, int sampleNumber, unsigned seed) {
    Sampler s(seed); MatrixX3d x; x.resize(sampleNumber, 3);
    for (int i = 0; i < sampleNumber; ++i) set_row(x, i, center + radius * s.in_unit_ball());
    return make_system(x, zeros(sampleNumber), 100.0 * 3.14 * radius * radius * radius, 1.4e5, 0.2);
}
// ParticleSystem.cpp:181-239: same ball, total mass 1300 * 3.14 r^3 (sic), sand constants
ParticleSystem ParticleSystem::SandBall(const Vector3d& center, double radius, int sampleNumber, unsigned seed) {
    Sampler s(seed); MatrixX3d x; x.resize(sampleNumber, 3);
    for (int i = 0; i < sampleNumber; ++i) set_row(x, i, center + radius * s.in_unit_ball());
    return make_system(x, zeros(sampleNumber), 1300.0 * 3.14 * radius * radius * radius, 3.537e5, 0.3);
}
// ParticleSystem.cpp:241-327: uniform box minus a ball centred on the (xmin, ymin, zmid) edge; the mass formula subtracts
// pi/4 r^3 (as the reference does)
ParticleSystem ParticleSystem::SandBlock(const Vector3d& bmin, const Vector3d& bmax, double holeRadius, int sampleNumber, unsigned seed) {
    if (holeRadius * 2.0 >= bmax[2] - bmin[2]) throw std::invalid_argument("SandBlock: the hole is too big");
    const Vector3d hole(bmin[0], bmin[1], 0.5 * (bmax[2] + bmin[2])), ext = bmax - bmin;
    Sampler s(seed); MatrixX3d x; x.resize(sampleNumber, 3);
    for (int i = 0; i < sampleNumber;) {
        const Vector3d p(bmin[0] + 0.5 * (s() + 1.0) * ext[0], bmin[1] + 0.5 * (s() + 1.0) * ext[1], bmin[2] + 0.5 * (s() + 1.0) * ext[2]);
        if ((p - hole).norm() >= holeRadius) set_row(x, i++, p);
    }
    const double mass = 1300.0 * (ext.prod() - 0.25 * kPi * holeRadius * holeRadius * holeRadius);
    return make_system(x, zeros(sampleNumber), mass, 3.537e5, 0.3);
}
// ParticleSystem.cpp:329-401: disc x uniform height, initial velocity (0,0,-1), mass 1300 pi r^2 h
ParticleSystem ParticleSystem::SandCylinder(const Vector3d& baseCenter, double radius, double height, int sampleNumber, unsigned seed) {
    Sampler s(seed); MatrixX3d x, v; x.resize(sampleNumber, 3); v.resize(sampleNumber, 3);
    for (int i = 0; i < sampleNumber; ++i) {
        double a, b; do { a = s(); b = s(); } while (a * a + b * b > 1.0);
        set_row(x, i, baseCenter + Vector3d(radius * a, radius * b, 0.5 * (s() + 1.0) * height));
        set_row(v, i, Vector3d(0.0, 0.0, -1.0));        // ParticleSystem.cpp:358-361: zero, column 2 = 1, whole matrix negated
    }
    return make_system(x, v, 1300.0 * kPi * radius * radius * height, 3.537e5, 0.3);
}

// =============================================================================================== LagrangianMesh
LagrangianMesh::LagrangianMesh(const MatrixX3d& vx, const MatrixX3i& F, const MatrixX3d& vv, const MatrixX3d& ev, const VectorXd& vm,
                               const VectorXd& vvol, const VectorXd& em, const VectorXd& evol, const MatrixX3d& d1, const MatrixX3d& d2,
                               const MatrixX3d& d3, const MatrixX3d& D1, const MatrixX3d& D2, const MatrixX3d& D3, double mu_, double lambda_,
                               double shearStiffness_, double stiffness_, double frictionCoeff_)
    : elementRestDirections_1_(D1), elementRestDirections_2_(D2), elementRestDirections_3_(D3), vertexPositions(vx), vertexVelocities(vv),
      elementVelocities(ev), vertexMasses(vm), elementMasses(em), vertexVolumes(vvol), elementVolumes(evol), faces(F),
      elementDirections_1(d1), elementDirections_2(d2), elementDirections_3(d3), mu(mu_), lambda(lambda_), shearStiffness(shearStiffness_),
      stiffness(stiffness_), frictionCoeff(frictionCoeff_) {
    const std::ptrdiff_t nv = vertexPositions.rows(), nf = faces.rows();
    for (std::ptrdiff_t f = 0; f < nf; ++f) for (int c = 0; c < 3; ++c)
        if (faces(f, c) < 0 || faces(f, c) >= nv) throw std::invalid_argument("LagrangianMesh: face index out of range");
    elementPositions.resize(nf, 3);
    MatrixX3d* vB[3] = { &vertexAffineMomenta_1, &vertexAffineMomenta_2, &vertexAffineMomenta_3 };
    MatrixX3d* eB[3] = { &elementAffineMomenta_1, &elementAffineMomenta_2, &elementAffineMomenta_3 };
    for (int a = 0; a < 3; ++a) { vB[a]->resize(nv, 3); vB[a]->setZero(); eB[a]->resize(nf, 3); eB[a]->setZero(); }   // LagrangianMesh.cpp:172-184
    updateElementPositions();                                                                                          // :186
}
void LagrangianMesh::updateElementPositions() {                                // LagrangianMesh.cpp:371-380
    for (std::ptrdiff_t f = 0; f < faces.rows(); ++f)
        for (int c = 0; c < 3; ++c)
            elementPositions(f, c) = (vertexPositions(faces(f, 0), c) + vertexPositions(faces(f, 1), c) + vertexPositions(faces(f, 2), c)) / 3.0;
}
void LagrangianMesh::bindConstraints(VectorXd* p) {                            // LagrangianMesh.cpp:354-363 (spins forever there)
    if (!p || p->size() != vertexPositions.rows()) throw std::invalid_argument("LagrangianMesh::bindConstraints: the constraints are not compatible");
    vertexIsFixed_ = p;
}
bool LagrangianMesh::vertexIsFixed(int v) const {                              // LagrangianMesh.cpp:462-481
    if (!vertexIsFixed_) return false;
    if (v < 0 || v >= vertexIsFixed_->size()) throw std::out_of_range("LagrangianMesh::vertexIsFixed: vertex index outside the range");
    return (*vertexIsFixed_)[v] != 0.0;
}

// LagrangianMesh.cpp:288-351: per face volume = area * thickness / 4 (Heron's formula, geometry.cpp:12-24), every incident
// vertex gets the same amount; rest directions (v2-v1, v3-v1, unit normal); masses = density * volume; Lame from (E, nu);
// friction coefficient = tan(angle).
LagrangianMesh LagrangianMesh::FromTriangles(const MatrixX3d& V, const MatrixX3i& F, double density, double thickness, double E, double nu,
                                             double shearStiffness, double stiffness, double frictionAngleInDegree) {
    const std::ptrdiff_t nv = V.rows(), nf = F.rows();
    MatrixX3d vv = zeros(nv), ev = zeros(nf), D1, D2, D3;
    D1.resize(nf, 3); D2.resize(nf, 3); D3.resize(nf, 3);
    VectorXd vvol(nv), evol(nf); vvol.setZero(); evol.setZero();
    for (std::ptrdiff_t f = 0; f < nf; ++f) {
        for (int c = 0; c < 3; ++c) if (F(f, c) < 0 || F(f, c) >= nv) throw std::invalid_argument("mesh face index out of range");
        const Vector3d a = get_row(V, F(f, 0)), b = get_row(V, F(f, 1)), c = get_row(V, F(f, 2));
        const double la = (b - a).norm(), lb = (c - b).norm(), lc = (a - c).norm(), p = 0.5 * (la + lb + lc);
        const double area = std::sqrt(p * (p - la) * (p - lb) * (p - lc));
        const double vol = 0.25 * area * thickness;
        evol[f] = vol; vvol[F(f, 0)] += vol; vvol[F(f, 1)] += vol; vvol[F(f, 2)] += vol;
        set_row(D1, f, b - a); set_row(D2, f, c - a); set_row(D3, f, (b - a).cross(c - a).normalized());
    }
    VectorXd vm = density * vvol, em = density * evol;
    const double lambda = E * nu / (1.0 + nu) / (1.0 - 2.0 * nu), mu = E / 2.0 / (1.0 + nu);
    return LagrangianMesh(V, F, vv, ev, vm, vvol, em, evol, D1, D2, D3, D1, D2, D3, mu, lambda, shearStiffness, stiffness,
                          std::tan(frictionAngleInDegree * kPi / 180.0));
}

LagrangianMesh LagrangianMesh::ObjMesh(const std::string& filename, double density, double thickness, double E, double nu,
                                       double shearStiffness, double stiffness, double frictionAngleInDegree) {
    std::ifstream fin(filename);
    if (!fin.is_open()) throw std::runtime_error("ObjMesh: cannot open the obj file " + filename);
    std::vector<double> pos; std::vector<int> idx; std::string line, tag;
    while (std::getline(fin, line)) {
        std::istringstream ls(line);
        if (!(ls >> tag)) continue;
        if (tag == "v") {
            float p[3];                                                        // the reference parses positions as float (LagrangianMesh.cpp:217)
            if (!(ls >> p[0] >> p[1] >> p[2])) throw std::runtime_error("ObjMesh: malformed vertex line in " + filename);
            pos.push_back(p[0]); pos.push_back(p[1]); pos.push_back(p[2]);
        } else if (tag == "f") {
            std::string tok; int got = 0;
            while (got < 3 && (ls >> tok)) { idx.push_back(std::atoi(tok.c_str())); ++got; }     // "a", "a/b", "a/b/c", "a//c"
            if (got != 3) throw std::runtime_error("ObjMesh: only triangulated faces are supported (" + filename + ")");
        }
    }
    MatrixX3d V; MatrixX3i F; V.resize((std::ptrdiff_t)pos.size() / 3, 3); F.resize((std::ptrdiff_t)idx.size() / 3, 3);
    for (std::ptrdiff_t i = 0; i < V.rows(); ++i) for (int c = 0; c < 3; ++c) V(i, c) = pos[(size_t)(3 * i + c)];
    for (std::ptrdiff_t f = 0; f < F.rows(); ++f) for (int c = 0; c < 3; ++c) F(f, c) = idx[(size_t)(3 * f + c)] - 1;
    return FromTriangles(V, F, density, thickness, E, nu, shearStiffness, stiffness, frictionAngleInDegree);
}

LagrangianMesh LagrangianMesh::SquareSheet(int n, const Vector3d& origin, double side, double density, double thickness, double E, double nu,
                                           double shearStiffness, double stiffness, double frictionAngleInDegree) {
    if (n < 2) throw std::invalid_argument("SquareSheet: n must be >= 2");
    MatrixX3d V; MatrixX3i F; V.resize((std::ptrdiff_t)n * n, 3); F.resize((std::ptrdiff_t)2 * (n - 1) * (n - 1), 3);
    const double e = side / (n - 1);
    for (int j = 0; j < n; ++j) for (int i = 0; i < n; ++i) set_row(V, (std::ptrdiff_t)j * n + i, origin + Vector3d(i * e, j * e, 0.0));
    std::ptrdiff_t f = 0;
    for (int j = 0; j + 1 < n; ++j) for (int i = 0; i + 1 < n; ++i) {
        const int a = j * n + i, b = a + 1, c = a + n, d = c + 1;
        F(f, 0) = a; F(f, 1) = b; F(f, 2) = d; ++f;
        F(f, 0) = a; F(f, 1) = d; F(f, 2) = c; ++f;
    }
    return FromTriangles(V, F, density, thickness, E, nu, shearStiffness, stiffness, frictionAngleInDegree);
}
