/* LevelSet.h -- collider level sets of the reference (LevelSet.h:5-12, LevelSet.cpp:8-42) plus the two primitives the
 * BASELINE configs need that the reference lacks (sphere-on-ground, box).  phi <= 0 is inside the collider; the solver
 * only ever evaluates phi / grad phi at grid nodes (HybridSolver.cpp:473-482). */
#ifndef AEP_HOST_LEVELSET_H
#define AEP_HOST_LEVELSET_H
#include <functional>
#include "EigenShim.h"

using LevelSet = std::function<double(const Eigen::Vector3d&)>;
using DLevelSet = std::function<Eigen::Vector3d(const Eigen::Vector3d&)>;

double groundLevelSet(const Eigen::Vector3d& x, double groundZ);
Eigen::Vector3d DgroundLevelSet(const Eigen::Vector3d& x, double groundZ);

double wall2groundLevelSet(const Eigen::Vector3d& x, double wallX, double wallY, double groundZ);
Eigen::Vector3d Dwall2groundLevelSet(const Eigen::Vector3d& x, double wallX, double wallY, double groundZ);

/* extensions (no reference counterpart) */
double sphereGroundLevelSet(const Eigen::Vector3d& x, const Eigen::Vector3d& center, double radius, double groundZ);
Eigen::Vector3d DsphereGroundLevelSet(const Eigen::Vector3d& x, const Eigen::Vector3d& center, double radius, double groundZ);
double boxLevelSet(const Eigen::Vector3d& x, const Eigen::Vector3d& bmin, const Eigen::Vector3d& bmax);
Eigen::Vector3d DboxLevelSet(const Eigen::Vector3d& x, const Eigen::Vector3d& bmin, const Eigen::Vector3d& bmax);
#endif
