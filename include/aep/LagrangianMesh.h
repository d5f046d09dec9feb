/* LagrangianMesh.h -- cloth mesh container with the reference's public members (LagrangianMesh.h:13-116).
 * The dead O(Nf^2) constructor helpers of the reference (buildFaceWings_, computeRestMetrics_, computeAreas_,
 * LagrangianMesh.cpp:13-125: their results are never read) are not reproduced; `faceWings`, `areas`, `inverseMetrics`
 * stay empty.  The cloth constitutive model itself runs on the GPU (csrc/aep_mesh.cuh). */
#ifndef AEP_HOST_LAGRANGIANMESH_H
#define AEP_HOST_LAGRANGIANMESH_H
#include <string>
#include <vector>
#include "EigenShim.h"

namespace igl { namespace viewer { class Viewer; } }

class LagrangianMesh {
private:
    igl::viewer::Viewer* viewer_ = nullptr;
    Eigen::VectorXd* vertexIsFixed_ = nullptr;  /* initialised (the reference leaves it dangling) */
    Eigen::MatrixX3d elementRestDirections_1_, elementRestDirections_2_, elementRestDirections_3_;
public:
    Eigen::MatrixX3d vertexPositions, elementPositions, vertexVelocities, elementVelocities;
    Eigen::VectorXd vertexMasses, elementMasses, vertexVolumes, elementVolumes;
    Eigen::MatrixX3d vertexAffineMomenta_1, vertexAffineMomenta_2, vertexAffineMomenta_3;
    Eigen::MatrixX3d elementAffineMomenta_1, elementAffineMomenta_2, elementAffineMomenta_3;
    Eigen::MatrixX3i faces;
    Eigen::MatrixX3i faceWings;
    Eigen::VectorXd areas;
    std::vector<Eigen::Matrix2d> inverseMetrics;
    Eigen::MatrixX3d elementDirections_1, elementDirections_2, elementDirections_3;
    double mu, lambda;
    double shearStiffness, stiffness;
    double frictionCoeff;

    LagrangianMesh(const Eigen::MatrixX3d& vertexPositions, const Eigen::MatrixX3i& faces,
                   const Eigen::MatrixX3d& vertexVelocities, const Eigen::MatrixX3d& elementVelocities,
                   const Eigen::VectorXd& vertexMasses, const Eigen::VectorXd& vertexVolumes,
                   const Eigen::VectorXd& elementMasses, const Eigen::VectorXd& elementVolumes,
                   const Eigen::MatrixX3d& elementDirections_1, const Eigen::MatrixX3d& elementDirections_2,
                   const Eigen::MatrixX3d& elementDirections_3, const Eigen::MatrixX3d& elementRestDirections_1,
                   const Eigen::MatrixX3d& elementRestDirections_2, const Eigen::MatrixX3d& elementRestDirections_3,
                   double mu, double lambda, double shearStiffness, double stiffness, double frictionCoeff);

    /* LagrangianMesh.cpp:197-352.  Face indices are read as int (the reference's unsigned short breaks above 65 535
     * vertices); "a/b/c" index triples are accepted. */
    static LagrangianMesh ObjMesh(const std::string& filename, double density, double thickness, double youngsModulus,
                                  double poissonRatio, double shearStiffness, double stiffness, double frictionAngleInDegree);
    /* the same constructor rules from in-memory geometry (regular n x n sheet for the BASELINE cloth configs) */
    static LagrangianMesh FromTriangles(const Eigen::MatrixX3d& V, const Eigen::MatrixX3i& F, double density, double thickness,
                                        double youngsModulus, double poissonRatio, double shearStiffness, double stiffness,
                                        double frictionAngleInDegree);
    static LagrangianMesh SquareSheet(int n, const Eigen::Vector3d& origin, double side, double density, double thickness,
                                      double youngsModulus, double poissonRatio, double shearStiffness, double stiffness,
                                      double frictionAngleInDegree);

    void bindViewer(igl::viewer::Viewer* viewer) { viewer_ = viewer; }
    void bindConstraints(Eigen::VectorXd* vertexIsFixed_p);
    void updateViewer() {}
    void updateElementPositions();

    const Eigen::MatrixX3d& elementRestDirections_1() const { return elementRestDirections_1_; }
    const Eigen::MatrixX3d& elementRestDirections_2() const { return elementRestDirections_2_; }
    const Eigen::MatrixX3d& elementRestDirections_3() const { return elementRestDirections_3_; }

    bool vertexIsFixed(int vertexID) const;
    const Eigen::VectorXd* constraints() const { return vertexIsFixed_; }
};
#endif
