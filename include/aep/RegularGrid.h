/* RegularGrid.h -- node-centred dense grid container with the reference's public surface (RegularGrid.h:15-60).
 * Nodes sit at min + (i,j,k) h, i in [0,res), h = (max-min)/res; index = k nx ny + j nx + i (RegularGrid.cpp:137-168).
 * On the B200 path the grid lives in HBM; `velocities / forces / masses` are host mirrors that HybridSolver::solve
 * refreshes when it returns (and on request), not per substep. */
#ifndef AEP_HOST_REGULARGRID_H
#define AEP_HOST_REGULARGRID_H
#include <tuple>
#include "EigenShim.h"

namespace igl { namespace viewer { class Viewer; } }

class RegularGrid {
private:
    Eigen::Vector3d minBound_, maxBound_;       /* Vector3d, not VectorXd: fixes the dangling-reference accessors */
    Eigen::Vector3i resolution_;
    Eigen::Vector3d h_;
    igl::viewer::Viewer* viewer_ = nullptr;
    mutable Eigen::MatrixX3d positions_;        /* node position table, built on first use (3 doubles per node) */
public:
    Eigen::MatrixX3d velocities;
    Eigen::MatrixX3d forces;
    Eigen::VectorXd masses;

    RegularGrid(const Eigen::VectorXd& minBound, const Eigen::VectorXd& maxBound, const Eigen::Vector3i& resolution);

    int toIndex(int i, int j, int k) const;
    std::tuple<int, int, int> toCoordinate(int index) const;

    void bindViewer(igl::viewer::Viewer* viewer) { viewer_ = viewer; }
    void updateViewer() {}                      /* rendering is out of scope; hook kept for source compatibility */
    void recomputeColors() {}
    int gridNumber() const { return resolution_[0] * resolution_[1] * resolution_[2]; }
    double gridVolume() const { return h_[0] * h_[1] * h_[2]; }
    const Eigen::Vector3d& minBound() const { return minBound_; }
    const Eigen::Vector3d& maxBound() const { return maxBound_; }
    const Eigen::Vector3d& h() const { return h_; }
    const Eigen::Vector3i& resolution() const { return resolution_; }
    const Eigen::MatrixX3d& positions() const;
    double max_velocity() const;
    double CFL_condition() const { return max_velocity() / h_.minCoeff(); }
    /* host mirrors are allocated on demand (a 512^3 grid's mirrors are 7.5 GB) */
    void allocateHostMirrors();
};
#endif
