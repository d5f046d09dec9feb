// aep_mesh.cuh -- LagrangianMesh (cloth) path.  PLACEHOLDER until the particle path is verified on hardware.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "aep_kernels.cuh"

namespace aep {
struct MeshState { long long nv = 0, nf = 0; int n_fixed = 0; };
inline void mesh_free(MeshState&) {}
inline int mesh_upload(MeshState&, const GridP&, const double*, const double*, int64_t, int64_t, const double*, const double*, const double*,
                       const double*, const double*, const int32_t*, const double*, const double*, const double*, const double*, const double*,
                       const double*, const double*, double, double, double, double, double, cudaStream_t) { return -1; }
inline int mesh_p2g(MeshState&, const GridP&, cudaStream_t, long long*) { return 0; }
inline int mesh_forces(MeshState&, const GridP&, cudaStream_t, long long*) { return 0; }
inline int mesh_pin(MeshState&, const GridP&, cudaStream_t, long long*) { return 0; }
inline int mesh_g2p(MeshState&, const GridP&, SimClock*, cudaStream_t, long long*) { return 0; }
inline int mesh_download(MeshState&, double*, double*, double*, double*, double*, double*, double*, cudaStream_t) { return 0; }
}  // namespace aep
