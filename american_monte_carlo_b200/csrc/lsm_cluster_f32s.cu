// One-cluster sweep kernels for small path sets, float path storage, float state (all degrees); see lsm_cluster.cuh.
#include "lsm_cluster.cuh"

namespace amc {

cudaError_t launch_cluster_f32s(int degree, const SweepArgs& a, cudaStream_t s) {
    return launch_cluster_d<float, float>(degree, a, s);
}

int64_t cluster_capacity_f32s(int degree) { return cluster_capacity_d<float, float>(degree); }

}  // namespace amc
