// One-cluster sweep kernels for small path sets, double path storage, double state (all degrees); see lsm_cluster.cuh.
#include "lsm_cluster.cuh"

namespace amc {

cudaError_t launch_cluster_f64(int degree, const SweepArgs& a, cudaStream_t s) {
    return launch_cluster_d<double, double>(degree, a, s);
}

int64_t cluster_capacity_f64(int degree) { return cluster_capacity_d<double, double>(degree); }

}  // namespace amc
