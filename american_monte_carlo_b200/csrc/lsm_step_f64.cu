// Step kernels for double path storage, double state (all degrees); see lsm_step.cuh.
#include "lsm_sweep.cuh"

namespace amc {

cudaError_t launch_step_f64(int degree, int grid, const StepArgs& a, cudaStream_t s, bool pdl, int n_batch) {
    return launch_step_d<double, double>(degree, grid, a, s, pdl, n_batch);
}

int step_occupancy_f64(int degree) { return occupancy_d<double, double>(degree); }

cudaError_t launch_sweep_f64(int degree, int grid, const SweepArgs& a, cudaStream_t s) {
    return launch_sweep_d<double, double, false>(degree, grid, a, s);
}

int sweep_occupancy_f64(int degree) { return sweep_occupancy_d<double, double, false>(degree); }

}  // namespace amc
