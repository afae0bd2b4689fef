// Step kernels for float path storage, double state (all degrees); see lsm_step.cuh.
#include "lsm_sweep.cuh"

namespace amc {

cudaError_t launch_step_f32(int degree, int grid, const StepArgs& a, cudaStream_t s, bool pdl, int n_batch) {
    return launch_step_d<float, double>(degree, grid, a, s, pdl, n_batch);
}

int step_occupancy_f32(int degree) { return occupancy_d<float, double>(degree); }

cudaError_t launch_sweep_f32(int degree, int grid, const SweepArgs& a, cudaStream_t s) {
    return launch_sweep_d<float, double, false>(degree, grid, a, s);
}

int sweep_occupancy_f32(int degree) { return sweep_occupancy_d<float, double, false>(degree); }

}  // namespace amc
