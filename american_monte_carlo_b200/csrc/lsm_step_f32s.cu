// Step kernels for float path storage, float state (all degrees); see lsm_step.cuh.
#include "lsm_sweep.cuh"

namespace amc {

cudaError_t launch_step_f32s(int degree, int grid, const StepArgs& a, cudaStream_t s, bool pdl, int n_batch) {
    return launch_step_d<float, float>(degree, grid, a, s, pdl, n_batch);
}

int step_occupancy_f32s(int degree) { return occupancy_d<float, float>(degree); }

cudaError_t launch_sweep_f32s(int degree, int grid, const SweepArgs& a, cudaStream_t s) {
    return launch_sweep_d<float, float, false>(degree, grid, a, s);
}

int sweep_occupancy_f32s(int degree) { return sweep_occupancy_d<float, float, false>(degree); }

}  // namespace amc
