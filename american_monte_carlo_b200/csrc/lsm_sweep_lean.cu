// Path-free sweep kernels (float prices regenerated from Philox counters; double or float state); see lsm_sweep.cuh.
#include "lsm_sweep.cuh"

namespace amc {

cudaError_t launch_sweep_lean(int state_f32, int degree, int grid, const SweepArgs& a, cudaStream_t s) {
    return state_f32 ? launch_sweep_d<float, float, true>(degree, grid, a, s)
                     : launch_sweep_d<float, double, true>(degree, grid, a, s);
}

int sweep_occupancy_lean(int state_f32, int degree) {
    return state_f32 ? sweep_occupancy_d<float, float, true>(degree) : sweep_occupancy_d<float, double, true>(degree);
}

}  // namespace amc
