// N1: the Metropolis accept / reject of a static-HMC transition, batched.
//
// Replaces, per chain (reference paths):
//   MetropolisIntegrationTransition._sample_n_step  transitions.py:275-315
//     accept_prob = 0 if isnan(h_init - h_prop) else exp(min(0, h_init - h_prop))   (:301-305)
//     accept_stat = accept_prob unless the integrator raised                            (:309)
//     accept iff no integrator error and uniform < accept_prob                          (:310-311)
//     dir: proposal flipped (:299), new state flipped again (:314)
//          -> unchanged on accept, negated on reject
// A chain whose trajectory failed at its first step has no proposal (`state_p is state`, :300):
// accept_prob = 0.  Memory-bound elementwise select; one launch per transition.
#pragma once
#include "common.cuh"

namespace mb200 {

__global__ void __launch_bounds__(256)
    metropolis_select_kernel(double* __restrict__ pos, double* __restrict__ mom,
                             const double* __restrict__ pos_prop,
                             const double* __restrict__ mom_prop,
                             const double* __restrict__ h_init, const double* __restrict__ h_prop,
                             const int32_t* __restrict__ status, const int32_t* __restrict__ n_done,
                             int32_t* __restrict__ dir, const double* __restrict__ uniforms,
                             int64_t n_chains, int dim, double* __restrict__ accept_prob,
                             double* __restrict__ accept_stat, int32_t* __restrict__ accepted) {
  const int64_t total = n_chains * dim;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t ch = idx / dim;
    const int j = (int)(idx - ch * dim);
    const bool error = status != nullptr && status[ch] != MB200_STATUS_OK;
    const bool moved = n_done == nullptr || n_done[ch] > 0;
    double prob = 0.0;
    if (moved) {
      const double h_diff = h_init[ch] - h_prop[ch];
      prob = (h_diff != h_diff) ? 0.0 : exp(fmin(0.0, h_diff));
    }
    const bool acc = !error && (uniforms[ch] < prob);
    if (acc) {
      pos[idx] = pos_prop[idx];
      mom[idx] = mom_prop[idx];
    }
    if (j == 0) {
      if (accept_prob != nullptr) accept_prob[ch] = prob;
      if (accept_stat != nullptr) accept_stat[ch] = error ? 0.0 : prob;
      if (accepted != nullptr) accepted[ch] = acc ? 1 : 0;
      if (dir != nullptr && !acc) dir[ch] = -dir[ch];
    }
  }
}

}  // namespace mb200
