// Device-side target models (closed registry; SURVEY.md section 7 "hard part 6").
//
// The reference takes arbitrary Python callables `neg_log_dens` / `grad_neg_log_dens`
// (systems.py:88-95); a fused GPU gradient cannot, so the benchmark targets are compiled in.
// The NumPy statements of the same models used by the oracle are in oracle/targets.py.
//
// "Pair" interface used by the Euclidean leapfrog kernels: a chain's position vector is owned
// in aligned pairs of consecutive coordinates (i, i+1), i even, by whatever threads the kernel
// layout chooses.  A gradient evaluation is
//   1. red[r] = sum over pairs of accumulate(...)         (per-chain sum reductions, NRED of them)
//   2. grad_pair(...) using the reduced values.
// Coordinates with index >= dim are phantom (value 0) and must not contribute.
#pragma once
#include "common.cuh"

namespace mb200 {

struct StdGaussianTarget {
  static constexpr int NRED = 0;
  __device__ StdGaussianTarget(const ModelArgs&, int) {}
  __device__ __forceinline__ void accumulate(int, double, double, double*) const {}
  __device__ __forceinline__ void grad_pair(int, double q0, double q1, const double*, double& g0,
                                            double& g1) const {
    g0 = q0;
    g1 = q1;
  }
  // contribution of the pair to l(q) (summed over pairs by the caller)
  __device__ __forceinline__ double nld_pair(int, double q0, double q1, const double*) const {
    return 0.5 * (q0 * q0 + q1 * q1);
  }
};

// v = q[0], x = q[1:]:  l = v^2/18 + (D-1) v/2 + exp(-v) |x|^2 / 2
struct NealFunnelTarget {
  static constexpr int NRED = 2;  // red[0] = |x|^2, red[1] = v
  int dim;
  __device__ NealFunnelTarget(const ModelArgs&, int d) : dim(d) {}
  __device__ __forceinline__ void accumulate(int i, double q0, double q1, double* red) const {
    if (i == 0) {
      red[0] += q1 * q1;
      red[1] += q0;
    } else {
      red[0] += q0 * q0 + q1 * q1;
    }
  }
  __device__ __forceinline__ void grad_pair(int i, double q0, double q1, const double* red,
                                            double& g0, double& g1) const {
    const double e = exp(-red[1]);
    g0 = (i == 0) ? (red[1] / 9.0 + 0.5 * (dim - 1) - 0.5 * e * red[0]) : e * q0;
    g1 = e * q1;
  }
  __device__ __forceinline__ double nld_pair(int i, double, double, const double* red) const {
    if (i != 0) return 0.0;
    const double v = red[1];
    return v * v / 18.0 + 0.5 * (dim - 1) * v + 0.5 * exp(-v) * red[0];
  }
};

// pairs (x, y) = (q[2k], q[2k+1]):  l = sum x^2/8 + (y - b x^2)^2 / 2
struct BananaTarget {
  static constexpr int NRED = 0;
  double b;
  __device__ BananaTarget(const ModelArgs& m, int) : b(m.tp[0]) {}
  __device__ __forceinline__ void accumulate(int, double, double, double*) const {}
  __device__ __forceinline__ void grad_pair(int, double x, double y, const double*, double& g0,
                                            double& g1) const {
    const double r = y - b * x * x;
    g0 = x / 4.0 - 2.0 * b * x * r;
    g1 = r;
  }
  __device__ __forceinline__ double nld_pair(int, double x, double y, const double*) const {
    const double r = y - b * x * x;
    return x * x / 8.0 + 0.5 * r * r;
  }
};

}  // namespace mb200
