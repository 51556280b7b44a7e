// Fixed-metric products shared by the warp-per-chain kernels (pair layout: lane l owns the
// coordinate pairs (2l + 64k, 2l + 64k + 1), k < KP).
#pragma once
#include "common.cuh"

namespace mb200 {

// v = M^-1 p for CPW vectors held by one warp.
//   identity : v = p                                   (matrices.py:504-508)
//   diagonal : v = (1/diag) * p                        (matrices.py:726-733, 783-784)
//   dense    : v = A p with the explicit symmetric inverse A (matrices.py:222-223, 1183-1188),
//              rows of A streamed through L1/L2, vectors broadcast from shared memory `psm`
//              (CPW * 64 * KP doubles per warp).
  // v = M^-1 p for the CPW chains of this warp
template <int KP, int CPW>
__device__ __forceinline__ void inv_metric_apply(int METRIC, const double* __restrict__ minv,
                                                    int dim, int lane, double* psm,
                                                    const double (&p)[CPW][2 * KP],
                                                    double (&v)[CPW][2 * KP]) {
  constexpr int NV = 2 * KP;
    if (METRIC == MB200_METRIC_IDENTITY) {
#pragma unroll
      for (int c = 0; c < CPW; ++c)
#pragma unroll
        for (int e = 0; e < NV; ++e) v[c][e] = p[c][e];
    } else if (METRIC == MB200_METRIC_DIAGONAL) {
#pragma unroll
      for (int k = 0; k < KP; ++k) {
        const int i = 2 * lane + 64 * k;
        const double d0 = (i < dim) ? minv[i] : 0.0;
        const double d1 = (i + 1 < dim) ? minv[i + 1] : 0.0;
#pragma unroll
        for (int c = 0; c < CPW; ++c) {
          v[c][2 * k] = d0 * p[c][2 * k];
          v[c][2 * k + 1] = d1 * p[c][2 * k + 1];
        }
      }
    } else {
      // stage momenta: psm[c * 64KP + i]
      constexpr int DP = 64 * KP;
      __syncwarp();
#pragma unroll
      for (int c = 0; c < CPW; ++c)
#pragma unroll
        for (int k = 0; k < KP; ++k) {
          const int i = 2 * lane + 64 * k;
          psm[c * DP + i] = p[c][2 * k];
          psm[c * DP + i + 1] = p[c][2 * k + 1];
        }
      __syncwarp();
#pragma unroll
      for (int c = 0; c < CPW; ++c)
#pragma unroll
        for (int e = 0; e < NV; ++e) v[c][e] = 0.0;
      const bool even = (dim & 1) == 0;
      for (int j = 0; j < dim; ++j) {
        const double* row = minv + (size_t)j * dim;  // A[j][:] == A[:][j] (symmetric)
        double a[NV];
#pragma unroll
        for (int k = 0; k < KP; ++k) {
          const int i = 2 * lane + 64 * k;
          if (even) {
            if (i < dim) {
              const double2 t = *reinterpret_cast<const double2*>(row + i);
              a[2 * k] = t.x;
              a[2 * k + 1] = t.y;
            } else {
              a[2 * k] = 0.0;
              a[2 * k + 1] = 0.0;
            }
          } else {
            a[2 * k] = (i < dim) ? row[i] : 0.0;
            a[2 * k + 1] = (i + 1 < dim) ? row[i + 1] : 0.0;
          }
        }
#pragma unroll
        for (int c = 0; c < CPW; ++c) {
          const double pj = psm[c * DP + j];
#pragma unroll
          for (int e = 0; e < NV; ++e) v[c][e] = fma(a[e], pj, v[c][e]);
        }
      }
    }
  }

}  // namespace mb200
