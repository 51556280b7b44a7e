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

// Experiment (NOT enabled by default; build with -DMB200_NUTS_LDS_MATVEC; parity-correct but
// measured SLOWER than the plain loop in round 1: 47 vs 64 M steps/s, profiles/r01_notes.md): v = A p for ONE chain per warp with the dense symmetric A staged in
// shared memory -- ld.shared.v2 row reads, four rows per trip with four independent accumulator
// sets (the profile of the plain loop shows one row load + 4 FMAs per trip with the loop / bounds
// logic around them and generic-address loads: profiles/r01_nuts_c1_ncu_lines.txt).  Requires
// even `dim`.  Summation order differs from inv_metric_apply (four partial sums).
template <int KP>
__device__ __forceinline__ void inv_metric_apply_staged(const double* s_minv, int dim, int lane,
                                                        double* psm, const double (&p)[1][2 * KP],
                                                        double (&v)[1][2 * KP]) {
  constexpr int NV = 2 * KP;
  __syncwarp();
#pragma unroll
  for (int k = 0; k < KP; ++k) {
    const int i = 2 * lane + 64 * k;
    psm[i] = p[0][2 * k];
    psm[i + 1] = p[0][2 * k + 1];
  }
  __syncwarp();
  const unsigned base = (unsigned)__cvta_generic_to_shared(s_minv);
  const unsigned pbase = (unsigned)__cvta_generic_to_shared(psm);
  double acc[4][NV];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int e = 0; e < NV; ++e) acc[r][e] = 0.0;
  const int dim4 = dim & ~3;
  for (int j = 0; j < dim4; j += 4) {
    double a[4][NV], pj[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      asm volatile("ld.shared.f64 %0, [%1];" : "=d"(pj[r]) : "r"(pbase + (unsigned)(j + r) * 8u));
#pragma unroll
      for (int k = 0; k < KP; ++k) {
        const int i = 2 * lane + 64 * k;
        if (i < dim) {
          asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];"
                       : "=d"(a[r][2 * k]), "=d"(a[r][2 * k + 1])
                       : "r"(base + (unsigned)((j + r) * dim + i) * 8u));
        } else {
          a[r][2 * k] = 0.0, a[r][2 * k + 1] = 0.0;
        }
      }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int e = 0; e < NV; ++e) acc[r][e] = fma(a[r][e], pj[r], acc[r][e]);
  }
  for (int j = dim4; j < dim; ++j) {
    const double pj = psm[j];
#pragma unroll
    for (int k = 0; k < KP; ++k) {
      const int i = 2 * lane + 64 * k;
      if (i < dim) {
        acc[0][2 * k] = fma(s_minv[(size_t)j * dim + i], pj, acc[0][2 * k]);
        acc[0][2 * k + 1] = fma(s_minv[(size_t)j * dim + i + 1], pj, acc[0][2 * k + 1]);
      }
    }
  }
#pragma unroll
  for (int e = 0; e < NV; ++e) v[0][e] = (acc[0][e] + acc[1][e]) + (acc[2][e] + acc[3][e]);
}

}  // namespace mb200
