// Microbenchmark: what does a scalar FP64 instruction cost on an SM sub-partition whose FP64 pipe
// is saturated with DMMA.8x8x4?  (K1 mixes ~140 scalar FP64 instructions with 256 DMMAs per
// warp-step; profiles/r01_notes.md.)
//
// One CTA per SM, 16 warps (4 per sub-partition).  Every warp runs `iters` rounds of
//   NDMMA dmma (8 independent accumulators)  +  NS scalar fp64 ops (kind / dependence selectable)
// Output: cycles per round per sub-partition, and the slope per scalar instruction.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// MODE 0: independent DFMA (8 chains); 1: one dependent DFMA chain; 2: independent DADD;
// 3: FP32 FFMA independent (control); 4: dependent chain but only warps with (warp>>2)==0 do it
// (the others do DMMA only); 5: SHFL.64 + DADD pairs (butterfly)
template <int NDMMA, int NS, int MODE>
__global__ void __launch_bounds__(512, 1) k_mix(double* out, int iters, long long* cyc) {
  double c0[8], c1[8];
  for (int i = 0; i < 8; i++) c0[i] = threadIdx.x * 1e-3 + i, c1[i] = i * 0.5;
  double a = threadIdx.x * 1e-6, b = 1.0 + threadIdx.x * 1e-7;
  double s[8];
  float f[8];
  for (int i = 0; i < 8; i++) s[i] = 1.0 + threadIdx.x * 1e-9 + i, f[i] = 1.0f + i;
  const double m = 1.0 + 1e-9, d = 1e-7;
  const int grp = threadIdx.x >> 7;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NDMMA; i++) dmma884(c0[i & 7], c1[i & 7], a, b);
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < NS; i++) s[i & 7] = fma(s[i & 7], m, d);
    } else if (MODE == 1) {
#pragma unroll
      for (int i = 0; i < NS; i++) s[0] = fma(s[0], m, d);
    } else if (MODE == 2) {
#pragma unroll
      for (int i = 0; i < NS; i++) s[i & 7] = s[i & 7] + d;
    } else if (MODE == 3) {
#pragma unroll
      for (int i = 0; i < NS; i++) f[i & 7] = fmaf(f[i & 7], 1.0000001f, 1e-7f);
    } else if (MODE == 4) {
      if (grp == 0) {
#pragma unroll
        for (int i = 0; i < NS; i++) s[0] = fma(s[0], m, d);
      }
    } else if (MODE == 5) {
#pragma unroll
      for (int i = 0; i < NS; i++) s[i & 7] += __shfl_xor_sync(0xffffffffu, s[i & 7], 1 + (i & 1));
    }
  }
  const long long t1 = clock64();
  double r = 0;
  for (int i = 0; i < 8; i++) r += c0[i] + c1[i] + s[i] + (double)f[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int NDMMA, int NS, int MODE>
static double run(double* out, long long* cyc, int iters) {
  k_mix<NDMMA, NS, MODE><<<148, 512>>>(out, iters, cyc);
  cudaDeviceSynchronize();
  k_mix<NDMMA, NS, MODE><<<148, 512>>>(out, iters, cyc);
  cudaDeviceSynchronize();
  long long h;
  cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  return (double)h / iters;
}

template <int MODE>
static void sweep(const char* name, double* out, long long* cyc) {
  const int iters = 4000;
  const double base = run<64, 0, MODE>(out, cyc, iters);
  const double v8 = run<64, 8, MODE>(out, cyc, iters);
  const double v16 = run<64, 16, MODE>(out, cyc, iters);
  const double v32 = run<64, 32, MODE>(out, cyc, iters);
  const double v64 = run<64, 64, MODE>(out, cyc, iters);
  // 4 warps per sub-partition: ideal DMMA-only round = 4 * 64 * 16 = 4096 cycles
  printf("%-34s base %.0f  +8: %.0f  +16: %.0f  +32: %.0f  +64: %.0f   cycles/scalar-instr/warp "
         "(per sub-partition): %.2f %.2f %.2f %.2f\n",
         name, base, v8, v16, v32, v64, (v8 - base) / (8 * 4), (v16 - base) / (16 * 4),
         (v32 - base) / (32 * 4), (v64 - base) / (64 * 4));
}

int main() {
  double* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 512 * sizeof(double));
  cudaMalloc(&cyc, sizeof(long long));
  printf("round = 64 DMMA.8x8x4 per warp, 4 warps per sub-partition (ideal 4096 cycles)\n");
  sweep<0>("independent DFMA", out, cyc);
  sweep<1>("dependent DFMA chain", out, cyc);
  sweep<2>("independent DADD", out, cyc);
  sweep<3>("independent FFMA (control)", out, cyc);
  sweep<4>("dependent DFMA, 1 warp of 4 only", out, cyc);
  sweep<5>("SHFL.64 + DADD", out, cyc);
  // scalar-only rates for reference
  {
    const int iters = 4000;
    const double v = run<0, 64, 0>(out, cyc, iters);
    printf("scalar only: 64 independent DFMA per warp, 4 warps/sub-partition: %.0f cycles "
           "(%.2f per instr)\n", v, v / 256);
    const double w = run<0, 64, 1>(out, cyc, iters);
    printf("scalar only: 64 dependent DFMA per warp: %.0f cycles (%.2f latency)\n", w, w / 64);
  }
  printf("status %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
