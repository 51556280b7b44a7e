// Stand-alone harness for the K1 tensor-core leapfrog kernel (mici_b200/csrc/leapfrog_dmma.cuh):
// C1-shaped synthetic input (8192 chains x 128, funnel target, SPD metric), CUDA-event timing of
// 50- and 200-step launches, a checksum to compare variants, and (with -DK1_TRACE) a per-warp
// phase timeline of CTA 0 (clock64 at the phase boundaries marked MB200_K1_TRACE in the kernel).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -fmad=false -lineinfo \
//        [-DK1_TRACE] -o k1_bench k1_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

#ifdef K1_TRACE
constexpr int TR_STEPS = 8, TR_PHASES = 6, TR_WARPS = 32;
__device__ long long g_trace[TR_WARPS][TR_STEPS][TR_PHASES];
#define MB200_K1_TRACE(phase)                                                          \
  do {                                                                                 \
    if (blockIdx.x == 0 && (threadIdx.x & 31) == 0 && s >= 20 && s < 20 + TR_STEPS)    \
      g_trace[threadIdx.x >> 5][s - 20][phase] = clock64();                            \
  } while (0)
__device__ long long g_mark[2][8];
__device__ unsigned long long g_gt[2][8];
#define MB200_K1_MARK(id)                                                              \
  do {                                                                                 \
    if ((blockIdx.x == 0 || blockIdx.x == 100) && threadIdx.x == 0) {                  \
      unsigned long long gt;                                                           \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));                           \
      g_mark[blockIdx.x == 100][id] = clock64();                                       \
      g_gt[blockIdx.x == 100][id] = gt;                                                \
    }                                                                                  \
  } while (0)
#endif
#include "../../mici_b200/csrc/leapfrog_dmma.cuh"

using namespace mb200;

int main(int argc, char** argv) {
  const int64_t n = argc > 1 ? atoll(argv[1]) : 8192;
  const int dim = argc > 2 ? atoi(argv[2]) : 128;
  std::vector<double> q(n * dim), p(n * dim), minv(dim * dim);
  srand(1234);
  auto rnd = [] { return (rand() / (double)RAND_MAX - 0.5) * 2.0; };
  for (auto& v : q) v = 0.1 * rnd();
  for (auto& v : p) v = rnd();
  // SPD "M^-1": I + G G^T / dim
  std::vector<double> g(dim * dim);
  for (auto& v : g) v = rnd();
  for (int i = 0; i < dim; ++i)
    for (int j = 0; j < dim; ++j) {
      double s = (i == j) ? 1.0 : 0.0;
      for (int k = 0; k < dim; ++k) s += g[i * dim + k] * g[j * dim + k] / dim;
      minv[i * dim + j] = s;
    }
  for (int i = 0; i < dim; ++i)
    for (int j = 0; j < i; ++j) minv[i * dim + j] = minv[j * dim + i];
  double *dq, *dp, *dqo, *dpo, *dm, *dh;
  int32_t *dst, *dnd;
  cudaMalloc(&dq, n * dim * 8), cudaMalloc(&dp, n * dim * 8), cudaMalloc(&dqo, n * dim * 8);
  cudaMalloc(&dpo, n * dim * 8), cudaMalloc(&dm, dim * dim * 8), cudaMalloc(&dh, n * 8);
  cudaMalloc(&dst, n * 4), cudaMalloc(&dnd, n * 4);
  cudaMemcpy(dq, q.data(), n * dim * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(dp, p.data(), n * dim * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(dm, minv.data(), dim * dim * 8, cudaMemcpyHostToDevice);
  ModelArgs m;
  memset(&m, 0, sizeof(m));
  m.target_id = MB200_TARGET_NEAL_FUNNEL;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  const bool with_h = argc > 3 ? atoi(argv[3]) != 0 : true;
  for (int steps : {1, 2, 10, 50, 200}) {
    float best = 1e30f;
    for (int rep = 0; rep < 6; ++rep) {
      cudaEventRecord(e0);
      int rc = leapfrog_dmma_dispatch(dq, dp, dqo, dpo, nullptr, nullptr, n, dim, 0.01, steps, dm, m, with_h ? dh : nullptr, dst,
                                      dnd, 0);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      if (rc != 0) { printf("dispatch rc=%d\n", rc); return 1; }
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      if (rep > 0 && ms < best) best = ms;
    }
    std::vector<double> qo(n * dim), ho(n);
    cudaMemcpy(qo.data(), dqo, n * dim * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(ho.data(), dh, n * 8, cudaMemcpyDeviceToHost);
    double cs = 0, hs = 0;
    for (auto v : qo) cs += v;
    for (auto v : ho) hs += v;
    const double rate = (double)n * steps / (best * 1e-3);
    printf("steps %3d: %.4f ms  %.4f G steps/s  DMMA %.1f%% of 37.1 TF  HBM-roofline %.3f  "
           "checksum q %.15e h %.15e  (%s)\n",
           steps, best, rate * 1e-9, rate * 2.0 * dim * dim / 37.1e12 * 100,
           rate * 32.0 * dim / 6566.4e9, cs, hs, cudaGetErrorString(cudaGetLastError()));
  }
#ifdef K1_TRACE
  {
    long long mk[2][8];
    unsigned long long gt[2][8];
    cudaMemcpyFromSymbol(mk, g_mark, sizeof(mk));
    cudaMemcpyFromSymbol(gt, g_gt, sizeof(gt));
    printf("marks of the last launch (200 steps), thread 0 of CTA 0 / CTA 100: 0 kernel start, 1 smem zeroed, "
           "2 A landed, 3 A scaled, 4 state loaded, 5 first kick done, 6 steps done, 7 stored\n");
    for (int b = 0; b < 2; ++b) {
      printf("CTA %3d cycles:", b ? 100 : 0);
      for (int i = 0; i < 8; ++i) printf(" %9lld", mk[b][i] - mk[b][0]);
      printf("\n        ns    :");
      for (int i = 0; i < 8; ++i) printf(" %9lld", (long long)(gt[b][i] - gt[0][0]));
      printf("\n");
    }
  }
  {
    static long long tr[TR_WARPS][TR_STEPS][TR_PHASES];
    cudaMemcpyFromSymbol(tr, g_trace, sizeof(tr));
    long long t0 = tr[0][0][0];
    for (int w = 0; w < DMMA_THREADS / 32; ++w)
      for (int s = 0; s < TR_STEPS; ++s)
        if (tr[w][s][0] < t0 && tr[w][s][0] > 0) t0 = tr[w][s][0];
    printf("trace (CTA 0, cycles relative to first event; phases: 0 drift start, 1 drift end, "
           "2 partials written, 3 reduce barrier passed, 4 kick done, 5 publish barrier passed)\n");
    for (int w = 0; w < DMMA_THREADS / 32; ++w)
      for (int s = 0; s < TR_STEPS; ++s) {
        printf("w%02d sp%d g%d s%d:", w, w & 3, w >> 2, s);
        for (int ph = 0; ph < TR_PHASES; ++ph) printf(" %7lld", tr[w][s][ph] - t0);
        printf("\n");
      }
  }
#endif
  return 0;
}
