// K1: explicit leapfrog, dense shared metric, fused target gradient, n_steps per launch --
// FP64 tensor-core kernel (DMMA m8n8k4) for dim <= 128.
//
// Replaces, for every chain at once (reference paths):
//   LeapfrogIntegrator._step          integrators.py:170-173
//   System.h1_flow                    systems.py:143-152      p -= (dt/2) grad l(q)
//   EuclideanMetricSystem.h2_flow     systems.py:362-363      q += dt * (M^-1 p)
//   explicit-inverse mat-vec          matrices.py:222-223 (ExplicitArrayMatrix @ vector)
//
// The one genuine contraction on the path is V = P * A  ([chains x D] . [D x D], A = M^-1
// explicit and symmetric): 2 D^2 flop per chain-step against 32 D bytes of state, i.e. above the
// B200 fp64 ridge, so the kernel is organised around the FP64 tensor pipe (tcgen05 has no fp64
// kind; DMMA.8x8x4 is the native sm_100a instruction, 37.1 TFLOP/s measured:
// profiles/r01_fp64_peak.txt).
//
// Work decomposition (one CTA per SM, 16 warps, 4 groups x 4 warps):
//   * a CTA owns up to 56 chains = 7 row tiles of 8 chains; groups 0-2 take two tiles each,
//     group 3 one -- 7 tile-quarters per SM sub-partition, which balances 8192 chains over
//     148 SMs x 4 sub-partitions (8192 / 148 = 55.4 chains per SM);
//   * warp w of a group (one per sub-partition) computes output columns [w*DP/4, (w+1)*DP/4) for
//     the group's tiles: accumulators and positions of that slice stay in registers in the DMMA
//     C-fragment layout for the whole launch, momenta in a shared-memory tile (HBM is touched
//     once on entry and once on exit);
//   * A lives in shared memory for the whole launch (staged by TMA bulk copies, row stride
//     padded by 4 doubles so A/B fragment loads are bank-conflict free); B fragments are read
//     through the symmetry A[k][n] = A[n][k] as 8 rows x 4 consecutive doubles;
//   * momenta are exchanged through the shared-memory tile once per step (A fragments), and the
//     per-chain reductions of the target gradient through per-warp partial sums; the groups
//     synchronise only internally (named barriers), so while one group is in its gradient /
//     update phase the other three keep the sub-partition's DMMA pipe fed (one warp alone
//     reaches 80 % of the pipe, two 96 %: profiles/r01_notes.md).
#pragma once
#include "targets.cuh"

namespace mb200 {

__device__ __forceinline__ void dmma_m8n8k4(double& c0, double& c1, double a, double b) {
  asm volatile(
      "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
      : "+d"(c0), "+d"(c1)
      : "d"(a), "d"(b));
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void named_barrier_sync(int id, int nthreads) {
#if defined(MB200_EXP) && MB200_EXP == 4  // experiment: no group barriers (racy, results invalid)
  __syncwarp();
#else
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
#endif
}

constexpr int DMMA_TILES_PER_CTA = 7;
#ifndef MB200_DMMA_GROUPS
#define MB200_DMMA_GROUPS 4
#endif
// groups of 4 warps; the 7 row tiles are dealt 2,2,2,1 (4 groups), 2,2,1,1,1 (5), 2,1,1,1,1,1 (6)
// or 1 each (7): more groups = more warps per sub-partition to cover each other's serial
// phases, at fewer registers per thread and less B-fragment reuse
constexpr int DMMA_GROUPS = MB200_DMMA_GROUPS;
__host__ __device__ constexpr int dmma_tile_start(int g) {
  // first tile of group g: groups [0, 7 - G) own two tiles, the rest one
  return g <= 7 - DMMA_GROUPS ? 2 * g : 2 * (7 - DMMA_GROUPS) + (g - (7 - DMMA_GROUPS));
}
__host__ __device__ constexpr int dmma_tile_count(int g) { return g < 7 - DMMA_GROUPS ? 2 : 1; }
constexpr int DMMA_THREADS = 32 * 4 * DMMA_GROUPS;
constexpr int DMMA_ROWS_PER_CTA = 8 * DMMA_TILES_PER_CTA;  // 56 chains
constexpr int DMMA_MAX_RED = 4;

template <int DP>
struct DmmaSmem {
  static constexpr int LDA = DP + 4;  // row stride (doubles): rows shift by 32 B mod 128 B
  double A[DP * LDA];
  double P[DMMA_ROWS_PER_CTA * LDA];
  double part[4][DMMA_MAX_RED][DMMA_ROWS_PER_CTA];  // [warp-in-group][reduction][row]
  unsigned long long mbar;
};

// One group's work: MT (1 or 2) row tiles starting at CTA-local row `row0`.
//
// Formulation used inside the kernel (exact re-parametrisation of integrators.py:170-173):
//   s = dir * p   (signed momentum; the sign flip is exact)
//   kick:  s -= (eps/2) * grad l(q)          == dir * (p - (dir*eps/2) * grad)
//   drift: q += s . (eps * A)                == q + (dir*eps) * (A p)
// so the per-chain direction only appears in the load and the store, sm.A holds eps*A (scaled
// once after the TMA lands), and the drift is a DMMA whose accumulator operand IS q: positions
// never leave the accumulator registers and the only fp64-ALU work per coordinate and step is
// one FMA for the target's reduction and one per half-kick.  (Every fp64-ALU instruction costs
// DMMA issue slots on the shared pipe -- measured ~8 cycles each, profiles/r01_notes.md.)
template <class Target, int DP, int MT>
__device__ __forceinline__ void leapfrog_dmma_group(
    DmmaSmem<DP>& sm, const Target& target, const double* q_in, const double* p_in,
    double* q_out, double* p_out, const int32_t* __restrict__ dir, int64_t n_chains, int dim,
    double step_size, int n_steps, double* __restrict__ h_out, int32_t* __restrict__ status,
    int32_t* __restrict__ n_done, int64_t chain0, int row0, int w, int lane, int bar_id) {
  constexpr int LDA = DmmaSmem<DP>::LDA;
  constexpr int NT = DP / 32;  // 8-column tiles per warp
  constexpr int KS = DP / 4;   // k steps
  constexpr int NRED = Target::NRED;
  static_assert(NRED + 2 <= DMMA_MAX_RED, "too many reductions");
  const int r = lane >> 2, c = lane & 3;
  const int col0 = w * (DP / 4);  // first column of this warp's slice
  const double mh = -0.5 * step_size;

  // registers: positions of the slice in C-fragment layout (row 8mt + r, columns
  // col0 + 8nt + 2c + {0,1}); signed momenta live in sm.P with the same ownership
  double q[MT][NT][2], sgn[MT];
  bool live[MT];
  double2* pslot[MT];  // &sm.P[row][col0 + 2c]; + 4*nt double2 per column tile

#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
    const int64_t ch = chain0 + row0 + 8 * mt + r;
    live[mt] = ch < n_chains;
    sgn[mt] = (live[mt] && dir != nullptr && dir[ch] < 0) ? -1.0 : 1.0;
    pslot[mt] = reinterpret_cast<double2*>(&sm.P[(row0 + 8 * mt + r) * LDA + col0 + 2 * c]);
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const int i = col0 + 8 * nt + 2 * c;
      double2 a = make_double2(0.0, 0.0), b = make_double2(0.0, 0.0);
      if (live[mt] && i < dim) {  // dim is even on this path
        a = *reinterpret_cast<const double2*>(q_in + (size_t)ch * dim + i);
        b = *reinterpret_cast<const double2*>(p_in + (size_t)ch * dim + i);
      }
      q[mt][nt][0] = a.x, q[mt][nt][1] = a.y;
      pslot[mt][4 * nt] = make_double2(sgn[mt] * b.x, sgn[mt] * b.y);
    }
  }

  double red[MT][NRED + 1];

  // per-chain sum reductions of the target over the full row: partial over this warp's slice,
  // exchanged through shared memory.  The barrier also orders "all A-fragment reads of sm.P
  // done" before the in-place momentum update that follows.
  auto reduce_rows = [&]() {
    if (NRED > 0) {
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        // independent per-tile terms, then a pairwise tree: keeps the dependent fp64 chain short
        // (every dependent op queues behind other warps' DMMAs on the shared pipe)
        double term[NT][NRED + 1];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
          for (int k = 0; k < NRED; ++k) term[nt][k] = 0.0;
          target.accumulate(col0 + 8 * nt + 2 * c, q[mt][nt][0], q[mt][nt][1], term[nt]);
        }
#pragma unroll
        for (int k = 0; k < NRED; ++k) {
          if (NT == 4) red[mt][k] = (term[0][k] + term[1][k]) + (term[2 % NT][k] + term[3 % NT][k]);
          else if (NT == 3) red[mt][k] = (term[0][k] + term[1][k]) + term[2 % NT][k];
          else if (NT == 2) red[mt][k] = term[0][k] + term[1 % NT][k];
          else red[mt][k] = term[0][k];
        }
#pragma unroll
        for (int k = 0; k < NRED; ++k) {
          double v = red[mt][k];
          v += __shfl_xor_sync(FULL_MASK, v, 1);
          v += __shfl_xor_sync(FULL_MASK, v, 2);
          if (c == 0) sm.part[w][k][row0 + 8 * mt + r] = v;
        }
      }
    }
    named_barrier_sync(bar_id, 128);
    if (NRED > 0) {
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int k = 0; k < NRED; ++k) {
          const int row = row0 + 8 * mt + r;
          red[mt][k] = ((sm.part[0][k][row] + sm.part[1][k][row]) + sm.part[2][k][row]) +
                       sm.part[3][k][row];
        }
    }
  };

  // s -= (eps/2) * grad l(q), `kicks` times (1 or 2: the two half-steps either side of a step
  // boundary stay two separately rounded updates, systems.py:152), one FMA per coordinate and
  // kick; then make the new momenta visible to the group.
  auto kick_and_publish = [&](int kicks) {
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      const double ks = target.kick_scalar(red[mt], mh);
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const int i = col0 + 8 * nt + 2 * c;
        double2 pv = pslot[mt][4 * nt];
        if (i < dim) {
          if (kicks >= 1) target.kick_pair(i, q[mt][nt][0], q[mt][nt][1], red[mt], mh, ks, pv.x, pv.y);
          if (kicks >= 2) target.kick_pair(i, q[mt][nt][0], q[mt][nt][1], red[mt], mh, ks, pv.x, pv.y);
        }
        pslot[mt][4 * nt] = pv;
      }
    }
    named_barrier_sync(bar_id, 128);
  };

  // acc += S * (eps A) on the tensor pipe (acc = q for the drift, acc = 0 for the energy)
  auto drift = [&](double (&acc)[MT][NT][2]) {
    const double* a_base = &sm.P[(row0 + r) * LDA + c];
    const double* b_base = &sm.A[(col0 + r) * LDA + c];
#pragma unroll 8
    for (int j = 0; j < KS; ++j) {
      double a[MT], b[NT];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) a[mt] = a_base[mt * 8 * LDA + 4 * j];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) b[nt] = b_base[nt * 8 * LDA + 4 * j];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) dmma_m8n8k4(acc[mt][nt][0], acc[mt][nt][1], a[mt], b[nt]);
    }
  };

  reduce_rows();
  kick_and_publish(n_steps > 0 ? 1 : 0);
#ifndef MB200_EXP
#define MB200_EXP 0
#endif
  for (int s = 0; s < n_steps; ++s) {
    drift(q);  // h2_flow (systems.py:363): q += dir*eps * (A p)
#if MB200_EXP == 2 || MB200_EXP == 5 || MB200_EXP == 6  // experiment: drift only (results invalid)
    continue;
#endif
    reduce_rows();
#if MB200_EXP == 3  // experiment: one barrier per step (results invalid)
    continue;
#endif
    // closes step s and (cached gradient) opens step s+1
    kick_and_publish(s + 1 < n_steps ? 2 : 1);
  }

#if MB200_EXP == 8  // experiment 8: re-run the loop with per-phase cycle counters (CTA 0 only)
  if (h_out != nullptr && blockIdx.x == 0) {
    long long t_drift = 0, t_red = 0, t_kick = 0;
    for (int s = 0; s < n_steps; ++s) {
      const long long t0 = clock64();
      drift(q);
      const long long t1 = clock64();
      reduce_rows();
      const long long t2 = clock64();
      kick_and_publish(2);
      const long long t3 = clock64();
      t_drift += t1 - t0, t_red += t2 - t1, t_kick += t3 - t2;
    }
    if (lane == 0) {
      const int wid = (bar_id - 1) * 4 + w;
      double* dbg = h_out + 4096;  // debug area in the h buffer (read by profiles/tools/phase_c1.py)
      dbg[wid * 4 + 0] = (double)t_drift / n_steps;
      dbg[wid * 4 + 1] = (double)t_red / n_steps;
      dbg[wid * 4 + 2] = (double)t_kick / n_steps;
      dbg[wid * 4 + 3] = (double)MT;
    }
    return;
  }
#endif
  // ---- store (p = dir * s)
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
    const int64_t ch = chain0 + row0 + 8 * mt + r;
    if (!live[mt]) continue;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const int i = col0 + 8 * nt + 2 * c;
      if (i < dim) {
        *reinterpret_cast<double2*>(q_out + (size_t)ch * dim + i) =
            make_double2(q[mt][nt][0], q[mt][nt][1]);
        const double2 sv = pslot[mt][4 * nt];
        *reinterpret_cast<double2*>(p_out + (size_t)ch * dim + i) =
            make_double2(sgn[mt] * sv.x, sgn[mt] * sv.y);
      }
    }
    if (w == 0 && c == 0) {
      if (status != nullptr) status[ch] = MB200_STATUS_OK;
      if (n_done != nullptr) n_done[ch] = n_steps;
    }
  }

  // ---- Hamiltonian of the final state: l(q) + p . (A p) / 2   (systems.py:187-196, 348-350)
  if (h_out != nullptr) {
    double l[MT], kin[MT], u[MT][NT][2];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      l[mt] = 0.0;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const int i = col0 + 8 * nt + 2 * c;
        if (i < dim) l[mt] += target.nld_pair(i, q[mt][nt][0], q[mt][nt][1], red[mt]);
        u[mt][nt][0] = 0.0, u[mt][nt][1] = 0.0;
      }
    }
    drift(u);  // u = s . (eps A); sm.P holds the final momenta
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      kin[mt] = 0.0;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const double2 sv = pslot[mt][4 * nt];
        kin[mt] = fma(sv.x, u[mt][nt][0], kin[mt]);
        kin[mt] = fma(sv.y, u[mt][nt][1], kin[mt]);
      }
      kin[mt] += __shfl_xor_sync(FULL_MASK, kin[mt], 1);
      kin[mt] += __shfl_xor_sync(FULL_MASK, kin[mt], 2);
      l[mt] += __shfl_xor_sync(FULL_MASK, l[mt], 1);
      l[mt] += __shfl_xor_sync(FULL_MASK, l[mt], 2);
    }
    named_barrier_sync(bar_id, 128);  // every warp has consumed the gradient partials
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
      if (c == 0) {
        sm.part[w][0][row0 + 8 * mt + r] = kin[mt];
        sm.part[w][1][row0 + 8 * mt + r] = l[mt];
      }
    named_barrier_sync(bar_id, 128);
    if (w == 0 && c == 0) {
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        if (!live[mt]) continue;
        const int row = row0 + 8 * mt + r;
        const double ks = ((sm.part[0][0][row] + sm.part[1][0][row]) + sm.part[2][0][row]) +
                          sm.part[3][0][row];
        const double ls = ((sm.part[0][1][row] + sm.part[1][1][row]) + sm.part[2][1][row]) +
                          sm.part[3][1][row];
        h_out[chain0 + row] = ls + 0.5 * (ks / step_size);
      }
    }
  }
}

template <class Target, int DP>
__global__ void __launch_bounds__(DMMA_THREADS, 1)
    leapfrog_dmma_kernel(const double* q_in, const double* p_in, double* q_out, double* p_out,
                         const int32_t* __restrict__ dir, int64_t n_chains, int dim,
                         double step_size, int n_steps, const double* __restrict__ minv,
                         ModelArgs model, double* __restrict__ h_out,
                         int32_t* __restrict__ status, int32_t* __restrict__ n_done) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  DmmaSmem<DP>& sm = *reinterpret_cast<DmmaSmem<DP>*>(smem_raw);
  constexpr int LDA = DmmaSmem<DP>::LDA;
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const int group = warp >> 2;  // 4 groups: tiles {0,1}, {2,3}, {4,5}, {6}
  // Column quarter owned by this warp.  warp & 3 is its SM sub-partition; rotating the quarters by
  // the group index puts each group's "coordinate 0" warp (which carries the target's per-chain
  // special work, e.g. exp(-v) of the funnel) on a different sub-partition -- otherwise one
  // sub-partition is systematically slower and the other three idle at the group barriers.
  const int w = ((warp & 3) + group) & 3;
  const Target target(model, dim);

  // ---- stage A = M^-1 into shared memory: one TMA bulk copy per row, one mbarrier
  const uint32_t mbar = smem_u32(&sm.mbar);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // zero the padding (rows >= dim, columns >= dim) -- disjoint from the TMA destinations
  for (int idx = tid; idx < DP * LDA; idx += blockDim.x) {
    const int row = idx / LDA, col = idx - row * LDA;
    if (row >= dim || col >= dim) sm.A[idx] = 0.0;
  }
  for (int idx = tid; idx < DMMA_ROWS_PER_CTA * LDA; idx += blockDim.x) sm.P[idx] = 0.0;
  __syncthreads();
  if (warp == 0) {  // the 32 lanes of warp 0 issue the row copies (one TMA bulk copy per row)
    const uint32_t row_bytes = (uint32_t)dim * 8u;
    if (lane == 0)
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar),
                   "r"(row_bytes * (uint32_t)dim)
                   : "memory");
    __syncwarp();
#pragma unroll 1
    for (int row = lane; row < dim; row += 32) {
      const unsigned long long src =
          reinterpret_cast<unsigned long long>(minv) + (unsigned long long)row * row_bytes;
      asm volatile(
          "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
          ::"r"(smem_u32(&sm.A[row * LDA])),
          "l"(src), "r"(row_bytes), "r"(mbar)
          : "memory");
    }
  }
  // wait for the bytes to land (phase 0), then scale the staged metric: sm.A = eps * A
  {
    uint32_t done = 0;
    while (!done) {
      asm volatile(
          "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n"
          " selp.u32 %0, 1, 0, p;\n}"
          : "=r"(done)
          : "r"(mbar)
          : "memory");
    }
  }
  for (int idx = tid; idx < DP * LDA; idx += blockDim.x) sm.A[idx] = step_size * sm.A[idx];
  __syncthreads();

  for (int64_t blk = blockIdx.x; blk * DMMA_ROWS_PER_CTA < n_chains; blk += gridDim.x) {
    const int64_t chain0 = blk * DMMA_ROWS_PER_CTA;
    const int64_t left = n_chains - chain0;
    const int tiles = (int)((left >= DMMA_ROWS_PER_CTA) ? DMMA_TILES_PER_CTA : (left + 7) / 8);
    const int row0 = 8 * dmma_tile_start(group);
    int mt = tiles - dmma_tile_start(group);
    mt = mt > dmma_tile_count(group) ? dmma_tile_count(group) : mt;
#define MB200_GROUP(MT)                                                                       \
  leapfrog_dmma_group<Target, DP, MT>(sm, target, q_in, p_in, q_out, p_out, dir, n_chains,    \
                                      dim, step_size, n_steps, h_out, status, n_done, chain0, \
                                      row0, w, lane, 1 + group)
#if defined(MB200_EXP) && MB200_EXP == 5  // experiment: 2 warps per sub-partition
    if (group >= 2) mt = 0;
#endif
#if defined(MB200_EXP) && MB200_EXP == 6  // experiment: 1 warp per sub-partition
    if (group >= 1) mt = 0;
#endif
    if (mt == 2) MB200_GROUP(2);
    else if (mt == 1) MB200_GROUP(1);
#undef MB200_GROUP
    __syncthreads();  // next block of chains reuses sm.P / sm.part
  }
}

template <class Target, int DP>
static int launch_dmma(const double* q_in, const double* p_in, double* q_out, double* p_out,
                       const int32_t* dir, int64_t n, int dim, double eps, int n_steps,
                       const double* minv, const ModelArgs& m, double* h_out, int32_t* status,
                       int32_t* n_done, cudaStream_t st, int sms) {
  auto kern = leapfrog_dmma_kernel<Target, DP>;
  const size_t smem = sizeof(DmmaSmem<DP>);
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) !=
      cudaSuccess)
    return MB200_ERR_CUDA;
  int64_t blocks = (n + DMMA_ROWS_PER_CTA - 1) / DMMA_ROWS_PER_CTA;
  if (blocks > sms) blocks = sms;  // persistent: CTAs loop over blocks of 56 chains
  kern<<<(unsigned)blocks, DMMA_THREADS, smem, st>>>(q_in, p_in, q_out, p_out, dir, n, dim, eps, n_steps,
                                            minv, m, h_out, status, n_done);
  return 0;
}

template <class Target>
static int dispatch_dmma_dim(const double* q_in, const double* p_in, double* q_out, double* p_out,
                             const int32_t* dir, int64_t n, int dim, double eps, int n_steps,
                             const double* minv, const ModelArgs& m, double* h_out,
                             int32_t* status, int32_t* n_done, cudaStream_t st, int sms) {
#define MB200_DM(DP)                                                                           \
  return launch_dmma<Target, DP>(q_in, p_in, q_out, p_out, dir, n, dim, eps, n_steps, minv, m, \
                                 h_out, status, n_done, st, sms)
  if (dim <= 32) MB200_DM(32);
  if (dim <= 64) MB200_DM(64);
  if (dim <= 96) MB200_DM(96);
  MB200_DM(128);
#undef MB200_DM
}

// Returns MB200_ERR_UNSUPPORTED when the shape is outside this kernel's domain (the caller
// then uses the general-dimension kernel).
static int leapfrog_dmma_dispatch(const double* q_in, const double* p_in, double* q_out,
                                  double* p_out, const int32_t* dir, int64_t n, int dim,
                                  double eps, int n_steps, const double* minv, const ModelArgs& m,
                                  double* h_out, int32_t* status, int32_t* n_done,
                                  cudaStream_t st) {
  if (dim > 128 || (dim & 1) || dim < 8) return MB200_ERR_UNSUPPORTED;
  if (!(eps != 0.0) || !isfinite(eps)) return MB200_ERR_UNSUPPORTED;  // eps*A formulation
  if ((reinterpret_cast<uintptr_t>(minv) & 15) != 0) return MB200_ERR_UNSUPPORTED;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
#define MB200_ARGS \
  q_in, p_in, q_out, p_out, dir, n, dim, eps, n_steps, minv, m, h_out, status, n_done, st, sms
  switch (m.target_id) {
    case MB200_TARGET_STD_GAUSSIAN: return dispatch_dmma_dim<StdGaussianTarget>(MB200_ARGS);
    case MB200_TARGET_NEAL_FUNNEL: return dispatch_dmma_dim<NealFunnelTarget>(MB200_ARGS);
    case MB200_TARGET_BANANA: return dispatch_dmma_dim<BananaTarget>(MB200_ARGS);
    default: return MB200_ERR_UNSUPPORTED;
  }
#undef MB200_ARGS
}

}  // namespace mb200
