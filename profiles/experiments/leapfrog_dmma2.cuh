// K1 (second generation): same arithmetic and shared-memory layout as leapfrog_dmma.cuh, but each
// warp owns TWO independent sets of row tiles that run half a step out of phase, and the group
// barriers are split-phase mbarriers (arrive ... work ... wait):
//
//     drift(A, s)  interleaved with  reduce / kick / publish of B (closing step s-1)
//     drift(B, s)  interleaved with  reduce / kick / publish of A (closing step s)
//
// A warp therefore never blocks on its group between a drift and the next: while the per-chain
// reduction of one set travels through shared memory, the warp issues the DMMAs of the other
// set.  (Measured with the first-generation kernel: a third of every warp's step was spent in
// the serial reduce -> kick -> publish phase and at its two group barriers, leaving the FP64
// tensor pipe 20 % idle: profiles/r01_notes.md.)
//
// CTA: 256 threads = 2 groups x 4 warps, one CTA per SM, 56 chains = 7 row tiles:
//   group 0: set A = tiles {0,1}, set B = tiles {2,3};  group 1: set A = {4,5}, set B = {6}
// -> 7 tile-quarters per SM sub-partition, as before.
#pragma once
#include <type_traits>

#include "leapfrog_dmma.cuh"

namespace mb200 {

struct SplitBarrier {
  uint32_t addr;
  uint32_t uses;  // completed waits: the next wait is for phase `uses`
  __device__ __forceinline__ void arrive(int lane) const {
    __syncwarp();
    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(addr) : "memory");
  }
  __device__ __forceinline__ void wait() {
    uint32_t done = 0;
    const uint32_t parity = uses & 1u;
    while (!done) {
      asm volatile(
          "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
          " selp.u32 %0, 1, 0, p;\n}"
          : "=r"(done)
          : "r"(addr), "r"(parity)
          : "memory");
    }
    ++uses;
  }
};

template <int DP>
struct Dmma2Smem {
  static constexpr int LDA = DP + 4;
  double A[DP * LDA];
  double P[DMMA_ROWS_PER_CTA * LDA];
  double part[4][DMMA_MAX_RED][DMMA_ROWS_PER_CTA];
  unsigned long long mbar;           // TMA
  unsigned long long gbar[2][2][2];  // [group][set][reduce / publish]
};

// register state of one set of MT row tiles (MT may be 0: empty set)
template <int MT, int NT, int NRED>
struct DmmaSet {
  double q[MT > 0 ? MT : 1][NT][2];
  double red[MT > 0 ? MT : 1][NRED + 1];
  double sgn[MT > 0 ? MT : 1];
  bool live[MT > 0 ? MT : 1];
  double2* pslot[MT > 0 ? MT : 1];
  int row0;
};

template <class Target, int DP, int MTA, int MTB>
__device__ __forceinline__ void leapfrog_dmma2_group(
    Dmma2Smem<DP>& sm, const Target& target, const double* q_in, const double* p_in,
    double* q_out, double* p_out, const int32_t* __restrict__ dir, int64_t n_chains, int dim,
    double step_size, int n_steps, double* __restrict__ h_out, int32_t* __restrict__ status,
    int32_t* __restrict__ n_done, int64_t chain0, int group, int w, int lane) {
  constexpr int LDA = Dmma2Smem<DP>::LDA;
  constexpr int NT = DP / 32;
  constexpr int KS = DP / 4;
  constexpr int KC = KS / 4;  // k steps per drift chunk
  constexpr int NRED = Target::NRED;
  constexpr bool HAS_B = MTB > 0;
  static_assert(NRED + 2 <= DMMA_MAX_RED, "too many reductions");
  const int r = lane >> 2, c = lane & 3;
  const int col0 = w * (DP / 4);
  const double mh = -0.5 * step_size;

  DmmaSet<MTA, NT, NRED> sa;
  DmmaSet<MTB, NT, NRED> sb;
  sa.row0 = group * 32;
  sb.row0 = group * 32 + 16;
  SplitBarrier b1a{smem_u32(&sm.gbar[group][0][0]), 0}, b2a{smem_u32(&sm.gbar[group][0][1]), 0};
  SplitBarrier b1b{smem_u32(&sm.gbar[group][1][0]), 0}, b2b{smem_u32(&sm.gbar[group][1][1]), 0};

  auto load = [&](auto& s, auto mt_tag) {
    constexpr int MT = decltype(mt_tag)::value;
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      const int64_t ch = chain0 + s.row0 + 8 * mt + r;
      s.live[mt] = ch < n_chains;
      s.sgn[mt] = (s.live[mt] && dir != nullptr && dir[ch] < 0) ? -1.0 : 1.0;
      s.pslot[mt] = reinterpret_cast<double2*>(&sm.P[(s.row0 + 8 * mt + r) * LDA + col0 + 2 * c]);
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const int i = col0 + 8 * nt + 2 * c;
        double2 a = make_double2(0.0, 0.0), b = make_double2(0.0, 0.0);
        if (s.live[mt] && i < dim) {
          a = *reinterpret_cast<const double2*>(q_in + (size_t)ch * dim + i);
          b = *reinterpret_cast<const double2*>(p_in + (size_t)ch * dim + i);
        }
        s.q[mt][nt][0] = a.x, s.q[mt][nt][1] = a.y;
        s.pslot[mt][4 * nt] = make_double2(s.sgn[mt] * b.x, s.sgn[mt] * b.y);
      }
    }
  };

  // partial per-chain reductions of the target over this warp's column slice -> shared memory
  auto reduce_part = [&](auto& s, auto mt_tag) {
    constexpr int MT = decltype(mt_tag)::value;
    if (NRED > 0) {
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        double term[NT][NRED + 1];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
          for (int k = 0; k < NRED; ++k) term[nt][k] = 0.0;
          target.accumulate(col0 + 8 * nt + 2 * c, s.q[mt][nt][0], s.q[mt][nt][1], term[nt]);
        }
#pragma unroll
        for (int k = 0; k < NRED; ++k) {
          double v;
          if (NT == 4) v = (term[0][k] + term[1][k]) + (term[2 % NT][k] + term[3 % NT][k]);
          else if (NT == 3) v = (term[0][k] + term[1][k]) + term[2 % NT][k];
          else if (NT == 2) v = term[0][k] + term[1 % NT][k];
          else v = term[0][k];
          v += __shfl_xor_sync(FULL_MASK, v, 1);
          v += __shfl_xor_sync(FULL_MASK, v, 2);
          if (c == 0) sm.part[w][k][s.row0 + 8 * mt + r] = v;
        }
      }
    }
  };

  // after the reduce barrier: total the partials, kick the signed momenta in place (`kicks`
  // separately rounded half-kicks, one FMA each), which also publishes them for the next drift
  auto finish = [&](auto& s, auto mt_tag, int kicks) {
    constexpr int MT = decltype(mt_tag)::value;
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      const int row = s.row0 + 8 * mt + r;
#pragma unroll
      for (int k = 0; k < NRED; ++k)
        s.red[mt][k] = ((sm.part[0][k][row] + sm.part[1][k][row]) + sm.part[2][k][row]) +
                       sm.part[3][k][row];
      const double ks = target.kick_scalar(s.red[mt], mh);
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const int i = col0 + 8 * nt + 2 * c;
        double2 pv = s.pslot[mt][4 * nt];
        if (i < dim) {
          if (kicks >= 1)
            target.kick_pair(i, s.q[mt][nt][0], s.q[mt][nt][1], s.red[mt], mh, ks, pv.x, pv.y);
          if (kicks >= 2)
            target.kick_pair(i, s.q[mt][nt][0], s.q[mt][nt][1], s.red[mt], mh, ks, pv.x, pv.y);
        }
        s.pslot[mt][4 * nt] = pv;
      }
    }
  };

  // acc += S (eps A) for k steps [j0, j1) on the tensor pipe
  auto drift_chunk = [&](auto& acc, int row0, auto mt_tag, int j0, int j1) {
    constexpr int MT = decltype(mt_tag)::value;
    const double* a_base = &sm.P[(row0 + r) * LDA + c];
    const double* b_base = &sm.A[(col0 + r) * LDA + c];
#pragma unroll
    for (int j = j0; j < j1; ++j) {
      double a[MT > 0 ? MT : 1], b[NT];
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

  using TA = std::integral_constant<int, MTA>;
  using TB = std::integral_constant<int, MTB>;

  // ---- prologue: load, first half-kick
  load(sa, TA{});
  if (HAS_B) load(sb, TB{});
  reduce_part(sa, TA{});
  b1a.arrive(lane);
  if (HAS_B) {
    reduce_part(sb, TB{});
    b1b.arrive(lane);
  }
  b1a.wait();
  finish(sa, TA{}, n_steps > 0 ? 1 : 0);
  b2a.arrive(lane);
  if (HAS_B) {
    b1b.wait();
    finish(sb, TB{}, n_steps > 0 ? 1 : 0);
    b2b.arrive(lane);
  }

  for (int s = 0; s < n_steps; ++s) {
    // ---- drift(A, s) || close step s-1 of B
    b2a.wait();
    drift_chunk(sa.q, sa.row0, TA{}, 0, KC);
    if (HAS_B && s > 0) {
      reduce_part(sb, TB{});
      b1b.arrive(lane);
    }
    drift_chunk(sa.q, sa.row0, TA{}, KC, 2 * KC);
    if (HAS_B && s > 0) {
      b1b.wait();
      finish(sb, TB{}, 2);
      b2b.arrive(lane);
    }
    drift_chunk(sa.q, sa.row0, TA{}, 2 * KC, KS);
    // ---- drift(B, s) || close step s of A
    if (HAS_B) {
      b2b.wait();
      drift_chunk(sb.q, sb.row0, TB{}, 0, KC);
      reduce_part(sa, TA{});
      b1a.arrive(lane);
      drift_chunk(sb.q, sb.row0, TB{}, KC, 2 * KC);
      b1a.wait();
      finish(sa, TA{}, s + 1 < n_steps ? 2 : 1);
      b2a.arrive(lane);
      drift_chunk(sb.q, sb.row0, TB{}, 2 * KC, KS);
    } else {
      reduce_part(sa, TA{});
      b1a.arrive(lane);
      b1a.wait();
      finish(sa, TA{}, s + 1 < n_steps ? 2 : 1);
      b2a.arrive(lane);
    }
  }
  if (HAS_B && n_steps > 0) {  // close the last step of B
    reduce_part(sb, TB{});
    b1b.arrive(lane);
    b1b.wait();
    finish(sb, TB{}, 1);
    b2b.arrive(lane);
  }
  b2a.wait();  // final momenta of the whole group are in sm.P
  if (HAS_B) b2b.wait();

  auto store = [&](auto& s, auto mt_tag) {
    constexpr int MT = decltype(mt_tag)::value;
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      const int64_t ch = chain0 + s.row0 + 8 * mt + r;
      if (!s.live[mt]) continue;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const int i = col0 + 8 * nt + 2 * c;
        if (i < dim) {
          *reinterpret_cast<double2*>(q_out + (size_t)ch * dim + i) =
              make_double2(s.q[mt][nt][0], s.q[mt][nt][1]);
          const double2 sv = s.pslot[mt][4 * nt];
          *reinterpret_cast<double2*>(p_out + (size_t)ch * dim + i) =
              make_double2(s.sgn[mt] * sv.x, s.sgn[mt] * sv.y);
        }
      }
      if (w == 0 && c == 0) {
        if (status != nullptr) status[ch] = MB200_STATUS_OK;
        if (n_done != nullptr) n_done[ch] = n_steps;
      }
    }
  };
  store(sa, TA{});
  if (HAS_B) store(sb, TB{});

  // ---- Hamiltonian of the final state: l(q) + p.(A p)/2  (systems.py:187-196, 348-350)
  if (h_out != nullptr) {
    auto energy = [&](auto& s, auto mt_tag, SplitBarrier& bx, SplitBarrier& by) {
      constexpr int MT = decltype(mt_tag)::value;
      double u[MT > 0 ? MT : 1][NT][2], l[MT > 0 ? MT : 1], kin[MT > 0 ? MT : 1];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        l[mt] = 0.0;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          const int i = col0 + 8 * nt + 2 * c;
          if (i < dim) l[mt] += target.nld_pair(i, s.q[mt][nt][0], s.q[mt][nt][1], s.red[mt]);
          u[mt][nt][0] = 0.0, u[mt][nt][1] = 0.0;
        }
      }
      drift_chunk(u, s.row0, mt_tag, 0, KS);
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        kin[mt] = 0.0;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          const double2 sv = s.pslot[mt][4 * nt];
          kin[mt] = fma(sv.x, u[mt][nt][0], kin[mt]);
          kin[mt] = fma(sv.y, u[mt][nt][1], kin[mt]);
        }
        kin[mt] += __shfl_xor_sync(FULL_MASK, kin[mt], 1);
        kin[mt] += __shfl_xor_sync(FULL_MASK, kin[mt], 2);
        l[mt] += __shfl_xor_sync(FULL_MASK, l[mt], 1);
        l[mt] += __shfl_xor_sync(FULL_MASK, l[mt], 2);
        if (c == 0) {
          sm.part[w][0][s.row0 + 8 * mt + r] = kin[mt];
          sm.part[w][1][s.row0 + 8 * mt + r] = l[mt];
        }
      }
      bx.arrive(lane);
      bx.wait();
      if (w == 0 && c == 0) {
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          if (!s.live[mt]) continue;
          const int row = s.row0 + 8 * mt + r;
          const double ks = ((sm.part[0][0][row] + sm.part[1][0][row]) + sm.part[2][0][row]) +
                            sm.part[3][0][row];
          const double ls = ((sm.part[0][1][row] + sm.part[1][1][row]) + sm.part[2][1][row]) +
                            sm.part[3][1][row];
          h_out[chain0 + row] = ls + 0.5 * (ks / step_size);
        }
      }
      (void)by;
    };
    energy(sa, TA{}, b1a, b2a);
    if (HAS_B) energy(sb, TB{}, b1b, b2b);
  }
}

template <class Target, int DP>
__global__ void __launch_bounds__(256, 1)
    leapfrog_dmma2_kernel(const double* q_in, const double* p_in, double* q_out, double* p_out,
                          const int32_t* __restrict__ dir, int64_t n_chains, int dim,
                          double step_size, int n_steps, const double* __restrict__ minv,
                          ModelArgs model, double* __restrict__ h_out,
                          int32_t* __restrict__ status, int32_t* __restrict__ n_done) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Dmma2Smem<DP>& sm = *reinterpret_cast<Dmma2Smem<DP>*>(smem_raw);
  constexpr int LDA = Dmma2Smem<DP>::LDA;
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const int group = warp >> 2;
  const int w = ((warp & 3) + group) & 3;  // column quarter (rotated per group, see v1)
  const Target target(model, dim);

  const uint32_t mbar = smem_u32(&sm.mbar);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int idx = tid; idx < DP * LDA; idx += blockDim.x) {
    const int row = idx / LDA, col = idx - row * LDA;
    if (row >= dim || col >= dim) sm.A[idx] = 0.0;
  }
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

  for (int64_t blk = blockIdx.x; blk * DMMA_ROWS_PER_CTA < n_chains; blk += gridDim.x) {
    // (re)initialise the per-block state: momentum tile and the split-phase group barriers
    __syncthreads();
    for (int idx = tid; idx < DMMA_ROWS_PER_CTA * LDA; idx += blockDim.x) sm.P[idx] = 0.0;
    if (tid < 8) {
      const uint32_t gb = smem_u32(&sm.gbar[tid >> 2][(tid >> 1) & 1][tid & 1]);
      if (blk != (int64_t)blockIdx.x) asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(gb));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 4;" ::"r"(gb));
    }
    if (tid == 0) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();

    const int64_t chain0 = blk * DMMA_ROWS_PER_CTA;
    const int64_t left = n_chains - chain0;
    const int tiles = (int)((left >= DMMA_ROWS_PER_CTA) ? DMMA_TILES_PER_CTA : (left + 7) / 8);
    int tg = tiles - 4 * group;
    tg = tg < 0 ? 0 : (tg > 4 ? 4 : tg);
    const int mta = tg < 2 ? tg : 2, mtb = tg - mta;
#define MB200_GROUP2(MA, MB_)                                                                  \
  leapfrog_dmma2_group<Target, DP, MA, MB_>(sm, target, q_in, p_in, q_out, p_out, dir,         \
                                            n_chains, dim, step_size, n_steps, h_out, status, \
                                            n_done, chain0, group, w, lane)
    if (mta == 2 && mtb == 2) MB200_GROUP2(2, 2);
    else if (mta == 2 && mtb == 1) MB200_GROUP2(2, 1);
    else if (mta == 2 && mtb == 0) MB200_GROUP2(2, 0);
    else if (mta == 1) MB200_GROUP2(1, 0);
#undef MB200_GROUP2
  }
}

template <class Target, int DP>
static int launch_dmma2(const double* q_in, const double* p_in, double* q_out, double* p_out,
                        const int32_t* dir, int64_t n, int dim, double eps, int n_steps,
                        const double* minv, const ModelArgs& m, double* h_out, int32_t* status,
                        int32_t* n_done, cudaStream_t st, int sms) {
  auto kern = leapfrog_dmma2_kernel<Target, DP>;
  const size_t smem = sizeof(Dmma2Smem<DP>);
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) !=
      cudaSuccess)
    return MB200_ERR_CUDA;
  int64_t blocks = (n + DMMA_ROWS_PER_CTA - 1) / DMMA_ROWS_PER_CTA;
  if (blocks > sms) blocks = sms;
  kern<<<(unsigned)blocks, 256, smem, st>>>(q_in, p_in, q_out, p_out, dir, n, dim, eps, n_steps,
                                            minv, m, h_out, status, n_done);
  return 0;
}

template <class Target>
static int dispatch_dmma2_dim(const double* q_in, const double* p_in, double* q_out,
                              double* p_out, const int32_t* dir, int64_t n, int dim, double eps,
                              int n_steps, const double* minv, const ModelArgs& m, double* h_out,
                              int32_t* status, int32_t* n_done, cudaStream_t st, int sms) {
#define MB200_DM(DP)                                                                            \
  return launch_dmma2<Target, DP>(q_in, p_in, q_out, p_out, dir, n, dim, eps, n_steps, minv, m, \
                                  h_out, status, n_done, st, sms)
  if (dim <= 32) MB200_DM(32);
  if (dim <= 64) MB200_DM(64);
  if (dim <= 96) MB200_DM(96);
  MB200_DM(128);
#undef MB200_DM
}

static int leapfrog_dmma2_dispatch(const double* q_in, const double* p_in, double* q_out,
                                   double* p_out, const int32_t* dir, int64_t n, int dim,
                                   double eps, int n_steps, const double* minv,
                                   const ModelArgs& m, double* h_out, int32_t* status,
                                   int32_t* n_done, cudaStream_t st) {
  if (dim > 128 || (dim & 1) || dim < 8) return MB200_ERR_UNSUPPORTED;
  if (!(eps != 0.0) || !isfinite(eps)) return MB200_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(minv) & 15) != 0) return MB200_ERR_UNSUPPORTED;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
#define MB200_ARGS \
  q_in, p_in, q_out, p_out, dir, n, dim, eps, n_steps, minv, m, h_out, status, n_done, st, sms
  switch (m.target_id) {
    case MB200_TARGET_STD_GAUSSIAN: return dispatch_dmma2_dim<StdGaussianTarget>(MB200_ARGS);
    case MB200_TARGET_NEAL_FUNNEL: return dispatch_dmma2_dim<NealFunnelTarget>(MB200_ARGS);
    case MB200_TARGET_BANANA: return dispatch_dmma2_dim<BananaTarget>(MB200_ARGS);
    default: return MB200_ERR_UNSUPPORTED;
  }
#undef MB200_ARGS
}

}  // namespace mb200
