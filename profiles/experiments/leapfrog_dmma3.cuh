// K1, warp-specialised generation: the arithmetic and shared-memory metric of leapfrog_dmma.cuh,
// with the two kinds of work of a leapfrog step given to two kinds of warps:
//
//   * 8 DRIFT warps (2 per SM sub-partition) do nothing but  q += s (eps A)  on the FP64 tensor
//     pipe.  Each owns a 32-column slice of TWO sets of row tiles whose steps are half a step out
//     of phase, so it always has a drift to issue; it never executes the per-chain reductions or
//     the momentum kicks and never waits on a CTA-wide or group-wide barrier.
//   * 8 UPDATE warps (2 per sub-partition) own whole rows (one row tile = 8 chains each): they
//     read the freshly drifted positions of a tile from a small export slot, evaluate the target's
//     per-chain reductions with warp shuffles (a row is 128 numbers = 4 per lane, no cross-warp
//     exchange), apply the half-kicks to the signed momenta in shared memory and hand the tile
//     back.
//   * hand-over is by split-phase mbarriers per tile set: q_ready (4 drift warps arrive, update
//     warps wait) and p_ready (update warps arrive, drift warps wait).  While set A is being
//     updated the drift warps are busy with set B, so the serial reduce -> kick chain (a third of a
//     step in the first-generation kernel) is off the tensor pipe's critical path.
//   * registers are re-partitioned with setmaxnreg: drift warps 200 (positions of both sets stay
//     in DMMA accumulators), update warps 56.
//
// CTA: 512 threads, one per SM, 56 chains = 7 row tiles: group 0 sets {0,1},{2,3}; group 1
// sets {4,5},{6} -- 7 tile-quarters per sub-partition, as in the other generations.
#pragma once
#include <type_traits>

#include "leapfrog_dmma2.cuh"

namespace mb200 {

template <int DP>
struct Dmma3Smem {
  static constexpr int LDA = DP + 4;
  double A[DP * LDA];                   // eps * M^-1
  double P[DMMA_ROWS_PER_CTA * LDA];    // signed momenta s = dir * p
  double Q[2][16 * LDA];                // per-group export slot: positions of the set just drifted
  double lrow[DMMA_ROWS_PER_CTA];       // l(q) per row of the final state (energy output)
  double kpart[4][DMMA_ROWS_PER_CTA];   // kinetic-energy partials per column quarter
  unsigned long long mbar;              // TMA
  unsigned long long qbar[2][2];        // [group][set] positions exported
  unsigned long long pbar[2][2];        // [group][set] momenta kicked
};

// ------------------------------------------------------------------------------ drift role
template <class Target, int DP, int MTA, int MTB>
__device__ __forceinline__ void dmma3_drift_role(
    Dmma3Smem<DP>& sm, const double* q_in, const double* p_in, double* q_out, double* p_out,
    const int32_t* __restrict__ dir, int64_t n_chains, int dim, double step_size, int n_steps,
    double* __restrict__ h_out, int32_t* __restrict__ status, int32_t* __restrict__ n_done,
    int64_t chain0, int group, int w, int lane) {
  constexpr int LDA = Dmma3Smem<DP>::LDA;
  constexpr int NT = DP / 32;
  constexpr int KS = DP / 4;
  constexpr bool HAS_B = MTB > 0;
  const int r = lane >> 2, c = lane & 3;
  const int col0 = w * (DP / 4);

  DmmaSet<MTA, NT, 0> sa;
  DmmaSet<MTB, NT, 0> sb;
  sa.row0 = group * 32;
  sb.row0 = group * 32 + 16;
  SplitBarrier qa{smem_u32(&sm.qbar[group][0]), 0}, pa{smem_u32(&sm.pbar[group][0]), 0};
  SplitBarrier qb{smem_u32(&sm.qbar[group][1]), 0}, pb{smem_u32(&sm.pbar[group][1]), 0};
  double* slot = sm.Q[group];

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
  // copy the positions of a set into the group's export slot (rows 8*mt + r of the slot)
  auto export_q = [&](auto& s, auto mt_tag) {
    constexpr int MT = decltype(mt_tag)::value;
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
        *reinterpret_cast<double2*>(&slot[(8 * mt + r) * LDA + col0 + 8 * nt + 2 * c]) =
            make_double2(s.q[mt][nt][0], s.q[mt][nt][1]);
  };
  auto drift = [&](auto& acc, int row0, auto mt_tag) {
    constexpr int MT = decltype(mt_tag)::value;
    const double* a_base = &sm.P[(row0 + r) * LDA + c];
    const double* b_base = &sm.A[(col0 + r) * LDA + c];
#pragma unroll 8
    for (int j = 0; j < KS; ++j) {
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

  // ---- prologue: hand the initial positions to the update warps (first half-kick)
  load(sa, TA{});
  if (HAS_B) load(sb, TB{});
  export_q(sa, TA{});
  qa.arrive(lane);
  pa.wait();  // A kicked; its update warps are done with the export slot
  if (HAS_B) {
    export_q(sb, TB{});
    qb.arrive(lane);
  }
  for (int s = 0; s < n_steps; ++s) {
    drift(sa.q, sa.row0, TA{});
    if (HAS_B) {
      pb.wait();  // B kicked (closing step s-1 / opening step s); slot free again
      export_q(sa, TA{});
      qa.arrive(lane);
      drift(sb.q, sb.row0, TB{});
      pa.wait();  // A kicked for step s+1 (or closed, after the last step)
      export_q(sb, TB{});
      qb.arrive(lane);
    } else {
      export_q(sa, TA{});
      qa.arrive(lane);
      pa.wait();
    }
  }
  if (HAS_B) pb.wait();  // the last kick of B

  // ---- store (p = dir * s)
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

  // ---- Hamiltonian: l(q) (left in sm.lrow by the update warps) + s.(eps A)s / (2 eps)
  if (h_out != nullptr) {
    auto kinetic = [&](auto& s, auto mt_tag) {
      constexpr int MT = decltype(mt_tag)::value;
      double u[MT > 0 ? MT : 1][NT][2];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) u[mt][nt][0] = 0.0, u[mt][nt][1] = 0.0;
      drift(u, s.row0, mt_tag);
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        double kin = 0.0;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          const double2 sv = s.pslot[mt][4 * nt];
          kin = fma(sv.x, u[mt][nt][0], kin);
          kin = fma(sv.y, u[mt][nt][1], kin);
        }
        kin += __shfl_xor_sync(FULL_MASK, kin, 1);
        kin += __shfl_xor_sync(FULL_MASK, kin, 2);
        if (c == 0) sm.kpart[w][s.row0 + 8 * mt + r] = kin;
      }
    };
    kinetic(sa, TA{});
    if (HAS_B) kinetic(sb, TB{});
    named_barrier_sync(1 + group, 128);  // the 4 drift warps of the group
    auto emit = [&](auto& s, auto mt_tag) {
      constexpr int MT = decltype(mt_tag)::value;
      if (w == 0 && c == 0) {
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          if (!s.live[mt]) continue;
          const int row = s.row0 + 8 * mt + r;
          const double ks = ((sm.kpart[0][row] + sm.kpart[1][row]) + sm.kpart[2][row]) +
                            sm.kpart[3][row];
          h_out[chain0 + row] = sm.lrow[row] + 0.5 * (ks / step_size);
        }
      }
    };
    emit(sa, TA{});
    if (HAS_B) emit(sb, TB{});
  }
}

// ----------------------------------------------------------------------------- update role
// One warp per row tile (8 chains), all 8 rows at once: lane = 4*row + cq owns the coordinate
// pairs cq + 4k (k = 0..DP/8-1) of its row, so the per-chain reductions are two xor-shuffles over
// the 4 lanes of a row and the target's special per-chain work (the funnel's exp(-v)) is evaluated
// for the 8 chains in one instruction stream.  Positions are streamed twice from the export slot
// (reduction pass, kick pass) instead of being held in registers (56 registers per thread).
template <class Target, int DP>
__device__ __forceinline__ void dmma3_update_role(Dmma3Smem<DP>& sm, const Target& target,
                                                  int dim, double step_size, int n_steps,
                                                  bool want_energy, int group, int set,
                                                  int tile_in_set, int lane) {
  constexpr int LDA = Dmma3Smem<DP>::LDA;
  constexpr int NK = DP / 8;  // pairs per lane
  constexpr int NRED = Target::NRED;
  const double mh = -0.5 * step_size;
  SplitBarrier qbar{smem_u32(&sm.qbar[group][set]), 0};
  const uint32_t pbar = smem_u32(&sm.pbar[group][set]);
  const int rr = lane >> 2, cq = lane & 3;
  const int row = group * 32 + set * 16 + 8 * tile_in_set + rr;  // CTA-local row
  const double2* qrow =
      reinterpret_cast<const double2*>(&sm.Q[group][(8 * tile_in_set + rr) * LDA]) + cq;
  double2* prow = reinterpret_cast<double2*>(&sm.P[row * LDA]) + cq;

  for (int phase = 0; phase <= n_steps; ++phase) {
    const int kicks = (n_steps == 0) ? 0 : ((phase == 0 || phase == n_steps) ? 1 : 2);
    qbar.wait();
    // ---- pass 1: per-chain reductions (four independent partial sets, then a tree)
    double red[NRED + 1];
    if (NRED > 0) {
      double part[4][NRED + 1];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int k = 0; k < NRED; ++k) part[a][k] = 0.0;
#pragma unroll
      for (int k = 0; k < NK; ++k) {
        const int i = 2 * (cq + 4 * k);
        if (i < dim) {
          const double2 qv = qrow[4 * k];
          target.accumulate(i, qv.x, qv.y, part[k & 3]);
        }
      }
#pragma unroll
      for (int k = 0; k < NRED; ++k) {
        double v = (part[0][k] + part[1][k]) + (part[2][k] + part[3][k]);
        v += __shfl_xor_sync(FULL_MASK, v, 1);
        v += __shfl_xor_sync(FULL_MASK, v, 2);
        red[k] = v;
      }
    }
    // ---- pass 2: half-kicks of the signed momenta in place (+ l(q) of the final state)
    const double ks = target.kick_scalar(red, mh);
    double l = 0.0;
#pragma unroll
    for (int k = 0; k < NK; ++k) {
      const int i = 2 * (cq + 4 * k);
      if (i < dim) {
        const double2 qv = qrow[4 * k];
        double2 pv = prow[4 * k];
        if (kicks >= 1) target.kick_pair(i, qv.x, qv.y, red, mh, ks, pv.x, pv.y);
        if (kicks >= 2) target.kick_pair(i, qv.x, qv.y, red, mh, ks, pv.x, pv.y);
        prow[4 * k] = pv;
        if (want_energy && phase == n_steps) l += target.nld_pair(i, qv.x, qv.y, red);
      }
    }
    if (want_energy && phase == n_steps) {
      l += __shfl_xor_sync(FULL_MASK, l, 1);
      l += __shfl_xor_sync(FULL_MASK, l, 2);
      if (cq == 0) sm.lrow[row] = l;
    }
    __syncwarp();
    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(pbar) : "memory");
  }
}

template <class Target, int DP>
__global__ void __launch_bounds__(512, 1)
    leapfrog_dmma3_kernel(const double* q_in, const double* p_in, double* q_out, double* p_out,
                          const int32_t* __restrict__ dir, int64_t n_chains, int dim,
                          double step_size, int n_steps, const double* __restrict__ minv,
                          ModelArgs model, double* __restrict__ h_out,
                          int32_t* __restrict__ status, int32_t* __restrict__ n_done) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Dmma3Smem<DP>& sm = *reinterpret_cast<Dmma3Smem<DP>*>(smem_raw);
  constexpr int LDA = Dmma3Smem<DP>::LDA;
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const Target target(model, dim);

  // ---- stage eps * A (as in the other generations)
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
  if (warp == 0) {
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

  const int64_t n_blocks = (n_chains + DMMA_ROWS_PER_CTA - 1) / DMMA_ROWS_PER_CTA;
  // per-block geometry (same for both roles)
  auto tiles_of = [&](int64_t blk) {
    const int64_t left = n_chains - blk * DMMA_ROWS_PER_CTA;
    return (int)((left >= DMMA_ROWS_PER_CTA) ? DMMA_TILES_PER_CTA : (left + 7) / 8);
  };

  if (warp < 8) {
    // ======================= DRIFT warps =======================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 200;");
    const int group = warp >> 2;
    const int w = ((warp & 3) + group) & 3;
    for (int64_t blk = blockIdx.x; blk < n_blocks; blk += gridDim.x) {
      __syncthreads();
      for (int idx = tid; idx < DMMA_ROWS_PER_CTA * LDA; idx += 256) sm.P[idx] = 0.0;
      const int tiles = tiles_of(blk);
      if (tid < 8) {  // [group][set] x {q, p} barriers
        const int g = tid >> 2, st = (tid >> 1) & 1, which = tid & 1;
        int tg = tiles - 4 * g;
        tg = tg < 0 ? 0 : (tg > 4 ? 4 : tg);
        const int mta = tg < 2 ? tg : 2, mtb = tg - mta;
        const int mt = st == 0 ? mta : mtb;
        const uint32_t b = smem_u32(which == 0 ? &sm.qbar[g][st] : &sm.pbar[g][st]);
        if (blk != (int64_t)blockIdx.x) asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(b));
        // q_ready: the 4 drift warps of the group; p_ready: one update warp per tile of the set
        const int count = which == 0 ? 4 : (mt > 0 ? mt : 1);
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(count));
      }
      if (tid == 0) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      __syncthreads();
      int tg = tiles - 4 * group;
      tg = tg < 0 ? 0 : (tg > 4 ? 4 : tg);
      const int mta = tg < 2 ? tg : 2, mtb = tg - mta;
      const int64_t chain0 = blk * DMMA_ROWS_PER_CTA;
#define MB200_DRIFT(MA, MB_)                                                                   \
  dmma3_drift_role<Target, DP, MA, MB_>(sm, q_in, p_in, q_out, p_out, dir, n_chains, dim,      \
                                        step_size, n_steps, h_out, status, n_done, chain0,     \
                                        group, w, lane)
      if (mta == 2 && mtb == 2) MB200_DRIFT(2, 2);
      else if (mta == 2 && mtb == 1) MB200_DRIFT(2, 1);
      else if (mta == 2 && mtb == 0) MB200_DRIFT(2, 0);
      else if (mta == 1) MB200_DRIFT(1, 0);
#undef MB200_DRIFT
    }
  } else {
    // ======================= UPDATE warps =======================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    const int u = warp - 8;           // 0..7
    const int group = u >> 2;         // 0,1
    const int set = (u >> 1) & 1;     // A, B
    const int tile_in_set = u & 1;
    for (int64_t blk = blockIdx.x; blk < n_blocks; blk += gridDim.x) {
      __syncthreads();
      for (int idx = tid - 256; idx < DMMA_ROWS_PER_CTA * LDA; idx += 256) sm.P[idx] = 0.0;
      __syncthreads();
      const int tiles = tiles_of(blk);
      int tg = tiles - 4 * group;
      tg = tg < 0 ? 0 : (tg > 4 ? 4 : tg);
      const int mta = tg < 2 ? tg : 2, mtb = tg - mta;
      const int mt = set == 0 ? mta : mtb;
      if (tile_in_set < mt)
        dmma3_update_role<Target, DP>(sm, target, dim, step_size, n_steps, h_out != nullptr,
                                      group, set, tile_in_set, lane);
    }
  }
}

template <class Target, int DP>
static int launch_dmma3(const double* q_in, const double* p_in, double* q_out, double* p_out,
                        const int32_t* dir, int64_t n, int dim, double eps, int n_steps,
                        const double* minv, const ModelArgs& m, double* h_out, int32_t* status,
                        int32_t* n_done, cudaStream_t st, int sms) {
  auto kern = leapfrog_dmma3_kernel<Target, DP>;
  const size_t smem = sizeof(Dmma3Smem<DP>);
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) !=
      cudaSuccess)
    return MB200_ERR_CUDA;
  int64_t blocks = (n + DMMA_ROWS_PER_CTA - 1) / DMMA_ROWS_PER_CTA;
  if (blocks > sms) blocks = sms;
  kern<<<(unsigned)blocks, 512, smem, st>>>(q_in, p_in, q_out, p_out, dir, n, dim, eps, n_steps,
                                            minv, m, h_out, status, n_done);
  return 0;
}

template <class Target>
static int dispatch_dmma3_dim(const double* q_in, const double* p_in, double* q_out,
                              double* p_out, const int32_t* dir, int64_t n, int dim, double eps,
                              int n_steps, const double* minv, const ModelArgs& m, double* h_out,
                              int32_t* status, int32_t* n_done, cudaStream_t st, int sms) {
#define MB200_DM(DP)                                                                            \
  return launch_dmma3<Target, DP>(q_in, p_in, q_out, p_out, dir, n, dim, eps, n_steps, minv, m, \
                                  h_out, status, n_done, st, sms)
  if (dim <= 32) MB200_DM(32);
  if (dim <= 64) MB200_DM(64);
  if (dim <= 96) MB200_DM(96);
  MB200_DM(128);
#undef MB200_DM
}

static int leapfrog_dmma3_dispatch(const double* q_in, const double* p_in, double* q_out,
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
    case MB200_TARGET_STD_GAUSSIAN: return dispatch_dmma3_dim<StdGaussianTarget>(MB200_ARGS);
    case MB200_TARGET_NEAL_FUNNEL: return dispatch_dmma3_dim<NealFunnelTarget>(MB200_ARGS);
    case MB200_TARGET_BANANA: return dispatch_dmma3_dim<BananaTarget>(MB200_ARGS);
    default: return MB200_ERR_UNSUPPORTED;
  }
#undef MB200_ARGS
}

}  // namespace mb200
