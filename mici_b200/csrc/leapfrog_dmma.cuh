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
#include <type_traits>

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
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// Profiling hook (profiles/tools/k1_bench.cu defines it to record clock64 per warp and phase);
// expands to nothing in the library build.
#ifndef MB200_K1_TRACE
#define MB200_K1_TRACE(phase)
#endif
#ifndef MB200_K1_MARK
#define MB200_K1_MARK(id)
#endif

constexpr int DMMA_TILES_PER_CTA = 7;
#ifndef MB200_DMMA_GROUPS
#define MB200_DMMA_GROUPS 4
#endif
// groups of 4 warps (one per SM sub-partition); the 7 row tiles are dealt as evenly as possible:
// 2,2,2,1 for 4 groups, 4,3 for 2 groups
constexpr int DMMA_GROUPS = MB200_DMMA_GROUPS;
__host__ __device__ constexpr int dmma_tile_count(int g) {
  return DMMA_TILES_PER_CTA / DMMA_GROUPS + (g < DMMA_TILES_PER_CTA % DMMA_GROUPS ? 1 : 0);
}
__host__ __device__ constexpr int dmma_tile_start(int g) {
  return g * (DMMA_TILES_PER_CTA / DMMA_GROUPS) +
         (g < DMMA_TILES_PER_CTA % DMMA_GROUPS ? g : DMMA_TILES_PER_CTA % DMMA_GROUPS);
}
constexpr int DMMA_MAX_MT = (DMMA_TILES_PER_CTA + DMMA_GROUPS - 1) / DMMA_GROUPS;
constexpr int DMMA_THREADS = 32 * 4 * DMMA_GROUPS;
constexpr int DMMA_ROWS_PER_CTA = 8 * DMMA_TILES_PER_CTA;  // 56 chains

template <int DP>
struct DmmaSmem {
  // row stride (doubles): rows shift by 64 B mod 128 B, so the 128-bit fragment loads of a
  // quarter-warp (2 rows x 4 lanes x 16 B) touch every bank once
  static constexpr int LDA = DP + 8;
  double A[DP * LDA];
  double P[DMMA_ROWS_PER_CTA * LDA];
  // per-chain exchange between the four warps that share a row: partial sums of the target's
  // reduction ([row][column quarter], one 32-byte line per row) and the scalar evaluated by the
  // owner of coordinate 0; `ex` is reused for the energy partials after the last step
  double psum[DMMA_ROWS_PER_CTA][4];
  double rscal[DMMA_ROWS_PER_CTA];
  double ex[2][DMMA_ROWS_PER_CTA][4];
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
// never leave the accumulator registers.
//
// Scheduling facts this is written around (profiles/r02_notes.md, fp64_arb / fp64_mix / k1_trace):
// a warp whose next instruction is a scalar FP64 operation makes NO progress while two or more
// other warps of its SM sub-partition stream DMMAs, so the per-step update phase of every group
// ends up running after the drifts of the whole sub-partition (the groups lock-step) and the
// step time is  drift (DMMA-pipe bound) + update phase (issue / latency bound).  The update phase
// is therefore kept as short as possible: no bounds logic (phantom coordinates are zero and stay
// zero under every registry target's kick), one FMA chain for the reduction, the per-chain scalar
// (funnel: exp(-v)) published by its owner instead of being summed, own momenta pre-loaded
// before the group barrier, two barriers per step.
//
// PC = true: per-chain step sizes (adaptive warm-up, adapters.py:40-235 per chain).  sm.A then holds
// A unscaled (step_size = 1) and the momentum tile holds  s = eps_c * dir * p:
//   kick:  s -= (eps_c^2 / 2) * grad l(q)      drift: q += s . A
// -- the same two flows, with eps_c applied on the momentum side instead of the matrix side.
template <class Target, int DP, int MT, bool PC>
__device__ __forceinline__ void leapfrog_dmma_group(
    DmmaSmem<DP>& sm, const Target& target, const double* q_in, const double* p_in,
    double* q_out, double* p_out, const int32_t* __restrict__ dir,
    const double* __restrict__ step_sizes, int64_t n_chains, int dim,
    double step_size, int n_steps, double* __restrict__ h_out, int32_t* __restrict__ status,
    int32_t* __restrict__ n_done, int64_t chain0, int row0, int w, int lane, int bar_id,
    int cta_threads, bool vec2, int32_t* __restrict__ counters) {
  constexpr int LDA = DmmaSmem<DP>::LDA;
  constexpr int NT = DP / 32;  // 8-column tiles per warp
  constexpr int KS = DP / 4;   // k steps
  const int r = lane >> 2, c = lane & 3;
  const int col0 = w * (DP / 4);  // first column of this warp's slice
  const double mh = -0.5 * step_size;
  // the lane that holds coordinate 0 of its rows (in q[mt][0][0])
  const bool owner = (w == 0) && (c == 0);

  // registers: positions of the slice in C-fragment layout (row 8mt + r, columns
  // col0 + 8nt + 2c + {0,1}); signed momenta live in sm.P with the same ownership
  double q[MT][NT][2], sgn[MT], mhr[PC ? MT : 1];
  bool live[MT];
  double2* pslot[MT];  // &sm.P[row][col0 + 2c]; + 4*nt double2 per column tile
  int row[MT];

#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
    row[mt] = row0 + 8 * mt + r;
    const int64_t ch = chain0 + row[mt];
    live[mt] = ch < n_chains;
    sgn[mt] = (live[mt] && dir != nullptr && dir[ch] < 0) ? -1.0 : 1.0;
    if (PC) {  // sgn becomes the load scale dir * eps_c
      const double e = live[mt] ? step_sizes[ch] : 0.0;
      sgn[mt] *= e;
      mhr[mt] = -0.5 * (e * e);
    }
    pslot[mt] = reinterpret_cast<double2*>(&sm.P[row[mt] * LDA + col0 + 2 * c]);
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const int i = col0 + 8 * nt + 2 * c;
      double2 a = make_double2(0.0, 0.0), b = make_double2(0.0, 0.0);
      if (live[mt] && i < dim) {
        if (vec2) {  // even dim, 16-byte aligned state arrays
          a = *reinterpret_cast<const double2*>(q_in + (size_t)ch * dim + i);
          b = *reinterpret_cast<const double2*>(p_in + (size_t)ch * dim + i);
        } else {
          a.x = q_in[(size_t)ch * dim + i], b.x = p_in[(size_t)ch * dim + i];
          if (i + 1 < dim) a.y = q_in[(size_t)ch * dim + i + 1], b.y = p_in[(size_t)ch * dim + i + 1];
        }
      }
      q[mt][nt][0] = a.x, q[mt][nt][1] = a.y;
      pslot[mt][4 * nt] = make_double2(sgn[mt] * b.x, sgn[mt] * b.y);
    }
  }

  int s = -1;  // current step (read by the profiling hook only)

  // ---- update phase, part 1 (before the group barrier): this warp's share of the per-chain
  // reduction and, on the owner lanes, the per-chain scalar
  auto publish_partials = [&]() {
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      if (Target::TILE_SUM) {
        // sum of squares over the slice: two FMA chains per row
        double t0 = q[mt][0][0] * q[mt][0][0];
        if (Target::COORD0) t0 = owner ? 0.0 : t0;
        t0 = fma(q[mt][0][1], q[mt][0][1], t0);
        double t1 = 0.0;
#pragma unroll
        for (int nt = 1; nt < NT; ++nt) {
          double& t = (nt & 1) ? t1 : t0;
          t = fma(q[mt][nt][0], q[mt][nt][0], t);
          t = fma(q[mt][nt][1], q[mt][nt][1], t);
        }
        double v = t0 + t1;
        v += __shfl_xor_sync(FULL_MASK, v, 1);
        v += __shfl_xor_sync(FULL_MASK, v, 2);
        if (c == 0) sm.psum[row[mt]][w] = v;
      }
      if (Target::ROW_SCALAR && owner) sm.rscal[row[mt]] = target.row_scalar(q[mt][0][0]);
    }
  };

  // ---- update phase, part 2: s -= (eps/2) * grad l(q), KICKS times (1 or 2: the two half-steps
  // either side of a step boundary stay two separately rounded updates, systems.py:152), one FMA
  // per coordinate and kick; then make the new momenta visible to the group.  The first barrier
  // also orders "all A-fragment reads of sm.P done" before the in-place update.
  auto kick_and_publish = [&](auto kicks_tag) {
    constexpr int KICKS = decltype(kicks_tag)::value;
    double2 pv[MT][NT];
    if (KICKS > 0) {
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) pv[mt][nt] = pslot[mt][4 * nt];  // own slots: no hazard
    }
    MB200_K1_TRACE(2);
    named_barrier_sync(bar_id, 128);
    MB200_K1_TRACE(3);
    if (KICKS == 0) return;
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      const double rs = Target::ROW_SCALAR ? sm.rscal[row[mt]] : 1.0;
      const double mhm = PC ? mhr[PC ? mt : 0] : mh;
      const double coef = target.kick_coef(mhm, rs);
      double c00 = coef, a00 = q[mt][0][0];
      if (Target::COORD0 && w == 0) {  // warp-uniform; only the owner lanes differ
        const double4 ps = *reinterpret_cast<const double4*>(&sm.psum[row[mt]][0]);
        const double g0 = target.grad0(q[mt][0][0], ((ps.x + ps.y) + ps.z) + ps.w, rs);
        c00 = owner ? mhm : coef;
        a00 = owner ? g0 : a00;
      }
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        double2 v = pv[mt][nt];
        if (Target::LINEAR) {
#pragma unroll
          for (int k = 0; k < KICKS; ++k) {
            v.x = (nt == 0) ? fma(c00, a00, v.x) : fma(coef, q[mt][nt][0], v.x);
            v.y = fma(coef, q[mt][nt][1], v.y);
          }
        } else {
#pragma unroll
          for (int k = 0; k < KICKS; ++k)
            target.kick_pair_nl(mhm, q[mt][nt][0], q[mt][nt][1], v.x, v.y);
        }
        pslot[mt][4 * nt] = v;
      }
    }
    MB200_K1_TRACE(4);
    named_barrier_sync(bar_id, 128);
    MB200_K1_TRACE(5);
  };
  using K0 = std::integral_constant<int, 0>;
  using K1 = std::integral_constant<int, 1>;
  using K2 = std::integral_constant<int, 2>;

  // acc += S * (eps A) on the tensor pipe (acc = q for the drift, acc = 0 for the energy).
  // Fragments are fetched with 128-bit loads: lane c of a row holds k = 8J + 2c and 8J + 2c + 1,
  // the two DMMAs of a k-pair contract {8J, 8J+2, 8J+4, 8J+6} and {8J+1, ..., 8J+7} (the pairing of
  // lanes with k is free as long as the A and B fragments agree).
  auto drift = [&](double (&acc)[MT][NT][2]) {
    const double2* a_base = reinterpret_cast<const double2*>(&sm.P[(row0 + r) * LDA + 2 * c]);
    const double2* b_base = reinterpret_cast<const double2*>(&sm.A[(col0 + r) * LDA + 2 * c]);
    // a single row tile gives a warp only NT independent accumulator chains (DMMA latency ~ 10
    // issue slots): the two halves of every k-pair then go to separate accumulator sets that are
    // added at the end (same products, summed in a different order)
    constexpr bool KSPLIT = (MT == 1);
    double acc2[KSPLIT ? NT : 1][2];
    if (KSPLIT) {
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) acc2[nt][0] = 0.0, acc2[nt][1] = 0.0;
    }
#pragma unroll 4
    for (int j = 0; j < KS / 2; ++j) {
      double2 a[MT], b[NT];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) a[mt] = a_base[mt * 4 * LDA + 4 * j];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) b[nt] = b_base[nt * 4 * LDA + 4 * j];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) dmma_m8n8k4(acc[mt][nt][0], acc[mt][nt][1], a[mt].x, b[nt].x);
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          if (KSPLIT) dmma_m8n8k4(acc2[nt][0], acc2[nt][1], a[mt].y, b[nt].y);
          else dmma_m8n8k4(acc[mt][nt][0], acc[mt][nt][1], a[mt].y, b[nt].y);
        }
    }
    if (KSPLIT) {
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) acc[0][nt][0] += acc2[nt][0], acc[0][nt][1] += acc2[nt][1];
    }
  };

  MB200_K1_MARK(4);
  publish_partials();
  if (n_steps > 0) kick_and_publish(K1{});
  else kick_and_publish(K0{});
  MB200_K1_MARK(5);
  for (s = 0; s < n_steps - 1; ++s) {
    MB200_K1_TRACE(0);
    drift(q);  // h2_flow (systems.py:363): q += dir*eps * (A p)
    MB200_K1_TRACE(1);
    publish_partials();
    kick_and_publish(K2{});  // closes step s and (cached gradient) opens step s+1
  }
  if (n_steps > 0) {
    drift(q);
    publish_partials();
    kick_and_publish(K1{});
  }

  MB200_K1_MARK(6);
  // ---- store (p = dir * s)
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
    const int64_t ch = chain0 + row[mt];
    if (!live[mt]) continue;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const int i = col0 + 8 * nt + 2 * c;
      if (i < dim) {
        const double2 sv = pslot[mt][4 * nt];
        double2 pv = make_double2(sgn[mt] * sv.x, sgn[mt] * sv.y);  // dir = +-1: exact
        if (PC) {
          // s = (dir eps_c) p  ->  p = s / (dir eps_c); a chain with eps_c = 0 has not moved
          if (sgn[mt] != 0.0) {
            pv = make_double2(sv.x / sgn[mt], sv.y / sgn[mt]);
          } else {
            pv.x = p_in[(size_t)ch * dim + i];
            pv.y = (i + 1 < dim) ? p_in[(size_t)ch * dim + i + 1] : 0.0;
          }
        }
        if (vec2) {
          *reinterpret_cast<double2*>(q_out + (size_t)ch * dim + i) =
              make_double2(q[mt][nt][0], q[mt][nt][1]);
          *reinterpret_cast<double2*>(p_out + (size_t)ch * dim + i) = pv;
        } else {
          q_out[(size_t)ch * dim + i] = q[mt][nt][0], p_out[(size_t)ch * dim + i] = pv.x;
          if (i + 1 < dim)
            q_out[(size_t)ch * dim + i + 1] = q[mt][nt][1], p_out[(size_t)ch * dim + i + 1] = pv.y;
        }
      }
    }
    if (w == 0 && c == 0) {
      if (status != nullptr) status[ch] = MB200_STATUS_OK;
      if (n_done != nullptr) n_done[ch] = n_steps;
      if (counters != nullptr) counters[ch * MB200_N_COUNTERS + MB200_COUNT_GRAD] += n_steps + 1;
    }
  }

  MB200_K1_MARK(7);
  // ---- Hamiltonian of the final state: l(q) + p . (A p) / 2   (systems.py:187-196, 348-350)
  // sm.psum / sm.rscal hold the reduction of the final positions (last update phase)
  if (h_out != nullptr) {
    double l[MT], kin[MT], u[MT][NT][2];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      double red[2] = {0.0, 1.0};
      if (Target::TILE_SUM) {
        const double4 ps = *reinterpret_cast<const double4*>(&sm.psum[row[mt]][0]);
        red[0] = ((ps.x + ps.y) + ps.z) + ps.w;
      }
      if (Target::ROW_SCALAR) red[1] = sm.rscal[row[mt]];
      l[mt] = 0.0;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const int i = col0 + 8 * nt + 2 * c;
        if (i < dim) l[mt] += target.nld_pair(i, q[mt][nt][0], q[mt][nt][1], red);
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
      if (c == 0) {
        sm.ex[0][row[mt]][w] = kin[mt];
        sm.ex[1][row[mt]][w] = l[mt];
      }
    }
    named_barrier_sync(bar_id, 128);
    if (w == 0 && c == 0) {
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        if (!live[mt]) continue;
        const double4 k4 = *reinterpret_cast<const double4*>(&sm.ex[0][row[mt]][0]);
        const double4 l4 = *reinterpret_cast<const double4*>(&sm.ex[1][row[mt]][0]);
        const double ks = ((k4.x + k4.y) + k4.z) + k4.w;
        const double ls = ((l4.x + l4.y) + l4.z) + l4.w;
        if (PC) {  // ks = eps_c^2 p.A p  (p.A p itself is unavailable for eps_c = 0: NaN)
          const double e = step_sizes[chain0 + row[mt]];
          h_out[chain0 + row[mt]] = ls + 0.5 * (ks / (e * e));
        } else {
          h_out[chain0 + row[mt]] = ls + 0.5 * (ks / step_size);
        }
      }
    }
  }
}

template <class Target, int DP, bool PC>
__global__ void __launch_bounds__(DMMA_THREADS, 1)
    leapfrog_dmma_kernel(const double* q_in, const double* p_in, double* q_out, double* p_out,
                         const int32_t* __restrict__ dir, const double* __restrict__ step_sizes,
                         int64_t n_chains, int dim,
                         double step_size, int n_steps, const double* __restrict__ minv,
                         ModelArgs model, double* __restrict__ h_out,
                         int32_t* __restrict__ status, int32_t* __restrict__ n_done,
                         int tiles_per_cta, int vec2, int tma_rows) {
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
  MB200_K1_MARK(0);

  // ---- stage A = M^-1 into shared memory: one TMA bulk copy per row, one mbarrier.  Issued
  // first; everything below until the wait overlaps the copies.
  const uint32_t mbar = smem_u32(&sm.mbar);
  if (tma_rows && warp == 0) {
    if (lane == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    const uint32_t row_bytes = (uint32_t)dim * 8u;
    if (lane == 0)
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar),
                   "r"(row_bytes * (uint32_t)dim)
                   : "memory");
    __syncwarp();
#pragma unroll 1
    for (int row = lane; row < dim; row += 32) {  // the 32 lanes issue the row copies
      const unsigned long long src =
          reinterpret_cast<unsigned long long>(minv) + (unsigned long long)row * row_bytes;
      asm volatile(
          "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
          ::"r"(smem_u32(&sm.A[row * LDA])),
          "l"(src), "r"(row_bytes), "r"(mbar)
          : "memory");
    }
  }
  // pull this CTA's first block of state rows towards L2 while A is in flight
  {
    const int64_t chain0 = (int64_t)blockIdx.x * (8 * tiles_per_cta);
    const int64_t left = n_chains - chain0;
    const int64_t rows = left < 8 * tiles_per_cta ? left : 8 * tiles_per_cta;
    const int64_t lines = (rows * dim * 8 + 127) / 128;
    for (int64_t i = tid; i < 2 * lines; i += blockDim.x) {
      const double* base = (i < lines ? q_in : p_in) + (size_t)chain0 * dim;
      const char* ptr = reinterpret_cast<const char*>(base) + (i < lines ? i : i - lines) * 128;
      asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr));
    }
  }
  if (!tma_rows) {  // odd dim / unaligned matrix: rows are not 16-byte multiples, plain copies
    for (int idx = tid; idx < dim * dim; idx += blockDim.x) {
      const int row = idx / dim, col = idx - row * dim;
      sm.A[row * LDA + col] = minv[idx];
    }
  }
  // zero the part of the padding that the fragment loads read (rows / columns in [dim, DP)) --
  // disjoint from the TMA destinations; sm.P needs none (every slot that is read is written by
  // the state load, phantom coordinates as zeros)
  if (dim < DP) {
    for (int idx = tid; idx < DP * DP; idx += blockDim.x) {
      const int row = idx / DP, col = idx - row * DP;
      if (row >= dim || col >= dim) sm.A[row * LDA + col] = 0.0;
    }
  }
  MB200_K1_MARK(1);
  // every thread polls the mbarrier below: its initialisation by warp 0 (and the zero-fill
  // above) must be visible first
  __syncthreads();
  // wait for the bytes to land (phase 0), then scale the staged metric: sm.A = eps * A
  if (tma_rows) {
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
  MB200_K1_MARK(2);
  if (!PC)  // per-chain step sizes: the matrix stays unscaled
#pragma unroll 4
  for (int idx = tid; idx < DP * (DP / 2); idx += blockDim.x) {
    const int row = idx / (DP / 2), c2 = idx - row * (DP / 2);
    double2* ptr = reinterpret_cast<double2*>(&sm.A[row * LDA]) + c2;
    double2 v = *ptr;
    v.x *= step_size, v.y *= step_size;
    *ptr = v;
  }
  __syncthreads();
  MB200_K1_MARK(3);

  // A CTA owns `tiles_per_cta` (<= 7) row tiles of 8 chains per pass -- 7 when the batch fills
  // the GPU (56 chains per SM), fewer for small batches so that the tiles spread over all SMs
  // (strong scaling: 1024 chains -> 128 CTAs of one tile instead of 19 CTAs of seven).  The
  // tiles of a pass are dealt to the 4 groups as evenly as possible (7 -> 2,2,2,1; 4 -> 1,1,1,1).
  const int rows_per_cta = 8 * tiles_per_cta;
  for (int64_t blk = blockIdx.x; blk * rows_per_cta < n_chains; blk += gridDim.x) {
    const int64_t chain0 = blk * rows_per_cta;
    const int64_t left = n_chains - chain0;
    const int tiles = (int)((left >= rows_per_cta) ? tiles_per_cta : (left + 7) / 8);
    const int base = tiles / DMMA_GROUPS, rem = tiles % DMMA_GROUPS;
    const int mt = base + (group < rem ? 1 : 0);
    const int row0 = 8 * (group * base + (group < rem ? group : rem));
    const int cta_threads = 128 * (tiles < DMMA_GROUPS ? tiles : DMMA_GROUPS);
#define MB200_GROUP(MT)                                                                       \
  leapfrog_dmma_group<Target, DP, MT, PC>(sm, target, q_in, p_in, q_out, p_out, dir,          \
                                          step_sizes, n_chains, dim, step_size, n_steps,      \
                                          h_out, status, n_done, chain0, row0, w, lane,       \
                                          1 + group, cta_threads, vec2 != 0, model.counters)
    if (DMMA_MAX_MT >= 4 && mt == 4) MB200_GROUP(4);
    else if (DMMA_MAX_MT >= 3 && mt == 3) MB200_GROUP(3);
    else if (mt == 2) MB200_GROUP(2);
    else if (mt == 1) MB200_GROUP(1);
#undef MB200_GROUP
    __syncthreads();  // next block of chains reuses sm.P / sm.part
  }
}

template <class Target, int DP, bool PC>
static int launch_dmma(const double* q_in, const double* p_in, double* q_out, double* p_out,
                       const int32_t* dir, const double* step_sizes, int64_t n, int dim,
                       double eps, int n_steps,
                       const double* minv, const ModelArgs& m, double* h_out, int32_t* status,
                       int32_t* n_done, cudaStream_t st, int sms) {
  auto kern = leapfrog_dmma_kernel<Target, DP, PC>;
  const size_t smem = sizeof(DmmaSmem<DP>);
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) !=
      cudaSuccess)
    return MB200_ERR_CUDA;
  const int64_t total_tiles = (n + 7) / 8;
  int64_t tpc = (total_tiles + sms - 1) / sms;  // tiles per CTA and pass
  if (tpc > DMMA_TILES_PER_CTA) tpc = DMMA_TILES_PER_CTA;
  int64_t blocks = (total_tiles + tpc - 1) / tpc;
  if (blocks > sms) blocks = sms;
  auto al16 = [](const void* ptr) { return (reinterpret_cast<uintptr_t>(ptr) & 15) == 0; };
  const int even = (dim & 1) == 0;
  const int vec2 = even && al16(q_in) && al16(p_in) && al16(q_out) && al16(p_out);
  const int tma_rows = even && al16(minv);  // persistent: CTAs loop over passes of tpc tiles
  kern<<<(unsigned)blocks, DMMA_THREADS, smem, st>>>(q_in, p_in, q_out, p_out, dir, step_sizes, n,
                                                     dim, PC ? 1.0 : eps, n_steps, minv, m, h_out,
                                                     status, n_done, (int)tpc, vec2, tma_rows);
  return 0;
}

template <class Target>
static int dispatch_dmma_dim(const double* q_in, const double* p_in, double* q_out, double* p_out,
                             const int32_t* dir, const double* step_sizes, int64_t n, int dim,
                             double eps, int n_steps,
                             const double* minv, const ModelArgs& m, double* h_out,
                             int32_t* status, int32_t* n_done, cudaStream_t st, int sms) {
#define MB200_DM(DP)                                                                        \
  return step_sizes != nullptr                                                              \
             ? launch_dmma<Target, DP, true>(q_in, p_in, q_out, p_out, dir, step_sizes, n,  \
                                             dim, eps, n_steps, minv, m, h_out, status,     \
                                             n_done, st, sms)                               \
             : launch_dmma<Target, DP, false>(q_in, p_in, q_out, p_out, dir, nullptr, n,    \
                                              dim, eps, n_steps, minv, m, h_out, status,    \
                                              n_done, st, sms)
  if (dim <= 32) MB200_DM(32);
  if (dim <= 64) MB200_DM(64);
  if (dim <= 96) MB200_DM(96);
  MB200_DM(128);
#undef MB200_DM
}

// Returns MB200_ERR_UNSUPPORTED when the shape is outside this kernel's domain (the caller
// then uses the general-dimension kernel).
int leapfrog_dmma_dispatch(const double* q_in, const double* p_in, double* q_out, double* p_out,
                           const int32_t* dir, const double* step_sizes, int64_t n, int dim,
                           double eps, int n_steps,
                           const double* minv, const ModelArgs& m, double* h_out, int32_t* status,
                           int32_t* n_done, cudaStream_t st) {
  if (dim > 128 || dim < 8) return MB200_ERR_UNSUPPORTED;
  if (step_sizes == nullptr && (!(eps != 0.0) || !isfinite(eps)))
    return MB200_ERR_UNSUPPORTED;  // eps*A formulation
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
#define MB200_ARGS                                                                             \
  q_in, p_in, q_out, p_out, dir, step_sizes, n, dim, eps, n_steps, minv, m, h_out, status, n_done, \
      st, sms
  switch (m.target_id) {
    case MB200_TARGET_STD_GAUSSIAN: return dispatch_dmma_dim<StdGaussianTarget>(MB200_ARGS);
    case MB200_TARGET_NEAL_FUNNEL: return dispatch_dmma_dim<NealFunnelTarget>(MB200_ARGS);
    case MB200_TARGET_BANANA: return dispatch_dmma_dim<BananaTarget>(MB200_ARGS);
    default: return MB200_ERR_UNSUPPORTED;
  }
#undef MB200_ARGS
}

}  // namespace mb200
