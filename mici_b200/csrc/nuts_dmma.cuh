// N4 on the tensor pipe: dynamic (NUTS) transitions for a SHARED DENSE Euclidean metric with
// dim <= 128, the chains of a CTA stepping leaf by leaf in lock-step.
//
// nuts.cuh gives every chain a warp and lets it run free; its mat-vec u = M^-1 grad l(q) then
// streams the whole staged metric from shared memory once per chain and leaf (1024 cycles of
// shared-memory bandwidth per leaf and SM: profiles/r02_notes.md).  Here the warps of a CTA still
// own one chain each and keep the tree bookkeeping of nuts.cuh word for word (same records, same
// order of random numbers, same merges: transitions.py:528-770), but meet once per leaf:
//   A  every warp writes the vector it needs multiplied (its chain's gradient) into a row of the
//      tile G [WARPS x DP] in shared memory,
//   B  all warps form U = G . M^-1 with DMMA.8x8x4 -- warp w owns DP/WARPS columns of U for all
//      rows, so the metric is read once per WARPS chains instead of once per chain,
//   C  every warp reads its row of U back and finishes its leaf (kicks, energy, subtree merges).
// One mat-vec per round for every live chain: INIT (v = M^-1 p), START of a doubling
// (u = M^-1 grad at the edge the tree is extended from), LEAF.  A chain whose tree is finished
// idles through the remaining rounds of its CTA (rows of G keep their last finite values).
//
// Subtree records (9 vectors: the two edges (q, p, v), the sum of momenta, the proposal (q, p))
// live in the per-chain global workspace as in nuts.cuh, but move less: a record buffer is handed
// from `cur` to a stack level by swapping slot numbers (no copy), an odd leaf waits for its sibling
// as 3 vectors, the even leaf is merged with it straight from registers, and every remaining copy
// loads all its vectors before it stores any (independent round trips to L2 instead of a chain).
//
// The products are the ones of nuts.cuh summed in a different order (tensor-pipe accumulation
// over k): trajectories agree with the oracle to rounding, which the fixtures cover at 1e-10.
#pragma once
#include "nuts.cuh"

namespace mb200 {

__device__ __forceinline__ void nuts_dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// A CTA holds GROUPS independent lock-step groups of 8 warps (8 chains: one DMMA row tile) that
// share the staged metric and synchronise among themselves only (named barriers): while one
// group is in its tensor-pipe phase the other does its bookkeeping.
template <int KP, int GROUPS>
struct NutsDmmaLayout {
  static constexpr int DP = 64 * KP;
  static constexpr int LDA = DP + 8;   // row stride: conflict-free 128-bit fragment loads
  static constexpr int WARPS = 8 * GROUPS;
  static constexpr int NTW = DP / 64;  // 8-column tiles of U per warp (8 warps span DP columns)
  static constexpr size_t smem_bytes() {
    return (size_t)(DP * LDA + 2 * WARPS * LDA) * sizeof(double);
  }
};

__device__ __forceinline__ void nuts_group_sync(int id, int nthreads) {
  asm volatile("barrier.cta.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ bool nuts_group_any(int id, int nthreads, bool pred) {
  uint32_t res;
  asm volatile(
      "{\n .reg .pred p, q;\n setp.ne.u32 q, %3, 0;\n"
      " barrier.cta.red.or.pred p, %1, %2, q;\n selp.u32 %0, 1, 0, p;\n}"
      : "=r"(res)
      : "r"(id), "r"(nthreads), "r"((uint32_t)pred)
      : "memory");
  return res != 0;
}

template <class Target, int KP, int GROUPS>
__global__ void __launch_bounds__(GROUPS * 256, 1)
    nuts_dmma_kernel(const double* __restrict__ q_in, const double* __restrict__ p_in,
                     double* __restrict__ q_out, double* __restrict__ p_out, int64_t n_chains,
                     int dim, double step_size, const double* __restrict__ minv, ModelArgs model,
                     NutsArgs a, double* __restrict__ workspace, double* __restrict__ h_out,
                     int32_t* __restrict__ n_step_out, double* __restrict__ av_accept_out,
                     double* __restrict__ reject_prob_out, int32_t* __restrict__ depth_out,
                     int32_t* __restrict__ diverging_out, int32_t* __restrict__ n_used_out,
                     int32_t* __restrict__ dir_out, int32_t* __restrict__ status) {
  using N = Nuts<Target, KP>;
  using K = LeapfrogGeneric<Target, KP, 1>;
  using L = NutsDmmaLayout<KP, GROUPS>;
  constexpr int NV = 2 * KP;
  constexpr int DP = L::DP, LDA = L::LDA, WARPS = L::WARPS, NTW = L::NTW, MT = 1;
  extern __shared__ __align__(16) double smem[];
  double* sA = smem;                 // M^-1, zero padded to DP x DP
  double* sG = sA + DP * LDA;        // operands  [WARPS][LDA]
  double* sU = sG + WARPS * LDA;     // products  [WARPS][LDA]
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int grp = warp >> 3, gw = warp & 7;  // lock-step group, warp (= chain row) within it
  const int bar_id = 1 + grp;
  for (int idx = threadIdx.x; idx < DP * DP; idx += blockDim.x) {
    const int row = idx / DP, col = idx - row * DP;
    sA[row * LDA + col] = (row < dim && col < dim) ? minv[(size_t)row * dim + col] : 0.0;
  }
  for (int idx = threadIdx.x; idx < WARPS * LDA; idx += blockDim.x) sG[idx] = 0.0;
  __syncthreads();

  const Target target(model, dim);
  const bool slice = a.slice != 0, euclid = a.euclidean_criterion != 0, extra = a.extra_checks != 0;
  const size_t ws_stride = (size_t)(7 + 2 + NUTS_REC * (1 + a.max_depth)) * DP;
  double2* g_row = reinterpret_cast<double2*>(sG + warp * LDA) + lane;        // + 32 k
  const double2* u_row = reinterpret_cast<const double2*>(sU + warp * LDA) + lane;
  // fragment bases of phase B (row r of every row tile; column col0 + r of every column tile)
  const int fr = lane >> 2, fc = lane & 3;
  const int col0 = gw * (8 * NTW);
  const double2* a_base = reinterpret_cast<const double2*>(sG + (8 * grp + fr) * LDA + 2 * fc);
  const double2* b_base = reinterpret_cast<const double2*>(sA + (col0 + fr) * LDA + 2 * fc);

  enum { PH_INIT = 0, PH_START = 1, PH_LEAF = 2 };

  for (int64_t base = ((int64_t)blockIdx.x * GROUPS + grp) * 8; base < n_chains;
       base += (int64_t)gridDim.x * WARPS) {
    const int64_t ch = base + gw;
    bool alive = ch < n_chains;
    const int64_t chs = alive ? ch : 0;  // dead warps of a ragged last block touch nothing
    double* tree = workspace + (size_t)chs * ws_stride;
    double* next = tree + 7 * DP;
    double* pool = next + 2 * DP;  // 1 + max_depth record buffers (cur + one per stack level)
    constexpr size_t REC = (size_t)NUTS_REC * DP;
    int cur_slot = 0;
    unsigned free_mask = 0u;
    unsigned char lvl_slot[NUTS_MAX_DEPTH];
    const double* uni = a.uniforms + (size_t)chs * a.n_uniforms;
    int n_used = 0;
    bool starved = false;
    auto uniform = [&]() -> double {
      if (n_used >= a.n_uniforms) {
        starved = true;
        return 0.5;
      }
      return uni[n_used++];
    };
    const double eps = (alive && a.step_sizes != nullptr) ? a.step_sizes[ch] : step_size;

    double q[1][NV], p[1][NV], v[1][NV], g[NV], u[NV];
    double red_q[Target::NRED + 1];  // the target's reduced sums at the current q (leaf phases)
#pragma unroll
    for (int r = 0; r < Target::NRED + 1; ++r) red_q[r] = 0.0;
#pragma unroll
    for (int e = 0; e < NV; ++e) {
      const int i = 2 * lane + 64 * (e >> 1) + (e & 1);
      q[0][e] = (alive && i < dim) ? q_in[(size_t)ch * dim + i] : 0.0;
      p[0][e] = (alive && i < dim) ? p_in[(size_t)ch * dim + i] : 0.0;
      v[0][e] = 0.0, g[e] = 0.0, u[e] = 0.0;
    }
    auto energy = [&]() -> double {  // System.h (systems.py:187-196) with v = M^-1 p current
      double kin = 0.0;
#pragma unroll
      for (int e = 0; e < NV; ++e) kin = fma(p[0][e], v[0][e], kin);
      kin = warp_sum(kin);
      return K::neg_log_dens(target, dim, lane, q[0]) + 0.5 * kin;
    };
    double h_init = 0.0, log_u = 0.0, w_tree = 0.0, h_next = 0.0;
    auto leaf_weight = [&](double h) -> double {
      return slice ? ((log_u <= -h) ? 1.0 : 0.0) : -h;
    };
    double lw[NUTS_MAX_DEPTH], lh[NUTS_MAX_DEPTH];  // weight / proposal energy per stack level
    double sum_accept = 0.0, reject_prob = 1.0, w_cur = 0.0, h_cur = 0.0, dt = 0.0;
    int n_step = 0, depth = 0, dirn = 1, k = 0, n_leaves = 1, phase = PH_INIT;
    bool diverging = false;
    // `dir` of the returned state object (see nuts.cuh)
    bool next_is_init = true, pos_is_init = true, neg_is_init = true;
    int init_dir = 1, next_dir = 1;

    while (true) {
      // ---------------------------------------------------------------- A: operand rows
      if (alive) {
        if (phase == PH_INIT) {
#pragma unroll
          for (int kk = 0; kk < KP; ++kk) g_row[32 * kk] = make_double2(p[0][2 * kk], p[0][2 * kk + 1]);
        } else {
          if (phase == PH_START) {
            dirn = (uniform() < 0.5) ? 1 : -1;  // transitions.py:729
            if (dirn == 1 ? pos_is_init : neg_is_init) init_dir = dirn;
            dt = dirn * eps;
            __syncwarp();
            N::ld(tree + (dirn == 1 ? PQ : NQ) * DP, lane, q[0]);
            N::ld(tree + (dirn == 1 ? PP : NP) * DP, lane, p[0]);
            N::ld(tree + (dirn == 1 ? PVEL : NVEL) * DP, lane, v[0]);
          } else {
            // LeapfrogIntegrator._step (integrators.py:170-173), first half kick and the drift;
            // v = M^-1 p follows the kick by linearity (u = M^-1 grad, see nuts.cuh)
#pragma unroll
            for (int e = 0; e < NV; ++e) {
              p[0][e] = __dsub_rn(p[0][e], __dmul_rn(0.5 * dt, g[e]));
              v[0][e] = __dsub_rn(v[0][e], __dmul_rn(0.5 * dt, u[e]));
            }
#pragma unroll
            for (int e = 0; e < NV; ++e) q[0][e] = __dadd_rn(q[0][e], __dmul_rn(dt, v[0][e]));
          }
          K::grad_keep(target, dim, lane, q[0], g, red_q);
#pragma unroll
          for (int kk = 0; kk < KP; ++kk) g_row[32 * kk] = make_double2(g[2 * kk], g[2 * kk + 1]);
        }
      }
      if (!nuts_group_any(bar_id, 256, alive)) break;

      // ---------------------------------------------------------------- B: U = G . M^-1
      {
        double acc[MT][NTW][2], acc2[MT][NTW][2];  // the halves of a k-pair: independent chains
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
          for (int nt = 0; nt < NTW; ++nt)
            acc[mt][nt][0] = acc[mt][nt][1] = acc2[mt][nt][0] = acc2[mt][nt][1] = 0.0;
#pragma unroll 4
        for (int j = 0; j < DP / 8; ++j) {
          double2 fa[MT], fb[NTW];
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) fa[mt] = a_base[mt * 4 * LDA + 4 * j];
#pragma unroll
          for (int nt = 0; nt < NTW; ++nt) fb[nt] = b_base[nt * 4 * LDA + 4 * j];
#pragma unroll
          for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int nt = 0; nt < NTW; ++nt) {
              nuts_dmma(acc[mt][nt][0], acc[mt][nt][1], fa[mt].x, fb[nt].x);
              nuts_dmma(acc2[mt][nt][0], acc2[mt][nt][1], fa[mt].y, fb[nt].y);
            }
        }
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
          for (int nt = 0; nt < NTW; ++nt)
            *reinterpret_cast<double2*>(sU + (8 * grp + fr) * LDA + col0 + 8 * nt + 2 * fc) =
                make_double2(acc[mt][nt][0] + acc2[mt][nt][0], acc[mt][nt][1] + acc2[mt][nt][1]);
      }
      nuts_group_sync(bar_id, 256);

      // ---------------------------------------------------------------- C: per-chain bookkeeping
      if (!alive) continue;
      double y[NV];
#pragma unroll
      for (int kk = 0; kk < KP; ++kk) {
        const double2 t = u_row[32 * kk];
        y[2 * kk] = t.x, y[2 * kk + 1] = t.y;
      }
      if (phase == PH_INIT) {
#pragma unroll
        for (int e = 0; e < NV; ++e) v[0][e] = y[e];
        h_init = energy();
        log_u = slice ? log(uniform()) - h_init : 0.0;  // transitions.py:832-839
        w_tree = leaf_weight(h_init);
        N::st(tree + NQ * DP, lane, q[0]), N::st(tree + PQ * DP, lane, q[0]);
        N::st(tree + NP * DP, lane, p[0]), N::st(tree + PP * DP, lane, p[0]);
        N::st(tree + SUMP * DP, lane, p[0]);
        N::st(tree + NVEL * DP, lane, v[0]), N::st(tree + PVEL * DP, lane, v[0]);
        N::st(next, lane, q[0]), N::st(next + DP, lane, p[0]);
        h_next = h_init;
        depth = 0;
        phase = PH_START;
        continue;
      }
#pragma unroll
      for (int e = 0; e < NV; ++e) u[e] = y[e];
      if (phase == PH_START) {
        w_cur = 0.0, h_cur = 0.0;
        n_leaves = 1 << depth;
        k = 1;
        cur_slot = 0;  // every level is empty at the start of a doubling
        free_mask = ((1u << (a.max_depth + 1)) - 1u) & ~1u;
        phase = PH_LEAF;
        continue;
      }
      // ---- finish leaf k of this doubling: second half kick
#pragma unroll
      for (int e = 0; e < NV; ++e) {
        p[0][e] = __dsub_rn(p[0][e], __dmul_rn(0.5 * dt, g[e]));
        v[0][e] = __dsub_rn(v[0][e], __dmul_rn(0.5 * dt, u[e]));
      }
      // System.h at the new state; l(q) reuses the sums the gradient reduced at this q
      double h;
      {
        double kin = 0.0;
#pragma unroll
        for (int e = 0; e < NV; ++e) kin = fma(p[0][e], v[0][e], kin);
        kin = warp_sum(kin);
        h = K::neg_log_dens_with(target, dim, lane, q[0], red_q) + 0.5 * kin;
      }
      if (h != h) h = INFINITY;  // transitions.py:626
      w_cur = leaf_weight(h);
      h_cur = h;
      const double h_diff = h_init - h;
      sum_accept += (h_diff != h_diff) ? 0.0 : exp(fmin(0.0, h_diff));
      ++n_step;
      bool terminate = false;
      double* cur = pool + (size_t)cur_slot * REC;
      if ((slice ? h + log_u : h - h_init) > a.max_delta_h) {  // _check_divergence
        diverging = true;
        terminate = true;
      } else if ((k & 1) && k < n_leaves) {
        // an odd-numbered leaf that is not the whole subtree waits on level 0 for its sibling:
        // 3 vectors (every slot of a leaf record aliases one of them)
        const int sl = __ffs(free_mask) - 1;
        free_mask &= ~(1u << sl);
        lvl_slot[0] = (unsigned char)sl;
        double* leaf = pool + (size_t)sl * REC;
        N::st(leaf + NQ * DP, lane, q[0]), N::st(leaf + NP * DP, lane, p[0]);
        N::st(leaf + NVEL * DP, lane, v[0]);
        lw[0] = w_cur, lh[0] = h_cur;
      } else if (k & 1) {
        // the single leaf of the first doubling is the whole subtree: a full record
        N::st(cur + NQ * DP, lane, q[0]), N::st(cur + PQ * DP, lane, q[0]);
        N::st(cur + RQ * DP, lane, q[0]);
        N::st(cur + NP * DP, lane, p[0]), N::st(cur + PP * DP, lane, p[0]);
        N::st(cur + RP * DP, lane, p[0]), N::st(cur + SUMP * DP, lane, p[0]);
        N::st(cur + NVEL * DP, lane, v[0]), N::st(cur + PVEL * DP, lane, v[0]);
      } else {
        // ---- level 0: merge the waiting leaf (inner) with this one (outer, in registers)
        {
          const int sl = lvl_slot[0];
          const double* inner = pool + (size_t)sl * REC;
          double lq[NV], lp[NV], lv[NV], sm[NV], wv[NV];
          N::ld(inner + NQ * DP, lane, lq);
          N::ld(inner + NP * DP, lane, lp);
          N::ld(inner + NVEL * DP, lane, lv);
          free_mask |= 1u << sl;
          const double w_new = nuts_add_w(slice, dirn == 1 ? lw[0] : w_cur,
                                          dirn == 1 ? w_cur : lw[0]);
          const bool take_outer = uniform() < nuts_ratio(slice, w_cur, w_new);
          // _termination_criterion of two leaves (no extra checks at merged depth 1):
          // sum of momenta neg + pos; w = pos.q - neg.q (euclidean) or the sum
#pragma unroll
          for (int e = 0; e < NV; ++e) {
            sm[e] = dirn == 1 ? lp[e] + p[0][e] : p[0][e] + lp[e];
            wv[e] = euclid ? (dirn == 1 ? q[0][e] - lq[e] : lq[e] - q[0][e]) : sm[e];
          }
          const double d1 = dirn == 1 ? N::dot(lv, wv) : N::dot(v[0], wv);
          const double d2 = dirn == 1 ? N::dot(v[0], wv) : N::dot(lv, wv);
          const bool stop = d1 < 0.0 || d2 < 0.0;
          N::st(cur + (dirn == 1 ? NQ : PQ) * DP, lane, lq);
          N::st(cur + (dirn == 1 ? NP : PP) * DP, lane, lp);
          N::st(cur + (dirn == 1 ? NVEL : PVEL) * DP, lane, lv);
          N::st(cur + (dirn == 1 ? PQ : NQ) * DP, lane, q[0]);
          N::st(cur + (dirn == 1 ? PP : NP) * DP, lane, p[0]);
          N::st(cur + (dirn == 1 ? PVEL : NVEL) * DP, lane, v[0]);
          N::st(cur + SUMP * DP, lane, sm);
          if (take_outer) {
            N::st(cur + RQ * DP, lane, q[0]), N::st(cur + RP * DP, lane, p[0]);
          } else {
            N::st(cur + RQ * DP, lane, lq), N::st(cur + RP * DP, lane, lp);
            h_cur = lh[0];
          }
          w_cur = w_new;
          if (stop) terminate = true;
        }
        int level = 1;
        for (int kk = k >> 1; !terminate && (kk & 1) == 0; kk >>= 1, ++level) {
          // merge the stored inner subtree of this level with the one just completed (outer)
          const int sl = lvl_slot[level];
          const double* inner = pool + (size_t)sl * REC;
          const double w_new = nuts_add_w(slice, dirn == 1 ? lw[level] : w_cur,
                                          dirn == 1 ? w_cur : lw[level]);
          const bool take_outer = uniform() < nuts_ratio(slice, w_cur, w_new);
          const double* neg = dirn == 1 ? inner : cur;
          const double* pos = dirn == 1 ? cur : inner;
          __syncwarp();
          const bool stop = N::turn(euclid, extra, neg, pos, level + 1, lane);
          // all loads first, then the stores
          double e0[NV], e1[NV], e2[NV], s1[NV], s2[NV], r0[NV], r1[NV];
          N::ld(inner + (dirn == 1 ? NQ : PQ) * DP, lane, e0);
          N::ld(inner + (dirn == 1 ? NP : PP) * DP, lane, e1);
          N::ld(inner + (dirn == 1 ? NVEL : PVEL) * DP, lane, e2);
          N::ld(cur + SUMP * DP, lane, s1);
          N::ld(inner + SUMP * DP, lane, s2);
          if (!take_outer) {
            N::ld(inner + RQ * DP, lane, r0);
            N::ld(inner + RP * DP, lane, r1);
          }
#pragma unroll
          for (int e = 0; e < NV; ++e)  // neg.sum_mom + pos.sum_mom
            s1[e] = dirn == 1 ? s2[e] + s1[e] : s1[e] + s2[e];
          N::st(cur + (dirn == 1 ? NQ : PQ) * DP, lane, e0);
          N::st(cur + (dirn == 1 ? NP : PP) * DP, lane, e1);
          N::st(cur + (dirn == 1 ? NVEL : PVEL) * DP, lane, e2);
          N::st(cur + SUMP * DP, lane, s1);
          if (!take_outer) {
            N::st(cur + RQ * DP, lane, r0);
            N::st(cur + RP * DP, lane, r1);
            h_cur = lh[level];
          }
          free_mask |= 1u << sl;
          w_cur = w_new;
          if (stop) terminate = true;
        }
        if (terminate) {
          // (level was advanced past the merge that stopped the tree: not used again)
        } else if (k < n_leaves) {
          // park the completed subtree on its level: hand over the buffer, take a free one
          lvl_slot[level] = (unsigned char)cur_slot;
          lw[level] = w_cur;
          lh[level] = h_cur;
          cur_slot = __ffs(free_mask) - 1;
          free_mask &= ~(1u << cur_slot);
        }
      }
      bool finished = terminate;
      if (!terminate) {
        if (k < n_leaves) {
          ++k;
        } else {
          // the doubling is complete: progressive sampling of the next state
          // (transitions.py:742-749), then merge it into the tree
          const double accept_prob = nuts_ratio(slice, w_cur, w_tree);
          const bool accept = uniform() < accept_prob;
          reject_prob *= 1.0 - accept_prob;
          if (dirn == 1) pos_is_init = false; else neg_is_init = false;
          const double* neg = dirn == 1 ? tree : cur;
          const double* pos = dirn == 1 ? cur : tree;
          __syncwarp();
          const bool stop = N::turn(euclid, extra, neg, pos, depth + 1, lane);
          // all loads first, then the stores
          double e0[NV], e1[NV], e2[NV], s1[NV], s2[NV], r0[NV], r1[NV];
          N::ld(cur + (dirn == 1 ? PQ : NQ) * DP, lane, e0);
          N::ld(cur + (dirn == 1 ? PP : NP) * DP, lane, e1);
          N::ld(cur + (dirn == 1 ? PVEL : NVEL) * DP, lane, e2);
          N::ld(tree + SUMP * DP, lane, s1);
          N::ld(cur + SUMP * DP, lane, s2);
          if (accept) {
            N::ld(cur + RQ * DP, lane, r0);
            N::ld(cur + RP * DP, lane, r1);
          }
#pragma unroll
          for (int e = 0; e < NV; ++e) s1[e] = dirn == 1 ? s1[e] + s2[e] : s2[e] + s1[e];
          N::st(tree + SUMP * DP, lane, s1);
          N::st(tree + (dirn == 1 ? PQ : NQ) * DP, lane, e0);
          N::st(tree + (dirn == 1 ? PP : NP) * DP, lane, e1);
          N::st(tree + (dirn == 1 ? PVEL : NVEL) * DP, lane, e2);
          if (accept) {
            N::st(next, lane, r0);
            N::st(next + DP, lane, r1);
            h_next = h_cur;
            next_is_init = false, next_dir = dirn;
          }
          w_tree = dirn == 1 ? nuts_add_w(slice, w_tree, w_cur) : nuts_add_w(slice, w_cur, w_tree);
          __syncwarp();
          if (stop) {
            finished = true;
          } else if (depth + 1 >= a.max_depth) {
            finished = true;  // `for depth in range(max_tree_depth)` ran out: depth stays
          } else {
            ++depth;
            phase = PH_START;
          }
        }
      }
      if (!finished) continue;
      // ---- the transition of this chain is complete: write it out and idle
      __syncwarp();
      double xo[NV], po[NV];
      N::ld(next, lane, xo);
      N::ld(next + DP, lane, po);
#pragma unroll
      for (int e = 0; e < NV; ++e) {
        const int i = 2 * lane + 64 * (e >> 1) + (e & 1);
        if (i < dim) {
          q_out[(size_t)ch * dim + i] = xo[e];
          p_out[(size_t)ch * dim + i] = po[e];
        }
      }
      if (lane == 0) {
        if (h_out != nullptr) h_out[ch] = h_next;
        if (n_step_out != nullptr) n_step_out[ch] = n_step;
        if (av_accept_out != nullptr) av_accept_out[ch] = n_step > 0 ? sum_accept / n_step : 0.0;
        if (reject_prob_out != nullptr) reject_prob_out[ch] = reject_prob;
        if (depth_out != nullptr) depth_out[ch] = depth;
        if (diverging_out != nullptr) diverging_out[ch] = diverging ? 1 : 0;
        if (n_used_out != nullptr) n_used_out[ch] = n_used;
        if (dir_out != nullptr) dir_out[ch] = next_is_init ? init_dir : next_dir;
        if (status != nullptr) status[ch] = starved ? MB200_STATUS_CONVERGENCE : MB200_STATUS_OK;
      }
      alive = false;
    }
  }
}

}  // namespace mb200
