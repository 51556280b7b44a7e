// N4: dynamic-length HMC transitions ("NUTS") on Euclidean-metric systems, one
// warp per chain -- every chain builds its own trajectory tree with its own control flow, so
// divergent tree depths cost nothing beyond the warp that owns the chain.
//
// Replaces, per chain (reference paths):
//   DynamicIntegrationTransition.sample / _build_tree / _termination_criterion / _merge_subtrees
//                                                   transitions.py:528-581, 610-770
//   MultinomialDynamicIntegrationTransition         transitions.py:773-809  (LogRepFloat weights,
//                                                   utils.py:50-71, 85-155)
//   SliceDynamicIntegrationTransition               transitions.py:812-858
//   euclidean_ / riemannian_no_u_turn_criterion     transitions.py:405-470
//   LeapfrogIntegrator.step + System.h per leaf     integrators.py:170-173, systems.py:187-196
//
// The recursion of `_build_tree` is unrolled into a binary-counter stack (after leaf number k of
// a doubling: one merge per trailing zero bit of k), exactly as oracle/mici_oracle.py's
// `nuts_transition`, which is bit-exact with the reference.  Random numbers: the caller supplies
// `n_uniforms` uniform variates per chain; the kernel consumes them in the reference's order
// ([slice: one at the start,] per doubling one for the direction, one per completed internal
// node in post-order, one for the progressive acceptance) and reports how many it used.
//
// Per-chain workspace in global memory (pair layout, DP = 64 KP doubles per vector):
//   tree  : NQ NP NV PQ PP PV SUM          (negative / positive edge: position, momentum,
//                                            velocity M^-1 p; sum of momenta)
//   next  : XQ XP                           (the state the transition returns)
//   cur   : NQ NP NV PQ PP PV SUM RQ RP     (subtree just completed; R = its proposal)
//   level[l] : same 9 vectors, l < max_tree_depth
#pragma once
#include "leapfrog_generic.cuh"

namespace mb200 {

constexpr int NUTS_MAX_DEPTH = 12;
constexpr int NUTS_REC = 9;  // vectors per subtree record
enum { NQ = 0, NP = 1, NVEL = 2, PQ = 3, PP = 4, PVEL = 5, SUMP = 6, RQ = 7, RP = 8 };

__host__ __device__ inline size_t nuts_workspace_doubles_per_chain(int dim, int max_depth) {
  const int kp = dim <= 64 ? 1 : dim <= 128 ? 2 : dim <= 256 ? 4 : dim <= 512 ? 8 : 16;
  return (size_t)(7 + 2 + NUTS_REC * (1 + max_depth)) * 64 * kp;
}

struct NutsArgs {
  int max_depth;
  double max_delta_h;
  int euclidean_criterion;  // 1: euclidean_no_u_turn_criterion, 0: riemannian_ (sum of momenta)
  int extra_checks;
  int slice;                // 1: SliceDynamicIntegrationTransition weights
  const double* uniforms;
  int n_uniforms;
  const double* step_sizes;
  int stage_metric;  // 1: copy the dense M^-1 into shared memory once per CTA
};

template <class Target, int KP>
struct Nuts {
  static constexpr int NV = 2 * KP;
  static constexpr int DP = 64 * KP;
  using K = LeapfrogGeneric<Target, KP, 1>;

  static __device__ __forceinline__ void ld(const double* vec, int lane, double (&a)[NV]) {
#pragma unroll
    for (int k = 0; k < KP; ++k) {
      const double2 t = *reinterpret_cast<const double2*>(vec + 2 * lane + 64 * k);
      a[2 * k] = t.x, a[2 * k + 1] = t.y;
    }
  }
  static __device__ __forceinline__ void st(double* vec, int lane, const double (&a)[NV]) {
#pragma unroll
    for (int k = 0; k < KP; ++k)
      *reinterpret_cast<double2*>(vec + 2 * lane + 64 * k) = make_double2(a[2 * k], a[2 * k + 1]);
  }
  static __device__ __forceinline__ void cp(double* dst, const double* src, int lane) {
#pragma unroll
    for (int k = 0; k < KP; ++k)
      *reinterpret_cast<double2*>(dst + 2 * lane + 64 * k) =
          *reinterpret_cast<const double2*>(src + 2 * lane + 64 * k);
  }
  static __device__ __forceinline__ double dot(const double (&a)[NV], const double (&b)[NV]) {
    double s = 0.0;
#pragma unroll
    for (int e = 0; e < NV; ++e) s = fma(a[e], b[e], s);
    return warp_sum(s);
  }

  // no-U-turn test between the states (q1, v1) and (q2, v2) (transitions.py:405-470):
  //   euclidean : w = q2 - q1 ;  riemannian : w = the supplied sum of momenta
  static __device__ bool no_u_turn(bool euclid, const double* q1, const double* v1,
                                   const double* q2, const double* v2, const double (&wsum)[NV],
                                   int lane) {
    double a[NV], b[NV], w[NV];
    if (euclid) {
      ld(q1, lane, a);
      ld(q2, lane, b);
#pragma unroll
      for (int e = 0; e < NV; ++e) w[e] = b[e] - a[e];
    } else {
#pragma unroll
      for (int e = 0; e < NV; ++e) w[e] = wsum[e];
    }
    ld(v1, lane, a);
    ld(v2, lane, b);
    const double d1 = dot(a, w), d2 = dot(b, w);
    return d1 < 0.0 || d2 < 0.0;
  }

  // _termination_criterion (transitions.py:528-556) for the tree that merges `neg` and `pos`
  // (records of >= 7 vectors) into one of depth `merged_depth`
  static __device__ bool turn(bool euclid, bool extra, const double* neg, const double* pos,
                              int merged_depth, int lane) {
    double s1[NV], s2[NV], t[NV];
    ld(neg + SUMP * DP, lane, s1);
    ld(pos + SUMP * DP, lane, s2);
#pragma unroll
    for (int e = 0; e < NV; ++e) t[e] = s1[e] + s2[e];
    if (no_u_turn(euclid, neg + NQ * DP, neg + NVEL * DP, pos + PQ * DP, pos + PVEL * DP, t, lane))
      return true;
    if (merged_depth > 1 && extra) {
      double m[NV];
      ld(pos + NP * DP, lane, m);
#pragma unroll
      for (int e = 0; e < NV; ++e) t[e] = s1[e] + m[e];  // neg.sum_mom + pos.negative.mom
      if (no_u_turn(euclid, neg + NQ * DP, neg + NVEL * DP, pos + NQ * DP, pos + NVEL * DP, t, lane))
        return true;
      ld(neg + PP * DP, lane, m);
#pragma unroll
      for (int e = 0; e < NV; ++e) t[e] = s2[e] + m[e];  // pos.sum_mom + neg.positive.mom
      if (no_u_turn(euclid, neg + PQ * DP, neg + PVEL * DP, pos + PQ * DP, pos + PVEL * DP, t, lane))
        return true;
    }
    return false;
  }
};

__device__ __forceinline__ double nuts_log1p_exp(double x) {  // utils.py:50-54
  return x > 0.0 ? x + log1p(exp(-x)) : log1p(exp(x));
}
__device__ __forceinline__ double nuts_log_sum_exp(double a, double b) {  // utils.py:65-71
  if (a == -INFINITY && b == -INFINITY) return -INFINITY;
  return a > b ? a + nuts_log1p_exp(b - a) : b + nuts_log1p_exp(a - b);
}
// _weight_function / LogRepFloat sum / _weight_ratio for the two variants
__device__ __forceinline__ double nuts_add_w(bool slice, double a, double b) {
  return slice ? a + b : nuts_log_sum_exp(a, b);
}
__device__ __forceinline__ double nuts_ratio(bool slice, double num, double den) {
  if (slice) return den > 0.0 ? fmin(num / den, 1.0) : fmin(num, 1.0);
  const double r = exp(num - den);  // NaN (both -inf) compares false below, like LogRepFloat
  return r > 1.0 ? 1.0 : r;
}

template <class Target, int KP>
__global__ void __launch_bounds__(KP == 2 ? 384 : 256)
    nuts_euclidean_kernel(const double* __restrict__ q_in, const double* __restrict__ p_in,
                          double* __restrict__ q_out, double* __restrict__ p_out, int64_t n_chains,
                          int dim, double step_size, int metric_kind,
                          const double* minv, ModelArgs model, NutsArgs a,
                          double* __restrict__ workspace, double* __restrict__ h_out,
                          int32_t* __restrict__ n_step_out, double* __restrict__ av_accept_out,
                          double* __restrict__ reject_prob_out, int32_t* __restrict__ depth_out,
                          int32_t* __restrict__ diverging_out, int32_t* __restrict__ n_used_out,
                          int32_t* __restrict__ dir_out, int32_t* __restrict__ status) {
  using N = Nuts<Target, KP>;
  using K = LeapfrogGeneric<Target, KP, 1>;
  constexpr int NV = 2 * KP;
  constexpr int DP = 64 * KP;
  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int wpb = blockDim.x >> 5;
  double* psm = smem + (size_t)warp * DP;
  if (a.stage_metric) {
    // every chain multiplies by the same dense M^-1 twice per leapfrog step: keep it in shared
    // memory for the life of the CTA instead of streaming it from L2 per chain
    double* s_minv = smem + (size_t)wpb * DP;
    for (int i = threadIdx.x; i < dim * dim; i += blockDim.x) s_minv[i] = minv[i];
    __syncthreads();
    minv = s_minv;
  }
#ifdef MB200_NUTS_LDS_MATVEC
  const bool fast_mv = a.stage_metric && (dim & 1) == 0;
#define MB200_NUTS_MATVEC(P, V)                                               \
  do {                                                                        \
    if (fast_mv) inv_metric_apply_staged<KP>(minv, dim, lane, psm, P, V);     \
    else inv_metric_apply<KP, 1>(metric_kind, minv, dim, lane, psm, P, V);    \
  } while (0)
#else
#define MB200_NUTS_MATVEC(P, V) inv_metric_apply<KP, 1>(metric_kind, minv, dim, lane, psm, P, V)
#endif
  const Target target(model, dim);
  const bool slice = a.slice != 0, euclid = a.euclidean_criterion != 0, extra = a.extra_checks != 0;
  const size_t ws_stride = (size_t)(7 + 2 + NUTS_REC * (1 + a.max_depth)) * DP;

  for (int64_t ch = (int64_t)blockIdx.x * wpb + warp; ch < n_chains;
       ch += (int64_t)gridDim.x * wpb) {
    double* tree = workspace + (size_t)ch * ws_stride;
    double* next = tree + 7 * DP;
    double* cur = next + 2 * DP;
    double* levels = cur + NUTS_REC * DP;
    const double* uni = a.uniforms + (size_t)ch * a.n_uniforms;
    int n_used = 0;
    bool starved = false;
    auto uniform = [&]() -> double {
      if (n_used >= a.n_uniforms) {
        starved = true;
        return 0.5;
      }
      return uni[n_used++];
    };
    const double eps = a.step_sizes != nullptr ? a.step_sizes[ch] : step_size;

    double q[1][NV], p[1][NV], v[1][NV], g[NV];
#pragma unroll
    for (int e = 0; e < NV; ++e) {
      const int i = 2 * lane + 64 * (e >> 1) + (e & 1);
      q[0][e] = (i < dim) ? q_in[(size_t)ch * dim + i] : 0.0;
      p[0][e] = (i < dim) ? p_in[(size_t)ch * dim + i] : 0.0;
    }
    auto energy = [&]() -> double {  // System.h (systems.py:187-196) with v = M^-1 p current
      double kin = 0.0;
#pragma unroll
      for (int e = 0; e < NV; ++e) kin = fma(p[0][e], v[0][e], kin);
      kin = warp_sum(kin);
      return K::neg_log_dens(target, dim, lane, q[0]) + 0.5 * kin;
    };
    MB200_NUTS_MATVEC(p, v);
    const double h_init = energy();
    const double log_u = slice ? log(uniform()) - h_init : 0.0;  // transitions.py:832-839
    auto leaf_weight = [&](double h) -> double {
      return slice ? ((log_u <= -h) ? 1.0 : 0.0) : -h;
    };
    double w_tree = leaf_weight(h_init);
    N::st(tree + NQ * DP, lane, q[0]), N::st(tree + PQ * DP, lane, q[0]);
    N::st(tree + NP * DP, lane, p[0]), N::st(tree + PP * DP, lane, p[0]);
    N::st(tree + SUMP * DP, lane, p[0]);
    N::st(tree + NVEL * DP, lane, v[0]), N::st(tree + PVEL * DP, lane, v[0]);
    N::st(next, lane, q[0]), N::st(next + DP, lane, p[0]);
    double h_next = h_init;

    double lw[NUTS_MAX_DEPTH], lh[NUTS_MAX_DEPTH];  // weight / proposal energy per stack level
    double sum_accept = 0.0, reject_prob = 1.0;
    int n_step = 0, depth = 0;
    bool diverging = false;
    // `dir` of the returned state object: leaf states carry the direction they were integrated
    // in; the initial state object is both edges of the tree at first and has its `dir`
    // overwritten (`state.dir = direction`, transitions.py:731) by every doubling that starts from
    // it.  The next adaptive stage's step-size search reads it (adapters.py:321).
    bool next_is_init = true, pos_is_init = true, neg_is_init = true;
    int init_dir = 1, next_dir = 1;
    for (depth = 0; depth < a.max_depth; ++depth) {
      const int dirn = (uniform() < 0.5) ? 1 : -1;  // transitions.py:729
      if (dirn == 1 ? pos_is_init : neg_is_init) init_dir = dirn;
      const double dt = dirn * eps;
      __syncwarp();
      N::ld(tree + (dirn == 1 ? PQ : NQ) * DP, lane, q[0]);
      N::ld(tree + (dirn == 1 ? PP : NP) * DP, lane, p[0]);
      N::ld(tree + (dirn == 1 ? PVEL : NVEL) * DP, lane, v[0]);
      K::grad(target, dim, lane, q[0], g);
      // ONE mat-vec per leaf: with u = M^-1 grad l(q) the velocity follows the momentum's half
      // kicks by linearity, v = M^-1 p:  v -= (dt/2) u.  (The reference applies M^-1 to p twice
      // per step -- for the drift and for dh_dmom of the new state; the two forms agree to
      // rounding.)
      double gu[1][NV], u[1][NV];
#pragma unroll
      for (int e = 0; e < NV; ++e) gu[0][e] = g[e];
      MB200_NUTS_MATVEC(gu, u);
      bool terminate = false;
      double w_cur = 0.0, h_cur = 0.0;
      const int n_leaves = 1 << depth;
      for (int k = 1; k <= n_leaves; ++k) {
        // LeapfrogIntegrator._step (integrators.py:170-173), two separately rounded half kicks
#pragma unroll
        for (int e = 0; e < NV; ++e) {
          p[0][e] = __dsub_rn(p[0][e], __dmul_rn(0.5 * dt, g[e]));
          v[0][e] = __dsub_rn(v[0][e], __dmul_rn(0.5 * dt, u[0][e]));
        }
#pragma unroll
        for (int e = 0; e < NV; ++e) q[0][e] = __dadd_rn(q[0][e], __dmul_rn(dt, v[0][e]));
        K::grad(target, dim, lane, q[0], g);
#pragma unroll
        for (int e = 0; e < NV; ++e) gu[0][e] = g[e];
        MB200_NUTS_MATVEC(gu, u);
#pragma unroll
        for (int e = 0; e < NV; ++e) {
          p[0][e] = __dsub_rn(p[0][e], __dmul_rn(0.5 * dt, g[e]));
          v[0][e] = __dsub_rn(v[0][e], __dmul_rn(0.5 * dt, u[0][e]));
        }
        double h = energy();
        if (h != h) h = INFINITY;  // transitions.py:626
        w_cur = leaf_weight(h);
        h_cur = h;
        // an odd-numbered leaf that is not the whole subtree is parked on level 0 untouched:
        // write it there directly instead of into `cur`
        const bool parked_leaf = (k & 1) && k < n_leaves;
        double* leaf = parked_leaf ? levels : cur;
        N::st(leaf + NQ * DP, lane, q[0]), N::st(leaf + PQ * DP, lane, q[0]);
        N::st(leaf + RQ * DP, lane, q[0]);
        N::st(leaf + NP * DP, lane, p[0]), N::st(leaf + PP * DP, lane, p[0]);
        N::st(leaf + RP * DP, lane, p[0]), N::st(leaf + SUMP * DP, lane, p[0]);
        N::st(leaf + NVEL * DP, lane, v[0]), N::st(leaf + PVEL * DP, lane, v[0]);
        const double h_diff = h_init - h;
        sum_accept += (h_diff != h_diff) ? 0.0 : exp(fmin(0.0, h_diff));
        ++n_step;
        if ((slice ? h + log_u : h - h_init) > a.max_delta_h) {  // _check_divergence
          diverging = true;
          terminate = true;
          break;
        }
        if (parked_leaf) {
          lw[0] = w_cur, lh[0] = h_cur;
          continue;
        }
        int level = 0;
        for (int kk = k; (kk & 1) == 0; kk >>= 1, ++level) {
          // merge the stored inner subtree of this level with the one just completed (outer)
          double* inner = levels + (size_t)level * NUTS_REC * DP;
          const double w_new = nuts_add_w(slice, dirn == 1 ? lw[level] : w_cur,
                                          dirn == 1 ? w_cur : lw[level]);
          const bool take_outer = uniform() < nuts_ratio(slice, w_cur, w_new);
          const double* neg = dirn == 1 ? inner : cur;
          const double* pos = dirn == 1 ? cur : inner;
          const bool stop = N::turn(euclid, extra, neg, pos, level + 1, lane);
          __syncwarp();
          if (dirn == 1) {
            N::cp(cur + NQ * DP, inner + NQ * DP, lane);
            N::cp(cur + NP * DP, inner + NP * DP, lane);
            N::cp(cur + NVEL * DP, inner + NVEL * DP, lane);
          } else {
            N::cp(cur + PQ * DP, inner + PQ * DP, lane);
            N::cp(cur + PP * DP, inner + PP * DP, lane);
            N::cp(cur + PVEL * DP, inner + PVEL * DP, lane);
          }
          double s1[NV], s2[NV];
          N::ld(cur + SUMP * DP, lane, s1);
          N::ld(inner + SUMP * DP, lane, s2);
#pragma unroll
          for (int e = 0; e < NV; ++e)  // neg.sum_mom + pos.sum_mom
            s1[e] = dirn == 1 ? s2[e] + s1[e] : s1[e] + s2[e];
          N::st(cur + SUMP * DP, lane, s1);
          if (!take_outer) {
            N::cp(cur + RQ * DP, inner + RQ * DP, lane);
            N::cp(cur + RP * DP, inner + RP * DP, lane);
            h_cur = lh[level];
          }
          __syncwarp();
          w_cur = w_new;
          if (stop) {
            terminate = true;
            break;
          }
        }
        if (terminate) break;
        if (k < n_leaves) {  // park the completed subtree on its level
          double* slot = levels + (size_t)level * NUTS_REC * DP;
#pragma unroll
          for (int r = 0; r < NUTS_REC; ++r) N::cp(slot + r * DP, cur + r * DP, lane);
          lw[level] = w_cur;
          lh[level] = h_cur;
          __syncwarp();
        }
      }
      if (terminate) break;
      // progressive sampling of the next state (transitions.py:742-749)
      const double accept_prob = nuts_ratio(slice, w_cur, w_tree);
      if (uniform() < accept_prob) {
        N::cp(next, cur + RQ * DP, lane);
        N::cp(next + DP, cur + RP * DP, lane);
        h_next = h_cur;
        next_is_init = false, next_dir = dirn;
      }
      reject_prob *= 1.0 - accept_prob;
      if (dirn == 1) pos_is_init = false; else neg_is_init = false;
      const double* neg = dirn == 1 ? tree : cur;
      const double* pos = dirn == 1 ? cur : tree;
      const bool stop = N::turn(euclid, extra, neg, pos, depth + 1, lane);
      __syncwarp();
      double s1[NV], s2[NV];
      N::ld(tree + SUMP * DP, lane, s1);
      N::ld(cur + SUMP * DP, lane, s2);
#pragma unroll
      for (int e = 0; e < NV; ++e) s1[e] = dirn == 1 ? s1[e] + s2[e] : s2[e] + s1[e];
      N::st(tree + SUMP * DP, lane, s1);
      if (dirn == 1) {
        N::cp(tree + PQ * DP, cur + PQ * DP, lane);
        N::cp(tree + PP * DP, cur + PP * DP, lane);
        N::cp(tree + PVEL * DP, cur + PVEL * DP, lane);
        w_tree = nuts_add_w(slice, w_tree, w_cur);
      } else {
        N::cp(tree + NQ * DP, cur + NQ * DP, lane);
        N::cp(tree + NP * DP, cur + NP * DP, lane);
        N::cp(tree + NVEL * DP, cur + NVEL * DP, lane);
        w_tree = nuts_add_w(slice, w_cur, w_tree);
      }
      __syncwarp();
      if (stop) break;
    }
    if (depth == a.max_depth) depth = a.max_depth - 1;  // `for depth in range(...)` ran out
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
  }
}

#undef MB200_NUTS_MATVEC

}  // namespace mb200
