// N4: dynamic-length HMC transitions ("NUTS") for ANY integrator / system pair -- the tree
// bookkeeping of DynamicIntegrationTransition (transitions.py:528-581, 610-770) as device kernels
// around the batched integrator step.
//
// The fused kernel (nuts.cuh) evaluates the leaf's leapfrog step itself and is limited to
// LeapfrogIntegrator on Euclidean systems.  Here the leaves come from outside: the host advances
// ALL chains by one batched `Integrator.step` per leaf (implicit, constrained, composition ...:
// whatever kernel the integrator launches), evaluates `system.h` and `system.dh_dmom` of the new
// states, and these kernels consume them -- one warp per chain, the per-chain tree in the same
// global workspace layout as nuts.cuh, the scalar state of the transition in a per-chain record
// that persists between launches.  Chains run their doublings in lock-step (2^depth leaves per
// doubling for every chain that is still growing its tree); a chain whose tree has terminated
// ignores the remaining leaves.  A failed integrator step (status != 0) terminates the tree and
// sets the matching flag, as the `except IntegratorError` of transitions.py:670-676 does.
//
//   nuts_generic_begin        initial state -> tree, [slice variable], flags
//   nuts_generic_start        per doubling: direction (one uniform), edge state out, active mask
//   nuts_generic_leaf         per leaf: weights, divergence test, binary-counter merges
//   nuts_generic_finish       per doubling: progressive sampling, tree merge, termination
//   nuts_generic_end          returned state and statistics
#pragma once
#include "nuts.cuh"

namespace mb200 {

struct NutsGenState {
  double h_init, log_u, w_tree, h_next, sum_accept, reject_prob, w_cur, h_cur;
  double lw[NUTS_MAX_DEPTH], lh[NUTS_MAX_DEPTH];
  int n_used, n_step, depth, dirn;
  int done, aborted;                 // transition finished / current doubling terminated early
  int diverging, conv_err, nonrev, starved;
  int next_is_init, pos_is_init, neg_is_init, init_dir, next_dir;
  int pad;
};

struct NutsGenArgs {
  int max_depth, slice, euclid, extra;
  double max_delta_h;
  const double* uniforms;
  int n_uniforms;
};

template <int KP>
struct NutsGen {
  using N = Nuts<StdGaussianTarget, KP>;
  static constexpr int NV = 2 * KP;
  static constexpr int DP = 64 * KP;

  static __device__ __forceinline__ void load_row(const double* src, int64_t ch, int dim, int lane,
                                                  double (&a)[NV]) {
#pragma unroll
    for (int e = 0; e < NV; ++e) {
      const int i = 2 * lane + 64 * (e >> 1) + (e & 1);
      a[e] = (i < dim) ? src[(size_t)ch * dim + i] : 0.0;
    }
  }
  static __device__ __forceinline__ void store_row(double* dst, int64_t ch, int dim, int lane,
                                                   const double (&a)[NV]) {
#pragma unroll
    for (int e = 0; e < NV; ++e) {
      const int i = 2 * lane + 64 * (e >> 1) + (e & 1);
      if (i < dim) dst[(size_t)ch * dim + i] = a[e];
    }
  }
  static __device__ __forceinline__ double uniform(NutsGenState& s, const NutsGenArgs& a,
                                                   int64_t ch) {
    if (s.n_used >= a.n_uniforms) {
      s.starved = 1;
      return 0.5;
    }
    return a.uniforms[(size_t)ch * a.n_uniforms + s.n_used++];
  }
  static __device__ __forceinline__ double leaf_weight(const NutsGenState& s, bool slice, double h) {
    return slice ? ((s.log_u <= -h) ? 1.0 : 0.0) : -h;
  }
};

template <int KP>
__global__ void __launch_bounds__(128)
    nuts_generic_begin_kernel(const double* __restrict__ q_in, const double* __restrict__ p_in,
                              const double* __restrict__ v_in, const double* __restrict__ h_in,
                              int64_t n_chains, int dim, NutsGenArgs a, double* __restrict__ ws,
                              NutsGenState* __restrict__ states) {
  using G = NutsGen<KP>;
  using N = typename G::N;
  constexpr int NV = G::NV, DP = G::DP;
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const size_t ws_stride = (size_t)(7 + 2 + NUTS_REC * (1 + a.max_depth)) * DP;
  for (int64_t ch = (int64_t)blockIdx.x * wpb + (threadIdx.x >> 5); ch < n_chains;
       ch += (int64_t)gridDim.x * wpb) {
    double* tree = ws + (size_t)ch * ws_stride;
    double* next = tree + 7 * DP;
    double q[NV], p[NV], v[NV];
    G::load_row(q_in, ch, dim, lane, q);
    G::load_row(p_in, ch, dim, lane, p);
    G::load_row(v_in, ch, dim, lane, v);
    NutsGenState s;
    s.h_init = h_in[ch];
    s.n_used = 0, s.starved = 0;
    s.log_u = a.slice ? log(G::uniform(s, a, ch)) - s.h_init : 0.0;  // transitions.py:832-839
    s.w_tree = G::leaf_weight(s, a.slice != 0, s.h_init);
    s.h_next = s.h_init;
    s.sum_accept = 0.0, s.reject_prob = 1.0, s.w_cur = 0.0, s.h_cur = 0.0;
    for (int l = 0; l < NUTS_MAX_DEPTH; ++l) s.lw[l] = 0.0, s.lh[l] = 0.0;
    s.n_step = 0, s.depth = 0, s.dirn = 1, s.done = 0, s.aborted = 0;
    s.diverging = 0, s.conv_err = 0, s.nonrev = 0;
    s.next_is_init = 1, s.pos_is_init = 1, s.neg_is_init = 1, s.init_dir = 1, s.next_dir = 1;
    s.pad = 0;
    N::st(tree + NQ * DP, lane, q), N::st(tree + PQ * DP, lane, q);
    N::st(tree + NP * DP, lane, p), N::st(tree + PP * DP, lane, p);
    N::st(tree + SUMP * DP, lane, p);
    N::st(tree + NVEL * DP, lane, v), N::st(tree + PVEL * DP, lane, v);
    N::st(next, lane, q), N::st(next + DP, lane, p);
    if (lane == 0) states[ch] = s;
  }
}

// Start of doubling `depth`: draw the direction, hand the edge state to the integrator.
template <int KP>
__global__ void __launch_bounds__(128)
    nuts_generic_start_kernel(int64_t n_chains, int dim, int depth, NutsGenArgs a,
                              double* __restrict__ ws, NutsGenState* __restrict__ states,
                              double* __restrict__ q_edge, double* __restrict__ p_edge,
                              int32_t* __restrict__ dir_out, int32_t* __restrict__ active) {
  using G = NutsGen<KP>;
  using N = typename G::N;
  constexpr int NV = G::NV, DP = G::DP;
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const size_t ws_stride = (size_t)(7 + 2 + NUTS_REC * (1 + a.max_depth)) * DP;
  for (int64_t ch = (int64_t)blockIdx.x * wpb + (threadIdx.x >> 5); ch < n_chains;
       ch += (int64_t)gridDim.x * wpb) {
    NutsGenState s = states[ch];
    if (s.done) {
      if (lane == 0) active[ch] = 0, dir_out[ch] = 1;
      continue;
    }
    double* tree = ws + (size_t)ch * ws_stride;
    const int dirn = (G::uniform(s, a, ch) < 0.5) ? 1 : -1;  // transitions.py:729
    if (dirn == 1 ? s.pos_is_init : s.neg_is_init) s.init_dir = dirn;
    s.dirn = dirn, s.depth = depth, s.aborted = 0, s.w_cur = 0.0, s.h_cur = 0.0;
    double q[NV], p[NV];
    N::ld(tree + (dirn == 1 ? PQ : NQ) * DP, lane, q);
    N::ld(tree + (dirn == 1 ? PP : NP) * DP, lane, p);
    G::store_row(q_edge, ch, dim, lane, q);
    G::store_row(p_edge, ch, dim, lane, p);
    if (lane == 0) {
      states[ch] = s;
      active[ch] = 1;
      dir_out[ch] = dirn;
    }
  }
}

// Leaf number k (1-based) of a doubling of n_leaves leaves: the new state (q, p), its velocity
// v = dh_dmom, energy h and the integrator's status.
template <int KP>
__global__ void __launch_bounds__(128)
    nuts_generic_leaf_kernel(const double* __restrict__ q_in, const double* __restrict__ p_in,
                             const double* __restrict__ v_in, const double* __restrict__ h_in,
                             const int32_t* __restrict__ status_in, int64_t n_chains, int dim,
                             int k, int n_leaves, NutsGenArgs a, double* __restrict__ ws,
                             NutsGenState* __restrict__ states, int32_t* __restrict__ active) {
  using G = NutsGen<KP>;
  using N = typename G::N;
  constexpr int NV = G::NV, DP = G::DP;
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const bool slice = a.slice != 0, euclid = a.euclid != 0, extra = a.extra != 0;
  const size_t ws_stride = (size_t)(7 + 2 + NUTS_REC * (1 + a.max_depth)) * DP;
  for (int64_t ch = (int64_t)blockIdx.x * wpb + (threadIdx.x >> 5); ch < n_chains;
       ch += (int64_t)gridDim.x * wpb) {
    NutsGenState s = states[ch];
    if (s.done || s.aborted) continue;
    double* tree = ws + (size_t)ch * ws_stride;
    double* cur = tree + 9 * DP;
    double* levels = cur + NUTS_REC * DP;
    const int dirn = s.dirn;
    const int st = status_in[ch];
    if (st != 0) {  // IntegratorError inside the step: the tree is terminated (transitions.py:670)
      if (st == MB200_STATUS_NON_REVERSIBLE) s.nonrev = 1;
      else s.conv_err = 1;
      s.aborted = 1;
      if (lane == 0) states[ch] = s, active[ch] = 0;
      continue;
    }
    double q[NV], p[NV], v[NV];
    G::load_row(q_in, ch, dim, lane, q);
    G::load_row(p_in, ch, dim, lane, p);
    G::load_row(v_in, ch, dim, lane, v);
    double h = h_in[ch];
    if (h != h) h = INFINITY;  // transitions.py:626
    double w_cur = G::leaf_weight(s, slice, h), h_cur = h;
    const bool parked_leaf = (k & 1) && k < n_leaves;
    double* leaf = parked_leaf ? levels : cur;
    N::st(leaf + NQ * DP, lane, q), N::st(leaf + PQ * DP, lane, q), N::st(leaf + RQ * DP, lane, q);
    N::st(leaf + NP * DP, lane, p), N::st(leaf + PP * DP, lane, p);
    N::st(leaf + RP * DP, lane, p), N::st(leaf + SUMP * DP, lane, p);
    N::st(leaf + NVEL * DP, lane, v), N::st(leaf + PVEL * DP, lane, v);
    const double h_diff = s.h_init - h;
    s.sum_accept += (h_diff != h_diff) ? 0.0 : exp(fmin(0.0, h_diff));
    ++s.n_step;
    bool terminate = false;
    if ((slice ? h + s.log_u : h - s.h_init) > a.max_delta_h) {  // _check_divergence
      s.diverging = 1;
      terminate = true;
    }
    if (!terminate) {
      if (parked_leaf) {
        s.lw[0] = w_cur, s.lh[0] = h_cur;
      } else {
        int level = 0;
        for (int kk = k; (kk & 1) == 0; kk >>= 1, ++level) {
          double* inner = levels + (size_t)level * NUTS_REC * DP;
          const double w_new = nuts_add_w(slice, dirn == 1 ? s.lw[level] : w_cur,
                                          dirn == 1 ? w_cur : s.lw[level]);
          const bool take_outer = G::uniform(s, a, ch) < nuts_ratio(slice, w_cur, w_new);
          const double* neg = dirn == 1 ? inner : cur;
          const double* pos = dirn == 1 ? cur : inner;
          __syncwarp();
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
          for (int e = 0; e < NV; ++e) s1[e] = dirn == 1 ? s2[e] + s1[e] : s1[e] + s2[e];
          N::st(cur + SUMP * DP, lane, s1);
          if (!take_outer) {
            N::cp(cur + RQ * DP, inner + RQ * DP, lane);
            N::cp(cur + RP * DP, inner + RP * DP, lane);
            h_cur = s.lh[level];
          }
          __syncwarp();
          w_cur = w_new;
          if (stop) {
            terminate = true;
            break;
          }
        }
        if (!terminate && k < n_leaves) {  // park the completed subtree on its level
          double* slot = levels + (size_t)level * NUTS_REC * DP;
#pragma unroll
          for (int r = 0; r < NUTS_REC; ++r) N::cp(slot + r * DP, cur + r * DP, lane);
          s.lw[level] = w_cur;
          s.lh[level] = h_cur;
        }
      }
    }
    s.w_cur = w_cur, s.h_cur = h_cur;
    if (terminate) s.aborted = 1;
    if (lane == 0) {
      states[ch] = s;
      if (terminate) active[ch] = 0;
    }
  }
}

// End of a doubling: progressive sampling of the next state (transitions.py:742-749), merge into
// the tree, no-U-turn test of the whole tree.
template <int KP>
__global__ void __launch_bounds__(128)
    nuts_generic_finish_kernel(int64_t n_chains, int dim, int depth, NutsGenArgs a,
                               double* __restrict__ ws, NutsGenState* __restrict__ states) {
  using G = NutsGen<KP>;
  using N = typename G::N;
  constexpr int NV = G::NV, DP = G::DP;
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const bool slice = a.slice != 0, euclid = a.euclid != 0, extra = a.extra != 0;
  const size_t ws_stride = (size_t)(7 + 2 + NUTS_REC * (1 + a.max_depth)) * DP;
  for (int64_t ch = (int64_t)blockIdx.x * wpb + (threadIdx.x >> 5); ch < n_chains;
       ch += (int64_t)gridDim.x * wpb) {
    NutsGenState s = states[ch];
    if (s.done) continue;
    if (s.aborted) {  // `if terminate: break` -- the transition ends at this depth
      s.done = 1;
      if (lane == 0) states[ch] = s;
      continue;
    }
    double* tree = ws + (size_t)ch * ws_stride;
    double* next = tree + 7 * DP;
    double* cur = next + 2 * DP;
    const int dirn = s.dirn;
    const double accept_prob = nuts_ratio(slice, s.w_cur, s.w_tree);
    if (G::uniform(s, a, ch) < accept_prob) {
      N::cp(next, cur + RQ * DP, lane);
      N::cp(next + DP, cur + RP * DP, lane);
      s.h_next = s.h_cur;
      s.next_is_init = 0, s.next_dir = dirn;
    }
    s.reject_prob *= 1.0 - accept_prob;
    if (dirn == 1) s.pos_is_init = 0; else s.neg_is_init = 0;
    const double* neg = dirn == 1 ? tree : cur;
    const double* pos = dirn == 1 ? cur : tree;
    __syncwarp();
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
      s.w_tree = nuts_add_w(slice, s.w_tree, s.w_cur);
    } else {
      N::cp(tree + NQ * DP, cur + NQ * DP, lane);
      N::cp(tree + NP * DP, cur + NP * DP, lane);
      N::cp(tree + NVEL * DP, cur + NVEL * DP, lane);
      s.w_tree = nuts_add_w(slice, s.w_cur, s.w_tree);
    }
    if (stop || depth + 1 >= a.max_depth) s.done = 1;
    if (lane == 0) states[ch] = s;
  }
}

template <int KP>
__global__ void __launch_bounds__(128)
    nuts_generic_end_kernel(int64_t n_chains, int dim, NutsGenArgs a, const double* __restrict__ ws,
                            const NutsGenState* __restrict__ states, double* __restrict__ q_out,
                            double* __restrict__ p_out, double* __restrict__ h_out,
                            int32_t* __restrict__ n_step_out, double* __restrict__ av_accept_out,
                            double* __restrict__ reject_prob_out, int32_t* __restrict__ depth_out,
                            int32_t* __restrict__ flags_out, int32_t* __restrict__ n_used_out,
                            int32_t* __restrict__ dir_out) {
  using G = NutsGen<KP>;
  using N = typename G::N;
  constexpr int NV = G::NV, DP = G::DP;
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const size_t ws_stride = (size_t)(7 + 2 + NUTS_REC * (1 + a.max_depth)) * DP;
  for (int64_t ch = (int64_t)blockIdx.x * wpb + (threadIdx.x >> 5); ch < n_chains;
       ch += (int64_t)gridDim.x * wpb) {
    const NutsGenState s = states[ch];
    const double* next = ws + (size_t)ch * ws_stride + 7 * DP;
    double q[NV], p[NV];
    N::ld(next, lane, q);
    N::ld(next + DP, lane, p);
    G::store_row(q_out, ch, dim, lane, q);
    G::store_row(p_out, ch, dim, lane, p);
    if (lane == 0) {
      h_out[ch] = s.h_next;
      n_step_out[ch] = s.n_step;
      av_accept_out[ch] = s.n_step > 0 ? s.sum_accept / s.n_step : 0.0;
      reject_prob_out[ch] = s.reject_prob;
      depth_out[ch] = s.depth;
      // bit 0 diverging, 1 convergence_error, 2 non_reversible_step, 3 ran out of uniforms
      flags_out[ch] = s.diverging | (s.conv_err << 1) | (s.nonrev << 2) | (s.starved << 3);
      n_used_out[ch] = s.n_used;
      dir_out[ch] = s.next_is_init ? s.init_dir : s.next_dir;
    }
  }
}

}  // namespace mb200
