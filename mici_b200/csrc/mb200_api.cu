// libmici_b200.so -- C-ABI entry points (include/mici_b200.h).  Host-side argument checking and
// kernel dispatch only; all arithmetic is in the .cuh kernels.
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "common.cuh"
#include "leapfrog_generic.cuh"
#ifndef MB200_NO_DMMA
#include "leapfrog_dmma.cuh"
#ifdef MB200_DMMA_V2  // experimental second-generation kernel (slower: profiles/r01_notes.md)
#include "leapfrog_dmma2.cuh"
#endif
#ifdef MB200_DMMA_V3  // warp-specialised generation (drift warps / update warps)
#include "leapfrog_dmma3.cuh"
#endif
#endif
#ifndef MB200_NO_CONSTRAINED
#include "constrained.cuh"
#endif
#ifndef MB200_NO_RIEMANNIAN
#include "riemannian.cuh"
#endif
#include "nuts.cuh"
#include "transitions.cuh"

namespace mb200 {

static thread_local char g_err[512] = "";

static int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

static int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(MB200_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
  return 0;
}

// Per-chain overrides travel from the *_per_chain entry points to the kernels through the calling
// thread only (thread-local, scoped): the plain entry points stay re-entrant and unchanged.
static thread_local const double* tl_step_sizes = nullptr;
static thread_local const int32_t* tl_n_steps = nullptr;
struct PerChainScope {
  PerChainScope(const double* eps, const int32_t* ns) { tl_step_sizes = eps, tl_n_steps = ns; }
  ~PerChainScope() { tl_step_sizes = nullptr, tl_n_steps = nullptr; }
};

static ModelArgs to_args(const mb200_model* m) {
  ModelArgs a;
  memset(&a, 0, sizeof(a));
  a.step_sizes = tl_step_sizes;
  a.n_steps_pc = tl_n_steps;
  a.target_id = m->target_id;
  for (int i = 0; i < MB200_MAX_PARAMS; ++i) a.tp[i] = m->target_params[i];
  a.taux = m->target_aux;
  a.rmetric_id = m->rmetric_id;
  for (int i = 0; i < MB200_MAX_PARAMS; ++i) a.mp[i] = m->rmetric_params[i];
  a.maux = m->rmetric_aux;
  return a;
}

static int num_sms() {
  static thread_local int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  return sms;
}

template <class Target, int KP, int CPW, bool GAUSS = false>
static int launch_generic(const double* q_in, const double* p_in, double* q_out, double* p_out,
                          const int32_t* dir, int64_t n, int dim, double eps, int n_steps,
                          const FlowSchedule& sched, int metric_kind, const double* minv, const ModelArgs& m, double* h_out,
                          int32_t* status, int32_t* n_done, cudaStream_t st) {
  constexpr int WARPS = 4;
  auto kern = leapfrog_generic_kernel<Target, KP, CPW, GAUSS>;
  const size_t smem = (size_t)WARPS * CPW * 64 * KP * sizeof(double);
  if (smem > 48 * 1024) {
    cudaError_t e =
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(MB200_ERR_CUDA, "smem attr: %s", cudaGetErrorString(e));
  }
  const int64_t groups = (n + CPW - 1) / CPW;
  int64_t blocks = (groups + WARPS - 1) / WARPS;
  const int64_t cap = (int64_t)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  kern<<<(unsigned)blocks, WARPS * 32, smem, st>>>(q_in, p_in, q_out, p_out, dir, n, dim, eps,
                                                   n_steps, sched, metric_kind, minv, m, h_out,
                                                   status, n_done);
  return check_launch("leapfrog_generic_kernel");
}

template <class Target>
static int dispatch_generic_dim(const double* q_in, const double* p_in, double* q_out,
                                double* p_out, const int32_t* dir, int64_t n, int dim, double eps,
                                int n_steps, const FlowSchedule& sched, int metric_kind, const double* minv,
                                const ModelArgs& m, double* h_out, int32_t* status,
                                int32_t* n_done, cudaStream_t st) {
#define MB200_GEN(KP, CPW)                                                                      \
  return sched.gaussian                                                                         \
             ? launch_generic<Target, KP, CPW, true>(q_in, p_in, q_out, p_out, dir, n, dim, eps, \
                                                     n_steps, sched, metric_kind, minv, m,      \
                                                     h_out, status, n_done, st)                 \
             : launch_generic<Target, KP, CPW, false>(q_in, p_in, q_out, p_out, dir, n, dim,    \
                                                      eps, n_steps, sched, metric_kind, minv,   \
                                                      m, h_out, status, n_done, st)
  if (dim <= 64) MB200_GEN(1, 4);
  if (dim <= 128) MB200_GEN(2, 4);
  if (dim <= 256) MB200_GEN(4, 2);
  if (dim <= 512) MB200_GEN(8, 1);
  if (dim <= 1024) MB200_GEN(16, 1);
#undef MB200_GEN
  return fail(MB200_ERR_UNSUPPORTED, "dim %d > 1024 not supported by the Euclidean leapfrog", dim);
}

static FlowSchedule leapfrog_schedule() {
  FlowSchedule s;
  memset(&s, 0, sizeof(s));
  s.n = 3;
  s.drift_mask = 0x2u;
  s.coef[0] = 0.5, s.coef[1] = 1.0, s.coef[2] = 0.5;
  return s;
}

static int leapfrog_euclidean_impl(const double* q_in, const double* p_in, double* q_out,
                                   double* p_out, const int32_t* dir, int64_t n, int dim,
                                   double eps, int n_steps, int metric_kind, const double* minv,
                                   const mb200_model* model, double* h_out, int32_t* status,
                                   int32_t* n_done, cudaStream_t st, bool allow_dmma,
                                   const FlowSchedule* schedule = nullptr) {
  const FlowSchedule sched = schedule ? *schedule : leapfrog_schedule();
  if (schedule) allow_dmma = false;
  if (n == 0 && dim >= 1 && n_steps >= 0) return 0;  // empty batch: nothing to do
  if (!q_in || !p_in || !q_out || !p_out || !model)
    return fail(MB200_ERR_INVALID_ARG, "null pointer argument");
  if (n < 0 || dim < 1 || n_steps < 0) return fail(MB200_ERR_INVALID_ARG, "bad sizes");
  if (metric_kind < 0 || metric_kind > 2) return fail(MB200_ERR_INVALID_ARG, "bad metric_kind");
  if (metric_kind != MB200_METRIC_IDENTITY && !minv)
    return fail(MB200_ERR_INVALID_ARG, "metric_inv is NULL");
  if (n == 0) return 0;
  const ModelArgs m = to_args(model);
  if (m.target_id == MB200_TARGET_BANANA && (dim & 1))
    return fail(MB200_ERR_INVALID_ARG, "banana target needs even dim");
#ifndef MB200_NO_DMMA
  if (allow_dmma && metric_kind == MB200_METRIC_DENSE && n_steps > 0) {
#if defined(MB200_DMMA_V3)
    int rc = leapfrog_dmma3_dispatch(q_in, p_in, q_out, p_out, dir, n, dim, eps, n_steps, minv, m,
                                     h_out, status, n_done, st);
#elif defined(MB200_DMMA_V2)
    int rc = leapfrog_dmma2_dispatch(q_in, p_in, q_out, p_out, dir, n, dim, eps, n_steps, minv, m,
                                     h_out, status, n_done, st);
#else
    int rc = leapfrog_dmma_dispatch(q_in, p_in, q_out, p_out, dir, n, dim, eps, n_steps, minv, m,
                                    h_out, status, n_done, st);
#endif
    if (rc == 0) return check_launch("leapfrog_dmma_kernel");
    if (rc != MB200_ERR_UNSUPPORTED) return fail(rc, "leapfrog_dmma launch failed");
  }
#endif
#define MB200_ARGS                                                                         \
  q_in, p_in, q_out, p_out, dir, n, dim, eps, n_steps, sched, metric_kind, minv, m, h_out, \
      status, n_done, st
  switch (m.target_id) {
    case MB200_TARGET_STD_GAUSSIAN:
      return dispatch_generic_dim<StdGaussianTarget>(MB200_ARGS);
    case MB200_TARGET_NEAL_FUNNEL:
      return dispatch_generic_dim<NealFunnelTarget>(MB200_ARGS);
    case MB200_TARGET_BANANA:
      return dispatch_generic_dim<BananaTarget>(MB200_ARGS);
    default:
      return fail(MB200_ERR_UNSUPPORTED, "target %d not available for Euclidean leapfrog",
                  m.target_id);
  }
#undef MB200_ARGS
}


template <class Target, int KP>
static int launch_eval(const double* q, const double* p, int64_t n, int dim, int metric_kind,
                       const double* minv, const ModelArgs& m, double* nld, double* grad,
                       double* vel, double* kin, cudaStream_t st) {
  constexpr int WARPS = 4;
  auto kern = euclidean_eval_kernel<Target, KP>;
  const size_t smem = (size_t)WARPS * 64 * KP * sizeof(double);
  int64_t blocks = (n + WARPS - 1) / WARPS;
  const int64_t cap = (int64_t)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  kern<<<(unsigned)blocks, WARPS * 32, smem, st>>>(q, p, n, dim, metric_kind, minv, m, nld, grad,
                                                   vel, kin);
  return check_launch("euclidean_eval_kernel");
}

template <class Target>
static int dispatch_eval_dim(const double* q, const double* p, int64_t n, int dim,
                             int metric_kind, const double* minv, const ModelArgs& m, double* nld,
                             double* grad, double* vel, double* kin, cudaStream_t st) {
#define MB200_EV(KP) \
  return launch_eval<Target, KP>(q, p, n, dim, metric_kind, minv, m, nld, grad, vel, kin, st)
  if (dim <= 64) MB200_EV(1);
  if (dim <= 128) MB200_EV(2);
  if (dim <= 256) MB200_EV(4);
  if (dim <= 512) MB200_EV(8);
  if (dim <= 1024) MB200_EV(16);
#undef MB200_EV
  return fail(MB200_ERR_UNSUPPORTED, "dim %d > 1024 not supported", dim);
}

#ifndef MB200_NO_CONSTRAINED
template <class Target, int KP>
static int launch_constrained(const double* q_in, const double* p_in, double* q_out,
                              double* p_out, const int32_t* dir, int64_t n, int dim, double eps,
                              int n_steps, int n_inner, int metric_kind, const double* minv,
                              const ModelArgs& m, double ctol, double ptol, double dtol,
                              int max_iters, double rev_tol, double* h_out, int32_t* status,
                              int32_t* n_done, int32_t* iters, cudaStream_t st, int proj_solver,
                              int max_ls) {
  constexpr int WARPS = 4;
  auto kern = constrained_leapfrog_kernel<Target, KP>;
  const size_t smem = (size_t)WARPS * (Target::NC > 1 ? Target::NC : 1) * 64 * KP * sizeof(double);
  int64_t blocks = (n + WARPS - 1) / WARPS;
  const int64_t cap = (int64_t)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  kern<<<(unsigned)blocks, WARPS * 32, smem, st>>>(q_in, p_in, q_out, p_out, dir, n, dim, eps,
                                                   n_steps, n_inner, metric_kind, minv, m, ctol,
                                                   ptol, dtol, max_iters, rev_tol, h_out, status,
                                                   n_done, iters, proj_solver, max_ls);
  return check_launch("constrained_leapfrog_kernel");
}
#endif

#ifndef MB200_NO_RIEMANNIAN
template <class Target, template <class> class MetricT>
static int launch_implicit(const double* q_in, const double* p_in, double* q_out, double* p_out,
                           const int32_t* dir, int64_t n, int dim, double eps, int n_steps,
                           const ModelArgs& m, double fp_tol, double fp_div, int fp_max,
                           double rev_tol, double* h_out, int32_t* status, int32_t* n_done,
                           int32_t* fp_iters, cudaStream_t st, int midpoint, int fp_solver) {
  auto kern = implicit_leapfrog_kernel<Target, MetricT>;
  int n_mats = MetricT<Target>::N_MATS;
  // SoftAbs: a third matrix enables warm-started eigensolves; use it when two CTAs still fit
  if (MetricT<Target>::SOFTABS && rm_smem_doubles(dim, 3) * sizeof(double) <= 113 * 1024) n_mats = 3;
  const size_t smem = rm_smem_doubles(dim, n_mats) * sizeof(double);
  if (smem > 227 * 1024)
    return fail(MB200_ERR_UNSUPPORTED,
                "dim %d: per-chain metric (%zu bytes) exceeds shared memory; not supported yet",
                dim, smem);
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return fail(MB200_ERR_CUDA, "smem attr: %s", cudaGetErrorString(e));
  int per_sm = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, RM_THREADS, smem);
  if (per_sm < 1) per_sm = 1;
  int64_t blocks = (int64_t)num_sms() * per_sm;
  if (blocks > n) blocks = n;
  kern<<<(unsigned)blocks, RM_THREADS, smem, st>>>(q_in, p_in, q_out, p_out, dir, n, dim, eps,
                                                   n_steps, m, fp_tol, fp_div, fp_max, rev_tol,
                                                   h_out, status, n_done, fp_iters, n_mats,
                                                   midpoint, fp_solver);
  return check_launch("implicit_leapfrog_kernel");
}

static int implicit_dispatch(const double* q_in, const double* p_in, double* q_out, double* p_out,
                             const int32_t* dir, int64_t n, int dim, double eps, int n_steps,
                             const ModelArgs& m, double fp_tol, double fp_div, int fp_max,
                             double rev_tol, double* h_out, int32_t* status, int32_t* n_done,
                             int32_t* fp_iters, cudaStream_t st, int midpoint = 0,
                             int fp_solver = 0) {
  if (fp_solver != MB200_FP_SOLVER_DIRECT && fp_solver != MB200_FP_SOLVER_STEFFENSEN)
    return fail(MB200_ERR_INVALID_ARG, "unknown fixed-point solver %d", fp_solver);
#define MB200_ARGS                                                                           \
  q_in, p_in, q_out, p_out, dir, n, dim, eps, n_steps, m, fp_tol, fp_div, fp_max, rev_tol,   \
      h_out, status, n_done, fp_iters, st, midpoint, fp_solver
  if (m.rmetric_id == MB200_RMETRIC_SOFTABS) {
    if (!(m.mp[0] > 0.0)) return fail(MB200_ERR_INVALID_ARG, "softabs_coeff must be positive");
    switch (m.target_id) {
      case MB200_TARGET_BANANA:
        if (dim & 1) return fail(MB200_ERR_INVALID_ARG, "banana target needs even dim");
        return launch_implicit<BananaRTarget, SoftAbsMetric>(MB200_ARGS);
      default:
        return fail(MB200_ERR_UNSUPPORTED, "target %d has no device Hessian / MTP (SoftAbs metric)",
                    m.target_id);
    }
  }
  if (m.rmetric_id == MB200_RMETRIC_RANK1) {
    if (!m.maux) return fail(MB200_ERR_INVALID_ARG, "rank-1 metric needs its base matrix (rmetric_aux)");
    if (m.target_id == MB200_TARGET_QUADRATIC && !m.taux)
      return fail(MB200_ERR_INVALID_ARG, "quadratic target needs its precision matrix");
    if (m.target_id == MB200_TARGET_BANANA && (dim & 1))
      return fail(MB200_ERR_INVALID_ARG, "banana target needs even dim");
    // per-chain Cholesky factor in shared memory when it fits (or when forced), else the
    // Sherman-Morrison form that never materialises M(q); mp[2] != 0 forces the latter
    const bool fits = rm_smem_doubles(dim, 1) * sizeof(double) <= 227 * 1024;
    const bool woodbury = !fits || m.mp[2] != 0.0;
    switch (m.target_id) {
      case MB200_TARGET_QUADRATIC:
        return woodbury ? launch_implicit<QuadraticRTarget, Rank1WoodburyMetric>(MB200_ARGS)
                        : launch_implicit<QuadraticRTarget, Rank1DenseMetric>(MB200_ARGS);
      case MB200_TARGET_STD_GAUSSIAN:
        return woodbury ? launch_implicit<StdGaussianRTarget, Rank1WoodburyMetric>(MB200_ARGS)
                        : launch_implicit<StdGaussianRTarget, Rank1DenseMetric>(MB200_ARGS);
      case MB200_TARGET_BANANA:
        return woodbury ? launch_implicit<BananaRTarget, Rank1WoodburyMetric>(MB200_ARGS)
                        : launch_implicit<BananaRTarget, Rank1DenseMetric>(MB200_ARGS);
      default:
        return fail(MB200_ERR_UNSUPPORTED, "target %d not available for Riemannian systems", m.target_id);
    }
  }
#undef MB200_ARGS
  return fail(MB200_ERR_INVALID_ARG, "unknown rmetric_id %d", m.rmetric_id);
}
#endif

#ifndef MB200_NO_CONSTRAINED
template <class Target, int KP>
static int launch_project(const double* q, const double* p_in, double* p_out, int64_t n, int dim,
                          int metric_kind, const double* minv, const ModelArgs& m,
                          cudaStream_t st) {
  constexpr int WARPS = 4;
  const size_t smem = (size_t)WARPS * (Target::NC > 1 ? Target::NC : 1) * 64 * KP * sizeof(double);
  int64_t blocks = (n + WARPS - 1) / WARPS;
  const int64_t cap = (int64_t)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  constrained_project_kernel<Target, KP><<<(unsigned)blocks, WARPS * 32, smem, st>>>(
      q, p_in, p_out, n, dim, metric_kind, minv, m);
  return check_launch("constrained_project_kernel");
}
#endif

#ifndef MB200_NO_RIEMANNIAN
template <class Target, template <class> class MetricT>
static int launch_sample_momentum(const double* q, const double* z, double* p_out, int64_t n,
                                  int dim, const ModelArgs& m, int32_t* status, cudaStream_t st) {
  auto kern = riemannian_sample_momentum_kernel<Target, MetricT>;
  const int n_mats = MetricT<Target>::N_MATS;
  const size_t smem = rm_smem_doubles(dim, n_mats) * sizeof(double);
  if (smem > 227 * 1024) return fail(MB200_ERR_UNSUPPORTED, "dim %d too large", dim);
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return fail(MB200_ERR_CUDA, "smem attr: %s", cudaGetErrorString(e));
  int64_t blocks = (int64_t)num_sms() * 2;
  if (blocks > n) blocks = n;
  kern<<<(unsigned)blocks, RM_THREADS, smem, st>>>(q, z, p_out, n, dim, m, status, n_mats);
  return check_launch("riemannian_sample_momentum_kernel");
}
#endif

template <class Target, int KP>
static int launch_nuts(const double* q_in, const double* p_in, double* q_out, double* p_out,
                       int64_t n, int dim, double eps, int metric_kind, const double* minv,
                       const ModelArgs& m, const NutsArgs& a, double* ws, double* h_out,
                       int32_t* n_step, double* av_accept, double* reject_prob, int32_t* depth,
                       int32_t* diverging, int32_t* n_used, int32_t* dir_out, int32_t* status,
                       cudaStream_t st) {
  auto kern = nuts_euclidean_kernel<Target, KP>;
  NutsArgs args = a;
  // dense metric that fits in shared memory next to the staging rows: the warps of a CTA share it
  // (12 warps for 64 < dim <= 128, where one CTA per SM fits; 8 otherwise -- measured)
  const size_t metric_bytes = (size_t)dim * dim * sizeof(double);
  const int staged_warps = KP == 2 ? 12 : 8;
  args.stage_metric = metric_kind == MB200_METRIC_DENSE &&
                      metric_bytes + staged_warps * 64 * KP * sizeof(double) <= 200 * 1024;
  const int warps = args.stage_metric ? staged_warps : 4;
  const size_t smem = (size_t)warps * 64 * KP * sizeof(double) + (args.stage_metric ? metric_bytes : 0);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(MB200_ERR_CUDA, "smem attr: %s", cudaGetErrorString(e));
  }
  int per_sm = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, warps * 32, smem);
  if (per_sm < 1) per_sm = 1;
  int64_t blocks = (n + warps - 1) / warps;
  const int64_t cap = (int64_t)num_sms() * per_sm;
  if (blocks > cap) blocks = cap;
  kern<<<(unsigned)blocks, warps * 32, smem, st>>>(
      q_in, p_in, q_out, p_out, n, dim, eps, metric_kind, minv, m, args, ws, h_out, n_step,
      av_accept, reject_prob, depth, diverging, n_used, dir_out, status);
  return check_launch("nuts_euclidean_kernel");
}

template <class Target>
static int dispatch_nuts_dim(const double* q_in, const double* p_in, double* q_out, double* p_out,
                             int64_t n, int dim, double eps, int metric_kind, const double* minv,
                             const ModelArgs& m, const NutsArgs& a, double* ws, double* h_out,
                             int32_t* n_step, double* av_accept, double* reject_prob,
                             int32_t* depth, int32_t* diverging, int32_t* n_used,
                             int32_t* dir_out, int32_t* status, cudaStream_t st) {
#define MB200_NUTS(KP)                                                                         \
  return launch_nuts<Target, KP>(q_in, p_in, q_out, p_out, n, dim, eps, metric_kind, minv, m, a, \
                                 ws, h_out, n_step, av_accept, reject_prob, depth, diverging,  \
                                 n_used, dir_out, status, st)
  if (dim <= 64) MB200_NUTS(1);
  if (dim <= 128) MB200_NUTS(2);
  if (dim <= 256) MB200_NUTS(4);
  if (dim <= 512) MB200_NUTS(8);
  if (dim <= 1024) MB200_NUTS(16);
#undef MB200_NUTS
  return fail(MB200_ERR_UNSUPPORTED, "dim %d > 1024 not supported", dim);
}

}  // namespace mb200

using namespace mb200;

extern "C" {

int mb200_version(void) { return MB200_VERSION; }

const char* mb200_last_error(void) { return g_err; }

int mb200_leapfrog_euclidean(const double* pos_in, const double* mom_in, double* pos_out,
                             double* mom_out, const int32_t* dir, int64_t n_chains, int32_t dim,
                             double step_size, int32_t n_steps, int32_t metric_kind,
                             const double* metric_inv, const mb200_model* model, double* h_out,
                             int32_t* status, int32_t* n_done, void* stream) {
  return leapfrog_euclidean_impl(pos_in, mom_in, pos_out, mom_out, dir, n_chains, dim, step_size,
                                 n_steps, metric_kind, metric_inv, model, h_out, status, n_done,
                                 (cudaStream_t)stream, true);
}

// Same arithmetic through the general-dimension kernel only (used by tests to cross-check the
// tensor-core kernel; not part of the reference-facing surface).
int mb200_leapfrog_euclidean_generic(const double* pos_in, const double* mom_in, double* pos_out,
                                     double* mom_out, const int32_t* dir, int64_t n_chains,
                                     int32_t dim, double step_size, int32_t n_steps,
                                     int32_t metric_kind, const double* metric_inv,
                                     const mb200_model* model, double* h_out, int32_t* status,
                                     int32_t* n_done, void* stream) {
  return leapfrog_euclidean_impl(pos_in, mom_in, pos_out, mom_out, dir, n_chains, dim, step_size,
                                 n_steps, metric_kind, metric_inv, model, h_out, status, n_done,
                                 (cudaStream_t)stream, false);
}

int mb200_hamiltonian_euclidean(const double* pos, const double* mom, int64_t n_chains,
                                int32_t dim, int32_t metric_kind, const double* metric_inv,
                                const mb200_model* model, double* h_out, void* stream) {
  if (n_chains == 0 && dim >= 1) return 0;
  if (!h_out) return fail(MB200_ERR_INVALID_ARG, "h_out is NULL");
  // zero leapfrog steps: loads the state, evaluates h, writes the (unchanged) state back in place
  return leapfrog_euclidean_impl(pos, mom, const_cast<double*>(pos), const_cast<double*>(mom),
                                 nullptr, n_chains, dim, 0.0, 0, metric_kind, metric_inv, model,
                                 h_out, nullptr, nullptr, (cudaStream_t)stream, false);
}

int mb200_euclidean_eval(const double* pos, const double* mom, int64_t n_chains, int32_t dim,
                         int32_t metric_kind, const double* metric_inv, const mb200_model* model,
                         double* nld_out, double* grad_out, double* vel_out, double* kin_out,
                         void* stream) {
  if (n_chains == 0 && dim >= 1) return 0;
  if (!pos || !mom || !model) return fail(MB200_ERR_INVALID_ARG, "null pointer argument");
  if (n_chains < 0 || dim < 1) return fail(MB200_ERR_INVALID_ARG, "bad sizes");
  if (metric_kind < 0 || metric_kind > 2) return fail(MB200_ERR_INVALID_ARG, "bad metric_kind");
  if (metric_kind != MB200_METRIC_IDENTITY && !metric_inv)
    return fail(MB200_ERR_INVALID_ARG, "metric_inv is NULL");
  if (n_chains == 0) return 0;
  const ModelArgs m = to_args(model);
  cudaStream_t st = (cudaStream_t)stream;
#define MB200_ARGS pos, mom, n_chains, dim, metric_kind, metric_inv, m, nld_out, grad_out, vel_out, kin_out, st
  switch (m.target_id) {
    case MB200_TARGET_STD_GAUSSIAN: return dispatch_eval_dim<StdGaussianTarget>(MB200_ARGS);
    case MB200_TARGET_NEAL_FUNNEL: return dispatch_eval_dim<NealFunnelTarget>(MB200_ARGS);
    case MB200_TARGET_BANANA: return dispatch_eval_dim<BananaTarget>(MB200_ARGS);
    default:
      return fail(MB200_ERR_UNSUPPORTED, "target %d not available for Euclidean eval", m.target_id);
  }
#undef MB200_ARGS
}

int mb200_constrained_leapfrog_euclidean(
    const double* pos_in, const double* mom_in, double* pos_out, double* mom_out,
    const int32_t* dir, int64_t n_chains, int32_t dim, double step_size, int32_t n_steps,
    int32_t n_inner_step, int32_t metric_kind, const double* metric_inv, const mb200_model* model,
    int32_t projection_solver, double constraint_tol, double position_tol, double divergence_tol,
    int32_t max_iters, int32_t max_line_search_iters, double reverse_check_tol, double* h_out,
    int32_t* status, int32_t* n_done, int32_t* newton_iters, void* stream) {
#ifdef MB200_NO_CONSTRAINED
  return fail(MB200_ERR_UNSUPPORTED, "constrained leapfrog not compiled in");
#else
  if (n_chains == 0 && dim >= 1) return 0;
  if (!pos_in || !mom_in || !pos_out || !mom_out || !model)
    return fail(MB200_ERR_INVALID_ARG, "null pointer argument");
  if (n_chains < 0 || dim < 1 || n_steps < 0 || n_inner_step < 1 || max_iters < 0 ||
      max_line_search_iters < 0)
    return fail(MB200_ERR_INVALID_ARG, "bad sizes");
  if (projection_solver < 0 || projection_solver > 2)
    return fail(MB200_ERR_INVALID_ARG, "unknown projection solver %d", projection_solver);
  if (metric_kind < 0 || metric_kind > 2) return fail(MB200_ERR_INVALID_ARG, "bad metric_kind");
  if (metric_kind != MB200_METRIC_IDENTITY && !metric_inv)
    return fail(MB200_ERR_INVALID_ARG, "metric_inv is NULL");
  if (n_chains == 0) return 0;
  const ModelArgs m = to_args(model);
  cudaStream_t st = (cudaStream_t)stream;
#define MB200_ARGS                                                                              \
  pos_in, mom_in, pos_out, mom_out, dir, n_chains, dim, step_size, n_steps, n_inner_step,       \
      metric_kind, metric_inv, m, constraint_tol, position_tol, divergence_tol, max_iters,      \
      reverse_check_tol, h_out, status, n_done, newton_iters, st, projection_solver,            \
      max_line_search_iters
  switch (m.target_id) {
    case MB200_TARGET_TORUS:
      if (dim != 3) return fail(MB200_ERR_INVALID_ARG, "torus target needs dim == 3");
      return launch_constrained<TorusTarget, 1>(MB200_ARGS);
    case MB200_TARGET_SPHERE:
      if (dim <= 64) return launch_constrained<SphereTarget, 1>(MB200_ARGS);
      if (dim <= 128) return launch_constrained<SphereTarget, 2>(MB200_ARGS);
      if (dim <= 256) return launch_constrained<SphereTarget, 4>(MB200_ARGS);
      return fail(MB200_ERR_UNSUPPORTED, "sphere target: dim %d > 256 not supported", dim);
    default:
      return fail(MB200_ERR_UNSUPPORTED, "target %d defines no constraint", m.target_id);
  }
#undef MB200_ARGS
#endif
}

#ifdef MB200_NO_RIEMANNIAN
int mb200_implicit_leapfrog_riemannian(const double*, const double*, double*, double*,
                                       const int32_t*, int64_t, int32_t, double, int32_t,
                                       const mb200_model*, int32_t, double, double, int32_t, double,
                                       double*, int32_t*, int32_t*, int32_t*, void*, int64_t,
                                       void*) {
  return fail(MB200_ERR_UNSUPPORTED, "implicit leapfrog not compiled in");
}
int64_t mb200_implicit_workspace_bytes(int64_t, int32_t, const mb200_model*) { return 0; }
int mb200_hamiltonian_riemannian(const double*, const double*, int64_t, int32_t,
                                 const mb200_model*, double*, int32_t*, void*, int64_t, void*) {
  return fail(MB200_ERR_UNSUPPORTED, "riemannian hamiltonian not compiled in");
}
int mb200_selftest_fixed_point(int32_t, int32_t, const double*, const double*, int64_t, int32_t,
                               double, double, int32_t, double*, int32_t*, int32_t*, void*) {
  return fail(MB200_ERR_UNSUPPORTED, "fixed-point self-test not compiled in");
}
#endif

#ifndef MB200_NO_RIEMANNIAN
int mb200_implicit_leapfrog_riemannian(
    const double* pos_in, const double* mom_in, double* pos_out, double* mom_out,
    const int32_t* dir, int64_t n_chains, int32_t dim, double step_size, int32_t n_steps,
    const mb200_model* model, int32_t fp_solver, double fp_convergence_tol,
    double fp_divergence_tol, int32_t fp_max_iters, double reverse_check_tol, double* h_out,
    int32_t* status, int32_t* n_done, int32_t* fp_iters, void* workspace, int64_t workspace_bytes,
    void* stream) {
  (void)workspace;
  (void)workspace_bytes;
  if (n_chains == 0 && dim >= 1) return 0;
  if (!pos_in || !mom_in || !pos_out || !mom_out || !model)
    return fail(MB200_ERR_INVALID_ARG, "null pointer argument");
  if (n_chains < 0 || dim < 1 || n_steps < 0 || fp_max_iters < 0)
    return fail(MB200_ERR_INVALID_ARG, "bad sizes");
  if (n_chains == 0) return 0;
  return implicit_dispatch(pos_in, mom_in, pos_out, mom_out, dir, n_chains, dim, step_size,
                           n_steps, to_args(model), fp_convergence_tol, fp_divergence_tol,
                           fp_max_iters, reverse_check_tol, h_out, status, n_done, fp_iters,
                           (cudaStream_t)stream, 0, fp_solver);
}

// every per-chain buffer of the implicit kernels lives in shared memory for the supported sizes
int64_t mb200_implicit_workspace_bytes(int64_t, int32_t, const mb200_model*) { return 0; }

int mb200_hamiltonian_riemannian(const double* pos, const double* mom, int64_t n_chains,
                                 int32_t dim, const mb200_model* model, double* h_out,
                                 int32_t* status, void* workspace, int64_t workspace_bytes,
                                 void* stream) {
  (void)workspace;
  (void)workspace_bytes;
  if (n_chains == 0 && dim >= 1) return 0;
  if (!pos || !mom || !model || !h_out) return fail(MB200_ERR_INVALID_ARG, "null pointer argument");
  if (n_chains < 0 || dim < 1) return fail(MB200_ERR_INVALID_ARG, "bad sizes");
  if (n_chains == 0) return 0;
  // zero steps: state written back unchanged in place, h evaluated
  return implicit_dispatch(pos, mom, const_cast<double*>(pos), const_cast<double*>(mom), nullptr,
                           n_chains, dim, 0.0, 0, to_args(model), 1e-9, 1e10, 100, 2e-8, h_out,
                           status, nullptr, nullptr, (cudaStream_t)stream);
}
#endif

int mb200_metropolis_select(double* pos, double* mom, const double* pos_prop,
                            const double* mom_prop, const double* h_init, const double* h_prop,
                            const int32_t* status, const int32_t* n_done, int32_t* dir,
                            const double* uniforms, int64_t n_chains, int32_t dim,
                            double* accept_prob, double* accept_stat, int32_t* accepted,
                            void* stream) {
  if (n_chains == 0 && dim >= 1) return 0;
  if (!pos || !mom || !pos_prop || !mom_prop || !h_init || !h_prop || !uniforms)
    return fail(MB200_ERR_INVALID_ARG, "null pointer argument");
  if (n_chains < 0 || dim < 1) return fail(MB200_ERR_INVALID_ARG, "bad sizes");
  if (n_chains == 0) return 0;
  int64_t blocks = (n_chains * dim + 255) / 256;
  const int64_t cap = (int64_t)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  metropolis_select_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      pos, mom, pos_prop, mom_prop, h_init, h_prop, status, n_done, dir, uniforms, n_chains, dim,
      accept_prob, accept_stat, accepted);
  return check_launch("metropolis_select_kernel");
}

#ifndef MB200_NO_RIEMANNIAN
int mb200_selftest_fixed_point(int32_t func_id, int32_t fp_solver, const double* x0,
                               const double* y, int64_t n, int32_t dim, double convergence_tol,
                                      double divergence_tol, int32_t max_iters, double* x_out,
                                      int32_t* iters_out, int32_t* status, void* stream) {
  if (!x0 || !y || !x_out || !iters_out || !status)
    return fail(MB200_ERR_INVALID_ARG, "null pointer argument");
  if (n < 0 || dim < 1 || func_id < 0 || func_id > 4) return fail(MB200_ERR_INVALID_ARG, "bad arguments");
  if (n == 0) return 0;
  const size_t smem = (size_t)(3 * dim + 40) * sizeof(double);
  int64_t blocks = n < 4096 ? n : 4096;
  fixed_point_selftest_kernel<<<(unsigned)blocks, 64, smem, (cudaStream_t)stream>>>(
      func_id, fp_solver, x0, y, n, dim, convergence_tol, divergence_tol, max_iters, x_out, iters_out,
      status);
  return check_launch("fixed_point_selftest_kernel");
}
#endif

int mb200_composition_euclidean(const double* pos_in, const double* mom_in, double* pos_out,
                                double* mom_out, const int32_t* dir, int64_t n_chains,
                                int32_t dim, double step_size, int32_t n_steps, int32_t n_flows,
                                const double* coefficients, int32_t initial_h1_flow_step,
                                int32_t metric_kind, const double* metric_inv,
                                const mb200_model* model, double* h_out, int32_t* status,
                                int32_t* n_done, void* stream) {
  if (!coefficients || n_flows < 1 || n_flows > MB200_MAX_FLOWS || (n_flows & 1) == 0)
    return fail(MB200_ERR_INVALID_ARG, "n_flows must be odd and in [1, %d]", MB200_MAX_FLOWS);
  FlowSchedule s;
  memset(&s, 0, sizeof(s));
  s.n = n_flows;
  for (int i = 0; i < n_flows; ++i) {
    s.coef[i] = coefficients[i];
    const bool is_a = (i & 1) == 0;  // flows alternate a, b, a, ... (integrators.py:279-281)
    const bool drift = initial_h1_flow_step ? !is_a : is_a;
    if (drift) s.drift_mask |= 1u << i;
  }
  return leapfrog_euclidean_impl(pos_in, mom_in, pos_out, mom_out, dir, n_chains, dim, step_size,
                                 n_steps, metric_kind, metric_inv, model, h_out, status, n_done,
                                 (cudaStream_t)stream, false, &s);
}

#ifndef MB200_NO_RIEMANNIAN
int mb200_selftest_eigh(const double* matrices, int64_t n_matrices, int32_t dim, int32_t warm_from,
                        double* eigval, double* eigvec, int32_t* status, void* stream) {
  if (!matrices || !eigval || !eigvec || !status)
    return fail(MB200_ERR_INVALID_ARG, "null pointer argument");
  if (n_matrices < 0 || dim < 1 || warm_from >= n_matrices)
    return fail(MB200_ERR_INVALID_ARG, "bad arguments");
  if (n_matrices == 0) return 0;
  const size_t smem = rm_smem_doubles(dim, 3) * sizeof(double);
  if (smem > 227 * 1024) return fail(MB200_ERR_UNSUPPORTED, "dim %d too large for the self-test", dim);
  cudaError_t e = cudaFuncSetAttribute(eigh_selftest_kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return fail(MB200_ERR_CUDA, "smem attr: %s", cudaGetErrorString(e));
  int64_t blocks = n_matrices < 1024 ? n_matrices : 1024;
  eigh_selftest_kernel<<<(unsigned)blocks, RM_THREADS, smem, (cudaStream_t)stream>>>(
      matrices, n_matrices, dim, warm_from, eigval, eigvec, status);
  return check_launch("eigh_selftest_kernel");
}
#else
int mb200_selftest_eigh(const double*, int64_t, int32_t, int32_t, double*, double*, int32_t*,
                        void*) {
  return fail(MB200_ERR_UNSUPPORTED, "eigh self-test not compiled in");
}
#endif

#ifndef MB200_NO_RIEMANNIAN
int mb200_implicit_midpoint_riemannian(
    const double* pos_in, const double* mom_in, double* pos_out, double* mom_out,
    const int32_t* dir, int64_t n_chains, int32_t dim, double step_size, int32_t n_steps,
    const mb200_model* model, int32_t fp_solver, double fp_convergence_tol,
    double fp_divergence_tol, int32_t fp_max_iters, double reverse_check_tol, double* h_out,
    int32_t* status, int32_t* n_done, int32_t* fp_iters, void* stream) {
  if (n_chains == 0 && dim >= 1) return 0;
  if (!pos_in || !mom_in || !pos_out || !mom_out || !model)
    return fail(MB200_ERR_INVALID_ARG, "null pointer argument");
  if (n_chains < 0 || dim < 1 || n_steps < 0 || fp_max_iters < 0)
    return fail(MB200_ERR_INVALID_ARG, "bad sizes");
  return implicit_dispatch(pos_in, mom_in, pos_out, mom_out, dir, n_chains, dim, step_size,
                           n_steps, to_args(model), fp_convergence_tol, fp_divergence_tol,
                           fp_max_iters, reverse_check_tol, h_out, status, n_done, fp_iters,
                           (cudaStream_t)stream, 1, fp_solver);
}
#else
int mb200_implicit_midpoint_riemannian(const double*, const double*, double*, double*,
                                       const int32_t*, int64_t, int32_t, double, int32_t,
                                       const mb200_model*, int32_t, double, double, int32_t, double,
                                       double*, int32_t*, int32_t*, int32_t*, void*) {
  return fail(MB200_ERR_UNSUPPORTED, "implicit midpoint not compiled in");
}
#endif

int mb200_project_onto_cotangent_space(const double* pos, const double* mom_in, double* mom_out,
                                       int64_t n_chains, int32_t dim, int32_t metric_kind,
                                       const double* metric_inv, const mb200_model* model,
                                       void* stream) {
#ifdef MB200_NO_CONSTRAINED
  return fail(MB200_ERR_UNSUPPORTED, "constrained systems not compiled in");
#else
  if (n_chains == 0 && dim >= 1) return 0;
  if (!pos || !mom_in || !mom_out || !model) return fail(MB200_ERR_INVALID_ARG, "null pointer argument");
  if (n_chains < 0 || dim < 1) return fail(MB200_ERR_INVALID_ARG, "bad sizes");
  if (metric_kind < 0 || metric_kind > 2) return fail(MB200_ERR_INVALID_ARG, "bad metric_kind");
  if (metric_kind != MB200_METRIC_IDENTITY && !metric_inv)
    return fail(MB200_ERR_INVALID_ARG, "metric_inv is NULL");
  const ModelArgs m = to_args(model);
  cudaStream_t st = (cudaStream_t)stream;
#define MB200_ARGS pos, mom_in, mom_out, n_chains, dim, metric_kind, metric_inv, m, st
  switch (m.target_id) {
    case MB200_TARGET_TORUS:
      if (dim != 3) return fail(MB200_ERR_INVALID_ARG, "torus target needs dim == 3");
      return launch_project<TorusTarget, 1>(MB200_ARGS);
    case MB200_TARGET_SPHERE:
      if (dim <= 64) return launch_project<SphereTarget, 1>(MB200_ARGS);
      if (dim <= 128) return launch_project<SphereTarget, 2>(MB200_ARGS);
      if (dim <= 256) return launch_project<SphereTarget, 4>(MB200_ARGS);
      return fail(MB200_ERR_UNSUPPORTED, "sphere target: dim %d > 256 not supported", dim);
    default:
      return fail(MB200_ERR_UNSUPPORTED, "target %d defines no constraint", m.target_id);
  }
#undef MB200_ARGS
#endif
}

int mb200_sample_momentum_riemannian(const double* pos, const double* normals, double* mom_out,
                                     int64_t n_chains, int32_t dim, const mb200_model* model,
                                     int32_t* status, void* stream) {
#ifdef MB200_NO_RIEMANNIAN
  return fail(MB200_ERR_UNSUPPORTED, "riemannian systems not compiled in");
#else
  if (n_chains == 0 && dim >= 1) return 0;
  if (!pos || !normals || !mom_out || !model) return fail(MB200_ERR_INVALID_ARG, "null pointer argument");
  if (n_chains < 0 || dim < 1) return fail(MB200_ERR_INVALID_ARG, "bad sizes");
  const ModelArgs m = to_args(model);
  cudaStream_t st = (cudaStream_t)stream;
#define MB200_ARGS pos, normals, mom_out, n_chains, dim, m, status, st
  if (m.rmetric_id == MB200_RMETRIC_SOFTABS) {
    if (m.target_id == MB200_TARGET_BANANA) return launch_sample_momentum<BananaRTarget, SoftAbsMetric>(MB200_ARGS);
    return fail(MB200_ERR_UNSUPPORTED, "target %d has no device Hessian", m.target_id);
  }
  if (m.rmetric_id == MB200_RMETRIC_RANK1) {
    if (!m.maux) return fail(MB200_ERR_INVALID_ARG, "rank-1 metric needs its base matrix");
    if (rm_smem_doubles(dim, 1) * sizeof(double) > 227 * 1024)
      return fail(MB200_ERR_UNSUPPORTED,
                  "dim %d: the Cholesky factor of M(q) does not fit in shared memory", dim);
    switch (m.target_id) {
      case MB200_TARGET_QUADRATIC: return launch_sample_momentum<QuadraticRTarget, Rank1DenseMetric>(MB200_ARGS);
      case MB200_TARGET_STD_GAUSSIAN: return launch_sample_momentum<StdGaussianRTarget, Rank1DenseMetric>(MB200_ARGS);
      case MB200_TARGET_BANANA: return launch_sample_momentum<BananaRTarget, Rank1DenseMetric>(MB200_ARGS);
      default: return fail(MB200_ERR_UNSUPPORTED, "target %d not available", m.target_id);
    }
  }
#undef MB200_ARGS
  return fail(MB200_ERR_INVALID_ARG, "unknown rmetric_id %d", m.rmetric_id);
#endif
}

int mb200_leapfrog_euclidean_per_chain(const double* pos_in, const double* mom_in, double* pos_out,
                                       double* mom_out, const int32_t* dir, int64_t n_chains,
                                       int32_t dim, const double* step_sizes,
                                       const int32_t* n_steps_per_chain, int32_t max_n_steps,
                                       int32_t n_flows, const double* coefficients,
                                       int32_t initial_h1_flow_step, int32_t metric_kind,
                                       const double* metric_inv, const mb200_model* model,
                                       double* h_out, int32_t* status, int32_t* n_done,
                                       void* stream) {
  if (n_chains > 0 && !step_sizes) return fail(MB200_ERR_INVALID_ARG, "step_sizes is NULL");
  FlowSchedule s = leapfrog_schedule();
  if (coefficients != nullptr) {
    if (n_flows < 1 || n_flows > MB200_MAX_FLOWS || (n_flows & 1) == 0)
      return fail(MB200_ERR_INVALID_ARG, "n_flows must be odd and in [1, %d]", MB200_MAX_FLOWS);
    memset(&s, 0, sizeof(s));
    s.n = n_flows;
    for (int i = 0; i < n_flows; ++i) {
      s.coef[i] = coefficients[i];
      const bool is_a = (i & 1) == 0;
      if (initial_h1_flow_step ? !is_a : is_a) s.drift_mask |= 1u << i;
    }
  }
  s.step_sizes = step_sizes;
  s.n_steps = n_steps_per_chain;
  return leapfrog_euclidean_impl(pos_in, mom_in, pos_out, mom_out, dir, n_chains, dim, 0.0,
                                 max_n_steps, metric_kind, metric_inv, model, h_out, status,
                                 n_done, (cudaStream_t)stream, false, &s);
}

int mb200_leapfrog_gaussian_euclidean(const double* pos_in, const double* mom_in, double* pos_out,
                                      double* mom_out, const int32_t* dir, int64_t n_chains,
                                      int32_t dim, double step_size, const double* step_sizes,
                                      int32_t n_steps, int32_t n_flows, const double* coefficients,
                                      int32_t initial_h1_flow_step, int32_t metric_kind,
                                      const double* metric_inv, const double* rotation,
                                      const mb200_model* model, double* h_out, int32_t* status,
                                      int32_t* n_done, void* stream) {
  FlowSchedule s = leapfrog_schedule();
  if (coefficients != nullptr) {
    if (n_flows < 1 || n_flows > MB200_MAX_FLOWS || (n_flows & 1) == 0)
      return fail(MB200_ERR_INVALID_ARG, "n_flows must be odd and in [1, %d]", MB200_MAX_FLOWS);
    memset(&s, 0, sizeof(s));
    s.n = n_flows;
    for (int i = 0; i < n_flows; ++i) {
      s.coef[i] = coefficients[i];
      const bool is_a = (i & 1) == 0;
      if (initial_h1_flow_step ? !is_a : is_a) s.drift_mask |= 1u << i;
    }
  }
  if (metric_kind != MB200_METRIC_IDENTITY && !rotation && n_chains > 0)
    return fail(MB200_ERR_INVALID_ARG, "rotation is NULL");
  if (metric_kind == MB200_METRIC_DENSE && step_sizes)
    return fail(MB200_ERR_UNSUPPORTED,
                "per-chain step sizes need per-chain rotation matrices for a dense metric");
  s.gaussian = 1;
  s.rot = rotation;
  s.step_sizes = step_sizes;
  return leapfrog_euclidean_impl(pos_in, mom_in, pos_out, mom_out, dir, n_chains, dim, step_size,
                                 n_steps, metric_kind, metric_inv, model, h_out, status, n_done,
                                 (cudaStream_t)stream, false, &s);
}

int mb200_constrained_leapfrog_euclidean_per_chain(
    const double* pos_in, const double* mom_in, double* pos_out, double* mom_out,
    const int32_t* dir, int64_t n_chains, int32_t dim, const double* step_sizes,
    const int32_t* n_steps_per_chain, int32_t max_n_steps, int32_t n_inner_step,
    int32_t metric_kind, const double* metric_inv, const mb200_model* model,
    int32_t projection_solver, double constraint_tol, double position_tol, double divergence_tol,
    int32_t max_iters, int32_t max_line_search_iters, double reverse_check_tol, double* h_out,
    int32_t* status, int32_t* n_done, int32_t* newton_iters, void* stream) {
  if (n_chains > 0 && !step_sizes) return fail(MB200_ERR_INVALID_ARG, "step_sizes is NULL");
  PerChainScope scope(step_sizes, n_steps_per_chain);
  return mb200_constrained_leapfrog_euclidean(
      pos_in, mom_in, pos_out, mom_out, dir, n_chains, dim, 0.0, max_n_steps, n_inner_step,
      metric_kind, metric_inv, model, projection_solver, constraint_tol, position_tol,
      divergence_tol, max_iters, max_line_search_iters, reverse_check_tol, h_out, status, n_done,
      newton_iters, stream);
}

int mb200_implicit_riemannian_per_chain(
    const double* pos_in, const double* mom_in, double* pos_out, double* mom_out,
    const int32_t* dir, int64_t n_chains, int32_t dim, const double* step_sizes,
    const int32_t* n_steps_per_chain, int32_t max_n_steps, int32_t midpoint,
    const mb200_model* model, int32_t fp_solver, double fp_convergence_tol,
    double fp_divergence_tol, int32_t fp_max_iters, double reverse_check_tol, double* h_out,
    int32_t* status, int32_t* n_done, int32_t* fp_iters, void* stream) {
  if (n_chains > 0 && !step_sizes) return fail(MB200_ERR_INVALID_ARG, "step_sizes is NULL");
  PerChainScope scope(step_sizes, n_steps_per_chain);
  if (midpoint)
    return mb200_implicit_midpoint_riemannian(pos_in, mom_in, pos_out, mom_out, dir, n_chains, dim,
                                              0.0, max_n_steps, model, fp_solver,
                                              fp_convergence_tol, fp_divergence_tol, fp_max_iters,
                                              reverse_check_tol, h_out, status, n_done, fp_iters,
                                              stream);
  return mb200_implicit_leapfrog_riemannian(pos_in, mom_in, pos_out, mom_out, dir, n_chains, dim,
                                            0.0, max_n_steps, model, fp_solver, fp_convergence_tol,
                                            fp_divergence_tol, fp_max_iters, reverse_check_tol,
                                            h_out, status, n_done, fp_iters, nullptr, 0, stream);
}

int64_t mb200_nuts_workspace_bytes(int64_t n_chains, int32_t dim, int32_t max_tree_depth) {
  if (n_chains < 0 || dim < 1 || dim > 1024 || max_tree_depth < 1 ||
      max_tree_depth > NUTS_MAX_DEPTH)
    return -1;
  return (int64_t)(nuts_workspace_doubles_per_chain(dim, max_tree_depth) * sizeof(double)) *
         n_chains;
}

int mb200_nuts_euclidean(const double* pos_in, const double* mom_in, double* pos_out,
                         double* mom_out, int64_t n_chains, int32_t dim, double step_size,
                         const double* step_sizes, int32_t metric_kind, const double* metric_inv,
                         const mb200_model* model, int32_t slice_variant,
                         int32_t euclidean_criterion, int32_t extra_subtree_checks,
                         int32_t max_tree_depth, double max_delta_h, const double* uniforms,
                         int32_t n_uniforms, void* workspace, int64_t workspace_bytes,
                         double* h_out, int32_t* n_step, double* av_metrop_accept_prob,
                         double* reject_prob, int32_t* tree_depth, int32_t* diverging,
                         int32_t* n_uniforms_used, int32_t* dir_out, int32_t* status,
                         void* stream) {
  if (n_chains == 0 && dim >= 1) return 0;
  if (!pos_in || !mom_in || !pos_out || !mom_out || !model || !uniforms || !workspace)
    return fail(MB200_ERR_INVALID_ARG, "null pointer argument");
  if (n_chains < 0 || dim < 1 || n_uniforms < 1) return fail(MB200_ERR_INVALID_ARG, "bad sizes");
  if (max_tree_depth < 1 || max_tree_depth > NUTS_MAX_DEPTH)
    return fail(MB200_ERR_INVALID_ARG, "max_tree_depth must be in [1, %d]", NUTS_MAX_DEPTH);
  if (metric_kind < 0 || metric_kind > 2) return fail(MB200_ERR_INVALID_ARG, "bad metric_kind");
  if (metric_kind != MB200_METRIC_IDENTITY && !metric_inv)
    return fail(MB200_ERR_INVALID_ARG, "metric_inv is NULL");
  const int64_t need = mb200_nuts_workspace_bytes(n_chains, dim, max_tree_depth);
  if (need < 0) return fail(MB200_ERR_UNSUPPORTED, "dim %d > 1024 not supported", dim);
  if (workspace_bytes < need)
    return fail(MB200_ERR_INVALID_ARG, "workspace too small: %lld < %lld bytes",
                (long long)workspace_bytes, (long long)need);
  const ModelArgs m = to_args(model);
  if (m.target_id == MB200_TARGET_BANANA && (dim & 1))
    return fail(MB200_ERR_INVALID_ARG, "banana target needs even dim");
  NutsArgs a;
  a.max_depth = max_tree_depth;
  a.max_delta_h = max_delta_h;
  a.euclidean_criterion = euclidean_criterion;
  a.extra_checks = extra_subtree_checks;
  a.slice = slice_variant;
  a.uniforms = uniforms;
  a.n_uniforms = n_uniforms;
  a.step_sizes = step_sizes;
  a.stage_metric = 0;
  cudaStream_t st = (cudaStream_t)stream;
#define MB200_ARGS                                                                             \
  pos_in, mom_in, pos_out, mom_out, n_chains, dim, step_size, metric_kind, metric_inv, m, a,   \
      (double*)workspace, h_out, n_step, av_metrop_accept_prob, reject_prob, tree_depth,       \
      diverging, n_uniforms_used, dir_out, status, st
  switch (m.target_id) {
    case MB200_TARGET_STD_GAUSSIAN: return dispatch_nuts_dim<StdGaussianTarget>(MB200_ARGS);
    case MB200_TARGET_NEAL_FUNNEL: return dispatch_nuts_dim<NealFunnelTarget>(MB200_ARGS);
    case MB200_TARGET_BANANA: return dispatch_nuts_dim<BananaTarget>(MB200_ARGS);
    default:
      return fail(MB200_ERR_UNSUPPORTED, "target %d not available for Euclidean NUTS",
                  m.target_id);
  }
#undef MB200_ARGS
}

int64_t mb200_host_scratch_bytes(int64_t n_chains, int32_t dim) {
  if (n_chains < 0 || dim < 1) return -1;
  return n_chains * ((int64_t)4 * dim * (int64_t)sizeof(double) + 2 * (int64_t)sizeof(int32_t));
}

int mb200_leapfrog_euclidean_host(const double* pos_in, const double* mom_in, double* pos_out,
                                  double* mom_out, const int32_t* dir, int64_t n_chains,
                                  int32_t dim, double step_size, int32_t n_steps,
                                  int32_t metric_kind, const double* metric_inv,
                                  const mb200_model* model, int32_t* status, int32_t n_chunks,
                                  void* const* streams, int32_t n_streams, void* scratch,
                                  int64_t scratch_bytes, int32_t synchronize) {
  if (n_chains == 0 && dim >= 1) return 0;
  if (!pos_in || !mom_in || !pos_out || !mom_out || !model || !streams || !scratch)
    return fail(MB200_ERR_INVALID_ARG, "null pointer argument");
  if (n_chains < 0 || dim < 1 || n_steps < 0 || n_chunks < 1 || n_streams < 1)
    return fail(MB200_ERR_INVALID_ARG, "bad sizes");
  if (scratch_bytes < mb200_host_scratch_bytes(n_chains, dim))
    return fail(MB200_ERR_INVALID_ARG, "scratch too small");
  const size_t nd = (size_t)n_chains * dim;
  double* d_qi = (double*)scratch;
  double* d_pi = d_qi + nd;
  double* d_qo = d_pi + nd;
  double* d_po = d_qo + nd;
  int32_t* d_status = (int32_t*)(d_po + nd);
  int32_t* d_dir = d_status + n_chains;
  // chunk boundaries on the granularity of a CTA of the kernel that will run (56 chains for the
  // tensor-core kernel, 16 for the general one), so that the chunks together launch no more CTAs
  // than one launch over all chains would
  const int64_t align = (metric_kind == MB200_METRIC_DENSE && dim <= 128) ? 56 : 16;
  int64_t per = (n_chains + n_chunks - 1) / n_chunks;
  per = (per + align - 1) / align * align;
  int c = 0;
  for (int64_t lo = 0; lo < n_chains; lo += per, ++c) {
    const int64_t len = (lo + per <= n_chains) ? per : n_chains - lo;
    cudaStream_t st = (cudaStream_t)streams[c % n_streams];
    const size_t off = (size_t)lo * dim, bytes = (size_t)len * dim * sizeof(double);
    cudaMemcpyAsync(d_qi + off, pos_in + off, bytes, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(d_pi + off, mom_in + off, bytes, cudaMemcpyHostToDevice, st);
    if (dir) cudaMemcpyAsync(d_dir + lo, dir + lo, len * sizeof(int32_t), cudaMemcpyHostToDevice, st);
    const int rc = mb200_leapfrog_euclidean(d_qi + off, d_pi + off, d_qo + off, d_po + off,
                                            dir ? d_dir + lo : nullptr, len, dim, step_size,
                                            n_steps, metric_kind, metric_inv, model, nullptr,
                                            d_status + lo, nullptr, st);
    if (rc != 0) return rc;
    cudaMemcpyAsync(pos_out + off, d_qo + off, bytes, cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(mom_out + off, d_po + off, bytes, cudaMemcpyDeviceToHost, st);
    if (status)
      cudaMemcpyAsync(status + lo, d_status + lo, len * sizeof(int32_t), cudaMemcpyDeviceToHost, st);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(MB200_ERR_CUDA, "host path: %s", cudaGetErrorString(e));
  if (synchronize) {
    const int used = c < n_streams ? c : n_streams;
    for (int i = 0; i < used; ++i) {
      e = cudaStreamSynchronize((cudaStream_t)streams[i]);
      if (e != cudaSuccess) return fail(MB200_ERR_CUDA, "host path: %s", cudaGetErrorString(e));
    }
  }
  return 0;
}

}  // extern "C"
