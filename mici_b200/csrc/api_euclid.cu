// libmici_b200.so -- C-ABI entry points (include/mici_b200.h): Euclidean-metric leapfrog family (general-dimension kernel, evaluation pieces, host-buffer path).
// Host-side argument checking and kernel dispatch only; all arithmetic is in the .cuh kernels.
#include "api_common.cuh"
#include "leapfrog_generic.cuh"

#include <condition_variable>
#include <cstring>
#include <functional>
#include <mutex>
#include <queue>
#include <string>
#include <thread>
#include <vector>

namespace mb200 {

// ---- host-buffer path for PAGEABLE memory ----------------------------------------------------
// cudaMemcpyAsync on pageable memory is staged by the driver through one internal buffer and
// serialises with the caller (measured: 9.7 ms per 8192 x 128 launch against 0.9 ms from pinned
// memory).  The library stages such buffers itself: a few worker threads each take one row
// block end to end -- memcpy into a pinned bounce buffer, H2D + kernel + D2H on the block's
// stream, wait for its event, memcpy out of the bounce buffer -- so the host copies of one block
// overlap the DMA and the kernel of the others.
class HostStager {
 public:
  static HostStager& get() {
    static HostStager* s = new HostStager;  // never destroyed: its detached workers outlive main
    return *s;
  }
  // grow-only pinned bounce buffer of at least `bytes`
  void* bounce(size_t bytes) {
    if (bytes > cap_) {
      if (buf_) cudaFreeHost(buf_);
      buf_ = nullptr, cap_ = 0;
      if (cudaHostAlloc(&buf_, bytes, cudaHostAllocPortable) != cudaSuccess) return nullptr;
      cap_ = bytes;
    }
    return buf_;
  }
  // runs the tasks on the pool (the caller's thread takes part), returns when all are done
  void run(std::vector<std::function<void()>>& tasks) {
    const int want = (int)tasks.size() - 1 < MAX_WORKERS ? (int)tasks.size() - 1 : MAX_WORKERS;
    {
      std::lock_guard<std::mutex> lk(mu_);
      while (n_workers_ < want) {
        std::thread([this] { loop(); }).detach();
        ++n_workers_;
      }
      pending_ = (int)tasks.size();
      for (auto& t : tasks) queue_.push(&t);
    }
    cv_.notify_all();
    for (;;) {  // help out
      std::function<void()>* t = nullptr;
      {
        std::lock_guard<std::mutex> lk(mu_);
        if (!queue_.empty()) t = queue_.front(), queue_.pop();
      }
      if (!t) break;
      (*t)();
      finish_one();
    }
    std::unique_lock<std::mutex> lk(mu_);
    done_cv_.wait(lk, [this] { return pending_ == 0; });
  }
  std::mutex call_mu;  // one host-path call at a time owns the bounce buffer

 private:
  static constexpr int MAX_WORKERS = 7;
  void finish_one() {
    std::lock_guard<std::mutex> lk(mu_);
    if (--pending_ == 0) done_cv_.notify_all();
  }
  void loop() {
    for (;;) {
      std::function<void()>* t;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [this] { return !queue_.empty(); });
        t = queue_.front(), queue_.pop();
      }
      (*t)();
      finish_one();
    }
  }
  std::mutex mu_;
  std::condition_variable cv_, done_cv_;
  std::queue<std::function<void()>*> queue_;
  int n_workers_ = 0;
  int pending_ = 0;
  void* buf_ = nullptr;
  size_t cap_ = 0;
};

static bool is_pageable(const void* ptr) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, ptr) != cudaSuccess) {
    cudaGetLastError();
    return true;
  }
  return a.type == cudaMemoryTypeUnregistered;
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
  if (n == 0 && dim >= 1 && n_steps >= 0) return 0;  // empty batch: nothing to do
  if (!q_in || !p_in || !q_out || !p_out || !model)
    return fail(MB200_ERR_INVALID_ARG, "null pointer argument");
  if (n < 0 || dim < 1 || n_steps < 0) return fail(MB200_ERR_INVALID_ARG, "bad sizes");
  if (metric_kind < 0 || metric_kind > 2) return fail(MB200_ERR_INVALID_ARG, "bad metric_kind");
  if (metric_kind != MB200_METRIC_IDENTITY && !minv)
    return fail(MB200_ERR_INVALID_ARG, "metric_inv is NULL");
  if (n == 0) return 0;
  const DeviceScope device_scope(q_in);
  const ModelArgs m = to_args(model);
  if (m.target_id == MB200_TARGET_BANANA && (dim & 1))
    return fail(MB200_ERR_INVALID_ARG, "banana target needs even dim");
  if (allow_dmma && metric_kind == MB200_METRIC_DENSE && n_steps > 0) {
    int rc = leapfrog_dmma_dispatch(q_in, p_in, q_out, p_out, dir, sched.step_sizes, n, dim, eps,
                                    n_steps, minv, m, h_out, status, n_done, st);
    if (rc == 0) return check_launch("leapfrog_dmma_kernel");
    if (rc != MB200_ERR_UNSUPPORTED) return fail(rc, "leapfrog_dmma launch failed");
  }
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

}  // namespace mb200

using namespace mb200;

extern "C" {

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
  const DeviceScope device_scope(pos);
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
  // plain leapfrog with one trajectory length: the DMMA kernel takes the step sizes per chain
  const bool dmma_ok = coefficients == nullptr && n_steps_per_chain == nullptr;
  return leapfrog_euclidean_impl(pos_in, mom_in, pos_out, mom_out, dir, n_chains, dim, 0.0,
                                 max_n_steps, metric_kind, metric_inv, model, h_out, status,
                                 n_done, (cudaStream_t)stream, dmma_ok, &s);
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
  const DeviceScope device_scope(scratch);
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
  if (is_pageable(pos_in) || is_pageable(mom_in) || is_pageable(pos_out) || is_pageable(mom_out)) {
    // pageable buffers: staged through pinned bounce buffers by the worker pool; the call is
    // synchronous (as CUDA's own pageable copies are)
    HostStager& hs = HostStager::get();
    std::lock_guard<std::mutex> call_lock(hs.call_mu);
    const size_t vec_bytes = nd * sizeof(double);
    char* pin = (char*)hs.bounce(4 * vec_bytes + 2 * (size_t)n_chains * sizeof(int32_t));
    if (!pin) return fail(MB200_ERR_CUDA, "host path: cannot allocate the pinned bounce buffer");
    double* b_qi = (double*)pin;
    double* b_pi = b_qi + nd;
    double* b_qo = b_pi + nd;
    double* b_po = b_qo + nd;
    int32_t* b_status = (int32_t*)(b_po + nd);
    int32_t* b_dir = b_status + n_chains;
    int dev = 0;
    cudaGetDevice(&dev);
    std::vector<std::function<void()>> tasks;
    std::vector<int> rcs;
    std::vector<std::string> msgs;
    int n_tasks = 0;
    for (int64_t lo = 0; lo < n_chains; lo += per) ++n_tasks;
    rcs.assign(n_tasks, 0), msgs.resize(n_tasks);
    int c = 0;
    for (int64_t lo = 0; lo < n_chains; lo += per, ++c) {
      const int64_t len = (lo + per <= n_chains) ? per : n_chains - lo;
      cudaStream_t st = (cudaStream_t)streams[c % n_streams];
      const size_t off = (size_t)lo * dim, bytes = (size_t)len * dim * sizeof(double);
      tasks.emplace_back([=, &rcs, &msgs] {
        cudaSetDevice(dev);
        memcpy(b_qi + off, pos_in + off, bytes);
        memcpy(b_pi + off, mom_in + off, bytes);
        if (dir) memcpy(b_dir + lo, dir + lo, len * sizeof(int32_t));
        cudaMemcpyAsync(d_qi + off, b_qi + off, bytes, cudaMemcpyHostToDevice, st);
        cudaMemcpyAsync(d_pi + off, b_pi + off, bytes, cudaMemcpyHostToDevice, st);
        if (dir)
          cudaMemcpyAsync(d_dir + lo, b_dir + lo, len * sizeof(int32_t), cudaMemcpyHostToDevice, st);
        int rc = mb200_leapfrog_euclidean(d_qi + off, d_pi + off, d_qo + off, d_po + off,
                                          dir ? d_dir + lo : nullptr, len, dim, step_size, n_steps,
                                          metric_kind, metric_inv, model, nullptr, d_status + lo,
                                          nullptr, st);
        if (rc != 0) {
          rcs[c] = rc, msgs[c] = g_err;
          return;
        }
        cudaMemcpyAsync(b_qo + off, d_qo + off, bytes, cudaMemcpyDeviceToHost, st);
        cudaMemcpyAsync(b_po + off, d_po + off, bytes, cudaMemcpyDeviceToHost, st);
        cudaMemcpyAsync(b_status + lo, d_status + lo, len * sizeof(int32_t),
                        cudaMemcpyDeviceToHost, st);
        // NB several blocks may share a stream: wait for THIS block's work only
        cudaEvent_t ev;
        cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
        cudaEventRecord(ev, st);
        const cudaError_t e = cudaEventSynchronize(ev);
        cudaEventDestroy(ev);
        if (e != cudaSuccess) {
          rcs[c] = MB200_ERR_CUDA, msgs[c] = cudaGetErrorString(e);
          return;
        }
        memcpy(pos_out + off, b_qo + off, bytes);
        memcpy(mom_out + off, b_po + off, bytes);
        if (status) memcpy(status + lo, b_status + lo, len * sizeof(int32_t));
      });
    }
    hs.run(tasks);
    for (int i = 0; i < n_tasks; ++i)
      if (rcs[i] != 0) return fail(rcs[i], "host path (block %d): %s", i, msgs[i].c_str());
    return 0;
  }
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
