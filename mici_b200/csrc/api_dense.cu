// libmici_b200.so -- C-ABI glue of the global-workspace dense metric policy (dense_global.cuh):
// implicit leapfrog / Hamiltonian / momentum refresh / velocity on Riemannian systems whose
// per-chain D x D metric does not fit in shared memory (config C4: D = 512), and the diagnostic
// entry point of the blocked factorisation.
#include "api_common.cuh"
#include "dense_global.cuh"

namespace mb200 {

static int dg_blocks(int64_t n) {
  const int64_t cap = (int64_t)num_sms();  // one CTA per SM (its shared memory is ~200 KB)
  return (int)(n < cap ? n : cap);
}

int64_t dense_global_workspace_bytes(int64_t n_chains, int dim) {
  return (int64_t)dg_blocks(n_chains) * (int64_t)dg_workspace_doubles(dim) * (int64_t)sizeof(double);
}

bool dense_global_supported(int dim) {
  return rm_smem_doubles(dim, RM_NMATS_GLOBAL) * sizeof(double) <= 227 * 1024;
}

template <class Target, template <class> class MetricT>
static int dg_launch_implicit(const double* q_in, const double* p_in, double* q_out, double* p_out,
                              const int32_t* dir, int64_t n, int dim, double eps, int n_steps,
                              ModelArgs m, double fp_tol, double fp_div, int fp_max, double rev_tol,
                              double* h_out, int32_t* status, int32_t* n_done, int32_t* fp_iters,
                              cudaStream_t st, int fp_solver, void* ws, int64_t ws_bytes) {
  auto kern = implicit_leapfrog_kernel<Target, MetricT>;
  const size_t smem = rm_smem_doubles(dim, RM_NMATS_GLOBAL) * sizeof(double);
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return fail(MB200_ERR_CUDA, "smem attr: %s", cudaGetErrorString(e));
  const int blocks = dg_blocks(n);
  DgScratch scratch(ws, ws_bytes, (size_t)dense_global_workspace_bytes(n, dim), st);
  if (scratch.ptr == nullptr) return fail(MB200_ERR_CUDA, "dense metric workspace allocation failed");
  m.workspace = scratch.ptr;
  m.ws_stride = dg_workspace_doubles(dim);
  kern<<<(unsigned)blocks, DG_THREADS, smem, st>>>(q_in, p_in, q_out, p_out, dir, n, dim, eps,
                                                   n_steps, m, fp_tol, fp_div, fp_max, rev_tol,
                                                   h_out, status, n_done, fp_iters,
                                                   RM_NMATS_GLOBAL, 0, fp_solver);
  return check_launch("implicit_leapfrog_kernel (global dense metric)");
}

int dense_global_implicit(const double* q_in, const double* p_in, double* q_out, double* p_out,
                          const int32_t* dir, int64_t n, int dim, double eps, int n_steps,
                          const ModelArgs& m, double fp_tol, double fp_div, int fp_max,
                          double rev_tol, double* h_out, int32_t* status, int32_t* n_done,
                          int32_t* fp_iters, cudaStream_t st, int midpoint, int fp_solver, void* ws,
                          int64_t ws_bytes) {
  if (midpoint)
    return fail(MB200_ERR_UNSUPPORTED,
                "implicit midpoint is not available for the global-workspace dense metric");
  if (!dense_global_supported(dim))
    return fail(MB200_ERR_UNSUPPORTED, "dim %d: panel buffers exceed shared memory", dim);
  if (!m.maux) return fail(MB200_ERR_INVALID_ARG, "dense metric needs its matrices (rmetric_aux)");
  if (m.target_id != MB200_TARGET_QUADRATIC)
    return fail(MB200_ERR_UNSUPPORTED,
                "target %d not compiled for the global-workspace dense metric", m.target_id);
  if (!m.taux) return fail(MB200_ERR_INVALID_ARG, "quadratic target needs its precision matrix");
#define MB200_ARGS                                                                         \
  q_in, p_in, q_out, p_out, dir, n, dim, eps, n_steps, m, fp_tol, fp_div, fp_max, rev_tol, \
      h_out, status, n_done, fp_iters, st, fp_solver, ws, ws_bytes
  if (m.rmetric_id == MB200_RMETRIC_HADAMARD)
    return dg_launch_implicit<QuadraticRTarget, GlobalDenseHadamard>(MB200_ARGS);
  return dg_launch_implicit<QuadraticRTarget, GlobalDenseRank1>(MB200_ARGS);
#undef MB200_ARGS
}

template <class Target, template <class> class MetricT, bool VELOCITY>
static int dg_launch_vec(const double* q, const double* v, double* out, int64_t n, int dim,
                         ModelArgs m, int32_t* status, cudaStream_t st) {
  const size_t smem = rm_smem_doubles(dim, RM_NMATS_GLOBAL) * sizeof(double);
  const int blocks = dg_blocks(n);
  DgScratch scratch(nullptr, 0, (size_t)dense_global_workspace_bytes(n, dim), st);
  if (scratch.ptr == nullptr) return fail(MB200_ERR_CUDA, "dense metric workspace allocation failed");
  m.workspace = scratch.ptr;
  m.ws_stride = dg_workspace_doubles(dim);
  cudaError_t e;
  if (VELOCITY) {
    auto kern = riemannian_velocity_kernel<Target, MetricT>;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(MB200_ERR_CUDA, "smem attr: %s", cudaGetErrorString(e));
    kern<<<(unsigned)blocks, DG_THREADS, smem, st>>>(q, v, out, n, dim, m, status, RM_NMATS_GLOBAL);
  } else {
    auto kern = riemannian_sample_momentum_kernel<Target, MetricT>;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(MB200_ERR_CUDA, "smem attr: %s", cudaGetErrorString(e));
    kern<<<(unsigned)blocks, DG_THREADS, smem, st>>>(q, v, out, n, dim, m, status, RM_NMATS_GLOBAL);
  }
  return check_launch("riemannian vector kernel (global dense metric)");
}

// velocity = 1: out = M(q)^-1 v ; velocity = 0: out = chol(M(q)) v
int dense_global_vector(const double* q, const double* v, double* out, int64_t n, int dim,
                        const ModelArgs& m, int32_t* status, cudaStream_t st, int velocity) {
  if (!dense_global_supported(dim))
    return fail(MB200_ERR_UNSUPPORTED, "dim %d: panel buffers exceed shared memory", dim);
  if (!m.maux) return fail(MB200_ERR_INVALID_ARG, "dense metric needs its matrices (rmetric_aux)");
  if (m.target_id != MB200_TARGET_QUADRATIC)
    return fail(MB200_ERR_UNSUPPORTED,
                "target %d not compiled for the global-workspace dense metric", m.target_id);
  const bool had = m.rmetric_id == MB200_RMETRIC_HADAMARD;
  if (velocity)
    return had ? dg_launch_vec<QuadraticRTarget, GlobalDenseHadamard, true>(q, v, out, n, dim, m, status, st)
               : dg_launch_vec<QuadraticRTarget, GlobalDenseRank1, true>(q, v, out, n, dim, m, status, st);
  return had ? dg_launch_vec<QuadraticRTarget, GlobalDenseHadamard, false>(q, v, out, n, dim, m, status, st)
             : dg_launch_vec<QuadraticRTarget, GlobalDenseRank1, false>(q, v, out, n, dim, m, status, st);
}

}  // namespace mb200

using namespace mb200;

extern "C" {

int mb200_selftest_dense_factor(const double* matrices, const double* rhs, int64_t n_matrices,
                                int32_t dim, double* chol_out, double* inv_out, double* sol_out,
                                double* logdet_out, int32_t* status, void* stream) {
  if (!matrices || !rhs || !chol_out || !inv_out || !sol_out || !logdet_out || !status)
    return fail(MB200_ERR_INVALID_ARG, "null pointer argument");
  if (n_matrices < 0 || dim < 1) return fail(MB200_ERR_INVALID_ARG, "bad sizes");
  if (n_matrices == 0) return 0;
  if (!dense_global_supported(dim))
    return fail(MB200_ERR_UNSUPPORTED, "dim %d: panel buffers exceed shared memory", dim);
  const DeviceScope device_scope(matrices);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t smem = rm_smem_doubles(dim, RM_NMATS_GLOBAL) * sizeof(double);
  cudaError_t e = cudaFuncSetAttribute(dense_global_selftest_kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return fail(MB200_ERR_CUDA, "smem attr: %s", cudaGetErrorString(e));
  const int blocks = dg_blocks(n_matrices);
  DgScratch scratch(nullptr, 0, (size_t)dense_global_workspace_bytes(n_matrices, dim), st);
  if (scratch.ptr == nullptr) return fail(MB200_ERR_CUDA, "workspace allocation failed");
  ModelArgs m;
  memset(&m, 0, sizeof(m));
  m.workspace = scratch.ptr;
  m.ws_stride = dg_workspace_doubles(dim);
  dense_global_selftest_kernel<<<(unsigned)blocks, DG_THREADS, smem, st>>>(
      matrices, rhs, n_matrices, dim, m, chol_out, inv_out, sol_out, logdet_out, status);
  return check_launch("dense_global_selftest_kernel");
}

}  // extern "C"
