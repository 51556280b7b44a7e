"""TEST INFRASTRUCTURE ONLY -- CPU (NumPy/SciPy) restatement of Mici's ``Integrator.step``.

This module is the *oracle* for the CUDA hot path.  It restates, one chain at a time and
with the same NumPy/SciPy calls in the same order, the arithmetic that the reference
performs in

* ``src/mici/integrators.py``  (``LeapfrogIntegrator._step`` :170-173,
  ``ImplicitLeapfrogIntegrator`` :482-544, ``ConstrainedLeapfrogIntegrator`` :929-984)
* ``src/mici/systems.py``      (``h1_flow`` :143-152, Euclidean ``h2_flow`` :352-363,
  Riemannian derivatives :1375-1402, constrained-system methods :786-873, :1006-1031)
* ``src/mici/solvers.py``      (``solve_fixed_point_direct`` :47-94,
  ``solve_projection_onto_manifold_newton`` :346-469)
* ``src/mici/matrices.py``     (dense SPD Cholesky / explicit inverse :1161-1188, :897-912,
  :1060-1061, LU solve :1311, :1371-1384, SoftAbs :1631-1685)

without the reference's memoising ``ChainState`` cache and ``Matrix`` object model (the
cache only removes repeated evaluations; it does not change any value).

Pinning: the reference's own tests hold no golden vectors for ``Integrator.step``
(SURVEY.md section 8c).  The oracle is therefore pinned against *outputs of the reference
itself*: ``oracle/make_golden.py`` imports the unmodified reference from
``/root/reference/src``, steps seeded inputs through it, and commits the results under
``tests/golden/``; ``tests/test_oracle.py`` checks this module against those fixtures
(and against the live reference when it is importable).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may
import this module.  Nothing under ``mici_b200/`` does.
"""

from __future__ import annotations

import math

import numpy as np
import numpy.linalg as nla
import scipy.linalg as sla

# Per-chain outcome codes shared with the CUDA path (include/mici_b200.h)
STATUS_OK = 0
STATUS_CONVERGENCE = 1  # mici.errors.ConvergenceError
STATUS_NON_REVERSIBLE = 2  # mici.errors.NonReversibleStepError
STATUS_LINALG = 3  # mici.errors.LinAlgError surfaced as ConvergenceError by the solvers


class OracleIntegratorError(RuntimeError):
    """Stand-in for mici.errors.IntegratorError carrying the per-chain status code."""

    def __init__(self, status, msg=""):
        super().__init__(msg)
        self.status = status


class _LinAlgError(RuntimeError):
    """Stand-in for mici.errors.LinAlgError (errors.py:22)."""


def maximum_norm(vct):
    """solvers.py:25-27."""
    return (abs(vct)).max()


# --------------------------------------------------------------------------------------
# matrices.py arithmetic
# --------------------------------------------------------------------------------------


def _chkfinite(a):
    """ExplicitArrayMatrix.__init__ (matrices.py:207-215): non-finite -> LinAlgError."""
    if not np.all(np.isfinite(a)):
        raise _LinAlgError("Array is not finite.")
    return a


def dense_spd_factor(m):
    """DenseDefiniteMatrix.factor (matrices.py:1161-1173): lower Cholesky factor."""
    try:
        return nla.cholesky(m)
    except nla.LinAlgError as e:
        raise _LinAlgError("Cholesky factorisation failed.") from e


def dense_spd_inverse(m, chol=None):
    """Explicit dense inverse exactly as the reference builds it.

    ``DensePositiveDefiniteMatrix._construct_inv`` (matrices.py:1209-1210) ->
    ``DenseDefiniteMatrix._construct_inv`` (:1183-1188) ->
    ``_BaseTriangularFactoredDefiniteMatrix._construct_inv`` (:979-980) builds
    ``TriangularFactoredDefiniteMatrix(factor=L.inv.T)`` whose ``.array`` (:1060-1061) is
    ``factor @ factor.array.T`` with ``factor = InverseTriangularMatrix(L.T, lower=False)``:

    * ``factor.array = solve_triangular(L.T, I, lower=False)``        (:906-912)
    * ``factor @ X    = solve_triangular(L.T, X, lower=False)``        (:897-903)
    """
    _chkfinite(m)
    if chol is None:
        chol = dense_spd_factor(m)
    lt = chol.T
    inv_lt = sla.solve_triangular(lt, np.identity(m.shape[0]), lower=False, check_finite=False)
    inv = sla.solve_triangular(lt, inv_lt.T, lower=False, check_finite=False)
    return _chkfinite(inv)


class IdentityMetric:
    """matrices.IdentityMatrix (matrices.py:491-554)."""

    kind = "identity"

    def inv_matvec(self, v):
        return v

    def sqrt_matvec(self, v):
        return v


class DiagonalMetric:
    """matrices.PositiveDiagonalMatrix (matrices.py:771-792, 709-768)."""

    kind = "diagonal"

    def __init__(self, diagonal):
        self.diagonal = np.asarray(diagonal, dtype=np.float64)
        self.inv_diagonal = 1.0 / self.diagonal

    def inv_matvec(self, v):
        return self.inv_diagonal * v

    def sqrt_matvec(self, v):
        return self.diagonal**0.5 * v


class DenseMetric:
    """matrices.DensePositiveDefiniteMatrix used as a fixed metric (systems.py:339-340)."""

    kind = "dense"

    def __init__(self, array):
        self.array = np.asarray(array, dtype=np.float64)
        self.chol = dense_spd_factor(self.array)
        self.inv_array = dense_spd_inverse(self.array, self.chol)

    def inv_matvec(self, v):
        # ExplicitArrayMatrix._left_matrix_multiply (matrices.py:222-223)
        return self.inv_array @ v

    def sqrt_matvec(self, v):
        return self.chol @ v


def coerce_metric(metric):
    """EuclideanMetricSystem.__init__ metric coercion (systems.py:332-346)."""
    if metric is None:
        return IdentityMetric()
    if hasattr(metric, "inv_matvec") and hasattr(metric, "sqrt_matvec"):
        return metric  # already a metric object ("else pass-through", systems.py:345-346)
    metric = np.asarray(metric)
    if metric.ndim == 1:
        return DiagonalMetric(metric)
    if metric.ndim == 2:
        return DenseMetric(metric)
    raise ValueError("metric must be None, 1D or 2D")


# --------------------------------------------------------------------------------------
# Explicit leapfrog on a Euclidean-metric system
# --------------------------------------------------------------------------------------


def euclidean_h(q, p, target, metric):
    """System.h = h1 + h2 (systems.py:187-196); h2 = 0.5 p . M^-1 p (:348-350)."""
    return target.neg_log_dens(q) + 0.5 * (p @ metric.inv_matvec(p))


def metric_eig(metric):
    """``(eigval, eigvec or None)`` as the reference's matrix classes expose them: ones / the
    diagonal with identity eigenvectors (matrices.py:519-528, 743-749), ``numpy.linalg.eigh`` of
    the array for a dense matrix (matrices.py:436-438)."""
    if metric.kind == "identity":
        return 1.0, None
    if metric.kind == "diagonal":
        return metric.diagonal, None
    if not hasattr(metric, "_eig"):
        metric._eig = nla.eigh(metric.array)
    return metric._eig


def gaussian_h2_flow(q, p, dt, metric):
    """GaussianEuclideanMetricSystem.h2_flow (systems.py:464-474)."""
    eigval, u = metric_eig(metric)
    omega = 1.0 / eigval**0.5
    sin_omega_dt, cos_omega_dt = np.sin(omega * dt), np.cos(omega * dt)
    ut_q = q if u is None else u.T @ q
    ut_p = p if u is None else u.T @ p
    q_new = cos_omega_dt * ut_q + (sin_omega_dt * omega) * ut_p
    p_new = cos_omega_dt * ut_p - (sin_omega_dt / omega) * ut_q
    if u is not None:
        q_new, p_new = u @ q_new, u @ p_new
    return q_new, p_new


def gaussian_euclidean_h(q, p, target, metric):
    """h1 + h2 with h2 = q.q/2 + p.M^-1 p/2 (systems.py:450-453)."""
    return target.neg_log_dens(q) + (0.5 * q @ q + 0.5 * p @ metric.inv_matvec(p))


def gaussian_composition_steps(q, p, time_step, n_steps, target, metric, coefficients=(0.5, 1.0, 0.5),
                               initial_h1_flow_step=True):
    """Leapfrog (default coefficients; integrators.py:170-173) or a symmetric composition
    (integrators.py:283-289) over the flows of a ``GaussianEuclideanMetricSystem``."""
    q = np.array(q, dtype=np.float64)
    p = np.array(p, dtype=np.float64)
    metric = coerce_metric(metric)
    grad = target.grad_neg_log_dens(q)
    for _ in range(n_steps):
        for i, c in enumerate(coefficients):
            if ((i % 2) == 0) == bool(initial_h1_flow_step):
                p -= (c * time_step) * grad
            else:
                q, p = gaussian_h2_flow(q, p, c * time_step, metric)
                grad = target.grad_neg_log_dens(q)
    return q, p


def leapfrog_steps(q, p, time_step, n_steps, target, metric):
    """``n_steps`` calls of ``LeapfrogIntegrator.step`` (integrators.py:63-80, 170-173).

    ``time_step = state.dir * step_size``.  The reference evaluates the gradient once per
    step because ``grad_neg_log_dens`` is memoised on ``pos`` (systems.py:109-119,
    states.py:248-258): the trailing half-step's gradient is reused by the next leading
    half-step.  The two half-step momentum updates are kept separate (two roundings).
    """
    q = np.array(q, dtype=np.float64)
    p = np.array(p, dtype=np.float64)
    metric = coerce_metric(metric)
    grad = target.grad_neg_log_dens(q)
    for _ in range(n_steps):
        p -= (0.5 * time_step) * grad  # h1_flow  systems.py:143-152
        q += time_step * metric.inv_matvec(p)  # h2_flow  systems.py:362-363
        grad = target.grad_neg_log_dens(q)
        p -= (0.5 * time_step) * grad
    return q, p


def composition_coefficients(free_coefficients):
    """SymmetricCompositionIntegrator.__init__ (integrators.py:266-277)."""
    n_free = len(free_coefficients)
    coefficients = list(free_coefficients)
    coefficients.append(0.5 - sum(free_coefficients[(n_free) % 2 :: 2]))
    coefficients.append(1 - 2 * sum(free_coefficients[(n_free + 1) % 2 :: 2]))
    return coefficients + coefficients[-2::-1]


BCSS_FREE_COEFFICIENTS = {  # integrators.py:312-314, 337-345, 367-377
    "bcss2": ((3 - 3**0.5) / 6,),
    "bcss3": (0.11888010966548, 0.29619504261126),
    "bcss4": (0.071353913450279725904, 0.191667800000000000000, 0.268548791161230105820),
}


def composition_steps(q, p, time_step, n_steps, target, metric, coefficients,
                      initial_h1_flow_step=True):
    """``n_steps`` of ``SymmetricCompositionIntegrator._step`` (integrators.py:283-289): flows
    alternate a, b, ..., a; a = h1_flow if ``initial_h1_flow_step`` else h2_flow; the gradient is
    re-evaluated whenever ``pos`` changed (cache on ``pos``)."""
    q = np.array(q, dtype=np.float64)
    p = np.array(p, dtype=np.float64)
    metric = coerce_metric(metric)
    grad = target.grad_neg_log_dens(q)
    for _ in range(n_steps):
        for i, c in enumerate(coefficients):
            is_h1 = ((i % 2) == 0) == bool(initial_h1_flow_step)
            if is_h1:
                p -= (c * time_step) * grad
            else:
                q += (c * time_step) * metric.inv_matvec(p)
                grad = target.grad_neg_log_dens(q)
    return q, p


# --------------------------------------------------------------------------------------
# solve_fixed_point_direct
# --------------------------------------------------------------------------------------


def solve_fixed_point_direct(func, x0, convergence_tol=1e-9, divergence_tol=1e10, max_iters=100):
    """solvers.py:47-94 with ``norm=maximum_norm``.  Returns ``(x, n_iters)``."""
    error = np.nan
    try:
        for i in range(max_iters):
            x = func(x0)
            error = maximum_norm(x - x0)
            if error > divergence_tol or np.isnan(error):
                raise OracleIntegratorError(STATUS_CONVERGENCE, f"diverged at iteration {i}")
            if error < convergence_tol:
                return x, i + 1
            x0 = x
    except (ValueError, _LinAlgError) as e:
        raise OracleIntegratorError(STATUS_CONVERGENCE, f"{type(e)} in fixed point solver") from e
    raise OracleIntegratorError(STATUS_CONVERGENCE, f"did not converge, last error {error}")


def solve_fixed_point_steffensen(func, x0, convergence_tol=1e-9, divergence_tol=1e10, max_iters=100):
    """solvers.py:97-154 with ``norm=maximum_norm``.  Returns ``(x, n_iters)``."""
    error = np.nan
    try:
        for i in range(max_iters):
            x1 = func(x0)
            x2 = func(x1)
            denom = x2 - 2 * x1 + x0
            denom[abs(denom) == 0.0] = np.finfo(x0.dtype).eps
            x = x0 - (x1 - x0) ** 2 / denom
            error = maximum_norm(x - x0)
            if error > divergence_tol or np.isnan(error):
                raise OracleIntegratorError(STATUS_CONVERGENCE, f"diverged at iteration {i}")
            if error < convergence_tol:
                return x, i + 1
            x0 = x
    except (ValueError, _LinAlgError) as e:
        raise OracleIntegratorError(STATUS_CONVERGENCE, f"{type(e)} in fixed point solver") from e
    raise OracleIntegratorError(STATUS_CONVERGENCE, f"did not converge, last error {error}")


FIXED_POINT_SOLVERS = {"direct": solve_fixed_point_direct, "steffensen": solve_fixed_point_steffensen}


# --------------------------------------------------------------------------------------
# Riemannian metrics (position dependent)
# --------------------------------------------------------------------------------------


class DenseRiemannianMetricValue:
    """``DensePositiveDefiniteMatrix(metric_func(q))`` (systems.py:1360-1373, 1727-1734)."""

    def __init__(self, array):
        self.array = _chkfinite(np.asarray(array))
        self.chol = dense_spd_factor(self.array)
        self.inv_array = dense_spd_inverse(self.array, self.chol)

    @property
    def log_abs_det(self):
        # _BaseTriangularFactoredDefiniteMatrix.log_abs_det (matrices.py:982-984)
        return 2 * np.log(np.abs(self.chol.diagonal())).sum()

    def inv_matvec(self, v):
        return self.inv_array @ v

    def sqrt_matvec(self, v):
        return self.chol @ v

    @property
    def grad_log_abs_det(self):
        # matrices.py:1175-1177
        return self.inv_array

    def grad_quadratic_form_inv(self, v):
        # matrices.py:1179-1181
        w = self.inv_array @ v
        return -np.outer(w, w)


class SoftAbsMetricValue:
    """``SoftAbsRegularizedPositiveDefiniteMatrix`` (matrices.py:1631-1685)."""

    def __init__(self, symmetric_array, softabs_coeff):
        self.coeff = softabs_coeff
        try:
            self.unreg_eigval, self.eigvec = nla.eigh(symmetric_array)
        except nla.LinAlgError as e:  # surfaces as ValueError-like failure inside solver
            raise _LinAlgError("eigh failed") from e
        self.eigval = self.softabs(self.unreg_eigval)
        if not np.all(self.eigval > 0):
            # EigendecomposedPositiveDefiniteMatrix.__init__ (matrices.py:1606-1609); NaN
            # eigenvalues land here too.  ValueError -> ConvergenceError inside solvers.
            raise ValueError("Eigenvalues must all be positive.")

    def softabs(self, x):
        return x / np.tanh(x * self.coeff)  # :1662-1664

    def grad_softabs(self, x):
        return 1.0 / np.tanh(self.coeff * x) - self.coeff * x / np.sinh(self.coeff * x) ** 2

    @property
    def log_abs_det(self):
        # SymmetricMatrix.log_abs_det (matrices.py:456-459)
        return np.log(np.abs(self.eigval)).sum()

    def inv_matvec(self, v):
        # EigendecomposedSymmetricMatrix._left_matrix_multiply (:1555-1556) with 1/eigval
        return self.eigvec @ ((1 / self.eigval) * (self.eigvec.T @ v))

    def sqrt_matvec(self, v):
        return self.eigvec @ (self.eigval**0.5 * (self.eigvec.T @ v))

    @property
    def grad_log_abs_det(self):
        # :1673-1676 -> EigendecomposedSymmetricMatrix._construct_array (:1565-1572)
        grad_eigval = self.grad_softabs(self.unreg_eigval) / self.eigval
        d = self.eigvec.shape[0]
        return self.eigvec @ (grad_eigval[:, None] * (self.eigvec.T @ np.identity(d)))

    def grad_quadratic_form_inv(self, vector):
        # :1678-1685
        num_j_mtx = self.eigval[:, None] - self.eigval[None, :]
        num_j_mtx += np.diag(self.grad_softabs(self.unreg_eigval))
        den_j_mtx = self.unreg_eigval[:, None] - self.unreg_eigval[None, :]
        np.fill_diagonal(den_j_mtx, 1)
        j_mtx = num_j_mtx / den_j_mtx
        e_vct = self.eigvec.T @ vector / self.eigval
        return -(self.eigvec @ (np.outer(e_vct, e_vct) * j_mtx) @ self.eigvec.T)


class RiemannianSystem:
    """``RiemannianMetricSystem`` derivative plumbing (systems.py:1360-1402).

    ``kind='dense'``  : DenseRiemannianMetricSystem  -- ``metric_model`` supplies
                        ``metric_func`` / ``vjp_metric_func``.
    ``kind='softabs'``: SoftAbsRiemannianMetricSystem -- metric is SoftAbs of the target
                        Hessian, VJP is the target's matrix-Tressian product (:1846-1920).
    """

    def __init__(self, target, kind, metric_model=None, softabs_coeff=1.0):
        self.target = target
        self.kind = kind
        self.metric_model = metric_model
        self.softabs_coeff = softabs_coeff
        self.n_metric_evals = 0

    def metric(self, q):
        self.n_metric_evals += 1
        if self.kind == "dense":
            return DenseRiemannianMetricValue(self.metric_model.metric_func(q))
        return SoftAbsMetricValue(self.target.hess_neg_log_dens(q), self.softabs_coeff)

    def vjp(self, q):
        if self.kind == "dense":
            return self.metric_model.vjp_metric_func(q)
        return self.target.mtp_neg_log_dens(q)

    def h(self, q, p):
        m = self.metric(q)
        # h1 (:1378-1379) + h2 (:1387-1388)
        return (self.target.neg_log_dens(q) + 0.5 * m.log_abs_det) + 0.5 * (p @ m.inv_matvec(p))

    def dh1_dpos(self, q, m=None):
        m = self.metric(q) if m is None else m
        return self.target.grad_neg_log_dens(q) + 0.5 * self.vjp(q)(m.grad_log_abs_det)

    def dh2_dpos(self, q, p, m=None):
        m = self.metric(q) if m is None else m
        return 0.5 * self.vjp(q)(m.grad_quadratic_form_inv(p))

    def dh2_dmom(self, q, p, m=None):
        m = self.metric(q) if m is None else m
        return m.inv_matvec(p)


def implicit_leapfrog_step(
    q,
    p,
    time_step,
    system,
    reverse_check_tol=2e-8,
    fixed_point_solver_kwargs=None,
    counts=None,
    fixed_point_solver="direct",
):
    """One ``ImplicitLeapfrogIntegrator.step`` (integrators.py:482-544).

    NB: ``_step`` passes ``time_step`` *unchanged* to all six sub-maps
    (:538-544), so one call advances time by ``2*time_step``; this is reproduced as is.
    Raises OracleIntegratorError with the status code on failure.  ``counts`` (dict) collects
    the fixed-point iteration counts of the four solves.
    """
    kw = {} if fixed_point_solver_kwargs is None else fixed_point_solver_kwargs
    q = np.array(q, dtype=np.float64)
    p = np.array(p, dtype=np.float64)
    dt = time_step
    its = []

    def solve(func, x0):
        x, n = FIXED_POINT_SOLVERS[fixed_point_solver](func, x0, **kw)
        its.append(n)
        return x

    def step_b_fwd(q, p, dt):
        # :496-502 -- fixed point in p at fixed q (metric cached on pos: one build)
        m = system.metric(q)
        p_init = p
        return solve(lambda mom: p_init - dt * system.dh2_dpos(q, mom, m), p_init)

    def step_c_adj(q, p, dt):
        # :530-536 -- fixed point in q, new metric every iteration
        q_init = q
        return solve(lambda pos: q_init + dt * system.dh2_dmom(pos, p), q_init)

    try:
        # _step_a :493-494
        p = p - dt * system.dh1_dpos(q)
        # _step_b_fwd
        p = step_b_fwd(q, p, dt)
        # _step_c_fwd :517-528
        q_init = q.copy()
        q = q + dt * system.dh2_dmom(q, p)
        q_back = step_c_adj(q.copy(), p, -dt)
        rev_diff = maximum_norm(q_back - q_init)
        if rev_diff > reverse_check_tol:
            raise OracleIntegratorError(STATUS_NON_REVERSIBLE, f"pos rev diff {rev_diff}")
        # _step_c_adj
        q = step_c_adj(q, p, dt)
        # _step_b_adj :504-515
        p_init = p.copy()
        p = p - dt * system.dh2_dpos(q, p)
        p_back = step_b_fwd(q, p.copy(), -dt)
        rev_diff = maximum_norm(p_back - p_init)
        if rev_diff > reverse_check_tol:
            raise OracleIntegratorError(STATUS_NON_REVERSIBLE, f"mom rev diff {rev_diff}")
        # _step_a
        p = p - dt * system.dh1_dpos(q)
    except (ValueError, _LinAlgError) as e:
        # Outside the solvers the reference would propagate these as ValueError/LinAlgError
        # (not IntegratorError); they only occur for non-finite states.
        raise OracleIntegratorError(STATUS_LINALG, str(e)) from e
    finally:
        if counts is not None:
            counts["fp_iters"] = its
    return q, p


def implicit_midpoint_step(q, p, time_step, system, reverse_check_tol=2e-8,
                           fixed_point_solver_kwargs=None, counts=None,
                           fixed_point_solver="direct"):
    """One ``ImplicitMidpointIntegrator.step`` (integrators.py:609-681) on a Riemannian system:
    ``_step_a_fwd(dt/2)`` (fixed point in the concatenated (pos, mom)), then ``_step_a_adj(dt/2)``
    (explicit Euler from the previous state + reversibility check)."""
    kw = {} if fixed_point_solver_kwargs is None else fixed_point_solver_kwargs
    q = np.array(q, dtype=np.float64)
    p = np.array(p, dtype=np.float64)
    its = []

    def dh_dmom(q, p, m):
        return system.dh2_dmom(q, p, m)

    def dh_dpos(q, p, m):  # System.dh_dpos: dh1_dpos + dh2_dpos (systems.py:198-201)
        return system.dh1_dpos(q, m) + system.dh2_dpos(q, p, m)

    def step_a_fwd(q, p, dt):
        z_init = np.concatenate([q, p])

        def func(z):
            zq, zp = np.split(z, 2)
            m = system.metric(zq)
            return z_init + np.concatenate([dt * dh_dmom(zq, zp, m), -dt * dh_dpos(zq, zp, m)])

        z, n = FIXED_POINT_SOLVERS[fixed_point_solver](func, z_init, **kw)
        its.append(n)
        return np.split(z, 2)

    try:
        dt = time_step / 2
        q, p = step_a_fwd(q, p, dt)
        m = system.metric(q)
        q_prev, p_prev = q.copy(), p.copy()
        q = q + dt * dh_dmom(q_prev, p_prev, m)
        p = p - dt * dh_dpos(q_prev, p_prev, m)
        q_back, p_back = step_a_fwd(q.copy(), p.copy(), -dt)
        rev_diff = maximum_norm(np.concatenate([q_back - q_prev, p_back - p_prev]))
        if rev_diff > reverse_check_tol:
            raise OracleIntegratorError(STATUS_NON_REVERSIBLE, f"rev diff {rev_diff}")
    except (ValueError, _LinAlgError) as e:
        raise OracleIntegratorError(STATUS_LINALG, str(e)) from e
    finally:
        if counts is not None:
            counts["fp_iters"] = its
    return q, p


# --------------------------------------------------------------------------------------
# Constrained leapfrog (RATTLE / geodesic integrator) on a Euclidean-metric system
# --------------------------------------------------------------------------------------


class ConstrainedSystem:
    """``DenseConstrainedEuclideanMetricSystem`` (systems.py:786-873, 1006-1031).  With
    ``dens_wrt_hausdorff=False`` the target density is given with respect to the Lebesgue measure
    of the ambient space and ``h1`` carries the correction ``log det gram / 2`` (:853-861)."""

    def __init__(self, target, metric=None, dens_wrt_hausdorff=True):
        self.target = target
        self.metric = coerce_metric(metric)
        self.dens_wrt_hausdorff = dens_wrt_hausdorff
        self.n_constr_evals = 0

    def gram(self, q):
        jac = self.jacob_constr(q)
        return jac @ self.inv_metric_mat(jac.T)  # systems.py:1013-1016

    def h1(self, q):
        if self.dens_wrt_hausdorff:
            return self.target.neg_log_dens(q)
        # log_det_sqrt_gram = 0.5 * DensePositiveDefiniteMatrix.log_abs_det (:836-838, 982-984)
        chol = np.linalg.cholesky(self.gram(q))
        return self.target.neg_log_dens(q) + 0.5 * (2 * np.log(np.abs(chol.diagonal())).sum())

    def dh1_dpos(self, q):
        if self.dens_wrt_hausdorff:
            return self.target.grad_neg_log_dens(q)
        # grad_log_det_sqrt_gram (:1024-1031): mhp_constr(inv_gram @ jac @ metric.inv)
        jac = self.jacob_constr(q)
        inv_gram = dense_spd_inverse(jac @ self.inv_metric_mat(jac.T))
        m = (inv_gram @ jac) @ self.inv_metric_full()
        return self.target.grad_neg_log_dens(q) + self.target.mhp_constr(q)(m)

    def inv_metric_full(self):
        """``metric.inv`` as the right operand of ``matrix @ metric.inv``."""
        m = self.metric
        if m.kind == "identity":
            return np.identity(self.jacob_constr_dim)
        if m.kind == "diagonal":
            return np.diag(m.inv_diagonal)
        return m.inv_array

    def constr(self, q):
        self.n_constr_evals += 1
        return self.target.constr(q)

    def jacob_constr(self, q):
        return self.target.jacob_constr(q)

    def inv_metric_mat(self, a):
        """``metric.inv @ a`` for a [D] vector or [D, K] matrix."""
        m = self.metric
        if m.kind == "identity":
            return a
        if m.kind == "diagonal":
            return m.inv_diagonal[:, None] * a if a.ndim == 2 else m.inv_diagonal * a
        return m.inv_array @ a

    def h(self, q, p):
        return self.h1(q) + 0.5 * (p @ self.inv_metric_mat(p))

    @property
    def jacob_constr_dim(self):
        return self.target.dim

    def project_onto_cotangent_space(self, mom, q):
        """systems.py:863-873: p -= J^T (gram^-1 (J (M^-1 p))), gram = J M^-1 J^T as a
        DensePositiveDefiniteMatrix whose ``.inv @ v`` is the explicit-inverse matvec."""
        jac = self.jacob_constr(q)
        gram = jac @ self.inv_metric_mat(jac.T)  # systems.py:1013-1016
        inv_gram = dense_spd_inverse(gram)
        mom = mom - jac.T @ (inv_gram @ (jac @ self.inv_metric_mat(mom)))
        return mom


def solve_projection_onto_manifold_newton(
    q,
    p,
    q_prev,
    time_step,
    system,
    constraint_tol=1e-9,
    position_tol=1e-8,
    divergence_tol=1e10,
    max_iters=50,
    counts=None,
):
    """solvers.py:346-469 for a Euclidean metric: ``dh2_flow_dmom = (|dt| M^-1, I)``
    (systems.py:794-799); residual Jacobian ``J (|dt| M^-1) J_prev^T`` is a
    ``DenseSquareMatrix`` solved by pivoted LU (systems.py:1020-1022, matrices.py:1311,
    1371-1384).  Returns ``(q, p)``; raises OracleIntegratorError(STATUS_CONVERGENCE)."""
    q = q.copy()
    p = p.copy()
    mu = np.zeros_like(q)
    jac_prev = system.jacob_constr(q_prev)
    adt = abs(time_step)
    error = np.nan
    try:
        for i in range(max_iters):
            jac = system.jacob_constr(q)
            c = system.constr(q)
            error = maximum_norm(c)
            # dt * metric.inv is a scaled matrix object in the reference:
            #  identity -> PositiveScaledIdentityMatrix (scalar*J^T),
            #  diagonal -> PositiveDiagonalMatrix(dt * 1/diag),
            #  dense    -> DensePositiveDefiniteMatrix(dt * inv_array)
            res_jac = _chkfinite(jac @ _scaled_inv_metric(system, adt, jac_prev.T))
            lu_piv = sla.lu_factor(res_jac, check_finite=False)
            delta_mu = jac_prev.T @ sla.lu_solve(lu_piv, c, 0, check_finite=False)
            delta_pos = _scaled_inv_metric(system, adt, delta_mu)
            if error > divergence_tol or np.isnan(error):
                raise OracleIntegratorError(STATUS_CONVERGENCE, f"Newton diverged at {i}")
            if error < constraint_tol and maximum_norm(delta_pos) < position_tol:
                p -= np.sign(time_step) * mu
                if counts is not None:
                    counts.setdefault("newton_iters", []).append(i + 1)
                return q, p
            mu += delta_mu
            q -= delta_pos
    except (ValueError, _LinAlgError) as e:
        raise OracleIntegratorError(STATUS_CONVERGENCE, f"{type(e)} in Newton solver") from e
    raise OracleIntegratorError(STATUS_CONVERGENCE, f"Newton did not converge, |c|={error}")


def solve_projection_onto_manifold_quasi_newton(
    q, p, q_prev, time_step, system, constraint_tol=1e-9, position_tol=1e-8,
    divergence_tol=1e10, max_iters=50, counts=None,
):
    """solvers.py:195-343 for a Euclidean metric: the Gram matrix J_prev (|dt| M^-1) J_prev^T is a
    DensePositiveDefiniteMatrix whose explicit inverse is applied every iteration."""
    q = q.copy()
    p = p.copy()
    mu = np.zeros_like(q)
    jac_prev = system.jacob_constr(q_prev)
    adt = abs(time_step)
    error = np.nan
    try:
        gram = _chkfinite(jac_prev @ _scaled_inv_metric(system, adt, jac_prev.T))
        inv_gram = dense_spd_inverse(gram)
        for i in range(max_iters):
            c = system.constr(q)
            error = maximum_norm(c)
            delta_mu = jac_prev.T @ (inv_gram @ c)
            delta_pos = _scaled_inv_metric(system, adt, delta_mu)
            if error > divergence_tol or np.isnan(error):
                raise OracleIntegratorError(STATUS_CONVERGENCE, f"quasi-Newton diverged at {i}")
            if error < constraint_tol and maximum_norm(delta_pos) < position_tol:
                p -= np.sign(time_step) * mu
                if counts is not None:
                    counts.setdefault("newton_iters", []).append(i + 1)
                return q, p
            mu += delta_mu
            q -= delta_pos
    except (ValueError, _LinAlgError) as e:
        raise OracleIntegratorError(STATUS_CONVERGENCE, f"{type(e)} in quasi-Newton solver") from e
    raise OracleIntegratorError(STATUS_CONVERGENCE, f"quasi-Newton did not converge, |c|={error}")


def solve_projection_onto_manifold_newton_with_line_search(
    q, p, q_prev, time_step, system, constraint_tol=1e-9, position_tol=1e-8,
    divergence_tol=1e10, max_iters=50, max_line_search_iters=10, counts=None,
):
    """solvers.py:472-614 for a Euclidean metric (order of tests and the line-search bookkeeping
    exactly as in the reference)."""
    q = q.copy()
    p = p.copy()
    mu = np.zeros_like(q)
    jac_prev = system.jacob_constr(q_prev)
    adt = abs(time_step)
    delta_pos, step_size = None, None
    error = np.nan
    for i in range(max_iters):
        try:
            jac = system.jacob_constr(q)
            c = system.constr(q)
            error = maximum_norm(c)
            if i > 0 and (error > divergence_tol or np.isnan(error)):
                raise OracleIntegratorError(STATUS_CONVERGENCE, f"Newton diverged at {i}")
            if error < constraint_tol and (
                i == 0 or maximum_norm(step_size * delta_pos) < position_tol
            ):
                p -= np.sign(time_step) * mu
                if counts is not None:
                    counts.setdefault("newton_iters", []).append(i + 1)
                return q, p
            res_jac = _chkfinite(jac @ _scaled_inv_metric(system, adt, jac_prev.T))
            lu_piv = sla.lu_factor(res_jac, check_finite=False)
            delta_mu = jac_prev.T @ sla.lu_solve(lu_piv, c, 0, check_finite=False)
            delta_pos = -_scaled_inv_metric(system, adt, delta_mu)
            pos_curr = q.copy()
            step_size = 1.0
            for _ in range(max_line_search_iters):
                q = pos_curr + step_size * delta_pos
                new_error = maximum_norm(system.constr(q))
                if new_error < error:
                    break
                step_size *= 0.5
            mu += step_size * delta_mu
        except (ValueError, _LinAlgError) as e:
            raise OracleIntegratorError(STATUS_CONVERGENCE, f"{type(e)} in Newton solver") from e
    raise OracleIntegratorError(STATUS_CONVERGENCE, f"Newton did not converge, |c|={error}")


PROJECTION_SOLVERS = {
    "newton": solve_projection_onto_manifold_newton,
    "quasi_newton": solve_projection_onto_manifold_quasi_newton,
    "newton_with_line_search": solve_projection_onto_manifold_newton_with_line_search,
}


def _scaled_inv_metric(system, scale, a):
    """``(scale * metric.inv) @ a`` with the reference's order of operations."""
    m = system.metric
    if m.kind == "identity":
        return scale * a  # PositiveScaledIdentityMatrix._left_matrix_multiply
    if m.kind == "diagonal":
        d = scale * m.inv_diagonal  # _scalar_multiply then diagonal product
        return d[:, None] * a if a.ndim == 2 else d * a
    return (scale * m.inv_array) @ a  # DenseDefiniteMatrix._scalar_multiply (:1138-1154)


def constrained_leapfrog_step(
    q,
    p,
    time_step,
    system,
    n_inner_step=1,
    reverse_check_tol=2e-8,
    projection_solver_kwargs=None,
    counts=None,
    projection_solver="newton",
):
    """One ``ConstrainedLeapfrogIntegrator.step`` (integrators.py:929-984)."""
    kw = {} if projection_solver_kwargs is None else projection_solver_kwargs
    q = np.array(q, dtype=np.float64)
    p = np.array(p, dtype=np.float64)
    tgt = system.target

    def h2_flow_retraction(q, p, q_prev, dt):
        # :929-942 ; h2_flow systems.py:362-363
        q = q + dt * system.inv_metric_mat(p)
        return PROJECTION_SOLVERS[projection_solver](q, p, q_prev, dt, system, counts=counts, **kw)

    try:
        # _step_a(dt/2) :947-949
        p = p - (0.5 * time_step) * system.dh1_dpos(q)
        p = system.project_onto_cotangent_space(p, q)
        # _step_b(dt) :951-979
        dt_inner = time_step / n_inner_step
        for _ in range(n_inner_step):
            q_prev = q.copy()
            q, p = h2_flow_retraction(q, p, q_prev, dt_inner)
            p = system.project_onto_cotangent_space(p, q)
            q_back, _ = h2_flow_retraction(q.copy(), p.copy(), q, -dt_inner)
            rev_diff = maximum_norm(q_back - q_prev)
            if rev_diff > reverse_check_tol:
                raise OracleIntegratorError(STATUS_NON_REVERSIBLE, f"rev diff {rev_diff}")
        # _step_a(dt/2)
        p = p - (0.5 * time_step) * system.dh1_dpos(q)
        p = system.project_onto_cotangent_space(p, q)
    except (ValueError, _LinAlgError) as e:
        raise OracleIntegratorError(STATUS_LINALG, str(e)) from e
    return q, p


# --------------------------------------------------------------------------------------
# Batch drivers (loop over chains, convert failures to status codes)
# --------------------------------------------------------------------------------------


def run_batch(step_fn, q, p, dirs, n_steps):
    """Apply ``step_fn(q_i, p_i, dir_i) -> (q_i, p_i)`` ``n_steps`` times to every chain.

    A chain whose step raises keeps the state it had *before* the failing step and records
    the status code; it takes no further steps (the reference's transitions terminate the
    trajectory on ``IntegratorError``: transitions.py:292-295).
    """
    q = np.array(q, dtype=np.float64)
    p = np.array(p, dtype=np.float64)
    n = q.shape[0]
    status = np.zeros(n, dtype=np.int32)
    n_done = np.zeros(n, dtype=np.int32)
    dirs = np.ones(n, dtype=np.int32) if dirs is None else np.broadcast_to(dirs, (n,))
    for i in range(n):
        qi, pi = q[i], p[i]
        for _ in range(n_steps):
            try:
                qi, pi = step_fn(qi, pi, int(dirs[i]))
            except OracleIntegratorError as e:
                status[i] = e.status
                break
            n_done[i] += 1
        q[i], p[i] = qi, pi
    return q, p, status, n_done


# --------------------------------------------------------------------------------------
# "Next" row N1: static-HMC transition (momentum refresh + Metropolis accept)
# --------------------------------------------------------------------------------------


def euclidean_sample_momentum(metric):
    """EuclideanMetricSystem.sample_momentum (systems.py:365-366)."""
    return lambda q, rng: metric.sqrt_matvec(rng.standard_normal(q.shape))


def constrained_sample_momentum(system):
    """ConstrainedTractableFlowSystem.sample_momentum (systems.py:613-616)."""
    return lambda q, rng: system.project_onto_cotangent_space(
        system.metric.sqrt_matvec(rng.standard_normal(q.shape)), q)


def riemannian_sample_momentum(system):
    """RiemannianMetricSystem.sample_momentum (systems.py:1401-1402)."""
    return lambda q, rng: system.metric(q).sqrt_matvec(rng.normal(size=q.shape))


def static_hmc_transition(q, p, d, rng, step_fn, h_fn, sample_momentum, n_step):
    """IndependentMomentumTransition.sample (transitions.py:136-142) followed by
    MetropolisIntegrationTransition._sample_n_step (transitions.py:275-315) for one chain.
    ``step_fn(q, p, dir) -> (q, p)`` raises OracleIntegratorError on failure;
    ``sample_momentum(q, rng)`` is one of the three constructors above.
    Returns ``(q, p, dir, stats)``."""
    p = sample_momentum(q, rng)
    if isinstance(n_step, tuple):
        # MetropolisRandomIntegrationTransition.sample (transitions.py:396-402)
        n_step = int(rng.integers(*n_step))
    h_init = h_fn(q, p)
    qp, pp, dp = q, p, d
    error = None
    n_done = 0
    try:
        for _ in range(n_step):
            qp, pp = step_fn(qp, pp, dp)
            n_done += 1
    except OracleIntegratorError as e:
        error = e.status
    else:
        dp = -dp
    if n_done > 0 or error is None:
        h_diff = h_init - h_fn(qp, pp)
        accept_prob = 0.0 if np.isnan(h_diff) else np.exp(min(0, h_diff))
    else:
        accept_prob = 0.0
    stats = {
        "n_step": n_done,
        "metrop_accept_prob": accept_prob,
        "accept_stat": accept_prob if error is None else 0.0,
        "convergence_error": error == STATUS_CONVERGENCE,
        "non_reversible_step": error == STATUS_NON_REVERSIBLE,
    }
    accepted = error is None and rng.uniform() < accept_prob
    if accepted:
        q, p, d = qp, pp, dp
    d = -d
    stats["accepted"] = bool(accepted)
    return q, p, d, stats


# --------------------------------------------------------------------------------------
# Adapters (row N3), one chain at a time, states as dicts -- adapters.py:126-648
# --------------------------------------------------------------------------------------

LOG_STEP_SIZE_REDUCERS = {
    # adapters.py:126-159
    "arithmetic_mean_log_step_size_reducer": lambda xs: sum(math.exp(x) for x in xs) / len(xs),
    "geometric_mean_log_step_size_reducer": lambda xs: math.exp(sum(xs) / len(xs)),
    "min_log_step_size_reducer": lambda xs: math.exp(min(xs)),
}


def find_init_step_size(q, p, d, step_eps_fn, h_fn, max_iters=100):
    """DualAveragingStepSizeAdapter._find_and_set_init_step_size (adapters.py:285-352).
    ``step_eps_fn(q, p, d, eps) -> (q, p)`` raises OracleIntegratorError on failure."""
    h_init = h_fn(q, p)
    if np.isnan(h_init):
        raise RuntimeError("Hamiltonian evaluating to NaN at initial state.")
    eps = 1
    threshold = math.log(2)
    too_big = False
    for s in range(max_iters):
        try:
            qn, pn = step_eps_fn(q, p, d, eps)
            delta_h = abs(h_init - h_fn(qn, pn))
            if s == 0 or np.isnan(delta_h):
                too_big = bool(np.isnan(delta_h) or delta_h > threshold)
            if (too_big and delta_h <= threshold) or (not too_big and delta_h > threshold):
                return eps
            eps = eps / 2 if too_big else eps * 2
        except OracleIntegratorError:
            too_big = True
            eps = eps / 2
    raise RuntimeError("Could not find reasonable initial step size")


class DualAveragingOracle:
    """adapters.py:172-391."""

    is_fast = True

    def __init__(self, adapt_stat_target=0.8, log_step_size_reg_target=None,
                 log_step_size_reg_coefficient=0.05, iter_decay_coeff=0.75, iter_offset=10,
                 max_init_step_size_iters=100,
                 log_step_size_reducer="arithmetic_mean_log_step_size_reducer"):
        self.target = adapt_stat_target
        self.reg_target = log_step_size_reg_target
        self.reg_coefficient = log_step_size_reg_coefficient
        self.decay = iter_decay_coeff
        self.offset = iter_offset
        self.max_init = max_init_step_size_iters
        self.reducer = LOG_STEP_SIZE_REDUCERS[log_step_size_reducer]

    def initialize(self, q, p, d, ctx):
        eps = find_init_step_size(q, p, d, ctx.step_eps, ctx.h, self.max_init)
        ctx.step_size = eps
        reg = math.log(10 * eps) if self.reg_target is None else self.reg_target
        return {"iter": 0, "smoothed_log_step_size": 0.0, "adapt_stat_error": 0.0,
                "log_step_size_reg_target": reg}

    def update(self, st, q, stats, ctx):  # adapters.py:354-373
        st["iter"] += 1
        w = 1 / (self.offset + st["iter"])
        st["adapt_stat_error"] *= 1 - w
        st["adapt_stat_error"] += w * (self.target - stats["accept_stat"])
        sw = (1 / st["iter"]) ** self.decay
        log_eps = st["log_step_size_reg_target"] - (
            st["adapt_stat_error"] * st["iter"] ** 0.5 / self.reg_coefficient)
        st["smoothed_log_step_size"] *= 1 - sw
        st["smoothed_log_step_size"] += sw * log_eps
        ctx.step_size = math.exp(log_eps)

    def finalize(self, states, ctx):  # adapters.py:375-390
        ctx.step_size = self.reducer([s["smoothed_log_step_size"] for s in states])
        return False


class OnlineVarianceOracle:
    """adapters.py:394-518."""

    is_fast = False

    def __init__(self, reg_iter_offset=5, reg_scale=1e-3):
        self.reg_iter_offset, self.reg_scale = reg_iter_offset, reg_scale

    def initialize(self, q, p, d, ctx):
        return {"iter": 0, "mean": np.zeros_like(q), "m2": np.zeros_like(q)}

    def update(self, st, q, stats, ctx):  # Welford, adapters.py:446-458
        st["iter"] += 1
        diff = q - st["mean"]
        st["mean"] += diff / st["iter"]
        st["m2"] += diff * (q - st["mean"])

    def _merge(self, states, outer):  # Chan et al. / Schubert-Gertz, adapters.py:487-505, 615-634
        for i, st in enumerate(states):
            if i == 0:
                n_iter, mean_est, m2 = st["iter"], st["mean"].copy(), st["m2"].copy()
            else:
                n_prev = n_iter
                n_iter += st["iter"]
                mean_diff = mean_est - st["mean"]
                mean_est *= n_prev
                mean_est += st["iter"] * st["mean"]
                mean_est /= n_iter
                m2 += st["m2"]
                m2 += (np.outer(mean_diff, mean_diff) if outer else mean_diff**2) * (
                    st["iter"] * n_prev) / n_iter
        return n_iter, m2

    def finalize(self, states, ctx):  # adapters.py:471-516
        n_iter, var_est = self._merge(states, outer=False)
        if n_iter < 2:
            raise RuntimeError("At least two chain samples required")
        var_est /= n_iter - 1
        if self.reg_iter_offset is not None and self.reg_iter_offset != 0:
            var_est *= n_iter / (self.reg_iter_offset + n_iter)
            var_est += self.reg_scale * (self.reg_iter_offset / (self.reg_iter_offset + n_iter))
        ctx.metric = DiagonalMetric(1.0 / var_est)  # PositiveDiagonalMatrix(var_est).inv
        return True  # momenta are resampled


class OnlineCovarianceOracle(OnlineVarianceOracle):
    """adapters.py:521-648."""

    def initialize(self, q, p, d, ctx):
        return {"iter": 0, "mean": np.zeros_like(q), "m2": np.zeros((q.shape[0],) * 2)}

    def update(self, st, q, stats, ctx):  # adapters.py:576-590
        st["iter"] += 1
        diff = q - st["mean"]
        st["mean"] += diff / st["iter"]
        st["m2"] += diff[None, :] * (q - st["mean"])[:, None]

    def finalize(self, states, ctx):  # adapters.py:603-648
        n_iter, covar = self._merge(states, outer=True)
        if n_iter < 2:
            raise RuntimeError("At least two chain samples required")
        covar /= n_iter - 1
        covar *= n_iter / (self.reg_iter_offset + n_iter)
        covar[np.diag_indices_from(covar)] += self.reg_scale * (
            self.reg_iter_offset / (self.reg_iter_offset + n_iter))
        ctx.metric = CovarianceFactoredMetric(covar)  # DensePositiveDefiniteMatrix(covar).inv
        return True


class CovarianceFactoredMetric:
    """``DensePositiveDefiniteMatrix(covar).inv`` as a metric: array ``L^-T L^-1`` with factor
    ``L^-T`` (matrices.py:1183-1188, 1209-1216); its own ``.inv`` multiplies by ``L L^T``
    (matrices.py:1041-1046, 1060-1061) and ``sqrt @ z`` solves ``L^T x = z`` (:897-903)."""

    kind = "dense"

    def __init__(self, covar):
        self.chol = dense_spd_factor(covar)
        self.array = dense_spd_inverse(covar, self.chol)
        self.inv_array = self.chol @ self.chol.T

    def inv_matvec(self, v):
        return self.inv_array @ v

    def sqrt_matvec(self, v):
        return sla.solve_triangular(self.chol.T, v, lower=False, check_finite=False)


# --------------------------------------------------------------------------------------
# Dynamic (NUTS) integration transitions -- transitions.py:415-860, restated ITERATIVELY
# (the recursion of `_build_tree` unrolled into a binary-counter stack: after leaf number k of a
# doubling, one merge per trailing zero bit of k) so that the device kernel can follow it line by
# line.  Random numbers are consumed in the reference's order: [slice: one at the start,]
# per doubling one for the direction, one per completed internal node in post-order, one for
# the progressive acceptance.
# --------------------------------------------------------------------------------------


def _log1p_exp(val):  # utils.py:50-54
    return val + math.log1p(math.exp(-val)) if val > 0.0 else math.log1p(math.exp(val))


def _log_sum_exp(a, b):  # utils.py:65-71
    if a == -math.inf and b == -math.inf:
        return -math.inf
    return a + _log1p_exp(b - a) if a > b else b + _log1p_exp(a - b)


def _exp(x):  # LogRepFloat.val (utils.py:110-115)
    try:
        return math.exp(x)
    except OverflowError:
        return math.inf


def _no_u_turn(kind, vel_fn, q1, p1, q2, p2, sum_mom):
    """euclidean_no_u_turn_criterion / riemannian_no_u_turn_criterion (transitions.py:405-470)."""
    w = (q2 - q1) if kind == "euclidean" else sum_mom
    return bool(np.sum(vel_fn(q1, p1) * w) < 0 or np.sum(vel_fn(q2, p2) * w) < 0)


def nuts_transition(q, p, uniform, step_fn, h_fn, vel_fn, max_tree_depth=10, max_delta_h=1000.0,
                    criterion="riemannian", extra_checks=True, variant="multinomial"):
    """DynamicIntegrationTransition.sample (transitions.py:712-770) with
    MultinomialDynamicIntegrationTransition (:773-809) or SliceDynamicIntegrationTransition
    (:812-858) weights.  ``uniform()`` returns the chain's next ``rng.uniform()``;
    ``step_fn(q, p, dir)`` raises OracleIntegratorError; ``vel_fn(q, p)`` is ``system.dh_dmom``.
    Returns ``(q, p, stats)``."""
    multinomial = variant == "multinomial"
    h_init = h_fn(q, p)
    log_u = None if multinomial else math.log(uniform()) - h_init  # :832-839

    def leaf_weight(h):  # _weight_function
        return -h if multinomial else int(log_u <= -h)

    def add_w(a, b):
        return _log_sum_exp(a, b) if multinomial else a + b

    def ratio(num, den):  # _weight_ratio, as the probability the comparison `u < .` sees
        if multinomial:
            return min(_exp(num - den), 1)  # NaN (both -inf) compares False, like LogRepFloat
        return min(num / den, 1) if den > 0 else min(num, 1)

    def turn(t, neg_sub, pos_sub):  # _termination_criterion (:528-556)
        if _no_u_turn(criterion, vel_fn, t["nq"], t["np"], t["pq"], t["pp"], t["sum"]):
            return True
        if t["depth"] > 1 and extra_checks:
            return _no_u_turn(criterion, vel_fn, neg_sub["nq"], neg_sub["np"], pos_sub["nq"],
                              pos_sub["np"], neg_sub["sum"] + pos_sub["np"]) or _no_u_turn(
                criterion, vel_fn, neg_sub["pq"], neg_sub["pp"], pos_sub["pq"], pos_sub["pp"],
                pos_sub["sum"] + neg_sub["pp"])
        return False

    def merge(neg_sub, pos_sub):  # _merge_subtrees (:571-581)
        return {"nq": neg_sub["nq"], "np": neg_sub["np"], "pq": pos_sub["pq"], "pp": pos_sub["pp"],
                "sum": neg_sub["sum"] + pos_sub["sum"], "w": add_w(neg_sub["w"], pos_sub["w"]),
                "depth": neg_sub["depth"] + 1}

    stats = {"n_step": 0, "reject_prob": 1.0, "diverging": False, "convergence_error": False,
             "non_reversible_step": False}
    sum_accept = 0.0
    tree = {"nq": q, "np": p, "pq": q, "pp": p, "sum": np.asarray(p), "w": leaf_weight(h_init),
            "depth": 0}
    next_q, next_p = q, p
    # `dir` of the returned state object.  Leaf states carry the direction they were integrated
    # in; the INITIAL state object is both edges of the tree at first and has its `dir` overwritten
    # (`state.dir = direction`, :731) by every doubling that starts from it.  The value is read
    # again by the step-size initialisation of the next adaptive stage (adapters.py:321).
    next_is_init, init_dir, next_dir = True, None, None
    edge_is_init = {1: True, -1: True}
    depth = 0
    for depth in range(max_tree_depth):
        direction = 2 * (uniform() < 0.5) - 1  # :729
        if edge_is_init[direction]:
            init_dir = direction
        cq, cp = (tree["pq"], tree["pp"]) if direction == 1 else (tree["nq"], tree["np"])
        # ---- _build_tree(depth, ...) unrolled (:610-710)
        stack, terminate, cur = {}, False, None
        for k in range(1, 2**depth + 1):
            try:
                cq, cp = step_fn(cq, cp, direction)
                h = h_fn(cq, cp)
                h = math.inf if np.isnan(h) else h
                cur = {"nq": cq, "np": cp, "pq": cq, "pp": cp, "sum": np.asarray(cp),
                       "w": leaf_weight(h), "depth": 0, "prop": (cq, cp)}
                h_diff = h_init - h
                sum_accept += 0.0 if np.isnan(h_diff) else math.exp(min(0, h_diff))
                stats["n_step"] += 1
                if (h - h_init if multinomial else h + log_u) > max_delta_h:  # _check_divergence
                    stats["diverging"] = True
                    terminate = True
            except OracleIntegratorError as e:
                stats["convergence_error"] |= e.status == STATUS_CONVERGENCE
                stats["non_reversible_step"] |= e.status == STATUS_NON_REVERSIBLE
                terminate = True
            if terminate:
                break
            level, kk = 0, k
            while kk % 2 == 0:  # one merge per trailing zero bit of the leaf count
                inner, outer = stack.pop(level), cur
                neg_sub, pos_sub = (inner, outer) if direction == 1 else (outer, inner)
                cur = merge(neg_sub, pos_sub)
                accept_outer = ratio(outer["w"], cur["w"])
                cur["prop"] = outer["prop"] if uniform() < accept_outer else inner["prop"]
                if turn(cur, neg_sub, pos_sub):
                    terminate = True
                    break
                level, kk = level + 1, kk // 2
            if terminate:
                break
            stack[level] = cur
        if terminate:
            break
        new = cur
        accept_prob = ratio(new["w"], tree["w"])  # :742
        if uniform() < accept_prob:
            next_q, next_p = new["prop"]
            next_is_init, next_dir = False, direction
        stats["reject_prob"] *= 1.0 - accept_prob
        edge_is_init[direction] = False
        neg_sub, pos_sub = (tree, new) if direction == 1 else (new, tree)
        tree = merge(neg_sub, pos_sub)
        if turn(tree, neg_sub, pos_sub):
            break
    stats["av_metrop_accept_prob"] = sum_accept / stats["n_step"] if stats["n_step"] > 0 else 0.0
    failed = stats["diverging"] or stats["convergence_error"] or stats["non_reversible_step"]
    stats["accept_stat"] = 0.0 if failed else stats["av_metrop_accept_prob"]
    stats["tree_depth"] = depth
    stats["dir"] = init_dir if next_is_init else next_dir
    return next_q, next_p, stats
