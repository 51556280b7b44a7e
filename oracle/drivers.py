"""TEST INFRASTRUCTURE ONLY -- drive a ``mici_b200.problems.Problem`` through

* the oracle port (``oracle/mici_oracle.py``)            -> ``oracle_run``
* the unmodified reference, when ``/root/reference/src`` exists -> ``reference_run``

Both return ``dict(pos, mom, status, n_done, h, ...)`` with identical conventions so that
``oracle/make_golden.py`` and the tests can compare them entry by entry.
"""

from __future__ import annotations

import os
import sys

import numpy as np

from . import mici_oracle as mo
from . import targets as tg

REFERENCE_SRC = "/root/reference/src"


def build_target(problem):
    name, kw = problem.target, problem.target_params
    if name == "std_gaussian":
        return tg.StdGaussian(**kw)
    if name == "neal_funnel":
        return tg.NealFunnel(**kw)
    if name == "banana":
        return tg.Banana(**kw)
    if name == "quadratic":
        return tg.Quadratic(**kw)
    if name == "torus":
        return tg.Torus(**kw)
    if name == "sphere":
        return tg.Sphere(**kw)
    raise KeyError(name)


def build_metric_model(problem):
    if problem.metric_model == "rank1":
        return tg.Rank1Metric(**problem.metric_params)
    if problem.metric_model is None:
        return None
    raise KeyError(problem.metric_model)


# ----------------------------------------------------------------------------- oracle


def oracle_step_fn(problem, counts=None, **overrides):
    """Return ``(step_fn(q, p, dir) -> (q, p), h_fn(q, p), system)`` for ``problem``."""
    target = build_target(problem)
    ikw = dict(problem.integrator_kwargs)
    ikw.update(overrides)
    eps = problem.step_size
    if problem.integrator == "leapfrog":
        metric = mo.coerce_metric(problem.metric)

        def step(q, p, d):
            return mo.leapfrog_steps(q, p, d * eps, 1, target, metric)

        return step, (lambda q, p: mo.euclidean_h(q, p, target, metric)), None
    if problem.integrator in mo.BCSS_FREE_COEFFICIENTS:
        metric = mo.coerce_metric(problem.metric)
        coefs = mo.composition_coefficients(mo.BCSS_FREE_COEFFICIENTS[problem.integrator])

        def step(q, p, d):
            return mo.composition_steps(q, p, d * eps, 1, target, metric, coefs)

        return step, (lambda q, p: mo.euclidean_h(q, p, target, metric)), None
    if problem.integrator in ("implicit_leapfrog", "implicit_midpoint"):
        kind = "softabs" if problem.system == "softabs_riemannian" else "dense"
        system = mo.RiemannianSystem(
            target,
            kind,
            metric_model=build_metric_model(problem),
            softabs_coeff=problem.system_kwargs.get("softabs_coeff", 1.0),
        )

        def step(q, p, d):
            c = {} if counts is None else counts
            fn = (mo.implicit_midpoint_step if problem.integrator == "implicit_midpoint"
                  else mo.implicit_leapfrog_step)
            out = fn(q, p, d * eps, system, counts=c, **ikw)
            if counts is not None:
                counts.setdefault("all_fp_iters", []).append(list(c.get("fp_iters", [])))
            return out

        return step, system.h, system
    if problem.integrator == "constrained_leapfrog":
        system = mo.ConstrainedSystem(target, problem.metric)

        def step(q, p, d):
            return mo.constrained_leapfrog_step(q, p, d * eps, system, counts=counts, **ikw)

        return step, system.h, system
    raise KeyError(problem.integrator)


def oracle_run(problem, n_steps, dirs=None, chains=None, counts=None, **overrides):
    """Step chains ``chains`` (default: all) of ``problem`` ``n_steps`` times through the oracle."""
    step, h_fn, _ = oracle_step_fn(problem, counts=counts, **overrides)
    sl = slice(None) if chains is None else chains
    q0, p0 = problem.pos[sl], problem.mom[sl]
    q, p, status, n_done = mo.run_batch(step, q0, p0, dirs, n_steps)
    h = np.array([_safe_h(h_fn, q[i], p[i]) for i in range(q.shape[0])])
    h0 = np.array([_safe_h(h_fn, q0[i], p0[i]) for i in range(q.shape[0])])
    return {"pos": q, "mom": p, "status": status, "n_done": n_done, "h": h, "h_init": h0}


def _safe_h(h_fn, q, p):
    try:
        return float(h_fn(q, p))
    except Exception:  # noqa: BLE001 - non-finite states
        return np.nan


# -------------------------------------------------------------------------- reference


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_SRC, "mici"))


def import_reference():
    """Import the unmodified reference package from /root/reference/src (read-only)."""
    if not reference_available():
        raise ImportError("reference not present on this machine")
    sys.dont_write_bytecode = True
    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)
    import mici  # noqa: PLC0415

    return mici


def build_reference(problem, **overrides):
    """Build the reference ``(system, integrator)`` pair for ``problem``.

    All derivative callables are passed explicitly (no autodiff backend).
    """
    mici = import_reference()
    target = build_target(problem)
    ikw = dict(problem.integrator_kwargs)
    ikw.update(overrides)
    if problem.system == "euclidean":
        system = mici.systems.EuclideanMetricSystem(
            neg_log_dens=target.neg_log_dens,
            metric=problem.metric,
            grad_neg_log_dens=target.grad_neg_log_dens,
        )
    elif problem.system == "softabs_riemannian":
        system = mici.systems.SoftAbsRiemannianMetricSystem(
            neg_log_dens=target.neg_log_dens,
            grad_neg_log_dens=target.grad_neg_log_dens,
            hess_neg_log_dens=target.hess_neg_log_dens,
            mtp_neg_log_dens=target.mtp_neg_log_dens,
            softabs_coeff=problem.system_kwargs.get("softabs_coeff", 1.0),
        )
    elif problem.system == "dense_riemannian":
        mm = build_metric_model(problem)
        system = mici.systems.DenseRiemannianMetricSystem(
            neg_log_dens=target.neg_log_dens,
            metric_func=mm.metric_func,
            vjp_metric_func=mm.vjp_metric_func,
            grad_neg_log_dens=target.grad_neg_log_dens,
        )
    elif problem.system == "constrained_euclidean":
        system = mici.systems.DenseConstrainedEuclideanMetricSystem(
            neg_log_dens=target.neg_log_dens,
            constr=target.constr,
            metric=problem.metric,
            dens_wrt_hausdorff=True,
            grad_neg_log_dens=target.grad_neg_log_dens,
            jacob_constr=target.jacob_constr,
        )
    else:
        raise KeyError(problem.system)
    cls = {
        "leapfrog": mici.integrators.LeapfrogIntegrator,
        "implicit_leapfrog": mici.integrators.ImplicitLeapfrogIntegrator,
        "constrained_leapfrog": mici.integrators.ConstrainedLeapfrogIntegrator,
        "implicit_midpoint": mici.integrators.ImplicitMidpointIntegrator,
        "bcss2": mici.integrators.BCSSTwoStageIntegrator,
        "bcss3": mici.integrators.BCSSThreeStageIntegrator,
        "bcss4": mici.integrators.BCSSFourStageIntegrator,
    }[problem.integrator]
    if isinstance(ikw.get("projection_solver"), str):
        ikw["projection_solver"] = getattr(
            mici.solvers, "solve_projection_onto_manifold_" + ikw["projection_solver"])
    if isinstance(ikw.get("fixed_point_solver"), str):
        ikw["fixed_point_solver"] = getattr(mici.solvers, "solve_fixed_point_" + ikw["fixed_point_solver"])
    return system, cls(system, problem.step_size, **ikw)


def reference_run(problem, n_steps, dirs=None, chains=None, **overrides):
    """Step chains through the unmodified reference, one ``ChainState`` at a time."""
    mici = import_reference()
    system, integrator = build_reference(problem, **overrides)
    sl = slice(None) if chains is None else chains
    q0, p0 = problem.pos[sl], problem.mom[sl]
    n = q0.shape[0]
    dirs = np.ones(n, dtype=np.int32) if dirs is None else np.broadcast_to(dirs, (n,))
    q, p = q0.copy(), p0.copy()
    status = np.zeros(n, dtype=np.int32)
    n_done = np.zeros(n, dtype=np.int32)
    h = np.full(n, np.nan)
    h0 = np.full(n, np.nan)
    call_counts = []
    for i in range(n):
        state = mici.states.ChainState(
            pos=q0[i].copy(), mom=p0[i].copy(), dir=int(dirs[i]), _call_counts={}
        )
        h0[i] = _safe_h(lambda *_: system.h(state), None, None)
        for _ in range(n_steps):
            try:
                state = integrator.step(state)
            except mici.errors.ConvergenceError:
                status[i] = mo.STATUS_CONVERGENCE
                break
            except mici.errors.NonReversibleStepError:
                status[i] = mo.STATUS_NON_REVERSIBLE
                break
            except (mici.errors.LinAlgError, ValueError):
                status[i] = mo.STATUS_LINALG
                break
            n_done[i] += 1
        q[i], p[i] = state.pos, state.mom
        h[i] = _safe_h(lambda *_: system.h(state), None, None)
        call_counts.append(dict(state._call_counts))
    return {
        "pos": q,
        "mom": p,
        "status": status,
        "n_done": n_done,
        "h": h,
        "h_init": h0,
        "call_counts": call_counts,
    }


# ------------------------------------------------------------------ static HMC (row N1)


def _sqrt_matvec(problem):
    metric = mo.coerce_metric(problem.metric)
    return metric.sqrt_matvec


def oracle_hmc(problem, n_iter, n_step, seed, chains=None):
    """``n_iter`` static-HMC iterations per chain through the oracle; per-chain generators
    ``default_rng([seed, chain])``.  Euclidean systems only."""
    step, h_fn, _ = oracle_step_fn(problem)
    sl = slice(None) if chains is None else chains
    q0, p0 = problem.pos[sl], problem.mom[sl]
    n = q0.shape[0]
    sqrt_mv = _sqrt_matvec(problem)
    pos = np.empty((n_iter, n, q0.shape[1]))
    stats = {k: np.empty((n_iter, n)) for k in ("n_step", "metrop_accept_prob", "accept_stat", "accepted")}
    dirs = np.ones(n, dtype=np.int32)
    for i in range(n):
        rng = np.random.default_rng([seed, i])
        q, p, d = q0[i].copy(), p0[i].copy(), 1
        for it in range(n_iter):
            q, p, d, st = mo.static_hmc_transition(q, p, d, rng, step, h_fn, sqrt_mv, n_step)
            pos[it, i] = q
            for k in stats:
                stats[k][it, i] = st[k]
        dirs[i] = d
    return {"pos": pos, "dir": dirs, **stats}


def reference_hmc(problem, n_iter, n_step, seed, chains=None):
    """The same through the unmodified reference transitions (transitions.py:129-142, 256-352)."""
    mici = import_reference()
    system, integrator = build_reference(problem)
    sl = slice(None) if chains is None else chains
    q0, p0 = problem.pos[sl], problem.mom[sl]
    n = q0.shape[0]
    mom_tr = mici.transitions.IndependentMomentumTransition(system)
    int_tr = mici.transitions.MetropolisStaticIntegrationTransition(system, integrator, n_step)
    pos = np.empty((n_iter, n, q0.shape[1]))
    stats = {k: np.empty((n_iter, n)) for k in ("n_step", "metrop_accept_prob", "accept_stat")}
    dirs = np.ones(n, dtype=np.int32)
    for i in range(n):
        rng = np.random.default_rng([seed, i])
        state = mici.states.ChainState(pos=q0[i].copy(), mom=p0[i].copy(), dir=1)
        for it in range(n_iter):
            state, _ = mom_tr.sample(state, rng)
            state, st = int_tr.sample(state, rng)
            pos[it, i] = state.pos
            for k in stats:
                stats[k][it, i] = st[k]
        dirs[i] = state.dir
    return {"pos": pos, "dir": dirs, **stats}
