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

# The unmodified reference: the read-only checkout in the build container, else the verbatim copy
# that oracle/build_ref.sh places under oracle/_ref/ (git-ignored; travels to the GPU box).
_REF_CANDIDATES = (
    "/root/reference/src",
    os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref"),
)
REFERENCE_SRC = next(
    (d for d in _REF_CANDIDATES if os.path.isdir(os.path.join(d, "mici"))), _REF_CANDIDATES[0]
)


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
    if name == "quartic":
        return tg.Quartic(**kw)
    if name == "torus":
        return tg.Torus(**kw)
    if name == "sphere":
        return tg.Sphere(**kw)
    if name == "multi_sphere":
        return tg.MultiSphere(**kw)
    raise KeyError(name)


def build_metric_model(problem):
    if problem.metric_model == "rank1":
        return tg.Rank1Metric(**{k: v for k, v in problem.metric_params.items()
                                 if k in ("base", "coeff")})
    if problem.metric_model == "hadamard":
        return tg.HadamardMetric(**{k: v for k, v in problem.metric_params.items()
                                    if k in ("base", "scale", "coeff")})
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
    if problem.system == "gaussian_euclidean":
        metric = mo.coerce_metric(problem.metric)
        coefs = (0.5, 1.0, 0.5)
        if problem.integrator in mo.BCSS_FREE_COEFFICIENTS:
            coefs = mo.composition_coefficients(mo.BCSS_FREE_COEFFICIENTS[problem.integrator])

        def step(q, p, d):
            return mo.gaussian_composition_steps(q, p, d * eps, 1, target, metric, coefs)

        return step, (lambda q, p: mo.gaussian_euclidean_h(q, p, target, metric)), None
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
        system = mo.ConstrainedSystem(
            target, problem.metric,
            dens_wrt_hausdorff=problem.system_kwargs.get("dens_wrt_hausdorff", True))

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
    """Import the unmodified reference package (``REFERENCE_SRC``, never written to)."""
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
    elif problem.system == "gaussian_euclidean":
        system = mici.systems.GaussianEuclideanMetricSystem(
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
            dens_wrt_hausdorff=problem.system_kwargs.get("dens_wrt_hausdorff", True),
            grad_neg_log_dens=target.grad_neg_log_dens,
            jacob_constr=target.jacob_constr,
            mhp_constr=target.mhp_constr,
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


def _sample_momentum(problem, system):
    if problem.system in ("euclidean", "gaussian_euclidean"):
        return mo.euclidean_sample_momentum(mo.coerce_metric(problem.metric))
    if problem.system == "constrained_euclidean":
        return mo.constrained_sample_momentum(system)
    return mo.riemannian_sample_momentum(system)


def oracle_hmc(problem, n_iter, n_step, seed, chains=None):
    """``n_iter`` static-HMC iterations per chain through the oracle; per-chain generators
    ``default_rng([seed, chain])``."""
    step, h_fn, system = oracle_step_fn(problem)
    sl = slice(None) if chains is None else chains
    q0, p0 = problem.pos[sl], problem.mom[sl]
    n = q0.shape[0]
    sample_mom = _sample_momentum(problem, system)
    pos = np.empty((n_iter, n, q0.shape[1]))
    stats = {k: np.empty((n_iter, n)) for k in ("n_step", "metrop_accept_prob", "accept_stat", "accepted")}
    dirs = np.ones(n, dtype=np.int32)
    for i in range(n):
        rng = np.random.default_rng([seed, i])
        q, p, d = q0[i].copy(), p0[i].copy(), 1
        for it in range(n_iter):
            q, p, d, st = mo.static_hmc_transition(q, p, d, rng, step, h_fn, sample_mom, n_step)
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
    if isinstance(n_step, tuple):
        int_tr = mici.transitions.MetropolisRandomIntegrationTransition(system, integrator, n_step)
    else:
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


# ------------------------------------------------------- adaptive staged sampling (row N3)


def _make_reference_adapters(mici, specs):
    out = []
    for name, kw in specs:
        cls = {
            "dual_averaging": mici.adapters.DualAveragingStepSizeAdapter,
            "online_variance": mici.adapters.OnlineVarianceMetricAdapter,
            "online_covariance": mici.adapters.OnlineCovarianceMetricAdapter,
        }[name]
        kw = dict(kw)
        if "log_step_size_reducer" in kw:
            kw["log_step_size_reducer"] = getattr(mici.adapters, kw["log_step_size_reducer"])
        out.append(cls(**kw))
    return out


def reference_sample_chains(problem, n_warm_up_iter, n_main_iter, n_step, seed, adapter_specs,
                            stager_kwargs=None, dynamic=None):
    """The reference's own ``StaticMetropolisHMC.sample_chains`` (samplers.py:1271-1432) with
    adapters and stager, sequential chains, warm-up traced.  Per-chain generators are the
    reference's (``default_rng(base.bit_generator.jumped(i))``, samplers.py:559-560)."""
    mici = import_reference()
    system, integrator = build_reference(problem)
    rng = np.random.default_rng(seed)
    if dynamic is None:
        sampler = mici.samplers.StaticMetropolisHMC(system, integrator, rng, n_step=n_step)
    else:  # samplers.py:1588-1690; `dynamic`: keyword arguments of the dynamic transition
        kw = dict(dynamic)
        variant = kw.pop("variant", "multinomial")
        # the sampler classes and the transition classes have different defaults
        # (samplers.py:1606-1609, 1714-1717 vs transitions.py:494-497): always pass both
        kw["termination_criterion"] = getattr(
            mici.transitions, kw.pop("criterion", "riemannian") + "_no_u_turn_criterion")
        kw["do_extra_subtree_checks"] = kw.pop("extra_checks", True)
        cls = (mici.samplers.DynamicMultinomialHMC if variant == "multinomial"
               else mici.samplers.DynamicSliceHMC)
        sampler = cls(system, integrator, rng, **kw)
    adapters = _make_reference_adapters(mici, adapter_specs)
    stager = None
    if stager_kwargs is not None:
        stager = mici.stagers.WindowedWarmUpStager(**stager_kwargs)
    init_states = [
        mici.states.ChainState(pos=problem.pos[i].copy(), mom=problem.mom[i].copy(), dir=1)
        for i in range(problem.n_chains)
    ]
    final_states, traces, stats = sampler.sample_chains(
        n_warm_up_iter, n_main_iter, init_states, adapters=adapters, stager=stager,
        trace_warm_up=True, n_worker=1, display_progress=False,
    )
    metric = getattr(system, "metric", None)
    if problem.system != "euclidean" or isinstance(metric, mici.matrices.IdentityMatrix):
        metric_arr = np.zeros(0)
    elif isinstance(metric, mici.matrices.PositiveDiagonalMatrix):
        metric_arr = np.asarray(metric.diagonal)
    else:
        metric_arr = np.asarray(metric.array)
    return {
        "pos": np.stack(traces["pos"], axis=1),  # [n_iter, n_chains, dim]
        "accept_stat": np.stack(stats["accept_stat"], axis=1),
        "n_step": np.stack(stats["n_step"], axis=1),
        **({} if dynamic is None else {"tree_depth": np.stack(stats["tree_depth"], axis=1),
                                      "diverging": np.stack(stats["diverging"], axis=1)}),
        "final_pos": np.stack([s.pos for s in final_states]),
        "final_mom": np.stack([s.mom for s in final_states]),
        "final_dir": np.array([s.dir for s in final_states], dtype=np.int32),
        "step_size": np.array(float(integrator.step_size)),
        "metric": metric_arr,
    }


class _AdaptiveContext:
    """What the adapters mutate: the integrator step size and (Euclidean systems) the metric."""

    def __init__(self, problem):
        import copy

        self.problem = copy.copy(problem)
        self.step_size = problem.step_size
        self.euclidean = problem.system == "euclidean" and problem.integrator == "leapfrog"
        if self.euclidean:
            self.target = build_target(problem)
            self.metric = mo.coerce_metric(problem.metric)
        else:
            _, self._h, self._system = oracle_step_fn(problem)
            self.metric = None

    def copy(self):
        c = object.__new__(type(self))
        c.__dict__.update(self.__dict__)
        return c

    def step_eps(self, q, p, d, eps):
        if self.euclidean:
            return mo.leapfrog_steps(q, p, d * eps, 1, self.target, self.metric)
        import copy

        prob = copy.copy(self.problem)
        prob.step_size = eps
        return oracle_step_fn(prob)[0](q, p, d)

    def step(self, q, p, d):
        return self.step_eps(q, p, d, self.step_size)

    def h(self, q, p):
        if self.euclidean:
            return mo.euclidean_h(q, p, self.target, self.metric)
        return self._h(q, p)

    def velocity(self, q, p):
        if self.euclidean:
            return self.metric.inv_matvec(p)
        return _velocity_fn(self.problem, self._system)(q, p)

    def sample_momentum(self):
        if self.euclidean:
            return mo.euclidean_sample_momentum(self.metric)
        return _sample_momentum(self.problem, self._system)


def _make_oracle_adapters(specs):
    cls = {"dual_averaging": mo.DualAveragingOracle, "online_variance": mo.OnlineVarianceOracle,
           "online_covariance": mo.OnlineCovarianceOracle}
    return [cls[name](**kw) for name, kw in specs]


def oracle_sample_chains(problem, stages, n_step, seed, adapter_specs, dynamic=None):
    """Staged adaptive static HMC through the oracle.  ``stages``: list of ``(n_iter, which)``
    with ``which`` one of ``"all"``, ``"fast"``, ``None`` (samplers.py:1075-1141 with the stage
    list of stagers.py).  Chains run one after the other inside every stage, each from a copy of
    the shared parameters, exactly like ``_sample_chains_sequential`` with per-chain deep-copied
    transitions (samplers.py:1118-1129); the adapters are finalised over all chains at the end
    of the stage (samplers.py:1131-1138)."""
    ctx = _AdaptiveContext(problem)
    adapters = _make_oracle_adapters(adapter_specs)
    n = problem.n_chains
    base = np.random.default_rng(seed)
    rngs = [np.random.default_rng(base.bit_generator.jumped(i)) for i in range(n)]
    q = [problem.pos[i].copy() for i in range(n)]
    p = [problem.mom[i].copy() for i in range(n)]
    d = [1] * n
    pos, acc, nst, eps_trace, depths, divs = [], [], [], [], [], []
    for n_iter, which in stages:
        active = [] if which is None else [a for a in adapters if which == "all" or a.is_fast]
        stage_pos = np.empty((n_iter, n, problem.dim))
        stage_acc = np.empty((n_iter, n))
        stage_nst = np.empty((n_iter, n))
        stage_eps = np.empty((n_iter, n))
        stage_depth = np.empty((n_iter, n))
        stage_div = np.empty((n_iter, n))
        chain_states = []
        for i in range(n):
            c = ctx.copy()
            states = [a.initialize(q[i], p[i], d[i], c) for a in active]
            for it in range(n_iter):
                stage_eps[it, i] = c.step_size
                if dynamic is None:
                    q[i], p[i], d[i], st = mo.static_hmc_transition(
                        q[i], p[i], d[i], rngs[i], c.step, c.h, c.sample_momentum(), n_step)
                else:
                    p[i] = c.sample_momentum()(q[i], rngs[i])
                    q[i], p[i], st = mo.nuts_transition(
                        q[i], p[i], rngs[i].uniform, c.step, c.h, c.velocity, **dynamic)
                    d[i] = st["dir"]
                    stage_depth[it, i], stage_div[it, i] = st["tree_depth"], st["diverging"]
                for a, a_st in zip(active, states):
                    a.update(a_st, q[i], st, c)
                stage_pos[it, i] = q[i]
                stage_acc[it, i] = st["accept_stat"]
                stage_nst[it, i] = st["n_step"]
            chain_states.append(states)
        for k, a in enumerate(active):
            if a.finalize([cs[k] for cs in chain_states], ctx):
                for i in range(n):
                    p[i] = ctx.sample_momentum()(q[i], rngs[i])
        pos.append(stage_pos), acc.append(stage_acc), nst.append(stage_nst)
        eps_trace.append(stage_eps), depths.append(stage_depth), divs.append(stage_div)
    metric = ctx.metric
    metric_arr = (np.zeros(0) if metric is None or metric.kind == "identity"
                  else metric.diagonal if metric.kind == "diagonal" else metric.array)
    return {
        "pos": np.concatenate(pos), "accept_stat": np.concatenate(acc),
        "n_step": np.concatenate(nst), "step_size_trace": np.concatenate(eps_trace),
        **({} if dynamic is None else {"tree_depth": np.concatenate(depths),
                                      "diverging": np.concatenate(divs)}),
        "final_pos": np.stack(q), "final_mom": np.stack(p),
        "final_dir": np.array(d, dtype=np.int32), "step_size": np.array(float(ctx.step_size)),
        "metric": metric_arr,
    }


# ------------------------------------------------------------- dynamic HMC / NUTS (row N4)


def _velocity_fn(problem, system):
    """``system.dh_dmom`` for the oracle systems."""
    if problem.system in ("euclidean", "gaussian_euclidean"):
        metric = mo.coerce_metric(problem.metric)
        return lambda q, p: metric.inv_matvec(p)
    if problem.system == "constrained_euclidean":
        return lambda q, p: system.inv_metric_mat(p)
    return lambda q, p: system.dh2_dmom(q, p)


NUTS_STATS = ("n_step", "av_metrop_accept_prob", "accept_stat", "reject_prob", "tree_depth",
              "diverging")


def oracle_nuts(problem, n_iter, seed, variant="multinomial", criterion="riemannian",
                max_tree_depth=10, max_delta_h=1000.0, extra_checks=True, chains=None):
    """``n_iter`` iterations of momentum refresh + dynamic integration transition per chain
    through the oracle (``mo.nuts_transition``); generators ``default_rng([seed, chain])``."""
    step, h_fn, system = oracle_step_fn(problem)
    vel = _velocity_fn(problem, system)
    sample_mom = _sample_momentum(problem, system)
    sl = slice(None) if chains is None else chains
    q0 = problem.pos[sl]
    n = q0.shape[0]
    pos = np.empty((n_iter, n, q0.shape[1]))
    dirs = np.empty((n_iter, n))
    stats = {k: np.empty((n_iter, n)) for k in NUTS_STATS}
    for i in range(n):
        rng = np.random.default_rng([seed, i])
        q = q0[i].copy()
        for it in range(n_iter):
            p = sample_mom(q, rng)
            q, p, st = mo.nuts_transition(
                q, p, rng.uniform, step, h_fn, vel, max_tree_depth=max_tree_depth,
                max_delta_h=max_delta_h, criterion=criterion, extra_checks=extra_checks,
                variant=variant)
            pos[it, i] = q
            dirs[it, i] = st["dir"]
            for k in NUTS_STATS:
                stats[k][it, i] = st[k]
    return {"pos": pos, "dir": dirs, **stats}


def reference_nuts(problem, n_iter, seed, variant="multinomial", criterion="riemannian",
                   max_tree_depth=10, max_delta_h=1000.0, extra_checks=True, chains=None):
    """The same through the reference's own transition classes (transitions.py:487-858)."""
    mici = import_reference()
    system, integrator = build_reference(problem)
    cls = (mici.transitions.MultinomialDynamicIntegrationTransition if variant == "multinomial"
           else mici.transitions.SliceDynamicIntegrationTransition)
    crit = (mici.transitions.riemannian_no_u_turn_criterion if criterion == "riemannian"
            else mici.transitions.euclidean_no_u_turn_criterion)
    int_tr = cls(system, integrator, max_tree_depth=max_tree_depth, max_delta_h=max_delta_h,
                 termination_criterion=crit, do_extra_subtree_checks=extra_checks)
    mom_tr = mici.transitions.IndependentMomentumTransition(system)
    sl = slice(None) if chains is None else chains
    q0, p0 = problem.pos[sl], problem.mom[sl]
    n = q0.shape[0]
    pos = np.empty((n_iter, n, q0.shape[1]))
    dirs = np.empty((n_iter, n))
    stats = {k: np.empty((n_iter, n)) for k in NUTS_STATS}
    for i in range(n):
        rng = np.random.default_rng([seed, i])
        state = mici.states.ChainState(pos=q0[i].copy(), mom=p0[i].copy(), dir=1)
        for it in range(n_iter):
            state, _ = mom_tr.sample(state, rng)
            state, st = int_tr.sample(state, rng)
            pos[it, i] = state.pos
            dirs[it, i] = state.dir
            for k in NUTS_STATS:
                stats[k][it, i] = st[k]
    return {"pos": pos, "dir": dirs, **stats}
