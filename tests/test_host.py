"""CPU tests of the host-side mirror of the reference interface (no kernels are launched)."""

import copy
import pickle

import numpy as np
import pytest
import torch

import mici_b200 as mb
from mici_b200 import engine, integrators, problems, solvers, systems, targets
from mici_b200.errors import (
    AdaptationError,
    ConvergenceError,
    IntegratorError,
    LinAlgError,
    NonReversibleStepError,
    ReadOnlyStateError,
    raise_for_status,
)


def test_chain_state_semantics():
    """reference tests/test_states.py: construct / attribute access / copy / read-only."""
    pos, mom = torch.zeros(4, 3, dtype=torch.float64), torch.ones(4, 3, dtype=torch.float64)
    s = mb.ChainState(pos=pos, mom=mom, dir=1)
    assert "pos" in s and "mom" in s and "dir" in s and "foo" not in s
    assert s.n_chains == 4 and s.dim == 3
    c = s.copy()
    c.pos[0, 0] = 5.0
    assert s.pos[0, 0] == 0.0  # deep copy of variables (states.py:263-279)
    r = s.copy(read_only=True)
    with pytest.raises(ReadOnlyStateError):
        r.pos = pos
    with pytest.raises(AttributeError):
        _ = s.nonexistent
    with pytest.raises(ValueError):
        mb.ChainState(_bad=1, pos=pos)
    s.status = torch.zeros(4, dtype=torch.int32)
    s.mom = mom * 2  # setting a variable drops derived per-call outputs
    assert s.status is None
    s2 = pickle.loads(pickle.dumps(s))
    assert torch.equal(s2.mom, s.mom)


def test_error_hierarchy_and_status_mapping():
    assert issubclass(ConvergenceError, IntegratorError)
    assert issubclass(NonReversibleStepError, IntegratorError)
    assert not issubclass(LinAlgError, IntegratorError)
    raise_for_status(0)
    with pytest.raises(ConvergenceError):
        raise_for_status(1)
    with pytest.raises(NonReversibleStepError):
        raise_for_status(2)
    with pytest.raises(LinAlgError):
        raise_for_status(3)


def test_metric_coercion_matches_reference_rules():
    """systems.py:332-346: None -> identity, 1-D -> diagonal, 2-D -> dense, else ValueError."""
    t = targets.StdGaussian(4)
    assert systems.EuclideanMetricSystem(t).metric.kind == systems.METRIC_IDENTITY
    d = systems.EuclideanMetricSystem(t, metric=np.array([1.0, 2.0, 3.0, 4.0]))
    assert d.metric.kind == systems.METRIC_DIAGONAL
    np.testing.assert_allclose(d.metric.inv, [1.0, 0.5, 1 / 3, 0.25])
    rng = np.random.default_rng(0)
    m = problems.dense_spd_metric(rng, 4)
    e = systems.EuclideanMetricSystem(t, metric=m)
    assert e.metric.kind == systems.METRIC_DENSE
    np.testing.assert_allclose(e.metric.inv @ m, np.identity(4), atol=1e-13)
    np.testing.assert_allclose(e.metric.sqrt @ e.metric.sqrt.T, m, atol=1e-13)
    with pytest.raises(ValueError):
        systems.EuclideanMetricSystem(t, metric=np.zeros((2, 2, 2)))
    with pytest.raises(ValueError):
        systems.EuclideanMetricSystem(t, metric=np.array([1.0, -1.0, 1.0, 1.0]))
    with pytest.raises(LinAlgError):
        systems.EuclideanMetricSystem(t, metric=-np.identity(4))
    e.metric = None  # assignable, as adapters do (adapters.py:513, 642)
    assert e.metric.kind == systems.METRIC_IDENTITY


def test_explicit_inverse_is_built_like_the_reference():
    from oracle import mici_oracle as mo

    rng = np.random.default_rng(1)
    m = problems.dense_spd_metric(rng, 24)
    _, inv = systems._explicit_spd_inverse(m)
    np.testing.assert_array_equal(inv, mo.DenseMetric(m).inv_array)


def test_system_rejects_python_callables_and_foreign_derivatives():
    with pytest.raises(TypeError):
        systems.EuclideanMetricSystem(lambda q: 0.0)
    with pytest.raises(ValueError):
        systems.EuclideanMetricSystem(targets.StdGaussian(3), grad_neg_log_dens=lambda q: q)
    with pytest.raises(ValueError):
        systems.DenseConstrainedEuclideanMetricSystem(targets.StdGaussian(3))
    # densities with respect to the Lebesgue measure are supported (systems.py:853-861)
    systems.DenseConstrainedEuclideanMetricSystem(targets.Torus(), dens_wrt_hausdorff=False)
    with pytest.raises(ValueError):
        systems.SoftAbsRiemannianMetricSystem(targets.Banana(4), softabs_coeff=0.0)


def test_call_counter_api_defaults():
    """``Integrator.count_calls`` (kernel-side counters): off by default, zero totals, chainable."""
    from mici_b200 import integrators

    eu = systems.EuclideanMetricSystem(targets.NealFunnel(8))
    integ = integrators.LeapfrogIntegrator(eu, 0.1)
    assert integ.call_counts is None and not integ._counting
    assert integ.call_count_totals() == dict.fromkeys(integ.COUNTER_NAMES, 0)
    assert integ.count_calls() is integ and integ._counting
    assert integ.count_calls(False).call_counts is None


def test_integrator_constructor_contracts():
    """integrators.py:52-80, 121-130, 438-446, 855-864: names, defaults, rejections."""
    eu = systems.EuclideanMetricSystem(targets.NealFunnel(8))
    lf = integrators.LeapfrogIntegrator(eu)
    assert lf.step_size is None and lf.system is eu
    s = mb.ChainState(pos=torch.zeros(2, 8, dtype=torch.float64),
                      mom=torch.zeros(2, 8, dtype=torch.float64), dir=1)
    with pytest.raises(AdaptationError):
        lf.step(s)  # step_size None (integrators.py:72-77)
    lf.step_size = 0.25  # adapters assign it (adapters.py:322-340)
    rm = systems.SoftAbsRiemannianMetricSystem(targets.Banana(4))
    with pytest.raises(ValueError):
        integrators.LeapfrogIntegrator(rm)  # no h1_flow / h2_flow (integrators.py:121-130)
    il = integrators.ImplicitLeapfrogIntegrator(rm, 0.1)
    assert il.reverse_check_tol == 2e-8 and il.fixed_point_solver is solvers.solve_fixed_point_direct
    assert il.fixed_point_solver.resolve_kwargs({"convergence_tol": 1e-12}) == {
        "convergence_tol": 1e-12, "divergence_tol": 1e10, "max_iters": 100}
    with pytest.raises(TypeError):
        il.fixed_point_solver.resolve_kwargs({"bogus": 1})
    with pytest.raises(ValueError):
        integrators.ImplicitLeapfrogIntegrator(rm, 0.1, fixed_point_solver=lambda f, x: x)
    st = integrators.ImplicitLeapfrogIntegrator(
        rm, 0.1, fixed_point_solver=solvers.solve_fixed_point_steffensen)
    assert st.fixed_point_solver.kind == 1
    cs = systems.DenseConstrainedEuclideanMetricSystem(targets.Torus())
    cl = integrators.ConstrainedLeapfrogIntegrator(cs, 0.1, n_inner_step=3)
    assert cl.n_inner_step == 3
    assert cl.projection_solver.defaults == {
        "constraint_tol": 1e-9, "position_tol": 1e-8, "divergence_tol": 1e10, "max_iters": 50}
    with pytest.raises(TypeError):
        integrators.ConstrainedLeapfrogIntegrator(eu, 0.1)


def test_integrators_survive_deepcopy_and_pickle():
    """samplers.py:1124-1129 deep-copies the integrator per chain; multiprocess pickles it."""
    for cfg, kw in (("C1", {"n_chains": 4, "dim": 8}), ("C2", {"n_chains": 4, "dim": 4}),
                    ("C3", {"n_chains": 4}), ("C4", {"n_chains": 2, "dim": 8})):
        integ = engine.build_integrator(problems.make_problem(cfg, **kw))
        for clone in (copy.deepcopy(integ), pickle.loads(pickle.dumps(integ))):
            assert type(clone) is type(integ)
            assert clone.step_size == integ.step_size
            assert clone.system.target.target_id == integ.system.target.target_id


def test_solver_markers_are_not_host_callables():
    from mici_b200.errors import Error

    with pytest.raises(Error):
        solvers.solve_fixed_point_direct(np.cos, np.array([1.0]))
    x = torch.tensor([[1.0, -3.0], [0.5, 0.25]])
    assert torch.equal(solvers.maximum_norm(x), torch.tensor([3.0, 0.5]))


def test_problem_configs_match_baseline_shapes():
    shapes = {"C0": (4, 10), "C1": (8192, 128), "C2": (2048, 64), "C3": (4096, 3)}
    for name, shp in shapes.items():
        p = problems.make_problem(name)
        assert p.pos.shape == shp and p.mom.shape == shp and p.pos.dtype == np.float64
        assert p.algorithmic_bytes_per_chain_step == 32 * shp[1]
    p = problems.make_problem("C4", n_chains=8)
    assert p.dim == 512
    # seeded: identical on every call
    a, b = problems.make_problem("C1", n_chains=8), problems.make_problem("C1", n_chains=8)
    np.testing.assert_array_equal(a.pos, b.pos)
    np.testing.assert_array_equal(a.metric, b.metric)


def test_compute_call_without_gpu_fails_loudly():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    prob = problems.make_problem("C1", n_chains=4, dim=8)
    integ = engine.build_integrator(prob)
    state = engine.build_state(prob, "cpu")
    with pytest.raises(Exception):  # noqa: B017 - no CPU fallback exists
        integ.step(state)


def test_stagers_match_reference_stage_lists():
    """stagers.py:79-291: the stage schedule stored with the adaptation fixtures (generated by
    the reference's stagers) and a sweep of closed-form properties."""
    from golden_util import load_adapt_case
    from mici_b200 import stagers
    from oracle.make_golden import ADAPT_CASES

    class Flag:
        def __init__(self, fast):
            self.is_fast = fast

    for name in ADAPT_CASES:
        _, specs, sk, n_warm, n_main, _, _, stages, _ = load_adapt_case(name)
        flags = [Flag(n == "dual_averaging") for n, _ in specs]
        if sk is not None:
            stager = stagers.WindowedWarmUpStager(**sk)
        elif all(f.is_fast for f in flags):
            stager = stagers.WarmUpStager()
        else:
            stager = stagers.WindowedWarmUpStager()
        mine = stager.stages(n_warm, n_main, flags)
        got = [(st.n_iter, None if st.adapters is None else
                ("all" if len(st.adapters) == len(flags) else "fast")) for st in mine.values()]
        if all(f.is_fast for f in flags):  # a fast-only list is both "all" and "fast"
            got = [(n, w if w is None else "all") for n, w in got]
            stages = [(n, w if w is None else "all") for n, w in stages]
        assert got == stages, name
    for n_warm in (1, 9, 100, 150, 151, 1000, 2777):
        n0, windows, n1 = stagers.WindowedWarmUpStager().slow_windows(n_warm)
        assert n0 + sum(windows) + n1 == n_warm and all(w > 0 for w in windows)


def test_adapter_reducers_and_moment_merge():
    """adapters.py:126-159 reducers; the Chan / Schubert-Gertz merge against a direct estimate."""
    import math

    import torch

    from mici_b200 import adapters

    logs = [math.log(0.1), math.log(0.2), math.log(0.4)]
    assert adapters.arithmetic_mean_log_step_size_reducer(logs) == pytest.approx(0.7 / 3, rel=1e-14)
    assert adapters.geometric_mean_log_step_size_reducer(logs) == pytest.approx(0.2, rel=1e-14)
    assert adapters.min_log_step_size_reducer(logs) == pytest.approx(0.1, rel=1e-14)
    rng = np.random.default_rng(0)
    x = torch.as_tensor(rng.standard_normal((3, 40, 5)))  # three "ranks" of 40 samples
    parts = []
    for r in range(3):
        mean = x[r].mean(0)
        parts.append((40, mean, (x[r] - mean).T @ (x[r] - mean)))
    n, mean, m2 = adapters._merge_moments(parts, outer=True)
    flat = x.reshape(-1, 5)
    assert n == 120
    np.testing.assert_allclose(mean.numpy(), flat.mean(0).numpy(), rtol=1e-13)
    np.testing.assert_allclose((m2 / (n - 1)).numpy(), np.cov(flat.numpy().T), rtol=1e-12, atol=1e-14)
    parts_v = [(c, m, torch.diagonal(s).clone()) for c, m, s in parts]
    _, _, v2 = adapters._merge_moments(parts_v, outer=False)
    np.testing.assert_allclose((v2 / (n - 1)).numpy(), flat.var(0, unbiased=True).numpy(), rtol=1e-12)


def test_covariance_factored_metric_matches_reference_semantics():
    """``DensePositiveDefiniteMatrix(covar).inv`` as assigned by the covariance adapter
    (adapters.py:642): array = explicit inverse, inv = L L^T, sqrt = L^-T."""
    import scipy.linalg as sla

    from mici_b200.systems import _FixedMetric

    rng = np.random.default_rng(1)
    a = rng.standard_normal((7, 7))
    covar = a @ a.T + 7 * np.eye(7)
    m = _FixedMetric.from_covariance(covar)
    chol = np.linalg.cholesky(covar)
    np.testing.assert_allclose(m.inv, covar, rtol=1e-13)
    np.testing.assert_allclose(m.array @ covar, np.eye(7), atol=1e-12)
    z = rng.standard_normal(7)
    np.testing.assert_allclose(m.sqrt @ z, sla.solve_triangular(chol.T, z, lower=False), rtol=1e-12)
    np.testing.assert_allclose(m.sqrt @ m.sqrt.T, m.array, rtol=1e-11, atol=1e-14)


@pytest.mark.parametrize("args", [(), (125, 50, 25, 3)])
@pytest.mark.parametrize("n_warm_up_iter", [5, 10, 100, 107, 500, 1003])
def test_windowed_stager_iteration_budget(args, n_warm_up_iter):
    """Mirror of the reference's ``StagerTests.test_stages`` (reference tests/test_stagers.py:8-25)."""
    from mici_b200 import stagers

    class Flag:
        def __init__(self, fast):
            self.is_fast = fast

    stages = stagers.WindowedWarmUpStager(*args).stages(n_warm_up_iter, 1, [Flag(True), Flag(False)])
    assert all(isinstance(k, str) and isinstance(v, stagers.ChainStage) for k, v in stages.items())
    assert all(v.n_iter >= 0 for v in stages.values())
    assert sum(v.n_iter for v in stages.values()) == n_warm_up_iter + 1
    stages = stagers.WarmUpStager().stages(1, 1, [Flag(True)])
    assert sum(v.n_iter for v in stages.values()) == 2


def test_transition_constructors_validate_like_reference():
    """transitions.py:338-341, 388-392, 507-509 and the fused-only restrictions."""
    from mici_b200 import integrators, systems, targets, transitions

    system = systems.EuclideanMetricSystem(targets.StdGaussian(4))
    integ = integrators.LeapfrogIntegrator(system, 0.1)
    with pytest.raises(ValueError):
        transitions.MetropolisStaticIntegrationTransition(system, integ, 0)
    with pytest.raises(ValueError):
        transitions.MetropolisRandomIntegrationTransition(system, integ, (3, 3))
    with pytest.raises(ValueError):
        transitions.MultinomialDynamicIntegrationTransition(system, integ, max_tree_depth=0)
    with pytest.raises(TypeError):
        transitions.DynamicIntegrationTransition(system, integ)
    # other integrators take the generic lock-step tree builder (nuts_generic.cuh)
    tr2 = transitions.SliceDynamicIntegrationTransition(
        system, integrators.BCSSTwoStageIntegrator(system, 0.1))
    assert not tr2._fused
    tr = transitions.SliceDynamicIntegrationTransition(system, integ, max_tree_depth=5)
    assert tr.n_uniforms == 2 * 5 + 2**5 + 1
    # systems the fused tree builder does not cover run leaf by leaf through their integrator
    torus = targets.make_target("torus")
    csys = systems.DenseConstrainedEuclideanMetricSystem(torus, torus)
    tr3 = transitions.MultinomialDynamicIntegrationTransition(
        csys, integrators.ConstrainedLeapfrogIntegrator(csys, 0.1))
    gsys = systems.GaussianEuclideanMetricSystem(targets.StdGaussian(4))
    tr4 = transitions.MultinomialDynamicIntegrationTransition(
        gsys, integrators.LeapfrogIntegrator(gsys, 0.1))
    assert not tr3._fused and not tr4._fused and tr._fused


def test_sampler_front_ends_mirror_reference_signatures_and_defaults():
    """samplers.py:1458-1465, 1527-1534, 1600-1611, 1708-1719: constructor signatures and the
    (different) defaults of the two dynamic samplers; per-chain generators as samplers.py:559-560."""
    import numpy as np

    from mici_b200 import integrators, samplers, systems, targets, transitions

    system = systems.EuclideanMetricSystem(targets.StdGaussian(4))
    integ = integrators.LeapfrogIntegrator(system, 0.1)
    rng = np.random.default_rng(1)
    s1 = samplers.StaticMetropolisHMC(system, integ, rng, 7)
    assert s1.n_step == 7 and s1.system is system and s1.rng is rng
    assert list(s1.transitions) == ["momentum_transition", "integration_transition"]
    s2 = samplers.RandomMetropolisHMC(system, integ, rng, (2, 9))
    assert s2.n_step_range == (2, 9)
    m = samplers.DynamicMultinomialHMC(system, integ, rng).transitions["integration_transition"]
    sl = samplers.DynamicSliceHMC(system, integ, rng).transitions["integration_transition"]
    assert m.termination_criterion is transitions.riemannian_no_u_turn_criterion
    assert m.do_extra_subtree_checks and m.max_tree_depth == 10 and m.max_delta_h == 1000
    assert sl.termination_criterion is transitions.euclidean_no_u_turn_criterion
    assert not sl.do_extra_subtree_checks
    gens = samplers._per_chain_rngs(rng, 3)
    want = [np.random.default_rng(rng.bit_generator.jumped(i)).uniform() for i in range(3)]
    assert [g.uniform() for g in gens] == want
    with pytest.raises(TypeError):
        s1.sample_chains(1, 1, np.zeros((2, 4)), not_an_argument=1)


class _OracleBackedIntegrator:
    """CPU stand-in for a device integrator: steps every chain through the oracle (NumPy) and
    reports status / Hamiltonian like ``Integrator.step_n`` -- lets the batched host logic of the
    adapters be checked on a machine without a GPU."""

    def __init__(self, problem):
        from oracle import drivers as dr

        self.ctx = dr._AdaptiveContext(problem)
        self.step_size = None

    def step_n(self, state, n_steps, *, return_h=False):
        import torch

        from mici_b200.states import ChainState
        from oracle import mici_oracle as mo

        assert n_steps == 1 and return_h
        q, p = state.pos.numpy(), state.mom.numpy()
        eps = self.step_size.numpy()
        n = q.shape[0]
        out_q, out_p = q.copy(), p.copy()
        status, h = np.zeros(n, dtype=np.int32), np.full(n, np.nan)
        for c in range(n):
            try:
                out_q[c], out_p[c] = self.ctx.step_eps(q[c], p[c], 1, float(eps[c]))
                h[c] = self.ctx.h(out_q[c], out_p[c])
            except mo.OracleIntegratorError as e:
                status[c] = e.status
        new = ChainState(pos=torch.as_tensor(out_q), mom=torch.as_tensor(out_p), dir=1)
        new.status, new.h = torch.as_tensor(status), torch.as_tensor(h)
        return new


class _OracleBackedSystem:
    def __init__(self, integrator):
        self.ctx = integrator.ctx

    def h(self, state):
        import torch

        q, p = state.pos.numpy(), state.mom.numpy()
        return torch.as_tensor(np.array([self.ctx.h(q[c], p[c]) for c in range(q.shape[0])]))


@pytest.mark.parametrize("cfg,kwargs", [("C1", {"n_chains": 24, "dim": 12}), ("C3", {"n_chains": 24}),
                                        ("C0", {"n_chains": 6, "dim": 10})])
def test_batched_initial_step_size_search_matches_oracle_on_cpu(cfg, kwargs):
    """adapters.py:285-352 as batched tensor logic (per-chain halving / doubling, NaN energies
    at step size 1 for the funnel, failed constrained steps for the torus) against the per-chain
    oracle search -- with an oracle-backed integrator, so the comparison is exact."""
    import warnings

    import torch

    from mici_b200 import adapters, problems
    from mici_b200.states import ChainState
    from oracle import mici_oracle as mo

    problem = problems.make_problem(cfg, **kwargs)
    integ = _OracleBackedIntegrator(problem)
    system = _OracleBackedSystem(integ)
    state = ChainState(pos=torch.as_tensor(problem.pos), mom=torch.as_tensor(problem.mom), dir=1)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        got = adapters.DualAveragingStepSizeAdapter()._find_and_set_init_step_size(
            state, system, integ).numpy()
        want = np.array([mo.find_init_step_size(problem.pos[c], problem.mom[c], 1,
                                                integ.ctx.step_eps, integ.ctx.h)
                         for c in range(problem.n_chains)])
    np.testing.assert_array_equal(got, want)
    assert integ.step_size is not None and torch.equal(integ.step_size, torch.as_tensor(got))
    if cfg != "C0":
        assert len(set(got.tolist())) > 1  # chains really end on different step sizes


def test_dual_averaging_update_matches_oracle_on_cpu():
    """adapters.py:354-373 for a batch against the scalar restatement, 40 updates."""
    import torch

    from mici_b200 import adapters
    from oracle import mici_oracle as mo

    n = 7
    rng = np.random.default_rng(4)
    accept = rng.uniform(0.0, 1.0, (40, n))
    eps0 = 2.0 ** rng.integers(-6, 1, n)
    ad = adapters.DualAveragingStepSizeAdapter(adapt_stat_target=0.7)
    st = {"iter": 0, "smoothed_log_step_size": torch.zeros(n, dtype=torch.float64),
          "adapt_stat_error": torch.zeros(n, dtype=torch.float64),
          "log_step_size_reg_target": torch.log(10 * torch.as_tensor(eps0))}
    tr = type("T", (), {"integrator": type("I", (), {"step_size": None})()})()
    od = mo.DualAveragingOracle(adapt_stat_target=0.7)

    class Ctx:
        step_size = None

    ost = [{"iter": 0, "smoothed_log_step_size": 0.0, "adapt_stat_error": 0.0,
            "log_step_size_reg_target": float(np.log(10 * eps0[c]))} for c in range(n)]
    ctxs = [Ctx() for _ in range(n)]
    for it in range(40):
        ad.update(st, None, {"accept_stat": torch.as_tensor(accept[it])}, tr)
        for c in range(n):
            od.update(ost[c], None, {"accept_stat": accept[it, c]}, ctxs[c])
        np.testing.assert_allclose(tr.integrator.step_size.numpy(),
                                   [ctxs[c].step_size for c in range(n)], rtol=1e-13)
    np.testing.assert_allclose(st["smoothed_log_step_size"].numpy(),
                               [o["smoothed_log_step_size"] for o in ost], rtol=1e-13, atol=1e-15)
    ad.finalize(st, None, tr, None)
    od.finalize(ost, ctxs[0])
    assert tr.integrator.step_size == pytest.approx(ctxs[0].step_size, rel=1e-13)


def test_sampler_chain_data_files_use_reference_names(tmp_path):
    """samplers.py:104-113, 247-254, 278-289: per-chain memmap file names and contents."""
    import torch

    from mici_b200 import samplers

    traces = {"pos": torch.arange(2 * 5 * 3, dtype=torch.float64).reshape(2, 5, 3),
              "hamiltonian": torch.arange(10, dtype=torch.float64).reshape(2, 5)}
    stats = {"accept_stat": torch.rand(2, 5, dtype=torch.float64),
             "n_step": torch.arange(10).reshape(2, 5)}
    samplers.write_chain_data(tmp_path, traces, stats, first_chain_index=4)
    names = sorted(p.name for p in tmp_path.iterdir())
    assert names == sorted([
        "trace_4_pos.npy", "trace_5_pos.npy", "trace_4_hamiltonian.npy", "trace_5_hamiltonian.npy",
        "stats_4_integration_transition_accept_stat.npy",
        "stats_5_integration_transition_accept_stat.npy",
        "stats_4_integration_transition_n_step.npy", "stats_5_integration_transition_n_step.npy"])
    got = np.load(tmp_path / "trace_5_pos.npy", mmap_mode="r")
    np.testing.assert_array_equal(got, traces["pos"][1].numpy())
    got = np.load(tmp_path / "stats_4_integration_transition_n_step.npy")
    np.testing.assert_array_equal(got, stats["n_step"][0].numpy())
