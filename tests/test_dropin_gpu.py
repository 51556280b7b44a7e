"""Drop-in proof: the STOCK reference sampler / transition layer
(``mici.samplers``, ``mici.transitions`` -- unmodified, imported from ``/root/reference/src`` or
from the copy under ``oracle/_ref``) drives a ``mici_b200`` system + integrator, one NumPy-held
``mici.states.ChainState`` per chain, exactly as it drives its own; the chains it produces are
compared with the all-reference run on the same seeds.

What the reference layer touches on the replaced objects (reference file:line):
``integrator.step(state)`` transitions.py:291, 657; ``system.h(state)`` transitions.py:281, 301;
``system.sample_momentum(state, rng)`` transitions.py:141; ``system.dh_dmom(state)``
transitions.py:434-435, 472-473 (dynamic criteria); ``state.copy()`` / ``state.dir`` flips.
"""

import numpy as np
import pytest

from mici_b200 import engine, problems
from oracle import drivers as dr

pytestmark = pytest.mark.gpu

needs_reference = pytest.mark.skipif(
    not dr.reference_available(), reason="reference package not available (oracle/_ref missing)"
)


def _run_stock_sampler(mici, sampler_cls, system, integrator, problem, n_iter, seed, **kw):
    rng = np.random.default_rng(seed)
    sampler = sampler_cls(system, integrator, rng, **kw)
    init = [mici.states.ChainState(pos=problem.pos[i].copy(), mom=None, dir=1)
            for i in range(problem.n_chains)]
    final, traces, stats = sampler.sample_chains(
        0, n_iter, init, adapters=[], n_worker=1, display_progress=False,
        trace_funcs=[lambda state: {"pos": state.pos}],
    )
    return (np.stack([np.asarray(s.pos) for s in final]), np.asarray(traces["pos"]),
            {k: np.asarray(v) for k, v in stats.items()})


CASES = {
    # name: (config, kwargs, sampler, n_iter, sampler kwargs)
    "c1_static": ("C1", {"n_chains": 4, "dim": 16}, "StaticMetropolisHMC", 6, {"n_step": 5}),
    "c0_static": ("C0", {"n_chains": 4, "dim": 10}, "StaticMetropolisHMC", 6, {"n_step": 7}),
    "c3_static": ("C3", {"n_chains": 4}, "StaticMetropolisHMC", 5, {"n_step": 4}),
    "c2_static": ("C2", {"n_chains": 3, "dim": 8}, "StaticMetropolisHMC", 3, {"n_step": 3}),
    "c4_static": ("C4", {"n_chains": 3, "dim": 12}, "StaticMetropolisHMC", 3, {"n_step": 3}),
    "c1_dynamic": ("C1", {"n_chains": 4, "dim": 10}, "DynamicMultinomialHMC", 4,
                   {"max_tree_depth": 4}),
    "c2_dynamic": ("C2", {"n_chains": 2, "dim": 8}, "DynamicMultinomialHMC", 2,
                   {"max_tree_depth": 3}),
}


@needs_reference
@pytest.mark.parametrize("name", sorted(CASES))
def test_stock_mici_sampler_over_mici_b200_integrator(name):
    cfg, kwargs, sampler_name, n_iter, skw = CASES[name]
    mici = dr.import_reference()
    problem = problems.make_problem(cfg, **kwargs)
    sampler_cls = getattr(mici.samplers, sampler_name)
    seed = 4242

    # all-reference run
    ref_system, ref_integrator = dr.build_reference(problem)
    ref_final, ref_trace, ref_stats = _run_stock_sampler(
        mici, sampler_cls, ref_system, ref_integrator, problem, n_iter, seed, **skw)

    # the same stock sampler over the CUDA system + integrator
    integrator = engine.build_integrator(problem)
    new_final, new_trace, new_stats = _run_stock_sampler(
        mici, sampler_cls, integrator.system, integrator, problem, n_iter, seed, **skw)

    np.testing.assert_array_equal(new_stats["n_step"], ref_stats["n_step"])
    np.testing.assert_array_equal(new_stats["convergence_error"], ref_stats["convergence_error"])
    np.testing.assert_array_equal(new_stats["non_reversible_step"],
                                  ref_stats["non_reversible_step"])
    np.testing.assert_allclose(new_stats["accept_stat"], ref_stats["accept_stat"],
                               rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(new_trace, ref_trace, rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(new_final, ref_final, rtol=1e-8, atol=1e-10)


@needs_reference
def test_integrator_step_raises_reference_compatible_errors():
    """``integrator.step`` on a single NumPy chain raises ``ConvergenceError`` /
    ``NonReversibleStepError`` that the reference's ``except IntegratorError`` clauses catch
    (transitions.py:292, 670)."""
    mici = dr.import_reference()
    problem = problems.make_problem("C3", n_chains=16)
    problem.step_size = 0.6
    integrator = engine.build_integrator(problem)
    raised = 0
    for i in range(problem.n_chains):
        state = mici.states.ChainState(pos=problem.pos[i].copy(), mom=problem.mom[i].copy(), dir=1)
        try:
            for _ in range(3):
                state = integrator.step(state)
        except Exception as e:  # noqa: BLE001
            raised += 1
            assert type(e).__name__ in ("ConvergenceError", "NonReversibleStepError"), repr(e)
    assert raised > 0
