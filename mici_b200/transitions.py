"""Batched Markov transitions around the integrators -- "next" row N1 of SURVEY.md 8(f).

Mirrors, for all chains at once, the two reference transitions that make up static HMC:

* ``IndependentMomentumTransition``           transitions.py:129-142
* ``MetropolisStaticIntegrationTransition``   transitions.py:256-352

so that a whole HMC iteration (momentum refresh, ``n_step`` fused integrator steps, energy,
accept / reject, direction flips) stays on the GPU: two kernel launches for the trajectory
(``system.h`` of the current state, then ``integrator.step_n(..., return_h=True)``) and one for
the Metropolis select (``mb200_metropolis_select``).

Random numbers.  ``rng`` is either a ``numpy.random.Generator`` (one stream for the whole batch),
a sequence of per-chain generators (each chain then consumes exactly the variates the reference
consumes from its own stream: ``standard_normal(dim)`` then one ``uniform()``, which is what the
parity tests use), or a ``torch.Generator`` on the device (variates generated on the GPU).
"""

from __future__ import annotations

import ctypes
from collections.abc import Sequence

import numpy as np
import torch

from . import _lib
from .states import ChainState
from .systems import _dir_tensor


def _normals(rng, shape, device):
    if isinstance(rng, torch.Generator):
        return torch.randn(shape, dtype=torch.float64, device=device, generator=rng)
    if isinstance(rng, Sequence):
        z = np.stack([g.standard_normal(shape[1:]) for g in rng])
    else:
        z = rng.standard_normal(shape)
    return torch.as_tensor(z, device=device)


def _uniforms(rng, n, device, mask=None):
    if isinstance(rng, torch.Generator):
        return torch.rand(n, dtype=torch.float64, device=device, generator=rng)
    if isinstance(rng, Sequence):
        # the reference draws `rng.uniform()` only for chains whose trajectory did not fail
        # (short-circuit in transitions.py:310); keep the per-chain streams in step with it
        u = np.array([g.uniform() if (mask is None or mask[i]) else 2.0 for i, g in enumerate(rng)])
    else:
        u = rng.uniform(size=n)
    return torch.as_tensor(u, device=device)


class IndependentMomentumTransition:
    """Resample every chain's momentum from N(0, M) (transitions.py:129-142)."""

    state_variables = frozenset({"mom"})
    statistic_types = None

    def __init__(self, system):
        self.system = system

    def sample(self, state, rng):
        state.mom = self.system.sample_momentum(state, rng)
        return state, None


def _integers(rng, lo, hi, n, device):
    """Per-chain ``rng.integers(lo, hi)`` (transitions.py:401)."""
    if isinstance(rng, torch.Generator):
        return torch.randint(lo, hi, (n,), dtype=torch.int32, device=device, generator=rng)
    if isinstance(rng, Sequence):
        k = np.array([g.integers(lo, hi) for g in rng], dtype=np.int32)
    else:
        k = rng.integers(lo, hi, size=n).astype(np.int32)
    return torch.as_tensor(k, device=device)


class MetropolisIntegrationTransition:
    """Trajectory + Metropolis accept step for all chains (transitions.py:235-315).  ``sample``
    returns ``(state, stats)`` where ``stats`` holds per-chain tensors with the reference's
    statistic names (transitions.py:226-232, 273)."""

    state_variables = frozenset({"pos", "mom", "dir"})

    def __init__(self, system, integrator):
        self.system = system
        self.integrator = integrator

    def _sample_n_step(self, state, n_step, rng):
        """``n_step``: one integer for all chains or an integer tensor ``[n_chains]``."""
        n, dim = state.pos.shape
        dev = state.pos.device
        h_init = self.system.h(state)
        prop = self.integrator.step_n(state, n_step, return_h=True)
        status, n_done = prop.status, prop.n_done
        dirs = _dir_tensor(state.dir if "dir" in state else 1, n, dev)
        if dirs is None:
            dirs = torch.ones(n, dtype=torch.int32, device=dev)
        else:
            dirs = dirs.clone()
        mask = None
        if isinstance(rng, Sequence):
            mask = (status == 0).cpu().numpy()
        u = _uniforms(rng, n, dev, mask)
        pos, mom = state.pos.clone(), state.mom.clone()
        accept_prob = torch.empty(n, dtype=torch.float64, device=dev)
        accept_stat = torch.empty(n, dtype=torch.float64, device=dev)
        accepted = torch.empty(n, dtype=torch.int32, device=dev)
        rc = _lib.load().mb200_metropolis_select(
            _lib.ptr(pos), _lib.ptr(mom), _lib.ptr(prop.pos), _lib.ptr(prop.mom),
            _lib.ptr(h_init), _lib.ptr(prop.h), _lib.ptr(status), _lib.ptr(n_done),
            _lib.ptr(dirs), _lib.ptr(u), n, dim, _lib.ptr(accept_prob), _lib.ptr(accept_stat),
            _lib.ptr(accepted), _lib.current_stream_ptr(dev),
        )
        _lib.check(rc, "mb200_metropolis_select")
        new = ChainState(pos=pos, mom=mom, dir=dirs)
        stats = {
            "n_step": n_done.to(torch.int64),
            "accept_stat": accept_stat,
            "metrop_accept_prob": accept_prob,
            "convergence_error": status == 1,
            "non_reversible_step": status == 2,
            "step_size": _step_size_stat(self.integrator.step_size, n, dev),
            "accepted": accepted.bool(),
        }
        return new, stats


def _step_size_stat(step_size, n, dev):
    if isinstance(step_size, torch.Tensor) and step_size.ndim == 1:
        return step_size.to(device=dev, dtype=torch.float64).clone()
    return torch.full((n,), float(step_size), dtype=torch.float64, device=dev)


class MetropolisStaticIntegrationTransition(MetropolisIntegrationTransition):
    """Static-trajectory HMC transition (transitions.py:318-352)."""

    def __init__(self, system, integrator, n_step):
        super().__init__(system, integrator)
        if n_step <= 0:
            raise ValueError("Number of integrator steps must be positive.")
        self.n_step = int(n_step)

    def sample(self, state, rng):
        return self._sample_n_step(state, self.n_step, rng)


class MetropolisRandomIntegrationTransition(MetropolisIntegrationTransition):
    """Trajectory length drawn per chain and per transition from ``rng.integers(lower, upper)``
    (transitions.py:355-402; NumPy's ``integers`` excludes ``upper``); all chains still advance
    in one launch (per-chain ``n_steps``, ``mb200_leapfrog_euclidean_per_chain``)."""

    def __init__(self, system, integrator, n_step_range):
        super().__init__(system, integrator)
        n_step_lower, n_step_upper = n_step_range
        if not (n_step_lower > 0 and n_step_lower < n_step_upper):
            raise ValueError("Range bounds must be non-negative and first entry less than last.")
        self.n_step_range = (int(n_step_lower), int(n_step_upper))

    def sample(self, state, rng):
        n = state.pos.shape[0]
        n_step = _integers(rng, *self.n_step_range, n, state.pos.device)
        return self._sample_n_step(state, n_step, rng)


def _run_stage(mom_tr, int_tr, state, rng, n_iter, adapters, record, all_stats, trace, group):
    """One sampling stage for every chain: the loop body of ``_sample_chain`` (samplers.py:459-513)
    then ``_finalize_adapters`` (samplers.py:1131-1138)."""
    adapters = adapters or []
    adapt_states = [a.initialize(state, int_tr) for a in adapters]
    for _ in range(n_iter):
        state, _ = mom_tr.sample(state, rng)
        state, st = int_tr.sample(state, rng)
        for a, a_state in zip(adapters, adapt_states):
            a.update(a_state, state, st, int_tr)
        if record:
            for k, v in st.items():
                all_stats.setdefault(k, []).append(v)
            if trace is not None:
                trace.append(state.pos.clone())
    for a, a_state in zip(adapters, adapt_states):
        a.finalize(a_state, state, int_tr, rng, group=group)
    return state, adapt_states


def sample_hmc(system, integrator, state, rng, n_iter, n_step, trace_pos=False, adapters=None,
               group=None):
    """``n_iter`` static-HMC iterations (momentum refresh + Metropolis transition) for every
    chain of ``state`` -- the inner loop of ``samplers._sample_chain`` (samplers.py:479-513) with
    the chain axis on the device; ``adapters`` (``mici_b200.adapters``) are initialised before,
    updated after every transition and finalised after the last one, as in one reference
    sampling stage.  Returns ``(final_state, stats, traces)``: ``stats`` per-key tensors
    ``[n_iter, n_chains]``, ``traces`` the positions ``[n_iter, n_chains, dim]`` if requested."""
    mom_tr = IndependentMomentumTransition(system)
    int_tr = MetropolisStaticIntegrationTransition(system, integrator, n_step)
    all_stats, trace = {}, ([] if trace_pos else None)
    state, _ = _run_stage(mom_tr, int_tr, state, rng, n_iter, adapters, True, all_stats, trace,
                          group)
    stats = {k: torch.stack(v) for k, v in all_stats.items()}
    return state, stats, (torch.stack(trace) if trace_pos else None)


def sample_chains(system, integrator, state, rng, n_warm_up_iter, n_main_iter, *, n_step=None,
                  n_step_range=None, adapters=None, stager=None, trace_warm_up=False,
                  trace_pos=True, group=None):
    """Staged sampling of all chains: ``HamiltonianMonteCarlo.sample_chains``
    (samplers.py:875-1141) for the static (``n_step``) or random (``n_step_range``) Metropolis
    HMC transitions, with the stage schedule of ``mici_b200.stagers`` (default: one warm-up stage
    if all adapters are fast, else windowed: samplers.py:1075-1082).  Adapter states are
    re-initialised at the start of every stage and finalised at its end (across all chains and,
    with a process group, across all ranks).  Returns ``(final_state, stats, traces)`` over the
    recorded stages (the main stage; also the warm-up if ``trace_warm_up``)."""
    from .stagers import WarmUpStager, WindowedWarmUpStager  # noqa: PLC0415

    if (n_step is None) == (n_step_range is None):
        raise ValueError("Give exactly one of `n_step` and `n_step_range`.")
    adapters = list(adapters or [])
    mom_tr = IndependentMomentumTransition(system)
    if n_step is not None:
        int_tr = MetropolisStaticIntegrationTransition(system, integrator, n_step)
    else:
        int_tr = MetropolisRandomIntegrationTransition(system, integrator, n_step_range)
    if stager is None:
        stager = WarmUpStager() if all(a.is_fast for a in adapters) else WindowedWarmUpStager()
    all_stats, trace = {}, ([] if trace_pos else None)
    for stage in stager.stages(n_warm_up_iter, n_main_iter, adapters,
                               trace_warm_up=trace_warm_up).values():
        state, _ = _run_stage(mom_tr, int_tr, state, rng, stage.n_iter, stage.adapters,
                              stage.record_stats, all_stats, trace if stage.trace else None,
                              group)
    stats = {k: torch.stack(v) for k, v in all_stats.items()}
    return state, stats, (torch.stack(trace) if trace_pos and trace else None)
