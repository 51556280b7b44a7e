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


class MetropolisStaticIntegrationTransition:
    """Static-trajectory HMC transition with Metropolis accept step for all chains
    (transitions.py:256-352).  ``sample`` returns ``(state, stats)`` where ``stats`` holds
    per-chain tensors with the reference's statistic names (transitions.py:226-232, 273)."""

    state_variables = frozenset({"pos", "mom", "dir"})

    def __init__(self, system, integrator, n_step):
        if n_step <= 0:
            raise ValueError("Number of integrator steps must be positive.")
        self.system = system
        self.integrator = integrator
        self.n_step = int(n_step)

    def sample(self, state, rng):
        n, dim = state.pos.shape
        dev = state.pos.device
        h_init = self.system.h(state)
        prop = self.integrator.step_n(state, self.n_step, return_h=True)
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
            "step_size": torch.full((n,), float(self.integrator.step_size), dtype=torch.float64,
                                    device=dev),
            "accepted": accepted.bool(),
        }
        return new, stats


def sample_hmc(system, integrator, state, rng, n_iter, n_step, trace_pos=False):
    """``n_iter`` static-HMC iterations (momentum refresh + Metropolis transition) for every
    chain of ``state`` -- the inner loop of ``samplers._sample_chain`` (samplers.py:479-513) with
    the chain axis on the device.  Returns ``(final_state, stats, traces)``: ``stats`` per-key
    tensors ``[n_iter, n_chains]``, ``traces`` the positions ``[n_iter, n_chains, dim]`` if
    requested."""
    mom_tr = IndependentMomentumTransition(system)
    int_tr = MetropolisStaticIntegrationTransition(system, integrator, n_step)
    all_stats, trace = {}, []
    for _ in range(n_iter):
        state, _ = mom_tr.sample(state, rng)
        state, st = int_tr.sample(state, rng)
        for k, v in st.items():
            all_stats.setdefault(k, []).append(v)
        if trace_pos:
            trace.append(state.pos.clone())
    stats = {k: torch.stack(v) for k, v in all_stats.items()}
    return state, stats, (torch.stack(trace) if trace_pos else None)
