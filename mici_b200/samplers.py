"""Sampler front ends with the reference's names and call signatures -- host-side glue over
``mici_b200.transitions.sample_chains`` (reference ``src/mici/samplers.py:1144-1700``):

* ``HamiltonianMonteCarlo``   samplers.py:1144-1432
* ``StaticMetropolisHMC``     samplers.py:1434-1498
* ``RandomMetropolisHMC``     samplers.py:1501-1572
* ``DynamicMultinomialHMC``   samplers.py:1575-1683
* ``DynamicSliceHMC``         samplers.py:1686-1796

One sampler object drives ALL chains at once on the device: there are no worker pools,
``n_worker`` and the progress-bar arguments of the reference are accepted and ignored.

``rng`` is a ``numpy.random.Generator`` -- every chain then gets the generator the reference
gives it (``default_rng(rng.bit_generator.jumped(i))``, samplers.py:559-560) and consumes exactly
the variates the reference consumes, so a run replays the reference's chains -- or a device
``torch.Generator`` (all variates generated on the GPU; the production setting).

``sample_chains`` returns ``HMCSampleChainsOutputs(final_states, traces, statistics)`` like the
reference; traces / statistics hold tensors indexed ``[chain, iteration, ...]`` (the reference
returns a list of per-chain arrays: ``traces["pos"][chain][iteration]`` reads the same).
"""

from __future__ import annotations

from typing import NamedTuple

import numpy as np
import torch

from .adapters import DualAveragingStepSizeAdapter
from .states import ChainState
from .transitions import (
    IndependentMomentumTransition,
    MetropolisRandomIntegrationTransition,
    MetropolisStaticIntegrationTransition,
    MultinomialDynamicIntegrationTransition,
    SliceDynamicIntegrationTransition,
    euclidean_no_u_turn_criterion,
    riemannian_no_u_turn_criterion,
    sample_chains,
)


class HMCSampleChainsOutputs(NamedTuple):
    """samplers.py:1170-1198."""

    final_states: ChainState
    traces: dict
    statistics: dict


def write_chain_data(memmap_path, traces, statistics, first_chain_index=0):
    """Write traces / statistics ``[chain, iteration, ...]`` as the per-chain ``.npy`` files the
    reference creates when memory-mapping chain data: ``trace_{i}_{key}.npy`` and
    ``stats_{i}_integration_transition_{key}.npy`` (samplers.py:104-113, 247-254, 278-289), so
    that tooling written against those files keeps working.  ``first_chain_index``: global index
    of this rank's first chain when chains are sharded over ranks."""
    from .traces import write_chain_traces  # noqa: PLC0415

    if not traces and not statistics:
        return {}, {}
    n = next(iter(traces.values())).shape[0] if traces else next(iter(statistics.values())).shape[0]
    indices = range(first_chain_index, first_chain_index + n)
    paths = write_chain_traces(memmap_path, "trace", traces, indices) if traces else {}
    stat_paths = write_chain_traces(
        memmap_path, "stats", {f"integration_transition_{k}": v for k, v in statistics.items()},
        indices)
    return paths, stat_paths


def _per_chain_rngs(rng, n_chain):
    """samplers.py:546-565."""
    if isinstance(rng, torch.Generator):
        return rng
    bit_generator = getattr(rng, "bit_generator", None)
    if bit_generator is not None and hasattr(bit_generator, "jumped"):
        return [np.random.default_rng(bit_generator.jumped(i)) for i in range(n_chain)]
    if bit_generator is not None and hasattr(bit_generator, "seed_seq"):
        return [np.random.default_rng(s) for s in bit_generator.seed_seq.spawn(n_chain)]
    raise ValueError(f"Unsupported random number generator type {type(rng)}.")


class HamiltonianMonteCarlo:
    """Momentum refresh followed by an integration transition, for all chains at once
    (samplers.py:1144-1432)."""

    def __init__(self, system, rng, integration_transition, momentum_transition=None):
        self._system = system
        self._rng = rng
        self.transitions = {
            "momentum_transition": (IndependentMomentumTransition(system)
                                    if momentum_transition is None else momentum_transition),
            "integration_transition": integration_transition,
        }

    @property
    def system(self):
        return self._system

    @property
    def rng(self):
        return self._rng

    def _preprocess_init_state(self, init_states, device):
        """Batched counterpart of samplers.py:1248-1261: an ``[n_chains, dim]`` array / tensor of
        positions or a batched ``ChainState``; missing momenta are drawn from ``self.rng``."""
        if isinstance(init_states, ChainState):
            state = init_states.copy()
        else:
            if isinstance(init_states, (list, tuple)):
                init_states = np.stack([np.asarray(getattr(s, "pos", s)) for s in init_states])
            pos = torch.as_tensor(init_states, dtype=torch.float64).to(device)
            state = ChainState(pos=pos, mom=None, dir=1)
        if state.pos.ndim != 2:
            raise ValueError("init_states must describe a batch: positions [n_chains, dim].")
        if "mom" not in state or state.mom is None:
            n = state.pos.shape[0]
            rng = self._rng
            if not isinstance(rng, torch.Generator):
                # the reference draws the momenta of the chains one after the other from the
                # sampler's own generator (samplers.py:1259-1260)
                rng = [rng] * n
            state.mom = self._system.sample_momentum(state, rng)
        return state

    def sample_chains(self, n_warm_up_iter, n_main_iter, init_states, *, adapters="default",
                      stager=None, trace_warm_up=False, trace_funcs=None, memmap_path=None,
                      first_chain_index=0, device="cuda", group=None, **ignored):
        """samplers.py:1271-1432.  ``adapters`` defaults to one ``DualAveragingStepSizeAdapter``
        (samplers.py:1405-1406); pass ``None`` or ``[]`` for none.  With ``memmap_path`` the
        traces and statistics are also written as the reference's per-chain ``.npy`` files."""
        unknown = set(ignored) - {"n_worker", "n_process", "use_thread_pool", "display_progress",
                                  "progress_bar_class", "max_threads_per_worker", "force_memmap",
                                  "monitor_stats"}
        if unknown:
            raise TypeError(f"unexpected keyword arguments {sorted(unknown)}")
        if trace_funcs is not None:
            raise NotImplementedError("custom trace functions: trace `pos` / `hamiltonian` only")
        state = self._preprocess_init_state(init_states, device)
        n = state.pos.shape[0]
        if adapters == "default":
            adapters = [DualAveragingStepSizeAdapter()]
        int_tr = self.transitions["integration_transition"]
        final, stats, trace = sample_chains(
            self._system, int_tr.integrator, state, _per_chain_rngs(self._rng, n), n_warm_up_iter,
            n_main_iter, integration_transition=int_tr,
            momentum_transition=self.transitions["momentum_transition"], adapters=adapters,
            stager=stager,
            trace_warm_up=trace_warm_up, trace_pos=True, trace_h=True, group=group)
        traces = {}
        if trace is not None:
            traces["pos"] = trace["pos"].transpose(0, 1)
            traces["hamiltonian"] = trace["hamiltonian"].transpose(0, 1)
        statistics = {k: v.transpose(0, 1) for k, v in stats.items()}
        if memmap_path is not None:
            write_chain_data(memmap_path, traces, statistics, first_chain_index)
        return HMCSampleChainsOutputs(final, traces, statistics)


class StaticMetropolisHMC(HamiltonianMonteCarlo):
    """Static integration time HMC with Metropolis accept step (samplers.py:1434-1498)."""

    def __init__(self, system, integrator, rng, n_step, momentum_transition=None):
        super().__init__(system, rng,
                         MetropolisStaticIntegrationTransition(system, integrator, n_step),
                         momentum_transition)

    @property
    def n_step(self):
        return self.transitions["integration_transition"].n_step

    @n_step.setter
    def n_step(self, value):
        self.transitions["integration_transition"].n_step = int(value)


class RandomMetropolisHMC(HamiltonianMonteCarlo):
    """Random integration time HMC with Metropolis accept step (samplers.py:1501-1572)."""

    def __init__(self, system, integrator, rng, n_step_range, momentum_transition=None):
        super().__init__(system, rng,
                         MetropolisRandomIntegrationTransition(system, integrator, n_step_range),
                         momentum_transition)

    @property
    def n_step_range(self):
        return self.transitions["integration_transition"].n_step_range


class _DynamicHMC(HamiltonianMonteCarlo):
    _transition_class = None

    def __init__(self, system, integrator, rng, *, max_tree_depth, max_delta_h,
                 termination_criterion, do_extra_subtree_checks, momentum_transition=None):
        super().__init__(
            system, rng,
            self._transition_class(system, integrator, max_tree_depth=max_tree_depth,
                                   max_delta_h=max_delta_h,
                                   termination_criterion=termination_criterion,
                                   do_extra_subtree_checks=do_extra_subtree_checks),
            momentum_transition)

    @property
    def max_tree_depth(self):
        return self.transitions["integration_transition"].max_tree_depth

    @property
    def max_delta_h(self):
        return self.transitions["integration_transition"].max_delta_h


class DynamicMultinomialHMC(_DynamicHMC):
    """Dynamic integration time HMC with multinomial sampling of the new state
    (samplers.py:1575-1683; defaults :1606-1609)."""

    _transition_class = MultinomialDynamicIntegrationTransition

    def __init__(self, system, integrator, rng, *, max_tree_depth=10, max_delta_h=1000,
                 termination_criterion=riemannian_no_u_turn_criterion,
                 do_extra_subtree_checks=True, momentum_transition=None):
        super().__init__(system, integrator, rng, max_tree_depth=max_tree_depth,
                         max_delta_h=max_delta_h, termination_criterion=termination_criterion,
                         do_extra_subtree_checks=do_extra_subtree_checks,
                         momentum_transition=momentum_transition)


class DynamicSliceHMC(_DynamicHMC):
    """Dynamic integration time HMC with slice sampling of the new state -- NUTS as in
    Hoffman & Gelman (samplers.py:1686-1796; defaults :1714-1717)."""

    _transition_class = SliceDynamicIntegrationTransition

    def __init__(self, system, integrator, rng, *, max_tree_depth=10, max_delta_h=1000.0,
                 termination_criterion=euclidean_no_u_turn_criterion,
                 do_extra_subtree_checks=False, momentum_transition=None):
        super().__init__(system, integrator, rng, max_tree_depth=max_tree_depth,
                         max_delta_h=max_delta_h, termination_criterion=termination_criterion,
                         do_extra_subtree_checks=do_extra_subtree_checks,
                         momentum_transition=momentum_transition)


__all__ = [
    "write_chain_data",
    "DynamicMultinomialHMC",
    "DynamicSliceHMC",
    "HMCSampleChainsOutputs",
    "HamiltonianMonteCarlo",
    "RandomMetropolisHMC",
    "StaticMetropolisHMC",
]
