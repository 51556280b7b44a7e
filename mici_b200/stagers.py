"""Warm-up stage schedules -- host-side mirror of ``mici.stagers`` (reference
``src/mici/stagers.py``), part of "next" row N3: which adapters run for how many iterations.

A stage is ``ChainStage(n_iter, adapters, trace, record_stats)``; ``adapters`` is a list of
batched ``mici_b200.adapters`` objects (the reference keys them by transition; the batched
sampler has exactly one adapted transition, the integration transition)."""

from __future__ import annotations

from typing import NamedTuple


class ChainStage(NamedTuple):
    """stagers.py:17-23."""

    n_iter: int
    adapters: list | None
    trace: bool
    record_stats: bool


class WarmUpStager:
    """One adaptive stage with all adapters, then the main stage (stagers.py:79-118)."""

    def stages(self, n_warm_up_iter, n_main_iter, adapters, *, trace_warm_up=False):
        out = {}
        if n_warm_up_iter > 0:
            out["Adaptive warm up"] = ChainStage(n_warm_up_iter, list(adapters), trace_warm_up,
                                                 trace_warm_up)
        if n_main_iter > 0:
            out["Main non-adaptive"] = ChainStage(n_main_iter, None, True, True)
        return out


class WindowedWarmUpStager:
    """Stan-style windowed warm-up (stagers.py:121-291): an initial fast stage, growing
    memoryless slow windows (all adapters, states reset per window), a final fast stage."""

    def __init__(self, n_init_slow_window_iter=25, n_init_fast_stage_iter=75,
                 n_final_fast_stage_iter=50, slow_window_multiplier=2.0):
        self.n_init_slow_window_iter = n_init_slow_window_iter
        self.n_init_fast_stage_iter = n_init_fast_stage_iter
        self.n_final_fast_stage_iter = n_final_fast_stage_iter
        self.slow_window_multiplier = slow_window_multiplier

    def slow_windows(self, n_warm_up_iter):
        """``(n_init_fast, [slow window sizes], n_final_fast)`` for a warm-up of the given
        length (stagers.py:216-268)."""
        n_fast0, n_slow0, n_fast1 = (self.n_init_fast_stage_iter, self.n_init_slow_window_iter,
                                     self.n_final_fast_stage_iter)
        if n_fast0 + n_slow0 + n_fast1 > n_warm_up_iter:
            n_fast0 = int(0.15 * n_warm_up_iter)
            n_fast1 = int(0.1 * n_warm_up_iter)
            n_slow0 = n_warm_up_iter - n_fast0 - n_fast1
        n_slow_total = n_warm_up_iter - n_fast0 - n_fast1
        windows, done, size = [], 0, n_slow0
        while done < n_slow_total:
            # a window that would leave less than one further full window takes all that is left
            if done + int((1 + self.slow_window_multiplier) * size) > n_slow_total:
                size = n_slow_total - done
            windows.append(size)
            done += size
            size = int(self.slow_window_multiplier * size)
        return n_fast0, windows, n_fast1

    def stages(self, n_warm_up_iter, n_main_iter, adapters, *, trace_warm_up=False):
        adapters = list(adapters)
        fast = [a for a in adapters if a.is_fast]
        out = {}
        if n_warm_up_iter > 0:
            n_fast0, windows, n_fast1 = self.slow_windows(n_warm_up_iter)
            out["Initial fast adaptive"] = ChainStage(n_fast0, fast, trace_warm_up, trace_warm_up)
            for i, n_iter in enumerate(windows):
                out[f"Slow adaptive ({i + 1}/{len(windows)})"] = ChainStage(
                    n_iter, adapters, trace_warm_up, trace_warm_up)
            out["Final fast adaptive"] = ChainStage(n_fast1, fast, trace_warm_up, trace_warm_up)
        if n_main_iter > 0:
            out["Main non-adaptive"] = ChainStage(n_main_iter, None, True, True)
        return out
