"""TEST / BENCH INFRASTRUCTURE ONLY -- time the UNMODIFIED reference (``mici``) on the host cores.

The CPU arm of ``bench.py`` (``--impl reference`` and the ``cpu_baseline`` objects): the
reference's own ``Integrator.step`` (``/root/reference/src/mici/integrators.py:63-80``) looped
over independent chains by a pool of worker processes, one per core this process may run on
(``os.sched_getaffinity``), ``OMP_NUM_THREADS=1`` -- BASELINE.md section 3, item 1.  The
reference package is imported from ``/root/reference/src`` in the build container and from the
verbatim copy ``oracle/_ref`` (``oracle/build_ref.sh``) on the GPU box; if neither exists the
oracle port (``oracle/cpu_baseline.py``) is timed instead and the result says ``kind: "port"``.

Every worker builds the config's ``(system, integrator)`` from the shared problem description,
steps its own chains for a bounded wall-clock budget and reports the leapfrog steps it completed
and its own loop time (pool start-up excluded).  ``IntegratorError``s are counted, not hidden: a
failed chain stops, as in ``transitions.py:292-295``.
"""

from __future__ import annotations

import multiprocessing as mp
import os
import time


def cgroup_cpu_quota():
    """CPUs this container may use according to its cgroup (v2 ``cpu.max`` / v1 cfs quota), or
    None when unlimited.  A box can show 128 schedulable CPUs and still be capped at 16."""
    try:
        with open("/sys/fs/cgroup/cpu.max") as f:
            quota, period = f.read().split()[:2]
        if quota != "max":
            return max(1, int(int(quota) / int(period) + 0.5))
    except (OSError, ValueError):
        pass
    try:
        with open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us") as f:
            quota = int(f.read())
        with open("/sys/fs/cgroup/cpu/cpu.cfs_period_us") as f:
            period = int(f.read())
        if quota > 0:
            return max(1, int(quota / period + 0.5))
    except (OSError, ValueError):
        pass
    return None


def usable_cores():
    """Worker count: the CPUs this process may be scheduled on, capped by the cgroup quota."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:  # pragma: no cover - non-Linux
        n = os.cpu_count() or 1
    quota = cgroup_cpu_quota()
    return n if quota is None else max(1, min(n, quota))


def _worker(args):
    cfg, kwargs, lo, hi, n_steps, budget_s, use_reference = args
    os.environ["OMP_NUM_THREADS"] = "1"
    import numpy as np  # noqa: PLC0415

    from mici_b200 import problems as pb  # noqa: PLC0415

    from . import drivers as dr  # noqa: PLC0415

    problem = pb.make_problem(cfg, **kwargs)
    q0, p0 = problem.pos[lo:hi], problem.mom[lo:hi]
    done = failed = 0
    if use_reference:
        mici = dr.import_reference()
        _, integrator = dr.build_reference(problem)

        def fresh(i):
            return mici.states.ChainState(pos=q0[i].copy(), mom=p0[i].copy(), dir=1)

        state = fresh(0)
        for _ in range(2):  # warm-up (imports, first-call caches)
            try:
                state = integrator.step(state)
            except mici.errors.Error:
                break
        t0 = time.perf_counter()
        rep = 0
        while True:
            for i in range(hi - lo):
                state = fresh(i)
                for _ in range(n_steps):
                    try:
                        state = integrator.step(state)
                    except mici.errors.IntegratorError:
                        failed += 1
                        break
                    done += 1
                if time.perf_counter() - t0 > budget_s:
                    break
            rep += 1
            if time.perf_counter() - t0 > budget_s:
                break
        return done, time.perf_counter() - t0, failed
    from . import mici_oracle as mo  # noqa: PLC0415

    step, _, _ = dr.oracle_step_fn(problem)
    mo.run_batch(step, q0[:1], p0[:1], None, 2)
    t0 = time.perf_counter()
    while True:
        for i in range(hi - lo):
            _, _, st, n_done = mo.run_batch(step, q0[i : i + 1], p0[i : i + 1], None, n_steps)
            done += int(n_done.sum())
            failed += int((np.asarray(st) != 0).sum())
            if time.perf_counter() - t0 > budget_s:
                break
        if time.perf_counter() - t0 > budget_s:
            break
    return done, time.perf_counter() - t0, failed


class Pool:
    """A spawn-context process pool kept alive across timed samples."""

    def __init__(self, n_workers=None):
        self.n_workers = n_workers or usable_cores()
        os.environ["OMP_NUM_THREADS"] = "1"
        self._pool = mp.get_context("spawn").Pool(self.n_workers)

    def map(self, jobs):
        return self._pool.map(_worker, jobs, chunksize=1)

    def close(self):
        self._pool.close()
        self._pool.join()


def run(cfg, kwargs, chains_per_worker, n_steps, budget_s, pool=None, n_workers=None, probe=True):
    """Time ``cfg`` on the pool.  Returns a dict with the aggregate rate (``value``), the per-core
    rate, a one-worker probe of the same loop (``probe_per_core``) and ``starved`` = the pool's
    per-core rate is below half of the probe's (oversubscribed / throttled host)."""
    from . import drivers as dr  # noqa: PLC0415

    use_reference = dr.reference_available()
    own = pool is None
    if own:
        pool = Pool(n_workers)
    n = pool.n_workers
    kw = dict(kwargs)
    kw["n_chains"] = n * chains_per_worker
    try:
        probe_rate = None
        if probe:
            d, t, _ = pool.map([(cfg, kw, 0, chains_per_worker, n_steps, min(budget_s, 2.0),
                                 use_reference)])[0]
            probe_rate = d / t if t > 0 else None
        jobs = [
            (cfg, kw, w * chains_per_worker, (w + 1) * chains_per_worker, n_steps, budget_s,
             use_reference)
            for w in range(n)
        ]
        res = pool.map(jobs)
    finally:
        if own:
            pool.close()
    total = sum(r[0] for r in res)
    slowest = max(r[1] for r in res)
    value = sum(r[0] / r[1] for r in res if r[1] > 0)  # workers time their own loops
    per_core = value / n
    out = {
        "value": value,
        "cores": n,
        "per_core": per_core,
        "probe_per_core": probe_rate,
        "starved": bool(probe_rate and per_core < 0.5 * probe_rate),
        "total_steps": total,
        "seconds": slowest,
        "failed_chains": sum(r[2] for r in res),
        "kind": "reference" if use_reference else "port",
        "schedulable_cpus": len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else None,
        "cgroup_cpu_quota": cgroup_cpu_quota(),
        "sample": (
            f"{n} workers x {chains_per_worker} chains x {n_steps} leapfrog steps, repeated for "
            f"{budget_s:g} s, of {cfg} ("
            + ("unmodified mici Integrator.step" if use_reference else "oracle port")
            + ", OMP_NUM_THREADS=1)"
        ),
    }
    return out
