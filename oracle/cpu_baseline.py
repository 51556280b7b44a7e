"""TEST / BENCH INFRASTRUCTURE ONLY -- time the oracle port on the host cores.

This is the CPU baseline printed beside the GPU number:
a process pool with one worker per host core, each looping the oracle's ``Integrator.step``
over its share of a bounded subsample of the workload's chains (per-chain cost does not depend
on the total chain count).  ``OMP_NUM_THREADS=1`` so BLAS does not oversubscribe.  Workers time
their own loops; throughput = total leapfrog steps / slowest worker's loop time (pool start-up
excluded).
"""

from __future__ import annotations

import multiprocessing as mp
import os
import time

import numpy as np


def _worker(args):
    cfg, kwargs, lo, hi, n_steps, reps = args
    os.environ["OMP_NUM_THREADS"] = "1"
    from mici_b200 import problems as pb  # noqa: PLC0415

    from . import drivers as dr  # noqa: PLC0415
    from . import mici_oracle as mo  # noqa: PLC0415

    problem = pb.make_problem(cfg, **kwargs)
    step, _, _ = dr.oracle_step_fn(problem)
    q0, p0 = problem.pos[lo:hi], problem.mom[lo:hi]
    mo.run_batch(step, q0[:1], p0[:1], None, 2)  # warm-up
    t0 = time.perf_counter()
    done = 0
    for _ in range(reps):
        _, _, _, n_done = mo.run_batch(step, q0, p0, None, n_steps)
        done += int(n_done.sum())
    return done, time.perf_counter() - t0


class Pool:
    """A spawn-context process pool kept alive across timed samples."""

    def __init__(self, n_workers=None):
        self.n_workers = n_workers or os.cpu_count() or 1
        os.environ["OMP_NUM_THREADS"] = "1"
        self._pool = mp.get_context("spawn").Pool(self.n_workers)

    def map(self, jobs):
        return self._pool.map(_worker, jobs, chunksize=1)

    def close(self):
        self._pool.close()
        self._pool.join()


def run(cfg, kwargs, chains_per_worker, n_steps, reps=1, n_workers=None, pool=None):
    """Returns dict(value=steps/s, cores, total_steps, seconds, sample)."""
    own = pool is None
    if own:
        pool = Pool(n_workers)
    n_workers = pool.n_workers
    jobs = [
        (cfg, kwargs, w * chains_per_worker, (w + 1) * chains_per_worker, n_steps, reps)
        for w in range(n_workers)
    ]
    try:
        res = pool.map(jobs)
    finally:
        if own:
            pool.close()
    total = sum(r[0] for r in res)
    slowest = max(r[1] for r in res)
    return {
        "value": total / slowest,
        "cores": n_workers,
        "total_steps": total,
        "seconds": slowest,
        "sample": (
            f"{n_workers} workers x {chains_per_worker} chains x {n_steps} leapfrog steps x {reps} "
            f"reps of {cfg} (oracle port, OMP_NUM_THREADS=1)"
        ),
    }
