"""Load the committed reference fixtures (tests/golden/*.npz, made by oracle/make_golden.py)."""

import os

import numpy as np

from mici_b200 import problems as pb
from oracle.make_golden import CASES, FAILURE_CASES, GOLDEN_DIR, input_checksum

RTOL = 1e-10  # north_star: 1e-10 rtol in fp64
ATOL = 1e-12


def case_names(failures=False):
    return list(FAILURE_CASES if failures else CASES)


def load_case(name):
    """Return (problem, dirs, overrides, golden npz dict)."""
    if name in CASES:
        cfg, kwargs, _, overrides = CASES[name]
        step_size = None
    else:
        cfg, kwargs, step_size, _, overrides = FAILURE_CASES[name]
    problem = pb.make_problem(cfg, **kwargs)
    if step_size is not None:
        problem.step_size = step_size
    g = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))
    np.testing.assert_allclose(input_checksum(problem), g["input_checksum"], rtol=1e-13)
    assert float(g["step_size"]) == problem.step_size
    return problem, g["dirs"], overrides, g


def assert_matches_golden(out, g, n_steps, rtol=RTOL, atol=ATOL, label="", kind_flip_frac=0.0):
    """Compare a run's (pos, mom, status, n_done[, h]) with the reference fixture.

    ``kind_flip_frac`` (failure-path fixtures only): which chains fail, and at which step, must
    match exactly, but for at most this fraction of the FAILED chains the failure kind
    (ConvergenceError vs NonReversibleStepError) may differ.  When a diverging Newton / line-search
    iteration is chaotic the kind is ill-conditioned: the reference itself flips between the two
    under 1e-15 relative perturbations of the input (checked for chain 8 of
    ``n4_line_search_torus_bigstep``: 10 / 30 split over 40 perturbed runs)."""
    ref_status = g[f"status_{n_steps}"]
    if kind_flip_frac > 0.0:
        np.testing.assert_array_equal(out["status"] != 0, ref_status != 0, err_msg=f"{label} failed")
        n_failed = max(int((ref_status != 0).sum()), 1)
        n_flip = int((out["status"] != ref_status).sum())
        assert n_flip <= kind_flip_frac * n_failed, (label, n_flip, n_failed)
    else:
        np.testing.assert_array_equal(out["status"], ref_status, err_msg=f"{label} status")
    np.testing.assert_array_equal(out["n_done"], g[f"n_done_{n_steps}"], err_msg=f"{label} n_done")
    for key in ("pos", "mom"):
        np.testing.assert_allclose(
            out[key], g[f"{key}_{n_steps}"], rtol=rtol, atol=atol, err_msg=f"{label} {key}"
        )
    if "h" in out and out["h"] is not None:
        ok = np.isfinite(g[f"h_{n_steps}"])
        np.testing.assert_allclose(
            np.asarray(out["h"])[ok], g[f"h_{n_steps}"][ok], rtol=rtol, atol=1e-9, err_msg=f"{label} h"
        )


def load_hmc_case(name):
    from oracle.make_golden import HMC_CASES

    cfg, kwargs, n_iter, n_step, seed = HMC_CASES[name]
    problem = pb.make_problem(cfg, **kwargs)
    g = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))
    problem.step_size = float(g["step_size"])
    np.testing.assert_allclose(input_checksum(problem), g["input_checksum"], rtol=1e-13)
    return problem, n_iter, n_step, seed, g


STAGE_NAMES = {0: None, 1: "fast", 2: "all"}


def load_adapt_case(name):
    from oracle.make_golden import ADAPT_CASES

    cfg, kwargs, specs, stager_kwargs, n_warm, n_main, n_step, seed = ADAPT_CASES[name]
    problem = pb.make_problem(cfg, **kwargs)
    g = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))
    np.testing.assert_allclose(input_checksum(problem), g["input_checksum"], rtol=1e-13)
    stages = [(int(n), STAGE_NAMES[int(w)]) for n, w in zip(g["stage_n_iter"], g["stage_which"])]
    return problem, specs, stager_kwargs, n_warm, n_main, n_step, seed, stages, g


def load_nuts_case(name):
    from oracle.make_golden import NUTS_CASES

    cfg, kwargs, eps, n_iter, seed, opts = NUTS_CASES[name]
    problem = pb.make_problem(cfg, **kwargs)
    problem.step_size = eps
    g = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))
    np.testing.assert_allclose(input_checksum(problem), g["input_checksum"], rtol=1e-13)
    return problem, n_iter, seed, opts, g
