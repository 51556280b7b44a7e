"""TEST INFRASTRUCTURE ONLY -- generate ``tests/golden/*.npz`` from the UNMODIFIED reference.

Run in the build container (where ``/root/reference`` exists):

    PYTHONDONTWRITEBYTECODE=1 OMP_NUM_THREADS=1 python -m oracle.make_golden

Every fixture stores the *outputs of the reference itself* (``mici`` imported from
``/root/reference/src``) for inputs that are regenerated deterministically from
``mici_b200.problems`` (seeded), plus a checksum of those inputs.  The GPU box has no
``/root/reference``; there the fixtures are the pinned statement of the reference's behaviour
that both the oracle port (``tests/test_oracle.py``) and the CUDA path
(``tests/test_parity_gpu.py``) are compared against.
"""

from __future__ import annotations

import os

import numpy as np

from mici_b200 import problems as pb

from . import drivers as dr

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# name -> (config, problem kwargs, step counts, integrator overrides)
CASES = {
    "c0_std_gaussian": ("C0", {}, (1, 5, 20), {}),
    "c1_funnel_dense": ("C1", {"n_chains": 32}, (1, 5, 20), {}),
    "c1_funnel_dense_d24": ("C1", {"n_chains": 16, "dim": 24}, (1, 20), {}),
    "c1_funnel_diag": ("C1", {"n_chains": 16, "dim": 40, "metric_kind": "diagonal"}, (1, 20), {}),
    "c1_funnel_identity": ("C1", {"n_chains": 16, "dim": 7, "metric_kind": "identity"}, (1, 20), {}),
    "c2_softabs_banana": ("C2", {"n_chains": 16}, (1, 5), {}),
    "c2_softabs_banana_d8": ("C2", {"n_chains": 32, "dim": 8}, (1, 5, 20), {}),
    "c3_torus": ("C3", {"n_chains": 64}, (1, 5, 20), {}),
    "c3_torus_inner3": ("C3", {"n_chains": 32}, (1, 5), {"n_inner_step": 3}),
    "n4_bcss2_funnel_d24": ("C1", {"n_chains": 12, "dim": 24, "integrator": "bcss2"}, (1, 5, 20), {}),
    "n4_bcss3_funnel_d40_diag": ("C1", {"n_chains": 8, "dim": 40, "metric_kind": "diagonal", "integrator": "bcss3"}, (1, 20), {}),
    "n4_bcss4_funnel_d130": ("C1", {"n_chains": 6, "dim": 130, "integrator": "bcss4"}, (1, 5), {}),
    "n4_midpoint_softabs_d8": ("C2", {"n_chains": 16, "dim": 8, "integrator": "implicit_midpoint"}, (1, 5, 20), {}),
    "n4_midpoint_dense_d32": ("C4", {"n_chains": 6, "dim": 32, "integrator": "implicit_midpoint"}, (1, 5), {}),
    "n4_steffensen_softabs_d8": ("C2", {"n_chains": 16, "dim": 8}, (1, 5), {"fixed_point_solver": "steffensen"}),
    "n4_steffensen_midpoint_dense_d16": ("C4", {"n_chains": 6, "dim": 16, "integrator": "implicit_midpoint"}, (1, 5), {"fixed_point_solver": "steffensen"}),
    "n4_quasi_newton_torus": ("C3", {"n_chains": 32}, (1, 5, 20), {"projection_solver": "quasi_newton"}),
    "n4_line_search_torus": ("C3", {"n_chains": 32}, (1, 5, 20), {"projection_solver": "newton_with_line_search"}),
    "n4_quasi_newton_sphere_dense_d10": ("S1", {"n_chains": 16, "dim": 10}, (1, 5), {"projection_solver": "quasi_newton", "n_inner_step": 2}),
    "n4_line_search_sphere_diag_d12": ("S1", {"n_chains": 16, "dim": 12, "metric_kind": "diagonal"}, (1, 5), {"projection_solver": "newton_with_line_search"}),
    "s1_sphere_dense_d10": ("S1", {"n_chains": 32, "dim": 10}, (1, 5, 20), {}),
    "s1_sphere_diag_d70_inner2": ("S1", {"n_chains": 8, "dim": 70, "metric_kind": "diagonal"}, (1, 5), {"n_inner_step": 2}),
    "s1_sphere_identity_d5": ("S1", {"n_chains": 16, "dim": 5, "metric_kind": "identity"}, (1, 20), {}),
    "c4_dense_riemannian_d64": ("C4", {"n_chains": 8, "dim": 64}, (1, 5), {}),
    "c4_dense_riemannian_d512": ("C4", {"n_chains": 2, "dim": 512}, (1,), {}),
}

# failure-path fixtures: step sizes chosen so that some chains raise IntegratorError
FAILURE_CASES = {
    "c2_softabs_banana_bigstep": ("C2", {"n_chains": 32, "dim": 8}, 0.6, (3,), {}),
    "c3_torus_bigstep": ("C3", {"n_chains": 64}, 0.4, (3,), {}),
    "n4_quasi_newton_torus_bigstep": ("C3", {"n_chains": 64}, 0.4, (3,), {"projection_solver": "quasi_newton"}),
    "n4_line_search_torus_bigstep": ("C3", {"n_chains": 64}, 0.45, (3,), {"projection_solver": "newton_with_line_search"}),
    "n4_midpoint_softabs_bigstep": ("C2", {"n_chains": 32, "dim": 8, "integrator": "implicit_midpoint"}, 0.9, (3,), {}),
}


def input_checksum(problem):
    return np.array([problem.pos.sum(), problem.mom.sum(), np.abs(problem.pos).sum()])


def mixed_dirs(n):
    d = np.ones(n, dtype=np.int32)
    d[1::3] = -1
    return d


def generate_case(name, cfg, kwargs, step_counts, overrides, step_size=None):
    problem = pb.make_problem(cfg, **kwargs)
    if step_size is not None:
        problem.step_size = step_size
    dirs = mixed_dirs(problem.n_chains)
    out = {
        "input_checksum": input_checksum(problem),
        "dirs": dirs,
        "step_size": np.array(problem.step_size),
        "step_counts": np.array(step_counts),
    }
    for n_steps in step_counts:
        r = dr.reference_run(problem, n_steps, dirs=dirs, **overrides)
        o = dr.oracle_run(problem, n_steps, dirs=dirs, **overrides)
        assert np.array_equal(r["status"], o["status"]), (name, r["status"], o["status"])
        assert np.array_equal(r["n_done"], o["n_done"]), name
        np.testing.assert_allclose(o["pos"], r["pos"], rtol=1e-12, atol=1e-14, err_msg=name)
        np.testing.assert_allclose(o["mom"], r["mom"], rtol=1e-12, atol=1e-14, err_msg=name)
        for key in ("pos", "mom", "status", "n_done", "h", "h_init"):
            out[f"{key}_{n_steps}"] = r[key]
        n_fail = int((r["status"] != 0).sum())
        print(f"{name:32s} steps={n_steps:3d} chains={problem.n_chains:3d} failed={n_fail}")
    np.savez(os.path.join(GOLDEN_DIR, name + ".npz"), **out)


def solver_known_answers():
    """Reference ``solve_fixed_point_direct`` on its own test problems
    (reference tests/test_solvers.py:25-39)."""
    mici = dr.import_reference()
    y = np.array([3.0, 5.0, 7.0])
    probs = {
        "babylonian": (lambda x: (y / x + x) / 2, np.ones_like(y)),
        "ratio": (lambda x: (x + y) / (x + 1), np.ones_like(y)),
        "cosine": (lambda x: np.cos(x), np.array([1.0])),
    }
    out = {}
    for k, (f, x0) in probs.items():
        for tol in (1e-6, 1e-8, 1e-10):
            out[f"{k}_{tol:g}"] = mici.solvers.solve_fixed_point_direct(f, x0, convergence_tol=tol)
            out[f"steffensen_{k}_{tol:g}"] = mici.solvers.solve_fixed_point_steffensen(
                f, x0, convergence_tol=tol)
    np.savez(os.path.join(GOLDEN_DIR, "solver_known_answers.npz"), **out)


HMC_CASES = {
    # name -> (config, kwargs, n_iter, n_step, seed)
    "hmc_c1_funnel_d16": ("C1", {"n_chains": 12, "dim": 16}, 6, 5, 11),
    "hmc_c0_std_gaussian": ("C0", {"n_chains": 6, "dim": 10}, 8, 7, 12),
}


def hmc_cases():
    """Static-HMC transitions (row N1) through the reference's own transition classes."""
    for name, (cfg, kwargs, n_iter, n_step, seed) in HMC_CASES.items():
        problem = pb.make_problem(cfg, **kwargs)
        if cfg == "C1":
            problem.step_size = 0.35
        r = dr.reference_hmc(problem, n_iter, n_step, seed)
        o = dr.oracle_hmc(problem, n_iter, n_step, seed)
        np.testing.assert_allclose(o["pos"], r["pos"], rtol=1e-12, atol=1e-14, err_msg=name)
        np.testing.assert_array_equal(o["dir"], r["dir"])
        np.testing.assert_allclose(o["metrop_accept_prob"], r["metrop_accept_prob"], rtol=1e-10)
        acc = float(o["accepted"].mean())
        print(f"{name:32s} iters={n_iter} n_step={n_step} accept rate={acc:.2f}")
        np.savez(os.path.join(GOLDEN_DIR, name + ".npz"), input_checksum=input_checksum(problem),
                 step_size=np.array(problem.step_size), accepted=o["accepted"],
                 **{k: r[k] for k in ("pos", "dir", "n_step", "metrop_accept_prob", "accept_stat")})


def main():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    hmc_cases()
    for name, (cfg, kwargs, steps, ov) in CASES.items():
        generate_case(name, cfg, kwargs, steps, ov)
    for name, (cfg, kwargs, eps, steps, ov) in FAILURE_CASES.items():
        generate_case(name, cfg, kwargs, steps, ov, step_size=eps)
    solver_known_answers()


if __name__ == "__main__":
    main()
