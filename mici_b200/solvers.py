"""Solver selectors mirroring ``mici.solvers`` (reference ``src/mici/solvers.py``).

In the reference these are Python functions that iterate over user callables
(``solve_fixed_point_direct`` solvers.py:47-94, ``solve_projection_onto_manifold_newton``
solvers.py:346-469).  In the batched engine the iteration runs *inside* the integrator kernels
(registers + per-chain convergence masks), so the objects here are the *names* a caller passes
as ``fixed_point_solver=`` / ``projection_solver=`` (defaults identical to the reference:
integrators.py:444, 862) together with the same keyword arguments and defaults.  Passing any
other solver is rejected -- there is no host-side fallback path.
"""

from __future__ import annotations

from .errors import Error


class _DeviceSolver:
    """Marker object for a solver that is fused into the CUDA integrator kernels."""

    def __init__(self, name, defaults, kind=0):
        self.__name__ = name
        self.defaults = dict(defaults)
        self.kind = kind  # MB200_FP_SOLVER_* for fixed-point solvers

    def resolve_kwargs(self, kwargs):
        out = dict(self.defaults)
        for k, v in (kwargs or {}).items():
            if k == "norm":
                if v is not maximum_norm:
                    raise ValueError("Only `maximum_norm` is available in the fused solvers.")
                continue
            if k not in out:
                raise TypeError(f"{self.__name__}() got an unexpected keyword argument {k!r}")
            out[k] = v
        return out

    def __call__(self, *args, **kwargs):
        raise Error(
            f"{self.__name__} runs inside the CUDA integrator kernels and cannot be called with "
            "Python callables; pass it as `fixed_point_solver=` / `projection_solver=`."
        )

    def __repr__(self):
        return f"<mici_b200 fused solver {self.__name__}>"


def maximum_norm(vct):
    """Maximum (L-infinity) norm (solvers.py:25-27) of a tensor along its last axis."""
    return abs(vct).amax(-1) if hasattr(vct, "amax") else abs(vct).max()


#: solvers.py:47-94 defaults
solve_fixed_point_direct = _DeviceSolver(
    "solve_fixed_point_direct",
    {"convergence_tol": 1e-9, "divergence_tol": 1e10, "max_iters": 100},
)

#: solvers.py:97-154 defaults (Aitken / Steffensen acceleration, two evaluations per iteration)
solve_fixed_point_steffensen = _DeviceSolver(
    "solve_fixed_point_steffensen",
    {"convergence_tol": 1e-9, "divergence_tol": 1e10, "max_iters": 100},
    kind=1,
)

#: solvers.py:346-469 defaults
solve_projection_onto_manifold_newton = _DeviceSolver(
    "solve_projection_onto_manifold_newton",
    {"constraint_tol": 1e-9, "position_tol": 1e-8, "divergence_tol": 1e10, "max_iters": 50},
)

#: solvers.py:195-343 defaults (residual Jacobian frozen at the previous state)
solve_projection_onto_manifold_quasi_newton = _DeviceSolver(
    "solve_projection_onto_manifold_quasi_newton",
    {"constraint_tol": 1e-9, "position_tol": 1e-8, "divergence_tol": 1e10, "max_iters": 50},
    kind=1,
)

#: solvers.py:472-614 defaults (full Newton direction with step halving)
solve_projection_onto_manifold_newton_with_line_search = _DeviceSolver(
    "solve_projection_onto_manifold_newton_with_line_search",
    {"constraint_tol": 1e-9, "position_tol": 1e-8, "divergence_tol": 1e10, "max_iters": 50,
     "max_line_search_iters": 10},
    kind=2,
)
