"""Synthetic benchmark / parity problems (BASELINE.json configs).

Host-side *input generation* only: shapes, seeds, shared metrics and initial states for the
five configurations C0..C4.  Everything is fp64 and generated with
``numpy.random.default_rng(BASE_SEED + k)`` so that the CUDA path, the oracle and the
reference all see identical inputs.  No integrator arithmetic lives here.
"""

from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

BASE_SEED = 20260924


@dataclass
class Problem:
    """A fully specified integrator workload.

    Attributes:
        name: config label (``"C0"`` .. ``"C4"``).
        integrator: ``"leapfrog"`` | ``"implicit_leapfrog"`` | ``"constrained_leapfrog"``.
        system: ``"euclidean"`` | ``"softabs_riemannian"`` | ``"dense_riemannian"`` |
            ``"constrained_euclidean"``.
        target: target-model name (``mici_b200.targets`` registry key).
        target_params: constructor kwargs of the target model.
        metric: ``None`` (identity), 1-D (diagonal) or 2-D (dense SPD) array -- the fixed
            metric of Euclidean systems.
        metric_model / metric_params: position-dependent metric (dense Riemannian only).
        step_size: integrator ``step_size``.
        pos, mom: ``[n_chains, dim]`` initial states.
    """

    name: str
    integrator: str
    system: str
    target: str
    target_params: dict
    step_size: float
    pos: np.ndarray
    mom: np.ndarray
    metric: np.ndarray | None = None
    metric_model: str | None = None
    metric_params: dict = field(default_factory=dict)
    system_kwargs: dict = field(default_factory=dict)
    integrator_kwargs: dict = field(default_factory=dict)

    @property
    def n_chains(self):
        return self.pos.shape[0]

    @property
    def dim(self):
        return self.pos.shape[1]

    @property
    def algorithmic_bytes_per_chain_step(self):
        """B_alg = 4 * D * 8: read q, p and write q, p in fp64."""
        return 32 * self.dim


def dense_spd_metric(rng, dim):
    """M = A A^T / D + I with A_ij ~ N(0, 1): condition number ~5."""
    a = rng.standard_normal((dim, dim))
    return a @ a.T / dim + np.identity(dim)


def c0_std_gaussian(n_chains=4, dim=10, seed=BASE_SEED + 0):
    rng = np.random.default_rng(seed)
    return Problem(
        name="C0",
        integrator="leapfrog",
        system="euclidean",
        target="std_gaussian",
        target_params={"dim": dim},
        step_size=0.1,
        pos=rng.standard_normal((n_chains, dim)),
        mom=rng.standard_normal((n_chains, dim)),
    )


def c1_funnel(n_chains=8192, dim=128, seed=BASE_SEED + 1, metric_kind="dense",
              integrator="leapfrog"):
    rng = np.random.default_rng(seed)
    metric = dense_spd_metric(rng, dim)
    pos = 0.1 * rng.standard_normal((n_chains, dim))
    z = rng.standard_normal((n_chains, dim))
    if metric_kind == "dense":
        mom = z @ np.linalg.cholesky(metric).T  # p0 = L z per chain
    elif metric_kind == "diagonal":
        metric = np.ascontiguousarray(metric.diagonal())
        mom = z * np.sqrt(metric)
    else:
        metric = None
        mom = z
    return Problem(
        name="C1",
        integrator=integrator,
        system="euclidean",
        target="neal_funnel",
        target_params={"dim": dim},
        step_size=0.01,
        pos=pos,
        mom=mom,
        metric=metric,
    )


def g1_gaussian_split(n_chains=256, dim=32, seed=BASE_SEED + 6, metric_kind="dense",
                      integrator="leapfrog", target="banana"):
    """Extra parity case for row N4: ``GaussianEuclideanMetricSystem`` (systems.py:369-474) --
    the registered target is the density relative to the standard Gaussian measure, the
    drift is the exact rotation in the eigenbasis of the metric."""
    rng = np.random.default_rng(seed)
    metric = dense_spd_metric(rng, dim)
    pos = 0.7 * rng.standard_normal((n_chains, dim))
    z = rng.standard_normal((n_chains, dim))
    if metric_kind == "dense":
        mom = z @ np.linalg.cholesky(metric).T
    elif metric_kind == "diagonal":
        metric = np.ascontiguousarray(metric.diagonal())
        mom = z * np.sqrt(metric)
    else:
        metric = None
        mom = z
    params = {"dim": dim, "b": 0.5} if target == "banana" else {"dim": dim}
    return Problem(
        name="G1",
        integrator=integrator,
        system="gaussian_euclidean",
        target=target,
        target_params=params,
        step_size=0.15,
        pos=pos,
        mom=mom,
        metric=metric,
    )


def c2_softabs_banana(n_chains=2048, dim=64, seed=BASE_SEED + 2, integrator="implicit_leapfrog"):
    rng = np.random.default_rng(seed)
    return Problem(
        name="C2",
        integrator=integrator,
        system="softabs_riemannian",
        target="banana",
        target_params={"dim": dim, "b": 0.5},
        step_size=0.1,
        pos=0.5 * rng.standard_normal((n_chains, dim)),
        mom=rng.standard_normal((n_chains, dim)),
        system_kwargs={"softabs_coeff": 1.0},
    )


def c6_softabs_quartic(n_chains=2048, dim=64, seed=BASE_SEED + 9, gamma=1.0,
                       integrator="implicit_leapfrog"):
    """C2's system (SoftAbs metric, implicit leapfrog) on a target with a DENSE Hessian:
    l = |q|^2/2 + (gamma/4) sum_m (a_m . q)^4 with D random directions.  The banana of C2 has a
    2 x 2 block-diagonal Hessian that a Jacobi eigensolver diagonalises in one round; this target
    needs every rotation of every sweep."""
    rng = np.random.default_rng(seed)
    directions = rng.standard_normal((dim, dim)) / np.sqrt(dim)
    return Problem(
        name="C6",
        integrator=integrator,
        system="softabs_riemannian",
        target="quartic",
        target_params={"directions": directions, "gamma": gamma},
        step_size=0.1,
        pos=0.5 * rng.standard_normal((n_chains, dim)),
        mom=rng.standard_normal((n_chains, dim)),
        system_kwargs={"softabs_coeff": 1.0},
    )


def c3_torus(n_chains=4096, seed=BASE_SEED + 3, R=1.0, r=0.5, alpha=0.9,
             dens_wrt_hausdorff=True):
    rng = np.random.default_rng(seed)
    theta, phi = rng.uniform(0, 2 * np.pi, size=(2, n_chains))
    pos = np.stack(
        [
            (R + r * np.cos(phi)) * np.cos(theta),
            (R + r * np.cos(phi)) * np.sin(theta),
            r * np.sin(phi),
        ],
        -1,
    )
    mom = rng.standard_normal((n_chains, 3))
    # project the initial momentum onto the cotangent space (identity metric):
    # p -= J^T (J J^T)^-1 J p with J = [2(rho-R)x/rho, 2(rho-R)y/rho, 2z]
    rho = np.sqrt(pos[:, 0] ** 2 + pos[:, 1] ** 2)
    f = 2.0 * (rho - R) / rho
    jac = np.stack([f * pos[:, 0], f * pos[:, 1], 2.0 * pos[:, 2]], -1)
    mom = mom - jac * ((jac * mom).sum(-1) / (jac * jac).sum(-1))[:, None]
    return Problem(
        name="C3",
        integrator="constrained_leapfrog",
        system="constrained_euclidean",
        target="torus",
        target_params={"R": R, "r": r, "alpha": alpha},
        step_size=0.1,
        pos=pos,
        mom=mom,
        integrator_kwargs={"n_inner_step": 1},
        system_kwargs={"dens_wrt_hausdorff": dens_wrt_hausdorff},
    )


def c4_dense_riemannian(n_chains=8192, dim=512, seed=BASE_SEED + 4, coeff=0.1,
                        integrator="implicit_leapfrog"):
    rng = np.random.default_rng(seed)
    g = rng.standard_normal((dim, dim))
    prec = np.identity(dim) + 0.1 * (g @ g.T) / dim
    base = dense_spd_metric(rng, dim)
    pos = 0.5 * rng.standard_normal((n_chains, dim))
    mom = rng.standard_normal((n_chains, dim)) @ np.linalg.cholesky(base).T
    return Problem(
        name="C4",
        integrator=integrator,
        system="dense_riemannian",
        target="quadratic",
        target_params={"prec": prec},
        step_size=0.05,
        pos=pos,
        mom=mom,
        metric_model="rank1",
        metric_params={"base": base, "coeff": coeff},
    )


def c5_dense_hadamard(n_chains=8192, dim=512, seed=BASE_SEED + 7, coeff=0.1,
                      integrator="implicit_leapfrog"):
    """C4's system with a metric that has NO low-rank structure: M(q) = B + c (q q^T) o S.  The
    quadratic target and the SPD matrices are generated like C4's; the metric can only be handled
    by the generic dense path (per-chain Cholesky + explicit inverse + dense VJP)."""
    rng = np.random.default_rng(seed)
    g = rng.standard_normal((dim, dim))
    prec = np.identity(dim) + 0.1 * (g @ g.T) / dim
    base = dense_spd_metric(rng, dim)
    scale = dense_spd_metric(rng, dim)
    pos = 0.5 * rng.standard_normal((n_chains, dim))
    mom = rng.standard_normal((n_chains, dim)) @ np.linalg.cholesky(base).T
    return Problem(
        name="C5",
        integrator=integrator,
        system="dense_riemannian",
        target="quadratic",
        target_params={"prec": prec},
        step_size=0.05,
        pos=pos,
        mom=mom,
        metric_model="hadamard",
        metric_params={"base": base, "scale": scale, "coeff": coeff},
    )


def sphere_constrained(n_chains=64, dim=10, seed=BASE_SEED + 5, metric_kind="dense",
                       dens_wrt_hausdorff=True):
    """Extra parity case for K6 beyond C3: unit sphere in R^dim, tilted Gaussian density,
    optional diagonal / dense metric (exercises the general-dimension constrained path)."""
    rng = np.random.default_rng(seed)
    pos = rng.standard_normal((n_chains, dim))
    pos /= np.linalg.norm(pos, axis=1, keepdims=True)
    mom = rng.standard_normal((n_chains, dim))
    if metric_kind == "dense":
        metric = dense_spd_metric(rng, dim)
    elif metric_kind == "diagonal":
        metric = rng.uniform(0.5, 2.0, dim)
    else:
        metric = None
    # project the momentum onto the cotangent space: p -= J^T (J M^-1 J^T)^-1 J M^-1 p, J = 2 q^T
    if metric is None:
        minv_j = pos
        minv_p = mom
    elif metric.ndim == 1:
        minv_j = pos / metric
        minv_p = mom / metric
    else:
        minv = np.linalg.inv(metric)
        minv_j = pos @ minv
        minv_p = mom @ minv
    lam = (pos * minv_p).sum(-1) / (pos * minv_j).sum(-1)
    mom = mom - lam[:, None] * pos
    return Problem(
        name="S1",
        integrator="constrained_leapfrog",
        system="constrained_euclidean",
        target="sphere",
        target_params={"dim": dim},
        step_size=0.2,
        pos=pos,
        mom=mom,
        metric=metric,
        integrator_kwargs={"n_inner_step": 1},
        system_kwargs={"dens_wrt_hausdorff": dens_wrt_hausdorff},
    )


def multi_sphere_constrained(n_chains=32, dim=16, n_constr=4, seed=BASE_SEED + 8,
                             metric_kind="dense", dens_wrt_hausdorff=True):
    """Extra parity case for K6 with SEVERAL constraints (C = n_constr <= 8): consecutive blocks
    of dim / n_constr coordinates each on their unit sphere; with a dense metric the Gram matrix
    and the Newton residual Jacobian are full C x C matrices."""
    rng = np.random.default_rng(seed)
    block = dim // n_constr
    pos = rng.standard_normal((n_chains, n_constr, block))
    pos /= np.linalg.norm(pos, axis=2, keepdims=True)
    pos = pos.reshape(n_chains, dim)
    mom = rng.standard_normal((n_chains, dim))
    if metric_kind == "dense":
        metric = dense_spd_metric(rng, dim)
        minv = np.linalg.inv(metric)
    elif metric_kind == "diagonal":
        metric = rng.uniform(0.5, 2.0, dim)
        minv = np.diag(1.0 / metric)
    else:
        metric = None
        minv = np.identity(dim)
    for i in range(n_chains):  # p -= J^T (J M^-1 J^T)^-1 J M^-1 p
        jac = np.zeros((n_constr, dim))
        for k in range(n_constr):
            jac[k, k * block:(k + 1) * block] = 2.0 * pos[i, k * block:(k + 1) * block]
        gram = jac @ minv @ jac.T
        mom[i] -= jac.T @ np.linalg.solve(gram, jac @ (minv @ mom[i]))
    return Problem(
        name="S2",
        integrator="constrained_leapfrog",
        system="constrained_euclidean",
        target="multi_sphere",
        target_params={"dim": dim, "n_constr": n_constr},
        step_size=0.15,
        pos=pos,
        mom=mom,
        metric=metric,
        integrator_kwargs={"n_inner_step": 1},
        system_kwargs={"dens_wrt_hausdorff": dens_wrt_hausdorff},
    )


CONFIGS = {
    "C0": c0_std_gaussian,
    "C1": c1_funnel,
    "C2": c2_softabs_banana,
    "C3": c3_torus,
    "C4": c4_dense_riemannian,
    "C5": c5_dense_hadamard,
    "C6": c6_softabs_quartic,
    "S1": sphere_constrained,
    "S2": multi_sphere_constrained,
    "G1": g1_gaussian_split,
}


def make_problem(name, **kwargs):
    """Build config ``name`` (``"C0"``..``"C4"``), optionally overriding sizes/seed."""
    return CONFIGS[name](**kwargs)
