"""Target-model registry of the CUDA path.

The reference accepts arbitrary Python callables for ``neg_log_dens`` and its derivatives
(``src/mici/systems.py:88-95``).  A fused GPU gradient cannot, so the engine ships a closed
registry of models compiled into ``libmici_b200.so`` (``mici_b200/csrc/targets.cuh``); an
instance of one of these classes is what is passed as ``neg_log_dens=`` to the systems in
``mici_b200.systems``.  Instances only carry ids and parameters -- no host arithmetic.
"""

from __future__ import annotations

import numpy as np

TARGET_STD_GAUSSIAN = 0
TARGET_NEAL_FUNNEL = 1
TARGET_BANANA = 2
TARGET_QUADRATIC = 3
TARGET_TORUS = 4
TARGET_SPHERE = 5
TARGET_MULTI_SPHERE = 6
TARGET_QUARTIC = 7

RMETRIC_SOFTABS = 0
RMETRIC_RANK1 = 1
RMETRIC_HADAMARD = 2


class Target:
    """Base class: ``target_id``, scalar ``params`` and an optional dense ``aux`` array."""

    target_id: int = -1
    name: str = ""
    n_constr: int = 0

    def __init__(self, dim, params=(), aux=None):
        self.dim = int(dim)
        self.params = tuple(float(x) for x in params)
        self.aux = None if aux is None else np.ascontiguousarray(aux, dtype=np.float64)

    def __repr__(self):
        return f"{type(self).__name__}(dim={self.dim}, params={self.params})"


class StdGaussian(Target):
    """l(q) = |q|^2 / 2."""

    target_id = TARGET_STD_GAUSSIAN
    name = "std_gaussian"

    def __init__(self, dim):
        super().__init__(dim)


class MultiSphere(Target):
    """``n_constr`` (2, 4 or 8) unit spheres on consecutive blocks of ``dim / n_constr``
    coordinates, c_k(q) = |q_block_k|^2 - 1; l = |q|^2/2 + q[0]."""

    target_id = TARGET_MULTI_SPHERE
    name = "multi_sphere"

    def __init__(self, dim, n_constr):
        if n_constr not in (2, 4, 8) or dim % n_constr:
            raise ValueError("n_constr must be 2, 4 or 8 and divide dim.")
        self.n_constr = n_constr
        super().__init__(dim, (float(n_constr),))


class NealFunnel(Target):
    """v = q[0], x = q[1:]; l = v^2/18 + (D-1) v/2 + exp(-v) |x|^2 / 2."""

    target_id = TARGET_NEAL_FUNNEL
    name = "neal_funnel"

    def __init__(self, dim):
        super().__init__(dim)


class Banana(Target):
    """Pairs (x, y): l = sum x^2/8 + (y - b x^2)^2 / 2."""

    target_id = TARGET_BANANA
    name = "banana"

    def __init__(self, dim, b=0.5):
        if dim % 2:
            raise ValueError("Banana target needs an even dimension.")
        super().__init__(dim, (b,))


class Quadratic(Target):
    """l(q) = q^T P q / 2 with dense SPD ``prec``."""

    target_id = TARGET_QUADRATIC
    name = "quadratic"

    def __init__(self, prec):
        prec = np.asarray(prec, dtype=np.float64)
        super().__init__(prec.shape[0], (), prec)


class Quartic(Target):
    """l(q) = |q|^2/2 + (gamma/4) sum_m (a_m . q)^4 with ``directions`` A [D x D]: dense Hessian
    and third-derivative tensor (SoftAbs systems)."""

    target_id = TARGET_QUARTIC
    name = "quartic"

    def __init__(self, directions, gamma=1.0):
        a = np.ascontiguousarray(directions, dtype=np.float64)
        if a.ndim != 2 or a.shape[0] != a.shape[1]:
            raise ValueError("`directions` must be a square matrix (D directions in R^D).")
        super().__init__(a.shape[1], (float(gamma),), a)


class Torus(Target):
    """Density on a torus in R^3 with constraint c(q) = (rho - R)^2 + z^2 - r^2
    (reference README.md:315-337)."""

    target_id = TARGET_TORUS
    name = "torus"
    n_constr = 1

    def __init__(self, R=1.0, r=0.5, alpha=0.9):
        super().__init__(3, (R, r, alpha))


class Sphere(Target):
    """l = |q|^2/2 + q[0] on the unit sphere c(q) = |q|^2 - 1."""

    target_id = TARGET_SPHERE
    name = "sphere"
    n_constr = 1

    def __init__(self, dim):
        super().__init__(dim)


class Rank1Metric:
    """Position-dependent dense metric M(q) = B + c q q^T."""

    rmetric_id = RMETRIC_RANK1
    name = "rank1"

    def __init__(self, base, coeff, force_low_rank_form=False, generic_rank1_vjp=False):
        """``force_low_rank_form``: evaluate the metric through the Sherman-Morrison identities
        against the shared explicit ``B^-1`` (O(D^2) per metric, never factorises) instead of the
        reference's per-chain Cholesky path -- an OPTIONAL policy; by default the metric is
        factorised per chain exactly as a generic dense metric.  ``generic_rank1_vjp``: form
        ``-(M^-1 p)(M^-1 p)^T`` explicitly and hand it to the dense VJP (matrices.py:1179-1181)
        instead of using the model's rank-one VJP."""
        base = np.ascontiguousarray(base, dtype=np.float64)
        # shared explicit inverse and log-determinant of B, built once on the host the way the
        # reference builds a fixed dense metric's inverse (matrices.py:1161-1188, 982-984)
        import scipy.linalg as sla  # noqa: PLC0415

        chol = np.linalg.cholesky(base)
        inv_lt = sla.solve_triangular(chol.T, np.identity(base.shape[0]), lower=False)
        inv = sla.solve_triangular(chol.T, inv_lt.T, lower=False)
        self.base = base
        self.aux = np.ascontiguousarray(np.concatenate([base.ravel(), inv.ravel()]))
        self.params = (
            float(coeff),
            float(2.0 * np.log(np.abs(chol.diagonal())).sum()),
            1.0 if force_low_rank_form else 0.0,
            1.0 if generic_rank1_vjp else 0.0,
        )


class HadamardMetric:
    """Position-dependent dense metric M(q) = B + c (q q^T) o S, B and S symmetric positive
    definite: a full-rank perturbation of B (Schur product theorem keeps it SPD), so only the
    generic dense path -- per-chain Cholesky, explicit inverse, dense VJP -- applies."""

    rmetric_id = RMETRIC_HADAMARD
    name = "hadamard"

    def __init__(self, base, scale, coeff, generic_rank1_vjp=False):
        base = np.ascontiguousarray(base, dtype=np.float64)
        scale = np.ascontiguousarray(scale, dtype=np.float64)
        if base.shape != scale.shape or base.ndim != 2 or base.shape[0] != base.shape[1]:
            raise ValueError("`base` and `scale` must be square matrices of the same shape.")
        self.base, self.scale = base, scale
        self.aux = np.ascontiguousarray(np.concatenate([base.ravel(), scale.ravel()]))
        self.params = (float(coeff), 0.0, 0.0, 1.0 if generic_rank1_vjp else 0.0)


REGISTRY = {
    cls.name: cls
    for cls in (StdGaussian, NealFunnel, Banana, Quadratic, Quartic, Torus, Sphere, MultiSphere)
}
METRIC_REGISTRY = {"rank1": Rank1Metric, "hadamard": HadamardMetric}


def make_target(name, **params):
    return REGISTRY[name](**params)


def make_metric_model(name, **params):
    return METRIC_REGISTRY[name](**params)
