"""TEST INFRASTRUCTURE ONLY -- NumPy statements of the benchmark target models.

These are the *user model callables* (``neg_log_dens``, ``grad_neg_log_dens``,
``hess_neg_log_dens``, ``mtp_neg_log_dens``, ``constr``, ``jacob_constr``,
``metric_func``, ``vjp_metric_func``) that the reference accepts as plain Python
functions (reference ``src/mici/systems.py:88-95, 776-784, 1310-1319, 1846-1859``).
The same closed-form models are compiled into ``libmici_b200.so`` as device
functors (``mici_b200/csrc/targets.cuh``); these NumPy versions exist so that the
*unmodified reference* (when importable) and the oracle port can be driven with
exactly the model the CUDA path integrates.  Model definitions follow SURVEY.md
section 8(d).

Nothing under ``mici_b200/`` imports this module.
"""

from __future__ import annotations

import numpy as np


class StdGaussian:
    """l(q) = 0.5 |q|^2  (config C0)."""

    name = "std_gaussian"

    def __init__(self, dim):
        self.dim = dim

    def neg_log_dens(self, q):
        return 0.5 * (q @ q)

    def grad_neg_log_dens(self, q):
        return q.copy()

    def hess_neg_log_dens(self, q):
        return np.identity(q.shape[0])

    def mtp_neg_log_dens(self, q):
        return lambda m: np.zeros_like(q)


class NealFunnel:
    """v = q[0], x = q[1:]; l = v^2/18 + (D-1) v / 2 + exp(-v) |x|^2 / 2  (config C1)."""

    name = "neal_funnel"

    def __init__(self, dim):
        self.dim = dim

    def neg_log_dens(self, q):
        v, x = q[0], q[1:]
        return v * v / 18.0 + 0.5 * (q.shape[0] - 1) * v + 0.5 * np.exp(-v) * (x @ x)

    def grad_neg_log_dens(self, q):
        v, x = q[0], q[1:]
        e = np.exp(-v)
        g = np.empty_like(q)
        g[0] = v / 9.0 + 0.5 * (q.shape[0] - 1) - 0.5 * e * (x @ x)
        g[1:] = e * x
        return g


class Banana:
    """Pairs (x, y) = (q[2k], q[2k+1]); l = sum x^2/8 + (y - b x^2)^2 / 2, b = 1/2 (C2)."""

    name = "banana"

    def __init__(self, dim, b=0.5):
        assert dim % 2 == 0
        self.dim = dim
        self.b = b

    def neg_log_dens(self, q):
        x, y = q[0::2], q[1::2]
        r = y - self.b * x * x
        return np.sum(x * x / 8.0 + 0.5 * r * r)

    def grad_neg_log_dens(self, q):
        x, y = q[0::2], q[1::2]
        r = y - self.b * x * x
        g = np.empty_like(q)
        g[0::2] = x / 4.0 - 2.0 * self.b * x * r
        g[1::2] = r
        return g

    def hess_neg_log_dens(self, q):
        x, y = q[0::2], q[1::2]
        b = self.b
        d = q.shape[0]
        h = np.zeros((d, d))
        i = np.arange(0, d, 2)
        h[i, i] = 0.25 - 2.0 * b * y + 6.0 * b * b * x * x
        h[i, i + 1] = -2.0 * b * x
        h[i + 1, i] = -2.0 * b * x
        h[i + 1, i + 1] = 1.0
        return h

    def mtp_neg_log_dens(self, q):
        x = q[0::2]
        b = self.b
        i = np.arange(0, q.shape[0], 2)

        def mtp(m):
            out = np.empty_like(q)
            out[0::2] = m[i, i] * (12.0 * b * b * x) - 2.0 * b * (m[i, i + 1] + m[i + 1, i])
            out[1::2] = -2.0 * b * m[i, i]
            return out

        return mtp


class Quartic:
    """l(q) = |q|^2 / 2 + (gamma / 4) sum_m (a_m . q)^4 with dense directions A [M x D]: the
    Hessian I + 3 gamma A^T diag((A q)^2) A and the third-derivative tensor are DENSE (no block
    structure for a SoftAbs eigensolver to exploit); mtp(V) = 6 gamma A^T ((A q) o diag(A V A^T))."""

    name = "quartic"

    def __init__(self, directions, gamma=1.0):
        self.a = np.asarray(directions)
        self.gamma = float(gamma)
        self.dim = self.a.shape[1]

    def neg_log_dens(self, q):
        s = self.a @ q
        return 0.5 * (q @ q) + 0.25 * self.gamma * np.sum(s**4)

    def grad_neg_log_dens(self, q):
        s = self.a @ q
        return q + self.gamma * (self.a.T @ s**3)

    def hess_neg_log_dens(self, q):
        s = self.a @ q
        return np.identity(self.dim) + 3.0 * self.gamma * ((self.a.T * s**2) @ self.a)

    def mtp_neg_log_dens(self, q):
        s = self.a @ q
        a, gamma = self.a, self.gamma
        return lambda m: 6.0 * gamma * (a.T @ (s * np.einsum("mi,ij,mj->m", a, m, a)))


class Quadratic:
    """l(q) = 0.5 q^T P q with dense SPD P (config C4)."""

    name = "quadratic"

    def __init__(self, prec):
        self.prec = np.asarray(prec)
        self.dim = self.prec.shape[0]

    def neg_log_dens(self, q):
        return 0.5 * (q @ (self.prec @ q))

    def grad_neg_log_dens(self, q):
        return self.prec @ q


class Rank1Metric:
    """Position-dependent dense metric M(q) = B + c q q^T (config C4)."""

    name = "rank1"

    def __init__(self, base, coeff):
        self.base = np.asarray(base)
        self.coeff = float(coeff)

    def metric_func(self, q):
        return self.base + self.coeff * np.outer(q, q)

    def vjp_metric_func(self, q):
        c = self.coeff
        return lambda v: c * ((v + v.T) @ q)


class HadamardMetric:
    """Position-dependent dense metric M(q) = B + c (q q^T) o S with B, S symmetric positive
    definite (``o`` = elementwise product).  M(q) is SPD by the Schur product theorem and, unlike
    the rank-1 model, is a full-rank perturbation of B: no low-rank identity applies, the
    reference's Cholesky / explicit-inverse path (matrices.py:1161-1188) is the only way through.
    dM_ij/dq_k = c S_ij (d_ik q_j + d_jk q_i), so vjp(V)_k = c sum_j (V_kj + V_jk) S_kj q_j."""

    name = "hadamard"

    def __init__(self, base, scale, coeff):
        self.base = np.asarray(base)
        self.scale = np.asarray(scale)
        self.coeff = float(coeff)

    def metric_func(self, q):
        return self.base + self.coeff * (np.outer(q, q) * self.scale)

    def vjp_metric_func(self, q):
        c, s = self.coeff, self.scale
        return lambda v: c * (((v + v.T) * s) @ q)


class Torus:
    """Density on a torus embedded in R^3 (config C3; reference README.md:315-337).

    rho = sqrt(x^2+y^2), theta = atan2(y, x), phi = atan2(z, rho - R),
    l = log1p(r cos(phi) / R) - log1p(alpha sin(4 theta) cos(phi)),
    c(q) = (rho - R)^2 + z^2 - r^2.
    """

    name = "torus"
    dim = 3
    n_constr = 1

    def __init__(self, R=1.0, r=0.5, alpha=0.9):
        self.R, self.r, self.alpha = float(R), float(r), float(alpha)

    def neg_log_dens(self, q):
        x, y, z = q
        rho = np.sqrt(x * x + y * y)
        theta = np.arctan2(y, x)
        phi = np.arctan2(z, rho - self.R)
        return np.log1p(self.r * np.cos(phi) / self.R) - np.log1p(
            np.sin(4 * theta) * np.cos(phi) * self.alpha
        )

    def grad_neg_log_dens(self, q):
        x, y, z = q
        a = self.r / self.R
        al = self.alpha
        rho2 = x * x + y * y
        rho = np.sqrt(rho2)
        u = rho - self.R
        theta = np.arctan2(y, x)
        phi = np.arctan2(z, u)
        s4, c4 = np.sin(4 * theta), np.cos(4 * theta)
        sp, cp = np.sin(phi), np.cos(phi)
        d1 = 1.0 + a * cp
        d2 = 1.0 + al * s4 * cp
        dl_dphi = -a * sp / d1 + al * s4 * sp / d2
        dl_dth = -4.0 * al * c4 * cp / d2
        w = u * u + z * z
        dphi_du = -z / w
        dphi_dz = u / w
        return np.array(
            [
                dl_dth * (-y / rho2) + dl_dphi * dphi_du * (x / rho),
                dl_dth * (x / rho2) + dl_dphi * dphi_du * (y / rho),
                dl_dphi * dphi_dz,
            ]
        )

    def constr(self, q):
        x, y, z = q
        rho = np.sqrt(x * x + y * y)
        return np.array([(rho - self.R) ** 2 + z * z - self.r**2])

    def jacob_constr(self, q):
        x, y, z = q
        rho = np.sqrt(x * x + y * y)
        f = 2.0 * (rho - self.R) / rho
        return np.array([[f * x, f * y, 2.0 * z]])

    def hess_constr(self, q):
        """Second derivatives of c = (rho - R)^2 + z^2 - r^2: with f = 2 (rho - R) / rho,
        d/dx (f x) = f + x df/dx, df/dx = 2 R x / rho^3."""
        x, y, z = q
        rho = np.sqrt(x * x + y * y)
        f = 2.0 * (rho - self.R) / rho
        g = 2.0 * self.R / rho**3
        return np.array([[[f + g * x * x, g * x * y, 0.0],
                          [g * x * y, f + g * y * y, 0.0],
                          [0.0, 0.0, 2.0]]])

    def mhp_constr(self, q):
        hess = self.hess_constr(q)
        return lambda m: np.sum(m[:, :, None] * hess, axis=(0, 1))


class Sphere:
    """Unit-sphere constraint c(q) = |q|^2 - 1 with Gaussian-tilted density (any D)."""

    name = "sphere"
    n_constr = 1

    def __init__(self, dim):
        self.dim = dim

    def neg_log_dens(self, q):
        return 0.5 * (q @ q) + q[0]

    def grad_neg_log_dens(self, q):
        g = q.copy()
        g[0] += 1.0
        return g

    def constr(self, q):
        return np.array([q @ q - 1.0])

    def jacob_constr(self, q):
        return 2.0 * q[None, :]

    def mhp_constr(self, q):  # hess[0] = 2 I
        return lambda m: 2.0 * m[0]


class MultiSphere:
    """``n_constr`` unit spheres: the coordinates are split into ``n_constr`` consecutive blocks
    of ``dim / n_constr`` and every block is constrained to its unit sphere,
    c_k(q) = |q_block_k|^2 - 1; tilted Gaussian density as ``Sphere``.  With a dense metric the
    Gram matrix J M^-1 J^T is a full ``n_constr x n_constr`` matrix (exercises the general C x C
    Cholesky / LU paths of the constrained integrator, C up to 8)."""

    name = "multi_sphere"

    def __init__(self, dim, n_constr):
        if dim % n_constr != 0:
            raise ValueError("dim must be a multiple of n_constr")
        self.dim, self.n_constr = dim, n_constr
        self.block = dim // n_constr

    def neg_log_dens(self, q):
        return 0.5 * (q @ q) + q[0]

    def grad_neg_log_dens(self, q):
        g = q.copy()
        g[0] += 1.0
        return g

    def constr(self, q):
        return (q.reshape(self.n_constr, self.block) ** 2).sum(-1) - 1.0

    def jacob_constr(self, q):
        jac = np.zeros((self.n_constr, self.dim))
        for k in range(self.n_constr):
            sl = slice(k * self.block, (k + 1) * self.block)
            jac[k, sl] = 2.0 * q[sl]
        return jac

    def mhp_constr(self, q):  # hess[k] = 2 I on block k
        def mhp(m):
            out = np.empty(self.dim)
            for k in range(self.n_constr):
                sl = slice(k * self.block, (k + 1) * self.block)
                out[sl] = 2.0 * m[k, sl]
            return out

        return mhp
