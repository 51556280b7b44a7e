"""Hamiltonian systems of the batched engine -- same class names, constructor keywords and
method meanings as ``mici.systems`` (reference ``src/mici/systems.py``) for the classes on the
hot path:

* ``EuclideanMetricSystem``                   systems.py:264-366
* ``DenseConstrainedEuclideanMetricSystem``   systems.py:619-873, 876-1031
* ``DenseRiemannianMetricSystem``             systems.py:1187-1402, 1710-1760
* ``SoftAbsRiemannianMetricSystem``           systems.py:1763-1920

Differences forced by the device: ``neg_log_dens`` is an instance of
``mici_b200.targets.Target`` (a model compiled into the library) instead of a Python callable,
its derivatives are implied, and states are batched ``mici_b200.states.ChainState`` s holding
``[n_chains, dim]`` fp64 CUDA tensors.  Methods return per-chain tensors.
"""

from __future__ import annotations

import ctypes

import numpy as np
import scipy.linalg as sla
import torch

from . import _lib
from .errors import LinAlgError
from .targets import RMETRIC_SOFTABS, HadamardMetric, Rank1Metric, Target

METRIC_IDENTITY, METRIC_DIAGONAL, METRIC_DENSE = 0, 1, 2


def _explicit_spd_inverse(array):
    """Dense explicit inverse built the way the reference builds ``metric.inv``: lower Cholesky
    factor (matrices.py:1161-1173) then two triangular solves against the identity
    (matrices.py:897-912, 1060-1061, 1183-1188).  Host side, once per metric assignment."""
    if not np.all(np.isfinite(array)):
        raise LinAlgError("Array is not finite.")
    try:
        chol = np.linalg.cholesky(array)
    except np.linalg.LinAlgError as e:
        raise LinAlgError("Cholesky factorisation failed.") from e
    inv_lt = sla.solve_triangular(chol.T, np.identity(array.shape[0]), lower=False, check_finite=False)
    inv = sla.solve_triangular(chol.T, inv_lt.T, lower=False, check_finite=False)
    return chol, inv


class _FixedMetric:
    """Host-side description of a fixed metric plus lazily uploaded device buffers."""

    def __init__(self, metric):
        if metric is None:
            self.kind, self.array, self.inv, self.sqrt = METRIC_IDENTITY, None, None, None
        else:
            if isinstance(metric, torch.Tensor):
                metric = metric.detach().cpu().numpy()
            metric = np.asarray(metric, dtype=np.float64)
            if metric.ndim == 1:
                if not np.all(metric > 0):
                    raise ValueError("Diagonal values must all be positive.")
                self.kind, self.array = METRIC_DIAGONAL, metric
                self.inv, self.sqrt = 1.0 / metric, metric**0.5
            elif metric.ndim == 2:
                self.kind, self.array = METRIC_DENSE, metric
                self.sqrt, self.inv = _explicit_spd_inverse(metric)
            else:
                msg = (
                    "If NumPy ndarray value is used for `metric` must be either 1D (diagonal "
                    "matrix) or 2D (dense positive definite matrix)."
                )
                raise ValueError(msg)
        self._dev = {}

    @classmethod
    def from_covariance(cls, covar):
        """The metric ``DensePositiveDefiniteMatrix(covar).inv`` that the covariance adapter
        assigns (adapters.py:642): with ``L = chol(covar)`` its array is the explicit inverse
        ``L^-T L^-1`` and its factor ``L^-T`` (matrices.py:1183-1188, 1209-1216), so ``metric.inv``
        multiplies by ``L L^T`` (matrices.py:1041-1046, 1060-1061) and ``metric.sqrt @ z`` solves
        ``L^T x = z`` (matrices.py:897-903) -- held here as the explicit upper-triangular factor."""
        covar = np.asarray(covar, dtype=np.float64)
        chol, explicit_inv = _explicit_spd_inverse(covar)
        self = cls.__new__(cls)
        self.kind, self.array = METRIC_DENSE, explicit_inv
        self.inv = chol @ chol.T
        self.sqrt = sla.solve_triangular(chol.T, np.identity(covar.shape[0]), lower=False,
                                         check_finite=False)
        self._dev = {}
        return self

    @property
    def shape(self):
        return (None, None) if self.array is None else (self.array.shape[0],) * 2

    def inv_device(self, device):
        """Device copy of 1/diag or of the explicit dense inverse (None for identity)."""
        if self.kind == METRIC_IDENTITY:
            return None
        key = ("inv", str(device))
        if key not in self._dev:
            self._dev[key] = torch.as_tensor(np.ascontiguousarray(self.inv), device=device)
        return self._dev[key]

    def sqrt_device(self, device):
        if self.kind == METRIC_IDENTITY:
            return None
        key = ("sqrt", str(device))
        if key not in self._dev:
            self._dev[key] = torch.as_tensor(np.ascontiguousarray(self.sqrt), device=device)
        return self._dev[key]

    def __getstate__(self):
        d = dict(self.__dict__)
        d["_dev"] = {}
        return d


def _batched(state):
    """View of a state as ([n, D] pos, [n, D] mom, dir tensor-or-None, squeeze flag).

    NumPy arrays (the reference's own ``ChainState`` storage, states.py:160-305) are accepted
    and moved to the current CUDA device; callers convert results back (``_like_input``)."""
    pos, mom = state.pos, state.mom
    if isinstance(pos, np.ndarray):
        pos = torch.as_tensor(np.ascontiguousarray(pos, dtype=np.float64), device="cuda")
        mom = None if mom is None else torch.as_tensor(
            np.ascontiguousarray(mom, dtype=np.float64), device="cuda")
    single = pos.ndim == 1
    if single:
        pos, mom = pos[None], (None if mom is None else mom[None])
    d = state.dir if "dir" in state else 1
    return pos, mom, d, single


def _like_input(ref, value):
    """Return ``value`` in the storage type of ``ref`` (NumPy in -> NumPy out)."""
    if isinstance(ref, np.ndarray) and isinstance(value, torch.Tensor):
        return value.cpu().numpy()
    return value


def _dir_tensor(d, n, device):
    """``dir`` as an int32 device tensor [n] (or None meaning all +1)."""
    if isinstance(d, torch.Tensor):
        return d.to(device=device, dtype=torch.int32).reshape(-1).contiguous()
    d = int(d)
    if d == 1:
        return None
    return torch.full((n,), d, dtype=torch.int32, device=device)


class System:
    """Base class (systems.py:39-229): holds the target model and builds ``mb200_model``."""

    def __init__(self, neg_log_dens, *, grad_neg_log_dens=None, backend=None):
        if not isinstance(neg_log_dens, Target):
            msg = (
                "mici_b200 systems take a `mici_b200.targets.Target` instance as `neg_log_dens` "
                "(models are compiled into the CUDA library); Python callables are not supported."
            )
            raise TypeError(msg)
        if grad_neg_log_dens is not None or backend is not None:
            raise ValueError("Derivatives are fused into the kernels; pass neither "
                             "`grad_neg_log_dens` nor `backend`.")
        self.target = neg_log_dens
        self._rmetric_id = 0
        self._rmetric_params = ()
        self._rmetric_aux = None
        self._dev = {}

    def __getstate__(self):
        d = dict(self.__dict__)
        d["_dev"] = {}
        return d

    def _aux_device(self, name, array, device):
        if array is None:
            return None
        key = (name, str(device))
        if key not in self._dev:
            self._dev[key] = torch.as_tensor(array, device=device).contiguous()
        return self._dev[key]

    def _model(self, device):
        m = _lib.Model()
        t = self.target
        m.target_id = t.target_id
        m.n_target_params = len(t.params)
        for i, v in enumerate(t.params):
            m.target_params[i] = v
        if not getattr(self, "dens_wrt_hausdorff", True):
            # constrained systems: density given with respect to the Lebesgue measure
            m.target_params[_lib.MAX_PARAMS - 1] = 1.0
        aux = self._aux_device("target_aux", t.aux, device)
        m.target_aux = None if aux is None else aux.data_ptr()
        m.rmetric_id = self._rmetric_id
        m.n_rmetric_params = len(self._rmetric_params)
        for i, v in enumerate(self._rmetric_params):
            m.rmetric_params[i] = v
        raux = self._aux_device("rmetric_aux", self._rmetric_aux, device)
        m.rmetric_aux = None if raux is None else raux.data_ptr()
        return m

    def dh_dmom(self, state):
        return self.dh2_dmom(state)


class TractableFlowSystem(System):
    """systems.py:232-261."""


class EuclideanMetricSystem(TractableFlowSystem):
    """Euclidean Hamiltonian system with a fixed metric (systems.py:264-366).

    ``metric``: ``None`` (identity), 1-D array (diagonal) or 2-D array (dense SPD), coerced as
    in systems.py:332-346.  Assignable (adapters set it: adapters.py:513, 642).
    """

    def __init__(self, neg_log_dens, *, metric=None, grad_neg_log_dens=None, backend=None):
        super().__init__(neg_log_dens, grad_neg_log_dens=grad_neg_log_dens, backend=backend)
        self.metric = metric

    @property
    def metric(self):
        return self._metric

    @metric.setter
    def metric(self, value):
        self._metric = value if isinstance(value, _FixedMetric) else _FixedMetric(value)

    def _eval(self, state, *, nld=False, grad=False, vel=False, kin=False):
        pos, mom, _, single = _batched(state)
        n, dim = pos.shape
        dev = pos.device
        lib = _lib.load()
        pos = pos.contiguous()
        mom = pos if mom is None else mom.contiguous()
        out = {}
        if nld:
            out["nld"] = torch.empty(n, dtype=torch.float64, device=dev)
        if grad:
            out["grad"] = torch.empty_like(pos)
        if vel:
            out["vel"] = torch.empty_like(pos)
        if kin:
            out["kin"] = torch.empty(n, dtype=torch.float64, device=dev)
        if nld or grad:
            model = self._model(dev)
        else:  # M^-1 p and p.M^-1 p do not involve the target (constrained targets have no
            model = _lib.Model()  # Euclidean-eval functor): neutral model
            model.target_id = 0
        rc = lib.mb200_euclidean_eval(
            _lib.ptr(pos), _lib.ptr(mom), n, dim, self._metric.kind,
            _lib.ptr(self._metric.inv_device(dev)), ctypes.byref(model),
            _lib.ptr(out.get("nld")), _lib.ptr(out.get("grad")), _lib.ptr(out.get("vel")),
            _lib.ptr(out.get("kin")), _lib.current_stream_ptr(dev),
        )
        _lib.check(rc, "mb200_euclidean_eval")
        if single:
            out = {k: v[0] for k, v in out.items()}
        return {k: _like_input(state.pos, v) for k, v in out.items()}  # NumPy in -> NumPy out

    def neg_log_dens(self, state):
        return self._eval(state, nld=True)["nld"]

    def grad_neg_log_dens(self, state):
        return self._eval(state, grad=True)["grad"]

    def h1(self, state):
        return self.neg_log_dens(state)

    def dh1_dpos(self, state):
        return self.grad_neg_log_dens(state)

    def h2(self, state):
        return self._eval(state, kin=True)["kin"]

    def dh2_dmom(self, state):
        return self._eval(state, vel=True)["vel"]

    def dh2_dpos(self, state):
        if isinstance(state.pos, np.ndarray):
            return np.zeros_like(state.pos)
        return torch.zeros_like(state.pos)

    def dh_dpos(self, state):
        return self.dh1_dpos(state)

    def h(self, state):
        """h = h1 + h2 (systems.py:187-196), one fused kernel."""
        pos, mom, _, single = _batched(state)
        n, dim = pos.shape
        dev = pos.device
        h = torch.empty(n, dtype=torch.float64, device=dev)
        model = self._model(dev)
        rc = _lib.load().mb200_hamiltonian_euclidean(
            _lib.ptr(pos.contiguous()), _lib.ptr(mom.contiguous()), n, dim, self._metric.kind,
            _lib.ptr(self._metric.inv_device(dev)), ctypes.byref(model), _lib.ptr(h),
            _lib.current_stream_ptr(dev),
        )
        _lib.check(rc, "mb200_hamiltonian_euclidean")
        return _like_input(state.pos, h[0] if single else h)

    def h1_flow(self, state, dt):
        """p -= dt * grad l(q) (systems.py:143-152); ``dt`` scalar or per-chain tensor."""
        state.mom = state.mom - _col(dt) * self.dh1_dpos(state)

    def h2_flow(self, state, dt):
        """q += dt * M^-1 p (systems.py:362-363)."""
        state.pos = state.pos + _col(dt) * self.dh2_dmom(state)

    def sample_momentum(self, state, rng):
        """``metric.sqrt @ N(0, I)`` (systems.py:365-366).  The variates come from ``rng`` (a
        NumPy generator, a sequence of per-chain NumPy generators, or a device
        ``torch.Generator``: see ``mici_b200.transitions``); the product ``L z`` runs on the
        device through ``mb200_euclidean_eval`` with the transposed factor in the metric slot."""
        from .transitions import _normals  # noqa: PLC0415

        pos = state.pos if state.pos.ndim == 2 else state.pos[None]
        # NumPy-held states (the reference's own ChainState storage): device work on the current
        # CUDA device, NumPy back out
        dev = torch.device("cuda") if isinstance(pos, np.ndarray) else pos.device
        z = _normals(rng, tuple(pos.shape), dev).contiguous()
        m = self._metric
        if m.kind != METRIC_IDENTITY:
            n, dim = z.shape
            out = torch.empty_like(z)
            key = ("sqrt_t", str(z.device))
            if key not in m._dev:
                fac = m.sqrt if m.kind == METRIC_DIAGONAL else np.ascontiguousarray(m.sqrt.T)
                m._dev[key] = torch.as_tensor(fac, device=z.device).contiguous()
            model = _lib.Model()  # the product L z does not involve the target
            model.target_id = 0
            rc = _lib.load().mb200_euclidean_eval(
                _lib.ptr(z), _lib.ptr(z), n, dim, m.kind, _lib.ptr(m._dev[key]),
                ctypes.byref(model), None, None, _lib.ptr(out), None,
                _lib.current_stream_ptr(z.device),
            )
            _lib.check(rc, "mb200_euclidean_eval")
            z = out
        return _like_input(state.pos, z if state.pos.ndim == 2 else z[0])


class GaussianEuclideanMetricSystem(EuclideanMetricSystem):
    """Euclidean system whose target density is given relative to the standard Gaussian measure
    (systems.py:369-474) -- "next" row N4: ``h1 = l(q)``, ``h2 = q.q/2 + p.M^-1 p/2`` and
    ``h2_flow`` is the exact rotation of ``(q, p)`` in the eigenbasis of ``M``.  The tractable-
    flow integrators (leapfrog, symmetric compositions) drive it through
    ``mb200_leapfrog_gaussian_euclidean``."""

    def h2(self, state):
        """``q.q/2 + p.M^-1 p/2`` (systems.py:450-453)."""
        pos = torch.as_tensor(state.pos)
        return super().h2(state) + _like_input(state.pos, 0.5 * (pos * pos).sum(-1))

    def dh2_dpos(self, state):
        """systems.py:460-462."""
        return state.pos

    def dh_dpos(self, state):
        return self.dh1_dpos(state) + state.pos

    def h(self, state):
        return self.h1(state) + self.h2(state)

    def _eig(self):
        """``(eigval, eigvec)`` of the metric as the reference obtains them: ``numpy.linalg.eigh``
        of the dense array (matrices.py:436-438); identity eigenvectors for identity / diagonal
        metrics (matrices.py:519-528, 743-749)."""
        m = self._metric
        if "eig" not in m._dev:
            m._dev["eig"] = np.linalg.eigh(m.array)
        return m._dev["eig"]

    def rotation_device(self, device, step_size, drift_coefficients):
        """Device operand ``rotation`` of ``mb200_leapfrog_gaussian_euclidean``."""
        m = self._metric
        if m.kind == METRIC_IDENTITY:
            return None
        if m.kind == METRIC_DIAGONAL:
            key = ("diag", str(device))
            if key not in m._dev:
                m._dev[key] = torch.as_tensor(np.ascontiguousarray(m.array), device=device)
            return m._dev[key]
        key = ("rot", str(device), float(step_size), tuple(float(c) for c in drift_coefficients))
        if key not in m._dev:
            # keep only the rotations of the most recent step sizes (adaptation visits many)
            stale = [k for k in m._dev if isinstance(k, tuple) and k and k[0] == "rot"]
            for k in stale[:-3]:
                del m._dev[k]
            eigval, u = self._eig()
            omega = 1.0 / eigval**0.5
            mats = []
            for c in drift_coefficients:
                t = float(c) * float(step_size)
                sn, cs = np.sin(omega * t), np.cos(omega * t)
                mats += [(u * cs) @ u.T, (u * (sn * omega)) @ u.T, -(u * (sn / omega)) @ u.T]
            m._dev[key] = torch.as_tensor(np.ascontiguousarray(np.stack(mats)), device=device)
        return m._dev[key]

    def h2_flow(self, state, dt):
        """Exact flow of ``h2`` over ``dt`` (systems.py:464-474), all chains in one launch."""
        from .integrators import _gaussian_flow  # noqa: PLC0415

        _gaussian_flow(self, state, dt)


def _col(dt):
    return dt[..., None] if isinstance(dt, torch.Tensor) and dt.ndim >= 1 else dt


class ConstrainedTractableFlowSystem(TractableFlowSystem):
    """systems.py:477-616."""

    def sample_momentum(self, state, rng):
        """Draw from N(0, M), then project onto the cotangent space (systems.py:613-616)."""
        mom = super().sample_momentum(state, rng)
        return self.project_onto_cotangent_space(mom, state)


class ConstrainedEuclideanMetricSystem(ConstrainedTractableFlowSystem, EuclideanMetricSystem):
    """Euclidean system subject to holonomic constraints (systems.py:619-873).

    ``constr`` must be the same ``Target`` instance as ``neg_log_dens`` (constrained targets
    carry their constraint function).  ``dens_wrt_hausdorff=False``: the target density is given
    with respect to the Lebesgue measure and ``h1`` / ``dh1_dpos`` carry ``log det gram / 2`` and
    its gradient through the constraint's matrix-Hessian product (systems.py:853-861, 1024-1031),
    fused into the kernels.
    """

    def __init__(self, neg_log_dens, constr=None, *, metric=None, dens_wrt_hausdorff=True,
                 grad_neg_log_dens=None, jacob_constr=None, backend=None):
        EuclideanMetricSystem.__init__(self, neg_log_dens, metric=metric,
                                       grad_neg_log_dens=grad_neg_log_dens, backend=backend)
        if constr is not None and constr is not neg_log_dens:
            raise ValueError("`constr` must be the target model passed as `neg_log_dens`.")
        if jacob_constr is not None:
            raise ValueError("The constraint Jacobian is fused into the kernels.")
        if neg_log_dens.n_constr < 1:
            raise ValueError(f"Target {neg_log_dens!r} defines no constraint function.")
        self.dens_wrt_hausdorff = bool(dens_wrt_hausdorff)

    def h(self, state):
        """``l(q) + p.M^-1 p/2`` (``dens_wrt_hausdorff=True``: systems.py:842-851, 187-196),
        evaluated by a zero-step launch of the constrained kernel."""
        pos, mom, _, single = _batched(state)
        n, dim = pos.shape
        dev = pos.device
        pos, mom = pos.contiguous(), mom.contiguous()
        h = torch.empty(n, dtype=torch.float64, device=dev)
        scratch_q, scratch_p = torch.empty_like(pos), torch.empty_like(mom)
        m = self._metric
        model = self._model(dev)
        rc = _lib.load().mb200_constrained_leapfrog_euclidean(
            _lib.ptr(pos), _lib.ptr(mom), _lib.ptr(scratch_q), _lib.ptr(scratch_p), None, n, dim,
            0.0, 0, 1, m.kind, _lib.ptr(m.inv_device(dev)), ctypes.byref(model), 0, 1e-9, 1e-8,
            1e10, 50, 10, 2e-8, _lib.ptr(h), None, None, None, _lib.current_stream_ptr(dev),
        )
        _lib.check(rc, "mb200_constrained_leapfrog_euclidean")
        return _like_input(state.pos, h[0] if single else h)

    def project_onto_cotangent_space(self, mom, state):
        """``mom - J^T (J M^-1 J^T)^-1 J M^-1 mom`` at ``state.pos`` (systems.py:863-873) for all
        chains in one launch (``mb200_project_onto_cotangent_space``)."""
        pos = state.pos
        single = pos.ndim == 1
        ref = mom
        pos_t = torch.as_tensor(pos)
        if pos_t.device.type != "cuda":
            pos_t = pos_t.to("cuda")
        pos_t = (pos_t[None] if single else pos_t).contiguous()
        mom_t = torch.as_tensor(mom).to(pos_t.device)
        mom_t = (mom_t[None] if mom_t.ndim == 1 else mom_t).contiguous()
        n, dim = pos_t.shape
        dev = pos_t.device
        out = torch.empty_like(mom_t)
        m = self._metric
        minv = None if m.kind == METRIC_IDENTITY else m.inv_device(dev)
        model = self._model(dev)
        rc = _lib.load().mb200_project_onto_cotangent_space(
            _lib.ptr(pos_t), _lib.ptr(mom_t), _lib.ptr(out), n, dim, m.kind, _lib.ptr(minv),
            ctypes.byref(model), _lib.current_stream_ptr(dev),
        )
        _lib.check(rc, "mb200_project_onto_cotangent_space")
        return _like_input(ref, out[0] if single else out)


class DenseConstrainedEuclideanMetricSystem(ConstrainedEuclideanMetricSystem):
    """systems.py:876-1031 (dense constraint Jacobian)."""

    def __init__(self, neg_log_dens, constr=None, *, metric=None, dens_wrt_hausdorff=True,
                 grad_neg_log_dens=None, jacob_constr=None, mhp_constr=None, backend=None):
        if mhp_constr is not None:
            raise ValueError("The constraint's matrix-Hessian product is fused into the kernels.")
        super().__init__(neg_log_dens, constr, metric=metric,
                         dens_wrt_hausdorff=dens_wrt_hausdorff,
                         grad_neg_log_dens=grad_neg_log_dens, jacob_constr=jacob_constr,
                         backend=backend)


class RiemannianMetricSystem(System):
    """Riemannian Hamiltonian system with a position-dependent metric (systems.py:1187-1402)."""

    def _workspace(self, n, dim, device):
        model = self._model(device)
        nbytes = int(_lib.load().mb200_implicit_workspace_bytes(n, dim, ctypes.byref(model)))
        key = ("ws", str(device))
        ws = self._dev.get(key)
        if ws is None or ws.numel() < nbytes:
            ws = torch.empty(max(nbytes, 8), dtype=torch.uint8, device=device)
            self._dev[key] = ws
        return ws

    def h(self, state):
        """l(q) + log|M(q)|/2 + p^T M(q)^-1 p / 2 (systems.py:1375-1390)."""
        pos, mom, _, single = _batched(state)
        n, dim = pos.shape
        dev = pos.device
        h = torch.empty(n, dtype=torch.float64, device=dev)
        status = torch.empty(n, dtype=torch.int32, device=dev)
        ws = self._workspace(n, dim, dev)
        model = self._model(dev)
        rc = _lib.load().mb200_hamiltonian_riemannian(
            _lib.ptr(pos.contiguous()), _lib.ptr(mom.contiguous()), n, dim, ctypes.byref(model),
            _lib.ptr(h), _lib.ptr(status), _lib.ptr(ws), ws.numel(), _lib.current_stream_ptr(dev),
        )
        _lib.check(rc, "mb200_hamiltonian_riemannian")
        return _like_input(state.pos, h[0] if single else h)

    def dh2_dmom(self, state):
        """``M(q)^-1 p`` (systems.py:1398-1399); not cached, as in the reference."""
        from .errors import LinAlgError  # noqa: PLC0415

        pos, mom, _, single = _batched(state)
        n, dim = pos.shape
        dev = pos.device
        vel = torch.empty((n, dim), dtype=torch.float64, device=dev)
        status = torch.empty(n, dtype=torch.int32, device=dev)
        model = self._model(dev)
        rc = _lib.load().mb200_dh_dmom_riemannian(
            _lib.ptr(pos.contiguous()), _lib.ptr(mom.contiguous()), _lib.ptr(vel), n, dim,
            ctypes.byref(model), _lib.ptr(status), _lib.current_stream_ptr(dev),
        )
        _lib.check(rc, "mb200_dh_dmom_riemannian")
        if bool((status != 0).any()):
            bad = int((status != 0).sum())
            raise LinAlgError(f"metric factorisation failed for {bad} of {n} chains")
        return _like_input(state.pos, vel[0] if single else vel)

    def sample_momentum(self, state, rng):
        """``metric(state).sqrt @ N(0, I)`` (systems.py:1401-1402): the factor of M(q) is built
        per chain on the device (``mb200_sample_momentum_riemannian``).  Chains whose metric
        cannot be built raise ``LinAlgError`` as in the reference."""
        from .errors import LinAlgError  # noqa: PLC0415
        from .transitions import _normals  # noqa: PLC0415

        pos = torch.as_tensor(state.pos)
        single = pos.ndim == 1
        if pos.device.type != "cuda":
            pos = pos.to("cuda")
        pos = (pos[None] if single else pos).contiguous()
        n, dim = pos.shape
        dev = pos.device
        z = _normals(rng, (n, dim), dev).contiguous()
        out = torch.empty_like(z)
        status = torch.empty(n, dtype=torch.int32, device=dev)
        model = self._model(dev)
        rc = _lib.load().mb200_sample_momentum_riemannian(
            _lib.ptr(pos), _lib.ptr(z), _lib.ptr(out), n, dim, ctypes.byref(model),
            _lib.ptr(status), _lib.current_stream_ptr(dev),
        )
        _lib.check(rc, "mb200_sample_momentum_riemannian")
        if bool((status != 0).any()):
            bad = int((status != 0).sum())
            raise LinAlgError(f"metric factorisation failed for {bad} of {n} chains")
        return _like_input(state.pos, out[0] if single else out)


class DenseRiemannianMetricSystem(RiemannianMetricSystem):
    """Dense position-dependent metric (systems.py:1710-1760): ``metric_func`` is a registered
    metric model -- ``mici_b200.targets.Rank1Metric`` (M(q) = B + c q q^T) or
    ``mici_b200.targets.HadamardMetric`` (M(q) = B + c (q q^T) o S, full rank).  Each chain's
    metric is factorised (Cholesky), inverted explicitly and differentiated through the model's
    VJP as the reference does (matrices.py:1161-1188, systems.py:1381-1399): in shared memory
    for D <= 160, in a per-CTA global workspace with DMMA-blocked routines beyond
    (csrc/dense_global.cuh)."""

    def __init__(self, neg_log_dens, metric_func, *, vjp_metric_func=None,
                 grad_neg_log_dens=None, backend=None):
        super().__init__(neg_log_dens, grad_neg_log_dens=grad_neg_log_dens, backend=backend)
        if not isinstance(metric_func, (Rank1Metric, HadamardMetric)):
            raise TypeError("`metric_func` must be a registered metric model "
                            "(Rank1Metric or HadamardMetric).")
        if vjp_metric_func is not None:
            raise ValueError("The metric VJP is fused into the kernels.")
        self.metric_model = metric_func
        self._rmetric_id = metric_func.rmetric_id
        self._rmetric_params = metric_func.params
        self._rmetric_aux = metric_func.aux


class SoftAbsRiemannianMetricSystem(RiemannianMetricSystem):
    """SoftAbs-regularised Hessian metric (systems.py:1763-1920)."""

    def __init__(self, neg_log_dens, *, grad_neg_log_dens=None, hess_neg_log_dens=None,
                 mtp_neg_log_dens=None, softabs_coeff=1.0, backend=None):
        super().__init__(neg_log_dens, grad_neg_log_dens=grad_neg_log_dens, backend=backend)
        if hess_neg_log_dens is not None or mtp_neg_log_dens is not None:
            raise ValueError("Hessian and MTP of the target are fused into the kernels.")
        if softabs_coeff <= 0:
            raise ValueError("softabs_coeff must be positive.")
        self.softabs_coeff = float(softabs_coeff)
        self._rmetric_id = RMETRIC_SOFTABS
        self._rmetric_params = (self.softabs_coeff,)
