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
from .targets import RMETRIC_SOFTABS, Rank1Metric, Target

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
        model = self._model(dev)
        rc = lib.mb200_euclidean_eval(
            _lib.ptr(pos), _lib.ptr(mom), n, dim, self._metric.kind,
            _lib.ptr(self._metric.inv_device(dev)), ctypes.byref(model),
            _lib.ptr(out.get("nld")), _lib.ptr(out.get("grad")), _lib.ptr(out.get("vel")),
            _lib.ptr(out.get("kin")), _lib.current_stream_ptr(dev),
        )
        _lib.check(rc, "mb200_euclidean_eval")
        if single:
            out = {k: v[0] for k, v in out.items()}
        return out

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
        z = _normals(rng, tuple(pos.shape), pos.device).contiguous()
        m = self._metric
        if m.kind != METRIC_IDENTITY:
            n, dim = z.shape
            out = torch.empty_like(z)
            key = ("sqrt_t", str(z.device))
            if key not in m._dev:
                fac = m.sqrt if m.kind == METRIC_DIAGONAL else np.ascontiguousarray(m.sqrt.T)
                m._dev[key] = torch.as_tensor(fac, device=z.device).contiguous()
            model = self._model(z.device)
            rc = _lib.load().mb200_euclidean_eval(
                _lib.ptr(z), _lib.ptr(z), n, dim, m.kind, _lib.ptr(m._dev[key]),
                ctypes.byref(model), None, None, _lib.ptr(out), None,
                _lib.current_stream_ptr(z.device),
            )
            _lib.check(rc, "mb200_euclidean_eval")
            z = out
        return z if state.pos.ndim == 2 else z[0]


def _col(dt):
    return dt[..., None] if isinstance(dt, torch.Tensor) and dt.ndim >= 1 else dt


class ConstrainedTractableFlowSystem(TractableFlowSystem):
    """systems.py:477-616."""


class ConstrainedEuclideanMetricSystem(ConstrainedTractableFlowSystem, EuclideanMetricSystem):
    """Euclidean system subject to holonomic constraints (systems.py:619-873).

    ``constr`` must be the same ``Target`` instance as ``neg_log_dens`` (constrained targets
    carry their constraint function); ``dens_wrt_hausdorff`` must be ``True``.
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
        if not dens_wrt_hausdorff:
            raise NotImplementedError("Only `dens_wrt_hausdorff=True` is implemented.")
        self.dens_wrt_hausdorff = dens_wrt_hausdorff


class DenseConstrainedEuclideanMetricSystem(ConstrainedEuclideanMetricSystem):
    """systems.py:876-1031 (dense constraint Jacobian)."""

    def __init__(self, neg_log_dens, constr=None, *, metric=None, dens_wrt_hausdorff=True,
                 grad_neg_log_dens=None, jacob_constr=None, mhp_constr=None, backend=None):
        if mhp_constr is not None:
            raise ValueError("`mhp_constr` is only used with `dens_wrt_hausdorff=False`.")
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


class DenseRiemannianMetricSystem(RiemannianMetricSystem):
    """Dense position-dependent metric (systems.py:1710-1760): ``metric_func`` is a
    ``mici_b200.targets.Rank1Metric`` model (M(q) = B + c q q^T)."""

    def __init__(self, neg_log_dens, metric_func, *, vjp_metric_func=None,
                 grad_neg_log_dens=None, backend=None):
        super().__init__(neg_log_dens, grad_neg_log_dens=grad_neg_log_dens, backend=backend)
        if not isinstance(metric_func, Rank1Metric):
            raise TypeError("`metric_func` must be a registered metric model (Rank1Metric).")
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
