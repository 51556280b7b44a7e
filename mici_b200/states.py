"""Batched Markov chain state -- the boundary type of the engine.

Mirrors the semantics of the reference ``ChainState`` (``src/mici/states.py:160-305``):
keyword-constructed variables (``pos``, ``mom``, ``dir``), attribute access, ``copy()`` giving
independent variable storage, optional read-only flag, ``in`` test.  What differs is the
storage: ``pos`` / ``mom`` are fp64 CUDA ``torch.Tensor`` s of shape ``[n_chains, dim]``
(row-major) and ``dir`` is ``+-1`` as a Python int (all chains) or an int32 tensor
``[n_chains]``.  The reference's memoising dict cache (states.py:37-157) is replaced by explicit
device buffers that the kernels recompute or carry: the per-chain
outcome of the last integrator call is exposed as ``state.status`` / ``state.n_done``.
"""

from __future__ import annotations

import copy as _copy

from .errors import ReadOnlyStateError

_AUX = ("status", "n_done", "h", "solver_iters")


class ChainState:
    """Batched chain state: ``ChainState(pos=..., mom=..., dir=1)``."""

    def __init__(self, *, _read_only=False, _aux=None, **variables):
        for name in variables:
            if name.startswith("_") or name == "copy":
                raise ValueError(f"Invalid state variable name {name!r}.")
        self.__dict__["_variables"] = variables
        self.__dict__["_aux"] = {} if _aux is None else _aux
        self.__dict__["_read_only"] = _read_only

    def __getattr__(self, name):
        d = self.__dict__
        if name in d.get("_variables", {}):
            return d["_variables"][name]
        if name in _AUX:
            return d.get("_aux", {}).get(name)
        raise AttributeError(f"'{type(self).__name__}' object has no attribute '{name}'")

    def __setattr__(self, name, value):
        if self._read_only:
            raise ReadOnlyStateError("ChainState instance is read-only.")
        if name in self._variables:
            self._variables[name] = value
            self._aux.clear()  # derived quantities depend on the variables
        elif name in _AUX:
            self._aux[name] = value
        else:
            super().__setattr__(name, value)

    def __contains__(self, name):
        return name in self._variables

    @property
    def n_chains(self):
        pos = self._variables["pos"]
        return 1 if pos.ndim == 1 else pos.shape[0]

    @property
    def dim(self):
        return self._variables["pos"].shape[-1]

    def copy(self, *, read_only=False):
        """Deep copy: variable tensors are cloned (states.py:263-279)."""

        def cp(v):
            if hasattr(v, "clone"):
                return v.clone()
            return _copy.copy(v)

        return type(self)(
            _read_only=read_only,
            _aux=dict(self._aux),
            **{k: cp(v) for k, v in self._variables.items()},
        )

    def __str__(self):
        return "(\n " + ",\n ".join(f"{k}={v}" for k, v in self._variables.items()) + ")"

    def __repr__(self):
        return type(self).__name__ + str(self)

    def __getstate__(self):
        return {"variables": self._variables, "aux": self._aux, "read_only": self._read_only}

    def __setstate__(self, state):
        self.__dict__["_variables"] = state["variables"]
        self.__dict__["_aux"] = state["aux"]
        self.__dict__["_read_only"] = state["read_only"]
