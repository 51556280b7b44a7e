"""mici_b200 -- B200-native batched-chain Hamiltonian integrator engine.

Drop-in for the ``Integrator.step`` hot path of matt-graham/mici (explicit leapfrog, implicit
generalised leapfrog, constrained leapfrog), evaluated over thousands of independent chains
per kernel launch.  Hand-written sm_100a CUDA behind a C ABI (``include/mici_b200.h``,
``libmici_b200.so``), bound here with ctypes; PyTorch tensors are only the device buffers.
There is no CPU fallback: without the built library every compute call raises.
"""

from . import (
    adapters,
    errors,
    integrators,
    problems,
    samplers,
    solvers,
    stagers,
    states,
    systems,
    targets,
    transitions,
)
from .states import ChainState

__all__ = [
    "ChainState",
    "adapters",
    "errors",
    "integrators",
    "problems",
    "samplers",
    "solvers",
    "stagers",
    "states",
    "systems",
    "targets",
    "transitions",
]
__version__ = "0.1.0"
