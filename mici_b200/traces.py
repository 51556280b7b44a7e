"""Trace / statistics write-out -- "next" row N2 of SURVEY.md 8(f).

The reference writes one ``.npy`` file per chain and traced variable, named
``{prefix}_{chain_index}_{key}.npy`` (``samplers.py:104-113``, opened with
``np.lib.format.open_memmap``: ``samplers.py:116-138``), which is what
``interop.convert_to_inference_data`` (``interop.py:54-96``) reads back.  Here the traces of all
chains of a rank live in one device buffer ``[n_iter, n_chains_local, ...]``; at write-out they
are gathered onto rank 0 with ONE collective (``torch.distributed.gather`` over NCCL / NVLink, or
gloo in the CPU tests) -- the only communication of a run -- and written in the reference's
per-chain layout so that downstream tooling is unchanged.
"""

from __future__ import annotations

from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

from . import parallel


def _valid_filename(string):
    """samplers.py:87-101."""
    return "".join(c for c in string if (c.isalnum() or c in "._- "))


def memmap_filenames(dir_path, prefix, key, indices):
    """samplers.py:104-113."""
    key_str = _valid_filename(str(key))
    return [Path(dir_path) / f"{prefix}_{index}_{key_str}.npy" for index in indices]


class TraceBuffer:
    """Device ring of per-iteration values for every local chain: ``[n_iter, n_chains, ...]``."""

    def __init__(self, n_iter, n_chains, item_shape=(), dtype=torch.float64, device="cuda"):
        self.data = torch.empty((n_iter, n_chains, *item_shape), dtype=dtype, device=device)
        self.n_written = 0

    def append(self, values):
        self.data[self.n_written].copy_(values)
        self.n_written += 1


def gather_traces(local, n_chains_total, dst=0, group=None):
    """Gather ``{key: [n_iter, n_local, ...]}`` onto ``dst`` as ``{key: [n_total, n_iter, ...]}``
    (chain-major, the reference's trace layout: one array per chain).  One collective per key;
    without an initialised process group the local data is returned."""
    out = {}
    distributed = dist.is_available() and dist.is_initialized()
    for key, val in local.items():
        chain_major = val.transpose(0, 1).contiguous()  # [n_local, n_iter, ...]
        if distributed:
            full = parallel.gather_rows(chain_major, n_chains_total, dst=dst, group=group)
        else:
            full = chain_major
        out[key] = full
    if distributed and dist.get_rank(group) != dst:
        return None
    return out


def write_chain_traces(dir_path, prefix, traces, chain_indices=None):
    """Write ``{key: [n_chains, n_iter, ...]}`` as one ``{prefix}_{index}_{key}.npy`` per chain and
    key, memory-mappable exactly like the reference's files.  Returns ``{key: [paths]}``."""
    dir_path = Path(dir_path)
    dir_path.mkdir(parents=True, exist_ok=True)
    paths = {}
    for key, val in traces.items():
        arr = val.detach().cpu().numpy() if isinstance(val, torch.Tensor) else np.asarray(val)
        idx = list(range(arr.shape[0])) if chain_indices is None else list(chain_indices)
        files = memmap_filenames(dir_path, prefix, key, idx)
        for f, chain in zip(files, arr):
            mm = np.lib.format.open_memmap(f, dtype=chain.dtype, mode="w+", shape=chain.shape)
            mm[:] = chain
            mm.flush()
            del mm
        paths[key] = files
    return paths
