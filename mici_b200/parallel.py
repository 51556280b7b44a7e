"""Multi-GPU plumbing: chains shard embarrassingly -- nothing on the step path.

One process per GPU (``torch.distributed``: ``nccl`` on GPUs, ``gloo`` on CPU for tests).  The
reference's only parallelism is a process pool over chains with per-chain ``.npy`` traces
(samplers.py:668-772, 116-138); here each rank owns a contiguous block of chain rows and the
single collective is the gather of final states / traces at write-out.
"""

from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(n_chains: int, rank: int, world_size: int) -> tuple[int, int]:
    """Contiguous block ``[lo, hi)`` of chain rows owned by ``rank`` (sizes differ by <= 1)."""
    base, rem = divmod(n_chains, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_rows(array, rank: int, world_size: int):
    lo, hi = shard_bounds(array.shape[0], rank, world_size)
    return array[lo:hi]


def gather_rows(local: torch.Tensor, n_total: int, dst: int = 0, group=None):
    """Gather row blocks (as laid out by ``shard_bounds``) onto ``dst`` (a rank OF ``group``);
    returns the full ``[n_total, ...]`` tensor on ``dst`` and ``None`` elsewhere.  One
    collective."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    # dist.gather addresses the destination by GLOBAL rank; `dst` is a rank of `group`
    global_dst = dst if group is None else dist.get_global_rank(group, dst)
    sizes = [shard_bounds(n_total, r, world)[1] - shard_bounds(n_total, r, world)[0]
             for r in range(world)]
    max_rows = max(sizes)
    pad = torch.zeros((max_rows, *local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    if rank == dst:
        bufs = [torch.empty_like(pad) for _ in range(world)]
        dist.gather(pad, bufs, dst=global_dst, group=group)
        return torch.cat([b[:s] for b, s in zip(bufs, sizes)], 0)
    dist.gather(pad, None, dst=global_dst, group=group)
    return None


def gather_state(state, n_total: int, dst: int = 0, group=None):
    """Write-out: gather ``pos``, ``mom`` and ``status`` of a sharded state on ``dst`` with ONE
    collective: the three arrays travel packed as ``[rows, 2 D + 1]`` float64 (a status code is
    exact in a double)."""
    pos, mom = state.pos, state.mom
    n, dim = pos.shape
    st = state.status
    packed = torch.empty((n, 2 * dim + 1), dtype=torch.float64, device=pos.device)
    packed[:, :dim] = pos
    packed[:, dim:2 * dim] = mom
    packed[:, 2 * dim] = -1.0 if st is None else st.to(torch.float64)
    full = gather_rows(packed, n_total, dst, group)
    if full is None:
        return None
    out = {"pos": full[:, :dim].contiguous(), "mom": full[:, dim:2 * dim].contiguous()}
    if st is not None:
        out["status"] = full[:, 2 * dim].to(torch.int32)
    return out
