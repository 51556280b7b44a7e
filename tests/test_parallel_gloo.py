"""CPU tests of the multi-GPU plumbing with the gloo backend, world_size 2: chains shard into
contiguous row blocks, nothing is exchanged on the step path, one gather at write-out."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mici_b200 import parallel, problems
from mici_b200.states import ChainState


def test_shard_bounds_partition_all_rows():
    for n in (0, 1, 7, 8, 8192, 65536 + 3):
        for world in (1, 2, 3, 8):
            spans = [parallel.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c and b >= a and d >= c
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_total, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    prob = problems.make_problem("C1", n_chains=n_total, dim=16)
    lo, hi = parallel.shard_bounds(n_total, rank, world)
    # every rank "steps" only its own rows (stand-in arithmetic: the kernels need a GPU);
    # what is tested is the sharding + write-out contract
    pos = torch.as_tensor(prob.pos[lo:hi]) * 2.0
    mom = torch.as_tensor(prob.mom[lo:hi]) + 1.0
    st = ChainState(pos=pos, mom=mom, dir=1)
    st.status = torch.full((hi - lo,), rank, dtype=torch.int32)
    gathered = parallel.gather_state(st, n_total, dst=0)
    if rank == 0:
        np.savez(out_path, pos=gathered["pos"].numpy(), mom=gathered["mom"].numpy(),
                 status=gathered["status"].numpy())
    else:
        assert gathered is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [10, 33])
def test_gather_at_write_out_gloo_world2(tmp_path, n_total):
    world = 2
    out = str(tmp_path / "gathered.npz")
    mp.spawn(_worker, args=(world, _free_port(), n_total, out), nprocs=world, join=True)
    g = np.load(out)
    prob = problems.make_problem("C1", n_chains=n_total, dim=16)
    np.testing.assert_array_equal(g["pos"], prob.pos * 2.0)
    np.testing.assert_array_equal(g["mom"], prob.mom + 1.0)
    lo1, _ = parallel.shard_bounds(n_total, 1, world)
    assert (g["status"][:lo1] == 0).all() and (g["status"][lo1:] == 1).all()


def _trace_worker(rank, world, port, n_total, n_iter, out_dir):
    from mici_b200 import traces

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = parallel.shard_bounds(n_total, rank, world)
    buf = traces.TraceBuffer(n_iter, hi - lo, (3,), device="cpu")
    stat = traces.TraceBuffer(n_iter, hi - lo, (), device="cpu")
    for it in range(n_iter):
        chains = torch.arange(lo, hi, dtype=torch.float64)
        buf.append(chains[:, None] * 100 + it + torch.arange(3, dtype=torch.float64)[None] / 10)
        stat.append(chains + it / 100)
    full = traces.gather_traces({"pos": buf.data, "accept_stat": stat.data}, n_total, dst=0)
    if rank == 0:
        traces.write_chain_traces(out_dir, "trace", full)
    else:
        assert full is None
    dist.barrier()
    dist.destroy_process_group()


def test_trace_write_out_gloo_world2_reference_file_layout(tmp_path):
    """Row N2: one gather at write-out, then one `{prefix}_{chain}_{key}.npy` per chain and key
    (reference samplers.py:104-138), readable as a memmap."""
    n_total, n_iter = 7, 5
    mp.spawn(_trace_worker, args=(2, _free_port(), n_total, n_iter, str(tmp_path)), nprocs=2,
             join=True)
    for c in range(n_total):
        pos = np.load(tmp_path / f"trace_{c}_pos.npy", mmap_mode="r")
        acc = np.load(tmp_path / f"trace_{c}_accept_stat.npy", mmap_mode="r")
        assert pos.shape == (n_iter, 3) and acc.shape == (n_iter,)
        for it in range(n_iter):
            np.testing.assert_allclose(pos[it], c * 100 + it + np.arange(3) / 10)
            np.testing.assert_allclose(acc[it], c + it / 100)
