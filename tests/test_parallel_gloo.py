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


class _FakeSystem:
    """Stand-in for the device system: records the assigned metric (the kernels need a GPU)."""

    metric = None

    def sample_momentum(self, state, rngs):
        return torch.zeros_like(state.pos)


class _FakeTransition:
    def __init__(self):
        self.system = _FakeSystem()
        self.integrator = type("I", (), {"step_size": None})()


def _adapt_worker(rank, world, port, n_total, n_iter, out_path):
    from mici_b200 import adapters

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = parallel.shard_bounds(n_total, rank, world)
    draws = np.random.default_rng(5).standard_normal((n_iter, n_total, 6)) * np.arange(1, 7)
    accept = np.random.default_rng(6).uniform(0.3, 1.0, (n_iter, n_total))
    state = ChainState(pos=torch.as_tensor(draws[0, lo:hi]), mom=torch.zeros(hi - lo, 6), dir=1)
    var_ad, cov_ad = adapters.OnlineVarianceMetricAdapter(), adapters.OnlineCovarianceMetricAdapter()
    da = adapters.DualAveragingStepSizeAdapter(
        log_step_size_reducer=adapters.geometric_mean_log_step_size_reducer)
    tr_v, tr_c, tr_d = _FakeTransition(), _FakeTransition(), _FakeTransition()
    st_v, st_c = var_ad.initialize(state, tr_v), cov_ad.initialize(state, tr_c)
    st_d = {"iter": 0, "smoothed_log_step_size": torch.zeros(hi - lo, dtype=torch.float64),
            "adapt_stat_error": torch.zeros(hi - lo, dtype=torch.float64),
            "log_step_size_reg_target": torch.full((hi - lo,), np.log(10 * 0.25),
                                                   dtype=torch.float64)}
    for it in range(n_iter):
        state.pos = torch.as_tensor(draws[it, lo:hi])
        stats = {"accept_stat": torch.as_tensor(accept[it, lo:hi])}
        var_ad.update(st_v, state, stats, tr_v)
        cov_ad.update(st_c, state, stats, tr_c)
        da.update(st_d, state, stats, tr_d)
    var_ad.finalize(st_v, state, tr_v, None)
    cov_ad.finalize(st_c, state, tr_c, None)
    da.finalize(st_d, state, tr_d, None)
    np.savez(out_path.format(rank=rank), diag=tr_v.system.metric, dense=tr_c.system.metric.array,
             dense_inv=tr_c.system.metric.inv, step_size=tr_d.integrator.step_size)
    dist.barrier()
    dist.destroy_process_group()


def test_adapter_finalize_merges_all_ranks_gloo_world2(tmp_path):
    """Row N3 across ranks: every rank adapts on its own chains; ``finalize`` merges the
    per-rank moments / log step sizes with one all_gather and every rank ends with the same
    parameters -- equal to the oracle's chain-by-chain merge (adapters.py:375-390, 471-516,
    603-648) over ALL chains."""
    from oracle import mici_oracle as mo

    world, n_total, n_iter = 2, 9, 12
    out = str(tmp_path / "adapt_{rank}.npz")
    mp.spawn(_adapt_worker, args=(world, _free_port(), n_total, n_iter, out), nprocs=world,
             join=True)
    g0, g1 = np.load(out.format(rank=0)), np.load(out.format(rank=1))
    for k in g0.files:
        np.testing.assert_array_equal(g0[k], g1[k])
    draws = np.random.default_rng(5).standard_normal((n_iter, n_total, 6)) * np.arange(1, 7)
    accept = np.random.default_rng(6).uniform(0.3, 1.0, (n_iter, n_total))

    class Ctx:
        metric, step_size = None, None

    ov, oc = mo.OnlineVarianceOracle(), mo.OnlineCovarianceOracle()
    od = mo.DualAveragingOracle(log_step_size_reducer="geometric_mean_log_step_size_reducer")
    sv, sc, sd = [], [], []
    for c in range(n_total):
        a, b = ov.initialize(draws[0, c], None, 1, None), oc.initialize(draws[0, c], None, 1, None)
        d = {"iter": 0, "smoothed_log_step_size": 0.0, "adapt_stat_error": 0.0,
             "log_step_size_reg_target": np.log(10 * 0.25)}
        ctx = Ctx()
        for it in range(n_iter):
            ov.update(a, draws[it, c], None, None)
            oc.update(b, draws[it, c], None, None)
            od.update(d, draws[it, c], {"accept_stat": accept[it, c]}, ctx)
        sv.append(a), sc.append(b), sd.append(d)
    cv, cc, cd = Ctx(), Ctx(), Ctx()
    ov.finalize(sv, cv), oc.finalize(sc, cc), od.finalize(sd, cd)
    np.testing.assert_allclose(g0["diag"], cv.metric.diagonal, rtol=1e-12)
    np.testing.assert_allclose(g0["dense"], cc.metric.array, rtol=1e-10, atol=1e-14)
    np.testing.assert_allclose(g0["dense_inv"], cc.metric.inv_array, rtol=1e-11, atol=1e-14)
    assert float(g0["step_size"]) == pytest.approx(cd.step_size, rel=1e-13)
