"""Adaptive warm-up with chains sharded over ranks (torchrun, NCCL): every rank adapts on its own
chains, `finalize` merges over all ranks with one all_gather; the result must equal a single-rank
run over ALL chains (same per-chain NumPy streams) and be identical on every rank.
Usage: python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 profiles/tools/adapt_multi.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch, torch.distributed as dist
from mici_b200 import adapters, engine, parallel, problems, transitions

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
dev = torch.device("cuda", torch.cuda.current_device())
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
N, SEED = 64, 5
prob = problems.make_problem("C1", n_chains=N, dim=24)


def run(lo, hi, group):
    integ = engine.build_integrator(prob)
    state = engine.build_state(prob, dev, chains=slice(lo, hi))
    rngs = [np.random.default_rng([SEED, i]) for i in range(lo, hi)]
    ads = [adapters.DualAveragingStepSizeAdapter(), adapters.OnlineCovarianceMetricAdapter()]
    final, stats, _ = transitions.sample_chains(integ.system, integ, state, rngs, 30, 3, n_step=5,
                                                adapters=ads, trace_pos=False, group=group)
    return integ.step_size, integ.system.metric.array, final


lo, hi = parallel.shard_bounds(N, rank, world)
eps, metric, final = run(lo, hi, None)
out = {"rank": rank, "world": world, "step_size": eps, "metric_trace": float(np.trace(metric))}
if world > 1:
    both = [None] * world
    dist.all_gather_object(both, (eps, metric))
    assert all(b[0] == both[0][0] and np.array_equal(b[1], both[0][1]) for b in both), "ranks disagree"
if rank == 0:
    eps1, metric1, _ = run(0, N, False)  # all chains on one rank, no collective
    out["single_rank_step_size"] = eps1
    out["rel_diff_step_size"] = abs(eps - eps1) / eps1
    out["max_rel_diff_metric"] = float(np.max(np.abs(metric - metric1) / np.abs(metric1).max()))
    print(json.dumps(out))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
