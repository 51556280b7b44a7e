"""C4 (dense Riemannian D=512, 8192 chains per GPU) under torchrun: weak scaling, max over ranks."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, torch.distributed as dist
from mici_b200 import engine, problems, parallel
rank = int(os.environ.get("RANK", 0)); local = int(os.environ.get("LOCAL_RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
dev = torch.device("cuda", local); torch.cuda.set_device(dev)
if world > 1: dist.init_process_group("nccl", device_id=dev)
prob = problems.make_problem("C4", n_chains=8192, seed=problems.BASE_SEED + 4 + 1000 * rank)
integ = engine.build_integrator(prob); state = engine.build_state(prob, dev)
out = integ.step_n(state, 1); torch.cuda.synchronize(dev)
if world > 1: dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
K = 3
e0.record()
for _ in range(K): out = integ.step_n(state, 1)
e1.record(); torch.cuda.synchronize(dev)
ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
done = torch.tensor([float(out.n_done.sum().item())], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX); dist.all_reduce(done)
    g = parallel.gather_state(out, 8192 * world, dst=0)   # write-out: the one collective
if rank == 0:
    print(json.dumps({"config": "C4", "n_gpus": world, "chains": 8192 * world, "dim": 512, "ms_per_batch_step": ms.item() / K,
                      "leapfrog_steps_per_s": done.item() / (ms.item() / K * 1e-3)}))
if world > 1: dist.destroy_process_group()
