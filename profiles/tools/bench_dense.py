"""Per-launch CUDA-event times of the dense Riemannian workloads (C4 dense / C4 low-rank / C5).
Usage: python profiles/tools/bench_dense.py [n_chains] [names...]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from mici_b200 import engine, problems

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
names = sys.argv[2:] or ["C4_low_rank", "C4", "C4_low_rank", "C5"]
for name in names:
    cfg = name.split("_")[0]
    prob = problems.make_problem(cfg, n_chains=n)
    if name.endswith("low_rank"):
        prob.metric_params = dict(prob.metric_params, force_low_rank_form=True)
    integ = engine.build_integrator(prob)
    state = engine.build_state(prob, "cuda:0")
    times = []
    for _ in range(4):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); out = integ.step_n(state, 1); b.record(); torch.cuda.synchronize()
        times.append(a.elapsed_time(b))
    print(json.dumps({"name": name, "chains": n, "ms": times,
                      "steps_per_s": n / (min(times) * 1e-3),
                      "ok": float((out.status == 0).float().mean())}))
