"""Device-time throughput of the non-headline configs (C0, C2, C3, C4) -- CUDA events, warm.
Usage: python profiles/tools/bench_configs.py [C2 C3 ...]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from mici_b200 import engine, problems

SPECS = {  # config -> (problem kwargs, leapfrog steps per launch, reps)
    "C0": ({}, 50, 10),
    "C2": ({}, 2, 5),
    "C3": ({}, 50, 10),
    "C4": ({"n_chains": 8192}, 1, 3),
    "C4small": ({"n_chains": 2048, "dim": 64}, 2, 5),
    "C6": ({}, 2, 5),
}
for name in (sys.argv[1:] or ["C0", "C2", "C3", "C4"]):
    kw, L, reps = SPECS[name]
    prob = problems.make_problem(name.replace("small", ""), **kw)
    integ = engine.build_integrator(prob)
    state = engine.build_state(prob, "cuda:0")
    out = integ.step_n(state, L); torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in ev:
        a.record(); out = integ.step_n(state, L); b.record()
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in ev)[reps // 2]
    ok = (out.status == 0).float().mean().item()
    done = int(out.n_done.sum().item())
    res = {"config": name, "chains": prob.n_chains, "dim": prob.dim, "steps_per_launch": L,
           "ms_per_launch": ms, "leapfrog_steps_per_s": done / (ms * 1e-3), "ok_fraction": ok}
    if out.solver_iters is not None:
        res["mean_solver_iters"] = out.solver_iters.double().mean(0).tolist() if out.solver_iters.ndim > 1 else out.solver_iters.double().mean().item()
    print(json.dumps(res))
