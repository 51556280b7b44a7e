"""Throughput of the fused dynamic-HMC (NUTS) kernel on C1 (8192 chains, D=128, dense metric):
leapfrog steps per second inside whole transitions, device-generated uniforms.
Usage: python profiles/tools/bench_nuts.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from mici_b200 import engine, problems, transitions

for cfg, kw, eps, depth in (("C1", {}, 0.01, 6), ("C1", {}, 0.01, 8), ("C1", {"dim": 64}, 0.01, 8),
                            ("C0", {"n_chains": 8192, "dim": 10}, 0.1, 8)):
    prob = problems.make_problem(cfg, **kw)
    prob.step_size = eps
    integ = engine.build_integrator(prob)
    state = engine.build_state(prob, "cuda:0")
    gen = torch.Generator(device="cuda:0"); gen.manual_seed(0)
    tr = transitions.MultinomialDynamicIntegrationTransition(integ.system, integ, max_tree_depth=depth)
    mom = transitions.IndependentMomentumTransition(integ.system)
    for _ in range(2):
        state, _ = mom.sample(state, gen); state, st = tr.sample(state, gen)
    torch.cuda.synchronize()
    reps, ms, steps = 5, [], 0
    for _ in range(reps):
        state, _ = mom.sample(state, gen)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); state, st = tr.sample(state, gen); b.record(); torch.cuda.synchronize()
        ms.append(a.elapsed_time(b)); steps += int(st["n_step"].sum().item())
    tot = sum(ms)
    print(json.dumps({"config": cfg, "chains": prob.n_chains, "dim": prob.dim, "max_tree_depth": depth,
                      "ms_per_transition": tot / reps, "mean_n_step": steps / reps / prob.n_chains,
                      "mean_tree_depth": float(st["tree_depth"].double().mean()),
                      "leapfrog_steps_per_s": steps / (tot * 1e-3),
                      "accept_stat": float(st["accept_stat"].mean())}))
