"""Tiny invocation of every kernel family (for compute-sanitizer memcheck / racecheck)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from mici_b200 import engine, problems, transitions
cases = [("C1", {"n_chains": 70}, 3), ("C1", {"n_chains": 9, "dim": 11}, 2), ("C1", {"n_chains": 5, "dim": 20, "integrator": "bcss3"}, 2),
         ("C2", {"n_chains": 3, "dim": 8}, 1), ("C3", {"n_chains": 40}, 2), ("S1", {"n_chains": 6, "dim": 10}, 2),
         ("C4", {"n_chains": 3, "dim": 16}, 1), ("C4", {"n_chains": 2, "dim": 200}, 1)]
for cfg, kw, L in cases:
    prob = problems.make_problem(cfg, **kw)
    integ = engine.build_integrator(prob)
    state = engine.build_state(prob, "cuda:0")
    out = integ.step_n(state, L, return_h=True)
    torch.cuda.synchronize()
    print(cfg, kw, "ok", bool(torch.isfinite(out.pos).all()), out.status.tolist()[:4])
prob = problems.make_problem("C1", n_chains=20, dim=16)
integ = engine.build_integrator(prob); state = engine.build_state(prob, "cuda:0")
gen = torch.Generator(device="cuda:0"); gen.manual_seed(1)
final, stats, _ = transitions.sample_hmc(integ.system, integ, state, gen, 2, 3)
torch.cuda.synchronize(); print("hmc ok", float(stats["accept_stat"].mean()))
