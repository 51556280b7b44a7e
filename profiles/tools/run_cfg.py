"""Small driver for ncu: a few launches of a config's integrator kernel (no timing claims)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from mici_b200 import engine, problems
cfg = sys.argv[1]; L = int(sys.argv[2]); reps = int(sys.argv[3]); n = int(sys.argv[4]) if len(sys.argv) > 4 else None
kw = {} if n is None else {"n_chains": n}
prob = problems.make_problem(cfg, **kw)
integ = engine.build_integrator(prob)
state = engine.build_state(prob, "cuda:0")
for _ in range(reps):
    out = integ.step_n(state, L)
torch.cuda.synchronize()
print("ok", float(out.pos.abs().mean()), float((out.status == 0).float().mean()))
