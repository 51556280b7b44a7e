"""Small driver for ncu: a few launches of the C1 leapfrog kernel (no timing claims)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from mici_b200 import engine, problems
L = int(sys.argv[1]) if len(sys.argv) > 1 else 50
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
prob = problems.make_problem("C1")
integ = engine.build_integrator(prob)
state = engine.build_state(prob, "cuda:0")
for _ in range(reps):
    out = integ.step_n(state, L)
torch.cuda.synchronize()
print("ok", float(out.pos.abs().mean()))
