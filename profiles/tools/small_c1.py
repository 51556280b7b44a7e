import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, numpy as np
from mici_b200 import engine, problems
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
L = int(sys.argv[2]) if len(sys.argv) > 2 else 1
prob = problems.make_problem("C1", n_chains=n)
integ = engine.build_integrator(prob)
state = engine.build_state(prob, "cuda:0")
out = integ.step_n(state, L, return_h=(len(sys.argv) > 3)); torch.cuda.synchronize()
from oracle import drivers
ref = drivers.oracle_run(prob, L)
print("maxdiff", np.abs(out.pos.cpu().numpy() - ref["pos"]).max(), np.abs(out.mom.cpu().numpy() - ref["mom"]).max())
