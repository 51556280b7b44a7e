"""Small driver for ncu: a few NUTS transitions on C1 (no timing claims).
Usage: python profiles/tools/run_nuts.py [n_chains] [max_tree_depth] [reps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from mici_b200 import engine, problems, transitions
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
depth = int(sys.argv[2]) if len(sys.argv) > 2 else 5
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
prob = problems.make_problem("C1", n_chains=n); prob.step_size = 0.01
integ = engine.build_integrator(prob); state = engine.build_state(prob, "cuda:0")
gen = torch.Generator(device="cuda:0"); gen.manual_seed(0)
tr = transitions.MultinomialDynamicIntegrationTransition(integ.system, integ, max_tree_depth=depth)
for _ in range(reps):
    state, st = tr.sample(state, gen)
torch.cuda.synchronize()
print("ok", float(st["n_step"].double().mean()))
