import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from mici_b200 import engine, problems
prob = problems.make_problem("C1")
integ = engine.build_integrator(prob)
state = engine.build_state(prob, "cuda:0")
out = integ.step_n(state, 100, return_h=True); torch.cuda.synchronize()
d = out.h[4096:4096 + 64].reshape(16, 4).cpu()
print("warp(group*4+quarter): drift / reduce(+bar1) / kick(+bar2) cycles per step, MT")
for i in range(16): print(i, [round(float(x)) for x in d[i]])
