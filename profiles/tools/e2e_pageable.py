"""Host path with PAGEABLE buffers (NumPy-backed in, fresh pageable out) vs number of row blocks.
Usage: python profiles/tools/e2e_pageable.py"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from mici_b200 import engine, problems

prob = problems.make_problem("C1")
integ = engine.build_integrator(prob)
dev = torch.device("cuda:0")
pos_p, mom_p = torch.from_numpy(np.array(prob.pos)), torch.from_numpy(np.array(prob.mom))
L, reps = 50, 15
for chunks in (4, 6, 8, 12, 16):
    for _ in range(3):
        integ.step_n_host(pos_p, mom_p, L, device=dev, n_chunks=chunks)
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        integ.step_n_host(pos_p, mom_p, L, device=dev, n_chunks=chunks)
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    ts.sort()
    print(json.dumps({"n_chunks": chunks, "ms_median": ts[len(ts) // 2], "ms_min": ts[0],
                      "steps_per_s": prob.n_chains * L / (ts[len(ts) // 2] * 1e-3)}))
