"""End-to-end (host buffers) time of one C1 bench step (50 leapfrog steps, 8192 chains) as a
function of the number of row chunks in Integrator.step_n_host.  CUDA events, median of reps.
Usage: python profiles/tools/e2e_chunks.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from mici_b200 import engine, problems

prob = problems.make_problem("C1")
integ = engine.build_integrator(prob)
dev = torch.device("cuda:0")
pos_h, mom_h = torch.as_tensor(prob.pos).pin_memory(), torch.as_tensor(prob.mom).pin_memory()
pos_o, mom_o = torch.empty_like(pos_h).pin_memory(), torch.empty_like(mom_h).pin_memory()
st_o = torch.empty(prob.n_chains, dtype=torch.int32).pin_memory()
L, reps = 50, 15
ref = integ.step_n(engine.build_state(prob, dev), L)
for chunks in (1, 2, 3, 4, 5, 6, 8, 12):
    def step():
        integ.step_n_host(pos_h, mom_h, L, out_pos=pos_o, out_mom=mom_o, out_status=st_o,
                          device=dev, n_chunks=chunks)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); step(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    # (the tensor-core kernel deals row tiles to its warps by batch size: split-k accumulators for
    # small launches, so chunked and whole-batch results agree to rounding, not bit for bit)
    torch.testing.assert_close(pos_o, ref.pos.cpu(), rtol=1e-9, atol=1e-10)
    torch.testing.assert_close(mom_o, ref.mom.cpu(), rtol=1e-9, atol=1e-10)
    assert int(st_o.abs().sum()) == 0
    ms = ts[len(ts) // 2]
    print(json.dumps({"n_chunks": chunks, "ms_median": ms, "ms_min": ts[0],
                      "e2e_steps_per_s": prob.n_chains * L / (ms * 1e-3)}))
