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
# --- entry points added late in round 1
from mici_b200 import adapters
for cfg, kw in (("C1", {"n_chains": 9, "dim": 12}), ("C1", {"n_chains": 5, "dim": 70, "metric_kind": "diagonal"}),
                ("C1", {"n_chains": 13, "dim": 128})):
    prob = problems.make_problem(cfg, **kw); prob.step_size = 0.1
    integ = engine.build_integrator(prob); state = engine.build_state(prob, "cuda:0")
    for cls in (transitions.MultinomialDynamicIntegrationTransition, transitions.SliceDynamicIntegrationTransition):
        tr = cls(integ.system, integ, max_tree_depth=4)
        st, stats = tr.sample(state, gen)
    torch.cuda.synchronize(); print("nuts ok", cfg, kw, stats["n_step"].tolist()[:4])
prob = problems.make_problem("C1", n_chains=11, dim=20)
integ = engine.build_integrator(prob); state = engine.build_state(prob, "cuda:0")
final, stats, _ = transitions.sample_chains(integ.system, integ, state, gen, 6, 2, n_step_range=(1, 4),
                                            adapters=[adapters.DualAveragingStepSizeAdapter(),
                                                      adapters.OnlineCovarianceMetricAdapter()])
torch.cuda.synchronize(); print("adaptive ok", integ.step_size)
for cfg, kw in (("C3", {"n_chains": 10}), ("C2", {"n_chains": 3, "dim": 6}), ("C4", {"n_chains": 3, "dim": 12})):
    prob = problems.make_problem(cfg, **kw)
    integ = engine.build_integrator(prob); state = engine.build_state(prob, "cuda:0")
    mom = integ.system.sample_momentum(state, gen)
    integ.step_size = torch.full((prob.n_chains,), prob.step_size, device="cuda:0", dtype=torch.float64)
    out = integ.step_n(state, torch.arange(prob.n_chains, device="cuda:0", dtype=torch.int32) % 3)
    torch.cuda.synchronize(); print("per-chain ok", cfg, bool(torch.isfinite(mom).all()), out.n_done.tolist()[:4])
for mk in ("identity", "diagonal", "dense"):
    prob = problems.make_problem("G1", n_chains=5, dim=18, metric_kind=mk)
    integ = engine.build_integrator(prob); state = engine.build_state(prob, "cuda:0")
    out = integ.step_n(state, 2, return_h=True)
    torch.cuda.synchronize(); print("gaussian ok", mk, bool(torch.isfinite(out.h).all()))
prob = problems.make_problem("C1", n_chains=130, dim=128)
integ = engine.build_integrator(prob)
p_h, m_h = torch.as_tensor(prob.pos).pin_memory(), torch.as_tensor(prob.mom).pin_memory()
q, p, s = integ.step_n_host(p_h, m_h, 2, device="cuda:0", n_chunks=3)
print("host path ok", bool(torch.isfinite(q).all()), int(s.abs().sum()))
