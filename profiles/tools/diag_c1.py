"""Per-launch timings of the C1 kernel with/without L2 flush + clock samples (diagnostic)."""
import sys, os, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, pynvml
from mici_b200 import engine, problems
pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
samples = []; stop = False
def sampler():
    while not stop:
        samples.append((time.perf_counter(), pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM),
                        pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0, pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)))
        time.sleep(0.002)
th = threading.Thread(target=sampler); th.start()
L = int(sys.argv[1]) if len(sys.argv) > 1 else 50
prob = problems.make_problem("C1")
integ = engine.build_integrator(prob)
state = engine.build_state(prob, "cuda:0")
flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda:0")
for _ in range(3): integ.step_n(state, L)
torch.cuda.synchronize()
for mode in ("noflush", "flush", "noflush_long"):
    K = 200 if mode == "noflush_long" else 20
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    t0 = time.perf_counter()
    for i in range(K):
        if mode == "flush": flush.fill_(1.0)
        ev[i][0].record(); integ.step_n(state, L); ev[i][1].record()
    thost = time.perf_counter() - t0
    torch.cuda.synchronize()
    ts = [a.elapsed_time(b) for a, b in ev]
    print(mode, "host enqueue ms/iter %.3f" % (thost / K * 1e3), "launch ms: min %.3f med %.3f max %.3f" % (min(ts), sorted(ts)[K // 2], max(ts)),
          "first5", ["%.3f" % t for t in ts[:5]], "last3", ["%.3f" % t for t in ts[-3:]])
stop = True; th.join()
clk = [s[1] for s in samples]; pw = [s[2] for s in samples]
print("clock samples", len(clk), "min", min(clk), "median", sorted(clk)[len(clk)//2], "max", max(clk), "power max", max(pw), "reasons", sorted(set(s[3] for s in samples)))
