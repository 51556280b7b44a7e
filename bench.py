#!/usr/bin/env python
"""Benchmark of the hot path: batched ``Integrator.step`` on B200.

    python bench.py --gpus N --steps K --warmup W          # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K --warmup W   # CPU reference arm

Workload (BASELINE.json configs[1], "C1"): dense-mass Euclidean leapfrog on Neal's funnel,
D = 128, 8192 chains per GPU, step size 0.01.  One bench "step" = one launch of
``LEAPFROG_PER_LAUNCH`` fused leapfrog steps over the whole batch (the trajectory loop of
``transitions.py:289-291``).  Metric: aggregate leapfrog steps / s = chains x leapfrog steps /
time.  Prints ONE JSON line (rank 0).
"""

from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "leapfrog steps/sec (aggregate over chains)"
UNIT = "leapfrog steps/s"
WORKLOAD = "C1: dense-mass Euclidean leapfrog, Neal's funnel D=128, 8192 chains per GPU"
N_CHAINS = 8192
DIM = 128
LEAPFROG_PER_LAUNCH = 50
FP64_PEAK_TFLOPS = 37.1  # measured DMMA peak on this pool's B200 (profiles/r01_fp64_peak.txt)
# dram__bytes_read.sum + dram__bytes_write.sum of one launch of the dominant kernel at this
# workload, from the `ncu --set full` capture summarised in
# profiles/r01_c1_dmma_v3_ncu_summary.txt (16 970 496 B read, 0 B written: the outputs are still
# in L2 when the kernel ends).  Algorithmic bytes per launch: 8192 x 4096 B = 33.5 MB.
NCU_DRAM_TRAFFIC_BYTES_PER_LAUNCH = 16970496
HBM_FALLBACK_GBS = 6650.0


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--leapfrog-per-launch", type=int, default=LEAPFROG_PER_LAUNCH)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def measured_hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:  # noqa: BLE001
        return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples SM clock / clock-event reasons during the timed region through NVML (no
    subprocesses: forking from the benchmark process would stall the launch thread)."""

    REASONS = {
        "hw_slowdown": 0x8,
        "sw_power_cap": 0x4,
        "sw_thermal_slowdown": 0x20,
        "hw_thermal_slowdown": 0x40,
    }

    def __init__(self, index):
        self.index = index
        self.rows = []
        self._stop = threading.Event()
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._nvml = None
        try:  # initialise NVML here, outside any timed region
            import pynvml

            pynvml.nvmlInit()
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self._max = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
            self._nvml = pynvml
        except Exception:  # noqa: BLE001
            pass

    def _run(self):
        pynvml, h, mx = self._nvml, getattr(self, "_h", None), getattr(self, "_max", None)
        if pynvml is None:
            return
        while not self._stop.is_set():
            try:
                self.rows.append((
                    pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), mx,
                    pynvml.nvmlDeviceGetCurrentClocksEventReasons(h),
                    pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0,
                ))
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(0.004)

    def __enter__(self):
        self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self._thread.join(timeout=10)

    def summary(self):
        sm = sorted(r[0] for r in self.rows)
        reasons = sorted(k for k, bit in self.REASONS.items() if any(r[2] & bit for r in self.rows))
        return {
            "sm_mhz": sm[len(sm) // 2] if sm else None,
            "sm_max_mhz": self.rows[0][1] if self.rows else None,
            "reasons": reasons,
            "samples": len(self.rows),
            "power_w_max": max((r[3] for r in self.rows), default=None),
        }


def cpu_baseline(kind_note=""):
    from oracle import cpu_baseline as cb

    cores = os.cpu_count() or 1
    res = cb.run("C1", {"n_chains": N_CHAINS, "dim": DIM}, chains_per_worker=8,
                 n_steps=LEAPFROG_PER_LAUNCH, reps=150, n_workers=cores)
    return {
        "value": res["value"],
        "unit": UNIT,
        "cores": res["cores"],
        "kind": "port",
        "sample": res["sample"] + kind_note,
        "seconds": res["seconds"],
    }


def run_reference(args, rank, world):
    """Reference arm: the reference's CPU algorithm (oracle port; the Python reference cannot
    travel to the GPU box) on all host cores, bounded sample per step."""
    if rank != 0:
        return
    from oracle import cpu_baseline as cb

    cores = os.cpu_count() or 1
    pool = cb.Pool(cores)
    vals = []
    for i in range(args.warmup + args.steps):
        if i == args.warmup:
            t0 = time.perf_counter()
        res = cb.run("C1", {"n_chains": N_CHAINS, "dim": DIM}, chains_per_worker=8,
                     n_steps=args.leapfrog_per_launch, reps=10, pool=pool)
        if i >= args.warmup:
            vals.append(res["value"])
    wall = time.perf_counter() - t0
    pool.close()
    value = sum(vals) / len(vals)
    line = {
        "impl": "reference",
        "metric": METRIC,
        "value": value,
        "unit": UNIT,
        "n_gpus": args.gpus,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": 1e3 * wall / args.steps,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "leapfrog_steps_per_launch": args.leapfrog_per_launch,
                   "integrator": "LeapfrogIntegrator", "metric": "dense"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": res["sample"]},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def run_cuda(args, rank, local_rank, world):
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline()  # before CUDA is initialised in this process

    import numpy as np
    import torch
    import torch.distributed as dist

    from mici_b200 import engine, parallel, problems

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    L = args.leapfrog_per_launch
    prob = problems.make_problem("C1", n_chains=N_CHAINS, dim=DIM,
                                 seed=problems.BASE_SEED + 1 + 1000 * rank)
    integ = engine.build_integrator(prob)
    state = engine.build_state(prob, dev)
    n, dim = state.pos.shape

    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)  # > 126 MB L2

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---------------- device-resident throughput (value) + per-launch kernel time
    for _ in range(args.warmup):
        integ.step_n(state, L)
    sync_all()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for _ in range(args.steps)]
    sampler = ClockSampler(local_rank)
    with sampler as clocks:
        for _ in range(2):  # sampler thread is running: absorb its start-up before timing
            integ.step_n(state, L)
        sync_all()
        out = state
        # Park the GPU while the host enqueues the whole timed sequence, so that the per-launch
        # event pairs contain device time only (no host launch gaps).  If the host was slower
        # than the parking time the measurement is repeated once with a longer park.
        park_ms = 3.0 * args.steps + 10.0
        for attempt in range(2):
            t_host = time.perf_counter()
            torch.cuda._sleep(int(park_ms * 1e-3 * 1.9e9))
            for i in range(args.steps):
                flush.fill_(float(i))  # evict q, p, M^-1 from L2 between timed launches
                ev[i][0].record()
                out = integ.step_n(state, L)
                ev[i][1].record()
            host_ms = (time.perf_counter() - t_host) * 1e3
            sync_all()
            if host_ms < 0.8 * park_ms:
                break
            park_ms = 2.0 * host_ms + 10.0
    times_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = torch.tensor([sum(times_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    assert int((out.status != 0).sum().item()) == 0 and bool(torch.isfinite(out.pos).all())

    # ---------------- end-to-end through the public API with HOST buffers
    pos_h = torch.as_tensor(prob.pos).pin_memory()
    mom_h = torch.as_tensor(prob.mom).pin_memory()
    pos_o = torch.empty_like(pos_h).pin_memory()
    mom_o = torch.empty_like(mom_h).pin_memory()
    st_o = torch.empty(n, dtype=torch.int32).pin_memory()

    def e2e_step():
        integ.step_n_host(pos_h, mom_h, L, out_pos=pos_o, out_mom=mom_o, out_status=st_o,
                          device=dev, n_chunks=8)

    for _ in range(args.warmup):
        e2e_step()
    sync_all()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        e2e_step()
    e1.record()
    sync_all()
    e2e_ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_ms = float(e2e_ms.item())
    _ = time.perf_counter() - t0

    # ---------------- write-out: the one collective (outside the timed step path)
    if world > 1:
        gathered = parallel.gather_state(out, n * world, dst=0)
        if rank == 0:
            assert gathered["pos"].shape[0] == n * world

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    chain_steps_per_launch = n * L
    value = world * chain_steps_per_launch * args.steps / (total_ms * 1e-3)
    e2e_value = world * chain_steps_per_launch * args.steps / (e2e_ms * 1e-3)
    b_alg = prob.algorithmic_bytes_per_chain_step  # 32 * D = 4096 B
    avg_launch_s = (sum(times_ms) / len(times_ms)) * 1e-3
    hbm_peak, peak_src = measured_hbm_peak()
    achieved_gbs = b_alg * chain_steps_per_launch / avg_launch_s / 1e9
    flops = (2.0 * dim * dim) * chain_steps_per_launch / avg_launch_s
    line = {
        "metric": METRIC,
        "value": value,
        "unit": UNIT,
        "n_gpus": world,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps,
        "ms_per_step_min_median_max": [min(times_ms), sorted(times_ms)[len(times_ms) // 2],
                                       max(times_ms)],
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f64",
        "data": "synthetic",
        "config": {
            "workload": WORKLOAD,
            "chains_per_gpu": n,
            "dim": dim,
            "leapfrog_steps_per_launch": L,
            "integrator": "LeapfrogIntegrator",
            "metric": "dense (explicit M^-1, shared)",
            "step_size": prob.step_size,
            "l2": "flushed between timed launches (256 MB fill)",
            "parallelism": f"chains sharded over {world} GPU(s), no collective on the step path",
        },
        "roofline": {
            "bound": "hbm",
            "achieved": achieved_gbs,
            "peak": hbm_peak,
            "unit": "GB/s",
            "frac": achieved_gbs / hbm_peak,
            "traffic": NCU_DRAM_TRAFFIC_BYTES_PER_LAUNCH if (n, dim) == (8192, 128) else None,
            "peak_source": peak_src,
            "kernel": "leapfrog_dmma_kernel<NealFunnelTarget,128>",
            "algorithmic_bytes_per_chain_step": b_alg,
            "fp64_tflops": flops / 1e12,
            "fp64_peak_tflops": FP64_PEAK_TFLOPS,
            "fp64_frac": flops / 1e12 / FP64_PEAK_TFLOPS,
        },
        "e2e": {
            "value": e2e_value,
            "unit": UNIT,
            "h2d_bytes_per_step": int(2 * n * dim * 8),
            "d2h_bytes_per_step": int(2 * n * dim * 8 + n * 4),
        },
        "gpu_launches": args.steps,
        "clocks": clocks.summary(),
    }
    if cpu is not None:
        line["cpu_baseline"] = cpu
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_cuda(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
