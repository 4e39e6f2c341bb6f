#!/usr/bin/env python
"""bench.py - Mcell-steps/s of the Natrix stable-fluids step on B200 (BASELINE.json's metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload NAME]

A "step" is one frame of the reference's demo loop over one synthetic workload (obstacles ->
FluidSimulator.update -> dye update -> velocity/dye impulses):

* every N: config 5 of BASELINE.json - weak scaling, one 32768 x 4096 row slab per GPU (global grid
  32768 x 4096 N), 200 Jacobi iterations per step, 64 circular obstacles per step AND per GPU (the
  1-GPU domain tiled N times along y, so cells and obstacle load per GPU are fixed); N > 1 runs under
  torchrun, one rank per GPU, halo exchange over NCCL overlapped with the interior Jacobi launches.
  The per-N values therefore form ONE series.
* the N = 1 line also carries `config3_4096`: config 3 - 4096^2 velocity + 4096^2 dye, 100 iterations,
  8 velocity + 8 dye splats and one circular obstacle per step (`--workload cfg3` makes it the primary).

One JSON line is printed by rank 0.  `value` is device-timed with inputs resident in HBM
(CUDA events on the simulator's stream, max over ranks); `e2e` is the same metric through the
Python API with a per-step host readback of a field statistic; `roofline` is the Jacobi kernel
(algorithmic 20 B per cell-sweep over its measured duration); `cpu_baseline` is the C/OpenMP
oracle port timed on this host's cores on a bounded sample of the same workload.
`--impl reference` times that CPU port alone (the reference's bgfx engine cannot run headless
here - DESIGN.md) on all host threads.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402

from natrix_b200 import workloads as W  # noqa: E402

METRIC = "Mcell-steps/s"
JACOBI_BYTES_PER_CELL_SWEEP = 20        # p 4 + div 4 + obstacles 8 -> p 4


def measured_peak_gbs():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks and throttle reasons sampled DURING the timed region."""

    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device = device
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.device}", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                 "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1])); smax.append(float(parts[2])); power.append(float(parts[3]))
            except ValueError:
                continue
            for name, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(smax), "power_w_max": max(power),
                "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------- CPU (oracle) arm
def time_cpu_port(w: W.Workload, steps: int, warmup: int, budget_s: float):
    """Times the C/OpenMP oracle port on all host threads.  Returns (Mcell-steps/s, info)."""
    from oracle import c_oracle

    threads = c_oracle.max_threads()
    sim, dye = W.build(w, c_oracle.COracleFluidSimulator, c_oracle.COracleSmoothParticlesArea if w.dye_size else None)
    t0 = time.perf_counter()
    W.run_step(w, sim, dye, 0)
    first = time.perf_counter() - t0
    done_warm = 1
    while done_warm < warmup and first * (done_warm + 2) < budget_s / 2:
        W.run_step(w, sim, dye, done_warm)
        done_warm += 1
    k = max(1, min(steps, int((budget_s - first * done_warm) / max(first, 1e-9))))
    t0 = time.perf_counter()
    for i in range(k):
        W.run_step(w, sim, dye, done_warm + i)
    dt = (time.perf_counter() - t0) / k
    value = w.cells / dt / 1e6
    sample = (f"{k} full steps of {w.name} ({w.width}x{w.height}, {w.iterations} Jacobi iterations"
              f"{', dye ' + 'x'.join(map(str, w.dye_size)) if w.dye_size else ''}) after {done_warm} warm-up, "
              f"{dt:.3f} s/step")
    return value, {"kind": "port", "cores": threads, "host_cpus": os.cpu_count(), "sample": sample,
                   "unit": METRIC, "value": value, "seconds_per_step": dt, "steps": k}


def run_reference_arm(args, w: W.Workload):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    value, info = time_cpu_port(w, args.steps, args.warmup, budget_s=150.0)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": METRIC, "n_gpus": args.gpus,
        "steps": info["steps"], "warmup": args.warmup, "ms_per_step": info["seconds_per_step"] * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": w.name, "grid": [w.width, w.height], "jacobi_iterations": w.iterations,
                   "dye": list(w.dye_size) if w.dye_size else None,
                   "note": "CPU port of the reference shaders (oracle/natrix_oracle.c, OpenMP); the reference's "
                           "bgfx engine cannot run headless here"},
        "cpu_baseline": {k: info[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": value, "unit": METRIC, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# --------------------------------------------------------------------------- GPU arm, N = 1
def impulse_bytes_per_step(w: W.Workload) -> int:
    """Bytes of host parameters that cross the C ABI per step (floats of obstacle / splat calls)."""
    b = 4 * 4 * len(w.circles) + 4 * 5 * w.splats_per_step + 4 * 8          # circles, add_velocity, step params
    if w.dye_size:
        b += 4 * 4 * w.splats_per_step + 4 * 3
    return b


def measure_single_gpu(args, w: W.Workload, with_cpu: bool, steps: int):
    """All N = 1 measurements of one workload; returns the JSON-line dict."""
    import torch

    from natrix_b200 import _lib as L
    from natrix_b200.core.fluid_simulator import FluidSimulator
    from natrix_b200.smooth_particles_area import SmoothParticlesArea

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the product path")
    dev = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(dev)
    sim, dye = W.build(w, FluidSimulator, SmoothParticlesArea if w.dye_size else None, device=dev)
    if args.pipeline is not None:
        sim.set_option(L.OPT_PIPELINE, args.pipeline)
    if args.depth is not None:
        sim.set_option(L.OPT_JACOBI_DEPTH, args.depth)
    if args.packed is not None:
        sim.set_option(L.OPT_PACKED, args.packed)
    sim.set_option(L.OPT_TIMING, 1)
    stream = torch.cuda.ExternalStream(sim.cuda_stream, device=dev)

    step = 0
    for _ in range(max(args.warmup, 3)):
        W.run_step(w, sim, dye, step); step += 1
    sim.synchronize()

    # ---- value: device-timed, inputs resident in HBM, no host readback inside the region
    sampler = ClockSampler(dev)
    sampler.start()
    launches0 = sim.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    jacobi_ms = []
    torch.cuda.synchronize()
    ev0.record(stream)
    for _ in range(steps):
        W.run_step(w, sim, dye, step); step += 1
    sim.synchronize()            # flushes queued impulses so they are inside the timed region
    ev1.record(stream)
    torch.cuda.synchronize()
    total_ms = ev0.elapsed_time(ev1)
    launches = sim.launch_count - launches0
    stage = sim.timings()        # per-stage CUDA events of the LAST timed step
    jacobi_ms.append(stage["jacobi"])
    ms_per_step = total_ms / steps
    value = w.cells / (ms_per_step * 1e-3) / 1e6

    # ---- e2e: same steps through the public API, per-step host sync + D2H of a field statistic
    t_e2e = []
    for _ in range(steps):
        t0 = time.perf_counter()
        W.run_step(w, sim, dye, step); step += 1
        ke = sim.stats("velocity")          # kinetic-energy style metric: 32 B device -> host, synchronises
        t_e2e.append(time.perf_counter() - t0)
        jacobi_ms.append(sim.timings()["jacobi"])
    clocks = sampler.stop()
    e2e_ms = 1e3 * sum(t_e2e) / len(t_e2e)
    e2e_value = w.cells / (e2e_ms * 1e-3) / 1e6

    # ---- roofline of the dominant kernel (temporally blocked Jacobi), measured live
    depth = sim.get_option(L.OPT_JACOBI_DEPTH)
    pipeline = sim.get_option(L.OPT_PIPELINE)
    jl = w.iterations if pipeline == 0 else -(-w.iterations // depth)
    peak, peak_src = measured_peak_gbs()
    jac_ms = statistics.mean(jacobi_ms)
    algo_bytes_step = JACOBI_BYTES_PER_CELL_SWEEP * w.cells * w.iterations
    achieved = algo_bytes_step / (jac_ms * 1e-3) / 1e9
    traffic = None
    tfile = ROOT / "profiles" / "jacobi_traffic.json"
    if tfile.exists():
        try:        # ncu dram__bytes_read.sum + dram__bytes_write.sum of one depth-8 launch of the nearest captured size
            caps = json.loads(tfile.read_text())["captures"]
            cap = min(caps, key=lambda c: abs(np.log(c["cells"] / w.cells)))
            traffic = cap["dram_bytes_per_cell_per_launch"] * w.cells
        except Exception:
            traffic = None
    dram_achieved = None if traffic is None or pipeline == 0 else traffic / (jac_ms / jl * 1e-3) / 1e9
    roofline = {"kernel": "k_jacobi_tb" if pipeline else "k_poisson_ref", "bound": "hbm", "achieved": achieved,
                "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic if pipeline else None,
                "dram_achieved": dram_achieved, "dram_frac": None if dram_achieved is None else dram_achieved / peak,
                "peak_source": peak_src, "launches_per_step": jl, "avg_launch_ms": jac_ms / jl,
                "algorithmic_bytes_per_launch": algo_bytes_step / jl,
                "note": "achieved/frac use ALGORITHMIC bytes = 20 B x cells x sweeps (reference field widths); one launch of "
                        "`depth` sweeps really moves ~13 B per cell (traffic, from ncu), so frac exceeds 1 by design and "
                        "dram_frac = traffic / launch time / peak is the physical HBM utilisation: the kernel is "
                        "instruction-issue bound, not HBM bound"}

    # ---- the other stages of the step against the same roofline (algorithmic B/cell: SURVEY 8(d))
    stage_algo = {"advect": 24 + 12 + 20 + (16 if w.viscosity > 0 else 0) + 20,   # the fused pre-projection kernel
                  "gradient": 28}
    stages_roofline = {}
    if pipeline == 1:
        for name, bpc in stage_algo.items():
            ms = stage[name]
            if ms > 0:
                gbs = bpc * w.cells / (ms * 1e-3) / 1e9
                stages_roofline[name] = {"algorithmic_bytes_per_cell": bpc, "ms": ms, "achieved": gbs, "frac": gbs / peak}

    # ---- CPU baseline beside it (bounded sample, all host threads)
    cpu = None
    if with_cpu and not args.no_cpu:
        _, info = time_cpu_port(w, steps=3, warmup=1, budget_s=30.0)
        cpu = {k: info[k] for k in ("value", "unit", "cores", "kind", "sample")}

    step_bytes = w.algorithmic_bytes_per_cell_step() * w.cells
    line = {
        "metric": METRIC, "value": value, "unit": METRIC, "n_gpus": 1, "steps": steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": w.name, "grid": [w.width, w.height], "jacobi_iterations": w.iterations,
                   "dye": list(w.dye_size) if w.dye_size else None, "splats_per_step": w.splats_per_step,
                   "obstacles_per_step": len(w.circles), "pipeline": pipeline, "jacobi_depth": depth,
                   "l2": "state (>= 560 MB) exceeds the 126 MB L2; no flush needed",
                   "algorithmic_GBps_full_step": step_bytes / (ms_per_step * 1e-3) / 1e9},
        "stage_ms": {k: round(v, 4) for k, v in stage.items()},
        "roofline": roofline, "stages_roofline": stages_roofline, "cpu_baseline": cpu,
        "e2e": {"value": e2e_value, "unit": METRIC, "ms_per_step": e2e_ms, "h2d_bytes_per_step": impulse_bytes_per_step(w),
                "d2h_bytes_per_step": 32,
                "note": "public Python API per step; inputs are the host-side impulse / obstacle parameters, the "
                        "result read back each step is the (sum, sumsq, min, max) of the velocity field"},
        "gpu_launches": int(launches), "clocks": clocks,
        "so": L.loaded_library_path(),
    }
    sim.destroy()
    return line


def run_single_gpu(args, w: W.Workload, secondary=None):
    line = measure_single_gpu(args, w, with_cpu=True, steps=args.steps)
    if secondary is not None:
        # the 4096^2 configuration the metric also quotes, measured in the same run
        sub = measure_single_gpu(args, secondary, with_cpu=False, steps=max(args.steps, 20))
        line["config3_4096"] = {k: sub[k] for k in ("value", "unit", "ms_per_step", "config", "stage_ms", "roofline", "stages_roofline", "e2e",
                                                     "gpu_launches")}
    print(json.dumps(line), flush=True)
    return 0


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="natrix_b200", choices=["natrix_b200", "reference"])
    ap.add_argument("--workload", default="auto", choices=["auto", "demo", "cfg2", "cfg3", "cfg4", "cfg5"])
    ap.add_argument("--size", type=int, default=None, help="override the grid size of cfg3/cfg4")
    ap.add_argument("--pipeline", type=int, default=None)
    ap.add_argument("--depth", type=int, default=None)
    ap.add_argument("--packed", type=int, default=None, help="0/1: f32x2 arithmetic in the Jacobi kernel")
    ap.add_argument("--no-obstacles", action="store_true", help="diagnostic: drop the per-step obstacles")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args(argv)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    n = max(args.gpus, 1)
    name = args.workload
    secondary = None
    if name == "auto":
        # one workload family for every N so that the per-N values form a weak-scaling series:
        # config 5, 32768 x 4096 cells per GPU.  The N = 1 line also carries config 3 (4096^2).
        name = "cfg5"
        if n == 1:
            secondary = W.cfg3_workload(4096)
    if name == "demo":
        w = W.demo_workload()
    elif name == "cfg2":
        w = W.cfg2_workload()
    elif name == "cfg3":
        w = W.cfg3_workload(args.size or 4096)
    elif name == "cfg4":
        w = W.cfg4_workload(args.size or 16384)
    else:
        w = W.cfg5_workload(n)

    if args.no_obstacles:
        w.circles = []
        w.name += "-noobst"
    if args.impl == "reference":
        if name == "cfg5":          # keep the CPU arm bounded: the 1-GPU slab of the weak-scaling grid
            w = W.cfg5_workload(1)
        return run_reference_arm(args, w)
    if n == 1 and world == 1:
        return run_single_gpu(args, w, secondary)
    from natrix_b200 import slabs

    return slabs.run_bench(args, w, METRIC, JACOBI_BYTES_PER_CELL_SWEEP, measured_peak_gbs, ClockSampler)


if __name__ == "__main__":
    sys.exit(main())
