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


def ncu_capture_for(kernel: str, cells: int):
    """profiles/kernel_traffic.json: per kernel, `ncu --set full` captures of THIS build (commit recorded in the file)
    at the bench sizes - DRAM bytes per launch, issue-slot utilisation, SM active / elapsed.  The capture whose cell
    count is nearest is scaled per cell."""
    tfile = ROOT / "profiles" / "kernel_traffic.json"
    try:
        caps = [c for c in json.loads(tfile.read_text())["captures"] if c["kernel"].startswith(kernel)]
        return min(caps, key=lambda c: abs(np.log(c["cells"] / cells))) if caps else None
    except Exception:
        return None


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


# --------------------------------------------------------------------------- CPU (reference / oracle) arm
def host_threads() -> int:
    """CPUs this process may run on.  torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arm sets its
    thread count explicitly from this instead."""
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_sample_of(w: W.Workload, max_cells: int) -> W.Workload:
    """A bounded sample of workload `w` for the CPU arm: the first `rows` rows of the grid at full width, same
    parameters and iteration count, with the obstacles / impulses of `w` that fall into the band (positions are
    normalised, so y is rescaled to the band).  The CPU rate in cell-steps per second does not depend on the number
    of rows (every stage is a row-streaming stencil), so the sample's rate is reported as the workload's."""
    if w.cells <= max_cells:
        return w
    rows = max(64, min(w.height, (max_cells // w.width) // 16 * 16))
    scale = w.height / rows
    circles = [(px, py * scale, r) for (px, py, r) in w.circles if (py * w.height - r) < rows]
    dye = None
    if w.dye_size:
        dye = (w.dye_size[0], max(16, int(round(w.dye_size[1] / scale))))
    import dataclasses
    return dataclasses.replace(w, name=f"{w.name}[rows 0..{rows - 1}]", height=rows, circles=circles, dye_size=dye)


def time_cpu_arm(w: W.Workload, steps: int, warmup: int, budget_s: float, max_cells: int, prefer: str = "reference"):
    """Times the reference's CPU implementation of the step on all host threads, on a bounded sample of `w`.

    kind "reference": oracle/_ref, the reference's own shader text compiled as C++ and dispatched in the reference's
    order (oracle/natrix_ref.py), when its prebuilt library is present; kind "port": the C/OpenMP restatement
    (oracle/natrix_oracle.c).  Returns (Mcell-steps/s, info)."""
    threads = host_threads()
    sw = cpu_sample_of(w, max_cells)
    kind = "port"
    if prefer == "reference":
        from oracle import natrix_ref as R
        if R.available():
            kind = "reference"
    if kind == "reference":
        variant = "literal" if sw.cells <= 2 ** 24 else "exact"      # float32 linear indices are exact up to 2^24 cells
        R.lib(variant).nref_set_threads(threads)
        sim_cls = lambda wd, ht, layout=None: R.RefFluidSimulator(wd, ht, layout, variant=variant)   # noqa: E731
        dye_cls = R.RefSmoothParticlesArea
        impl = f"oracle/_ref/libnatrix_ref{'_exact' if variant == 'exact' else ''}.so (reference shaders compiled as C++, OpenMP over rows)"
    else:
        from oracle import c_oracle
        c_oracle.lib().nox_set_threads(threads)
        sim_cls, dye_cls = c_oracle.COracleFluidSimulator, c_oracle.COracleSmoothParticlesArea
        impl = "oracle/natrix_oracle.c (C/OpenMP restatement of the reference shaders)"
    sim, dye = W.build(sw, sim_cls, dye_cls if sw.dye_size else None)
    t0 = time.perf_counter()
    W.run_step(sw, sim, dye, 0)
    first = time.perf_counter() - t0
    done_warm = 1
    while done_warm < warmup and first * (done_warm + 2) < budget_s / 2:
        W.run_step(sw, sim, dye, done_warm)
        done_warm += 1
    k = max(1, min(steps, int((budget_s - first * done_warm) / max(first, 1e-9))))
    t0 = time.perf_counter()
    for i in range(k):
        W.run_step(sw, sim, dye, done_warm + i)
    dt = (time.perf_counter() - t0) / k
    value = sw.cells / dt / 1e6
    sample = (f"{k} full steps of {sw.name} ({sw.width}x{sw.height} = {sw.cells / 1e6:.1f} Mcells"
              f"{'' if sw is w else f', a band of the {w.width}x{w.height} workload'}, {sw.iterations} Jacobi iterations, "
              f"{len(sw.circles)} obstacles{', dye ' + 'x'.join(map(str, sw.dye_size)) if sw.dye_size else ''}) after "
              f"{done_warm} warm-up, {dt:.3f} s/step, {threads} threads; {impl}")
    return value, {"kind": kind, "cores": threads, "host_cpus": os.cpu_count(), "sample": sample, "unit": METRIC,
                   "value": value, "seconds_per_step": dt, "steps": k, "sample_grid": [sw.width, sw.height]}


def run_reference_arm(args, w: W.Workload):
    """--impl reference: the reference's own CPU implementation of the step (oracle/_ref, else the C port) on all the
    host threads this process may use, on a bounded sample of the SAME workload the GPU arm runs at this N.  Under
    torchrun rank 0 alone runs it; the other ranks exit 0."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    value, info = time_cpu_arm(w, args.steps, max(args.warmup, 1), budget_s=120.0, max_cells=args.cpu_cells)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": METRIC, "n_gpus": args.gpus,
        "steps": info["steps"], "warmup": args.warmup, "ms_per_step": info["seconds_per_step"] * 1e3 * w.cells / (info["sample_grid"][0] * info["sample_grid"][1]),
        "higher_is_better": True, "scaling": "weak" if "weak" in w.name else "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": w.name, "grid": [w.width, w.height], "jacobi_iterations": w.iterations,
                   "dye": list(w.dye_size) if w.dye_size else None,
                   "note": "the reference's CPU path on the host cores (its bgfx engine cannot run headless here): value is "
                           "the cell-step rate measured on the bounded sample named in cpu_baseline.sample, ms_per_step is "
                           "that rate applied to the full grid"},
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
    if args.jacobi_kernel is not None:
        sim.set_option(L.OPT_JACOBI_KERNEL, args.jacobi_kernel)
    if args.smem_depth is not None:
        sim.set_option(L.OPT_SMEM_DEPTH, args.smem_depth)
    sim.set_option(L.OPT_TIMING, 1)
    stream = torch.cuda.ExternalStream(sim.cuda_stream, device=dev)

    step = 0
    for _ in range(max(args.warmup, 3)):
        W.run_step(w, sim, dye, step); step += 1
    sim.stats("velocity")        # warm-up covers the e2e leg's readback kernel too (CUDA loads a kernel on its first launch)
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
    e2e_median_ms, e2e_max_ms = 1e3 * statistics.median(t_e2e), 1e3 * max(t_e2e)
    e2e_value = w.cells / (e2e_ms * 1e-3) / 1e6

    # ---- roofline of the dominant kernel (the Jacobi sweeps), measured live
    pipeline = sim.get_option(L.OPT_PIPELINE)
    jk = sim.get_option(L.OPT_JACOBI_KERNEL)            # the kernel in use: 0 (pipeline 0), 1 TMA streaming, 2 shared memory
    depth = sim.get_option(L.OPT_SMEM_DEPTH if jk == 2 else L.OPT_JACOBI_DEPTH)
    jl = w.iterations if pipeline == 0 else -(-w.iterations // depth)
    kname = {0: "k_poisson_ref", 1: "k_jacobi_tb", 2: "k_jacobi_smem"}[jk]
    peak, peak_src = measured_peak_gbs()
    jac_ms = statistics.mean(jacobi_ms)
    algo_bytes_step = JACOBI_BYTES_PER_CELL_SWEEP * w.cells * w.iterations
    algo_achieved = algo_bytes_step / (jac_ms * 1e-3) / 1e9
    # bytes one launch MUST move: p 4 + div 4 + mask 1 read, p 4 written, per cell
    min_bytes = 13 * w.cells
    cap = ncu_capture_for(kname, w.cells)
    traffic = None if cap is None else cap["dram_bytes_per_cell_per_launch"] * w.cells
    phys = traffic if traffic is not None else min_bytes
    achieved = phys / (jac_ms / jl * 1e-3) / 1e9
    roofline = {"kernel": kname, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": None if cap is None else cap.get("source"),
                "min_bytes_per_launch": min_bytes, "min_bytes_frac": min_bytes / (jac_ms / jl * 1e-3) / 1e9 / peak,
                "algorithmic_achieved": algo_achieved, "algorithmic_frac": algo_achieved / peak,
                "peak_source": peak_src, "launches_per_step": jl, "avg_launch_ms": jac_ms / jl,
                "algorithmic_bytes_per_launch": algo_bytes_step / jl,
                "issue_active_pct": None if cap is None else cap.get("issue_active_pct"),
                "sm_active_over_elapsed": None if cap is None else cap.get("sm_active_over_elapsed"),
                "note": "achieved/frac are PHYSICAL: DRAM bytes of one launch over the launch duration measured live here, against "
                        "the measured HBM peak.  The bytes are ncu's dram__bytes_read+write of this build at this size "
                        "(profiles/kernel_traffic.json, `traffic`) or, when no capture of this size exists (`traffic` null), "
                        "the 13 B per cell the launch must move (min_bytes_*).  algorithmic_* is SURVEY 8(d)'s figure, "
                        "20 B x cells x sweeps per launch of `depth` sweeps over the same duration: it exceeds the peak by "
                        "design, because temporal blocking reads p, div and the mask once per `depth` sweeps"}

    # ---- the other stages of the step: physical bytes (what the kernel must move) next to the algorithmic ones
    stage_algo = {"advect": 24 + 12 + 20 + (16 if w.viscosity > 0 else 0) + 20,   # the fused pre-projection kernel
                  "gradient": 28}
    stage_phys = {"advect": 26, "gradient": 21}      # vel 8 + obs 1 -> vel 8, vort 4, div 4, mask 1; p 4 + mask 1 + vel 8 -> vel 8
    stage_kernel = {"advect": "k_preproject", "gradient": "k_gradient_mask4"}
    stages_roofline = {}
    if pipeline == 1:
        for name, bpc in stage_algo.items():
            ms = stage[name]
            if ms > 0:
                gbs = stage_phys[name] * w.cells / (ms * 1e-3) / 1e9
                c2 = ncu_capture_for(stage_kernel[name], w.cells)
                stages_roofline[name] = {"kernel": stage_kernel[name], "ms": ms, "bytes_per_cell": stage_phys[name],
                                         "achieved": gbs, "frac": gbs / peak,
                                         "algorithmic_bytes_per_cell": bpc,
                                         "algorithmic_frac": bpc * w.cells / (ms * 1e-3) / 1e9 / peak,
                                         "ncu_dram_bytes_per_cell": None if c2 is None else c2["dram_bytes_per_cell_per_launch"]}

    # ---- CPU baseline beside it (bounded sample, all host threads)
    cpu = None
    if with_cpu and not args.no_cpu:
        _, info = time_cpu_arm(w, steps=3, warmup=1, budget_s=25.0, max_cells=args.cpu_cells)
        cpu = {k: info[k] for k in ("value", "unit", "cores", "kind", "sample")}
        _, port = time_cpu_arm(w, steps=2, warmup=1, budget_s=10.0, max_cells=args.cpu_cells, prefer="port")
        cpu["c_port"] = {k: port[k] for k in ("value", "cores", "sample")}     # the hand-written C/OpenMP restatement, for context

    step_bytes = w.algorithmic_bytes_per_cell_step() * w.cells
    line = {
        "metric": METRIC, "value": value, "unit": METRIC, "n_gpus": 1, "steps": steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": w.name, "grid": [w.width, w.height], "jacobi_iterations": w.iterations,
                   "dye": list(w.dye_size) if w.dye_size else None, "splats_per_step": w.splats_per_step,
                   "obstacles_per_step": len(w.circles), "pipeline": pipeline, "jacobi_kernel": kname, "jacobi_depth": depth,
                   "l2": "state (>= 560 MB) exceeds the 126 MB L2; no flush needed",
                   "algorithmic_GBps_full_step": step_bytes / (ms_per_step * 1e-3) / 1e9},
        "stage_ms": {k: round(v, 4) for k, v in stage.items()},
        "roofline": roofline, "stages_roofline": stages_roofline, "cpu_baseline": cpu,
        "e2e": {"value": e2e_value, "unit": METRIC, "ms_per_step": e2e_ms, "ms_per_step_median": e2e_median_ms,
                "ms_per_step_max": e2e_max_ms, "h2d_bytes_per_step": impulse_bytes_per_step(w),
                "d2h_bytes_per_step": 32,
                "note": "public Python API per step; inputs are the host-side impulse / obstacle parameters, the "
                        "result read back each step is the (sum, sumsq, min, max) of the velocity field"},
        "gpu_launches": int(launches), "clocks": clocks,
        "plan_cache": dict(zip(("hits", "misses"), sim.plan_cache_stats())),
        "so": L.loaded_library_path(),
    }
    sim.destroy()
    return line


SUB_KEYS = ("value", "unit", "ms_per_step", "config", "stage_ms", "roofline", "stages_roofline", "e2e", "gpu_launches", "plan_cache")


def run_single_gpu(args, w: W.Workload, secondary=None):
    line = measure_single_gpu(args, w, with_cpu=True, steps=args.steps)
    if secondary is not None:
        # the 4096^2 configuration the metric also quotes, measured in the same run
        sub = measure_single_gpu(args, secondary, with_cpu=False, steps=max(args.steps, 20))
        line["config3_4096"] = {k: sub[k] for k in SUB_KEYS}
        # config 5 with the 64 circles drifting 3 cells per step: a new obstacle set - and a new tile plan for the
        # Jacobi kernel - every step (the static workload re-adds identical circles and always hits the plan cache)
        import dataclasses
        moving = dataclasses.replace(w, name=w.name + "-moving", drift=3.0)
        sub = measure_single_gpu(args, moving, with_cpu=False, steps=min(args.steps, 10))
        line["config5_moving"] = {k: sub[k] for k in ("value", "unit", "ms_per_step", "stage_ms", "e2e", "gpu_launches", "plan_cache")}
        # the 1-GPU point of config 4 (16384^2, strong scaling; the N > 1 lines carry `config4_strong`)
        sub = measure_single_gpu(args, W.cfg4_workload(16384), with_cpu=False, steps=min(args.steps, 10))
        line["config4_16384"] = {k: sub[k] for k in ("value", "unit", "ms_per_step", "config", "stage_ms", "e2e", "gpu_launches")}
    print(json.dumps(line), flush=True)
    return 0


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="natrix_b200", choices=["natrix_b200", "reference"])
    ap.add_argument("--workload", default="auto", choices=["auto", "demo", "cfg2", "cfg3", "cfg4", "cfg5", "cfg5m"])
    ap.add_argument("--size", type=int, default=None, help="override the grid size of cfg3/cfg4")
    ap.add_argument("--pipeline", type=int, default=None)
    ap.add_argument("--depth", type=int, default=None)
    ap.add_argument("--packed", type=int, default=None, help="0/1: f32x2 arithmetic in the Jacobi kernel")
    ap.add_argument("--jacobi-kernel", type=int, default=None, help="0 auto, 1 TMA register-streaming kernel, 2 shared-memory kernel")
    ap.add_argument("--smem-depth", type=int, default=None, help="sweeps per launch of the shared-memory Jacobi kernel")
    ap.add_argument("--no-obstacles", action="store_true", help="diagnostic: drop the per-step obstacles")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--cpu-cells", type=int, default=8 * 1024 * 1024,
                    help="CPU arm: largest sample (cells) of the workload that is actually stepped on the host")
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
        if name == "cfg5m":
            w.name, w.drift = w.name + "-moving", 3.0

    if args.no_obstacles:
        w.circles = []
        w.name += "-noobst"
    if args.impl == "reference":
        return run_reference_arm(args, w)         # the SAME workload as the GPU arm at this N, on a bounded sample
    if n == 1 and world == 1:
        return run_single_gpu(args, w, secondary)
    from natrix_b200 import slabs

    return slabs.run_bench(args, w, METRIC, JACOBI_BYTES_PER_CELL_SWEEP, measured_peak_gbs, ClockSampler)


if __name__ == "__main__":
    sys.exit(main())
