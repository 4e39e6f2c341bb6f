"""Regenerates the golden fixtures in this directory FROM THE NUMPY ORACLE.

    python tests/golden/make_golden.py

The reference itself cannot produce vectors (its bgfx engine does not run here and it ships
no fixtures: "parity unpinned", see oracle/natrix_oracle.py), so these files pin the oracle
against accidental change, and give the GPU tests a fixed answer that does not depend on the
oracle code being importable.

* small_case.npz     - every field after each of 4 steps of a 96x64 grid (dye 192x128) with a
                       circle, a static triangle, splats, vorticity and viscosity;
* digests.json       - SHA-256 of the raw float32 bytes of every field for the BASELINE.json
                       configurations that fit a CPU test (config 1 after 1 and 3 steps, a
                       256x256 cut of config 2 after 2 steps).
"""
from __future__ import annotations

import hashlib
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from natrix_b200 import workloads as W                                     # noqa: E402
from oracle.natrix_oracle import OracleFluidSimulator, OracleSmoothParticlesArea  # noqa: E402

HERE = Path(__file__).resolve().parent


def small_case(sim_cls, dye_cls, steps=4):
    """The scripted small case; returns {f"{field}_{step}": array}."""
    w, h = 96, 64
    sim = sim_cls(w, h, None)
    sim.vorticity, sim.viscosity, sim.iterations = 1.0, 0.5, 20
    W.set_velocity(sim, W.random_velocity(w, h, seed=0))
    dye = dye_cls(2 * w, 2 * h, sim, None)
    dye.dissipation = 0.98
    out = {}
    for k in range(steps):
        sim.add_circle_obstacle((0.5, 0.5), 9.0)
        sim.add_triangle_obstacle((0.1, 0.1), (0.3, 0.15), (0.2, 0.4), static=(k % 2 == 0))
        sim.update(W.DT)
        dye.update(W.DT)
        pos = (0.5 + 0.3 * np.cos(0.1 * k), 0.5 + 0.3 * np.sin(0.1 * k))
        sim.add_velocity(pos, (0.3, -0.2), 8.0)
        dye.add_particles(pos, 25.0, 0.04)
        for name, arr in W.fields_of(sim, dye).items():
            out[f"{name}_{k}"] = arr.copy()
    return out


def digest(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a, dtype=np.float32).tobytes()).hexdigest()


def run_digests(sim_cls, dye_cls):
    res = {}
    w1 = W.demo_workload()
    sim, dye = W.build(w1, sim_cls, dye_cls)
    for k in range(3):
        W.run_step(w1, sim, dye, k)
        if k in (0, 2):
            res[f"config1_step{k + 1}"] = {n: digest(a) for n, a in W.fields_of(sim, dye).items()}
    w2 = W.cfg2_workload()
    w2.width = w2.height = 256
    w2.circles = [(px, py, r / 4.0) for (px, py, r) in w2.circles]
    sim, dye = W.build(w2, sim_cls, None)
    for k in range(2):
        W.run_step(w2, sim, None, k)
    res["config2_256_step2"] = {n: digest(a) for n, a in W.fields_of(sim).items()}
    return res


if __name__ == "__main__":
    np.savez_compressed(HERE / "small_case.npz", **small_case(OracleFluidSimulator, OracleSmoothParticlesArea))
    (HERE / "digests.json").write_text(json.dumps(run_digests(OracleFluidSimulator, OracleSmoothParticlesArea),
                                                  indent=1, sort_keys=True) + "\n")
    print("wrote", HERE / "small_case.npz", (HERE / "small_case.npz").stat().st_size, "bytes")
