"""Deterministic synthetic workloads for the configurations BASELINE.json names
(SURVEY.md section 8(d)).  Engine-agnostic host logic: every driver only uses the reference's
public API (``add_circle_obstacle``, ``update``, ``add_velocity``, ``add_particles`` ...), so the
same function drives the CUDA product, the NumPy oracle and the C oracle.

Initial velocity is the one thing the reference API cannot express (it has no upload path,
SURVEY Q16); ``set_velocity`` uses whichever extension the engine offers.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, List, Optional, Tuple

import numpy as np

DT = float(np.float32(1.0 / 60.0))      # dt = float32(1/60)


@dataclass
class Workload:
    name: str
    width: int
    height: int
    iterations: int
    vorticity: float
    viscosity: float
    dye_size: Optional[Tuple[int, int]] = None
    dye_dissipation: float = 0.98
    circles: List[Tuple[float, float, float]] = field(default_factory=list)   # (px, py, radius)
    splats_per_step: int = 0
    splat_radius: float = 32.0
    dye_radius: float = 250.0
    dye_strength: float = 0.04
    orbit_seed: Optional[int] = None
    init: str = "zero"                   # "zero" | "smooth" | "random"
    drift: float = 0.0                   # cells per step the obstacles move along +x (wrapping): a new obstacle set every step

    @property
    def cells(self) -> int:
        return self.width * self.height

    def algorithmic_bytes_per_cell_step(self) -> int:
        """B_step(N) = 132 + 20 N (116 + 20 N when viscosity == 0), plus the
        impulses and dye passes this workload issues per step."""
        b = (132 if self.viscosity > 0 else 116) + 20 * self.iterations
        b += 16 * self.splats_per_step
        if self.dye_size:
            dye_ratio = (self.dye_size[0] * self.dye_size[1]) / self.cells
            b += int(round((24 + 8 * self.splats_per_step) * dye_ratio))
        return b


def demo_workload() -> Workload:
    """config 1 - demo/simulation_demo.py defaults (:100-110, :220-237, :73-74, :367)."""
    return Workload("demo-640x360", 640, 360, 50, 1.0, 0.5, dye_size=(1280, 720), circles=[(0.5, 0.5, 40.0)],
                    splats_per_step=1, splat_radius=32.0, init="zero")


def cfg2_workload() -> Workload:
    """config 2 - 1024^2, 50 iterations, vorticity confinement and circular obstacles."""
    return Workload("1024sq-N50", 1024, 1024, 50, 1.0, 0.1,
                    circles=[(0.25, 0.25, 40.0), (0.75, 0.25, 60.0), (0.5, 0.6, 80.0), (0.3, 0.8, 25.0)],
                    init="smooth")


def cfg3_workload(size: int = 4096) -> Workload:
    """config 3 - 4096^2 velocity + dye, 100 iterations, continuous splat impulses."""
    return Workload(f"{size}sq-N100-splats-dye", size, size, 100, 1.0, 0.0, dye_size=(size, size),
                    circles=[(0.5, 0.5, 256.0 * size / 4096.0)], splats_per_step=8,
                    splat_radius=64.0 * size / 4096.0, dye_radius=250.0 * size / 4096.0, orbit_seed=1,
                    init="smooth")


def cfg4_workload(size: int = 16384) -> Workload:
    """config 4 - 16384^2 strong scaling (as config 3, no dye)."""
    return Workload(f"{size}sq-N100-strong", size, size, 100, 1.0, 0.0,
                    circles=[(0.5, 0.5, 256.0 * size / 4096.0)], splats_per_step=8,
                    splat_radius=64.0 * size / 4096.0, orbit_seed=1, init="smooth")


def cfg5_workload(n_gpus: int = 1, width: int = 32768, rows_per_gpu: int = 4096) -> Workload:
    """config 5 - 32768^2 weak scaling: 32768 x 4096 cells, 200 iterations and 64 circles PER GPU.

    The N-GPU domain is the 1-GPU domain tiled N times along y (the same 64 circles in every 4096-row band),
    so the work per GPU - obstacle load included - does not change with N."""
    height = rows_per_gpu * n_gpus
    rng = np.random.default_rng(2)
    band = [(float(rng.uniform(0.05, 0.95)), float(rng.uniform(0.05, 0.95)), float(rng.uniform(64.0, 512.0)))
            for _ in range(64)]
    circles = [(x, (g + y) / n_gpus, r) for g in range(n_gpus) for (x, y, r) in band]
    return Workload(f"{width}x{height}-N200-weak", width, height, 200, 1.0, 0.0, circles=circles, init="smooth")


def smooth_velocity(width: int, height: int, row0: int = 0, rows: Optional[int] = None) -> np.ndarray:
    """v0 = 0.5 (sin(2pi 3x/W) cos(2pi 2y/H), -cos(2pi 3x/W) sin(2pi 2y/H)), float64 -> float32."""
    rows = height if rows is None else rows
    x = np.arange(width, dtype=np.float64)[None, :]
    y = np.arange(row0, row0 + rows, dtype=np.float64)[:, None]
    ax = 2.0 * np.pi * 3.0 * x / width
    ay = 2.0 * np.pi * 2.0 * y / height
    v = np.empty((rows, width, 2), np.float32)
    v[..., 0] = (0.5 * np.sin(ax) * np.cos(ay)).astype(np.float32)
    v[..., 1] = (-0.5 * np.cos(ax) * np.sin(ay)).astype(np.float32)
    return v


def random_velocity(width: int, height: int, seed: int = 0) -> np.ndarray:
    rng = np.random.default_rng(seed)
    return (0.5 * rng.uniform(-1.0, 1.0, (height, width, 2))).astype(np.float32)


def initial_velocity(w: Workload) -> Optional[np.ndarray]:
    if w.init == "smooth":
        return smooth_velocity(w.width, w.height)
    if w.init == "random":
        return random_velocity(w.width, w.height)
    return None


def set_velocity(sim, v0: np.ndarray) -> None:
    if hasattr(sim, "upload"):
        sim.upload("velocity", v0)       # CUDA product (natrix_copy_in)
    else:
        sim.velocity = v0                # oracles


def configure(sim, w: Workload) -> None:
    sim.vorticity = w.vorticity
    sim.viscosity = w.viscosity
    sim.iterations = w.iterations


def orbit_positions(w: Workload, step: int) -> List[Tuple[float, float, float, float]]:
    """(px, py, vx, vy) of the scripted impulses of one step.  Config 1 replays a mouse drag on
    a circle (velocity = 10 x the position delta, simulation_demo.py:233); the seeded configs
    move `splats_per_step` emitters on fixed orbits."""
    out = []
    if w.orbit_seed is None:
        def pos(k):
            return 0.5 + 0.3 * math.cos(0.1 * k), 0.5 + 0.3 * math.sin(0.1 * k)
        (x1, y1), (x0, y0) = pos(step), pos(step - 1)
        out.append((x1, y1, 10.0 * (x1 - x0), 10.0 * (y1 - y0)))
        return out[: w.splats_per_step]
    rng = np.random.default_rng(w.orbit_seed)
    for _ in range(w.splats_per_step):
        cx, cy = rng.uniform(0.25, 0.75, 2)
        rad = rng.uniform(0.05, 0.2)
        om = rng.uniform(0.05, 0.2) * (1 if rng.uniform() < 0.5 else -1)
        ph = rng.uniform(0, 2 * math.pi)
        a1, a0 = ph + om * step, ph + om * (step - 1)
        x1, y1 = cx + rad * math.cos(a1), cy + rad * math.sin(a1)
        x0, y0 = cx + rad * math.cos(a0), cy + rad * math.sin(a0)
        out.append((float(x1), float(y1), float(10.0 * (x1 - x0)), float(10.0 * (y1 - y0))))
    return out


def circles_at(w: Workload, step: int) -> List[Tuple[float, float, float]]:
    """The obstacles of one step: fixed, or (w.drift != 0) moved along +x by `drift` cells per step, wrapping."""
    if not w.drift:
        return w.circles
    return [((px + step * w.drift / w.width) % 1.0, py, r) for (px, py, r) in w.circles]


def run_step(w: Workload, sim, dye, step: int, dt: float = DT) -> None:
    """One frame in the demo's call order (simulation_demo.py:220-237): obstacles ->
    fluid.update -> dye.update -> impulses."""
    for (px, py, r) in circles_at(w, step):
        sim.add_circle_obstacle((px, py), r)
    sim.update(dt)
    if dye is not None:
        dye.update(dt)
    for (px, py, vx, vy) in orbit_positions(w, step):
        sim.add_velocity((px, py), (vx, vy), w.splat_radius)
        if dye is not None:
            dye.add_particles((px, py), w.dye_radius, w.dye_strength)


def build(w: Workload, sim_cls: Callable, dye_cls: Optional[Callable], **sim_kwargs):
    sim = sim_cls(w.width, w.height, None, **sim_kwargs)
    configure(sim, w)
    v0 = initial_velocity(w)
    if v0 is not None:
        set_velocity(sim, v0)
    dye = None
    if w.dye_size and dye_cls is not None:
        dye = dye_cls(w.dye_size[0], w.dye_size[1], sim, None)
        dye.dissipation = w.dye_dissipation
    return sim, dye


def fields_of(sim, dye=None) -> dict:
    """All fields as NumPy arrays, whichever engine."""
    if hasattr(sim, "download"):
        out = {k: sim.download(k) for k in ("velocity", "pressure", "divergence", "vorticity")}
        if dye is not None:
            out["dye"] = dye.download()
        return out
    out = {"velocity": np.array(sim.velocity), "pressure": np.array(sim.pressure),
           "divergence": np.array(sim.divergence), "vorticity": np.array(sim.vorticity_field)}
    if dye is not None:
        out["dye"] = np.array(dye.particles)
    return out


def random_scenario(seed: int, max_size: int = 160):
    """A seeded script of public-API calls (sizes, parameters, obstacles, impulses, dye splats, per-frame dt)
    for differential tests: ``play_scenario`` replays it on any engine (CUDA product, NumPy oracle, C oracle)."""
    rng = np.random.default_rng(1000 + seed)
    w = int(rng.choice([int(rng.integers(1, max_size)), 16 * int(rng.integers(16, max(17, max_size // 8 + 17)))]))
    h = int(rng.integers(1, max_size))
    scn = {
        "size": (w, h), "dye_size": (int(rng.integers(1, 2 * max_size)), int(rng.integers(1, 2 * max_size))),
        "iterations": int(rng.integers(1, 30)), "speed": float(rng.choice([1.0, 120.0, 500.0, 1000.0])),
        "dissipation": float(rng.choice([1.0, 0.97, 0.5])), "vorticity": float(rng.choice([0.0, 1.0, 7.5])),
        "viscosity": float(rng.choice([0.0, 0.1, 3.0])), "has_borders": bool(rng.integers(0, 2)),
        "dye_dissipation": float(rng.choice([1.0, 0.98])), "v0_scale": float(rng.choice([0.3, 1.0, 1.7])),
        "v0_seed": int(rng.integers(0, 2**31)), "frames": [],
    }
    for k in range(int(rng.integers(2, 5))):
        frame = {"dt": 0.0 if (k == 0 and rng.uniform() < 0.3) else float(np.float32(rng.choice([1 / 60, 1 / 30, 0.004]))),
                 "circles": [], "triangles": [], "splats": [], "dye": []}
        for _ in range(int(rng.integers(0, 4))):
            frame["circles"].append((float(rng.uniform(-0.1, 1.1)), float(rng.uniform(-0.1, 1.1)),
                                     float(rng.choice([0.0, 1.5, rng.uniform(1, 0.4 * max(w, h) + 2)])), bool(rng.integers(0, 2))))
        for _ in range(int(rng.integers(0, 2))):
            frame["triangles"].append(tuple(float(v) for v in rng.uniform(-0.1, 1.1, 6)) + (bool(rng.integers(0, 2)),))
        for _ in range(int(rng.integers(0, 4))):
            frame["splats"].append((float(rng.uniform(0, 1)), float(rng.uniform(0, 1)), float(rng.uniform(-3, 3)),
                                    float(rng.uniform(-3, 3)), float(rng.uniform(0.5, 0.3 * max(w, h) + 1))))
        for _ in range(int(rng.integers(0, 3))):
            frame["dye"].append((float(rng.uniform(0, 1)), float(rng.uniform(0, 1)), float(rng.uniform(1, max_size / 2)),
                                 float(rng.uniform(0.01, 400.0))))
        scn["frames"].append(frame)
    return scn


def play_scenario(scn, sim_cls: Callable, dye_cls: Optional[Callable], on_frame: Optional[Callable] = None):
    """Replay a ``random_scenario`` through the reference's public API (demo call order); returns (sim, dye)."""
    w, h = scn["size"]
    sim = sim_cls(w, h, None)
    for p in ("iterations", "speed", "dissipation", "vorticity", "viscosity", "has_borders"):
        setattr(sim, p, scn[p])
    rng = np.random.default_rng(scn["v0_seed"])
    set_velocity(sim, (scn["v0_scale"] * rng.uniform(-1.0, 1.0, (h, w, 2))).astype(np.float32))
    dye = None
    if dye_cls is not None:
        dye = dye_cls(scn["dye_size"][0], scn["dye_size"][1], sim, None)
        dye.dissipation, dye.speed = scn["dye_dissipation"], scn["speed"]
    for k, f in enumerate(scn["frames"]):
        for (px, py, r, static) in f["circles"]:
            sim.add_circle_obstacle((px, py), r, static)
        for (*p, static) in f["triangles"]:
            sim.add_triangle_obstacle((p[0], p[1]), (p[2], p[3]), (p[4], p[5]), static)
        sim.update(f["dt"])
        if dye is not None:
            dye.update(f["dt"])
        for (px, py, vx, vy, r) in f["splats"]:
            sim.add_velocity((px, py), (vx, vy), r)
        if dye is not None:
            for (px, py, r, s) in f["dye"]:
                dye.add_particles((px, py), r, s)
        if on_frame is not None:
            on_frame(k, sim, dye)
    return sim, dye
