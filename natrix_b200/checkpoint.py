"""Checkpoint / restore of a simulation (SURVEY 5 "checkpoint/resume": absent in the reference, whose
buffers never leave the GPU).

What a step carries over is small: the READ velocity, the obstacle map stamped since the last update, the
pressure (only read again when warm starting), the simulator's parameters, and per dye field the dye
buffer with its two parameters.  Divergence and vorticity are outputs of the next step.  A restored run
continues bit-identically (tests/test_gpu_parity.py, tests/test_host_logic.py).

Engine-agnostic: works on the CUDA classes (download / upload through the C ABI) and on the oracle
classes (plain arrays)."""
from __future__ import annotations

import numpy as np

FORMAT = 1
_PARAMS = ("speed", "iterations", "dissipation", "vorticity", "viscosity", "has_borders", "simulate")


def _get(sim, name):
    if hasattr(sim, "download"):
        return sim.download(name)
    return np.array(getattr(sim, name), copy=True)


def _set(sim, name, value):
    if hasattr(sim, "upload"):
        sim.upload(name, value)
    elif name == "pressure":                       # the oracle's pressure has no setter: write the READ buffer
        sim._p[sim.PRESSURE_READ] = np.array(value, copy=True)
    else:
        setattr(sim, name, np.array(value, copy=True))


def _dye_get(dye):
    return dye.download() if hasattr(dye, "download") else np.array(dye.particles, copy=True)


def _dye_set(dye, value):
    if hasattr(dye, "upload"):
        dye.upload(value)
    else:
        dye.particles = np.array(value, copy=True)


def state_dict(sim, dyes=()) -> dict:
    """Everything the next step depends on, as host arrays / scalars."""
    state = {"format": np.int64(FORMAT), "size": np.array([sim.width, sim.height], np.int64)}
    for p in _PARAMS:
        state[f"param_{p}"] = np.float64(getattr(sim, p))
    for f in ("velocity", "pressure", "obstacles"):
        state[f] = _get(sim, f)
    state["n_dyes"] = np.int64(len(dyes))
    for i, d in enumerate(dyes):
        state[f"dye{i}"] = _dye_get(d)
        state[f"dye{i}_params"] = np.array([d.speed, d.dissipation, float(d.simulate)], np.float64)
    return state


def load_state_dict(state, sim, dyes=()) -> None:
    if int(state["format"]) != FORMAT:
        raise ValueError(f"unknown checkpoint format {int(state['format'])}")
    if tuple(int(v) for v in state["size"]) != (sim.width, sim.height):
        raise ValueError(f"checkpoint is {tuple(int(v) for v in state['size'])}, simulator is {(sim.width, sim.height)}")
    if int(state["n_dyes"]) != len(dyes):
        raise ValueError(f"checkpoint holds {int(state['n_dyes'])} dye fields, {len(dyes)} given")
    for p in _PARAMS:
        v = float(state[f"param_{p}"])
        setattr(sim, p, int(v) if p == "iterations" else bool(v) if p in ("has_borders", "simulate") else v)
    for f in ("velocity", "pressure", "obstacles"):
        _set(sim, f, state[f])
    for i, d in enumerate(dyes):
        if state[f"dye{i}"].shape != (d.height, d.width):
            raise ValueError(f"dye field {i}: checkpoint is {state[f'dye{i}'].shape}, field is {(d.height, d.width)}")
        _dye_set(d, state[f"dye{i}"])
        d.speed, d.dissipation, d.simulate = float(state[f"dye{i}_params"][0]), float(state[f"dye{i}_params"][1]), bool(
            state[f"dye{i}_params"][2])


def save(path, sim, dyes=()) -> None:
    np.savez(path, **state_dict(sim, dyes))


def load(path, sim, dyes=()) -> None:
    with np.load(path) as z:
        load_state_dict({k: z[k] for k in z.files}, sim, dyes)
