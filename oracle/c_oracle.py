"""ctypes front-end for oracle/natrix_oracle.c (TEST INFRASTRUCTURE ONLY).

Mirrors the surface of ``OracleFluidSimulator`` / ``OracleSmoothParticlesArea`` so tests can
run the two restatements side by side and bench.py can time the multi-threaded CPU baseline.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_SO = _HERE / "_build" / "libnatrix_oracle.so"


def build(force: bool = False) -> Path:
    src = _HERE / "natrix_oracle.c"
    if force or not _SO.exists() or _SO.stat().st_mtime < src.stat().st_mtime:
        subprocess.check_call(["make", "-s", "-C", str(_HERE)])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(str(_SO))
        vp, f, i, d = C.c_void_p, C.c_float, C.c_int, C.c_double
        L.nox_create.restype = vp
        L.nox_create.argtypes = [i, i]
        L.nox_destroy.argtypes = [vp]
        L.nox_set_params.argtypes = [vp, f, i, f, f, d, i]
        L.nox_field.restype = C.POINTER(C.c_float)
        L.nox_field.argtypes = [vp, i]
        L.nox_add_velocity.argtypes = [vp, f, f, f, f, f]
        L.nox_add_circle_obstacle.argtypes = [vp, f, f, f, i]
        L.nox_add_triangle_obstacle.argtypes = [vp, f, f, f, f, f, f, i]
        L.nox_step.argtypes = [vp, f]
        L.nox_poisson_sweeps.argtypes = [vp, i]
        L.nox_dye_create.restype = vp
        L.nox_dye_create.argtypes = [vp, i, i]
        L.nox_dye_destroy.argtypes = [vp]
        L.nox_dye_field.restype = C.POINTER(C.c_float)
        L.nox_dye_field.argtypes = [vp]
        L.nox_dye_add.argtypes = [vp, f, f, f, f]
        L.nox_dye_step.argtypes = [vp, f, f, f]
        L.nox_set_threads.argtypes = [i]
        L.nox_max_threads.restype = i
        _lib = L
    return _lib


VELOCITY, PRESSURE, DIVERGENCE, VORTICITY, OBSTACLES = range(5)


class COracleFluidSimulator:
    def __init__(self, width, height, vertex_layout=None, threads: int | None = None):
        self._L = lib()
        if threads:
            self._L.nox_set_threads(int(threads))
        self.width, self.height = int(width), int(height)
        self._h = self._L.nox_create(self.width, self.height)
        self.speed, self.iterations, self.dissipation = 500.0, 50, 1.0
        self.vorticity, self.viscosity = 0.0, 0.1
        self.has_borders, self.simulate = True, True

    def _push(self):
        self._L.nox_set_params(self._h, self.speed, int(self.iterations), self.dissipation,
                               self.vorticity, float(self.viscosity), int(bool(self.has_borders)))

    def _view(self, field, comps):
        ptr = self._L.nox_field(self._h, field)
        shape = (self.height, self.width, comps) if comps > 1 else (self.height, self.width)
        return np.ctypeslib.as_array(ptr, shape=shape)

    @property
    def velocity(self):
        return self._view(VELOCITY, 2)

    @velocity.setter
    def velocity(self, v):
        self._view(VELOCITY, 2)[...] = np.asarray(v, np.float32).reshape(self.height, self.width, 2)

    pressure = property(lambda s: s._view(PRESSURE, 1))
    divergence = property(lambda s: s._view(DIVERGENCE, 1))
    vorticity_field = property(lambda s: s._view(VORTICITY, 1))
    obstacles = property(lambda s: s._view(OBSTACLES, 2))

    def add_velocity(self, position, velocity, radius):
        if self.simulate:
            self._L.nox_add_velocity(self._h, position[0], position[1], velocity[0], velocity[1], radius)

    def add_circle_obstacle(self, position, radius, static=False):
        if self.simulate:
            self._L.nox_add_circle_obstacle(self._h, position[0], position[1], radius, int(static))

    def add_triangle_obstacle(self, p1, p2, p3, static=False):
        if self.simulate:
            self._L.nox_add_triangle_obstacle(self._h, p1[0], p1[1], p2[0], p2[1], p3[0], p3[1], int(static))

    def update(self, time_delta):
        if self.simulate:
            self._push()
            self._L.nox_step(self._h, time_delta)

    def poisson_sweeps(self, n):
        self._L.nox_poisson_sweeps(self._h, int(n))

    def destroy(self):
        if self._h:
            self._L.nox_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


class COracleSmoothParticlesArea:
    def __init__(self, width, height, fluid_simulation: COracleFluidSimulator, vertex_layout=None):
        self._L = lib()
        self.width, self.height = int(width), int(height)
        self.fluid_simulation = fluid_simulation
        self._h = self._L.nox_dye_create(fluid_simulation._h, self.width, self.height)
        self.speed, self.dissipation, self.simulate = 500.0, 1.0, True

    @property
    def particles(self):
        return np.ctypeslib.as_array(self._L.nox_dye_field(self._h), shape=(self.height, self.width))

    def add_particles(self, position, radius, strength):
        if self.simulate:
            self._L.nox_dye_add(self._h, position[0], position[1], radius, strength)

    def update(self, time_delta):
        if self.simulate:
            self._L.nox_dye_step(self._h, time_delta, self.speed, self.dissipation)

    def destroy(self):
        if self._h:
            self._L.nox_dye_destroy(self._h)
            self._h = None


def max_threads() -> int:
    return int(lib().nox_max_threads())


def host_cores() -> int:
    return os.cpu_count() or 1
