"""Drop-in ``SmoothParticlesArea`` (the dye field) backed by libnatrix_b200.so.

Mirrors ref: demo/smooth_particles_area.py:15-211 - constructor ``(width, height,
fluid_simulation, vertex_layout)``, validated ``speed`` / ``dissipation`` properties, the
``simulate`` attribute, ``add_particles``, ``update`` and ``destroy``.  The dye grid may have a
different resolution from the velocity grid; it reads the simulator's current velocity and
obstacle buffers through the simulator handle (the reference relied on bgfx slot bindings the
simulator left behind).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from natrix_b200 import _lib as L
from natrix_b200.core.fluid_simulator import DeviceField, FluidSimulator


class SmoothParticlesArea:
    PARTICLES_IN = 0
    PARTICLES_OUT = 1

    simulate = True

    def __init__(self, width: int, height: int, fluid_simulation: FluidSimulator, vertex_layout=None, slab=None):
        """``slab=(row0, rows, halo)`` holds one row slab of the ``width x height`` dye grid on a simulator
        slab (multi-GPU, natrix_b200.slabs.SlabSmoothParticlesArea); buffers then cover ``rows`` rows."""
        self.fluid_simulation = fluid_simulation
        self.vertex_layout = vertex_layout
        self._width = int(width)
        self._height = int(height)
        self._speed = 500.0
        self._dissipation = 1.0
        self._lib = L.lib()
        handle = C.c_void_p()
        if slab is None:
            self._rows = self._height
            L.check(self._lib.natrix_dye_create(fluid_simulation._handle(), self._width, self._height,
                                                C.byref(handle)))
        else:
            row0, rows, halo = (int(v) for v in slab)
            self._rows = rows
            L.check(self._lib.natrix_dye_create_slab(fluid_simulation._handle(), self._width, self._height, row0, rows,
                                                     halo, C.byref(handle)))
        self._h = handle
        fluid_simulation._dyes.append(self)

    @property
    def width(self):
        return self._width

    @property
    def height(self):
        return self._height

    @property
    def speed(self):
        return self._speed

    @speed.setter
    def speed(self, value):
        if value > 0:
            self._speed = value
        else:
            raise ValueError("'Speed' should be greater than zero")

    @property
    def dissipation(self):
        return self._dissipation

    @dissipation.setter
    def dissipation(self, value):
        if value > 0:
            self._dissipation = value
        else:
            raise ValueError("'Dissipation' should be grater than zero")

    def _handle(self):
        if not self._h:
            raise L.NatrixError(-3, "dye field was destroyed")
        return self._h

    def add_particles(self, position: tuple, radius: float, strength: float):
        if self.simulate:
            L.check(self._lib.natrix_dye_add(self._handle(), position[0], position[1], radius, strength))

    def update(self, time_delta: float):
        if self.simulate:
            L.check(self._lib.natrix_dye_step(self._handle(), time_delta, self._speed, self._dissipation))

    def destroy(self):
        if self._h:
            self._lib.natrix_dye_destroy(self._h)
            self._h = None
            if self in self.fluid_simulation._dyes:
                self.fluid_simulation._dyes.remove(self)

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass

    # ---- extensions (the reference renders the buffer on the GPU and never reads it back)
    def get_particles_buffer(self) -> DeviceField:
        ptr, nbytes = C.c_void_p(), C.c_size_t()
        L.check(self._lib.natrix_dye_field_ptr(self._handle(), C.byref(ptr), C.byref(nbytes)))
        return DeviceField(ptr.value, (self._rows, self._width), np.float32, self,
                           self.fluid_simulation.cuda_stream)

    @property
    def device(self):
        return self.fluid_simulation.device

    def _download_ptr(self, fld):
        return self.download()

    def download(self) -> np.ndarray:
        out = np.empty((self._rows, self._width), np.float32)
        L.check(self._lib.natrix_dye_copy_out(self._handle(), out.ctypes.data_as(C.c_void_p), out.nbytes))
        return out

    def upload(self, array) -> None:
        arr = np.ascontiguousarray(array, dtype=np.float32).reshape(self._rows, self._width)
        L.check(self._lib.natrix_dye_copy_in(self._handle(), arr.ctypes.data_as(C.c_void_p), arr.nbytes))

    def export_rgba8(self) -> np.ndarray:
        """The dye as the (H, W, 4) uint8 image the demo's compute shader writes for rendering
        (ref: demo/shaders/demo.ComputeShader.comp:9-21)."""
        out = np.empty((self._rows, self._width, 4), np.uint8)
        L.check(self._lib.natrix_dye_export_rgba8(self._handle(), out.ctypes.data_as(C.c_void_p), out.nbytes, 0))
        return out

    def render_frame(self, quiver_tile: float = 0.0) -> np.ndarray:
        """The frame the demo draws, as an (H, W, 4) uint8 image, top row first: the dye through the plasma / fbm
        colour map over the clear colour, plus the velocity arrows when ``quiver_tile`` > 0
        (ref: demo/shaders/demo.FieldFragmentShader.frag, demo.QuiverFragmentShader.frag, simulation_demo.py:249-281)."""
        out = np.empty((self._rows, self._width, 4), np.uint8)
        L.check(self._lib.natrix_render_frame(self._handle(), out.ctypes.data_as(C.c_void_p), out.nbytes, 0,
                                              float(quiver_tile)))
        return out

    def stats(self):
        out = (C.c_double * 4)()
        L.check(self._lib.natrix_dye_stats(self._handle(), out))
        return tuple(out)
