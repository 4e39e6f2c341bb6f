"""oracle/_ref front-end: the REFERENCE'S OWN shader text as the parity oracle.  TEST INFRASTRUCTURE ONLY.

``oracle/_ref/libnatrix_ref*.so`` is built by ``oracle/Makefile`` from the 14 compute shaders where they lie
under /root/reference (natrix/core/shaders/originals/*.comp + common.sh + constants.sh,
demo/shaders/shader.{AddParticle,AdvectParticle}.comp), compiled UNMODIFIED as C++ behind
``oracle/ref_shim/bgfx_compute.sh``.  Two ways to drive it:

* ``RefFluidSimulator`` / ``RefSmoothParticlesArea`` (this file): the dispatch sequence of
  ``natrix/core/fluid_simulator.py:116-280`` and ``demo/smooth_particles_area.py:70-104`` restated call by
  call (set uniforms, bind slots, dispatch, flip) against the shim runtime.  This one travels to the GPU box
  together with the prebuilt ``.so`` and is the oracle of the ``-m gpu`` parity tests.
* ``reference_classes()``: the UNMODIFIED reference Python classes, imported from /root/reference with
  ``oracle/ref_shim/fake_modules`` standing in for ``pybgfx`` / ``decouple``.  Only possible where
  /root/reference exists (the build container); ``tests/test_ref_oracle.py`` uses it to prove the restated
  driver issues exactly the reference's dispatches, and ``tests/golden/make_golden_ref.py`` to freeze vectors.

Variants (``lib(variant)``): ``literal`` - float32 linear indices exactly as written (exact up to 2**24 cells,
SURVEY Q1); ``exact`` - indices carried exactly, for larger grids; ``lerp`` - ``mix`` in the HLSL ``lerp``
form ``x + a*(y-x)`` instead of the GLSL specification's ``x*(1-a) + y*a`` (the documented alternative
interpretation; the product and both restated oracles use the GLSL form).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys
from math import ceil
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
REFERENCE_ROOT = Path(os.environ.get("NATRIX_REFERENCE_ROOT", "/root/reference"))
_SO = {"literal": _HERE / "_ref" / "libnatrix_ref.so",
       "exact": _HERE / "_ref" / "libnatrix_ref_exact.so",
       "lerp": _HERE / "_ref" / "libnatrix_ref_lerp.so"}
_libs = {}

GROUP = 16                                  # constants.sh:7, common/constants.py:7
# slot numbering, constants.sh:9-18 / common/constants.py:8-17
VELOCITY_IN, VELOCITY_OUT, PRESSURE_IN, PRESSURE_OUT, VORTICITY, DIVERGENCE, OBSTACLES, GENERIC = 1, 2, 3, 4, 5, 6, 7, 8
PARTICLES_IN, PARTICLES_OUT = 9, 10


def reference_available() -> bool:
    return (REFERENCE_ROOT / "natrix" / "core" / "shaders" / "originals" / "shader.Poisson.comp").is_file()


def available(variant: str = "literal") -> bool:
    return _SO[variant].is_file() or reference_available()


def build() -> None:
    """Compile the shim libraries (needs /root/reference; a no-op for the _ref targets elsewhere)."""
    subprocess.check_call(["make", "-s", "-C", str(_HERE)])


def lib(variant: str = "literal"):
    if variant not in _libs:
        if reference_available():
            build()
        if not _SO[variant].is_file():
            raise FileNotFoundError(f"{_SO[variant]}: oracle/_ref is built from /root/reference (make -C oracle)")
        L = C.CDLL(str(_SO[variant]))
        L.nref_build_info.restype = C.c_char_p
        L.nref_create_uniform.argtypes = [C.c_char_p]
        L.nref_set_uniform.argtypes = [C.c_int, C.POINTER(C.c_float), C.c_int]
        L.nref_create_buffer.argtypes = [C.c_size_t]
        L.nref_buffer_ptr.restype = C.c_void_p
        L.nref_buffer_ptr.argtypes = [C.c_int]
        L.nref_destroy_buffer.argtypes = [C.c_int]
        L.nref_set_buffer.argtypes = [C.c_int, C.c_int]
        L.nref_create_program.argtypes = [C.c_char_p]
        L.nref_dispatch.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int]
        L.nref_set_threads.argtypes = [C.c_int]
        L.nref_max_threads.restype = C.c_int
        L.nref_program_count.restype = C.c_int
        L.nref_program_name.restype = C.c_char_p
        L.nref_program_name.argtypes = [C.c_int]
        _libs[variant] = L
    return _libs[variant]


def _view(L, handle: int, shape) -> np.ndarray:
    ptr = C.cast(L.nref_buffer_ptr(handle), C.POINTER(C.c_float))
    return np.ctypeslib.as_array(ptr, shape=tuple(shape))


class _Runtime:
    """Thin helper over the shim's bgfx-like C calls."""

    def __init__(self, variant: str, threads: int | None):
        self.L = lib(variant)
        if threads:
            self.L.nref_set_threads(int(threads))
        self._uniforms = {}
        self.dispatch_log = None

    def uniform(self, name, *values):
        h = self._uniforms.get(name)
        if h is None:
            h = self._uniforms[name] = self.L.nref_create_uniform(name.encode())
        arr = (C.c_float * len(values))(*values)        # the reference builds a c_float array: double -> float32 here
        self.L.nref_set_uniform(h, arr, len(values))

    def program(self, name):
        p = self.L.nref_create_program(name.encode())
        if p < 0:
            raise FileNotFoundError(name)
        return (p, name)

    def dispatch(self, program, gx, gy):
        if self.dispatch_log is not None:
            self.dispatch_log.append(program[1])
        if self.L.nref_dispatch(program[0], gx, gy, 1) != 0:
            raise RuntimeError(program[1])


class RefFluidSimulator:
    """The dispatch sequence of natrix/core/fluid_simulator.py over the compiled reference shaders.

    Buffers: velocity x2 and obstacles 2 floats per cell, divergence and vorticity 1; PRESSURE 2 floats per cell -
    the shaders declare it ``vec2`` (shader.Poisson.comp:11,13) and the reference's 24x over-allocation
     is what makes that stride fit; component x is the pressure."""

    def __init__(self, width, height, vertex_layout=None, variant: str = "literal", threads: int | None = None):
        self.rt = _Runtime(variant, threads)
        L = self.rt.L
        self.width, self.height = int(width), int(height)
        self.speed, self.iterations, self.dissipation = 500.0, 50, 1.0      # fluid_simulator.py:28-32
        self.vorticity, self.viscosity = 0.0, 0.1
        self.has_borders, self.simulate = True, True                        # :34-35
        self.warm_start = False         # opt-in extension of the product (NATRIX_OPT_WARM_START), not the reference
        self.VELOCITY_READ, self.VELOCITY_WRITE = 0, 1                      # :16-17
        self.PRESSURE_READ, self.PRESSURE_WRITE = 0, 1                      # :19-20
        # _load_compute_kernels, :370-442
        names = ["AddVelocity", "InitBoundaries", "AdvectVelocity", "Divergence", "Poisson", "SubtractGradient",
                 "CalcVorticity", "ApplyVorticity", "AddCircleObstacle", "AddTriangleObstacle", "ClearBuffer", "Viscosity"]
        self.k = {n: self.rt.program(f"shader.{n}.comp") for n in names}
        # _set_size, :282-290
        self._num_cells = self.width * self.height
        self.gx = int(ceil(float(self.width) / float(GROUP)))
        self.gy = int(ceil(float(self.height) / float(GROUP)))
        # _create_buffers, :357-368 (zero-filled)
        c = self._num_cells
        self._velocity_buffer = [L.nref_create_buffer(8 * c), L.nref_create_buffer(8 * c)]
        self._pressure_buffer = [L.nref_create_buffer(8 * c), L.nref_create_buffer(8 * c)]
        self._divergence_buffer = L.nref_create_buffer(4 * c)       # bound to slot 5 = _Vorticity in the shaders (Q4)
        self._vorticity_buffer = L.nref_create_buffer(4 * c)        # bound to slot 6 = _Divergence in the shaders (Q4)
        self._obstacles_buffer = L.nref_create_buffer(8 * c)
        self._init_compute_kernels()

    # ------------------------------------------------------------------ fluid_simulator.py:338-355
    def _init_compute_kernels(self):
        L = self.rt.L
        self.rt.uniform("_Size", self.width, self.height)
        L.nref_set_buffer(VELOCITY_IN, self._velocity_buffer[self.VELOCITY_READ])
        L.nref_set_buffer(VELOCITY_OUT, self._velocity_buffer[self.VELOCITY_WRITE])
        L.nref_set_buffer(PRESSURE_IN, self._pressure_buffer[self.PRESSURE_READ])
        L.nref_set_buffer(PRESSURE_OUT, self._pressure_buffer[self.PRESSURE_WRITE])
        L.nref_set_buffer(5, self._divergence_buffer)
        L.nref_set_buffer(6, self._vorticity_buffer)
        L.nref_set_buffer(OBSTACLES, self._obstacles_buffer)

    # ------------------------------------------------------------------ :315-336
    def _update_params(self, time_delta):
        self.rt.uniform("_ElapsedTime", time_delta)
        self.rt.uniform("_Speed", self.speed)
        self.rt.uniform("_Dissipation", self.dissipation)
        self.rt.uniform("_VorticityScale", self.vorticity)
        if self.viscosity > 0.0:
            centre_factor = 1.0 / self.viscosity
            stencil_factor = 1.0 / (4.0 + centre_factor)
            self.rt.uniform("_Alpha", centre_factor)
            self.rt.uniform("_rBeta", stencil_factor)

    # ------------------------------------------------------------------ :444-474
    def _flip_velocity_buffer(self):
        self.VELOCITY_READ, self.VELOCITY_WRITE = self.VELOCITY_WRITE, self.VELOCITY_READ
        self.rt.L.nref_set_buffer(VELOCITY_IN, self._velocity_buffer[self.VELOCITY_READ])
        self.rt.L.nref_set_buffer(VELOCITY_OUT, self._velocity_buffer[self.VELOCITY_WRITE])

    def _flip_pressure_buffer(self):
        self.PRESSURE_READ, self.PRESSURE_WRITE = self.PRESSURE_WRITE, self.PRESSURE_READ
        self.rt.L.nref_set_buffer(PRESSURE_IN, self._pressure_buffer[self.PRESSURE_READ])
        self.rt.L.nref_set_buffer(PRESSURE_OUT, self._pressure_buffer[self.PRESSURE_WRITE])

    def _run(self, name):
        self.rt.dispatch(self.k[name], self.gx, self.gy)

    # ------------------------------------------------------------------ :116-131
    def add_velocity(self, position, velocity, radius):
        if self.simulate:
            self._init_compute_kernels()
            self.rt.uniform("_Position", position[0], position[1])
            self.rt.uniform("_Value", velocity[0], velocity[1])
            self.rt.uniform("_Radius", radius)
            self._run("AddVelocity")
            self._flip_velocity_buffer()

    # ------------------------------------------------------------------ :135-153
    def add_circle_obstacle(self, position, radius, static=False):
        if self.simulate:
            self._init_compute_kernels()
            self.rt.uniform("_Position", position[0], position[1])
            self.rt.uniform("_Radius", radius)
            self.rt.uniform("_Static", 1.0 if static else 0.0)
            self._run("AddCircleObstacle")

    # ------------------------------------------------------------------ :156-172
    def add_triangle_obstacle(self, p1, p2, p3, static=False):
        if self.simulate:
            self._init_compute_kernels()
            self.rt.uniform("_P1", p1[0], p1[1])
            self.rt.uniform("_P2", p2[0], p2[1])
            self.rt.uniform("_P3", p3[0], p3[1])
            self.rt.uniform("_Static", 1.0 if static else 0.0)
            self._run("AddTriangleObstacle")

    # ------------------------------------------------------------------ :174-280
    def update(self, time_delta, skip_pressure_clear: bool = False):
        if not self.simulate:
            return
        L = self.rt.L
        self._init_compute_kernels()
        self._update_params(time_delta)
        if self.has_borders:                                    # :181-188
            self._run("InitBoundaries")
        self._run("AdvectVelocity")                             # :191-198
        self._flip_velocity_buffer()
        self._run("CalcVorticity")                              # :201-207
        self._run("ApplyVorticity")                             # :210-217
        self._flip_velocity_buffer()
        if self.viscosity > 0.0:                                # :220-228
            self._run("Viscosity")
            self._flip_velocity_buffer()
        self._run("Divergence")                                 # :231-233
        if not (skip_pressure_clear or self.warm_start):        # :236-248 (skipped only by the opt-in warm start)
            L.nref_set_buffer(GENERIC, self._pressure_buffer[self.PRESSURE_READ])
            self._run("ClearBuffer")
            L.nref_set_buffer(PRESSURE_IN, self._pressure_buffer[self.PRESSURE_READ])
        for _ in range(int(self.iterations)):                   # :251-255
            self._run("Poisson")
            self._flip_pressure_buffer()
        self._run("SubtractGradient")                           # :258-265
        self._flip_velocity_buffer()
        L.nref_set_buffer(GENERIC, self._obstacles_buffer)      # :268-280
        self._run("ClearBuffer")
        L.nref_set_buffer(OBSTACLES, self._obstacles_buffer)

    def poisson_sweeps(self, n):
        """n Poisson dispatches + flips on the current state (the hot loop alone, :251-255)."""
        self._init_compute_kernels()
        for _ in range(int(n)):
            self._run("Poisson")
            self._flip_pressure_buffer()

    # ------------------------------------------------------------------ test-side views of the buffers
    def _v(self, handle, comps):
        shape = (self.height, self.width, comps) if comps > 1 else (self.height, self.width)
        return _view(self.rt.L, handle, shape)

    @property
    def velocity(self):
        return self._v(self._velocity_buffer[self.VELOCITY_READ], 2)

    @velocity.setter
    def velocity(self, v):
        self._v(self._velocity_buffer[self.VELOCITY_READ], 2)[...] = np.asarray(v, np.float32).reshape(self.height, self.width, 2)

    @property
    def pressure(self):
        return self._v(self._pressure_buffer[self.PRESSURE_READ], 2)[..., 0]

    @pressure.setter
    def pressure(self, p):
        self._v(self._pressure_buffer[self.PRESSURE_READ], 2)[...] = np.asarray(p, np.float32).reshape(self.height, self.width, 1)

    divergence = property(lambda s: s._v(s._vorticity_buffer, 1))          # slot 6 = _Divergence (Q4)
    vorticity_field = property(lambda s: s._v(s._divergence_buffer, 1))    # slot 5 = _Vorticity (Q4)
    obstacles = property(lambda s: s._v(s._obstacles_buffer, 2))

    def destroy(self):                                                     # :476-515
        for h in self._velocity_buffer + self._pressure_buffer + [self._divergence_buffer, self._vorticity_buffer,
                                                                  self._obstacles_buffer]:
            self.rt.L.nref_destroy_buffer(h)
        self._velocity_buffer = self._pressure_buffer = []


class RefSmoothParticlesArea:
    """demo/smooth_particles_area.py over the compiled dye shaders.  Like the reference it never binds slots 1
    and 7: the dye shaders read whatever velocity / obstacle buffers the simulator left bound."""

    def __init__(self, width, height, fluid_simulation: RefFluidSimulator, vertex_layout=None, full_grid_splat=True):
        self.fluid_simulation = fluid_simulation
        self.rt = fluid_simulation.rt
        L = self.rt.L
        self.width, self.height = int(width), int(height)
        self.speed, self.dissipation, self.simulate = 500.0, 1.0, True     # smooth_particles_area.py:22-25
        self.PARTICLES_IN, self.PARTICLES_OUT = 0, 1
        self.k_add = self.rt.program("shader.AddParticle.comp")
        self.k_advect = self.rt.program("shader.AdvectParticle.comp")
        self.gx = int(ceil(float(self.width) / float(GROUP)))              # :109-121
        self.gy = int(ceil(float(self.height) / float(GROUP)))
        # the reference dispatches add_particles over (groups_x, groups_x) (:80-82); every graded
        # configuration is square or landscape, where that covers the grid; `full_grid_splat` is the fix.
        self.full_grid_splat = full_grid_splat
        c = self.width * self.height
        self._particles_buffer = [L.nref_create_buffer(4 * c), L.nref_create_buffer(4 * c)]
        self._init_compute_kernels()

    def _init_compute_kernels(self):                                       # :137-156
        self.rt.uniform("_ParticleSize", self.width, self.height)
        self.rt.uniform("_VelocitySize", self.fluid_simulation.width, self.fluid_simulation.height)
        self.rt.L.nref_set_buffer(PARTICLES_IN, self._particles_buffer[self.PARTICLES_IN])
        self.rt.L.nref_set_buffer(PARTICLES_OUT, self._particles_buffer[self.PARTICLES_OUT])

    def _flip_buffer(self):                                                # :176-190
        self.PARTICLES_IN, self.PARTICLES_OUT = self.PARTICLES_OUT, self.PARTICLES_IN
        self.rt.L.nref_set_buffer(PARTICLES_IN, self._particles_buffer[self.PARTICLES_IN])
        self.rt.L.nref_set_buffer(PARTICLES_OUT, self._particles_buffer[self.PARTICLES_OUT])

    def add_particles(self, position, radius, strength):                   # :70-83
        if self.simulate:
            self._init_compute_kernels()
            self.rt.uniform("_Position", position[0], position[1])
            self.rt.uniform("_Value", strength)
            self.rt.uniform("_Radius", radius)
            self.rt.dispatch(self.k_add, self.gx, self.gy if self.full_grid_splat else self.gx)
            self._flip_buffer()

    def update(self, time_delta):                                          # :85-104
        if self.simulate:
            self._init_compute_kernels()
            self.rt.uniform("_Dissipation", self.dissipation)
            self.rt.uniform("_ElapsedTime", time_delta)
            self.rt.uniform("_Speed", self.speed)
            self.rt.dispatch(self.k_advect, self.gx, self.gy)
            self._flip_buffer()

    @property
    def particles(self):
        return _view(self.rt.L, self._particles_buffer[self.PARTICLES_IN], (self.height, self.width))

    def destroy(self):
        for h in self._particles_buffer:
            self.rt.L.nref_destroy_buffer(h)
        self._particles_buffer = []


# ---------------------------------------------------------------------------- the unmodified reference Python
def reference_classes(variant: str = "literal", layout_stride: int = 8):
    """(FluidSimulator, SmoothParticlesArea, bgfx) - the reference's own classes, imported from
    /root/reference with the pybgfx / decouple stand-ins, wrapped only to add field views.

    ``layout_stride``: bytes per "vertex" of the layout handed to the constructors.  The demo's is 24
    (simulation_demo.py:95-98), which makes every buffer 24x larger than its contents; anything >= 2
    holds the vec2-declared pressure, and 8 keeps large grids affordable."""
    if not reference_available():
        raise FileNotFoundError(REFERENCE_ROOT)
    lib(variant)
    fake = str(_HERE / "ref_shim" / "fake_modules")
    for p in (fake, str(REFERENCE_ROOT)):
        if p not in sys.path:
            sys.path.append(p)
    import pybgfx                                                            # the stand-in
    pybgfx.use_library(_SO[variant])
    # this repo ships shims named `natrix` / `demo` (re-exports of the product); the reference's packages of the
    # same names must win for this import, so load them under private names straight from their files.
    import importlib.util

    def load(name, path, package_paths=None):
        spec = importlib.util.spec_from_file_location(name, path, submodule_search_locations=package_paths)
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
        return mod

    saved = {k: sys.modules.get(k) for k in list(sys.modules) if k == "natrix" or k.startswith("natrix.") or k == "demo"
             or k.startswith("demo.")}
    for k in saved:
        del sys.modules[k]
    try:
        load("natrix", REFERENCE_ROOT / "natrix" / "__init__.py", [str(REFERENCE_ROOT / "natrix")])
        import natrix.core.fluid_simulator as ref_fs                         # noqa: E402  (reference file, unmodified)
        ref_spa = load("natrix_reference_demo_spa", REFERENCE_ROOT / "demo" / "smooth_particles_area.py")
    finally:
        ref_mods = {k: v for k, v in sys.modules.items() if k == "natrix" or k.startswith("natrix.")}
        for k in ref_mods:
            del sys.modules[k]
        sys.modules.update({k: v for k, v in saved.items() if v is not None})
    L = lib(variant)
    layout = pybgfx.bgfx.VertexLayout(layout_stride)

    class ReferenceFluidSimulator(ref_fs.FluidSimulator):
        def __init__(self, width, height, vertex_layout=None):
            super().__init__(width, height, layout if vertex_layout is None else vertex_layout)

        def _v(self, handle, comps):
            shape = (self.height, self.width, comps) if comps > 1 else (self.height, self.width)
            return _view(L, handle.id, shape)

        @property
        def velocity(self):
            return self._v(self._velocity_buffer[self.VELOCITY_READ], 2)

        @velocity.setter
        def velocity(self, v):
            self.velocity[...] = np.asarray(v, np.float32).reshape(self.height, self.width, 2)

        pressure = property(lambda s: s._v(s._pressure_buffer[s.PRESSURE_READ], 2)[..., 0])
        divergence = property(lambda s: s._v(s._vorticity_buffer, 1))
        vorticity_field = property(lambda s: s._v(s._divergence_buffer, 1))
        obstacles = property(lambda s: s._v(s._obstacles_buffer, 2))

    class ReferenceSmoothParticlesArea(ref_spa.SmoothParticlesArea):
        def __init__(self, width, height, fluid_simulation, vertex_layout=None):
            super().__init__(width, height, fluid_simulation, layout if vertex_layout is None else vertex_layout)

        @property
        def particles(self):
            return _view(L, self._particles_buffer[self.PARTICLES_IN].id, (self._height, self._width))

    return ReferenceFluidSimulator, ReferenceSmoothParticlesArea, pybgfx.bgfx


def max_threads(variant: str = "literal") -> int:
    return int(lib(variant).nref_max_threads())
