"""Drop-in ``FluidSimulator`` backed by libnatrix_b200.so (hand-written sm_100a CUDA kernels).

Mirrors the public surface of the reference class (ref: natrix/core/fluid_simulator.py:15-515):
constructor ``(width, height, vertex_layout)``, the validated properties ``speed``,
``iterations``, ``dissipation``, ``vorticity``, ``viscosity`` (same ValueError messages,
ref :58-111), the attributes ``has_borders`` / ``simulate``, and the methods ``add_velocity``,
``add_circle_obstacle``, ``add_triangle_obstacle``, ``update``, ``get_velocity_buffer`` and
``destroy``.  Where the reference enqueued bgfx dispatches, this class makes one C-ABI call.

Additions (the reference has no host upload / readback at all): ``download``,
``upload``, ``stats``, ``synchronize`` and the ``DeviceField`` objects returned by
``get_velocity_buffer`` / ``field`` (zero-copy: ``__cuda_array_interface__``, DLPack via torch).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np

from natrix_b200 import _lib as L

_FIELD_IDS = {
    "velocity": L.VELOCITY,
    "pressure": L.PRESSURE,
    "divergence": L.DIVERGENCE,
    "vorticity": L.VORTICITY,
    "obstacles": L.OBSTACLES,
    "nbmask": L.NBMASK,
    "div4": L.DIV4,          # internal: 0.25 * divergence, the copy the Jacobi sweeps read
}


class DeviceField:
    """A view of one simulator field in GPU memory (what ``get_velocity_buffer`` hands out in
    place of the reference's bgfx buffer handle, ref: fluid_simulator.py:113-114).

    The pointer addresses the simulator's CURRENT read buffer and is invalidated by the next
    mutating call on the simulator (ping-pong flip), exactly like the bgfx handle was.
    """

    def __init__(self, ptr: int, shape: Tuple[int, ...], dtype, owner, stream: int):
        self.data_ptr = int(ptr)
        self.shape = tuple(shape)
        self.dtype = np.dtype(dtype)
        self._owner = owner
        self._stream = stream

    @property
    def nbytes(self) -> int:
        return int(np.prod(self.shape)) * self.dtype.itemsize

    @property
    def __cuda_array_interface__(self):
        return {
            "shape": self.shape,
            "typestr": self.dtype.str,
            "data": (self.data_ptr, False),
            "version": 3,
            "strides": None,
            "stream": self._stream or None,
        }

    def torch(self):
        """Zero-copy ``torch.Tensor`` aliasing the field (torch is optional plumbing)."""
        import torch

        return torch.as_tensor(self, device=f"cuda:{self._owner.device}")

    def __dlpack__(self, stream=None):
        return self.torch().__dlpack__(stream=stream)

    def __dlpack_device__(self):
        return self.torch().__dlpack_device__()

    def numpy(self) -> np.ndarray:
        """Synchronising device -> host copy."""
        return self._owner._download_ptr(self)


class FluidSimulator:
    VELOCITY_READ = 0
    VELOCITY_WRITE = 1

    PRESSURE_READ = 0
    PRESSURE_WRITE = 1

    has_borders = True
    simulate = True

    def __init__(self, width: int, height: int, vertex_layout=None, device: int = 0,
                 slab: Optional[Tuple[int, int, int]] = None):
        """``vertex_layout`` is accepted for signature compatibility and ignored (it only sized
        bgfx vertex buffers).  ``slab=(row0, rows, halo)`` creates one row slab of a
        ``width x height`` global grid (multi-GPU, see natrix_b200.slabs)."""
        self._width = int(width)
        self._height = int(height)
        self.vertex_layout = vertex_layout
        self.device = int(device)
        self._speed = 500.0
        self._iterations = 50
        self._dissipation = 1.0
        self._vorticity = 0.0
        self._viscosity = 0.1
        self._lib = L.lib()
        handle = C.c_void_p()
        if slab is None:
            L.check(self._lib.natrix_create(self._width, self._height, self.device, C.byref(handle)))
            self._row0, self._rows, self._halo = 0, self._height, 0
        else:
            row0, rows, halo = (int(v) for v in slab)
            L.check(self._lib.natrix_create_slab(self._width, self._height, row0, rows, halo, self.device,
                                                 C.byref(handle)))
            self._row0, self._rows, self._halo = row0, rows, halo
        self._h = handle
        self._num_cells = self._width * self._rows
        self._dyes = []
        stream = C.c_void_p()
        L.check(self._lib.natrix_stream(self._h, C.byref(stream)))
        self.cuda_stream = stream.value or 0

    # ------------------------------------------------------------------ properties
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
    def iterations(self):
        return self._iterations

    @iterations.setter
    def iterations(self, value):
        if value > 0:
            self._iterations = value
        else:
            raise ValueError("'Iterations' should be grater than zero")

    @property
    def dissipation(self):
        return self._dissipation

    @dissipation.setter
    def dissipation(self, value):
        if value > 0:
            self._dissipation = value
        else:
            raise ValueError("'Dissipation' should be grater than zero")

    @property
    def vorticity(self):
        return self._vorticity

    @vorticity.setter
    def vorticity(self, value):
        if value >= 0:
            self._vorticity = value
        else:
            raise ValueError("'Vorticity' should be grater or equal than zero")

    @property
    def viscosity(self):
        return self._viscosity

    @viscosity.setter
    def viscosity(self, value):
        if value >= 0.0:
            self._viscosity = value
        else:
            raise ValueError("'Viscosity' should be greater or equal than zero")

    # ------------------------------------------------------------------ reference methods
    def _handle(self):
        if not self._h:
            raise L.NatrixError(-3, "simulator was destroyed")
        return self._h

    def _push_params(self):
        # ref: _update_params (fluid_simulator.py:315-336); viscosity stays a double so that
        # alpha / rBeta are derived exactly like the reference derives them
        L.check(self._lib.natrix_set_params(self._handle(), self._speed, int(self._iterations),
                                            self._dissipation, self._vorticity, float(self._viscosity),
                                            1 if self.has_borders else 0))

    def get_velocity_buffer(self) -> DeviceField:
        return self.field("velocity")

    def add_velocity(self, position: tuple, velocity: tuple, radius: float):
        if self.simulate:
            L.check(self._lib.natrix_add_velocity(self._handle(), position[0], position[1], velocity[0],
                                                  velocity[1], radius))

    # position in normalised local space, radius in cells (ref :133-134)
    def add_circle_obstacle(self, position: tuple, radius: float, static=False):
        if self.simulate:
            L.check(self._lib.natrix_add_circle_obstacle(self._handle(), position[0], position[1], radius,
                                                         1 if static else 0))

    # points in normalised local space
    def add_triangle_obstacle(self, p1: tuple, p2: tuple, p3: tuple, static=False):
        if self.simulate:
            L.check(self._lib.natrix_add_triangle_obstacle(self._handle(), p1[0], p1[1], p2[0], p2[1], p3[0],
                                                           p3[1], 1 if static else 0))

    def update(self, time_delta: float):
        if self.simulate:
            self._push_params()
            L.check(self._lib.natrix_step(self._handle(), time_delta))

    def destroy(self):
        for dye in list(self._dyes):
            dye.destroy()
        if self._h:
            self._lib.natrix_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass

    # ------------------------------------------------------------------ extensions
    def set_option(self, option: int, value: int):
        L.check(self._lib.natrix_set_option(self._handle(), int(option), int(value)))

    def get_option(self, option: int) -> int:
        out = C.c_int()
        L.check(self._lib.natrix_get_option(self._handle(), int(option), C.byref(out)))
        return out.value

    @property
    def warm_start(self) -> bool:
        """NOT reference behaviour (default False): keep the previous step's pressure as the Jacobi initial
        guess instead of clearing it (fluid_simulator.py:236-248), so that fewer iterations reach the same
        residual."""
        return bool(self.get_option(L.OPT_WARM_START))

    @warm_start.setter
    def warm_start(self, value: bool):
        self.set_option(L.OPT_WARM_START, 1 if value else 0)

    _SOLVERS = ("jacobi", "sor", "multigrid")

    @property
    def solver(self) -> str:
        """NOT reference behaviour (default "jacobi", the reference's loop, fluid_simulator.py:251-255): "sor" runs
        `iterations` red-black SOR sweeps (`sor_omega`), "multigrid" `iterations` V(mg_smooth, mg_smooth) cycles on the same
        linear system.  Full grids, fused pipeline."""
        return self._SOLVERS[self.get_option(L.OPT_SOLVER)]

    @solver.setter
    def solver(self, name: str):
        self.set_option(L.OPT_SOLVER, self._SOLVERS.index(name))

    @property
    def sor_omega(self) -> float:
        return self.get_option(L.OPT_SOR_OMEGA_MILLI) / 1000.0

    @sor_omega.setter
    def sor_omega(self, value: float):
        self.set_option(L.OPT_SOR_OMEGA_MILLI, int(round(value * 1000)))

    @property
    def mg_smooth(self) -> int:
        return self.get_option(L.OPT_MG_SMOOTH)

    @mg_smooth.setter
    def mg_smooth(self, value: int):
        self.set_option(L.OPT_MG_SMOOTH, int(value))

    def _shape(self, fid: int):
        comps = L.FIELD_COMPONENTS[fid]
        return (self._rows, self._width, comps) if comps > 1 else (self._rows, self._width)

    def field(self, name: str) -> DeviceField:
        fid = _FIELD_IDS[name]
        ptr, nbytes = C.c_void_p(), C.c_size_t()
        L.check(self._lib.natrix_field_ptr(self._handle(), fid, C.byref(ptr), C.byref(nbytes)))
        if fid in (L.OBSTACLES, L.NBMASK):      # kept as 1 byte per cell in HBM
            return DeviceField(ptr.value, (self._rows, self._width), np.uint8, self, self.cuda_stream)
        return DeviceField(ptr.value, self._shape(fid), np.float32, self, self.cuda_stream)

    def download(self, name: str) -> np.ndarray:
        """Synchronising copy of a field to a new NumPy array (obstacles come back in the
        reference's float2 encoding)."""
        fid = _FIELD_IDS[name]
        if fid == L.NBMASK:
            out = np.empty((self._rows, self._width), np.uint8)
        else:
            out = np.empty(self._shape(fid), np.float32)
        L.check(self._lib.natrix_copy_out(self._handle(), fid, out.ctypes.data_as(C.c_void_p), out.nbytes))
        return out

    def upload(self, name: str, array) -> None:
        fid = _FIELD_IDS[name]
        dtype = np.uint8 if fid == L.NBMASK else np.float32
        shape = (self._rows, self._width) if fid == L.NBMASK else self._shape(fid)
        arr = np.ascontiguousarray(array, dtype=dtype).reshape(shape)
        L.check(self._lib.natrix_copy_in(self._handle(), fid, arr.ctypes.data_as(C.c_void_p), arr.nbytes))

    def _download_ptr(self, fld: DeviceField) -> np.ndarray:
        for name, fid in _FIELD_IDS.items():
            cur = self.field(name)
            if cur.data_ptr == fld.data_ptr:
                return self.download(name)
        raise L.NatrixError(-3, "stale DeviceField: the simulator has flipped its buffers since")

    def stats(self, name: str):
        """(sum, sum of squares, min, max) of a float field, reduced on the GPU in a fixed order."""
        out = (C.c_double * 4)()
        L.check(self._lib.natrix_field_stats(self._handle(), _FIELD_IDS[name], out))
        return tuple(out)

    def synchronize(self):
        L.check(self._lib.natrix_sync(self._handle()))

    def timings(self):
        ms = (C.c_float * 6)()
        L.check(self._lib.natrix_get_timings(self._handle(), ms, 6))
        return dict(zip(("advect", "forces", "divergence", "jacobi", "gradient", "clear"), ms))

    def plan_cache_stats(self):
        """(hits, misses) of the Jacobi kernel's tile-plan cache (a miss = a new obstacle set)."""
        h, m = C.c_ulonglong(), C.c_ulonglong()
        L.check(self._lib.natrix_debug_plan_stats(self._handle(), C.byref(h), C.byref(m)))
        return h.value, m.value

    @property
    def launch_count(self) -> int:
        out = C.c_ulonglong()
        L.check(self._lib.natrix_launch_count(self._handle(), C.byref(out)))
        return out.value
