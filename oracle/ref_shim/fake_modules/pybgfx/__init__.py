"""Stand-in for the `pybgfx` package (bgfx-python 2.0.1, poetry.lock:76-86) - TEST INFRASTRUCTURE ONLY.

Exactly the names natrix/core/fluid_simulator.py, natrix/core/utils/shaders_utils.py and
demo/smooth_particles_area.py import, routed to oracle/_ref/libnatrix_ref*.so (the reference's own
shader text compiled as C++, see ../bgfx_compute.sh and ../runtime.cpp).  With this directory on
sys.path the UNMODIFIED reference classes run here; oracle/natrix_ref.py does the plumbing.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

_lib = None


def use_library(path) -> None:
    """Select the shim build (literal float indexing / exact indexing / lerp-form mix)."""
    global _lib
    L = C.CDLL(str(path))
    L.nref_build_info.restype = C.c_char_p
    L.nref_create_uniform.argtypes = [C.c_char_p]
    L.nref_set_uniform.argtypes = [C.c_int, C.POINTER(C.c_float), C.c_int]
    L.nref_create_buffer.argtypes = [C.c_size_t]
    L.nref_buffer_ptr.restype = C.c_void_p
    L.nref_buffer_ptr.argtypes = [C.c_int]
    L.nref_buffer_bytes.restype = C.c_size_t
    L.nref_buffer_bytes.argtypes = [C.c_int]
    L.nref_destroy_buffer.argtypes = [C.c_int]
    L.nref_set_buffer.argtypes = [C.c_int, C.c_int]
    L.nref_create_program.argtypes = [C.c_char_p]
    L.nref_dispatch.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int]
    L.nref_set_threads.argtypes = [C.c_int]
    L.nref_max_threads.restype = C.c_int
    _lib = L


def library():
    if _lib is None:
        raise RuntimeError("pybgfx stand-in: call use_library(<oracle/_ref/libnatrix_ref*.so>) first")
    return _lib


class _Handle:
    kind = "?"

    def __init__(self, ident, **kw):
        self.id = ident
        self.__dict__.update(kw)


class UniformHandle(_Handle):
    kind = "uniform"


class BufferHandle(_Handle):
    kind = "buffer"


class ShaderHandle(_Handle):
    kind = "shader"


class ProgramHandle(_Handle):
    kind = "program"


class _VertexLayout:
    """bgfx.VertexLayout: only the stride matters (createDynamicVertexBuffer allocates num * stride bytes).
    `begin().add(attrib, num, type).end()` as in demo/simulation_demo.py:95-98 (Float attributes)."""

    def __init__(self, stride: int = 0):
        self.stride = stride

    def begin(self):
        self.stride = 0
        return self

    def add(self, attrib, num, attrib_type, normalized=False, as_int=False):
        self.stride += 4 * int(num)
        return self

    def end(self):
        return self


class _Enum:
    def __init__(self, **kw):
        self.__dict__.update(kw)


class _Bgfx:
    VertexLayout = _VertexLayout
    UniformType = _Enum(Sampler=0, End=1, Vec4=2, Mat3=3, Mat4=4)
    Access = _Enum(Read=0, Write=1, ReadWrite=2)
    Attrib = _Enum(Position=0, TexCoord0=10)
    AttribType = _Enum(Float=4)

    dispatch_log = None          # optional list: names of the programs dispatched, in order

    @staticmethod
    def createUniform(name, uniform_type, num=1):
        return UniformHandle(library().nref_create_uniform(name.encode()), name=name)

    @staticmethod
    def setUniform(handle, value, num=1):
        # `value` is what as_void_ptr() returned: the ctypes float array itself.  bgfx would copy a whole
        # vec4 from that address (past the end of a 1- or 2-float array); only the components
        # the array holds are defined, and only those are ever declared by the shaders.
        n = len(value)
        library().nref_set_uniform(handle.id, C.cast(value, C.POINTER(C.c_float)), n)

    @staticmethod
    def createDynamicVertexBuffer(num, layout, flags=0):
        nbytes = int(num) * int(layout.stride)
        ident = library().nref_create_buffer(nbytes)
        if ident < 0:
            raise MemoryError(nbytes)
        return BufferHandle(ident, bytes=nbytes)

    @staticmethod
    def setBuffer(stage, handle, access):
        library().nref_set_buffer(int(stage), handle.id)

    @staticmethod
    def createProgram(shader, destroy_shaders=False):
        ident = library().nref_create_program(shader.name.encode())
        if ident < 0:
            raise FileNotFoundError(f"shader {shader.name} is not part of this shim build")
        return ProgramHandle(ident, name=shader.name)

    @classmethod
    def dispatch(cls, view_id, program, num_x=1, num_y=1, num_z=1, flags=0):
        if cls.dispatch_log is not None:
            cls.dispatch_log.append(program.name)
        rc = library().nref_dispatch(program.id, int(num_x), int(num_y), int(num_z))
        if rc != 0:
            raise RuntimeError(f"dispatch of {program.name} failed")

    @staticmethod
    def destroy(handle):
        if handle.kind == "buffer" and handle.id >= 0:
            library().nref_destroy_buffer(handle.id)
            handle.id = -1


bgfx = _Bgfx
