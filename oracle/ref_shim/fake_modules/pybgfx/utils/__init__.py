"""pybgfx.utils stand-in (natrix/core/fluid_simulator.py:6)."""


def as_void_ptr(obj):
    """The real helper casts a ctypes object to `void*`; the stand-in's setUniform needs the element count
    too, so the array is passed through unchanged."""
    return obj
