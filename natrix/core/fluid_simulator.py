"""Same import path as the reference module (ref: natrix/core/fluid_simulator.py)."""
from natrix_b200.core.fluid_simulator import DeviceField, FluidSimulator  # noqa: F401

__all__ = ["FluidSimulator", "DeviceField"]
