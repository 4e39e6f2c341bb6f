"""natrix_b200 - B200-native implementation of Natrix's per-step stable-fluids pipeline.

``natrix_b200.core.fluid_simulator.FluidSimulator`` and
``natrix_b200.smooth_particles_area.SmoothParticlesArea`` keep the reference's Python API;
all arithmetic runs in libnatrix_b200.so (hand-written sm_100a CUDA, C ABI in
include/natrix_b200.h).  The ``natrix`` package at the repository root re-exports the same
classes under the reference's import path.
"""
from natrix_b200._lib import NatrixError  # noqa: F401

__all__ = ["NatrixError"]
__version__ = "0.1.0"
