"""Import-path shim: ``natrix.core.fluid_simulator.FluidSimulator`` resolves to the B200 build
(ref: natrix/__init__.py:1)."""
from .core import *  # noqa: F401,F403
