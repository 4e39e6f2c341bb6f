"""Same import path as the reference module (ref: demo/smooth_particles_area.py)."""
from natrix_b200.smooth_particles_area import SmoothParticlesArea  # noqa: F401

__all__ = ["SmoothParticlesArea"]
