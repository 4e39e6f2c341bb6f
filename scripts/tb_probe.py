"""Smallest possible exercise of the temporally blocked Jacobi kernel (for compute-sanitizer):
python scripts/tb_probe.py [width height [depth]] - compares pipeline 1 with pipeline 0."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np  # noqa: E402

from natrix_b200 import _lib as L, workloads as W  # noqa: E402
from natrix_b200.core.fluid_simulator import FluidSimulator  # noqa: E402

w, h = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (512, 256)
depth = int(sys.argv[3]) if len(sys.argv) > 3 else 8
out = []
for pipeline in (0, 1):
    s = FluidSimulator(w, h)
    s.set_option(L.OPT_PIPELINE, pipeline)
    s.set_option(L.OPT_JACOBI_DEPTH, depth)
    s.vorticity, s.viscosity, s.iterations = 1.0, 0.2, 19
    s.upload("velocity", W.random_velocity(w, h, 3))
    s.add_circle_obstacle((0.5, 0.5), 20.0)
    s.update(W.DT)
    out.append(W.fields_of(s))
    s.destroy()
for k in out[0]:
    print(k, "bit-identical:", np.array_equal(out[0][k], out[1][k]), "max|diff|",
          float(np.abs(out[0][k] - out[1][k]).max()))
