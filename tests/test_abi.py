"""CPU tests of the drop-in boundary: the shared library loads, exports every symbol that
include/natrix_b200.h declares, fails loudly without a GPU, and the Python mirror keeps the
reference's names / validation (ref: natrix/core/fluid_simulator.py:58-111)."""
import ctypes
import re

import pytest

from conftest import ROOT
from natrix_b200 import _lib as L


def _declared_symbols():
    text = (ROOT / "include" / "natrix_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(natrix_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_table_agree():
    assert _declared_symbols() == sorted(L.SIGNATURES)


def test_library_exports_every_declared_symbol():
    assert L.LIB_PATH.exists(), "build libnatrix_b200.so first (__graft_entry__.build())"
    lib = ctypes.CDLL(str(L.LIB_PATH))
    for name in _declared_symbols():
        assert hasattr(lib, name), f"{name} is declared in include/natrix_b200.h but not exported"


def test_version_and_error_strings_are_callable_without_gpu():
    lib = L.lib()
    assert b"natrix_b200" in lib.natrix_version()
    assert isinstance(lib.natrix_last_error(), bytes)


def test_null_handles_are_rejected_not_dereferenced():
    lib = L.lib()
    assert lib.natrix_step(None, 0.1) == -1
    assert lib.natrix_set_params(None, 1.0, 1, 1.0, 0.0, 0.0, 1) == -1
    assert lib.natrix_dye_step(None, 0.1, 1.0, 1.0) == -1
    assert b"null" in lib.natrix_last_error()
    assert lib.natrix_destroy(None) == 0


def test_no_cpu_fallback_without_a_device():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from natrix_b200.core.fluid_simulator import FluidSimulator

    with pytest.raises(L.NatrixError):
        FluidSimulator(64, 64, None)


def test_reference_import_paths_resolve_to_the_b200_classes():
    from demo.smooth_particles_area import SmoothParticlesArea
    from natrix.core.fluid_simulator import FluidSimulator
    from natrix_b200.core.fluid_simulator import FluidSimulator as F2
    from natrix_b200.smooth_particles_area import SmoothParticlesArea as S2

    assert FluidSimulator is F2 and SmoothParticlesArea is S2
    for name in ("add_velocity", "add_circle_obstacle", "add_triangle_obstacle", "update",
                 "get_velocity_buffer", "destroy", "width", "height", "speed", "iterations",
                 "dissipation", "vorticity", "viscosity", "has_borders", "simulate"):
        assert hasattr(FluidSimulator, name), name
    for name in ("add_particles", "update", "destroy", "speed", "dissipation", "simulate"):
        assert hasattr(SmoothParticlesArea, name), name


def test_product_never_imports_the_oracle():
    for path in (ROOT / "natrix_b200").rglob("*.py"):
        text = path.read_text()
        assert "oracle" not in text.replace("the oracle", "").replace("NumPy oracle", "").replace(
            "C oracle", "").replace("oracles", "").replace("float32 oracle", ""), path
    for path in (ROOT / "natrix_b200" / "csrc").glob("*"):
        if path.suffix in (".cu", ".cuh", ".h"):
            assert "#include \"../../oracle" not in path.read_text()


def test_no_kernel_uses_the_predicate_output_of_vimnmx_relu():
    """Found the hard way: when the result of __vimin_s32_relu(v, m) is later compared with m, ptxas 12.9
    folds the compare into VIMNMX.RELU's predicate output - and that predicate came out true for every cell on B200 (a
    reference-order advect kernel wrote all zeros).  The shipped kernels only use the value form (predicates PT, PT);
    this scan keeps it that way.  Also: sm_100a is the only architecture in the library."""
    import re
    import shutil
    import subprocess

    import pytest

    from natrix_b200 import _lib as L

    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump is not on PATH")
    sass = subprocess.run(["cuobjdump", "-sass", str(L.LIB_PATH)], capture_output=True, text=True, timeout=600).stdout
    assert set(re.findall(r"arch = (sm_\w+)", sass)) == {"sm_100a"}
    forms = set(re.findall(r"VIMNMX\.RELU\s+R\d+, (\w+), (\w+)", sass))
    assert forms, "the fused pre-projection kernel clamps its gather corners with VIMNMX.RELU"
    assert forms == {("PT", "PT")}, forms
    assert "UTMALDG" in sass and "FFMA2" in sass          # TMA bulk tensor loads, 2-wide fp32: the Blackwell paths are in
