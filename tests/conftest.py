import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_sessionstart(session):
    """The shared libraries are build artefacts (git-ignored): on a fresh checkout build them once (nvcc
    cross-compiles for sm_100a without a GPU), exactly as __graft_entry__.build() does."""
    missing = [p for p in (ROOT / "natrix_b200" / "libnatrix_b200.so", ROOT / "oracle" / "_build" / "libnatrix_oracle.so",
                           ROOT / "examples" / "_build" / "c_host", ROOT / "examples" / "_build" / "c_host_multi") if not p.exists()]
    if missing:
        import __graft_entry__

        __graft_entry__.build()


# Tolerance of BASELINE.json's north_star: rel. tol 1e-5 per field per step, stated as
# max|a-b| <= 1e-5 * max|b|.  The kernels are written to be bit-identical to
# the oracle (no FMA, IEEE div/sqrt, same operand order); tests report how many elements differ
# at all, and most of them assert exact equality on top of the tolerance.
REL_TOL = 1e-5


def field_report(got: np.ndarray, want: np.ndarray):
    got = np.asarray(got)
    want = np.asarray(want)
    assert got.shape == want.shape, (got.shape, want.shape)
    scale = float(np.max(np.abs(want))) if want.size else 0.0
    err = float(np.max(np.abs(got.astype(np.float64) - want.astype(np.float64)))) if want.size else 0.0
    ndiff = int(np.count_nonzero(got != want))
    return err, scale, ndiff


def assert_fields_close(got: dict, want: dict, what: str = "", exact: bool = False):
    for name, w in want.items():
        err, scale, ndiff = field_report(got[name], w)
        assert np.all(np.isfinite(got[name])), f"{what}{name}: non-finite values"
        assert err <= REL_TOL * max(scale, 1e-30) or err == 0.0, (
            f"{what}{name}: max|err| {err:.3e} > {REL_TOL} * max|ref| {scale:.3e} ({ndiff} elements differ)")
        if exact:
            assert ndiff == 0, f"{what}{name}: {ndiff} elements are not bit-identical (max err {err:.3e})"


def parity_oracle():
    """(sim, dye, sim_mt, dye_mt, kind): the oracle classes the GPU parity tests compare against.  oracle/_ref (the
    reference's own shader text, compiled) when its library is present; else the restated oracles."""
    from oracle import natrix_ref as R

    if R.available():
        return (R.RefFluidSimulator, R.RefSmoothParticlesArea, R.RefFluidSimulator, R.RefSmoothParticlesArea,
                "reference shader text compiled as C++ (oracle/_ref/libnatrix_ref.so)")
    from oracle.c_oracle import COracleFluidSimulator, COracleSmoothParticlesArea
    from oracle.natrix_oracle import OracleFluidSimulator, OracleSmoothParticlesArea

    return (OracleFluidSimulator, OracleSmoothParticlesArea, COracleFluidSimulator, COracleSmoothParticlesArea,
            "restated oracles (oracle/natrix_oracle.py, oracle/natrix_oracle.c) - oracle/_ref missing")


def pytest_report_header(config):
    return f"natrix parity oracle: {parity_oracle()[4]}"


@pytest.fixture(scope="session")
def golden_dir():
    return ROOT / "tests" / "golden"
