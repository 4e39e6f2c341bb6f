"""CPU tests of host-side logic that needs no device: property validation (same messages as
the reference), workload generators, slab partitioning."""
import numpy as np
import pytest

from natrix_b200 import workloads as W
from natrix_b200.core.fluid_simulator import FluidSimulator
from natrix_b200.smooth_particles_area import SmoothParticlesArea


class _NoDevice(FluidSimulator):
    """Exercises the property setters without creating a device handle."""

    def __init__(self):          # noqa: D401 - deliberately skips the C-ABI call
        self._speed, self._iterations, self._dissipation = 500.0, 50, 1.0
        self._vorticity, self._viscosity = 0.0, 0.1
        self._h, self._dyes = None, []


def test_defaults_match_reference():
    s = _NoDevice()
    assert (s.speed, s.iterations, s.dissipation, s.vorticity, s.viscosity) == (500.0, 50, 1.0, 0.0, 0.1)
    assert FluidSimulator.has_borders is True and FluidSimulator.simulate is True
    assert SmoothParticlesArea.simulate is True


@pytest.mark.parametrize("attr,bad,msg", [
    ("speed", 0, "'Speed' should be greater than zero"),
    ("iterations", 0, "'Iterations' should be grater than zero"),
    ("dissipation", -1.0, "'Dissipation' should be grater than zero"),
    ("vorticity", -0.1, "'Vorticity' should be grater or equal than zero"),
    ("viscosity", -0.1, "'Viscosity' should be greater or equal than zero"),
])
def test_property_validation_messages(attr, bad, msg):
    s = _NoDevice()
    with pytest.raises(ValueError) as e:
        setattr(s, attr, bad)
    assert str(e.value) == msg


def test_zero_vorticity_and_viscosity_are_legal():
    s = _NoDevice()
    s.vorticity = 0
    s.viscosity = 0.0
    assert s.viscosity == 0.0


def test_workloads_are_deterministic_and_match_baseline_configs():
    w3 = W.cfg3_workload()
    assert (w3.width, w3.height, w3.iterations, w3.splats_per_step) == (4096, 4096, 100, 8)
    assert W.orbit_positions(w3, 5) == W.orbit_positions(w3, 5)
    assert len(W.orbit_positions(w3, 5)) == 8
    w1 = W.demo_workload()
    assert (w1.width, w1.height, w1.iterations, w1.viscosity, w1.dye_size) == (640, 360, 50, 0.5, (1280, 720))
    assert W.cfg2_workload().algorithmic_bytes_per_cell_step() == 132 + 20 * 50
    w5 = W.cfg5_workload(8)
    assert (w5.width, w5.height, w5.iterations, len(w5.circles)) == (32768, 32768, 200, 64 * 8)   # 64 per GPU band


def test_smooth_velocity_slab_equals_window_of_full_field():
    full = W.smooth_velocity(128, 96)
    part = W.smooth_velocity(128, 96, row0=32, rows=40)
    assert np.array_equal(full[32:72], part)
    assert np.max(np.abs(full)) <= 0.5


def test_png_writer_round_trip(tmp_path):
    """headless demo driver: the RGBA8 PNG it writes decodes back to the same pixels."""
    import struct
    import zlib

    from natrix_b200.headless_demo import write_png

    img = np.random.default_rng(1).integers(0, 256, (37, 53, 4), dtype=np.uint8)
    path = tmp_path / "t.png"
    write_png(path, img)
    b = path.read_bytes()
    assert b[:8] == b"\x89PNG\r\n\x1a\n"
    i, idat, tags = 8, b"", []
    while i < len(b):
        n, = struct.unpack(">I", b[i:i + 4])
        tag, data = b[i + 4:i + 8], b[i + 8:i + 8 + n]
        assert struct.unpack(">I", b[i + 8 + n:i + 12 + n])[0] == zlib.crc32(tag + data) & 0xFFFFFFFF
        tags.append(tag)
        if tag == b"IHDR":
            assert struct.unpack(">IIBBBBB", data) == (53, 37, 8, 6, 0, 0, 0)
        if tag == b"IDAT":
            idat += data
        i += 12 + n
    assert tags[0] == b"IHDR" and tags[-1] == b"IEND"
    raw = np.frombuffer(zlib.decompress(idat), np.uint8).reshape(37, 1 + 53 * 4)
    assert (raw[:, 0] == 0).all() and np.array_equal(raw[:, 1:].reshape(37, 53, 4), img)


def test_checkpoint_round_trip_continues_bit_identically(tmp_path):
    """natrix_b200.checkpoint on the oracle classes: a run restored from a file written after 3 frames
    (with an obstacle stamped and not yet consumed) continues exactly like the run that never stopped."""
    from natrix_b200 import checkpoint
    from oracle.natrix_oracle import OracleFluidSimulator, OracleSmoothParticlesArea

    def make():
        w = W.demo_workload()
        w.width, w.height, w.dye_size, w.iterations, w.init = 48, 36, (96, 72), 7, "random"
        return (w, *W.build(w, OracleFluidSimulator, OracleSmoothParticlesArea))

    w, a, ad = make()
    for k in range(3):
        W.run_step(w, a, ad, k)
    a.add_circle_obstacle((0.3, 0.6), 5.0)                      # pending obstacle: part of the state
    checkpoint.save(tmp_path / "c.npz", a, [ad])
    _, b, bd = make()
    b.iterations, b.vorticity = 3, 9.0                          # overwritten by the checkpoint
    checkpoint.load(tmp_path / "c.npz", b, [bd])
    assert (b.iterations, b.vorticity, b.viscosity) == (a.iterations, a.vorticity, a.viscosity)
    for k in range(3, 6):
        W.run_step(w, a, ad, k)
        W.run_step(w, b, bd, k)
    for name in ("velocity", "pressure", "divergence"):
        assert np.array_equal(getattr(a, name), getattr(b, name)), name
    assert np.array_equal(ad.particles, bd.particles)
    _, c, cd = make()
    with pytest.raises(ValueError):
        checkpoint.load(tmp_path / "c.npz", c, [])              # dye count mismatch


def _plan(width, depth, r0, r1, boxes, max_tiles):
    import ctypes as C

    from natrix_b200 import _lib as L

    flat = (C.c_int * max(1, 4 * len(boxes)))(*[int(v) for b in boxes for v in b])
    cap = 4 * max_tiles + 64
    out = (C.c_int * (4 * cap))()
    n = L.check(L.lib().natrix_debug_plan_tiles(width, depth, r0, r1, flat, len(boxes), max_tiles, out, cap))
    assert n <= cap
    return np.array(out[:4 * n], dtype=np.int64).reshape(n, 4)


@pytest.mark.parametrize("seed", range(6))
def test_jacobi_tile_plan_partitions_every_strip(seed):
    """The obstacle-aware tile planner of the temporally blocked Jacobi kernel (host code, no device needed):
    whatever the obstacle hints, every strip's tiles cover the row range exactly once."""
    rng = np.random.default_rng(seed)
    width = int(rng.choice([256, 1024, 4096, 32768]))
    depth = int(rng.integers(1, 9))
    r0 = int(rng.integers(-24, 1))
    r1 = r0 + int(rng.choice([9, 64, 360, 4096 + 48]))
    boxes = []
    for _ in range(int(rng.integers(0, 40))):
        if rng.uniform() < 0.6:                                  # a circle: (cx, -1 - r, cy, 0)
            boxes.append((rng.integers(0, width), -1 - int(rng.integers(1, 600)), rng.integers(r0 - 50, r1 + 50), 0))
        else:
            x0, y0 = int(rng.integers(0, width)), int(rng.integers(r0 - 20, r1))
            boxes.append((x0, x0 + int(rng.integers(1, 900)), y0, y0 + int(rng.integers(1, 700))))
    max_tiles = int(rng.choice([148 * 12, 148 * 16, 200]))
    tiles = _plan(width, depth, r0, r1, boxes, max_tiles)
    pitch = 128 - 2 * (4 if depth <= 4 else 8)
    nstrips = -(-width // pitch)
    assert len(tiles) >= nstrips and set(tiles[:, 0]) == set(range(nstrips))
    for st in range(nstrips):
        t = tiles[tiles[:, 0] == st]
        t = t[np.argsort(t[:, 1])]
        assert t[0, 1] == r0 and t[-1, 2] == r1, f"strip {st} does not span the rows"
        assert (t[1:, 1] == t[:-1, 2]).all() and (t[:, 2] > t[:, 1]).all(), f"strip {st} has a gap or an overlap"
    assert np.array_equal(tiles, _plan(width, depth, r0, r1, boxes, max_tiles))      # deterministic
    if len(boxes) == 0 and (r1 - r0) * nstrips >= 8 * max_tiles:
        assert len(tiles) <= max_tiles + nstrips                                       # about one tile per warp
    if (r1 - r0) * nstrips >= 16 * max_tiles and len(boxes) == 0:
        assert len(tiles) >= max_tiles                                                 # no resident warp is left without a tile


def test_jacobi_tile_plan_is_shorter_where_obstacles_are():
    free = _plan(4096, 8, 0, 4096, [], 148 * 12)
    circ = _plan(4096, 8, 0, 4096, [(2048, -1 - 700, 2048, 0)], 148 * 12)
    h = lambda t, st: (t[t[:, 0] == st][:, 2] - t[t[:, 0] == st][:, 1])
    mid, edge = 2048 // 112, 0
    assert h(circ, mid).min() < h(free, mid).min()             # the strip through the circle is cut finer ...
    assert h(circ, edge).mean() >= h(free, edge).mean()        # ... and the free strips get the longer tiles


def test_bench_reference_arm_prints_the_contract_line():
    """bench.py --impl reference (the CPU port timed on the host cores) needs no GPU: one JSON line with the
    keys the measurement contract names."""
    import json
    import subprocess
    import sys

    from conftest import ROOT

    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--workload", "demo",
                          "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "Mcell-steps/s" == d["unit"] and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["config"]["workload"] == "demo-640x360" and d["dtype"] == "f32" and d["data"] == "synthetic"
    cb = d["cpu_baseline"]
    from oracle import natrix_ref
    assert cb["kind"] == ("reference" if natrix_ref.available() else "port")
    assert cb["cores"] >= 1 and cb["value"] == d["value"] and "demo-640x360" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and d["vs_baseline"] is None


def test_bench_roofline_reads_the_committed_ncu_captures():
    """bench.py turns profiles/kernel_traffic.json (ncu DRAM bytes per cell of the captured kernels) into the physical
    roofline of its JSON line: the file must parse, name the kernels the bench asks for, and hold sane per-cell bytes."""
    import importlib.util
    import json

    from conftest import ROOT

    spec = importlib.util.spec_from_file_location("bench_mod", ROOT / "bench.py")
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    data = json.loads((ROOT / "profiles" / "kernel_traffic.json").read_text())
    assert data["captures"] and all(c["dram_bytes_per_cell_per_launch"] > 0 for c in data["captures"])
    big = bench.ncu_capture_for("k_jacobi_tb", 32768 * 4096)
    mid = bench.ncu_capture_for("k_jacobi_tb", 4096 * 4096)
    assert big["cells"] == 32768 * 4096 and mid["cells"] == 4096 * 4096
    # a launch of the temporally blocked kernel must move at least 13 B per cell (p, div4, mask in; p out) minus what
    # the 126 MB L2 keeps at 4096^2, and not much more than that
    assert 13.0 <= big["dram_bytes_per_cell_per_launch"] < 18.0
    assert 9.0 <= mid["dram_bytes_per_cell_per_launch"] < 18.0
    assert 0.0 < big["sm_active_over_elapsed"] <= 1.0 and 0.0 < big["issue_active_pct"] <= 100.0
    assert bench.ncu_capture_for("k_preproject", 32768 * 4096)["dram_bytes_per_cell_per_launch"] >= 26.0
    assert bench.ncu_capture_for("k_no_such_kernel", 1) is None
