"""GPU parity tests (pytest -m gpu): the CUDA path, called through the C ABI (ctypes), against
the oracle on the same seeded inputs.  Bar: BASELINE.json's rel. tol 1e-5 per field per step
(conftest.REL_TOL); in fact the kernels are written to be bit-identical and most tests assert
that too.  Full-size cases use size-independent properties (fused pipeline == reference-order
pipeline bit for bit, temporal blocking depth independence)."""
import importlib.util

import numpy as np
import pytest

from conftest import ROOT, assert_fields_close, field_report
from natrix_b200 import _lib as L
from natrix_b200 import workloads as W
from natrix_b200.core.fluid_simulator import FluidSimulator
from natrix_b200.smooth_particles_area import SmoothParticlesArea
from conftest import parity_oracle

# The oracle of every test below is oracle/_ref - the reference's own shader text compiled as C++ and driven in
# the reference's dispatch order (oracle/natrix_ref.py) - whenever its prebuilt library is present (it is built
# where /root/reference exists and travels to the GPU box); the restated C / NumPy oracles, proven bit-identical
# to it by tests/test_ref_oracle.py, are the fallback.  The names keep saying which restatement a test used to
# take; `ORACLE_KIND` says what is actually in use (also printed in the pytest header).
(OracleFluidSimulator, OracleSmoothParticlesArea, COracleFluidSimulator, COracleSmoothParticlesArea,
 ORACLE_KIND) = parity_oracle()

pytestmark = pytest.mark.gpu


def test_the_parity_oracle_is_the_reference_shader_text():
    if not ORACLE_KIND.startswith("reference"):
        pytest.skip(f"oracle/_ref not present here; parity checked against: {ORACLE_KIND}")
    from oracle import natrix_ref as R
    assert b"source=" in R.lib("literal").nref_build_info()


def _mg():
    spec = importlib.util.spec_from_file_location("make_golden", ROOT / "tests" / "golden" / "make_golden.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


TB, SMEM = 1, 2          # NATRIX_OPT_JACOBI_KERNEL: the TMA register-streaming kernel / the shared-memory kernel
# (pipeline, Jacobi kernel): the reference-order pipeline, and the fused pipeline with each of its two Jacobi kernels
# forced - small grids would otherwise only ever reach the shared-memory one
PIPELINES = [(0, None), (1, TB), (1, SMEM)]


def _sim_cls(pipeline, depth=None, kernel=None):
    def make(w, h, layout=None, **kw):
        s = FluidSimulator(w, h, layout, **kw)
        s.set_option(L.OPT_PIPELINE, pipeline)
        if kernel == TB and (w % 16 or w < 128):
            pass                                   # TMA cannot address this width: the library's own choice
        elif kernel is not None:
            s.set_option(L.OPT_JACOBI_KERNEL, kernel)
            assert s.get_option(L.OPT_JACOBI_KERNEL) == (kernel if pipeline else 0)
        if depth is not None:
            s.set_option(L.OPT_SMEM_DEPTH if kernel == SMEM else L.OPT_JACOBI_DEPTH, depth)
        return s
    return make


@pytest.mark.parametrize("pipeline,kernel", PIPELINES)
def test_small_case_matches_golden_fixture(pipeline, kernel, golden_dir):
    want = np.load(golden_dir / "small_case.npz")
    got = _mg().small_case(_sim_cls(pipeline, kernel=kernel), SmoothParticlesArea)
    assert_fields_close(got, {k: want[k] for k in want.files}, f"pipeline {pipeline} kernel {kernel}: ", exact=True)


@pytest.mark.parametrize("pipeline,kernel", PIPELINES)
def test_config1_demo_ten_steps_vs_numpy_oracle(pipeline, kernel):
    """config 1: demo grid 640x360 + dye 1280x720, random initial velocity, 10 frames."""
    w = W.demo_workload()
    w.init = "random"
    gsim, gdye = W.build(w, _sim_cls(pipeline, kernel=kernel), SmoothParticlesArea)
    osim, odye = W.build(w, OracleFluidSimulator, OracleSmoothParticlesArea)
    for k in range(10):
        W.run_step(w, gsim, gdye, k)
        W.run_step(w, osim, odye, k)
        assert_fields_close(W.fields_of(gsim, gdye), W.fields_of(osim, odye), f"step {k}: ", exact=True)
    gsim.destroy()


def test_config1_zero_state_first_frame_dt0():
    """the demo's first frame has dt = 0 and an all-zero state."""
    w = W.demo_workload()
    gsim, gdye = W.build(w, FluidSimulator, SmoothParticlesArea)
    osim, odye = W.build(w, OracleFluidSimulator, OracleSmoothParticlesArea)
    for k, dt in enumerate((0.0, W.DT, W.DT)):
        W.run_step(w, gsim, gdye, k, dt)
        W.run_step(w, osim, odye, k, dt)
        assert_fields_close(W.fields_of(gsim, gdye), W.fields_of(osim, odye), f"frame {k}: ", exact=True)


@pytest.mark.parametrize("pipeline,kernel", PIPELINES)
def test_config2_1024_twenty_steps_vs_c_oracle(pipeline, kernel):
    """config 2: 1024^2, 50 iterations, vorticity confinement, 4 circular obstacles."""
    w = W.cfg2_workload()
    gsim, _ = W.build(w, _sim_cls(pipeline, kernel=kernel), None)
    osim, _ = W.build(w, COracleFluidSimulator, None)
    for k in range(20):
        W.run_step(w, gsim, None, k)
        W.run_step(w, osim, None, k)
        if k in (0, 1, 9, 19):
            assert_fields_close(W.fields_of(gsim), W.fields_of(osim), f"step {k}: ", exact=True)


@pytest.mark.parametrize("kernel,depth", [(TB, d) for d in range(1, 9)] + [(SMEM, d) for d in (1, 2, 5, 8, 11, 13, 16)])
def test_temporal_blocking_is_depth_independent(kernel, depth):
    """T sweeps per launch must equal T launches of one sweep, bit for bit, for every T and for both temporally
    blocked kernels, including the remainder block (iterations = 19 is not a multiple of any T > 1)."""
    w, h = 640, 296
    rng = np.random.default_rng(depth)
    v0 = (0.6 * rng.uniform(-1, 1, (h, w, 2))).astype(np.float32)
    out = []
    for pipeline, d in ((0, None), (1, depth)):
        s = _sim_cls(pipeline, d, kernel if pipeline else None)(w, h)
        s.vorticity, s.viscosity, s.iterations = 0.7, 0.2, 19
        s.upload("velocity", v0)
        for _ in range(2):
            s.add_circle_obstacle((0.3, 0.4), 30.0)
            s.add_circle_obstacle((0.0, 0.0), 25.0)          # touches the corner
            s.add_circle_obstacle((0.999, 0.5), 12.0)        # touches the right edge
            s.add_triangle_obstacle((0.6, 0.2), (0.9, 0.3), (0.7, 0.8), static=True)
            s.update(W.DT)
        out.append(W.fields_of(s))
        s.destroy()
    assert_fields_close(out[1], out[0], f"depth {depth}: ", exact=True)


@pytest.mark.parametrize("kernel", [TB, SMEM])
def test_subnormal_divergence_takes_the_two_step_form(kernel):
    """The temporally blocked kernels read 0.25 * divergence and end a sweep in one fma (common.cuh NB_RAW), which is
    bit-identical to the shader's (sum - b) * 0.25 except where 0.25 * b is inexact: a non-zero |b| < 2^-124.  Those
    cells must carry the NB_RAW bit and take the two-step form.  Velocities around 1e-38 in the left half of the grid
    make such divergences; the right half holds ordinary values, so both forms run side by side in one row."""
    w, h = 256, 96
    rng = np.random.default_rng(5)
    v0 = (0.5 * rng.uniform(-1, 1, (h, w, 2))).astype(np.float32)
    v0[:, : w // 2] = (rng.integers(-40, 40, (h, w // 2, 2)).astype(np.float64) * 2.0 ** -149 * 7).astype(np.float32)
    v0[10:30, 20:60] = 0.0
    outs = []
    for pipeline, k in ((0, None), (1, kernel)):
        s = _sim_cls(pipeline, kernel=k)(w, h)
        s.vorticity, s.viscosity, s.iterations, s.has_borders = 0.0, 0.0, 21, False
        s.upload("velocity", v0)
        s.add_circle_obstacle((0.3, 0.5), 12.0)
        s.update(0.0)                                  # dt = 0: the back-trace returns the cell itself, the tiny values survive
        outs.append(W.fields_of(s))
        if pipeline:
            m, b, b4 = s.download("nbmask"), s.download("divergence"), s.download("div4")
            raw = (m & 16) != 0
            assert raw.any() and not raw[:, w // 2 + 2:].any()
            assert np.array_equal(b4[raw], b[raw]) and np.array_equal(b4[~raw], (b * np.float32(0.25))[~raw])
            assert np.array_equal((b4 * np.float32(4.0))[~raw], b[~raw])
            tiny = (b != 0) & (np.abs(b) < 2.0 ** -124)
            assert tiny.any() and not (raw & ~tiny).any()
        s.destroy()
    assert_fields_close(outs[1], outs[0], f"kernel {kernel} vs reference-order pipeline: ", exact=True)
    o = OracleFluidSimulator(w, h)
    o.vorticity, o.viscosity, o.iterations, o.has_borders = 0.0, 0.0, 21, False
    o.velocity = v0
    o.add_circle_obstacle((0.3, 0.5), 12.0)
    o.update(0.0)
    assert_fields_close(outs[1], W.fields_of(o), f"kernel {kernel} vs oracle: ", exact=True)
    assert float(np.abs(outs[1]["pressure"]).max()) > 0.0


@pytest.mark.parametrize("w,h", [(97, 61), (256, 40), (272, 33), (1000, 24), (16, 300), (1, 9), (130, 1)])
def test_ragged_and_degenerate_sizes_vs_oracle(w, h):
    rng = np.random.default_rng(w + 1000 * h)
    v0 = rng.uniform(-1.5, 1.5, (h, w, 2)).astype(np.float32)   # |v| > 1: back-traces leave the grid
    for pipeline, kernel in PIPELINES:
        if kernel == TB and (w % 16 or w < 128):
            continue                                # the same run as (1, SMEM): TMA cannot address this width
        g = _sim_cls(pipeline, kernel=kernel)(w, h)
        o = OracleFluidSimulator(w, h)
        for s in (g, o):
            s.vorticity, s.viscosity, s.iterations, s.speed = 2.0, 0.3, 11, 300.0
        g.upload("velocity", v0)
        o.velocity = v0
        for k in range(3):
            for s in (g, o):
                s.add_circle_obstacle((0.4, 0.6), min(w, h) / 5.0)
                s.update(W.DT)
                s.add_velocity((0.5, 0.5), (0.7, -0.4), 4.0)
            assert_fields_close(W.fields_of(g), W.fields_of(o), f"{w}x{h} p{pipeline} step {k}: ", exact=True)
        g.destroy()


@pytest.mark.parametrize("borders,visc,vort", [(False, 0.0, 0.0), (True, 0.0, 3.0), (False, 2.0, 0.5)])
def test_parameter_corners_vs_oracle(borders, visc, vort):
    w, h = 320, 200
    v0 = W.random_velocity(w, h, seed=11)
    g, o = FluidSimulator(w, h), OracleFluidSimulator(w, h)
    for s in (g, o):
        s.has_borders, s.viscosity, s.vorticity, s.iterations, s.dissipation = borders, visc, vort, 23, 0.97
    g.upload("velocity", v0)
    o.velocity = v0
    for k in range(3):
        for s in (g, o):
            s.add_triangle_obstacle((0.2, 0.2), (0.5, 0.25), (0.3, 0.7))
            s.update(W.DT)
        assert_fields_close(W.fields_of(g), W.fields_of(o), f"step {k}: ", exact=True)


def test_impulses_and_obstacles_vs_oracle():
    w, h = 300, 180
    g, o = FluidSimulator(w, h), OracleFluidSimulator(w, h)
    v0 = (3.0 * W.random_velocity(w, h, seed=2)).astype(np.float32)        # |v| up to 1.5
    g.upload("velocity", v0)
    o.velocity = v0
    for s in (g, o):
        s.add_circle_obstacle((0.25, 0.5), 20.0)
        s.add_circle_obstacle((1.2, 0.5), 80.0, static=True)               # centre outside the grid
        s.add_circle_obstacle((0.5, 0.5), -1.0)                            # negative radius: nothing
        s.add_triangle_obstacle((0.5, 0.1), (0.9, 0.2), (0.6, 0.9), static=True)
        s.add_triangle_obstacle((0.1, 0.6), (0.3, 0.7), (0.2, 0.95), static=False)
        s.add_velocity((0.5, 0.5), (0.4, -0.9), 25.0)
        s.add_velocity((0.1, 0.9), (-2.0, 2.0), 60.0)
    assert np.array_equal(g.download("obstacles"), o.obstacles)
    got = g.download("velocity")
    assert np.array_equal(got, o.velocity)
    assert np.all(np.abs(got) <= 1.0)                                       # Q7: every cell clamped
    g.update(W.DT)
    o.update(W.DT)
    assert not g.download("obstacles").any()                                # Q8: cleared by update
    assert_fields_close(W.fields_of(g), W.fields_of(o), exact=True)


def test_collinear_triangle_is_cleared_like_any_other_obstacle():
    """an oblique collinear triangle selects the whole line through its points (all three edge functions vanish there),
    far outside its bounding box; the fused pipeline clears only the rows it believes were stamped, so it must believe
    every row was (the reference clears every cell each step)"""
    w, h = 256, 192
    g, o = FluidSimulator(w, h), OracleFluidSimulator(w, h)
    v0 = W.random_velocity(w, h, seed=5)
    g.upload("velocity", v0)
    o.velocity = v0
    for k in range(2):
        for s in (g, o):
            s.add_triangle_obstacle((0.25, 0.25), (0.5, 0.5), (0.75, 0.75))
            s.add_circle_obstacle((0.5, 0.2), 1e10)                 # a radius far beyond 2^31 cells
        assert np.array_equal(g.download("obstacles"), o.obstacles)
        for s in (g, o):
            s.update(W.DT)
        assert not g.download("obstacles").any()
        assert_fields_close(W.fields_of(g), W.fields_of(o), f"step {k}: ", exact=True)


def test_simulate_false_gates_every_mutator():
    g = FluidSimulator(64, 64)
    v0 = W.random_velocity(64, 64, 1)
    g.upload("velocity", v0)
    g.simulate = False
    g.add_velocity((0.5, 0.5), (1, 1), 10.0)
    g.add_circle_obstacle((0.5, 0.5), 10.0)
    g.update(W.DT)
    assert np.array_equal(g.download("velocity"), v0)
    assert not g.download("obstacles").any()


def test_dye_cross_resolution_vs_oracle():
    w, h = 200, 120
    g, o = FluidSimulator(w, h), OracleFluidSimulator(w, h)
    gd, od = SmoothParticlesArea(333, 250, g), OracleSmoothParticlesArea(333, 250, o)
    v0 = W.random_velocity(w, h, 4)
    g.upload("velocity", v0)
    o.velocity = v0
    for d in (gd, od):
        d.dissipation, d.speed = 0.95, 400.0
        d.add_particles((0.5, 0.5), 60.0, 3.0)
        d.add_particles((0.2, 0.7), 30.0, 300.0)       # saturates the 255 clamp
    for k in range(3):
        for s, d in ((g, gd), (o, od)):
            s.add_circle_obstacle((0.6, 0.4), 15.0)
            d.update(W.DT)                              # sees the obstacle (added before update)
            s.update(W.DT)
            d.update(W.DT)                              # obstacle map is empty again
        err, scale, ndiff = field_report(gd.download(), od.particles)
        assert ndiff == 0, (k, err, scale)


def test_zero_copy_views_and_stats():
    import torch

    g = FluidSimulator(256, 128)
    v0 = W.random_velocity(256, 128, 9)
    g.upload("velocity", v0)
    buf = g.get_velocity_buffer()
    t = torch.as_tensor(buf, device="cuda:0")
    g.synchronize()
    assert t.shape == (128, 256, 2) and t.data_ptr() == buf.data_ptr
    assert np.array_equal(t.cpu().numpy(), v0)
    assert np.array_equal(buf.numpy(), v0)
    s, q, lo, hi = g.stats("velocity")
    assert np.isclose(s, v0.astype(np.float64).sum(), rtol=1e-12, atol=1e-9)
    assert np.isclose(q, (v0.astype(np.float64) ** 2).sum(), rtol=1e-12)
    assert (lo, hi) == (float(v0.min()), float(v0.max()))


def test_errors_are_loud():
    g = FluidSimulator(64, 64)
    g.destroy()
    with pytest.raises(L.NatrixError):
        g.update(W.DT)
    with pytest.raises(L.NatrixError):
        FluidSimulator(0, 10)
    g2 = FluidSimulator(32, 32)
    import ctypes as C
    buf = np.zeros(5, np.uint8)
    rc = g2._lib.natrix_copy_in(g2._handle(), 0, buf.ctypes.data_as(C.c_void_p), 5)
    assert rc == -1 and b"size" in g2._lib.natrix_last_error()
    with pytest.raises(ValueError):
        g2.iterations = 0


def test_config3_4096_fused_equals_reference_order_pipeline():
    """config 3 at BASELINE.json's full size.  The CPU oracle is too slow for 4096^2 x 100
    sweeps x several steps, so the full-size check is the size-independent property that the
    fused / temporally blocked pipeline is bit-identical to the one-kernel-per-shader pipeline
    (which the smaller cases pin to the oracle), plus one oracle-checked step with 12 sweeps."""
    w = W.cfg3_workload()
    a, ad = W.build(w, _sim_cls(0), SmoothParticlesArea)
    b, bd = W.build(w, _sim_cls(1), SmoothParticlesArea)
    for k in range(2):
        W.run_step(w, a, ad, k)
        W.run_step(w, b, bd, k)
    fa, fb = W.fields_of(a, ad), W.fields_of(b, bd)
    assert_fields_close(fb, fa, "4096^2 fused vs reference-order: ", exact=True)
    assert float(np.abs(fa["velocity"]).max()) > 0.1 and float(fa["dye"].max()) > 0.0
    a.destroy()
    # one step against the C oracle, 12 sweeps
    w.iterations = 12
    o, od = W.build(w, COracleFluidSimulator, COracleSmoothParticlesArea)
    b2, bd2 = W.build(w, _sim_cls(1), SmoothParticlesArea)
    W.run_step(w, o, od, 0)
    W.run_step(w, b2, bd2, 0)
    assert_fields_close(W.fields_of(b2, bd2), W.fields_of(o, od), "4096^2 vs C oracle: ", exact=True)


def test_dye_vectorised_path_cross_resolution_vs_oracle():
    """dye width % 4 == 0 takes the 4-cells-per-thread kernel; resolution differs from the velocity grid"""
    w, h = 200, 120
    g, o = FluidSimulator(w, h), OracleFluidSimulator(w, h)
    gd, od = SmoothParticlesArea(336, 252, g), OracleSmoothParticlesArea(336, 252, o)
    v0 = W.random_velocity(w, h, 6)
    g.upload("velocity", v0)
    o.velocity = v0
    for d in (gd, od):
        d.dissipation = 0.97
        d.add_particles((0.5, 0.5), 70.0, 5.0)
        d.add_particles((0.9, 0.1), 40.0, 1.0)
    for k in range(3):
        for s, d in ((g, gd), (o, od)):
            s.add_circle_obstacle((0.6, 0.4), 15.0)
            d.update(W.DT)
            s.update(W.DT)
            d.update(W.DT)
            d.add_particles((0.3, 0.3 + 0.1 * k), 20.0, 0.5)
        err, scale, ndiff = field_report(gd.download(), od.particles)
        assert ndiff == 0 and scale > 0, (k, err, scale)


def test_all_cell_clamp_with_localised_overshoot_vs_oracle():
    """only a few cells exceed |v| = 1 (one band of rows): add_velocity elsewhere must still clamp them (Q7)"""
    w, h = 512, 320
    v0 = (0.4 * W.random_velocity(w, h, 8)).astype(np.float32)
    v0[200:203, 40:60] = 1.7
    v0[10, 500] = -3.0
    g, o = FluidSimulator(w, h), OracleFluidSimulator(w, h)
    g.upload("velocity", v0)
    o.velocity = v0
    for s in (g, o):
        s.iterations = 9
        s.add_velocity((0.8, 0.9), (0.2, 0.1), 10.0)
    assert np.array_equal(g.download("velocity"), o.velocity)
    for k in range(2):
        for s in (g, o):
            s.update(W.DT)
            s.add_velocity((0.1, 0.1), (0.5, -0.5), 12.0)
        assert_fields_close(W.fields_of(g), W.fields_of(o), f"step {k}: ", exact=True)


def test_results_do_not_depend_on_scheduling_hints(monkeypatch):
    """chunk heights (obstacle-aware planning, NATRIX_TB_CHUNK / KAPPA) and the launch shape only change
    how the rows are cut, never a bit of the result"""
    w, h = 768, 400
    v0 = W.random_velocity(w, h, 12)
    outs = []
    for env in ({}, {"NATRIX_TB_CHUNK": "20"}, {"NATRIX_TB_KAPPA": "3.5", "NATRIX_TB_SHAPE": "0"}):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        s = _sim_cls(1, kernel=TB)(w, h)
        s.vorticity, s.viscosity, s.iterations = 1.0, 0.0, 21
        s.upload("velocity", v0)
        for _ in range(2):
            s.add_circle_obstacle((0.5, 0.3), 35.0)
            s.add_triangle_obstacle((0.1, 0.6), (0.4, 0.65), (0.2, 0.9))
            s.update(W.DT)
        outs.append(W.fields_of(s))
        s.destroy()
        for k in env:
            monkeypatch.delenv(k)
    for other in outs[1:]:
        assert_fields_close(other, outs[0], exact=True)


def test_dye_rgba8_export_vs_oracle():
    """the RGBA8 image of demo.ComputeShader.comp, on the device"""
    from oracle.natrix_oracle import dye_to_rgba8

    g = FluidSimulator(128, 96)
    gd = SmoothParticlesArea(200, 150, g)
    rng = np.random.default_rng(3)
    dye = rng.uniform(-0.2, 1.4, (150, 200)).astype(np.float32)
    dye[0, :5] = [0.0, 1.0, 0.5, 0.5 + 1.0 / 510.0, 2.0]
    gd.upload(dye)
    img = gd.export_rgba8()
    assert img.shape == (150, 200, 4) and img.dtype == np.uint8
    assert np.array_equal(img, dye_to_rgba8(dye))


FRAME_TOL = 2      # 8-bit levels: rand() = fract(sin(n) * 43758.5) amplifies one ulp of sin() to ~2/255 of colour


@pytest.mark.parametrize("tile", [0.0, 8.0, 32.0])
def test_render_frame_vs_oracle(tile):
    """the demo's frame (field colour map over the clear colour, optional quiver overlay)
    against the NumPy transcription of the two fragment shaders, after a few demo steps."""
    from oracle.natrix_oracle import render_frame

    w = W.demo_workload()
    sim, dye = W.build(w, FluidSimulator, SmoothParticlesArea)
    for k in range(6):
        W.run_step(w, sim, dye, k)
    dye.add_particles((0.3, 0.6), 120.0, 0.8)                   # pending splat: the frame must include it
    img = dye.render_frame(tile)
    ref = render_frame(dye.download(), sim.download("velocity"), tile)
    assert img.shape == (w.dye_size[1], w.dye_size[0], 4) and img.dtype == np.uint8
    diff = np.abs(img.astype(np.int32) - ref.astype(np.int32))
    assert float((diff > FRAME_TOL).mean()) < 1e-4, f"{int((diff > FRAME_TOL).sum())} channel values off by more than {FRAME_TOL}"
    assert int(diff.max()) <= 16                                # an arrow edge may flip a pixel's coverage by one rounding
    if tile == 0.0:
        bg = dye.download()[::-1] <= 0.0                       # no dye: exactly the clear colour 0x1a0427ff
        assert bg.any() and (img[bg] == np.array([26, 4, 39, 255], np.uint8)).all()
    assert len(np.unique(img.reshape(-1, 4), axis=0)) > 50     # not a constant image


def test_config5_slab_size_fused_equals_reference_order_pipeline():
    """config 5 at its per-GPU size (32768 x 4096, 64 circles): tile planner, select / solid / free bodies
    and the fused pre-projection against the one-kernel-per-shader pipeline, bit for bit (24 sweeps)."""
    w = W.cfg5_workload(1)
    w.iterations = 24
    a, _ = W.build(w, _sim_cls(0), None)
    b, _ = W.build(w, _sim_cls(1), None)
    for k in range(2):
        W.run_step(w, a, None, k)
        W.run_step(w, b, None, k)
    for name in ("velocity", "pressure", "divergence", "vorticity"):
        fa, fb = a.download(name), b.download(name)
        assert np.array_equal(fa, fb), f"{name}: {int(np.count_nonzero(fa != fb))} cells differ"
        assert float(np.abs(fa).max()) > 0.0
        del fa, fb


def test_moving_obstacles_every_step_a_new_tile_plan_vs_oracle():
    """Obstacles that move every step give the Jacobi kernel a new tile plan every step: 40 steps on a 512 x 256 grid
    walk through more plans than the cache has slots (cut + asynchronous upload + slot reuse on the step path).
    Every field of the last step and a few on the way, bit for bit against the oracle."""
    import dataclasses
    w = W.Workload("moving-512x256", 512, 256, 17, 1.0, 0.1, circles=[(0.2, 0.3, 18.0), (0.6, 0.7, 30.0), (0.9, 0.5, 12.0)],
                   splats_per_step=1, splat_radius=20.0, init="random", drift=5.0)
    g, _ = W.build(w, _sim_cls(1, kernel=TB), None)
    o, _ = W.build(w, OracleFluidSimulator, None)
    for k in range(40):
        W.run_step(w, g, None, k)
        W.run_step(w, o, None, k)
        if k in (0, 7, 39):
            assert_fields_close(W.fields_of(g), W.fields_of(o), exact=True)
    hits, misses = g.plan_cache_stats()
    assert misses >= 40 and hits >= 40, (hits, misses)          # a new plan per step, reused by the step's later launches
    static = dataclasses.replace(w, drift=0.0)
    g2, _ = W.build(static, _sim_cls(1, kernel=TB), None)
    for k in range(5):
        W.run_step(static, g2, None, k)
    assert g2.plan_cache_stats()[1] <= 4                        # the same circles every step: the plans are found again


def test_headless_demo_writes_frames(tmp_path):
    """the headless replay of the demo loop produces PNG frames that are not blank."""
    from natrix_b200.headless_demo import run

    written, fps = run(12, 6, tmp_path, quiver=32.0, width=640, height=360)
    assert [p.name for p in written] == ["frame_00006.png", "frame_00012.png"] and fps > 0
    assert all(p.stat().st_size > 2000 for p in written)


@pytest.mark.parametrize("pipeline,kernel", PIPELINES)
def test_warm_start_option_vs_oracle(pipeline, kernel):
    """NATRIX_OPT_WARM_START (not reference behaviour): the pressure of the previous step is the
    initial guess; still bit-identical to the oracle run with the pressure clear skipped, and different from the
    reference-order run."""
    w = W.cfg2_workload()
    w.width, w.height, w.iterations = 384, 256, 21
    g, _ = W.build(w, _sim_cls(pipeline, kernel=kernel), None)
    o, _ = W.build(w, OracleFluidSimulator, None)
    cold, _ = W.build(w, _sim_cls(pipeline, kernel=kernel), None)
    g.warm_start = True
    o.warm_start = True
    assert g.warm_start and not cold.warm_start
    for k in range(4):
        for s in (g, o, cold):
            W.run_step(w, s, None, k)
        assert_fields_close(W.fields_of(g), W.fields_of(o), f"step {k}: ", exact=True)
    assert not np.array_equal(g.download("pressure"), cold.download("pressure"))


@pytest.mark.parametrize("solver,iters,size", [("sor", 30, (256, 160)), ("multigrid", 3, (256, 160)), ("multigrid", 2, (130, 66)),
                                               ("sor", 7, (97, 61)), ("multigrid", 2, (512, 512))])
def test_solver_extensions_vs_oracle(solver, iters, size):
    """SURVEY 8(f)-4 (NOT reference behaviour): red-black SOR and multigrid V-cycles behind NATRIX_OPT_SOLVER solve the
    system the reference's Jacobi loop iterates on; every field bit-identical to the NumPy restatement
    (oracle rb_sor_sweep / mg_v_cycle), on grids with 5, 2 (odd coarse side), 1 and 6 levels."""
    from oracle.natrix_oracle import OracleFluidSimulator as NumpyOracle     # the reference's shaders have no such solver

    w, h = size
    v0 = W.random_velocity(w, h, seed=21)
    g, o = FluidSimulator(w, h), NumpyOracle(w, h)
    g.solver = solver
    o.solver = solver
    assert g.solver == solver
    for s in (g, o):
        s.vorticity, s.viscosity, s.iterations = 1.5, 0.0, iters
    if solver == "sor":
        g.sor_omega = o.sor_omega = 1.7
    else:
        g.mg_smooth = o.mg_smooth = 2
    g.upload("velocity", v0)
    o.velocity = v0
    for k in range(3):
        for s in (g, o):
            s.add_circle_obstacle((0.35, 0.55), min(w, h) / 6.0)
            s.add_triangle_obstacle((0.6, 0.2), (0.9, 0.3), (0.7, 0.8))
            s.update(W.DT)
            s.add_velocity((0.5, 0.5), (0.6, -0.3), 9.0)
        assert_fields_close(W.fields_of(g), W.fields_of(o), f"{solver} {w}x{h} step {k}: ", exact=True)
    g.destroy()


def test_solver_extensions_leave_less_residual_than_jacobi_for_the_time():
    """what the extensions are for: the RMS residual of the pressure system after one solve, against the solve's
    device time - 3 V(2,2) cycles beat 200 Jacobi sweeps on a 1024^2 grid with obstacles"""
    from oracle import natrix_oracle as O

    w = W.cfg2_workload()
    out = {}
    for name, solver, iters in (("jacobi", "jacobi", 200), ("multigrid", "multigrid", 3)):
        s, _ = W.build(w, FluidSimulator, None)
        s.solver, s.iterations = solver, iters
        s.set_option(L.OPT_TIMING, 1)
        for k in range(2):
            for (px, py, r) in w.circles:
                s.add_circle_obstacle((px, py), r)
            obstacles = s.download("obstacles")
            s.update(W.DT)
        p, div = s.download("pressure"), s.download("divergence")
        nb = O.neighbours(O.solid(obstacles))
        r = O.poisson_sweep(p, div, None, nb) * np.float32(4.0) - np.float32(4.0) * p
        fluid = ~O.solid(obstacles)
        out[name] = (float(np.sqrt(np.mean(r[fluid].astype(np.float64) ** 2))), s.timings()["jacobi"])
        s.destroy()
    assert out["multigrid"][0] < out["jacobi"][0], out
    print("residual, ms:", out)


def test_checkpoint_restores_a_run_bit_identically(tmp_path):
    """natrix_b200.checkpoint through the C ABI: velocity, pressure, pending obstacles, parameters and the dye."""
    from natrix_b200 import checkpoint

    w = W.demo_workload()
    w.init = "random"
    a, ad = W.build(w, FluidSimulator, SmoothParticlesArea)
    for k in range(4):
        W.run_step(w, a, ad, k)
    a.add_circle_obstacle((0.3, 0.6), 25.0)
    a.add_triangle_obstacle((0.6, 0.2), (0.8, 0.3), (0.7, 0.6), True)
    checkpoint.save(tmp_path / "c.npz", a, [ad])
    b, bd = W.build(w, FluidSimulator, SmoothParticlesArea)
    b.iterations = 3
    checkpoint.load(tmp_path / "c.npz", b, [bd])
    assert b.iterations == a.iterations
    for k in range(4, 7):
        W.run_step(w, a, ad, k)
        W.run_step(w, b, bd, k)
    assert_fields_close(W.fields_of(a, ad), W.fields_of(b, bd), "after restore: ", exact=True)


def test_pure_c_host_matches_the_python_mirror():
    """examples/c_host.c drives the demo loop through the C ABI with no Python in the process; the field
    statistics it prints equal those of the same loop through the ctypes mirror."""
    import ctypes
    import subprocess

    exe = ROOT / "examples" / "_build" / "c_host"
    if not exe.exists():
        pytest.skip("examples/_build/c_host is not built (python -c 'import __graft_entry__ as g; g.build()')")
    frames = 12
    out = subprocess.run([str(exe), str(frames)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    got = {ln.split()[0]: [float(v) for v in ln.split()[1:]] for ln in out.stdout.splitlines()
           if ln.split()[0] in ("velocity", "pressure", "dye")}

    libm = ctypes.CDLL("libm.so.6")
    for fn in (libm.cosf, libm.sinf):
        fn.restype, fn.argtypes = ctypes.c_float, [ctypes.c_float]
    f32 = np.float32
    sim = FluidSimulator(640, 360, None)
    sim.vorticity, sim.viscosity, sim.iterations = 1.0, 0.5, 50
    dye = SmoothParticlesArea(1280, 720, sim, None)
    dye.dissipation = 0.98
    dt = float(f32(1.0) / f32(60.0))

    def pos(k):
        a = float(f32(0.1) * f32(k))
        return f32(0.5) + f32(0.3) * f32(libm.cosf(a)), f32(0.5) + f32(0.3) * f32(libm.sinf(a))

    for k in range(frames):
        sim.add_circle_obstacle((0.5, 0.5), 40.0)
        sim.update(dt)
        dye.update(dt)
        (x1, y1), (x0, y0) = pos(k), pos(k - 1)
        sim.add_velocity((float(x1), float(y1)), (float(f32(10.0) * (x1 - x0)), float(f32(10.0) * (y1 - y0))), 32.0)
        dye.add_particles((float(x1), float(y1)), 250.0, 0.04)
    want = {"velocity": sim.stats("velocity"), "pressure": sim.stats("pressure"), "dye": dye.stats()}
    for name in want:
        assert got[name] == list(want[name]), f"{name}: C host {got[name]} != mirror {list(want[name])}"


@pytest.mark.parametrize("pipeline,kernel", PIPELINES)
@pytest.mark.parametrize("seed", range(20))
def test_random_scenarios_vs_c_oracle(seed, pipeline, kernel):
    """Differential test on seeded random scripts of public-API calls (workloads.random_scenario): ragged and
    1-cell-wide grids, every parameter corner, obstacles partly outside the grid, zero radii, dt = 0, dye grids of
    unrelated size - every field of every frame bit-identical to the C oracle."""
    scn = W.random_scenario(seed)
    frames = {}
    W.play_scenario(scn, COracleFluidSimulator, COracleSmoothParticlesArea,
                    lambda k, s, d: frames.__setitem__(k, {n: a.copy() for n, a in W.fields_of(s, d).items()}))

    def check(k, s, d):
        assert_fields_close(W.fields_of(s, d), frames[k], f"seed {seed} {scn['size']} frame {k}: ", exact=True)

    W.play_scenario(scn, _sim_cls(pipeline, kernel=kernel), SmoothParticlesArea, check)
