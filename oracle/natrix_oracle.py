"""CPU oracle for the Natrix per-step stable-fluids pipeline (float32 NumPy).

TEST INFRASTRUCTURE ONLY.  Nothing under ``natrix_b200/`` may import this file; it is
used by ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl
reference`` legs of ``bench.py`` as the checker and as the reported CPU baseline.

PARITY UNPINNED: the reference (fbertola/Natrix) ships no tests, golden vectors or
known-answer fixtures for this path (its ``tests/`` holds one empty file), and its bgfx
execution engine (``bgfx-python 2.0.1``, poetry.lock:76-86) is an un-vendored native
dependency that is absent here, so the reference cannot be run to produce vectors either.
This file is therefore a literal restatement of the arithmetic that IS in the reference
tree (the GLSL compute shaders) plus the dispatch order of ``FluidSimulator.update``.  It
is cross-checked bit-for-bit against an independently written C restatement
(``oracle/natrix_oracle.c``) and frozen by ``tests/golden/``.

Every function cites the reference file:line it follows (paths relative to the
reference root).  Interpretation choices:

* linear indices are integers (the shaders compute them in float32, exact only up to
  2**24 cells - Q1);
* pressure is one float per cell (Q2); all buffers start at zero (Q16);
* ``mix(a, b, t) = a*(1-t) + b*t`` (GLSL definition), never fused;
* ``inversesqrt(x) = 1/sqrt(x)`` with IEEE sqrt and division;
* all scalars are narrowed double -> float32 exactly where the reference builds a
  ``c_float`` (fluid_simulator.py:119-130, 315-336);
* expressions are evaluated left to right in float32 with no FMA contraction.
"""
from __future__ import annotations

import numpy as np

F = np.float32
_ONE = F(1.0)
_HALF = F(0.5)
_QUARTER = F(0.25)
_EPS = F(2.4414e-4)


# --------------------------------------------------------------------------- helpers
def _grid(width: int, height: int):
    """gl_GlobalInvocationID.xy as float32 planes of shape (H, W)."""
    xs = np.arange(width, dtype=F)[None, :]
    ys = np.arange(height, dtype=F)[:, None]
    return xs, ys


def solid(obstacles: np.ndarray) -> np.ndarray:
    """``obstacle.x > 0 || obstacle.y > 0`` (shader.AdvectVelocity.comp:30 and every
    other consumer)."""
    return (obstacles[..., 0] > 0) | (obstacles[..., 1] > 0)


def neighbours(a: np.ndarray):
    """common.sh:9-19 GetNeighbours: clamp-to-edge L, R, B(y-1), T(y+1) gathers."""
    left = np.concatenate([a[:, :1], a[:, :-1]], axis=1)
    right = np.concatenate([a[:, 1:], a[:, -1:]], axis=1)
    bottom = np.concatenate([a[:1], a[:-1]], axis=0)
    top = np.concatenate([a[1:], a[-1:]], axis=0)
    return left, right, bottom, top


def mix(a, b, t):
    """GLSL mix: a*(1-t) + b*t, each product rounded separately."""
    return a * (_ONE - t) + b * t


def _clamp(x, lo, hi):
    return np.minimum(np.maximum(x, lo), hi)


def _bilinear_corners(fx, fy, width, height):
    """shader.AdvectVelocity.comp:38-42 - clamped floor/ceil corners and the UNclamped
    delta (Q6)."""
    zero = F(0.0)
    bx, by = F(width - 1), F(height - 1)
    trx = _clamp(np.ceil(fx), zero, bx)
    try_ = _clamp(np.ceil(fy), zero, by)
    blx = _clamp(np.floor(fx), zero, bx)
    bly = _clamp(np.floor(fy), zero, by)
    dx = fx - blx
    dy = fy - bly
    return (trx.astype(np.int64), try_.astype(np.int64), blx.astype(np.int64),
            bly.astype(np.int64), dx, dy)


# ------------------------------------------------------------------- velocity stages
def init_boundaries(vel: np.ndarray) -> None:
    """shader.InitBoundaries.comp:14-34 - zero the four border lines IN PLACE."""
    vel[0, :, :] = 0
    vel[-1, :, :] = 0
    vel[:, 0, :] = 0
    vel[:, -1, :] = 0


def advect_velocity(vel, obstacles, dt, speed, dissipation):
    """shader.AdvectVelocity.comp:27-50."""
    h, w = vel.shape[:2]
    dt, speed, dissipation = F(dt), F(speed), F(dissipation)
    xs, ys = _grid(w, h)
    fx = xs - vel[..., 0] * dt * speed
    fy = ys - vel[..., 1] * dt * speed
    trx, try_, blx, bly, dx, dy = _bilinear_corners(fx, fy, w, h)
    lt = vel[try_, blx]
    rt = vel[try_, trx]
    lb = vel[bly, blx]
    rb = vel[bly, trx]
    dx2, dy2 = dx[..., None], dy[..., None]
    h1 = mix(lt, rt, dx2)
    h2 = mix(lb, rb, dx2)
    out = _clamp(mix(h2, h1, dy2) * dissipation, F(-1.0), F(1.0))
    out[solid(obstacles)] = 0
    return out.astype(F, copy=False)


def calc_vorticity(vel):
    """shader.CalcVorticity.comp:20-26."""
    vl, vr, vb, vt = neighbours(vel)
    return _HALF * ((vr[..., 1] - vl[..., 1]) - (vt[..., 0] - vb[..., 0]))


def apply_vorticity(vel, vort, dt, scale):
    """shader.ApplyVorticity.comp:26-39."""
    dt, scale = F(dt), F(scale)
    wl, wr, wb, wt = neighbours(vort)
    fx = _HALF * (np.abs(wt) - np.abs(wb))
    fy = _HALF * (np.abs(wr) - np.abs(wl))
    mag = np.maximum(_EPS, fx * fx + fy * fy)
    inv = _ONE / np.sqrt(mag)
    fx = fx * inv
    fy = fy * inv
    k = scale * vort                      # _VorticityScale * vC, then * vec2(1, -1)
    fx = fx * k
    fy = fy * (-k)
    out = np.empty_like(vel)
    out[..., 0] = vel[..., 0] + fx * dt
    out[..., 1] = vel[..., 1] + fy * dt
    return out


def viscosity_alpha_rbeta(viscosity: float):
    """fluid_simulator.py:327-336 - computed in Python double, narrowed to c_float."""
    centre = 1.0 / viscosity
    stencil = 1.0 / (4.0 + centre)
    return F(centre), F(stencil)


def viscosity_sweep(vel, alpha, rbeta):
    """shader.Viscosity.comp:24-31."""
    x1, x2, y1, y2 = neighbours(vel)
    return (x1 + x2 + y1 + y2 + vel * F(alpha)) * F(rbeta)


def divergence(vel, obstacles):
    """shader.Divergence.comp:22-40."""
    vl, vr, vb, vt = neighbours(vel)
    sl, sr, sb, st = neighbours(solid(obstacles))
    zero = F(0.0)
    x1 = np.where(sl, zero, vl[..., 0])
    x2 = np.where(sr, zero, vr[..., 0])
    y1 = np.where(sb, zero, vb[..., 1])
    y2 = np.where(st, zero, vt[..., 1])
    return _HALF * ((x2 - x1) + (y2 - y1))


def poisson_sweep(p, div, obstacles, solid_nb=None):
    """shader.Poisson.comp:24-37 (one Jacobi sweep).  ``solid_nb`` may carry the
    precomputed neighbour-solid masks (they do not change within a step)."""
    pl, pr, pb, pt = neighbours(p)
    sl, sr, sb, st = solid_nb if solid_nb is not None else neighbours(solid(obstacles))
    x1 = np.where(sl, p, pl)
    x2 = np.where(sr, p, pr)
    y1 = np.where(sb, p, pb)
    y2 = np.where(st, p, pt)
    return (x1 + x2 + y1 + y2 - div) * _QUARTER


def subtract_gradient(vel, p, obstacles):
    """shader.SubtractGradient.comp:24-46."""
    pl, pr, pb, pt = neighbours(p)
    sl, sr, sb, st = neighbours(solid(obstacles))
    x1 = np.where(sl, p, pl)
    x2 = np.where(sr, p, pr)
    y1 = np.where(sb, p, pb)
    y2 = np.where(st, p, pt)
    out = np.empty_like(vel)
    out[..., 0] = vel[..., 0] - _HALF * (x2 - x1)
    out[..., 1] = vel[..., 1] - _HALF * (y2 - y1)
    return out


def _splat_len(px, py, width, height):
    """``distance(_Position * _Size, vec2(gl_GlobalInvocationID))``
    (shader.AddVelocity.comp:27-30)."""
    sx = F(px) * F(width)
    sy = F(py) * F(height)
    xs, ys = _grid(width, height)
    ddx = sx - xs
    ddy = sy - ys
    return np.sqrt(ddx * ddx + ddy * ddy)


def add_velocity(vel, position, value, radius):
    """shader.AddVelocity.comp:26-35 - note the clamp applies to EVERY cell (Q7)."""
    h, w = vel.shape[:2]
    r = F(radius)
    ln = _splat_len(position[0], position[1], w, h)
    inside = ln <= r
    fall = (r - ln)
    out = vel.copy()
    for c in (0, 1):
        add = F(value[c]) * fall / r
        out[..., c] = np.where(inside, vel[..., c] + add, vel[..., c])
    return _clamp(out, F(-1.0), F(1.0))


def add_circle_obstacle(obstacles, position, radius, static=False) -> None:
    """shader.AddCircleObstacle.comp:24-36 - both branches write (1, 0) (Q17)."""
    h, w = obstacles.shape[:2]
    inside = _splat_len(position[0], position[1], w, h) <= F(radius)
    obstacles[inside] = (1.0, 0.0)


def _tri_sign(p1x, p1y, p2x, p2y, p3x, p3y):
    """shader.AddTriangleObstacle.comp:19-22."""
    return ((p1x - p3x) * (p2y - p3y)) - ((p2x - p3x) * (p1y - p3y))


def add_triangle_obstacle(obstacles, p1, p2, p3, static=False) -> None:
    """shader.AddTriangleObstacle.comp:24-51."""
    h, w = obstacles.shape[:2]
    xs, ys = _grid(w, h)
    ptx = np.broadcast_to(xs / F(w), (h, w))
    pty = np.broadcast_to(ys / F(h), (h, w))
    a = (F(p1[0]), F(p1[1]))
    b = (F(p2[0]), F(p2[1]))
    c = (F(p3[0]), F(p3[1]))
    b1 = _tri_sign(ptx, pty, a[0], a[1], b[0], b[1]) < 0
    b2 = _tri_sign(ptx, pty, b[0], b[1], c[0], c[1]) < 0
    b3 = _tri_sign(ptx, pty, c[0], c[1], a[0], a[1]) < 0
    inside = (b1 == b2) & (b2 == b3)
    obstacles[inside] = (0.0, 1.0) if static else (1.0, 0.0)


# ------------------------------------------------------------------------ dye stages
def add_particles(dye, position, radius, value):
    """demo/shaders/shader.AddParticle.comp:25-34."""
    h, w = dye.shape
    r = F(radius)
    ln = _splat_len(position[0], position[1], w, h)
    splat = _clamp(dye + F(value) * (r - ln) / r, F(0.0), F(255.0))
    return np.where(ln <= r, splat, dye).astype(F, copy=False)


def advect_particles(dye, vel, obstacles, dt, speed, dissipation):
    """demo/shaders/shader.AdvectParticle.comp:21-70."""
    ph, pw = dye.shape
    vh, vw = vel.shape[:2]
    dt, speed, dissipation = F(dt), F(speed), F(dissipation)
    xs, ys = _grid(pw, ph)
    nx = np.broadcast_to((xs / F(pw)) * F(vw), (ph, pw))
    ny = np.broadcast_to((ys / F(ph)) * F(vh), (ph, pw))
    ox = nx.astype(np.int64)              # uint(fNormalisedPos.x): truncation
    oy = ny.astype(np.int64)
    is_solid = solid(obstacles)[oy, ox]
    # GetVelocity (:21-35)
    trx, try_, blx, bly, dx, dy = _bilinear_corners(nx, ny, vw, vh)
    lt, rt, lb, rb = vel[try_, blx], vel[try_, trx], vel[bly, blx], vel[bly, trx]
    dx2, dy2 = dx[..., None], dy[..., None]
    h1 = mix(lt, rt, dx2)
    h2 = mix(lb, rb, dx2)
    ratio = np.array([F(pw) / F(vw), F(ph) / F(vh)], dtype=F)
    v = mix(h2, h1, dy2) * ratio
    fx = xs - v[..., 0] * dt * speed
    fy = ys - v[..., 1] * dt * speed
    trx, try_, blx, bly, dx, dy = _bilinear_corners(fx, fy, pw, ph)
    g1 = mix(dye[try_, blx], dye[try_, trx], dx)
    g2 = mix(dye[bly, blx], dye[bly, trx], dx)
    out = mix(g2, g1, dy) * dissipation
    out[is_solid] = 0
    return out.astype(F, copy=False)


def dye_to_rgba8(dye):
    """demo/shaders/demo.ComputeShader.comp:9-21 - imageStore of vec4(v, v, v, v) into an rgba8 (unorm)
    image: every channel = round-to-nearest-even of clamp(v, 0, 1) * 255."""
    c = np.rint(_clamp(dye, F(0.0), F(1.0)) * F(255.0)).astype(np.uint8)
    return np.repeat(c[..., None], 4, axis=-1)


# ------------------------------------------------------------------ frame rendering
def _fract(x):
    return x - np.floor(x)


def field_colour_lut():
    """plasma(fbm(c)) for the 256 values c = k / 255 an rgba8 texel can take
    (demo/shaders/demo.FieldFragmentShader.frag:21-33 plasma, :73-92 rand / noise / fbm, :96-99 main).
    Returns (256, 3) float32.  rand() multiplies sin() by 43758.5, so one ulp of sin() moves a colour by
    up to ~2/255: comparisons with another sin() implementation need that tolerance."""
    c = (np.arange(256, dtype=F) / F(255.0)).astype(F)

    def rand(n):
        return _fract(np.sin(n).astype(F) * F(43758.5453123))

    def noise(p):
        fl, fc = np.floor(p), _fract(p)
        return mix(rand(fl), rand(fl + F(1.0)), fc)

    v, a, x = np.zeros_like(c), F(0.5), c.copy()
    for _ in range(5):
        v = v + a * noise(x)
        x = x * F(2.0) + F(100.0)
        a = a * F(0.5)
    coeff = np.array([[0.05873234392399702, 0.02333670892565664, 0.5433401826748754],
                      [2.176514634195958, 0.2383834171260182, 0.7539604599784036],
                      [-2.689460476458034, -7.455851135738909, 3.110799939717086],
                      [6.130348345893603, 42.3461881477227, -28.51885465332158],
                      [-11.10743619062271, -82.66631109428045, 60.13984767418263],
                      [10.02306557647065, 71.41361770095349, -54.07218655560067],
                      [-3.658713842777788, -22.93153465461149, 18.19190778539828]], dtype=F)
    t = v[:, None]
    out = np.broadcast_to(coeff[6], (256, 3)).astype(F)
    for k in range(5, -1, -1):
        out = coeff[k] + t * out
    return out.astype(F)


def _unorm8(c):
    return np.rint(_clamp(c, F(0.0), F(1.0)) * F(255.0)).astype(F)


def _blend8(src_rgb, src_a, dst):
    """BGFX_STATE_BLEND_ALPHA on an rgba8 target: src * a + dst * (1 - a) per channel, rounded to 8 bits.
    dst is (H, W, 4) float32 in [0, 1]."""
    ia = F(1.0) - src_a
    out = np.empty_like(dst)
    for ch in range(3):
        out[..., ch] = _unorm8(src_rgb[..., ch] * src_a + dst[..., ch] * ia) / F(255.0)
    out[..., 3] = _unorm8(src_a * src_a + dst[..., 3] * ia) / F(255.0)
    return out


def _line(px, py, ax, ay, bx, by):
    """demo/shaders/demo.QuiverFragmentShader.frag:20-28."""
    cx, cy = (ax + bx) * F(0.5), (ay + by) * F(0.5)
    ex, ey = bx - ax, by - ay
    ln = np.sqrt(ex * ex + ey * ey)
    with np.errstate(invalid="ignore", divide="ignore"):
        dx, dy = ex / ln, ey / ln
    rx, ry = px - cx, py - cy
    d1 = np.abs(rx * dy + ry * (-dx))
    d2 = np.abs(rx * dx + ry * dy) - F(0.5) * ln
    return np.maximum(d1, d2)


def render_frame(dye, vel=None, quiver_tile=0.0):
    """The frame demo/simulation_demo.py:249-281 draws, as (H, W, 4) uint8, top row first.

    Pass 1: the rgba8 dye texture (dye_to_rgba8) through plasma(fbm(texel)) with alpha = texel
    (demo.FieldFragmentShader.frag:96-99), alpha-blended over the clear colour 0x1a0427ff (:94).
    Pass 2 (quiver_tile > 0): demo.QuiverFragmentShader.frag:14-70, white arrows with alpha = 1 - dist.
    Conventions as in natrix_b200/csrc/render.cu: the quad covers the framebuffer, which has the dye grid's
    size; framebuffer y grows upwards (dye row 0 at the bottom); each pass is rounded to 8 bits."""
    h, w = dye.shape
    c8 = dye_to_rgba8(dye)[..., 0].astype(np.int64)
    lut = field_colour_lut()
    dst = np.empty((h, w, 4), F)
    dst[...] = np.array([26.0, 4.0, 39.0, 255.0], F) / F(255.0)
    fb = _blend8(lut[c8], (c8.astype(F) / F(255.0)).astype(F), dst)           # indexed by framebuffer row (bottom = 0)
    if quiver_tile > 0:
        t = F(quiver_tile)
        vh, vw = vel.shape[:2]
        fx = (np.arange(w, dtype=F) + F(0.5))[None, :] + np.zeros((h, 1), F)
        fy = (np.arange(h, dtype=F) + F(0.5))[:, None] + np.zeros((1, w), F)
        cx, cy = (np.floor(fx / t) + F(0.5)) * t, (np.floor(fy / t) + F(0.5)) * t
        qx = (F(1.0) - cx / F(w)) * F(vw)
        qy = (F(1.0) - cy / F(h)) * F(vh)
        tx, ty, bx, by, dx, dy = _bilinear_corners(qx, qy, vw, vh)
        h1x = mix(vel[ty, bx, 0], vel[ty, tx, 0], dx)
        h2x = mix(vel[by, bx, 0], vel[by, tx, 0], dx)
        vx = F(-1.0) * mix(h2x, h1x, dy) * (F(w) / F(vw))
        vy = F(-1.0) * vel[by, bx, 1] * (F(h) / F(vh))       # mix(a, b, vec2(t, 0)): y is never interpolated
        vx, vy = vx * t * F(0.4), vy * t * F(0.4)
        px, py = fx - cx, fy - cy
        mag = np.sqrt(vx * vx + vy * vy)
        with np.errstate(invalid="ignore", divide="ignore"):
            ux, uy = vx / mag, vy / mag
        m2 = _clamp(mag, F(0.0), t * F(0.5))
        ax, ay = ux * m2, uy * m2
        shaft = _line(px, py, ax, ay, -ax, -ay)
        hd1 = _line(px, py, ax, ay, F(0.4) * ax + F(0.2) * (-ay), F(0.4) * ay + F(0.2) * ax)
        hd2 = _line(px, py, ax, ay, F(0.4) * ax + F(0.2) * ay, F(0.4) * ay + F(0.2) * (-ax))
        dist = np.where(mag > F(0.001), np.minimum(shaft, np.minimum(hd1, hd2)), F(1.0)).astype(F)
        white = np.ones((h, w, 3), F)
        fb = _blend8(white, F(1.0) - _clamp(dist, F(0.0), F(1.0)), fb)
    return np.rint(fb[::-1] * F(255.0)).astype(np.uint8)


# ------------------------------------------------------------------ simulator mirror
# ---------------------------------------------------------------------------------------------------------------
# Pressure solvers that are NOT reference behaviour (product: natrix_b200/csrc/solvers.cu, opt-in
# through NATRIX_OPT_SOLVER).  They solve the same system as shader.Poisson.comp:24-37 iterates on -
# x1 + x2 + y1 + y2 - 4 p = div with the centre substituted for solid / outside neighbours - and are restated here
# operation by operation so that the CUDA kernels can be checked bit for bit.
def rb_sor_sweep(p, rhs, nb, omega):
    """One red-black successive over-relaxation sweep: cells with (x + y) even, then the others, each
    p <- p + omega * (gs - p) with gs the shader's Jacobi update of that cell from the CURRENT field."""
    h, w = p.shape
    yy, xx = np.mgrid[0:h, 0:w]
    red = ((xx + yy) & 1) == 0
    for colour in (red, ~red):
        gs = poisson_sweep(p, rhs, None, nb)
        p = np.where(colour, p + F(omega) * (gs - p), p)
    return p


def mg_restrict(a):
    """full weighting of a cell-centred grid: the mean of the 2 x 2 children"""
    return F(0.25) * (a[0::2, 0::2] + a[1::2, 0::2] + a[0::2, 1::2] + a[1::2, 1::2])


def mg_prolong(a):
    """cell-centred bilinear interpolation (weights 9/16, 3/16, 3/16, 1/16), clamp-to-edge; rows first, then columns"""
    def up(x, axis):
        lo = np.concatenate([np.take(x, [0], axis), np.take(x, range(x.shape[axis] - 1), axis)], axis)
        hi = np.concatenate([np.take(x, range(1, x.shape[axis]), axis), np.take(x, [-1], axis)], axis)
        even, odd = F(0.75) * x + F(0.25) * lo, F(0.75) * x + F(0.25) * hi
        out = np.stack([even, odd], axis=axis + 1)
        shape = list(x.shape)
        shape[axis] *= 2
        return out.reshape(shape)
    return up(up(a, 0), 1)


def mg_levels(solid0, min_size=16):
    """solid maps of the multigrid hierarchy: a coarse cell is solid when all four children are; coarsening stops
    when a side is odd or the next level would be smaller than min_size / 2 on a side"""
    solids = [solid0]
    while min(solids[-1].shape) >= min_size and solids[-1].shape[0] % 2 == 0 and solids[-1].shape[1] % 2 == 0:
        s = solids[-1]
        solids.append(s[0::2, 0::2] & s[1::2, 0::2] & s[0::2, 1::2] & s[1::2, 1::2])
    return solids


def mg_v_cycle(p, rhs, solids, level, nu):
    """One V(nu, nu) cycle with a red-black Gauss-Seidel smoother.  rhs plays the role of `div` at this level."""
    nb = neighbours(solids[level])
    for _ in range(nu):
        p = rb_sor_sweep(p, rhs, nb, 1.0)
    if level + 1 < len(solids):
        # residual of x1 + x2 + y1 + y2 - 4 p = rhs with the same neighbour substitution
        pl, pr, pb, pt = neighbours(p)
        sl, sr, sb, st = nb
        lap = np.where(sl, p, pl) + np.where(sr, p, pr) + np.where(sb, p, pb) + np.where(st, p, pt) - F(4.0) * p
        # solid cells carry no equation (their own update only ever sees themselves): whatever divergence the
        # stencil left in them must not reach the coarse grid, where it would be "corrected" again every cycle
        # (without this line the iteration diverges on deep hierarchies: 1024^2 with one circle, 8 levels)
        r = np.where(solids[level], F(0.0), rhs - lap)
        coarse_rhs = F(4.0) * mg_restrict(r)                     # h -> 2h: the right-hand side scales by 4
        e = mg_v_cycle(np.zeros_like(coarse_rhs), coarse_rhs, solids, level + 1, nu)
        p = p + mg_prolong(e)
    for _ in range(nu):
        p = rb_sor_sweep(p, rhs, nb, 1.0)
    return p


class OracleFluidSimulator:
    """State machine of natrix/core/fluid_simulator.py:15-515 over NumPy arrays.

    Field shapes: velocity (2, H, W, 2) ping-pong, pressure (2, H, W) ping-pong,
    divergence / vorticity (H, W), obstacles (H, W, 2).  Ping-pong indices follow
    fluid_simulator.py:16-20 and the flips at :444-474.
    """

    def __init__(self, width: int, height: int, vertex_layout=None):
        self.width, self.height = int(width), int(height)
        self.speed = 500.0            # fluid_simulator.py:28-35 defaults
        self.iterations = 50
        self.dissipation = 1.0
        self.vorticity = 0.0
        self.viscosity = 0.1
        self.has_borders = True
        self.simulate = True
        h, w = self.height, self.width
        self._vel = [np.zeros((h, w, 2), F), np.zeros((h, w, 2), F)]
        self._p = [np.zeros((h, w), F), np.zeros((h, w), F)]
        self.divergence = np.zeros((h, w), F)
        self.vorticity_field = np.zeros((h, w), F)
        self.obstacles = np.zeros((h, w, 2), F)
        self.VELOCITY_READ, self.VELOCITY_WRITE = 0, 1
        self.PRESSURE_READ, self.PRESSURE_WRITE = 0, 1

    # -- accessors
    @property
    def velocity(self):
        return self._vel[self.VELOCITY_READ]

    @velocity.setter
    def velocity(self, v):
        self._vel[self.VELOCITY_READ] = np.ascontiguousarray(v, dtype=F).reshape(
            self.height, self.width, 2).copy()

    @property
    def pressure(self):
        return self._p[self.PRESSURE_READ]

    def _flip_v(self):
        self.VELOCITY_READ, self.VELOCITY_WRITE = self.VELOCITY_WRITE, self.VELOCITY_READ

    def _flip_p(self):
        self.PRESSURE_READ, self.PRESSURE_WRITE = self.PRESSURE_WRITE, self.PRESSURE_READ

    # -- mutators (fluid_simulator.py:116-172)
    def add_velocity(self, position, velocity, radius):
        if self.simulate:
            self._vel[self.VELOCITY_WRITE] = add_velocity(self.velocity, position, velocity, radius)
            self._flip_v()

    def add_circle_obstacle(self, position, radius, static=False):
        if self.simulate:
            add_circle_obstacle(self.obstacles, position, radius, static)

    def add_triangle_obstacle(self, p1, p2, p3, static=False):
        if self.simulate:
            add_triangle_obstacle(self.obstacles, p1, p2, p3, static)

    # -- the hot path (fluid_simulator.py:174-280)
    def update(self, time_delta: float):
        if not self.simulate:
            return
        dt = F(time_delta)
        if self.has_borders:
            init_boundaries(self._vel[self.VELOCITY_READ])                      # :181-188
        self._vel[self.VELOCITY_WRITE] = advect_velocity(
            self.velocity, self.obstacles, dt, self.speed, self.dissipation)    # :191-198
        self._flip_v()
        self.vorticity_field = calc_vorticity(self.velocity)                    # :201-207
        self._vel[self.VELOCITY_WRITE] = apply_vorticity(
            self.velocity, self.vorticity_field, dt, self.vorticity)            # :210-217
        self._flip_v()
        if self.viscosity > 0.0:                                                # :220-228
            alpha, rbeta = viscosity_alpha_rbeta(self.viscosity)
            self._vel[self.VELOCITY_WRITE] = viscosity_sweep(self.velocity, alpha, rbeta)
            self._flip_v()
        self.divergence = divergence(self.velocity, self.obstacles)             # :231-233
        if not getattr(self, "warm_start", False):      # warm_start: opt-in extension of the product, not the reference
            self._p[self.PRESSURE_READ] = np.zeros_like(self._p[0])             # :236-248
        nb = neighbours(solid(self.obstacles))
        solver = getattr(self, "solver", "jacobi")      # "sor" / "multigrid": opt-in extensions of the product
        if solver == "sor":
            for _ in range(int(self.iterations)):
                self._p[self.PRESSURE_READ] = rb_sor_sweep(self.pressure, self.divergence, nb, getattr(self, "sor_omega", 1.9))
        elif solver == "multigrid":
            solids = mg_levels(solid(self.obstacles))
            for _ in range(int(self.iterations)):
                self._p[self.PRESSURE_READ] = mg_v_cycle(self.pressure, self.divergence, solids, 0, int(getattr(self, "mg_smooth", 2)))
        else:
            for _ in range(int(self.iterations)):                               # :251-255
                self._p[self.PRESSURE_WRITE] = poisson_sweep(
                    self.pressure, self.divergence, self.obstacles, nb)
                self._flip_p()
        self._vel[self.VELOCITY_WRITE] = subtract_gradient(
            self.velocity, self.pressure, self.obstacles)                       # :258-265
        self._flip_v()
        self.obstacles = np.zeros_like(self.obstacles)                          # :268-280


class OracleSmoothParticlesArea:
    """demo/smooth_particles_area.py:15-211 over NumPy arrays (the "dye" field)."""

    def __init__(self, width, height, fluid_simulation: OracleFluidSimulator, vertex_layout=None):
        self.width, self.height = int(width), int(height)
        self.fluid_simulation = fluid_simulation
        self.speed = 500.0
        self.dissipation = 1.0
        self.simulate = True
        self.particles = np.zeros((self.height, self.width), F)

    def add_particles(self, position, radius, strength):
        if self.simulate:
            self.particles = add_particles(self.particles, position, radius, strength)

    def update(self, time_delta):
        if self.simulate:
            sim = self.fluid_simulation
            self.particles = advect_particles(
                self.particles, sim.velocity, sim.obstacles, time_delta, self.speed,
                self.dissipation)
