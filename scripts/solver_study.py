"""CPU study for SURVEY 8(f)-4 (not product code, no GPU): how many stencil passes do red-black SOR and a
geometric V-cycle need to leave the residual that N Jacobi sweeps leave, on the reference's own discretisation
(shader.Poisson.comp: Neumann by neighbour substitution at obstacles and grid edges)?

    python scripts/solver_study.py [size]

Residual = RMS over fluid cells of the algebraic residual x1 + x2 + y1 + y2 - 4 p - div of that system, for
one solve from p = 0 (the divergence left after the projection is no convergence measure here: the shaders
take divergence and gradient over 2 cells but solve the compact 5-point system, so it has a floor that no
solver lowers).  Work is counted in full-grid stencil passes
(a red-black sweep = 1 pass, a V-cycle level l costs 4^-l per smoothing pass)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np  # noqa: E402

from natrix_b200 import workloads as W  # noqa: E402
from oracle import natrix_oracle as O  # noqa: E402

F = np.float32


def setup(n):
    rng = np.random.default_rng(3)
    vel = W.smooth_velocity(n, n) + (0.2 * rng.uniform(-1, 1, (n, n, 2))).astype(F)
    obs = np.zeros((n, n, 2), F)
    for (px, py, r) in [(0.25, 0.25, 0.04 * n), (0.75, 0.3, 0.06 * n), (0.5, 0.6, 0.08 * n), (0.3, 0.8, 0.025 * n)]:
        O.add_circle_obstacle(obs, (px, py), r)
    return vel, obs


def residual(vel, p, obs):
    div = O.divergence(vel, obs)
    nb = O.neighbours(O.solid(obs))
    r = O.poisson_sweep(p, div, None, nb) * F(4.0) - F(4.0) * p        # x1 + x2 + y1 + y2 - div - 4 p
    fluid = ~O.solid(obs)
    return float(np.sqrt(np.mean(r[fluid].astype(np.float64) ** 2)))


def jacobi(p, div, nb, sweeps):
    for _ in range(sweeps):
        p = O.poisson_sweep(p, div, None, nb)
    return p


def rb_sor(p, div, nb, sweeps, omega):
    for _ in range(sweeps):
        p = O.rb_sor_sweep(p, div, nb, omega)
    return p


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    vel, obs = setup(n)
    div = O.divergence(vel, obs)
    nb = O.neighbours(O.solid(obs))
    zero = np.zeros((n, n), F)
    print(f"{n}x{n}, 4 circles; residual at p = 0: {residual(vel, zero, obs):.3e}")
    print(f"{'method':34s} {'passes':>8s} {'rms residual':>15s}")
    for s in (25, 50, 100, 200, 400, 800):
        print(f"{'Jacobi (reference)':34s} {s:8d} {residual(vel, jacobi(zero, div, nb, s), obs):15.3e}")
    for omega in (1.0, 1.7, 1.9):
        for s in (25, 50, 100):
            print(f"{f'red-black SOR, omega = {omega}':34s} {s:8d} {residual(vel, rb_sor(zero, div, nb, s, omega), obs):15.3e}")
    solids = O.mg_levels(O.solid(obs))
    per_cycle = sum((2 * 2 + 0.5) * 4.0 ** -lv for lv in range(len(solids)))          # fine-grid passes of one V(2,2)
    for cycles in (1, 2, 4, 8):
        p = zero
        for _ in range(cycles):
            p = O.mg_v_cycle(p, div, solids, 0, 2)
        print(f"{f'V(2,2) x {cycles}, red-black smoother':34s} {cycles * per_cycle:8.1f} {residual(vel, p, obs):15.3e}")


if __name__ == "__main__":
    main()
