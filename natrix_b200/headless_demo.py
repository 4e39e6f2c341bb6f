"""Headless replay of the reference demo: demo/simulation_demo.py:220-237 frame order
(circle obstacle -> fluid.update -> dye.update -> scripted "mouse" impulses) on a 1280 x 720 framebuffer,
frames rendered on the device by natrix_render_frame and written as PNG files.

    python -m natrix_b200.headless_demo --frames 300 --every 30 --out frames/ [--quiver 32]

There is no window, no bgfx and no ImGui: the scripted drag of workloads.demo_workload() stands in for the
mouse.  Needs a CUDA device (no CPU fallback)."""
from __future__ import annotations

import argparse
import struct
import time
import zlib
from pathlib import Path

import numpy as np


def write_png(path, rgba: np.ndarray) -> None:
    """Minimal RGBA8 PNG writer (zlib + CRC from the standard library)."""
    h, w, c = rgba.shape
    assert c == 4 and rgba.dtype == np.uint8
    raw = np.concatenate([np.zeros((h, 1), np.uint8), rgba.reshape(h, w * 4)], axis=1).tobytes()   # filter 0 per row

    def chunk(tag: bytes, data: bytes) -> bytes:
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)

    png = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 6, 0, 0, 0))
    png += chunk(b"IDAT", zlib.compress(raw, 6)) + chunk(b"IEND", b"")
    Path(path).write_bytes(png)


def run(frames: int, every: int, out_dir, quiver: float = 0.0, width: int = 1280, height: int = 720, device: int = 0):
    from natrix_b200 import workloads as W
    from natrix_b200.core.fluid_simulator import FluidSimulator
    from natrix_b200.smooth_particles_area import SmoothParticlesArea

    w = W.demo_workload()
    w.width, w.height, w.dye_size = width // 2, height // 2, (width, height)       # simulation_demo.py:100-111
    sim, dye = W.build(w, FluidSimulator, SmoothParticlesArea, device=device)
    out_dir = Path(out_dir)
    out_dir.mkdir(parents=True, exist_ok=True)
    written, t0 = [], time.perf_counter()
    for k in range(frames):
        W.run_step(w, sim, dye, k)
        if every > 0 and (k + 1) % every == 0:
            path = out_dir / f"frame_{k + 1:05d}.png"
            write_png(path, dye.render_frame(quiver))
            written.append(path)
    sim.synchronize()
    return written, frames / (time.perf_counter() - t0)


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--frames", type=int, default=300)
    ap.add_argument("--every", type=int, default=30, help="write every n-th frame (0: none)")
    ap.add_argument("--out", default="frames")
    ap.add_argument("--quiver", type=float, default=0.0, help="arrow tile size in pixels (the demo offers 8, 16, 32, 64)")
    ap.add_argument("--device", type=int, default=0)
    args = ap.parse_args(argv)
    written, fps = run(args.frames, args.every, args.out, args.quiver, device=args.device)
    print(f"{args.frames} frames at {fps:.0f} frames/s, {len(written)} PNG files in {args.out}/")
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
