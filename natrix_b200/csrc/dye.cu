// dye.cu - the "smooth particles area" (dye) kernels.
// ref: demo/shaders/shader.AddParticle.comp:25-34, demo/shaders/shader.AdvectParticle.comp:21-70,
//      demo/smooth_particles_area.py:70-104.
#include "kernels.h"

namespace natrix {
namespace {

constexpr int BX = 64, BY = 4, BT = BX * BY;   // 2-D blocks: back-traced gathers of nearby rows hit L1

// Geometry: dg describes the dye rows this handle holds (w = dye width, hg = global dye height, rows
// [y0, y0 + hl) plus `halo` rows either side), vg the simulator's velocity rows.  Arrays are indexed by
// local row; the arithmetic only ever sees GLOBAL cell coordinates, so a slab reproduces the full grid.
// A gather that leaves the held rows raises *err (slabs only: the full grid cannot trip it).
template <bool SLAB = true>
__device__ __forceinline__ int held_row(const Geom& g, int gy, int* __restrict__ err) {
    if (!SLAB) return gy;                        // full grid: y0 = 0 and every clamped row is held
    const int lo = g.y0 - g.halo, hi = g.y0 + g.hl + g.halo - 1;
    if (gy < lo || gy > hi) {
        *err = 1;
        gy = clampi(gy, lo, hi);
    }
    return gy - g.y0;
}

// K splats applied in sequence per cell: identical arithmetic to K AddParticle dispatches
// (each dispatch is a pure per-cell map, the ping-pong flip carries no cross-cell dependency).
__global__ void __launch_bounds__(BT)
k_dye_add(const float* __restrict__ din, float* __restrict__ dout, const Geom dg, int r0, int r1,
          const __grid_constant__ SplatDBatch b) {
    const int x = blockIdx.x * BX + threadIdx.x;
    const int y = r0 + blockIdx.y * BY + threadIdx.y;
    if (x >= dg.w || y >= r1) return;
    const ptrdiff_t pos = lin(dg, x, y);
    float v = din[pos];
    const float fxp = (float)x, fyp = (float)(y + dg.y0);
    for (int i = 0; i < b.n; ++i) {
        const SplatD s = b.s[i];
        const float ex = s.sx - fxp, ey = s.sy - fyp;
        const float len = sqrtf(ex * ex + ey * ey);
        if (len <= s.r) v = clampf(v + s.value * (s.r - len) / s.r, 0.0f, 255.0f);
    }
    dout[pos] = v;
}

__global__ void __launch_bounds__(BT)
k_dye_advect(const float* __restrict__ din, float* __restrict__ dout, const Geom dg,
             const float2* __restrict__ vel, const uint8_t* __restrict__ obs, const Geom vg,
             float dt, float speed, float diss, int* __restrict__ err) {
    const int x = blockIdx.x * BX + threadIdx.x;
    const int y = blockIdx.y * BY + threadIdx.y;
    if (x >= dg.w || y >= dg.hl) return;
    const int pw = dg.w, ph = dg.hg, vw = vg.w, vh = vg.hg;
    const int gy = y + dg.y0;
    const ptrdiff_t pos = lin(dg, x, y);
    // fNormalisedPos (:46) and the obstacle lookup at its truncation (:47-50)
    const float nx = ((float)x / (float)pw) * (float)vw;
    const float ny = ((float)gy / (float)ph) * (float)vh;
    if (obs[lin(vg, (int)(unsigned)nx, held_row(vg, (int)(unsigned)ny, err))] != OBS_FREE) { dout[pos] = 0.0f; return; }
    // GetVelocity (:21-35): bilinear sample of the velocity grid, scaled to dye cells
    const Corners c = corners(nx, ny, vw, vh);
    const int vty = held_row(vg, c.ty, err), vby = held_row(vg, c.by, err);
    const float2 lt = vel[lin(vg, c.bx, vty)], rt = vel[lin(vg, c.tx, vty)];
    const float2 lb = vel[lin(vg, c.bx, vby)], rb = vel[lin(vg, c.tx, vby)];
    const float rx = (float)pw / (float)vw, ry = (float)ph / (float)vh;
    const float vx = mixf(mixf(lb.x, rb.x, c.dx), mixf(lt.x, rt.x, c.dx), c.dy) * rx;
    const float vy = mixf(mixf(lb.y, rb.y, c.dx), mixf(lt.y, rt.y, c.dx), c.dy) * ry;
    // back-trace in dye cells and gather the dye with the same clamp rule (:57-69)
    const float fx = (float)x - vx * dt * speed;
    const float fy = (float)gy - vy * dt * speed;
    const Corners q = corners(fx, fy, pw, ph);
    const int dty = held_row(dg, q.ty, err), dby = held_row(dg, q.by, err);
    const float g1 = mixf(din[lin(dg, q.bx, dty)], din[lin(dg, q.tx, dty)], q.dx);
    const float g2 = mixf(din[lin(dg, q.bx, dby)], din[lin(dg, q.tx, dby)], q.dx);
    dout[pos] = mixf(g2, g1, q.dy) * diss;
}

// Same arithmetic, 4 dye cells of one row per thread.  The row-only half of the velocity sample
// (normalised y, its clamped floor / ceil rows and weight, the obstacle row) is computed once per
// thread, and the normalised coordinates come from tables filled by k_dye_tables with the exact
// expression of the shader, so no per-cell IEEE division is left.
constexpr int D4X = 32, D4Y = 8;
__global__ void k_dye_tables(float* __restrict__ nx, float* __restrict__ ny, const Geom dg, int vw, int vh) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < dg.w) nx[i] = ((float)i / (float)dg.w) * (float)vw;        // fNormalisedPos.x (:46)
    if (i < dg.hl) ny[i] = ((float)(i + dg.y0) / (float)dg.hg) * (float)vh;
}

// A thread's 4 cells are 32 columns apart (x0, x0 + 32, x0 + 64, x0 + 96): the 32 lanes of a warp then gather
// from neighbouring addresses, 8-9 sectors per request instead of ~27 with 4 consecutive cells per thread
// (that layout kept the L1 wavefront pipe 91 % busy - the kernel's limiter, profiles/r1c_cfg3_k_dye_advect*).
template <bool SLAB>
__global__ void __launch_bounds__(D4X * D4Y)
k_dye_advect4(const float* __restrict__ din, float* __restrict__ dout, const Geom dg, const float2* __restrict__ vel,
              const uint8_t* __restrict__ obs, const Geom vg, const float* __restrict__ nxt,
              const float* __restrict__ nyt, float rx, float ry, float dt, float speed, float diss,
              int* __restrict__ err) {
    const int x0 = blockIdx.x * (D4X * 4) + threadIdx.x;
    const int y = blockIdx.y * D4Y + threadIdx.y;
    if (x0 >= dg.w || y >= dg.hl) return;
    const int pw = dg.w, ph = dg.hg, vw = vg.w, vh = vg.hg;
    const int gy = y + dg.y0;
    const float ny = nyt[y];
    const float my = (float)(vh - 1), mx = (float)(vw - 1);
    const int vty = (int)clampf(ceilf(ny), 0.0f, my), vby = (int)clampf(floorf(ny), 0.0f, my);
    const float vdy = ny - (float)vby;
    const float2* vrow_t = vel + lin(vg, 0, held_row<SLAB>(vg, vty, err));
    const float2* vrow_b = vel + lin(vg, 0, held_row<SLAB>(vg, vby, err));
    const uint8_t* orow = obs + lin(vg, 0, held_row<SLAB>(vg, (int)(unsigned)ny, err));
    float* drow = dout + lin(dg, 0, y);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int x = x0 + 32 * j;
        if (x >= pw) break;
        const float nx = nxt[x];
        const int vtx = (int)clampf(ceilf(nx), 0.0f, mx), vbx = (int)clampf(floorf(nx), 0.0f, mx);
        const float vdx = nx - (float)vbx;
        const float2 lt = vrow_t[vbx], rt = vrow_t[vtx], lb = vrow_b[vbx], rb = vrow_b[vtx];
        const float vx = mixf(mixf(lb.x, rb.x, vdx), mixf(lt.x, rt.x, vdx), vdy) * rx;
        const float vy = mixf(mixf(lb.y, rb.y, vdx), mixf(lt.y, rt.y, vdx), vdy) * ry;
        const float fx = (float)x - vx * dt * speed;
        const float fy = (float)gy - vy * dt * speed;
        const Corners q = corners(fx, fy, pw, ph);
        const float* drow_t = din + lin(dg, 0, held_row<SLAB>(dg, q.ty, err));
        const float* drow_b = din + lin(dg, 0, held_row<SLAB>(dg, q.by, err));
        const float g1 = mixf(drow_t[q.bx], drow_t[q.tx], q.dx);
        const float g2 = mixf(drow_b[q.bx], drow_b[q.tx], q.dx);
        const float r = mixf(g2, g1, q.dy) * diss;
        __stcs(drow + x, orow[(unsigned)nx] != OBS_FREE ? 0.0f : r);
    }
}

// ref: demo/shaders/demo.ComputeShader.comp:9-21 - the dye value replicated into the four channels of
// an rgba8 (unorm) image: round-to-nearest of clamp(v, 0, 1) * 255 per channel.
__global__ void __launch_bounds__(256) k_dye_rgba8(const float* __restrict__ dye, uint32_t* __restrict__ out, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t c = __float2uint_rn(clampf(dye[i], 0.0f, 1.0f) * 255.0f);
    out[i] = c * 0x01010101u;
}

}  // namespace

int launch_dye_rgba8(const float* dye, uint32_t* out, size_t n, cudaStream_t st) {
    if (!n) return 0;
    k_dye_rgba8<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(dye, out, n);
    return 1;
}

int launch_dye_add(const float* din, float* dout, Geom dg, int r0, int r1, const SplatDBatch& b, cudaStream_t st) {
    if (r1 <= r0) return 0;
    dim3 grid((dg.w + BX - 1) / BX, (r1 - r0 + BY - 1) / BY, 1);
    k_dye_add<<<grid, dim3(BX, BY, 1), 0, st>>>(din, dout, dg, r0, r1, b);
    return 1;
}
int launch_dye_advect(const float* din, float* dout, Geom dg, const float2* vel, const uint8_t* obs, Geom vg,
                      float dt, float speed, float diss, int* err, cudaStream_t st) {
    dim3 grid((dg.w + BX - 1) / BX, (dg.hl + BY - 1) / BY, 1);
    k_dye_advect<<<grid, dim3(BX, BY, 1), 0, st>>>(din, dout, dg, vel, obs, vg, dt, speed, diss, err);
    return 1;
}
int launch_dye_tables(float* nx, float* ny, Geom dg, int vw, int vh, cudaStream_t st) {
    const int n = dg.w > dg.hl ? dg.w : dg.hl;
    k_dye_tables<<<(n + 255) / 256, 256, 0, st>>>(nx, ny, dg, vw, vh);
    return 1;
}
int launch_dye_advect4(const float* din, float* dout, Geom dg, const float2* vel, const uint8_t* obs, Geom vg,
                       const float* nx, const float* ny, float dt, float speed, float diss, int* err, cudaStream_t st) {
    // _ParticleSize / _VelocitySize (:34): IEEE single division, same on host and device
    const float rx = (float)dg.w / (float)vg.w, ry = (float)dg.hg / (float)vg.hg;
    dim3 grid((dg.w / 4 + D4X - 1) / D4X, (dg.hl + D4Y - 1) / D4Y, 1);
    if (dg.hl == dg.hg && vg.hl == vg.hg)
        k_dye_advect4<false><<<grid, dim3(D4X, D4Y, 1), 0, st>>>(din, dout, dg, vel, obs, vg, nx, ny, rx, ry, dt, speed, diss, err);
    else
        k_dye_advect4<true><<<grid, dim3(D4X, D4Y, 1), 0, st>>>(din, dout, dg, vel, obs, vg, nx, ny, rx, ry, dt, speed, diss, err);
    return 1;
}

}  // namespace natrix
