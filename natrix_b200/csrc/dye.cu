// dye.cu - the "smooth particles area" (dye) kernels.
// ref: demo/shaders/shader.AddParticle.comp:25-34, demo/shaders/shader.AdvectParticle.comp:21-70,
//      demo/smooth_particles_area.py:70-104.
#include "kernels.h"

namespace natrix {
namespace {

constexpr int BX = 64, BY = 4, BT = BX * BY;   // 2-D blocks: back-traced gathers of nearby rows hit L1

// K splats applied in sequence per cell: identical arithmetic to K AddParticle dispatches
// (each dispatch is a pure per-cell map, the ping-pong flip carries no cross-cell dependency).
__global__ void __launch_bounds__(BT)
k_dye_add(const float* __restrict__ din, float* __restrict__ dout, int pw, int ph,
          const __grid_constant__ SplatDBatch b) {
    const int x = blockIdx.x * BX + threadIdx.x;
    const int y = blockIdx.y * BY + threadIdx.y;
    if (x >= pw || y >= ph) return;
    const size_t pos = (size_t)y * pw + x;
    float v = din[pos];
    const float fxp = (float)x, fyp = (float)y;
    for (int i = 0; i < b.n; ++i) {
        const SplatD s = b.s[i];
        const float ex = s.sx - fxp, ey = s.sy - fyp;
        const float len = sqrtf(ex * ex + ey * ey);
        if (len <= s.r) v = clampf(v + s.value * (s.r - len) / s.r, 0.0f, 255.0f);
    }
    dout[pos] = v;
}

__global__ void __launch_bounds__(BT)
k_dye_advect(const float* __restrict__ din, float* __restrict__ dout, int pw, int ph,
             const float2* __restrict__ vel, const uint8_t* __restrict__ obs, int vw, int vh,
             float dt, float speed, float diss) {
    const int x = blockIdx.x * BX + threadIdx.x;
    const int y = blockIdx.y * BY + threadIdx.y;
    if (x >= pw || y >= ph) return;
    const size_t pos = (size_t)y * pw + x;
    // fNormalisedPos (:46) and the obstacle lookup at its truncation (:47-50)
    const float nx = ((float)x / (float)pw) * (float)vw;
    const float ny = ((float)y / (float)ph) * (float)vh;
    const size_t opos = (size_t)(unsigned)ny * vw + (unsigned)nx;
    if (obs[opos] != OBS_FREE) { dout[pos] = 0.0f; return; }
    // GetVelocity (:21-35): bilinear sample of the velocity grid, scaled to dye cells
    const Corners c = corners(nx, ny, vw, vh);
    const float2 lt = vel[(size_t)c.ty * vw + c.bx], rt = vel[(size_t)c.ty * vw + c.tx];
    const float2 lb = vel[(size_t)c.by * vw + c.bx], rb = vel[(size_t)c.by * vw + c.tx];
    const float rx = (float)pw / (float)vw, ry = (float)ph / (float)vh;
    const float vx = mixf(mixf(lb.x, rb.x, c.dx), mixf(lt.x, rt.x, c.dx), c.dy) * rx;
    const float vy = mixf(mixf(lb.y, rb.y, c.dx), mixf(lt.y, rt.y, c.dx), c.dy) * ry;
    // back-trace in dye cells and gather the dye with the same clamp rule (:57-69)
    const float fx = (float)x - vx * dt * speed;
    const float fy = (float)y - vy * dt * speed;
    const Corners q = corners(fx, fy, pw, ph);
    const float g1 = mixf(din[(size_t)q.ty * pw + q.bx], din[(size_t)q.ty * pw + q.tx], q.dx);
    const float g2 = mixf(din[(size_t)q.by * pw + q.bx], din[(size_t)q.by * pw + q.tx], q.dx);
    dout[pos] = mixf(g2, g1, q.dy) * diss;
}

}  // namespace

int launch_dye_add(const float* din, float* dout, int pw, int ph, const SplatDBatch& b, cudaStream_t st) {
    dim3 grid((pw + BX - 1) / BX, (ph + BY - 1) / BY, 1);
    k_dye_add<<<grid, dim3(BX, BY, 1), 0, st>>>(din, dout, pw, ph, b);
    return 1;
}
int launch_dye_advect(const float* din, float* dout, int pw, int ph, const float2* vel, const uint8_t* obs,
                      int vw, int vh, float dt, float speed, float diss, cudaStream_t st) {
    dim3 grid((pw + BX - 1) / BX, (ph + BY - 1) / BY, 1);
    k_dye_advect<<<grid, dim3(BX, BY, 1), 0, st>>>(din, dout, pw, ph, vel, obs, vw, vh, dt, speed, diss);
    return 1;
}

}  // namespace natrix
