// common.cuh - shared host/device definitions for libnatrix_b200 (sm_100a only).
//
// Arithmetic rules (so that every kernel is bit-identical to oracle/natrix_oracle.py):
//   * this library is compiled with -fmad=false: no FMA contraction anywhere;
//   * IEEE division and sqrt (nvcc defaults, never --use_fast_math);
//   * GLSL mix(a,b,t) = a*(1-t) + b*t; inversesqrt(x) = 1/sqrtf(x);
//   * expressions keep the operand order of the reference shaders.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

#include "../../include/natrix_b200.h"

#ifndef __CUDA_ARCH_LIST__
#endif

namespace natrix {

// Geometry of one row slab (the whole grid when y0 == 0 && hl == hg).
// Arrays are indexed by LOCAL row ly in [-halo, hl+halo); global row gy = y0 + ly.
struct Geom {
    int w;      // width (cells)
    int hg;     // global height
    int y0;     // global row of local row 0
    int hl;     // rows owned by this slab
    int halo;   // extra rows allocated above and below
};

// encodings of the 1-byte obstacle map (information-equivalent to the reference's float2:
// every consumer only tests x > 0 || y > 0)
enum : uint8_t { OBS_FREE = 0, OBS_DYNAMIC = 1 /* (1,0) */, OBS_STATIC = 2 /* (0,1) */ };

// bits of the blocked-neighbour mask consumed by the Jacobi and gradient kernels:
// set when that neighbour is solid OR lies outside the global domain; in both cases
// shader.Poisson.comp:32-35 / shader.SubtractGradient.comp:35-42 use the centre pressure.
enum : uint8_t { NB_L = 1, NB_R = 2, NB_B = 4, NB_T = 8,
                 // The Jacobi kernels read the divergence PRE-SCALED: b4 = 0.25 * b, so that a sweep ends in ONE fused
                 // multiply-add, fma(sum, 0.25, -b4), instead of the shader's subtract-then-scale.  The two are bit-identical
                 // whenever 0.25 * b is exact (proof and brute-force check in DESIGN.md 5.1: scaling by a power of two
                 // commutes with rounding, and where the scaled result is subnormal the shader's first rounding is either
                 // exact or a tie that rounds to the same multiple of four).  0.25 * b is inexact only for a non-zero
                 // |b| < 2^-124 whose last two mantissa bits are not both zero; such a cell stores b itself in the scaled
                 // field and carries this bit, which sends its row to the select body, where the shader's two-step form is used.
                 NB_RAW = 16 };

// the scaled divergence the Jacobi kernels read, and whether the cell has to take the unfused form (see NB_RAW)
__device__ __forceinline__ float scaled_divergence(float b, bool* raw) {
    const float b4 = b * 0.25f;
    *raw = b4 * 4.0f != b;
    return *raw ? b : b4;
}

__host__ __device__ __forceinline__ int clampi(int v, int lo, int hi) {
    return v < lo ? lo : (v > hi ? hi : v);
}
__device__ __forceinline__ float clampf(float v, float lo, float hi) {
    return fminf(fmaxf(v, lo), hi);
}
__device__ __forceinline__ float mixf(float a, float b, float t) {
    return a * (1.0f - t) + b * t;   // -fmad=false keeps the two roundings
}

// Clamped floor/ceil corners with the UNclamped delta
// (ref: shader.AdvectVelocity.comp:38-42).
struct Corners { int tx, ty, bx, by; float dx, dy; };
__device__ __forceinline__ Corners corners(float fx, float fy, int w, int h) {
    Corners c;
    const float mx = (float)(w - 1), my = (float)(h - 1);
    c.tx = (int)clampf(ceilf(fx), 0.0f, mx);
    c.ty = (int)clampf(ceilf(fy), 0.0f, my);
    c.bx = (int)clampf(floorf(fx), 0.0f, mx);
    c.by = (int)clampf(floorf(fy), 0.0f, my);
    c.dx = fx - (float)c.bx;
    c.dy = fy - (float)c.by;
    return c;
}

__device__ __forceinline__ ptrdiff_t lin(const Geom& g, int x, int ly) {
    return (ptrdiff_t)ly * g.w + x;
}

// Velocity load with shader.InitBoundaries.comp:14-34 folded in when FOLD: the four border lines
// of the READ buffer read as zero (the in-place zeroing is not observable after the step).
template <bool FOLD>
__device__ __forceinline__ float2 load_vel(const float2* __restrict__ v, const Geom& g, int x, int gy) {
    if (FOLD && (x == 0 || x == g.w - 1 || gy == 0 || gy == g.hg - 1)) return make_float2(0.0f, 0.0f);
    return v[lin(g, x, gy - g.y0)];
}

// ref: shader.AdvectVelocity.comp:36-49 for one non-solid cell (x, gy); reports gathers that leave the
// rows this slab holds through *err (multi-GPU only; the full grid can never trip it)
template <bool FOLD>
__device__ __forceinline__ float2 advect_cell(const float2* __restrict__ vin, const Geom& g, int x, int gy,
                                              float2 vel, float dt, float speed, float diss, int* __restrict__ err) {
    const float fx = (float)x - vel.x * dt * speed;
    const float fy = (float)gy - vel.y * dt * speed;
    Corners c = corners(fx, fy, g.w, g.hg);
    const int lo = g.y0 - g.halo, hi = g.y0 + g.hl + g.halo - 1;
    if (c.by < lo || c.ty > hi) {
        *err = 1;
        c.by = clampi(c.by, lo, hi);
        c.ty = clampi(c.ty, lo, hi);
    }
    const float2 lt = load_vel<FOLD>(vin, g, c.bx, c.ty);
    const float2 rt = load_vel<FOLD>(vin, g, c.tx, c.ty);
    const float2 lb = load_vel<FOLD>(vin, g, c.bx, c.by);
    const float2 rb = load_vel<FOLD>(vin, g, c.tx, c.by);
    const float h1x = mixf(lt.x, rt.x, c.dx), h1y = mixf(lt.y, rt.y, c.dx);
    const float h2x = mixf(lb.x, rb.x, c.dx), h2y = mixf(lb.y, rb.y, c.dx);
    float2 o;
    o.x = clampf(mixf(h2x, h1x, c.dy) * diss, -1.0f, 1.0f);
    o.y = clampf(mixf(h2y, h1y, c.dy) * diss, -1.0f, 1.0f);
    return o;
}

// ref: shader.ApplyVorticity.comp:33-39 - the confinement force times dt for one cell
__device__ __forceinline__ float2 confinement_force(float wL, float wR, float wB, float wT, float wC, float scale,
                                                    float dt) {
    float fx = 0.5f * (fabsf(wT) - fabsf(wB));
    float fy = 0.5f * (fabsf(wR) - fabsf(wL));
    const float m = fmaxf(2.4414e-4f, fx * fx + fy * fy);
    const float inv = 1.0f / sqrtf(m);
    fx = fx * inv;
    fy = fy * inv;
    const float k = scale * wC;
    fx = fx * k;
    fy = fy * (-k);
    return make_float2(fx * dt, fy * dt);
}

// m ? a : b for m all-ones / all-zero, without occupying a predicate register
__device__ __forceinline__ float bitsel(float a, float b, uint32_t m) {
    return __uint_as_float((__float_as_uint(a) & m) | (__float_as_uint(b) & ~m));
}

// streaming stores: data written once per kernel should not pollute L1
__device__ __forceinline__ void stg_stream(float4* p, float4 v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

}  // namespace natrix
