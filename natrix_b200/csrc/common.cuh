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
enum : uint8_t { NB_L = 1, NB_R = 2, NB_B = 4, NB_T = 8 };

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

// streaming global accesses: data touched once per kernel should not pollute L1
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_stream(float4* p, float4 v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

}  // namespace natrix
