// render.cu - the demo's frame as an RGBA8 image, so that a headless viewer
// can replace the bgfx render path.
// ref: demo/shaders/demo.ComputeShader.comp:9-21        dye -> rgba8 texture
//      demo/shaders/demo.FieldFragmentShader.frag:21-33,73-99   plasma(fbm(texel.x)), alpha = texel.x
//      demo/shaders/demo.QuiverFragmentShader.frag:14-70        velocity arrows, white, alpha = 1 - dist
//      demo/simulation_demo.py:94 (clear colour 0x1a0427ff), :249-281 (alpha blending, pass order)
//
// Conventions (the reference leaves them to the graphics backend; stated here once):
//   - the full-screen quad is taken to cover the framebuffer exactly and the framebuffer has the dye
//     grid's size (simulation_demo.py:109-111, :121-128), so every fragment samples one texel centre;
//   - framebuffer y grows upwards (gl_FragCoord, dye row 0 at the bottom); the image is written top row first;
//   - the render target is RGBA8: the result of each pass is rounded to 8 bits before the next blends over it.
#include "kernels.h"

namespace natrix {
namespace {

__device__ __forceinline__ float fractf(float x) { return x - floorf(x); }
// rand / noise / fbm: demo.FieldFragmentShader.frag:73-92
__device__ __forceinline__ float rand1(float n) { return fractf(sinf(n) * 43758.5453123f); }
__device__ __forceinline__ float noise1(float p) {
    const float fl = floorf(p), fc = fractf(p);
    return mixf(rand1(fl), rand1(fl + 1.0f), fc);
}
__device__ float fbm1(float x) {
    float v = 0.0f, a = 0.5f;
    for (int i = 0; i < 5; ++i) {
        v = v + a * noise1(x);
        x = x * 2.0f + 100.0f;
        a = a * 0.5f;
    }
    return v;
}
// plasma: demo.FieldFragmentShader.frag:21-33 (Horner, one channel)
__device__ __forceinline__ float horner6(float t, float c0, float c1, float c2, float c3, float c4, float c5, float c6) {
    return c0 + t * (c1 + t * (c2 + t * (c3 + t * (c4 + t * (c5 + t * c6)))));
}

// The texture is rgba8, so texel.x takes 256 values: the whole colour map is a 256-entry table.
__global__ void k_field_lut(float4* __restrict__ lut) {
    const int k = threadIdx.x;
    const float c = (float)k / 255.0f;
    const float t = fbm1(c);
    float4 o;
    o.x = horner6(t, 0.05873234392399702f, 2.176514634195958f, -2.689460476458034f, 6.130348345893603f,
                  -11.10743619062271f, 10.02306557647065f, -3.658713842777788f);
    o.y = horner6(t, 0.02333670892565664f, 0.2383834171260182f, -7.455851135738909f, 42.3461881477227f,
                  -82.66631109428045f, 71.41361770095349f, -22.93153465461149f);
    o.z = horner6(t, 0.5433401826748754f, 0.7539604599784036f, 3.110799939717086f, -28.51885465332158f,
                  60.13984767418263f, -54.07218655560067f, 18.19190778539828f);
    o.w = c;
    lut[k] = o;
}

__device__ __forceinline__ float unorm8(float c) { return rintf(clampf(c, 0.0f, 1.0f) * 255.0f); }

// BGFX_STATE_BLEND_ALPHA: src * src.a + dst * (1 - src.a) on all four channels, result stored as rgba8
__device__ __forceinline__ float4 blend8(float4 src, float4 dst) {
    const float ia = 1.0f - src.w;
    float4 o;
    o.x = unorm8(src.x * src.w + dst.x * ia) / 255.0f;
    o.y = unorm8(src.y * src.w + dst.y * ia) / 255.0f;
    o.z = unorm8(src.z * src.w + dst.z * ia) / 255.0f;
    o.w = unorm8(src.w * src.w + dst.w * ia) / 255.0f;
    return o;
}

// demo.QuiverFragmentShader.frag:20-28
__device__ __forceinline__ float line_dist(float px, float py, float ax, float ay, float bx, float by) {
    const float cx = (ax + bx) * 0.5f, cy = (ay + by) * 0.5f;
    const float ex = bx - ax, ey = by - ay;
    const float len = sqrtf(ex * ex + ey * ey);
    const float dx = ex / len, dy = ey / len;
    const float rx = px - cx, ry = py - cy;
    const float d1 = fabsf(rx * dy + ry * (-dx));
    const float d2 = fabsf(rx * dx + ry * dy) - 0.5f * len;
    return fmaxf(d1, d2);
}

__global__ void __launch_bounds__(256)
k_render_frame(const float* __restrict__ dye, const float4* __restrict__ lut, const float2* __restrict__ vel,
               uint32_t* __restrict__ out, int w, int h, int vw, int vh, float tile) {
    const int x = blockIdx.x * 32 + threadIdx.x;
    const int r = blockIdx.y * 8 + threadIdx.y;          // image row, top first
    if (x >= w || r >= h) return;
    const int yb = h - 1 - r;                            // framebuffer / dye row
    // demo.ComputeShader.comp:20 then demo.FieldFragmentShader.frag:96-99
    const int c8 = (int)unorm8(dye[(size_t)yb * w + x]);
    const float4 l = lut[c8];
    float4 px = blend8(make_float4(l.x, l.y, l.z, l.w),
                       make_float4(26.0f / 255.0f, 4.0f / 255.0f, 39.0f / 255.0f, 1.0f));   // clear 0x1a0427ff
    if (tile > 0.0f) {
        // demo.QuiverFragmentShader.frag:66-70 with gl_FragCoord = pixel centre
        const float fx = (float)x + 0.5f, fy = (float)yb + 0.5f;
        const float cx = (floorf(fx / tile) + 0.5f) * tile, cy = (floorf(fy / tile) + 0.5f) * tile;   // :16-18
        // _field(centre) :49-64 - note mix(a, b, vec2(t, 0)): only the x component is interpolated
        const float qx = (1.0f - cx / (float)w) * (float)vw, qy = (1.0f - cy / (float)h) * (float)vh;
        const Corners c = corners(qx, qy, vw, vh);
        const float2 lt = vel[(size_t)c.ty * vw + c.bx], rt = vel[(size_t)c.ty * vw + c.tx];
        const float2 lb = vel[(size_t)c.by * vw + c.bx], rb = vel[(size_t)c.by * vw + c.tx];
        const float h1x = mixf(lt.x, rt.x, c.dx), h2x = mixf(lb.x, rb.x, c.dx);
        float vx = -1.0f * mixf(h2x, h1x, c.dy) * ((float)w / (float)vw);
        float vy = -1.0f * lb.y * ((float)h / (float)vh);
        vx = vx * tile * 0.4f;
        vy = vy * tile * 0.4f;
        // _vector :30-47
        const float pxr = fx - cx, pyr = fy - cy;
        float mag = sqrtf(vx * vx + vy * vy), dist = 1.0f;
        if (mag > 0.001f) {
            const float dx = vx / mag, dy = vy / mag;
            mag = clampf(mag, 0.0f, tile * 0.5f);
            vx = dx * mag;
            vy = dy * mag;
            const float shaft = line_dist(pxr, pyr, vx, vy, -vx, -vy);
            const float h1 = line_dist(pxr, pyr, vx, vy, 0.4f * vx + 0.2f * (-vy), 0.4f * vy + 0.2f * vx);
            const float h2 = line_dist(pxr, pyr, vx, vy, 0.4f * vx + 0.2f * vy, 0.4f * vy + 0.2f * (-vx));
            dist = fminf(shaft, fminf(h1, h2));
        }
        px = blend8(make_float4(1.0f, 1.0f, 1.0f, 1.0f - clampf(dist, 0.0f, 1.0f)), px);
    }
    const uint32_t R = (uint32_t)(px.x * 255.0f + 0.5f), G = (uint32_t)(px.y * 255.0f + 0.5f);
    const uint32_t B = (uint32_t)(px.z * 255.0f + 0.5f), A = (uint32_t)(px.w * 255.0f + 0.5f);
    out[(size_t)r * w + x] = R | (G << 8) | (B << 16) | (A << 24);
}

}  // namespace

int launch_field_lut(float4* lut, cudaStream_t st) {
    k_field_lut<<<1, 256, 0, st>>>(lut);
    return 1;
}

int launch_render_frame(const float* dye, const float4* lut, const float2* vel, uint32_t* out, int w, int h, int vw,
                        int vh, float tile, cudaStream_t st) {
    dim3 grid((w + 31) / 32, (h + 7) / 8, 1);
    k_render_frame<<<grid, dim3(32, 8, 1), 0, st>>>(dye, lut, vel, out, w, h, vw, vh, tile);
    return 1;
}

}  // namespace natrix
