// bgfx_compute.sh - C++ stand-in for bgfx's shader prelude.  TEST INFRASTRUCTURE ONLY.
//
// The reference's compute shaders (natrix/core/shaders/originals/*.comp, demo/shaders/shader.*.comp)
// start with `#include "bgfx_compute.sh"`, a header of the un-vendored bgfx-python 2.0.1 dependency
// (poetry.lock:76-86) that maps one GLSL-like dialect onto GLSL / HLSL / SPIR-V / MSL.  This file is
// that header for a fifth back end: g++.  With it on the include path the UNMODIFIED shader text
// compiles as C++17 (oracle/ref_shim/shader_tu.cpp includes one shader per translation unit, straight
// from /root/reference), so the arithmetic the parity tests are pinned to is the reference's own
// source, not a restatement of it.
//
// What the dialect needs, and the definition each item gets here (everything the shader text itself
// does not define):
//   * vec2 / ivec2 / uvec2 / uvec3 / uvec4 / vec4 with the constructors and swizzles the shaders use;
//   * BUFFER_RO / BUFFER_WR / BUFFER_RW(name, type, slot): a typed view of whatever buffer the runtime
//     (runtime.cpp, mirroring bgfx.setBuffer) has bound to `slot` when the dispatch starts;
//   * `uniform T name;` : a namespace-scope variable the runtime fills by NAME (bgfx.setUniform);
//   * NUM_THREADS(x, y, z): records the work-group size the dispatcher multiplies group counts by;
//   * gl_GlobalInvocationID: thread-local uvec3;
//   * builtins, with their GLSL 4.50 specification definitions evaluated in IEEE binary32, one
//     rounding per operation, no contraction (the TU is compiled with -ffp-contract=off):
//       mix(x, y, a)      = x*(1-a) + y*a          (GLSL 8.3; -DNATRIX_REF_MIX_LERP selects the HLSL
//                                                   lerp form x + a*(y-x), which bgfx's HLSL/Metal
//                                                   back ends would emit - a documented switch)
//       clamp(x, lo, hi)  = min(max(x, lo), hi)
//       distance(a, b)    = length(a - b),  length(v) = sqrt(v.x*v.x + v.y*v.y)
//       dot(a, b)         = a.x*b.x + a.y*b.y
//       inversesqrt(x)    = 1 / sqrt(x)            (correctly rounded sqrt and divide)
//       floor, ceil, abs, max, min: exact
//   * HLSL-style implicit conversions the shader text relies on (it only compiles through bgfx's
//     HLSL-flavoured preprocessor): vec2 -> float takes .x (`float p = _PressureIn[pos]` with
//     `_PressureIn` declared vec2, shader.Poisson.comp:11,26), float -> vec2 replicates
//     (`_PressureOut[pos] = scalar`, shader.Poisson.comp:37; `_Buffer[pos] = 0.0f`,
//     shader.ClearBuffer.comp:17), scalar arguments of mix() widen to the vector argument
//     (shader.AdvectParticle.comp:67-69).
//   * index arithmetic: the shaders compute linear indices in float (`uint pos = gid.y * _Size.x +
//     gid.x` with `uniform vec2 _Size`).  Default build: literally that (builtin unsigned * float),
//     exact up to 2^24 cells.  -DNATRIX_REF_EXACT_INDEX makes `uint` a class whose products with a
//     float are carried in double, i.e. the evident intent, for grids above 2^24 cells.
#ifndef NATRIX_REF_BGFX_COMPUTE_SH
#define NATRIX_REF_BGFX_COMPUTE_SH

#include <cmath>
#include <cstddef>
#include <cstdint>

namespace natrix_ref {

// ----------------------------------------------------------------------------------- scalar uint
#ifdef NATRIX_REF_EXACT_INDEX
struct Index {                      // uint * float, carried exactly
    double v;
};
struct UInt {
    uint32_t v;
    UInt() : v(0) {}
    UInt(uint32_t a) : v(a) {}
    UInt(int a) : v((uint32_t)a) {}
    explicit UInt(float a) : v((uint32_t)a) {}
    explicit UInt(double a) : v((uint32_t)a) {}
    UInt(Index a) : v((uint32_t)a.v) {}
    explicit operator float() const { return (float)v; }
    explicit operator int() const { return (int)v; }
    explicit operator size_t() const { return v; }
};
inline Index operator*(UInt a, float b) { return {(double)a.v * (double)b}; }
inline Index operator+(Index a, UInt b) { return {a.v + (double)b.v}; }
inline UInt operator*(UInt a, UInt b) { return UInt(a.v * b.v); }
inline UInt operator+(UInt a, UInt b) { return UInt(a.v + b.v); }
inline bool operator>=(UInt a, float b) { return (float)a.v >= b; }
inline bool operator==(UInt a, float b) { return (float)a.v == b; }
inline bool operator==(UInt a, unsigned b) { return a.v == b; }
inline float operator-(float a, UInt b) { return a - (float)b.v; }
inline size_t to_index(UInt a) { return a.v; }
inline size_t to_index(Index a) { return (size_t)a.v; }
typedef UInt uint_t;
#else
typedef unsigned int uint_t;
inline size_t to_index(unsigned a) { return a; }
inline size_t to_index(float a) { return (size_t)(unsigned)a; }   // HLSL float -> uint index (shader.AdvectVelocity.comp:43)
#endif

// ----------------------------------------------------------------------------------- vectors
struct ivec2;
struct uvec2;
struct uvec3;

struct vec2 {
    float x, y;
    vec2() : x(0), y(0) {}
    template <class A, class B> vec2(A a, B b) : x((float)a), y((float)b) {}
    explicit vec2(float a) : x(a), y(a) {}          // vec2(delta.x); implicit float -> vec2 only on buffer stores
    explicit vec2(const ivec2& v);
    vec2(const uvec2& v);                            // distance(vec2, gid.xy), shader.AddParticle.comp:30
    explicit vec2(const uvec3& v);                   // vec2(gl_GlobalInvocationID), shader.AddCircleObstacle.comp:25
    operator float() const { return x; }             // HLSL vector -> scalar truncation
    vec2& operator*=(const vec2& o) { x = x * o.x; y = y * o.y; return *this; }
};
struct ivec2 {
    int x, y;
    ivec2() : x(0), y(0) {}
    ivec2(int a, int b) : x(a), y(b) {}
    ivec2(float a, float b) : x((int)a), y((int)b) {}
    explicit ivec2(const vec2& v) : x((int)v.x), y((int)v.y) {}
    explicit ivec2(const uvec2& v);
};
struct uvec2 {
    uint_t x, y;
};
struct uvec3 {
    uint_t x, y, z;
    uvec2 xy;                                        // the only swizzle used; kept in step by the dispatcher
};
struct uvec4 {
    uint_t x, y, z, w;
};
struct vec4 {
    float x, y, z, w;
};
inline vec2::vec2(const ivec2& v) : x((float)v.x), y((float)v.y) {}
inline vec2::vec2(const uvec2& v) : x((float)v.x), y((float)v.y) {}
inline vec2::vec2(const uvec3& v) : x((float)v.x), y((float)v.y) {}
inline ivec2::ivec2(const uvec2& v) : x((int)v.x), y((int)v.y) {}

inline vec2 operator+(const vec2& a, const vec2& b) { return vec2(a.x + b.x, a.y + b.y); }
inline vec2 operator-(const vec2& a, const vec2& b) { return vec2(a.x - b.x, a.y - b.y); }
inline vec2 operator*(const vec2& a, const vec2& b) { return vec2(a.x * b.x, a.y * b.y); }
inline vec2 operator/(const vec2& a, const vec2& b) { return vec2(a.x / b.x, a.y / b.y); }
inline vec2 operator*(const vec2& a, float b) { return vec2(a.x * b, a.y * b); }
inline vec2 operator*(float a, const vec2& b) { return vec2(a * b.x, a * b.y); }
inline vec2 operator/(const vec2& a, float b) { return vec2(a.x / b, a.y / b); }

// ----------------------------------------------------------------------------------- builtins
inline float abs(float a) { return std::fabs(a); }
inline float max(float a, float b) { return a < b ? b : a; }
inline float min(float a, float b) { return b < a ? b : a; }
inline int clamp(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
inline float clamp(float v, float lo, float hi) { return min(max(v, lo), hi); }
inline vec2 clamp(const vec2& v, const vec2& lo, const vec2& hi) {
    return vec2(clamp(v.x, lo.x, hi.x), clamp(v.y, lo.y, hi.y));
}
inline vec2 floor(const vec2& v) { return vec2(std::floor(v.x), std::floor(v.y)); }
inline vec2 ceil(const vec2& v) { return vec2(std::ceil(v.x), std::ceil(v.y)); }
inline float dot(const vec2& a, const vec2& b) { return a.x * b.x + a.y * b.y; }
inline float length(const vec2& v) { return std::sqrt(v.x * v.x + v.y * v.y); }
inline float distance(const vec2& a, const vec2& b) { return length(a - b); }
inline float inversesqrt(float a) { return 1.0f / std::sqrt(a); }
inline float mix(float x, float y, float a) {
#ifdef NATRIX_REF_MIX_LERP
    return x + a * (y - x);
#else
    return x * (1.0f - a) + y * a;
#endif
}
// mix(float, float, vec2) (shader.AdvectParticle.comp:67-69): the scalars widen to the vector argument
inline vec2 mix(float x, float y, const vec2& a) { return vec2(mix(x, y, a.x), mix(x, y, a.y)); }
inline vec2 mix(const vec2& x, const vec2& y, const vec2& a) { return vec2(mix(x.x, y.x, a.x), mix(x.y, y.y, a.y)); }

// ----------------------------------------------------------------------------------- buffers
// The slot table is the runtime's copy of bgfx's compute bindings (bgfx.setBuffer(stage, handle, access)).
// Every BUFFER_* declaration registers itself; the dispatcher re-points all views before a dispatch.
struct BufferView {
    int slot;
    void* base;
    BufferView* next;
};
void register_buffer(BufferView* b);

template <class T> struct Element {                 // one element, with the HLSL scalar <-> vector rules
    T* p;
    operator T() const { return *p; }
    Element& operator=(const T& v) { *p = v; return *this; }
};
template <> struct Element<vec2> {
    vec2* p;
    operator vec2() const { return *p; }
    operator float() const { return p->x; }
    Element& operator=(const vec2& v) { *p = v; return *this; }
    Element& operator=(float v) { p->x = v; p->y = v; return *this; }
    float x() const { return p->x; }
};
template <class T> struct Buffer : BufferView {
    explicit Buffer(int s) { slot = s; base = nullptr; next = nullptr; register_buffer(this); }
    template <class I> Element<T> operator[](I i) const { return Element<T>{(T*)base + to_index(i)}; }
};

// `_VelocityIn[n.x].x` (shader.Divergence.comp:24): member access on the proxy needs the value
template <class T> struct ReadBuffer : BufferView {
    explicit ReadBuffer(int s) { slot = s; base = nullptr; next = nullptr; register_buffer(this); }
    template <class I> const T& operator[](I i) const { return ((const T*)base)[to_index(i)]; }
};

// ----------------------------------------------------------------------------------- per-dispatch state
extern thread_local uvec3 tl_global_invocation_id;
struct GroupSize {
    GroupSize(int x, int y, int z);
};
}  // namespace natrix_ref

// The names the shader text uses are global in GLSL.
using natrix_ref::vec2;
using natrix_ref::ivec2;
using natrix_ref::uvec2;
using natrix_ref::uvec3;
using natrix_ref::uvec4;
using natrix_ref::vec4;
using natrix_ref::abs;
using natrix_ref::max;
using natrix_ref::min;
using natrix_ref::clamp;
using natrix_ref::floor;
using natrix_ref::ceil;
using natrix_ref::dot;
using natrix_ref::length;
using natrix_ref::distance;
using natrix_ref::inversesqrt;
using natrix_ref::mix;
#ifdef NATRIX_REF_EXACT_INDEX
#define uint natrix_ref::UInt
#else
typedef unsigned int uint;
#endif

#define gl_GlobalInvocationID (natrix_ref::tl_global_invocation_id)
#define uniform
#define BUFFER_RO(_name, _type, _slot) static natrix_ref::ReadBuffer<_type> _name(_slot)
#define BUFFER_WR(_name, _type, _slot) static natrix_ref::Buffer<_type> _name(_slot)
#define BUFFER_RW(_name, _type, _slot) static natrix_ref::Buffer<_type> _name(_slot)
#define NUM_THREADS(_x, _y, _z) static natrix_ref::GroupSize natrix_ref_group_size(_x, _y, _z);

#endif  // NATRIX_REF_BGFX_COMPUTE_SH
