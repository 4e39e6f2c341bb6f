// jacobi_tb.cu - temporally blocked pressure-Jacobi sweeps for sm_100a.
//
// ref: shader.Poisson.comp:24-37 applied `depth` times (fluid_simulator.py:251-255).  One launch
// advances the pressure field by T = depth sweeps while reading p, div and the blocked-neighbour
// mask once and writing p once: 13 B/cell per T sweeps instead of 20 B/cell per sweep.
//
// Decomposition.  The output rows are cut into tiles of SW = 128 columns x CH rows; ONE WARP owns
// one tile and needs no other warp: a lane holds 4 adjacent columns (one float4, so its shared-
// memory reads are bank-conflict free) and keeps, for every time level 0..T-1, the two most
// recent rows in registers (2*T*4 floats).  The warp marches down its tile
// one input row at a time; when input row i arrives, level t produces row i-t from the three rows
// i-t-1, i-t, i-t+1 of level t-1 (left/right neighbours across lanes by warp shuffle), so row i-T
// of level T leaves the registers T rows behind the load front.  Halo cells (T columns on each
// side of the strip, T rows above and below the chunk) are recomputed redundantly; they see
// exactly the same inputs in exactly the same order as in a 1-sweep-per-launch kernel, so the
// result is bit-identical.
//
// Staging.  Input rows come through TMA (cp.async.bulk.tensor.2d, SASS UTMALDG): each warp runs
// its own ring of 4-row x 128-column boxes for p (8 rows), div (16 rows: a div row is needed again by every
// level for T more iterations) and the mask (8 rows), completed on 2 warp-private mbarriers; lane 0 puts
// the next box in flight as soon as the current one has landed.  Out-of-bounds parts of a box (strip halos beyond the
// grid, rows beyond the slab) are zero-filled by TMA; those values only ever feed cells whose
// results are discarded, because every in-domain cell next to the edge carries the "blocked"
// bit for that direction and substitutes its own pressure (the shader's clamp-to-edge rule).
//
// Obstacles.  Warps whose rows in flight have an all-zero mask run a select-free body
// (4 FP instructions per cell-sweep); otherwise a body with 4 selects per cell.
//
// Scaled divergence.  The kernel reads b4 = 0.25 b (written next to the divergence by the stage that computes it)
// and ends a sweep in one fused multiply-add, fma(sum, 0.25, -b4), bit-identical to the shader's (sum - b) * 0.25;
// cells where 0.25 b would be inexact carry NB_RAW in the mask and take the two-step form (common.cuh).
#include <cuda.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <type_traits>
#include <vector>

#include "jacobi_tb.h"
#include "kernels.h"

namespace natrix {

namespace {

constexpr int SW = 128;                 // strip width: columns per warp tile, 4 per lane
// Launch shapes (V): 0 = 8 warps x 2 blocks/SM (128 regs/thread); 1 = 12 warps x 1 block/SM (168 regs/thread)
template <int V>
struct Shape {
    static constexpr int WARPS = V == 0 ? 8 : V == 1 ? 12 : V == 2 ? 14 : 15;     // 2: 144 regs/thread, 3: 136
    static constexpr int BLOCKS = V == 0 ? 2 : 1;
};
constexpr int NUM_SHAPES = 4;
constexpr int SHAPE_WARPS[NUM_SHAPES] = {8, 12, 14, 15};
constexpr int P_ROWS = 8, D_ROWS = 16;
constexpr int GROUP = 4;                // rows per TMA box
constexpr int NBAR = 2;                 // mbarriers per warp: the group being consumed + the one in flight

// The mask is 1 byte per cell, and TMA wants the first byte of a box 16-byte aligned in global
// memory: the strip origin x0 is only a multiple of 4, so the mask box starts at x0 rounded down
// to 16 and is MBOX = 144 bytes wide; each 4-row box gets its own 128 B aligned slot.
constexpr int MBOX = SW + 16;
constexpr int M_SLOT = 640;             // >= GROUP * MBOX, multiple of 128
constexpr int M_SLOTS = 4;             // 16 rows: like div, a mask row is needed again for T more iterations
static_assert(GROUP * MBOX <= M_SLOT && M_SLOT % 128 == 0, "mask slot too small");
static_assert(P_ROWS == 2 * GROUP && D_ROWS >= JACOBI_TB_MAX_DEPTH + 2 * GROUP, "ring depths");

struct __align__(128) WarpSmem {
    float p[P_ROWS][SW];
    float d[D_ROWS][SW];
    uint8_t m[M_SLOTS][M_SLOT];
    unsigned long long bar[NBAR];
    unsigned char pad[128 - NBAR * 8];
};
static_assert(sizeof(WarpSmem) % 128 == 0, "per-warp shared memory must keep 128 B alignment");


struct TBParams {
    float* pout;      // local row 0 of the output field
    int w;            // grid width
    int halo;         // tensor-map row = local row + halo
    int r0, r1;       // output rows [r0, r1)
    const int4* tiles;    // [ntiles] (strip, first output row, end output row, -) in device memory
    int ntiles;
    int hx;           // halo columns on each side of a strip (>= T, multiple of 4)
    unsigned long long* trace;   // developer tracing (NATRIX_TB_TRACE): per tile (start ns, end ns, smid, -), else null
};
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ unsigned smid() {
    unsigned v;
    asm volatile("mov.u32 %0, %smid;" : "=r"(v));
    return v;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int x, int y, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(dst), "l"(map), "r"(x), "r"(y), "r"(bar) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

// shared-memory byte addresses of this lane's 16-byte column group in the three rings
// plus all-ones flags for the lane holding grid column 0 (as its column 0) / width-1 (as column 3)
struct LaneAddr { uint32_t p, d, m, edge_l, edge_r, mkeep; };

// blocked-neighbour bits (one byte per column) of input row i for this lane's 4 columns, straight
// from the TMA-filled ring; the grid-edge L / R bits are dropped (handled through edge_l / edge_r)
__device__ __forceinline__ uint32_t raw_mask_of_row(const LaneAddr& sa, int i) {
    return lds32(sa.m + (uint32_t)((i >> 2) & (M_SLOTS - 1)) * M_SLOT + (uint32_t)(i & 3) * MBOX);
}
__device__ __forceinline__ uint32_t mask_of_row(const LaneAddr& sa, int i) { return raw_mask_of_row(sa, i) & sa.mkeep; }

// One input row: advance every time level by one row.  a[t][.] holds the two newest rows of
// level t (slot PAR = older, PAR^1 = newer); afterwards slot PAR holds the newest.
// Column j of a lane is strip column 4*lane + j.
template <int T, bool PZERO, int PAR, bool SLOW>
__device__ __forceinline__ void row_step(float (&a)[T][2][4], const LaneAddr& sa, int i, int lane,
                                         float (&out)[4]) {
    constexpr unsigned FULL = 0xffffffffu;
    float nw[4];
    if (PZERO) {
#pragma unroll
        for (int c = 0; c < 4; ++c) nw[c] = 0.0f;
    } else {
        const float4 v = lds128(sa.p + (uint32_t)(i & (P_ROWS - 1)) * (SW * 4));
        nw[0] = v.x; nw[1] = v.y; nw[2] = v.z; nw[3] = v.w;
    }
    const int lane_l = (lane + 31) & 31, lane_r = (lane + 1) & 31;
#pragma unroll
    for (int t = 1; t <= T; ++t) {
        float(&old)[4] = a[t - 1][PAR];
        float(&mid)[4] = a[t - 1][PAR ^ 1];
        const float4 dv = lds128(sa.d + (uint32_t)((i - t) & (D_ROWS - 1)) * (SW * 4));
        const float d[4] = {dv.x, dv.y, dv.z, dv.w};
        // left / right neighbours across lanes (lanes 0 and 31 receive wrapped garbage for the
        // strip's outermost columns, which lie in the discarded halo)
        // at the grid's left / right edge the neighbour is the cell itself (clamp-to-edge), which
        // is what the blocked bit would select; doing it here keeps edge strips on the fast body
        const float sl = bitsel(mid[0], __shfl_sync(FULL, mid[3], lane_l), sa.edge_l);
        const float sr = bitsel(mid[3], __shfl_sync(FULL, mid[0], lane_r), sa.edge_r);
        const uint32_t mrow_bits = SLOW ? mask_of_row(sa, i - t) : 0u;
        float res[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float C = mid[j];
            float x1 = j > 0 ? mid[j - 1] : sl;
            float x2 = j < 3 ? mid[j + 1] : sr;
            float y1 = old[j];       // row - 1 ("B")
            float y2 = nw[j];        // row + 1 ("T")
            if (SLOW) {
                const uint32_t bits = mrow_bits >> (8 * j);
                x1 = (bits & NB_L) ? C : x1;
                x2 = (bits & NB_R) ? C : x2;
                y1 = (bits & NB_B) ? C : y1;
                y2 = (bits & NB_T) ? C : y2;
            }
            const float sum = x1 + x2 + y1 + y2;
            res[j] = __fmaf_rn(sum, 0.25f, -d[j]);
            if (SLOW && ((mrow_bits >> (8 * j)) & NB_RAW)) res[j] = (sum - d[j]) * 0.25f;    // d holds b itself here
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) { old[c] = nw[c]; nw[c] = res[c]; }
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) out[c] = nw[c];
}

// Packed variant of row_step: the same arithmetic on Blackwell's 2-wide fp32 instructions
// (FADD2 / FMUL2, PTX add/mul.rn.f32x2: each half is an IEEE-rounded fp32 operation, so results
// stay bit-identical).  A lane's 4 columns are two aligned register pairs (c0,c1), (c2,c3).  The
// horizontal sum x1 + x2 mixes columns across pairs, so it is formed by scalar adds written
// straight into an aligned pair; the vertical neighbours, the divergence and the scale are
// already pair-aligned: 2 + 4 instructions per 2 cells instead of 10.
__device__ __forceinline__ float2 neg2(float2 v) { return make_float2(-v.x, -v.y); }

// A/B switches for timing experiments only (scripts/ab_build.sh); the shipped library is built with the defaults.
// NATRIX_TB_FMA=0 ends a sweep in add + multiply on the SCALED divergence: same instruction mix as the unfused
// form, wrong numbers.
#ifndef NATRIX_TB_FMA
#define NATRIX_TB_FMA 1
#endif
#ifndef NATRIX_TB_FILL
#define NATRIX_TB_FILL 0      // measured: the fill bodies are cold code (instruction-cache misses) and cost what they save
#endif
// the last two operations of a sweep on a pair of cells: fma(sum, 0.25, -b4) == (sum - b) * 0.25 (common.cuh NB_RAW)
__device__ __forceinline__ float2 finish2(float2 sum, float2 b4) {
#if NATRIX_TB_FMA
    return __ffma2_rn(sum, make_float2(0.25f, 0.25f), neg2(b4));
#else
    return __fmul2_rn(__fadd2_rn(sum, neg2(b4)), make_float2(0.25f, 0.25f));
#endif
}

// RAW: some cell of the rows in flight carries NB_RAW (its ring entry is b itself, not 0.25 b): those cells redo the
// last step in the shader's two-step form.  Practically never instantiated at run time (a non-zero |b| < 2^-124).
template <int T, bool PZERO, int PAR, bool SLOW, bool RAW = false>
__device__ __forceinline__ void row_step2(float2 (&a)[T][2][2], const LaneAddr& sa, int i, int lane,
                                          float2 (&out)[2]) {
    constexpr unsigned FULL = 0xffffffffu;
    float2 nw[2];
    if (PZERO) {
        nw[0] = nw[1] = make_float2(0.0f, 0.0f);
    } else {
        const float4 v = lds128(sa.p + (uint32_t)(i & (P_ROWS - 1)) * (SW * 4));
        nw[0] = make_float2(v.x, v.y);
        nw[1] = make_float2(v.z, v.w);
    }
    const int lane_l = (lane + 31) & 31, lane_r = (lane + 1) & 31;
#pragma unroll
    for (int t = 1; t <= T; ++t) {
        float2(&old)[2] = a[t - 1][PAR];
        float2(&mid)[2] = a[t - 1][PAR ^ 1];
        const float4 dv = lds128(sa.d + (uint32_t)((i - t) & (D_ROWS - 1)) * (SW * 4));
        const float sl = bitsel(mid[0].x, __shfl_sync(FULL, mid[1].y, lane_l), sa.edge_l);
        const float sr = bitsel(mid[1].y, __shfl_sync(FULL, mid[0].x, lane_r), sa.edge_r);
        float2 s0, s1, y10 = old[0], y11 = old[1], y20 = nw[0], y21 = nw[1];
        uint32_t raw = 0u;
        if (SLOW) {
            const uint32_t m = mask_of_row(sa, i - t);
            if (RAW) raw = m & (0x01010101u * NB_RAW);
            const float c0 = mid[0].x, c1 = mid[0].y, c2 = mid[1].x, c3 = mid[1].y;
            s0.x = ((m & (NB_L << 0)) ? c0 : sl) + ((m & (NB_R << 0)) ? c0 : c1);
            s0.y = ((m & (NB_L << 8)) ? c1 : c0) + ((m & (NB_R << 8)) ? c1 : c2);
            s1.x = ((m & (NB_L << 16)) ? c2 : c1) + ((m & (NB_R << 16)) ? c2 : c3);
            s1.y = ((m & (NB_L << 24)) ? c3 : c2) + ((m & (NB_R << 24)) ? c3 : sr);
            y10.x = (m & (NB_B << 0)) ? c0 : y10.x;   y20.x = (m & (NB_T << 0)) ? c0 : y20.x;
            y10.y = (m & (NB_B << 8)) ? c1 : y10.y;   y20.y = (m & (NB_T << 8)) ? c1 : y20.y;
            y11.x = (m & (NB_B << 16)) ? c2 : y11.x;  y21.x = (m & (NB_T << 16)) ? c2 : y21.x;
            y11.y = (m & (NB_B << 24)) ? c3 : y11.y;  y21.y = (m & (NB_T << 24)) ? c3 : y21.y;
        } else {
            s0 = make_float2(sl + mid[0].y, mid[0].x + mid[1].x);
            s1 = make_float2(mid[0].y + mid[1].y, mid[1].x + sr);
        }
        // ((x1 + x2) + y1) + y2 in the shader's left-to-right order, then fma(sum, 0.25, -b4) == (sum - b) * 0.25
        const float2 t0 = __fadd2_rn(__fadd2_rn(s0, y10), y20), t1 = __fadd2_rn(__fadd2_rn(s1, y11), y21);
        float2 r0 = finish2(t0, make_float2(dv.x, dv.y));
        float2 r1 = finish2(t1, make_float2(dv.z, dv.w));
        if (SLOW && RAW && raw) {                        // the ring holds b itself for these cells
            if (raw & (NB_RAW << 0)) r0.x = (t0.x - dv.x) * 0.25f;
            if (raw & (NB_RAW << 8)) r0.y = (t0.y - dv.y) * 0.25f;
            if (raw & (NB_RAW << 16)) r1.x = (t1.x - dv.z) * 0.25f;
            if (raw & (NB_RAW << 24)) r1.y = (t1.y - dv.w) * 0.25f;
        }
        old[0] = nw[0]; old[1] = nw[1];
        nw[0] = r0; nw[1] = r1;
    }
    out[0] = nw[0]; out[1] = nw[1];
}

// Select-free packed row step in two phases.  Phase A forms (x1 + x2) + y1 of EVERY level from the rows
// kept from earlier iterations - nothing in it depends on this iteration's new rows, and after it the
// "older" slot of every level is dead.  Phase B is the short dependent chain through the levels
// (+ y2, then fma(., 0.25, -b4)): each level's new row can be written straight into the dead slot of the level
// below, so the state rotates without register moves.
template <int T, bool PZERO, int PAR>
__device__ __forceinline__ void row_step2_fast(float2 (&a)[T][2][2], const LaneAddr& sa, int i, int lane,
                                               float2 (&out)[2]) {
    constexpr unsigned FULL = 0xffffffffu;
    const int lane_l = (lane + 31) & 31, lane_r = (lane + 1) & 31;
    float2 part[T][2];
#pragma unroll
    for (int t = 1; t <= T; ++t) {
        const float2(&old)[2] = a[t - 1][PAR];
        const float2(&mid)[2] = a[t - 1][PAR ^ 1];
        const float sl = bitsel(mid[0].x, __shfl_sync(FULL, mid[1].y, lane_l), sa.edge_l);
        const float sr = bitsel(mid[1].y, __shfl_sync(FULL, mid[0].x, lane_r), sa.edge_r);
        const float2 s0 = make_float2(sl + mid[0].y, mid[0].x + mid[1].x);
        const float2 s1 = make_float2(mid[0].y + mid[1].y, mid[1].x + sr);
        part[t - 1][0] = __fadd2_rn(s0, old[0]);
        part[t - 1][1] = __fadd2_rn(s1, old[1]);
    }
    float2 nw0, nw1;
    if (PZERO) {
        nw0 = nw1 = make_float2(0.0f, 0.0f);
    } else {
        const float4 v = lds128(sa.p + (uint32_t)(i & (P_ROWS - 1)) * (SW * 4));
        nw0 = make_float2(v.x, v.y);
        nw1 = make_float2(v.z, v.w);
    }
#pragma unroll
    for (int t = 1; t <= T; ++t) {
        const float4 dv = lds128(sa.d + (uint32_t)((i - t) & (D_ROWS - 1)) * (SW * 4));
        const float2 r0 = finish2(__fadd2_rn(part[t - 1][0], nw0), make_float2(dv.x, dv.y));
        const float2 r1 = finish2(__fadd2_rn(part[t - 1][1], nw1), make_float2(dv.z, dv.w));
        a[t - 1][PAR][0] = nw0;
        a[t - 1][PAR][1] = nw1;
        nw0 = r0;
        nw1 = r1;
    }
    out[0] = nw0;
    out[1] = nw1;
}

// The first 2 T input rows of a tile only fill the pipeline: level t produces its first row that anything will
// ever read (row t of the tile's input window) at input row 2 t, so at input row i only levels 1 .. LV = i / 2 have
// work to do.  This is row_step2_fast cut off after level LV < T: the newest row of level LV is parked in that
// level's state instead of leaving the kernel.  (With every level running from row 0, as the other bodies do, the
// T (T + 1) level-steps skipped here produce values no later step consumes: 9 % of a 85-row tile at T = 8.)
template <int T, bool PZERO, int PAR, int LV>
__device__ __forceinline__ void row_step2_fill(float2 (&a)[T][2][2], const LaneAddr& sa, int i, int lane) {
    static_assert(LV < T, "the fill body stops below the last level");
    constexpr unsigned FULL = 0xffffffffu;
    const int lane_l = (lane + 31) & 31, lane_r = (lane + 1) & 31;
    float2 part[LV > 0 ? LV : 1][2];
#pragma unroll
    for (int t = 1; t <= LV; ++t) {
        const float2(&old)[2] = a[t - 1][PAR];
        const float2(&mid)[2] = a[t - 1][PAR ^ 1];
        const float sl = bitsel(mid[0].x, __shfl_sync(FULL, mid[1].y, lane_l), sa.edge_l);
        const float sr = bitsel(mid[1].y, __shfl_sync(FULL, mid[0].x, lane_r), sa.edge_r);
        const float2 s0 = make_float2(sl + mid[0].y, mid[0].x + mid[1].x);
        const float2 s1 = make_float2(mid[0].y + mid[1].y, mid[1].x + sr);
        part[t - 1][0] = __fadd2_rn(s0, old[0]);
        part[t - 1][1] = __fadd2_rn(s1, old[1]);
    }
    float2 nw0, nw1;
    if (PZERO) {
        nw0 = nw1 = make_float2(0.0f, 0.0f);
    } else {
        const float4 v = lds128(sa.p + (uint32_t)(i & (P_ROWS - 1)) * (SW * 4));
        nw0 = make_float2(v.x, v.y);
        nw1 = make_float2(v.z, v.w);
    }
#pragma unroll
    for (int t = 1; t <= LV; ++t) {
        const float4 dv = lds128(sa.d + (uint32_t)((i - t) & (D_ROWS - 1)) * (SW * 4));
        const float2 r0 = finish2(__fadd2_rn(part[t - 1][0], nw0), make_float2(dv.x, dv.y));
        const float2 r1 = finish2(__fadd2_rn(part[t - 1][1], nw1), make_float2(dv.z, dv.w));
        a[t - 1][PAR][0] = nw0;
        a[t - 1][PAR][1] = nw1;
        nw0 = r0;
        nw1 = r1;
    }
    a[LV][PAR][0] = nw0;
    a[LV][PAR][1] = nw1;
}

// Rows deep inside an obstacle: every cell of the strip has all four neighbours blocked, in every row in
// flight (and none carries NB_RAW), so each level is ((C + C) + C) + C - b, * 0.25 of the cell itself
// (shader.Poisson.comp:32-37 with all four substitutions): no neighbours, no shuffles, no selects.
template <int T, bool PZERO, int PAR>
__device__ __forceinline__ void row_step2_solid(float2 (&a)[T][2][2], const LaneAddr& sa, int i, float2 (&out)[2]) {
    float2 nw0, nw1;
    if (PZERO) {
        nw0 = nw1 = make_float2(0.0f, 0.0f);
    } else {
        const float4 v = lds128(sa.p + (uint32_t)(i & (P_ROWS - 1)) * (SW * 4));
        nw0 = make_float2(v.x, v.y);
        nw1 = make_float2(v.z, v.w);
    }
#pragma unroll
    for (int t = 1; t <= T; ++t) {
        const float2 c0 = a[t - 1][PAR ^ 1][0], c1 = a[t - 1][PAR ^ 1][1];
        const float4 dv = lds128(sa.d + (uint32_t)((i - t) & (D_ROWS - 1)) * (SW * 4));
        const float2 r0 = finish2(__fadd2_rn(__fadd2_rn(__fadd2_rn(c0, c0), c0), c0), make_float2(dv.x, dv.y));
        const float2 r1 = finish2(__fadd2_rn(__fadd2_rn(__fadd2_rn(c1, c1), c1), c1), make_float2(dv.z, dv.w));
        a[t - 1][PAR][0] = nw0;
        a[t - 1][PAR][1] = nw1;
        nw0 = r0;
        nw1 = r1;
    }
    out[0] = nw0;
    out[1] = nw1;
}

template <int T, bool PZERO, bool PACKED, int V>
__global__ void __launch_bounds__(Shape<V>::WARPS * 32, Shape<V>::BLOCKS)
k_jacobi_tb(const __grid_constant__ CUtensorMap map_p, const __grid_constant__ CUtensorMap map_d,
            const __grid_constant__ CUtensorMap map_m, const TBParams prm) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int WARPS = Shape<V>::WARPS;
    const int tile = blockIdx.x * WARPS + warp;
    if (tile >= prm.ntiles) return;                 // warps are independent: no block barrier below
    WarpSmem& S = reinterpret_cast<WarpSmem*>(smem_raw)[warp];

    const int4 td = prm.tiles[tile];
    const int strip = td.x, out_lo = td.y, out_hi = td.z;
    const int x0 = strip * (SW - 2 * prm.hx) - prm.hx;      // first strip column (may be < 0)
    const int y_first = out_lo - T;                          // first input row (local)
    const int nrows = (out_hi - out_lo) + 2 * T;
    const int ngroups = (nrows + GROUP - 1) / GROUP;   // extra rows of the last group are computed and dropped

    const uint32_t bar0 = smem_u32(&S.bar[0]);
    const uint32_t p_addr = smem_u32(&S.p[0][0]), d_addr = smem_u32(&S.d[0][0]), m_addr = smem_u32(&S.m[0][0]);
    if ((p_addr & 127u) != 0u) __trap();             // TMA destinations must be 128 B aligned
    if (lane == 0) {
#pragma unroll
        for (int b = 0; b < NBAR; ++b) mbar_init(bar0 + 8 * b, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();

    auto issue = [&](int g) {   // lane 0 only
        const uint32_t bar = bar0 + 8 * (g & (NBAR - 1));
        const int row = y_first + GROUP * g + prm.halo;
        constexpr uint32_t BYTES = GROUP * (SW * (PZERO ? 4u : 8u) + MBOX);
        mbar_expect_tx(bar, BYTES);
        if (!PZERO) tma_load_2d(p_addr + ((GROUP * g) & (P_ROWS - 1)) * (SW * 4), &map_p, x0, row, bar);
        tma_load_2d(d_addr + ((GROUP * g) & (D_ROWS - 1)) * (SW * 4), &map_d, x0, row, bar);
        tma_load_2d(m_addr + (g & (M_SLOTS - 1)) * M_SLOT, &map_m, x0 & ~15, row, bar);
    };
    // Programmatic dependent launch: this grid may become resident while the previous kernel of the stream (the
    // previous Jacobi launch, which wrote the pressure this one reads and read the buffer this one writes) is
    // still draining; everything above touched no field.  Wait for that kernel's completion and memory flush,
    // THEN let the next launch start its own prologue - so at most two launches are ever in flight and a launch
    // never overtakes the readers of its output buffer.  Both are no-ops in a plain launch.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    // per-tile trace (NATRIX_TB_TRACE): the clock starts once the tile may touch the fields
    if (prm.trace && lane == 0) { prm.trace[4 * tile] = globaltimer_ns(); prm.trace[4 * tile + 2] = smid(); }
    if (lane == 0) issue(0);

    float a[PACKED ? 1 : T][2][4];                    // scalar state
    float2 a2[PACKED ? T : 1][2][2];                  // packed state (two aligned pairs per row)
#pragma unroll
    for (int t = 0; t < T; ++t) {
        if constexpr (PACKED) {
            a2[t][0][0] = a2[t][0][1] = a2[t][1][0] = a2[t][1][1] = make_float2(0.0f, 0.0f);
        } else {
#pragma unroll
            for (int c = 0; c < 4; ++c) { a[t][0][c] = 0.0f; a[t][1][c] = 0.0f; }
        }
    }
    uint32_t busy = 0u;                               // bit t-1: row i-t has a non-zero mask somewhere in the warp
    constexpr uint32_t BUSY_MASK = (1u << T) - 1u;

    const int xa = x0 + 4 * lane;                     // this lane's first grid column
    const uint32_t edge_l = xa == 0 ? 0xffffffffu : 0u, edge_r = xa + 4 == prm.w ? 0xffffffffu : 0u;
    // the grid-edge L / R bits are honoured through edge_l / edge_r, so they need not make a row busy
    const LaneAddr sa{p_addr + 16u * lane, d_addr + 16u * lane, m_addr + (uint32_t)(x0 & 15) + 4u * lane, edge_l, edge_r,
                      0x1f1f1f1fu & ~((edge_l & (uint32_t)NB_L) | (edge_r & ((uint32_t)NB_R << 24)))};
    const bool st_ok = 4 * lane >= prm.hx && 4 * lane < SW - prm.hx && xa < prm.w;
    float* const out_col = prm.pout + xa;

    auto store_row = [&](int i, float r0, float r1, float r2, float r3) {
        const int ly = y_first + i - T;               // row of level T that just completed
        if (st_ok && ly >= out_lo && ly < out_hi)
            stg_stream(reinterpret_cast<float4*>(out_col + (ptrdiff_t)ly * prm.w), make_float4(r0, r1, r2, r3));
    };
    uint32_t solid = 0u;                              // bit t-1: every cell of row i-t (whole strip) is fully blocked
    uint32_t rawb = 0u;                               // bit t-1: some cell of row i-t carries NB_RAW
    auto fast_row = [&](auto par, int i) {            // no row in flight has mask bits
        constexpr int PAR = decltype(par)::value;
        if constexpr (PACKED) {
            float2 r[2];
            row_step2_fast<T, PZERO, PAR>(a2, sa, i, lane, r);
            store_row(i, r[0].x, r[0].y, r[1].x, r[1].y);
        } else {
            float r[4];
            row_step<T, PZERO, PAR, false>(a, sa, i, lane, r);
            store_row(i, r[0], r[1], r[2], r[3]);
        }
    };
    auto one_row = [&](auto par, int i) {             // picks the body from the history of the rows in flight
        constexpr int PAR = decltype(par)::value;
        if (!busy) {
            fast_row(par, i);
        } else if constexpr (PACKED) {
            float2 r[2];
            if (solid == BUSY_MASK) row_step2_solid<T, PZERO, PAR>(a2, sa, i, r);
            else if (rawb) row_step2<T, PZERO, PAR, true, true>(a2, sa, i, lane, r);
            else row_step2<T, PZERO, PAR, true>(a2, sa, i, lane, r);
            store_row(i, r[0].x, r[0].y, r[1].x, r[1].y);
        } else {
            float r[4];
            row_step<T, PZERO, PAR, true>(a, sa, i, lane, r);
            store_row(i, r[0], r[1], r[2], r[3]);
        }
    };
    using I0 = std::integral_constant<int, 0>;
    using I1 = std::integral_constant<int, 1>;
    // groups whose four rows all lie in the fill phase (input rows < 2 T): T / 2 of them
    constexpr int FILL_GROUPS = (PACKED && NATRIX_TB_FILL) ? T / 2 : 0;
    auto fill_rows = [&](auto gc, int i0) {           // the four rows of fill group G = decltype(gc)::value
        constexpr int G = decltype(gc)::value;
        if constexpr (PACKED && G < FILL_GROUPS) {
            row_step2_fill<T, PZERO, 0, (4 * G + 0) / 2>(a2, sa, i0, lane);
            row_step2_fill<T, PZERO, 1, (4 * G + 1) / 2>(a2, sa, i0 + 1, lane);
            row_step2_fill<T, PZERO, 0, (4 * G + 2) / 2>(a2, sa, i0 + 2, lane);
            row_step2_fill<T, PZERO, 1, (4 * G + 3) / 2>(a2, sa, i0 + 3, lane);
        }
    };
    auto fill_group = [&](int g, int i0) {
        switch (g) {
        case 0: fill_rows(std::integral_constant<int, 0>{}, i0); break;
        case 1: fill_rows(std::integral_constant<int, 1>{}, i0); break;
        case 2: fill_rows(std::integral_constant<int, 2>{}, i0); break;
        default: fill_rows(std::integral_constant<int, 3>{}, i0); break;
        }
    };

    // One TMA group (4 rows) per iteration: wait for it, put the next group in flight, consume it.
    // The select-free body runs when no row in flight has a mask bit anywhere in the warp.
    // No __syncwarp is needed before a ring slot is refilled: the slot's last readers are at least one
    // group back, and every row ends in full-mask shuffles / votes that all lanes must have reached.
    for (int g = 0; g < ngroups; ++g) {
        while (!mbar_try_wait(bar0 + 8 * (g & (NBAR - 1)), (g / NBAR) & 1)) {}
        const int i0 = GROUP * g;
        const uint32_t w0 = raw_mask_of_row(sa, i0), w1 = raw_mask_of_row(sa, i0 + 1);
        const uint32_t w2 = raw_mask_of_row(sa, i0 + 2), w3 = raw_mask_of_row(sa, i0 + 3);
        const uint32_t a0 = __any_sync(0xffffffffu, (w0 & sa.mkeep) != 0u) ? 1u : 0u, a1 = __any_sync(0xffffffffu, (w1 & sa.mkeep) != 0u) ? 1u : 0u;
        const uint32_t a2_ = __any_sync(0xffffffffu, (w2 & sa.mkeep) != 0u) ? 1u : 0u, a3 = __any_sync(0xffffffffu, (w3 & sa.mkeep) != 0u) ? 1u : 0u;
        if (lane == 0 && g + 1 < ngroups) issue(g + 1);
        if (busy | a0 | a1 | a2_ | a3) {
            // some row in flight (or arriving) carries mask bits: per row, the body its rows in flight need
            constexpr uint32_t ALL = 0x0f0f0f0fu, BITS = 0x1f1f1f1fu, RAWS = 0x01010101u * NB_RAW;
            // solid: all four neighbours blocked and no NB_RAW, in every cell of the strip
            const uint32_t s0 = __all_sync(0xffffffffu, (w0 & BITS) == ALL) ? 1u : 0u, s1 = __all_sync(0xffffffffu, (w1 & BITS) == ALL) ? 1u : 0u;
            const uint32_t s2 = __all_sync(0xffffffffu, (w2 & BITS) == ALL) ? 1u : 0u, s3 = __all_sync(0xffffffffu, (w3 & BITS) == ALL) ? 1u : 0u;
            const uint32_t q = __ballot_sync(0xffffffffu, ((w0 | w1 | w2 | w3) & RAWS) != 0u) ? 1u : 0u;   // per group is enough
            one_row(I0{}, i0);
            busy = ((busy << 1) | a0) & BUSY_MASK; solid = ((solid << 1) | s0) & BUSY_MASK; rawb = ((rawb << 1) | q) & BUSY_MASK;
            one_row(I1{}, i0 + 1);
            busy = ((busy << 1) | a1) & BUSY_MASK; solid = ((solid << 1) | s1) & BUSY_MASK; rawb = ((rawb << 1) | q) & BUSY_MASK;
            one_row(I0{}, i0 + 2);
            busy = ((busy << 1) | a2_) & BUSY_MASK; solid = ((solid << 1) | s2) & BUSY_MASK; rawb = ((rawb << 1) | q) & BUSY_MASK;
            one_row(I1{}, i0 + 3);
            busy = ((busy << 1) | a3) & BUSY_MASK; solid = ((solid << 1) | s3) & BUSY_MASK; rawb = ((rawb << 1) | q) & BUSY_MASK;
        } else if (PACKED && g < FILL_GROUPS) {
            // pipeline fill: input rows 4 g .. 4 g + 3 feed levels 1 .. (4 g + h) / 2 only
            solid = 0u;
            rawb = 0u;
            fill_group(g, i0);
        } else {
            solid = 0u;
            rawb = 0u;
#pragma unroll 1
            for (int h = 0; h < GROUP; h += 2) {
                fast_row(I0{}, i0 + h);
                fast_row(I1{}, i0 + h + 1);
            }
        }
    }
    if (prm.trace && lane == 0) prm.trace[4 * tile + 1] = globaltimer_ns();
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct MapEntry {
    const void* base;
    int w;
    size_t rows;
    int elem;
    CUtensorMap map;
};

using KernelFn = void (*)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const TBParams);

template <int T, int V>
KernelFn pick(bool pzero, bool packed) {
    if (packed) return pzero ? k_jacobi_tb<T, true, true, V> : k_jacobi_tb<T, false, true, V>;
    return pzero ? k_jacobi_tb<T, true, false, V> : k_jacobi_tb<T, false, false, V>;
}

template <int V>
KernelFn kernel_for_shape(int depth, bool pzero, bool packed) {
    switch (depth) {
    case 1: return pick<1, V>(pzero, packed);
    case 2: return pick<2, V>(pzero, packed);
    case 3: return pick<3, V>(pzero, packed);
    case 4: return pick<4, V>(pzero, packed);
    case 5: return pick<5, V>(pzero, packed);
    case 6: return pick<6, V>(pzero, packed);
    case 7: return pick<7, V>(pzero, packed);
    case 8: return pick<8, V>(pzero, packed);
    default: return nullptr;
    }
}

KernelFn kernel_for(int depth, bool pzero, bool packed, int shape) {
    switch (shape) {
    case 1: return kernel_for_shape<1>(depth, pzero, packed);
    case 2: return kernel_for_shape<2>(depth, pzero, packed);
    case 3: return kernel_for_shape<3>(depth, pzero, packed);
    default: return kernel_for_shape<0>(depth, pzero, packed);
    }
}

}  // namespace

struct JacobiTB {
    EncodeTiledFn encode = nullptr;
    std::string err;
    int sm_count = 148;
    int chunk_override = 0;
    std::vector<MapEntry> maps;
    // Tile plans: how the output cells are cut into one tile per resident warp.  Rows of a strip that
    // carry obstacles (from the boxes the host stamped this step) cost about `kappa` times a free row
    // (select body), so tiles are cut shorter there and every tile takes about the same time.
    struct Plan {
        std::vector<int> key;         // geometry + obstacle boxes
        std::vector<int4> tiles;
        int4* d_tiles = nullptr;
    };
    double sigma = 1.0;
    // measured (ncu, 32768x4096): 256 B promotion fetches 7 % more DRAM bytes than 128 B / none for the same time
    int l2_promotion = CU_TENSOR_MAP_L2_PROMOTION_L2_128B;      // NATRIX_TB_L2PROMO = 0 none, 1 64 B, 2 128 B, 3 256 B
    double kappa = 2.6;           // measured (this build, 200 sweeps at 32768x4096 with 64 circles / 100 at 4096^2): 2.0 -> 9.95 / 0.684 ms, 2.3 -> 9.91 / 0.661, 2.6 -> 9.65 / 0.662, 3.0 -> 9.56 / 0.682

    static constexpr int PU = 4;  // planning granularity (rows) = one TMA group

    // boxes: nboxes x (x0, x1, y0, y1), global columns, local rows, half-open; or a circle (cx, -1 - r, cy, 0).
    static std::vector<int4> cut_tiles(int w, int r0, int r1, int hx, int depth, const int* boxes, int nboxes,
                                       int max_tiles, double kappa, int chunk_override, double sigma = 1.0,
                                       int warps_per_block = 0) {
        const int pitch = SW - 2 * hx, nstrips = (w + pitch - 1) / pitch;
        const int rows = r1 - r0, nu = (rows + PU - 1) / PU;
        // per strip, prefix counts of planning units that will run the select body ("heavy": some cell
        // of the strip has a blocked neighbour) or the neighbour-free body ("solid": the strip is wholly
        // inside a circle there, with `depth` rows of margin for the rows in flight)
        std::vector<int> pre((size_t)nstrips * (nu + 1), 0), pres((size_t)nstrips * (nu + 1), 0);
        std::vector<char> marks((size_t)nstrips * nu, 0);
        // boxes outside, strips inside: a box only visits the strips whose columns it can reach (a conservative
        // range; the exact test follows), so the cost is per box-strip overlap, not boxes x strips
        for (int pass = 0; pass < 2; ++pass) {               // pass 0: heavy rows, pass 1: solid rows override
            for (int k = 0; k < nboxes; ++k) {
                const int* bx = boxes + 4 * k;
                const bool circle = bx[1] < bx[0];
                if (!circle && pass == 1) continue;
                const double xlo = circle ? (double)bx[0] - (double)(-1 - bx[1]) - 2.0 : (double)bx[0];
                const double xhi = circle ? (double)bx[0] + (double)(-1 - bx[1]) + 2.0 : (double)bx[1];
                const int s_lo = std::max(0, (int)std::floor((xlo + hx - SW - 1) / pitch) - 1);
                const int s_hi = std::min(nstrips - 1, (int)std::ceil((xhi + hx + 1) / pitch) + 1);
                for (int st_ = s_lo; st_ <= s_hi; ++st_) {
                    const int c0 = st_ * pitch - hx - 1, c1 = st_ * pitch - hx + SW + 1;   // columns whose masks the strip reads
                    char* mark = &marks[(size_t)st_ * nu];
                    int ya, yb;
                    if (circle) {
                        // circle (cx, -1 - r, cy, -): only the rows where it really crosses this strip's columns
                        const double cx = bx[0], r = (double)(-1 - bx[1]), cy = bx[2];
                        if (pass == 0) {
                            const double rr = r + 2.0;
                            const double dx = std::max(0.0, std::max((double)c0 - cx, cx - (double)(c1 - 1)));
                            if (dx > rr) continue;
                            const double hh = std::sqrt(rr * rr - dx * dx);
                            ya = std::max((int)std::floor(cy - hh) - 1, r0);
                            yb = std::min((int)std::ceil(cy + hh) + 2 + depth, r1);   // + the rows it stays in flight
                        } else {
                            const double rr = r - 2.0;
                            const double dmax = std::max(std::fabs((double)c0 - cx), std::fabs((double)(c1 - 1) - cx));
                            if (c0 < 0 || c1 > w || dmax >= rr) continue;
                            const double hf = std::sqrt(rr * rr - dmax * dmax) - depth - 1;
                            ya = std::max((int)std::ceil(cy - hf), r0);
                            yb = std::min((int)std::floor(cy + hf), r1);
                        }
                    } else {
                        if (bx[1] <= c0 || bx[0] >= c1) continue;
                        ya = std::max(bx[2] - 1, r0);
                        yb = std::min(bx[3] + 1 + depth, r1);
                    }
                    if (yb <= ya) continue;
                    if (pass == 0) {
                        for (int u = (ya - r0) / PU; u < nu && r0 + u * PU < yb; ++u) mark[u] = 1;
                    } else {                                 // only units lying wholly inside
                        for (int u = (ya - r0 + PU - 1) / PU; u < nu && r0 + (u + 1) * PU <= yb; ++u) mark[u] = 2;
                    }
                }
            }
        }
        for (int st_ = 0; st_ < nstrips; ++st_) {
            const char* mark = &marks[(size_t)st_ * nu];
            int* p = &pre[(size_t)st_ * (nu + 1)];
            int* q = &pres[(size_t)st_ * (nu + 1)];
            for (int u = 0; u < nu; ++u) { p[u + 1] = p[u] + (mark[u] == 1); q[u + 1] = q[u] + (mark[u] == 2); }
        }
        // cost of a neighbour-free ("solid") row relative to a free row: measured equal on B200 (per-tile trace,
        // 4096^2 with one r = 256 circle: 112-row solid tiles 0.54 us/row, free tiles 0.545 us/row)
        const double SIGMA = sigma;
        // Cost of planning units [ua, ub) of one strip, in free-row equivalents.  Fitted on per-tile traces
        // (NATRIX_TB_TRACE, B200): free 0.42-0.44 us/row, select body 0.96-1.12 us/row (kappa), neighbour-free body
        // as a free row (sigma 1.0), and a constant of ~23 rows per tile: the 2 * depth warm-up rows plus start-up.
        const double warm = 2.0 * depth + 6.0;
        auto cost = [&](int st_, int ua, int ub) {
            const int* p = &pre[(size_t)st_ * (nu + 1)];
            const int* q = &pres[(size_t)st_ * (nu + 1)];
            const int hv = p[ub] - p[ua], so = q[ub] - q[ua];
            return (double)(ub - ua) * PU + (kappa - 1.0) * hv * PU + (SIGMA - 1.0) * so * PU + warm * (hv > 0 ? kappa : 1.0);
        };
        std::vector<int4> tiles;
        if (chunk_override > 0) {
            for (int y = r0; y < r1; y += chunk_override)
                for (int st_ = 0; st_ < nstrips; ++st_) tiles.push_back(make_int4(st_, y, std::min(r1, y + chunk_override), 0));
            return tiles;
        }
        // greedy cut of one strip under a cost limit: unit boundaries (ends) of its tiles
        // (the planner runs on the host between the pre-projection launch and the first Jacobi launch; with moving
        // obstacles it runs every step, so its cost matters: the search for a tile's end gallops out from the previous
        // tile's length - tiles of a strip are of similar length - instead of bisecting the whole strip)
        auto cut_strip = [&](int st_, double limit, std::vector<int>* ends) {
            int n = 0, ua = 0, len = 0;
            while (ua < nu) {
                // largest ub in (ua, nu] with cost(ua, ub) <= limit (at least one unit); cost is monotone in ub.
                // invariant: cost(ua, good) <= limit (or good == ua + 1), cost(ua, bad) > limit (or bad == nu + 1)
                auto fits = [&](int ub) { return cost(st_, ua, ub) <= limit; };
                int good = ua + 1, bad = nu + 1;
                const int probe = len > 0 ? std::min(nu, ua + len) : nu;
                int step = std::max(1, len / 8);
                if (probe > good) {
                    if (fits(probe)) {
                        good = probe;
                        for (;;) {
                            const int nx = std::min(nu, good + step);
                            if (nx == good) break;
                            if (fits(nx)) { good = nx; step *= 2; } else { bad = nx; break; }
                        }
                    } else {
                        bad = probe;
                        for (;;) {
                            const int nx = std::max(ua + 1, bad - step);
                            if (nx <= good) break;
                            if (fits(nx)) { good = nx; break; }
                            bad = nx;
                            step *= 2;
                        }
                    }
                }
                while (bad - good > 1) {
                    const int mid = (good + bad) / 2;
                    if (fits(mid)) good = mid; else bad = mid;
                }
                if (ends) ends->push_back(good);
                len = good - ua;
                ua = good;
                ++n;
            }
            return n;
        };
        auto count_all = [&](double limit) {
            int n = 0;
            for (int st_ = 0; st_ < nstrips; ++st_) n += cut_strip(st_, limit, nullptr);
            return n;
        };
        // 1. the smallest common limit that needs no more tiles than there are resident warps
        //    (bisection down to half a row; it starts from the perfect-packing bound, which no cut can beat)
        double lo = 0.0, hi = 0.0, sum = 0.0;
        for (int st_ = 0; st_ < nstrips; ++st_) { const double c = cost(st_, 0, nu); hi = std::max(hi, c); sum += c; }
        lo = std::min(hi, sum / max_tiles);
        for (int it = 0; it < 24 && hi - lo > 0.5; ++it) {
            const double mid = 0.5 * (lo + hi);
            if (count_all(mid) <= max_tiles) hi = mid; else lo = mid;
        }
        // 2. per strip: with its number of tiles fixed, the smallest limit that still fits (evens the strip's tiles
        //    out instead of leaving a short remainder); then hand the warps still unused, one at a time, to the strip
        //    whose costliest tile is the largest.
        std::vector<int> k(nstrips);
        std::vector<double> lim(nstrips);
        auto tighten = [&](int st_) {                       // smallest limit that cuts strip st_ into <= k[st_] tiles
            // k tiles cannot all cost less than the strip's cost shared out evenly (less the one warm-up it counts)
            double a = std::max(0.0, (cost(st_, 0, nu) - warm * kappa) / k[st_]), b = lim[st_];
            for (int it = 0; it < 20 && b - a > 0.5; ++it) {
                const double mid = 0.5 * (a + b);
                if (cut_strip(st_, mid, nullptr) <= k[st_]) b = mid; else a = mid;
            }
            lim[st_] = b;
        };
        int total = 0;
        for (int st_ = 0; st_ < nstrips; ++st_) {
            k[st_] = cut_strip(st_, hi, nullptr);
            lim[st_] = hi;
            total += k[st_];
            tighten(st_);
        }
        while (total < max_tiles) {
            int worst = 0;
            for (int st_ = 1; st_ < nstrips; ++st_)
                if (lim[st_] > lim[worst]) worst = st_;
            if (k[worst] >= nu) break;                       // already one unit per tile
            const double before = lim[worst];
            ++k[worst];
            ++total;
            tighten(worst);
            if (lim[worst] >= before) break;                 // a single heavy unit bounds the span: nothing left to gain
        }
        for (int st_ = 0; st_ < nstrips; ++st_) {
            std::vector<int> ends;
            cut_strip(st_, lim[st_], &ends);
            const int* p = &pre[(size_t)st_ * (nu + 1)];
            const int* q = &pres[(size_t)st_ * (nu + 1)];
            int ua = 0;
            for (int ub : ends) {
                // .w = heavy units | solid units << 16 (diagnostics: NATRIX_TB_TRACE, natrix_debug_plan_tiles)
                tiles.push_back(make_int4(st_, r0 + ua * PU, std::min(r1, r0 + ub * PU), (p[ub] - p[ua]) | ((q[ub] - q[ua]) << 16)));
                ua = ub;
            }
        }
        // 3. order = placement: tile i runs on warp i % warps of block i / warps.  Neighbouring warps should stream
        //    neighbouring strips of the same rows, and the warps of one SM share its issue slots, so the tiles with
        //    select-body rows (2 - 3x the instructions) are dealt out over the blocks instead of sitting together
        //    (measured: the slowest tiles were heavy ones sharing an SM with other heavy ones, +12 % over the model).
        std::stable_sort(tiles.begin(), tiles.end(), [](const int4& a, const int4& b) { return a.y < b.y; });
        if (warps_per_block > 0 && (int)tiles.size() > warps_per_block) {
            const int n = (int)tiles.size(), nblocks = (n + warps_per_block - 1) / warps_per_block;
            std::vector<int> order(n);
            for (int i = 0; i < n; ++i) order[i] = i;
            std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return (tiles[a].w & 0xffff) > (tiles[b].w & 0xffff); });
            int nheavy = 0;
            while (nheavy < n && (tiles[order[nheavy]].w & 0xffff) != 0) ++nheavy;
            // heavy tiles go round-robin over the blocks (heaviest first); the free ones fill the remaining slots in
            // row order.  Slots of the last, partial block are filled last.
            std::vector<std::vector<int>> blocks(nblocks);
            for (int i = 0; i < nheavy; ++i) blocks[i % nblocks].push_back(order[i]);
            std::vector<int> rest(order.begin() + nheavy, order.end());
            std::sort(rest.begin(), rest.end());
            size_t r = 0;
            for (int b = 0; b < nblocks; ++b) {
                const int cap = std::min(warps_per_block, n - b * warps_per_block);
                while ((int)blocks[b].size() < cap && r < rest.size()) blocks[b].push_back(rest[r++]);
            }
            std::vector<int4> placed;
            placed.reserve(n);
            bool ok = r == rest.size();
            for (int b = 0; b < nblocks && ok; ++b) {
                const int cap = std::min(warps_per_block, n - b * warps_per_block);
                if ((int)blocks[b].size() != cap) ok = false;
                for (int i : blocks[b]) placed.push_back(tiles[i]);
            }
            if (ok) tiles.swap(placed);                      // (more heavy tiles than fit evenly: keep the row order)
        }
        return tiles;
    }

    // Plan cache without a cliff.  Per launching stream, SLOTS fixed-size slots in one device arena with a pinned
    // host mirror, allocated on first use.  A miss (a new obstacle set - every step when obstacles move) costs the
    // host-side cut and an asynchronous upload on the launching stream: no cudaMalloc, no stream synchronisation
    // and no cudaFree on the step path.  A slot is only ever read by launches on its own stream, so overwriting
    // the least recently used one is ordered behind its readers by the stream itself; the pinned mirror of the
    // slot is guarded by an event recorded behind its previous upload (long complete unless the host has run more
    // than SLOTS misses ahead of the GPU).
    static constexpr int SLOTS = 32;
    struct Pool {
        cudaStream_t stream = nullptr;
        int slot_cap = 0;                          // tiles per slot
        int4* d_arena = nullptr;
        int4* h_arena = nullptr;
        std::vector<Plan> plans;
        cudaEvent_t uploaded[SLOTS] = {};
        unsigned long long used[SLOTS] = {};
    };
    std::vector<Pool*> pools;
    unsigned long long use_clock = 0, plan_hits = 0, plan_misses = 0;

    const Plan* plan_for(int w, int r0, int r1, int hx, int depth, const int* boxes, int nboxes, int max_tiles,
                         int warps_per_block, cudaStream_t st) {
        std::vector<int> key = {w, r0, r1, hx, depth, max_tiles, chunk_override, warps_per_block};
        key.insert(key.end(), boxes, boxes + 4 * nboxes);
        Pool* pool = nullptr;
        for (Pool* q : pools)
            if (q->stream == st) pool = q;
        if (!pool) {
            pool = new Pool();
            pool->stream = st;
            pool->plans.resize(SLOTS);
            pools.push_back(pool);
        }
        for (int i = 0; i < SLOTS; ++i) {
            Plan& p = pool->plans[i];
            if (p.d_tiles && p.key == key) {
                pool->used[i] = ++use_clock;
                ++plan_hits;
                return &p;
            }
        }
        ++plan_misses;
        std::vector<int4> tiles = cut_tiles(w, r0, r1, hx, depth, boxes, nboxes, max_tiles, kappa, chunk_override, sigma, warps_per_block);
        const int pitch = SW - 2 * hx, nstrips = (w + pitch - 1) / pitch;
        const int need = std::max((int)tiles.size(), sm_count * 32 + 2 * nstrips);
        if (need > pool->slot_cap) {
            // (re)allocate the arena: first use, or more tiles than any plan before (not the steady state)
            cudaStreamSynchronize(st);
            cudaFree(pool->d_arena);
            if (pool->h_arena) cudaFreeHost(pool->h_arena);
            pool->d_arena = nullptr; pool->h_arena = nullptr;
            pool->slot_cap = need + need / 4;
            if (cudaMalloc((void**)&pool->d_arena, (size_t)SLOTS * pool->slot_cap * sizeof(int4)) != cudaSuccess ||
                cudaMallocHost((void**)&pool->h_arena, (size_t)SLOTS * pool->slot_cap * sizeof(int4)) != cudaSuccess) {
                err = "tile plan arena allocation failed";
                return nullptr;
            }
            for (int i = 0; i < SLOTS; ++i) {
                pool->plans[i] = Plan();
                pool->used[i] = 0;
                if (!pool->uploaded[i] && cudaEventCreateWithFlags(&pool->uploaded[i], cudaEventDisableTiming) != cudaSuccess) {
                    err = "cudaEventCreate(tile plan)";
                    return nullptr;
                }
            }
        }
        int victim = 0;
        for (int i = 1; i < SLOTS; ++i)
            if (pool->used[i] < pool->used[victim]) victim = i;
        Plan& p = pool->plans[victim];
        if (p.d_tiles && cudaEventSynchronize(pool->uploaded[victim]) != cudaSuccess) { err = "cudaEventSynchronize(tile plan)"; return nullptr; }
        p.key = std::move(key);
        p.tiles = std::move(tiles);
        p.d_tiles = pool->d_arena + (size_t)victim * pool->slot_cap;
        int4* host = pool->h_arena + (size_t)victim * pool->slot_cap;
        std::copy(p.tiles.begin(), p.tiles.end(), host);
        if (cudaMemcpyAsync(p.d_tiles, host, p.tiles.size() * sizeof(int4), cudaMemcpyHostToDevice, st) != cudaSuccess ||
            cudaEventRecord(pool->uploaded[victim], st) != cudaSuccess) { err = "tile plan upload failed"; return nullptr; }
        pool->used[victim] = ++use_clock;
        return &p;
    }
    // developer tracing: NATRIX_TB_TRACE=<file> dumps per-tile (strip, rows, start, end, SM) of launch number
    // NATRIX_TB_TRACE_LAUNCH (default 40) as CSV; costs a device sync on that launch only
    std::string trace_path;
    int trace_launch = 40, launch_no = 0;
    unsigned long long* d_trace = nullptr;
    // programmatic dependent launch of consecutive Jacobi launches (NATRIX_TB_PDL=0 turns it off): the next
    // launch's blocks take an SM as soon as this launch's block there has exited and wait in their prologue
    int pdl = 1;
    int reserve_sms = 0;          // SMs the next launches leave free (for an exchange kernel running beside them)
    bool attr_set[JACOBI_TB_MAX_DEPTH + 1][2][2][NUM_SHAPES] = {};
    int shape = 1;                // measured on B200 at 4096^2: 12 warps x 168 registers beats 2 x 8 warps x 128

    const CUtensorMap* map_for(const void* base, int w, size_t rows, int elem) {
        for (const MapEntry& e : maps)
            if (e.base == base && e.w == w && e.rows == rows && e.elem == elem) return &e.map;
        const int box_w = elem == 4 ? SW : MBOX;
        if (!encode) {
            void* fn = nullptr;
            cudaDriverEntryPointQueryResult q;
            cudaError_t ce = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
            if (ce != cudaSuccess || q != cudaDriverEntryPointSuccess || !fn) {
                err = "cuTensorMapEncodeTiled is not available from the CUDA driver";
                return nullptr;
            }
            encode = (EncodeTiledFn)fn;
        }
        MapEntry e{base, w, rows, elem, {}};
        const cuuint64_t gdim[2] = {(cuuint64_t)w, (cuuint64_t)rows};
        const cuuint64_t gstride[1] = {(cuuint64_t)w * (cuuint64_t)elem};
        const cuuint32_t box[2] = {(cuuint32_t)box_w, (cuuint32_t)GROUP};
        const cuuint32_t estride[2] = {1, 1};
        CUresult r = encode(&e.map, elem == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_UINT8, 2,
                            const_cast<void*>(base), gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_NONE, (CUtensorMapL2promotion)l2_promotion,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            err = "cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r);
            return nullptr;
        }
        if (maps.size() >= 64) maps.erase(maps.begin());
        maps.push_back(e);
        return &maps.back().map;
    }
};

JacobiTB* jacobi_tb_create() {
    JacobiTB* tb = new JacobiTB();
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) tb->sm_count = n;
    }
    if (const char* e = getenv("NATRIX_TB_CHUNK")) tb->chunk_override = atoi(e);
    if (const char* e = getenv("NATRIX_TB_SHAPE")) tb->shape = atoi(e) % NUM_SHAPES;
    if (const char* e = getenv("NATRIX_TB_KAPPA")) tb->kappa = atof(e);
    if (const char* e = getenv("NATRIX_TB_SIGMA")) tb->sigma = atof(e);
    if (const char* e = getenv("NATRIX_TB_L2PROMO")) tb->l2_promotion = atoi(e) & 3;
    if (const char* e = getenv("NATRIX_TB_PDL")) tb->pdl = atoi(e) != 0;
    if (const char* e = getenv("NATRIX_TB_TRACE")) tb->trace_path = e;
    if (const char* e = getenv("NATRIX_TB_TRACE_LAUNCH")) tb->trace_launch = atoi(e);
    return tb;
}

void jacobi_tb_destroy(JacobiTB* tb) {
    if (!tb) return;
    for (JacobiTB::Pool* q : tb->pools) {
        cudaFree(q->d_arena);
        if (q->h_arena) cudaFreeHost(q->h_arena);
        for (cudaEvent_t e : q->uploaded) if (e) cudaEventDestroy(e);
        delete q;
    }
    delete tb;
}
void jacobi_tb_plan_stats(JacobiTB* tb, unsigned long long* hits, unsigned long long* misses) {
    *hits = tb ? tb->plan_hits : 0;
    *misses = tb ? tb->plan_misses : 0;
}
void jacobi_tb_reserve_sms(JacobiTB* tb, int n) { if (tb) tb->reserve_sms = n < 0 ? 0 : n; }

const char* jacobi_tb_error(JacobiTB* tb) { return tb ? tb->err.c_str() : "null JacobiTB"; }

int jacobi_tb_plan_debug(int w, int depth, int r0, int r1, const int* boxes, int nboxes, int max_tiles, int* out4,
                         int cap) {
    const int hx = depth <= 4 ? 4 : 8;
    const std::vector<int4> tiles = JacobiTB::cut_tiles(w, r0, r1, hx, depth, boxes, nboxes, max_tiles, 2.3, 0, 1.0, 12);
    for (size_t i = 0; i < tiles.size() && (int)i < cap; ++i) {
        out4[4 * i] = tiles[i].x; out4[4 * i + 1] = tiles[i].y; out4[4 * i + 2] = tiles[i].z; out4[4 * i + 3] = tiles[i].w;
    }
    return (int)tiles.size();
}

bool jacobi_tb_supported(const Geom& g) {
    // TMA needs 16-byte row pitches for the float fields and the byte mask
    return g.w % 16 == 0 && g.w >= SW;
}

int jacobi_tb_launch(JacobiTB* tb, const float* pin, const float* div4, const uint8_t* nbmask, float* pout, Geom g,
                     int depth, int r0, int r1, bool p_is_zero, int packed, const int* boxes, int nboxes,
                     cudaStream_t st) {
    if (!tb) return -1;
    if (depth < 1 || depth > JACOBI_TB_MAX_DEPTH) { tb->err = "depth out of range"; return -1; }
    if (!jacobi_tb_supported(g)) { tb->err = "grid width must be a multiple of 16 and >= 256"; return -1; }
    if (r1 <= r0) return 0;
    const size_t rows_alloc = (size_t)g.hl + 2 * (size_t)g.halo;
    const ptrdiff_t off = (ptrdiff_t)g.halo * g.w;
    const CUtensorMap* mp = tb->map_for(pin - off, g.w, rows_alloc, 4);
    if (!mp) return -1;
    const CUtensorMap map_p = *mp;            // copy: the cache vector may reallocate
    const CUtensorMap* md = tb->map_for(div4 - off, g.w, rows_alloc, 4);
    if (!md) return -1;
    const CUtensorMap map_d = *md;
    const CUtensorMap* mm = tb->map_for(nbmask - off, g.w, rows_alloc, 1);
    if (!mm) return -1;
    const CUtensorMap map_m = *mm;

    const int warps = SHAPE_WARPS[tb->shape];
    const size_t smem = (size_t)warps * sizeof(WarpSmem);
    TBParams prm;
    prm.pout = pout;
    prm.w = g.w;
    prm.halo = g.halo;
    prm.r0 = r0;
    prm.r1 = r1;
    prm.hx = depth <= 4 ? 4 : 8;
    // one tile per resident warp; tile heights follow the obstacle boxes (see Plan)
    const int max_tiles = std::max(1, tb->sm_count - tb->reserve_sms) * warps * (tb->shape == 0 ? 2 : 1);
    // The grid's first and last row carry the B / T blocked bits (clamp-to-edge), so the rows next to them run
    // the select body while they are in flight: tell the planner (measured: the top band of tiles took 52.5 us
    // against 48 us for every other band at 4096^2).
    std::vector<int> hints(boxes, boxes + 4 * (size_t)nboxes);
    for (int edge : {-g.y0, g.hg - 1 - g.y0})
        if (edge >= r0 - depth && edge < r1 + depth) { hints.insert(hints.end(), {0, g.w, edge, edge + 1}); }
    boxes = hints.data();
    nboxes = (int)hints.size() / 4;
    const JacobiTB::Plan* plan = tb->plan_for(g.w, r0, r1, prm.hx, depth, boxes, nboxes, max_tiles, warps, st);
    if (!plan) return -1;
    prm.tiles = plan->d_tiles;
    prm.ntiles = (int)plan->tiles.size();
    const int blocks = (prm.ntiles + warps - 1) / warps;

    KernelFn fn = kernel_for(depth, p_is_zero, packed != 0, tb->shape);
    bool& attr_done = tb->attr_set[depth][p_is_zero ? 1 : 0][packed ? 1 : 0][tb->shape];
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute((const void*)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { tb->err = std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e); return -1; }
        attr_done = true;
    }
    prm.trace = nullptr;
    const bool tracing = !tb->trace_path.empty() && tb->launch_no == tb->trace_launch;
    ++tb->launch_no;
    if (tracing && cudaMalloc((void**)&tb->d_trace, (size_t)prm.ntiles * 32) == cudaSuccess) {
        cudaMemsetAsync(tb->d_trace, 0, (size_t)prm.ntiles * 32, st);
        prm.trace = tb->d_trace;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)blocks);
    cfg.blockDim = dim3((unsigned)(warps * 32));
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (tb->pdl && !tracing) ? 1 : 0;
    cudaError_t e = cudaLaunchKernelEx(&cfg, fn, map_p, map_d, map_m, prm);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) { tb->err = std::string("k_jacobi_tb launch: ") + cudaGetErrorString(e); return -1; }
    if (tracing && prm.trace) {
        std::vector<unsigned long long> h((size_t)prm.ntiles * 4);
        cudaStreamSynchronize(st);
        cudaMemcpy(h.data(), tb->d_trace, h.size() * 8, cudaMemcpyDeviceToHost);
        cudaFree(tb->d_trace);
        tb->d_trace = nullptr;
        if (FILE* f = fopen(tb->trace_path.c_str(), "w")) {
            fprintf(f, "tile,strip,row0,row1,start_ns,end_ns,smid,depth,w,r0,r1,heavy_rows,solid_rows\n");
            for (int i = 0; i < prm.ntiles; ++i)
                fprintf(f, "%d,%d,%d,%d,%llu,%llu,%llu,%d,%d,%d,%d,%d,%d\n", i, plan->tiles[i].x, plan->tiles[i].y, plan->tiles[i].z,
                        h[4 * i], h[4 * i + 1], h[4 * i + 2], depth, g.w, r0, r1, (plan->tiles[i].w & 0xffff) * JacobiTB::PU,
                        (plan->tiles[i].w >> 16) * JacobiTB::PU);
            fclose(f);
        }
    }
    return 1;
}

}  // namespace natrix
