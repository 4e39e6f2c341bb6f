// fused.cu - the fused pipeline's non-Jacobi kernels (NATRIX_OPT_PIPELINE = 1).
//
//  * k_preproject: ONE pass for everything the reference does between the start of update() and the
//    Poisson loop - [InitBoundaries] -> AdvectVelocity -> CalcVorticity -> ApplyVorticity ->
//    [Viscosity] -> Divergence (+ the blocked-neighbour mask).  Reads velocity 8 B + obstacles 1 B per
//    cell, writes velocity 8 B, vorticity 4 B, divergence 4 B, scaled divergence 4 B, mask 1 B = 30 B/cell, instead of
//    92 B/cell (108 with viscosity) for the five dispatches.
//  * impulse kernels that touch only the bounding boxes of the splats, in place;
//  * the gradient subtraction driven by the mask.
//
// All of them produce bit-identical results to the one-kernel-per-shader versions in
// stages_ref.cu (same expressions, same operand order, -fmad=false).
#include <algorithm>
#include <type_traits>

#include "kernels.h"

namespace natrix {
namespace {

// ------------------------------------------------------------------------------------ pre-projection
//
// Streaming decomposition (same idea as jacobi_tb.cu): one warp owns a strip of 128 columns (4 per
// lane, 4 halo columns on each side recomputed) and marches down its chunk of rows.  When the
// advected row `ly` is produced, the vorticity of row ly-1, the confined velocity of row ly-2,
// [the viscous velocity of row ly-3] and the divergence of the row above that follow from rolling
// 3-row windows held in registers; left/right neighbours come from the adjacent lanes by shuffle.
constexpr int PSW = 128;      // strip width
constexpr int PHX = 4;        // halo columns per side = number of x-neighbour stages

struct PreParams {
    int r0, r1;               // output rows (local)
    int ch, nstrips, ntiles;
    float dt, speed, diss, scale, alpha, rbeta;
};

__device__ __forceinline__ void copy4(float2 (&d)[4], const float2 (&s)[4]) {
#pragma unroll
    for (int j = 0; j < 4; ++j) d[j] = s[j];
}
__device__ __forceinline__ void copy4(float (&d)[4], const float (&s)[4]) {
#pragma unroll
    for (int j = 0; j < 4; ++j) d[j] = s[j];
}

// Clamped floor / ceil corners of a back-traced position with the UNclamped delta (ref:
// shader.AdvectVelocity.comp:38-42), in integers: int(clamp(floor(f), 0, m)) == clamp(int_floor(f), 0, m)
// for every float f - cvt.rmi / cvt.rpi saturate where the float clamp would have cut anyway, and NaN gives 0
// both ways - so this is common.cuh corners() with 4 conversions and 4 integer clamps (VIMNMX.RELU) instead of
// 4 roundings, 8 float min / max and 4 conversions.
__device__ __forceinline__ Corners corners_int(float fx, float fy, int w, int h) {
    Corners c;
    c.bx = __vimin_s32_relu(__float2int_rd(fx), w - 1);      // max(min(v, w - 1), 0) in one instruction
    c.tx = __vimin_s32_relu(__float2int_ru(fx), w - 1);
    c.by = __vimin_s32_relu(__float2int_rd(fy), h - 1);
    c.ty = __vimin_s32_relu(__float2int_ru(fy), h - 1);
    c.dx = fx - (float)c.bx;
    c.dy = fy - (float)c.by;
    return c;
}

// The bilinear back-trace of one cell (ref: shader.AdvectVelocity.comp:36-49; same expressions as common.cuh
// advect_cell) in two halves, so that the four gathers of a cell can be in flight across a whole row of
// arithmetic: issue = corners, addresses, loads; finish = the three mixes.  SLAB adds the check that the gathered
// rows are rows this slab holds.
struct Gather { float2 lt, rt, lb, rb; float dx, dy; };
template <bool SLAB>
__device__ __forceinline__ void gather_issue(Gather& q, const float2* __restrict__ rowp, const Geom& g, float fxc, int gy,
                                             int gyw, float fyc, float2 vel, float dt, float speed, int* __restrict__ err) {
    const float fx = fxc - vel.x * dt * speed;          // fxc = (float)x, fyc = (float)gy
    const float fy = fyc - vel.y * dt * speed;
    Corners c = corners_int(fx, fy, g.w, g.hg);
    if (SLAB) {
        const int lo = g.y0 - g.halo, hi = g.y0 + g.hl + g.halo - 1;
        if (c.by < lo || c.ty > hi) *err = 1;
        c.by = clampi(c.by, lo, hi);
        c.ty = clampi(c.ty, lo, hi);
    }
    // 32-bit cell offsets from the first cell of the row the warp is on (rowp; gyw = gy * width): one multiply-add
    // per gathered row, one add and one 64-bit multiply-add per load (launch_preproject checks the grid has < 2^31 cells)
    const int ob = c.by * g.w - gyw, ot = c.ty * g.w - gyw;
    q.lt = __ldg(rowp + (ot + c.bx)); q.rt = __ldg(rowp + (ot + c.tx));      // ld.global.nc, as the compiler
    q.lb = __ldg(rowp + (ob + c.bx)); q.rb = __ldg(rowp + (ob + c.tx));      // infers for a plain __restrict__ read
    q.dx = c.dx; q.dy = c.dy;
}
// (The three mixes are scalar on purpose: written with the 2-wide fp32 intrinsics, ptxas 12.9 contracts
// mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even under --fmad=false, which rounds once where the shader rounds twice.)
__device__ __forceinline__ float2 gather_finish(const Gather& q, float diss) {
    const float h1x = mixf(q.lt.x, q.rt.x, q.dx), h1y = mixf(q.lt.y, q.rt.y, q.dx);
    const float h2x = mixf(q.lb.x, q.rb.x, q.dx), h2y = mixf(q.lb.y, q.rb.y, q.dx);
    float2 o;
    o.x = clampf(mixf(h2x, h1x, q.dy) * diss, -1.0f, 1.0f);
    o.y = clampf(mixf(h2y, h1y, q.dy) * diss, -1.0f, 1.0f);
    return o;
}

// InitBoundaries (shader.InitBoundaries.comp:14-34) is NOT folded in here: when has_borders is set the
// border lines of the READ buffer are zeroed in place by k_zero_borders first, exactly like the
// reference's dispatch does (2 (W + H) cells; cheaper than testing every gathered corner).
// Both variants run up to 12 warps per block, one block per SM (up to 168 registers); small grids get smaller blocks.
// PIPE = false: the gathers of a row are consumed in the same iteration (small grids: shorter warm-up per tile).
// PIPE = true : the gathers of row ly+1 are issued before the vorticity / confinement / divergence arithmetic
//               of row ly and consumed one iteration later; the rolling windows rotate without register moves.
template <bool VISCOUS, bool SLAB, bool PIPE>
__global__ void __launch_bounds__(384, 1)
k_preproject(const float2* __restrict__ vin, const uint8_t* __restrict__ obs, float2* __restrict__ vout,
             float* __restrict__ vort, float* __restrict__ div, float* __restrict__ div4, uint8_t* __restrict__ nbmask, const Geom g,
             const PreParams prm, int* __restrict__ err) {
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int DEPTH = VISCOUS ? 4 : 3;            // rows between the advected row and the divergence row
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = blockIdx.x * (blockDim.x >> 5) + warp;
    if (tile >= prm.ntiles) return;
    const int chunk = tile / prm.nstrips, strip = tile - chunk * prm.nstrips;
    const int x0 = strip * (PSW - 2 * PHX) - PHX;
    const int xa = x0 + 4 * lane;                      // first of this lane's 4 columns (multiple of 4)
    const int out_lo = prm.r0 + chunk * prm.ch;
    const int out_hi = min(out_lo + prm.ch, prm.r1);
    const bool inside = xa >= 0 && xa + 3 < g.w;      // all 4 columns are grid columns
    const bool st_ok = 4 * lane >= PHX && 4 * lane < PSW - PHX && inside;
    const uint32_t edge_l = xa == 0 ? 0xffffffffu : 0u, edge_r = xa + 4 == g.w ? 0xffffffffu : 0u;
    const int lane_l = (lane + 31) & 31, lane_r = (lane + 1) & 31;
    const int xc0 = clampi(xa, 0, g.w - 4);            // clamped column group for memory safety (halo lanes)

    // Rolling windows: three rows per stage - older, middle, newest - in three register slots whose roles rotate
    // from one row to the next.  The row loop is unrolled three times so that every slot index is a compile-time
    // constant: no register ever moves (the earlier two-slot version spent a tenth of its instructions on moves).
    float2 A[3][4], B[3][4], C[3][4];                  // advected / confined / viscous velocity
    float Wv[3][4];                                    // vorticity
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            A[k][j] = B[k][j] = C[k][j] = make_float2(0.0f, 0.0f);
            Wv[k][j] = 0.0f;
        }
    const float fxc[4] = {(float)xc0, (float)(xc0 + 1), (float)(xc0 + 2), (float)(xc0 + 3)};

    // The centre cells of a row (obstacle word + 4 velocities) are the only loads that miss to DRAM - the
    // back-trace gathers land on rows this warp has just streamed through.  They are fetched one row ahead
    // so that their latency overlaps the arithmetic of the current row.
    uint32_t ow_n = 0u;
    float4 c01_n = make_float4(0.0f, 0.0f, 0.0f, 0.0f), c23_n = c01_n;
    auto fetch_row = [&](int ly_) {
        const int gy_ = g.y0 + ly_;
        if (gy_ >= 0 && gy_ < g.hg && ly_ < out_hi + DEPTH) {
            const ptrdiff_t base = lin(g, xc0, ly_);
            ow_n = *reinterpret_cast<const uint32_t*>(obs + base);
            c01_n = *reinterpret_cast<const float4*>(vin + base);
            c23_n = *reinterpret_cast<const float4*>(vin + base + 2);
        }
    };
    fetch_row(out_lo - DEPTH);

    // gathers in flight for the row about to be advected (PIPE: issued one row ahead of their use)
    Gather G[4];
    uint32_t g_ow = 0u;
    bool g_valid = false;
    auto issue_row = [&](int ly_) {          // consumes the fetched centre cells of row ly_, fetches row ly_+1
        const int gy_ = g.y0 + ly_;
        const uint32_t ow = ow_n;
        const float4 c01 = c01_n, c23 = c23_n;
        fetch_row(ly_ + 1);
        g_valid = gy_ >= 0 && gy_ < g.hg && ly_ < out_hi + DEPTH;
        if (g_valid) {
            const float2 cv[4] = {make_float2(c01.x, c01.y), make_float2(c01.z, c01.w), make_float2(c23.x, c23.y),
                                  make_float2(c23.z, c23.w)};
            const float fyc = (float)gy_;
            const float2* rowp = vin + (ptrdiff_t)ly_ * g.w;
            asm volatile("" : "+l"(rowp));           // keep the row pointer a value of its own (else it is re-derived per load)
            const int gyw = gy_ * g.w;
#pragma unroll
            for (int j = 0; j < 4; ++j) gather_issue<SLAB>(G[j], rowp, g, fxc[j], gy_, gyw, fyc, cv[j], prm.dt, prm.speed, err);
            g_ow = ow;
        }
    };
    if (PIPE) issue_row(out_lo - DEPTH);

    // one row: slot K holds the OLDER row of every window, K + 1 the middle one, K + 2 receives the newest.
    // ROT (the PIPE variant, 168 registers): the roles rotate; otherwise (128 registers, small grids) the slots are
    // fixed and the rows move down one slot after every row, which needs fewer live registers.
    constexpr bool ROT = PIPE;
    auto row = [&](auto kc, const int ly) {
        constexpr int K0 = ROT ? decltype(kc)::value % 3 : 0, K1 = (K0 + 1) % 3, K2 = (K0 + 2) % 3;
        float2 (&A0)[4] = A[K0]; float2 (&A1)[4] = A[K1]; float2 (&An)[4] = A[K2];
        float (&W0)[4] = Wv[K0]; float (&W1)[4] = Wv[K1]; float (&Wn)[4] = Wv[K2];
        float2 (&B0)[4] = B[K0]; float2 (&B1)[4] = B[K1]; float2 (&Bn)[4] = B[K2];
        const int gy = g.y0 + ly;
        // the clamp-to-edge fix-ups below only fire next to the grid's first / last row: one warp-uniform branch
        // keeps their 64 selects off every other row
        const bool near_edge = gy <= DEPTH + 1 || gy >= g.hg - 1;
        // ---- stage 0: advect row ly (ref: shader.AdvectVelocity.comp:27-50).  All 4 cells are traced
        // without branching (16 independent gathers in flight); solid cells are zeroed afterwards.
        if (!PIPE) issue_row(ly);
        // (finished for all four cells and selected afterwards: a branch per cell costs more than the mixes of a solid cell)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 fin = gather_finish(G[j], prm.diss);
            const bool keep = g_valid && !((g_ow >> (8 * j)) & 0xffu);
            An[j] = make_float2(keep ? fin.x : 0.0f, keep ? fin.y : 0.0f);
        }
        if (PIPE) issue_row(ly + 1);

        // ---- stage 1: vorticity of row r1 = ly-1 (ref: shader.CalcVorticity.comp:20-26)
        // clamp-to-edge in y: at the first / last grid row the missing neighbour row is the row itself
        const int r1 = ly - 1, g1 = gy - 1;
        if (near_edge) {
            if (g1 == 0) copy4(A0, A1);
            if (g1 == g.hg - 1) copy4(An, A1);
        }
        {
            const float ly_ = bitsel(A1[0].y, __shfl_sync(FULL, A1[3].y, lane_l), edge_l);
            const float ry_ = bitsel(A1[3].y, __shfl_sync(FULL, A1[0].y, lane_r), edge_r);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float vLy = j > 0 ? A1[j - 1].y : ly_;
                const float vRy = j < 3 ? A1[j + 1].y : ry_;
                Wn[j] = 0.5f * ((vRy - vLy) - (An[j].x - A0[j].x));
            }
            if (st_ok && r1 >= out_lo && r1 < out_hi)
                stg_stream(reinterpret_cast<float4*>(vort + lin(g, xa, r1)), make_float4(Wn[0], Wn[1], Wn[2], Wn[3]));
        }

        // ---- stage 2: confinement on row r2 = ly-2 (ref: shader.ApplyVorticity.comp:26-39)
        const int g2 = gy - 2;
        if (near_edge) {
            if (g2 == 0) copy4(W0, W1);
            if (g2 == g.hg - 1) copy4(Wn, W1);
        }
        {
            const float wl = bitsel(W1[0], __shfl_sync(FULL, W1[3], lane_l), edge_l);
            const float wr = bitsel(W1[3], __shfl_sync(FULL, W1[0], lane_r), edge_r);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float wL = j > 0 ? W1[j - 1] : wl;
                const float wR = j < 3 ? W1[j + 1] : wr;
                const float2 f = confinement_force(wL, wR, W0[j], Wn[j], W1[j], prm.scale, prm.dt);
                Bn[j] = make_float2(A0[j].x + f.x, A0[j].y + f.y);      // A0 is row ly-2 here
            }
        }

        // ---- stage 3 (optional): viscosity on row r3 = ly-3 (ref: shader.Viscosity.comp:24-31)
        float2 (&Fn)[4] = VISCOUS ? C[K2] : B[K2];           // newest row of the final pre-projection velocity
        const int rF = VISCOUS ? ly - 3 : ly - 2;            // its local row
        if (VISCOUS) {
            const int g3 = gy - 3;
            if (near_edge) {
                if (g3 == 0) copy4(B0, B1);
                if (g3 == g.hg - 1) copy4(Bn, B1);
            }
            const float lx = bitsel(B1[0].x, __shfl_sync(FULL, B1[3].x, lane_l), edge_l);
            const float ly2 = bitsel(B1[0].y, __shfl_sync(FULL, B1[3].y, lane_l), edge_l);
            const float rx = bitsel(B1[3].x, __shfl_sync(FULL, B1[0].x, lane_r), edge_r);
            const float ry = bitsel(B1[3].y, __shfl_sync(FULL, B1[0].y, lane_r), edge_r);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 x1 = j > 0 ? B1[j - 1] : make_float2(lx, ly2);
                const float2 x2 = j < 3 ? B1[j + 1] : make_float2(rx, ry);
                Fn[j].x = (x1.x + x2.x + B0[j].x + Bn[j].x + B1[j].x * prm.alpha) * prm.rbeta;
                Fn[j].y = (x1.y + x2.y + B0[j].y + Bn[j].y + B1[j].y * prm.alpha) * prm.rbeta;
            }
        }
        if (st_ok && rF >= out_lo && rF < out_hi) {
            float4* dst = reinterpret_cast<float4*>(vout + lin(g, xa, rF));
            stg_stream(dst, make_float4(Fn[0].x, Fn[0].y, Fn[1].x, Fn[1].y));
            stg_stream(dst + 1, make_float4(Fn[2].x, Fn[2].y, Fn[3].x, Fn[3].y));
        }

        // ---- stage 4: divergence + blocked-neighbour mask of row rd = rF-1
        //      (ref: shader.Divergence.comp:22-40; mask bits as in stages_ref.cu k_divergence)
        float2 (&F0)[4] = VISCOUS ? C[K0] : B[K0];      // row rd-1
        float2 (&F1)[4] = VISCOUS ? C[K1] : B[K1];      // row rd
        const int rd = rF - 1, gd = g.y0 + rd;
        if (rd >= out_lo && rd < out_hi) {
            if (near_edge) {
                if (gd == 0) copy4(F0, F1);
                if (gd == g.hg - 1) copy4(Fn, F1);
            }
            const uint32_t oM = *reinterpret_cast<const uint32_t*>(obs + lin(g, xc0, rd));
            const uint32_t oB = *reinterpret_cast<const uint32_t*>(obs + lin(g, xc0, max(gd - 1, 0) - g.y0));
            const uint32_t oT = *reinterpret_cast<const uint32_t*>(obs + lin(g, xc0, min(gd + 1, g.hg - 1) - g.y0));
            // solid flags of the 4 cells as one word, bit 0 of each byte (obstacle bytes are 0, 1 or 2; the bit that
            // the shift drags in from the next byte lands on bit 7 and is masked off)
            const uint32_t ONES = 0x01010101u;
            const uint32_t sM = (oM | (oM >> 1)) & ONES, sB4 = (oB | (oB >> 1)) & ONES, sT4 = (oT | (oT >> 1)) & ONES;
            const uint32_t sLw = __shfl_sync(FULL, sM, lane_l), sRw = __shfl_sync(FULL, sM, lane_r);
            const uint32_t sl0 = edge_l ? (sM & 1u) : (sLw >> 24);            // solid flag of column xa-1 (the cell itself at the grid's edge)
            const uint32_t sr3 = edge_r ? (sM >> 24) : (sRw & 1u);            // ... of column xa+4
            const uint32_t sL4 = (sM << 8) | sl0, sR4 = (sM >> 8) | (sr3 << 24);
            const float lx = bitsel(F1[0].x, __shfl_sync(FULL, F1[3].x, lane_l), edge_l);
            const float rx = bitsel(F1[3].x, __shfl_sync(FULL, F1[0].x, lane_r), edge_r);
            float dv[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float x1 = (sL4 & (1u << (8 * j))) ? 0.0f : (j > 0 ? F1[j - 1].x : lx);
                const float x2 = (sR4 & (1u << (8 * j))) ? 0.0f : (j < 3 ? F1[j + 1].x : rx);
                const float y1 = (sB4 & (1u << (8 * j))) ? 0.0f : F0[j].y;
                const float y2 = (sT4 & (1u << (8 * j))) ? 0.0f : Fn[j].y;
                dv[j] = 0.5f * ((x2 - x1) + (y2 - y1));
            }
            // blocked-neighbour mask: the solid flags moved to their bit, plus the grid's edges
            uint32_t mword = sL4 | (sR4 << 1) | (sB4 << 2) | (sT4 << 3);
            mword |= (edge_l & (uint32_t)NB_L) | (edge_r & ((uint32_t)NB_R << 24));
            if (gd == 0) mword |= ONES * NB_B;
            if (gd == g.hg - 1) mword |= ONES * NB_T;
            // the scaled copy the Jacobi kernels read (common.cuh NB_RAW): 0.25 b, exact for every b but a non-zero
            // |b| < 2^-124 - those cells keep b and carry the NB_RAW bit
            const float2 q = make_float2(0.25f, 0.25f), four = make_float2(4.0f, 4.0f);
            float2 s01 = __fmul2_rn(make_float2(dv[0], dv[1]), q), s23 = __fmul2_rn(make_float2(dv[2], dv[3]), q);
            const float2 u01 = __fmul2_rn(s01, four), u23 = __fmul2_rn(s23, four);
            if (((__float_as_uint(u01.x) ^ __float_as_uint(dv[0])) | (__float_as_uint(u01.y) ^ __float_as_uint(dv[1])) |
                 (__float_as_uint(u23.x) ^ __float_as_uint(dv[2])) | (__float_as_uint(u23.y) ^ __float_as_uint(dv[3]))) != 0u) {
                float sc[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    bool raw;
                    sc[j] = scaled_divergence(dv[j], &raw);
                    if (raw) mword |= (uint32_t)NB_RAW << (8 * j);
                }
                s01 = make_float2(sc[0], sc[1]);
                s23 = make_float2(sc[2], sc[3]);
            }
            if (st_ok) {
                stg_stream(reinterpret_cast<float4*>(div + lin(g, xa, rd)), make_float4(dv[0], dv[1], dv[2], dv[3]));
                stg_stream(reinterpret_cast<float4*>(div4 + lin(g, xa, rd)), make_float4(s01.x, s01.y, s23.x, s23.y));
                *reinterpret_cast<uint32_t*>(nbmask + lin(g, xa, rd)) = mword;
            }
        }
        if (!ROT) {
            copy4(A[0], A[1]); copy4(A[1], A[2]);
            copy4(Wv[0], Wv[1]); copy4(Wv[1], Wv[2]);
            copy4(B[0], B[1]); copy4(B[1], B[2]);
            if (VISCOUS) { copy4(C[0], C[1]); copy4(C[1], C[2]); }
        }
    };

    if (ROT) {
        // rows beyond out_hi + DEPTH - 1 (at most two, to complete a group of three) load and store nothing
        for (int ly = out_lo - DEPTH; ly < out_hi + DEPTH; ly += 3) {
            row(std::integral_constant<int, 0>{}, ly);
            row(std::integral_constant<int, 1>{}, ly + 1);
            row(std::integral_constant<int, 2>{}, ly + 2);
        }
    } else {
        for (int ly = out_lo - DEPTH; ly < out_hi + DEPTH; ++ly) row(std::integral_constant<int, 0>{}, ly);
    }
}

// ------------------------------------------------------------------------------------------- impulses
constexpr int SBX = 64, SBY = 4;

struct Box { int x0, x1, y0, y1; };                 // global cell coordinates, half-open
struct SplatVBoxes { int n; SplatV s[MAX_SPLATS]; Box b[MAX_SPLATS]; };
struct SplatDBoxes { int n; SplatD s[MAX_SPLATS]; Box b[MAX_SPLATS]; };

__device__ __forceinline__ bool in_box(const Box& b, int x, int y) {
    return x >= b.x0 && x < b.x1 && y >= b.y0 && y < b.y1;
}

// ref: shader.AddVelocity.comp:26-35, b.n dispatches applied in order, in place, to the cells inside
// at least one splat's bounding box; the cell is handled by the block of the FIRST box containing it.
__global__ void __launch_bounds__(SBX * SBY)
k_splat_velocity_boxes(float2* __restrict__ vel, const Geom g, const __grid_constant__ SplatVBoxes b) {
    const int i = blockIdx.z;
    const Box bx = b.b[i];
    const int x = bx.x0 + blockIdx.x * SBX + threadIdx.x;
    const int gy = bx.y0 + blockIdx.y * SBY + threadIdx.y;
    if (x >= bx.x1 || gy >= bx.y1) return;
    for (int k = 0; k < i; ++k)
        if (in_box(b.b[k], x, gy)) return;
    const ptrdiff_t pos = lin(g, x, gy - g.y0);
    float2 v = vel[pos];
    const float fxp = (float)x, fyp = (float)gy;
    for (int k = 0; k < b.n; ++k) {
        const SplatV s = b.s[k];
        const float ex = s.sx - fxp, ey = s.sy - fyp;
        const float len = sqrtf(ex * ex + ey * ey);
        if (len <= s.r) {
            const float fall = s.r - len;
            v.x = v.x + s.vx * fall / s.r;
            v.y = v.y + s.vy * fall / s.r;
        }
        v.x = clampf(v.x, -1.0f, 1.0f);
        v.y = clampf(v.y, -1.0f, 1.0f);
    }
    vel[pos] = v;
}

// The clamp that every AddVelocity dispatch applies to ALL cells, for the cells outside
// every box.  It only has work to do where some |v| > 1, which the kernel that produced the field
// recorded per band of OVER_BAND rows; blocks of clean bands leave at once.
__global__ void __launch_bounds__(256)
k_clamp_outside_boxes(float2* __restrict__ vel, const Geom g, int r0, int r1, const __grid_constant__ SplatVBoxes b,
                      const int* __restrict__ over1) {
    const int band = blockIdx.y;
    if (over1[band] == 0) return;
    const int lo = max(r0, band * OVER_BAND - g.halo), hi = min(r1, (band + 1) * OVER_BAND - g.halo);
    const int n = (hi - lo) * g.w;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int ly = lo + i / g.w, x = i % g.w;
        bool skip = false;
        for (int k = 0; k < b.n; ++k) skip = skip || in_box(b.b[k], x, g.y0 + ly);
        if (skip) continue;
        const float2 v = vel[lin(g, x, ly)];
        const float2 c = make_float2(clampf(v.x, -1.0f, 1.0f), clampf(v.y, -1.0f, 1.0f));
        if (c.x != v.x || c.y != v.y) vel[lin(g, x, ly)] = c;        // rare: most cells are already in range
    }
}

// ref: demo/shaders/shader.AddParticle.comp:25-34, b.n dispatches in order, in place, boxes only
__global__ void __launch_bounds__(SBX * SBY)
k_splat_dye_boxes(float* __restrict__ dye, const Geom dg, const __grid_constant__ SplatDBoxes b) {
    const int i = blockIdx.z;
    const Box bx = b.b[i];                       // boxes are in global dye cells, clipped to the rows held
    const int x = bx.x0 + blockIdx.x * SBX + threadIdx.x;
    const int gy = bx.y0 + blockIdx.y * SBY + threadIdx.y;
    if (x >= bx.x1 || gy >= bx.y1) return;
    for (int k = 0; k < i; ++k)
        if (in_box(b.b[k], x, gy)) return;
    const ptrdiff_t pos = lin(dg, x, gy - dg.y0);
    float v = dye[pos];
    const float fxp = (float)x, fyp = (float)gy;
    for (int k = 0; k < b.n; ++k) {
        const SplatD s = b.s[k];
        const float ex = s.sx - fxp, ey = s.sy - fyp;
        const float len = sqrtf(ex * ex + ey * ey);
        if (len <= s.r) v = clampf(v + s.value * (s.r - len) / s.r, 0.0f, 255.0f);
    }
    dye[pos] = v;
}

// ref: shader.AddCircleObstacle.comp:24-36 for up to MAX_CIRCLES queued circles in one launch, each on
// its own bounding box (blockIdx.z); overlapping circles store the same value.
struct CircleBoxes { int n; float sx[MAX_CIRCLES], sy[MAX_CIRCLES], r[MAX_CIRCLES]; Box b[MAX_CIRCLES]; int first[MAX_CIRCLES + 1]; };
constexpr int CBX = 32, CBY = 8, CCELLS = 4;     // a block covers 128 x 8 cells, 4 consecutive cells per thread
// Blocks are numbered circle by circle (first[i] = first block of circle i), so the grid holds exactly the
// blocks the bounding boxes need whatever the mix of radii.
__global__ void __launch_bounds__(CBX * CBY)
k_add_circles(uint8_t* __restrict__ obs, const Geom g, const __grid_constant__ CircleBoxes c) {
    int i = 0;
    while (i + 1 < c.n && (int)blockIdx.x >= c.first[i + 1]) ++i;
    const Box bx = c.b[i];
    const int xa = bx.x0 & ~3;                                   // 4-cell groups aligned in x
    const int bw = (bx.x1 - xa + CBX * CCELLS - 1) / (CBX * CCELLS);
    const int lb = (int)blockIdx.x - c.first[i];
    const int x = xa + ((lb % bw) * CBX + threadIdx.x) * CCELLS;
    const int gy = bx.y0 + (lb / bw) * CBY + threadIdx.y;
    if (x >= bx.x1 || gy >= bx.y1) return;
    const float ey = c.sy[i] - (float)gy, r = c.r[i];
    uint32_t in = 0u;
#pragma unroll
    for (int j = 0; j < CCELLS; ++j) {
        const float ex = c.sx[i] - (float)(x + j);
        if (x + j >= bx.x0 && x + j < bx.x1 && sqrtf(ex * ex + ey * ey) <= r) in |= 1u << j;
    }
    if (!in) return;
    uint8_t* row = obs + lin(g, x, gy - g.y0);
    if (in == 0xfu && (reinterpret_cast<uintptr_t>(row) & 3u) == 0) {
        *reinterpret_cast<uint32_t*>(row) = OBS_DYNAMIC * 0x01010101u;
    } else {
#pragma unroll
        for (int j = 0; j < CCELLS; ++j)
            if (in & (1u << j)) row[j] = OBS_DYNAMIC;
    }
}

// Cells with sqrt(ex^2 + ey^2) <= r lie within r + 2 of the centre along each axis (sqrt is monotone and
// >= |ex| up to one rounding); clip to [xlo, xhi) x [ylo, yhi).  A negative or NaN radius selects nothing.
Box splat_box(float sx, float sy, float r, int xlo, int xhi, int ylo, int yhi) {
    Box b{0, 0, 0, 0};
    if (!(r >= 0.0f)) return b;
    const double m = (double)r + 2.0;
    const double x0 = (double)sx - m, x1 = (double)sx + m, y0 = (double)sy - m, y1 = (double)sy + m;
    if (!(x1 >= xlo && x0 <= xhi && y1 >= ylo && y0 <= yhi)) return b;     // also rejects NaN centres
    b.x0 = x0 > xlo ? (int)x0 : xlo;
    b.x1 = x1 < xhi - 1 ? (int)x1 + 1 : xhi;
    b.y0 = y0 > ylo ? (int)y0 : ylo;
    b.y1 = y1 < yhi - 1 ? (int)y1 + 1 : yhi;
    if (b.x1 <= b.x0 || b.y1 <= b.y0) b = Box{0, 0, 0, 0};
    return b;
}

template <class Boxes>
dim3 boxes_grid(const Boxes& b) {
    int mw = 0, mh = 0;
    for (int i = 0; i < b.n; ++i) {
        mw = b.b[i].x1 - b.b[i].x0 > mw ? b.b[i].x1 - b.b[i].x0 : mw;
        mh = b.b[i].y1 - b.b[i].y0 > mh ? b.b[i].y1 - b.b[i].y0 : mh;
    }
    return dim3((mw + SBX - 1) / SBX, (mh + SBY - 1) / SBY, b.n);
}

// ---------------------------------------------------------------------------------------------- borders
// ref: shader.InitBoundaries.comp:14-34 - zero the four border lines of the READ velocity in place.
// One thread per border cell of rows [r0, r1): the two side columns, plus the first / last grid row.
__global__ void __launch_bounds__(256)
k_zero_borders(float2* __restrict__ vel, const Geom g, int r0, int r1) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int nrows = r1 - r0;
    const float2 z = make_float2(0.0f, 0.0f);
    if (i < nrows) {
        vel[lin(g, 0, r0 + i)] = z;
        vel[lin(g, g.w - 1, r0 + i)] = z;
    } else if (i < nrows + g.w) {
        const int top = -g.y0;                       // local row of grid row 0
        if (top >= r0 && top < r1) vel[lin(g, i - nrows, top)] = z;
    } else if (i < nrows + 2 * g.w) {
        const int bot = g.hg - 1 - g.y0;
        if (bot >= r0 && bot < r1) vel[lin(g, i - nrows - g.w, bot)] = z;
    }
}

// ------------------------------------------------------------------------------------------- gradient
// ref: shader.SubtractGradient.comp:24-46 via the blocked-neighbour mask; also records, per band of
// OVER_BAND rows, whether any |v| > 1 leaves the step (lets the next add_velocity skip the all-cell clamp
// for the bands that cannot need it).
constexpr int GBX = 64, GBY = 4;
__global__ void __launch_bounds__(GBX * GBY)
k_gradient_mask(const float2* __restrict__ vin, const float* __restrict__ p, const uint8_t* __restrict__ nbmask,
                float2* __restrict__ vout, const Geom g, int r0, int r1, int* __restrict__ over1) {
    const int x = blockIdx.x * GBX + threadIdx.x;
    const int ly = r0 + (int)blockIdx.y * GBY + threadIdx.y;
    if (x >= g.w || ly >= r1) return;
    const ptrdiff_t pos = lin(g, x, ly);
    const uint8_t m = nbmask[pos];
    const float c = p[pos];
    const float x1 = (m & NB_L) ? c : p[pos - 1];
    const float x2 = (m & NB_R) ? c : p[pos + 1];
    const float y1 = (m & NB_B) ? c : p[pos - g.w];
    const float y2 = (m & NB_T) ? c : p[pos + g.w];
    float2 v = vin[pos];
    v.x = v.x - 0.5f * (x2 - x1);
    v.y = v.y - 0.5f * (y2 - y1);
    vout[pos] = v;
    if (fabsf(v.x) > 1.0f || fabsf(v.y) > 1.0f) over1[(ly + g.halo) / OVER_BAND] = 1;
}

// Same, 4 cells per thread (width % 4 == 0): three float4 pressure loads (row, row-1, row+1 - every
// address is independent of the mask, so all loads are in flight at once), two float4 velocity loads,
// one mask word; the columns left / right of the group come from the neighbouring lanes.
constexpr int G4X = 32, G4Y = 8;
__global__ void __launch_bounds__(G4X * G4Y)
k_gradient_mask4(const float2* __restrict__ vin, const float* __restrict__ p, const uint8_t* __restrict__ nbmask,
                 float2* __restrict__ vout, const Geom g, int r0, int r1, int* __restrict__ over1) {
    const int x = (blockIdx.x * G4X + threadIdx.x) * 4;
    const int ly = r0 + (int)blockIdx.y * G4Y + threadIdx.y;
    const bool live = x < g.w && ly < r1;
    const int xs = live ? x : 0, ys = live ? ly : r0;                 // dead threads still take part in the shuffles
    const int gy = g.y0 + ys;
    const ptrdiff_t pos = lin(g, xs, ys);
    const ptrdiff_t up = lin(g, xs, max(gy - 1, 0) - g.y0), dn = lin(g, xs, min(gy + 1, g.hg - 1) - g.y0);
    const float4 pc = *reinterpret_cast<const float4*>(p + pos);
    const float4 pb = *reinterpret_cast<const float4*>(p + up);       // row - 1 ("B")
    const float4 pt = *reinterpret_cast<const float4*>(p + dn);       // row + 1 ("T")
    const uint32_t mw = *reinterpret_cast<const uint32_t*>(nbmask + pos);
    const float4 v01 = *reinterpret_cast<const float4*>(vin + pos);
    const float4 v23 = *reinterpret_cast<const float4*>(vin + pos + 2);
    float pl = __shfl_up_sync(0xffffffffu, pc.w, 1), pr = __shfl_down_sync(0xffffffffu, pc.x, 1);
    if (threadIdx.x == 0) pl = p[pos - (xs > 0 ? 1 : 0)];             // group at the start of the warp's span
    if (threadIdx.x == G4X - 1) pr = p[pos + (xs + 4 < g.w ? 4 : 3)];
    if (!live) return;
    const float c[4] = {pc.x, pc.y, pc.z, pc.w};
    const float b[4] = {pb.x, pb.y, pb.z, pb.w};
    const float t[4] = {pt.x, pt.y, pt.z, pt.w};
    const float vx[4] = {v01.x, v01.z, v23.x, v23.z}, vy[4] = {v01.y, v01.w, v23.y, v23.w};
    float ox[4], oy[4];
    bool over = false;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const uint32_t m = mw >> (8 * j);
        const float x1 = (m & NB_L) ? c[j] : (j > 0 ? c[j - 1] : pl);
        const float x2 = (m & NB_R) ? c[j] : (j < 3 ? c[j + 1] : pr);
        const float y1 = (m & NB_B) ? c[j] : b[j];
        const float y2 = (m & NB_T) ? c[j] : t[j];
        ox[j] = vx[j] - 0.5f * (x2 - x1);
        oy[j] = vy[j] - 0.5f * (y2 - y1);
        over = over || fabsf(ox[j]) > 1.0f || fabsf(oy[j]) > 1.0f;
    }
    float4* dst = reinterpret_cast<float4*>(vout + pos);
    stg_stream(dst, make_float4(ox[0], oy[0], ox[1], oy[1]));
    stg_stream(dst + 1, make_float4(ox[2], oy[2], ox[3], oy[3]));
    if (over) over1[(ly + g.halo) / OVER_BAND] = 1;
}

}  // namespace

bool preproject_supported(const Geom& g) {
    // the gathers use 32-bit cell offsets: global row x width must stay below 2^31
    return g.w % 4 == 0 && g.w >= 8 && (size_t)g.w * (size_t)g.hg < ((size_t)1 << 31);
}

int launch_zero_borders(float2* vel, Geom g, int r0, int r1, cudaStream_t st) {
    if (r1 <= r0) return 0;
    const int n = (r1 - r0) + 2 * g.w;
    k_zero_borders<<<(n + 255) / 256, 256, 0, st>>>(vel, g, r0, r1);
    return 1;
}

int launch_preproject(const float2* vin, const uint8_t* obs, float2* vout, float* vort, float* div, float* div4,
                      uint8_t* nbmask, Geom g, int r0, int r1, float dt, float speed, float diss, float scale, bool viscous, float alpha,
                      float rbeta, int sm_count, int* err, cudaStream_t st) {
    if (r1 <= r0) return 0;
    PreParams prm;
    prm.r0 = r0; prm.r1 = r1;
    prm.dt = dt; prm.speed = speed; prm.diss = diss; prm.scale = scale; prm.alpha = alpha; prm.rbeta = rbeta;
    prm.nstrips = (g.w + (PSW - 2 * PHX) - 1) / (PSW - 2 * PHX);
    const int rows = r1 - r0;
    // PIPE pays off once the grid fills the GPU; small grids are latency-bound per warp and prefer the variant
    // with more, lighter CTAs (measured: 640x360 and 1024^2 vs 4096^2 and 32768x4096)
    static const int pipe_env = [] { const char* e = getenv("NATRIX_PRE_PIPE"); return e ? (atoi(e) != 0 ? 1 : 0) : -1; }();
    const bool pipe = pipe_env >= 0 ? pipe_env != 0 : (size_t)g.w * (size_t)(r1 - r0) >= ((size_t)4 << 20);
    const int warps = 12, resident = 1;
    int nchunks = (sm_count * warps * resident) / prm.nstrips;  // one tile per resident warp
    if (nchunks < 1) nchunks = 1;
    int ch = (rows + nchunks - 1) / nchunks;
    static const int min_ch = [] { const char* e = getenv("NATRIX_PRE_MINCH"); return e && atoi(e) > 0 ? atoi(e) : 4; }();
    if (ch < min_ch) ch = min_ch;
    prm.ch = ch;
    nchunks = (rows + ch - 1) / ch;
    prm.ntiles = prm.nstrips * nchunks;
    // a small grid has fewer tiles than the GPU has resident warps (640 x 360: 540): spread them over all the SMs
    // in smaller blocks instead of filling a third of the SMs with 12 warps each - every tile is one dependent
    // chain of rows, and a warp that shares its scheduler with two others walks it more slowly
    // (rounded up: never more blocks than SMs, the kernel runs one block per SM)
    const int bw = std::max(2, std::min(warps, (prm.ntiles + sm_count - 1) / sm_count));
    const int blocks = (prm.ntiles + bw - 1) / bw;
    const bool slab = g.hl != g.hg;
#define NATRIX_PRE(V, S, P) k_preproject<V, S, P><<<blocks, bw * 32, 0, st>>>(vin, obs, vout, vort, div, div4, nbmask, g, prm, err)
    if (pipe) {
        if (viscous) { if (slab) NATRIX_PRE(true, true, true); else NATRIX_PRE(true, false, true); }
        else { if (slab) NATRIX_PRE(false, true, true); else NATRIX_PRE(false, false, true); }
    } else {
        if (viscous) { if (slab) NATRIX_PRE(true, true, false); else NATRIX_PRE(true, false, false); }
        else { if (slab) NATRIX_PRE(false, true, false); else NATRIX_PRE(false, false, false); }
    }
#undef NATRIX_PRE
    return 1;
}

int launch_gradient_mask(const float2* vin, const float* p, const uint8_t* nbmask, float2* vout, Geom g, int r0, int r1,
                         int* over1, cudaStream_t st) {
    if (r1 <= r0) return 0;
    if (g.w % 4 == 0) {
        dim3 grid((g.w / 4 + G4X - 1) / G4X, (r1 - r0 + G4Y - 1) / G4Y, 1);
        k_gradient_mask4<<<grid, dim3(G4X, G4Y, 1), 0, st>>>(vin, p, nbmask, vout, g, r0, r1, over1);
    } else {
        dim3 grid((g.w + GBX - 1) / GBX, (r1 - r0 + GBY - 1) / GBY, 1);
        k_gradient_mask<<<grid, dim3(GBX, GBY, 1), 0, st>>>(vin, p, nbmask, vout, g, r0, r1, over1);
    }
    return 1;
}

int launch_splat_velocity_boxes(float2* vel, Geom g, int r0, int r1, const SplatV* splats, int n, const int* over1,
                                int sm_count, cudaStream_t st) {
    if (n <= 0 || r1 <= r0) return 0;
    SplatVBoxes b;
    b.n = n;
    for (int i = 0; i < n; ++i) {
        b.s[i] = splats[i];
        b.b[i] = splat_box(splats[i].sx, splats[i].sy, splats[i].r, 0, g.w, g.y0 + r0, g.y0 + r1);
    }
    int launched = 0;
    (void)sm_count;
    const int nbands = (g.hl + 2 * g.halo + OVER_BAND - 1) / OVER_BAND;
    k_clamp_outside_boxes<<<dim3(16, nbands, 1), 256, 0, st>>>(vel, g, r0, r1, b, over1);
    ++launched;
    const dim3 grid = boxes_grid(b);
    if (grid.x > 0 && grid.y > 0) {
        k_splat_velocity_boxes<<<grid, dim3(SBX, SBY, 1), 0, st>>>(vel, g, b);
        ++launched;
    }
    return launched;
}

int launch_add_circles(uint8_t* obs, Geom g, int r0, int r1, const float* sxyr, int n, cudaStream_t st) {
    int launched = 0;
    for (int base = 0; base < n;) {
        CircleBoxes c;
        c.n = 0;
        c.first[0] = 0;
        int i = base;
        for (; i < n && c.n < MAX_CIRCLES; ++i) {
            const Box b = splat_box(sxyr[3 * i], sxyr[3 * i + 1], sxyr[3 * i + 2], 0, g.w, g.y0 + r0, g.y0 + r1);
            if (b.x1 <= b.x0) continue;
            c.sx[c.n] = sxyr[3 * i]; c.sy[c.n] = sxyr[3 * i + 1]; c.r[c.n] = sxyr[3 * i + 2]; c.b[c.n] = b;
            const int bw = (b.x1 - (b.x0 & ~3) + CBX * CCELLS - 1) / (CBX * CCELLS), bh = (b.y1 - b.y0 + CBY - 1) / CBY;
            c.first[c.n + 1] = c.first[c.n] + bw * bh;
            ++c.n;
        }
        base = i;
        if (c.n == 0 || c.first[c.n] == 0) continue;
        k_add_circles<<<c.first[c.n], dim3(CBX, CBY, 1), 0, st>>>(obs, g, c);
        ++launched;
    }
    return launched;
}

int launch_splat_dye_boxes(float* dye, Geom dg, int r0, int r1, const SplatD* splats, int n, cudaStream_t st) {
    if (n <= 0 || r1 <= r0) return 0;
    SplatDBoxes b;
    b.n = n;
    for (int i = 0; i < n; ++i) {
        b.s[i] = splats[i];
        b.b[i] = splat_box(splats[i].sx, splats[i].sy, splats[i].r, 0, dg.w, dg.y0 + r0, dg.y0 + r1);
    }
    const dim3 grid = boxes_grid(b);
    if (grid.x == 0 || grid.y == 0) return 0;
    k_splat_dye_boxes<<<grid, dim3(SBX, SBY, 1), 0, st>>>(dye, dg, b);
    return 1;
}

}  // namespace natrix
