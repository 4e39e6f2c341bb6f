// jacobi_smem.cu - pressure-Jacobi sweeps for SMALL grids (the demo's 640 x 360, config 2's 1024^2) and for
// widths the TMA kernel cannot address.
//
// ref: shader.Poisson.comp:24-37 applied `depth` times (fluid_simulator.py:251-255).
//
// jacobi_tb.cu streams one tile per WARP through registers: the right shape when every warp has hundreds of
// rows to march, but a launch can never be shorter than the (2 * depth + rows) row steps one warp runs back to
// back - about 20 us at 640 x 360, where a tile has 4 output rows behind 16 warm-up rows, and the step's 50
// sweeps become seven such launches one after the other.  Small grids are latency-bound, so this kernel turns the
// decomposition round: one tile per BLOCK, the whole padded tile (p twice, div, mask: 13 B per cell) in shared
// memory, all 1024 threads on every sweep with a block barrier between sweeps.  A sweep of a 10 000-cell tile
// is ~2 000 cycles, and depth can go to 16 (the halo of `depth` cells per side is recomputed; compute is cheap
// when the grid fits a single wave of blocks).  The launch that runs the last sweeps of a step also subtracts the
// pressure gradient from the velocity (GRAD): the final pressure tile, its neighbours and the mask are on chip already.
//
// Arithmetic and operand order are those of every other Jacobi kernel here - ((x1 + x2) + y1) + y2, then
// fma(sum, 0.25, -b4) on the pre-scaled divergence (== (sum - b) * 0.25, common.cuh NB_RAW), centre substituted
// for blocked neighbours - so results are bit-identical.  Cells of the padded tile that
// lie outside the grid (or outside the rows the slab holds) are zero with all four neighbours blocked: they stay
// zero and nothing in the grid ever reads them, because every in-grid cell next to the edge carries the blocked
// bit for that direction (the shader's clamp-to-edge rule, written into the mask by the divergence stage).
#include <algorithm>
#include <cstdlib>

#include "kernels.h"

namespace natrix {
namespace {

constexpr int SM_THREADS = 1024;
constexpr int SM_WARPS = SM_THREADS / 32;
constexpr int SM_MAX_PW = 128;              // padded tile width: one warp covers a row, 4 cells per lane

struct SmemParams {
    const float* pin;
    const float* div;       // the pre-scaled divergence
    const uint8_t* nbm;
    float* pout;
    int w;
    int row_lo, row_hi;     // local rows that exist AND lie inside the global grid
    int r0, r1;             // output rows
    int tx, ty, ntx;        // output tile size, tiles per row of tiles
    int depth, hx;          // halo rows per side (>= sweeps); halo columns per side (multiple of 4, >= depth)
    int pw, ph;             // padded tile: pw = tx + 2 hx (<= 128, multiple of 4), ph = ty + 2 depth
    int p_zero;
    int sweeps;             // sweeps this launch runs; depth = sweeps, or sweeps + 1 when the gradient follows
    // GRAD: the launch that runs the step's last sweeps also subtracts the pressure gradient (one more halo cell)
    const float2* vin;
    float2* vout;
    int* over1;             // per band of OVER_BAND allocated rows: some |v| > 1 was written (see fused.cu)
    int halo_rows;          // Geom::halo, for the band index
};

__device__ __forceinline__ float jacobi_cell(float c, float l, float r, float b, float t, float d, uint32_t m) {
    const float x1 = (m & NB_L) ? c : l;
    const float x2 = (m & NB_R) ? c : r;
    const float y1 = (m & NB_B) ? c : b;
    const float y2 = (m & NB_T) ? c : t;
    const float sum = x1 + x2 + y1 + y2;
    return (m & NB_RAW) ? (sum - d) * 0.25f : __fmaf_rn(sum, 0.25f, -d);     // d is 0.25 b, or b itself under NB_RAW
}

// VEC: width % 4 == 0, so a lane's 4 columns are inside or outside the grid together and global accesses are
// 16 B (p, div) / 4 B (mask) wide; otherwise every cell is bounds-checked and moved on its own.
template <bool VEC, bool GRAD>
__global__ void __launch_bounds__(SM_THREADS, 1)
k_jacobi_smem(const SmemParams prm) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int pw = prm.pw, ph = prm.ph, cells = pw * ph;
    float* buf0 = reinterpret_cast<float*>(smem_raw);
    float* buf1 = buf0 + cells;
    float* dv = buf1 + cells;
    uint8_t* mk = reinterpret_cast<uint8_t*>(dv + cells);

    const int tcol = blockIdx.x % prm.ntx, trow = blockIdx.x / prm.ntx;
    const int X0 = tcol * prm.tx - prm.hx;                  // grid column of padded column 0 (multiple of 4)
    const int Y0 = prm.r0 + trow * prm.ty - prm.depth;      // local row of padded row 0
    const int groups = pw >> 2;
    const bool lane_on = lane < groups;
    const int px = 4 * (lane_on ? lane : groups - 1);       // idle lanes shadow the last group (shuffles stay full-warp)
    const int x = X0 + px;

    // programmatic dependent launch: everything above touched no field (see jacobi_tb.cu)
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    // ---- fill the padded tile
    for (int row = warp; row < ph; row += SM_WARPS) {
        const int ly = Y0 + row;
        const bool row_ok = ly >= prm.row_lo && ly < prm.row_hi;
        float4 p4 = make_float4(0.0f, 0.0f, 0.0f, 0.0f), d4 = p4;
        uint32_t m4 = 0x0f0f0f0fu;
        const ptrdiff_t base = (ptrdiff_t)ly * prm.w + x;
        if (VEC) {
            if (row_ok && x >= 0 && x < prm.w) {
                if (!prm.p_zero) p4 = *reinterpret_cast<const float4*>(prm.pin + base);
                d4 = *reinterpret_cast<const float4*>(prm.div + base);
                m4 = *reinterpret_cast<const uint32_t*>(prm.nbm + base);
            }
        } else if (row_ok) {
            float pv[4] = {0.0f, 0.0f, 0.0f, 0.0f}, dd[4] = {0.0f, 0.0f, 0.0f, 0.0f};
            m4 = 0u;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint32_t m = 0x0fu;
                if (x + j >= 0 && x + j < prm.w) {
                    if (!prm.p_zero) pv[j] = prm.pin[base + j];
                    dd[j] = prm.div[base + j];
                    m = prm.nbm[base + j];
                }
                m4 |= m << (8 * j);
            }
            p4 = make_float4(pv[0], pv[1], pv[2], pv[3]);
            d4 = make_float4(dd[0], dd[1], dd[2], dd[3]);
        }
        if (lane_on) {
            const int o = row * pw + px;
            *reinterpret_cast<float4*>(buf0 + o) = p4;
            *reinterpret_cast<float4*>(dv + o) = d4;
            *reinterpret_cast<uint32_t*>(mk + o) = m4;
        }
    }
    __syncthreads();

    // ---- `depth` sweeps in shared memory.  Sweep s only needs rows [s, ph - s): what lies outside has already
    // been reached by the garbage that creeps in from the tile's border one cell per sweep.
    float* src = buf0;
    float* dst = buf1;
    for (int s = 1; s <= prm.sweeps; ++s) {
        for (int row = s + warp; row < ph - s; row += SM_WARPS) {
            const int o = row * pw + px;
            const float4 c = *reinterpret_cast<const float4*>(src + o);
            const float4 b = *reinterpret_cast<const float4*>(src + o - pw);     // row - 1 ("B")
            const float4 t = *reinterpret_cast<const float4*>(src + o + pw);     // row + 1 ("T")
            const float4 d = *reinterpret_cast<const float4*>(dv + o);
            const uint32_t m = *reinterpret_cast<const uint32_t*>(mk + o);
            // left / right neighbours across lanes; the outermost columns of the padded tile receive garbage,
            // which is what the halo is for
            const float l = __shfl_up_sync(0xffffffffu, c.w, 1);
            const float r = __shfl_down_sync(0xffffffffu, c.x, 1);
            float4 n;
            if (__any_sync(0xffffffffu, m != 0u)) {
                n.x = jacobi_cell(c.x, l, c.y, b.x, t.x, d.x, m);
                n.y = jacobi_cell(c.y, c.x, c.z, b.y, t.y, d.y, m >> 8);
                n.z = jacobi_cell(c.z, c.y, c.w, b.z, t.z, d.z, m >> 16);
                n.w = jacobi_cell(c.w, c.z, r, b.w, t.w, d.w, m >> 24);
            } else {
                n.x = __fmaf_rn(l + c.y + b.x + t.x, 0.25f, -d.x);
                n.y = __fmaf_rn(c.x + c.z + b.y + t.y, 0.25f, -d.y);
                n.z = __fmaf_rn(c.y + c.w + b.z + t.z, 0.25f, -d.z);
                n.w = __fmaf_rn(c.z + r + b.w + t.w, 0.25f, -d.w);
            }
            if (lane_on) *reinterpret_cast<float4*>(dst + o) = n;
        }
        __syncthreads();
        float* tmp = src; src = dst; dst = tmp;
    }

    // ---- the tile's own cells leave; GRAD: and the projected velocity with them (ref: shader.SubtractGradient.comp:24-46
    // through the blocked-neighbour mask, same expressions as k_gradient_mask4) - the pressure tile, its mask and one
    // valid ring of neighbours are in shared memory already
    const bool col_out = lane_on && px >= prm.hx && px < prm.hx + prm.tx;
    for (int row = prm.depth + warp; row < prm.depth + prm.ty; row += SM_WARPS) {
        const int ly = Y0 + row;
        const float4 v = *reinterpret_cast<const float4*>(src + row * pw + px);
        float pl = 0.0f, pr = 0.0f;
        if (GRAD) {                                     // all lanes take part in the shuffles
            pl = __shfl_up_sync(0xffffffffu, v.w, 1);
            pr = __shfl_down_sync(0xffffffffu, v.x, 1);
        }
        if (ly >= prm.r1 || !col_out) continue;
        const ptrdiff_t base = (ptrdiff_t)ly * prm.w + x;
        if (VEC) {
            if (x < prm.w) *reinterpret_cast<float4*>(prm.pout + base) = v;
        } else {
            const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (x + j < prm.w) prm.pout[base + j] = vv[j];
        }
        if (GRAD) {
            const float4 pb = *reinterpret_cast<const float4*>(src + (row - 1) * pw + px);       // row - 1 ("B")
            const float4 pt = *reinterpret_cast<const float4*>(src + (row + 1) * pw + px);       // row + 1 ("T")
            const uint32_t mw = *reinterpret_cast<const uint32_t*>(mk + row * pw + px);
            const float c[4] = {v.x, v.y, v.z, v.w}, b[4] = {pb.x, pb.y, pb.z, pb.w}, t[4] = {pt.x, pt.y, pt.z, pt.w};
            float vx[4] = {0.0f, 0.0f, 0.0f, 0.0f}, vy[4] = {0.0f, 0.0f, 0.0f, 0.0f};
            if (VEC) {
                if (x < prm.w) {
                    const float4 v01 = *reinterpret_cast<const float4*>(prm.vin + base), v23 = *reinterpret_cast<const float4*>(prm.vin + base + 2);
                    vx[0] = v01.x; vy[0] = v01.y; vx[1] = v01.z; vy[1] = v01.w; vx[2] = v23.x; vy[2] = v23.y; vx[3] = v23.z; vy[3] = v23.w;
                }
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (x + j < prm.w) { const float2 q = prm.vin[base + j]; vx[j] = q.x; vy[j] = q.y; }
            }
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
                over = over || ((x + j < prm.w) && (fabsf(ox[j]) > 1.0f || fabsf(oy[j]) > 1.0f));
            }
            if (VEC) {
                if (x < prm.w) {
                    float4* dstv = reinterpret_cast<float4*>(prm.vout + base);
                    dstv[0] = make_float4(ox[0], oy[0], ox[1], oy[1]);
                    dstv[1] = make_float4(ox[2], oy[2], ox[3], oy[3]);
                }
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (x + j < prm.w) prm.vout[base + j] = make_float2(ox[j], oy[j]);
            }
            if (over) prm.over1[(ly + prm.halo_rows) / OVER_BAND] = 1;
        }
    }
}

int env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}

}  // namespace

int jacobi_smem_max_depth() {
    static const int d = std::min(16, std::max(1, env_int("NATRIX_SMEM_DEPTH", 16)));
    return d;
}

// Grids up to this many cells take the shared-memory kernel (measured on B200); larger ones the
// register-streaming TMA kernel.
size_t jacobi_smem_cell_limit() {
    static const size_t n = (size_t)env_int("NATRIX_SMEM_CELLS", 1536 * 1024);
    return n;
}

int launch_jacobi_smem(const float* pin, const float* div4, const uint8_t* nbmask, float* pout, Geom g, int depth,
                       int r0, int r1, bool p_is_zero, int sm_count, cudaStream_t st, const float2* grad_vin, float2* grad_vout,
                       int* over1) {
    if (r1 <= r0) return 0;
    if (depth < 1 || depth > 16) return -1;
    static const size_t smem_cap = (size_t)env_int("NATRIX_SMEM_KB", 200) * 1024;
    const bool grad = grad_vin != nullptr;
    SmemParams prm;
    prm.pin = pin; prm.div = div4; prm.nbm = nbmask; prm.pout = pout;
    prm.vin = grad_vin; prm.vout = grad_vout; prm.over1 = over1; prm.halo_rows = g.halo;
    prm.w = g.w;
    prm.row_lo = std::max(-g.halo, -g.y0);
    prm.row_hi = std::min(g.hl + g.halo, g.hg - g.y0);
    prm.r0 = r0; prm.r1 = r1;
    prm.sweeps = depth;
    prm.depth = depth + (grad ? 1 : 0);      // the gradient reads one more ring of final pressures
    prm.hx = (prm.depth + 3) & ~3;
    prm.p_zero = p_is_zero ? 1 : 0;
    const int rows = r1 - r0;
    const int tx_max = SM_MAX_PW - 2 * prm.hx;
    prm.ntx = (g.w + tx_max - 1) / tx_max;
    prm.tx = (((g.w + prm.ntx - 1) / prm.ntx) + 3) & ~3;
    prm.pw = prm.tx + 2 * prm.hx;
    // rows of tiles: fill whole waves of one block per SM; more waves while a tile does not fit shared memory
    int nty = 1;
    for (int waves = 1;; ++waves) {
        nty = std::max(1, (sm_count * waves) / prm.ntx);
        nty = std::min(nty, rows);
        prm.ty = (rows + nty - 1) / nty;
        prm.ph = prm.ty + 2 * prm.depth;
        if ((size_t)prm.pw * prm.ph * 13 <= smem_cap || prm.ty == 1) break;
    }
    nty = (rows + prm.ty - 1) / prm.ty;
    const size_t smem = (size_t)prm.pw * prm.ph * 13;
    const bool vec = g.w % 4 == 0;
    using Fn = void (*)(const SmemParams);
    const Fn fn = grad ? (vec ? k_jacobi_smem<true, true> : k_jacobi_smem<false, true>)
                       : (vec ? k_jacobi_smem<true, false> : k_jacobi_smem<false, false>);
    // the opt-in to > 48 KB of dynamic shared memory is per device (a process may drive several)
    static bool attr_set[64][4] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    dev &= 63;
    const int slot = (vec ? 1 : 0) + (grad ? 2 : 0);
    if (!attr_set[dev][slot]) {
        if (cudaFuncSetAttribute((const void*)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cap) != cudaSuccess) return -1;
        attr_set[dev][slot] = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(prm.ntx * nty));
    cfg.blockDim = dim3(SM_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    static const int pdl = env_int("NATRIX_TB_PDL", 1);
    cfg.numAttrs = pdl ? 1 : 0;
    if (cudaLaunchKernelEx(&cfg, fn, prm) != cudaSuccess) return -1;
    return 1;
}

}  // namespace natrix
