// solvers.cu - pressure solvers that are NOT reference behaviour, opt-in through
// NATRIX_OPT_SOLVER: red-black successive over-relaxation and a geometric multigrid V-cycle.
//
// Both solve the system the reference's Jacobi loop iterates on (shader.Poisson.comp:24-37,
// fluid_simulator.py:251-255): x1 + x2 + y1 + y2 - 4 p = div, with the centre pressure substituted for a
// neighbour that is solid or outside the grid (the 4 bits of the blocked-neighbour mask).  The reference offers
// only Jacobi; these exist because 2 - 4 V(2,2) cycles leave the residual of 400 - 800 Jacobi sweeps
// (scripts/solver_study.py).  Every kernel follows oracle/natrix_oracle.py (rb_sor_sweep, mg_restrict, mg_prolong,
// mg_v_cycle) operation by operation - same operand order, no FMA contraction - so the results are bit-identical to
// that NumPy restatement.  Full grids only (no slab exchange is defined for them).
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "kernels.h"

namespace natrix {
namespace {

constexpr int SBX_ = 64, SBY_ = 4;

// the shader's Jacobi update of one cell from the current field (shader.Poisson.comp:32-37)
__device__ __forceinline__ float gs_cell(const float* __restrict__ p, const float* __restrict__ rhs,
                                         const uint8_t* __restrict__ mask, int w, ptrdiff_t pos, float c) {
    const uint32_t m = mask[pos];
    const float x1 = (m & NB_L) ? c : p[pos - 1];
    const float x2 = (m & NB_R) ? c : p[pos + 1];
    const float y1 = (m & NB_B) ? c : p[pos - w];
    const float y2 = (m & NB_T) ? c : p[pos + w];
    return (x1 + x2 + y1 + y2 - rhs[pos]) * 0.25f;
}

// One colour of a red-black sweep, in place: p <- p + omega * (gs - p) for the cells with (x + y) & 1 == colour.
// A cell's four neighbours have the other colour, so nothing this launch writes is read by it.
__global__ void __launch_bounds__(SBX_ * SBY_)
k_sor_colour(float* p, const float* __restrict__ rhs, const uint8_t* __restrict__ mask, int w, int h, int colour, float omega) {
    const int y = blockIdx.y * SBY_ + threadIdx.y;
    const int x = 2 * (blockIdx.x * SBX_ + threadIdx.x) + ((y + colour) & 1);
    if (x >= w || y >= h) return;
    const ptrdiff_t pos = (ptrdiff_t)y * w + x;
    const float c = p[pos];
    const float gs = gs_cell(p, rhs, mask, w, pos, c);
    p[pos] = c + omega * (gs - c);
}

// coarse solid map: a coarse cell is solid when all four children are (fine: any non-zero byte is solid)
__global__ void __launch_bounds__(256)
k_mg_coarsen(const uint8_t* __restrict__ fine, int wf, uint8_t* __restrict__ coarse, int wc, int hc) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= wc || y >= hc) return;
    const uint8_t* f = fine + (ptrdiff_t)(2 * y) * wf + 2 * x;
    coarse[(ptrdiff_t)y * wc + x] = (f[0] && f[wf] && f[1] && f[wf + 1]) ? 1 : 0;
}

// blocked-neighbour mask of a level from its solid map (bits as written by the divergence stage)
__global__ void __launch_bounds__(256)
k_mg_mask(const uint8_t* __restrict__ solid, uint8_t* __restrict__ mask, int w, int h) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w || y >= h) return;
    const ptrdiff_t pos = (ptrdiff_t)y * w + x;
    uint8_t m = 0;
    if (x == 0 || solid[pos - 1]) m |= NB_L;
    if (x == w - 1 || solid[pos + 1]) m |= NB_R;
    if (y == 0 || solid[pos - w]) m |= NB_B;
    if (y == h - 1 || solid[pos + w]) m |= NB_T;
    mask[pos] = m;
}

// residual of one fine cell: rhs - (x1 + x2 + y1 + y2 - 4 p); zero in a solid cell, which carries no equation
// (see oracle mg_v_cycle: its leftover divergence must not reach the coarse grid)
__device__ __forceinline__ float residual_cell(const float* __restrict__ p, const float* __restrict__ rhs,
                                               const uint8_t* __restrict__ mask, const uint8_t* __restrict__ solid, int w,
                                               ptrdiff_t pos) {
    if (solid[pos]) return 0.0f;
    const uint32_t m = mask[pos];
    const float c = p[pos];
    const float x1 = (m & NB_L) ? c : p[pos - 1];
    const float x2 = (m & NB_R) ? c : p[pos + 1];
    const float y1 = (m & NB_B) ? c : p[pos - w];
    const float y2 = (m & NB_T) ? c : p[pos + w];
    const float lap = x1 + x2 + y1 + y2 - 4.0f * c;
    return rhs[pos] - lap;
}

// coarse right-hand side = 4 * mean of the four children's residuals (h -> 2h scales the right-hand side by 4)
__global__ void __launch_bounds__(256)
k_mg_residual_restrict(const float* __restrict__ p, const float* __restrict__ rhs, const uint8_t* __restrict__ mask,
                       const uint8_t* __restrict__ solid, int wf, float* __restrict__ coarse_rhs, int wc, int hc) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= wc || y >= hc) return;
    const ptrdiff_t f = (ptrdiff_t)(2 * y) * wf + 2 * x;
    const float r00 = residual_cell(p, rhs, mask, solid, wf, f), r10 = residual_cell(p, rhs, mask, solid, wf, f + wf);       // (row, col) = (2y, 2x), (2y+1, 2x)
    const float r01 = residual_cell(p, rhs, mask, solid, wf, f + 1), r11 = residual_cell(p, rhs, mask, solid, wf, f + wf + 1);
    coarse_rhs[(ptrdiff_t)y * wc + x] = 4.0f * (0.25f * (r00 + r10 + r01 + r11));
}

// p += cell-centred bilinear interpolation of the coarse correction (rows first, then columns; clamp-to-edge)
__global__ void __launch_bounds__(256)
k_mg_prolong_add(float* __restrict__ p, int wf, int hf, const float* __restrict__ e, int wc, int hc) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= wf || y >= hf) return;
    const int i = y >> 1, j = x >> 1;
    const int i2 = (y & 1) ? min(i + 1, hc - 1) : max(i - 1, 0);
    const int j2 = (x & 1) ? min(j + 1, wc - 1) : max(j - 1, 0);
    const float* e0 = e + (ptrdiff_t)i * wc;
    const float* e1 = e + (ptrdiff_t)i2 * wc;
    const float uc = 0.75f * e0[j] + 0.25f * e1[j];
    const float un = 0.75f * e0[j2] + 0.25f * e1[j2];
    const ptrdiff_t pos = (ptrdiff_t)y * wf + x;
    p[pos] = p[pos] + (0.75f * uc + 0.25f * un);
}

dim3 rows_grid(int w, int h, int bx = 256) { return dim3((unsigned)((w + bx - 1) / bx), (unsigned)h, 1); }

// ---- several red-black sweeps per launch: the temporally blocked form of k_sor_colour.  One tile per block in shared
// memory (pressure, right-hand side, mask: 9 B per cell), updated in place THERE colour by colour with a block barrier between
// the half-sweeps; every half-sweep lets the garbage at the tile's border creep in by one cell, so `sweeps` sweeps
// need a halo of 2 * sweeps cells, recomputed from the neighbouring tiles' cells.  A cell sees exactly the values the
// one-colour-per-launch kernel would show it, in the same operand order: bit-identical, at 13 B per cell of global
// traffic per launch instead of 26 B per cell per sweep.
constexpr int RB_THREADS = 512, RB_WARPS = RB_THREADS / 32, RB_PW = 128;
struct RbParams {
    const float* pin;       // read (tile + halo) ...
    float* pout;            // ... and written (tile) in DIFFERENT buffers: a neighbouring tile of the same launch reads pin
    const float* rhs;
    const uint8_t* mask;
    int w, h;
    int tx, ty, ntx;        // output tile, tiles per row of tiles
    int halo, ph;           // halo cells per side (2 * sweeps); padded rows ty + 2 halo (padded width is RB_PW)
    int sweeps;
    float omega;
};

__global__ void __launch_bounds__(RB_THREADS)
k_sor_tile(const RbParams prm) {
    extern __shared__ __align__(16) unsigned char rb_smem[];
    const int cells = RB_PW * prm.ph;
    float* sp = reinterpret_cast<float*>(rb_smem);
    float* sr = sp + cells;
    uint8_t* sm = reinterpret_cast<uint8_t*>(sr + cells);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tcol = blockIdx.x % prm.ntx, trow = blockIdx.x / prm.ntx;
    const int X0 = tcol * prm.tx - prm.halo, Y0 = trow * prm.ty - prm.halo;      // grid coordinates of padded cell (0, 0)
    // fill: cells outside the grid are zero with every neighbour blocked (they stay zero and nothing in the grid reads
    // them: in-grid edge cells carry the blocked bit for that direction)
    for (int row = warp; row < prm.ph; row += RB_WARPS) {
        const int y = Y0 + row;
        for (int col = lane; col < RB_PW; col += 32) {
            const int x = X0 + col;
            float pv = 0.0f, rv = 0.0f;
            uint8_t mv = 0x0f;
            if (x >= 0 && x < prm.w && y >= 0 && y < prm.h) {
                const ptrdiff_t pos = (ptrdiff_t)y * prm.w + x;
                pv = prm.pin[pos]; rv = prm.rhs[pos]; mv = prm.mask[pos];
            }
            const int o = row * RB_PW + col;
            sp[o] = pv; sr[o] = rv; sm[o] = mv;
        }
    }
    __syncthreads();
    for (int hs = 0; hs < 2 * prm.sweeps; ++hs) {
        const int colour = hs & 1;
        // the padded border row / column has no neighbour in the tile: it is never updated (and never needed)
        for (int row = 1 + warp; row < prm.ph - 1; row += RB_WARPS) {
            const int par = (X0 + Y0 + row + colour) & 1;          // first padded column of this colour in the row
            for (int col = par + 2 * lane; col < RB_PW - 1; col += 64) {
                if (col == 0) continue;
                const int o = row * RB_PW + col;
                const uint32_t m = sm[o];
                const float c = sp[o];
                const float x1 = (m & NB_L) ? c : sp[o - 1];
                const float x2 = (m & NB_R) ? c : sp[o + 1];
                const float y1 = (m & NB_B) ? c : sp[o - RB_PW];
                const float y2 = (m & NB_T) ? c : sp[o + RB_PW];
                const float gs = (x1 + x2 + y1 + y2 - sr[o]) * 0.25f;
                sp[o] = c + prm.omega * (gs - c);
            }
        }
        __syncthreads();
    }
    for (int row = prm.halo + warp; row < prm.halo + prm.ty; row += RB_WARPS) {
        const int y = Y0 + row;
        if (y >= prm.h) break;
        for (int col = prm.halo + lane; col < prm.halo + prm.tx; col += 32) {
            const int x = X0 + col;
            if (x < prm.w) prm.pout[(ptrdiff_t)y * prm.w + x] = sp[row * RB_PW + col];
        }
    }
}

}  // namespace

int launch_sor_sweep(float* p, const float* rhs, const uint8_t* mask, int w, int h, float omega, cudaStream_t st) {
    const dim3 grid((unsigned)(((w + 1) / 2 + SBX_ - 1) / SBX_), (unsigned)((h + SBY_ - 1) / SBY_), 1);
    for (int colour = 0; colour < 2; ++colour)
        k_sor_colour<<<grid, dim3(SBX_, SBY_, 1), 0, st>>>(p, rhs, mask, w, h, colour, omega);
    return 2;
}

// `sweeps` red-black sweeps on *p.  With tiles (default, up to 4 sweeps per launch) every launch reads *p and writes
// *scratch and the two pointers are swapped, so on return *p is the buffer holding the result; NATRIX_SOR_BLOCK=0
// selects the one-colour-per-launch kernels, which work in place.  Returns the kernels launched, -1 on a launch error.
int launch_sor_sweeps(float** p, float** scratch, const float* rhs, const uint8_t* mask, int w, int h, float omega, int sweeps,
                      cudaStream_t st) {
    static const int per_launch = [] { const char* e = getenv("NATRIX_SOR_BLOCK"); return e ? std::min(8, atoi(e)) : 4; }();
    int launched = 0;
    if (per_launch <= 0 || !*scratch) {
        for (int k = 0; k < sweeps; ++k) launched += launch_sor_sweep(*p, rhs, mask, w, h, omega, st);
        return launched;
    }
    static bool attr_set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    constexpr int SMEM_CAP = 96 * 1024;          // two blocks per SM
    if (!attr_set[dev]) {
        if (cudaFuncSetAttribute((const void*)k_sor_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_CAP) != cudaSuccess) return -1;
        attr_set[dev] = true;
    }
    for (int left = sweeps; left > 0;) {
        const int n = std::min(left, per_launch);
        RbParams prm;
        prm.pin = *p; prm.pout = *scratch; prm.rhs = rhs; prm.mask = mask;
        prm.w = w; prm.h = h; prm.sweeps = n; prm.omega = omega;
        prm.halo = 2 * n;
        const int tx_max = RB_PW - 2 * prm.halo;
        prm.ntx = (w + tx_max - 1) / tx_max;
        prm.tx = (w + prm.ntx - 1) / prm.ntx;
        const int ph_max = SMEM_CAP / (RB_PW * 9);
        const int ty_max = std::max(1, ph_max - 2 * prm.halo);
        const int nty = (h + ty_max - 1) / ty_max;
        prm.ty = (h + nty - 1) / nty;
        prm.ph = prm.ty + 2 * prm.halo;
        k_sor_tile<<<prm.ntx * nty, RB_THREADS, (size_t)RB_PW * prm.ph * 9, st>>>(prm);
        if (cudaGetLastError() != cudaSuccess) return -1;
        std::swap(*p, *scratch);
        ++launched;
        left -= n;
    }
    return launched;
}

// ---- multigrid hierarchy: level 0 is the simulator's own pressure / divergence / mask / obstacle map
struct Multigrid {
    struct Level { int w = 0, h = 0; float *p = nullptr, *p2 = nullptr, *rhs = nullptr; uint8_t *solid = nullptr, *mask = nullptr; };
    std::vector<Level> lv;       // lv[0] holds sizes only
};

Multigrid* multigrid_create(int w, int h) {
    Multigrid* mg = new Multigrid();
    Multigrid::Level l0;
    l0.w = w; l0.h = h;
    mg->lv.push_back(l0);
    // oracle mg_levels: coarsen while the smaller side is >= 16 and both sides are even
    while (std::min(mg->lv.back().w, mg->lv.back().h) >= 16 && mg->lv.back().w % 2 == 0 && mg->lv.back().h % 2 == 0) {
        Multigrid::Level l;
        l.w = mg->lv.back().w / 2; l.h = mg->lv.back().h / 2;
        const size_t n = (size_t)l.w * l.h;
        if (cudaMalloc((void**)&l.p, n * 4) != cudaSuccess || cudaMalloc((void**)&l.p2, n * 4) != cudaSuccess ||
            cudaMalloc((void**)&l.rhs, n * 4) != cudaSuccess ||
            cudaMalloc((void**)&l.solid, n) != cudaSuccess || cudaMalloc((void**)&l.mask, n) != cudaSuccess) {
            cudaFree(l.p); cudaFree(l.p2); cudaFree(l.rhs); cudaFree(l.solid); cudaFree(l.mask);
            multigrid_destroy(mg);
            return nullptr;
        }
        mg->lv.push_back(l);
    }
    return mg;
}

void multigrid_destroy(Multigrid* mg) {
    if (!mg) return;
    for (size_t l = 1; l < mg->lv.size(); ++l) {
        cudaFree(mg->lv[l].p); cudaFree(mg->lv[l].p2); cudaFree(mg->lv[l].rhs); cudaFree(mg->lv[l].solid); cudaFree(mg->lv[l].mask);
    }
    delete mg;
}

int multigrid_levels(const Multigrid* mg) { return mg ? (int)mg->lv.size() : 0; }

// `cycles` V(nu, nu) cycles on (*p, rhs) of the full grid; obs = the step's obstacle bytes, mask = its blocked-neighbour
// mask.  *p / *scratch are the simulator's two pressure buffers: the smoother ping-pongs between them and on return *p is
// the one holding the result.  Returns the kernels launched, -1 on a launch error.
int multigrid_solve(Multigrid* mg, float** p, float** scratch, const float* rhs, const uint8_t* obs, const uint8_t* mask,
                    int cycles, int nu, cudaStream_t st) {
    int launched = 0;
    std::vector<Multigrid::Level>& lv = mg->lv;
    lv[0].p = *p; lv[0].p2 = *scratch;
    lv[0].rhs = const_cast<float*>(rhs); lv[0].solid = const_cast<uint8_t*>(obs); lv[0].mask = const_cast<uint8_t*>(mask);
    // the hierarchy of solid maps and masks follows this step's obstacles
    for (size_t l = 1; l < lv.size(); ++l) {
        k_mg_coarsen<<<rows_grid(lv[l].w, lv[l].h), 256, 0, st>>>(lv[l - 1].solid, lv[l - 1].w, lv[l].solid, lv[l].w, lv[l].h);
        k_mg_mask<<<rows_grid(lv[l].w, lv[l].h), 256, 0, st>>>(lv[l].solid, lv[l].mask, lv[l].w, lv[l].h);
        launched += 2;
    }
    auto smooth = [&](int l) {
        const int n = launch_sor_sweeps(&lv[l].p, &lv[l].p2, lv[l].rhs, lv[l].mask, lv[l].w, lv[l].h, 1.0f, nu, st);
        if (n < 0) return false;
        launched += n;
        return true;
    };
    const int last = (int)lv.size() - 1;
    for (int c = 0; c < cycles; ++c) {
        for (int l = 0; l <= last; ++l) {                       // down: smooth, restrict the residual
            if (!smooth(l)) return -1;
            if (l == last) break;
            k_mg_residual_restrict<<<rows_grid(lv[l + 1].w, lv[l + 1].h), 256, 0, st>>>(lv[l].p, lv[l].rhs, lv[l].mask, lv[l].solid, lv[l].w,
                                                                                         lv[l + 1].rhs, lv[l + 1].w, lv[l + 1].h);
            cudaMemsetAsync(lv[l + 1].p, 0, (size_t)lv[l + 1].w * lv[l + 1].h * 4, st);
            launched += 2;
        }
        for (int l = last; l >= 0; --l) {                       // up: (the coarsest level smooths twice in a row), correct, smooth
            if (l < last) {
                k_mg_prolong_add<<<rows_grid(lv[l].w, lv[l].h), 256, 0, st>>>(lv[l].p, lv[l].w, lv[l].h, lv[l + 1].p, lv[l + 1].w, lv[l + 1].h);
                launched += 1;
            }
            if (!smooth(l)) return -1;
        }
    }
    *p = lv[0].p;
    *scratch = lv[0].p2;
    return launched;
}

}  // namespace natrix
