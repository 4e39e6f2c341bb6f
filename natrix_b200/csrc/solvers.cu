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

}  // namespace

int launch_sor_sweep(float* p, const float* rhs, const uint8_t* mask, int w, int h, float omega, cudaStream_t st) {
    const dim3 grid((unsigned)(((w + 1) / 2 + SBX_ - 1) / SBX_), (unsigned)((h + SBY_ - 1) / SBY_), 1);
    for (int colour = 0; colour < 2; ++colour)
        k_sor_colour<<<grid, dim3(SBX_, SBY_, 1), 0, st>>>(p, rhs, mask, w, h, colour, omega);
    return 2;
}

// ---- multigrid hierarchy: level 0 is the simulator's own pressure / divergence / mask / obstacle map
struct Multigrid {
    struct Level { int w = 0, h = 0; float *p = nullptr, *rhs = nullptr; uint8_t *solid = nullptr, *mask = nullptr; };
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
        if (cudaMalloc((void**)&l.p, n * 4) != cudaSuccess || cudaMalloc((void**)&l.rhs, n * 4) != cudaSuccess ||
            cudaMalloc((void**)&l.solid, n) != cudaSuccess || cudaMalloc((void**)&l.mask, n) != cudaSuccess) {
            cudaFree(l.p); cudaFree(l.rhs); cudaFree(l.solid); cudaFree(l.mask);
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
        cudaFree(mg->lv[l].p); cudaFree(mg->lv[l].rhs); cudaFree(mg->lv[l].solid); cudaFree(mg->lv[l].mask);
    }
    delete mg;
}

int multigrid_levels(const Multigrid* mg) { return mg ? (int)mg->lv.size() : 0; }

// `cycles` V(nu, nu) cycles on (p, rhs) of the full grid; obs = the step's obstacle bytes, mask = its blocked-neighbour mask
int multigrid_solve(Multigrid* mg, float* p, const float* rhs, const uint8_t* obs, const uint8_t* mask, int cycles, int nu,
                    cudaStream_t st) {
    int launched = 0;
    std::vector<Multigrid::Level>& lv = mg->lv;
    lv[0].p = p; lv[0].rhs = const_cast<float*>(rhs); lv[0].solid = const_cast<uint8_t*>(obs); lv[0].mask = const_cast<uint8_t*>(mask);
    // the hierarchy of solid maps and masks follows this step's obstacles
    for (size_t l = 1; l < lv.size(); ++l) {
        k_mg_coarsen<<<rows_grid(lv[l].w, lv[l].h), 256, 0, st>>>(lv[l - 1].solid, lv[l - 1].w, lv[l].solid, lv[l].w, lv[l].h);
        k_mg_mask<<<rows_grid(lv[l].w, lv[l].h), 256, 0, st>>>(lv[l].solid, lv[l].mask, lv[l].w, lv[l].h);
        launched += 2;
    }
    const int last = (int)lv.size() - 1;
    for (int c = 0; c < cycles; ++c) {
        for (int l = 0; l <= last; ++l) {                       // down: smooth, restrict the residual
            for (int k = 0; k < nu; ++k) launched += launch_sor_sweep(lv[l].p, lv[l].rhs, lv[l].mask, lv[l].w, lv[l].h, 1.0f, st);
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
            for (int k = 0; k < nu; ++k) launched += launch_sor_sweep(lv[l].p, lv[l].rhs, lv[l].mask, lv[l].w, lv[l].h, 1.0f, st);
        }
    }
    return launched;
}

}  // namespace natrix
