// stages_ref.cu - the "reference-order" pipeline: one sm_100a kernel per reference compute
// shader, launched in the reference's dispatch order (NATRIX_OPT_PIPELINE = 0).  It exists so
// that the fused / temporally blocked kernels can be checked bit-for-bit against a plain
// implementation on the GPU at sizes where the CPU oracle is too slow; it is also what the
// impulse / obstacle / utility entry points use.
//
// Differences from the shaders that are NOT observable: integer linear indices,
// 1-byte obstacle map (Q17), slab-relative row addressing.
#include "kernels.h"

namespace natrix {

namespace {

// 256 threads per block as 64 x 4 cells: a warp reads 128/256 contiguous bytes of one row, and the
// rows above / below that the 5-point stencils need are served from L1 within the block
constexpr int BX = 64, BY = 4, BT = BX * BY;

// common.sh:9-19 GetNeighbours with clamp-to-edge on the GLOBAL domain, as local indices
struct Nbr { ptrdiff_t l, r, b, t; };
__device__ __forceinline__ Nbr neighbours(const Geom& g, int x, int ly) {
    const int gy = g.y0 + ly;
    Nbr n;
    n.l = lin(g, max(x - 1, 0), ly);
    n.r = lin(g, min(x + 1, g.w - 1), ly);
    n.b = lin(g, x, max(gy - 1, 0) - g.y0);
    n.t = lin(g, x, min(gy + 1, g.hg - 1) - g.y0);
    return n;
}

#define CELL_PROLOGUE                                        \
    const int x = blockIdx.x * BX + threadIdx.x;             \
    const int ly = r0 + (int)blockIdx.y * BY + threadIdx.y;  \
    if (x >= g.w || ly >= r1) return;                        \
    const ptrdiff_t pos = lin(g, x, ly);

// ref: shader.InitBoundaries.comp:14-34
__global__ void __launch_bounds__(BT) k_init_boundaries(float2* __restrict__ vel, Geom g, int r0, int r1) {
    CELL_PROLOGUE
    const int gy = g.y0 + ly;
    if (x == 0 || x == g.w - 1 || gy == 0 || gy == g.hg - 1) vel[pos] = make_float2(0.0f, 0.0f);
}

// ref: shader.AdvectVelocity.comp:27-50 (arithmetic in common.cuh advect_cell)
template <bool FOLD>
__global__ void __launch_bounds__(BT)
k_advect(const float2* __restrict__ vin, const uint8_t* __restrict__ obs, float2* __restrict__ vout,
         Geom g, int r0, int r1, float dt, float speed, float diss, int* __restrict__ err) {
    CELL_PROLOGUE
    if (obs[pos] != OBS_FREE) { vout[pos] = make_float2(0.0f, 0.0f); return; }
    const int gy = g.y0 + ly;
    vout[pos] = advect_cell<FOLD>(vin, g, x, gy, load_vel<FOLD>(vin, g, x, gy), dt, speed, diss, err);
}

// ref: shader.CalcVorticity.comp:20-26
__global__ void __launch_bounds__(BT)
k_vorticity(const float2* __restrict__ vel, float* __restrict__ vort, Geom g, int r0, int r1) {
    CELL_PROLOGUE
    const Nbr n = neighbours(g, x, ly);
    const float2 vL = vel[n.l], vR = vel[n.r], vB = vel[n.b], vT = vel[n.t];
    vort[pos] = 0.5f * ((vR.y - vL.y) - (vT.x - vB.x));
}

// ref: shader.ApplyVorticity.comp:26-39
__global__ void __launch_bounds__(BT)
k_confinement(const float2* __restrict__ vin, const float* __restrict__ vort, float2* __restrict__ vout,
              Geom g, int r0, int r1, float dt, float scale) {
    CELL_PROLOGUE
    const Nbr n = neighbours(g, x, ly);
    const float2 f = confinement_force(vort[n.l], vort[n.r], vort[n.b], vort[n.t], vort[pos], scale, dt);
    const float2 v = vin[pos];
    vout[pos] = make_float2(v.x + f.x, v.y + f.y);
}

// ref: shader.Viscosity.comp:24-31
__global__ void __launch_bounds__(BT)
k_viscosity(const float2* __restrict__ vin, float2* __restrict__ vout, Geom g, int r0, int r1,
            float alpha, float rbeta) {
    CELL_PROLOGUE
    const Nbr n = neighbours(g, x, ly);
    const float2 x1 = vin[n.l], x2 = vin[n.r], y1 = vin[n.b], y2 = vin[n.t], b = vin[pos];
    float2 o;
    o.x = (x1.x + x2.x + y1.x + y2.x + b.x * alpha) * rbeta;
    o.y = (x1.y + x2.y + y1.y + y2.y + b.y * alpha) * rbeta;
    vout[pos] = o;
}

// ref: shader.Divergence.comp:22-40.  Also emits the blocked-neighbour mask (nullable) that
// the mask-based Jacobi / gradient kernels consume.
__global__ void __launch_bounds__(BT)
k_divergence(const float2* __restrict__ vel, const uint8_t* __restrict__ obs, float* __restrict__ div,
             float* __restrict__ div4, uint8_t* __restrict__ nbmask, Geom g, int r0, int r1) {
    CELL_PROLOGUE
    const Nbr n = neighbours(g, x, ly);
    const bool sL = obs[n.l] != OBS_FREE, sR = obs[n.r] != OBS_FREE;
    const bool sB = obs[n.b] != OBS_FREE, sT = obs[n.t] != OBS_FREE;
    const float x1 = sL ? 0.0f : vel[n.l].x;
    const float x2 = sR ? 0.0f : vel[n.r].x;
    const float y1 = sB ? 0.0f : vel[n.b].y;
    const float y2 = sT ? 0.0f : vel[n.t].y;
    const float b = 0.5f * ((x2 - x1) + (y2 - y1));
    div[pos] = b;
    if (nbmask) {
        const int gy = g.y0 + ly;
        bool raw;
        div4[pos] = scaled_divergence(b, &raw);       // what the temporally blocked Jacobi kernels read (common.cuh NB_RAW)
        uint8_t m = raw ? NB_RAW : 0;
        if (sL || x == 0) m |= NB_L;
        if (sR || x == g.w - 1) m |= NB_R;
        if (sB || gy == 0) m |= NB_B;
        if (sT || gy == g.hg - 1) m |= NB_T;
        nbmask[pos] = m;
    }
}

// div4 and the NB_RAW bit of the mask from div (see common.cuh), for a divergence or a mask the host uploaded
__global__ void __launch_bounds__(256)
k_rescale_divergence(const float* __restrict__ div, float* __restrict__ div4, uint8_t* __restrict__ nbmask, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    bool raw;
    div4[i] = scaled_divergence(div[i], &raw);
    nbmask[i] = (uint8_t)((nbmask[i] & 0x0f) | (raw ? NB_RAW : 0));
}

// ref: shader.Poisson.comp:24-37 (reads the obstacle map like the shader does)
__global__ void __launch_bounds__(BT)
k_poisson_ref(const float* __restrict__ pin, const float* __restrict__ div,
              const uint8_t* __restrict__ obs, float* __restrict__ pout, Geom g, int r0, int r1) {
    CELL_PROLOGUE
    const Nbr n = neighbours(g, x, ly);
    const float p = pin[pos];
    const float x1 = obs[n.l] != OBS_FREE ? p : pin[n.l];
    const float x2 = obs[n.r] != OBS_FREE ? p : pin[n.r];
    const float y1 = obs[n.b] != OBS_FREE ? p : pin[n.b];
    const float y2 = obs[n.t] != OBS_FREE ? p : pin[n.t];
    pout[pos] = (x1 + x2 + y1 + y2 - div[pos]) * 0.25f;
}

// ref: shader.SubtractGradient.comp:24-46
__global__ void __launch_bounds__(BT)
k_gradient_ref(const float2* __restrict__ vin, const float* __restrict__ p,
               const uint8_t* __restrict__ obs, float2* __restrict__ vout, Geom g, int r0, int r1) {
    CELL_PROLOGUE
    const Nbr n = neighbours(g, x, ly);
    const float c = p[pos];
    const float x1 = obs[n.l] != OBS_FREE ? c : p[n.l];
    const float x2 = obs[n.r] != OBS_FREE ? c : p[n.r];
    const float y1 = obs[n.b] != OBS_FREE ? c : p[n.b];
    const float y2 = obs[n.t] != OBS_FREE ? c : p[n.t];
    float2 v = vin[pos];
    v.x = v.x - 0.5f * (x2 - x1);
    v.y = v.y - 0.5f * (y2 - y1);
    vout[pos] = v;
}

// ref: shader.AddVelocity.comp:26-35, applied b.n times in sequence per cell (each application
// includes the all-cell clamp) - identical arithmetic to b.n separate dispatches.
__global__ void __launch_bounds__(BT)
k_add_velocity(const float2* __restrict__ vin, float2* __restrict__ vout, Geom g, int r0, int r1,
               const __grid_constant__ SplatVBatch b) {
    CELL_PROLOGUE
    float2 v = vin[pos];
    const float fxp = (float)x, fyp = (float)(g.y0 + ly);
    for (int i = 0; i < b.n; ++i) {
        const SplatV s = b.s[i];
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
    vout[pos] = v;
}

// ref: shader.AddCircleObstacle.comp:24-36 (x0/y0c: offset of the launched window)
__global__ void __launch_bounds__(BT)
k_add_circle(uint8_t* __restrict__ obs, Geom g, int x0, int xe, int r0, int r1, float sx, float sy,
             float radius) {
    const int x = x0 + blockIdx.x * BX + threadIdx.x;
    const int ly = r0 + (int)blockIdx.y * BY + threadIdx.y;
    if (x >= xe || ly >= r1) return;
    const float ex = sx - (float)x, ey = sy - (float)(g.y0 + ly);
    if (sqrtf(ex * ex + ey * ey) <= radius) obs[lin(g, x, ly)] = OBS_DYNAMIC;
}

// ref: shader.AddTriangleObstacle.comp:19-51
__device__ __forceinline__ float tri_sign(float ax, float ay, float bx, float by, float cx, float cy) {
    return ((ax - cx) * (by - cy)) - ((bx - cx) * (ay - cy));
}
__global__ void __launch_bounds__(BT)
k_add_triangle(uint8_t* __restrict__ obs, Geom g, int r0, int r1, float p1x, float p1y, float p2x,
               float p2y, float p3x, float p3y, int is_static) {
    CELL_PROLOGUE
    const float tx = (float)x / (float)g.w, ty = (float)(g.y0 + ly) / (float)g.hg;
    const bool b1 = tri_sign(tx, ty, p1x, p1y, p2x, p2y) < 0.0f;
    const bool b2 = tri_sign(tx, ty, p2x, p2y, p3x, p3y) < 0.0f;
    const bool b3 = tri_sign(tx, ty, p3x, p3y, p1x, p1y) < 0.0f;
    if (b1 == b2 && b2 == b3) obs[pos] = is_static ? OBS_STATIC : OBS_DYNAMIC;
}

__global__ void k_obs_expand(const uint8_t* __restrict__ obs, float2* __restrict__ out, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint8_t o = obs[i];
    out[i] = make_float2(o == OBS_DYNAMIC ? 1.0f : 0.0f, o == OBS_STATIC ? 1.0f : 0.0f);
}
__global__ void k_obs_pack(const float2* __restrict__ in, uint8_t* __restrict__ obs, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float2 o = in[i];
    obs[i] = o.x > 0.0f ? OBS_DYNAMIC : (o.y > 0.0f ? OBS_STATIC : OBS_FREE);
}

// deterministic two-pass reduction: fixed grid, fixed strides, fixed tree order
constexpr int ST_BLOCKS = 1024, ST_THREADS = 256;
struct Stat { double s, q, lo, hi; };
__device__ __forceinline__ Stat stat_merge(Stat a, Stat b) {
    return Stat{a.s + b.s, a.q + b.q, fmin(a.lo, b.lo), fmax(a.hi, b.hi)};
}
__device__ Stat block_reduce(Stat v) {
    __shared__ Stat sh[ST_THREADS];
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int o = ST_THREADS / 2; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) sh[threadIdx.x] = stat_merge(sh[threadIdx.x], sh[threadIdx.x + o]);
        __syncthreads();
    }
    return sh[0];
}
__global__ void __launch_bounds__(ST_THREADS)
k_stats_partial(const float* __restrict__ d, size_t n, double* __restrict__ scratch) {
    Stat v{0.0, 0.0, INFINITY, -INFINITY};
    for (size_t i = (size_t)blockIdx.x * ST_THREADS + threadIdx.x; i < n; i += (size_t)ST_BLOCKS * ST_THREADS) {
        const double f = (double)d[i];
        v = stat_merge(v, Stat{f, f * f, f, f});
    }
    v = block_reduce(v);
    if (threadIdx.x == 0) {
        scratch[4 * blockIdx.x + 0] = v.s; scratch[4 * blockIdx.x + 1] = v.q;
        scratch[4 * blockIdx.x + 2] = v.lo; scratch[4 * blockIdx.x + 3] = v.hi;
    }
}
__global__ void __launch_bounds__(ST_THREADS)
k_stats_final(const double* __restrict__ scratch, double* __restrict__ out4) {
    Stat v{0.0, 0.0, INFINITY, -INFINITY};
    for (int i = threadIdx.x; i < ST_BLOCKS; i += ST_THREADS)
        v = stat_merge(v, Stat{scratch[4 * i], scratch[4 * i + 1], scratch[4 * i + 2], scratch[4 * i + 3]});
    v = block_reduce(v);
    if (threadIdx.x == 0) { out4[0] = v.s; out4[1] = v.q; out4[2] = v.lo; out4[3] = v.hi; }
}

inline dim3 cell_grid(const Geom& g, int r0, int r1) {
    return dim3((g.w + BX - 1) / BX, (r1 - r0 + BY - 1) / BY, 1);
}
const dim3 CELL_BLOCK(BX, BY, 1);

}  // namespace

#define ROWS_OR_RETURN if (r1 <= r0) return 0;

int launch_init_boundaries(float2* vel, Geom g, int r0, int r1, cudaStream_t st) {
    ROWS_OR_RETURN
    k_init_boundaries<<<cell_grid(g, r0, r1), CELL_BLOCK, 0, st>>>(vel, g, r0, r1);
    return 1;
}
int launch_advect(const float2* vin, const uint8_t* obs, float2* vout, Geom g, int r0, int r1,
                  float dt, float speed, float diss, bool fold, int* err, cudaStream_t st) {
    ROWS_OR_RETURN
    if (fold) k_advect<true><<<cell_grid(g, r0, r1), CELL_BLOCK, 0, st>>>(vin, obs, vout, g, r0, r1, dt, speed, diss, err);
    else k_advect<false><<<cell_grid(g, r0, r1), CELL_BLOCK, 0, st>>>(vin, obs, vout, g, r0, r1, dt, speed, diss, err);
    return 1;
}
int launch_vorticity(const float2* vel, float* vort, Geom g, int r0, int r1, cudaStream_t st) {
    ROWS_OR_RETURN
    k_vorticity<<<cell_grid(g, r0, r1), CELL_BLOCK, 0, st>>>(vel, vort, g, r0, r1);
    return 1;
}
int launch_confinement(const float2* vin, const float* vort, float2* vout, Geom g, int r0, int r1,
                       float dt, float scale, cudaStream_t st) {
    ROWS_OR_RETURN
    k_confinement<<<cell_grid(g, r0, r1), CELL_BLOCK, 0, st>>>(vin, vort, vout, g, r0, r1, dt, scale);
    return 1;
}
int launch_viscosity(const float2* vin, float2* vout, Geom g, int r0, int r1, float alpha, float rbeta,
                     cudaStream_t st) {
    ROWS_OR_RETURN
    k_viscosity<<<cell_grid(g, r0, r1), CELL_BLOCK, 0, st>>>(vin, vout, g, r0, r1, alpha, rbeta);
    return 1;
}
int launch_divergence(const float2* vel, const uint8_t* obs, float* div, float* div4, uint8_t* nbmask, Geom g, int r0,
                      int r1, cudaStream_t st) {
    ROWS_OR_RETURN
    k_divergence<<<cell_grid(g, r0, r1), CELL_BLOCK, 0, st>>>(vel, obs, div, div4, nbmask, g, r0, r1);
    return 1;
}
int launch_rescale_divergence(const float* div, float* div4, uint8_t* nbmask, size_t n, cudaStream_t st) {
    if (n == 0) return 0;
    k_rescale_divergence<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(div, div4, nbmask, n);
    return 1;
}
int launch_poisson_ref(const float* pin, const float* div, const uint8_t* obs, float* pout, Geom g, int r0,
                       int r1, cudaStream_t st) {
    ROWS_OR_RETURN
    k_poisson_ref<<<cell_grid(g, r0, r1), CELL_BLOCK, 0, st>>>(pin, div, obs, pout, g, r0, r1);
    return 1;
}
int launch_gradient_ref(const float2* vin, const float* p, const uint8_t* obs, float2* vout, Geom g, int r0,
                        int r1, cudaStream_t st) {
    ROWS_OR_RETURN
    k_gradient_ref<<<cell_grid(g, r0, r1), CELL_BLOCK, 0, st>>>(vin, p, obs, vout, g, r0, r1);
    return 1;
}
int launch_add_velocity(const float2* vin, float2* vout, Geom g, int r0, int r1, const SplatVBatch& b,
                        cudaStream_t st) {
    ROWS_OR_RETURN
    k_add_velocity<<<cell_grid(g, r0, r1), CELL_BLOCK, 0, st>>>(vin, vout, g, r0, r1, b);
    return 1;
}
int launch_add_circle(uint8_t* obs, Geom g, int r0, int r1, float sx, float sy, float radius, bool bbox,
                      cudaStream_t st) {
    int x0 = 0, xe = g.w;
    if (bbox) {
        // cells farther than radius+2 from the centre along one axis cannot satisfy
        // sqrt(ex^2+ey^2) <= radius (sqrt is monotone and >= |ex| up to one rounding)
        if (!(radius >= 0.0f)) return 0;           // negative or NaN radius marks nothing
        const double m = (double)radius + 2.0;
        const double xlo = (double)sx - m, xhi = (double)sx + m;
        const double ylo = (double)sy - m - g.y0, yhi = (double)sy + m - g.y0;
        if (xhi < 0 || xlo > g.w || yhi < r0 || ylo > r1) return 0;
        x0 = xlo > 0 ? (int)xlo : 0;
        xe = xhi < g.w - 1 ? (int)xhi + 1 : g.w;
        r0 = ylo > r0 ? (int)ylo : r0;
        r1 = yhi < r1 - 1 ? (int)yhi + 1 : r1;
    }
    if (r1 <= r0 || xe <= x0) return 0;
    dim3 grid((xe - x0 + BX - 1) / BX, (r1 - r0 + BY - 1) / BY, 1);
    k_add_circle<<<grid, CELL_BLOCK, 0, st>>>(obs, g, x0, xe, r0, r1, sx, sy, radius);
    return 1;
}
int launch_add_triangle(uint8_t* obs, Geom g, int r0, int r1, float p1x, float p1y, float p2x, float p2y,
                        float p3x, float p3y, int is_static, cudaStream_t st) {
    ROWS_OR_RETURN
    k_add_triangle<<<cell_grid(g, r0, r1), CELL_BLOCK, 0, st>>>(obs, g, r0, r1, p1x, p1y, p2x, p2y, p3x, p3y, is_static);
    return 1;
}
int launch_obs_expand(const uint8_t* obs, float2* out, size_t n, cudaStream_t st) {
    if (!n) return 0;
    k_obs_expand<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(obs, out, n);
    return 1;
}
int launch_obs_pack(const float2* in, uint8_t* obs, size_t n, cudaStream_t st) {
    if (!n) return 0;
    k_obs_pack<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(in, obs, n);
    return 1;
}
int launch_stats(const float* data, size_t n, double* scratch, double* out4_dev, cudaStream_t st) {
    k_stats_partial<<<ST_BLOCKS, ST_THREADS, 0, st>>>(data, n, scratch);
    k_stats_final<<<1, ST_THREADS, 0, st>>>(scratch, out4_dev);
    return 2;
}

}  // namespace natrix
