// kernels.h - host-callable launchers of every CUDA kernel in libnatrix_b200.
// All pointers address LOCAL row 0 of a slab (halo rows live at negative / >= hl indices).
// Row ranges [r0, r1) are local rows.  Every launcher enqueues on `st` and returns the
// number of kernels it launched (for natrix_launch_count).
#pragma once
#include "common.cuh"

namespace natrix {

struct SplatV { float sx, sy, vx, vy, r; };       // add_velocity: splat_pos = pos * size
struct SplatD { float sx, sy, r, value; };        // add_particles
constexpr int MAX_SPLATS = 32;                    // per batched launch
constexpr int MAX_CIRCLES = 64;                   // circles rasterised per launch (fused.cu)
constexpr int OVER_BAND = 64;                     // rows per "some |v| > 1 here" flag (fused.cu)
struct SplatVBatch { int n; SplatV s[MAX_SPLATS]; };
struct SplatDBatch { int n; SplatD s[MAX_SPLATS]; };

// ---- reference-order pipeline: one kernel per reference shader (stages_ref.cu) -------------
int launch_init_boundaries(float2* vel, Geom g, int r0, int r1, cudaStream_t st);
int launch_advect(const float2* vin, const uint8_t* obs, float2* vout, Geom g, int r0, int r1,
                  float dt, float speed, float diss, bool fold_borders, int* err, cudaStream_t st);
int launch_vorticity(const float2* vel, float* vort, Geom g, int r0, int r1, cudaStream_t st);
int launch_confinement(const float2* vin, const float* vort, float2* vout, Geom g, int r0, int r1,
                       float dt, float scale, cudaStream_t st);
int launch_viscosity(const float2* vin, float2* vout, Geom g, int r0, int r1, float alpha,
                     float rbeta, cudaStream_t st);
// div4 / nbmask (nullable together): the pre-scaled divergence and the blocked-neighbour mask (with the NB_RAW bit)
// that the temporally blocked Jacobi kernels and the mask-driven gradient read
int launch_divergence(const float2* vel, const uint8_t* obs, float* div, float* div4, uint8_t* nbmask, Geom g,
                      int r0, int r1, cudaStream_t st);
// div4 and the NB_RAW bit of the mask re-derived from div (after the host uploaded a divergence or a mask)
int launch_rescale_divergence(const float* div, float* div4, uint8_t* nbmask, size_t n, cudaStream_t st);
int launch_poisson_ref(const float* pin, const float* div, const uint8_t* obs, float* pout, Geom g,
                       int r0, int r1, cudaStream_t st);
int launch_gradient_ref(const float2* vin, const float* p, const uint8_t* obs, float2* vout, Geom g,
                        int r0, int r1, cudaStream_t st);
int launch_add_velocity(const float2* vin, float2* vout, Geom g, int r0, int r1,
                        const SplatVBatch& b, cudaStream_t st);
int launch_add_circle(uint8_t* obs, Geom g, int r0, int r1, float sx, float sy, float radius,
                      bool bbox, cudaStream_t st);
int launch_add_triangle(uint8_t* obs, Geom g, int r0, int r1, float p1x, float p1y, float p2x,
                        float p2y, float p3x, float p3y, int is_static, cudaStream_t st);
int launch_obs_expand(const uint8_t* obs, float2* out, size_t n, cudaStream_t st);
int launch_obs_pack(const float2* in, uint8_t* obs, size_t n, cudaStream_t st);
// deterministic sum / sumsq / min / max of n floats; scratch holds >= 4*1024 doubles
int launch_stats(const float* data, size_t n, double* scratch, double* out4_dev, cudaStream_t st);

// ---- dye (dye.cu) ----------------------------------------------------------------------------
// dg: the dye rows held (Geom of the dye grid), vg: the simulator's velocity rows; arrays are row-0 views.
// err: device int raised when a gather leaves the held rows (slabs only).
int launch_dye_add(const float* din, float* dout, Geom dg, int r0, int r1, const SplatDBatch& b, cudaStream_t st);
int launch_dye_advect(const float* din, float* dout, Geom dg, const float2* vel, const uint8_t* obs, Geom vg,
                      float dt, float speed, float diss, int* err, cudaStream_t st);

int launch_dye_rgba8(const float* dye, uint32_t* out, size_t n, cudaStream_t st);
// 4 cells per thread with precomputed normalised-coordinate tables (width % 4 == 0)
int launch_dye_tables(float* nx, float* ny, Geom dg, int vw, int vh, cudaStream_t st);
int launch_dye_advect4(const float* din, float* dout, Geom dg, const float2* vel, const uint8_t* obs, Geom vg,
                       const float* nx, const float* ny, float dt, float speed, float diss, int* err, cudaStream_t st);

// ---- frame rendering (render.cu): the demo's field colour map and quiver overlay
int launch_field_lut(float4* lut, cudaStream_t st);
int launch_render_frame(const float* dye, const float4* lut, const float2* vel, uint32_t* out, int w, int h, int vw,
                        int vh, float tile, cudaStream_t st);

// ---- fused pipeline (fused.cu / jacobi_tb.cu) ----------------------------------------------------
// `depth` (1..16) sweeps pin -> pout for rows [r0, r1), one tile per block in shared memory (jacobi_smem.cu): any
// width, meant for small grids.  Same contract as jacobi_tb_launch; returns launches, or -1 on a launch error.
// grad_vin != null: the launch also subtracts the gradient of its result from grad_vin into grad_vout for the same
// rows (shader.SubtractGradient.comp:24-46) and records |v| > 1 per band in over1 - for the launch with the step's last sweeps.
int launch_jacobi_smem(const float* pin, const float* div4, const uint8_t* nbmask, float* pout, Geom g, int depth,
                       int r0, int r1, bool p_is_zero, int sm_count, cudaStream_t st, const float2* grad_vin = nullptr,
                       float2* grad_vout = nullptr, int* over1 = nullptr);
int jacobi_smem_max_depth();             // default sweeps per launch (NATRIX_SMEM_DEPTH, <= 16)
size_t jacobi_smem_cell_limit();         // grids up to this many cells prefer the shared-memory kernel
// over1: one device int per band of OVER_BAND allocated rows, set when some |v| > 1 is written there
int launch_gradient_mask(const float2* vin, const float* p, const uint8_t* nbmask, float2* vout,
                         Geom g, int r0, int r1, int* over1, cudaStream_t st);
// fused advect + vorticity + confinement + [viscosity] + divergence + mask, rows [r0, r1)
bool preproject_supported(const Geom& g);
int launch_preproject(const float2* vin, const uint8_t* obs, float2* vout, float* vort, float* div, float* div4,
                      uint8_t* nbmask, Geom g, int r0, int r1, float dt, float speed, float diss, float scale,
                      bool viscous, float alpha, float rbeta, int sm_count, int* err, cudaStream_t st);
// InitBoundaries restricted to the border cells of rows [r0, r1), in place
int launch_zero_borders(float2* vel, Geom g, int r0, int r1, cudaStream_t st);
// in-place impulses restricted to the splats' bounding boxes (n <= MAX_SPLATS)
int launch_splat_velocity_boxes(float2* vel, Geom g, int r0, int r1, const SplatV* splats, int n,
                                const int* over1, int sm_count, cudaStream_t st);
int launch_splat_dye_boxes(float* dye, Geom dg, int r0, int r1, const SplatD* splats, int n, cudaStream_t st);
// n queued circles (sx, sy, radius triples, in cells) rasterised on their bounding boxes, rows [r0, r1)
int launch_add_circles(uint8_t* obs, Geom g, int r0, int r1, const float* sxyr, int n, cudaStream_t st);

// ---- opt-in solvers that are not reference behaviour (solvers.cu); full grids only
// one red-black SOR sweep (both colours) in place on p; returns the kernels launched
int launch_sor_sweep(float* p, const float* rhs, const uint8_t* mask, int w, int h, float omega, cudaStream_t st);
// `sweeps` sweeps, up to 4 per launch in shared-memory tiles: every launch reads *p and writes *scratch and swaps the two,
// so on return *p holds the result (NATRIX_SOR_BLOCK=0 or a null *scratch: the in-place one-colour kernels).  -1 on error.
int launch_sor_sweeps(float** p, float** scratch, const float* rhs, const uint8_t* mask, int w, int h, float omega, int sweeps,
                      cudaStream_t st);
struct Multigrid;
Multigrid* multigrid_create(int w, int h);        // nullptr when device memory ran out
void multigrid_destroy(Multigrid* mg);
int multigrid_levels(const Multigrid* mg);
// `cycles` V(nu, nu) cycles with a red-black Gauss-Seidel smoother on the full grid's (*p, rhs); *p / *scratch are the two
// pressure buffers (on return *p holds the result); obs / mask are the step's obstacle bytes and blocked-neighbour mask.
// Returns the kernels launched, -1 on a launch error.
int multigrid_solve(Multigrid* mg, float** p, float** scratch, const float* rhs, const uint8_t* obs, const uint8_t* mask,
                    int cycles, int nu, cudaStream_t st);

}  // namespace natrix
