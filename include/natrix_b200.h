/* natrix_b200.h - C ABI of libnatrix_b200.so
 *
 * This is the drop-in boundary for ONE path of fbertola/Natrix: the per-step stable-fluids
 * pipeline that natrix/core/fluid_simulator.py drives through bgfx-python.  Every entry
 * point below replaces a group of bgfx calls made by the reference (cited as
 * "ref: file:line", paths relative to the reference root).  Plain C types only: no
 * torch, no C++ in the signatures; bind it with ctypes / cffi / cgo / JNI.
 *
 * Conventions
 *   - every call returns 0 on success and a negative natrix_status on failure;
 *     natrix_last_error() returns a thread-local, NUL-terminated description;
 *   - handles are owned by the caller and released with natrix_destroy / natrix_dye_destroy
 *     (ref: FluidSimulator.destroy, fluid_simulator.py:476-515);
 *   - a handle is not thread-safe; calls are enqueued on the simulator's CUDA stream and
 *     return before the GPU finishes (the reference is deferred too: bgfx.dispatch only
 *     records, work runs at bgfx.frame()).  natrix_copy_out / natrix_sync / natrix_field_stats
 *     synchronise;
 *   - there is no CPU fallback: without a CUDA device every call fails with
 *     NATRIX_ERR_CUDA.
 *
 * Layout of fields (row-major, idx = y*width + x, float32):
 *   VELOCITY  float2 per cell (x, y)            8 B   ref: fluid_simulator.py:357-361
 *   PRESSURE  float  per cell                   4 B   ref: :362-365 (scalar)
 *   DIVERGENCE, VORTICITY  float per cell       4 B   ref: :366-367
 *   OBSTACLES float2 per cell, (1,0)/(0,1)/0    8 B   ref: :368 (kept as 1 byte/cell in HBM;
 *                                                     expanded on copy_out)
 *   DYE       float  per dye cell               4 B   ref: demo/smooth_particles_area.py:158-162
 */
#ifndef NATRIX_B200_H
#define NATRIX_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct natrix_sim natrix_sim;
typedef struct natrix_dye natrix_dye;

enum natrix_status {
    NATRIX_OK = 0,
    NATRIX_ERR_ARG = -1,     /* bad argument (null handle, size mismatch, unknown id) */
    NATRIX_ERR_CUDA = -2,    /* CUDA runtime / driver failure, message has the CUDA error */
    NATRIX_ERR_STATE = -3,   /* call not valid in the current state */
    NATRIX_ERR_RANGE = -4    /* a slab's halo was too small for this step (multi-GPU only) */
};

enum natrix_field {
    NATRIX_VELOCITY = 0,
    NATRIX_PRESSURE = 1,
    NATRIX_DIVERGENCE = 2,
    NATRIX_VORTICITY = 3,
    NATRIX_OBSTACLES = 4,
    NATRIX_NBMASK = 5,       /* internal blocked-neighbour mask, 1 byte per cell: bits 0-3 = L/R/B/T neighbour is solid
                                or outside the grid, bit 4 = the cell's scaled divergence holds b itself (see DIV4) */
    NATRIX_DIV4 = 6          /* internal: 0.25 * divergence, the copy the Jacobi sweeps read (the sweep then ends in one
                                fused multiply-add, bit-identical to the shader's (sum - b) * 0.25); the field a slab
                                exchanges with its neighbours before the first sweep */
};

enum natrix_option {
    NATRIX_OPT_PIPELINE = 0, /* 0 = one kernel per reference shader (reference dispatch order);
                                1 = fused / temporally blocked kernels (default)            */
    NATRIX_OPT_JACOBI_DEPTH = 1, /* sweeps per launch of the temporally blocked Jacobi kernel */
    NATRIX_OPT_TIMING = 2,   /* 1 = record per-stage CUDA events (natrix_get_timings)        */
    NATRIX_OPT_WARM_START = 3, /* NOT reference behaviour, default 0: 1 = keep the previous
                                step's pressure as the initial guess instead of clearing it            */
    NATRIX_OPT_PACKED = 4,   /* 1 = f32x2 packed arithmetic in the Jacobi kernel              */
    NATRIX_OPT_JACOBI_KERNEL = 5, /* which temporally blocked kernel runs the sweeps of pipeline 1: 0 = auto (default:
                                shared-memory tiles for small, latency-bound grids and for widths TMA cannot
                                address, register streaming behind TMA otherwise), 1 = TMA kernel, 2 = shared-memory
                                kernel.  natrix_get_option reports the kernel in use (1 or 2; 0 under pipeline 0) */
    NATRIX_OPT_SMEM_DEPTH = 6, /* sweeps per launch of the shared-memory Jacobi kernel, 1..16 (default 16)  */
    /* Pressure solvers that are NOT reference behaviour (the reference only has the Jacobi loop of
       fluid_simulator.py:251-255).  Same linear system, full grids only, fused pipeline only; `iterations` then counts
       red-black SOR sweeps / V-cycles.  Checked bit for bit against oracle/natrix_oracle.py rb_sor_sweep / mg_v_cycle. */
    NATRIX_OPT_SOLVER = 7,   /* 0 = Jacobi (default, the reference), 1 = red-black SOR, 2 = multigrid V-cycles   */
    NATRIX_OPT_SOR_OMEGA_MILLI = 8, /* over-relaxation factor x 1000 of solver 1 (default 1900)                     */
    NATRIX_OPT_MG_SMOOTH = 9 /* red-black Gauss-Seidel sweeps before and after each coarse-grid correction (default 2) */
};

/* ---- lifetime ------------------------------------------------------------------------
 * ref: FluidSimulator.__init__ (fluid_simulator.py:37-48): createUniform x15,
 * createProgram x12, createDynamicVertexBuffer x7 (shaders_utils.py:7-12), setBuffer
 * slots 1-7.  All fields start at zero. */
int natrix_create(int width, int height, int device, natrix_sim** out);

/* One row slab of a taller global grid (multi-GPU, one process per GPU).  The slab owns
 * global rows [row0, row0+rows) of a width x global_height grid and keeps `halo` extra
 * rows above and below that the host fills by halo exchange (see natrix_halo_*). */
int natrix_create_slab(int width, int global_height, int row0, int rows, int halo, int device,
                       natrix_sim** out);
int natrix_destroy(natrix_sim* sim);

/* ---- parameters ------------------------------------------------------------------------
 * ref: property setters fluid_simulator.py:58-111 and _update_params :315-336 (setUniform
 * _Speed, _Dissipation, _VorticityScale, _Alpha, _rBeta).  viscosity is a double because the
 * reference derives alpha = 1/viscosity and rBeta = 1/(4+alpha) in Python double before
 * narrowing to c_float; viscosity == 0 skips the viscosity pass (:220). */
int natrix_set_params(natrix_sim* sim, float speed, int iterations, float dissipation,
                      float vorticity, double viscosity, int has_borders);
int natrix_set_option(natrix_sim* sim, int option, int value);
int natrix_get_option(natrix_sim* sim, int option, int* value);

/* ---- impulses and obstacles ------------------------------------------------------------
 * ref: add_velocity :116-131 (shader.AddVelocity.comp), add_circle_obstacle :135-153
 * (shader.AddCircleObstacle.comp), add_triangle_obstacle :156-172
 * (shader.AddTriangleObstacle.comp).  Positions are normalised [0,1]; radius in cells. */
int natrix_add_velocity(natrix_sim* sim, float px, float py, float vx, float vy, float radius);
int natrix_add_circle_obstacle(natrix_sim* sim, float px, float py, float radius, int is_static);
int natrix_add_triangle_obstacle(natrix_sim* sim, float p1x, float p1y, float p2x, float p2y,
                                 float p3x, float p3y, int is_static);

/* ---- the hot path ----------------------------------------------------------------------
 * ref: FluidSimulator.update :174-280: [InitBoundaries] -> AdvectVelocity -> CalcVorticity ->
 * ApplyVorticity -> [Viscosity] -> Divergence -> clear pressure -> iterations x Poisson ->
 * SubtractGradient -> clear obstacles.  On a slab handle with a communicator (natrix_comm_init) the same one
 * call runs the slab's share of the step, halo exchanges included; every rank calls it with the same dt. */
int natrix_step(natrix_sim* sim, float dt);

/* ---- multi-GPU: row slabs with the halo exchange inside the library ---------------------------------
 * New capability (the reference is single-device): one slab handle per GPU - one process or
 * one host thread per GPU - holding the rows natrix_create_slab names, in the standard partition: rank r of
 * N holds H/N rows from r*(H/N) + min(r, H%N), the first H%N ranks one more.  Rank 0 calls
 * natrix_comm_unique_id and hands the 128 bytes to every rank by whatever means the host has (a file, MPI,
 * torch.distributed); then every rank calls natrix_comm_init (collective).  From there natrix_step and
 * natrix_dye_step exchange halo rows with the two neighbouring ranks themselves - NCCL send/recv over
 * NVLink, bound at run time from libnccl.so.2 (NATRIX_NCCL_LIB overrides the path); the pressure exchanges
 * of the Jacobi phase run on a second stream under the interior launches (NATRIX_SLAB_OVERLAP=0 disables).
 * Results are bit-identical to the single-GPU run.  The natrix_step_phase / natrix_halo_region calls below
 * remain for hosts that bring their own transport. */
int natrix_comm_unique_id(void* id128);
int natrix_comm_init(natrix_sim* sim, const void* id128, int rank, int world);
/* exchanges (NCCL groups) issued so far and bytes sent to neighbours by this handle */
int natrix_comm_stats(natrix_sim* sim, unsigned long long* exchanges, unsigned long long* bytes);

/* Multi-GPU pieces of one step (slab handles) for hosts with their own transport.  Phases, in order:
 *   0 advect (needs VELOCITY halo of natrix_halo_rows_needed(sim, 0, dt) rows)
 *   1 vorticity+confinement(+viscosity)+divergence+mask (needs post-advect VELOCITY halo 4)
 *   2 `sweeps` Jacobi sweeps (needs PRESSURE halo `sweeps`, DIVERGENCE+NBMASK halo `sweeps`)
 *   3 subtract gradient + clear obstacles (needs PRESSURE halo 1)
 * Overlapped form of phase 2, for a group of `sweeps` <= halo sweeps between two exchanges:
 *   4 starts the group's interior rows (they need no halo) on the simulator's stream and returns;
 *   5 runs the group's edge rows on the stream natrix_comm_stream returns - where the host has queued
 *     the PRESSURE (first group of a step: DIVERGENCE + NBMASK) exchange of `sweeps` rows between the
 *     two calls - and the rest of the interior on the simulator's stream, which finally waits for the
 *     edges.  4 and 5 come in pairs with the same `sweeps`; the slab needs >= 2 * sweeps rows. */
int natrix_step_phase(natrix_sim* sim, int phase, float dt, int sweeps);
int natrix_halo_rows_needed(natrix_sim* sim, int phase, float dt);
/* Device addresses of the rows to send / to receive for one field:
 * side 0 = towards lower row indices ("up"), 1 = towards higher ("down").
 * send = the slab's own first/last `rows` rows; recv = the halo rows beyond them. */
int natrix_halo_region(natrix_sim* sim, int field, int side, int rows, void** send_ptr,
                       void** recv_ptr, size_t* bytes);

/* ---- field access ------------------------------------------------------------------------
 * The reference exposes only get_velocity_buffer() (:113-114, a bgfx handle) and has no host
 * upload / readback at all.  natrix_field_ptr is the zero-copy equivalent of
 * that handle (device pointer of the CURRENT read buffer; invalidated by the next mutating
 * call); copy_in / copy_out / field_stats are additive extensions used by tests and viewers.
 * For OBSTACLES copy_out/copy_in convert from/to the reference's float2 encoding. */
int natrix_field_ptr(natrix_sim* sim, int field, void** dev_ptr, size_t* bytes);
int natrix_copy_out(natrix_sim* sim, int field, void* host, size_t bytes);
int natrix_copy_in(natrix_sim* sim, int field, const void* host, size_t bytes);
/* out[0..3] = sum, sum of squares, min, max over all components (deterministic order). */
int natrix_field_stats(natrix_sim* sim, int field, double* out4);

/* ---- dye ("smooth particles area") ----------------------------------------------------
 * ref: demo/smooth_particles_area.py:15-211, demo/shaders/shader.AddParticle.comp,
 * shader.AdvectParticle.comp.  The dye grid may have a different resolution from the
 * velocity grid; it reads the simulator's CURRENT velocity and obstacle buffers (the
 * reference relies on the slot-1 / slot-7 bindings the simulator left behind). */
int natrix_dye_create(natrix_sim* sim, int width, int height, natrix_dye** out);
/* Multi-GPU: the rows [row0, row0+rows) of a width x global_height dye grid, with `halo`
 * extra rows either side, attached to a simulator slab.  Before natrix_dye_step the host exchanges
 * natrix_dye_halo_rows_needed(dye, 0, ..) rows of the simulator's VELOCITY (natrix_halo_region; take the
 * maximum over ranks) and (dye, 1, ..) rows of the dye (natrix_dye_halo_region) with both neighbours.
 * field_ptr / copy_out / copy_in / stats / export_rgba8 then cover the slab's own rows only. */
int natrix_dye_create_slab(natrix_sim* sim, int width, int global_height, int row0, int rows, int halo,
                           natrix_dye** out);
int natrix_dye_halo_rows_needed(natrix_dye* dye, int which, float dt, float speed);
int natrix_dye_halo_region(natrix_dye* dye, int side, int rows, void** send_ptr, void** recv_ptr,
                           size_t* bytes);
int natrix_dye_destroy(natrix_dye* dye);
int natrix_dye_add(natrix_dye* dye, float px, float py, float radius, float strength);
int natrix_dye_step(natrix_dye* dye, float dt, float speed, float dissipation);
int natrix_dye_field_ptr(natrix_dye* dye, void** dev_ptr, size_t* bytes);
int natrix_dye_copy_out(natrix_dye* dye, void* host, size_t bytes);
int natrix_dye_copy_in(natrix_dye* dye, const void* host, size_t bytes);
int natrix_dye_stats(natrix_dye* dye, double* out4);
/* the dye buffer as the RGBA8 image the demo renders, one texel per dye cell, all four
 * channels = the dye value converted like an rgba8 unorm imageStore: round(clamp(v, 0, 1) * 255).
 * ref: demo/shaders/demo.ComputeShader.comp:9-21 (imageStore(InputTexture, coord, vec4(v, v, v, v))).
 * `out` may be a device or a host pointer (is_device says which); bytes = width * height * 4. */
int natrix_dye_export_rgba8(natrix_dye* dye, void* out, size_t bytes, int is_device);
/* the frame the demo draws, as a width x height RGBA8 image (top row first) with
 * the framebuffer taken to have the dye grid's size: the rgba8 dye texture above coloured by
 * plasma(fbm(texel)) with alpha = texel (ref: demo/shaders/demo.FieldFragmentShader.frag:21-33,73-99),
 * alpha-blended over the clear colour 0x1a0427ff (ref: demo/simulation_demo.py:94, :249-254) and, when
 * quiver_tile > 0 (the demo offers 8, 16, 32, 64), the white velocity arrows of
 * demo/shaders/demo.QuiverFragmentShader.frag:14-70 blended on top (simulation_demo.py:256-281).
 * Full-grid handles only.  `out` may be a device or a host pointer; bytes = width * height * 4. */
int natrix_render_frame(natrix_dye* dye, void* out, size_t bytes, int is_device, float quiver_tile);

/* ---- synchronisation / introspection ------------------------------------------------- */
int natrix_sync(natrix_sim* sim);
/* CUDA stream the simulator enqueues on (a cudaStream_t), for event timing by the caller. */
int natrix_stream(natrix_sim* sim, void** stream);
/* Second, high-priority stream of a slab handle: halo exchanges queued here overlap the interior
 * Jacobi launches of natrix_step_phase(sim, 4, ..) (see above). */
int natrix_comm_stream(natrix_sim* sim, void** stream);
/* Per-stage milliseconds of the last step when NATRIX_OPT_TIMING is on.  Index:
 * 0 advect, 1 vorticity/confinement/viscosity, 2 divergence, 3 jacobi, 4 gradient, 5 clears. */
int natrix_get_timings(natrix_sim* sim, float* ms, int n);
/* Number of kernels (and memsets) this handle has launched since creation. */
int natrix_launch_count(natrix_sim* sim, unsigned long long* kernels);
/* Host-only introspection (no CUDA call, works without a device): the tile plan of the temporally blocked
 * Jacobi kernel for `depth` sweeps over rows [row0, row1) of a `width`-column grid, given the obstacle hints
 * of a step - nboxes x (x0, x1, y0, y1) half-open boxes, or circles encoded as (cx, -1 - radius, cy, 0).
 * Writes up to `cap` tiles as (strip, first row, end row, diagnostic tag); a strip is 120 (depth <= 4) or 112 output
 * columns wide.  Returns the number of tiles.  Results of a step never depend on the plan. */
int natrix_debug_plan_tiles(int width, int depth, int row0, int row1, const int* boxes, int nboxes,
                            int max_tiles, int* out4, int cap);
/* Tile-plan cache of the temporally blocked Jacobi kernel: launches that reused a plan / that cut and uploaded a
 * new one (a new obstacle set; asynchronous, no allocation on the step path). */
int natrix_debug_plan_stats(natrix_sim* sim, unsigned long long* hits, unsigned long long* misses);
const char* natrix_last_error(void);
const char* natrix_version(void);

#ifdef __cplusplus
}
#endif
#endif /* NATRIX_B200_H */
