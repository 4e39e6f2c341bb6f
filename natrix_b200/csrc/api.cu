// api.cu - the extern "C" boundary of libnatrix_b200.so (see include/natrix_b200.h) and the
// orchestration of one simulation step (ref: FluidSimulator.update, fluid_simulator.py:174-280).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <algorithm>
#include <string>
#include <vector>

#include <dlfcn.h>
#include <nvtx3/nvToolsExt.h>     // header-only NVTX 3: a few ns per call unless a profiler is attached

#include "kernels.h"
#include "jacobi_tb.h"

using namespace natrix;

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}

#define CU(expr)                                                                         \
    do {                                                                                 \
        cudaError_t e__ = (expr);                                                        \
        if (e__ != cudaSuccess)                                                          \
            return fail(NATRIX_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__)); \
    } while (0)

#define NEED(cond, msg) \
    do { if (!(cond)) return fail(NATRIX_ERR_ARG, msg); } while (0)

// ---- NCCL, bound at run time: the library has no link-time dependency on it, and a process that already
// holds a libnccl.so.2 (torch's bundled one) keeps using that copy.  Only the few entry points of the halo
// exchange are bound; the prototypes follow nccl.h (ncclUniqueId is a 128-byte struct passed by value).
struct NcclId { char internal[128]; };
struct Nccl {
    void* h = nullptr;
    int (*GetUniqueId)(NcclId*) = nullptr;
    int (*CommInitRank)(void**, int, NcclId, int) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*Send)(const void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    std::string why;
};
constexpr int NCCL_UINT8 = 1;        // ncclUint8 (nccl.h)

Nccl* nccl() {
    static Nccl n;
    if (n.h || !n.why.empty()) return &n;
    const char* env = getenv("NATRIX_NCCL_LIB");
    if (env) n.h = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
    if (!n.h) n.h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);     // the copy this process already uses
    if (!n.h) n.h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!n.h) n.h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!n.h) { n.why = std::string("libnccl.so.2 not found (set NATRIX_NCCL_LIB): ") + dlerror(); return &n; }
    auto sym = [&](const char* name) { void* p = dlsym(n.h, name); if (!p) n.why = std::string("missing NCCL symbol ") + name; return p; };
    n.GetUniqueId = (int (*)(NcclId*))sym("ncclGetUniqueId");
    n.CommInitRank = (int (*)(void**, int, NcclId, int))sym("ncclCommInitRank");
    n.CommDestroy = (int (*)(void*))sym("ncclCommDestroy");
    n.Send = (int (*)(const void*, size_t, int, int, void*, cudaStream_t))sym("ncclSend");
    n.Recv = (int (*)(void*, size_t, int, int, void*, cudaStream_t))sym("ncclRecv");
    n.GroupStart = (int (*)())sym("ncclGroupStart");
    n.GroupEnd = (int (*)())sym("ncclGroupEnd");
    n.GetErrorString = (const char* (*)(int))sym("ncclGetErrorString");
    if (!n.why.empty()) n.h = nullptr;
    return &n;
}

#define NC(expr)                                                                              \
    do {                                                                                      \
        int r__ = (expr);                                                                     \
        if (r__ != 0)                                                                         \
            return fail(NATRIX_ERR_CUDA, std::string(#expr) + ": " + nccl()->GetErrorString(r__)); \
    } while (0)

// NVTX range around one stage of the step (tracing); shows up per stage in Nsight Systems / ncu --nvtx
struct Range {
    explicit Range(const char* name) { nvtxRangePushA(name); }
    ~Range() { nvtxRangePop(); }
};

enum Stage { ST_ADVECT = 0, ST_VORT, ST_DIV, ST_JACOBI, ST_GRAD, ST_CLEAR, ST_COUNT };

}  // namespace

struct natrix_sim {
    Geom g{};
    int device = 0;
    cudaStream_t st = nullptr;
    size_t rows_alloc = 0, cells_alloc = 0;     // including halos
    // allocations (base) and row-0 views
    float2* vel_base[2] = {nullptr, nullptr};
    float* p_base[2] = {nullptr, nullptr};
    float *div_base = nullptr, *vort_base = nullptr, *div4_base = nullptr;
    uint8_t *obs_base = nullptr, *nbm_base = nullptr;
    float2* vel[2] = {nullptr, nullptr};
    float* p[2] = {nullptr, nullptr};
    float *div = nullptr, *vort = nullptr;
    float* div4 = nullptr;                       // 0.25 * div, what the temporally blocked Jacobi kernels read (common.cuh NB_RAW)
    uint8_t *obs = nullptr, *nbm = nullptr;
    int vr = 0, pr = 0;                          // VELOCITY_READ / PRESSURE_READ indices
    // parameters (fluid_simulator.py:28-35 defaults)
    float speed = 500.0f, dissipation = 1.0f, vorticity = 0.0f;
    float alpha = (float)(1.0 / 0.1), rbeta = (float)(1.0 / (4.0 + 1.0 / 0.1));
    int iterations = 50, has_borders = 1, viscous = 1;
    // options
    int pipeline = 1, jacobi_depth = 8, timing = 0, packed = 1, warm_start = 0;
    // solvers that are not reference behaviour (solvers.cu): 0 Jacobi (the reference), 1 red-black SOR, 2 multigrid
    int solver = 0, sor_omega_milli = 1900, mg_smooth = 2;
    Multigrid* mg = nullptr;
    int jacobi_kernel = 0;                       // NATRIX_OPT_JACOBI_KERNEL: 0 auto, 1 TMA register streaming, 2 shared memory
    int smem_depth = 0;                          // sweeps per launch of the shared-memory kernel (set at create)
    // gradient subtraction as the epilogue of the step's last shared-memory Jacobi launch (full grids; NATRIX_SMEM_GRAD=0
    // turns it off): natrix_step asks for it, phase_jacobi does it, phase_project then skips its own launch
    int smem_grad = 1;
    bool grad_wanted = false, grad_done = false;
    // bookkeeping
    std::vector<SplatV> pending;                 // add_velocity calls not yet applied (pipeline 1)
    std::vector<float> circles;                  // queued add_circle_obstacle calls (sx, sy, r), pipeline 1
    std::vector<int> heavy;                      // merged [lo, hi) local-row intervals stamped with obstacles this step
    std::vector<int> boxes;                      // (x0, x1, y0, y1) per obstacle stamped this step (scheduling hint)
    bool obs_dirty = false, p_is_zero = false, fused_pre = false;
    bool first_block = false;                    // no Jacobi launch of this step has been queued yet
    int* d_err = nullptr;                        // [0] advection left the slab's halo; [1..] per band of OVER_BAND
                                                 // rows: some |v| > 1 in the READ velocity
    int nbands = 0;
    int* h_err = nullptr;
    cudaEvent_t err_event = nullptr;
    bool err_pending = false;
    int sm_count = 148;
    double *d_scratch = nullptr, *d_out4 = nullptr, *h_out4 = nullptr;
    float2* d_tmp2 = nullptr;                    // staging for OBSTACLES copy in/out
    unsigned long long launches = 0;
    cudaEvent_t ev[ST_COUNT + 1] = {};
    float stage_ms[ST_COUNT] = {};
    JacobiTB* tb = nullptr;
    std::vector<natrix_dye*> dyes;
    // slabs, overlapped halo exchange: the exchange and the edge zones of a Jacobi group run on st_edge
    // while the interior runs on st (phase_jacobi_interior / phase_jacobi_edges)
    cudaStream_t st_edge = nullptr;
    cudaEvent_t ev_group = nullptr, ev_edges = nullptr;
    int group_open = 0;                          // sweeps of the group whose interior is queued, else 0
    // halo rows of the READ velocity that were filled (natrix_halo_region / the library's own exchange) since
    // that buffer was last written: the range check of the back-traces compares against these, not against
    // the allocation
    int vel_halo_valid = 0;
    // the library's own halo exchange (natrix_comm_init): NCCL communicator over the slabs, rank order = row order
    void* comm = nullptr;
    int comm_rank = -1, comm_world = 0;
    int overlap = 1;                             // NATRIX_SLAB_OVERLAP
    int xfirst = 0, reserve_sms = 0;             // NATRIX_SLAB_XFIRST, NATRIX_SLAB_RESERVE (see step_slab)
    unsigned long long exchanges = 0, exchanged_bytes = 0;

    int ext_lo(int k) const { int lo = -k; if (g.y0 + lo < 0) lo = -g.y0; return lo < -g.halo ? -g.halo : lo; }
    int ext_hi(int k) const {
        int hi = g.hl + k;
        if (g.y0 + hi > g.hg) hi = g.hg - g.y0;
        return hi > g.hl + g.halo ? g.hl + g.halo : hi;
    }
};

struct natrix_dye {
    natrix_sim* sim = nullptr;
    Geom g{};                                   // dye rows held: w, global height, row0, rows, halo
    float* base[2] = {nullptr, nullptr};        // allocations (halo rows included)
    float* d[2] = {nullptr, nullptr};           // row-0 views
    float* tables = nullptr;                    // normalised x (w floats) then y (one per own row) coordinates
    uint32_t* rgba = nullptr;                   // staging for host RGBA8 export
    float4* lut = nullptr;                      // 256-entry field colour map (render.cu), built on first use
    int rd = 0;
    int halo_valid = 0;                         // halo rows of the READ dye buffer filled since it was last written
    std::vector<SplatD> pending;

    size_t own_cells() const { return (size_t)g.w * g.hl; }
    // rows of the global dye grid this handle holds, as local indices (halo clipped at the grid's ends)
    int held_lo() const { return g.y0 - g.halo < 0 ? -g.y0 : -g.halo; }
    int held_hi() const { return g.y0 + g.hl + g.halo > g.hg ? g.hg - g.y0 : g.hl + g.halo; }
};

namespace {

template <class T>
cudaError_t alloc_rows(T** base, T** view, const natrix_sim* s) {
    cudaError_t e = cudaMalloc((void**)base, s->cells_alloc * sizeof(T));
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(*base, 0, s->cells_alloc * sizeof(T), s->st);
    *view = *base + (size_t)s->g.halo * s->g.w;
    return e;
}

// remember which rows carry obstacles (scheduling hint for the Jacobi kernel); keeps the list merged
void mark_heavy_rows(natrix_sim* s, double glo, double ghi, double xlo = -1e30, double xhi = 1e30, bool circle = false) {
    const int margin = 2;
    const int by0 = std::max((int)std::floor(std::max(glo, -1e9)) - margin - s->g.y0, -s->g.halo);
    const int by1 = std::min((int)std::ceil(std::min(ghi, 1e9)) + margin + 1 - s->g.y0, s->g.hl + s->g.halo);
    if (by1 <= by0) return;                          // does not touch the rows this slab holds
    if (circle && 0.5 * (ghi - glo) < 1e6 && s->boxes.size() < 4 * 256) {
        // a circle keeps its geometry: (cx, -1 - r, cy local, 0), see jacobi_tb.h
        const double r = 0.5 * (ghi - glo);
        s->boxes.insert(s->boxes.end(), {(int)std::lround(0.5 * (xlo + xhi)), -1 - (int)std::ceil(r),
                                         (int)std::lround(0.5 * (glo + ghi)) - s->g.y0, 0});
    } else {
        const int bx0 = (int)std::max(0.0, std::floor(xlo) - margin), bx1 = (int)std::min((double)s->g.w, std::ceil(xhi) + margin + 1);
        if (bx1 > bx0) {
            if (s->boxes.size() >= 4 * 256) {        // too many to plan around: one box over everything
                s->boxes = {0, s->g.w, -s->g.halo, s->g.hl + s->g.halo};
            } else {
                s->boxes.insert(s->boxes.end(), {bx0, bx1, by0, by1});
            }
        }
    }
    // (clamped in double first: a radius beyond 2^31 must not overflow the casts)
    int lo = (int)std::floor(std::max(glo, -1e9)) - margin - s->g.y0, hi = (int)std::ceil(std::min(ghi, 1e9)) + margin + 1 - s->g.y0;
    lo = std::max(lo, -s->g.halo);
    hi = std::min(hi, s->g.hl + s->g.halo);
    if (hi <= lo) return;
    std::vector<int> out;
    bool placed = false;
    for (size_t k = 0; k + 1 < s->heavy.size(); k += 2) {
        int a = s->heavy[k], b = s->heavy[k + 1];
        if (b < lo) { out.push_back(a); out.push_back(b); }
        else if (a > hi) {
            if (!placed) { out.push_back(lo); out.push_back(hi); placed = true; }
            out.push_back(a); out.push_back(b);
        } else { lo = std::min(lo, a); hi = std::max(hi, b); }
    }
    if (!placed) { out.push_back(lo); out.push_back(hi); }
    if (out.size() > 64) { out = {out.front(), out.back()}; }      // too fragmented: one interval
    s->heavy.swap(out);
}

int select_device(const natrix_sim* s) {
    CU(cudaSetDevice(s->device));
    return 0;
}

void stamp(natrix_sim* s, int i) {
    if (s->timing) cudaEventRecord(s->ev[i], s->st);
}

// apply queued add_velocity calls in one pass (pipeline 1) - identical per-cell arithmetic
int flush_splats(natrix_sim* s) {
    Range nvtx_range("natrix.add_velocity");
    size_t i = 0;
    const int lo = s->ext_lo(s->g.halo), hi = s->ext_hi(s->g.halo);
    while (i < s->pending.size()) {
        if (s->pipeline == 0) {
            // one full-grid dispatch + ping-pong flip per call, like the reference
            SplatVBatch b;
            b.n = 1;
            b.s[0] = s->pending[i++];
            s->launches += launch_add_velocity(s->vel[s->vr], s->vel[1 - s->vr], s->g, lo, hi, b, s->st);
            s->vr = 1 - s->vr;
        } else {
            // up to MAX_SPLATS calls in one in-place pass over their bounding boxes; cells outside
            // only need the all-cell clamp, and only if some |v| > 1 (d_err[1])
            const int n = (int)std::min<size_t>(MAX_SPLATS, s->pending.size() - i);
            s->launches += launch_splat_velocity_boxes(s->vel[s->vr], s->g, lo, hi, &s->pending[i], n, s->d_err + 1,
                                                       s->sm_count, s->st);
            CU(cudaMemsetAsync(s->d_err + 1, 0, s->nbands * sizeof(int), s->st));      // every |v| <= 1 now
            i += n;
        }
    }
    s->pending.clear();
    CU(cudaGetLastError());
    return 0;
}

// rasterise the queued circles in one launch (pipeline 1); must run before anything reads or
// overwrites the obstacle map
int flush_circles(natrix_sim* s) {
    if (s->circles.empty()) return 0;
    s->launches += launch_add_circles(s->obs, s->g, s->ext_lo(s->g.halo), s->ext_hi(s->g.halo), s->circles.data(),
                                      (int)s->circles.size() / 3, s->st);
    s->circles.clear();
    CU(cudaGetLastError());
    return 0;
}

int flush_dye(natrix_dye* d) {
    natrix_sim* s = d->sim;
    size_t i = 0;
    const int lo = d->held_lo(), hi = d->held_hi();      // halo rows too: both neighbours apply the same splats
    while (i < d->pending.size()) {
        if (s->pipeline == 0) {
            SplatDBatch b;
            b.n = 1;
            b.s[0] = d->pending[i++];
            s->launches += launch_dye_add(d->d[d->rd], d->d[1 - d->rd], d->g, lo, hi, b, s->st);
            d->rd = 1 - d->rd;
        } else {
            const int n = (int)std::min<size_t>(MAX_SPLATS, d->pending.size() - i);
            s->launches += launch_splat_dye_boxes(d->d[d->rd], d->g, lo, hi, &d->pending[i], n, s->st);
            i += n;
        }
    }
    d->pending.clear();
    CU(cudaGetLastError());
    return 0;
}

// Slabs only (the full grid never sets the flag): did an advection back-trace leave the rows this
// slab holds?  The flag is fetched asynchronously after every step and looked at when the copy has
// landed - at the latest by the next step or natrix_sync - so the check never drains the stream.
int check_range_flag(natrix_sim* s, bool wait) {
    if (s->g.hl == s->g.hg) return 0;
    if (s->err_pending) {
        if (wait) CU(cudaEventSynchronize(s->err_event));
        const cudaError_t q = cudaEventQuery(s->err_event);
        if (q == cudaSuccess) {
            s->err_pending = false;
            if (*s->h_err) {
                *s->h_err = 0;
                CU(cudaMemsetAsync(s->d_err, 0, sizeof(int), s->st));
                return fail(NATRIX_ERR_RANGE, "a back-trace (velocity or dye advection) left the slab's halo rows; enlarge the halo");
            }
        } else if (q != cudaErrorNotReady) {
            CU(q);
        }
    }
    if (!s->err_pending && !wait) {
        CU(cudaMemcpyAsync(s->h_err, s->d_err, sizeof(int), cudaMemcpyDeviceToHost, s->st));
        CU(cudaEventRecord(s->err_event, s->st));
        s->err_pending = true;
    }
    return 0;
}

// ---- the four phases of a step -------------------------------------------------------------
int phase_advect(natrix_sim* s, float dt) {
    Range nvtx_range("natrix.pre_projection");
    const Geom& g = s->g;
    if (int rc = flush_splats(s)) return rc;
    if (int rc = flush_circles(s)) return rc;
    stamp(s, ST_ADVECT);
    const bool fold = s->pipeline != 0 && s->has_borders;
    // the back-traces may only gather from halo rows that were filled for THIS velocity buffer (the kernels use
    // Geom::halo for nothing but that range check): rows beyond them hold stale ping-pong data
    Geom gv = g;
    gv.halo = std::min(g.halo, s->vel_halo_valid);
    s->vel_halo_valid = 0;                       // the buffer that becomes READ has no exchanged rows yet
    if (g.hl != g.hg && gv.halo < std::min(4, g.halo) && (g.y0 > 0 || g.y0 + g.hl < g.hg))
        return fail(NATRIX_ERR_STATE, "the slab's VELOCITY halo was not exchanged before the step (natrix_halo_rows_needed rows)");
    s->fused_pre = s->pipeline != 0 && preproject_supported(g);
    if (s->fused_pre) {
        // advect + vorticity + confinement + [viscosity] + divergence + mask in one pass; the
        // velocity buffer flips once (the intermediate velocities never reach memory)
        if (s->has_borders)
            s->launches += launch_zero_borders(s->vel[s->vr], g, s->ext_lo(g.halo), s->ext_hi(g.halo), s->st);
        s->launches += launch_preproject(s->vel[s->vr], s->obs, s->vel[1 - s->vr], s->vort, s->div, s->div4, s->nbm, gv, 0, g.hl,
                                         dt, s->speed, s->dissipation, s->vorticity, s->viscous != 0, s->alpha,
                                         s->rbeta, s->sm_count, s->d_err, s->st);
        s->vr = 1 - s->vr;
        CU(cudaGetLastError());
        return 0;
    }
    if (s->has_borders && !fold)
        s->launches += launch_init_boundaries(s->vel[s->vr], g, s->ext_lo(g.halo), s->ext_hi(g.halo), s->st);
    s->launches += launch_advect(s->vel[s->vr], s->obs, s->vel[1 - s->vr], gv, s->ext_lo(4), s->ext_hi(4), dt,
                                 s->speed, s->dissipation, fold, s->d_err, s->st);
    s->vr = 1 - s->vr;
    CU(cudaGetLastError());
    return 0;
}

int phase_forces(natrix_sim* s, float dt) {
    const Geom& g = s->g;
    stamp(s, ST_VORT);
    s->first_block = true;
    if (s->fused_pre) {
        stamp(s, ST_DIV);
        // clear pressure (fluid_simulator.py:236-248): the first temporally blocked launch (either kernel) treats
        // p as zero without reading it, so no fill is needed.
        // NATRIX_OPT_WARM_START keeps the previous step's pressure as the initial guess instead.
        s->p_is_zero = !s->warm_start;
        return 0;
    }
    s->launches += launch_vorticity(s->vel[s->vr], s->vort, g, s->ext_lo(3), s->ext_hi(3), s->st);
    s->launches += launch_confinement(s->vel[s->vr], s->vort, s->vel[1 - s->vr], g, s->ext_lo(2), s->ext_hi(2),
                                      dt, s->vorticity, s->st);
    s->vr = 1 - s->vr;
    if (s->viscous) {
        s->launches += launch_viscosity(s->vel[s->vr], s->vel[1 - s->vr], g, s->ext_lo(1), s->ext_hi(1),
                                        s->alpha, s->rbeta, s->st);
        s->vr = 1 - s->vr;
    }
    stamp(s, ST_DIV);
    s->launches += launch_divergence(s->vel[s->vr], s->obs, s->div, s->div4, s->nbm, g, 0, g.hl, s->st);
    // clear pressure (fluid_simulator.py:236-248); halo rows included so the first Jacobi
    // block needs no exchange
    if (!s->warm_start) {
        CU(cudaMemsetAsync(s->p_base[s->pr], 0, s->cells_alloc * sizeof(float), s->st));
        s->launches += 1;
    }
    s->p_is_zero = !s->warm_start;
    CU(cudaGetLastError());
    return 0;
}

// Which temporally blocked kernel runs the sweeps of pipeline 1: the shared-memory one (jacobi_smem.cu) for
// grids small enough to be latency-bound and for widths TMA cannot address, else the register-streaming TMA
// kernel (jacobi_tb.cu).  NATRIX_OPT_JACOBI_KERNEL forces either (tests, A/B measurements).
bool use_smem_kernel(const natrix_sim* s) {
    if (s->pipeline == 0) return false;
    if (s->jacobi_kernel == 2 || !jacobi_tb_supported(s->g)) return true;
    if (s->jacobi_kernel == 1) return false;
    return s->g.hl == s->g.hg && s->cells_alloc <= jacobi_smem_cell_limit();
}

int jacobi_launch_depth(const natrix_sim* s) {
    if (s->pipeline == 0) return 1;
    if (use_smem_kernel(s))      // slabs keep one depth for both kernels: the exchange schedule is built on it
        return s->g.hl == s->g.hg ? s->smem_depth : std::min(s->smem_depth, s->jacobi_depth);
    return s->jacobi_depth;
}

// `sweeps` Jacobi sweeps; requires p, div, nbmask valid on ext(sweeps) (exchange done by caller)
// One launch of `depth` Jacobi sweeps over local rows [r0, r1): p[src] -> p[1 - src] on stream st.
int jacobi_rows(natrix_sim* s, int src, int depth, int r0, int r1, bool p_zero, cudaStream_t st, bool with_gradient = false) {
    if (r1 <= r0) return 0;
    const Geom& g = s->g;
    if (s->pipeline == 0) {
        s->launches += launch_poisson_ref(s->p[src], s->div, s->obs, s->p[1 - src], g, r0, r1, st);
    } else if (use_smem_kernel(s)) {
        // small grids (latency-bound) and widths TMA cannot address: one tile per block in shared memory
        const int n = launch_jacobi_smem(s->p[src], s->div4, s->nbm, s->p[1 - src], g, depth, r0, r1, p_zero, s->sm_count, st,
                                         with_gradient ? s->vel[s->vr] : nullptr, with_gradient ? s->vel[1 - s->vr] : nullptr,
                                         s->d_err + 1);
        if (n < 0) return fail(NATRIX_ERR_CUDA, std::string("k_jacobi_smem launch: ") + cudaGetErrorString(cudaGetLastError()));
        s->launches += n;
    } else {
        int n = jacobi_tb_launch(s->tb, s->p[src], s->div4, s->nbm, s->p[1 - src], g, depth, r0, r1, p_zero, s->packed,
                                 s->boxes.data(), (int)s->boxes.size() / 4, st);
        if (n < 0) return fail(NATRIX_ERR_CUDA, std::string("jacobi_tb: ") + jacobi_tb_error(s->tb));
        s->launches += n;
    }
    return 0;
}

int phase_jacobi(natrix_sim* s, int sweeps) {
    Range nvtx_range("natrix.jacobi");
    int left = sweeps;
    while (left > 0) {
        int t = std::min(left, jacobi_launch_depth(s));
        if (use_smem_kernel(s)) {
            // the fewest launches the depth allows, sweeps spread evenly over them (50 = 13 + 13 + 12 + 12)
            const int launches = (left + t - 1) / t;
            t = (left + launches - 1) / launches;
        }
        // the launch with the step's last sweeps takes the gradient subtraction along (shared-memory kernel, full grid)
        const bool grad = s->grad_wanted && left == t && use_smem_kernel(s) && s->g.hl == s->g.hg;
        if (int rc = jacobi_rows(s, s->pr, t, s->ext_lo(left - t), s->ext_hi(left - t), s->p_is_zero, s->st, grad)) return rc;
        if (grad) s->grad_done = true;
        s->pr = 1 - s->pr;
        left -= t;
        s->p_is_zero = false;
    }
    CU(cudaGetLastError());
    return 0;
}

// ---- overlapped halo exchange (slabs) ----------------------------------------------------------------
// A group of `sweeps` Jacobi sweeps between two pressure exchanges.  Only its FIRST launch is cut by rows:
// the interior - the rows whose value after that launch depends on no halo row - runs on the main stream
// beside the exchange, the two edge zones on st_edge behind the exchange; the remaining launches of the
// group cover all rows in one piece on the main stream, which waits for the edge zones first.
//
//   launch 1 (depth d_1):  interior I_1 = [d_1, hl - d_1)                          p[src] -> p[1 - src]
//                          edges    B_1 = [ext_lo(sweeps - d_1), d_1) and the mirror image
//   launch j > 1:          F_j = [ext_lo(sweeps - c_j), ext_hi(sweeps - c_j)),  c_j = d_1 + .. + d_j
//
//   main    | I_1 ....................... | wait E | F_2 ........ | F_3 ........ |  next group
//   st_edge | wait G | exchange X | B_1 | E
//
// I_1 reads own rows of p[src] only and writes rows of p[1 - src] that B_1 does not; X writes halo rows of
// p[src] (and of the divergence and the mask in the first group of a step) that I_1 never reads.  F_2 writes
// p[src], the buffer X sends from and B_1 reads: it is ordered behind both through E.  (An earlier schedule
// cut every launch of the group: the edge blocks of launch j then shared the SMs with interior launch
// j + 1 - one wave of blocks that each own a whole SM - and delayed it by their own duration every time;
// measured at 2 GPUs, 32768 x 4096 per GPU: 0.58 ms of a 11.9 ms step.)
std::vector<int> group_depths(int sweeps, int depth) {
    std::vector<int> d;
    if (sweeps % depth) d.push_back(sweeps % depth);
    for (int k = 0; k < sweeps / depth; ++k) d.push_back(depth);
    return d;
}

int ensure_edge_stream(natrix_sim* s) {
    if (s->st_edge) return 0;
    int lo = 0, hi = 0;
    CU(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    // highest priority: the exchange and the edge blocks are few and the rest of the group waits for them
    CU(cudaStreamCreateWithPriority(&s->st_edge, cudaStreamNonBlocking, hi));
    CU(cudaEventCreateWithFlags(&s->ev_group, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&s->ev_edges, cudaEventDisableTiming));
    return 0;
}

// phase 4: the interior of the group's first launch; the host queues the exchange on st_edge next
int phase_jacobi_interior(natrix_sim* s, int sweeps) {
    Range nvtx_range("natrix.jacobi.interior");
    const Geom& g = s->g;
    const bool up = g.y0 > 0, down = g.y0 + g.hl < g.hg;
    if (s->group_open) return fail(NATRIX_ERR_STATE, "a Jacobi group is open: call phase 5 first");
    if ((up || down) && sweeps > g.halo) return fail(NATRIX_ERR_ARG, "group is deeper than the slab's halo");
    if (g.hl < 2 * sweeps) return fail(NATRIX_ERR_STATE, "slab is too short to split into interior and edges");
    if (int rc = ensure_edge_stream(s)) return rc;
    // everything queued so far (divergence, the previous group) precedes the exchange and the edge zones
    CU(cudaEventRecord(s->ev_group, s->st));
    CU(cudaStreamWaitEvent(s->st_edge, s->ev_group, 0));
    const int d1 = group_depths(sweeps, jacobi_launch_depth(s))[0];
    if (int rc = jacobi_rows(s, s->pr, d1, up ? d1 : 0, down ? g.hl - d1 : g.hl, s->p_is_zero, s->st)) return rc;
    s->group_open = sweeps;
    CU(cudaGetLastError());
    return 0;
}

// phase 5: the exchange is queued on st_edge; the first launch's edge zones there, the rest of the group on main
int phase_jacobi_edges(natrix_sim* s, int sweeps) {
    Range nvtx_range("natrix.jacobi.edges");
    const Geom& g = s->g;
    const bool up = g.y0 > 0, down = g.y0 + g.hl < g.hg;
    if (s->group_open != sweeps) return fail(NATRIX_ERR_STATE, "phase 5 must follow phase 4 with the same sweeps");
    const std::vector<int> depths = group_depths(sweeps, jacobi_launch_depth(s));
    int src = s->pr, done = depths[0];
    if (up)
        if (int rc = jacobi_rows(s, src, depths[0], s->ext_lo(sweeps - done), done, s->p_is_zero, s->st_edge)) return rc;
    if (down)
        if (int rc = jacobi_rows(s, src, depths[0], g.hl - done, s->ext_hi(sweeps - done), s->p_is_zero, s->st_edge)) return rc;
    CU(cudaEventRecord(s->ev_edges, s->st_edge));
    CU(cudaStreamWaitEvent(s->st, s->ev_edges, 0));
    src = 1 - src;
    for (size_t j = 1; j < depths.size(); ++j) {
        done += depths[j];
        if (int rc = jacobi_rows(s, src, depths[j], s->ext_lo(sweeps - done), s->ext_hi(sweeps - done), false, s->st)) return rc;
        src = 1 - src;
    }
    s->pr = src;
    s->p_is_zero = false;
    s->group_open = 0;
    CU(cudaGetLastError());
    return 0;
}

// NATRIX_OPT_SOLVER 1 / 2 in place of the Jacobi sweeps (full grid): `iterations` red-black SOR sweeps or V-cycles on
// the pressure buffer in place, from zero unless the simulator warm-starts
int phase_solver(natrix_sim* s) {
    Range nvtx_range(s->solver == 1 ? "natrix.sor" : "natrix.multigrid");
    const Geom& g = s->g;
    if (g.hl != g.hg) return fail(NATRIX_ERR_STATE, "the SOR / multigrid solvers run on a full grid only");
    float* p = s->p[s->pr];
    if (s->p_is_zero) CU(cudaMemsetAsync(p, 0, (size_t)g.w * g.hl * sizeof(float), s->st));
    s->p_is_zero = false;
    // the tiled smoother ping-pongs between the two pressure buffers: whichever holds the result becomes PRESSURE_READ
    float* other = s->p[1 - s->pr];
    int n;
    if (s->solver == 1) {
        n = launch_sor_sweeps(&p, &other, s->div, s->nbm, g.w, g.hl, (float)(s->sor_omega_milli * 1e-3), s->iterations, s->st);
    } else {
        if (!s->mg) s->mg = multigrid_create(g.w, g.hl);
        if (!s->mg) return fail(NATRIX_ERR_CUDA, "multigrid: out of device memory");
        n = multigrid_solve(s->mg, &p, &other, s->div, s->obs, s->nbm, s->iterations, s->mg_smooth, s->st);
    }
    if (n < 0) return fail(NATRIX_ERR_CUDA, std::string("solver launch: ") + cudaGetErrorString(cudaGetLastError()));
    s->launches += n;
    if (p != s->p[s->pr]) s->pr = 1 - s->pr;
    CU(cudaGetLastError());
    return 0;
}

int phase_project(natrix_sim* s) {
    Range nvtx_range("natrix.subtract_gradient");
    const Geom& g = s->g;
    stamp(s, ST_GRAD);
    if (s->grad_done)
        s->grad_done = false;                    // the last Jacobi launch wrote the projected velocity already
    else if (s->pipeline == 0)
        s->launches += launch_gradient_ref(s->vel[s->vr], s->p[s->pr], s->obs, s->vel[1 - s->vr], g, 0, g.hl, s->st);
    else
        s->launches += launch_gradient_mask(s->vel[s->vr], s->p[s->pr], s->nbm, s->vel[1 - s->vr], g, 0, g.hl,
                                            s->d_err + 1, s->st);
    s->vr = 1 - s->vr;
    stamp(s, ST_CLEAR);
    // clear obstacles (fluid_simulator.py:268-280); skipped when nothing was stamped since the
    // last clear (the map is already all zero)
    if (s->pipeline == 0) {
        CU(cudaMemsetAsync(s->obs_base, 0, s->cells_alloc, s->st));
        s->launches += 1;
    } else if (s->obs_dirty) {
        // only the rows that were stamped since the last clear can be non-zero
        for (size_t k = 0; k + 1 < s->heavy.size(); k += 2) {
            const int lo = s->heavy[k], hi = s->heavy[k + 1];
            CU(cudaMemsetAsync(s->obs + (ptrdiff_t)lo * g.w, 0, (size_t)(hi - lo) * g.w, s->st));
            s->launches += 1;
        }
    }
    s->obs_dirty = false;
    s->heavy.clear();
    s->boxes.clear();
    stamp(s, ST_COUNT);
    CU(cudaGetLastError());
    return 0;
}

int field_info(natrix_sim* s, int field, void** base_row0, size_t* elem) {
    switch (field) {
    case NATRIX_VELOCITY: *base_row0 = s->vel[s->vr]; *elem = sizeof(float2); return 0;
    case NATRIX_PRESSURE: *base_row0 = s->p[s->pr]; *elem = sizeof(float); return 0;
    case NATRIX_DIVERGENCE: *base_row0 = s->div; *elem = sizeof(float); return 0;
    case NATRIX_VORTICITY: *base_row0 = s->vort; *elem = sizeof(float); return 0;
    case NATRIX_OBSTACLES: *base_row0 = s->obs; *elem = sizeof(uint8_t); return 0;
    case NATRIX_NBMASK: *base_row0 = s->nbm; *elem = sizeof(uint8_t); return 0;
    case NATRIX_DIV4: *base_row0 = s->div4; *elem = sizeof(float); return 0;
    default: return fail(NATRIX_ERR_ARG, "unknown field id");
    }
}

// ---- the library's own halo exchange (natrix_comm_init) ------------------------------------------------
// send = the slab's own first / last `rows` rows of a field, recv = the halo rows beyond them
void halo_ptrs(char* row0, size_t row_bytes, int hl, int side, int rows, char** send, char** recv) {
    if (side == 0) {
        *send = row0;
        *recv = row0 - (ptrdiff_t)rows * (ptrdiff_t)row_bytes;
    } else {
        *send = row0 + (size_t)(hl - rows) * row_bytes;
        *recv = row0 + (size_t)hl * row_bytes;
    }
}

// One NCCL group: `rows` halo rows of every listed field swapped with both neighbouring ranks, on stream st.
// Point-to-point only - the path has no collective.
int exchange_fields(natrix_sim* s, const int* fields, int nfields, int rows, cudaStream_t st) {
    Range nvtx_range("natrix.halo_exchange");
    if (rows <= 0 || !s->comm) return 0;
    const Geom& g = s->g;
    if (rows > g.halo || rows > g.hl)
        return fail(NATRIX_ERR_RANGE, "this step needs " + std::to_string(rows) + " halo rows but the slab was created with " +
                                      std::to_string(g.halo));
    Nccl* n = nccl();
    NC(n->GroupStart());
    for (int side = 0; side < 2; ++side) {
        const int peer = side == 0 ? s->comm_rank - 1 : s->comm_rank + 1;
        if (side == 0 ? g.y0 <= 0 : g.y0 + g.hl >= g.hg) continue;
        for (int k = 0; k < nfields; ++k) {
            void* row0 = nullptr;
            size_t elem = 0;
            if (int rc = field_info(s, fields[k], &row0, &elem)) { n->GroupEnd(); return rc; }
            char *send, *recv;
            halo_ptrs((char*)row0, (size_t)g.w * elem, g.hl, side, rows, &send, &recv);
            const size_t bytes = (size_t)rows * g.w * elem;
            NC(n->Send(send, bytes, NCCL_UINT8, peer, s->comm, st));
            NC(n->Recv(recv, bytes, NCCL_UINT8, peer, s->comm, st));
            s->exchanged_bytes += bytes;
        }
    }
    NC(n->GroupEnd());
    s->exchanges += 1;
    for (int k = 0; k < nfields; ++k)
        if (fields[k] == NATRIX_VELOCITY) s->vel_halo_valid = rows;
    return 0;
}

// the same for the rows of a dye field (its own geometry), on the simulator's stream
int exchange_dye(natrix_dye* d, int rows) {
    Range nvtx_range("natrix.halo_exchange.dye");
    natrix_sim* s = d->sim;
    if (rows <= 0 || !s->comm) return 0;
    const Geom& g = d->g;
    if (rows > g.halo || rows > g.hl)
        return fail(NATRIX_ERR_RANGE, "this dye step needs " + std::to_string(rows) + " halo rows but the dye slab was created with " +
                                      std::to_string(g.halo));
    Nccl* n = nccl();
    const size_t row_bytes = (size_t)g.w * sizeof(float), bytes = (size_t)rows * row_bytes;
    NC(n->GroupStart());
    for (int side = 0; side < 2; ++side) {
        if (side == 0 ? g.y0 <= 0 : g.y0 + g.hl >= g.hg) continue;
        const int peer = side == 0 ? s->comm_rank - 1 : s->comm_rank + 1;
        char *send, *recv;
        halo_ptrs((char*)d->d[d->rd], row_bytes, g.hl, side, rows, &send, &recv);
        NC(n->Send(send, bytes, NCCL_UINT8, peer, s->comm, s->st));
        NC(n->Recv(recv, bytes, NCCL_UINT8, peer, s->comm, s->st));
        s->exchanged_bytes += bytes;
    }
    NC(n->GroupEnd());
    s->exchanges += 1;
    d->halo_valid = rows;
    return 0;
}

int velocity_rows_for_step(const natrix_sim* s, float dt) {
    // |v| <= 1 after add_velocity / advect clamps; the projection can exceed it slightly, so the reach is padded
    // and the kernel reports (NATRIX_ERR_RANGE) if a back-trace still leaves the exchanged rows.
    const double reach = std::ceil(1.25 * (double)dt * (double)s->speed) + 1.0;
    return (int)reach + 4;
}

// The whole step of one slab, exchanges included (ref: FluidSimulator.update, fluid_simulator.py:174-280; the
// exchange schedule is DESIGN.md section 6, the executable model of it natrix_b200/slabs.py SlabSimulator.update).
int step_slab(natrix_sim* s, float dt) {
    const Geom& g = s->g;
    if (int rc = flush_splats(s)) return rc;
    const int vel = NATRIX_VELOCITY, prs = NATRIX_PRESSURE;
    // every allocated halo row of the velocity, not just the rows_needed a |v| <= 1 flow reaches: the projection
    // and the confinement force can push |v| past 1, and 48 rows cost microseconds over NVLink.  A back-trace
    // that still leaves them raises NATRIX_ERR_RANGE.
    if (int rc = exchange_fields(s, &vel, 1, std::max(std::min(velocity_rows_for_step(s, dt), g.halo), std::min(g.halo, g.hl)), s->st)) return rc;
    if (int rc = phase_advect(s, dt)) return rc;
    if (int rc = phase_forces(s, dt)) return rc;
    // Several Jacobi launches per exchange: with k * depth halo rows the slab recomputes the rows its neighbour
    // owns for the first k - 1 launches instead of exchanging after every one.  A partial group goes first so
    // that launch depths never decrease (phase_jacobi_interior / phase_jacobi_edges).
    const int n = s->iterations, depth = jacobi_launch_depth(s);
    const int span = std::max(1, g.halo / depth) * depth;
    std::vector<int> groups;
    if (n % span) groups.push_back(n % span);
    for (int k = 0; k < n / span; ++k) groups.push_back(span);
    const bool overlap = s->overlap && s->comm_world > 1 && g.hl >= 2 * span;
    stamp(s, ST_JACOBI);
    s->first_block = false;
    for (size_t i = 0; i < groups.size(); ++i) {
        const int t = groups[i];
        cudaStream_t xs = overlap ? s->st_edge : s->st;
        if (overlap && s->xfirst) {
            // the exchange kernel is queued BEFORE the interior launch and that launch leaves it `reserve` SMs:
            // an interior launch is one wave of blocks that each own a whole SM, so an exchange queued behind it
            // would only start when that wave drains
            CU(cudaEventRecord(s->ev_group, s->st));
            CU(cudaStreamWaitEvent(s->st_edge, s->ev_group, 0));
        } else if (overlap) {
            jacobi_tb_reserve_sms(s->tb, s->reserve_sms);
            const int rc = phase_jacobi_interior(s, t);
            jacobi_tb_reserve_sms(s->tb, 0);
            if (rc) return rc;
        }
        if (i == 0) {
            // p starts at zero, halos included - unless the simulator warm-starts from the last step's pressure
            // the fused pipeline's sweeps read the scaled divergence, the reference-order pipeline the divergence itself
            const int first[3] = {s->pipeline != 0 ? NATRIX_DIV4 : NATRIX_DIVERGENCE, NATRIX_NBMASK, NATRIX_PRESSURE};
            if (int rc = exchange_fields(s, first, s->warm_start ? 3 : 2, std::min(span, n), xs)) return rc;
        } else {
            if (int rc = exchange_fields(s, &prs, 1, t, xs)) return rc;
        }
        if (overlap && s->xfirst) {
            jacobi_tb_reserve_sms(s->tb, s->reserve_sms);
            const int rc = phase_jacobi_interior(s, t);
            jacobi_tb_reserve_sms(s->tb, 0);
            if (rc) return rc;
        }
        if (int rc = overlap ? phase_jacobi_edges(s, t) : phase_jacobi(s, t)) return rc;
    }
    if (int rc = exchange_fields(s, &prs, 1, 1, s->st)) return rc;
    if (int rc = phase_project(s)) return rc;
    return check_range_flag(s, false);
}

// Standard row partition (the first height % world ranks hold one extra row): (row0, rows) of `rank`.
void partition_rows(int height, int world, int rank, int* row0, int* rows) {
    const int base = height / world, extra = height % world;
    *rows = base + (rank < extra ? 1 : 0);
    *row0 = rank * base + std::min(rank, extra);
}

// Rows of post-projection velocity a dye slab samples beyond its simulator slab, maximum over ranks (every rank
// must exchange the same count); the shader's float32 (y / dye_height) * grid_height at the first and last dye
// row of each slab (ref: demo/shaders/shader.AdvectParticle.comp:46).
int dye_velocity_rows(int dye_height, int grid_height, int world) {
    int need = 0;
    for (int r = 0; r < world; ++r) {
        int p0, pn, v0, vn;
        partition_rows(dye_height, world, r, &p0, &pn);
        partition_rows(grid_height, world, r, &v0, &vn);
        const float lo = ((float)p0 / (float)dye_height) * (float)grid_height;
        const float hi = ((float)(p0 + pn - 1) / (float)dye_height) * (float)grid_height;
        const int first = std::max(0, (int)std::floor(lo)), last = std::min(grid_height - 1, (int)std::ceil(hi));
        need = std::max(need, std::max(v0 - first, last - (v0 + vn - 1)));
    }
    return need;
}

}  // namespace

extern "C" {

const char* natrix_last_error(void) { return g_err.c_str(); }
const char* natrix_version(void) { return "natrix_b200 0.1 (sm_100a)"; }

int natrix_create_slab(int width, int global_height, int row0, int rows, int halo, int device,
                       natrix_sim** out) {
    NEED(out, "out is null");
    *out = nullptr;
    NEED(width > 0 && global_height > 0, "width and height must be positive");
    NEED(rows > 0 && row0 >= 0 && row0 + rows <= global_height, "slab rows outside the global grid");
    NEED(halo >= 0, "halo must be >= 0");
    NEED(rows == global_height ? true : halo >= 1, "a partial slab needs halo >= 1");
    int ndev = 0;
    CU(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(NATRIX_ERR_CUDA, "no such CUDA device");
    CU(cudaSetDevice(device));
    natrix_sim* s = new natrix_sim();
    s->device = device;
    s->g = Geom{width, global_height, row0, rows, halo};
    s->rows_alloc = (size_t)rows + 2 * (size_t)halo;
    s->cells_alloc = s->rows_alloc * (size_t)width;
    cudaError_t e = cudaStreamCreateWithFlags(&s->st, cudaStreamNonBlocking);
    for (int i = 0; i < 2 && e == cudaSuccess; ++i) {
        e = alloc_rows(&s->vel_base[i], &s->vel[i], s);
        if (e == cudaSuccess) e = alloc_rows(&s->p_base[i], &s->p[i], s);
    }
    if (e == cudaSuccess) e = alloc_rows(&s->div_base, &s->div, s);
    if (e == cudaSuccess) e = alloc_rows(&s->vort_base, &s->vort, s);
    if (e == cudaSuccess) e = alloc_rows(&s->div4_base, &s->div4, s);
    if (e == cudaSuccess) e = alloc_rows(&s->obs_base, &s->obs, s);
    if (e == cudaSuccess) e = alloc_rows(&s->nbm_base, &s->nbm, s);
    s->nbands = (int)((s->rows_alloc + OVER_BAND - 1) / OVER_BAND);
    if (e == cudaSuccess) e = cudaMalloc((void**)&s->d_err, (1 + s->nbands) * sizeof(int));
    if (e == cudaSuccess) e = cudaMemsetAsync(s->d_err, 0, (1 + s->nbands) * sizeof(int), s->st);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&s->sm_count, cudaDevAttrMultiProcessorCount, device);
    if (e == cudaSuccess) e = cudaMallocHost((void**)&s->h_err, sizeof(int));
    if (e == cudaSuccess) { *s->h_err = 0; e = cudaEventCreateWithFlags(&s->err_event, cudaEventDisableTiming); }
    if (e == cudaSuccess) e = cudaMalloc((void**)&s->d_scratch, 4 * 1024 * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc((void**)&s->d_out4, 4 * sizeof(double));
    if (e == cudaSuccess) e = cudaMallocHost((void**)&s->h_out4, 4 * sizeof(double));
    for (int i = 0; i <= ST_COUNT && e == cudaSuccess; ++i) e = cudaEventCreate(&s->ev[i]);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s->st);
    if (e != cudaSuccess) {
        std::string msg = std::string("natrix_create: ") + cudaGetErrorString(e);
        natrix_destroy(s);
        return fail(NATRIX_ERR_CUDA, msg);
    }
    s->tb = jacobi_tb_create();
    s->smem_depth = jacobi_smem_max_depth();
    if (const char* e = getenv("NATRIX_SMEM_GRAD")) s->smem_grad = atoi(e) != 0;
    *out = s;
    return 0;
}

int natrix_create(int width, int height, int device, natrix_sim** out) {
    return natrix_create_slab(width, height, 0, height, 0, device, out);
}

int natrix_destroy(natrix_sim* s) {
    if (!s) return 0;
    cudaSetDevice(s->device);
    if (s->st) cudaStreamSynchronize(s->st);
    for (natrix_dye* d : s->dyes) d->sim = nullptr;   // orphaned dye handles stay destroyable
    if (s->comm) { nccl()->CommDestroy(s->comm); s->comm = nullptr; }
    jacobi_tb_destroy(s->tb);
    multigrid_destroy(s->mg);
    for (int i = 0; i < 2; ++i) { cudaFree(s->vel_base[i]); cudaFree(s->p_base[i]); }
    cudaFree(s->div_base); cudaFree(s->vort_base); cudaFree(s->div4_base); cudaFree(s->obs_base); cudaFree(s->nbm_base);
    cudaFree(s->d_err); cudaFree(s->d_scratch); cudaFree(s->d_out4); cudaFree(s->d_tmp2);
    if (s->h_err) cudaFreeHost(s->h_err);
    if (s->err_event) cudaEventDestroy(s->err_event);
    if (s->h_out4) cudaFreeHost(s->h_out4);
    for (int i = 0; i <= ST_COUNT; ++i) if (s->ev[i]) cudaEventDestroy(s->ev[i]);
    if (s->st_edge) { cudaStreamSynchronize(s->st_edge); cudaStreamDestroy(s->st_edge); }
    if (s->ev_group) cudaEventDestroy(s->ev_group);
    if (s->ev_edges) cudaEventDestroy(s->ev_edges);
    if (s->st) cudaStreamDestroy(s->st);
    delete s;
    return 0;
}

int natrix_set_params(natrix_sim* s, float speed, int iterations, float dissipation, float vorticity,
                      double viscosity, int has_borders) {
    NEED(s, "null simulator");
    // same validation as the property setters (fluid_simulator.py:58-111)
    NEED(speed > 0, "'Speed' should be greater than zero");
    NEED(iterations > 0, "'Iterations' should be grater than zero");
    NEED(dissipation > 0, "'Dissipation' should be grater than zero");
    NEED(vorticity >= 0, "'Vorticity' should be grater or equal than zero");
    NEED(viscosity >= 0.0, "'Viscosity' should be greater or equal than zero");
    s->speed = speed; s->iterations = iterations; s->dissipation = dissipation;
    s->vorticity = vorticity; s->has_borders = has_borders ? 1 : 0;
    s->viscous = viscosity > 0.0;
    if (s->viscous) {           // fluid_simulator.py:327-336, in double then narrowed
        const double centre = 1.0 / viscosity;
        s->alpha = (float)centre;
        s->rbeta = (float)(1.0 / (4.0 + centre));
    }
    return 0;
}

int natrix_set_option(natrix_sim* s, int option, int value) {
    NEED(s, "null simulator");
    switch (option) {
    case NATRIX_OPT_PIPELINE:
        NEED(value == 0 || value == 1, "pipeline must be 0 or 1");
        if (int rc = select_device(s)) return rc;
        if (int rc = flush_splats(s)) return rc;
        if (int rc = flush_circles(s)) return rc;
        CU(cudaMemsetAsync(s->d_err + 1, 1, s->nbands * sizeof(int), s->st));   // pipeline 0 does not track |v| > 1
        s->pipeline = value; return 0;
    case NATRIX_OPT_JACOBI_DEPTH:
        NEED(value >= 1 && value <= JACOBI_TB_MAX_DEPTH, "jacobi depth out of range");
        s->jacobi_depth = value; return 0;
    case NATRIX_OPT_TIMING: s->timing = value ? 1 : 0; return 0;
    case NATRIX_OPT_PACKED: s->packed = value ? 1 : 0; return 0;
    case NATRIX_OPT_WARM_START: s->warm_start = value ? 1 : 0; return 0;
    case NATRIX_OPT_SOLVER:
        NEED(value >= 0 && value <= 2, "solver must be 0 (Jacobi, the reference), 1 (red-black SOR) or 2 (multigrid)");
        NEED(value == 0 || s->g.hl == s->g.hg, "the SOR / multigrid solvers run on a full grid only");
        s->solver = value; return 0;
    case NATRIX_OPT_SOR_OMEGA_MILLI:
        NEED(value > 0 && value < 2000, "SOR omega (x 1000) must lie in (0, 2000)");
        s->sor_omega_milli = value; return 0;
    case NATRIX_OPT_MG_SMOOTH:
        NEED(value >= 1 && value <= 8, "multigrid smoothing sweeps must lie in 1..8");
        s->mg_smooth = value; return 0;
    case NATRIX_OPT_JACOBI_KERNEL:
        NEED(value >= 0 && value <= 2, "jacobi kernel must be 0 (auto), 1 (TMA register streaming) or 2 (shared memory)");
        NEED(value != 1 || jacobi_tb_supported(s->g), "the TMA kernel needs width % 16 == 0 and width >= 128");
        s->jacobi_kernel = value; return 0;
    case NATRIX_OPT_SMEM_DEPTH:
        NEED(value >= 1 && value <= 16, "shared-memory jacobi depth out of range (1..16)");
        s->smem_depth = value; return 0;
    default: return fail(NATRIX_ERR_ARG, "unknown option id");
    }
}

int natrix_get_option(natrix_sim* s, int option, int* value) {
    NEED(s && value, "null argument");
    switch (option) {
    case NATRIX_OPT_PIPELINE: *value = s->pipeline; return 0;
    case NATRIX_OPT_JACOBI_DEPTH: *value = s->jacobi_depth; return 0;
    case NATRIX_OPT_TIMING: *value = s->timing; return 0;
    case NATRIX_OPT_PACKED: *value = s->packed; return 0;
    case NATRIX_OPT_WARM_START: *value = s->warm_start; return 0;
    case NATRIX_OPT_JACOBI_KERNEL: *value = s->pipeline == 0 ? 0 : (use_smem_kernel(s) ? 2 : 1); return 0;   // the one in use
    case NATRIX_OPT_SMEM_DEPTH: *value = s->smem_depth; return 0;
    case NATRIX_OPT_SOLVER: *value = s->solver; return 0;
    case NATRIX_OPT_SOR_OMEGA_MILLI: *value = s->sor_omega_milli; return 0;
    case NATRIX_OPT_MG_SMOOTH: *value = s->mg_smooth; return 0;
    default: return fail(NATRIX_ERR_ARG, "unknown option id");
    }
}

int natrix_add_velocity(natrix_sim* s, float px, float py, float vx, float vy, float radius) {
    NEED(s, "null simulator");
    // splat_pos = _Position * _Size (shader.AddVelocity.comp:27), float32 products
    SplatV sp{px * (float)s->g.w, py * (float)s->g.hg, vx, vy, radius};
    s->pending.push_back(sp);
    if (s->pipeline == 0) {
        if (int rc = select_device(s)) return rc;
        return flush_splats(s);          // one dispatch per call, like the reference
    }
    return 0;
}

int natrix_add_circle_obstacle(natrix_sim* s, float px, float py, float radius, int is_static) {
    NEED(s, "null simulator");
    (void)is_static;                      // both shader branches write (1,0): SURVEY Q17
    if (int rc = select_device(s)) return rc;
    const Geom& g = s->g;
    if (s->pipeline != 0) {
        // splat_pos = _Position * _Size (shader.AddCircleObstacle.comp:25), float32 products
        s->circles.insert(s->circles.end(), {px * (float)g.w, py * (float)g.hg, radius});
    } else {
        s->launches += launch_add_circle(s->obs, g, s->ext_lo(g.halo), s->ext_hi(g.halo), px * (float)g.w,
                                         py * (float)g.hg, radius, false, s->st);
    }
    if (radius >= 0.0f)
        mark_heavy_rows(s, (double)py * g.hg - radius, (double)py * g.hg + radius, (double)px * g.w - radius,
                        (double)px * g.w + radius, true);
    s->obs_dirty = true;
    CU(cudaGetLastError());
    return 0;
}

int natrix_add_triangle_obstacle(natrix_sim* s, float p1x, float p1y, float p2x, float p2y, float p3x,
                                 float p3y, int is_static) {
    NEED(s, "null simulator");
    if (int rc = select_device(s)) return rc;
    if (int rc = flush_circles(s)) return rc;        // keep the stamping order (a static triangle stores another code)
    const Geom& g = s->g;
    s->launches += launch_add_triangle(s->obs, g, s->ext_lo(g.halo), s->ext_hi(g.halo), p1x, p1y, p2x, p2y,
                                       p3x, p3y, is_static, s->st);
    {
        const double ys[3] = {(double)p1y * g.hg, (double)p2y * g.hg, (double)p3y * g.hg};
        const double ymin = std::min(ys[0], std::min(ys[1], ys[2])), ymax = std::max(ys[0], std::max(ys[1], ys[2]));
        // a degenerate triangle selects whole lines of cells (see stages_ref.cu): call every row heavy
        const double xs[3] = {(double)p1x * g.w, (double)p2x * g.w, (double)p3x * g.w};
        const double xmin = std::min(xs[0], std::min(xs[1], xs[2])), xmax = std::max(xs[0], std::max(xs[1], xs[2]));
        // ... and so does any (nearly) collinear triple, whatever its bounding box: all three edge functions vanish
        // along the whole line through the points.  Non-finite vertices select unpredictably: every row as well.
        const double cross = std::fabs((xs[1] - xs[0]) * (ys[2] - ys[0]) - (xs[2] - xs[0]) * (ys[1] - ys[0]));
        const double extent = std::max(xmax - xmin, ymax - ymin);
        const bool finite = std::isfinite(xmin) && std::isfinite(xmax) && std::isfinite(ymin) && std::isfinite(ymax);
        if (!finite || ymax - ymin < 1.0 || xmax - xmin < 1.0 || cross < 1e-3 * extent * extent + 1.0)
            mark_heavy_rows(s, 0.0, (double)g.hg);
        else mark_heavy_rows(s, ymin, ymax, xmin, xmax);
    }
    s->obs_dirty = true;
    CU(cudaGetLastError());
    return 0;
}

int natrix_step_phase(natrix_sim* s, int phase, float dt, int sweeps) {
    NEED(s, "null simulator");
    if (int rc = select_device(s)) return rc;
    switch (phase) {
    case 0: return phase_advect(s, dt);
    case 1: return phase_forces(s, dt);
    case 2:
        NEED(sweeps > 0, "sweeps must be positive");
        NEED(s->solver == 0, "the SOR / multigrid solvers are driven by natrix_step on a full grid, not phase by phase");
        if (s->first_block) stamp(s, ST_JACOBI);     // first block of the step: the stage spans all blocks + exchanges
        s->first_block = false;
        return phase_jacobi(s, sweeps);
    case 3: {
        if (int rc = phase_project(s)) return rc;
        return check_range_flag(s, false);
    }
    case 4:
        NEED(sweeps > 0, "sweeps must be positive");
        if (s->first_block) stamp(s, ST_JACOBI);
        s->first_block = false;
        return phase_jacobi_interior(s, sweeps);
    case 5: return phase_jacobi_edges(s, sweeps);
    default: return fail(NATRIX_ERR_ARG, "phase must be 0..5");
    }
}

int natrix_halo_rows_needed(natrix_sim* s, int phase, float dt) {
    if (!s) return fail(NATRIX_ERR_ARG, "null simulator");
    switch (phase) {
    case 0: {
        // |v| <= 1 after add_velocity / advect clamps; the projection can exceed it slightly, so
        // the reach is padded and the kernel reports (NATRIX_ERR_RANGE) if it was not enough.
        const double reach = std::ceil(1.25 * (double)dt * (double)s->speed) + 1.0;
        return (int)reach + 4;
    }
    case 1: return 0;                 // phase 0 already produced rows ext(4)
    case 2: return jacobi_launch_depth(s);   // per launch; a group of k launches between two exchanges needs k times that
    case 3: return 1;
    default: return fail(NATRIX_ERR_ARG, "phase must be 0..3");
    }
}

int natrix_halo_region(natrix_sim* s, int field, int side, int rows, void** send_ptr, void** recv_ptr,
                       size_t* bytes) {
    NEED(s && send_ptr && recv_ptr && bytes, "null argument");
    NEED(side == 0 || side == 1, "side must be 0 or 1");
    NEED(rows >= 0 && rows <= s->g.halo && rows <= s->g.hl, "rows exceed the slab's halo");
    if (int rc = select_device(s)) return rc;
    if (field == NATRIX_VELOCITY)
        if (int rc = flush_splats(s)) return rc;
    void* row0 = nullptr;
    size_t elem = 0;
    if (int rc = field_info(s, field, &row0, &elem)) return rc;
    const size_t row_bytes = (size_t)s->g.w * elem;
    halo_ptrs((char*)row0, row_bytes, s->g.hl, side, rows, (char**)send_ptr, (char**)recv_ptr);
    *bytes = (size_t)rows * row_bytes;
    if (field == NATRIX_VELOCITY) s->vel_halo_valid = rows;      // the host is about to fill them
    return 0;
}

int natrix_comm_unique_id(void* id128) {
    NEED(id128, "null argument");
    Nccl* n = nccl();
    if (!n->h) return fail(NATRIX_ERR_STATE, n->why);
    NC(n->GetUniqueId((NcclId*)id128));
    return 0;
}

int natrix_comm_init(natrix_sim* s, const void* id128, int rank, int world) {
    NEED(s && id128, "null argument");
    NEED(world >= 1 && rank >= 0 && rank < world, "rank outside the world");
    NEED(!s->comm, "the simulator already has a communicator");
    int row0 = 0, rows = 0;
    partition_rows(s->g.hg, world, rank, &row0, &rows);
    NEED(row0 == s->g.y0 && rows == s->g.hl,
         "slab rows must follow the standard partition: rank r holds rows [r*(H/N) + min(r, H%N), ...), H%N ranks one extra");
    Nccl* n = nccl();
    if (!n->h) return fail(NATRIX_ERR_STATE, n->why);
    if (int rc = select_device(s)) return rc;
    if (int rc = ensure_edge_stream(s)) return rc;
    NcclId id;
    memcpy(&id, id128, sizeof(id));
    NC(n->CommInitRank(&s->comm, world, id, rank));
    s->comm_rank = rank;
    s->comm_world = world;
    if (const char* e = getenv("NATRIX_SLAB_OVERLAP")) s->overlap = atoi(e) != 0;
    if (const char* e = getenv("NATRIX_SLAB_XFIRST")) s->xfirst = atoi(e) != 0;
    if (const char* e = getenv("NATRIX_SLAB_RESERVE")) s->reserve_sms = atoi(e);
    return 0;
}

int natrix_comm_stats(natrix_sim* s, unsigned long long* exchanges, unsigned long long* bytes) {
    NEED(s && exchanges && bytes, "null argument");
    *exchanges = s->exchanges;
    *bytes = s->exchanged_bytes;
    return 0;
}

int natrix_step(natrix_sim* s, float dt) {
    Range nvtx_range("natrix.step");
    NEED(s, "null simulator");
    if (int rc = select_device(s)) return rc;
    if (s->g.hl != s->g.hg) {
        if (!s->comm)
            return fail(NATRIX_ERR_STATE, "natrix_step on a slab needs natrix_comm_init (or drive natrix_step_phase with your own exchange)");
        return step_slab(s, dt);
    }
    if (int rc = phase_advect(s, dt)) return rc;
    if (int rc = phase_forces(s, dt)) return rc;
    stamp(s, ST_JACOBI);
    s->grad_wanted = s->smem_grad != 0;          // this call runs ALL the step's sweeps: the last launch may project too
    const int rc_solve = s->solver != 0 && s->pipeline != 0 ? phase_solver(s) : phase_jacobi(s, s->iterations);
    s->grad_wanted = false;
    if (rc_solve) return rc_solve;
    if (int rc = phase_project(s)) return rc;
    return 0;
}

int natrix_field_ptr(natrix_sim* s, int field, void** dev_ptr, size_t* bytes) {
    NEED(s && dev_ptr, "null argument");
    if (int rc = select_device(s)) return rc;
    if (field == NATRIX_VELOCITY)
        if (int rc = flush_splats(s)) return rc;
    if (field == NATRIX_OBSTACLES)
        if (int rc = flush_circles(s)) return rc;
    size_t elem = 0;
    if (int rc = field_info(s, field, dev_ptr, &elem)) return rc;
    if (bytes) *bytes = (size_t)s->g.w * s->g.hl * elem;
    return 0;
}

int natrix_copy_out(natrix_sim* s, int field, void* host, size_t bytes) {
    NEED(s && host, "null argument");
    if (int rc = select_device(s)) return rc;
    if (field == NATRIX_VELOCITY)
        if (int rc = flush_splats(s)) return rc;
    void* src = nullptr;
    size_t elem = 0;
    if (int rc = field_info(s, field, &src, &elem)) return rc;
    const size_t n = (size_t)s->g.w * s->g.hl;
    if (field == NATRIX_OBSTACLES) {
        if (int rc = flush_circles(s)) return rc;
        NEED(bytes == n * sizeof(float2), "OBSTACLES copy_out expects width*height*8 bytes");
        if (!s->d_tmp2) CU(cudaMalloc((void**)&s->d_tmp2, n * sizeof(float2)));
        s->launches += launch_obs_expand(s->obs, s->d_tmp2, n, s->st);
        src = s->d_tmp2;
        elem = sizeof(float2);
    }
    NEED(bytes == n * elem, "copy_out size does not match the field");
    CU(cudaMemcpyAsync(host, src, bytes, cudaMemcpyDeviceToHost, s->st));
    CU(cudaStreamSynchronize(s->st));
    return 0;
}

int natrix_copy_in(natrix_sim* s, int field, const void* host, size_t bytes) {
    NEED(s && host, "null argument");
    if (int rc = select_device(s)) return rc;
    if (field == NATRIX_VELOCITY)
        if (int rc = flush_splats(s)) return rc;
    void* dst = nullptr;
    size_t elem = 0;
    if (int rc = field_info(s, field, &dst, &elem)) return rc;
    const size_t n = (size_t)s->g.w * s->g.hl;
    if (field == NATRIX_OBSTACLES) {
        if (int rc = flush_circles(s)) return rc;
        NEED(bytes == n * sizeof(float2), "OBSTACLES copy_in expects width*height*8 bytes");
        if (!s->d_tmp2) CU(cudaMalloc((void**)&s->d_tmp2, n * sizeof(float2)));
        CU(cudaMemcpyAsync(s->d_tmp2, host, bytes, cudaMemcpyHostToDevice, s->st));
        s->launches += launch_obs_pack(s->d_tmp2, s->obs, n, s->st);
        mark_heavy_rows(s, 0.0, (double)s->g.hg);
        s->obs_dirty = true;
        CU(cudaStreamSynchronize(s->st));
        return 0;
    }
    NEED(bytes == n * elem, "copy_in size does not match the field");
    CU(cudaMemcpyAsync(dst, host, bytes, cudaMemcpyHostToDevice, s->st));
    if (field == NATRIX_DIVERGENCE || field == NATRIX_NBMASK)    // keep the scaled copy and its NB_RAW bits in step
        s->launches += launch_rescale_divergence(s->div, s->div4, s->nbm, n, s->st);
    CU(cudaStreamSynchronize(s->st));
    if (field == NATRIX_PRESSURE) s->p_is_zero = false;
    if (field == NATRIX_VELOCITY) CU(cudaMemsetAsync(s->d_err + 1, 1, s->nbands * sizeof(int), s->st));   // unknown range
    return 0;
}

int natrix_field_stats(natrix_sim* s, int field, double* out4) {
    Range nvtx_range("natrix.field_stats");
    NEED(s && out4, "null argument");
    NEED((field >= NATRIX_VELOCITY && field <= NATRIX_VORTICITY) || field == NATRIX_DIV4, "stats are defined for float fields");
    if (int rc = select_device(s)) return rc;
    if (field == NATRIX_VELOCITY)
        if (int rc = flush_splats(s)) return rc;
    void* src = nullptr;
    size_t elem = 0;
    if (int rc = field_info(s, field, &src, &elem)) return rc;
    const size_t nfl = (size_t)s->g.w * s->g.hl * (elem / sizeof(float));
    s->launches += launch_stats((const float*)src, nfl, s->d_scratch, s->d_out4, s->st);
    CU(cudaMemcpyAsync(s->h_out4, s->d_out4, 4 * sizeof(double), cudaMemcpyDeviceToHost, s->st));
    CU(cudaStreamSynchronize(s->st));
    memcpy(out4, s->h_out4, 4 * sizeof(double));
    return 0;
}

// ---- dye ---------------------------------------------------------------------------------------
int natrix_dye_create_slab(natrix_sim* s, int width, int global_height, int row0, int rows, int halo,
                           natrix_dye** out) {
    NEED(s && out, "null argument");
    *out = nullptr;
    NEED(width > 0 && global_height > 0, "width and height must be positive");
    NEED(row0 >= 0 && rows > 0 && row0 + rows <= global_height && halo >= 0, "bad dye slab geometry");
    const bool full = rows == global_height;
    NEED(full == (s->g.hl == s->g.hg), "a dye slab needs a simulator slab (and a full dye grid a full simulator)");
    if (int rc = select_device(s)) return rc;
    natrix_dye* d = new natrix_dye();
    d->sim = s;
    d->g = Geom{width, global_height, row0, rows, halo};
    const size_t cells = (size_t)width * ((size_t)rows + 2 * (size_t)halo);
    cudaError_t e = cudaSuccess;
    for (int i = 0; i < 2 && e == cudaSuccess; ++i) {
        e = cudaMalloc((void**)&d->base[i], cells * sizeof(float));
        if (e == cudaSuccess) e = cudaMemsetAsync(d->base[i], 0, cells * sizeof(float), s->st);
        d->d[i] = d->base[i] + (size_t)halo * width;
    }
    if (e == cudaSuccess) e = cudaMalloc((void**)&d->tables, (size_t)(width + rows) * sizeof(float));
    if (e != cudaSuccess) {
        cudaFree(d->base[0]); cudaFree(d->base[1]); cudaFree(d->tables); delete d;
        return fail(NATRIX_ERR_CUDA, std::string("natrix_dye_create: ") + cudaGetErrorString(e));
    }
    s->launches += launch_dye_tables(d->tables, d->tables + width, d->g, s->g.w, s->g.hg, s->st);
    s->dyes.push_back(d);
    *out = d;
    return 0;
}

int natrix_dye_create(natrix_sim* s, int width, int height, natrix_dye** out) {
    NEED(s, "null argument");
    NEED(s->g.hl == s->g.hg, "dye fields on a simulator slab are created with natrix_dye_create_slab");
    return natrix_dye_create_slab(s, width, height, 0, height, 0, out);
}

int natrix_dye_destroy(natrix_dye* d) {
    if (!d) return 0;
    if (d->sim) {
        cudaSetDevice(d->sim->device);
        cudaStreamSynchronize(d->sim->st);
        auto& v = d->sim->dyes;
        for (size_t i = 0; i < v.size(); ++i) if (v[i] == d) { v.erase(v.begin() + i); break; }
    }
    cudaFree(d->base[0]); cudaFree(d->base[1]); cudaFree(d->tables); cudaFree(d->rgba); cudaFree(d->lut);
    delete d;
    return 0;
}

#define DYE_LIVE(d) NEED((d) && (d)->sim, "dye handle is null or its simulator was destroyed")

int natrix_dye_add(natrix_dye* d, float px, float py, float radius, float strength) {
    DYE_LIVE(d);
    SplatD sp{px * (float)d->g.w, py * (float)d->g.hg, radius, strength};
    d->pending.push_back(sp);
    if (d->sim->pipeline == 0) {
        if (int rc = select_device(d->sim)) return rc;
        return flush_dye(d);
    }
    return 0;
}

int natrix_dye_step(natrix_dye* d, float dt, float speed, float dissipation) {
    Range nvtx_range("natrix.dye_step");
    DYE_LIVE(d);
    natrix_sim* s = d->sim;
    if (int rc = select_device(s)) return rc;
    if (int rc = flush_splats(s)) return rc;     // the advect reads the CURRENT velocity
    if (int rc = flush_circles(s)) return rc;    // ... and the CURRENT obstacle map
    if (int rc = flush_dye(d)) return rc;
    if (s->comm) {
        // the library's own exchange: the post-projection velocity rows the dye samples beyond the simulator slab
        // and the dye rows within back-trace reach (both neighbours apply the same add_particles calls first)
        const int vel = NATRIX_VELOCITY;
        if (int rc = exchange_fields(s, &vel, 1, dye_velocity_rows(d->g.hg, s->g.hg, s->comm_world), s->st)) return rc;
        const int need = (int)std::ceil(1.25 * (double)dt * (double)speed * ((double)d->g.hg / (double)s->g.hg)) + 2;
        if (int rc = exchange_dye(d, std::min(need, std::min(d->g.halo, d->g.hl)))) return rc;
    }
    // gathers may only touch halo rows filled for the CURRENT buffers (Geom::halo is the kernels' range limit)
    Geom dg = d->g, vg = s->g;
    dg.halo = std::min(dg.halo, d->halo_valid);
    vg.halo = std::min(vg.halo, s->vel_halo_valid);
    d->halo_valid = 0;
    if (s->pipeline != 0 && d->g.w % 4 == 0)
        s->launches += launch_dye_advect4(d->d[d->rd], d->d[1 - d->rd], dg, s->vel[s->vr], s->obs, vg, d->tables,
                                          d->tables + d->g.w, dt, speed, dissipation, s->d_err, s->st);
    else
        s->launches += launch_dye_advect(d->d[d->rd], d->d[1 - d->rd], dg, s->vel[s->vr], s->obs, vg, dt, speed,
                                         dissipation, s->d_err, s->st);
    d->rd = 1 - d->rd;
    CU(cudaGetLastError());
    if (s->g.hl != s->g.hg) return check_range_flag(s, false);
    return 0;
}

// Halo rows a dye slab needs before natrix_dye_step: which = 0 rows of the simulator's VELOCITY (post
// projection) around the rows the dye samples, 1 rows of the dye itself (back-trace reach + bilinear).
// The velocity figure depends on how the two grids are cut: the host takes the maximum over ranks.
int natrix_dye_halo_rows_needed(natrix_dye* d, int which, float dt, float speed) {
    if (!d || !d->sim) return fail(NATRIX_ERR_ARG, "dye handle is null or its simulator was destroyed");
    const Geom& g = d->g;
    const Geom& vg = d->sim->g;
    if (which == 0) {
        // the shader's float32 expression for the first and last own dye row (monotone in y)
        const float lo = ((float)g.y0 / (float)g.hg) * (float)vg.hg;
        const float hi = ((float)(g.y0 + g.hl - 1) / (float)g.hg) * (float)vg.hg;
        const int first = std::max(0, (int)std::floor(lo)), last = std::min(vg.hg - 1, (int)std::ceil(hi));
        return std::max(0, std::max(vg.y0 - first, last - (vg.y0 + vg.hl - 1)));
    }
    if (which == 1) {
        const double ry = (double)g.hg / (double)vg.hg;
        return (int)std::ceil(1.25 * (double)dt * (double)speed * ry) + 2;
    }
    return fail(NATRIX_ERR_ARG, "which must be 0 (velocity) or 1 (dye)");
}

// As natrix_halo_region, for the dye rows.  Pending add_particles calls are applied first so that both
// neighbours exchange the same state.
int natrix_dye_halo_region(natrix_dye* d, int side, int rows, void** send_ptr, void** recv_ptr, size_t* bytes) {
    DYE_LIVE(d);
    NEED(send_ptr && recv_ptr && bytes, "null argument");
    NEED(side == 0 || side == 1, "side must be 0 or 1");
    NEED(rows >= 0 && rows <= d->g.halo && rows <= d->g.hl, "rows exceed the dye slab's halo");
    if (int rc = select_device(d->sim)) return rc;
    if (int rc = flush_dye(d)) return rc;
    const size_t row_bytes = (size_t)d->g.w * sizeof(float);
    halo_ptrs((char*)d->d[d->rd], row_bytes, d->g.hl, side, rows, (char**)send_ptr, (char**)recv_ptr);
    *bytes = (size_t)rows * row_bytes;
    d->halo_valid = rows;                       // the host is about to fill them
    return 0;
}

int natrix_dye_field_ptr(natrix_dye* d, void** dev_ptr, size_t* bytes) {
    DYE_LIVE(d);
    NEED(dev_ptr, "null argument");
    if (int rc = select_device(d->sim)) return rc;
    if (int rc = flush_dye(d)) return rc;
    *dev_ptr = d->d[d->rd];
    if (bytes) *bytes = d->own_cells() * sizeof(float);
    return 0;
}

int natrix_dye_copy_out(natrix_dye* d, void* host, size_t bytes) {
    DYE_LIVE(d);
    NEED(host && bytes == d->own_cells() * sizeof(float), "dye copy_out size mismatch");
    if (int rc = select_device(d->sim)) return rc;
    if (int rc = flush_dye(d)) return rc;
    CU(cudaMemcpyAsync(host, d->d[d->rd], bytes, cudaMemcpyDeviceToHost, d->sim->st));
    CU(cudaStreamSynchronize(d->sim->st));
    return 0;
}

int natrix_dye_copy_in(natrix_dye* d, const void* host, size_t bytes) {
    DYE_LIVE(d);
    NEED(host && bytes == d->own_cells() * sizeof(float), "dye copy_in size mismatch");
    if (int rc = select_device(d->sim)) return rc;
    if (int rc = flush_dye(d)) return rc;
    CU(cudaMemcpyAsync(d->d[d->rd], host, bytes, cudaMemcpyHostToDevice, d->sim->st));
    CU(cudaStreamSynchronize(d->sim->st));
    return 0;
}

int natrix_dye_stats(natrix_dye* d, double* out4) {
    DYE_LIVE(d);
    NEED(out4, "null argument");
    natrix_sim* s = d->sim;
    if (int rc = select_device(s)) return rc;
    if (int rc = flush_dye(d)) return rc;
    s->launches += launch_stats(d->d[d->rd], d->own_cells(), s->d_scratch, s->d_out4, s->st);
    CU(cudaMemcpyAsync(s->h_out4, s->d_out4, 4 * sizeof(double), cudaMemcpyDeviceToHost, s->st));
    CU(cudaStreamSynchronize(s->st));
    memcpy(out4, s->h_out4, 4 * sizeof(double));
    return 0;
}

int natrix_dye_export_rgba8(natrix_dye* d, void* out, size_t bytes, int is_device) {
    DYE_LIVE(d);
    const size_t n = d->own_cells();
    NEED(out && bytes == n * 4, "rgba8 export expects width*rows*4 bytes");
    natrix_sim* s = d->sim;
    if (int rc = select_device(s)) return rc;
    if (int rc = flush_dye(d)) return rc;
    if (is_device) {
        s->launches += launch_dye_rgba8(d->d[d->rd], (uint32_t*)out, n, s->st);
        CU(cudaGetLastError());
        return 0;
    }
    if (!d->rgba) CU(cudaMalloc((void**)&d->rgba, n * 4));
    s->launches += launch_dye_rgba8(d->d[d->rd], d->rgba, n, s->st);
    CU(cudaMemcpyAsync(out, d->rgba, bytes, cudaMemcpyDeviceToHost, s->st));
    CU(cudaStreamSynchronize(s->st));
    return 0;
}

int natrix_render_frame(natrix_dye* d, void* out, size_t bytes, int is_device, float quiver_tile) {
    DYE_LIVE(d);
    natrix_sim* s = d->sim;
    NEED(s->g.hl == s->g.hg, "frames are rendered from the full grid (gather the slabs first)");
    const size_t n = d->own_cells();
    NEED(out && bytes == n * 4, "a frame is width*height*4 bytes");
    if (int rc = select_device(s)) return rc;
    if (int rc = flush_dye(d)) return rc;
    if (quiver_tile > 0.0f)
        if (int rc = flush_splats(s)) return rc;             // the overlay shows the CURRENT velocity
    if (!d->lut) {
        CU(cudaMalloc((void**)&d->lut, 256 * sizeof(float4)));
        s->launches += launch_field_lut(d->lut, s->st);
    }
    uint32_t* dst = (uint32_t*)out;
    if (!is_device) {
        if (!d->rgba) CU(cudaMalloc((void**)&d->rgba, n * 4));
        dst = d->rgba;
    }
    s->launches += launch_render_frame(d->d[d->rd], d->lut, s->vel[s->vr], dst, d->g.w, d->g.hl, s->g.w, s->g.hg,
                                       quiver_tile, s->st);
    CU(cudaGetLastError());
    if (!is_device) {
        CU(cudaMemcpyAsync(out, d->rgba, bytes, cudaMemcpyDeviceToHost, s->st));
        CU(cudaStreamSynchronize(s->st));
    }
    return 0;
}

// ---- sync / introspection ------------------------------------------------------------------------
int natrix_sync(natrix_sim* s) {
    NEED(s, "null simulator");
    if (int rc = select_device(s)) return rc;
    if (int rc = flush_splats(s)) return rc;
    if (int rc = flush_circles(s)) return rc;
    for (natrix_dye* d : s->dyes)
        if (int rc = flush_dye(d)) return rc;
    CU(cudaStreamSynchronize(s->st));
    if (int rc = check_range_flag(s, false)) return rc;      // fetch the latest flag ...
    return check_range_flag(s, true);                        // ... and report it now
}

int natrix_stream(natrix_sim* s, void** stream) {
    NEED(s && stream, "null argument");
    *stream = (void*)s->st;
    return 0;
}

int natrix_comm_stream(natrix_sim* s, void** stream) {
    NEED(s && stream, "null argument");
    if (int rc = select_device(s)) return rc;
    if (int rc = ensure_edge_stream(s)) return rc;
    *stream = (void*)s->st_edge;
    return 0;
}

int natrix_get_timings(natrix_sim* s, float* ms, int n) {
    NEED(s && ms, "null argument");
    NEED(s->timing, "enable NATRIX_OPT_TIMING before the step");
    if (int rc = select_device(s)) return rc;
    CU(cudaStreamSynchronize(s->st));
    for (int i = 0; i < ST_COUNT; ++i) {
        float t = 0.0f;
        CU(cudaEventElapsedTime(&t, s->ev[i], s->ev[i + 1]));
        s->stage_ms[i] = t;
    }
    for (int i = 0; i < n; ++i) ms[i] = i < ST_COUNT ? s->stage_ms[i] : 0.0f;
    return 0;
}

int natrix_debug_plan_tiles(int width, int depth, int row0, int row1, const int* boxes, int nboxes, int max_tiles,
                            int* out4, int cap) {
    NEED(width >= 256 && width % 16 == 0, "the temporally blocked kernel needs width % 16 == 0 and width >= 256");
    NEED(depth >= 1 && depth <= JACOBI_TB_MAX_DEPTH && row1 > row0 && max_tiles > 0, "bad plan arguments");
    NEED((nboxes == 0 || boxes) && nboxes >= 0 && (cap == 0 || out4) && cap >= 0, "null argument");
    return jacobi_tb_plan_debug(width, depth, row0, row1, boxes, nboxes, max_tiles, out4, cap);
}

int natrix_debug_plan_stats(natrix_sim* s, unsigned long long* hits, unsigned long long* misses) {
    NEED(s && hits && misses, "null argument");
    jacobi_tb_plan_stats(s->tb, hits, misses);
    return 0;
}

int natrix_launch_count(natrix_sim* s, unsigned long long* kernels) {
    NEED(s && kernels, "null argument");
    *kernels = s->launches;
    return 0;
}

}  // extern "C"
