/* CPU oracle (C / OpenMP) for the Natrix per-step stable-fluids pipeline.
 *
 * TEST INFRASTRUCTURE ONLY - never linked into, loaded by or called from the product
 * (natrix_b200/).  Two jobs: (1) an independently written second restatement that must
 * agree BIT FOR BIT with oracle/natrix_oracle.py (tests/test_oracle.py), (2) the
 * multi-threaded CPU baseline that bench.py times ("cpu_baseline", "--impl reference").
 *
 * PARITY UNPINNED: the reference has no golden vectors / KATs for this path and its
 * bgfx engine cannot run here; see the header of natrix_oracle.py.
 *
 * Each routine cites the reference shader it restates (paths relative to the reference
 * root; "core/" = natrix/core/shaders/originals/, "demo/" = demo/shaders/).
 * Build: see oracle/Makefile (gcc -O2 -fopenmp -ffp-contract=off: IEEE float32, no FMA).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct nox_sim {
    int w, h;
    float *vel[2];   /* float2 per cell, ping-pong   (fluid_simulator.py:357-361) */
    float *p[2];     /* float per cell, ping-pong    (:362-365, scalar pressure)  */
    float *div, *vort;
    float *obs;      /* float2 per cell */
    int vr, pr;      /* VELOCITY_READ / PRESSURE_READ (:16-20) */
    float speed, dissipation, vorticity, alpha, rbeta;
    int iterations, has_borders, viscous;
} nox_sim;

typedef struct nox_dye {
    nox_sim *sim;
    int w, h;
    float *d[2];
    int rd;
} nox_dye;

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
static inline float clampf(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }
static inline float mixf(float a, float b, float t) { return a * (1.0f - t) + b * t; }
static inline int is_solid(const float *obs, size_t i) { return obs[2 * i] > 0.0f || obs[2 * i + 1] > 0.0f; }

/* common.sh:9-19 */
#define NEIGHBOURS(x, y, w, h)                                   \
    size_t nL = (size_t)(y) * (w) + clampi((x)-1, 0, (w)-1);     \
    size_t nR = (size_t)(y) * (w) + clampi((x) + 1, 0, (w)-1);   \
    size_t nB = (size_t)clampi((y)-1, 0, (h)-1) * (w) + (x);     \
    size_t nT = (size_t)clampi((y) + 1, 0, (h)-1) * (w) + (x);

void nox_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
int nox_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

nox_sim *nox_create(int w, int h) {
    nox_sim *s = (nox_sim *)calloc(1, sizeof(nox_sim));
    size_t c = (size_t)w * h;
    s->w = w; s->h = h;
    for (int i = 0; i < 2; ++i) {
        s->vel[i] = (float *)calloc(2 * c, sizeof(float));
        s->p[i] = (float *)calloc(c, sizeof(float));
    }
    s->div = (float *)calloc(c, sizeof(float));
    s->vort = (float *)calloc(c, sizeof(float));
    s->obs = (float *)calloc(2 * c, sizeof(float));
    s->vr = 0; s->pr = 0;
    s->speed = 500.0f; s->dissipation = 1.0f; s->vorticity = 0.0f;     /* :28-35 */
    s->iterations = 50; s->has_borders = 1;
    s->viscous = 1; s->alpha = (float)(1.0 / 0.1); s->rbeta = (float)(1.0 / (4.0 + 1.0 / 0.1));
    return s;
}

void nox_destroy(nox_sim *s) {
    if (!s) return;
    for (int i = 0; i < 2; ++i) { free(s->vel[i]); free(s->p[i]); }
    free(s->div); free(s->vort); free(s->obs); free(s);
}

/* fluid_simulator.py:315-336 - alpha / rBeta derived in double, narrowed to float */
void nox_set_params(nox_sim *s, float speed, int iterations, float dissipation, float vorticity,
                    double viscosity, int has_borders) {
    s->speed = speed; s->iterations = iterations; s->dissipation = dissipation;
    s->vorticity = vorticity; s->has_borders = has_borders;
    s->viscous = viscosity > 0.0;
    if (s->viscous) {
        double centre = 1.0 / viscosity;
        s->alpha = (float)centre;
        s->rbeta = (float)(1.0 / (4.0 + centre));
    }
}

/* field ids shared with include/natrix_b200.h */
float *nox_field(nox_sim *s, int field) {
    switch (field) {
    case 0: return s->vel[s->vr];
    case 1: return s->p[s->pr];
    case 2: return s->div;
    case 3: return s->vort;
    case 4: return s->obs;
    default: return NULL;
    }
}

/* core/shader.AddVelocity.comp:26-35 */
void nox_add_velocity(nox_sim *s, float px, float py, float vx, float vy, float radius) {
    const int w = s->w, h = s->h;
    const float *in = s->vel[s->vr];
    float *out = s->vel[1 - s->vr];
    const float sx = px * (float)w, sy = py * (float)h;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            size_t i = (size_t)y * w + x;
            float a = in[2 * i], b = in[2 * i + 1];
            float ex = sx - (float)x, ey = sy - (float)y;
            float len = sqrtf(ex * ex + ey * ey);
            if (len <= radius) {
                float fall = radius - len;
                a = a + vx * fall / radius;
                b = b + vy * fall / radius;
            }
            out[2 * i] = clampf(a, -1.0f, 1.0f);
            out[2 * i + 1] = clampf(b, -1.0f, 1.0f);
        }
    s->vr = 1 - s->vr;
}

/* core/shader.AddCircleObstacle.comp:24-36 */
void nox_add_circle_obstacle(nox_sim *s, float px, float py, float radius, int is_static) {
    (void)is_static;
    const int w = s->w, h = s->h;
    const float sx = px * (float)w, sy = py * (float)h;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            float ex = sx - (float)x, ey = sy - (float)y;
            if (sqrtf(ex * ex + ey * ey) <= radius) {
                size_t i = (size_t)y * w + x;
                s->obs[2 * i] = 1.0f; s->obs[2 * i + 1] = 0.0f;
            }
        }
}

/* core/shader.AddTriangleObstacle.comp:19-51 */
static inline float tri_sign(float ax, float ay, float bx, float by, float cx, float cy) {
    return ((ax - cx) * (by - cy)) - ((bx - cx) * (ay - cy));
}
void nox_add_triangle_obstacle(nox_sim *s, float p1x, float p1y, float p2x, float p2y, float p3x,
                               float p3y, int is_static) {
    const int w = s->w, h = s->h;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            float tx = (float)x / (float)w, ty = (float)y / (float)h;
            int b1 = tri_sign(tx, ty, p1x, p1y, p2x, p2y) < 0.0f;
            int b2 = tri_sign(tx, ty, p2x, p2y, p3x, p3y) < 0.0f;
            int b3 = tri_sign(tx, ty, p3x, p3y, p1x, p1y) < 0.0f;
            if (b1 == b2 && b2 == b3) {
                size_t i = (size_t)y * w + x;
                s->obs[2 * i] = is_static ? 0.0f : 1.0f;
                s->obs[2 * i + 1] = is_static ? 1.0f : 0.0f;
            }
        }
}

/* clamped floor/ceil corners with unclamped delta: core/shader.AdvectVelocity.comp:38-42 */
typedef struct { int tx, ty, bx, by; float dx, dy; } corners_t;
static inline corners_t corners(float fx, float fy, int w, int h) {
    corners_t c;
    float mx = (float)(w - 1), my = (float)(h - 1);
    c.tx = (int)clampf(ceilf(fx), 0.0f, mx);
    c.ty = (int)clampf(ceilf(fy), 0.0f, my);
    c.bx = (int)clampf(floorf(fx), 0.0f, mx);
    c.by = (int)clampf(floorf(fy), 0.0f, my);
    c.dx = fx - (float)c.bx;
    c.dy = fy - (float)c.by;
    return c;
}

static void st_init_boundaries(nox_sim *s) { /* core/shader.InitBoundaries.comp:14-34 */
    const int w = s->w, h = s->h;
    float *v = s->vel[s->vr];
    for (int x = 0; x < w; ++x) {
        size_t a = x, b = (size_t)(h - 1) * w + x;
        v[2 * a] = v[2 * a + 1] = 0.0f; v[2 * b] = v[2 * b + 1] = 0.0f;
    }
    for (int y = 0; y < h; ++y) {
        size_t a = (size_t)y * w, b = (size_t)y * w + (w - 1);
        v[2 * a] = v[2 * a + 1] = 0.0f; v[2 * b] = v[2 * b + 1] = 0.0f;
    }
}

static void st_advect(nox_sim *s, float dt) { /* core/shader.AdvectVelocity.comp:27-50 */
    const int w = s->w, h = s->h;
    const float *in = s->vel[s->vr];
    float *out = s->vel[1 - s->vr];
    const float speed = s->speed, diss = s->dissipation;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            size_t i = (size_t)y * w + x;
            if (is_solid(s->obs, i)) { out[2 * i] = 0.0f; out[2 * i + 1] = 0.0f; continue; }
            float fx = (float)x - in[2 * i] * dt * speed;
            float fy = (float)y - in[2 * i + 1] * dt * speed;
            corners_t c = corners(fx, fy, w, h);
            size_t lt = (size_t)c.ty * w + c.bx, rt = (size_t)c.ty * w + c.tx;
            size_t lb = (size_t)c.by * w + c.bx, rb = (size_t)c.by * w + c.tx;
            for (int k = 0; k < 2; ++k) {
                float h1 = mixf(in[2 * lt + k], in[2 * rt + k], c.dx);
                float h2 = mixf(in[2 * lb + k], in[2 * rb + k], c.dx);
                out[2 * i + k] = clampf(mixf(h2, h1, c.dy) * diss, -1.0f, 1.0f);
            }
        }
    s->vr = 1 - s->vr;
}

static void st_vorticity(nox_sim *s) { /* core/shader.CalcVorticity.comp:20-26 */
    const int w = s->w, h = s->h;
    const float *v = s->vel[s->vr];
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            NEIGHBOURS(x, y, w, h)
            s->vort[(size_t)y * w + x] =
                0.5f * ((v[2 * nR + 1] - v[2 * nL + 1]) - (v[2 * nT] - v[2 * nB]));
        }
}

static void st_confinement(nox_sim *s, float dt) { /* core/shader.ApplyVorticity.comp:26-39 */
    const int w = s->w, h = s->h;
    const float *in = s->vel[s->vr], *o = s->vort;
    float *out = s->vel[1 - s->vr];
    const float scale = s->vorticity;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            size_t i = (size_t)y * w + x;
            NEIGHBOURS(x, y, w, h)
            float fx = 0.5f * (fabsf(o[nT]) - fabsf(o[nB]));
            float fy = 0.5f * (fabsf(o[nR]) - fabsf(o[nL]));
            float m = fmaxf(2.4414e-4f, fx * fx + fy * fy);
            float inv = 1.0f / sqrtf(m);
            fx = fx * inv; fy = fy * inv;
            float k = scale * o[i];
            fx = fx * k; fy = fy * (-k);
            out[2 * i] = in[2 * i] + fx * dt;
            out[2 * i + 1] = in[2 * i + 1] + fy * dt;
        }
    s->vr = 1 - s->vr;
}

static void st_viscosity(nox_sim *s) { /* core/shader.Viscosity.comp:24-31 */
    const int w = s->w, h = s->h;
    const float *in = s->vel[s->vr];
    float *out = s->vel[1 - s->vr];
    const float alpha = s->alpha, rbeta = s->rbeta;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            size_t i = (size_t)y * w + x;
            NEIGHBOURS(x, y, w, h)
            for (int k = 0; k < 2; ++k)
                out[2 * i + k] = (in[2 * nL + k] + in[2 * nR + k] + in[2 * nB + k] + in[2 * nT + k] +
                                  in[2 * i + k] * alpha) * rbeta;
        }
    s->vr = 1 - s->vr;
}

static void st_divergence(nox_sim *s) { /* core/shader.Divergence.comp:22-40 */
    const int w = s->w, h = s->h;
    const float *v = s->vel[s->vr];
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            NEIGHBOURS(x, y, w, h)
            float x1 = is_solid(s->obs, nL) ? 0.0f : v[2 * nL];
            float x2 = is_solid(s->obs, nR) ? 0.0f : v[2 * nR];
            float y1 = is_solid(s->obs, nB) ? 0.0f : v[2 * nB + 1];
            float y2 = is_solid(s->obs, nT) ? 0.0f : v[2 * nT + 1];
            s->div[(size_t)y * w + x] = 0.5f * ((x2 - x1) + (y2 - y1));
        }
}

static void st_poisson(nox_sim *s) { /* core/shader.Poisson.comp:24-37 */
    const int w = s->w, h = s->h;
    const float *in = s->p[s->pr];
    float *out = s->p[1 - s->pr];
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            size_t i = (size_t)y * w + x;
            NEIGHBOURS(x, y, w, h)
            float p = in[i];
            float x1 = is_solid(s->obs, nL) ? p : in[nL];
            float x2 = is_solid(s->obs, nR) ? p : in[nR];
            float y1 = is_solid(s->obs, nB) ? p : in[nB];
            float y2 = is_solid(s->obs, nT) ? p : in[nT];
            out[i] = (x1 + x2 + y1 + y2 - s->div[i]) * 0.25f;
        }
    s->pr = 1 - s->pr;
}

static void st_gradient(nox_sim *s) { /* core/shader.SubtractGradient.comp:24-46 */
    const int w = s->w, h = s->h;
    const float *p = s->p[s->pr], *in = s->vel[s->vr];
    float *out = s->vel[1 - s->vr];
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            size_t i = (size_t)y * w + x;
            NEIGHBOURS(x, y, w, h)
            float c = p[i];
            float x1 = is_solid(s->obs, nL) ? c : p[nL];
            float x2 = is_solid(s->obs, nR) ? c : p[nR];
            float y1 = is_solid(s->obs, nB) ? c : p[nB];
            float y2 = is_solid(s->obs, nT) ? c : p[nT];
            out[2 * i] = in[2 * i] - 0.5f * (x2 - x1);
            out[2 * i + 1] = in[2 * i + 1] - 0.5f * (y2 - y1);
        }
    s->vr = 1 - s->vr;
}

/* natrix/core/fluid_simulator.py:174-280 */
void nox_step(nox_sim *s, float dt) {
    size_t c = (size_t)s->w * s->h;
    if (s->has_borders) st_init_boundaries(s);
    st_advect(s, dt);
    st_vorticity(s);
    st_confinement(s, dt);
    if (s->viscous) st_viscosity(s);
    st_divergence(s);
    memset(s->p[s->pr], 0, c * sizeof(float));
    for (int i = 0; i < s->iterations; ++i) st_poisson(s);
    st_gradient(s);
    memset(s->obs, 0, 2 * c * sizeof(float));
}

/* only the Jacobi loop: used by bench.py to time a bounded sample of the hot loop */
void nox_poisson_sweeps(nox_sim *s, int n) { for (int i = 0; i < n; ++i) st_poisson(s); }

/* ------------------------------------------------------------------ dye ("particles") */
nox_dye *nox_dye_create(nox_sim *s, int w, int h) {
    nox_dye *d = (nox_dye *)calloc(1, sizeof(nox_dye));
    d->sim = s; d->w = w; d->h = h;
    d->d[0] = (float *)calloc((size_t)w * h, sizeof(float));
    d->d[1] = (float *)calloc((size_t)w * h, sizeof(float));
    return d;
}
void nox_dye_destroy(nox_dye *d) { if (d) { free(d->d[0]); free(d->d[1]); free(d); } }
float *nox_dye_field(nox_dye *d) { return d->d[d->rd]; }

/* demo/shader.AddParticle.comp:25-34 */
void nox_dye_add(nox_dye *d, float px, float py, float radius, float value) {
    const int w = d->w, h = d->h;
    const float *in = d->d[d->rd];
    float *out = d->d[1 - d->rd];
    const float sx = px * (float)w, sy = py * (float)h;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            size_t i = (size_t)y * w + x;
            float ex = sx - (float)x, ey = sy - (float)y;
            float len = sqrtf(ex * ex + ey * ey);
            float v = in[i];
            if (len <= radius) v = clampf(v + value * (radius - len) / radius, 0.0f, 255.0f);
            out[i] = v;
        }
    d->rd = 1 - d->rd;
}

/* demo/shader.AdvectParticle.comp:21-70 */
void nox_dye_step(nox_dye *d, float dt, float speed, float dissipation) {
    const int pw = d->w, ph = d->h, vw = d->sim->w, vh = d->sim->h;
    const float *in = d->d[d->rd], *vel = d->sim->vel[d->sim->vr], *obs = d->sim->obs;
    float *out = d->d[1 - d->rd];
    const float rx = (float)pw / (float)vw, ry = (float)ph / (float)vh;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < ph; ++y)
        for (int x = 0; x < pw; ++x) {
            size_t i = (size_t)y * pw + x;
            float nx = ((float)x / (float)pw) * (float)vw;
            float ny = ((float)y / (float)ph) * (float)vh;
            size_t oi = (size_t)(unsigned)ny * vw + (unsigned)nx;
            if (is_solid(obs, oi)) { out[i] = 0.0f; continue; }
            corners_t c = corners(nx, ny, vw, vh);
            size_t lt = (size_t)c.ty * vw + c.bx, rt = (size_t)c.ty * vw + c.tx;
            size_t lb = (size_t)c.by * vw + c.bx, rb = (size_t)c.by * vw + c.tx;
            float v[2];
            for (int k = 0; k < 2; ++k) {
                float h1 = mixf(vel[2 * lt + k], vel[2 * rt + k], c.dx);
                float h2 = mixf(vel[2 * lb + k], vel[2 * rb + k], c.dx);
                v[k] = mixf(h2, h1, c.dy) * (k ? ry : rx);
            }
            float fx = (float)x - v[0] * dt * speed;
            float fy = (float)y - v[1] * dt * speed;
            corners_t q = corners(fx, fy, pw, ph);
            float g1 = mixf(in[(size_t)q.ty * pw + q.bx], in[(size_t)q.ty * pw + q.tx], q.dx);
            float g2 = mixf(in[(size_t)q.by * pw + q.bx], in[(size_t)q.by * pw + q.tx], q.dx);
            out[i] = mixf(g2, g1, q.dy) * dissipation;
        }
    d->rd = 1 - d->rd;
}
