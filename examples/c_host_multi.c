/* c_host_multi.c - a multi-GPU host with no Python in it: one thread per GPU, one slab handle per thread, the
 * halo exchange inside libnatrix_b200.so (natrix_comm_init + natrix_step).  It runs the demo loop of the
 * reference (demo/simulation_demo.py:220-237) on a grid cut into N row slabs and, on GPU 0, on the whole grid,
 * and compares every slab's rows of velocity, pressure and dye with the single-GPU fields byte for byte.
 *
 *   gcc -O2 -pthread -Iinclude examples/c_host_multi.c -o examples/_build/c_host_multi -Lnatrix_b200 \
 *       -lnatrix_b200 -Wl,-rpath,$PWD/natrix_b200 -lm
 *   examples/_build/c_host_multi [gpus [frames [width height]]]
 */
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "natrix_b200.h"

#define HALO 48
#define MAX_GPUS 8

typedef struct {
    int rank, world, frames, width, height;
    const unsigned char* id;
    float* vel;        /* out: this slab's rows of velocity (width * rows * 2 floats) */
    float* prs;
    float* dye;        /* dye grid = 2x the simulation grid, cut the same way */
    int row0, rows;
    int status;
    char error[512];
} rank_args;

#define RCHECK(call)                                                                          \
    do {                                                                                      \
        int rc__ = (call);                                                                    \
        if (rc__ < 0) {                                                                       \
            snprintf(a->error, sizeof(a->error), "%s failed (%d): %s", #call, rc__, natrix_last_error()); \
            a->status = rc__;                                                                 \
            return NULL;                                                                      \
        }                                                                                     \
    } while (0)

static void partition(int height, int world, int rank, int* row0, int* rows) {
    const int base = height / world, extra = height % world;
    *rows = base + (rank < extra ? 1 : 0);
    *row0 = rank * base + (rank < extra ? rank : extra);
}

/* the demo's frame loop (simulation_demo.py:220-237) on whatever handle it is given */
static int frames_loop(natrix_sim* sim, natrix_dye* dye, int frames) {
    const float dt = 1.0f / 60.0f;
    for (int k = 0; k < frames; ++k) {
        int rc;
        if ((rc = natrix_add_circle_obstacle(sim, 0.5f, 0.5f, 40.0f, 0)) < 0) return rc;
        if ((rc = natrix_step(sim, dt)) < 0) return rc;
        if ((rc = natrix_dye_step(dye, dt, 500.0f, 0.98f)) < 0) return rc;
        const float x1 = 0.5f + 0.3f * cosf(0.1f * k), y1 = 0.5f + 0.3f * sinf(0.1f * k);
        const float x0 = 0.5f + 0.3f * cosf(0.1f * (k - 1)), y0 = 0.5f + 0.3f * sinf(0.1f * (k - 1));
        if ((rc = natrix_add_velocity(sim, x1, y1, 10.0f * (x1 - x0), 10.0f * (y1 - y0), 32.0f)) < 0) return rc;
        if ((rc = natrix_dye_add(dye, x1, y1, 250.0f, 0.04f)) < 0) return rc;
    }
    return 0;
}

static void* rank_main(void* p) {
    rank_args* a = (rank_args*)p;
    natrix_sim* sim = NULL;
    natrix_dye* dye = NULL;
    int drow0, drows;
    partition(a->height, a->world, a->rank, &a->row0, &a->rows);
    partition(2 * a->height, a->world, a->rank, &drow0, &drows);
    RCHECK(natrix_create_slab(a->width, a->height, a->row0, a->rows, HALO, a->rank, &sim));
    RCHECK(natrix_set_params(sim, 500.0f, 50, 1.0f, 1.0f, 0.5, 1));
    RCHECK(natrix_comm_init(sim, a->id, a->rank, a->world));          /* collective over the rank threads */
    RCHECK(natrix_dye_create_slab(sim, 2 * a->width, 2 * a->height, drow0, drows, 2 * HALO, &dye));
    RCHECK(frames_loop(sim, dye, a->frames));
    const size_t cells = (size_t)a->width * a->rows, dcells = (size_t)2 * a->width * drows;
    a->vel = (float*)malloc(cells * 8);
    a->prs = (float*)malloc(cells * 4);
    a->dye = (float*)malloc(dcells * 4);
    RCHECK(natrix_copy_out(sim, NATRIX_VELOCITY, a->vel, cells * 8));
    RCHECK(natrix_copy_out(sim, NATRIX_PRESSURE, a->prs, cells * 4));
    RCHECK(natrix_dye_copy_out(dye, a->dye, dcells * 4));
    RCHECK(natrix_sync(sim));                                          /* reports a halo that was too small */
    RCHECK(natrix_dye_destroy(dye));
    RCHECK(natrix_destroy(sim));
    return NULL;
}

int main(int argc, char** argv) {
    const int world = argc > 1 ? atoi(argv[1]) : 2;
    const int frames = argc > 2 ? atoi(argv[2]) : 10;
    const int width = argc > 4 ? atoi(argv[3]) : 640, height = argc > 4 ? atoi(argv[4]) : 720;
    if (world < 1 || world > MAX_GPUS) { fprintf(stderr, "gpus must be 1..%d\n", MAX_GPUS); return 2; }

    /* the single-GPU run of the whole grid */
    natrix_sim* sim = NULL;
    natrix_dye* dye = NULL;
    if (natrix_create(width, height, 0, &sim) < 0 || natrix_set_params(sim, 500.0f, 50, 1.0f, 1.0f, 0.5, 1) < 0 ||
        natrix_dye_create(sim, 2 * width, 2 * height, &dye) < 0 || frames_loop(sim, dye, frames) < 0) {
        fprintf(stderr, "single-GPU run failed: %s\n", natrix_last_error());
        return 1;
    }
    const size_t cells = (size_t)width * height;
    float* vel = (float*)malloc(cells * 8);
    float* prs = (float*)malloc(cells * 4);
    float* dy = (float*)malloc(cells * 16);
    if (natrix_copy_out(sim, NATRIX_VELOCITY, vel, cells * 8) < 0 || natrix_copy_out(sim, NATRIX_PRESSURE, prs, cells * 4) < 0 ||
        natrix_dye_copy_out(dye, dy, cells * 16) < 0) {
        fprintf(stderr, "copy_out failed: %s\n", natrix_last_error());
        return 1;
    }
    natrix_dye_destroy(dye);
    natrix_destroy(sim);

    /* the same loop on `world` slabs, one host thread and one GPU each */
    unsigned char id[128];
    if (natrix_comm_unique_id(id) < 0) { fprintf(stderr, "natrix_comm_unique_id: %s\n", natrix_last_error()); return 1; }
    pthread_t th[MAX_GPUS];
    rank_args args[MAX_GPUS];
    memset(args, 0, sizeof(args));
    for (int r = 0; r < world; ++r) {
        args[r].rank = r; args[r].world = world; args[r].frames = frames; args[r].width = width; args[r].height = height;
        args[r].id = id;
        pthread_create(&th[r], NULL, rank_main, &args[r]);
    }
    int bad = 0;
    for (int r = 0; r < world; ++r) {
        pthread_join(th[r], NULL);
        if (args[r].status < 0) { fprintf(stderr, "rank %d: %s\n", r, args[r].error); bad = 1; continue; }
        int drow0, drows;
        partition(2 * height, world, r, &drow0, &drows);
        const int v = memcmp(args[r].vel, vel + (size_t)args[r].row0 * width * 2, (size_t)args[r].rows * width * 8) == 0;
        const int p = memcmp(args[r].prs, prs + (size_t)args[r].row0 * width, (size_t)args[r].rows * width * 4) == 0;
        const int d = memcmp(args[r].dye, dy + (size_t)drow0 * 2 * width, (size_t)drows * 2 * width * 4) == 0;
        printf("rank %d rows [%d, %d): velocity %s pressure %s dye %s\n", r, args[r].row0, args[r].row0 + args[r].rows,
               v ? "identical" : "DIFFERS", p ? "identical" : "DIFFERS", d ? "identical" : "DIFFERS");
        bad |= !(v && p && d);
    }
    printf("C_HOST_MULTI %s gpus=%d grid=%dx%d frames=%d\n", bad ? "FAIL" : "PASS", world, width, height, frames);
    return bad;
}
