/* c_host.c - a host with no Python in it: the demo loop of the reference (demo/simulation_demo.py:220-237)
 * driven through the C ABI of libnatrix_b200.so alone.  Built by __graft_entry__.build() (plain gcc, no
 * CUDA headers needed) and run by tests/test_gpu_parity.py, which compares the printed field statistics
 * with the same loop driven through the ctypes mirror.
 *
 *   gcc -O2 -Iinclude examples/c_host.c -o examples/_build/c_host -Lnatrix_b200 -lnatrix_b200 \
 *       -Wl,-rpath,$PWD/natrix_b200
 *   examples/_build/c_host [frames]
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "natrix_b200.h"

#define CHECK(call)                                                                   \
    do {                                                                              \
        int rc__ = (call);                                                            \
        if (rc__ < 0) {                                                               \
            fprintf(stderr, "%s failed (%d): %s\n", #call, rc__, natrix_last_error()); \
            return 1;                                                                 \
        }                                                                             \
    } while (0)

int main(int argc, char** argv) {
    const int frames = argc > 1 ? atoi(argv[1]) : 20;
    const float dt = 1.0f / 60.0f;
    natrix_sim* sim = NULL;
    natrix_dye* dye = NULL;
    CHECK(natrix_create(640, 360, 0, &sim));                                  /* simulation_demo.py:100-102 */
    CHECK(natrix_set_params(sim, 500.0f, 50, 1.0f, 1.0f, 0.5, 1));             /* :103-105 + defaults */
    CHECK(natrix_dye_create(sim, 1280, 720, &dye));                           /* :107-110 */
    for (int k = 0; k < frames; ++k) {
        CHECK(natrix_add_circle_obstacle(sim, 0.5f, 0.5f, 40.0f, 0));         /* :220 */
        CHECK(natrix_step(sim, dt));                                          /* :222 */
        CHECK(natrix_dye_step(dye, dt, 500.0f, 0.98f));                       /* :223 */
        /* a scripted "mouse drag" on a circle, velocity = 10 x the position delta (:225-235) */
        const float x1 = 0.5f + 0.3f * cosf(0.1f * k), y1 = 0.5f + 0.3f * sinf(0.1f * k);
        const float x0 = 0.5f + 0.3f * cosf(0.1f * (k - 1)), y0 = 0.5f + 0.3f * sinf(0.1f * (k - 1));
        CHECK(natrix_add_velocity(sim, x1, y1, 10.0f * (x1 - x0), 10.0f * (y1 - y0), 32.0f));
        CHECK(natrix_dye_add(dye, x1, y1, 250.0f, 0.04f));
    }
    double v[4], p[4], d[4];
    CHECK(natrix_field_stats(sim, NATRIX_VELOCITY, v));
    CHECK(natrix_field_stats(sim, NATRIX_PRESSURE, p));
    CHECK(natrix_dye_stats(dye, d));
    unsigned long long launches = 0;
    CHECK(natrix_launch_count(sim, &launches));
    printf("version %s\nframes %d launches %llu\n", natrix_version(), frames, launches);
    printf("velocity %.17g %.17g %.17g %.17g\n", v[0], v[1], v[2], v[3]);
    printf("pressure %.17g %.17g %.17g %.17g\n", p[0], p[1], p[2], p[3]);
    printf("dye %.17g %.17g %.17g %.17g\n", d[0], d[1], d[2], d[3]);
    CHECK(natrix_dye_destroy(dye));
    CHECK(natrix_destroy(sim));
    return 0;
}
