/*
 * examples/reference_host_patch.c -- the seam of patches/sim_gravity_gpu.patch as
 * a small compilable unit for readers who do not have the reference at hand: a
 * record with the reference's object_t layout (sim_gravity.c:83-105), worker-0's
 * RUN body replaced by gravity_cuda_step_objects(), and the unchanged host
 * bookkeeping (day test, PATH ring, sim_time; :1229-1260).  The patch itself is
 * applied to the REAL sim_gravity.c and run on the GPU by tests/test_reference_patch.py
 * (INTEGRATION.md section 2); tests/test_integration_snippet.py compiles and links
 * this unit against libgravity_cuda and, with a GPU, runs two_body-like data through it.
 */
#include "gravity_cuda.h"

#include <stdbool.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define MAX_OBJECT 200
#define MAX_NAME 32
#define MAX_PATH (260 * 365 + 260 / 4)

typedef struct {                 /* same field order and types as the reference */
    char    NAME[MAX_NAME];
    double  X, Y, VX, VY, MASS, RADIUS;
    double  NEXT_X, NEXT_Y, NEXT_VX, NEXT_VY;
    int32_t DISP_COLOR, DISP_R, MAX_PATH_DISP_DFLT;
    struct path_s { double X, Y; } PATH[MAX_PATH];
} object_t;

typedef struct {
    char      sim_name[MAX_NAME];
    int64_t   sim_time;
    int32_t   max_object;
    object_t *object[MAX_OBJECT];
} sim_t;

static sim_t sim;
static int32_t DT = 1;
static volatile int64_t path_tail;
static gravity_cuda_t *gpu;

_Static_assert(offsetof(object_t, X) == 32 && offsetof(object_t, Y) == 40 && offsetof(object_t, VX) == 48 &&
               offsetof(object_t, VY) == 56 && offsetof(object_t, MASS) == 64, "reference object_t layout");

/* the body of `if (local_state == STATE_RUN)` after the patch */
static int run_one_step(void)
{
    static int32_t last_day;
    int32_t dt = DT, day, k;
    bool day_changed;

    if (gravity_cuda_step_objects(gpu, (void **)sim.object, sim.max_object, offsetof(object_t, X),
                                  offsetof(object_t, Y), offsetof(object_t, VX), offsetof(object_t, VY),
                                  offsetof(object_t, MASS), dt, 1) != 0) {
        fprintf(stderr, "ERROR %s\n", gravity_cuda_last_error(gpu));
        return -1;
    }
    day = sim.sim_time / 86400;
    day_changed = day != last_day;
    last_day = day;
    if (day_changed) {
        for (k = 0; k < sim.max_object; k++) {
            object_t *obj = sim.object[k];
            obj->PATH[path_tail % MAX_PATH].X = obj->X;
            obj->PATH[path_tail % MAX_PATH].Y = obj->Y;
        }
    }
    sim.sim_time += dt;
    if (day_changed) __sync_fetch_and_add(&path_tail, 1);
    return 0;
}

int main(int argc, char **argv)
{
    const int steps = argc > 1 ? atoi(argv[1]) : 3;
    int i;
    /* two_body_circular-like initial conditions, written out rather than parsed */
    static const double init[2][5] = {{-1.4E11, 0, 0, -8.8E3, 2000000E24}, {1.4E11, 0, 0, 17.6E3, 1000000E24}};
    sim.max_object = 2;
    for (i = 0; i < 2; i++) {
        object_t *o = calloc(1, sizeof(object_t));
        snprintf(o->NAME, MAX_NAME, "body%d", i + 1);
        o->X = init[i][0]; o->Y = init[i][1]; o->VX = init[i][2]; o->VY = init[i][3]; o->MASS = init[i][4];
        sim.object[i] = o;
    }
    if (gravity_cuda_create(&gpu, sim.max_object, 1) != 0) {
        fprintf(stderr, "GPU Init Failed: %s\n", gravity_cuda_last_error(NULL));
        return 1;
    }
    {
        struct timespec t0, t1;
        clock_gettime(CLOCK_MONOTONIC, &t0);
        for (i = 0; i < steps; i++) {
            if (run_one_step() != 0) return 1;
        }
        clock_gettime(CLOCK_MONOTONIC, &t1);
        /* what one drop-in call costs: gather, upload, step, download, scatter */
        fprintf(stderr, "%d steps, %.1f us per gravity_cuda_step_objects call\n", steps,
                ((t1.tv_sec - t0.tv_sec) * 1e9 + (t1.tv_nsec - t0.tv_nsec)) / 1e3 / (steps > 0 ? steps : 1));
    }
    for (i = 0; i < 2; i++) {
        printf("%s %a %a %a %a\n", sim.object[i]->NAME, sim.object[i]->X, sim.object[i]->Y, sim.object[i]->VX,
               sim.object[i]->VY);
    }
    gravity_cuda_destroy(gpu);
    return 0;
}
