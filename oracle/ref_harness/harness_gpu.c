/*
 * oracle/ref_harness/harness_gpu.c -- TEST INFRASTRUCTURE, never shipped.
 *
 * Drives the reference's own translation unit WITH patches/sim_gravity_gpu.patch
 * applied (a temporary copy made at build time by tools/apply_reference_patch.py;
 * nothing from the reference is kept in this repository) against libgravity_cuda:
 * the reference's sim_gravity_init() parses the scenario, creates the GPU handle
 * and starts its one worker; the reference's sim_gravity_thread() calls
 * gravity_cuda_step_objects() where its k-loop used to be and then runs its own,
 * unmodified commit block (day test, PATH ring, sim_time); sim_gravity_terminate()
 * joins the worker and destroys the handle.  max_thread is 1, as in the patch.
 *
 * Metering exactly K steps without touching the patched source: two names are
 * redirected on the compiler command line, like pthread_create in the CPU harness:
 *   gravity_cuda_step_objects -> hook_step_objects: calls the real entry point and,
 *       once the requested number of steps is reached, sets state = STATE_STOP
 *       from inside the worker itself, before its commit block runs;
 *   usleep -> hook_usleep: the worker only sleeps in STATE_STOP
 *       (sim_gravity.c:1263-1264), i.e. after the commit block of the last step.
 *
 * usage: ref_gravity_gpu <scenario> <steps> [<steps> ...]   (ascending, cumulative)
 * output: the same JSON lines as ref_gravity (harness.c).
 */
#include <unistd.h>
#include "sim_gravity_gpu_patched.c"

#undef gravity_cuda_step_objects
#undef usleep
extern int gravity_cuda_step_objects(gravity_cuda_t *h, void **objects, int64_t n, size_t off_x, size_t off_y,
                                     size_t off_vx, size_t off_vy, size_t off_mass, int32_t dt, int64_t nsteps);
extern int usleep(useconds_t usec);

static volatile long long steps_done, steps_target, shim_calls;
static volatile long long idle_calls;
static volatile int       step_failed;

/* GPU_BATCHED: the unit was patched with patches/sim_gravity_gpu_batched.patch, whose worker
 * asks for a whole batch of steps per shim call (up to the next PATH sample, at most
 * gpu_batch_limit); the harness meters by setting that limit to the steps still wanted. */
#ifdef GPU_BATCHED
#define SET_BATCH_LIMIT(n) (gpu_batch_limit = (n) > 0 ? (n) : 1)
#else
#define SET_BATCH_LIMIT(n) ((void)0)
#endif

int hook_step_objects(gravity_cuda_t *h, void **objects, int64_t n, size_t off_x, size_t off_y, size_t off_vx,
                      size_t off_vy, size_t off_mass, int32_t dt, int64_t nsteps)
{
    int rc = gravity_cuda_step_objects(h, objects, n, off_x, off_y, off_vx, off_vy, off_mass, dt, nsteps);
    if (rc != 0) {
        step_failed = 1;
        return rc;
    }
    shim_calls++;
    steps_done += nsteps;
    if (steps_done >= steps_target) state = STATE_STOP;     /* the commit block of this step still runs */
    SET_BATCH_LIMIT(steps_target - steps_done);
    return rc;
}

int hook_usleep(useconds_t usec)
{
    (void)usec;
    __sync_fetch_and_add(&idle_calls, 1);
    return usleep(100);
}

#include "path_digest.h"

static void dump(FILE *out, const char *path, long long steps)
{
    int i;
    fprintf(out, "{\"scenario\": \"%s\", \"steps\": %lld, \"sim_time\": %lld, \"dt\": %d, \"n\": %d, \"sim_width\": \"%a\", "
                 "\"path_tail\": %lld, \"path_hash\": \"%016llx\", \"max_thread\": %d, \"shim_calls\": %lld, \"bodies\": [",
            path, steps, (long long)sim.sim_time, DT, sim.max_object, sim_width, (long long)path_tail,
            (unsigned long long)reference_path_digest(), max_thread, (long long)shim_calls);
    for (i = 0; i < sim.max_object; i++) {
        object_t *o = sim.object[i];
        fprintf(out, "%s{\"name\": \"%s\", \"mass\": \"%a\", \"radius\": \"%a\", \"x\": \"%a\", \"y\": \"%a\", \"vx\": \"%a\", \"vy\": \"%a\"}",
                i ? ", " : "", o->NAME, o->MASS, o->RADIUS, o->X, o->Y, o->VX, o->VY);
    }
    fprintf(out, "]}\n");
    fflush(out);
}

int main(int argc, char **argv)
{
    int a;

    if (argc < 3) {
        fprintf(stderr, "usage: %s <scenario> <steps> [<steps> ...]\n", argv[0]);
        return 2;
    }
    if (sim_gravity_init(argv[1]) != 0) {      /* parser + gravity_cuda_create + worker, all the reference's code */
        printf("{\"scenario\": \"%s\", \"error\": [\"%s\", \"%s\", \"%s\"]}\n",
               argv[1], error_str[0], error_str[1], error_str[2]);
        return 1;
    }
    for (a = 2; a < argc; a++) {
        long long target = atoll(argv[a]), seen;
        if (target < steps_done) {
            fprintf(stderr, "ref_harness: step counts must ascend\n");
            return 2;
        }
        if (target > steps_done && sim.max_object > 0) {
            steps_target = target;
            SET_BATCH_LIMIT(target - steps_done);
            __sync_synchronize();
            state = STATE_RUN;                 /* what the RUN button does, sim_gravity.c:953-955 */
            while (steps_done < target && !step_failed) usleep(50);
            if (step_failed) {
                printf("{\"scenario\": \"%s\", \"error\": [\"GPU Step Failed\", \"%s\", \"\"]}\n", argv[1],
                       gravity_cuda_last_error(gpu));
                sim_gravity_terminate();
                return 1;
            }
        }
        /* the worker sleeps only in STATE_STOP, after the commit block of the last step */
        seen = idle_calls;
        while (idle_calls < seen + 2) usleep(50);
        dump(stdout, argv[1], steps_done);
    }
    sim_gravity_terminate();                   /* joins the worker, destroys the GPU handle */
    return 0;
}
