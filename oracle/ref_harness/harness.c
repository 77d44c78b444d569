/*
 * oracle/ref_harness/harness.c -- TEST INFRASTRUCTURE, never shipped.
 *
 * "Tier A" oracle: drives the UNMODIFIED reference translation unit
 * (/root/reference/sim_gravity.c, pulled in below by #include from where it
 * lies; nothing from it is copied into this repository) for an exact number
 * of steps and prints the resulting body state as C99 hex floats.
 *
 * How exactly-K steps are obtained without touching the reference:
 *   - sim_gravity_init() (sim_gravity.c:197) parses the scenario with the
 *     reference's own parser; its pthread_create (sim_gravity.c:441) is
 *     renamed on the command line to a no-op hook, so no worker starts yet.
 *   - we then set max_thread = N+1, start N real sim_gravity_thread(i)
 *     workers (sim_gravity.c:1130) and make this thread barrier participant
 *     id = N.  Ownership is k = id, id+max_thread, ... (sim_gravity.c:1183),
 *     so participant N owns no body.  Every step is two barrier calls
 *     (sim_gravity.c:1140 and :1226); thread 0 latches `state` only when all
 *     participants have arrived (sim_gravity.c:1301-1314), so the harness
 *     decides step by step whether the workers RUN, STOP or TERMINATE.
 *   - to read a consistent state we request STATE_STOP for one round: worker
 *     0 reaches that barrier only after its commit (sim_gravity.c:1238-1255).
 * Works for N <= MAX_THREAD-1 = 15 bodies, which covers all six scenarios; a run
 * that only asks for step 0 (the parsed initial state) works for any N.
 *
 * usage: ref_gravity <scenario> <steps> [<steps> ...]   (ascending, cumulative)
 * output: one JSON object per requested step count, one per line.
 */
#include "sim_gravity.c"

#undef pthread_create
extern int pthread_create(pthread_t *, const pthread_attr_t *, void *(*)(void *), void *);

#include "path_digest.h"

static void dump(FILE *out, const char *path, long long steps)
{
    int i;
    fprintf(out, "{\"scenario\": \"%s\", \"steps\": %lld, \"sim_time\": %lld, \"dt\": %d, \"n\": %d, \"sim_width\": \"%a\", "
                 "\"path_tail\": %lld, \"path_hash\": \"%016llx\", \"bodies\": [",
            path, steps, (long long)sim.sim_time, DT, sim.max_object, sim_width, (long long)path_tail,
            (unsigned long long)reference_path_digest());
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
    pthread_t tid[MAX_THREAD];
    long long done = 0;
    int i, a, n, me;

    if (argc < 3) {
        fprintf(stderr, "usage: %s <scenario> <steps> [<steps> ...]\n", argv[0]);
        return 2;
    }
    if (sim_gravity_init(argv[1]) != 0) {
        printf("{\"scenario\": \"%s\", \"error\": [\"%s\", \"%s\", \"%s\"]}\n",
               argv[1], error_str[0], error_str[1], error_str[2]);
        return 1;
    }
    n = sim.max_object;
    if (atoll(argv[argc - 1]) == 0) {
        /* parse only (every requested step count is 0): no workers needed, so any N the
         * reference's parser accepts (up to MAX_OBJECT) can be dumped */
        for (a = 2; a < argc; a++) dump(stdout, argv[1], 0);
        return 0;
    }
    if (n + 1 > MAX_THREAD) {
        fprintf(stderr, "ref_harness: N=%d exceeds MAX_THREAD-1\n", n);
        return 3;
    }

    max_thread = n + 1;
    me = n;
    state = STATE_STOP;
    for (i = 0; i < n; i++) {
        pthread_create(&tid[i], NULL, sim_gravity_thread, (void *)(long)i);
    }

    for (a = 2; a < argc; a++) {
        long long target = atoll(argv[a]);
        if (target < done) {
            fprintf(stderr, "ref_harness: step counts must ascend\n");
            return 2;
        }
        state = STATE_RUN;
        __sync_synchronize();
        while (done < target) {
            sim_gravity_barrier(me);   /* B1: workers see RUN and compute      */
            sim_gravity_barrier(me);   /* B2: all NEXT_* written, 0 commits    */
            done++;
        }
        state = STATE_STOP;
        __sync_synchronize();
        sim_gravity_barrier(me);       /* worker 0 arrives only after commit   */
        dump(stdout, argv[1], done);
    }

    state = STATE_TERMINATE_THREADS;
    __sync_synchronize();
    sim_gravity_barrier(me);
    for (i = 0; i < n; i++) {
        pthread_join(tid[i], NULL);
    }
    return 0;
}
