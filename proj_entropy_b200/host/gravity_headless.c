/*
 * gravity_headless.c -- headless C host for libgravity_cuda.
 *
 * This is the reference's worker loop (sim_gravity_thread,
 * sim_gravity.c:1130-1273) with the display split off: the body of
 * `if (local_state == STATE_RUN)` is one call into libgravity_cuda, and what
 * the reference keeps around it stays here in host C -- the simulation clock
 * (sim_time += DT, :1255), the PATH rings (:1248-1251, :1258-1260) and the RATE
 * Y/S readout (:912-919).  The PATH samples themselves are taken ON THE DEVICE
 * (gravity_cuda_step_path): the step kernels carry the commit block's day test
 * (:1232-1236), so a batch of any length -- also at DT = 1 day, where every commit
 * is a sample -- is one call, and the host fetches the new samples once per batch.
 *
 * usage: gravity_headless (--scenario FILE | --plummer N [--seed S] | --resume SNAPSHOT)
 *            [--steps K] [--dt SECONDS] [--gpus P] [--kernel auto|tiled|strict]
 *            [--exchange nccl|peer] [--checkpoint K1,K2,...] [--save SNAPSHOT]
 *            [--save-scenario FILE] [--path-batch ROWS] [--no-path] [--quiet]
 * prints one JSON line per checkpoint (body state as C99 hex floats) and a
 * final timing line.  --save-scenario writes the final state in the reference's
 * own scenario format (at most 200 bodies), readable by the original app.
 */
#include "gravity_cuda.h"
#include "gravity_host.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define MAX_PATH_SAMPLES (260 * 365 + 260 / 4) /* MAX_PATH, sim_gravity.c:38 */
#define YEAR_SECS (365.25 * 86400)

typedef struct { double x, y; } path_point_t;

static uint64_t microsec_now(void) /* CLOCK_MONOTONIC, as util.c:1635-1641 */
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (uint64_t)ts.tv_sec * 1000000 + (uint64_t)ts.tv_nsec / 1000;
}

static void dump_state(const char *label, long long steps, const gravity_scenario_t *s)
{
    int i;
    printf("{\"scenario\": \"%s\", \"steps\": %lld, \"sim_time\": %lld, \"dt\": %d, \"n\": %d, \"bodies\": [",
           label, steps, (long long)s->sim_time, s->dt, s->n);
    for (i = 0; i < s->n; i++) {
        printf("%s{\"name\": \"%s\", \"mass\": \"%a\", \"x\": \"%a\", \"y\": \"%a\", \"vx\": \"%a\", \"vy\": \"%a\"}",
               i ? ", " : "", s->name[i], s->mass[i], s->x[i], s->y[i], s->vx[i], s->vy[i]);
    }
    printf("]}\n");
    fflush(stdout);
}

/* Digest of the PATH rings, samples in order, bodies in index order, X then Y, each
 * double folded as a 64-bit word FNV-1a style -- the value the reference harness
 * prints for the reference's own rings (tests compare the two). */
static uint64_t path_digest(const path_point_t *path, int n, long long tail)
{
    uint64_t h = 14695981039346656037ULL;
    long long k, first = tail > MAX_PATH_SAMPLES ? tail - MAX_PATH_SAMPLES : 0;
    int i;
    if (path == NULL) return h;
    for (k = first; k < tail; k++) {
        for (i = 0; i < n; i++) {
            uint64_t w[2];
            memcpy(w, &path[(size_t)i * MAX_PATH_SAMPLES + (size_t)(k % MAX_PATH_SAMPLES)], sizeof(w));
            h = (h ^ w[0]) * 1099511628211ULL;
            h = (h ^ w[1]) * 1099511628211ULL;
        }
    }
    return h;
}

static int usage(const char *argv0)
{
    fprintf(stderr,
            "usage: %s (--scenario FILE | --plummer N [--seed S] | --resume SNAPSHOT) [--steps K] [--dt SEC]\n"
            "          [--gpus P] [--kernel auto|tiled|strict] [--exchange nccl|peer] [--checkpoint K1,K2,...]\n"
            "          [--save SNAPSHOT] [--save-scenario FILE] [--path-batch ROWS] [--no-path] [--quiet]\n",
            argv0);
    return 2;
}

int main(int argc, char **argv)
{
    const char *scenario_file = NULL, *checkpoints = NULL, *kernel = "auto", *resume_file = NULL, *save_file = NULL;
    const char *save_scenario_file = NULL, *exchange = NULL;
    long long plummer_n = 0, steps = 1000, seed = 1, done = 0, next_cp, path_batch = 4096, path_reads = 0;
    double *ring_x = NULL, *ring_y = NULL, *resume_path = NULL;
    int64_t resume_tail = 0, resume_count = 0;
    int dt_override = 0, gpus = 1, quiet = 0, want_path = 1, a, i;
    gravity_scenario_t *sim = NULL;
    gravity_cuda_t *h = NULL;
    path_point_t *path = NULL;
    long long path_tail = 0, device_tail = 0;
    int32_t last_day = 0; /* function-static in the reference, starts at 0 */
    uint64_t t0, t1;
    char *cp_cursor = NULL;

    for (a = 1; a < argc; a++) {
        if (!strcmp(argv[a], "--scenario") && a + 1 < argc) scenario_file = argv[++a];
        else if (!strcmp(argv[a], "--plummer") && a + 1 < argc) plummer_n = atoll(argv[++a]);
        else if (!strcmp(argv[a], "--seed") && a + 1 < argc) seed = atoll(argv[++a]);
        else if (!strcmp(argv[a], "--steps") && a + 1 < argc) steps = atoll(argv[++a]);
        else if (!strcmp(argv[a], "--dt") && a + 1 < argc) dt_override = atoi(argv[++a]);
        else if (!strcmp(argv[a], "--gpus") && a + 1 < argc) gpus = atoi(argv[++a]);
        else if (!strcmp(argv[a], "--kernel") && a + 1 < argc) kernel = argv[++a];
        else if (!strcmp(argv[a], "--exchange") && a + 1 < argc) exchange = argv[++a];
        else if (!strcmp(argv[a], "--checkpoint") && a + 1 < argc) checkpoints = argv[++a];
        else if (!strcmp(argv[a], "--resume") && a + 1 < argc) resume_file = argv[++a];
        else if (!strcmp(argv[a], "--save") && a + 1 < argc) save_file = argv[++a];
        else if (!strcmp(argv[a], "--save-scenario") && a + 1 < argc) save_scenario_file = argv[++a];
        else if (!strcmp(argv[a], "--path-batch") && a + 1 < argc) path_batch = atoll(argv[++a]);
        else if (!strcmp(argv[a], "--no-path")) want_path = 0;
        else if (!strcmp(argv[a], "--quiet")) quiet = 1;
        else return usage(argv[0]);
    }
    if ((scenario_file != NULL) + (plummer_n != 0) + (resume_file != NULL) != 1) return usage(argv[0]);

    /* ---- initial conditions (sim_gravity_init) ---- */
    if (scenario_file) {
        if (gravity_scenario_load(scenario_file, &sim) != 0) {
            fprintf(stderr, "ERROR %s '%s' '%s'\n", sim ? sim->error_str[0] : "Out Of Memory",
                    sim ? sim->error_str[1] : "", sim ? sim->error_str[2] : "");
            gravity_scenario_free(sim);
            return 1;
        }
    } else if (resume_file) {
        int64_t rn = 0, rt = 0;
        int32_t rdt = 0, rld = 0;
        double *m, *x, *y, *vx, *vy;
        if (gravity_snapshot_read_path(resume_file, &rn, &rt, &rdt, &rld, &m, &x, &y, &vx, &vy, &resume_tail,
                                       &resume_count, &resume_path) != 0) {
            fprintf(stderr, "ERROR cannot read snapshot '%s'\n", resume_file);
            return 1;
        }
        sim = calloc(1, sizeof(*sim));
        sim->n = (int32_t)rn;
        sim->dt = rdt;
        sim->sim_time = rt;
        last_day = rld;
        sim->sim_width = 50 * GRAVITY_AU;
        snprintf(sim->sim_name, sizeof(sim->sim_name), "SNAPSHOT");
        sim->mass = calloc((size_t)rn * 6, sizeof(double));
        sim->radius = sim->mass + rn;
        sim->x = sim->radius + rn;
        sim->y = sim->x + rn;
        sim->vx = sim->y + rn;
        sim->vy = sim->vx + rn;
        memcpy(sim->mass, m, sizeof(double) * (size_t)rn);
        memcpy(sim->x, x, sizeof(double) * (size_t)rn);
        memcpy(sim->y, y, sizeof(double) * (size_t)rn);
        memcpy(sim->vx, vx, sizeof(double) * (size_t)rn);
        memcpy(sim->vy, vy, sizeof(double) * (size_t)rn);
        free(m);
        sim->name = calloc((size_t)rn, GRAVITY_MAX_NAME);
        sim->disp_color = calloc((size_t)rn * 2, sizeof(int32_t));
        sim->max_path_disp_dflt = sim->disp_color + rn;
    } else {
        sim = calloc(1, sizeof(*sim));
        sim->n = (int32_t)plummer_n;
        sim->dt = 1;
        sim->sim_width = 50 * GRAVITY_AU;
        snprintf(sim->sim_name, sizeof(sim->sim_name), "PLUMMER_%lld", plummer_n);
        sim->mass = calloc((size_t)plummer_n * 6, sizeof(double));
        sim->radius = sim->mass + plummer_n;
        sim->x = sim->radius + plummer_n;
        sim->y = sim->x + plummer_n;
        sim->vx = sim->y + plummer_n;
        sim->vy = sim->vx + plummer_n;
        sim->name = calloc((size_t)plummer_n, GRAVITY_MAX_NAME);
        sim->disp_color = calloc((size_t)plummer_n * 2, sizeof(int32_t));
        sim->max_path_disp_dflt = sim->disp_color + plummer_n;
        if (gravity_plummer_generate(plummer_n, (uint64_t)seed, sim->mass, sim->x, sim->y, sim->vx, sim->vy) != 0) {
            fprintf(stderr, "ERROR plummer generator failed\n");
            return 1;
        }
        for (i = 0; i < sim->n; i++) {      /* display-only fields, for --save-scenario */
            sim->radius[i] = 1.0E6;
            sim->disp_color[i] = 9;
            sim->max_path_disp_dflt[i] = 30;
        }
    }
    if (dt_override > 0) sim->dt = dt_override;

    /* ---- the GPU replaces the worker pool (sim_gravity.c:425-442) ---- */
    if (gravity_cuda_create(&h, sim->n, gpus) != 0) {
        fprintf(stderr, "ERROR %s\n", gravity_cuda_last_error(NULL));
        return 1;
    }
    if (strcmp(kernel, "auto") != 0) {
        if (gravity_cuda_set_option(h, "kernel", !strcmp(kernel, "tiled") ? GRAVITY_CUDA_KERNEL_TILED
                                                                           : GRAVITY_CUDA_KERNEL_STRICT) != 0) {
            fprintf(stderr, "ERROR %s\n", gravity_cuda_last_error(h));
            return 1;
        }
    }
    if (exchange != NULL) {
        /* how the GPUs hand each other the new positions (barrier B2 + the commit copy): an NCCL
         * all-gather per step, or stores into the peers' memory fused into the step kernel */
        if (strcmp(exchange, "nccl") != 0 && strcmp(exchange, "peer") != 0) return usage(argv[0]);
        if (gravity_cuda_set_option(h, "exchange", !strcmp(exchange, "peer")) != 0) {
            fprintf(stderr, "ERROR %s\n", gravity_cuda_last_error(h));
            return 1;
        }
    }
    if (gravity_cuda_upload(h, sim->mass, sim->x, sim->y, sim->vx, sim->vy) != 0) {
        fprintf(stderr, "ERROR %s\n", gravity_cuda_last_error(h));
        return 1;
    }
    if (want_path && sim->n <= GRAVITY_MAX_OBJECT && sim->dt >= 1 && sim->sim_time >= 0) {
        if (path_batch < 3) path_batch = 3;
        if (path_batch > MAX_PATH_SAMPLES) path_batch = MAX_PATH_SAMPLES;
        path = calloc((size_t)sim->n * MAX_PATH_SAMPLES, sizeof(path_point_t));
        ring_x = malloc(sizeof(double) * (size_t)path_batch * (size_t)sim->n);
        ring_y = malloc(sizeof(double) * (size_t)path_batch * (size_t)sim->n);
        if (!path || !ring_x || !ring_y || gravity_cuda_path_enable(h, path_batch) != 0) {
            fprintf(stderr, "ERROR %s\n", gravity_cuda_last_error(h));
            return 1;
        }
        if (resume_path != NULL) {
            /* the PATH rings as they were when the snapshot was taken: the run continues them */
            long long k;
            for (k = 0; k < resume_count; k++) {
                const size_t slot = (size_t)((resume_tail - resume_count + k) % MAX_PATH_SAMPLES);
                for (i = 0; i < sim->n; i++) {
                    path[(size_t)i * MAX_PATH_SAMPLES + slot].x = resume_path[((size_t)k * sim->n + i) * 2];
                    path[(size_t)i * MAX_PATH_SAMPLES + slot].y = resume_path[((size_t)k * sim->n + i) * 2 + 1];
                }
            }
            path_tail = resume_tail;
        }
    }
    free(resume_path);

    if (checkpoints) cp_cursor = strdup(checkpoints);
    next_cp = -1;
    if (cp_cursor) {
        char *tok = strtok(cp_cursor, ",");
        next_cp = tok ? atoll(tok) : -1;
        if (next_cp == 0) {
            dump_state(sim->sim_name, 0, sim);
            tok = strtok(NULL, ",");
            next_cp = tok ? atoll(tok) : -1;
        }
    }

    /* ---- STATE_RUN ---- */
    t0 = microsec_now();
    while (done < steps) {
        long long batch = steps - done;
        if (next_cp > done && next_cp - done < batch) batch = next_cp - done;
        if (path) {
            /* no more samples per batch than the device ring has rows: k steps take at most
             * k samples, and at most k*dt/86400 + 2 */
            long long cap_steps = path_batch;
            if (sim->dt < 86400) cap_steps = (path_batch - 2) * (86400 / sim->dt);
            long long tail_after = 0;
            long long k;
            const long long tail_base = path_tail - device_tail;    /* samples taken before this process */
            if (batch > cap_steps) batch = cap_steps;
            if (gravity_cuda_step_path(h, sim->dt, batch, sim->sim_time, &last_day, (int64_t *)&tail_after) != 0) {
                fprintf(stderr, "ERROR %s\n", gravity_cuda_last_error(h));
                return 1;
            }
            /* the commit block's PATH writes (sim_gravity.c:1248-1251, :1258-1260), fetched once per batch */
            if (tail_after > device_tail) {
                const long long cnt = tail_after - device_tail;
                if (gravity_cuda_path_read(h, device_tail, cnt, ring_x, ring_y) != 0) {
                    fprintf(stderr, "ERROR %s\n", gravity_cuda_last_error(h));
                    return 1;
                }
                path_reads++;
                for (k = 0; k < cnt; k++) {
                    const size_t slot = (size_t)((path_tail + k) % MAX_PATH_SAMPLES);
                    for (i = 0; i < sim->n; i++) {
                        path[(size_t)i * MAX_PATH_SAMPLES + slot].x = ring_x[(size_t)k * sim->n + i];
                        path[(size_t)i * MAX_PATH_SAMPLES + slot].y = ring_y[(size_t)k * sim->n + i];
                    }
                }
                device_tail = tail_after;
                path_tail = tail_base + device_tail;
            }
        } else {
            if (gravity_cuda_step(h, sim->dt, batch) != 0) {
                fprintf(stderr, "ERROR %s\n", gravity_cuda_last_error(h));
                return 1;
            }
            /* the day test still runs at every commit (sim_gravity.c:1234-1236): what it leaves behind */
            last_day = (int32_t)((sim->sim_time + (batch - 1) * (long long)sim->dt) / 86400);
        }
        sim->sim_time += (long long)sim->dt * batch;            /* :1255 */
        done += batch;
        if (done == next_cp) {
            char *tok;
            if (gravity_cuda_download(h, sim->x, sim->y, sim->vx, sim->vy) != 0) return 1;
            dump_state(sim->sim_name, done, sim);
            tok = strtok(NULL, ",");
            next_cp = tok ? atoll(tok) : -1;
        }
    }
    t1 = microsec_now();

    if (save_file) {
        /* state, clock, the day memory and -- so that a resumed run continues them -- the PATH rings */
        const long long stored = path ? (path_tail < MAX_PATH_SAMPLES ? path_tail : MAX_PATH_SAMPLES) : 0;
        double *rings = stored > 0 ? malloc(sizeof(double) * 2 * (size_t)stored * (size_t)sim->n) : NULL;
        long long k;
        for (k = 0; rings != NULL && k < stored; k++) {
            const size_t slot = (size_t)((path_tail - stored + k) % MAX_PATH_SAMPLES);
            for (i = 0; i < sim->n; i++) {
                rings[((size_t)k * sim->n + i) * 2] = path[(size_t)i * MAX_PATH_SAMPLES + slot].x;
                rings[((size_t)k * sim->n + i) * 2 + 1] = path[(size_t)i * MAX_PATH_SAMPLES + slot].y;
            }
        }
        if (gravity_cuda_download(h, sim->x, sim->y, sim->vx, sim->vy) != 0 ||
            gravity_snapshot_write_path(save_file, sim->n, sim->sim_time, sim->dt, last_day, sim->mass, sim->x, sim->y,
                                        sim->vx, sim->vy, rings ? path_tail : 0, rings ? stored : 0, rings) != 0) {
            fprintf(stderr, "ERROR cannot write snapshot '%s'\n", save_file);
            return 1;
        }
        free(rings);
    }
    if (save_scenario_file) {
        int rc;
        if (gravity_cuda_download(h, sim->x, sim->y, sim->vx, sim->vy) != 0) return 1;
        rc = gravity_scenario_save(sim, save_scenario_file);
        if (rc != 0) {
            fprintf(stderr, "ERROR cannot write scenario '%s': %s\n", save_scenario_file, gravity_scenario_save_error(rc));
            return 1;
        }
    }
    if (!quiet) {
        double secs = (double)(t1 - t0) / 1e6;
        double kin = 0, pot = 0;
        long long launches = 0;
        gravity_cuda_energy(h, &kin, &pot);
        gravity_cuda_launch_count(h, (int64_t *)&launches);
        printf("{\"summary\": \"%s\", \"n\": %d, \"steps\": %lld, \"dt\": %d, \"gpus\": %d, \"wall_s\": %.6f, "
               "\"pairs_per_s\": %.6g, \"rate_years_per_s\": %.6g, \"path_samples\": %lld, \"path_hash\": \"%016llx\", "
               "\"path_reads\": %lld, \"sim_time\": %lld, \"last_day\": %d, "
               "\"kinetic\": \"%a\", \"potential\": \"%a\", \"kernel_launches\": %lld}\n",
               sim->sim_name, sim->n, done, sim->dt, gpus, secs,
               secs > 0 ? (double)sim->n * (sim->n - 1) * (double)done / secs : 0.0,
               secs > 0 ? (double)sim->dt * (double)done / YEAR_SECS / secs : 0.0, path_tail,
               (unsigned long long)path_digest(path, sim->n, path_tail), path_reads, (long long)sim->sim_time, last_day,
               kin, pot, launches);
    }

    gravity_cuda_destroy(h);
    free(path);
    free(ring_x);
    free(ring_y);
    free(cp_cursor);
    gravity_scenario_free(sim);
    return 0;
}
