/*
 * oracle/gravity_oracle.c -- TEST INFRASTRUCTURE.  Not part of the product.
 * See gravity_oracle.h for scope and parity status (PINNED against the
 * unmodified reference translation unit via tests/golden/).
 *
 * Build with the reference's arithmetic: -O2, no -ffast-math, no -march
 * (/root/reference/Makefile:17-19), plus -ffp-contract=off so that "no FMA"
 * is explicit rather than a property of the default x86-64 target.
 */
#include "gravity_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ---- one row: sim_gravity.c:1186-1208 ---------------------------------- */

/* The evaluated body is half-drifted (sim_gravity.c:1186-1187), every other
 * body is read un-drifted (:1196-1197); self is skipped by index (:1191) and
 * exactly coincident pairs are skipped (:1199-1201).  G multiplies the
 * finished sum once (:1207-1208).  Sum order is j ascending. */
static void row_accel(int64_t n, const double *mass, const double *x, const double *y,
                      const double *vx, const double *vy, int32_t dt, int64_t k,
                      double *ax, double *ay)
{
    const double px = x[k] + vx[k] * dt / 2;
    const double py = y[k] + vy[k] * dt / 2;
    double sx = 0, sy = 0;
    int64_t j;

    for (j = 0; j < n; j++) {
        double m, dx, dy, r, t;
        if (j == k) continue;
        m  = mass[j];
        dx = x[j] - px;
        dy = y[j] - py;
        r  = sqrt(dx * dx + dy * dy);
        if (r == 0) continue;
        t  = m / (r * r * r);
        sx = sx + t * dx;
        sy = sy + t * dy;
    }
    *ax = ORACLE_G * sx;
    *ay = ORACLE_G * sy;
}

static void row_accel_ld(int64_t n, const double *mass, const double *x, const double *y,
                         const double *vx, const double *vy, int32_t dt, int64_t k,
                         double *ax, double *ay)
{
    /* the half-drifted position is formed in double exactly as the reference
     * forms it: it is an input of the sum, not part of the sum's error. */
    const double px = x[k] + vx[k] * dt / 2;
    const double py = y[k] + vy[k] * dt / 2;
    long double sx = 0, sy = 0;
    int64_t j;

    for (j = 0; j < n; j++) {
        long double dx, dy, r, t;
        if (j == k) continue;
        dx = (long double)x[j] - px;
        dy = (long double)y[j] - py;
        r  = sqrtl(dx * dx + dy * dy);
        if (r == 0) continue;
        t  = (long double)mass[j] / (r * r * r);
        sx += t * dx;
        sy += t * dy;
    }
    *ax = (double)((long double)ORACLE_G * sx);
    *ay = (double)((long double)ORACLE_G * sy);
}

/* ---- row-parallel driver ------------------------------------------------ */

typedef void (*row_fn)(int64_t, const double *, const double *, const double *,
                       const double *, const double *, int32_t, int64_t, double *, double *);

typedef struct {
    row_fn fn;
    int64_t n;
    const double *mass, *x, *y, *vx, *vy;
    int32_t dt;
    const int64_t *rows;
    int64_t nrows;
    double *ax, *ay;
    int tid, nthreads;
} accel_job_t;

static void *accel_worker(void *arg)
{
    accel_job_t *jb = arg;
    int64_t r;
    /* strided ownership, as k = thread_id; k += max_thread (sim_gravity.c:1183) */
    for (r = jb->tid; r < jb->nrows; r += jb->nthreads) {
        int64_t k = jb->rows ? jb->rows[r] : r;
        jb->fn(jb->n, jb->mass, jb->x, jb->y, jb->vx, jb->vy, jb->dt, k, &jb->ax[r], &jb->ay[r]);
    }
    return NULL;
}

static void accel_rows(row_fn fn, int64_t n, const double *mass, const double *x, const double *y,
                       const double *vx, const double *vy, int32_t dt,
                       const int64_t *rows, int64_t nrows, int nthreads, double *ax, double *ay)
{
    int t;
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    if ((int64_t)nthreads > nrows) nthreads = nrows > 0 ? (int)nrows : 1;
    accel_job_t jobs[256];
    pthread_t th[256];
    for (t = 0; t < nthreads; t++) {
        accel_job_t jb = { fn, n, mass, x, y, vx, vy, dt, rows, nrows, ax, ay, t, nthreads };
        jobs[t] = jb;
    }
    if (nthreads == 1) {
        accel_worker(&jobs[0]);
        return;
    }
    for (t = 0; t < nthreads; t++) pthread_create(&th[t], NULL, accel_worker, &jobs[t]);
    for (t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
}

void oracle_accel(int64_t n, const double *mass, const double *x, const double *y,
                  const double *vx, const double *vy, int32_t dt,
                  const int64_t *rows, int64_t nrows, int nthreads, double *ax, double *ay)
{
    accel_rows(row_accel, n, mass, x, y, vx, vy, dt, rows, nrows, nthreads, ax, ay);
}

void oracle_accel_ld(int64_t n, const double *mass, const double *x, const double *y,
                     const double *vx, const double *vy, int32_t dt,
                     const int64_t *rows, int64_t nrows, int nthreads, double *ax, double *ay)
{
    accel_rows(row_accel_ld, n, mass, x, y, vx, vy, dt, rows, nrows, nthreads, ax, ay);
}

/* ---- full step: sim_gravity.c:1183-1222 + commit :1242-1245 -------------- */

void oracle_step(int64_t n, const double *mass, double *x, double *y,
                 double *vx, double *vy, int32_t dt, int64_t nsteps, int nthreads)
{
    double *ax, *ay, *nx, *ny;
    int64_t s, k;

    if (n <= 0 || nsteps <= 0) return;
    ax = malloc(sizeof(double) * 4 * (size_t)n);
    ay = ax + n;
    nx = ay + n;
    ny = nx + n;

    for (s = 0; s < nsteps; s++) {
        if (nthreads <= 1 || n < 64) {
            for (k = 0; k < n; k++) row_accel(n, mass, x, y, vx, vy, dt, k, &ax[k], &ay[k]);
        } else {
            accel_rows(row_accel, n, mass, x, y, vx, vy, dt, NULL, n, nthreads, ax, ay);
        }
        /* sim_gravity.c:1210-1221: every read below is of time-t state */
        for (k = 0; k < n; k++) {
            const double v1x = vx[k], v1y = vy[k];
            const double ddx = v1x * dt + 0.5 * ax[k] * dt * dt;
            const double ddy = v1y * dt + 0.5 * ay[k] * dt * dt;
            nx[k] = x[k] + ddx;
            ny[k] = y[k] + ddy;
            vx[k] = v1x + ax[k] * dt;   /* velocity is read only by its owner, */
            vy[k] = v1y + ay[k] * dt;   /* so it can be committed immediately  */
        }
        /* sim_gravity.c:1242-1243: positions commit after every body is done */
        memcpy(x, nx, sizeof(double) * (size_t)n);
        memcpy(y, ny, sizeof(double) * (size_t)n);
    }
    free(ax);
}

/* ---- energy diagnostic --------------------------------------------------- */

typedef struct {
    int64_t n;
    const double *mass, *x, *y;
    int tid, nthreads;
    long double pot;
} pot_job_t;

static void *pot_worker(void *arg)
{
    pot_job_t *jb = arg;
    long double acc = 0;
    int64_t i, j;
    for (i = jb->tid; i < jb->n; i += jb->nthreads) {
        long double row = 0;
        for (j = i + 1; j < jb->n; j++) {
            long double dx = (long double)jb->x[j] - jb->x[i];
            long double dy = (long double)jb->y[j] - jb->y[i];
            long double r = sqrtl(dx * dx + dy * dy);
            if (r == 0) continue;
            row += (long double)jb->mass[j] / r;
        }
        acc += row * jb->mass[i];
    }
    jb->pot = acc;
    return NULL;
}

void oracle_energy(int64_t n, const double *mass, const double *x, const double *y,
                   const double *vx, const double *vy, int nthreads,
                   double *kinetic, double *potential)
{
    long double ke = 0, pe = 0;
    pot_job_t jobs[256];
    pthread_t th[256];
    int64_t i;
    int t;

    for (i = 0; i < n; i++) {
        ke += 0.5L * mass[i] * ((long double)vx[i] * vx[i] + (long double)vy[i] * vy[i]);
    }
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    for (t = 0; t < nthreads; t++) {
        pot_job_t jb = { n, mass, x, y, t, nthreads, 0 };
        jobs[t] = jb;
    }
    if (nthreads == 1) {
        pot_worker(&jobs[0]);
    } else {
        for (t = 0; t < nthreads; t++) pthread_create(&th[t], NULL, pot_worker, &jobs[t]);
        for (t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    }
    for (t = 0; t < nthreads; t++) pe += jobs[t].pot;
    *kinetic = (double)ke;
    *potential = (double)(-(long double)ORACLE_G * pe);
}
