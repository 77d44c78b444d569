/*
 * oracle/gravity_oracle.h -- TEST INFRASTRUCTURE.  Not part of the product.
 *
 * CPU restatement ("Tier B") of the all-pairs gravity step that is inline in
 * the reference worker thread (/root/reference/sim_gravity.c:1183-1255), on a
 * struct-of-arrays state so that N is not capped at MAX_OBJECT = 200
 * (sim_gravity.c:35).  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load this library; the
 * product (libgravity_cuda) never links or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle_pinned.py checks this code
 * bit-for-bit against outputs of the unmodified reference translation unit
 * (oracle/_ref/ref_gravity, built from /root/reference by oracle/Makefile)
 * recorded in tests/golden/, and against that binary live when it is present.
 */
#ifndef GRAVITY_ORACLE_H
#define GRAVITY_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* sim_gravity.c:29 -- the reference's constant, not CODATA. */
#define ORACLE_G 6.673E-11

/* AX, AY exactly as at sim_gravity.c:1207-1208 for the listed bodies.
 * rows == NULL means rows 0..nrows-1.  Work is split over `nthreads` host
 * threads by striding the row list, as sim_gravity.c:1183 strides k. */
void oracle_accel(int64_t n, const double *mass, const double *x, const double *y,
                  const double *vx, const double *vy, int32_t dt,
                  const int64_t *rows, int64_t nrows, int nthreads,
                  double *ax, double *ay);

/* nsteps full steps in place: sim_gravity.c:1183-1222 then the commit at
 * :1242-1245.  The caller advances sim_time (+= dt per step, :1255). */
void oracle_step(int64_t n, const double *mass, double *x, double *y,
                 double *vx, double *vy, int32_t dt, int64_t nsteps, int nthreads);

/* Same accelerations accumulated in x87 long double with correctly rounded
 * long double sqrt/div: a higher-precision yardstick for accuracy reports
 * (the reference author's note at sim_gravity.c:7-15 contemplates it). */
void oracle_accel_ld(int64_t n, const double *mass, const double *x, const double *y,
                     const double *vx, const double *vy, int32_t dt,
                     const int64_t *rows, int64_t nrows, int nthreads,
                     double *ax, double *ay);

/* Energy diagnostic (the reference has none): kinetic = 1/2 sum m |v|^2,
 * potential = - sum_{i<j} G m_i m_j / r_ij, skipping r_ij == 0, evaluated on
 * the un-drifted positions.  Accumulated in long double, returned as double. */
void oracle_energy(int64_t n, const double *mass, const double *x, const double *y,
                   const double *vx, const double *vy, int nthreads,
                   double *kinetic, double *potential);

#ifdef __cplusplus
}
#endif
#endif
