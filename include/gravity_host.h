/*
 * include/gravity_host.h -- host-side pieces either side of the hot path,
 * written in C because the reference host is C:
 *
 *   - the scenario-file reader (the input contract of sim_gravity_init,
 *     sim_gravity.c:272-402, tables :138-167), without any SDL dependency;
 *   - the synthetic planar Plummer-model generator used for large-N runs
 *     (no reference equivalent; specified in SURVEY.md section 8d);
 *   - the once-per-day PATH sampling that surrounds the step in the reference
 *     worker (sim_gravity.c:1232-1236, :1248-1251, :1258-1260), expressed as
 *     "how many steps may be batched before the next sample".
 *
 * Error convention as in the reference: 0 on success, -1 on failure with up to
 * three message rows (sim_gravity.c:136), each at most 199 characters.
 */
#ifndef GRAVITY_HOST_H
#define GRAVITY_HOST_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GRAVITY_AU          1.495978707E11     /* sim_gravity.c:30 */
#define GRAVITY_LY          9.4605284E15       /* sim_gravity.c:31 */
#define GRAVITY_MAX_OBJECT  200                /* sim_gravity.c:35 */
#define GRAVITY_MAX_NAME    32                 /* sim_gravity.c:37 */
#define GRAVITY_MAX_PATH_DISP_DAYS (256 * 365.25)   /* sim_gravity.c:40 */

/* SoA image of sim_t + object_t (sim_gravity.c:83-112), minus the display-only
 * PATH ring.  All arrays have n entries and are owned by the struct. */
typedef struct gravity_scenario {
    char     sim_name[GRAVITY_MAX_NAME];   /* upper-cased basename, :411-414 */
    int64_t  sim_time;                     /* seconds, starts at 0 (:415)    */
    int32_t  dt;                           /* DT, snapped (:320-329)         */
    int32_t  n;                            /* sim.max_object                 */
    double   sim_width;                    /* metres, snapped (:299-311)     */
    char   (*name)[GRAVITY_MAX_NAME];
    double  *mass, *radius, *x, *y, *vx, *vy;
    int32_t *disp_color;                   /* index into the reference's colour table */
    int32_t *max_path_disp_dflt;           /* days, clamped (:352-356)       */
    char     error_str[3][200];
} gravity_scenario_t;

/* Reads `pathname` with the semantics of sim_gravity_init.  *out is always set
 * to a struct the caller frees (even on failure, so that error_str can be
 * read) unless memory ran out.  A file without PARAM lines inherits DT and
 * SIM_WIDTH from the previously loaded scenario in the reference
 * (sim_gravity.c:222-223 read the statics before they are reset); pass those
 * values as inherit_dt / inherit_sim_width (0 / 0.0 reproduce a fresh
 * process).  max_objects <= 0 means the reference's cap of 200. */
int gravity_scenario_load_ex(const char *pathname, int32_t inherit_dt, double inherit_sim_width,
                             int32_t max_objects, gravity_scenario_t **out);
int gravity_scenario_load(const char *pathname, gravity_scenario_t **out);
void gravity_scenario_free(gravity_scenario_t *s);

/* Writes the state back out in the reference's file format (absolute
 * coordinates, no indentation), the inverse of the reader; the reference's
 * ENABLE_LOGGING dump (sim_gravity.c:444-467) is the nearest equivalent.
 * The file is written so that the REFERENCE's own parser (sim_gravity.c:272-402)
 * rebuilds the same bits: names are made single tokens (empty -> B<index>, white
 * space -> '_', no leading '#', not the word PARAM; :277-289, :341-349),
 * SIM_WIDTH is written as the snapping table's own literal (:299-311), the DT
 * lines are left out when dt == 0 ("inherit", :222-223), and states the format
 * cannot carry are refused instead of being altered silently:
 *   GRAVITY_SAVE_E_TOO_MANY  n > 200, the reference stops reading there (:398-401)
 *   GRAVITY_SAVE_E_DT        dt is not in valid_dt[] and would be snapped (:320-329)
 *   GRAVITY_SAVE_E_VALUE     a non-finite body field
 * One thing the format itself loses: -0.0 reloads as +0.0 (X + relto[0].X, :380). */
#define GRAVITY_SAVE_E_IO        (-1)
#define GRAVITY_SAVE_E_TOO_MANY  (-2)
#define GRAVITY_SAVE_E_DT        (-3)
#define GRAVITY_SAVE_E_VALUE     (-4)
#define GRAVITY_SAVE_E_ARG       (-5)
int gravity_scenario_save(const gravity_scenario_t *s, const char *pathname);
const char *gravity_scenario_save_error(int code);

/* Planar projection of a Plummer model (Aarseth, Henon & Wielen 1974
 * sampling), SI units: total mass 1.989E30 kg split equally, scale radius
 * 1 AU, G = 6.673E-11; x,y,vx,vy kept, z,vz dropped because the reference is
 * 2-D (sim_gravity_help.h:17-20); centre-of-mass position and velocity
 * removed.  xoshiro256** seeded through splitmix64(seed), draws strictly
 * sequential, so the bytes depend on (n, seed) only. */
int gravity_plummer_generate(int64_t n, uint64_t seed, double *mass, double *x, double *y,
                             double *vx, double *vy);

/* Binary snapshot of a running simulation: checkpoint / resume for any N (the
 * reference has neither; its state lives only in RAM and re-selecting a scenario
 * restarts from sim_time = 0, sim_gravity.c:415).  Layout, little endian:
 * char magic[8] = "GRAVSNP1"; int64 n; int64 sim_time; int32 dt; int32 last_day;
 * then mass, x, y, vx, vy as n doubles each (write() now emits version 2 with no samples,
 * see below; read() takes either).  read() allocates one block that
 * the caller frees through *mass (x, y, vx, vy point into it). */
int gravity_snapshot_write(const char *pathname, int64_t n, int64_t sim_time, int32_t dt, int32_t last_day,
                           const double *mass, const double *x, const double *y,
                           const double *vx, const double *vy);
int gravity_snapshot_read(const char *pathname, int64_t *n, int64_t *sim_time, int32_t *dt, int32_t *last_day,
                          double **mass, double **x, double **y, double **vx, double **vy);

/* The same with the PATH rings (sim_gravity.c:101-104, :1248-1260), so that a resumed run
 * continues them: magic "GRAVSNP2", the v1 header followed by int64 path_tail (samples taken
 * so far) and int64 path_count (samples stored: the most recent ones, oldest first), the five
 * columns, then path_count rows of n {X,Y} pairs.  read_path() accepts both versions (a v1
 * file has no samples); *path_xy is a separate block the caller frees (NULL when empty);
 * the three path pointers may be NULL. */
int gravity_snapshot_write_path(const char *pathname, int64_t n, int64_t sim_time, int32_t dt, int32_t last_day,
                                const double *mass, const double *x, const double *y,
                                const double *vx, const double *vy,
                                int64_t path_tail, int64_t path_count, const double *path_xy);
int gravity_snapshot_read_path(const char *pathname, int64_t *n, int64_t *sim_time, int32_t *dt, int32_t *last_day,
                               double **mass, double **x, double **y, double **vx, double **vy,
                               int64_t *path_tail, int64_t *path_count, double **path_xy);

/* How many steps of size dt may run from sim_time before the reference would
 * take its next PATH sample: the commit at sim_gravity.c:1232-1236 samples the
 * just-committed positions when sim_time/86400 (taken BEFORE sim_time += DT,
 * :1255) differs from the value seen at the previous commit. */
int64_t gravity_steps_until_day_change(int64_t sim_time, int32_t dt, int32_t last_day);

#ifdef __cplusplus
}
#endif
#endif
