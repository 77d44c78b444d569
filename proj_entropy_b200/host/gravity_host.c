/*
 * gravity_host.c -- scenario reader, Plummer generator and day-sampling
 * arithmetic of libgravity_host (see include/gravity_host.h).
 *
 * The reader reproduces the behaviour of sim_gravity_init
 * (sim_gravity.c:197-422) without SDL: same comment / blank-line rules, same
 * PARAM handling and snapping tables, same "indent = relative to the previous
 * object one level up" rule, same error rows.  tests/test_scenario.py checks
 * it field by field against what the reference's own parser produced
 * (tests/golden/).
 */
#include "gravity_host.h"

#include <ctype.h>
#include <libgen.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ---- snapping tables, sim_gravity.c:138-167 -------------------------------- */

static const int32_t k_valid_dt[] = {
    1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 20, 30, 40, 50, 60, 120, 180, 240, 300, 360, 420, 480, 540,
    600, 1200, 1800, 2400, 3000, 3600, 2 * 3600, 3 * 3600, 4 * 3600, 5 * 3600, 6 * 3600, 86400,
};
#define N_VALID_DT ((int)(sizeof(k_valid_dt) / sizeof(k_valid_dt[0])))

/* .01, .02, .05 AU ... 5,000,000 AU: 1-2-5 per decade over nine decades */
#define N_VALID_WIDTH 27
static double valid_width(int i)
{
    /* the reference writes literals such as .02*AU and .5*AU; build the same
     * doubles from the same decimal literals */
    static const double lit[N_VALID_WIDTH] = {
        .01, .02, .05, .1, .2, .5, 1, 2, 5, 10, 20, 50, 100, 200, 500, 1000, 2000, 5000,
        10000, 20000, 50000, 100000, 200000, 500000, 1000000, 2000000, 5000000};
    return lit[i] * GRAVITY_AU;
}

static const char *const k_colors[] = {"PURPLE", "BLUE", "LIGHT_BLUE", "GREEN", "YELLOW",
                                       "ORANGE", "RED", "GRAY", "PINK", "WHITE"};

static int32_t color_index(const char *s)
{
    int i;
    for (i = 0; i < 10; i++) {
        if (strcmp(s, k_colors[i]) == 0) return i;
    }
    return 9; /* unknown names fall back to WHITE, util.h COLOR_STR_TO_COLOR */
}

/* ---- file access with the semantics of util.c:1753-1806 -------------------- */

#define FILE_CAP 1000000

static char *slurp(const char *pathname)
{
    FILE *fp = fopen(pathname, "r");
    char *buf;
    size_t len;
    if (fp == NULL) return NULL;
    buf = calloc(1, FILE_CAP);
    if (buf == NULL) {
        fclose(fp);
        return NULL;
    }
    len = fread(buf, 1, FILE_CAP - 1, fp);
    fclose(fp);
    if (len == 0) { /* an empty file fails to open in the reference too */
        free(buf);
        return NULL;
    }
    return buf;
}

static char *next_line(char *buf, size_t *offset)
{
    char *s, *nl;
    if (buf[*offset] == '\0') return NULL;
    s = buf + *offset;
    nl = strchr(s, '\n');
    if (nl != NULL) {
        *nl = '\0';
        *offset = (size_t)(nl - buf) + 1;
    } else {
        *offset = (size_t)(s - buf) + strlen(s);
    }
    return s;
}

/* ---- scenario -------------------------------------------------------------- */

static gravity_scenario_t *scenario_alloc(int32_t cap)
{
    gravity_scenario_t *s = calloc(1, sizeof(*s));
    if (s == NULL) return NULL;
    s->name = calloc((size_t)cap, GRAVITY_MAX_NAME);
    s->mass = calloc((size_t)cap * 6, sizeof(double));
    s->disp_color = calloc((size_t)cap * 2, sizeof(int32_t));
    if (!s->name || !s->mass || !s->disp_color) {
        gravity_scenario_free(s);
        return NULL;
    }
    s->radius = s->mass + cap;
    s->x = s->radius + cap;
    s->y = s->x + cap;
    s->vx = s->y + cap;
    s->vy = s->vx + cap;
    s->max_path_disp_dflt = s->disp_color + cap;
    return s;
}

void gravity_scenario_free(gravity_scenario_t *s)
{
    if (s == NULL) return;
    free(s->name);
    free(s->mass);
    free(s->disp_color);
    free(s);
}

static int parse_error(gravity_scenario_t *s, const char *what, const char *pathname, int line)
{
    snprintf(s->error_str[0], 200, "%s", what);
    snprintf(s->error_str[1], 200, "%s", pathname);
    if (line > 0) snprintf(s->error_str[2], 200, "line %d", line);
    s->n = 0;
    return -1;
}

int gravity_scenario_load_ex(const char *pathname, int32_t inherit_dt, double inherit_sim_width,
                             int32_t max_objects, gravity_scenario_t **out)
{
    enum { MAX_RELTO = 100 };
    struct { double x, y, vx, vy; } relto[MAX_RELTO];
    gravity_scenario_t *s;
    char *buf, *ln, *base, tmp[200];
    size_t offset = 0;
    int line = 0, last_idx = 0, i;
    int32_t new_dt = inherit_dt;
    double new_width = inherit_sim_width;

    if (out == NULL || pathname == NULL) return -1;
    if (max_objects <= 0) max_objects = GRAVITY_MAX_OBJECT;
    *out = s = scenario_alloc(max_objects);
    if (s == NULL) return -1;
    memset(relto, 0, sizeof(relto));

    /* defaults in force while the file is read, sim_gravity.c:252-260 */
    snprintf(s->sim_name, GRAVITY_MAX_NAME, "GRAVITY");
    s->sim_width = 3 * GRAVITY_AU;
    s->dt = 1;

    buf = slurp(pathname);
    if (buf == NULL) return parse_error(s, "Open Failed", pathname, 0);

    while ((ln = next_line(buf, &offset)) != NULL) {
        char name[GRAVITY_MAX_NAME], color[100];
        double x, y, vx, vy, mass, radius;
        int disp, idx, cnt;

        line++;
        if (ln[0] == '#') continue;                       /* :277-279 */
        for (i = 0; ln[i] == ' '; i++) {}
        if (ln[i] == '\0') continue;                      /* :280-285 */

        if (strncmp(ln, "PARAM ", 6) == 0) {              /* :288-338 */
            char pname[100], punits[100];
            double pvalue;
            if (sscanf(ln + 6, "%99s %lf %99s", pname, &pvalue, punits) != 3) goto bad_param;
            if (strcmp(pname, "SIM_WIDTH") == 0 &&
                (strcmp(punits, "AU") == 0 || strcmp(punits, "LY") == 0)) {
                new_width = pvalue * (strcmp(punits, "AU") == 0 ? GRAVITY_AU : GRAVITY_LY);
                for (i = N_VALID_WIDTH - 1; i >= 0; i--) {   /* snap down */
                    if (new_width >= valid_width(i)) {
                        new_width = valid_width(i);
                        break;
                    }
                }
                if (i == -1) new_width = valid_width(0);
            } else if ((strcmp(pname, "DT_LINUX") == 0 || strcmp(pname, "DT_ANDROID") == 0) &&
                       strcmp(punits, "SEC") == 0) {
                if (strcmp(pname, "DT_ANDROID") == 0) continue;   /* other platform's line */
                /* truncation, :320.  Out-of-range and NaN values are undefined in C;
                 * the reference's x86-64 build yields INT32_MIN for them
                 * (cvttsd2si), which then snaps to the table floor: do the same,
                 * explicitly. */
                new_dt = (pvalue > -2147483649.0 && pvalue < 2147483648.0) ? (int32_t)pvalue : INT32_MIN;
                for (i = N_VALID_DT - 1; i >= 0; i--) {
                    if (new_dt >= k_valid_dt[i]) {
                        new_dt = k_valid_dt[i];
                        break;
                    }
                }
                if (i == -1) new_dt = k_valid_dt[0];
            } else {
                goto bad_param;
            }
            continue;
        bad_param:
            free(buf);
            return parse_error(s, "Invalid Parameter", pathname, line);
        }

        cnt = sscanf(ln, "%31s %lf %lf %lf %lf %lf %lf %99s %d", name, &x, &y, &vx, &vy, &mass,
                     &radius, color, &disp);
        if (cnt != 9) {                                   /* :343-349 */
            free(buf);
            return parse_error(s, "Invalid Object", pathname, line);
        }
        if (disp < 0) disp = 0;                           /* :352-356 */
        else if (disp > GRAVITY_MAX_PATH_DISP_DAYS) disp = (int)GRAVITY_MAX_PATH_DISP_DAYS;

        for (idx = 0; ln[idx] == ' '; idx++) {}           /* :359-368 */
        idx++;
        if (idx > last_idx + 1 || idx >= MAX_RELTO) {
            free(buf);
            return parse_error(s, "Invalid Indent", pathname, line);
        }
        last_idx = idx;

        i = s->n;
        memset(s->name[i], 0, GRAVITY_MAX_NAME);
        snprintf(s->name[i], GRAVITY_MAX_NAME, "%s", name);
        s->x[i]  = x + relto[idx - 1].x;                  /* :380-383 */
        s->y[i]  = y + relto[idx - 1].y;
        s->vx[i] = vx + relto[idx - 1].vx;
        s->vy[i] = vy + relto[idx - 1].vy;
        s->mass[i] = mass;
        s->radius[i] = radius;
        s->disp_color[i] = color_index(color);
        s->max_path_disp_dflt[i] = disp;
        relto[idx].x  = s->x[i];                          /* :389-393 */
        relto[idx].y  = s->y[i];
        relto[idx].vx = s->vx[i];
        relto[idx].vy = s->vy[i];
        s->n++;
        if (s->n == max_objects) break;                   /* :398-401 */
    }
    free(buf);

    snprintf(tmp, sizeof(tmp), "%s", pathname);           /* :410-414 */
    base = basename(tmp);
    memset(s->sim_name, 0, GRAVITY_MAX_NAME);
    strncpy(s->sim_name, base, GRAVITY_MAX_NAME - 1);
    for (i = 0; s->sim_name[i]; i++) s->sim_name[i] = (char)toupper((unsigned char)s->sim_name[i]);
    s->sim_time = 0;
    s->sim_width = new_width;
    s->dt = new_dt;
    return 0;
}

int gravity_scenario_load(const char *pathname, gravity_scenario_t **out)
{
    return gravity_scenario_load_ex(pathname, 0, 0.0, 0, out);
}

/* A NAME the reference's `sscanf("%s ...")` (sim_gravity.c:341-349) reads back as
 * exactly one token and that its line classifier (:277-289) does not mistake for
 * a comment or a PARAM line: white space and control characters become '_', an
 * empty name becomes B<index>, a leading '#' and the bare word PARAM get a '_'
 * in front.  An empty NAME would otherwise leave 8 fields behind a leading
 * blank, i.e. "Invalid Object" -- and that blank would also count as one level of
 * indentation (:358-368). */
static void writable_name(const char *src, int index, char out[GRAVITY_MAX_NAME])
{
    char tmp[GRAVITY_MAX_NAME];
    int i, len = 0, printable = 0;
    for (i = 0; i < GRAVITY_MAX_NAME - 1 && src != NULL && src[i] != '\0'; i++) {
        const unsigned char c = (unsigned char)src[i];
        tmp[len++] = (c <= ' ' || c == 0x7f) ? '_' : (char)c;
        if (c > ' ' && c != 0x7f) printable = 1;
    }
    tmp[len] = '\0';
    if (!printable) {
        snprintf(out, GRAVITY_MAX_NAME, "B%d", index);
    } else if (tmp[0] == '#' || strcmp(tmp, "PARAM") == 0) {
        snprintf(out, GRAVITY_MAX_NAME, "_%.*s", GRAVITY_MAX_NAME - 2, tmp);
    } else {
        snprintf(out, GRAVITY_MAX_NAME, "%s", tmp);
    }
}

const char *gravity_scenario_save_error(int code)
{
    switch (code) {
    case 0: return "ok";
    case GRAVITY_SAVE_E_IO: return "cannot write the file";
    case GRAVITY_SAVE_E_TOO_MANY: return "more than 200 bodies: the reference stops reading at MAX_OBJECT";
    case GRAVITY_SAVE_E_DT: return "DT is not in the reference's table of step sizes and would be snapped on load";
    case GRAVITY_SAVE_E_VALUE: return "a body field is not finite";
    default: return "invalid argument";
    }
}

int gravity_scenario_save(const gravity_scenario_t *s, const char *pathname)
{
    FILE *fp;
    int i, wi, ok = 1;
    if (s == NULL || pathname == NULL || s->n < 0) return GRAVITY_SAVE_E_ARG;
    /* the reference silently stops at MAX_OBJECT bodies (:398-401): a longer file
     * would load as a different system */
    if (s->n > GRAVITY_MAX_OBJECT) return GRAVITY_SAVE_E_TOO_MANY;
    /* DT is snapped down to valid_dt[] on load (:320-329); dt == 0 means "no PARAM
     * line" (the loader's inherit value) and is written as no line at all */
    if (s->dt != 0) {
        for (i = 0; i < N_VALID_DT && k_valid_dt[i] != s->dt; i++) {}
        if (i == N_VALID_DT) return GRAVITY_SAVE_E_DT;
    }
    for (i = 0; i < s->n; i++) {
        if (!isfinite(s->x[i]) || !isfinite(s->y[i]) || !isfinite(s->vx[i]) || !isfinite(s->vy[i]) ||
            !isfinite(s->mass[i]) || !isfinite(s->radius[i]))
            return GRAVITY_SAVE_E_VALUE;
    }
    fp = fopen(pathname, "w");
    if (fp == NULL) return GRAVITY_SAVE_E_IO;
    ok &= fprintf(fp, "# %s -- written by libgravity_host at sim_time %lld s\n", s->sim_name,
                  (long long)s->sim_time) > 0;
    /* SIM_WIDTH is snapped DOWN to valid_sim_width[] on load (:299-311).  Writing
     * sim_width/AU and multiplying by AU again can land one ulp below the table
     * entry and fall to the next smaller one, so the table's own decimal literal
     * is written instead: the reference then rebuilds the identical double. */
    if (s->sim_width > 0) {
        static const char *const lit[N_VALID_WIDTH] = {
            ".01", ".02", ".05", ".1", ".2", ".5", "1", "2", "5", "10", "20", "50", "100", "200", "500", "1000",
            "2000", "5000", "10000", "20000", "50000", "100000", "200000", "500000", "1000000", "2000000",
            "5000000"};
        for (wi = N_VALID_WIDTH - 1; wi > 0 && s->sim_width < valid_width(wi); wi--) {}
        ok &= fprintf(fp, "PARAM SIM_WIDTH %s AU\n", lit[wi]) > 0;
    }
    if (s->dt != 0) {
        ok &= fprintf(fp, "PARAM DT_LINUX %d SEC\n", s->dt) > 0;
        ok &= fprintf(fp, "PARAM DT_ANDROID %d SEC\n", s->dt) > 0;
    }
    for (i = 0; i < s->n; i++) {
        char name[GRAVITY_MAX_NAME];
        int c = s->disp_color ? s->disp_color[i] : 9;
        writable_name(s->name ? s->name[i] : NULL, i, name);
        /* absolute coordinates, no indentation: relto[0] is all zero (:228), and
         * %.17g round-trips every finite double through glibc's %lf */
        ok &= fprintf(fp, "%s %.17g %.17g %.17g %.17g %.17g %.17g %s %d\n", name, s->x[i], s->y[i],
                      s->vx[i], s->vy[i], s->mass[i], s->radius[i], k_colors[(c >= 0 && c < 10) ? c : 9],
                      s->max_path_disp_dflt ? s->max_path_disp_dflt[i] : 0) > 0;
    }
    ok &= fclose(fp) == 0;
    return ok ? 0 : GRAVITY_SAVE_E_IO;
}

/* ---- day sampling, sim_gravity.c:1232-1236 ---------------------------------- */

int64_t gravity_steps_until_day_change(int64_t sim_time, int32_t dt, int32_t last_day)
{
    int64_t boundary, gap;
    if ((int32_t)(sim_time / 86400) != last_day) return 1;
    if (dt <= 0) return INT64_MAX;
    boundary = ((int64_t)last_day + 1) * 86400;   /* first sim_time whose day differs */
    gap = boundary - sim_time;                    /* > 0 here */
    /* the commit of step s sees sim_time + (s-1)*dt */
    return (gap + dt - 1) / dt + 1;
}

/* ---- binary snapshots ----------------------------------------------------------- */

static const char k_snap_magic1[8] = {'G', 'R', 'A', 'V', 'S', 'N', 'P', '1'};
static const char k_snap_magic2[8] = {'G', 'R', 'A', 'V', 'S', 'N', 'P', '2'};

int gravity_snapshot_write_path(const char *pathname, int64_t n, int64_t sim_time, int32_t dt, int32_t last_day,
                                const double *mass, const double *x, const double *y,
                                const double *vx, const double *vy,
                                int64_t path_tail, int64_t path_count, const double *path_xy)
{
    const double *cols[5] = {mass, x, y, vx, vy};
    FILE *fp;
    int c, ok = 1;
    if (!pathname || n < 1 || !mass || !x || !y || !vx || !vy) return -1;
    if (path_count < 0 || path_tail < path_count || (path_count > 0 && !path_xy)) return -1;
    fp = fopen(pathname, "wb");
    if (fp == NULL) return -1;
    ok &= fwrite(k_snap_magic2, 1, 8, fp) == 8;
    ok &= fwrite(&n, sizeof(n), 1, fp) == 1;
    ok &= fwrite(&sim_time, sizeof(sim_time), 1, fp) == 1;
    ok &= fwrite(&dt, sizeof(dt), 1, fp) == 1;
    ok &= fwrite(&last_day, sizeof(last_day), 1, fp) == 1;
    ok &= fwrite(&path_tail, sizeof(path_tail), 1, fp) == 1;
    ok &= fwrite(&path_count, sizeof(path_count), 1, fp) == 1;
    for (c = 0; c < 5 && ok; c++) ok &= fwrite(cols[c], sizeof(double), (size_t)n, fp) == (size_t)n;
    if (ok && path_count > 0) {
        const size_t cnt = (size_t)path_count * (size_t)n * 2;
        ok &= fwrite(path_xy, sizeof(double), cnt, fp) == cnt;
    }
    ok &= fclose(fp) == 0;
    return ok ? 0 : -1;
}

int gravity_snapshot_write(const char *pathname, int64_t n, int64_t sim_time, int32_t dt, int32_t last_day,
                           const double *mass, const double *x, const double *y,
                           const double *vx, const double *vy)
{
    return gravity_snapshot_write_path(pathname, n, sim_time, dt, last_day, mass, x, y, vx, vy, 0, 0, NULL);
}

int gravity_snapshot_read_path(const char *pathname, int64_t *n, int64_t *sim_time, int32_t *dt, int32_t *last_day,
                               double **mass, double **x, double **y, double **vx, double **vy,
                               int64_t *path_tail, int64_t *path_count, double **path_xy)
{
    char magic[8];
    int64_t cnt = 0, t = 0, tail = 0, pcount = 0;
    int32_t d = 0, ld = 0;
    double *block, *path = NULL;
    FILE *fp;
    int v2;
    if (!pathname || !n || !sim_time || !dt || !last_day || !mass || !x || !y || !vx || !vy) return -1;
    fp = fopen(pathname, "rb");
    if (fp == NULL) return -1;
    if (fread(magic, 1, 8, fp) != 8) goto bad;
    v2 = memcmp(magic, k_snap_magic2, 8) == 0;
    if (!v2 && memcmp(magic, k_snap_magic1, 8) != 0) goto bad;
    if (fread(&cnt, sizeof(cnt), 1, fp) != 1 || fread(&t, sizeof(t), 1, fp) != 1 ||
        fread(&d, sizeof(d), 1, fp) != 1 || fread(&ld, sizeof(ld), 1, fp) != 1 || cnt < 1 ||
        cnt > ((int64_t)1 << 40))
        goto bad;
    if (v2 && (fread(&tail, sizeof(tail), 1, fp) != 1 || fread(&pcount, sizeof(pcount), 1, fp) != 1 || pcount < 0 ||
               tail < pcount || pcount > ((int64_t)1 << 32) / cnt))
        goto bad;
    block = malloc(sizeof(double) * 5 * (size_t)cnt);
    if (block == NULL || fread(block, sizeof(double), 5 * (size_t)cnt, fp) != 5 * (size_t)cnt) {
        free(block);
        goto bad;
    }
    if (pcount > 0 && path_xy != NULL) {
        const size_t words = (size_t)pcount * (size_t)cnt * 2;
        path = malloc(sizeof(double) * words);
        if (path == NULL || fread(path, sizeof(double), words, fp) != words) {
            free(path);
            free(block);
            goto bad;
        }
    }
    fclose(fp);
    *n = cnt;
    *sim_time = t;
    *dt = d;
    *last_day = ld;
    *mass = block;
    *x = block + cnt;
    *y = block + 2 * cnt;
    *vx = block + 3 * cnt;
    *vy = block + 4 * cnt;
    if (path_tail) *path_tail = tail;
    if (path_count) *path_count = path ? pcount : 0;
    if (path_xy) *path_xy = path;
    return 0;
bad:
    fclose(fp);
    return -1;
}

int gravity_snapshot_read(const char *pathname, int64_t *n, int64_t *sim_time, int32_t *dt, int32_t *last_day,
                          double **mass, double **x, double **y, double **vx, double **vy)
{
    return gravity_snapshot_read_path(pathname, n, sim_time, dt, last_day, mass, x, y, vx, vy, NULL, NULL, NULL);
}

/* ---- Plummer model ----------------------------------------------------------- */

typedef struct { uint64_t s[4]; } rng_t;

static uint64_t splitmix64(uint64_t *state)
{
    uint64_t z = (*state += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

static uint64_t rotl64(uint64_t v, int k) { return (v << k) | (v >> (64 - k)); }

static uint64_t xoshiro256ss(rng_t *r)
{
    const uint64_t result = rotl64(r->s[1] * 5, 7) * 9;
    const uint64_t t = r->s[1] << 17;
    r->s[2] ^= r->s[0];
    r->s[3] ^= r->s[1];
    r->s[1] ^= r->s[2];
    r->s[0] ^= r->s[3];
    r->s[2] ^= t;
    r->s[3] = rotl64(r->s[3], 45);
    return result;
}

static double uniform01(rng_t *r) { return (double)(xoshiro256ss(r) >> 11) * 0x1.0p-53; }

int gravity_plummer_generate(int64_t n, uint64_t seed, double *mass, double *x, double *y,
                             double *vx, double *vy)
{
    const double G = 6.673E-11, M = 1.989E30, a = GRAVITY_AU;
    const double two_pi = 6.283185307179586476925286766559;
    const double vscale = sqrt(G * M / a);
    rng_t rng;
    uint64_t sm = seed;
    long double cx = 0, cy = 0, cvx = 0, cvy = 0;
    int64_t i;

    if (n < 1 || !mass || !x || !y || !vx || !vy) return -1;
    for (i = 0; i < 4; i++) rng.s[i] = splitmix64(&sm);

    for (i = 0; i < n; i++) {
        double r, z, rho, phi, q, v, vz, vrho;
        /* radius from the cumulative mass fraction u1: r = a / sqrt(u1^(-2/3) - 1) */
        for (;;) {
            const double u1 = uniform01(&rng);
            if (u1 <= 0.0) continue;
            r = a / sqrt(pow(u1, -2.0 / 3.0) - 1.0);
            if (r <= 20.0 * a) break;
        }
        /* isotropic direction; only the projection on the x-y plane is kept */
        z   = (1.0 - 2.0 * uniform01(&rng)) * r;
        phi = two_pi * uniform01(&rng);
        rho = sqrt(r * r - z * z);
        x[i] = rho * cos(phi);
        y[i] = rho * sin(phi);
        /* speed as a fraction q of the escape speed, von Neumann rejection on
         * g(q) = q^2 (1-q^2)^(7/2) <= 0.1 */
        for (;;) {
            const double u4 = uniform01(&rng);
            const double u5 = uniform01(&rng);
            if (0.1 * u5 < u4 * u4 * pow(1.0 - u4 * u4, 3.5)) {
                q = u4;
                break;
            }
        }
        v    = q * sqrt(2.0) * pow(1.0 + (r * r) / (a * a), -0.25) * vscale;
        vz   = (1.0 - 2.0 * uniform01(&rng)) * v;
        phi  = two_pi * uniform01(&rng);
        vrho = sqrt(v * v - vz * vz);
        vx[i] = vrho * cos(phi);
        vy[i] = vrho * sin(phi);
        mass[i] = M / (double)n;
        cx += x[i];
        cy += y[i];
        cvx += vx[i];
        cvy += vy[i];
    }
    /* equal masses: the centre of mass is the plain mean */
    {
        const double mx = (double)(cx / n), my = (double)(cy / n);
        const double mvx = (double)(cvx / n), mvy = (double)(cvy / n);
        for (i = 0; i < n; i++) {
            x[i] -= mx;
            y[i] -= my;
            vx[i] -= mvx;
            vy[i] -= mvy;
        }
    }
    return 0;
}
