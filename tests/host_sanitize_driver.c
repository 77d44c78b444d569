/* Test driver (built with -fsanitize=address,undefined by
 * tests/test_host_sanitizers.py): loads every file named on the command line
 * with the product's scenario reader, round-trips what loads through save/load
 * and snapshots, and exercises the Plummer generator and the day arithmetic. */
#include "gravity_host.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

int main(int argc, char **argv)
{
    int a, ok = 0, bad = 0;
    char tmp[512];
    for (a = 2; a < argc; a++) {
        gravity_scenario_t *s = NULL;
        int rc = gravity_scenario_load(argv[a], &s);
        if (rc == 0 && s->n > 0) {
            gravity_scenario_t *t = NULL;
            int64_t n, st;
            int32_t dt, ld;
            double *m, *x, *y, *vx, *vy;
            snprintf(tmp, sizeof(tmp), "%s/roundtrip", argv[1]);
            if (gravity_scenario_save(s, tmp) != 0 || gravity_scenario_load_ex(tmp, 1, 1.0, 1000, &t) != 0 || t->n != s->n) {
                fprintf(stderr, "round trip failed for %s\n", argv[a]);
                return 3;
            }
            gravity_scenario_free(t);
            snprintf(tmp, sizeof(tmp), "%s/snap", argv[1]);
            if (gravity_snapshot_write(tmp, s->n, 12345, s->dt, 0, s->mass, s->x, s->y, s->vx, s->vy) != 0 ||
                gravity_snapshot_read(tmp, &n, &st, &dt, &ld, &m, &x, &y, &vx, &vy) != 0 || n != s->n ||
                memcmp(x, s->x, sizeof(double) * (size_t)n) != 0) {
                fprintf(stderr, "snapshot failed for %s\n", argv[a]);
                return 4;
            }
            free(m);
            ok++;
        } else {
            bad++;
        }
        gravity_scenario_free(s);
    }
    {
        enum { N = 3000 };
        double *buf = malloc(sizeof(double) * 5 * N);
        if (gravity_plummer_generate(N, 7, buf, buf + N, buf + 2 * N, buf + 3 * N, buf + 4 * N) != 0) return 5;
        free(buf);
    }
    {
        /* the writer on states a reader never produces: empty and awkward names (31 bytes, blanks,
         * '#', "PARAM"), 200 and 201 bodies, a DT outside the table, a NaN */
        enum { N = 201 };
        static char names[N][GRAVITY_MAX_NAME];
        static double col[6][N];
        static int32_t ints[2][N];
        gravity_scenario_t s, *t = NULL;
        int i;
        memset(&s, 0, sizeof(s));
        for (i = 0; i < N; i++) {
            col[0][i] = 1e24 + i; col[1][i] = 1e6; col[2][i] = i * 1e9; col[3][i] = -i * 1e9; col[4][i] = i; col[5][i] = -i;
            ints[0][i] = i % 12 - 1; ints[1][i] = i;
            if (i % 5 == 1) memset(names[i], 'x', GRAVITY_MAX_NAME - 1);
            if (i % 5 == 2) snprintf(names[i], GRAVITY_MAX_NAME, " a b\tc ");
            if (i % 5 == 3) snprintf(names[i], GRAVITY_MAX_NAME, "#PARAM");
            if (i % 5 == 4) snprintf(names[i], GRAVITY_MAX_NAME, "PARAM");
        }
        snprintf(s.sim_name, sizeof(s.sim_name), "SANITIZE");
        s.dt = 60; s.n = N; s.sim_width = 7.7 * GRAVITY_AU; s.name = names;
        s.mass = col[0]; s.radius = col[1]; s.x = col[2]; s.y = col[3]; s.vx = col[4]; s.vy = col[5];
        s.disp_color = ints[0]; s.max_path_disp_dflt = ints[1];
        snprintf(tmp, sizeof(tmp), "%s/awkward", argv[1]);
        if (gravity_scenario_save(&s, tmp) != GRAVITY_SAVE_E_TOO_MANY) return 7;
        s.n = 200;
        if (gravity_scenario_save(&s, tmp) != 0 || gravity_scenario_load(tmp, &t) != 0 || t->n != 200 ||
            memcmp(t->x, s.x, sizeof(double) * 200) != 0 || strcmp(t->name[0], "B0") != 0 || strcmp(t->name[4], "_PARAM") != 0)
            return 8;
        gravity_scenario_free(t);
        s.dt = 61;
        if (gravity_scenario_save(&s, tmp) != GRAVITY_SAVE_E_DT) return 9;
        s.dt = 0; col[2][17] = 0.0 / 0.0;
        if (gravity_scenario_save(&s, tmp) != GRAVITY_SAVE_E_VALUE || gravity_scenario_save_error(GRAVITY_SAVE_E_VALUE)[0] == 0) return 10;
    }
    {
        long long acc = 0;
        int t, d;
        for (t = 0; t < 400000; t += 7919)
            for (d = 1; d <= 86400; d *= 7) acc += gravity_steps_until_day_change(t, d, t / 86400 % 3);
        if (acc <= 0) return 6;
    }
    printf("loaded %d rejected %d\n", ok, bad);
    return 0;
}
