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
        long long acc = 0;
        int t, d;
        for (t = 0; t < 400000; t += 7919)
            for (d = 1; d <= 86400; d *= 7) acc += gravity_steps_until_day_change(t, d, t / 86400 % 3);
        if (acc <= 0) return 6;
    }
    printf("loaded %d rejected %d\n", ok, bad);
    return 0;
}
