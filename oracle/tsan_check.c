/* oracle/tsan_check.c -- TEST INFRASTRUCTURE: drives the threaded oracle under
 * ThreadSanitizer (gcc -fsanitize=thread).  The reference's own worker pool has
 * known benign races (non-volatile `state`, `DT`); the
 * oracle port must have none, since its threads own disjoint rows. */
#include "gravity_oracle.h"

#include <stdio.h>
#include <stdlib.h>

int main(void)
{
    enum { N = 400 };
    static double m[N], x[N], y[N], vx[N], vy[N], ax[N], ay[N];
    unsigned s = 12345u;
    int i;
    for (i = 0; i < N; i++) {
        s = s * 1664525u + 1013904223u; x[i] = (double)(s >> 8) * 1e5;
        s = s * 1664525u + 1013904223u; y[i] = (double)(s >> 8) * 1e5;
        s = s * 1664525u + 1013904223u; vx[i] = (double)(s >> 16) - 32768.0;
        s = s * 1664525u + 1013904223u; vy[i] = (double)(s >> 16) - 32768.0;
        m[i] = 1e24 + i * 1e20;
    }
    oracle_accel(N, m, x, y, vx, vy, 7, NULL, N, 6, ax, ay);
    oracle_step(N, m, x, y, vx, vy, 7, 5, 6);
    {
        double ke, pe;
        oracle_energy(N, m, x, y, vx, vy, 6, &ke, &pe);
        printf("ok %a %a %a\n", ax[0], x[1], ke + pe);
    }
    return 0;
}
