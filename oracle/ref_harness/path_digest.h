/*
 * oracle/ref_harness/path_digest.h -- TEST INFRASTRUCTURE, never shipped.
 *
 * Included by the harnesses AFTER the reference translation unit, so that it can
 * see the reference's static state.  Digest of the PATH rings the reference's
 * commit block fills (sim_gravity.c:1248-1251): samples 0..path_tail-1 in sample
 * order, bodies in index order, X then Y, each double taken as a 64-bit word and
 * folded FNV-1a style (h = (h ^ word) * 1099511628211, h0 = 14695981039346656037).
 * tests/ and gravity_headless compute the same digest over their own samples.
 */
static uint64_t reference_path_digest(void)
{
    uint64_t h = 14695981039346656037ULL;
    int64_t k, tail = path_tail, first = tail > MAX_PATH ? tail - MAX_PATH : 0;
    int i;
    for (k = first; k < tail; k++) {
        for (i = 0; i < sim.max_object; i++) {
            uint64_t w[2];
            memcpy(w, &sim.object[i]->PATH[k % MAX_PATH], sizeof(w));
            h = (h ^ w[0]) * 1099511628211ULL;
            h = (h ^ w[1]) * 1099511628211ULL;
        }
    }
    return h;
}
