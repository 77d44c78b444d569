// tools/fp64_mix.cu -- developer microbenchmark: the instruction MIX of one
// pair interaction (13 FP64 ops + 1 MUFU.RSQ64H) without its dependency chain,
// to find the ceiling the hardware allows for this mix and which ingredient
// costs what.  Prints FP64 warp-instructions per cycle per SM sub-partition.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CH 4

__device__ __forceinline__ double seed(double x) { double r; asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); return r; }

template <int MODE>
__global__ void __launch_bounds__(256) k(double *out, int iters, double a, double b, const double *smem_src)
{
    __shared__ double sh[512];
    sh[threadIdx.x] = smem_src ? smem_src[threadIdx.x] : 1.0;
    sh[threadIdx.x + 256] = 2.0;
    __syncthreads();
    double p[CH], q[CH], s[CH], t[CH], u[CH], ax[CH], ay[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) {
        p[c] = a + (threadIdx.x * 8 + c) * 1e-9; q[c] = b + (threadIdx.x * 8 + c) * 1e-9;
        s[c] = p[c] * 1.5; t[c] = q[c] * 1.25; u[c] = p[c] + q[c]; ax[c] = 0; ay[c] = 0;
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            // 11 two-operand FP64 ops, independent across the statements as far as possible
            p[c] = p[c] * q[c];             // 1
            s[c] = s[c] + t[c];             // 2
            u[c] = fma(u[c], u[c], p[c]);   // 3
            q[c] = q[c] * s[c];             // 4
            t[c] = fma(t[c], u[c], 1.5);    // 5
            p[c] = p[c] + u[c];             // 6
            s[c] = s[c] * t[c];             // 7
            u[c] = fma(u[c], q[c], 1.0);    // 8
            q[c] = q[c] + p[c];             // 9
            t[c] = t[c] * s[c];             // 10
            s[c] = fma(s[c], s[c], q[c]);   // 11
            if (MODE == 0) {                // two more two-operand ops (all-2-cycle reference)
                ax[c] = ax[c] + u[c];
                ay[c] = ay[c] * t[c];
            }
            if (MODE >= 1) {                // the two accumulates: three distinct operands each
                ax[c] = fma(t[c], u[c], ax[c]);
                ay[c] = fma(t[c], p[c], ay[c]);
            }
            if (MODE >= 2) {                // + the seed
                u[c] = u[c] + seed(s[c]);   // NOTE: adds one extra DADD (14 FP64 ops)
            }
            if (MODE >= 3) {                // + one broadcast LDS.128 per 2 interactions and MOVs
                if ((c & 1) == 0) { double2 v = *reinterpret_cast<double2 *>(&sh[(it * 2) & 255]); q[c] += v.x * 1e-30; }
            }
        }
    }
    double r = 0;
#pragma unroll
    for (int c = 0; c < CH; ++c) r += p[c] + q[c] + s[c] + t[c] + u[c] + ax[c] + ay[c];
    if (r == 12345.678) out[0] = r;
}

template <int MODE>
void run(const char *name, int sms, int bps, double ghz, int fp64_per_group)
{
    double *out; cudaMalloc(&out, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        k<MODE><<<sms * bps, 256>>>(out, iters, 0.999999, 1e-9, nullptr);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double groups = (double)sms * bps * 8 * (double)iters * CH;      // warp-level interaction groups
    const double cyc_per_group = best * 1e-3 * ghz * 1e9 * sms * 4 / groups;
    printf("{\"mix\": \"%s\", \"blocks_per_sm\": %d, \"ms\": %.3f, \"cycles_per_group_per_smsp\": %.2f, \"fp64_ops_in_group\": %d}\n",
           name, bps, best, cyc_per_group, fp64_per_group);
    cudaFree(out);
}

int main(int argc, char **argv)
{
    int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const double ghz = argc > 1 ? atof(argv[1]) : 1.965;
    for (int bps = 1; bps <= 2; ++bps) {
        if (bps == 1) { run<0>("13 two-operand", sms, 1, ghz, 13); run<1>("11 two-operand + 2 three-operand fma", sms, 1, ghz, 13); run<2>("... + mufu.rsq64h (+1 dadd)", sms, 1, ghz, 14); run<3>("... + lds", sms, 1, ghz, 14); }
        else          { run<0>("13 two-operand", sms, 2, ghz, 13); run<1>("11 two-operand + 2 three-operand fma", sms, 2, ghz, 13); run<2>("... + mufu.rsq64h (+1 dadd)", sms, 2, ghz, 14); run<3>("... + lds", sms, 2, ghz, 14); }
    }
    return 0;
}
