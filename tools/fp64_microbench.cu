// tools/fp64_microbench.cu -- developer microbenchmarks of the sm_100a FP64
// pipe: how many warp-level FP64 instructions per cycle per SM sub-partition
// does the hardware sustain for different operand patterns?  Informs the
// instruction budget of the tiled pair loop, k_tiled_steps.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_microbench tools/fp64_microbench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CHAINS 8

template <int MODE>
__global__ void __launch_bounds__(256) k(double *out, int iters, double a, double b)
{
    double acc[CHAINS], x[CHAINS], y[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) {
        acc[c] = threadIdx.x + c;
        x[c]   = a + (threadIdx.x * 8 + c) * 1e-9;
        y[c]   = b + (threadIdx.x * 8 + c) * 1e-9;
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int c = 0; c < CHAINS; ++c) {
                if (MODE == 0) acc[c] = fma(acc[c], a, b);              // 1 distinct + 2 shared (reuse)
                if (MODE == 1) acc[c] = fma(x[c], y[c], acc[c]);        // 3 distinct
                if (MODE == 2) acc[c] = fma(x[c], x[c], acc[c]);        // 2 distinct
                if (MODE == 3) acc[c] = acc[c] * x[c];                  // DMUL 2 distinct
                if (MODE == 4) acc[c] = acc[c] + x[c];                  // DADD 2 distinct
                if (MODE == 5) acc[c] = fma(acc[c], x[c], 1.5);         // 2 distinct + immediate
                if (MODE == 6) acc[c] = acc[c] * acc[c];                // DMUL 1 distinct
                if (MODE == 7) {                                        // pair of FMAs sharing one operand
                    acc[c] = fma(x[c], y[(c + 1) % CHAINS], acc[c]);
                }
                if (MODE == 8) {                                        // MUFU.RSQ64H only
                    double r;
                    asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(acc[c]));
                    acc[c] = r;
                }
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += acc[c];
    if (s == 12345.678) out[0] = s;
}

template <int MODE>
void run(const char *name, int sms, int blocks_per_sm, double clock_ghz)
{
    double *out;
    cudaMalloc(&out, 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    int iters = 20000;
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        k<MODE><<<sms * blocks_per_sm, 256>>>(out, iters, 0.999999, 1e-9);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double warp_instr = (double)sms * blocks_per_sm * 8 * (double)iters * 4 * CHAINS;
    const double per_smsp_per_s = warp_instr / (sms * 4) / (best * 1e-3);
    printf("{\"mode\": \"%s\", \"blocks_per_sm\": %d, \"ms\": %.3f, \"warp_instr_per_cycle_per_smsp_at_%.3fGHz\": %.4f}\n",
           name, blocks_per_sm, best, clock_ghz, per_smsp_per_s / (clock_ghz * 1e9));
    cudaFree(out);
}

int main(int argc, char **argv)
{
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const double ghz = argc > 1 ? atof(argv[1]) : 1.965;
    for (int bps : {1, 2, 4}) {
        if (bps == 1) { run<0>("dfma acc,a,b (2 shared)", sms, 1, ghz); run<1>("dfma 3 distinct", sms, 1, ghz); run<2>("dfma x,x,acc", sms, 1, ghz); run<3>("dmul 2 distinct", sms, 1, ghz); run<4>("dadd 2 distinct", sms, 1, ghz); run<5>("dfma acc,x,imm", sms, 1, ghz); run<6>("dmul acc,acc", sms, 1, ghz); run<7>("dfma x,y',acc", sms, 1, ghz); run<8>("mufu.rsq64h", sms, 1, ghz); }
        if (bps == 2) { run<0>("dfma acc,a,b (2 shared)", sms, 2, ghz); run<1>("dfma 3 distinct", sms, 2, ghz); run<2>("dfma x,x,acc", sms, 2, ghz); run<3>("dmul 2 distinct", sms, 2, ghz); run<4>("dadd 2 distinct", sms, 2, ghz); run<5>("dfma acc,x,imm", sms, 2, ghz); run<6>("dmul acc,acc", sms, 2, ghz); run<7>("dfma x,y',acc", sms, 2, ghz); run<8>("mufu.rsq64h", sms, 2, ghz); }
        if (bps == 4) { run<0>("dfma acc,a,b (2 shared)", sms, 4, ghz); run<1>("dfma 3 distinct", sms, 4, ghz); run<2>("dfma x,x,acc", sms, 4, ghz); run<3>("dmul 2 distinct", sms, 4, ghz); run<4>("dadd 2 distinct", sms, 4, ghz); run<5>("dfma acc,x,imm", sms, 4, ghz); run<6>("dmul acc,acc", sms, 4, ghz); run<7>("dfma x,y',acc", sms, 4, ghz); run<8>("mufu.rsq64h", sms, 4, ghz); }
    }
    return 0;
}
