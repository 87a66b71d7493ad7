// Microbenchmark (B200, sm_100a): issue / pipe rate of scalar FP32 instructions against the packed FFMA2 / FADD2 / FMUL2 forms.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/fp32x2 tools/ubench/fp32x2.cu ; run on the GPU box.
// Prints warp-instructions per cycle per SM sub-partition (clock rate taken from cudaDeviceProp at run time) for each loop.
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
#define ILP 8
template <int MODE>
__global__ void __launch_bounds__(256) bench(float * out, float b, float c)
{
    float a[ILP];
    float2 p[ILP / 2];
    for (int k = 0; k < ILP; ++k) a[k] = threadIdx.x * 1e-3f + k;
    for (int k = 0; k < ILP / 2; ++k) p[k] = make_float2(a[2 * k], a[2 * k + 1]);
    const float2 b2 = make_float2(b, b * 1.0001f), c2 = make_float2(c, c * 0.999f);
    unsigned u = threadIdx.x;
#pragma unroll 1
    for (int i = 0; i < ITERS; ++i) {
#pragma unroll
        for (int k = 0; k < ILP; ++k) {
            if (MODE == 0) a[k] = __fmaf_rn(a[k], b, c);
            if (MODE == 1) a[k] = __fadd_rn(a[k], b);
            if (MODE == 2) a[k] = __fmul_rn(a[k], b);
            if (MODE == 3 && k < ILP / 2) { p[k] = __ffma2_rn(p[k], b2, c2); }
            if (MODE == 4 && k < ILP / 2) { p[k] = __fadd2_rn(p[k], b2); }
            if (MODE == 5 && k < ILP / 2) { p[k] = __fmul2_rn(p[k], b2); }
            if (MODE == 6) { /* FFMA2 interleaved with ALU work (PRMT) */
                if (k < ILP / 2) p[k] = __ffma2_rn(p[k], b2, c2);
                else u = __byte_perm(u, __float_as_uint(p[k - ILP / 2].x), 0x2103u + k);
            }
            if (MODE == 7) { /* FFMA interleaved with PRMT */
                if (k < ILP / 2) a[k] = __fmaf_rn(a[k], b, c);
                else u = __byte_perm(u, __float_as_uint(a[k - ILP / 2]), 0x2103u + k);
            }
            if (MODE == 8) a[k] = __fadd_rz(a[k], b);
            if (MODE == 9) a[k] = fminf(fmaxf(__fadd_rn(a[k], b), 0.0f), 1.0f); /* FADD.SAT */
        }
    }
    float s = __uint_as_float(u) * 1e-30f;
    for (int k = 0; k < ILP; ++k) s += a[k];
    for (int k = 0; k < ILP / 2; ++k) s += p[k].x + p[k].y;
    if (s == 123.456f) out[0] = s;
}
template <int MODE>
static void run(const char * name, double instPerIter, float * d, double hz, int sms)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int blocks = sms * 8;
    bench<MODE><<<blocks, 256>>>(d, 1.0001f, 0.5f);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int r = 0; r < 5; ++r) bench<MODE><<<blocks, 256>>>(d, 1.0001f, 0.5f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double warpInst = 5.0 * blocks * 8.0 * ITERS * instPerIter;
    printf("%-28s %8.3f ms  %.3f warp-inst/cycle/SMSP\n", name, ms / 5, warpInst / (ms * 1e-3 * hz * sms * 4));
}
int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double hz = khz * 1e3;
    printf("%s, %d SMs, %.0f MHz (nominal max; the loops below assume the GPU runs at it)\n", p.name, p.multiProcessorCount, hz / 1e6);
    float * d;
    cudaMalloc(&d, 4);
    run<0>("FFMA (3 registers)", ILP, d, hz, p.multiProcessorCount);
    run<1>("FADD", ILP, d, hz, p.multiProcessorCount);
    run<2>("FMUL", ILP, d, hz, p.multiProcessorCount);
    run<8>("FADD.RZ", ILP, d, hz, p.multiProcessorCount);
    run<9>("FADD.SAT", ILP, d, hz, p.multiProcessorCount);
    run<3>("FFMA2 (2 FMAs each)", ILP / 2, d, hz, p.multiProcessorCount);
    run<4>("FADD2", ILP / 2, d, hz, p.multiProcessorCount);
    run<5>("FMUL2", ILP / 2, d, hz, p.multiProcessorCount);
    run<6>("FFMA2 + PRMT 1:1", ILP, d, hz, p.multiProcessorCount);
    run<7>("FFMA + PRMT 1:1", ILP, d, hz, p.multiProcessorCount);
    return 0;
}
