// Microbenchmark: issue rate of scalar FFMA vs packed FFMA2 (and FMUL/FADD, MUFU.RSQ/RCP) on sm_100a.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp32_rate fp32_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
template <int MODE>
__global__ void __launch_bounds__(256) k(float *out, float a, float b, long long *clk) {
    float x[8];
    float2 y[8];
    double z[8];
    for (int i = 0; i < 8; ++i) { x[i] = threadIdx.x * 1e-3f + i; y[i] = make_float2(x[i], x[i] + 1.f); z[i] = x[i]; }
    float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) x[i] = fmaf(x[i], a, b);
            if (MODE == 1) y[i] = __ffma2_rn(y[i], a2, b2);
            if (MODE == 2) x[i] = __fmul_rn(x[i], a);
            if (MODE == 3) x[i] = __fadd_rn(x[i], a);
            if (MODE == 4) x[i] = rsqrtf(x[i]);
            if (MODE == 5) x[i] = __fdividef(a, x[i]);
            if (MODE == 6) x[i] = a / x[i];
            if (MODE == 7) x[i] = sqrtf(x[i]);
            if (MODE == 8) x[i] = __fsqrt_rn(x[i]) ;
            if (MODE == 9) z[i] = fma(z[i], (double)a, (double)b);
            if (MODE == 10) z[i] = (double)a / z[i];
            if (MODE == 11) z[i] = sqrt(z[i]);
        }
    }
    long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < 8; ++i) s += x[i] + y[i].x + y[i].y + (float)z[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}
template <int MODE>
void run(const char *name, int flop_per_op) {
    float *out; long long *clk, h;
    int sms = 148, bps = 4;   // 4 x 256 threads = 32 warps / SM
    cudaMalloc(&out, sms * bps * 256 * 4); cudaMalloc(&clk, 8);
    k<MODE><<<sms * bps, 256>>>(out, 1.0001f, 0.5f, clk);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<sms * bps, 256>>>(out, 1.0001f, 0.5f, clk);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
    double ops = (double)sms * bps * 256 * ITERS * 8;
    double warp_instr_per_clk_smsp = (double)(bps * 8) * ITERS * 8 / 4 / (double)h;
    printf("%-28s %8.3f ms  %7.2f Tops/s (%6.2f TFLOP/s)  clk %lld  warp-instr/clk/SMSP %.3f\n", name, ms, ops / ms / 1e9,
           ops * flop_per_op / ms / 1e9, h, warp_instr_per_clk_smsp);
    cudaFree(out); cudaFree(clk);
}
int main() {
    run<0>("FFMA scalar", 2);
    run<1>("FFMA2 packed (per float2)", 4);
    run<2>("FMUL scalar", 1);
    run<3>("FADD scalar", 1);
    run<4>("rsqrtf (MUFU.RSQ)", 1);
    run<5>("__fdividef (MUFU.RCP+FMUL)", 1);
    run<6>("IEEE a / x", 1);
    run<7>("sqrtf IEEE", 1);
    run<8>("__fsqrt_rn", 1);
    run<9>("DFMA", 2);
    run<10>("IEEE double a / x", 1);
    run<11>("IEEE double sqrt", 1);
    return 0;
}
