// Microbenchmark: FFMA vs FFMA2 vs MUFU.TANH issue throughput on one SM-filling grid.
#include <cuda_runtime.h>
#include <cstdio>
template <int MODE>
__global__ void k(float* out, int iters) {
    float a[8]; float2 b[8];
    for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x * 0.001f + i; b[i] = make_float2(a[i], a[i] + 0.5f); }
    const float c = 1.0001f, d = 0.0001f;
    const float2 c2 = make_float2(c, c), d2 = make_float2(d, d);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) a[i] = fmaf(a[i], c, d);
            else if (MODE == 1) b[i] = __ffma2_rn(b[i], c2, d2);
            else if (MODE == 2) { float y; asm volatile("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(a[i])); a[i] = y; }
            else if (MODE == 3) { a[i] = fmaf(a[i], c, d); float y; asm volatile("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(b[i].x)); b[i].x = y; }
        }
    }
    float s = 0;
    for (int i = 0; i < 8; ++i) s += a[i] + b[i].x + b[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char* name, double flop_per_inst) {
    float* out; cudaMalloc(&out, 148 * 8 * 512 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    k<MODE><<<148 * 4, 512>>>(out, 100);
    cudaEventRecord(e0);
    k<MODE><<<148 * 4, 512>>>(out, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double warp_inst = 148.0 * 4 * 16 * iters * 8;
    printf("%-14s %.3f ms  %.1f G warp-inst/s  = %.2f warp-inst/clk/SM @1.9GHz, %.1f TFLOP/s\n", name, ms, warp_inst / ms / 1e6,
           warp_inst / ms / 1e6 / 148 / 1.9, warp_inst * 32 * flop_per_inst / ms / 1e9);
    cudaFree(out);
}
int main() { run<0>("FFMA", 2); run<1>("FFMA2", 4); run<2>("MUFU.TANH", 0); run<3>("FFMA+TANH", 2); return 0; }
