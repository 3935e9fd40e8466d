// Microbenchmark: throughput of ex2.approx.ftz.f32 vs ex2.approx.f16x2 vs fma.rn.f32x2 per SM (B200).
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

template <int MODE>
__global__ void k(float* out, int iters) {
    float a0 = threadIdx.x * 1e-3f, a1 = a0 + 0.1f, a2 = a0 + 0.2f, a3 = a0 + 0.3f, a4 = a0 + .4f, a5 = a0 + .5f, a6 = a0 + .6f, a7 = a0 + .7f;
    uint32_t h0 = threadIdx.x, h1 = h0 + 1, h2 = h0 + 2, h3 = h0 + 3, h4 = h0 + 4, h5 = h0 + 5, h6 = h0 + 6, h7 = h0 + 7;
    uint64_t d0 = threadIdx.x, d1 = d0 + 1, d2 = d0 + 2, d3 = d0 + 3, d4 = d0 + 4, d5 = d0 + 5, d6 = d0 + 6, d7 = d0 + 7;
    for (int i = 0; i < iters; ++i) {
        if (MODE == 0) {
#define E(x) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x));
            E(a0) E(a1) E(a2) E(a3) E(a4) E(a5) E(a6) E(a7)
        } else if (MODE == 1) {
#define H(x) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(x));
            H(h0) H(h1) H(h2) H(h3) H(h4) H(h5) H(h6) H(h7)
        } else if (MODE == 2) {
#define D(x) asm volatile("fma.rn.f32x2 %0, %0, %0, %0;" : "+l"(x));
            D(d0) D(d1) D(d2) D(d3) D(d4) D(d5) D(d6) D(d7)
        } else if (MODE == 3) {
#define C(x, y) asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(x) : "f"(y), "f"(y));
            C(h0, a0) C(h1, a1) C(h2, a2) C(h3, a3) C(h4, a4) C(h5, a5) C(h6, a6) C(h7, a7)
            a0 += __uint_as_float(h0); a1 += __uint_as_float(h1);
        } else if (MODE == 4) {
#define X(x) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(x));
            X(h0) X(h1) X(h2) X(h3) X(h4) X(h5) X(h6) X(h7)
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 + __uint_as_float(h0 ^ h1 ^ h2 ^ h3 ^ h4 ^ h5 ^ h6 ^ h7) +
                                                 (float)(d0 ^ d1 ^ d2 ^ d3 ^ d4 ^ d5 ^ d6 ^ d7);
}

template <int MODE>
void run(const char* name, int lanes_per_op) {
    float* out;
    cudaMalloc(&out, 148 * 1024 * 4);
    const int iters = 20000;
    k<MODE><<<148, 1024>>>(out, 100);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<148, 1024>>>(out, iters);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    int clk;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    double ops = 8.0 * iters * 1024;                 // thread-level instructions per SM
    double cyc = ms * 1e-3 * clk * 1e3;
    printf("%-28s %.3f ms  thread-instr/clk/SM = %.2f  elements/clk/SM = %.2f (at max clock %d kHz) err=%s\n", name, ms, ops / cyc,
           ops * lanes_per_op / cyc, clk, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
}

int main() {
    run<0>("ex2.approx.ftz.f32", 1);
    run<1>("ex2.approx.f16x2", 2);
    run<4>("ex2.approx.ftz.bf16x2", 2);
    run<2>("fma.rn.f32x2", 2);
    run<3>("cvt.rn.f16x2.f32 (+2 fadd)", 2);
    return 0;
}
