// Microbenchmark: tcgen05.ld (TMEM -> registers) throughput per SM on B200, for 4 / 8 / 16 warps per CTA.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../musediffusion_b200/csrc/common.cuh"
namespace md { void set_last_error(const char*, ...) {} int check_cuda(cudaError_t, const char*) { return 0; } }
using namespace md;

__global__ void k(float* out, int iters, long long* cyc) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) tmem_alloc<512>(&slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        uint32_t r0[32], r1[32], r2[32], r3[32];
        const uint32_t c = ((warp >> 2) * 128) & 511;
        tmem_ld32(base + c, r0);
        tmem_ld32(base + c + 32, r1);
        tmem_ld32(base + c + 64, r2);
        tmem_ld32(base + c + 96, r3);
        tc_wait_ld();
#pragma unroll
        for (int j = 0; j < 32; ++j) acc ^= r0[j] ^ r1[j] ^ r2[j] ^ r3[j];
    }
    const long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float(acc);
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(slot);
}

int main() {
    float* out; long long* cyc; long long h;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
    for (int warps : {4, 8, 16}) {
        const int iters = 2000;
        k<<<148, warps * 32>>>(out, 10, cyc);
        k<<<148, warps * 32>>>(out, iters, cyc);
        cudaDeviceSynchronize();
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        const double bytes = (double)warps * iters * 4 * 32 * 32 * 4;   // per SM
        printf("tcgen05.ld 32x32b.x32, %2d warps/CTA: %.1f B/clk/SM  (%lld cycles, err=%s)\n", warps, bytes / h, h, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
