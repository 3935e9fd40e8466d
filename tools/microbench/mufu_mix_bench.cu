// Can ONE warp per SM sub-partition keep the MUFU busy (1 ex2 per 8 cycles) while also issuing the rest of the
// softmax instruction mix (FFMA2 arg, FADD2 sum, F2FP pack) in the MUFU's shadow?  4 warps per CTA, 1 CTA per SM.
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint64_t f2_pack(float lo, float hi) { uint64_t d; asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi)); return d; }
__device__ __forceinline__ void f2_unpack(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t f2_fma(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ uint64_t f2_add(uint64_t a, uint64_t b) { uint64_t d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

template <int MODE>
__global__ void k(float* out, const float* in, int iters, long long* cyc) {
    float s[64];
    for (int i = 0; i < 64; ++i) s[i] = in[(threadIdx.x + i * 37) & 1023];
    uint64_t acc0 = 0, acc1 = 0;
    uint32_t pk = 0;
    const uint64_t sc = f2_pack(1.4426950f, 1.4426950f), ng = f2_pack(-3.f, -3.f);
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < 32; ++c) {
            float a0, a1;
            if (MODE >= 1) { f2_unpack(f2_fma(f2_pack(s[2 * c], s[2 * c + 1]), sc, ng), a0, a1); } else { a0 = s[2 * c]; a1 = s[2 * c + 1]; }
            const float p0 = ex2(a0), p1 = ex2(a1);
            if (MODE >= 2) { if (c & 1) acc1 = f2_add(acc1, f2_pack(p0, p1)); else acc0 = f2_add(acc0, f2_pack(p0, p1)); }
            if (MODE >= 3) { __nv_bfloat162 v = __floats2bfloat162_rn(p0, p1); pk ^= *reinterpret_cast<uint32_t*>(&v); }
            s[2 * c] = p0 * 0.5f - 1.0f; s[2 * c + 1] = p1;   // keep a dependency across iterations
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
    float r = 0; for (int i = 0; i < 64; ++i) r += s[i];
    float q0, q1; f2_unpack(f2_add(acc0, acc1), q0, q1);
    out[blockIdx.x * blockDim.x + threadIdx.x] = r + q0 + q1 + __uint_as_float(pk);
}
template <int MODE> void run(const char* name, int warps) {
    float *out, *in; long long *cyc, h;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&in, 4096); cudaMemset(in, 0, 4096); cudaMalloc(&cyc, 8);
    const int iters = 2000;
    k<MODE><<<148, warps * 32>>>(out, in, 10, cyc);
    k<MODE><<<148, warps * 32>>>(out, in, iters, cyc);
    cudaDeviceSynchronize();
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-44s %2d warps/SM: %.2f cycles per MUFU warp-instruction per SMSP (8.0 = pipe-bound)  err=%s\n", name, warps,
           (double)h / (iters * 64.0 * (warps / 4.0)), cudaGetErrorString(cudaGetLastError()));
}
int main() {
    for (int w : {4, 8}) {
        if (w == 4) { run<0>("ex2 only", 4); run<1>("+ FFMA2 argument", 4); run<2>("+ FADD2 row sum", 4); run<3>("+ F2FP pack (full softmax mix)", 4); }
        else { run<0>("ex2 only", 8); run<3>("full softmax mix", 8); }
    }
    run<3>("full softmax mix", 12);
    run<3>("full softmax mix", 16);
    return 0;
}
