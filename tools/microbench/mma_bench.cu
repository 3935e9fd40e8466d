// Microbenchmark: tensor-pipe time of the attention MMAs on B200, without any softmax / handshakes.
//   mode 0: S = Q K^T only   (SS, 128x128x64 = four K=16 MMAs per call)
//   mode 1: O += P V only    (TS, 128x64x128 = eight K=16 MMAs per call, P from TMEM)
//   mode 2: the kernel's order for two Q tiles: QK0 QK1 PV0 PV1 per key block
//   mode 3: as mode 2 but issued by two warps (one per tile), like the kernel
// Reports cycles per "key block of one tile" (QK + PV = 2 x 256 cycles at the pipe floor).
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../musediffusion_b200/csrc/common.cuh"
namespace md { void set_last_error(const char*, ...) {} int check_cuda(cudaError_t, const char*) { return 0; } }
using namespace md;

__global__ void __launch_bounds__(128, 1) k(int mode, int iters, long long* cyc) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;                 // 2 x 16 KB
    uint8_t* sK = smem + 32768;         // 16 KB
    uint8_t* sV = smem + 49152;         // 16 KB
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 65536);   // [0] sink, [1],[2] done
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_init(&bars[2], 1); mbar_init(&bars[3], 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc<512>(&slot);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = slot;
    constexpr uint32_t idesc_qk = make_idesc_bf16(128, 128, 0);
    constexpr uint32_t idesc_pv = make_idesc_bf16(128, 64, 1);
    const uint64_t kd = make_sdesc_sw128(smem_u32(sK)), vd = make_sdesc_sw128(smem_u32(sV));
    long long t0 = 0, t1 = 0;
    if (mode < 3) {
        if (warp == 0) {
            t0 = clock64();
            for (int i = 0; i < iters; ++i) {
                if (mode == 0 || mode == 2) {
                    umma_qk64_commit_w(tb + 0, make_sdesc_sw128(smem_u32(sQ)), kd, idesc_qk, &bars[0]);
                    umma_qk64_commit_w(tb + 128, make_sdesc_sw128(smem_u32(sQ + 16384)), kd, idesc_qk, &bars[0]);
                }
                if (mode == 1 || mode == 2) {
                    umma_pv128_commit_w(tb + 384, tb + 256, vd, idesc_pv, 1, &bars[3]);
                    umma_pv128_commit_w(tb + 448, tb + 320, vd, idesc_pv, 1, &bars[3]);
                }
            }
            tc_commit_w(&bars[1]);
            mbar_wait(&bars[1], 0);
            t1 = clock64();
        }
    } else {
        if (warp < 2) {
            t0 = clock64();
            for (int i = 0; i < iters; ++i) {
                umma_qk64_commit_w(tb + warp * 128, make_sdesc_sw128(smem_u32(sQ + warp * 16384)), kd, idesc_qk, &bars[0]);
                umma_pv128_commit_w(tb + 384 + warp * 64, tb + 256 + warp * 64, vd, idesc_pv, 1, &bars[3]);
            }
            tc_commit_w(&bars[1 + warp]);
            mbar_wait(&bars[1 + warp], 0);
            t1 = clock64();
        }
    }
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tb);
}

int main() {
    long long* cyc; long long h;
    cudaMalloc(&cyc, 8);
    const int smem = 65536 + 1024 + 256;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const char* names[] = {"QK only (SS 128x128x64) x2 tiles", "PV only (TS 128x64x128) x2 tiles", "QK0 QK1 PV0 PV1, one issuing warp", "QK+PV per tile, two issuing warps"};
    for (int mode = 0; mode < 4; ++mode) {
        const int iters = 4000;
        k<<<148, 128, smem>>>(mode, 10, cyc);
        k<<<148, 128, smem>>>(mode, iters, cyc);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("%-40s %8.1f cycles per key block pair (2 tiles)   [%s]\n", names[mode], (double)h / iters, cudaGetErrorString(e));
    }
    return 0;
}
