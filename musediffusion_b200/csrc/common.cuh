// Blackwell (sm_100a) PTX helpers shared by the kernels of the MuseDiffusion sampling path:
// mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (alloc / mma / commit / ld / st), UMMA descriptors.
// Hand-written inline PTX; nothing here depends on CUTLASS or torch.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace md {

#define MD_DEVINL __device__ __forceinline__

// ----------------------------------------------------------------------------------------------
// error plumbing for the C-ABI layer
// ----------------------------------------------------------------------------------------------
void set_last_error(const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);

// ----------------------------------------------------------------------------------------------
// per-device host state.  A process may drive several GPUs (one cudaSetDevice per thread / per call): everything the
// library remembers between calls — SM count, "dynamic shared memory attribute already raised" flags, the schedule
// tables — is indexed by the calling thread's current device, never process-global.
// ----------------------------------------------------------------------------------------------
constexpr int kMaxDevices = 64;
int current_device();      // cudaGetDevice(), clamped to [0, kMaxDevices)
int num_sms();             // multiprocessor count of the current device
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (kernel, device); `done` = that kernel's static flag array
template <typename Kernel>
inline int ensure_dyn_smem(Kernel kern, int bytes, bool (&done)[kMaxDevices], const char* what) {
    const int dev = current_device();
    if (done[dev]) return 0;
    if (check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes), what)) return -2;
    done[dev] = true;
    return 0;
}
// environment tuning switches are read ONCE per process (never per call)
int env_int(const char* name, int dflt);

// ----------------------------------------------------------------------------------------------
// misc
// ----------------------------------------------------------------------------------------------
MD_DEVINL uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

MD_DEVINL uint32_t elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}\n"
        : "=r"(pred));
    return pred;
}

MD_DEVINL void named_bar_sync(uint32_t id, uint32_t nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ----------------------------------------------------------------------------------------------
// mbarrier
// ----------------------------------------------------------------------------------------------
MD_DEVINL void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
MD_DEVINL void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
MD_DEVINL void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

MD_DEVINL void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
MD_DEVINL void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
MD_DEVINL uint32_t mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok;
}
// Bounded wait: a protocol bug must become a trap (an error the host sees), never a hung GPU.
#ifndef MD_MBAR_SPIN_LIMIT
#define MD_MBAR_SPIN_LIMIT (1u << 28)
#endif
MD_DEVINL void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > MD_MBAR_SPIN_LIMIT) {
            printf("musediff_b200: mbarrier wait timed out (block %d thread %d bar %u parity %u)\n", blockIdx.x,
                   threadIdx.x, smem_u32(bar), parity);
            __trap();
        }
    }
}
// The same wait for control warps (TMA producer, MMA issuers) that share a scheduler with compute warps: the suspend-time
// hint keeps the warp asleep in hardware until the phase completes instead of re-polling every ~100 cycles.
MD_DEVINL void mbar_wait_idle(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    for (;;) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred P;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, P;\n\t}\n"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity), "r"(20000u)
            : "memory");
        if (ok) return;
        if (++spins > (MD_MBAR_SPIN_LIMIT >> 4)) {
            printf("musediff_b200: mbarrier wait timed out (block %d thread %d bar %u parity %u)\n", blockIdx.x,
                   threadIdx.x, smem_u32(bar), parity);
            __trap();
        }
    }
}

// ----------------------------------------------------------------------------------------------
// TMA
// ----------------------------------------------------------------------------------------------
MD_DEVINL void tma_prefetch_desc(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(m) : "memory");
}
MD_DEVINL void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
MD_DEVINL void tma_load_3d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
MD_DEVINL void tma_store_2d(const CUtensorMap* m, const void* smem_src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(m),
                 "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
                 : "memory");
}
MD_DEVINL void tma_store_3d(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(m),
                 "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
MD_DEVINL void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
MD_DEVINL void tma_store_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}

// ----------------------------------------------------------------------------------------------
// tcgen05: TMEM allocation, fences, commit, ld / st
// ----------------------------------------------------------------------------------------------
template <uint32_t kCols>
MD_DEVINL void tmem_alloc(uint32_t* smem_result) {  // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
                 "n"(kCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
MD_DEVINL void tmem_dealloc(uint32_t taddr) {  // same warp that allocated
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
MD_DEVINL void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
MD_DEVINL void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
MD_DEVINL void tc_commit(uint64_t* bar) {  // arrives on `bar` when all prior tcgen05.mma of this thread retire
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
MD_DEVINL void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
MD_DEVINL void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem], bf16 x bf16 -> fp32, issued by ONE thread.
MD_DEVINL void umma_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]  (A = bf16 pairs packed in 32-bit TMEM columns, lane = row)
MD_DEVINL void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
        ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}

// tcgen05.ld 32 lanes x 32-bit, N consecutive columns per thread (thread i of the warp <-> lane base+i).
MD_DEVINL void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
MD_DEVINL void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
MD_DEVINL void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
          "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
// first 16 registers of a 32-register group (P packed in place over the S registers it was computed from)
MD_DEVINL void tmem_st16_lo(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
          "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
// d = bf16x2(lo, hi), with d tied to an existing variable so that the result stays in that variable's register
MD_DEVINL void pack_bf16x2_into(uint32_t& d, float lo, float hi) {
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "+r"(d) : "f"(hi), "f"(lo));
}
MD_DEVINL void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
          "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
          "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
          "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}

// ----------------------------------------------------------------------------------------------
// Warp-collective ("_w") forms: executed by ALL 32 lanes of a converged warp with warp-uniform operands; the
// instruction itself is predicated on elect.sync inside the asm block.  Keeping the operands in warp-uniform control
// flow lets ptxas hold descriptors / addresses in uniform registers (UTCHMMA, UTMALDG and UTCBAR take UR operands);
// issuing from inside an `if (lane == 0)` region instead costs an ELECT + R2UR sequence per instruction, which was
// measured to be slower than the 32-cycle N=64 MMAs it feeds.  elect.sync deterministically picks the same lane, so
// the tcgen05.commit of a warp tracks the tcgen05.mma issued by that warp.
// ----------------------------------------------------------------------------------------------
MD_DEVINL void umma_ss_w(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
MD_DEVINL void umma_ts_w(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
        ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
MD_DEVINL void tc_commit_w(uint64_t* bar) {
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}\n"
        ::"r"(smem_u32(bar))
        : "memory");
}
MD_DEVINL void mbar_arrive_expect_tx_w(uint64_t* bar, uint32_t bytes) {
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n\t}\n"
        ::"r"(smem_u32(bar)), "r"(bytes)
        : "memory");
}
MD_DEVINL void tma_load_2d_w(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n\t}\n"
        ::"r"(smem_u32(smem_dst)), "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
MD_DEVINL void tma_load_3d_w(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n\t}\n"
        ::"r"(smem_u32(smem_dst)), "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

// ---- 2-CTA (cta_group::2) forms: one CTA pair = one 256-row MMA; the leader CTA issues, operands are read from both
// CTAs' shared memory (each holds its own 128 A rows and half of the B rows), commits are multicast to both CTAs.
MD_DEVINL uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
MD_DEVINL void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
template <uint32_t kCols>
MD_DEVINL void tmem_alloc_2cta(uint32_t* smem_result) {  // one full warp in EACH CTA of the pair, same warp id
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "n"(kCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
MD_DEVINL void tmem_dealloc_2cta(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
// arrive (count 1) on the barrier at the same shared-memory offset in CTA `rank` of the cluster
MD_DEVINL void mbar_arrive_remote(uint64_t* bar, uint32_t rank) {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}\n"
        ::"r"(smem_u32(bar)), "r"(rank)
        : "memory");
}
MD_DEVINL void mbar_arrive_remote_w(uint64_t* bar, uint32_t rank) {     // warp-collective, one elected lane arrives
    asm volatile(
        "{\n\t.reg .pred q;\n\t.reg .b32 ra;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "@q mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}\n"
        ::"r"(smem_u32(bar)), "r"(rank)
        : "memory");
}
// TMA load whose completion bytes are credited to the LEADER CTA's barrier (peer bit of the address cleared)
MD_DEVINL void tma_load_2d_2cta_w(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n\t}\n"
        ::"r"(smem_u32(smem_dst)), "l"(m), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1)
        : "memory");
}
MD_DEVINL void umma_ss_2cta_w(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "@q tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
MD_DEVINL void tc_commit_2cta_w(uint64_t* bar) {    // arrives on `bar` (same offset) in BOTH CTAs of the pair
    asm volatile(
        "{\n\t.reg .pred q;\n\t.reg .b16 msk;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "mov.b16 msk, 3;\n\t"
        "@q tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], msk;\n\t}\n"
        ::"r"(smem_u32(bar))
        : "memory");
}

// Batched forms: ONE election per group of MMAs + the commit that follows (the per-instruction ELECT / R2UR
// sequence of the single forms costs more issue time than a 32-cycle N=64 MMA takes to execute).
//   S[tmem_d] = Q K^T over 64 dims: four K=16 MMAs (SS), then commit -> bar
MD_DEVINL void umma_qk64_commit_w(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint64_t* bar) {
    asm volatile(
        "{\n\t.reg .pred q, pt, pf;\n\t.reg .b64 da, db;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "setp.eq.b32 pt, 0, 0;\n\t"
        "setp.ne.b32 pf, 0, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, pf;\n\t"
        "add.u64 da, %1, 2;\n\tadd.u64 db, %2, 2;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, pt;\n\t"
        "add.u64 da, %1, 4;\n\tadd.u64 db, %2, 4;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, pt;\n\t"
        "add.u64 da, %1, 6;\n\tadd.u64 db, %2, 6;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, pt;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%4];\n\t}\n"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(smem_u32(bar))
        : "memory");
}
//   O[tmem_d] (+)= P[tmem_a] V over 128 keys: eight K=16 MMAs (TS; P advances 8 TMEM columns, V 2048 B per step),
//   then commit -> bar
MD_DEVINL void umma_pv128_commit_w(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate,
                                   uint64_t* bar) {
    asm volatile(
        "{\n\t.reg .pred q, pt, p0;\n\t.reg .b64 db;\n\t.reg .b32 ta;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "setp.eq.b32 pt, 0, 0;\n\t"
        "setp.ne.b32 p0, %4, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p0;\n\t"
        "add.u32 ta, %1, 8;\n\tadd.u64 db, %2, 128;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [ta], db, %3, pt;\n\t"
        "add.u32 ta, %1, 16;\n\tadd.u64 db, %2, 256;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [ta], db, %3, pt;\n\t"
        "add.u32 ta, %1, 24;\n\tadd.u64 db, %2, 384;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [ta], db, %3, pt;\n\t"
        "add.u32 ta, %1, 32;\n\tadd.u64 db, %2, 512;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [ta], db, %3, pt;\n\t"
        "add.u32 ta, %1, 40;\n\tadd.u64 db, %2, 640;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [ta], db, %3, pt;\n\t"
        "add.u32 ta, %1, 48;\n\tadd.u64 db, %2, 768;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [ta], db, %3, pt;\n\t"
        "add.u32 ta, %1, 56;\n\tadd.u64 db, %2, 896;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [ta], db, %3, pt;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%5];\n\t}\n"
        ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(smem_u32(bar))
        : "memory");
}

//   the same without the trailing commit (a later tcgen05.commit of the same thread covers these MMAs)
MD_DEVINL void umma_pv128_w(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred q, pt, p0;\n\t.reg .b64 db;\n\t.reg .b32 ta;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "setp.eq.b32 pt, 0, 0;\n\t"
        "setp.ne.b32 p0, %4, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p0;\n\t"
        "add.u32 ta, %1, 8;\n\tadd.u64 db, %2, 128;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [ta], db, %3, pt;\n\t"
        "add.u32 ta, %1, 16;\n\tadd.u64 db, %2, 256;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [ta], db, %3, pt;\n\t"
        "add.u32 ta, %1, 24;\n\tadd.u64 db, %2, 384;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [ta], db, %3, pt;\n\t"
        "add.u32 ta, %1, 32;\n\tadd.u64 db, %2, 512;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [ta], db, %3, pt;\n\t"
        "add.u32 ta, %1, 40;\n\tadd.u64 db, %2, 640;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [ta], db, %3, pt;\n\t"
        "add.u32 ta, %1, 48;\n\tadd.u64 db, %2, 768;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [ta], db, %3, pt;\n\t"
        "add.u32 ta, %1, 56;\n\tadd.u64 db, %2, 896;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [ta], db, %3, pt;\n\t"
        "}\n"
        ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}

//   half-width last key block: O[tmem_d] (+)= P[tmem_a] V over 64 keys: four K=16 MMAs, with / without the commit
MD_DEVINL void umma_pv64_commit_w(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate,
                                  uint64_t* bar) {
    asm volatile(
        "{\n\t.reg .pred q, pt, p0;\n\t.reg .b64 db;\n\t.reg .b32 ta;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "setp.eq.b32 pt, 0, 0;\n\t"
        "setp.ne.b32 p0, %4, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p0;\n\t"
        "add.u32 ta, %1, 8;\n\tadd.u64 db, %2, 128;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [ta], db, %3, pt;\n\t"
        "add.u32 ta, %1, 16;\n\tadd.u64 db, %2, 256;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [ta], db, %3, pt;\n\t"
        "add.u32 ta, %1, 24;\n\tadd.u64 db, %2, 384;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [ta], db, %3, pt;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%5];\n\t}\n"
        ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(smem_u32(bar))
        : "memory");
}
MD_DEVINL void umma_pv64_w(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred q, pt, p0;\n\t.reg .b64 db;\n\t.reg .b32 ta;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "setp.eq.b32 pt, 0, 0;\n\t"
        "setp.ne.b32 p0, %4, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p0;\n\t"
        "add.u32 ta, %1, 8;\n\tadd.u64 db, %2, 128;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [ta], db, %3, pt;\n\t"
        "add.u32 ta, %1, 16;\n\tadd.u64 db, %2, 256;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [ta], db, %3, pt;\n\t"
        "add.u32 ta, %1, 24;\n\tadd.u64 db, %2, 384;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [ta], db, %3, pt;\n\t"
        "}\n"
        ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}

// ----------------------------------------------------------------------------------------------
// UMMA descriptors (bit layout: PTX ISA "tcgen05 shared memory descriptor" / "instruction descriptor")
// ----------------------------------------------------------------------------------------------
// Instruction descriptor, kind::f16, bf16 A/B, fp32 accumulate.
//   [4,6) D format (1 = f32)  [7,10) A format (1 = bf16)  [10,13) B format (1 = bf16)
//   [15] A major (0 = K)      [16] B major (0 = K, 1 = MN) [17,23) N >> 3   [24,29) M >> 4
__host__ __device__ constexpr uint32_t make_idesc_bf16(uint32_t M, uint32_t N, uint32_t b_mn_major = 0) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (b_mn_major << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
// Shared-memory matrix descriptor for a 128B-swizzled tile whose rows are 128 bytes (64 bf16) and whose
// 8-row groups are 1024 bytes apart (what a TMA box with CU_TENSOR_MAP_SWIZZLE_128B and a 64-element inner
// dimension produces).  Valid for K-major operands (rows = M/N index) and for the MN-major B operand whose
// single 64-wide N atom is contiguous (rows = K index).
//   [0,14) addr >> 4   [16,30) LBO >> 4   [32,46) SBO >> 4   [46,48) version = 1   [61,64) layout (2 = SWIZZLE_128B)
MD_DEVINL uint64_t make_sdesc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;              // LBO (ignored for swizzled layouts with a single atom in that mode)
    d |= (uint64_t)(1024u >> 4) << 32;   // SBO = 1024 B between 8-row groups
    d |= (uint64_t)1 << 46;              // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;              // SWIZZLE_128B
    return d;
}

// ----------------------------------------------------------------------------------------------
// small math helpers
// ----------------------------------------------------------------------------------------------
MD_DEVINL uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
MD_DEVINL float2 unpack_bf16x2(uint32_t u) {
    __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
    return __bfloat1622float2(v);
}
MD_DEVINL float fast_tanh(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
MD_DEVINL float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
MD_DEVINL float fast_rcp(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// ---- packed fp32x2 arithmetic (Blackwell FFMA2 / FADD2: two lanes per issue slot) ----
MD_DEVINL uint64_t f2_pack(float lo, float hi) {
    uint64_t d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi));
    return d;
}
MD_DEVINL void f2_unpack(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
MD_DEVINL uint64_t f2_fma(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
MD_DEVINL uint64_t f2_mul(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
MD_DEVINL uint64_t f2_add(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
// 2^x for two lanes on the FMA pipe (no MUFU): Cody-Waite split x = n + f, n = rint(x) through the 1.5*2^23 magic
// constant, 2^f by a degree-3 minimax polynomial on [-0.5, 0.5] (max rel. error 7.6e-5, far below bf16 rounding of
// the result), exponent restored by adding n << 23 to the bit pattern.  Inputs must already be clamped to >= -125.
MD_DEVINL uint64_t f2_exp2_poly(uint64_t x) {
    const uint64_t magic = f2_pack(12582912.0f, 12582912.0f);
    const uint64_t nmagic = f2_pack(-12582912.0f, -12582912.0f);
    const uint64_t neg1 = f2_pack(-1.0f, -1.0f);
    const uint64_t r = f2_add(x, magic);                 // low mantissa bits = rint(x)
    const uint64_t n = f2_add(r, nmagic);                // rint(x) as float
    const uint64_t f = f2_fma(n, neg1, x);               // x - n in [-0.5, 0.5]
    uint64_t p = f2_fma(f2_pack(0.05520550534129143f, 0.05520550534129143f), f, f2_pack(0.24261397123336792f, 0.24261397123336792f));
    p = f2_fma(p, f, f2_pack(0.6932547688484192f, 0.6932547688484192f));
    p = f2_fma(p, f, f2_pack(0.9999276995658875f, 0.9999276995658875f));
    float p0, p1, r0, r1;
    f2_unpack(p, p0, p1);
    f2_unpack(r, r0, r1);
    const uint32_t b0 = __float_as_uint(p0) + (__float_as_uint(r0) << 23);
    const uint32_t b1 = __float_as_uint(p1) + (__float_as_uint(r1) << 23);
    return f2_pack(__uint_as_float(b0), __uint_as_float(b1));
}

// erf-GELU (HF ACT2FN["gelu"]): 0.5 x (1 + erf(x / sqrt2)); erf by Abramowitz-Stegun 7.1.26 (|err| < 1.5e-7).
// (A packed-FFMA2 degree-15 polynomial variant was measured slower in the FFN1 epilogue: the MUFU rcp/ex2 of this
// form run beside the FMA pipe.)
MD_DEVINL float gelu_erf(float x) {
    const float z = fabsf(x) * 0.70710678118654752f;
    const float t = fast_rcp(fmaf(0.3275911f, z, 1.0f));
    float p = fmaf(1.061405429f, t, -1.453152027f);
    p = fmaf(p, t, 1.421413741f);
    p = fmaf(p, t, -0.284496736f);
    p = fmaf(p, t, 0.254829592f);
    p *= t;
    const float e = fast_exp2(-z * z * 1.4426950408889634f);
    const float erf_abs = fmaf(-p, e, 1.0f);           // erf(|x|/sqrt2)
    const float half_x = 0.5f * x;
    return fmaf(fabsf(half_x), erf_abs, half_x);        // 0.5x + 0.5|x| erf(|x|/sqrt2)  == 0.5x(1+erf(x/sqrt2))
}

// The same erf-GELU for two values per issue slot, written as x * Phi(x) with Phi(x) = 1 / (1 + 2^(x q(x^2))):
// the logit of the Gaussian CDF is odd, so no |x| / sign handling is needed, and a degree-9 odd polynomial (fitted
// minimax on the absolute GELU error over |x| <= 12, monotone beyond) gives |gelu error| < 3.5e-6 in fp32 — under half a
// bf16 ulp of any output above 1e-3.  12 issue slots per pair (6 packed FMA-pipe + 2 ex2 + 2 rcp + 2 packed) against
// ~30 for two scalar evaluations above; the FFN1 epilogue is issue-bound.  Coefficients carry the -log2(e) factor.
MD_DEVINL uint64_t gelu_erf_x2(uint64_t x) {
    const uint64_t x2 = f2_mul(x, x);
    uint64_t q = f2_fma(f2_pack(-3.2289849514199886e-06f, -3.2289849514199886e-06f), x2, f2_pack(8.823807002045214e-05f, 8.823807002045214e-05f));
    q = f2_fma(q, x2, f2_pack(0.00036027480382472277f, 0.00036027480382472277f));
    q = f2_fma(q, x2, f2_pack(-0.10522668808698654f, -0.10522668808698654f));
    q = f2_fma(q, x2, f2_pack(-2.3020453453063965f, -2.3020453453063965f));
    float g0, g1;
    f2_unpack(f2_mul(q, x), g0, g1);
    const uint64_t d = f2_add(f2_pack(fast_exp2(g0), fast_exp2(g1)), f2_pack(1.0f, 1.0f));
    float d0, d1;
    f2_unpack(d, d0, d1);
    return f2_mul(x, f2_pack(fast_rcp(d0), fast_rcp(d1)));
}

}  // namespace md
