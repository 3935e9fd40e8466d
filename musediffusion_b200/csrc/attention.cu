// Fused non-causal, unmasked multi-head attention for the BERT-style encoder of the denoiser (sm_100a).
//
//   ctx[b, l, h, :] = softmax_j( q[b, l, h, :] . k[b, j, h, :] ) v[b, j, h, :]        (q pre-scaled by 1/sqrt(64))
//
// Replaces HF BertSelfAttention's eager softmax(QK^T/sqrt(d))V (transformers modeling_bert.py, called from the
// reference at MuseDiffusion/models/network.py:151 with hidden states only: no attention mask, no head mask), which
// materialises [B, 12, L, L] fp32 scores in HBM.  Here the scores never leave the SM:
//   * S = Q K^T by tcgen05.mma (SS) into TMEM, one 128x128 fp32 tile per Q tile,
//   * each softmax thread owns one row of S (tcgen05.ld 32x32b), keeps the running max / sum in registers, writes
//     P = exp2(..) as packed bf16 into its own TMEM columns,
//   * O += P V by tcgen05.mma with A = P from TMEM (TS form), B = V tile (MN-major, 128B swizzle) from shared memory,
//   * O is rescaled lazily in TMEM only when the row max grew by more than 2^32 (exact after final normalisation).
// One persistent CTA per SM works on TWO streams of S tiles at a time so the tensor pipe computes S for one stream while
// the other stream's softmax runs.  A work item is one of
//   PAIR   two consecutive 128-row Q tiles of one (b, h); both share every K/V stage brought in by TMA;
//   SPLIT  the odd last Q tile of a (b, h) (L = 2096: 17 tiles, the 17th holds 48 rows): its key blocks are divided
//          between the two streams (flash-decoding style), each accumulates its own (m, l, O) and warpgroup 1 merges the
//          two partial results from TMEM — the tile costs about half a PAIR slot instead of a whole one with one
//          warpgroup idle;
//   SINGLE the same when there is only one key block (L <= 128): stream 1 idles.
// Padding is not paid for twice: only the LAST key block is masked (its own code path — the select per element costs a
// quarter of the loop's issue slots), a last block with <= 64 keys runs at half width (N = 64 Q K^T, 4 instead of 8
// P V MMAs, 64 exponentials per row), and softmax warps whose 32 rows all lie beyond L only keep the barriers moving.
//
// Roles (384 threads = 3 warpgroups): warpgroup 0 = {warp 0: TMA producer, warp 1: TMEM owner + MMA issuer of stream 0,
// warp 2: MMA issuer of stream 1, warp 3: idle}, warpgroup 1 = softmax/correction/epilogue for stream 0, warpgroup 2 =
// same for stream 1.  setmaxnreg moves registers from warpgroup 0 to the softmax warpgroups (a full S row lives in
// registers).  At head dim 64 the MUFU / FMA / ALU mix of the softmax, not the tensor pipe, is the binding resource.
#include <stdlib.h>

#include "common.cuh"
#include "musediff_b200.h"

namespace md {
int make_tmap_bf16_3d(CUtensorMap* tm, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1_elems,
                      uint64_t stride2_elems, uint32_t box0, uint32_t box1);

constexpr int ATT_BQ = 128;        // rows per Q tile
constexpr int ATT_BKV = 128;       // keys per K/V stage
constexpr int ATT_DH = 64;
constexpr int ATT_STAGES = 4;      // power of two (ring position -> stage / phase by shifts)
constexpr int kAttThreads = 384, kRegsProducer = 72, kRegsSoftmax = 216;   // 128*72 + 256*216 = 64512 = 384 x 168
constexpr int ATT_TILE_BYTES = 128 * 64 * 2;   // every smem tile is 128 rows x 128 B
constexpr int kAttXchgBytes = 2 * 128 * 4;     // SPLIT items: stream 1 hands its (m, l) per row to stream 0
constexpr int kAttSmem = 1024 + (4 + 2 * ATT_STAGES + 2) * ATT_TILE_BYTES + 512 + kAttXchgBytes;   // Q double-buffered across work items; 2 output staging tiles
static_assert(kAttSmem <= 232448, "attention shared memory exceeds the 227 KB per-CTA limit");
// TMEM columns (all 512 used): S and P have separate homes so that S(j+1) = Q K^T can be issued as soon as the softmax
// warpgroup has READ S(j) into registers — the tensor pipe's latency leaves the softmax critical path.
constexpr uint32_t TM_S0 = 0, TM_S1 = 128, TM_P0 = 256, TM_P1 = 320, TM_O0 = 384, TM_O1 = 448;
constexpr float kLog2e = 1.4426950408889634f;
// Lazy-rescale threshold in log2 units: P and the running sums may grow to 2^32 times their value under an exact running
// max before O is rescaled — far inside the bf16 / fp32 exponent range (2^127), and it makes the TMEM round trip of the
// rescale rare (8, the usual choice for fp16 P, rescaled in ~20 % of the key blocks and cost ~2 %).  MD_ATT_THR overrides.
constexpr float kRescaleThreshold = 32.0f;

struct AttArgs {
    int B, L, NH;
    int n_qtiles;       // ceil(L/128)
    int n_items;        // work items per (b, h): n_qtiles / 2 PAIRs + (n_qtiles odd ? 1 SPLIT / SINGLE : 0)
    int n_kv;           // ceil(L/128)
    int kv_last;        // keys in the last key block: L - (n_kv - 1) * 128, in (0, 128]
    int half_j;         // index of the key block that runs at half width (n_kv - 1 when kv_last <= 64), else -1
    int total_work;     // B * NH * n_items
    int step_it, step_h, step_b;   // gridDim.x work items ahead = (step_b sequences, step_h heads, step_it items) ahead: the
                                   // persistent loops advance (b, h, item) by additions (no divisions by run-time values)
    float rescale_thr;  // lazy-rescale threshold in log2 units (see kRescaleThreshold)
    long long* trace;   // optional [role 10][event 8][step 64] clock64 stamps of CTA 0 (debug / tuning)
};
#define MD_TRACE(role, ev, step)                                                                          \
    do {                                                                                                  \
        if (kTrace && a.trace != nullptr && blockIdx.x == 0 && lane == 0 && (step) < 64)                  \
            a.trace[((role) * 8 + (ev)) * 64 + (step)] = clock64();                                       \
    } while (0)

enum { ITEM_PAIR = 0, ITEM_SPLIT = 1, ITEM_SINGLE = 2 };
struct Work { int b, h, qt0, mode; };
// Persistent work-item cursor: item w = blockIdx.x + k * gridDim.x of the (b, h, item) lattice, advanced incrementally.
struct WorkCursor {
    int w, b, h, it;
    MD_DEVINL void init(const AttArgs& a) {
        w = blockIdx.x;
        it = w % a.n_items;
        const int bh = w / a.n_items;
        h = bh % a.NH;
        b = bh / a.NH;
    }
    MD_DEVINL bool valid(const AttArgs& a) const { return w < a.total_work; }
    MD_DEVINL void next(const AttArgs& a) {
        w += gridDim.x;
        it += a.step_it;
        int carry = 0;
        if (it >= a.n_items) { it -= a.n_items; carry = 1; }
        h += a.step_h + carry;
        b += a.step_b;
        if (h >= a.NH) { h -= a.NH; ++b; }
    }
    MD_DEVINL Work get(const AttArgs& a) const {
        Work r;
        r.b = b; r.h = h;
        if (it < (a.n_qtiles >> 1)) { r.qt0 = 2 * it; r.mode = ITEM_PAIR; }
        else { r.qt0 = a.n_qtiles - 1; r.mode = (a.n_kv >= 2) ? ITEM_SPLIT : ITEM_SINGLE; }
        return r;
    }
};
// key blocks [jb, jb + nblk) that stream x of a work item walks (nblk = 0: the stream idles)
MD_DEVINL void stream_range(const Work& wk, int x, const AttArgs& a, int& jb, int& nblk) {
    if (wk.mode == ITEM_SPLIT) {
        const int n0 = a.n_kv >> 1;
        jb = x ? n0 : 0;
        nblk = x ? a.n_kv - n0 : n0;
    } else {
        jb = 0;
        nblk = (wk.mode == ITEM_PAIR || x == 0) ? a.n_kv : 0;
    }
}
// K/V ring position (0 .. n_kv-1 inside the work item) that holds local block jj of stream x.  PAIR / SINGLE: both streams
// consume the same block from the same stage.  SPLIT: the producer interleaves the two streams' blocks
// (x0 b0, x1 b0, x0 b1, ...; stream 1's extra block when n_kv is odd comes last).
// Branch-free: position = min(jj * mul + add, n_kv - 1) with (mul, add) = (1, 0) PAIR / SINGLE, (2, x) SPLIT.
MD_DEVINL void ring_pos_coef(const Work& wk, int x, int& mul, int& add) {
    mul = (wk.mode == ITEM_SPLIT) ? 2 : 1;
    add = (wk.mode == ITEM_SPLIT) ? x : 0;
}
// the inverse for the producer: which key block goes into ring position q
MD_DEVINL int block_of_pos(const Work& wk, int q, const AttArgs& a) {
    if (wk.mode != ITEM_SPLIT) return q;
    const int n0 = a.n_kv >> 1;
    if (q == 2 * n0) return a.n_kv - 1;
    return (q & 1) ? n0 + (q >> 1) : (q >> 1);
}

// 32-bit shared-address forms (one base register per stream + immediate offsets instead of one generic pointer per barrier)
MD_DEVINL void mbar_arrive_a(uint32_t addr) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(addr) : "memory"); }
template <bool kSleep = false>      // kSleep: suspend-time hint for waits that are expected to be long (work-item boundaries)
MD_DEVINL void mbar_wait_a(uint32_t addr, uint32_t parity) {
    uint32_t spins = 0;
    for (;;) {
        uint32_t ok;
        if (kSleep)
            asm volatile(
                "{\n\t.reg .pred P;\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2, %3;\n\t"
                "selp.u32 %0, 1, 0, P;\n\t}\n"
                : "=r"(ok)
                : "r"(addr), "r"(parity), "r"(20000u)
                : "memory");
        else
            asm volatile(
                "{\n\t.reg .pred P;\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
                "selp.u32 %0, 1, 0, P;\n\t}\n"
                : "=r"(ok)
                : "r"(addr), "r"(parity)
                : "memory");
        if (ok) return;
        if (++spins > MD_MBAR_SPIN_LIMIT) {
            printf("musediff_b200: mbarrier wait timed out (block %d thread %d bar %u parity %u)\n", blockIdx.x, threadIdx.x, addr, parity);
            __trap();
        }
    }
}
// per-stream barrier block (64 B): byte offsets from the stream's base address
constexpr uint32_t SB_S_FULL = 0, SB_P_FULL = 8, SB_O_FULL = 16, SB_O_EMPTY = 24, SB_S_FREE = 32, SB_P_FREE = 40;

MD_DEVINL void mbar_arrive_w(uint64_t* bar) {     // warp-collective, one elected lane arrives
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q mbarrier.arrive.shared::cta.b64 _, [%0];\n\t}\n"
        ::"r"(smem_u32(bar))
        : "memory");
}

struct SoftmaxState {
    float m_used, l_sum;    // running reference max (log2 units) and row sum relative to it
    uint32_t scnt, fcnt;    // phases of s_full / p_free
};

// One key block of one S row.  NGA = 32-column groups of S that exist (4 = full block, 2 = half-width last block);
// kMask = columns >= `valid` are padding (only ever the last key block).  One mbarrier wait per key block: S(j+1)'s commit
// also covers P V(j-1) (tcgen05.commit tracks every earlier MMA of the issuing thread), so waiting for it after the
// exponentials of block j both frees P / O for rewriting and pre-pays the wait at the top of block j+1.
template <int kPoly, int NGA, bool kMask, bool kTrace>
MD_DEVINL void softmax_block(const AttArgs& a, uint32_t tS, uint32_t tP, uint32_t tO, uint32_t sb, int valid, bool first,
                             bool has_next, SoftmaxState& st, int trole, int lane) {
    MD_TRACE(trole, 1, (int)st.scnt);
    uint32_t s[NGA][32];
#pragma unroll
    for (int g = 0; g < NGA; ++g) tmem_ld32(tS + 32 * g, s[g]);
    tc_wait_ld();
    tc_fence_before();
    mbar_arrive_a(sb + SB_S_FREE);     // S is in registers: the next Q K^T may overwrite it
    MD_TRACE(trole, 2, (int)st.scnt);
    if (kMask) {
#pragma unroll
        for (int g = 0; g < NGA; ++g)
#pragma unroll
            for (int c = 0; c < 32; ++c)
                if (g * 32 + c >= valid) s[g][c] = 0xff800000u;   // -inf
    }
    float mx0 = __uint_as_float(s[0][0]), mx1 = __uint_as_float(s[0][1]), mx2 = __uint_as_float(s[0][2]),
          mx3 = __uint_as_float(s[0][3]);
#pragma unroll
    for (int g = 0; g < NGA; ++g)
#pragma unroll
        for (int c = (g == 0 ? 4 : 0); c < 32; c += 4) {
            mx0 = fmaxf(mx0, __uint_as_float(s[g][c]));
            mx1 = fmaxf(mx1, __uint_as_float(s[g][c + 1]));
            mx2 = fmaxf(mx2, __uint_as_float(s[g][c + 2]));
            mx3 = fmaxf(mx3, __uint_as_float(s[g][c + 3]));
        }
    const float mrow = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
    MD_TRACE(trole, 3, (int)st.scnt);
    const float mb = mrow * kLog2e;        // column 0 of every block is a real key, so mrow is finite
    // lazy rescale decision (the running output lives in TMEM and is only touched after the exponentials, once P V of
    // block j-1 has retired — the wait is then off the critical path); branch-free: the first block starts from
    // m_used = -inf, so it always "rescales" an empty sum by 2^-inf = 0
    const bool need = mb > st.m_used + a.rescale_thr;
    const bool any_resc = __any_sync(0xffffffffu, need) && !first;
    const float f_resc = need ? fast_exp2(st.m_used - mb) : 1.0f;
    st.m_used = need ? mb : st.m_used;
    st.l_sum *= f_resc;
    // p = 2^(s log2e - m): packed FFMA2 for the argument, then kPoly of every 8 pairs take the FMA-pipe polynomial and
    // the rest the MUFU (16 ex2/clk/SM is the binding unit at head dim 64); packed FADD2 sums.
    uint64_t acc_a = 0, acc_b = 0;     // two independent packed accumulators (bit pattern of +0.0f, +0.0f)
    const uint64_t l2e2 = f2_pack(kLog2e, kLog2e);
    const uint64_t negm2 = f2_pack(-st.m_used, -st.m_used);
    MD_TRACE(trole, 4, (int)st.scnt);
    uint32_t pk[NGA][16];
#pragma unroll
    for (int g = 0; g < NGA; ++g) {
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            const uint64_t arg = f2_fma(f2_pack(__uint_as_float(s[g][2 * c]), __uint_as_float(s[g][2 * c + 1])), l2e2, negm2);
            float a0, a1, p0, p1;
            f2_unpack(arg, a0, a1);
            uint64_t p2;
            if (kPoly == 9) {                       // timing experiment only: no exponential at all
                p2 = arg; p0 = a0; p1 = a1;
            } else if ((c & 7) < kPoly) {
                p2 = f2_exp2_poly(f2_pack(fmaxf(a0, -125.0f), fmaxf(a1, -125.0f)));
                f2_unpack(p2, p0, p1);
            } else {
                p0 = fast_exp2(a0);
                p1 = fast_exp2(a1);
                p2 = f2_pack(p0, p1);
            }
            if (c & 1) acc_b = f2_add(acc_b, p2); else acc_a = f2_add(acc_a, p2);
            pk[g][c] = pack_bf16x2(p0, p1);
        }
    }
    MD_TRACE(trole, 6, (int)st.scnt);
    if (has_next) {
        mbar_wait_a(sb + SB_S_FULL, (st.scnt + 1) & 1);      // S(j+1) done => P V(j-1) retired: O and P may be touched
    } else if (!first) {
        mbar_wait_a(sb + SB_P_FREE, st.fcnt & 1);
        ++st.fcnt;
    }
    tc_fence_after();
    if (any_resc) {
#pragma unroll
        for (int g = 0; g < ATT_DH / 32; ++g) {
            uint32_t o[32];
            tmem_ld32(tO + 32 * g, o);
            tc_wait_ld();
#pragma unroll
            for (int c = 0; c < 32; ++c) o[c] = __float_as_uint(__uint_as_float(o[c]) * f_resc);
            tmem_st32(tO + 32 * g, o);
        }
    }
#pragma unroll
    for (int g = 0; g < NGA; ++g) tmem_st16(tP + g * 16, pk[g]);
    {
        float q0, q1;
        f2_unpack(f2_add(acc_a, acc_b), q0, q1);
        st.l_sum += q0 + q1;
    }
    tc_wait_st();
    tc_fence_before();
    mbar_arrive_a(sb + SB_P_FULL);
    MD_TRACE(trole, 7, (int)st.scnt);
    ++st.scnt;
}

// The same barrier traffic for a warp whose 32 rows all lie beyond L (last Q tile): nothing is loaded, computed or stored
// (the P V MMA reads stale P for these lanes; their O rows are never written out).  The waits keep the arrivals of this warp
// inside the phase the other warps are in.
MD_DEVINL void idle_block(uint32_t sb, bool first, bool has_next, SoftmaxState& st) {
    mbar_arrive_a(sb + SB_S_FREE);
    if (has_next) {
        mbar_wait_a<true>(sb + SB_S_FULL, (st.scnt + 1) & 1);
    } else if (!first) {
        mbar_wait_a<true>(sb + SB_P_FREE, st.fcnt & 1);
        ++st.fcnt;
    }
    mbar_arrive_a(sb + SB_P_FULL);
    ++st.scnt;
}

// O row (64 fp32 in TMEM columns) * scale -> bf16 -> 128 B of the staging tile (128B swizzle: 16 B chunk ^= row % 8)
MD_DEVINL void stage_row(uint8_t* tile, int r, const uint32_t (&o)[2][32], float scale) {
    const uint32_t row_addr = smem_u32(tile) + r * 128;
#pragma unroll
    for (int gg = 0; gg < 2; ++gg)
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const int chunk = gg * 4 + g;
            const uint32_t addr = row_addr + ((chunk ^ (r & 7)) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr),
                         "r"(pack_bf16x2(__uint_as_float(o[gg][g * 8 + 0]) * scale, __uint_as_float(o[gg][g * 8 + 1]) * scale)),
                         "r"(pack_bf16x2(__uint_as_float(o[gg][g * 8 + 2]) * scale, __uint_as_float(o[gg][g * 8 + 3]) * scale)),
                         "r"(pack_bf16x2(__uint_as_float(o[gg][g * 8 + 4]) * scale, __uint_as_float(o[gg][g * 8 + 5]) * scale)),
                         "r"(pack_bf16x2(__uint_as_float(o[gg][g * 8 + 6]) * scale, __uint_as_float(o[gg][g * 8 + 7]) * scale))
                         : "memory");
        }
}

template <int kPoly, bool kTrace = false>
__global__ void __launch_bounds__(kAttThreads, 1)
attention_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmOut, const AttArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;                                     // 2 buffers x 2 tiles: the next work item's Q is prefetched
    uint8_t* sK = smem + 4 * ATT_TILE_BYTES;                // ATT_STAGES tiles
    uint8_t* sV = sK + ATT_STAGES * ATT_TILE_BYTES;         // ATT_STAGES tiles
    uint8_t* sO = sV + ATT_STAGES * ATT_TILE_BYTES;         // 2 tiles: normalised output staged for the TMA store
    uint64_t* bars = reinterpret_cast<uint64_t*>(sO + 2 * ATT_TILE_BYTES);
    uint64_t* q_full = bars;                // [2]
    uint64_t* q_empty = bars + 2;           // [2]
    uint64_t* k_full = bars + 4;            // [STAGES]
    uint64_t* k_empty = k_full + ATT_STAGES;
    uint64_t* v_full = k_empty + ATT_STAGES;
    uint64_t* v_empty = v_full + ATT_STAGES;
    // per-stream barriers, one 64 B block per stream (SB_* offsets): s_full, p_full, o_full, o_empty,
    // s_free (softmax has read S into registers -> next Q K^T may overwrite it), p_free (last key block only: P V of
    // block n-2 retired; earlier blocks learn it from s_full(j+1))
    uint8_t* sbar = reinterpret_cast<uint8_t*>(bars) + 192;
    auto SB = [sbar](int x, uint32_t off) { return reinterpret_cast<uint64_t*>(sbar + x * 64 + off); };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sbar + 128);
    float* sXchg = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 512);     // [2][128]: (m, l) of stream 1 (SPLIT items)

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int H = a.NH * ATT_DH;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQKV);
        tma_prefetch_desc(&tmOut);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&q_full[i], 1);
            mbar_init(&q_empty[i], 2);         // both MMA warps (one per stream) release the shared stages
        }
        for (int s = 0; s < ATT_STAGES; ++s) {
            mbar_init(&k_full[s], 1);
            mbar_init(&k_empty[s], 2);         // PAIR: both MMA warps; SPLIT: the owning MMA warp + the producer itself
            mbar_init(&v_full[s], 1);
            mbar_init(&v_empty[s], 2);
        }
        for (int x = 0; x < 2; ++x) {
            mbar_init(SB(x, SB_S_FULL), 1);
            mbar_init(SB(x, SB_P_FULL), 128);
            mbar_init(SB(x, SB_O_FULL), 1);
            mbar_init(SB(x, SB_O_EMPTY), 128);
            mbar_init(SB(x, SB_S_FREE), 128);
            mbar_init(SB(x, SB_P_FREE), 1);
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    // (each role re-reads the TMEM base from shared memory below: a value kept live across the setmaxnreg boundaries is
    // placed in local memory by ptxas and re-loaded from there — L2 latency, the L1 is all shared memory here — at every use)
#define MD_TMEM_BASE() ({ uint32_t v_; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v_) : "r"(smem_u32(tmem_slot))); v_; })

    if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsProducer));
    if (warp == 0) {
        // =========================================================== TMA producer (whole warp, elected issue)
        uint32_t wcnt = 0;
        uint32_t rp = 0;                       // K/V ring position, carried across work items
        WorkCursor cur;
        for (cur.init(a); cur.valid(a); cur.next(a), ++wcnt) {
            const Work wk = cur.get(a);
            const int n_q = (wk.mode == ITEM_PAIR) ? 2 : 1;        // SPLIT: both streams read the same Q tile
            const int qb = wcnt & 1;               // Q buffer of this work item
            mbar_wait_idle(&q_empty[qb], ((wcnt >> 1) & 1) ^ 1);
            mbar_arrive_expect_tx_w(&q_full[qb], n_q * ATT_TILE_BYTES);
            for (int x = 0; x < n_q; ++x)
                tma_load_3d_w(sQ + (qb * 2 + x) * ATT_TILE_BYTES, &tmQKV, &q_full[qb], wk.h * ATT_DH, (wk.qt0 + x) * ATT_BQ, wk.b);
            for (int q = 0; q < a.n_kv; ++q, ++rp) {
                const int j = block_of_pos(wk, q, a);
                const uint32_t st = rp & (ATT_STAGES - 1), ph = (rp / ATT_STAGES) & 1;
                mbar_wait_idle(&k_empty[st], ph ^ 1);
                mbar_arrive_expect_tx_w(&k_full[st], ATT_TILE_BYTES);
                tma_load_3d_w(sK + st * ATT_TILE_BYTES, &tmQKV, &k_full[st], H + wk.h * ATT_DH, j * ATT_BKV, wk.b);
                if (wk.mode == ITEM_SPLIT) mbar_arrive_w(&k_empty[st]);      // stands in for the stream that skips this stage
                mbar_wait_idle(&v_empty[st], ph ^ 1);
                mbar_arrive_expect_tx_w(&v_full[st], ATT_TILE_BYTES);
                tma_load_3d_w(sV + st * ATT_TILE_BYTES, &tmQKV, &v_full[st], 2 * H + wk.h * ATT_DH, j * ATT_BKV, wk.b);
                if (wk.mode == ITEM_SPLIT) mbar_arrive_w(&v_empty[st]);
            }
        }
    } else if (warp == 1 || warp == 2) {
        // =========================================================== MMA issuers: warp 1 -> stream 0, warp 2 -> stream 1
        // (whole warp runs the loop, one elected lane issues).  One issuing warp per stream: the issue stream of a
        // single warp (waits, descriptor moves, 24 small MMAs per key block) was measured to be the critical path.
        const int x = warp - 1;
        const uint32_t tmem_base = MD_TMEM_BASE();
        constexpr uint32_t idesc_qk = make_idesc_bf16(ATT_BQ, ATT_BKV, 0);
        constexpr uint32_t idesc_qk_half = make_idesc_bf16(ATT_BQ, ATT_BKV / 2, 0);
        constexpr uint32_t idesc_pv = make_idesc_bf16(ATT_BQ, ATT_DH, 1);   // B (= V) is MN-major
        const uint32_t tS = tmem_base + (x ? TM_S1 : TM_S0);
        const uint32_t tP = tmem_base + (x ? TM_P1 : TM_P0);
        const uint32_t tO = tmem_base + (x ? TM_O1 : TM_O0);
        uint32_t wcnt = 0;
        uint32_t rp0 = 0;    // ring position of the work item's first key block, carried across work items
        uint32_t pcnt = 0;   // P tiles consumed (phase of p_full)
        uint32_t ocnt = 0;   // work items with output (phase of o_empty)
        uint32_t qcnt = 0;   // Q K^T issued (phase of s_free)
        uint64_t* const s_full = SB(x, SB_S_FULL);
        uint64_t* const p_full = SB(x, SB_P_FULL);
        uint64_t* const o_full = SB(x, SB_O_FULL);
        uint64_t* const o_empty = SB(x, SB_O_EMPTY);
        uint64_t* const s_free = SB(x, SB_S_FREE);
        uint64_t* const p_free = SB(x, SB_P_FREE);
        WorkCursor cur;
        for (cur.init(a); cur.valid(a); cur.next(a), ++wcnt, rp0 += a.n_kv) {
            const Work wk = cur.get(a);
            int jb, nblk;
            stream_range(wk, x, a, jb, nblk);
            const int qb = wcnt & 1;
            const uint64_t qd = make_sdesc_sw128(smem_u32(sQ + (qb * 2 + (wk.mode == ITEM_PAIR ? x : 0)) * ATT_TILE_BYTES));
            mbar_wait_idle(&q_full[qb], (wcnt >> 1) & 1);
            if (nblk == 0) {
                // SINGLE item, stream 1: only keep the shared stages moving
                for (int q = 0; q < a.n_kv; ++q) {
                    const uint32_t rp = rp0 + q, st = rp & (ATT_STAGES - 1), ph = (rp / ATT_STAGES) & 1;
                    mbar_wait_idle(&k_full[st], ph);
                    tc_commit_w(&k_empty[st]);
                    mbar_wait_idle(&v_full[st], ph);
                    tc_commit_w(&v_empty[st]);
                }
                tc_commit_w(&q_empty[qb]);
                continue;
            }
            int pmul, padd;
            ring_pos_coef(wk, x, pmul, padd);
            const int last_pos = a.n_kv - 1;
            {   // S(0)
                const uint32_t rp = rp0 + min(padd, last_pos), st = rp & (ATT_STAGES - 1), ph = (rp / ATT_STAGES) & 1;
                mbar_wait_idle(&k_full[st], ph);
                if (qcnt > 0) mbar_wait_idle(s_free, (qcnt - 1) & 1);
                ++qcnt;
                tc_fence_after();
                umma_qk64_commit_w(tS, qd, make_sdesc_sw128(smem_u32(sK + st * ATT_TILE_BYTES)),
                                   jb == a.half_j ? idesc_qk_half : idesc_qk, s_full);
                tc_commit_w(&k_empty[st]);
            }
            for (int jj = 0; jj < nblk; ++jj) {
                if (jj + 1 < nblk) {
                    // S(j+1): needs only K[j+1] and the softmax warpgroup's READ of S(j)
                    const uint32_t rp = rp0 + min((jj + 1) * pmul + padd, last_pos), st = rp & (ATT_STAGES - 1), ph = (rp / ATT_STAGES) & 1;
                    mbar_wait_idle(&k_full[st], ph);
                    mbar_wait_idle(s_free, (qcnt - 1) & 1);
                    ++qcnt;
                    tc_fence_after();
                    MD_TRACE(x, 0, (int)(wcnt * a.n_kv + jj));
                    umma_qk64_commit_w(tS, qd, make_sdesc_sw128(smem_u32(sK + st * ATT_TILE_BYTES)),
                                       jb + jj + 1 == a.half_j ? idesc_qk_half : idesc_qk, s_full);
                    tc_commit_w(&k_empty[st]);
                } else if (jj > 0) {
                    tc_commit_w(p_free);     // tracks P V(n-2) for the last block's softmax
                }
                const uint32_t rp = rp0 + min(jj * pmul + padd, last_pos), st = rp & (ATT_STAGES - 1), ph = (rp / ATT_STAGES) & 1;
                mbar_wait_idle(&v_full[st], ph);
                if (jj == 0) mbar_wait_idle(o_empty, (ocnt & 1) ^ 1);   // previous item's O drained
                MD_TRACE(x, 1, (int)(wcnt * a.n_kv + jj));
                mbar_wait_idle(p_full, pcnt & 1);
                ++pcnt;
                tc_fence_after();
                MD_TRACE(x, 2, (int)(wcnt * a.n_kv + jj));
                const uint64_t vd = make_sdesc_sw128(smem_u32(sV + st * ATT_TILE_BYTES));
                const uint32_t accum = jj > 0 ? 1u : 0u;
                if (jb + jj == a.half_j) umma_pv64_w(tO, tP, vd, idesc_pv, accum);     // 64 keys = 4 K-steps
                else umma_pv128_w(tO, tP, vd, idesc_pv, accum);                          // covered by the commit behind Q K^T(j+2) ...
                if (jj + 1 == nblk) {                                                    // ... or, for the last block, by this one
                    tc_commit_w(o_full);
                    ++ocnt;
                }
                tc_commit_w(&v_empty[st]);
            }
            tc_commit_w(&q_empty[qb]);
        }
    }
    } else {
        // =========================================================== softmax / correction / epilogue
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegsSoftmax));
        const uint32_t tmem_base = MD_TMEM_BASE();
        const int x = (warp - 4) >> 2;          // which stream this warp works on
        const int quad = warp & 3;              // TMEM lane quadrant
        const int r = quad * 32 + lane;         // row inside the Q tile
        const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
        const uint32_t tS = tmem_base + (x ? TM_S1 : TM_S0) + lane_addr;
        const uint32_t tP = tmem_base + (x ? TM_P1 : TM_P0) + lane_addr;
        const uint32_t tO = tmem_base + (x ? TM_O1 : TM_O0) + lane_addr;
        const int trole = 2 + x * 4 + quad;
        const uint32_t sb = smem_u32(sbar) + x * 64;      // this stream's barrier block
        SoftmaxState st;
        st.scnt = 0; st.fcnt = 0;
        uint32_t ocnt = 0;
        WorkCursor cur;
        for (cur.init(a); cur.valid(a); cur.next(a)) {
            const Work wk = cur.get(a);
            int jb, nblk;
            stream_range(wk, x, a, jb, nblk);
            if (nblk == 0) continue;                              // SINGLE item: stream 1 has nothing to do
            const int qt = wk.qt0 + (wk.mode == ITEM_PAIR ? x : 0);
            const bool rows_exist = (qt * ATT_BQ + quad * 32) < a.L;      // warp-uniform
            st.m_used = -INFINITY; st.l_sum = 0.f;
            mbar_wait_a<true>(sb + SB_S_FULL, st.scnt & 1);      // long at a work-item boundary: sleep
            tc_fence_after();
            for (int jj = 0; jj < nblk; ++jj) {
                const int j = jb + jj;
                const bool has_next = (jj + 1 < nblk), first = (jj == 0);
                MD_TRACE(trole, 0, (int)st.scnt);
                if (!rows_exist) {
                    idle_block(sb, first, has_next, st);
                } else if (j == a.n_kv - 1 && a.kv_last < ATT_BKV) {
                    if (a.kv_last <= ATT_BKV / 2)
                        softmax_block<kPoly, 2, true, kTrace>(a, tS, tP, tO, sb, a.kv_last, first, has_next, st, trole, lane);
                    else
                        softmax_block<kPoly, 4, true, kTrace>(a, tS, tP, tO, sb, a.kv_last, first, has_next, st, trole, lane);
                } else {
                    softmax_block<kPoly, 4, false, kTrace>(a, tS, tP, tO, sb, ATT_BKV, first, has_next, st, trole, lane);
                }
            }
            // ---- epilogue: O / l -> bf16 -> global
            mbar_wait_a<true>(sb + SB_O_FULL, ocnt & 1);
            ++ocnt;
            tc_fence_after();
            uint32_t o[2][32];
            float scale;
            if (wk.mode == ITEM_SPLIT) {
                // Stream 1 hands (m, l) to stream 0, which also reads stream 1's O straight from TMEM (both warpgroups address
                // the same lane quadrants).  Both warpgroups pass through the SAME two bar.sync instructions (compute-sanitizer's
                // synccheck rejects one named barrier reached at two program counters), and both execute the merge arithmetic:
                // a register array defined under a predicate would be live around the whole work-item loop for ptxas
                // (64 registers spilled per key block); stream 1 discards its copy.
                if (x == 1) {
                    sXchg[r] = st.m_used;
                    sXchg[128 + r] = st.l_sum;
                }
                tc_fence_before();
                asm volatile("bar.sync 2, 256;" ::: "memory");       // (m, l) visible, O1 complete
                tc_fence_after();
                const float m1 = sXchg[r], l1 = sXchg[128 + r];
                const float m = fmaxf(st.m_used, m1);
                const float w0 = fast_exp2(st.m_used - m), w1 = fast_exp2(m1 - m);
                scale = 1.0f / (st.l_sum * w0 + l1 * w1);
                uint32_t o1[2][32];
                tmem_ld32(tO, o[0]);
                tmem_ld32(tO + 32, o[1]);
                tmem_ld32(tmem_base + TM_O1 + lane_addr, o1[0]);
                tmem_ld32(tmem_base + TM_O1 + lane_addr + 32, o1[1]);
                tc_wait_ld();
#pragma unroll
                for (int g = 0; g < 2; ++g)
#pragma unroll
                    for (int c = 0; c < 32; ++c)
                        o[g][c] = __float_as_uint(fmaf(__uint_as_float(o[g][c]), w0, __uint_as_float(o1[g][c]) * w1));
                tc_fence_before();
                asm volatile("bar.sync 3, 256;" ::: "memory");       // stream 0 has read O1
                mbar_arrive_a(sb + SB_O_EMPTY);
                if (x == 1) continue;
            } else {
                scale = 1.0f / st.l_sum;
                tmem_ld32(tO, o[0]);
                tmem_ld32(tO + 32, o[1]);
                tc_wait_ld();
                tc_fence_before();
                mbar_arrive_a(sb + SB_O_EMPTY);
            }
            // Row r -> 128 B of the staging tile, then ONE bulk tensor store per tile.  (A direct st.global from the
            // row-per-thread layout touches 32 different lines per warp instruction and was measured to stall the other
            // stream's TMEM traffic for ~3000 cycles per work item.)  Rows >= L of the last tile are clipped by the tensor map.
            const bool issuer = (r == 0);
            if (issuer) tma_store_wait_read<0>();          // the previous item's store has finished reading the tile
            if (x == 0) asm volatile("bar.sync 6, 128;" ::: "memory");
            else asm volatile("bar.sync 7, 128;" ::: "memory");
            stage_row(sO + x * ATT_TILE_BYTES, r, o, scale);      // rows >= L hold garbage: the tensor map clips them
            fence_proxy_async_smem();
            if (x == 0) asm volatile("bar.sync 6, 128;" ::: "memory");
            else asm volatile("bar.sync 7, 128;" ::: "memory");
            if (issuer) {
                tma_store_3d(&tmOut, sO + x * ATT_TILE_BYTES, wk.h * ATT_DH, qt * ATT_BQ, wk.b);
                tma_store_commit();
            }
        }
    }
    if (warp >= 4 && (threadIdx.x & 127) == 0) tma_store_wait_read<0>();   // staging tiles must outlive the bulk stores' reads
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<512>(MD_TMEM_BASE());
#undef MD_TMEM_BASE
}

}  // namespace md

using namespace md;

extern "C" __attribute__((visibility("default"))) int md_attention_bf16(const void* qkv, void* out, int B, int L, int NH, int DH, cudaStream_t stream) {
    if (DH != ATT_DH) { set_last_error("md_attention_bf16: head dim %d unsupported (kernel is specialised for 64)", DH); return MD_ERR_ARG; }
    if (B <= 0 || L <= 0 || NH <= 0) { set_last_error("md_attention_bf16: empty problem B=%d L=%d NH=%d", B, L, NH); return MD_ERR_ARG; }
    const int H = NH * DH;
    CUtensorMap tm;
    if (int e = make_tmap_bf16_3d(&tm, qkv, 3 * H, L, B, 3 * H, (uint64_t)L * 3 * H, 64, 128)) return e;
    AttArgs a;
    a.B = B; a.L = L; a.NH = NH;
    a.n_qtiles = (L + ATT_BQ - 1) / ATT_BQ;
    a.n_items = (a.n_qtiles >> 1) + (a.n_qtiles & 1);
    a.n_kv = (L + ATT_BKV - 1) / ATT_BKV;
    a.kv_last = L - (a.n_kv - 1) * ATT_BKV;
    a.half_j = a.kv_last <= ATT_BKV / 2 ? a.n_kv - 1 : -1;
    a.total_work = B * NH * a.n_items;
    static long long* const trace_ptr = getenv("MD_ATT_TRACE_PTR") ? reinterpret_cast<long long*>(strtoull(getenv("MD_ATT_TRACE_PTR"), nullptr, 0)) : nullptr;
    a.trace = trace_ptr;     // tools/att_trace.py sets it before the first call; read once
    CUtensorMap tmo;
    if (int e = make_tmap_bf16_3d(&tmo, out, H, L, B, H, (uint64_t)L * H, 64, 128)) return e;
    typedef void (*KernelFn)(const CUtensorMap, const CUtensorMap, const AttArgs);
    // tuning switch: MD_ATT_POLY = how many of every 8 element pairs compute 2^x on the FMA pipe instead of the MUFU.
    // Alone at full clock 0 and 3 are fastest (4.39-4.44 ms at B = 256 against 4.58 for 2); inside the power-capped step
    // the clock is set by the energy of the whole step and the variant with the fewest instructions (0: MUFU only) gives
    // the shortest in-step attention time (profiles/r2_attention_ab.md) -> default 0
    static const int poly = env_int("MD_ATT_POLY", 0);
    static const KernelFn kern = trace_ptr != nullptr ? attention_kernel<2, true>
                                 : poly == 9 ? attention_kernel<9> : poly >= 4 ? attention_kernel<4> : poly == 3 ? attention_kernel<3>
                                 : poly == 2 ? attention_kernel<2> : poly == 1 ? attention_kernel<1> : attention_kernel<0>;
    static bool attr_set[kMaxDevices] = {false};
    if (ensure_dyn_smem(kern, kAttSmem, attr_set, "cudaFuncSetAttribute(attention)")) return MD_ERR_CUDA;
    static const float thr = getenv("MD_ATT_THR") ? (float)atof(getenv("MD_ATT_THR")) : kRescaleThreshold;
    a.rescale_thr = thr;
    const int grid = a.total_work < num_sms() ? a.total_work : num_sms();
    a.step_it = grid % a.n_items;
    a.step_h = (grid / a.n_items) % NH;
    a.step_b = (grid / a.n_items) / NH;
    kern<<<grid, kAttThreads, kAttSmem, stream>>>(tm, tmo, a);
    return check_cuda(cudaGetLastError(), "attention launch");
}
