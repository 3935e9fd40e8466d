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
// One persistent CTA per SM works on TWO 128-row Q tiles of the same (b, h) so the tensor pipe computes S for one
// tile while the other tile's softmax runs, and both tiles share every K/V stage brought in by TMA.
//
// Roles (384 threads = 3 warpgroups): warpgroup 0 = {warp 0: TMA producer, warp 1: TMEM owner + MMA issuer of Q
// tile 0, warp 2: MMA issuer of Q tile 1, warp 3: idle}, warpgroup 1 = softmax/correction/epilogue for Q tile 0, warpgroup 2 = same for Q tile 1.
// setmaxnreg moves registers from warpgroup 0 to the softmax warpgroups (a full S row lives in registers).  At head dim
// 64 the MUFU / FMA / ALU mix of the softmax, not the tensor pipe, is the binding resource; a pair of named barriers
// can make the two softmax warpgroups take turns on the exp2 phase (MD_ATT_TURNS, off by default: measured slower once
// the P V wait had been moved behind the exponentials).
#include <stdlib.h>

#include "common.cuh"
#include "musediff_b200.h"

namespace md {
int num_sms();
int make_tmap_bf16_3d(CUtensorMap* tm, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1_elems,
                      uint64_t stride2_elems, uint32_t box0, uint32_t box1);

constexpr int ATT_BQ = 128;        // rows per Q tile
constexpr int ATT_BKV = 128;       // keys per K/V stage
constexpr int ATT_DH = 64;
constexpr int ATT_STAGES = 3;
// kSplit = softmax warps per TMEM lane quadrant of a Q tile: 1 -> one thread owns a whole 128-key row of S,
// 2 -> two threads own 64 keys each (twice the warps to hide the serial LDTM -> max -> exp -> STTM chain).
template <int kSplit> struct AttCfg;
template <> struct AttCfg<1> { static constexpr int kThreads = 384, kRegsProducer = 72, kRegsSoftmax = 216; };   // 128*72 + 256*216 = 64512 = 384 x 168
template <> struct AttCfg<2> { static constexpr int kThreads = 640, kRegsProducer = 56, kRegsSoftmax = 104; };   // 128*56 + 512*104 = 60416 <= 640 x 96
constexpr int ATT_TILE_BYTES = 128 * 64 * 2;   // every smem tile is 128 rows x 128 B
constexpr int kAttXchgBytes = (2 * 2 * 2 * 128 + 2 * 2 * 128) * 4;   // row-max (double buffered) and row-sum exchange between the two column halves
constexpr int kAttSmem = 1024 + (4 + 2 * ATT_STAGES + 2) * ATT_TILE_BYTES + 512 + kAttXchgBytes;   // Q double-buffered across work items; 2 output staging tiles
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
    int n_pairs;        // ceil(ceil(L/128) / 2)
    int n_qtiles;       // ceil(L/128)
    int n_kv;           // ceil(L/128)
    int total_work;     // B * NH * n_pairs
    __nv_bfloat16* out; // [B*L, NH*64]
    float rescale_thr;  // lazy-rescale threshold in log2 units (see kRescaleThreshold)
    int turn_every;     // 1: the two Q tiles alternate on the exp2 phase every key block; 0: only on block 0 (phase offset)
    long long* trace;   // optional [role 10][event 8][step 64] clock64 stamps of CTA 0 (debug / tuning)
};
#define MD_TRACE(role, ev, step)                                                                          \
    do {                                                                                                  \
        if (kTrace && a.trace != nullptr && blockIdx.x == 0 && lane == 0 && (step) < 64)                  \
            a.trace[((role) * 8 + (ev)) * 64 + (step)] = clock64();                                       \
    } while (0)

struct Work { int b, h, pair; };
MD_DEVINL Work decode_work(int w, const AttArgs& a) {
    Work r;
    r.pair = w % a.n_pairs;
    const int bh = w / a.n_pairs;
    r.h = bh % a.NH;
    r.b = bh / a.NH;
    return r;
}

template <bool kTurns, int kN>
MD_DEVINL void turn_wait(int x) {
    if (!kTurns) return;
    if (x == 0) asm volatile("bar.sync 2, %0;" ::"n"(kN) : "memory");
    else asm volatile("bar.sync 3, %0;" ::"n"(kN) : "memory");
}
template <bool kTurns, int kN>
MD_DEVINL void turn_pass(int x) {
    if (!kTurns) return;
    if (x == 0) asm volatile("bar.arrive 3, %0;" ::"n"(kN) : "memory");
    else asm volatile("bar.arrive 2, %0;" ::"n"(kN) : "memory");
}

template <bool kTurns, int kPoly, int kSplit, bool kTrace = false>
__global__ void __launch_bounds__(AttCfg<kSplit>::kThreads, 1)
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
    uint64_t* s_full = v_empty + ATT_STAGES;   // [2]
    uint64_t* p_full = s_full + 2;             // [2]
    uint64_t* o_full = p_full + 2;             // [2]
    uint64_t* o_empty = o_full + 2;            // [2]
    uint64_t* s_free = o_empty + 2;            // [2] softmax has read S into registers -> next Q K^T may overwrite it
    uint64_t* p_free = s_free + 2;             // [2] last key block only: P V of block n-2 retired (earlier blocks learn it from s_full(j+1))
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(p_free + 2);
    float* sMax = reinterpret_cast<float*>(bars) + 128;     // [tile][buf][half][128]   (barriers occupy < 512 B)
    float* sSum = sMax + 2 * 2 * 2 * 128;                   // [tile][half][128]
    constexpr int kTileThreads = 128 * kSplit;              // softmax threads per Q tile

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int H = a.NH * ATT_DH;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQKV);
        tma_prefetch_desc(&tmOut);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&q_full[i], 1);
            mbar_init(&q_empty[i], 2);         // both MMA warps (one per Q tile) release the shared stages
        }
        for (int s = 0; s < ATT_STAGES; ++s) {
            mbar_init(&k_full[s], 1);
            mbar_init(&k_empty[s], 2);
            mbar_init(&v_full[s], 1);
            mbar_init(&v_empty[s], 2);
        }
        for (int x = 0; x < 2; ++x) {
            mbar_init(&s_full[x], 1);
            mbar_init(&p_full[x], kTileThreads);
            mbar_init(&o_full[x], 1);
            mbar_init(&o_empty[x], kTileThreads);
            mbar_init(&s_free[x], kTileThreads);
            mbar_init(&p_free[x], 1);
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(AttCfg<kSplit>::kRegsProducer));
    if (warp == 0) {
        // =========================================================== TMA producer (whole warp, elected issue)
        {
            uint32_t wcnt = 0;
            uint32_t st = 0, ph = 0;               // K/V ring position, carried across work items
            for (int w = blockIdx.x; w < a.total_work; w += gridDim.x, ++wcnt) {
                const Work wk = decode_work(w, a);
                const int qt0 = wk.pair * 2;
                const int n_active = (qt0 + 1 < a.n_qtiles) ? 2 : 1;
                const int qb = wcnt & 1;               // Q buffer of this work item
                mbar_wait_idle(&q_empty[qb], ((wcnt >> 1) & 1) ^ 1);
                mbar_arrive_expect_tx_w(&q_full[qb], n_active * ATT_TILE_BYTES);
                for (int x = 0; x < n_active; ++x)
                    tma_load_3d_w(sQ + (qb * 2 + x) * ATT_TILE_BYTES, &tmQKV, &q_full[qb], wk.h * ATT_DH, (qt0 + x) * ATT_BQ, wk.b);
                for (int j = 0; j < a.n_kv; ++j) {
                    mbar_wait_idle(&k_empty[st], ph ^ 1);
                    mbar_arrive_expect_tx_w(&k_full[st], ATT_TILE_BYTES);
                    tma_load_3d_w(sK + st * ATT_TILE_BYTES, &tmQKV, &k_full[st], H + wk.h * ATT_DH, j * ATT_BKV, wk.b);
                    mbar_wait_idle(&v_empty[st], ph ^ 1);
                    mbar_arrive_expect_tx_w(&v_full[st], ATT_TILE_BYTES);
                    tma_load_3d_w(sV + st * ATT_TILE_BYTES, &tmQKV, &v_full[st], 2 * H + wk.h * ATT_DH, j * ATT_BKV, wk.b);
                    if (++st == ATT_STAGES) { st = 0; ph ^= 1; }
                }
            }
        }
    } else {
        // =========================================================== MMA issuers: warp 1 -> Q tile 0, warp 2 -> Q tile 1
        // (whole warp runs the loop, one elected lane issues).  One issuing warp per tile: the issue stream of a
        // single warp (waits, descriptor moves, 24 small MMAs per key block) was measured to be the critical path.
        if (warp == 1 || warp == 2) {
            const int x = warp - 1;
            constexpr uint32_t idesc_qk = make_idesc_bf16(ATT_BQ, ATT_BKV, 0);
            constexpr uint32_t idesc_pv = make_idesc_bf16(ATT_BQ, ATT_DH, 1);   // B (= V) is MN-major
            const uint32_t tS = tmem_base + (x ? TM_S1 : TM_S0);
            const uint32_t tP = tmem_base + (x ? TM_P1 : TM_P0);
            const uint32_t tO = tmem_base + (x ? TM_O1 : TM_O0);
            uint32_t wcnt = 0;
            uint32_t st = 0, ph = 0;   // K/V ring position of key block j, carried across work items
            uint32_t pcnt = 0;   // P tiles consumed (phase of p_full)
            uint32_t ocnt = 0;   // work items (phase of o_empty)
            uint32_t qcnt = 0;   // Q K^T issued (phase of s_free)
            for (int w = blockIdx.x; w < a.total_work; w += gridDim.x, ++wcnt) {
                const Work wk = decode_work(w, a);
                const bool active = (wk.pair * 2 + x < a.n_qtiles);
                const int qb = wcnt & 1;
                const uint64_t qd = make_sdesc_sw128(smem_u32(sQ + (qb * 2 + x) * ATT_TILE_BYTES));
                mbar_wait_idle(&q_full[qb], (wcnt >> 1) & 1);
                {   // S(0)
                    mbar_wait_idle(&k_full[st], ph);
                    if (active) {
                        if (qcnt > 0) mbar_wait_idle(&s_free[x], (qcnt - 1) & 1);
                        ++qcnt;
                        tc_fence_after();
                        umma_qk64_commit_w(tS, qd, make_sdesc_sw128(smem_u32(sK + st * ATT_TILE_BYTES)), idesc_qk, &s_full[x]);
                    }
                    tc_commit_w(&k_empty[st]);
                }
                for (int j = 0; j < a.n_kv; ++j) {
                    const bool has_next = (j + 1 < a.n_kv);
                    uint32_t st_n = st + 1, ph_n = ph;
                    if (st_n == ATT_STAGES) { st_n = 0; ph_n ^= 1; }
                    if (has_next) {
                        // S(j+1): needs only K[j+1] and the softmax warpgroup's READ of S(j)
                        mbar_wait_idle(&k_full[st_n], ph_n);
                        if (active) {
                            mbar_wait_idle(&s_free[x], (qcnt - 1) & 1);
                            ++qcnt;
                            tc_fence_after();
                            MD_TRACE(x, 0, (int)(wcnt * a.n_kv + j));
                            umma_qk64_commit_w(tS, qd, make_sdesc_sw128(smem_u32(sK + st_n * ATT_TILE_BYTES)), idesc_qk, &s_full[x]);
                        }
                        tc_commit_w(&k_empty[st_n]);
                    }
                    if (active && !has_next && j > 0) tc_commit_w(&p_free[x]);     // tracks P V(n-2) for the last block's softmax
                    mbar_wait_idle(&v_full[st], ph);
                    if (active) {
                        if (j == 0) mbar_wait_idle(&o_empty[x], (ocnt & 1) ^ 1);   // previous item's O drained
                        MD_TRACE(x, 1, (int)(wcnt * a.n_kv + j));
                        mbar_wait_idle(&p_full[x], pcnt & 1);
                        ++pcnt;
                        tc_fence_after();
                        MD_TRACE(x, 2, (int)(wcnt * a.n_kv + j));
                        const uint64_t vd = make_sdesc_sw128(smem_u32(sV + st * ATT_TILE_BYTES));
                        if (has_next) {
                            umma_pv128_w(tO, tP, vd, idesc_pv, j > 0 ? 1u : 0u);    // covered by the commit behind Q K^T(j+2)
                        } else {
                            umma_pv128_commit_w(tO, tP, vd, idesc_pv, j > 0 ? 1u : 0u, &o_full[x]);
                            ++ocnt;
                        }
                    }
                    tc_commit_w(&v_empty[st]);
                    st = st_n;
                    ph = ph_n;
                }
                tc_commit_w(&q_empty[qb]);
            }
        }
    }
    } else {
        // =========================================================== softmax / correction / epilogue
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(AttCfg<kSplit>::kRegsSoftmax));
        constexpr int NC = ATT_BKV / kSplit;     // keys (S columns) per thread
        constexpr int NG = NC / 32;              // 32-column chunks per thread
        constexpr int OC = ATT_DH / kSplit;      // O columns per thread
        const int x = (warp - 4) / (4 * kSplit);            // which Q tile this warp works on
        const int half = ((warp - 4) % (4 * kSplit)) >> 2;  // which column half (kSplit == 2)
        // turn-taking on the exp2 phase: tile x syncs on barrier (2 + x) and hands over by arriving on (3 - x)
        if (kTurns && x == 1) turn_pass<kTurns, 2 * kTileThreads>(1);   // tile 0 goes first
        const int quad = warp & 3;              // TMEM lane quadrant
        const int r = quad * 32 + lane;         // row inside the Q tile
        const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
        const uint32_t tS = tmem_base + (x ? TM_S1 : TM_S0) + lane_addr + half * NC;
        const uint32_t tP = tmem_base + (x ? TM_P1 : TM_P0) + lane_addr + half * (NC / 2);
        const uint32_t tO = tmem_base + (x ? TM_O1 : TM_O0) + lane_addr + half * OC;
        uint32_t scnt = 0, ocnt = 0, fcnt = 0;
        for (int w = blockIdx.x; w < a.total_work; w += gridDim.x) {
            const Work wk = decode_work(w, a);
            const int qt = wk.pair * 2 + x;
            if (qt >= a.n_qtiles) {
                // phantom tile of the last pair: keep the turn-taking handshake in step with the other tile
                for (int j = 0; j < (a.turn_every ? a.n_kv : 1); ++j) {
                    turn_wait<kTurns, 2 * kTileThreads>(x);
                    turn_pass<kTurns, 2 * kTileThreads>(x);
                }
                continue;
            }
            float m_used = -INFINITY, l_sum = 0.f;
            // One mbarrier wait per key block: S(j+1)'s commit also covers P V(j-1) (tcgen05.commit tracks every earlier
            // MMA of the issuing thread), so waiting for it after the exponentials of block j both frees P / O for
            // rewriting and pre-pays the wait at the top of block j+1.
            mbar_wait_idle(&s_full[x], scnt & 1);      // long at a work-item boundary (a whole item for a phantom tile's warps): sleep
            tc_fence_after();
            for (int j = 0; j < a.n_kv; ++j, ++scnt) {
                const bool tr = (half == 0);
                if (tr) MD_TRACE(2 + x * 4 + quad, 0, (int)scnt);
                if (tr) MD_TRACE(2 + x * 4 + quad, 1, (int)scnt);
                uint32_t s[NG][32];
#pragma unroll
                for (int g = 0; g < NG; ++g) tmem_ld32(tS + 32 * g, s[g]);
                tc_wait_ld();
                tc_fence_before();
                mbar_arrive(&s_free[x]);               // S is in registers: the next Q K^T may overwrite it
                if (tr) MD_TRACE(2 + x * 4 + quad, 2, (int)scnt);
                const int valid = a.L - j * ATT_BKV - half * NC;   // keys of this thread's columns that exist
                if (valid < NC) {
#pragma unroll
                    for (int g = 0; g < NG; ++g)
#pragma unroll
                        for (int c = 0; c < 32; ++c)
                            if (g * 32 + c >= valid) s[g][c] = 0xff800000u;   // -inf
                }
                float mx0 = __uint_as_float(s[0][0]), mx1 = __uint_as_float(s[0][1]), mx2 = __uint_as_float(s[0][2]),
                      mx3 = __uint_as_float(s[0][3]);
#pragma unroll
                for (int g = 0; g < NG; ++g)
#pragma unroll
                    for (int c = (g == 0 ? 4 : 0); c < 32; c += 4) {
                        mx0 = fmaxf(mx0, __uint_as_float(s[g][c]));
                        mx1 = fmaxf(mx1, __uint_as_float(s[g][c + 1]));
                        mx2 = fmaxf(mx2, __uint_as_float(s[g][c + 2]));
                        mx3 = fmaxf(mx3, __uint_as_float(s[g][c + 3]));
                    }
                float mrow = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
                if (kSplit == 2) {
                    // the two threads of a row exchange their partial maxima (double-buffered by block parity)
                    float* mx = sMax + ((x * 2 + (j & 1)) * 2) * 128;
                    mx[half * 128 + r] = mrow;
                    if (x == 0) asm volatile("bar.sync 4, %0;" ::"n"(kTileThreads) : "memory");
                    else asm volatile("bar.sync 5, %0;" ::"n"(kTileThreads) : "memory");
                    mrow = fmaxf(mrow, mx[(half ^ 1) * 128 + r]);
                }
                if (tr) MD_TRACE(2 + x * 4 + quad, 3, (int)scnt);
                const float mb = mrow * kLog2e;        // -inf only if the whole block is masked for this row: impossible (block 0 ...)
                // lazy rescale decision (the running output lives in TMEM and is only touched after the exponentials,
                // once P V of block j-1 has retired — the wait is then off the critical path)
                // (branch-free: block 0 starts from m_used = -inf, so it always "rescales" an empty sum by 2^-inf = 0)
                const bool need = mb > m_used + a.rescale_thr;
                const bool any_resc = __any_sync(0xffffffffu, need) && j > 0;
                const float f_resc = need ? fast_exp2(m_used - mb) : 1.0f;
                m_used = need ? mb : m_used;
                l_sum *= f_resc;
                // p = 2^(s log2e - m): packed FFMA2 for the argument, then kPoly of every 8 pairs take the FMA-pipe
                // polynomial and the rest the MUFU (16 ex2/clk/SM is the binding unit at head dim 64); packed FADD2 sums.
                uint64_t acc_a = 0, acc_b = 0;     // two independent packed accumulators (bit pattern of +0.0f, +0.0f)
                const uint64_t l2e2 = f2_pack(kLog2e, kLog2e);
                const uint64_t negm2 = f2_pack(-m_used, -m_used);
                if (tr) MD_TRACE(2 + x * 4 + quad, 4, (int)scnt);
                const bool turn = (j == 0) || a.turn_every;
                if (turn) turn_wait<kTurns, 2 * kTileThreads>(x);                       // my turn on the MUFU pipe
                if (tr) MD_TRACE(2 + x * 4 + quad, 5, (int)scnt);
                uint32_t pk[NG][16];
#pragma unroll
                for (int g = 0; g < NG; ++g) {
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
                if (turn) turn_pass<kTurns, 2 * kTileThreads>(x);                       // hand the MUFU pipe to the other tile
                if (tr) MD_TRACE(2 + x * 4 + quad, 6, (int)scnt);
                {
                    if (j + 1 < a.n_kv) {
                        mbar_wait(&s_full[x], (scnt + 1) & 1);      // S(j+1) done => P V(j-1) retired: O and P may be touched
                    } else if (j > 0) {
                        mbar_wait(&p_free[x], fcnt & 1);
                        ++fcnt;
                    }
                    tc_fence_after();
                    if (any_resc) {
#pragma unroll
                        for (int g = 0; g < OC / 32; ++g) {
                            uint32_t o[32];
                            tmem_ld32(tO + 32 * g, o);
                            tc_wait_ld();
#pragma unroll
                            for (int c = 0; c < 32; ++c) o[c] = __float_as_uint(__uint_as_float(o[c]) * f_resc);
                            tmem_st32(tO + 32 * g, o);
                        }
                    }
                }
#pragma unroll
                for (int g = 0; g < NG; ++g) tmem_st16(tP + g * 16, pk[g]);
                {
                    float q0, q1;
                    f2_unpack(f2_add(acc_a, acc_b), q0, q1);
                    l_sum += q0 + q1;
                }
                tc_wait_st();
                tc_fence_before();
                mbar_arrive(&p_full[x]);
                if (tr) MD_TRACE(2 + x * 4 + quad, 7, (int)scnt);
            }
            // ---- epilogue: O / l -> bf16 -> global
            if (kSplit == 2) {
                float* sm = sSum + x * 2 * 128;
                sm[half * 128 + r] = l_sum;
                if (x == 0) asm volatile("bar.sync 4, %0;" ::"n"(kTileThreads) : "memory");
                else asm volatile("bar.sync 5, %0;" ::"n"(kTileThreads) : "memory");
                l_sum += sm[(half ^ 1) * 128 + r];
            }
            mbar_wait_idle(&o_full[x], ocnt & 1);
            ++ocnt;
            tc_fence_after();
            uint32_t o[OC / 32][32];
#pragma unroll
            for (int g = 0; g < OC / 32; ++g) tmem_ld32(tO + 32 * g, o[g]);
            tc_wait_ld();
            tc_fence_before();
            mbar_arrive(&o_empty[x]);
            // Row r -> 128 B of the staging tile (128B swizzle: 16 B chunk ^= row % 8, conflict-free), then ONE bulk
            // tensor store per tile.  (A direct st.global from the row-per-thread layout touches 32 different lines per
            // warp instruction and was measured to stall the other tile's TMEM traffic for ~3000 cycles per work item.)
            // Rows >= L of the last tile are clipped by the tensor map.
            const bool issuer = (r == 0 && half == 0);
            if (issuer) tma_store_wait_read<0>();          // the previous item's store has finished reading the tile
            if (x == 0) asm volatile("bar.sync 6, %0;" ::"n"(kTileThreads) : "memory");
            else asm volatile("bar.sync 7, %0;" ::"n"(kTileThreads) : "memory");
            {
                const float inv = 1.0f / l_sum;
                const uint32_t row_addr = smem_u32(sO + x * ATT_TILE_BYTES) + r * 128;
#pragma unroll
                for (int gg = 0; gg < OC / 32; ++gg)
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        const int chunk = half * (OC / 8) + gg * 4 + g;
                        const uint32_t addr = row_addr + ((chunk ^ (r & 7)) << 4);
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr),
                                     "r"(pack_bf16x2(__uint_as_float(o[gg][g * 8 + 0]) * inv, __uint_as_float(o[gg][g * 8 + 1]) * inv)),
                                     "r"(pack_bf16x2(__uint_as_float(o[gg][g * 8 + 2]) * inv, __uint_as_float(o[gg][g * 8 + 3]) * inv)),
                                     "r"(pack_bf16x2(__uint_as_float(o[gg][g * 8 + 4]) * inv, __uint_as_float(o[gg][g * 8 + 5]) * inv)),
                                     "r"(pack_bf16x2(__uint_as_float(o[gg][g * 8 + 6]) * inv, __uint_as_float(o[gg][g * 8 + 7]) * inv))
                                     : "memory");
                    }
            }
            fence_proxy_async_smem();
            if (x == 0) asm volatile("bar.sync 6, %0;" ::"n"(kTileThreads) : "memory");
            else asm volatile("bar.sync 7, %0;" ::"n"(kTileThreads) : "memory");
            if (issuer) {
                tma_store_3d(&tmOut, sO + x * ATT_TILE_BYTES, wk.h * ATT_DH, qt * ATT_BQ, wk.b);
                tma_store_commit();
            }
        }
    }
    if (warp >= 4 && (threadIdx.x & 127) == 0) tma_store_wait_read<0>();   // staging tiles must outlive the bulk stores' reads
    __syncwarp();
    if (kTurns && warp >= 4 && warp < 4 + 4 * kSplit) turn_wait<kTurns, 2 * kTileThreads>(0);   // absorb tile 1's final hand-over
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<512>(tmem_base);
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
    a.n_pairs = (a.n_qtiles + 1) / 2;
    a.n_kv = (L + ATT_BKV - 1) / ATT_BKV;
    a.total_work = B * NH * a.n_pairs;
    a.out = reinterpret_cast<__nv_bfloat16*>(out);
    static long long* const trace_ptr = getenv("MD_ATT_TRACE_PTR") ? reinterpret_cast<long long*>(strtoull(getenv("MD_ATT_TRACE_PTR"), nullptr, 0)) : nullptr;
    a.trace = trace_ptr;     // tools/att_trace.py sets it before the first call; read once
    CUtensorMap tmo;
    if (int e = make_tmap_bf16_3d(&tmo, out, H, L, B, H, (uint64_t)L * H, 64, 128)) return e;
    typedef void (*KernelFn)(const CUtensorMap, const CUtensorMap, const AttArgs);
    static KernelFn kern = nullptr;
    static int threads = 0;
    static int turn_every = 1;
    if (kern == nullptr) {
        // tuning switches (defaults are the measured best): MD_ATT_TURNS = MUFU turn-taking between the two Q tiles,
        // MD_ATT_POLY = how many of every 8 element pairs compute 2^x on the FMA pipe instead of the MUFU,
        // MD_ATT_SPLIT = softmax threads per S row (1 or 2)
        const char* e = getenv("MD_ATT_TURNS");
        const int turns = e ? atoi(e) : 0;     // 0 none, 1 every key block, 2 first block of a work item only
        turn_every = (turns == 1);
        e = getenv("MD_ATT_POLY");
        const int poly = e ? atoi(e) : 2;
        e = getenv("MD_ATT_SPLIT");
        const int split = e ? atoi(e) : 1;
#define MD_ATT_PICK(T_, S_)                                                                                               \
    (poly == 9 ? attention_kernel<T_, 9, S_> : poly >= 4 ? attention_kernel<T_, 4, S_> : poly == 3 ? attention_kernel<T_, 3, S_> \
     : poly == 2 ? attention_kernel<T_, 2, S_> : poly == 1 ? attention_kernel<T_, 1, S_> : attention_kernel<T_, 0, S_>)
        if (split == 2) kern = turns ? MD_ATT_PICK(true, 2) : MD_ATT_PICK(false, 2);
        else kern = turns ? MD_ATT_PICK(true, 1) : MD_ATT_PICK(false, 1);
#undef MD_ATT_PICK
        if (trace_ptr != nullptr) {      // timeline tracing build of the selected split (tools/att_trace.py)
            if (split == 2) kern = turns ? attention_kernel<true, 0, 2, true> : attention_kernel<false, 0, 2, true>;
            else kern = turns ? attention_kernel<true, 2, 1, true> : attention_kernel<false, 2, 1, true>;
        }
        threads = split == 2 ? AttCfg<2>::kThreads : AttCfg<1>::kThreads;
    }
    static bool attr_set[kMaxDevices] = {false};
    if (ensure_dyn_smem(kern, kAttSmem, attr_set, "cudaFuncSetAttribute(attention)")) return MD_ERR_CUDA;
    a.turn_every = turn_every;
    static const float thr = getenv("MD_ATT_THR") ? (float)atof(getenv("MD_ATT_THR")) : kRescaleThreshold;
    a.rescale_thr = thr;
    const int grid = a.total_work < num_sms() ? a.total_work : num_sms();
    kern<<<grid, threads, kAttSmem, stream>>>(tm, tmo, a);
    return check_cuda(cudaGetLastError(), "attention launch");
}
