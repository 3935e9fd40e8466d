// Sample-quality metrics on the decoded note sequences, batched on the GPU:
//   get_vectors (MuseDiffusion/metric.py:4-75): rhythm [32] / melody [12] / harmony [12] vector of every sequence — one warp
//     per sequence walks the tokens in lockstep (the walk is a state machine, but the 32 rhythm slots of a bar are
//     independent: lane k owns slot k; lanes 0..11 own the melody / harmony bins);
//   the token counts behind Controllability_Pitch / Controllability_Velocity (:131-169) by a lane-strided pass;
//   MSIM Gram matrix + nearest neighbour of ONNC (:89-109): one warp per row of the N x N similarity.
// The reference computes note amplitudes in Python floats (float64) and rounds them when they are stored into the
// float32 rhythm vector; the same is done here.  Norms are warp-tree sums (the reference's are torch.norm): vectors
// agree to ~1e-7, the tests allow 1e-5.
#include <stdint.h>

#include "common.cuh"
#include "musediff_b200.h"

namespace md {

namespace {
struct MetricArgs {
    const int32_t* notes;     // [B, Ln]
    const int32_t* note_len;  // [B]
    const int32_t* meta;      // [B, 11]
    float* vectors;           // [B, 56] = rhythm 32 | melody 12 | harmony 12
    int32_t* status;          // [B] 0 ok, 1 the reference raises on this sequence
    int32_t* stats;           // [B, 4] = pitch token sum, pitch token count, velocity token count, velocity tokens out of range
    int B, Ln;
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
}  // namespace

__global__ void __launch_bounds__(32) sequence_metrics_kernel(const MetricArgs a) {
    const int b = blockIdx.x, lane = threadIdx.x;
    const int n = a.note_len[b];
    const int32_t* midi = a.notes + (size_t)b * a.Ln;
    const int32_t* meta = a.meta + (size_t)b * 11;

    // ---- Controllability_Pitch / _Velocity counts (metric.py:131-169)
    {
        const int min_vel = meta[7] - 524, max_vel = meta[8] - 524;
        int psum = 0, pcnt = 0, vcnt = 0, vwrong = 0;
        for (int i = lane; i < n; i += 32) {
            const int t = midi[i];
            if (t >= 3 && t <= 130) { psum += t; ++pcnt; }
            if (t >= 131 && t <= 194) {
                ++vcnt;
                if (!((min_vel == 130 || min_vel <= t) && (max_vel == 195 || t <= max_vel))) ++vwrong;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            psum += __shfl_xor_sync(0xffffffffu, psum, o);
            pcnt += __shfl_xor_sync(0xffffffffu, pcnt, o);
            vcnt += __shfl_xor_sync(0xffffffffu, vcnt, o);
            vwrong += __shfl_xor_sync(0xffffffffu, vwrong, o);
        }
        if (lane == 0) {
            int32_t* s = a.stats + (size_t)b * 4;
            s[0] = psum; s[1] = pcnt; s[2] = vcnt; s[3] = vwrong;
        }
    }

    // ---- get_vectors (metric.py:4-75): every lane walks the same tokens; lane k owns rhythm slot k, lanes < 12 the bins
    float rhythm = 1e-8f, tmp = 1e-8f;
    float melody = 1e-8f, harmony = 0.f;          // meaningful on lanes 0..11
    int cur = -1, prev = -1, prev_startp = -1, startp = -1;
    bool have_startp = false, ok = true;
    int i = 0;
    while (i < n && midi[i] != 2) ++i;            // first BAR (the reference runs off the end -> IndexError)
    if (i >= n) ok = false;
    ++i;
    while (ok) {
        if (i >= n) { ok = false; break; }
        const int t0 = midi[i];
        if (t0 <= 2) {
            tmp = tmp / sqrtf(warp_sum(tmp * tmp));
            rhythm += tmp;
            tmp = 1e-8f;
            ++i;
            if (t0 == 2) { prev_startp = -1; continue; }
            if (!have_startp) { ok = false; break; }                         // unbound `startp` in the reference
            if (prev_startp != startp && prev >= 0 && lane == ((cur - prev) % 12 + 12) % 12) melody += 1.0f;
            break;
        }
        if (!(t0 >= 432 && t0 <= 559)) { ok = false; break; }                 // "position not found"
        startp = t0 - 432;
        have_startp = true;
        if (i + 1 >= n) { ok = false; break; }
        const int t1 = midi[i + 1];
        if (t1 >= 195 && t1 <= 303) { i += 2; continue; }
        if (i + 3 >= n) { ok = false; break; }
        const int t2 = midi[i + 2], t3 = midi[i + 3];
        if (!(t1 >= 131 && t1 <= 194 && t2 >= 3 && t2 <= 130 && t3 >= 304 && t3 <= 431)) { ok = false; break; }   // "wrong format"
        const int pitch = t2;
        const int endp = startp + t3 - 303;
        if (lane == pitch % 12) harmony += 1.0f;
        {
            const int t = 4 * lane;                                          // for t in range(0, min(128, endp), 4)
            if (t < min(128, endp) && t >= startp) {
                const double amp0 = 0.00542676376 * (double)(t1 - 130) * 2.0 + 0.310801;
                const double amp = amp0 * amp0;
                const double decay = 1.0 - (double)(t - startp) / 128.0;
                const float val = (float)(amp * (decay > 0.0 ? decay : 0.0));
                if (val > tmp) tmp = val;
            }
        }
        if (cur >= 0 && prev_startp != startp) {
            if (prev >= 0 && lane == ((cur - prev) % 12 + 12) % 12) melody += 1.0f;
            prev = cur;
            cur = pitch;
        }
        cur = max(pitch, cur);
        prev_startp = startp;
        i += 4;
    }
    float* v = a.vectors + (size_t)b * 56;
    if (ok) {
        const float rn = sqrtf(warp_sum(rhythm * rhythm));
        const float mn = sqrtf(warp_sum(lane < 12 ? melody * melody : 0.f));
        const float hn = sqrtf(warp_sum(lane < 12 ? harmony * harmony : 0.f));
        v[lane] = rhythm / rn;
        if (lane < 12) {
            v[32 + lane] = melody / mn;
            v[44 + lane] = harmony / hn;
        }
    } else {
        v[lane] = 0.f;
        if (lane < 24) v[32 + lane] = 0.f;
    }
    if (lane == 0) a.status[b] = ok ? 0 : 1;
}

// MSIM Gram matrix and its row arg-max with the diagonal zeroed (metric.py:99-107): one warp per row
__global__ void __launch_bounds__(32) onnc_kernel(const float* __restrict__ vec, int N, float* __restrict__ msim,
                                                  int32_t* __restrict__ most_sim) {
    const int i = blockIdx.x, lane = threadIdx.x;
    __shared__ float vi[56];
    for (int k = lane; k < 56; k += 32) vi[k] = vec[(size_t)i * 56 + k];
    __syncwarp();
    float best = -INFINITY;
    int best_j = 0x7fffffff;
    for (int j = lane; j < N; j += 32) {
        const float* vj = vec + (size_t)j * 56;
        float r = 0.f, m = 0.f, h = 0.f;
#pragma unroll 8
        for (int k = 0; k < 32; ++k) r = fmaf(vi[k], vj[k], r);
#pragma unroll
        for (int k = 0; k < 12; ++k) { m = fmaf(vi[32 + k], vj[32 + k], m); h = fmaf(vi[44 + k], vj[44 + k], h); }
        float s = r * m * h;
        if (j == i) s = 0.f;
        if (msim != nullptr) msim[(size_t)i * N + j] = s;
        if (s > best) { best = s; best_j = j; }            // strict >: the first maximum wins, as torch.argmax
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oj = __shfl_xor_sync(0xffffffffu, best_j, o);
        if (ob > best || (ob == best && oj < best_j)) { best = ob; best_j = oj; }
    }
    if (lane == 0) most_sim[i] = best_j;
}

}  // namespace md

using namespace md;

extern "C" __attribute__((visibility("default"))) int md_sequence_metrics(const int32_t* notes, const int32_t* note_len, const int32_t* meta,
                                                                         int B, int Ln, float* vectors, int32_t* status, int32_t* stats,
                                                                         cudaStream_t stream) {
    if (B < 0 || Ln <= 0) { set_last_error("md_sequence_metrics: bad shape B=%d Ln=%d", B, Ln); return MD_ERR_ARG; }
    if (B == 0) return MD_OK;
    if (!notes || !note_len || !meta || !vectors || !status || !stats) { set_last_error("md_sequence_metrics: null pointer"); return MD_ERR_ARG; }
    MetricArgs a;
    a.notes = notes; a.note_len = note_len; a.meta = meta; a.vectors = vectors; a.status = status; a.stats = stats; a.B = B; a.Ln = Ln;
    sequence_metrics_kernel<<<B, 32, 0, stream>>>(a);
    return check_cuda(cudaGetLastError(), "sequence_metrics launch");
}

extern "C" __attribute__((visibility("default"))) int md_onnc(const float* vectors, int N, float* msim, int32_t* most_sim, cudaStream_t stream) {
    if (N < 0) { set_last_error("md_onnc: bad N=%d", N); return MD_ERR_ARG; }
    if (N == 0) return MD_OK;
    if (!vectors || !most_sim) { set_last_error("md_onnc: null pointer"); return MD_ERR_ARG; }
    onnc_kernel<<<N, 32, 0, stream>>>(vectors, N, msim, most_sim);
    return check_cuda(cudaGetLastError(), "onnc launch");
}
