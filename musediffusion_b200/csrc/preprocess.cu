// Modification-mode input preparation, batched on the GPU: one warp per (src, trg) pair does
// what the reference's dataset preprocessing does per row in Python —
//   merge_and_mask  (MuseDiffusion/data/preprocess.py:30-58): every chord token of the target and the position token in
//                   front of it move behind the meta; row = [*src, EOS, *trg'], mask 0 over src + EOS, 1 over trg',
//   helper_filter   (:73-81): rows longer than seq_len are reported through their length and left as padding,
//   collate_batches (MuseDiffusion/data/wrapper.py:90-126): zero-padded ids, one-padded mask.
// Fully data-parallel: chord flags by ballot, destinations by warp prefix counts (two passes over the target: count,
// then place).  numpy's negative-index wrap is kept: a chord token at index 0 pairs with the LAST target token.
#include <stdint.h>

#include "common.cuh"
#include "musediff_b200.h"

namespace md {

namespace {
struct MergeArgs {
    const int32_t* src;       // [B, Ls]
    const int32_t* src_len;   // [B]
    const int32_t* trg;       // [B, Lt]
    const int32_t* trg_len;   // [B]
    int32_t* input_ids;       // [B, seq_len]
    int32_t* input_mask;      // [B, seq_len]
    int32_t* length;          // [B]
    int B, Ls, Lt, seq_len, end_token;
};
__device__ __forceinline__ bool is_chord(int t) { return t >= 195 && t <= 303; }   // preprocess.py:42
}  // namespace

__global__ void __launch_bounds__(32) merge_and_mask_kernel(const MergeArgs a) {
    const int b = blockIdx.x;
    const int lane = threadIdx.x;
    const int ns = a.src_len[b], nt = a.trg_len[b];
    const int32_t* src = a.src + (size_t)b * a.Ls;
    const int32_t* trg = a.trg + (size_t)b * a.Lt;
    int32_t* ids = a.input_ids + (size_t)b * a.seq_len;
    int32_t* msk = a.input_mask + (size_t)b * a.seq_len;
    const bool chord0 = nt > 0 && is_chord(trg[0]);

    // pass 1: number of chord tokens and of target tokens that stay
    int n_ch = 0, n_keep = 0;
    for (int base = 0; base < nt; base += 32) {
        const int i = base + lane;
        bool ch = false, keep = false;
        if (i < nt) {
            ch = is_chord(trg[i]);
            // removed: chord tokens, the token in front of a chord token, and (index -1 wrap) the last token if trg[0] is a chord
            keep = !(ch || (i + 1 < nt && is_chord(trg[i + 1])) || (i == nt - 1 && chord0));
        }
        n_ch += __popc(__ballot_sync(0xffffffffu, ch));
        n_keep += __popc(__ballot_sync(0xffffffffu, keep));
    }
    const int n_src2 = ns + 2 * n_ch;                 // src + moved (position, chord) pairs
    const int total = n_src2 + 1 + n_keep;
    if (lane == 0) a.length[b] = total;
    if (total > a.seq_len) {                           // helper_filter drops it: leave an all-padding row
        for (int i = lane; i < a.seq_len; i += 32) { ids[i] = 0; msk[i] = 1; }
        return;
    }
    for (int i = lane; i < ns; i += 32) { ids[i] = src[i]; msk[i] = 0; }
    // pass 2: place the pairs behind src and the kept tokens behind the EOS
    int c_ch = 0, c_keep = 0;
    for (int base = 0; base < nt; base += 32) {
        const int i = base + lane;
        bool ch = false, keep = false;
        int t = 0;
        if (i < nt) {
            t = trg[i];
            ch = is_chord(t);
            keep = !(ch || (i + 1 < nt && is_chord(trg[i + 1])) || (i == nt - 1 && chord0));
        }
        const unsigned m_ch = __ballot_sync(0xffffffffu, ch), m_keep = __ballot_sync(0xffffffffu, keep);
        const unsigned below = (1u << lane) - 1;
        if (ch) {
            const int k = c_ch + __popc(m_ch & below);
            ids[ns + 2 * k] = trg[i == 0 ? nt - 1 : i - 1];
            ids[ns + 2 * k + 1] = t;
            msk[ns + 2 * k] = 0;
            msk[ns + 2 * k + 1] = 0;
        }
        if (keep) {
            const int k = c_keep + __popc(m_keep & below);
            ids[n_src2 + 1 + k] = t;
            msk[n_src2 + 1 + k] = 1;
        }
        c_ch += __popc(m_ch);
        c_keep += __popc(m_keep);
    }
    if (lane == 0) { ids[n_src2] = a.end_token; msk[n_src2] = 0; }
    for (int i = total + lane; i < a.seq_len; i += 32) { ids[i] = 0; msk[i] = 1; }
}

}  // namespace md

using namespace md;

extern "C" __attribute__((visibility("default"))) int md_merge_and_mask(const int32_t* src, const int32_t* src_len, const int32_t* trg,
                                                                       const int32_t* trg_len, int B, int Ls, int Lt, int seq_len,
                                                                       int end_token, int32_t* input_ids, int32_t* input_mask,
                                                                       int32_t* length, cudaStream_t stream) {
    if (B < 0 || Ls < 0 || Lt < 0 || seq_len <= 0) { set_last_error("md_merge_and_mask: bad shape B=%d Ls=%d Lt=%d seq_len=%d", B, Ls, Lt, seq_len); return MD_ERR_ARG; }
    if (B == 0) return MD_OK;
    if (!src_len || !trg_len || !input_ids || !input_mask || !length || (Ls > 0 && !src) || (Lt > 0 && !trg)) {
        set_last_error("md_merge_and_mask: null pointer");
        return MD_ERR_ARG;
    }
    MergeArgs a;
    a.src = src; a.src_len = src_len; a.trg = trg; a.trg_len = trg_len; a.input_ids = input_ids; a.input_mask = input_mask;
    a.length = length; a.B = B; a.Ls = Ls; a.Lt = Lt; a.seq_len = seq_len; a.end_token = end_token;
    merge_and_mask_kernel<<<B, 32, 0, stream>>>(a);
    return check_cuda(cudaGetLastError(), "merge_and_mask launch");
}
