// Token-level half of the post-sampling decode, batched on the GPU: one warp per sampled
// sequence does everything `SequenceToMidi.decode` does before the MIDI writer is called
// (MuseDiffusion/utils/decode_util.py:192-199 split_meta_midi, :72-82 remove_padding, :84-141 restore_chord,
// :143-155 validate_once, :157-184 validate_rigidly) and reports the reference's outcome as a status code, so the
// rank-sequential host loop of run/sample.py:222-294 only has to write the MIDI files of the rows that survived.
//
// The token row is staged in shared memory with coalesced loads; the scans that are data-parallel (mask sum, first
// EOS, BAR list, the (position, velocity, pitch, duration) 4-gram search, the copy-out) are warp-cooperative, the
// splice list of restore_chord and the strict grammar walk are sequential by nature and run on lane 0 over shared
// memory.  Integer work: the parity tests require bit-exact agreement with the reference's outcome on every row.
#include <stdint.h>

#include "common.cuh"
#include "musediff_b200.h"

namespace md {

namespace {
constexpr int TOK_EOS = 1, TOK_BAR = 2, TOK_PITCH = 3, TOK_VEL = 131, TOK_CHORD = 195, TOK_DUR = 304, TOK_POS = 432, TOK_BPM = 560;

struct DecodeArgs {
    const int32_t* tokens;   // [B, L]
    const int32_t* mask;     // [B, L]
    int32_t* status;         // [B]
    int32_t* note_len;       // [B]
    int32_t* notes;          // [B, 2L]
    int32_t* meta;           // [B, 11]
    int B, L, strict;
};

// numpy indexing of the reference: negative indices wrap, anything else out of range raises IndexError
__device__ __forceinline__ bool np_at(const int32_t* s, int n, int i, int& v) {
    if (i < 0) i += n;
    if (i < 0 || i >= n) return false;
    v = s[i];
    return true;
}
}  // namespace

__global__ void __launch_bounds__(32) decode_prepare_kernel(const DecodeArgs a) {
    extern __shared__ int32_t sm[];
    const int L = a.L;
    int32_t* row = sm;                 // [L]   the sampled row
    int32_t* seq = row + L;            // [2L]  note part after remove_padding (+ the BARs restore_chord may insert)
    int32_t* out = seq + 2 * L;        // [2L]  restored note sequence
    int32_t* bars = out + 2 * L;       // [2L]  indices of BAR tokens in seq
    const int b = blockIdx.x;
    const int lane = threadIdx.x;
    const int32_t* tok = a.tokens + (size_t)b * L;
    const int32_t* msk = a.mask + (size_t)b * L;

    // ---- split_meta_midi (decode_util.py:192-196): len_meta = len(seq) - sum(mask)
    int msum = 0;
    for (int i = lane; i < L; i += 32) {
        row[i] = tok[i];
        msum += msk[i];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) msum += __shfl_xor_sync(0xffffffffu, msum, o);
    __syncwarp();
    const int len_meta = L - msum;
    // encoded_meta = seq[:len_meta - 1], note_seq = seq[len_meta:]   (python slice semantics for odd masks)
    int m_end = len_meta - 1;
    if (m_end < 0) m_end = max(L + m_end, 0);
    m_end = min(m_end, L);
    int n_start = len_meta < 0 ? max(L + len_meta, 0) : min(len_meta, L);
    const int n_chord = max(m_end - 11, 0);
    const int32_t* chord = row + 11;

    // ---- remove_padding (:72-82): cut after the first EOS
    int eos = -1;
    for (int base = n_start; base < L && eos < 0; base += 32) {
        const int i = base + lane;
        const unsigned hit = __ballot_sync(0xffffffffu, i < L && row[i] == TOK_EOS);
        if (hit) eos = base + __ffs(hit) - 1;
    }
    int status = MD_DECODE_OK;
    int n = 0, out_len = 0;
    bool have_notes = false;
    if (eos < 0) {
        status = MD_DECODE_NO_EOS;
    } else {
        n = eos - n_start + 1;
        // ---- BAR list of the note part and chord-bar count (:91-93), warp-cooperative
        int n_bars = 0, n_cb = 0;
        for (int base = 0; base < n; base += 32) {
            const int i = base + lane;
            const int v = i < n ? row[n_start + i] : -1;
            if (i < n) seq[i] = v;
            const unsigned hit = __ballot_sync(0xffffffffu, v == TOK_BAR);
            if (v == TOK_BAR) bars[n_bars + __popc(hit & ((1u << lane) - 1))] = i;
            n_bars += __popc(hit);
        }
        for (int base = 0; base < n_chord; base += 32) {
            const int i = base + lane;
            n_cb += __popc(__ballot_sync(0xffffffffu, i < n_chord && chord[i] == TOK_POS));
        }
        __syncwarp();
        if (lane == 0) {
            // ---- restore_chord (:84-141), sequential splice list
            int first = 0;
            if (n_bars == n_cb) {
                first = 0;
            } else if (n_bars == n_cb + 1) {
                first = 1;
            } else if (n_bars < n_cb) {
                // np.insert(seq, -1, 2) x diff: BARs in front of the last token
                const int diff = n_cb - n_bars;
                const int last_tok = seq[n - 1];
                for (int k = 0; k < diff; ++k) {
                    seq[n - 1 + k] = TOK_BAR;
                    bars[n_bars + k] = n - 1 + k;
                }
                n += diff;
                seq[n - 1] = last_tok;
                n_bars = n_cb;
                first = 0;
            } else {
                status = MD_DECODE_RESTORE_FAILED;
            }
            if (status == MD_DECODE_OK && first >= n_bars) status = MD_DECODE_INDEX_ERROR;     // bar_idx[first] raises
            if (status == MD_DECODE_OK) {
                const int cap = 2 * L;                       // beyond it only the length is tracked (-> MD_DECODE_TOO_LONG)
                auto copy = [&](int lo, int hi) {            // out += seq[lo:hi]
                    lo = max(lo, 0);
                    hi = min(hi, n);
                    for (int i = lo; i < hi; ++i) {
                        if (out_len < cap) out[out_len] = seq[i];
                        ++out_len;
                    }
                };
                auto emit_chord = [&](int i) {               // out += chord_info[i:i+2]
                    for (int k = i; k < min(i + 2, n_chord); ++k) {
                        if (out_len < cap) out[out_len] = chord[k];
                        ++out_len;
                    }
                };
                copy(0, bars[first] + 1);
                emit_chord(0);
                int bar_count = first;
                int last = bars[first];
                for (int i = 2; i < n_chord && status == MD_DECODE_OK; i += 2) {
                    if (chord[i] == TOK_POS) {
                        if (bar_count + 1 >= n_bars) { status = MD_DECODE_INDEX_ERROR; break; }
                        copy(last + 1, bars[bar_count + 1] + 1);
                        emit_chord(i);
                        ++bar_count;
                        last = bars[bar_count];
                    } else {
                        // last index c inside the current bar with POSITION <= seq[c] < chord_info[i]
                        const int lo = bars[bar_count];
                        const int hi = (bar_count != n_bars - 1) ? bars[bar_count + 1] : n;
                        int c = -1;
                        for (int k = hi - 1; k > lo; --k)
                            if (seq[k] >= TOK_POS && seq[k] < chord[i]) { c = k; break; }
                        if (c < 0) {
                            emit_chord(i);
                        } else {
                            copy(last + 1, c + 4);
                            emit_chord(i);
                            last = c + 3;
                        }
                    }
                }
                if (status == MD_DECODE_OK) copy(last + 1, n);
                if (status == MD_DECODE_OK && out_len > cap) status = MD_DECODE_TOO_LONG;
            }
            if (status != MD_DECODE_OK) out_len = 0;
        }
        status = __shfl_sync(0xffffffffu, status, 0);
        out_len = __shfl_sync(0xffffffffu, out_len, 0);
        __syncwarp();
        have_notes = (status == MD_DECODE_OK);
        if (have_notes) {
            // ---- validate_once (:143-155): any idx <= len - 3 with (seq[idx-1], seq[idx], seq[idx+1], seq[idx+2]) a note;
            //      seq[idx - 1] at idx = 0 reads the last token, as numpy does
            bool found = false;
            for (int base = 0; base < out_len - 2 && !found; base += 32) {
                const int idx = base + lane;
                bool hit = false;
                if (idx + 2 <= out_len - 1) {
                    const int t = out[idx], tp = out[idx == 0 ? out_len - 1 : idx - 1], t1 = out[idx + 1], t2 = out[idx + 2];
                    hit = t >= TOK_VEL && t < TOK_CHORD && tp >= TOK_POS && tp < TOK_BPM && t1 >= TOK_PITCH && t1 < TOK_VEL &&
                          t2 >= TOK_DUR && t2 < TOK_POS;
                }
                found = __any_sync(0xffffffffu, hit);
            }
            if (!found) status = MD_DECODE_VALIDATION_FAILED;
            // ---- validate_rigidly (:157-184), sequential walk
            if (found && a.strict) {
                if (lane == 0) {
                    int i = 0;
                    int st = MD_DECODE_STRICT_FAILED;
                    while (i < out_len) {
                        const int t = out[i];
                        if (t == TOK_EOS) { st = MD_DECODE_OK; break; }
                        if (t == TOK_BAR) { ++i; continue; }
                        if (!(t >= TOK_POS && t < TOK_BPM)) break;
                        int t1, t2, t3;
                        if (!np_at(out, out_len, i + 1, t1)) { st = MD_DECODE_INDEX_ERROR; break; }
                        if (t1 >= TOK_VEL && t1 < TOK_CHORD) {
                            // all([seq[i+2] in ..., seq[i+3] in ...]) evaluates both look-aheads first
                            if (!np_at(out, out_len, i + 2, t2) || !np_at(out, out_len, i + 3, t3)) { st = MD_DECODE_INDEX_ERROR; break; }
                            if (t2 >= TOK_PITCH && t2 < TOK_VEL && t3 >= TOK_DUR && t3 < TOK_POS) { i += 4; continue; }
                            break;
                        }
                        if (t1 >= TOK_CHORD && t1 < TOK_DUR) { i += 2; continue; }
                        break;
                    }
                    status = st;
                }
                status = __shfl_sync(0xffffffffu, status, 0);
            }
        }
    }
    // ---- outputs (zero padded)
    int32_t* o_notes = a.notes + (size_t)b * 2 * L;
    const int keep = have_notes ? out_len : 0;
    for (int i = lane; i < 2 * L; i += 32) o_notes[i] = i < keep ? out[i] : 0;
    if (lane < 11) a.meta[(size_t)b * 11 + lane] = (have_notes && lane < m_end) ? row[lane] : 0;
    if (lane == 0) {
        a.status[b] = status;
        a.note_len[b] = keep;
    }
}

}  // namespace md

using namespace md;

extern "C" __attribute__((visibility("default"))) int md_decode_prepare(const int32_t* tokens, const int32_t* mask, int B, int L,
                                                                       int strict, int32_t* status, int32_t* note_len,
                                                                       int32_t* notes, int32_t* meta, cudaStream_t stream) {
    if (B < 0 || L <= 0) { set_last_error("md_decode_prepare: bad shape B=%d L=%d", B, L); return MD_ERR_ARG; }
    if (B == 0) return MD_OK;
    if (!tokens || !mask || !status || !note_len || !notes || !meta) { set_last_error("md_decode_prepare: null pointer"); return MD_ERR_ARG; }
    const size_t smem = (size_t)7 * L * sizeof(int32_t);
    if (smem > 200 * 1024) { set_last_error("md_decode_prepare: seq_len %d needs %zu B of shared memory (limit 200 KB)", L, smem); return MD_ERR_ARG; }
    static size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
        if (check_cuda(cudaFuncSetAttribute(decode_prepare_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                       "cudaFuncSetAttribute(decode_prepare)"))
            return MD_ERR_CUDA;
        configured = smem;
    }
    DecodeArgs a;
    a.tokens = tokens; a.mask = mask; a.status = status; a.note_len = note_len; a.notes = notes; a.meta = meta;
    a.B = B; a.L = L; a.strict = strict;
    decode_prepare_kernel<<<B, 32, smem, stream>>>(a);
    return check_cuda(cudaGetLastError(), "decode_prepare launch");
}
