"""
ORACLE — TEST INFRASTRUCTURE ONLY.  Never imported by the product (`musediffusion_b200/`).

CPU restatement (numpy) of the step in front of the sampling path in modification mode, SURVEY.md §8(f) row 2:

    merge_and_mask   (MuseDiffusion/data/preprocess.py:30-58)   every chord token of the target and the position token in
                     front of it move behind the meta (`src`); row = [*src, EOS, *trg], mask 0 over src + EOS, 1 over trg
    helper_filter    (:73-81)                                   rows longer than seq_len are dropped
    collate_batches  (MuseDiffusion/data/wrapper.py:90-126)     zero-padded ids, one-padded mask, lengths

numpy indexing semantics are kept (a chord token at index 0 pairs with index -1 = the LAST token, preprocess.py:44-47).
Pinned by tests/golden/merge_and_mask.npz (the unmodified reference's helper_tokenize + collate_batches on the same rows).
"""
import numpy as np

CHORD_LO, CHORD_HI = 195, 303      # preprocess.py:42


def merge_and_mask(src, trg, end_token=1):
    """preprocess.py:38-58 for one (src, trg) pair -> (input_ids, input_mask, length)."""
    src = np.asarray(src, dtype=np.int64)
    trg = np.asarray(trg, dtype=np.int64)
    chord = (trg >= CHORD_LO) & (trg <= CHORD_HI)
    idx = np.repeat(np.nonzero(chord)[0], 2)
    idx[::2] -= 1                                  # the position token in front of each chord token (index -1 wraps)
    keep = np.ones(trg.shape, dtype=bool)
    keep[idx] = False
    src = np.concatenate([src, trg[idx]])
    trg = trg[keep]
    ids = np.concatenate([src, [end_token], trg])
    mask = np.concatenate([np.zeros(len(src) + 1, np.int64), np.ones(len(trg), np.int64)])
    return ids, mask, len(ids)


def collate(rows, seq_len):
    """wrapper.py:90-126 with an explicit seq_len: rows = list of (ids, mask, length), all lengths <= seq_len."""
    B = len(rows)
    input_ids = np.zeros((B, seq_len), np.int64)
    input_mask = np.ones((B, seq_len), np.int64)
    length = np.zeros((B,), np.int64)
    for b, (ids, mask, n) in enumerate(rows):
        input_ids[b, :n] = ids
        input_mask[b, :n] = mask
        length[b] = n
    return input_ids, input_mask, length


def merge_and_mask_batch(src, src_len, trg, trg_len, seq_len, end_token=1):
    """Batch form with the padded layout of the C-ABI (`md_merge_and_mask`): rows whose merged length exceeds seq_len (the
    ones helper_filter drops) keep their true length and get an all-padding row."""
    B = len(src_len)
    input_ids = np.zeros((B, seq_len), np.int32)
    input_mask = np.ones((B, seq_len), np.int32)
    length = np.zeros((B,), np.int32)
    for b in range(B):
        ids, mask, n = merge_and_mask(src[b, :src_len[b]], trg[b, :trg_len[b]], end_token)
        length[b] = n
        if n <= seq_len:
            input_ids[b, :n] = ids
            input_mask[b, :n] = mask
    return input_ids, input_mask, length


def merge_cases(seed=7, n_rows=120, max_trg=200):
    """(src, trg) pairs in the raw-dataset format (preprocess.py:9-23: src = 11 meta tokens, trg = REMI-like events with
    the (position, chord) pairs still inline), plus rows that hit the numpy corner cases."""
    rng = np.random.default_rng(seed)
    srcs, trgs = [], []

    def meta():
        return [int(rng.integers(lo, hi + 1)) for lo, hi in
                [(560, 600), (601, 625), (626, 629), (630, 637), (638, 640), (641, 649), (650, 652), (653, 718), (653, 718),
                 (719, 725), (726, 728)]]

    def note():
        return [int(rng.integers(432, 560)), int(rng.integers(131, 195)), int(rng.integers(3, 131)), int(rng.integers(304, 432))]

    for _ in range(n_rows):
        t = []
        for _ in range(int(rng.integers(1, 9))):
            t += [2, 432, int(rng.integers(195, 304))]
            for _ in range(int(rng.integers(0, 5))):
                t += note()
                if rng.random() < 0.15:
                    t += [432 + 16 * int(rng.integers(1, 8)), int(rng.integers(195, 304))]
        t = t[:max_trg - 1] + [1]
        srcs.append(meta()); trgs.append(t)
    srcs.append(meta()); trgs.append([200, 2, 440, 150, 60, 310, 1])                 # chord token at index 0: pairs with trg[-1]
    srcs.append(meta()); trgs.append([2, 432, 200, 201, 202, 440, 150, 60, 310, 1])   # adjacent chord tokens: duplicates in src
    srcs.append(meta()); trgs.append([2, 440, 150, 60, 310, 1])                       # no chord at all
    srcs.append(meta()); trgs.append([1])
    srcs.append(meta()); trgs.append([250])                                           # single chord token: index 0 and -1 coincide
    srcs.append([]); trgs.append([2, 432, 200, 1])                                     # empty src
    Ls, Lt = max(len(s) for s in srcs), max(len(t) for t in trgs)
    src = np.zeros((len(srcs), max(Ls, 1)), np.int64); trg = np.zeros((len(trgs), Lt), np.int64)
    for b, (s, t) in enumerate(zip(srcs, trgs)):
        src[b, :len(s)] = s; trg[b, :len(t)] = t
    return src, np.array([len(s) for s in srcs], np.int64), trg, np.array([len(t) for t in trgs], np.int64)


def well_formed_rows(seed, n_rows, max_bars=6, max_notes=5):
    """REMI-like targets as the dataset has them: per bar a (432, chord) pair, notes in position order, optional chord
    changes (432 + 16k, chord) placed in front of the first note at or after that position."""
    rng = np.random.default_rng(seed)
    rows = []
    for _ in range(n_rows):
        meta = [570, 610, 627, 631, 639, 642, 651, 660, 700, 720, 727]
        t = []
        for _ in range(int(rng.integers(1, max_bars + 1))):
            t += [2, 432, int(rng.integers(195, 304))]
            pos = np.sort(rng.integers(432, 560, size=int(rng.integers(0, max_notes + 1))))
            changes = [432 + 16 * int(k) for k in np.sort(rng.choice(np.arange(1, 8), size=int(rng.integers(0, 3)), replace=False))]
            for p_ in pos:
                while changes and changes[0] <= p_:
                    t += [changes.pop(0), int(rng.integers(195, 304))]
                t += [int(p_), int(rng.integers(131, 195)), int(rng.integers(3, 131)), int(rng.integers(304, 432))]
            for c in changes:
                t += [c, int(rng.integers(195, 304))]
        rows.append((meta, t + [1]))
    return rows
