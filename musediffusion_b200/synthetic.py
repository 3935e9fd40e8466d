"""Synthetic ComMU-shaped token batches (no dataset download is possible offline): the input formats of
`meta_to_batch` (MuseDiffusion/utils/decode_util.py:221-230, generation) and `collate_batches`
(MuseDiffusion/data/wrapper.py:90-126 with the masks of data/preprocess.py:50-56, modification)."""
import numpy as np

# one token per meta field, ranges from commu/preprocessor/encoder/event_tokens.py:308-329 (TOKEN_OFFSET)
META_RANGES = [(560, 600), (601, 625), (626, 629), (630, 637), (638, 640), (641, 649), (650, 652), (653, 718),
               (653, 718), (719, 725), (726, 728)]
CHORD_MARK, CHORD_LO, CHORD_HI = 432, 195, 303


def make_prefix(rng):
    """11 meta tokens then (position, chord) pairs per bar as MetaToSequence builds them (decode_util.py:25-46)."""
    toks = [int(rng.integers(lo, hi + 1)) for lo, hi in META_RANGES]
    for _ in range(int(rng.choice([4, 8, 16]))):
        toks += [CHORD_MARK, int(rng.integers(CHORD_LO, CHORD_HI + 1))]
        if rng.random() < 0.25:
            toks += [CHORD_MARK + 16 * int(rng.integers(1, 8)), int(rng.integers(CHORD_LO, CHORD_HI + 1))]
    return toks


def make_synthetic_batch(mode, B, L, seed=105, per_row_prefix=False):
    rng = np.random.default_rng(seed)
    if mode == "generation":
        ids = np.zeros((B, L), dtype=np.int32)
        msk = np.ones((B, L), dtype=np.int32)
        prefix = make_prefix(rng)
        for b in range(B):
            if per_row_prefix and b:
                prefix = make_prefix(rng)
            n = min(len(prefix), L - 1)
            ids[b, :n] = prefix[:n]
            msk[b, :n + 1] = 0
        return {"input_ids": ids, "input_mask": msk}
    if mode == "modification":
        ids = np.zeros((B, L), dtype=np.int64)
        msk = np.ones((B, L), dtype=np.int64)
        length = np.zeros((B,), dtype=np.int64)
        for b in range(B):
            prefix = make_prefix(rng)
            n = min(len(prefix), max(L // 2 - 1, 1))
            total = int(rng.integers(min(64, L), L + 1))
            row = list(prefix[:n]) + [1]
            k = 0
            while len(row) < total - 1:
                if k % 6 == 0:
                    row.append(2)
                row += [int(rng.integers(432, 560)), int(rng.integers(131, 195)), int(rng.integers(3, 131)),
                        int(rng.integers(304, 432))]
                k += 1
            row = row[:total - 1] + [1]
            ids[b, :len(row)] = row
            msk[b, :n + 1] = 0
            length[b] = len(row)
        return {"input_ids": ids, "input_mask": msk, "length": length}
    raise ValueError(mode)
