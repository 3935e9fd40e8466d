"""CPU tests: oracle/preprocess_oracle.py against tests/golden/merge_and_mask.npz, produced
by the UNMODIFIED reference (`helper_tokenize`, `helper_filter`, `collate_batches`) on the same raw rows."""
import os

import numpy as np
import pytest

import decode_oracle as D
import preprocess_oracle as P


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(os.path.join(golden_dir, "merge_and_mask.npz"), allow_pickle=False)


def test_cases_are_the_committed_ones(g):
    src, src_len, trg, trg_len = P.merge_cases()
    for a, k in zip((src, src_len, trg, trg_len), ("src", "src_len", "trg", "trg_len")):
        assert np.array_equal(a, g[k]), k


def test_oracle_matches_reference(g):
    L = int(g["seq_len"])
    ids, mask, length = P.merge_and_mask_batch(g["src"], g["src_len"], g["trg"], g["trg_len"], L)
    assert np.array_equal(length, g["length"])
    keep = length <= L
    assert 0 < keep.sum() < len(keep)                       # the fixture has both kept and filtered rows
    assert np.array_equal(ids[keep], g["kept_input_ids"])
    assert np.array_equal(mask[keep], g["kept_input_mask"])
    assert np.array_equal(length[keep], g["kept_length"])
    assert (ids[~keep] == 0).all() and (mask[~keep] == 1).all()


def test_numpy_wrap_corner_cases():
    meta = list(range(560, 571))
    ids, mask, n = P.merge_and_mask(meta, [200, 2, 440, 150, 60, 310, 1])      # chord at index 0 pairs with trg[-1]
    assert ids.tolist() == meta + [1, 200] + [1] + [2, 440, 150, 60, 310] and n == 19
    assert mask.tolist() == [0] * 14 + [1] * 5
    ids, _, _ = P.merge_and_mask(meta, [2, 432, 200, 201, 440, 1])             # adjacent chords: the shared token is duplicated
    assert ids.tolist() == meta + [432, 200, 200, 201] + [1] + [2, 440, 1]


def test_round_trip_with_restore_chord():
    """merge_and_mask moves the (position, chord) pairs into the meta, restore_chord (decode side, row 1) puts them
    back: for well-formed rows decode(encode(trg)) == trg, and the strict grammar accepts the result."""
    n_strict_ok = 0
    for meta, t in P.well_formed_rows(seed=11, n_rows=80):
        ids, mask, n = P.merge_and_mask(meta, t)
        st, notes, m11 = D.decode_prepare(ids, mask, strict=True)
        assert st in (D.OK, D.VALIDATION_FAILED), st          # VALIDATION_FAILED: a row without a single note
        assert notes.tolist() == t
        assert m11.tolist() == meta
        n_strict_ok += st == D.OK
    assert n_strict_ok >= 60
