"""CPU tests: oracle/decode_oracle.py (the token-level half of SequenceToMidi.decode)
against tests/golden/decode_prepare.npz, which oracle/make_golden.py produced by running the UNMODIFIED reference
(`MuseDiffusion/utils/decode_util.py` SequenceToMidi.split_meta_midi + validate_generated_sequence) on the same rows."""
import os

import numpy as np
import pytest

import decode_oracle as D


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(os.path.join(golden_dir, "decode_prepare.npz"), allow_pickle=False)


def test_cases_are_the_committed_ones(g):
    tok, msk = D.decode_cases()
    assert np.array_equal(tok, g["tokens"]) and np.array_equal(msk, g["masks"])


@pytest.mark.parametrize("strict", [0, 1])
def test_oracle_matches_reference_outcomes(g, strict):
    status, note_len, notes, meta = D.decode_prepare_batch(g["tokens"], g["masks"], bool(strict))
    assert np.array_equal(status, g["status_%d" % strict])
    assert np.array_equal(note_len, g["note_len_%d" % strict])
    assert np.array_equal(notes, g["notes_%d" % strict])
    assert np.array_equal(meta, g["meta_%d" % strict])


def test_fixture_covers_every_outcome(g):
    seen = set(np.unique(g["status_1"]).tolist())
    assert {D.OK, D.NO_EOS, D.RESTORE_FAILED, D.VALIDATION_FAILED, D.STRICT_FAILED, D.INDEX_ERROR} <= seen


def test_known_small_cases():
    meta = [570, 610, 627, 631, 639, 642, 651, 660, 700, 720, 727]
    note = [440, 150, 60, 310]
    # one chord bar, one BAR: chord pair goes right behind the BAR
    row = meta + [432, 200] + [1] + [2] + note + [1] + [0] * 10
    mask = [0] * 14 + [1] * (len(row) - 14)
    st, ns, mt = D.decode_prepare(np.array(row), np.array(mask), strict=True)
    assert st == D.OK and ns.tolist() == [2, 432, 200] + note + [1] and mt.tolist() == meta
    # two BARs more than chord bars
    row = meta + [432, 200] + [1] + [2, 2, 2] + note + [1]
    mask = [0] * 14 + [1] * (len(row) - 14)
    assert D.decode_prepare(np.array(row), np.array(mask))[0] == D.RESTORE_FAILED
    # no EOS behind the notes
    row = meta + [432, 200] + [1] + [2] + note
    mask = [0] * 14 + [1] * (len(row) - 14)
    assert D.decode_prepare(np.array(row), np.array(mask))[0] == D.NO_EOS
    # missing BARs are inserted in front of the EOS; no complete note -> validate_once fails
    row = meta + [432, 200, 432, 201] + [1] + [2, 440, 150] + [1]
    mask = [0] * 16 + [1] * (len(row) - 16)
    st, ns, _ = D.decode_prepare(np.array(row), np.array(mask))
    assert st == D.VALIDATION_FAILED and ns.tolist() == [2, 432, 200, 440, 150, 2, 432, 201, 1]
    # "... position velocity EOS": the eager look-ahead of validate_rigidly runs off the end (reference: IndexError)
    row = meta + [432, 200] + [1] + [2] + note + [441, 151] + [1]
    mask = [0] * 14 + [1] * (len(row) - 14)
    assert D.decode_prepare(np.array(row), np.array(mask), strict=True)[0] == D.INDEX_ERROR
    assert D.decode_prepare(np.array(row), np.array(mask), strict=False)[0] == D.OK


def test_degenerate_meta_overflows_the_padded_layout():
    tok, msk = D.too_long_case()
    st, ns, _ = D.decode_prepare(tok[0], msk[0])
    assert st == D.OK and len(ns) > 2 * tok.shape[1]           # the reference itself has no bound
    status, note_len, notes, meta = D.decode_prepare_batch(tok, msk)
    assert status[0] == D.TOO_LONG and note_len[0] == 0 and not notes.any() and not meta.any()
