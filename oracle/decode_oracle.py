"""
ORACLE — TEST INFRASTRUCTURE ONLY.  Never imported by the product (`musediffusion_b200/`).

CPU restatement (plain Python over numpy int arrays) of the token-level half of the reference's post-sampling decode,
SURVEY.md §8(f) row 1: everything `SequenceToMidi.decode` does BEFORE it hands the sequence to the MIDI writer
(MuseDiffusion/utils/decode_util.py:207-214):

    split_meta_midi  (decode_util.py:192-199)   meta / note split from the mask, then
    remove_padding   (:72-82)                   cut after the first EOS (token 1)           -> "NO EOS TOKEN"
    restore_chord    (:84-141)                  re-insert the (position, chord) pairs of the meta  -> "RESTORE_CHORD FROM META FAILED"
    validate_once    (:143-155)                 at least one (position, velocity, pitch, duration) 4-gram
    validate_rigidly (:157-184)                 strict grammar walk (only when strict_validation)

The reference signals failures with `SequenceToMidiError(msg)`; any other exception (IndexError from an out-of-range
look-ahead, ...) is re-raised by `batch_decode_*` and aborts the run (decode_util.py:283-300).  Both are kept here as
status codes so that the CUDA kernel can be compared bit for bit:

    0 OK   1 NO_EOS   2 RESTORE_FAILED   3 VALIDATION_FAILED   4 STRICT_FAILED   5 INDEX_ERROR (reference would crash)
    6 TOO_LONG (batch layout only: restored sequence longer than the 2L columns of the padded output)

Pinned by tests/golden/decode_prepare.npz (generated from the unmodified reference by oracle/make_golden.py).
"""
import numpy as np

OK, NO_EOS, RESTORE_FAILED, VALIDATION_FAILED, STRICT_FAILED, INDEX_ERROR, TOO_LONG = 0, 1, 2, 3, 4, 5, 6
STATUS_TEXT = {OK: "OK", NO_EOS: "NO EOS TOKEN", RESTORE_FAILED: "RESTORE_CHORD FROM META FAILED",
               VALIDATION_FAILED: "VALIDATION OF SEQUENCE FAILED", STRICT_FAILED: "STRICT VALIDATION OF SEQUENCE FAILED",
               INDEX_ERROR: "IndexError (the reference aborts)"}

# commu/preprocessor/encoder/event_tokens.py:308-329 (TOKEN_OFFSET)
EOS, BAR, PITCH, NOTE_VELOCITY, CHORD_START, NOTE_DURATION, POSITION, BPM = 1, 2, 3, 131, 195, 304, 432, 560


class _Fail(Exception):
    def __init__(self, code):
        self.code = code


def _at(seq, i):
    """numpy indexing semantics of the reference: negative indices wrap, out of range raises (-> INDEX_ERROR)."""
    n = len(seq)
    if i < -n or i >= n:
        raise _Fail(INDEX_ERROR)
    return int(seq[i])


def remove_padding(seq):
    """decode_util.py:72-82"""
    hits = np.nonzero(np.asarray(seq) == EOS)[0]
    if len(hits) == 0:
        raise _Fail(NO_EOS)
    return np.asarray(seq)[:int(hits[0]) + 1]


def restore_chord(seq, meta):
    """decode_util.py:84-141.  `meta` = 11 meta tokens followed by (position, chord) pairs."""
    seq = np.asarray(seq)
    new_meta = np.asarray(meta)[:11]
    chord = np.asarray(meta)[11:]
    n_chord_bars = int(np.sum(chord == POSITION))
    bars = np.nonzero(seq == BAR)[0]
    if len(bars) == n_chord_bars:
        first = 0
    elif len(bars) == n_chord_bars + 1:
        first = 1
    elif len(bars) < n_chord_bars:
        for _ in range(n_chord_bars - len(bars)):        # np.insert(seq, -1, 2): a BAR in front of the last token
            seq = np.concatenate([seq[:-1], [BAR], seq[-1:]]) if len(seq) else _raise(INDEX_ERROR)
        bars = np.nonzero(seq == BAR)[0]
        first = 0
    else:
        raise _Fail(RESTORE_FAILED)
    if first >= len(bars):
        raise _Fail(INDEX_ERROR)
    out = list(seq[:int(bars[first]) + 1]) + list(chord[:2])
    bar_count = first
    last = int(bars[first])
    for i in range(2, len(chord), 2):
        if int(chord[i]) == POSITION:
            if bar_count + 1 >= len(bars):
                raise _Fail(INDEX_ERROR)
            out += list(seq[last + 1:int(bars[bar_count + 1]) + 1]) + list(chord[i:i + 2])
            bar_count += 1
            last = int(bars[bar_count])
        else:
            cand = np.nonzero((seq >= POSITION) & (seq < int(chord[i])))[0]
            if bar_count != len(bars) - 1:
                ok = cand[(cand > bars[bar_count]) & (cand < bars[bar_count + 1])]
            else:
                ok = cand[cand > bars[bar_count]]
            if len(ok) == 0:
                out += list(chord[i:i + 2])
            else:
                c = int(ok[-1])
                out += list(seq[last + 1:c + 4]) + list(chord[i:i + 2])
                last = c + 3
    out += list(seq[last + 1:])
    return np.asarray(out, dtype=np.int64), new_meta


def _raise(code):
    raise _Fail(code)


def validate_once(seq):
    """decode_util.py:143-155 (note seq[idx - 1] at idx = 0 reads the LAST token, as numpy does)."""
    n = len(seq)
    for idx in range(n):
        if idx + 2 > n - 1:
            break
        if (NOTE_VELOCITY <= _at(seq, idx) < CHORD_START and POSITION <= _at(seq, idx - 1) < BPM
                and PITCH <= _at(seq, idx + 1) < NOTE_VELOCITY and NOTE_DURATION <= _at(seq, idx + 2) < POSITION):
            return
    raise _Fail(VALIDATION_FAILED)


def validate_rigidly(seq):
    """decode_util.py:157-184 (the eager `all([...])` reads seq[i + 3] even when seq[i + 2] already fails)."""
    i, n = 0, len(seq)
    while True:
        if i >= n:
            break
        t = _at(seq, i)
        if t == EOS:
            return
        if t == BAR:
            i += 1
            continue
        if not (POSITION <= t < BPM):
            break
        t1 = _at(seq, i + 1)
        if NOTE_VELOCITY <= t1 < CHORD_START:
            t2, t3 = _at(seq, i + 2), _at(seq, i + 3)
            if PITCH <= t2 < NOTE_VELOCITY and NOTE_DURATION <= t3 < POSITION:
                i += 4
                continue
            break
        if CHORD_START <= t1 < NOTE_DURATION:
            i += 2
            continue
        break
    raise _Fail(STRICT_FAILED)


def decode_prepare(seq, input_mask, strict=False):
    """split_meta_midi (decode_util.py:192-199) + validate_generated_sequence (:186-190) for ONE sequence.
    Returns (status, note_seq, meta); note_seq / meta are the arrays the MIDI writer would get (empty on failure
    before they exist)."""
    seq = np.asarray(seq)
    empty = np.zeros((0,), dtype=np.int64)
    try:
        len_meta = len(seq) - int(np.asarray(input_mask).sum())
        meta = seq[:len_meta - 1] if len_meta - 1 >= 0 else seq[:len_meta - 1]      # python slice semantics
        notes = remove_padding(seq[len_meta:])
        notes, meta11 = restore_chord(notes, meta)
    except _Fail as f:
        return f.code, empty, empty
    try:
        validate_once(notes)
        if strict:
            validate_rigidly(notes)
    except _Fail as f:
        return f.code, notes, np.asarray(meta11, dtype=np.int64)
    return OK, notes, np.asarray(meta11, dtype=np.int64)


def decode_prepare_batch(tokens, masks, strict=False):
    """Batch form with the padded layout of the C-ABI (`md_decode_prepare`): status [B] int32, note_len [B] int32,
    notes [B, 2L] int32 (zero padded), meta [B, 11] int32."""
    tokens, masks = np.asarray(tokens), np.asarray(masks)
    B, L = tokens.shape
    status = np.zeros((B,), np.int32)
    note_len = np.zeros((B,), np.int32)
    notes = np.zeros((B, 2 * L), np.int32)
    meta = np.zeros((B, 11), np.int32)
    for b in range(B):
        st, ns, mt = decode_prepare(tokens[b], masks[b], strict)
        if len(ns) > 2 * L:
            st, ns, mt = TOO_LONG, ns[:0], mt[:0]
        status[b] = st
        note_len[b] = len(ns)
        notes[b, :len(ns)] = ns
        meta[b, :len(mt)] = mt[:11]
    return status, note_len, notes, meta


def decode_cases(seed=2024, L=160, n_rand=160):
    """Token rows + masks exercising every branch of split_meta_midi / restore_chord / validate_* (decode_util.py:72-199):
    well-formed rows with fewer / equal / one more / many more BARs than chord bars, intra-bar chord changes, rows
    without EOS, truncated notes, random substitutions.  Shared by the golden script and the GPU parity tests."""
    rng = np.random.default_rng(seed)
    rows, masks = [], []

    def note():
        return [int(rng.integers(432, 560)), int(rng.integers(131, 195)), int(rng.integers(3, 131)), int(rng.integers(304, 432))]

    def build(n_chord_bars, n_bars, notes_per_bar, changes=0.3, tail=(1,)):
        meta = [int(rng.integers(lo, hi + 1)) for lo, hi in
                [(560, 600), (601, 625), (626, 629), (630, 637), (638, 640), (641, 649), (650, 652), (653, 718), (653, 718),
                 (719, 725), (726, 728)]]
        chord = []
        for _ in range(n_chord_bars):
            chord += [432, int(rng.integers(195, 304))]
            if rng.random() < changes:
                chord += [432 + 16 * int(rng.integers(1, 8)), int(rng.integers(195, 304))]
        body = []
        for _ in range(n_bars):
            body.append(2)
            for _ in range(int(rng.integers(0, notes_per_bar + 1))):
                body += note()
        row = meta + chord + [1] + body + list(tail)
        row = row[:L]
        m = [0] * min(len(meta) + len(chord) + 1, L) + [1] * (L - min(len(meta) + len(chord) + 1, L))
        row = row + [0] * (L - len(row))
        return row, m

    for nb in (1, 2, 4, 8):
        for delta in (-2, -1, 0, 1, 2, 3):
            if nb + delta < 0:
                continue
            rows_m = build(nb, nb + delta, 3)
            rows.append(rows_m[0]); masks.append(rows_m[1])
    r, m = build(4, 4, 3, tail=())                      # no EOS after the notes
    rows.append(r); masks.append(m)
    r, m = build(0, 0, 0)                               # no chords, no bars: bar_idx[0] raises in the reference
    rows.append(r); masks.append(m)
    r, m = build(2, 2, 2, tail=(int(rng.integers(432, 560)), int(rng.integers(131, 195)), 1))   # "... pos vel EOS"
    rows.append(r); masks.append(m)
    r, m = build(2, 2, 2, tail=(int(rng.integers(432, 560)), 1))                                 # "... pos EOS"
    rows.append(r); masks.append(m)
    r, m = build(2, 2, 0)                               # bars but not a single note
    rows.append(r); masks.append(m)
    for _ in range(n_rand):                             # random rows, then random damage
        nb = int(rng.choice([1, 2, 3, 4, 6]))
        r, m = build(nb, max(nb + int(rng.integers(-1, 2)), 0), int(rng.integers(1, 4)), changes=float(rng.random()) * 0.8)
        r = np.array(r)
        start = int(np.sum(np.array(m) == 0))
        kind = int(rng.integers(0, 6))
        if kind == 1:                                   # substitute a few note tokens
            for _ in range(int(rng.integers(1, 6))):
                r[int(rng.integers(start, L))] = int(rng.integers(0, 729))
        elif kind == 2:                                 # drop every EOS in the note part
            r[start:][r[start:] == 1] = int(rng.integers(3, 131))
        elif kind == 3:                                 # early EOS
            r[int(rng.integers(start, min(start + 12, L)))] = 1
        elif kind == 4:                                 # fully random note part (what an untrained model emits)
            r[start:] = rng.integers(0, 729, size=L - start)
        rows.append(r.tolist()); masks.append(m)
    return np.array(rows, dtype=np.int64), np.array(masks, dtype=np.int64)


def too_long_case(L=256, pairs=24):
    """A row whose meta alternates late / early intra-bar chord positions: restore_chord's `last_idx` jumps backwards
    and the same notes are copied again and again (decode_util.py:118-138) until the result exceeds 2L tokens."""
    meta = [570, 610, 627, 631, 639, 642, 651, 660, 700, 720, 727]
    chord = [432, 200]
    for _ in range(pairs):
        chord += [432 + 16 * 7, 201, 432 + 16 * 1, 202]
    n_notes = (L - len(meta) - len(chord) - 4) // 4
    body = [2, 433, 150, 60, 310]
    for _ in range(n_notes - 1):
        body += [500, 150, 60, 310]
    row = meta + chord + [1] + body + [1]
    row = row + [0] * (L - len(row))
    m = len(meta) + len(chord) + 1
    mask = [0] * m + [1] * (L - m)
    return np.array([row], dtype=np.int64), np.array([mask], dtype=np.int64)
