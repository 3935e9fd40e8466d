"""Note sequence -> MIDI content -> Standard MIDI File (last stage), against what the reference's
own write_midi (commu/preprocessor/encoder/encoder_utils.py:386-497) produced for the same inputs
(tests/golden/midi_decode.json, oracle/make_golden.py `midi`)."""
import json
import os
import sys

import numpy as np
import pytest

from musediffusion_b200 import midi

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle"))


@pytest.fixture(scope="module")
def gold():
    return json.load(open(os.path.join(HERE, "golden", "midi_decode.json")))


def check(got, want):
    assert got.ticks_per_beat == want["ticks_per_beat"] == 480
    assert [[got.tempo, 0]] == want["tempo"]
    assert [[got.numerator, got.denominator, 0]] == want["time_signature"]
    assert [[got.key_name, 0]] == want["key"]
    assert got.notes.tolist() == want["notes"]
    assert [[t, int(k)] for t, k in zip(got.marker_texts, got.marker_times)] == want["markers"]
    assert ["OOV: %d" % w for w in got.oov] == want["oov"]


def test_direct_cases_match_reference(gold):
    assert len(gold["direct"]) >= 40 and sum(len(r["notes"]) for r in gold["direct"]) > 500
    assert {tuple(r["time_signature"][0][:2]) for r in gold["direct"]} == {(4, 4), (3, 4), (6, 8), (12, 8)}
    for want in gold["direct"]:
        check(midi.decode_event_sequence(np.array(want["note_seq"]), np.array(want["meta"])), want)


def test_rows_after_restore_chord_match_reference(gold):
    """the rows of the decode fixture that pass validate_once: restored note sequence + meta from the fixture itself"""
    prep = np.load(os.path.join(HERE, "golden", "decode_prepare.npz"))
    seen = errors = 0
    for want in gold["rows"]:
        b = want["row"]
        assert prep["status_0"][b] == 0
        ns, mt = prep["notes_0"][b, :prep["note_len_0"][b]], prep["meta_0"][b]
        if "error" in want:
            with pytest.raises(KeyError):
                midi.decode_event_sequence(ns, mt)
            assert want["error"] == "KeyError"
            errors += 1
        else:
            check(midi.decode_event_sequence(ns, mt), want)
            seen += 1
    assert seen > 40 and errors > 0


def test_smf_round_trip(tmp_path, gold):
    for k, want in enumerate(gold["direct"][:12]):
        m = midi.decode_event_sequence(np.array(want["note_seq"]), np.array(want["meta"]))
        path = tmp_path / ("case%d.midi" % k)
        m.dump(str(path))
        tpb, (conductor, track) = midi.read_smf(path.read_bytes())
        assert tpb == 480
        meta = {kind[1]: (t, raw) for t, kind, raw in conductor if isinstance(kind, tuple)}
        assert meta[0x58][1][:2] == bytes((m.numerator, {4: 2, 8: 3}[m.denominator]))
        assert int.from_bytes(meta[0x51][1], "big") == round(60_000_000 / m.tempo)
        assert meta[0x59][1][1] == int(m.key_name.endswith("minor")) and -7 <= int.from_bytes(meta[0x59][1][:1], "big", signed=True) <= 7
        assert [(t, raw.decode()) for t, kind, raw in conductor if kind == (0xFF, 0x06)] == \
            sorted(zip(m.marker_times.tolist(), m.marker_texts), key=lambda e: e[0])
        assert conductor[-1][1] == (0xFF, 0x2F) and track[-1][1] == (0xFF, 0x2F)
        assert track[0] == (0, 0xC0, b"\x00")
        ons = sorted((t, raw[0], raw[1]) for t, st, raw in track if st == 0x90)
        offs = sorted((t, raw[0]) for t, st, raw in track if st == 0x80)
        assert ons == sorted((s, p, v) for v, p, s, e in m.notes.tolist())
        assert offs == sorted((e, p) for v, p, s, e in m.notes.tolist())
        assert [t for t, _, _ in track] == sorted(t for t, _, _ in track)          # delta times never negative


def test_key_signature_table():
    sf = {}
    for name in midi.KEY_NAMES:
        m = midi.DecodedMidi(120, 4, 4, name, np.zeros((0, 4), np.int64), np.zeros((0,), np.int64), [])
        _, (conductor, _) = midi.read_smf(m.to_bytes())
        raw = [r for _, k, r in conductor if k == (0xFF, 0x59)][0]
        sf[name] = int.from_bytes(raw[:1], "big", signed=True)
    assert sf["cmajor"] == 0 and sf["aminor"] == 0 and sf["gmajor"] == 1 and sf["eminor"] == 1
    assert sf["fmajor"] == -1 and sf["dminor"] == -1 and sf["ebmajor"] == -3 and sf["cminor"] == -3 and sf["bmajor"] == 5


def test_unknown_tempo_cannot_be_written():
    m = midi.decode_event_sequence(np.array([2, 1]), np.array([560, 602, 627, 631, 638, 642, 651, 654, 654, 720, 727]))
    assert m.tempo == 0
    with pytest.raises(ZeroDivisionError):
        m.to_bytes()
