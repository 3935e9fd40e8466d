"""Generation-mode input format: meta dict -> token prefix -> batch, against vectors written by the
reference's own MetaToSequence / MetaEncoder (oracle/make_golden.py `meta`)."""
import argparse
import json
import os

import numpy as np
import pytest

from musediffusion_b200 import meta

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def gold():
    return json.load(open(os.path.join(GOLD, "meta_encode.json")))


def test_meta_to_sequence_matches_reference(gold):
    assert len(gold["cases"]) >= 50
    for case in gold["cases"]:
        assert meta.meta_to_sequence(case["meta"]) == case["tokens"], case["meta"]


def test_unknown_tokens_and_errors(gold):
    base = {k: v for k, v in gold["cases"][0]["meta"].items() if k != "chord_progression"}
    assert len(gold["unknown"]) == 10                     # every field but num_measures has an "unknown" token
    for u in gold["unknown"]:
        d = dict(base)
        d[u["field"]] = "unknown"
        assert meta.encode_meta(d) == u["tokens"], u["field"]
    for e in gold["errors"]:
        d = dict(base)
        d[e["field"]] = e["value"]
        with pytest.raises(meta.UnprocessableMidiError):
            meta.encode_meta(d)


def test_tables_cover_the_vocabulary():
    assert len(meta.CHORD_MAP) == 109 and max(meta.CHORD_MAP.values()) == 303 and meta.CHORD_MAP["NN"] == 303
    assert len(meta.KEY_MAP) == 34 and len(meta.INST_MAP) == 62
    assert meta.OFFSET["rhythm"] + 1 + max(meta.RHYTHM_MAP.values()) == 728      # last id of the 729-token vocabulary


def test_meta_to_batch_matches_fixture():
    ref = np.load(os.path.join(GOLD, "meta_batch.npz"))
    m = {"bpm": 70, "audio_key": "aminor", "time_signature": "4/4", "pitch_range": "mid_high", "num_measures": 8,
         "inst": "acoustic_piano", "genre": "newage", "min_velocity": 60, "max_velocity": 80, "track_role": "main_melody",
         "rhythm": "standard", "chord_progression": "-".join((["Am"] * 8 + ["G"] * 8 + ["F"] * 8 + ["E"] * 8) * 2)}
    b = meta.meta_to_batch(m, 2, 64)
    assert b["input_ids"].dtype == b["input_mask"].dtype and str(b["input_ids"].dtype) == "torch.int32"
    assert np.array_equal(b["input_ids"].numpy(), ref["input_ids"])
    assert np.array_equal(b["input_mask"].numpy(), ref["input_mask"])
    with pytest.raises(RuntimeError):
        meta.meta_to_batch(m, 2, 27)                      # prefix + separator slot does not fit
    with pytest.raises(AssertionError):
        meta.encode_chord(["Am"] * 7)


def test_cli_flags_round_trip(tmp_path, gold):
    p = meta.add_meta_arguments(argparse.ArgumentParser())
    assert meta.meta_from_args(p.parse_args([])) is None
    case = gold["cases"][3]["meta"]
    argv = []
    for k, v in case.items():
        argv += ["--" + k, str(v)]
    got = meta.meta_from_args(p.parse_args(argv))
    assert meta.meta_to_sequence(got) == gold["cases"][3]["tokens"]
    # list-literal chord syntax (config/sample.py:173-177) and the json route
    listed = dict(case, chord_progression=str(case["chord_progression"].split("-")))
    jf = tmp_path / "meta.json"
    jf.write_text(json.dumps(listed))
    got = meta.meta_from_args(p.parse_args(["--meta_json", str(jf)]))
    assert meta.meta_to_sequence(got) == gold["cases"][3]["tokens"]
    with pytest.raises(ValueError):
        meta.meta_from_args(p.parse_args(["--bpm", "70"]))
