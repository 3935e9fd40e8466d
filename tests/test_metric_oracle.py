"""CPU tests: oracle/metric_oracle.py against tests/golden/metrics.npz, produced by the
UNMODIFIED reference's MuseDiffusion/metric.py (get_vectors, ONNC, Controllability_Pitch / _Velocity)."""
import os

import numpy as np
import pytest

import metric_oracle as M


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(os.path.join(golden_dir, "metrics.npz"), allow_pickle=False)


def rows_of(g):
    return [g["midis"][b, :g["lens"][b]] for b in range(len(g["lens"]))]


def test_cases_are_the_committed_ones(g):
    metas, midis, lens = M.metric_cases()
    assert np.array_equal(metas, g["metas"]) and np.array_equal(midis, g["midis"]) and np.array_equal(lens, g["lens"])


def test_vectors_match_reference(g):
    out = [M.get_vectors(r) for r in rows_of(g)]
    assert all(o[0] == 0 for o in out)
    for k, name in ((1, "rhythm"), (2, "melody"), (3, "harmony")):
        got = np.stack([o[k] for o in out])
        assert np.abs(got - g[name]).max() < 1e-6, name          # float32 norms: torch.norm's summation order differs


def test_onnc_and_controllability_match_reference(g):
    score, most, msim = M.onnc(g["rhythm"], g["melody"], g["harmony"])
    assert np.array_equal(most, g["most_sim"])
    assert abs(score - float(g["onnc"])) < 1e-6
    assert np.abs(msim - g["msim"]).max() < 1e-6
    assert M.controllability_pitch(g["metas"], rows_of(g)) == tuple(g["cp"].tolist())
    assert M.controllability_velocity(g["metas"], rows_of(g)) == tuple(g["cv"].tolist())


def test_sequences_the_reference_raises_on():
    assert M.get_vectors([440, 150, 60, 310, 1])[0] == 1                 # no BAR: the first scan runs off the end
    assert M.get_vectors([2, 440, 150, 60, 1])[0] == 1                   # note cut short ("wrong format")
    assert M.get_vectors([2, 150, 60, 310, 1])[0] == 1                   # "position not found"
    assert M.get_vectors([2, 1])[0] == 1                                 # no note at all: unbound `startp`
    assert M.get_vectors([2, 440, 150, 60, 310])[0] == 1                 # no EOS: runs off the end
    assert M.get_vectors([2, 440, 150, 60, 310, 1])[0] == 0
