"""GPU parity tests of md_sequence_metrics / md_onnc through the C-ABI against the fixture
generated from the unmodified reference (MuseDiffusion/metric.py) and against oracle/metric_oracle.py on fresh rows."""
import os

import numpy as np
import pytest
import torch

import metric_oracle as M

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():  # pragma: no cover
    pytest.skip("needs a CUDA device", allow_module_level=True)

from musediffusion_b200 import metric, ops  # noqa: E402

DEV = torch.device("cuda:0")
TOL = 1e-5        # float32 vectors: the warp-tree norms differ from torch.norm's summation order in the last bits


def rows_of(g):
    return [g["midis"][b, :g["lens"][b]] for b in range(len(g["lens"]))]


def test_vectors_and_counts_match_reference_fixture(golden_dir):
    g = np.load(os.path.join(golden_dir, "metrics.npz"), allow_pickle=False)
    vec, status, stats = ops.sequence_metrics(torch.from_numpy(g["midis"]).to(DEV), torch.from_numpy(g["lens"]).to(DEV),
                                              torch.from_numpy(g["metas"]).to(DEV))
    vec = vec.cpu().numpy()
    assert int(status.max()) == 0
    assert np.abs(vec[:, :32] - g["rhythm"]).max() < TOL
    assert np.abs(vec[:, 32:44] - g["melody"]).max() < TOL
    assert np.abs(vec[:, 44:] - g["harmony"]).max() < TOL
    rows = rows_of(g)
    assert metric.Controllability_Pitch(g["metas"], rows, device=DEV) == tuple(g["cp"].tolist())
    assert metric.Controllability_Velocity(g["metas"], rows, device=DEV) == tuple(g["cv"].tolist())


def test_onnc_matches_reference_fixture(golden_dir):
    g = np.load(os.path.join(golden_dir, "metrics.npz"), allow_pickle=False)
    score, msim, most = metric.ONNC(rows_of(g), return_MSIM=True, return_mostsim=True, device=DEV)
    assert np.array_equal(most.cpu().numpy(), g["most_sim"])
    assert abs(float(score) - float(g["onnc"])) < 1e-6
    assert np.abs(msim.cpu().numpy() - g["msim"]).max() < TOL
    assert abs(float(metric.ONNC(rows_of(g), device=DEV)) - float(g["onnc"])) < 1e-6
    r, m, h = metric.get_vectors(rows_of(g)[3], device=DEV)
    assert np.abs(r.cpu().numpy() - g["rhythm"][3]).max() < TOL and m.shape == (12,) and h.shape == (12,)


@pytest.mark.parametrize("seed,n", [(11, 40), (12, 300)])
def test_matches_oracle_on_fresh_rows(seed, n):
    metas, midis, lens = M.metric_cases(seed=seed, n=n)
    rows = [midis[b, :lens[b]] for b in range(len(lens))]
    want = [M.get_vectors(r) for r in rows]
    vec, status, _ = ops.sequence_metrics(torch.from_numpy(midis).to(DEV), torch.from_numpy(lens).to(DEV),
                                          torch.from_numpy(metas).to(DEV))
    vec = vec.cpu().numpy()
    assert np.array_equal(status.cpu().numpy(), np.array([w[0] for w in want]))
    assert np.abs(vec[:, :32] - np.stack([w[1] for w in want])).max() < TOL
    assert np.abs(vec[:, 32:44] - np.stack([w[2] for w in want])).max() < TOL
    assert np.abs(vec[:, 44:] - np.stack([w[3] for w in want])).max() < TOL
    assert metric.Controllability_Pitch(metas, rows, device=DEV) == M.controllability_pitch(metas, rows)
    assert metric.Controllability_Velocity(metas, rows, device=DEV) == M.controllability_velocity(metas, rows)
    score, most, msim = M.onnc(np.stack([w[1] for w in want]), np.stack([w[2] for w in want]), np.stack([w[3] for w in want]))
    got_score, got_most = metric.ONNC(rows, return_mostsim=True, device=DEV)
    margin_ok = np.sort(msim, axis=1)[:, -1] - np.sort(msim, axis=1)[:, -2] > 1e-5      # rows whose nearest neighbour is not a near tie
    assert np.array_equal(got_most.cpu().numpy()[margin_ok], most[margin_ok])
    assert abs(float(got_score) - score) <= (~margin_ok).sum() / len(rows) + 1e-6


def test_sequences_the_reference_raises_on():
    bad = [[440, 150, 60, 310, 1], [2, 440, 150, 60, 1], [2, 150, 60, 310, 1], [2, 1], [2, 440, 150, 60, 310]]
    good = [2, 440, 150, 60, 310, 1]
    rows = bad + [good]
    Ln = max(len(r) for r in rows)
    arr = np.zeros((len(rows), Ln), np.int64)
    for b, r in enumerate(rows):
        arr[b, :len(r)] = r
    vec, status, _ = ops.sequence_metrics(torch.from_numpy(arr).to(DEV), torch.tensor([len(r) for r in rows]).to(DEV),
                                          torch.zeros(len(rows), 11, dtype=torch.int64).to(DEV))
    assert status.cpu().tolist() == [1, 1, 1, 1, 1, 0]
    assert float(vec[:5].abs().max()) == 0.0
    with pytest.raises(ValueError):
        metric.get_vectors(bad[1], device=DEV)
    with pytest.raises(ValueError):
        metric.ONNC([good, bad[0]], device=DEV)


def test_cli_batch_metrics_accumulate_like_the_reference(golden_dir, capsys):
    """sample.batch_metrics (run/sample.py:244-280): ONNC over ground truth + generated of the VALID rows, CP / CV of the
    generated ones, ONNC weighted by the valid count."""
    from musediffusion_b200 import decode_util, metric, sample
    g = np.load(os.path.join(golden_dir, "decode_prepare.npz"))
    dev = torch.device("cuda:0")
    ok = [b for b in range(len(g["status_0"])) if g["status_0"][b] == 0]
    notes = [g["notes_0"][b, :g["note_len_0"][b]] for b in ok]
    _, status, _ = metric._run(notes, device=dev)
    rows = [b for b, s in zip(ok, status.cpu().numpy()) if s == 0][:10]
    assert len(rows) >= 6
    rows = rows + [b for b in range(len(g["status_0"])) if g["status_0"][b] == 3][:2]        # two invalid rows ride along
    tok, mask = torch.from_numpy(g["tokens"][rows]).to(dev), torch.from_numpy(g["masks"][rows]).to(dev)
    prep = decode_util.prepare_batch(tok, mask)
    total = dict(onnc_sum=0.0, onnc_count=0, total_total_p=0, total_wrong_p=0, total_total_v=0, total_wrong_v=0)
    sample.batch_metrics(total, prep, g["tokens"][rows], mask, 4)
    n = len(rows) - 2
    gen = [prep.note_seqs[k] for k in range(n)]
    metas = [prep.metas[k] for k in range(n)]
    onnc = float(metric.ONNC(tuple(gen) + tuple(gen), device=dev))
    assert total["onnc_count"] == n and abs(total["onnc_sum"] - n * onnc) < 1e-6
    assert (total["total_total_p"], total["total_wrong_p"]) == metric.Controllability_Pitch(metas, gen, device=dev)
    assert (total["total_total_v"], total["total_wrong_v"]) == metric.Controllability_Velocity(metas, gen, device=dev)
    out = capsys.readouterr().out
    assert "Metric of Batch 4" in out and "ONNC: %.6f" % onnc in out
