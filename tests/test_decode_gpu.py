"""GPU parity tests of md_decode_prepare through the C-ABI: bit-exact against the fixture
generated from the unmodified reference and against oracle/decode_oracle.py on fresh rows (integer work)."""
import os

import numpy as np
import pytest
import torch

import decode_oracle as D

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():  # pragma: no cover
    pytest.skip("needs a CUDA device", allow_module_level=True)

from musediffusion_b200 import _lib, decode_util, ops  # noqa: E402

DEV = torch.device("cuda:0")


def run_kernel(tokens, masks, strict):
    out = ops.decode_prepare(torch.as_tensor(tokens).to(DEV), torch.as_tensor(masks).to(DEV), bool(strict))
    torch.cuda.synchronize()
    return [t.cpu().numpy() for t in out]


@pytest.mark.parametrize("strict", [0, 1])
def test_matches_reference_fixture(golden_dir, strict):
    g = np.load(os.path.join(golden_dir, "decode_prepare.npz"), allow_pickle=False)
    status, note_len, notes, meta = run_kernel(g["tokens"], g["masks"], strict)
    assert np.array_equal(status, g["status_%d" % strict])
    assert np.array_equal(note_len, g["note_len_%d" % strict])
    assert np.array_equal(notes, g["notes_%d" % strict])
    assert np.array_equal(meta, g["meta_%d" % strict])


@pytest.mark.parametrize("seed,L,n", [(1, 64, 200), (2, 333, 300), (3, 2096, 96)])
def test_matches_oracle_on_fresh_rows(seed, L, n):
    tokens, masks = D.decode_cases(seed=seed, L=L, n_rand=n)
    for strict in (False, True):
        want = D.decode_prepare_batch(tokens, masks, strict)
        got = run_kernel(tokens, masks, strict)
        for w, k, name in zip(want, got, ["status", "note_len", "notes", "meta"]):
            assert np.array_equal(w, k), (name, strict)


def test_full_length_well_formed_rows_survive():
    """BASELINE-size rows (L = 2096) built like the modification-mode data: every row must come back OK under the
    strict grammar, with the chord pairs spliced in (restored length = notes + chord tokens)."""
    from musediffusion_b200.synthetic import make_synthetic_batch
    cond = make_synthetic_batch("modification", 64, 2096, seed=5)
    tokens, masks = cond["input_ids"], cond["input_mask"]
    want = D.decode_prepare_batch(tokens, masks, True)
    got = run_kernel(tokens, masks, True)
    for w, k in zip(want, got):
        assert np.array_equal(w, k)
    assert set(np.unique(got[0]).tolist()) <= {D.OK, D.RESTORE_FAILED, D.VALIDATION_FAILED, D.STRICT_FAILED}


def test_degenerate_meta_overflow_is_reported():
    tok, msk = D.too_long_case()
    want = D.decode_prepare_batch(tok, msk, False)
    got = run_kernel(tok, msk, False)
    assert want[0][0] == D.TOO_LONG
    for w, k in zip(want, got):
        assert np.array_equal(w, k)


def test_host_mirror_reports_like_decode_batch(golden_dir):
    g = np.load(os.path.join(golden_dir, "decode_prepare.npz"), allow_pickle=False)
    prep = decode_util.prepare_batch(g["tokens"], g["masks"], strict_validation=False)
    ref_status = g["status_0"]
    assert prep.valid_count == int((ref_status == 0).sum())
    assert prep.invalid_idxes == np.nonzero(ref_status != 0)[0].tolist()
    b = int(np.nonzero(ref_status == 0)[0][0])
    assert np.array_equal(prep.note_seqs[b], g["notes_0"][b, :g["note_len_0"][b]])
    assert np.array_equal(prep.metas[b], g["meta_0"][b])
    lines = []
    ok_only = decode_util.PreparedBatch(prep.status, prep.note_seqs, prep.metas, prep.valid_count,
                                        [i for i in prep.invalid_idxes if ref_status[i] != D.INDEX_ERROR])
    decode_util.report_failures(ok_only, batch_index=3, previous_count=100, print_fn=lines.append)
    assert len(lines) == len(ok_only.invalid_idxes)
    i0 = ok_only.invalid_idxes[0]
    assert lines[0] == "<Warning> Batch 3 Index %d (Original: %d) - Generation Failure: %s" % (
        i0, 100 + i0, D.STATUS_TEXT[int(ref_status[i0])])
    if (ref_status == D.INDEX_ERROR).any():
        with pytest.raises(IndexError):
            decode_util.report_failures(prep, batch_index=3, previous_count=100, print_fn=lines.append)


def test_rejects_cpu_tensors_and_bad_shapes():
    with pytest.raises(_lib.MuseDiffLibraryError):
        ops.decode_prepare(torch.zeros(2, 8, dtype=torch.int32), torch.zeros(2, 8, dtype=torch.int32))
    with pytest.raises(ValueError):
        decode_util.prepare_batch(np.zeros((2, 8), np.int64), np.zeros((2, 9), np.int64))


def test_cli_generation_end_to_end(tmp_path, capsys):
    """`python -m musediffusion_b200 generation ...` on a small random-init checkpoint (2-layer encoder, seq_len 64):
    checkpoint + training_args.json in, tokens.npy + decode_status.npy out, failure warnings worded like the reference."""
    import json
    from musediffusion_b200 import sample
    from musediffusion_b200.initialization import create_model_and_diffusion
    targs = dict(hidden_dim=128, hidden_t_dim=128, vocab_size=729, seq_len=64, dropout=0.1, noise_schedule="sqrt",
                 diffusion_steps=20, timestep_respacing="", rescale_timesteps=True, predict_xstart=True,
                 encoder_config=dict(num_hidden_layers=2))
    torch.manual_seed(0)
    model, _ = create_model_and_diffusion(**targs)
    ck = tmp_path / "run1"
    ck.mkdir()
    torch.save(model.state_dict(), ck / "model_000001.pt")
    (ck / "training_args.json").write_text(json.dumps(targs))
    out = tmp_path / "out"
    sample.main(["generation", "--model_path", str(ck / "model_000001.pt"), "--step", "20", "--batch_size", "3",
                 "--num_samples", "6", "--out_dir", str(out), "--strict_validation"])
    d = out / "run1" / "model_000001.pt.generation.samples"
    tokens = np.load(d / "tokens.npy")
    status = np.load(d / "decode_status.npy")
    assert tokens.shape == (6, 64) and status.shape == (6,)
    from musediffusion_b200.synthetic import make_synthetic_batch
    b = make_synthetic_batch("generation", 6, 64, seed=105)
    want = D.decode_prepare_batch(tokens, b["input_mask"], True)
    assert np.array_equal(status, want[0])
    printed = capsys.readouterr().out
    assert printed.count("Generation Failure") == int((status != 0).sum())
    assert "(%d valid)" % int((status == 0).sum()) in printed
