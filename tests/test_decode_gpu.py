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
    checkpoint + training_args.json in, tokens.npy + decode_status.npy + generated_*.midi out, summaries worded like the reference."""
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
    assert "Generation Failure" not in printed            # generation skips failures quietly (decode_util.py:357-359)
    assert printed.count("Summary of Trial") == 2
    assert "(%d valid)" % int((status == 0).sum()) in printed
    assert sorted(f for f in os.listdir(d) if f.endswith(".midi")) == ["generated_%07d.midi" % k for k in range(int((status == 0).sum()))]
    # the reference's meta flags (config/sample.py:157-254) instead of the synthetic prefix
    from musediffusion_b200 import meta
    flags = {"bpm": 70, "audio_key": "aminor", "time_signature": "4/4", "pitch_range": "mid_high", "num_measures": 8,
             "inst": "acoustic_piano", "genre": "newage", "min_velocity": 60, "max_velocity": 80, "track_role": "main_melody",
             "rhythm": "standard", "chord_progression": "Am-Am-Am-Am-G-G-G-G-F-F-F-F-E-E-E-E"}
    argv = []
    for k, v in flags.items():
        argv += ["--" + k, str(v)]
    out2 = tmp_path / "out2"
    sample.main(["generation", "--model_path", str(ck / "model_000001.pt"), "--step", "20", "--batch_size", "2",
                 "--num_samples", "3", "--out_dir", str(out2)] + argv)
    tokens = np.load(out2 / "run1" / "model_000001.pt.generation.samples" / "tokens.npy")
    prefix = meta.meta_to_sequence(flags)
    assert tokens.shape == (3, 64) and (tokens[:, :len(prefix)] == np.array(prefix)).all()


@pytest.mark.parametrize("mode", ["modification", "generation"])
def test_decode_batch_writes_the_midi_files_of_the_reference(tmp_path, golden_dir, capsys, mode):
    """decode_batch (utils/decode_util.py:233-384): rows of the decode fixture -> files named / numbered as the reference's
    two writers do, each holding the notes the reference's write_midi produced (tests/golden/midi_decode.json)."""
    import json
    from musediffusion_b200 import midi
    g = np.load(os.path.join(golden_dir, "decode_prepare.npz"))
    gold = {r["row"]: r for r in json.load(open(os.path.join(golden_dir, "midi_decode.json")))["rows"]}
    good = [b for b, r in gold.items() if "error" not in r]
    bad = [b for b in range(len(g["status_0"])) if g["status_0"][b] in (1, 2, 3)][:7]
    rows = sorted(good[:20] + bad)
    valid, invalid = decode_util.decode_batch(mode, g["tokens"][rows], g["masks"][rows], batch_index=3, previous_count=60,
                                              output_dir=str(tmp_path), return_indices=True)
    assert valid == 20 and invalid == [k for k, b in enumerate(rows) if b in bad]
    printed = capsys.readouterr().out
    k_valid = 0
    for k, b in enumerate(rows):
        if b in bad:
            continue
        name = "%07d_batch%05d_%04d.midi" % (60 + k, 3, k) if mode == "modification" else "generated_%07d.midi" % (60 + k_valid)
        k_valid += 1
        tpb, (conductor, track) = midi.read_smf((tmp_path / name).read_bytes())
        ons = sorted((t, raw[0], raw[1]) for t, st, raw in track if st == 0x90)
        assert ons == sorted((s, p, v) for v, p, s, e in gold[b]["notes"])
        assert [[raw.decode(), t] for t, kind, raw in conductor if kind == (0xFF, 0x06)] == sorted(gold[b]["markers"], key=lambda m: m[1])
    assert len(os.listdir(tmp_path)) == 20
    if mode == "modification":
        assert printed.count("Generation Failure") == len(bad) and "Summary of Batch 3" in printed and " * 7 sequences are invalid." in printed
    else:
        assert "Generation Failure" not in printed and "Summary of Trial 3" in printed and " * Totally 80 sequences are converted." in printed
    raising = [b for b, r in gold.items() if "error" in r][:1]
    with pytest.raises(KeyError):                         # an "unknown" key / time-signature token aborts the reference too
        decode_util.decode_batch(mode, g["tokens"][raising], g["masks"][raising], 0, 0, str(tmp_path))
