"""The drop-in, proven through the reference's OWN entry point on a GPU: the unmodified
`MuseDiffusion.run.sample.main()` (run/sample.py:22-311, from the verbatim copy oracle/_ref/ that oracle/build_ref.sh makes)
is run twice on the same reference-written checkpoint directory (`model_000000.pt` + `TrainSettings(...).json()`), the same
inputs and the same noise stream — once with its stock modules on the host CPU, once with the three `sys.modules`
swaps of INTEGRATION.md so that `create_model_and_diffusion`, the loops and the rounding callback resolve to
musediffusion_b200 on cuda:0.  Tokens handed to `decode_batch` must be the reference's under the north-star rule
(oracle/parity_rule.py): ids equal at every rounding call except below-margin positions."""
import os

import numpy as np
import pytest
import torch

import musediff_oracle as O
import parity_rule as R
import ref_harness as H

pytestmark = pytest.mark.gpu
if not torch.cuda.is_available():  # pragma: no cover
    pytest.skip("needs a CUDA device", allow_module_level=True)
if not H.available():  # pragma: no cover
    pytest.skip("reference copy oracle/_ref/ missing: run `sh oracle/build_ref.sh` in the build container", allow_module_level=True)

MARGIN_TOL, DOWNSTREAM_TOL = 0.5, 1.0
L = 64
META = ["--bpm", "70", "--audio_key", "aminor", "--time_signature", "4/4", "--pitch_range", "mid_high", "--num_measures", "8",
        "--inst", "acoustic_piano", "--genre", "newage", "--min_velocity", "60", "--max_velocity", "80", "--track_role",
        "main_melody", "--rhythm", "standard", "--chord_progression", "Am-Am-Am-Am-G-G-G-G-F-F-F-F-E-E-E-E"]


@pytest.fixture(scope="module")
def checkpoint(tmp_path_factory):
    d = tmp_path_factory.mktemp("dropin")
    H.install()
    p = O.make_random_params(seed=7, seq_len=L)
    return H.write_checkpoint_dir(str(d / "diffusion_models"), p, seq_len=L), str(d / "out")


def compare(tag, ref, got):
    assert len(ref["tokens"]) == len(got["tokens"]) >= 1
    n_batches = len(ref["tokens"])
    calls = len(ref["step_ids"]) // n_batches
    assert len(got["step_ids"]) == len(ref["step_ids"]) and calls >= 1
    for b in range(n_batches):
        rt, gt, mask = ref["tokens"][b], got["tokens"][b], ref["masks"][b]
        assert gt.dtype == rt.dtype and gt.shape == rt.shape and np.array_equal(got["masks"][b], mask)
        ref_ids = np.stack(ref["step_ids"][b * calls:(b + 1) * calls])
        ref_margin = np.stack(ref["step_margin"][b * calls:(b + 1) * calls])
        got_ids = np.stack(got["step_ids"][b * calls:(b + 1) * calls])
        free = mask != 0
        rep = R.chain_report(ref_ids, ref_margin, got_ids, free, MARGIN_TOL, DOWNSTREAM_TOL)
        tk = R.token_report(rt, gt, free, rep["never_divergent"])
        print("%s batch %d: %d rounding calls, ids equal at %.5f of (call, position) pairs, %d below-margin positions "
              "(max margin %.3f), tokens equal at %.5f; violations %d + %d"
              % (tag, b, calls, rep["id_agreement_all_calls"], rep["positions_ever_divergent"],
                 max(rep["max_margin_primary"], rep["max_margin_downstream"]), tk["token_agreement"], rep["violations"],
                 tk["token_violations"]))
        assert rep["violations"] == 0 and tk["token_violations"] == 0, (rep, tk)
        assert np.array_equal(gt[~free & (rt > 0)], rt[~free & (rt > 0)])        # the conditioning prefix comes back untouched


def test_reference_main_generation_dropin(checkpoint):
    model_path, out_dir = checkpoint
    argv = ["--step", "20", "--batch_size", "2", "--num_samples", "2"] + META          # DDIM, gap 100
    ref = H.run_main("generation", model_path, out_dir, argv, stream_seed=5)
    got = H.run_main("generation", model_path, out_dir, argv, dropin=True, stream_seed=5)
    compare("generation / ddim20", ref, got)


def test_reference_main_modification_dropin(checkpoint):
    model_path, out_dir = checkpoint
    conds = [O.make_synthetic_batch("modification", 3, L, seed=3 + i) for i in range(2)]
    batches = [{k: torch.from_numpy(v) for k, v in c.items()} for c in conds]
    argv = ["--step", "2000", "--batch_size", "3", "--strength", "0.006", "--use_corruption", "false"]   # DDPM, t_enc = 12, top_p = 1
    ref = H.run_main("modification", model_path, out_dir, argv, batches=batches, stream_seed=6)
    got = H.run_main("modification", model_path, out_dir, argv, batches=batches, dropin=True, stream_seed=6)
    compare("modification / ddpm x12", ref, got)
    argv = ["--step", "50", "--batch_size", "3", "--strength", "0.5", "--use_corruption", "false"]       # DDIM gap 40, t_enc = 25
    ref = H.run_main("modification", model_path, out_dir, argv, batches=batches, stream_seed=8)
    got = H.run_main("modification", model_path, out_dir, argv, batches=batches, dropin=True, stream_seed=8)
    compare("modification / ddim50 x25", ref, got)


def test_reference_main_ends_in_midi_files(checkpoint, capsys):
    """fourth swap: `decode_batch` of the reference's main() resolves to the package's (GPU validity filter + MIDI writer);
    the run's own bookkeeping (valid counts, log file) goes on as with the reference's decoder."""
    model_path, out_dir = checkpoint
    conds = [O.make_synthetic_batch("modification", 3, L, seed=30 + i) for i in range(2)]
    batches = [{k: torch.from_numpy(v) for k, v in c.items()} for c in conds]
    argv = ["--step", "20", "--batch_size", "3", "--strength", "0.5", "--use_corruption", "false"]
    got = H.run_main("modification", model_path, out_dir, argv, batches=batches, dropin=True, stream_seed=9, midi_tail=True)
    assert len(got["valid"]) == 2
    files = sorted(f for f in os.listdir(got["output_dir"]) if f.endswith(".midi"))
    want = []
    for b, (count, invalid) in enumerate(got["valid"]):
        assert count + len(invalid) == 3
        want += ["%07d_batch%05d_%04d.midi" % (b * 3 + k, b, k) for k in range(3) if k not in invalid]
    assert files == sorted(want)
    printed = capsys.readouterr().out
    assert printed.count("Summary of Batch") >= 2
