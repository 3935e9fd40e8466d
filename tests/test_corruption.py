"""SURVEY.md section 8(f) row 2 (second half): the modification-mode corruptions reproduce the UNMODIFIED reference
(`MuseDiffusion/data/corruption.py`, fixture written by oracle/make_golden.py::golden_corruption) token for token and leave
the shared random stream at the same position."""
import os

import numpy as np
import pytest
import torch

from musediffusion_b200 import corruption as C
from musediffusion_b200.initialization import seed_all

CONFIGS = {"mt": ("mt", 1, 1.0, None), "mn": ("mn", 1, 1.0, None), "rn": ("rn", 1, 1.0, None), "rr": ("rr", 1, 1.0, None),
           "default": ("mt,mn,rn,rr", 4, 0.5, None), "kw": ("rr,mt,rn", 2, 0.7, "dict(p=0.15, count=2)")}


@pytest.mark.parametrize("tag", list(CONFIGS))
def test_corruptions_match_reference(golden_dir, tag):
    g = np.load(os.path.join(golden_dir, "corruption.npz"))
    rows = g["rows"]
    corr = C.Corruptions.from_config(*CONFIGS[tag])
    C.generator.seed(1234)
    want = [g[tag][i, :n] for i, n in enumerate(g[tag + "_len"])]       # a rotation behind a masked EOS can lengthen a row
    before = rows.copy()
    got = [corr(torch.from_numpy(r)).numpy() for r in rows]
    assert all(np.array_equal(a, b) for a, b in zip(got, want)), [i for i, (a, b) in enumerate(zip(got, want)) if not np.array_equal(a, b)]
    assert C.generator.random() == float(g[tag + "_next"])              # consumed exactly as many draws as the reference
    assert np.array_equal(rows, before)                                 # inputs untouched (inplace=False)
    assert any(len(a) != len(r) or (a != r).any() for a, r in zip(got, rows))
    # numpy rows work too and give the same result
    C.generator.seed(1234)
    got_np = [corr(r) for r in rows]
    assert all(np.array_equal(a, b) for a, b in zip(got_np, want))


def test_seed_all_seeds_the_corruption_stream_and_edge_cases():
    seed_all(105, deterministic=True)
    a = C.generator.random()
    seed_all(105, deterministic=True)
    assert C.generator.random() == a
    row = torch.tensor([600, 610, 626, 630, 638, 641, 650, 660, 670, 720, 726, 1, 131, 5, 310, 1, 0, 0])
    # a velocity token at index 12 with the note cut by the row end is skipped (idx + 3 > len), like the reference
    short = torch.tensor([600, 610, 626, 630, 638, 641, 650, 660, 670, 720, 726, 1, 440, 140, 60])
    C.generator.seed(1)
    assert torch.equal(C.masking_note(short, 1.0), short)
    C.generator.seed(1)
    out = C.masking_note(row, 1.0)
    assert out[11:15].tolist() == [0, 0, 0, 0] and out[15] == 1 and torch.equal(row[12:15], torch.tensor([131, 5, 310]))
    with pytest.raises(AssertionError):
        C.random_rotating(row, 1)                                        # fewer than two bars (corruption.py:176)
    with pytest.raises(AssertionError):
        C.Corruptions(("mt",), 2, 0.5)                                   # corr_max > len(corr_available)


def test_cli_corruption_flags_default_to_the_training_run(golden_dir):
    """run/sample.py:125-135: --use_corruption / --corr_* left unset take training_args.json's values."""
    import json
    from types import SimpleNamespace
    from musediffusion_b200.sample import corruption_from_args, create_parser
    targs = SimpleNamespace(**json.load(open(os.path.join(golden_dir, "training_args_default.json"))))
    args = create_parser().parse_args(["modification", "--model_path", "m.pt"])
    c = corruption_from_args(args, targs)
    assert c is not None and (c.corr_max, c.corr_p, len(c.corr_available)) == (targs.corr_max, targs.corr_p, 4)
    args = create_parser().parse_args(["modification", "--model_path", "m.pt", "--corr_available", "mt,rr", "--corr_max", "1", "--corr_p", "1.0"])
    c = corruption_from_args(args, targs)
    assert (c.corr_max, c.corr_p, len(c.corr_available)) == (1, 1.0, 2)
    args = create_parser().parse_args(["modification", "--model_path", "m.pt", "--use_corruption", "false"])
    assert corruption_from_args(args, targs) is None
