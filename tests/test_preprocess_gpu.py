"""GPU parity tests of md_merge_and_mask through the C-ABI: bit-exact against the fixture
generated from the unmodified reference, against oracle/preprocess_oracle.py on fresh rows, and the encode -> decode
round trip through both kernels at the BASELINE sequence length."""
import os

import numpy as np
import pytest
import torch

import decode_oracle as D
import preprocess_oracle as P

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():  # pragma: no cover
    pytest.skip("needs a CUDA device", allow_module_level=True)

from musediffusion_b200 import decode_util, ops, preprocess  # noqa: E402

DEV = torch.device("cuda:0")


def run_kernel(src, src_len, trg, trg_len, seq_len):
    out = ops.merge_and_mask(*[torch.as_tensor(x).to(DEV) for x in (src, src_len, trg, trg_len)], seq_len)
    torch.cuda.synchronize()
    return [t.cpu().numpy() for t in out]


def test_matches_reference_fixture(golden_dir):
    g = np.load(os.path.join(golden_dir, "merge_and_mask.npz"), allow_pickle=False)
    L = int(g["seq_len"])
    ids, mask, length = run_kernel(g["src"], g["src_len"], g["trg"], g["trg_len"], L)
    assert np.array_equal(length, g["length"])
    keep = length <= L
    assert np.array_equal(ids[keep], g["kept_input_ids"])
    assert np.array_equal(mask[keep], g["kept_input_mask"])
    assert (ids[~keep] == 0).all() and (mask[~keep] == 1).all()


@pytest.mark.parametrize("seed,n,max_trg,seq_len", [(1, 50, 40, 48), (2, 300, 500, 333), (3, 200, 2000, 2096)])
def test_matches_oracle_on_fresh_rows(seed, n, max_trg, seq_len):
    src, src_len, trg, trg_len = P.merge_cases(seed=seed, n_rows=n, max_trg=max_trg)
    want = P.merge_and_mask_batch(src, src_len, trg, trg_len, seq_len)
    got = run_kernel(src, src_len, trg, trg_len, seq_len)
    for w, k, name in zip(want, got, ["input_ids", "input_mask", "length"]):
        assert np.array_equal(w, k), name


def test_round_trip_through_both_kernels_at_full_length():
    """encode (merge_and_mask) -> decode (split + restore_chord + strict validation) on 256 well-formed rows padded to
    L = 2096: the restored note sequence is the original target, the meta comes back, the strict grammar accepts it."""
    rows = P.well_formed_rows(seed=21, n_rows=256, max_bars=40, max_notes=10)
    col = preprocess.merge_and_collate([m for m, _ in rows], [t for _, t in rows], seq_len=2096, device=DEV)
    assert col["input_ids"].shape == (256, 2096) and col["input_ids"].dtype == torch.long
    prep = decode_util.prepare_batch(col["input_ids"], col["input_mask"], strict_validation=True)
    assert prep.valid_count >= 250
    for b, (meta, t) in enumerate(rows):
        assert prep.status[b] in (D.OK, D.VALIDATION_FAILED)
        assert prep.note_seqs[b].tolist() == t
        assert prep.metas[b].tolist() == meta


def test_host_mirror_filters_like_helper_filter(golden_dir):
    g = np.load(os.path.join(golden_dir, "merge_and_mask.npz"), allow_pickle=False)
    B = len(g["src_len"])
    src_rows = [g["src"][b, :g["src_len"][b]].tolist() for b in range(B)]
    trg_rows = [g["trg"][b, :g["trg_len"][b]].tolist() for b in range(B)]
    col = preprocess.merge_and_collate(src_rows, trg_rows, seq_len=int(g["seq_len"]), device=DEV)
    assert np.array_equal(col["input_ids"].cpu().numpy(), g["kept_input_ids"])
    assert np.array_equal(col["input_mask"].cpu().numpy(), g["kept_input_mask"])
    assert np.array_equal(col["length"].cpu().numpy(), g["kept_length"])
    assert np.array_equal(col["kept_index"].cpu().numpy(), np.nonzero(g["length"] <= int(g["seq_len"]))[0])
    with pytest.raises(ValueError):
        preprocess.merge_and_collate(src_rows, trg_rows[:-1], seq_len=96, device=DEV)
