"""Host mirror of the modification-mode input preparation: the reference's
`helper_tokenize` / `merge_and_mask` (MuseDiffusion/data/preprocess.py:26-70), `helper_filter` (:73-81) and
`collate_batches` (MuseDiffusion/data/wrapper.py:90-126), for a whole list of raw rows in one kernel launch
(`md_merge_and_mask`).  The corruption functions (data/corruption.py) draw from Python's RNG and are not part of this."""
import numpy as np
import torch

from . import ops


def merge_and_collate(src_rows, trg_rows, seq_len, end_token=1, device=None, dtype=torch.long):
    """src_rows / trg_rows: lists of integer sequences (the raw `{'src': ..., 'trg': ...}` columns, preprocess.py:9-23).

    Returns a dict shaped like `collate_batches`' result for the rows that survive `helper_filter`
    (`input_ids`, `input_mask`, `length`: [B_kept, seq_len] / [B_kept], `dtype`), plus `kept_index` (positions of the
    surviving rows in the input lists) and `all_length` (merged length of every input row)."""
    if len(src_rows) != len(trg_rows):
        raise ValueError("src_rows and trg_rows must have the same number of rows")
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    B = len(src_rows)
    Ls = max([len(s) for s in src_rows] + [1])
    Lt = max([len(t) for t in trg_rows] + [1])
    src = np.zeros((B, Ls), np.int32)
    trg = np.zeros((B, Lt), np.int32)
    for b, (s, t) in enumerate(zip(src_rows, trg_rows)):
        src[b, :len(s)] = s
        trg[b, :len(t)] = t
    src_len = np.array([len(s) for s in src_rows], np.int32)
    trg_len = np.array([len(t) for t in trg_rows], np.int32)
    ids, mask, length = ops.merge_and_mask(torch.from_numpy(src).to(dev), torch.from_numpy(src_len).to(dev),
                                           torch.from_numpy(trg).to(dev), torch.from_numpy(trg_len).to(dev), seq_len, end_token)
    keep = torch.nonzero(length <= seq_len).flatten()
    return {"input_ids": ids[keep].to(dtype), "input_mask": mask[keep].to(dtype), "length": length[keep].to(dtype),
            "kept_index": keep, "all_length": length}
