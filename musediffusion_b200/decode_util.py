"""Host mirror of the token-level half of the reference's post-sampling decode.

The reference walks the sampled batch row by row in Python, rank after rank (`run/sample.py:222-294` ->
`decode_batch` -> `batch_decode_seq2seq` / `batch_decode_generation`, `utils/decode_util.py:233-384`); each row goes
through `SequenceToMidi.decode` (:207-214) = split_meta_midi + remove_padding + restore_chord + validate_* and only
then to the MIDI writer.  Here the whole batch takes one kernel launch (`md_decode_prepare`); what comes back is, per
row, the reference's outcome (OK or the text of the SequenceToMidiError it would raise), the restored note sequence
and the 11 meta tokens — exactly the two arrays `decode_event_sequence` (:201-205) needs.  Writing MIDI files stays
with the reference (`miditoolkit` is not part of this path).
"""
from collections import namedtuple

import numpy as np
import torch

from . import ops

OK, NO_EOS, RESTORE_FAILED, VALIDATION_FAILED, STRICT_FAILED, INDEX_ERROR, TOO_LONG = range(7)
STATUS_TEXT = {
    OK: "OK",
    NO_EOS: "NO EOS TOKEN",
    RESTORE_FAILED: "RESTORE_CHORD FROM META FAILED",
    VALIDATION_FAILED: "VALIDATION OF SEQUENCE FAILED",
    STRICT_FAILED: "STRICT VALIDATION OF SEQUENCE FAILED",
    INDEX_ERROR: "IndexError",
    TOO_LONG: "RESTORED SEQUENCE TOO LONG",
}


class SequenceToMidiError(Exception):
    """same name as the reference's exception (decode_util.py:53-54)"""


PreparedBatch = namedtuple("PreparedBatch", "status note_seqs metas valid_count invalid_idxes")


def prepare_batch(sequences, input_ids_mask_ori, strict_validation=False, device=None):
    """Batched `split_meta_midi` + `validate_generated_sequence` (decode_util.py:186-199).

    sequences / input_ids_mask_ori: [B, L] integer arrays or tensors (what `decode_batch` receives, :233-257).
    Returns PreparedBatch(status int32 [B] (numpy), note_seqs: list of int64 arrays (None where they do not exist),
    metas: list of 11-token arrays, valid_count, invalid_idxes: sorted list) — `valid_count` / `invalid_idxes` are what
    `decode_batch(..., return_indices=True)` reports to `run/sample.py:231-243`."""
    dev = torch.device(device) if device is not None else (sequences.device if torch.is_tensor(sequences) and sequences.is_cuda
                                                          else torch.device("cuda", torch.cuda.current_device()))
    tok = torch.as_tensor(np.asarray(sequences) if not torch.is_tensor(sequences) else sequences).to(dev)
    msk = torch.as_tensor(np.asarray(input_ids_mask_ori) if not torch.is_tensor(input_ids_mask_ori) else input_ids_mask_ori).to(dev)
    if tok.dim() != 2 or tok.shape != msk.shape:
        raise ValueError("sequences and input_ids_mask_ori must both be [B, L]")
    status, note_len, notes, meta = ops.decode_prepare(tok, msk, strict_validation)
    status, note_len, notes, meta = status.cpu().numpy(), note_len.cpu().numpy(), notes.cpu().numpy(), meta.cpu().numpy()
    note_seqs, metas, invalid = [], [], []
    for b in range(len(status)):
        has = status[b] in (OK, VALIDATION_FAILED, STRICT_FAILED) or (status[b] == INDEX_ERROR and note_len[b] > 0)
        note_seqs.append(notes[b, :note_len[b]].astype(np.int64) if has else None)
        metas.append(meta[b].astype(np.int64) if has else None)
        if status[b] != OK:
            invalid.append(b)
    return PreparedBatch(status, note_seqs, metas, len(status) - len(invalid), invalid)


def report_failures(prepared, batch_index, previous_count, print_fn=print):
    """the warnings `batch_decode_*` prints for rows that fail (decode_util.py:287-293), same wording"""
    for index in prepared.invalid_idxes:
        code = int(prepared.status[index])
        if code == INDEX_ERROR:
            raise IndexError("row %d of batch %d: the reference's decode aborts here (index out of range while validating)"
                             % (index, batch_index))
        print_fn("<Warning> Batch %d Index %d (Original: %d) - Generation Failure: %s"
                 % (batch_index, index, previous_count + index, STATUS_TEXT[code]))
