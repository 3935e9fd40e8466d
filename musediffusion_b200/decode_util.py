"""Host mirror of the token-level half of the reference's post-sampling decode.

The reference walks the sampled batch row by row in Python, rank after rank (`run/sample.py:222-294` ->
`decode_batch` -> `batch_decode_seq2seq` / `batch_decode_generation`, `utils/decode_util.py:233-384`); each row goes
through `SequenceToMidi.decode` (:207-214) = split_meta_midi + remove_padding + restore_chord + validate_* and only
then to the MIDI writer.  Here the whole batch takes one kernel launch (`md_decode_prepare`); what comes back is, per
row, the reference's outcome (OK or the text of the SequenceToMidiError it would raise), the restored note sequence
and the 11 meta tokens — exactly the two arrays `decode_event_sequence` (:201-205) needs.  `decode_batch` then turns
every valid row into a MIDI file through `midi.py` (vectorised event walk + a Standard-MIDI-File writer; `miditoolkit`
is not needed), with the reference's file names, warnings and summaries.
"""
import os

from collections import namedtuple

import numpy as np
import torch

from . import midi, ops

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


def decode_batch(mode, sequences, input_ids_mask_ori, batch_index, previous_count, output_dir, return_indices=False,
                 strict_validation=False, prepared=None):
    """`decode_batch` (decode_util.py:233-257) with both writers behind it: modification keeps the original index in the
    file name and reports every failure (`batch_decode_seq2seq`, :260-331); generation numbers the VALID files
    consecutively from `previous_count` and skips failures quietly (`batch_decode_generation`, :334-384).  Rows the
    reference would abort on (IndexError inside its validation) raise here too.  Returns valid_count, or
    (valid_count, invalid_idxes) with return_indices.  `prepared` (not in the reference): a PreparedBatch the caller already
    holds for these rows."""
    if mode not in ("generation", "modification"):
        raise AssertionError("Unknown decoding mode")
    prep = prepared if prepared is not None else prepare_batch(sequences, input_ids_mask_ori, strict_validation=strict_validation)
    n, valid_index = len(prep.status), previous_count
    for index in range(n):
        code = int(prep.status[index])
        if code == INDEX_ERROR:
            raise IndexError("row %d of batch %d: the reference's decode aborts here (index out of range while validating)"
                             % (index, batch_index))
        if code != OK:
            if mode == "modification":
                print("<Warning> Batch %d Index %d (Original: %d) - Generation Failure: %s"
                      % (batch_index, index, previous_count + index, STATUS_TEXT[code]))
            continue
        decoded = midi.decode_event_sequence(prep.note_seqs[index], prep.metas[index])
        log = " ".join("OOV: %d" % w for w in decoded.oov)
        if mode == "modification":
            if log:
                print("<Warning> Batch %d Index %d (Original: %d) - %s" % (batch_index, index, previous_count + index, log))
            name = "%07d_batch%05d_%04d.midi" % (previous_count + index, batch_index, index)
        else:
            if log:
                print("<Warning> Index %d - %s" % (valid_index, log))
            name = "generated_%07d.midi" % valid_index
        decoded.dump(os.path.join(output_dir, name))
        valid_index += 1
    valid_count = valid_index - previous_count
    if mode == "modification":
        print("\n" + (" Summary of Batch %d " % batch_index).center(60, "=") + "\n"
              " * Original index: from %d to %d\n * %d valid sequences are converted to midi into path:\n     %s\n"
              " * %d sequences are invalid.\n" % (previous_count, previous_count + n, valid_count, os.path.abspath(output_dir),
                                                 len(prep.invalid_idxes))
              + (" * Index (in batch %d) of invalid sequence:\n    %s\n" % (batch_index, prep.invalid_idxes)
                 if prep.invalid_idxes else "") + "=" * 60 + "\n")
    else:
        print("\n" + (" Summary of Trial %d " % batch_index).center(60, "=") + "\n"
              " * %d valid sequences are converted to midi into path:\n     %s\n * Totally %d sequences are converted.\n"
              % (valid_count, os.path.abspath(output_dir), valid_index) + "=" * 60 + "\n")
    return (valid_count, prep.invalid_idxes) if return_indices else valid_count
