"""Host mirror of the reference's sample-quality metrics (MuseDiffusion/metric.py) with the
same function names and return conventions, computed for the whole list in two kernel launches (`md_sequence_metrics`,
`md_onnc`) instead of per-token Python loops."""
import numpy as np
import torch

from . import ops

PITCH_RANGE = {631: (3, 38), 632: (39, 50), 633: (51, 62), 634: (63, 74), 635: (75, 86), 636: (87, 98), 637: (99, 130)}


def _pad(midilist, device):
    B = len(midilist)
    Ln = max([len(m) for m in midilist] + [1])
    arr = np.zeros((B, Ln), np.int32)
    for b, m in enumerate(midilist):
        arr[b, :len(m)] = np.asarray(m.cpu() if torch.is_tensor(m) else m)
    lens = np.array([len(m) for m in midilist], np.int32)
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    return torch.from_numpy(arr).to(dev), torch.from_numpy(lens).to(dev), dev


def _run(midilist, metas=None, device=None):
    notes, lens, dev = _pad(midilist, device)
    if metas is None:
        meta = torch.zeros((len(midilist), 11), dtype=torch.int32, device=dev)
    else:
        meta = torch.as_tensor(np.asarray([np.asarray(m.cpu() if torch.is_tensor(m) else m)[:11] for m in metas])).to(dev)
    return ops.sequence_metrics(notes, lens, meta)


def get_vectors(midi, note_len=128, device=None):
    """metric.py:4-75 for one sequence -> (rhythm [32], melody [12], harmony [12]) float32 tensors."""
    if note_len != 128:
        raise NotImplementedError("the kernel is specialised for note_len = 128 (the only value the reference uses)")
    vec, status, _ = _run([midi], device=device)
    if int(status[0]) != 0:
        raise ValueError("wrong midi format (the reference's get_vectors raises on this sequence)")
    return vec[0, :32], vec[0, 32:44], vec[0, 44:56]


def ONNC(midilist, return_vectors=False, return_MSIM=False, return_mostsim=False, device=None):
    """metric.py:89-117: the first half of `midilist` is ground truth, the second half generated."""
    vec, status, _ = _run(midilist, device=device)
    if int(status.max()) != 0:
        raise ValueError("wrong midi format in row %d" % int(torch.nonzero(status)[0]))
    most, msim = ops.onnc_nearest(vec, want_msim=return_MSIM)
    most = most.long()
    n = len(midilist)
    half = n // 2
    onnc = ((most[:half] < half).sum() + (most[half:] >= half).sum()) / n
    if not any([return_vectors, return_MSIM, return_mostsim]):
        return onnc
    out = [onnc]
    if return_vectors:
        out.append([[vec[:, :32], vec[:, 32:44], vec[:, 44:56]]])
    if return_MSIM:
        out.append(msim)
    if return_mostsim:
        out.append(most)
    return out


def Controllability_Pitch(metas, midis, device=None):
    """metric.py:131-149 -> (total, num_wrong)"""
    _, _, stats = _run(midis, metas, device)
    stats = stats.cpu().numpy().astype(np.int64)
    wrong = 0
    for meta, (psum, pcnt, _, _) in zip(metas, stats):
        pr = int(meta[3])
        if pr != 630:
            lo, hi = PITCH_RANGE[pr]
            if pcnt == 0 or not (lo * pcnt <= psum <= hi * pcnt):       # mean of an empty selection is NaN: counted wrong
                wrong += 1
    return len(metas), wrong


def Controllability_Velocity(metas, midis, device=None):
    """metric.py:152-169 -> (total, num_wrong)"""
    _, _, stats = _run(midis, metas, device)
    stats = stats.cpu().numpy().astype(np.int64)
    total = wrong = 0
    for meta, (_, _, vcnt, vwrong) in zip(metas, stats):
        if int(meta[8]) - 524 != 130:
            total += int(vcnt)
            wrong += int(vwrong)
    return total, wrong
