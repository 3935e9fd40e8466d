"""Modification-mode input corruptions — host mirror of MuseDiffusion/data/corruption.py (:8-195), same names, same
arguments, same results for the same seed (second half).

The four corruptions draw from ONE sequential `random.Random` stream (`generator`, seeded by `seed_all`,
utils/initialization.py:15-26) and how many numbers a row consumes depends on its tokens AND on earlier draws (which
corruptions fire, in which shuffled order, rejection loops inside `randint`), so row r+1 cannot start before row r is
finished: there is no data-parallel formulation that keeps the reference's outputs, and the work per row is a few hundred
integer operations.  It therefore stays on the host; what runs on the GPU of this pipeline stage is the deterministic
part (`md_merge_and_mask`, preprocess.py).  Here the per-token Python loops of the reference are replaced by bulk draws
from the same generator followed by vectorised numpy updates wherever the draw count is known up front.

Token ranges (commu/preprocessor/encoder/event_tokens.py): EOS 1, BAR 2, pitch 3-130, velocity 131-194, duration 304-431,
position 432-559.  A row is [meta(11) .. EOS .. notes .. EOS? .. zero padding]."""
import random

import numpy as np
import torch

generator = random.Random()         # corruption.py:6 — seeded together with python / numpy / torch by seed_all


def _as_array(seq, inplace):
    """1-D integer tensor / array -> (numpy view or copy to work on, wrap-back function)."""
    if isinstance(seq, torch.Tensor):
        assert seq.ndim == 1
        work = seq if inplace else seq.clone()
        return work.numpy(), lambda a: work              # the numpy view shares the tensor's memory
    arr = np.asarray(seq)
    assert arr.ndim == 1
    work = arr if inplace else arr.copy()
    return work, lambda a: a


def _draws(n):
    """n consecutive `generator.random()` values, in stream order."""
    rnd = generator.random
    return np.fromiter((rnd() for _ in range(n)), dtype=np.float64, count=n)


def masking_token(seq, p, inplace=False):
    """corruption.py:99-113 ('mt', p = 0.3): every token from index 12 up to (not including) the first EOS is set to 0 with
    probability p; one draw per visited token."""
    a, wrap = _as_array(seq, inplace)
    tail = a[12:]
    eos = np.flatnonzero(tail == 1)
    n = int(eos[0]) if eos.size else int(tail.shape[0])
    hit = _draws(n) < p
    tail[:n][hit] = 0
    return wrap(a)


def _velocity_slots(a):
    """indices of velocity tokens that the reference does not skip (`if idx + 3 > len(seq): continue`)"""
    idx = np.flatnonzero((a >= 131) & (a <= 194))
    return idx[idx + 3 <= a.shape[0]]


def masking_note(seq, p, inplace=False):
    """corruption.py:116-133 ('mn', p = 0.5): a note (position, velocity, pitch, duration) is zeroed with probability p; one
    draw per velocity token.  `corrupted[idx-1:idx+3]` keeps Python's slice meaning for idx = 0 (empty slice)."""
    a, wrap = _as_array(seq, inplace)
    idx = _velocity_slots(a)
    hit = _draws(idx.shape[0]) < p
    for i in idx[hit]:
        a[int(i) - 1:int(i) + 3] = 0
    return wrap(a)


def randomize_note(seq, p, inplace=False):
    """corruption.py:136-163 ('rn', p = 0.5): with probability p a note gets a new velocity / pitch / duration
    (`randint` inclusive ranges 131-194, 3-130, 304-431).  The draws interleave (one uniform, then three rejection-sampled
    integers only if it fired), so this one walks the velocity tokens in order."""
    a, wrap = _as_array(seq, inplace)
    for i in _velocity_slots(a):
        if generator.random() < p:
            i = int(i)
            a[i] = generator.randint(131, 194)
            a[i + 1] = generator.randint(3, 130)
            a[i + 2] = generator.randint(304, 431)
    return wrap(a)


def random_rotating(seq, count, inplace=False):
    """corruption.py:166-195 ('rr', count = 3): `count` times, two bars picked by `generator.sample` swap places.  As in the
    reference the bar boundaries are those of the ORIGINAL row (they are not re-derived after a swap), the last bar ends at
    the last EOS, and the result is a new row (never in place)."""
    is_tensor = isinstance(seq, torch.Tensor)
    a = (seq if inplace else seq.clone()).numpy() if is_tensor else (np.asarray(seq) if inplace else np.asarray(seq).copy())
    src = seq.numpy() if is_tensor else np.asarray(seq)
    bar_idx = np.flatnonzero(src == 2)
    eos_idx = int(np.flatnonzero(src == 1)[-1])
    for _ in range(count):
        assert len(bar_idx) > 1
        first, second = sorted(generator.sample(range(0, len(bar_idx)), 2))
        b1s, b2s = int(bar_idx[first]), int(bar_idx[second])
        b1e = int(bar_idx[first + 1])
        b2e = int(bar_idx[second + 1]) if second < len(bar_idx) - 1 else eos_idx
        a = np.concatenate([a[:b1s], a[b2s:b2e], a[b1e:b2s], a[b1s:b1e], a[b2e:]])
    return torch.from_numpy(a) if is_tensor else a


class Corruptions:
    """corruption.py:9-96: a shuffled subset of the registered corruptions, each applied with probability corr_p."""

    MAP = {"mt": (masking_token, ["p"], {"p": 0.3}), "mn": (masking_note, ["p"], {"p": 0.5}),
           "rn": (randomize_note, ["p"], {"p": 0.5}), "rr": (random_rotating, ["count"], {"count": 3})}

    @classmethod
    def from_config(cls, corr_available, corr_max, corr_p, corr_kwargs=None):
        return cls(corr_available=tuple(corr_available.split(",")), corr_max=int(corr_max), corr_p=float(corr_p),
                   corr_kwargs=eval(corr_kwargs) if corr_kwargs else None)      # the reference evals this config string too (:25)

    def __init__(self, corr_available, corr_max, corr_p, corr_kwargs=None):
        assert all(key in self.MAP or callable(key) for key in corr_available)
        assert 0 <= corr_max <= len(corr_available) and 0 <= corr_p <= 1
        assert corr_kwargs is None or isinstance(corr_kwargs, dict)
        self.corr_available = tuple(self.get(key, corr_kwargs) for key in corr_available)
        self.corr_max, self.corr_p, self.corr_kwargs = corr_max, corr_p, corr_kwargs

    @classmethod
    def get(cls, key, update_kwargs=None, inplace=None):
        if callable(key):
            return key
        func, required, defaults = cls.MAP[key]
        defaults = dict(defaults)
        if update_kwargs is not None:
            defaults.update(update_kwargs)
        kwargs = {k: defaults[k] for k in required}
        if inplace is not None:
            kwargs.update(inplace=inplace)
        if kwargs:
            from functools import partial
            func = partial(func, **kwargs)
        return func

    def __call__(self, seq, inplace=False):
        assert seq.ndim == 1
        corrupted = seq if inplace else (seq.clone() if isinstance(seq, torch.Tensor) else np.array(seq))
        fns = list(self.corr_available)
        generator.shuffle(fns)
        for fn in fns[:self.corr_max]:
            if generator.random() > 1 - self.corr_p:
                corrupted = fn(corrupted, inplace=True)
        return corrupted


def corrupt_batch(corruption, input_ids):
    """MidiSequenceDataset.__getitem__ for a list / batch of rows (data/wrapper.py:73-86): rows are corrupted one after the
    other in order (the stream is sequential); returns (corrupted rows, correct rows)."""
    rows = list(input_ids)
    return [corruption(r) for r in rows], rows
