"""
ORACLE — TEST INFRASTRUCTURE ONLY.  Never imported by the product (`musediffusion_b200/`).

CPU restatement (numpy) of the sample-quality metrics the reference computes on the decoded note sequences,
SURVEY.md §8(f) row 4 (MuseDiffusion/metric.py, called from run/sample.py:245-272):

    get_vectors               (metric.py:4-75)     rhythm [32] / melody [12] / harmony [12] vectors of one sequence
    MSIM / ONNC               (:78-117)            similarity = product of the three cosine Gram matrices; 1-NN classifier score
    Controllability_Pitch     (:131-149)           rows whose mean pitch falls outside the range named by the meta
    Controllability_Velocity  (:152-169)           velocity tokens outside [min_vel, max_vel] of the meta

The reference mixes float32 tensors with Python floats (float64): amplitudes are computed in float64 and rounded when they
are stored into the float32 rhythm vector; that is kept.  Sequences the reference would raise on (no BAR, bad grammar,
running off the end) are reported as status 1 with zero vectors.

Pinned by tests/golden/metrics.npz (the unmodified reference on the same sequences).
"""
import numpy as np

F32 = np.float32
PITCH_RANGE = {631: (3, 38), 632: (39, 50), 633: (51, 62), 634: (63, 74), 635: (75, 86), 636: (87, 98), 637: (99, 130)}


def get_vectors(midi, note_len=128):
    """metric.py:4-75 -> (status, rhythm[32], melody[12], harmony[12]) float32."""
    midi = [int(v) for v in midi]
    n = len(midi)
    zero = (1, np.zeros(32, F32), np.zeros(12, F32), np.zeros(12, F32))

    def at(k):
        if k >= n:
            raise IndexError
        return midi[k]

    try:
        i = 0
        while at(i) != 2:
            i += 1
        i += 1
        rhythm = np.full(32, 1e-8, F32)
        tmp = np.full(32, 1e-8, F32)
        melody = np.full(12, 1e-8, F32)
        harmony = np.zeros(12, F32)
        cur, prev, prev_startp = -1, -1, -1
        startp = None
        while True:
            if at(i) <= 2:
                tmp = tmp / F32(np.sqrt(np.sum(tmp * tmp, dtype=F32)))
                rhythm = rhythm + tmp
                tmp = np.full(32, 1e-8, F32)
                i += 1
                if midi[i - 1] == 2:
                    prev_startp = -1
                    continue
                if startp is None:
                    raise NameError                      # the reference reads an unbound `startp` here
                if prev_startp != startp and prev >= 0:
                    melody[(cur - prev) % 12] += F32(1)
                break
            if not 432 <= at(i) <= 559:
                raise ValueError
            startp = at(i) - 432
            if 195 <= at(i + 1) <= 303:
                i += 2
                continue
            v1, v2, v3 = at(i + 1), at(i + 2), at(i + 3)
            if not (131 <= v1 <= 194 and 3 <= v2 <= 130 and 304 <= v3 <= 431):
                raise ValueError
            pitch = v2
            endp = startp + v3 - 303
            harmony[pitch % 12] += F32(1)
            for t in range(0, min(128, endp), 4):
                if t < startp:
                    continue
                amp = (0.00542676376 * (v1 - 130) * 2 + 0.310801) ** 2                 # float64, as in the reference
                val = F32(amp * max(0, 1 - (t - startp) / note_len))                   # rounded when stored / compared
                if val > tmp[t // 4]:
                    tmp[t // 4] = val
            if cur >= 0:
                if prev_startp != startp:
                    if prev >= 0:
                        melody[(cur - prev) % 12] += F32(1)
                    prev = cur
                    cur = pitch
            cur = max(pitch, cur)
            prev_startp = startp
            i += 4
    except (IndexError, ValueError, NameError):
        return zero
    norm = lambda v: v / F32(np.sqrt(np.sum(v * v, dtype=F32)))
    with np.errstate(invalid="ignore", divide="ignore"):
        return 0, norm(rhythm), norm(melody), norm(harmony)


def onnc(rhythm, melody, harmony):
    """metric.py:89-117 on stacked vectors [N, 32] / [N, 12] / [N, 12] -> (onnc, most_sim[N], msim[N, N])."""
    msim = (rhythm @ rhythm.T) * (melody @ melody.T) * (harmony @ harmony.T)
    msim = msim.astype(F32)
    np.fill_diagonal(msim, 0)
    most = np.argmax(msim, axis=1)
    N = len(most)
    half = N // 2
    score = (int((most[:half] < half).sum()) + int((most[half:] >= half).sum())) / N
    return score, most, msim


def controllability_pitch(metas, midis):
    """metric.py:131-149 -> (total, num_wrong)"""
    wrong = 0
    for meta, midi in zip(metas, midis):
        pr = int(meta[3])
        if pr != 630:
            midi = np.asarray(midi)
            pitch = midi[(midi >= 3) & (midi <= 130)]
            mean = float(pitch.mean()) if len(pitch) else float("nan")
            lo, hi = PITCH_RANGE[pr]
            if not (lo <= mean <= hi):
                wrong += 1
    return len(metas), wrong


def controllability_velocity(metas, midis):
    """metric.py:152-169 -> (total, num_wrong)"""
    total = wrong = 0
    for meta, midi in zip(metas, midis):
        lo, hi = int(meta[7]) - 524, int(meta[8]) - 524
        if hi != 130:
            midi = np.asarray(midi)
            vel = midi[(midi >= 131) & (midi <= 194)]
            total += len(vel)
            for e in vel:
                if not ((lo == 130 or lo <= e) and (hi == 195 or e <= hi)):
                    wrong += 1
    return total, wrong


def metric_cases(seed=5, n=64):
    """Strictly valid note sequences (what run/sample.py feeds the metrics: rows that passed validate_rigidly) with their
    11-token metas, built by preprocess_oracle.well_formed_rows, plus their chord-restored form."""
    import preprocess_oracle as P
    rng = np.random.default_rng(seed)
    metas, midis = [], []
    for meta, t in P.well_formed_rows(seed=seed, n_rows=n, max_bars=8, max_notes=6):
        meta = list(meta)
        meta[3] = int(rng.integers(630, 638))
        meta[7] = int(rng.integers(653, 719))
        meta[8] = int(rng.integers(653, 719))
        if not any(131 <= v <= 194 for v in t):
            continue                                     # a row without a single note never reaches the metrics
        metas.append(meta)
        midis.append(t)
    L = max(len(m) for m in midis)
    arr = np.zeros((len(midis), L), np.int64)
    for b, m in enumerate(midis):
        arr[b, :len(m)] = m
    return np.array(metas, np.int64), arr, np.array([len(m) for m in midis], np.int64)
