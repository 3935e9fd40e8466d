"""Generation-mode input format: meta dict -> token prefix -> `{'input_ids','input_mask'}` batch.

Mirror of `MetaToSequence` / `meta_to_batch` (MuseDiffusion/utils/decode_util.py:16-50,221-230), whose meta half is the
ComMU `MetaEncoder` (commu/preprocessor/encoder/meta.py:108-250) over the vocabulary offsets of
commu/preprocessor/encoder/event_tokens.py:308-329 and the category tables of commu/preprocessor/utils/constants.py.
Host-side integer bookkeeping (eleven meta tokens + the chord prefix, once per run) — table-driven here; the product
path starts at the batch this returns.  Pinned by tests/golden/meta_encode.json (vectors written from the reference's
own classes)."""
import json
import math

import torch

UNKNOWN = "unknown"                         # constants.py:29

# vocabulary offsets (event_tokens.py:308-329): each meta field owns [offset, next offset); offset itself = "unknown"
CHORD_START, POSITION = 195, 432
OFFSET = {"bpm": 560, "audio_key": 601, "time_signature": 626, "pitch_range": 630, "num_measures": 638, "inst": 641,
          "genre": 650, "velocity": 653, "track_role": 719, "rhythm": 726}
MAX_BPM, BPM_INTERVAL, VELOCITY_INTERVAL = 200, 5, 2

FIELDS = ("bpm", "audio_key", "time_signature", "pitch_range", "num_measures", "inst", "genre", "min_velocity",
          "max_velocity", "track_role", "rhythm")                      # container.py:24-35 declaration order

_ROOTS = ("c", "c#", "d", "d#", "e", "f", "f#", "g", "g#", "a", "a#", "b")
_FLATS = {"db": "c#", "eb": "d#", "gb": "f#", "ab": "g#", "bb": "a#"}
KEY_MAP = {r + q: i + 12 * k for k, q in enumerate(("major", "minor")) for i, r in enumerate(_ROOTS)}
KEY_MAP.update({f + q: KEY_MAP[s + q] for f, s in _FLATS.items() for q in ("major", "minor")})   # constants.py:35-70
TIME_SIG_MAP = {"4/4": 0, "3/4": 1, "6/8": 2, "12/8": 3}
PITCH_RANGE_MAP = {n: i for i, n in enumerate(("very_low", "low", "mid_low", "mid", "mid_high", "high", "very_high"))}
GENRE_MAP = {"newage": 0, "cinematic": 1}
TRACK_ROLE_MAP = {n: i for i, n in enumerate(("main_melody", "sub_melody", "accompaniment", "bass", "pad", "riff"))}
RHYTHM_MAP = {"standard": 0, "triplet": 1}
_INST_FAMILIES = {                                                                               # constants.py:93-156
    0: "acoustic_piano electric_piano harpsichord keyboard organ",
    1: "accordion synth_lead",
    2: "bell celesta glockenspiel marimba synth_bell vibraphone xylophone orgel",
    3: "acoustic_bass acoustic_guitar banjo electric_bass electric_guitar_clean electric_guitar_distortion harp mandolin "
       "nylon_guitar oud sitar synth_bass synth_bass_808 synth_bass_wobble ukulele zither yanggeum",
    4: "fiddle pad_synth string_cello string_double_bass string_ensemble string_viola string_violin synth_pad",
    5: "bassoon brass_ensemble clarinet flute horn oboe recorder trombone trumpet tuba synth_brass sax bamboo_flute",
    6: "drums_full drums_tops percussion timpani",
    7: "choir synth_pluck synth_voice whistle",
    8: "vocal",
}
INST_MAP = {name: fam for fam, names in _INST_FAMILIES.items() for name in names.split()}
CATEGORY = {"audio_key": KEY_MAP, "time_signature": TIME_SIG_MAP, "pitch_range": PITCH_RANGE_MAP, "inst": INST_MAP,
            "genre": GENRE_MAP, "track_role": TRACK_ROLE_MAP, "rhythm": RHYTHM_MAP}

# chord vocabulary (event_tokens.py:195-303): roots a..g# x nine qualities, then "NN"; keys capitalised like
# MetaToSequence.chord_map (decode_util.py:20-23)
_CHORD_ROOTS = ("a", "a#", "b", "c", "c#", "d", "d#", "e", "f", "f#", "g", "g#")
_CHORD_QUALITIES = ("", "7", "+", "dim", "m", "m7", "m7b5", "maj7", "sus4")
CHORD_MAP = {(r + q)[0].upper() + (r + q)[1:]: CHORD_START + 9 * i + j
             for i, r in enumerate(_CHORD_ROOTS) for j, q in enumerate(_CHORD_QUALITIES)}
CHORD_MAP["NN"] = CHORD_START + 9 * len(_CHORD_ROOTS)


class UnprocessableMidiError(ValueError):
    """commu/preprocessor/utils/exceptions.py: raised for a value outside the tables."""


def encode_field(name, value):
    """One meta token.  "unknown" -> the field's offset (meta.py:84-101; num_measures raises instead, :157-158); a known
    value -> offset + 1 + index, except bpm whose bins start at the offset itself (meta.py:50) and num_measures whose
    three classes sit directly at 638/639/640 (meta.py:56-58)."""
    base = OFFSET["velocity" if name.endswith("velocity") else name]
    if name == "num_measures":
        if value == UNKNOWN:
            raise UnprocessableMidiError("num measures unknown")
        n = math.floor(value)
        if n not in (4, 5, 8, 9, 16, 17):
            raise UnprocessableMidiError("num measures ValueError: %s" % n)
        return base + (0 if n < 8 else 1 if n < 16 else 2)
    if value == UNKNOWN:
        return base
    if name == "bpm":
        return base + max(min(value, MAX_BPM) // BPM_INTERVAL, 1)
    if name == "min_velocity":
        return base + 1 + math.floor(value / VELOCITY_INTERVAL)
    if name == "max_velocity":
        return base + 1 + math.ceil(value / VELOCITY_INTERVAL)
    try:
        return base + 1 + CATEGORY[name][value]
    except KeyError:
        raise UnprocessableMidiError("%s KeyError: %s" % (name, value)) from None


def encode_meta(midi_meta):
    """Eleven tokens in declaration order (meta.py:230-241)."""
    return [encode_field(name, midi_meta[name]) for name in FIELDS]


def normalize_chord_progression(text):
    """config/sample.py:173-177: list-literal punctuation -> the dash-separated form."""
    mapping = {",": "-", "[": "", "]": "", "'": "", " ": ""}
    return "".join(mapping.get(c, c) for c in text)


def encode_chord(chords):
    """decode_util.py:25-39: eight chord slots per bar; each bar opens with Position 0 (432) + its chord, and every change
    inside the bar adds Position (432 + 16 * slot) WITHOUT the new chord's token (the reference's behaviour, kept)."""
    if len(chords) % 8:
        raise AssertionError("chord progression must hold 8 entries per bar, got %d" % len(chords))
    out = []
    for bar in range(0, len(chords), 8):
        out += [POSITION, CHORD_MAP[chords[bar]]]
        current = chords[bar]
        for slot in range(1, 8):
            if chords[bar + slot] != current:
                out.append(POSITION + 16 * slot)
                current = chords[bar + slot]
    return out


def meta_to_sequence(midi_meta_dict):
    """MetaToSequence.execute (decode_util.py:44-47)."""
    return encode_meta(midi_meta_dict) + encode_chord(midi_meta_dict["chord_progression"].split("-"))


def meta_to_batch(midi_meta_dict, batch_size, seq_len, device=None):
    """decode_util.py:221-230: every row = the prefix then zeros; the mask frees everything after prefix + 1 (the slot the
    datasets put the separator in)."""
    prefix = torch.tensor(meta_to_sequence(midi_meta_dict), dtype=torch.int32)
    if len(prefix) + 1 > seq_len:
        raise RuntimeError("meta prefix of %d tokens does not fit seq_len %d" % (len(prefix), seq_len))
    input_ids = torch.zeros(batch_size, seq_len, dtype=torch.int32)
    input_ids[:, :len(prefix)] = prefix
    input_mask = torch.ones(batch_size, seq_len, dtype=torch.int32)
    input_mask[:, :len(prefix) + 1] = 0
    batch = {"input_ids": input_ids, "input_mask": input_mask}
    return batch if device is None else {k: v.to(device) for k, v in batch.items()}


_INT_FIELDS, _FLOAT_FIELDS = ("bpm", "min_velocity", "max_velocity"), ("num_measures",)


def add_meta_arguments(parser):
    """config/sample.py:157-193,223-230: one flag per meta field + chord_progression, or --meta_json for all of them."""
    group = parser.add_argument_group(title="meta")
    group.add_argument("--meta_json", type=str, default=None, help="json file holding every meta field below")
    for name in FIELDS:
        if name in CATEGORY:
            group.add_argument("--" + name, type=str, choices=tuple(CATEGORY[name]), default=None,
                               metavar="{%s}" % ", ".join(CATEGORY[name]))
        else:
            group.add_argument("--" + name, type=int if name in _INT_FIELDS else float, default=None)
    group.add_argument("--chord_progression", type=str, default=None)
    return parser


def meta_from_args(args):
    """config/sample.py:236-254: the dict `meta_to_batch` takes, or None when no meta flag was given.  Like the reference,
    a partial set of flags is an error."""
    if getattr(args, "meta_json", None):
        with open(args.meta_json) as f:
            meta = json.load(f)
    else:
        meta = {k: getattr(args, k, None) for k in FIELDS + ("chord_progression",)}
        if all(v is None for v in meta.values()):
            return None
    missing = [k for k in FIELDS + ("chord_progression",) if meta.get(k) is None]
    if missing:
        raise ValueError("meta fields missing: %s" % ", ".join(missing))
    for k in _INT_FIELDS:
        meta[k] = int(meta[k])
    meta["num_measures"] = float(meta["num_measures"])
    meta["chord_progression"] = normalize_chord_progression(meta["chord_progression"])
    return meta
