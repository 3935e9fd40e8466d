"""Note sequence -> MIDI file: the last stage of the reference's post-sampling decode.

`SequenceToMidi.decode_event_sequence` (MuseDiffusion/utils/decode_util.py:201-205) hands the restored note sequence and
the eleven meta tokens to ComMU's `EventSequenceEncoder.decode` (commu/preprocessor/encoder/encoder.py:72-96) ->
`write_midi` (encoder_utils.py:386-497), which walks event objects one by one and fills a `miditoolkit.MidiFile`;
`.dump()` then serialises through mido.  Neither library is part of this image, and neither is needed: the event walk
is four vectorised comparisons over the token array, and a Standard MIDI File is a few dozen bytes of framing.
`decode_event_sequence` returns the same content the reference puts into its MidiFile (pinned note for note against
the reference's own write_midi in tests/golden/midi_decode.json); `DecodedMidi.dump` writes it as SMF format 1 with the
track layout miditoolkit uses (conductor track: time signature, tempo, markers, key; one instrument track).  Host side.
"""
import struct
from dataclasses import dataclass, field

import numpy as np

# event vocabulary (commu/preprocessor/encoder/event_tokens.py:308-329 + encoder_utils.py:47-57): word -> event
EOS, BAR, PITCH, VELOCITY, CHORD, DURATION, POSITION, META = 1, 2, 3, 131, 195, 304, 432, 560
TICKS_PER_BEAT, POSITION_RESOLUTION, BPM_INTERVAL = 480, 128, 5                      # constants.py:23-26
TIME_SIGNATURES = ("4/4", "3/4", "6/8", "12/8")                                      # constants.py:74-81
VELOCITY_BINS = np.linspace(2, 127, 64, dtype=int)                                   # encoder_utils.py:17-18
_K_BAR, _K_PITCH, _K_VEL, _K_CHORD, _K_DUR, _K_POS = range(6)
_CHORD_ROOTS = ("a", "a#", "b", "c", "c#", "d", "d#", "e", "f", "f#", "g", "g#")
_CHORD_QUALITIES = ("", "7", "+", "dim", "m", "m7", "m7b5", "maj7", "sus4")
CHORD_NAMES = tuple(r + q for r in _CHORD_ROOTS for q in _CHORD_QUALITIES) + ("NN",)  # event_tokens.py:195-303
# KEY_NUM_MAP (constants.py:72) inverts KEY_MAP, so the LAST spelling of each pitch class wins: the flat one
_KEY_ROOTS = ("c", "db", "d", "eb", "e", "f", "gb", "g", "ab", "a", "bb", "b")
KEY_NAMES = tuple(r + m for m in ("major", "minor") for r in _KEY_ROOTS)
# sharps (+) / flats (-) of each major key by pitch class, for the SMF key-signature event
_MAJOR_SF = (0, -5, 2, -3, 4, -1, -6, 1, -4, 3, -2, 5)


@dataclass
class DecodedMidi:
    """What write_midi leaves in its MidiFile (encoder_utils.py:462-497)."""
    tempo: int                                  # bpm
    numerator: int
    denominator: int
    key_name: str
    notes: np.ndarray                           # int64 [n, 4]: velocity, pitch, start tick, end tick (miditoolkit.Note order)
    marker_times: np.ndarray                    # int64 [m]
    marker_texts: list                          # chord names
    oov: list = field(default_factory=list)     # words outside the event vocabulary (the reference prints "OOV: w" for each)
    ticks_per_beat: int = TICKS_PER_BEAT

    def dump(self, path):
        with open(path, "wb") as f:
            f.write(self.to_bytes())

    def to_bytes(self):
        """SMF format 1, two tracks."""
        key = KEY_NAMES.index(self.key_name)
        minor = key >= 12
        sf = _MAJOR_SF[(key + 3) % 12] if minor else _MAJOR_SF[key]                # a minor key signs like its relative major
        conductor = [(0, b"\xff\x58\x04" + bytes((self.numerator, self.denominator.bit_length() - 1, 24, 8))),
                     (0, b"\xff\x51\x03" + struct.pack(">I", int(round(60_000_000 / self.tempo)))[1:])]
        for t, text in zip(self.marker_times.tolist(), self.marker_texts):
            raw = text.encode("latin-1")
            conductor.append((t, b"\xff\x06" + _varlen(len(raw)) + raw))
        conductor.append((0, b"\xff\x59\x02" + struct.pack(">bB", sf, int(minor))))
        conductor.sort(key=lambda e: e[0])                                          # stable: insertion order inside a tick
        events = [(0, 1, 0, b"\xc0\x00")]                                           # program 0 on channel 0 (Instrument(0))
        for vel, pitch, start, end in self.notes.tolist():
            events.append((start, 2, pitch, bytes((0x90, pitch, vel))))
            events.append((end, 0, pitch, bytes((0x80, pitch, vel))))
        events.sort(key=lambda e: e[:3])                                            # note-offs ahead of note-ons inside a tick
        return (b"MThd" + struct.pack(">IHHH", 6, 1, 2, self.ticks_per_beat)
                + _track(conductor) + _track([(t, raw) for t, _, _, raw in events]))


def _varlen(n):
    out = [n & 0x7F]
    n >>= 7
    while n:
        out.append(0x80 | (n & 0x7F))
        n >>= 7
    return bytes(reversed(out))


def _track(timed):
    body, now = bytearray(), 0
    for t, raw in timed:
        body += _varlen(t - now) + raw
        now = t
    body += b"\x00\xff\x2f\x00"
    return b"MTrk" + struct.pack(">I", len(body)) + bytes(body)


def decode_event_sequence(note_seq, encoded_meta):
    """`SequenceToMidi.decode_event_sequence(note_seq, encoded_meta)` (decode_util.py:201-205).

    Behaviours of write_midi kept on purpose: only words 2..559 are events (EOS is dropped silently, anything else is
    reported as OOV and dropped, encoder_utils.py:370-384); the walk stops three events before the end (:397), so a Bar or
    a chord among the last three events is not seen; a Bar at event 0 does not advance the bar counter (:398); starts come
    from `linspace(bar start, bar end, 128, endpoint=False, dtype=int)` (:431-439), i.e. floor(position * ticks_per_bar /
    128); durations from `arange(step, ticks_per_bar + 1, step)` with step = int(ticks_per_bar / 128) (encoder.py:83-88).
    An "unknown" key / time-signature token raises KeyError, as the reference's table lookups do."""
    words = np.asarray(note_seq, dtype=np.int64).reshape(-1)
    meta = [int(v) for v in np.asarray(encoded_meta).reshape(-1)[:11]]
    bpm_word, key_word, ts_word = meta[0], meta[1], meta[2]
    ts_index, key_index = ts_word - 627, key_word - 602
    if not 0 <= ts_index < len(TIME_SIGNATURES):
        raise KeyError(ts_index)
    numerator, denominator = (int(v) for v in TIME_SIGNATURES[ts_index].split("/"))
    ticks_per_bar = TICKS_PER_BEAT * int(numerator / denominator * 4)
    step = int(ticks_per_bar / POSITION_RESOLUTION)

    is_event = (words >= BAR) & (words < META)
    oov = words[~is_event & (words != EOS)].tolist()
    ev = words[is_event]
    kind = np.searchsorted(np.array([PITCH, VELOCITY, CHORD, DURATION, POSITION]), ev, side="right")
    n = len(ev)
    seen = np.arange(n) < n - 3                                          # range(len(events) - 3)

    def at(k, offset):                                                   # kind[i + offset] == k, False past the end
        out = np.zeros(n, dtype=bool)
        out[:max(n - offset, 0)] = kind[offset:] == k
        return out
    bars = seen & (kind == _K_BAR) & (np.arange(n) > 0)
    note_at = seen & (kind == _K_POS) & at(_K_VEL, 1) & at(_K_PITCH, 2) & at(_K_DUR, 3)
    chord_at = seen & (kind == _K_POS) & at(_K_CHORD, 1)
    bar_of = np.cumsum(bars)                                             # bars counted before (and at) each event

    i = np.nonzero(note_at)[0]
    start = bar_of[i] * ticks_per_bar + (ev[i] - POSITION) * ticks_per_bar // POSITION_RESOLUTION
    duration = (ev[i + 3] - DURATION + 1) * step
    notes = np.stack([VELOCITY_BINS[ev[i + 1] - VELOCITY], ev[i + 2] - PITCH, start, start + duration], axis=1) \
        if len(i) else np.zeros((0, 4), dtype=np.int64)
    c = np.nonzero(chord_at)[0]
    marker_times = bar_of[c] * ticks_per_bar + (ev[c] - POSITION) * ticks_per_bar // POSITION_RESOLUTION
    marker_texts = [CHORD_NAMES[w - CHORD] for w in ev[c + 1].tolist()]
    if not 0 <= key_index < len(KEY_NAMES):
        raise KeyError(key_index)
    return DecodedMidi(tempo=(bpm_word - META) * BPM_INTERVAL, numerator=numerator, denominator=denominator,
                       key_name=KEY_NAMES[key_index], notes=notes.astype(np.int64), marker_times=marker_times.astype(np.int64),
                       marker_texts=marker_texts, oov=oov)


def read_smf(data):
    """Minimal reader for the files `DecodedMidi.dump` writes (format 1, no running status): returns
    (ticks_per_beat, [track -> list of (absolute tick, status byte / meta type, payload bytes)])."""
    if data[:4] != b"MThd":
        raise ValueError("not a MIDI file")
    hlen, fmt, ntracks, tpb = struct.unpack(">IHHH", data[4:14])
    pos, tracks = 8 + hlen, []
    for _ in range(ntracks):
        if data[pos:pos + 4] != b"MTrk":
            raise ValueError("track chunk expected at byte %d" % pos)
        end = pos + 8 + struct.unpack(">I", data[pos + 4:pos + 8])[0]
        pos += 8
        now, out = 0, []
        while pos < end:
            delta = 0
            while True:
                byte = data[pos]
                pos += 1
                delta = (delta << 7) | (byte & 0x7F)
                if not byte & 0x80:
                    break
            now += delta
            status = data[pos]
            if status == 0xFF:
                kind = data[pos + 1]
                pos += 2
                length = 0
                while True:
                    byte = data[pos]
                    pos += 1
                    length = (length << 7) | (byte & 0x7F)
                    if not byte & 0x80:
                        break
                out.append((now, (0xFF, kind), data[pos:pos + length]))
                pos += length
            else:
                length = 1 if status & 0xF0 in (0xC0, 0xD0) else 2
                out.append((now, status, data[pos + 1:pos + 1 + length]))
                pos += 1 + length
        tracks.append(out)
    return tpb, tracks
