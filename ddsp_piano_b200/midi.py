"""Host-side MIDI front end pieces.

``MIDIRoll2Conditioning`` mirrors ``ddsp_piano/utils/midi_encoders.py``: pianoroll
``[n_frames, 88, 2]`` (activity, onset velocity) -> polyphonic conditioning
``[n_frames, n_synths, 2]`` (pitch, velocity per channel) + polyphony.  The per-frame voice
allocation is a sequential scan; it runs in C++ inside ``libb200ddsp.so`` (host code, no GPU) and is
called with NumPy arrays.  Like the reference, one object = one allocation state = one piece:
``__call__`` starts from a fresh state.
"""
import ctypes
import struct

import numpy as np

from . import _lib


class MIDIRoll2Conditioning:
    def __init__(self, n_synths=16):
        self.n_synths = n_synths

    def __call__(self, roll):
        roll = np.ascontiguousarray(roll, dtype=np.float32)
        if roll.ndim != 3 or roll.shape[-1] != 2:
            raise ValueError(f'roll must be [n_frames, n_pitches, 2], got {roll.shape}')
        n_frames, n_pitches, _ = roll.shape
        if self.n_synths > n_pitches:
            raise ValueError(f'n_synths={self.n_synths} exceeds the {n_pitches} pitches of the roll')
        cond = np.empty([n_frames, self.n_synths, 2], np.float32)
        poly = np.empty([n_frames], np.float32)
        rc = _lib.load().b200ddsp_midi_roll_to_conditioning(
            roll.ctypes.data_as(ctypes.c_void_p), n_frames, n_pitches, self.n_synths,
            ctypes.c_float(21.0), cond.ctypes.data_as(ctypes.c_void_p),
            poly.ctypes.data_as(ctypes.c_void_p))
        if rc != 0:
            raise ValueError(f'b200ddsp_midi_roll_to_conditioning failed: {_lib.STATUS_NAMES.get(rc, rc)}')
        return cond, poly


# ---------------------------------------------------------------------------------------------
# MIDI file -> conditioning (reference utils/io_utils.py:77-137)
# ---------------------------------------------------------------------------------------------
# The reference reads the file with note_seq (midi_file_to_note_sequence = pretty_midi underneath),
# extends notes over the sustain pedal (note_seq.apply_sustain_control_changes) and rasterises with
# note_seq.sequences_lib.sequence_to_pianoroll.  note_seq / pretty_midi are not in the container, so
# their behaviour is RESTATED here from the published sources and is parity unpinned (DESIGN.md
# section 7): Standard MIDI File parsing is fixed by the file format; the restated conventions that
# matter for the model input (which track carries the tempo map, the float64 operation order of the
# tick -> seconds table, how notes are paired and grouped into instruments, the order in which equal-time
# events are processed, frame rounding) are spelled out where they are applied.
#
# A note is [start_s, end_s, pitch, velocity, instrument]; a control change is
# [time_s, number, value, instrument]; both lists are in note_seq's sequence order (instrument by
# instrument, each in file order).

def _vlq(data, pos):
    value = 0
    while True:
        b = data[pos]
        pos += 1
        value = (value << 7) | (b & 0x7f)
        if not b & 0x80:
            return value, pos


def _parse_tracks(data, path):
    """Standard MIDI File chunks -> (division, [[(absolute tick, kind, channel, d1, d2)] per track]);
    kind in 'on' / 'off' / 'cc' / 'program' / 'tempo' (tempo: d1 = microseconds per quarter)."""
    if data[:4] != b'MThd':
        raise ValueError(f'{path}: not a Standard MIDI File')
    hlen, _, n_tracks, division = struct.unpack('>IHHH', data[4:14])
    if division & 0x8000:
        raise ValueError(f'{path}: SMPTE time division is not supported')
    pos, tracks = 8 + hlen, []
    data_bytes = {0x80: 2, 0x90: 2, 0xa0: 2, 0xb0: 2, 0xc0: 1, 0xd0: 1, 0xe0: 2}
    for _ in range(n_tracks):
        if data[pos:pos + 4] != b'MTrk':
            raise ValueError(f'{path}: corrupt track chunk')
        tlen = struct.unpack('>I', data[pos + 4:pos + 8])[0]
        p, end, tick, status, events = pos + 8, pos + 8 + tlen, 0, 0, []
        if end > len(data):
            raise ValueError(f'{path}: truncated track chunk')
        while p < end:
            delta, p = _vlq(data, p)
            tick += delta
            b = data[p]
            if b == 0xff:                              # meta event
                kind = data[p + 1]
                n, p = _vlq(data, p + 2)
                if kind == 0x51 and n == 3:
                    events.append((tick, 'tempo', 0, int.from_bytes(data[p:p + 3], 'big'), 0))
                p += n
            elif b in (0xf0, 0xf7):                    # sysex
                n, p = _vlq(data, p + 1)
                p += n
            elif b >= 0xf1:                            # system common / real time (not valid in files)
                p += 1 + {0xf1: 1, 0xf2: 2, 0xf3: 1}.get(b, 0)
            else:
                if b & 0x80:
                    status = b
                    p += 1
                kind, channel = status & 0xf0, status & 0x0f
                if kind not in data_bytes:
                    raise ValueError(f'{path}: data byte without a running status')
                d1 = data[p]
                d2 = data[p + 1] if data_bytes[kind] == 2 else 0
                p += data_bytes[kind]
                if kind == 0x90 and d2 > 0:
                    events.append((tick, 'on', channel, d1, d2))
                elif kind == 0x80 or kind == 0x90:     # note-on with velocity 0 is a note-off
                    events.append((tick, 'off', channel, d1, 0))
                elif kind == 0xb0:
                    events.append((tick, 'cc', channel, d1, d2))
                elif kind == 0xc0:
                    events.append((tick, 'program', channel, d1, 0))
        tracks.append(events)
        pos = end
    return division, tracks


def read_midi(path):
    """Standard MIDI File (format 0/1, metrical time) -> (notes, control_changes, total_time), what
    note_seq.midi_file_to_note_sequence holds after pretty_midi.PrettyMIDI(file).  Conventions:

    * the tempo map is read from track 0 only; a tempo event at tick 0 replaces the 120 bpm default, a later
      one opens a new interval unless its scale repeats the last one; seconds(tick) =
      seconds(interval start) + scale * (tick - interval start) with scale = 60 / ((6e7 / us) * division),
      in float64 in that order (frame rounding downstream depends on the last bit);
    * a note-off closes every open note-on of its (channel, pitch) in its track that started on an earlier
      tick; notes struck on the very tick of the note-off stay open;
    * an instrument is a (program, channel, track) triple in order of first note END; control changes
      seen on a (channel, track) before its first note join that instrument, those of (channel, track)
      pairs without any note are dropped;
    * total_time = the latest note end (control changes do not extend it)."""
    with open(path, 'rb') as f:
        data = f.read()
    try:
        division, tracks = _parse_tracks(data, path)
    except (IndexError, struct.error) as e:            # an event or chunk header cut short
        raise ValueError(f'{path}: truncated MIDI data') from e
    if division == 0:
        raise ValueError(f'{path}: zero ticks per quarter note')
    scales = [(0, 60.0 / (120.0 * division))]
    for tick, kind, _, us, _ in (tracks[0] if tracks else []):
        if kind != 'tempo':
            continue
        scale = 60.0 / ((6e7 / us) * division)
        if tick == 0:
            scales = [(0, scale)]
        elif scale != scales[-1][1]:
            scales.append((tick, scale))
    starts, t = [], 0.0
    for i, (tk, scale) in enumerate(scales):
        if i:
            t = t + scales[i - 1][1] * (tk - scales[i - 1][0])
        starts.append(t)
    ticks_of = [tk for tk, _ in scales]

    def seconds(tick):
        import bisect
        k = bisect.bisect_right(ticks_of, tick) - 1
        return starts[k] + scales[k][1] * (tick - ticks_of[k])

    instruments, stragglers = {}, {}                    # (program, channel, track) -> [notes, ccs]

    def instrument(program, channel, track, create):
        key = (program, channel, track)
        if key in instruments:
            return instruments[key]
        if not create and (channel, track) in stragglers:
            return stragglers[(channel, track)]
        if create:
            instruments[key] = stragglers.pop((channel, track), None) or [[], []]
            return instruments[key]
        stragglers[(channel, track)] = [[], []]
        return stragglers[(channel, track)]

    for track, events in enumerate(tracks):
        open_notes, program = {}, [0] * 16
        for tick, kind, channel, d1, d2 in events:
            if kind == 'program':
                program[channel] = d1
            elif kind == 'on':
                open_notes.setdefault((channel, d1), []).append((tick, d2))
            elif kind == 'off':
                key = (channel, d1)
                if key in open_notes:
                    close = [(s, v) for s, v in open_notes[key] if s != tick]
                    keep = [(s, v) for s, v in open_notes[key] if s == tick]
                    for s, v in close:      # pretty_midi looks the instrument up per closed note only
                        instrument(program[channel], channel, track, True)[0].append(
                            [seconds(s), seconds(tick), d1, v])
                    if close and keep:
                        open_notes[key] = keep
                    else:
                        del open_notes[key]
            elif kind == 'cc':
                instrument(program[channel], channel, track, False)[1].append([seconds(tick), d1, d2])
    notes, ccs = [], []
    for index, (ns, cs) in enumerate(instruments.values()):
        notes += [n_ + [index] for n_ in ns]
        ccs += [c + [index] for c in cs]
    total_time = max([n_[1] for n_ in notes] + [0.0])
    return notes, ccs, total_time


def apply_sustain_control_changes(notes, control_changes, total_time=None, sustain_control_number=64):
    """note_seq.apply_sustain_control_changes: while the sustain pedal of its instrument (value >= 64) is
    down a note rings until the pedal is released or the same pitch is struck again on that instrument,
    whichever comes first.  Events at equal times are processed in the order sustain-on, sustain-off,
    note-on, note-off.  total_time grows with the notes ended by a pedal release; notes still ringing after
    the last event end there, and that time becomes total_time.  Returns (notes, total_time)."""
    ON, OFF, NOTE_ON, NOTE_OFF = 0, 1, 2, 3
    inst = lambda x, at: x[at] if len(x) > at else 0
    notes = [list(n_) for n_ in notes]
    if total_time is None:
        total_time = max([n_[1] for n_ in notes] + [0.0])
    events = [(n_[0], NOTE_ON, n_, inst(n_, 4)) for n_ in notes] + [(n_[1], NOTE_OFF, n_, inst(n_, 4)) for n_ in notes]
    for c in control_changes:
        if c[1] == sustain_control_number:
            events.append((c[0], ON if c[2] >= 64 else OFF, None, inst(c, 3)))
    events.sort(key=lambda e: (e[0], e[1]))
    active, sustain, removed, time = {}, {}, set(), 0
    for time, kind, note, i in events:
        held = active.setdefault(i, [])
        if kind == ON:
            sustain[i] = True
        elif kind == OFF:
            sustain[i] = False
            still = []
            for a in held:
                if a[1] < time:
                    a[1] = time                         # it was ringing on the pedal: ends now
                    total_time = max(total_time, time)
                else:
                    still.append(a)
            active[i] = still
        elif kind == NOTE_ON:
            if sustain.get(i, False):
                still = []
                for a in held:
                    if a[2] == note[2]:
                        a[1] = time                     # same pitch struck again
                        if a[0] == a[1]:
                            removed.add(id(a))
                    else:
                        still.append(a)
                active[i] = held = still
            held.append(note)
        elif not sustain.get(i, False):
            active[i] = [a for a in held if a is not note]
    for held in active.values():                        # still ringing after the last event
        for a in held:
            a[1] = time
            total_time = time
    return [n_ for n_ in notes if id(n_) not in removed], total_time


def sequence_to_pianoroll(notes, control_changes, total_time, frames_per_second=250, min_pitch=21,
                          max_pitch=108, onset_window=1, max_velocity=127.0):
    """note_seq.sequences_lib.sequence_to_pianoroll with its defaults (onset_mode='window'):
    a note is active from floor(start * fps) to ceil(end * fps); its onset velocity,
    velocity / max_velocity, is written on the onset frame +- onset_window frames, notes taken by
    start time (stable), later ones overwriting; control_changes[frame, number] = value + 1 on the
    frame of the event (0 = no event), in sequence order.
    Returns (active [T, 88], onset_velocities [T, 88], control_changes [T, 128])."""
    import math
    n_frames = int(total_time * frames_per_second + 1)
    n_pitches = max_pitch - min_pitch + 1
    active = np.zeros([n_frames, n_pitches], np.float32)
    onset_vel = np.zeros([n_frames, n_pitches], np.float32)
    ccs = np.zeros([n_frames, 128], np.int32)
    for n_ in sorted(notes, key=lambda n_: n_[0]):
        start, end, pitch, velocity = n_[:4]
        if pitch < min_pitch or pitch > max_pitch:
            continue
        s = int(start * frames_per_second)
        e = max(s, int(math.ceil(end * frames_per_second)))
        active[s:e, pitch - min_pitch] = 1.0
        o0, o1 = max(0, s - onset_window), min(n_frames, s + onset_window + 1)
        onset_vel[o0:o1, pitch - min_pitch] = velocity / max_velocity
    for c in control_changes:
        frame = int(c[0] * frames_per_second)
        if frame < n_frames:
            ccs[frame, c[1]] = c[2] + 1
    return active, onset_vel, ccs


def ensure_sequence_length(sequence, length, right=True):
    """utils/io_utils.py:204-224."""
    n = sequence.shape[0]
    if n == length:
        return sequence
    if n > length:
        return sequence[:length] if right else sequence[-length:]
    pad = [(0, length - n) if right else (length - n, 0)] + [(0, 0)] * (sequence.ndim - 1)
    return np.pad(sequence, pad_width=pad)


def load_midi_as_conditioning(mid_path, n_synths=16, frame_rate=250, duration=None, warm_up_duration=0.,
                              onset_window=1):
    """utils/io_utils.py:85-137: MIDI file -> {'conditioning' [1, F, n_synths, 2], 'pedal' [1, F, 4],
    'duration'}; the dict a PianoModel takes (plus 'piano_model')."""
    notes, ccs, total_time = read_midi(mid_path)
    notes, total_time = apply_sustain_control_changes(notes, ccs, total_time)   # :77-82
    active, onset_vel, cc_roll = sequence_to_pianoroll(notes, ccs, total_time, frame_rate, 21, 108,
                                                       onset_window)       # :104-107
    midi_roll = np.stack((active, onset_vel), axis=-1)                     # :109
    pedals = (cc_roll[:, 64:68] / 128.).astype(np.float32)                 # :110
    conditioning, _ = MIDIRoll2Conditioning(n_synths)(midi_roll)           # :113-114
    if duration is None:                                                   # :117-120
        target = int(np.ceil(total_time) * frame_rate)
    else:
        target = int(duration * frame_rate)
    conditioning = ensure_sequence_length(conditioning, target)
    pedals = ensure_sequence_length(pedals, target)
    if warm_up_duration > 0.:                                              # :127-130
        n = target + int(warm_up_duration * frame_rate)
        conditioning = ensure_sequence_length(conditioning, n, right=False)
        pedals = ensure_sequence_length(pedals, n, right=False)
    return {'conditioning': conditioning[np.newaxis, ...], 'pedal': pedals[np.newaxis, ...],
            'duration': target / frame_rate + warm_up_duration}
