"""Host-side MIDI front end pieces.

``MIDIRoll2Conditioning`` mirrors ``ddsp_piano/utils/midi_encoders.py``: pianoroll
``[n_frames, 88, 2]`` (activity, onset velocity) -> polyphonic conditioning
``[n_frames, n_synths, 2]`` (pitch, velocity per channel) + polyphony.  The per-frame voice
allocation is a sequential scan; it runs in C++ inside ``libb200ddsp.so`` (host code, no GPU) and is
called with NumPy arrays.  Like the reference, one object = one allocation state = one piece:
``__call__`` starts from a fresh state.
"""
import ctypes

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
# section 7): Standard MIDI File parsing and the tempo map are fixed by the file format; the
# restated conventions that matter for the model input are spelled out where they are applied.

def _vlq(data, pos):
    value = 0
    while True:
        b = data[pos]
        pos += 1
        value = (value << 7) | (b & 0x7f)
        if not b & 0x80:
            return value, pos


def read_midi(path):
    """Standard MIDI File (format 0/1, metrical time) -> (notes, control_changes, end_time):
    notes = [[start_s, end_s, pitch, velocity]], control_changes = [[time_s, number, value]].
    pretty_midi conventions: note-on with velocity 0 is a note-off; a note-off closes every open
    note-on of its (channel, pitch) that started on an earlier tick; tempo events of any track
    apply to all tracks; end_time = the last note end / control change."""
    import struct
    with open(path, 'rb') as f:
        data = f.read()
    if data[:4] != b'MThd':
        raise ValueError(f'{path}: not a Standard MIDI File')
    hlen, fmt, n_tracks, division = struct.unpack('>IHHH', data[4:14])
    if division & 0x8000:
        raise ValueError('SMPTE time division is not supported')
    pos = 8 + hlen
    events, tempi = [], [(0, 500000)]                  # (tick, kind, a, b, c); default 120 bpm
    for _ in range(n_tracks):
        if data[pos:pos + 4] != b'MTrk':
            raise ValueError('corrupt track chunk')
        tlen = struct.unpack('>I', data[pos + 4:pos + 8])[0]
        p, end, tick, status = pos + 8, pos + 8 + tlen, 0, 0
        while p < end:
            delta, p = _vlq(data, p)
            tick += delta
            b = data[p]
            if b == 0xff:                              # meta event
                kind = data[p + 1]
                n, p = _vlq(data, p + 2)
                if kind == 0x51 and n == 3:
                    tempi.append((tick, int.from_bytes(data[p:p + 3], 'big')))
                p += n
            elif b in (0xf0, 0xf7):                    # sysex
                n, p = _vlq(data, p + 1)
                p += n
            else:
                if b & 0x80:
                    status = b
                    p += 1
                kind, channel = status & 0xf0, status & 0x0f
                if kind in (0xc0, 0xd0):
                    p += 1
                else:
                    d1, d2 = data[p], data[p + 1]
                    p += 2
                    if kind == 0x90 and d2 > 0:
                        events.append((tick, 'on', channel, d1, d2))
                    elif kind == 0x80 or kind == 0x90:
                        events.append((tick, 'off', channel, d1, 0))
                    elif kind == 0xb0:
                        events.append((tick, 'cc', channel, d1, d2))
        pos = end
    # tempo map: seconds at each tempo change
    tempi = sorted(set(tempi))
    marks, t = [], 0.0
    for i, (tk, us) in enumerate(tempi):
        if i:
            t += (tk - tempi[i - 1][0]) * tempi[i - 1][1] * 1e-6 / division
        marks.append((tk, t, us))

    def seconds(tick):
        k = 0
        for i, m in enumerate(marks):
            if m[0] <= tick:
                k = i
        tk, t0, us = marks[k]
        return t0 + (tick - tk) * us * 1e-6 / division

    notes, ccs, open_notes = [], [], {}
    for tick, kind, channel, d1, d2 in sorted(events, key=lambda e: e[0]):   # stable: file order per tick
        if kind == 'on':
            open_notes.setdefault((channel, d1), []).append((tick, d2))
        elif kind == 'off':
            key = (channel, d1)
            if key in open_notes:
                close = [(s, v) for s, v in open_notes[key] if s != tick]
                keep = [(s, v) for s, v in open_notes[key] if s == tick]
                for s, v in close:
                    notes.append([seconds(s), seconds(tick), d1, v])
                if close and keep:
                    open_notes[key] = keep
                else:
                    del open_notes[key]
        else:
            ccs.append([seconds(tick), d1, d2])
    notes.sort(key=lambda n_: n_[0])
    end_time = max([n_[1] for n_ in notes] + [c[0] for c in ccs] + [0.0])
    return notes, ccs, end_time


def apply_sustain_control_changes(notes, control_changes, sustain_control_number=64):
    """note_seq.apply_sustain_control_changes: while the sustain pedal (value >= 64) is down a note
    rings until the pedal is released or the same pitch is struck again, whichever comes first.
    Events at equal times are processed in the order sustain-on, sustain-off, note-on, note-off.
    Returns (notes, total_time)."""
    ON, OFF, NOTE_ON, NOTE_OFF = 0, 1, 2, 3
    notes = [list(n_) for n_ in notes]
    events = []
    for n_ in notes:
        events.append((n_[0], NOTE_ON, n_))
        events.append((n_[1], NOTE_OFF, n_))
    for t, number, value in control_changes:
        if number == sustain_control_number:
            events.append((t, ON if value >= 64 else OFF, None))
    events.sort(key=lambda e: (e[0], e[1]))
    active, sustain, removed, time = [], False, set(), 0.0
    for time, kind, note in events:
        if kind == ON:
            sustain = True
        elif kind == OFF:
            sustain = False
            still = []
            for a in active:
                if a[1] < time:
                    a[1] = time                         # it was ringing on the pedal: ends now
                else:
                    still.append(a)
            active = still
        elif kind == NOTE_ON:
            if sustain:
                still = []
                for a in active:
                    if a[2] == note[2]:
                        a[1] = time                     # same pitch struck again
                        if a[0] == a[1]:
                            removed.add(id(a))
                    else:
                        still.append(a)
                active = still
            active.append(note)
        else:
            if not sustain and any(a is note for a in active):
                active = [a for a in active if a is not note]
    for a in active:                                    # still ringing at the end of the piece
        a[1] = max(a[1], time)
    notes = [n_ for n_ in notes if id(n_) not in removed]
    total = max([n_[1] for n_ in notes] + [c[0] for c in control_changes] + [0.0])
    return notes, total


def sequence_to_pianoroll(notes, control_changes, total_time, frames_per_second=250, min_pitch=21,
                          max_pitch=108, onset_window=1, max_velocity=127.0):
    """note_seq.sequences_lib.sequence_to_pianoroll with its defaults (onset_mode='window'):
    a note is active from floor(start * fps) to ceil(end * fps) (at least one frame); its onset
    velocity, velocity / max_velocity, is written on the onset frame +- onset_window frames;
    control_changes[frame, number] = value + 1 on the frame of the event (0 = no event).
    Returns (active [T, 88], onset_velocities [T, 88], control_changes [T, 128])."""
    import math
    n_frames = int(total_time * frames_per_second + 1)
    n_pitches = max_pitch - min_pitch + 1
    active = np.zeros([n_frames, n_pitches], np.float32)
    onset_vel = np.zeros([n_frames, n_pitches], np.float32)
    ccs = np.zeros([n_frames, 128], np.int32)
    for start, end, pitch, velocity in sorted(notes, key=lambda n_: n_[0]):
        if pitch < min_pitch or pitch > max_pitch:
            continue
        s = int(start * frames_per_second)
        e = max(s + 1, int(math.ceil(end * frames_per_second)))
        active[s:e, pitch - min_pitch] = 1.0
        o0, o1 = max(0, s - onset_window), min(n_frames, s + onset_window + 1)
        onset_vel[o0:o1, pitch - min_pitch] = velocity / max_velocity
    for t, number, value in sorted(control_changes, key=lambda c: c[0]):
        frame = int(t * frames_per_second)
        if frame < n_frames:
            ccs[frame, number] = value + 1
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
    notes, ccs, _ = read_midi(mid_path)
    notes, total_time = apply_sustain_control_changes(notes, ccs)          # :77-82
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
