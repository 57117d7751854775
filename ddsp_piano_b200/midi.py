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
