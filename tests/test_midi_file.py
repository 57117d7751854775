"""MIDI file front end (ddsp_piano_b200/midi.py, reference utils/io_utils.py:77-137) on files written
by the test itself.  note_seq / pretty_midi are absent: the conventions are restated (parity
unpinned) and what is tested here is the file format, the tempo map and the stated conventions."""
import struct

import numpy as np
import pytest

from ddsp_piano_b200 import midi


def vlq(n):
    out = [n & 0x7f]
    n >>= 7
    while n:
        out.append((n & 0x7f) | 0x80)
        n >>= 7
    return bytes(reversed(out))


def write_smf(path, tracks, division=480, fmt=1):
    """tracks: lists of (absolute tick, bytes of the event without delta time)."""
    chunks = b''
    for tr in tracks:
        body, last = b'', 0
        for tick, ev in sorted(tr, key=lambda e: e[0]):
            body += vlq(tick - last) + ev
            last = tick
        body += vlq(0) + b'\xff\x2f\x00'
        chunks += b'MTrk' + struct.pack('>I', len(body)) + body
    with open(path, 'wb') as f:
        f.write(b'MThd' + struct.pack('>IHHH', 6, fmt, len(tracks), division) + chunks)


def on(p, v, ch=0): return bytes([0x90 | ch, p, v])
def off(p, ch=0): return bytes([0x80 | ch, p, 0])
def cc(n, v, ch=0): return bytes([0xb0 | ch, n, v])
def tempo(us): return b'\xff\x51\x03' + us.to_bytes(3, 'big')


def test_read_midi_tempo_map_and_note_pairing(tmp_path):
    path = str(tmp_path / 'a.mid')
    # 480 ticks per quarter; 120 bpm for one quarter (0.5 s), then 60 bpm (1 s per quarter)
    write_smf(path, [[(0, tempo(500000)), (480, tempo(1000000))],
                     [(0, on(60, 100)), (480, off(60)), (480, on(64, 80)), (960, on(64, 0)),   # vel 0 = off
                      (240, cc(64, 127)), (1200, cc(64, 0))]])
    notes, ccs, end = midi.read_midi(path)
    assert [[round(x, 6) for x in n[:2]] + n[2:] for n in notes] == [[0.0, 0.5, 60, 100, 0], [0.5, 1.5, 64, 80, 0]]
    assert [[round(c[0], 6)] + c[1:] for c in ccs] == [[0.25, 64, 127, 0], [2.0, 64, 0, 0]]
    assert abs(end - 1.5) < 1e-9                       # total_time = the latest note end


def test_read_midi_pretty_midi_conventions(tmp_path):
    path = str(tmp_path / 'p.mid')
    # tempo events outside track 0 are ignored; a repeated tempo opens no new interval; a program change
    # makes a new instrument; control changes of a channel without notes are dropped
    write_smf(path, [[(0, tempo(500000)), (480, tempo(500000)), (960, tempo(250000))],
                     [(0, tempo(1000000)), (0, cc(64, 127)), (0, on(60, 100)), (480, off(60)),
                      (480, bytes([0xc0, 5])), (480, on(62, 90)), (1440, off(62)), (0, cc(64, 10, ch=3))],
                     [(0, on(60, 70, ch=1)), (960, off(60, ch=1))]])
    notes, ccs, end = midi.read_midi(path)
    assert notes == [[0.0, 0.5, 60, 100, 0], [0.5, 1.25, 62, 90, 1], [0.0, 1.0, 60, 70, 2]]
    assert ccs == [[0.0, 64, 127, 0]] and end == 1.25
    # seconds(tick) = start of the interval + scale * ticks, with pretty_midi's scale arithmetic
    scale = 60.0 / ((6e7 / 500000) * 480)
    assert notes[1][1] == scale * 960 + (60.0 / ((6e7 / 250000) * 480)) * 480
    # the pedal of instrument 0 holds only its own notes (note_seq keeps one pedal state per instrument)
    out, total = midi.apply_sustain_control_changes(notes, ccs, end)
    assert [n[1] for n in out] == [1.25, 1.25, 1.0] and total == 1.25


def test_read_midi_rejects_garbage(tmp_path):
    path = str(tmp_path / 'x.mid')
    with open(path, 'wb') as f:
        f.write(b'RIFF....')
    with pytest.raises(ValueError):
        midi.read_midi(path)
    with open(path, 'wb') as f:
        f.write(b'MThd' + struct.pack('>IHHH', 6, 1, 1, 0x8000 | 25) + b'MTrk' + struct.pack('>I', 0))
    with pytest.raises(ValueError, match='SMPTE'):
        midi.read_midi(path)
    with open(path, 'wb') as f:
        f.write(b'MThd' + struct.pack('>IHHH', 6, 1, 1, 480) + b'MTrk' + struct.pack('>I', 400) + b'\x00')
    with pytest.raises(ValueError, match='truncated'):
        midi.read_midi(path)
    with open(path, 'wb') as f:                          # an event cut short inside a well-formed chunk
        f.write(b'MThd' + struct.pack('>IHHH', 6, 1, 1, 480) + b'MTrk' + struct.pack('>I', 2) + b'\x00\x90')
    with pytest.raises(ValueError, match='truncated'):
        midi.read_midi(path)
    with open(path, 'wb') as f:
        f.write(b'MThd' + struct.pack('>IHHH', 6, 1, 1, 0) + b'MTrk' + struct.pack('>I', 0))
    with pytest.raises(ValueError, match='zero ticks'):
        midi.read_midi(path)


def test_written_notes_come_back(tmp_path):
    """Property: any set of notes without same-pitch overlaps, written as note-on / note-off pairs at any
    constant tempo, is read back with its ticks * scale times, pitches and velocities."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=40, deadline=None)
    @given(st.lists(st.tuples(st.integers(0, 5000), st.integers(1, 2000), st.integers(0, 127), st.integers(1, 127)),
                    min_size=1, max_size=30),
           st.integers(200000, 1500000), st.sampled_from([96, 240, 480, 960]))
    def check(raw, us, division):
        busy, kept = {}, []
        for start, length, pitch, vel in raw:
            if all(start + length <= s0 or start >= e0 for s0, e0 in busy.get(pitch, [])):
                busy.setdefault(pitch, []).append((start, start + length))
                kept.append((start, start + length, pitch, vel))
        path = str(tmp_path / 'h.mid')
        events = [(0, tempo(us))]
        for s0, e0, pitch, vel in kept:
            events += [(s0, on(pitch, vel)), (e0, off(pitch))]
        # write_smf sorts by tick (stable): a note-off on the tick of the next note-on of its pitch must come
        # first, as in a real file
        events.sort(key=lambda e: (e[0], 0 if e[1][0] & 0xf0 == 0x80 else 1))
        write_smf(path, [events], division=division)
        notes, ccs, total = midi.read_midi(path)
        scale = 60.0 / ((6e7 / us) * division)
        want = sorted((scale * s0, scale * e0, pitch, vel, 0) for s0, e0, pitch, vel in kept)
        assert sorted(tuple(n) for n in notes) == want
        assert ccs == [] and total == max(n[1] for n in want)

    check()


def test_running_status_and_format_0(tmp_path):
    path = str(tmp_path / 'b.mid')
    body = vlq(0) + on(60, 90) + vlq(240) + bytes([62, 70]) + vlq(240) + off(60) + vlq(0) + off(62) + \
        vlq(0) + b'\xff\x2f\x00'
    with open(path, 'wb') as f:
        f.write(b'MThd' + struct.pack('>IHHH', 6, 0, 1, 480) + b'MTrk' + struct.pack('>I', len(body)) + body)
    notes, _, _ = midi.read_midi(path)
    assert [(n[2], n[3], round(n[0], 6), round(n[1], 6)) for n in notes] == [(60, 90, 0.0, 0.5), (62, 70, 0.25, 0.5)]


def test_sustain_pedal_extends_notes():
    notes = [[0.0, 0.2, 60, 100], [0.5, 0.6, 60, 90], [0.1, 0.3, 64, 80]]
    ccs = [[0.05, 64, 127], [1.0, 64, 0]]
    out, total = midi.apply_sustain_control_changes(notes, ccs)
    by = {(n[2], n[0]): n[1] for n in out}
    assert by[(60, 0.0)] == 0.5            # rings on the pedal until the same pitch is struck again
    assert by[(60, 0.5)] == 1.0            # ... or until the pedal comes up
    assert by[(64, 0.1)] == 1.0
    assert total == 1.0
    out2, _ = midi.apply_sustain_control_changes(notes, [])
    assert sorted(n[1] for n in out2) == [0.2, 0.3, 0.6]


def test_pianoroll_conventions():
    notes = [[0.1, 0.2, 60, 127], [0.1004, 0.1008, 21, 64], [0.0, 1.0, 20, 100]]   # pitch 20 is out of range
    active, onset, ccs = midi.sequence_to_pianoroll(notes, [[0.3, 64, 127], [0.3, 67, 0]], 1.0, 250)
    assert active.shape == (251, 88) and ccs.shape == (251, 128)
    assert active[:, 60 - 21].nonzero()[0].tolist() == list(range(25, 50))
    assert active[:, 0].nonzero()[0].tolist() == [25]                              # at least one frame
    assert onset[:, 60 - 21].nonzero()[0].tolist() == [24, 25, 26] and onset[25, 39] == 1.0
    assert abs(onset[25, 0] - 64 / 127) < 1e-7
    assert ccs[75, 64] == 128 and ccs[75, 67] == 1 and ccs.sum() == 129            # value + 1 on the event frame


def test_load_midi_as_conditioning_shapes(tmp_path):
    path = str(tmp_path / 'c.mid')
    write_smf(path, [[(0, tempo(500000)), (0, on(60, 100)), (0, on(64, 100)), (960, off(60)), (960, off(64)),
                      (480, cc(64, 100)), (1440, cc(64, 0))]])
    x = midi.load_midi_as_conditioning(path, n_synths=16, frame_rate=250, warm_up_duration=0.5)
    assert x['conditioning'].shape == (1, 2 * 250 + 125, 16, 2) and x['pedal'].shape == (1, 625, 4)
    assert x['duration'] == 2.5
    pitches = x['conditioning'][0, 125 + 10, :, 0]
    assert sorted(pitches[pitches > 0].tolist()) == [60.0, 64.0]
    assert np.all(x['conditioning'][0, :125] == 0)                                  # warm-up padding
    held = x['conditioning'][0, 125 + 300, :, 0]                                    # 1.2 s: pedal still down
    assert sorted(held[held > 0].tolist()) == [60.0, 64.0]
    assert np.all(x['conditioning'][0, 125 + 380:, :, 0] == 0)                      # pedal up at 1.5 s
    assert x['pedal'][0, 125 + 125, 0] == pytest.approx(101 / 128)
    fixed = midi.load_midi_as_conditioning(path, duration=1.0)
    assert fixed['conditioning'].shape == (1, 250, 16, 2)


def _script():
    import importlib.util
    import os
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'scripts', 'synthesize_midi_file.py')
    spec = importlib.util.spec_from_file_location('synthesize_midi_file', path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_script_arguments_wav_and_normalisation(tmp_path):
    """scripts/synthesize_midi_file.py: the reference's argument list (synthesize_midi_file.py:12-36),
    the 16-bit wav writer and the dBFS normalisation (io_utils.py:245-253)."""
    import wave
    s = _script()
    a = s.process_args(['in.mid', 'out.wav'])
    assert (a.model, a.piano_type, a.warm_up, a.duration, a.normalize, a.unreverbed) == ('v2', 9, 0.5, None, None, False)
    a = s.process_args(['-m', 'dafx22', '-wu', '0', '-d', '3', '-n', '-20', '-u', '--piano_type', '2', 'in.mid', 'out.wav'])
    assert (a.model, a.piano_type, a.warm_up, a.duration, a.normalize, a.unreverbed) == ('dafx22', 2, 0.0, 3.0, -20.0, True)
    t = np.arange(2400, dtype=np.float32) / 24000
    x = (0.25 * np.sin(2 * np.pi * 440 * t)).astype(np.float32)
    y = s.normalize_dbfs(x, -20.0)
    assert abs(20 * np.log10(np.sqrt(np.mean(y.astype(np.float64) ** 2))) + 20.0) < 1e-4
    assert s.normalize_dbfs(np.zeros(8, np.float32), -20.0).tolist() == [0.0] * 8
    path = str(tmp_path / 'o.wav')
    s.write_wav(path, np.array([0.0, 0.5, -0.5, 2.0, -2.0], np.float32), 24000)
    with wave.open(path, 'rb') as f:
        assert (f.getnchannels(), f.getsampwidth(), f.getframerate(), f.getnframes()) == (1, 2, 24000, 5)
        assert np.frombuffer(f.readframes(5), '<i2').tolist() == [0, 16384, -16384, 32767, -32768]


@pytest.mark.gpu
def test_script_midi_file_to_wav(tmp_path):
    """The whole config-1 path as a user runs it: MIDI file -> wav (reverberated and dry) with the shipped
    v2 weights, piano 9, 0.5 s warm-up cut from the output."""
    import wave
    s = _script()
    mid, out = str(tmp_path / 'a4.mid'), str(tmp_path / 'a4.wav')
    write_smf(mid, [[(0, tempo(500000)), (0, on(69, 100)), (960, off(69))]])
    s.main(s.process_args(['-u', '-n', '-20', mid, out]))
    for path in (out, out + '_unreverbed.wav'):
        with wave.open(path, 'rb') as f:
            assert (f.getframerate(), f.getnframes()) == (24000, 24000)
            x = np.frombuffer(f.readframes(24000), '<i2').astype(np.float64) / 32768
        assert abs(20 * np.log10(np.sqrt(np.mean(x ** 2))) + 20.0) < 0.5
        n = 8192
        spec = np.abs(np.fft.rfft(x[2048:2048 + n] * np.hanning(n), 4 * n))
        assert abs(np.argmax(spec) * 24000 / (4 * n) - 440.0) < 4.0


def test_script_main_with_a_stub_model(tmp_path, monkeypatch):
    """The script's own logic (warm-up cut, the two output files, the piano id) around a stub model:
    no GPU needed."""
    import wave
    import torch
    import ddsp_piano_b200 as dp
    s = _script()
    seen = {}

    class Stub:
        def __call__(self, inputs):
            seen.update(inputs)
            n = inputs['conditioning'].shape[1] * 96
            wet = torch.full([1, n], 0.25)
            wet[0, :12000] = 1.0                         # the warm-up part must not reach the file
            return {'audio_synth': wet, 'add': {'signal': wet * 0.5}}

    monkeypatch.setattr(dp, 'maestro_v2_model', lambda *a, **k: Stub())
    mid, out = str(tmp_path / 'n.mid'), str(tmp_path / 'n.wav')
    write_smf(mid, [[(0, tempo(500000)), (0, on(60, 100)), (960, off(60))]])
    s.main(s.process_args(['-u', '--piano_type', '3', mid, out]))
    assert seen['piano_model'].tolist() == [[3]] and seen['conditioning'].shape == (1, 375, 16, 2)
    for path, level in ((out, 8192), (out + '_unreverbed.wav', 4096)):
        with wave.open(path, 'rb') as f:
            assert (f.getframerate(), f.getnframes()) == (24000, 24000)
            assert set(np.frombuffer(f.readframes(24000), '<i2').tolist()) == {level}
