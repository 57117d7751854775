"""A whole piece through the MIDI-file path on the GPU: a synthetic 60 s piano part (broken chords in both
hands, up to 10 sounding notes with the pedal, pedal changes every bar) written as a Standard MIDI File,
then load_midi_as_conditioning -> PianoModel (shipped weights, one clip of F = 15 125 frames) -> audio.
Prints the time of each stage for both shipped models.
usage: python scripts/long_piece_timing.py [seconds]"""
import os
import struct
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import ddsp_piano_b200 as dp
from ddsp_piano_b200 import midi

HERE = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def vlq(n):
    out = [n & 0x7f]
    n >>= 7
    while n:
        out.append((n & 0x7f) | 0x80)
        n >>= 7
    return bytes(reversed(out))


def write_piece(path, seconds, division=480):
    """120 bpm, 4/4: left hand one broken chord per bar in quavers, right hand semiquavers; pedal per bar."""
    rng = np.random.default_rng(0)
    chords = [(48, 52, 55, 60), (45, 48, 52, 57), (41, 45, 48, 53), (43, 47, 50, 55)]
    events, bar_ticks = [(0, b'\xff\x51\x03' + (500000).to_bytes(3, 'big'))], 4 * division
    for bar in range(int(seconds / 2.0)):
        t0, chord = bar * bar_ticks, chords[bar % 4]
        events.append((t0 + 10, bytes([0xb0, 64, 127])))
        events.append((t0 + bar_ticks - 30, bytes([0xb0, 64, 0])))
        for k in range(8):
            p = chord[k % 4]
            events.append((t0 + k * division // 2, bytes([0x90, p, int(rng.integers(50, 90))])))
            events.append((t0 + (k + 1) * division // 2 - 5, bytes([0x80, p, 0])))
        for k in range(16):
            p = chord[(k * 3) % 4] + 24 + (12 if k % 5 == 0 else 0)
            events.append((t0 + k * division // 4, bytes([0x90, p, int(rng.integers(60, 110))])))
            events.append((t0 + (k + 1) * division // 4 - 5, bytes([0x80, p, 0])))
    body, last = b'', 0
    for tick, ev in sorted(events, key=lambda e: e[0]):
        body += vlq(tick - last) + ev
        last = tick
    body += vlq(0) + b'\xff\x2f\x00'
    with open(path, 'wb') as f:
        f.write(b'MThd' + struct.pack('>IHHH', 6, 0, 1, division) + b'MTrk' + struct.pack('>I', len(body)) + body)


def main():
    seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
    path = os.path.join(tempfile.mkdtemp(), 'piece.mid')
    write_piece(path, seconds)
    t0 = time.perf_counter()
    inputs = midi.load_midi_as_conditioning(path, warm_up_duration=0.5)
    t_midi = time.perf_counter() - t0
    cond = inputs['conditioning'][0]
    print(f'MIDI file -> conditioning {list(inputs["conditioning"].shape)} in {t_midi * 1e3:.0f} ms (host); '
          f'polyphony max {int((cond[:, :, 0] > 0).sum(1).max())}, mean {(cond[:, :, 0] > 0).sum(1).mean():.1f}')
    for name, build, sr in (('dafx22 16 kHz', lambda: dp.dafx22_model(os.path.join(HERE, 'tests/golden/dafx22_weights.npz'), device='cuda:0'), 16000),
                            ('maestro-v2 24 kHz', lambda: dp.maestro_v2_model(os.path.join(HERE, 'tests/golden/v2_weights.npz'), device='cuda:0'), 24000)):
        model = build()
        inputs['piano_model'] = np.array([[min(9, model.n_instruments - 1)]], np.int64)   # the dafx22 fixture keeps 2 IRs
        out = model(inputs)
        torch.cuda.synchronize()
        tc, ts = [], []
        for _ in range(3):
            t0 = time.perf_counter()
            f = model.compute_controls(inputs)
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            out = model.processor_group(f, return_outputs_dict=True)
            torch.cuda.synchronize()
            t2 = time.perf_counter()
            tc.append(t1 - t0)
            ts.append(t2 - t1)
        audio = out['signal']
        assert audio.shape[1] == int(round(inputs['duration'] * sr)) and bool(torch.isfinite(audio).all())
        c, s = np.median(tc) * 1e3, np.median(ts) * 1e3
        print(f'{name}: control-rate graph {c:.1f} ms + synthesis {s:.1f} ms = {c + s:.1f} ms for '
              f'{inputs["duration"]:.1f} s of audio ({inputs["duration"] / ((c + s) * 1e-3):.0f} x real time); '
              f'peak |audio| {float(audio.abs().max()):.3f}, rms {float(audio.pow(2).mean().sqrt()):.4f}')


if __name__ == '__main__':
    main()
