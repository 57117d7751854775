"""BASELINE config 1 on the GPU: one 3 s clip (+0.5 s warm-up), 4 sustained notes on 16 voice channels,
the shipped dafx22 weights at 16 kHz (and the v2 weights at 24 kHz) from MIDI conditioning to audio.
usage: python scripts/config1_timing.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ddsp_piano_b200 as dp

HERE = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
F, P = 875, 16
cond = np.zeros([1, F, P, 2], np.float32)
for v, pitch in enumerate((48, 60, 64, 67)):            # notes on at frame 0 (+warm-up), off at 625
    cond[0, 125:750, v, 0] = pitch
    cond[0, 125, v, 1] = 80 / 127
feats = {'conditioning': cond, 'pedal': np.zeros([1, F, 4], np.float32), 'piano_model': np.zeros([1, 1], np.int64)}
for name, build, sr in (('dafx22 (16 kHz, H96, M64, 1.5 s IR)', lambda: dp.dafx22_model(os.path.join(HERE, 'tests/golden/dafx22_weights.npz'), device='cuda:0'), 16000),
                        ('maestro-v2 (24 kHz, H128, M96, FDN IR 2 s)', lambda: dp.maestro_v2_model(os.path.join(HERE, 'tests/golden/v2_weights.npz'), device='cuda:0'), 24000)):
    model = build()
    for _ in range(3):
        out = model(feats)
    torch.cuda.synchronize()
    ts, tc = [], []
    for _ in range(10):
        t0 = time.perf_counter()
        f = model.compute_controls(feats)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        out = model.processor_group(f, return_outputs_dict=True)
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        tc.append(t1 - t0); ts.append(t2 - t1)
    audio_s = F / 250.0
    c, s = np.median(tc) * 1e3, np.median(ts) * 1e3
    print(f'{name}: control-rate graph {c:.2f} ms + synthesis {s:.2f} ms = {c + s:.2f} ms for {audio_s:.1f} s of audio '
          f'({audio_s / ((c + s) * 1e-3):.0f} x real time), peak |audio| {float(out["signal"].abs().max()):.3f}')
