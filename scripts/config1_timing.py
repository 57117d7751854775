"""BASELINE config 1 on the GPU: one 3 s clip (+0.5 s warm-up), 4 sustained notes on 16 voice channels,
the shipped dafx22 weights at 16 kHz (and the v2 weights at 24 kHz) from MIDI conditioning to audio.
usage: python scripts/config1_timing.py      (bench.py reports the same numbers under "config1")"""
import os
import sys
import time

HERE = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if HERE not in sys.path:
    sys.path.insert(0, HERE)
import numpy as np
import torch

F, P = 875, 16
MODELS = (('dafx22', 'dafx22 (16 kHz, H96, M64, 1.5 s IR)', 'dafx22_model', 'tests/golden/dafx22_weights.npz'),
          ('maestro_v2', 'maestro-v2 (24 kHz, H128, M96, FDN IR 2 s)', 'maestro_v2_model', 'tests/golden/v2_weights.npz'))


def features():
    cond = np.zeros([1, F, P, 2], np.float32)
    for v, pitch in enumerate((48, 60, 64, 67)):        # notes on at frame 0 (+warm-up), off at 625
        cond[0, 125:750, v, 0] = pitch
        cond[0, 125, v, 1] = 80 / 127
    return {'conditioning': cond, 'pedal': np.zeros([1, F, 4], np.float32), 'piano_model': np.zeros([1, 1], np.int64)}


def measure(device='cuda:0', reps=10):
    """{model: {control_rate_ms, synthesis_ms, total_ms, audio_s, rtf, peak}}: wall clock around synchronised
    calls, median of `reps` after 3 warm-up forwards."""
    import ddsp_piano_b200 as dp
    feats, results = features(), {}
    for key, name, factory, weights in MODELS:
        model = getattr(dp, factory)(os.path.join(HERE, weights), device=device)
        for _ in range(3):
            out = model(feats)
        torch.cuda.synchronize()
        ts, tc = [], []
        for _ in range(reps):
            t0 = time.perf_counter()
            f = model.compute_controls(feats)
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            out = model.processor_group(f, return_outputs_dict=True)
            torch.cuda.synchronize()
            t2 = time.perf_counter()
            tc.append(t1 - t0)
            ts.append(t2 - t1)
        audio_s = F / 250.0
        c, s = float(np.median(tc) * 1e3), float(np.median(ts) * 1e3)
        results[key] = {'what': name, 'control_rate_ms': c, 'synthesis_ms': s, 'total_ms': c + s, 'audio_s': audio_s,
                        'rtf': audio_s / ((c + s) * 1e-3), 'peak': float(out['signal'].abs().max())}
    return results


if __name__ == '__main__':
    for r in measure().values():
        print(f"{r['what']}: control-rate graph {r['control_rate_ms']:.2f} ms + synthesis {r['synthesis_ms']:.2f} ms = "
              f"{r['total_ms']:.2f} ms for {r['audio_s']:.1f} s of audio ({r['rtf']:.0f} x real time), "
              f"peak |audio| {r['peak']:.3f}")
