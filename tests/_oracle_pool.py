"""Test infrastructure: the oracle's polyphonic forward with the (voice, clip) units spread over the host
cores (multiprocessing, spawn), for parity checks at BASELINE's full sizes.  Same arithmetic and the same
summation order as oracle/ddsp_piano_np.py::polyphonic_forward."""
import multiprocessing as mp
import os

import numpy as np


def _voice(args):
    (sr, amp, hd, inh, f0, mags, noise) = args
    from oracle import ddsp_piano_np as ref
    c = ref.additive_controls(amp, hd, inh, f0, sample_rate=sr)
    a = ref.additive_signal(**c, sample_rate=sr, inference=True)
    n = ref.noise_signal(ref.noise_controls(mags)['magnitudes'], noise)
    return a, n


def polyphonic_forward_clips(x, noises, clips, sr, reverb=True):
    """x: stacked [P, B, F, C] controls (+ reverb_ir [B, L]); noises: list of [B, N] per voice.
    Returns {clip: (dry [N], wet [N] or None)} for the requested clips."""
    from oracle import ddsp_piano_np as ref
    P = x['f0_hz'].shape[0]
    jobs = [(sr, x['amplitudes'][v][b:b + 1], x['harmonic_distribution'][v][b:b + 1], x['inharm_coef'][v][b:b + 1],
             x['f0_hz'][v][b:b + 1], x['magnitudes'][v][b:b + 1], noises[v][b:b + 1])
            for b in clips for v in range(P)]
    with mp.get_context('spawn').Pool(min(os.cpu_count() or 1, len(jobs))) as pool:
        parts = pool.map(_voice, jobs, chunksize=1)
    out = {}
    for i, b in enumerate(clips):
        dry = None
        for v in range(P):
            a, n = parts[i * P + v]
            dry = (n + a) if dry is None else (dry + n) + a          # polyphonic_dag.py:28-37
        wet = ref.reverb_signal(dry, x['reverb_ir'][b:b + 1]) if reverb else None
        out[b] = (dry[0], None if wet is None else wet[0])
    return out
