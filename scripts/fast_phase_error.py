#!/usr/bin/env python
"""Error table of `fast_phase` (b200ddsp_config.fast_phase = 1): the additive synth with closed-form
double-precision unit start phases against
  ideal   oracle additive_signal_exact_sum: the reference's signal model (its float32 controls, its legacy resize
          coordinates, its window cross-fade) evaluated in exact arithmetic (float64 omegas, sums, cosines)
  ref32   oracle additive_signal in float32: the reference's own arithmetic -- float32 omegas (the same rounded
          value added thousands of times for a held partial) and the float32 angular_cumsum
beside the bit-faithful kernels against ref32, and ref32 against ideal (= how much of the reference's output is
float32 rounding noise of its own phase).  max|y - ref| / max|ref| per case.

    python scripts/fast_phase_error.py > profiles/r02_fast_phase_error_table.txt      (needs a B200)
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import ddsp_piano_b200 as dp                      # noqa: E402
from oracle import ddsp_piano_np as ref           # noqa: E402  (checker)
from test_gpu_parity import voice_inputs          # noqa: E402


def rel(a, b):
    return float(np.max(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64))) / np.max(np.abs(b)))


def case(name, sr, ctl32):
    dev = torch.device('cuda:0')
    exact = ref.additive_signal_exact_sum(**ctl32, sample_rate=sr)      # the ideal model
    ref32 = ref.additive_signal(**ctl32, sample_rate=sr, inference=True)
    t = {k: torch.from_numpy(v).to(dev) for k, v in ctl32.items()}
    out = {}
    for fast in (False, True):
        synth = dp.MultiInharmonic(frame_rate=250, sample_rate=sr, inference=True, fast_phase=fast, name='additive')
        out[fast] = synth.get_signal(**t).cpu().numpy()
    print(f'{name:44s} {rel(out[True], exact):10.2e} {rel(out[True], ref32):12.2e} {rel(out[False], ref32):12.2e} '
          f'{rel(ref32, exact):12.2e}')


def main():
    print(f'{"case":44s} {"fast/ideal":>10s} {"fast/ref32":>12s} {"faithful/ref32":>12s} {"ref32/ideal":>12s}')
    gold = os.path.join(ROOT, 'tests', 'golden')
    for name in ('additive_24k_inference', 'additive_48k_h128', 'additive_24k_single_string'):
        g = dict(np.load(os.path.join(gold, name + '.npz')))
        ctl = {k[4:]: g[k] for k in ('ctl_amplitudes', 'ctl_harmonic_distribution', 'ctl_harmonic_shifts', 'ctl_f0_hz')}
        case('golden ' + name, int(g['sample_rate']), ctl)
    for sr, F, B, H, S in ((24000, 750, 2, 96, 2), (48000, 750, 1, 128, 2), (24000, 750, 2, 128, 1), (16000, 750, 2, 96, 2)):
        x = voice_inputs(np.random.default_rng(sr + H), B, F, H, S, 8)
        ctl = ref.additive_controls(x['amplitudes'], x['harmonic_distribution'], x['inharm_coef'], x['f0_hz'],
                                    sample_rate=sr)
        case(f'3 s clip with onsets sr={sr} H={H} S={S} B={B}', sr, ctl)
    import bench
    w = bench.WORKLOADS['full']
    x = bench.synthetic_inputs(w, seed=0)
    for v in (0, 7):
        ctl = ref.additive_controls(x['amplitudes'][v][5:6], x['harmonic_distribution'][v][5:6], x['inharm_coef'][v][5:6],
                                    x['f0_hz'][v][5:6], sample_rate=w['sr'])
        case(f'configs[2] clip 5 voice {v}', w['sr'], ctl)
    w = bench.WORKLOADS['stress']
    x = bench.synthetic_inputs(w, seed=3, B=2)
    ctl = ref.additive_controls(x['amplitudes'][3][:1], x['harmonic_distribution'][3][:1], x['inharm_coef'][3][:1],
                                x['f0_hz'][3][:1], sample_rate=w['sr'])
    case('configs[4] (48 kHz, H128) one voice-clip', w['sr'], ctl)


if __name__ == '__main__':
    main()
