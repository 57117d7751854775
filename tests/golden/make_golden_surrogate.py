#!/usr/bin/env python
"""Golden vectors for SurrogateAdditive, produced like make_golden.py: the
reference's own ``ddsp_piano/modules/surrogate_synth.py`` (and the ``inharm_synth.py`` it imports)
is EXECUTED unmodified over the NumPy stand-ins for tensorflow/gin/ddsp under oracle/tf_shim.

    python tests/golden/make_golden_surrogate.py     # writes tests/golden/surrogate_*.npz
"""
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import REF, load_reference_modules, make_voice_inputs  # noqa: E402


def main():
    load_reference_modules()
    full = 'ddsp_piano.modules.surrogate_synth'
    spec = importlib.util.spec_from_file_location(
        full, os.path.join(REF, 'ddsp_piano', 'modules', 'surrogate_synth.py'))
    sur = importlib.util.module_from_spec(spec)
    sys.modules[full] = sur
    spec.loader.exec_module(sur)
    rng = np.random.default_rng(20221018)
    for name, sr, F, B, H, onsets in (('surrogate_16k', 16000, 40, 2, 96, True),
                                      ('surrogate_24k_h40', 24000, 25, 1, 40, False)):
        x = make_voice_inputs(rng, B, F, H, 1, 8, onsets=onsets)
        decays = rng.uniform(0.97, 1.02, [B, F, H]).astype(np.float32)        # some above 1: clipped
        decays[:, :, ::7] *= -1.0                                             # sign is ignored (abs)
        # frames since the last onset, as SurrogateModule produces it (an integer count per frame)
        decay_time = np.broadcast_to((np.arange(F) % 13).astype(np.float32)[None, :, None], [B, F, 1]).copy()
        synth = sur.SurrogateAdditive(frame_rate=250, sample_rate=sr, inference=True, name='inharmonic')
        ctl = synth.get_controls(x['amplitudes'], decays, decay_time, x['harmonic_distribution'],
                                 x['inharm_coef'], x['f0_hz'])
        sig = synth.get_signal(**ctl)
        out = {f'in_{k}': v for k, v in x.items() if k != 'magnitudes'}
        out.update(in_decays=decays, in_decay_time=decay_time, sample_rate=np.int64(sr),
                   signal=np.asarray(sig, np.float32))
        out.update({f'ctl_{k}': np.asarray(v, np.float32) for k, v in ctl.items()})
        path = os.path.join(HERE, name + '.npz')
        np.savez_compressed(path, **out)
        print(f'{name}: {os.path.getsize(path) / 1024:.1f} KiB, max|signal| = {np.max(np.abs(sig)):.4f}')


if __name__ == '__main__':
    main()
