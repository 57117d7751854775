#!/usr/bin/env python
"""A/B of the reverb's radix-64 FFT passes: plain strided loads against bulk asynchronous staging
(B200DDSP_FFT_BULK=1: cp.async.bulk + mbarrier, csrc/reverb.cuh).  Times the stand-alone reverb
(b200ddsp_reverb: 3 forward + 3 inverse passes of 2^18 points, 4 + 4 of 2^19) with CUDA events, L2 warm
(the ping-pong buffers live there) and checks the result against the float64 convolution of one clip.

    B200DDSP_FFT_BULK=0 python scripts/reverb_ab.py; B200DDSP_FFT_BULK=1 python scripts/reverb_ab.py
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ddsp_piano_b200 as dp                               # noqa: E402
from ddsp_piano_b200.processors import _DEFAULT_CFG        # noqa: E402


def main():
    dev = torch.device('cuda:0')
    eng = dp.get_engine(dev, **{**_DEFAULT_CFG, 'sample_rate': 24000})
    out = {'bulk': int(os.environ.get('B200DDSP_FFT_BULK', '0'))}
    from scipy.signal import fftconvolve
    for B, N, L in ((16, 72000, 72000), (16, 144000, 144000), (8, 72000, 72000)):
        rng = np.random.default_rng(N)
        audio = (rng.standard_normal([B, N]) * 0.1).astype(np.float32)
        ir = (rng.standard_normal([B, L]) * np.exp(-6 * np.arange(L) / L) * 1e-2).astype(np.float32)
        a, h = torch.from_numpy(audio).to(dev), torch.from_numpy(ir).to(dev)
        for _ in range(5):
            y = eng.reverb(a, h)
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(50)]
        for e0, e1 in evs:
            e0.record()
            y = eng.reverb(a, h)
            e1.record()
        torch.cuda.synchronize()
        ms = sorted(e0.elapsed_time(e1) for e0, e1 in evs)
        hh = ir[0].astype(np.float64).copy()
        hh[0] = 0
        want = fftconvolve(audio[0].astype(np.float64), hh)[:N] + audio[0]
        err = float(np.max(np.abs(y[0].cpu().numpy() - want)) / np.max(np.abs(want)))
        out[f'B{B}_N{N}_L{L}'] = {'median_ms': ms[len(ms) // 2], 'min_ms': ms[0], 'rel_err_vs_float64': err}
    print(json.dumps(out))


if __name__ == '__main__':
    main()
