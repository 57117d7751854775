"""Print max|y - oracle| / max|oracle| of the additive synth on a few shapes (development aid).
usage: python scripts/additive_error.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np, torch
import ddsp_piano_b200 as dp
from oracle import ddsp_piano_np as ref
from test_gpu_parity import voice_inputs, rel_err, cu
dev = torch.device('cuda:0')
for sr, F, B, H, S in [(24000, 250, 2, 96, 2), (16000, 190, 1, 96, 2), (48000, 60, 1, 128, 2), (24000, 42, 1, 20, 4)]:
    rng = np.random.default_rng(sr + F)
    x = voice_inputs(rng, B, F, H, S, 8)
    ctl = ref.additive_controls(x['amplitudes'], x['harmonic_distribution'], x['inharm_coef'], x['f0_hz'], sample_rate=sr)
    want = ref.additive_signal(**ctl, sample_rate=sr, inference=True)
    synth = dp.MultiInharmonic(frame_rate=250, sample_rate=sr, inference=True, name='additive')
    got = synth(cu(x['amplitudes'], dev), cu(x['harmonic_distribution'], dev), cu(x['inharm_coef'], dev), cu(x['f0_hz'], dev))
    print(f'sr={sr} F={F} B={B} H={H} S={S}: rel err {rel_err(got, want):.3e}')
