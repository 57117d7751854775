import time, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
import ddsp_piano_b200 as dp
w = bench.WORKLOADS['full']
dev = torch.device('cuda:0')
x = {k: torch.from_numpy(v).to(dev) for k, v in bench.synthetic_inputs(w, 0).items()}
P = w['P']
additive = dp.MultiInharmonic(frame_rate=250, sample_rate=w['sr'], inference=True, name='additive')
noise = dp.DynamicSizeFilteredNoise(frame_rate=250, sample_rate=w['sr'], name='noise', seed=1)
group = dp.ProcessorGroup(dag=dp.polyphonic_dag(additive=additive, noise=noise, reverb=dp.Reverb(),
    additive_controls=['amplitudes', 'harmonic_distribution', 'inharm_coef', 'f0_hz'],
    noise_controls=['magnitudes'], reverb_controls=['reverb_ir'], n_synths=P))
def feats():
    f = {f'{k}_{v}': x[k][v] for k in ('amplitudes', 'harmonic_distribution', 'inharm_coef', 'f0_hz', 'magnitudes') for v in range(P)}
    f['reverb_ir'] = x['reverb_ir']
    return f
for _ in range(5): group(feats())
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(50): f = feats()
t1 = time.perf_counter()
print('features dict build us', (t1 - t0) / 50 * 1e6)
ts = []
for _ in range(20):
    f = feats(); torch.cuda.synchronize()
    t0 = time.perf_counter(); group(f); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    ts.append(((t1 - t0) * 1e6, (t2 - t0) * 1e6))
import numpy as np
print('call returns after us (median)', np.median([a for a, b in ts]), 'complete after us', np.median([b for a, b in ts]))
