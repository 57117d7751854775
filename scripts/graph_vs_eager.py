"""Eager launches vs CUDA-graph replay of one ProcessorGroup forward (config 3, resident inputs):
how much of the step is launch gaps?  usage: python scripts/graph_vs_eager.py"""
import os, sys
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
f = {f'{k}_{v}': x[k][v] for k in ('amplitudes', 'harmonic_distribution', 'inharm_coef', 'f0_hz', 'magnitudes') for v in range(P)}
f['reverb_ir'] = x['reverb_ir']
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(5): group(dict(f))
torch.cuda.synchronize()
def timed(fn, n=20):
    tot = 0.0
    for _ in range(n):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / n
print('eager  ms/step', timed(lambda: group(dict(f))))
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    group(dict(f))
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    out = group(dict(f))
g.replay(); torch.cuda.synchronize()
print('graph  ms/step', timed(lambda: g.replay()))
print('eager  ms/step', timed(lambda: group(dict(f))))
