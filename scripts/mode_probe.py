"""Is the 1.03 / 1.10 ms mode of the synthesis stage tied to the process or to the engine (its streams)?
Builds several engines in one process (a new handle = new streams, new workspace) and times each."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
import ddsp_piano_b200 as dp
w = bench.WORKLOADS['full']
dev = torch.device('cuda:0')
x = {k: torch.from_numpy(v).to(dev) for k, v in bench.synthetic_inputs(w, 0).items()}
P = w['P']
f = {f'{k}_{v}': x[k][v] for k in ('amplitudes', 'harmonic_distribution', 'inharm_coef', 'f0_hz', 'magnitudes') for v in range(P)}
f['reverb_ir'] = x['reverb_ir']
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timed(fn, n=30):
    tot = 0.0
    for _ in range(n):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / n
for i in range(8):
    additive = dp.MultiInharmonic(frame_rate=250, sample_rate=w['sr'], inference=True, name='additive',
                                  min_frequency=20.0 + 1e-3 * i)
    noise = dp.DynamicSizeFilteredNoise(frame_rate=250, sample_rate=w['sr'], name='noise', seed=1)
    group = dp.ProcessorGroup(dag=dp.polyphonic_dag(additive=additive, noise=noise, reverb=dp.Reverb(),
        additive_controls=['amplitudes', 'harmonic_distribution', 'inharm_coef', 'f0_hz'],
        noise_controls=['magnitudes'], reverb_controls=['reverb_ir'], n_synths=P))
    for _ in range(5): group(dict(f))
    torch.cuda.synchronize()
    a = timed(lambda: group(dict(f)))
    b = timed(lambda: group(dict(f)))
    print(f'engine {i}: {a:.3f} {b:.3f} ms/step')
