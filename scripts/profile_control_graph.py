import os, sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import ddsp_piano_b200 as dp
import bench
dev = torch.device('cuda:0')
model = dp.dafx22_model(os.path.join('/root/repo', 'tests/golden/dafx22_weights.npz'), device=dev)
B, F, P = 16, 750, 16
rng = np.random.default_rng(0)
cond = np.zeros([B, F, P, 2], np.float32)
for b in range(B):
    for v in range(P):
        k = 0
        while k < F:
            seg = int(rng.integers(40, 200))
            cond[b, k:k + seg, v, 0] = rng.integers(21, 109)
            cond[b, k, v, 1] = rng.uniform(0.2, 1.0)
            k += seg
x = {'conditioning': cond, 'pedal': np.zeros([B, F, 4], np.float32), 'piano_model': np.zeros([B, 1], np.int64)}
for _ in range(3):
    f = model.compute_controls(x)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    f = model.compute_controls(x)
torch.cuda.synchronize()
print('compute_controls ms', (time.perf_counter() - t0) / 5 * 1e3)
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        f = model.compute_controls(x)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=25, max_name_column_width=60))
