"""Run under torchrun (one rank per GPU): the config-4 timeline reverb with the NCCL neighbour
exchange, checked against the float64 convolution of the concatenated timeline.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/multi_gpu_timeline.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    local = int(os.environ.get('LOCAL_RANK', rank))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    import ddsp_piano_b200 as dp
    from ddsp_piano_b200 import sharding
    from ddsp_piano_b200.processors import _DEFAULT_CFG
    eng = dp.get_engine(dev, **{**_DEFAULT_CFG, 'sample_rate': 24000})
    S, N, L = 4, 7200, 7200
    rng = np.random.default_rng(99)                      # same timeline on every rank
    dry_all = (rng.standard_normal([world * S, N]) * 0.1).astype(np.float32)
    ir = (rng.standard_normal([L]) * np.exp(-6 * np.arange(L) / L) * 1e-2).astype(np.float32)
    lo, hi = sharding.clip_shard(world * S, rank, world)
    wet = sharding.timeline_reverb(torch.from_numpy(dry_all[lo:hi]).to(dev), torch.from_numpy(ir).to(dev),
                                   eng.reverb_full, rank, world)
    torch.cuda.synchronize()
    h = ir.astype(np.float64).copy()
    h[0] = 0
    x = dry_all.astype(np.float64).reshape(-1)
    want = (np.convolve(x, h)[:x.size] + x).reshape(world * S, N)[lo:hi]
    err = float(np.max(np.abs(wet.cpu().numpy() - want)) / np.max(np.abs(want)))
    # the same through the fused kernel: overlap-add + carry into the successor's buffer over peer memory
    peer = sharding.PeerTimeline(eng, hi - lo, N, L, rank, world)
    errs = []
    for _ in range(3):                                   # the buffers are reused call after call
        wet2 = peer.reverb(torch.from_numpy(dry_all[lo:hi]).to(dev), torch.from_numpy(ir).to(dev))
        torch.cuda.synchronize()
        errs.append(float(np.max(np.abs(wet2.cpu().numpy() - want)) / np.max(np.abs(want))))
    same = bool(torch.equal(wet2, peer.reverb(torch.from_numpy(dry_all[lo:hi]).to(dev),
                                              torch.from_numpy(ir).to(dev))))
    peer.close()
    print(f'rank {rank}: peer-memory timeline rel err {max(errs):.3e}, repeatable {same}', flush=True)
    err = max(err, max(errs)) if same else 1.0
    ok = torch.tensor([1.0 if err < 2e-5 else 0.0], device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    print(f'rank {rank}: timeline reverb rel err {err:.3e}', flush=True)
    dist.barrier()
    dist.destroy_process_group()
    if ok.item() != 1.0:
        sys.exit(1)
    if rank == 0:
        print('TIMELINE_OK')


if __name__ == '__main__':
    main()
