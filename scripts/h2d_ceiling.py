#!/usr/bin/env python
"""Host-to-device ceiling of one node with N ranks copying at once: how fast can
130.56 MB of pinned control tensors (one bench step of configs[2]) reach each GPU when 1, 2, 4, 8 ranks
copy concurrently, with no kernels running?  This is the bound of bench.py's `e2e` figure.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29533 scripts/h2d_ceiling.py > gpurun_out/h2d_ceiling_nN.json

Variants: ordinary page-locked memory (cudaHostAlloc default), write-combined page-locked memory, and the
copy split over two streams (two copy engines).  Times are CUDA events around 20 back-to-back copies,
all ranks released by a barrier; the slowest rank is what an end-to-end step sees.
"""
import ctypes
import json
import os
import sys

import torch
import torch.distributed as dist

BYTES = 130_560_000
REPS = 20


def main():
    rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', rank))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    rt = ctypes.CDLL('libcudart.so')
    rt.cudaHostAlloc.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_size_t, ctypes.c_uint]
    rt.cudaMemcpyAsync.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
    dst = torch.empty(BYTES, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def host(flags):
        p = ctypes.c_void_p()
        assert rt.cudaHostAlloc(ctypes.byref(p), BYTES, flags) == 0
        ctypes.memset(p, 1, BYTES)                     # first touch on this rank's CPUs
        return p

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def run(src, split, kind=1):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        half = BYTES // 2 // 256 * 256
        for timed in (False, True):
            barrier()
            e0.record(s1)
            s2.wait_stream(s1)
            for _ in range(REPS if timed else 3):
                a, b = (dst.data_ptr(), src.value) if kind == 1 else (src.value, dst.data_ptr())
                if split:
                    rt.cudaMemcpyAsync(a, b, half, kind, ctypes.c_void_p(s1.cuda_stream))
                    rt.cudaMemcpyAsync(a + half, b + half, BYTES - half, kind, ctypes.c_void_p(s2.cuda_stream))
                else:
                    rt.cudaMemcpyAsync(a, b, BYTES, kind, ctypes.c_void_p(s1.cuda_stream))
            s1.wait_stream(s2)
            e1.record(s1)
            barrier()
        return BYTES * REPS / (e0.elapsed_time(e1) * 1e-3) / 1e9

    plain, wc = host(0), host(4)                       # cudaHostAllocDefault, cudaHostAllocWriteCombined
    res = {'h2d_pinned': run(plain, False), 'h2d_pinned_two_streams': run(plain, True),
           'h2d_write_combined': run(wc, False), 'd2h_pinned': run(plain, False, kind=2)}
    if world > 1:
        alls = [None] * world
        dist.all_gather_object(alls, res)
    else:
        alls = [res]
    if rank == 0:
        out = {'n_ranks': world, 'bytes_per_copy': BYTES, 'reps': REPS, 'unit': 'GB/s per rank',
               'host_cpus': os.cpu_count()}
        for k in res:
            v = [a[k] for a in alls]
            out[k] = {'min': min(v), 'max': max(v), 'sum': sum(v), 'per_rank': [round(x, 2) for x in v]}
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
