"""N > 1 host logic on CPU: world_size-2 (and 3) gloo groups run the timeline reverb exchange
(ddsp_piano_b200/sharding.py) with a numpy convolution standing in for the CUDA kernel, and the
result is compared with the reverb of the concatenated timeline computed in one piece."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ddsp_piano_b200 import sharding


def test_clip_shard_partitions_exactly():
    for n in (0, 1, 7, 16, 128, 129):
        for world in (1, 2, 3, 8):
            spans = [sharding.clip_shard(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.clip_shard(4, 2, 2)


def conv_full_numpy(dry, ir):
    """'valid'-padded convolution per segment with ir[0] masked (ddsp.effects.Reverb)."""
    out = []
    for x, h in zip(dry.numpy().astype(np.float64), ir.numpy().astype(np.float64)):
        h = h.copy()
        h[0] = 0.0
        out.append(np.convolve(x, h))
    return torch.from_numpy(np.stack(out).astype(np.float32))


def reference_timeline(dry_all, ir, add_dry=True):
    h = ir.numpy().astype(np.float64).copy()
    h[0] = 0.0
    x = dry_all.numpy().astype(np.float64).reshape(-1)
    wet = np.convolve(x, h)[:x.size]
    return torch.from_numpy((wet + (x if add_dry else 0.0)).astype(np.float32)).reshape(dry_all.shape)


def test_overlap_add_single_rank_long_tail():
    g = torch.Generator().manual_seed(0)
    S, N, L = 5, 40, 95            # the tail spans more than two segments
    dry = torch.randn(S, N, generator=g)
    ir = torch.randn(L, generator=g) * 0.1
    wet = sharding.timeline_reverb(dry, ir, conv_full_numpy)
    assert torch.allclose(wet, reference_timeline(dry, ir), atol=2e-6)
    with pytest.raises(ValueError):
        sharding.timeline_reverb(dry[:2], torch.randn(3 * N), conv_full_numpy)   # tail > span


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, S, N, L, add_dry, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(1234)
        dry_all = torch.randn(world * S, N, generator=g)        # same timeline on every rank
        ir = torch.randn(L, generator=g) * 0.1
        lo, hi = sharding.clip_shard(world * S, rank, world)
        wet = sharding.timeline_reverb(dry_all[lo:hi].clone(), ir, conv_full_numpy, rank, world,
                                       add_dry=add_dry)
        want = reference_timeline(dry_all, ir, add_dry)[lo:hi]
        q.put((rank, float((wet - want).abs().max()), float(want.abs().max())))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world,S,N,L,add_dry', [(2, 4, 96, 96, True),     # config-4 shape: L == N
                                                 (2, 3, 64, 150, True),    # tail spans 3 segments
                                                 (3, 2, 50, 20, False)])
def test_timeline_reverb_across_ranks(world, S, N, L, add_dry):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, S, N, L, add_dry, q))
             for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in results) == list(range(world))
    for _, err, scale in results:
        assert err <= 2e-6 * max(scale, 1.0)


def _phase_worker(rank, world, port, q):
    """Each rank synthesises its third of a clip with the oracle's segment form; the chunk-offset carry
    travels down the ranks through sharding.chain_carry."""
    from oracle import ddsp_piano_np as ref
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(11)                         # same clip on every rank
        sr, B, H, S, seg = 24000, 1, 16, 2, 125
        F = world * seg
        f0 = 220.0 * 2 ** rng.uniform(0, 2, [B, 1, 1]) * (1 + 1e-3 * np.arange(S))[None, None, :]
        f0 = np.broadcast_to(f0, [B, F, S]).astype(np.float32).copy()
        ctl = ref.additive_controls(rng.standard_normal([B, F, 1]).astype(np.float32),
                                    rng.standard_normal([B, F, H]).astype(np.float32),
                                    rng.uniform(1e-4, 1e-3, [B, F, 1]).astype(np.float32), f0, sample_rate=sr)

        def finish(carry_in):
            y, carry = ref.additive_signal_segment(**ctl, frames=(rank * seg, (rank + 1) * seg),
                                                   carry=carry_in.numpy(), sample_rate=sr)
            return y, torch.from_numpy(carry)

        got = sharding.chain_carry(finish, torch.zeros(S, B, H), rank, world)
        whole = ref.additive_signal(**ctl, sample_rate=sr, inference=True)
        n = seg * 96
        q.put((rank, bool(np.array_equal(got, whole[:, rank * n:(rank + 1) * n]))))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 3])
def test_phase_carry_chain_across_ranks_is_bit_exact(world):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_phase_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(results) == [(r, True) for r in range(world)]
