"""Run under torchrun (one rank per GPU): BASELINE config 4 -- one timeline cut into spans over the
ranks; oscillator phase state and reverb tail are handed over INSIDE the kernels through NVLink peer
memory (ddsp_piano_b200.sharding.SpanChain, csrc/link.cuh).  Every rank also synthesises the WHOLE
timeline on its own GPU and checks its span bit for bit (dry) and against the float64 reverb (wet).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/multi_gpu_timeline.py [--small]
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

KEYS = ('amplitudes', 'harmonic_distribution', 'inharm_coef', 'f0_hz', 'magnitudes')


def main():
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    local = int(os.environ.get('LOCAL_RANK', rank))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    import ddsp_piano_b200 as dp
    from ddsp_piano_b200 import _lib, sharding
    from ddsp_piano_b200.processors import _DEFAULT_CFG
    from scipy.signal import fftconvolve
    small = '--small' in sys.argv
    # full size: 16 segments of 3 s per rank, 16 voices, the 3 s impulse response (config 4 per rank)
    sr, P, B, H, S, M = 24000, (4 if small else 16), 1, 96, 2, 64
    seg, n_seg, L = (125, 4, 9000) if small else (750, 16, 72000)
    U = sr // 250
    F_rank = seg * n_seg
    total = world * F_rank
    eng = dp.get_engine(dev, **{**_DEFAULT_CFG, 'sample_rate': sr, 'n_noise_bands': M})

    rng = np.random.default_rng(4)                         # the same timeline on every rank
    n_notes = total // seg                                 # one pitch per voice and segment
    voices = []
    for v in range(P):
        midi = rng.integers(21, 109, size=n_notes)
        hz = 440.0 * 2.0 ** ((midi - 69) / 12.0)
        f0 = np.repeat(hz, seg)[None, :, None] * (1 + 1e-3 * np.arange(S))[None, None, :]
        voices.append({'f0_hz': f0.astype(np.float32),
                       'inharm_coef': rng.uniform(1e-4, 1e-3, [B, total, 1]).astype(np.float32),
                       'amplitudes': rng.standard_normal([B, total, 1]).astype(np.float32),
                       'harmonic_distribution': rng.standard_normal([B, total, H]).astype(np.float32),
                       'magnitudes': rng.standard_normal([B, total, M]).astype(np.float32)})
    ir = (rng.standard_normal([B, L]) * np.exp(-6 * np.arange(L) / L) * 1e-2).astype(np.float32)

    def sl(k0, k1):
        return [{k: torch.from_numpy(np.ascontiguousarray(v[k][:, k0:k1])).to(dev) for k in KEYS} for v in voices]

    in0, out0, F_in = sharding.span_of(rank, world, F_rank)
    mine = sl(in0, in0 + F_in)
    ir_d = torch.from_numpy(ir).to(dev)
    chain = sharding.SpanChain(eng, P * B * S * H, B * (L - 1), rank, world)
    n_calls = 5                                            # back to back: double buffering + acknowledgements
    for _ in range(n_calls):
        phase, tail = chain.links()
        span = _lib.Span(in_first_frame=in0, out_first_frame=out0, n_out_frames=F_rank, total_frames=total,
                         phase=phase)
        dry, wet = eng.forward_timeline(mine, ir_d, span, seg, tail=tail, seed=11)
    torch.cuda.synchronize()
    chain.check()

    # the whole timeline on this GPU alone
    whole, _ = eng.forward_polyphonic(sl(0, total), seed=11)
    torch.cuda.synchronize()
    lo, hi = out0 * U, (out0 + F_rank) * U
    dry_equal = bool(torch.equal(dry, whole[:, lo:hi]))
    x = whole.cpu().numpy().astype(np.float64)
    h = ir.astype(np.float64).copy()
    h[:, 0] = 0
    want = np.stack([fftconvolve(x[b], h[b])[:x.shape[1]] for b in range(B)]) + x
    err = float(np.max(np.abs(wet.cpu().numpy() - want[:, lo:hi])) / np.max(np.abs(want)))
    print(f'rank {rank}: span frames [{out0}, {out0 + F_rank}) of {total}: dry bit-equal to the whole-timeline '
          f'call: {dry_equal}; wet rel err vs float64 {err:.3e}', flush=True)
    ok = torch.tensor([1.0 if (dry_equal and err < 2e-5) else 0.0], device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    chain.close()
    dist.barrier()
    dist.destroy_process_group()
    if ok.item() != 1.0:
        sys.exit(1)
    if rank == 0:
        print('TIMELINE_OK')


if __name__ == '__main__':
    main()
