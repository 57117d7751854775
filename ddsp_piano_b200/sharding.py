"""Multi-GPU partitioning of the synthesis path.  One process per GPU.

Two regimes:

* independent clips (BASELINE configs 2/3/5): clips are sharded over ranks, no data-path
  collective -- :func:`clip_shard`;
* one long timeline cut into fixed-length segments (BASELINE config 4): segments are
  synthesised independently (each is its own clip for the additive and noise processors,
  exactly as the reference's segment pipeline treats them), but the reverb is a convolution of
  the CONCATENATED dry timeline, so the last ``L-1`` wet samples of every segment spill into
  its successors.  Inside a rank that is a local overlap-add; across ranks it is one
  neighbour exchange (``send`` to rank+1 / ``recv`` from rank-1) of an ``L-1``-sample carry that
  is added to the head of the receiving rank's timeline -- :func:`timeline_reverb`.

A third piece, :func:`chain_carry`, is the rank-to-rank protocol for oscillator phase continuity
across segments (the carried state is specified in DESIGN.md section 6; the kernels do not take a
phase seed yet).

The functions are backend-agnostic (``torch.distributed`` with NCCL on GPUs, gloo in the CPU
tests); the convolution itself is injected (``conv_full``), on the GPU it is
``Engine.reverb_full`` (hand-written FFT convolution, 'valid'-padded).
"""
import torch
import torch.distributed as dist


def clip_shard(n_clips, rank, world):
    """Contiguous, balanced shard [start, stop) of ``n_clips`` independent clips."""
    if not (0 <= rank < world):
        raise ValueError(f'rank {rank} outside world of {world}')
    base, extra = divmod(n_clips, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def overlap_add_segments(wet_full, n_samples):
    """Local part of the timeline reverb.

    wet_full: [S, N + L - 1] 'valid'-padded convolution of each of the rank's S consecutive
    segments.  Returns (wet [S, N], carry [L - 1]) where ``wet`` already contains the spill of
    every local segment into its local successors and ``carry`` is what spills past the end of
    the rank's span (to be added at the start of the next rank's timeline)."""
    S, total = wet_full.shape
    N = n_samples
    tail = total - N                     # L - 1
    if tail > S * N:
        raise ValueError(f'reverb tail ({tail} samples) is longer than the rank\'s span '
                         f'({S} segments x {N}); use fewer ranks or longer spans')
    span = torch.zeros(S * N + tail, dtype=wet_full.dtype, device=wet_full.device)
    # deterministic order: segment by segment (tails may cover several successors)
    for i in range(S):
        span[i * N:i * N + total] += wet_full[i]
    return span[:S * N].reshape(S, N).clone(), span[S * N:].clone()


def exchange_carry(carry, rank, world, group=None):
    """Send this rank's carry to rank+1 and return the carry received from rank-1 (zeros on
    rank 0: the timeline starts there; the last rank's carry falls off the end)."""
    recv = torch.zeros_like(carry)
    if world == 1:
        return recv
    ops = []
    if rank + 1 < world:
        ops.append(dist.P2POp(dist.isend, carry.contiguous(), rank + 1, group))
    if rank > 0:
        ops.append(dist.P2POp(dist.irecv, recv, rank - 1, group))
    for req in dist.batch_isend_irecv(ops):
        req.wait()
    return recv


def timeline_reverb(dry, ir, conv_full, rank=0, world=1, add_dry=True, group=None):
    """Reverb of a timeline whose consecutive segments are spread over ranks.

    dry: [S, N] this rank's consecutive dry segments (rank r holds segments r*S .. r*S+S-1);
    ir: [L] (one impulse response for the whole timeline);
    conv_full(dry [S, N], ir [S, L]) -> [S, N + L - 1]: linear convolution with ir[0] masked
    (ddsp.effects.Reverb semantics), no dry signal added.
    Returns wet [S, N] == reverb(concatenated timeline)[this rank's span]."""
    S, N = dry.shape
    if ir.dim() != 1:
        raise ValueError('timeline_reverb takes a single impulse response [L]')
    L = ir.shape[0]
    wet_full = conv_full(dry, ir[None, :].expand(S, L).contiguous())
    if tuple(wet_full.shape) != (S, N + L - 1):
        raise ValueError(f'conv_full returned {tuple(wet_full.shape)}, expected {(S, N + L - 1)}')
    wet, carry = overlap_add_segments(wet_full, N)
    incoming = exchange_carry(carry, rank, world, group)
    flat = wet.reshape(-1)
    flat[:incoming.shape[0]] += incoming          # PeerTimeline fuses this into the overlap-add kernel
    wet = flat.reshape(S, N)
    return wet + dry if add_dry else wet


def chain_carry(finish, carry_like, rank, world, group=None):
    """Order-preserving chain over ranks for state that must be accumulated in timeline order (the
    oscillators' float32 chunk-offset sums across timeline segments, SURVEY 8e-i: float32 addition is
    not associative, so a tree scan would not reproduce the whole-clip phases bit for bit).

    ``finish(carry_in) -> (result, carry_out)`` completes this rank's span from the state at its start;
    everything that does not depend on the carry (the in-chunk phase sums, the bulk of the work) belongs
    BEFORE this call so that ranks only serialise on the seed.  Rank 0 starts from zeros; the carry is
    ``carry_like``-shaped (12 KB per clip at config 2).  Returns ``result``."""
    carry_in = torch.zeros_like(carry_like)
    if rank > 0:
        dist.recv(carry_in, src=rank - 1, group=group)
    result, carry_out = finish(carry_in)
    if rank + 1 < world:
        if carry_out.shape != carry_like.shape or carry_out.dtype != carry_like.dtype:
            raise ValueError(f'carry changed from {tuple(carry_like.shape)} {carry_like.dtype} to '
                             f'{tuple(carry_out.shape)} {carry_out.dtype}')
        dist.send(carry_out.contiguous(), dst=rank + 1, group=group)
    return result


class _DeviceBuffer:
    """A raw device allocation as a ``__cuda_array_interface__`` object (for torch.as_tensor)."""

    def __init__(self, ptr, n_floats):
        self.__cuda_array_interface__ = {'shape': (n_floats,), 'typestr': '<f4', 'data': (ptr, False),
                                         'version': 2}


class PeerTimeline:
    """The GPU form of :func:`timeline_reverb`: ONE kernel does the local overlap-add and adds the
    carry straight into the successor's output buffer over NVLink peer memory
    (``csrc/timeline.cuh``) -- no send/recv, no staging buffer, no separate add.

    Every rank owns a peer-visible output buffer [S * N] (``b200ddsp_peer_alloc``); the CUDA IPC
    handles are exchanged once, here, through ``torch.distributed`` and rank r maps the buffer of
    rank r + 1.  Per call: zero the head of the own buffer, barrier, launch, barrier."""

    def __init__(self, engine, n_segments, n_samples, ir_length, rank, world, group=None):
        self.eng, self.S, self.N, self.L = engine, n_segments, n_samples, ir_length
        self.rank, self.world, self.group = rank, world, group
        if ir_length - 1 > n_segments * n_samples:
            raise ValueError(f'reverb tail ({ir_length - 1} samples) is longer than the rank\'s span '
                             f'({n_segments} segments x {n_samples})')
        self.ptr, handle = engine.peer_alloc(n_segments * n_samples * 4)
        self.out = torch.as_tensor(_DeviceBuffer(self.ptr, n_segments * n_samples), device=engine.device)
        handles = [None] * world
        if world > 1:
            dist.all_gather_object(handles, handle, group=group)
        self.peer_head = engine.peer_open(handles[rank + 1]) if rank + 1 < world else 0

    def reverb(self, dry, ir, add_dry=True):
        """dry [S, N] (this rank's consecutive segments), ir [L] -> wet [S, N], a view of the
        peer-visible buffer (valid until the next call)."""
        S, N, L = self.S, self.N, self.L
        if tuple(dry.shape) != (S, N) or tuple(ir.shape) != (L,):
            raise ValueError(f'expected dry {(S, N)} and ir {(L,)}, got {tuple(dry.shape)}, {tuple(ir.shape)}')
        wet_full = self.eng.reverb_full(dry, ir[None, :].expand(S, L).contiguous())
        self.out[:L - 1].zero_()
        if self.world > 1:
            dist.barrier(group=self.group)           # every head is zero before anyone adds into it
        self.eng.timeline_overlap_add(wet_full, dry if add_dry else None, self.ptr, self.peer_head, S, N, L)
        if self.world > 1:
            dist.barrier(group=self.group)           # every carry has landed
        return self.out.view(S, N)

    def close(self):
        if self.world > 1:
            dist.barrier(group=self.group)
        if self.peer_head:
            self.eng.peer_close(self.peer_head)
            self.peer_head = 0
        if self.world > 1:
            dist.barrier(group=self.group)
        if self.ptr:
            self.out = None
            self.eng.peer_free(self.ptr)
            self.ptr = 0
