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

On the GPUs both exchanges of a timeline -- the oscillators' phase state and the reverb tail -- happen
INSIDE the kernels of ``b200ddsp_forward_timeline`` through NVLink peer memory; :class:`SpanChain` owns
the inboxes and counters.  :func:`timeline_reverb` and :func:`chain_carry` are the same two protocols
written with ``torch.distributed`` send/recv: the host-side model of the chain, exercised with the gloo
backend on CPU (``tests/test_sharding_gloo.py``); the convolution / the span computation is injected.
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
    bad = carry_out.shape != carry_like.shape or carry_out.dtype != carry_like.dtype
    if rank + 1 < world:
        # the successor is blocked in recv: it always gets a message (NaNs if this rank's carry is
        # unusable, so that the failure travels down the chain instead of hanging it)
        dist.send(torch.full_like(carry_like, float('nan')) if bad else carry_out.contiguous(),
                  dst=rank + 1, group=group)
    if bad:
        raise ValueError(f'carry changed from {tuple(carry_like.shape)} {carry_like.dtype} to '
                         f'{tuple(carry_out.shape)} {carry_out.dtype}')
    return result


class _DeviceBuffer:
    """A raw device allocation as a ``__cuda_array_interface__`` object (for torch.as_tensor)."""

    def __init__(self, ptr, n, typestr='<f4'):
        self.__cuda_array_interface__ = {'shape': (n,), 'typestr': typestr, 'data': (ptr, False),
                                         'version': 2}


def local_link(seed=None, carry=None, epoch=1):
    """A hand-off inside ONE process and stream (spans of a timeline synthesised one after the other on
    the same GPU): the payloads are ordinary device tensors, stream order is the synchronisation."""
    from . import _lib
    return _lib.Link(seed=seed.data_ptr() if seed is not None else None,
                     carry=carry.data_ptr() if carry is not None else None, epoch=epoch)


class SpanChain:
    """Rank r's place in a chain of spans of one timeline (BASELINE config 4): the inboxes and counters
    of the two stream-ordered hand-offs -- oscillator phase state and reverb tail -- that
    ``b200ddsp_forward_timeline`` performs INSIDE its kernels over NVLink peer memory (``csrc/link.cuh``,
    ``csrc/timeline.cuh``).  Nothing here runs per step except building two small structs: no barrier,
    no NCCL call on the data path (``torch.distributed`` is used once, to exchange the CUDA IPC handles).

    Mailbox layout (one peer-visible allocation per rank, zero-filled):
    8 x u64 counters [phase ready, phase ack, tail ready, tail ack, phase scratch x 2, tail scratch x 2],
    then 2 phase slots of ``n_phase`` floats, then 2 tail slots of ``n_tail`` floats."""

    COUNTER_BYTES = 256

    def __init__(self, engine, n_phase, n_tail, rank=0, world=1, group=None):
        if world > 1 and dist.get_backend(group) != 'nccl':
            raise ValueError('SpanChain hands state over through CUDA peer memory: it needs the NCCL '
                             'process group of one node (use timeline_reverb / chain_carry on CPU groups)')
        self.eng, self.rank, self.world, self.group = engine, rank, world, group
        self.n_phase, self.n_tail = int(n_phase), int(n_tail)
        al = lambda x: (x + 255) // 256 * 256
        self.phase_off = self.COUNTER_BYTES
        self.phase_slot = al(4 * self.n_phase)
        self.tail_off = self.phase_off + 2 * self.phase_slot
        self.tail_slot = al(4 * self.n_tail)
        nbytes = self.tail_off + 2 * self.tail_slot
        self.ptr, handle = engine.peer_alloc(nbytes)
        self.counters = torch.as_tensor(_DeviceBuffer(self.ptr, 8, '<i8'), device=engine.device)
        handles = [None] * world
        if world > 1:
            dist.all_gather_object(handles, handle, group=group)
        self.prev = engine.peer_open(handles[rank - 1]) if rank > 0 else 0
        self.next = engine.peer_open(handles[rank + 1]) if rank + 1 < world else 0
        self.epoch = 0
        if world > 1:
            dist.barrier(group=group)

    def links(self):
        """Advance the call counter and return (phase link, tail link) for this call."""
        from . import _lib
        self.epoch += 1
        e, slot = self.epoch, self.epoch & 1

        def link(ready, ack, scratch, off, slot_bytes):
            lk = _lib.Link(epoch=e, scratch=self.ptr + 8 * scratch)
            if self.prev:
                lk.seed = self.ptr + off + slot * slot_bytes
                lk.seed_ready = self.ptr + 8 * ready
                lk.seed_ack = self.prev + 8 * ack
            if self.next:
                lk.carry = self.next + off + slot * slot_bytes
                lk.carry_ready = self.next + 8 * ready
                lk.carry_ack = self.ptr + 8 * ack
            return lk

        return (link(0, 1, 4, self.phase_off, self.phase_slot),
                link(2, 3, 6, self.tail_off, self.tail_slot))

    def check(self):
        """Raise if a wait inside a kernel timed out (a peer never delivered).  Synchronises."""
        c = self.counters.cpu().tolist()
        if c[5] or c[7]:
            raise RuntimeError(f'rank {self.rank}: span hand-off timed out (phase {c[5]:#x}, tail {c[7]:#x}); '
                               f'counters {c[:4]} at epoch {self.epoch}')

    def close(self):
        torch.cuda.synchronize(self.eng.device)
        if self.world > 1:
            dist.barrier(group=self.group)
        for attr in ('prev', 'next'):
            if getattr(self, attr):
                self.eng.peer_close(getattr(self, attr))
                setattr(self, attr, 0)
        if self.world > 1:
            dist.barrier(group=self.group)
        if self.ptr:
            self.counters = None
            self.eng.peer_free(self.ptr)
            self.ptr = 0


def span_of(rank, world, frames_per_rank):
    """(in_first_frame, out_first_frame, n_input_frames) of rank ``rank``'s span when every rank
    synthesises ``frames_per_rank`` frames of a timeline of ``world * frames_per_rank``: one frame of
    halo on either side except at the two ends of the timeline."""
    total = world * frames_per_rank
    out0 = rank * frames_per_rank
    in0 = max(out0 - 1, 0)
    in1 = min(out0 + frames_per_rank + 1, total)
    return in0, out0, in1 - in0
