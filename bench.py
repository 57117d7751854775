#!/usr/bin/env python
"""Benchmark of the DDSP-Piano synthesis hot path (BASELINE.json metric: real-time factor,
audio seconds per wall second, 24 kHz, batch 16, poly 16).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload full|dry|stress|timeline]
    python bench.py --impl reference ...        # the CPU oracle on the host cores

A step = one forward of the whole polyphonic ProcessorGroup (16 voices x (additive + noise), running
sum, reverb) over one batch of synthetic control tensors.  `value` is timed on the device with inputs
resident in HBM; `e2e` goes through the reference-facing ProcessorGroup call with pinned HOST inputs
(H2D of every control tensor and D2H of the audio inside the timed region).

Workloads (BASELINE.json configs):
  full      configs[2]  16 independent 3 s clips, full chain incl. 3 s reverb   (default at N = 1)
  dry       configs[1]  the same without the reverb
  stress    configs[4]  48 kHz / poly 32 / 128 partials
  timeline  configs[3]  ONE timeline of N x 16 contiguous 3 s segments, 16 per GPU: oscillator phases,
                        resamplers, noise FIR and reverb run THROUGH the segment and rank boundaries
                        (the reference synthesises a piece in one pass); phase state and reverb tail
                        cross ranks inside the kernels over NVLink peer memory   (default at N > 1)
At N = 1 the line carries `dry`, `stress` and `timeline` (one rank alone) as extra keys.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    'full': dict(name='configs[2]: batch16 x 3s, poly16, H96, M64, additive+noise+3s reverb, 24kHz',
                 sr=24000, B=16, P=16, S=2, H=96, M=64, F=750, L=72000),
    'dry': dict(name='configs[1]: batch16 x 3s, poly16, H96, M64, additive+noise (no reverb), 24kHz',
                sr=24000, B=16, P=16, S=2, H=96, M=64, F=750, L=0),
    'stress': dict(name='configs[4]: batch16 x 3s, poly32, H128, M96, full chain, 48kHz',
                   sr=48000, B=16, P=32, S=2, H=128, M=96, F=750, L=144000),
    # per rank: B = 16 consecutive 3 s segments of ONE timeline (a single row of 12 000 frames)
    'timeline': dict(name='configs[3]: one timeline of n_gpus x 16 contiguous 3 s segments (16 per GPU), poly16, '
                          'H96, M64, additive+noise+3s reverb through every boundary, 24kHz',
                     sr=24000, B=16, P=16, S=2, H=96, M=64, F=750, L=72000),
}
METRIC = 'real-time factor (audio-sec/wall-sec) @24kHz batch16 poly16; HBM GB/s %peak'   # BASELINE.json
UNIT = 'x real time'
CONTROL_KEYS = ('amplitudes', 'harmonic_distribution', 'inharm_coef', 'f0_hz', 'magnitudes')
# FMA-pipe lane-cycles per live oscillator-sample, counted from the SASS of the two kernels
# (2 scalar FMUL + 5 packed ops per pair of chains in the phase chain = 6; + cross-fade,
# turn count every 4th sample, wrap, cosine accumulate = 11.5 in the synthesis pass)
CYCLES_SYNTH, CYCLES_PHASE = 11.5, 6.0


# ------------------------------------------------------------------------------------------
# synthetic inputs
# ------------------------------------------------------------------------------------------

def synthetic_inputs(w, seed, B=None, held_notes=False):
    """SURVEY.md 8d config 2/3 distributions, pre-get_controls, stacked [P, B, F, C] float32.
    f0 is constant over frames as BASELINE.md section 3 states.  inharm_coef is drawn per FRAME
    by default (the distribution is stated per element), which makes every partial frequency
    move every frame -- the most expensive case for the kernels; held_notes=True draws it per
    (voice, clip) like the reference model does (InharmonicityNetwork is a function of pitch)."""
    B = w['B'] if B is None else B
    P, F, H, S, M, L = w['P'], w['F'], w['H'], w['S'], w['M'], w['L']
    rng = np.random.default_rng(seed)
    midi = rng.integers(21, 109, size=[P, B, 1, 1])
    f0 = 440.0 * 2.0 ** ((midi - 69) / 12.0) * (1.0 + 1e-3 * np.arange(S))[None, None, None, :]
    x = {
        'f0_hz': np.broadcast_to(f0, [P, B, F, S]).astype(np.float32).copy(),
        'inharm_coef': (np.broadcast_to(rng.uniform(1e-4, 1e-3, [P, B, 1, 1]), [P, B, F, 1])
                        if held_notes else rng.uniform(1e-4, 1e-3, [P, B, F, 1])
                        ).astype(np.float32).copy(),
        'amplitudes': rng.standard_normal([P, B, F, 1], dtype=np.float32),
        'harmonic_distribution': rng.standard_normal([P, B, F, H], dtype=np.float32),
        'magnitudes': rng.standard_normal([P, B, F, M], dtype=np.float32),
    }
    if L:
        x['reverb_ir'] = impulse_response(rng, B, L)
    return x


def impulse_response(rng, B, L):
    t = np.arange(L) / L
    return (rng.standard_normal([B, L]) * np.exp(-6 * t) * 1e-2).astype(np.float32)


def timeline_segment(w, g, v):
    """Controls of global segment g (3 s) of voice v of the benchmark timeline: same distributions as
    synthetic_inputs, one pitch per (voice, segment), a function of (g, v) only so that every rank can
    produce its neighbours' halo frames.  Dict of [F, C] float32."""
    F, H, S, M = w['F'], w['H'], w['S'], w['M']
    rng = np.random.default_rng([20251017, g, v])
    hz = 440.0 * 2.0 ** ((int(rng.integers(21, 109)) - 69) / 12.0)
    return {'f0_hz': np.broadcast_to(hz * (1.0 + 1e-3 * np.arange(S))[None, :], [F, S]).astype(np.float32),
            'inharm_coef': rng.uniform(1e-4, 1e-3, [F, 1]).astype(np.float32),
            'amplitudes': rng.standard_normal([F, 1], dtype=np.float32),
            'harmonic_distribution': rng.standard_normal([F, H], dtype=np.float32),
            'magnitudes': rng.standard_normal([F, M], dtype=np.float32)}


def timeline_inputs(w, rank, world):
    """This rank's span of the timeline: stacked [P, 1, F_in, C] over its input frames (16 segments plus
    one halo frame either side, except at the two ends of the timeline) + the timeline's impulse response."""
    from ddsp_piano_b200 import sharding
    n_seg, F, P = w['B'], w['F'], w['P']
    in0, out0, F_in = sharding.span_of(rank, world, n_seg * F)
    g0, g1 = in0 // F, (in0 + F_in - 1) // F
    x = {k: [] for k in CONTROL_KEYS}
    for v in range(P):
        segs = [timeline_segment(w, g, v) for g in range(g0, g1 + 1)]
        for k in CONTROL_KEYS:
            row = np.concatenate([s[k] for s in segs], axis=0)
            x[k].append(row[in0 - g0 * F:in0 - g0 * F + F_in][None])
    x = {k: np.ascontiguousarray(np.stack(v)) for k, v in x.items()}
    x['reverb_ir'] = impulse_response(np.random.default_rng(20251017), 1, w['L'])
    return x, (in0, out0, F_in)


def algorithmic_bytes(w, B):
    """raw control tensors read once + impulse responses + outputs written once."""
    P, F, H, S, M, L = w['P'], w['F'], w['H'], w['S'], w['M'], w['L']
    N = F * (w['sr'] // 250)
    return P * B * F * (1 + H + 1 + S + M) * 4 + B * L * 4 + B * N * 4 * (2 if L else 1)


def live_chain_samples(w, f0, inharm, carry_all=False):
    """Oscillator-samples the kernels actually run for these inputs ([P, B, F] f0 of string 0 and
    inharmonicity): they drop 16-partial half-groups that lie above Nyquist (or belong to a muted voice)
    in every frame a 1000-sample chunk touches; a unit with nh live half-groups runs nh
    chains on each of the warp's 32 lanes.  Returns (synthesis pass, phase pass); the phase pass of a
    timeline span follows every partial through every chunk (carry_all)."""
    F, H, S, sr = f0.shape[-1], w['H'], w['S'], w['sr']
    U = sr // 250
    N = F * U
    f0 = f0.astype(np.float64)
    binh = np.maximum(inharm.astype(np.float64), 0.0)
    n = np.arange(1, H + 1, dtype=np.float64)
    freq = f0[..., None] * n * np.sqrt(1.0 + binh[..., None] * n * n)    # [P, B, F, H]
    can = (freq < sr / 2.0) & (f0[..., None] > 20.0)
    top = np.where(can.any(-1), H - np.argmax(can[..., ::-1], axis=-1), 0)   # 1 + highest live partial
    nh_frame = -(-top // 16)
    lanes = 16 if S % 2 == 0 else 32
    nh_chunk = []
    for t0 in range(0, N, 1000):
        t1 = min(N, t0 + 1000) - 1
        k0, k1 = t0 // U, min(F - 1, t1 // U + 1)
        nh_chunk.append((nh_frame[..., k0:k1 + 1].max(-1), t1 + 1 - t0))
    per = lambda nh: nh if S % 2 == 0 else -(-nh // 2)
    synth = sum(int(per(nh).sum()) * lanes * S * n_s for nh, n_s in nh_chunk)
    if carry_all:
        phase = int(np.prod(f0.shape[:-1])) * int(per(np.int64(-(-H // 16)))) * lanes * S * N
    else:
        # a chunk's end phase matters only to later chunks: suffix maximum of the live half-groups
        later = np.zeros_like(nh_chunk[0][0])
        phase = 0
        for i in range(len(nh_chunk) - 2, -1, -1):
            later = np.maximum(later, nh_chunk[i + 1][0])
            phase += int(per(later).sum()) * lanes * S * nh_chunk[i][1]
    return synth, int(phase)


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(self.index), f'--query-gpu={self.Q}',
                 '--format=csv,noheader,nounits', '-lms', '100'],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith('active'):
                        reasons.add(n)
            except Exception:
                pass
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': mx,
                'reasons': sorted(reasons), 'samples': len(sm)}


# ------------------------------------------------------------------------------------------
# CPU side: the oracle as the reported baseline / reference arm.  Everything is MEASURED: a step runs
# the whole workload of one GPU (P x B (voice, clip) units + the reverb) on all host cores.
# ------------------------------------------------------------------------------------------

def _oracle_voice_clip(args):
    """One (voice, clip) of the clip workloads through the numpy oracle: get_controls + get_signal of
    the additive and noise processors."""
    (sr, F, H, S, M, seed) = args
    from oracle import ddsp_piano_np as ref
    rng = np.random.default_rng(seed)
    hz = 440.0 * 2.0 ** ((int(rng.integers(21, 109)) - 69) / 12.0)
    f0 = np.broadcast_to((hz * (1.0 + 1e-3 * np.arange(S)))[None, None, :], [1, F, S]).astype(np.float32)
    amp = rng.standard_normal([1, F, 1], dtype=np.float32)
    hd = rng.standard_normal([1, F, H], dtype=np.float32)
    inh = rng.uniform(1e-4, 1e-3, [1, F, 1]).astype(np.float32)
    mags = rng.standard_normal([1, F, M], dtype=np.float32)
    noise = rng.uniform(-1, 1, [1, F * (sr // 250)]).astype(np.float32)
    c = ref.additive_controls(amp, hd, inh, f0, sample_rate=sr)
    a = ref.additive_signal(**c, sample_rate=sr, inference=True)
    n = ref.noise_signal(ref.noise_controls(mags)['magnitudes'], noise)
    return float(np.abs(a + n).max())


def _oracle_timeline_voice(args):
    """One voice of a timeline span through the oracle's segment forms (phase state carried from
    segment to segment, one frame of halo: oracle/ddsp_piano_np.py::additive_signal_segment,
    noise_signal_segment) -- the CPU restatement of what one GPU does per step of the timeline workload."""
    (w, v, n_seg) = args
    from oracle import ddsp_piano_np as ref
    sr, F = w['sr'], w['F']
    U = sr // 250
    segs = [timeline_segment(w, g, v) for g in range(n_seg)]
    ctl_in = {k: np.concatenate([s[k] for s in segs], axis=0)[None] for k in CONTROL_KEYS}
    ctl = ref.additive_controls(ctl_in['amplitudes'], ctl_in['harmonic_distribution'], ctl_in['inharm_coef'],
                                ctl_in['f0_hz'], sample_rate=sr)
    mags = ref.noise_controls(ctl_in['magnitudes'])['magnitudes']
    noise = np.random.default_rng(v).uniform(-1, 1, [1, n_seg * F * U]).astype(np.float32)
    carry, peak = None, 0.0
    for g in range(n_seg):
        a, carry = ref.additive_signal_segment(**ctl, frames=(g * F, (g + 1) * F), carry=carry, sample_rate=sr)
        n = ref.noise_signal_segment(mags, noise, frames=(g * F, (g + 1) * F), sample_rate=sr)
        peak = max(peak, float(np.abs(a + n).max()))
    return peak


def _oracle_reverb(args):
    (N, L, seed) = args
    from oracle import ddsp_piano_np as ref
    rng = np.random.default_rng(seed)
    audio = rng.standard_normal([1, N], dtype=np.float32)
    ir = rng.standard_normal([1, L], dtype=np.float32)
    return float(np.abs(ref.reverb_signal(audio, ir)).max())


def cpu_step(w, workload, pool):
    """One step of one GPU's workload on the host cores through the oracle.  Returns measured seconds."""
    sr, F, H, S, M, L, B, P = (w[k] for k in ('sr', 'F', 'H', 'S', 'M', 'L', 'B', 'P'))
    N = F * (sr // 250)
    t0 = time.perf_counter()
    if workload == 'timeline':
        pool.map(_oracle_timeline_voice, [(w, v, B) for v in range(P)], chunksize=1)
        pool.map(_oracle_reverb, [(B * N, L, 5)], chunksize=1)          # one pass over the whole span
    else:
        pool.map(_oracle_voice_clip, [(sr, F, H, S, M, 1000 + i) for i in range(P * B)], chunksize=1)
        if L:
            pool.map(_oracle_reverb, [(N, L, 5 + b) for b in range(B)], chunksize=1)
    return time.perf_counter() - t0


def cpu_sample_desc(w, workload, cores, secs):
    P, B = w['P'], w['B']
    if workload == 'timeline':
        return (f'one GPU\'s span of the timeline, complete: {P} voices x {B} consecutive 3 s segments through the '
                f'oracle\'s segment forms (carried phase state, halo) + one reverb pass over the span, {cores} worker '
                f'processes, measured {secs:.2f} s per step (not extrapolated)')
    return (f'the whole batch of one GPU: {P * B} (voice, clip) units (get_controls + get_signal, additive + noise)'
            + (f' + {B} reverbs' if w['L'] else '') + f' on {cores} worker processes, measured {secs:.2f} s per step '
            '(not extrapolated)')


def run_reference(args, w, workload):
    """--impl reference: the reference's TF2 CPU path is not installable (tensorflow / ddsp absent here
    and on the GPU box, no index: profiles/r02_tf_probe.txt); its CPU restatement (oracle/, numpy) is
    timed instead, one worker process per host core.  Every step runs one GPU's complete workload and
    ms_per_step is its measured wall time; the number of steps is capped by a time budget."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    budget_s = float(os.environ.get('B200DDSP_REFERENCE_BUDGET_S', '150'))
    t_begin = time.perf_counter()
    secs = []
    with mp.get_context('spawn').Pool(cores) as pool:
        pool.map(_oracle_voice_clip, [(w['sr'], 25, w['H'], w['S'], w['M'], i) for i in range(cores)])   # imports
        warm = cpu_step(w, workload, pool) if args.warmup > 0 else None   # one full warm-up step
        for _ in range(args.steps):
            secs.append(cpu_step(w, workload, pool))
            if time.perf_counter() - t_begin + 1.3 * secs[-1] > budget_s:
                break
    sec = float(np.mean(secs))
    audio_sec = w['B'] * w['F'] / 250.0
    value = audio_sec / sec
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': len(secs), 'steps_requested': args.steps, 'warmup': 1 if warm is not None else 0,
        'ms_per_step': 1e3 * sec, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic', 'config': {'workload': w['name']},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                         'sample': cpu_sample_desc(w, workload, cores, sec), 'extrapolated': False,
                         'step_seconds': secs, 'warmup_step_seconds': warm},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
        'note': 'host CPUs do not multiply with --gpus: the value is one host\'s throughput on one GPU\'s share of '
                'the workload, whatever N is',
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------
# GPU side
# ------------------------------------------------------------------------------------------

def bind_to_gpu_numa_node(index):
    """Pin this rank to the CPUs of the NUMA node its GPU hangs off, BEFORE the pinned host buffers
    are allocated (first touch places them on that node): with several ranks per host the
    host-to-device copies otherwise cross the socket interconnect.  Returns the node or None."""
    try:
        bus = subprocess.run(['nvidia-smi', '--query-gpu=pci.bus_id', '--format=csv,noheader', '-i',
                              str(index)], capture_output=True, text=True, timeout=20).stdout.strip()
        bus = bus.lower()
        if bus.count(':') == 2 and len(bus.split(':')[0]) == 8:      # 00000000:1b:00.0 -> 0000:1b:00.0
            bus = bus[4:]
        with open(f'/sys/bus/pci/devices/{bus}/numa_node') as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(f'/sys/devices/system/node/node{node}/cpulist') as f:
            cpus = set()
            for part in f.read().strip().split(','):
                lo, _, hi = part.partition('-')
                cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


class GpuBench:
    """Shared state of the GPU arm: device, process group, timing helpers."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        import __graft_entry__
        self.torch, self.dist, self.args = torch, dist, args
        self.rank = int(os.environ.get('RANK', '0'))
        self.local = int(os.environ.get('LOCAL_RANK', '0'))
        self.world = int(os.environ.get('WORLD_SIZE', '1'))
        if self.rank == 0:
            __graft_entry__.build()
        torch.cuda.set_device(self.local)
        self.dev = torch.device('cuda', self.local)
        if self.world > 1:
            dist.init_process_group('nccl', device_id=self.dev)
            dist.barrier()
        import ddsp_piano_b200 as dp
        self.dp = dp
        self.numa_node = bind_to_gpu_numa_node(self.local) if self.world > 1 else None
        self.flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=self.dev)   # > 126 MB L2
        # the write leaves the L2 full of DIRTY lines whose write-back would be charged to the first
        # kernels of the timed step; reading a second buffer afterwards leaves it cold and clean
        self.flush_read = torch.zeros_like(self.flush) if os.environ.get('B200DDSP_FLUSH_READ', '1') != '0' else None

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def timed(self, fn, steps, eng=None, stages=None):
        """K steps, each bracketed by CUDA events on the launching stream, L2 flushed between steps
        (outside the brackets), a barrier + synchronize on both sides.  Returns total ms (this rank)."""
        torch = self.torch
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
               for _ in range(steps)]
        self.barrier()
        for i in range(steps):
            self.flush.zero_()
            if self.flush_read is not None:
                self.flush_read.sum()
            evs[i][0].record()
            fn()
            evs[i][1].record()
            if stages is not None:
                for k, v in eng.last_stage_ms().items():
                    stages[k] = stages.get(k, 0.0) + v
        self.barrier()
        return sum(a.elapsed_time(b) for a, b in evs)

    def reduce_max(self, v):
        if self.world == 1:
            return v
        t = self.torch.tensor([v], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def group(self, w, fast_phase=False):
        dp = self.dp
        sr, P, L = w['sr'], w['P'], w['L']
        additive = dp.MultiInharmonic(frame_rate=250, sample_rate=sr, inference=True, name='additive',
                                      fast_phase=fast_phase)
        noise = dp.DynamicSizeFilteredNoise(frame_rate=250, sample_rate=sr, name='noise', seed=1234)
        reverb = dp.Reverb(trainable=False) if L else None
        group = dp.ProcessorGroup(dag=dp.polyphonic_dag(
            additive=additive, noise=noise, reverb=reverb,
            additive_controls=['amplitudes', 'harmonic_distribution', 'inharm_coef', 'f0_hz'],
            noise_controls=['magnitudes'], reverb_controls=['reverb_ir'] if L else [], n_synths=P))
        from ddsp_piano_b200.processors import _DEFAULT_CFG
        cfg = {**_DEFAULT_CFG, **additive.engine_config(), **noise.engine_config(w['M'])}
        if reverb is not None:
            cfg.update(reverb.engine_config())
        return group, dp.get_engine(self.dev, **cfg)

    @staticmethod
    def features(parents, P, L):
        f = {f'{k}_{v}': parents[k][v] for k in CONTROL_KEYS for v in range(P)}
        if L:
            f['reverb_ir'] = parents['reverb_ir']
        return f


def measure_clips(gb, w, steps, warmup, held=False, e2e=True, fast_phase=False):
    """Independent clips sharded over ranks (no collective on the data path)."""
    torch = gb.torch
    P, L, B, F = w['P'], w['L'], w['B'], w['F']
    N = F * (w['sr'] // 250)
    x_np = synthetic_inputs(w, seed=gb.rank, held_notes=held)
    host = {k: torch.from_numpy(v).pin_memory() for k, v in x_np.items()}
    resident = {k: v.to(gb.dev) for k, v in host.items()}
    group, eng = gb.group(w, fast_phase)
    eng.set_profiling(True)
    # the per-voice views exist once (Parallelizer.unparallelize hands them out the same way);
    # every step passes a fresh shallow copy of the dict because ProcessorGroup extends it
    fr, fh = gb.features(resident, P, L), gb.features(host, P, L)
    for _ in range(warmup):
        group(dict(fr), return_outputs_dict=False)
    l0 = gb.dp.total_launches()
    stages = {}
    ms = gb.timed(lambda: group(dict(fr), return_outputs_dict=False), steps, eng, stages)
    launches = gb.dp.total_launches() - l0
    out = {'ms': gb.reduce_max(ms) / steps, 'stages': {k: v / steps for k, v in stages.items()},
           'launches': launches, 'inputs': x_np, 'eng': eng, 'audio_sec': gb.world * B * F / 250.0}
    if e2e:
        # pinned HOST control tensors in, pinned HOST audio out: the ProcessorGroup routes CPU features
        # to b200ddsp_forward_polyphonic_host (H2D + kernels + D2H of dry and wet on the timed stream)
        for _ in range(3):
            group(dict(fh), return_outputs_dict=False)
        out['e2e_ms'] = gb.reduce_max(gb.timed(lambda: group(dict(fh), return_outputs_dict=False), steps)) / steps
        out['h2d'] = sum(v.numel() * 4 for v in host.values())
        out['d2h'] = B * N * 4 * (2 if L else 1)      # the reference returns both outs['add'] and the wet signal
    eng.set_profiling(False)
    return out


def measure_timeline(gb, w, steps, warmup):
    """configs[3]: this rank's 16 segments of the timeline; phase state and reverb tail cross the ranks
    inside the kernels (sharding.SpanChain: NVLink peer memory, stream-ordered, no host barrier)."""
    torch = gb.torch
    from ddsp_piano_b200 import _lib, sharding
    P, L, n_seg, F, S, H = w['P'], w['L'], w['B'], w['F'], w['S'], w['H']
    U = w['sr'] // 250
    x_np, (in0, out0, F_in) = timeline_inputs(w, gb.rank, gb.world)
    host = {k: torch.from_numpy(v).pin_memory() for k, v in x_np.items()}
    resident = {k: v.to(gb.dev) for k, v in host.items()}
    group, eng = gb.group(w)
    eng.set_profiling(True)
    chain = sharding.SpanChain(eng, P * S * H, L - 1, gb.rank, gb.world)
    fr, fh = gb.features(resident, P, L), gb.features(host, P, L)
    total = gb.world * n_seg * F

    def step(feats):
        phase, tail = chain.links()
        span = _lib.Span(in_first_frame=in0, out_first_frame=out0, n_out_frames=n_seg * F, total_frames=total,
                         phase=phase)
        return group(dict(feats), return_outputs_dict=False,
                     timeline={'span': span, 'seg_frames': F, 'tail': tail})

    for _ in range(warmup):
        step(fr)
    l0 = gb.dp.total_launches()
    stages = {}
    ms = gb.timed(lambda: step(fr), steps, eng, stages)
    launches = gb.dp.total_launches() - l0
    for _ in range(3):
        step(fh)
    e2e_ms = gb.timed(lambda: step(fh), steps)
    chain.check()
    chain.close()
    eng.set_profiling(False)
    N = n_seg * F * U
    return {'ms': gb.reduce_max(ms) / steps, 'e2e_ms': gb.reduce_max(e2e_ms) / steps,
            'stages': {k: v / steps for k, v in stages.items()}, 'launches': launches, 'inputs': x_np, 'eng': eng,
            'audio_sec': gb.world * n_seg * F / 250.0, 'h2d': sum(v.numel() * 4 for v in host.values()),
            'd2h': 2 * N * 4,
            'exchange_bytes_per_step': (P * S * H + (L - 1)) * 4 * (gb.world - 1)}


def ncu_traffic(workload):
    """dram bytes per launch set of the dominant kernel from a committed ncu capture of THIS build
    (profiles/r02_ncu_traffic.json, written from the .ncu-rep by profiles/ncu_summary.py), or None."""
    try:
        with open(os.path.join(ROOT, 'profiles', 'r02_ncu_traffic.json')) as f:
            t = json.load(f)
        return t.get(workload)
    except Exception:
        return None


def roofline_of(gb, w, m, fma_peak, clocks, workload):
    """Dominant kernel = the oscillator bank (additive_synth_kernel<NH,2>, one launch per bucket,
    concurrent).  HBM view per the contract (SURVEY 8d algorithmic bytes over the kernel's measured
    duration) and, beside it, the roof that binds: the FP32 FMA pipe against a peak MEASURED at the
    start of this run (b200ddsp_measure_fma_rate)."""
    peaks, peaks_src = {}, 'fallback'
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            peaks, peaks_src = json.load(f), 'measured'
    except Exception:
        pass
    hbm_peak = float(peaks.get('hbm_gbs', 6650.0))
    timeline = workload == 'timeline'
    B = 1 if timeline else w['B']
    wb = dict(w, F=w['F'] * (w['B'] if timeline else 1))
    ab = algorithmic_bytes(wb, B)
    osc_ms, ends_ms = m['stages'].get('oscillators', 0.0), m['stages'].get('phase_ends', 0.0)
    x = m['inputs']
    if timeline:    # the output frames of the span only
        koff = x['f0_hz'].shape[2] - wb['F']
        lo = 1 if (koff >= 1 and gb.rank > 0) else 0
        f0, inh = x['f0_hz'][:, :, lo:lo + wb['F'], 0], x['inharm_coef'][:, :, lo:lo + wb['F'], 0]
    else:
        f0, inh = x['f0_hz'][..., 0], x['inharm_coef'][..., 0]
    live_synth, live_phase = live_chain_samples(w, f0, inh, carry_all=timeline and gb.world > 1)
    achieved = ab / (osc_ms * 1e-3) / 1e9 if osc_ms > 0 else None
    traffic = ncu_traffic(workload) if gb.world == 1 else None
    fp32 = None
    if osc_ms > 0 and fma_peak:
        need_s, need_p = live_synth * CYCLES_SYNTH, live_phase * CYCLES_PHASE
        fp32 = {'bound': 'fp32', 'unit': 'G lane-FMA/s',
                'peak': fma_peak['packed'] / 1e9,
                'peak_source': 'measured at the start of this run: b200ddsp_measure_fma_rate (independent FFMA2 '
                               f'chains on every SM, best of 3); scalar FFMA {fma_peak["scalar"] / 1e9:.0f}; '
                               'nominal 148 SMs x 128 lanes x SM clock = '
                               f'{148 * 128 * (clocks.get("sm_max_mhz") or 1965.0) / 1e3:.0f}',
                'oscillators': {'achieved': need_s / (osc_ms * 1e-3) / 1e9,
                                'frac': need_s / (osc_ms * 1e-3) / fma_peak['packed'],
                                'lane_cycles_per_live_oscillator_sample': CYCLES_SYNTH,
                                'live_oscillator_samples': live_synth, 'kernel_ms': osc_ms},
                'phase_pass': ({'achieved': need_p / (ends_ms * 1e-3) / 1e9,
                                'frac': need_p / (ends_ms * 1e-3) / fma_peak['packed'],
                                'lane_cycles_per_live_oscillator_sample': CYCLES_PHASE,
                                'live_oscillator_samples': live_phase, 'kernel_ms': ends_ms}
                               if ends_ms > 0 else None),
                'both_passes_frac': ((need_s + need_p) / ((osc_ms + ends_ms) * 1e-3) / fma_peak['packed'])}
    return {
        'bound': 'hbm',
        'kernel': 'additive_synth_kernel<NH,2>, one launch per bucket of live half-groups, concurrent (the '
                  'oscillator bank)',
        'achieved': achieved, 'peak': hbm_peak, 'unit': 'GB/s',
        'frac': (achieved / hbm_peak) if achieved else None,
        'traffic': traffic['dram_bytes'] if traffic else None,
        'traffic_source': traffic['source'] if traffic else 'no ncu capture of this build for this workload',
        'peak_source': f'{peaks_src} (MEASURED_PEAKS.json hbm_gbs)' if peaks_src == 'measured'
        else 'fallback 6650 GB/s (B200_PROFILING.md)',
        'algorithmic_bytes_per_launch': ab,
        'algorithmic_bytes_definition': 'SURVEY.md 8d: raw control tensors read once + impulse responses + dry and '
                                        'wet outputs written once = 121.4 B per output sample at configs[2]',
        'kernel_ms': osc_ms, 'kernel_share_of_step': osc_ms / m['ms'] if m['ms'] else None,
        'whole_step_GBps': ab / (m['ms'] * 1e-3) / 1e9,
        'note': 'the oscillator bank is bound by the FP32 FMA pipe, not by HBM (SURVEY.md fact 5: ~500 flop per '
                'algorithmic byte): the HBM fraction is small by construction, `fp32` is the roof that binds',
        'fp32': fp32,
        'stage_ms': m['stages'],
    }


def run_gpu(args, workload):
    gb = GpuBench(args)
    w = WORKLOADS[workload]
    steps, warmup = args.steps, max(args.warmup, 3)
    sampler = ClockSampler(gb.local)
    sampler.start()
    # FP32 peak first (also brings the clocks up before the timed region)
    eng0 = gb.group(WORKLOADS['full'])[1]
    fma_peak = {'packed': eng0.measure_fma_rate(True), 'scalar': eng0.measure_fma_rate(False)}
    if workload == 'timeline':
        m = measure_timeline(gb, w, steps, warmup)
    else:
        m = measure_clips(gb, w, steps, warmup)
    clocks = sampler.stop()
    value = m['audio_sec'] / (m['ms'] * 1e-3)
    e2e_value = m['audio_sec'] / (m['e2e_ms'] * 1e-3)
    extras = {}

    def brief(mm):
        return {'value': mm['audio_sec'] / (mm['ms'] * 1e-3), 'unit': UNIT, 'ms_per_step': mm['ms'],
                'e2e': ({'value': mm['audio_sec'] / (mm['e2e_ms'] * 1e-3), 'ms_per_step': mm['e2e_ms'],
                         'h2d_bytes_per_step': mm['h2d'], 'd2h_bytes_per_step': mm['d2h']} if 'e2e_ms' in mm else None),
                'stage_ms': mm['stages'], 'gpu_launches_per_step': mm['launches'] / max(mm.get('steps', 1), 1)}

    xsteps = max(3, min(steps, 10))
    if not args.no_extras:
        if workload == 'full':
            # same workload with note-constant inharmonicity (the reference model's behaviour)
            mh = measure_clips(gb, w, steps, 3, held=True, e2e=False)
            extras['held_notes_variant'] = dict(brief(dict(mh, steps=steps)), what=(
                'same workload, inharm_coef constant per (voice, clip) as in the reference model: partial '
                'frequencies are constant between frames'))
        if workload == 'full':
            mf = measure_clips(gb, w, steps, 3, e2e=False, fast_phase=True)
            extras['fast_phase_variant'] = dict(brief(dict(mf, steps=steps)), what=(
                'same workload with b200ddsp_config.fast_phase = 1 (opt-in): closed-form double-precision unit start '
                'phases instead of the bit-faithful phase pass; within 1e-3 of the exact signal model but NOT within '
                '1e-4 of the reference (profiles/r02_fast_phase_error_table.txt) -- never the headline'))
        if workload == 'full' and gb.world == 1:
            for name in ('dry', 'stress', 'timeline'):
                wx = WORKLOADS[name]
                mx = (measure_timeline if name == 'timeline' else measure_clips)(gb, wx, xsteps, 3)
                extras[name] = dict(brief(dict(mx, steps=xsteps)), workload=wx['name'], steps=xsteps)
                if name == 'timeline':
                    extras[name]['roofline_fp32'] = roofline_of(gb, wx, mx, fma_peak, clocks, name)['fp32']
        extras['e2e_from_conditioning'] = from_conditioning_line(gb)
        if workload == 'timeline' and gb.world > 1:
            # continuity with round 1's scaling series: independent clips, no exchange at all
            mc = measure_clips(gb, WORKLOADS['full'], xsteps, 3)
            extras['independent_clips'] = dict(brief(dict(mc, steps=xsteps)), workload=WORKLOADS['full']['name'],
                                               steps=xsteps, sharding='clips over ranks, no collective')

    if gb.rank == 0:
        cores = os.cpu_count() or 1
        cpu = None
        if gb.world == 1 and not args.no_cpu_baseline:
            import multiprocessing as mp
            with mp.get_context('spawn').Pool(cores) as pool:
                pool.map(_oracle_voice_clip, [(w['sr'], 25, w['H'], w['S'], w['M'], i) for i in range(cores)])
                sec = cpu_step(w, workload, pool)
            cpu = {'value': (w['B'] * w['F'] / 250.0) / sec, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                   'sample': cpu_sample_desc(w, workload, cores, sec), 'extrapolated': False}
        timeline = workload == 'timeline'
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': gb.world, 'steps': steps,
            'warmup': warmup, 'ms_per_step': m['ms'], 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': w['name'], 'voices': w['P'], 'substrings': w['S'], 'partials': w['H'],
                       'noise_bands': w['M'], 'frames_per_segment': w['F'], 'segments_per_gpu': w['B'],
                       'samples_per_segment': w['F'] * (w['sr'] // 250), 'reverb_taps': w['L'],
                       'sample_rate': w['sr'], 'noise': 'in-kernel Philox',
                       'phase': 'bit-faithful float32 angular_cumsum',
                       'l2': 'flushed between timed steps: 256 MB written, then 256 MB of another buffer read (cold '
                             'and clean: no write-back of the flush itself inside the timed step)',
                       'sharding': ('contiguous spans of one timeline over ranks; phase state '
                                    f'({w["P"] * w["S"] * w["H"] * 4} B) and reverb tail ({(w["L"] - 1) * 4} B) handed '
                                    'to the successor inside the kernels over NVLink peer memory (stream-ordered, '
                                    'no host barrier, no NCCL call on the data path)') if timeline else
                                   'independent clips over ranks, no collective',
                       'host_numa_binding': gb.numa_node},
            'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': m['h2d'],
                    'd2h_bytes_per_step': m['d2h'], 'ms_per_step': m['e2e_ms'],
                    'h2d_GBps_if_copy_bound': m['h2d'] / (m['e2e_ms'] * 1e-3) / 1e9,
                    'note': 'bounded by the host-to-device copy of the control tensors over PCIe; the kernels run '
                            'under the copies (DESIGN.md section 5).  Host ceiling measured with no kernels running '
                            '(scripts/h2d_ceiling.py, profiles/r02_h2d_ceiling_n*.json): 55.5 GB/s per rank alone, '
                            '29.3 with 4 ranks copying, 23.4 with 8 (one host, 237 GB/s aggregate)'},
            'gpu_launches': m['launches'], 'clocks': clocks,
            'roofline': roofline_of(gb, w, m, fma_peak, clocks, workload),
        }
        if timeline:
            line['exchange_bytes_per_step'] = m['exchange_bytes_per_step']
            line['config']['note'] = ('per-rank work = 16 segments like configs[2], plus the phase chain of EVERY partial '
                                      'through every chunk (what sounds after the span is unknown to the rank); the '
                                      'N = 1 line carries the same workload on one rank under "timeline"')
        line.update(extras)
        if cpu is not None:
            line['cpu_baseline'] = cpu
        if gb.world == 1 and workload == 'full' and not args.no_extras:
            line['config1'] = config1_line()
        print(json.dumps(line))
    if gb.world > 1:
        gb.dist.barrier()
        gb.dist.destroy_process_group()


def from_conditioning_line(gb, steps=10):
    """`e2e_from_conditioning`: what the reference pipeline actually moves to the device is the MIDI
    conditioning, not the control tensors (piano_model.py:146-164): batch 16 x 3 s of polyphonic
    conditioning [16, 750, 16, 2] + pedals (1.7 MB, pinned host memory) -> control-rate graph (dafx22.gin with
    the shipped weights, its native 16 kHz / H96 / M64 / 1.5 s IR) -> the synthesis kernels -> audio back on
    the host, all inside CUDA events.  Random notes: 4 of the 16 channels sound at any time."""
    try:
        torch = gb.torch
        B, F, P, sr = 16, 750, 16, 16000
        rng = np.random.default_rng(5 + gb.rank)
        cond = np.zeros([B, F, P, 2], np.float32)
        for b in range(B):
            for v in range(4):
                t = 0
                while t < F - 40:
                    n = int(rng.integers(40, 200))
                    cond[b, t:t + n - 10, v, 0] = rng.integers(36, 96)
                    cond[b, t, v, 1] = rng.uniform(0.3, 1.0)
                    t += n
        host = {'conditioning': torch.from_numpy(cond).pin_memory(),
                'pedal': torch.zeros([B, F, 4]).pin_memory(),
                'piano_model': torch.zeros([B, 1], dtype=torch.int64).pin_memory()}
        model = gb.dp.dafx22_model(os.path.join(ROOT, 'tests', 'golden', 'dafx22_weights.npz'), device=gb.dev,
                                   sample_rate=sr, inference=True)
        out_host = torch.empty([B, F * (sr // 250)], dtype=torch.float32).pin_memory()

        def step():
            feats = {k: v.to(gb.dev, non_blocking=True) for k, v in host.items()}
            out = model(feats)
            out_host.copy_(model.get_audio_from_outputs(out), non_blocking=True)

        for _ in range(3):
            step()
        ms = gb.reduce_max(gb.timed(step, steps)) / steps
        h2d = sum(v.numel() * v.element_size() for v in host.values())
        return {'value': gb.world * B * F / 250.0 / (ms * 1e-3), 'unit': UNIT, 'ms_per_step': ms,
                'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': out_host.numel() * 4, 'steps': steps,
                'what': 'MIDI conditioning (pinned host) -> dafx22 control-rate graph (shipped weights, 16 kHz) -> '
                        'synthesis kernels -> audio on the host; batch 16 x 3 s per GPU, 4 sounding voices of 16'}
    except Exception as e:                               # noqa: BLE001 (informational key)
        return {'error': f'{type(e).__name__}: {e}'}


def config1_line():
    """BASELINE configs[0] beside the headline (after the timed region, never part of it): one 3 s clip from
    MIDI conditioning to audio through the control-rate graph and the kernels with the shipped weights
    (scripts/config1_timing.py).  A failure here is reported in the line, it does not take the bench with it."""
    try:
        import importlib.util
        spec = importlib.util.spec_from_file_location(
            'config1_timing', os.path.join(ROOT, 'scripts', 'config1_timing.py'))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        out = mod.measure()
        out['what'] = ('configs[0]: single 3 s MIDI clip (+0.5 s warm-up), 4 notes on 16 channels, shipped weights, '
                       'MIDI conditioning -> audio; wall clock, median of 10')
        return out
    except Exception as e:                               # noqa: BLE001 (informational key)
        return {'error': f'{type(e).__name__}: {e}'}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default=None, choices=sorted(WORKLOADS))
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-extras', action='store_true')
    args = ap.parse_args()
    world = int(os.environ.get('WORLD_SIZE', str(args.gpus if args.impl == 'reference' else 1)))
    workload = args.workload or ('timeline' if max(world, args.gpus) > 1 else 'full')
    if args.impl == 'reference':
        run_reference(args, WORKLOADS[workload], workload)
    else:
        run_gpu(args, workload)


if __name__ == '__main__':
    main()
