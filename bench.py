#!/usr/bin/env python
"""Benchmark of the DDSP-Piano synthesis hot path (BASELINE.json metric: real-time factor,
audio seconds per wall second, 24 kHz, batch 16, poly 16).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload full|dry|stress]
    python bench.py --impl reference ...        # the CPU oracle on the host cores

A step = one forward of the whole polyphonic ProcessorGroup (16 voices x (additive + noise),
running sum, reverb) over one batch of synthetic control tensors.  `value` is timed on the
device with inputs resident in HBM; `e2e` goes through the reference-facing ProcessorGroup
call with pinned HOST inputs (H2D of every control tensor and D2H of the audio inside the
timed region).  Multi-GPU: the batch of clips is sharded, one process per GPU, no collective
on the data path -> weak scaling, value = all clips / max-over-ranks time.
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
    # BASELINE.json configs[2]: batch 16 x 3 s, poly 16, 96 partials, 64 noise bands, full chain
    # including the 3 s convolution reverb, 24 kHz
    'full': dict(name='configs[2]: batch16 x 3s, poly16, H96, M64, additive+noise+3s reverb, 24kHz',
                 sr=24000, B=16, P=16, S=2, H=96, M=64, F=750, L=72000),
    # configs[1]: same without the reverb
    'dry': dict(name='configs[1]: batch16 x 3s, poly16, H96, M64, additive+noise (no reverb), 24kHz',
                sr=24000, B=16, P=16, S=2, H=96, M=64, F=750, L=0),
    # configs[4]: 48 kHz / poly 32 / 128 partials stress
    'stress': dict(name='configs[4]: batch16 x 3s, poly32, H128, M96, full chain, 48kHz',
                   sr=48000, B=16, P=32, S=2, H=128, M=96, F=750, L=144000),
}
METRIC = 'real-time factor (audio-sec/wall-sec) @24kHz batch16 poly16; HBM GB/s %peak'   # BASELINE.json
UNIT = 'x real time'


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
        t = np.arange(L) / L
        x['reverb_ir'] = (rng.standard_normal([B, L]) * np.exp(-6 * t) * 1e-2).astype(np.float32)
    return x


def algorithmic_bytes(w, B, G):
    """Bytes each stage must move once (DESIGN.md 'Algorithmic bytes')."""
    P, F, H, S, M, L = w['P'], w['F'], w['H'], w['S'], w['M'], w['L']
    U = w['sr'] // 250
    N = F * U
    n_chunks = -(-N // 1000)
    R = P * B
    return {
        'forward': R * F * (1 + H + 1 + S + M) * 4 + B * L * 4 + B * N * 4 * (2 if L else 1),
        'oscillators': R * F * (1 + 2 * H + S) * 4 + R * S * n_chunks * H * 4 + G * B * N * 4,
        'n_samples': B * N,
    }


def live_chain_samples(w, x):
    """Oscillator-samples the synthesis kernels actually run for these inputs: the kernels drop
    16-partial half-groups that lie above Nyquist (or belong to a muted voice) in every frame a
    1000-sample chunk touches, and a unit with nh live half-groups runs nh chains
    on each of the warp's 32 lanes.  Host-side restatement of that bookkeeping (the liveness rule
    of the controls kernel: a partial can sound if f0 n sqrt(1 + B n^2) < sr / 2 and f0 > 20 Hz)."""
    P, F, H, S = w['P'], w['F'], w['H'], w['S']
    sr = w['sr']
    U = sr // 250
    N = F * U
    f0 = x['f0_hz'][..., 0].astype(np.float64)                           # [P, B, F]
    binh = np.maximum(x['inharm_coef'][..., 0].astype(np.float64), 0.0)
    n = np.arange(1, H + 1, dtype=np.float64)
    freq = f0[..., None] * n * np.sqrt(1.0 + binh[..., None] * n * n)    # [P, B, F, H]
    can = (freq < sr / 2.0) & (f0[..., None] > 20.0)
    top = np.where(can.any(-1), H - np.argmax(can[..., ::-1], axis=-1), 0)   # 1 + highest live partial
    nh_frame = -(-top // 16)                                             # [P, B, F]
    total = 0
    for t0 in range(0, N, 1000):
        t1 = min(N, t0 + 1000) - 1
        k0, k1 = t0 // U, min(F - 1, t1 // U + 1)
        nh = nh_frame[..., k0:k1 + 1].max(-1)                            # [P, B]
        lanes_per_string = 16 if S % 2 == 0 else 32
        chains = nh if S % 2 == 0 else -(-nh // 2)
        total += int(chains.sum()) * lanes_per_string * S * (t1 + 1 - t0)
    return total


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
# CPU side: the oracle as the reported baseline / reference arm
# ------------------------------------------------------------------------------------------

def _oracle_voice_clip(args):
    """One (voice, clip) of the workload through the numpy oracle: get_controls + get_signal of
    the additive and noise processors.  Returns seconds."""
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
    t0 = time.perf_counter()
    c = ref.additive_controls(amp, hd, inh, f0, sample_rate=sr)
    a = ref.additive_signal(**c, sample_rate=sr, inference=True)
    n = ref.noise_signal(ref.noise_controls(mags)['magnitudes'], noise)
    _ = a + n
    return time.perf_counter() - t0


def _oracle_reverb(sr, F, L, seed):
    from oracle import ddsp_piano_np as ref
    rng = np.random.default_rng(seed)
    N = F * (sr // 250)
    audio = rng.standard_normal([1, N], dtype=np.float32)
    ir = rng.standard_normal([1, L], dtype=np.float32)
    t0 = time.perf_counter()
    ref.reverb_signal(audio, ir)
    return time.perf_counter() - t0


def cpu_sample(w, n_voice_clips, cores, pool=None):
    """Time `n_voice_clips` (voice, clip) units of the workload through the oracle on `cores`
    worker processes and scale to the whole batch.  Returns (rtf, seconds, description)."""
    sr, F, H, S, M, L, B, P = (w[k] for k in ('sr', 'F', 'H', 'S', 'M', 'L', 'B', 'P'))
    jobs = [(sr, F, H, S, M, 1000 + i) for i in range(n_voice_clips)]
    t0 = time.perf_counter()
    if pool is not None:
        pool.map(_oracle_voice_clip, jobs, chunksize=1)
    else:
        for j in jobs:
            _oracle_voice_clip(j)
    t_voices = time.perf_counter() - t0
    t_rev = _oracle_reverb(sr, F, L, 5) if L else 0.0
    # whole batch = P*B voice-clips (embarrassingly parallel over `cores`) + B reverbs
    t_full = t_voices * (P * B) / n_voice_clips + t_rev * B / max(cores, 1)
    audio_sec = B * F / 250.0
    desc = (f'{n_voice_clips} of {P * B} (voice, clip) units (get_controls+get_signal, additive+'
            f'noise) on {cores} worker process(es) in {t_voices:.2f}s'
            + (f' + 1 of {B} reverbs in {t_rev:.3f}s' if L else '') + ', scaled to the whole batch')
    return audio_sec / t_full, t_voices + t_rev, desc


def run_reference(args, w):
    """--impl reference: the reference's TF2 CPU path is not installable here (tensorflow/ddsp
    absent, no network); its CPU restatement (oracle/, numpy) is timed instead, one worker
    process per host core, on a bounded sample of the same workload."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    per_step = 2 * cores            # two (voice, clip) units per core per step: ~1.5 s of wall
    vals, secs, desc = [], [], ''
    with mp.get_context('spawn').Pool(cores) as pool:
        pool.map(_oracle_voice_clip, [(w['sr'], 25, w['H'], w['S'], w['M'], i)
                                      for i in range(cores)])      # import numpy in the workers
        for _ in range(min(args.warmup, 2)):
            cpu_sample(w, cores, cores, pool)
        for _ in range(args.steps):
            rtf, s, desc = cpu_sample(w, per_step, cores, pool)
            vals.append(rtf)
            secs.append(s)
    value = float(np.mean(vals))
    audio_sec = w['B'] * w['F'] / 250.0
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * audio_sec / value,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
        'data': 'synthetic', 'config': {'workload': w['name']},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                         'sample': desc},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
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
        import subprocess
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


def run_gpu(args, w):
    import torch
    import torch.distributed as dist
    import __graft_entry__
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if rank == 0:
        __graft_entry__.build()
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
        dist.barrier()
    import ddsp_piano_b200 as dp
    from ddsp_piano_b200.processors import _DEFAULT_CFG
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    numa_node = bind_to_gpu_numa_node(local) if world > 1 else None
    sr, B, P, S, H, M, F, L = (w[k] for k in ('sr', 'B', 'P', 'S', 'H', 'M', 'F', 'L'))
    U = sr // 250
    N = F * U

    # weak scaling: every rank synthesises its own batch of B clips (independent MIDI segments)
    x_np = synthetic_inputs(w, seed=rank)
    host = {k: torch.from_numpy(v).pin_memory() for k, v in x_np.items()}
    resident = {k: v.to(dev) for k, v in host.items()}

    additive = dp.MultiInharmonic(frame_rate=250, sample_rate=sr, inference=True, name='additive')
    noise = dp.DynamicSizeFilteredNoise(frame_rate=250, sample_rate=sr, name='noise', seed=1234)
    reverb = dp.Reverb(trainable=False) if L else None
    group = dp.ProcessorGroup(dag=dp.polyphonic_dag(
        additive=additive, noise=noise, reverb=reverb,
        additive_controls=['amplitudes', 'harmonic_distribution', 'inharm_coef', 'f0_hz'],
        noise_controls=['magnitudes'], reverb_controls=['reverb_ir'] if L else [], n_synths=P))

    def features(parents):
        f = {f'{k}_{v}': parents[k][v] for k in ('amplitudes', 'harmonic_distribution',
                                                 'inharm_coef', 'f0_hz', 'magnitudes')
             for v in range(P)}
        if L:
            f['reverb_ir'] = parents['reverb_ir']
        return f

    cfg = {**_DEFAULT_CFG, **additive.engine_config(), **noise.engine_config(M)}
    if reverb is not None:
        cfg.update(reverb.engine_config())
    eng = dp.get_engine(dev, **cfg)
    eng.set_profiling(True)

    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)   # > 126 MB L2

    # the per-voice views exist once (Parallelizer.unparallelize hands them out the same way);
    # every step passes a fresh shallow copy of the dict because ProcessorGroup extends it
    feats_resident = features(resident)
    feats_host = features(host)

    def step_resident():
        return group(dict(feats_resident), return_outputs_dict=False)

    h2d = sum(v.numel() * 4 for v in host.values())

    def step_e2e():
        # pinned HOST control tensors in, pinned HOST audio out: the ProcessorGroup routes CPU
        # features to b200ddsp_forward_polyphonic_host (H2D + kernels + D2H on the timed stream)
        return group(dict(feats_host), return_outputs_dict=False)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, stages=None):
        """K steps, each bracketed by CUDA events on the launching stream, L2 flushed between
        steps (outside the brackets).  Returns total ms."""
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
               for _ in range(steps)]
        barrier()
        for i in range(steps):
            flush.zero_()
            evs[i][0].record()
            fn()
            evs[i][1].record()
            if stages is not None:
                s = eng.last_stage_ms()
                for k, v in s.items():
                    stages[k] = stages.get(k, 0.0) + v
        barrier()
        return sum(a.elapsed_time(b) for a, b in evs)

    # clocks / throttle reasons are sampled from the warm-up to the end of the last timed loop
    # (the resident loop alone lasts ~60 ms, shorter than nvidia-smi's start-up)
    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(max(args.warmup, 3)):
        step_resident()
    launches0 = dp.total_launches()
    stages = {}
    total_ms = timed(step_resident, args.steps, stages)
    launches = dp.total_launches() - launches0

    for _ in range(3):
        step_e2e()
    e2e_ms = timed(step_e2e, args.steps)

    # same workload with note-constant inharmonicity (the reference model's behaviour)
    held = {k: torch.from_numpy(v).to(dev)
            for k, v in synthetic_inputs(w, seed=rank, held_notes=True).items()}
    feats_held = features(held)
    for _ in range(3):
        group(dict(feats_held), return_outputs_dict=False)
    held_stages = {}
    held_ms = timed(lambda: group(dict(feats_held), return_outputs_dict=False), args.steps, held_stages)
    clocks = sampler.stop()

    def reduce_max(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    total_ms = reduce_max(total_ms)
    e2e_ms = reduce_max(e2e_ms)
    held_ms = reduce_max(held_ms)
    audio_sec = world * B * F / 250.0
    ms_per_step = total_ms / args.steps
    value = audio_sec / (ms_per_step * 1e-3)
    e2e_value = audio_sec / (e2e_ms / args.steps * 1e-3)

    if rank == 0:
        peaks, peaks_src = {}, 'fallback'
        try:
            with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
                peaks, peaks_src = json.load(f), 'measured'
        except Exception:
            pass
        hbm_peak = float(peaks.get('hbm_gbs', 6650.0))
        G = 1
        while G < P and (-(-N // 1000)) * B * G < 4 * 148:
            G *= 2
        ab = algorithmic_bytes(w, B, min(G, P))
        osc_ms = stages.get('oscillators', 0.0) / args.steps
        achieved = ab['oscillators'] / (osc_ms * 1e-3) / 1e9 if osc_ms > 0 else None
        # issue-slot view of the same kernel: counted FP32 instructions per oscillator-sample
        osc_samples = B * P * S * H * N
        # dram__bytes_read+write of the six bucket launches of one step, from the ncu --set full
        # capture of this same command (profiles/r01_prof10_summary.txt); config 3 at N=1 only
        traffic = 101.7e6 if (args.workload == 'full' and world == 1) else None
        live = live_chain_samples(w, synthetic_inputs(w, seed=rank))
        ends_ms = stages.get('phase_ends', 0.0) / args.steps
        sm_hz = (clocks.get('sm_mhz') or 1965.0) * 1e6
        roofline = {
            'bound': 'hbm',
            'kernel': 'additive_synth_kernel<NH,2> x 6 buckets, concurrent (the oscillator bank)',
            'achieved': achieved, 'peak': hbm_peak, 'unit': 'GB/s',
            'frac': (achieved / hbm_peak) if achieved else None, 'traffic': traffic,
            'traffic_source': 'profiles/r01_prof10_summary.txt (sum over the 6 bucket launches; below '
                              'the algorithmic bytes because silent partial groups are never read)',
            'peak_source': f'{peaks_src} (MEASURED_PEAKS.json hbm_gbs)' if peaks_src == 'measured'
            else 'fallback 6650 GB/s (B200_PROFILING.md)',
            'algorithmic_bytes_per_launch': ab['oscillators'], 'kernel_ms': osc_ms,
            'kernel_share_of_step': osc_ms / ms_per_step if ms_per_step else None,
            'note': 'the oscillator bank is bound by the FP32 (FMA) pipe, not by HBM (SURVEY.md fact 5): '
                    'the bit-faithful phase chain needs 11.5 FMA-pipe cycles per oscillator-sample (+6 in the phase pass); see '
                    'fma_pipe, oscillator_samples_per_s and DESIGN.md section 4',
            'oscillator_samples_per_s': osc_samples / (osc_ms * 1e-3) if osc_ms > 0 else None,
            # the bound that matters, measured live: FMA-pipe cycles the bit-faithful algorithm needs
            # for the oscillator-samples on live lanes (11.5 per sample in the synthesis pass, 6 in
            # the phase pass) over the pipe cycles available in the measured time
            # (148 SMs x 4 schedulers x 32 lanes at the SM clock under load)
            'fp32_pipe': ({'live_oscillator_samples': live,
                           'synthesis': {'needed_lane_cycles': live * 11.5,
                                         'frac': live * 11.5 / (osc_ms * 1e-3 * 148 * 128 * sm_hz)},
                           'phase_pass': ({'needed_lane_cycles': live * 6.0 * (1 - 1000.0 / N),
                                           'frac': live * 6.0 * (1 - 1000.0 / N) /
                                                   (ends_ms * 1e-3 * 148 * 128 * sm_hz)}
                                          if ends_ms > 0 else None),
                           'peak_lane_cycles_per_s': 148 * 128 * sm_hz}
                          if osc_ms > 0 else None),
            # what actually bounds the stage.  Static facts from the ncu capture of this command
            # (profiles/r01_prof10_summary.txt, config 3): warp instructions of the six bucket
            # launches and the FMA-pipe activity of the two dominant kernels; live: the share of
            # the issue slots of the measured time those instructions fill
            'fma_pipe': ({'ncu_pipe_fma_cycles_active_pct': {'additive_synth_kernel<6,2> (largest bucket)': 61.0,
                                                             'additive_fast_kernel<2,ends> (phase pass)': 77.0},
                          'warp_instructions_per_launch_set': 8.560e8,
                          'issue_slots_per_s': 148 * 4 * sm_hz,
                          'issue_frac': 8.560e8 / (osc_ms * 1e-3) / (148 * 4 * sm_hz),
                          'source': 'profiles/r01_prof10_summary.txt'}
                         if (args.workload == 'full' and world == 1 and osc_ms > 0) else None),
            'whole_step_GBps': ab['forward'] / (ms_per_step * 1e-3) / 1e9,
            'stage_ms': {k: v / args.steps for k, v in stages.items()},
        }
        cores = os.cpu_count() or 1
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            rtf, secs, desc = cpu_sample(w, 16, 1)
            cpu = {'value': rtf, 'unit': UNIT, 'cores': 1, 'kind': 'port', 'sample': desc,
                   'host_cores_available': cores}
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
            'warmup': max(args.warmup, 3), 'ms_per_step': ms_per_step, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': w['name'], 'clips_per_gpu': B, 'voices': P, 'substrings': S,
                       'partials': H, 'noise_bands': M, 'frames': F, 'samples_per_clip': N,
                       'reverb_taps': L, 'sample_rate': sr, 'noise': 'in-kernel Philox',
                       'phase': 'bit-faithful float32 angular_cumsum',
                       'l2': 'flushed (256 MB write) between timed steps',
                       'sharding': 'clips over ranks, no collective',
                       'host_numa_binding': numa_node},
            'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': h2d,
                    'd2h_bytes_per_step': B * N * 4, 'ms_per_step': e2e_ms / args.steps,
                    'h2d_GBps_if_copy_bound': h2d / (e2e_ms / args.steps * 1e-3) / 1e9,
                    'note': 'bounded by the host-to-device copy of the control tensors over PCIe; '
                            'the kernels run under the copies (DESIGN.md section 5)'},
            'gpu_launches': launches, 'clocks': clocks, 'roofline': roofline,
            'held_notes_variant': {
                'what': 'same workload, inharm_coef constant per (voice, clip) as in the reference '
                        'model: partial frequencies are constant between frames',
                'value': audio_sec / (held_ms / args.steps * 1e-3), 'unit': UNIT,
                'ms_per_step': held_ms / args.steps,
                'stage_ms': {k: v / args.steps for k, v in held_stages.items()}},
        }
        if cpu is not None:
            line['cpu_baseline'] = cpu
        if world == 1 and args.workload == 'full':
            line['config1'] = config1_line()
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def config1_line():
    """BASELINE configs[0] beside the headline (after the timed region, never part of it): one 3 s clip from
    MIDI conditioning to audio through the control-rate graph and the kernels with the shipped weights
    (scripts/config1_timing.py).  A failure here is reported in the line, it does not take the bench with it."""
    try:
        import importlib.util
        spec = importlib.util.spec_from_file_location(
            'config1_timing', os.path.join(os.path.dirname(os.path.abspath(__file__)), 'scripts', 'config1_timing.py'))
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
    ap.add_argument('--workload', default='full', choices=sorted(WORKLOADS))
    ap.add_argument('--no-cpu-baseline', action='store_true')
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    if args.impl == 'reference':
        run_reference(args, w)
    else:
        run_gpu(args, w)


if __name__ == '__main__':
    main()
