"""BASELINE config 4 on the GPU: one timeline synthesised as consecutive spans.

The reference runs a whole piece through ONE pass (synthesize_midi_file.py:52-54,73), so a span must
reproduce, for its samples, exactly what the whole-timeline call produces: the tests compare the span
calls BIT FOR BIT with the whole-timeline call of the same kernels (which test_gpu_parity.py pins to the
oracle), the carried phase state bit for bit with the oracle's segment form, and the timeline reverb
with the float64 convolution of the concatenated dry signal.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

torch = pytest.importorskip('torch')

from oracle import ddsp_piano_np as ref            # noqa: E402  (checker only)
from test_gpu_parity import TIGHT, cu, rel_err, voice_inputs, dp, dev   # noqa: E402,F401

pytestmark = pytest.mark.gpu

KEYS = ('amplitudes', 'harmonic_distribution', 'inharm_coef', 'f0_hz', 'magnitudes')


def engine_for(dp, dev, sr, M, **extra):
    from ddsp_piano_b200.processors import _DEFAULT_CFG
    return dp.get_engine(dev, **{**_DEFAULT_CFG, 'sample_rate': sr, 'n_noise_bands': M, **extra})


def timeline_inputs(seed, P, B, F, H, S, M, U, with_noise=True):
    rng = np.random.default_rng(seed)
    voices = [voice_inputs(rng, B, F, H, S, M) for _ in range(P)]
    if with_noise:
        for v in voices:
            v['noise'] = rng.uniform(-1, 1, [B, F * U]).astype(np.float32)
    return voices


def slice_voices(voices, k0, k1, U, dev):
    out = []
    for v in voices:
        d = {k: cu(v[k][:, k0:k1], dev) for k in KEYS}
        if 'noise' in v:
            d['noise'] = cu(v['noise'][:, k0 * U:k1 * U], dev)
        out.append(d)
    return out


def run_spans(dp, eng, voices, bounds, total, U, dev, P, B, S, H, seed=0):
    """Synthesise the timeline span by span on one GPU (local links: stream order is the hand-off).
    Returns (dry [B, total * U], list of carries [P, B, S, H] after every span)."""
    from ddsp_piano_b200 import _lib, sharding
    carries, pieces = [], []
    prev = None
    for i, (o0, o1) in enumerate(bounds):
        k0, k1 = max(o0 - 1, 0), min(o1 + 1, total)
        carry = torch.full([P, B, S, H], float('nan'), dtype=torch.float32, device=dev)
        span = _lib.Span(in_first_frame=k0, out_first_frame=o0, n_out_frames=o1 - o0, total_frames=total,
                         phase=sharding.local_link(seed=prev, carry=carry, epoch=i + 1))
        pieces.append(eng.forward_span(slice_voices(voices, k0, k1, U, dev), span, seed=seed))
        carries.append(carry)
        prev = carry
    return torch.cat(pieces, dim=1), carries


def test_offset_scan_in_several_tiles(dp, dev, monkeypatch):
    """The chunk-offset scan keeps a whole span in one shared-memory tile when it fits (1536 chunks); longer
    spans go through it tile by tile.  Forced here with tiles of 7 chunks on a 36-chunk clip and on spans with
    a carried phase state: same bits."""
    sr, H, S, M, P, B = 24000, 96, 2, 64, 3, 2
    U = sr // 250
    bounds = [(0, 125), (125, 375)]
    total = bounds[-1][1]
    voices = timeline_inputs(21, P, B, total, H, S, M, U, with_noise=False)
    eng = engine_for(dp, dev, sr, M)
    whole, _ = eng.forward_polyphonic(slice_voices(voices, 0, total, U, dev), seed=9)
    spans, carries = run_spans(dp, eng, voices, bounds, total, U, dev, P, B, S, H, seed=9)
    monkeypatch.setenv('B200DDSP_OFFSETS_TILE', '7')
    whole7, _ = eng.forward_polyphonic(slice_voices(voices, 0, total, U, dev), seed=9)
    spans7, carries7 = run_spans(dp, eng, voices, bounds, total, U, dev, P, B, S, H, seed=9)
    torch.cuda.synchronize()
    assert torch.equal(whole7, whole) and torch.equal(spans7, spans) and torch.equal(spans, whole)
    for a, b in zip(carries, carries7):
        assert torch.equal(a, b)


@pytest.mark.parametrize('sr,H,S,M,bounds', [
    (24000, 96, 2, 64, [(0, 375), (375, 1000), (1000, 1500)]),     # BASELINE shapes, unequal spans
    (24000, 128, 1, 96, [(0, 125), (125, 250), (250, 500)]),       # v2 shapes: one string, 190-tap noise FIR
    (48000, 128, 2, 96, [(0, 250), (250, 375)]),                   # stress shapes, U = 192
])
def test_spans_equal_the_whole_timeline_bit_for_bit(dp, dev, sr, H, S, M, bounds):
    P, B = 3, 2
    U = sr // 250
    total = bounds[-1][1]
    eng = engine_for(dp, dev, sr, M)
    voices = timeline_inputs(sr + H, P, B, total, H, S, M, U)
    whole, _ = eng.forward_polyphonic(slice_voices(voices, 0, total, U, dev))
    got, carries = run_spans(dp, eng, voices, bounds, total, U, dev, P, B, S, H)
    torch.cuda.synchronize()
    assert float(whole.abs().max()) > 1e-3
    assert torch.equal(got, whole)
    # the same with the noise drawn in-kernel: the Philox counter is the GLOBAL sample index
    for v in voices:
        del v['noise']
    whole_p, _ = eng.forward_polyphonic(slice_voices(voices, 0, total, U, dev), seed=77)
    got_p, _ = run_spans(dp, eng, voices, bounds, total, U, dev, P, B, S, H, seed=77)
    assert torch.equal(got_p, whole_p)
    assert not torch.equal(whole_p, whole)
    assert all(bool(torch.isfinite(c).all()) for c in carries)


def test_span_phase_state_equals_the_oracle_segment_form(dp, dev):
    """The state a span hands on == oracle/ddsp_piano_np.py::additive_signal_segment's carry, bit for
    bit, and the span's additive audio is within tolerance of the oracle's segment audio."""
    sr, P, B, H, S, M = 24000, 2, 1, 96, 2, 64
    U = sr // 250
    bounds = [(0, 125), (125, 375), (375, 500)]
    total = bounds[-1][1]
    eng = engine_for(dp, dev, sr, M)
    voices = timeline_inputs(5, P, B, total, H, S, M, U)
    for v in voices:
        v['magnitudes'][:] = -40.0                     # noise off: the dry span is the additive signal
    got, carries = run_spans(dp, eng, voices, bounds, total, U, dev, P, B, S, H)
    torch.cuda.synchronize()
    want_audio = np.zeros([B, total * U], np.float32)
    for vi, v in enumerate(voices):
        ctl = ref.additive_controls(v['amplitudes'], v['harmonic_distribution'], v['inharm_coef'], v['f0_hz'],
                                    sample_rate=sr)
        carry = None
        for i, fr in enumerate(bounds):
            audio, carry = ref.additive_signal_segment(**ctl, frames=fr, carry=carry, sample_rate=sr)
            want_audio[:, fr[0] * U:fr[1] * U] += audio
            # CUDA payload is [P, B, S, H]; the oracle's is [S, B, H]
            np.testing.assert_array_equal(carries[i][vi].permute(1, 0, 2).cpu().numpy(), carry)
    assert rel_err(got, want_audio) < TIGHT


def test_span_argument_checks(dp, dev):
    from ddsp_piano_b200 import _lib, sharding
    sr, P, B, H, S, M = 24000, 1, 1, 32, 2, 64
    U = sr // 250
    eng = engine_for(dp, dev, sr, M)
    voices = timeline_inputs(1, P, B, 300, H, S, M, U, with_noise=False)
    seed = torch.zeros([P, B, S, H], device=dev)

    def call(k0, k1, o0, n, total=300, with_seed=True):
        span = _lib.Span(in_first_frame=k0, out_first_frame=o0, n_out_frames=n, total_frames=total,
                         phase=sharding.local_link(seed=seed if with_seed else None))
        return eng.forward_span(slice_voices(voices, k0, k1, U, dev), span)

    call(124, 251, 125, 125)                                        # fine
    with pytest.raises(ValueError, match='halo is needed before'):
        call(125, 251, 125, 125)
    with pytest.raises(ValueError, match='halo is needed after'):
        call(124, 250, 125, 125)
    with pytest.raises(ValueError, match='multiple of the 1000-sample chunk'):
        call(99, 226, 100, 125)
    with pytest.raises(ValueError, match='needs a phase seed'):
        call(124, 251, 125, 125, with_seed=False)
    with pytest.raises(ValueError, match='of a timeline of'):
        call(124, 251, 125, 200)
    call(249, 300, 250, 50)                                         # the end of the timeline needs no halo


def float64_reverb(dry, ir, add_dry=True):
    """ddsp.effects.Reverb on the whole timeline in float64 (FFT convolution)."""
    from scipy.signal import fftconvolve
    h = ir.astype(np.float64).copy()
    h[:, 0] = 0
    x = dry.astype(np.float64)
    wet = np.stack([fftconvolve(x[b], h[b])[:x.shape[1]] for b in range(x.shape[0])])
    return wet + x if add_dry else wet


@pytest.mark.parametrize('B,n_seg,N,L,n_spans', [(2, 3, 2400, 5000, 3),     # tail covers two segments
                                               (1, 4, 7200, 7200, 2),
                                               (3, 1, 4096, 1000, 4)])
def test_timeline_reverb_spans_on_one_gpu(dp, dev, B, n_seg, N, L, n_spans):
    """b200ddsp_timeline_reverb span after span (local tail links) == reverb of the whole timeline."""
    from ddsp_piano_b200 import sharding
    eng = engine_for(dp, dev, 24000, 64)
    rng = np.random.default_rng(B * 100 + n_seg)
    span = n_seg * N
    dry = (rng.standard_normal([B, n_spans * span]) * 0.1).astype(np.float32)
    ir = (rng.standard_normal([B, L]) * np.exp(-6 * np.arange(L) / L) * 1e-2).astype(np.float32)
    want = float64_reverb(dry, ir)
    ir_d = cu(ir, dev)
    out, prev = [], None
    for i in range(n_spans):
        carry = torch.full([B, L - 1], float('nan'), dtype=torch.float32, device=dev)
        link = sharding.local_link(seed=prev, carry=carry, epoch=i + 1)
        out.append(eng.timeline_reverb(cu(dry[:, i * span:(i + 1) * span], dev), ir_d, n_seg, tail=link))
        prev = carry
    got = torch.cat(out, dim=1)
    assert rel_err(got, want) < TIGHT
    # without a successor nothing is produced, without a predecessor nothing is consumed
    alone = eng.timeline_reverb(cu(dry[:, :span], dev), ir_d, n_seg)
    assert torch.equal(alone, out[0])


def test_forward_timeline_on_one_gpu(dp, dev):
    """The fused call (span forward + timeline reverb) span after span, device and host inputs, against
    the whole-timeline dry signal (bit for bit) and its float64 reverb."""
    from ddsp_piano_b200 import _lib, sharding
    sr, P, B, H, S, M, L, seg = 24000, 3, 1, 96, 2, 64, 9000, 125
    U = sr // 250
    bounds = [(0, 250), (250, 500), (500, 750)]
    total = bounds[-1][1]
    eng = engine_for(dp, dev, sr, M)
    voices = timeline_inputs(9, P, B, total, H, S, M, U, with_noise=False)
    rng = np.random.default_rng(2)
    ir = (rng.standard_normal([B, L]) * np.exp(-6 * np.arange(L) / L) * 1e-2).astype(np.float32)
    whole, _ = eng.forward_polyphonic(slice_voices(voices, 0, total, U, dev), seed=5)
    want_wet = float64_reverb(whole.cpu().numpy(), ir)
    for host in (False, True):
        dries, wets, prev_p, prev_t = [], [], None, None
        for i, (o0, o1) in enumerate(bounds):
            k0, k1 = max(o0 - 1, 0), min(o1 + 1, total)
            cp = torch.empty([P, B, S, H], dtype=torch.float32, device=dev)
            ct = torch.empty([B, L - 1], dtype=torch.float32, device=dev)
            span = _lib.Span(in_first_frame=k0, out_first_frame=o0, n_out_frames=o1 - o0, total_frames=total,
                             phase=sharding.local_link(seed=prev_p, carry=cp, epoch=i + 1))
            tail = sharding.local_link(seed=prev_t, carry=ct, epoch=i + 1)
            vs = slice_voices(voices, k0, k1, U, dev)
            irt = cu(ir, dev)
            if host:
                vs = [{k: t.cpu().pin_memory() for k, t in v.items()} for v in vs]
                irt = torch.from_numpy(ir).pin_memory()
            dry, wet = eng.forward_timeline(vs, irt, span, seg, tail=tail, seed=5)
            torch.cuda.synchronize()
            dries.append(dry.clone().to(dev))
            wets.append(wet.clone().to(dev))
            prev_p, prev_t = cp, ct
        assert torch.equal(torch.cat(dries, dim=1), whole)
        assert rel_err(torch.cat(wets, dim=1), want_wet) < TIGHT


def test_timeline_two_gpus_peer_memory():
    """Config 4 across real GPUs: spans on 2 ranks, phase state and reverb tail handed over inside the
    kernels through NVLink peer memory (skipped on a single-GPU box; the driver's scaling run and
    profiles/ hold the multi-GPU evidence)."""
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    proc = subprocess.run(
        [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
         '--master-addr', '127.0.0.1', '--master-port', '29517',
         os.path.join(root, 'tests', 'multi_gpu_timeline.py')],
        capture_output=True, text=True, timeout=900)
    assert proc.returncode == 0 and 'TIMELINE_OK' in proc.stdout, proc.stdout + proc.stderr
