"""Control-rate graph of dafx22.gin: ddsp_piano_b200/model.py against the numpy
restatement oracle/piano_model_np.py on the shipped weights (tests/golden/dafx22_weights.npz,
exported by tests/golden/make_model_weights.py).  The Keras/ddsp layer semantics underneath both
are restated, not pinned (see the oracle's header)."""
import os

import numpy as np
import pytest
import torch

from oracle import piano_model_np as ref

HERE = os.path.dirname(os.path.abspath(__file__))
WEIGHTS = os.path.join(HERE, 'golden', 'dafx22_weights.npz')


@pytest.fixture(scope='module')
def weights():
    from ddsp_piano_b200.checkpoint import NpzWeights
    return NpzWeights(WEIGHTS)


def midi_clip(B=2, T=300, P=16, seed=0):
    """A few notes per clip on the first voices, the rest silent (pitch 0), pedal ramps."""
    rng = np.random.default_rng(seed)
    cond = np.zeros([B, T, P, 2], np.float32)
    for b in range(B):
        for v in range(4):
            t = int(rng.integers(0, T // 4))
            while t < T - 20:
                dur = int(rng.integers(20, 120))
                cond[b, t:t + dur, v, 0] = rng.integers(30, 100)
                cond[b, t, v, 1] = rng.uniform(0.2, 1.0)
                t += dur + int(rng.integers(5, 60))
    pedal = np.zeros([B, T, 4], np.float32)
    pedal[:, T // 3:2 * T // 3, 0] = 0.8
    return cond, pedal, np.arange(B).reshape(B, 1) % 2


# ---- oracle known answers -------------------------------------------------------------------

def test_note_release_holds_the_note_for_the_release_time():
    T, rel = 400, 0.4                                   # 100 frames at 250 Hz
    pitch = np.zeros([1, T, 1], np.float32)
    pitch[0, 10:30, 0] = 60
    out = ref.note_release(pitch, rel, 250)[0, :, 0]
    assert np.all(out[:10] == 0) and np.all(out[10:30] == 60)
    # the counter starts at the first silent frame: held while steps <= 100, i.e. 101 more frames
    assert np.all(out[30:131] == 60) and np.all(out[131:] == 0)


def test_note_release_new_note_replaces_the_held_one():
    pitch = np.zeros([1, 200, 1], np.float32)
    pitch[0, 5:10, 0] = 50
    pitch[0, 40:45, 0] = 72
    out = ref.note_release(pitch, 1.0, 250)[0, :, 0]
    assert np.all(out[10:40] == 50) and np.all(out[40:] == 72)


def test_inharmonicity_and_detuner_known_answers(weights):
    w = ref.load_weights(weights)
    pitch = np.array([[[21.0], [60.0], [108.0]]], np.float32)
    b = ref.inharmonicity(pitch, None, w['inharm'])[0, :, 0]
    # Rigaud et al. tessitura model: a valley between the bass and treble bridges
    assert b[1] < b[0] and b[1] < b[2] and 1e-5 < b[1] < 1e-3
    f0 = ref.detuner(pitch, None, np.zeros([1, 2], np.float32), np.zeros(2, np.float32))
    np.testing.assert_allclose(f0[0, 1], 440.0 * 2 ** ((60 - 69) / 12), rtol=1e-6)
    # silent voices (pitch 0) land at 8.18 Hz, below min_frequency: the synth mutes them
    assert ref.detuner(np.zeros([1, 1, 1], np.float32), None, *w['detuner'])[0, 0, 0] < 20.0


def test_gru_restatement_matches_torch_gru():
    from ddsp_piano_b200.model import GRU
    rng = np.random.default_rng(1)
    i, u = 7, 12
    k, r, b = (rng.standard_normal(s).astype(np.float32) * 0.4 for s in ([i, 3 * u], [u, 3 * u], [2, 3 * u]))
    x = rng.standard_normal([3, 40, i]).astype(np.float32)
    want = ref.gru(x, k, r, b)
    got = GRU(k, r, b, torch.device('cpu'))(torch.from_numpy(x)).numpy()
    np.testing.assert_allclose(got, want, atol=2e-6)


def test_controls_on_cpu_match_the_oracle(weights, monkeypatch):
    """The torch graph (on CPU, with the note-release kernel replaced by the oracle's recurrence)
    against the numpy restatement, shipped weights."""
    from ddsp_piano_b200 import model as M
    cond, pedal, pm = midi_clip()
    w = ref.load_weights(weights)
    want = ref.control_graph(cond, pedal, pm, w)

    class CpuRelease:
        def __init__(self, dur, fr):
            self.dur, self.fr = dur, fr

        def __call__(self, conditioning):
            return torch.from_numpy(ref.note_release(conditioning[..., 0:1].numpy(), self.dur, self.fr))

    class CpuReverb:
        def __init__(self, emb, **kw):
            self.emb = emb

        def __call__(self, pm_):
            return self.emb[torch.as_tensor(pm_).long().reshape(-1)]

    monkeypatch.setattr(M, 'NoteRelease', CpuRelease)
    monkeypatch.setattr(M, 'MultiInstrumentReverb', CpuReverb)
    model = M.dafx22_model(weights, device='cpu')
    got = model.compute_controls({'conditioning': cond, 'pedal': pedal, 'piano_model': pm})
    for key in ('f0_hz', 'inharm_coef', 'amplitudes', 'harmonic_distribution', 'magnitudes'):
        g, r_ = got[key].numpy(), want[key]
        assert g.shape == r_.shape, key
        err = np.max(np.abs(g - r_)) / np.max(np.abs(r_))
        assert err < 1e-4, (key, err)
        assert got[f'{key}_3'].data_ptr() == got[key][3].data_ptr()          # per-voice views
    assert got['reverb_ir'].shape == (2, 24000)


# ---- GPU -------------------------------------------------------------------------------------

@pytest.mark.gpu
def test_note_release_kernel_is_bit_exact():
    import ddsp_piano_b200 as dp
    dev = torch.device('cuda:0')
    cond, _, _ = midi_clip(B=3, T=700, seed=3)
    rows = cond.transpose(2, 0, 1, 3).reshape(-1, 700, 2)
    want = ref.note_release(rows[..., 0:1], 1.0, 250)
    eng = dp.get_engine(dev, **dp.processors._DEFAULT_CFG)
    got = eng.note_release(torch.from_numpy(rows).to(dev), 250.0).cpu().numpy()
    assert np.array_equal(got, want)
    got1 = eng.note_release(torch.from_numpy(rows[..., 0:1].copy()).to(dev), 250.0).cpu().numpy()
    assert np.array_equal(got1, want)


@pytest.mark.gpu
def test_dafx22_controls_match_the_oracle_on_gpu(weights):
    import ddsp_piano_b200 as dp
    cond, pedal, pm = midi_clip(B=2, T=300)
    want = ref.control_graph(cond, pedal, pm, ref.load_weights(weights))
    model = dp.dafx22_model(weights, device='cuda:0')
    got = model.compute_controls({'conditioning': cond, 'pedal': pedal, 'piano_model': pm})
    assert np.array_equal(got['extended_pitch'].cpu().numpy().reshape(want['extended_pitch'].shape),
                          want['extended_pitch'])
    for key in ('f0_hz', 'inharm_coef', 'amplitudes', 'harmonic_distribution', 'magnitudes'):
        g, r_ = got[key].cpu().numpy(), want[key]
        err = np.max(np.abs(g - r_)) / np.max(np.abs(r_))
        assert err < 1e-4, (key, err)


@pytest.mark.gpu
def test_dafx22_midi_to_audio(weights):
    """MIDI conditioning -> audio with the shipped weights: one sustained A4 sounds at 440 Hz during
    the note, decays after the note-off, and the clip is silent before the onset."""
    import ddsp_piano_b200 as dp
    B, T, P, sr = 1, 750, 16, 16000
    cond = np.zeros([B, T, P, 2], np.float32)
    cond[0, 250:500, 0, 0] = 69
    cond[0, 250, 0, 1] = 0.8
    model = dp.dafx22_model(weights, device='cuda:0', sample_rate=sr, inference=True)
    out = model({'conditioning': cond, 'pedal': np.zeros([B, T, 4], np.float32),
                 'piano_model': np.zeros([B, 1], np.int64)})
    audio = model.get_audio_from_outputs(out).cpu().numpy()
    U = sr // 250
    assert audio.shape == (B, T * U) and np.all(np.isfinite(audio))
    dry = out['add']['signal'].cpu().numpy()[0]
    before, during, after = dry[100 * U:240 * U], dry[260 * U:500 * U], dry[700 * U:]
    rms = lambda x: float(np.sqrt(np.mean(x.astype(np.float64) ** 2)))
    # (the 16 noise synths hiss at a low level whether or not a note sounds)
    assert rms(during) > 5 * rms(before) and rms(during) > 5 * rms(after)
    n = 8192
    spec = np.abs(np.fft.rfft(during[2048:2048 + n] * np.hanning(n), 4 * n))
    peak_hz = np.argmax(spec) * sr / (4 * n)
    assert abs(peak_hz - 440.0) < 3.0, peak_hz


@pytest.mark.gpu
def test_dafx22_reverb_ir_is_the_unmasked_checkpoint_row(weights):
    """configs/dafx22.gin binds %inference to MultiInharmonic only (dafx22.gin:108-110); its
    MultiInstrumentReverb keeps the constructor default inference=False, so the reference convolves with
    the learnt impulse response as stored -- no exponential decay mask (sub_modules.py:339-365)."""
    import torch
    import ddsp_piano_b200 as dp
    row = weights.tensor('model/reverb_model/reverb_dict/layer_with_weights-0/embeddings/.ATTRIBUTES/VARIABLE_VALUE')
    pm = torch.tensor([[1], [0]], dtype=torch.int64, device='cuda:0')      # the fixture keeps two of the ten rows
    model = dp.dafx22_model(weights, device='cuda:0', inference=True)
    ir = model.reverb_model(pm)
    np.testing.assert_array_equal(ir.cpu().numpy(), row[[1, 0]])
    masked = dp.dafx22_model(weights, device='cuda:0', inference=True, reverb_decay_mask=True).reverb_model(pm)
    np.testing.assert_array_equal(masked[:, :16000].cpu().numpy(), row[[1, 0], :16000])
    assert float((masked[:, 16001:] - ir[:, 16001:]).abs().max()) > 0


# ---- maestro-v2.gin (the script default) --------------------------------------------------------

V2_WEIGHTS = os.path.join(HERE, 'golden', 'v2_weights.npz')


@pytest.fixture(scope='module')
def v2_weights():
    from ddsp_piano_b200.checkpoint import NpzWeights
    return NpzWeights(V2_WEIGHTS)


def test_joint_tuning_known_answers(v2_weights):
    """Rigaud's model: no detuning at the reference pitch, octaves stretched away from it."""
    w = ref.load_weights_v2(v2_weights)['tuning']
    pm = np.array([5])                                  # pitch_ref = 64.0, K = 4.51, alpha = 24 (maestro-v2.gin)
    pitch = np.array([[[64.0], [40.0], [88.0]]], np.float32)
    f0, inharm = ref.joint_inharm_tuning(pitch, pm, w)
    et = 440.0 * 2.0 ** ((pitch[0, :, 0] - 69.0) / 12.0)
    assert abs(f0[0, 0, 0] / et[0] - 1.0) < 1e-6        # at pitch_ref: ratio = 1, detuning = 1
    assert f0[0, 1, 0] < et[1] and f0[0, 2, 0] > et[2]  # Railsback curve: flat bass, sharp treble
    assert np.all(inharm > 0) and inharm[0, 2, 0] > inharm[0, 0, 0] > inharm[0, 1, 0]   # treble bridge dominates


def test_v2_controls_on_cpu_match_the_oracle(v2_weights, monkeypatch):
    from ddsp_piano_b200 import model as M
    cond, pedal, pm = midi_clip(B=2, T=200, seed=5)
    want = ref.control_graph_v2(cond, pedal, pm, ref.load_weights_v2(v2_weights))

    class CpuRelease:
        def __init__(self, dur, fr):
            self.dur, self.fr = dur, fr

        def __call__(self, conditioning):
            return torch.from_numpy(ref.note_release(conditioning[..., 0:1].numpy(), self.dur, self.fr))

    class NoReverb:
        def __init__(self, *a, **kw):
            pass

        def __call__(self, pm_):
            return torch.zeros(pm_.shape[0], 8)

    monkeypatch.setattr(M, 'NoteRelease', CpuRelease)
    monkeypatch.setattr(M, 'MultiInstrumentFeedbackDelayReverb', NoReverb)
    model = M.maestro_v2_model(v2_weights, device='cpu')
    got = model.compute_controls({'conditioning': cond, 'pedal': pedal, 'piano_model': pm})
    for key in ('f0_hz', 'inharm_coef', 'amplitudes', 'harmonic_distribution', 'magnitudes'):
        g, r_ = got[key].numpy(), want[key]
        assert g.shape == r_.shape, (key, g.shape, r_.shape)
        err = np.max(np.abs(g - r_)) / np.max(np.abs(r_))
        assert err < 1e-4, (key, err)
    assert got['f0_hz'].shape[-1] == 1 and got['harmonic_distribution'].shape[-1] == 128


@pytest.mark.gpu
def test_v2_controls_match_the_oracle_on_gpu(v2_weights):
    import ddsp_piano_b200 as dp
    cond, pedal, pm = midi_clip(B=2, T=200, seed=5)
    want = ref.control_graph_v2(cond, pedal, pm, ref.load_weights_v2(v2_weights))
    model = dp.maestro_v2_model(v2_weights, device='cuda:0')
    got = model.compute_controls({'conditioning': cond, 'pedal': pedal, 'piano_model': pm})
    for key in ('f0_hz', 'inharm_coef', 'amplitudes', 'harmonic_distribution', 'magnitudes'):
        g, r_ = got[key].cpu().numpy(), want[key]
        err = np.max(np.abs(g - r_)) / np.max(np.abs(r_))
        assert err < 1e-4, (key, err)
    assert got['reverb_ir'].shape == (2, 48000)


@pytest.mark.gpu
def test_v2_reverb_ir_is_cached_for_frozen_weights(v2_weights):
    """inference=True: the network's response is computed once per set of instrument ids; inference=False
    evaluates it on every call like the reference (sub_modules.py:413-446)."""
    import ddsp_piano_b200 as dp
    cond, pedal, pm = midi_clip(B=2, T=50, seed=6)
    x = {'conditioning': cond, 'pedal': pedal, 'piano_model': pm}
    model = dp.maestro_v2_model(v2_weights, device='cuda:0')
    a = model.compute_controls(x)['reverb_ir']
    b = model.compute_controls(x)['reverb_ir']
    assert b is a and len(model.reverb_model.ir_cache) == 1
    fresh = dp.maestro_v2_model(v2_weights, device='cuda:0', inference=False)
    c = fresh.compute_controls(x)['reverb_ir']
    d = fresh.compute_controls(x)['reverb_ir']
    assert d is not c and torch.equal(c, d) and torch.equal(a, c) and not fresh.reverb_model.ir_cache


@pytest.mark.gpu
def test_v2_midi_to_audio(v2_weights):
    """The reference's default model (maestro-v2.gin, 24 kHz, one string per note, FDN reverb) from
    MIDI conditioning to audio: a sustained A4 peaks at its (slightly stretched) 440 Hz."""
    import ddsp_piano_b200 as dp
    B, T, P, sr = 1, 750, 16, 24000
    cond = np.zeros([B, T, P, 2], np.float32)
    cond[0, 250:500, 0, 0] = 69
    cond[0, 250, 0, 1] = 0.8
    model = dp.maestro_v2_model(v2_weights, device='cuda:0')
    out = model({'conditioning': cond, 'pedal': np.zeros([B, T, 4], np.float32),
                 'piano_model': np.zeros([B, 1], np.int64)})
    audio = model.get_audio_from_outputs(out).cpu().numpy()
    U = sr // 250
    assert audio.shape == (B, T * U) and np.all(np.isfinite(audio))
    dry = out['add']['signal'].cpu().numpy()[0]
    before, during, after = dry[100 * U:240 * U], dry[260 * U:500 * U], dry[700 * U:]
    rms = lambda x: float(np.sqrt(np.mean(x.astype(np.float64) ** 2)))
    assert rms(during) > 5 * rms(before) and rms(during) > 5 * rms(after)
    n = 16384
    spec = np.abs(np.fft.rfft(during[2048:2048 + n] * np.hanning(n), 4 * n))
    peak_hz = np.argmax(spec) * sr / (4 * n)
    assert abs(peak_hz - 440.0) < 4.0, peak_hz
    wet_tail = audio[0, 700 * U:]
    assert rms(wet_tail) > rms(after)                   # the reverb rings on after the dry decay


@pytest.mark.gpu
def test_midi_file_to_audio(weights, tmp_path):
    """The reference's synthesize_midi_file.py path without note_seq / TensorFlow: a MIDI file
    written here (A4 held for 1 s with the sustain pedal down) -> load_midi_as_conditioning ->
    PianoModel (shipped dafx22 weights) -> audio; the spectrum peaks at 440 Hz and the note keeps
    ringing on the pedal after its note-off."""
    import struct
    import ddsp_piano_b200 as dp
    from ddsp_piano_b200 import midi

    def vlq(n):
        out = [n & 0x7f]
        n >>= 7
        while n:
            out.append((n & 0x7f) | 0x80)
            n >>= 7
        return bytes(reversed(out))

    events = [(0, b'\xff\x51\x03' + (500000).to_bytes(3, 'big')), (0, bytes([0xb0, 64, 127])),
              (480, bytes([0x90, 69, 100])), (1440, bytes([0x80, 69, 0])), (2400, bytes([0xb0, 64, 0]))]
    body, last = b'', 0
    for tick, ev in events:
        body += vlq(tick - last) + ev
        last = tick
    body += vlq(0) + b'\xff\x2f\x00'
    path = str(tmp_path / 'a4.mid')
    with open(path, 'wb') as f:
        f.write(b'MThd' + struct.pack('>IHHH', 6, 0, 1, 480) + b'MTrk' + struct.pack('>I', len(body)) + body)
    x = midi.load_midi_as_conditioning(path, n_synths=16, frame_rate=250, warm_up_duration=0.5)
    assert x['conditioning'].shape[1] == int(3 * 250 + 125)
    model = dp.dafx22_model(weights, device='cuda:0', sample_rate=16000, inference=True)
    out = model({'conditioning': x['conditioning'], 'pedal': x['pedal'],
                 'piano_model': np.zeros([1, 1], np.int64)})
    dry = out['add']['signal'].cpu().numpy()[0]
    U = 64
    note = dry[(125 + 130) * U:(125 + 370) * U]          # 0.5 s + note-on at 0.5 s
    n = 8192
    spec = np.abs(np.fft.rfft(note[1024:1024 + n] * np.hanning(n), 4 * n))
    assert abs(np.argmax(spec) * 16000 / (4 * n) - 440.0) < 3.0
    rms = lambda v: float(np.sqrt(np.mean(v.astype(np.float64) ** 2)))
    pedal_tail = dry[(125 + 400) * U:(125 + 600) * U]    # after the note-off (1.5 s), pedal down until 2.5 s
    after = dry[(125 + 700) * U:]
    assert rms(pedal_tail) > 3 * rms(after)


def test_piano_model_id_out_of_range_is_a_value_error(weights, v2_weights):
    """An instrument id beyond the smallest per-instrument table is refused on the host (the reference's
    embedding lookup raises InvalidArgumentError); nothing reaches the device."""
    from ddsp_piano_b200 import model as M
    cond, pedal = np.zeros([1, 4, 16, 2], np.float32), np.zeros([1, 4, 4], np.float32)
    for factory, path, n in ((M.dafx22_model, weights, 2), (M.maestro_v2_model, v2_weights, 10)):
        model = factory(path, device='cpu')
        assert model.n_instruments == n                 # the dafx22 fixture keeps 2 of the 10 shipped IRs
        for bad in ([[n]], [[-1]]):
            with pytest.raises(ValueError, match='piano_model ids'):
                model.compute_controls({'conditioning': cond, 'pedal': pedal, 'piano_model': bad})


@pytest.mark.gpu
@pytest.mark.parametrize('rows,F,units,inputs', [(1, 300, 64, 32), (16, 300, 192, 52), (37, 65, 192, 35),
                                                 (256, 40, 192, 52), (3, 50, 128, 8), (20, 50, 256, 8)])
def test_gru_recurrence_kernel(rows, F, units, inputs):
    """b200ddsp_gru_recurrence (one launch for all frames, clusters sharing the state over distributed
    shared memory) against torch.nn.GRU in float64 on the CPU: same recurrence (Keras reset_after=True),
    tolerance 2e-5 on a state bounded by 1."""
    from ddsp_piano_b200.engine import get_engine
    from ddsp_piano_b200.processors import _DEFAULT_CFG
    g = torch.Generator().manual_seed(rows * 1000 + units)
    gru = torch.nn.GRU(inputs, units, batch_first=True).double()
    with torch.no_grad():
        for p_ in gru.parameters():
            # +-2/sqrt(u): a contracting recurrence (with N(0, 0.3) weights at u >= 192 it is chaotic and
            # torch's own float32 and float64 runs part by O(1) within 300 frames)
            p_.copy_((torch.rand(p_.shape, generator=g, dtype=torch.float64) * 2 - 1) * (2.0 / units ** 0.5))
    x = torch.randn(rows, F, inputs, generator=g, dtype=torch.float64)
    with torch.no_grad():
        want = gru(x)[0].numpy()
        xp = (x @ gru.weight_ih_l0.t() + gru.bias_ih_l0).float()
    eng = get_engine(torch.device('cuda:0'), **_DEFAULT_CFG)
    got = eng.gru_recurrence(xp.cuda(), gru.weight_hh_l0.detach().float().cuda(), gru.bias_hh_l0.detach().float().cuda())
    torch.cuda.synchronize()
    assert got.shape == (rows, F, units)
    err = np.max(np.abs(got.cpu().numpy() - want))
    assert err < 2e-5, err
