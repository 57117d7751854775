"""CPU tests of the oracle: restatement vs the golden vectors produced by executing the
reference's own Python (tests/golden/make_golden.py), plus the analytical known-answer
tests of SURVEY.md section 8c.  No GPU needed."""
import os

import numpy as np
import pytest

from oracle import ddsp_core_np as core
from oracle import ddsp_piano_np as ref


def load(golden_dir, name):
    return dict(np.load(os.path.join(golden_dir, name + '.npz'), allow_pickle=False))


ADDITIVE = ['additive_24k_inference', 'additive_16k_training', 'additive_48k_h128',
            'additive_16k_exp_tanh_prenorm', 'additive_24k_single_string']


@pytest.mark.parametrize('name', ADDITIVE)
def test_additive_restatement_matches_reference_python(golden_dir, name):
    g = load(golden_dir, name)
    sr = int(g['sample_rate'])
    ctl = ref.additive_controls(
        g['in_amplitudes'], g['in_harmonic_distribution'], g['in_inharm_coef'], g['in_f0_hz'],
        sample_rate=sr, scale_fn=str(g['scale_fn']),
        normalize_after_nyquist_cut=bool(g['normalize_after_nyquist_cut']))
    for k in ('amplitudes', 'harmonic_distribution', 'harmonic_shifts', 'f0_hz'):
        np.testing.assert_array_equal(ctl[k], g['ctl_' + k])      # same ops, same order: bit exact
    sig = ref.additive_signal(**ctl, sample_rate=sr, inference=bool(g['inference']))
    np.testing.assert_array_equal(sig, g['signal'])


@pytest.mark.parametrize('name', ['noise_24k_m64', 'noise_48k_m96', 'noise_16k_m64'])
def test_noise_restatement_matches_reference_python(golden_dir, name):
    g = load(golden_dir, name)
    ctl = ref.noise_controls(g['in_magnitudes'])
    np.testing.assert_array_equal(ctl['magnitudes'], g['ctl_magnitudes'])
    sig = ref.noise_signal(ctl['magnitudes'], g['noise'])
    np.testing.assert_array_equal(sig, g['signal'])


@pytest.mark.parametrize('name', ['reverb_n2400_l1000', 'reverb_n2400_l2400_wet'])
def test_reverb_restatement(golden_dir, name):
    g = load(golden_dir, name)
    sig = ref.reverb_signal(g['audio'], g['ir'], add_dry=bool(g['add_dry']))
    np.testing.assert_array_equal(sig, g['signal'])
    # independent check: float64 direct convolution
    for b in range(g['audio'].shape[0]):
        ir = g['ir'][b].astype(np.float64).copy()
        ir[0] = 0
        wet = np.convolve(g['audio'][b].astype(np.float64), ir)[:g['audio'].shape[1]]
        want = wet + (g['audio'][b] if bool(g['add_dry']) else 0)
        assert np.max(np.abs(sig[b] - want)) <= 2e-6 * max(1.0, np.max(np.abs(want)))


def test_dag_restatement_matches_reference_python(golden_dir):
    g = load(golden_dir, 'dag_24k_p3')
    P, sr = int(g['n_synths']), int(g['sample_rate'])
    assert list(g['node_names']) == ['additive', 'noise', 'add'] * P + ['reverb']
    feats = {k[3:]: v for k, v in g.items() if k.startswith('in_')}
    out = ref.polyphonic_forward(feats, n_synths=P, sample_rate=sr,
                                 noise_by_voice=[g[f'noise_{v}'] for v in range(P)])
    np.testing.assert_array_equal(out['dry'], g['dry'])
    np.testing.assert_array_equal(out['signal'], g['signal'])
    np.testing.assert_array_equal(out['additive'][-1], g['last_additive'])
    np.testing.assert_array_equal(out['noise'][-1], g['last_noise'])
    # voice 2 is silent (f0 below min_frequency): its additive output is exactly zero
    assert not np.any(out['additive'][2])


# ---------------- analytical known-answer tests -----------------

def _const_controls(B, F, H, amp, f0, S=1):
    amps = np.full([B, F, 1], amp, np.float32)
    hd = np.full([B, F, H], 1.0 / H, np.float32)
    shifts = np.zeros([B, F, H], np.float32)
    f0_hz = np.full([B, F, S], f0, np.float32)
    return amps, hd, shifts, f0_hz


def test_kat1_inclusive_cumsum_quarter_rate():
    sr = 16000
    amps, hd, shifts, f0 = _const_controls(1, 4, 1, 1.0, sr / 4)
    y = ref.additive_signal(amps, hd, shifts, f0, sample_rate=sr, inference=False)
    t = np.arange(y.shape[1])
    want = np.cos(np.pi / 2 * (t + 1))                       # first sample is cos(pi/2) = 0
    assert np.max(np.abs(y[0, :64] - want[:64])) < 1e-4      # float32 phase accumulation
    y64 = ref.additive_signal(*[a.astype(np.float64) for a in (amps, hd, shifts, f0)],
                              sample_rate=sr, inference=False)
    assert y64.dtype == np.float64 and np.max(np.abs(y64[0] - want)) < 1e-11


def test_kat2_window_upsampling_is_partition_of_unity():
    x = np.full([1, 10, 3], 0.75, np.float32)
    y = core.upsample_with_windows(x, 10 * 96)
    assert np.max(np.abs(y - 0.75)) < 1e-6
    # closed form y[kU+r] = x[k] w[r+U] + x[k+1] w[r]
    rng = np.random.default_rng(0)
    x = rng.standard_normal([2, 7, 4])
    y = core.upsample_with_windows(x, 7 * 32)
    w = core.hann_window(64, np.float64)
    xe = np.concatenate([x, x[:, -1:]], 1)
    want = (xe[:, :-1, None, :] * w[None, None, 32:, None] +
            xe[:, 1:, None, :] * w[None, None, :32, None]).reshape(2, 7 * 32, 4)
    assert np.max(np.abs(y - want)) < 1e-14


def test_hann_window_of_odd_length_is_symmetric():
    """tf.signal.hann_window(periodic=True) divides by window_length + even - 1: an odd-length window is
    the symmetric one (latent in the configured shapes -- noise IRs of 126 / 190 taps and the 2U resampling
    window are even -- but a cropped IR, window_size < 2 (M - 1) with window_size = 257, would use it)."""
    k = np.arange(257)
    np.testing.assert_allclose(core.hann_window(257, np.float64), 0.5 - 0.5 * np.cos(2 * np.pi * k / 256), atol=1e-15)
    w = core.hann_window(257, np.float32)
    assert abs(float(w[128]) - 1.0) < 1e-6 and abs(float(w[0])) < 1e-7 and abs(float(w[256])) < 1e-6
    k = np.arange(64)
    np.testing.assert_allclose(core.hann_window(64, np.float64), 0.5 - 0.5 * np.cos(2 * np.pi * k / 64), atol=1e-15)


def test_kat3_partials_above_nyquist_are_silent():
    sr = 16000
    B, F, H = 1, 4, 8
    amps = np.zeros([B, F, 1], np.float32)
    hd = np.zeros([B, F, H], np.float32)
    f0 = np.full([B, F, 1], 1500.0, np.float32)          # partials 6.. are >= 8 kHz
    ctl = ref.additive_controls(amps, hd, np.zeros([B, F, 1], np.float32), f0, sample_rate=sr)
    assert np.all(ctl['harmonic_distribution'][..., 5:] == 0)
    assert np.all(ctl['harmonic_distribution'][..., :5] > 0)
    assert np.allclose(ctl['harmonic_distribution'].sum(-1), 1.0, atol=1e-6)


def test_kat4_below_min_frequency_voice_is_exactly_zero():
    rng = np.random.default_rng(1)
    B, F, H = 2, 6, 16
    ctl = ref.additive_controls(rng.standard_normal([B, F, 1]).astype(np.float32),
                                rng.standard_normal([B, F, H]).astype(np.float32),
                                np.full([B, F, 1], 5e-4, np.float32),
                                np.full([B, F, 2], 8.18, np.float32), sample_rate=24000)
    assert not np.any(ctl['amplitudes'])
    y = ref.additive_signal(**ctl, sample_rate=24000)
    assert not np.any(y)


def test_kat5_angular_cumsum_vs_float64():
    rng = np.random.default_rng(2)
    om = rng.uniform(0.0, 3.0, [1, 3500, 4]).astype(np.float32)
    ph32 = core.angular_cumsum(om)
    ph64 = np.mod(np.cumsum(om.astype(np.float64), axis=1), 2 * np.pi)
    d = np.abs(np.angle(np.exp(1j * (ph32 - ph64))))
    assert d.max() < 5e-3
    assert ph32.min() >= 0 and ph32.max() < 2 * np.pi + 1e-6
    # chunk boundary: sample 1000 = (wrapped total of chunk 0) + om[1000]
    c0 = np.mod(np.cumsum(om[0, :1000], axis=0, dtype=np.float32)[-1], np.float32(2 * np.pi))
    np.testing.assert_array_equal(ph32[0, 1000], np.mod(om[0, 1000] + c0, np.float32(2 * np.pi)))


def test_kat6_identical_substrings_equal_single_string():
    rng = np.random.default_rng(3)
    B, F, H = 1, 5, 12
    a = rng.standard_normal([B, F, 1]).astype(np.float32)
    hd = rng.standard_normal([B, F, H]).astype(np.float32)
    ic = np.full([B, F, 1], 3e-4, np.float32)
    f0 = np.full([B, F, 1], 220.0, np.float32)
    c1 = ref.additive_controls(a, hd, ic, f0, sample_rate=16000)
    c2 = ref.additive_controls(a, hd, ic, np.concatenate([f0, f0], -1), sample_rate=16000)
    y1 = ref.additive_signal(**c1, sample_rate=16000)
    y2 = ref.additive_signal(**c2, sample_rate=16000)
    assert np.max(np.abs(y1 - y2)) < 1e-6


def test_kat7_flat_magnitudes_delay_noise_by_two_samples():
    rng = np.random.default_rng(4)
    F, U, M = 12, 96, 64
    noise = rng.uniform(-1, 1, [1, F * U])
    y = ref.noise_signal(np.ones([1, F, M]), noise)          # float64 path
    assert np.max(np.abs(y[0, 2:] - noise[0, :-2])) < 1e-12
    assert np.max(np.abs(y[0, :2])) < 1e-12


def test_kat8_single_frame_support():
    F, U, M = 10, 96, 64
    lir = 2 * (M - 1)
    start = (lir - 1) // 2 - 1
    mags = np.zeros([1, F, M])
    mags[0, 4] = np.linspace(1, 2, M)
    noise = np.ones([1, F * U])
    y = ref.noise_signal(mags, noise)[0]
    nz = np.nonzero(np.abs(y) > 1e-13)[0]
    assert nz.min() >= 4 * U - start and nz.max() < 4 * U - start + U + lir - 1


def test_kat9_fft_path_equals_direct_convolution():
    rng = np.random.default_rng(5)
    F, U, M = 6, 192, 96
    mags = rng.uniform(0, 2, [1, F, M])
    noise = rng.uniform(-1, 1, [1, F * U])
    y = ref.noise_signal(mags, noise)[0]
    ir = core.frequency_impulse_response(mags, 257)[0]
    lir = ir.shape[-1]
    z = np.zeros(F * U + lir - 1)
    for k in range(F):
        z[k * U:k * U + U + lir - 1] += np.convolve(noise[0, k * U:(k + 1) * U], ir[k])
    start = (lir - 1) // 2 - 1
    assert np.max(np.abs(y - z[start:start + F * U])) < 1e-12


def test_kat10_11_reverb_identities():
    rng = np.random.default_rng(6)
    x = rng.standard_normal([2, 500])
    ir = np.zeros([2, 100])
    ir[:, 0], ir[:, 1] = 7.0, 1.0
    y = ref.reverb_signal(x, ir)
    want = x.copy()
    want[:, 1:] += x[:, :-1]
    assert np.max(np.abs(y - want)) < 1e-12                   # ir[0] is masked whatever it is
    ir = np.zeros([2, 100])
    ir[:, 0] = 1.0
    assert np.max(np.abs(ref.reverb_signal(x, ir) - x)) < 1e-12


def test_legacy_bilinear_holds_last_frame_and_matches_float_coordinates():
    F, U = 750, 96
    x = np.arange(F, dtype=np.float32)[None, :, None] * 10
    y = core.resample(x, F * U)[0, :, 0]
    assert np.all(y[-U:] == x[0, -1, 0])
    scale = np.float32(F) / np.float32(F * U)
    assert float(scale) > 1.0 / U                             # SURVEY 7 "hard parts": floor never flips
    t = np.arange(F * U)
    assert np.array_equal(np.floor(t.astype(np.float32) * scale).astype(np.int64), t // U)


def test_exponential_decay_mask():
    ir = np.ones([1, 24000], np.float32)
    m = ref.exponential_decay_mask(ir)
    assert np.all(m[0, :16000] == 1) and abs(m[0, -1] - np.exp(-4)) < 1e-6


def test_reverb_real_ir_fixture(golden_dir):
    """SURVEY 8c KAT 12: the shipped dafx22 impulse response (row 0 of reverb_dict, with the
    inference-time decay mask of sub_modules.py:339-349) against float64 direct convolution."""
    ir = load(golden_dir, 'dafx22_reverb_ir_row0')['ir']
    assert ir.shape == (24000,) and ir.dtype == np.float32
    masked = ref.exponential_decay_mask(ir[None, :])
    assert np.array_equal(masked[0, :16000], ir[:16000])            # decay starts at sample 16000
    assert abs(masked[0, -1] / ir[-1] - np.exp(-4.0)) < 1e-6
    rng = np.random.default_rng(12)
    audio = (rng.standard_normal([1, 8000]) * 0.1).astype(np.float32)
    got = ref.reverb_signal(audio, masked)
    h = masked[0].astype(np.float64).copy()
    h[0] = 0.0
    want = np.convolve(audio[0].astype(np.float64), h)[:8000] + audio[0]
    assert np.max(np.abs(got[0] - want)) <= 2e-6 * np.max(np.abs(want))


@pytest.mark.parametrize('name', ['fdn_sr2000', 'fdn_sr8000'])
def test_fdn_restatement_matches_reference_python(golden_dir, name):
    """the FDN reverb IR generator (modules/fdn_reverb.py) restated in
    oracle/fdn_np.py vs the reference's own code executed over the stand-in."""
    from oracle import fdn_np
    g = load(golden_dir, name)
    keys = ('input_gain', 'output_gain', 'gain_allpass', 'delays_allpass', 'time_rev_0_sec',
            'alpha_tone', 'early_ir')
    ir = fdn_np.fdn_ir(*[g[k] for k in keys], sampling_rate=float(g['sampling_rate']))
    assert ir.dtype == np.float32 and ir.shape == g['ir'].shape
    assert np.max(np.abs(ir - g['ir'])) <= 2e-6 * np.max(np.abs(g['ir']))
    sig = fdn_np.fdn_signal(g['audio'], g['ir'])
    np.testing.assert_array_equal(sig, g['signal'])
    # the early FIR is the head of the response; the late part decays
    assert np.max(np.abs(ir[-200:])) < 0.2 * np.max(np.abs(ir))


def test_fdn_restatement_six_lines(golden_dir):
    """The 6-line network of configs/ENSTDkCl-*.gin:118-122 (trainable delays): restatement vs the
    reference's code executed over the stand-in (tests/golden/make_golden.py::make_fdn6)."""
    from oracle import fdn_np
    g = load(golden_dir, 'fdn6_sr4000')
    keys = ('input_gain', 'output_gain', 'gain_allpass', 'delays_allpass', 'time_rev_0_sec',
            'alpha_tone', 'early_ir')
    assert g['input_gain'].shape == (6,) and g['delay_values'].shape == (6,)
    ir = fdn_np.fdn_ir(*[g[k] for k in keys], sampling_rate=float(g['sampling_rate']),
                       delay_values=g['delay_values'])
    assert np.max(np.abs(ir - g['ir'])) <= 2e-6 * np.max(np.abs(g['ir']))
    np.testing.assert_array_equal(fdn_np.fdn_signal(g['audio'], g['ir']), g['signal'])


@pytest.mark.parametrize('name', ['surrogate_16k', 'surrogate_24k_h40'])
def test_surrogate_restatement_matches_reference_execution(golden_dir, name):
    """SurrogateAdditive (surrogate_synth.py), goldens from tests/golden/make_golden_surrogate.py:
    the restatement reproduces the reference-executed controls and signal bit for bit."""
    g = load(golden_dir, name)
    sr = int(g['sample_rate'])
    ctl = ref.surrogate_controls(g['in_amplitudes'], g['in_decays'], g['in_decay_time'],
                                 g['in_harmonic_distribution'], g['in_inharm_coef'], g['in_f0_hz'],
                                 sample_rate=sr)
    for k, v in ctl.items():
        np.testing.assert_array_equal(np.asarray(v, np.float32), g['ctl_' + k], err_msg=k)
    assert np.all((g['ctl_decays'] >= 1e-5) & (g['ctl_decays'] <= 1.0))
    sig = ref.surrogate_signal(**ctl, sample_rate=sr, inference=True)
    np.testing.assert_array_equal(sig, g['signal'])


def test_additive_segments_with_carried_state_equal_the_whole_clip():
    """SURVEY 8e-i/ii as a specification (oracle only; no kernel does this yet): frames [f0, f1) of a long
    clip synthesised from one frame of halo on each side and the carried float32 chunk-offset sum are BIT
    identical to the same samples of the whole-clip synthesis; restarting the phase per segment is not."""
    rng = np.random.default_rng(3)
    sr, B, F, H, S = 24000, 2, 375, 24, 2                 # three segments of 125 frames = 12 000 samples
    f0 = (110.0 * 2 ** rng.uniform(0, 3, [B, 1, 1]) * (1 + 1e-3 * np.arange(S))[None, None, :])
    f0 = np.broadcast_to(f0, [B, F, S]).astype(np.float32).copy()
    f0[:, 200:, :] *= np.float32(1.122)                   # a pitch change inside the second segment
    ctl = ref.additive_controls(rng.standard_normal([B, F, 1]).astype(np.float32),
                                rng.standard_normal([B, F, H]).astype(np.float32),
                                rng.uniform(1e-4, 1e-3, [B, F, 1]).astype(np.float32), f0, sample_rate=sr)
    whole = ref.additive_signal(**ctl, sample_rate=sr, inference=True)
    carry, parts = None, []
    for lo in (0, 125, 250):
        y, carry = ref.additive_signal_segment(**ctl, frames=(lo, lo + 125), carry=carry, sample_rate=sr)
        parts.append(y)
        assert carry.shape == (S, B, H) and carry.dtype == np.float32
    assert np.array_equal(np.concatenate(parts, axis=1), whole)
    restarted, _ = ref.additive_signal_segment(**ctl, frames=(125, 250), carry=None, sample_rate=sr)
    assert not np.array_equal(restarted, whole[:, 12000:24000])
    with pytest.raises(ValueError):
        ref.additive_signal_segment(**ctl, frames=(1, 126), carry=None, sample_rate=sr)


@pytest.mark.parametrize('M', [64, 96])
def test_noise_segments_with_frame_halo_equal_the_whole_clip(M):
    """SURVEY 8e-iii as a specification (oracle only): a segment of the filtered noise from its frames plus
    ceil(fft / U) - 1 frames of halo before and one after is BIT identical to the whole-clip synthesis."""
    rng = np.random.default_rng(5)
    sr, B, F, U = 24000, 2, 60, 96
    mags = ref.noise_controls(rng.standard_normal([B, F, M]).astype(np.float32) * 2 + 3)['magnitudes']
    noise = rng.uniform(-1, 1, [B, F * U]).astype(np.float32)
    whole = ref.noise_signal(mags, noise)
    parts = [ref.noise_signal_segment(mags, noise, frames=(lo, hi), sample_rate=sr)
             for lo, hi in ((0, 20), (20, 27), (27, 60))]
    assert np.array_equal(np.concatenate(parts, axis=1), whole)
    # without the halo (the segment treated as its own clip, as the kernels do today) the head differs
    alone = ref.noise_signal(mags[:, 20:27], noise[:, 20 * U:27 * U])
    assert not np.array_equal(alone, whole[:, 20 * U:27 * U])
