"""NumPy restatement of the reference's synthesis processors (the hot path).

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Each function cites the
reference lines it follows (paths relative to ``/root/reference/ddsp_piano``).
The dtype of the inputs selects the precision: float32 is the TF-faithful
restatement, float64 the ground-truth variant.

Pinned against ``tests/golden/*.npz`` -- vectors produced by executing the
reference's own ``modules/inharm_synth.py`` / ``filtered_noise_synth.py`` /
``polyphonic_dag.py`` over the NumPy ``tensorflow``/``ddsp`` stand-in of
``oracle/tf_shim`` (generator: ``tests/golden/make_golden.py``).  The ddsp/TF
layer underneath is restated, not imported: parity for it is unpinned.
"""
import numpy as np

from . import ddsp_core_np as core

SCALE_EXP_SIGMOID = 'exp_sigmoid'
SCALE_EXP_TANH = 'exp_tanh'
SCALE_NONE = None


def exp_tanh(x, max_value=2., exponent=10., gain=1., threshold=1e-7):
    """modules/inharm_synth.py:8-17 -- tanh flavoured exp_sigmoid."""
    x = core.tf_float32(x)
    dt = x.dtype.type
    pos = dt(0.5) * (np.tanh(dt(gain) * x) + dt(1.))
    return dt(max_value) * pos ** dt(np.log(exponent)) + dt(threshold)


def _scale(name):
    if name in (SCALE_EXP_SIGMOID, 'core.exp_sigmoid'):
        return core.exp_sigmoid
    if name == SCALE_EXP_TANH:
        return exp_tanh
    if name is None:
        return None
    if callable(name):
        return name
    raise ValueError(f'unknown scale_fn {name!r}')


def inharmonic_frequencies(f0_hz, inharm_coef, n_harmonics):
    """modules/inharm_synth.py:20-46: partial k at f0*k*sqrt(1 + B*k^2)."""
    dt = f0_hz.dtype.type
    k = np.linspace(1.0, float(n_harmonics), int(n_harmonics)).astype(dt)[None, None, :]
    stretch = np.sqrt(np.power(k, dt(2)) * inharm_coef + dt(1.))
    return f0_hz * k * stretch, stretch - dt(1.)


def oscillator_bank(freq_env, amp_env, sample_rate, angular):
    """modules/inharm_synth.py:49-84 (cos_oscillator_bank), sum_sinusoids=True."""
    dt = freq_env.dtype.type
    amp_env = core.remove_above_nyquist(freq_env, amp_env, sample_rate)     # :65-67
    omega = freq_env * dt(2.0 * np.pi)                                        # :69
    omega = omega / dt(float(sample_rate))                                    # :70
    if angular:
        phase = core.angular_cumsum(omega)                                    # :75
    else:
        phase = np.cumsum(omega, axis=1, dtype=dt)                            # :77
    return np.sum(amp_env * np.cos(phase), axis=-1, dtype=dt)                 # :80-83


def additive_controls(amplitudes, harmonic_distribution, inharm_coef, f0_hz, *,
                      sample_rate, min_frequency=20, scale_fn=SCALE_EXP_SIGMOID,
                      normalize_after_nyquist_cut=True, normalize_below_nyquist=True):
    """MultiInharmonic.get_controls, modules/inharm_synth.py:254-270 on top of
    InHarmonic.get_controls :167-219.  f0_hz is [B, F, S]."""
    amplitudes = core.tf_float32(amplitudes)
    harmonic_distribution = core.tf_float32(harmonic_distribution)
    inharm_coef = core.tf_float32(inharm_coef)
    f0_hz = core.tf_float32(f0_hz)
    dt = f0_hz.dtype.type
    f0_first = f0_hz[..., 0:1]                                                # :262
    fn = _scale(scale_fn)
    inharm_coef = np.maximum(inharm_coef, dt(0.))                             # :183
    if fn is not None:                                                        # :184-186
        amplitudes = fn(amplitudes)
        harmonic_distribution = fn(harmonic_distribution)
    n_harm = int(harmonic_distribution.shape[-1])
    partial_hz, shifts = inharmonic_frequencies(f0_first, inharm_coef, n_harm)
    if not normalize_after_nyquist_cut:                                       # :194-198
        harmonic_distribution = core.safe_divide(
            harmonic_distribution, np.sum(harmonic_distribution, -1, keepdims=True))
    if normalize_below_nyquist:                                               # :200-208
        harmonic_distribution = core.remove_above_nyquist(
            partial_hz, harmonic_distribution, sample_rate)
        amplitudes = amplitudes * (f0_first > dt(min_frequency)).astype(dt)
    if normalize_after_nyquist_cut:                                           # :210-214
        harmonic_distribution = core.safe_divide(
            harmonic_distribution, np.sum(harmonic_distribution, -1, keepdims=True))
    amplitudes = amplitudes / dt(f0_hz.shape[-1])                             # :269
    return {'amplitudes': amplitudes,
            'harmonic_distribution': harmonic_distribution,
            'harmonic_shifts': shifts,
            'f0_hz': f0_hz}


def additive_signal(amplitudes, harmonic_distribution, harmonic_shifts, f0_hz, *,
                    sample_rate, frame_rate=250, inference=True):
    """MultiInharmonic.get_signal, modules/inharm_synth.py:272-293: one full
    harmonic_synthesis (:87-127) per substring, summed."""
    amplitudes = core.tf_float32(amplitudes)
    harmonic_distribution = core.tf_float32(harmonic_distribution)
    harmonic_shifts = core.tf_float32(harmonic_shifts)
    f0_hz = core.tf_float32(f0_hz)
    dt = f0_hz.dtype.type
    upsampling = int(sample_rate / frame_rate)                                # :163-165
    n_samples = upsampling * f0_hz.shape[1]                                   # :240
    n_harm = harmonic_distribution.shape[-1]
    audio = None
    for s in range(f0_hz.shape[-1]):
        partial_hz = core.get_harmonic_frequencies(f0_hz[..., s:s + 1], n_harm)   # :106
        partial_hz = partial_hz * (dt(1.0) + harmonic_shifts)                 # :107-108
        partial_amp = amplitudes * harmonic_distribution                      # :111-114
        freq_env = core.resample(partial_hz, n_samples)                       # :117
        amp_env = core.resample(partial_amp, n_samples, method='window')      # :118-119
        y = oscillator_bank(freq_env, amp_env, sample_rate, inference)        # :122-126
        audio = y if audio is None else audio + y                             # :286-292
    return audio


def additive_signal_exact_sum(amplitudes, harmonic_distribution, harmonic_shifts, f0_hz, *,
                              sample_rate, frame_rate=250, rounded_omega=False):
    """The reference's signal model in EXACT arithmetic -- ground truth for ``fast_phase`` (SURVEY 7 'hard
    parts': a fast mode validated against a float64 oracle).  Envelopes come from the float32 controls along
    the reference's own resampling (legacy bilinear coordinates float32(i) * float32(F / N), window
    cross-fade); then the angular frequency (F_k + g_k lerp) 2 pi / sr, its running sum, the cosines and the
    sum over partials are evaluated in float64.  What the reference adds on top of this model is float32
    rounding: of every omega (systematic for a held partial: the same rounded value is added thousands of
    times) and of the running sum itself (angular_cumsum); ``rounded_omega=True`` keeps the first kind
    (the float32 omegas of modules/inharm_synth.py:69-70, summed exactly) to tell the two apart.
    additive_signal on float64 INPUTS is a third thing: it also resamples along float64 coordinates."""
    amplitudes = core.tf_float32(amplitudes).astype(np.float32)
    harmonic_distribution = core.tf_float32(harmonic_distribution).astype(np.float32)
    harmonic_shifts = core.tf_float32(harmonic_shifts).astype(np.float32)
    f0_hz = core.tf_float32(f0_hz).astype(np.float32)
    dt = np.float32
    n_frames = f0_hz.shape[1]
    upsampling = int(sample_rate / frame_rate)
    n_samples = upsampling * n_frames
    n_harm = harmonic_distribution.shape[-1]
    # legacy bilinear coordinates (core._resize_bilinear_legacy), float32 like the reference
    pos = np.arange(n_samples).astype(dt) * (dt(n_frames) / dt(n_samples))
    lower = np.floor(pos).astype(np.int64)
    upper = np.minimum(np.ceil(pos).astype(np.int64), n_frames - 1)
    lerp = (pos - np.floor(pos)).astype(dt)
    audio = None
    for s in range(f0_hz.shape[-1]):
        partial_hz = core.get_harmonic_frequencies(f0_hz[..., s:s + 1], n_harm) * (dt(1.0) + harmonic_shifts)
        partial_amp = amplitudes * harmonic_distribution
        freq_env = core.resample(partial_hz, n_samples)                               # float32: the Nyquist mask
        amp_env = core.resample(partial_amp, n_samples, method='window')
        amp_env = core.remove_above_nyquist(freq_env, amp_env, sample_rate)
        if rounded_omega:
            omega = ((freq_env * dt(2.0 * np.pi)) / dt(float(sample_rate))).astype(np.float64)
        else:
            top, bottom = partial_hz[:, lower, :], partial_hz[:, upper, :]
            slope = (bottom - top).astype(np.float64)                                  # float32 difference, like the resize
            omega = (top.astype(np.float64) + slope * lerp[None, :, None].astype(np.float64)) * (2.0 * np.pi / sample_rate)
        phase = np.mod(np.cumsum(omega, axis=1), 2.0 * np.pi)
        y = np.sum(amp_env.astype(np.float64) * np.cos(phase), axis=-1)
        audio = y if audio is None else audio + y
    return audio


def surrogate_controls(amplitudes, decays, decay_time, harmonic_distribution, inharm_coef, f0_hz, *,
                       sample_rate, min_frequency=20, scale_fn=SCALE_EXP_SIGMOID,
                       normalize_harm_distribution=True, normalize_below_nyquist=True):
    """SurrogateAdditive.get_controls, modules/surrogate_synth.py:132-189.  f0_hz is [B, F, 1]."""
    amplitudes = core.tf_float32(amplitudes)
    harmonic_distribution = core.tf_float32(harmonic_distribution)
    inharm_coef = core.tf_float32(inharm_coef)
    f0_hz = core.tf_float32(f0_hz)
    dt = f0_hz.dtype.type
    fn = _scale(scale_fn)
    if fn is not None:                                                        # :152-154
        amplitudes = fn(amplitudes)
        harmonic_distribution = fn(harmonic_distribution)
    inharm_coef = np.maximum(inharm_coef, dt(0.))                             # :157
    n_harm = int(harmonic_distribution.shape[-1])
    partial_hz, shifts = inharmonic_frequencies(f0_hz, inharm_coef, n_harm)   # :159-161
    if decays is not None:                                                    # :163-171
        decays = np.maximum(np.minimum(core.tf_float32(decays), dt(1.)), dt(1e-5))
        decays = np.where(partial_hz >= dt(sample_rate / 2.), np.ones_like(decays), decays)
    if normalize_below_nyquist:                                               # :172-181
        harmonic_distribution = core.remove_above_nyquist(partial_hz, harmonic_distribution, sample_rate)
        amplitudes = amplitudes * (f0_hz > dt(min_frequency)).astype(dt)
    if normalize_harm_distribution:                                           # :183-187
        harmonic_distribution = core.safe_divide(
            harmonic_distribution, np.sum(harmonic_distribution, -1, keepdims=True))
    return {'amplitudes': amplitudes, 'decays': decays, 'decay_time': decay_time,
            'harmonic_distribution': harmonic_distribution, 'harmonic_shifts': shifts, 'f0_hz': f0_hz}


def surrogate_signal(amplitudes, decays, decay_time, harmonic_distribution, harmonic_shifts, f0_hz, *,
                     sample_rate, frame_rate=250, inference=True):
    """SurrogateAdditive.get_signal -> surrogate_harmonic_synthesis, modules/surrogate_synth.py:11-104:
    the additive bank with every partial's amplitude envelope multiplied by |decay|^t, t = samples
    since the (frame-rate) decay_time origin, both held over each control frame (tf.repeat)."""
    amplitudes = core.tf_float32(amplitudes)
    harmonic_distribution = core.tf_float32(harmonic_distribution)
    harmonic_shifts = core.tf_float32(harmonic_shifts)
    f0_hz = core.tf_float32(f0_hz)
    dt = f0_hz.dtype.type
    upsampling = int(sample_rate / frame_rate)
    n_frames = f0_hz.shape[1]
    n_samples = upsampling * n_frames                                         # :49
    n_harm = harmonic_distribution.shape[-1]
    partial_hz = core.get_harmonic_frequencies(f0_hz, n_harm) * (dt(1.0) + harmonic_shifts)   # :61-64
    partial_amp = amplitudes * harmonic_distribution                          # :67-70
    freq_env = core.resample(partial_hz, n_samples)                           # :73
    amp_env = core.resample(partial_amp, n_samples, method='window')          # :74-75
    if decays is not None and decay_time is not None:                         # :78-97
        decay_env = np.repeat(core.tf_float32(decays), upsampling, axis=1)
        t = np.repeat(core.tf_float32(decay_time), upsampling, axis=1) * dt(upsampling)
        t = t + np.tile(np.arange(upsampling, dtype=dt), n_frames)[None, :, None]
        amp_env = amp_env * np.power(np.abs(decay_env), t.astype(dt))
    return oscillator_bank(freq_env, amp_env, sample_rate, inference)         # :100-103


def noise_controls(magnitudes, *, initial_bias=-5.0, scale_fn=SCALE_EXP_SIGMOID):
    """ddsp.synths.FilteredNoise.get_controls (inherited by
    modules/filtered_noise_synth.py:13): scale_fn(magnitudes + initial_bias)."""
    magnitudes = core.tf_float32(magnitudes)
    fn = _scale(scale_fn)
    if fn is not None:
        magnitudes = fn(magnitudes + magnitudes.dtype.type(initial_bias))
    return {'magnitudes': magnitudes}


def noise_signal(magnitudes, noise, *, window_size=257):
    """DynamicSizeFilteredNoise.get_signal, modules/filtered_noise_synth.py:27-42.
    The reference draws ``tf.random.uniform([B, N], -1, 1)`` unseeded (:39-40);
    the oracle takes that tensor as an argument so parity is definable."""
    magnitudes = core.tf_float32(magnitudes)
    noise = core.tf_float32(noise)
    return core.frequency_filter(noise, magnitudes, window_size=window_size)    # :41-42


def reverb_signal(audio, ir, *, add_dry=True):
    """ddsp.effects.Reverb.get_signal (configured configs/dafx22.gin:99-100,111):
    zero ir[:, 0], one-block FFT convolution with delay_compensation=0, plus dry."""
    audio, ir = core.tf_float32(audio), core.tf_float32(ir)
    if ir.ndim == 1:
        ir = ir[None, :]
    if ir.ndim == 3:
        ir = ir[:, :, 0]
    ir = np.concatenate([np.zeros([ir.shape[0], 1], ir.dtype), ir[:, 1:]], axis=1)
    wet = core.fft_convolve(audio, ir, padding='same', delay_compensation=0)
    return (wet + audio) if add_dry else wet


def exponential_decay_mask(ir, decay_exponent=4., decay_start=16000):
    """modules/sub_modules.py:339-349 (MultiInstrumentReverb, inference only)."""
    ir = core.tf_float32(ir)
    dt = ir.dtype.type
    length = ir.shape[-1]
    time = np.linspace(0.0, 1.0, length - decay_start).astype(dt)
    mask = np.concatenate([np.ones(decay_start, dt), np.exp(-dt(decay_exponent) * time)])
    return ir * mask[None, :]


def polyphonic_forward(features, *, n_synths, sample_rate, frame_rate=250,
                       inference=True, noise_by_voice=None, add_dry=True,
                       scale_fn=SCALE_EXP_SIGMOID, noise_scale_fn=SCALE_EXP_SIGMOID,
                       normalize_after_nyquist_cut=True, normalize_below_nyquist=True,
                       window_size=257, initial_bias=-5.0, reverb=True):
    """The DAG of modules/polyphonic_dag.py:21-42 as wired by configs/dafx22.gin:91-100:
    per voice additive + noise, running sum 'add' (MultiAdd, inharm_synth.py:296-309,
    python sum() = ((0 + a) + b) + c), then reverb on the sum.

    features: dict with amplitudes_i, harmonic_distribution_i, inharm_coef_i, f0_hz_i,
    magnitudes_i (i < n_synths) and reverb_ir.  noise_by_voice: list of [B, N] arrays.
    Returns dict(dry=[B, N], signal=[B, N], additive=[...per voice], noise=[...])."""
    dry = None
    add_sigs, noise_sigs = [], []
    for v in range(n_synths):
        c = additive_controls(features[f'amplitudes_{v}'],
                              features[f'harmonic_distribution_{v}'],
                              features[f'inharm_coef_{v}'], features[f'f0_hz_{v}'],
                              sample_rate=sample_rate, scale_fn=scale_fn,
                              normalize_after_nyquist_cut=normalize_after_nyquist_cut,
                              normalize_below_nyquist=normalize_below_nyquist)
        a = additive_signal(**c, sample_rate=sample_rate, frame_rate=frame_rate,
                            inference=inference)
        m = noise_controls(features[f'magnitudes_{v}'], initial_bias=initial_bias,
                           scale_fn=noise_scale_fn)
        n = noise_signal(m['magnitudes'], noise_by_voice[v], window_size=window_size)
        add_sigs.append(a)
        noise_sigs.append(n)
        dry = (n + a) if dry is None else (dry + n) + a        # polyphonic_dag.py:28-37
    out = dry
    if reverb:
        out = reverb_signal(dry, features['reverb_ir'], add_dry=add_dry)
    return {'dry': dry, 'signal': out, 'additive': add_sigs, 'noise': noise_sigs}


# ---------------------------------------------------------------------------------------------
# Timeline segments: the additive synth of frames [f0, f1) of a long clip, bit for
# bit what additive_signal gives for those samples on the whole clip.  This is the SPECIFICATION of
# the carried state for a segment-sharded oscillator bank (no kernel implements it yet, DESIGN.md
# section 6): what has to cross a segment boundary is
#   * one frame of halo on each side for the two resamplers: the window resampler reads frame k + 1,
#     and the legacy bilinear coordinate float32(i) * float32(F / N) of the GLOBAL sample index i can
#     round below the frame of sample i, so the first samples of a segment may read frame f0 - 1 (it
#     does not at U = 64, 96, 192, where float32(1 / U) is exact or rounds up; the assert keeps watch);
#   * per (clip, substring, partial) the float32 running sum of the chunk-end phases mod 2 pi that
#     angular_cumsum (ddsp.core, chunk_size 1000) accumulates UNWRAPPED over the chunks before the
#     segment, so f0 * U must be a multiple of 1000 (72 000 samples per 3 s segment at 24 kHz is).
# ---------------------------------------------------------------------------------------------
def additive_signal_segment(amplitudes, harmonic_distribution, harmonic_shifts, f0_hz, *, frames,
                            carry=None, sample_rate, frame_rate=250, chunk_size=1000):
    """Controls of the WHOLE clip ([B, F, ...], the output of additive_controls) + frames = (f0, f1)
    + carry [S, B, H] from the previous segment (None at f0 = 0) -> (audio [B, (f1 - f0) * U],
    carry for the next segment).  Only frames f0 - 1 .. f1 of the controls are read.  inference=True
    (angular_cumsum) only: the plain cumsum of training mode carries its unbounded float32 phase."""
    amplitudes = core.tf_float32(amplitudes)
    harmonic_distribution = core.tf_float32(harmonic_distribution)
    harmonic_shifts = core.tf_float32(harmonic_shifts)
    f0_hz = core.tf_float32(f0_hz)
    dt = f0_hz.dtype.type
    two_pi = dt(2.0 * np.pi)
    f_lo, f_hi = frames
    n_frames = f0_hz.shape[1]
    U = int(sample_rate / frame_rate)
    n_total = U * n_frames
    i_lo, i_hi = f_lo * U, f_hi * U
    if i_lo % chunk_size or (i_hi % chunk_size and f_hi != n_frames):
        raise ValueError('segment boundaries must fall on angular_cumsum chunk boundaries')
    n_harm = harmonic_distribution.shape[-1]
    S = f0_hz.shape[-1]
    # legacy bilinear coordinates of the global sample indices (core._resize_bilinear_legacy)
    scale = dt(n_frames) / dt(n_total)
    pos = np.arange(i_lo, i_hi).astype(dt) * scale
    pos_floor = np.floor(pos)
    lower = np.maximum(pos_floor.astype(np.int64), 0)
    upper = np.minimum(np.ceil(pos).astype(np.int64), n_frames - 1)
    lerp = (pos - pos_floor).astype(dt)
    assert lower.min() >= max(f_lo - 1, 0) and upper.max() <= min(f_hi, n_frames - 1)      # the halo
    # window resampler in closed form (core.upsample_with_windows): y[kU + r] = x[k] w[r + U] + x[k + 1] w[r]
    window = core.hann_window(2 * U, dt)
    k = np.arange(i_lo, i_hi) // U
    r = np.arange(i_lo, i_hi) % U
    k_next = np.minimum(k + 1, n_frames - 1)                                  # add_endpoint: last frame held
    partial_amp = amplitudes * harmonic_distribution                          # :111-114
    amp_env = partial_amp[:, k_next, :] * window[r][None, :, None] + partial_amp[:, k, :] * window[r + U][None, :, None]
    audio, carry_out = None, []
    for s in range(S):
        partial_hz = core.get_harmonic_frequencies(f0_hz[..., s:s + 1], n_harm) * (dt(1.0) + harmonic_shifts)
        top, bottom = partial_hz[:, lower, :], partial_hz[:, upper, :]
        freq_env = top + (bottom - top) * lerp[None, :, None]                 # :117
        amp = core.remove_above_nyquist(freq_env, amp_env, sample_rate)       # :65-67
        omega = freq_env * two_pi / dt(float(sample_rate))                    # :69-70
        B, n = omega.shape[:2]
        pad = (-n) % chunk_size
        if pad:
            omega = np.pad(omega, [(0, 0), (0, pad), (0, 0)])
        chunks = omega.reshape(B, -1, chunk_size, n_harm)
        phase = np.cumsum(chunks, axis=2, dtype=dt)
        ends = np.mod(phase[:, :, -1, :], two_pi)                             # [B, chunks, H]
        start = np.zeros([B, 1, n_harm], dt) if carry is None else np.asarray(carry[s], dt)[:, None, :]
        running = np.cumsum(np.concatenate([start, ends], axis=1), axis=1, dtype=dt)      # unwrapped, sequential
        phase = np.mod(phase + np.mod(running[:, :-1], two_pi)[:, :, None, :], two_pi)
        phase = phase.reshape(B, -1, n_harm)[:, :n]
        carry_out.append(running[:, -1])
        y = np.sum(amp * np.cos(phase), axis=-1, dtype=dt)                    # :80-83
        audio = y if audio is None else audio + y
    return audio, np.stack(carry_out)


def noise_signal_segment(magnitudes, noise, *, frames, sample_rate, frame_rate=250, window_size=257):
    """SURVEY 8e-iii as a specification: samples [f0 * U, f1 * U) of noise_signal on the whole clip,
    bit for bit, from the frames that reach into them.  Frame f of the noise is filtered by its own
    impulse response and laid down at f * U with a tail of fft_size - U samples, and the output is read
    (Lir - 1) // 2 - 1 samples late (ddsp.core.fft_convolve, delay_compensation = -1), so a segment
    needs ceil(fft_size / U) - 1 frames of halo before it (2 frames for M = 64, 5 for M = 96 at 24 kHz:
    the zero-padded end of each transform is rounding noise, not zeros, and the whole-clip sum includes
    it) and one frame after it; the per-sample sums run from the newest frame to the oldest, as
    tf.signal.overlap_and_add (restated in core.fft_convolve) orders them.
    magnitudes [B, F, M] (after noise_controls) and noise [B, N] are those of the WHOLE clip; only the
    halo-extended range is read."""
    magnitudes = core.tf_float32(magnitudes)
    noise = core.tf_float32(noise)
    dt = noise.dtype.type
    B, F = magnitudes.shape[:2]
    U = int(sample_rate / frame_rate)
    f_lo, f_hi = frames
    ir_size = 2 * (magnitudes.shape[-1] - 1)
    if 1 <= window_size < ir_size:                              # a shorter window crops the response
        ir_size = window_size
    fft_size = core.get_fft_size(U, ir_size, power_of_2=True)
    n_seg = -(-fft_size // U)
    start = (ir_size - 1) // 2 - 1
    fa, fb = max(0, f_lo - (n_seg - 1)), min(F, f_hi + 1)
    ir = core.frequency_impulse_response(magnitudes[:, fa:fb], window_size=window_size)
    assert ir.shape[-1] == ir_size
    audio_frames = noise[:, fa * U:fb * U].reshape(B, fb - fa, U)
    fr = np.fft.irfft(np.fft.rfft(audio_frames, fft_size) * np.fft.rfft(ir, fft_size), fft_size).astype(dt)
    fr = np.pad(fr, [(0, 0), (0, 0), (0, n_seg * U - fft_size)]).reshape(B, fb - fa, n_seg, U)
    i_lo, i_hi = f_lo * U + start, f_hi * U + start            # positions in the un-cropped overlap-add
    j_lo, j_hi = i_lo // U, (i_hi - 1) // U
    blocks = np.zeros([B, j_hi - j_lo + 1, U], dt)
    for s in range(n_seg):                                      # newest frame first
        for j in range(j_lo, j_hi + 1):
            f = j - s
            if fa <= f < fb:
                blocks[:, j - j_lo] += fr[:, f - fa, s]
    flat = blocks.reshape(B, -1)
    return flat[:, i_lo - j_lo * U:i_hi - j_lo * U]
