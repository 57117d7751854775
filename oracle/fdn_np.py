"""NumPy restatement of the reference's feedback-delay-network reverb IR generator
(``ddsp_piano/modules/fdn_reverb.py:178-360``).

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Pinned against ``tests/golden/fdn_*.npz``,
which were produced by executing the reference's own ``fdn_reverb.py`` over the NumPy
``tensorflow`` stand-in (``tests/golden/make_golden.py::make_fdn``).

The network (fixed configuration of the reference: 8 delay lines, Householder mixing matrix,
one-pole reverberation-time control, 4 Schroeder allpasses per line, FIR early reflections) is
sampled on ``n/2 + 1`` frequencies, ``n = 2 * sampling_rate``:

    H[k] = c^T  D_k (I - F_k D_k)^-1  b          late_ir = irfft(H)          (:321-337)
    D_k  = diag( z^-floor(m_d) * (eta_d + z^-1) / (1 + eta_d z^-1) )         (:241-264)
    F_k  = diag(g_d / (1 - p_d z^-1 + 1e-8)) . (0.5 * 11^T - I) . diag(prod_j allpass_dj)   (:266-312)

and the IR is ``early_ir`` (zero padded) + ``late_ir`` (:339-360).
"""
import numpy as np

DEFAULT_DELAYS = np.array([233, 311, 421, 461, 587, 613, 789, 891], np.float32)   # fdn_reverb.py:96


def fdn_transfer(input_gain, output_gain, gain_allpass, delays_allpass, time_rev_0_sec, alpha_tone,
                 sampling_rate, delay_values=DEFAULT_DELAYS):
    """H[k] for k = 0 .. n/2 (complex64), n = int(2 * sampling_rate)."""
    f32, c64 = np.float32, np.complex64
    sr = f32(sampling_rate)
    n = int(2 * sr)
    nb = n // 2 + 1
    delay_values = np.asarray(delay_values, f32)
    lines = delay_values.shape[0]
    b = np.asarray(input_gain, f32).astype(c64)
    c = np.asarray(output_gain, f32).astype(c64)
    ga = np.asarray(gain_allpass, f32)
    da = np.asarray(delays_allpass, f32)
    t0 = np.asarray(time_rev_0_sec, f32).reshape(-1)[0]
    alpha = np.asarray(alpha_tone, f32).reshape(-1)[0]

    wk = (2 * np.pi * np.arange(nb).astype(f32) / n).astype(c64)            # :233-238
    zinv = np.exp(-1j * wk)                                                  # z^-1 on the unit circle
    # integer delays + first-order allpass interpolation of the fractional part   (:241-262)
    whole = np.floor(delay_values)
    z_d = np.exp(-1j * wk[:, None] * whole.astype(c64)[None, :])
    d_eta = (delay_values - whole).astype(c64)
    eta = (1 - d_eta) / (1 + d_eta)
    interp = (eta[None, :] + zinv[:, None]) / (1 + eta[None, :] * zinv[:, None])
    D = z_d * interp                                                          # [nb, lines] diagonal of D_k

    # one-pole low-pass per line sets the frequency dependent reverberation time   (:266-288)
    delay_sec = (delay_values + np.sum(da, axis=-1, dtype=f32)) / sr
    k = np.power(f32(10.0), -3 * delay_sec / t0)
    kpi = np.power(f32(10.0), -3 * delay_sec / (alpha * t0))
    g = (2 * k * kpi / (k + kpi)).astype(c64)
    p = ((k - kpi) / (k + kpi)).astype(c64)
    lowpass = g[None, :] / (1 - p[None, :] * zinv[:, None] + 1e-8)           # [nb, lines]

    # four Schroeder allpasses in series per line   (:290-308)
    z_del = np.exp(1j * wk[:, None, None] * da.astype(c64)[None, :, :])
    gac = ga.astype(c64)[None, :, :]
    allpass = np.prod((1 + gac * z_del) / (gac + z_del), axis=-1, dtype=c64)  # [nb, lines]

    mixing = (-np.eye(lines, dtype=f32) + 0.5 * np.ones([lines, lines], f32)).astype(c64)   # :118-120
    # feedback matrix F = diag(lowpass) . mixing . diag(allpass)   (:310-312)
    F = (lowpass[:, :, None] * mixing[None, :, :]) @ _diag(allpass)
    M = np.eye(lines, dtype=c64)[None] - F @ _diag(D)                        # :328
    core = _diag(D) @ np.linalg.inv(M)                                       # :326-330
    H = np.squeeze(c[None, None, :] @ (core @ b[None, :, None]))             # :322-334
    return H.astype(c64)


def _diag(x):
    out = np.zeros(x.shape + (x.shape[-1],), dtype=x.dtype)
    i = np.arange(x.shape[-1])
    out[..., i, i] = x
    return out


def fdn_ir(input_gain, output_gain, gain_allpass, delays_allpass, time_rev_0_sec, alpha_tone, early_ir,
           sampling_rate, delay_values=DEFAULT_DELAYS):
    """FeedbackDelayNetwork.get_ir (fdn_reverb.py:339-360): early FIR + late FDN response."""
    H = fdn_transfer(input_gain, output_gain, gain_allpass, delays_allpass, time_rev_0_sec, alpha_tone,
                     sampling_rate, delay_values)
    late = np.fft.irfft(H).astype(np.float32)
    early = np.asarray(early_ir, np.float32).reshape(-1)
    if late.shape[0] > early.shape[0]:
        early = np.pad(early, [(0, late.shape[0] - early.shape[0])])
    return early[:late.shape[0]] + late


def fdn_signal(audio, ir):
    """FeedbackDelayNetwork.get_signal (fdn_reverb.py:406-410): plain convolution, no dry path, no
    masking of ir[0] (unlike ddsp.effects.Reverb)."""
    from . import ddsp_core_np as core
    return core.fft_convolve(np.asarray(audio, np.float32), np.asarray(ir, np.float32)[None, :],
                             padding='same', delay_compensation=0)
