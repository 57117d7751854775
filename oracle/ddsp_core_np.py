"""NumPy restatement of the ``ddsp.core`` primitives the DDSP-Piano hot path calls.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  PARITY UNPINNED for this file:
``ddsp`` (pip, "tested on v3.2.0 and v3.7.0", reference ``README.md:12-14``) and
TensorFlow are third-party packages that are neither vendored under
``/root/reference`` nor installable in the build container, so every function
below restates the published ddsp v3.7.0 algorithm (``ddsp/core.py``) and the
TensorFlow CPU kernel semantics it lowers to.  Reference call sites are cited on
each function (paths relative to ``/root/reference/ddsp_piano``).

Two precisions from one code path: every function computes in the dtype of its
input array.  ``float32`` input reproduces TensorFlow's float32 evaluation order
(sequential in-chunk cumsum, ``floormod``, legacy bilinear coordinates computed
as ``float(i) * scale``, unfused multiply/add); ``float64`` input gives the
ground-truth variant of the same algorithm.
"""
import numpy as np


def _dt(x):
    return np.asarray(x).dtype.type


def tf_float32(x):
    """ddsp.core.tf_float32 -- cast to float32 (kept as float64 in f64 mode)."""
    x = np.asarray(x)
    if x.dtype == np.float64:
        return x
    return x.astype(np.float32)


def sigmoid(x):
    dt = _dt(x)
    return dt(1.0) / (dt(1.0) + np.exp(-x))


def exp_sigmoid(x, exponent=10.0, max_value=2.0, threshold=1e-7):
    """ddsp.core.exp_sigmoid: max_value * sigmoid(x)**log(exponent) + threshold.

    Default ``scale_fn`` of InHarmonic (modules/inharm_synth.py:149) and of
    ddsp.synths.FilteredNoise (base of modules/filtered_noise_synth.py:13).
    """
    x = tf_float32(x)
    dt = _dt(x)
    return dt(max_value) * sigmoid(x) ** dt(np.log(exponent)) + dt(threshold)


def safe_divide(numerator, denominator, eps=1e-7):
    """ddsp.core.safe_divide (call sites modules/inharm_synth.py:195,211)."""
    dt = _dt(denominator)
    safe_denominator = np.where(denominator == 0.0, dt(eps), denominator)
    return numerator / safe_denominator


def remove_above_nyquist(frequency_envelopes, amplitude_envelopes, sample_rate=16000):
    """ddsp.core.remove_above_nyquist (call sites modules/inharm_synth.py:65,201)."""
    frequency_envelopes = tf_float32(frequency_envelopes)
    amplitude_envelopes = tf_float32(amplitude_envelopes)
    dt = _dt(frequency_envelopes)
    return np.where(frequency_envelopes >= dt(sample_rate / 2.0),
                    np.zeros_like(amplitude_envelopes), amplitude_envelopes)


def get_harmonic_frequencies(frequencies, n_harmonics):
    """ddsp.core.get_harmonic_frequencies (call site modules/inharm_synth.py:106)."""
    frequencies = tf_float32(frequencies)
    dt = _dt(frequencies)
    f_ratios = np.linspace(1.0, float(n_harmonics), int(n_harmonics)).astype(dt)
    return frequencies * f_ratios[np.newaxis, np.newaxis, :]


def hann_window(window_length, dtype=np.float32):
    """tf.signal.hann_window(periodic=True): a - b*cos(2*pi*n/N), evaluated in dtype."""
    dt = np.dtype(dtype).type
    if window_length == 1:
        return np.ones([1], dtype=dt)
    # tf.signal window_ops._raised_cosine_window: n = window_length + periodic * even - 1, so a periodic
    # window of EVEN length divides by window_length and one of ODD length by window_length - 1 (the
    # periodic flag has no effect there: hann(257)[k] = 0.5 - 0.5 cos(2 pi k / 256))
    even = 1 - window_length % 2
    n = dt(window_length + even - 1)
    count = np.arange(window_length).astype(dt)
    cos_arg = dt(2 * np.pi) * count / n
    return (dt(0.5) - dt(0.5) * np.cos(cos_arg)).astype(dt)


def upsample_with_windows(inputs, n_timesteps, add_endpoint=True):
    """ddsp.core.upsample_with_windows (reached from modules/inharm_synth.py:118-119
    via resample(method='window')): overlapping Hann windows, hop = n_timesteps/F.
    """
    inputs = tf_float32(inputs)
    dt = _dt(inputs)
    if inputs.ndim != 3:
        raise ValueError('Upsample_with_windows() only supports 3 dimensions, '
                         'not {}.'.format(inputs.shape))
    if add_endpoint:
        inputs = np.concatenate([inputs, inputs[:, -1:, :]], axis=1)
    n_frames = int(inputs.shape[1])
    n_intervals = n_frames - 1
    if n_frames >= n_timesteps:
        raise ValueError('Upsample with windows cannot be used for downsampling'
                         'More input frames ({}) than output timesteps ({})'.format(
                             n_frames, n_timesteps))
    if n_timesteps % n_intervals != 0.0:
        minus_one = '' if add_endpoint else ' - 1'
        raise ValueError('For upsampling, the target the number of timesteps must be '
                         'divisible by the number of input frames{}. (timesteps:{}, '
                         'frames:{}, add_endpoint={}).'.format(
                             minus_one, n_timesteps, n_frames, add_endpoint))
    hop_size = n_timesteps // n_intervals
    window_length = 2 * hop_size
    window = hann_window(window_length, dt)
    # [B, C, frames, window] then overlap-and-add with hop = hop_size.
    x = np.transpose(inputs, [0, 2, 1])[:, :, :, np.newaxis]
    x_windowed = x * window[np.newaxis, np.newaxis, np.newaxis, :]
    b, c = x_windowed.shape[:2]
    out = np.zeros([b, c, (n_frames + 1) * hop_size], dtype=dt)
    first, second = x_windowed[..., :hop_size], x_windowed[..., hop_size:]
    out[:, :, :n_frames * hop_size] += first.reshape(b, c, -1)
    out[:, :, hop_size:] += second.reshape(b, c, -1)
    out = np.transpose(out, [0, 2, 1])
    return out[:, hop_size:-hop_size, :]


def _resize_bilinear_legacy(inputs, n_timesteps):
    """tf.compat.v1.image.resize(BILINEAR, align_corners=False) along axis 1 of a
    [B, F, C] tensor, i.e. TensorFlow's legacy ResizeBilinear CPU kernel with
    half_pixel_centers=False: scale = F / float(N); in = float(i) * scale;
    lower = floor(in); upper = min(ceil(in), F - 1); lerp = in - floor(in);
    out = top + (bottom - top) * lerp (separate multiply and add).
    """
    dt = _dt(inputs)
    n_frames = inputs.shape[1]
    scale = dt(n_frames) / dt(n_timesteps)
    pos = np.arange(n_timesteps).astype(dt) * scale
    pos_floor = np.floor(pos)
    lower = np.maximum(pos_floor.astype(np.int64), 0)
    upper = np.minimum(np.ceil(pos).astype(np.int64), n_frames - 1)
    lerp = (pos - pos_floor).astype(dt)
    top = inputs[:, lower, :]
    bottom = inputs[:, upper, :]
    return top + (bottom - top) * lerp[np.newaxis, :, np.newaxis]


def resample(inputs, n_timesteps, method='linear', add_endpoint=True):
    """ddsp.core.resample for 3-D [B, F, C] input (call sites
    modules/inharm_synth.py:117 'linear' and :118-119 'window')."""
    inputs = tf_float32(inputs)
    if inputs.ndim != 3:
        raise ValueError('oracle resample() restates the 3-D case only')
    if method == 'linear':
        if not add_endpoint:
            raise ValueError('oracle restates add_endpoint=True only')
        return _resize_bilinear_legacy(inputs, n_timesteps)
    elif method == 'window':
        return upsample_with_windows(inputs, n_timesteps, add_endpoint)
    raise ValueError('Method ({}) is invalid. Must be one of {}.'.format(
        method, "['linear', 'window']"))


def angular_cumsum(angular_frequency, chunk_size=1000):
    """ddsp.core.angular_cumsum (call site modules/inharm_synth.py:75): cumsum in
    chunks of 1000 samples, each chunk wrapped to [0, 2*pi) before being carried
    into the next.  In float32 the in-chunk cumsum is a sequential float32 sum
    (Eigen scan on the TF CPU path), which np.cumsum reproduces.
    """
    dt = _dt(angular_frequency)
    two_pi = dt(2.0 * np.pi)
    n_batch, n_time = angular_frequency.shape[:2]
    ch_shape = list(angular_frequency.shape[2:])
    remainder = n_time % chunk_size
    if remainder:
        pad = [(0, 0)] * angular_frequency.ndim
        pad[1] = (0, chunk_size - remainder)
        angular_frequency = np.pad(angular_frequency, pad)
    length = angular_frequency.shape[1]
    n_chunks = length // chunk_size
    chunks = angular_frequency.reshape([n_batch, n_chunks, chunk_size] + ch_shape)
    phase = np.cumsum(chunks, axis=2, dtype=dt)
    offsets = np.mod(phase[:, :, -1:, ...], two_pi)
    offsets = np.concatenate([np.zeros_like(offsets[:, :1]), offsets], axis=1)[:, :-1]
    offsets = np.mod(np.cumsum(offsets, axis=1, dtype=dt), two_pi)
    phase = phase + offsets
    phase = np.mod(phase, two_pi)
    phase = phase.reshape([n_batch, length] + ch_shape)
    if remainder:
        phase = phase[:, :n_time]
    return phase


def apply_window_to_impulse_response(impulse_response, window_size=0, causal=False):
    """ddsp.core.apply_window_to_impulse_response: zero-phase IR -> windowed causal IR."""
    impulse_response = tf_float32(impulse_response)
    dt = _dt(impulse_response)
    if causal:
        impulse_response = np.fft.fftshift(impulse_response, axes=-1)
    ir_size = int(impulse_response.shape[-1])
    if (window_size <= 0) or (window_size > ir_size):
        window_size = ir_size
    window = hann_window(window_size, dt)
    padding = ir_size - window_size
    if padding > 0:
        half_idx = (window_size + 1) // 2
        window = np.concatenate([window[half_idx:], np.zeros([padding], dt),
                                 window[:half_idx]], axis=0)
    else:
        window = np.fft.fftshift(window, axes=-1)
    impulse_response = window * impulse_response
    if padding > 0:
        first_half_start = (ir_size - (half_idx - 1)) + 1
        second_half_end = half_idx + 1
        impulse_response = np.concatenate([impulse_response[..., first_half_start:],
                                           impulse_response[..., :second_half_end]],
                                          axis=-1)
    else:
        impulse_response = np.fft.fftshift(impulse_response, axes=-1)
    return impulse_response


def frequency_impulse_response(magnitudes, window_size=0):
    """ddsp.core.frequency_impulse_response: irfft of real magnitudes (zero phase),
    Hann-windowed, rotated to causal form.  Length 2*(M-1)."""
    magnitudes = tf_float32(magnitudes)
    dt = _dt(magnitudes)
    cdt = np.complex64 if dt == np.float32 else np.complex128
    impulse_response = np.fft.irfft(magnitudes.astype(cdt), axis=-1).astype(dt)
    return apply_window_to_impulse_response(impulse_response, window_size)


def get_fft_size(frame_size, ir_size, power_of_2=True):
    """ddsp.core.get_fft_size."""
    convolved_frame_size = ir_size + frame_size - 1
    if power_of_2:
        fft_size = int(2 ** np.ceil(np.log2(convolved_frame_size)))
    else:
        raise ValueError('oracle restates power_of_2=True only')
    return fft_size


def crop_and_compensate_delay(audio, audio_size, ir_size, padding, delay_compensation):
    """ddsp.core.crop_and_compensate_delay."""
    if padding == 'valid':
        crop_size = ir_size + audio_size - 1
    elif padding == 'same':
        crop_size = audio_size
    else:
        raise ValueError('Padding must be \'valid\' or \'same\', instead '
                         'of {}.'.format(padding))
    total_size = int(audio.shape[-1])
    crop = total_size - crop_size
    start = ((ir_size - 1) // 2 - 1 if delay_compensation < 0 else delay_compensation)
    end = crop - start
    return audio[:, start:total_size - end]


def fft_convolve(audio, impulse_response, padding='same', delay_compensation=-1):
    """ddsp.core.fft_convolve (call site modules/fdn_reverb.py:409; reached from
    modules/filtered_noise_synth.py:41 through frequency_filter and from
    ddsp.effects.Reverb.get_signal): framed FFT convolution + overlap-add.
    audio [B, N]; impulse_response [B, Lir] or [B, n_frames, Lir]."""
    audio, impulse_response = tf_float32(audio), tf_float32(impulse_response)
    dt = _dt(audio)
    batch_size, audio_size = audio.shape
    if impulse_response.ndim == 2:
        impulse_response = impulse_response[:, np.newaxis, :]
    if impulse_response.shape[0] == 1 and batch_size > 1:
        impulse_response = np.tile(impulse_response, [batch_size, 1, 1])
    batch_size_ir, n_ir_frames, ir_size = impulse_response.shape
    if batch_size != batch_size_ir:
        raise ValueError('Batch size of audio ({}) and impulse response ({}) must '
                         'be the same.'.format(batch_size, batch_size_ir))
    frame_size = int(np.ceil(audio_size / n_ir_frames))
    hop_size = frame_size
    # tf.signal.frame(pad_end=True)
    n_audio_frames = -(-audio_size // hop_size)
    padded = np.zeros([batch_size, n_audio_frames * frame_size], dt)
    padded[:, :audio_size] = audio
    audio_frames = padded.reshape(batch_size, n_audio_frames, frame_size)
    if n_audio_frames != n_ir_frames:
        raise ValueError(
            'Number of Audio frames ({}) and impulse response frames ({}) do not '
            'match. For small hop size = ceil(audio_size / n_ir_frames), '
            'number of impulse response frames must be a multiple of the audio '
            'size.'.format(n_audio_frames, n_ir_frames))
    fft_size = get_fft_size(frame_size, ir_size, power_of_2=True)
    audio_fft = np.fft.rfft(audio_frames, fft_size)
    ir_fft = np.fft.rfft(impulse_response, fft_size)
    audio_ir_fft = audio_fft * ir_fft
    audio_frames_out = np.fft.irfft(audio_ir_fft, fft_size).astype(dt)
    # tf.signal.overlap_and_add(frames, hop)
    total = (n_audio_frames - 1) * hop_size + fft_size
    audio_out = np.zeros([batch_size, total], dt)
    if fft_size <= 2 * hop_size or n_audio_frames == 1:
        for i in range(n_audio_frames):
            audio_out[:, i * hop_size:i * hop_size + fft_size] += audio_frames_out[:, i]
    else:
        # Same sums, vectorised over frames: segment s of every frame lands s hops later.
        n_seg = -(-fft_size // hop_size)
        pad_w = n_seg * hop_size - fft_size
        fr = np.pad(audio_frames_out, [(0, 0), (0, 0), (0, pad_w)])
        fr = fr.reshape(batch_size, n_audio_frames, n_seg, hop_size)
        acc = np.zeros([batch_size, n_audio_frames + n_seg - 1, hop_size], dt)
        for s in range(n_seg):
            acc[:, s:s + n_audio_frames] += fr[:, :, s]
        audio_out = acc.reshape(batch_size, -1)[:, :total]
    return crop_and_compensate_delay(audio_out, audio_size, ir_size, padding,
                                     delay_compensation)


def frequency_filter(audio, magnitudes, window_size=0, padding='same'):
    """ddsp.core.frequency_filter (call site modules/filtered_noise_synth.py:41)."""
    impulse_response = frequency_impulse_response(magnitudes, window_size=window_size)
    return fft_convolve(audio, impulse_response, padding=padding)


def nested_lookup(nested_key, nested_dict, delimiter='/'):
    """ddsp.core.nested_lookup: 'a/b' -> nested_dict['a']['b']."""
    keys = nested_key.split(delimiter)
    value = nested_dict
    for key in keys:
        try:
            value = value[key]
        except KeyError:
            raise KeyError(f'Key \'{key}\' as a part of nested key \'{nested_key}\' '
                           'not found during nested dictionary lookup, out of '
                           f'available keys: {list(nested_dict.keys())}')
    return value
