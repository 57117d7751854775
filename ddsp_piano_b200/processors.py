"""The reference's operator API for the synthesis hot path, backed by the sm_100a kernels.

Same class names, constructor arguments, method names, dictionary keys and DAG format as
the reference, so ``configs/dafx22.gin:91-111`` can bind these in place of its own:

  reference                                                      here
  ---------------------------------------------------------------------------------------
  ddsp.processors.Processor / ProcessorGroup (ddsp v3.7.0)        Processor / ProcessorGroup
  modules/inharm_synth.py:130-244  InHarmonic                     InHarmonic
  modules/inharm_synth.py:247-293  MultiInharmonic                MultiInharmonic
  modules/inharm_synth.py:296-309  MultiAdd                       MultiAdd
  modules/filtered_noise_synth.py:12-42 DynamicSizeFilteredNoise  DynamicSizeFilteredNoise
  ddsp.effects.Reverb (trainable=False)                           Reverb
  modules/polyphonic_dag.py:6-42   polyphonic_dag                 polyphonic_dag

Tensors are ``torch.Tensor`` on a CUDA device (float32); every ``get_signal`` runs
hand-written CUDA through the C ABI (``include/b200ddsp.h``).  ``ProcessorGroup`` walks the
DAG node by node exactly like ddsp's ``DAGLayer``; when the DAG has the shape
``polyphonic_dag`` produces it instead issues ONE ``b200ddsp_forward_polyphonic`` call
(``fused=True``, the default).
"""
import os

import torch

from .engine import get_engine, scale_fn_id


def exp_sigmoid(x, exponent=10.0, max_value=2.0, threshold=1e-7):
    """ddsp.core.exp_sigmoid.  Passed as ``scale_fn`` it only selects the in-kernel scaling;
    called directly it is evaluated with torch ops on ``x``'s device."""
    x = torch.as_tensor(x, dtype=torch.float32)
    return max_value * torch.sigmoid(x) ** torch.log(torch.tensor(exponent)).item() + threshold


def exp_tanh(x, max_value=2., exponent=10., gain=1., threshold=1e-7):
    """modules/inharm_synth.py:13-17."""
    x = torch.as_tensor(x, dtype=torch.float32)
    y = max_value * (0.5 * (torch.tanh(gain * x) + 1.)) ** torch.log(torch.tensor(exponent)).item()
    return y + threshold


def nested_lookup(nested_key, nested_dict, delimiter='/'):
    """ddsp.core.nested_lookup."""
    value = nested_dict
    for key in nested_key.split(delimiter):
        try:
            value = value[key]
        except KeyError:
            raise KeyError(f'Key \'{key}\' as a part of nested key \'{nested_key}\' '
                           'not found during nested dictionary lookup, out of '
                           f'available keys: {list(nested_dict.keys())}')
    return value


_DEFAULT_CFG = dict(sample_rate=16000, frame_rate=250, min_frequency=20.0, additive_scale_fn=0,
                    normalize_after_nyquist_cut=1, normalize_below_nyquist=1, inference=1,
                    noise_scale_fn=0, noise_initial_bias=-5.0, noise_window_size=257,
                    reverb_add_dry=1, n_noise_bands=0, fast_phase=0)


def _device_of(*tensors):
    for t in tensors:
        if isinstance(t, torch.Tensor):
            return t.device
    return torch.device('cuda', torch.cuda.current_device())


class Processor:
    """ddsp.processors.Processor: ``__call__`` = ``get_signal(**get_controls(*args))``."""

    def __init__(self, name, trainable=True):
        self.name = name
        self.trainable = trainable

    def __call__(self, *args, return_outputs_dict=False, **kwargs):
        for k in ['training', 'mask']:
            kwargs.pop(k, None)
        controls = self.get_controls(*args, **kwargs)
        signal = self.get_signal(**controls)
        if return_outputs_dict:
            return dict(signal=signal, controls=controls)
        return signal

    def get_controls(self, *args, **kwargs):
        raise NotImplementedError

    def get_signal(self, *args, **kwargs):
        raise NotImplementedError

    def engine_config(self):
        """ABI configuration fields this processor determines."""
        return {}


class InHarmonic(Processor):
    """modules/inharm_synth.py:130-244: bank of inharmonic cosine oscillators.

    ``fast_phase`` (not in the reference): start every 256-sample unit from a closed-form double-precision
    phase instead of reproducing the reference's float32 ``angular_cumsum`` bit for bit -- about 20 % less
    time per forward, output within ~3e-4 of the exact (float64) synthesis but NOT within 1e-4 of the
    reference, whose own float32 phase drifts (``b200ddsp_config.fast_phase``)."""

    def __init__(self, frame_rate=250, sample_rate=16000, min_frequency=20,
                 scale_fn=exp_sigmoid, normalize_after_nyquist_cut=True,
                 normalize_below_nyquist=True, inference=False, name='inharmonic', fast_phase=False):
        self.frame_rate = frame_rate
        self.sample_rate = sample_rate
        self.min_frequency = min_frequency
        self.normalize_after_nyquist_cut = normalize_after_nyquist_cut
        self.scale_fn = scale_fn
        self.normalize_below_nyquist = normalize_below_nyquist
        self.inference = inference
        self.fast_phase = fast_phase
        super().__init__(name=name)

    @property
    def upsampling(self):
        return int(self.sample_rate / self.frame_rate)

    def engine_config(self):
        return dict(sample_rate=int(self.sample_rate), frame_rate=int(self.frame_rate),
                    min_frequency=float(self.min_frequency),
                    additive_scale_fn=scale_fn_id(self.scale_fn),
                    normalize_after_nyquist_cut=int(bool(self.normalize_after_nyquist_cut)),
                    normalize_below_nyquist=int(bool(self.normalize_below_nyquist)),
                    inference=int(bool(self.inference)), fast_phase=int(bool(self.fast_phase)))

    def _engine(self, *tensors):
        return get_engine(_device_of(*tensors), **{**_DEFAULT_CFG, **self.engine_config()})

    def get_controls(self, amplitudes, harmonic_distribution, inharm_coef, f0_hz):
        return self._engine(amplitudes, f0_hz).additive_controls(
            amplitudes, harmonic_distribution, inharm_coef, f0_hz)

    def get_signal(self, amplitudes, harmonic_distribution, harmonic_shifts, f0_hz):
        return self._engine(amplitudes, f0_hz).additive_signal(
            amplitudes, harmonic_distribution, harmonic_shifts, f0_hz)


class SurrogateAdditive(InHarmonic):
    """modules/surrogate_synth.py:107-214 (configs/surrogate.gin): the inharmonic bank of one string
    with exponentially decaying partial amplitudes, ``|decays| ** (decay_time * U + r)`` per sample
    (B. Hayes, sinusoidal frequency estimation by gradient descent).  Forward only (both cumsum modes of
    ``inference``); it runs on the generic oscillator kernel."""

    def __init__(self, frame_rate=250, sample_rate=16000, min_frequency=20,
                 normalize_harm_distribution=True, scale_fn=exp_sigmoid, normalize_below_nyquist=True,
                 inference=False, name='inharmonic'):
        super().__init__(frame_rate=frame_rate, sample_rate=sample_rate, min_frequency=min_frequency,
                         scale_fn=scale_fn, normalize_after_nyquist_cut=True,
                         normalize_below_nyquist=normalize_below_nyquist, inference=inference, name=name)
        self.normalize_harm_distribution = normalize_harm_distribution

    def engine_config(self):
        cfg = super().engine_config()
        if not self.normalize_harm_distribution:               # surrogate_synth.py:183-187 skipped
            cfg['normalize_after_nyquist_cut'] = 2
        return cfg

    def get_controls(self, amplitudes, decays, decay_time, harmonic_distribution, inharm_coef, f0_hz):
        eng = self._engine(amplitudes, f0_hz)
        ctl = eng.additive_controls(amplitudes, harmonic_distribution, inharm_coef, f0_hz)
        return {'amplitudes': ctl['amplitudes'],
                'decays': None if decays is None else eng.surrogate_decays(decays, inharm_coef, f0_hz),
                'decay_time': decay_time, 'harmonic_distribution': ctl['harmonic_distribution'],
                'harmonic_shifts': ctl['harmonic_shifts'], 'f0_hz': ctl['f0_hz']}

    def get_signal(self, amplitudes, decays, decay_time, harmonic_distribution, harmonic_shifts, f0_hz):
        eng = self._engine(amplitudes, f0_hz)
        if decays is None or decay_time is None:                               # surrogate_synth.py:78
            return eng.additive_signal(amplitudes, harmonic_distribution, harmonic_shifts, f0_hz)
        return eng.surrogate_signal(amplitudes, decays, decay_time, harmonic_distribution,
                                    harmonic_shifts, f0_hz)


class MultiInharmonic(InHarmonic):
    """modules/inharm_synth.py:247-293: one oscillator bank per substring (f0_hz [B, F, S]),
    sharing amplitudes (pre-divided by S) and inharmonic shifts.  The S syntheses and their
    sum are one kernel launch here."""

    def __init__(self, name='multi_inharmonic', **kwargs):
        super().__init__(name=name, **kwargs)


class MultiAdd(Processor):
    """modules/inharm_synth.py:296-309: sum an arbitrary number of signals."""

    def __init__(self, name='add'):
        super().__init__(name=name)

    def get_controls(self, *signals):
        return {f'signal_{i}': s for i, s in enumerate(signals)}

    def get_signal(self, **signals):
        return sum(signals.values())


class DynamicSizeFilteredNoise(Processor):
    """modules/filtered_noise_synth.py:12-42 on top of ddsp.synths.FilteredNoise
    (``n_samples`` is accepted and ignored, like in the reference: the length follows the
    controls).  The reference draws unseeded uniform noise per call; here it comes from an
    in-kernel Philox stream keyed by (``seed``, call counter), or -- for parity tests -- from
    tensors queued with :meth:`push_noise`."""

    def __init__(self, frame_rate=250, sample_rate=16000, n_samples=64000, window_size=257,
                 scale_fn=exp_sigmoid, initial_bias=-5.0, name='filtered_noise', seed=None):
        super().__init__(name=name)
        self.frame_rate = frame_rate
        self.sample_rate = sample_rate
        self.n_samples = n_samples
        self.window_size = window_size
        self.scale_fn = scale_fn
        self.initial_bias = initial_bias
        self.seed = int.from_bytes(os.urandom(8), 'little') if seed is None else int(seed)
        self._calls = 0
        self._injected = []

    @property
    def upsampling(self):
        return int(self.sample_rate / self.frame_rate)

    def push_noise(self, noise):
        """Queue a [B, N] tensor to be filtered by the next ``get_signal`` instead of fresh
        noise (FIFO)."""
        self._injected.append(noise)

    def pop_noise(self):
        return self._injected.pop(0) if self._injected else None

    def next_stream_id(self, n=1):
        sid = self._calls
        self._calls += n
        return sid

    def engine_config(self, n_bands=0):
        return dict(sample_rate=int(self.sample_rate), frame_rate=int(self.frame_rate),
                    noise_scale_fn=scale_fn_id(self.scale_fn),
                    noise_initial_bias=float(self.initial_bias),
                    noise_window_size=int(self.window_size), n_noise_bands=int(n_bands))

    def _engine(self, magnitudes):
        cfg = {**_DEFAULT_CFG, **self.engine_config(magnitudes.shape[-1])}
        return get_engine(_device_of(magnitudes), **cfg)

    def get_controls(self, magnitudes):
        return self._engine(magnitudes).noise_controls(magnitudes)

    def get_signal(self, magnitudes):
        return self._engine(magnitudes).noise_signal(
            magnitudes, noise=self.pop_noise(), seed=self.seed & (2 ** 64 - 1),
            stream_id=self.next_stream_id())


class Reverb(Processor):
    """ddsp.effects.Reverb with ``trainable=False`` (configs/dafx22.gin:99-100,111): the
    impulse response is an input; ``ir[:, 0]`` is masked and the dry signal added."""

    def __init__(self, trainable=False, reverb_length=48000, add_dry=True, name='reverb'):
        super().__init__(name=name, trainable=trainable)
        if trainable:
            raise ValueError('Reverb(trainable=True) owns a tf.Variable in ddsp; the hot path '
                             'implements the trainable=False form the reference configures')
        self._reverb_length = reverb_length
        self._add_dry = add_dry

    def engine_config(self):
        return dict(reverb_add_dry=int(bool(self._add_dry)))

    def get_controls(self, audio, ir=None):
        if ir is None:
            raise ValueError('Must provide "ir" tensor if Reverb trainable=False.')
        return {'audio': audio, 'ir': ir}

    def get_signal(self, audio, ir):
        eng = get_engine(_device_of(audio, ir), **{**_DEFAULT_CFG, **self.engine_config()})
        return eng.reverb(audio, ir)


class MultiInstrumentReverb:
    """modules/sub_modules.py:300-365: one learnt impulse response per instrument (an embedding
    table [n_instruments, reverb_length], e.g. ``Checkpoint.tensor('reverb_model/reverb_dict')``);
    ``__call__(piano_model [B, 1])`` returns ``reverb_ir [B, L]``, with the exponential decay mask
    (:339-349) when ``inference=True``.  The lookup is an index; the mask is a CUDA kernel."""

    def __init__(self, embeddings, n_instruments=None, reverb_duration=None, sample_rate=16000,
                 inference=False, name='reverb_model'):
        self.name = name
        self.embeddings = embeddings
        self.sample_rate = sample_rate
        self.inference = inference
        self.n_instruments = embeddings.shape[0] if n_instruments is None else n_instruments
        self.reverb_duration = (embeddings.shape[1] / sample_rate if reverb_duration is None
                                else reverb_duration)

    @property
    def reverb_length(self):
        return int(self.reverb_duration * self.sample_rate)

    def exponential_decay_mask(self, ir, decay_exponent=4., decay_start=16000):
        eng = get_engine(_device_of(ir), **_DEFAULT_CFG)
        return eng.ir_decay_mask(ir, decay_exponent, decay_start)

    def __call__(self, piano_model):
        idx = torch.as_tensor(piano_model, device=self.embeddings.device).long()
        if self.n_instruments == 1:
            idx = torch.zeros_like(idx)
        ir = self.embeddings[idx]                       # [B, 1, L] for piano_model [B, 1]
        if ir.dim() == 3:
            ir = ir[:, 0]
        ir = ir.contiguous()
        if self.inference:
            ir = self.exponential_decay_mask(ir)
        return ir

    call = __call__


class FeedbackDelayNetwork(Processor):
    """modules/fdn_reverb.py:20-410: frequency-sampled feedback delay network (Householder mixing,
    one-pole reverberation-time control, 4 allpasses per line, FIR early reflections).  ``get_ir`` is
    what ``MultiInstrumentFeedbackDelayReverb`` (modules/sub_modules.py:368-446, the reverb model of
    configs/maestro-v2.gin) calls to produce ``reverb_ir``; as a DAG processor (configs/ENSTDkCl-*.gin)
    ``get_signal`` convolves the dry audio with that IR.

    ``trainable=False`` (8 lines, fdn_reverb.py:95-97): the parameters are inputs of ``get_controls``.
    ``trainable=True`` (:121-174, the ENSTDkCl gins with 6 lines and trainable delays): the processor owns
    them -- ``self.parameters``, drawn like the reference's initialisers or set with
    :meth:`load_parameters` from a checkpoint -- and ``get_controls(audio_dry)`` needs nothing else.
    Forward only: nothing here computes gradients."""

    PARAMETERS = ('input_gain', 'output_gain', 'gain_allpass', 'delays_allpass', 'time_rev_0_sec',
                  'alpha_tone', 'early_ir')

    def __init__(self, trainable=False, name='DelayNetwork', sampling_rate=16000.0, delay_lines=8,
                 delay_values=None, delays_allpass=None, early_ir_length=200, early_reflections=6,
                 time_control_bands=6, delay_trainable=False, seed=0):
        super().__init__(name=name, trainable=trainable)
        self.sampling_rate = float(sampling_rate)
        self.freq_points = int(2 * self.sampling_rate)
        self.delay_values = delay_values
        self.delays_allpass = delays_allpass
        self.early_ir_length = early_ir_length
        self.delay_lines = delay_lines
        self.delay_trainable = delay_trainable
        self.parameters = None
        self._seed = seed
        self.build()

    def __len__(self):
        return self.delay_lines

    def build(self, input_shape=None):
        """fdn_reverb.py:92-175: fixed delay values unless they are trainable; the variables of the
        trainable form, from the reference's initialisers (normal(0.25, 0.1) gains, normal(400, 60) delays,
        normal(2, 0.5) >= 0 reverberation time, normal(0, 0.1) early response and tone)."""
        if self.delay_values is None and not (self.trainable and self.delay_trainable):
            self.delay_values = [233., 311., 421., 461., 587., 613., 789., 891.]
            self.delay_lines = len(self.delay_values)
        if self.delay_lines not in (6, 8):
            raise ValueError(f'delay_lines={self.delay_lines}: the kernels are built for 8 lines '
                             '(fdn_reverb.py:30) or 6 (configs/ENSTDkCl-*.gin)')
        if self.trainable and self.parameters is None:
            g = torch.Generator().manual_seed(self._seed)
            D = self.delay_lines
            normal = lambda mean, std, *shape: torch.normal(mean, std, size=shape, generator=g)
            self.parameters = {
                'early_ir': normal(0.0, 0.1, self.early_ir_length),
                'input_gain': normal(0.25, 0.1, D), 'output_gain': normal(0.25, 0.1, D),
                'time_rev_0_sec': normal(2.0, 0.5, 1).clamp_min(0.0), 'alpha_tone': normal(0.0, 0.1, 1),
                'delays_allpass': normal(400.0, 60.0, D, 4), 'gain_allpass': normal(0.25, 0.1, D, 4)}
            if self.delay_trainable and self.delay_values is None:
                self.parameters['delay_values'] = normal(400.0, 60.0, D)

    def load_parameters(self, values):
        """Set the variables of the trainable form (raw values: ``alpha_tone`` BEFORE its sigmoid, like the
        reference's variable ``_alpha_tone``)."""
        if not self.trainable:
            raise ValueError('only FeedbackDelayNetwork(trainable=True) owns parameters')
        for k, v in values.items():
            if k not in self.parameters and k != 'delay_values':
                raise KeyError(k)
            self.parameters[k] = torch.as_tensor(v, dtype=torch.float32).reshape(
                self.parameters[k].shape if k in self.parameters else [self.delay_lines])

    def _engine(self, *tensors):
        return get_engine(_device_of(*tensors), **_DEFAULT_CFG)

    def get_ir(self, input_gain, output_gain, gain_allpass, delays_allpass, time_rev_0_sec,
               alpha_tone, early_ir, device=None):
        delays = self.delay_values
        if self.trainable and self.parameters.get('delay_values') is not None:
            delays = self.parameters['delay_values']
        eng = get_engine(device, **_DEFAULT_CFG) if device is not None else self._engine(input_gain, early_ir)
        return eng.fdn_ir(input_gain, output_gain, gain_allpass, delays_allpass, time_rev_0_sec, alpha_tone,
                          early_ir, self.sampling_rate, delays)

    def get_controls(self, audio_dry=None, input_gain=None, output_gain=None, gain_allpass=None,
                     delays_allpass=None, time_rev_0_sec=None, alpha_tone=None, early_ir=None):
        if self.trainable:                                              # fdn_reverb.py:382-391
            dev = _device_of(audio_dry)
            p = {k: self.parameters[k].to(dev) for k in self.PARAMETERS}
            p['alpha_tone'] = torch.sigmoid(p['alpha_tone'])
            ir = self.get_ir(**p, device=dev)
        else:
            ir = self.get_ir(input_gain, output_gain, gain_allpass, delays_allpass, time_rev_0_sec,
                             alpha_tone, early_ir)
        return {'audio': audio_dry, 'ir': ir}

    def get_signal(self, audio, ir):
        # fft_convolve(audio, ir[None], delay_compensation=0): no dry path, ir[0] kept
        return self._engine(audio, ir).fft_convolve(audio, ir)


def polyphonic_dag(additive, noise, reverb=None,
                   additive_controls=['amps', 'harmonic_distribution', 'f0_hz'],
                   noise_controls=['noise_magnitudes'], reverb_controls=[], n_synths=16):
    """modules/polyphonic_dag.py:6-42: per voice additive_i, noise_i and a running sum
    'add', then the reverb on 'add/signal'."""
    add = MultiAdd(name='add')
    dag = [(additive, [c + '_0' for c in additive_controls]),
           (noise, [c + '_0' for c in noise_controls]),
           (add, [noise.name + '/signal', additive.name + '/signal'])]
    for i in range(1, n_synths):
        dag.append((additive, [c + f'_{i}' for c in additive_controls]))
        dag.append((noise, [c + f'_{i}' for c in noise_controls]))
        dag.append((add, ['add/signal', noise.name + '/signal', additive.name + '/signal']))
    if reverb is not None:
        dag.append((reverb, ['add/signal'] + reverb_controls))
    return dag


class ProcessorGroup:
    """ddsp.processors.ProcessorGroup (a ``DAGLayer``): runs ``(processor, [input keys])``
    nodes in order over a growing outputs dict; ``'a/b'`` keys are nested lookups; each
    node's ``{'signal', 'controls'}`` is stored under the processor's name and the last
    one aliased as ``'out'``.  Call site: modules/piano_model.py:160-164.

    ``fused=True`` (default): a DAG of the ``polyphonic_dag`` shape runs as one C-ABI call.
    The outputs dict then holds ``add``, ``out`` (and ``reverb``); per-voice entries
    (``additive``, ``noise``), which in the reference hold the LAST voice only because the
    processor objects are shared, are produced only by the node-by-node walk
    (``fused=False``)."""

    def __init__(self, dag, name='processor_group', fused=True):
        self.name = name
        self.dag = []
        self._modules = {}
        for node in dag:
            module, keys = node[0], node[1]
            self._modules[module.name] = module
            self.dag.append((module.name, list(keys)))
        self.fused = fused
        self._plan = self._match_polyphonic() if fused else None

    @property
    def processors(self):
        return [self._modules[k] for k, _ in self.dag]

    # -- pattern match --------------------------------------------------------------------
    def _match_polyphonic(self):
        dag = self.dag
        n = len(dag)
        has_reverb = n % 3 == 1
        if n < 3 or n % 3 == 2:
            return None
        P = n // 3
        add_name, noise_name, sum_name = dag[0][0], dag[1][0], dag[2][0]
        additive, noise, add = (self._modules[add_name], self._modules[noise_name],
                                self._modules[sum_name])
        if not (isinstance(additive, InHarmonic) and isinstance(noise, DynamicSizeFilteredNoise)
                and isinstance(add, MultiAdd)):
            return None
        voices = []
        for i in range(P):
            a, z, s = dag[3 * i], dag[3 * i + 1], dag[3 * i + 2]
            if (a[0], z[0], s[0]) != (add_name, noise_name, sum_name):
                return None
            if len(a[1]) != 4 or len(z[1]) != 1:
                return None
            want = [noise_name + '/signal', add_name + '/signal']
            if i > 0:
                want = [sum_name + '/signal'] + want
            if s[1] != want:
                return None
            voices.append((a[1], z[1][0]))
        reverb, ir_key = None, None
        if has_reverb:
            r = dag[-1]
            reverb = self._modules[r[0]]
            if r[1][0] != sum_name + '/signal':
                return None
            if isinstance(reverb, FeedbackDelayNetwork):
                # configs/ENSTDkCl-*.gin:99-100: the network itself closes the DAG (its parameters are its
                # own, or -- trainable=False -- 7 more feature keys)
                if len(r[1]) != (1 if reverb.trainable else 8):
                    return None
                ir_key = list(r[1][1:])
            elif isinstance(reverb, Reverb) and len(r[1]) == 2:
                ir_key = r[1][1]
            else:
                return None
        if int(additive.sample_rate) != int(noise.sample_rate) or \
                int(additive.frame_rate) != int(noise.frame_rate):
            return None
        return dict(additive=additive, noise=noise, add=add, reverb=reverb, voices=voices,
                    ir_key=ir_key)

    # -- execution ------------------------------------------------------------------------
    def _run_fused(self, outputs, timeline=None):
        plan = self._plan
        additive, noise, reverb = plan['additive'], plan['noise'], plan['reverb']
        voices = []
        def get(k):
            return outputs[k] if k in outputs else nested_lookup(k, outputs)

        for add_keys, mag_key in plan['voices']:
            voices.append({'amplitudes': get(add_keys[0]), 'harmonic_distribution': get(add_keys[1]),
                           'inharm_coef': get(add_keys[2]), 'f0_hz': get(add_keys[3]),
                           'magnitudes': get(mag_key), 'noise': noise.pop_noise()})
        M = voices[0]['magnitudes'].shape[-1]
        cfg = {**_DEFAULT_CFG, **additive.engine_config(), **noise.engine_config(M)}
        fdn_tail = isinstance(reverb, FeedbackDelayNetwork)
        if reverb is not None and not fdn_tail:
            cfg.update(reverb.engine_config())
        f0_0 = voices[0]['f0_hz']
        on_host = not (isinstance(f0_0, torch.Tensor) and f0_0.device.type == 'cuda')
        dev = torch.device('cuda', torch.cuda.current_device()) if on_host else f0_0.device
        eng = get_engine(dev, **cfg)
        ir = nested_lookup(plan['ir_key'], outputs) if reverb is not None and not fdn_tail else None
        seed = (noise.seed + 0x9E3779B97F4A7C15 * noise.next_stream_id(len(voices))) & (2 ** 64 - 1)
        if fdn_tail:
            # additive + noise + sums fused as above; the network's own two steps close the DAG:
            # get_controls (the impulse response, two kernels) and get_signal (FFT convolution, no dry path)
            if timeline is not None or on_host:
                raise ValueError('a DAG closed by FeedbackDelayNetwork takes whole clips of device tensors')
            dry, _ = eng.forward_polyphonic(voices, reverb_ir=None, seed=seed)
            ctl = reverb.get_controls(dry, *[nested_lookup(k, outputs) for k in plan['ir_key']])
            outputs[plan['add'].name] = {'signal': dry, 'controls': {}}
            outputs[reverb.name] = {'signal': reverb.get_signal(ctl['audio'], ctl['ir']), 'controls': ctl}
            outputs['out'] = outputs[reverb.name]
            return outputs
        if timeline is not None:
            # one span of a timeline (sharding.SpanChain): the features cover the span's input frames
            if reverb is None:
                raise ValueError('a timeline call needs the reverb node (its tail is what crosses spans)')
            dry, wet = eng.forward_timeline(voices, ir, timeline['span'], timeline['seg_frames'],
                                            tail=timeline.get('tail'), seed=seed)
        elif on_host:
            # HOST features (numpy / CPU tensors): staged copies overlap the kernels; the signals
            # come back as pinned host tensors, valid after torch.cuda.current_stream().synchronize()
            dry, wet = eng.forward_polyphonic_host(voices, reverb_ir=ir, seed=seed)
        else:
            dry, wet = eng.forward_polyphonic(voices, reverb_ir=ir, seed=seed)
        outputs[plan['add'].name] = {'signal': dry, 'controls': {}}
        last = outputs[plan['add'].name]
        if reverb is not None:
            outputs[reverb.name] = {'signal': wet, 'controls': {'audio': dry, 'ir': ir}}
            last = outputs[reverb.name]
        outputs['out'] = last
        return outputs

    def get_controls(self, inputs, timeline=None, **kwargs):
        outputs = inputs
        if self._plan is not None:
            return self._run_fused(outputs, timeline)
        if timeline is not None:
            raise ValueError('timeline calls run through the fused polyphonic DAG only')
        module_outputs = None
        for module_key, input_keys in self.dag:
            module = self._modules[module_key]
            args = [nested_lookup(key, outputs) for key in input_keys]
            module_outputs = module(*args, return_outputs_dict=True, **kwargs)
            outputs[module_key] = module_outputs
        outputs['out'] = module_outputs
        return outputs

    def get_signal(self, outputs):
        return outputs['out']['signal']

    def __call__(self, inputs, return_outputs_dict=False, timeline=None, **kwargs):
        """``timeline`` (no counterpart in the reference, which runs a piece in one pass): a dict
        ``{'span': _lib.Span, 'seg_frames': int, 'tail': _lib.Link or None}`` makes this call ONE SPAN of
        a longer timeline -- ``inputs`` then cover the span's input frames (one halo frame either side)
        and the signals its output frames; see ``sharding.SpanChain`` and DESIGN.md section 6."""
        controls = self.get_controls(inputs, timeline=timeline, **kwargs)
        signal = self.get_signal(controls)
        if return_outputs_dict:
            return dict(signal=signal, controls=controls)
        return signal
