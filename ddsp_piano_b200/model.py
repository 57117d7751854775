"""The control-rate graph of ``configs/dafx22.gin`` in front of the synthesis
kernels: MIDI conditioning -> 250 Hz control tensors -> ``ProcessorGroup`` -> audio.

Mirrors the reference's module structure (``ddsp_piano/modules/piano_model.py:146-169`` and the
sub-modules of ``modules/sub_modules.py`` named below) so that a ``PianoModel`` here is called
like the reference's: ``model({'conditioning': [B, F, P, 2], 'pedal': [B, F, 4],
'piano_model': [B, 1]})`` returns the processor group's controls plus ``'audio_synth'``.

This is caller-side plumbing at 1/96 of the audio rate, not the hot path: the dense layers are
torch library calls (cuBLAS, TF32 disabled); the two parts that are sequential in time, the
note-release recurrence and the recurrence of the two GRUs, are CUDA kernels behind the C ABI
(``b200ddsp_note_release``, ``b200ddsp_gru_recurrence``: one launch for all frames).  Weights come from the reference's shipped TensorFlow checkpoint
through the TF-free reader in ``checkpoint.py``.

Parity status: checked in ``tests/test_model.py`` against a numpy restatement of the same graph on
the shipped weights; the Keras / ddsp layer semantics underneath are restated, not pinned
(TensorFlow is not in the container) -- see DESIGN.md section 8.
"""
import os

import numpy as np
import torch

from .checkpoint import Checkpoint, NpzWeights
from .engine import get_engine
from .processors import (_DEFAULT_CFG, DynamicSizeFilteredNoise, MultiInharmonic,
                         MultiInstrumentReverb, ProcessorGroup, Reverb, polyphonic_dag)

MIDI_NORM = 128.0


def _t(x, device):
    return torch.as_tensor(np.ascontiguousarray(x), dtype=torch.float32, device=device)


class Dense:
    """tf.keras.layers.Dense: ``act(x @ kernel + bias)``."""

    def __init__(self, kernel, bias, activation=None):
        self.kernel, self.bias, self.activation = kernel, bias, activation

    def __call__(self, x):
        y = torch.matmul(x, self.kernel) + self.bias
        return self.activation(y) if self.activation is not None else y


# 'cluster': the persistent cluster kernel of this library; 'cudnn': torch.nn.GRU (A/B timing only)
GRU_IMPL = os.environ.get('B200DDSP_GRU', 'cluster')


def leaky_relu(x):
    return torch.nn.functional.leaky_relu(x, 0.2)          # tf.nn.leaky_relu default alpha


class GRU:
    """tf.keras.layers.GRU(units, return_sequences=True) with TF2's ``reset_after=True``: the same
    recurrence as torch.nn.GRU once the gates are reordered from Keras' (z, r, h) to torch's
    (r, z, n) and the [2, 3u] bias is split into input and recurrent halves."""

    def __init__(self, kernel, recurrent_kernel, bias, device):
        u = recurrent_kernel.shape[0]
        order = np.concatenate([np.arange(u, 2 * u), np.arange(0, u), np.arange(2 * u, 3 * u)])
        self.gru = torch.nn.GRU(kernel.shape[0], u, batch_first=True).to(device)
        with torch.no_grad():
            self.gru.weight_ih_l0.copy_(_t(kernel[:, order].T, device))
            self.gru.weight_hh_l0.copy_(_t(recurrent_kernel[:, order].T, device))
            self.gru.bias_ih_l0.copy_(_t(bias[0][order], device))
            self.gru.bias_hh_l0.copy_(_t(bias[1][order], device))
        self.gru.requires_grad_(False)
        self.gru.flatten_parameters()

    def __call__(self, x):
        u = self.gru.hidden_size
        if x.is_cuda and GRU_IMPL == 'cluster' and u in (64, 128, 192, 256):
            # input projections of all frames as one GEMM, then the recurrence as ONE launch
            # (csrc/control_rate.cuh::gru_recurrence_kernel) instead of a library step per frame
            with torch.no_grad():
                R, F, C = x.shape
                xp = torch.addmm(self.gru.bias_ih_l0, x.reshape(R * F, C), self.gru.weight_ih_l0.t())
                return get_engine(x.device, **_DEFAULT_CFG).gru_recurrence(
                    xp.reshape(R, F, 3 * u), self.gru.weight_hh_l0, self.gru.bias_hh_l0)
        with torch.backends.cudnn.flags(enabled=True, allow_tf32=False), torch.no_grad():
            return self.gru(x.contiguous())[0]


class Normalize:
    """ddsp.training.nn.Normalize('layer'): moments over time and channels (every axis but the
    batch), learnt scale/shift [1, 1, 1, C]; ``norm_axes='channels'`` = per-frame moments."""

    def __init__(self, scale, shift, norm_axes='time_channels', eps=1e-5):
        self.scale, self.shift = scale.reshape(1, 1, -1), shift.reshape(1, 1, -1)
        self.dims = (1, 2) if norm_axes == 'time_channels' else (2,)
        self.eps = eps

    def __call__(self, x):
        mean = x.mean(dim=self.dims, keepdim=True)
        var = x.var(dim=self.dims, unbiased=False, keepdim=True)
        return (x - mean) / torch.sqrt(var + self.eps) * self.scale + self.shift


class OneHotZEncoder:
    """sub_modules.py:183-251: instrument id -> z, global_inharm, global_detuning, held over the
    clip (``resample`` of a one-frame embedding is a broadcast)."""

    def __init__(self, embedding, inharm_embedding, detune_embedding):
        self.embedding, self.inharm_embedding, self.detune_embedding = \
            embedding, inharm_embedding, detune_embedding
        self.n_instruments = embedding.shape[0]

    def __call__(self, piano_model, n_frames):
        idx = torch.as_tensor(piano_model, device=self.embedding.device).long().reshape(-1)
        if self.n_instruments == 1:
            idx = torch.zeros_like(idx)
        hold = lambda e: e[idx][:, None, :].expand(-1, n_frames, -1)
        return hold(self.embedding), hold(self.inharm_embedding), hold(self.detune_embedding)


class ContextNetwork:
    """sub_modules.py:18-65 with the layers of dafx22.gin:60-71: Dense 32 (leaky_relu), GRU 64,
    Normalize, then OutputSplitsLayer's dense_out -> context [B, F, 32]."""

    def __init__(self, dense, gru, norm, dense_out, normalize_pitch=False):
        self.dense, self.gru, self.norm, self.dense_out = dense, gru, norm, dense_out
        self.normalize_pitch = normalize_pitch

    def __call__(self, conditioning, pedal, z):
        B, F = conditioning.shape[:2]
        if self.normalize_pitch:
            conditioning = conditioning / torch.tensor([MIDI_NORM, 1.0], device=conditioning.device)
        x = torch.cat([conditioning.reshape(B, F, -1), pedal, z], dim=-1)   # collapse_last_axis :41-48
        return self.dense_out(self.norm(self.gru(self.dense(x))))


class NoteRelease:
    """sub_modules.py:1174-1188 over F0ProcessorCell (:1114-1171): CUDA kernel, one thread per row."""

    def __init__(self, release_duration, frame_rate=250):
        self.release_duration, self.frame_rate = float(release_duration), frame_rate

    def __call__(self, conditioning):
        eng = get_engine(conditioning.device, **_DEFAULT_CFG)
        return eng.note_release(conditioning, np.float32(self.release_duration) * np.float32(self.frame_rate))


class InharmonicityNetwork:
    """sub_modules.py:611-701."""

    def __init__(self, model_specific_weight, slopes, offsets, slopes_modifier, offsets_modifier):
        self.model_specific_weight = model_specific_weight
        self.slopes, self.offsets = slopes + slopes_modifier, offsets + offsets_modifier

    def __call__(self, extended_pitch, global_inharm=None):
        asym = self.slopes * (extended_pitch / MIDI_NORM + self.offsets)
        if global_inharm is not None:
            g = global_inharm * 10.0
            asym = asym + self.model_specific_weight * torch.cat([torch.zeros_like(g), g], dim=-1)
        return torch.exp(asym).sum(dim=-1, keepdim=True)


class Detuner:
    """sub_modules.py:903-943."""

    def __init__(self, layer, n_substrings=2, use_detune=True):
        self.layer, self.n_substrings, self.use_detune = layer, n_substrings, use_detune

    def __call__(self, extended_pitch, global_detuning=None):
        pitch = extended_pitch
        if self.use_detune:
            detuning = torch.tanh(self.layer(extended_pitch / MIDI_NORM))
            if global_detuning is not None:
                detuning = detuning + torch.tanh(global_detuning)
            pitch = extended_pitch + detuning
        return 440.0 * torch.exp2((pitch - 69.0) / 12.0)                     # ddsp.core.midi_to_hz


class MonophonicNetwork:
    """sub_modules.py:455-496 with the layers of dafx22.gin:73-88: Dense 128, GRU 192, Dense 192,
    Normalize, dense_out -> (amplitudes 1, harmonic_distribution H, magnitudes M)."""

    def __init__(self, dense1, gru, dense2, norm, dense_out, output_splits):
        self.dense1, self.gru, self.dense2, self.norm, self.dense_out = dense1, gru, dense2, norm, dense_out
        self.output_splits = output_splits

    def __call__(self, conditioning, extended_pitch, context):
        scale = torch.tensor([MIDI_NORM, 1.0], device=conditioning.device)
        x = torch.cat([extended_pitch / MIDI_NORM, conditioning / scale, context], dim=-1)
        y = self.dense_out(self.norm(self.dense2(self.gru(self.dense1(x)))))
        out, at = {}, 0
        for key, dim in self.output_splits:
            out[key] = y[..., at:at + dim]
            at += dim
        return out


class Parallelizer:
    """sub_modules.py:528-602: merges the polyphony axis into the batch (voice-major rows
    v * B + b) for the monophonic modules and hands the results back as stacked [P, B, F, C]
    tensors plus per-voice views ``key_i`` -- the layout the fused forward takes zero-copy."""

    def __init__(self, n_synths=16,
                 global_keys=('conditioning', 'context', 'global_inharm', 'global_detuning'),
                 mono_keys=('f0_hz', 'inharm_coef', 'amplitudes', 'harmonic_distribution', 'magnitudes')):
        self.n_synths, self.global_keys, self.mono_keys = n_synths, global_keys, mono_keys

    def parallelize(self, features):
        P = self.n_synths
        for k in self.global_keys:
            x = features[k]
            if x.dim() == 4:                                                 # [B, F, P, C] -> [P, B, F, C]
                x = x.permute(2, 0, 1, 3)
            else:
                x = x.unsqueeze(0).expand(P, *x.shape)
            self.batch_size = x.shape[1]
            features[k] = x.reshape(P * x.shape[1], *x.shape[2:])
        return features

    def unparallelize(self, features):
        P = self.n_synths
        for k in self.mono_keys:
            x = features[k].contiguous()
            x = x.reshape(P, self.batch_size, *x.shape[1:])
            features[k] = x
            for i in range(P):
                features[f'{k}_{i}'] = x[i]
        return features

    def __call__(self, features, parallelize=True):
        return self.parallelize(features) if parallelize else self.unparallelize(features)


class PianoModel:
    """piano_model.py:12-169, inference only: global features -> parallelize -> monophonic
    features -> unparallelize -> processor group."""

    def __init__(self, z_encoder, note_release, context_network, parallelizer, monophonic_network,
                 inharm_model, detuner, reverb_model, processor_group, device='cuda'):
        self.device = torch.device(device)
        self.z_encoder, self.note_release, self.context_network = z_encoder, note_release, context_network
        self.parallelizer, self.monophonic_network = parallelizer, monophonic_network
        self.inharm_model, self.detuner, self.reverb_model = inharm_model, detuner, reverb_model
        self.processor_group = processor_group

    @property
    def n_synths(self):
        return self.parallelizer.n_synths

    @property
    def sample_rate(self):
        return self.processor_group.processors[0].sample_rate

    @property
    def n_instruments(self):
        """Rows of the smallest per-instrument table (tables with one row serve every instrument)."""
        rows = [getattr(m, 'n_instruments', None) for m in (self.z_encoder, self.reverb_model)]
        if isinstance(self.context_network, FiLMContextNetwork):
            rows.append(self.context_network.piano_id_head.shape[0])
        if isinstance(self.inharm_model, JointParametricInharmTuning):
            rows += [e.shape[0] for e in self.inharm_model.e.values()]
        rows = [r for r in rows if r is not None and r > 1]
        return min(rows) if rows else None

    def compute_controls(self, features):
        """Everything before the processor group (piano_model.py:146-158)."""
        f = dict(features)
        dev = self.device
        ids, n = torch.as_tensor(f['piano_model']).reshape(-1), self.n_instruments
        if n is not None and ids.numel() and (int(ids.min()) < 0 or int(ids.max()) >= n):
            # an embedding lookup out of range is an InvalidArgumentError in the reference; on the
            # device it would be an assert that takes the CUDA context with it
            raise ValueError(f'piano_model ids must lie in [0, {n}), got {ids.tolist()}')
        f['conditioning'] = torch.as_tensor(f['conditioning'], dtype=torch.float32, device=dev)
        f['pedal'] = torch.as_tensor(f['pedal'], dtype=torch.float32, device=dev)
        f['piano_model'] = torch.as_tensor(f['piano_model'], device=dev).long().reshape(-1, 1)
        n_frames = f['conditioning'].shape[1]
        # compute_global_features :118-128
        if self.z_encoder is not None:
            f['z'], f['global_inharm'], f['global_detuning'] = self.z_encoder(f['piano_model'], n_frames)
            f['context'] = self.context_network(f['conditioning'], f['pedal'], f['z'])
        else:
            f['context'] = self.context_network(f['conditioning'], f['pedal'], f['piano_model'])
        if getattr(self.reverb_model, 'cache_ir', False):
            f['reverb_ir'] = self.reverb_model(f['piano_model'], key=tuple(ids.tolist()))
        else:
            f['reverb_ir'] = self.reverb_model(f['piano_model'])
        f = self.parallelizer(f, parallelize=True)
        # compute_monophonic_features :130-142
        f['extended_pitch'] = self.note_release(f['conditioning'])
        if self.detuner is not None:
            f['inharm_coef'] = self.inharm_model(f['extended_pitch'], f['global_inharm'])
            f['f0_hz'] = self.detuner(f['extended_pitch'], f['global_detuning'])
        else:
            f['f0_hz'], f['inharm_coef'] = self.inharm_model(f['extended_pitch'], f['piano_model'])
        f.update(self.monophonic_network(f['conditioning'], f['extended_pitch'], f['context']))
        return self.parallelizer(f, parallelize=False)

    def __call__(self, features, training=False):
        if training:
            raise ValueError('the B200 path is inference only (no losses, no gradients)')
        f = self.compute_controls(features)
        pg_out = self.processor_group(f, return_outputs_dict=True)          # piano_model.py:160
        outputs = pg_out['controls']
        outputs['audio_synth'] = pg_out['signal']
        return outputs

    def get_audio_from_outputs(self, outputs):
        return outputs['audio_synth']


def dafx22_model(checkpoint_prefix, device='cuda', sample_rate=16000, frame_rate=250, n_synths=16,
                 inference=True, norm_axes='time_channels', seed=0, reverb_decay_mask=False):
    """The model ``configs/dafx22.gin`` builds, restored from the shipped weights
    (``model_weights/dafx22/ckpt-0``; 16 kHz, 96 partials, 64 noise bands, 1.5 s reverb).

    ``inference`` is the gin macro ``%inference`` that ``synthesize_midi_file.py:53`` sets: in
    ``dafx22.gin`` it reaches ``inharm_synth.MultiInharmonic`` only (angular cumsum).  The gin never
    binds ``MultiInstrumentReverb.inference`` (constructor default False, sub_modules.py:300-337), so
    the reference's dafx22 audio uses the learnt impulse response UNMASKED; ``reverb_decay_mask=True``
    opts into the exponential decay mask of sub_modules.py:339-349."""
    device = torch.device(device)
    if isinstance(checkpoint_prefix, (Checkpoint, NpzWeights)):
        ck = checkpoint_prefix
    elif str(checkpoint_prefix).endswith('.npz'):
        ck = NpzWeights(checkpoint_prefix)
    else:
        ck = Checkpoint(checkpoint_prefix)
    t = lambda name: _t(ck.tensor(f'model/{name}/.ATTRIBUTES/VARIABLE_VALUE'), device)
    raw = lambda name: ck.tensor(f'model/{name}/.ATTRIBUTES/VARIABLE_VALUE')
    dense = lambda p, act=None: Dense(t(f'{p}/kernel'), t(f'{p}/bias'), act)
    gru = lambda p: GRU(raw(f'{p}/cell/kernel'), raw(f'{p}/cell/recurrent_kernel'), raw(f'{p}/cell/bias'), device)
    norm = lambda p: Normalize(t(f'{p}/scale'), t(f'{p}/shift'), norm_axes)
    cn, mn = 'context_network/model/layer_with_weights-', 'monophonic_network/model/layer_with_weights-'
    out_dim = raw('monophonic_network/dense_out/bias').shape[0]
    n_mags = 64
    splits = (('amplitudes', 1), ('harmonic_distribution', out_dim - 1 - n_mags), ('magnitudes', n_mags))
    additive = MultiInharmonic(frame_rate=frame_rate, sample_rate=sample_rate, inference=inference,
                               name='additive')
    noise = DynamicSizeFilteredNoise(frame_rate=frame_rate, sample_rate=sample_rate, name='noise',
                                     seed=seed)
    dag = polyphonic_dag(additive=additive, noise=noise, reverb=Reverb(trainable=False),
                         additive_controls=['amplitudes', 'harmonic_distribution', 'inharm_coef', 'f0_hz'],
                         noise_controls=['magnitudes'], reverb_controls=['reverb_ir'], n_synths=n_synths)
    return PianoModel(
        z_encoder=OneHotZEncoder(t('z_encoder/embedding/embeddings'),
                                 t('z_encoder/inharm_embedding/embeddings'),
                                 t('z_encoder/detune_embedding/embeddings')),
        note_release=NoteRelease(float(raw('note_release/layer/cell/release_duration')), frame_rate),
        context_network=ContextNetwork(dense(cn + '0', leaky_relu), gru(cn + '1'), norm(cn + '2'),
                                       dense('context_network/dense_out')),
        parallelizer=Parallelizer(n_synths),
        monophonic_network=MonophonicNetwork(dense(mn + '0', leaky_relu), gru(mn + '1'),
                                             dense(mn + '2', leaky_relu), norm(mn + '3'),
                                             dense('monophonic_network/dense_out'), splits),
        inharm_model=InharmonicityNetwork(*(t(f'inharm_model/{k}') for k in
                                            ('model_specific_weight', 'slopes', 'offsets',
                                             'slopes_modifier', 'offsets_modifier'))),
        detuner=Detuner(dense('detuner/layer')),
        reverb_model=MultiInstrumentReverb(t('reverb_model/reverb_dict/layer_with_weights-0/embeddings'),
                                           sample_rate=sample_rate, inference=reverb_decay_mask),
        processor_group=ProcessorGroup(dag=dag), device=device)


# ---- configs/maestro-v2.gin: the script default (synthesize_midi_file.py:14-17) ----------------

class LayerNormalization:
    """tf.keras.layers.LayerNormalization(): last axis, epsilon 1e-3."""

    def __init__(self, gamma, beta, eps=1e-3):
        self.gamma, self.beta, self.eps = gamma, beta, eps

    def __call__(self, x):
        return torch.nn.functional.layer_norm(x, (x.shape[-1],), self.gamma, self.beta, self.eps)


class FcStack:
    """ddsp.training.nn.FcStack(ch, layers): [Dense, LayerNormalization, leaky_relu] x layers."""

    def __init__(self, layers):
        self.layers = layers                       # [(Dense, LayerNormalization)]

    def __call__(self, x):
        for dense, norm in self.layers:
            x = leaky_relu(norm(dense(x)))
        return x


class FiLMContextNetwork:
    """sub_modules.py:97-180."""

    def __init__(self, conditioning_head, pedal_head, piano_id_head, main_dense0, main_gru, main_dense2,
                 main_norm, film_input_reshape, output_layer):
        self.conditioning_head, self.pedal_head, self.piano_id_head = conditioning_head, pedal_head, piano_id_head
        self.main_dense0, self.main_gru, self.main_dense2, self.main_norm = \
            main_dense0, main_gru, main_dense2, main_norm
        self.film_input_reshape, self.output_layer = film_input_reshape, output_layer

    def apply_film(self, features, piano_feat):
        film_coef, film_bias = self.film_input_reshape(piano_feat).chunk(2, dim=-1)
        return features * film_coef + film_bias

    def __call__(self, conditioning, pedal, piano_model):
        B, F = conditioning.shape[:2]
        scale = torch.tensor([MIDI_NORM, 1.0], device=conditioning.device)
        conditioning_feat = self.conditioning_head((conditioning / scale).reshape(B, F, -1))
        pedal_feat = self.pedal_head(pedal)
        piano_feat = self.piano_id_head[piano_model.reshape(-1)][:, None, :]
        x = torch.cat([conditioning_feat, pedal_feat], dim=-1)
        x = leaky_relu(self.main_norm(self.main_dense2(self.main_gru(self.main_dense0(x)))))
        return self.output_layer(self.apply_film(x, piano_feat))


class MonophonicDeepNetwork:
    """sub_modules.py:499-525."""

    def __init__(self, input_stacks, rnn, out_stack, dense_out, output_splits):
        self.input_stacks, self.rnn, self.out_stack, self.dense_out = input_stacks, rnn, out_stack, dense_out
        self.output_splits = output_splits

    def __call__(self, conditioning, extended_pitch, context):
        scale = torch.tensor([MIDI_NORM, 1.0], device=conditioning.device)
        a = self.input_stacks[0](extended_pitch / MIDI_NORM)
        b = self.input_stacks[1](conditioning / scale)
        c = self.input_stacks[2](context)
        x = self.rnn(torch.cat([a, b, c], dim=-1))
        y = self.dense_out(self.out_stack(torch.cat([a, b, c, x], dim=-1)))
        out, at = {}, 0
        for key, dim in self.output_splits:
            out[key] = y[..., at:at + dim]
            at += dim
        return out


class JointParametricInharmTuning:
    """sub_modules.py:763-876 (Rigaud et al., DAFx-11): inharmonicity and octave-stretched tuning
    along the tessitura from seven per-instrument parameters."""

    def __init__(self, **embeddings):
        self.e = embeddings                        # alpha_b, beta_b, alpha_t, beta_t, pitch_ref, K, alpha

    def _p(self, name, piano_model):
        return self.e[name][piano_model.reshape(-1)][:, None, :]

    def get_inharm(self, pitch, pm):
        return torch.exp(self._p('alpha_b', pm) * pitch + self._p('beta_b', pm)) + \
            torch.exp(self._p('alpha_t', pm) * pitch + self._p('beta_t', pm))

    def get_deviation_from_ET(self, pitch, pm):
        hz = lambda n: 440.0 * torch.exp2((n - 69.0) / 12.0)
        ref = self._p('pitch_ref', pm)
        ratio = hz(pitch) / hz(ref)
        rho = 1.0 + self._p('K', pm) * (1.0 - torch.tanh((pitch - ref) / self._p('alpha', pm))) / 2.0
        detuning = 1.0 + self.get_inharm(ref, pm) * (ratio * rho) ** 2
        detuning = detuning / (1.0 + self.get_inharm(pitch, pm) * rho ** 2)
        return torch.sqrt(detuning)

    def __call__(self, extended_pitch, piano_model):
        inharm_coef = self.get_inharm(extended_pitch, piano_model)
        f0_hz = 440.0 * torch.exp2((extended_pitch - 69.0) / 12.0) * \
            self.get_deviation_from_ET(extended_pitch, piano_model)
        return f0_hz, inharm_coef


class MultiInstrumentFeedbackDelayReverb:
    """sub_modules.py:368-446: per-instrument FDN parameters (embeddings) -> ``reverb_ir`` through
    ``FeedbackDelayNetwork.get_ir`` (the CUDA IR generator of csrc/fdn.cuh)."""

    def __init__(self, embeddings, sample_rate=24000, cache_ir=False):
        from .processors import FeedbackDelayNetwork
        self.e = embeddings
        self.n_instruments = embeddings['_input_gain'].shape[0]
        self.reverb_model = FeedbackDelayNetwork(trainable=False, sampling_rate=float(sample_rate))
        self.reverb_model.build(None)
        # The reference evaluates the network on every call (its parameters are being trained).  With frozen
        # weights the response of an instrument never changes: cache_ir keeps the last few by instrument ids
        # (0.4 ms of double-precision inverse DFT per call otherwise).  Clear `ir_cache` after changing `e`.
        self.cache_ir = cache_ir
        self.ir_cache = {}

    def __call__(self, piano_model, key=None):
        if self.cache_ir and key is not None and key in self.ir_cache:
            return self.ir_cache[key]
        ir = self._compute(piano_model)
        if self.cache_ir and key is not None:
            if len(self.ir_cache) >= 16:
                self.ir_cache.pop(next(iter(self.ir_cache)))
            self.ir_cache[key] = ir
        return ir

    def _compute(self, piano_model):
        idx = piano_model.reshape(-1)
        if self.n_instruments == 1:
            idx = torch.zeros_like(idx)
        split = lambda x: torch.stack(x.chunk(4, dim=-1), dim=-1)            # reshape_embedding :427-429
        g = lambda name: self.e[name][idx]
        return self.reverb_model.get_ir(
            g('_input_gain'), g('_output_gain'), split(g('_gain_allpass')).contiguous(),
            split(g('_delays_allpass')).contiguous(), torch.relu(g('_time_rev_0_sec')),
            torch.sigmoid(g('_alpha_tone')), g('_early_ir'))


def maestro_v2_model(checkpoint_prefix, device='cuda', sample_rate=24000, frame_rate=250, n_synths=16,
                     inference=True, seed=0, cache_reverb_ir=None):
    """The model ``configs/maestro-v2.gin`` builds, restored from the shipped weights
    (``model_weights/v2/ckpt-225000``; 24 kHz, 128 partials, 96 noise bands, one string per note,
    2 s feedback-delay-network reverb).  ``cache_reverb_ir`` (default: ``inference``): compute the
    network's impulse response once per set of instrument ids instead of on every call."""
    device = torch.device(device)
    if isinstance(checkpoint_prefix, (Checkpoint, NpzWeights)):
        ck = checkpoint_prefix
    elif str(checkpoint_prefix).endswith('.npz'):
        ck = NpzWeights(checkpoint_prefix)
    else:
        ck = Checkpoint(checkpoint_prefix)
    raw = lambda name: ck.tensor(f'model/{name}/.ATTRIBUTES/VARIABLE_VALUE')
    t = lambda name: _t(raw(name), device)
    dense = lambda p, act=None: Dense(t(f'{p}/kernel'), t(f'{p}/bias'), act)
    gru = lambda p: GRU(raw(f'{p}/cell/kernel'), raw(f'{p}/cell/recurrent_kernel'), raw(f'{p}/cell/bias'), device)
    ln = lambda p: LayerNormalization(t(f'{p}/gamma'), t(f'{p}/beta'))
    lw = 'layer_with_weights-'
    stack = lambda p, n: FcStack([(dense(f'{p}/{lw}{i}/{lw}0'), ln(f'{p}/{lw}{i}/{lw}1')) for i in range(n)])
    cn, mn = 'context_network', 'monophonic_network'
    out_dim = raw(f'{mn}/dense_out/bias').shape[0]
    n_mags = 96
    splits = (('amplitudes', 1), ('harmonic_distribution', out_dim - 1 - n_mags), ('magnitudes', n_mags))
    additive = MultiInharmonic(frame_rate=frame_rate, sample_rate=sample_rate, inference=inference,
                               name='additive')
    noise = DynamicSizeFilteredNoise(frame_rate=frame_rate, sample_rate=sample_rate, name='noise',
                                     seed=seed)
    dag = polyphonic_dag(additive=additive, noise=noise, reverb=Reverb(trainable=False),
                         additive_controls=['amplitudes', 'harmonic_distribution', 'inharm_coef', 'f0_hz'],
                         noise_controls=['magnitudes'], reverb_controls=['reverb_ir'], n_synths=n_synths)
    return PianoModel(
        z_encoder=None,
        note_release=NoteRelease(float(raw('note_release/layer/cell/release_duration')), frame_rate),
        context_network=FiLMContextNetwork(
            stack(f'{cn}/conditioning_head', 2), stack(f'{cn}/pedal_head', 2),
            t(f'{cn}/piano_id_head/embeddings'), dense(f'{cn}/main_model/{lw}0', leaky_relu),
            gru(f'{cn}/main_model/{lw}1'), dense(f'{cn}/main_model/{lw}2'), ln(f'{cn}/main_model/{lw}3'),
            dense(f'{cn}/film_input_reshape'), stack(f'{cn}/output_layer', 2)),
        parallelizer=Parallelizer(n_synths, global_keys=('conditioning', 'context', 'piano_model')),
        monophonic_network=MonophonicDeepNetwork(
            [stack(f'{mn}/input_stacks/{i}', 3) for i in range(3)], gru(f'{mn}/model/{lw}0/rnn'),
            stack(f'{mn}/out_stack', 3), dense(f'{mn}/dense_out'), splits),
        inharm_model=JointParametricInharmTuning(
            **{k: t(f'inharm_model/{k}/embeddings') for k in
               ('alpha_b', 'beta_b', 'alpha_t', 'beta_t', 'pitch_ref', 'K', 'alpha')}),
        detuner=None,
        reverb_model=MultiInstrumentFeedbackDelayReverb(
            {k: t(f'reverb_model/{k}/embeddings') for k in
             ('_input_gain', '_output_gain', '_gain_allpass', '_delays_allpass', '_time_rev_0_sec',
              '_alpha_tone', '_early_ir')}, sample_rate=sample_rate, cache_ir=bool(inference if cache_reverb_ir is None else cache_reverb_ir)),
        processor_group=ProcessorGroup(dag=dag), device=device)
