"""CPU oracle (TEST INFRASTRUCTURE ONLY) for the control-rate graph of ``configs/dafx22.gin``
(SURVEY 8f rank 1): a numpy float32 restatement of what ``PianoModel.call``
(reference ``ddsp_piano/modules/piano_model.py:146-169``) computes between the MIDI conditioning
and the processor group -- the control tensors the synthesis kernels consume.

Only ``tests/`` may import this module; the product path is ``ddsp_piano_b200/model.py``.

Parity status: the reference's own code (``modules/sub_modules.py``: OneHotZEncoder :183-251,
ContextNetwork :18-65, NoteRelease / F0ProcessorCell :1114-1188, InharmonicityNetwork :611-701,
Detuner :903-943, MonophonicNetwork :455-496, Parallelizer :528-602) is restated line by line.
The third-party layers underneath -- ``tf.keras.layers.{Dense, GRU, Embedding}``,
``tf.nn.leaky_relu``, ``ddsp.training.nn.{Normalize, OutputSplitsLayer}``,
``ddsp.core.{midi_to_hz, resample}`` -- are NOT in the container (TensorFlow/ddsp absent, no
network), so their published semantics are restated here and this layer is **parity unpinned**:

* Dense: ``act(x @ kernel + bias)``; ``leaky_relu`` with TensorFlow's default ``alpha = 0.2``.
* GRU (TF2 default ``reset_after=True``, sigmoid/tanh): kernel ``[in, 3u]`` and recurrent kernel
  ``[u, 3u]`` in gate order (z, r, h), bias ``[2, 3u]`` = (input bias, recurrent bias):
  ``z = s(x Wz + bz + h Uz + cz)``, ``r = s(x Wr + br + h Ur + cr)``,
  ``hh = tanh(x Wh + bh + r * (h Uh + ch))``, ``h' = z * h + (1 - z) * hh``.  The checkpoint's bias
  shape [2, 3u] confirms ``reset_after``.
* ``nn.Normalize('layer')``: learnt ``scale``/``shift`` of shape [1, 1, 1, C] (as in the checkpoint)
  applied to ``(x - mean) / sqrt(var + 1e-5)``, moments over every axis but the batch (time and
  channels); ``norm_axes='channels'`` switches to per-frame moments for comparison.
* ``nn.OutputSplitsLayer``: a final ``Dense(sum(dims))`` named ``dense_out`` followed by a split.
* ``midi_to_hz(n) = 440 * 2 ** ((n - 69) / 12)``; ``resample`` of a one-frame embedding to
  ``n_frames`` is a broadcast.
"""
import numpy as np

F32 = np.float32
MIDI_NORM = F32(128.0)


def leaky_relu(x, alpha=0.2):
    return np.where(x > 0, x, F32(alpha) * x).astype(F32)


def dense(x, kernel, bias, activation=None):
    y = (x.astype(F32) @ kernel.astype(F32) + bias.astype(F32)).astype(F32)
    return activation(y) if activation is not None else y


def _sigmoid(x):
    return (1.0 / (1.0 + np.exp(-x.astype(np.float64)))).astype(F32)


def gru(x, kernel, recurrent_kernel, bias):
    """Keras GRU, reset_after=True, return_sequences=True, zero initial state.  x: [R, T, in]."""
    R, T, _ = x.shape
    u = recurrent_kernel.shape[0]
    xw = (x.astype(F32) @ kernel.astype(F32) + bias[0].astype(F32)).astype(F32)     # [R, T, 3u]
    h = np.zeros([R, u], F32)
    out = np.empty([R, T, u], F32)
    for t in range(T):
        hu = (h @ recurrent_kernel.astype(F32) + bias[1].astype(F32)).astype(F32)
        z = _sigmoid(xw[:, t, :u] + hu[:, :u])
        r = _sigmoid(xw[:, t, u:2 * u] + hu[:, u:2 * u])
        hh = np.tanh(xw[:, t, 2 * u:] + r * hu[:, 2 * u:]).astype(F32)
        h = (z * h + (F32(1) - z) * hh).astype(F32)
        out[:, t] = h
    return out


def normalize(x, scale, shift, norm_axes='time_channels', eps=1e-5):
    """ddsp.training.nn.Normalize('layer') on [R, T, C]."""
    axes = (1, 2) if norm_axes == 'time_channels' else (2,)
    x64 = x.astype(np.float64)
    mean = x64.mean(axis=axes, keepdims=True)
    var = x64.var(axis=axes, keepdims=True)
    y = ((x64 - mean) / np.sqrt(var + eps)).astype(F32)
    return (y * scale.reshape(1, 1, -1).astype(F32) + shift.reshape(1, 1, -1).astype(F32)).astype(F32)


def midi_to_hz(notes):
    return (F32(440.0) * np.exp2((notes.astype(F32) - F32(69.0)) / F32(12.0))).astype(F32)


def note_release(active_pitch, release_duration, frame_rate=250):
    """sub_modules.py:1138-1171 unrolled over time.  active_pitch: [R, T, 1] -> [R, T, 1]."""
    R, T, _ = active_pitch.shape
    sat = lambda v, thr: np.minimum(np.maximum(v - thr, F32(0)), F32(1)).astype(F32)
    previous = np.zeros([R, 1], F32)
    steps = np.zeros([R, 1], F32)
    limit = F32(F32(release_duration) * F32(frame_rate))
    out = np.empty([R, T, 1], F32)
    for t in range(T):
        note = active_pitch[:, t].astype(F32)
        activity = sat(note, F32(0))                                            # :1151
        release_end = sat(steps, limit)                                         # :1154-1156
        y = (activity * note + (F32(1) - activity) * previous * (F32(1) - release_end)).astype(F32)
        steps = ((steps + F32(1)) * (F32(1) - activity) * (F32(1) - release_end)).astype(F32)
        previous = y
        out[:, t] = y
    return out


def inharmonicity(extended_pitch, global_inharm, w):
    """InharmonicityNetwork.call, sub_modules.py:667-701."""
    reduced = (extended_pitch.astype(F32) / MIDI_NORM).astype(F32)
    slopes = (w['slopes'] + w['slopes_modifier']).astype(F32)
    offsets = (w['offsets'] + w['offsets_modifier']).astype(F32)
    asym = (slopes * (reduced + offsets)).astype(F32)                           # [R, T, 2]
    if global_inharm is not None:
        g = (global_inharm.astype(F32) * F32(10.0)).astype(F32)
        g = np.concatenate([np.zeros_like(g), g], axis=-1)                      # bass bridge only
        asym = (asym + w['model_specific_weight'].astype(F32) * g).astype(F32)
    return np.exp(asym).sum(axis=-1, keepdims=True).astype(F32)


def detuner(extended_pitch, global_detuning, kernel, bias, use_detune=True):
    """Detuner.call, sub_modules.py:923-943 -> f0_hz [R, T, n_substrings]."""
    pitch = extended_pitch.astype(F32)
    if use_detune:
        det = np.tanh(dense(pitch / MIDI_NORM, kernel, bias)).astype(F32)
        if global_detuning is not None:
            det = (det + np.tanh(global_detuning.astype(F32))).astype(F32)
        pitch = (pitch + det).astype(F32)
    return midi_to_hz(pitch)


def control_graph(conditioning, pedal, piano_model, w, n_synths=None, norm_axes='time_channels',
                  use_detune=True, frame_rate=250):
    """conditioning [B, T, P, 2], pedal [B, T, 4], piano_model [B] int -> the stacked control
    tensors of the processor group: amplitudes [P, B, T, 1], harmonic_distribution [P, B, T, H],
    magnitudes [P, B, T, M], f0_hz [P, B, T, S], inharm_coef [P, B, T, 1] (+ intermediates)."""
    conditioning = conditioning.astype(F32)
    pedal = pedal.astype(F32)
    B, T, P, _ = conditioning.shape
    assert n_synths in (None, P)
    pm = np.asarray(piano_model).reshape(B).astype(np.int64)
    if w['z_embedding'].shape[0] == 1:                                          # :231-232
        pm = np.zeros_like(pm)
    # OneHotZEncoder.call :229-251 (one frame resampled to n_frames = a broadcast)
    z = np.repeat(w['z_embedding'][pm][:, None, :], T, axis=1).astype(F32)
    g_inh = np.repeat(w['z_inharm'][pm][:, None, :], T, axis=1).astype(F32)
    g_det = np.repeat(w['z_detune'][pm][:, None, :], T, axis=1).astype(F32)
    # ContextNetwork.compute_output :50-65 (normalize_pitch=False in dafx22.gin)
    x = np.concatenate([conditioning.reshape(B, T, 2 * P), pedal, z], axis=-1)
    x = dense(x, *w['context_dense'], activation=leaky_relu)
    x = gru(x, *w['context_gru'])
    x = normalize(x, *w['context_norm'], norm_axes=norm_axes)
    context = dense(x, *w['context_out'])                                       # [B, T, 32]
    # Parallelizer.parallelize :583-587: voice-major rows v * B + b
    par = lambda a: np.repeat(a[None], P, axis=0).reshape(P * B, T, a.shape[-1])
    cond_p = conditioning.transpose(2, 0, 1, 3).reshape(P * B, T, 2)
    context_p, g_inh_p, g_det_p = par(context), par(g_inh), par(g_det)
    # monophonic features, piano_model.py:132-144
    ext = note_release(cond_p[..., 0:1], w['release_duration'], frame_rate)
    inharm = inharmonicity(ext, g_inh_p, w['inharm'])
    f0 = detuner(ext, g_det_p, *w['detuner'], use_detune=use_detune)
    # MonophonicNetwork.compute_output :478-496
    x = np.concatenate([ext / MIDI_NORM, cond_p / np.array([MIDI_NORM, 1.0], F32), context_p],
                       axis=-1).astype(F32)
    x = dense(x, *w['mono_dense1'], activation=leaky_relu)
    x = gru(x, *w['mono_gru'])
    x = dense(x, *w['mono_dense2'], activation=leaky_relu)
    x = normalize(x, *w['mono_norm'], norm_axes=norm_axes)
    y = dense(x, *w['mono_out'])
    H, M = w['n_harmonics'], w['n_magnitudes']
    un = lambda a: a.reshape(P, B, T, a.shape[-1])                              # :576-596
    return dict(amplitudes=un(y[..., 0:1]), harmonic_distribution=un(y[..., 1:1 + H]),
                magnitudes=un(y[..., 1 + H:1 + H + M]), f0_hz=un(f0), inharm_coef=un(inharm),
                extended_pitch=un(ext), context=context)


def load_weights(ckpt, n_harmonics=96, n_magnitudes=64):
    """Checkpoint (ddsp_piano_b200.checkpoint.Checkpoint of model_weights/dafx22/ckpt-0) ->
    the dict ``control_graph`` takes; keys of SURVEY appendix B."""
    t = lambda name: ckpt.tensor(f'model/{name}/.ATTRIBUTES/VARIABLE_VALUE')
    pair = lambda prefix, a='kernel', b='bias': (t(f'{prefix}/{a}'), t(f'{prefix}/{b}'))
    gru_w = lambda prefix: (t(f'{prefix}/cell/kernel'), t(f'{prefix}/cell/recurrent_kernel'),
                            t(f'{prefix}/cell/bias'))
    return dict(
        z_embedding=t('z_encoder/embedding/embeddings'),
        z_inharm=t('z_encoder/inharm_embedding/embeddings'),
        z_detune=t('z_encoder/detune_embedding/embeddings'),
        context_dense=pair('context_network/model/layer_with_weights-0'),
        context_gru=gru_w('context_network/model/layer_with_weights-1'),
        context_norm=pair('context_network/model/layer_with_weights-2', 'scale', 'shift'),
        context_out=pair('context_network/dense_out'),
        mono_dense1=pair('monophonic_network/model/layer_with_weights-0'),
        mono_gru=gru_w('monophonic_network/model/layer_with_weights-1'),
        mono_dense2=pair('monophonic_network/model/layer_with_weights-2'),
        mono_norm=pair('monophonic_network/model/layer_with_weights-3', 'scale', 'shift'),
        mono_out=pair('monophonic_network/dense_out'),
        detuner=pair('detuner/layer'),
        inharm={k: t(f'inharm_model/{k}') for k in
                ('model_specific_weight', 'slopes', 'offsets', 'slopes_modifier', 'offsets_modifier')},
        release_duration=float(t('note_release/layer/cell/release_duration')),
        n_harmonics=n_harmonics, n_magnitudes=n_magnitudes)


# ------------------------------------------------------------------------------------------------
# configs/maestro-v2.gin (the script default, synthesize_midi_file.py:14-17): FiLM context network,
# deep monophonic network, joint parametric inharmonicity/tuning.  Additional third-party layers,
# restated and unpinned like the ones above:
# * ``ddsp.training.nn.FcStack(ch, layers)``: ``layers`` x [Dense(ch), LayerNormalization(),
#   leaky_relu] (checkpoint: kernel/bias + gamma/beta per layer).
# * ``tf.keras.layers.LayerNormalization``: moments over the last axis, epsilon 1e-3, gamma/beta.
# * ``nn.Rnn(ch, 'gru')`` = GRU(ch, return_sequences=True); ``nn.get_embedding`` = Embedding.
# ------------------------------------------------------------------------------------------------

def layer_norm(x, gamma, beta, eps=1e-3):
    x64 = x.astype(np.float64)
    mean = x64.mean(axis=-1, keepdims=True)
    var = x64.var(axis=-1, keepdims=True)
    return (((x64 - mean) / np.sqrt(var + eps)).astype(F32) * gamma.astype(F32) + beta.astype(F32)).astype(F32)


def fc_stack(x, layers):
    for kernel, bias, gamma, beta in layers:
        x = leaky_relu(layer_norm(dense(x, kernel, bias), gamma, beta))
    return x


def joint_inharm_tuning(extended_pitch, piano_model, w):
    """JointParametricInharmTuning.call, sub_modules.py:833-876.  extended_pitch [R, T, 1],
    piano_model [R] -> f0_hz [R, T, 1], inharm_coef [R, T, 1]."""
    e = lambda name: w[name][piano_model][:, None, :].astype(F32)             # [R, 1, 1]
    pitch = extended_pitch.astype(F32)

    def inharm(p):                                                            # :833-837
        return (np.exp(e('alpha_b') * p + e('beta_b')) + np.exp(e('alpha_t') * p + e('beta_t'))).astype(F32)

    ref_pitch = e('pitch_ref')
    ratio = (midi_to_hz(pitch) / midi_to_hz(ref_pitch)).astype(F32)           # :842
    rst = ((F32(1) - np.tanh((pitch - ref_pitch) / e('alpha'))) / F32(2)).astype(F32)   # :830-831
    rho = (F32(1) + e('K') * rst).astype(F32)                                 # :845-847
    det = (F32(1) + inharm(ref_pitch) * (ratio * rho) ** 2).astype(F32)       # :849
    det = (det / (F32(1) + inharm(pitch) * rho ** 2)).astype(F32)             # :850
    det = np.sqrt(det).astype(F32)
    return (midi_to_hz(pitch) * det).astype(F32), inharm(pitch)


def control_graph_v2(conditioning, pedal, piano_model, w, frame_rate=250):
    """maestro-v2: conditioning [B, T, P, 2], pedal [B, T, 4], piano_model [B] -> stacked controls
    (f0_hz has ONE string: [P, B, T, 1])."""
    conditioning = conditioning.astype(F32)
    pedal = pedal.astype(F32)
    B, T, P, _ = conditioning.shape
    pm = np.asarray(piano_model).reshape(B).astype(np.int64)
    scale = np.array([MIDI_NORM, 1.0], F32)
    # FiLMContextNetwork.call, sub_modules.py:153-180
    cond_feat = fc_stack((conditioning / scale).reshape(B, T, 2 * P), w['ctx_conditioning_head'])
    pedal_feat = fc_stack(pedal, w['ctx_pedal_head'])
    piano_feat = w['ctx_piano_id'][pm][:, None, :].astype(F32)                # [B, 1, 32]
    x = np.concatenate([cond_feat, pedal_feat], axis=-1)
    x = dense(x, *w['ctx_main_dense0'], activation=leaky_relu)
    x = gru(x, *w['ctx_main_gru'])
    x = leaky_relu(layer_norm(dense(x, *w['ctx_main_dense2']), *w['ctx_main_ln']))
    film = dense(piano_feat, *w['ctx_film'])                                  # :140-151
    half = film.shape[-1] // 2
    x = (x * film[..., :half] + film[..., half:]).astype(F32)
    context = fc_stack(x, w['ctx_output'])
    # Parallelizer, global_keys = (conditioning, context, piano_model)  (maestro-v2.gin:36-38)
    cond_p = conditioning.transpose(2, 0, 1, 3).reshape(P * B, T, 2)
    context_p = np.repeat(context[None], P, axis=0).reshape(P * B, T, -1)
    pm_p = np.tile(pm, P)
    ext = note_release(cond_p[..., 0:1], w['release_duration'], frame_rate)
    f0, inharm = joint_inharm_tuning(ext, pm_p, w['tuning'])
    # MonophonicDeepNetwork.compute_output, sub_modules.py:510-525
    a = fc_stack(ext / MIDI_NORM, w['mono_in0'])
    b = fc_stack(cond_p / scale, w['mono_in1'])
    c = fc_stack(context_p, w['mono_in2'])
    x = np.concatenate([a, b, c], axis=-1)
    x = gru(x, *w['mono_gru'])
    x = np.concatenate([a, b, c, x], axis=-1)
    y = dense(fc_stack(x, w['mono_out_stack']), *w['mono_out'])
    H, M = w['n_harmonics'], w['n_magnitudes']
    un = lambda t_: t_.reshape(P, B, T, t_.shape[-1])
    return dict(amplitudes=un(y[..., 0:1]), harmonic_distribution=un(y[..., 1:1 + H]),
                magnitudes=un(y[..., 1 + H:1 + H + M]), f0_hz=un(f0), inharm_coef=un(inharm),
                extended_pitch=un(ext), context=context)


def load_weights_v2(ckpt, n_harmonics=128, n_magnitudes=96):
    t = lambda name: ckpt.tensor(f'model/{name}/.ATTRIBUTES/VARIABLE_VALUE')
    pair = lambda prefix, a='kernel', b='bias': (t(f'{prefix}/{a}'), t(f'{prefix}/{b}'))

    def stack(prefix, n):
        return [(t(f'{prefix}/layer_with_weights-{i}/layer_with_weights-0/kernel'),
                 t(f'{prefix}/layer_with_weights-{i}/layer_with_weights-0/bias'),
                 t(f'{prefix}/layer_with_weights-{i}/layer_with_weights-1/gamma'),
                 t(f'{prefix}/layer_with_weights-{i}/layer_with_weights-1/beta')) for i in range(n)]

    gru_w = lambda prefix: (t(f'{prefix}/cell/kernel'), t(f'{prefix}/cell/recurrent_kernel'),
                            t(f'{prefix}/cell/bias'))
    cn, mn = 'context_network', 'monophonic_network'
    return dict(
        ctx_conditioning_head=stack(f'{cn}/conditioning_head', 2),
        ctx_pedal_head=stack(f'{cn}/pedal_head', 2),
        ctx_piano_id=t(f'{cn}/piano_id_head/embeddings'),
        ctx_main_dense0=pair(f'{cn}/main_model/layer_with_weights-0'),
        ctx_main_gru=gru_w(f'{cn}/main_model/layer_with_weights-1'),
        ctx_main_dense2=pair(f'{cn}/main_model/layer_with_weights-2'),
        ctx_main_ln=pair(f'{cn}/main_model/layer_with_weights-3', 'gamma', 'beta'),
        ctx_film=pair(f'{cn}/film_input_reshape'),
        ctx_output=stack(f'{cn}/output_layer', 2),
        mono_in0=stack(f'{mn}/input_stacks/0', 3), mono_in1=stack(f'{mn}/input_stacks/1', 3),
        mono_in2=stack(f'{mn}/input_stacks/2', 3),
        mono_gru=gru_w(f'{mn}/model/layer_with_weights-0/rnn'),
        mono_out_stack=stack(f'{mn}/out_stack', 3),
        mono_out=pair(f'{mn}/dense_out'),
        tuning={k: t(f'inharm_model/{k}/embeddings') for k in
                ('alpha_b', 'beta_b', 'alpha_t', 'beta_t', 'pitch_ref', 'K', 'alpha')},
        release_duration=float(t('note_release/layer/cell/release_duration')),
        n_harmonics=n_harmonics, n_magnitudes=n_magnitudes)
