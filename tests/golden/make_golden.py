#!/usr/bin/env python
"""Generate tests/golden/*.npz by EXECUTING the reference's own Python for the hot path.

Runs in the build container only (needs /root/reference, which does not exist on the
GPU box).  The reference's ``ddsp_piano/modules/{inharm_synth,filtered_noise_synth,
polyphonic_dag}.py`` are loaded unmodified by file path; ``tensorflow``/``gin``/``ddsp``
resolve to the NumPy stand-ins under ``oracle/tf_shim`` (TensorFlow and ddsp are not
installable here -- see oracle/__init__.py for what that does and does not pin).

    python tests/golden/make_golden.py            # rewrites tests/golden/*.npz
"""
import importlib.util
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get('DDSP_PIANO_REFERENCE', '/root/reference')


def load_reference_modules():
    sys.path.insert(0, os.path.join(ROOT, 'oracle', 'tf_shim'))
    sys.path.insert(0, ROOT)
    for pkg in ('ddsp_piano', 'ddsp_piano.modules'):
        m = types.ModuleType(pkg)
        m.__path__ = []
        sys.modules[pkg] = m
    mods = {}
    for name in ('inharm_synth', 'filtered_noise_synth', 'polyphonic_dag'):
        full = f'ddsp_piano.modules.{name}'
        spec = importlib.util.spec_from_file_location(
            full, os.path.join(REF, 'ddsp_piano', 'modules', name + '.py'))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[full] = mod
        spec.loader.exec_module(mod)
        mods[name] = mod
    return mods


def midi_hz(m):
    return 440.0 * 2.0 ** ((np.asarray(m, np.float64) - 69.0) / 12.0)


def make_voice_inputs(rng, B, F, H, S, M, onsets=True, silent=False):
    """Pre-get_controls tensors for one voice (BASELINE.md section 3 distributions)."""
    f0 = np.empty([B, F, S], np.float32)
    for b in range(B):
        k = 0
        while k < F:
            seg = int(rng.integers(4, 9)) if onsets else F
            if silent or (onsets and rng.random() < 0.25):
                hz = 8.1758                     # MIDI pitch 0: gated by min_frequency
            else:
                hz = float(midi_hz(rng.integers(21, 109)))
            for s in range(S):
                f0[b, k:k + seg, s] = hz * (1.0 + 1e-3 * s)
            k += seg
    return {
        'amplitudes': rng.standard_normal([B, F, 1]).astype(np.float32),
        'harmonic_distribution': rng.standard_normal([B, F, H]).astype(np.float32),
        'inharm_coef': rng.uniform(1e-4, 1e-3, [B, F, 1]).astype(np.float32),
        'f0_hz': f0,
        'magnitudes': rng.standard_normal([B, F, M]).astype(np.float32),
    }


def main():
    mods = load_reference_modules()
    import tensorflow as tf_shim           # the stand-in
    from ddsp import processors, effects   # the stand-in
    inh, fns, dag = mods['inharm_synth'], mods['filtered_noise_synth'], mods['polyphonic_dag']
    rng = np.random.default_rng(20221017)

    def save(name, **arrays):
        path = os.path.join(HERE, name + '.npz')
        np.savez_compressed(path, **arrays)
        print(f'{name}: {os.path.getsize(path) / 1024:.1f} KiB')

    # ---- additive: MultiInharmonic.get_controls + get_signal ------------------------
    add_cases = {
        # name: (sr, F, B, H, S, ctor kwargs)
        'additive_24k_inference': (24000, 25, 2, 96, 2, dict(inference=True)),
        'additive_16k_training': (16000, 20, 2, 32, 2, dict(inference=False)),
        'additive_48k_h128': (48000, 12, 1, 128, 2, dict(inference=True)),
        'additive_16k_exp_tanh_prenorm': (16000, 20, 1, 48, 3, dict(
            inference=True, scale_fn=inh.exp_tanh, normalize_after_nyquist_cut=False)),
        'additive_24k_single_string': (24000, 12, 1, 64, 1, dict(inference=True)),
    }
    for name, (sr, F, B, H, S, kw) in add_cases.items():
        x = make_voice_inputs(rng, B, F, H, S, 8)
        synth = inh.MultiInharmonic(frame_rate=250, sample_rate=sr, name='additive', **kw)
        ctl = synth.get_controls(x['amplitudes'].copy(), x['harmonic_distribution'].copy(),
                                 x['inharm_coef'].copy(), x['f0_hz'].copy())
        sig = synth.get_signal(**{k: v.copy() for k, v in ctl.items()})
        assert sig.dtype == np.float32 and sig.shape == (B, F * (sr // 250))
        save(name, sample_rate=sr, inference=bool(kw.get('inference', False)),
             scale_fn='exp_tanh' if 'scale_fn' in kw else 'exp_sigmoid',
             normalize_after_nyquist_cut=kw.get('normalize_after_nyquist_cut', True),
             in_amplitudes=x['amplitudes'], in_harmonic_distribution=x['harmonic_distribution'],
             in_inharm_coef=x['inharm_coef'], in_f0_hz=x['f0_hz'],
             ctl_amplitudes=ctl['amplitudes'],
             ctl_harmonic_distribution=ctl['harmonic_distribution'],
             ctl_harmonic_shifts=ctl['harmonic_shifts'], ctl_f0_hz=ctl['f0_hz'],
             signal=sig)

    # ---- noise: DynamicSizeFilteredNoise with injected uniform noise ----------------
    for name, (sr, F, B, M) in {'noise_24k_m64': (24000, 25, 2, 64),
                                'noise_48k_m96': (48000, 10, 1, 96),
                                'noise_16k_m64': (16000, 20, 1, 64)}.items():
        mags = rng.standard_normal([B, F, M]).astype(np.float32) * 2.0 + 3.0
        noise = rng.uniform(-1, 1, [B, F * (sr // 250)]).astype(np.float32)
        synth = fns.DynamicSizeFilteredNoise(frame_rate=250, sample_rate=sr, name='noise')
        ctl = synth.get_controls(mags.copy())
        tf_shim.push_noise(noise)
        sig = synth.get_signal(**ctl)
        assert sig.dtype == np.float32 and sig.shape == noise.shape
        save(name, sample_rate=sr, in_magnitudes=mags, noise=noise,
             ctl_magnitudes=ctl['magnitudes'], signal=sig)

    # ---- reverb: ddsp.effects.Reverb (stand-in; no reference source exists) ---------
    for name, (N, L, B, add_dry) in {'reverb_n2400_l1000': (2400, 1000, 2, True),
                                     'reverb_n2400_l2400_wet': (2400, 2400, 1, False)}.items():
        audio = rng.standard_normal([B, N]).astype(np.float32) * 0.1
        t = np.arange(L) / L
        ir = (rng.standard_normal([B, L]) * np.exp(-6 * t) * 1e-2).astype(np.float32)
        rv = effects.Reverb(trainable=False, add_dry=add_dry)
        sig = rv(audio, ir)
        save(name, add_dry=add_dry, audio=audio, ir=ir, signal=sig)

    # ---- the DAG: reference polyphonic_dag through ProcessorGroup -------------------
    sr, F, B, H, S, M, P, L = 24000, 25, 2, 96, 2, 64, 3, 1500
    feats, noises = {}, []
    for v in range(P):
        x = make_voice_inputs(rng, B, F, H, S, M, silent=(v == 2))
        for k, a in x.items():
            feats[f'{k}_{v}'] = a
        noises.append(rng.uniform(-1, 1, [B, F * (sr // 250)]).astype(np.float32))
    t = np.arange(L) / L
    feats['reverb_ir'] = (rng.standard_normal([B, L]) * np.exp(-6 * t) * 1e-2).astype(np.float32)
    nodes = dag.polyphonic_dag(
        additive=inh.MultiInharmonic(frame_rate=250, sample_rate=sr, inference=True,
                                     name='additive'),
        additive_controls=['amplitudes', 'harmonic_distribution', 'inharm_coef', 'f0_hz'],
        noise=fns.DynamicSizeFilteredNoise(frame_rate=250, sample_rate=sr, name='noise'),
        noise_controls=['magnitudes'],
        reverb=effects.Reverb(trainable=False), reverb_controls=['reverb_ir'], n_synths=P)
    group = processors.ProcessorGroup(dag=nodes)
    for n in noises:
        tf_shim.push_noise(n)
    out = group({k: v.copy() for k, v in feats.items()}, return_outputs_dict=True)
    arrays = {f'in_{k}': v for k, v in feats.items()}
    for v in range(P):
        arrays[f'noise_{v}'] = noises[v]
    save('dag_24k_p3', sample_rate=sr, n_synths=P,
         node_names=np.array([n[0].name for n in nodes]),
         dry=out['controls']['add']['signal'], signal=out['signal'],
         last_additive=out['controls']['additive']['signal'],
         last_noise=out['controls']['noise']['signal'], **arrays)


if __name__ == '__main__':
    main()


def make_ir_fixture():
    """Row 0 of the reverb impulse-response embedding of the SHIPPED dafx22 checkpoint
    (ddsp_piano/model_weights/dafx22/ckpt-0.data-00000-of-00001: tensor
    model/reverb_model/reverb_dict/layer_with_weights-0/embeddings [10, 24000] float32 LE at byte
    offset 308892) -- a real IR for the reverb parity tests.

        python -c "import sys; sys.path.insert(0, 'tests/golden'); import make_golden as m; m.make_ir_fixture()"
    """
    path = os.path.join(REF, 'ddsp_piano', 'model_weights', 'dafx22', 'ckpt-0.data-00000-of-00001')
    with open(path, 'rb') as f:
        f.seek(308892)
        ir = np.frombuffer(f.read(10 * 24000 * 4), dtype='<f4').reshape(10, 24000)
    assert np.isfinite(ir).all() and abs(float(ir[0, 1]) - 3.1801593) < 1e-6
    out = os.path.join(HERE, 'dafx22_reverb_ir_row0.npz')
    np.savez_compressed(out, ir=ir[0].astype(np.float32), sample_rate=16000, piano_model=0)
    print(f'dafx22_reverb_ir_row0: {os.path.getsize(out) / 1024:.1f} KiB')


def make_fdn():
    """tests/golden/fdn_*.npz: the reference's modules/fdn_reverb.py executed over the stand-in.

        python -c "import sys; sys.path.insert(0, 'tests/golden'); import make_golden as m; m.make_fdn()"
    """
    load_reference_modules()
    spec = importlib.util.spec_from_file_location(
        'ddsp_piano.modules.fdn_reverb', os.path.join(REF, 'ddsp_piano', 'modules', 'fdn_reverb.py'))
    mod = importlib.util.module_from_spec(spec)
    sys.modules['ddsp_piano.modules.fdn_reverb'] = mod
    spec.loader.exec_module(mod)
    rng = np.random.default_rng(20230928)
    for name, sr, n_audio in (('fdn_sr2000', 2000.0, 1500), ('fdn_sr8000', 8000.0, 5000)):
        fdn = mod.FeedbackDelayNetwork(trainable=False, sampling_rate=sr)
        fdn.build(None)
        # the initialisers of MultiInstrumentFeedbackDelayReverb (sub_modules.py:384-411) and the
        # activations of its call (:431-438)
        p = dict(
            input_gain=rng.normal(0.25, 0.1, 8).astype(np.float32),
            output_gain=rng.normal(0.25, 0.1, 8).astype(np.float32),
            gain_allpass=rng.normal(0.25, 0.1, [8, 4]).astype(np.float32),
            delays_allpass=rng.normal(400, 60, [8, 4]).astype(np.float32),
            time_rev_0_sec=np.maximum(rng.normal(2, 0.5, [1]), 0).astype(np.float32),
            alpha_tone=(1 / (1 + np.exp(-rng.normal(0, 0.1, [1])))).astype(np.float32),
            early_ir=rng.normal(0, 0.1, [200]).astype(np.float32))
        audio = (rng.standard_normal([2, n_audio]) * 0.1).astype(np.float32)
        ctl = fdn.get_controls(audio_dry=audio, **{k: v.copy() for k, v in p.items()})
        ir = ctl['ir']
        assert ir.dtype == np.float32 and ir.shape == (int(2 * sr),)
        sig = fdn.get_signal(ctl['audio'], ir)
        path = os.path.join(HERE, name + '.npz')
        np.savez_compressed(path, sampling_rate=sr, audio=audio, ir=ir, signal=sig, **p)
        print(f'{name}: {os.path.getsize(path) / 1024:.1f} KiB')


def make_fdn6():
    """tests/golden/fdn6_sr4000.npz: the 6-line network of configs/ENSTDkCl-*.gin:118-122 -- the reference's
    modules/fdn_reverb.py executed over the stand-in with 6 delay values / gains / allpass rows (its
    variables' shapes, drawn from its initialisers, fdn_reverb.py:121-174), then used as the DAG's last
    processor: get_controls(audio_dry) -> get_signal.

        python -c "import sys; sys.path.insert(0, 'tests/golden'); import make_golden as m; m.make_fdn6()"
    """
    load_reference_modules()
    spec = importlib.util.spec_from_file_location(
        'ddsp_piano.modules.fdn_reverb', os.path.join(REF, 'ddsp_piano', 'modules', 'fdn_reverb.py'))
    mod = importlib.util.module_from_spec(spec)
    sys.modules['ddsp_piano.modules.fdn_reverb'] = mod
    spec.loader.exec_module(mod)
    rng = np.random.default_rng(20240117)
    sr, n_audio, D = 4000.0, 3000, 6
    delay_values = rng.normal(400, 60, D).astype(np.float32)           # "Delay values" (delay_trainable)
    fdn = mod.FeedbackDelayNetwork(trainable=False, sampling_rate=sr, delay_lines=D, delay_values=delay_values)
    fdn.build(None)
    assert tuple(fdn.mixing_matrix.shape) == (D, D)
    p = dict(
        input_gain=rng.normal(0.25, 0.1, D).astype(np.float32),
        output_gain=rng.normal(0.25, 0.1, D).astype(np.float32),
        gain_allpass=rng.normal(0.25, 0.1, [D, 4]).astype(np.float32),
        delays_allpass=rng.normal(400, 60, [D, 4]).astype(np.float32),
        time_rev_0_sec=np.maximum(rng.normal(2, 0.5, [1]), 0).astype(np.float32),
        alpha_tone=(1 / (1 + np.exp(-rng.normal(0, 0.1, [1])))).astype(np.float32),
        early_ir=rng.normal(0, 0.1, [200]).astype(np.float32))
    audio = (rng.standard_normal([2, n_audio]) * 0.1).astype(np.float32)
    ctl = fdn.get_controls(audio_dry=audio, **{k: v.copy() for k, v in p.items()})
    ir = ctl['ir']
    assert ir.dtype == np.float32 and ir.shape == (int(2 * sr),)
    sig = fdn.get_signal(ctl['audio'], ir)
    path = os.path.join(HERE, 'fdn6_sr4000.npz')
    np.savez_compressed(path, sampling_rate=sr, audio=audio, ir=ir, signal=sig, delay_values=delay_values, **p)
    print(f'fdn6_sr4000: {os.path.getsize(path) / 1024:.1f} KiB')


def make_checkpoint_fixtures():
    """Small fixtures cut from the SHIPPED checkpoints with ddsp_piano_b200/checkpoint.py (no TF):
    * dafx22_ckpt-0.index         -- verbatim copy of model_weights/dafx22/ckpt-0.index (2.5 KB), for
                                     the parser test;
    * v2_fdn_params_piano0.npz    -- the feedback-delay-network parameters of instrument 0 in
                                     model_weights/v2/ckpt-225000 (activations applied as in
                                     sub_modules.py:431-438), for the FDN kernels.

        python -c "import sys; sys.path.insert(0, 'tests/golden'); import make_golden as m; m.make_checkpoint_fixtures()"
    """
    import shutil
    sys.path.insert(0, ROOT)
    from ddsp_piano_b200.checkpoint import Checkpoint
    weights = os.path.join(REF, 'ddsp_piano', 'model_weights')
    shutil.copyfile(os.path.join(weights, 'dafx22', 'ckpt-0.index'),
                    os.path.join(HERE, 'dafx22_ckpt-0.index'))
    p = Checkpoint(os.path.join(weights, 'v2', 'ckpt-225000')).fdn_parameters(0)
    np.savez_compressed(os.path.join(HERE, 'v2_fdn_params_piano0.npz'), sampling_rate=24000.0, **p)


def make_midi_conditioning():
    """tests/golden/midi_conditioning.npz: the reference's OWN utils/midi_encoders.py (pure NumPy,
    imported unmodified -- no stand-in involved) run on synthetic pianorolls.

        python -c "import sys; sys.path.insert(0, 'tests/golden'); import make_golden as m; m.make_midi_conditioning()"
    """
    spec = importlib.util.spec_from_file_location(
        'ref_midi_encoders', os.path.join(REF, 'ddsp_piano', 'utils', 'midi_encoders.py'))
    enc = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(enc)
    rng = np.random.default_rng(88)

    def random_roll(n_frames, n_notes, max_len):
        roll = np.zeros([n_frames, 88, 2], np.float32)
        for _ in range(n_notes):
            p = int(rng.integers(0, 88))
            t0 = int(rng.integers(0, n_frames - 2))
            t1 = min(n_frames, t0 + int(rng.integers(2, max_len)))
            if roll[max(t0 - 1, 0):t1 + 1, p, 0].any():
                continue                                   # keep notes of one pitch apart
            roll[t0:t1, p, 0] = 1.0
            roll[t0, p, 1] = float(rng.integers(1, 128)) / 127.0    # onset velocity, onset frame only
        return roll

    cases = {'sparse': random_roll(500, 60, 80), 'dense': random_roll(400, 400, 120),
             'chords': random_roll(300, 150, 200)}
    # more simultaneous notes than channels for a while
    over = np.zeros([120, 88, 2], np.float32)
    over[10:90, 20:44, 0] = 1.0
    over[10, 20:44, 1] = 0.5
    over[50:110, 60:64, 0] = 1.0
    over[50, 60:64, 1] = 0.8
    cases['overflow'] = over
    arrays = {}
    for name, roll in cases.items():
        for n_synths in (16, 4):
            cond, poly = enc.MIDIRoll2Conditioning(n_synths)(roll.copy())
            arrays[f'{name}_cond{n_synths}'] = cond.astype(np.float32)
            arrays[f'{name}_poly'] = poly.astype(np.float32)
        arrays[f'{name}_roll'] = roll
    path = os.path.join(HERE, 'midi_conditioning.npz')
    np.savez_compressed(path, **arrays)
    print(f'midi_conditioning: {os.path.getsize(path) / 1024:.1f} KiB')
