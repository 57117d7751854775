"""CPU-side tests: the C-ABI library builds, loads and exports every symbol the header
declares (no compute calls -- there is no GPU here), and the host-side mirror of the
reference's operator API behaves like ddsp's (DAG format, key strings, pattern match,
error behaviour, no silent CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def lib_path():
    import __graft_entry__
    __graft_entry__.build()
    from ddsp_piano_b200 import _lib
    return _lib.LIB_PATH


def header_functions():
    text = open(os.path.join(ROOT, 'include', 'b200ddsp.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(b200ddsp_[a-z_0-9]+)\s*\(', text)))


def test_header_declares_the_documented_surface():
    names = header_functions()
    for must in ['b200ddsp_create', 'b200ddsp_destroy', 'b200ddsp_additive_controls',
                 'b200ddsp_additive_signal', 'b200ddsp_noise_controls', 'b200ddsp_noise_signal',
                 'b200ddsp_reverb', 'b200ddsp_forward_polyphonic',
                 'b200ddsp_forward_polyphonic_host', 'b200ddsp_last_error']:
        assert must in names


def test_library_exports_every_declared_symbol(lib_path):
    lib = ctypes.CDLL(lib_path)
    missing = [n for n in header_functions() if not hasattr(lib, n)]
    assert not missing, missing
    lib.b200ddsp_version.restype = ctypes.c_int
    assert lib.b200ddsp_version() == 100


def test_python_binding_covers_the_header(lib_path):
    from ddsp_piano_b200 import _lib
    assert sorted(_lib.EXPORTS) == header_functions()
    lib = _lib.load()
    for n in _lib.EXPORTS:
        assert getattr(lib, n).restype is not None or n == 'never'


def test_struct_layouts_match_the_header():
    from ddsp_piano_b200 import _lib
    text = open(os.path.join(ROOT, 'include', 'b200ddsp.h')).read()
    cfg = re.search(r'typedef struct \{(.*?)\} b200ddsp_config;', text, flags=re.S).group(1)
    cfg = re.sub(r'/\*.*?\*/', '', cfg, flags=re.S)
    fields = re.findall(r'\b(?:int|float)\s+([a-z_0-9]+)\s*;', cfg)
    assert fields == [f[0] for f in _lib.Config._fields_]
    voice = re.search(r'typedef struct \{(.*?)\} b200ddsp_voice;', text, flags=re.S).group(1)
    voice = re.sub(r'/\*.*?\*/', '', voice, flags=re.S)
    vfields = re.findall(r'const float\*\s+([a-z_0-9]+)\s*;', voice)
    assert vfields == [f[0] for f in _lib.Voice._fields_]
    assert ctypes.sizeof(_lib.Voice) == 6 * ctypes.sizeof(ctypes.c_void_p)
    def body(name):
        end = text.index('} ' + name + ';')
        return text[text.rindex('typedef struct {', 0, end) + len('typedef struct {'):end]

    link = body('b200ddsp_link')
    link = re.sub(r'/\*.*?\*/', '', link, flags=re.S)
    lfields = re.findall(r'\b(?:const float\*|float\*|unsigned long long\*?)\s+([a-z_0-9]+)\s*;', link)
    assert lfields == [f[0] for f in _lib.Link._fields_]
    assert ctypes.sizeof(_lib.Link) == 8 * 8
    span = body('b200ddsp_span')
    span = re.sub(r'/\*.*?\*/', '', span, flags=re.S)
    sfields = re.findall(r'\b(?:long long|int|b200ddsp_link)\s+([a-z_0-9]+)\s*;', span)
    assert sfields == [f[0] for f in _lib.Span._fields_]
    assert ctypes.sizeof(_lib.Span) == 4 * 8 + ctypes.sizeof(_lib.Link)


def test_create_fails_loudly_without_a_gpu(lib_path):
    """No CPU fallback: creating a handle without a CUDA device is an error, not a detour."""
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    from ddsp_piano_b200 import _lib
    lib = _lib.load()
    cfg = _lib.Config(sample_rate=24000, frame_rate=250, min_frequency=20.0, additive_scale_fn=0,
                      normalize_after_nyquist_cut=1, normalize_below_nyquist=1, inference=1,
                      noise_scale_fn=0, noise_initial_bias=-5.0, noise_window_size=257,
                      reverb_add_dry=1, n_noise_bands=64, fast_phase=0)
    h = ctypes.c_void_p()
    rc = lib.b200ddsp_create(ctypes.byref(cfg), ctypes.byref(h))
    assert rc == -4 and not h.value                      # B200DDSP_CUDA_ERROR
    assert b'cuda' in lib.b200ddsp_last_error(None).lower()
    import ddsp_piano_b200 as dp
    synth = dp.MultiInharmonic(sample_rate=24000, inference=True)
    z = torch.zeros(1, 4, 1)
    with pytest.raises(RuntimeError):
        synth.get_signal(z, torch.zeros(1, 4, 8), torch.zeros(1, 4, 8), z)


def test_product_does_not_import_the_oracle():
    """oracle/ is test infrastructure: nothing under the package may reference it."""
    pkg = os.path.join(ROOT, 'ddsp_piano_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                src = open(os.path.join(dirpath, f)).read()
                assert 'oracle' not in src, os.path.join(dirpath, f)
                assert '/root/reference' not in src


# ------------------------------- operator API (host logic) ------------------------------------

def test_polyphonic_dag_matches_reference_structure():
    """modules/polyphonic_dag.py:21-42: node order, key strings, shared processor objects."""
    import ddsp_piano_b200 as dp
    additive = dp.MultiInharmonic(sample_rate=24000, inference=True, name='additive')
    noise = dp.DynamicSizeFilteredNoise(sample_rate=24000, name='noise')
    reverb = dp.Reverb(trainable=False)
    dag = dp.polyphonic_dag(additive, noise, reverb,
                            additive_controls=['amplitudes', 'harmonic_distribution',
                                               'inharm_coef', 'f0_hz'],
                            noise_controls=['magnitudes'], reverb_controls=['reverb_ir'],
                            n_synths=3)
    assert len(dag) == 3 * 3 + 1
    assert dag[0] == (additive, ['amplitudes_0', 'harmonic_distribution_0', 'inharm_coef_0', 'f0_hz_0'])
    assert dag[1] == (noise, ['magnitudes_0'])
    assert dag[2][0].name == 'add' and dag[2][1] == ['noise/signal', 'additive/signal']
    assert dag[5][1] == ['add/signal', 'noise/signal', 'additive/signal']
    assert dag[3][0] is additive and dag[4][0] is noise and dag[5][0] is dag[2][0]
    assert dag[-1] == (reverb, ['add/signal', 'reverb_ir'])
    # default key names of the reference signature
    d2 = dp.polyphonic_dag(additive, noise, n_synths=1)
    assert d2[0][1] == ['amps_0', 'harmonic_distribution_0', 'f0_hz_0']
    assert d2[1][1] == ['noise_magnitudes_0'] and len(d2) == 3


def test_processor_group_plan_and_processors_order():
    import ddsp_piano_b200 as dp
    additive = dp.MultiInharmonic(sample_rate=16000, inference=True, name='additive')
    noise = dp.DynamicSizeFilteredNoise(sample_rate=16000, name='noise')
    kw = dict(additive_controls=['amplitudes', 'harmonic_distribution', 'inharm_coef', 'f0_hz'],
              noise_controls=['magnitudes'])
    g = dp.ProcessorGroup(dp.polyphonic_dag(additive, noise, dp.Reverb(), reverb_controls=['reverb_ir'],
                                            n_synths=4, **kw))
    assert g._plan is not None and len(g._plan['voices']) == 4 and g._plan['ir_key'] == 'reverb_ir'
    # processors in DAG order, additive and noise first (synthesize_from_csv.py:99)
    assert [p.name for p in g.processors[:3]] == ['additive', 'noise', 'add']
    assert dp.ProcessorGroup(dp.polyphonic_dag(additive, noise, n_synths=2, **kw))._plan['reverb'] is None
    # not the polyphonic pattern -> node-by-node walk
    assert dp.ProcessorGroup([(additive, ['a', 'b', 'c', 'd'])])._plan is None
    assert dp.ProcessorGroup(dp.polyphonic_dag(additive, noise, n_synths=2, **kw), fused=False)._plan is None
    # mismatched sample rates cannot be fused
    other = dp.DynamicSizeFilteredNoise(sample_rate=24000, name='noise')
    assert dp.ProcessorGroup(dp.polyphonic_dag(additive, other, n_synths=2, **kw))._plan is None


def test_dag_walk_semantics_with_stub_processors():
    """ddsp DAGLayer semantics: nested 'a/b' lookups, outputs stored under the processor name,
    'out' aliases the last node, input dict is extended in place, training/mask kwargs dropped."""
    import ddsp_piano_b200 as dp

    class Gain(dp.Processor):
        def __init__(self, g, name):
            super().__init__(name=name)
            self.g = g

        def get_controls(self, x):
            return {'x': x}

        def get_signal(self, x):
            return x * self.g

    add = dp.MultiAdd(name='add')
    group = dp.ProcessorGroup([(Gain(2.0, 'a'), ['in']), (Gain(3.0, 'b'), ['a/signal']),
                               (add, ['a/signal', 'b/signal'])])
    feats = {'in': np.array([1.0, 2.0])}
    out = group(feats, return_outputs_dict=True, training=True)
    np.testing.assert_array_equal(out['signal'], [8.0, 16.0])
    assert out['controls'] is feats and feats['out'] is feats['add']
    assert list(feats['add']['controls']) == ['signal_0', 'signal_1']
    with pytest.raises(KeyError):
        dp.nested_lookup('a/nope', feats)
    assert np.array_equal(group({'in': np.array([1.0])}), [8.0])


def test_scale_fn_and_constructor_contract():
    import ddsp_piano_b200 as dp
    from ddsp_piano_b200.engine import scale_fn_id
    assert scale_fn_id(dp.exp_sigmoid) == 0 and scale_fn_id('exp_sigmoid') == 0
    assert scale_fn_id(dp.exp_tanh) == 1 and scale_fn_id(None) == 2
    with pytest.raises(ValueError):
        scale_fn_id(lambda x: x)
    s = dp.MultiInharmonic(frame_rate=250, sample_rate=24000, min_frequency=20, inference=True)
    assert s.name == 'multi_inharmonic' and s.upsampling == 96 and s.sample_rate == 24000
    assert dp.InHarmonic().name == 'inharmonic' and dp.InHarmonic().inference is False
    n = dp.DynamicSizeFilteredNoise(frame_rate=250, sample_rate=16000)
    assert (n.upsampling, n.window_size, n.initial_bias, n.name) == (64, 257, -5.0, 'filtered_noise')
    with pytest.raises(ValueError):
        dp.Reverb(trainable=True)
    with pytest.raises(ValueError):
        dp.Reverb().get_controls(np.zeros([1, 4]))
    # exp_sigmoid / exp_tanh called directly agree with the oracle's
    import torch
    from oracle import ddsp_core_np as core
    from oracle import ddsp_piano_np as ref
    x = np.linspace(-8, 8, 33).astype(np.float32)
    np.testing.assert_allclose(dp.exp_sigmoid(torch.from_numpy(x)).numpy(), core.exp_sigmoid(x), rtol=2e-6)
    np.testing.assert_allclose(dp.exp_tanh(torch.from_numpy(x)).numpy(), ref.exp_tanh(x), rtol=2e-5, atol=1e-9)


# ------------------------------- TF checkpoint reader (no TensorFlow) --------------------------

def test_checkpoint_index_parser_matches_survey_appendix_b():
    """ddsp_piano_b200/checkpoint.py on the shipped dafx22 index (fixture = verbatim copy of
    model_weights/dafx22/ckpt-0.index): tensor names, shapes and byte offsets of SURVEY appendix B."""
    import shutil
    import tempfile
    from ddsp_piano_b200.checkpoint import Checkpoint
    with tempfile.TemporaryDirectory() as d:
        shutil.copyfile(os.path.join(ROOT, 'tests', 'golden', 'dafx22_ckpt-0.index'),
                        os.path.join(d, 'ckpt-0.index'))
        ck = Checkpoint(os.path.join(d, 'ckpt-0'))
    suffix = '/.ATTRIBUTES/VARIABLE_VALUE'
    want = {
        'model/reverb_model/reverb_dict/layer_with_weights-0/embeddings': ([10, 24000], 308892),
        'model/monophonic_network/dense_out/kernel': ([192, 161], 9092),
        'model/context_network/dense_out/kernel': ([64, 32], 772),
        'model/detuner/layer/kernel': ([1, 2], 133384),
        'model/note_release/layer/cell/release_duration': ([], 133400),
        'model/z_encoder/embedding/embeddings': ([10, 16], 52),
    }
    for key, (shape, offset) in want.items():
        e = ck.entries[key + suffix]
        assert e['shape'] == shape and e['offset'] == offset and e['dtype'] == 1
        assert e['size'] == 4 * int(np.prod(shape, dtype=np.int64))
    assert ck.n_shards == 1 and len(ck.keys()) == 35
    with pytest.raises(KeyError):
        ck.tensor('no/such/variable')


@pytest.mark.skipif(not os.path.isdir('/root/reference/ddsp_piano/model_weights'),
                    reason='shipped checkpoints only exist in the build container')
def test_checkpoint_tensors_from_shipped_weights(golden_dir):
    from ddsp_piano_b200.checkpoint import Checkpoint
    weights = '/root/reference/ddsp_piano/model_weights'
    ir = Checkpoint(os.path.join(weights, 'dafx22', 'ckpt-0')).reverb_ir(0)
    fixture = np.load(os.path.join(golden_dir, 'dafx22_reverb_ir_row0.npz'))['ir']
    np.testing.assert_array_equal(ir, fixture)
    p = Checkpoint(os.path.join(weights, 'v2', 'ckpt-225000')).fdn_parameters(0)
    g = np.load(os.path.join(golden_dir, 'v2_fdn_params_piano0.npz'))
    for k, v in p.items():
        np.testing.assert_array_equal(v, g[k])
    assert p['gain_allpass'].shape == (8, 4) and p['early_ir'].shape == (200,)


# ------------------------------- MIDI front end: voice allocation ------------------------------

@pytest.mark.parametrize('case', ['sparse', 'dense', 'chords', 'overflow'])
@pytest.mark.parametrize('n_synths', [16, 4])
def test_midi_roll_to_conditioning_matches_reference(lib_path, case, n_synths):
    """SURVEY 8f row 3 (voice allocation): the C++ port inside libb200ddsp.so vs outputs of the
    reference's own utils/midi_encoders.py (imported unmodified by tests/golden/make_golden.py).
    Integer/index work: bit exact."""
    from ddsp_piano_b200.midi import MIDIRoll2Conditioning
    g = np.load(os.path.join(ROOT, 'tests', 'golden', 'midi_conditioning.npz'))
    roll = g[f'{case}_roll']
    cond, poly = MIDIRoll2Conditioning(n_synths)(roll)
    np.testing.assert_array_equal(poly, g[f'{case}_poly'])
    np.testing.assert_array_equal(cond, g[f'{case}_cond{n_synths}'])
    # properties: every channel holds an active pitch of the frame or 0; a held note never moves
    pitches = cond[..., 0]
    for t in range(1, len(pitches)):
        for ch in range(n_synths):
            p = pitches[t, ch]
            if p != 0 and p in pitches[t - 1]:
                assert pitches[t - 1, ch] == p


def test_midi_roll_shape_errors(lib_path):
    from ddsp_piano_b200.midi import MIDIRoll2Conditioning
    with pytest.raises(ValueError):
        MIDIRoll2Conditioning(16)(np.zeros([10, 88], np.float32))
    with pytest.raises(ValueError):
        MIDIRoll2Conditioning(16)(np.zeros([10, 8, 2], np.float32))
    cond, poly = MIDIRoll2Conditioning(16)(np.zeros([0, 88, 2], np.float32))
    assert cond.shape == (0, 16, 2) and poly.shape == (0,)


# ------------------------------- bench helpers (host logic) -------------------------------------

def test_bench_live_chain_samples_matches_brute_force():
    """bench.live_chain_samples restates the kernels' liveness bookkeeping (16-partial half-groups,
    maximum over the frames a 1000-sample chunk touches, both substrings on every lane)."""
    import bench
    w = dict(P=2, B=2, F=30, H=40, S=2, sr=24000)
    rng = np.random.default_rng(3)
    f0 = np.empty([2, 2, 30, 2], np.float32)
    f0[..., 0] = np.array([[110.0, 8.1758], [1500.0, 440.0]], np.float32)[:, :, None]
    f0[1, 1, 15:, 0] = 3000.0                            # a pitch change inside the clip
    f0[..., 1] = f0[..., 0] * 1.001
    x = {'f0_hz': f0, 'inharm_coef': rng.uniform(1e-4, 1e-3, [2, 2, 30, 1]).astype(np.float32)}
    U, N = 96, 30 * 96
    n = np.arange(1, 41)
    total = 0
    for v in range(2):
        for b in range(2):
            live = np.zeros(30, int)
            for k in range(30):
                fk, bk = float(f0[v, b, k, 0]), float(x['inharm_coef'][v, b, k, 0])
                ok = (fk * n * np.sqrt(1 + bk * n * n) < 12000.0) & (fk > 20.0)
                live[k] = 0 if not ok.any() else -(-(int(np.max(np.nonzero(ok)[0])) + 1) // 16)
            for t0 in range(0, N, 1000):
                t1 = min(N, t0 + 1000) - 1
                nh = live[t0 // U:min(29, t1 // U + 1) + 1].max()
                total += int(nh) * 32 * (t1 + 1 - t0)
    synth, phase = bench.live_chain_samples(w, f0[..., 0], x['inharm_coef'][..., 0])
    assert synth == total
    # the phase pass follows a half-group through chunk c only if a LATER chunk sounds; a span that hands
    # its state on follows every half-group (3 for H = 40) through every chunk
    assert 0 < phase < synth
    assert bench.live_chain_samples(w, f0[..., 0], x['inharm_coef'][..., 0], carry_all=True)[1] == 4 * 3 * 32 * N
    assert bench.bind_to_gpu_numa_node(0) in (None, 0, 1, 2, 3)   # never raises without a GPU


def test_bench_config1_key_never_raises():
    """The informational "config1" key of the bench line (configs[0] beside the headline) reports a failure
    instead of raising: without a GPU it is an error entry, on a GPU the two shipped models."""
    import torch
    import bench
    out = bench.config1_line()
    if torch.cuda.is_available():
        assert {'dafx22', 'maestro_v2', 'what'} <= set(out) and out['dafx22']['rtf'] > 50
    else:
        assert set(out) == {'error'}


def test_checkpoint_directory_resolves_to_its_latest_prefix(tmp_path):
    """The reference's --ckpt default is a DIRECTORY (model_weights/v2/): the reader resolves it through
    the `checkpoint` state file like tf.train.latest_checkpoint; a prefix passes through unchanged."""
    from ddsp_piano_b200.checkpoint import latest_checkpoint
    (tmp_path / 'checkpoint').write_text('model_checkpoint_path: "ckpt-225000"\nall_model_checkpoint_paths: "ckpt-225000"\n')
    assert latest_checkpoint(str(tmp_path)) == str(tmp_path / 'ckpt-225000')
    assert latest_checkpoint(str(tmp_path / 'ckpt-7')) == str(tmp_path / 'ckpt-7')
    empty = tmp_path / 'empty'
    empty.mkdir()
    with pytest.raises(FileNotFoundError):
        latest_checkpoint(str(empty))


def test_gin_registration_registers_the_drop_in_classes(monkeypatch):
    """ddsp_piano_b200.gin_registration calls gin.external_configurable for every class
    the reference's gin files bind (configs/dafx22.gin:91-100); gin itself is not a dependency of this package,
    so the test hands it a recording stand-in, and checks the failure mode without it."""
    import importlib
    import sys
    import types
    import ddsp_piano_b200 as dp
    calls = {}
    fake = types.ModuleType('gin')
    fake.external_configurable = lambda fn, name=None, module=None: calls.setdefault(name, (fn, module)) and fn
    monkeypatch.setitem(sys.modules, 'gin', fake)
    sys.modules.pop('ddsp_piano_b200.gin_registration', None)
    reg = importlib.import_module('ddsp_piano_b200.gin_registration')
    for name in ('MultiInharmonic', 'DynamicSizeFilteredNoise', 'Reverb', 'MultiAdd', 'ProcessorGroup',
                 'polyphonic_dag', 'exp_tanh', 'FeedbackDelayNetwork'):
        assert calls[name] == (getattr(dp, name), 'ddsp_piano_b200')
    assert set(reg.registered) == set(reg.CONFIGURABLES)
    sys.modules.pop('ddsp_piano_b200.gin_registration', None)
    monkeypatch.setitem(sys.modules, 'gin', None)           # import gin -> ImportError
    with pytest.raises(ImportError, match='gin-config'):
        importlib.import_module('ddsp_piano_b200.gin_registration')
    sys.modules.pop('ddsp_piano_b200.gin_registration', None)
