"""ctypes binding of ``libb200ddsp.so`` (C ABI declared in ``include/b200ddsp.h``).

The library is built in-tree by :func:`build` (``nvcc`` for sm_100a) and must be present:
there is no CPU or PyTorch fallback -- loading fails loudly instead.
"""
import ctypes
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, 'csrc')
# B200DDSP_LIB: developer switch for A/B timing of an alternative in-tree build of the same sources
LIB_PATH = os.environ.get('B200DDSP_LIB') or os.path.join(HERE, 'libb200ddsp.so')
HEADER = os.path.join(ROOT, 'include', 'b200ddsp.h')

OK = 0
STATUS_NAMES = {0: 'OK', -1: 'BAD_SHAPE', -2: 'BAD_ALIGN', -3: 'UNSUPPORTED_CONFIG',
                -4: 'CUDA_ERROR', -5: 'WORKSPACE_TOO_SMALL', -6: 'BAD_ARGUMENT'}
SCALE_FN_IDS = {'exp_sigmoid': 0, 'exp_tanh': 1, None: 2, 'none': 2}
MAX_VOICES = 64

c_float_p = ctypes.POINTER(ctypes.c_float)


class Config(ctypes.Structure):
    """``b200ddsp_config``."""
    _fields_ = [('sample_rate', ctypes.c_int), ('frame_rate', ctypes.c_int),
                ('min_frequency', ctypes.c_float), ('additive_scale_fn', ctypes.c_int),
                ('normalize_after_nyquist_cut', ctypes.c_int),
                ('normalize_below_nyquist', ctypes.c_int), ('inference', ctypes.c_int),
                ('noise_scale_fn', ctypes.c_int), ('noise_initial_bias', ctypes.c_float),
                ('noise_window_size', ctypes.c_int), ('reverb_add_dry', ctypes.c_int),
                ('n_noise_bands', ctypes.c_int), ('fast_phase', ctypes.c_int)]


class Voice(ctypes.Structure):
    """``b200ddsp_voice``: device pointers of one voice's raw control tensors."""
    _fields_ = [('amplitudes', ctypes.c_void_p), ('harmonic_distribution', ctypes.c_void_p),
                ('inharm_coef', ctypes.c_void_p), ('f0_hz', ctypes.c_void_p),
                ('magnitudes', ctypes.c_void_p), ('noise', ctypes.c_void_p)]


class Link(ctypes.Structure):
    """``b200ddsp_link``: one rank's end of a stream-ordered hand-off through peer memory."""
    _fields_ = [('seed', ctypes.c_void_p), ('seed_ready', ctypes.c_void_p), ('seed_ack', ctypes.c_void_p),
                ('carry', ctypes.c_void_p), ('carry_ready', ctypes.c_void_p), ('carry_ack', ctypes.c_void_p),
                ('epoch', ctypes.c_ulonglong), ('scratch', ctypes.c_void_p)]


class Span(ctypes.Structure):
    """``b200ddsp_span``: which frames of a timeline a call reads and synthesises."""
    _fields_ = [('in_first_frame', ctypes.c_longlong), ('out_first_frame', ctypes.c_longlong),
                ('n_out_frames', ctypes.c_int), ('total_frames', ctypes.c_longlong), ('phase', Link)]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC)
                  if f.endswith(('.cu', '.cuh', '.inl'))) + [HEADER]


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    built = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(s) > built for s in sources())


def build(force=False, verbose=False):
    """Compile ``csrc/b200ddsp.cu`` into ``libb200ddsp.so`` for sm_100a (cross-compiles
    without a GPU)."""
    if not force and not needs_build():
        return LIB_PATH
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    cmd = [nvcc, '-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
           '-Xcompiler', '-fPIC', '-shared', '-o', LIB_PATH, os.path.join(CSRC, 'b200ddsp.cu')]
    cmd[1:1] = os.environ.get('B200DDSP_NVCC_FLAGS', '').split()
    if verbose:
        cmd.insert(1, '-Xptxas=-v')
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError('nvcc failed:\n' + proc.stdout + proc.stderr)
    if verbose:
        print(proc.stderr)
    return LIB_PATH


_LIB = None


def load():
    """Load the shared library and declare its prototypes.  Raises if it is missing: the
    product has no other execution path."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f'{LIB_PATH} is missing: build it with `python -c "import __graft_entry__ as g; '
            'g.build()"` (nvcc, sm_100a).  There is no CPU fallback.')
    lib = ctypes.CDLL(LIB_PATH)
    vp, sz, u64, ci = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint64, ctypes.c_int
    lib.b200ddsp_version.restype = ci
    lib.b200ddsp_version.argtypes = []
    lib.b200ddsp_last_error.restype = ctypes.c_char_p
    lib.b200ddsp_last_error.argtypes = [vp]
    lib.b200ddsp_create.restype = ci
    lib.b200ddsp_create.argtypes = [ctypes.POINTER(Config), ctypes.POINTER(vp)]
    lib.b200ddsp_destroy.restype = ci
    lib.b200ddsp_destroy.argtypes = [vp]
    lib.b200ddsp_workspace_bytes.restype = sz
    lib.b200ddsp_workspace_bytes.argtypes = [vp, ci, ci, ci, ci, ci, ci, ci]
    lib.b200ddsp_additive_workspace_bytes.restype = sz
    lib.b200ddsp_additive_workspace_bytes.argtypes = [vp, ci, ci, ci, ci]
    lib.b200ddsp_additive_controls.restype = ci
    lib.b200ddsp_additive_controls.argtypes = [vp] + [vp] * 7 + [ci, ci, ci, ci, vp]
    lib.b200ddsp_additive_signal.restype = ci
    lib.b200ddsp_additive_signal.argtypes = [vp] + [vp] * 5 + [ci, ci, ci, ci, ci, vp, sz, vp]
    lib.b200ddsp_noise_controls.restype = ci
    lib.b200ddsp_noise_controls.argtypes = [vp, vp, vp, sz, vp]
    lib.b200ddsp_noise_signal.restype = ci
    lib.b200ddsp_noise_signal.argtypes = [vp, vp, vp, u64, u64, vp, ci, ci, ci, ci, vp, sz, vp]
    lib.b200ddsp_noise_workspace_bytes.restype = sz
    lib.b200ddsp_noise_workspace_bytes.argtypes = [vp, ci, ci, ci]
    lib.b200ddsp_reverb.restype = ci
    lib.b200ddsp_reverb.argtypes = [vp, vp, vp, vp, ci, ci, ci, vp, sz, vp]
    lib.b200ddsp_reverb_full.restype = ci
    lib.b200ddsp_reverb_full.argtypes = [vp, vp, vp, vp, ci, ci, ci, vp, sz, vp]
    lib.b200ddsp_ir_decay_mask.restype = ci
    lib.b200ddsp_ir_decay_mask.argtypes = [vp, vp, vp, ci, ci, ctypes.c_float, ci, vp]
    lib.b200ddsp_surrogate_decays.restype = ci
    lib.b200ddsp_surrogate_decays.argtypes = [vp, vp, vp, vp, vp, ci, ci, ci, vp]
    lib.b200ddsp_surrogate_signal.restype = ci
    lib.b200ddsp_surrogate_signal.argtypes = [vp] + [vp] * 7 + [ci, ci, ci, vp, sz, vp]
    lib.b200ddsp_peer_alloc.restype = ci
    lib.b200ddsp_peer_alloc.argtypes = [vp, sz, ctypes.POINTER(vp), ctypes.c_char_p]
    lib.b200ddsp_peer_free.restype = ci
    lib.b200ddsp_peer_free.argtypes = [vp, vp]
    lib.b200ddsp_peer_open.restype = ci
    lib.b200ddsp_peer_open.argtypes = [vp, ctypes.c_char_p, ctypes.POINTER(vp)]
    lib.b200ddsp_peer_close.restype = ci
    lib.b200ddsp_peer_close.argtypes = [vp, vp]
    span_p, link_p = ctypes.POINTER(Span), ctypes.POINTER(Link)
    lib.b200ddsp_forward_span.restype = ci
    lib.b200ddsp_forward_span.argtypes = [vp, ctypes.POINTER(Voice), ci, vp, ci, ci, ci, ci, ci, u64,
                                          span_p, vp, sz, vp]
    lib.b200ddsp_forward_timeline.restype = ci
    lib.b200ddsp_forward_timeline.argtypes = [vp, ctypes.POINTER(Voice), ci, vp, vp, vp, ci, ci, ci, ci, ci,
                                              ci, ci, u64, span_p, link_p, vp, sz, vp]
    lib.b200ddsp_forward_timeline_host.restype = ci
    lib.b200ddsp_forward_timeline_host.argtypes = lib.b200ddsp_forward_timeline.argtypes
    lib.b200ddsp_timeline_workspace_bytes.restype = sz
    lib.b200ddsp_timeline_workspace_bytes.argtypes = [vp] + [ci] * 11
    lib.b200ddsp_timeline_reverb.restype = ci
    lib.b200ddsp_timeline_reverb.argtypes = [vp, vp, vp, vp, ci, ci, ci, ci, link_p, vp, sz, vp]
    lib.b200ddsp_timeline_reverb_workspace_bytes.restype = sz
    lib.b200ddsp_timeline_reverb_workspace_bytes.argtypes = [vp, ci, ci, ci, ci]
    lib.b200ddsp_note_release.restype = ci
    lib.b200ddsp_note_release.argtypes = [vp, vp, vp, ci, ci, ci, ctypes.c_float, vp]
    lib.b200ddsp_gru_recurrence.restype = ci
    lib.b200ddsp_gru_recurrence.argtypes = [vp, vp, vp, vp, vp, ci, ci, ci, vp]
    lib.b200ddsp_fft_convolve.restype = ci
    lib.b200ddsp_fft_convolve.argtypes = [vp, vp, vp, vp, ci, ci, ci, ci, vp, sz, vp]
    lib.b200ddsp_fdn_ir.restype = ci
    lib.b200ddsp_fdn_ir.argtypes = [vp] + [vp] * 7 + [ci, c_float_p, ci, ctypes.c_float, vp, ci, vp, sz, vp]
    lib.b200ddsp_fdn_workspace_bytes.restype = sz
    lib.b200ddsp_fdn_workspace_bytes.argtypes = [vp, ctypes.c_float, ci]
    lib.b200ddsp_forward_polyphonic.restype = ci
    lib.b200ddsp_forward_polyphonic.argtypes = [vp, ctypes.POINTER(Voice), ci, vp, vp, vp,
                                                ci, ci, ci, ci, ci, ci, u64, vp, sz, vp]
    lib.b200ddsp_forward_polyphonic_host.restype = ci
    lib.b200ddsp_forward_polyphonic_host.argtypes = [vp, ctypes.POINTER(Voice), ci, vp, vp, vp,
                                                     ci, ci, ci, ci, ci, ci, u64, vp, sz, vp]
    lib.b200ddsp_workspace_bytes_host.restype = sz
    lib.b200ddsp_workspace_bytes_host.argtypes = [vp, ci, ci, ci, ci, ci, ci, ci, ci]
    lib.b200ddsp_midi_roll_to_conditioning.restype = ci
    lib.b200ddsp_midi_roll_to_conditioning.argtypes = [vp, ci, ci, ci, ctypes.c_float, vp, vp]
    lib.b200ddsp_measure_fma_rate.restype = ci
    lib.b200ddsp_measure_fma_rate.argtypes = [vp, ci, ctypes.POINTER(ctypes.c_double), vp]
    lib.b200ddsp_launch_count.restype = u64
    lib.b200ddsp_launch_count.argtypes = [vp]
    lib.b200ddsp_set_profiling.restype = ci
    lib.b200ddsp_set_profiling.argtypes = [vp, ci]
    lib.b200ddsp_last_stage_ms.restype = ci
    lib.b200ddsp_last_stage_ms.argtypes = [vp, c_float_p]
    _LIB = lib
    return lib


EXPORTS = ['b200ddsp_version', 'b200ddsp_last_error', 'b200ddsp_create', 'b200ddsp_destroy',
           'b200ddsp_workspace_bytes', 'b200ddsp_additive_workspace_bytes',
           'b200ddsp_additive_controls', 'b200ddsp_additive_signal',
           'b200ddsp_noise_controls', 'b200ddsp_noise_signal', 'b200ddsp_noise_workspace_bytes', 'b200ddsp_reverb', 'b200ddsp_reverb_full', 'b200ddsp_fft_convolve', 'b200ddsp_ir_decay_mask', 'b200ddsp_note_release', 'b200ddsp_gru_recurrence', 'b200ddsp_surrogate_decays', 'b200ddsp_surrogate_signal', 'b200ddsp_peer_alloc', 'b200ddsp_peer_free',
           'b200ddsp_peer_open', 'b200ddsp_peer_close', 'b200ddsp_forward_span', 'b200ddsp_forward_timeline',
           'b200ddsp_forward_timeline_host', 'b200ddsp_timeline_workspace_bytes', 'b200ddsp_timeline_reverb',
           'b200ddsp_timeline_reverb_workspace_bytes', 'b200ddsp_fdn_ir',
           'b200ddsp_fdn_workspace_bytes',
           'b200ddsp_forward_polyphonic', 'b200ddsp_forward_polyphonic_host',
           'b200ddsp_workspace_bytes_host', 'b200ddsp_midi_roll_to_conditioning',
           'b200ddsp_launch_count', 'b200ddsp_measure_fma_rate', 'b200ddsp_set_profiling',
           'b200ddsp_last_stage_ms']
CONV_MASK_IR0, CONV_ADD_DRY, CONV_FULL = 1, 2, 4
STAGES = ['controls', 'phase_ends', 'phase_scan', 'oscillators', 'noise', 'reverb', 'mix']
