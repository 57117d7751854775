"""Host-side owner of a ``b200ddsp_handle``: validates tensors, owns the scratch workspace
(a torch CUDA buffer -- PyTorch is used for device memory and streams only) and turns
status codes into the exceptions the reference raises (``ValueError`` for shape problems,
like ddsp; ``RuntimeError`` for CUDA failures).
"""
import ctypes
import threading

import torch

from . import _lib

_ENGINES = {}
_LOCK = threading.Lock()


def scale_fn_id(fn):
    """Map the reference's ``scale_fn`` argument (``core.exp_sigmoid``, ``exp_tanh``
    [modules/inharm_synth.py:13-17] or ``None``) to the ABI enum."""
    if fn is None:
        return 2
    name = fn if isinstance(fn, str) else getattr(fn, '__name__', None)
    if name in ('exp_sigmoid', 'core.exp_sigmoid'):
        return 0
    if name == 'exp_tanh':
        return 1
    if name in ('none', 'None'):
        return 2
    raise ValueError(f'unsupported scale_fn {fn!r}: the CUDA path implements exp_sigmoid, '
                     'exp_tanh and None')


class Engine:
    """One handle = one processor configuration on one device."""

    def __init__(self, device, **cfg):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError('b200ddsp needs a CUDA device (sm_100a); there is no CPU fallback')
        self.device = torch.device(device)
        self.cfg = _lib.Config(**cfg)
        self.handle = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            rc = self.lib.b200ddsp_create(ctypes.byref(self.cfg), ctypes.byref(self.handle))
        if rc != 0:
            msg = self.lib.b200ddsp_last_error(None).decode()
            raise (ValueError if rc in (-1, -3, -6) else RuntimeError)(
                f'b200ddsp_create: {_lib.STATUS_NAMES.get(rc, rc)}: {msg}')
        self._workspace = None
        self._host_out = {}
        self._ws_sizes = {}
        self._keep_alive = None

    def __del__(self):
        try:
            if getattr(self, 'handle', None):
                self.lib.b200ddsp_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    # -- helpers --------------------------------------------------------------------------
    def check(self, rc):
        if rc == 0:
            return
        msg = self.lib.b200ddsp_last_error(self.handle).decode()
        text = f'b200ddsp {_lib.STATUS_NAMES.get(rc, rc)}: {msg}'
        if rc in (-1, -2, -3, -6):
            raise ValueError(text)
        raise RuntimeError(text)

    def tensor(self, x, name, ndim=None):
        """ddsp.core.tf_float32 + device/contiguity requirements of the ABI."""
        if not isinstance(x, torch.Tensor):
            x = torch.as_tensor(x, dtype=torch.float32, device=self.device)
        if x.device != self.device:
            raise ValueError(f'{name} is on {x.device}, engine is on {self.device}')
        if x.dtype != torch.float32:
            x = x.to(torch.float32)
        if not x.is_contiguous():
            x = x.contiguous()
        if ndim is not None and x.dim() != ndim:
            raise ValueError(f'{name} must have {ndim} dimensions, got shape {tuple(x.shape)}')
        return x

    def workspace(self, nbytes):
        if self._workspace is None or self._workspace.numel() < nbytes:
            self._workspace = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=self.device)
        return self._workspace

    def stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    @property
    def upsampling(self):
        return int(self.cfg.sample_rate / self.cfg.frame_rate)

    def launch_count(self):
        return int(self.lib.b200ddsp_launch_count(self.handle))

    def set_profiling(self, enable):
        self.check(self.lib.b200ddsp_set_profiling(self.handle, int(bool(enable))))

    def measure_fma_rate(self, packed=True):
        """FMA-pipe lane-operations per second this GPU sustains right now (bench.py's FP32 peak)."""
        v = ctypes.c_double()
        with torch.cuda.device(self.device):
            self.check(self.lib.b200ddsp_measure_fma_rate(self.handle, int(packed), ctypes.byref(v),
                                                          self.stream()))
        return float(v.value)

    def last_stage_ms(self):
        """Device time of each stage of the most recent call (needs set_profiling(True))."""
        ms = (ctypes.c_float * len(_lib.STAGES))()
        self.check(self.lib.b200ddsp_last_stage_ms(self.handle, ms))
        return dict(zip(_lib.STAGES, [float(x) for x in ms]))

    # -- entry points ---------------------------------------------------------------------
    def additive_controls(self, amplitudes, harmonic_distribution, inharm_coef, f0_hz):
        amplitudes = self.tensor(amplitudes, 'amplitudes', 3)
        hd = self.tensor(harmonic_distribution, 'harmonic_distribution', 3)
        inharm_coef = self.tensor(inharm_coef, 'inharm_coef', 3)
        f0_hz = self.tensor(f0_hz, 'f0_hz', 3)
        B, F, H = hd.shape
        S = f0_hz.shape[-1]
        for t, n, c in ((amplitudes, 'amplitudes', 1), (inharm_coef, 'inharm_coef', 1),
                        (f0_hz, 'f0_hz', S)):
            if tuple(t.shape) != (B, F, c):
                raise ValueError(f'{n} has shape {tuple(t.shape)}, expected {(B, F, c)}')
        amp_out = torch.empty_like(amplitudes)
        hd_out = torch.empty_like(hd)
        shifts_out = torch.empty_like(hd)
        with torch.cuda.device(self.device):
            self.check(self.lib.b200ddsp_additive_controls(
                self.handle, amplitudes.data_ptr(), hd.data_ptr(), inharm_coef.data_ptr(),
                f0_hz.data_ptr(), amp_out.data_ptr(), hd_out.data_ptr(), shifts_out.data_ptr(),
                B, F, H, S, self.stream()))
        return {'amplitudes': amp_out, 'harmonic_distribution': hd_out,
                'harmonic_shifts': shifts_out, 'f0_hz': f0_hz}

    def additive_signal(self, amplitudes, harmonic_distribution, harmonic_shifts, f0_hz):
        amplitudes = self.tensor(amplitudes, 'amplitudes', 3)
        hd = self.tensor(harmonic_distribution, 'harmonic_distribution', 3)
        shifts = self.tensor(harmonic_shifts, 'harmonic_shifts', 3)
        f0_hz = self.tensor(f0_hz, 'f0_hz', 3)
        B, F, H = hd.shape
        S = f0_hz.shape[-1]
        if tuple(amplitudes.shape) != (B, F, 1) or tuple(shifts.shape) != (B, F, H) or \
                tuple(f0_hz.shape[:2]) != (B, F):
            raise ValueError('additive controls have inconsistent shapes: '
                             f'{tuple(amplitudes.shape)}, {tuple(hd.shape)}, '
                             f'{tuple(shifts.shape)}, {tuple(f0_hz.shape)}')
        N = F * self.upsampling
        out = torch.empty([B, N], dtype=torch.float32, device=self.device)
        ws = self.workspace(self.lib.b200ddsp_additive_workspace_bytes(self.handle, B, F, H, S))
        with torch.cuda.device(self.device):
            self.check(self.lib.b200ddsp_additive_signal(
                self.handle, amplitudes.data_ptr(), hd.data_ptr(), shifts.data_ptr(),
                f0_hz.data_ptr(), out.data_ptr(), B, F, H, S, 0, ws.data_ptr(), ws.numel(),
                self.stream()))
        return out

    def surrogate_decays(self, decays, inharm_coef, f0_hz):
        """SurrogateAdditive.get_controls, decay clipping (surrogate_synth.py:163-171)."""
        decays = self.tensor(decays, 'decays', 3)
        inharm_coef = self.tensor(inharm_coef, 'inharm_coef', 3)
        f0_hz = self.tensor(f0_hz, 'f0_hz', 3)
        B, F, H = decays.shape
        if tuple(inharm_coef.shape) != (B, F, 1) or tuple(f0_hz.shape) != (B, F, 1):
            raise ValueError(f'inharm_coef {tuple(inharm_coef.shape)} / f0_hz {tuple(f0_hz.shape)} '
                             f'must be {(B, F, 1)}')
        out = torch.empty_like(decays)
        with torch.cuda.device(self.device):
            self.check(self.lib.b200ddsp_surrogate_decays(
                self.handle, decays.data_ptr(), inharm_coef.data_ptr(), f0_hz.data_ptr(),
                out.data_ptr(), B, F, H, self.stream()))
        return out

    def surrogate_signal(self, amplitudes, decays, decay_time, harmonic_distribution, harmonic_shifts,
                         f0_hz):
        amplitudes = self.tensor(amplitudes, 'amplitudes', 3)
        decays = self.tensor(decays, 'decays', 3)
        decay_time = self.tensor(decay_time, 'decay_time', 3)
        hd = self.tensor(harmonic_distribution, 'harmonic_distribution', 3)
        shifts = self.tensor(harmonic_shifts, 'harmonic_shifts', 3)
        f0_hz = self.tensor(f0_hz, 'f0_hz', 3)
        B, F, H = hd.shape
        for t, n, c in ((amplitudes, 'amplitudes', 1), (decays, 'decays', H), (decay_time, 'decay_time', 1),
                        (shifts, 'harmonic_shifts', H), (f0_hz, 'f0_hz', 1)):
            if tuple(t.shape) != (B, F, c):
                raise ValueError(f'{n} has shape {tuple(t.shape)}, expected {(B, F, c)}')
        out = torch.empty([B, F * self.upsampling], dtype=torch.float32, device=self.device)
        ws = self.workspace(self.lib.b200ddsp_additive_workspace_bytes(self.handle, B, F, H, 1))
        with torch.cuda.device(self.device):
            self.check(self.lib.b200ddsp_surrogate_signal(
                self.handle, amplitudes.data_ptr(), decays.data_ptr(), decay_time.data_ptr(),
                hd.data_ptr(), shifts.data_ptr(), f0_hz.data_ptr(), out.data_ptr(), B, F, H,
                ws.data_ptr(), ws.numel(), self.stream()))
        return out

    def noise_controls(self, magnitudes):
        magnitudes = self.tensor(magnitudes, 'magnitudes')
        out = torch.empty_like(magnitudes)
        with torch.cuda.device(self.device):
            self.check(self.lib.b200ddsp_noise_controls(
                self.handle, magnitudes.data_ptr(), out.data_ptr(), magnitudes.numel(),
                self.stream()))
        return {'magnitudes': out}

    def noise_signal(self, magnitudes, noise=None, seed=0, stream_id=0):
        magnitudes = self.tensor(magnitudes, 'magnitudes', 3)
        B, F, M = magnitudes.shape
        N = F * self.upsampling
        if noise is not None:
            noise = self.tensor(noise, 'noise', 2)
            if tuple(noise.shape) != (B, N):
                raise ValueError(f'noise has shape {tuple(noise.shape)}, expected {(B, N)}')
        out = torch.empty([B, N], dtype=torch.float32, device=self.device)
        ws = self.workspace(self.lib.b200ddsp_noise_workspace_bytes(self.handle, B, F, M))
        with torch.cuda.device(self.device):
            self.check(self.lib.b200ddsp_noise_signal(
                self.handle, magnitudes.data_ptr(), noise.data_ptr() if noise is not None else None,
                seed, stream_id, out.data_ptr(), B, F, M, 0, ws.data_ptr(), ws.numel(),
                self.stream()))
        return out

    def reverb_full(self, audio, ir):
        """'valid'-padded wet signal [B, N + L - 1] (no dry), see sharding.timeline_reverb."""
        return self.reverb(audio, ir, full=True)

    def ir_decay_mask(self, ir, decay_exponent=4.0, decay_start=16000):
        ir = self.tensor(ir, 'ir', 2)
        out = torch.empty_like(ir)
        with torch.cuda.device(self.device):
            self.check(self.lib.b200ddsp_ir_decay_mask(
                self.handle, ir.data_ptr(), out.data_ptr(), ir.shape[0], ir.shape[1],
                float(decay_exponent), int(decay_start), self.stream()))
        return out

    # -- peer-visible buffers (sharding.SpanChain) --------------------------------------------------
    def peer_alloc(self, nbytes):
        """cudaMalloc + CUDA IPC handle: (device pointer, 64-byte handle)."""
        import ctypes
        ptr = ctypes.c_void_p()
        handle = ctypes.create_string_buffer(64)
        with torch.cuda.device(self.device):
            self.check(self.lib.b200ddsp_peer_alloc(self.handle, nbytes, ctypes.byref(ptr), handle))
        return ptr.value, handle.raw

    def peer_open(self, handle):
        import ctypes
        ptr = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            self.check(self.lib.b200ddsp_peer_open(self.handle, handle, ctypes.byref(ptr)))
        return ptr.value

    def peer_close(self, ptr):
        with torch.cuda.device(self.device):
            self.check(self.lib.b200ddsp_peer_close(self.handle, ptr))

    def peer_free(self, ptr):
        with torch.cuda.device(self.device):
            self.check(self.lib.b200ddsp_peer_free(self.handle, ptr))

    # -- spans of a timeline (sharding.SpanChain builds the links) ----------------------------------
    def _span_voices(self, voices, reverb_ir, host):
        arr, keep, (P, B, F, H, S, M, _), any_noise = self._voice_array(voices, host=host)
        ir, L = None, 0
        if reverb_ir is not None:
            ir = self._impulse_response(reverb_ir, B, host=host)
            L = ir.shape[1]
        return arr, keep, (P, B, F, H, S, M, L), any_noise, ir

    def forward_span(self, voices, span, seed=0):
        """Additive + noise of one span of B timelines (``span``: a ``_lib.Span``; control tensors cover
        the span's INPUT frames) -> dry [B, n_out_frames * U]."""
        arr, keep, (P, B, F, H, S, M, _), _, _ = self._span_voices(voices, None, host=False)
        ws = self.workspace(self.lib.b200ddsp_workspace_bytes(self.handle, P, B, F, H, S, M, 0))
        dry = torch.empty([B, span.n_out_frames * self.upsampling], dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            self.check(self.lib.b200ddsp_forward_span(
                self.handle, arr, P, dry.data_ptr(), B, F, H, S, M, seed, ctypes.byref(span),
                ws.data_ptr(), ws.numel(), self.stream()))
        return dry

    def forward_timeline(self, voices, reverb_ir, span, seg_frames, tail=None, seed=0, want_dry=True):
        """The span forward + the reverb of the timeline (segment convolutions, overlap-add, tail
        hand-off through ``tail``: a ``_lib.Link`` or None).  HOST control tensors take the staged
        entry point and come back as pinned host tensors (see forward_polyphonic_host).
        Returns (dry or None, wet), each [B, n_out_frames * U]."""
        f0_0 = voices[0]['f0_hz']
        host = not (isinstance(f0_0, torch.Tensor) and f0_0.device.type == 'cuda')
        arr, keep, (P, B, F, H, S, M, L), any_noise, ir = self._span_voices(voices, reverb_ir, host=host)
        N = span.n_out_frames * self.upsampling
        key = ('timeline', host, P, B, F, H, S, M, L, span.n_out_frames, seg_frames, any_noise)
        nbytes = self._ws_sizes.get(key)
        if nbytes is None:
            nbytes = self._ws_sizes[key] = self.lib.b200ddsp_timeline_workspace_bytes(
                self.handle, P, B, F, H, S, M, L, span.n_out_frames, seg_frames, int(host), int(any_noise))
        if nbytes == 0:
            raise ValueError('forward_timeline: bad shapes')
        ws = self.workspace(nbytes)
        if host:
            bufs = self._host_out.get((B, N))
            if bufs is None:
                bufs = self._host_out[(B, N)] = (torch.empty([B, N], dtype=torch.float32).pin_memory(),
                                                 torch.empty([B, N], dtype=torch.float32).pin_memory())
            dry, wet = (bufs[0] if want_dry else None), bufs[1]
            self._keep_alive = (keep, ir)
            fn = self.lib.b200ddsp_forward_timeline_host
        else:
            dry = torch.empty([B, N], dtype=torch.float32, device=self.device) if want_dry else None
            wet = torch.empty([B, N], dtype=torch.float32, device=self.device)
            fn = self.lib.b200ddsp_forward_timeline
        with torch.cuda.device(self.device):
            self.check(fn(self.handle, arr, P, ir.data_ptr(), dry.data_ptr() if dry is not None else None,
                          wet.data_ptr(), B, F, H, S, M, L, seg_frames, seed, ctypes.byref(span),
                          ctypes.byref(tail) if tail is not None else None, ws.data_ptr(), ws.numel(),
                          self.stream()))
        return dry, wet

    def timeline_reverb(self, dry, ir, n_seg, tail=None):
        """Reverb of a dry span [B, n_seg * N] with the timelines' impulse responses [B, L] (segment
        convolutions + overlap-add + tail hand-off); the reverb stage of forward_timeline alone."""
        dry = self.tensor(dry, 'dry', 2)
        B = dry.shape[0]
        ir = self._impulse_response(ir, B, host=False)
        L = ir.shape[1]
        if dry.shape[1] % n_seg:
            raise ValueError(f'span of {dry.shape[1]} samples is not {n_seg} equal segments')
        N = dry.shape[1] // n_seg
        ws = self.workspace(self.lib.b200ddsp_timeline_reverb_workspace_bytes(self.handle, B, n_seg, N, L))
        out = torch.empty_like(dry)
        with torch.cuda.device(self.device):
            self.check(self.lib.b200ddsp_timeline_reverb(
                self.handle, dry.data_ptr(), ir.data_ptr(), out.data_ptr(), B, n_seg, N, L,
                ctypes.byref(tail) if tail is not None else None, ws.data_ptr(), ws.numel(), self.stream()))
        return out

    def note_release(self, conditioning, release_frames):
        """NoteRelease over conditioning [rows, F, 2] (pitch column) or active pitch [rows, F, 1]
        -> extended_pitch [rows, F, 1]."""
        c = self.tensor(conditioning, 'conditioning', 3)
        out = torch.empty(c.shape[0], c.shape[1], 1, dtype=torch.float32, device=c.device)
        with torch.cuda.device(self.device):
            self.check(self.lib.b200ddsp_note_release(
                self.handle, c.data_ptr(), out.data_ptr(), c.shape[0], c.shape[1], c.shape[2],
                float(release_frames), self.stream()))
        return out

    def gru_recurrence(self, x_proj, w_hh, b_hh):
        """GRU recurrence over x_proj [rows, F, 3u] (input projections, gates r, z, n) -> [rows, F, u]."""
        x = self.tensor(x_proj, 'x_proj', 3)
        w, b = self.tensor(w_hh, 'w_hh', 2), self.tensor(b_hh, 'b_hh', 1)
        u = w.shape[1]
        if w.shape[0] != 3 * u or b.shape[0] != 3 * u or x.shape[2] != 3 * u:
            raise ValueError(f'GRU shapes disagree: x_proj {tuple(x.shape)}, w_hh {tuple(w.shape)}, b_hh {tuple(b.shape)}')
        out = torch.empty(x.shape[0], x.shape[1], u, dtype=torch.float32, device=x.device)
        with torch.cuda.device(self.device):
            self.check(self.lib.b200ddsp_gru_recurrence(
                self.handle, x.data_ptr(), w.data_ptr(), b.data_ptr(), out.data_ptr(), x.shape[0], x.shape[1],
                u, self.stream()))
        return out

    def fft_convolve(self, audio, ir, mask_ir0=False, add_dry=False, full=False):
        """ddsp.core.fft_convolve(audio, ir, delay_compensation=0) + the reverbs' options."""
        flags = (_lib.CONV_MASK_IR0 if mask_ir0 else 0) | (_lib.CONV_ADD_DRY if add_dry else 0) | \
            (_lib.CONV_FULL if full else 0)
        return self.reverb(audio, ir, full=full, flags=flags)

    def fdn_ir(self, input_gain, output_gain, gain_allpass, delays_allpass, time_rev_0_sec,
               alpha_tone, early_ir, sampling_rate, delay_values=None):
        """FeedbackDelayNetwork.get_ir for a batch of parameter rows -> [B, int(2 * sampling_rate)].
        The number of delay lines D (8, or 6 as in configs/ENSTDkCl-*.gin) is the last axis of
        ``input_gain``."""
        ig = self.tensor(input_gain, 'input_gain')
        batched = ig.dim() == 2
        B = ig.shape[0] if batched else 1
        D = int(ig.shape[-1])
        def prep(x, name, shape):
            x = self.tensor(x, name).reshape(shape)
            return x if x.is_contiguous() else x.contiguous()
        ig = prep(ig, 'input_gain', [B, D])
        og = prep(output_gain, 'output_gain', [B, D])
        ga = prep(gain_allpass, 'gain_allpass', [B, D, 4])
        da = prep(delays_allpass, 'delays_allpass', [B, D, 4])
        t0 = prep(time_rev_0_sec, 'time_rev_0_sec', [B])
        al = prep(alpha_tone, 'alpha_tone', [B])
        er = self.tensor(early_ir, 'early_ir')
        er = prep(er, 'early_ir', [B, er.numel() // B])
        E = er.shape[1]
        n = int(2 * float(sampling_rate))
        dv = None
        if delay_values is not None:
            vals = [float(v) for v in torch.as_tensor(delay_values).reshape(-1).tolist()]
            if len(vals) != D:
                raise ValueError(f'{len(vals)} delay values for {D} delay lines')
            dv = (ctypes.c_float * D)(*vals)
        ws = self.workspace(self.lib.b200ddsp_fdn_workspace_bytes(self.handle, float(sampling_rate), B))
        out = torch.empty([B, n], dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            self.check(self.lib.b200ddsp_fdn_ir(
                self.handle, ig.data_ptr(), og.data_ptr(), ga.data_ptr(), da.data_ptr(), t0.data_ptr(),
                al.data_ptr(), er.data_ptr(), E, dv, D, float(sampling_rate), out.data_ptr(), B,
                ws.data_ptr(), ws.numel(), self.stream()))
        return out if batched else out[0]

    def reverb(self, audio, ir, full=False, flags=None):
        audio = self.tensor(audio, 'audio', 2)
        ir = self.tensor(ir, 'ir')
        if ir.dim() == 1:
            ir = ir[None, :]
        if ir.dim() == 3:
            ir = ir[:, :, 0].contiguous()
        B, N = audio.shape
        if ir.shape[0] == 1 and B > 1:
            ir = ir.expand(B, -1).contiguous()
        if ir.shape[0] != B:
            # ddsp.core.fft_convolve's message
            raise ValueError('Batch size of audio ({}) and impulse response ({}) must '
                             'be the same.'.format(B, ir.shape[0]))
        L = ir.shape[1]
        n = 2
        while n < N + L - 1:
            n *= 2
        al = lambda x: (x + 255) // 256 * 256
        ws = self.workspace(al(n * 8 + B * 32) + 2 * al(B * n * 8))
        out = torch.empty([B, N + L - 1] if full else [B, N], dtype=torch.float32,
                          device=self.device)
        with torch.cuda.device(self.device):
            if flags is not None:
                self.check(self.lib.b200ddsp_fft_convolve(
                    self.handle, audio.data_ptr(), ir.data_ptr(), out.data_ptr(), B, N, L, flags,
                    ws.data_ptr(), ws.numel(), self.stream()))
            else:
                fn = self.lib.b200ddsp_reverb_full if full else self.lib.b200ddsp_reverb
                self.check(fn(self.handle, audio.data_ptr(), ir.data_ptr(), out.data_ptr(), B, N, L,
                              ws.data_ptr(), ws.numel(), self.stream()))
        return out

    def _voice_array(self, voices, host):
        """Validate the per-voice control tensors and fill the b200ddsp_voice array.  This runs
        on every forward, ahead of the first kernel launch, so the common case (float32,
        contiguous, right device) is checked without conversions."""
        P = len(voices)
        if P < 1 or P > _lib.MAX_VOICES:
            raise ValueError(f'n_synths={P} outside [1, {_lib.MAX_VOICES}]')
        arr = (_lib.Voice * P)()
        keep = []
        dev = torch.device('cpu') if host else self.device
        f32 = torch.float32
        v0 = voices[0]
        hd0 = v0['harmonic_distribution']
        if not isinstance(hd0, torch.Tensor) or hd0.dim() != 3:
            hd0 = _host_tensor(hd0, 'harmonic_distribution_0', 3) if host else \
                self.tensor(hd0, 'harmonic_distribution_0', 3)
        B, F, H = hd0.shape
        f00 = v0['f0_hz'] if isinstance(v0['f0_hz'], torch.Tensor) else torch.as_tensor(v0['f0_hz'])
        m0 = v0['magnitudes'] if isinstance(v0['magnitudes'], torch.Tensor) else \
            torch.as_tensor(v0['magnitudes'])
        S, M = f00.shape[-1], m0.shape[-1]
        N = F * self.upsampling
        want = (('amplitudes', (B, F, 1)), ('harmonic_distribution', (B, F, H)),
                ('inharm_coef', (B, F, 1)), ('f0_hz', (B, F, S)), ('magnitudes', (B, F, M)))
        any_noise = False
        for i, v in enumerate(voices):
            ptrs = []
            for k, shape in want:
                t = v[k]
                if not (isinstance(t, torch.Tensor) and t.dtype is f32 and t.device == dev
                        and t.is_contiguous()):
                    t = _host_tensor(t, f'{k}_{i}') if host else self.tensor(t, f'{k}_{i}')
                if tuple(t.shape) != shape:
                    raise ValueError(f'{k}_{i} has shape {tuple(t.shape)}, expected {shape}')
                keep.append(t)
                ptrs.append(t.data_ptr())
            nz = v.get('noise')
            if nz is not None:
                nz = _host_tensor(nz, f'noise_{i}', 2) if host else self.tensor(nz, f'noise_{i}', 2)
                if tuple(nz.shape) != (B, N):
                    raise ValueError(f'noise_{i} has shape {tuple(nz.shape)}, expected {(B, N)}')
                keep.append(nz)
                any_noise = True
            arr[i] = _lib.Voice(*ptrs, nz.data_ptr() if nz is not None else None)
        return arr, keep, (P, B, F, H, S, M, N), any_noise

    def _impulse_response(self, reverb_ir, B, host):
        ir = _host_tensor(reverb_ir, 'reverb_ir') if host else self.tensor(reverb_ir, 'reverb_ir')
        if ir.dim() == 1:
            ir = ir[None, :]
        if ir.dim() == 3:
            ir = ir[:, :, 0].contiguous()
        if ir.shape[0] == 1 and B > 1:
            ir = ir.expand(B, -1).contiguous()
        if ir.shape[0] != B:
            raise ValueError('Batch size of audio ({}) and impulse response ({}) must '
                             'be the same.'.format(B, ir.shape[0]))
        return ir

    def forward_polyphonic_host(self, voices, reverb_ir=None, seed=0, want_dry=True):
        """Same as forward_polyphonic for HOST tensors (pinned memory recommended): the H2D
        copies are staged group by group and overlap the kernels; returns pinned host tensors
        (dry or None, wet or None) that are valid once the current stream has been synchronised
        and are REUSED by the next call on this engine."""
        arr, keep, (P, B, F, H, S, M, N), any_noise = self._voice_array(voices, host=True)
        L = 0
        ir = None
        if reverb_ir is not None:
            ir = self._impulse_response(reverb_ir, B, host=True)
            L = ir.shape[1]
        key = ('host', P, B, F, H, S, M, L, any_noise)
        nbytes = self._ws_sizes.get(key)
        if nbytes is None:
            nbytes = self._ws_sizes[key] = self.lib.b200ddsp_workspace_bytes_host(
                self.handle, P, B, F, H, S, M, L, int(any_noise))
        ws = self.workspace(nbytes)
        bufs = self._host_out.get((B, N))
        if bufs is None:
            bufs = self._host_out[(B, N)] = (torch.empty([B, N], dtype=torch.float32).pin_memory(),
                                             torch.empty([B, N], dtype=torch.float32).pin_memory())
        dry = bufs[0] if (want_dry or ir is None) else None
        wet = bufs[1] if ir is not None else None
        self._keep_alive = (keep, ir)        # host buffers must outlive the enqueued copies
        with torch.cuda.device(self.device):
            self.check(self.lib.b200ddsp_forward_polyphonic_host(
                self.handle, arr, P, ir.data_ptr() if ir is not None else None,
                dry.data_ptr() if dry is not None else None,
                wet.data_ptr() if wet is not None else None, B, F, H, S, M, L, seed,
                ws.data_ptr(), ws.numel(), self.stream()))
        return dry, wet

    def forward_polyphonic(self, voices, reverb_ir=None, seed=0):
        """voices: list of dicts with keys amplitudes, harmonic_distribution, inharm_coef,
        f0_hz, magnitudes and optionally noise.  Returns (dry, wet or None)."""
        arr, keep, (P, B, F, H, S, M, N), _ = self._voice_array(voices, host=False)
        L = 0
        ir = None
        if reverb_ir is not None:
            ir = self._impulse_response(reverb_ir, B, host=False)
            L = ir.shape[1]
        key = ('dev', P, B, F, H, S, M, L)
        nbytes = self._ws_sizes.get(key)
        if nbytes is None:
            nbytes = self._ws_sizes[key] = self.lib.b200ddsp_workspace_bytes(
                self.handle, P, B, F, H, S, M, L)
        ws = self.workspace(nbytes)
        dry = torch.empty([B, N], dtype=torch.float32, device=self.device)
        wet = torch.empty_like(dry) if ir is not None else None
        with torch.cuda.device(self.device):
            self.check(self.lib.b200ddsp_forward_polyphonic(
                self.handle, arr, P, ir.data_ptr() if ir is not None else None, dry.data_ptr(),
                wet.data_ptr() if wet is not None else None, B, F, H, S, M, L, seed,
                ws.data_ptr(), ws.numel(), self.stream()))
        return dry, wet


def _host_tensor(x, name, ndim=None):
    if not isinstance(x, torch.Tensor):
        x = torch.as_tensor(x, dtype=torch.float32)
    if x.device.type != 'cpu':
        raise ValueError(f'{name} is on {x.device}; the host entry point takes CPU tensors')
    if x.dtype != torch.float32:
        x = x.to(torch.float32)
    if not x.is_contiguous():
        x = x.contiguous()
    if ndim is not None and x.dim() != ndim:
        raise ValueError(f'{name} must have {ndim} dimensions, got shape {tuple(x.shape)}')
    return x


def get_engine(device, **cfg):
    """Engines are cached per (device, configuration)."""
    device = torch.device(device)
    if device.type != 'cuda':
        raise RuntimeError(f'b200ddsp processors run on CUDA tensors only (got {device}); '
                           'there is no CPU fallback')
    if device.index is None:
        device = torch.device('cuda', torch.cuda.current_device())
    key = (str(device),) + tuple(sorted(cfg.items()))
    with _LOCK:
        eng = _ENGINES.get(key)
        if eng is None:
            eng = _ENGINES[key] = Engine(device, **cfg)
        return eng


def total_launches():
    return sum(e.launch_count() for e in _ENGINES.values())
