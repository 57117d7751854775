/*
 * b200ddsp.h -- C ABI of libb200ddsp.so: the DDSP-Piano per-sample synthesis hot path
 * (inharmonic additive oscillator bank, filtered-noise synth, convolution reverb) as
 * hand-written CUDA kernels for sm_100a (NVIDIA B200).
 *
 * The reference (lrenault/ddsp-piano) is pure Python/TensorFlow and has no FFI for this
 * path; the "binding" it would use is a Python processor class whose get_controls /
 * get_signal forward to these entry points through ctypes.  Each
 * entry point cites the reference interface it replaces (paths under
 * /root/reference/ddsp_piano/).
 *
 * Conventions
 *   - plain C types only; every tensor is a DEVICE pointer to contiguous row-major
 *     float32, 16-byte aligned, owned by the caller (allocated e.g. by torch);
 *   - B batch (clips), F control frames (250 Hz), U = sample_rate / frame_rate,
 *     N = F*U samples, H partials, S substrings per voice, M noise bands, L reverb taps,
 *     P voices;
 *   - every call only enqueues work on `stream` (a cudaStream_t passed as void*) and
 *     returns; nothing allocates device memory after b200ddsp_create(), so calls are
 *     CUDA-graph capturable; scratch comes from the caller's `workspace`;
 *   - return value 0 = OK, negative = b200ddsp_status; text via b200ddsp_last_error();
 *   - a handle may be used by one host thread at a time; handles are independent; one
 *     handle per device (multi-GPU = one process and one handle per GPU).
 */
#ifndef B200DDSP_H_
#define B200DDSP_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200DDSP_VERSION 100 /* 0.1.0 */
#define B200DDSP_MAX_VOICES 64

typedef enum {
  B200DDSP_OK = 0,
  B200DDSP_BAD_SHAPE = -1,
  B200DDSP_BAD_ALIGN = -2,
  B200DDSP_UNSUPPORTED_CONFIG = -3,
  B200DDSP_CUDA_ERROR = -4,
  B200DDSP_WORKSPACE_TOO_SMALL = -5,
  B200DDSP_BAD_ARGUMENT = -6
} b200ddsp_status;

/* scale_fn ids: ddsp.core.exp_sigmoid (default, modules/inharm_synth.py:149),
 * exp_tanh (modules/inharm_synth.py:13-17), or scale_fn=None. */
typedef enum {
  B200DDSP_SCALE_EXP_SIGMOID = 0,
  B200DDSP_SCALE_EXP_TANH = 1,
  B200DDSP_SCALE_NONE = 2
} b200ddsp_scale_fn;

/* Mirrors the constructor arguments of the three processors of configs/dafx22.gin:91-111:
 * MultiInharmonic (modules/inharm_synth.py:145-153,251-252), DynamicSizeFilteredNoise
 * (modules/filtered_noise_synth.py:18-21 + ddsp.synths.FilteredNoise) and
 * ddsp.effects.Reverb. */
typedef struct {
  int sample_rate;                 /* Hz; U = sample_rate / frame_rate (inharm_synth.py:163-165) */
  int frame_rate;                  /* Hz, 250 */
  float min_frequency;             /* 20: voices with f0 <= this are muted (inharm_synth.py:207-208) */
  int additive_scale_fn;           /* b200ddsp_scale_fn */
  int normalize_after_nyquist_cut; /* inharm_synth.py:210-214 (default 1); 0: normalise BEFORE the cut instead
                                      (:194-198); 2: never (SurrogateAdditive(normalize_harm_distribution=
                                      False), modules/surrogate_synth.py:183-187) */
  int normalize_below_nyquist;     /* inharm_synth.py:200-208 (default 1) */
  int inference;                   /* 1: ddsp angular_cumsum (chunks of 1000); 0: plain cumsum over the
                                      clip (inharm_synth.py:73-77), one serial float32 chain per
                                      oscillator (whole clips only, no spans) */
  int noise_scale_fn;              /* b200ddsp_scale_fn (FilteredNoise.scale_fn) */
  float noise_initial_bias;        /* -5.0 */
  int noise_window_size;           /* 257 */
  int reverb_add_dry;              /* effects.Reverb add_dry (default 1) */
  int n_noise_bands;               /* M: the FIR window / cosine tables are built for this M */
  int fast_phase;                  /* 0 (default): bit-faithful float32 phase accumulation in the
                                      reference's summation order (two passes over the phase chain).
                                      1: the phase at the start of every 256-sample synthesis unit is
                                      evaluated in closed form in double precision from the frame-rate
                                      controls (no phase pass, ~20 % less time per forward); closer to
                                      the exact phase than the reference's float32 cumsum and
                                      therefore up to ~1e-2 rad away from it in the highest partials:
                                      NOT within 1e-4 of the reference, validated against the same signal
                                      model in float64 instead (tests/test_gpu_parity.py).
                                      inference = 1, fast additive path, whole clips only. */
} b200ddsp_config;

typedef struct b200ddsp_handle b200ddsp_handle;

/* Device pointers of one voice's RAW (pre-get_controls) control tensors, i.e. the
 * features `amplitudes_i, harmonic_distribution_i, inharm_coef_i, f0_hz_i, magnitudes_i`
 * that Parallelizer.unparallelize produces (modules/sub_modules.py:589-596). */
typedef struct {
  const float* amplitudes;            /* [B, F, 1] */
  const float* harmonic_distribution; /* [B, F, H] */
  const float* inharm_coef;           /* [B, F, 1] */
  const float* f0_hz;                 /* [B, F, S] */
  const float* magnitudes;            /* [B, F, M] */
  const float* noise;                 /* [B, N] uniform [-1,1) samples to filter, or NULL to draw
                                         them in-kernel with Philox (filtered_noise_synth.py:39-40
                                         draws them unseeded) */
} b200ddsp_voice;

int b200ddsp_version(void);

/* Message for the last failure on this handle (or of b200ddsp_create when h == NULL). */
const char* b200ddsp_last_error(const b200ddsp_handle* h);

/* Builds the per-configuration tables (Hann tables, noise-FIR cosine table, FFT twiddles)
 * on the current CUDA device.  The only place that allocates device memory. */
int b200ddsp_create(const b200ddsp_config* cfg, b200ddsp_handle** out);
int b200ddsp_destroy(b200ddsp_handle* h);

/* Scratch bytes needed by the calls below for these shapes (take the max over the calls
 * you make).  P = voices per b200ddsp_forward_polyphonic call (1 for per-voice calls). */
size_t b200ddsp_workspace_bytes(const b200ddsp_handle* h, int P, int B, int F, int H, int S,
                                int M, int L);

/* Scratch bytes of b200ddsp_additive_signal alone (chunk offsets + liveness tables). */
size_t b200ddsp_additive_workspace_bytes(const b200ddsp_handle* h, int B, int F, int H, int S);

/* MultiInharmonic.get_controls -- modules/inharm_synth.py:254-270 over :167-219.
 * rows = B (or P*B for a stacked [P,B,...] tensor).  f0_hz is passed through unchanged by
 * the reference, so it has no output.  Outputs: amplitudes [rows,F,1],
 * harmonic_distribution [rows,F,H], harmonic_shifts [rows,F,H]. */
int b200ddsp_additive_controls(b200ddsp_handle* h, const float* amplitudes,
                               const float* harmonic_distribution, const float* inharm_coef,
                               const float* f0_hz, float* amplitudes_out,
                               float* harmonic_distribution_out, float* harmonic_shifts_out,
                               int rows, int F, int H, int S, void* stream);

/* MultiInharmonic.get_signal -- modules/inharm_synth.py:272-293 (harmonic_synthesis :87-127,
 * cos_oscillator_bank :49-84).  Inputs are get_controls outputs.  out [B, N]; if
 * accumulate != 0 the result is added to `out` (the MultiAdd node, inharm_synth.py:296-309). */
int b200ddsp_additive_signal(b200ddsp_handle* h, const float* amplitudes,
                             const float* harmonic_distribution, const float* harmonic_shifts,
                             const float* f0_hz, float* out, int B, int F, int H, int S,
                             int accumulate, void* workspace, size_t workspace_bytes,
                             void* stream);

/* ddsp.synths.FilteredNoise.get_controls (base class of
 * modules/filtered_noise_synth.py:13): scale_fn(magnitudes + initial_bias). n = element count. */
int b200ddsp_noise_controls(b200ddsp_handle* h, const float* magnitudes, float* magnitudes_out,
                            size_t n, void* stream);

/* DynamicSizeFilteredNoise.get_signal -- modules/filtered_noise_synth.py:27-42:
 * time-varying linear-phase FIR (ddsp.core.frequency_filter) over uniform noise.
 * magnitudes [B,F,M] are get_controls outputs.  noise [B,N] or NULL (then Philox4x32-10
 * keyed by (seed, stream_id)).  out [B,N], added to when accumulate != 0. */
int b200ddsp_noise_signal(b200ddsp_handle* h, const float* magnitudes, const float* noise,
                          uint64_t seed, uint64_t stream_id, float* out, int B, int F, int M,
                          int accumulate, void* workspace, size_t workspace_bytes, void* stream);
/* Scratch bytes of b200ddsp_noise_signal (the per-frame FIR taps). */
size_t b200ddsp_noise_workspace_bytes(const b200ddsp_handle* h, int B, int F, int M);

/* ddsp.effects.Reverb.get_signal (configs/dafx22.gin:99-100,111; ir producer
 * modules/sub_modules.py:351-365): out = conv(audio, ir with ir[:,0]=0)[:, :N] (+ audio if
 * add_dry).  audio [B,N], ir [B,L], out [B,N]; out may not alias audio. */
int b200ddsp_reverb(b200ddsp_handle* h, const float* audio, const float* ir, float* out, int B,
                    int N, int L, void* workspace, size_t workspace_bytes, void* stream);

/* The same convolution with padding='valid': out_full [B, N + L - 1] = conv(audio, ir with
 * ir[:,0]=0), no dry signal.  Building block of the multi-GPU timeline reverb, where the last
 * L-1 samples of a segment are overlap-added into its successors (ddsp.core.fft_convolve with
 * padding='valid', delay_compensation=0). */
int b200ddsp_reverb_full(b200ddsp_handle* h, const float* audio, const float* ir, float* out_full,
                         int B, int N, int L, void* workspace, size_t workspace_bytes, void* stream);

/* MultiInstrumentReverb.exponential_decay_mask -- modules/sub_modules.py:339-349 (applied to the
 * impulse response when the model is built with inference=True, :361-363): out[b, i] = ir[b, i] for
 * i < decay_start, ir[b, i] * exp(-decay_exponent * t) after, t = linspace(0, 1, L - decay_start).
 * The reference hard-codes decay_exponent = 4, decay_start = 16000.  out may alias ir. */
int b200ddsp_ir_decay_mask(b200ddsp_handle* h, const float* ir, float* out, int B, int L,
                           float decay_exponent, int decay_start, void* stream);

/* SurrogateAdditive -- modules/surrogate_synth.py:107-214 (configs/surrogate.gin:106): the inharmonic
 * bank of ONE string (f0_hz [B, F, 1]) whose partial amplitudes are multiplied by |decay|^t, t = samples
 * since the frame-rate decay_time origin (surrogate_harmonic_synthesis :78-97).  get_controls =
 * b200ddsp_additive_controls with S = 1 (no pre-normalisation) for amplitudes / harmonic_distribution /
 * harmonic_shifts, plus b200ddsp_surrogate_decays (clip to [1e-5, 1], 1 above Nyquist, :163-171);
 * get_signal = b200ddsp_surrogate_signal (decays [B, F, H], decay_time [B, F, 1]; workspace of
 * b200ddsp_additive_workspace_bytes(h, B, F, H, 1)).  Both inference modes (:212). */
int b200ddsp_surrogate_decays(b200ddsp_handle* h, const float* decays, const float* inharm_coef,
                              const float* f0_hz, float* decays_out, int B, int F, int H, void* stream);
int b200ddsp_surrogate_signal(b200ddsp_handle* h, const float* amplitudes, const float* decays,
                              const float* decay_time, const float* harmonic_distribution,
                              const float* harmonic_shifts, const float* f0_hz, float* out, int B, int F,
                              int H, void* workspace, size_t workspace_bytes, void* stream);

/* Peer-visible device buffers for the hand-off between the spans of a timeline (below): plain
 * cudaMalloc (zero-filled) + a CUDA IPC handle (64 bytes) that another process on the same node maps
 * with b200ddsp_peer_open (cudaIpcOpenMemHandle, peer access over NVLink). */
int b200ddsp_peer_alloc(b200ddsp_handle* h, size_t bytes, void** dev_ptr, unsigned char* ipc_handle64);
int b200ddsp_peer_free(b200ddsp_handle* h, void* dev_ptr);
int b200ddsp_peer_open(b200ddsp_handle* h, const unsigned char* ipc_handle64, void** peer_ptr);
int b200ddsp_peer_close(b200ddsp_handle* h, void* peer_ptr);

/* ---- one timeline cut into spans (BASELINE config 4) -----------------------------------
 * The reference synthesises a whole piece in ONE pass (synthesize_midi_file.py:52-54,73: %duration is
 * rebound to the piece), so the float32 phase of every oscillator (core.angular_cumsum,
 * modules/inharm_synth.py:73-75), both resamplers (:117-119) and the noise FIR
 * (modules/filtered_noise_synth.py:41) run THROUGH every 3 s boundary, and the legacy bilinear
 * coordinate float(i) * float(F / N) is taken on the GLOBAL sample index.  A span call synthesises
 * frames [out_first_frame, out_first_frame + n_out_frames) of such a timeline, bit for bit what the
 * whole-timeline call produces for those samples, from
 *   - control tensors that cover input frames [in_first_frame, in_first_frame + F) with at least one
 *     frame of halo on either side of the output frames (none is needed where the span touches the
 *     start / the end of the timeline: the reference zero-pads the noise there and holds the last
 *     frame, core.resample add_endpoint);
 *   - the phase state at the span's first sample: per (voice, timeline, substring, partial) the float32
 *     running sum of the chunk-end phases mod 2 pi that angular_cumsum accumulates over the chunks
 *     before the span (zeros at the start of the timeline).  out_first_frame * U must be a multiple of
 *     the 1000-sample chunk.  The call returns the same state at the span's end.
 * Across GPUs the state travels rank to rank in timeline order (float32 addition is not associative,
 * so this is a chain, not a tree), and so do the L - 1 samples of reverb tail that spill into the next
 * span.  Both hand-offs are stream-ordered through peer memory: the producing kernel stores the
 * payload into the successor's inbox (NVLink P2P) and then raises a counter there; the consuming
 * kernel spins on that counter.  No host synchronisation, no NCCL call on the data path. */

/* One rank's end of such a hand-off.  Counters are 64-bit, only grow, and carry the call number
 * `epoch` (1, 2, ...).  Inboxes are double buffered by the caller (slot = epoch & 1); the producer
 * waits for ack >= epoch - 2 before it overwrites a slot.  Any pointer may be NULL: no predecessor
 * (seed = zeros) / no successor / payload already in place (same-process use). */
typedef struct {
  const float* seed;                /* LOCAL inbox slot the predecessor wrote for this epoch */
  unsigned long long* seed_ready;   /* LOCAL counter: >= epoch once `seed` is complete */
  unsigned long long* seed_ack;     /* PEER counter (predecessor's): set to epoch once `seed` was read */
  float* carry;                     /* PEER inbox slot of the successor for this epoch (or local memory) */
  unsigned long long* carry_ready;  /* PEER counter (successor's seed_ready) */
  unsigned long long* carry_ack;    /* LOCAL counter the successor acknowledges into */
  unsigned long long epoch;
  unsigned long long* scratch;      /* LOCAL, 2 words, zero-initialised once: CTA arrival counter + error
                                       word (non-zero after a wait timed out, ~4 s) */
} b200ddsp_link;

typedef struct {
  long long in_first_frame;    /* global index of input frame 0 of the control tensors */
  long long out_first_frame;   /* global index of the first synthesised frame */
  int n_out_frames;            /* frames synthesised: dry_out is [B, n_out_frames * U] */
  long long total_frames;      /* frames of the whole timeline */
  b200ddsp_link phase;         /* payload [P, B, S, H] float32 */
} b200ddsp_span;

/* b200ddsp_forward_polyphonic for one span of B timelines: voices[v] tensors are [B, F, C] over the
 * INPUT frames; an injected noise tensor is [B, F * U] over the same frames.  No reverb node (see
 * b200ddsp_forward_timeline).  Fast additive path only (U % 8 == 0, H <= 128, inference = 1).
 * Workspace: b200ddsp_workspace_bytes(h, P, B, F, H, S, M, 0). */
int b200ddsp_forward_span(b200ddsp_handle* h, const b200ddsp_voice* voices, int P, float* dry_out, int B,
                          int F, int H, int S, int M, uint64_t seed, const b200ddsp_span* span,
                          void* workspace, size_t workspace_bytes, void* stream);

/* The span forward followed by the reverb of the timeline (ddsp.effects.Reverb on the WHOLE piece,
 * configs/dafx22.gin:99-100): the dry span is convolved segment by segment (seg_frames frames each,
 * padding='valid', one FFT block per segment) with the timeline's impulse response reverb_ir [B, L];
 * ONE kernel then overlap-adds the segments, adds the dry signal (add_dry) and the predecessor's tail
 * (`tail` link, payload [B, L - 1]) and hands the L - 1 samples that spill past the span to the
 * successor.  dry_out / wet_out [B, n_out_frames * U] (dry_out may be NULL).  n_out_frames must be a
 * multiple of seg_frames and L - 1 <= n_out_frames * U.
 * Workspace: b200ddsp_timeline_workspace_bytes(). */
int b200ddsp_forward_timeline(b200ddsp_handle* h, const b200ddsp_voice* voices, int P,
                              const float* reverb_ir, float* dry_out, float* wet_out, int B, int F, int H,
                              int S, int M, int L, int seg_frames, uint64_t seed,
                              const b200ddsp_span* span, const b200ddsp_link* tail, void* workspace,
                              size_t workspace_bytes, void* stream);
/* The same from HOST control tensors / into HOST audio (see b200ddsp_forward_polyphonic_host). */
int b200ddsp_forward_timeline_host(b200ddsp_handle* h, const b200ddsp_voice* voices_host, int P,
                                   const float* reverb_ir_host, float* dry_out_host, float* wet_out_host,
                                   int B, int F, int H, int S, int M, int L, int seg_frames, uint64_t seed,
                                   const b200ddsp_span* span, const b200ddsp_link* tail, void* workspace,
                                   size_t workspace_bytes, void* stream);
size_t b200ddsp_timeline_workspace_bytes(const b200ddsp_handle* h, int P, int B, int F, int H, int S, int M,
                                         int L, int n_out_frames, int seg_frames, int host_inputs,
                                         int with_noise);

/* The reverb stage of b200ddsp_forward_timeline alone: dry [B, n_seg * N] device -> out. */
int b200ddsp_timeline_reverb(b200ddsp_handle* h, const float* dry, const float* reverb_ir, float* out,
                             int B, int n_seg, int N, int L, const b200ddsp_link* tail, void* workspace,
                             size_t workspace_bytes, void* stream);
size_t b200ddsp_timeline_reverb_workspace_bytes(const b200ddsp_handle* h, int B, int n_seg, int N, int L);

/* NoteRelease -- modules/sub_modules.py:1174-1188 (tfkl.RNN over F0ProcessorCell, :1114-1171): holds
 * each voice's last played MIDI note for `release_frames` (= release_duration * frame_rate; the
 * shipped dafx22 weights have release_duration = 1 s) frames after its note-off, so that the partial
 * frequencies stay defined while the note rings out.  active_pitch is read with a stride of
 * in_stride floats between frames (2 for the pitch column of conditioning [rows, F, 2], :1184);
 * extended_pitch is [rows, F].  Control-rate helper of the caller row SURVEY 8f-1. */
int b200ddsp_note_release(b200ddsp_handle* h, const float* active_pitch, float* extended_pitch,
                          int rows, int F, int in_stride, float release_frames, void* stream);

/* The time-sequential part of tf.keras.layers.GRU(units, return_sequences=True) with TF2's default
 * reset_after=True, zero initial state (the GRUs of ContextNetwork / MonophonicNetwork, reference
 * configs/dafx22.gin:64-72, 77-86, and of their v2 forms, modules/sub_modules.py:121, 503), all F frames in ONE launch: clusters of
 * CTAs hold the recurrent weights in shared memory and exchange the hidden state through distributed
 * shared memory.  x_proj [rows, F, 3 * units] = x W_i + b_i (the caller's GEMM); w_hh [3 * units, units];
 * b_hh [3 * units]; gates in the order (r, z, n), i.e. Keras' (z, r, h) columns swapped as for
 * torch.nn.GRU; out [rows, F, units].  units in {64, 128, 192, 256}. */
int b200ddsp_gru_recurrence(b200ddsp_handle* h, const float* x_proj, const float* w_hh,
                            const float* b_hh, float* out, int rows, int F, int units, void* stream);

/* ddsp.core.fft_convolve(audio, ir, padding, delay_compensation=0) with the options the reverbs
 * of the reference use.  flags: B200DDSP_CONV_MASK_IR0 zeroes ir[:,0] (effects.Reverb masks the dry
 * tap), B200DDSP_CONV_ADD_DRY adds the input (effects.Reverb add_dry), B200DDSP_CONV_FULL writes
 * N + L - 1 samples (padding='valid') instead of N.  flags = 0 is FeedbackDelayNetwork.get_signal
 * (modules/fdn_reverb.py:406-410). */
#define B200DDSP_CONV_MASK_IR0 1
#define B200DDSP_CONV_ADD_DRY 2
#define B200DDSP_CONV_FULL 4
int b200ddsp_fft_convolve(b200ddsp_handle* h, const float* audio, const float* ir, float* out, int B,
                          int N, int L, int flags, void* workspace, size_t workspace_bytes,
                          void* stream);

/* FeedbackDelayNetwork.get_ir -- modules/fdn_reverb.py:339-360 over get_late_ir :178-337: the
 * impulse response of the feedback delay network that modules/sub_modules.py:368-446
 * (MultiInstrumentFeedbackDelayReverb, configs/maestro-v2.gin:118-122) feeds to effects.Reverb, and
 * that configs/ENSTDkCl-*.gin:99-100,118-122 put at the end of the DAG itself.  delay_lines = D: 8
 * (fdn_reverb.py:30) or 6 (ENSTDkCl gins).  All parameters are device tensors, one row per batch
 * element: input_gain, output_gain [B,D], gain_allpass, delays_allpass [B,D,4], time_rev_0_sec,
 * alpha_tone [B], early_ir [B,E].  delay_values: HOST array of D delay-line lengths, or NULL (D = 8
 * only) for the reference's fixed values (fdn_reverb.py:96).  ir_out [B, n], n = (int)(2 *
 * sampling_rate).  workspace >= b200ddsp_fdn_workspace_bytes(). */
int b200ddsp_fdn_ir(b200ddsp_handle* h, const float* input_gain, const float* output_gain,
                    const float* gain_allpass, const float* delays_allpass,
                    const float* time_rev_0_sec, const float* alpha_tone, const float* early_ir, int E,
                    const float* delay_values, int delay_lines, float sampling_rate, float* ir_out,
                    int B, void* workspace, size_t workspace_bytes, void* stream);
size_t b200ddsp_fdn_workspace_bytes(const b200ddsp_handle* h, float sampling_rate, int B);

/* The whole DAG of modules/polyphonic_dag.py:21-42 as wired by configs/dafx22.gin:91-100,
 * entered at modules/piano_model.py:160: for every voice get_controls + get_signal of the
 * additive and noise processors, the running MultiAdd sum, then the reverb.
 * voices: HOST array of P structs of device pointers.  reverb_ir [B,L] or NULL (no reverb
 * node).  dry_out [B,N] = outputs['add']['signal']; wet_out [B,N] = outputs['reverb']['signal']
 * (ignored when reverb_ir is NULL). */
int b200ddsp_forward_polyphonic(b200ddsp_handle* h, const b200ddsp_voice* voices, int P,
                                const float* reverb_ir, float* dry_out, float* wet_out, int B,
                                int F, int H, int S, int M, int L, uint64_t seed,
                                void* workspace, size_t workspace_bytes, void* stream);

/* The same DAG fed from HOST memory -- what modules/piano_model.py:160 looks like to a caller
 * whose control tensors live on the host.  Every pointer in `voices_host`, `reverb_ir_host`,
 * `dry_out_host` and `wet_out_host` is a HOST pointer (page-locked memory for asynchronous,
 * full-speed copies); either output may be NULL.  The library stages the inputs in `workspace`
 * (device memory) on an internal copy stream, in voice groups, so that the H2D copies of one
 * group overlap the kernels of the previous one; the result is copied back on `stream`.  The
 * call only enqueues; the host buffers must stay valid until `stream` has drained. */
int b200ddsp_forward_polyphonic_host(b200ddsp_handle* h, const b200ddsp_voice* voices_host, int P,
                                     const float* reverb_ir_host, float* dry_out_host,
                                     float* wet_out_host, int B, int F, int H, int S, int M, int L,
                                     uint64_t seed, void* workspace, size_t workspace_bytes,
                                     void* stream);
size_t b200ddsp_workspace_bytes_host(const b200ddsp_handle* h, int P, int B, int F, int H, int S,
                                     int M, int L, int with_noise);

/* HOST function (no GPU work): MIDIRoll2Conditioning.__call__ of a fresh object --
 * utils/midi_encoders.py:33-104, called from utils/io_utils.py:115-116.  roll [n_frames, n_pitches,
 * 2] = stacked (activity, onset velocity) pianorolls at the control rate (n_pitches = 88,
 * first_pitch = 21 in the reference); conditioning [n_frames, n_synths, 2] = (pitch, velocity) per
 * polyphonic channel, a sounding note keeping its channel; polyphony [n_frames] (may be NULL) =
 * number of active notes before the reduction to n_synths.  All pointers are HOST pointers. */
int b200ddsp_midi_roll_to_conditioning(const float* roll, int n_frames, int n_pitches, int n_synths,
                                       float first_pitch, float* conditioning, float* polyphony);

/* Measured FP32 peak for the roofline of the oscillator bank (which is bound by the FMA pipe, not by
 * HBM): runs independent FFMA2 (packed != 0) or scalar FFMA chains on every SM for about a
 * millisecond and returns the achieved FMA-pipe lane-operations per second (one FFMA2 = 2 lane
 * operations; multiply by 2 for FLOP/s).  Timed with CUDA events on `stream`; synchronises it;
 * allocates and frees its own 600 KB -- a measurement, not part of the synthesis path. */
int b200ddsp_measure_fma_rate(b200ddsp_handle* h, int packed, double* lane_ops_per_s, void* stream);

/* Number of kernel launches enqueued by this handle since creation (bench.py's
 * gpu_launches claim is read from here). */
uint64_t b200ddsp_launch_count(const b200ddsp_handle* h);

/* Optional per-stage device timing (bench.py's roofline figures).  When enabled, every call
 * brackets its stages with CUDA events (created in b200ddsp_create) on the caller's stream;
 * b200ddsp_last_stage_ms waits for the events of the most recent call and returns the
 * duration of each stage in milliseconds (0 for stages that did not run).  Not graph
 * capturable while enabled. */
typedef enum {
  B200DDSP_STAGE_CONTROLS = 0,      /* get_controls kernels */
  B200DDSP_STAGE_PHASE_ENDS = 1,    /* additive pass 1: chunk end phases */
  B200DDSP_STAGE_PHASE_SCAN = 2,    /* chunk ends -> chunk offsets */
  B200DDSP_STAGE_OSCILLATORS = 3,   /* additive pass 2: the oscillator bank */
  B200DDSP_STAGE_NOISE = 4,         /* noise taps GEMM + FIR of every voice */
  B200DDSP_STAGE_REVERB = 5,        /* FFT convolution */
  B200DDSP_STAGE_MIX = 6,           /* dry = noise slices + additive partial signals */
  B200DDSP_N_STAGES = 7
} b200ddsp_stage;
int b200ddsp_set_profiling(b200ddsp_handle* h, int enable);
int b200ddsp_last_stage_ms(b200ddsp_handle* h, float* ms /* [B200DDSP_N_STAGES] */);

#ifdef __cplusplus
}
#endif
#endif /* B200DDSP_H_ */
