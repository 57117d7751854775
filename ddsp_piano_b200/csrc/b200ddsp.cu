// libb200ddsp.so -- C ABI (include/b200ddsp.h) over the sm_100a kernels of this directory.
// Host side only: argument validation, workspace carving, launch configuration.  Nothing here
// allocates device memory after b200ddsp_create(); every launch goes to the caller's stream.
#include "../../include/b200ddsp.h"

#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <type_traits>
#include <new>
#include <string>
#include <utility>
#include <vector>

#include "additive.cuh"
#include "additive_fast.cuh"
#include "common.cuh"
#include "controls.cuh"
#include "fdn.cuh"
#include "noise.cuh"
#include "reverb.cuh"
#include "control_rate.cuh"
#include "timeline.cuh"
#include "ubench.cuh"

using namespace b200ddsp;

struct b200ddsp_handle {
  b200ddsp_config cfg;
  int U = 0;
  int device = 0;
  int n_sms = 148;
  float* d_window = nullptr;   // hann(2U)
  float* d_cmat_t = nullptr;   // [M][M-1] noise IR matrix
  bool fast_div = false;       // two-word-reciprocal division == IEEE division for this sample rate
  std::map<std::pair<int, int>, bool> uniform_lerp;   // (F, N) -> floor(float(t)*scale) == t/U
  unsigned long long launches = 0;
  cudaStream_t copy_stream = nullptr;   // H2D staging of the host-input entry point
  cudaStream_t d2h_stream = nullptr;    // the dry signal goes home while the reverb still runs
  cudaEvent_t ev_dry_ready = nullptr, ev_dry_home = nullptr;
  cudaStream_t aux_stream[kMaxGroups - 1] = {};   // the synthesis buckets run concurrently
  cudaEvent_t ev_fork = nullptr, ev_join[kMaxGroups - 1] = {};
  bool offsets_smem_opted_in = false;
  cudaStream_t noise_stream = nullptr;  // the noise synth runs beside the oscillator bank
  cudaEvent_t ev_noise_fork = nullptr, ev_noise_join = nullptr;
  cudaStream_t hd_stream = nullptr;     // harmonic_distribution controls run beside the phase pass
  cudaEvent_t ev_hd_fork = nullptr, ev_hd_done[8] = {}, ev_ir_spectra = nullptr;
  cudaEvent_t ev_group[8] = {};
  cudaEvent_t ev_mags[4] = {}, ev_ir = nullptr, ev_enter = nullptr, ev_small = nullptr;
  bool profiling = false;
  bool trace = false;
  std::vector<cudaEvent_t> trace_ev;      // [0] = entry of the call
  std::vector<std::string> trace_what;
  size_t trace_n = 0;
  cudaEvent_t ev_begin[B200DDSP_N_STAGES] = {};
  cudaEvent_t ev_end[B200DDSP_N_STAGES] = {};
  bool ev_used[B200DDSP_N_STAGES] = {};
  char err[512] = {0};
};

// Brackets one stage with events when profiling is on.
struct StageTimer {
  b200ddsp_handle* h;
  int stage;
  cudaStream_t st;
  StageTimer(b200ddsp_handle* h_, int stage_, cudaStream_t st_) : h(h_), stage(stage_), st(st_) {
    if (h->profiling) {
      cudaEventRecord(h->ev_begin[stage], st);
      h->ev_used[stage] = true;
    }
  }
  ~StageTimer() {
    if (h->profiling) cudaEventRecord(h->ev_end[stage], st);
  }
};

static void reset_stage_flags(b200ddsp_handle* h) {
  for (int i = 0; i < B200DDSP_N_STAGES; ++i) h->ev_used[i] = false;
}

static void trace_launch(b200ddsp_handle* h, const char* what, cudaStream_t stream) {
  if (!h->trace || !h->profiling || h->trace_n >= h->trace_ev.size()) return;
  char tag[96];
  const char* name = stream == h->noise_stream ? "noise" : stream == h->hd_stream ? "side" :
                     stream == h->copy_stream ? "copy" : "main";
  for (int i = 0; i < kMaxGroups - 1; ++i)
    if (stream == h->aux_stream[i]) name = "aux";
  snprintf(tag, sizeof tag, "%-5s %s", name, what);
  cudaEventRecord(h->trace_ev[h->trace_n], stream);
  h->trace_what[h->trace_n++] = tag;
}

static thread_local char g_create_err[512] = {0};

static int fail(b200ddsp_handle* h, int code, const char* fmt, ...) {
  char* dst = h ? h->err : g_create_err;
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(dst, 512, fmt, ap);
  va_end(ap);
  return code;
}

#define CUDA_TRY(h, expr)                                                                    \
  do {                                                                                       \
    cudaError_t e_ = (expr);                                                                 \
    if (e_ != cudaSuccess)                                                                   \
      return fail(h, B200DDSP_CUDA_ERROR, "%s failed: %s", #expr, cudaGetErrorString(e_));   \
  } while (0)

#define CHECK_LAUNCH(h, what)                                                                \
  do {                                                                                       \
    cudaError_t e_ = cudaGetLastError();                                                     \
    if (e_ != cudaSuccess)                                                                   \
      return fail(h, B200DDSP_CUDA_ERROR, "launch of %s failed: %s", what,                   \
                  cudaGetErrorString(e_));                                                   \
    (h)->launches++;                                                                         \
  } while (0)

// Developer trace (B200DDSP_TRACE=1 + profiling on): an event after every launch of the forward path;
// b200ddsp_last_stage_ms prints when each launch COMPLETED relative to the call's entry, by stream.
#define CHECK_LAUNCH_ON(h, what, stream)                                                     \
  do {                                                                                       \
    CHECK_LAUNCH(h, what);                                                                   \
    trace_launch(h, what, stream);                                                           \
  } while (0)

// tuning knobs for experiments (not part of the ABI)
static int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return (v && *v) ? atoi(v) : dflt;
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

static size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// ---------------------------------------------------------------------------------------------
// host-side tables and checks
// ---------------------------------------------------------------------------------------------

// tf.signal.hann_window(n, periodic=True) evaluated in float32 like TF does.
static std::vector<float> hann_periodic(int n) {
  std::vector<float> w(n);
  for (int i = 0; i < n; ++i) {
    const float arg = (float)(2.0 * M_PI) * (float)i / (float)n;
    w[i] = 0.5f - 0.5f * (float)cos((double)arg);
  }
  return w;
}

// Noise FIR as a matrix: tap i = M-1+d (d = 0..M-2) of
//   c[i] = hann(Lir)[i] * irfft(m)[(i + Lir/2) mod Lir]
//        = hann[i]/Lir * ( m_0 + (-1)^(i+M-1) m_{M-1} + 2 sum_{j=1}^{M-2} (-1)^j m_j cos(2 pi j i / Lir) )
// (ddsp.core.frequency_impulse_response + apply_window_to_impulse_response, window_size >= Lir).
// window_size < Lir would crop the IR; the reference never configures that (window_size = 257).
static int cmat_pitch_for(int M) { return (M - 1 + 7) & ~7; }   // rows padded for 16-byte loads

static std::vector<float> noise_ir_matrix_t(int M) {
  const int lir = 2 * (M - 1), nd = M - 1, pitch = cmat_pitch_for(M);
  std::vector<float> hann = hann_periodic(lir);
  std::vector<float> c((size_t)M * pitch, 0.f);
  for (int d = 0; d < nd; ++d) {
    const int i = M - 1 + d;
    const double wgt = (double)hann[i] / (double)lir;
    for (int j = 0; j < M; ++j) {
      double coef;
      if (j == 0) {
        coef = 1.0;
      } else if (j == M - 1) {
        coef = ((i + M - 1) & 1) ? -1.0 : 1.0;
      } else {
        const long long ji = ((long long)j * i) % lir;
        coef = 2.0 * ((j & 1) ? -1.0 : 1.0) * cos(2.0 * M_PI * (double)ji / (double)lir);
      }
      c[(size_t)j * pitch + d] = (float)(wgt * coef);
    }
  }
  return c;
}

// Is fma(x, r, fl(x * r_lo)), with r + r_lo the two-word float32 reciprocal of sr, equal to x / sr
// for every float32 mantissa?  (The identity is scale invariant away from underflow/overflow,
// so one binade suffices.)
static bool verify_fast_division(float sr) {
  const float r = 1.0f / sr;
  const float r_lo = (float)(1.0 / (double)sr - (double)r);
  for (uint32_t m = 0; m < (1u << 23); ++m) {
    const uint32_t bits = (127u << 23) | m;
    float x;
    memcpy(&x, &bits, 4);
    volatile float lo = x * r_lo;
    const float q = fmaf(x, r, lo);
    volatile float want = x / sr;
    if (q != want) return false;
  }
  return true;
}

// Legacy bilinear resize: is floor(float(t) * float(F/N)) == t / U for every output sample, so
// that the lerp frame is the amplitude frame?  (False e.g. when float(1/U) rounds down.)
static bool lerp_is_uniform(b200ddsp_handle* h, int F, int N, int U) {
  auto key = std::make_pair(F, N);
  auto it = h->uniform_lerp.find(key);
  if (it != h->uniform_lerp.end()) return it->second;
  volatile float scale = (float)F / (float)N;
  bool ok = true;
  for (int t = 0; t < N && ok; ++t) {
    volatile float in = (float)t * scale;
    ok = ((int)floorf(in) == t / U);
  }
  h->uniform_lerp[key] = ok;
  return ok;
}

// The generic path still assumes floor(in) is the amplitude frame or the one below it.
static bool lerp_is_supported(int F, int N, int U) {
  volatile float scale = (float)F / (float)N;
  for (int t = 0; t < N; ++t) {
    volatile float in = (float)t * scale;
    const int lo = (int)floorf(in), k = t / U;
    if (lo != k && lo != k - 1) return false;
  }
  return true;
}

// ---------------------------------------------------------------------------------------------
// lifetime
// ---------------------------------------------------------------------------------------------

extern "C" int b200ddsp_version(void) { return B200DDSP_VERSION; }

extern "C" const char* b200ddsp_last_error(const b200ddsp_handle* h) {
  return h ? h->err : g_create_err;
}

extern "C" uint64_t b200ddsp_launch_count(const b200ddsp_handle* h) { return h ? h->launches : 0; }

extern "C" int b200ddsp_create(const b200ddsp_config* cfg, b200ddsp_handle** out) {
  if (!cfg || !out) return fail(nullptr, B200DDSP_BAD_ARGUMENT, "cfg and out must be non-null");
  *out = nullptr;
  if (cfg->sample_rate <= 0 || cfg->frame_rate <= 0 || cfg->sample_rate < cfg->frame_rate)
    return fail(nullptr, B200DDSP_UNSUPPORTED_CONFIG, "bad sample_rate/frame_rate %d/%d",
                cfg->sample_rate, cfg->frame_rate);
  if (cfg->fast_phase != 0 && cfg->fast_phase != 1)
    return fail(nullptr, B200DDSP_UNSUPPORTED_CONFIG, "fast_phase must be 0 or 1");
  if (cfg->fast_phase && !cfg->inference)
    return fail(nullptr, B200DDSP_UNSUPPORTED_CONFIG,
                "fast_phase replaces the chunked angular_cumsum of inference=1; with inference=0 the "
                "reference's single plain cumsum has no chunk structure to start from");
  for (int fn : {cfg->additive_scale_fn, cfg->noise_scale_fn})
    if (fn < 0 || fn > 2) return fail(nullptr, B200DDSP_UNSUPPORTED_CONFIG, "bad scale_fn %d", fn);
  const int M = cfg->n_noise_bands;
  if (M != 0 && (M < 3 || M > 1024))
    return fail(nullptr, B200DDSP_UNSUPPORTED_CONFIG, "n_noise_bands %d out of range", M);
  if (M != 0 && cfg->noise_window_size > 0 && cfg->noise_window_size < 2 * (M - 1))
    return fail(nullptr, B200DDSP_UNSUPPORTED_CONFIG,
                "noise_window_size %d < IR length %d (cropped IR) is not implemented",
                cfg->noise_window_size, 2 * (M - 1));
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess)
    return fail(nullptr, B200DDSP_CUDA_ERROR, "cudaGetDevice: %s", cudaGetErrorString(e));
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, dev);
  if (e != cudaSuccess)
    return fail(nullptr, B200DDSP_CUDA_ERROR, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
  if (prop.major != 10)
    return fail(nullptr, B200DDSP_UNSUPPORTED_CONFIG,
                "device %d is sm_%d%d; this library carries sm_100a code only", dev, prop.major,
                prop.minor);

  b200ddsp_handle* h = new (std::nothrow) b200ddsp_handle();
  if (!h) return fail(nullptr, B200DDSP_CUDA_ERROR, "out of host memory");
  h->cfg = *cfg;
  h->device = dev;
  if (getenv("B200DDSP_TRACE")) {
    h->trace = true;
    h->trace_ev.resize(256);
    h->trace_what.resize(256);
    for (auto& e : h->trace_ev) cudaEventCreate(&e);
  }
  h->n_sms = prop.multiProcessorCount;
  h->U = (int)((double)cfg->sample_rate / (double)cfg->frame_rate);   // inharm_synth.py:163-165
  h->fast_div = verify_fast_division((float)cfg->sample_rate);

  std::vector<float> win = hann_periodic(2 * h->U);
  if (cudaMalloc(&h->d_window, win.size() * sizeof(float)) != cudaSuccess ||
      cudaMemcpy(h->d_window, win.data(), win.size() * sizeof(float), cudaMemcpyHostToDevice) !=
          cudaSuccess) {
    fail(nullptr, B200DDSP_CUDA_ERROR, "window table: %s", cudaGetErrorString(cudaGetLastError()));
    b200ddsp_destroy(h);
    return B200DDSP_CUDA_ERROR;
  }
  if (M != 0) {
    std::vector<float> cm = noise_ir_matrix_t(M);
    if (cudaMalloc(&h->d_cmat_t, cm.size() * sizeof(float)) != cudaSuccess ||
        cudaMemcpy(h->d_cmat_t, cm.data(), cm.size() * sizeof(float), cudaMemcpyHostToDevice) !=
            cudaSuccess) {
      fail(nullptr, B200DDSP_CUDA_ERROR, "noise matrix: %s",
           cudaGetErrorString(cudaGetLastError()));
      b200ddsp_destroy(h);
      return B200DDSP_CUDA_ERROR;
    }
  }
  for (int i = 0; i < B200DDSP_N_STAGES; ++i) {
    if (cudaEventCreate(&h->ev_begin[i]) != cudaSuccess || cudaEventCreate(&h->ev_end[i]) != cudaSuccess) {
      fail(nullptr, B200DDSP_CUDA_ERROR, "event creation: %s", cudaGetErrorString(cudaGetLastError()));
      b200ddsp_destroy(h);
      return B200DDSP_CUDA_ERROR;
    }
  }
  {
    bool ok = cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaStreamCreateWithFlags(&h->d2h_stream, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&h->ev_dry_ready, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&h->ev_dry_home, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming) == cudaSuccess;
    for (int i = 0; i < kMaxGroups - 1 && ok; ++i) {
      ok = cudaStreamCreateWithFlags(&h->aux_stream[i], cudaStreamNonBlocking) == cudaSuccess &&
           cudaEventCreateWithFlags(&h->ev_join[i], cudaEventDisableTiming) == cudaSuccess;
    }
    for (int i = 0; i < 8 && ok; ++i)
      ok = cudaEventCreateWithFlags(&h->ev_group[i], cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaStreamCreateWithFlags(&h->noise_stream, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&h->ev_noise_fork, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&h->ev_noise_join, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaStreamCreateWithFlags(&h->hd_stream, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&h->ev_hd_fork, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&h->ev_ir_spectra, cudaEventDisableTiming) == cudaSuccess;
    for (int i = 0; i < 8 && ok; ++i)
      ok = cudaEventCreateWithFlags(&h->ev_hd_done[i], cudaEventDisableTiming) == cudaSuccess;
    for (int i = 0; i < 4 && ok; ++i)
      ok = cudaEventCreateWithFlags(&h->ev_mags[i], cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&h->ev_ir, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&h->ev_enter, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&h->ev_small, cudaEventDisableTiming) == cudaSuccess;
    if (!ok) {
      fail(nullptr, B200DDSP_CUDA_ERROR, "copy stream: %s", cudaGetErrorString(cudaGetLastError()));
      b200ddsp_destroy(h);
      return B200DDSP_CUDA_ERROR;
    }
  }
  *out = h;
  return B200DDSP_OK;
}

extern "C" int b200ddsp_set_profiling(b200ddsp_handle* h, int enable) {
  if (!h) return B200DDSP_BAD_ARGUMENT;
  h->profiling = enable != 0;
  reset_stage_flags(h);
  return B200DDSP_OK;
}

extern "C" int b200ddsp_last_stage_ms(b200ddsp_handle* h, float* ms) {
  if (!h || !ms) return B200DDSP_BAD_ARGUMENT;
  for (int i = 0; i < B200DDSP_N_STAGES; ++i) {
    ms[i] = 0.f;
    if (!h->ev_used[i]) continue;
    CUDA_TRY(h, cudaEventSynchronize(h->ev_end[i]));
    CUDA_TRY(h, cudaEventElapsedTime(&ms[i], h->ev_begin[i], h->ev_end[i]));
  }
  if (h->trace && h->trace_n > 1) {
    static int dumps = 0;
    if (dumps++ < env_int("B200DDSP_TRACE", 1)) {
      for (size_t i = 1; i < h->trace_n; ++i) {
        float t = 0.f;
        cudaEventSynchronize(h->trace_ev[i]);
        cudaEventElapsedTime(&t, h->trace_ev[0], h->trace_ev[i]);
        fprintf(stderr, "trace %8.1f us  %s\n", t * 1e3f, h->trace_what[i].c_str());
      }
      fprintf(stderr, "trace ----\n");
    }
  }
  return B200DDSP_OK;
}

extern "C" int b200ddsp_destroy(b200ddsp_handle* h) {
  if (!h) return B200DDSP_OK;
  if (h->d_window) cudaFree(h->d_window);
  if (h->d_cmat_t) cudaFree(h->d_cmat_t);
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  if (h->d2h_stream) cudaStreamDestroy(h->d2h_stream);
  if (h->ev_dry_ready) cudaEventDestroy(h->ev_dry_ready);
  if (h->ev_dry_home) cudaEventDestroy(h->ev_dry_home);
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->noise_stream) cudaStreamDestroy(h->noise_stream);
  if (h->ev_noise_fork) cudaEventDestroy(h->ev_noise_fork);
  if (h->ev_noise_join) cudaEventDestroy(h->ev_noise_join);
  if (h->hd_stream) cudaStreamDestroy(h->hd_stream);
  if (h->ev_hd_fork) cudaEventDestroy(h->ev_hd_fork);
  if (h->ev_ir_spectra) cudaEventDestroy(h->ev_ir_spectra);
  for (int i = 0; i < 8; ++i)
    if (h->ev_hd_done[i]) cudaEventDestroy(h->ev_hd_done[i]);
  for (int i = 0; i < kMaxGroups - 1; ++i) {
    if (h->aux_stream[i]) cudaStreamDestroy(h->aux_stream[i]);
    if (h->ev_join[i]) cudaEventDestroy(h->ev_join[i]);
  }
  for (int i = 0; i < 8; ++i)
    if (h->ev_group[i]) cudaEventDestroy(h->ev_group[i]);
  for (int i = 0; i < 4; ++i)
    if (h->ev_mags[i]) cudaEventDestroy(h->ev_mags[i]);
  if (h->ev_ir) cudaEventDestroy(h->ev_ir);
  if (h->ev_enter) cudaEventDestroy(h->ev_enter);
  if (h->ev_small) cudaEventDestroy(h->ev_small);
  for (int i = 0; i < B200DDSP_N_STAGES; ++i) {
    if (h->ev_begin[i]) cudaEventDestroy(h->ev_begin[i]);
    if (h->ev_end[i]) cudaEventDestroy(h->ev_end[i]);
  }
  delete h;
  return B200DDSP_OK;
}

// ---------------------------------------------------------------------------------------------
// workspace
// ---------------------------------------------------------------------------------------------

static int fft_size_for(int N, int L) {
  int n = 2;
  while (n < N + L - 1) n <<= 1;   // ddsp.core.get_fft_size(power_of_2=True), frame = N
  return n;
}

constexpr int kNoiseSlices = 4;   // voice slices of the FIR stage (parallelism + overlap with copies)


// ddsp.core.angular_cumsum wraps the phase every 1000 samples (inference=True); with
// inference=False the reference uses one plain cumsum over the clip (inharm_synth.py:73-77):
// a single "chunk" of N samples.
static int chunk_for(const b200ddsp_handle* h, int N) { return h->cfg.inference ? kAngularChunk : N; }
static int n_chunks_for(const b200ddsp_handle* h, int N) {
  const int c = chunk_for(h, N);
  return (N + c - 1) / c;
}

// Voice groups: split the voices over gridDim.z until the grid fills the GPU.
static int voice_groups_for(int P, int B, int n_chunks) {
  int G = 1;
  while (G < P && (long long)n_chunks * B * G < 4 * 148) G *= 2;
  return G < P ? G : P;
}

static int substrings_per_pass(int S) { return (S % 2 == 0) ? 2 : 1; }

// Synthesis units per chunk on the fast additive path (additive_fast.cuh, kSubLen); the plain cumsum of
// inference=0 has no phase pass that could record the state inside its single chunk.
static int sub_units_for(const b200ddsp_handle* h, int N) {
  if (!h->cfg.inference) return 1;
  const int n = (chunk_for(h, N) + kSubLen - 1) / kSubLen;
  const int forced = env_int("B200DDSP_SUB_UNITS", 0);   // 1 = whole-chunk units (A/B timing)
  return forced == 1 ? 1 : n;
}

// Device buffers of the additive synth for R = P*B rows (byte offsets into a workspace).
struct AdditiveLayout {
  size_t offsets;    // float [R*S, n_chunks, H]   chunk end phases, then chunk offsets
  size_t mids;       // float [R*S, n_chunks, n_sub - 1, H]  phase accumulator at the sub-unit boundaries
  size_t na_frame;   // u8 [R, F]                  live partial groups per frame
  size_t cut_index;  // u16 [R, F]                 partials below Nyquist per frame (prep kernel -> hd kernel)
  size_t synth_na;   // u8 [R, n_chunks]           live partial groups per chunk
  size_t ends_na;    // u8 [R, n_chunks]           groups whose end phase a later chunk needs
  size_t lerp;       // float [N]                  legacy-bilinear lerp weight per sample
  size_t lerp_sums;  // double [F + n_chunks * n_sub]  fast_phase: lerp sums per frame / up to every unit start
  size_t plan;       // AdditivePlan
  size_t lists;      // int [kPlanSlots][kMaxGroups][R * n_chunks]
  size_t chunk_count; // int [kPlanSlots][kMaxGroups][n_chunks]      units per (slot, bucket, chunk)
  size_t chunk_first; // int [kPlanSlots][kMaxGroups][n_chunks + 1]  first work item of every chunk
  size_t partials;   // float [n_partials, B, N]   partial signals for the mixer
  size_t end;
};

// Partial signals: one per (voice, substring set) on the fast path, one per voice group on the
// generic path; sized for the larger of the two so that the choice can be made at run time.
static size_t max_partials(int P, int S) {
  const size_t fast = (size_t)P * (S / substrings_per_pass(S));
  return fast > (size_t)P ? fast : (size_t)P;
}

static AdditiveLayout carve_additive(const b200ddsp_handle* h, size_t at, int P, int B, int F, int H,
                                     int S) {
  AdditiveLayout a{};
  const int U = h->U;
  const size_t R = (size_t)P * B, N = (size_t)F * U;
  const size_t n_chunks = (size_t)n_chunks_for(h, (int)N);
  size_t o = at;
  auto take = [&](size_t bytes) { size_t p = o; o += align_up(bytes); return p; };
  a.offsets = take(R * S * n_chunks * H * 4);
  const size_t n_sub = (size_t)sub_units_for(h, (int)N);
  a.mids = take(R * S * n_chunks * n_sub * H * 4);   // n_sub - 1 per chunk, n_sub with fast_phase
  a.na_frame = take(R * F);
  a.cut_index = take(R * F * 2);
  a.synth_na = take(R * n_chunks);
  a.ends_na = take(R * n_chunks);
  a.lerp = take(N * 4);
  a.lerp_sums = take(((size_t)F + n_chunks * n_sub) * 8);
  // plan and per-chunk unit counts are contiguous: one memset clears both
  a.plan = take(align_up(sizeof(AdditivePlan)) + (size_t)kPlanSlots * kMaxGroups * n_chunks * 4);
  a.chunk_count = a.plan + align_up(sizeof(AdditivePlan));
  a.lists = take((size_t)kPlanSlots * kMaxGroups * R * n_chunks * 4);
  a.chunk_first = take((size_t)kPlanSlots * kMaxGroups * (n_chunks + 1) * 4);
  a.partials = take(max_partials(P, S) * B * N * 4);
  a.end = o;
  return a;
}

// Voices are processed in `n_groups` consecutive groups (1 for device inputs; several for host
// inputs, so that the H2D copies of one group overlap the kernels of the previous one).
struct WorkspaceLayout {
  size_t amp, hd, shifts, f0, noise_part, tw, buf_a, buf_b, total;
  AdditiveLayout add;
  PlanGroups groups;
  int n_chunks, nfft;
};

static WorkspaceLayout carve(const b200ddsp_handle* h, int P, int B, int F, int H, int S, int M, int L,
                             int n_groups) {
  WorkspaceLayout w{};
  const int U = h->U;
  const size_t R = (size_t)P * B, N = (size_t)F * U;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t at = o; o += align_up(bytes); return at; };
  w.n_chunks = n_chunks_for(h, (int)N);
  if (n_groups > P) n_groups = P;
  if (n_groups > kMaxVoiceGroups) n_groups = kMaxVoiceGroups;
  if (n_groups < 1) n_groups = 1;
  w.groups.n_groups = n_groups;
  for (int g = 0; g <= n_groups; ++g) w.groups.first_voice[g] = (int)((long long)P * g / n_groups);
  w.amp = take(R * F * 4);
  w.hd = take(R * F * H * 4);
  w.shifts = take(R * F * H * 4);
  w.f0 = take(R * F * S * 4);
  w.noise_part = take((size_t)kNoiseSlices * B * N * 4);
  w.add = carve_additive(h, o, P, B, F, H, S);
  o = w.add.end;
  if (L > 0) {
    w.nfft = fft_size_for((int)N, L);
    w.tw = take((size_t)w.nfft * 8 + (size_t)B * 32);   // twiddles + per-clip scales + maxima
    // B + 1 rows: the split reverb keeps the IR spectra in the first ceil(B/2) rows and works on the
    // dry signal in the next ceil(B/2)
    w.buf_a = take((size_t)(B + 1) * w.nfft * 8);
    w.buf_b = take((size_t)(B + 1) * w.nfft * 8);
  }
  w.total = o;
  return w;
}

extern "C" size_t b200ddsp_workspace_bytes(const b200ddsp_handle* h, int P, int B, int F, int H,
                                           int S, int M, int L) {
  if (!h || P < 1 || B < 1 || F < 1) return 0;
  return carve(h, P, B, F, H > 0 ? H : 1, S > 0 ? S : 1, M > 1 ? M : 2, L,
               env_int("B200DDSP_DEV_GROUPS", 1)).total;
}

extern "C" size_t b200ddsp_additive_workspace_bytes(const b200ddsp_handle* h, int B, int F, int H,
                                                    int S) {
  if (!h || B < 1 || F < 1 || H < 1 || S < 1) return 0;
  return carve_additive(h, 0, 1, B, F, H, S).end;
}

// ---------------------------------------------------------------------------------------------
// additive
// ---------------------------------------------------------------------------------------------

static int check_common(b200ddsp_handle* h, int B, int F) {
  if (!h) return B200DDSP_BAD_ARGUMENT;
  if (B < 1 || F < 1) return fail(h, B200DDSP_BAD_SHAPE, "B=%d F=%d must be positive", B, F);
  if ((long long)F * h->U > (1 << 24))
    return fail(h, B200DDSP_BAD_SHAPE,
                "N = F*U = %lld exceeds 2^24 samples (float32 sample index of the reference's "
                "resize is no longer exact); split the timeline", (long long)F * h->U);
  if (B > 65535) return fail(h, B200DDSP_BAD_SHAPE, "B=%d exceeds 65535", B);
  return B200DDSP_OK;
}

// get_controls = prep (amplitudes, shifts, f0 copy, liveness) + hd (harmonic distribution)
static void launch_additive_prep(const AdditiveControlsArgs& a, const AdditiveControlsPtrs& p, int P,
                                 cudaStream_t st) {
  const int warps_needed = (a.n_frames_voice + kFramesPerWarp - 1) / kFramesPerWarp;
  dim3 grid((warps_needed + 7) / 8, P);
  switch ((a.H + 31) / 32) {
    case 1: additive_prep_kernel<1><<<grid, 256, 0, st>>>(a, p); break;
    case 2: additive_prep_kernel<2><<<grid, 256, 0, st>>>(a, p); break;
    case 3: additive_prep_kernel<3><<<grid, 256, 0, st>>>(a, p); break;
    case 4: additive_prep_kernel<4><<<grid, 256, 0, st>>>(a, p); break;
    case 5:
    case 6: additive_prep_kernel<6><<<grid, 256, 0, st>>>(a, p); break;
    default: additive_prep_kernel<8><<<grid, 256, 0, st>>>(a, p); break;
  }
}

static void launch_additive_hd(const AdditiveControlsArgs& a, const AdditiveControlsPtrs& p, int P,
                               cudaStream_t st) {
  const int warps_needed = (a.n_frames_voice + kFramesPerWarp - 1) / kFramesPerWarp;
  dim3 grid((warps_needed + 7) / 8, P);
  switch ((a.H + 31) / 32) {
    case 1: additive_hd_kernel<1><<<grid, 256, 0, st>>>(a, p); break;
    case 2: additive_hd_kernel<2><<<grid, 256, 0, st>>>(a, p); break;
    case 3: additive_hd_kernel<3><<<grid, 256, 0, st>>>(a, p); break;
    case 4: additive_hd_kernel<4><<<grid, 256, 0, st>>>(a, p); break;
    case 5:
    case 6: additive_hd_kernel<6><<<grid, 256, 0, st>>>(a, p); break;
    default: additive_hd_kernel<8><<<grid, 256, 0, st>>>(a, p); break;
  }
}

static AdditiveControlsArgs controls_args(const b200ddsp_handle* h, int rows_per_voice_frames, int H,
                                          int S) {
  AdditiveControlsArgs a{};
  a.n_frames_voice = rows_per_voice_frames;
  a.H = H;
  a.S = S;
  a.nyquist = (float)(h->cfg.sample_rate / 2.0);
  a.min_frequency = h->cfg.min_frequency;
  a.scale_fn = h->cfg.additive_scale_fn;
  a.normalize_after = h->cfg.normalize_after_nyquist_cut;
  a.normalize_below = h->cfg.normalize_below_nyquist;
  return a;
}

extern "C" int b200ddsp_additive_controls(b200ddsp_handle* h, const float* amplitudes,
                                          const float* harmonic_distribution,
                                          const float* inharm_coef, const float* f0_hz,
                                          float* amplitudes_out, float* harmonic_distribution_out,
                                          float* harmonic_shifts_out, int rows, int F, int H, int S,
                                          void* stream) {
  if (int rc = check_common(h, rows, F)) return rc;
  if (H < 1 || H > 32 * kMaxHarmonicsPerLane)
    return fail(h, B200DDSP_BAD_SHAPE, "H=%d outside [1, %d]", H, 32 * kMaxHarmonicsPerLane);
  if (S < 1 || S > 32) return fail(h, B200DDSP_BAD_SHAPE, "S=%d outside [1, 32]", S);
  if (!amplitudes || !harmonic_distribution || !inharm_coef || !f0_hz || !amplitudes_out ||
      !harmonic_distribution_out || !harmonic_shifts_out)
    return fail(h, B200DDSP_BAD_ARGUMENT, "null tensor pointer");
  AdditiveControlsPtrs p{};
  p.amp_in[0] = amplitudes;
  p.hd_in[0] = harmonic_distribution;
  p.inharm_in[0] = inharm_coef;
  p.f0_in[0] = f0_hz;
  AdditiveControlsArgs a = controls_args(h, rows * F, H, S);
  a.amp_out = amplitudes_out;
  a.hd_out = harmonic_distribution_out;
  a.shifts_out = harmonic_shifts_out;
  a.f0_out = nullptr;
  a.na_frame = nullptr;
  launch_additive_prep(a, p, 1, (cudaStream_t)stream);
  CHECK_LAUNCH(h, "additive_prep_kernel");
  launch_additive_hd(a, p, 1, (cudaStream_t)stream);
  CHECK_LAUNCH(h, "additive_hd_kernel");
  return B200DDSP_OK;
}

static bool additive_fast_path(b200ddsp_handle* h, int F, int H) {
  const int U = h->U;
  return h->fast_div && (U % 8 == 0) && (chunk_for(h, F * U) % 8 == 0) && H <= 16 * kMaxGroups &&
         lerp_is_uniform(h, F, F * U, U);
}

// A span of a timeline as the kernels see it (b200ddsp_span validated and reduced; null = the call
// covers whole clips).
struct SpanInfo {
  int koff;            // input frame of output frame 0
  int F_out;           // frames synthesised
  int tg0;             // global sample index of output sample 0
  int total_frames;    // frames of the whole timeline
  bool seeded;         // not the start of the timeline: chunk 0 has an offset
  bool carry;          // the phase state at the span's end is wanted
  Link phase;
};

static Link to_link(const b200ddsp_link& l) {
  Link k{};
  k.seed = l.seed; k.seed_ready = l.seed_ready; k.seed_ack = l.seed_ack;
  k.carry = l.carry; k.carry_ready = l.carry_ready; k.carry_ack = l.carry_ack;
  k.epoch = l.epoch; k.scratch = l.scratch;
  return k;
}

// What the mixer needs to know about the additive partial signals.
struct AdditiveResult {
  const float* partials;        // [n_partials, B, N]
  const unsigned char* live;    // [P*B, n_chunks] or nullptr
  int n_partials, sets;
};

// One additive synthesis over stacked controls (R = P*B rows), split in the phases the
// polyphonic forward interleaves with its copies:
//   additive_begin        validate, choose fast/generic path, fill the kernel arguments
//   additive_phase_pass   liveness per chunk, work lists, chunk end phases, chunk offsets
//                         (needs f0, shifts and na_frame only -- not hd/amp)
//   additive_synth_group  the oscillator bank of one voice group -> partial signals
struct AdditiveRun {
  AdditiveArgs a;
  AdditiveFastArgs fa;
  bool fast;
  int sets, n_chunks, P, B, F, H, S, G;
  unsigned char *na_frame, *synth_na, *ends_na;
  unsigned short* cut_index;
  PlanGroups groups;
  float* partials;
  double* frame_sum;
  const SpanInfo* span;
};

static int additive_begin(b200ddsp_handle* h, AdditiveRun* r, const float* amp, const float* hd,
                          const float* shifts, const float* f0, char* base, const AdditiveLayout& lay,
                          int P, int B, int F, int H, int S, const PlanGroups& groups,
                          const float* decays = nullptr, const float* decay_time = nullptr,
                          const SpanInfo* span = nullptr) {
  const int U = h->U, N = (span ? span->F_out : F) * U;
  // legacy-bilinear coordinates are those of the whole timeline
  const int F_all = span ? span->total_frames : F, N_all = F_all * U;
  r->span = span;
  if (H < 1 || H > 256) return fail(h, B200DDSP_BAD_SHAPE, "H=%d outside [1, 256]", H);
  if (S < 1 || S > 32) return fail(h, B200DDSP_BAD_SHAPE, "S=%d outside [1, 32]", S);
  r->fast = additive_fast_path(h, F_all, H) && decays == nullptr;   // the surrogate runs on the generic kernel
  if (span && !r->fast)
    return fail(h, B200DDSP_UNSUPPORTED_CONFIG,
                "spans of a timeline are implemented on the fast additive path only (U %% 8 == 0, "
                "H <= 128, uniform resize coordinates); got U=%d H=%d", U, H);
  if (span && h->cfg.fast_phase)
    return fail(h, B200DDSP_UNSUPPORTED_CONFIG,
                "spans of a timeline carry the reference's float32 phase state: not available with fast_phase");
  if (h->cfg.fast_phase && !r->fast)
    return fail(h, B200DDSP_UNSUPPORTED_CONFIG,
                "fast_phase is implemented on the fast additive path only (U %% 8 == 0, H <= 128); got U=%d H=%d",
                U, H);
  if (span && !h->cfg.inference)
    return fail(h, B200DDSP_UNSUPPORTED_CONFIG,
                "spans of a timeline need inference=1: the plain cumsum of training mode carries an "
                "unbounded float32 phase");
  if (!r->fast && !lerp_is_uniform(h, F, N, U) && !lerp_is_supported(F, N, U))
    return fail(h, B200DDSP_UNSUPPORTED_CONFIG,
                "legacy-bilinear source frame departs from t/U by more than one frame (F=%d N=%d)",
                F, N);
  r->n_chunks = n_chunks_for(h, N);
  if (r->n_chunks > 12 * 1024)
    return fail(h, B200DDSP_BAD_SHAPE, "timeline of %d chunks is too long for one call", r->n_chunks);
  r->P = P; r->B = B; r->F = F; r->H = H; r->S = S;
  r->sets = r->fast ? S / substrings_per_pass(S) : 1;
  r->groups = groups;
  // generic path: voice groups inside one launch (only when the forward itself is not grouped)
  r->G = (groups.n_groups == 1) ? voice_groups_for(P, B, r->n_chunks) : 1;
  r->na_frame = (unsigned char*)(base + lay.na_frame);
  r->cut_index = (unsigned short*)(base + lay.cut_index);
  r->synth_na = (unsigned char*)(base + lay.synth_na);
  r->ends_na = (unsigned char*)(base + lay.ends_na);
  r->partials = (float*)(base + lay.partials);
  r->frame_sum = (double*)(base + lay.lerp_sums);
  AdditiveArgs& a = r->a;
  a = AdditiveArgs{};
  a.amp = amp; a.hd = hd; a.shifts = shifts; a.f0 = f0;
  a.decays = decays; a.decay_time = decay_time;
  a.offsets = (float*)(base + lay.offsets);
  a.mids = (float*)(base + lay.mids);
  a.out = r->partials;
  a.window = h->d_window;
  a.B = B; a.P = P; a.F = F; a.H = H; a.S = S; a.U = U; a.N = N;
  a.chunk = chunk_for(h, N);
  a.n_chunks = r->n_chunks;
  a.n_sub = r->fast ? sub_units_for(h, N) : 1;
  a.fast_phase = h->cfg.fast_phase ? 1 : 0;
  a.voices_per_group = (P + r->G - 1) / r->G;
  a.koff = span ? span->koff : 0;
  a.seeded = (span && span->seeded) ? 1 : 0;
  a.accumulate = 0;
  a.plain = (!r->fast && !h->cfg.inference) ? 1 : 0;
  a.scale = (float)F_all / (float)N_all;
  a.nyquist = (float)(h->cfg.sample_rate / 2.0);
  a.sr = (float)h->cfg.sample_rate;
  a.inv_sr = 1.0f / a.sr;
  a.inv_sr_lo = (float)(1.0 / (double)a.sr - (double)a.inv_sr);
  r->fa = AdditiveFastArgs{};
  r->fa.a = a;
  r->fa.plan = (AdditivePlan*)(base + lay.plan);
  r->fa.lists = (int*)(base + lay.lists);
  r->fa.chunk_count = (int*)(base + lay.chunk_count);
  r->fa.chunk_first = (int*)(base + lay.chunk_first);
  r->fa.lerp = (float*)(base + lay.lerp);
  r->fa.sp = substrings_per_pass(S);
  const size_t smem = (size_t)(((2 * U + 31) & ~31) + kAddWarps * kMaxChunk) * sizeof(float);
  if (!r->fast && smem > 48 * 1024)
    return fail(h, B200DDSP_UNSUPPORTED_CONFIG, "upsampling factor U=%d too large", U);
  return B200DDSP_OK;
}

static AdditiveResult additive_result(const AdditiveRun& r) {
  AdditiveResult res{};
  res.partials = r.partials;
  res.live = r.fast ? r.synth_na : nullptr;
  res.sets = r.sets;
  res.n_partials = r.fast ? r.P * r.sets : (r.groups.n_groups == 1 ? r.G : r.groups.n_groups);
  return res;
}

template <int HP>
static void launch_additive_hp(const AdditiveArgs& a, bool ends_only, dim3 grid, int threads,
                               size_t smem, cudaStream_t st) {
  if (ends_only) additive_kernel<HP, 1, false, true><<<grid, threads, 0, st>>>(a);
  else additive_kernel<HP, 1, false, false><<<grid, threads, smem, st>>>(a);
}

// generic path (any U, per-sample lerp frame, IEEE division)
static void launch_additive_generic(const AdditiveArgs& a, bool ends_only, dim3 grid, cudaStream_t st) {
  const int n_pairs = a.voices_per_group * a.S;
  const int threads = 32 * (n_pairs < kAddWarps ? n_pairs : kAddWarps);
  const size_t smem = (size_t)(((2 * a.U + 31) & ~31) + kAddWarps * kMaxChunk) * sizeof(float);
  switch ((a.H + 31) / 32) {
    case 1: launch_additive_hp<1>(a, ends_only, grid, threads, smem, st); break;
    case 2: launch_additive_hp<2>(a, ends_only, grid, threads, smem, st); break;
    case 3: launch_additive_hp<3>(a, ends_only, grid, threads, smem, st); break;
    case 4: launch_additive_hp<4>(a, ends_only, grid, threads, smem, st); break;
    case 5:
    case 6: launch_additive_hp<6>(a, ends_only, grid, threads, smem, st); break;
    default: launch_additive_hp<8>(a, ends_only, grid, threads, smem, st); break;
  }
}

// the phase pass: persistent kernel, every warp pulls (unit, substring set) items until the list is empty
static void launch_additive_ends(const AdditiveFastArgs& fa, int grid, cudaStream_t st) {
  if (fa.sp == 2) additive_fast_kernel<2, true><<<grid, kAddThreads, 0, st>>>(fa);
  else additive_fast_kernel<1, true><<<grid, kAddThreads, 0, st>>>(fa);
}

template <int NH>
static void launch_synth_bucket(const AdditiveFastArgs& fa, bool plain, int grid, size_t smem,
                                cudaStream_t st) {
  const int threads = kSynthWarps * 32;
  if (fa.sp == 2) {
    if (plain) additive_synth_kernel<NH, 2, true><<<grid, threads, smem, st>>>(fa);
    else additive_synth_kernel<NH, 2, false><<<grid, threads, smem, st>>>(fa);
  } else {
    if (plain) additive_synth_kernel<NH, 1, true><<<grid, threads, smem, st>>>(fa);
    else additive_synth_kernel<NH, 1, false><<<grid, threads, smem, st>>>(fa);
  }
}

// persistent grids: every warp pulls work until its list is empty
static int persistent_grid(const b200ddsp_handle* h, long long max_items, int ctas_per_sm) {
  const long long want = (max_items + kAddWarps - 1) / kAddWarps;
  const long long cap = (long long)h->n_sms * ctas_per_sm;
  return (int)(want < cap ? (want > 0 ? want : 1) : cap);
}

// chunk end phases -> chunk offsets: the whole span in one shared-memory tile when it fits (additive.cuh)
static void launch_offsets(b200ddsp_handle* h, OffsetsArgs& oa, cudaStream_t st) {
  int cap = env_int("B200DDSP_OFFSETS_TILE", kOffTileMax);   // tests force several tiles on short clips
  cap = cap < 1 ? 1 : (cap > kOffTileMax ? kOffTileMax : cap);
  oa.tile_chunks = oa.n_chunks < cap ? (oa.n_chunks > 0 ? oa.n_chunks : 1) : cap;
  const size_t smem = (size_t)oa.tile_chunks * (33 * sizeof(float) + 1);
  if (!h->offsets_smem_opted_in) {   // per handle = per device
    cudaFuncSetAttribute(additive_offsets_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)((size_t)kOffTileMax * (33 * sizeof(float) + 1)));
    h->offsets_smem_opted_in = true;
  }
  const int threads = oa.n_chunks > 128 ? 512 : 256;
  additive_offsets_kernel<<<oa.n_osc_rows * ((oa.H + 31) / 32), threads, smem, st>>>(oa);
}

// small_kernels_done / ends_done: optional events recorded after the work lists are built / after the
// phase pass proper, in front of the scan.
static int additive_phase_pass(b200ddsp_handle* h, AdditiveRun& r, bool na_frame_ready,
                               cudaStream_t st, cudaEvent_t small_kernels_done = nullptr,
                               cudaEvent_t ends_done = nullptr) {
  const AdditiveArgs& a = r.a;
  const int R = r.P * r.B;
  const bool carry = r.span && r.span->carry;
  if (r.fast) {
    {
      StageTimer tm(h, B200DDSP_STAGE_PHASE_SCAN, st);
      if (!na_frame_ready) {
        additive_alive_frames_kernel<<<(R * r.F + 7) / 8, 256, 0, st>>>(a.amp, a.hd, r.na_frame,
                                                                        R * r.F, r.H);
        CHECK_LAUNCH_ON(h, "additive_alive_frames_kernel", st);
      }
      additive_lerp_kernel<<<(a.N + 255) / 256, 256, 0, st>>>((float*)r.fa.lerp, a.N, a.U, a.scale,
                                                               r.span ? r.span->tg0 : 0);
      CHECK_LAUNCH_ON(h, "additive_lerp_kernel", st);
      additive_alive_chunks_kernel<<<R, 128, (size_t)r.n_chunks, st>>>(
          r.na_frame, r.synth_na, r.ends_na, r.F, a.U, a.N, a.chunk, r.n_chunks, a.koff,
          carry ? (r.H + 15) / 16 : 0);
      CHECK_LAUNCH_ON(h, "additive_alive_chunks_kernel", st);
      CUDA_TRY(h, cudaMemsetAsync(r.fa.plan, 0, align_up(sizeof(AdditivePlan)) +
                                                    (size_t)kPlanSlots * kMaxGroups * r.n_chunks * 4, st));
      const int n_units = R * r.n_chunks;
      additive_plan_kernel<<<(n_units + 255) / 256, 256, 0, st>>>(
          r.synth_na, r.ends_na, r.fa.plan, (int*)r.fa.lists, (int*)r.fa.chunk_count, n_units, r.n_chunks, r.B,
          r.groups, carry ? 1 : 0, a.n_sub);
      CHECK_LAUNCH_ON(h, "additive_plan_kernel", st);
      additive_plan_items_kernel<<<(kPlanSlots - 1) * kMaxGroups, 32, 0, st>>>(
          r.fa.chunk_count, (int*)r.fa.chunk_first, r.fa.plan, r.n_chunks, r.fa.sp, h->cfg.inference ? 0 : 1);
      CHECK_LAUNCH_ON(h, "additive_plan_items_kernel", st);
    }
    if (small_kernels_done) {
      CUDA_TRY(h, cudaEventRecord(small_kernels_done, st));
      small_kernels_done = nullptr;
    }
    if (a.fast_phase) {   // closed-form unit start phases instead of pass 1 + scan
      StageTimer tm(h, B200DDSP_STAGE_PHASE_ENDS, st);
      const int n_osc = R * r.S * r.H, n_frames = a.N / a.U, n_units = r.n_chunks * a.n_sub;
      additive_lerp_sums_kernel<<<(n_frames + n_units + 127) / 128, 128, 0, st>>>(
          r.fa.lerp, r.frame_sum, r.frame_sum + n_frames, a.N, a.U, a.chunk, r.n_chunks, a.n_sub);
      CHECK_LAUNCH_ON(h, "additive_lerp_sums_kernel", st);
      (void)n_osc;
      additive_closed_phase_kernel<<<R * r.S * ((r.H + 31) / 32), kClosedSegs * 32, 0, st>>>(a, r.frame_sum,
                                                                                            r.frame_sum + n_frames);
      CHECK_LAUNCH_ON(h, "additive_closed_phase_kernel", st);
      return B200DDSP_OK;
    }
    if (r.n_chunks > 1 || carry || a.n_sub > 1) {
      StageTimer tm(h, B200DDSP_STAGE_PHASE_ENDS, st);
      AdditiveFastArgs fa = r.fa;
      fa.slot = 0;
      launch_additive_ends(fa, persistent_grid(h, (long long)R * r.n_chunks * r.sets, 4), st);
      CHECK_LAUNCH_ON(h, "additive_fast_kernel<ends>", st);
    }
    if (ends_done) CUDA_TRY(h, cudaEventRecord(ends_done, st));
    if (r.n_chunks > 1 || carry || (r.span && r.span->seeded)) {
      if (r.span && r.span->phase.seed && r.span->phase.seed_ready) {
        // the predecessor's state may still be on its way: hold the stream, not the SMs (link.cuh)
        link_gate_kernel<<<1, 32, 0, st>>>(r.span->phase.seed_ready, r.span->phase.epoch, r.span->phase.scratch);
        CHECK_LAUNCH_ON(h, "link_gate_kernel", st);
      }
      OffsetsArgs oa{};
      oa.offsets = a.offsets;
      oa.ends_na = r.ends_na;
      oa.n_osc_rows = R * r.S; oa.n_chunks = r.n_chunks; oa.H = r.H; oa.S = r.S;
      oa.carry_all = carry ? 1 : 0;
      if (r.span) oa.link = r.span->phase;
      launch_offsets(h, oa, st);
      CHECK_LAUNCH_ON(h, "additive_offsets_kernel", st);
    }
    return B200DDSP_OK;
  }
  if (r.n_chunks > 1) {
    {
      StageTimer tm(h, B200DDSP_STAGE_PHASE_ENDS, st);
      launch_additive_generic(a, true, dim3(r.n_chunks - 1, r.B, r.G), st);
      CHECK_LAUNCH_ON(h, "additive_kernel<ends>", st);
    }
    StageTimer tm(h, B200DDSP_STAGE_PHASE_SCAN, st);
    OffsetsArgs oa{};
    oa.offsets = a.offsets;
    oa.n_osc_rows = R * r.S; oa.n_chunks = r.n_chunks; oa.H = r.H; oa.S = r.S;
    launch_offsets(h, oa, st);
    CHECK_LAUNCH_ON(h, "additive_offsets_kernel", st);
  }
  if (small_kernels_done) CUDA_TRY(h, cudaEventRecord(small_kernels_done, st));
  return B200DDSP_OK;
}

static int additive_synth_group(b200ddsp_handle* h, AdditiveRun& r, int g, cudaStream_t st) {
  StageTimer tm(h, B200DDSP_STAGE_OSCILLATORS, st);
  const int v0 = r.groups.first_voice[g], Pg = r.groups.first_voice[g + 1] - v0;
  if (r.fast) {
    AdditiveFastArgs fa = r.fa;
    fa.slot = 1 + g;
    // Hann half-window + one reduction tile per warp
    const size_t smem = (size_t)(((r.a.U + 3) & ~3) + kSynthWarps * kRedTileFloats) * sizeof(float);
    const bool plain = !h->cfg.inference;
    // one kernel per bucket (number of live 16-partial half-groups), heaviest on the caller's
    // stream, the others on auxiliary streams so that the buckets overlap; grids are sized for the
    // largest possible bucket, surplus CTAs exit at once
    const long long max_items = (long long)Pg * r.B * r.n_chunks * r.sets * r.a.n_sub;
    const int grid = (int)((max_items + kSynthWarps - 1) / kSynthWarps);
    const int n_buckets = (r.H + 15) / 16;
    CUDA_TRY(h, cudaEventRecord(h->ev_fork, st));
    // stream of bucket nh: the heaviest on the caller's stream, every other bucket on its own
    // auxiliary stream.  Measured alternatives, all slower: buckets sharing streams (any grouping,
    // +5..40 %), higher stream priority for the small buckets (+10 %) or for the heavy ones (+3 %,
    // and the host-input path loses 25 %).  Run to run the stage still lands in one of two modes
    // (1.03 / 1.10 ms at config 3) depending on how the block scheduler interleaves the launches;
    // one merged launch over all buckets in heaviest-first order is deterministic but takes 1.20-1.30.
    bool used[kMaxGroups] = {};
    for (int nh = n_buckets; nh >= 1; --nh) {
      const int si = (nh == n_buckets) ? 0 : n_buckets - nh;
      cudaStream_t s = (si == 0) ? st : h->aux_stream[si - 1];
      if (si > 0 && !used[si]) {
        CUDA_TRY(h, cudaStreamWaitEvent(s, h->ev_fork, 0));
        used[si] = true;
      }
      switch (nh) {
        case 1: launch_synth_bucket<1>(fa, plain, grid, smem, s); break;
        case 2: launch_synth_bucket<2>(fa, plain, grid, smem, s); break;
        case 3: launch_synth_bucket<3>(fa, plain, grid, smem, s); break;
        case 4: launch_synth_bucket<4>(fa, plain, grid, smem, s); break;
        case 5: launch_synth_bucket<5>(fa, plain, grid, smem, s); break;
        case 6: launch_synth_bucket<6>(fa, plain, grid, smem, s); break;
        case 7: launch_synth_bucket<7>(fa, plain, grid, smem, s); break;
        default: launch_synth_bucket<8>(fa, plain, grid, smem, s); break;
      }
      {
        char what[40];
        snprintf(what, sizeof what, "additive_synth_kernel<%d>", nh);
        CHECK_LAUNCH_ON(h, what, s);
      }
    }
    int aux = 0;
    for (int si = 1; si < kMaxGroups; ++si) {
      if (!used[si]) continue;
      CUDA_TRY(h, cudaEventRecord(h->ev_join[si - 1], h->aux_stream[si - 1]));
      CUDA_TRY(h, cudaStreamWaitEvent(st, h->ev_join[si - 1], 0));
      ++aux;
    }
    (void)aux;
    return B200DDSP_OK;
  }
  // generic kernel over the group's rows; one partial signal per launch-internal voice group
  AdditiveArgs a = r.a;
  const size_t row0 = (size_t)v0 * r.B;
  a.amp += row0 * r.F;
  a.hd += row0 * r.F * r.H;
  a.shifts += row0 * r.F * r.H;
  a.f0 += row0 * r.F * r.S;
  a.offsets += row0 * r.S * r.n_chunks * r.H;
  a.P = Pg;
  const int G = (r.groups.n_groups == 1) ? r.G : 1;
  a.voices_per_group = (Pg + G - 1) / G;
  a.out = r.partials + (size_t)(r.groups.n_groups == 1 ? 0 : g) * r.B * r.a.N;
  if (a.plain && !a.accumulate)   // the plain-cumsum kernel adds into its output (additive.cuh)
    CUDA_TRY(h, cudaMemsetAsync(a.out, 0, (size_t)G * r.B * r.a.N * sizeof(float), st));
  launch_additive_generic(a, false, dim3(r.n_chunks, r.B, G), st);
  CHECK_LAUNCH_ON(h, "additive_kernel<synth>", st);
  return B200DDSP_OK;
}

static int additive_signal_impl(b200ddsp_handle* h, const float* amplitudes,
                                const float* harmonic_distribution, const float* harmonic_shifts,
                                const float* f0_hz, const float* decays, const float* decay_time,
                                float* out, int B, int F, int H, int S, int accumulate, void* workspace,
                                size_t workspace_bytes, void* stream) {
  if (int rc = check_common(h, B, F)) return rc;
  if (!amplitudes || !harmonic_distribution || !harmonic_shifts || !f0_hz || !out)
    return fail(h, B200DDSP_BAD_ARGUMENT, "null tensor pointer");
  if (H < 1 || H > 256) return fail(h, B200DDSP_BAD_SHAPE, "H=%d outside [1, 256]", H);
  if (S < 1 || S > 32) return fail(h, B200DDSP_BAD_SHAPE, "S=%d outside [1, 32]", S);
  const AdditiveLayout lay = carve_additive(h, 0, 1, B, F, H, S);
  if (!workspace || workspace_bytes < lay.end)
    return fail(h, B200DDSP_WORKSPACE_TOO_SMALL, "additive_signal needs %zu workspace bytes, got %zu",
                lay.end, workspace_bytes);
  if (!aligned16(workspace)) return fail(h, B200DDSP_BAD_ALIGN, "workspace must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  reset_stage_flags(h);
  PlanGroups groups{};
  groups.n_groups = 1;
  groups.first_voice[0] = 0;
  groups.first_voice[1] = 1;
  AdditiveRun run;
  if (int rc = additive_begin(h, &run, amplitudes, harmonic_distribution, harmonic_shifts, f0_hz,
                              (char*)workspace, lay, 1, B, F, H, S, groups, decays, decay_time))
    return rc;
  if (int rc = additive_phase_pass(h, run, false, st)) return rc;
  if (int rc = additive_synth_group(h, run, 0, st)) return rc;
  const AdditiveResult res = additive_result(run);
  PartialSumArgs ps{};
  ps.partials = res.partials;
  ps.live = res.live;
  ps.out = out;
  ps.n_partials = res.n_partials;
  ps.sets = res.sets;
  ps.B = B;
  ps.N = F * h->U;
  ps.chunk = chunk_for(h, ps.N);
  ps.n_chunks = n_chunks_for(h, ps.N);
  ps.accumulate = accumulate;
  additive_sum_partials_kernel<<<dim3((ps.N + 255) / 256, B), 256, 0, st>>>(ps);
  CHECK_LAUNCH(h, "additive_sum_partials_kernel");
  return B200DDSP_OK;
}

// ---------------------------------------------------------------------------------------------
// noise
// ---------------------------------------------------------------------------------------------

extern "C" int b200ddsp_additive_signal(b200ddsp_handle* h, const float* amplitudes,
                                        const float* harmonic_distribution,
                                        const float* harmonic_shifts, const float* f0_hz, float* out,
                                        int B, int F, int H, int S, int accumulate, void* workspace,
                                        size_t workspace_bytes, void* stream) {
  return additive_signal_impl(h, amplitudes, harmonic_distribution, harmonic_shifts, f0_hz, nullptr, nullptr,
                              out, B, F, H, S, accumulate, workspace, workspace_bytes, stream);
}

extern "C" int b200ddsp_surrogate_signal(b200ddsp_handle* h, const float* amplitudes, const float* decays,
                                         const float* decay_time, const float* harmonic_distribution,
                                         const float* harmonic_shifts, const float* f0_hz, float* out,
                                         int B, int F, int H, void* workspace, size_t workspace_bytes,
                                         void* stream) {
  if (!h) return B200DDSP_BAD_ARGUMENT;
  if (!decays || !decay_time) return fail(h, B200DDSP_BAD_ARGUMENT, "null decay tensor");
  return additive_signal_impl(h, amplitudes, harmonic_distribution, harmonic_shifts, f0_hz, decays,
                              decay_time, out, B, F, H, 1, 0, workspace, workspace_bytes, stream);
}

extern "C" int b200ddsp_surrogate_decays(b200ddsp_handle* h, const float* decays, const float* inharm_coef,
                                         const float* f0_hz, float* out, int B, int F, int H,
                                         void* stream) {
  if (int rc = check_common(h, B, F)) return rc;
  if (!decays || !inharm_coef || !f0_hz || !out) return fail(h, B200DDSP_BAD_ARGUMENT, "null tensor pointer");
  if (H < 1 || H > 256) return fail(h, B200DDSP_BAD_SHAPE, "H=%d outside [1, 256]", H);
  const size_t n = (size_t)B * F * H;
  surrogate_decays_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      decays, inharm_coef, f0_hz, out, B * F, H, (float)(h->cfg.sample_rate / 2.0));
  CHECK_LAUNCH(h, "surrogate_decays_kernel");
  return B200DDSP_OK;
}

extern "C" int b200ddsp_noise_controls(b200ddsp_handle* h, const float* magnitudes,
                                       float* magnitudes_out, size_t n, void* stream) {
  if (!h) return B200DDSP_BAD_ARGUMENT;
  if (!magnitudes || !magnitudes_out) return fail(h, B200DDSP_BAD_ARGUMENT, "null tensor pointer");
  if (n == 0) return B200DDSP_OK;
  const int blocks = (int)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
  noise_controls_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(
      magnitudes, magnitudes_out, n, h->cfg.noise_initial_bias, h->cfg.noise_scale_fn);
  CHECK_LAUNCH(h, "noise_controls_kernel");
  return B200DDSP_OK;
}

// The noise synth of voices [v0, v1) (get_controls scaling when scale_fn != 2, taps, noise, FIR: one
// kernel, noise.cuh), written as `n_slices` noise signals starting at slice index `slice0` of
// noise_part [*, B, N].  vp: magnitudes (and optional injected noise) of ALL voices.
static int run_noise_voices(b200ddsp_handle* h, int scale_fn, const NoiseVoicePtrs& vp, int v0, int v1,
                            int slice0, int n_slices, float* noise_part, int B, int F, int M,
                            unsigned long long seed, unsigned long long stream_id, cudaStream_t st,
                            const SpanInfo* span = nullptr, long long in_first_frame = 0) {
  const int U = h->U;
  const int F_out = span ? span->F_out : F;
  if (M != h->cfg.n_noise_bands || !h->d_cmat_t)
    return fail(h, B200DDSP_BAD_SHAPE, "M=%d but the handle was created for n_noise_bands=%d", M,
                h->cfg.n_noise_bands);
  if (U % 8 != 0)
    return fail(h, B200DDSP_UNSUPPORTED_CONFIG, "noise synth needs U %% 8 == 0 (U=%d)", U);
  if (U > 8 * 4 * 16)
    return fail(h, B200DDSP_UNSUPPORTED_CONFIG, "noise synth supports U <= 512 (U=%d)", U);
  StageTimer tm(h, B200DDSP_STAGE_NOISE, st);
  NoiseArgs a{};
  a.cmat_t = h->d_cmat_t;
  a.cmat_pitch = cmat_pitch_for(M);
  a.out = noise_part;
  a.v_begin = v0;
  a.v_end = v1;
  a.slice0 = slice0;
  a.B = B; a.F = F; a.M = M; a.U = U; a.N = F_out * U;
  a.koff = span ? span->koff : 0;
  a.sample0 = (unsigned long long)in_first_frame * (unsigned long long)U;
  a.halo_before = (M + U - 1) / U;
  a.halo_after = (U + M - 5) / U;
  a.scale_fn = scale_fn;
  a.bias = h->cfg.noise_initial_bias;
  a.seed = seed;
  a.stream_id = stream_id;
  // bulk asynchronous copies need 16-byte aligned rows: M % 4 == 0 and aligned tensors (the ABI asks for
  // 16-byte alignment; a misaligned view falls back to ordinary loads instead of faulting)
  a.bulk = (M % 4 == 0) && env_int("B200DDSP_NOISE_BULK", 1);
  for (int v = v0; v < v1; ++v) a.bulk = a.bulk && aligned16(vp.mags[v]);
  const NoiseSmemLayout L(M, U, a.halo_before, a.halo_after);
  const size_t smem = (size_t)L.total_floats * sizeof(float);
  if (smem > 227 * 1024)
    return fail(h, B200DDSP_UNSUPPORTED_CONFIG, "noise tile needs %zu bytes of shared memory", smem);
  // warps split the U/8 blocks of a frame: one block per thread up to 12 warps, else 2..4 per thread
  // on up to 16 warps
  const int n_blocks = U / 8;
  const int nb = n_blocks <= 12 ? 1 : (n_blocks <= 32 ? 2 : (n_blocks + 15) / 16);
  const int warps = (n_blocks + nb - 1) / nb;
  dim3 grid((F_out + kNoiseFrames - 1) / kNoiseFrames, B, n_slices);
  auto launch = [&](auto kernel) -> cudaError_t {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kernel<<<grid, warps * 32, smem, st>>>(a, vp);
    return cudaSuccess;
  };
  cudaError_t e = cudaSuccess;
  switch (nb) {
    case 1: e = launch(noise_synth_kernel<1>); break;
    case 2: e = launch(noise_synth_kernel<2>); break;
    case 3: e = launch(noise_synth_kernel<3>); break;
    default: e = launch(noise_synth_kernel<4>); break;
  }
  if (e != cudaSuccess)
    return fail(h, B200DDSP_CUDA_ERROR, "noise_synth_kernel attribute: %s", cudaGetErrorString(e));
  CHECK_LAUNCH_ON(h, "noise_synth_kernel", st);
  return B200DDSP_OK;
}

// dry = noise slices + additive partial signals (the MultiAdd nodes of the DAG)
static int run_mix(b200ddsp_handle* h, const float* noise_part, int n_noise, const AdditiveResult* mix,
                   float* out, int B, int N, int accumulate, cudaStream_t st) {
  MixArgs m{};
  m.noise = noise_part;
  m.n_noise = n_noise;
  m.partials = mix ? mix->partials : nullptr;
  m.live = mix ? mix->live : nullptr;
  m.n_partials = mix ? mix->n_partials : 0;
  m.sets = mix ? mix->sets : 1;
  m.out = out;
  m.B = B;
  m.N = N;
  m.chunk = chunk_for(h, N);
  m.n_chunks = n_chunks_for(h, N);
  m.accumulate = accumulate;
  if (N % 4 != 0 || m.chunk % 4 != 0)
    return fail(h, B200DDSP_UNSUPPORTED_CONFIG, "mixer needs N %% 4 == 0 (N=%d)", N);
  StageTimer tm(h, B200DDSP_STAGE_MIX, st);
  mix_kernel<<<dim3((N / 4 + 255) / 256, B), 256, 0, st>>>(m);
  CHECK_LAUNCH_ON(h, "mix_kernel", st);
  return B200DDSP_OK;
}

extern "C" size_t b200ddsp_noise_workspace_bytes(const b200ddsp_handle* h, int B, int F, int M) {
  if (!h || B < 1 || F < 1 || M < 2) return 0;
  (void)M;
  return align_up((size_t)B * F * h->U * 4);
}

extern "C" int b200ddsp_noise_signal(b200ddsp_handle* h, const float* magnitudes, const float* noise,
                                     uint64_t seed, uint64_t stream_id, float* out, int B, int F,
                                     int M, int accumulate, void* workspace, size_t workspace_bytes,
                                     void* stream) {
  if (int rc = check_common(h, B, F)) return rc;
  if (!magnitudes || !out) return fail(h, B200DDSP_BAD_ARGUMENT, "null tensor pointer");
  const size_t need = b200ddsp_noise_workspace_bytes(h, B, F, M);
  if (!workspace || workspace_bytes < need)
    return fail(h, B200DDSP_WORKSPACE_TOO_SMALL, "noise_signal needs %zu workspace bytes, got %zu", need,
                workspace_bytes);
  if (!aligned16(workspace)) return fail(h, B200DDSP_BAD_ALIGN, "workspace must be 16-byte aligned");
  NoiseVoicePtrs vp{};
  vp.mags[0] = magnitudes;
  vp.noise[0] = noise;
  reset_stage_flags(h);
  float* part = (float*)workspace;
  if (int rc = run_noise_voices(h, B200DDSP_SCALE_NONE, vp, 0, 1, 0, 1, part, B, F, M, seed, stream_id,
                                (cudaStream_t)stream))
    return rc;
  return run_mix(h, part, 1, nullptr, out, B, F * h->U, accumulate, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------------------------
// reverb
// ---------------------------------------------------------------------------------------------

template <class Loader, class Storer>
static void launch_fft_pass(int R, const Loader& ld, const Storer& st, const float2* tw, int n, int Ns,
                            int batches, cudaStream_t s) {
  if (R == 64) {
    dim3 grid64(n / 64 / kFft64J, batches);
    if constexpr (std::is_same<Loader, LoadComplex>::value) {
      static const int bulk = env_int("B200DDSP_FFT_BULK", 0);
      if (bulk == 2) {   // persistent, two-stage pipeline
        static const int per_sm = env_int("B200DDSP_FFT_PIPE_CTAS", 6);
        const int tiles = n / 64 / kFft64J * batches;
        const int grid = tiles < 148 * per_sm ? tiles : 148 * per_sm;
        if (Ns < 16) fft_pass64_pipelined_kernel<Storer, true><<<grid, kFft64Threads, 0, s>>>(ld.src, st, tw, n, Ns, batches);
        else fft_pass64_pipelined_kernel<Storer, false><<<grid, kFft64Threads, 0, s>>>(ld.src, st, tw, n, Ns, batches);
        return;
      }
      if (bulk) {
        if (Ns < 16) fft_pass64_kernel<Loader, Storer, true, true><<<grid64, kFft64Threads, 0, s>>>(ld, st, tw, n, Ns);
        else fft_pass64_kernel<Loader, Storer, false, true><<<grid64, kFft64Threads, 0, s>>>(ld, st, tw, n, Ns);
        return;
      }
    }
    if (Ns < 16) fft_pass64_kernel<Loader, Storer, true><<<grid64, kFft64Threads, 0, s>>>(ld, st, tw, n, Ns);
    else fft_pass64_kernel<Loader, Storer, false><<<grid64, kFft64Threads, 0, s>>>(ld, st, tw, n, Ns);
    return;
  }
  dim3 grid((n / R + kFftThreads - 1) / kFftThreads, batches);
  switch (R) {
    case 2: fft_pass_kernel<2><<<grid, kFftThreads, 0, s>>>(ld, st, tw, n, Ns); break;
    case 4: fft_pass_kernel<4><<<grid, kFftThreads, 0, s>>>(ld, st, tw, n, Ns); break;
    case 8: fft_pass_kernel<8><<<grid, kFftThreads, 0, s>>>(ld, st, tw, n, Ns); break;
    default: fft_pass_kernel<16><<<grid, kFftThreads, 0, s>>>(ld, st, tw, n, Ns); break;
  }
}

// Radix plan: as many radix-64 passes (shared-memory kernel) as possible, the remaining bits as
// small register-radix passes in front.  n < 1024 keeps the plain radix-16 plan.
static std::vector<int> fft_radices(int n) {
  int p = 0;
  while ((1 << p) < n) ++p;
  std::vector<int> r;
  if (n < 1024 || env_int("B200DDSP_FFT_RADIX16", 0)) {
    if (p % 4) r.push_back(1 << (p % 4));
    for (int i = 0; i < p / 4; ++i) r.push_back(16);
    return r;
  }
  const int rem = p % 6;
  if (rem == 5) { r.push_back(8); r.push_back(4); }
  else if (rem > 0) r.push_back(1 << rem);
  for (int i = 0; i < p / 6; ++i) r.push_back(64);
  return r;
}

// flags: B200DDSP_CONV_*; -1 = the handle's effects.Reverb configuration (mask ir[0], add_dry).
static int run_reverb(b200ddsp_handle* h, const float* audio, const float* ir, float* out, int B,
                      int N, int L, float2* tw, float2* buf_a, float2* buf_b, cudaStream_t st,
                      int flags = -1) {
  if (flags < 0) flags = B200DDSP_CONV_MASK_IR0 | (h->cfg.reverb_add_dry ? B200DDSP_CONV_ADD_DRY : 0);
  const bool full_output = (flags & B200DDSP_CONV_FULL) != 0;
  const int add_dry = (!full_output && (flags & B200DDSP_CONV_ADD_DRY)) ? 1 : 0;
  const int first_tap = (flags & B200DDSP_CONV_MASK_IR0) ? 1 : 0;
  StageTimer tm(h, B200DDSP_STAGE_REVERB, st);
  const int n = fft_size_for(N, L);
  const std::vector<int> radices = fft_radices(n);
  const int n_pass = (int)radices.size();
  fft_twiddle_kernel<<<(n + 255) / 256, 256, 0, st>>>(tw, n);
  CHECK_LAUNCH(h, "fft_twiddle_kernel");
  float4* scales = reinterpret_cast<float4*>(tw + n);   // [B], right behind the twiddle table
  unsigned int* maxima = reinterpret_cast<unsigned int*>(scales + B);   // [B][2]
  CUDA_TRY(h, cudaMemsetAsync(maxima, 0, (size_t)B * 8, st));
  reverb_maxima_kernel<<<dim3(32, B), 256, 0, st>>>(audio, ir, maxima, N, L, first_tap);
  CHECK_LAUNCH(h, "reverb_maxima_kernel");
  reverb_scales_kernel<<<(B + 63) / 64, 64, 0, st>>>(maxima, scales, B);
  CHECK_LAUNCH(h, "reverb_scales_kernel");
  // forward: (audio, ir) -> Z
  float2* src = nullptr;
  float2* dst = buf_a;
  int Ns = 1;
  for (int i = 0; i < n_pass; ++i) {
    const StoreComplex sto{dst, n};
    if (i == 0) {
      launch_fft_pass(radices[i], LoadAudioIr{audio, ir, scales, N, L, first_tap}, sto, tw, n, Ns, B, st);
    } else {
      launch_fft_pass(radices[i], LoadComplex{src, n}, sto, tw, n, Ns, B, st);
    }
    CHECK_LAUNCH(h, "fft_pass_kernel<forward>");
    Ns *= radices[i];
    src = dst;
    dst = (dst == buf_a) ? buf_b : buf_a;
  }
  // spectrum product, two clips per inverse transform
  const int pairs = (B + 1) / 2;
  {
    dim3 grid((n / 2 + 1 + 255) / 256, pairs);
    reverb_spectrum_kernel<<<grid, 256, 0, st>>>(src, dst, n, B);
    CHECK_LAUNCH(h, "reverb_spectrum_kernel");
    src = dst;
    dst = (dst == buf_a) ? buf_b : buf_a;
  }
  // inverse (as a forward transform of the conjugate), last pass writes the wet+dry signal
  Ns = 1;
  for (int i = 0; i < n_pass; ++i) {
    const LoadComplex ld{src, n};
    if (i == n_pass - 1) {
      const StoreWetPair sto{out, audio, scales, N, full_output ? N + L - 1 : N, B, 1.0f / (float)n,
                             add_dry};
      launch_fft_pass(radices[i], ld, sto, tw, n, Ns, pairs, st);
    } else {
      launch_fft_pass(radices[i], ld, StoreComplex{dst, n}, tw, n, Ns, pairs, st);
    }
    CHECK_LAUNCH(h, "fft_pass_kernel<inverse>");
    Ns *= radices[i];
    src = dst;
    dst = (dst == buf_a) ? buf_b : buf_a;
  }
  return B200DDSP_OK;
}

// The same convolution in two phases for the polyphonic forward (reverb.cuh, "split form").
// Phase 1 (needs the impulse responses only): twiddles, IR scales, forward transforms of IR pairs
// into rows [0, pairs) of the ping-pong buffers; returns where the spectra ended up.
static int run_reverb_ir_phase(b200ddsp_handle* h, const float* ir, int B, int N, int L, float2* tw,
                               float2* buf_a, float2* buf_b, const float2** ir_spectra,
                               cudaStream_t st) {
  const int n = fft_size_for(N, L);
  const std::vector<int> radices = fft_radices(n);
  const int pairs = (B + 1) / 2;
  fft_twiddle_kernel<<<(n + 255) / 256, 256, 0, st>>>(tw, n);
  CHECK_LAUNCH_ON(h, "fft_twiddle_kernel", st);
  float4* scales = reinterpret_cast<float4*>(tw + n);
  unsigned int* maxima = reinterpret_cast<unsigned int*>(scales + B);
  CUDA_TRY(h, cudaMemsetAsync(maxima, 0, (size_t)B * 8, st));
  reverb_maxima1_kernel<<<dim3(32, B), 256, 0, st>>>(ir, maxima, L, 1, 1);
  CHECK_LAUNCH_ON(h, "reverb_maxima1_kernel", st);
  reverb_scales_kernel<<<(B + 63) / 64, 64, 0, st>>>(maxima, scales, B);   // audio part still 1
  CHECK_LAUNCH_ON(h, "reverb_scales_kernel", st);
  float2* src = nullptr;
  float2* dst = buf_a;
  int Ns = 1;
  for (size_t i = 0; i < radices.size(); ++i) {
    const StoreComplex sto{dst, n};
    if (i == 0) launch_fft_pass(radices[i], LoadRealPair{ir, scales, L, 1, B, 1}, sto, tw, n, Ns, pairs, st);
    else launch_fft_pass(radices[i], LoadComplex{src, n}, sto, tw, n, Ns, pairs, st);
    CHECK_LAUNCH_ON(h, "fft_pass_kernel<ir>", st);
    Ns *= radices[i];
    src = dst;
    dst = (dst == buf_a) ? buf_b : buf_a;
  }
  *ir_spectra = src;
  return B200DDSP_OK;
}

// Phase 2 (the tail of the forward): dry pairs -> spectra, product with the IR spectra, inverse.
static int run_reverb_audio_phase(b200ddsp_handle* h, const float* audio, const float2* ir_spectra,
                                  float* out, int B, int N, int L, float2* tw, float2* buf_a,
                                  float2* buf_b, cudaStream_t st) {
  StageTimer tm(h, B200DDSP_STAGE_REVERB, st);
  const int n = fft_size_for(N, L);
  const std::vector<int> radices = fft_radices(n);
  const int n_pass = (int)radices.size();
  const int pairs = (B + 1) / 2;
  float4* scales = reinterpret_cast<float4*>(tw + n);
  unsigned int* maxima = reinterpret_cast<unsigned int*>(scales + B);
  reverb_maxima1_kernel<<<dim3(32, B), 256, 0, st>>>(audio, maxima, N, 0, 0);
  CHECK_LAUNCH_ON(h, "reverb_maxima1_kernel", st);
  reverb_scales_kernel<<<(B + 63) / 64, 64, 0, st>>>(maxima, scales, B);
  CHECK_LAUNCH_ON(h, "reverb_scales_kernel", st);
  float2* wa = buf_a + (size_t)pairs * n;   // rows [pairs, 2 pairs) of both buffers
  float2* wb = buf_b + (size_t)pairs * n;
  float2* src = nullptr;
  float2* dst = wa;
  int Ns = 1;
  for (int i = 0; i < n_pass; ++i) {
    const StoreComplex sto{dst, n};
    if (i == 0) launch_fft_pass(radices[i], LoadRealPair{audio, scales, N, 0, B, 0}, sto, tw, n, Ns, pairs, st);
    else launch_fft_pass(radices[i], LoadComplex{src, n}, sto, tw, n, Ns, pairs, st);
    CHECK_LAUNCH_ON(h, "fft_pass_kernel<forward>", st);
    Ns *= radices[i];
    src = dst;
    dst = (dst == wa) ? wb : wa;
  }
  {
    dim3 grid((n / 2 + 1 + 255) / 256, pairs);
    reverb_spectrum_split_kernel<<<grid, 256, 0, st>>>(src, ir_spectra, dst, n);
    CHECK_LAUNCH_ON(h, "reverb_spectrum_split_kernel", st);
    src = dst;
    dst = (dst == wa) ? wb : wa;
  }
  Ns = 1;
  for (int i = 0; i < n_pass; ++i) {
    const LoadComplex ld{src, n};
    if (i == n_pass - 1) {
      const StoreWetPair sto{out, audio, scales, N, N, B, 1.0f / (float)n, h->cfg.reverb_add_dry ? 1 : 0};
      launch_fft_pass(radices[i], ld, sto, tw, n, Ns, pairs, st);
    } else {
      launch_fft_pass(radices[i], ld, StoreComplex{dst, n}, tw, n, Ns, pairs, st);
    }
    CHECK_LAUNCH_ON(h, "fft_pass_kernel<inverse>", st);
    Ns *= radices[i];
    src = dst;
    dst = (dst == wa) ? wb : wa;
  }
  return B200DDSP_OK;
}

static int reverb_entry(b200ddsp_handle* h, const float* audio, const float* ir, float* out,
                        int B, int N, int L, void* workspace, size_t workspace_bytes,
                        void* stream, int flags) {
  if (!h) return B200DDSP_BAD_ARGUMENT;
  if (B < 1 || N < 1 || L < 1 || B > 65535)
    return fail(h, B200DDSP_BAD_SHAPE, "B=%d N=%d L=%d must be positive", B, N, L);
  if ((long long)N + L - 1 > (1ll << 28))
    return fail(h, B200DDSP_BAD_SHAPE, "N + L - 1 = %lld exceeds 2^28", (long long)N + L - 1);
  if (!audio || !ir || !out) return fail(h, B200DDSP_BAD_ARGUMENT, "null tensor pointer");
  if (out == audio) return fail(h, B200DDSP_BAD_ARGUMENT, "out may not alias audio");
  const int n = fft_size_for(N, L);
  const size_t tw_b = align_up((size_t)n * 8 + (size_t)B * 32), buf_b = align_up((size_t)B * n * 8);
  const size_t need = tw_b + 2 * buf_b;
  if (!workspace || workspace_bytes < need)
    return fail(h, B200DDSP_WORKSPACE_TOO_SMALL, "reverb needs %zu workspace bytes, got %zu", need,
                workspace_bytes);
  if (!aligned16(workspace)) return fail(h, B200DDSP_BAD_ALIGN, "workspace must be 16-byte aligned");
  char* w = (char*)workspace;
  reset_stage_flags(h);
  return run_reverb(h, audio, ir, out, B, N, L, (float2*)w, (float2*)(w + tw_b),
                    (float2*)(w + tw_b + buf_b), (cudaStream_t)stream, flags);
}

extern "C" int b200ddsp_reverb(b200ddsp_handle* h, const float* audio, const float* ir, float* out,
                               int B, int N, int L, void* workspace, size_t workspace_bytes,
                               void* stream) {
  return reverb_entry(h, audio, ir, out, B, N, L, workspace, workspace_bytes, stream, -1);
}

extern "C" int b200ddsp_reverb_full(b200ddsp_handle* h, const float* audio, const float* ir,
                                    float* out_full, int B, int N, int L, void* workspace,
                                    size_t workspace_bytes, void* stream) {
  return reverb_entry(h, audio, ir, out_full, B, N, L, workspace, workspace_bytes, stream,
                      B200DDSP_CONV_MASK_IR0 | B200DDSP_CONV_FULL);
}

extern "C" int b200ddsp_ir_decay_mask(b200ddsp_handle* h, const float* ir, float* out, int B, int L,
                                      float decay_exponent, int decay_start, void* stream) {
  if (!h) return B200DDSP_BAD_ARGUMENT;
  if (!ir || !out) return fail(h, B200DDSP_BAD_ARGUMENT, "null tensor pointer");
  if (B < 1 || L < 1 || B > 65535) return fail(h, B200DDSP_BAD_SHAPE, "B=%d L=%d", B, L);
  if (decay_start < 0 || decay_start >= L)
    return fail(h, B200DDSP_BAD_SHAPE,
                "decay_start=%d must lie inside the impulse response (L=%d): the reference's "
                "tf.linspace(0, 1, L - decay_start) is empty otherwise", decay_start, L);
  ir_decay_mask_kernel<<<dim3((L + 255) / 256, B), 256, 0, (cudaStream_t)stream>>>(ir, out, L, decay_start,
                                                                                 decay_exponent);
  CHECK_LAUNCH(h, "ir_decay_mask_kernel");
  return B200DDSP_OK;
}

extern "C" int b200ddsp_fft_convolve(b200ddsp_handle* h, const float* audio, const float* ir, float* out,
                                     int B, int N, int L, int flags, void* workspace,
                                     size_t workspace_bytes, void* stream) {
  if (flags < 0 || flags > 7) return fail(h, B200DDSP_BAD_ARGUMENT, "bad convolution flags %d", flags);
  return reverb_entry(h, audio, ir, out, B, N, L, workspace, workspace_bytes, stream, flags);
}

// ---------------------------------------------------------------------------------------------
// reverb of a span of a timeline (timeline.cuh): per-segment 'valid' convolution with the timeline's
// impulse response, then ONE kernel for the overlap-add and both ends of the tail hand-off
// ---------------------------------------------------------------------------------------------
struct TimelineReverbPlan {
  const float* ir;            // [B, L] device
  int B, n_seg, N, L, nfft, add_dry;
  float2 *tw, *buf_a, *buf_b; // twiddles; ping-pong buffers of B + ceil(rows / 2) transforms each
  float4 *scales, *ir_scales; // [rows], [B]
  unsigned int* maxima;       // [rows + B][2]
  float* wet_full;            // [rows, N + L - 1]
  mutable const float2* ir_spectra;
  Link tail;
};

struct TimelineReverbLayout { size_t tw, buf_a, buf_b, wet_full, total; int nfft; };

static TimelineReverbLayout carve_timeline_reverb(size_t at, int B, int n_seg, int N, int L) {
  TimelineReverbLayout t{};
  const size_t rows = (size_t)B * n_seg, pairs = (rows + 1) / 2;
  t.nfft = fft_size_for(N, L);
  size_t o = at;
  auto take = [&](size_t bytes) { size_t p = o; o += align_up(bytes); return p; };
  t.tw = take((size_t)t.nfft * 8 + (rows + B) * 16 + (rows + B) * 8);
  t.buf_a = take((B + pairs) * (size_t)t.nfft * 8);
  t.buf_b = take((B + pairs) * (size_t)t.nfft * 8);
  t.wet_full = take(rows * (size_t)(N + L - 1) * 4);
  t.total = o;
  return t;
}

static TimelineReverbPlan timeline_plan(const b200ddsp_handle* h, char* base, const TimelineReverbLayout& lay,
                                        const float* ir, int B, int n_seg, int N, int L, const Link& tail) {
  TimelineReverbPlan p{};
  const size_t rows = (size_t)B * n_seg;
  p.ir = ir; p.B = B; p.n_seg = n_seg; p.N = N; p.L = L; p.nfft = lay.nfft;
  p.add_dry = h->cfg.reverb_add_dry ? 1 : 0;
  p.tw = (float2*)(base + lay.tw);
  p.scales = reinterpret_cast<float4*>(p.tw + lay.nfft);
  p.ir_scales = p.scales + rows;
  p.maxima = reinterpret_cast<unsigned int*>(p.ir_scales + B);
  p.buf_a = (float2*)(base + lay.buf_a);
  p.buf_b = (float2*)(base + lay.buf_b);
  p.wet_full = (float*)(base + lay.wet_full);
  p.tail = tail;
  return p;
}

static int timeline_reverb_ir_phase(b200ddsp_handle* h, const TimelineReverbPlan& p, cudaStream_t st) {
  const int n = p.nfft, rows = p.B * p.n_seg;
  const std::vector<int> radices = fft_radices(n);
  fft_twiddle_kernel<<<(n + 255) / 256, 256, 0, st>>>(p.tw, n);
  CHECK_LAUNCH_ON(h, "fft_twiddle_kernel", st);
  CUDA_TRY(h, cudaMemsetAsync(p.maxima, 0, (size_t)(rows + p.B) * 8, st));
  unsigned int* max_ir = p.maxima + 2 * rows;
  reverb_maxima1_kernel<<<dim3(32, p.B), 256, 0, st>>>(p.ir, max_ir, p.L, 1, 1);
  CHECK_LAUNCH_ON(h, "reverb_maxima1_kernel", st);
  timeline_scales_kernel<<<(rows + 63) / 64, 64, 0, st>>>(nullptr, max_ir, p.scales, p.ir_scales, rows, p.n_seg);
  CHECK_LAUNCH_ON(h, "timeline_scales_kernel", st);
  float2* src = nullptr;
  float2* dst = p.buf_a;
  int Ns = 1;
  for (size_t i = 0; i < radices.size(); ++i) {
    const StoreComplex sto{dst, n};
    if (i == 0) launch_fft_pass(radices[i], LoadRealSingle{p.ir, p.ir_scales, p.L, 1}, sto, p.tw, n, Ns, p.B, st);
    else launch_fft_pass(radices[i], LoadComplex{src, n}, sto, p.tw, n, Ns, p.B, st);
    CHECK_LAUNCH_ON(h, "fft_pass_kernel<ir>", st);
    Ns *= radices[i];
    src = dst;
    dst = (dst == p.buf_a) ? p.buf_b : p.buf_a;
  }
  p.ir_spectra = src;
  return B200DDSP_OK;
}

static int timeline_reverb_audio_phase(b200ddsp_handle* h, const TimelineReverbPlan& p, const float* dry,
                                       float* out, cudaStream_t st) {
  StageTimer tm(h, B200DDSP_STAGE_REVERB, st);
  const int n = p.nfft, rows = p.B * p.n_seg, pairs = (rows + 1) / 2;
  const std::vector<int> radices = fft_radices(n);
  const int n_pass = (int)radices.size();
  reverb_maxima1_kernel<<<dim3(32, rows), 256, 0, st>>>(dry, p.maxima, p.N, 0, 0);
  CHECK_LAUNCH_ON(h, "reverb_maxima1_kernel", st);
  timeline_scales_kernel<<<(rows + 63) / 64, 64, 0, st>>>(p.maxima, p.maxima + 2 * rows, p.scales, nullptr, rows,
                                                          p.n_seg);
  CHECK_LAUNCH_ON(h, "timeline_scales_kernel", st);
  float2* wa = p.buf_a + (size_t)p.B * n;
  float2* wb = p.buf_b + (size_t)p.B * n;
  float2* src = nullptr;
  float2* dst = wa;
  int Ns = 1;
  for (int i = 0; i < n_pass; ++i) {
    const StoreComplex sto{dst, n};
    if (i == 0) launch_fft_pass(radices[i], LoadRealPair{dry, p.scales, p.N, 0, rows, 0}, sto, p.tw, n, Ns, pairs, st);
    else launch_fft_pass(radices[i], LoadComplex{src, n}, sto, p.tw, n, Ns, pairs, st);
    CHECK_LAUNCH_ON(h, "fft_pass_kernel<forward>", st);
    Ns *= radices[i];
    src = dst;
    dst = (dst == wa) ? wb : wa;
  }
  {
    dim3 grid((n / 2 + 1 + 255) / 256, pairs);
    timeline_spectrum_kernel<<<grid, 256, 0, st>>>(src, p.ir_spectra, dst, n, rows, p.n_seg);
    CHECK_LAUNCH_ON(h, "timeline_spectrum_kernel", st);
    src = dst;
    dst = (dst == wa) ? wb : wa;
  }
  const int total = p.N + p.L - 1;
  Ns = 1;
  for (int i = 0; i < n_pass; ++i) {
    const LoadComplex ld{src, n};
    if (i == n_pass - 1) {
      const StoreWetPair sto{p.wet_full, dry, p.scales, p.N, total, rows, 1.0f / (float)n, 0};
      launch_fft_pass(radices[i], ld, sto, p.tw, n, Ns, pairs, st);
    } else {
      launch_fft_pass(radices[i], ld, StoreComplex{dst, n}, p.tw, n, Ns, pairs, st);
    }
    CHECK_LAUNCH_ON(h, "fft_pass_kernel<inverse>", st);
    Ns *= radices[i];
    src = dst;
    dst = (dst == wa) ? wb : wa;
  }
  TimelineTailArgs ta{};
  ta.wet_full = p.wet_full;
  ta.dry = p.add_dry ? dry : nullptr;
  ta.out = out;
  ta.B = p.B; ta.n_seg = p.n_seg; ta.N = p.N; ta.total = total;
  const long long span = (long long)p.n_seg * p.N;
  const int tail = p.L - 1;
  ta.tail_ctas = p.tail.carry ? (tail + 255) / 256 : 0;
  ta.body_ctas = (int)((span - tail + 255) / 256);
  ta.head_ctas = (tail + 255) / 256;
  ta.link = p.tail;
  timeline_tail_kernel<<<dim3(ta.tail_ctas + ta.body_ctas + ta.head_ctas, p.B), 256, 0, st>>>(ta);
  CHECK_LAUNCH_ON(h, "timeline_tail_kernel", st);
  return B200DDSP_OK;
}

static int check_timeline_reverb(b200ddsp_handle* h, int B, int n_seg, int N, int L) {
  if (B < 1 || n_seg < 1 || N < 1 || L < 1 || B > 65535 || (long long)B * n_seg > 65535)
    return fail(h, B200DDSP_BAD_SHAPE, "B=%d n_seg=%d N=%d L=%d", B, n_seg, N, L);
  if ((long long)(L - 1) > (long long)n_seg * N)
    return fail(h, B200DDSP_BAD_SHAPE,
                "the reverb tail (%d samples) is longer than the span (%d segments x %d)", L - 1, n_seg, N);
  if ((long long)N + L - 1 > (1ll << 28))
    return fail(h, B200DDSP_BAD_SHAPE, "N + L - 1 = %lld exceeds 2^28", (long long)N + L - 1);
  return B200DDSP_OK;
}

extern "C" size_t b200ddsp_timeline_reverb_workspace_bytes(const b200ddsp_handle* h, int B, int n_seg, int N,
                                                           int L) {
  if (!h || B < 1 || n_seg < 1 || N < 1 || L < 1) return 0;
  return carve_timeline_reverb(0, B, n_seg, N, L).total;
}

extern "C" int b200ddsp_timeline_reverb(b200ddsp_handle* h, const float* dry, const float* reverb_ir,
                                        float* out, int B, int n_seg, int N, int L, const b200ddsp_link* tail,
                                        void* workspace, size_t workspace_bytes, void* stream) {
  if (!h) return B200DDSP_BAD_ARGUMENT;
  if (!dry || !reverb_ir || !out) return fail(h, B200DDSP_BAD_ARGUMENT, "null tensor pointer");
  if (out == dry) return fail(h, B200DDSP_BAD_ARGUMENT, "out may not alias dry");
  if (int rc = check_timeline_reverb(h, B, n_seg, N, L)) return rc;
  const TimelineReverbLayout lay = carve_timeline_reverb(0, B, n_seg, N, L);
  if (!workspace || workspace_bytes < lay.total)
    return fail(h, B200DDSP_WORKSPACE_TOO_SMALL, "timeline_reverb needs %zu workspace bytes, got %zu",
                lay.total, workspace_bytes);
  if (!aligned16(workspace)) return fail(h, B200DDSP_BAD_ALIGN, "workspace must be 16-byte aligned");
  reset_stage_flags(h);
  const TimelineReverbPlan p = timeline_plan(h, (char*)workspace, lay, reverb_ir, B, n_seg, N, L,
                                             tail ? to_link(*tail) : Link{});
  cudaStream_t st = (cudaStream_t)stream;
  if (int rc = timeline_reverb_ir_phase(h, p, st)) return rc;
  return timeline_reverb_audio_phase(h, p, dry, out, st);
}

// ---------------------------------------------------------------------------------------------
// feedback-delay-network impulse response
// ---------------------------------------------------------------------------------------------

extern "C" size_t b200ddsp_fdn_workspace_bytes(const b200ddsp_handle* h, float sampling_rate, int B) {
  if (!h || B < 1 || !(sampling_rate >= 1.0f)) return 0;
  const int n = (int)(2.0f * sampling_rate);
  return align_up((size_t)B * (n / 2 + 1) * sizeof(float2));
}

extern "C" int b200ddsp_fdn_ir(b200ddsp_handle* h, const float* input_gain, const float* output_gain,
                               const float* gain_allpass, const float* delays_allpass,
                               const float* time_rev_0_sec, const float* alpha_tone,
                               const float* early_ir, int E, const float* delay_values,
                               int delay_lines, float sampling_rate, float* ir_out, int B,
                               void* workspace, size_t workspace_bytes, void* stream) {
  if (!h) return B200DDSP_BAD_ARGUMENT;
  if (delay_lines != 6 && delay_lines != 8)
    return fail(h, B200DDSP_UNSUPPORTED_CONFIG, "delay_lines=%d: the network is built for 8 lines "
                "(fdn_reverb.py:30) or 6 (configs/ENSTDkCl-*.gin)", delay_lines);
  if (delay_lines != 8 && !delay_values)
    return fail(h, B200DDSP_BAD_ARGUMENT, "delay_values are required when delay_lines != 8");
  if (!input_gain || !output_gain || !gain_allpass || !delays_allpass || !time_rev_0_sec || !alpha_tone ||
      !ir_out || (E > 0 && !early_ir))
    return fail(h, B200DDSP_BAD_ARGUMENT, "null tensor pointer");
  if (B < 1 || B > 65535 || E < 0) return fail(h, B200DDSP_BAD_SHAPE, "B=%d E=%d", B, E);
  if (!(sampling_rate >= 1.0f) || sampling_rate > 4.0e6f)
    return fail(h, B200DDSP_BAD_SHAPE, "sampling_rate %g out of range", (double)sampling_rate);
  const int n = (int)(2.0f * sampling_rate);            // freq_points = int(2 * sampling_rate), :83
  if (n < 2 || (n & 1)) return fail(h, B200DDSP_UNSUPPORTED_CONFIG, "freq_points=%d must be even", n);
  const size_t need = b200ddsp_fdn_workspace_bytes(h, sampling_rate, B);
  if (!workspace || workspace_bytes < need)
    return fail(h, B200DDSP_WORKSPACE_TOO_SMALL, "fdn_ir needs %zu workspace bytes, got %zu", need,
                workspace_bytes);
  if (!aligned16(workspace)) return fail(h, B200DDSP_BAD_ALIGN, "workspace must be 16-byte aligned");
  static const float kDefaultDelays[kFdnLines] = {233, 311, 421, 461, 587, 613, 789, 891};
  FdnArgs a{};
  a.input_gain = input_gain; a.output_gain = output_gain;
  a.gain_allpass = gain_allpass; a.delays_allpass = delays_allpass;
  a.time_rev_0_sec = time_rev_0_sec; a.alpha_tone = alpha_tone;
  a.early_ir = early_ir;
  a.H = (float2*)workspace;
  a.ir = ir_out;
  for (int d = 0; d < delay_lines; ++d) a.delay_values[d] = delay_values ? delay_values[d] : kDefaultDelays[d];
  a.sampling_rate = sampling_rate;
  a.n = n; a.E = E < n ? E : n; a.B = B;
  cudaStream_t st = (cudaStream_t)stream;
  if (delay_lines == 8) fdn_transfer_kernel<8><<<dim3((n / 2 + 1 + 127) / 128, B), 128, 0, st>>>(a);
  else fdn_transfer_kernel<6><<<dim3((n / 2 + 1 + 127) / 128, B), 128, 0, st>>>(a);
  CHECK_LAUNCH(h, "fdn_transfer_kernel");
  fdn_irfft_kernel<<<dim3((n + 255) / 256, B), 256, 0, st>>>(a);
  CHECK_LAUNCH(h, "fdn_irfft_kernel");
  return B200DDSP_OK;
}

// ---------------------------------------------------------------------------------------------
// the whole DAG
// ---------------------------------------------------------------------------------------------

// Optional cross-stream ordering for the host-input entry point: the main stream waits for
// `small_ready` before the prep kernel, `group_ready[g]` before it touches the harmonic
// distribution of voice group g, `mags_ready[h]` before the noise stage of voice half h,
// `ir_ready` before the reverb.
// host-input path: the magnitudes arrive in this many voice parts (one noise slice each)
static int noise_parts(int P) { return P < kNoiseSlices ? 1 : kNoiseSlices; }

struct ForwardSync {
  float* dry_host;                              // where the dry signal goes as soon as it is mixed (or nullptr)
  cudaEvent_t small_ready;                      // amplitudes, inharm_coef, f0_hz of all voices
  cudaEvent_t group_ready[kMaxVoiceGroups];     // harmonic_distribution of voice group g
  cudaEvent_t mags_ready[4];                    // magnitudes of voice part p (noise_parts(P) parts)
  cudaEvent_t ir_ready;
};

static int forward_core(b200ddsp_handle* h, const b200ddsp_voice* voices, int P, const float* reverb_ir,
                        float* dry_out, float* wet_out, int B, int F, int H, int S, int M, int L,
                        uint64_t seed, char* base, const WorkspaceLayout& w, const ForwardSync* sync,
                        cudaStream_t st, const SpanInfo* span = nullptr, long long in_first_frame = 0,
                        const TimelineReverbPlan* tl = nullptr) {
  const int U = h->U, N = (span ? span->F_out : F) * U;
  float* amp = (float*)(base + w.amp);
  float* hd = (float*)(base + w.hd);
  float* shifts = (float*)(base + w.shifts);
  float* f0 = (float*)(base + w.f0);

  NoiseVoicePtrs vp{};
  AdditiveControlsPtrs cp{};
  for (int v = 0; v < P; ++v) {
    const b200ddsp_voice& vc = voices[v];
    if (!vc.amplitudes || !vc.harmonic_distribution || !vc.inharm_coef || !vc.f0_hz || !vc.magnitudes)
      return fail(h, B200DDSP_BAD_ARGUMENT, "voice %d has a null control tensor", v);
    vp.mags[v] = vc.magnitudes;
    vp.noise[v] = vc.noise;
    cp.amp_in[v] = vc.amplitudes;
    cp.hd_in[v] = vc.harmonic_distribution;
    cp.inharm_in[v] = vc.inharm_coef;
    cp.f0_in[v] = vc.f0_hz;
  }
  reset_stage_flags(h);
  h->trace_n = 0;
  trace_launch(h, "entry", st);
  AdditiveRun run;
  if (int rc = additive_begin(h, &run, amp, hd, shifts, f0, base, w.add, P, B, F, H, S, w.groups, nullptr,
                              nullptr, span))
    return rc;
  AdditiveControlsArgs ca = controls_args(h, B * F, H, S);
  ca.amp_out = amp; ca.hd_out = hd; ca.shifts_out = shifts; ca.f0_out = f0;
  ca.na_frame = run.fast ? run.na_frame : nullptr;
  ca.cut_index = run.cut_index;

  // 3. noise of every voice -> noise slices; FilteredNoise.get_controls is fused into the taps
  //    GEMM's operand load.  It depends on the magnitudes only, so it is enqueued on its own stream
  //    and shares the SMs with the phase pass and the oscillator bank (neither fills every issue
  //    slot); the mixer joins the two.  The voices are processed in halves (each half = its share of
  //    the noise slices), so that on the host-input path the FIR of the first half runs while the
  //    second half is still copying.
  float* noise_part = (float*)(base + w.noise_part);
  const int n_slices = P < kNoiseSlices ? P : kNoiseSlices;
  auto enqueue_noise = [&](cudaStream_t ns) -> int {
    const int halves = sync ? noise_parts(P) : 1;
    for (int hf = 0; hf < halves; ++hf) {
      const int v0 = P * hf / halves, v1 = P * (hf + 1) / halves;
      const int s0 = n_slices * hf / halves, s1 = n_slices * (hf + 1) / halves;
      if (sync) CUDA_TRY(h, cudaStreamWaitEvent(ns, sync->mags_ready[hf], 0));
      if (int rc = run_noise_voices(h, h->cfg.noise_scale_fn, vp, v0, v1, s0, s1 - s0, noise_part, B,
                                    F, M, seed, 0, ns, span, in_first_frame))
        return rc;
    }
    return B200DDSP_OK;
  };
  // 0 after the oscillators, 1 from the start, 2 after the work lists, 3 after the scan, 4 after the phase
  // pass in front of the scan (measured on whole clips: 1 and 2 equal, 0 is 2 % slower).  A span that takes
  // its phase state from a PEER waits for it in front of the scan -- the later in the chain, the longer --
  // so there the noise synth is started at that point (4) and runs under the wait.
  static const int noise_env = env_int("B200DDSP_NOISE_STREAM", -1);
  const bool waits_for_peer = span && span->phase.seed && span->phase.seed_ready;
  const int noise_mode = noise_env >= 0 ? noise_env : (waits_for_peer && run.fast ? 4 : 1);
  const bool noise_beside = noise_mode != 0;
  auto fork_noise = [&](bool record) -> int {
    if (record) CUDA_TRY(h, cudaEventRecord(h->ev_noise_fork, st));
    CUDA_TRY(h, cudaStreamWaitEvent(h->noise_stream, h->ev_noise_fork, 0));
    if (int rc = enqueue_noise(h->noise_stream)) return rc;
    CUDA_TRY(h, cudaEventRecord(h->ev_noise_join, h->noise_stream));
    return B200DDSP_OK;
  };
  if (noise_mode == 1)
    if (int rc = fork_noise(true)) return rc;

  // 1. everything that does not need harmonic_distribution: amplitudes, inharmonic shifts,
  //    liveness, then the phase pass of ALL voices (chunk end phases -> chunk offsets)
  if (sync) CUDA_TRY(h, cudaStreamWaitEvent(st, sync->small_ready, 0));
  {
    StageTimer tm(h, B200DDSP_STAGE_CONTROLS, st);
    launch_additive_prep(ca, cp, P, st);
    CHECK_LAUNCH_ON(h, "additive_prep_kernel", st);
  }

  // harmonic_distribution (scale, Nyquist cut, normalise: an HBM stream over 147 MB) needs nothing
  // from the phase pass, which is bound by the FP32 pipe: it runs beside it on its own stream, one
  // launch per voice group as the group's copy arrives.  It starts after the prep kernel, which leaves it the
  // index of the first partial above Nyquist per frame (96 square roots per frame it need not repeat).
  CUDA_TRY(h, cudaEventRecord(h->ev_hd_fork, st));
  CUDA_TRY(h, cudaStreamWaitEvent(h->hd_stream, h->ev_hd_fork, 0));
  for (int g = 0; g < w.groups.n_groups; ++g) {
    const int v0 = w.groups.first_voice[g], Pg = w.groups.first_voice[g + 1] - v0;
    if (sync) CUDA_TRY(h, cudaStreamWaitEvent(h->hd_stream, sync->group_ready[g], 0));
    AdditiveControlsPtrs gp{};
    for (int i = 0; i < Pg; ++i) {
      gp.hd_in[i] = cp.hd_in[v0 + i];
      gp.inharm_in[i] = cp.inharm_in[v0 + i];
      gp.f0_in[i] = cp.f0_in[v0 + i];
    }
    AdditiveControlsArgs ga = ca;
    ga.hd_out = hd + (size_t)v0 * B * F * H;
    if (ga.cut_index) ga.cut_index += (size_t)v0 * B * F;
    launch_additive_hd(ga, gp, Pg, h->hd_stream);
    CHECK_LAUNCH_ON(h, "additive_hd_kernel", h->hd_stream);
    CUDA_TRY(h, cudaEventRecord(h->ev_hd_done[g], h->hd_stream));
  }
  // the impulse responses are known now, the dry signal only at the very end: their spectra are
  // prepared on the same side stream, which leaves the tail with 8 + 8 transforms instead of 16 + 8
  static const int reverb_split = env_int("B200DDSP_REVERB_SPLIT", 1);
  const float2* ir_spectra = nullptr;
  if (tl) {   // timeline reverb: the impulse responses' spectra, early and off the critical path
    if (sync) CUDA_TRY(h, cudaStreamWaitEvent(h->hd_stream, sync->ir_ready, 0));
    if (int rc = timeline_reverb_ir_phase(h, *tl, h->hd_stream)) return rc;
    CUDA_TRY(h, cudaEventRecord(h->ev_ir_spectra, h->hd_stream));
  } else if (reverb_ir && reverb_split) {
    if (sync) CUDA_TRY(h, cudaStreamWaitEvent(h->hd_stream, sync->ir_ready, 0));
    if (int rc = run_reverb_ir_phase(h, reverb_ir, B, N, L, (float2*)(base + w.tw), (float2*)(base + w.buf_a),
                                     (float2*)(base + w.buf_b), &ir_spectra, h->hd_stream))
      return rc;
    CUDA_TRY(h, cudaEventRecord(h->ev_ir_spectra, h->hd_stream));
  }

  // (the phase pass of ALL voices follows: chunk end phases -> chunk offsets)
  // the noise joins once the small latency-bound kernels of the phase pass are behind us (the
  // event is recorded after the work lists are built): it then shares the SMs with the long
  // phase and oscillator kernels only
  if (int rc = additive_phase_pass(h, run, run.fast, st, noise_mode == 2 ? h->ev_noise_fork : nullptr,
                                   noise_mode == 4 ? h->ev_noise_fork : nullptr))
    return rc;
  if (noise_mode >= 2)
    if (int rc = fork_noise(noise_mode == 3)) return rc;

  // 2. per voice group: the oscillator bank -> partial signals (its harmonic distribution was
  //    prepared on the side stream, see below)
  for (int g = 0; g < w.groups.n_groups; ++g) {
    CUDA_TRY(h, cudaStreamWaitEvent(st, h->ev_hd_done[g], 0));
    if (int rc = additive_synth_group(h, run, g, st)) return rc;
  }
  const AdditiveResult mix = additive_result(run);

  if (!noise_beside) {
    if (int rc = enqueue_noise(st)) return rc;
  } else {
    CUDA_TRY(h, cudaStreamWaitEvent(st, h->ev_noise_join, 0));
  }
  if (int rc = run_mix(h, noise_part, n_slices, &mix, dry_out, B, N, 0, st)) return rc;
  if (sync && sync->dry_host) {   // host-input path: the dry signal travels home under the reverb
    CUDA_TRY(h, cudaEventRecord(h->ev_dry_ready, st));
    CUDA_TRY(h, cudaStreamWaitEvent(h->d2h_stream, h->ev_dry_ready, 0));
    CUDA_TRY(h, cudaMemcpyAsync(sync->dry_host, dry_out, (size_t)B * N * 4, cudaMemcpyDeviceToHost, h->d2h_stream));
    CUDA_TRY(h, cudaEventRecord(h->ev_dry_home, h->d2h_stream));
  }
  if (tl) {
    CUDA_TRY(h, cudaStreamWaitEvent(st, h->ev_ir_spectra, 0));
    return timeline_reverb_audio_phase(h, *tl, dry_out, wet_out, st);
  }
  // 4. reverb -> wet
  if (reverb_ir && ir_spectra) {
    CUDA_TRY(h, cudaStreamWaitEvent(st, h->ev_ir_spectra, 0));
    return run_reverb_audio_phase(h, dry_out, ir_spectra, wet_out, B, N, L, (float2*)(base + w.tw),
                                  (float2*)(base + w.buf_a), (float2*)(base + w.buf_b), st);
  }
  if (reverb_ir) {
    if (sync) CUDA_TRY(h, cudaStreamWaitEvent(st, sync->ir_ready, 0));
    return run_reverb(h, dry_out, reverb_ir, wet_out, B, N, L, (float2*)(base + w.tw),
                      (float2*)(base + w.buf_a), (float2*)(base + w.buf_b), st);
  }
  return B200DDSP_OK;
}

static int check_forward_args(b200ddsp_handle* h, const b200ddsp_voice* voices, int P, int B, int F,
                              int H, int S, int M, int L, bool has_ir) {
  if (int rc = check_common(h, B, F)) return rc;
  if (!voices || P < 1 || P > B200DDSP_MAX_VOICES)
    return fail(h, B200DDSP_BAD_SHAPE, "P=%d outside [1, %d]", P, B200DDSP_MAX_VOICES);
  if (H < 1 || H > 256) return fail(h, B200DDSP_BAD_SHAPE, "H=%d outside [1, 256]", H);
  if (S < 1 || S > 32) return fail(h, B200DDSP_BAD_SHAPE, "S=%d outside [1, 32]", S);
  if (M < 3) return fail(h, B200DDSP_BAD_SHAPE, "M=%d must be at least 3", M);
  if (has_ir && L < 1) return fail(h, B200DDSP_BAD_SHAPE, "reverb_ir given but L < 1");
  return B200DDSP_OK;
}

extern "C" int b200ddsp_forward_polyphonic(b200ddsp_handle* h, const b200ddsp_voice* voices, int P,
                                           const float* reverb_ir, float* dry_out, float* wet_out,
                                           int B, int F, int H, int S, int M, int L, uint64_t seed,
                                           void* workspace, size_t workspace_bytes, void* stream) {
  if (int rc = check_forward_args(h, voices, P, B, F, H, S, M, L, reverb_ir != nullptr)) return rc;
  if (!dry_out) return fail(h, B200DDSP_BAD_ARGUMENT, "dry_out is null");
  if (reverb_ir && !wet_out) return fail(h, B200DDSP_BAD_ARGUMENT, "reverb_ir given but wet_out is null");
  if (reverb_ir && wet_out == dry_out)
    return fail(h, B200DDSP_BAD_ARGUMENT, "wet_out may not alias dry_out");
  const WorkspaceLayout w = carve(h, P, B, F, H, S, M, reverb_ir ? L : 0, env_int("B200DDSP_DEV_GROUPS", 1));
  if (!workspace || workspace_bytes < w.total)
    return fail(h, B200DDSP_WORKSPACE_TOO_SMALL, "forward_polyphonic needs %zu workspace bytes, got %zu",
                w.total, workspace_bytes);
  if (!aligned16(workspace)) return fail(h, B200DDSP_BAD_ALIGN, "workspace must be 16-byte aligned");
  return forward_core(h, voices, P, reverb_ir, dry_out, wet_out, B, F, H, S, M, L, seed,
                      (char*)workspace, w, nullptr, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------------------------
// staging of HOST inputs
// ---------------------------------------------------------------------------------------------
struct HostStaging {
  size_t amp, hd, inh, f0, mags, noise, ir, dry, wet, core, total;
};

// F = input frames, F_out = frames synthesised (== F for whole clips); `core_bytes` of scratch follow.
static HostStaging carve_host(int P, int B, int F, int F_out, int H, int S, int M, int L, int U, bool noise,
                              size_t core_bytes) {
  HostStaging s{};
  const size_t R = (size_t)P * B, N_in = (size_t)F * U, N = (size_t)F_out * U;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t at = o; o += align_up(bytes); return at; };
  s.amp = take(R * F * 4);
  s.hd = take(R * F * H * 4);
  s.inh = take(R * F * 4);
  s.f0 = take(R * F * S * 4);
  s.mags = take(R * F * (size_t)M * 4);
  if (noise) s.noise = take(R * N_in * 4);
  if (L > 0) s.ir = take((size_t)B * L * 4);
  s.dry = take((size_t)B * N * 4);
  if (L > 0) s.wet = take((size_t)B * N * 4);
  s.core = take(core_bytes);
  s.total = o;
  return s;
}

static int host_copy_groups(int P) {
  const int g = env_int("B200DDSP_HOST_GROUPS", 2);
  return P >= g ? g : P;
}

// ---------------------------------------------------------------------------------------------
// spans of a timeline (device inputs)
// ---------------------------------------------------------------------------------------------
static int make_span(b200ddsp_handle* h, const b200ddsp_span* sp, int F, SpanInfo* out) {
  if (!sp) return fail(h, B200DDSP_BAD_ARGUMENT, "span is null");
  const int U = h->U;
  const long long in0 = sp->in_first_frame, out0 = sp->out_first_frame, total = sp->total_frames;
  const long long Fo = sp->n_out_frames;
  if (in0 < 0 || out0 < in0 || Fo < 1 || total < 1 || out0 + Fo > total || in0 + F > total)
    return fail(h, B200DDSP_BAD_SHAPE,
                "span: input frames [%lld, %lld), output frames [%lld, %lld) of a timeline of %lld", in0,
                in0 + F, out0, out0 + Fo, total);
  if (total * U > (1ll << 24))
    return fail(h, B200DDSP_BAD_SHAPE,
                "timeline of %lld samples exceeds 2^24 (the float32 sample index of the reference's "
                "resize is no longer exact)", total * U);
  if (out0 > 0 && out0 == in0)
    return fail(h, B200DDSP_BAD_SHAPE, "span: one frame of halo is needed before output frame %lld", out0);
  if (out0 + Fo < total && in0 + F < out0 + Fo + 1)
    return fail(h, B200DDSP_BAD_SHAPE, "span: one frame of halo is needed after output frame %lld",
                out0 + Fo - 1);
  if ((out0 * U) % kAngularChunk != 0)
    return fail(h, B200DDSP_BAD_SHAPE,
                "span: the first output sample (%lld) must lie on a multiple of the %d-sample chunk of "
                "angular_cumsum", out0 * U, kAngularChunk);
  if (sp->phase.carry && ((out0 + Fo) * U) % kAngularChunk != 0 && out0 + Fo != total)
    return fail(h, B200DDSP_BAD_SHAPE, "span: a phase carry needs the span to end on a chunk boundary");
  if (out0 > 0 && !sp->phase.seed)
    return fail(h, B200DDSP_BAD_ARGUMENT, "span: output frame %lld > 0 needs a phase seed", out0);
  out->koff = (int)(out0 - in0);
  out->F_out = (int)Fo;
  out->tg0 = (int)(out0 * U);
  out->total_frames = (int)total;
  out->seeded = out0 > 0;
  out->carry = sp->phase.carry != nullptr;
  out->phase = to_link(sp->phase);
  if (out0 == 0) { out->phase.seed = nullptr; out->phase.seed_ready = nullptr; out->phase.seed_ack = nullptr; }
  return B200DDSP_OK;
}

extern "C" int b200ddsp_forward_span(b200ddsp_handle* h, const b200ddsp_voice* voices, int P, float* dry_out,
                                     int B, int F, int H, int S, int M, uint64_t seed,
                                     const b200ddsp_span* span, void* workspace, size_t workspace_bytes,
                                     void* stream) {
  if (int rc = check_forward_args(h, voices, P, B, F, H, S, M, 0, false)) return rc;
  if (!dry_out) return fail(h, B200DDSP_BAD_ARGUMENT, "dry_out is null");
  SpanInfo si{};
  if (int rc = make_span(h, span, F, &si)) return rc;
  const WorkspaceLayout w = carve(h, P, B, F, H, S, M, 0, 1);
  if (!workspace || workspace_bytes < w.total)
    return fail(h, B200DDSP_WORKSPACE_TOO_SMALL, "forward_span needs %zu workspace bytes, got %zu", w.total,
                workspace_bytes);
  if (!aligned16(workspace)) return fail(h, B200DDSP_BAD_ALIGN, "workspace must be 16-byte aligned");
  return forward_core(h, voices, P, nullptr, dry_out, nullptr, B, F, H, S, M, 0, seed, (char*)workspace, w,
                      nullptr, (cudaStream_t)stream, &si, span->in_first_frame);
}

static int check_timeline_args(b200ddsp_handle* h, const b200ddsp_span* span, int B, int L, int seg_frames) {
  if (!span) return fail(h, B200DDSP_BAD_ARGUMENT, "span is null");
  if (seg_frames < 1 || span->n_out_frames % seg_frames != 0)
    return fail(h, B200DDSP_BAD_SHAPE, "n_out_frames=%d is not a multiple of seg_frames=%d",
                span->n_out_frames, seg_frames);
  return check_timeline_reverb(h, B, span->n_out_frames / seg_frames, seg_frames * h->U, L);
}

struct TimelineLayout { WorkspaceLayout core; TimelineReverbLayout rev; size_t dry, total; };

static TimelineLayout carve_timeline(const b200ddsp_handle* h, int P, int B, int F, int H, int S, int M, int L,
                                     int n_out_frames, int seg_frames, int n_groups) {
  TimelineLayout t{};
  t.core = carve(h, P, B, F, H, S, M, 0, n_groups);
  t.rev = carve_timeline_reverb(t.core.total, B, n_out_frames / seg_frames, seg_frames * h->U, L);
  t.dry = t.rev.total;                                   // dry span when the caller does not want it
  t.total = t.dry + align_up((size_t)B * n_out_frames * h->U * 4);
  return t;
}

extern "C" int b200ddsp_forward_timeline(b200ddsp_handle* h, const b200ddsp_voice* voices, int P,
                                         const float* reverb_ir, float* dry_out, float* wet_out, int B,
                                         int F, int H, int S, int M, int L, int seg_frames, uint64_t seed,
                                         const b200ddsp_span* span, const b200ddsp_link* tail,
                                         void* workspace, size_t workspace_bytes, void* stream) {
  if (int rc = check_forward_args(h, voices, P, B, F, H, S, M, L, true)) return rc;
  if (!reverb_ir || !wet_out) return fail(h, B200DDSP_BAD_ARGUMENT, "reverb_ir / wet_out is null");
  if (wet_out == dry_out) return fail(h, B200DDSP_BAD_ARGUMENT, "wet_out may not alias dry_out");
  if (int rc = check_timeline_args(h, span, B, L, seg_frames)) return rc;
  SpanInfo si{};
  if (int rc = make_span(h, span, F, &si)) return rc;
  const TimelineLayout t = carve_timeline(h, P, B, F, H, S, M, L, si.F_out, seg_frames, 1);
  if (!workspace || workspace_bytes < t.total)
    return fail(h, B200DDSP_WORKSPACE_TOO_SMALL, "forward_timeline needs %zu workspace bytes, got %zu",
                t.total, workspace_bytes);
  if (!aligned16(workspace)) return fail(h, B200DDSP_BAD_ALIGN, "workspace must be 16-byte aligned");
  char* base = (char*)workspace;
  const TimelineReverbPlan tl = timeline_plan(h, base, t.rev, reverb_ir, B, si.F_out / seg_frames,
                                              seg_frames * h->U, L, tail ? to_link(*tail) : Link{});
  float* dry = dry_out ? dry_out : (float*)(base + t.dry);
  return forward_core(h, voices, P, nullptr, dry, wet_out, B, F, H, S, M, L, seed, base, t.core, nullptr,
                      (cudaStream_t)stream, &si, span->in_first_frame, &tl);
}

extern "C" size_t b200ddsp_timeline_workspace_bytes(const b200ddsp_handle* h, int P, int B, int F, int H, int S,
                                                    int M, int L, int n_out_frames, int seg_frames,
                                                    int host_inputs, int with_noise) {
  if (!h || P < 1 || B < 1 || F < 1 || H < 1 || S < 1 || M < 3 || L < 1 || seg_frames < 1 ||
      n_out_frames < seg_frames)
    return 0;
  const TimelineLayout t = carve_timeline(h, P, B, F, H, S, M, L, n_out_frames, seg_frames,
                                          host_inputs ? host_copy_groups(P) : 1);
  if (!host_inputs) return t.total;
  return carve_host(P, B, F, n_out_frames, H, S, M, L, h->U, with_noise != 0, t.total).total;
}

// ---------------------------------------------------------------------------------------------
// the whole DAG from HOST buffers
// ---------------------------------------------------------------------------------------------
extern "C" size_t b200ddsp_workspace_bytes_host(const b200ddsp_handle* h, int P, int B, int F, int H,
                                                int S, int M, int L, int with_noise) {
  if (!h || P < 1 || B < 1 || F < 1 || H < 1 || S < 1 || M < 3) return 0;
  const WorkspaceLayout w = carve(h, P, B, F, H, S, M, L, host_copy_groups(P));
  return carve_host(P, B, F, F, H, S, M, L, h->U, with_noise != 0, w.total).total;
}

// Shared by b200ddsp_forward_polyphonic_host (span == nullptr: whole clips, one-block reverb) and
// b200ddsp_forward_timeline_host (a span of a timeline, segment reverb + tail hand-off).
static int forward_host_impl(b200ddsp_handle* h, const b200ddsp_voice* voices_host, int P,
                             const float* reverb_ir_host, float* dry_out_host, float* wet_out_host, int B,
                             int F, int H, int S, int M, int L, uint64_t seed, const b200ddsp_span* span,
                             int seg_frames, const b200ddsp_link* tail, void* workspace,
                             size_t workspace_bytes, void* stream) {
  if (int rc = check_forward_args(h, voices_host, P, B, F, H, S, M, L, reverb_ir_host != nullptr))
    return rc;
  if (!dry_out_host && !wet_out_host)
    return fail(h, B200DDSP_BAD_ARGUMENT, "neither dry_out_host nor wet_out_host given");
  if (wet_out_host && !reverb_ir_host)
    return fail(h, B200DDSP_BAD_ARGUMENT, "wet_out_host given without reverb_ir_host");
  const int U = h->U;
  SpanInfo si{};
  if (span) {
    if (!reverb_ir_host || !wet_out_host)
      return fail(h, B200DDSP_BAD_ARGUMENT, "forward_timeline_host needs reverb_ir_host and wet_out_host");
    if (int rc = check_timeline_args(h, span, B, L, seg_frames)) return rc;
    if (int rc = make_span(h, span, F, &si)) return rc;
  }
  const int F_out = span ? si.F_out : F, N = F_out * U, N_in = F * U;
  bool any_noise = false;
  for (int v = 0; v < P; ++v) any_noise |= voices_host[v].noise != nullptr;
  const int Lr = reverb_ir_host ? L : 0;
  WorkspaceLayout w{};
  TimelineLayout tlay{};
  if (span) {
    tlay = carve_timeline(h, P, B, F, H, S, M, L, F_out, seg_frames, host_copy_groups(P));
    w = tlay.core;
  } else {
    w = carve(h, P, B, F, H, S, M, Lr, host_copy_groups(P));
  }
  const HostStaging hs = carve_host(P, B, F, F_out, H, S, M, Lr, U, any_noise, span ? tlay.total : w.total);
  if (!workspace || workspace_bytes < hs.total)
    return fail(h, B200DDSP_WORKSPACE_TOO_SMALL,
                "forward from host inputs needs %zu workspace bytes, got %zu", hs.total, workspace_bytes);
  if (!aligned16(workspace)) return fail(h, B200DDSP_BAD_ALIGN, "workspace must be 16-byte aligned");
  for (int v = 0; v < P; ++v) {
    const b200ddsp_voice& vc = voices_host[v];
    if (!vc.amplitudes || !vc.harmonic_distribution || !vc.inharm_coef || !vc.f0_hz || !vc.magnitudes)
      return fail(h, B200DDSP_BAD_ARGUMENT, "voice %d has a null control tensor", v);
  }
  char* base = (char*)workspace;
  cudaStream_t st = (cudaStream_t)stream;
  cudaStream_t cs = h->copy_stream;
  const size_t BF = (size_t)B * F;

  // device-side voices over the staging area
  b200ddsp_voice dev[B200DDSP_MAX_VOICES];
  for (int v = 0; v < P; ++v) {
    dev[v].amplitudes = (float*)(base + hs.amp) + v * BF;
    dev[v].harmonic_distribution = (float*)(base + hs.hd) + v * BF * H;
    dev[v].inharm_coef = (float*)(base + hs.inh) + v * BF;
    dev[v].f0_hz = (float*)(base + hs.f0) + v * BF * S;
    dev[v].magnitudes = (float*)(base + hs.mags) + v * BF * M;
    dev[v].noise = voices_host[v].noise ? (float*)(base + hs.noise) + (size_t)v * B * N_in : nullptr;
  }
  // one memcpy per run of voices whose host tensors are contiguous (stacked [P,B,F,C] parents)
  auto copy_runs = [&](int v0, int v1, size_t elems, auto host_of, auto dev_of) -> cudaError_t {
    int v = v0;
    while (v < v1) {
      int run = 1;
      while (v + run < v1 && host_of(v + run) == host_of(v) + (size_t)run * elems) ++run;
      cudaError_t e = cudaMemcpyAsync((void*)dev_of(v), host_of(v), (size_t)run * elems * 4,
                                      cudaMemcpyHostToDevice, cs);
      if (e != cudaSuccess) return e;
      v += run;
    }
    return cudaSuccess;
  };
  // the copy stream may not overwrite the staging area before the previous call (ordered on
  // its own stream) has consumed it, nor before work already queued on `st`
  CUDA_TRY(h, cudaEventRecord(h->ev_enter, st));
  CUDA_TRY(h, cudaStreamWaitEvent(cs, h->ev_enter, 0));
  ForwardSync sync{};
  sync.dry_host = dry_out_host;
  // small tensors of every voice first: they are all the phase pass needs
  CUDA_TRY(h, copy_runs(0, P, BF, [&](int v) { return voices_host[v].amplitudes; },
                        [&](int v) { return dev[v].amplitudes; }));
  CUDA_TRY(h, copy_runs(0, P, BF, [&](int v) { return voices_host[v].inharm_coef; },
                        [&](int v) { return dev[v].inharm_coef; }));
  CUDA_TRY(h, copy_runs(0, P, BF * S, [&](int v) { return voices_host[v].f0_hz; },
                        [&](int v) { return dev[v].f0_hz; }));
  CUDA_TRY(h, cudaEventRecord(h->ev_small, cs));
  sync.small_ready = h->ev_small;
  for (int g = 0; g < w.groups.n_groups; ++g) {
    const int v0 = w.groups.first_voice[g], v1 = w.groups.first_voice[g + 1];
    CUDA_TRY(h, copy_runs(v0, v1, BF * H, [&](int v) { return voices_host[v].harmonic_distribution; },
                          [&](int v) { return dev[v].harmonic_distribution; }));
    CUDA_TRY(h, cudaEventRecord(h->ev_group[g], cs));
    sync.group_ready[g] = h->ev_group[g];
  }
  // the impulse responses are small and nothing waits for them until the very end: sending them
  // now keeps them out of the tail after the last large copy
  float* ir_dev = nullptr;
  if (reverb_ir_host) {
    ir_dev = (float*)(base + hs.ir);
    CUDA_TRY(h, cudaMemcpyAsync(ir_dev, reverb_ir_host, (size_t)B * L * 4, cudaMemcpyHostToDevice, cs));
  }
  CUDA_TRY(h, cudaEventRecord(h->ev_ir, cs));
  sync.ir_ready = h->ev_ir;
  const int n_parts = noise_parts(P);
  for (int hf = 0; hf < n_parts; ++hf) {
    const int v0 = P * hf / n_parts, v1 = P * (hf + 1) / n_parts;
    CUDA_TRY(h, copy_runs(v0, v1, BF * M, [&](int v) { return voices_host[v].magnitudes; },
                          [&](int v) { return dev[v].magnitudes; }));
    for (int v = v0; v < v1; ++v)
      if (voices_host[v].noise)
        CUDA_TRY(h, cudaMemcpyAsync((void*)dev[v].noise, voices_host[v].noise, (size_t)B * N_in * 4,
                                    cudaMemcpyHostToDevice, cs));
    CUDA_TRY(h, cudaEventRecord(h->ev_mags[hf], cs));
    sync.mags_ready[hf] = h->ev_mags[hf];
  }

  float* dry_dev = (float*)(base + hs.dry);
  float* wet_dev = reverb_ir_host ? (float*)(base + hs.wet) : nullptr;
  if (span) {
    const TimelineReverbPlan tl = timeline_plan(h, base + hs.core, tlay.rev, ir_dev, B, F_out / seg_frames,
                                                seg_frames * U, L, tail ? to_link(*tail) : Link{});
    if (int rc = forward_core(h, dev, P, nullptr, dry_dev, wet_dev, B, F, H, S, M, L, seed, base + hs.core, w,
                              &sync, st, &si, span->in_first_frame, &tl))
      return rc;
  } else if (int rc = forward_core(h, dev, P, ir_dev, dry_dev, wet_dev, B, F, H, S, M, L, seed, base + hs.core,
                                   w, &sync, st)) {
    return rc;
  }
  if (dry_out_host) CUDA_TRY(h, cudaStreamWaitEvent(st, h->ev_dry_home, 0));   // copied under the reverb (forward_core)
  if (wet_out_host)
    CUDA_TRY(h, cudaMemcpyAsync(wet_out_host, wet_dev, (size_t)B * N * 4, cudaMemcpyDeviceToHost, st));
  return B200DDSP_OK;
}

extern "C" int b200ddsp_forward_polyphonic_host(b200ddsp_handle* h, const b200ddsp_voice* voices_host,
                                                int P, const float* reverb_ir_host,
                                                float* dry_out_host, float* wet_out_host, int B, int F,
                                                int H, int S, int M, int L, uint64_t seed,
                                                void* workspace, size_t workspace_bytes, void* stream) {
  return forward_host_impl(h, voices_host, P, reverb_ir_host, dry_out_host, wet_out_host, B, F, H, S, M, L,
                           seed, nullptr, 0, nullptr, workspace, workspace_bytes, stream);
}

extern "C" int b200ddsp_forward_timeline_host(b200ddsp_handle* h, const b200ddsp_voice* voices_host, int P,
                                              const float* reverb_ir_host, float* dry_out_host,
                                              float* wet_out_host, int B, int F, int H, int S, int M, int L,
                                              int seg_frames, uint64_t seed, const b200ddsp_span* span,
                                              const b200ddsp_link* tail, void* workspace,
                                              size_t workspace_bytes, void* stream) {
  if (!span) return fail(h, B200DDSP_BAD_ARGUMENT, "span is null");
  return forward_host_impl(h, voices_host, P, reverb_ir_host, dry_out_host, wet_out_host, B, F, H, S, M, L,
                           seed, span, seg_frames, tail, workspace, workspace_bytes, stream);
}

// ---------------------------------------------------------------------------------------------
// peer-visible buffers (the inboxes of the span hand-off, link.cuh)
// ---------------------------------------------------------------------------------------------
extern "C" int b200ddsp_peer_alloc(b200ddsp_handle* h, size_t bytes, void** dev_ptr,
                                   unsigned char* ipc_handle64) {
  if (!h) return B200DDSP_BAD_ARGUMENT;
  if (!dev_ptr || !ipc_handle64 || bytes == 0)
    return fail(h, B200DDSP_BAD_ARGUMENT, "peer_alloc: null output or zero size");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "the ABI passes IPC handles as 64 bytes");
  void* p = nullptr;
  CUDA_TRY(h, cudaMalloc(&p, bytes));
  if (cudaMemset(p, 0, bytes) != cudaSuccess) {   // counters of a link start at zero
    cudaFree(p);
    return fail(h, B200DDSP_CUDA_ERROR, "cudaMemset of a peer buffer failed");
  }
  cudaIpcMemHandle_t ipc;
  cudaError_t e = cudaIpcGetMemHandle(&ipc, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    return fail(h, B200DDSP_CUDA_ERROR, "cudaIpcGetMemHandle: %s", cudaGetErrorString(e));
  }
  memcpy(ipc_handle64, &ipc, 64);
  *dev_ptr = p;
  return B200DDSP_OK;
}

extern "C" int b200ddsp_peer_free(b200ddsp_handle* h, void* dev_ptr) {
  if (!h) return B200DDSP_BAD_ARGUMENT;
  if (dev_ptr) CUDA_TRY(h, cudaFree(dev_ptr));
  return B200DDSP_OK;
}

extern "C" int b200ddsp_peer_open(b200ddsp_handle* h, const unsigned char* ipc_handle64, void** peer_ptr) {
  if (!h) return B200DDSP_BAD_ARGUMENT;
  if (!ipc_handle64 || !peer_ptr) return fail(h, B200DDSP_BAD_ARGUMENT, "peer_open: null argument");
  cudaIpcMemHandle_t ipc;
  memcpy(&ipc, ipc_handle64, 64);
  CUDA_TRY(h, cudaIpcOpenMemHandle(peer_ptr, ipc, cudaIpcMemLazyEnablePeerAccess));
  return B200DDSP_OK;
}

extern "C" int b200ddsp_peer_close(b200ddsp_handle* h, void* peer_ptr) {
  if (!h) return B200DDSP_BAD_ARGUMENT;
  if (peer_ptr) CUDA_TRY(h, cudaIpcCloseMemHandle(peer_ptr));
  return B200DDSP_OK;
}

// ---------------------------------------------------------------------------------------------
// measured FP32 peak (bench.py's roofline denominator)
// ---------------------------------------------------------------------------------------------
extern "C" int b200ddsp_measure_fma_rate(b200ddsp_handle* h, int packed, double* lane_ops_per_s,
                                         void* stream) {
  if (!h) return B200DDSP_BAD_ARGUMENT;
  if (!lane_ops_per_s) return fail(h, B200DDSP_BAD_ARGUMENT, "null output");
  cudaStream_t st = (cudaStream_t)stream;
  const int ctas = h->n_sms, threads = 1024, iters = 1 << 15;   // 8 warps per scheduler, ~1 ms
  float* out = nullptr;
  CUDA_TRY(h, cudaMalloc(&out, (size_t)ctas * threads * sizeof(float)));   // not on the synthesis path
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 0.f;
  for (int rep = 0; rep < 4; ++rep) {   // first repetition warms the clocks
    cudaEventRecord(e0, st);
    if (packed) fma_rate_kernel<true><<<ctas, threads, 0, st>>>(out, 0.999f, iters);
    else fma_rate_kernel<false><<<ctas, threads, 0, st>>>(out, 0.999f, iters);
    cudaEventRecord(e1, st);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && (best == 0.f || ms < best)) best = ms;
  }
  const cudaError_t e = cudaGetLastError();
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  if (e != cudaSuccess || best <= 0.f)
    return fail(h, B200DDSP_CUDA_ERROR, "fma_rate_kernel: %s", cudaGetErrorString(e));
  h->launches += 4;
  *lane_ops_per_s = (double)ctas * threads * iters * kUbenchChains * 2.0 / (best * 1e-3);
  return B200DDSP_OK;
}

// ---------------------------------------------------------------------------------------------
// control-rate helpers
// ---------------------------------------------------------------------------------------------
extern "C" int b200ddsp_note_release(b200ddsp_handle* h, const float* active_pitch, float* extended_pitch,
                                     int rows, int F, int in_stride, float release_frames,
                                     void* stream) {
  if (!h) return B200DDSP_BAD_ARGUMENT;
  if (!active_pitch || !extended_pitch) return fail(h, B200DDSP_BAD_ARGUMENT, "null tensor pointer");
  if (rows < 1 || F < 1 || in_stride < 1)
    return fail(h, B200DDSP_BAD_SHAPE, "rows=%d F=%d in_stride=%d", rows, F, in_stride);
  if (!(release_frames >= 0.f))
    return fail(h, B200DDSP_BAD_ARGUMENT, "release_frames=%g must be >= 0", (double)release_frames);
  note_release_kernel<<<(rows + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
      active_pitch, extended_pitch, rows, F, in_stride, release_frames);
  CHECK_LAUNCH(h, "note_release_kernel");
  return B200DDSP_OK;
}

template <int UC, int CL>
static int launch_gru(b200ddsp_handle* h, const float* x_proj, const float* w_hh, const float* b_hh,
                      float* out, int rows, int F, cudaStream_t st) {
  constexpr int u = UC * CL;
  // rows per cluster: as few as keeps every cluster resident at once (the frame loop is latency-bound,
  // so small groups on many SMs win), at most what 1024 threads hold (4 lanes per unit and row pair)
  const int rb_max = 2 * (1024 / (4 * UC)) < 16 ? 2 * (1024 / (4 * UC)) : 16;
  static const int async_exchange = env_int("B200DDSP_GRU_ASYNC", 1);
  auto kernel = (CL > 1 && async_exchange) ? gru_recurrence_kernel<UC, CL, true> : gru_recurrence_kernel<UC, CL, false>;
  auto smem_for = [&](int rb) { return (size_t)(3 * UC * (u + 16) + 2 * rb * u) * sizeof(float) + 16; };   // + two mbarriers
  CUDA_TRY(h, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_for(rb_max)));
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cfg.stream = st;
  // How many clusters of a given shape the GPU holds at once (several small CTAs share an SM) is the driver's
  // to say; B200DDSP_GRU_RB forces a group size for A/B timing.
  static const int forced_rb = env_int("B200DDSP_GRU_RB", 0);
  static int cached_rows = -1, cached_rb = 0;   // the occupancy queries cost host time: once per row count
  int RB = rb_max;
  if (cached_rows == rows) RB = cached_rb;
  else for (int rb = 2; rb <= rb_max; rb *= 2) {
    const int groups_rb = (rows + rb - 1) / rb;
    cfg.gridDim = dim3((unsigned)(groups_rb * CL));
    cfg.blockDim = dim3((unsigned)(UC * (rb / 2) * 4));
    cfg.dynamicSmemBytes = smem_for(rb);
    int resident = 0;
    if (cudaOccupancyMaxActiveClusters(&resident, kernel, &cfg) != cudaSuccess) {
      cudaGetLastError();
      resident = h->n_sms / CL;
    }
    if (forced_rb == rb || (forced_rb == 0 && groups_rb <= resident)) {
      RB = rb;
      break;
    }
  }
  cached_rows = rows;
  cached_rb = RB;
  const int groups = (rows + RB - 1) / RB;
  cfg.gridDim = dim3((unsigned)(groups * CL));
  cfg.blockDim = dim3((unsigned)(UC * (RB / 2) * 4));
  cfg.dynamicSmemBytes = smem_for(RB);
  CUDA_TRY(h, cudaLaunchKernelEx(&cfg, kernel, x_proj, w_hh, b_hh, out, rows, F, RB));
  CHECK_LAUNCH(h, "gru_recurrence_kernel");
  return B200DDSP_OK;
}

extern "C" int b200ddsp_gru_recurrence(b200ddsp_handle* h, const float* x_proj, const float* w_hh,
                                       const float* b_hh, float* out, int rows, int F, int units,
                                       void* stream) {
  if (!h) return B200DDSP_BAD_ARGUMENT;
  if (!x_proj || !w_hh || !b_hh || !out) return fail(h, B200DDSP_BAD_ARGUMENT, "null tensor pointer");
  if (rows < 1 || F < 1) return fail(h, B200DDSP_BAD_SHAPE, "rows=%d F=%d", rows, F);
  if ((((uintptr_t)x_proj | (uintptr_t)w_hh | (uintptr_t)b_hh | (uintptr_t)out) & 3) != 0)
    return fail(h, B200DDSP_BAD_ALIGN, "tensors must be float32-aligned");
  cudaStream_t st = (cudaStream_t)stream;
  switch (units) {
    case 64: return launch_gru<64, 1>(h, x_proj, w_hh, b_hh, out, rows, F, st);
    case 128: return launch_gru<32, 4>(h, x_proj, w_hh, b_hh, out, rows, F, st);
    case 192: return launch_gru<24, 8>(h, x_proj, w_hh, b_hh, out, rows, F, st);
    case 256: return launch_gru<32, 8>(h, x_proj, w_hh, b_hh, out, rows, F, st);
    default:
      return fail(h, B200DDSP_UNSUPPORTED_CONFIG, "GRU of %d units (supported: 64, 128, 192, 256)", units);
  }
}

// ---------------------------------------------------------------------------------------------
// host-side front end
// ---------------------------------------------------------------------------------------------
#include "midi_host.inl"
