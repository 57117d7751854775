// Control-rate kernels: MultiInharmonic.get_controls and FilteredNoise.get_controls.
#pragma once
#include "common.cuh"

namespace b200ddsp {

// Per-voice input tensors (blockIdx.y = voice); outputs are stacked voice-major [P*B, F, .].
struct AdditiveControlsPtrs {
  const float* amp_in[B200DDSP_MAX_VOICES_INTERNAL];     // [B, F, 1]
  const float* hd_in[B200DDSP_MAX_VOICES_INTERNAL];      // [B, F, H]
  const float* inharm_in[B200DDSP_MAX_VOICES_INTERNAL];  // [B, F, 1]
  const float* f0_in[B200DDSP_MAX_VOICES_INTERNAL];      // [B, F, S]
};

struct AdditiveControlsArgs {
  float* amp_out;          // [P*B, F, 1]
  float* hd_out;           // [P*B, F, H]
  float* shifts_out;       // [P*B, F, H]
  float* f0_out;           // [P*B, F, S] copy (get_controls returns f0_hz unchanged), or nullptr
  unsigned char* na_frame; // [P*B, F] number of leading 16-partial half-groups that can sound
                           // (additive_fast.cuh), or nullptr
  unsigned short* cut_index;  // [P*B, F] number of partials below Nyquist = index of the first one the cut of
                              // get_controls removes: written by the prep kernel, read by the hd kernel
                              // (which otherwise recomputes every partial's frequency), or nullptr
  int n_frames_voice;      // B * F
  int H, S;
  float nyquist, min_frequency;
  int scale_fn, normalize_after, normalize_below;
};

// MultiInharmonic.get_controls (modules/inharm_synth.py:254-270 over :167-219) is split in two
// kernels along its data dependencies, so that the host-input pipeline can start the phase pass
// before the (large) harmonic_distribution tensors have arrived:
//   additive_prep_kernel  needs f0_hz, inharm_coef, amplitudes -> amplitudes, harmonic_shifts,
//                         f0 copy, per-frame liveness of the partial groups
//   additive_hd_kernel    needs harmonic_distribution (+ f0_hz, inharm_coef for the Nyquist cut)
//                         -> harmonic_distribution
// One warp per kFramesPerWarp (row, frame)s; lanes stride over partials (HP per lane); all loads
// of the warp's frames are issued before any arithmetic.
constexpr int kMaxHarmonicsPerLane = 8;  // H <= 256
constexpr int kFramesPerWarp = 2;

__device__ __forceinline__ float inharm_factor(int h, float binh) {
  const float n = (float)(h + 1);
  return sqrtf(__fadd_rn(__fmul_rn(__fmul_rn(n, n), binh), 1.0f));   // get_inharmonic_freq :37-39
}

template <int HP>
__global__ void __launch_bounds__(256) additive_prep_kernel(const AdditiveControlsArgs a,
                                                            const AdditiveControlsPtrs p) {
  const int lane = threadIdx.x & 31;
  const int vf0 = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * kFramesPerWarp;
  if (vf0 >= a.n_frames_voice) return;
  const int v = blockIdx.y;
  float f0s[kFramesPerWarp], inh[kFramesPerWarp], amps[kFramesPerWarp];
#pragma unroll
  for (int e = 0; e < kFramesPerWarp; ++e) {
    const int vf = min(vf0 + e, a.n_frames_voice - 1);
    f0s[e] = __ldg(p.f0_in[v] + (size_t)vf * a.S);                  // f0_hz[..., 0:1]   (:262)
    inh[e] = __ldg(p.inharm_in[v] + vf);
    amps[e] = __ldg(p.amp_in[v] + vf);
  }
#pragma unroll
  for (int e = 0; e < kFramesPerWarp; ++e) {
    const int vf = vf0 + e;
    if (vf >= a.n_frames_voice) break;
    const size_t rf = (size_t)v * a.n_frames_voice + vf;            // stacked output frame
    const float f0 = f0s[e];
    const float binh = fmaxf(inh[e], 0.f);                          // :183
    float amp = apply_scale_fn(amps[e], a.scale_fn);                // :184-186
    if (a.normalize_below) amp = __fmul_rn(amp, (f0 > a.min_frequency) ? 1.0f : 0.0f);   // :207-208
    // :269; by a power of two the division is a multiplication by its (exact) reciprocal
    const float amp_final = ((a.S & (a.S - 1)) == 0)
                                ? __fmul_rn(amp, __int_as_float(0x3f800000 - ((31 - __clz(a.S)) << 23)))
                                : __fdiv_rn(amp, (float)a.S);
    if (a.f0_out != nullptr && lane < a.S)
      a.f0_out[rf * a.S + lane] = __ldg(p.f0_in[v] + (size_t)vf * a.S + lane);
    int na = 0, n_below = 0;
#pragma unroll
    for (int j = 0; j < HP; ++j) {
      const int h = lane + 32 * j;
      bool can_sound = false;
      if (h < a.H) {
        const float fac = inharm_factor(h, binh);
        const float fi = __fmul_rn(__fmul_rn(f0, (float)(h + 1)), fac);      // :42
        a.shifts_out[rf * a.H + h] = __fadd_rn(fac, -1.0f);                  // :44
        // the Nyquist cut of get_controls (:200-205) zeroes this partial in this frame
        can_sound = !a.normalize_below || !(fi >= a.nyquist);
      }
      const unsigned m = __ballot_sync(0xffffffffu, can_sound);
      if (m) na = (m >> 16) ? 2 * j + 2 : 2 * j + 1;
      // every operation of fi is monotonic in h (f0 > 0; for f0 <= 0 nothing is ever cut), so the partials
      // below Nyquist are the first n_below ones (without the cut: all H of them)
      n_below += __popc(m);
    }
    if (lane == 0) {
      a.amp_out[rf] = amp_final;
      if (a.na_frame != nullptr) a.na_frame[rf] = (unsigned char)(amp_final != 0.f ? na : 0);
      if (a.cut_index != nullptr) a.cut_index[rf] = (unsigned short)n_below;
    }
  }
}

template <int HP>
__global__ void __launch_bounds__(256) additive_hd_kernel(const AdditiveControlsArgs a,
                                                          const AdditiveControlsPtrs p) {
  const int lane = threadIdx.x & 31;
  const int vf0 = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * kFramesPerWarp;
  if (vf0 >= a.n_frames_voice) return;
  const int v = blockIdx.y;
  float raw[kFramesPerWarp][HP], f0s[kFramesPerWarp], inh[kFramesPerWarp];
  const ScaleFn scale(a.scale_fn);
#pragma unroll
  for (int e = 0; e < kFramesPerWarp; ++e) {
    const int vf = min(vf0 + e, a.n_frames_voice - 1);
    f0s[e] = __ldg(p.f0_in[v] + (size_t)vf * a.S);
    inh[e] = __ldg(p.inharm_in[v] + vf);
#pragma unroll
    for (int j = 0; j < HP; ++j) {
      const int h = lane + 32 * j;
      raw[e][j] = (h < a.H) ? __ldg(p.hd_in[v] + (size_t)vf * a.H + h) : 0.f;
    }
  }
#pragma unroll
  for (int e = 0; e < kFramesPerWarp; ++e) {
    const int vf = vf0 + e;
    if (vf >= a.n_frames_voice) break;
    const size_t rf = (size_t)v * a.n_frames_voice + vf;
    const float f0 = f0s[e];
    const float binh = fmaxf(inh[e], 0.f);
    float d[HP];
    bool cut[HP];
    float sum = 0.f;
    // partials below Nyquist = the first n_below (their frequencies are monotonic in h, see the prep kernel,
    // which leaves the count in cut_index; without it the warp counts them here)
    int n_below = 0;
    if (a.cut_index != nullptr) {
      n_below = (int)a.cut_index[rf];
    } else {
#pragma unroll
      for (int j = 0; j < HP; ++j) {
        const int h = lane + 32 * j;
        bool below = false;
        if (h < a.H) below = !(__fmul_rn(__fmul_rn(f0, (float)(h + 1)), inharm_factor(h, binh)) >= a.nyquist);
        n_below += __popc(__ballot_sync(0xffffffffu, below));
      }
    }
#pragma unroll
    for (int j = 0; j < HP; ++j) {
      const int h = lane + 32 * j;
      d[j] = 0.f;
      cut[j] = h >= n_below;
      if (h < a.H) {
        d[j] = scale(raw[e][j]);                                    // :184-186
        sum += d[j];
      }
    }
    // safe_divide by the frame's sum: one IEEE reciprocal per frame and a multiply per partial (within an ulp
    // of the quotient; these are amplitudes, not phases: nothing accumulates)
    if (a.normalize_after == 0) {                                    // :194-198
      sum = warp_sum(sum);
      const float inv = __frcp_rn((sum == 0.f) ? 1e-7f : sum);
#pragma unroll
      for (int j = 0; j < HP; ++j) d[j] = __fmul_rn(d[j], inv);
    }
    if (a.normalize_below) {                                         // :200-205
      sum = 0.f;
#pragma unroll
      for (int j = 0; j < HP; ++j) {
        if (cut[j]) d[j] = 0.f;
        sum += d[j];
      }
    }
    if (a.normalize_after == 1) {                                    // :210-214 (2: never normalised)
      sum = warp_sum(sum);
      const float inv = __frcp_rn((sum == 0.f) ? 1e-7f : sum);
#pragma unroll
      for (int j = 0; j < HP; ++j) d[j] = __fmul_rn(d[j], inv);
    }
#pragma unroll
    for (int j = 0; j < HP; ++j) {
      const int h = lane + 32 * j;
      if (h < a.H) a.hd_out[rf * a.H + h] = d[j];
    }
  }
}

// FilteredNoise.get_controls: scale_fn(magnitudes + initial_bias).
__global__ void __launch_bounds__(256) noise_controls_kernel(const float* __restrict__ in,
                                                             float* __restrict__ out, size_t n,
                                                             float bias, int scale_fn) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    float x = in[i];
    out[i] = (scale_fn == 2) ? x : apply_scale_fn(__fadd_rn(x, bias), scale_fn);
  }
}

// SurrogateAdditive.get_controls, decay part (surrogate_synth.py:163-171): clip to [1e-5, 1], and
// 1 (no decay) for partials whose inharmonic frequency is at or above Nyquist.
__global__ void __launch_bounds__(256) surrogate_decays_kernel(const float* __restrict__ decays,
                                                               const float* __restrict__ inharm_coef,
                                                               const float* __restrict__ f0_hz,
                                                               float* __restrict__ out, int n_frames,
                                                               int H, float nyquist) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)n_frames * H) return;
  const int fr = (int)(i / H), h = (int)(i - (size_t)fr * H);
  const float binh = fmaxf(__ldg(inharm_coef + fr), 0.f);
  const float fi = __fmul_rn(__fmul_rn(__ldg(f0_hz + fr), (float)(h + 1)), inharm_factor(h, binh));
  const float d = fmaxf(fminf(__ldg(decays + i), 1.f), 1e-5f);
  out[i] = (fi >= nyquist) ? 1.f : d;
}

}  // namespace b200ddsp
