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
  unsigned char* na_frame; // [P*B, F] 1 + highest live 32-partial group (additive_fast.cuh), or nullptr
  int n_frames_voice;      // B * F
  int H, S;
  float nyquist, min_frequency;
  int scale_fn, normalize_after, normalize_below;
};

// One warp per kFramesPerWarp (row, frame)s; lanes stride over partials (HP per lane).  Follows
// modules/inharm_synth.py:167-219 (InHarmonic.get_controls) as called from :254-270.  All loads
// of the warp's frames are issued before any arithmetic: the kernel is a 175 MB stream.
constexpr int kMaxHarmonicsPerLane = 8;  // H <= 256
constexpr int kFramesPerWarp = 2;

template <int HP>
__global__ void __launch_bounds__(256) additive_controls_kernel(const AdditiveControlsArgs a,
                                                                const AdditiveControlsPtrs p) {
  const int lane = threadIdx.x & 31;
  const int vf0 = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * kFramesPerWarp;
  if (vf0 >= a.n_frames_voice) return;
  const int v = blockIdx.y;
  float raw[kFramesPerWarp][HP], f0s[kFramesPerWarp], inh[kFramesPerWarp], amps[kFramesPerWarp];
#pragma unroll
  for (int e = 0; e < kFramesPerWarp; ++e) {
    const int vf = min(vf0 + e, a.n_frames_voice - 1);
    f0s[e] = __ldg(p.f0_in[v] + (size_t)vf * a.S);                  // f0_hz[..., 0:1]   (:262)
    inh[e] = __ldg(p.inharm_in[v] + vf);
    amps[e] = __ldg(p.amp_in[v] + vf);
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
    const size_t rf = (size_t)v * a.n_frames_voice + vf;            // stacked output frame
    const float f0 = f0s[e];
    const float binh = fmaxf(inh[e], 0.f);                          // :183
    float amp = apply_scale_fn(amps[e], a.scale_fn);                // :184-186
    if (a.f0_out != nullptr && lane < a.S)
      a.f0_out[rf * a.S + lane] = __ldg(p.f0_in[v] + (size_t)vf * a.S + lane);
    float d[HP], fi[HP];
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < HP; ++j) {
      const int h = lane + 32 * j;
      d[j] = 0.f;
      fi[j] = 0.f;
      if (h < a.H) {
        const float n = (float)(h + 1);
        const float fac = sqrtf(__fadd_rn(__fmul_rn(__fmul_rn(n, n), binh), 1.0f));   // :37-39
        fi[j] = __fmul_rn(__fmul_rn(f0, n), fac);                            // :42
        a.shifts_out[rf * a.H + h] = __fadd_rn(fac, -1.0f);                  // :44
        d[j] = apply_scale_fn(raw[e][j], a.scale_fn);
        sum += d[j];
      }
    }
    if (!a.normalize_after) {                                        // :194-198
      sum = warp_sum(sum);
      const float den = (sum == 0.f) ? 1e-7f : sum;
#pragma unroll
      for (int j = 0; j < HP; ++j) d[j] = __fdiv_rn(d[j], den);
    }
    if (a.normalize_below) {                                         // :200-208
      sum = 0.f;
#pragma unroll
      for (int j = 0; j < HP; ++j) {
        if (fi[j] >= a.nyquist) d[j] = 0.f;
        sum += d[j];
      }
      amp = __fmul_rn(amp, (f0 > a.min_frequency) ? 1.0f : 0.0f);
    }
    if (a.normalize_after) {                                         // :210-214
      sum = warp_sum(sum);
      const float den = (sum == 0.f) ? 1e-7f : sum;
#pragma unroll
      for (int j = 0; j < HP; ++j) d[j] = __fdiv_rn(d[j], den);
    }
    int na = 0;
#pragma unroll
    for (int j = 0; j < HP; ++j) {
      const int h = lane + 32 * j;
      if (h < a.H) a.hd_out[rf * a.H + h] = d[j];
      if (a.na_frame != nullptr) {
        const bool live = (h < a.H) && (d[j] != 0.f);
        if (__ballot_sync(0xffffffffu, live)) na = j + 1;
      }
    }
    const float amp_final = __fdiv_rn(amp, (float)a.S);              // :269
    if (lane == 0) {
      a.amp_out[rf] = amp_final;
      if (a.na_frame != nullptr) a.na_frame[rf] = (unsigned char)(amp_final != 0.f ? na : 0);
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

}  // namespace b200ddsp
