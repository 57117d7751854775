// Shared helpers for the b200ddsp kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define B200DDSP_MAX_VOICES_INTERNAL 64

namespace b200ddsp {

constexpr float kTwoPi = 6.283185307179586f;        // float32(2*pi), what TF/NumPy use
constexpr float kInvTwoPi = 0.15915494309189535f;
constexpr float kRoundMagic = 12582912.0f;           // 1.5 * 2^23: fma(x, c, magic) - magic = rint(x*c)
constexpr int kAngularChunk = 1000;                  // ddsp.core.angular_cumsum chunk_size
constexpr int kWarp = 32;

// floormod(x, float32(2*pi)) exactly as TF's FloorMod / np.mod compute it (std::fmod plus
// sign fix), without a loop: n = rint(x/2pi) by the magic-number trick, r = fma(-n, 2pi, x)
// is exact for |x| < 1e6 (exhaustively checked on the host against fmodf), then fold to [0, 2pi).
// r itself, in [-pi, pi], is what the oscillator feeds to cos.
__device__ __forceinline__ float wrap_to_pi(float x) {
  float t = __fmaf_rn(x, kInvTwoPi, kRoundMagic);
  float n = __fadd_rn(t, -kRoundMagic);
  return __fmaf_rn(-n, kTwoPi, x);
}

__device__ __forceinline__ float floormod_two_pi_fast(float x) {
  float r = wrap_to_pi(x);
  return r < 0.f ? __fadd_rn(r, kTwoPi) : r;
}

// General-range version (any finite x): std::fmod semantics + sign fix.
__device__ __forceinline__ float floormod_two_pi(float x) {
  if (fabsf(x) < 1.0e6f) return floormod_two_pi_fast(x);
  float m = fmodf(x, kTwoPi);
  return (m != 0.f && m < 0.f) ? __fadd_rn(m, kTwoPi) : m;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ddsp.core.exp_sigmoid / reference exp_tanh (modules/inharm_synth.py:8-17) / identity.
__device__ __forceinline__ float apply_scale_fn(float x, int fn) {
  const float kLog10 = 2.302585092994046f;
  if (fn == 0) {
    float s = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x)));
    return __fadd_rn(__fmul_rn(2.0f, powf(s, kLog10)), 1e-7f);
  } else if (fn == 1) {
    float s = __fmul_rn(0.5f, __fadd_rn(tanhf(x), 1.0f));
    return __fadd_rn(__fmul_rn(2.0f, powf(s, kLog10)), 1e-7f);
  }
  return x;
}

}  // namespace b200ddsp
