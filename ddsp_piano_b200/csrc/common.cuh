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

// cos of a float32 phase of any magnitude (inference=False: the plain cumsum reaches 1e5 rad):
// reduce modulo the true 2 pi in double precision, then the hardware cosine.
__device__ __forceinline__ float cos_large(float x) {
  const double xd = (double)x;
  const double n = rint(xd * 0.15915494309189535);
  return __cosf((float)fma(-n, 6.283185307179586, xd));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ddsp.core.exp_sigmoid / reference exp_tanh (modules/inharm_synth.py:8-17) / identity.
//   exp_sigmoid(x) = 2 * sigmoid(x)^ln(10) + 1e-7 = 2^(1 - ln10 * log2(1 + 2^(-x log2 e))) + 1e-7
//   exp_tanh(x)    = 2 * (0.5 (tanh x + 1))^ln(10) + 1e-7, and 0.5 (tanh x + 1) = sigmoid(2x)
// evaluated with the hardware exp2/log2 approximations: relative error below 1e-6, against a
// parity budget of 1e-4 on the audio (amplitudes are not accumulated, unlike the phase).
// One MUFU each: the flush-to-zero forms need no range scaling around the instruction, and nothing here
// depends on a subnormal (e^-z underflows to 0 where sigmoid is 1 to within 1e-38; the result has 1e-7 added).
__device__ __forceinline__ float ex2_ftz(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2_ftz(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// The same function with its mode decided once, outside the per-element loop: no branch per element.
struct ScaleFn {
  float c;        // -log2(e) for exp_sigmoid, twice that for exp_tanh (the doubling is exact)
  bool identity;
  __device__ __forceinline__ explicit ScaleFn(int fn)
      : c(fn == 1 ? -2.0f * 1.4426950408889634f : -1.4426950408889634f), identity(fn == 2) {}
  __device__ __forceinline__ float operator()(float x) const {
    const float l = lg2_ftz(1.0f + ex2_ftz(x * c));
    const float y = ex2_ftz(__fmaf_rn(-2.302585092994046f, l, 1.0f)) + 1e-7f;
    return identity ? x : y;
  }
};

__device__ __forceinline__ float apply_scale_fn(float x, int fn) {
  if (fn == 2) return x;
  const float kLog2e = 1.4426950408889634f, kLn10 = 2.302585092994046f;
  const float z = (fn == 1) ? 2.0f * x : x;
  const float t = ex2_ftz(-z * kLog2e);            // e^-z, +inf for very negative z
  const float l = lg2_ftz(1.0f + t);               // -log2(sigmoid(z))
  return ex2_ftz(__fmaf_rn(-kLn10, l, 1.0f)) + 1e-7f;
}

// ---- asynchronous bulk copy + mbarrier (PTX; SASS: UBLKCP / SYNCS) -----------------------------------
__device__ __forceinline__ unsigned int smem_u32(const void* p) {
  return (unsigned int)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(void* bar, unsigned int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(void* bar, unsigned int bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(void* bar, unsigned int parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
// global -> shared, `bytes` a multiple of 16, both addresses 16-byte aligned; completion is counted on `bar`
__device__ __forceinline__ void bulk_load(void* dst, const void* src, unsigned int bytes, void* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

}  // namespace b200ddsp
