// Measured FP32 peak for bench.py's roofline (the oscillator bank is bound by the FMA pipe, SURVEY
// fact 5, and MEASURED_PEAKS.json holds no FP32 figure): independent FFMA2 / FFMA chains on every SM.
#pragma once
#include "common.cuh"

namespace b200ddsp {

constexpr int kUbenchChains = 8;

template <bool PACKED>
__global__ void __launch_bounds__(1024) fma_rate_kernel(float* out, float s, int iters) {
  float2 p[kUbenchChains];
#pragma unroll
  for (int i = 0; i < kUbenchChains; ++i) p[i] = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f - i);
  const float2 s2 = make_float2(s, s), c2 = make_float2(1e-3f, 2e-3f);
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < kUbenchChains; ++i) {
      if (PACKED) {
        p[i] = __ffma2_rn(p[i], s2, c2);
      } else {
        p[i].x = __fmaf_rn(p[i].x, s, 1e-3f);
        p[i].y = __fmaf_rn(p[i].y, s, 2e-3f);
      }
    }
  }
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < kUbenchChains; ++i) acc += p[i].x + p[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

}  // namespace b200ddsp
