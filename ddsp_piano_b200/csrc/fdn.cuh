// Feedback-delay-network reverb: impulse-response generator (reference
// modules/fdn_reverb.py:178-360, FeedbackDelayNetwork.get_late_ir / get_ir).
// The IR feeds the convolution reverb (reverb.cuh) unchanged.
//
// The network (8 delay lines, Householder mixing, one-pole reverberation-time control, four
// Schroeder allpasses per line) is sampled on n/2+1 frequencies, n = 2 * sampling_rate:
//     H[k] = c^T D_k (I - F_k D_k)^-1 b ,      late_ir = irfft(H) ,      ir = early_ir + late_ir
// fdn_transfer_kernel  one thread per (frequency bin, batch row): builds the lines x lines complex system
//                      in float32 complex arithmetic and solves it by Gaussian elimination with
//                      partial pivoting.  Phase angles are formed exactly as the reference forms
//                      them -- float32 w_k = (2pi * k) / n, float32 product w_k * delay -- because
//                      at delays of ~900 samples that rounding (2.4e-4 rad) is part of its output.
// fdn_irfft_kernel     n is 2 * sampling_rate (48 000 at 24 kHz: not a power of two), so the
//                      inverse real DFT is evaluated directly: one thread per output sample walks
//                      the bins with a double-precision phasor recurrence (no table, no memory
//                      traffic; 24 001 x 48 000 x 6 DFMA = 0.4 ms of FP64 on a B200, once per model).
#pragma once
#include "common.cuh"

namespace b200ddsp {

constexpr int kFdnLines = 8;        // at most; the kernel is instantiated for 6 (configs/ENSTDkCl-*.gin) and 8
constexpr int kFdnAllpass = 4;

struct FdnArgs {
  const float* input_gain;      // [B, lines]
  const float* output_gain;     // [B, lines]
  const float* gain_allpass;    // [B, lines, 4]
  const float* delays_allpass;  // [B, lines, 4]
  const float* time_rev_0_sec;  // [B]
  const float* alpha_tone;      // [B]
  const float* early_ir;        // [B, E]
  float2* H;                    // [B, n/2 + 1]
  float* ir;                    // [B, n]
  float delay_values[kFdnLines];
  float sampling_rate;
  int n, E, B;
};

__device__ __forceinline__ float2 c_mul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 c_div(float2 a, float2 b) {
  const float den = b.x * b.x + b.y * b.y;
  return make_float2((a.x * b.x + a.y * b.y) / den, (a.y * b.x - a.x * b.y) / den);
}
__device__ __forceinline__ float2 c_expi(float theta) {   // exp(i theta), accurate for large theta
  float s, c;
  sincosf(theta, &s, &c);
  return make_float2(c, s);
}

template <int NL>
__global__ void __launch_bounds__(128) fdn_transfer_kernel(const FdnArgs a) {
  const int nb = a.n / 2 + 1;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (k >= nb) return;
  // w_k exactly as the reference forms it: float32 (2 pi * k) / n      (fdn_reverb.py:233-238)
  const float wk = __fdiv_rn(__fmul_rn(kTwoPi, (float)k), (float)a.n);
  const float2 zinv = c_expi(-wk);
  const float t0 = a.time_rev_0_sec[b];
  const float at0 = __fmul_rn(a.alpha_tone[b], t0);

  float2 D[NL], lowpass[NL], allpass[NL];
#pragma unroll
  for (int d = 0; d < NL; ++d) {
    const float dv = a.delay_values[d];
    const float whole = floorf(dv);
    // integer delay + first-order allpass interpolation of the fractional part   (:241-262)
    const float2 z_d = c_expi(-__fmul_rn(wk, whole));
    const float frac = dv - whole;
    const float eta = (1.0f - frac) / (1.0f + frac);
    const float2 interp = c_div(make_float2(eta + zinv.x, zinv.y),
                                make_float2(1.0f + eta * zinv.x, eta * zinv.y));
    D[d] = c_mul(z_d, interp);
    // one-pole low-pass: frequency dependent reverberation time   (:266-288)
    float sum_ap = 0.f;
    float2 ap = make_float2(1.f, 0.f);
#pragma unroll
    for (int j = 0; j < kFdnAllpass; ++j) {
      const float da = a.delays_allpass[((size_t)b * NL + d) * kFdnAllpass + j];
      const float ga = a.gain_allpass[((size_t)b * NL + d) * kFdnAllpass + j];
      sum_ap += da;
      const float2 zdel = c_expi(__fmul_rn(wk, da));                        // :298
      ap = c_mul(ap, c_div(make_float2(1.0f + ga * zdel.x, ga * zdel.y),
                           make_float2(ga + zdel.x, zdel.y)));              // :301-304
    }
    allpass[d] = ap;
    const float delay_sec = (dv + sum_ap) / a.sampling_rate;
    const float kk = powf(10.0f, -3.0f * delay_sec / t0);
    const float kpi = powf(10.0f, -3.0f * delay_sec / at0);
    const float g = 2.0f * kk * kpi / (kk + kpi);
    const float p = (kk - kpi) / (kk + kpi);
    lowpass[d] = c_div(make_float2(g, 0.f), make_float2(1.0f - p * zinv.x + 1e-8f, -p * zinv.y));
  }
  // M = I - F D,  F = diag(lowpass) (0.5 * 11^T - I) diag(allpass);  augmented with b   (:310-330)
  float2 M[NL][NL + 1];
#pragma unroll
  for (int i = 0; i < NL; ++i) {
#pragma unroll
    for (int j = 0; j < NL; ++j) {
      const float mix = (i == j) ? -0.5f : 0.5f;
      const float2 f = c_mul(c_mul(lowpass[i], make_float2(mix, 0.f)), allpass[j]);
      const float2 fd = c_mul(f, D[j]);
      M[i][j] = make_float2((i == j ? 1.0f : 0.0f) - fd.x, -fd.y);
    }
    M[i][NL] = make_float2(a.input_gain[(size_t)b * NL + i], 0.f);
  }
  // Gaussian elimination with partial pivoting (row swaps done by value: fully unrolled, no
  // dynamic indexing)
#pragma unroll
  for (int c = 0; c < NL; ++c) {
#pragma unroll
    for (int r = c + 1; r < NL; ++r) {
      const float mc = M[c][c].x * M[c][c].x + M[c][c].y * M[c][c].y;
      const float mr = M[r][c].x * M[r][c].x + M[r][c].y * M[r][c].y;
      if (mr > mc) {
#pragma unroll
        for (int j = c; j <= NL; ++j) {
          const float2 t = M[c][j];
          M[c][j] = M[r][j];
          M[r][j] = t;
        }
      }
    }
    const float2 inv = c_div(make_float2(1.f, 0.f), M[c][c]);
#pragma unroll
    for (int r = c + 1; r < NL; ++r) {
      const float2 f = c_mul(M[r][c], inv);
#pragma unroll
      for (int j = c + 1; j <= NL; ++j) {
        const float2 t = c_mul(f, M[c][j]);
        M[r][j] = make_float2(M[r][j].x - t.x, M[r][j].y - t.y);
      }
    }
  }
  float2 x[NL];
#pragma unroll
  for (int i = NL - 1; i >= 0; --i) {
    float2 s = M[i][NL];
#pragma unroll
    for (int j = i + 1; j < NL; ++j) {
      const float2 t = c_mul(M[i][j], x[j]);
      s = make_float2(s.x - t.x, s.y - t.y);
    }
    x[i] = c_div(s, M[i][i]);
  }
  // H = c^T D x      (:322-334)
  float2 Hk = make_float2(0.f, 0.f);
#pragma unroll
  for (int i = 0; i < NL; ++i) {
    const float2 t = c_mul(D[i], x[i]);
    const float c = a.output_gain[(size_t)b * NL + i];
    Hk.x += c * t.x;
    Hk.y += c * t.y;
  }
  a.H[(size_t)b * nb + k] = Hk;
}

// ir[t] = early[t] + (1/n) (Re H_0 + (-1)^t Re H_{n/2} + 2 sum_{k=1}^{n/2-1} Re(H_k e^{+2 pi i k t / n}))
constexpr int kFdnTile = 1024;

__global__ void __launch_bounds__(256) fdn_irfft_kernel(const FdnArgs a) {
  __shared__ float2 Hs[kFdnTile];
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  const int half = a.n / 2;                  // n is even: n = int(2 * sampling_rate)
  const float2* H = a.H + (size_t)b * (half + 1);
  double rs, rc;
  sincospi(2.0 * (double)(t % a.n) / (double)a.n, &rs, &rc);   // r = e^{2 pi i t / n}
  double wc = rc, ws = rs;                                      // phasor for k = 1
  double acc = 0.0;
  for (int k0 = 1; k0 < half; k0 += kFdnTile) {
    const int len = min(kFdnTile, half - k0);
    __syncthreads();
    for (int i = threadIdx.x; i < len; i += blockDim.x) Hs[i] = H[k0 + i];
    __syncthreads();
    for (int i = 0; i < len; ++i) {
      const float2 h = Hs[i];
      acc = fma((double)h.x, wc, acc);
      acc = fma(-(double)h.y, ws, acc);
      const double nc = fma(wc, rc, -ws * rs);
      ws = fma(wc, rs, ws * rc);
      wc = nc;
    }
  }
  if (t >= a.n) return;
  double v = 2.0 * acc + (double)H[0].x + ((t & 1) ? -1.0 : 1.0) * (double)H[half].x;
  float out = (float)(v / (double)a.n);
  if (t < a.E) out = __fadd_rn(a.early_ir[(size_t)b * a.E + t], out);      // :352-360
  a.ir[(size_t)b * a.n + t] = out;
}

}  // namespace b200ddsp
