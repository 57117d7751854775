// Microbenchmark: how fast can an SM issue the instruction mix of the oscillator bank's synthesis pass,
// with its dependency structure but without memory, branches or reductions?  Per pair of oscillators and
// sample (additive_fast.cuh::osc_group_h, general / nocheck variant):
//   chain   m = fma2(g, lerp, 0); f = add2(F, m); x = mul2(f, 2pi); lo = mul2(x, r_lo); om = fma2(x, r, lo);
//           ph = add2(ph, om)
//   synth   amp = fma2(dA, w, A); x2 = add2(ph, off); [every 4th sample: n = add2(fma2(x2, 1/2pi, magic), -magic)]
//           r = fma2(n, -2pi, x2); c = (cos(r.x), cos(r.y)) [2 FMUL + 2 MUFU]; acc = fma2(amp, c, acc)
// NP pairs per thread (3 = the 6-chain bucket), W warps per scheduler.  Prints FMA-pipe lane-cycles per
// oscillator-sample actually achieved against the 11 that the mix needs at 128 lanes per SM and clock.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o synth_mix synth_mix.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int NP, bool WITH_COS>
__global__ void __launch_bounds__(1024) k(float* out, const float* in, int iters, long long* cyc) {
  float2 ph[NP], F[NP], g[NP], off[NP], A[NP], dA[NP], nw[NP], acc[4];
  for (int i = 0; i < NP; ++i) {
    const float b = in[threadIdx.x % 32 + i];
    ph[i] = make_float2(b, b * 0.5f); F[i] = make_float2(440.f * (i + 1), 441.f * (i + 1));
    g[i] = make_float2(b * 1e-3f, b * 2e-3f); off[i] = make_float2(b, 1.f + b); A[i] = make_float2(0.1f, 0.2f);
    dA[i] = make_float2(1e-3f, -1e-3f); nw[i] = make_float2(0.f, 0.f);
  }
  for (int i = 0; i < 4; ++i) acc[i] = make_float2(0.f, 0.f);
  const float2 two_pi = make_float2(6.2831855f, 6.2831855f), r2 = make_float2(4.1666666e-5f, 4.1666666e-5f),
               rlo = make_float2(1e-12f, 1e-12f), inv2pi = make_float2(0.15915494f, 0.15915494f),
               magic = make_float2(12582912.f, 12582912.f), nmagic = make_float2(-12582912.f, -12582912.f),
               ntwo_pi = make_float2(-6.2831855f, -6.2831855f), zero = make_float2(0.f, 0.f);
  float lerp = in[0] * 1e-3f, w = in[1] * 1e-3f;
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      lerp += 1.0f / 96; w += 1.0f / 96;
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        const float2 m = __ffma2_rn(g[i], make_float2(lerp, lerp), zero);
        const float2 f = __fadd2_rn(F[i], m);
        const float2 x = __fmul2_rn(f, two_pi);
        const float2 om = __ffma2_rn(x, r2, __fmul2_rn(x, rlo));
        ph[i] = __fadd2_rn(ph[i], om);
        if (WITH_COS) {
          const float2 amp = __ffma2_rn(dA[i], make_float2(w, w), A[i]);
          const float2 x2 = __fadd2_rn(ph[i], off[i]);
          if (j == 0) nw[i] = __fadd2_rn(__ffma2_rn(x2, inv2pi, magic), nmagic);
          const float2 r = __ffma2_rn(nw[i], ntwo_pi, x2);
          acc[j] = __ffma2_rn(amp, make_float2(__cosf(r.x), __cosf(r.y)), acc[j]);
        }
      }
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
  for (int i = 0; i < NP; ++i) s += ph[i].x + ph[i].y;
  for (int i = 0; i < 4; ++i) s += acc[i].x + acc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int NP, bool WITH_COS>
void run(int warps_per_smsp) {
  float *out, *in; long long* cyc;
  const int threads = 128 * warps_per_smsp, iters = 4096;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&in, 1024); cudaMalloc(&cyc, 8);
  cudaMemset(in, 0, 1024);
  k<NP, WITH_COS><<<148, threads>>>(out, in, iters, cyc);
  k<NP, WITH_COS><<<148, threads>>>(out, in, iters, cyc);
  cudaDeviceSynchronize();
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  // oscillator-samples per SM: threads * iters * 4 samples * 2 NP oscillators; lane-cycles available: c * 128
  const double osc = (double)threads * iters * 4 * 2 * NP;
  const double need = WITH_COS ? 11.0 + 0.5 : 6.0;   // FMA-pipe lane-cycles per oscillator-sample of this mix
  printf("pairs/thread=%d %s warps/SMSP=%d: %.2f lane-cycles per oscillator-sample (mix needs %.1f) -> %.0f %% of 128 lanes/clk\n",
         NP, WITH_COS ? "synthesis" : "phase only", warps_per_smsp, c * 128.0 / osc, need, 100.0 * need * osc / (c * 128.0));
  cudaFree(out); cudaFree(in); cudaFree(cyc);
}

int main() {
  for (int w : {4, 6, 7, 8}) { run<3, true>(w); run<3, false>(w); }
  for (int w : {7, 9}) { run<1, true>(w); run<2, true>(w); }
  return 0;
}
