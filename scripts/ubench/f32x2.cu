// Microbenchmark: throughput of packed fp32x2 arithmetic (FFMA2/FADD2/FMUL2, sm_100+) against the
// scalar forms, alone and mixed with MUFU.COS, per SM sub-partition.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32x2 f32x2.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long pk(float a, float b) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void upk(unsigned long long v, float& a, float& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}

constexpr int ITERS = 4096;
constexpr int CH = 8;   // independent chains per thread (pairs for the x2 kernels)

// MODE 0: scalar FFMA, 2*CH chains     MODE 1: FFMA2, CH pair-chains (same flops as mode 0)
// MODE 2: scalar FADD                  MODE 3: FADD2
// MODE 4: scalar FMUL+FADD alternating MODE 5: FMUL2+FADD2
// MODE 6: scalar FFMA x2CH + 2 MUFU per iteration   MODE 7: FFMA2 x CH + 2 MUFU per iteration
// MODE 8: scalar: per chain pair 12 FFMA + 2 MUFU + 2 IADD (kernel-like mix)  MODE 9: same with FFMA2
template <int MODE>
__global__ void __launch_bounds__(1024) k(float* out, float s, long long* cyc) {
  float x[2 * CH];
  unsigned long long p[CH];
#pragma unroll
  for (int i = 0; i < 2 * CH; ++i) x[i] = threadIdx.x * 1e-3f + i;
#pragma unroll
  for (int i = 0; i < CH; ++i) p[i] = pk(x[2 * i], x[2 * i + 1]);
  const unsigned long long s2 = pk(s, s), c2 = pk(1e-3f, 2e-3f);
  float m = 0.f;
  int ia = threadIdx.x;
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 2 * CH; ++i) x[i] = __fmaf_rn(x[i], s, 1e-3f * 1.0f);
    } else if (MODE == 1) {
#pragma unroll
      for (int i = 0; i < CH; ++i) p[i] = fma2(p[i], s2, c2);
    } else if (MODE == 2) {
#pragma unroll
      for (int i = 0; i < 2 * CH; ++i) x[i] = __fadd_rn(x[i], s);
    } else if (MODE == 3) {
#pragma unroll
      for (int i = 0; i < CH; ++i) p[i] = add2(p[i], s2);
    } else if (MODE == 4) {
#pragma unroll
      for (int i = 0; i < 2 * CH; ++i) x[i] = __fadd_rn(__fmul_rn(x[i], s), s);
    } else if (MODE == 5) {
#pragma unroll
      for (int i = 0; i < CH; ++i) p[i] = add2(mul2(p[i], s2), s2);
    } else if (MODE == 6) {
#pragma unroll
      for (int i = 0; i < 2 * CH; ++i) x[i] = __fmaf_rn(x[i], s, 1e-3f);
      m += __cosf(x[0]) + __cosf(x[1]);
    } else if (MODE == 7) {
#pragma unroll
      for (int i = 0; i < CH; ++i) p[i] = fma2(p[i], s2, c2);
      float a, b;
      upk(p[0], a, b);
      m += __cosf(a) + __cosf(b);
    } else if (MODE == 8) {
#pragma unroll
      for (int i = 0; i < CH; ++i) {
        float a = x[2 * i], b = x[2 * i + 1];
#pragma unroll
        for (int j = 0; j < 6; ++j) { a = __fmaf_rn(a, s, 1e-3f); b = __fmaf_rn(b, s, 2e-3f); }
        x[2 * i] = a; x[2 * i + 1] = b;
        m = __fmaf_rn(a, __cosf(a), m);
        m = __fmaf_rn(b, __cosf(b), m);
        ia = ia * 3 + 1;
      }
    } else if (MODE == 9) {
#pragma unroll
      for (int i = 0; i < CH; ++i) {
        unsigned long long v = p[i];
#pragma unroll
        for (int j = 0; j < 6; ++j) v = fma2(v, s2, c2);
        p[i] = v;
        float a, b;
        upk(v, a, b);
        m = __fmaf_rn(a, __cosf(a), m);
        m = __fmaf_rn(b, __cosf(b), m);
        ia = ia * 3 + 1;
      }
    }
  }
  const long long t1 = clock64();
  float acc = m + ia;
#pragma unroll
  for (int i = 0; i < 2 * CH; ++i) acc += x[i];
#pragma unroll
  for (int i = 0; i < CH; ++i) { float a, b; upk(p[i], a, b); acc += a + b; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int MODE>
void run(const char* name, double flops_per_iter_thread, int warps_per_smsp) {
  float* out; long long* cyc;
  const int threads = 128 * warps_per_smsp / 1;   // 4 SMSPs x warps x 32 / ... one CTA per SM
  cudaMalloc(&out, 148 * 1024 * sizeof(float));
  cudaMalloc(&cyc, 8);
  k<MODE><<<148, threads>>>(out, 0.999f, cyc);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<148, threads>>>(out, 0.999f, cyc);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  const double warp_iters = (double)ITERS * warps_per_smsp;           // per SMSP
  printf("%-44s warps/SMSP=%d cycles=%lld  cycles per (warp-iteration)=%.2f  lane-ops/clk/SM=%.1f  %.3f ms\n",
         name, warps_per_smsp, c, c / warp_iters, flops_per_iter_thread * 32 * 4 * warps_per_smsp * ITERS / (double)c, ms);
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int w : {2, 4, 6, 8}) {
    run<0>("scalar FFMA x16", 16, w);
    run<1>("FFMA2 x8 (16 lanes-ops)", 16, w);
    run<2>("scalar FADD x16", 16, w);
    run<3>("FADD2 x8", 16, w);
    run<4>("scalar FMUL+FADD x16", 32, w);
    run<5>("FMUL2+FADD2 x8", 32, w);
    run<6>("scalar FFMA x16 + 2 cos", 16, w);
    run<7>("FFMA2 x8 + 2 cos", 16, w);
    run<8>("scalar 8x(12 FFMA + 2 cos + 2 FFMA + IMAD)", 8 * 14, w);
    run<9>("packed 8x(6 FFMA2 + 2 cos + 2 FFMA + IMAD)", 8 * 14, w);
  }
  return 0;
}
