// Does FFMA2/FADD2 occupy the issue port for one cycle (and the FMA pipe for two) or for two?
// Per loop iteration: 8 packed ops on 8 independent pair-chains + K integer ALU ops (LOP3/IADD3,
// other pipe).  If the packed op takes one issue slot, cycles stay ~16 until 8 + K + loop overhead > 16.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32x2_issue f32x2_issue.cu
#include <cstdio>
#include <cuda_runtime.h>
constexpr int ITERS = 4096;

template <int KIND, int K, bool XU = false>   // KIND 0: FFMA2, 1: FADD2, 2: scalar FFMA x16, 3: FFMA2 + FADD2 alternating (4+4)
__global__ void __launch_bounds__(1024) k(float* out, float s, long long* cyc, unsigned mask) {
  float m[4] = {s, s * 0.5f, s * 0.25f, s * 0.125f};
  float2 p[8];
  float x[16];
#pragma unroll
  for (int i = 0; i < 8; ++i) p[i] = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f + i);
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = threadIdx.x * 1e-3f + i;
  unsigned a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 77u + i;
  const float2 s2 = make_float2(s, s), c2 = make_float2(1e-3f, 2e-3f);
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
    if (KIND == 0) {
#pragma unroll
      for (int i = 0; i < 8; ++i) p[i] = __ffma2_rn(p[i], s2, c2);
    } else if (KIND == 1) {
#pragma unroll
      for (int i = 0; i < 8; ++i) p[i] = __fadd2_rn(p[i], c2);
    } else if (KIND == 2) {
#pragma unroll
      for (int i = 0; i < 16; ++i) x[i] = __fmaf_rn(x[i], s, 1e-3f);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) p[i] = (i & 1) ? __fadd2_rn(p[i], c2) : __ffma2_rn(p[i], s2, c2);
    }
#pragma unroll
    for (int j = 0; j < K; ++j) { if (XU) m[j & 3] = exp2f(m[j & 3]); else a[j & 7] = a[j & 7] ^ mask; }   // one LOP3 (ALU pipe) or one MUFU.EX2 (XU pipe)
  }
  const long long t1 = clock64();
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) acc += p[i].x + p[i].y + (float)a[i] + m[i & 3];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int KIND, int K, bool XU = false>
void run(const char* name) {
  float* out; long long* cyc;
  const int w = 8, threads = 128 * w;
  cudaMalloc(&out, 148 * 1024 * sizeof(float));
  cudaMalloc(&cyc, 8);
  k<KIND, K, XU><<<148, threads>>>(out, 0.999f, cyc, 0x55aa55aau);
  k<KIND, K, XU><<<148, threads>>>(out, 0.999f, cyc, 0x55aa55aau);
  cudaDeviceSynchronize();
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-28s %s K=%2d  cycles per warp-iteration per SMSP = %.2f\n", name, XU ? "MUFU" : "LOP3", K, c / ((double)ITERS * w));
  cudaFree(out); cudaFree(cyc);
}

int main() {
  run<0, 0>("FFMA2 x8"); run<0, 2>("FFMA2 x8"); run<0, 4>("FFMA2 x8"); run<0, 6>("FFMA2 x8"); run<0, 8>("FFMA2 x8");
  run<0, 1, true>("FFMA2 x8"); run<0, 2, true>("FFMA2 x8");
  run<1, 0>("FADD2 x8"); run<1, 4>("FADD2 x8"); run<1, 8>("FADD2 x8"); run<1, 2, true>("FADD2 x8");
  run<3, 0>("FFMA2 x4 + FADD2 x4"); run<3, 4>("FFMA2 x4 + FADD2 x4"); run<3, 8>("FFMA2 x4 + FADD2 x4");
  run<2, 0>("scalar FFMA x16"); run<2, 2>("scalar FFMA x16"); run<2, 4>("scalar FFMA x16"); run<2, 8>("scalar FFMA x16");
  run<2, 1, true>("scalar FFMA x16"); run<2, 2, true>("scalar FFMA x16");
  return 0;
}
