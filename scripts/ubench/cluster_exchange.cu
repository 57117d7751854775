// Microbenchmark for the frame loop of gru_recurrence_kernel (csrc/control_rate.cuh): what does ONE exchange of
// the hidden state between the 8 CTAs of a cluster cost per frame?
//   mode 0: distributed-shared-memory stores + cooperative_groups cluster.sync()   (what the kernel does now;
//           compiles to MEMBAR.ALL.GPU + UCGABAR_ARV/WAIT + CCTL.IVALL, profiles/r01_gru_sass_summary.txt)
//   mode 1: st.async.shared::cluster ... mbarrier::complete_tx::bytes into an mbarrier of the destination CTA;
//           each CTA arms its own mbarrier with the bytes it expects and waits on it: no cluster barrier and
//           no fence in the loop
// Both variants move the same data (every CTA writes UC values for each of RB rows into all 8 CTAs, two
// alternating buffers) and check it: each value encodes (frame, writer rank, index), the reader sums what it
// received and compares with the closed form at the end.
// NOT RUN YET (written when the round's GPU minutes were spent); builds with
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o cluster_exchange cluster_exchange.cu
// usage: ./cluster_exchange [frames]
#include <cooperative_groups.h>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

constexpr int CL = 8, UC = 24, RB = 2, U = CL * UC;      // u = 192, the row group of one clip
constexpr int THREADS = UC * RB;                         // one thread per (row, own unit)

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned map_to_rank(unsigned addr, unsigned rank) {
  unsigned r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void st_async_f32(unsigned remote_addr, float v, unsigned remote_bar) {
  asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(remote_addr),
               "r"(__float_as_uint(v)), "r"(remote_bar)
               : "memory");
}

template <int MODE>
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(THREADS)
exchange_kernel(int frames, double* sums, long long* cycles) {
  __shared__ __align__(16) float hs[2][RB][U];
  __shared__ __align__(8) unsigned long long bars[2];
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned rank = cluster.block_rank();
  const int tid = threadIdx.x, row = tid / UC, jl = tid % UC, j = rank * UC + jl;
  for (int i = tid; i < 2 * RB * U; i += THREADS) (&hs[0][0][0])[i] = 0.f;
  if (MODE == 1 && tid == 0) {
    mbar_init(smem_u32(&bars[0]), 1);
    mbar_init(smem_u32(&bars[1]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  cluster.sync();
  double acc = 0.0;
  const long long t0 = clock64();
  for (int t = 0; t < frames; ++t) {
    const int nxt = (t & 1) ^ 1;
    const float v = (float)((t % 251) * 1000 + j);                 // exact in float32
    if (MODE == 0) {
#pragma unroll
      for (unsigned c = 0; c < CL; ++c) cluster.map_shared_rank(&hs[nxt][row][j], c)[0] = v;
      cluster.sync();
    } else {
      const unsigned bar = smem_u32(&bars[nxt]);
      if (tid == 0) mbar_expect_tx(bar, CL * UC * RB * 4);         // everything this CTA receives this frame
      const unsigned dst = smem_u32(&hs[nxt][row][j]);
#pragma unroll
      for (unsigned c = 0; c < CL; ++c) st_async_f32(map_to_rank(dst, c), v, map_to_rank(bar, c));
      mbar_wait(bar, (t >> 1) & 1);                                // each barrier is used every other frame
    }
    // "use" the exchanged state: every thread reads the whole row it owns (as the dot products do)
    float s = 0.f;
#pragma unroll 8
    for (int k = 0; k < U; k += UC) s += hs[nxt][row][k + jl];
    acc += s;
  }
  const long long t1 = clock64();
  cluster.sync();                                                  // nobody leaves while peers may still write
  sums[blockIdx.x * THREADS + tid] = acc;
  if (tid == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
static void run(const char* name, int frames) {
  double* sums;
  long long* cycles;
  cudaMalloc(&sums, CL * THREADS * sizeof(double));
  cudaMalloc(&cycles, CL * sizeof(long long));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  exchange_kernel<MODE><<<CL, THREADS>>>(16, sums, cycles);        // warm-up
  cudaEventRecord(e0);
  exchange_kernel<MODE><<<CL, THREADS>>>(frames, sums, cycles);
  cudaEventRecord(e1);
  cudaError_t err = cudaDeviceSynchronize();
  if (err != cudaSuccess) {
    printf("%s: %s\n", name, cudaGetErrorString(err));
    exit(1);
  }
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  static double h_sums[CL * THREADS];
  long long h_cycles[CL];
  cudaMemcpy(h_sums, sums, sizeof(h_sums), cudaMemcpyDeviceToHost);
  cudaMemcpy(h_cycles, cycles, sizeof(h_cycles), cudaMemcpyDeviceToHost);
  // thread (row, jl) of any CTA sums, per frame, the values of units jl, jl + UC, ... = ranks 0..7
  int bad = 0;
  for (int b = 0; b < CL; ++b)
    for (int tid = 0; tid < THREADS; ++tid) {
      const int jl = tid % UC;
      double want = 0.0;
      for (int t = 0; t < frames; ++t)
        for (int c = 0; c < CL; ++c) want += (double)((t % 251) * 1000 + c * UC + jl);
      if (h_sums[b * THREADS + tid] != want) ++bad;
    }
  printf("%-44s %8.3f us per frame (%lld cycles), %s\n", name, ms * 1e3 / frames, h_cycles[0] / frames,
         bad ? "DATA MISMATCH" : "data ok");
  cudaFree(sums);
  cudaFree(cycles);
}

int main(int argc, char** argv) {
  const int frames = argc > 1 ? atoi(argv[1]) : 20000;
  run<0>("DSMEM stores + cluster.sync()", frames);
  run<1>("st.async + mbarrier complete_tx (no barrier)", frames);
  return 0;
}
