// Stream-ordered hand-off of state between consecutive spans of a timeline through peer memory
// (b200ddsp_link, include/b200ddsp.h; protocol described in timeline.cuh).
#pragma once
#include "common.cuh"

namespace b200ddsp {

struct Link {   // b200ddsp_link by value (device copy)
  const float* seed;
  unsigned long long* seed_ready;
  unsigned long long* seed_ack;
  float* carry;
  unsigned long long* carry_ready;
  unsigned long long* carry_ack;
  unsigned long long epoch;
  unsigned long long* scratch;   // [0] CTA arrival counter, [1] error word
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// payload written by a peer GPU into this GPU's memory: read around L1 (a stale line of an earlier
// epoch may sit there)
__device__ __forceinline__ float ld_inbox(const float* p) {
  float v;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}

constexpr long long kLinkTimeoutCycles = 8000000000ll;   // ~4 s at 1.9 GHz

// One thread waits until *counter >= want (nullptr: nothing to wait for); call before __syncthreads().
__device__ __forceinline__ void link_wait(const unsigned long long* counter, unsigned long long want,
                                          unsigned long long* scratch) {
  if (counter == nullptr) return;
  const long long t0 = clock64();
  unsigned int ns = 32;
  while (ld_acquire_sys(counter) < want) {
    if (clock64() - t0 > kLinkTimeoutCycles) {
      if (scratch) atomicExch(scratch + 1, 0xDEADull);
      break;
    }
    __nanosleep(ns);
    if (ns < 1024) ns *= 2;
  }
}

// One thread on the stream instead of a grid of waiting CTAs: launched in front of a kernel that reads an
// inbox, it holds the STREAM until the peer's payload has arrived and leaves the SMs to whatever runs on the
// other streams meanwhile (a grid that spins in its CTAs keeps every thread slot of the GPU occupied).
__global__ void __launch_bounds__(32) link_gate_kernel(const unsigned long long* counter, unsigned long long want,
                                                       unsigned long long* scratch) {
  if (threadIdx.x == 0) link_wait(counter, want, scratch);
}

// Called by every CTA (all threads) after its last store of a payload / last read of an inbox: the
// CTA that arrives last raises the counters.  `n_ctas` CTAs take part.
// `i_wrote`: this thread stored part of the payload (only those threads pay for the system-scope fence).
__device__ __forceinline__ void link_arrive(const Link& lk, unsigned int n_ctas, bool wrote_carry,
                                            bool read_seed, bool i_wrote = true) {
  if (lk.scratch == nullptr || !(wrote_carry || read_seed)) return;   // nothing to signal (uniform over the grid)
  if (wrote_carry && i_wrote) __threadfence_system();   // this thread's payload stores are visible system-wide
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned long long prev = atomicAdd(lk.scratch, 1ull);
    if (prev + 1 == (unsigned long long)n_ctas) {
      __threadfence_system();
      *lk.scratch = 0ull;   // next call on this stream starts from zero
      if (wrote_carry && lk.carry_ready) st_release_sys(lk.carry_ready, lk.epoch);
      if (read_seed && lk.seed_ack) st_release_sys(lk.seed_ack, lk.epoch);
    }
  }
}

}  // namespace b200ddsp
