// Control-rate (250 Hz) helpers that feed the synthesis path.
#pragma once
#include <cooperative_groups.h>

#include "common.cuh"

namespace b200ddsp {

// NoteRelease (reference modules/sub_modules.py:1174-1188): an RNN over the frames of one voice
// whose cell, F0ProcessorCell.call (:1138-1171), holds the last played note for `release_frames`
// frames after its note-off.  State = (note in memory, frames since note-off), both float32 and
// updated with the reference's own arithmetic (products of saturated ramps, not branches):
//   activity    = min(relu(pitch), 1)
//   release_end = min(relu(steps - release_frames), 1)
//   out         = activity * pitch + (1 - activity) * previous * (1 - release_end)
//   steps       = (steps + 1) * (1 - activity) * (1 - release_end)
// One thread per voice row; the recurrence is sequential in time and a row is 750 frames.
__global__ void __launch_bounds__(128) note_release_kernel(const float* __restrict__ pitch,
                                                           float* __restrict__ out, int rows, int F,
                                                           int in_stride, float release_frames) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= rows) return;
  const float* p = pitch + (size_t)row * F * in_stride;
  float* o = out + (size_t)row * F;
  float previous = 0.f, steps = 0.f;
  for (int k = 0; k < F; ++k) {
    const float note = p[(size_t)k * in_stride];
    const float activity = fminf(fmaxf(note, 0.f), 1.f);
    const float release_end = fminf(fmaxf(__fadd_rn(steps, -release_frames), 0.f), 1.f);
    const float keep = __fmul_rn(__fmul_rn(__fadd_rn(1.f, -activity), previous),
                                 __fadd_rn(1.f, -release_end));
    const float y = __fadd_rn(__fmul_rn(activity, note), keep);
    steps = __fmul_rn(__fmul_rn(__fadd_rn(steps, 1.f), __fadd_rn(1.f, -activity)),
                      __fadd_rn(1.f, -release_end));
    previous = y;
    o[k] = y;
  }
}

// ---------------------------------------------------------------------------------------------
// GRU recurrence (tf.keras.layers.GRU(units, return_sequences=True), TF2 default reset_after=True:
// the GRUs of ContextNetwork / MonophonicNetwork, reference configs/dafx22.gin:64-72, 77-86, and of
// FiLMContextNetwork / MonophonicDeepNetwork, modules/sub_modules.py:121, 503).
// ---------------------------------------------------------------------------------------------
// The input projections x W_i + b_i of all frames are one library GEMM upstream; what is left is the
// part that is sequential in time:
//   r = sigmoid(xr + h Wr + br)   z = sigmoid(xz + h Wz + bz)   n = tanh(xn + r * (h Wn + bn))
//   h' = (1 - z) * n + z * h
// A frame is a [rows, u] x [u, 3u] product with rows = voices x clips (16 for one clip): far too small
// for a launch per frame, which is what a library GRU costs (6.5 us per frame measured, 100 ms for a
// 60 s piece).  Here ONE launch walks all frames: a cluster of CL CTAs owns a group of RB rows for the
// whole sequence, CTA c keeps the recurrent weights of its UC = u / CL units in shared memory
// (3 * UC * u floats: 55 KB for u = 192 on 8 CTAs) and the hidden state of the group in two shared
// buffers; after each frame every CTA writes its UC new state values into the next buffer of ALL CTAs
// of the cluster (distributed shared memory) and one cluster barrier ends the frame.
// Thread = (row pair p, unit j, quarter q of the k range): 4 adjacent lanes split the dot products
// (float4 reads, k interleaved by 16 so that a quarter-warp reads 2 x 64 contiguous bytes on disjoint
// banks: row stride u + 16 floats), reduce with two shuffles, lanes q = 0, 1 finish rows 2p, 2p + 1.
// ASYNC (CL > 1): the new state travels with st.async.shared::cluster ... mbarrier::complete_tx::bytes into an
// mbarrier of the DESTINATION CTA, which every CTA arms with the bytes it expects per frame and waits on:
// no cluster barrier and no GPU-scope fence in the frame loop (cluster.sync() compiles to MEMBAR.ALL.GPU +
// UCGABAR arrive/wait).  Measured in isolation (scripts/ubench/cluster_exchange.cu,
// profiles/r02_ubench_cluster_exchange.txt): 793 -> 410 cycles per frame.  Why no barrier is needed: a CTA can
// only write frame t + 1's state into a peer's buffer after it has received ALL of frame t's state, which the
// peer sends after it has finished reading that very buffer.
__device__ __forceinline__ unsigned int map_shared_u32(unsigned int addr, unsigned int rank) {
  unsigned int r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_async_f32(unsigned int remote_addr, float v, unsigned int remote_bar) {
  asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(remote_addr),
               "r"(__float_as_uint(v)), "r"(remote_bar)
               : "memory");
}

template <int UC, int CL, bool ASYNC = false>
__global__ void __launch_bounds__(UC * 32 > 1024 ? 1024 : UC * 32, 1)
gru_recurrence_kernel(const float* __restrict__ x_proj,   // [rows, F, 3u]  gates (r, z, n), bias b_i included
                      const float* __restrict__ w_hh,     // [3u, u]        torch weight_hh layout
                      const float* __restrict__ b_hh,     // [3u]
                      float* __restrict__ out,            // [rows, F, u]
                      int rows, int F, int RB) {
  constexpr int u = UC * CL;
  constexpr int WS = u + 16;                              // padded row of the weight tile
  extern __shared__ __align__(16) float gru_smem[];
  float* Ws = gru_smem;                                   // [3][UC][WS]
  float* hs = gru_smem + 3 * UC * WS;                     // [2][RB][u]
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(hs + 2 * RB * u);   // [2] (ASYNC)
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (CL > 1) ? (int)cluster.block_rank() : 0;
  const int group = blockIdx.x / CL;
  const int tid = threadIdx.x;
  const int q = tid & 3, jl = (tid >> 2) % UC, p = (tid >> 2) / UC;
  const int j = rank * UC + jl;                           // unit of this thread
  const int row0 = group * RB;

  for (int i = tid; i < 3 * UC * u; i += blockDim.x) {    // recurrent weights of the own units
    const int g = i / (UC * u), r = i % (UC * u), jj = r / u, k = r % u;
    Ws[(g * UC + jj) * WS + k] = w_hh[(size_t)(g * u + rank * UC + jj) * u + k];
  }
  for (int i = tid; i < 2 * RB * u; i += blockDim.x) hs[i] = 0.f;

  // lanes q = 0, 1 own the state of (row 2p + q, unit j)
  const int lrow = 2 * p + q;                             // meaningful for q < 2
  const int row = row0 + lrow;
  const bool owner = q < 2 && row < rows;
  float b_r = 0.f, b_z = 0.f, b_n = 0.f, h_own = 0.f;
  if (owner) { b_r = b_hh[j]; b_z = b_hh[u + j]; b_n = b_hh[2 * u + j]; }
  const float* xp = x_proj + ((size_t)row * F) * (3 * u) + j;
  float* op = out + ((size_t)row * F) * u + j;

  // lane q writes the new state into the buffers of CTAs q, q + 4, ... of the cluster
  constexpr int NR = (CL + 3) / 4;
  float* remote[NR];
#pragma unroll
  for (int i = 0; i < NR; ++i) {
    const int c = q + 4 * i;
    remote[i] = c >= CL ? nullptr : (CL > 1 ? cluster.map_shared_rank(hs, c) : hs);
  }
  if (ASYNC && tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
  }
  if (CL > 1) cluster.sync(); else __syncthreads();       // weights, zero state and barriers in place everywhere
  const unsigned int hs_u32 = smem_u32(hs), bars_u32 = smem_u32(bars);

  const float4* wr = reinterpret_cast<const float4*>(Ws + (0 * UC + jl) * WS) + q;
  const float4* wz = reinterpret_cast<const float4*>(Ws + (1 * UC + jl) * WS) + q;
  const float4* wn = reinterpret_cast<const float4*>(Ws + (2 * UC + jl) * WS) + q;
  const unsigned base_lane = (tid & 31) & ~3u;

  float xr_n = 0.f, xz_n = 0.f, xn_n = 0.f;               // input projections, fetched one frame ahead
  if (owner) { xr_n = __ldg(xp); xz_n = __ldg(xp + u); xn_n = __ldg(xp + 2 * u); }
  for (int t = 0; t < F; ++t) {
    const int cur = t & 1;
    const float xr = xr_n, xz = xz_n, xn = xn_n;
    if (owner && t + 1 < F) {                             // in flight under this frame's dot products and exchange
      xr_n = __ldg(xp + (size_t)(t + 1) * 3 * u);
      xz_n = __ldg(xp + (size_t)(t + 1) * 3 * u + u);
      xn_n = __ldg(xp + (size_t)(t + 1) * 3 * u + 2 * u);
    }
    if (ASYNC && tid == 0) mbar_expect_tx(&bars[cur ^ 1], (unsigned int)(RB * u * sizeof(float)));
    const float4* ha = reinterpret_cast<const float4*>(hs + (cur * RB + 2 * p) * u) + q;
    const float4* hb = ha + u / 4;
    float ar0 = 0.f, az0 = 0.f, an0 = 0.f, ar1 = 0.f, az1 = 0.f, an1 = 0.f;
#pragma unroll
    for (int i = 0; i < u / 16; ++i) {
      const float4 a = ha[4 * i], b = hb[4 * i];
      const float4 r4 = wr[4 * i], z4 = wz[4 * i], n4 = wn[4 * i];
      ar0 = fmaf(r4.x, a.x, ar0); az0 = fmaf(z4.x, a.x, az0); an0 = fmaf(n4.x, a.x, an0);
      ar1 = fmaf(r4.x, b.x, ar1); az1 = fmaf(z4.x, b.x, az1); an1 = fmaf(n4.x, b.x, an1);
      ar0 = fmaf(r4.y, a.y, ar0); az0 = fmaf(z4.y, a.y, az0); an0 = fmaf(n4.y, a.y, an0);
      ar1 = fmaf(r4.y, b.y, ar1); az1 = fmaf(z4.y, b.y, az1); an1 = fmaf(n4.y, b.y, an1);
      ar0 = fmaf(r4.z, a.z, ar0); az0 = fmaf(z4.z, a.z, az0); an0 = fmaf(n4.z, a.z, an0);
      ar1 = fmaf(r4.z, b.z, ar1); az1 = fmaf(z4.z, b.z, az1); an1 = fmaf(n4.z, b.z, an1);
      ar0 = fmaf(r4.w, a.w, ar0); az0 = fmaf(z4.w, a.w, az0); an0 = fmaf(n4.w, a.w, an0);
      ar1 = fmaf(r4.w, b.w, ar1); az1 = fmaf(z4.w, b.w, az1); an1 = fmaf(n4.w, b.w, an1);
    }
#pragma unroll
    for (int o = 1; o <= 2; o <<= 1) {
      ar0 += __shfl_xor_sync(0xffffffffu, ar0, o); az0 += __shfl_xor_sync(0xffffffffu, az0, o);
      an0 += __shfl_xor_sync(0xffffffffu, an0, o); ar1 += __shfl_xor_sync(0xffffffffu, ar1, o);
      az1 += __shfl_xor_sync(0xffffffffu, az1, o); an1 += __shfl_xor_sync(0xffffffffu, an1, o);
    }
    const float sr = q == 0 ? ar0 : ar1, sz = q == 0 ? az0 : az1, sn = q == 0 ? an0 : an1;
    // precise expf / tanhf / IEEE division on purpose: with the hardware approximations (__expf, __fdividef:
    // 3e-7 per gate) the frame loses ~150 cycles but the controls of a 750-frame clip drift past 1e-4 of the
    // restatement (the recurrence amplifies per-frame rounding; measured, round 2)
    const float r = 1.f / (1.f + expf(-(xr + sr + b_r)));
    const float z = 1.f / (1.f + expf(-(xz + sz + b_z)));
    const float n = tanhf(xn + r * (sn + b_n));
    const float h_new = owner ? (1.f - z) * n + z * h_own : 0.f;
    h_own = h_new;
    if (owner) op[(size_t)t * u] = h_new;
    // both rows of the pair to all four lanes, then out to the cluster
    const float va = __shfl_sync(0xffffffffu, h_new, base_lane), vb = __shfl_sync(0xffffffffu, h_new, base_lane + 1);
    const int at = ((cur ^ 1) * RB + 2 * p) * u + j;
    if (ASYNC) {
#pragma unroll
      for (int i = 0; i < NR; ++i) {
        const unsigned int c = q + 4 * i;
        if (c < CL) {
          const unsigned int bar = map_shared_u32(bars_u32 + 8u * (cur ^ 1), c);
          st_async_f32(map_shared_u32(hs_u32 + 4u * at, c), va, bar);
          st_async_f32(map_shared_u32(hs_u32 + 4u * (at + u), c), vb, bar);
        }
      }
      mbar_wait(&bars[cur ^ 1], (unsigned int)((t >> 1) & 1));   // each barrier is used every other frame
    } else {
#pragma unroll
      for (int i = 0; i < NR; ++i)
        if (remote[i] != nullptr) { remote[i][at] = va; remote[i][at + u] = vb; }
      if (CL > 1) cluster.sync(); else __syncthreads();
    }
  }
  if (ASYNC) cluster.sync();                              // nobody leaves while peers may still write
}

}  // namespace b200ddsp
