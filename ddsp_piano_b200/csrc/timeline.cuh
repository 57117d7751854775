// Timeline reverb across GPUs (BASELINE config 4): a long dry timeline is cut into
// consecutive segments spread over ranks; the reverb of the CONCATENATED timeline is the
// 'valid'-padded convolution of every segment, overlap-added at hop N -- and the last L-1 samples
// of a rank's span spill into the head of the next rank's span.
#pragma once
#include "common.cuh"

namespace b200ddsp {

// One launch does the local overlap-add AND the exchange: output sample t of the rank's span
// gathers wet_full[i][t - i N] over the segments i that cover it (ascending i: a fixed order);
// samples past the span's end are the carry, summed the same way and added straight into the
// head of the successor's output buffer through its peer mapping (NVLink P2P, one float
// atomicAdd per sample).  The head region [0, L-1) of every buffer therefore receives exactly
// two contributions -- the local sum and the predecessor's carry -- each as ONE atomicAdd into
// zeroed memory, so the result does not depend on their arrival order (a + b == b + a).
__global__ void __launch_bounds__(256) timeline_overlap_add_kernel(
    const float* __restrict__ wet_full,   // [S, N + L - 1]
    const float* __restrict__ dry,        // [S, N] or nullptr (add_dry)
    float* __restrict__ out,              // [S * N]; out[0 .. L-1) zeroed before any rank launches
    float* __restrict__ peer_head,        // successor's out (its first L - 1 samples), or nullptr
    int S, int N, int total) {
  const long long span = (long long)S * N;
  const int tail = total - N;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= span + tail) return;
  if (t >= span && peer_head == nullptr) return;
  const int i_hi = (int)min((long long)(S - 1), t / N);
  const long long first = t - total + 1;                       // segment start must be > first - 1
  const int i_lo = first <= 0 ? 0 : (int)((first + N - 1) / N);
  float acc = 0.f;
  for (int i = i_lo; i <= i_hi; ++i) acc += wet_full[(size_t)i * total + (size_t)(t - (long long)i * N)];
  if (t < span) {
    if (dry != nullptr) acc += dry[t];
    if (t < tail) atomicAdd(out + t, acc);
    else out[t] = acc;
  } else {
    atomicAdd(peer_head + (t - span), acc);
  }
}

}  // namespace b200ddsp
