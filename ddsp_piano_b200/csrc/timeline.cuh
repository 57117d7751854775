// One timeline cut into spans over GPUs (BASELINE config 4): the stream-ordered hand-off
// of state between consecutive spans through peer memory, and the reverb of a span.
//
// The reference synthesises a piece in ONE pass (synthesize_midi_file.py:52-54,73), so two things cross
// every span boundary: the oscillators' phase state (additive.cuh: additive_offsets_kernel) and the
// L - 1 samples of reverb tail.  Both use the same protocol (b200ddsp_link, include/b200ddsp.h): the
// producing kernel stores the payload straight into the successor's inbox over NVLink P2P, fences at
// system scope, and the last of its CTAs raises a counter in the successor's memory; the consuming
// kernel polls that counter (a local L2 read), then reads the inbox.  Inboxes are double buffered and
// acknowledged, so a rank may run at most two calls ahead of its successor.  Every wait is bounded
// (~4 s): a lost peer turns into an error word, never into a hung GPU.
#pragma once
#include "common.cuh"
#include "link.cuh"
#include "reverb.cuh"

namespace b200ddsp {

// ---- reverb of a span ---------------------------------------------------------------------------
// Rows are the span's segments ([B * n_seg, N], timeline-major); every segment of timeline b is
// convolved with impulse response b.  scales[row] = (2^-ea(row), 2^-ei(b), 2^(ea + ei)), see
// reverb.cuh (per-row power-of-two normalisation).
__global__ void timeline_scales_kernel(const unsigned int* __restrict__ max_audio,   // [rows][2], .x used
                                       const unsigned int* __restrict__ max_ir,      // [B][2], .y used
                                       float4* __restrict__ scales, float4* __restrict__ ir_scales,
                                       int rows, int n_seg) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const int b = r / n_seg;
  const float ma = max_audio ? __uint_as_float(max_audio[2 * r]) : 0.f;
  const float mi = __uint_as_float(max_ir[2 * b + 1]);
  int ea = 0, ei = 0;
  if (ma > 0.f && ma < 3.0e38f) frexpf(ma, &ea);
  if (mi > 0.f && mi < 3.0e38f) frexpf(mi, &ei);
  ea = max(-60, min(60, ea));
  ei = max(-60, min(60, ei));
  scales[r] = make_float4(ldexpf(1.f, -ea), ldexpf(1.f, -ei), ldexpf(1.f, ea + ei), 0.f);
  if (ir_scales != nullptr && r == b * n_seg) ir_scales[b] = make_float4(1.f, ldexpf(1.f, -ei), 1.f, 0.f);
}

// z = ir[b] * 2^-ei (real), tap 0 masked (Reverb._mask_dry_ir), zero padded
struct LoadRealSingle {
  const float* x; const float4* scales; int len, first;
  __device__ __forceinline__ float2 operator()(int b, int i) const {
    if (i < first || i >= len) return make_float2(0.f, 0.f);
    return make_float2(__ldg(x + (size_t)b * len + i) * __ldg(scales + b).y, 0.f);
  }
};

// Za = FFT(a0 + i a1) of a pair of segment rows, ZH[b] = FFT(ir_b) (Hermitian: the IR is real) ->
// V[k] = conj Y0[k] + i conj Y1[k], Y = A * H of the row's timeline.
__global__ void __launch_bounds__(256) timeline_spectrum_kernel(const float2* __restrict__ Za,
                                                                const float2* __restrict__ ZH,
                                                                float2* __restrict__ V, int n, int rows,
                                                                int n_seg) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;   // 0 .. n/2
  const int pair = blockIdx.y;
  if (k > n / 2) return;
  const int kn = (n - k) & (n - 1);
  const int r0 = 2 * pair, r1 = min(r0 + 1, rows - 1);
  const float2 za = Za[(size_t)pair * n + k], wa = Za[(size_t)pair * n + kn];
  const float2 A0 = make_float2(0.5f * (za.x + wa.x), 0.5f * (za.y - wa.y));
  const float2 A1 = make_float2(0.5f * (za.y + wa.y), -0.5f * (za.x - wa.x));
  const float2 H0 = ZH[(size_t)(r0 / n_seg) * n + k];
  const float2 H1 = ZH[(size_t)(r1 / n_seg) * n + k];
  const float2 y0 = cmul(A0, H0), y1 = cmul(A1, H1);
  V[(size_t)pair * n + k] = make_float2(y0.x + y1.y, y1.x - y0.y);
  if (kn != k) V[(size_t)pair * n + kn] = make_float2(y0.x - y1.y, y0.y + y1.x);
}

// Overlap-add of the span's 'valid'-padded segment convolutions + the two ends of the tail hand-off,
// ONE launch.  Output sample t of timeline b gathers wet_full[b, i][t - i N] over the segments i that
// cover it, in ascending i (a fixed order), plus the dry sample, plus -- for t < L - 1 -- the tail the
// predecessor handed over.  Samples past the span's end are this span's tail: summed the same way and
// stored into the successor's inbox.
// Grid order matters for liveness: the CTAs that PRODUCE the tail come first (they never wait for a
// predecessor), the CTAs that CONSUME one come last, so a waiting CTA can never keep a producing CTA
// of the same launch from being scheduled.
struct TimelineTailArgs {
  const float* wet_full;   // [B, n_seg, N + L - 1]
  const float* dry;        // [B, n_seg * N] or nullptr (add_dry)
  float* out;              // [B, n_seg * N]
  int B, n_seg, N, total;  // total = N + L - 1
  int tail_ctas, body_ctas, head_ctas;   // per timeline, 256 samples each
  Link link;
};

__global__ void __launch_bounds__(256) timeline_tail_kernel(const TimelineTailArgs a) {
  const int b = blockIdx.y;
  const int tail = a.total - a.N;                    // L - 1
  const long long span = (long long)a.n_seg * a.N;
  const Link& lk = a.link;
  const bool produce = lk.carry != nullptr;
  const bool consume = lk.seed != nullptr;
  const unsigned int n_ctas = gridDim.x * gridDim.y;
  int cta = blockIdx.x;
  long long t;        // sample index relative to the span start; >= span: this span's tail
  int region;         // 0 tail (produce), 1 body, 2 head (consume)
  if (cta < a.tail_ctas) {
    region = 0;
    t = span + (long long)cta * 256 + threadIdx.x;
  } else if (cta < a.tail_ctas + a.body_ctas) {
    region = 1;
    t = tail + (long long)(cta - a.tail_ctas) * 256 + threadIdx.x;
  } else {
    region = 2;
    t = (long long)(cta - a.tail_ctas - a.body_ctas) * 256 + threadIdx.x;
  }
  const bool in_range = (region == 0) ? (t < span + tail) : (region == 1) ? (t < span) : (t < tail && t < span);
  float acc = 0.f;
  if (in_range && (region != 0 || produce)) {
    const int i_hi = (int)min((long long)(a.n_seg - 1), t / a.N);
    const long long first = t - a.total + 1;
    const int i_lo = first <= 0 ? 0 : (int)((first + a.N - 1) / a.N);
    const float* w = a.wet_full + (size_t)b * a.n_seg * a.total;
    for (int i = i_lo; i <= i_hi; ++i) acc += w[(size_t)i * a.total + (size_t)(t - (long long)i * a.N)];
  }
  if (region == 0) {
    if (produce) {
      if (threadIdx.x == 0 && lk.carry_ack != nullptr && lk.epoch > 2)
        link_wait(lk.carry_ack, lk.epoch - 2, lk.scratch);   // the slot's previous payload was consumed
      __syncthreads();
      if (in_range) lk.carry[(size_t)b * tail + (size_t)(t - span)] = acc;
    }
  } else {
    if (region == 2 && consume) {
      if (threadIdx.x == 0) link_wait(lk.seed_ready, lk.epoch, lk.scratch);
      __syncthreads();
      if (in_range) acc += ld_inbox(lk.seed + (size_t)b * tail + (size_t)t);
    }
    if (in_range) {
      if (a.dry != nullptr) acc += a.dry[(size_t)b * span + t];
      a.out[(size_t)b * span + t] = acc;
    }
  }
  link_arrive(lk, n_ctas, produce, consume, region == 0);
}

}  // namespace b200ddsp
