// Fast path of the additive oscillator bank (same maths and rounding as additive.cuh, which
// stays as the generic path): used when U % 8 == 0, floor(float(t) * scale) == t / U for
// every sample, and the two-word-reciprocal division is exact for the sample rate (all checked
// on the host).
//
// What it adds over the generic kernel -- none of it changes a single output bit of the phase:
//  * a warp owns one (voice, clip, 1000-sample chunk).  S even: the two substrings of a pair ride
//    on the two half-warps and chain j of a lane is partial 16 j + (lane & 15); S odd: one
//    substring on the whole warp, chain j = partial 32 j + lane.  Chains are processed two at a
//    time in packed float32x2 registers (FFMA2 / FADD2 / FMUL2), an odd last chain is packed over
//    pairs of consecutive samples;
//  * dead partials are not synthesised: get_controls zeroes every partial above Nyquist and mutes
//    voices below min_frequency, so 16-partial half-groups whose amplitude is zero in every frame
//    touching a chunk are dropped for that chunk (nh = number of live half-groups, from the
//    controls kernel / additive_alive_frames_kernel), and the phase-only pass skips a half-group
//    once no later chunk needs its offset;
//  * per control frame the warp picks the cheapest exact variant:
//      steady   both frame endpoints have identical partial frequencies (a held note):
//               omega is constant over the frame -> 1 FADD per sample in the phase chain
//      general  legacy-bilinear lerp per sample (6 FMA-pipe cycles in the chain)
//    crossed with
//      silent   every amplitude of the frame pair is zero: phase chain only
//      nocheck  no sounding partial can reach Nyquist inside the frame: the per-sample mask of
//               cos_oscillator_bank (inharm_synth.py:65-67) is elided
//      check    with the mask
//  * the unrolled body is 4 samples (not 32) so that the variants a SM executes concurrently
//    stay inside the 32 KB instruction cache (the first version of this kernel spent its top
//    stall reason on instruction fetch).
#pragma once
#include "additive.cuh"


namespace b200ddsp {

struct AdditivePlan;

struct AdditiveFastArgs {
  AdditiveArgs a;                  // a.out: [P * sets, B, N] partial signals
  AdditivePlan* plan;              // bucket counts + work counters
  const int* lists;                // [kPlanSlots][kMaxGroups][R * n_chunks] units by slot and bucket: slot 0 (phase
                                   // pass) a flat list of units; synthesis slots: rows grouped by chunk, entry
                                   // c * R + j = row of the j-th unit of chunk c
  const int* chunk_count;          // [kPlanSlots][kMaxGroups][n_chunks]      units per (slot, bucket, chunk)
  const int* chunk_first;          // [kPlanSlots][kMaxGroups][n_chunks + 1]  first work item of chunk c (prefix of
                                   // ceil(count / units per warp)), written by additive_plan_items_kernel
  const float* lerp;               // [N] legacy-bilinear lerp weight of every sample (additive_lerp_kernel)
  int slot;                        // work list to drain: 0 = phase ends (all voices),
                                   // 1 + g = synthesis of voice group g
  int sp;                          // substrings per pass (1 or 2)
};

// ---- lerp table --------------------------------------------------------------------------------
// lerp[t] = in - floor(in), in = float(t) * float(F/N): the legacy ResizeBilinear weight.  It depends
// on the sample index only, so it is computed once per call instead of once per oscillator-sample.
// Spans of a timeline: the coordinate is taken on the GLOBAL sample index tg0 + t and `scale` is that
// of the whole timeline (the reference resizes the whole piece in one call).
__global__ void __launch_bounds__(256) additive_lerp_kernel(float* __restrict__ lerp, int N, int U,
                                                            float scale, int tg0) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= N) return;
  const int tg = tg0 + t;
  const float in = __fmul_rn((float)tg, scale);
  lerp[t] = __fadd_rn(in, -(float)(tg / U));   // fast path: floor(in) == tg / U (checked on the host)
}

// ---- liveness scan ---------------------------------------------------------------------------
// na_frame[row, k] = 1 + index of the highest 16-partial half-group with a non-zero partial
// amplitude in frame k (0 = silent frame).  One warp per (row, frame).
__global__ void __launch_bounds__(256) additive_alive_frames_kernel(
    const float* __restrict__ amp, const float* __restrict__ hd, unsigned char* __restrict__ na_frame,
    int n_row_frames, int H) {
  const int rf = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (rf >= n_row_frames) return;
  const int lane = threadIdx.x & 31;
  int na = 0;
  if (__ldg(amp + rf) != 0.f) {
    for (int q = 0; q * 32 < H; ++q) {
      const int h = lane + 32 * q;
      const bool live = (h < H) && (__ldg(hd + (size_t)rf * H + h) != 0.f);
      const unsigned m = __ballot_sync(0xffffffffu, live);
      if (m) na = (m >> 16) ? 2 * q + 2 : 2 * q + 1;
    }
  }
  if (lane == 0) na_frame[rf] = (unsigned char)na;
}

// synth_na[row, c] = max na_frame over the frames chunk c reads (frames of its samples plus the
// next one: the amplitude envelope cross-fades towards frame k+1); ends_na[row, c] = max
// synth_na over chunks > c (a group's chunk-end phase matters only if a later chunk sounds).
// One CTA per row; shared memory holds the row's synth_na.
__global__ void __launch_bounds__(128) additive_alive_chunks_kernel(
    const unsigned char* __restrict__ na_frame, unsigned char* __restrict__ synth_na,
    unsigned char* __restrict__ ends_na, int F, int U, int N, int chunk, int n_chunks, int koff,
    int carry_na) {
  extern __shared__ unsigned char sna[];   // [n_chunks]
  const int row = blockIdx.x;
  const unsigned char* nf = na_frame + (size_t)row * F;
  for (int c = threadIdx.x; c < n_chunks; c += blockDim.x) {
    const int t0 = c * chunk, t1 = min(N, t0 + chunk) - 1;
    const int k0 = koff + t0 / U, k1 = min(F - 1, koff + t1 / U + 1);
    int m = 0;
    for (int k = k0; k <= k1; ++k) m = max(m, (int)nf[k]);
    sna[c] = (unsigned char)m;
    synth_na[(size_t)row * n_chunks + c] = (unsigned char)m;
  }
  __syncthreads();
  // suffix maximum: each thread owns a contiguous span, spans are stitched through shared memory
  __shared__ int span_max[128];
  const int per = (n_chunks + blockDim.x - 1) / blockDim.x;
  const int lo = threadIdx.x * per, hi = min(n_chunks, lo + per);
  int m = 0;
  for (int c = lo; c < hi; ++c) m = max(m, (int)sna[c]);
  span_max[threadIdx.x] = m;
  __syncthreads();
  int later = 0;
  for (int j = threadIdx.x + 1; j < (int)blockDim.x; ++j) later = max(later, span_max[j]);
  // carry_na > 0 (a span whose phase state is handed on): what sounds after the span is not known
  // here, so every half-group's chain is followed through every chunk
  for (int c = hi - 1; c >= lo; --c) {
    ends_na[(size_t)row * n_chunks + c] = (unsigned char)(carry_na > 0 ? carry_na : later);
    later = max(later, (int)sna[c]);
  }
}

enum { kAmpSilent = 0, kAmpNoCheck = 1, kAmpCheck = 2 };
// Synthesis units are SUB-chunks: the phase pass stores the in-chunk accumulator at every kSubLen-th
// sample, so pass 2 can enter a chunk at those points with bit-identical state.  A 1000-sample chunk
// becomes 4 units (256 + 256 + 256 + 232): four times as many, four times shorter work items -- the
// tail of the stage and the single-wave buckets shrink accordingly (a 1000-sample unit held its warp
// for ~0.2 ms whatever else the GPU had to do).
constexpr int kSubLen = 256;    // multiple of the 8-sample phase body and of kWrapEvery
constexpr int kOscUnroll = 4;   // samples per unrolled body of the synthesis pass
#ifndef B200DDSP_SYNTH_STEP
#define B200DDSP_SYNTH_STEP 4
#endif
constexpr int kSynthStep = B200DDSP_SYNTH_STEP;   // samples per loop trip of the synthesis pass (4 or 8)

// Rows and outputs of the units that share a warp in the synthesis pass (osc_chunk_h, LW < 16); row < 0: none.
struct WarpUnits {
  int row[4];
  float* out[4];
};

// Cross-lane reduction of the synthesis pass through shared memory.  Every 4-sample group leaves 4 partial
// sums per lane; summing them over the 32 lanes with a shuffle butterfly costs 6 SHFL + 6 FSEL + 10 FADD and a
// 16-byte store per group.  Instead each lane parks its 4 values in a per-warp tile [32 lanes][kRedPitch]
// (one STS.128; pitch 36 = 4 mod 32 words: conflict free) and after 8 groups lane s adds column s over the
// 32 rows (conflict free, fixed order -> bitwise reproducible) and stores one coalesced 128-byte run.
constexpr int kRedPitch = 36;
constexpr int kRedTileFloats = 32 * kRedPitch;          // per warp

__device__ __forceinline__ void flush_tile(const float* tile, int lane, int n, float* dst) {
  __syncwarp();
  if (lane < n) {
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
    for (int L = 0; L < 32; L += 4) {
      s0 += tile[(L + 0) * kRedPitch + lane];
      s1 += tile[(L + 1) * kRedPitch + lane];
      s2 += tile[(L + 2) * kRedPitch + lane];
      s3 += tile[(L + 3) * kRedPitch + lane];
    }
    dst[lane] = (s0 + s1) + (s2 + s3);
  }
  __syncwarp();
}

// The same for UPW units sharing the warp: rows [u * 32 / UPW, (u + 1) * 32 / UPW) belong to unit u.
template <int UPW>
__device__ __forceinline__ void flush_tile_units(const float* tile, int lane, int n, const WarpUnits& units,
                                                 int at) {
  constexpr int RU = 32 / UPW;                            // rows per unit (8 or 16)
  __syncwarp();
  if (lane < n) {
#pragma unroll
    for (int u = 0; u < UPW; ++u) {
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
      for (int L = 0; L < RU; L += 4) {
        s0 += tile[(u * RU + L + 0) * kRedPitch + lane];
        s1 += tile[(u * RU + L + 1) * kRedPitch + lane];
        s2 += tile[(u * RU + L + 2) * kRedPitch + lane];
        s3 += tile[(u * RU + L + 3) * kRedPitch + lane];
      }
      if (units.out[u] != nullptr) units.out[u][at + lane] = (s0 + s1) + (s2 + s3);
    }
  }
  __syncwarp();
}

// 32 lanes x 4 values -> every lane returns the sum over lanes of y[lane & 3].
__device__ __forceinline__ float transpose_reduce4(float (&y)[4], int lane) {
#pragma unroll
  for (int o = 2; o >= 1; o >>= 1) {
    const bool up = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < o; ++i) {
      const float send = up ? y[i] : y[i + o];
      const float keep = up ? y[i + o] : y[i];
      y[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
  float v = y[0];
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  v += __shfl_xor_sync(0xffffffffu, v, 8);
  v += __shfl_xor_sync(0xffffffffu, v, 16);
  return v;
}

// ---- lane layout and packed arithmetic ----------------------------------------------------------
// LW = 16 (S even): lanes 0-15 carry substring s0, lanes 16-31 substring s0 + 1; chain j of a lane
// is partial 16 j + (lane & 15).  LW = 32 (S odd): one substring, chain j = partial 32 j + lane.
// A unit with nh live 16-partial half-groups therefore runs nh chains per lane (LW = 16) with no
// idle lanes beyond the last half-group: with 32-partial groups and the substrings in the two
// halves of a register, 29 % of the lanes of the benchmark's pitch distribution computed partials
// above Nyquist; at 16-partial granularity it is 15 %.
// Chains are processed two at a time in packed float32x2 registers (sm_100: FFMA2 / FADD2 /
// FMUL2): a packed operation takes ONE issue slot for two oscillators (it still occupies the FMA
// pipe for two cycles), which frees issue slots for the MUFU, the shuffles and the address
// arithmetic; an odd last chain is packed over pairs of consecutive samples instead.  Every
// packed operation rounds each half exactly like its scalar form (.rn), so the phase stays
// bit-identical.  One trap: ptxas 12.9 contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 (it
// honours .rn only on the scalar forms), so the one product that must stay unfused --
// (bottom - top) * lerp of the legacy bilinear resize -- is computed with two scalar __fmul_rn.
template <int NC>
struct OscStateH {
  float ph[NC];    // in-chunk float32 phase accumulator
  float F[NC];     // partial frequency of frame k
  float g[NC];     // general frames: F(k+1) - F(k), rounded once like the resize kernel's
                   // bottom - top; steady frames: the constant omega
  float off[NC];   // chunk offset (synthesis pass)
  float A[NC];     // partial amplitude of frame k
  float dA[NC];    // A(k+1) - A(k): the Hann cross-fade is evaluated as A + dA * w[r]
};

__device__ __forceinline__ float2 splat2(float v) { return make_float2(v, v); }

template <int NC, int LW, bool WITH_AMP>
__device__ __forceinline__ void load_frame_h(const AdditiveArgs& a, int row, int s, int k, int l,
                                             float (&F)[NC], float (&A)[NC], bool muted = false) {
  const size_t base = ((size_t)row * a.F + k) * a.H;
  const float amp = (WITH_AMP && !muted) ? __ldg(a.amp + (size_t)row * a.F + k) : 0.f;
  const float f0 = __ldg(a.f0 + ((size_t)row * a.F + k) * a.S + s);
#pragma unroll
  for (int j = 0; j < NC; ++j) {
    const int h = l + LW * j;
    float sh = 0.f, hdv = 0.f;
    if (h < a.H) {
      sh = __ldg(a.shifts + base + h);
      if (WITH_AMP) hdv = __ldg(a.hd + base + h);
    }
    F[j] = (h < a.H) ? __fmul_rn(__fmul_rn(f0, (float)(h + 1)), __fadd_rn(1.0f, sh)) : 0.f;   // :106-108
    A[j] = __fmul_rn(amp, hdv);                                                                // :111-114
  }
}

// Load frames k and min(k + 1, F - 1) and derive the frame's variant (both frames are re-read at
// every frame boundary: one extra L1 hit per 96 samples buys 3 NC registers of carried state).
template <int NC, int LW, bool WITH_AMP>
__device__ __forceinline__ void enter_frame_h(const AdditiveArgs& a, int row, int s, int k, int l,
                                              OscStateH<NC>& st, bool& steady, int& amp_mode,
                                              bool muted = false) {
  float Fn[NC], An[NC];
  load_frame_h<NC, LW, WITH_AMP>(a, row, s, k, l, st.F, st.A, muted);
  load_frame_h<NC, LW, WITH_AMP>(a, row, s, min(k + 1, a.F - 1), l, Fn, An, muted);
  bool all_steady = true, any_live = false, any_risky = false;
  // f stays within [min(F, Fn), max(F, Fn) * (1 + 2^-22)] over the frame (one rounding in
  // bottom - top, one in the product, one in the sum), hence the margin
  const float nyq_lo = a.nyquist * (1.0f - 1e-6f);
#pragma unroll
  for (int j = 0; j < NC; ++j) {
    const bool live = WITH_AMP && (st.A[j] != 0.f || An[j] != 0.f);
    any_live |= live;
    st.dA[j] = An[j] - st.A[j];
    st.g[j] = __fadd_rn(Fn[j], -st.F[j]);
    all_steady &= (st.g[j] == 0.f);
    any_risky |= live && (fmaxf(st.F[j], Fn[j]) >= nyq_lo);
  }
  steady = __all_sync(0xffffffffu, all_steady);
  if (steady) {   // omega of a steady frame: f = F + 0 * lerp = F
#pragma unroll
    for (int j = 0; j < NC; ++j)
      st.g[j] = div_sr<true>(__fmul_rn(st.F[j], kTwoPi), a.sr, a.inv_sr, a.inv_sr_lo);
  }
  amp_mode = kAmpSilent;
  if (WITH_AMP && __any_sync(0xffffffffu, any_live))
    amp_mode = __any_sync(0xffffffffu, any_risky) ? kAmpCheck : kAmpNoCheck;
}

// The cosine needs phase + offset modulo 2 pi (the reference's floormod); the number of whole turns
// n = rint(x / 2 pi) is recomputed every kWrapEvery samples only: a sounding partial advances by
// less than pi per sample, so with the turns of up to three samples ago the argument x - n 2 pi
// (still exact: one FMA) lies in [-pi, 4 pi), where the hardware cosine is as accurate as on
// [-pi, pi] to within 1.5e-6 rad.  Saves 1.5 of 13 FMA-pipe cycles per oscillator-sample.
// The unfused product (bottom - top) * lerp as ONE packed instruction: fma(g, lerp, +0) rounds g * lerp
// exactly once like the multiply (the two differ only in the sign of a zero product, which the following
// F + m erases because partial frequencies are positive), but -- unlike mul.rn.f32x2, or fma with -0
// which ptxas first rewrites to a multiply -- ptxas may not contract it into the add that follows:
// +0 is not the additive identity of IEEE arithmetic (-0 + +0 = +0).  One issue slot instead of two.
#ifndef B200DDSP_LERP_FMA
#define B200DDSP_LERP_FMA 1
#endif
constexpr bool kLerpProductAsFma = B200DDSP_LERP_FMA != 0;
constexpr int kWrapEvery = 4;   // measured: 1 -> 2 -> 4 = 1.157 -> 1.111 -> 1.088 ms for the stage, error 4e-7 -> 9e-7

// UNROLL consecutive samples (inside one control frame) of every chain of the lane.
// win = shared Hann table positioned at the first sample's offset r inside the frame.
template <int NC, bool STEADY, int AMP, int UNROLL, bool PLAIN>
__device__ __forceinline__ void osc_group_h(const AdditiveArgs& a, OscStateH<NC>& st, const float* win,
                                            const float (&fr)[UNROLL], float (&y)[kOscUnroll]) {
  static_assert(AMP == kAmpSilent || UNROLL == 4, "window loads are float4");
  constexpr int NP = NC / 2;          // packed pairs of chains
  constexpr bool ODD = (NC & 1) != 0; // plus one scalar chain
  float wr[4] = {0.f, 0.f, 0.f, 0.f};
  if (AMP != kAmpSilent) {   // r is a multiple of 4: the load is 16-byte aligned
    const float4 r4 = *reinterpret_cast<const float4*>(win);
    wr[0] = r4.x; wr[1] = r4.y; wr[2] = r4.z; wr[3] = r4.w;
  }
  const float2 two_pi2 = splat2(kTwoPi), neg_two_pi2 = splat2(-kTwoPi);
  const float2 inv_sr2 = splat2(a.inv_sr), inv_sr_lo2 = splat2(a.inv_sr_lo);
  const float2 inv_two_pi2 = splat2(kInvTwoPi);
  const float2 magic2 = splat2(kRoundMagic), neg_magic2 = splat2(-kRoundMagic);
  float2 ph[NP > 0 ? NP : 1], acc[4];
#pragma unroll
  for (int i = 0; i < NP; ++i) ph[i] = make_float2(st.ph[2 * i], st.ph[2 * i + 1]);
#pragma unroll
  for (int i = 0; i < 4; ++i) acc[i] = make_float2(0.f, 0.f);
  float2 acct[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};   // odd chain: {y0, y1}, {y2, y3}
  float2 nwrap[NP > 0 ? NP : 1], nwrap_t = make_float2(0.f, 0.f);    // whole turns taken out of the phase
#pragma unroll
  for (int i = 0; i < (NP > 0 ? NP : 1); ++i) nwrap[i] = make_float2(0.f, 0.f);
  constexpr int L = NC - 1;   // the odd last chain
#pragma unroll
  for (int j = 0; j < UNROLL; ++j) {
    const float w0 = wr[j & 3];                // rising half of hann(2U): weight of frame k+1
#pragma unroll
    for (int i = 0; i < NP; ++i) {
      const int c0 = 2 * i, c1 = 2 * i + 1;
      float2 om, f = make_float2(0.f, 0.f);
      if (STEADY) {
        om = make_float2(st.g[c0], st.g[c1]);
      } else {
        // top + (bottom - top) * lerp, product and sum rounded separately (scalar products: see above)
        float2 m;
        if (kLerpProductAsFma) m = __ffma2_rn(make_float2(st.g[c0], st.g[c1]), splat2(fr[j]), make_float2(0.f, 0.f));
        else m = make_float2(__fmul_rn(st.g[c0], fr[j]), __fmul_rn(st.g[c1], fr[j]));
        f = __fadd2_rn(make_float2(st.F[c0], st.F[c1]), m);
        const float2 x = __fmul2_rn(f, two_pi2);                               // :69
        om = __ffma2_rn(x, inv_sr2, __fmul2_rn(x, inv_sr_lo2));                // :70, see div_sr
      }
      ph[i] = __fadd2_rn(ph[i], om);                                           // cumsum
      if (AMP != kAmpSilent) {
        // Hann cross-fade of the frame amplitudes A w[r+U] + An w[r], with w[r+U] = 1 - w[r]
        float2 amp = __ffma2_rn(make_float2(st.dA[c0], st.dA[c1]), splat2(w0),
                                make_float2(st.A[c0], st.A[c1]));
        if (AMP == kAmpCheck) {                                                // :65-67
          const float2 fc = STEADY ? make_float2(st.F[c0], st.F[c1]) : f;
          amp.x = (fc.x >= a.nyquist) ? 0.f : amp.x;
          amp.y = (fc.y >= a.nyquist) ? 0.f : amp.y;
        }
        float2 c;
        if (PLAIN) {
          c = make_float2(cos_large(ph[i].x), cos_large(ph[i].y));             // tf.cos(tf.cumsum)
        } else {
          const float2 x = __fadd2_rn(ph[i], make_float2(st.off[c0], st.off[c1]));
          if (kWrapEvery == 1 || (j & (kWrapEvery - 1)) == 0)                  // see kWrapEvery
            nwrap[i] = __fadd2_rn(__ffma2_rn(x, inv_two_pi2, magic2), neg_magic2);
          const float2 r = __ffma2_rn(nwrap[i], neg_two_pi2, x);               // wrap_to_pi
          c = make_float2(__cosf(r.x), __cosf(r.y));
        }
        acc[j & 3] = __ffma2_rn(amp, c, acc[j & 3]);                           // :80-83
      }
    }
    if (ODD && (j & 1) == 0) {
      // the odd last chain is packed over TIME instead: samples j and j + 1 share every operation
      // except the two sequential adds of the phase accumulator
      float2 om, f = make_float2(0.f, 0.f);
      if (STEADY) {
        om = splat2(st.g[L]);
      } else {
        float2 m;
        if (kLerpProductAsFma) m = __ffma2_rn(splat2(st.g[L]), make_float2(fr[j], fr[j + 1]), make_float2(0.f, 0.f));
        else m = make_float2(__fmul_rn(st.g[L], fr[j]), __fmul_rn(st.g[L], fr[j + 1]));
        f = __fadd2_rn(splat2(st.F[L]), m);
        const float2 x = __fmul2_rn(f, two_pi2);
        om = __ffma2_rn(x, inv_sr2, __fmul2_rn(x, inv_sr_lo2));
      }
      const float pa = __fadd_rn(st.ph[L], om.x);
      const float pb = __fadd_rn(pa, om.y);
      st.ph[L] = pb;
      if (AMP != kAmpSilent) {
        float2 amp = __ffma2_rn(splat2(st.dA[L]), make_float2(w0, wr[(j + 1) & 3]), splat2(st.A[L]));
        if (AMP == kAmpCheck) {
          const float2 fc = STEADY ? splat2(st.F[L]) : f;
          amp.x = (fc.x >= a.nyquist) ? 0.f : amp.x;
          amp.y = (fc.y >= a.nyquist) ? 0.f : amp.y;
        }
        float2 c;
        if (PLAIN) {
          c = make_float2(cos_large(pa), cos_large(pb));
        } else {
          const float2 x = __fadd2_rn(make_float2(pa, pb), splat2(st.off[L]));
          if (kWrapEvery == 1 || (j & (kWrapEvery - 1)) == 0)
            nwrap_t = __fadd2_rn(__ffma2_rn(x, inv_two_pi2, magic2), neg_magic2);
          const float2 r = __ffma2_rn(nwrap_t, neg_two_pi2, x);
          c = make_float2(__cosf(r.x), __cosf(r.y));
        }
        acct[(j & 3) >> 1] = __ffma2_rn(amp, c, acct[(j & 3) >> 1]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < NP; ++i) { st.ph[2 * i] = ph[i].x; st.ph[2 * i + 1] = ph[i].y; }
  if (AMP != kAmpSilent && NP > 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) y[i] += acc[i].x + acc[i].y;
  }
  if (AMP != kAmpSilent && ODD) {
    y[0] += acct[0].x; y[1] += acct[0].y; y[2] += acct[1].x; y[3] += acct[1].y;
  }
}

// One (row, substring pair, chunk) on one warp, half-warp layout.  ENDS_ONLY: phase chain only over the
// whole chunk, writes the accumulator at the sub-unit boundaries and the chunk end phases; otherwise
// sub-unit q of the chunk: writes its audio to `row_out` (the chunk's base).
// LW = 8 / LW = 4 (synthesis pass, S even, few live half-groups): TWO / FOUR units of the same chunk share
// the warp -- unit u on lanes [u * 2 LW, (u + 1) * 2 LW), its two substrings on the halves of that range,
// chain j = partial LW j + (lane & (LW - 1)).  The per-sample work that does not depend on the oscillator
// (loop, lerp and window loads, variant dispatch, reduction) is shared by 2 / 4 times as many chains, and a
// lane carries 2 NH / 4 NH chains instead of NH: buckets of 16 or 32 live partials stop being one or two
// dependent chains per lane under a fixed per-sample overhead.  `units`: row and output of every unit of
// the warp (row < 0: none).
template <int NC, int LW, bool ENDS_ONLY, bool PLAIN>
__device__ __forceinline__ void osc_chunk_h(const AdditiveArgs& a, const float* fa_lerp, int row, int s0,
                                            int c, int q, int lane, const float* win, float* row_out,
                                            float* tile = nullptr, const WarpUnits* units = nullptr) {
  const int t0 = c * a.chunk;
  const int tc1 = min(a.N, t0 + a.chunk);              // end of the chunk
  const int ts = ENDS_ONLY ? t0 : t0 + q * kSubLen;    // first / one-past-last sample of this unit
  const int t1 = (ENDS_ONLY || a.n_sub == 1) ? tc1 : min(tc1, ts + kSubLen);
  const int l = lane & (LW - 1);
  const int s = (LW < 16) ? s0 + ((lane / LW) & 1) : s0 + lane / LW;   // LW = 16: two substrings on the half-warps
  bool muted = false;                                  // lanes of a missing unit
  if constexpr (LW < 16) {
    const int r_u = units->row[lane / (2 * LW)];
    muted = r_u < 0;
    row = muted ? units->row[0] : r_u;
  }
  OscStateH<NC> st;
  int k = ts / a.U;
  int r = ts - k * a.U;
  k += a.koff;                                         // input frame (spans carry halo frames in front)
  bool steady;
  int amp_mode;
  enter_frame_h<NC, LW, !ENDS_ONLY>(a, row, s, k, l, st, steady, amp_mode, muted);
  const size_t osc_chunk = ((size_t)row * a.S + s) * a.n_chunks + c;
#pragma unroll
  for (int j = 0; j < NC; ++j) {
    st.ph[j] = 0.f;
    st.off[j] = 0.f;
    const int h = l + LW * j;
    if (!ENDS_ONLY && h < a.H) {
      if (a.fast_phase) {
        st.off[j] = a.mids[(osc_chunk * a.n_sub + q) * a.H + h];
      } else {
        if (c > 0 || a.seeded) st.off[j] = a.offsets[osc_chunk * a.H + h];
        if (q > 0) st.ph[j] = a.mids[(osc_chunk * (a.n_sub - 1) + (q - 1)) * a.H + h];
      }
    }
  }
  // samples per loop trip: 8 (chunk, frame and sub-unit lengths are multiples of 8).  The synthesis pass runs
  // its 4-sample body kSynthStep / 4 times per trip: frame check, lerp load and variant dispatch once per trip
  constexpr int STEP = ENDS_ONLY ? 8 : kSynthStep;
  for (int t = ts; t < t1; t += STEP, r += STEP) {
    if (ENDS_ONLY && a.n_sub > 1 && t > t0 && ((t - t0) & (kSubLen - 1)) == 0) {
      const int qq = (t - t0) / kSubLen - 1;           // state at the start of sub-unit qq + 1
#pragma unroll
      for (int j = 0; j < NC; ++j) {
        const int h = l + LW * j;
        if (h < a.H) a.mids[(osc_chunk * (a.n_sub - 1) + qq) * a.H + h] = st.ph[j];
      }
    }
    if (r == a.U) {
      r = 0;
      ++k;
      enter_frame_h<NC, LW, !ENDS_ONLY>(a, row, s, k, l, st, steady, amp_mode, muted);
    }
    // legacy-bilinear lerp weights of the trip's samples (table built by additive_lerp_kernel);
    // steady frames (held notes) never touch the table.  Fetching one group ahead was measured
    // and does not pay: the other resident warps already cover the L1 latency.
    float fr[STEP];
#pragma unroll
    for (int j4 = 0; j4 < STEP / 4; ++j4) {
      float4 l4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (!steady) l4 = __ldg(reinterpret_cast<const float4*>(fa_lerp + t) + j4);
      fr[4 * j4] = l4.x; fr[4 * j4 + 1] = l4.y; fr[4 * j4 + 2] = l4.z; fr[4 * j4 + 3] = l4.w;
    }
    const float* w = win + r;
    if constexpr (ENDS_ONLY) {
      float y[kOscUnroll] = {0.f, 0.f, 0.f, 0.f};
      if (steady) osc_group_h<NC, true, kAmpSilent, STEP, false>(a, st, w, fr, y);
      else osc_group_h<NC, false, kAmpSilent, STEP, false>(a, st, w, fr, y);
    } else {
      constexpr int NG = STEP / kOscUnroll;            // 4-sample groups per trip
      float y[NG][kOscUnroll];
      float f4[NG][kOscUnroll];
#pragma unroll
      for (int g = 0; g < NG; ++g)
#pragma unroll
        for (int i = 0; i < kOscUnroll; ++i) { y[g][i] = 0.f; f4[g][i] = fr[4 * g + i]; }
      if (amp_mode == kAmpSilent) {
#pragma unroll
        for (int g = 0; g < NG; ++g) {
          if (steady) osc_group_h<NC, true, kAmpSilent, kOscUnroll, false>(a, st, w + 4 * g, f4[g], y[g]);
          else osc_group_h<NC, false, kAmpSilent, kOscUnroll, false>(a, st, w + 4 * g, f4[g], y[g]);
        }
      } else if (amp_mode == kAmpNoCheck) {
#pragma unroll
        for (int g = 0; g < NG; ++g) {
          if (steady) osc_group_h<NC, true, kAmpNoCheck, kOscUnroll, PLAIN>(a, st, w + 4 * g, f4[g], y[g]);
          else osc_group_h<NC, false, kAmpNoCheck, kOscUnroll, PLAIN>(a, st, w + 4 * g, f4[g], y[g]);
        }
      } else {
#pragma unroll
        for (int g = 0; g < NG; ++g) {
          if (steady) osc_group_h<NC, true, kAmpCheck, kOscUnroll, PLAIN>(a, st, w + 4 * g, f4[g], y[g]);
          else osc_group_h<NC, false, kAmpCheck, kOscUnroll, PLAIN>(a, st, w + 4 * g, f4[g], y[g]);
        }
      }
      // park the trip's partial sums (zeros for a silent one); every 32 samples, and at the unit's end, the
      // warp adds the tile's columns and writes up to 32 samples
      const int gi = ((t - ts) >> 2) & 7;              // first 4-sample group of the trip inside its 32-sample block
#pragma unroll
      for (int g = 0; g < NG; ++g)
        *reinterpret_cast<float4*>(tile + lane * kRedPitch + 4 * (gi + g)) =
            make_float4(y[g][0], y[g][1], y[g][2], y[g][3]);
      if (gi + NG == 8 || t + STEP >= t1) {
        const int at = (t - t0) - 4 * gi;
        if constexpr (LW < 16) flush_tile_units<16 / LW>(tile, lane, 4 * (gi + NG), *units, at);
        else flush_tile(tile, lane, 4 * (gi + NG), row_out + at);
      }
    }
  }
  if (ENDS_ONLY) {
#pragma unroll
    for (int j = 0; j < NC; ++j) {
      const int h = l + LW * j;
      if (h < a.H)
        a.offsets[osc_chunk * a.H + h] = floormod_two_pi(st.ph[j]);
    }
  }
}

// nh = live 16-partial half-groups of the unit.  S even: substring pairs on the half-warps, nh chains
// per lane; S odd: one substring per pass on the whole warp, lanes own partials lane + 32 j,
// (nh + 1) / 2 chains per lane.
// CMAX = largest number of chains per lane that can occur (prunes the switch, and with it the
// register budget of the kernel, to what the configured H needs).
template <int SP, bool ENDS_ONLY, bool PLAIN, int CMAX = 8>
__device__ __forceinline__ void osc_chunk_dispatch(const AdditiveArgs& a, const float* fa_lerp, int nh,
                                                   int row, int s0, int c, int q, int lane,
                                                   const float* win, float* row_out) {
  constexpr int LW = (SP == 2) ? 16 : 32;
  const int chains = (SP == 2) ? nh : (nh + 1) / 2;
#define B200DDSP_CHAIN_CASE(N)                                                                   \
  if constexpr (CMAX >= N) {                                                                     \
    if (chains == N || (N == CMAX && chains > N)) {                                              \
      osc_chunk_h<N, LW, ENDS_ONLY, PLAIN>(a, fa_lerp, row, s0, c, q, lane, win, row_out);       \
      return;                                                                                    \
    }                                                                                            \
  }
  B200DDSP_CHAIN_CASE(1) B200DDSP_CHAIN_CASE(2) B200DDSP_CHAIN_CASE(3) B200DDSP_CHAIN_CASE(4)
  B200DDSP_CHAIN_CASE(5) B200DDSP_CHAIN_CASE(6) B200DDSP_CHAIN_CASE(7) B200DDSP_CHAIN_CASE(8)
#undef B200DDSP_CHAIN_CASE
}

// ---- fast_phase: closed-form unit start phases ---------------------------------------------------
// b200ddsp_config.fast_phase = 1 trades the bit-faithful float32 phase for speed: instead of following the
// reference's 1000 sequential float32 adds per chunk (pass 1) and its float32 sum of chunk ends (scan),
// the phase at the start of every synthesis unit is evaluated in double precision from the frame-rate
// controls.  The signal model stays the reference's -- partial frequencies move along ITS legacy
// bilinear coordinates, lerp[t] of additive_lerp_kernel -- so within frame k
//   sum_{r < n} (F_k + g_k lerp[kU + r]) = n F_k + g_k sum_{r < n} lerp[kU + r],   g_k = F_{k+1} - F_k,
// with the lerp sums taken once per call (additive_lerp_sums_kernel: one per frame, one per unit start).
// One thread per oscillator then walks the frames (a few double operations each: microseconds for the
// whole batch against 0.4 ms for pass 1).  Inside a unit (<= 256 samples) pass 2 still accumulates in
// float32 from that start.  What is removed is the rounding noise of the reference's float32 running sum
// (up to ~1e-3 rad per chunk in the highest partials, more over a clip), so the output is NOT within
// 1e-4 of the reference: opt-in, validated against the same signal model evaluated in float64
// (tests/test_gpu_parity.py::test_fast_phase_against_the_exact_model, profiles/r02_fast_phase_error_table.txt).
__global__ void __launch_bounds__(128) additive_lerp_sums_kernel(const float* __restrict__ lerp,
                                                                 double* __restrict__ frame_sum,   // [F_out]
                                                                 double* __restrict__ unit_sum,    // [n_chunks * n_sub]
                                                                 int N, int U, int chunk, int n_chunks, int n_sub) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int n_frames = N / U, n_units = n_chunks * n_sub;
  if (i < n_frames) {
    double acc = 0.0;
    for (int r = 0; r < U; ++r) acc += (double)lerp[i * U + r];
    frame_sum[i] = acc;
  } else if (i < n_frames + n_units) {
    const int u = i - n_frames, c = u / n_sub, q = u - c * n_sub;
    const long long ts = (long long)c * chunk + (long long)q * kSubLen;
    double acc = 0.0;
    if (ts < N) {
      const int k = (int)(ts / U), n = (int)(ts - (long long)k * U);
      for (int r = 0; r < n; ++r) acc += (double)lerp[k * U + r];
    }
    unit_sum[u] = acc;
  }
}

// CTA = (oscillator row rs, block of 32 partials); lane = partial, warp = one of 8 time segments of the
// clip.  Phase A: every warp sums the turns of its segment's frames; a prefix over the 8 segment totals
// gives each warp its start; phase B: the warps walk their segments again and write the unit start phases.
// Eight times the parallelism of one thread per oscillator for twice the (coalesced, 8-deep batched) loads.
constexpr int kClosedSegs = 8;
constexpr int kClosedBatch = 8;

__global__ void __launch_bounds__(kClosedSegs * 32) additive_closed_phase_kernel(
    const AdditiveArgs a, const double* __restrict__ frame_sum, const double* __restrict__ unit_sum) {
  __shared__ double seg_turns[kClosedSegs][32];
  const int hb = (a.H + 31) / 32;
  const int rs = blockIdx.x / hb, h = (blockIdx.x - rs * hb) * 32 + (threadIdx.x & 31);
  const int seg = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = rs / a.S, s = rs - row * a.S;
  const int n_frames = a.N / a.U;                       // frames synthesised
  const int per = (n_frames + kClosedSegs - 1) / kClosedSegs;
  const int k_lo = seg * per, k_hi = min(n_frames, k_lo + per);
  const double inv_sr = 1.0 / (double)a.sr;
  const float nf = (float)(h + 1);
  const bool live = h < a.H;
  const float* f0p = a.f0 + (size_t)row * a.F * a.S + s;
  const float* shp = a.shifts + (size_t)row * a.F * a.H + (live ? h : 0);
  auto freq = [&](int k) {                              // frame-rate partial frequency, float32 like the kernels
    k = min(k + a.koff, a.F - 1);
    return __fmul_rn(__fmul_rn(__ldg(f0p + (size_t)k * a.S), nf), __fadd_rn(1.0f, __ldg(shp + (size_t)k * a.H)));
  };
  float* dst = a.mids + (size_t)rs * a.n_chunks * a.n_sub * a.H + (live ? h : 0);
#pragma unroll 1
  for (int pass = 0; pass < 2; ++pass) {
    double turns = 0.0;                                 // phase / 2 pi before the first sample of frame k
    if (pass == 1) {
      for (int j = 0; j < seg; ++j) turns += seg_turns[j][lane];
      turns -= floor(turns);
    }
    // next unit start at or after the segment's first sample
    long long ts = 0;
    int c = 0, q = 0;
    if (pass == 1) {
      const long long t_lo = (long long)k_lo * a.U;
      c = (int)(t_lo / a.chunk);
      q = (int)((t_lo - (long long)c * a.chunk + kSubLen - 1) / kSubLen);
      if (q >= a.n_sub || (long long)c * a.chunk + (long long)q * kSubLen >= min((long long)a.N, (long long)(c + 1) * a.chunk)) {
        q = 0;
        ++c;
      }
      ts = (long long)c * a.chunk + (long long)q * kSubLen;
    }
    for (int k0 = k_lo; k0 < k_hi; k0 += kClosedBatch) {
      float Fb[kClosedBatch + 1];
#pragma unroll
      for (int i = 0; i <= kClosedBatch; ++i) Fb[i] = freq(min(k0 + i, n_frames + 1));
#pragma unroll
      for (int i = 0; i < kClosedBatch; ++i) {
        const int k = k0 + i;
        if (k < k_hi) {
          const double g = (double)__fadd_rn(Fb[i + 1], -Fb[i]), F = (double)Fb[i];
          if (pass == 1) {
            const long long t_end = (long long)(k + 1) * a.U;
            while (c < a.n_chunks && ts < t_end) {      // unit starts inside this frame
              const double n = (double)(ts - (long long)k * a.U);
              const double at = turns + (n * F + g * unit_sum[c * a.n_sub + q]) * inv_sr;
              if (live) dst[((size_t)c * a.n_sub + q) * a.H] = (float)((at - floor(at)) * 6.283185307179586);
              const long long chunk_end = min((long long)a.N, (long long)(c + 1) * a.chunk);
              if (++q == a.n_sub || (long long)c * a.chunk + (long long)q * kSubLen >= chunk_end) {
                q = 0;
                ++c;
              }
              ts = (long long)c * a.chunk + (long long)q * kSubLen;
            }
          }
          turns += ((double)a.U * F + g * frame_sum[k]) * inv_sr;
          turns -= floor(turns);
        }
      }
    }
    if (pass == 0) {
      seg_turns[seg][lane] = turns;
      __syncthreads();
    }
  }
}

// ---- work lists ------------------------------------------------------------------------------
// The unit of work is one (row, chunk) = 1000 samples of one voice of one clip; its cost is
// proportional to the number of live partial groups.  additive_plan_kernel buckets the units by
// that number so that the persistent kernels below can hand them out heaviest first (longest
// processing time first: the tail of the launch is made of the cheapest units).
constexpr int kMaxGroups = 8;        // buckets = live 16-partial half-groups; H <= 128 on the fast path
constexpr int kMaxVoiceGroups = 8;   // voice groups of one forward (host-input pipelining)
constexpr int kPlanSlots = 1 + kMaxVoiceGroups;

struct AdditivePlan {
  int count[kPlanSlots][kMaxGroups];   // [slot][na - 1]  number of units in the bucket
  int next[kPlanSlots];                // work counter of the persistent kernel draining the slot
  // zeroed by cudaMemsetAsync before every plan
};

struct PlanGroups {
  int n_groups;
  int first_voice[kMaxVoiceGroups + 1];
};

// Warp-aggregated list append: the lanes that append to the same counter elect a leader, which
// reserves their slots with ONE atomicAdd (18 k same-address atomics took 24 us; this takes 3).
__device__ __forceinline__ int list_append_slot(int* counter, bool active, int key) {
  const unsigned peers = __match_any_sync(0xffffffffu, active ? key : -1);
  int pos = -1;
  if (active) {
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(peers) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(counter, __popc(peers));
    base = __shfl_sync(peers, base, leader);
    pos = base + __popc(peers & ((1u << lane) - 1u));
  }
  return pos;
}

// Units per warp in the synthesis pass: buckets of few live half-groups run four (one half-group) or two
// (two or three) units of the same chunk on one warp (osc_chunk_h, LW = 4 / 8).
__host__ __device__ constexpr int synth_pack(int nh, int sp, bool plain) {
  return (sp != 2 || plain) ? 1 : nh == 1 ? 4 : nh <= 3 ? 2 : 1;
}

__global__ void __launch_bounds__(256) additive_plan_kernel(
    const unsigned char* __restrict__ synth_na, const unsigned char* __restrict__ ends_na,
    AdditivePlan* plan, int* __restrict__ lists, int* __restrict__ chunk_count, int n_units, int n_chunks,
    int B, const PlanGroups groups, int carry_all, int n_sub) {
  // lists: [kPlanSlots][kMaxGroups][n_units]; chunk_count: [kPlanSlots][kMaxGroups][n_chunks], zeroed
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool in_range = i < n_units;
  const int row = in_range ? i / n_chunks : 0, c = in_range ? i - row * n_chunks : 0;
  const int ns = in_range ? synth_na[i] : 0;
  const int R = n_units / n_chunks;
  if (ns > 0) {   // synthesis: rows grouped by chunk (threads of a warp have consecutive chunks: no contention)
    const int v = row / B;
    int g = 0;
    while (g + 1 < groups.n_groups && v >= groups.first_voice[g + 1]) ++g;
    const int key = (1 + g) * kMaxGroups + ns - 1;
    const int j = atomicAdd(chunk_count + (size_t)key * n_chunks + c, 1);
    lists[(size_t)key * n_units + (size_t)c * R + j] = row;
  }
  // pass 1 follows a half-group through chunk c if a later chunk needs its end phase (ends_na) or if
  // pass 2 enters this chunk at a sub-unit boundary (n_sub > 1: it needs the accumulator there)
  int ne = in_range ? ends_na[i] : 0;
  bool ends = ne > 0 && (c < n_chunks - 1 || carry_all);
  if (n_sub > 1 && ns > 0) {
    ne = ends ? max(ne, ns) : ns;
    ends = true;
  }
  {
    const int key = ne - 1;
    const int pos = list_append_slot(ends ? &plan->count[0][ne - 1] : nullptr, ends, key);
    if (ends) lists[(size_t)key * n_units + pos] = i;
  }
}

// Work items of the synthesis pass: chunk c of (slot, bucket) contributes ceil(count / pack) items; one
// warp per (slot, bucket) scans the chunks and leaves the first item of every chunk in chunk_first and the
// total in plan->count.  The order of the rows inside a chunk follows the atomics of the plan kernel; every
// unit's output is computed independently of its partner, so the result does not depend on it.
__global__ void __launch_bounds__(32) additive_plan_items_kernel(const int* __restrict__ chunk_count,
                                                                 int* __restrict__ chunk_first,
                                                                 AdditivePlan* plan, int n_chunks, int sp,
                                                                 int plain) {
  const int key = kMaxGroups + blockIdx.x;               // slots 1 .. kPlanSlots - 1
  const int slot = key / kMaxGroups, b = key - slot * kMaxGroups;
  const int pack = synth_pack(b + 1, sp, plain != 0);
  const int lane = threadIdx.x;
  const int* cnt = chunk_count + (size_t)key * n_chunks;
  int* first = chunk_first + (size_t)key * (n_chunks + 1);
  int base = 0;
  for (int c0 = 0; c0 < n_chunks; c0 += 32) {
    const int c = c0 + lane;
    const int items = c < n_chunks ? (cnt[c] + pack - 1) / pack : 0;
    int incl = items;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int up = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += up;
    }
    if (c < n_chunks) first[c] = base + incl - items;
    base += __shfl_sync(0xffffffffu, incl, 31);
  }
  if (lane == 0) {
    first[n_chunks] = base;
    plan->count[slot][b] = base;
  }
}

// Persistent kernel: every warp pulls (unit, substring set) items from a global counter until
// the lists are exhausted.  No block-level synchronisation after the Hann table is staged, so
// a warp that drew cheap items simply draws more of them.  Each item writes its own region of
// `out` ([P * sets, B, N] partial signals, summed in a fixed order by the mixer), which keeps the
// result independent of the scheduling order.
template <int SP, bool ENDS_ONLY, bool PLAIN = false>
__global__ void __launch_bounds__(kAddThreads, 3) additive_fast_kernel(const AdditiveFastArgs fa) {
  static_assert(ENDS_ONLY, "the persistent kernel is the phase pass; synthesis runs per bucket (additive_synth_kernel)");
  const AdditiveArgs& a = fa.a;
  const float* win = nullptr;
  const int lane = threadIdx.x & 31;
  const int kind = fa.slot;
  const int sets = a.S / SP;
  const int n_sub = ENDS_ONLY ? 1 : a.n_sub;
  const int per_unit = sets * n_sub;               // items per listed unit: (set, sub-unit)
  const int n_units = a.P * a.B * a.n_chunks;
  int bucket_end[kMaxGroups];   // cumulative item counts, heaviest bucket first
  {
    int acc = 0;
#pragma unroll
    for (int i = 0; i < kMaxGroups; ++i) {
      acc += fa.plan->count[kind][kMaxGroups - 1 - i] * per_unit;
      bucket_end[i] = acc;
    }
  }
  for (;;) {
    int item = 0;
    if (lane == 0) item = atomicAdd(&fa.plan->next[kind], 1);
    item = __shfl_sync(0xffffffffu, item, 0);
    if (item >= bucket_end[kMaxGroups - 1]) break;
    int bi = 0, begin = 0;
#pragma unroll
    for (int i = 0; i < kMaxGroups - 1; ++i)
      if (item >= bucket_end[i]) { bi = i + 1; begin = bucket_end[i]; }
    const int na = kMaxGroups - bi;   // live half-groups of the bucket
    const int local = item - begin;
    const int unit = fa.lists[(size_t)(kind * kMaxGroups + na - 1) * n_units + local / per_unit];
    const int rem = local - (local / per_unit) * per_unit;
    const int set = rem / n_sub, q = rem - set * n_sub;
    const int row = unit / a.n_chunks;
    const int c = unit - row * a.n_chunks;
    const int v = row / a.B, b = row - v * a.B;
    float* out = ENDS_ONLY ? nullptr
                           : a.out + (((size_t)v * sets + set) * a.B + b) * a.N + (size_t)c * a.chunk;
    osc_chunk_dispatch<SP, ENDS_ONLY, PLAIN>(a, fa.lerp, na, row, set * SP, c, q, lane, win, out);
  }
}

// out[b, t] (+)= sum over partial signals p of partials[p, b, t], skipping (voice, chunk) units
// that were never written because nothing sounds there (live == nullptr: all written).
struct PartialSumArgs {
  const float* partials;          // [n_partials, B, N]
  const unsigned char* live;      // [P * B, n_chunks] synth_na, or nullptr
  float* out;                     // [B, N]
  int n_partials, sets, B, N, chunk, n_chunks, accumulate;
};

__global__ void __launch_bounds__(256) additive_sum_partials_kernel(const PartialSumArgs s) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (t >= s.N) return;
  const int c = t / s.chunk;
  float acc = 0.f;
  for (int p = 0; p < s.n_partials; ++p) {
    const int v = p / s.sets;
    if (s.live == nullptr || s.live[((size_t)v * s.B + b) * s.n_chunks + c] != 0)
      acc += s.partials[((size_t)p * s.B + b) * s.N + t];
  }
  float* o = s.out + (size_t)b * s.N + t;
  *o = s.accumulate ? (*o + acc) : acc;
}

// Synthesis pass of ONE bucket (units with exactly NA live partial groups): a warp per (unit,
// substring set), 4 warps per CTA, plain grid in list order.  Compiling the buckets as separate
// kernels lets each use only the registers its NA needs (48 .. 128), i.e. 40 .. 16 resident warps
// per SM instead of 16 for all; the host launches the four buckets on four streams so that they
// fill each other's tails.  The grid is sized for the largest possible bucket; surplus CTAs exit.
#ifndef B200DDSP_SYNTH_WARPS
#define B200DDSP_SYNTH_WARPS 4
#endif
constexpr int kSynthWarps = B200DDSP_SYNTH_WARPS;
// resident CTAs per SM the compiler must allow (= register budget) by chains per lane: 9 x 128
// threads at 56 registers, 8 at 64, 7 at 72, 5 at 96.  Measured on config 3: the higher occupancy
// is worth 4 % of the stage over leaving the choice to ptxas.
__host__ __device__ constexpr int synth_min_ctas(int chains) {
  return chains <= 2 ? 9 : chains <= 4 ? 8 : chains <= 6 ? 7 : 5;
}

template <int NH, int SP, bool PLAIN>
__global__ void __launch_bounds__(kSynthWarps * 32,
                                  synth_min_ctas(SP == 2 ? NH * synth_pack(NH, SP, PLAIN) : (NH + 1) / 2) * 4 / kSynthWarps)
additive_synth_kernel(const AdditiveFastArgs fa) {
  const AdditiveArgs& a = fa.a;
  extern __shared__ __align__(16) float smem[];
  float* win = smem;                                   // [U] rising half of hann(2U)
  float* tile = smem + ((a.U + 3) & ~3) + (threadIdx.x >> 5) * kRedTileFloats;   // reduction tile of this warp
  const int lane = threadIdx.x & 31;
  constexpr int PACK = synth_pack(NH, SP, PLAIN);                // units per warp
  const int sets = a.S / SP;
  const int per_item = sets * a.n_sub;                           // warps per work item: (set, sub-unit)
  const int n_warp_items = fa.plan->count[fa.slot][NH - 1] * per_item;
  if ((int)blockIdx.x * kSynthWarps >= n_warp_items) return;     // whole CTA has nothing to do
  for (int i = threadIdx.x; i < a.U; i += blockDim.x) win[i] = a.window[i];
  __syncthreads();
  const int wi = blockIdx.x * kSynthWarps + (threadIdx.x >> 5);
  if (wi >= n_warp_items) return;
  const int item = wi / per_item, rem = wi - item * per_item;
  const int set = rem / a.n_sub, q = rem - set * a.n_sub;        // the sub-units of a chunk share a CTA
  // chunk of the item: first[c] <= item < first[c + 1]
  const int key = fa.slot * kMaxGroups + NH - 1;
  const int* first = fa.chunk_first + (size_t)key * (a.n_chunks + 1);
  int lo = 0, hi = a.n_chunks;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(first + mid) <= item) lo = mid; else hi = mid;
  }
  const int c = lo;
  const int R = a.P * a.B, n_units = R * a.n_chunks;
  const int j0 = (item - __ldg(first + c)) * PACK;
  const int* rows = fa.lists + (size_t)key * n_units + (size_t)c * R;
  const int row = rows[j0];
  auto out_of = [&](int r) {
    const int v = r / a.B, b = r - v * a.B;
    return a.out + (((size_t)v * sets + set) * a.B + b) * a.N + (size_t)c * a.chunk;
  };
  if constexpr (SP == 2 && PACK > 1) {
    const int n_here = __ldg(fa.chunk_count + (size_t)key * a.n_chunks + c);
    WarpUnits units;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      units.row[u] = (u < PACK && j0 + u < n_here) ? rows[j0 + u] : -1;
      units.out[u] = units.row[u] >= 0 ? out_of(units.row[u]) : nullptr;
    }
    osc_chunk_h<PACK * NH, 16 / PACK, false, PLAIN>(a, fa.lerp, row, set * SP, c, q, lane, win, nullptr, tile,
                                                    &units);
  } else if constexpr (SP == 2) {
    osc_chunk_h<NH, 16, false, PLAIN>(a, fa.lerp, row, set * SP, c, q, lane, win, out_of(row), tile);
  } else {
    osc_chunk_h<(NH + 1) / 2, 32, false, PLAIN>(a, fa.lerp, row, set, c, q, lane, win, out_of(row), tile);
  }
}

}  // namespace b200ddsp
