// Filtered-noise synthesiser: DynamicSizeFilteredNoise.get_signal (reference
// modules/filtered_noise_synth.py:27-42) = ddsp.core.frequency_filter(uniform noise, magnitudes).
//
// Per control frame k the reference builds a linear-phase FIR from the M magnitudes
//   c_k[i] = hann(Lir)[i] * irfft(m_k)[(i + Lir/2) mod Lir],  Lir = 2(M-1),  c_k[0] = 0
// convolves it with that frame's U noise samples (zero-padded FFTs of 256/512 points) and
// overlap-adds at hop U; the result is cropped at start = (Lir-1)/2 - 1.  In the time domain
//   z[n] = sum_{m=1}^{Lir-1} x[n-m] * c_{frame(n-m)}[m],     noise_out[t] = z[t + start].
// With Lir = 126 (190 at M = 96) taps the direct form costs fewer flops than three FFTs per
// frame, needs no transposes, and is exact, so that is what runs here:
//   stage 1  c_k = Cmat x m_k      (inverse real DFT + window as a [M-1 x M] matrix applied to
//                                   34 frames at a time; Cmat is built once in create())
//   stage 2  FIR, register tiled: a thread owns 8 consecutive outputs of one frame and slides
//            an 8-tap register window over the taps: 2 shared loads per 8 FMAs.
// CTA = (tile of 32 output frames, clip).  Lanes are FRAMES (odd shared-memory pitches make the
// frame-strided accesses conflict free), warps split the 8-sample blocks of a frame.  The CTA
// loops over the voices, accumulates their noise in a shared output tile and finally adds
// the additive partial sums: it is also the MultiAdd mixer (inharm_synth.py:296-309).
#pragma once
#include "common.cuh"

namespace b200ddsp {

constexpr int kNoiseFrames = 32;   // output frames per CTA (= lanes)
constexpr int kNoiseWarps = 4;
constexpr int kNoiseThreads = kNoiseWarps * kWarp;
constexpr int kTapPad = 16;        // zero taps either side of c_k (8-aligned input blocks + the
                                   // 15-tap register window overhang by up to 13 taps)

struct NoiseVoicePtrs {
  const float* mags[B200DDSP_MAX_VOICES_INTERNAL];    // [B, F, M] scaled magnitudes
  const float* noise[B200DDSP_MAX_VOICES_INTERNAL];   // [B, N] or nullptr (Philox)
};

struct NoiseArgs {
  const float* cmat_t;   // [M][M-1]: cmat_t[j*(M-1) + d] -> tap M-1+d (and its mirror M-1-d)
  const float* partials; // [n_partials, B, N] additive partial signals to mix in, or nullptr
  const unsigned char* live;   // [P * B, n_chunks]: partial p = (voice p / sets) is only defined
                               // where live != 0 (additive fast path), or nullptr
  float* out;            // [B, N]
  int n_partials, sets, chunk, n_chunks;
  int accumulate;        // out += result
  int P, B, F, M, U, N;
  int halo_before, halo_after;   // input halo in FRAMES either side of the tile
  unsigned long long seed, stream_id;
};

struct NoiseSmemLayout {
  int n_in;       // input frames held: kNoiseFrames + halo_before + halo_after
  int pitch_x;    // U | 1
  int pitch_c;    // (Lir + 2*kTapPad) | 1
  int pitch_m;    // frames rounded up to 4, +4 (float4 broadcast loads)
  int pitch_o;    // U | 1
  int off_x, off_c, off_m, off_out, total_floats;
  __host__ __device__ NoiseSmemLayout(int M, int U, int hb, int ha) {
    const int lir = 2 * (M - 1);
    n_in = kNoiseFrames + hb + ha;
    pitch_x = U | 1;
    pitch_c = (lir + 2 * kTapPad) | 1;
    pitch_m = ((n_in + 3) & ~3) + 4;
    pitch_o = U | 1;
    off_x = 0;
    off_c = (off_x + n_in * pitch_x + 3) & ~3;
    off_m = (off_c + n_in * pitch_c + 3) & ~3;
    off_out = (off_m + M * pitch_m + 3) & ~3;
    total_floats = off_out + kNoiseFrames * pitch_o;
  }
};

// Philox4x32-10 (Salmon et al. 2011), the generator family TF's random ops use.  The
// reference draws unseeded noise (filtered_noise_synth.py:39-40), so only the distribution
// matters: uniform on [-1, 1) with 23 random mantissa bits, like tf.random.uniform.
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const unsigned int hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
    const unsigned int hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += 0x9E3779B9u;
    key.y += 0xBB67AE85u;
  }
  return ctr;
}

__device__ __forceinline__ float uniform_pm1(unsigned int bits) {
  // [1,2) from 23 mantissa bits, minus 1 -> [0,1); * 2 - 1 -> [-1,1)
  const float u = __uint_as_float((bits >> 9) | 0x3f800000u) - 1.0f;
  return __fmaf_rn(u, 2.0f, -1.0f);
}

template <int DUMMY = 0>
__global__ void __launch_bounds__(kNoiseThreads) noise_fir_kernel(const NoiseArgs a,
                                                                 const NoiseVoicePtrs vp) {
  extern __shared__ __align__(16) float smem[];
  const NoiseSmemLayout L(a.M, a.U, a.halo_before, a.halo_after);
  float* xs = smem + L.off_x;     // [n_in][pitch_x]  noise samples by input frame
  float* cs = smem + L.off_c;     // [n_in][pitch_c]  zero-padded taps by input frame
  float* ms = smem + L.off_m;     // [M][pitch_m]     magnitudes, band-major
  float* os = smem + L.off_out;   // [kNoiseFrames][pitch_o] output tile
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.y;
  const int f_tile = blockIdx.x * kNoiseFrames;        // first output frame
  const int k_first = f_tile - a.halo_before;          // first input frame held (may be < 0)
  const int U = a.U, M = a.M, lir = 2 * (M - 1);
  const int start = (lir - 1) / 2 - 1;                 // crop_and_compensate_delay
  const int n_blocks = U / 8;

  for (int i = threadIdx.x; i < kNoiseFrames * L.pitch_o; i += kNoiseThreads) os[i] = 0.f;
  // taps outside [1, Lir-1] stay zero for the whole kernel
  for (int i = threadIdx.x; i < L.n_in * L.pitch_c; i += kNoiseThreads) cs[i] = 0.f;

  for (int v = 0; v < a.P; ++v) {
    __syncthreads();   // previous voice's FIR is done with xs/cs/ms
    // ---- stage 0: stage this voice's magnitudes and noise for the held input frames ------
    const float* mags = vp.mags[v] + (size_t)b * a.F * M;
    for (int i = threadIdx.x; i < L.n_in * M; i += kNoiseThreads) {
      const int fi = i / M, j = i - fi * M;
      const int k = k_first + fi;
      ms[j * L.pitch_m + fi] = (k >= 0 && k < a.F) ? __ldg(mags + (size_t)k * M + j) : 0.f;
    }
    const float* nz = vp.noise[v];
    if (nz != nullptr) {
      nz += (size_t)b * a.N;
      for (int i = threadIdx.x; i < L.n_in * U; i += kNoiseThreads) {
        const int fi = i / U, j = i - fi * U;
        const int k = k_first + fi;
        xs[fi * L.pitch_x + j] = (k >= 0 && k < a.F) ? __ldg(nz + (size_t)k * U + j) : 0.f;
      }
    } else {
      // counter = (sample index / 4, clip, voice, stream), key = seed
      const uint2 key = make_uint2((unsigned int)a.seed, (unsigned int)(a.seed >> 32));
      for (int i = threadIdx.x; i < L.n_in * U / 4; i += kNoiseThreads) {
        const int fi = (i * 4) / U, j = i * 4 - fi * U;
        const int k = k_first + fi;
        float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k >= 0 && k < a.F) {
          const unsigned int blk = (unsigned int)(((size_t)k * U + j) >> 2);
          const uint4 bits = philox4x32_10(
              make_uint4(blk, (unsigned int)b, (unsigned int)v + (unsigned int)a.stream_id,
                         (unsigned int)(a.stream_id >> 32)), key);
          r = make_float4(uniform_pm1(bits.x), uniform_pm1(bits.y), uniform_pm1(bits.z),
                          uniform_pm1(bits.w));
        }
        float* dst = xs + fi * L.pitch_x + j;
        dst[0] = r.x; dst[1] = r.y; dst[2] = r.z; dst[3] = r.w;
      }
    }
    __syncthreads();

    // ---- stage 1: taps c_k = Cmat x m_k for every held frame ------------------------------
    // thread = (tap d, group of 4 frames); Cmat row read coalesced through L1, magnitudes as
    // one float4 shared broadcast.
    {
      const int n_fg = (L.n_in + 3) / 4;
      const int n_d = M - 1;
      for (int w = threadIdx.x; w < n_d * n_fg; w += kNoiseThreads) {
        const int fg = w / n_d, d = w - fg * n_d;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        const float* cm = a.cmat_t + d;
        const float* mrow = ms + fg * 4;
#pragma unroll 4
        for (int j = 0; j < M; ++j) {
          const float cj = __ldg(cm + (size_t)j * n_d);
          const float4 m4 = *reinterpret_cast<const float4*>(mrow + j * L.pitch_m);
          acc.x = __fmaf_rn(cj, m4.x, acc.x);
          acc.y = __fmaf_rn(cj, m4.y, acc.y);
          acc.z = __fmaf_rn(cj, m4.z, acc.z);
          acc.w = __fmaf_rn(cj, m4.w, acc.w);
        }
        const float accs[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int fi = fg * 4 + e;
          if (fi < L.n_in) {
            float* c = cs + fi * L.pitch_c + kTapPad;
            c[M - 1 + d] = accs[e];
            if (d > 0) c[M - 1 - d] = accs[e];     // linear phase: symmetric about M-1
          }
        }
      }
    }
    __syncthreads();

    // ---- stage 2: FIR.  lane = output frame, warp walks 8-sample blocks -------------------
    {
      const int fo = lane;                               // output frame within the tile
      for (int blk = warp; blk < n_blocks; blk += kNoiseWarps) {
        const int i0 = blk * 8;
        float acc[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[q] = 0.f;
        // z index of output q: n = (f*U + i0 + q) + start; input e = i0 + q + start - m relative
        // to the start of output frame f, m in [1, Lir-1].
        const int e_hi = i0 + 7 + start - 1;             // largest input offset used
        const int e_lo = i0 + start - (lir - 1);         // smallest
        const int d_lo = (e_lo >= 0) ? e_lo / U : -((-e_lo + U - 1) / U);
        const int d_hi = e_hi / U;                       // e_hi >= 0
        for (int d = d_lo; d <= d_hi; ++d) {
          const int fi = fo + a.halo_before + d;         // held input frame index
          const float* xrow = xs + fi * L.pitch_x;
          const float* crow = cs + fi * L.pitch_c + kTapPad;
          // 8-aligned input range (U % 8 == 0); taps outside [1, Lir-1] read the zero padding
          const int j_lo = max(0, e_lo - d * U) & ~7;
          const int j_hi = min(U - 1, e_hi - d * U) | 7;
          // tap of output q for input j+u: m0 - u + q with m0 = i0 + start - (d*U + j)
          int m0 = i0 + start - (d * U + j_lo);
          float w[15];                                   // taps m0-7 .. m0+7
#pragma unroll
          for (int i = 0; i < 15; ++i) w[i] = crow[m0 - 7 + i];
          for (int j = j_lo; j <= j_hi; j += 8) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const float xv = xrow[j + u];
#pragma unroll
              for (int q = 0; q < 8; ++q) acc[q] = __fmaf_rn(xv, w[7 - u + q], acc[q]);
            }
            m0 -= 8;
            if (j + 8 <= j_hi) {
#pragma unroll
              for (int i = 0; i < 7; ++i) w[8 + i] = w[i];
#pragma unroll
              for (int i = 0; i < 8; ++i) w[i] = crow[m0 - 7 + i];
            }
          }
        }
        float* o = os + fo * L.pitch_o + i0;
#pragma unroll
        for (int q = 0; q < 8; ++q) o[q] += acc[q];
      }
    }
  }
  __syncthreads();

  // ---- mix: additive partial sums + noise tile -> out ------------------------------------
  const int t_tile = f_tile * U;
  const int len = min(kNoiseFrames * U, a.N - t_tile);
  float* out = a.out + (size_t)b * a.N + t_tile;
  for (int i = threadIdx.x; i < len; i += kNoiseThreads) {
    float acc = os[(i / U) * L.pitch_o + (i % U)];
    const int c = (t_tile + i) / a.chunk;
    for (int p = 0; p < a.n_partials; ++p) {
      if (a.live == nullptr || a.live[((size_t)(p / a.sets) * a.B + b) * a.n_chunks + c] != 0)
        acc += a.partials[((size_t)p * a.B + b) * a.N + t_tile + i];
    }
    if (a.accumulate) acc += out[i];
    out[i] = acc;
  }
}

}  // namespace b200ddsp
