// Filtered-noise synthesiser: DynamicSizeFilteredNoise.get_signal (reference
// modules/filtered_noise_synth.py:27-42) = ddsp.core.frequency_filter(uniform noise, magnitudes).
//
// Per control frame k the reference builds a linear-phase FIR from the M magnitudes
//   c_k[i] = hann(Lir)[i] * irfft(m_k)[(i + Lir/2) mod Lir],  Lir = 2(M-1),  c_k[0] = 0
// convolves it with that frame's U noise samples (zero-padded FFTs of 256/512 points) and
// overlap-adds at hop U; the result is cropped at start = (Lir-1)/2 - 1.  In the time domain
//   z[n] = sum_{m=1}^{Lir-1} x[n-m] * c_{frame(n-m)}[m],     noise_out[t] = z[t + start].
// With Lir = 126 (190 at M = 96) taps the direct form costs fewer flops than three FFTs per
// frame, needs no transposes, and is exact, so that is what runs here, in two kernels:
//
//  noise_taps_kernel   c_k = Cmat x scale_fn(m_k + bias): the inverse real DFT and the window
//                      folded into one [M x M-1] matrix (built once in create()); the filter is
//                      symmetric about tap M-1 so only taps M-1 .. 2M-3 are computed and
//                      stored.  A register-tiled FP32 GEMM over all frames of all voices; the
//                      magnitudes' get_controls scaling is fused into the operand load.
//  noise_fir_kernel    CTA = (tile of 32 output frames, clip, voice slice); it loops over the
//                      voices of its slice, so most of the MultiAdd node (inharm_synth.py:296-309)
//                      is register accumulation.  Lanes are FRAMES (shared-memory pitches = 4 mod
//                      32 make the frame-strided 128-bit accesses conflict free); a thread owns 8
//                      consecutive outputs of its frame and slides a 15-tap register window over
//                      the taps: 4 LDS.128 per 64 FMAs.
//  mix_kernel          dry = sum of the noise slices + the additive partial signals.
#pragma once
#include "common.cuh"

namespace b200ddsp {

// ---- taps GEMM ------------------------------------------------------------------------------
constexpr int kTapsTileF = 64;   // frames per CTA
constexpr int kTapsTileD = 64;   // taps per CTA
constexpr int kTapsTileK = 32;   // bands per shared-memory stage

struct NoiseTapsPtrs {
  const float* mags[B200DDSP_MAX_VOICES_INTERNAL];   // [B, F, M] raw or scaled magnitudes per voice
};

struct NoiseTapsArgs {
  const float* cmat_t;   // [M][M-1]
  float* taps;           // [P*B*F][tap_pitch]: taps M-1 .. 2M-3 of every frame
  float* mags_out;       // [P*B*F][M] scaled magnitudes (get_controls output) or nullptr
  int frames_per_voice;  // B * F
  int M, tap_pitch;
  int scale_fn;          // b200ddsp_scale_fn to apply on load; 2 = magnitudes are already scaled
  float bias;
};

__global__ void __launch_bounds__(256) noise_taps_kernel(const NoiseTapsArgs a,
                                                         const NoiseTapsPtrs vp) {
  __shared__ float As[kTapsTileF][kTapsTileK + 1];
  __shared__ __align__(16) float Cs[kTapsTileK][kTapsTileD];
  const int v = blockIdx.z;
  const int f0 = blockIdx.x * kTapsTileF;   // frame within the voice
  const int d0 = blockIdx.y * kTapsTileD;
  const int nd = a.M - 1;
  const int tid = threadIdx.x;
  const int tf = tid >> 4, td = tid & 15;   // 16 x 16 threads, 4 x 4 outputs each
  const float* mags = vp.mags[v];
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < a.M; k0 += kTapsTileK) {
    // A tile: 64 frames x 32 bands, scaled on the way in
    for (int i = tid; i < kTapsTileF * kTapsTileK; i += 256) {
      const int fr = i / kTapsTileK, kk = i - fr * kTapsTileK;
      const int f = f0 + fr, k = k0 + kk;
      float x = 0.f;
      if (f < a.frames_per_voice && k < a.M) {
        x = __ldg(mags + (size_t)f * a.M + k);
        if (a.scale_fn != 2) x = apply_scale_fn(__fadd_rn(x, a.bias), a.scale_fn);
        if (a.mags_out != nullptr && blockIdx.y == 0)
          a.mags_out[((size_t)v * a.frames_per_voice + f) * a.M + k] = x;
      }
      As[fr][kk] = x;
    }
    for (int i = tid; i < kTapsTileK * kTapsTileD; i += 256) {
      const int kk = i / kTapsTileD, dd = i - kk * kTapsTileD;
      const int k = k0 + kk, d = d0 + dd;
      Cs[kk][dd] = (k < a.M && d < nd) ? __ldg(a.cmat_t + (size_t)k * nd + d) : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int kk = 0; kk < kTapsTileK; ++kk) {
      const float4 c4 = *reinterpret_cast<const float4*>(&Cs[kk][td * 4]);
      const float cv[4] = {c4.x, c4.y, c4.z, c4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float av = As[tf * 4 + i][kk];
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = __fmaf_rn(av, cv[j], acc[i][j]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int f = f0 + tf * 4 + i;
    if (f >= a.frames_per_voice) continue;
    float* dst = a.taps + ((size_t)v * a.frames_per_voice + f) * a.tap_pitch + d0 + td * 4;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (d0 + td * 4 + j < a.tap_pitch) dst[j] = acc[i][j];
  }
}

// ---- FIR ---------------------------------------------------------------------------------------
constexpr int kNoiseFrames = 32;   // output frames per CTA (= lanes)
constexpr int kTapPad = 16;        // zero taps either side of c_k (8-aligned input blocks + the
                                   // 16-tap register window overhang by up to 14 taps)

struct NoiseVoicePtrs {
  const float* noise[B200DDSP_MAX_VOICES_INTERNAL];   // [B, N] or nullptr (Philox)
};

struct NoiseArgs {
  const float* taps;     // [P*B*F][tap_pitch] from noise_taps_kernel
  float* out;            // [n_slices, B, N] noise of each voice slice
  int v_begin, v_end;    // voices handled by this launch, split evenly over gridDim.z slices
  int slice0;            // index of the launch's first slice in `out`
  int B, F, M, U, N, tap_pitch;  // F = input frames (rows of `taps`, frames of an injected noise tensor),
                                 // N = output samples
  int koff;                      // input frame of output frame 0 (spans of a timeline: halo in front)
  unsigned long long sample0;    // global sample index of input frame 0 (Philox counter base)
  int halo_before, halo_after;   // input halo in FRAMES either side of the tile
  unsigned long long seed, stream_id;
};

__host__ __device__ inline int pitch_4mod32(int n) {   // smallest p >= n with p = 4 (mod 32)
  return n + ((4 - n) % 32 + 32) % 32;
}

struct NoiseSmemLayout {
  int n_in;       // input frames held: kNoiseFrames + halo_before + halo_after
  int pitch_x;    // >= U,  = 4 (mod 32)
  int pitch_c;    // >= Lir + 2*kTapPad + 4,  = 4 (mod 32)
  int tap_shift;  // 0..3: makes the register-window loads 16-byte aligned
  int off_x, off_c, total_floats;
  __host__ __device__ NoiseSmemLayout(int M, int U, int hb, int ha) {
    const int lir = 2 * (M - 1);
    const int start = (lir - 1) / 2 - 1;
    n_in = kNoiseFrames + hb + ha;
    pitch_x = pitch_4mod32(U);
    pitch_c = pitch_4mod32(lir + 2 * kTapPad + 4);
    tap_shift = ((7 - start) % 4 + 4) % 4;
    off_x = 0;
    off_c = n_in * pitch_x;
    total_floats = off_c + n_in * pitch_c;
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

__device__ __forceinline__ void load4(float* dst, const float* src) {
  const float4 v = *reinterpret_cast<const float4*>(src);
  dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
}

// NB = 8-sample blocks per thread: blocks warp, warp + W, ... of the thread's frame.
template <int NB>
__global__ void __launch_bounds__(512) noise_fir_kernel(const NoiseArgs a, const NoiseVoicePtrs vp) {
  extern __shared__ __align__(16) float smem[];
  const NoiseSmemLayout L(a.M, a.U, a.halo_before, a.halo_after);
  float* xs = smem + L.off_x;     // [n_in][pitch_x]  noise samples by input frame
  float* cs = smem + L.off_c;     // [n_in][pitch_c]  zero-padded taps by input frame
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n_threads = blockDim.x, n_warps = blockDim.x >> 5;
  const int b = blockIdx.y;
  const int f_tile = blockIdx.x * kNoiseFrames;        // first output frame
  const int k_first = f_tile + a.koff - a.halo_before; // first input frame held (may be < 0)
  const int U = a.U, M = a.M, lir = 2 * (M - 1);
  const int start = (lir - 1) / 2 - 1;                 // crop_and_compensate_delay
  const int n_blocks = U / 8;
  const int tap0 = kTapPad + L.tap_shift;              // position of tap 0 inside a taps row
  // this CTA's voices
  const int n_v = a.v_end - a.v_begin;
  const int v_lo = a.v_begin + (int)(((long long)n_v * blockIdx.z) / gridDim.z);
  const int v_hi = a.v_begin + (int)(((long long)n_v * (blockIdx.z + 1)) / gridDim.z);

  float acc[NB][8];
#pragma unroll
  for (int i = 0; i < NB; ++i)
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[i][q] = 0.f;
  // taps outside [1, Lir-1] stay zero for the whole kernel
  for (int i = threadIdx.x; i < L.n_in * L.pitch_c; i += n_threads) cs[i] = 0.f;

  for (int v = v_lo; v < v_hi; ++v) {
    __syncthreads();   // previous voice's FIR is done with xs/cs (and the zero fill is visible)
    // ---- stage this voice's taps and noise for the held input frames ----------------------
    const float* taps = a.taps + ((size_t)v * a.B + b) * a.F * a.tap_pitch;
    for (int i = threadIdx.x; i < L.n_in * (M - 1); i += n_threads) {
      const int fi = i / (M - 1), d = i - fi * (M - 1);
      const int k = k_first + fi;
      const float c = (k >= 0 && k < a.F) ? __ldg(taps + (size_t)k * a.tap_pitch + d) : 0.f;
      float* row = cs + fi * L.pitch_c + tap0;
      row[M - 1 + d] = c;
      if (d > 0) row[M - 1 - d] = c;                   // linear phase: symmetric about tap M-1
    }
    const float* nz = vp.noise[v];
    if (nz != nullptr) {
      nz += (size_t)b * a.F * U;
      for (int i = threadIdx.x; i < L.n_in * U; i += n_threads) {
        const int fi = i / U, j = i - fi * U;
        const int k = k_first + fi;
        xs[fi * L.pitch_x + j] = (k >= 0 && k < a.F) ? __ldg(nz + (size_t)k * U + j) : 0.f;
      }
    } else {
      // counter = (sample index / 4, clip, voice + stream, stream >> 32), key = seed
      const uint2 key = make_uint2((unsigned int)a.seed, (unsigned int)(a.seed >> 32));
      for (int i = threadIdx.x; i < L.n_in * U / 4; i += n_threads) {
        const int fi = (i * 4) / U, j = i * 4 - fi * U;
        const int k = k_first + fi;
        float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k >= 0 && k < a.F) {
          const unsigned int blk = (unsigned int)((a.sample0 + (size_t)k * U + j) >> 2);
          const uint4 bits = philox4x32_10(
              make_uint4(blk, (unsigned int)b, (unsigned int)v + (unsigned int)a.stream_id,
                         (unsigned int)(a.stream_id >> 32)), key);
          r = make_float4(uniform_pm1(bits.x), uniform_pm1(bits.y), uniform_pm1(bits.z),
                          uniform_pm1(bits.w));
        }
        *reinterpret_cast<float4*>(xs + fi * L.pitch_x + j) = r;
      }
    }
    __syncthreads();

    // ---- FIR.  lane = output frame, warp walks its 8-sample blocks -------------------------
    const int fo = lane;                               // output frame within the tile
#pragma unroll
    for (int ib = 0; ib < NB; ++ib) {
      const int blk = warp + ib * n_warps;
      if (blk < n_blocks) {
        const int i0 = blk * 8;
        // z index of output q: n = (f*U + i0 + q) + start; input offset e = i0 + q + start - m
        // relative to the start of output frame f, m in [1, Lir-1].
        const int e_hi = i0 + 7 + start - 1;             // largest input offset used
        const int e_lo = i0 + start - (lir - 1);         // smallest
        const int d_lo = (e_lo >= 0) ? e_lo / U : -((-e_lo + U - 1) / U);
        const int d_hi = e_hi / U;                       // e_hi >= 0
        for (int d = d_lo; d <= d_hi; ++d) {
          const int fi = fo + a.halo_before + d;         // held input frame index
          const float* xrow = xs + fi * L.pitch_x;
          const float* crow = cs + fi * L.pitch_c + tap0;
          // 8-aligned input range (U % 8 == 0); taps outside [1, Lir-1] read the zero padding
          const int j_lo = max(0, e_lo - d * U) & ~7;
          const int j_hi = min(U - 1, e_hi - d * U) | 7;
          // tap of output q for input j+u: m0 - u + q with m0 = i0 + start - (d*U + j);
          // m0 = start (mod 8), so crow + m0 - 7 is 16-byte aligned by the choice of tap_shift
          int m0 = i0 + start - (d * U + j_lo);
          float w[16];                                   // taps m0-7 .. m0+8 (the last is unused)
#pragma unroll
          for (int i = 0; i < 16; i += 4) load4(w + i, crow + m0 - 7 + i);
          for (int j = j_lo; j <= j_hi; j += 8) {
            float x[8];
            load4(x, xrow + j);
            load4(x + 4, xrow + j + 4);
#pragma unroll
            for (int u = 0; u < 8; ++u)
#pragma unroll
              for (int q = 0; q < 8; ++q) acc[ib][q] = __fmaf_rn(x[u], w[7 - u + q], acc[ib][q]);
            m0 -= 8;
            if (j + 8 <= j_hi) {
#pragma unroll
              for (int i = 0; i < 7; ++i) w[8 + i] = w[i];
              load4(w, crow + m0 - 7);
              load4(w + 4, crow + m0 - 3);
            }
          }
        }
      }
    }
  }
  __syncthreads();

  // ---- through shared memory (for coalescing) to the slice's noise signal -------------------
  float* os = xs;   // [kNoiseFrames][pitch_x], reuses the noise tile
#pragma unroll
  for (int ib = 0; ib < NB; ++ib) {
    const int blk = warp + ib * n_warps;
    if (blk < n_blocks) {
      float* o = os + lane * L.pitch_x + blk * 8;
      *reinterpret_cast<float4*>(o) = make_float4(acc[ib][0], acc[ib][1], acc[ib][2], acc[ib][3]);
      *reinterpret_cast<float4*>(o + 4) = make_float4(acc[ib][4], acc[ib][5], acc[ib][6], acc[ib][7]);
    }
  }
  __syncthreads();
  const int t_tile = f_tile * U;
  const int len = min(kNoiseFrames * U, a.N - t_tile);
  float* out = a.out + ((size_t)(a.slice0 + blockIdx.z) * a.B + b) * a.N + t_tile;
  for (int i = threadIdx.x; i < len; i += n_threads) out[i] = os[(i / U) * L.pitch_x + (i % U)];
}

// ---- mix ---------------------------------------------------------------------------------------
// out[b, t] (+)= sum of the noise slices + sum of the additive partial signals (skipping (voice,
// chunk) units that were never written because nothing sounds there).  Fixed summation order.
struct MixArgs {
  const float* noise;          // [n_noise, B, N] or nullptr
  const float* partials;       // [n_partials, B, N] or nullptr
  const unsigned char* live;   // [P * B, n_chunks] synth_na (partial p belongs to voice p / sets), or nullptr
  float* out;                  // [B, N]
  int n_noise, n_partials, sets, B, N, chunk, n_chunks, accumulate;
};

__global__ void __launch_bounds__(256) mix_kernel(const MixArgs m) {
  const int t = (blockIdx.x * blockDim.x + threadIdx.x) * 4;   // N and chunk are multiples of 4
  const int b = blockIdx.y;
  if (t >= m.N) return;
  const int c = t / m.chunk;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int z = 0; z < m.n_noise; ++z) {
    const float4 v = *reinterpret_cast<const float4*>(m.noise + ((size_t)z * m.B + b) * m.N + t);
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  for (int p = 0; p < m.n_partials; ++p) {
    if (m.live == nullptr || m.live[((size_t)(p / m.sets) * m.B + b) * m.n_chunks + c] != 0) {
      const float4 v = *reinterpret_cast<const float4*>(m.partials + ((size_t)p * m.B + b) * m.N + t);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  }
  float4* o = reinterpret_cast<float4*>(m.out + (size_t)b * m.N + t);
  if (m.accumulate) {
    const float4 v = *o;
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  *o = acc;
}

}  // namespace b200ddsp
