// Filtered-noise synthesiser: DynamicSizeFilteredNoise.get_signal (reference
// modules/filtered_noise_synth.py:27-42) = ddsp.core.frequency_filter(uniform noise, magnitudes).
//
// Per control frame k the reference builds a linear-phase FIR from the M magnitudes
//   c_k[i] = hann(Lir)[i] * irfft(m_k)[(i + Lir/2) mod Lir],  Lir = 2(M-1),  c_k[0] = 0
// convolves it with that frame's U noise samples (zero-padded FFTs of 256/512 points) and
// overlap-adds at hop U; the result is cropped at start = (Lir-1)/2 - 1.  In the time domain
//   z[n] = sum_{m=1}^{Lir-1} x[n-m] * c_{frame(n-m)}[m],     noise_out[t] = z[t + start].
// With Lir = 126 (190 at M = 96) taps the direct form costs fewer flops than three FFTs per
// frame, needs no transposes, and is exact, so that is what runs here, in ONE kernel:
//
//  noise_synth_kernel  CTA = (tile of 32 output frames, clip, voice slice); it loops over the voices
//                      of its slice two at a time, so most of the MultiAdd node
//                      (inharm_synth.py:296-309) is register accumulation.  Per voice pair:
//    load     the raw magnitudes of the tile's frames (+ halo) are contiguous in the caller's
//             [B, F, M] tensors: ONE bulk asynchronous copy per voice (cp.async.bulk, the 1-D form of
//             TMA) lands them in shared memory and signals an mbarrier.  The copy for the NEXT pair
//             is issued as soon as this pair's magnitudes have been consumed, so it flies under the
//             taps / noise / FIR phases below;
//    scale    FilteredNoise.get_controls, scale_fn(m + bias), both voices interleaved (m_v0, m_v1);
//    taps     c_k = Cmat x m_k: the inverse real DFT and the window folded into one [M x M-1] matrix
//             (built once in create()); the filter is symmetric about tap M-1, so only taps
//             M-1 .. 2M-3 are computed and mirrored into shared memory; thread = 4 frames x 4 taps;
//    noise    Philox4x32-10 (or the injected tensor of the parity tests), interleaved like the taps;
//    FIR      lanes are FRAMES (shared-memory pitches = 4 mod 32 words make the frame-strided 128-bit
//             accesses conflict free); a thread owns 8 consecutive outputs of its frame and slides a
//             16-tap register window over the taps.  Every multiply-accumulate is a packed FFMA2
//             (sm_100) over the two voices: half the issue slots of the scalar form, which was
//             issue-bound.
//  mix_kernel          dry = sum of the noise slices + the additive partial signals.
#pragma once
#include "common.cuh"

namespace b200ddsp {

constexpr int kNoiseFrames = 32;   // output frames per CTA (= lanes)
constexpr int kTapPad = 16;        // zero taps either side of c_k (8-aligned input blocks + the
                                   // 16-tap register window overhang by up to 14 taps)

struct NoiseVoicePtrs {
  const float* mags[B200DDSP_MAX_VOICES_INTERNAL];    // [B, F, M] raw (or already scaled) magnitudes
  const float* noise[B200DDSP_MAX_VOICES_INTERNAL];   // [B, F * U] or nullptr (Philox)
};

struct NoiseArgs {
  const float* cmat_t;   // [M][cmat_pitch] taps matrix, rows padded to a multiple of 8 with zeros
  int cmat_pitch;
  float* out;            // [n_slices, B, N] noise of each voice slice
  int v_begin, v_end;    // voices handled by this launch, split evenly over gridDim.z slices
  int slice0;            // index of the launch's first slice in `out`
  int B, F, M, U, N;     // F = input frames (rows of the magnitudes, frames of an injected noise tensor),
                         // N = output samples
  int koff;                      // input frame of output frame 0 (spans of a timeline: halo in front)
  unsigned long long sample0;    // global sample index of input frame 0 (Philox counter base)
  int halo_before, halo_after;   // input halo in FRAMES either side of the tile
  int scale_fn;          // b200ddsp_scale_fn applied to the magnitudes; 2 = already scaled
  float bias;
  int bulk;              // magnitudes rows are 16-byte aligned (M % 4 == 0): bulk asynchronous copies
  unsigned long long seed, stream_id;
};

__host__ __device__ inline int pitch_4mod32(int n) {   // smallest p >= n with p = 4 (mod 32)
  return n + ((4 - n) % 32 + 32) % 32;
}

// Shared memory of one CTA (floats).  Taps and noise of the two voices are INTERLEAVED --
// (c_v0[k], c_v1[k]), (x_v0[j], x_v1[j]) -- so that the FIR's operands are naturally aligned pairs.
struct NoiseSmemLayout {
  int n_in;       // input frames held: kNoiseFrames + halo_before + halo_after
  int pitch_x;    // float2 per row, >= U;                   2 * pitch_x = 4 (mod 32) words
  int pitch_c;    // float2 per row, >= Lir + 2*kTapPad + 2; 2 * pitch_c = 4 (mod 32) words
  int pitch_f;    // frames per band row of the scaled magnitudes: n_in rounded up to 4
  int tap_shift;  // 0..1: makes the register-window loads 16-byte aligned
  int off_raw;    // [2][n_in][M] raw magnitudes of the voice pair (bulk-copy destination)
  int off_m;      // [M][pitch_f] float2 scaled magnitudes, band-major: 4 frames of a band = 32 contiguous bytes
  int off_x;      // [n_in][pitch_x] float2 noise
  int off_c;      // [n_in][pitch_c] float2 taps
  int off_bar;    // mbarrier (8 bytes)
  int total_floats;
  __host__ __device__ NoiseSmemLayout(int M, int U, int hb, int ha) {
    const int lir = 2 * (M - 1);
    const int start = (lir - 1) / 2 - 1;
    n_in = kNoiseFrames + hb + ha;
    pitch_x = pitch_4mod32(2 * U) / 2;
    pitch_c = pitch_4mod32(2 * (lir + 2 * kTapPad + 2)) / 2;
    tap_shift = (kTapPad + start - 7) & 1;
    off_raw = 0;
    off_m = (2 * n_in * M + 31) & ~31;
    pitch_f = (n_in + 3) & ~3;
    off_x = off_m + ((2 * M * pitch_f + 31) & ~31);
    off_c = off_x + ((2 * n_in * pitch_x + 31) & ~31);
    off_bar = off_c + 2 * n_in * pitch_c;
    total_floats = off_bar + 4;
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

__device__ __forceinline__ void load2x2(float2* dst, const float2* src) {   // 16 bytes = two pairs
  const float4 v = *reinterpret_cast<const float4*>(src);
  dst[0] = make_float2(v.x, v.y);
  dst[1] = make_float2(v.z, v.w);
}

// 8 inputs x[0..7] (both voices) into 8 outputs: acc[q] += x[u] * w[7 - u + q], w = lo[0..7] | hi[0..7].
__device__ __forceinline__ void fir_step8(float2 (&acc)[8], const float2 (&lo)[8], const float2 (&hi)[8],
                                          const float2* xrow) {
#pragma unroll
  for (int h4 = 0; h4 < 2; ++h4) {                       // inputs in two halves: 8 registers instead of 16
    float2 x[4];
    load2x2(x, xrow + 4 * h4);
    load2x2(x + 2, xrow + 4 * h4 + 2);
#pragma unroll
    for (int u4 = 0; u4 < 4; ++u4) {
      const int u = 4 * h4 + u4;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int i = 7 - u + q;
        acc[q] = __ffma2_rn(x[u4], i < 8 ? lo[i] : hi[i - 8], acc[q]);
      }
    }
  }
}

// NB = 8-sample blocks per thread: blocks warp, warp + W, ... of the thread's frame.
// NB = 1 (U <= 96): at most 12 warps and 85 registers, so that two CTAs share an SM.
template <int NB>
__global__ void __launch_bounds__(NB == 1 ? 384 : 512, NB == 1 ? 2 : 1)
noise_synth_kernel(const NoiseArgs a, const NoiseVoicePtrs vp) {
  extern __shared__ __align__(16) float smem[];
  const NoiseSmemLayout L(a.M, a.U, a.halo_before, a.halo_after);
  float* raw = smem + L.off_raw;                              // [2][n_in][M]
  float2* ms = reinterpret_cast<float2*>(smem + L.off_m);     // [M][pitch_f]  scaled magnitudes of both voices
  float2* xs = reinterpret_cast<float2*>(smem + L.off_x);     // [n_in][pitch_x]  noise of both voices
  float2* cs = reinterpret_cast<float2*>(smem + L.off_c);     // [n_in][pitch_c]  zero-padded taps of both
  void* bar = smem + L.off_bar;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n_threads = blockDim.x, n_warps = blockDim.x >> 5;
  const int b = blockIdx.y;
  const int f_tile = blockIdx.x * kNoiseFrames;        // first output frame
  const int k_first = f_tile + a.koff - a.halo_before; // first input frame held (may be < 0)
  const int U = a.U, M = a.M, lir = 2 * (M - 1);
  const int start = (lir - 1) / 2 - 1;                 // crop_and_compensate_delay
  const int n_blocks = U / 8;
  const int tap0 = kTapPad + L.tap_shift;              // position of tap 0 inside a taps row
  // frames of the held range that exist: [k_lo, k_hi)
  const int k_lo = max(k_first, 0), k_hi = min(k_first + L.n_in, a.F);
  // this CTA's voices
  const int n_v = a.v_end - a.v_begin;
  const int v_lo = a.v_begin + (int)(((long long)n_v * blockIdx.z) / gridDim.z);
  const int v_hi = a.v_begin + (int)(((long long)n_v * (blockIdx.z + 1)) / gridDim.z);

  // one elected thread starts the copies of a voice pair and tells the mbarrier how many bytes to expect
  auto start_loads = [&](int v) {
    const unsigned int bytes = (unsigned int)(k_hi - k_lo) * M * 4u;
    const int nv = (v + 1 < v_hi) ? 2 : 1;
    mbar_expect_tx(bar, bytes * nv);
    for (int e = 0; e < nv; ++e)
      bulk_load(raw + (e * L.n_in + (k_lo - k_first)) * M, vp.mags[v + e] + ((size_t)b * a.F + k_lo) * M, bytes, bar);
  };
  if (a.bulk && threadIdx.x == 0) mbar_init(bar, 1);

  float2 acc[NB][8];
#pragma unroll
  for (int i = 0; i < NB; ++i)
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[i][q] = make_float2(0.f, 0.f);
  // taps outside [1, Lir-1] stay zero for the whole kernel
  for (int i = threadIdx.x; i < L.n_in * L.pitch_c; i += n_threads) cs[i] = make_float2(0.f, 0.f);
  __syncthreads();
  if (a.bulk && threadIdx.x == 0 && v_lo < v_hi && k_hi > k_lo) start_loads(v_lo);

  unsigned int parity = 0;
  for (int v = v_lo; v < v_hi; v += 2) {
    const bool have1 = v + 1 < v_hi;                   // odd voice count: the second half is silence
    // ---- noise of the pair for the held input frames (needs nothing from the magnitudes: it runs while
    //      their copy is in flight): a warp per frame, 4 samples of both voices per thread
    {
      const float* nz0 = vp.noise[v];
      const float* nz1 = have1 ? vp.noise[v + 1] : nullptr;
      if (nz0 != nullptr) nz0 += (size_t)b * a.F * U;
      if (nz1 != nullptr) nz1 += (size_t)b * a.F * U;
      const uint2 key = make_uint2((unsigned int)a.seed, (unsigned int)(a.seed >> 32));
      for (int fi = warp; fi < L.n_in; fi += n_warps) {
        const int k = k_first + fi;
        const bool in = k >= k_lo && k < k_hi;
        float2* xrow = xs + fi * L.pitch_x;
        for (int j = 4 * lane; j < U; j += 128) {
          float4 r0 = make_float4(0.f, 0.f, 0.f, 0.f), r1 = r0;
          if (in) {
            // Philox counter = (global sample index / 4, clip, voice + stream, stream >> 32), key = seed
            const unsigned int blk = (unsigned int)((a.sample0 + (size_t)k * U + j) >> 2);
            if (nz0 != nullptr) {
              r0 = __ldg(reinterpret_cast<const float4*>(nz0 + (size_t)k * U + j));
            } else {
              const uint4 bits = philox4x32_10(make_uint4(blk, (unsigned int)b, (unsigned int)v + (unsigned int)a.stream_id,
                                                          (unsigned int)(a.stream_id >> 32)), key);
              r0 = make_float4(uniform_pm1(bits.x), uniform_pm1(bits.y), uniform_pm1(bits.z), uniform_pm1(bits.w));
            }
            if (nz1 != nullptr) {
              r1 = __ldg(reinterpret_cast<const float4*>(nz1 + (size_t)k * U + j));
            } else if (have1) {
              const uint4 bits = philox4x32_10(make_uint4(blk, (unsigned int)b, (unsigned int)(v + 1) + (unsigned int)a.stream_id,
                                                          (unsigned int)(a.stream_id >> 32)), key);
              r1 = make_float4(uniform_pm1(bits.x), uniform_pm1(bits.y), uniform_pm1(bits.z), uniform_pm1(bits.w));
            }
          }
          *reinterpret_cast<float4*>(xrow + j) = make_float4(r0.x, r1.x, r0.y, r1.y);
          *reinterpret_cast<float4*>(xrow + j + 2) = make_float4(r0.z, r1.z, r0.w, r1.w);
        }
      }
    }
    // ---- magnitudes of the pair in shared memory ---------------------------------------------------
    if (a.bulk) {
      if (k_hi > k_lo) mbar_wait(bar, parity);
      parity ^= 1u;
    } else {
      for (int e = 0; e < (have1 ? 2 : 1); ++e)
        for (int i = threadIdx.x; i < (k_hi - k_lo) * M; i += n_threads)
          raw[(e * L.n_in + (k_lo - k_first)) * M + i] = __ldg(vp.mags[v + e] + ((size_t)b * a.F + k_lo) * M + i);
      __syncthreads();
    }
    // ---- scale (FilteredNoise.get_controls), interleave the two voices, transpose to band-major:
    //      consecutive lanes take consecutive frames (reads at stride M words, M odd: no bank conflict)
    const ScaleFn scale(a.scale_fn);
    for (int i = threadIdx.x; i < M * L.pitch_f; i += n_threads) {
      const int j = i / L.pitch_f, fi = i - j * L.pitch_f;
      const int k = k_first + fi;
      float2 m = make_float2(0.f, 0.f);
      if (fi < L.n_in && k >= k_lo && k < k_hi) {
        m.x = raw[fi * M + j];
        if (have1) m.y = raw[(L.n_in + fi) * M + j];
        if (a.scale_fn != 2) {
          m.x = scale(__fadd_rn(m.x, a.bias));
          if (have1) m.y = scale(__fadd_rn(m.y, a.bias));
        }
      }
      ms[i] = m;
    }
    __syncthreads();   // scaled magnitudes complete; the raw buffer is free again
    if (a.bulk && threadIdx.x == 0 && v + 2 < v_hi && k_hi > k_lo) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic reads before the async writes
      start_loads(v + 2);
    }
    // ---- taps: c[fi][d] = sum_k Cmat[k][d] * m[fi][k] for both voices.  Thread = 4 frames x 4 taps
    //      (16 FFMA2 per band); per band it reads its 4 frames' magnitude pairs (32 contiguous bytes, shared by
    //      the half warp) and 4 matrix entries (16 bytes of the padded, L1-resident table): 4 shared-memory/L1
    //      wavefronts per 32 issue cycles of FFMA2.  The first version (a thread per tap over 6 frames) needed 4
    //      wavefronts per 12 cycles on every one of the 4 schedulers -- more than the one L1 data path delivers
    //      -- and spent 40 % of the kernel here.  Operands of band k+1 are loaded under the FMAs of band k.
    {
      const int nd = M - 1, tgs = a.cmat_pitch / 4, fgs = L.n_in / 4, mstep = L.pitch_f / 2;
      const int n_tiles = tgs * fgs;
      for (int w = threadIdx.x; w < n_tiles; w += n_threads) {
        const int fg = w / tgs, tg = w - fg * tgs;
        const float4* mcol = reinterpret_cast<const float4*>(ms + 4 * fg);
        const float4* ccol = reinterpret_cast<const float4*>(a.cmat_t) + tg;
        float2 acc[4][4];
#pragma unroll
        for (int f = 0; f < 4; ++f)
#pragma unroll
          for (int t = 0; t < 4; ++t) acc[f][t] = make_float2(0.f, 0.f);
        auto band = [&](const float4& p01, const float4& p23, const float4& pc) {
          const float2 mf[4] = {make_float2(p01.x, p01.y), make_float2(p01.z, p01.w),
                                make_float2(p23.x, p23.y), make_float2(p23.z, p23.w)};
          const float ct[4] = {pc.x, pc.y, pc.z, pc.w};
#pragma unroll
          for (int f = 0; f < 4; ++f)
#pragma unroll
            for (int t = 0; t < 4; ++t) acc[f][t] = __ffma2_rn(make_float2(ct[t], ct[t]), mf[f], acc[f][t]);
        };
        float4 a01 = mcol[0], a23 = mcol[1], ac = __ldg(ccol);      // two operand sets, rotated by name
        int k = 0;
#pragma unroll 1
        for (; k + 2 <= M; k += 2) {
          const float4 b01 = mcol[mstep], b23 = mcol[mstep + 1], bc = __ldg(ccol + tgs);
          band(a01, a23, ac);
          mcol += 2 * mstep;
          ccol += 2 * tgs;
          if (k + 2 < M) {
            a01 = mcol[0];
            a23 = mcol[1];
            ac = __ldg(ccol);
          }
          band(b01, b23, bc);
        }
        if (k < M) band(a01, a23, ac);
#pragma unroll
        for (int f = 0; f < 4; ++f) {
          float2* row = cs + (4 * fg + f) * L.pitch_c + tap0;
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const int d = 4 * tg + t;
            if (d < nd) {
              row[M - 1 + d] = acc[f][t];
              if (d > 0) row[M - 1 - d] = acc[f][t];       // linear phase: symmetric about tap M-1
            }
          }
        }
      }
      // the n_in % 4 frames left over (the halo makes n_in = 34): one output per thread, on the warps after
      // the tiles' so that each scheduler gets one tile warp and one of these
      const int f_rest = 4 * fgs, n_rest = (L.n_in - f_rest) * a.cmat_pitch;
      const int t0 = ((n_tiles + 31) & ~31) % n_threads;
      for (int w = (threadIdx.x - t0 + n_threads) % n_threads; w < n_rest; w += n_threads) {
        const int fr = w / a.cmat_pitch, d = w - fr * a.cmat_pitch;
        const float2* mcol = ms + f_rest + fr;
        const float* ccol = a.cmat_t + d;
        float2 acc = make_float2(0.f, 0.f);
#pragma unroll 5
        for (int k = 0; k < M; ++k) {
          const float c = __ldg(ccol + k * a.cmat_pitch);
          acc = __ffma2_rn(make_float2(c, c), mcol[k * L.pitch_f], acc);
        }
        if (d < nd) {
          float2* row = cs + (f_rest + fr) * L.pitch_c + tap0;
          row[M - 1 + d] = acc;
          if (d > 0) row[M - 1 - d] = acc;
        }
      }
    }
    __syncthreads();   // taps and noise tile complete

    // ---- FIR.  lane = output frame, warp walks its 8-sample blocks -------------------------------
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
          const float2* xrow = xs + fi * L.pitch_x;
          const float2* crow = cs + fi * L.pitch_c + tap0;
          // 8-aligned input range (U % 8 == 0); taps outside [1, Lir-1] read the zero padding
          const int j_lo = max(0, e_lo - d * U) & ~7;
          const int j_hi = min(U - 1, e_hi - d * U) | 7;
          // tap of output q for input j+u: m0 - u + q with m0 = i0 + start - (d*U + j);
          // tap0 + m0 - 7 is even by the choice of tap_shift: 16-byte aligned pairs
          int m0 = i0 + start - (d * U + j_lo);
          // register window of 16 taps m0-7 .. m0+8 in two banks of 8: after 8 inputs the low bank
          // becomes the high one and the other bank is reloaded -- the loop is unrolled by two so that
          // the banks swap NAMES instead of contents (the moves of a shifting window ran on the FMA pipe)
          float2 wa[8], wb[8];
#pragma unroll
          for (int i = 0; i < 8; i += 2) {
            load2x2(wa + i, crow + m0 - 7 + i);
            load2x2(wb + i, crow + m0 + 1 + i);
          }
          for (int j = j_lo; j <= j_hi; j += 16) {
            fir_step8(acc[ib], wa, wb, xrow + j);
            if (j + 8 > j_hi) break;
            m0 -= 8;
#pragma unroll
            for (int i = 0; i < 8; i += 2) load2x2(wb + i, crow + m0 - 7 + i);
            fir_step8(acc[ib], wb, wa, xrow + j + 8);
            m0 -= 8;
            if (j + 16 <= j_hi) {
#pragma unroll
              for (int i = 0; i < 8; i += 2) load2x2(wa + i, crow + m0 - 7 + i);
            }
          }
        }
      }
    }
    __syncthreads();   // the FIR is done with the noise tile and the taps
  }

  // ---- through shared memory (for coalescing) to the slice's noise signal -------------------
  float* os = reinterpret_cast<float*>(xs);   // [kNoiseFrames][U + 4], reuses the noise tile
  const int pitch_o = U + 4;
#pragma unroll
  for (int ib = 0; ib < NB; ++ib) {
    const int blk = warp + ib * n_warps;
    if (blk < n_blocks) {
      float* o = os + lane * pitch_o + blk * 8;
      *reinterpret_cast<float4*>(o) = make_float4(acc[ib][0].x + acc[ib][0].y, acc[ib][1].x + acc[ib][1].y,
                                                  acc[ib][2].x + acc[ib][2].y, acc[ib][3].x + acc[ib][3].y);
      *reinterpret_cast<float4*>(o + 4) = make_float4(acc[ib][4].x + acc[ib][4].y, acc[ib][5].x + acc[ib][5].y,
                                                      acc[ib][6].x + acc[ib][6].y, acc[ib][7].x + acc[ib][7].y);
    }
  }
  __syncthreads();
  const int t_tile = f_tile * U;
  const int len = min(kNoiseFrames * U, a.N - t_tile);
  float* out = a.out + ((size_t)(a.slice0 + blockIdx.z) * a.B + b) * a.N + t_tile;
  for (int i = threadIdx.x; i < len; i += n_threads) out[i] = os[(i / U) * pitch_o + (i % U)];
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
