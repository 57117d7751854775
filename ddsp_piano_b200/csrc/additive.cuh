// The inharmonic additive oscillator bank: MultiInharmonic.get_signal
// (reference modules/inharm_synth.py:272-293 -> harmonic_synthesis :87-127 ->
// cos_oscillator_bank :49-84, with ddsp.core.resample / angular_cumsum underneath).
//
// Data flow per (clip b, voice v, substring s, partial h), sample t = k*U + r:
//   F_k[h]  = (f0[k,s] * n) * (1 + shift_k[h])                 frame-rate partial frequency
//   A_k[h]  = amp_k * hd_k[h]                                   frame-rate partial amplitude
//   f[t]    = F_lo + (F_hi - F_lo) * (in - lo),  in = float(t) * float(F/N)   (legacy bilinear)
//   a[t]    = A_k * w[r+U] + A_{k+1} * w[r]                     (Hann overlap-add upsampling)
//   omega   = (f * 2pi) / sr ;  phase = cumsum within the 1000-sample chunk (float32,
//             sequential) ; p = floormod(phase + chunk_offset, 2pi) ; y[t] += a * cos(p)
//
// The float32 phase path is reproduced operation for operation (unfused mul/add, IEEE
// division, sequential adds) because rounding noise of the reference's cumsum is part of
// its output; everything that is not accumulated (amplitudes, cos, the sums over
// partials/voices) only has to be accurate.
//
// Three launches: (1) `ends` pass = the phase chain alone, one value per (oscillator,
// chunk); (2) a tiny scan turning chunk ends into chunk offsets; (3) the synthesis pass.
// Work decomposition of (1) and (3): CTA = (chunk, clip, voice group); each warp takes
// (voice, substring) pairs round-robin; lane l owns partials l, l+32, ... (HP per lane),
// so per-sample work that does not depend on the partial (lerp weight, window weights) is
// shared by HP oscillators and the per-frame control loads are coalesced.
#pragma once
#include "common.cuh"
#include "link.cuh"

namespace b200ddsp {

constexpr int kAddWarps = 8;
constexpr int kAddThreads = kAddWarps * kWarp;
constexpr int kMaxChunk = 1024;   // smem row length; ddsp's chunk_size is 1000

struct AdditiveArgs {
  const float* amp;     // [R, F]     R = P*B rows, row = v*B + b (voice-major, sub_modules.py:589-596)
  const float* hd;      // [R, F, H]
  const float* shifts;  // [R, F, H]
  const float* f0;      // [R, F, S]
  float* offsets;       // [R*S, n_chunks, H]: chunk end phases (pass 1), then chunk offsets (scan)
  float* mids;          // [R*S, n_chunks, n_sub - 1, H]: in-chunk phase accumulator at samples kSubLen, 2 kSubLen,
                        // ... of every chunk (fast path, pass 1), so that pass 2 can start inside a chunk;
                        // fast_phase: [R*S, n_chunks, n_sub, H] phase (mod 2 pi) at the start of every unit
  int n_sub;            // synthesis units per chunk (fast path): ceil(chunk / kSubLen), or 1
  int fast_phase;       // unit start phases come from the closed form (additive_closed_phase_kernel)
  const float* decays;      // [R, F, H] or nullptr: SurrogateAdditive (surrogate_synth.py:78-97), generic
  const float* decay_time;  // [R, F]                kernel only: amplitude *= |decay|^(decay_time U + r)
  float* out;           // [G, B, N]
  const float* window;  // [2U]  tf.signal.hann_window(2U, periodic)
  int B, P, F, H, S, U, N;
  int chunk, n_chunks;
  int voices_per_group;
  int koff;             // input frame of output sample 0 (spans: halo frames in front; else 0)
  int seeded;           // chunk 0 has an offset too (a span that continues a timeline)
  int accumulate;       // out += (only honoured when gridDim.z == 1)
  int plain;            // generic kernel, inference = 0: one plain cumsum over the clip (chunk = N), no wrap
  float scale;          // float32(F) / float32(N)
  float nyquist;        // float32(sr / 2)
  float sr;             // float32(sr)
  float inv_sr;         // float32(1 / sr)
  float inv_sr_lo;      // float32(1 / sr - inv_sr): second word of the reciprocal
};

// x / sr, correctly rounded.  FASTDIV: the reciprocal is held as two float32 words r + r_lo
// (relative error 2^-48), so fma(x, r, fl(x * r_lo)) is the quotient to 2^-47 before its single
// rounding -- and a float32 quotient by an integer sample rate can only come that close to a
// rounding boundary when it lies on it.  b200ddsp_create() verifies on the host, over all 2^23
// mantissas, that this equals IEEE division for the configured sample rate before selecting
// this path.  Two FMA-pipe operations instead of the classic three (x*r, remainder, correction).
template <bool FASTDIV>
__device__ __forceinline__ float div_sr(float x, float sr, float inv_sr, float inv_sr_lo) {
  if (FASTDIV) return __fmaf_rn(x, inv_sr, __fmul_rn(x, inv_sr_lo));
  return __fdiv_rn(x, sr);
}

__device__ __forceinline__ float transpose_reduce32(float (&y)[32], int lane) {
  // 32 lanes x 32 values -> lane l returns sum over lanes of y[l].  31 shuffles.
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    const bool up = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < o; ++i) {
      const float send = up ? y[i] : y[i + o];
      const float keep = up ? y[i + o] : y[i];
      y[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
  return y[0];
}

template <int HP>
struct FrameRegs {
  float F[HP];  // partial frequencies of a control frame
  float A[HP];  // partial amplitudes of a control frame
};

template <int HP, bool WITH_AMP>
__device__ __forceinline__ void load_frame(const AdditiveArgs& a, int row, int s, int k, int lane,
                                           FrameRegs<HP>& fr) {
  const float f0 = __ldg(a.f0 + ((size_t)row * a.F + k) * a.S + s);
  const float amp = WITH_AMP ? __ldg(a.amp + (size_t)row * a.F + k) : 0.f;
  const size_t base = ((size_t)row * a.F + k) * a.H;
#pragma unroll
  for (int q = 0; q < HP; ++q) {
    const int h = lane + 32 * q;
    fr.F[q] = 0.f;
    fr.A[q] = 0.f;
    if (h < a.H) {
      const float n = (float)(h + 1);
      const float sh = __ldg(a.shifts + base + h);
      // get_harmonic_frequencies (f0 * n), then *= (1.0 + shifts)   inharm_synth.py:106-108
      fr.F[q] = __fmul_rn(__fmul_rn(f0, n), __fadd_rn(1.0f, sh));
      if (WITH_AMP) fr.A[q] = __fmul_rn(amp, __ldg(a.hd + base + h));   // :111-114
    }
  }
}

// G samples per group: every group lies inside one control frame and one chunk
// (G = 8 needs U % 8 == 0 and chunk % 8 == 0; G = 1 is the generic path).
// FAST: floor(float(t)*scale) == t / U for every t (checked on the host) and the 3-op
// division is exact; otherwise the lerp frame is decided per sample and __fdiv_rn is used.
template <int HP, int G, bool FAST, bool ENDS_ONLY>
__global__ void __launch_bounds__(kAddThreads) additive_kernel(const AdditiveArgs a) {
  extern __shared__ float smem[];
  float* win = smem;                                   // [2U]
  float* rows = smem + ((2 * a.U + 31) & ~31);         // [kAddWarps][kMaxChunk]
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int n_threads = blockDim.x, n_warps = blockDim.x >> 5;   // <= kAddThreads / kAddWarps
  const int c = blockIdx.x, b = blockIdx.y, g = blockIdx.z;
  const int t0 = c * a.chunk;
  const int t1 = min(a.N, t0 + a.chunk);

  if (!ENDS_ONLY) {
    for (int i = threadIdx.x; i < 2 * a.U; i += n_threads) win[i] = a.window[i];
    for (int i = threadIdx.x; i < n_warps * kMaxChunk; i += n_threads) rows[i] = 0.f;
    __syncthreads();
  }

  const int v_begin = g * a.voices_per_group;
  const int v_end = min(a.P, v_begin + a.voices_per_group);
  const int n_pairs = (v_end - v_begin) * a.S;
  const float two_pi = kTwoPi;

  for (int pair = warp; pair < n_pairs; pair += n_warps) {
    const int v = v_begin + pair / a.S;
    const int s = pair - (pair / a.S) * a.S;
    const int row = v * a.B + b;

    int k = t0 / a.U;
    int r = t0 - k * a.U;
    FrameRegs<HP> cur, nxt;
    float dF[HP], ph[HP], off[HP];
    float Fprev[HP], dFprev[HP];   // generic path only: the frame below (lo == k-1)
    float D[HP], T0 = 0.f;         // surrogate: |decay| of frame k per partial, decay_time * U
    const bool surrogate = !ENDS_ONLY && a.decays != nullptr;
    auto load_decays = [&](int kk) {
      T0 = __fmul_rn(__ldg(a.decay_time + (size_t)row * a.F + kk), (float)a.U);
#pragma unroll
      for (int q = 0; q < HP; ++q) {
        const int h = lane + 32 * q;
        D[q] = (h < a.H) ? fabsf(__ldg(a.decays + ((size_t)row * a.F + kk) * a.H + h)) : 1.f;
      }
    };
    if (surrogate) load_decays(k);
    load_frame<HP, !ENDS_ONLY>(a, row, s, k, lane, cur);
    load_frame<HP, !ENDS_ONLY>(a, row, s, min(k + 1, a.F - 1), lane, nxt);
    const size_t off_base = (((size_t)row * a.S + s) * a.n_chunks + c) * a.H;
#pragma unroll
    for (int q = 0; q < HP; ++q) {
      dF[q] = __fadd_rn(nxt.F[q], -cur.F[q]);
      ph[q] = 0.f;
      off[q] = 0.f;
      Fprev[q] = cur.F[q];
      dFprev[q] = 0.f;
      const int h = lane + 32 * q;
      if (!ENDS_ONLY && c > 0 && h < a.H) off[q] = a.offsets[off_base + h];
    }
    if (!FAST && k > 0) {
      FrameRegs<HP> prv;
      load_frame<HP, false>(a, row, s, k - 1, lane, prv);
#pragma unroll
      for (int q = 0; q < HP; ++q) {
        Fprev[q] = prv.F[q];
        dFprev[q] = __fadd_rn(cur.F[q], -prv.F[q]);
      }
    }

    float tf = (float)t0;
    int t = t0;
    int pos = 0;
    while (t < t1) {
      float y[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) y[i] = 0.f;
#pragma unroll
      for (int gi = 0; gi < 32 / G; ++gi) {
        if (t < t1) {
          if (r == a.U) {   // next control frame (warp-uniform)
            r = 0;
            ++k;
#pragma unroll
            for (int q = 0; q < HP; ++q) {
              Fprev[q] = cur.F[q];
              dFprev[q] = dF[q];
              cur.F[q] = nxt.F[q];
              cur.A[q] = nxt.A[q];
            }
            load_frame<HP, !ENDS_ONLY>(a, row, s, min(k + 1, a.F - 1), lane, nxt);
#pragma unroll
            for (int q = 0; q < HP; ++q) dF[q] = __fadd_rn(nxt.F[q], -cur.F[q]);
            if (surrogate) load_decays(k);
          }
          const float kf = (float)k;
#pragma unroll
          for (int j = 0; j < G; ++j) {
            // legacy ResizeBilinear coordinates: in = float(i) * scale, lerp = in - floor(in)
            const float in = __fmul_rn(tf, a.scale);
            float frac = __fadd_rn(in, -kf);
            bool below = false;
            if (!FAST) {
              const float lo = floorf(in);
              below = lo < kf;
              frac = __fadd_rn(in, -lo);
            }
            float w0 = 0.f, w1 = 0.f;
            if (!ENDS_ONLY) {
              w0 = win[r + j];          // rising half: weight of frame k+1
              w1 = win[r + j + a.U];    // falling half: weight of frame k
            }
#pragma unroll
            for (int q = 0; q < HP; ++q) {
              float f;
              if (!FAST && below) {
                f = __fadd_rn(Fprev[q], __fmul_rn(dFprev[q], frac));
              } else {
                f = __fadd_rn(cur.F[q], __fmul_rn(dF[q], frac));       // top + (bottom-top)*lerp
              }
              const float om = div_sr<FAST>(__fmul_rn(f, two_pi), a.sr, a.inv_sr, a.inv_sr_lo);  // :69-70
              ph[q] = __fadd_rn(ph[q], om);                            // in-chunk cumsum
              if (!ENDS_ONLY) {
                float amp = __fmaf_rn(cur.A[q], w1, __fmul_rn(nxt.A[q], w0));
                if (surrogate) amp = __fmul_rn(amp, powf(D[q], __fadd_rn(T0, (float)(r + j))));
                amp = (f >= a.nyquist) ? 0.f : amp;                    // :65-67
                const float cv = a.plain ? cos_large(ph[q])                     // tf.cos(tf.cumsum), :76-77
                                         : __cosf(wrap_to_pi(__fadd_rn(ph[q], off[q])));
                y[gi * G + j] = __fmaf_rn(amp, cv, y[gi * G + j]);              // :80-83
              }
            }
            tf += 1.0f;
          }
          r += G;
          t += G;
        }
      }
      if (!ENDS_ONLY) {
        const float ysum = transpose_reduce32(y, lane);
        if (a.plain) {
          // one "chunk" = the whole clip: no staging row; the (voice, substring) warps add straight into the
          // zeroed output (float atomics: the order of the voices is not fixed in this mode)
          if (t0 + pos + lane < t1) atomicAdd(a.out + ((size_t)g * a.B + b) * a.N + t0 + pos + lane, ysum);
        } else {
          rows[warp * kMaxChunk + pos + lane] += ysum;   // pos + lane < kMaxChunk always
        }
      }
      pos += 32;
    }

    if (ENDS_ONLY) {
      // chunk end phase mod 2pi (ddsp angular_cumsum: offsets = phase[:, :, -1] % 2pi)
#pragma unroll
      for (int q = 0; q < HP; ++q) {
        const int h = lane + 32 * q;
        if (h < a.H) a.offsets[off_base + h] = floormod_two_pi(ph[q]);
      }
    }
  }

  if (!ENDS_ONLY && !a.plain) {
    __syncthreads();
    float* out = a.out + ((size_t)g * a.B + b) * a.N + t0;
    const int len = t1 - t0;
    for (int i = threadIdx.x; i < len; i += n_threads) {
      float acc = rows[i];
      for (int w = 1; w < n_warps; ++w) acc += rows[w * kMaxChunk + i];
      if (a.accumulate) acc += out[i];
      out[i] = acc;
    }
  }
}

// Chunk end phases -> chunk offsets, in place (ddsp angular_cumsum: shift down one chunk, cumsum over
// chunks in float32, then mod 2pi).  The running sum is sequential per oscillator (float32 addition is
// not associative and the reference's order is part of its output), so: one CTA per (row, substring,
// block of 32 partials); all threads stage a tile of chunk ends in shared memory (coalesced: 32
// consecutive partials per chunk), ONE warp walks the tile adding sequentially, all threads store the
// wrapped offsets.
// ends_na (optional): 16-partial half-groups >= ends_na[row, c] were not computed for chunk c and count
// as 0 (no later chunk reads their offset).
// Spans of a timeline (b200ddsp_span): the sum starts from the predecessor's state `link.seed` instead
// of 0 and its final value (including the LAST chunk's end, carry_all) goes to `link.carry` -- which
// may be the successor GPU's inbox; see link.cuh for the hand-off.
constexpr int kOffTileMax = 1536;   // chunks per shared-memory tile at most (33 floats each: 198 KB)

struct OffsetsArgs {
  float* offsets;                 // [n_osc_rows, n_chunks, H]
  const unsigned char* ends_na;   // [n_osc_rows / S, n_chunks] or nullptr
  int n_osc_rows, n_chunks, H, S;
  int carry_all;                  // the last chunk's end phase is part of the sum (a carry is wanted)
  int tile_chunks;                // chunks per tile: min(n_chunks, kOffTileMax); dynamic smem = 33 floats each
  Link link;                      // payload [n_osc_rows, H]
};

// The tile holds as many chunks as fit (a config-4 span of 1152 chunks is ONE tile): all loads of the tile are
// in flight together, one warp then adds through it, and -- what the next rank of a chain is waiting for --
// the carry leaves and its flag is raised BEFORE the wrapped offsets are written back.
__global__ void __launch_bounds__(512) additive_offsets_kernel(const OffsetsArgs a) {
  extern __shared__ float off_tile[];                  // [tile_chunks][33] floats, then [tile_chunks] bytes
  unsigned char* na_s = reinterpret_cast<unsigned char*>(off_tile + (size_t)a.tile_chunks * 33);
  const int n_warps = blockDim.x >> 5;                 // 8, or 16 for long spans (more loads in flight)
  const int hb = (a.H + 31) / 32;
  const int rs = blockIdx.x / hb, h0 = (blockIdx.x - rs * hb) * 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int h = h0 + lane;
  const int group = h >> 4;   // liveness is counted in 16-partial half-groups
  const unsigned char* na = a.ends_na ? a.ends_na + (size_t)(rs / a.S) * a.n_chunks : nullptr;
  float* p = a.offsets + (size_t)rs * a.n_chunks * a.H + h;
  const Link& lk = a.link;
  float cum = 0.f;
  if (lk.seed != nullptr) {
    if (threadIdx.x == 0) link_wait(lk.seed_ready, lk.epoch, lk.scratch);   // (a gate kernel has waited already)
    __syncthreads();
    if (warp == 0 && h < a.H) cum = ld_inbox(lk.seed + (size_t)rs * a.H + h);
  }
  const int last_end = a.carry_all ? a.n_chunks : a.n_chunks - 1;   // chunks whose end phase counts
  bool arrived = false;
  for (int c0 = 0; c0 < a.n_chunks; c0 += a.tile_chunks) {
    const int nc = min(a.tile_chunks, a.n_chunks - c0);
    for (int j = threadIdx.x; j < nc; j += blockDim.x) na_s[j] = na ? na[c0 + j] : (unsigned char)255;
    __syncthreads();
    // 16 rows per warp and round, every load unconditional (clamped to a valid address; what does not count is
    // replaced by 0 afterwards) so that all 16 are in flight before the first is stored
    const float* p_safe = a.offsets + (size_t)rs * a.n_chunks * a.H + min(h, a.H - 1);
    for (int j0 = warp; j0 < nc; j0 += 16 * n_warps) {
      float v[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        const int j = j0 + u * n_warps;
        v[u] = 0.f;
        if (j < nc) v[u] = p_safe[(size_t)(c0 + j) * a.H];      // predicated, not branched: still 16 in flight
      }
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        const int j = j0 + u * n_warps;
        if (j < nc) {
          const bool have = (h < a.H) && (c0 + j < last_end) && (group < (int)na_s[j]);
          off_tile[j * 33 + lane] = have ? v[u] : 0.f;
        }
      }
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll 8
      for (int j = 0; j < nc; ++j) {
        const float e = off_tile[j * 33 + lane];
        off_tile[j * 33 + lane] = cum;   // offset of chunk c0 + j, unwrapped
        cum = __fadd_rn(cum, e);
      }
    }
    if (c0 + nc >= a.n_chunks) {         // last tile: hand the state on first
      if (lk.carry != nullptr) {
        if (threadIdx.x == 0 && lk.carry_ack != nullptr && lk.epoch > 2)
          link_wait(lk.carry_ack, lk.epoch - 2, lk.scratch);
        __syncthreads();
        if (warp == 0 && h < a.H) lk.carry[(size_t)rs * a.H + h] = cum;
      }
      link_arrive(lk, gridDim.x, lk.carry != nullptr, lk.seed != nullptr, warp == 0);
      arrived = true;
    }
    __syncthreads();
    for (int j = warp; j < nc; j += n_warps)
      if (h < a.H) p[(size_t)(c0 + j) * a.H] = floormod_two_pi(off_tile[j * 33 + lane]);
    __syncthreads();
  }
  if (!arrived) link_arrive(lk, gridDim.x, lk.carry != nullptr, lk.seed != nullptr);   // n_chunks == 0
}

}  // namespace b200ddsp
