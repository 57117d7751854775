// Convolution reverb: ddsp.effects.Reverb.get_signal as configured by the reference
// (configs/dafx22.gin:99-100,111): ir[:, 0] = 0, wet = fft_convolve(audio, ir, 'same',
// delay_compensation=0) -- ONE block of size n = 2^ceil(log2(N + L - 1)) -- out = wet (+ audio).
//
// Both operands are real, so one complex transform carries both: z = audio + i*ir,
//   X[k] = (Z[k] + conj Z[n-k]) / 2,   H[k] = (Z[k] - conj Z[n-k]) / 2i,   Y = X * H,
// and because the wet signals are real, two clips share one inverse transform:
//   FFT(conj Ya + i conj Yb) = n * (ya + i yb).
// The transform is a Stockham autosort FFT: radix-16 passes (plus one radix-2/4/8 pass when
// log2 n is not a multiple of 4) between two ping-pong buffers that stay resident in the
// 126 MB L2 (2 MB per clip at n = 2^18).  Twiddles come from a table exp(-2 pi i q / n)
// evaluated in double precision, so the float32 error is the butterflies' only.
#pragma once
#include "common.cuh"

namespace b200ddsp {

constexpr int kFftThreads = 256;

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(__fmaf_rn(a.x, b.x, -a.y * b.y), __fmaf_rn(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 mul_neg_i(float2 a) { return make_float2(a.y, -a.x); }

__device__ __forceinline__ void dft2(float2& a, float2& b) {
  const float2 t = a;
  a = cadd(t, b);
  b = csub(t, b);
}

__device__ __forceinline__ void dft4(float2& v0, float2& v1, float2& v2, float2& v3) {
  const float2 t0 = cadd(v0, v2), t1 = csub(v0, v2);
  const float2 t2 = cadd(v1, v3), t3 = mul_neg_i(csub(v1, v3));
  v0 = cadd(t0, t2);
  v1 = cadd(t1, t3);
  v2 = csub(t0, t2);
  v3 = csub(t1, t3);
}

// exp(-2 pi i m / 16), m = 0..15 (only m = b*c with b, c < 4 is used)
__device__ __forceinline__ float2 w16(int m) {
  const float c1 = 0.92387953251128674f, s1 = 0.38268343236508977f, h = 0.70710678118654752f;
  switch (m & 15) {
    case 0: return make_float2(1.f, 0.f);
    case 1: return make_float2(c1, -s1);
    case 2: return make_float2(h, -h);
    case 3: return make_float2(s1, -c1);
    case 4: return make_float2(0.f, -1.f);
    case 5: return make_float2(-s1, -c1);
    case 6: return make_float2(-h, -h);
    case 7: return make_float2(-c1, -s1);
    case 8: return make_float2(-1.f, 0.f);
    case 9: return make_float2(-c1, s1);
    case 10: return make_float2(-h, h);
    case 11: return make_float2(-s1, c1);
    case 12: return make_float2(0.f, 1.f);
    case 13: return make_float2(s1, c1);
    case 14: return make_float2(h, h);
    default: return make_float2(c1, s1);
  }
}

// In-register forward DFT of R points, natural order in and out.
template <int R>
__device__ __forceinline__ void dft(float2 (&v)[R]);

template <>
__device__ __forceinline__ void dft<2>(float2 (&v)[2]) { dft2(v[0], v[1]); }

template <>
__device__ __forceinline__ void dft<4>(float2 (&v)[4]) { dft4(v[0], v[1], v[2], v[3]); }

template <>
__device__ __forceinline__ void dft<8>(float2 (&v)[8]) {
  // r = 2a + b, q = c + 4d: DFT4 over a, twiddle W8^(bc), DFT2 over b
  dft4(v[0], v[2], v[4], v[6]);   // b = 0: T0[c] in v[2c]
  dft4(v[1], v[3], v[5], v[7]);   // b = 1: T1[c] in v[2c+1]
#pragma unroll
  for (int c = 1; c < 4; ++c) v[2 * c + 1] = cmul(v[2 * c + 1], w16(2 * c));
  float2 o[8];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    float2 a = v[2 * c], b = v[2 * c + 1];
    dft2(a, b);
    o[c] = a;
    o[c + 4] = b;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = o[i];
}

template <>
__device__ __forceinline__ void dft<16>(float2 (&v)[16]) {
  // r = 4a + b, q = c + 4d: DFT4 over a, twiddle W16^(bc), DFT4 over b
#pragma unroll
  for (int b = 0; b < 4; ++b) dft4(v[b], v[4 + b], v[8 + b], v[12 + b]);   // T_b[c] in v[4c+b]
#pragma unroll
  for (int b = 1; b < 4; ++b)
#pragma unroll
    for (int c = 1; c < 4; ++c) v[4 * c + b] = cmul(v[4 * c + b], w16(b * c));
  float2 o[16];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    float2 t0 = v[4 * c], t1 = v[4 * c + 1], t2 = v[4 * c + 2], t3 = v[4 * c + 3];
    dft4(t0, t1, t2, t3);
    o[c] = t0;
    o[c + 4] = t1;
    o[c + 8] = t2;
    o[c + 12] = t3;
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = o[i];
}

// ---- loaders / storers ---------------------------------------------------------------------
struct LoadComplex {
  const float2* src; int n;
  __device__ __forceinline__ float2 operator()(int batch, int i) const {
    return src[(size_t)batch * n + i];
  }
};
struct StoreComplex {
  float2* dst; int n;
  __device__ __forceinline__ void operator()(int batch, int i, float2 v) const {
    dst[(size_t)batch * n + i] = v;
  }
};
// Per-clip power-of-two normalisation.  Audio and IR travel through ONE complex transform and are
// separated afterwards as (Z[k] +- conj Z[n-k]) / 2, so rounding noise of the larger operand leaks
// into the smaller one: without normalisation the error of the wet signal grows like
// |audio| / |ir| (quadratically with the input level).  Scaling both to [0.5, 1) by exact powers
// of two (undone exactly in the final store) makes the error independent of the levels.
// scales[b] = (2^-ea, 2^-ei, 2^(ea+ei)) with 2^ea >= max|audio[b]|, 2^ei >= max|ir[b, 1:]|.
// Step 1: per-clip maxima (bit patterns of non-negative floats order like unsigned ints), many
// CTAs per clip; maxima must be zeroed beforehand.  Step 2: one thread per clip -> scales.
__global__ void __launch_bounds__(256) reverb_maxima_kernel(const float* __restrict__ audio,
                                                            const float* __restrict__ ir,
                                                            unsigned int* __restrict__ maxima, int N,
                                                            int L, int first_tap) {
  const int b = blockIdx.y;
  const int stride = gridDim.x * blockDim.x;
  const int i0 = blockIdx.x * blockDim.x + threadIdx.x;
  float ma = 0.f, mi = 0.f;
  for (int i = i0; i < N; i += stride) ma = fmaxf(ma, fabsf(__ldg(audio + (size_t)b * N + i)));
  for (int i = first_tap + i0; i < L; i += stride) mi = fmaxf(mi, fabsf(__ldg(ir + (size_t)b * L + i)));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ma = fmaxf(ma, __shfl_xor_sync(0xffffffffu, ma, o));
    mi = fmaxf(mi, __shfl_xor_sync(0xffffffffu, mi, o));
  }
  if ((threadIdx.x & 31) == 0) {
    // NaN compares false everywhere above and is dropped; +inf survives and disables scaling
    atomicMax(maxima + 2 * b, __float_as_uint(ma));
    atomicMax(maxima + 2 * b + 1, __float_as_uint(mi));
  }
}

__global__ void reverb_scales_kernel(const unsigned int* __restrict__ maxima,
                                     float4* __restrict__ scales, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float ma = __uint_as_float(maxima[2 * b]), mi = __uint_as_float(maxima[2 * b + 1]);
  // exponent e with 2^e > m (finite); m == 0 or non-finite -> no scaling
  int ea = 0, ei = 0;
  if (ma > 0.f && ma < 3.0e38f) frexpf(ma, &ea);
  if (mi > 0.f && mi < 3.0e38f) frexpf(mi, &ei);
  ea = max(-60, min(60, ea));
  ei = max(-60, min(60, ei));
  scales[b] = make_float4(ldexpf(1.f, -ea), ldexpf(1.f, -ei), ldexpf(1.f, ea + ei), 0.f);
}

// z = audio / 2^ea + i * ir / 2^ei, zero padded to n; first_tap = 1 masks ir[0]
// (Reverb._mask_dry_ir)
struct LoadAudioIr {
  const float* audio; const float* ir; const float4* scales; int N, L, first_tap;
  __device__ __forceinline__ float2 operator()(int batch, int i) const {
    const float4 sc = __ldg(scales + batch);
    const float re = (i < N) ? __ldg(audio + (size_t)batch * N + i) * sc.x : 0.f;
    const float im = (i >= first_tap && i < L) ? __ldg(ir + (size_t)batch * L + i) * sc.y : 0.f;
    return make_float2(re, im);
  }
};
// real part -> clip 2*batch, imaginary part -> clip 2*batch+1; crop to n_out samples
// (N for padding 'same' with delay_compensation = 0; N + L - 1 for the 'valid' form used by the
// timeline overlap-add), scale by 1/n, add the dry signal (add_dry needs n_out <= N)
struct StoreWetPair {
  float* out; const float* audio; const float4* scales; int N, n_out, B; float inv_n; int add_dry;
  __device__ __forceinline__ void operator()(int batch, int i, float2 v) const {
    if (i >= n_out) return;
    const int b0 = 2 * batch, b1 = b0 + 1;
    float y0 = v.x * inv_n * __ldg(scales + b0).z;
    if (add_dry) y0 = __fadd_rn(y0, __ldg(audio + (size_t)b0 * N + i));
    out[(size_t)b0 * n_out + i] = y0;
    if (b1 < B) {
      float y1 = v.y * inv_n * __ldg(scales + b1).z;
      if (add_dry) y1 = __fadd_rn(y1, __ldg(audio + (size_t)b1 * N + i));
      out[(size_t)b1 * n_out + i] = y1;
    }
  }
};

// ---- split form: impulse responses first, audio later ---------------------------------------------
// Inside the polyphonic forward the impulse responses are known from the start while the dry signal
// exists only at the very end, so the forward transforms of the IRs (two real IRs per complex
// transform) run early on a side stream and the tail of the forward is left with 8 + 8 transforms
// (two dry clips per forward transform, two wet clips per inverse) instead of 16 + 8.
// maxima[2 b + which]: which = 0 audio, 1 impulse response (entries zeroed beforehand).
__global__ void __launch_bounds__(256) reverb_maxima1_kernel(const float* __restrict__ x,
                                                             unsigned int* __restrict__ maxima, int len,
                                                             int first, int which) {
  const int b = blockIdx.y;
  const int stride = gridDim.x * blockDim.x;
  float m = 0.f;
  for (int i = first + blockIdx.x * blockDim.x + threadIdx.x; i < len; i += stride)
    m = fmaxf(m, fabsf(__ldg(x + (size_t)b * len + i)));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(maxima + 2 * b + which, __float_as_uint(m));
}

// z = x[2 pair] * s + i * x[2 pair + 1] * s', zero padded; which selects the audio (.x) or the
// impulse-response (.y) scale of each clip; first = 1 masks tap 0 (Reverb._mask_dry_ir)
struct LoadRealPair {
  const float* x; const float4* scales; int len, first, B, which;
  __device__ __forceinline__ float2 operator()(int pair, int i) const {
    const int b0 = 2 * pair, b1 = b0 + 1;
    if (i < first || i >= len) return make_float2(0.f, 0.f);
    const float4 s0 = __ldg(scales + b0);
    const float re = __ldg(x + (size_t)b0 * len + i) * (which ? s0.y : s0.x);
    float im = 0.f;
    if (b1 < B) {
      const float4 s1 = __ldg(scales + b1);
      im = __ldg(x + (size_t)b1 * len + i) * (which ? s1.y : s1.x);
    }
    return make_float2(re, im);
  }
};

// Za = FFT(a0 + i a1), Zh = FFT(h0 + i h1) of one pair of clips -> V[k] = conj Y0[k] + i conj Y1[k],
// Y0 = A0 H0, Y1 = A1 H1 (the real signals separated by Hermitian symmetry).
__global__ void __launch_bounds__(256) reverb_spectrum_split_kernel(const float2* __restrict__ Za,
                                                                     const float2* __restrict__ Zh,
                                                                     float2* __restrict__ V, int n) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;   // 0 .. n/2
  const int pair = blockIdx.y;
  if (k > n / 2) return;
  const int kn = (n - k) & (n - 1);
  const float2 za = Za[(size_t)pair * n + k], wa = Za[(size_t)pair * n + kn];
  const float2 zh = Zh[(size_t)pair * n + k], wh = Zh[(size_t)pair * n + kn];
  // even part -> first signal, odd part / i -> second signal
  const float2 A0 = make_float2(0.5f * (za.x + wa.x), 0.5f * (za.y - wa.y));
  const float2 A1 = make_float2(0.5f * (za.y + wa.y), -0.5f * (za.x - wa.x));
  const float2 H0 = make_float2(0.5f * (zh.x + wh.x), 0.5f * (zh.y - wh.y));
  const float2 H1 = make_float2(0.5f * (zh.y + wh.y), -0.5f * (zh.x - wh.x));
  const float2 y0 = cmul(A0, H0), y1 = cmul(A1, H1);
  V[(size_t)pair * n + k] = make_float2(y0.x + y1.y, y1.x - y0.y);
  if (kn != k) V[(size_t)pair * n + kn] = make_float2(y0.x - y1.y, y0.y + y1.x);
}

// One Stockham pass of radix R: sub-transforms of length Ns -> Ns*R.
template <int R, class Loader, class Storer>
__global__ void __launch_bounds__(kFftThreads) fft_pass_kernel(const Loader ld, const Storer st,
                                                              const float2* __restrict__ tw,
                                                              int n, int Ns) {
  const int j = blockIdx.x * kFftThreads + threadIdx.x;
  const int batch = blockIdx.y;
  const int stride = n / R;
  if (j >= stride) return;
  float2 v[R];
#pragma unroll
  for (int r = 0; r < R; ++r) v[r] = ld(batch, j + r * stride);
  const int k = j & (Ns - 1);
  if (Ns > 1) {
    const int step = k * (stride / Ns);    // k * n / (Ns * R)
#pragma unroll
    for (int r = 1; r < R; ++r) v[r] = cmul(v[r], __ldg(tw + r * step));
  }
  dft<R>(v);
  const int base = (j - k) * R + k;
#pragma unroll
  for (int r = 0; r < R; ++r) st(batch, base + r * Ns, v[r]);
}

// One Stockham pass of radix 64 = 8 x 8 through shared memory: a CTA of 128 threads takes 16
// consecutive butterflies (1024 points).  Stage 1: thread (jj, a) loads points r = a + 8m with
// the inter-pass twiddle, does the 8-point DFT over m, applies W64^(a c); stage 2, after a
// shared-memory exchange: thread (jj, c) does the 8-point DFT over a and owns outputs q = c + 8d.
// Global loads are 128-byte runs (16 consecutive j); stores are 128-byte runs when Ns >= 16 and
// go through shared memory (SMALL_NS) when the output runs are shorter.
constexpr int kFft64J = 16;
constexpr int kFft64Threads = kFft64J * 8;

// BULK (complex source only): the CTA's tile -- 64 runs of 16 consecutive points, one per residue r --
// is staged in shared memory by 64 bulk asynchronous copies of 128 bytes (cp.async.bulk, the 1-D form of
// TMA; one per thread of the first two warps) that complete on an mbarrier, instead of 8 strided 8-byte
// loads per thread.  A/B against the plain loads: DESIGN.md 4.3 (B200DDSP_FFT_BULK).
template <class Loader, class Storer, bool SMALL_NS, bool BULK = false>
__global__ void __launch_bounds__(kFft64Threads) fft_pass64_kernel(const Loader ld, const Storer st,
                                                                  const float2* __restrict__ tw,
                                                                  int n, int Ns) {
  __shared__ float2 S[kFft64J * 65];
  __shared__ __align__(128) float2 T[BULK ? 64 * kFft64J : 1];
  __shared__ __align__(8) unsigned long long bar;
  const int jj = threadIdx.x & (kFft64J - 1);
  const int a = threadIdx.x >> 4;                 // stage 1: residue of r mod 8; stage 2: c
  const int batch = blockIdx.y;
  const int stride = n >> 6;                      // n / 64
  const int j = blockIdx.x * kFft64J + jj;        // n >= 1024: every j is valid
  const int k = j & (Ns - 1);
  float2 v[8];
  if constexpr (BULK) {
    if (threadIdx.x == 0) {
      mbar_init(&bar, 1);
      mbar_expect_tx(&bar, 64 * kFft64J * (unsigned int)sizeof(float2));
    }
    __syncthreads();
    if (threadIdx.x < 64)
      bulk_load(&T[threadIdx.x * kFft64J], ld.src + (size_t)batch * n + (size_t)blockIdx.x * kFft64J +
                                               (size_t)threadIdx.x * stride,
                kFft64J * (unsigned int)sizeof(float2), &bar);
    mbar_wait(&bar, 0);
#pragma unroll
    for (int m = 0; m < 8; ++m) v[m] = T[(a + 8 * m) * kFft64J + jj];
  } else {
#pragma unroll
    for (int m = 0; m < 8; ++m) v[m] = ld(batch, j + (a + 8 * m) * stride);
  }
  if (Ns > 1) {
    const int step = k * (stride / Ns);           // k * n / (Ns * 64)
#pragma unroll
    for (int m = 0; m < 8; ++m) {
      const int r = a + 8 * m;
      if (r > 0) v[m] = cmul(v[m], __ldg(tw + r * step));
    }
  }
  dft<8>(v);                                      // over m -> index c
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    float2 x = v[c];
    if (a * c > 0) x = cmul(x, __ldg(tw + (a * c) * stride));   // W64^(a c)
    S[jj * 65 + c * 8 + a] = x;
  }
  __syncthreads();
  const int c = a;
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = S[jj * 65 + c * 8 + i];
  dft<8>(v);                                      // over a -> index d, output q = c + 8 d
  if (!SMALL_NS) {
    const int base = (j - k) * 64 + k;
#pragma unroll
    for (int d = 0; d < 8; ++d) st(batch, base + (c + 8 * d) * Ns, v[d]);
  } else {
    // outputs of the CTA's 16 butterflies: q-major inside a butterfly when Ns == 1; gather them in
    // shared memory and let consecutive threads store consecutive addresses as far as they go
    __syncthreads();
#pragma unroll
    for (int d = 0; d < 8; ++d) S[jj * 65 + c + 8 * d] = v[d];
    __syncthreads();
    // output address of (j, q) = (j - k) * 64 + k + q * Ns: for fixed q-block the run over k is
    // contiguous; enumerate (jblock, q, k) so that k is fastest
    const int j0 = blockIdx.x * kFft64J;          // multiple of 16 >= Ns (Ns in {1,2,4,8})
    for (int e = threadIdx.x; e < kFft64J * 64; e += kFft64Threads) {
      const int kk = e & (Ns - 1);
      const int q = (e / Ns) & 63;
      const int jb = e / (Ns * 64);               // which Ns-block of the CTA's 16 j's
      const int jl = jb * Ns + kk;                // local j
      const int jg = j0 + jl;
      st(batch, (jg - kk) * 64 + kk + q * Ns, S[jl * 65 + q]);
    }
  }
}

// The same pass as a PERSISTENT kernel with a two-stage bulk-asynchronous pipeline (complex source only):
// a CTA walks tiles (batch, 16 consecutive butterflies) with stride gridDim.x; while it transforms tile i
// out of stage i & 1, the 64 bulk copies of tile i + 1 (128 bytes each, cp.async.bulk -> UBLKCP) fill the
// other stage and complete on that stage's mbarrier.  The global-load latency that the one-shot kernel
// covers with occupancy alone (stalls: long_scoreboard + lg_throttle) is hidden behind the butterflies.
template <class Storer, bool SMALL_NS>
__global__ void __launch_bounds__(kFft64Threads) fft_pass64_pipelined_kernel(const float2* __restrict__ src,
                                                                            const Storer st,
                                                                            const float2* __restrict__ tw,
                                                                            int n, int Ns, int batches) {
  __shared__ float2 S[kFft64J * 65];
  __shared__ __align__(128) float2 T[2][64 * kFft64J];
  __shared__ __align__(8) unsigned long long bar[2];
  const int jj = threadIdx.x & (kFft64J - 1);
  const int a = threadIdx.x >> 4;                 // stage 1: residue of r mod 8; stage 2: c
  const int stride = n >> 6;                      // n / 64
  const int tiles_per_batch = stride / kFft64J;
  const int n_tiles = tiles_per_batch * batches;
  if (threadIdx.x == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
  }
  __syncthreads();
  auto issue = [&](int tile, int stage) {         // threads 0..63: one 128-byte run each
    const int batch = tile / tiles_per_batch, jb = tile - batch * tiles_per_batch;
    if (threadIdx.x == 0) mbar_expect_tx(&bar[stage], 64 * kFft64J * (unsigned int)sizeof(float2));
    // the expect_tx of thread 0 and the copies of the other threads may reach the barrier in any order: the
    // phase cannot complete before the arrival that carries the expected byte count
    if (threadIdx.x < 64)
      bulk_load(&T[stage][threadIdx.x * kFft64J],
                src + (size_t)batch * n + (size_t)jb * kFft64J + (size_t)threadIdx.x * stride,
                kFft64J * (unsigned int)sizeof(float2), &bar[stage]);
  };
  int tile = blockIdx.x;
  if (tile < n_tiles) issue(tile, 0);
  unsigned int parity[2] = {0u, 0u};
  for (int it = 0; tile < n_tiles; tile += gridDim.x, ++it) {
    const int stage = it & 1;
    const int next = tile + gridDim.x;
    if (next < n_tiles) issue(next, stage ^ 1);   // stage ^ 1 was released by the barrier that ended tile it - 1
    mbar_wait(&bar[stage], parity[stage]);
    parity[stage] ^= 1u;
    const int batch = tile / tiles_per_batch;
    const int j = (tile - batch * tiles_per_batch) * kFft64J + jj;
    const int k = j & (Ns - 1);
    float2 v[8];
#pragma unroll
    for (int m = 0; m < 8; ++m) v[m] = T[stage][(a + 8 * m) * kFft64J + jj];
    if (Ns > 1) {
      const int step = k * (stride / Ns);
#pragma unroll
      for (int m = 0; m < 8; ++m) {
        const int r = a + 8 * m;
        if (r > 0) v[m] = cmul(v[m], __ldg(tw + r * step));
      }
    }
    dft<8>(v);
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      float2 x = v[c];
      if (a * c > 0) x = cmul(x, __ldg(tw + (a * c) * stride));
      S[jj * 65 + c * 8 + a] = x;
    }
    __syncthreads();
    const int c = a;
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = S[jj * 65 + c * 8 + i];
    dft<8>(v);
    if (!SMALL_NS) {
      const int base = (j - k) * 64 + k;
#pragma unroll
      for (int d = 0; d < 8; ++d) st(batch, base + (c + 8 * d) * Ns, v[d]);
    } else {
      __syncthreads();
#pragma unroll
      for (int d = 0; d < 8; ++d) S[jj * 65 + c + 8 * d] = v[d];
      __syncthreads();
      const int j0 = (tile - batch * tiles_per_batch) * kFft64J;
      for (int e = threadIdx.x; e < kFft64J * 64; e += kFft64Threads) {
        const int kk = e & (Ns - 1);
        const int q = (e / Ns) & 63;
        const int jb = e / (Ns * 64);
        const int jl = jb * Ns + kk;
        const int jg = j0 + jl;
        st(batch, (jg - kk) * 64 + kk + q * Ns, S[jl * 65 + q]);
      }
    }
    __syncthreads();   // S and T[stage] are free again (generic reads done before the next async writes)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
}

// MultiInstrumentReverb.exponential_decay_mask (modules/sub_modules.py:339-349, inference only):
// out = ir * concat(ones(start), exp(-e * linspace(0, 1, L - start))), tf.linspace evaluated in
// float32 (start + i * step, last element exactly 1).
__global__ void __launch_bounds__(256) ir_decay_mask_kernel(const float* __restrict__ ir,
                                                            float* __restrict__ out, int L, int start,
                                                            float exponent) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= L) return;
  const size_t o = (size_t)blockIdx.y * L + i;
  float m = 1.0f;
  if (i >= start) {
    const int n = L - start, j = i - start;
    const float step = (n > 1) ? __fdiv_rn(1.0f, (float)(n - 1)) : 0.f;
    const float t = (j == n - 1 && n > 1) ? 1.0f : __fmul_rn((float)j, step);
    m = expf(-exponent * t);
  }
  out[o] = ir[o] * m;
}

// tw[q] = exp(-2 pi i q / n)
__global__ void __launch_bounds__(256) fft_twiddle_kernel(float2* tw, int n) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n) return;
  double s, c;
  sincospi(-2.0 * (double)q / (double)n, &s, &c);
  tw[q] = make_float2((float)c, (float)s);
}

// Z (B clips) -> V (ceil(B/2) pairs): V[k] = conj Ya[k] + i conj Yb[k], Y = X * H.
__global__ void __launch_bounds__(256) reverb_spectrum_kernel(const float2* __restrict__ Z,
                                                               float2* __restrict__ V, int n,
                                                               int B) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;   // 0 .. n/2
  const int pair = blockIdx.y;
  if (k > n / 2) return;
  const int kn = (n - k) & (n - 1);
  float2 y[2];
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const int b = 2 * pair + e;
    y[e] = make_float2(0.f, 0.f);
    if (b < B) {
      const float2 z = Z[(size_t)b * n + k], w = Z[(size_t)b * n + kn];
      const float2 X = make_float2(0.5f * (z.x + w.x), 0.5f * (z.y - w.y));
      const float2 H = make_float2(0.5f * (z.y + w.y), -0.5f * (z.x - w.x));
      y[e] = cmul(X, H);
    }
  }
  // V[k]   = conj(Ya) + i conj(Yb) = (ya.x + yb.y) + i (yb.x - ya.y)
  // V[n-k] = Ya + i Yb             = (ya.x - yb.y) + i (ya.y + yb.x)      (Y is Hermitian)
  V[(size_t)pair * n + k] = make_float2(y[0].x + y[1].y, y[1].x - y[0].y);
  if (kn != k) V[(size_t)pair * n + kn] = make_float2(y[0].x - y[1].y, y[0].y + y[1].x);
}

}  // namespace b200ddsp
