// Fbank front-end in ONE kernel for sm_100a: windowed DFT -> power -> mel filterbank -> dB (+ per-utterance max).
//
// Replaces torch.stft (cuFFT) + cm_fbank_logmel for the two transform sizes the reference YAMLs use (n_fft = 400 and 512,
// hop 160; speechbrain.lobes.features.Fbank: reference call sites train_CTC.py:285, train_S2S.py:349, YAML
// hparams/CTC/conmamba_large.yaml:322-326).  The cuFFT route writes the complex STFT (B, F, T) to HBM - 206 MB for
// 64 x 20 s - and reads it back; here the samples are read once (1.3 MB per second of batch) and only the (B, T, M) log-mel
// features are written.
//
// A CTA owns 8 consecutive frames of one utterance.  The n_fft-point real DFT is evaluated as a two-stage decomposition
// n = N2*n1 + n2, k = k1 + 16*k2 (N = 16 * N2, N2 = 25 or 32):
//   stage 1  A[n2][k1]  = sum_n1 x[N2*n1 + n2] * W16^(n1*k1)           16-point DFTs of real data (k1 = 0..8, rest by symmetry)
//            A'[n2][k1] = A[n2][k1] * WN^(n2*k1)                       twiddles from a per-CTA shared-memory table
//   stage 2  X[k1 + 16*k2] = sum_n2 A'[n2][k1] * WN2^(n2*k2)           N2-point DFTs, only the bins k <= N/2
// Every loop is fully unrolled and the W16 / WN2 factors are compile-time constants (tools/gen_dft_tables.py), i.e. FFMA
// immediates: no table loads in the inner loops, ~44 k FFMA per frame at N = 512 against 263 k for the direct DFT.  fp32
// throughout (a tensor-core DFT would need three TF32 products per term to keep weak bins above the rounding floor of the
// strong ones - more issue slots than this form).  Epilogue as in fbank.cu: sparse triangular filterbank over each filter's
// support (host-provided [lo, hi) per mel), dB, running max by ordered-int atomics; cm_fbank_floor applies max - top_db.
#include <climits>
#include <math_constants.h>

#include "common.cuh"
#include "dft_tables.cuh"

namespace cm {
namespace dft {

constexpr int kDF = 8;            // frames per CTA
constexpr int kThreads = 256;

template <int N2> struct Tw;
template <> struct Tw<25> {
  static __device__ __forceinline__ constexpr float c(int i) { return kCos25[i]; }
  static __device__ __forceinline__ constexpr float s(int i) { return kSin25[i]; }
};
template <> struct Tw<32> {
  static __device__ __forceinline__ constexpr float c(int i) { return kCos32[i]; }
  static __device__ __forceinline__ constexpr float s(int i) { return kSin32[i]; }
};

template <int N2>
struct Smem {
  static constexpr int N = 16 * N2, NB = N / 2 + 1;
  union {
    float xw[kDF][N];                                   // windowed frames                         (stage 0-1)
    struct {
      float pw[NB][kDF + 1];                            // power spectrum                          (stage 2 - epilogue)
      float ot[kDF][129];                               // dB tile (nmels <= 128)
    } e;
  } u;
  float ar[kDF][N2][16], ai[kDF][N2][16];               // A' = stage-1 output times the inter-stage twiddle
  float twc[N2][16], tws[N2][16];                       // cos / sin of 2*pi*n2*k1/N
  float wmax[kThreads / 32];
};

// stage 2 for the k2 range [K0, K1) of one (frame, k1): N2-point DFT with compile-time twiddles
template <int N2, int K0, int K1>
__device__ __forceinline__ void stage2(const float* __restrict__ ar, const float* __restrict__ ai, float (*pw)[kDF + 1], const int f,
                                       const int k1) {
  constexpr int NK = K1 - K0, NB = 16 * N2 / 2 + 1;
  float xr[NK], xi[NK];
#pragma unroll
  for (int q = 0; q < NK; ++q) { xr[q] = 0.f; xi[q] = 0.f; }
#pragma unroll
  for (int n2 = 0; n2 < N2; ++n2) {
    const float a = ar[n2 * 16], b = ai[n2 * 16];
#pragma unroll
    for (int q = 0; q < NK; ++q) {
      const int idx = (n2 * (K0 + q)) % N2;             // compile-time after unrolling
      const float c = Tw<N2>::c(idx), s = Tw<N2>::s(idx);
      // (a + i b) * (c - i s)
      xr[q] = fmaf(a, c, xr[q]); xr[q] = fmaf(b, s, xr[q]);
      xi[q] = fmaf(b, c, xi[q]); xi[q] = fmaf(-a, s, xi[q]);
    }
  }
#pragma unroll
  for (int q = 0; q < NK; ++q) {
    const int k = k1 + 16 * (K0 + q);
    if (k < NB) pw[k][f] = fmaf(xr[q], xr[q], xi[q] * xi[q]);
  }
}

template <int N2>
__global__ void __launch_bounds__(kThreads) fbank_dft_kernel(const cm_fbank_wav_args p) {
  using S = Smem<N2>;
  constexpr int N = S::N, NB = S::NB;
  constexpr int K2 = N2 / 2 + 1;                         // k2 = 0 .. N2/2 covers every bin k <= N/2
  constexpr int K2a = (K2 + 1) / 2;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  S& sm = *reinterpret_cast<S*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * kDF;

  // inter-stage twiddle table (once per CTA)
  for (int i = tid; i < N2 * 16; i += kThreads) {
    const int n2 = i >> 4, k1 = i & 15;
    float s, c;
    sincospif(2.0f * (float)((n2 * k1) % N) / (float)N, &s, &c);
    sm.twc[n2][k1] = c; sm.tws[n2][k1] = s;
  }
  // stage 0: windowed frames (centre padding with zeros: frame t covers samples [t*hop - N/2, t*hop + N/2))
  const float* wv = p.wav + (int64_t)b * p.wav_sb;
  for (int i = tid; i < kDF * N; i += kThreads) {
    const int f = i / N, n = i - f * N;
    const int t = t0 + f;
    const int s = t * p.hop - N / 2 + n;
    float v = 0.f;
    if (t < p.frames && s >= 0 && s < p.n_samples) v = __ldg(wv + s) * __ldg(p.window + n);
    sm.u.xw[f][n] = v;
  }
  __syncthreads();

  // stage 1: thread = (frame, n2)
  if (tid < kDF * N2) {
    const int f = tid / N2, n2 = tid - f * N2;
    float x[16];
#pragma unroll
    for (int n1 = 0; n1 < 16; ++n1) x[n1] = sm.u.xw[f][N2 * n1 + n2];
    float re[9], im[9];
#pragma unroll
    for (int k1 = 0; k1 <= 8; ++k1) {
      float r = 0.f, m = 0.f;
#pragma unroll
      for (int n1 = 0; n1 < 16; ++n1) {
        const int idx = (n1 * k1) & 15;
        r = fmaf(x[n1], kCos16[idx], r);
        m = fmaf(-x[n1], kSin16[idx], m);
      }
      re[k1] = r; im[k1] = m;
    }
#pragma unroll
    for (int k1 = 0; k1 < 16; ++k1) {
      const float r = k1 <= 8 ? re[k1] : re[16 - k1];
      const float m = k1 <= 8 ? im[k1] : -im[16 - k1];   // real input: A[16 - k1] = conj(A[k1])
      const float c = sm.twc[n2][k1], s = sm.tws[n2][k1];
      sm.ar[f][n2][k1] = fmaf(r, c, m * s);              // (r + i m) * (c - i s)
      sm.ai[f][n2][k1] = fmaf(m, c, -r * s);
    }
  }
  __syncthreads();

  // stage 2: thread = (frame, k1, half of the k2 range); the halves are warp-uniform
  {
    const int f = (tid & 127) >> 4, k1 = tid & 15;
    const float* ar = &sm.ar[f][0][k1];
    const float* ai = &sm.ai[f][0][k1];
    if (tid < 128) stage2<N2, 0, K2a>(ar, ai, sm.u.e.pw, f, k1);
    else stage2<N2, K2a, K2>(ar, ai, sm.u.e.pw, f, k1);
  }
  __syncthreads();

  // mel projection over each filter's support, dB, running max.  thread -> (frame = tid & 7, mel = (tid >> 3) + 32 i)
  float tmax = -CUDART_INF_F;
  {
    const int f = tid & (kDF - 1);
    const float k10 = p.multiplier * 0.30102999566398120f;   // multiplier * log10(2)
    for (int m = tid >> 3; m < p.nmels; m += kThreads / kDF) {
      const int lo = __ldg(p.band + 2 * m), hi = __ldg(p.band + 2 * m + 1);
      float acc = 0.f;
      for (int k = lo; k < hi; ++k) acc = fmaf(sm.u.e.pw[k][f], __ldg(p.fbank + (int64_t)k * p.nmels + m), acc);
      const float db = k10 * lg2(fmaxf(acc, p.amin)) - p.db_offset;
      sm.u.e.ot[f][m] = db;
      if (t0 + f < p.frames) tmax = fmaxf(tmax, db);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, o));
  if (lane == 0) sm.wmax[warp] = tmax;
  __syncthreads();

  const int nfr = min(kDF, p.frames - t0);
  float* outp = p.out + ((int64_t)b * p.frames + t0) * p.nmels;
  for (int i = tid; i < nfr * p.nmels; i += kThreads) {
    const int fr = i / p.nmels, m = i - fr * p.nmels;
    outp[i] = sm.u.e.ot[fr][m];
  }
  if (tid == 0) {
    float mx = sm.wmax[0];
#pragma unroll
    for (int w = 1; w < kThreads / 32; ++w) mx = fmaxf(mx, sm.wmax[w]);
    int* addr = reinterpret_cast<int*>(p.utt_max + b);   // float max via ordered-int atomics (valid for mixed signs)
    if (mx >= 0.f) atomicMax(addr, __float_as_int(mx));
    else atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(mx));
  }
}

template <int N2>
static int launch(const cm_fbank_wav_args& a, cudaStream_t st) {
  const size_t smem = sizeof(Smem<N2>);
  auto kern = fbank_dft_kernel<N2>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);   // per device
  if (e != cudaSuccess) return (int)e;
  kern<<<dim3(cm_ceil_div(a.frames, kDF), a.batch), kThreads, smem, st>>>(a);
  CM_LAUNCH_CHECK();
  return 0;
}

}  // namespace dft
}  // namespace cm

extern "C" int cm_fbank_wav_supported(int32_t n_fft) { return (n_fft == 400 || n_fft == 512) ? 1 : 0; }

extern "C" int cm_fbank_wav_logmel(const cm_fbank_wav_args* a, void* stream) {
  if (a == nullptr) return CM_ERR_BAD_ARG;
  if (a->batch <= 0 || a->frames <= 0 || a->n_samples <= 0 || a->nmels <= 0 || a->hop <= 0) return CM_ERR_BAD_ARG;
  if (!a->wav || !a->window || !a->fbank || !a->band || !a->out || !a->utt_max) return CM_ERR_BAD_ARG;
  if (a->batch > 65535 || a->nmels > 128) return CM_ERR_UNSUPPORTED;
  if ((int64_t)a->frames * a->hop > (int64_t)INT32_MAX / 2) return CM_ERR_UNSUPPORTED;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (a->n_fft == 400) return cm::dft::launch<25>(*a, st);
  if (a->n_fft == 512) return cm::dft::launch<32>(*a, st);
  return CM_ERR_UNSUPPORTED;
}
