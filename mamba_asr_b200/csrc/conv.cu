// Depthwise causal conv1d (+SiLU) forward / backward for sm_100a, both BiMamba-v2 directions from one read.
//
// Replaces causal_conv1d_cuda.causal_conv1d_fwd / _bwd (reference call sites
// modules/mamba/selective_scan_interface.py:182,244,286 and modules/mamba/bimamba.py:282-287).
// Widths 2..4 are evaluated as 4 taps with left-zero-padded weights w4[j] = w[j-(4-W)]:
//   causal      s_f[l] = bias_f + sum_j w4_f[j] * x[l-3+j]
//   anticausal  s_b[l] = bias_b + sum_j w4_b[j] * x[l+3-j]      (conv1d_b on the flipped sequence, in place)
//
// Channel-last kernels (x.sd == 1, the layout the B200 Mamba block uses): one thread owns VEC adjacent channels
// and 16 consecutive time steps; all 22 row loads (16 + 3 halo each side) are issued before the first FMA, every
// access is a 64..256-byte coalesced row segment, HBM traffic is x once in, each output once out.
// Generic kernels (any strides, e.g. the reference's time-contiguous (B, D, L)): lanes along time, halo re-reads
// served by L1.  Weight/bias gradients are reduced in-CTA and written as per-(batch, 64-step chunk) partial rows
// that cm_reduce_rows() sums in a fixed order (no atomics, deterministic).
#include <algorithm>
#include <cstdlib>
#include <type_traits>

#include "common.cuh"

namespace cm {

#ifndef CM_CONV_TL
#define CM_CONV_TL 16
#endif
#ifndef CM_CONV_TY
#define CM_CONV_TY 4
#endif
constexpr int kTL = CM_CONV_TL;   // time steps per thread (channel-last kernels)
constexpr int kTY = CM_CONV_TY;   // threadIdx.y: time chunks per CTA  -> kTL * kTY = 64 steps per CTA
constexpr int kChunk = kTL * kTY;

template <int BYTES> struct RawVec;
template <> struct RawVec<2> { using type = unsigned short; };
template <> struct RawVec<4> { using type = uint32_t; };
template <> struct RawVec<8> { using type = uint2; };
template <> struct RawVec<16> { using type = uint4; };

template <typename T> __device__ __forceinline__ float bits_to_float(uint32_t lo16);
template <> __device__ __forceinline__ float bits_to_float<__nv_bfloat16>(uint32_t v) { return __uint_as_float(v << 16); }
template <> __device__ __forceinline__ float bits_to_float<__half>(uint32_t v) {
  return __half2float(__ushort_as_half(static_cast<unsigned short>(v)));
}
template <typename T> __device__ __forceinline__ uint32_t float_to_bits(float f);
template <> __device__ __forceinline__ uint32_t float_to_bits<__nv_bfloat16>(float f) {
  return __bfloat16_as_ushort(__float2bfloat16_rn(f));
}
template <> __device__ __forceinline__ uint32_t float_to_bits<__half>(float f) {
  return __half_as_ushort(__float2half_rn(f));
}

// VEC adjacent elements <-> fp32 registers through one aligned load/store.  ld_raw / cvt are split so that a kernel can
// issue all of a tile's loads back to back and unpack afterwards (an unpack right behind each load serialises them).
template <typename T, int VEC>
struct VecIO {
  using Raw = typename RawVec<sizeof(T) * VEC>::type;
  static __device__ __forceinline__ Raw ld_raw(const T* p) { return __ldg(reinterpret_cast<const Raw*>(p)); }
  static __device__ __forceinline__ Raw zero() {
    Raw r;
    memset(&r, 0, sizeof(Raw));
    return r;
  }
  static __device__ __forceinline__ void cvt(const Raw& r, float (&o)[VEC]) {
    if constexpr (sizeof(T) == 4) {
      const float* f = reinterpret_cast<const float*>(&r);
#pragma unroll
      for (int i = 0; i < VEC; ++i) o[i] = f[i];
    } else {
      const unsigned short* h = reinterpret_cast<const unsigned short*>(&r);
#pragma unroll
      for (int i = 0; i < VEC; ++i) o[i] = bits_to_float<T>(h[i]);
    }
  }
  static __device__ __forceinline__ void ld(const T* p, float (&o)[VEC]) { cvt(ld_raw(p), o); }
  static __device__ __forceinline__ void st(T* p, const float (&v)[VEC]) {
    Raw r;
    if constexpr (sizeof(T) == 4) {
      float* f = reinterpret_cast<float*>(&r);
#pragma unroll
      for (int i = 0; i < VEC; ++i) f[i] = v[i];
    } else {
      unsigned short* h = reinterpret_cast<unsigned short*>(&r);
#pragma unroll
      for (int i = 0; i < VEC; ++i) h[i] = static_cast<unsigned short>(float_to_bits<T>(v[i]));
    }
    *reinterpret_cast<Raw*>(p) = r;
  }
};

// PRECISE (fp32 I/O): ex2 + rcp sigmoid; 16-bit I/O: one MUFU.TANH (error far below the output rounding)
template <bool PRECISE>
__device__ __forceinline__ float silu_f(float s) { return s * sigmoid_sel<PRECISE>(s); }
template <bool PRECISE>
__device__ __forceinline__ float silu_grad(float s) {
  const float sig = sigmoid_sel<PRECISE>(s);
  return sig * fmaf(s, 1.f - sig, 1.f);
}

// 4-tap weights of VEC channels, left-zero-padded
template <int VEC>
struct Taps {
  float w[VEC][4], b[VEC];
  __device__ __forceinline__ void load(const cm_conv_dir& dr, int d0, int W) {
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int k = j - (4 - W);
        w[v][j] = (k >= 0) ? __ldg(dr.weight + (int64_t)(d0 + v) * W + k) : 0.f;
      }
      b[v] = dr.bias ? __ldg(dr.bias + d0 + v) : 0.f;
    }
  }
};

// ---------------------------------------------------------------------------------------------------
// channel-last forward
// ---------------------------------------------------------------------------------------------------
template <typename T, int VEC>
__global__ void __launch_bounds__(32 * kTY) conv_fwd_cl_kernel(const cm_conv_args p) {
  const int d0 = (blockIdx.x * 32 + threadIdx.x) * VEC;
  const int l0 = (blockIdx.y * kTY + threadIdx.y) * kTL;
  const int L = p.seqlen;
  if (d0 >= p.dim || l0 >= L) return;
  const int b = blockIdx.z;
  const bool silu = (p.flags & CM_FLAG_SILU) != 0;

  using VIO = VecIO<T, VEC>;
  typename VIO::Raw xr[kTL + 6];
  const T* xp = static_cast<const T*>(p.x.ptr) + b * p.x.sb + d0;
#pragma unroll
  for (int i = 0; i < kTL + 6; ++i) {          // all 22 row loads in flight before the first unpack
    const int l = l0 - 3 + i;
    xr[i] = (l >= 0 && l < L) ? VIO::ld_raw(xp + (int64_t)l * p.x.sl) : VIO::zero();
  }
  float xw[kTL + 6][VEC];
#pragma unroll
  for (int i = 0; i < kTL + 6; ++i) VIO::cvt(xr[i], xw[i]);
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    if (r < p.ndir) {
      const cm_conv_dir& dr = p.dir[r];
      Taps<VEC> tp;
      tp.load(dr, d0, p.width);
      const bool anti = dr.anticausal != 0;
      T* op = static_cast<T*>(dr.out.ptr) + b * dr.out.sb + d0;
#pragma unroll
      for (int i = 0; i < kTL; ++i) {
        const int l = l0 + i;
        if (l < L) {
          float o[VEC];
#pragma unroll
          for (int v = 0; v < VEC; ++v) {
            float acc = tp.b[v];
#pragma unroll
            for (int j = 0; j < 4; ++j) acc = fmaf(tp.w[v][j], anti ? xw[i + 6 - j][v] : xw[i + j][v], acc);
            o[v] = silu ? silu_f<sizeof(T) == 4>(acc) : acc;
          }
          VecIO<T, VEC>::st(op + (int64_t)l * dr.out.sl, o);
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// channel-last backward: dx (summed over directions) + per-CTA partial dweight / dbias
// ---------------------------------------------------------------------------------------------------
template <typename T, int VEC>
__global__ void __launch_bounds__(32 * kTY) conv_bwd_cl_kernel(const cm_conv_args p) {
  __shared__ float red[kTY][32][VEC * 5 + 1];
  const int d0 = (blockIdx.x * 32 + threadIdx.x) * VEC;
  const int l0 = (blockIdx.y * kTY + threadIdx.y) * kTL;
  const int L = p.seqlen, W = p.width;
  const bool active = (d0 < p.dim) && (l0 < L);
  const int b = blockIdx.z;
  const bool silu = (p.flags & CM_FLAG_SILU) != 0;
  const int dsafe = (d0 < p.dim) ? d0 : 0;

  using VIO = VecIO<T, VEC>;
  // every global load of the thread is issued first: x window (22 rows) and both upstream-gradient windows (19 rows)
  typename VIO::Raw xr[kTL + 6], gr[2][kTL + 6];
#pragma unroll
  for (int i = 0; i < kTL + 6; ++i) { xr[i] = VIO::zero(); gr[0][i] = VIO::zero(); gr[1][i] = VIO::zero(); }
  if (active) {
    const T* xp = static_cast<const T*>(p.x.ptr) + b * p.x.sb + d0;
#pragma unroll
    for (int i = 0; i < kTL + 6; ++i) {
      const int l = l0 - 3 + i;
      if (l >= 0 && l < L) xr[i] = VIO::ld_raw(xp + (int64_t)l * p.x.sl);
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      if (r < p.ndir) {
        const cm_conv_dir& dr = p.dir[r];
        const bool anti = dr.anticausal != 0;
        const T* gp = static_cast<const T*>(dr.out.ptr) + b * dr.out.sb + d0;
#pragma unroll
        for (int q = 0; q < kTL + 6; ++q) {
          const bool need = anti ? (q < kTL + 3) : (q >= 3);
          const int lq = l0 - 3 + q;
          if (need && lq >= 0 && lq < L) gr[r][q] = VIO::ld_raw(gp + (int64_t)lq * dr.out.sl);
        }
      }
    }
  }
  float xw[kTL + 6][VEC];
  float dxa[kTL][VEC];
#pragma unroll
  for (int i = 0; i < kTL + 6; ++i) VIO::cvt(xr[i], xw[i]);
#pragma unroll
  for (int i = 0; i < kTL; ++i)
#pragma unroll
    for (int v = 0; v < VEC; ++v) dxa[i][v] = 0.f;

#pragma unroll
  for (int r = 0; r < 2; ++r) {
    if (r < p.ndir) {   // CTA-uniform
      const cm_conv_dir& dr = p.dir[r];
      const bool anti = dr.anticausal != 0;
      Taps<VEC> tp;
      tp.load(dr, dsafe, W);
      float dw[VEC][4], db[VEC];
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        db[v] = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) dw[v][j] = 0.f;
      }
      if (active) {
        // positions lq = l0 - 3 + q whose upstream gradient reaches this thread's dx rows:
        //   causal:     l' in [l0, l0 + TL + 3)   -> q in [3, TL + 6)
        //   anticausal: l' in [l0 - 3, l0 + TL)   -> q in [0, TL + 3)
#pragma unroll
        for (int q = 0; q < kTL + 6; ++q) {
          const bool need = anti ? (q < kTL + 3) : (q >= 3);
          const int lq = l0 - 3 + q;
          if (need && lq >= 0 && lq < L) {
            float go[VEC];
            VIO::cvt(gr[r][q], go);
            const bool own = (q >= 3) && (q < kTL + 3);   // lq in [l0, l0 + TL)
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
              float g = go[v];
              if (silu) {
                float s = tp.b[v];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  // causal tap j reads x[lq-3+j] = xw[q-3+j+... ] ; window index of time t is t-(l0-3)
                  const int xi = anti ? (q + 3 - j) : (q - 3 + j);
                  const float xv = (xi >= 0 && xi < kTL + 6) ? xw[xi][v] : 0.f;
                  s = fmaf(tp.w[v][j], xv, s);
                }
                g *= silu_grad<sizeof(T) == 4>(s);
              }
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int xi = anti ? (q + 3 - j) : (q - 3 + j);   // window index of the input this tap touches
                const int di = xi - 3;                            // dx row (time l0 + di)
                if (di >= 0 && di < kTL) dxa[di][v] = fmaf(tp.w[v][j], g, dxa[di][v]);
                if (own) {
                  const float xv = (xi >= 0 && xi < kTL + 6) ? xw[xi][v] : 0.f;
                  dw[v][j] = fmaf(g, xv, dw[v][j]);
                }
              }
              if (own) db[v] += g;
            }
          }
        }
      }
      // reduce the kTY time chunks of this CTA, one partial row per (batch, 64-step chunk)
      __syncthreads();
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
#pragma unroll
        for (int j = 0; j < 4; ++j) red[threadIdx.y][threadIdx.x][v * 5 + j] = dw[v][j];
        red[threadIdx.y][threadIdx.x][v * 5 + 4] = db[v];
      }
      __syncthreads();
      if (threadIdx.y == 0 && d0 < p.dim) {
        const int64_t prow = (int64_t)b * gridDim.y + blockIdx.y;
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float acc = 0.f;
#pragma unroll
            for (int y = 0; y < kTY; ++y) acc += red[y][threadIdx.x][v * 5 + j];
            const int k = j - (4 - W);
            if (k >= 0) dr.dweight_part[(prow * p.dim + d0 + v) * W + k] = acc;
          }
          if (dr.dbias_part) {
            float acc = 0.f;
#pragma unroll
            for (int y = 0; y < kTY; ++y) acc += red[y][threadIdx.x][v * 5 + 4];
            dr.dbias_part[prow * p.dim + d0 + v] = acc;
          }
        }
      }
    }
  }
  if (active) {
    T* dxp = static_cast<T*>(p.dx.ptr) + b * p.dx.sb + d0;
#pragma unroll
    for (int i = 0; i < kTL; ++i) {
      const int l = l0 + i;
      if (l < L) VecIO<T, VEC>::st(dxp + (int64_t)l * p.dx.sl, dxa[i]);
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// channel-last backward, sliding-window form (round 2): a thread owns TWO adjacent channels and a whole 64-step chunk
// ---------------------------------------------------------------------------------------------------
// conv_bwd_cl_kernel above scatters every upstream-gradient row into a 16-step register window of dx: 112 instructions per
// (channel, step) with both directions, 168 registers, 3 CTAs per SM (profiles/r02_conv_bwd_cfg3_shipped_ncu.txt: issue
// bound at 17 % occupancy, 0.14 of the HBM peak).  Here the thread walks its chunk in time order and keeps 4-deep windows
// in registers (the loop is unrolled by 4, so the rotation is register renaming):
//   X  = x[l-3 .. l]                  both pre-activations of a step come from the same window:
//        s_f[l]   = b_f + sum_j w_f[j] x[l-3+j]          s_b[l-3] = b_b + sum_j w_b[j] x[l-j]
//   GF = ge_f[l-3 .. l]               effective gradients  ge = g * silu'(s)
//   GB = ge_b[l-6 .. l-3]             (the anticausal direction runs three steps behind: its s needs x up to l)
//   dx[l-3] = sum_j w_f[j] GF[l-j] + sum_j w_b[j] GB[l-6+j]
// dweight / dbias accumulate in registers over the chunk and are written as the chunk's partial row directly (the thread
// owns the whole chunk: no shared-memory reduction).  State pairs are packed (FFMA2 on the two channels).  Per chunk a
// thread reads 70 rows of x and 67 / 70 of the two gradients for 64 rows of dx: 5-9 % halo instead of 37 %.
template <typename T> struct PairLd;      // two adjacent elements -> float2 (ld_raw / cvt split: loads are issued groups ahead)
template <> struct PairLd<float> {
  using Raw = float2;
  static __device__ __forceinline__ Raw ld_raw(const float* p) { return __ldg(reinterpret_cast<const float2*>(p)); }
  static __device__ __forceinline__ Raw zero() { return make_float2(0.f, 0.f); }
  static __device__ __forceinline__ float2 cvt(Raw r) { return r; }
  static __device__ __forceinline__ float2 ld(const float* p) { return __ldg(reinterpret_cast<const float2*>(p)); }
  static __device__ __forceinline__ void st(float* p, float2 v) { *reinterpret_cast<float2*>(p) = v; }
};
template <> struct PairLd<__nv_bfloat16> {
  using Raw = uint32_t;
  static __device__ __forceinline__ Raw ld_raw(const __nv_bfloat16* p) { return __ldg(reinterpret_cast<const uint32_t*>(p)); }
  static __device__ __forceinline__ Raw zero() { return 0u; }
  static __device__ __forceinline__ float2 cvt(Raw r) { return make_float2(__uint_as_float(r << 16), __uint_as_float(r & 0xffff0000u)); }
  static __device__ __forceinline__ float2 ld(const __nv_bfloat16* p) { return cvt(ld_raw(p)); }
  static __device__ __forceinline__ void st(__nv_bfloat16* p, float2 v) {
    *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(v.x, v.y);
  }
};
template <> struct PairLd<__half> {
  using Raw = uint32_t;
  static __device__ __forceinline__ Raw ld_raw(const __half* p) { return __ldg(reinterpret_cast<const uint32_t*>(p)); }
  static __device__ __forceinline__ Raw zero() { return 0u; }
  static __device__ __forceinline__ float2 cvt(Raw r) { return __half22float2(*reinterpret_cast<const __half2*>(&r)); }
  static __device__ __forceinline__ float2 ld(const __half* p) { return cvt(ld_raw(p)); }
  static __device__ __forceinline__ void st(__half* p, float2 v) { *reinterpret_cast<__half2*>(p) = __floats2half2_rn(v.x, v.y); }
};

template <bool PRECISE>
__device__ __forceinline__ float2 silu_grad2(float2 s) {
  return make_float2(silu_grad<PRECISE>(s.x), silu_grad<PRECISE>(s.y));
}

// M0 / M1: mode of direction slot 0 / 1: 0 causal, 1 anticausal, -1 absent
#ifndef CM_CONV_SW_MINB
#define CM_CONV_SW_MINB 3
#endif
template <typename T, bool SILU, int M0, int M1>
__global__ void __launch_bounds__(128, CM_CONV_SW_MINB) conv_bwd_sw_kernel(const cm_conv_args p) {
  constexpr bool PRECISE = sizeof(T) == 4;
  constexpr int NS = (M0 >= 0 ? 1 : 0) + (M1 >= 0 ? 1 : 0);
  const int d0 = (blockIdx.x * 128 + threadIdx.x) * 2;
  if (d0 >= p.dim) return;
  const int L = p.seqlen, W = p.width;
  const int l0 = blockIdx.y * kChunk, l1 = min(l0 + kChunk, L);
  const int b = blockIdx.z;
  const T* xp = static_cast<const T*>(p.x.ptr) + b * p.x.sb + d0;
  T* dxp = static_cast<T*>(p.dx.ptr) + b * p.dx.sb + d0;

  // per direction slot: taps (pairs over the two channels), upstream-gradient pointer, accumulators
  float2 w[2][4], bias[2], dw[2][4], db[2];
  const T* gp[2];
  int64_t gsl[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int mode = r == 0 ? M0 : M1;
    if (mode < 0) continue;
    const cm_conv_dir& dr = p.dir[r];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = j - (4 - W);
      w[r][j] = (k >= 0) ? make_float2(__ldg(dr.weight + (int64_t)d0 * W + k), __ldg(dr.weight + (int64_t)(d0 + 1) * W + k))
                         : make_float2(0.f, 0.f);
      dw[r][j] = make_float2(0.f, 0.f);
    }
    bias[r] = dr.bias ? make_float2(__ldg(dr.bias + d0), __ldg(dr.bias + d0 + 1)) : make_float2(0.f, 0.f);
    db[r] = make_float2(0.f, 0.f);
    gp[r] = static_cast<const T*>(dr.out.ptr) + b * dr.out.sb + d0;
    gsl[r] = dr.out.sl;
  }
  const float2 zero2 = make_float2(0.f, 0.f);
  auto ldx = [&](int l) { return (l >= 0 && l < L) ? PairLd<T>::ld(xp + (int64_t)l * p.x.sl) : zero2; };
  auto ldg = [&](int r, int l) { return (l >= 0 && l < L) ? PairLd<T>::ld(gp[r] + (int64_t)l * gsl[r]) : zero2; };

  // windows, slot = step & 3
  float2 X[4], GE[2][4], GD[2][4];          // GD: raw upstream gradient of an anticausal slot, delayed by three steps
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    X[i] = zero2;
#pragma unroll
    for (int r = 0; r < 2; ++r) { GE[r][i] = zero2; GD[r][i] = zero2; }
  }
  // steps l0-3 .. l0-1 only fill the windows: x, and the delayed gradients of the anticausal slots
#pragma unroll
  for (int i = 1; i <= 3; ++i) {
    X[i] = ldx(l0 - 4 + i);                  // slot of step l is (l - l0) & 3: steps l0-3, l0-2, l0-1 -> slots 1, 2, 3
    if (M0 == 1) GD[0][i] = ldg(0, l0 - 4 + i);
    if (M1 == 1) GD[1][i] = ldg(1, l0 - 4 + i);
  }

  // one step.  PH = (l - l0) & 3 is a compile-time constant inside the unrolled body
  auto step = [&](auto ph_c, const int l, const float2 xv, const float2 g0v, const float2 g1v) {
    constexpr int PH = decltype(ph_c)::value;
    constexpr int S0 = PH, S1 = (PH + 3) & 3, S2 = (PH + 2) & 3, S3 = (PH + 1) & 3;   // slots of steps l, l-1, l-2, l-3
    X[S0] = xv;
    float2 dxv = zero2;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int mode = r == 0 ? M0 : M1;
      if (mode < 0) continue;
      if (mode == 0) {
        // causal: s_f[l] = b + w0 x[l-3] + w1 x[l-2] + w2 x[l-1] + w3 x[l];  ge_f[l] joins the window
        float2 ge = r == 0 ? g0v : g1v;
        if (SILU) {
          float2 s = ffma2(w[r][0], X[S3], bias[r]);
          s = ffma2(w[r][1], X[S2], s); s = ffma2(w[r][2], X[S1], s); s = ffma2(w[r][3], X[S0], s);
          ge = fmul2(ge, silu_grad2<PRECISE>(s));
        }
        GE[r][S0] = ge;
        // dx[l-3] += w0 ge[l-3] + w1 ge[l-2] + w2 ge[l-1] + w3 ge[l]  ... with tap j on ge[l-3+(3-j)] = ge[l-j]
        dxv = ffma2(w[r][0], GE[r][S0], dxv); dxv = ffma2(w[r][1], GE[r][S1], dxv);
        dxv = ffma2(w[r][2], GE[r][S2], dxv); dxv = ffma2(w[r][3], GE[r][S3], dxv);
        const float2 go = (l < l1) ? ge : zero2;          // every step is owned by exactly one chunk
        dw[r][0] = ffma2(go, X[S3], dw[r][0]); dw[r][1] = ffma2(go, X[S2], dw[r][1]);
        dw[r][2] = ffma2(go, X[S1], dw[r][2]); dw[r][3] = ffma2(go, X[S0], dw[r][3]);
        db[r] = fadd2(db[r], go);
      } else {
        // anticausal, three steps behind: s_b[l-3] = b + w0 x[l] + w1 x[l-1] + w2 x[l-2] + w3 x[l-3]
        float2 ge = GD[r][S3];                             // upstream gradient of step l-3
        GD[r][S0] = r == 0 ? g0v : g1v;
        if (SILU) {
          float2 s = ffma2(w[r][0], X[S0], bias[r]);
          s = ffma2(w[r][1], X[S1], s); s = ffma2(w[r][2], X[S2], s); s = ffma2(w[r][3], X[S3], s);
          ge = fmul2(ge, silu_grad2<PRECISE>(s));
        }
        // window of effective gradients: slot S0 <- ge_b[l-3]; S1, S2, S3 hold ge_b[l-4], ge_b[l-5], ge_b[l-6]
        GE[r][S0] = ge;
        // dx[l-3] += sum_j w[j] ge_b[l-6+j]
        dxv = ffma2(w[r][0], GE[r][S3], dxv); dxv = ffma2(w[r][1], GE[r][S2], dxv);
        dxv = ffma2(w[r][2], GE[r][S1], dxv); dxv = ffma2(w[r][3], GE[r][S0], dxv);
        const float2 go = (l - 3 >= l0 && l - 3 < l1) ? ge : zero2;
        dw[r][0] = ffma2(go, X[S0], dw[r][0]); dw[r][1] = ffma2(go, X[S1], dw[r][1]);
        dw[r][2] = ffma2(go, X[S2], dw[r][2]); dw[r][3] = ffma2(go, X[S3], dw[r][3]);
        db[r] = fadd2(db[r], go);
      }
    }
    if (l - 3 >= l0 && l - 3 < l1) PairLd<T>::st(dxp + (int64_t)(l - 3) * p.dx.sl, dxv);
  };
  // steps l0 .. l1+2 (the last three only flush the windows), in groups of 4 so that the window slots are compile-time.
  // The rows of a group are loaded TWO GROUPS AHEAD into raw registers (12 loads in flight per thread at any time): the
  // first version loaded each row where it was used and ran at the speed of one DRAM round trip per step.
  using PL = PairLd<T>;
  typename PL::Raw rx[2][4], rg0[2][4], rg1[2][4];
  auto load_group = [&](auto buf_c, const int lb) {
    constexpr int B = decltype(buf_c)::value;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int l = lb + i;
      const bool in = l < L;                 // l >= l0 >= 0
      rx[B][i] = in ? PL::ld_raw(xp + (int64_t)l * p.x.sl) : PL::zero();
      rg0[B][i] = (M0 >= 0 && in) ? PL::ld_raw(gp[0] + (int64_t)l * gsl[0]) : PL::zero();
      rg1[B][i] = (M1 >= 0 && in) ? PL::ld_raw(gp[1] + (int64_t)l * gsl[1]) : PL::zero();
    }
  };
  auto run_group = [&](auto buf_c, const int lb) {
    constexpr int B = decltype(buf_c)::value;
    float2 xv[4], g0v[4], g1v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { xv[i] = PL::cvt(rx[B][i]); g0v[i] = PL::cvt(rg0[B][i]); g1v[i] = PL::cvt(rg1[B][i]); }
    load_group(buf_c, lb + 8);               // refill the buffer just unpacked: the rows of the group after next
    step(std::integral_constant<int, 0>{}, lb, xv[0], g0v[0], g1v[0]);
    step(std::integral_constant<int, 1>{}, lb + 1, xv[1], g0v[1], g1v[1]);
    step(std::integral_constant<int, 2>{}, lb + 2, xv[2], g0v[2], g1v[2]);
    step(std::integral_constant<int, 3>{}, lb + 3, xv[3], g0v[3], g1v[3]);
  };
  const int lend = l1 + 3;
  load_group(std::integral_constant<int, 0>{}, l0);
  load_group(std::integral_constant<int, 1>{}, l0 + 4);
#pragma unroll 1
  for (int l = l0; l < lend; l += 8) {
    run_group(std::integral_constant<int, 0>{}, l);
    run_group(std::integral_constant<int, 1>{}, l + 4);
  }

  const int64_t prow = (int64_t)b * gridDim.y + blockIdx.y;
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int mode = r == 0 ? M0 : M1;
    if (mode < 0) continue;
    const cm_conv_dir& dr = p.dir[r];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = j - (4 - W);
      if (k >= 0) {
        dr.dweight_part[(prow * p.dim + d0) * W + k] = dw[r][j].x;
        dr.dweight_part[(prow * p.dim + d0 + 1) * W + k] = dw[r][j].y;
      }
    }
    if (dr.dbias_part) {
      dr.dbias_part[prow * p.dim + d0] = db[r].x;
      dr.dbias_part[prow * p.dim + d0 + 1] = db[r].y;
    }
  }
  (void)NS;
}

// Forward in the same sliding-window form: a thread owns two adjacent channels and a 64-step chunk; both directions' outputs of
// a step come from one 4-deep window of x (the anticausal output of step l-3 needs x up to l, so it is produced three steps
// late).  ~15 instructions per (channel, step) against 45 of conv_fwd_cl_kernel, rows loaded two groups ahead.
template <typename T, bool SILU, int M0, int M1>
__global__ void __launch_bounds__(128) conv_fwd_sw_kernel(const cm_conv_args p) {
  constexpr bool PRECISE = sizeof(T) == 4;
  const int d0 = (blockIdx.x * 128 + threadIdx.x) * 2;
  if (d0 >= p.dim) return;
  const int L = p.seqlen, W = p.width;
  const int l0 = blockIdx.y * kChunk, l1 = min(l0 + kChunk, L);
  const int b = blockIdx.z;
  const T* xp = static_cast<const T*>(p.x.ptr) + b * p.x.sb + d0;
  float2 w[2][4], bias[2];
  T* op[2];
  int64_t osl[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int mode = r == 0 ? M0 : M1;
    if (mode < 0) continue;
    const cm_conv_dir& dr = p.dir[r];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = j - (4 - W);
      w[r][j] = (k >= 0) ? make_float2(__ldg(dr.weight + (int64_t)d0 * W + k), __ldg(dr.weight + (int64_t)(d0 + 1) * W + k))
                         : make_float2(0.f, 0.f);
    }
    bias[r] = dr.bias ? make_float2(__ldg(dr.bias + d0), __ldg(dr.bias + d0 + 1)) : make_float2(0.f, 0.f);
    op[r] = static_cast<T*>(dr.out.ptr) + b * dr.out.sb + d0;
    osl[r] = dr.out.sl;
  }
  using PL = PairLd<T>;
  const float2 zero2 = make_float2(0.f, 0.f);
  float2 X[4];
  X[0] = zero2;
#pragma unroll
  for (int i = 1; i <= 3; ++i) {
    const int l = l0 - 4 + i;
    X[i] = (l >= 0) ? PL::ld(xp + (int64_t)l * p.x.sl) : zero2;
  }
  auto act = [&](float2 s) {
    return SILU ? make_float2(silu_f<PRECISE>(s.x), silu_f<PRECISE>(s.y)) : s;
  };
  auto step = [&](auto ph_c, const int l, const float2 xv) {
    constexpr int PH = decltype(ph_c)::value;
    constexpr int S0 = PH, S1 = (PH + 3) & 3, S2 = (PH + 2) & 3, S3 = (PH + 1) & 3;   // slots of steps l, l-1, l-2, l-3
    X[S0] = xv;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int mode = r == 0 ? M0 : M1;
      if (mode < 0) continue;
      if (mode == 0) {
        if (l < l1) {
          float2 s = ffma2(w[r][0], X[S3], bias[r]);
          s = ffma2(w[r][1], X[S2], s); s = ffma2(w[r][2], X[S1], s); s = ffma2(w[r][3], X[S0], s);
          PL::st(op[r] + (int64_t)l * osl[r], act(s));
        }
      } else {
        if (l - 3 >= l0 && l - 3 < l1) {
          float2 s = ffma2(w[r][0], X[S0], bias[r]);
          s = ffma2(w[r][1], X[S1], s); s = ffma2(w[r][2], X[S2], s); s = ffma2(w[r][3], X[S3], s);
          PL::st(op[r] + (int64_t)(l - 3) * osl[r], act(s));
        }
      }
    }
  };
  typename PL::Raw rx[2][4];
  auto load_group = [&](auto buf_c, const int lb) {
    constexpr int B = decltype(buf_c)::value;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int l = lb + i;
      rx[B][i] = (l < L) ? PL::ld_raw(xp + (int64_t)l * p.x.sl) : PL::zero();
    }
  };
  auto run_group = [&](auto buf_c, const int lb) {
    constexpr int B = decltype(buf_c)::value;
    float2 xv[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) xv[i] = PL::cvt(rx[B][i]);
    load_group(buf_c, lb + 8);
    step(std::integral_constant<int, 0>{}, lb, xv[0]);
    step(std::integral_constant<int, 1>{}, lb + 1, xv[1]);
    step(std::integral_constant<int, 2>{}, lb + 2, xv[2]);
    step(std::integral_constant<int, 3>{}, lb + 3, xv[3]);
  };
  const bool any_anti = (M0 == 1) || (M1 == 1);
  const int lend = any_anti ? l1 + 3 : l1;
  load_group(std::integral_constant<int, 0>{}, l0);
  load_group(std::integral_constant<int, 1>{}, l0 + 4);
#pragma unroll 1
  for (int l = l0; l < lend; l += 8) {
    run_group(std::integral_constant<int, 0>{}, l);
    run_group(std::integral_constant<int, 1>{}, l + 4);
  }
}

template <typename T>
static bool launch_conv_fwd_sw(const cm_conv_args& a, cudaStream_t st) {
  if (getenv("CM_CONV_NO_SW") != nullptr) return false;
  const dim3 grid(cm_ceil_div(a.dim / 2, 128), cm_ceil_div(a.seqlen, kChunk), a.batch);
  const bool silu = (a.flags & CM_FLAG_SILU) != 0;
  const int m0 = a.dir[0].anticausal ? 1 : 0, m1 = a.ndir == 2 ? (a.dir[1].anticausal ? 1 : 0) : -1;
#define CM_SWF(S, A0, A1) conv_fwd_sw_kernel<T, S, A0, A1><<<grid, 128, 0, st>>>(a)
  if (silu) {
    if (m0 == 0 && m1 == 1) CM_SWF(true, 0, 1);
    else if (m0 == 1 && m1 == 0) CM_SWF(true, 1, 0);
    else if (m0 == 0 && m1 == -1) CM_SWF(true, 0, -1);
    else if (m0 == 1 && m1 == -1) CM_SWF(true, 1, -1);
    else return false;
  } else {
    if (m0 == 0 && m1 == 1) CM_SWF(false, 0, 1);
    else if (m0 == 1 && m1 == 0) CM_SWF(false, 1, 0);
    else if (m0 == 0 && m1 == -1) CM_SWF(false, 0, -1);
    else if (m0 == 1 && m1 == -1) CM_SWF(false, 1, -1);
    else return false;
  }
#undef CM_SWF
  return true;
}

template <typename T>
static bool launch_conv_bwd_sw(const cm_conv_args& a, cudaStream_t st) {
  if (getenv("CM_CONV_NO_SW") != nullptr) return false;
  const dim3 grid(cm_ceil_div(a.dim / 2, 128), cm_ceil_div(a.seqlen, kChunk), a.batch);
  const bool silu = (a.flags & CM_FLAG_SILU) != 0;
  const int m0 = a.dir[0].anticausal ? 1 : 0, m1 = a.ndir == 2 ? (a.dir[1].anticausal ? 1 : 0) : -1;
#define CM_SW(S, A0, A1) conv_bwd_sw_kernel<T, S, A0, A1><<<grid, 128, 0, st>>>(a)
  if (silu) {
    if (m0 == 0 && m1 == 1) CM_SW(true, 0, 1);
    else if (m0 == 1 && m1 == 0) CM_SW(true, 1, 0);
    else if (m0 == 0 && m1 == -1) CM_SW(true, 0, -1);
    else if (m0 == 1 && m1 == -1) CM_SW(true, 1, -1);
    else return false;
  } else {
    if (m0 == 0 && m1 == 1) CM_SW(false, 0, 1);
    else if (m0 == 1 && m1 == 0) CM_SW(false, 1, 0);
    else if (m0 == 0 && m1 == -1) CM_SW(false, 0, -1);
    else if (m0 == 1 && m1 == -1) CM_SW(false, 1, -1);
    else return false;
  }
#undef CM_SW
  return true;
}

// ---------------------------------------------------------------------------------------------------
// generic-stride kernels: one thread per (b, d, l), lanes along time
// ---------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ float ldx(const T* base, int64_t sl, int l, int L) {
  return (l >= 0 && l < L) ? Elem<T>::ld(base + (int64_t)l * sl) : 0.f;
}

template <typename T>
__global__ void __launch_bounds__(128) conv_fwd_generic_kernel(const cm_conv_args p) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  const int d = blockIdx.y, b = blockIdx.z, L = p.seqlen, W = p.width;
  if (l >= L) return;
  const bool silu = (p.flags & CM_FLAG_SILU) != 0;
  const T* xp = static_cast<const T*>(p.x.ptr) + b * p.x.sb + d * p.x.sd;
  float xv[7];
#pragma unroll
  for (int i = 0; i < 7; ++i) xv[i] = ldx(xp, p.x.sl, l - 3 + i, L);
  for (int r = 0; r < p.ndir; ++r) {
    const cm_conv_dir& dr = p.dir[r];
    Taps<1> tp;
    tp.load(dr, d, W);
    float acc = tp.b[0];
#pragma unroll
    for (int j = 0; j < 4; ++j) acc = fmaf(tp.w[0][j], dr.anticausal ? xv[6 - j] : xv[j], acc);
    Elem<T>::st(static_cast<T*>(dr.out.ptr) + b * dr.out.sb + d * dr.out.sd + (int64_t)l * dr.out.sl,
                silu ? silu_f<sizeof(T) == 4>(acc) : acc);
  }
}

template <typename T>
__global__ void __launch_bounds__(kChunk) conv_bwd_generic_kernel(const cm_conv_args p) {
  __shared__ float red[kChunk / 32][5];
  const int l = blockIdx.x * kChunk + threadIdx.x;
  const int d = blockIdx.y, b = blockIdx.z, L = p.seqlen, W = p.width;
  const bool silu = (p.flags & CM_FLAG_SILU) != 0;
  const bool inb = l < L;
  const T* xp = static_cast<const T*>(p.x.ptr) + b * p.x.sb + d * p.x.sd;
  float xv[13];   // x[l-6 .. l+6]
#pragma unroll
  for (int i = 0; i < 13; ++i) xv[i] = inb ? ldx(xp, p.x.sl, l - 6 + i, L) : 0.f;
  float dx = 0.f;
  for (int r = 0; r < p.ndir; ++r) {
    const cm_conv_dir& dr = p.dir[r];
    const bool anti = dr.anticausal != 0;
    Taps<1> tp;
    tp.load(dr, d, W);
    const T* gp = static_cast<const T*>(dr.out.ptr) + b * dr.out.sb + d * dr.out.sd;
    float dw[4] = {0.f, 0.f, 0.f, 0.f}, db = 0.f;
    if (inb) {
      // upstream positions l' = l + o that touch x[l]:  causal o in [0,3] (tap j = 3-o), anticausal o in [-3,0] (tap j = 3+o)
#pragma unroll
      for (int oo = 0; oo < 4; ++oo) {
        const int o = anti ? -oo : oo;
        const int lp = l + o;
        if (lp >= 0 && lp < L) {
          float g = Elem<T>::ld(gp + (int64_t)lp * dr.out.sl);
          if (silu) {
            float s = tp.b[0];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int off = anti ? (o + 3 - j) : (o - 3 + j);   // input time relative to l
              s = fmaf(tp.w[0][j], xv[6 + off], s);
            }
            g *= silu_grad<sizeof(T) == 4>(s);
          }
          dx = fmaf(tp.w[0][3 - oo], g, dx);
          if (oo == 0) {   // own position: weight / bias gradients
            db = g;
#pragma unroll
            for (int j = 0; j < 4; ++j) dw[j] = g * xv[6 + (anti ? (3 - j) : (j - 3))];
          }
        }
      }
    }
    // CTA reduce of (dw[4], db)
    float vals[5] = {dw[0], dw[1], dw[2], dw[3], db};
#pragma unroll
    for (int i = 0; i < 5; ++i) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) vals[i] += __shfl_xor_sync(0xffffffffu, vals[i], o);
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
      for (int i = 0; i < 5; ++i) red[threadIdx.x >> 5][i] = vals[i];
    }
    __syncthreads();
    if (threadIdx.x < 5) {
      float acc = 0.f;
#pragma unroll
      for (int w = 0; w < kChunk / 32; ++w) acc += red[w][threadIdx.x];
      const int64_t prow = (int64_t)b * gridDim.x + blockIdx.x;
      if (threadIdx.x < 4) {
        const int k = threadIdx.x - (4 - W);
        if (k >= 0) dr.dweight_part[(prow * p.dim + d) * W + k] = acc;
      } else if (dr.dbias_part) {
        dr.dbias_part[prow * p.dim + d] = acc;
      }
    }
  }
  if (inb) Elem<T>::st(static_cast<T*>(p.dx.ptr) + b * p.dx.sb + d * p.dx.sd + (int64_t)l * p.dx.sl, dx);
}

// single-token update --------------------------------------------------------------------------------
template <typename T>
__global__ void conv_update_kernel(const T* __restrict__ x, T* __restrict__ state, const float* __restrict__ w,
                                   const float* __restrict__ bias, T* __restrict__ out, int batch, int dim, int W,
                                   uint32_t flags) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)batch * dim) return;
  const int d = idx % dim;
  T* st = state + idx * W;
  float acc = bias ? __ldg(bias + d) : 0.f;
  for (int k = 0; k < W; ++k) {
    const float v = (k + 1 < W) ? Elem<T>::round(static_cast<float>(st[k + 1])) : Elem<T>::ld(x + idx);
    Elem<T>::st(st + k, v);
    acc = fmaf(__ldg(w + (int64_t)d * W + k), v, acc);
  }
  Elem<T>::st(out + idx, (flags & CM_FLAG_SILU) ? silu_f<sizeof(T) == 4>(acc) : acc);
}

// ---- host dispatch ---------------------------------------------------------------------------------
static bool aligned_to(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

template <typename T>
static int pick_vec(const cm_conv_args& a, bool bwd) {
  // channel-last fast path needs unit channel stride everywhere
  if (a.x.sd != 1) return 0;
  if (bwd && a.dx.sd != 1) return 0;
  for (int r = 0; r < a.ndir; ++r)
    if (a.dir[r].out.sd != 1) return 0;
  int maxv = 2;   // measured on B200: 2 channels per thread (88 registers, 5+ CTAs/SM) beats 4 (154 registers)
  if (const char* e = getenv("CM_CONV_VEC")) maxv = std::max(1, std::min(maxv, atoi(e)));   // tuning experiments
  for (int vec = maxv; vec >= 1; vec >>= 1) {
    const size_t bytes = sizeof(T) * vec;
    if (bytes > 16) continue;
    bool ok = (a.dim % vec) == 0;
    auto chk = [&](const cm_tensor3& t) {
      ok = ok && aligned_to(t.ptr, bytes) && (t.sb % vec) == 0 && (t.sl % vec) == 0;
    };
    chk(a.x);
    if (bwd) chk(a.dx);
    for (int r = 0; r < a.ndir; ++r) chk(a.dir[r].out);
    if (ok) return vec;
  }
  return 1;
}

template <typename T>
static int launch_conv_t(const cm_conv_args& a, bool bwd, cudaStream_t st) {
  const int vec = pick_vec<T>(a, bwd);
  if (vec > 0) {
    const dim3 block(32, kTY);
    const dim3 grid(cm_ceil_div(cm_ceil_div(a.dim, vec), 32), cm_ceil_div(a.seqlen, kChunk), a.batch);
    if (!bwd) {
      if (vec >= 2 && (a.dim % 2) == 0 && launch_conv_fwd_sw<T>(a, st)) { /* sliding-window kernel */ }
      else if (vec == 4) conv_fwd_cl_kernel<T, 4><<<grid, block, 0, st>>>(a);
      else if (vec == 2) conv_fwd_cl_kernel<T, 2><<<grid, block, 0, st>>>(a);
      else conv_fwd_cl_kernel<T, 1><<<grid, block, 0, st>>>(a);
    } else {
      if (vec == 2 && launch_conv_bwd_sw<T>(a, st)) { /* sliding-window kernel */ }
      else if (vec == 2) conv_bwd_cl_kernel<T, 2><<<grid, block, 0, st>>>(a);
      else conv_bwd_cl_kernel<T, 1><<<grid, block, 0, st>>>(a);
    }
  } else {
    if (a.dim > 65535) return CM_ERR_UNSUPPORTED;
    if (!bwd) {
      const dim3 grid(cm_ceil_div(a.seqlen, 128), a.dim, a.batch);
      conv_fwd_generic_kernel<T><<<grid, 128, 0, st>>>(a);
    } else {
      const dim3 grid(cm_ceil_div(a.seqlen, kChunk), a.dim, a.batch);
      conv_bwd_generic_kernel<T><<<grid, kChunk, 0, st>>>(a);
    }
  }
  CM_LAUNCH_CHECK();
  return 0;
}

static int check_conv(const cm_conv_args& a, bool bwd) {
  if (a.batch <= 0 || a.dim <= 0 || a.seqlen <= 0) return CM_ERR_BAD_ARG;
  if (a.ndir != 1 && a.ndir != 2) return CM_ERR_BAD_ARG;
  if (!dtype_ok(a.dtype) || a.x.ptr == nullptr) return CM_ERR_BAD_ARG;
  if (a.width < 2 || a.width > CM_CONV_MAX_WIDTH) return CM_ERR_UNSUPPORTED;
  if (a.batch > 65535) return CM_ERR_UNSUPPORTED;
  for (int r = 0; r < a.ndir; ++r) {
    if (!a.dir[r].weight || !a.dir[r].out.ptr) return CM_ERR_BAD_ARG;
    if (bwd && !a.dir[r].dweight_part) return CM_ERR_BAD_ARG;
  }
  if (bwd && a.dx.ptr == nullptr) return CM_ERR_BAD_ARG;
  return 0;
}

}  // namespace cm

extern "C" int cm_conv_num_part(int32_t batch, int32_t seqlen) {
  if (batch <= 0 || seqlen <= 0) return 0;
  return batch * cm_ceil_div(seqlen, cm::kChunk);
}

static int conv_dispatch(const cm_conv_args* args, bool bwd, void* stream) {
  if (args == nullptr) return CM_ERR_BAD_ARG;
  const int e = cm::check_conv(*args, bwd);
  if (e) return e;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (args->dtype) {
    case CM_F32: return cm::launch_conv_t<float>(*args, bwd, st);
    case CM_BF16: return cm::launch_conv_t<__nv_bfloat16>(*args, bwd, st);
    default: return cm::launch_conv_t<__half>(*args, bwd, st);
  }
}

extern "C" int cm_conv_fwd(const cm_conv_args* args, void* stream) { return conv_dispatch(args, false, stream); }
extern "C" int cm_conv_bwd(const cm_conv_args* args, void* stream) { return conv_dispatch(args, true, stream); }

extern "C" int cm_conv_update(const void* x, void* conv_state, const float* weight, const float* bias, void* out,
                              int32_t batch, int32_t dim, int32_t width, int32_t dtype, uint32_t flags, void* stream) {
  if (!x || !conv_state || !weight || !out || batch <= 0 || dim <= 0) return CM_ERR_BAD_ARG;
  if (!cm::dtype_ok(dtype)) return CM_ERR_BAD_ARG;
  if (width < 2 || width > CM_CONV_MAX_WIDTH) return CM_ERR_UNSUPPORTED;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int threads = 128;
  const unsigned blocks = (unsigned)(((int64_t)batch * dim + threads - 1) / threads);
  switch (dtype) {
    case CM_F32:
      cm::conv_update_kernel<float><<<blocks, threads, 0, st>>>(static_cast<const float*>(x), static_cast<float*>(conv_state),
                                                              weight, bias, static_cast<float*>(out), batch, dim, width, flags);
      break;
    case CM_BF16:
      cm::conv_update_kernel<__nv_bfloat16><<<blocks, threads, 0, st>>>(
          static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(conv_state), weight, bias,
          static_cast<__nv_bfloat16*>(out), batch, dim, width, flags);
      break;
    default:
      cm::conv_update_kernel<__half><<<blocks, threads, 0, st>>>(static_cast<const __half*>(x), static_cast<__half*>(conv_state),
                                                               weight, bias, static_cast<__half*>(out), batch, dim, width, flags);
      break;
  }
  CM_LAUNCH_CHECK();
  return 0;
}
