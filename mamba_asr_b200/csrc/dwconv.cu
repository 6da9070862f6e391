// Depthwise conv1d over time, channel-last, for sm_100a: the k = 31 convolution of the ConMamba convolution module
// (reference modules/Conmamba.py:281-290: nn.Conv1d(C, C, kernel_size, padding, groups=C)).
//
//   y[b, l, c] = bias[c] + sum_k w[c, k] * x[b, l - pad_left + k, c]            (zero outside [0, L))
//
// torch evaluates it on (B, C, L) tensors through conv_depthwise2d kernels - measured on B200 at 64 x 501 x 256: 179 us
// forward, 211 us backward-data, 396 us backward-weight per layer, plus the two transposes around it; the op itself is
// 254 MFMA and 33 MB per layer (7 us of FP32 issue, 5 us of HBM).
//
// Layout: a lane owns ONE channel (a warp row is 32 consecutive channels: coalesced in the channel-last activations),
// keeps the K taps and a K-deep sliding window of the input in registers and walks a slab of time steps; the loop is
// unrolled by K so that the window rotation is register renaming.  One new row per output, K FFMA per output.
//   forward / backward-data : the same kernel (backward-data = taps flipped, pad_left' = K-1-pad_left, no bias)
//   backward-weight         : dw[c,k] += dy[l] * x[l - pad_left + k]  with the same window; 4 warps of a CTA are summed in
//                             shared memory and each CTA writes one partial row (fixed-order reduction by cm_reduce_multi:
//                             deterministic, no atomics)
// Roof: HBM / FP32 issue (balanced).  Algorithmic bytes per position: 2s forward, 2s backward-data, 2s backward-weight.
#include <cstdlib>

#include "common.cuh"

namespace cm {

constexpr int kDwWarps = 4;

template <int K> struct DwCfg {
  static constexpr int NCH = (64 + K - 1) / K;   // chunks of K outputs per warp slab
  static constexpr int TW = NCH * K;             // outputs per warp slab
};

template <typename T, int K>
__global__ void __launch_bounds__(32 * kDwWarps) dwconv_fwd_kernel(const cm_dwconv_args p) {
  constexpr int TW = DwCfg<K>::TW;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  const bool act = c < p.dim;
  const int b = blockIdx.z;
  const int t0 = (blockIdx.y * kDwWarps + warp) * TW;
  const int L = p.seqlen;
  if (t0 >= L) return;
  const int cc = act ? c : 0;
  float w[K];
#pragma unroll
  for (int k = 0; k < K; ++k) w[k] = act ? __ldg(p.weight + (int64_t)cc * K + (p.flip ? K - 1 - k : k)) : 0.f;
  const float bias = (act && p.bias != nullptr) ? __ldg(p.bias + cc) : 0.f;
  const T* xp = static_cast<const T*>(p.x.ptr) + b * p.x.sb + cc * p.x.sd;
  T* yp = static_cast<T*>(p.y.ptr) + b * p.y.sb + cc * p.y.sd;
  const int64_t xsl = p.x.sl, ysl = p.y.sl;
  const int pad = p.pad_left;
  auto ldx = [&](int r) -> float { return (act && r >= 0 && r < L) ? Elem<T>::ld(xp + r * xsl) : 0.f; };
  float win[K];
#pragma unroll
  for (int i = 0; i < K - 1; ++i) win[i] = ldx(t0 - pad + i);
  win[K - 1] = 0.f;
#pragma unroll 1
  for (int o = 0; o < TW; o += K) {
    if (t0 + o >= L) break;
#pragma unroll
    for (int j = 0; j < K; ++j) {
      const int l = t0 + o + j;
      win[(K - 1 + j) % K] = ldx(l - pad + K - 1);
      float acc = bias;
#pragma unroll
      for (int k = 0; k < K; ++k) acc = fmaf(w[k], win[(j + k) % K], acc);
      if (act && l < L) Elem<T>::st(yp + l * ysl, acc);
    }
  }
}

template <typename T, int K>
__global__ void __launch_bounds__(32 * kDwWarps) dwconv_bwd_weight_kernel(const cm_dwconv_args p) {
  constexpr int TW = DwCfg<K>::TW;
  __shared__ float red[kDwWarps][K + 1][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  const bool act = c < p.dim;
  const int b = blockIdx.z;
  const int t0 = (blockIdx.y * kDwWarps + warp) * TW;
  const int L = p.seqlen;
  const int cc = act ? c : 0;
  const T* xp = static_cast<const T*>(p.x.ptr) + b * p.x.sb + cc * p.x.sd;
  const T* gp = static_cast<const T*>(p.dy.ptr) + b * p.dy.sb + cc * p.dy.sd;
  const int64_t xsl = p.x.sl, gsl = p.dy.sl;
  const int pad = p.pad_left;
  float dw[K], db = 0.f;
#pragma unroll
  for (int k = 0; k < K; ++k) dw[k] = 0.f;
  if (t0 < L) {
    auto ldx = [&](int r) -> float { return (act && r >= 0 && r < L) ? Elem<T>::ld(xp + r * xsl) : 0.f; };
    float win[K];
#pragma unroll
    for (int i = 0; i < K - 1; ++i) win[i] = ldx(t0 - pad + i);
    win[K - 1] = 0.f;
#pragma unroll 1
    for (int o = 0; o < TW; o += K) {
      if (t0 + o >= L) break;
#pragma unroll
      for (int j = 0; j < K; ++j) {
        const int l = t0 + o + j;
        win[(K - 1 + j) % K] = ldx(l - pad + K - 1);
        const float g = (act && l < L) ? Elem<T>::ld(gp + l * gsl) : 0.f;
        db += g;
#pragma unroll
        for (int k = 0; k < K; ++k) dw[k] = fmaf(g, win[(j + k) % K], dw[k]);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < K; ++k) red[warp][k][lane] = dw[k];
  red[warp][K][lane] = db;
  __syncthreads();
  // fixed-order sum over the CTA's warps; one partial row per CTA
  const int64_t part = (int64_t)blockIdx.z * gridDim.y + blockIdx.y;
  for (int i = threadIdx.x; i < (K + 1) * 32; i += blockDim.x) {
    const int k = i / 32, ln = i % 32;
    const int ch = blockIdx.x * 32 + ln;
    if (ch >= p.dim) continue;
    float a = 0.f;
#pragma unroll
    for (int wv = 0; wv < kDwWarps; ++wv) a += red[wv][k][ln];
    if (k < K) p.dweight_part[(part * p.dim + ch) * K + k] = a;
    else if (p.dbias_part != nullptr) p.dbias_part[part * p.dim + ch] = a;
  }
}

// ---- shared-memory tiled kernels (round 2) --------------------------------------------------------------------------
// The kernels above load each input row where the window needs it: one DRAM round trip per 31 outputs in front of a
// dependent FMA chain, 126 registers, 49 / 60 us per launch at 64 x 501 x 256 against a 5 us HBM / 7 us FP32 balance
// (0.10 of the HBM peak).  Here a warp first brings the whole slab it needs - TW + K - 1 rows of 32 channels - into its
// private shared-memory tile with cp.async (16-byte chunks, every row in flight at once, rows outside [0, L) and channels
// outside [0, dim) zero-filled), then walks it in groups of R = 8 outputs: R + K - 1 conflict-free LDS per lane feed R * K
// FFMA from registers (taps in registers as before).  No global-memory latency inside the arithmetic, 80 registers.
constexpr int kDwR = 8;                      // outputs per register group
// Forward tile kernel, launch shape.  One warp per CTA: with the 4-warp CTAs of the backward-weight kernel the ConMamba-large
// launch (64 x 501 x 256) was 1024 CTAs of which half had two idle warps, in 1.38 waves of 5 CTAs / SM - the SMs were active
// 71 % of the 45 us (profiles/r02_dwconv_fwd_tile_kernel_cfg3_ncu.txt).  Single-warp CTAs of 96 registers and < 10 KB of
// shared memory are all resident at once (21 per SM).  A warp's slab is FTW outputs = a whole number of periods of its
// circular register window (below).
template <int K> struct DwFwd {
  static constexpr int WB = (kDwR + K - 1 + kDwR - 1) / kDwR * kDwR;   // window registers: R + K - 1 rounded up to groups
  static constexpr int PG = WB / kDwR;                                 // groups per period of the circular window
  static constexpr int FTW = WB * ((96 + WB - 1) / WB);                // outputs per warp slab (K = 31: 120)
  static constexpr int RX = FTW + K - 1;                               // input rows of the slab
};

// store under a predicate the compiler cannot sink the arithmetic into (an `if` around the store moves the 31 FFMA of the
// output under the branch as well: BSSY / BRA / BSYNC per output)
template <typename T>
__device__ __forceinline__ void dw_st_pred(T* ptr, float v, bool ok) {
  if constexpr (sizeof(T) == 4) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %2, 0;\n@p st.global.f32 [%0], %1;\n}" ::"l"(ptr), "f"(v), "r"((int)ok) : "memory");
  } else {
    T t;
    Elem<T>::st(&t, v);
    const unsigned short r = *reinterpret_cast<const unsigned short*>(&t);
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %2, 0;\n@p st.global.b16 [%0], %1;\n}" ::"l"(ptr), "h"(r), "r"((int)ok) : "memory");
  }
}

__device__ __forceinline__ void cp_async16_zfill(void* dst, const void* src, bool valid) {
  const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(dst));
  const int n = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(n) : "memory");
}

template <int K> struct DwTile {
  static constexpr int TW = DwCfg<K>::TW;
  static constexpr int NG = (TW + kDwR - 1) / kDwR;        // groups per slab
  static constexpr int RX = NG * kDwR + K - 1;             // input rows of a warp tile
  static constexpr int RG = NG * kDwR;                     // upstream-gradient rows (backward-weight)
};

// rows [r0, r0 + nrows) x channels [c_blk, c_blk + 32) of a (batch b) channel-last tensor -> tile[nrows][32]
template <typename T>
__device__ __forceinline__ void dw_load_tile(T* tile, const T* base, int64_t sl, int r0, int nrows, int L, int c_blk, int dim,
                                             int lane) {
  constexpr int EPC = 16 / (int)sizeof(T);    // elements per 16-byte chunk
  constexpr int CPR = 32 / EPC;               // chunks per row
  for (int i = lane; i < nrows * CPR; i += 32) {
    const int row = i / CPR, ch = i - row * CPR;
    const int r = r0 + row, c = c_blk + ch * EPC;
    const bool ok = r >= 0 && r < L && c < dim;
    cp_async16_zfill(tile + row * 32 + ch * EPC, ok ? base + (int64_t)r * sl + c : base, ok);
  }
}

template <typename T, int K>
__global__ void __maxnreg__(96) dwconv_fwd_tile_kernel(const cm_dwconv_args p) {
  using DF = DwFwd<K>;
  constexpr int WB = DF::WB;
  extern __shared__ __align__(16) unsigned char dw_smem[];
  const int lane = threadIdx.x;
  const int c_blk = blockIdx.x * 32;
  const int c = c_blk + lane;
  const bool act = c < p.dim;
  const int b = blockIdx.z;
  const int t0 = blockIdx.y * DF::FTW;
  const int L = p.seqlen;
  T* tile = reinterpret_cast<T*>(dw_smem);
  dw_load_tile<T>(tile, static_cast<const T*>(p.x.ptr) + b * p.x.sb, p.x.sl, t0 - p.pad_left, DF::RX, L, c_blk, p.dim, lane);
  asm volatile("cp.async.commit_group;" ::: "memory");
  const int cc = act ? c : 0;
  float w[K];
#pragma unroll
  for (int k = 0; k < K; ++k) w[k] = act ? __ldg(p.weight + (int64_t)cc * K + (p.flip ? K - 1 - k : k)) : 0.f;
  const float bias = (act && p.bias != nullptr) ? __ldg(p.bias + cc) : 0.f;
  T* yp = static_cast<T*>(p.y.ptr) + b * p.y.sb + cc;
  const int64_t ysl = p.y.sl;
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncwarp();
  const T* col = tile + lane;
  auto ld_in = [&](int row) -> float {
    if constexpr (sizeof(T) == 4) return reinterpret_cast<const float*>(col)[row * 32];
    else return Elem<T>::cvt(reinterpret_cast<const unsigned short*>(col)[row * 32]);
  };
  // Circular register window: input row x of the slab lives in in[x % WB].  A group of kDwR outputs loads only its kDwR new
  // rows; a period of PG groups is unrolled so that every index is a compile-time register name (no moves), and the loop
  // over periods is rolled (K = 31: 5 groups = 1240 FFMA per iteration; the fully unrolled slab is 64 KB of code).
  float in[WB];
#pragma unroll
  for (int i = 0; i < K - 1; ++i) in[i] = ld_in(i);
#pragma unroll 1
  for (int o = 0; o < DF::FTW; o += WB) {
#pragma unroll
    for (int g = 0; g < DF::PG; ++g) {
      const int og = o + g * kDwR;
      if (t0 + og >= L) return;               // an exit, not a join
#pragma unroll
      for (int j = 0; j < kDwR; ++j) in[(g * kDwR + K - 1 + j) % WB] = ld_in(og + K - 1 + j);
#pragma unroll
      for (int j = 0; j < kDwR; ++j) {
        float acc = bias;
#pragma unroll
        for (int k = 0; k < K; ++k) acc = fmaf(w[k], in[(g * kDwR + j + k) % WB], acc);
        const int l = t0 + og + j;
        dw_st_pred<T>(yp + l * ysl, acc, act && l < L);
      }
    }
  }
}

template <typename T, int K>
__global__ void __launch_bounds__(32 * kDwWarps) dwconv_bwdw_tile_kernel(const cm_dwconv_args p) {
  using DT = DwTile<K>;
  extern __shared__ __align__(16) unsigned char dw_smem[];
  __shared__ float red[kDwWarps][K + 1][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c_blk = blockIdx.x * 32;
  const int c = c_blk + lane;
  const bool act = c < p.dim;
  const int b = blockIdx.z;
  const int t0 = (blockIdx.y * kDwWarps + warp) * DT::TW;
  const int L = p.seqlen;
  float dw[K], db = 0.f;
#pragma unroll
  for (int k = 0; k < K; ++k) dw[k] = 0.f;
  if (t0 < L) {
    T* xt = reinterpret_cast<T*>(dw_smem) + (size_t)warp * (DT::RX + DT::RG) * 32;
    T* gt = xt + DT::RX * 32;
    dw_load_tile<T>(xt, static_cast<const T*>(p.x.ptr) + b * p.x.sb, p.x.sl, t0 - p.pad_left, DT::RX, L, c_blk, p.dim, lane);
    // upstream-gradient rows of this slab only (rows of the next slab are zeroed: they belong to the next warp)
    dw_load_tile<T>(gt, static_cast<const T*>(p.dy.ptr) + b * p.dy.sb, p.dy.sl, t0, DT::RG, min(L, t0 + DT::TW), c_blk, p.dim, lane);
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    const T* xc = xt + lane;
    const T* gc = gt + lane;
#pragma unroll 1
    for (int o = 0; o < DT::TW; o += kDwR) {
      if (t0 + o >= L) break;
      float in[kDwR + K - 1], g[kDwR];
#pragma unroll
      for (int i = 0; i < kDwR + K - 1; ++i) {
        if constexpr (sizeof(T) == 4) in[i] = reinterpret_cast<const float*>(xc)[(o + i) * 32];
        else in[i] = Elem<T>::cvt(reinterpret_cast<const unsigned short*>(xc)[(o + i) * 32]);
      }
#pragma unroll
      for (int j = 0; j < kDwR; ++j) {
        if constexpr (sizeof(T) == 4) g[j] = reinterpret_cast<const float*>(gc)[(o + j) * 32];
        else g[j] = Elem<T>::cvt(reinterpret_cast<const unsigned short*>(gc)[(o + j) * 32]);
        db += g[j];
      }
#pragma unroll
      for (int k = 0; k < K; ++k) {
#pragma unroll
        for (int j = 0; j < kDwR; ++j) dw[k] = fmaf(g[j], in[j + k], dw[k]);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < K; ++k) red[warp][k][lane] = dw[k];
  red[warp][K][lane] = db;
  __syncthreads();
  // fixed-order sum over the CTA's warps; one partial row per CTA
  const int64_t part = (int64_t)blockIdx.z * gridDim.y + blockIdx.y;
  for (int i = threadIdx.x; i < (K + 1) * 32; i += blockDim.x) {
    const int k = i / 32, ln = i % 32;
    const int ch = blockIdx.x * 32 + ln;
    if (ch >= p.dim) continue;
    float a = 0.f;
#pragma unroll
    for (int wv = 0; wv < kDwWarps; ++wv) a += red[wv][k][ln];
    if (k < K) p.dweight_part[(part * p.dim + ch) * K + k] = a;
    else if (p.dbias_part != nullptr) p.dbias_part[part * p.dim + ch] = a;
  }
}

// the tiled kernels need channel-last tensors whose rows can be fetched in 16-byte chunks
template <typename T>
static bool dw_tile_ok(const cm_dwconv_args& a, bool wgrad) {
  if (getenv("CM_DWCONV_NO_TILE") != nullptr) return false;
  constexpr int EPC = 16 / (int)sizeof(T);
  auto ok = [&](const cm_tensor3& t) {
    return t.ptr != nullptr && t.sd == 1 && (reinterpret_cast<uintptr_t>(t.ptr) & 15) == 0 && t.sl % EPC == 0 && t.sb % EPC == 0;
  };
  if (a.dim % EPC != 0 || !ok(a.x)) return false;
  if (wgrad) return ok(a.dy);
  return a.y.ptr != nullptr && a.y.sd == 1;
}

template <typename T, int K>
static int dw_launch(const cm_dwconv_args& a, bool wgrad, cudaStream_t st) {
  constexpr int TW = DwCfg<K>::TW;
  const dim3 grid(cm_ceil_div(a.dim, 32), cm_ceil_div(a.seqlen, kDwWarps * TW), a.batch);
  if (dw_tile_ok<T>(a, wgrad)) {
    using DT = DwTile<K>;
    if (wgrad) {
      const size_t smem = (size_t)kDwWarps * (DT::RX + DT::RG) * 32 * sizeof(T);
      auto kern = dwconv_bwdw_tile_kernel<T, K>;
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);   // per device
      if (e != cudaSuccess) return (int)e;
      kern<<<grid, 32 * kDwWarps, smem, st>>>(a);
    } else {
      const size_t smem = (size_t)DwFwd<K>::RX * 32 * sizeof(T);
      const dim3 fgrid(cm_ceil_div(a.dim, 32), cm_ceil_div(a.seqlen, DwFwd<K>::FTW), a.batch);
      auto kern = dwconv_fwd_tile_kernel<T, K>;
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return (int)e;
      kern<<<fgrid, 32, smem, st>>>(a);
    }
    CM_LAUNCH_CHECK();
    return 0;
  }
  if (wgrad) dwconv_bwd_weight_kernel<T, K><<<grid, 32 * kDwWarps, 0, st>>>(a);
  else dwconv_fwd_kernel<T, K><<<grid, 32 * kDwWarps, 0, st>>>(a);
  CM_LAUNCH_CHECK();
  return 0;
}

template <typename T>
static int dw_dispatch(const cm_dwconv_args& a, bool wgrad, cudaStream_t st) {
  switch (a.ksize) {
    case 3: return dw_launch<T, 3>(a, wgrad, st);
    case 7: return dw_launch<T, 7>(a, wgrad, st);
    case 15: return dw_launch<T, 15>(a, wgrad, st);
    case 31: return dw_launch<T, 31>(a, wgrad, st);
    default: return CM_ERR_UNSUPPORTED;
  }
}

static int dw_slab(int ksize) {
  switch (ksize) {
    case 3: return DwCfg<3>::TW;
    case 7: return DwCfg<7>::TW;
    case 15: return DwCfg<15>::TW;
    case 31: return DwCfg<31>::TW;
    default: return 0;
  }
}

}  // namespace cm

extern "C" int cm_dwconv_num_part(int32_t batch, int32_t seqlen, int32_t ksize) {
  const int tw = cm::dw_slab(ksize);
  if (batch <= 0 || seqlen <= 0 || tw == 0) return CM_ERR_BAD_ARG;
  return batch * cm_ceil_div(seqlen, cm::kDwWarps * tw);
}

static int dw_check(const cm_dwconv_args* a) {
  if (a == nullptr || a->batch <= 0 || a->dim <= 0 || a->seqlen <= 0 || a->x.ptr == nullptr) return CM_ERR_BAD_ARG;
  if (!cm::dtype_ok(a->dtype)) return CM_ERR_BAD_ARG;
  if (a->batch > 65535) return CM_ERR_UNSUPPORTED;
  if (cm::dw_slab(a->ksize) == 0) return CM_ERR_UNSUPPORTED;
  if (a->pad_left < 0 || a->pad_left >= a->ksize) return CM_ERR_BAD_ARG;
  return 0;
}

extern "C" int cm_dwconv_fwd(const cm_dwconv_args* a, void* stream) {
  if (int e = dw_check(a)) return e;
  if (a->y.ptr == nullptr || a->weight == nullptr) return CM_ERR_BAD_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (a->dtype) {
    case CM_F32: return cm::dw_dispatch<float>(*a, false, st);
    case CM_BF16: return cm::dw_dispatch<__nv_bfloat16>(*a, false, st);
    default: return cm::dw_dispatch<__half>(*a, false, st);
  }
}

extern "C" int cm_dwconv_bwd_weight(const cm_dwconv_args* a, void* stream) {
  if (int e = dw_check(a)) return e;
  if (a->dy.ptr == nullptr || a->dweight_part == nullptr) return CM_ERR_BAD_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (a->dtype) {
    case CM_F32: return cm::dw_dispatch<float>(*a, true, st);
    case CM_BF16: return cm::dw_dispatch<__nv_bfloat16>(*a, true, st);
    default: return cm::dw_dispatch<__half>(*a, true, st);
  }
}
