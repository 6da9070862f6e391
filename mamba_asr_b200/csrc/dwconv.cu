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

// ---- tensor-pipe forward / backward-data for 16-bit tensors (round 2, last session) ---------------------------------------
// The tile kernel above is FP32-issue-bound: 31 FFMA-instructions per 32 outputs.  Per channel the conv is a banded Toeplitz
// product, so it fits mma.sync.m16n8k16 (bf16 / fp16 operands, fp32 accumulate):
//     Y[i][n] = sum_s sum_j A_s[i][j] * B_s[j][n],   i = time inside a 16-block, n = one of 8 consecutive 16-blocks,
//     A_s[i][j] = w[16 s + j - i]  (0 outside the taps),   B_s[j][n] = x[16 n + 16 s + j]  (slab-relative time),  s < NS
// i.e. 128 outputs of one channel take NS = 3 MMAs (K = 31) whose B fragments are two aligned 32-bit loads from a
// [channel][time] shared tile, and the A fragments are built once per (CTA, channel) from the taps (a Toeplitz block has only
// 3 distinct register pairs per slice: A[g+8][2t+8..] = A[g][2t..]).  The taps are rounded to the tensors' 16-bit type - what
// the reference's autocast conv does with its fp32 weights.  CTA = 4 warps = 32 channels x 128 times of one batch: 16-byte
// global loads scattered into the channel-major tile (odd word stride: conflict-free) -> a warp takes 8 channels in turn and
// keeps their accumulators, so that a lane stores 16-byte rows (8 adjacent channels) of its 4 times.  fp32 tensors keep the
// FFMA kernel.
constexpr int kDmT = 128;                    // output times per CTA
template <int K> struct DwMma {
  static constexpr int NS = (16 + K - 1 + 15) / 16;      // K slices of 16 input times (K = 31: 3)
  static constexpr int RX = kDmT - 16 + 16 * NS;         // slab rows (input times) the MMAs read (K = 31: 160)
  static constexpr int XT = RX + 2;                      // row stride of the transposed tile in elements: 81 words for K = 31
  static_assert((XT / 2) % 2 == 1, "odd word stride keeps the transpose conflict-free");
};

template <typename T> struct Mma16;
template <> struct Mma16<__nv_bfloat16> {
  static __device__ __forceinline__ void mma(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  static __device__ __forceinline__ unsigned short cvt(float v) {
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    return *reinterpret_cast<const unsigned short*>(&h);
  }
  static __device__ __forceinline__ uint32_t pack(float a, float b) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
  }
};
template <> struct Mma16<__half> {
  static __device__ __forceinline__ void mma(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  static __device__ __forceinline__ unsigned short cvt(float v) {
    const __half h = __float2half_rn(v);
    return *reinterpret_cast<const unsigned short*>(&h);
  }
  static __device__ __forceinline__ uint32_t pack(float a, float b) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
  }
};

template <typename T, int K>
__global__ void __launch_bounds__(128) dwconv_fwd_mma_kernel(const cm_dwconv_args p) {
  using DM = DwMma<K>;
  constexpr int NS = DM::NS, RX = DM::RX, XT = DM::XT;
  constexpr int WP = 64;                       // tap-pair row: entry k + 16 = (w[k], w[k + 1]) for k in [-16, 48)
  static_assert(NS <= 3, "the tap-pair row covers three K slices");
  extern __shared__ __align__(16) unsigned char dw_smem[];
  uint32_t* wp = reinterpret_cast<uint32_t*>(dw_smem);                         // [32][WP]   packed 16-bit tap pairs
  unsigned short* xt = reinterpret_cast<unsigned short*>(wp + 32 * WP);        // [32][XT]   slab, channel-major
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int c_blk = blockIdx.x * 32;
  const int b = blockIdx.z;
  const int t0 = blockIdx.y * kDmT;
  const int L = p.seqlen;
  // slab rows t0 - pad_left .. + RX: every 16-byte chunk (one time, 8 channels) is requested first, ...
  constexpr int NCH = (RX * 4 + 127) / 128;     // chunks per thread
  uint4 chunk[NCH];
  {
    const T* xb = static_cast<const T*>(p.x.ptr) + b * p.x.sb;
#pragma unroll
    for (int q = 0; q < NCH; ++q) {
      const int i = tid + 128 * q;
      const int row = i >> 2, ch = i & 3;
      const int r = t0 - p.pad_left + row, c = c_blk + ch * 8;
      const bool ok = i < RX * 4 && r >= 0 && r < L && c < p.dim;
      chunk[q] = ok ? __ldg(reinterpret_cast<const uint4*>(xb + (int64_t)r * p.x.sl + c)) : make_uint4(0u, 0u, 0u, 0u);
    }
  }
  // taps of the CTA's 32 channels as PAIRS (w[k], w[k + 1]) rounded to T, zero outside [0, K): a Toeplitz fragment register
  // is then one 32-bit load (as single 16-bit entries the fragment build was 18 loads + 9 merges per channel and lane)
  // (all 32 tap loads of a thread are requested before the first is used: rolled, this loop was 16 L2 round trips in a row)
  {
    constexpr int NW = 32 * WP / 128;
    float v0[NW], v1[NW];
#pragma unroll
    for (int q = 0; q < NW; ++q) {
      const int i = tid + 128 * q;
      const int c = i >> 6, k = (i & 63) - 16;
      const int ch = c_blk + c;
      const bool in0 = ch < p.dim && k >= 0 && k < K, in1 = ch < p.dim && k + 1 >= 0 && k + 1 < K;
      const float* wr = p.weight + (int64_t)(ch < p.dim ? ch : 0) * K;
      v0[q] = __ldg(wr + (in0 ? (p.flip ? K - 1 - k : k) : 0));
      v1[q] = __ldg(wr + (in1 ? (p.flip ? K - 2 - k : k + 1) : 0));
      if (!in0) v0[q] = 0.f;
      if (!in1) v1[q] = 0.f;
    }
#pragma unroll
    for (int q = 0; q < NW; ++q) wp[tid + 128 * q] = Mma16<T>::pack(v0[q], v1[q]);
  }
  // ... then scattered into the channel-major tile: 8 two-byte stores per chunk.  A warp covers 8 rows x 4 chunks; for one
  // channel of the chunk the banks are 8 * chunk + row / 2 (word stride XT / 2 is odd): conflict-free
#pragma unroll
  for (int q = 0; q < NCH; ++q) {
    const int i = tid + 128 * q;
    if (i < RX * 4) {
      const int row = i >> 2, ch = i & 3;
      unsigned short* dst = xt + (ch * 8) * XT + row;
      const uint32_t w4[4] = {chunk[q].x, chunk[q].y, chunk[q].z, chunk[q].w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        dst[(2 * e) * XT] = (unsigned short)(w4[e] & 0xffffu);
        dst[(2 * e + 1) * XT] = (unsigned short)(w4[e] >> 16);
      }
    }
  }
  __syncthreads();
  const int g = lane >> 2, t = lane & 3;
  const int cw = c_blk + 8 * warp;              // the warp's 8 channels
  if (cw >= p.dim) return;                      // dim is a multiple of 8: a warp's channels are all inside or all outside
  float acc[8][4];
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int cl = 8 * warp + q;                // channel inside the CTA
    const uint32_t* wrow = wp + cl * WP + 16;
    const unsigned short* xrow = xt + cl * XT;
    const float bias = p.bias != nullptr ? __ldg(p.bias + cw + q) : 0.f;
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[q][e] = bias;
#pragma unroll
    for (int sl = 0; sl < NS; ++sl) {
      // A_s[i][j] = w[16 s + j - i]: a0 = (row g, cols 2t, 2t+1), a1 = (row g + 8, same cols), a2 = (row g, cols 2t + 8, + 9),
      // a3 = (row g + 8, cols 2t + 8, + 9) = a0
      const int kb = 16 * sl + 2 * t - g;
      const uint32_t a0 = wrow[kb], a1 = wrow[kb - 8], a2 = wrow[kb + 8];
      // B_s[j][n] = x[16 n + 16 s + j]: b0 = (rows 2t, 2t+1, col g), b1 = (rows 2t + 8, + 9, col g)
      const uint32_t* xb32 = reinterpret_cast<const uint32_t*>(xrow + 16 * g + 16 * sl + 2 * t);
      Mma16<T>::mma(acc[q], a0, a1, a2, a0, xb32[0], xb32[4]);
    }
  }
  // accumulators of channel q: (time 16 * 2t + g, 16 * (2t + 1) + g) in slots 0, 1 and the same + 8 in slots 2, 3
  T* yb = static_cast<T*>(p.y.ptr) + b * p.y.sb + cw;
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int tt = t0 + 16 * (2 * t + (e & 1)) + g + 8 * (e >> 1);
    if (tt < L) {
      uint4 v;
      v.x = Mma16<T>::pack(acc[0][e], acc[1][e]); v.y = Mma16<T>::pack(acc[2][e], acc[3][e]);
      v.z = Mma16<T>::pack(acc[4][e], acc[5][e]); v.w = Mma16<T>::pack(acc[6][e], acc[7][e]);
      *reinterpret_cast<uint4*>(yb + (int64_t)tt * p.y.sl) = v;
    }
  }
}

template <typename T>
static bool dw_mma_ok(const cm_dwconv_args& a) {
  // default for 16-bit tensors (17.5 us against 22.5 us for the FFMA tile kernel at 64 x 501 x 256 in the step);
  // CM_DWCONV_NO_MMA=1 keeps the FFMA kernel (fp32 taps)
  if (sizeof(T) != 2 || getenv("CM_DWCONV_NO_MMA") != nullptr) return false;
  // 16-byte output rows of 8 channels: y rows and the channel offset of a warp must be 16-byte aligned
  return a.y.ptr != nullptr && a.y.sd == 1 && (reinterpret_cast<uintptr_t>(a.y.ptr) & 15) == 0 && a.y.sl % 8 == 0 && a.y.sb % 8 == 0;
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
    } else if (dw_mma_ok<T>(a)) {
      if constexpr (sizeof(T) == 2) {
        using DM = DwMma<K>;
        const size_t smem = (size_t)32 * DM::XT * 2 + (size_t)32 * 64 * 4;
        const dim3 mgrid(cm_ceil_div(a.dim, 32), cm_ceil_div(a.seqlen, kDmT), a.batch);
        dwconv_fwd_mma_kernel<T, K><<<mgrid, 128, smem, st>>>(a);
      }
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
