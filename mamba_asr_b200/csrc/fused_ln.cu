// Residual add + dropout + LayerNorm in one pass, forward and backward, for sm_100a.
//
// SURVEY.md section 8(f) rank 2 ("fusable LN+residual", reference modules/Conmamba.py:638-649).  Every sub-block of a
// ConMamba layer ends in   s = a + alpha * dropout(b)   and the next one starts with   y = LayerNorm(s)
// (ffn_module1 -> norm1, mamba + skip -> convolution_module.layer_norm, conv module -> ffn_module2[0], ffn_module2 ->
// norm2).  As separate torch kernels that is dropout (read b, write b', write mask), scale, add (read a, read b', write
// s) and the norm (read s, write y): 25 bytes per element at fp32 residual / bf16 branch; fused it is 13 (read a, b; write
// s, y, mask), and backward likewise folds the norm's dx, the residual gradient, the dropout mask and alpha into one pass
// that also produces the dgamma / dbeta partial rows.
//
// Layout and method are those of layernorm.cu's pair-vectorised kernels: one warp owns two rows at a time, a lane owns
// column pairs lane + 32 i, statistics by warp shuffles in fp32.  The dropout mask comes from a counter-based hash of
// (seed, call id, element index): no RNG state in the kernel, a fresh mask per call site and per step (the seed is read
// from device memory, so a CUDA-graph replay sees the value the host advanced before it); the mask is stored (1 byte per
// element) and re-read by backward.  Roof: HBM.
#include <cstdlib>

#include "common.cuh"

namespace cm {

constexpr int kFlWarps = 8;

template <typename T> struct Fl2;
template <> struct Fl2<float> {
  static __device__ __forceinline__ float2 ld(const float* p) { return __ldg(reinterpret_cast<const float2*>(p)); }
  static __device__ __forceinline__ void st(float* p, float2 v) { *reinterpret_cast<float2*>(p) = v; }
};
template <> struct Fl2<__nv_bfloat16> {
  static __device__ __forceinline__ float2 ld(const __nv_bfloat16* p) {
    const uint32_t r = __ldg(reinterpret_cast<const uint32_t*>(p));
    return make_float2(__uint_as_float(r << 16), __uint_as_float(r & 0xffff0000u));
  }
  static __device__ __forceinline__ void st(__nv_bfloat16* p, float2 v) {
    *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(v.x, v.y);
  }
};

__device__ __forceinline__ uint32_t mix32(uint32_t x) {   // "lowbias32" integer finaliser
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}
__device__ __forceinline__ uint32_t drop_key(const int64_t* seed, uint32_t call_id) {
  const uint64_t s = seed ? static_cast<uint64_t>(*seed) : 0x243F6A8885A308D3ull;
  return mix32(static_cast<uint32_t>(s) ^ mix32(static_cast<uint32_t>(s >> 32) + call_id * 0x9E3779B9u + 0x85EBCA6Bu));
}
// Keep bits of the four columns 4 qi .. 4 qi + 3 of a row: the row enters once (row_key, hoisted out of the column loop),
// a quad costs two integer finalisers (4 x 16-bit uniform samples against thr = p * 65536).  Bit e = keep column 4 qi + e.
__device__ __forceinline__ uint32_t row_key(uint32_t key, int64_t row) {
  return mix32(key ^ (static_cast<uint32_t>(row) * 0x9E3779B9u) ^ (static_cast<uint32_t>(static_cast<uint64_t>(row) >> 32) * 0x68E31DA4u));
}
__device__ __forceinline__ uint32_t keep4(uint32_t rk, int qi, uint32_t thr) {
  const uint32_t h = mix32(rk + static_cast<uint32_t>(qi) * 0x85EBCA6Bu);
  const uint32_t h2 = mix32(h + 0x6A09E667u);
  return ((h & 0xffffu) >= thr ? 1u : 0u) | ((h >> 16) >= thr ? 2u : 0u) | ((h2 & 0xffffu) >= thr ? 4u : 0u) |
         ((h2 >> 16) >= thr ? 8u : 0u);
}
// the pair kernels take the two bits of column pair pi from its quad
__device__ __forceinline__ uint32_t keep2(uint32_t rk, int pi, uint32_t thr) { return (keep4(rk, pi >> 1, thr) >> (2 * (pi & 1))) & 3u; }

template <typename T> struct Fl4;
template <> struct Fl4<float> {
  static __device__ __forceinline__ void ld(const float* p, float* o) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(p));
    o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
  }
  static __device__ __forceinline__ void st(float* p, const float* v) { *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]); }
};
template <> struct Fl4<__nv_bfloat16> {
  static __device__ __forceinline__ void ld(const __nv_bfloat16* p, float* o) {
    const uint2 r = __ldg(reinterpret_cast<const uint2*>(p));
    o[0] = __uint_as_float(r.x << 16); o[1] = __uint_as_float(r.x & 0xffff0000u);
    o[2] = __uint_as_float(r.y << 16); o[3] = __uint_as_float(r.y & 0xffff0000u);
  }
  static __device__ __forceinline__ void st(__nv_bfloat16* p, const float* v) {
    const __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
    uint2 o;
    o.x = *reinterpret_cast<const uint32_t*>(&a);
    o.y = *reinterpret_cast<const uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p) = o;
  }
};

template <typename Ta, typename Tb, typename Ty, int NPP>
__global__ void __launch_bounds__(32 * kFlWarps) add_ln_fwd_kernel(const cm_add_ln_args A) {
  const int lane = threadIdx.x & 31;
  const int64_t row0 = ((int64_t)blockIdx.x * kFlWarps + (threadIdx.x >> 5)) * 2;
  if (row0 >= A.rows) return;
  const bool two = row0 + 1 < A.rows;
  const int C = A.cols, np = C >> 1;
  const Ta* a = static_cast<const Ta*>(A.a);
  const Tb* b = static_cast<const Tb*>(A.b);
  const bool drop = A.p_drop > 0.f && b != nullptr;
  const uint32_t thr = drop ? (uint32_t)(A.p_drop * 65536.0f) : 0u;
  const float keep_scale = drop ? A.alpha / (1.0f - A.p_drop) : A.alpha;
  const uint32_t key = drop ? drop_key(A.seed, A.call_id) : 0u;
  if (A.key != nullptr && blockIdx.x == 0 && threadIdx.x == 0) *A.key = key;
  float2 v[2][NPP];
  float s[2] = {0.f, 0.f};
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int64_t row = row0 + (two ? r : 0);
    const uint32_t rk = drop ? row_key(key, row) : 0u;
#pragma unroll
    for (int i = 0; i < NPP; ++i) {
      const int pi = lane + 32 * i;
      float2 t = make_float2(0.f, 0.f);
      if (pi < np) {
        t = Fl2<Ta>::ld(a + row * A.a_stride + 2 * pi);
        if (b != nullptr) {
          const float2 bv = Fl2<Tb>::ld(b + row * A.b_stride + 2 * pi);
          float kx = keep_scale, ky = keep_scale;
          if (drop) {
            const uint32_t kb = keep2(rk, pi, thr);
            const bool k0 = kb & 1u, k1 = kb & 2u;
            kx = k0 ? keep_scale : 0.f; ky = k1 ? keep_scale : 0.f;
            if (A.mask != nullptr && (r == 0 || two))
              *reinterpret_cast<uchar2*>(A.mask + row * (int64_t)C + 2 * pi) = make_uchar2(k0 ? 1 : 0, k1 ? 1 : 0);
          }
          t.x = fmaf(kx, bv.x, t.x); t.y = fmaf(ky, bv.y, t.y);
        }
        if (A.s != nullptr && (r == 0 || two)) Fl2<Ta>::st(static_cast<Ta*>(A.s) + row * A.s_stride + 2 * pi, t);
        // the statistics are taken from the value as the residual stream stores it (what backward re-reads)
        if (sizeof(Ta) == 2) t = make_float2(Elem<Ta>::round(t.x), Elem<Ta>::round(t.y));
      }
      v[r][i] = t;
      s[r] += t.x + t.y;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s[0] += __shfl_xor_sync(0xffffffffu, s[0], o);
    s[1] += __shfl_xor_sync(0xffffffffu, s[1], o);
  }
  const float invC = 1.f / (float)C;
  float mu[2] = {s[0] * invC, s[1] * invC}, q[2] = {0.f, 0.f};
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int i = 0; i < NPP; ++i) {
      const int pi = lane + 32 * i;
      if (pi < np) {
        const float d0 = v[r][i].x - mu[r], d1 = v[r][i].y - mu[r];
        q[r] = fmaf(d0, d0, fmaf(d1, d1, q[r]));
      }
    }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    q[0] += __shfl_xor_sync(0xffffffffu, q[0], o);
    q[1] += __shfl_xor_sync(0xffffffffu, q[1], o);
  }
  const float rs[2] = {rsqrtf(q[0] * invC + A.eps), rsqrtf(q[1] * invC + A.eps)};
  Ty* y = static_cast<Ty*>(A.y);
#pragma unroll
  for (int i = 0; i < NPP; ++i) {
    const int pi = lane + 32 * i;
    if (pi < np) {
      const float2 g = A.gamma ? __ldg(reinterpret_cast<const float2*>(A.gamma + 2 * pi)) : make_float2(1.f, 1.f);
      const float2 bb = A.beta ? __ldg(reinterpret_cast<const float2*>(A.beta + 2 * pi)) : make_float2(0.f, 0.f);
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        if (r == 1 && !two) break;
        Fl2<Ty>::st(y + (row0 + r) * A.y_stride + 2 * pi,
                    make_float2(fmaf((v[r][i].x - mu[r]) * rs[r], g.x, bb.x), fmaf((v[r][i].y - mu[r]) * rs[r], g.y, bb.y)));
      }
    }
  }
  if (lane == 0) {
    A.mean[row0] = mu[0]; A.rstd[row0] = rs[0];
    if (two) { A.mean[row0 + 1] = mu[1]; A.rstd[row0 + 1] = rs[1]; }
  }
}

// backward: total = LayerNorm-backward(dy; s) + ds ;  da = total ;  db = alpha * mask / (1 - p) * total
// One row per warp iteration, every load of the row (s, dy, ds, mask) issued before the first use: with the residual
// gradient and the mask fetched after the row reductions the kernel paid two memory round trips per row (measured 68 us
// for 32064 x 256 = 2.0 TB/s); two rows in flight per warp with all loads up front need 150 registers (1 CTA / SM).
// Measured at 32064 x 256 (fp32 residual, bf16 branch): 4 CTAs / SM (64 registers, 24 B spilled) 46 us, 2 CTAs 48 us,
// 3 CTAs 59 us.
#ifndef CM_FL_BWD_MINB
#define CM_FL_BWD_MINB 4
#endif
template <typename Ta, typename Tb, typename Ty, int NPP>
__global__ void __launch_bounds__(32 * kFlWarps, (NPP <= 4 ? CM_FL_BWD_MINB : 1)) add_ln_bwd_kernel(const cm_add_ln_args A) {
  __shared__ float2 red[kFlWarps][32 * NPP + 1];   // reused for dgamma, then dbeta
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int C = A.cols, np = C >> 1;
  const Ta* sx = static_cast<const Ta*>(A.s);
  const Ty* dy = static_cast<const Ty*>(A.dy);
  const Ta* ds = static_cast<const Ta*>(A.ds);
  Ta* da = static_cast<Ta*>(A.da);
  Tb* db_out = static_cast<Tb*>(A.db);
  const bool drop = A.p_drop > 0.f && db_out != nullptr;
  const bool regen = drop && A.mask == nullptr;
  const uint32_t thr = drop ? (uint32_t)(A.p_drop * 65536.0f) : 0u;
  const uint32_t key = regen ? __ldg(A.key) : 0u;
  const float keep_scale = drop ? A.alpha / (1.0f - A.p_drop) : A.alpha;
  float2 g[NPP], dg[NPP], db[NPP];
#pragma unroll
  for (int i = 0; i < NPP; ++i) {
    const int pi = lane + 32 * i;
    g[i] = (A.gamma && pi < np) ? __ldg(reinterpret_cast<const float2*>(A.gamma + 2 * pi)) : make_float2(1.f, 1.f);
    dg[i] = make_float2(0.f, 0.f); db[i] = make_float2(0.f, 0.f);
  }
  const float invC = 1.f / (float)C;
  const int64_t rstep = (int64_t)gridDim.x * kFlWarps;
  for (int64_t row = (int64_t)blockIdx.x * kFlWarps + warp; row < A.rows; row += rstep) {
    float2 xh[NPP], gy[NPP], e[NPP];
    unsigned short mk[NPP];
    const float mu = __ldg(A.mean + row), rs = __ldg(A.rstd + row);
#pragma unroll
    for (int i = 0; i < NPP; ++i) {
      const int pi = lane + 32 * i;
      const bool in = pi < np;
      xh[i] = in ? Fl2<Ta>::ld(sx + row * A.s_stride + 2 * pi) : make_float2(0.f, 0.f);
      gy[i] = in ? Fl2<Ty>::ld(dy + row * A.dy_stride + 2 * pi) : make_float2(0.f, 0.f);
      e[i] = (ds != nullptr && in) ? Fl2<Ta>::ld(ds + row * A.ds_stride + 2 * pi) : make_float2(0.f, 0.f);
      mk[i] = (drop && !regen && in) ? __ldg(reinterpret_cast<const unsigned short*>(A.mask + row * (int64_t)C + 2 * pi))
                                     : (unsigned short)0x0101;
    }
    if (regen) {
      const uint32_t rk = row_key(key, row);
#pragma unroll
      for (int i = 0; i < NPP; ++i) {
        const uint32_t kb = keep2(rk, lane + 32 * i, thr);
        mk[i] = (unsigned short)((kb & 1u) | ((kb & 2u) << 7));
      }
    }
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NPP; ++i) {
      const int pi = lane + 32 * i;
      const float2 dv = gy[i];
      const float2 h = (pi < np) ? make_float2((xh[i].x - mu) * rs, (xh[i].y - mu) * rs) : make_float2(0.f, 0.f);
      xh[i] = h;
      gy[i] = make_float2(dv.x * g[i].x, dv.y * g[i].y);
      dg[i].x = fmaf(dv.x, h.x, dg[i].x); dg[i].y = fmaf(dv.y, h.y, dg[i].y);
      db[i].x += dv.x; db[i].y += dv.y;
      s1 += gy[i].x + gy[i].y;
      s2 = fmaf(gy[i].x, h.x, fmaf(gy[i].y, h.y, s2));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    const float m1 = s1 * invC, m2 = s2 * invC;
#pragma unroll
    for (int i = 0; i < NPP; ++i) {
      const int pi = lane + 32 * i;
      if (pi < np) {
        const float2 t = make_float2(fmaf(rs, gy[i].x - m1 - xh[i].x * m2, e[i].x), fmaf(rs, gy[i].y - m1 - xh[i].y * m2, e[i].y));
        Fl2<Ta>::st(da + row * A.da_stride + 2 * pi, t);
        if (db_out != nullptr) {
          const float kx = (mk[i] & 0x00ffu) ? keep_scale : 0.f, ky = (mk[i] & 0xff00u) ? keep_scale : 0.f;
          Fl2<Tb>::st(db_out + row * A.db_stride + 2 * pi, make_float2(kx * t.x, ky * t.y));
        }
      }
    }
  }
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
    if (pass) __syncthreads();
#pragma unroll
    for (int i = 0; i < NPP; ++i) red[warp][lane + 32 * i] = pass ? db[i] : dg[i];
    __syncthreads();
    float* dst = pass ? A.dbeta_part : A.dgamma_part;
    for (int pi = threadIdx.x; pi < np; pi += blockDim.x) {
      float2 acc = make_float2(0.f, 0.f);
#pragma unroll
      for (int w = 0; w < kFlWarps; ++w) { acc.x += red[w][pi].x; acc.y += red[w][pi].y; }
      *reinterpret_cast<float2*>(dst + (int64_t)blockIdx.x * C + 2 * pi) = acc;
    }
  }
}

// ---- quad-vectorised kernels (cols % 4 == 0, strides % 4 == 0, 16-byte aligned fp32 / 8-byte aligned 16-bit rows) -------
// A lane owns NQ quads of four consecutive columns (lane + 32 i): 16-byte fp32 and 8-byte 16-bit accesses, half the memory
// instructions of the pair kernels above, one keep4() per quad.  ncu of the pair forward at 32064 x 256
// (profiles/r01_fused_addln_gelu_cfg3_ncu.txt): 560 warp instructions per row, issue slots 63 % busy, ALU pipe 44 % (the
// per-pair hash with its 64-bit index) - issue-bound at 58 % of the HBM peak; the pair backward: 64 registers with 26 B
// of spills per thread, long-scoreboard 17.5 stalls per issue.  The dropout mask is regenerated in backward from the
// forward's key (A.key) unless the caller asks for the stored byte mask (A.mask).
template <typename Ta, typename Tb, typename Ty, int NQ>
__global__ void __launch_bounds__(32 * kFlWarps) add_ln_fwd_q_kernel(const cm_add_ln_args A) {
  const int lane = threadIdx.x & 31;
  const int64_t row0 = ((int64_t)blockIdx.x * kFlWarps + (threadIdx.x >> 5)) * 2;
  const Ta* a = static_cast<const Ta*>(A.a);
  const Tb* b = static_cast<const Tb*>(A.b);
  const bool drop = A.p_drop > 0.f && b != nullptr;
  const uint32_t thr = drop ? (uint32_t)(A.p_drop * 65536.0f) : 0u;
  const float keep_scale = drop ? A.alpha / (1.0f - A.p_drop) : A.alpha;
  const uint32_t key = drop ? drop_key(A.seed, A.call_id) : 0u;
  if (A.key != nullptr && blockIdx.x == 0 && threadIdx.x == 0) *A.key = key;
  if (row0 >= A.rows) return;
  const bool two = row0 + 1 < A.rows;
  const int C = A.cols, nq = C >> 2;
  float v[2][NQ][4];
  float s[2] = {0.f, 0.f};
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int64_t row = row0 + (two ? r : 0);
    const bool live = r == 0 || two;
    const uint32_t rk = drop ? row_key(key, row) : 0u;
#pragma unroll
    for (int i = 0; i < NQ; ++i) {
      const int qi = lane + 32 * i;
      float t[4] = {0.f, 0.f, 0.f, 0.f};
      if (qi < nq) {
        Fl4<Ta>::ld(a + row * A.a_stride + 4 * qi, t);
        if (b != nullptr) {
          float bv[4];
          Fl4<Tb>::ld(b + row * A.b_stride + 4 * qi, bv);
          const uint32_t kb = drop ? keep4(rk, qi, thr) : 0xfu;
#pragma unroll
          for (int e = 0; e < 4; ++e) t[e] = fmaf(((kb >> e) & 1u) ? keep_scale : 0.f, bv[e], t[e]);
          if (drop && A.mask != nullptr && live)
            *reinterpret_cast<uint32_t*>(A.mask + row * (int64_t)C + 4 * qi) =
                (kb & 1u) | ((kb & 2u) << 7) | ((kb & 4u) << 14) | ((kb & 8u) << 21);
        }
        if (A.s != nullptr && live) Fl4<Ta>::st(static_cast<Ta*>(A.s) + row * A.s_stride + 4 * qi, t);
        // the statistics are taken from the value as the residual stream stores it (what backward re-reads)
        if (sizeof(Ta) == 2) {
#pragma unroll
          for (int e = 0; e < 4; ++e) t[e] = Elem<Ta>::round(t[e]);
        }
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) v[r][i][e] = t[e];
      s[r] += (t[0] + t[1]) + (t[2] + t[3]);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s[0] += __shfl_xor_sync(0xffffffffu, s[0], o);
    s[1] += __shfl_xor_sync(0xffffffffu, s[1], o);
  }
  const float invC = 1.f / (float)C;
  const float mu[2] = {s[0] * invC, s[1] * invC};
  float q[2] = {0.f, 0.f};
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int i = 0; i < NQ; ++i) {
      const bool in = lane + 32 * i < nq;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float d = in ? v[r][i][e] - mu[r] : 0.f;
        v[r][i][e] = d;
        q[r] = fmaf(d, d, q[r]);
      }
    }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    q[0] += __shfl_xor_sync(0xffffffffu, q[0], o);
    q[1] += __shfl_xor_sync(0xffffffffu, q[1], o);
  }
  const float rs[2] = {rsqrtf(q[0] * invC + A.eps), rsqrtf(q[1] * invC + A.eps)};
  Ty* y = static_cast<Ty*>(A.y);
#pragma unroll
  for (int i = 0; i < NQ; ++i) {
    const int qi = lane + 32 * i;
    if (qi < nq) {
      const float4 g = A.gamma ? __ldg(reinterpret_cast<const float4*>(A.gamma) + qi) : make_float4(1.f, 1.f, 1.f, 1.f);
      const float4 bb = A.beta ? __ldg(reinterpret_cast<const float4*>(A.beta) + qi) : make_float4(0.f, 0.f, 0.f, 0.f);
      const float g4[4] = {g.x, g.y, g.z, g.w}, b4[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        if (r == 1 && !two) break;
        float o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) o[e] = fmaf(v[r][i][e] * rs[r], g4[e], b4[e]);
        Fl4<Ty>::st(y + (row0 + r) * A.y_stride + 4 * qi, o);
      }
    }
  }
  if (lane == 0) {
    A.mean[row0] = mu[0]; A.rstd[row0] = rs[0];
    if (two) { A.mean[row0 + 1] = mu[1]; A.rstd[row0 + 1] = rs[1]; }
  }
}

// backward, quads: one row per warp iteration, every load of the row issued before the first use; gamma is re-read per row
// (L1) instead of living in registers, the keep bits are regenerated after the row reductions (no mask registers).
// A.dbsum_part != NULL: also the column sums of db (the bias gradient of the Linear that produced b), one more partial row.
#ifndef CM_FLQ_MINB
#define CM_FLQ_MINB 3
#endif
// NST > 0: the rows of s, ds and dy come through shared memory - lane 0 of a warp requests the warp's rows NST - 1
// iterations ahead as 1-D bulk copies (cp.async.bulk, one mbarrier per warp and stage), so a warp keeps NST - 1 rows of
// loads in flight while it reduces and stores the current one.  With one row per warp and iteration in registers the SM
// holds ~30 KB of loads in flight on average, below the latency x bandwidth product of HBM3e (53-65 % of the copy peak).
__device__ __forceinline__ uint32_t fl_smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void fl_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(fl_smem_u32(dst)), "l"(src), "r"(bytes), "r"(fl_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fl_mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(fl_smem_u32(bar)), "r"(parity) : "memory");
  } while (!done);
}
template <typename T> struct Fl4s;     // a quad out of a staged row (plain shared-memory loads)
template <> struct Fl4s<float> {
  static __device__ __forceinline__ void ld(const unsigned char* p, float* o) {
    const float4 v = *reinterpret_cast<const float4*>(p);
    o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
  }
};
template <> struct Fl4s<__nv_bfloat16> {
  static __device__ __forceinline__ void ld(const unsigned char* p, float* o) {
    const uint2 r = *reinterpret_cast<const uint2*>(p);
    o[0] = __uint_as_float(r.x << 16); o[1] = __uint_as_float(r.x & 0xffff0000u);
    o[2] = __uint_as_float(r.y << 16); o[3] = __uint_as_float(r.y & 0xffff0000u);
  }
};

// ACT: dy is the gradient of gelu(LayerNorm(s)) (the GELU epilogue of cm_layernorm_fwd, routed here by cm_layernorm_bwd):
// the pre-activation h * gamma + beta is recomputed and dy multiplied by gelu'(.) on the way in.
template <typename Ta, typename Tb, typename Ty, int NQ, bool COLSUM, int NST = 0, bool ACT = false>
__global__ void __launch_bounds__(32 * kFlWarps, (NQ <= 2 ? CM_FLQ_MINB : NQ <= 4 ? 2 : 1)) add_ln_bwd_q_kernel(const cm_add_ln_args A) {
  __shared__ float4 red[kFlWarps][32 * NQ + 1];   // reused for dgamma, dbeta and the db column sums
  extern __shared__ __align__(128) unsigned char fl_stage[];          // NST > 0: [warp][stage][s | ds | dy]
  __shared__ uint64_t fl_bar[kFlWarps][NST > 0 ? NST : 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int C = A.cols, nq = C >> 2;
  const Ta* sx = static_cast<const Ta*>(A.s);
  const Ty* dy = static_cast<const Ty*>(A.dy);
  const Ta* ds = static_cast<const Ta*>(A.ds);
  Ta* da = static_cast<Ta*>(A.da);
  Tb* db_out = static_cast<Tb*>(A.db);
  const bool drop = A.p_drop > 0.f && db_out != nullptr;
  const bool regen = drop && A.mask == nullptr;
  const uint32_t thr = drop ? (uint32_t)(A.p_drop * 65536.0f) : 0u;
  const uint32_t key = regen ? __ldg(A.key) : 0u;
  const float keep_scale = drop ? A.alpha / (1.0f - A.p_drop) : A.alpha;
  float dg[NQ][4], dbt[NQ][4], dbs[COLSUM ? NQ : 1][4];
#pragma unroll
  for (int i = 0; i < NQ; ++i)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      dg[i][e] = dbt[i][e] = 0.f;
      if (COLSUM) dbs[i][e] = 0.f;
    }
  const float invC = 1.f / (float)C;
  const int64_t rstep = (int64_t)gridDim.x * kFlWarps;
  const uint32_t seg_a = (uint32_t)C * (uint32_t)sizeof(Ta), seg_y = (uint32_t)C * (uint32_t)sizeof(Ty);
  const uint32_t row_bytes = 2 * seg_a + seg_y;
  unsigned char* const stage0 = fl_stage + (size_t)warp * (NST > 0 ? NST : 1) * row_bytes;
  auto request = [&](int64_t r, int st_i) {         // lane 0: the three row segments of row r into stage st_i
    uint64_t* bar = &fl_bar[warp][st_i];
    unsigned char* dst = stage0 + (size_t)st_i * row_bytes;
    const uint32_t total = seg_a + seg_y + (ds != nullptr ? seg_a : 0u);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fl_smem_u32(bar)), "r"(total) : "memory");
    fl_bulk_g2s(dst, sx + r * A.s_stride, seg_a, bar);
    if (ds != nullptr) fl_bulk_g2s(dst + seg_a, ds + r * A.ds_stride, seg_a, bar);
    fl_bulk_g2s(dst + 2 * seg_a, dy + r * A.dy_stride, seg_y, bar);
  };
  int st_cur = 0;
  uint32_t st_par = 0;
  if (NST > 0) {
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < (NST > 0 ? NST : 1); ++i)
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(fl_smem_u32(&fl_bar[warp][i])) : "memory");
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      int64_t r = (int64_t)blockIdx.x * kFlWarps + warp;
      for (int i = 0; i < NST - 1 && r < A.rows; ++i, r += rstep) request(r, i);
    }
    __syncwarp();
  }
  for (int64_t row = (int64_t)blockIdx.x * kFlWarps + warp; row < A.rows; row += rstep) {
    float xh[NQ][4], gy[NQ][4], e4[NQ][4];
    const float mu = __ldg(A.mean + row), rs = __ldg(A.rstd + row);
    if (NST > 0) {
      // the stage read in the previous iteration is free (every lane's loads fed the shuffles of that iteration): refill it
      __syncwarp();
      if (lane == 0) {
        const int64_t rn = row + (int64_t)(NST - 1) * rstep;
        if (rn < A.rows) {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          request(rn, st_cur == 0 ? NST - 1 : st_cur - 1);
        }
      }
      fl_mbar_wait(&fl_bar[warp][st_cur], st_par);
    }
    const unsigned char* const stg = stage0 + (size_t)st_cur * row_bytes;
#pragma unroll
    for (int i = 0; i < NQ; ++i) {
      const int qi = lane + 32 * i;
      const bool in = qi < nq;
#pragma unroll
      for (int e = 0; e < 4; ++e) xh[i][e] = gy[i][e] = e4[i][e] = 0.f;
      if (in) {
        if (NST > 0) {
          Fl4s<Ta>::ld(stg + 4 * sizeof(Ta) * qi, xh[i]);
          Fl4s<Ty>::ld(stg + 2 * seg_a + 4 * sizeof(Ty) * qi, gy[i]);
          if (ds != nullptr) Fl4s<Ta>::ld(stg + seg_a + 4 * sizeof(Ta) * qi, e4[i]);
        } else {
          Fl4<Ta>::ld(sx + row * A.s_stride + 4 * qi, xh[i]);
          Fl4<Ty>::ld(dy + row * A.dy_stride + 4 * qi, gy[i]);
          if (ds != nullptr) Fl4<Ta>::ld(ds + row * A.ds_stride + 4 * qi, e4[i]);
        }
      }
    }
    if (NST > 0) {
      if (++st_cur == NST) { st_cur = 0; st_par ^= 1u; }
    }
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NQ; ++i) {
      const int qi = lane + 32 * i;
      const bool in = qi < nq;
      const float4 gv = (A.gamma && in) ? __ldg(reinterpret_cast<const float4*>(A.gamma) + qi) : make_float4(1.f, 1.f, 1.f, 1.f);
      const float g4[4] = {gv.x, gv.y, gv.z, gv.w};
      const float4 bv = (ACT && A.beta && in) ? __ldg(reinterpret_cast<const float4*>(A.beta) + qi) : make_float4(0.f, 0.f, 0.f, 0.f);
      const float b4[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float dv = gy[i][e];
        const float h = in ? (xh[i][e] - mu) * rs : 0.f;
        if (ACT) dv *= gelu_grad_f(fmaf(h, g4[e], b4[e]));
        xh[i][e] = h;
        const float gg = dv * g4[e];
        gy[i][e] = gg;
        dg[i][e] = fmaf(dv, h, dg[i][e]);
        dbt[i][e] += dv;
        s1 += gg;
        s2 = fmaf(gg, h, s2);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    const float m1 = s1 * invC, m2 = s2 * invC;
    const uint32_t rk = regen ? row_key(key, row) : 0u;
#pragma unroll
    for (int i = 0; i < NQ; ++i) {
      const int qi = lane + 32 * i;
      if (qi < nq) {
        float t[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) t[e] = fmaf(rs, gy[i][e] - m1 - xh[i][e] * m2, e4[i][e]);
        Fl4<Ta>::st(da + row * A.da_stride + 4 * qi, t);
        if (db_out != nullptr) {
          uint32_t kb = 0xfu;
          if (regen) {
            kb = keep4(rk, qi, thr);
          } else if (drop) {
            const uint32_t m = __ldg(reinterpret_cast<const uint32_t*>(A.mask + row * (int64_t)C + 4 * qi));
            kb = ((m & 0xffu) ? 1u : 0u) | ((m & 0xff00u) ? 2u : 0u) | ((m & 0xff0000u) ? 4u : 0u) | ((m & 0xff000000u) ? 8u : 0u);
          }
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            t[e] *= ((kb >> e) & 1u) ? keep_scale : 0.f;
            if (COLSUM) dbs[i][e] += t[e];
          }
          Fl4<Tb>::st(db_out + row * A.db_stride + 4 * qi, t);
        }
      }
    }
  }
  constexpr int npass = COLSUM ? 3 : 2;
#pragma unroll
  for (int pass = 0; pass < npass; ++pass) {
    if (pass) __syncthreads();
#pragma unroll
    for (int i = 0; i < NQ; ++i) {
      float4 v4;
      if (pass == 0) v4 = make_float4(dg[i][0], dg[i][1], dg[i][2], dg[i][3]);
      else if (pass == 1) v4 = make_float4(dbt[i][0], dbt[i][1], dbt[i][2], dbt[i][3]);
      else v4 = make_float4(dbs[COLSUM ? i : 0][0], dbs[COLSUM ? i : 0][1], dbs[COLSUM ? i : 0][2], dbs[COLSUM ? i : 0][3]);
      red[warp][lane + 32 * i] = v4;
    }
    __syncthreads();
    float* dst = pass == 0 ? A.dgamma_part : pass == 1 ? A.dbeta_part : A.dbsum_part;
    for (int qi = threadIdx.x; qi < nq; qi += blockDim.x) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int w = 0; w < kFlWarps; ++w) {
        const float4 t = red[w][qi];
        acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
      }
      *reinterpret_cast<float4*>(dst + (int64_t)blockIdx.x * C + 4 * qi) = acc;
    }
  }
}

static bool al(const void* p, size_t n) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) % n) == 0; }

template <typename T> static bool quad_al(const void* p, int64_t stride) {
  return p == nullptr || ((reinterpret_cast<uintptr_t>(p) % (4 * sizeof(T))) == 0 && (stride & 3) == 0);
}
template <typename Ta, typename Tb, typename Ty>
static bool quad_ok(const cm_add_ln_args& a, bool bwd) {
  static const bool off = getenv("CM_ADD_LN_NO_QUAD") != nullptr;      // A/B switch: the pair kernels
  if (off || (a.cols & 3)) return false;
  if (!al(a.gamma, 16) || !al(a.beta, 16) || !al(a.mask, 4)) return false;
  if (!bwd)
    return quad_al<Ta>(a.a, a.a_stride) && quad_al<Tb>(a.b, a.b_stride) && quad_al<Ta>(a.s, a.s_stride) && quad_al<Ty>(a.y, a.y_stride);
  return quad_al<Ta>(a.s, a.s_stride) && quad_al<Ty>(a.dy, a.dy_stride) && quad_al<Ta>(a.ds, a.ds_stride) &&
         quad_al<Ta>(a.da, a.da_stride) && quad_al<Tb>(a.db, a.db_stride) && al(a.dgamma_part, 16) && al(a.dbeta_part, 16) &&
         al(a.dbsum_part, 16);
}

// the staged backward (bulk copies): whole 16-byte units per row segment, 16-byte aligned rows
template <typename Ta, typename Ty>
static bool staged_ok(const cm_add_ln_args& a) {
  static const bool off = getenv("CM_ADD_LN_NO_STAGE") != nullptr;     // A/B switch: rows loaded straight into registers
  if (off || (a.cols & 7)) return false;
  auto ok = [](const void* p, int64_t stride, size_t es) {
    return p == nullptr || ((reinterpret_cast<uintptr_t>(p) & 15) == 0 && ((stride * (int64_t)es) & 15) == 0);
  };
  return ok(a.s, a.s_stride, sizeof(Ta)) && ok(a.ds, a.ds_stride, sizeof(Ta)) && ok(a.dy, a.dy_stride, sizeof(Ty));
}
template <typename Ta, typename Tb, typename Ty, bool ACT = false>
static int add_ln_bwd_staged(const cm_add_ln_args& a, int nblk, cudaStream_t st) {
  const int C = a.cols;
  const size_t row_bytes = (size_t)C * (2 * sizeof(Ta) + sizeof(Ty));
  // three stages (two rows of loads in flight per warp) while 3 CTAs / SM still fit beside the reduction buffer, else two
  const bool three = 3 * (3 * kFlWarps * row_bytes + 10 * 1024) <= 227 * 1024;
  const size_t smem = (three ? 3 : 2) * kFlWarps * row_bytes;
  const bool cs = !ACT && a.dbsum_part != nullptr && a.db != nullptr;
#define FL_BS(N, CS, NS)                                                                            \
  do {                                                                                              \
    auto kern = add_ln_bwd_q_kernel<Ta, Tb, Ty, N, CS, NS, ACT>;                                    \
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    if (e != cudaSuccess) return (int)e;                                                            \
    kern<<<nblk, 32 * kFlWarps, smem, st>>>(a);                                                     \
  } while (0)
#define FL_BS2(N)                                                                                   \
  do {                                                                                              \
    if (cs) { if (three) FL_BS(N, (!ACT), 3); else FL_BS(N, (!ACT), 2); }                           \
    else    { if (three) FL_BS(N, false, 3); else FL_BS(N, false, 2); }                             \
  } while (0)
  if (C <= 128) FL_BS2(1); else FL_BS2(2);
#undef FL_BS2
#undef FL_BS
  CM_LAUNCH_CHECK();
  return 0;
}

template <typename Ta, typename Tb, typename Ty>
static int add_ln_launch(const cm_add_ln_args& a, bool bwd, cudaStream_t st) {
  const int C = a.cols;
  if (quad_ok<Ta, Tb, Ty>(a, bwd)) {
    if (!bwd) {
      const unsigned grid = (unsigned)((a.rows + 2 * kFlWarps - 1) / (2 * kFlWarps));
#define FL_FQ(N) add_ln_fwd_q_kernel<Ta, Tb, Ty, N><<<grid, 32 * kFlWarps, 0, st>>>(a)
      if (C <= 128) FL_FQ(1); else if (C <= 256) FL_FQ(2); else if (C <= 512) FL_FQ(4); else FL_FQ(8);
#undef FL_FQ
    } else {
      const int nblk = cm_add_ln_num_part(a.rows, a.cols);
      if (C <= 256 && staged_ok<Ta, Ty>(a)) return add_ln_bwd_staged<Ta, Tb, Ty>(a, nblk, st);
#define FL_BQ(N)                                                                                   \
  do {                                                                                             \
    if (a.dbsum_part != nullptr && a.db != nullptr)                                                \
      add_ln_bwd_q_kernel<Ta, Tb, Ty, N, true><<<nblk, 32 * kFlWarps, 0, st>>>(a);                 \
    else                                                                                           \
      add_ln_bwd_q_kernel<Ta, Tb, Ty, N, false><<<nblk, 32 * kFlWarps, 0, st>>>(a);                \
  } while (0)
      if (C <= 128) FL_BQ(1); else if (C <= 256) FL_BQ(2); else if (C <= 512) FL_BQ(4); else FL_BQ(8);
#undef FL_BQ
    }
    CM_LAUNCH_CHECK();
    return 0;
  }
  if (a.dbsum_part != nullptr) return CM_ERR_UNSUPPORTED;     // the column sums of db exist in the quad kernels only
  if (!bwd) {
    const unsigned grid = (unsigned)((a.rows + 2 * kFlWarps - 1) / (2 * kFlWarps));
#define FL_F(N) add_ln_fwd_kernel<Ta, Tb, Ty, N><<<grid, 32 * kFlWarps, 0, st>>>(a)
    if (C <= 192) FL_F(3); else if (C <= 256) FL_F(4); else if (C <= 512) FL_F(8); else FL_F(16);
#undef FL_F
  } else {
    const int nblk = cm_add_ln_num_part(a.rows, a.cols);
#define FL_B(N) add_ln_bwd_kernel<Ta, Tb, Ty, N><<<nblk, 32 * kFlWarps, 0, st>>>(a)
    if (C <= 192) FL_B(3); else if (C <= 256) FL_B(4); else if (C <= 512) FL_B(8); else FL_B(16);
#undef FL_B
  }
  CM_LAUNCH_CHECK();
  return 0;
}

static int add_ln_dispatch(const cm_add_ln_args& a, bool bwd, cudaStream_t st) {
#define GO(TA, TB, TY) return add_ln_launch<TA, TB, TY>(a, bwd, st)
  const int bd = a.b ? a.b_dtype : (a.db ? a.b_dtype : a.a_dtype == CM_F32 ? CM_F32 : CM_BF16);
  if (a.a_dtype == CM_F32 && bd == CM_BF16 && a.y_dtype == CM_BF16) GO(float, __nv_bfloat16, __nv_bfloat16);
  if (a.a_dtype == CM_F32 && bd == CM_BF16 && a.y_dtype == CM_F32) GO(float, __nv_bfloat16, float);
  if (a.a_dtype == CM_F32 && bd == CM_F32 && a.y_dtype == CM_F32) GO(float, float, float);
  if (a.a_dtype == CM_BF16 && bd == CM_BF16 && a.y_dtype == CM_BF16) GO(__nv_bfloat16, __nv_bfloat16, __nv_bfloat16);
  if (a.a_dtype == CM_BF16 && bd == CM_BF16 && a.y_dtype == CM_F32) GO(__nv_bfloat16, __nv_bfloat16, float);
#undef GO
  return CM_ERR_UNSUPPORTED;
}

// cm_layernorm_bwd through the quad kernels above (layernorm.cu): LayerNorm backward is cm_add_ln_bwd without ds / db.  `a`
// carries s = x, da = dx, alpha = 1, p_drop = 0; `act`: the forward had the GELU epilogue (beta needed).  Any grid is valid
// (persistent row loop); nblk = the partial rows the caller allocated.  CM_ERR_UNSUPPORTED: geometry not covered here.
template <typename Ta, typename Ty, bool ACT>
static int ln_bwd_routed_t(const cm_add_ln_args& a, int nblk, cudaStream_t st) {
  using Tb = __nv_bfloat16;                                   // no b / db in this use
  if (!quad_ok<Ta, Tb, Ty>(a, true)) return CM_ERR_UNSUPPORTED;
  if (ACT && !al(a.beta, 16)) return CM_ERR_UNSUPPORTED;
  const int C = a.cols;
  if (C <= 256 && staged_ok<Ta, Ty>(a)) return add_ln_bwd_staged<Ta, Tb, Ty, ACT>(a, nblk, st);
#define FL_BR(N) add_ln_bwd_q_kernel<Ta, Tb, Ty, N, false, 0, ACT><<<nblk, 32 * kFlWarps, 0, st>>>(a)
  if (C <= 128) FL_BR(1); else if (C <= 256) FL_BR(2); else if (C <= 512) FL_BR(4); else FL_BR(8);
#undef FL_BR
  CM_LAUNCH_CHECK();
  return 0;
}
int ln_bwd_routed(const cm_add_ln_args& a, int nblk, bool act, cudaStream_t st) {
#define GO(TA, TY) return act ? ln_bwd_routed_t<TA, TY, true>(a, nblk, st) : ln_bwd_routed_t<TA, TY, false>(a, nblk, st)
  if (a.a_dtype == CM_F32 && a.y_dtype == CM_BF16) GO(float, __nv_bfloat16);
  if (a.a_dtype == CM_F32 && a.y_dtype == CM_F32) GO(float, float);
  if (a.a_dtype == CM_BF16 && a.y_dtype == CM_BF16) GO(__nv_bfloat16, __nv_bfloat16);
  if (a.a_dtype == CM_BF16 && a.y_dtype == CM_F32) GO(__nv_bfloat16, float);
#undef GO
  return CM_ERR_UNSUPPORTED;
}

}  // namespace cm

static int add_ln_common_ok(const cm_add_ln_args* a) {
  if (!a || a->rows <= 0 || a->cols <= 0 || !a->mean || !a->rstd) return CM_ERR_BAD_ARG;
  if (a->cols > 1024 || (a->cols & 1)) return CM_ERR_UNSUPPORTED;
  if (a->p_drop < 0.f || a->p_drop >= 1.f) return CM_ERR_BAD_ARG;
  if (!cm::al(a->gamma, 8) || !cm::al(a->beta, 8) || !cm::al(a->mask, 2)) return CM_ERR_UNSUPPORTED;
  return 0;
}

extern "C" int cm_add_ln_fwd(const cm_add_ln_args* a, void* stream) {
  if (int rc = add_ln_common_ok(a)) return rc;
  if (!a->a || !a->y) return CM_ERR_BAD_ARG;
  if (a->mask != nullptr && (a->b == nullptr || a->p_drop <= 0.f)) return CM_ERR_BAD_ARG;
  if (a->p_drop > 0.f && a->b != nullptr && a->mask == nullptr && a->key == nullptr) return CM_ERR_BAD_ARG;   // nothing for backward
  if (!cm::al(a->a, 8) || !cm::al(a->b, 4) || !cm::al(a->s, 8) || !cm::al(a->y, 4)) return CM_ERR_UNSUPPORTED;
  if ((a->a_stride | a->b_stride | a->s_stride | a->y_stride) & 1) return CM_ERR_UNSUPPORTED;
  return cm::add_ln_dispatch(*a, false, static_cast<cudaStream_t>(stream));
}

extern "C" int cm_add_ln_bwd(const cm_add_ln_args* a, void* stream) {
  if (int rc = add_ln_common_ok(a)) return rc;
  if (!a->s || !a->dy || !a->da || !a->dgamma_part || !a->dbeta_part) return CM_ERR_BAD_ARG;
  if (a->p_drop > 0.f && a->db != nullptr && a->mask == nullptr && a->key == nullptr) return CM_ERR_BAD_ARG;
  if (!cm::al(a->s, 8) || !cm::al(a->dy, 4) || !cm::al(a->ds, 8) || !cm::al(a->da, 8) || !cm::al(a->db, 4) ||
      !cm::al(a->dgamma_part, 8) || !cm::al(a->dbeta_part, 8))
    return CM_ERR_UNSUPPORTED;
  if ((a->s_stride | a->dy_stride | a->ds_stride | a->da_stride | a->db_stride) & 1) return CM_ERR_UNSUPPORTED;
  return cm::add_ln_dispatch(*a, true, static_cast<cudaStream_t>(stream));
}

extern "C" int cm_add_ln_dbsum_supported(int32_t cols, int64_t min_stride) {
  return (cols > 0 && cols <= 1024 && (cols & 3) == 0 && (min_stride & 3) == 0 && getenv("CM_ADD_LN_NO_QUAD") == nullptr) ? 1 : 0;
}

// partial rows (= CTAs) of cm_add_ln_bwd: one wave of persistent CTAs (8 warps each) at the residency the register budget of
// the row width allows - 3 per SM up to 256 columns (80 registers, no spills; measured 37.6 us against 42.1 us for 4 CTAs
// at 64 registers with spills, 32064 x 256), 2 up to 512, 1 beyond
extern "C" int cm_add_ln_num_part(int64_t rows, int32_t cols) {
  const int64_t need = (rows + cm::kFlWarps - 1) / cm::kFlWarps;
  const int64_t cap = 148 * (cols <= 256 ? 3 : cols <= 512 ? 2 : 1);
  return (int)(need < cap ? (need < 1 ? 1 : need) : cap);
}
