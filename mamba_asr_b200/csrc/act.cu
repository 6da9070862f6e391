// GELU (exact, erf) + dropout in one pass, forward and backward, for sm_100a.
//
// the position-wise feed-forward modules of a ConMamba layer are
// Linear -> activation -> Dropout -> Linear (speechbrain PositionalwiseFeedForward; reference modules/Conmamba.py:595-621
// with activation = GELU from Transformer.py:740-751), the FLOP majority of the layer.  As torch ops the
// activation and the dropout are two passes over the (rows, d_ffn) tensor forward (9 bytes per element at bf16) and two
// backward (11 bytes); fused they are one each (5 and 7 bytes).  Eight elements per thread (one 16-byte access at 16-bit
// types), grid-stride over the flat tensor.  Dropout mask: the counter-based hash of fused_ln.cu, stored as one byte per
// element for backward.  Roof: HBM.
#include "common.cuh"

namespace cm {

__device__ __forceinline__ uint32_t act_mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}
__device__ __forceinline__ uint32_t act_key(const int64_t* seed, uint32_t call_id) {
  const uint64_t s = seed ? static_cast<uint64_t>(*seed) : 0x243F6A8885A308D3ull;
  return act_mix32(static_cast<uint32_t>(s) ^ act_mix32(static_cast<uint32_t>(s >> 32) + call_id * 0x9E3779B9u + 0x85EBCA6Bu));
}

template <typename T> struct Vec8;
template <> struct Vec8<float> {
  static __device__ __forceinline__ void ld(const float* p, float* o) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
  }
  static __device__ __forceinline__ void st(float* p, const float* v) {
    reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
    reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
  }
};
template <> struct Vec8<__nv_bfloat16> {
  static __device__ __forceinline__ void ld(const __nv_bfloat16* p, float* o) {
    const uint4 r = __ldg(reinterpret_cast<const uint4*>(p));
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) { o[2 * i] = __uint_as_float(w[i] << 16); o[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
  }
  static __device__ __forceinline__ void st(__nv_bfloat16* p, const float* v) {
    uint4 r;
    __nv_bfloat162 t;
    t = __floats2bfloat162_rn(v[0], v[1]); r.x = *reinterpret_cast<uint32_t*>(&t);
    t = __floats2bfloat162_rn(v[2], v[3]); r.y = *reinterpret_cast<uint32_t*>(&t);
    t = __floats2bfloat162_rn(v[4], v[5]); r.z = *reinterpret_cast<uint32_t*>(&t);
    t = __floats2bfloat162_rn(v[6], v[7]); r.w = *reinterpret_cast<uint32_t*>(&t);
    *reinterpret_cast<uint4*>(p) = r;
  }
};
template <> struct Vec8<__half> {
  static __device__ __forceinline__ void ld(const __half* p, float* o) {
    const uint4 r = __ldg(reinterpret_cast<const uint4*>(p));
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
      o[2 * i] = f.x; o[2 * i + 1] = f.y;
    }
  }
  static __device__ __forceinline__ void st(__half* p, const float* v) {
    uint4 r;
    __half2 t;
    t = __floats2half2_rn(v[0], v[1]); r.x = *reinterpret_cast<uint32_t*>(&t);
    t = __floats2half2_rn(v[2], v[3]); r.y = *reinterpret_cast<uint32_t*>(&t);
    t = __floats2half2_rn(v[4], v[5]); r.z = *reinterpret_cast<uint32_t*>(&t);
    t = __floats2half2_rn(v[6], v[7]); r.w = *reinterpret_cast<uint32_t*>(&t);
    *reinterpret_cast<uint4*>(p) = r;
  }
};

// gelu_parts / gelu_f / gelu_grad_f: common.cuh (shared with ln_act.cu)

// keep bits of 8 consecutive elements starting at flat index 8 * v: four hashes (16-bit samples) of the pair indices
// 4 v .. 4 v + 3; the high word of the index is mixed once
__device__ __forceinline__ uint32_t keep8(uint32_t key, int64_t v, uint32_t thr) {
  uint32_t bits = 0;
  const uint32_t hi = act_mix32(static_cast<uint32_t>(static_cast<uint64_t>(v) >> 30) + 0x68E31DA4u) ^ key;
  const uint32_t lo = static_cast<uint32_t>(v) << 2;
#pragma unroll
  for (int h = 0; h < 4; ++h) {
    const uint32_t r = act_mix32(((lo + h) * 0x9E3779B9u) ^ hi);
    bits |= ((r & 0xffffu) >= thr ? 1u : 0u) << (2 * h);
    bits |= ((r >> 16) >= thr ? 1u : 0u) << (2 * h + 1);
  }
  return bits;
}

template <typename T>
__global__ void __launch_bounds__(256) gelu_dropout_fwd_kernel(const T* __restrict__ x, T* __restrict__ y,
                                                               uint8_t* __restrict__ mask, int64_t n8, float p,
                                                               const int64_t* __restrict__ seed, uint32_t call_id,
                                                               uint32_t* __restrict__ key_out, bool drop,
                                                               uint8_t* __restrict__ bits) {
  const uint32_t thr = drop ? (uint32_t)(p * 65536.0f) : 0u;
  const float scale = drop ? 1.0f / (1.0f - p) : 1.0f;
  const uint32_t key = drop ? act_key(seed, call_id) : 0u;
  if (key_out != nullptr && blockIdx.x == 0 && threadIdx.x == 0) *key_out = key;   // backward regenerates the mask from it
  // two 16-byte chunks per thread and iteration, both loads issued before the first use: one chunk per thread keeps only
  // 32 KB per SM in flight at full occupancy, below the latency x bandwidth product of HBM3e
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < n8; v += 2 * stride) {
    const int64_t v1 = v + stride;
    const bool two = v1 < n8;
    float a[2][8];
    Vec8<T>::ld(x + 8 * v, a[0]);
    if (two) Vec8<T>::ld(x + 8 * v1, a[1]);
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      if (c == 1 && !two) break;
      const int64_t vv = c ? v1 : v;
      const uint32_t kb = drop ? keep8(key, vv, thr) : 0xffu;
#pragma unroll
      // the keep bit as a factor (scale or 0), not as a branch: written as a select around gelu_f ptxas put every element's
      // GELU under its own divergent branch (BSSY / BRA / BSYNC per element, no overlap between the 8 chains)
      for (int i = 0; i < 8; ++i) a[c][i] = (((kb >> i) & 1u) ? scale : 0.f) * gelu_f(a[c][i]);
      Vec8<T>::st(y + 8 * vv, a[c]);
      if (bits != nullptr) bits[vv] = (uint8_t)kb;          // one keep bit per element: what backward reads instead of re-hashing
      if (mask != nullptr) {
        uint2 m;
        m.x = (kb & 1u) | ((kb & 2u) << 7) | ((kb & 4u) << 14) | ((kb & 8u) << 21);
        m.y = ((kb >> 4) & 1u) | (((kb >> 4) & 2u) << 7) | (((kb >> 4) & 4u) << 14) | (((kb >> 4) & 8u) << 21);
        *reinterpret_cast<uint2*>(mask + 8 * vv) = m;
      }
    }
  }
}

// mask source of the backward: the byte mask the forward stored, or the forward's key (the same hash, regenerated)
// COLS8 > 0 (vec8 groups per row; 256 % COLS8 == 0 so a thread's eight columns never change across its grid-stride
// iterations): also the column sums of dx, one partial row per CTA (fixed order) - the bias gradient of the Linear that
// produced x, which would otherwise be one more pass over dx (cm_colsum).
template <typename T, bool COLSUM>
__global__ void __launch_bounds__(256) gelu_dropout_bwd_kernel(const T* __restrict__ x, const T* __restrict__ dy,
                                                               const uint8_t* __restrict__ mask,
                                                               const uint32_t* __restrict__ key_in, T* __restrict__ dx,
                                                               int64_t n8, float p, int cols8, float* __restrict__ cs_part,
                                                               const uint8_t* __restrict__ bits) {
  const bool drop = p > 0.f;
  const bool regen = drop && mask == nullptr && bits == nullptr;
  const float scale = drop ? 1.0f / (1.0f - p) : 1.0f;
  const uint32_t thr = drop ? (uint32_t)(p * 65536.0f) : 0u;
  const uint32_t key = regen ? __ldg(key_in) : 0u;
  float cs[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) cs[i] = 0.f;
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < n8; v += (int64_t)gridDim.x * blockDim.x) {
    float a[8], g[8];
    Vec8<T>::ld(x + 8 * v, a);
    Vec8<T>::ld(dy + 8 * v, g);
    uint32_t kb = 0xffu;
    if (drop && bits != nullptr) {
      kb = __ldg(bits + v);                // the forward's keep bits (1 byte per 8 elements): no hash in backward
    } else if (regen) {
      kb = keep8(key, v, thr);
    } else if (drop) {
      const uint2 m = __ldg(reinterpret_cast<const uint2*>(mask + 8 * v));
      kb = 0u;
#pragma unroll
      for (int i = 0; i < 8; ++i) kb |= ((((i < 4 ? m.x : m.y) >> (8 * (i & 3))) & 0xffu) ? 1u : 0u) << i;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = ((((kb >> i) & 1u) ? scale : 0.f) * g[i]) * gelu_grad_f(a[i]);   // factor, not branch
    Vec8<T>::st(dx + 8 * v, a);
    if (COLSUM) {
#pragma unroll
      for (int i = 0; i < 8; ++i) cs[i] += a[i];       // the fp32 values, before the store rounds them
    }
  }
  if (COLSUM) {
    __shared__ float4 part_s[256][2];
    part_s[threadIdx.x][0] = make_float4(cs[0], cs[1], cs[2], cs[3]);
    part_s[threadIdx.x][1] = make_float4(cs[4], cs[5], cs[6], cs[7]);
    __syncthreads();
    // column c <- threads t = c / 8 + j * cols8 (their vec index is congruent to c / 8 modulo cols8), increasing j
    float* out = cs_part + (int64_t)blockIdx.x * cols8 * 8;
    for (int c = threadIdx.x; c < cols8 * 8; c += 256) {
      float sum = 0.f;
      for (int t = c >> 3; t < 256; t += cols8) sum += reinterpret_cast<const float*>(&part_s[t][0])[c & 7];
      out[c] = sum;
    }
  }
}

static unsigned act_grid(int64_t n8) {
  const int64_t need = (n8 + 255) / 256;
  const int64_t cap = 148 * 8;
  return (unsigned)(need < cap ? (need < 1 ? 1 : need) : cap);
}

}  // namespace cm

static bool act_al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

static int act_fwd_launch(const void* x, void* y, uint8_t* mask, int64_t n, int32_t dtype, float p_drop, const int64_t* seed,
                          uint32_t call_id, uint32_t* key_out, bool drop, void* stream, uint8_t* bits = nullptr) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t n8 = n >> 3;
  const unsigned grid = cm::act_grid(n8);
  switch (dtype) {
    case CM_F32: cm::gelu_dropout_fwd_kernel<float><<<grid, 256, 0, st>>>(static_cast<const float*>(x), static_cast<float*>(y), mask, n8, p_drop, seed, call_id, key_out, drop, bits); break;
    case CM_BF16: cm::gelu_dropout_fwd_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(y), mask, n8, p_drop, seed, call_id, key_out, drop, bits); break;
    default: cm::gelu_dropout_fwd_kernel<__half><<<grid, 256, 0, st>>>(static_cast<const __half*>(x), static_cast<__half*>(y), mask, n8, p_drop, seed, call_id, key_out, drop, bits); break;
  }
  CM_LAUNCH_CHECK();
  return 0;
}

template <bool COLSUM>
static int act_bwd_launch(const void* x, const void* dy, const uint8_t* mask, const uint32_t* key, void* dx, int64_t n,
                          int32_t dtype, float p_drop, int cols8, float* cs_part, void* stream, const uint8_t* bits = nullptr) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t n8 = n >> 3;
  const unsigned grid = cm::act_grid(n8);
  switch (dtype) {
    case CM_F32: cm::gelu_dropout_bwd_kernel<float, COLSUM><<<grid, 256, 0, st>>>(static_cast<const float*>(x), static_cast<const float*>(dy), mask, key, static_cast<float*>(dx), n8, p_drop, cols8, cs_part, bits); break;
    case CM_BF16: cm::gelu_dropout_bwd_kernel<__nv_bfloat16, COLSUM><<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(dy), mask, key, static_cast<__nv_bfloat16*>(dx), n8, p_drop, cols8, cs_part, bits); break;
    default: cm::gelu_dropout_bwd_kernel<__half, COLSUM><<<grid, 256, 0, st>>>(static_cast<const __half*>(x), static_cast<const __half*>(dy), mask, key, static_cast<__half*>(dx), n8, p_drop, cols8, cs_part, bits); break;
  }
  CM_LAUNCH_CHECK();
  return 0;
}

// ---- GLU over the last dimension, forward and backward ------------------------------------------------------------------
// The ConMamba convolution module gates its pointwise-conv output before the depthwise conv (reference modules/Conmamba.py:
// 268-279: Conv1d(C, 2C, 1) -> nn.GLU(dim=1)); channel-last that is y[r, c] = h[r, c] * sigmoid(h[r, C + c]).  torch's
// glu / glu_backward kernels run it at 2.1 TB/s on B200 (23 / 36 us per layer at 32064 x 512 bf16); these are the same passes
// with 16-byte accesses, two chunks per thread in flight.  (Gating inside the depthwise-conv kernels was built and measured
// slower: they are FP32-issue-bound, profiles/r02_glu_in_dwconv_negative.txt.)  Roof: HBM; bytes per gated element: 3 s
// forward, 5 s backward.
namespace cm {

template <typename T>
__global__ void __launch_bounds__(256) glu_fwd_kernel(const T* __restrict__ h, T* __restrict__ y, int64_t rows, int32_t c8,
                                                      int64_t h_stride, int64_t y_stride) {
  const int64_t n = rows * c8, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += 2 * stride) {
    const int64_t j = i + stride;
    const bool two = j < n;
    const int64_t r0 = i / c8, r1 = two ? j / c8 : r0;
    const int k0 = (int)(i - r0 * c8) * 8, k1 = two ? (int)(j - r1 * c8) * 8 : k0;
    float a0[8], b0[8], a1[8], b1[8];
    Vec8<T>::ld(h + r0 * h_stride + k0, a0);
    Vec8<T>::ld(h + r0 * h_stride + 8 * c8 + k0, b0);
    if (two) { Vec8<T>::ld(h + r1 * h_stride + k1, a1); Vec8<T>::ld(h + r1 * h_stride + 8 * c8 + k1, b1); }
#pragma unroll
    for (int e = 0; e < 8; ++e) a0[e] *= sigmoidf_fast(b0[e]);
    Vec8<T>::st(y + r0 * y_stride + k0, a0);
    if (two) {
#pragma unroll
      for (int e = 0; e < 8; ++e) a1[e] *= sigmoidf_fast(b1[e]);
      Vec8<T>::st(y + r1 * y_stride + k1, a1);
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256) glu_bwd_kernel(const T* __restrict__ h, const T* __restrict__ dy, T* __restrict__ dh,
                                                      int64_t rows, int32_t c8, int64_t h_stride, int64_t dy_stride,
                                                      int64_t dh_stride) {
  const int64_t n = rows * c8, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int64_t r = i / c8;
    const int k = (int)(i - r * c8) * 8;
    float a[8], b[8], g[8];
    Vec8<T>::ld(h + r * h_stride + k, a);
    Vec8<T>::ld(h + r * h_stride + 8 * c8 + k, b);
    Vec8<T>::ld(dy + r * dy_stride + k, g);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float sg = sigmoidf_fast(b[e]);
      const float gs = g[e] * sg;
      b[e] = gs * a[e] * (1.0f - sg);
      a[e] = gs;
    }
    Vec8<T>::st(dh + r * dh_stride + k, a);
    Vec8<T>::st(dh + r * dh_stride + 8 * c8 + k, b);
  }
}

template <typename T>
static int glu_launch(const void* h, const void* dy, void* out, int64_t rows, int32_t dim, int64_t h_stride, int64_t dy_stride,
                      int64_t out_stride, bool bwd, void* stream) {
  const int c8 = dim / 8;
  const int64_t n = rows * c8;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (bwd) {
    const unsigned grid = (unsigned)((n + 255) / 256 < 148 * 64 ? (n + 255) / 256 : 148 * 64);
    glu_bwd_kernel<T><<<grid, 256, 0, st>>>(static_cast<const T*>(h), static_cast<const T*>(dy), static_cast<T*>(out), rows, c8,
                                            h_stride, dy_stride, out_stride);
  } else {
    const int64_t want = (n + 511) / 512;
    const unsigned grid = (unsigned)(want < 148 * 64 ? (want > 0 ? want : 1) : 148 * 64);
    glu_fwd_kernel<T><<<grid, 256, 0, st>>>(static_cast<const T*>(h), static_cast<T*>(out), rows, c8, h_stride, out_stride);
  }
  CM_LAUNCH_CHECK();
  return 0;
}

static int glu_check(const void* h, const void* other, const void* out, int64_t rows, int32_t dim, int64_t s0, int64_t s1,
                     int64_t s2, int32_t dtype, bool bwd) {
  if (!h || !out || (bwd && !other) || rows <= 0 || dim <= 0 || !dtype_ok(dtype)) return CM_ERR_BAD_ARG;
  const int es = dtype == CM_F32 ? 4 : 2;
  // 16-byte chunks of 8 elements: with fp32 the chunk is two float4 (32-byte alignment of the 8-element group is not needed)
  if ((dim & 7) || ((s0 | s1 | s2) & 7) || s0 < 2 * (int64_t)dim) return CM_ERR_UNSUPPORTED;
  auto al = [&](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if (!al(h) || !al(out) || (bwd && !al(other))) return CM_ERR_UNSUPPORTED;
  (void)es;
  return 0;
}

}  // namespace cm

extern "C" int cm_glu_fwd(const void* h, void* y, int64_t rows, int32_t dim, int64_t h_stride, int64_t y_stride, int32_t dtype,
                          void* stream) {
  if (int e = cm::glu_check(h, nullptr, y, rows, dim, h_stride, y_stride, 8, dtype, false)) return e;
  if (y_stride < dim) return CM_ERR_BAD_ARG;
  switch (dtype) {
    case CM_F32: return cm::glu_launch<float>(h, nullptr, y, rows, dim, h_stride, 0, y_stride, false, stream);
    case CM_BF16: return cm::glu_launch<__nv_bfloat16>(h, nullptr, y, rows, dim, h_stride, 0, y_stride, false, stream);
    default: return cm::glu_launch<__half>(h, nullptr, y, rows, dim, h_stride, 0, y_stride, false, stream);
  }
}

extern "C" int cm_glu_bwd(const void* h, const void* dy, void* dh, int64_t rows, int32_t dim, int64_t h_stride, int64_t dy_stride,
                          int64_t dh_stride, int32_t dtype, void* stream) {
  if (int e = cm::glu_check(h, dy, dh, rows, dim, h_stride, dy_stride, dh_stride, dtype, true)) return e;
  if (dy_stride < dim || dh_stride < 2 * (int64_t)dim) return CM_ERR_BAD_ARG;
  switch (dtype) {
    case CM_F32: return cm::glu_launch<float>(h, dy, dh, rows, dim, h_stride, dy_stride, dh_stride, true, stream);
    case CM_BF16: return cm::glu_launch<__nv_bfloat16>(h, dy, dh, rows, dim, h_stride, dy_stride, dh_stride, true, stream);
    default: return cm::glu_launch<__half>(h, dy, dh, rows, dim, h_stride, dy_stride, dh_stride, true, stream);
  }
}

extern "C" int cm_gelu_dropout_fwd(const void* x, void* y, uint8_t* mask, int64_t n, int32_t dtype, float p_drop,
                                   const int64_t* seed, uint32_t call_id, void* stream) {
  if (!x || !y || n <= 0 || !cm::dtype_ok(dtype) || p_drop < 0.f || p_drop >= 1.f) return CM_ERR_BAD_ARG;
  if ((n & 7) || !act_al16(x) || !act_al16(y) || (mask && (reinterpret_cast<uintptr_t>(mask) & 7))) return CM_ERR_UNSUPPORTED;
  if (mask && p_drop <= 0.f) return CM_ERR_BAD_ARG;
  return act_fwd_launch(x, y, mask, n, dtype, p_drop, seed, call_id, nullptr, mask != nullptr, stream);
}

extern "C" int cm_gelu_dropout_bwd(const void* x, const void* dy, const uint8_t* mask, void* dx, int64_t n, int32_t dtype,
                                   float p_drop, void* stream) {
  if (!x || !dy || !dx || n <= 0 || !cm::dtype_ok(dtype) || p_drop < 0.f || p_drop >= 1.f) return CM_ERR_BAD_ARG;
  if ((n & 7) || !act_al16(x) || !act_al16(dy) || !act_al16(dx) || (mask && (reinterpret_cast<uintptr_t>(mask) & 7)))
    return CM_ERR_UNSUPPORTED;
  return act_bwd_launch<false>(x, dy, mask, nullptr, dx, n, dtype, mask ? p_drop : 0.f, 0, nullptr, stream);
}

// struct form: the mask regenerated in backward from the forward's key (nothing stored), column sums of dx
static bool act_cols_ok(int64_t n, int32_t cols) {
  return cols > 0 && (cols & 7) == 0 && 256 % (cols >> 3) == 0 && n % cols == 0;
}

extern "C" int cm_act_colsum_supported(int64_t n, int32_t cols) { return act_cols_ok(n, cols) ? 1 : 0; }

extern "C" int cm_act_num_part(int64_t n) { return n > 0 ? (int)cm::act_grid(n >> 3) : 1; }

extern "C" int cm_gelu_dropout_fwd_v2(const cm_act_args* a, void* stream) {
  if (!a || !a->x || !a->y || a->n <= 0 || !cm::dtype_ok(a->dtype) || a->p_drop < 0.f || a->p_drop >= 1.f) return CM_ERR_BAD_ARG;
  if ((a->n & 7) || !act_al16(a->x) || !act_al16(a->y) || (a->mask && (reinterpret_cast<uintptr_t>(a->mask) & 7)))
    return CM_ERR_UNSUPPORTED;
  if (a->p_drop > 0.f && !a->mask && !a->key && !a->keep_bits) return CM_ERR_BAD_ARG;   // backward could not rebuild the mask
  return act_fwd_launch(a->x, a->y, a->p_drop > 0.f ? a->mask : nullptr, a->n, a->dtype, a->p_drop, a->seed, a->call_id,
                        a->key, a->p_drop > 0.f, stream, a->p_drop > 0.f ? a->keep_bits : nullptr);
}

extern "C" int cm_gelu_dropout_bwd_v2(const cm_act_args* a, void* stream) {
  if (!a || !a->x || !a->dy || !a->dx || a->n <= 0 || !cm::dtype_ok(a->dtype) || a->p_drop < 0.f || a->p_drop >= 1.f)
    return CM_ERR_BAD_ARG;
  if ((a->n & 7) || !act_al16(a->x) || !act_al16(a->dy) || !act_al16(a->dx) || (a->mask && (reinterpret_cast<uintptr_t>(a->mask) & 7)))
    return CM_ERR_UNSUPPORTED;
  if (a->p_drop > 0.f && !a->mask && !a->key && !a->keep_bits) return CM_ERR_BAD_ARG;
  const uint8_t* bits = a->p_drop > 0.f ? a->keep_bits : nullptr;
  if (a->cols > 0) {
    if (!a->colsum_part) return CM_ERR_BAD_ARG;
    if (!act_cols_ok(a->n, a->cols)) return CM_ERR_UNSUPPORTED;
    return act_bwd_launch<true>(a->x, a->dy, a->mask, a->key, a->dx, a->n, a->dtype, a->p_drop, a->cols >> 3, a->colsum_part, stream, bits);
  }
  return act_bwd_launch<false>(a->x, a->dy, a->mask, a->key, a->dx, a->n, a->dtype, a->p_drop, 0, nullptr, stream, bits);
}
