// Selective-scan forward, warpgroup-specialised kernel for sm_100a ("wg" path, round 2): the default for channel-last
// 16-state launches with TMA-addressable operands and no time windows.
//
// Mathematics, bidirectional stash/combine protocol and checkpoint contract are those of scan_fwd.cu (see its header;
// reference semantics: modules/mamba/selective_scan_interface.py:106-157 and modules/mamba/bimamba.py:223-253).
//
// The round-1 default (scan_fwd_sp.cu) was bound by issue slots, and more than half of the instructions it issued belonged
// to its IO role (64-bit address arithmetic, predicates and branches around per-thread global loads; a reduction of the
// per-lane partial outputs through shared memory; two mbarrier rings).  This kernel keeps the lane mapping (a lane owns
// states 4m..4m+3 of two adjacent channels, state pairs packed in 64-bit registers) and changes everything around it:
//
//   * CTA = two warpgroups with their own register budgets (setmaxnreg): 4 recurrence warps (88 registers) and 4 service
//     warps (40); 4 CTAs per SM = 16 recurrence warps (12 before).  A CTA owns two GROUPS of 32 channels: the two time
//     directions of one channel block (bidirectional launch) or two channel blocks (unidirectional).  Per group: 2 recurrence
//     warps, 1 producer warp, 1 TMA warp.
//   * ALL GLOBAL LOADS ARE TMA TILES (cp.async.bulk.tensor, box = 32 channels x 16 steps: u, delta, B|C, and in the gated
//     ranges z and the partner direction's stash), issued by the TMA warp up to three tiles ahead; rows outside [0, L) and
//     channels outside [0, dim) are zero-filled by the hardware.  The producer warp turns a raw tile into fp32 operand rows
//     (softplus, dt*u; B|C widened) - no address arithmetic, no predicate on any load.
//   * THE SUM OVER STATES y = sum_n C_n h_n (4 lanes of a channel pair) IS DONE BY THE TENSOR CORE: mma.sync m16n8k8 TF32
//     with the lane's partial sum (split into two TF32 terms for fp32 I/O) as the A operand and a one-hot column selector
//     as B, accumulated over 8 steps - afterwards lane (pair, m) holds y of steps m, m+4, m+8, m+12 of its tile for both
//     channels and finishes them itself: D*u skip, stash / combine with the partner direction / out_scale*silu(z), two
//     channels per store.  No partial-sum buffer, no y ring, no second pair of barriers.
//   * bidirectional fusion as before: range 0 stashes pre-gate sums in `out`, the recurrence warps of both directions
//     arrive at a named barrier, the TMA warps wait on it before they fetch the first tile of range 1 (which reads the
//     partner's stash back through L2); no flip, no second output tensor, no add kernel.
//
// Requirements (else cm_scan_fwd falls through to scan_fwd_sp.cu / scan_fwd_cl.cu / scan_fwd.cu): unit channel stride,
// dstate == 16, dim a multiple of 32, variable B/C as one 32-element B|C row per step, 16-byte aligned rows, no time
// windows (those launches stay on scan_fwd_sp.cu).
#include <climits>
#include <cstddef>
#include <cstdlib>

#include "common.cuh"
#include "sp_common.cuh"
#include "tma.cuh"

namespace cm {
namespace wgf {

using cm::sp::Pair;

constexpr int kT = 16;                    // steps per tile
constexpr int kCH = 32;                   // channels per group
constexpr int kNP = kCH / 2;              // channel pairs per group
constexpr int kRS = 3;                    // raw (TMA) ring depth
#ifndef CM_FWDWG_RREG
#define CM_FWDWG_RREG 88
#endif
#ifndef CM_FWDWG_IREG
#define CM_FWDWG_IREG 40
#endif
#ifndef CM_FWDWG_MINB
#define CM_FWDWG_MINB 4
#endif

struct FwdDir {
  CUtensorMap m_u, m_dl, m_bc;
  char *out, *pre;                         // byte pointers at (batch 0, channel 0, PROCESSED step 0)
  int64_t out_sb, pre_sb;                  // batch strides (bytes)
  int32_t out_ss, pre_ss;                  // bytes per processed step (signed)
  int32_t s1, reverse;
  const float* A;
  int64_t A_sd, A_sn;
  const float *Dskip, *bias;
  float* ckpt;
  int64_t ckpt_sb, ckpt_sd;
  float* last;
  int64_t ls_sb, ls_sd, ls_sn;
};
struct alignas(64) FwdParams {
  CUtensorMap m_z, m_st;                   // gate z; `out` read back (the partner direction's stash)
  FwdDir dir[2];
  int32_t L, dim, ndir, has_z, has_pre, pad0;
  float scale;
  int32_t pad1;
};

template <typename T>
struct alignas(128) Raw {                  // one TMA stage of a group: [step (ascending time)][channel]
  T u[kT][kCH], dl[kT][kCH], z[kT][kCH], st[kT][kCH], bc[kT][32];
};
struct alignas(128) Ops {                  // one operand slot of a group (fp32, processed order)
  float4 dd[kT][kNP];                      // (dt0, dt1, dt0*u0, dt1*u1)
  float bc[kT][32];                        // B[0..15] | C[0..15]
};
template <typename T>
struct alignas(128) GroupSmem {
  Raw<T> raw[kRS];
  Ops ops[2];
  uint64_t raw_full[kRS], raw_empty[kRS], in_full[2], in_empty[2];
};

__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WGF_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WGF_DONE;\n\t"
      "bra WGF_WAIT;\n\t"
      "WGF_DONE:\n\t}"
      ::"r"(tma::smem_u32(b)), "r"(parity) : "memory");
}
// with a suspend-time hint: the service warps wait whole tiles (see scan_bwd_wg.cu)
__device__ __forceinline__ void mbar_wait_long(uint64_t* b, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WGFL_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
      "@p bra WGFL_DONE;\n\t"
      "bra WGFL_WAIT;\n\t"
      "WGFL_DONE:\n\t}"
      ::"r"(tma::smem_u32(b)), "r"(parity), "r"(100000u) : "memory");
}
__device__ __forceinline__ void warp_arrive(uint64_t* b, int lane) {
  __syncwarp();
  if (lane == 0) tma::mbar_arrive(b);
}

// D += A * B, m16n8k8 TF32 (A row-major 16x8, B col-major 8x8 with both halves equal, fp32 accumulate)
__device__ __forceinline__ void mma_tf32(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b) {
  asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%8}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b));
}
// Sum of x (rows g) and y (rows g + 8) over the four lanes of a quad, added into the selected column of the accumulator.
// PRECISE: two TF32 terms per value (upper 19 bits + remainder, |error| < 2^-21); else one term rounded to nearest.
template <bool PRECISE>
__device__ __forceinline__ void quad_sum_mma(float (&acc)[4], float x, float y, uint32_t sel) {
  if (PRECISE) {
    const uint32_t hx = __float_as_uint(x) & 0xffffe000u, hy = __float_as_uint(y) & 0xffffe000u;
    const float lx = x - __uint_as_float(hx), ly = y - __uint_as_float(hy);
    mma_tf32(acc, hx, hy, __float_as_uint(lx), __float_as_uint(ly), sel);
  } else {
    mma_tf32(acc, __float_as_uint(x) + 0x1000u, __float_as_uint(y) + 0x1000u, 0u, 0u, sel);
  }
}

template <typename T> struct SmemPair;    // two adjacent elements of a raw tile -> float2
template <> struct SmemPair<float> {
  static __device__ __forceinline__ float2 ld(const float* p) { return *reinterpret_cast<const float2*>(p); }
};
template <> struct SmemPair<__nv_bfloat16> {
  static __device__ __forceinline__ float2 ld(const __nv_bfloat16* p) {
    const uint32_t r = *reinterpret_cast<const uint32_t*>(p);
    return make_float2(__uint_as_float(r << 16), __uint_as_float(r & 0xffff0000u));
  }
};
template <> struct SmemPair<__half> {
  static __device__ __forceinline__ float2 ld(const __half* p) { return __half22float2(*reinterpret_cast<const __half2*>(p)); }
};
template <typename T> struct SmemOct;     // eight adjacent elements of a raw tile -> two float4
template <> struct SmemOct<float> {
  static __device__ __forceinline__ void ld(const float* p, float4* a, float4* b) {
    *a = *reinterpret_cast<const float4*>(p); *b = *reinterpret_cast<const float4*>(p + 4);
  }
};
template <> struct SmemOct<__nv_bfloat16> {
  static __device__ __forceinline__ void ld(const __nv_bfloat16* p, float4* a, float4* b) {
    const uint4 r = *reinterpret_cast<const uint4*>(p);
    *a = make_float4(__uint_as_float(r.x << 16), __uint_as_float(r.x & 0xffff0000u), __uint_as_float(r.y << 16),
                     __uint_as_float(r.y & 0xffff0000u));
    *b = make_float4(__uint_as_float(r.z << 16), __uint_as_float(r.z & 0xffff0000u), __uint_as_float(r.w << 16),
                     __uint_as_float(r.w & 0xffff0000u));
  }
};
template <> struct SmemOct<__half> {
  static __device__ __forceinline__ void ld(const __half* p, float4* a, float4* b) {
    const float2 v0 = __half22float2(*reinterpret_cast<const __half2*>(p)), v1 = __half22float2(*reinterpret_cast<const __half2*>(p + 2));
    const float2 v2 = __half22float2(*reinterpret_cast<const __half2*>(p + 4)), v3 = __half22float2(*reinterpret_cast<const __half2*>(p + 6));
    *a = make_float4(v0.x, v0.y, v1.x, v1.y); *b = make_float4(v2.x, v2.y, v3.x, v3.y);
  }
};

enum { FM_UNI = 0, FM_STASH = 1, FM_COMBINE = 2 };

// what a group works on
struct Group {
  int dir, c_base;
};
__device__ __forceinline__ Group group_of(const FwdParams& P, int gi) {
  Group g;
  g.dir = P.ndir == 2 ? gi : 0;
  g.c_base = (P.ndir == 2 ? blockIdx.x : 2 * blockIdx.x + gi) * kCH;
  return g;
}
// processed-step range r of a direction and its mode
__device__ __forceinline__ void range_of(const FwdParams& P, const FwdDir& d, int r, int* s_begin, int* s_end, int* mode) {
  if (P.ndir == 1) { *s_begin = 0; *s_end = P.L; *mode = FM_UNI; return; }
  *s_begin = r == 0 ? 0 : d.s1;
  *s_end = r == 0 ? d.s1 : P.L;
  *mode = r == 0 ? FM_STASH : FM_COMBINE;
}

// ---- recurrence warps (2 per group) ------------------------------------------------------------------------------------
template <int V> struct IntC { static constexpr int value = V; };

template <typename T>
__device__ __forceinline__ void scan_role(const FwdParams& P, GroupSmem<T>& S, const int gt, const int gi) {
  using P2 = Pair<T>;
  constexpr int ES = (int)sizeof(T);
  constexpr int RB = kCH * ES;               // bytes per row of a raw tile
  constexpr bool PRECISE = sizeof(T) == 4;
  const Group grp = group_of(P, gi);
  const FwdDir& d = P.dir[grp.dir];
  const int warp = gt >> 5, lane = gt & 31;
  const int b = blockIdx.y;
  const int g = lane >> 2, m = lane & 3;
  const float gf = (float)g;
  const int lp = warp * 8 + g;              // channel pair inside the group
  const int c0 = grp.c_base + 2 * lp;
  const bool ch_ok = c0 < P.dim;
  const bool rev = d.reverse != 0;
  // [channel c][state pair j]: states 4m + 2j, 4m + 2j + 1 of channel c0 + c
  float2 kA[2][2], h[2][2];
  {
    const float* A0 = d.A + (int64_t)(ch_ok ? c0 : 0) * d.A_sd + (int64_t)(4 * m) * d.A_sn;
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const float* a = A0 + c * d.A_sd + (2 * j) * d.A_sn;
        kA[c][j] = make_float2(__ldg(a) * kLog2e, __ldg(a + d.A_sn) * kLog2e);
        h[c][j] = make_float2(0.f, 0.f);
      }
  }
  float2 Dsk = make_float2(0.f, 0.f);
  if (d.Dskip && ch_ok) Dsk = make_float2(__ldg(d.Dskip + c0), __ldg(d.Dskip + c0 + 1));
  // loop invariants pinned in registers (read through P.dir[dir] they would be re-fetched from the constant bank, with a
  // register index, at every use inside the unrolled tile)
  int out_ss = d.out_ss, pre_ss = d.pre_ss;
  // output rows of this lane: processed step m of the tile, then every fourth step
  char* po = d.out + b * d.out_sb + (int64_t)c0 * ES + (int64_t)m * out_ss;
  char* ppre = (P.has_pre && d.pre) ? d.pre + b * d.pre_sb + (int64_t)c0 * ES + (int64_t)m * pre_ss : nullptr;
  float* ckq = (d.ckpt && ch_ok) ? d.ckpt + b * d.ckpt_sb + (int64_t)c0 * d.ckpt_sd + 4 * m : nullptr;   // next checkpoint
  int64_t ckpt_sd = d.ckpt_sd;
  float scale = P.scale;
  // this lane's element in a raw tile: row of processed step m, then every fourth processed step
  int rrow = ((rev ? kT - 1 - m : m) * kCH + 2 * lp) * ES;
  int rstep = (rev ? -4 : 4) * RB;
  asm volatile("" : "+r"(out_ss), "+r"(pre_ss), "+r"(rrow), "+r"(rstep), "+f"(scale), "+l"(po), "+l"(ppre), "+l"(ckq), "+l"(ckpt_sd));

  int it = 0;                               // tile counter over both ranges: ops slot = it & 1, raw stage = it % kRS
  int stage = 0;
  uint32_t stage_par = 0;

  // one range of processed steps [s_begin, s_end) in a compile-time mode
  auto run_range = [&](auto mode_c, auto hasz_c, const int s_begin, const int s_end) {
    constexpr int MODE = decltype(mode_c)::value;
    constexpr bool HAS_Z = decltype(hasz_c)::value != 0 && MODE != FM_STASH;
    const int nst = s_end - s_begin;
    const int ntile = nst > 0 ? cm_ceil_div(nst, kT) : 0;
#pragma unroll 1
    for (int t = 0; t < ntile; ++t, ++it) {
      const int slot = it & 1;
      const int sb0 = s_begin + t * kT;
      const int nvalid = s_end - sb0;       // steps of this tile inside the range (may exceed kT)
      const Ops& O = S.ops[slot];
      mbar_wait(&S.in_full[slot], (it >> 1) & 1);
      const float4* ddb = &O.dd[0][lp];
      const float* bcb = &O.bc[0][4 * m];
      float Y[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
      for (int k = 0; k < kT; ++k) {
        if ((k % CM_SCAN_CKPT_STEPS) == 0 && ckq != nullptr && k < nvalid) {
          *reinterpret_cast<float4*>(ckq) = make_float4(h[0][0].x, h[0][0].y, h[0][1].x, h[0][1].y);
          *reinterpret_cast<float4*>(ckq + ckpt_sd) = make_float4(h[1][0].x, h[1][0].y, h[1][1].x, h[1][1].y);
          ckq += 16;
        }
        const float4 dd = ddb[k * kNP];
        const float4 bb = *reinterpret_cast<const float4*>(bcb + k * 32);
        const float4 cc = *reinterpret_cast<const float4*>(bcb + k * 32 + 16);
        const float2 Bp[2] = {make_float2(bb.x, bb.y), make_float2(bb.z, bb.w)};
        const float2 Cp[2] = {make_float2(cc.x, cc.y), make_float2(cc.z, cc.w)};
        float yc[2];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const float dtc = c ? dd.y : dd.x, duc = c ? dd.w : dd.z;
          const float2 dt2 = make_float2(dtc, dtc), du2 = make_float2(duc, duc);
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const float2 x = fmul2(dt2, kA[c][j]);
            const float2 a = make_float2(ex2(x.x), ex2(x.y));
            h[c][j] = ffma2(a, h[c][j], fmul2(du2, Bp[j]));
          }
          const float2 pp = ffma2(Cp[1], h[c][1], fmul2(Cp[0], h[c][0]));
          yc[c] = pp.x + pp.y;
        }
        uint32_t sel;                        // 1.0f in the lanes that hold column 2*(k & 3) + ((k >> 2) & 1) of B
        asm("set.eq.f32.f32 %0, %1, %2;" : "=r"(sel) : "f"(gf), "f"((float)(2 * (k & 3) + ((k >> 2) & 1))));
        quad_sum_mma<PRECISE>(Y[k >> 3], yc[0], yc[1], sel);
      }
      warp_arrive(&S.in_empty[slot], lane);   // the operand slot may be refilled
      // ---- steps m, m+4, m+8, m+12 of the tile: skip term, stash / combine / gate, store
      mbar_wait(&S.raw_full[stage], stage_par);            // (complete long ago: makes the TMA writes visible to this thread)
      const char* rb = reinterpret_cast<const char*>(&S.raw[stage]) + rrow;
      char* pos = po + (int64_t)sb0 * out_ss;
      char* pps = ppre != nullptr ? ppre + (int64_t)sb0 * pre_ss : nullptr;
      auto unit = [&](auto e_c, const bool ok) {
        constexpr int e = decltype(e_c)::value;
        const char* rr = rb + e * rstep;
        const float2 u2 = SmemPair<T>::ld(reinterpret_cast<const T*>(rr + offsetof(Raw<T>, u)));
        const float2 y = ffma2(Dsk, u2, make_float2(Y[e >> 1][e & 1], Y[e >> 1][2 + (e & 1)]));
        char* dsto = pos + (int64_t)(4 * e) * out_ss;
        if (MODE == FM_STASH) {
          if (ok) P2::st(dsto, y);
        } else {
          float2 tot = y;
          if (MODE == FM_COMBINE) tot = fadd2(tot, SmemPair<T>::ld(reinterpret_cast<const T*>(rr + offsetof(Raw<T>, st))));
          float2 val = fmul2(tot, make_float2(scale, scale));
          if (HAS_Z) {
            const float2 zz = SmemPair<T>::ld(reinterpret_cast<const T*>(rr + offsetof(Raw<T>, z)));
            val.x *= zz.x * sigmoid_sel<PRECISE>(zz.x);
            val.y *= zz.y * sigmoid_sel<PRECISE>(zz.y);
          }
          if (ok) {
            if (pps != nullptr) P2::st(pps + (int64_t)(4 * e) * pre_ss, tot);
            P2::st(dsto, val);
          }
        }
      };
      if (ch_ok) {
        if (nvalid >= kT) {                  // full tile: no per-step predicate
          unit(IntC<0>{}, true); unit(IntC<1>{}, true); unit(IntC<2>{}, true); unit(IntC<3>{}, true);
        } else {
          unit(IntC<0>{}, m < nvalid); unit(IntC<1>{}, m + 4 < nvalid); unit(IntC<2>{}, m + 8 < nvalid); unit(IntC<3>{}, m + 12 < nvalid);
        }
      }
      warp_arrive(&S.raw_empty[stage], lane);
      if (++stage == kRS) { stage = 0; stage_par ^= 1; }
    }
  };

  if (P.ndir == 2) {
    if (P.has_z) { run_range(IntC<FM_STASH>{}, IntC<1>{}, 0, d.s1); } else { run_range(IntC<FM_STASH>{}, IntC<0>{}, 0, d.s1); }
    // this direction's stash of its half is complete: let the TMA warps fetch the partner's stash (range 1)
    __threadfence();
    tma::fence_proxy_async_all();
    asm volatile("bar.arrive 1, 192;" ::: "memory");
    if (P.has_z) run_range(IntC<FM_COMBINE>{}, IntC<1>{}, d.s1, P.L); else run_range(IntC<FM_COMBINE>{}, IntC<0>{}, d.s1, P.L);
  } else {
    if (P.has_z) run_range(IntC<FM_UNI>{}, IntC<1>{}, 0, P.L); else run_range(IntC<FM_UNI>{}, IntC<0>{}, 0, P.L);
  }
  if (d.last != nullptr && ch_ok) {
    float* ls = d.last + b * d.ls_sb + (int64_t)c0 * d.ls_sd + 4 * m * d.ls_sn;
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        ls[c * d.ls_sd + (2 * j) * d.ls_sn] = h[c][j].x;
        ls[c * d.ls_sd + (2 * j + 1) * d.ls_sn] = h[c][j].y;
      }
  }
}

// ---- producer warp (1 per group): raw tile -> fp32 operand rows ----------------------------------------------------------
template <typename T, bool SOFTPLUS>
__device__ __forceinline__ void producer_role(const FwdParams& P, GroupSmem<T>& S, const int lane, const int gi) {
  constexpr bool PRECISE = sizeof(T) == 4;
  const Group grp = group_of(P, gi);
  const FwdDir& d = P.dir[grp.dir];
  const bool rev = d.reverse != 0;
  const int cp = lane & (kNP - 1), k0 = lane >> 4;      // unit e = (step k0 + 2e, channel pair cp), e = 0..7
  const int cu = grp.c_base + 2 * cp;
  float2 bias = make_float2(0.f, 0.f);
  if (d.bias && cu < P.dim) bias = make_float2(__ldg(d.bias + cu), __ldg(d.bias + cu + 1));
  const int bc_row = lane >> 2, bc_col = (lane & 3) * 8;   // B|C widening: rows bc_row and bc_row + 8, eight columns

  int it = 0, stage = 0;
  uint32_t stage_par = 0;
  const int nrange = P.ndir == 2 ? 2 : 1;
#pragma unroll 1
  for (int range = 0; range < nrange; ++range) {
    int s_begin, s_end, mode;
    range_of(P, d, range, &s_begin, &s_end, &mode);
    const int nst = s_end - s_begin;
    const int ntile = nst > 0 ? cm_ceil_div(nst, kT) : 0;
#pragma unroll 1
    for (int t = 0; t < ntile; ++t, ++it) {
      const int slot = it & 1;
      const int nvalid = nst - t * kT;
      mbar_wait_long(&S.raw_full[stage], stage_par);
      if (it >= 2) mbar_wait_long(&S.in_empty[slot], ((it >> 1) & 1) ^ 1);
      const Raw<T>& R = S.raw[stage];
      Ops& O = S.ops[slot];
#pragma unroll 4
      for (int e = 0; e < kT / 2; ++e) {
        const int k = k0 + 2 * e;
        const int r = rev ? kT - 1 - k : k;
        const float2 u2 = SmemPair<T>::ld(&R.u[r][2 * cp]);
        float2 dt = fadd2(SmemPair<T>::ld(&R.dl[r][2 * cp]), bias);
        if (SOFTPLUS) { dt.x = softplus_fwd<PRECISE>(dt.x); dt.y = softplus_fwd<PRECISE>(dt.y); }
        if (k >= nvalid) dt = make_float2(0.f, 0.f);   // a = 1, input 0: a missing step leaves the state unchanged
        const float2 du = fmul2(dt, u2);
        O.dd[k][cp] = make_float4(dt.x, dt.y, du.x, du.y);
      }
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int k = bc_row + 8 * e;
        const int r = rev ? kT - 1 - k : k;
        float4 v0, v1;
        SmemOct<T>::ld(&R.bc[r][bc_col], &v0, &v1);
        *reinterpret_cast<float4*>(&O.bc[k][bc_col]) = v0;
        *reinterpret_cast<float4*>(&O.bc[k][bc_col + 4]) = v1;
      }
      warp_arrive(&S.in_full[slot], lane);
      if (lane == 0) tma::mbar_arrive(&S.raw_empty[stage]);   // (after the __syncwarp of warp_arrive: every lane has read the stage)
      if (++stage == kRS) { stage = 0; stage_par ^= 1; }
    }
  }
}

// ---- TMA warp (1 per group): one lane streams the raw tiles ---------------------------------------------------------------
template <typename T>
__device__ __forceinline__ void tma_role(const FwdParams& P, GroupSmem<T>& S, const int lane, const int gi) {
  constexpr int ES = (int)sizeof(T);
  constexpr uint32_t kTile = (uint32_t)(kT * kCH * ES);
  const Group grp = group_of(P, gi);
  const FwdDir& d = P.dir[grp.dir];
  const bool rev = d.reverse != 0;
  const int b = blockIdx.y;
  const bool has_z = P.has_z != 0;
  int it = 0, stage = 0;
  uint32_t stage_par = 0;
  const int nrange = P.ndir == 2 ? 2 : 1;
#pragma unroll 1
  for (int range = 0; range < nrange; ++range) {
    int s_begin, s_end, mode;
    range_of(P, d, range, &s_begin, &s_end, &mode);
    const int nst = s_end - s_begin;
    const int ntile = nst > 0 ? cm_ceil_div(nst, kT) : 0;
    if (range == 1) {
      // the partner direction's stash of this half must be complete (the recurrence warps of both directions arrive)
      asm volatile("bar.sync 1, 192;" ::: "memory");
      tma::fence_proxy_async_all();
    }
    const bool need_z = has_z && mode != FM_STASH;
    const bool need_st = mode == FM_COMBINE;
#pragma unroll 1
    for (int t = 0; t < ntile; ++t, ++it) {
      if (lane == 0) {
        if (it >= kRS) mbar_wait_long(&S.raw_empty[stage], stage_par ^ 1);   // tile it - kRS has been consumed
        const int sb0 = s_begin + t * kT;
        const int t0 = rev ? (P.L - sb0 - kT) : sb0;
        Raw<T>& R = S.raw[stage];
        uint64_t* bar = &S.raw_full[stage];
        tma::fence_proxy_async_smem();
        tma::mbar_expect_tx(bar, kTile * (3u + (need_z ? 1u : 0u) + (need_st ? 1u : 0u)));
        tma::load_3d(&R.u[0][0], &d.m_u, bar, grp.c_base, t0, b);
        tma::load_3d(&R.dl[0][0], &d.m_dl, bar, grp.c_base, t0, b);
        tma::load_3d(&R.bc[0][0], &d.m_bc, bar, 0, t0, b);
        if (need_z) tma::load_3d(&R.z[0][0], &P.m_z, bar, grp.c_base, t0, b);
        if (need_st) tma::load_3d(&R.st[0][0], &P.m_st, bar, grp.c_base, t0, b);
      }
      if (++stage == kRS) { stage = 0; stage_par ^= 1; }
    }
  }
}

template <typename T, bool SOFTPLUS>
__global__ void __launch_bounds__(256, CM_FWDWG_MINB) scan_fwd_wg_kernel(const __grid_constant__ FwdParams P) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  GroupSmem<T>* S = reinterpret_cast<GroupSmem<T>*>(smem_raw);
  const int tid = threadIdx.x;
  if (tid < 2) {
    GroupSmem<T>& G = S[tid];
    for (int i = 0; i < kRS; ++i) {
      tma::mbar_init(&G.raw_full[i], 1);
      tma::mbar_init(&G.raw_empty[i], 3);   // 2 recurrence warps + the producer warp
    }
    for (int i = 0; i < 2; ++i) {
      tma::mbar_init(&G.in_full[i], 1);
      tma::mbar_init(&G.in_empty[i], 2);
    }
    tma::fence_barrier_init();
  }
  __syncthreads();
  // thread layout: [recurrence group 0 (2 warps) | recurrence group 1 | producer 0 | producer 1 | TMA 0 | TMA 1]
  if (tid < 128) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(CM_FWDWG_RREG));
    const int gi = tid >> 6;
    scan_role<T>(P, S[gi], tid & 63, gi);
  } else {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(CM_FWDWG_IREG));
    const int w = (tid - 128) >> 5, lane = tid & 31;
    if (w < 2) producer_role<T, SOFTPLUS>(P, S[w], lane, w);
    else tma_role<T>(P, S[w - 2], lane, w - 2);
  }
}

// ---- host --------------------------------------------------------------------------------------------------------------
static bool step_stride32(int64_t sl_elems, int es, bool reverse, int32_t* out) {
  const int64_t v = (reverse ? -sl_elems : sl_elems) * es;
  if (v > INT32_MAX / 2 || v < INT32_MIN / 2) return false;
  *out = (int32_t)v;
  return true;
}

template <typename T>
static bool build_params(const cm_scan_fwd_args& a, FwdParams* P) {
  constexpr int ES = (int)sizeof(T);
  if (a.dstate != 16 || a.dim % kCH != 0 || a.seqlen < 1) return false;
  auto pair_ok = [&](const cm_tensor3& t) {
    return t.ptr != nullptr && t.sd == 1 && (reinterpret_cast<uintptr_t>(t.ptr) % (2 * ES)) == 0 && t.sb % 2 == 0 && t.sl % 2 == 0;
  };
  if (!pair_ok(a.out)) return false;
  if (a.out_pre.ptr != nullptr && !pair_ok(a.out_pre)) return false;
  P->L = a.seqlen; P->dim = a.dim; P->ndir = a.ndir;
  P->has_z = a.z.ptr != nullptr; P->has_pre = a.out_pre.ptr != nullptr;
  P->scale = a.out_scale; P->pad0 = P->pad1 = 0;
  const int64_t L = a.seqlen, Bt = a.batch, D = a.dim;
  if (a.z.ptr != nullptr && (a.z.sd != 1 || !tma::make_map_blc(&P->m_z, a.z.ptr, ES, D, L, Bt, a.z.sl, a.z.sb, kCH, kT))) return false;
  if (a.ndir == 2 && !tma::make_map_blc(&P->m_st, a.out.ptr, ES, D, L, Bt, a.out.sl, a.out.sb, kCH, kT)) return false;
  for (int r = 0; r < a.ndir; ++r) {
    const cm_scan_dir& s = a.dir[r];
    FwdDir& d = P->dir[r];
    if (s.bc_const) return false;
    if (s.u.sd != 1 || s.delta.sd != 1 || s.Bm.sd != 1 || s.Cm.sd != 1) return false;
    // one 32-element B|C row per step
    if (static_cast<const char*>(s.Cm.ptr) != static_cast<const char*>(s.Bm.ptr) + 16 * ES || s.Cm.sl != s.Bm.sl || s.Cm.sb != s.Bm.sb)
      return false;
    if (s.ckpt != nullptr && ((reinterpret_cast<uintptr_t>(s.ckpt) & 15) != 0 || (s.ckpt_sb % 4) != 0 || (s.ckpt_sd % 4) != 0))
      return false;
    if (!tma::make_map_blc(&d.m_u, s.u.ptr, ES, D, L, Bt, s.u.sl, s.u.sb, kCH, kT)) return false;
    if (!tma::make_map_blc(&d.m_dl, s.delta.ptr, ES, D, L, Bt, s.delta.sl, s.delta.sb, kCH, kT)) return false;
    if (!tma::make_map_blc(&d.m_bc, s.Bm.ptr, ES, 32, L, Bt, s.Bm.sl, s.Bm.sb, 32, kT)) return false;
    const bool rev = s.reverse != 0;
    d.reverse = rev;
    const int64_t l0 = rev ? L - 1 : 0;
    d.out = static_cast<char*>(a.out.ptr) + l0 * a.out.sl * ES;
    d.out_sb = a.out.sb * ES;
    if (!step_stride32(a.out.sl, ES, rev, &d.out_ss)) return false;
    d.pre = nullptr; d.pre_sb = 0; d.pre_ss = 0;
    if (a.out_pre.ptr != nullptr) {
      d.pre = static_cast<char*>(a.out_pre.ptr) + l0 * a.out_pre.sl * ES;
      d.pre_sb = a.out_pre.sb * ES;
      if (!step_stride32(a.out_pre.sl, ES, rev, &d.pre_ss)) return false;
    }
    d.s1 = cm_first_range(a.seqlen, a.ndir, s.reverse);
    d.A = s.A; d.A_sd = s.A_sd; d.A_sn = s.A_sn;
    d.Dskip = s.Dskip; d.bias = s.delta_bias;
    d.ckpt = s.ckpt; d.ckpt_sb = s.ckpt_sb; d.ckpt_sd = s.ckpt_sd;
    d.last = s.last_state; d.ls_sb = s.ls_sb; d.ls_sd = s.ls_sd; d.ls_sn = s.ls_sn;
  }
  return true;
}

template <typename T>
static int try_t(const cm_scan_fwd_args& a, cudaStream_t st, int* rc) {
  FwdParams P;
  if (!build_params<T>(a, &P)) return 0;
  const size_t smem = 2 * sizeof(GroupSmem<T>);
  const bool sp = (a.flags & CM_FLAG_DELTA_SOFTPLUS) != 0;
  const int blocks_c = a.dim / kCH;
  const unsigned gx = (unsigned)(a.ndir == 2 ? blocks_c : (blocks_c + 1) / 2);
  auto launch = [&](auto kern) -> int {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);   // per device
    if (e != cudaSuccess) return (int)e;
    kern<<<dim3(gx, a.batch, 1), 256, smem, st>>>(P);
    e = cudaGetLastError();
    return e == cudaSuccess ? 0 : (int)e;
  };
  *rc = sp ? launch(scan_fwd_wg_kernel<T, true>) : launch(scan_fwd_wg_kernel<T, false>);
  return 1;
}

}  // namespace wgf

// returns 1 if launched (result in *rc), 0 if the warpgroup path does not apply
int scan_fwd_try_warpgroup(const cm_scan_fwd_args& a, cudaStream_t st, int* rc) {
  if (getenv("CM_SCAN_NO_WG") != nullptr) return 0;
  switch (a.dtype) {
    case CM_F32: return wgf::try_t<float>(a, st, rc);
    case CM_BF16: return wgf::try_t<__nv_bfloat16>(a, st, rc);
    default: return wgf::try_t<__half>(a, st, rc);
  }
}

}  // namespace cm
