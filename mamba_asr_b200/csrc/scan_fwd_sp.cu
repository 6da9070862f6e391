// Selective-scan forward, state-parallel channel-last kernel for sm_100a ("sp" path).
//
// Mathematics, bidirectional stash/combine protocol and checkpoint contract are those of scan_fwd.cu (see its header;
// reference semantics: modules/mamba/selective_scan_interface.py:106-157 and modules/mamba/bimamba.py:223-253).
// What differs is the mapping of work to lanes.  scan_fwd.cu / scan_fwd_cl.cu give one lane a whole channel (16 states
// in registers); that is op-optimal but leaves batch*dim*ndir/32 warps - 576 warps for 32 x 288 channels, fewer than
// the 592 SM sub-partitions of a B200 - and every step pays the per-channel scalar work (softplus, gate, conversions,
// predicates) at 1/32 lane efficiency in front of a latency-exposed recurrence.
//
// Here a LANE OWNS ONE STATE n of CPL adjacent channels: a warp = 2 half-warps x 16 states = 2*CPL channels, so there
// are 16/CPL times more warps, and the time loop of a lane is the bare recurrence
//     LDS (dt,du)  LDS (B_n,C_n)  FMUL2  MUFU.EX2 x CPL  FMUL2  FFMA2  FMUL2  STS
// - about 9 issue slots per 16 MUFU cycles for CPL = 2: the XU pipe (16 ex2/clk/SM) is the binding unit, as it must be
// for this op on B200.  Everything that is per (step, channel) rather than per (step, channel,
// state) is done by the same threads in vectorised phases around the recurrence, one TILE of kT steps at a time:
//   pre-phase : each thread converts ONE (step, channel-pair): raw loads (issued a whole tile ahead, straight into
//               registers), softplus, dt*u  -> fp32 (dt, du) rows in shared memory; B/C rows -> fp32 (B_n, C_n) pairs
//   recurrence: writes the products C_n*h_n to a padded shared tile
//   epilogue  : the thread that converted a (step, channel-pair) sums its 16 products (LDS.128, conflict-free), adds
//               D*u, stashes / combines with the partner direction / gates, and stores 2 channels with one access.
// The two time directions of a channel block share a CTA (one 128-thread group each, private named barriers) and meet
// in the middle exactly like scan_fwd.cu: range 0 stashes pre-gate sums in `out`, one __syncthreads(), range 1 reads the
// partner's stash back (L2), adds, gates once, writes the final value and `out_pre`.
//
// Requirements (else the launcher falls through to the other kernels): unit channel stride, dstate == 16, variable
// B/C with unit state stride, dim a multiple of the CTA's channel count, pair/quad-aligned rows.
#include <cstdlib>
#include <type_traits>

#include "common.cuh"

namespace cm {
namespace sp {

constexpr int kT = 16;          // steps per tile
constexpr int kNW = 4;          // warps per direction group
constexpr int kGT = kNW * 32;   // threads per direction group

template <int CPL>
struct Cfg {
  static constexpr int CH = kNW * 2 * CPL;        // channels per CTA
  static constexpr int NG = CH / CPL;             // channel groups per CTA (= lanes' channel blocks) = 2*kNW
  static constexpr int PROW = 16 * CPL + 4;       // padded row of the product tile (floats)
  static constexpr int NPAIR = CH / 2;            // channel pairs per step
  static constexpr int UPT = kT * NPAIR / kGT;    // (step, pair) units per thread
  static constexpr int KSTRIDE = kGT / NPAIR;     // step distance between a thread's units
  static_assert(kT * NPAIR % kGT == 0 && UPT >= 1, "tile must divide over the group");
};

template <int CPL>
struct DirSmem {
  using C = Cfg<CPL>;
  float dd[2][kT][C::NG][2 * CPL];   // per channel group: dt[CPL] then du[CPL]
  float2 bc[2][kT][16];              // (B_n, C_n)
  float p[kT][C::NG][C::PROW];       // products C_n*h_n: element [n*CPL + j]
};

// ---- paired element I/O ------------------------------------------------------------------------------------------
template <typename T> struct Pair;
template <> struct Pair<float> {
  using Raw = float2;
  static __device__ __forceinline__ Raw ld_nc(const void* p) {
    Raw r; asm volatile("ld.global.nc.v2.f32 {%0,%1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p)); return r;
  }
  static __device__ __forceinline__ Raw ld_cg(const void* p) {
    Raw r; asm volatile("ld.global.cg.v2.f32 {%0,%1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p)); return r;
  }
  static __device__ __forceinline__ Raw zero() { return make_float2(0.f, 0.f); }
  static __device__ __forceinline__ float2 cvt(Raw r) { return r; }
  static __device__ __forceinline__ void st(void* p, float2 v) { *reinterpret_cast<float2*>(p) = v; }
};
template <> struct Pair<__nv_bfloat16> {
  using Raw = uint32_t;
  static __device__ __forceinline__ Raw ld_nc(const void* p) {
    Raw r; asm volatile("ld.global.nc.b32 %0, [%1];" : "=r"(r) : "l"(p)); return r;
  }
  static __device__ __forceinline__ Raw ld_cg(const void* p) {
    Raw r; asm volatile("ld.global.cg.b32 %0, [%1];" : "=r"(r) : "l"(p)); return r;
  }
  static __device__ __forceinline__ Raw zero() { return 0u; }
  static __device__ __forceinline__ float2 cvt(Raw r) {
    return make_float2(__uint_as_float(r << 16), __uint_as_float(r & 0xffff0000u));
  }
  static __device__ __forceinline__ void st(void* p, float2 v) {
    *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(v.x, v.y);
  }
};
template <> struct Pair<__half> {
  using Raw = uint32_t;
  static __device__ __forceinline__ Raw ld_nc(const void* p) {
    Raw r; asm volatile("ld.global.nc.b32 %0, [%1];" : "=r"(r) : "l"(p)); return r;
  }
  static __device__ __forceinline__ Raw ld_cg(const void* p) {
    Raw r; asm volatile("ld.global.cg.b32 %0, [%1];" : "=r"(r) : "l"(p)); return r;
  }
  static __device__ __forceinline__ Raw zero() { return 0u; }
  static __device__ __forceinline__ float2 cvt(Raw r) {
    return __half22float2(*reinterpret_cast<const __half2*>(&r));
  }
  static __device__ __forceinline__ void st(void* p, float2 v) {
    *reinterpret_cast<__half2*>(p) = __floats2half2_rn(v.x, v.y);
  }
};

// four consecutive elements (one quarter of a B or C row)
template <typename T> struct Quad;
template <> struct Quad<float> {
  using Raw = float4;
  static __device__ __forceinline__ Raw ld_nc(const void* p) {
    Raw r;
    asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
  }
  static __device__ __forceinline__ Raw zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
  static __device__ __forceinline__ void cvt(Raw r, float* o) { o[0] = r.x; o[1] = r.y; o[2] = r.z; o[3] = r.w; }
};
template <> struct Quad<__nv_bfloat16> {
  using Raw = uint2;
  static __device__ __forceinline__ Raw ld_nc(const void* p) {
    Raw r;
    asm volatile("ld.global.nc.v2.b32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
  }
  static __device__ __forceinline__ Raw zero() { return make_uint2(0u, 0u); }
  static __device__ __forceinline__ void cvt(Raw r, float* o) {
    o[0] = __uint_as_float(r.x << 16); o[1] = __uint_as_float(r.x & 0xffff0000u);
    o[2] = __uint_as_float(r.y << 16); o[3] = __uint_as_float(r.y & 0xffff0000u);
  }
};
template <> struct Quad<__half> {
  using Raw = uint2;
  static __device__ __forceinline__ Raw ld_nc(const void* p) {
    Raw r;
    asm volatile("ld.global.nc.v2.b32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
  }
  static __device__ __forceinline__ Raw zero() { return make_uint2(0u, 0u); }
  static __device__ __forceinline__ void cvt(Raw r, float* o) {
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&r.x));
    const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&r.y));
    o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y;
  }
};

// private barrier of a direction group (immediate ids: a register id would make ptxas reserve all 16 barriers)
__device__ __forceinline__ void group_bar(int grp) {
  if (grp == 0) asm volatile("bar.sync 1, %0;" ::"n"(kGT) : "memory");
  else asm volatile("bar.sync 2, %0;" ::"n"(kGT) : "memory");
}

enum { FM_UNI = 0, FM_STASH = 1, FM_COMBINE = 2 };

#ifndef CM_FWDSP_MINB
#define CM_FWDSP_MINB 3
#endif
constexpr int kSub = 4;   // steps per software-pipelined sub-block of the recurrence

// One direction group (128 threads) of a CTA.  DIR is a template parameter so that every p.dir[DIR] field is a
// constant-bank operand (no indexed LDC, no registers spent on strides).
template <typename T, int CPL, int NDIR, int DIR>
__device__ __forceinline__ void run_dir(const cm_scan_fwd_args& p, DirSmem<CPL>& S, const int gt) {
  using C = Cfg<CPL>;
  using P2 = Pair<T>;
  using Q4 = Quad<T>;
  constexpr int ES = (int)sizeof(T);
  constexpr bool PRECISE = sizeof(T) == 4;
  const cm_scan_dir& dp = p.dir[DIR];
  const int warp = gt >> 5, lane = gt & 31;
  const int b = blockIdx.y;
  const int c_base = blockIdx.x * C::CH;
  const int L = p.seqlen;
  const bool rev = dp.reverse != 0;
  const bool softplus = (p.flags & CM_FLAG_DELTA_SOFTPLUS) != 0;
  const bool has_z = p.z.ptr != nullptr;

  // ---- recurrence identity: state n of channel group g
  const int hw = lane >> 4, n = lane & 15;
  const int g = warp * 2 + hw;
  const int cs = c_base + g * CPL;
  float kA[CPL], h[CPL];
#pragma unroll
  for (int j = 0; j < CPL; ++j) {
    kA[j] = __ldg(dp.A + (int64_t)(cs + j) * dp.A_sd + n * dp.A_sn) * kLog2e;
    h[j] = 0.f;
  }
  float* ckp = dp.ckpt ? dp.ckpt + b * dp.ckpt_sb + (int64_t)cs * dp.ckpt_sd + n : nullptr;

  // ---- unit identity (pre-phase and epilogue): channel pair cp at steps k0 + i*KSTRIDE of every tile
  const int cp = gt % C::NPAIR, k0 = gt / C::NPAIR;
  const int cu = c_base + 2 * cp;
  float bias[2], Dsk[2];
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    bias[j] = dp.delta_bias ? __ldg(dp.delta_bias + cu + j) : 0.f;
    Dsk[j] = dp.Dskip ? __ldg(dp.Dskip + cu + j) : 0.f;
  }
  const int64_t sgn = rev ? -1 : 1, l0 = rev ? (L - 1) : 0;
  // byte pointer of this thread's first unit at processed step s
  auto at = [&](const cm_tensor3& t, int s) {
    return static_cast<char*>(t.ptr) + (b * t.sb + (int64_t)cu * t.sd + (l0 + sgn * s) * t.sl) * ES;
  };
  // B/C quarter rows: thread (row kq, part) of the tile
  const int kq = gt >> 3, part = gt & 7, q4 = (part & 3) * 4;
  const cm_tensor3& bct = (part < 4) ? dp.Bm : dp.Cm;

  typename P2::Raw ru[C::UPT], rdl[C::UPT];
  typename Q4::Raw rbc;
  float2 uu[C::UPT];

  const int s1 = cm_first_range(L, NDIR, dp.reverse);
#pragma unroll 1
  for (int range = 0; range < NDIR; ++range) {
    if (NDIR == 2 && range == 1) __syncthreads();   // partner's stash of the other half is complete
    const int mode = (NDIR == 1) ? FM_UNI : (range == 0 ? FM_STASH : FM_COMBINE);
    const int s_begin = range == 0 ? 0 : s1, s_end = range == 0 ? s1 : L;
    int jck = range == 0 ? 0 : cm_ceil_div(s1, CM_SCAN_CKPT_STEPS);   // next checkpoint slot
    const int nst = s_end - s_begin;
    if (nst <= 0) continue;
    const int ntile = cm_ceil_div(nst, kT);
    const bool need_z = has_z && mode != FM_STASH;
    const bool need_st = mode == FM_COMBINE;

    // advancing pointers: *_n = the tile being prefetched, *_c = the tile being finished
    const char* pu_n = at(dp.u, s_begin + k0);
    const char* pdl_n = at(dp.delta, s_begin + k0);
    const char* pbc_n = static_cast<const char*>(bct.ptr) + (b * bct.sb + (l0 + sgn * (s_begin + kq)) * bct.sl + q4) * ES;
    const char* pz_c = has_z ? at(p.z, s_begin + k0) : nullptr;
    char* po_c = at(p.out, s_begin + k0);
    char* ppre_c = p.out_pre.ptr ? at(p.out_pre, s_begin + k0) : nullptr;
    int rem_n = nst - k0;       // > i*KSTRIDE  <=> unit i of the prefetched tile exists
    int remq_n = nst - kq;      // > 0 <=> B/C row of the prefetched tile exists

    auto load_raw = [&]() {
#pragma unroll
      for (int i = 0; i < C::UPT; ++i) {
        if (rem_n > i * C::KSTRIDE) {
          ru[i] = P2::ld_nc(pu_n + (int64_t)i * C::KSTRIDE * sgn * dp.u.sl * ES);
          rdl[i] = P2::ld_nc(pdl_n + (int64_t)i * C::KSTRIDE * sgn * dp.delta.sl * ES);
        } else {
          ru[i] = P2::zero();
          rdl[i] = P2::zero();
        }
      }
      rbc = (remq_n > 0) ? Q4::ld_nc(pbc_n) : Q4::zero();
    };
    auto advance_n = [&]() {
      pu_n += (int64_t)kT * sgn * dp.u.sl * ES;
      pdl_n += (int64_t)kT * sgn * dp.delta.sl * ES;
      pbc_n += (int64_t)kT * sgn * bct.sl * ES;
      rem_n -= kT;
      remq_n -= kT;
    };
    // converts the raw registers (loaded with the `rem` in force at load time, passed here as remv)
    auto convert_store = [&](int buf, int remv) {
#pragma unroll
      for (int i = 0; i < C::UPT; ++i) {
        const int k = k0 + i * C::KSTRIDE;
        const bool valid = remv > i * C::KSTRIDE;
        const float2 u2 = P2::cvt(ru[i]);
        const float2 d2 = P2::cvt(rdl[i]);
        float dt0 = d2.x + bias[0], dt1 = d2.y + bias[1];
        if (softplus) { dt0 = softplus_fwd<PRECISE>(dt0); dt1 = softplus_fwd<PRECISE>(dt1); }
        if (!valid) { dt0 = 0.f; dt1 = 0.f; }   // a = 1, input 0: the state passes through a missing step unchanged
        uu[i] = u2;
        float* row = &S.dd[buf][k][(2 * cp) / CPL][0];
        if constexpr (CPL == 2) {
          *reinterpret_cast<float4*>(row) = make_float4(dt0, dt1, dt0 * u2.x, dt1 * u2.y);
        } else {
          const int j = (2 * cp) % CPL;
          *reinterpret_cast<float2*>(row + j) = make_float2(dt0, dt1);
          *reinterpret_cast<float2*>(row + CPL + j) = make_float2(dt0 * u2.x, dt1 * u2.y);
        }
      }
      float v[4];
      Q4::cvt(rbc, v);
      float* dst = reinterpret_cast<float*>(&S.bc[buf][kq][q4]) + (part >> 2);
#pragma unroll
      for (int i = 0; i < 4; ++i) dst[2 * i] = v[i];
    };

    load_raw();
    convert_store(0, rem_n);
    advance_n();
    group_bar(DIR);

    int rem_c = nst - k0;       // validity of the tile being finished
#pragma unroll 1
    for (int t = 0; t < ntile; ++t) {
      const int buf = t & 1;
      load_raw();               // tile t+1 (all-zero beyond the range)
      // epilogue operands of this tile, in flight during the recurrence
      typename P2::Raw rz[C::UPT], rst[C::UPT];
#pragma unroll
      for (int i = 0; i < C::UPT; ++i) {
        const bool v = rem_c > i * C::KSTRIDE;
        rz[i] = (need_z && v) ? P2::ld_nc(pz_c + (int64_t)i * C::KSTRIDE * sgn * p.z.sl * ES) : P2::zero();
        rst[i] = (need_st && v) ? P2::ld_cg(po_c + (int64_t)i * C::KSTRIDE * sgn * p.out.sl * ES) : P2::zero();
      }

      // ---- recurrence over the tile: sub-blocks of kSub steps, operands of the next sub-block loaded while the
      // dependent chain of the current one runs.  Missing steps of a last partial tile carry dt = 0 (identity).
      const int nvalid = nst - t * kT;    // steps of this tile inside the range (may exceed kT)
      {
        const float* ddb = &S.dd[buf][0][g][0];
        const float2* bcb = &S.bc[buf][0][n];
        float* pb = &S.p[0][g][n * CPL];
        constexpr int DDS = C::NG * 2 * CPL;      // floats per step in dd
        constexpr int PS = C::NG * C::PROW;       // floats per step in p
        float ddr[kSub][2 * CPL];
        float2 bcr[kSub];
        auto lds_sub = [&](int sb) {
#pragma unroll
          for (int i = 0; i < kSub; ++i) {
            const float* src = ddb + (sb * kSub + i) * DDS;
            if constexpr (CPL == 2) {
              const float4 v = *reinterpret_cast<const float4*>(src);
              ddr[i][0] = v.x; ddr[i][1] = v.y; ddr[i][2] = v.z; ddr[i][3] = v.w;
            } else {
              const float4 v = *reinterpret_cast<const float4*>(src);
              const float4 w = *reinterpret_cast<const float4*>(src + 4);
              ddr[i][0] = v.x; ddr[i][1] = v.y; ddr[i][2] = v.z; ddr[i][3] = v.w;
              ddr[i][4] = w.x; ddr[i][5] = w.y; ddr[i][6] = w.z; ddr[i][7] = w.w;
            }
            bcr[i] = bcb[(sb * kSub + i) * 16];
          }
        };
        lds_sub(0);
#pragma unroll
        for (int sb = 0; sb < kT / kSub; ++sb) {
          if (((sb * kSub) % CM_SCAN_CKPT_STEPS) == 0) {
            if (ckp != nullptr && sb * kSub < nvalid) {
              float* dst = ckp + (int64_t)jck * 16;
#pragma unroll
              for (int j = 0; j < CPL; ++j) dst[j * dp.ckpt_sd] = h[j];
            }
            ++jck;
          }
          // exponent arguments, decays and input products of the whole sub-block
          float a[kSub][CPL], ub[kSub][CPL], cc[kSub];
#pragma unroll
          for (int i = 0; i < kSub; ++i) {
#pragma unroll
            for (int j = 0; j < CPL; j += 2) {
              const float2 x = fmul2(make_float2(ddr[i][j], ddr[i][j + 1]), make_float2(kA[j], kA[j + 1]));
              a[i][j] = x.x; a[i][j + 1] = x.y;
            }
          }
#pragma unroll
          for (int i = 0; i < kSub; ++i) {
#pragma unroll
            for (int j = 0; j < CPL; ++j) a[i][j] = ex2(a[i][j]);
          }
#pragma unroll
          for (int i = 0; i < kSub; ++i) {
#pragma unroll
            for (int j = 0; j < CPL; j += 2) {
              const float2 w = fmul2(make_float2(ddr[i][CPL + j], ddr[i][CPL + j + 1]), make_float2(bcr[i].x, bcr[i].x));
              ub[i][j] = w.x; ub[i][j + 1] = w.y;
            }
            cc[i] = bcr[i].y;
          }
          if (sb + 1 < kT / kSub) lds_sub(sb + 1);
          // the dependent chain
#pragma unroll
          for (int i = 0; i < kSub; ++i) {
            float pv[CPL];
#pragma unroll
            for (int j = 0; j < CPL; j += 2) {
              const float2 hn = ffma2(make_float2(a[i][j], a[i][j + 1]), make_float2(h[j], h[j + 1]),
                                      make_float2(ub[i][j], ub[i][j + 1]));
              h[j] = hn.x; h[j + 1] = hn.y;
              const float2 pp = fmul2(make_float2(cc[i], cc[i]), hn);
              pv[j] = pp.x; pv[j + 1] = pp.y;
            }
            float* pr = pb + (sb * kSub + i) * PS;
            if constexpr (CPL == 2) *reinterpret_cast<float2*>(pr) = make_float2(pv[0], pv[1]);
            else *reinterpret_cast<float4*>(pr) = make_float4(pv[0], pv[1], pv[2], pv[3]);
          }
        }
        // slots are counted per EXISTING checkpoint: undo the count of a slot whose step lies beyond the range
        if (nvalid <= CM_SCAN_CKPT_STEPS) --jck;
      }
      group_bar(DIR);

      // ---- epilogue: one (step, channel pair) per unit
#pragma unroll
      for (int i = 0; i < C::UPT; ++i) {
        const int k = k0 + i * C::KSTRIDE;
        float2 y;
        {
          // row of channel group (2cp)/CPL; this pair's products sit at [n*CPL + jj], jj = (2cp)%CPL + {0,1}
          const float* row = &S.p[k][(2 * cp) / CPL][(2 * cp) % CPL];
          float2 acc0 = make_float2(0.f, 0.f), acc1 = make_float2(0.f, 0.f);
          if constexpr (CPL == 2) {
            const float4* r = reinterpret_cast<const float4*>(row);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float4 v = r[q];
              acc0 = fadd2(acc0, make_float2(v.x, v.y));
              acc1 = fadd2(acc1, make_float2(v.z, v.w));
            }
          } else {
#pragma unroll
            for (int nn = 0; nn < 16; nn += 2) {
              acc0 = fadd2(acc0, *reinterpret_cast<const float2*>(row + nn * CPL));
              acc1 = fadd2(acc1, *reinterpret_cast<const float2*>(row + (nn + 1) * CPL));
            }
          }
          y = fadd2(acc0, acc1);
        }
        y.x = fmaf(Dsk[0], uu[i].x, y.x);
        y.y = fmaf(Dsk[1], uu[i].y, y.y);
        if (rem_c > i * C::KSTRIDE) {
          char* po = po_c + (int64_t)i * C::KSTRIDE * sgn * p.out.sl * ES;
          if (mode == FM_STASH) {
            P2::st(po, y);
          } else {
            const float2 stv = P2::cvt(rst[i]);
            const float2 tot = make_float2(y.x + stv.x, y.y + stv.y);
            float2 val = make_float2(tot.x * p.out_scale, tot.y * p.out_scale);
            if (need_z) {
              const float2 zz = P2::cvt(rz[i]);
              val.x *= zz.x * sigmoid_sel<PRECISE>(zz.x);
              val.y *= zz.y * sigmoid_sel<PRECISE>(zz.y);
            }
            if (ppre_c != nullptr) P2::st(ppre_c + (int64_t)i * C::KSTRIDE * sgn * p.out_pre.sl * ES, tot);
            P2::st(po, val);
          }
        }
      }
      // advance the finished-tile pointers
      if (has_z) pz_c += (int64_t)kT * sgn * p.z.sl * ES;
      po_c += (int64_t)kT * sgn * p.out.sl * ES;
      if (ppre_c != nullptr) ppre_c += (int64_t)kT * sgn * p.out_pre.sl * ES;
      rem_c -= kT;
      convert_store(buf ^ 1, rem_n);   // tile t+1 (zeros past the end: harmless, never consumed)
      advance_n();
      group_bar(DIR);
    }
  }

  if (dp.last_state != nullptr) {
    float* ls = dp.last_state + b * dp.ls_sb + (int64_t)cs * dp.ls_sd + n * dp.ls_sn;
#pragma unroll
    for (int j = 0; j < CPL; ++j) ls[j * dp.ls_sd] = h[j];
  }
}

template <typename T, int CPL, int NDIR>
__global__ void __launch_bounds__(NDIR* kGT, CM_FWDSP_MINB) scan_fwd_sp_kernel(const __grid_constant__ cm_scan_fwd_args p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  DirSmem<CPL>* S = reinterpret_cast<DirSmem<CPL>*>(smem_raw);
  const int tid = threadIdx.x;
  if (NDIR == 1 || tid < kGT) run_dir<T, CPL, NDIR, 0>(p, S[0], tid);
  else run_dir<T, CPL, NDIR, 1>(p, S[1], tid - kGT);
}

template <typename T>
static bool t_ok(const cm_tensor3& t, int64_t quantum) {
  const int64_t es = sizeof(T);
  return t.ptr != nullptr && t.sd == 1 && (reinterpret_cast<uintptr_t>(t.ptr) % (quantum * es)) == 0 &&
         t.sb % quantum == 0 && t.sl % quantum == 0;
}

template <typename T>
static bool fwd_sp_ok(const cm_scan_fwd_args& a, int cpl) {
  const int ch = kNW * 2 * cpl;
  if (a.dstate != 16 || a.dim % ch != 0) return false;
  if (!t_ok<T>(a.out, 2)) return false;
  if (a.z.ptr != nullptr && !t_ok<T>(a.z, 2)) return false;
  if (a.out_pre.ptr != nullptr && !t_ok<T>(a.out_pre, 2)) return false;
  for (int r = 0; r < a.ndir; ++r) {
    const cm_scan_dir& d = a.dir[r];
    if (d.bc_const) return false;
    if (!t_ok<T>(d.u, 2) || !t_ok<T>(d.delta, 2) || !t_ok<T>(d.Bm, 4) || !t_ok<T>(d.Cm, 4)) return false;
  }
  return true;
}

template <typename T, int CPL, int NDIR>
static int launch_one(const cm_scan_fwd_args& a, cudaStream_t st) {
  const size_t smem = sizeof(DirSmem<CPL>) * NDIR;
  auto kern = scan_fwd_sp_kernel<T, CPL, NDIR>;
  static bool attr_done = false;   // idempotent attribute; a benign race sets it twice
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    attr_done = true;
  }
  kern<<<dim3(a.dim / Cfg<CPL>::CH, a.batch), NDIR * kGT, smem, st>>>(a);
  CM_LAUNCH_CHECK();
  return 0;
}

template <typename T>
static int launch_t(const cm_scan_fwd_args& a, int cpl, cudaStream_t st) {
  if (a.ndir == 2) {
    if (cpl == 2) return launch_one<T, 2, 2>(a, st);
    return launch_one<T, 4, 2>(a, st);
  }
  if (cpl == 2) return launch_one<T, 2, 1>(a, st);
  return launch_one<T, 4, 1>(a, st);
}

}  // namespace sp

// returns 1 if launched (result in *rc), 0 if the state-parallel path does not apply
int scan_fwd_try_state_parallel(const cm_scan_fwd_args& a, cudaStream_t st, int* rc) {
  int cpl = 2;
  if (const char* e = getenv("CM_SP_CPL")) cpl = atoi(e);   // A/B measurements only
  if (cpl != 2 && cpl != 4) cpl = 2;
  while (cpl > 2 && a.dim % (sp::kNW * 2 * cpl) != 0) cpl >>= 1;
  switch (a.dtype) {
    case CM_F32:
      if (!sp::fwd_sp_ok<float>(a, cpl)) return 0;
      *rc = sp::launch_t<float>(a, cpl, st);
      return 1;
    case CM_BF16:
      if (!sp::fwd_sp_ok<__nv_bfloat16>(a, cpl)) return 0;
      *rc = sp::launch_t<__nv_bfloat16>(a, cpl, st);
      return 1;
    default:
      if (!sp::fwd_sp_ok<__half>(a, cpl)) return 0;
      *rc = sp::launch_t<__half>(a, cpl, st);
      return 1;
  }
}

}  // namespace cm
