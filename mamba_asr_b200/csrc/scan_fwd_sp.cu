// Selective-scan forward, state-parallel channel-last kernel for sm_100a ("sp" path) - the default forward kernel.
//
// Mathematics, bidirectional stash/combine protocol and checkpoint contract are those of scan_fwd.cu (see its header;
// reference semantics: modules/mamba/selective_scan_interface.py:106-157 and modules/mamba/bimamba.py:223-253).
// What differs is the mapping of work to lanes.  scan_fwd.cu / scan_fwd_cl.cu give one lane a whole channel (16 states
// in registers): op-optimal, but only batch*dim*ndir/32 warps exist - 576 for 32 x 288 channels, fewer than the 592 SM
// sub-partitions of a B200 - and the per-channel scalar work (softplus, gate, conversions, predicates) runs in front
// of a latency-exposed recurrence at 1/32 lane efficiency.
//
// Here a LANE OWNS FOUR STATES (4m..4m+3) OF TWO ADJACENT CHANNELS: 4 lanes cover a channel pair, a warp 8 pairs, and
// the kernel is WARP-SPECIALISED.  Per time direction a CTA (32 channels) runs
//   2 recurrence warps: the bare time loop on packed fp32 pairs, per step
//        LDS.128 (dt0,dt1,du0,du1)  2 LDS.128 (B[4], C[4])  4 FMUL2  8 MUFU.EX2  4 FMUL2  4 FFMA2  FMUL2 3 FFMA2  STS.64
//      = 28 issue slots per 64 XU cycles, then a warp-local reduction of the 4 partial sums of each (step, pair);
//   2 IO warps: raw global loads issued a whole tile ahead straight into registers, softplus, dt*u -> fp32
//      (dt0,dt1,du0,du1) rows and fp32 B/C rows in a 2-tile shared ring; epilogue: D*u skip term, stash / combine with
//      the partner direction / gate, 2 channels per store.
// The roles meet only through mbarriers (full/empty pairs on the operand ring and on the y ring), so the recurrence
// warps never see a global-memory latency.  WHY 4 x 2: measured with ncu, the binding unit of the 2-states-x-2-channels
// variant was neither the MUFU pipe (45 %) nor issue (42 %) but the LSU data pipe (l1tex__data_pipe_lsu_wavefronts
// 90 %): every state update needs its operands delivered to registers through the 128 B/clk shared-memory pipe,
// 12/S + 8/C bytes per update for S states x C channels per lane (dt,du,y per channel; B,C per state).  2x2 = 10 B,
// 4x2 = 7 B, which puts the pipe at 18 updates/clk/SM against the MUFU pipe's 16.
// The two time directions of a channel block share a CTA and meet in the middle exactly like scan_fwd.cu: range 0
// stashes pre-gate sums in `out`, one __syncthreads(), range 1 reads the partner's stash back (L2), adds, gates once,
// writes the final value and `out_pre`.
//
// Requirements (else the launcher falls through to the other kernels): unit channel stride, dstate == 16, variable
// B/C with unit state stride, dim a multiple of 32, pair/quad-aligned rows, 32-bit byte strides per step.
#include <climits>
#include <cstdlib>

#include "sp_common.cuh"

namespace cm {
namespace sp {

// kernel-side view of one direction: byte pointers at (batch 0, channel 0, PROCESSED step 0) and signed byte strides
// per processed step, so that every address is one IMAD.WIDE with a constant-bank stride
struct FwdDir {
  const char *u, *dl, *B, *C, *z;
  char *out, *pre;
  int64_t u_sb, dl_sb, b_sb, c_sb, z_sb, out_sb, pre_sb;   // batch strides (bytes)
  int32_t u_ss, dl_ss, b_ss, c_ss, z_ss, out_ss, pre_ss;   // bytes per processed step
  int32_t s1;                                              // length of the first range (cm_first_range)
  int32_t reverse, pad0;
  const float* A;
  int64_t A_sd, A_sn;
  const float *Dskip, *bias;
  float* ckpt;
  int64_t ckpt_sb, ckpt_sd;
  float* last;
  int64_t ls_sb, ls_sd, ls_sn;
};
struct FwdParams {
  int32_t L;
  uint32_t flags;
  float scale;
  int32_t dim;
  // time windows (small-batch long sequences): grid.z = window; W == 0: one window = the whole sequence
  int32_t W, nwin, summary, ndir;
  float* ws_state;     // [batch][ndir][nwin][dim][16]: pass 1 writes each window's end state from zero, the combine kernel
                       // turns it in place into the window's INCOMING state, pass 2 starts from it
  float* ws_sumdt;     // [batch][ndir][nwin][dim]: sum of Delta over the window (decay of the window = exp(A * sum))
  FwdDir dir[2];
};

#ifndef CM_FWDSP_IO
#define CM_FWDSP_IO 64
#endif
constexpr int kIO = CM_FWDSP_IO; // IO threads per direction (2 or 4 warps)
constexpr int kNBC = 128 / kIO;  // B/C quarter rows per IO thread and tile
constexpr int kIU = kT * kNP / kIO;   // (step, pair) units per IO thread and tile (4)
constexpr int kIKS = kIO / kNP;       // step distance between an IO thread's units (4)

struct FwdSmem {
  float4 dd[2][kT][kNP];      // (dt0, dt1, du0, du1) of a channel pair            [ring of 2 tiles]
  float bc[2][kT][32];        // B[0..15] | C[0..15]
  float p[kNW][kT][8][8];     // warp-private partial products of the tile in flight: [pair in warp][2*lane_in_pair + channel]
  float2 y[2][kT][kNP];       // sum over states of C*h per (step, channel pair)     [ring of 2 tiles]
  uint64_t in_full[2], in_empty[2], p_full[2], p_empty[2];   // p_* guard the y ring
};

enum { FM_UNI = 0, FM_STASH = 1, FM_COMBINE = 2 };

#ifndef CM_FWDSP_MINB
#define CM_FWDSP_MINB 3
#endif
#ifndef CM_FWDSP_SUB
#define CM_FWDSP_SUB 1
#endif
constexpr int kSub = CM_FWDSP_SUB;   // steps per software-pipelined sub-block of the recurrence

// ---- mbarrier primitives (shared::cta) ------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "MBAR_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra MBAR_DONE;\n\t"
      "bra MBAR_WAIT;\n\t"
      "MBAR_DONE:\n\t}"
      ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
// order-pinned shared load / ex2 (volatile: ptxas keeps volatile asm statements in program order)
__device__ __forceinline__ float4 lds128v(const void* p) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(smem_u32(p)));
  return v;
}
__device__ __forceinline__ float ex2v(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// one arrival per warp, after the warp's own shared-memory accesses are ordered
__device__ __forceinline__ void warp_arrive(uint64_t* b, int lane) {
  __syncwarp();
  if (lane == 0) mbar_arrive(b);
}

// Tile counter `it` runs over BOTH ranges of a direction; ring slot = it & 1, use count of a slot = it >> 1.
//   in_full[slot]  : IO warps  -> recurrence warps   (dd/bc of tile `it` are in shared memory)
//   in_empty[slot] : recurrence -> IO                 (tile `it` has been consumed; slot may be refilled with it+2)
//   p_full[slot]   : recurrence -> IO                 (y of tile `it` is complete)
//   p_empty[slot]  : IO -> recurrence                 (y of tile `it` has been consumed; slot free for it+2)
// The 8 partial products of a (step, channel pair) come from the 8 lanes of ONE warp, so the recurrence warps reduce
// them themselves after the tile (warp-private buffer, __syncwarp only) - the XU-bound warps have the issue slots.

// Processed-step ranges of this CTA's window for one direction: range 0 = [lo, mid), range 1 = [mid, hi).  In a
// bidirectional launch the ascending direction first covers the first half of the window's time span and the descending
// one the second half (stash), then each finishes the other half (combine).  Summary launches have a single range.
template <int NDIR>
__device__ __forceinline__ void window_ranges(const FwdParams& P, const FwdDir& d, int* lo, int* mid, int* hi) {
  const int L = P.L;
  if (P.W == 0) {
    *lo = 0; *hi = L; *mid = (NDIR == 2) ? d.s1 : L;
    return;
  }
  const int t0 = blockIdx.z * P.W, t1 = min(t0 + P.W, L), n = t1 - t0;
  const int first = (NDIR == 2 && !P.summary) ? (d.reverse ? n - cm_mid(n) : cm_mid(n)) : n;
  *lo = d.reverse ? L - t1 : t0;
  *hi = d.reverse ? L - t0 : t1;
  *mid = *lo + first;
}
__device__ __forceinline__ int64_t ws_row(const FwdParams& P, int b, int dir, int w) {
  return ((int64_t)b * P.ndir + dir) * P.nwin + w;
}

// ---- recurrence warps (2 per direction): lane = states 4m..4m+3 of channel pair pr -------------------------------
#ifndef CM_FWDSP_RTDIR
#define CM_FWDSP_RTDIR 1
#endif
// The direction is a run-time value for this role (it reads a handful of FwdDir fields, all before the time loop): as a
// template parameter a bidirectional CTA kept two copies of the unrolled 16-step recurrence live in the instruction cache.
template <typename T, int NDIR>
__device__ __forceinline__ void scan_role(const FwdParams& P, FwdSmem& S, const int gt, const int DIR) {
  const FwdDir& d = P.dir[DIR];
  const int warp = gt >> 5, lane = gt & 31;
  const int b = blockIdx.y;
  const int c_base = blockIdx.x * kCH;
  const int g = lane >> 2, m = lane & 3;
  const int pr = warp * 8 + g;
  const int c0 = c_base + 2 * pr;
  float2 kA[4], h[4];   // [state 4m + j] ; .x channel c0, .y channel c0 + 1
  {
    const float* A0 = d.A + (int64_t)c0 * d.A_sd;
    const float* A1 = A0 + d.A_sd;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      kA[j] = make_float2(__ldg(A0 + (4 * m + j) * d.A_sn) * kLog2e, __ldg(A1 + (4 * m + j) * d.A_sn) * kLog2e);
      h[j] = make_float2(0.f, 0.f);
    }
  }
  float* ckp = d.ckpt ? d.ckpt + b * d.ckpt_sb + (int64_t)c0 * d.ckpt_sd + 4 * m : nullptr;
  const bool summary = P.summary != 0;
  float* wsp = P.ws_state ? P.ws_state + (ws_row(P, b, DIR, blockIdx.z) * P.dim + c0) * 16 + 4 * m : nullptr;
  if (wsp != nullptr && !summary) {               // pass 2 of a windowed launch: the window's incoming state
    const float4 v0 = *reinterpret_cast<const float4*>(wsp), v1 = *reinterpret_cast<const float4*>(wsp + 16);
    h[0] = make_float2(v0.x, v1.x); h[1] = make_float2(v0.y, v1.y); h[2] = make_float2(v0.z, v1.z); h[3] = make_float2(v0.w, v1.w);
  }
  int s_lo, s_mid, s_hi;
  window_ranges<NDIR>(P, d, &s_lo, &s_mid, &s_hi);
  int it = 0;
#pragma unroll 1
  for (int range = 0; range < NDIR; ++range) {
    if (NDIR == 2 && range == 1) __syncthreads();   // partner's stash of the other half is complete
    const int s_begin = range == 0 ? s_lo : s_mid, s_end = (NDIR == 1 || range == 1) ? s_hi : s_mid;
    int jck = range == 0 ? 0 : cm_ceil_div(d.s1, CM_SCAN_CKPT_STEPS);   // next checkpoint slot (unwindowed launches only)
    const int nst = s_end - s_begin;
    const int ntile = nst > 0 ? cm_ceil_div(nst, kT) : 0;
#pragma unroll 1
    for (int t = 0; t < ntile; ++t, ++it) {
      const int slot = it & 1;
      const uint32_t par = (it >> 1) & 1;
      const int nvalid = nst - t * kT;    // steps of this tile inside the range (may exceed kT)
      mbar_wait(&S.in_full[slot], par);
      if (it >= 2 && !summary) mbar_wait(&S.p_empty[slot], par ^ 1);
      const float4* ddb = &S.dd[slot][0][pr];
      const float* bcb = &S.bc[slot][0][4 * m];
      float* pb = &S.p[warp][0][g][2 * m];
      float4 ddr[kSub], br[kSub], cr[kSub];
      auto lds_sub = [&](int sb) {
#pragma unroll
        for (int i = 0; i < kSub; ++i) {
          ddr[i] = ddb[(sb * kSub + i) * kNP];
          br[i] = *reinterpret_cast<const float4*>(bcb + (sb * kSub + i) * 32);
          cr[i] = *reinterpret_cast<const float4*>(bcb + (sb * kSub + i) * 32 + 16);
        }
      };
      lds_sub(0);
#pragma unroll
      for (int sb = 0; sb < kT / kSub; ++sb) {
        if (((sb * kSub) % CM_SCAN_CKPT_STEPS) == 0) {
          if (ckp != nullptr && sb * kSub < nvalid) {
            float* dst = ckp + (int64_t)jck * 16;
            *reinterpret_cast<float4*>(dst) = make_float4(h[0].x, h[1].x, h[2].x, h[3].x);
            *reinterpret_cast<float4*>(dst + d.ckpt_sd) = make_float4(h[0].y, h[1].y, h[2].y, h[3].y);
          }
          ++jck;
        }
        float2 a[kSub][4], ub[kSub][4];
        float4 cc[kSub];
#pragma unroll
        for (int i = 0; i < kSub; ++i) {
          const float2 dt = make_float2(ddr[i].x, ddr[i].y), du = make_float2(ddr[i].z, ddr[i].w);
          const float bb[4] = {br[i].x, br[i].y, br[i].z, br[i].w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            a[i][j] = fmul2(dt, kA[j]);
            ub[i][j] = fmul2(du, make_float2(bb[j], bb[j]));
          }
          cc[i] = cr[i];
        }
        if (sb + 1 < kT / kSub) lds_sub(sb + 1);
#pragma unroll
        for (int i = 0; i < kSub; ++i) {
#pragma unroll
          for (int j = 0; j < 4; ++j) a[i][j] = make_float2(ex2(a[i][j].x), ex2(a[i][j].y));
        }
#pragma unroll
        for (int i = 0; i < kSub; ++i) {
          const float c4[4] = {cc[i].x, cc[i].y, cc[i].z, cc[i].w};
#pragma unroll
          for (int j = 0; j < 4; ++j) h[j] = ffma2(a[i][j], h[j], ub[i][j]);
          if (!summary) {
            float2 pp = fmul2(make_float2(c4[0], c4[0]), h[0]);
#pragma unroll
            for (int j = 1; j < 4; ++j) pp = ffma2(make_float2(c4[j], c4[j]), h[j], pp);
            *reinterpret_cast<float2*>(pb + (sb * kSub + i) * 64) = pp;
          }
        }
      }
      __syncwarp();
      if (summary) {                                // state-only pass: nothing to hand to the IO warps
        if (lane == 0) mbar_arrive(&S.in_empty[slot]);
        continue;
      }
      // reduce this warp's 128 (step, pair) rows of 8 floats (4 lanes x 2 channels): lane -> rows lane + 32 j.  The two
      // 16-byte halves are read in an order that depends on the row, which keeps the LDS.128 conflict-free.
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int r = lane + 32 * j;
        const float4* row = reinterpret_cast<const float4*>(&S.p[warp][0][0][0] + r * 8);
        const int sw = (r >> 2) & 1;
        const float4 v0 = row[sw], v1 = row[sw ^ 1];
        const float2 acc = fadd2(fadd2(make_float2(v0.x, v0.y), make_float2(v0.z, v0.w)),
                                 fadd2(make_float2(v1.x, v1.y), make_float2(v1.z, v1.w)));
        S.y[slot][r >> 3][warp * 8 + (r & 7)] = acc;
      }
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&S.in_empty[slot]);
        mbar_arrive(&S.p_full[slot]);
      }
    }
  }
  if (summary) {                                  // pass 1: end state of the window started from zero
    *reinterpret_cast<float4*>(wsp) = make_float4(h[0].x, h[1].x, h[2].x, h[3].x);
    *reinterpret_cast<float4*>(wsp + 16) = make_float4(h[0].y, h[1].y, h[2].y, h[3].y);
    return;
  }
  const bool final_window = P.W == 0 || blockIdx.z == (d.reverse ? 0 : P.nwin - 1);
  if (d.last != nullptr && final_window) {
    float* ls = d.last + b * d.ls_sb + (int64_t)c0 * d.ls_sd + 4 * m * d.ls_sn;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      ls[j * d.ls_sn] = h[j].x;
      ls[d.ls_sd + j * d.ls_sn] = h[j].y;
    }
  }
}

// ---- IO warps (2 per direction): thread = channel pair cp at steps k0 + 4i of every tile ----------------------------
#ifndef CM_FWDSP_RTDIR_IO
#define CM_FWDSP_RTDIR_IO 0
#endif
template <typename T, int NDIR>
__device__ __forceinline__ void io_role(const FwdParams& P, FwdSmem& S, const int io, const int DIR) {
  using P2 = Pair<T>;
  using Q4 = Quad<T>;
  constexpr int ES = (int)sizeof(T);
  constexpr bool PRECISE = sizeof(T) == 4;
  const FwdDir& d = P.dir[DIR];
  const int lane = io & 31;
  const int b = blockIdx.y;
  const int c_base = blockIdx.x * kCH;
  const bool softplus = (P.flags & CM_FLAG_DELTA_SOFTPLUS) != 0;
  const bool has_z = d.z != nullptr;
  const int cp = io & (kNP - 1), k0 = io / kNP;
  const int cu = c_base + 2 * cp;
  float2 bias = make_float2(0.f, 0.f), Dsk = make_float2(0.f, 0.f);
  if (d.bias) bias = make_float2(__ldg(d.bias + cu), __ldg(d.bias + cu + 1));
  if (d.Dskip) Dsk = make_float2(__ldg(d.Dskip + cu), __ldg(d.Dskip + cu + 1));
  const char* pu = d.u + b * d.u_sb + cu * ES;
  const char* pdl = d.dl + b * d.dl_sb + cu * ES;
  const char* pz = has_z ? d.z + b * d.z_sb + cu * ES : nullptr;
  char* po = d.out + b * d.out_sb + cu * ES;
  char* ppre = d.pre ? d.pre + b * d.pre_sb + cu * ES : nullptr;
  // B/C quarter rows: rows kq + j*kIO/8 of the tile, quarter `part`
  const int kq = io >> 3, part = io & 7, isC = part >> 2, q4 = (part & 3) * 4;
  const char* pbc = (isC ? d.C + b * d.c_sb : d.B + b * d.b_sb) + q4 * ES;
  const int bc_ss = isC ? d.c_ss : d.b_ss;
  const int bc_dst = isC * 16 + q4;   // floats: row = B[0..15] | C[0..15]

  typename P2::Raw ru[kIU], rdl[kIU], uprev[kIU], rz[kIU], rst[kIU];
  typename Q4::Raw rbc[kNBC];
  const bool summary = P.summary != 0;
  float2 sumdt = make_float2(0.f, 0.f);
  int s_lo, s_mid, s_hi;
  window_ranges<NDIR>(P, d, &s_lo, &s_mid, &s_hi);
  int it = 0;
#pragma unroll 1
  for (int range = 0; range < NDIR; ++range) {
    if (NDIR == 2 && range == 1) __syncthreads();   // partner's stash of the other half is complete
    const int mode = (NDIR == 1) ? FM_UNI : (range == 0 ? FM_STASH : FM_COMBINE);
    const int s_begin = range == 0 ? s_lo : s_mid, s_end = (NDIR == 1 || range == 1) ? s_hi : s_mid;
    const int nst = s_end - s_begin;
    const int ntile = nst > 0 ? cm_ceil_div(nst, kT) : 0;
    const bool need_z = has_z && mode != FM_STASH;
    const bool need_st = mode == FM_COMBINE;

    auto load_raw = [&](int sb0) {
#pragma unroll
      for (int i = 0; i < kIU; ++i) {
        const int s = sb0 + k0 + i * kIKS;
        if (s < s_end) {
          ru[i] = P2::ld_nc(pu + (int64_t)s * d.u_ss);
          rdl[i] = P2::ld_nc(pdl + (int64_t)s * d.dl_ss);
        } else {
          ru[i] = P2::zero();
          rdl[i] = P2::zero();
        }
      }
#pragma unroll
      for (int j = 0; j < kNBC; ++j) {
        const int s = sb0 + kq + (kIO / 8) * j;
        rbc[j] = (s < s_end) ? Q4::ld_nc(pbc + (int64_t)s * bc_ss) : Q4::zero();
      }
    };
    auto convert_store = [&](int sb0, int slot) {
#pragma unroll
      for (int i = 0; i < kIU; ++i) {
        const int k = k0 + i * kIKS;
        const float2 u2 = P2::cvt(ru[i]);
        float2 dt = fadd2(P2::cvt(rdl[i]), bias);
        if (softplus) { dt.x = softplus_fwd<PRECISE>(dt.x); dt.y = softplus_fwd<PRECISE>(dt.y); }
        if (sb0 + k >= s_end) dt = make_float2(0.f, 0.f);   // a = 1, input 0: a missing step leaves the state unchanged
        sumdt = fadd2(sumdt, dt);
        const float2 du = fmul2(dt, u2);
        S.dd[slot][k][cp] = make_float4(dt.x, dt.y, du.x, du.y);
      }
#pragma unroll
      for (int j = 0; j < kNBC; ++j) {
        float v[4];
        Q4::cvt(rbc[j], v);
        *reinterpret_cast<float4*>(&S.bc[slot][kq + (kIO / 8) * j][bc_dst]) = make_float4(v[0], v[1], v[2], v[3]);
      }
    };
    // epilogue operands (gate, partner's stash) of the tile starting at sb0
    auto load_epi = [&](int sb0) {
#pragma unroll
      for (int i = 0; i < kIU; ++i) {
        const int s = sb0 + k0 + i * kIKS;
        const bool v = s < s_end;
        rz[i] = (need_z && v) ? P2::ld_nc(pz + (int64_t)s * d.z_ss) : P2::zero();
        rst[i] = (need_st && v) ? P2::ld_cg(po + (int64_t)s * d.out_ss) : P2::zero();
      }
    };
    // finishes tile (sb0, ring counter itf): skip term, stash / combine / gate, store
    auto epilogue = [&](int sb0, int itf) {
      const int slot = itf & 1;
      mbar_wait(&S.p_full[slot], (itf >> 1) & 1);
      float2 yv[kIU];
#pragma unroll
      for (int i = 0; i < kIU; ++i) yv[i] = S.y[slot][k0 + i * kIKS][cp];
      warp_arrive(&S.p_empty[slot], lane);
#pragma unroll
      for (int i = 0; i < kIU; ++i) {
        const int s = sb0 + k0 + i * kIKS;
        const float2 y = ffma2(Dsk, P2::cvt(uprev[i]), yv[i]);
        if (s < s_end) {
          char* dsto = po + (int64_t)s * d.out_ss;
          if (mode == FM_STASH) {
            P2::st(dsto, y);
          } else {
            const float2 tot = fadd2(y, P2::cvt(rst[i]));
            float2 val = fmul2(tot, make_float2(P.scale, P.scale));
            if (need_z) {
              const float2 zz = P2::cvt(rz[i]);
              val.x *= zz.x * sigmoid_sel<PRECISE>(zz.x);
              val.y *= zz.y * sigmoid_sel<PRECISE>(zz.y);
            }
            if (ppre != nullptr) P2::st(ppre + (int64_t)s * d.pre_ss, tot);
            P2::st(dsto, val);
          }
        }
      }
    };

    // every global load is issued one full iteration before its use
    if (ntile > 0) load_raw(s_begin);
#pragma unroll 1
    for (int t = 0; t < ntile; ++t, ++it) {
      const int slot = it & 1;
      const int sb0 = s_begin + t * kT;
      if (it >= 2) mbar_wait(&S.in_empty[slot], ((it >> 1) & 1) ^ 1);
      convert_store(sb0, slot);
      warp_arrive(&S.in_full[slot], lane);
      typename P2::Raw ucur[kIU];
#pragma unroll
      for (int i = 0; i < kIU; ++i) ucur[i] = ru[i];
      if (t + 1 < ntile) load_raw(sb0 + kT);
      if (!summary) {
        if (t > 0) epilogue(sb0 - kT, it - 1);      // uses uprev, rz, rst of tile t-1
        load_epi(sb0);
      }
#pragma unroll
      for (int i = 0; i < kIU; ++i) uprev[i] = ucur[i];
    }
    if (ntile > 0 && !summary) epilogue(s_begin + (ntile - 1) * kT, it - 1);
  }
  if (summary) {
    // sum of Delta over the window per channel: the kIKS threads that share a channel pair combine through shared memory
    // (the y ring is idle in this pass; the recurrence warps never touch it)
    float2* scratch = &S.y[0][0][0];
    if (DIR == 0) asm volatile("bar.sync 1, %0;" ::"n"(kIO) : "memory"); else asm volatile("bar.sync 2, %0;" ::"n"(kIO) : "memory");
    scratch[io] = sumdt;
    if (DIR == 0) asm volatile("bar.sync 1, %0;" ::"n"(kIO) : "memory"); else asm volatile("bar.sync 2, %0;" ::"n"(kIO) : "memory");
    if (io < kNP) {
      float2 a = scratch[io];
#pragma unroll
      for (int q = 1; q < kIKS; ++q) a = fadd2(a, scratch[io + q * kNP]);
      float* dst = P.ws_sumdt + ws_row(P, b, DIR, blockIdx.z) * P.dim + c_base + 2 * io;
      dst[0] = a.x; dst[1] = a.y;
    }
  }
}

template <typename T, int NDIR>
__global__ void __launch_bounds__(NDIR*(kGT + kIO), CM_FWDSP_MINB) scan_fwd_sp_kernel(const __grid_constant__ FwdParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  FwdSmem* S = reinterpret_cast<FwdSmem*>(smem_raw);
  const int tid = threadIdx.x;
  if (tid < NDIR) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&S[tid].in_full[i], kIO / 32);
      mbar_init(&S[tid].in_empty[i], kNW);
      mbar_init(&S[tid].p_full[i], kNW);
      mbar_init(&S[tid].p_empty[i], kIO / 32);
    }
  }
  __syncthreads();
  // thread layout: [recurrence dir 0 | recurrence dir 1 | IO dir 0 | IO dir 1]
  if (tid < NDIR * kGT) {
#if CM_FWDSP_RTDIR
    const int dir = (NDIR == 2 && tid >= kGT) ? 1 : 0;
    scan_role<T, NDIR>(P, S[dir], tid - dir * kGT, dir);
#else
    if (NDIR == 1 || tid < kGT) scan_role<T, NDIR>(P, S[0], tid, 0);
    else scan_role<T, NDIR>(P, S[1], tid - kGT, 1);
#endif
  } else {
    const int io = tid - NDIR * kGT;
#if CM_FWDSP_RTDIR_IO
    const int dir = (NDIR == 2 && io >= kIO) ? 1 : 0;
    io_role<T, NDIR>(P, S[dir], io - dir * kIO, dir);
#else
    if (NDIR == 1 || io < kIO) io_role<T, NDIR>(P, S[0], io, 0);
    else io_role<T, NDIR>(P, S[1], io - kIO, 1);
#endif
  }
}

template <typename T>
static bool t_ok(const cm_tensor3& t, int64_t quantum) {
  const int64_t es = sizeof(T);
  return t.ptr != nullptr && t.sd == 1 && (reinterpret_cast<uintptr_t>(t.ptr) % (quantum * es)) == 0 &&
         t.sb % quantum == 0 && t.sl % quantum == 0;
}

template <typename T>
static bool build_params(const cm_scan_fwd_args& a, FwdParams* P) {
  constexpr int ES = (int)sizeof(T);
  if (a.dstate != 16 || a.dim % kCH != 0) return false;
  if (!t_ok<T>(a.out, 2)) return false;
  if (a.z.ptr != nullptr && !t_ok<T>(a.z, 2)) return false;
  if (a.out_pre.ptr != nullptr && !t_ok<T>(a.out_pre, 2)) return false;
  P->L = a.seqlen;
  P->flags = a.flags;
  P->scale = a.out_scale;
  P->dim = a.dim;
  P->W = 0; P->nwin = 1; P->summary = 0; P->ndir = a.ndir;
  P->ws_state = nullptr; P->ws_sumdt = nullptr;
  for (int r = 0; r < a.ndir; ++r) {
    const cm_scan_dir& s = a.dir[r];
    FwdDir& d = P->dir[r];
    d.reverse = s.reverse != 0; d.pad0 = 0;
    if (s.bc_const) return false;
    if (!t_ok<T>(s.u, 2) || !t_ok<T>(s.delta, 2) || !t_ok<T>(s.Bm, 4) || !t_ok<T>(s.Cm, 4)) return false;
    if (s.ckpt != nullptr && ((reinterpret_cast<uintptr_t>(s.ckpt) & 15) != 0 || (s.ckpt_sb % 4) != 0 || (s.ckpt_sd % 4) != 0))
      return false;   // checkpoints are written as float4
    const bool rev = s.reverse != 0;
    const int64_t l0 = rev ? a.seqlen - 1 : 0;
    auto bp = [&](const cm_tensor3& t) { return static_cast<char*>(t.ptr) + l0 * t.sl * ES; };
    bool ok = true;
    ok &= step_stride(s.u.sl, ES, rev, a.seqlen, &d.u_ss);
    ok &= step_stride(s.delta.sl, ES, rev, a.seqlen, &d.dl_ss);
    ok &= step_stride(s.Bm.sl, ES, rev, a.seqlen, &d.b_ss);
    ok &= step_stride(s.Cm.sl, ES, rev, a.seqlen, &d.c_ss);
    ok &= step_stride(a.out.sl, ES, rev, a.seqlen, &d.out_ss);
    d.z_ss = d.pre_ss = 0;
    if (a.z.ptr) ok &= step_stride(a.z.sl, ES, rev, a.seqlen, &d.z_ss);
    if (a.out_pre.ptr) ok &= step_stride(a.out_pre.sl, ES, rev, a.seqlen, &d.pre_ss);
    if (!ok) return false;
    d.u = bp(s.u); d.dl = bp(s.delta); d.B = bp(s.Bm); d.C = bp(s.Cm);
    d.z = a.z.ptr ? bp(a.z) : nullptr;
    d.out = bp(a.out);
    d.pre = a.out_pre.ptr ? bp(a.out_pre) : nullptr;
    d.u_sb = s.u.sb * ES; d.dl_sb = s.delta.sb * ES; d.b_sb = s.Bm.sb * ES; d.c_sb = s.Cm.sb * ES;
    d.z_sb = a.z.sb * ES; d.out_sb = a.out.sb * ES; d.pre_sb = a.out_pre.sb * ES;
    d.s1 = cm_first_range(a.seqlen, a.ndir, s.reverse);
    d.A = s.A; d.A_sd = s.A_sd; d.A_sn = s.A_sn;
    d.Dskip = s.Dskip; d.bias = s.delta_bias;
    d.ckpt = s.ckpt; d.ckpt_sb = s.ckpt_sb; d.ckpt_sd = s.ckpt_sd;
    d.last = s.last_state; d.ls_sb = s.ls_sb; d.ls_sd = s.ls_sd; d.ls_sn = s.ls_sn;
  }
  return true;
}

// Serial pass over the windows of a row: turns each window's end-state-from-zero into its incoming state, in place.
__global__ void __launch_bounds__(256) window_combine_kernel(const FwdParams P, int batch) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)batch * P.ndir * P.dim * 16;
  if (idx >= total) return;
  const int n = (int)(idx & 15);
  const int c = (int)((idx >> 4) % P.dim);
  const int dir = (int)((idx / (16 * (int64_t)P.dim)) % P.ndir);
  const int b = (int)(idx / (16 * (int64_t)P.dim * P.ndir));
  const FwdDir& d = P.dir[dir];
  const float kA = __ldg(d.A + (int64_t)c * d.A_sd + n * d.A_sn) * kLog2e;
  float h = 0.f;
  for (int i = 0; i < P.nwin; ++i) {
    const int w = d.reverse ? P.nwin - 1 - i : i;        // processing order of the direction
    const int64_t row = ws_row(P, b, dir, w);
    float* sp_ = P.ws_state + (row * P.dim + c) * 16 + n;
    const float end0 = *sp_;
    const float sd = P.ws_sumdt[row * P.dim + c];
    *sp_ = h;
    h = fmaf(ex2(kA * sd), h, end0);
  }
}

// Windows over time pay the recurrence twice but multiply the number of CTAs: worth it when a whole-sequence launch
// leaves most SMs idle (inference on a few long utterances, BASELINE config 5).  Training launches (checkpoints) are
// never windowed: the backward kernel walks whole sequences.
static int plan_windows(const cm_scan_fwd_args& a, int* W) {
  *W = 0;
  for (int r = 0; r < a.ndir; ++r)
    if (a.dir[r].ckpt != nullptr) return 1;
  if (getenv("CM_SCAN_NO_WINDOWS") != nullptr) return 1;
  const int64_t ctas = (int64_t)(a.dim / kCH) * a.batch;
  const int64_t slots = 148 * 3;
  if (ctas * 2 > slots || a.seqlen < 1024) return 1;
  int want = (int)((slots + ctas - 1) / ctas);
  if (const char* e = getenv("CM_SCAN_WINDOWS")) want = atoi(e);   // A/B measurements
  if (want < 2) return 1;
  int w = (a.seqlen + want - 1) / want;
  w = (w + 63) / 64 * 64;
  if (w < 256) w = 256;
  const int nwin = (a.seqlen + w - 1) / w;
  if (nwin < 2) return 1;
  *W = w;
  return nwin;
}
static int64_t window_bytes(const cm_scan_fwd_args& a, int nwin) {
  return (int64_t)a.batch * a.ndir * nwin * a.dim * (16 + 1) * (int64_t)sizeof(float);
}

template <typename T, int NDIR>
static int launch_one(const FwdParams& P, const cm_scan_fwd_args& a, cudaStream_t st) {
  const size_t smem = sizeof(FwdSmem) * NDIR;
  auto kern = scan_fwd_sp_kernel<T, NDIR>;
  // per-device attribute, set on every launch (a host-side table lookup): a process-wide "done" flag would leave the
  // second GPU of a multi-device process at the 48 KB default
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  kern<<<dim3(a.dim / kCH, a.batch, P.nwin), NDIR * (kGT + kIO), smem, st>>>(P);
  CM_LAUNCH_CHECK();
  return 0;
}

template <typename T>
static int try_t(const cm_scan_fwd_args& a, cudaStream_t st, int* rc) {
  FwdParams P;
  if (!build_params<T>(a, &P)) return 0;
  int W = 0;
  const int nwin = plan_windows(a, &W);
  if (nwin > 1 && a.workspace != nullptr && a.workspace_bytes >= window_bytes(a, nwin) &&
      (reinterpret_cast<uintptr_t>(a.workspace) & 15) == 0) {
    P.W = W; P.nwin = nwin;
    P.ws_state = static_cast<float*>(a.workspace);
    P.ws_sumdt = P.ws_state + (int64_t)a.batch * a.ndir * nwin * a.dim * 16;
    P.summary = 1;
    *rc = (a.ndir == 2) ? launch_one<T, 2>(P, a, st) : launch_one<T, 1>(P, a, st);
    if (*rc) return 1;
    const int64_t total = (int64_t)a.batch * a.ndir * a.dim * 16;
    window_combine_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(P, a.batch);
    {
      cudaError_t e = cudaGetLastError();
      if (e != cudaSuccess) { *rc = (int)e; return 1; }
    }
    P.summary = 0;
  }
  *rc = (a.ndir == 2) ? launch_one<T, 2>(P, a, st) : launch_one<T, 1>(P, a, st);
  return 1;
}

}  // namespace sp

// bytes of caller-provided workspace that let cm_scan_fwd split this launch into time windows (0: not useful)
int64_t scan_fwd_sp_workspace_bytes(const cm_scan_fwd_args& a) {
  if (a.dstate != 16 || a.dim % sp::kCH != 0) return 0;
  int W = 0;
  const int nwin = sp::plan_windows(a, &W);
  return nwin > 1 ? sp::window_bytes(a, nwin) : 0;
}

// returns 1 if launched (result in *rc), 0 if the state-parallel path does not apply
int scan_fwd_try_state_parallel(const cm_scan_fwd_args& a, cudaStream_t st, int* rc) {
  switch (a.dtype) {
    case CM_F32: return sp::try_t<float>(a, st, rc);
    case CM_BF16: return sp::try_t<__nv_bfloat16>(a, st, rc);
    default: return sp::try_t<__half>(a, st, rc);
  }
}

}  // namespace cm
