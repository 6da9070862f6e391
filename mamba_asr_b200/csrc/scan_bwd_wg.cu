// Selective-scan backward, warpgroup-specialised kernel for sm_100a ("wg" path, round 2): the default for channel-last
// 16-state launches with TMA-addressable operands.  Same mathematics, checkpoint and partial-sum contracts as scan_bwd.cu
// (see its header; adjoint of modules/mamba/selective_scan_interface.py:106-157).
//
// What the round-1 kernel (scan_bwd_sp.cu) was bound by, from its ncu captures: 2 recurrence warps
// per SM sub-partition at 128 registers (the IO warps held as many registers as the recurrence warps), 58 shared-memory
// wavefronts per warp-step of which 26 were the partial sums written to and re-read from shared memory, and an IO role
// that issued a third of all instructions (64-bit address arithmetic and predicates of its global loads).  This kernel:
//
//   * CTA = TWO WARPGROUPS with their own register budgets (setmaxnreg): 4 recurrence warps at 120 registers, 4 IO warps at
//     40; 3 CTAs per SM = 12 recurrence warps (8 before).  A CTA owns 64 channels of one (batch, direction).
//   * ALL GLOBAL LOADS ARE TMA TILES (cp.async.bulk.tensor, box = 64 channels x 8 steps) issued by one thread three tiles
//     ahead: u, delta, dout, z, out_pre, the B|C rows and the tile's fp32 state checkpoint; rows outside [0, L) and channels
//     outside [0, dim) are zero-filled by the hardware, so no load has an address computation or a predicate.
//   * lane = states 4m..4m+3 of two adjacent channels, state pairs packed in one 64-bit register (FFMA2 / FMUL2 operate on
//     two states; dt, dt*u and dy are broadcast operands; B and C pairs come straight out of LDS.128).
//   * THE SUMS OVER STATES (r1 = sum_n lambda*B, r2 = sum_n lambda*(a h)*A: 4 lanes of a channel pair) ARE DONE BY THE
//     TENSOR CORE: mma.sync m16n8k8 TF32 with the value split into two TF32 terms (hi + lo, |error| < 2^-21) as the A
//     operand and a one-hot column selector as B, accumulated over the tile - after 8 steps lane (pair, m) holds r1, r2 of
//     steps m and m+4 for both channels, and finishes du / ddelta / dD / d(bias) for them itself (no result ring, no
//     second trip through shared memory).  HMMA.1688 runs on its own pipe at the rate of one MUFU (measured:
//     tools/ubench/mma_red.cu).
//   * THE SUMS OVER CHANNELS (dB, dC: 8 channel pairs of a warp) are a recursive-halving exchange over the three pair
//     bits of the lane index (7 SHFL + 7 FADD + 14 SEL per step instead of 12 STS/LDS wavefronts + a transposed re-read);
//     one 128-byte row per warp and step goes to shared memory, the IO warps add the four warps and write ONE partial
//     row per (batch, 64-channel slab, step): half the partial traffic of the 32-channel slabs.
//     Deterministic: fixed-order sums, no atomics (the reference kernel accumulates dB/dC with fp32 atomics).
//   * history: the recomputed states of the last 4 steps of a tile stay in registers, the first 4 are parked in a
//     warp-private shared buffer (as before); a*h_{k-1} is obtained as h_k - du*B (never divides by a decay).
//
// Requirements (else cm_scan_bwd falls through to scan_bwd_sp.cu / scan_bwd_cl.cu / scan_bwd.cu): unit channel stride,
// dstate == 16, dim a multiple of 32 and >= 64, variable B/C as one 32-element B|C row per step, 16-byte aligned rows,
// lanes_per_channel in {0, 1}; the caller sizes the dB/dC partial tensor with cm_scan_bwd_slab_channels() (64).
#include <climits>
#include <cstdlib>

#include "common.cuh"
#include "sp_common.cuh"
#include "tma.cuh"

namespace cm {
namespace wgb {

using cm::sp::Pair;

constexpr int kTB = CM_SCAN_CKPT_STEPS;   // steps per tile (8)
constexpr int kRW = 4;                    // recurrence warps (warpgroup 0)
constexpr int kIW = 4;                    // IO warps (warpgroup 1)
#ifndef CM_BWDWG_IA
#define CM_BWDWG_IA 2
#endif
constexpr int kIA = CM_BWDWG_IA;          // IO warps that work (the others only give their registers away and exit)
constexpr int kRT = kRW * 32, kIT = kIA * 32;
constexpr int kCH = 64;                   // channels per CTA (= slab width of the dB/dC partial tensor)
constexpr int kNP = kCH / 2;              // channel pairs per CTA
constexpr int kRS = 2;                    // raw (TMA) ring depth
#ifndef CM_BWDWG_HREG
#define CM_BWDWG_HREG 4
#endif
constexpr int kHR = CM_BWDWG_HREG;        // last kHR steps of a tile keep their recomputed states in registers
#ifndef CM_BWDWG_RREG
#define CM_BWDWG_RREG 120
#endif
#ifndef CM_BWDWG_IREG
#define CM_BWDWG_IREG 40
#endif
#ifndef CM_BWDWG_FENCE
#define CM_BWDWG_FENCE 2
#endif
#ifndef CM_BWDWG_PIPE
#define CM_BWDWG_PIPE 0   // decays of the next step evaluated one step ahead: measured 0.569 vs 0.556 ms (ptxas already overlaps them)
#endif
constexpr float kLn2f = 0.6931471805599453f;
// -DCM_ABL_*: timing ablations (tools/_run_r2n.sh); they change the results and are never part of the product build
#ifdef CM_ABL_NOEX2
__device__ __forceinline__ float ex2r(float x) { return x + 1.0f; }
#else
__device__ __forceinline__ float ex2r(float x) { return ex2(x); }
#endif
// dB / dC exchange buffer of a warp: the dC block of a step starts 144 floats after its dB block (128 + a 16-float pad, so
// that the half-warp that reads dB and dC pairs side by side hits 32 distinct banks), the second step 272 floats after the first
constexpr int kPbWhich = 8 * 16 + 16, kPbStep = kPbWhich + 8 * 16;
static_assert(kTB == 8, "the column selector of the state sums assumes 8-step tiles");

struct BwdDir {
  CUtensorMap m_u, m_dl, m_bc, m_ck;
  char *du, *ddl;                          // byte pointers at (batch 0, channel 0, PROCESSED step 0)
  int64_t du_sb, ddl_sb;
  int32_t du_ss, ddl_ss;                   // bytes per processed step (signed)
  int32_t s1, reverse, write_dz, pad0;
  const float* A;
  int64_t A_sd, A_sn;
  const float *Dskip, *bias;
  float *dBC_part, *dA_part, *dD_part, *dbias_part;
  int64_t part_l0;                         // float offset of processed step 0 inside one [L][32] slab
  int32_t part_ss, pad1;                   // floats per processed step (+-32)
};
struct alignas(64) BwdParams {
  CUtensorMap m_go, m_z, m_pre;            // dout, z, out_pre
  BwdDir dir[2];
  char* dz;                                // byte pointer at (batch 0, channel 0, time 0)
  int64_t dz_sb;
  int32_t dz_sl;                           // bytes per TIME step
  int32_t L, ndir, n_slab, dim, has_z;
  uint32_t flags;
  float scale;
};

template <typename T>
struct alignas(128) Raw {                  // one TMA stage: [step (ascending time)][channel]
  T u[kTB][kCH], dl[kTB][kCH], go[kTB][kCH], z[kTB][kCH], pre[kTB][kCH];
  T bc[kTB][32];
};
struct alignas(128) Ops {                  // one operand slot (fp32, processed order)
  float ck[2][kNP][16];                    // state checkpoint of the tile: [channel parity][pair][state]      (TMA)
  float4 dd[kTB][kNP];                     // (dt0, dt1, dt0*u0, dt1*u1)
  float4 us[kTB][kNP];                     // (u0, u1, sigmoid(delta+bias)0, 1)
  float2 dy[kTB][kNP];                     // gated output gradient of the pair
  float bc[kTB][32];                       // B[0..15] | C[0..15]
};
template <typename T>
struct alignas(128) Smem {
  Raw<T> raw[kRS];
  Ops ops[2];
  float4 hs[kRW][kTB - kHR > 0 ? kTB - kHR : 1][2][32];   // recomputed states of the first steps of the tile in flight
  float pb[kRW][2 * kPbStep];              // per-lane dB[4] / dC[4] of two steps: [step][dB|dC][pair][state], see kPbWhich
  float bcw[2][kRW][kTB][32];              // per-warp sums over its 16 channels: dB[16] | dC[16]     [result ring, 2 tiles]
  uint64_t raw_full[kRS], ck_full[2], in_full[2], in_empty[2], out_full[2], out_empty[2];
};

// ---- mbarrier helpers (labels are scoped by the braces) -----------------------------------------------------------------
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WG_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WG_DONE;\n\t"
      "bra WG_WAIT;\n\t"
      "WG_DONE:\n\t}"
      ::"r"(tma::smem_u32(b)), "r"(parity) : "memory");
}
// the same with a suspend-time hint: the IO warps wait a whole tile (microseconds) for the recurrence warps - without the
// hint the try_wait returns after a short system-defined interval and the retry loop of the two idle warps issued 12 % of
// all instructions of the kernel (profiles/r02_scan_bwd_wg_cfg3_v1_ncu.txt)
__device__ __forceinline__ void mbar_wait_long(uint64_t* b, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WGL_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
      "@p bra WGL_DONE;\n\t"
      "bra WGL_WAIT;\n\t"
      "WGL_DONE:\n\t}"
      ::"r"(tma::smem_u32(b)), "r"(parity), "r"(100000u) : "memory");
}
__device__ __forceinline__ void warp_arrive(uint64_t* b, int lane) {
  __syncwarp();
  if (lane == 0) tma::mbar_arrive(b);
}

// D += A * B, m16n8k8 TF32 (A row-major 16x8, B col-major 8x8, fp32 accumulate)
__device__ __forceinline__ void mma_tf32(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b) {
  asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%8}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b));
}
// Sum of x (rows g) and y (rows g + 8) over the four lanes of a quad, added into column `sel` of the accumulator: the value
// is split into its upper 19 bits (exact in TF32) and the remainder, which the tensor core truncates to TF32 itself.
// 16-bit I/O (PRECISE = false): one TF32 term rounded to nearest (relative error 2^-11, an eighth of the output's ulp).
template <bool PRECISE>
__device__ __forceinline__ void quad_sum_mma(float (&acc)[4], float x, float y, uint32_t sel) {
  if (PRECISE) {
    const uint32_t hx = __float_as_uint(x) & 0xffffe000u, hy = __float_as_uint(y) & 0xffffe000u;
    const float lx = x - __uint_as_float(hx), ly = y - __uint_as_float(hy);
    mma_tf32(acc, hx, hy, __float_as_uint(lx), __float_as_uint(ly), sel);
  } else {
    // round to nearest by adding half a TF32 ulp to the magnitude (the tensor core drops the low 13 bits); cvt.rna.tf32
    // expands to four instructions with an Inf / NaN test
    mma_tf32(acc, __float_as_uint(x) + 0x1000u, __float_as_uint(y) + 0x1000u, 0u, 0u, sel);
  }
}

// Tiles are visited in reverse processing order: range 1 (if bidirectional) last tile first, then range 0.
struct TileSeq {
  int n1, n0, s1, L, j1;
  __device__ __forceinline__ TileSeq(int L_, int ndir, int s1_) {
    L = L_;
    s1 = ndir == 2 ? s1_ : L_;
    n0 = cm_ceil_div(s1, kTB);
    n1 = ndir == 2 ? cm_ceil_div(L - s1, kTB) : 0;
    j1 = n0;
  }
  __device__ __forceinline__ int total() const { return n0 + n1; }
  // -> first processed step of tile i, one-past-last step of its range, checkpoint slot
  __device__ __forceinline__ void get(int i, int* sb0, int* s_end, int* slot) const {
    if (i < n1) {
      const int tt = n1 - 1 - i;
      *sb0 = s1 + tt * kTB; *s_end = L; *slot = j1 + tt;
    } else {
      const int tt = n0 - 1 - (i - n1);
      *sb0 = tt * kTB; *s_end = s1; *slot = tt;
    }
  }
};

// ---- recurrence warps --------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ void scan_role(const BwdParams& P, Smem<T>& S, const int gt, const int dir) {
  using P2 = Pair<T>;
  constexpr int ES = (int)sizeof(T);
  constexpr bool PRECISE = sizeof(T) == 4;
  const BwdDir& d = P.dir[dir];
  const int warp = gt >> 5, lane = gt & 31;
  const int b = blockIdx.y;
  const int c_base = blockIdx.x * kCH;
  const int g = lane >> 2, m = lane & 3;
  const float gf = (float)g;
  const int lp = warp * 8 + g;              // channel pair inside the CTA
  const int c0 = c_base + 2 * lp;
  const bool ch_ok = c0 < P.dim;            // dim is a multiple of 32: the last CTA may own 32 channels only
  // [channel c][state pair i]: states 4m + 2i, 4m + 2i + 1 of channel c0 + c
  float2 kA[2][2], mu[2][2], dA[2][2];
  {
    const float* A0 = d.A + (int64_t)(ch_ok ? c0 : 0) * d.A_sd + (int64_t)(4 * m) * d.A_sn;
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const float* a = A0 + c * d.A_sd + (2 * i) * d.A_sn;
        kA[c][i] = make_float2(__ldg(a) * kLog2e, __ldg(a + d.A_sn) * kLog2e);
        mu[c][i] = make_float2(0.f, 0.f);
        dA[c][i] = make_float2(0.f, 0.f);
      }
  }
  float2 Dsk = make_float2(0.f, 0.f);
  if (d.Dskip && ch_ok) Dsk = make_float2(__ldg(d.Dskip + c0), __ldg(d.Dskip + c0 + 1));
  char* pdu = d.du + b * d.du_sb + (int64_t)c0 * ES;
  char* pddl = d.ddl + b * d.ddl_sb + (int64_t)c0 * ES;
  float2 dD_acc = make_float2(0.f, 0.f), db_acc = make_float2(0.f, 0.f);
  // one-hot selector columns of the tensor-core state sums: step k goes to column 2*(k & 3) + (k >> 2), so that lane
  // (pair g, m) ends up with steps m (accumulator slots 0 / 2) and m + 4 (slots 1 / 3)
  // sums over the warp's 8 channel pairs (dB, dC): two steps at a time through a warp-private buffer; lane -> (step parity,
  // dB|dC, two adjacent states)
  float* const pbw = &S.pb[warp][g * 16 + 4 * m];
  const int rsp = lane >> 4, rwhich = (lane >> 3) & 1, rn = (lane & 7) * 2;
  const float* const pbr = &S.pb[warp][rsp * kPbStep + rwhich * kPbWhich + rn];
  const int bcw_off = rsp * 32 + rwhich * 16 + rn;

  const TileSeq seq(P.L, P.ndir, d.s1);
  const int ntot = seq.total();
#pragma unroll 1
  for (int i = 0; i < ntot; ++i) {
    const int slot = i & 1;
    const uint32_t par = (i >> 1) & 1;
    Ops& O = S.ops[slot];
    int sb0, s_end, cslot;
    seq.get(i, &sb0, &s_end, &cslot);
    mbar_wait(&S.ck_full[slot], par);
    float2 h[2][2];
    {
      const float4 k0 = *reinterpret_cast<const float4*>(&O.ck[0][lp][4 * m]);
      const float4 k1 = *reinterpret_cast<const float4*>(&O.ck[1][lp][4 * m]);
      h[0][0] = make_float2(k0.x, k0.y); h[0][1] = make_float2(k0.z, k0.w);
      h[1][0] = make_float2(k1.x, k1.y); h[1][1] = make_float2(k1.z, k1.w);
    }
    mbar_wait(&S.in_full[slot], par);
    const float4* ddb = &O.dd[0][lp];
    const float2* dyb = &O.dy[0][lp];
    const float* bcb = &O.bc[0][4 * m];
    // ---- forward: recompute the states of the tile
    float2 hist[kHR > 0 ? kHR : 1][2][2];
#ifdef CM_ABL_NOFWD
#pragma unroll
    for (int k = 0; k < kHR; ++k)
#pragma unroll
      for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int j = 0; j < 2; ++j) hist[k][c][j] = h[c][j];
    if (mu[0][0].x == 123.456f)
#endif
#if CM_BWDWG_PIPE
    // software pipeline: the decays a = 2^(dt * kA) of step k + 1 are evaluated (LDS -> FMUL2 -> MUFU) while step k's state
    // update runs, so that neither the shared-memory nor the MUFU latency sits in front of the dependent FFMA2 chain
    float2 an[2][2];
    float4 ddn = ddb[0];
    auto decays = [&](const float4& dd_, float2 (&o)[2][2]) {
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const float dtc = c ? dd_.y : dd_.x;
        const float2 dt2 = make_float2(dtc, dtc);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const float2 x = fmul2(dt2, kA[c][j]);
          o[c][j] = make_float2(ex2r(x.x), ex2r(x.y));
        }
      }
    };
    decays(ddn, an);
#endif
#pragma unroll
    for (int k = 0; k < kTB; ++k) {
#if CM_BWDWG_PIPE
      const float4 dd = ddn;
      float2 ac[2][2];
#pragma unroll
      for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int j = 0; j < 2; ++j) ac[c][j] = an[c][j];
      if (k + 1 < kTB) { ddn = ddb[(k + 1) * kNP]; decays(ddn, an); }
#else
      const float4 dd = ddb[k * kNP];
#endif
      const float4 bb = *reinterpret_cast<const float4*>(bcb + k * 32);
      const float2 Bp[2] = {make_float2(bb.x, bb.y), make_float2(bb.z, bb.w)};
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const float dtc = c ? dd.y : dd.x, duc = c ? dd.w : dd.z;
        const float2 dt2 = make_float2(dtc, dtc), du2 = make_float2(duc, duc);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
#if CM_BWDWG_PIPE
          const float2 a = ac[c][j];
          (void)dt2;
#else
          const float2 x = fmul2(dt2, kA[c][j]);
          const float2 a = make_float2(ex2r(x.x), ex2r(x.y));
#endif
          h[c][j] = ffma2(a, h[c][j], fmul2(du2, Bp[j]));
        }
      }
      if (k >= kTB - kHR) {
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
          for (int j = 0; j < 2; ++j) hist[k - (kTB - kHR)][c][j] = h[c][j];
      } else {
        S.hs[warp][k][0][lane] = make_float4(h[0][0].x, h[0][0].y, h[0][1].x, h[0][1].y);
        S.hs[warp][k][1][lane] = make_float4(h[1][0].x, h[1][0].y, h[1][1].x, h[1][1].y);
      }
    }
    if (i >= 2) mbar_wait(&S.out_empty[slot], par ^ 1);   // the IO warps have drained bcw[slot] of tile i - 2
    float* bcw = &S.bcw[slot][warp][0][bcw_off];
    float R1[4] = {0.f, 0.f, 0.f, 0.f}, R2[4] = {0.f, 0.f, 0.f, 0.f};
    float2 pv[8];                                          // dB / dC exchange of a step pair: summed one step later
    auto pb_sum = [&](int kk) {                            // kk = the even step of the pair whose values are in pv
      float2 acc = fadd2(fadd2(pv[0], pv[1]), fadd2(pv[2], pv[3]));
      acc = fadd2(acc, fadd2(fadd2(pv[4], pv[5]), fadd2(pv[6], pv[7])));
      *reinterpret_cast<float2*>(bcw + kk * 32) = acc;
    };
    // ---- reverse sweep.  lambda = dy*C + mu ; mu <- a*lambda ; the adjoint of the decay is q = lambda * a * h_{k-1} = mu_new * h_{k-1}
    // (h_{k-1}: the previous entry of the history, the checkpoint for the first step of the tile)
    auto hist_at = [&](int j, float2 (&o)[2][2]) {         // recomputed state after step j of the tile; j = -1: the checkpoint
      if (j >= kTB - kHR) {
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
          for (int q = 0; q < 2; ++q) o[c][q] = hist[j - (kTB - kHR)][c][q];
      } else {
        const float4 h0 = j >= 0 ? S.hs[warp][j >= 0 ? j : 0][0][lane] : *reinterpret_cast<const float4*>(&O.ck[0][lp][4 * m]);
        const float4 h1 = j >= 0 ? S.hs[warp][j >= 0 ? j : 0][1][lane] : *reinterpret_cast<const float4*>(&O.ck[1][lp][4 * m]);
        o[0][0] = make_float2(h0.x, h0.y); o[0][1] = make_float2(h0.z, h0.w);
        o[1][0] = make_float2(h1.x, h1.y); o[1][1] = make_float2(h1.z, h1.w);
      }
    };
    float2 hk[2][2];
    hist_at(kTB - 1, hk);
#if CM_BWDWG_PIPE
    float4 ddr = ddb[(kTB - 1) * kNP];
    float2 ar[2][2];
    decays(ddr, ar);
#endif
#pragma unroll
    for (int k = kTB - 1; k >= 0; --k) {
#if CM_BWDWG_FENCE > 0
      if ((k % CM_BWDWG_FENCE) == CM_BWDWG_FENCE - 1) asm volatile("" ::: "memory");   // scheduling fence: bounds how far ptxas hoists operand loads
#endif
#if CM_BWDWG_PIPE
      const float4 dd = ddr;
      float2 acur[2][2];
#pragma unroll
      for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int j = 0; j < 2; ++j) acur[c][j] = ar[c][j];
      if (k > 0) { ddr = ddb[(k - 1) * kNP]; decays(ddr, ar); }
#else
      const float4 dd = ddb[k * kNP];
#endif
      const float2 dy = dyb[k * kNP];
      const float4 bb = *reinterpret_cast<const float4*>(bcb + k * 32);
      const float4 cc = *reinterpret_cast<const float4*>(bcb + k * 32 + 16);
      const float2 Bp[2] = {make_float2(bb.x, bb.y), make_float2(bb.z, bb.w)};
      const float2 Cp[2] = {make_float2(cc.x, cc.y), make_float2(cc.z, cc.w)};
      float2 hp[2][2];
      hist_at(k - 1, hp);
      float2 accB[2], accC[2], r1a[2], r2a[2];
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const float dtc = c ? dd.y : dd.x, duc = c ? dd.w : dd.z, dyc = c ? dy.y : dy.x;
        const float2 dt2 = make_float2(dtc, dtc), du2 = make_float2(duc, duc);
        const float2 dy2 = make_float2(dyc, dyc);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
#if CM_BWDWG_PIPE
          const float2 a = acur[c][j];
#else
          const float2 x = fmul2(dt2, kA[c][j]);
          const float2 a = make_float2(ex2r(x.x), ex2r(x.y));
#endif
          const float2 lam = ffma2(dy2, Cp[j], mu[c][j]);
          accC[j] = c == 0 ? fmul2(dy2, hk[c][j]) : ffma2(dy2, hk[c][j], accC[j]);
          accB[j] = c == 0 ? fmul2(du2, lam) : ffma2(du2, lam, accB[j]);
          mu[c][j] = fmul2(a, lam);
          const float2 q = fmul2(mu[c][j], hp[c][j]);
          dA[c][j] = ffma2(q, dt2, dA[c][j]);
          r1a[c] = j == 0 ? fmul2(lam, Bp[j]) : ffma2(lam, Bp[j], r1a[c]);
          r2a[c] = j == 0 ? fmul2(q, kA[c][j]) : ffma2(q, kA[c][j], r2a[c]);
        }
      }
#pragma unroll
      for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int j = 0; j < 2; ++j) hk[c][j] = hp[c][j];
      // sums over the 16 states of a channel: 4 in the lane, 4 lanes on the tensor core
      {
        uint32_t sel;                                      // 1.0f in the lanes that hold column 2*(k & 3) + (k >> 2) of B
        asm("set.eq.f32.f32 %0, %1, %2;" : "=r"(sel) : "f"(gf), "f"((float)(2 * (k & 3) + (k >> 2))));
#ifdef CM_ABL_NOMMA
        R1[k & 3] += r1a[0].x + r1a[0].y + r1a[1].x + r1a[1].y + __uint_as_float(sel);
        R2[k & 3] += r2a[0].x + r2a[0].y + r2a[1].x + r2a[1].y;
#else
        quad_sum_mma<PRECISE>(R1, r1a[0].x + r1a[0].y, r1a[1].x + r1a[1].y, sel);
        quad_sum_mma<PRECISE>(R2, r2a[0].x + r2a[0].y, r2a[1].x + r2a[1].y, sel);
#endif
      }
      // sums over the warp's 8 channel pairs
#ifdef CM_ABL_NOPB
      if (accB[0].x == 123.456f)
#endif
      *reinterpret_cast<float4*>(pbw + (k & 1) * kPbStep) = make_float4(accB[0].x, accB[0].y, accB[1].x, accB[1].y);
#ifdef CM_ABL_NOPB
      if (accC[0].x == 123.456f)
#endif
      *reinterpret_cast<float4*>(pbw + (k & 1) * kPbStep + kPbWhich) = make_float4(accC[0].x, accC[0].y, accC[1].x, accC[1].y);
#ifdef CM_ABL_NOPB
      if (k == 0) { for (int gg = 0; gg < 8; ++gg) pv[gg] = accB[0]; pb_sum(0); }
      if (false) {
#else
      if ((k & 1) == 0) {
#endif
        __syncwarp();
#pragma unroll
        for (int gg = 0; gg < 8; ++gg) pv[gg] = *reinterpret_cast<const float2*>(pbr + gg * 16);
        __syncwarp();
        if (k == 0) pb_sum(0);
      } else if (k < kTB - 1) {
        pb_sum(k + 1);
      }
    }
    // ---- steps m and m + 4 of the tile: du, ddelta, and the running per-channel sums
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int k = m + 4 * e;
      const int s = sb0 + k;
      const float4 dd = ddb[k * kNP];
      const float4 us = O.us[k][lp];
      const float2 dy = dyb[k * kNP];
      const float2 r1 = make_float2(R1[e], R1[2 + e]), r2 = make_float2(R2[e], R2[2 + e]);
      const float2 dt = make_float2(dd.x, dd.y), u2 = make_float2(us.x, us.y), sig = make_float2(us.z, us.w);
      const float2 duo = ffma2(dt, r1, fmul2(dy, Dsk));
      const float2 ddt = ffma2(r2, make_float2(kLn2f, kLn2f), fmul2(u2, r1));
      const float2 ddl = fmul2(ddt, sig);
      if (s < s_end) {
        dD_acc = ffma2(dy, u2, dD_acc);
        db_acc = fadd2(db_acc, ddl);
        if (ch_ok) {
          P2::st(pdu + (int64_t)s * d.du_ss, duo);
          P2::st(pddl + (int64_t)s * d.ddl_ss, ddl);
        }
      }
    }
    __syncwarp();
    if (lane == 0) {
      tma::mbar_arrive(&S.in_empty[slot]);
      tma::mbar_arrive(&S.out_full[slot]);
    }
  }
  // ---- per-row sums over time
  // dD, d(delta_bias): the four lanes of a pair hold disjoint steps
  dD_acc.x += __shfl_xor_sync(0xffffffffu, dD_acc.x, 1); dD_acc.y += __shfl_xor_sync(0xffffffffu, dD_acc.y, 1);
  db_acc.x += __shfl_xor_sync(0xffffffffu, db_acc.x, 1); db_acc.y += __shfl_xor_sync(0xffffffffu, db_acc.y, 1);
  dD_acc.x += __shfl_xor_sync(0xffffffffu, dD_acc.x, 2); dD_acc.y += __shfl_xor_sync(0xffffffffu, dD_acc.y, 2);
  db_acc.x += __shfl_xor_sync(0xffffffffu, db_acc.x, 2); db_acc.y += __shfl_xor_sync(0xffffffffu, db_acc.y, 2);
  if (ch_ok) {
    const int64_t row = (int64_t)b * P.dim + c0;
    float* da = d.dA_part + row * 16 + 4 * m;
    *reinterpret_cast<float4*>(da) = make_float4(dA[0][0].x, dA[0][0].y, dA[0][1].x, dA[0][1].y);
    *reinterpret_cast<float4*>(da + 16) = make_float4(dA[1][0].x, dA[1][0].y, dA[1][1].x, dA[1][1].y);
    if (m == 0) {
      if (d.dD_part) { d.dD_part[row] = dD_acc.x; d.dD_part[row + 1] = dD_acc.y; }
      if (d.dbias_part) { d.dbias_part[row] = db_acc.x; d.dbias_part[row + 1] = db_acc.y; }
    }
  }
}

// ---- IO warps ----------------------------------------------------------------------------------------------------------
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

template <typename T> struct SmemQuad;    // four adjacent elements of a raw tile -> float4
template <> struct SmemQuad<float> {
  static __device__ __forceinline__ float4 ld(const float* p) { return *reinterpret_cast<const float4*>(p); }
};
template <> struct SmemQuad<__nv_bfloat16> {
  static __device__ __forceinline__ float4 ld(const __nv_bfloat16* p) {
    const uint2 r = *reinterpret_cast<const uint2*>(p);
    return make_float4(__uint_as_float(r.x << 16), __uint_as_float(r.x & 0xffff0000u), __uint_as_float(r.y << 16),
                       __uint_as_float(r.y & 0xffff0000u));
  }
};
template <> struct SmemQuad<__half> {
  static __device__ __forceinline__ float4 ld(const __half* p) {
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(p)), b = __half22float2(*reinterpret_cast<const __half2*>(p + 2));
    return make_float4(a.x, a.y, b.x, b.y);
  }
};

template <typename T, bool SOFTPLUS, bool HAS_Z>
__device__ __forceinline__ void io_role(const BwdParams& P, Smem<T>& S, const int io, const int dir) {
  using P2 = Pair<T>;
  constexpr int ES = (int)sizeof(T);
  constexpr bool PRECISE = sizeof(T) == 4;
  constexpr uint32_t kRawBytes = (uint32_t)((HAS_Z ? 5 : 3) * kTB * kCH * ES + kTB * 32 * ES);
  constexpr uint32_t kCkBytes = (uint32_t)(2 * kNP * 16 * sizeof(float));
  constexpr int kUnits = kTB * kNP / kIT;               // (step, channel pair) units per thread and tile (4)
  constexpr int kStepInc = kIT / kNP;                   // step distance between a thread's units (2)
  const BwdDir& d = P.dir[dir];
  const int lane = io & 31;
  const int b = blockIdx.y;
  const int c_base = blockIdx.x * kCH;
  const bool rev = d.reverse != 0;
  const bool do_dz = HAS_Z && d.write_dz != 0;
  const int cp = io & (kNP - 1), k0 = io / kNP;         // unit e = (step k0 + kStepInc * e, channel pair cp)
  const int cu = c_base + 2 * cp;
  const bool ch_ok = cu < P.dim;
  float2 bias = make_float2(0.f, 0.f);
  if (d.bias && ch_ok) bias = make_float2(__ldg(d.bias + cu), __ldg(d.bias + cu + 1));
  char* pdz = do_dz ? P.dz + b * P.dz_sb + (int64_t)cu * ES : nullptr;
  const int bc_row = io >> 3, bc_col = (io & 7) * 4;    // B|C conversion: (step, four columns) per thread
  // dB/dC partial rows of this CTA's 64-channel slab: thread -> (step, float4 column)
  const int pk = io >> 3, part = io & 7;
  float* ppart = d.dBC_part + ((int64_t)b * P.n_slab + blockIdx.x) * (int64_t)P.L * 32 + d.part_l0 + 4 * part;
  const float scale = P.scale;
  const int L = P.L;

  const TileSeq seq(P.L, P.ndir, d.s1);
  const int ntot = seq.total();

  auto issue_raw = [&](int i) {           // thread 0
    int sb0, s_end, cslot;
    seq.get(i, &sb0, &s_end, &cslot);
    const int st = i & 1;
    const int t0 = rev ? (L - sb0 - kTB) : sb0;
    Raw<T>& R = S.raw[st];
    uint64_t* bar = &S.raw_full[st];
    tma::mbar_expect_tx(bar, kRawBytes);
    tma::load_3d(&R.u[0][0], &d.m_u, bar, c_base, t0, b);
    tma::load_3d(&R.dl[0][0], &d.m_dl, bar, c_base, t0, b);
    tma::load_3d(&R.go[0][0], &P.m_go, bar, c_base, t0, b);
    if (HAS_Z) {
      tma::load_3d(&R.z[0][0], &P.m_z, bar, c_base, t0, b);
      tma::load_3d(&R.pre[0][0], &P.m_pre, bar, c_base, t0, b);
    }
    tma::load_3d(&R.bc[0][0], &d.m_bc, bar, 0, t0, b);
  };
  auto issue_ck = [&](int i) {            // thread 0
    int sb0, s_end, cslot;
    seq.get(i, &sb0, &s_end, &cslot);
    uint64_t* bar = &S.ck_full[i & 1];
    tma::mbar_expect_tx(bar, kCkBytes);
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(tma::smem_u32(&S.ops[i & 1].ck[0][0][0])), "l"(reinterpret_cast<uint64_t>(&d.m_ck)), "r"(tma::smem_u32(bar)),
          "r"(0), "r"(cslot), "r"(c_base / 2), "r"(0), "r"(b)
        : "memory");
  };
  auto produce = [&](int slot, int sb0, int s_end) {
    const Raw<T>& R = S.raw[slot];
    Ops& O = S.ops[slot];
#pragma unroll 2
    for (int e = 0; e < kUnits; ++e) {
      const int k = k0 + kStepInc * e;
      const int s = sb0 + k;
      const bool valid = s < s_end;
      const int r = rev ? kTB - 1 - k : k;
      const float2 u2 = SmemPair<T>::ld(&R.u[r][2 * cp]);
      const float2 x = fadd2(SmemPair<T>::ld(&R.dl[r][2 * cp]), bias);
      float2 dt = x, sig = make_float2(1.f, 1.f);
      if (SOFTPLUS) {
        if (PRECISE) {
          dt = make_float2(softplus_fwd<true>(x.x), softplus_fwd<true>(x.y));
          sig = make_float2(softplus_grad(x.x), softplus_grad(x.y));
        } else {
          // one exponential serves softplus and its derivative: e = exp(x), softplus = ln(1 + e), sigmoid = e / (1 + e)
          const float ex = ex2(x.x * kLog2e), ey = ex2(x.y * kLog2e);
          const float px = 1.0f + ex, py = 1.0f + ey;
          dt.x = x.x > 20.0f ? x.x : kLn2 * lg2(px);
          dt.y = x.y > 20.0f ? x.y : kLn2 * lg2(py);
          sig.x = x.x > 20.0f ? 1.0f : ex * rcp(px);
          sig.y = x.y > 20.0f ? 1.0f : ey * rcp(py);
        }
      }
      const float2 go = fmul2(SmemPair<T>::ld(&R.go[r][2 * cp]), make_float2(scale, scale));
      float2 dy = go;
      if (HAS_Z) {
        const float2 zz = SmemPair<T>::ld(&R.z[r][2 * cp]);
        const float2 sz = make_float2(sigmoid_sel<PRECISE>(zz.x), sigmoid_sel<PRECISE>(zz.y));
        const float2 gs = fmul2(go, sz);
        dy = fmul2(gs, zz);
        if (do_dz && valid && ch_ok) {
          const float2 pre = SmemPair<T>::ld(&R.pre[r][2 * cp]);
          const int tt = rev ? L - 1 - s : s;
          const float2 dsilu = make_float2(fmaf(zz.x, 1.f - sz.x, 1.f), fmaf(zz.y, 1.f - sz.y, 1.f));
          P2::st(pdz + (int64_t)tt * P.dz_sl, fmul2(fmul2(gs, pre), dsilu));
        }
      }
      if (!valid) { dt = make_float2(0.f, 0.f); dy = make_float2(0.f, 0.f); }   // identity step
      const float2 du = fmul2(dt, u2);
      O.dd[k][cp] = make_float4(dt.x, dt.y, du.x, du.y);
      O.us[k][cp] = make_float4(u2.x, u2.y, sig.x, sig.y);
      O.dy[k][cp] = dy;
    }
    if (io < 64) {
      const int r = rev ? kTB - 1 - bc_row : bc_row;
      *reinterpret_cast<float4*>(&O.bc[bc_row][bc_col]) = SmemQuad<T>::ld(&R.bc[r][bc_col]);
    }
  };
  auto rows_out = [&](int i, int sb0, int s_end) {
    const int slot = i & 1;
    mbar_wait_long(&S.out_full[slot], (i >> 1) & 1);
    float4 acc = *reinterpret_cast<const float4*>(&S.bcw[slot][0][pk][4 * part]);
#pragma unroll
    for (int w = 1; w < kRW; ++w) {
      const float4 v = *reinterpret_cast<const float4*>(&S.bcw[slot][w][pk][4 * part]);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    warp_arrive(&S.out_empty[slot], lane);
    const int s = sb0 + pk;
    if (s < s_end) *reinterpret_cast<float4*>(ppart + (int64_t)s * d.part_ss) = acc;
  };

  if (io == 0) {
    for (int i = 0; i < kRS && i < ntot; ++i) issue_raw(i);
    for (int i = 0; i < 2 && i < ntot; ++i) issue_ck(i);
  }
  int psb0 = 0, ps_end = 0;
#pragma unroll 1
  for (int i = 0; i < ntot; ++i) {
    const int slot = i & 1;
    const uint32_t par = (i >> 1) & 1;
    int sb0, s_end, cslot;
    seq.get(i, &sb0, &s_end, &cslot);
    if (i >= 2) {
      mbar_wait_long(&S.in_empty[slot], par ^ 1);           // the recurrence warps are done with ops[slot] (tile i - 2)
      if (io == 0) {
        tma::fence_proxy_async_smem();
        issue_ck(i);
      }
    }
    mbar_wait(&S.raw_full[slot], par);
    produce(slot, sb0, s_end);
    warp_arrive(&S.in_full[slot], lane);
    asm volatile("bar.sync 1, %0;" ::"n"(kIT) : "memory");  // every IO thread has read raw stage `slot`
    if (io == 0 && i + kRS < ntot) {
      tma::fence_proxy_async_smem();
      issue_raw(i + kRS);
    }
    if (i > 0 && io < 64) rows_out(i - 1, psb0, ps_end);
    psb0 = sb0; ps_end = s_end;
  }
  if (ntot > 0 && io < 64) rows_out(ntot - 1, psb0, ps_end);
}

template <typename T, bool SOFTPLUS, bool HAS_Z>
__global__ void __launch_bounds__(kRT + kIW * 32, 3) scan_bwd_wg_kernel(const __grid_constant__ BwdParams P) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Smem<T>& S = *reinterpret_cast<Smem<T>*>(smem_raw);
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int i = 0; i < kRS; ++i) tma::mbar_init(&S.raw_full[i], 1);
    for (int i = 0; i < 2; ++i) {
      tma::mbar_init(&S.ck_full[i], 1);
      tma::mbar_init(&S.in_full[i], kIA);
      tma::mbar_init(&S.in_empty[i], kRW);
      tma::mbar_init(&S.out_full[i], kRW);
      tma::mbar_init(&S.out_empty[i], 2);
    }
    tma::fence_barrier_init();
  }
  __syncthreads();
  if (tid < kRT) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(CM_BWDWG_RREG));
    scan_role<T>(P, S, tid, blockIdx.z);
  } else {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(CM_BWDWG_IREG));
    if (tid - kRT < kIT) io_role<T, SOFTPLUS, HAS_Z>(P, S, tid - kRT, blockIdx.z);
  }
}

// ---- host --------------------------------------------------------------------------------------------------------------
static bool step_stride32(int64_t sl_elems, int es, bool reverse, int32_t* out) {
  const int64_t v = (reverse ? -sl_elems : sl_elems) * es;
  if (v > INT32_MAX / 2 || v < INT32_MIN / 2) return false;
  *out = (int32_t)v;
  return true;
}

// 5-D map over the fp32 checkpoints [batch][dim][nck][16]: (state, slot, channel pair, channel parity, batch)
static bool make_map_ckpt(CUtensorMap* m, const float* base, int64_t dim, int64_t nck, int64_t batch, int64_t sd, int64_t sb) {
  tma::EncodeTiledFn fn = tma::encode_fn();
  if (fn == nullptr) return false;
  cuuint64_t dims[5] = {16, (cuuint64_t)nck, (cuuint64_t)(dim / 2), 2, (cuuint64_t)batch};
  const uint64_t sbb = batch > 1 ? (uint64_t)sb * 4 : (uint64_t)dim * (uint64_t)sd * 4;
  cuuint64_t strides[4] = {64, (cuuint64_t)sd * 8, (cuuint64_t)sd * 4, sbb};
  cuuint32_t box[5] = {16, 1, (cuuint32_t)kNP, 2, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  for (int i = 0; i < 4; ++i)
    if ((strides[i] & 15) != 0 || strides[i] >= (1ull << 40) || strides[i] == 0) return false;
  const CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

template <typename T>
static bool build_params(const cm_scan_bwd_args& a, BwdParams* P, bool maps) {
  constexpr int ES = (int)sizeof(T);
  if (a.dstate != 16 || a.dim % 32 != 0 || a.dim < kCH || a.seqlen < 1) return false;
  if (a.lanes_per_channel != 0 && a.lanes_per_channel != 1) return false;
  if (a.dout.ptr == nullptr || a.dout.sd != 1) return false;
  const bool has_z = a.z.ptr != nullptr;
  if (has_z && (a.z.sd != 1 || a.out_pre.ptr == nullptr || a.out_pre.sd != 1 || a.dz.ptr == nullptr || a.dz.sd != 1)) return false;
  const int64_t L = a.seqlen, Bt = a.batch, D = a.dim;
  P->L = a.seqlen; P->ndir = a.ndir; P->n_slab = (a.dim + kCH - 1) / kCH; P->dim = a.dim; P->has_z = has_z;
  P->flags = a.flags; P->scale = a.out_scale;
  P->dz = nullptr; P->dz_sb = 0; P->dz_sl = 0;
  if (has_z) {
    if ((reinterpret_cast<uintptr_t>(a.dz.ptr) % (2 * ES)) != 0 || a.dz.sb % 2 != 0 || a.dz.sl % 2 != 0) return false;
    P->dz = static_cast<char*>(a.dz.ptr);
    P->dz_sb = a.dz.sb * ES;
    if (!step_stride32(a.dz.sl, ES, false, &P->dz_sl)) return false;
  }
  auto ok_map = [&](CUtensorMap* m, const cm_tensor3& t, int64_t ch, int box_c) {
    if (t.ptr == nullptr || t.sd != 1) return false;
    // TMA addressability is a property of the view (alignment, strides): checked identically with or without encoding
    const uint64_t sl = (uint64_t)t.sl * ES, sb = (uint64_t)t.sb * ES;
    if ((reinterpret_cast<uintptr_t>(t.ptr) & 15) != 0 || (sl & 15) != 0 || t.sl <= 0) return false;
    if (Bt > 1 && ((sb & 15) != 0 || t.sb <= 0)) return false;
    if (Bt == 1 && ((sl * (uint64_t)L) & 15) != 0) return false;
    if (!maps) return true;
    return tma::make_map_blc(m, t.ptr, ES, ch, L, Bt, t.sl, t.sb, box_c, kTB);
  };
  if (!ok_map(&P->m_go, a.dout, D, kCH)) return false;
  if (has_z && (!ok_map(&P->m_z, a.z, D, kCH) || !ok_map(&P->m_pre, a.out_pre, D, kCH))) return false;
  const int64_t nck = cm_scan_num_ckpt(a.seqlen, a.ndir);
  for (int r = 0; r < a.ndir; ++r) {
    const cm_scan_bwd_dir& sd = a.dir[r];
    const cm_scan_dir& s = sd.in;
    BwdDir& d = P->dir[r];
    if (s.bc_const) return false;
    if (s.Bm.ptr == nullptr || s.Cm.ptr == nullptr || s.Bm.sd != 1 || s.Cm.sd != 1) return false;
    if (static_cast<const char*>(s.Cm.ptr) != static_cast<const char*>(s.Bm.ptr) + 16 * ES || s.Cm.sl != s.Bm.sl || s.Cm.sb != s.Bm.sb)
      return false;
    auto pair_ok = [&](const cm_tensor3& t) {
      return t.ptr != nullptr && t.sd == 1 && (reinterpret_cast<uintptr_t>(t.ptr) % (2 * ES)) == 0 && t.sb % 2 == 0 && t.sl % 2 == 0;
    };
    if (!pair_ok(sd.du) || !pair_ok(sd.ddelta)) return false;
    if (s.ckpt == nullptr || (reinterpret_cast<uintptr_t>(s.ckpt) & 15) != 0 || (s.ckpt_sb % 4) != 0 || (s.ckpt_sd % 4) != 0 ||
        s.ckpt_sd <= 0)
      return false;
    if ((reinterpret_cast<uintptr_t>(sd.dBC_part) & 15) != 0 || (reinterpret_cast<uintptr_t>(sd.dA_part) & 15) != 0) return false;
    if (!ok_map(&d.m_u, s.u, D, kCH) || !ok_map(&d.m_dl, s.delta, D, kCH) || !ok_map(&d.m_bc, s.Bm, 32, 32)) return false;
    if (maps && !make_map_ckpt(&d.m_ck, s.ckpt, D, nck, Bt, s.ckpt_sd, s.ckpt_sb)) return false;
    const bool rev = s.reverse != 0;
    const int64_t l0 = rev ? L - 1 : 0;
    d.reverse = rev; d.write_dz = (r == 0); d.pad0 = d.pad1 = 0;
    d.du = static_cast<char*>(sd.du.ptr) + l0 * sd.du.sl * ES;
    d.ddl = static_cast<char*>(sd.ddelta.ptr) + l0 * sd.ddelta.sl * ES;
    d.du_sb = sd.du.sb * ES; d.ddl_sb = sd.ddelta.sb * ES;
    if (!step_stride32(sd.du.sl, ES, rev, &d.du_ss) || !step_stride32(sd.ddelta.sl, ES, rev, &d.ddl_ss)) return false;
    d.s1 = cm_first_range(a.seqlen, a.ndir, s.reverse);
    d.A = s.A; d.A_sd = s.A_sd; d.A_sn = s.A_sn;
    d.Dskip = s.Dskip; d.bias = s.delta_bias;
    d.dBC_part = sd.dBC_part; d.dA_part = sd.dA_part; d.dD_part = sd.dD_part; d.dbias_part = sd.dbias_part;
    d.part_l0 = l0 * 32;
    d.part_ss = rev ? -32 : 32;
  }
  return true;
}

template <typename T>
static int try_t(const cm_scan_bwd_args& a, cudaStream_t st, int* rc) {
  BwdParams P;
  if (!build_params<T>(a, &P, true)) return 0;
  const size_t smem = sizeof(Smem<T>);
  const bool sp = (a.flags & CM_FLAG_DELTA_SOFTPLUS) != 0, hz = a.z.ptr != nullptr;
  auto launch = [&](auto kern) -> int {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);   // per device
    if (e != cudaSuccess) return (int)e;
    kern<<<dim3(P.n_slab, a.batch, a.ndir), kRT + kIW * 32, smem, st>>>(P);
    e = cudaGetLastError();
    return e == cudaSuccess ? 0 : (int)e;
  };
  if (sp) *rc = hz ? launch(scan_bwd_wg_kernel<T, true, true>) : launch(scan_bwd_wg_kernel<T, true, false>);
  else *rc = hz ? launch(scan_bwd_wg_kernel<T, false, true>) : launch(scan_bwd_wg_kernel<T, false, false>);
  return 1;
}

}  // namespace wgb

// 1 if cm_scan_bwd will take the warpgroup kernel for these arguments (a pure function of the argument block: the caller
// sizes the dB/dC partial tensor from it through cm_scan_bwd_slab_channels)
int scan_bwd_warpgroup_applies(const cm_scan_bwd_args& a) {
  if (getenv("CM_SCAN_NO_WG") != nullptr || getenv("CM_SCAN_NO_SP") != nullptr || getenv("CM_SCAN_GENERIC") != nullptr) return 0;
  // Below one full wave of 3 CTAs per SM the 32-channel CTAs of scan_bwd_sp.cu spread the rows over more SMs (measured at
  // the ConMamba-small shape, 320 CTAs: 0.198 ms against 0.146 ms); CM_SCAN_WG=1 forces this kernel for measurements.
  const char* force = getenv("CM_SCAN_WG");
  const int64_t ctas = (int64_t)((a.dim + wgb::kCH - 1) / wgb::kCH) * a.batch * a.ndir;
  if ((force == nullptr || force[0] == '0') && ctas < 3 * 148) return 0;
  wgb::BwdParams P;
  switch (a.dtype) {
    case CM_F32: return wgb::build_params<float>(a, &P, false) ? 1 : 0;
    case CM_BF16: return wgb::build_params<__nv_bfloat16>(a, &P, false) ? 1 : 0;
    default: return wgb::build_params<__half>(a, &P, false) ? 1 : 0;
  }
}
int scan_bwd_warpgroup_slab() { return wgb::kCH; }

int scan_bwd_try_warpgroup(const cm_scan_bwd_args& a, cudaStream_t st, int* rc) {
  switch (a.dtype) {
    case CM_F32: return wgb::try_t<float>(a, st, rc);
    case CM_BF16: return wgb::try_t<__nv_bfloat16>(a, st, rc);
    default: return wgb::try_t<__half>(a, st, rc);
  }
}

}  // namespace cm
