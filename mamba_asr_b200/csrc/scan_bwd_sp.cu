// Selective-scan backward, warp-specialised state-parallel channel-last kernel for sm_100a ("sp" path).
//
// Same mathematics, checkpoint and partial-sum contracts as scan_bwd.cu (see its header; adjoint of
// modules/mamba/selective_scan_interface.py:106-157), same lane mapping and role split as
// scan_fwd_sp.cu: a lane owns states 4m..4m+3 of two adjacent channels, a CTA covers 32 channels of ONE direction
// (the two directions are independent in backward: grid.z) with
//   2 recurrence warps - per 8-step tile (= the checkpoint interval, walked last tile first):
//        forward  : reload the checkpoint, recompute h_k for the tile, parking them in a warp-private shared history;
//        reverse  : lambda = dy*C + mu ; mu <- a*lambda ; a*h_{k-1} is obtained as h_k - du*B (never divides by a decay);
//                   per step the lane emits  dB[4], dC[4]  (summed over its 2 channels)  and  r1 = sum_j lambda_j B_j,
//                   r2 = sum_j lambda_j (a h_{k-1})_j kA_j  (summed over its 4 states);  dA accumulates in registers;
//        reduce   : warp-local sums over the 4 lanes of a pair (r1, r2) and over the warp's 8 pairs (dB, dC);
//   2 IO warps - raw loads one tile ahead into registers; softplus, sigmoid, gate and dz; fp32 operand rows into a
//        2-tile shared ring;  outputs du = dy*D + dt*r1, ddelta = (ln2*r2 + u*r1)*sigmoid(delta+bias), the per-channel
//        dD / d(delta_bias) sums, and the 32-channel slab row of the dB/dC partial tensor (128-byte coalesced rows).
// mbarrier full/empty pairs guard the operand ring and the result ring; nothing else synchronises the roles.
// Deterministic: fixed-order sums, no atomics (the reference kernel accumulates dB/dC with fp32 atomics).
//
// Requirements (else cm_scan_bwd falls through to scan_bwd_cl.cu / scan_bwd.cu): as scan_fwd_sp.cu, plus
// lanes_per_channel in {0, 1} (the dB/dC partial tensor then has 32-channel slabs).
#include <climits>
#include <cstdlib>
#include <type_traits>

#include "sp_common.cuh"

namespace cm {
namespace spb {

using namespace cm::sp;

#ifndef CM_BWDSP_MINB
#define CM_BWDSP_MINB 4
#endif
#ifndef CM_BWDSP_UNROLL
#define CM_BWDSP_UNROLL 2
#endif
constexpr int kUnr = CM_BWDSP_UNROLL;
#ifndef CM_BWDSP_HREG
#define CM_BWDSP_HREG 4
#endif
constexpr int kHR = CM_BWDSP_HREG;        // last kHR steps of a tile keep their recomputed states in registers, the rest in smem
constexpr int kTB = CM_SCAN_CKPT_STEPS;   // steps per tile (8)
constexpr int kIO = 64;                   // IO threads (2 warps)
constexpr int kIU = kTB * kNP / kIO;      // (step, pair) units per IO thread and tile (2)
constexpr int kIKS = kIO / kNP;           // step distance between an IO thread's units (4)
constexpr float kLn2f = 0.6931471805599453f;

struct BwdDir {
  const char *u, *dl, *B, *C, *dout, *z, *pre;   // byte pointers at (batch 0, channel 0, PROCESSED step 0)
  char *du, *ddl, *dz;
  int64_t u_sb, dl_sb, b_sb, c_sb, dout_sb, z_sb, pre_sb, du_sb, ddl_sb, dz_sb;   // batch strides (bytes)
  int32_t u_ss, dl_ss, b_ss, c_ss, dout_ss, z_ss, pre_ss, du_ss, ddl_ss, dz_ss;   // bytes per processed step
  int32_t s1;            // length of the first range
  int32_t write_dz;
  const float* A;
  int64_t A_sd, A_sn;
  const float *Dskip, *bias;
  const float* ckpt;
  int64_t ckpt_sb, ckpt_sd;
  float *dBC_part, *dA_part, *dD_part, *dbias_part;
  int64_t part_l0;       // float offset of processed step 0 inside one [L][32] slab
  int32_t part_ss;       // floats per processed step (+-32)
  int32_t pad;
};
struct BwdParams {
  int32_t L, ndir, n_slab;
  uint32_t flags;
  float scale;
  int32_t pad;
  BwdDir dir[2];
};

struct BwdSmem {
  float4 dd[2][kTB][kNP];        // (dt0, dt1, du0, du1)                              [operand ring, 2 tiles]
  float2 dy[2][kTB][kNP];        // gated output gradient of the pair
  float bc[2][kTB][32];          // B[0..15] | C[0..15]
  float4 hs[kNW][kTB - kHR > 0 ? kTB - kHR : 1][2][32];   // recomputed states of the first steps of the tile in flight
  float4 pb[kNW][kTB][2][8][4];  // per-lane dB[4] (0) and dC[4] (1)                 (warp-private)
  float4 pr[kNW][kTB][8][4];     // per-lane (r1_0, r1_1, r2_0, r2_1)                (warp-private)
  float4 r[2][kTB][kNP];         // sums over states                                 [result ring, 2 tiles]
  float bcw[2][kNW][kTB][32];    // per-warp sums over its 16 channels: dB[16] | dC[16]
  float red[kIO][4];             // final dD / dbias reduction
  uint64_t in_full[2], in_empty[2], out_full[2], out_empty[2];
};

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
__device__ __forceinline__ void warp_arrive(uint64_t* b, int lane) {
  __syncwarp();
  if (lane == 0) mbar_arrive(b);
}

// Tiles are visited in reverse processing order: range 1 (if bidirectional) last tile first, then range 0.
struct TileSeq {
  int n1, n0, s1, L, j1;   // tiles of range 1 / range 0, split point, length, first checkpoint slot of range 1
  __device__ __forceinline__ TileSeq(int L_, int ndir, int s1_) {
    L = L_;
    s1 = ndir == 2 ? s1_ : L_;
    n0 = cm_ceil_div(s1, kTB);
    n1 = ndir == 2 ? cm_ceil_div(L - s1, kTB) : 0;
    j1 = n0;   // = ceil(s1 / CM_SCAN_CKPT_STEPS)
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

// ---- recurrence warps ------------------------------------------------------------------------------------------------
#ifndef CM_BWDSP_RTDIR
#define CM_BWDSP_RTDIR 1
#endif
// The direction is a run-time value here (the role touches five fields of BwdDir, all before the time loop): as a template
// parameter the kernel carried two copies of the 8-step-unrolled recompute + reverse sweep.
__device__ __forceinline__ void scan_role(const BwdParams& P, BwdSmem& S, const int gt, const int dir) {
  const BwdDir& d = P.dir[dir];
  const int warp = gt >> 5, lane = gt & 31;
  const int b = blockIdx.y;
  const int c_base = blockIdx.x * kCH;
  const int g = lane >> 2, m = lane & 3;
  const int pr = warp * 8 + g;
  const int c0 = c_base + 2 * pr;
  float2 kA[4], mu[4], dA[4];   // [state 4m + j] ; .x channel c0, .y channel c0 + 1
  {
    const float* A0 = d.A + (int64_t)c0 * d.A_sd;
    const float* A1 = A0 + d.A_sd;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      kA[j] = make_float2(__ldg(A0 + (4 * m + j) * d.A_sn) * kLog2e, __ldg(A1 + (4 * m + j) * d.A_sn) * kLog2e);
      mu[j] = make_float2(0.f, 0.f);
      dA[j] = make_float2(0.f, 0.f);
    }
  }
  const float* ckp = d.ckpt + b * d.ckpt_sb + (int64_t)c0 * d.ckpt_sd + 4 * m;
  const TileSeq seq(P.L, P.ndir, d.s1);
  const int ntot = seq.total();

  int sb0, s_end, cslot;
  float4 ck0, ck1;
  if (ntot > 0) {
    seq.get(0, &sb0, &s_end, &cslot);
    ck0 = __ldg(reinterpret_cast<const float4*>(ckp + (int64_t)cslot * 16));
    ck1 = __ldg(reinterpret_cast<const float4*>(ckp + (int64_t)cslot * 16 + d.ckpt_sd));
  }
#pragma unroll 1
  for (int i = 0; i < ntot; ++i) {
    const int slot = i & 1;
    const uint32_t par = (i >> 1) & 1;
    float2 h[4] = {make_float2(ck0.x, ck1.x), make_float2(ck0.y, ck1.y), make_float2(ck0.z, ck1.z), make_float2(ck0.w, ck1.w)};
    if (i + 1 < ntot) {   // next tile's checkpoint: in flight during this tile
      seq.get(i + 1, &sb0, &s_end, &cslot);
      ck0 = __ldg(reinterpret_cast<const float4*>(ckp + (int64_t)cslot * 16));
      ck1 = __ldg(reinterpret_cast<const float4*>(ckp + (int64_t)cslot * 16 + d.ckpt_sd));
    }
    mbar_wait(&S.in_full[slot], par);
    const float4* ddb = &S.dd[slot][0][pr];
    const float2* dyb = &S.dy[slot][0][pr];
    const float* bcb = &S.bc[slot][0][4 * m];
    // ---- forward: recompute the states of the tile.  The history of the last kHR steps stays in registers, the first
    // steps are parked in shared memory (all 8 in shared memory cost 16 of the 86 LSU wavefronts per warp-step that bound
    // the first version of this kernel; all 8 in registers spill)
    float2 hist[kHR > 0 ? kHR : 1][4];
#ifdef CM_ABL_NO_FWD
#pragma unroll
    for (int k = 0; k < kHR; ++k)
#pragma unroll
      for (int j = 0; j < 4; ++j) hist[k][j] = h[j];
    if (mu[1].x == 123.456f)
#endif
#pragma unroll
    for (int k = 0; k < kTB; ++k) {
      const float4 dd = ddb[k * kNP];
      const float4 bb = *reinterpret_cast<const float4*>(bcb + k * 32);
      const float2 dt = make_float2(dd.x, dd.y), du = make_float2(dd.z, dd.w);
      const float b4[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 x = fmul2(dt, kA[j]);
        const float2 a = make_float2(ex2(x.x), ex2(x.y));
        h[j] = ffma2(a, h[j], fmul2(du, make_float2(b4[j], b4[j])));
      }
      if (k >= kTB - kHR) {
#pragma unroll
        for (int j = 0; j < 4; ++j) hist[k - (kTB - kHR)][j] = h[j];
      } else {
        S.hs[warp][k][0][lane] = make_float4(h[0].x, h[1].x, h[2].x, h[3].x);
        S.hs[warp][k][1][lane] = make_float4(h[0].y, h[1].y, h[2].y, h[3].y);
      }
    }
    // ---- reverse sweep
#pragma unroll
    for (int k = kTB - 1; k >= 0; --k) {
      if ((k % kUnr) == kUnr - 1) asm volatile("" ::: "memory");   // scheduling fence: bounds how far ptxas hoists the
                                                                  // operand loads of later steps (register pressure)
      const float4 dd = ddb[k * kNP];
      const float2 dy = dyb[k * kNP];
      const float4 bb = *reinterpret_cast<const float4*>(bcb + k * 32);
      const float4 cc = *reinterpret_cast<const float4*>(bcb + k * 32 + 16);
      const float2 dt = make_float2(dd.x, dd.y), du = make_float2(dd.z, dd.w);
      const float b4[4] = {bb.x, bb.y, bb.z, bb.w}, c4[4] = {cc.x, cc.y, cc.z, cc.w};
      float2 hk[4];
      if (k >= kTB - kHR) {
#pragma unroll
        for (int j = 0; j < 4; ++j) hk[j] = hist[k - (kTB - kHR)][j];
      } else {
        const float4 hx = S.hs[warp][k][0][lane], hy = S.hs[warp][k][1][lane];
        hk[0] = make_float2(hx.x, hy.x); hk[1] = make_float2(hx.y, hy.y);
        hk[2] = make_float2(hx.z, hy.z); hk[3] = make_float2(hx.w, hy.w);
      }
      float dBv[4], dCv[4];
      float2 r1 = make_float2(0.f, 0.f), r2 = make_float2(0.f, 0.f);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 x = fmul2(dt, kA[j]);
        const float2 a = make_float2(ex2(x.x), ex2(x.y));
        const float2 lam = ffma2(dy, make_float2(c4[j], c4[j]), mu[j]);
        const float2 pc = fmul2(dy, hk[j]);
        dCv[j] = pc.x + pc.y;
        const float2 pbv = fmul2(lam, du);
        dBv[j] = pbv.x + pbv.y;
        const float2 t = ffma2(du, make_float2(-b4[j], -b4[j]), hk[j]);   // a * h_{k-1}
        const float2 q = fmul2(lam, t);
        dA[j] = ffma2(q, dt, dA[j]);
        r1 = ffma2(lam, make_float2(b4[j], b4[j]), r1);
        r2 = ffma2(q, kA[j], r2);
        mu[j] = fmul2(a, lam);
      }
#ifdef CM_ABL_NO_PB
      if (dBv[0] == 123.456f)
#endif
      S.pb[warp][k][0][g][m] = make_float4(dBv[0], dBv[1], dBv[2], dBv[3]);
#ifdef CM_ABL_NO_PB
      if (dCv[0] == 123.456f)
#endif
      S.pb[warp][k][1][g][m] = make_float4(dCv[0], dCv[1], dCv[2], dCv[3]);
#ifdef CM_ABL_NO_PB
      if (r1.x == 123.456f)
#endif
      S.pr[warp][k][g][m] = make_float4(r1.x, r1.y, r2.x, r2.y);
    }
    __syncwarp();
    if (i >= 2) mbar_wait(&S.out_empty[slot], par ^ 1);
#ifdef CM_ABL_NO_PB
    if (mu[0].x == 123.456f) {
#else
    {
#endif
    // ---- warp-local reductions
    // (step, pair) -> sum over the pair's 4 lanes of (r1, r2): 64 units, 2 per lane
#pragma unroll
    for (int jj = 0; jj < 2; ++jj) {
      const int ui = lane + 32 * jj;          // = k * 8 + g
      const float4* src = &S.pr[warp][0][0][0] + ui * 4;
      const int rot = ui >> 1;                // lanes 64 B apart: a rotated start keeps the 8 lanes of a phase on 8 bank groups
      const float4 v0 = src[rot & 3], v1 = src[(rot + 1) & 3], v2 = src[(rot + 2) & 3], v3 = src[(rot + 3) & 3];
      const float2 s0 = fadd2(fadd2(make_float2(v0.x, v0.y), make_float2(v1.x, v1.y)),
                              fadd2(make_float2(v2.x, v2.y), make_float2(v3.x, v3.y)));
      const float2 s1 = fadd2(fadd2(make_float2(v0.z, v0.w), make_float2(v1.z, v1.w)),
                              fadd2(make_float2(v2.z, v2.w), make_float2(v3.z, v3.w)));
      S.r[slot][ui >> 3][warp * 8 + (ui & 7)] = make_float4(s0.x, s0.y, s1.x, s1.y);
    }
    // (step, lane-in-pair, dB|dC) -> sum over the warp's 8 pairs: 64 float4 outputs, 2 per lane
#pragma unroll
    for (int jj = 0; jj < 2; ++jj) {
      const int oi = lane + 32 * jj;          // = k * 8 + mm * 2 + which
      const int k = oi >> 3, mm = (oi >> 1) & 3, which = oi & 1;
      // the dB and dC lanes of a phase start on different halves of the 8 pairs (bank groups 4*gg + mm)
      float2 a0 = make_float2(0.f, 0.f), a1 = make_float2(0.f, 0.f);
#pragma unroll
      for (int gg = 0; gg < 8; ++gg) {
        const float4 v = S.pb[warp][k][which][gg ^ which][mm];
        a0 = fadd2(a0, make_float2(v.x, v.y));
        a1 = fadd2(a1, make_float2(v.z, v.w));
      }
      *reinterpret_cast<float4*>(&S.bcw[slot][warp][k][which * 16 + 4 * mm]) = make_float4(a0.x, a0.y, a1.x, a1.y);
    }
    }
    __syncwarp();
    if (lane == 0) {
      mbar_arrive(&S.in_empty[slot]);
      mbar_arrive(&S.out_full[slot]);
    }
  }
  // dA: per-row sums over time
  float* da = d.dA_part + ((int64_t)b * gridDim.x * kCH + c0) * 16 + 4 * m;
  *reinterpret_cast<float4*>(da) = make_float4(dA[0].x, dA[1].x, dA[2].x, dA[3].x);
  *reinterpret_cast<float4*>(da + 16) = make_float4(dA[0].y, dA[1].y, dA[2].y, dA[3].y);
}

// ---- IO warps ----------------------------------------------------------------------------------------------------------
// The IO role reads its BwdDir fields inside the loops (constant-bank operands); with a run-time direction they move to
// registers / indexed constant loads and the kernel is 3 % slower (measured), so this role stays templated in effect.
#ifndef CM_BWDSP_RTDIR_IO
#define CM_BWDSP_RTDIR_IO 0
#endif
template <typename T>
__device__ __forceinline__ void io_role(const BwdParams& P, BwdSmem& S, const int io, const int DIR) {
  using P2 = Pair<T>;
  using Q4 = Quad<T>;
  constexpr int ES = (int)sizeof(T);
  constexpr bool PRECISE = sizeof(T) == 4;
  const BwdDir& d = P.dir[DIR];
  const int lane = io & 31;
  const int b = blockIdx.y;
  const int c_base = blockIdx.x * kCH;
  const bool softplus = (P.flags & CM_FLAG_DELTA_SOFTPLUS) != 0;
  const bool has_z = d.z != nullptr;
  const bool do_dz = has_z && d.write_dz != 0;
  const int cp = io & (kNP - 1), k0 = io / kNP;
  const int cu = c_base + 2 * cp;
  float2 bias = make_float2(0.f, 0.f), Dsk = make_float2(0.f, 0.f);
  if (d.bias) bias = make_float2(__ldg(d.bias + cu), __ldg(d.bias + cu + 1));
  if (d.Dskip) Dsk = make_float2(__ldg(d.Dskip + cu), __ldg(d.Dskip + cu + 1));
  const char* pu = d.u + b * d.u_sb + cu * ES;
  const char* pdl = d.dl + b * d.dl_sb + cu * ES;
  const char* pgo = d.dout + b * d.dout_sb + cu * ES;
  const char* pz = has_z ? d.z + b * d.z_sb + cu * ES : nullptr;
  const char* ppre = do_dz ? d.pre + b * d.pre_sb + cu * ES : nullptr;
  char* pdu = d.du + b * d.du_sb + cu * ES;
  char* pddl = d.ddl + b * d.ddl_sb + cu * ES;
  char* pdz = do_dz ? d.dz + b * d.dz_sb + cu * ES : nullptr;
  // B/C quarter rows: one per thread and tile
  const int kq = io >> 3, part = io & 7, isC = part >> 2, q4 = (part & 3) * 4;
  const char* pbc = (isC ? d.C + b * d.c_sb : d.B + b * d.b_sb) + q4 * ES;
  const int bc_ss = isC ? d.c_ss : d.b_ss;
  const int bc_dst = isC * 16 + q4;
  // dB/dC partial rows of this CTA's 32-channel slab: thread -> (step kq, float4 column part)
  float* ppart = d.dBC_part + ((int64_t)b * P.n_slab + blockIdx.x) * (int64_t)P.L * 32 + d.part_l0 + 4 * part;

  const TileSeq seq(P.L, P.ndir, d.s1);
  const int ntot = seq.total();
  typename P2::Raw ru[kIU], rdl[kIU], rgo[kIU], rz[kIU], rpre[kIU];
  typename Q4::Raw rbc;
  struct Keep { float2 u, dt, sig, dy; };
  Keep cur[kIU], prev[kIU];
  float2 dD_acc = make_float2(0.f, 0.f), db_acc = make_float2(0.f, 0.f);

  auto load_raw = [&](int i) {
    int sb0, s_end, cslot;
    seq.get(i, &sb0, &s_end, &cslot);
#pragma unroll
    for (int e = 0; e < kIU; ++e) {
      const int s = sb0 + k0 + e * kIKS;
      const bool v = s < s_end;
      ru[e] = v ? P2::ld_nc(pu + (int64_t)s * d.u_ss) : P2::zero();
      rdl[e] = v ? P2::ld_nc(pdl + (int64_t)s * d.dl_ss) : P2::zero();
      rgo[e] = v ? P2::ld_nc(pgo + (int64_t)s * d.dout_ss) : P2::zero();
      rz[e] = (has_z && v) ? P2::ld_nc(pz + (int64_t)s * d.z_ss) : P2::zero();
      rpre[e] = (do_dz && v) ? P2::ld_nc(ppre + (int64_t)s * d.pre_ss) : P2::zero();
    }
    const int s = sb0 + kq;
    rbc = (s < s_end) ? Q4::ld_nc(pbc + (int64_t)s * bc_ss) : Q4::zero();
  };
  auto produce = [&](int i) {
    int sb0, s_end, cslot;
    seq.get(i, &sb0, &s_end, &cslot);
    const int slot = i & 1;
#pragma unroll
    for (int e = 0; e < kIU; ++e) {
      const int k = k0 + e * kIKS;
      const int s = sb0 + k;
      const bool valid = s < s_end;
      const float2 u2 = P2::cvt(ru[e]);
      const float2 x = fadd2(P2::cvt(rdl[e]), bias);
      float2 dt = x, sig = make_float2(1.f, 1.f);
      if (softplus) {
        dt = make_float2(softplus_fwd<PRECISE>(x.x), softplus_fwd<PRECISE>(x.y));
        sig = make_float2(softplus_grad(x.x), softplus_grad(x.y));
      }
      const float2 go = fmul2(P2::cvt(rgo[e]), make_float2(P.scale, P.scale));
      float2 dy = go;
      if (has_z) {
        const float2 zz = P2::cvt(rz[e]);
        const float2 sz = make_float2(sigmoid_sel<PRECISE>(zz.x), sigmoid_sel<PRECISE>(zz.y));
        dy = make_float2(go.x * zz.x * sz.x, go.y * zz.y * sz.y);
        if (do_dz && valid) {
          const float2 pre = P2::cvt(rpre[e]);
          P2::st(pdz + (int64_t)s * d.dz_ss, make_float2(go.x * pre.x * sz.x * fmaf(zz.x, 1.f - sz.x, 1.f),
                                                         go.y * pre.y * sz.y * fmaf(zz.y, 1.f - sz.y, 1.f)));
        }
      }
      if (!valid) { dt = make_float2(0.f, 0.f); dy = make_float2(0.f, 0.f); }   // identity step
      const float2 du = fmul2(dt, u2);
      S.dd[slot][k][cp] = make_float4(dt.x, dt.y, du.x, du.y);
      S.dy[slot][k][cp] = dy;
      cur[e].u = u2; cur[e].dt = dt; cur[e].sig = sig; cur[e].dy = dy;
    }
    float v[4];
    Q4::cvt(rbc, v);
    *reinterpret_cast<float4*>(&S.bc[slot][kq][bc_dst]) = make_float4(v[0], v[1], v[2], v[3]);
  };
  auto consume = [&](int i) {
    int sb0, s_end, cslot;
    seq.get(i, &sb0, &s_end, &cslot);
    const int slot = i & 1;
    mbar_wait(&S.out_full[slot], (i >> 1) & 1);
    float4 rv[kIU];
#pragma unroll
    for (int e = 0; e < kIU; ++e) rv[e] = S.r[slot][k0 + e * kIKS][cp];
    const float4 w0 = *reinterpret_cast<const float4*>(&S.bcw[slot][0][kq][4 * part]);
    const float4 w1 = *reinterpret_cast<const float4*>(&S.bcw[slot][1][kq][4 * part]);
    warp_arrive(&S.out_empty[slot], lane);
#pragma unroll
    for (int e = 0; e < kIU; ++e) {
      const int s = sb0 + k0 + e * kIKS;
      if (s < s_end) {
        const Keep& kp = prev[e];
        const float2 r1 = make_float2(rv[e].x, rv[e].y), r2 = make_float2(rv[e].z, rv[e].w);
        const float2 duo = ffma2(kp.dt, r1, fmul2(kp.dy, Dsk));
        const float2 ddt = ffma2(r2, make_float2(kLn2f, kLn2f), fmul2(kp.u, r1));
        const float2 ddl = fmul2(ddt, kp.sig);
        dD_acc = ffma2(kp.dy, kp.u, dD_acc);
        db_acc = fadd2(db_acc, ddl);
        P2::st(pdu + (int64_t)s * d.du_ss, duo);
        P2::st(pddl + (int64_t)s * d.ddl_ss, ddl);
      }
    }
    const int s = sb0 + kq;
    if (s < s_end)
      *reinterpret_cast<float4*>(ppart + (int64_t)s * d.part_ss) = make_float4(w0.x + w1.x, w0.y + w1.y, w0.z + w1.z, w0.w + w1.w);
  };

  if (ntot > 0) load_raw(0);
#pragma unroll 1
  for (int i = 0; i < ntot; ++i) {
    const int slot = i & 1;
    if (i >= 2) mbar_wait(&S.in_empty[slot], ((i >> 1) & 1) ^ 1);
    produce(i);
    warp_arrive(&S.in_full[slot], lane);
    if (i + 1 < ntot) load_raw(i + 1);
    if (i > 0) consume(i - 1);
#pragma unroll
    for (int e = 0; e < kIU; ++e) prev[e] = cur[e];
  }
  if (ntot > 0) consume(ntot - 1);

  // per-channel sums over time: 4 threads (k0 = 0..3) share a channel pair
  S.red[io][0] = dD_acc.x; S.red[io][1] = dD_acc.y; S.red[io][2] = db_acc.x; S.red[io][3] = db_acc.y;
  asm volatile("bar.sync 1, %0;" ::"n"(kIO) : "memory");
  if (io < kNP) {
    float4 acc = *reinterpret_cast<const float4*>(S.red[io]);
#pragma unroll
    for (int q = 1; q < kIKS; ++q) {
      const float4 v = *reinterpret_cast<const float4*>(S.red[io + q * kNP]);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    const int64_t row = (int64_t)b * gridDim.x * kCH + c_base + 2 * io;
    if (d.dD_part) { d.dD_part[row] = acc.x; d.dD_part[row + 1] = acc.y; }
    if (d.dbias_part) { d.dbias_part[row] = acc.z; d.dbias_part[row + 1] = acc.w; }
  }
}


template <typename T>
__global__ void __launch_bounds__(kGT + kIO, CM_BWDSP_MINB) scan_bwd_sp_kernel(const __grid_constant__ BwdParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  BwdSmem& S = *reinterpret_cast<BwdSmem*>(smem_raw);
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&S.in_full[i], kIO / 32);
      mbar_init(&S.in_empty[i], kNW);
      mbar_init(&S.out_full[i], kNW);
      mbar_init(&S.out_empty[i], kIO / 32);
    }
  }
  __syncthreads();
  if (tid < kGT) {
#if CM_BWDSP_RTDIR
    scan_role(P, S, tid, blockIdx.z);
#else
    if (blockIdx.z == 0) scan_role(P, S, tid, 0); else scan_role(P, S, tid, 1);
#endif
  } else {
#if CM_BWDSP_RTDIR_IO
    io_role<T>(P, S, tid - kGT, blockIdx.z);
#else
    if (blockIdx.z == 0) io_role<T>(P, S, tid - kGT, 0); else io_role<T>(P, S, tid - kGT, 1);
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
static bool build_params(const cm_scan_bwd_args& a, BwdParams* P) {
  constexpr int ES = (int)sizeof(T);
  if (a.dstate != 16 || a.dim % kCH != 0) return false;
  if (!t_ok<T>(a.dout, 2)) return false;
  if (a.z.ptr != nullptr && (!t_ok<T>(a.z, 2) || !t_ok<T>(a.out_pre, 2) || !t_ok<T>(a.dz, 2))) return false;
  P->L = a.seqlen;
  P->ndir = a.ndir;
  P->n_slab = a.dim / kCH;
  P->flags = a.flags;
  P->scale = a.out_scale;
  P->pad = 0;
  for (int r = 0; r < a.ndir; ++r) {
    const cm_scan_bwd_dir& sd = a.dir[r];
    const cm_scan_dir& s = sd.in;
    BwdDir& d = P->dir[r];
    if (s.bc_const) return false;
    if (!t_ok<T>(s.u, 2) || !t_ok<T>(s.delta, 2) || !t_ok<T>(s.Bm, 4) || !t_ok<T>(s.Cm, 4)) return false;
    if (!t_ok<T>(sd.du, 2) || !t_ok<T>(sd.ddelta, 2)) return false;
    if (s.ckpt == nullptr || (reinterpret_cast<uintptr_t>(s.ckpt) & 15) != 0 || (s.ckpt_sb % 4) != 0 || (s.ckpt_sd % 4) != 0)
      return false;
    if ((reinterpret_cast<uintptr_t>(sd.dBC_part) & 15) != 0 || (reinterpret_cast<uintptr_t>(sd.dA_part) & 15) != 0) return false;
    const bool rev = s.reverse != 0;
    const int64_t l0 = rev ? a.seqlen - 1 : 0;
    auto bp = [&](const cm_tensor3& t) { return static_cast<char*>(t.ptr) + l0 * t.sl * ES; };
    bool ok = true;
    ok &= step_stride(s.u.sl, ES, rev, a.seqlen, &d.u_ss);
    ok &= step_stride(s.delta.sl, ES, rev, a.seqlen, &d.dl_ss);
    ok &= step_stride(s.Bm.sl, ES, rev, a.seqlen, &d.b_ss);
    ok &= step_stride(s.Cm.sl, ES, rev, a.seqlen, &d.c_ss);
    ok &= step_stride(a.dout.sl, ES, rev, a.seqlen, &d.dout_ss);
    ok &= step_stride(sd.du.sl, ES, rev, a.seqlen, &d.du_ss);
    ok &= step_stride(sd.ddelta.sl, ES, rev, a.seqlen, &d.ddl_ss);
    d.z_ss = d.pre_ss = d.dz_ss = 0;
    if (a.z.ptr) {
      ok &= step_stride(a.z.sl, ES, rev, a.seqlen, &d.z_ss);
      ok &= step_stride(a.out_pre.sl, ES, rev, a.seqlen, &d.pre_ss);
      ok &= step_stride(a.dz.sl, ES, rev, a.seqlen, &d.dz_ss);
    }
    if (!ok) return false;
    d.u = bp(s.u); d.dl = bp(s.delta); d.B = bp(s.Bm); d.C = bp(s.Cm); d.dout = bp(a.dout);
    d.z = a.z.ptr ? bp(a.z) : nullptr;
    d.pre = a.z.ptr ? bp(a.out_pre) : nullptr;
    d.dz = a.z.ptr ? bp(a.dz) : nullptr;
    d.du = bp(sd.du); d.ddl = bp(sd.ddelta);
    d.u_sb = s.u.sb * ES; d.dl_sb = s.delta.sb * ES; d.b_sb = s.Bm.sb * ES; d.c_sb = s.Cm.sb * ES;
    d.dout_sb = a.dout.sb * ES; d.z_sb = a.z.sb * ES; d.pre_sb = a.out_pre.sb * ES; d.dz_sb = a.dz.sb * ES;
    d.du_sb = sd.du.sb * ES; d.ddl_sb = sd.ddelta.sb * ES;
    d.s1 = cm_first_range(a.seqlen, a.ndir, s.reverse);
    d.write_dz = (r == 0);
    d.A = s.A; d.A_sd = s.A_sd; d.A_sn = s.A_sn;
    d.Dskip = s.Dskip; d.bias = s.delta_bias;
    d.ckpt = s.ckpt; d.ckpt_sb = s.ckpt_sb; d.ckpt_sd = s.ckpt_sd;
    d.dBC_part = sd.dBC_part; d.dA_part = sd.dA_part; d.dD_part = sd.dD_part; d.dbias_part = sd.dbias_part;
    d.part_l0 = l0 * 32;
    d.part_ss = rev ? -32 : 32;
    d.pad = 0;
  }
  return true;
}

template <typename T>
static int try_t(const cm_scan_bwd_args& a, cudaStream_t st, int* rc) {
  BwdParams P;
  if (!build_params<T>(a, &P)) return 0;
  const size_t smem = sizeof(BwdSmem);
  auto kern = scan_bwd_sp_kernel<T>;
  {   // per-device attribute: set on every launch (see scan_fwd_sp.cu)
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { *rc = (int)e; return 1; }
  }
  kern<<<dim3(a.dim / kCH, a.batch, a.ndir), kGT + kIO, smem, st>>>(P);
  cudaError_t e = cudaGetLastError();
  *rc = (e == cudaSuccess) ? 0 : (int)e;
  return 1;
}

}  // namespace spb

// returns 1 if launched (result in *rc), 0 if the state-parallel path does not apply
int scan_bwd_try_state_parallel(const cm_scan_bwd_args& a, cudaStream_t st, int* rc) {
  switch (a.dtype) {
    case CM_F32: return spb::try_t<float>(a, st, rc);
    case CM_BF16: return spb::try_t<__nv_bfloat16>(a, st, rc);
    default: return spb::try_t<__half>(a, st, rc);
  }
}

}  // namespace cm
