// Selective-scan backward, "lc" kernel (lane = channel, TMA-staged) for sm_100a.  Parity-green but slower than the
// state-parallel kernel at every measured shape, so cm_scan_bwd takes it only with CM_SCAN_LC_BWD=1 (kept: it is the
// measured data point behind the analysis in DESIGN.md section 3.2, and the TMA / slab-negotiation plumbing is shared).
//
// Same mathematics, checkpoint and partial-sum contracts as scan_bwd.cu (see its header; adjoint of
// modules/mamba/selective_scan_interface.py:106-157).  What changes is the mapping:
//
//   * A LANE OWNS ONE CHANNEL and its 16 states (8 packed fp32 pairs).  B and C reach the lanes as broadcast LDS.128 from a
//     per-warp fp32 copy of the tile's B|C rows; dt, dt*u and the gated output gradient dy are per-lane scalars that the
//     packed FMUL2 / FFMA2 broadcast.  The per-channel reductions over the states (r1 = sum_n lambda*B, r2 = sum_n g*A,
//     which give du and d delta) are in-lane sums: no exchange at all.  (The state-parallel kernel of round 1 moved every
//     operand of every state update through the shared-memory pipe - 58 wavefronts per 256 updates - and reduced r1, r2
//     across lanes.)
//   * A CTA is 4 warps = 128 channels of one (batch, direction); tiles of 8 steps (= the checkpoint interval) are walked last
//     tile first.  One thread issues the TMA tile loads (cp.async.bulk.tensor, box 128 channels x 8 steps) of u, delta,
//     dout, z, out_pre and B|C two tiles ahead into a 2-stage ring guarded by mbarriers (transaction bytes).
//   * Per tile and warp: S) per-step scalars (softplus, gate, dz);  F) forward recompute of the 8 states from the
//     checkpoint - the history h_{k-1} of steps 0..3 is parked in shared memory, of steps 4..7 it stays in registers, the
//     decays a_k of steps 5..7 are parked in shared memory (the others are re-evaluated on the MUFU pipe in R: this is the
//     split that fits two CTAs per SM);  R) reverse sweep lambda = dy*C + mu ; mu <- a*lambda ; g = mu*h_{k-1} ;
//     dA += g*dt ; r2 += g*A ; r1 += lambda*B ; per step the lane emits du, d delta and 32 products (dB[16] | dC[16]) that
//     have to be summed over the channels.
//   * The sum over channels: the 32 x 32 (lane x value) block of a step is transposed through a padded shared-memory buffer
//     (8 STS.128 + 8 LDS.128 per lane, conflict-free), quarter-warps are combined with 8 shuffles, and the four warps of
//     the CTA are combined after the tile's single CTA barrier - ONE 128-byte partial row per (batch, 128-channel block,
//     step) goes to global memory (the round-1 kernel wrote one per 32 channels: 4x the partial traffic).
//     Deterministic: fixed-order sums, no atomics (the reference kernel accumulates dB/dC with fp32 atomics).
//
// Requirements (else cm_scan_bwd falls through to the other kernels): as scan_fwd_lc.cu, dim a multiple of 128,
// lanes_per_channel in {0, 1}; the caller sizes the dB/dC partial tensor with cm_scan_bwd_slab_channels().
#include <climits>
#include <cstdlib>

#include "common.cuh"
#include "tma.cuh"

namespace cm {
namespace lcb {

constexpr int kTB = CM_SCAN_CKPT_STEPS;   // steps per tile (8)
constexpr int kW = 4;                     // warps per CTA
constexpr int kCH = 32 * kW;              // channels per CTA (= slab width of the dB/dC partial tensor)
constexpr int kStages = 2;
constexpr int kHS = 4;                    // h_{k-1} of steps 0..kHS-1 in shared memory, of steps kHS..7 in registers
constexpr int kAS = 3;                    // a_k of the last kAS steps in shared memory, the others re-evaluated in R
constexpr int kTrPad = 36;                // floats per lane row of the transposition buffer (conflict-free 16-byte chunks)
constexpr float kLn2f = 0.6931471805599453f;

struct BwdDir {
  CUtensorMap m_u, m_dl, m_bc;
  char *du, *ddl;                          // byte pointers at (batch 0, channel 0, PROCESSED step 0)
  int64_t du_sb, ddl_sb;
  int32_t du_ss, ddl_ss;                   // bytes per processed step (signed)
  int32_t s1, reverse, write_dz, pad0;
  const float* A;
  int64_t A_sd, A_sn;
  const float *Dskip, *bias;
  const float* ckpt;
  int64_t ckpt_sb, ckpt_sd;
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
  int32_t L, ndir, n_blk, dim, has_z;
  uint32_t flags;
  float scale;
};

template <typename T>
struct Stage {                             // one tile of the CTA's per-step operands: [step][channel]
  T u[kTB][kCH], dl[kTB][kCH], go[kTB][kCH], z[kTB][kCH], pre[kTB][kCH];
  T bc[kTB][32];
};
struct WarpBuf {
  float bcf[kTB][32];                      // fp32 B|C rows of the tile, processed order
  float4 hs[kHS][4][32];                   // h_{k-1}, k < kHS: [step][quad of states][lane]
  float4 as[kAS][4][32];                   // a_k, k >= kTB - kAS
  float tr[32][kTrPad];                    // (lane x value) block of one step
};
template <typename T>
struct alignas(128) Smem {
  Stage<T> stage[kStages];
  WarpBuf w[kW];
  float red[2][kW][kTB][32];               // per-warp sums over its 32 channels: dB[16] | dC[16], double-buffered over tiles
  uint64_t full[kStages];
};

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

template <typename T> struct Ld;
template <> struct Ld<float> {
  static __device__ __forceinline__ float at(const char* a) { return *reinterpret_cast<const float*>(a); }
};
template <> struct Ld<__nv_bfloat16> {
  static __device__ __forceinline__ float at(const char* a) {
    return __uint_as_float(static_cast<uint32_t>(*reinterpret_cast<const unsigned short*>(a)) << 16);
  }
};
template <> struct Ld<__half> {
  static __device__ __forceinline__ float at(const char* a) { return __half2float(*reinterpret_cast<const __half*>(a)); }
};

// predicated global store of one element (see scan_fwd_lc.cu)
template <typename T> struct St;
template <> struct St<float> {
  static __device__ __forceinline__ void pred(void* p, float v, bool ok) {
    asm volatile("{ .reg .pred q; setp.ne.b32 q, %2, 0; @q st.global.f32 [%0], %1; }" ::"l"(p), "f"(v), "r"((int)ok) : "memory");
  }
};
template <> struct St<__nv_bfloat16> {
  static __device__ __forceinline__ void pred(void* p, float v, bool ok) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(0.f), "f"(v));
    asm volatile("{ .reg .pred q; setp.ne.b32 q, %2, 0; @q st.global.b16 [%0], %1; }" ::"l"(p), "r"(r), "r"((int)ok) : "memory");
  }
};
template <> struct St<__half> {
  static __device__ __forceinline__ void pred(void* p, float v, bool ok) {
    uint32_t r;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(0.f), "f"(v));
    asm volatile("{ .reg .pred q; setp.ne.b32 q, %2, 0; @q st.global.b16 [%0], %1; }" ::"l"(p), "r"(r), "r"((int)ok) : "memory");
  }
};

__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }
// ex2 that the compiler may not merge with an earlier evaluation of the same argument: the reverse sweep RE-evaluates the
// decays of the first steps instead of keeping them alive in registers across the tile
__device__ __forceinline__ float ex2v(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <typename T, bool SOFTPLUS>
__global__ void __launch_bounds__(kW * 32, 2) scan_bwd_lc_kernel(const __grid_constant__ BwdParams P) {
  constexpr int ES = (int)sizeof(T);
  constexpr int ROWB = kCH * ES;                      // bytes per tile row of a per-channel operand
  constexpr int BCROWB = 32 * ES;
  constexpr bool PRECISE = sizeof(T) == 4;
  constexpr uint32_t kChanTile = kTB * ROWB;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Smem<T>& S = *reinterpret_cast<Smem<T>*>(smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int DIR = blockIdx.z;
  const BwdDir& d = P.dir[DIR];
  const int b = blockIdx.y;
  const int c_blk = blockIdx.x * kCH;
  const int ch = c_blk + warp * 32 + lane;
  const int L = P.L;
  const bool rev = d.reverse != 0;
  const bool has_z = P.has_z != 0;
  const bool do_dz = has_z && d.write_dz != 0;
  WarpBuf& W = S.w[warp];

  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < kStages; ++i) tma::mbar_init(&S.full[i], 1);
    tma::fence_barrier_init();
  }
  __syncthreads();

  float2 kA[8], mu[8], dA[8];
  {
    const float* Ap = d.A + (int64_t)ch * d.A_sd;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      kA[j] = f2(__ldg(Ap + (2 * j) * d.A_sn) * kLog2e, __ldg(Ap + (2 * j + 1) * d.A_sn) * kLog2e);
      mu[j] = f2(0.f, 0.f);
      dA[j] = f2(0.f, 0.f);
    }
  }
  float bias_r = d.bias ? __ldg(d.bias + ch) : 0.f;
  float Dsk_r = d.Dskip ? __ldg(d.Dskip + ch) : 0.f;
  float scale_r = P.scale;
  const float* ckp = d.ckpt + b * d.ckpt_sb + (int64_t)ch * d.ckpt_sd;
  char* pdu = d.du + b * d.du_sb + (int64_t)ch * ES;
  char* pddl = d.ddl + b * d.ddl_sb + (int64_t)ch * ES;
  char* pdz = do_dz ? P.dz + b * P.dz_sb + (int64_t)ch * ES : nullptr;
  int du_ss = d.du_ss, ddl_ss = d.ddl_ss, dz_sl = P.dz_sl;
  int rowstep = rev ? -ROWB : ROWB, bcstep = rev ? -BCROWB : BCROWB;
  asm volatile("" : "+r"(rowstep), "+r"(bcstep), "+r"(du_ss), "+r"(ddl_ss), "+r"(dz_sl), "+f"(bias_r), "+f"(Dsk_r), "+f"(scale_r),
               "+l"(pdu), "+l"(pddl), "+l"(pdz), "+l"(ckp));
  const int first = rev ? (kTB - 1) * ROWB : 0, bcfirst = rev ? (kTB - 1) * BCROWB : 0;
  float dD_acc = 0.f, db_acc = 0.f;

  const TileSeq seq(L, P.ndir, d.s1);
  const int ntot = seq.total();

  auto issue = [&](int i) {                 // thread 0: tile i -> ring slot i & 1
    int sb0, s_end, cslot;
    seq.get(i, &sb0, &s_end, &cslot);
    const int slot = i & 1;
    const int t0 = rev ? (L - sb0 - kTB) : sb0;
    Stage<T>& st = S.stage[slot];
    uint64_t* bar = &S.full[slot];
    tma::mbar_expect_tx(bar, kChanTile * (3u + (has_z ? 1u : 0u) + (do_dz ? 1u : 0u)) + kTB * BCROWB);
    tma::load_3d(&st.u[0][0], &d.m_u, bar, c_blk, t0, b);
    tma::load_3d(&st.dl[0][0], &d.m_dl, bar, c_blk, t0, b);
    tma::load_3d(&st.go[0][0], &P.m_go, bar, c_blk, t0, b);
    if (has_z) tma::load_3d(&st.z[0][0], &P.m_z, bar, c_blk, t0, b);
    if (do_dz) tma::load_3d(&st.pre[0][0], &P.m_pre, bar, c_blk, t0, b);
    tma::load_3d(&st.bc[0][0], &d.m_bc, bar, 0, t0, b);
  };
  if (threadIdx.x == 0) {
    if (ntot > 0) issue(0);
    if (ntot > 1) issue(1);
  }

  float* const bcf = &W.bcf[0][0];
  float* const trw = &W.tr[lane][0];                       // this lane's row of the transposition buffer
  // reader side of the transposition: chunk (lane & 7) of the 8 source lanes of quarter (lane >> 3)
  const float* const trr = &W.tr[(lane >> 3) * 8][(lane & 7) * 4];

#pragma unroll 1
  for (int i = 0; i < ntot; ++i) {
    const int slot = i & 1, pbuf = i & 1;
    int sb0, s_end, cslot;
    seq.get(i, &sb0, &s_end, &cslot);
    const int nvalid = s_end - sb0;            // steps of this tile inside its range (may exceed kTB)
    // the tile's checkpoint: in flight under the scalar phase
    float4 ck[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) ck[q] = __ldg(reinterpret_cast<const float4*>(ckp + (int64_t)cslot * 16) + q);

    tma::mbar_wait(&S.full[slot], (i >> 1) & 1);
    const char* const st0 = reinterpret_cast<const char*>(&S.stage[slot]);
    const char* const col = st0 + first + (warp * 32 + lane) * ES;                 // this lane's column, processed step 0
    {
      // B|C rows to fp32, processed order: lane = element
      const char* bcc = st0 + 5 * kChanTile + bcfirst + lane * ES;
      float v[kTB];
#pragma unroll
      for (int k = 0; k < kTB; ++k) v[k] = Ld<T>::at(bcc + k * bcstep);
#pragma unroll
      for (int k = 0; k < kTB; ++k) bcf[k * 32 + lane] = v[k];
    }
    // ---- S: per-step scalars -------------------------------------------------------------------------------------
    float uu[kTB], dtv[kTB], duv[kTB], dyv[kTB], sgv[kTB];
#pragma unroll
    for (int k = 0; k < kTB; ++k) {
      const char* a = col + k * rowstep;
      const bool valid = k < nvalid;
      const float u = Ld<T>::at(a);
      const float x = Ld<T>::at(a + kChanTile) + bias_r;
      float dt = x, sg = 1.f;
      if (SOFTPLUS) {
        dt = softplus_fwd<PRECISE>(x);
        sg = softplus_grad(x);
      }
      const float go = Ld<T>::at(a + 2 * kChanTile) * scale_r;
      float dy = go;
      if (has_z) {
        const float zz = Ld<T>::at(a + 3 * kChanTile);
        const float sz = sigmoid_sel<PRECISE>(zz);
        dy = go * zz * sz;
        if (do_dz) {
          const float pre = Ld<T>::at(a + 4 * kChanTile);
          const int64_t t = rev ? (int64_t)(L - 1 - (sb0 + k)) : (int64_t)(sb0 + k);
          St<T>::pred(pdz + t * dz_sl, go * pre * sz * fmaf(zz, 1.f - sz, 1.f), valid);
        }
      }
      if (!valid) { dt = 0.f; dy = 0.f; }        // identity step
      uu[k] = u; dtv[k] = dt; duv[k] = dt * u; dyv[k] = dy; sgv[k] = sg;
    }
    __syncwarp();                                // bcf complete

    // ---- F: recompute the states of the tile ---------------------------------------------------------------------
    float2 h[8] = {f2(ck[0].x, ck[0].y), f2(ck[0].z, ck[0].w), f2(ck[1].x, ck[1].y), f2(ck[1].z, ck[1].w),
                   f2(ck[2].x, ck[2].y), f2(ck[2].z, ck[2].w), f2(ck[3].x, ck[3].y), f2(ck[3].z, ck[3].w)};
    float2 hist[kTB - kHS][8];                   // hist[j] = h_{kHS + j - 1}: state BEFORE step kHS + j
#pragma unroll
    for (int k = 0; k < kTB; ++k) {
      if (k < kHS) {
#pragma unroll
        for (int q = 0; q < 4; ++q) W.hs[k][q][lane] = make_float4(h[2 * q].x, h[2 * q].y, h[2 * q + 1].x, h[2 * q + 1].y);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) hist[k - kHS][j] = h[j];
      }
      const float2 dt2 = f2(dtv[k], dtv[k]), du2 = f2(duv[k], duv[k]);
      float4 b4[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) b4[q] = *reinterpret_cast<const float4*>(bcf + k * 32 + q * 4);
      float2 a2[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) a2[j] = fmul2(dt2, kA[j]);
#pragma unroll
      for (int j = 0; j < 8; ++j) a2[j] = f2(ex2(a2[j].x), ex2(a2[j].y));
      if (k >= kTB - kAS) {
#pragma unroll
        for (int q = 0; q < 4; ++q)
          W.as[k - (kTB - kAS)][q][lane] = make_float4(a2[2 * q].x, a2[2 * q].y, a2[2 * q + 1].x, a2[2 * q + 1].y);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        h[2 * q] = ffma2(a2[2 * q], h[2 * q], fmul2(du2, f2(b4[q].x, b4[q].y)));
        h[2 * q + 1] = ffma2(a2[2 * q + 1], h[2 * q + 1], fmul2(du2, f2(b4[q].z, b4[q].w)));
      }
    }

    // ---- R: reverse sweep --------------------------------------------------------------------------------------------
    float2 hk[8];                                // h_k of the current step (for dC)
#pragma unroll
    for (int j = 0; j < 8; ++j) hk[j] = h[j];
#pragma unroll
    for (int k = kTB - 1; k >= 0; --k) {
      const float2 dt2 = f2(dtv[k], dtv[k]), du2 = f2(duv[k], duv[k]), dy2 = f2(dyv[k], dyv[k]);
      float2 a2[8], hp[8];
      if (k >= kTB - kAS) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 v = W.as[k - (kTB - kAS)][q][lane];
          a2[2 * q] = f2(v.x, v.y); a2[2 * q + 1] = f2(v.z, v.w);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) a2[j] = fmul2(dt2, kA[j]);
#pragma unroll
        for (int j = 0; j < 8; ++j) a2[j] = f2(ex2v(a2[j].x), ex2v(a2[j].y));
      }
      if (k < kHS) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 v = W.hs[k][q][lane];
          hp[2 * q] = f2(v.x, v.y); hp[2 * q + 1] = f2(v.z, v.w);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) hp[j] = hist[k - kHS][j];
      }
      float4 b4[4], c4[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) b4[q] = *reinterpret_cast<const float4*>(bcf + k * 32 + q * 4);
#pragma unroll
      for (int q = 0; q < 4; ++q) c4[q] = *reinterpret_cast<const float4*>(bcf + k * 32 + 16 + q * 4);
      float2 r1 = f2(0.f, 0.f), r2 = f2(0.f, 0.f);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float2 bb[2] = {f2(b4[q].x, b4[q].y), f2(b4[q].z, b4[q].w)};
        const float2 cc[2] = {f2(c4[q].x, c4[q].y), f2(c4[q].z, c4[q].w)};
        float2 pB[2], pC[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int j = 2 * q + e;
          const float2 lam = ffma2(dy2, cc[e], mu[j]);
          pC[e] = fmul2(dy2, hk[j]);
          pB[e] = fmul2(lam, du2);
          r1 = ffma2(lam, bb[e], r1);
          mu[j] = fmul2(a2[j], lam);
          const float2 g = fmul2(mu[j], hp[j]);
          r2 = ffma2(g, kA[j], r2);
          dA[j] = ffma2(g, dt2, dA[j]);
        }
        *reinterpret_cast<float4*>(trw + 4 * q) = make_float4(pB[0].x, pB[0].y, pB[1].x, pB[1].y);
        *reinterpret_cast<float4*>(trw + 16 + 4 * q) = make_float4(pC[0].x, pC[0].y, pC[1].x, pC[1].y);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) hk[j] = hp[j];
      // per-channel outputs of the step
      {
        const float r1s = r1.x + r1.y, r2s = r2.x + r2.y;
        const float duo = fmaf(dtv[k], r1s, dyv[k] * Dsk_r);
        const float ddt = fmaf(r2s, kLn2f, uu[k] * r1s);
        const float ddl = ddt * sgv[k];
        dD_acc = fmaf(dyv[k], uu[k], dD_acc);
        const bool valid = k < nvalid;
        db_acc += valid ? ddl : 0.f;
        const int64_t s = sb0 + k;
        St<T>::pred(pdu + s * du_ss, duo, valid);
        St<T>::pred(pddl + s * ddl_ss, ddl, valid);
      }
      // sum of the 32 products over the warp's 32 channels: transpose through shared memory
      __syncwarp();
      {
        float4 acc = *reinterpret_cast<const float4*>(trr);
#pragma unroll
        for (int s = 1; s < 8; ++s) {
          const float4 v = *reinterpret_cast<const float4*>(trr + s * kTrPad);
          const float2 lo = fadd2(f2(acc.x, acc.y), f2(v.x, v.y)), hi = fadd2(f2(acc.z, acc.w), f2(v.z, v.w));
          acc = make_float4(lo.x, lo.y, hi.x, hi.y);
        }
#pragma unroll
        for (int m = 8; m <= 16; m <<= 1) {
          const float4 o = make_float4(__shfl_xor_sync(0xffffffffu, acc.x, m), __shfl_xor_sync(0xffffffffu, acc.y, m),
                                       __shfl_xor_sync(0xffffffffu, acc.z, m), __shfl_xor_sync(0xffffffffu, acc.w, m));
          const float2 lo = fadd2(f2(acc.x, acc.y), f2(o.x, o.y)), hi = fadd2(f2(acc.z, acc.w), f2(o.z, o.w));
          acc = make_float4(lo.x, lo.y, hi.x, hi.y);
        }
        if (lane < 8) *reinterpret_cast<float4*>(&S.red[pbuf][warp][k][4 * lane]) = acc;
      }
      __syncwarp();                              // tr may be overwritten by the next step
    }

    __syncthreads();                             // every warp is done with this stage; red[pbuf] is complete
    if (threadIdx.x == 0 && i + kStages < ntot) {
      tma::fence_proxy_async_smem();
      issue(i + kStages);
    }
    // sum over the CTA's warps -> one 128-byte partial row per step
    {
      float* part = d.dBC_part + ((int64_t)b * P.n_blk + blockIdx.x) * (int64_t)L * 32 + d.part_l0;
#pragma unroll
      for (int o = threadIdx.x; o < kTB * 32; o += kW * 32) {
        const int k = o >> 5, v = o & 31;
        float acc = S.red[pbuf][0][k][v];
#pragma unroll
        for (int w = 1; w < kW; ++w) acc += S.red[pbuf][w][k][v];
        if (k < nvalid) part[(int64_t)(sb0 + k) * d.part_ss + v] = acc;
      }
    }
  }

  // per-row sums over time
  {
    const int64_t row = (int64_t)b * P.dim + ch;
    float4* da = reinterpret_cast<float4*>(d.dA_part + row * 16);
#pragma unroll
    for (int q = 0; q < 4; ++q) da[q] = make_float4(dA[2 * q].x, dA[2 * q].y, dA[2 * q + 1].x, dA[2 * q + 1].y);
    if (d.dD_part) d.dD_part[row] = dD_acc;
    if (d.dbias_part) d.dbias_part[row] = db_acc;
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
static bool build_params(const cm_scan_bwd_args& a, BwdParams* P, bool maps) {
  constexpr int ES = (int)sizeof(T);
  if (a.dstate != 16 || a.dim % kCH != 0) return false;
  if (a.lanes_per_channel != 0 && a.lanes_per_channel != 1) return false;
  if (a.dout.ptr == nullptr || a.dout.sd != 1) return false;
  const bool has_z = a.z.ptr != nullptr;
  if (has_z && (a.z.sd != 1 || a.out_pre.ptr == nullptr || a.out_pre.sd != 1 || a.dz.ptr == nullptr || a.dz.sd != 1)) return false;
  const int64_t L = a.seqlen, Bt = a.batch, D = a.dim;
  P->L = a.seqlen; P->ndir = a.ndir; P->n_blk = a.dim / kCH; P->dim = a.dim; P->has_z = has_z;
  P->flags = a.flags; P->scale = a.out_scale;
  P->dz = nullptr; P->dz_sb = 0; P->dz_sl = 0;
  if (has_z) {
    P->dz = static_cast<char*>(a.dz.ptr);
    P->dz_sb = a.dz.sb * ES;
    if (!step_stride32(a.dz.sl, ES, false, &P->dz_sl)) return false;
  }
  auto ok_map = [&](CUtensorMap* m, const cm_tensor3& t, int64_t ch, int box_c) {
    if (t.sd != 1) return false;
    // TMA addressability is a property of the view (alignment, strides): checked identically with or without encoding
    const uint64_t sl = (uint64_t)t.sl * ES, sb = (uint64_t)t.sb * ES;
    if ((reinterpret_cast<uintptr_t>(t.ptr) & 15) != 0 || (sl & 15) != 0 || t.sl <= 0) return false;
    if (Bt > 1 && ((sb & 15) != 0 || t.sb <= 0)) return false;
    if (!maps) return true;
    return tma::make_map_blc(m, t.ptr, ES, ch, L, Bt, t.sl, t.sb, box_c, kTB);
  };
  if (!ok_map(&P->m_go, a.dout, D, kCH)) return false;
  if (has_z && (!ok_map(&P->m_z, a.z, D, kCH) || !ok_map(&P->m_pre, a.out_pre, D, kCH))) return false;
  for (int r = 0; r < a.ndir; ++r) {
    const cm_scan_bwd_dir& sd = a.dir[r];
    const cm_scan_dir& s = sd.in;
    BwdDir& d = P->dir[r];
    if (s.bc_const) return false;
    if (s.Bm.sd != 1 || s.Cm.sd != 1) return false;
    if (static_cast<const char*>(s.Cm.ptr) != static_cast<const char*>(s.Bm.ptr) + 16 * ES || s.Cm.sl != s.Bm.sl || s.Cm.sb != s.Bm.sb)
      return false;
    if (sd.du.sd != 1 || sd.ddelta.sd != 1) return false;
    if (s.ckpt == nullptr || (reinterpret_cast<uintptr_t>(s.ckpt) & 15) != 0 || (s.ckpt_sb % 4) != 0 || (s.ckpt_sd % 4) != 0)
      return false;
    if ((reinterpret_cast<uintptr_t>(sd.dBC_part) & 15) != 0 || (reinterpret_cast<uintptr_t>(sd.dA_part) & 15) != 0) return false;
    if (!ok_map(&d.m_u, s.u, D, kCH) || !ok_map(&d.m_dl, s.delta, D, kCH) || !ok_map(&d.m_bc, s.Bm, 32, 32)) return false;
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
    d.ckpt = s.ckpt; d.ckpt_sb = s.ckpt_sb; d.ckpt_sd = s.ckpt_sd;
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
  const bool sp = (a.flags & CM_FLAG_DELTA_SOFTPLUS) != 0;
  auto launch = [&](auto kern) -> int {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);   // per device
    if (e != cudaSuccess) return (int)e;
    kern<<<dim3(a.dim / kCH, a.batch, a.ndir), kW * 32, smem, st>>>(P);
    e = cudaGetLastError();
    return e == cudaSuccess ? 0 : (int)e;
  };
  *rc = sp ? launch(scan_bwd_lc_kernel<T, true>) : launch(scan_bwd_lc_kernel<T, false>);
  return 1;
}

}  // namespace lcb

// 1 if cm_scan_bwd will take the lane-per-channel TMA kernel for these arguments (a pure function of the argument block:
// the caller sizes the dB/dC partial tensor from it through cm_scan_bwd_slab_channels)
int scan_bwd_lane_channel_applies(const cm_scan_bwd_args& a) {
  // Off by default: measured 0.87 ms against 0.63 ms of the state-parallel kernel at the ConMamba-large shape (two fat
  // warps per SM sub-partition issue 0.32 instructions per cycle).  CM_SCAN_LC_BWD=1 selects it.
  const char* on = getenv("CM_SCAN_LC_BWD");
  if (on == nullptr || on[0] == '0') return 0;
  if (getenv("CM_SCAN_NO_LC") != nullptr || getenv("CM_SCAN_NO_SP") != nullptr || getenv("CM_SCAN_GENERIC") != nullptr) return 0;
  lcb::BwdParams P;
  switch (a.dtype) {
    case CM_F32: return lcb::build_params<float>(a, &P, false) ? 1 : 0;
    case CM_BF16: return lcb::build_params<__nv_bfloat16>(a, &P, false) ? 1 : 0;
    default: return lcb::build_params<__half>(a, &P, false) ? 1 : 0;
  }
}
int scan_bwd_lane_channel_slab() { return lcb::kCH; }

// returns 1 if launched (result in *rc), 0 if the lane-per-channel TMA path does not apply
int scan_bwd_try_lane_channel(const cm_scan_bwd_args& a, cudaStream_t st, int* rc) {
  switch (a.dtype) {
    case CM_F32: return lcb::try_t<float>(a, st, rc);
    case CM_BF16: return lcb::try_t<__nv_bfloat16>(a, st, rc);
    default: return lcb::try_t<__half>(a, st, rc);
  }
}

}  // namespace cm
