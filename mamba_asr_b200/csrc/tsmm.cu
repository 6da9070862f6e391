// Tall-skinny weight-gradient GEMM  C[M, N] = sum_r A[r, m] * B[r, n]  (A^T B over tens of thousands of rows) for sm_100a.
//
// These are the weight gradients of the Mamba block's skinny projections (reference
// modules/mamba/selective_scan_interface.py:277-283): d dt_proj.weight = ddelta^T x_dbl[:, :R] (D x R, R = 9..32) and
// d x_proj.weight = dx_dbl^T conv1d_out ((R + 2N) x D, evaluated here as (conv1d_out^T dx_dbl)^T).  The reduction dimension
// is batch * L (12032 .. 32064) and one operand is at most 64 columns wide, so the product is a column reduction, not a
// GEMM tile problem: cuBLAS needs 17-26 us for each of them on B200 (a batched split over the utterances plus a sum;
// 13 % of the ConMamba-small step), the bytes are worth 2-5 us.
//
// One CTA owns 64 columns of A and a chunk of rows: 4 warps, each a 16 x N strip of the output held in registers
// (mma.sync m16n8k16 bf16 -> fp32; this is tensor-core work the north star allots to the projections), operand tiles of 128
// rows staged by cp.async into a double-buffered shared ring (32-row stages left the kernel waiting on one memory round
// trip per two k-steps: 14 us at 12032 x 288 x 48) and read with ldmatrix.trans (both operands are stored
// row = reduction index).  Every CTA writes its partial [M-tile x N] block; cm_reduce_multi sums the row chunks in a fixed
// order (deterministic, no atomics).  Roof: HBM - rows * (M + N) * 2 bytes.
#include <mma.h>

#include "common.cuh"

namespace cm {

constexpr int kTsBM = 64;      // columns of A per CTA
constexpr int kTsKS = 128;     // rows per pipeline stage: one memory round trip buys 8 k-steps of MMAs
constexpr int kTsKCmin = 256;  // rows per CTA (row chunk) are a multiple of this; see tsmm_chunk_rows()
constexpr int kTsMaxN = 64;
constexpr int kTsAPad = kTsBM + 8;    // 144-byte rows: ldmatrix row addresses fall on distinct 16-byte bank groups

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
  const uint32_t s = static_cast<uint32_t>(__cvta_generic_to_shared(smem));
  const int sz = valid ? 16 : 0;     // src-size 0: the 16 bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void ldsm_x4_t(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, const void* p) {
  const uint32_t s = static_cast<uint32_t>(__cvta_generic_to_shared(p));
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(s));
}
__device__ __forceinline__ void ldsm_x2_t(uint32_t& r0, uint32_t& r1, const void* p) {
  const uint32_t s = static_cast<uint32_t>(__cvta_generic_to_shared(p));
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(s));
}
template <typename T> struct MmaBf;
template <> struct MmaBf<__nv_bfloat16> {
  static __device__ __forceinline__ void mma(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  }
};
template <> struct MmaBf<__half> {
  static __device__ __forceinline__ void mma(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  }
};

// NT = number of 8-column tiles of B (N = 8 * NT)
template <typename T, int NT>
__global__ void __launch_bounds__(128) tsmm_kernel(const T* __restrict__ A, int64_t lda, const T* __restrict__ B, int64_t ldb,
                                                   float* __restrict__ part, int64_t rows, int M, int kc) {
  constexpr int N = 8 * NT;
  constexpr int BPAD = N + 8;                       // 16-byte aligned rows, odd multiple of 16 bytes when N % 16 == 0
  extern __shared__ __align__(16) unsigned char ts_smem[];
  T (*sA)[kTsKS][kTsAPad] = reinterpret_cast<T (*)[kTsKS][kTsAPad]>(ts_smem);
  T (*sB)[kTsKS][BPAD] = reinterpret_cast<T (*)[kTsKS][BPAD]>(ts_smem + 2 * sizeof(T) * kTsKS * kTsAPad);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * kTsBM;
  const int64_t r_begin = (int64_t)blockIdx.y * kc;
  const int64_t r_end = min(r_begin + (int64_t)kc, rows);
  const int nstage = (int)((r_end - r_begin + kTsKS - 1) / kTsKS);

  auto load_stage = [&](int st, int buf) {
    const int64_t r0 = r_begin + (int64_t)st * kTsKS;
    // A tile: kTsKS rows x 64 columns, chunks of 8 elements
#pragma unroll
    for (int i = 0; i < kTsKS * 8 / 128; ++i) {
      const int ch = tid + 128 * i;
      const int r = ch >> 3, c8 = (ch & 7) * 8;
      const bool ok = (r0 + r < r_end) && (m0 + c8 < M);
      cp_async16(&sA[buf][r][c8], ok ? A + (r0 + r) * lda + m0 + c8 : A, ok);
    }
    // B tile: kTsKS rows x N columns
    for (int ch = tid; ch < kTsKS * NT; ch += 128) {
      const int r = ch / NT, c8 = (ch % NT) * 8;
      const bool okb = r0 + r < r_end;
      cp_async16(&sB[buf][r][c8], okb ? B + (r0 + r) * ldb + c8 : B, okb);
    }
    cp_async_commit();
  };

  float acc[NT][4];
#pragma unroll
  for (int j = 0; j < NT; ++j)
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[j][q] = 0.f;

  if (nstage > 0) load_stage(0, 0);
  for (int st = 0; st < nstage; ++st) {
    const int buf = st & 1;
    if (st + 1 < nstage) {
      load_stage(st + 1, buf ^ 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
#pragma unroll
    for (int ks = 0; ks < kTsKS; ks += 16) {
      // A operand (m16 x k16) of this warp's strip: element (m, k) = sA[ks + k][16 * warp + m]; four transposed 8x8 blocks
      uint32_t a[4];
      {
        const int mat = lane >> 3, rr = lane & 7;                 // matrix id, row of that 8x8 block (= k index)
        const int krow = ks + rr + ((mat & 2) ? 8 : 0);
        const int mcol = 16 * warp + ((mat & 1) ? 8 : 0);
        ldsm_x4_t(a[0], a[1], a[2], a[3], &sA[buf][krow][mcol]);
      }
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        uint32_t b[2];
        const int rr = lane & 7, mat = (lane >> 3) & 1;
        ldsm_x2_t(b[0], b[1], &sB[buf][ks + rr + 8 * mat][8 * j]);
        MmaBf<T>::mma(acc[j], a, b);
      }
    }
    __syncthreads();
  }
  // partial block of this row chunk: part[chunk][m][n]
  float* dst = part + ((int64_t)blockIdx.y * M) * N;
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int j = 0; j < NT; ++j) {
    const int mA = m0 + 16 * warp + g, mB = mA + 8, n = 8 * j + 2 * t;
    if (mA < M) *reinterpret_cast<float2*>(dst + (int64_t)mA * N + n) = make_float2(acc[j][0], acc[j][1]);
    if (mB < M) *reinterpret_cast<float2*>(dst + (int64_t)mB * N + n) = make_float2(acc[j][2], acc[j][3]);
  }
}

// rows per CTA: enough CTAs to fill the machine twice, few enough chunks that the partial blocks stay small
static int tsmm_chunk_rows(int64_t rows, int M) {
  const int64_t mt = (M + kTsBM - 1) / kTsBM;
  int64_t kc = rows * mt / (2 * 148);
  kc = (kc + kTsKCmin - 1) / kTsKCmin * kTsKCmin;
  if (kc < kTsKCmin) kc = kTsKCmin;
  if (kc > 2048) kc = 2048;
  return (int)kc;
}

template <typename T, int NT>
static int tsmm_launch_nt(const void* A, int64_t lda, const void* B, int64_t ldb, float* part, int64_t rows, int M,
                          cudaStream_t st) {
  const int kc = tsmm_chunk_rows(rows, M);
  const dim3 grid((M + kTsBM - 1) / kTsBM, (unsigned)((rows + kc - 1) / kc));
  const size_t smem = 2 * sizeof(T) * kTsKS * (kTsAPad + 8 * NT + 8);
  auto kern = tsmm_kernel<T, NT>;
  {   // per-device attribute: set on every launch (a process-wide flag would miss the second GPU of a process)
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  kern<<<grid, 128, smem, st>>>(static_cast<const T*>(A), lda, static_cast<const T*>(B), ldb, part, rows, M, kc);
  CM_LAUNCH_CHECK();
  return 0;
}

template <typename T>
static int tsmm_launch(const void* A, int64_t lda, const void* B, int64_t ldb, float* part, int64_t rows, int M, int N,
                       cudaStream_t st) {
  switch (N / 8) {
    case 1: return tsmm_launch_nt<T, 1>(A, lda, B, ldb, part, rows, M, st);
    case 2: return tsmm_launch_nt<T, 2>(A, lda, B, ldb, part, rows, M, st);
    case 3: return tsmm_launch_nt<T, 3>(A, lda, B, ldb, part, rows, M, st);
    case 4: return tsmm_launch_nt<T, 4>(A, lda, B, ldb, part, rows, M, st);
    case 5: return tsmm_launch_nt<T, 5>(A, lda, B, ldb, part, rows, M, st);
    case 6: return tsmm_launch_nt<T, 6>(A, lda, B, ldb, part, rows, M, st);
    case 7: return tsmm_launch_nt<T, 7>(A, lda, B, ldb, part, rows, M, st);
    case 8: return tsmm_launch_nt<T, 8>(A, lda, B, ldb, part, rows, M, st);
    default: return CM_ERR_UNSUPPORTED;
  }
}

}  // namespace cm

extern "C" int cm_tsmm_num_part(int64_t rows, int32_t M) {
  if (rows <= 0 || M <= 0) return CM_ERR_BAD_ARG;
  const int kc = cm::tsmm_chunk_rows(rows, M);
  return (int)((rows + kc - 1) / kc);
}

extern "C" int cm_tsmm(const void* A, int64_t lda, const void* B, int64_t ldb, float* part, int64_t rows, int32_t M, int32_t N,
                       int32_t dtype, void* stream) {
  if (!A || !B || !part || rows <= 0 || M <= 0 || N <= 0) return CM_ERR_BAD_ARG;
  if (dtype != CM_BF16 && dtype != CM_F16) return CM_ERR_UNSUPPORTED;
  if (N > cm::kTsMaxN || (N & 7) || (M & 7) || (lda & 7) || (ldb & 7)) return CM_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(B) & 15) || (reinterpret_cast<uintptr_t>(part) & 7))
    return CM_ERR_UNSUPPORTED;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == CM_BF16) return cm::tsmm_launch<__nv_bfloat16>(A, lda, B, ldb, part, rows, M, N, st);
  return cm::tsmm_launch<__half>(A, lda, B, ldb, part, rows, M, N, st);
}
