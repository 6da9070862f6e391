// cm_reduce_batch: the fixed-order column sums of MANY partial-sum buffers in one launch (sm_100a).
//
// Every backward kernel of the library leaves its cross-CTA sums (dgamma / dbeta of the LayerNorms, conv and depthwise-conv
// dweight / dbias, dA / dD / d(delta_bias) of the scan, bias-gradient column sums, split-K weight-gradient partial products)
// as per-CTA partial rows [rows][cols] fp32 and a reducer adds the rows in a fixed order (deterministic, no atomics; the
// reference's causal-conv1d / selective-scan kernels use fp32 atomics: SURVEY.md section 2.2).  Reduced where they are
// produced that is 221 cm_reduce_multi + 219 at::reduce_kernel launches of ~4 us per ConMamba-large training step (1.9 of
// 42 ms, profiles/r02_step_profile_large.txt): each is a few hundred rows of a few hundred columns - pure latency.  The host
// side (kernels.reduce_many with deferral on) queues the jobs of a whole backward pass and this kernel runs them 64 at a time:
// a 1-D grid, each CTA finds its job by binary search over the prefix sums of the jobs' CTA counts.
//   tall jobs (rows > 32): a CTA of 32 x 8 threads owns 128 columns (a thread 4: one 16-byte access per row where the job's
//                          geometry allows); its 8 row lanes stride over the rows with 8 loads in flight per thread and are
//                          combined through shared memory in lane order;
//   wide jobs (rows <= 32, the split-K sums: cols up to 2^18): a thread owns 4 consecutive columns, a CTA 1024.
// Roof: HBM / L2 latency; algorithmic bytes: rows * cols * 4 read, cols * 4 written per job.
#include "common.cuh"

namespace cm {

struct ReduceBatch {
  cm_reduce_job2 j[CM_REDUCE_BATCH_MAX];
  int32_t first[CM_REDUCE_BATCH_MAX + 1];   // first CTA of job i; first[njobs] = grid size
  int32_t njobs;
};

constexpr int kRbWide = 32;          // jobs of at most this many rows take the wide path
constexpr int kRbLanes = 8;          // row lanes of a tall CTA
constexpr int kRbFly = 8;            // independent 16-byte loads in flight per thread

// four consecutive columns starting at c of row pointer `row`: one 16-byte load where the job's geometry allows
__device__ __forceinline__ float4 rb_ld4(const float* row, int64_t c, int64_t cols, bool vec) {
  if (vec) return __ldg(reinterpret_cast<const float4*>(row + c));
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c < cols) v.x = __ldg(row + c);
  if (c + 1 < cols) v.y = __ldg(row + c + 1);
  if (c + 2 < cols) v.z = __ldg(row + c + 2);
  if (c + 3 < cols) v.w = __ldg(row + c + 3);
  return v;
}
__device__ __forceinline__ void rb_st4(float* out, int64_t c, int64_t cols, bool vec, const float4& v) {
  if (vec) { *reinterpret_cast<float4*>(out + c) = v; return; }
  if (c < cols) out[c] = v.x;
  if (c + 1 < cols) out[c + 1] = v.y;
  if (c + 2 < cols) out[c + 2] = v.z;
  if (c + 3 < cols) out[c + 3] = v.w;
}
__device__ __forceinline__ void rb_add(float4& a, const float4& v) { a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w; }

// CTA = 32 x 8 threads.  (The first version used 32 x 32 threads per 32 columns: 4000 - 6500 CTAs of 1024 threads per launch,
// 1.6 - 2 TB/s with 45 % of the issue slots busy - CTA turnover, not memory, profiles/r02_reduce_batch_kernel_ncu.txt.)
//   wide job: a thread owns 4 columns, the CTA 1024; rows added in index order, kRbFly loads in flight
//   tall job: a thread owns 4 columns of a row lane, the CTA 128 columns x 8 lanes; lane y adds rows y, y + 8, ... in groups of
//             kRbFly, the lanes are then added in lane order through shared memory
__global__ void __launch_bounds__(32 * kRbLanes) reduce_batch_kernel(const __grid_constant__ ReduceBatch rb) {
  __shared__ float4 sm[kRbLanes][32];
  // job of this CTA: largest i with first[i] <= blockIdx.x
  int lo = 0, hi = rb.njobs - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (rb.first[mid] <= (int)blockIdx.x) lo = mid;
    else hi = mid - 1;
  }
  const cm_reduce_job2& job = rb.j[lo];
  const int blk = blockIdx.x - rb.first[lo];
  const int64_t rows = job.rows, cols = job.cols, stride = job.stride;
  const bool vec = ((cols | stride) & 3) == 0 &&
                   ((reinterpret_cast<uintptr_t>(job.part) | reinterpret_cast<uintptr_t>(job.out)) & 15) == 0;
  if (rows <= kRbWide) {
    const int64_t c = ((int64_t)blk * (32 * kRbLanes) + threadIdx.y * 32 + threadIdx.x) * 4;
    if (c >= cols) return;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int64_t r = 0;
    for (; r + kRbFly <= rows; r += kRbFly) {
      float4 v[kRbFly];
#pragma unroll
      for (int i = 0; i < kRbFly; ++i) v[i] = rb_ld4(job.part + (r + i) * stride, c, cols, vec);
#pragma unroll
      for (int i = 0; i < kRbFly; ++i) rb_add(acc, v[i]);
    }
    for (; r < rows; ++r) rb_add(acc, rb_ld4(job.part + r * stride, c, cols, vec));
    rb_st4(job.out, c, cols, vec, acc);
    return;
  }
  const int64_t c = ((int64_t)blk * 32 + threadIdx.x) * 4;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c < cols) {
    int64_t r = threadIdx.y;
    for (; r + (kRbFly - 1) * kRbLanes < rows; r += kRbFly * kRbLanes) {
      float4 v[kRbFly];
#pragma unroll
      for (int i = 0; i < kRbFly; ++i) v[i] = rb_ld4(job.part + (r + i * kRbLanes) * stride, c, cols, vec);
#pragma unroll
      for (int i = 0; i < kRbFly; ++i) rb_add(acc, v[i]);
    }
    for (; r < rows; r += kRbLanes) rb_add(acc, rb_ld4(job.part + r * stride, c, cols, vec));
  }
  sm[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && c < cols) {
    float4 t = sm[0][threadIdx.x];
#pragma unroll
    for (int y = 1; y < kRbLanes; ++y) rb_add(t, sm[y][threadIdx.x]);
    rb_st4(job.out, c, cols, vec, t);
  }
}

}  // namespace cm

extern "C" int cm_reduce_batch(const cm_reduce_job2* jobs, int32_t njobs, void* stream) {
  if (jobs == nullptr || njobs <= 0 || njobs > CM_REDUCE_BATCH_MAX) return CM_ERR_BAD_ARG;
  cm::ReduceBatch rb;
  int64_t total = 0;
  for (int i = 0; i < njobs; ++i) {
    const cm_reduce_job2& j = jobs[i];
    if (!j.part || !j.out || j.rows <= 0 || j.cols <= 0 || j.stride < j.cols) return CM_ERR_BAD_ARG;
    rb.j[i] = j;
    rb.first[i] = (int32_t)total;
    total += (j.rows <= cm::kRbWide) ? (j.cols + 1023) / 1024 : (j.cols + 127) / 128;
    if (total > 0x7fffffff) return CM_ERR_UNSUPPORTED;
  }
  for (int i = njobs; i < CM_REDUCE_BATCH_MAX; ++i) rb.j[i] = jobs[0];
  for (int i = njobs; i <= CM_REDUCE_BATCH_MAX; ++i) rb.first[i] = (int32_t)total;
  rb.njobs = njobs;
  cm::reduce_batch_kernel<<<(unsigned)total, dim3(32, cm::kRbLanes), 0, static_cast<cudaStream_t>(stream)>>>(rb);
  CM_LAUNCH_CHECK();
  return 0;
}
