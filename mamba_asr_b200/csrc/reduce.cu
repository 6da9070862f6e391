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
//   tall jobs (rows > 32): a CTA of 32 x 32 threads owns 32 columns; the 32 warps stride over the rows with 8 loads in flight
//                          per thread and are combined through shared memory in warp order;
//   wide jobs (rows <= 32, the split-K sums: cols up to 2^18): a thread owns 4 consecutive columns (one 16-byte access per
//                          row where the job's geometry allows), a CTA 4096 columns.
// Roof: HBM / L2 latency; algorithmic bytes: rows * cols * 4 read, cols * 4 written per job.
#include "common.cuh"

namespace cm {

struct ReduceBatch {
  cm_reduce_job2 j[CM_REDUCE_BATCH_MAX];
  int32_t first[CM_REDUCE_BATCH_MAX + 1];   // first CTA of job i; first[njobs] = grid size
  int32_t njobs;
};

constexpr int kRbWide = 32;          // jobs of at most this many rows take the wide path
constexpr int kRbFly = 8;

__global__ void __launch_bounds__(1024) reduce_batch_kernel(const __grid_constant__ ReduceBatch rb) {
  __shared__ float sm[32][33];
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
  const int tid = threadIdx.y * 32 + threadIdx.x;
  if (rows <= kRbWide) {
    const int64_t c0 = ((int64_t)blk * 1024 + tid) * 4;
    if (c0 >= cols) return;
    const bool vec = ((cols | stride) & 3) == 0 && ((reinterpret_cast<uintptr_t>(job.part) | reinterpret_cast<uintptr_t>(job.out)) & 15) == 0;
    if (vec) {
      const float4* src = reinterpret_cast<const float4*>(job.part + c0);
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      int64_t r = 0;
      for (; r + 4 <= rows; r += 4) {
        float4 v[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = __ldg(src + (r + i) * (stride >> 2));
#pragma unroll
        for (int i = 0; i < 4; ++i) { acc.x += v[i].x; acc.y += v[i].y; acc.z += v[i].z; acc.w += v[i].w; }
      }
      for (; r < rows; ++r) {
        const float4 v = __ldg(src + r * (stride >> 2));
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
      *reinterpret_cast<float4*>(job.out + c0) = acc;
    } else {
      for (int k = 0; k < 4 && c0 + k < cols; ++k) {
        float acc = 0.f;
        for (int64_t r = 0; r < rows; ++r) acc += __ldg(job.part + r * stride + c0 + k);
        job.out[c0 + k] = acc;
      }
    }
    return;
  }
  const int64_t c = (int64_t)blk * 32 + threadIdx.x;
  float acc[kRbFly];
#pragma unroll
  for (int i = 0; i < kRbFly; ++i) acc[i] = 0.f;
  if (c < cols) {
    const float* src = job.part + c;
    int64_t r = threadIdx.y;
    for (; r + (kRbFly - 1) * 32 < rows; r += kRbFly * 32) {
      float v[kRbFly];
#pragma unroll
      for (int i = 0; i < kRbFly; ++i) v[i] = __ldg(src + (r + i * 32) * stride);
#pragma unroll
      for (int i = 0; i < kRbFly; ++i) acc[i] += v[i];
    }
    float v[kRbFly];
#pragma unroll
    for (int i = 0; i < kRbFly; ++i) v[i] = (r + i * 32 < rows) ? __ldg(src + (r + i * 32) * stride) : 0.f;
#pragma unroll
    for (int i = 0; i < kRbFly; ++i) acc[i] += v[i];
  }
  sm[threadIdx.y][threadIdx.x] = ((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7]));
  __syncthreads();
  if (threadIdx.y == 0 && c < cols) {
    float t = sm[0][threadIdx.x];
#pragma unroll
    for (int y = 1; y < 32; ++y) t += sm[y][threadIdx.x];
    job.out[c] = t;
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
    total += (j.rows <= cm::kRbWide) ? (j.cols + 4095) / 4096 : (j.cols + 31) / 32;
    if (total > 0x7fffffff) return CM_ERR_UNSUPPORTED;
  }
  for (int i = njobs; i < CM_REDUCE_BATCH_MAX; ++i) rb.j[i] = jobs[0];
  for (int i = njobs; i <= CM_REDUCE_BATCH_MAX; ++i) rb.first[i] = (int32_t)total;
  rb.njobs = njobs;
  cm::reduce_batch_kernel<<<(unsigned)total, dim3(32, 32), 0, static_cast<cudaStream_t>(stream)>>>(rb);
  CM_LAUNCH_CHECK();
  return 0;
}
