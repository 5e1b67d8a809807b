// Device Philox generators that reproduce torch.randn / torch.rand / Tensor.uniform_ on CUDA
// bit-for-bit (same seed, same generator offset), plus slice generation for batch sharding.
//
// Replaces, on the hot path: NoiseGenerator.rand_like (reference py/noise_generation.py:133-155)
// and the direct torch.randn calls in PyramidNoiseGenerator.generate (:632-640),
// PerlinOldNoiseGenerator.perlin_noise (:465-469) and PowerNoiseItem (py/nodes/powernoise.py:396-401).
#include "common.cuh"
#include "../../include/sonar_b200.h"

namespace sonar {

enum DistKind : int { DIST_NORMAL = 0, DIST_UNIFORM = 1 };

// One "pair" = one Philox block of one emulated ATen thread = 4 output elements
//   li = t + T*(4k + lane).
// The launch covers pairs p in [0, T*k_count): t = p % T, k = k_lo + p / T. With k_lo = 0 and
// k_count = ceil(numel / 4T) this is exactly ATen's grid-stride loop. A slice [begin, end) of
// the virtual tensor only keeps the lanes that fall inside it and writes them at (li - begin).
// The four lanes of a call land T elements apart. Calls that lie wholly inside the slice (all but the first and the
// last call row of a slice) take one pointer and three pointer bumps instead of four 64-bit range tests.
__device__ __forceinline__ void store_lanes(float* __restrict__ out, const float4& v, int64_t li0, int64_t T, int64_t begin,
                                            int64_t end) {
  if (li0 >= begin && li0 + 3 * T < end) {
    float* p = out + (li0 - begin);
    p[0] = v.x;
    p += T;
    p[0] = v.y;
    p += T;
    p[0] = v.z;
    p += T;
    p[0] = v.w;
    return;
  }
  const float vals[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int lane = 0; lane < 4; ++lane) {
    const int64_t li = li0 + T * lane;
    if (li >= begin && li < end) out[li - begin] = vals[lane];
  }
}

template <int KIND>
__global__ void __launch_bounds__(kBlock)
philox_fill_kernel(float* __restrict__ out, int64_t begin, int64_t end, PhiloxStream s, uint32_t k_lo, uint32_t k_hi,
                   float p0, float p1) {
  const int64_t T = s.threads;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  // virtual ATen thread vt (Philox subsequence) x call k; no 64-bit div/mod in the loop
  for (int64_t vt = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; vt < T; vt += nthreads) {
    for (uint32_t k = k_lo; k <= k_hi; ++k) {
      const int64_t li0 = vt + T * (int64_t)(4 * (uint64_t)k);
      if (li0 >= end) break;
      if (li0 + 3 * T < begin) continue;
      float4 v;
      if (KIND == DIST_NORMAL) {
        v = philox_normal4(s, (uint32_t)vt, k);
        v.x = v.x * p1 + p0;  // at::transformation::normal: val * std + mean
        v.y = v.y * p1 + p0;
        v.z = v.z * p1 + p0;
        v.w = v.w * p1 + p0;
      } else {
        v = philox_uniform4(s, (uint32_t)vt, k);
        v.x = uniform_transform(v.x, p0, p1);
        v.y = uniform_transform(v.y, p0, p1);
        v.z = uniform_transform(v.z, p0, p1);
        v.w = uniform_transform(v.w, p0, p1);
      }
      store_lanes(out, v, li0, T, begin, end);
    }
  }
}

// Several draws in ONE launch (blockIdx.y = draw): a pyramid sample needs its base, its full-size
// level and 2-3 tiny coarse levels, a Perlin sample its base and two angle grids -- five / three
// launches of a few microseconds of work each, i.e. mostly launch latency. Draws keep their own ATen
// geometry (T = 256 * grid_blocks), offset and transform, so every value is what a separate
// torch.randn / Tensor.uniform_ call would have produced.
struct FillBatchDev {
  float* out[SONAR_FILL_BATCH_MAX];
  int64_t begin[SONAR_FILL_BATCH_MAX], end[SONAR_FILL_BATCH_MAX];
  uint64_t offset[SONAR_FILL_BATCH_MAX];
  uint32_t threads[SONAR_FILL_BATCH_MAX], k_lo[SONAR_FILL_BATCH_MAX], k_hi[SONAR_FILL_BATCH_MAX];
  int32_t kind[SONAR_FILL_BATCH_MAX];
  float p0[SONAR_FILL_BATCH_MAX], p1[SONAR_FILL_BATCH_MAX];
  uint64_t seed;
};

__global__ void __launch_bounds__(kBlock)
philox_fill_batch_kernel(FillBatchDev b) {
  const int d = blockIdx.y;
  const PhiloxStream s{b.seed, b.offset[d], b.threads[d]};
  const int64_t T = s.threads, begin = b.begin[d], end = b.end[d];
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  float* __restrict__ out = b.out[d];
  const float p0 = b.p0[d], p1 = b.p1[d];
  const bool normal = b.kind[d] == DIST_NORMAL;
  const uint32_t k_lo = b.k_lo[d], k_hi = b.k_hi[d];
  for (int64_t vt = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; vt < T; vt += nthreads) {
    for (uint32_t k = k_lo; k <= k_hi; ++k) {
      const int64_t li0 = vt + T * (int64_t)(4 * (uint64_t)k);
      if (li0 >= end) break;
      if (li0 + 3 * T < begin) continue;
      float4 v;
      if (normal) {
        v = philox_normal4(s, (uint32_t)vt, k);
        v.x = v.x * p1 + p0;
        v.y = v.y * p1 + p0;
        v.z = v.z * p1 + p0;
        v.w = v.w * p1 + p0;
      } else {
        v = philox_uniform4(s, (uint32_t)vt, k);
        v.x = uniform_transform(v.x, p0, p1);
        v.y = uniform_transform(v.y, p0, p1);
        v.z = uniform_transform(v.z, p0, p1);
        v.w = uniform_transform(v.w, p0, p1);
      }
      store_lanes(out, v, li0, T, begin, end);
    }
  }
}

template <int KIND>
int launch_fill(float* out, int64_t begin, int64_t count, int64_t numel_total, uint64_t seed,
                uint64_t offset, uint32_t grid_blocks, float p0, float p1, cudaStream_t stream) {
  if (count <= 0) return 0;
  if (grid_blocks == 0 || begin < 0 || begin + count > numel_total) return (int)cudaErrorInvalidValue;
  PhiloxStream s{seed, offset, grid_blocks * (uint32_t)kBlock};
  const int64_t T = s.threads;
  const int64_t end = begin + count;
  // rows r = li / T touched by the slice -> Philox calls k = r / 4
  const int64_t k_lo = (begin / T) / 4;
  const int64_t k_hi = ((end - 1) / T) / 4;
  const int grid = streaming_grid_shared(T, kBlock, 1);
  philox_fill_kernel<KIND><<<grid, kBlock, 0, stream>>>(out, begin, end, s, (uint32_t)k_lo, (uint32_t)k_hi, p0, p1);
  SONAR_LAUNCH_CHECK();
  return 0;
}

}  // namespace sonar

extern "C" {

int sonar_philox_policy(int64_t numel, uint32_t* grid_blocks, uint64_t* counter_offset) {
  // ATen/native/cuda/DistributionTemplates.h:50-62 (calc_execution_policy), unroll = 4
  if (numel <= 0) {
    if (grid_blocks) *grid_blocks = 0;
    if (counter_offset) *counter_offset = 0;
    return 0;
  }
  const sonar::DeviceInfo& di = sonar::device_info();
  uint64_t n = (uint64_t)numel;
  uint64_t grid = (n + sonar::kBlock - 1) / sonar::kBlock;
  uint64_t cap = (uint64_t)di.sm_count * (uint64_t)(di.max_threads_per_sm / sonar::kBlock);
  if (grid > cap) grid = cap;
  if (grid_blocks) *grid_blocks = (uint32_t)grid;
  if (counter_offset) *counter_offset = ((n - 1) / ((uint64_t)sonar::kBlock * grid * 4) + 1) * 4;
  return 0;
}

int sonar_philox_normal_f32(float* out, int64_t begin, int64_t count, int64_t numel_total, uint64_t seed,
                            uint64_t offset, uint32_t grid_blocks, float mean, float std, void* stream) {
  return sonar::launch_fill<sonar::DIST_NORMAL>(out, begin, count, numel_total, seed, offset, grid_blocks, mean,
                                                std, (cudaStream_t)stream);
}

int sonar_philox_uniform_f32(float* out, int64_t begin, int64_t count, int64_t numel_total, uint64_t seed,
                             uint64_t offset, uint32_t grid_blocks, float from, float to, void* stream) {
  return sonar::launch_fill<sonar::DIST_UNIFORM>(out, begin, count, numel_total, seed, offset, grid_blocks, from,
                                                 to, (cudaStream_t)stream);
}

int sonar_philox_fill_batch(const SonarFillBatch* batch, void* stream) {
  using namespace sonar;
  if (batch == nullptr || batch->n < 0 || batch->n > SONAR_FILL_BATCH_MAX) return (int)cudaErrorInvalidValue;
  FillBatchDev b;
  b.seed = batch->seed;
  int n = 0, gx = 1;
  for (int i = 0; i < batch->n; ++i) {
    const SonarFillDesc& d = batch->draws[i];
    if (d.count <= 0) continue;
    if (d.out == nullptr || d.grid_blocks == 0 || d.begin < 0 || d.begin + d.count > d.numel_total ||
        (d.kind != DIST_NORMAL && d.kind != DIST_UNIFORM))
      return (int)cudaErrorInvalidValue;
    const int64_t T = (int64_t)d.grid_blocks * kBlock, end = d.begin + d.count;
    b.out[n] = d.out;
    b.begin[n] = d.begin;
    b.end[n] = end;
    b.offset[n] = d.offset;
    b.threads[n] = (uint32_t)T;
    b.k_lo[n] = (uint32_t)((d.begin / T) / 4);
    b.k_hi[n] = (uint32_t)(((end - 1) / T) / 4);
    b.kind[n] = d.kind;
    b.p0[n] = d.p0;
    b.p1[n] = d.p1;
    const int g = streaming_grid_shared(T, kBlock, 1);
    if (g > gx) gx = g;
    ++n;
  }
  if (n == 0) return 0;
  for (int i = n; i < SONAR_FILL_BATCH_MAX; ++i) {  // unused slots: well-defined values
    b.out[i] = nullptr;
    b.begin[i] = b.end[i] = 0;
    b.offset[i] = 0;
    b.threads[i] = kBlock;
    b.k_lo[i] = 1;
    b.k_hi[i] = 0;
    b.kind[i] = 0;
    b.p0[i] = b.p1[i] = 0.0f;
  }
  philox_fill_batch_kernel<<<dim3((unsigned)gx, (unsigned)n), kBlock, 0, (cudaStream_t)stream>>>(b);
  SONAR_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
