// Global moments + the conditional normalisation of `scale_noise`.
//
// Reference: py/utils.py:85-106. The reference reduces mean/std over the WHOLE tensor (batch
// included), syncs them to the host with .item(), then conditionally centres / rescales:
//     if |mean| > thr: noise -= mean;  if |1 - std| > thr: noise /= std;  noise *= factor
// with thr = 2.5 / sqrt(numel) and the unbiased std. Here the two sums live in a device buffer
// (so a batch-sharded run can all-reduce them) and the conditional is evaluated on the device by
// the apply kernel itself: no host round trip.
#include "common.cuh"
#include "../../include/sonar_b200.h"

namespace sonar {

// ---------------------------------------------------------------------------------------------
// moments: sums[0] += sum(x), sums[1] += sum(x^2), double accumulation
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock)
moments_kernel(const float* __restrict__ x, int64_t n, int vec_ok, double* __restrict__ sums) {
  __shared__ double scratch[64];
  double s = 0.0, ss = 0.0;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  if (vec_ok) {
    const int64_t n4 = n >> 2;
    for (int64_t i = tid; i < n4; i += stride) {
      const float4 v = ld4_stream(x + 4 * i);
      // pairwise in float for the 4-vector keeps the fp64 pipe at 1/4 rate of the loads
      const double a = (double)v.x + (double)v.y, b = (double)v.z + (double)v.w;
      s += a + b;
      ss += ((double)v.x * v.x + (double)v.y * v.y) + ((double)v.z * v.z + (double)v.w * v.w);
    }
    for (int64_t i = (n4 << 2) + tid; i < n; i += stride) {
      const double v = x[i];
      s += v;
      ss += v * v;
    }
  } else {
    for (int64_t i = tid; i < n; i += stride) {
      const double v = x[i];
      s += v;
      ss += v * v;
    }
  }
  block_sum2(s, ss, scratch);
  if (threadIdx.x == 0) {
    atomicAdd(&sums[0], s);
    atomicAdd(&sums[1], ss);
  }
}

// Moments of a Philox normal draw restricted to the slice [begin, end) without materialising it.
__global__ void __launch_bounds__(kBlock)
philox_normal_moments_kernel(int64_t begin, int64_t end, PhiloxStream st, uint32_t k_lo, uint32_t k_hi,
                             double* __restrict__ sums) {
  __shared__ double scratch[64];
  double s = 0.0, ss = 0.0;
  const int64_t T = st.threads;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  for (int64_t vt = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; vt < T; vt += nthreads) {
    float fs = 0.0f, fss = 0.0f;  // per-thread partials stay fp32 (<= 4*calls terms), block/global sums fp64
    for (uint32_t k = k_lo; k <= k_hi; ++k) {
      const int64_t li0 = vt + T * (int64_t)(4 * (uint64_t)k);
      if (li0 >= end) break;
      if (li0 + 3 * T < begin) continue;
      const float4 v = philox_normal4(st, (uint32_t)vt, k);
      const float vals[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int lane = 0; lane < 4; ++lane) {
        const int64_t li = li0 + T * lane;
        if (li >= begin && li < end) {
          fs += vals[lane];
          fss += vals[lane] * vals[lane];
        }
      }
    }
    s += (double)fs;
    ss += (double)fss;
  }
  block_sum2(s, ss, scratch);
  if (threadIdx.x == 0) {
    atomicAdd(&sums[0], s);
    atomicAdd(&sums[1], ss);
  }
}

// Moments of SEVERAL future draws of the same shape in one launch (blockIdx.y = draw): a sampler
// knows, at its first noise request, the generator offsets of every remaining ancestral-noise draw
// of the run, so the statistics scale_noise needs are reduced once, ahead of time, and each sampler
// step becomes a single launch that regenerates its normals in registers (samplers.py).
constexpr int kMomentsBatch = 64;
struct OffsetBatch {
  uint64_t offset[kMomentsBatch];
};

__global__ void __launch_bounds__(kBlock)
philox_normal_moments_batch_kernel(int64_t begin, int64_t end, PhiloxStream st, OffsetBatch offs, uint32_t k_lo,
                                   uint32_t k_hi, double* __restrict__ sums) {
  __shared__ double scratch[64];
  st.offset = offs.offset[blockIdx.y];
  double s = 0.0, ss = 0.0;
  const int64_t T = st.threads;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  for (int64_t vt = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; vt < T; vt += nthreads) {
    float fs = 0.0f, fss = 0.0f;
    for (uint32_t k = k_lo; k <= k_hi; ++k) {
      const int64_t li0 = vt + T * (int64_t)(4 * (uint64_t)k);
      if (li0 >= end) break;
      if (li0 + 3 * T < begin) continue;
      float4 v;
      if (li0 + 2 * T < end) {
        v = philox_normal4(st, (uint32_t)vt, k);
      } else {  // lanes 2, 3 lie beyond the slice: one Box-Muller is enough
        const float2 lo = philox_normal2_lo(st, (uint32_t)vt, k);
        v = make_float4(lo.x, lo.y, 0.f, 0.f);
      }
      const float vals[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int lane = 0; lane < 4; ++lane) {
        const int64_t li = li0 + T * lane;
        if (li >= begin && li < end) {
          fs += vals[lane];
          fss += vals[lane] * vals[lane];
        }
      }
    }
    s += (double)fs;
    ss += (double)fss;
  }
  block_sum2(s, ss, scratch);
  if (threadIdx.x == 0) {
    atomicAdd(&sums[2 * blockIdx.y], s);
    atomicAdd(&sums[2 * blockIdx.y + 1], ss);
  }
}

// sums[i] -> the scale_noise decision of draw i as float[4] = {mean, std, subtract-mean flag, divide flag}:
// evaluated ONCE per draw (fp64), so the kernels that consume it load 16 bytes instead of running a
// fp64 divide + sqrt behind a barrier in every CTA.
__global__ void norm_decisions_kernel(const double* __restrict__ sums, int n, int64_t count, float threshold_std_devs,
                                      float* __restrict__ decisions) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const NormDecision d = decide_normalisation(sums + 2 * i, count, threshold_std_devs);
  reinterpret_cast<float4*>(decisions)[i] = make_float4(d.mean, d.std, d.sub_mean ? 1.0f : 0.0f, d.div_std ? 1.0f : 0.0f);
}

// One pass: materialise the slice of a Philox normal draw AND reduce its moments (for tensors too
// large to keep the normals in registers across a grid barrier: write once, read once).
__global__ void __launch_bounds__(kBlock)
philox_normal_fill_moments_kernel(float* __restrict__ out, int64_t begin, int64_t end, PhiloxStream st, uint32_t k_lo,
                                  uint32_t k_hi, double* __restrict__ sums) {
  __shared__ double scratch[64];
  double s = 0.0, ss = 0.0;
  const int64_t T = st.threads;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  for (int64_t vt = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; vt < T; vt += nthreads) {
    float fs = 0.0f, fss = 0.0f;
    for (uint32_t k = k_lo; k <= k_hi; ++k) {
      const int64_t li0 = vt + T * (int64_t)(4 * (uint64_t)k);
      if (li0 >= end) break;
      if (li0 + 3 * T < begin) continue;
      float4 v;
      if (li0 + 2 * T < end) {
        v = philox_normal4(st, (uint32_t)vt, k);
      } else {  // lanes 2, 3 lie beyond the slice: one Box-Muller is enough
        const float2 lo = philox_normal2_lo(st, (uint32_t)vt, k);
        v = make_float4(lo.x, lo.y, 0.f, 0.f);
      }
      const float vals[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int lane = 0; lane < 4; ++lane) {
        const int64_t li = li0 + T * lane;
        if (li >= begin && li < end) {
          out[li - begin] = vals[lane];
          fs += vals[lane];
          fss += vals[lane] * vals[lane];
        }
      }
    }
    s += (double)fs;
    ss += (double)fss;
  }
  block_sum2(s, ss, scratch);
  if (threadIdx.x == 0) {
    atomicAdd(&sums[0], s);
    atomicAdd(&sums[1], ss);
  }
}

// ---------------------------------------------------------------------------------------------
// scale_noise apply
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock)
scale_noise_kernel(const float* __restrict__ x, float* __restrict__ out, int64_t n, int vec_ok,
                   const double* __restrict__ sums, int64_t count, float factor, float threshold_std_devs,
                   const double* __restrict__ peer_mailbox, int peer_world, double peer_epoch) {
  __shared__ double peer_sums[2];
  __shared__ NormDecision nd_slot;
  const NormDecision nd = peer_world > 1 ? decide_normalisation_peers(peer_mailbox, peer_world, peer_epoch, count,
                                                                       threshold_std_devs, peer_sums)
                                         : decide_normalisation_block(sums, count, threshold_std_devs, &nd_slot);
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  if (vec_ok) {
    const int64_t n4 = n >> 2;
    for (int64_t i = tid; i < n4; i += stride) {
      float4 v = ld4(x + 4 * i);  // plain load: `out` may alias `x`
      v.x = apply_norm(v.x, nd) * factor;
      v.y = apply_norm(v.y, nd) * factor;
      v.z = apply_norm(v.z, nd) * factor;
      v.w = apply_norm(v.w, nd) * factor;
      st4(out + 4 * i, v);
    }
    for (int64_t i = (n4 << 2) + tid; i < n; i += stride) out[i] = apply_norm(x[i], nd) * factor;
  } else {
    for (int64_t i = tid; i < n; i += stride) out[i] = apply_norm(x[i], nd) * factor;
  }
}

// out = x * (scale / std(x)) with the unbiased std taken from device sums
// (GreenTestNoiseGenerator: noise *= scale / noise.std(), py/noise_generation.py:703)
__global__ void __launch_bounds__(kBlock)
scale_by_std_kernel(const float* __restrict__ x, float* __restrict__ out, int64_t n, const double* __restrict__ sums,
                    int64_t count, float scale) {
  const double cnt = (double)count;
  double var = (sums[1] - sums[0] * sums[0] / cnt) / (cnt - 1.0);
  if (var < 0.0) var = 0.0;
  const float mult = scale / (float)sqrt(var);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = x[i] * mult;
}

}  // namespace sonar

extern "C" {

int sonar_moments_f32(const float* x, int64_t n, double* sums, void* stream) {
  if (n <= 0) return 0;
  const int vec_ok = sonar::aligned16(x) ? 1 : 0;
  const int grid = sonar::streaming_grid((n + 3) / 4, sonar::kBlock, 2);
  sonar::moments_kernel<<<grid, sonar::kBlock, 0, (cudaStream_t)stream>>>(x, n, vec_ok, sums);
  SONAR_LAUNCH_CHECK();
  return 0;
}

int sonar_philox_normal_moments(int64_t begin, int64_t count, int64_t numel_total, uint64_t seed, uint64_t offset,
                                uint32_t grid_blocks, double* sums, void* stream) {
  if (count <= 0) return 0;
  if (grid_blocks == 0 || begin < 0 || begin + count > numel_total) return (int)cudaErrorInvalidValue;
  sonar::PhiloxStream s{seed, offset, grid_blocks * (uint32_t)sonar::kBlock};
  const int64_t T = s.threads, end = begin + count;
  const int64_t k_lo = (begin / T) / 4, k_hi = ((end - 1) / T) / 4;
  const int grid = sonar::streaming_grid(T, sonar::kBlock, 1);
  sonar::philox_normal_moments_kernel<<<grid, sonar::kBlock, 0, (cudaStream_t)stream>>>(begin, end, s, (uint32_t)k_lo,
                                                                                       (uint32_t)k_hi, sums);
  SONAR_LAUNCH_CHECK();
  return 0;
}

int sonar_philox_normal_moments_batch(const uint64_t* offsets_host, int n_draws, int64_t begin, int64_t count,
                                      int64_t numel_total, uint64_t seed, uint32_t grid_blocks, double* sums,
                                      void* stream) {
  using namespace sonar;
  if (n_draws <= 0 || count <= 0) return 0;
  if (offsets_host == nullptr || sums == nullptr || grid_blocks == 0 || begin < 0 || begin + count > numel_total)
    return (int)cudaErrorInvalidValue;
  PhiloxStream s{seed, 0, grid_blocks * (uint32_t)kBlock};
  const int64_t T = s.threads, end = begin + count;
  const int64_t k_lo = (begin / T) / 4, k_hi = ((end - 1) / T) / 4;
  SONAR_CUDA_TRY(cudaMemsetAsync(sums, 0, 2 * sizeof(double) * (size_t)n_draws, (cudaStream_t)stream));
  // 8 Philox calls per thread: one block reduction + two fp64 atomics per 2048 calls instead of per 256 (a
  // 29-draw look-ahead table over 524,288 normals each: 100 -> 70 us)
  const int gx = streaming_grid(T, kBlock * 8, 1);
  for (int first = 0; first < n_draws; first += kMomentsBatch) {
    const int m = n_draws - first < kMomentsBatch ? n_draws - first : kMomentsBatch;
    OffsetBatch offs;
    for (int i = 0; i < kMomentsBatch; ++i) offs.offset[i] = offsets_host[first + (i < m ? i : 0)];
    philox_normal_moments_batch_kernel<<<dim3((unsigned)gx, (unsigned)m), kBlock, 0, (cudaStream_t)stream>>>(
        begin, end, s, offs, (uint32_t)k_lo, (uint32_t)k_hi, sums + 2 * first);
    SONAR_LAUNCH_CHECK();
  }
  return 0;
}

int sonar_norm_decisions(const double* sums, int n, int64_t count, float threshold_std_devs, float* decisions,
                         void* stream) {
  if (n <= 0) return 0;
  if (sums == nullptr || decisions == nullptr || count <= 0 || (reinterpret_cast<uintptr_t>(decisions) & 15u))
    return (int)cudaErrorInvalidValue;
  sonar::norm_decisions_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(sums, n, count, threshold_std_devs, decisions);
  SONAR_LAUNCH_CHECK();
  return 0;
}

int sonar_philox_normal_fill_moments_f32(float* out, int64_t begin, int64_t count, int64_t numel_total, uint64_t seed,
                                         uint64_t offset, uint32_t grid_blocks, double* sums, void* stream) {
  if (count <= 0) return 0;
  if (grid_blocks == 0 || begin < 0 || begin + count > numel_total) return (int)cudaErrorInvalidValue;
  sonar::PhiloxStream s{seed, offset, grid_blocks * (uint32_t)sonar::kBlock};
  const int64_t T = s.threads, end = begin + count;
  const int64_t k_lo = (begin / T) / 4, k_hi = ((end - 1) / T) / 4;
  SONAR_CUDA_TRY(cudaMemsetAsync(sums, 0, 2 * sizeof(double), (cudaStream_t)stream));
  const int grid = sonar::streaming_grid(T, sonar::kBlock, 1);
  sonar::philox_normal_fill_moments_kernel<<<grid, sonar::kBlock, 0, (cudaStream_t)stream>>>(
      out, begin, end, s, (uint32_t)k_lo, (uint32_t)k_hi, sums);
  SONAR_LAUNCH_CHECK();
  return 0;
}

int sonar_scale_noise_f32(const float* x, float* out, int64_t n, const double* sums, int64_t count, float factor,
                          float threshold_std_devs, void* stream) {
  if (n <= 0) return 0;
  const int vec_ok = (sonar::aligned16(x) && sonar::aligned16(out)) ? 1 : 0;
  const int grid = sonar::streaming_grid((n + 3) / 4, sonar::kBlock, 2);
  sonar::scale_noise_kernel<<<grid, sonar::kBlock, 0, (cudaStream_t)stream>>>(x, out, n, vec_ok, sums, count, factor,
                                                                             threshold_std_devs, nullptr, 0, 0.0);
  SONAR_LAUNCH_CHECK();
  return 0;
}

int sonar_scale_noise_peers_f32(const float* x, float* out, int64_t n, const double* mailbox, int world, double epoch,
                                int64_t count, float factor, float threshold_std_devs, void* stream) {
  if (n <= 0) return 0;
  if (mailbox == nullptr || world < 2 || world > SONAR_PEER_MAX_RANKS) return (int)cudaErrorInvalidValue;
  const int vec_ok = (sonar::aligned16(x) && sonar::aligned16(out)) ? 1 : 0;
  const int grid = sonar::streaming_grid((n + 3) / 4, sonar::kBlock, 2);
  sonar::scale_noise_kernel<<<grid, sonar::kBlock, 0, (cudaStream_t)stream>>>(x, out, n, vec_ok, nullptr, count, factor,
                                                                             threshold_std_devs, mailbox, world, epoch);
  SONAR_LAUNCH_CHECK();
  return 0;
}

int sonar_scale_by_std_f32(const float* x, float* out, int64_t n, const double* sums, int64_t count, float scale,
                           void* stream) {
  if (n <= 0) return 0;
  const int grid = sonar::streaming_grid(n, sonar::kBlock, 4);
  sonar::scale_by_std_kernel<<<grid, sonar::kBlock, 0, (cudaStream_t)stream>>>(x, out, n, sums, count, scale);
  SONAR_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
