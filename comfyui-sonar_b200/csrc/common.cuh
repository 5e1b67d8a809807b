// Shared device helpers for the sonar_b200 kernels (sm_100a).
//
// Everything here is bandwidth-oriented plumbing: vector loads/stores, warp/block reductions,
// the torch-compatible two-branch lerp, and a random-access Philox4x32-10 front end that reproduces
// the element -> (thread, call, lane) mapping of ATen's CUDA distribution kernels so that a fused
// kernel can regenerate "torch.randn(device='cuda')" values in registers instead of reading them
// from HBM.
//
// NOTE: this translation unit family must NOT be compiled with --use_fast_math: curand's Box-Muller
// uses logf/sqrtf (accurate) and __sincosf (fast) explicitly and the bit pattern depends on it.
#pragma once

#include <cuda_runtime.h>
#include <curand_kernel.h>
#include <stdint.h>

#define SONAR_LAUNCH_CHECK()                         \
  do {                                               \
    cudaError_t err__ = cudaGetLastError();          \
    if (err__ != cudaSuccess) return (int)err__;     \
  } while (0)

#define SONAR_CUDA_TRY(expr)                         \
  do {                                               \
    cudaError_t err__ = (expr);                      \
    if (err__ != cudaSuccess) return (int)err__;     \
  } while (0)

namespace sonar {

constexpr int kBlock = 256;          // ATen's distribution block size; also our default
constexpr int kSmCountB200 = 148;

// ---------------------------------------------------------------------------------------------
// device properties (cached per device)
// ---------------------------------------------------------------------------------------------
struct DeviceInfo {
  int sm_count;
  int max_threads_per_sm;
  int max_smem_optin;
};

inline const DeviceInfo& device_info() {
  static thread_local int cached_dev = -1;
  static thread_local DeviceInfo info{kSmCountB200, 2048, 227 * 1024};
  int dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess && dev != cached_dev) {
    cudaDeviceGetAttribute(&info.sm_count, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&info.max_threads_per_sm, cudaDevAttrMaxThreadsPerMultiProcessor, dev);
    cudaDeviceGetAttribute(&info.max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    cached_dev = dev;
  }
  return info;
}

// Grid for a streaming element-wise kernel: enough CTAs to cover the work, capped at a whole number
// of waves (SMs x resident CTAs of 256 threads) so the tail wave is never ragged.
inline int streaming_grid(int64_t work_items, int items_per_block, int waves = 4) {
  const DeviceInfo& di = device_info();
  int64_t need = (work_items + items_per_block - 1) / items_per_block;
  int64_t cap = (int64_t)di.sm_count * (di.max_threads_per_sm / kBlock) * waves;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}

// Co-scheduling hint (sonar_set_grid_limit, api.cu): when > 0, the grid-stride kernels that honour it launch at
// most this many CTAs per SM, leaving the other thread slots of every SM to a kernel on another stream -- the
// ALU-bound noise producers (Philox fill, FFT) and the HBM-bound fused step then run on the same SMs at the same
// time instead of one after the other.
int& grid_limit_ctas_per_sm();

inline int streaming_grid_shared(int64_t work_items, int items_per_block, int waves = 4) {
  const int limit = grid_limit_ctas_per_sm();
  if (limit <= 0) return streaming_grid(work_items, items_per_block, waves);
  const DeviceInfo& di = device_info();
  int64_t need = (work_items + items_per_block - 1) / items_per_block;
  const int64_t cap = (int64_t)di.sm_count * limit;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}

// ---------------------------------------------------------------------------------------------
// blend functions (reference: py/utils.py:17-21 BLENDING_MODES; ATen/native/Lerp.h:21-35)
// ---------------------------------------------------------------------------------------------
enum BlendMode : int { BLEND_LERP = 0, BLEND_INJECT = 1, BLEND_SUBTRACT_B = 2 };

// Rounding discipline (parity at 1e-5 over dozens of chained steps): where the reference issues SEPARATE
// eager ops (mul, then add) every intermediate is rounded, so the kernels use __fmul_rn / __fadd_rn /
// __fsub_rn, which nvcc never contracts into an FMA; where ATen itself fuses (torch.lerp is
// fmadd(coeff, end - start, base), ATen/native/cpu/LerpKernel.cpp and Lerp.h -- checked bit-exact against
// torch.lerp on the build host) the kernels call fmaf explicitly.
__device__ __forceinline__ float torch_lerp(float a, float b, float w) {
  const float d = __fsub_rn(b, a);
  return fabsf(w) < 0.5f ? fmaf(w, d, a) : fmaf(-d, __fsub_rn(1.0f, w), b);
}

__device__ __forceinline__ double torch_lerp(double a, double b, double w) {
  const double d = __dsub_rn(b, a);
  return fabs(w) < 0.5 ? fma(w, d, a) : fma(-d, __dsub_rn(1.0, w), b);
}

__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ double sub_rn(double a, double b) { return __dsub_rn(a, b); }

template <typename T>
__device__ __forceinline__ T blend(int mode, T a, T b, T t) {
  switch (mode) {
    case BLEND_INJECT: return add_rn(mul_rn(b, t), a);      // (b * t).add_(a)
    case BLEND_SUBTRACT_B: return sub_rn(a, mul_rn(b, t));  // a - b * t
    default: return torch_lerp(a, b, t);
  }
}

// ---------------------------------------------------------------------------------------------
// reductions
// ---------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-wide sum of two doubles; result valid in thread 0. `scratch` needs 2*32 doubles.
__device__ __forceinline__ void block_sum2(double& a, double& b, double* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  a = warp_sum(a);
  b = warp_sum(b);
  if (lane == 0) {
    scratch[warp] = a;
    scratch[32 + warp] = b;
  }
  __syncthreads();
  if (warp == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    a = lane < nw ? scratch[lane] : 0.0;
    b = lane < nw ? scratch[32 + lane] : 0.0;
    a = warp_sum(a);
    b = warp_sum(b);
  }
}

// Producer kernels can reduce the moments of what they write (so scale_noise needs no separate read
// pass): per-thread fp32 partials (s, ss) -> fp64 block sum -> two atomics per block into sums[0..1],
// which must be zero on entry. `clear` (optional) is ANOTHER double[2] that the launch zeroes: callers
// recycle a ring of slots without a memset node per launch (slot i's producer clears slot i+1).
// All threads of the block must call. sums == nullptr: nothing happens.
__device__ __forceinline__ void commit_moments(double* __restrict__ sums, double* __restrict__ clear, float s, float ss) {
  if (sums == nullptr) return;  // launch-uniform
  __shared__ double moments_scratch[64];
  double ds = (double)s, dss = (double)ss;
  block_sum2(ds, dss, moments_scratch);
  if (threadIdx.x == 0) {
    atomicAdd(&sums[0], ds);
    atomicAdd(&sums[1], dss);
    if (clear != nullptr && blockIdx.x == 0 && blockIdx.y == 0) {
      clear[0] = 0.0;
      clear[1] = 0.0;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Philox, ATen-compatible random access
//
// ATen (ATen/native/cuda/DistributionTemplates.h:50-82): thread `t` of a grid of T = grid*256
// threads runs curand_init(seed, /*subsequence=*/t, /*offset=*/off) and its k-th curand_*4 call
// yields the values of elements  li = t + T*(4k + lane), lane = 0..3.
// curand_init + k calls  ==  Philox4x32-10(counter = {off/4 + k (64 bit), t (64 bit)}, key = seed).
// ---------------------------------------------------------------------------------------------
struct PhiloxStream {
  uint64_t seed;
  uint64_t offset;     // torch generator offset at the draw (multiple of 4)
  uint32_t threads;    // T = grid_blocks * 256 of the emulated ATen launch
};

// Philox4x32-10 (same rounds / constants as curand_Philox4x32_10, curand_philox4x32_x.h) written with
// one 32x32->64 multiply per lane pair per round (IMAD.WIDE) instead of separate mulhi / mullo.
// 32 x 32 -> 64 bit product as ONE IMAD.WIDE.U32 (the C++ form `(uint64_t)a * b` followed by shifts leaves an
// add of a zero high word per product in the SASS: 20 wasted instructions per Philox call).
__device__ __forceinline__ void mul_wide_u32(uint32_t a, uint32_t b, uint32_t& lo, uint32_t& hi) {
  asm("{\n\t.reg .u64 t;\n\tmul.wide.u32 t, %2, %3;\n\tmov.b64 {%0, %1}, t;\n\t}" : "=r"(lo), "=r"(hi) : "r"(a), "r"(b));
}

__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                               uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t lo0, hi0, lo1, hi1;
    mul_wide_u32(0xD2511F53u, c0, lo0, hi0);
    mul_wide_u32(0xCD9E8D57u, c2, lo1, hi1);
    const uint32_t n0 = hi1 ^ c1 ^ k0;
    const uint32_t n2 = hi0 ^ c3 ^ k1;
    c1 = lo1;
    c3 = lo0;
    c0 = n0;
    c2 = n2;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}

__device__ __forceinline__ uint4 philox_raw(const PhiloxStream& s, uint32_t thread, uint64_t call) {
  const uint64_t c = (s.offset >> 2) + call;
  return philox4x32_10((uint32_t)c, (uint32_t)(c >> 32), thread, 0u, (uint32_t)s.seed, (uint32_t)(s.seed >> 32));
}

// curand_normal4 on the k-th block of thread `thread` (curand_normal.h: curand_box_muller4)
__device__ __forceinline__ float4 philox_normal4(const PhiloxStream& s, uint32_t thread, uint64_t call) {
  const uint4 r = philox_raw(s, thread, call);
  const float2 a = _curand_box_muller(r.x, r.y);
  const float2 b = _curand_box_muller(r.z, r.w);
  return make_float4(a.x, a.y, b.x, b.y);
}

// Only the first Box-Muller pair (lanes 0, 1) -- for draws where lanes 2, 3 fall outside the tensor
__device__ __forceinline__ float2 philox_normal2_lo(const PhiloxStream& s, uint32_t thread, uint64_t call) {
  const uint4 r = philox_raw(s, thread, call);
  return _curand_box_muller(r.x, r.y);
}

// One lane of curand_normal4: lanes 0, 1 are the Box-Muller pair of (r.x, r.y), lanes 2, 3 of (r.z, r.w)
__device__ __forceinline__ float philox_normal_lane(const uint4& r, int lane) {
  const float2 n = _curand_box_muller(lane < 2 ? r.x : r.z, lane < 2 ? r.y : r.w);
  return (lane & 1) ? n.y : n.x;
}

// curand_uniform4 (values in (0, 1])
__device__ __forceinline__ float4 philox_uniform4(const PhiloxStream& s, uint32_t thread, uint64_t call) {
  return _curand_uniform4(philox_raw(s, thread, call));
}

// at::uniform_kernel's transform: rand * range + from, with the (0,1] -> [0,1) bound reversal
// (ATen/native/cuda/DistributionTemplates.h:487-501)
__device__ __forceinline__ float uniform_transform(float r, float from, float to) {
  const float range = to - from;
  const float v = r * range + from;
  return v == to ? from : v;
}

// ---------------------------------------------------------------------------------------------
// scale_noise decision (reference py/utils.py:100-106), evaluated on the device from the two
// global sums so that no .item() host sync is needed. mean/std are rounded to float first because
// the reference compares the float32 results of noise.mean()/noise.std() in Python doubles.
// ---------------------------------------------------------------------------------------------
struct NormDecision {
  float mean;
  float std;
  int sub_mean;
  int div_std;
};

__device__ __forceinline__ NormDecision decide_normalisation(const double* __restrict__ sums, int64_t count,
                                                             float threshold_std_devs) {
  NormDecision d{0.0f, 1.0f, 0, 0};
  if (sums == nullptr || count <= 0) return d;
  const double n = (double)count;
  const double s = sums[0], ss = sums[1];
  const double mean = s / n;
  double var = (ss - s * s / n) / (n - 1.0);  // unbiased; n == 1 -> NaN like torch.std
  if (var < 0.0) var = 0.0;
  d.mean = (float)mean;
  d.std = (float)sqrt(var);
  const double thr = (double)threshold_std_devs / sqrt(n);
  d.sub_mean = fabs((double)d.mean) > thr ? 1 : 0;
  d.div_std = fabs(1.0 - (double)d.std) > thr ? 1 : 0;
  return d;
}

// Bounded spin on a peer flag: a rank that died, took another code path or issued its exchange on another
// stream must not leave this GPU in a kernel that can never be killed. After kPeerWaitNs of waiting the kernel
// traps -- the context fails loudly (cudaErrorLaunchFailure at the next API call) instead of hanging.
constexpr unsigned long long kPeerWaitNs = 30ull * 1000ull * 1000ull * 1000ull;

__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ void wait_for_epoch(const volatile double* flag, double epoch) {
  if (*flag == epoch) return;
  const unsigned long long t0 = global_timer_ns();
  unsigned spins = 0;
  while (*flag != epoch) {
    if ((++spins & 0x3ffu) == 0 && global_timer_ns() - t0 > kPeerWaitNs) __trap();
  }
}

// Batch-sharded statistics over peer memory (peer.cu): wait for the `world` partial sums of `epoch`
// in the local mailbox, add them in rank order, take the same decision on every rank.
// All threads of the block must call; `shared2` is 2 doubles of shared memory.
__device__ __forceinline__ NormDecision decide_normalisation_peers(const double* mailbox, int world, double epoch,
                                                                   int64_t count, float threshold_std_devs,
                                                                   double* shared2) {
  if (threadIdx.x == 0) {
    const int parity = ((long long)epoch) & 1;
    const volatile double* slots = mailbox + (size_t)parity * 8 * 4;
    double s = 0.0, ss = 0.0;
    for (int r = 0; r < world; ++r) {
      wait_for_epoch(&slots[r * 4 + 2], epoch);
      __threadfence();
      s += slots[r * 4 + 0];
      ss += slots[r * 4 + 1];
    }
    shared2[0] = s;
    shared2[1] = ss;
  }
  __syncthreads();
  return decide_normalisation(shared2, count, threshold_std_devs);
}

// a / b for a divisor that is the same for every element: reciprocal computed once by the caller,
// one Newton step on the residual -> the correctly rounded quotient whenever a/b is a normal number
// (3 instructions instead of the ~12 of the generic IEEE division sequence).
__device__ __forceinline__ float div_by(float a, float b, float inv_b) {
  const float q = a * inv_b;
  const float r = fmaf(-q, b, a);
  return fmaf(r, inv_b, q);
}

__device__ __forceinline__ float apply_norm(float v, const NormDecision& d) {
  if (d.sub_mean) v -= d.mean;
  if (d.div_std) v /= d.std;
  return v;
}

// Block-cooperative decision: one thread does the fp64 arithmetic, everybody reads the result.
// All threads of the block must call. `slot` is a NormDecision in shared memory.
__device__ __forceinline__ NormDecision decide_normalisation_block(const double* __restrict__ sums, int64_t count,
                                                                   float threshold_std_devs, NormDecision* slot) {
  if (threadIdx.x == 0) *slot = decide_normalisation(sums, count, threshold_std_devs);
  __syncthreads();
  return *slot;
}

// ---------------------------------------------------------------------------------------------
// vector access helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, const float4& v) { *reinterpret_cast<float4*>(p) = v; }

__device__ __forceinline__ float4 ld4_stream(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}

__device__ __forceinline__ void st4_stream(float* p, const float4& v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w)
               : "memory");
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace sonar
