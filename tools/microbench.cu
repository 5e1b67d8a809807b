// Standalone micro-benchmarks that size the streaming kernels of sonar_b200 on a real B200.
// Not part of the product: it answers "what does the hardware give a 3-read/2-write float4 stream,
// a global reduction tail, a Philox normal draw" at the tensor sizes of BASELINE.json's configs.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu
//   gpurun -- ./tools/microbench > gpurun_out/microbench.txt
#include <cuda_runtime.h>
#include <curand_kernel.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

#define CK(x)                                                                       \
  do {                                                                              \
    cudaError_t e = (x);                                                            \
    if (e != cudaSuccess) {                                                         \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
      exit(1);                                                                      \
    }                                                                               \
  } while (0)

static float* g_flush = nullptr;
static const size_t kFlushBytes = 256u << 20;

__global__ void flush_kernel(float4* p, size_t n4) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n4; i += stride) p[i] = make_float4(1.f, 2.f, 3.f, 4.f);
}

static void flush_l2(cudaStream_t s) { flush_kernel<<<148 * 8, 256, 0, s>>>((float4*)g_flush, kFlushBytes / 16); }

// time `reps` single launches, each after an L2 flush; returns median microseconds
template <typename F>
static float time_cold(F launch, int reps = 15) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  std::vector<float> us;
  for (int r = 0; r < reps + 3; ++r) {
    flush_l2(0);
    CK(cudaEventRecord(e0, 0));
    launch();
    CK(cudaEventRecord(e1, 0));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (r >= 3) us.push_back(ms * 1e3f);
  }
  std::sort(us.begin(), us.end());
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return us[us.size() / 2];
}

// warm: `inner` back-to-back launches between events (L2-resident when the working set fits)
template <typename F>
static float time_warm(F launch, int inner = 20, int reps = 7) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  std::vector<float> us;
  for (int r = 0; r < reps + 2; ++r) {
    CK(cudaEventRecord(e0, 0));
    for (int i = 0; i < inner; ++i) launch();
    CK(cudaEventRecord(e1, 0));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (r >= 2) us.push_back(ms * 1e3f / inner);
  }
  std::sort(us.begin(), us.end());
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return us[us.size() / 2];
}

__global__ void empty_kernel() {}

// ------------------------------------------------------------------------------------------
// 3-read / 2-write stream (the shape of the fused sonar step), U float4 per thread per iteration
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 ldnc(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}
__device__ __forceinline__ void stcs(float4* p, float4 v) {
  asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

template <int U, int HINT>
__global__ void __launch_bounds__(256) stream32_kernel(const float4* __restrict__ a, const float4* __restrict__ b,
                                                       const float4* __restrict__ c, float4* __restrict__ o1,
                                                       float4* __restrict__ o2, int64_t n4) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + (U - 1) * stride < n4; i += U * stride) {
    float4 va[U], vb[U], vc[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (HINT) {
        va[u] = ldnc(a + i + u * stride);
        vb[u] = ldnc(b + i + u * stride);
        vc[u] = ldnc(c + i + u * stride);
      } else {
        va[u] = a[i + u * stride];
        vb[u] = b[i + u * stride];
        vc[u] = c[i + u * stride];
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float4 r1 = make_float4(va[u].x * 0.5f + vb[u].x, va[u].y * 0.5f + vb[u].y, va[u].z * 0.5f + vb[u].z,
                              va[u].w * 0.5f + vb[u].w);
      float4 r2 = make_float4(vc[u].x * 0.25f + r1.x, vc[u].y * 0.25f + r1.y, vc[u].z * 0.25f + r1.z,
                              vc[u].w * 0.25f + r1.w);
      if (HINT) {
        stcs(o1 + i + u * stride, r1);
        stcs(o2 + i + u * stride, r2);
      } else {
        o1[i + u * stride] = r1;
        o2[i + u * stride] = r2;
      }
    }
  }
  for (; i < n4; i += stride) {
    float4 va = a[i], vb = b[i], vc = c[i];
    float4 r1 = make_float4(va.x * 0.5f + vb.x, va.y * 0.5f + vb.y, va.z * 0.5f + vb.z, va.w * 0.5f + vb.w);
    o1[i] = r1;
    o2[i] = make_float4(vc.x * 0.25f + r1.x, vc.y * 0.25f + r1.y, vc.z * 0.25f + r1.z, vc.w * 0.25f + r1.w);
  }
}

// ------------------------------------------------------------------------------------------
// reductions: sum + sum of squares of n floats, three tails
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum_d(double v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_sum_f(float v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// TAIL 0: two double atomics per block on one address pair; 1: per-block partials, no combine;
// 2: per-block partials + last-block-done combine (threadfence + counter)
template <int TAIL, int U>
__global__ void __launch_bounds__(256) moments_kernel(const float4* __restrict__ x, int64_t n4, double* sums,
                                                      double* partials, unsigned* counter) {
  __shared__ double sh[16];
  __shared__ bool last;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  float s = 0.f, ss = 0.f;
  for (; i + (U - 1) * stride < n4; i += U * stride) {
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = ldnc(x + i + u * stride);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      s += (v[u].x + v[u].y) + (v[u].z + v[u].w);
      ss += (v[u].x * v[u].x + v[u].y * v[u].y) + (v[u].z * v[u].z + v[u].w * v[u].w);
    }
  }
  for (; i < n4; i += stride) {
    float4 v = ldnc(x + i);
    s += (v.x + v.y) + (v.z + v.w);
    ss += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
  }
  double ds = warp_sum_d((double)s), dss = warp_sum_d((double)ss);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) {
    sh[warp] = ds;
    sh[8 + warp] = dss;
  }
  __syncthreads();
  if (warp == 0) {
    ds = lane < 8 ? sh[lane] : 0.0;
    dss = lane < 8 ? sh[8 + lane] : 0.0;
    ds = warp_sum_d(ds);
    dss = warp_sum_d(dss);
    if (lane == 0) {
      if (TAIL == 0) {
        atomicAdd(&sums[0], ds);
        atomicAdd(&sums[1], dss);
      } else {
        partials[2 * blockIdx.x] = ds;
        partials[2 * blockIdx.x + 1] = dss;
      }
    }
  }
  if (TAIL == 2) {
    if (threadIdx.x == 0) {
      __threadfence();
      unsigned t = atomicAdd(counter, 1u);
      last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (last) {
      __threadfence();
      double a = 0.0, b = 0.0;
      for (int j = threadIdx.x; j < (int)gridDim.x; j += blockDim.x) {
        a += ((volatile double*)partials)[2 * j];
        b += ((volatile double*)partials)[2 * j + 1];
      }
      a = warp_sum_d(a);
      b = warp_sum_d(b);
      __syncthreads();
      if (lane == 0) {
        sh[warp] = a;
        sh[8 + warp] = b;
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        double ta = 0, tb = 0;
        for (int w = 0; w < 8; ++w) {
          ta += sh[w];
          tb += sh[8 + w];
        }
        sums[0] = ta;
        sums[1] = tb;
        *counter = 0;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// Philox normal draws (ATen mapping), curand's device functions vs a hand-scheduled version
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) philox_curand_kernel(float* out, int64_t n, uint64_t seed, uint64_t offset) {
  const int64_t T = (int64_t)gridDim.x * blockDim.x;
  const int64_t vt = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
  for (int64_t k = 0; vt + T * 4 * k < n; ++k) {
    const uint64_t c = (offset >> 2) + k;
    uint4 ctr = make_uint4((uint32_t)c, (uint32_t)(c >> 32), (uint32_t)vt, 0u);
    uint4 r = curand_Philox4x32_10(ctr, key);
    float2 a = _curand_box_muller(r.x, r.y);
    float2 b = _curand_box_muller(r.z, r.w);
    const int64_t li = vt + T * 4 * k;
    out[li] = a.x;
    if (li + T < n) out[li + T] = a.y;
    if (li + 2 * T < n) out[li + 2 * T] = b.x;
    if (li + 3 * T < n) out[li + 3 * T] = b.y;
  }
}

__device__ __forceinline__ uint4 philox10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    c1 = (uint32_t)p1;
    c3 = (uint32_t)p0;
    c0 = n0;
    c2 = n2;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}

// moments only (no store): the ALU ceiling of "regenerate the draw in registers"
__global__ void __launch_bounds__(256) philox_moments_only_kernel(double* sums, int64_t n, uint64_t seed, uint64_t offset) {
  __shared__ double sh[16];
  const int64_t T = (int64_t)gridDim.x * blockDim.x;
  const int64_t vt = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  float s = 0.f, ss = 0.f;
  for (int64_t k = 0; vt + T * 4 * k < n; ++k) {
    const uint64_t c = (offset >> 2) + k;
    uint4 r = philox10((uint32_t)c, (uint32_t)(c >> 32), (uint32_t)vt, 0u, (uint32_t)seed, (uint32_t)(seed >> 32));
    float2 a = _curand_box_muller(r.x, r.y);
    float2 b = _curand_box_muller(r.z, r.w);
    const int64_t li = vt + T * 4 * k;
    s += a.x;
    ss += a.x * a.x;
    if (li + T < n) { s += a.y; ss += a.y * a.y; }
    if (li + 2 * T < n) { s += b.x; ss += b.x * b.x; }
    if (li + 3 * T < n) { s += b.y; ss += b.y * b.y; }
  }
  double ds = warp_sum_d((double)s), dss = warp_sum_d((double)ss);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { sh[warp] = ds; sh[8 + warp] = dss; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0, b = 0;
    for (int w = 0; w < 8; ++w) { a += sh[w]; b += sh[8 + w]; }
    atomicAdd(&sums[0], a);
    atomicAdd(&sums[1], b);
  }
}

// ------------------------------------------------------------------------------------------
int main() {
  CK(cudaSetDevice(0));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  printf("device %s, %d SMs, L2 %d MB\n", prop.name, prop.multiProcessorCount, prop.l2CacheSize >> 20);
  CK(cudaMalloc(&g_flush, kFlushBytes));

  printf("empty kernel, events around one launch (cold): %.2f us; back-to-back: %.2f us\n",
         time_cold([] { empty_kernel<<<1, 32>>>(); }), time_warm([] { empty_kernel<<<1, 32>>>(); }, 50));
  printf("empty kernel 1184x256: cold %.2f us; b2b %.2f us\n", time_cold([] { empty_kernel<<<1184, 256>>>(); }),
         time_warm([] { empty_kernel<<<1184, 256>>>(); }, 50));

  const int64_t sizes[] = {524288, 4194304, 7603200, 60825600};
  for (int64_t n : sizes) {
    float *a, *b, *c, *o1, *o2;
    CK(cudaMalloc(&a, n * 4)); CK(cudaMalloc(&b, n * 4)); CK(cudaMalloc(&c, n * 4));
    CK(cudaMalloc(&o1, n * 4)); CK(cudaMalloc(&o2, n * 4));
    CK(cudaMemset(a, 0, n * 4)); CK(cudaMemset(b, 0, n * 4)); CK(cudaMemset(c, 0, n * 4));
    const int64_t n4 = n / 4;
    const double bytes = 20.0 * n;
    printf("\n== stream 3R/2W, n = %lld floats (%.1f MB moved) ==\n", (long long)n, bytes / 1e6);
    const int grids[] = {148, 296, 592, 1184, 2368, 0};
    for (int g : grids) {
      int grid = g == 0 ? (int)((n4 + 255) / 256) : g;
      if (g != 0 && (int64_t)g * 256 > n4) continue;
#define RUN(U, H)                                                                                                   \
  {                                                                                                                 \
    float tc = time_cold([&] { stream32_kernel<U, H><<<grid, 256>>>((float4*)a, (float4*)b, (float4*)c, (float4*)o1, (float4*)o2, n4); }); \
    float tw = time_warm([&] { stream32_kernel<U, H><<<grid, 256>>>((float4*)a, (float4*)b, (float4*)c, (float4*)o1, (float4*)o2, n4); }); \
    printf("  grid %6d U=%d hint=%d : cold %7.2f us (%6.0f GB/s)   warm %7.2f us (%6.0f GB/s)\n", grid, U, H, tc,      \
           bytes / tc / 1e3, tw, bytes / tw / 1e3);                                                                 \
  }
      RUN(1, 0) RUN(1, 1) RUN(2, 1) RUN(4, 1)
#undef RUN
    }
    // reductions
    double* sums; double* partials; unsigned* counter;
    CK(cudaMalloc(&sums, 16)); CK(cudaMalloc(&partials, 16 * 8192)); CK(cudaMalloc(&counter, 4));
    CK(cudaMemset(sums, 0, 16)); CK(cudaMemset(counter, 0, 4));
    printf("-- moments of n floats (%.1f MB read)\n", 4.0 * n / 1e6);
    for (int g : {148, 296, 592, 1184, 2368}) {
      if ((int64_t)g * 256 > n4) continue;
#define RUNM(T, U)                                                                                              \
  {                                                                                                             \
    float tc = time_cold([&] { moments_kernel<T, U><<<g, 256>>>((float4*)a, n4, sums, partials, counter); });    \
    printf("  grid %5d tail=%d U=%d: cold %7.2f us (%6.0f GB/s)\n", g, T, U, tc, 4.0 * n / tc / 1e3);            \
  }
      RUNM(0, 4) RUNM(1, 4) RUNM(2, 4) RUNM(2, 8)
#undef RUNM
    }
    // philox
    printf("-- philox normal fill of n floats\n");
    for (int g : {296, 592, 1184}) {
      float tc = time_cold([&] { philox_curand_kernel<<<g, 256>>>(o1, n, 1234, 0); });
      float tm = time_cold([&] { philox_moments_only_kernel<<<g, 256>>>(sums, n, 1234, 0); });
      printf("  grid %5d: fill cold %7.2f us (%6.1f Gnormal/s)   moments-only %7.2f us (%6.1f Gnormal/s)\n", g, tc,
             n / tc / 1e3, tm, n / tm / 1e3);
    }
    cudaFree(a); cudaFree(b); cudaFree(c); cudaFree(o1); cudaFree(o2); cudaFree(sums); cudaFree(partials); cudaFree(counter);
  }
  return 0;
}
