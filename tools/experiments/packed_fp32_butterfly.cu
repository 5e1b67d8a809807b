// Round-2 experiment (not part of libsonar_b200.so): is the packed two-wide fp32 arithmetic of sm_100
// (add.f32x2 / fma.rn.f32x2, FADD2 / FFMA2 in SASS) worth using for the FFT butterflies of csrc/spectral.cu?
// Complex values are two-wide by nature: a complex add is one FADD2, a complex multiply by a twiddle is
// FMUL2 + FFMA2 plus one swizzle. The spectral kernel is issue bound (profiles/r01i), so halving the number of
// floating-point instructions of the butterflies (35 % of what it issues) is the cheapest lever left.
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o packed_fp32_butterfly packed_fp32_butterfly.cu
//   ./packed_fp32_butterfly            -> ns per radix-5 butterfly + twiddles, scalar vs packed, and max |diff|
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul_conj(float2 a, float2 w) {
  return make_float2(a.x * w.x + a.y * w.y, a.y * w.x - a.x * w.y);
}

// packed forms
__device__ __forceinline__ float2 padd(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 psub(float2 a, float2 b) { return __ffma2_rn(b, make_float2(-1.0f, -1.0f), a); }
__device__ __forceinline__ float2 pscale(float2 a, float s) { return __fmul2_rn(a, make_float2(s, s)); }
__device__ __forceinline__ float2 pfma(float2 a, float s, float2 c) { return __ffma2_rn(a, make_float2(s, s), c); }
// a * conj(w) = (a.x w.x + a.y w.y, a.y w.x - a.x w.y) = a * (w.x, w.x) + (a.y, -a.x) * (w.y, w.y)
__device__ __forceinline__ float2 pmul_conj(float2 a, float2 w) {
  return __ffma2_rn(make_float2(a.y, -a.x), make_float2(w.y, w.y), __fmul2_rn(a, make_float2(w.x, w.x)));
}

constexpr float c1 = 0.30901699437494742f, c2 = -0.80901699437494742f;
constexpr float s1 = 0.95105651629515357f, s2 = 0.58778525229247313f;

__device__ __forceinline__ void radix5_scalar(float2 (&v)[5], const float2 (&w)[5]) {
#pragma unroll
  for (int t = 1; t < 5; ++t) v[t] = cmul_conj(v[t], w[t]);
  const float2 v0 = v[0];
  const float2 a1 = cadd(v[1], v[4]), a2 = cadd(v[2], v[3]), b1 = csub(v[1], v[4]), b2 = csub(v[2], v[3]);
  const float2 m1 = make_float2(v0.x + c1 * a1.x + c2 * a2.x, v0.y + c1 * a1.y + c2 * a2.y);
  const float2 m2 = make_float2(v0.x + c2 * a1.x + c1 * a2.x, v0.y + c2 * a1.y + c1 * a2.y);
  const float2 e1 = make_float2(s1 * b1.x + s2 * b2.x, s1 * b1.y + s2 * b2.y);
  const float2 e2 = make_float2(s2 * b1.x - s1 * b2.x, s2 * b1.y - s1 * b2.y);
  const float2 n1 = make_float2(-e1.y, e1.x), n2 = make_float2(-e2.y, e2.x);
  v[0] = make_float2(v0.x + a1.x + a2.x, v0.y + a1.y + a2.y);
  v[1] = cadd(m1, n1);
  v[2] = cadd(m2, n2);
  v[3] = csub(m2, n2);
  v[4] = csub(m1, n1);
}

__device__ __forceinline__ void radix5_packed(float2 (&v)[5], const float2 (&w)[5]) {
#pragma unroll
  for (int t = 1; t < 5; ++t) v[t] = pmul_conj(v[t], w[t]);
  const float2 v0 = v[0];
  const float2 a1 = padd(v[1], v[4]), a2 = padd(v[2], v[3]), b1 = psub(v[1], v[4]), b2 = psub(v[2], v[3]);
  const float2 m1 = pfma(a2, c2, pfma(a1, c1, v0));
  const float2 m2 = pfma(a2, c1, pfma(a1, c2, v0));
  const float2 e1 = pfma(b2, s2, pscale(b1, s1));
  const float2 e2 = pfma(b2, -s1, pscale(b1, s2));
  const float2 n1 = make_float2(-e1.y, e1.x), n2 = make_float2(-e2.y, e2.x);
  v[0] = padd(v0, padd(a1, a2));
  v[1] = padd(m1, n1);
  v[2] = padd(m2, n2);
  v[3] = psub(m2, n2);
  v[4] = psub(m1, n1);
}

template <bool PACKED>
__global__ void __launch_bounds__(256) butterfly_loop(float2* out, int iters) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  float2 v[5], w[5];
#pragma unroll
  for (int t = 0; t < 5; ++t) {
    v[t] = make_float2(0.001f * (tid % 97) + 0.1f * t, 0.002f * (tid % 89) - 0.1f * t);
    float s, c;
    __sincosf(0.01f * (t + 1) * (tid % 31), &s, &c);
    w[t] = make_float2(c, s);
  }
  for (int i = 0; i < iters; ++i) {
    if (PACKED)
      radix5_packed(v, w);
    else
      radix5_scalar(v, w);
#pragma unroll
    for (int t = 0; t < 5; ++t) v[t] = make_float2(v[t].x * 0.2f, v[t].y * 0.2f);  // keep the values bounded
  }
#pragma unroll
  for (int t = 0; t < 5; ++t) out[tid * 5 + t] = v[t];
}

int main() {
  const int blocks = 148 * 8, threads = 256, iters = 2000;
  float2 *a, *b;
  cudaMalloc(&a, sizeof(float2) * 5 * blocks * threads);
  cudaMalloc(&b, sizeof(float2) * 5 * blocks * threads);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float ms[2];
  for (int rep = 0; rep < 2; ++rep) {  // first pass warms up
    cudaEventRecord(e0);
    butterfly_loop<false><<<blocks, threads>>>(a, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms[0], e0, e1);
    cudaEventRecord(e0);
    butterfly_loop<true><<<blocks, threads>>>(b, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms[1], e0, e1);
  }
  const size_t n = (size_t)5 * blocks * threads;
  float2* ha = new float2[n];
  float2* hb = new float2[n];
  cudaMemcpy(ha, a, n * sizeof(float2), cudaMemcpyDeviceToHost);
  cudaMemcpy(hb, b, n * sizeof(float2), cudaMemcpyDeviceToHost);
  double diff = 0;
  for (size_t i = 0; i < n; ++i) {
    diff = fmax(diff, fabs((double)ha[i].x - hb[i].x));
    diff = fmax(diff, fabs((double)ha[i].y - hb[i].y));
  }
  const double butterflies = (double)blocks * threads * iters;
  printf("radix-5 butterfly + 4 twiddles: scalar %.3f ms (%.3f ps each), packed %.3f ms (%.3f ps each), max |diff| %.3g\n",
         ms[0], ms[0] * 1e9 / butterflies, ms[1], ms[1] * 1e9 / butterflies, diff);
  printf("cudaGetLastError: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
