// Reference-latent guidance epilogue of the Sonar samplers.
//
// Reference: SonarGuidanceMixin py/sonar.py:323-411 -- guidance_shift :372-378 (the standardised
// reference latent takes the per-batch-item mean / unbiased std of x or of the denoised prediction),
// guidance_linear :400-411 (blend x towards it), guidance_euler :380-398 (Euler step towards it).
// Upstream: 2 reductions + ~6 element-wise passes per guided step; here one per-item moments launch and
// one apply launch.
#include "common.cuh"
#include "../../include/sonar_b200.h"

namespace sonar {

constexpr int kItemChunks = 64;  // CTAs per item in the moments pass

// sums[2*item], sums[2*item+1] += sum / sum of squares of this CTA's chunk of the item
__global__ void __launch_bounds__(kBlock)
item_moments_kernel(const float* __restrict__ x, int64_t per_item, int vec_ok, double* __restrict__ sums) {
  __shared__ double scratch[64];
  const int item = blockIdx.y;
  const float* base = x + (int64_t)item * per_item;
  double s = 0.0, ss = 0.0;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t done = 0;
  if (vec_ok) {
    const int64_t n4 = per_item >> 2;
    for (int64_t i = tid; i < n4; i += stride) {
      const float4 v = ld4_stream(base + 4 * i);
      s += ((double)v.x + (double)v.y) + ((double)v.z + (double)v.w);
      ss += ((double)v.x * v.x + (double)v.y * v.y) + ((double)v.z * v.z + (double)v.w * v.w);
    }
    done = n4 << 2;
  }
  for (int64_t i = done + tid; i < per_item; i += stride) {
    const double v = base[i];
    s += v;
    ss += v * v;
  }
  block_sum2(s, ss, scratch);
  if (threadIdx.x == 0) {
    atomicAdd(&sums[2 * item], s);
    atomicAdd(&sums[2 * item + 1], ss);
  }
}

template <int VEC>
__global__ void __launch_bounds__(kBlock)
guidance_kernel(SonarGuidanceParams p) {
  const int item = blockIdx.y;
  // mean and unbiased std of the statistics source for this item, rounded to float like t.mean() / t.std()
  float mean = 0.0f, stdv = 1.0f;
  if (p.item_sums != nullptr) {
    const double n = (double)p.per_item, s = p.item_sums[2 * item], ss = p.item_sums[2 * item + 1];
    double var = (ss - s * s / n) / (n - 1.0);
    if (var < 0.0) var = 0.0;
    mean = (float)(s / n);
    stdv = (float)sqrt(var);
  }
  const float* x = p.x + (int64_t)item * p.per_item;
  const float* ref = p.ref + (p.ref_items == 1 ? 0 : (int64_t)item * p.per_item);
  float* out = p.out + (int64_t)item * p.per_item;
  const float inv_sigma = 1.0f / p.sigma;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  auto one = [&](float xv, float rv) {
    const float target = p.item_sums != nullptr ? __fadd_rn(__fmul_rn(rv, stdv), mean) : rv;  // (ref * std).add_(mean)
    if (p.kind == SONAR_GUIDANCE_EULER) {
      const float d = div_by(xv - target, p.sigma, inv_sigma);  // to_d(x, sigma, target)
      return __fadd_rn(__fmul_rn(d, p.dt), xv);                 // (d * dt).add_(x)
    }
    return blend<float>(p.blend_mode, xv, target, p.factor);
  };
  int64_t done = 0;
  if (VEC == 4) {
    const int64_t n4 = p.per_item >> 2;
    for (int64_t i = tid; i < n4; i += stride) {
      const float4 xv = ld4(x + 4 * i), rv = ld4(ref + 4 * i);
      st4(out + 4 * i, make_float4(one(xv.x, rv.x), one(xv.y, rv.y), one(xv.z, rv.z), one(xv.w, rv.w)));
    }
    done = n4 << 2;
  }
  for (int64_t i = done + tid; i < p.per_item; i += stride) out[i] = one(x[i], ref[i]);
}

}  // namespace sonar

extern "C" {

int sonar_item_moments_f32(const float* x, int64_t items, int64_t per_item, double* sums, void* stream) {
  using namespace sonar;
  if (items <= 0 || per_item <= 0) return 0;
  if (x == nullptr || sums == nullptr || items > 65535) return (int)cudaErrorInvalidValue;
  SONAR_CUDA_TRY(cudaMemsetAsync(sums, 0, 2 * sizeof(double) * (size_t)items, (cudaStream_t)stream));
  const int vec_ok = (aligned16(x) && (per_item % 4 == 0)) ? 1 : 0;
  int64_t chunks = (per_item + (int64_t)kBlock * 8 - 1) / ((int64_t)kBlock * 8);
  if (chunks > kItemChunks) chunks = kItemChunks;
  item_moments_kernel<<<dim3((unsigned)chunks, (unsigned)items), kBlock, 0, (cudaStream_t)stream>>>(x, per_item, vec_ok, sums);
  SONAR_LAUNCH_CHECK();
  return 0;
}

int sonar_guidance_f32(const SonarGuidanceParams* params, void* stream) {
  using namespace sonar;
  if (params == nullptr) return (int)cudaErrorInvalidValue;
  const SonarGuidanceParams& p = *params;
  if (p.items <= 0 || p.per_item <= 0) return 0;
  if (p.x == nullptr || p.ref == nullptr || p.out == nullptr || p.items > 65535 ||
      (p.ref_items != 1 && p.ref_items != p.items) ||
      (p.kind != SONAR_GUIDANCE_LINEAR && p.kind != SONAR_GUIDANCE_EULER) || (p.kind == SONAR_GUIDANCE_EULER && p.sigma == 0.0f))
    return (int)cudaErrorInvalidValue;
  const bool vec = p.per_item % 4 == 0 && aligned16(p.x) && aligned16(p.ref) && aligned16(p.out);
  int64_t chunks = (p.per_item + (int64_t)kBlock * 4 - 1) / ((int64_t)kBlock * 4);
  const int64_t cap = (int64_t)device_info().sm_count * 8 / p.items + 1;
  if (chunks > cap) chunks = cap;
  const dim3 grid((unsigned)chunks, (unsigned)p.items);
  if (vec)
    guidance_kernel<4><<<grid, kBlock, 0, (cudaStream_t)stream>>>(p);
  else
    guidance_kernel<1><<<grid, kBlock, 0, (cudaStream_t)stream>>>(p);
  SONAR_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
