// FreeU-Extreme epilogue around the spectral filter.
//
// Reference: FreeUExtremeConfig.get_scale py/nodes/freeu_extreme.py:183-194 ("hidden mean": per batch
// item, the channel mean of the activation rescaled to [0, 1] by its own min / max, then
// 1 + (scale - 1) * that) and FreeUExtremeConfig.apply :203-227 (filtered slice * scale, written back
// over the channel slice, optionally through a BLENDING_MODES function). The filter itself (ffilter
// :10-29) is sonar_spectral_filter_f32 with real input.
//
// Two launches per application: (1) channel mean + per-block (min, max) partials, (2) one pass over the
// slice that reduces the partials in its prologue and applies scale and blend while writing the
// activation in place. Both are pure streams: 4 B/element read for (1), 8-12 B/element for (2).
#include "common.cuh"
#include "../../include/sonar_b200.h"

namespace sonar {

constexpr int kFreeuRangeBlocks = 64;  // (min, max) partials per batch item

// hidden[b][p] = mean_c h[b][c][p]; partial[b][blockIdx.x] = (min, max) of this block's pixels.
template <int VEC>
__global__ void __launch_bounds__(kBlock)
freeu_hidden_mean_kernel(const float* __restrict__ h, float* __restrict__ hidden, float2* __restrict__ partial,
                         int channels, int64_t hw) {
  const int b = blockIdx.y;
  const float* src = h + (int64_t)b * channels * hw;
  float* dst = hidden + (int64_t)b * hw;
  const float count = (float)channels;  // torch.mean: sum / count
  float mn = INFINITY, mx = -INFINITY;
  const int64_t nvec = hw / VEC;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
    float acc[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) acc[v] = 0.0f;
    const float* p = src + i * VEC;
    for (int c = 0; c < channels; ++c, p += hw) {
      if (VEC == 4) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(p));
        acc[0] += q.x;
        acc[1 % VEC] += q.y;
        acc[2 % VEC] += q.z;
        acc[3 % VEC] += q.w;
      } else {
        acc[0] += __ldg(p);
      }
    }
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      acc[v] = acc[v] / count;
      mn = fminf(mn, acc[v]);
      mx = fmaxf(mx, acc[v]);
    }
    if (VEC == 4)
      *reinterpret_cast<float4*>(dst + i * VEC) = make_float4(acc[0], acc[1 % VEC], acc[2 % VEC], acc[3 % VEC]);
    else
      dst[i] = acc[0];
  }
  __shared__ float smin[kBlock / 32], smax[kBlock / 32];
  mn = warp_min(mn);
  mx = warp_max(mx);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) {
    smin[warp] = mn;
    smax[warp] = mx;
  }
  __syncthreads();
  if (warp == 0) {
    mn = lane < kBlock / 32 ? smin[lane] : INFINITY;
    mx = lane < kBlock / 32 ? smax[lane] : -INFINITY;
    mn = warp_min(mn);
    mx = warp_max(mx);
    if (lane == 0) partial[(int64_t)b * kFreeuRangeBlocks + blockIdx.x] = make_float2(mn, mx);
  }
  if (blockIdx.x == 0)  // neutral entries for the partial slots no block owns
    for (int i = gridDim.x + threadIdx.x; i < kFreeuRangeBlocks; i += blockDim.x)
      partial[(int64_t)b * kFreeuRangeBlocks + i] = make_float2(INFINITY, -INFINITY);
}

// x[b][off + c][p] = blend(x, src * sc, t), src = filtered[b][c][p] (or x itself), sc = scale or
// 1 + (scale - 1) * (hidden[b][p] - min_b) / (max_b - min_b). One block row per (b, c) plane of the slice.
template <int VEC>
__global__ void __launch_bounds__(kBlock)
freeu_apply_kernel(SonarFreeuParams P) {
  __shared__ float2 rng_s;
  const int64_t planes = P.batch * P.slice_channels;
  for (int64_t plane = blockIdx.y; plane < planes; plane += gridDim.y) {
    const int64_t b = plane / P.slice_channels, c = plane - b * P.slice_channels;
    float mn = 0.0f, span = 1.0f;
    if (P.hidden != nullptr) {
      __syncthreads();
      if (threadIdx.x < 32) {
        float lo = INFINITY, hi = -INFINITY;
        const float2* part = reinterpret_cast<const float2*>(P.hidden_range) + b * kFreeuRangeBlocks;
        for (int i = threadIdx.x; i < kFreeuRangeBlocks; i += 32) {
          const float2 v = part[i];
          lo = fminf(lo, v.x);
          hi = fmaxf(hi, v.y);
        }
        lo = warp_min(lo);
        hi = warp_max(hi);
        if (threadIdx.x == 0) rng_s = make_float2(lo, hi);
      }
      __syncthreads();
      mn = rng_s.x;
      span = __fsub_rn(rng_s.y, rng_s.x);
    }
    float* xp = P.x + (b * P.channels + P.slice_offset + c) * P.hw;
    const float* fp = P.filtered != nullptr ? P.filtered + plane * P.hw : xp;
    const float* hp = P.hidden != nullptr ? P.hidden + b * P.hw : nullptr;
    const int64_t nvec = P.hw / VEC;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
      float xv[VEC], fv[VEC], hv[VEC];
      if (VEC == 4) {
        const float4 q = *reinterpret_cast<const float4*>(xp + i * 4);
        xv[0] = q.x, xv[1 % VEC] = q.y, xv[2 % VEC] = q.z, xv[3 % VEC] = q.w;
        const float4 f = *reinterpret_cast<const float4*>(fp + i * 4);
        fv[0] = f.x, fv[1 % VEC] = f.y, fv[2 % VEC] = f.z, fv[3 % VEC] = f.w;
        if (hp != nullptr) {
          const float4 g = __ldg(reinterpret_cast<const float4*>(hp + i * 4));
          hv[0] = g.x, hv[1 % VEC] = g.y, hv[2 % VEC] = g.z, hv[3 % VEC] = g.w;
        }
      } else {
        xv[0] = xp[i];
        fv[0] = fp[i];
        if (hp != nullptr) hv[0] = __ldg(hp + i);
      }
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        float sc = P.scale;
        if (hp != nullptr) {
          const float t = __fdiv_rn(__fsub_rn(hv[v], mn), span);
          sc = __fadd_rn(1.0f, __fmul_rn(P.scale_minus_one, t));
        }
        const float y = __fmul_rn(fv[v], sc);
        xv[v] = P.use_blend ? blend<float>(P.blend_mode, xv[v], y, P.blend) : y;
      }
      if (VEC == 4)
        *reinterpret_cast<float4*>(xp + i * 4) = make_float4(xv[0], xv[1 % VEC], xv[2 % VEC], xv[3 % VEC]);
      else
        xp[i] = xv[0];
    }
  }
}

static int freeu_range_blocks(int64_t hw, int vec) {
  int64_t blocks = (hw / vec + kBlock - 1) / kBlock;
  if (blocks < 1) blocks = 1;
  return (int)(blocks > kFreeuRangeBlocks ? kFreeuRangeBlocks : blocks);
}

}  // namespace sonar

extern "C" {

int64_t sonar_freeu_range_bytes(int64_t batch) { return batch * sonar::kFreeuRangeBlocks * (int64_t)sizeof(float2); }

int sonar_freeu_hidden_mean_f32(const float* h, float* hidden, void* range_partial, int64_t batch, int64_t channels,
                                int64_t hw, void* stream) {
  using namespace sonar;
  if (batch <= 0 || channels <= 0 || hw <= 0) return 0;
  if (batch > 65535 || channels > INT32_MAX || h == nullptr || hidden == nullptr || range_partial == nullptr)
    return (int)cudaErrorInvalidValue;
  const bool vec = (hw % 4 == 0) && aligned16(h) && aligned16(hidden);
  const int blocks = freeu_range_blocks(hw, vec ? 4 : 1);
  const dim3 grid((unsigned)blocks, (unsigned)batch);
  if (vec)
    freeu_hidden_mean_kernel<4><<<grid, kBlock, 0, (cudaStream_t)stream>>>(h, hidden, (float2*)range_partial, (int)channels, hw);
  else
    freeu_hidden_mean_kernel<1><<<grid, kBlock, 0, (cudaStream_t)stream>>>(h, hidden, (float2*)range_partial, (int)channels, hw);
  SONAR_LAUNCH_CHECK();
  return 0;
}

int sonar_freeu_apply_f32(const SonarFreeuParams* params, void* stream) {
  using namespace sonar;
  if (params == nullptr) return (int)cudaErrorInvalidValue;
  const SonarFreeuParams& p = *params;
  if (p.batch <= 0 || p.slice_channels <= 0 || p.hw <= 0) return 0;
  if (p.x == nullptr || p.slice_offset < 0 || p.slice_offset + p.slice_channels > p.channels) return (int)cudaErrorInvalidValue;
  if ((p.hidden == nullptr) != (p.hidden_range == nullptr)) return (int)cudaErrorInvalidValue;
  const bool vec = (p.hw % 4 == 0) && aligned16(p.x) && (p.filtered == nullptr || aligned16(p.filtered)) &&
                   (p.hidden == nullptr || aligned16(p.hidden));
  const int64_t planes = p.batch * p.slice_channels;
  int64_t gx = (p.hw / (vec ? 4 : 1) + kBlock - 1) / kBlock;
  if (gx < 1) gx = 1;
  if (gx > 64) gx = 64;
  const dim3 grid((unsigned)gx, (unsigned)(planes > 65535 ? 65535 : planes));
  if (vec)
    freeu_apply_kernel<4><<<grid, kBlock, 0, (cudaStream_t)stream>>>(p);
  else
    freeu_apply_kernel<1><<<grid, kBlock, 0, (cudaStream_t)stream>>>(p);
  SONAR_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
