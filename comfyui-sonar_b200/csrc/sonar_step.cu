// Fused Sonar momentum step: momentum mix + both history updates + Euler / DPM-Solver++(SDE) half
// step + ancestral noise injection, one read of (x, denoised, history[, noise]) and one write of
// (x', history') per element.
//
// Reference (py/sonar.py): update_hist :227-236, momentum_mix :238-260, get_momentum_denoised
// :262-283, get_momentum_d :285-307, momentum_step :309-320, SonarEulerAncestral.step :541-573,
// SonarDPMPPSDE.momentum_step :649-735. The reference issues ~15 full-tensor ATen passes per Euler-a
// step (46 ops) and 4 host syncs; this is one launch and no sync.
//
// The ancestral noise can be (a) absent, (b) a tensor, (c) regenerated in registers from the
// Philox stream torch.randn(device='cuda') would have produced, optionally with the conditional
// global normalisation of scale_noise applied from device-resident sums (stats pre-pass in
// stats.cu: zero HBM bytes for the noise).
#include <cooperative_groups.h>
#include <stdlib.h>

#include "common.cuh"
#include "../../include/sonar_b200.h"

namespace sonar {

struct StepElem {
  float x_out;
  float h_out;
};

__device__ __forceinline__ float update_history(const SonarStepParams& p, bool have_h, float h, float v) {
  // update_hist: history = v if history is None else blend(v * md_scale, h * hd_scale, hd_ratio)
  return have_h ? blend<float>(p.history_blend, v * p.md_scale, h * p.hd_scale, p.hd_ratio) : v;
}

// Per-launch constants derived once per thread from the parameter block.
struct StepConsts {
  float inv_sigma;
  float inv_hist_div;
  bool m_is_one, denoised_mode, have_h0, have_h_first_mix, scale_hist;
};

__device__ __forceinline__ StepConsts make_consts(const SonarStepParams& p) {
  StepConsts c;
  c.inv_sigma = 1.0f / p.sigma;
  c.inv_hist_div = 1.0f / p.hist_in_div;
  c.m_is_one = p.momentum == 1.0f;
  c.denoised_mode = p.mode == SONAR_MODE_DENOISED;
  c.have_h0 = p.hist_state != SONAR_HIST_NONE;
  c.have_h_first_mix = p.hist_state == SONAR_HIST_PRESENT;
  c.scale_hist = p.hist_in_div != 1.0f;
  return c;
}

__device__ __forceinline__ StepElem step_element(const SonarStepParams& p, const StepConsts& c, float x, float den,
                                                 float h_raw, float noise) {
  bool have_h = c.have_h0;
  float h = have_h ? (c.scale_hist ? div_by(h_raw, p.hist_in_div, c.inv_hist_div) : h_raw) : 0.0f;

  // ---- get_momentum_denoised ----
  float md = den;
  if (!c.m_is_one && c.have_h_first_mix && c.denoised_mode) md = blend<float>(p.momentum_blend, h * p.sigma, den, p.momentum);
  if (p.history_active) {
    h = update_history(p, have_h, h, div_by(den, p.sigma, c.inv_sigma));
    have_h = true;
  }
  const float den_eff = p.momentum_active ? md : den;

  // ---- derivative / DPM-Solver++ difference term ----
  const float d = p.kind == SONAR_STEP_EULER ? div_by(x - den_eff, p.sigma, c.inv_sigma) : p.c0 * den_eff;

  // ---- get_momentum_d ----
  float d_out = d;
  if (!c.m_is_one && !c.denoised_mode) {
    const float mom_d = have_h ? blend<float>(p.momentum_blend, h, d, p.momentum) : d;
    if (p.history_active) {
      h = update_history(p, have_h, h, p.mode == SONAR_MODE_NEW ? d : mom_d);
      have_h = true;
    }
    d_out = p.momentum_active ? mom_d : d;
  }

  StepElem r;
  r.x_out = p.kind == SONAR_STEP_EULER ? d_out * p.c0 + x : p.c1 * x - d_out;
  if (p.noise_kind != SONAR_NOISE_NONE) r.x_out = r.x_out + noise * p.noise_scale;
  r.h_out = h;
  return r;
}

// scale_noise on load: (v - mean) / std * factor with the division done by reciprocal + Newton step
struct NoiseNorm {
  float mean, std, inv_std, factor;
  bool sub_mean, div_std;
};

__device__ __forceinline__ NoiseNorm make_noise_norm(const NormDecision& d, float factor) {
  NoiseNorm n;
  n.mean = d.mean;
  n.std = d.std;
  n.inv_std = 1.0f / d.std;
  n.factor = factor;
  n.sub_mean = d.sub_mean != 0;
  n.div_std = d.div_std != 0;
  return n;
}

__device__ __forceinline__ float norm_noise_value(float v, const NoiseNorm& n) {
  if (n.sub_mean) v -= n.mean;
  if (n.div_std) v = div_by(v, n.std, n.inv_std);
  return v * n.factor;
}

// ---- contiguous float4 variant (no noise / tensor noise) ----
__device__ __forceinline__ NormDecision tensor_noise_decision(const SonarStepParams& p, bool norm_noise,
                                                              NormDecision* slot, double* peer_sums) {
  if (!norm_noise) return NormDecision{0.f, 1.f, 0, 0};
  if (p.peer_world > 1)
    return decide_normalisation_peers(p.peer_mailbox, p.peer_world, p.peer_epoch, p.noise_count,
                                      p.noise_threshold_std_devs, peer_sums);
  return decide_normalisation_block(p.noise_sums, p.noise_count, p.noise_threshold_std_devs, slot);
}

__global__ void __launch_bounds__(kBlock)
sonar_step_vec_kernel(SonarStepParams p) {
  __shared__ NormDecision nd_slot;
  __shared__ double peer_sums[2];
  const bool has_h_in = p.hist_state != SONAR_HIST_NONE;
  const bool has_noise = p.noise_kind == SONAR_NOISE_TENSOR || p.noise_kind == SONAR_NOISE_TENSOR_NORMALIZED;
  const bool write_h = p.hist_out != nullptr;
  // raw Gaussian tensor + device-resident sums: apply scale_noise on load
  const bool norm_noise = p.noise_kind == SONAR_NOISE_TENSOR_NORMALIZED;
  const NoiseNorm nn = make_noise_norm(tensor_noise_decision(p, norm_noise, &nd_slot, peer_sums),
                                       norm_noise ? p.noise_factor : 1.0f);
  const StepConsts c = make_consts(p);
  const int64_t n4 = p.n >> 2;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = tid; i < n4; i += stride) {
    const float4 x = ld4_stream(p.x + 4 * i);
    const float4 dn = ld4_stream(p.denoised + 4 * i);
    const float4 h = has_h_in ? ld4(p.hist_in + 4 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 nz = has_noise ? ld4_stream(p.noise + 4 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
    if (norm_noise) {
      nz.x = norm_noise_value(nz.x, nn);
      nz.y = norm_noise_value(nz.y, nn);
      nz.z = norm_noise_value(nz.z, nn);
      nz.w = norm_noise_value(nz.w, nn);
    }
    const StepElem a = step_element(p, c, x.x, dn.x, h.x, nz.x);
    const StepElem b = step_element(p, c, x.y, dn.y, h.y, nz.y);
    const StepElem e = step_element(p, c, x.z, dn.z, h.z, nz.z);
    const StepElem d = step_element(p, c, x.w, dn.w, h.w, nz.w);
    st4(p.x_out + 4 * i, make_float4(a.x_out, b.x_out, e.x_out, d.x_out));
    if (write_h) st4(p.hist_out + 4 * i, make_float4(a.h_out, b.h_out, e.h_out, d.h_out));
  }
  for (int64_t i = (n4 << 2) + tid; i < p.n; i += stride) {
    const float nzs = has_noise ? (norm_noise ? norm_noise_value(p.noise[i], nn) : p.noise[i]) : 0.f;
    const StepElem a = step_element(p, c, p.x[i], p.denoised[i], has_h_in ? p.hist_in[i] : 0.f, nzs);
    p.x_out[i] = a.x_out;
    if (write_h) p.hist_out[i] = a.h_out;
  }
}

__global__ void __launch_bounds__(kBlock)
sonar_step_scalar_kernel(SonarStepParams p) {
  __shared__ NormDecision nd_slot;
  __shared__ double peer_sums[2];
  const bool has_h_in = p.hist_state != SONAR_HIST_NONE;
  const bool has_noise = p.noise_kind == SONAR_NOISE_TENSOR || p.noise_kind == SONAR_NOISE_TENSOR_NORMALIZED;
  const bool write_h = p.hist_out != nullptr;
  const bool norm_noise = p.noise_kind == SONAR_NOISE_TENSOR_NORMALIZED;
  const NoiseNorm nn = make_noise_norm(tensor_noise_decision(p, norm_noise, &nd_slot, peer_sums),
                                       norm_noise ? p.noise_factor : 1.0f);
  const StepConsts c = make_consts(p);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += (int64_t)gridDim.x * blockDim.x) {
    const float nzs = has_noise ? (norm_noise ? norm_noise_value(p.noise[i], nn) : p.noise[i]) : 0.f;
    const StepElem a = step_element(p, c, p.x[i], p.denoised[i], has_h_in ? p.hist_in[i] : 0.f, nzs);
    p.x_out[i] = a.x_out;
    if (write_h) p.hist_out[i] = a.h_out;
  }
}

// ---- Philox variants: CUDA thread <-> virtual ATen thread vt (Philox subsequence), call k ----
// A (vt, k) pair owns the 4 elements li = vt + T*(4k + lane): T apart, so consecutive threads touch
// consecutive addresses (coalesced scalar accesses). Element index li is GLOBAL (position in the
// un-sharded noise tensor); the local tensors hold the slice [noise_begin, noise_begin + n).
__device__ __forceinline__ void step_pair(const SonarStepParams& p, const StepConsts& c, const NoiseNorm& nn,
                                          const float z[4], int64_t li0, int64_t T, int64_t begin, int64_t end,
                                          bool has_h_in, bool write_h) {
  float xs[4], ds[4], hs[4];
  bool ok[4];
#pragma unroll
  for (int lane = 0; lane < 4; ++lane) {  // issue every load of the pair before any arithmetic
    const int64_t li = li0 + T * lane;
    ok[lane] = li >= begin && li < end;
    const int64_t i = ok[lane] ? li - begin : 0;
    xs[lane] = ok[lane] ? __ldg(p.x + i) : 0.0f;
    ds[lane] = ok[lane] ? __ldg(p.denoised + i) : 0.0f;
    hs[lane] = (ok[lane] && has_h_in) ? p.hist_in[i] : 0.0f;
  }
#pragma unroll
  for (int lane = 0; lane < 4; ++lane) {
    if (!ok[lane]) continue;
    const int64_t i = li0 + T * lane - begin;
    const float nz = norm_noise_value(z[lane], nn);
    const StepElem a = step_element(p, c, xs[lane], ds[lane], hs[lane], nz);
    p.x_out[i] = a.x_out;
    if (write_h) p.hist_out[i] = a.h_out;
  }
}

__global__ void __launch_bounds__(kBlock)
sonar_step_philox_kernel(SonarStepParams p, PhiloxStream st, uint32_t k_lo, uint32_t k_hi) {
  const bool has_h_in = p.hist_state != SONAR_HIST_NONE;
  const bool write_h = p.hist_out != nullptr;
  __shared__ NormDecision nd_slot;
  const NormDecision nd = p.noise_kind == SONAR_NOISE_PHILOX_NORMALIZED
                              ? decide_normalisation_block(p.noise_sums, p.noise_count, p.noise_threshold_std_devs, &nd_slot)
                              : NormDecision{0.f, 1.f, 0, 0};
  const NoiseNorm nn = make_noise_norm(nd, p.noise_factor);
  const StepConsts c = make_consts(p);
  const int64_t T = st.threads;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  const int64_t begin = p.noise_begin, end = p.noise_begin + p.n;
  for (int64_t vt = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; vt < T; vt += nthreads) {
    for (uint32_t k = k_lo; k <= k_hi; ++k) {
      const int64_t li0 = vt + T * (int64_t)(4 * (uint64_t)k);
      if (li0 >= end) break;
      if (li0 + 3 * T < begin) continue;
      const float4 z4 = philox_normal4(st, (uint32_t)vt, k);
      const float z[4] = {z4.x, z4.y, z4.z, z4.w};
      step_pair(p, c, nn, z, li0, T, begin, end, has_h_in, write_h);
    }
  }
}

// Single-launch variant for small tensors (launch-bound regime): phase 1 draws the Philox normals
// into registers and reduces their moments, a grid-wide barrier publishes the global sums, phase 2
// applies the conditional normalisation to the SAME registers and performs the step. The whole
// grid must be co-resident (cooperative launch); each thread owns at most kCoopPairs (vt, k) pairs.
// The double[2] sums slot is zeroed for the next launch by the kernel itself (ping-pong slots).
constexpr int kCoopPairs = 4;

// materialise + moments in one pass (same kernel as stats.cu's, local to this TU)
__global__ void __launch_bounds__(kBlock)
philox_fill_moments_device(float* __restrict__ out, int64_t begin, int64_t end, PhiloxStream st, uint32_t k_lo,
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
      } else {
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

// moments of the un-materialised Philox normal draw (same kernel as stats.cu's, local to this TU)
__global__ void __launch_bounds__(kBlock)
philox_normal_moments_device(int64_t begin, int64_t end, PhiloxStream st, uint32_t k_lo, uint32_t k_hi,
                             double* __restrict__ sums) {
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

__global__ void __launch_bounds__(kBlock, 4)
sonar_step_coop_kernel(SonarStepParams p, PhiloxStream st, uint32_t calls, double* __restrict__ slot,
                       double* __restrict__ next_slot) {
  __shared__ double scratch[64];
  const bool has_h_in = p.hist_state != SONAR_HIST_NONE;
  const bool write_h = p.hist_out != nullptr;
  const int64_t T = st.threads;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t end = p.n;  // cooperative path: un-sharded draw, begin == 0
  float z[kCoopPairs][4];
  float fs = 0.0f, fss = 0.0f;
#pragma unroll
  for (int j = 0; j < kCoopPairs; ++j) {
    const int64_t vt = tid + (int64_t)(j / calls) * nthreads;
    const uint32_t k = (uint32_t)j % calls;
    const int64_t li0 = vt + T * (int64_t)(4 * (uint64_t)k);
    const bool live = vt < T && (uint32_t)j < calls * (uint32_t)((T + nthreads - 1) / nthreads) && li0 < end;
    if (live) {
      if (li0 + 2 * T < end) {
        const float4 z4 = philox_normal4(st, (uint32_t)vt, k);
        z[j][0] = z4.x; z[j][1] = z4.y; z[j][2] = z4.z; z[j][3] = z4.w;
      } else {  // lanes 2, 3 lie beyond the tensor: one Box-Muller is enough
        const float2 z2 = philox_normal2_lo(st, (uint32_t)vt, k);
        z[j][0] = z2.x; z[j][1] = z2.y; z[j][2] = 0.0f; z[j][3] = 0.0f;
      }
#pragma unroll
      for (int lane = 0; lane < 4; ++lane) {
        const int64_t li = li0 + T * lane;
        if (li < end) {
          fs += z[j][lane];
          fss += z[j][lane] * z[j][lane];
        }
      }
    } else {
      z[j][0] = z[j][1] = z[j][2] = z[j][3] = 0.0f;
    }
  }
  double s = (double)fs, ss = (double)fss;
  block_sum2(s, ss, scratch);
  if (threadIdx.x == 0) {
    atomicAdd(&slot[0], s);
    atomicAdd(&slot[1], ss);
  }
  cooperative_groups::this_grid().sync();
  __shared__ NormDecision nd_slot;
  const NoiseNorm nn = make_noise_norm(
      decide_normalisation_block(slot, p.noise_count, p.noise_threshold_std_devs, &nd_slot), p.noise_factor);
  const StepConsts c = make_consts(p);
  if (tid == 0) {
    next_slot[0] = 0.0;
    next_slot[1] = 0.0;
  }
#pragma unroll
  for (int j = 0; j < kCoopPairs; ++j) {
    const int64_t vt = tid + (int64_t)(j / calls) * nthreads;
    const uint32_t k = (uint32_t)j % calls;
    const int64_t li0 = vt + T * (int64_t)(4 * (uint64_t)k);
    const bool live = vt < T && (uint32_t)j < calls * (uint32_t)((T + nthreads - 1) / nthreads) && li0 < end;
    if (live) step_pair(p, c, nn, z[j], li0, T, 0, end, has_h_in, write_h);
  }
}

}  // namespace sonar

namespace sonar {
static int g_coop_enabled = -1;  // -1: decide from the environment on first use

static int coop_blocks_per_sm() {
  static thread_local int occupancy = -1;
  if (occupancy < 0) {
    int dev = 0, coop = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    occupancy = 0;
    if (coop) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occupancy, sonar_step_coop_kernel, kBlock, 0);
  }
  // Measured on B200 (profiles/): at 8x4x128x128 the cooperative launch takes 18 us against ~14 us for
  // "materialise + moments" followed by the float4 step kernel -- the ALU-bound Philox phase and the
  // memory phase cannot overlap across the grid barrier. Hence opt-in (SONAR_B200_COOP=1 or
  // sonar_step_enable_cooperative(1)).
  if (g_coop_enabled < 0) g_coop_enabled = getenv("SONAR_B200_COOP") != nullptr ? 1 : 0;
  return g_coop_enabled ? occupancy : 0;
}

// cooperative grid for a draw of `grid_blocks` emulated ATen blocks covering n elements, or 0
static int64_t coop_grid_for(int64_t n, uint32_t grid_blocks) {
  const int64_t T = (int64_t)grid_blocks * kBlock;
  if (T <= 0 || n <= 0) return 0;
  int64_t g = (int64_t)coop_blocks_per_sm() * device_info().sm_count;
  if (g > (int64_t)grid_blocks) g = grid_blocks;
  if (g <= 0) return 0;
  const int64_t calls = ((n - 1) / T) / 4 + 1;
  const int64_t pairs = ((T + g * kBlock - 1) / (g * kBlock)) * calls;
  return pairs <= kCoopPairs ? g : 0;
}
}  // namespace sonar

extern "C" int sonar_step_enable_cooperative(int enable) {
  sonar::g_coop_enabled = enable ? 1 : 0;
  return 0;
}

extern "C" int sonar_step_single_launch_ok(int64_t n, uint32_t philox_grid_blocks) {
  return sonar::coop_grid_for(n, philox_grid_blocks) > 0 ? 1 : 0;
}

extern "C" int sonar_step_f32(const SonarStepParams* params, void* stream_) {
  using namespace sonar;
  if (params == nullptr) return (int)cudaErrorInvalidValue;
  SonarStepParams p = *params;
  if (p.n <= 0) return 0;
  if (p.x == nullptr || p.denoised == nullptr || p.x_out == nullptr) return (int)cudaErrorInvalidValue;
  if (p.hist_state != SONAR_HIST_NONE && p.hist_in == nullptr) return (int)cudaErrorInvalidValue;
  if ((p.noise_kind == SONAR_NOISE_TENSOR || p.noise_kind == SONAR_NOISE_TENSOR_NORMALIZED) && p.noise == nullptr)
    return (int)cudaErrorInvalidValue;
  if (p.noise_kind == SONAR_NOISE_TENSOR_NORMALIZED && p.noise_sums == nullptr && p.peer_world <= 1)
    return (int)cudaErrorInvalidValue;
  if (p.peer_world > 1 && (p.peer_mailbox == nullptr || p.peer_world > SONAR_PEER_MAX_RANKS))
    return (int)cudaErrorInvalidValue;
  if (p.noise_kind == SONAR_NOISE_PHILOX_NORMALIZED && p.noise_sums == nullptr && p.sums_scratch == nullptr)
    return (int)cudaErrorInvalidValue;
  cudaStream_t stream = (cudaStream_t)stream_;

  if (p.noise_kind == SONAR_NOISE_PHILOX || p.noise_kind == SONAR_NOISE_PHILOX_NORMALIZED) {
    if (p.philox_grid_blocks == 0 || p.noise_begin < 0 || p.noise_begin + p.n > p.noise_numel_total)
      return (int)cudaErrorInvalidValue;
    PhiloxStream st{p.philox_seed, p.philox_offset, p.philox_grid_blocks * (uint32_t)kBlock};
    const int64_t T = st.threads, end = p.noise_begin + p.n;
    const int64_t k_lo = (p.noise_begin / T) / 4, k_hi = ((end - 1) / T) / 4;
    const bool self_stats = p.noise_kind == SONAR_NOISE_PHILOX_NORMALIZED && p.noise_sums == nullptr;
    if (self_stats) {
      // the caller left the moments pre-pass to us (un-sharded draw): sums_scratch is double[4],
      // two ping-pong slots, zero-initialised once by the caller
      const bool sharded = p.peer_world > 1;
      if (p.sums_scratch == nullptr || (!sharded && (p.noise_begin != 0 || p.n != p.noise_numel_total)))
        return (int)cudaErrorInvalidValue;
      double* slot = p.sums_scratch + 2 * (p.sums_parity & 1);
      double* next_slot = p.sums_scratch + 2 * ((p.sums_parity & 1) ^ 1);
      if (sharded) {
        // batch-sharded: materialise + moments of THIS rank's slice, store the two partial sums into
        // every rank's mailbox over NVLink, then the step kernel waits for all partials on the device
        if (p.noise == nullptr) return (int)cudaErrorInvalidValue;
        SONAR_CUDA_TRY(cudaMemsetAsync(slot, 0, 2 * sizeof(double), stream));
        philox_fill_moments_device<<<streaming_grid(T, kBlock, 1), kBlock, 0, stream>>>(
            const_cast<float*>(p.noise), p.noise_begin, end, st, (uint32_t)k_lo, (uint32_t)k_hi, slot);
        SONAR_LAUNCH_CHECK();
        const int rc = sonar_peer_publish_sums(p.peer_targets, p.peer_rank, p.peer_world, slot, p.peer_epoch, stream);
        if (rc != 0) return rc;
        p.noise_kind = SONAR_NOISE_TENSOR_NORMALIZED;
        p.noise_sums = slot;
        goto dense_step;
      }
      p.noise_count = p.n;
      const int64_t calls = k_hi + 1;
      const int64_t coop_grid = coop_grid_for(p.n, p.philox_grid_blocks);
      if (coop_grid > 0) {
        uint32_t calls32 = (uint32_t)calls;
        void* args[] = {&p, &st, &calls32, &slot, &next_slot};
        SONAR_CUDA_TRY(cudaLaunchCooperativeKernel((void*)sonar_step_coop_kernel, dim3((unsigned)coop_grid), dim3(kBlock),
                                                   args, 0, stream));
        return 0;
      }
      // Default path: ONE pass that materialises the normals into the caller's scratch (p.noise) while
      // reducing their moments, then the float4 step kernel normalises them on load. Two launches, one
      // C-ABI call; 8 B/element of extra traffic buys a single Philox evaluation per element.
      SONAR_CUDA_TRY(cudaMemsetAsync(slot, 0, 2 * sizeof(double), stream));
      const int grid_m = streaming_grid(T, kBlock, 1);
      if (p.noise != nullptr) {
        philox_fill_moments_device<<<grid_m, kBlock, 0, stream>>>(const_cast<float*>(p.noise), 0, end, st,
                                                                  (uint32_t)k_lo, (uint32_t)k_hi, slot);
        SONAR_LAUNCH_CHECK();
        p.noise_kind = SONAR_NOISE_TENSOR_NORMALIZED;
        p.noise_sums = slot;
        goto dense_step;
      }
      // no scratch given: moments pre-pass, then regenerate the normals inside the step kernel
      philox_normal_moments_device<<<grid_m, kBlock, 0, stream>>>(0, end, st, (uint32_t)k_lo, (uint32_t)k_hi, slot);
      SONAR_LAUNCH_CHECK();
      p.noise_sums = slot;
      p.sums_parity = 0;
    }
    const int grid = streaming_grid(T, kBlock, 1);
    sonar_step_philox_kernel<<<grid, kBlock, 0, stream>>>(p, st, (uint32_t)k_lo, (uint32_t)k_hi);
    SONAR_LAUNCH_CHECK();
    return 0;
  }

dense_step:
  const bool vec_ok = aligned16(p.x) && aligned16(p.denoised) && aligned16(p.x_out) &&
                      (p.hist_in == nullptr || aligned16(p.hist_in)) &&
                      (p.hist_out == nullptr || aligned16(p.hist_out)) && (p.noise == nullptr || aligned16(p.noise));
  if (vec_ok) {
    const int grid = streaming_grid((p.n + 3) / 4, kBlock, 2);
    sonar_step_vec_kernel<<<grid, kBlock, 0, stream>>>(p);
  } else {
    const int grid = streaming_grid(p.n, kBlock, 2);
    sonar_step_scalar_kernel<<<grid, kBlock, 0, stream>>>(p);
  }
  SONAR_LAUNCH_CHECK();
  return 0;
}
