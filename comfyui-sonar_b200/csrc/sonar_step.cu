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

__device__ __forceinline__ StepElem step_element(const SonarStepParams& p, float x, float den, float h_raw,
                                                 float noise) {
  const bool m_is_one = p.momentum == 1.0f;
  const bool denoised_mode = p.mode == SONAR_MODE_DENOISED;
  bool have_h = p.hist_state != SONAR_HIST_NONE;
  const bool have_h_first_mix = p.hist_state == SONAR_HIST_PRESENT;
  float h = have_h ? h_raw / p.hist_in_div : 0.0f;

  // ---- get_momentum_denoised ----
  float md = den;
  if (!m_is_one && have_h_first_mix && denoised_mode) md = blend<float>(p.momentum_blend, h * p.sigma, den, p.momentum);
  if (p.history_active) {
    h = update_history(p, have_h, h, den / p.sigma);
    have_h = true;
  }
  const float den_eff = p.momentum_active ? md : den;

  // ---- derivative / DPM-Solver++ difference term ----
  const float d = p.kind == SONAR_STEP_EULER ? (x - den_eff) / p.sigma : p.c0 * den_eff;

  // ---- get_momentum_d ----
  float d_out = d;
  if (!m_is_one && !denoised_mode) {
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

// ---- contiguous float4 variant (no noise / tensor noise) ----
__global__ void __launch_bounds__(kBlock)
sonar_step_vec_kernel(SonarStepParams p) {
  const bool has_h_in = p.hist_state != SONAR_HIST_NONE;
  const bool has_noise = p.noise_kind == SONAR_NOISE_TENSOR;
  const bool write_h = p.hist_out != nullptr;
  const int64_t n4 = p.n >> 2;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = tid; i < n4; i += stride) {
    const float4 x = ld4_stream(p.x + 4 * i);
    const float4 dn = ld4_stream(p.denoised + 4 * i);
    const float4 h = has_h_in ? ld4(p.hist_in + 4 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 nz = has_noise ? ld4_stream(p.noise + 4 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
    const StepElem a = step_element(p, x.x, dn.x, h.x, nz.x);
    const StepElem b = step_element(p, x.y, dn.y, h.y, nz.y);
    const StepElem c = step_element(p, x.z, dn.z, h.z, nz.z);
    const StepElem d = step_element(p, x.w, dn.w, h.w, nz.w);
    st4(p.x_out + 4 * i, make_float4(a.x_out, b.x_out, c.x_out, d.x_out));
    if (write_h) st4(p.hist_out + 4 * i, make_float4(a.h_out, b.h_out, c.h_out, d.h_out));
  }
  for (int64_t i = (n4 << 2) + tid; i < p.n; i += stride) {
    const StepElem a = step_element(p, p.x[i], p.denoised[i], has_h_in ? p.hist_in[i] : 0.f, has_noise ? p.noise[i] : 0.f);
    p.x_out[i] = a.x_out;
    if (write_h) p.hist_out[i] = a.h_out;
  }
}

__global__ void __launch_bounds__(kBlock)
sonar_step_scalar_kernel(SonarStepParams p) {
  const bool has_h_in = p.hist_state != SONAR_HIST_NONE;
  const bool has_noise = p.noise_kind == SONAR_NOISE_TENSOR;
  const bool write_h = p.hist_out != nullptr;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += (int64_t)gridDim.x * blockDim.x) {
    const StepElem a = step_element(p, p.x[i], p.denoised[i], has_h_in ? p.hist_in[i] : 0.f, has_noise ? p.noise[i] : 0.f);
    p.x_out[i] = a.x_out;
    if (write_h) p.hist_out[i] = a.h_out;
  }
}

// ---- Philox variant: thread <-> (emulated ATen thread t, call k); 4 elements T apart ----
// Element index li below is GLOBAL (position in the un-sharded noise tensor); the local tensors
// hold the slice [noise_begin, noise_begin + n).
__global__ void __launch_bounds__(kBlock)
sonar_step_philox_kernel(SonarStepParams p, PhiloxStream st, uint32_t k_lo, int64_t n_pairs) {
  const bool has_h_in = p.hist_state != SONAR_HIST_NONE;
  const bool write_h = p.hist_out != nullptr;
  const NormDecision nd = p.noise_kind == SONAR_NOISE_PHILOX_NORMALIZED
                              ? decide_normalisation(p.noise_sums, p.noise_count, p.noise_threshold_std_devs)
                              : NormDecision{0.f, 1.f, 0, 0};
  const int64_t T = st.threads;
  const int64_t begin = p.noise_begin, end = p.noise_begin + p.n;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n_pairs;
       q += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t t = (uint32_t)(q % T);
    const uint64_t k = k_lo + (uint64_t)(q / T);
    const int64_t li0 = (int64_t)t + T * (int64_t)(4 * k);
    if (li0 >= end) continue;  // whole pair beyond the slice (li grows with lane)
    const float4 z = philox_normal4(st, t, k);
    const float zs[4] = {z.x, z.y, z.z, z.w};
#pragma unroll
    for (int lane = 0; lane < 4; ++lane) {
      const int64_t li = li0 + T * lane;
      if (li >= begin && li < end) {
        const int64_t i = li - begin;
        const float nz = apply_norm(zs[lane], nd) * p.noise_factor;
        const StepElem a = step_element(p, p.x[i], p.denoised[i], has_h_in ? p.hist_in[i] : 0.f, nz);
        p.x_out[i] = a.x_out;
        if (write_h) p.hist_out[i] = a.h_out;
      }
    }
  }
}

}  // namespace sonar

extern "C" int sonar_step_f32(const SonarStepParams* params, void* stream_) {
  using namespace sonar;
  if (params == nullptr) return (int)cudaErrorInvalidValue;
  SonarStepParams p = *params;
  if (p.n <= 0) return 0;
  if (p.x == nullptr || p.denoised == nullptr || p.x_out == nullptr) return (int)cudaErrorInvalidValue;
  if (p.hist_state != SONAR_HIST_NONE && p.hist_in == nullptr) return (int)cudaErrorInvalidValue;
  if (p.noise_kind == SONAR_NOISE_TENSOR && p.noise == nullptr) return (int)cudaErrorInvalidValue;
  if (p.noise_kind == SONAR_NOISE_PHILOX_NORMALIZED && p.noise_sums == nullptr) return (int)cudaErrorInvalidValue;
  cudaStream_t stream = (cudaStream_t)stream_;

  if (p.noise_kind == SONAR_NOISE_PHILOX || p.noise_kind == SONAR_NOISE_PHILOX_NORMALIZED) {
    if (p.philox_grid_blocks == 0 || p.noise_begin < 0 || p.noise_begin + p.n > p.noise_numel_total)
      return (int)cudaErrorInvalidValue;
    PhiloxStream st{p.philox_seed, p.philox_offset, p.philox_grid_blocks * (uint32_t)kBlock};
    const int64_t T = st.threads, end = p.noise_begin + p.n;
    const int64_t k_lo = (p.noise_begin / T) / 4, k_hi = ((end - 1) / T) / 4;
    const int64_t n_pairs = T * (k_hi - k_lo + 1);
    const int grid = streaming_grid(n_pairs, kBlock, 1);
    sonar_step_philox_kernel<<<grid, kBlock, 0, stream>>>(p, st, (uint32_t)k_lo, n_pairs);
    SONAR_LAUNCH_CHECK();
    return 0;
  }

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
