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
  return have_h ? blend<float>(p.history_blend, __fmul_rn(v, p.md_scale), __fmul_rn(h, p.hd_scale), p.hd_ratio) : v;
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
  if (!c.m_is_one && c.have_h_first_mix && c.denoised_mode)
    md = blend<float>(p.momentum_blend, __fmul_rn(h, p.sigma), den, p.momentum);
  if (p.history_active) {
    h = update_history(p, have_h, h, div_by(den, p.sigma, c.inv_sigma));
    have_h = true;
  }
  const float den_eff = p.momentum_active ? md : den;

  // ---- derivative / DPM-Solver++ difference term ----
  const float d = p.kind == SONAR_STEP_EULER ? div_by(__fsub_rn(x, den_eff), p.sigma, c.inv_sigma) : __fmul_rn(p.c0, den_eff);

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
  // (d * dt).add_(x)  /  ((sigma_fn(s) / sigma_fn(t)) * x).sub_(d): every eager op rounds
  r.x_out = p.kind == SONAR_STEP_EULER ? __fadd_rn(__fmul_rn(d_out, p.c0), x) : __fsub_rn(__fmul_rn(p.c1, x), d_out);
  if (p.noise_kind != SONAR_NOISE_NONE) r.x_out = __fadd_rn(r.x_out, __fmul_rn(noise, p.noise_scale));
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
  if (n.sub_mean) v = __fsub_rn(v, n.mean);
  if (n.div_std) v = div_by(v, n.std, n.inv_std);
  return __fmul_rn(v, n.factor);
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

// ---------------------------------------------------------------------------------------------
// Specialised ("fast") element: the configuration every stock Sonar sampler runs with -- momentum
// and history both active, lerp blends, CLASSIC/NEW mode, momentum != 1, history not rescaled on
// load -- with the sampler kind, the mode and the presence of a history as template parameters.
// The generic step_element above evaluates ~10 launch-uniform conditions and two blend switches per
// element (160 issued instructions per element in the round-1 ncu capture, 75 % issue-slot busy: the
// kernel was instruction-bound, not memory-bound); this form is ~30.
// ---------------------------------------------------------------------------------------------
struct LerpW {
  float w, omw;
  bool small;  // |w| < 0.5: a + w*(b-a), else b - (b-a)*(1-w)   (ATen/native/Lerp.h:21-35)
};

__device__ __forceinline__ LerpW make_lerp(float w) { return LerpW{w, 1.0f - w, fabsf(w) < 0.5f}; }

__device__ __forceinline__ float lerp_u(float a, float b, const LerpW& l) {  // == torch_lerp (common.cuh)
  const float d = __fsub_rn(b, a);
  return l.small ? fmaf(l.w, d, a) : fmaf(-d, l.omw, b);
}

struct FastConsts {
  LerpW mom, hist;
  float sigma, inv_sigma, c0, c1, hd_scale, md_scale, noise_scale;
};

__device__ __forceinline__ FastConsts make_fast_consts(const SonarStepParams& p) {
  FastConsts c;
  c.mom = make_lerp(p.momentum);
  c.hist = make_lerp(p.hd_ratio);
  c.sigma = p.sigma;
  c.inv_sigma = 1.0f / p.sigma;
  c.c0 = p.c0;
  c.c1 = p.c1;
  c.hd_scale = p.hd_scale;
  c.md_scale = p.md_scale;
  c.noise_scale = p.noise_scale;
  return c;
}

template <int KIND, bool NEW_MODE, bool HAVE_H, bool NOISE>
__device__ __forceinline__ StepElem step_element_fast(const FastConsts& c, float x, float den, float h, float noise) {
  // get_momentum_denoised: history <- update(denoised / sigma)
  const float den_s = div_by(den, c.sigma, c.inv_sigma);
  const float h1 = HAVE_H ? lerp_u(__fmul_rn(den_s, c.md_scale), __fmul_rn(h, c.hd_scale), c.hist) : den_s;
  // derivative (Euler) or DPM-Solver++ difference term
  const float d = KIND == SONAR_STEP_EULER ? div_by(__fsub_rn(x, den), c.sigma, c.inv_sigma) : __fmul_rn(c.c0, den);
  // get_momentum_d: mix with the history, second history update
  const float mom_d = lerp_u(h1, d, c.mom);
  StepElem r;
  r.h_out = lerp_u(__fmul_rn(NEW_MODE ? d : mom_d, c.md_scale), __fmul_rn(h1, c.hd_scale), c.hist);
  // (d * dt).add_(x)  /  (c1 * x).sub_(d)  then  + noise * (s_noise * sigma_up): every eager op rounds
  r.x_out = KIND == SONAR_STEP_EULER ? __fadd_rn(__fmul_rn(mom_d, c.c0), x) : __fsub_rn(__fmul_rn(c.c1, x), mom_d);
  if (NOISE) r.x_out = __fadd_rn(r.x_out, __fmul_rn(noise, c.noise_scale));
  return r;
}

// NOISE: 0 none, 1 tensor, 2 raw Gaussian tensor normalised on load
template <int KIND, bool NEW_MODE, bool HAVE_H, int NOISE>
__global__ void __launch_bounds__(kBlock)
sonar_step_fast_vec_kernel(SonarStepParams p) {
  __shared__ NormDecision nd_slot;
  __shared__ double peer_sums[2];
  const NoiseNorm nn = make_noise_norm(tensor_noise_decision(p, NOISE == 2, &nd_slot, peer_sums),
                                       NOISE == 2 ? p.noise_factor : 1.0f);
  const FastConsts c = make_fast_consts(p);
  const int64_t n4 = p.n >> 2;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = tid; i < n4; i += stride) {
    const float4 x = ld4_stream(p.x + 4 * i);
    const float4 dn = ld4_stream(p.denoised + 4 * i);
    const float4 h = HAVE_H ? ld4(p.hist_in + 4 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 nz = NOISE ? ld4_stream(p.noise + 4 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
    if (NOISE == 2) {
      nz.x = norm_noise_value(nz.x, nn);
      nz.y = norm_noise_value(nz.y, nn);
      nz.z = norm_noise_value(nz.z, nn);
      nz.w = norm_noise_value(nz.w, nn);
    }
    const StepElem a = step_element_fast<KIND, NEW_MODE, HAVE_H, NOISE != 0>(c, x.x, dn.x, h.x, nz.x);
    const StepElem b = step_element_fast<KIND, NEW_MODE, HAVE_H, NOISE != 0>(c, x.y, dn.y, h.y, nz.y);
    const StepElem e = step_element_fast<KIND, NEW_MODE, HAVE_H, NOISE != 0>(c, x.z, dn.z, h.z, nz.z);
    const StepElem d = step_element_fast<KIND, NEW_MODE, HAVE_H, NOISE != 0>(c, x.w, dn.w, h.w, nz.w);
    st4(p.x_out + 4 * i, make_float4(a.x_out, b.x_out, e.x_out, d.x_out));
    st4(p.hist_out + 4 * i, make_float4(a.h_out, b.h_out, e.h_out, d.h_out));
  }
  for (int64_t i = (n4 << 2) + tid; i < p.n; i += stride) {
    const float nzs = NOISE ? (NOISE == 2 ? norm_noise_value(p.noise[i], nn) : p.noise[i]) : 0.f;
    const StepElem a =
        step_element_fast<KIND, NEW_MODE, HAVE_H, NOISE != 0>(c, p.x[i], p.denoised[i], HAVE_H ? p.hist_in[i] : 0.f, nzs);
    p.x_out[i] = a.x_out;
    p.hist_out[i] = a.h_out;
  }
}

// Philox noise regenerated in registers: CUDA thread <-> virtual ATen thread vt, call k. The 4 lanes
// of a (vt, k) pair are T elements apart, so consecutive threads touch consecutive addresses: every
// access is a coalesced 128-byte warp transaction. All loads of a pair are issued before the Philox
// rounds and the Box-Muller transform, which then run under the loads' latency.
// FULL: all four lanes of the pair lie inside the slice (no per-lane predicates or index clamps).
// LANES = 2: the draw has at most two rows (numel <= 2T, e.g. an SDXL latent batch): lanes 2, 3 never
// exist, one Box-Muller pair per Philox call, and the kernel fits 32 registers -> 8 CTAs per SM, i.e.
// the whole emulated ATen grid (148 x 8 CTAs) is resident in a single wave.
template <int KIND, bool NEW_MODE, bool HAVE_H, bool FULL, int LANES>
__device__ __forceinline__ void step_pair_fast(const SonarStepParams& p, const FastConsts& c, const NoiseNorm& nn,
                                               const PhiloxStream& st, uint32_t vt, uint32_t k, int64_t li0, int64_t T,
                                               int64_t begin, int64_t end) {
  float xs[LANES], ds[LANES], hs[LANES];
  bool ok[LANES];
  int64_t idx[LANES];
#pragma unroll
  for (int lane = 0; lane < LANES; ++lane) {
    const int64_t li = li0 + T * lane;
    ok[lane] = FULL || (li >= begin && li < end);
    idx[lane] = ok[lane] ? li - begin : 0;
    xs[lane] = __ldg(p.x + idx[lane]);
    ds[lane] = __ldg(p.denoised + idx[lane]);
    hs[lane] = HAVE_H ? p.hist_in[idx[lane]] : 0.0f;
  }
  float z[4];
  if (LANES == 4 && (FULL || ok[LANES - 2] || ok[LANES - 1])) {
    const float4 z4 = philox_normal4(st, vt, k);
    z[0] = z4.x; z[1] = z4.y; z[2] = z4.z; z[3] = z4.w;
  } else {  // lanes 2, 3 lie outside the slice: one Box-Muller is enough
    const float2 z2 = philox_normal2_lo(st, vt, k);
    z[0] = z2.x; z[1] = z2.y; z[2] = 0.0f; z[3] = 0.0f;
  }
#pragma unroll
  for (int lane = 0; lane < LANES; ++lane) {
    if (!ok[lane]) continue;
    const StepElem a =
        step_element_fast<KIND, NEW_MODE, HAVE_H, true>(c, xs[lane], ds[lane], hs[lane], norm_noise_value(z[lane], nn));
    p.x_out[idx[lane]] = a.x_out;
    p.hist_out[idx[lane]] = a.h_out;
  }
}

// scale_noise decision for the regenerated normals: either precomputed once per draw
// (sonar_norm_decisions: 16 bytes every thread loads, no barrier) or evaluated by thread 0 of each
// CTA from the raw sums (a fp64 divide + sqrt behind a __syncthreads: ~23 % of the warp stall
// cycles of the C2 launch in the round-1 ncu capture).
__device__ __forceinline__ NormDecision philox_noise_decision(const SonarStepParams& p, NormDecision* slot) {
  if (p.noise_kind != SONAR_NOISE_PHILOX_NORMALIZED) return NormDecision{0.f, 1.f, 0, 0};
  if (p.noise_decision != nullptr) {  // launch-uniform
    const float4 d = __ldg(reinterpret_cast<const float4*>(p.noise_decision));
    return NormDecision{d.x, d.y, d.z != 0.0f ? 1 : 0, d.w != 0.0f ? 1 : 0};
  }
  return decide_normalisation_block(p.noise_sums, p.noise_count, p.noise_threshold_std_devs, slot);
}

constexpr int kPhiloxStepBlock = 128;  // finer-grained CTAs: the 1.6-wave tail of 256-thread CTAs costs ~10 %

template <int KIND, bool NEW_MODE, bool HAVE_H>
__global__ void __launch_bounds__(kPhiloxStepBlock)
sonar_step_fast_philox_kernel(SonarStepParams p, PhiloxStream st, uint32_t k_lo, uint32_t k_hi) {
  __shared__ NormDecision nd_slot;
  const NoiseNorm nn = make_noise_norm(philox_noise_decision(p, &nd_slot), p.noise_factor);
  const FastConsts c = make_fast_consts(p);
  const int64_t T = st.threads;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  const int64_t begin = p.noise_begin, end = p.noise_begin + p.n;
  for (int64_t vt = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; vt < T; vt += nthreads) {
    for (uint32_t k = k_lo; k <= k_hi; ++k) {
      const int64_t li0 = vt + T * (int64_t)(4 * (uint64_t)k);
      if (li0 >= end) break;
      if (li0 + 3 * T < begin) continue;
      if (li0 >= begin && li0 + 3 * T < end)
        step_pair_fast<KIND, NEW_MODE, HAVE_H, true, 4>(p, c, nn, st, (uint32_t)vt, k, li0, T, begin, end);
      else
        step_pair_fast<KIND, NEW_MODE, HAVE_H, false, 4>(p, c, nn, st, (uint32_t)vt, k, li0, T, begin, end);
    }
  }
}

// Vectorised Philox variant for draws of more than two rows: one thread owns FOUR consecutive virtual ATen threads
// (vt .. vt + 3) of one call k, i.e. for every lane four CONSECUTIVE elements -- x, denoised and the history move as
// 16-byte loads and stores (the scalar form above issues one 4-byte access per element: LSU-instruction bound at
// 0.62 of the measured HBM peak at one video latent). The 16 normals of the four calls stay in registers; the three
// float4 loads of a lane are issued before that lane's arithmetic. Needs begin, n and T to be multiples of 4 and
// 16-byte aligned tensors (the launcher checks; anything else takes the scalar kernel).
template <int KIND, bool NEW_MODE, bool HAVE_H>
__global__ void __launch_bounds__(kBlock)
sonar_step_fast_philox4_kernel(SonarStepParams p, PhiloxStream st, uint32_t k_lo, uint32_t n_calls, uint32_t groups) {
  __shared__ NormDecision nd_slot;
  const NoiseNorm nn = make_noise_norm(philox_noise_decision(p, &nd_slot), p.noise_factor);
  const FastConsts c = make_fast_consts(p);
  const int64_t T = st.threads;
  const int64_t begin = p.noise_begin, end = p.noise_begin + p.n;
  const int64_t items = (int64_t)n_calls * groups;  // (call, group of four virtual threads), group fastest
  for (int64_t item = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; item < items; item += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t kk = (uint32_t)(item / groups);
    const uint32_t vt = (uint32_t)(item - (int64_t)kk * groups) * 4u;
    const uint32_t k = k_lo + kk;
    const int64_t li0 = (int64_t)vt + T * (int64_t)(4 * (uint64_t)k);
    if (li0 >= end || li0 + 3 * T + 4 <= begin) continue;
    float z[4][4];  // [virtual thread][lane]
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 z4 = philox_normal4(st, vt + j, k);
      z[j][0] = z4.x; z[j][1] = z4.y; z[j][2] = z4.z; z[j][3] = z4.w;
    }
#pragma unroll
    for (int lane = 0; lane < 4; ++lane) {
      const int64_t li = li0 + T * lane;
      if (li < begin || li >= end) continue;  // whole groups of four are inside or outside (begin, n multiples of 4)
      const int64_t i = li - begin;
      const float4 x = ld4_stream(p.x + i);
      const float4 dn = ld4_stream(p.denoised + i);
      const float4 h = HAVE_H ? ld4(p.hist_in + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      const StepElem a = step_element_fast<KIND, NEW_MODE, HAVE_H, true>(c, x.x, dn.x, h.x, norm_noise_value(z[0][lane], nn));
      const StepElem b = step_element_fast<KIND, NEW_MODE, HAVE_H, true>(c, x.y, dn.y, h.y, norm_noise_value(z[1][lane], nn));
      const StepElem e = step_element_fast<KIND, NEW_MODE, HAVE_H, true>(c, x.z, dn.z, h.z, norm_noise_value(z[2][lane], nn));
      const StepElem d = step_element_fast<KIND, NEW_MODE, HAVE_H, true>(c, x.w, dn.w, h.w, norm_noise_value(z[3][lane], nn));
      st4(p.x_out + i, make_float4(a.x_out, b.x_out, e.x_out, d.x_out));
      st4(p.hist_out + i, make_float4(a.h_out, b.h_out, e.h_out, d.h_out));
    }
  }
}

// Draws of at most two rows (numel_total <= 2T, un-sharded or sharded): k == 0, lanes 0 and 1 only.
template <int KIND, bool NEW_MODE, bool HAVE_H>
__global__ void __launch_bounds__(kBlock, 8)
sonar_step_fast_philox2_kernel(SonarStepParams p, PhiloxStream st) {
  __shared__ NormDecision nd_slot;
  const NoiseNorm nn = make_noise_norm(philox_noise_decision(p, &nd_slot), p.noise_factor);
  const FastConsts c = make_fast_consts(p);
  const int64_t T = st.threads;
  const int64_t begin = p.noise_begin, end = p.noise_begin + p.n;
  const int64_t vt = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // grid == the emulated ATen grid
  if (vt >= T || vt >= end || vt + T < begin) return;
  if (vt >= begin && vt + T < end)
    step_pair_fast<KIND, NEW_MODE, HAVE_H, true, 2>(p, c, nn, st, (uint32_t)vt, 0u, vt, T, begin, end);
  else
    step_pair_fast<KIND, NEW_MODE, HAVE_H, false, 2>(p, c, nn, st, (uint32_t)vt, 0u, vt, T, begin, end);
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
  const NormDecision nd = p.noise_kind != SONAR_NOISE_PHILOX_NORMALIZED
                              ? NormDecision{0.f, 1.f, 0, 0}
                              : (p.noise_decision != nullptr
                                     ? NormDecision{p.noise_decision[0], p.noise_decision[1], p.noise_decision[2] != 0.0f ? 1 : 0,
                                                    p.noise_decision[3] != 0.0f ? 1 : 0}
                                     : decide_normalisation_block(p.noise_sums, p.noise_count, p.noise_threshold_std_devs, &nd_slot));
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

}  // namespace sonar

namespace sonar {
// the specialised kernels cover the configuration of the stock samplers (see step_element_fast)
static bool fast_config(const SonarStepParams& p) {
  return p.mode != SONAR_MODE_DENOISED && p.momentum != 1.0f && p.momentum_active && p.history_active &&
         p.momentum_blend == SONAR_BLEND_LERP && p.history_blend == SONAR_BLEND_LERP && p.hist_out != nullptr &&
         (p.hist_state == SONAR_HIST_NONE || p.hist_in_div == 1.0f);
}

template <int KIND, bool NEW_MODE, bool HAVE_H>
static void launch_fast_vec(const SonarStepParams& p, int grid, cudaStream_t stream) {
  const bool tensor = p.noise_kind == SONAR_NOISE_TENSOR, norm = p.noise_kind == SONAR_NOISE_TENSOR_NORMALIZED;
  if (norm)
    sonar_step_fast_vec_kernel<KIND, NEW_MODE, HAVE_H, 2><<<grid, kBlock, 0, stream>>>(p);
  else if (tensor)
    sonar_step_fast_vec_kernel<KIND, NEW_MODE, HAVE_H, 1><<<grid, kBlock, 0, stream>>>(p);
  else
    sonar_step_fast_vec_kernel<KIND, NEW_MODE, HAVE_H, 0><<<grid, kBlock, 0, stream>>>(p);
}

template <int KIND, bool NEW_MODE, bool HAVE_H>
static void launch_fast_philox(const SonarStepParams& p, const PhiloxStream& st, uint32_t k_lo, uint32_t k_hi,
                               cudaStream_t stream) {
  const int64_t T = st.threads;
  if (p.noise_numel_total <= 2 * T) {  // two rows at most: single-wave kernel, grid == emulated ATen grid
    sonar_step_fast_philox2_kernel<KIND, NEW_MODE, HAVE_H><<<(unsigned)(T / kBlock), kBlock, 0, stream>>>(p, st);
    return;
  }
  const bool vec4 = (T & 3) == 0 && (p.noise_begin & 3) == 0 && (p.n & 3) == 0 && aligned16(p.x) && aligned16(p.denoised) &&
                    aligned16(p.x_out) && aligned16(p.hist_out) && (p.hist_in == nullptr || aligned16(p.hist_in));
  if (vec4) {
    const uint32_t groups = (uint32_t)(T / 4), n_calls = k_hi - k_lo + 1;
    const int grid = streaming_grid((int64_t)groups * n_calls, kBlock, 4);
    sonar_step_fast_philox4_kernel<KIND, NEW_MODE, HAVE_H><<<grid, kBlock, 0, stream>>>(p, st, k_lo, n_calls, groups);
    return;
  }
  const int grid = streaming_grid(T, kPhiloxStepBlock, 1);
  sonar_step_fast_philox_kernel<KIND, NEW_MODE, HAVE_H><<<grid, kPhiloxStepBlock, 0, stream>>>(p, st, k_lo, k_hi);
}

// expands to the 8 (kind, mode, history) instantiations of LAUNCH<...>(args)
#define SONAR_DISPATCH_FAST(LAUNCH, p, ...)                                                       \
  do {                                                                                            \
    const bool euler__ = (p).kind == SONAR_STEP_EULER, new__ = (p).mode == SONAR_MODE_NEW;        \
    const bool have_h__ = (p).hist_state != SONAR_HIST_NONE;                                      \
    if (euler__) {                                                                                \
      if (new__) {                                                                                \
        if (have_h__) LAUNCH<SONAR_STEP_EULER, true, true>(__VA_ARGS__);                          \
        else LAUNCH<SONAR_STEP_EULER, true, false>(__VA_ARGS__);                                  \
      } else {                                                                                    \
        if (have_h__) LAUNCH<SONAR_STEP_EULER, false, true>(__VA_ARGS__);                         \
        else LAUNCH<SONAR_STEP_EULER, false, false>(__VA_ARGS__);                                 \
      }                                                                                           \
    } else {                                                                                      \
      if (new__) {                                                                                \
        if (have_h__) LAUNCH<SONAR_STEP_DPMPP, true, true>(__VA_ARGS__);                          \
        else LAUNCH<SONAR_STEP_DPMPP, true, false>(__VA_ARGS__);                                  \
      } else {                                                                                    \
        if (have_h__) LAUNCH<SONAR_STEP_DPMPP, false, true>(__VA_ARGS__);                         \
        else LAUNCH<SONAR_STEP_DPMPP, false, false>(__VA_ARGS__);                                 \
      }                                                                                           \
    }                                                                                             \
  } while (0)
}  // namespace sonar

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
  if (p.noise_kind == SONAR_NOISE_PHILOX_NORMALIZED && p.noise_sums == nullptr && p.noise_decision == nullptr)
    return (int)cudaErrorInvalidValue;
  if (p.noise_decision != nullptr && (reinterpret_cast<uintptr_t>(p.noise_decision) & 15u)) return (int)cudaErrorInvalidValue;
  cudaStream_t stream = (cudaStream_t)stream_;

  if (p.noise_kind == SONAR_NOISE_PHILOX || p.noise_kind == SONAR_NOISE_PHILOX_NORMALIZED) {
    if (p.philox_grid_blocks == 0 || p.noise_begin < 0 || p.noise_begin + p.n > p.noise_numel_total)
      return (int)cudaErrorInvalidValue;
    PhiloxStream st{p.philox_seed, p.philox_offset, p.philox_grid_blocks * (uint32_t)kBlock};
    const int64_t T = st.threads, end = p.noise_begin + p.n;
    const int64_t k_lo = (p.noise_begin / T) / 4, k_hi = ((end - 1) / T) / 4;
    const int grid = streaming_grid(T, kBlock, 1);
    if (fast_config(p))
      SONAR_DISPATCH_FAST(launch_fast_philox, p, p, st, (uint32_t)k_lo, (uint32_t)k_hi, stream);
    else
      sonar_step_philox_kernel<<<grid, kBlock, 0, stream>>>(p, st, (uint32_t)k_lo, (uint32_t)k_hi);
    SONAR_LAUNCH_CHECK();
    return 0;
  }

  const bool vec_ok = aligned16(p.x) && aligned16(p.denoised) && aligned16(p.x_out) &&
                      (p.hist_in == nullptr || aligned16(p.hist_in)) &&
                      (p.hist_out == nullptr || aligned16(p.hist_out)) && (p.noise == nullptr || aligned16(p.noise));
  if (vec_ok) {
    const int grid = streaming_grid_shared((p.n + 3) / 4, kBlock, 2);
    if (fast_config(p))
      SONAR_DISPATCH_FAST(launch_fast_vec, p, p, grid, stream);
    else
      sonar_step_vec_kernel<<<grid, kBlock, 0, stream>>>(p);
  } else {
    const int grid = streaming_grid(p.n, kBlock, 2);
    sonar_step_scalar_kernel<<<grid, kBlock, 0, stream>>>(p);
  }
  SONAR_LAUNCH_CHECK();
  return 0;
}
