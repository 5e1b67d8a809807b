/*
 * sonar_b200 -- C ABI of libsonar_b200.so (hand-written sm_100a CUDA kernels).
 *
 * The reference (blepping/ComfyUI-sonar) is pure Python on eager PyTorch and has no FFI of its own
 * (SURVEY.md section 8b): the drop-in boundary is the ComfyUI node surface, kept in Python by the
 * package `comfyui-sonar_b200/`. This header is the NEW seam underneath that surface: each entry
 * point replaces a group of eager ATen passes in the reference, cited as file:line below.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in `_host`;
 *   - tensors are dense, row-major (NCHW / planes x H x W), fp32 unless the name says f64;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises;
 *   - return value: 0 on success, otherwise the cudaError_t of the failed call / launch
 *     (cudaErrorInvalidValue == 1 for argument errors). No exceptions, no global state.
 *   - there is NO CPU fallback behind any of these symbols.
 */
#ifndef SONAR_B200_H_
#define SONAR_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SONAR_B200_ABI_VERSION 9
int sonar_abi_version(void);
/* Binds the calling thread of THIS library's CUDA runtime to `device` (the library links cudart
 * statically; one process per GPU normally makes this a no-op). */
int sonar_set_device(int device);
/* Co-scheduling hint for the calling thread's next launches: when ctas_per_sm > 0 the grid-stride streaming kernels
 * (Philox fills, the vectorised fused step) launch at most that many CTAs per SM, so that a kernel enqueued on
 * ANOTHER stream finds free thread slots on every SM and the two run side by side (the ALU-bound noise producers under
 * the HBM-bound step). 0 restores the default grids. No reference counterpart: eager PyTorch runs one kernel at a time. */
int sonar_set_grid_limit(int ctas_per_sm);

/* ------------------------------------------------------------------------------------------------
 * Philox generators, bit-exact with torch.randn / torch.rand / Tensor.uniform_ on CUDA.
 * replaces: NoiseGenerator.rand_like                       py/noise_generation.py:133-155
 *           torch.randn in PyramidNoiseGenerator.generate  py/noise_generation.py:632-640
 *           Tensor.uniform_ in perlin_noise                py/noise_generation.py:465-469
 *           complex64 torch.randn in PowerNoiseItem        py/nodes/powernoise.py:396-401
 * mirrors : ATen/native/cuda/DistributionTemplates.h:50-82 (launch policy and element mapping)
 *
 * A draw of `numel_total` elements is identified by (seed, offset, grid_blocks) where offset is the
 * torch CUDA generator's philox offset BEFORE the draw and grid_blocks comes from
 * sonar_philox_policy(). [begin, begin+count) selects a slice of the flattened draw (batch
 * sharding); out[0] receives element `begin`. A complex64 randn is a float normal draw over
 * 2*numel floats with std = 1/sqrt(2) (interleaved re, im).
 * ---------------------------------------------------------------------------------------------- */
int sonar_philox_policy(int64_t numel, uint32_t* grid_blocks_host, uint64_t* counter_offset_host);
int sonar_philox_normal_f32(float* out, int64_t begin, int64_t count, int64_t numel_total, uint64_t seed,
                            uint64_t offset, uint32_t grid_blocks, float mean, float std, void* stream);
int sonar_philox_uniform_f32(float* out, int64_t begin, int64_t count, int64_t numel_total, uint64_t seed,
                             uint64_t offset, uint32_t grid_blocks, float from, float to, void* stream);

/* Several draws in one launch (a pyramid sample = base + every level; a Perlin sample = base + angle
 * grids). Each draw keeps its own geometry / offset / transform (kind 0: normal with mean p0, std p1;
 * kind 1: uniform on [p0, p1)), so the values equal those of separate calls; all share `seed`. */
#define SONAR_FILL_BATCH_MAX 32
typedef struct SonarFillDesc {
  float* out;
  int64_t begin;
  int64_t count;
  int64_t numel_total;
  uint64_t offset;
  uint32_t grid_blocks;
  int32_t kind;
  float p0;
  float p1;
} SonarFillDesc;
typedef struct SonarFillBatch {
  int32_t n;
  uint64_t seed;
  SonarFillDesc draws[SONAR_FILL_BATCH_MAX];
} SonarFillBatch;
int sonar_philox_fill_batch(const SonarFillBatch* batch_host, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Global moments and scale_noise.
 * replaces: scale_noise                                    py/utils.py:85-106
 *           chain accumulation + normalisation             py/noise.py:189-194
 * `sums` is double[2] = {sum, sum of squares}; the kernels ACCUMULATE into it (zero it first),
 * so a batch-sharded run can all-reduce the two doubles between the moments pass and the apply
 * pass. `count` is the GLOBAL element count the sums cover. `out` may alias `x`.
 * ---------------------------------------------------------------------------------------------- */
int sonar_moments_f32(const float* x, int64_t n, double* sums, void* stream);
int sonar_philox_normal_moments(int64_t begin, int64_t count, int64_t numel_total, uint64_t seed, uint64_t offset,
                                uint32_t grid_blocks, double* sums, void* stream);
/* Moments of n_draws draws of the same geometry (same seed / grid_blocks / slice), one launch:
 * sums[2*i], sums[2*i+1] are OVERWRITTEN with the sum and the sum of squares of the slice of the draw
 * whose generator offset is offsets_host[i]. A sampler calls this once, at its first noise request,
 * for the ancestral-noise draws of ALL its remaining steps (their offsets are known in advance), so
 * that each step is a single sonar_step_f32 launch with SONAR_NOISE_PHILOX_NORMALIZED. */
int sonar_philox_normal_moments_batch(const uint64_t* offsets_host, int n_draws, int64_t begin, int64_t count,
                                      int64_t numel_total, uint64_t seed, uint32_t grid_blocks, double* sums,
                                      void* stream);
/* decisions[4*i .. 4*i+3] = {mean, std, subtract-mean flag, divide flag} of scale_noise (py/utils.py:100-106)
 * for the n statistics sums[2*i], sums[2*i+1] over `count` elements each. */
int sonar_norm_decisions(const double* sums, int n, int64_t count, float threshold_std_devs, float* decisions,
                         void* stream);
/* materialise the slice AND write (not accumulate) its moments into sums, one pass */
int sonar_philox_normal_fill_moments_f32(float* out, int64_t begin, int64_t count, int64_t numel_total, uint64_t seed,
                                         uint64_t offset, uint32_t grid_blocks, double* sums, void* stream);
int sonar_scale_noise_f32(const float* x, float* out, int64_t n, const double* sums, int64_t count, float factor,
                          float threshold_std_devs, void* stream);
/* out = x * (scale / unbiased_std) -- GreenTestNoiseGenerator.generate, py/noise_generation.py:703 */
int sonar_scale_by_std_f32(const float* x, float* out, int64_t n, const double* sums, int64_t count, float scale,
                           void* stream);

/* ------------------------------------------------------------------------------------------------
 * Fused Sonar momentum step.
 * replaces: SonarBase.update_hist/momentum_mix/get_momentum_denoised/get_momentum_d/momentum_step
 *                                                          py/sonar.py:227-320
 *           SonarEuler.step / SonarEulerAncestral.step     py/sonar.py:460-480, :541-573
 *           SonarDPMPPSDE.momentum_step (each half step)   py/sonar.py:649-735
 * ---------------------------------------------------------------------------------------------- */
enum { SONAR_STEP_EULER = 0, SONAR_STEP_DPMPP = 1 };
enum { SONAR_MODE_CLASSIC = 0, SONAR_MODE_NEW = 1, SONAR_MODE_DENOISED = 2 };
enum { SONAR_BLEND_LERP = 0, SONAR_BLEND_INJECT = 1, SONAR_BLEND_SUBTRACT_B = 2 };
/* history at entry: none (first step, ZERO init) / present / initialised by this very call
 * (SAMPLE, SAMPLE_NORM, RAND init: used by the updates but not by the first momentum mix,
 * py/sonar.py:272-282) */
enum { SONAR_HIST_NONE = 0, SONAR_HIST_PRESENT = 1, SONAR_HIST_INIT = 2 };
enum {
  SONAR_NOISE_NONE = 0,
  SONAR_NOISE_TENSOR = 1,            /* noise read from `noise` */
  SONAR_NOISE_PHILOX = 2,            /* torch.randn(device='cuda') regenerated in registers */
  SONAR_NOISE_PHILOX_NORMALIZED = 3, /* ... followed by scale_noise decided from `noise_sums` */
  SONAR_NOISE_TENSOR_NORMALIZED = 4  /* raw Gaussian in `noise`, scale_noise applied on load from `noise_sums` */
};

typedef struct SonarStepParams {
  const float* x;        /* sample at sigma                                   */
  const float* denoised; /* model output                                      */
  const float* hist_in;  /* history_d, may alias hist_out; NULL if HIST_NONE  */
  const float* noise;    /* SONAR_NOISE_TENSOR only                           */
  float* x_out;          /* must not alias x                                  */
  float* hist_out;       /* NULL: do not store the history                    */
  int64_t n;             /* elements in this (local) tensor                   */

  int32_t kind;            /* SONAR_STEP_*                                     */
  int32_t mode;            /* SONAR_MODE_*                                     */
  int32_t momentum_blend;  /* SONAR_BLEND_*                                    */
  int32_t history_blend;   /* SONAR_BLEND_*                                    */
  int32_t hist_state;      /* SONAR_HIST_*                                     */
  int32_t momentum_active; /* check_step(step)                                 */
  int32_t history_active;  /* check_step(step, is_history) && momentum_hist!=1 */
  int32_t noise_kind;      /* SONAR_NOISE_*                                    */

  float momentum;
  float sigma;       /* sigma the denoised was evaluated at                   */
  float c0;          /* EULER: dt = sigma_down - sigma; DPMPP: expm1(t - s)   */
  float c1;          /* DPMPP: sigma_fn(s) / sigma_fn(t); unused for EULER    */
  float hd_ratio;    /* history_ratios, py/sonar.py:208-219                   */
  float hd_scale;
  float md_scale;
  float hist_in_div; /* history = hist_in / hist_in_div (SAMPLE_NORM init)    */
  float noise_scale; /* s_noise * sigma_up                                    */
  float noise_factor;            /* factor applied by scale_noise (philox kinds) */
  float noise_threshold_std_devs;

  /* Philox noise: the draw torch.randn(x.shape) would make, this tensor = slice at noise_begin */
  uint64_t philox_seed;
  uint64_t philox_offset;
  uint32_t philox_grid_blocks;
  int64_t noise_begin;
  int64_t noise_numel_total;
  /* double[2] = {sum, sum of squares} of the WHOLE noise tensor (all ranks), device resident:
   * PHILOX_NORMALIZED: reduced ahead of time by sonar_philox_normal_moments_batch (and all-reduced
   * when the batch is sharded); TENSOR_NORMALIZED: reduced by the producer of `noise`. */
  const double* noise_sums;
  int64_t noise_count; /* global element count behind noise_sums */
  /* PHILOX_NORMALIZED, optional: float[4] = {mean, std, subtract-mean flag, divide flag} from
   * sonar_norm_decisions (16-byte aligned). Takes precedence over noise_sums: the decision is then a
   * 16-byte load instead of fp64 arithmetic behind a barrier in every CTA. */
  const float* noise_decision;
  /* TENSOR_NORMALIZED with batch-sharded statistics over peer memory (see sonar_peer_*): when
   * peer_world > 1 the kernel waits until all peer_world partial sums of `peer_epoch` have landed in
   * the LOCAL mailbox `peer_mailbox` and normalises with their total (noise_sums is ignored). */
  int32_t peer_world;
  const double* peer_mailbox;
  double peer_epoch;
} SonarStepParams;

int sonar_step_f32(const SonarStepParams* params_host, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Reference-latent guidance epilogue (runs right after the fused step of a guided sampler).
 * replaces: SonarGuidanceMixin.guidance_shift / guidance_linear / guidance_euler
 *                                                          py/sonar.py:372-411
 * sonar_item_moments_f32: sums[2*i], sums[2*i+1] are OVERWRITTEN with the sum / sum of squares of item i
 * of a dense (items, per_item) tensor (the per-batch-item mean / unbiased std of guidance_shift).
 * sonar_guidance_f32:  target = ref * std_i + mean_i  (statistics from item_sums; NULL = no shift)
 *   LINEAR: out = blend(x, target, factor)          EULER: out = x + (x - target) / sigma * dt
 * ref is (items, per_item), or (1, per_item) broadcast over the batch (ref_items == 1). out may alias x.
 * ---------------------------------------------------------------------------------------------- */
enum { SONAR_GUIDANCE_LINEAR = 0, SONAR_GUIDANCE_EULER = 1 };
typedef struct SonarGuidanceParams {
  const float* x;
  const float* ref;
  const double* item_sums; /* double[items][2] */
  float* out;
  int64_t items;
  int64_t per_item;
  int32_t ref_items;
  int32_t kind;       /* SONAR_GUIDANCE_* */
  int32_t blend_mode; /* SONAR_BLEND_*, LINEAR only */
  float factor;       /* LINEAR only */
  float sigma;        /* EULER only */
  float dt;           /* EULER only: (sigma_next - sigma) * factor */
} SonarGuidanceParams;
int sonar_item_moments_f32(const float* x, int64_t items, int64_t per_item, double* sums, void* stream);
int sonar_guidance_f32(const SonarGuidanceParams* params_host, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Pyramid family: fused multi-level resample-and-accumulate.
 *   out[p,y,x] = base_scale*base[p,y,x] + sum_i weights[i] * resample(levels[i][p] -> HxW)[y,x]
 * replaces: PyramidNoiseGenerator.generate                 py/noise_generation.py:621-649
 *           HighresPyramidNoiseGenerator.generate          py/noise_generation.py:539-564
 *           PyramidOldNoiseGenerator.generate              py/noise_generation.py:579-606
 *           utils.scale_samples -> common_upscale          py/utils.py:58-67
 * Level tensors are (planes, level_h, level_w); level sizes are decided on the host from the CPU
 * generator draws exactly as the reference does (:627-630). `base` may be NULL.
 * ---------------------------------------------------------------------------------------------- */
#define SONAR_PYRAMID_MAX_LEVELS 16
enum { SONAR_RESAMPLE_BILINEAR = 0, SONAR_RESAMPLE_NEAREST_EXACT = 1, SONAR_RESAMPLE_AREA = 2 };

typedef struct SonarPyramidParams {
  float* out;
  const float* base;
  const float* levels[SONAR_PYRAMID_MAX_LEVELS];
  int32_t level_h[SONAR_PYRAMID_MAX_LEVELS];
  int32_t level_w[SONAR_PYRAMID_MAX_LEVELS];
  float weights[SONAR_PYRAMID_MAX_LEVELS];
  int64_t planes;
  int32_t H;
  int32_t W;
  int32_t n_levels;
  int32_t mode; /* SONAR_RESAMPLE_* */
  float base_scale;
  double* sums;       /* optional double[2], zero on entry: += {sum, sum of squares} of `out` (same launch) */
  double* sums_clear; /* optional double[2] the launch zeroes (next slot of the caller's ring) */
} SonarPyramidParams;

int sonar_pyramid_accum_f32(const SonarPyramidParams* params_host, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Perlin (grid cell == one pixel): out[b,c,y,x] = base[b,c,y,x]/div_fac + sum_it stencil(angles[it][c])[y,x]
 * replaces: PerlinOldNoiseGenerator.generate/perlin_noise/perlin_noise_tensor
 *                                                          py/noise_generation.py:353-493
 * angles[it] is (C, H+1, W+1) uniform in [0, 2pi); the stencil is shared by all B batch items.
 * ---------------------------------------------------------------------------------------------- */
#define SONAR_PERLIN_MAX_ITERS 8
typedef struct SonarPerlinParams {
  float* out;
  const float* base; /* (B,C,H,W) uniform [0,1); NULL = zeros */
  const float* angles[SONAR_PERLIN_MAX_ITERS];
  int32_t B;
  int32_t C;
  int32_t H;
  int32_t W;
  int32_t iterations;
  int32_t blend_mode; /* SONAR_BLEND_* */
  float div_fac;
  double* sums;       /* optional double[2], zero on entry: += {sum, sum of squares} of `out` */
  double* sums_clear; /* optional double[2] the launch zeroes */
} SonarPerlinParams;

int sonar_perlin_accum_f32(const SonarPerlinParams* params_host, void* stream);

/* ------------------------------------------------------------------------------------------------
 * RNG-fused noise synthesis: pyramid / Perlin noise (and their blend) computed element-wise from the Philox stream.
 * replaces: PyramidNoiseGenerator.generate                 py/noise_generation.py:621-649
 *           PerlinOldNoiseGenerator.generate               py/noise_generation.py:478-493
 *           BlendedNoise noise_sampler (scalar weight)     py/noise.py:1391-1405
 * out[i] = blend(mode, A(i), B(i), blend_t) (or A(i) when b.kind == NONE) for the elements [begin, begin + n) of a
 * dense (B_total, C, H, W) tensor of numel_total elements; every full-size draw of the graph is a torch draw of
 * numel_total values (geometry: grid_blocks from sonar_philox_policy) at its own generator offset and never touches
 * memory. A term:
 *   PYRAMID: base normal draw (base_offset) * base_scale + sum_l weights[l] * resample(level l); levels[l] != NULL is a
 *            materialised coarse level (local planes, level_h[l], level_w[l]); levels[l] == NULL (l == full_level, at
 *            most one) is the full-size level, a normal draw at level_offset[l]. mode: BILINEAR or NEAREST_EXACT.
 *   PERLIN : uniform [uniform_from, uniform_to) base draw / div_fac + sum_it tables[it][c, y, x], tables from
 *            sonar_perlin_tables_f32 (the stencil of iteration `it`, (C, H, W), shared by the batch).
 * sonar_perlin_tables_f32: tables[it][c,y,x] = 2 x 2 gradient stencil of pixel (y, x) whose corner angles are the
 * elements of a uniform [0, 2 pi) draw of C (H+1) (W+1) values at offsets[it] (regenerated in registers).
 * ---------------------------------------------------------------------------------------------- */
#define SONAR_MIX_MAX_TABLES 4
enum { SONAR_TERM_NONE = 0, SONAR_TERM_PYRAMID = 1, SONAR_TERM_PERLIN = 2 };
typedef struct SonarMixTerm {
  int32_t kind;       /* SONAR_TERM_* */
  int32_t n_levels;   /* PYRAMID */
  int32_t mode;       /* PYRAMID: SONAR_RESAMPLE_BILINEAR / _NEAREST_EXACT */
  int32_t full_level; /* PYRAMID: index of the full-size level, -1 = none */
  int32_t iterations; /* PERLIN */
  float base_scale;   /* PYRAMID */
  float div_fac;      /* PERLIN */
  float uniform_from; /* PERLIN: range of the base draw */
  float uniform_to;
  uint64_t base_offset;
  uint64_t level_offset[SONAR_PYRAMID_MAX_LEVELS];
  const float* levels[SONAR_PYRAMID_MAX_LEVELS];
  int32_t level_h[SONAR_PYRAMID_MAX_LEVELS];
  int32_t level_w[SONAR_PYRAMID_MAX_LEVELS];
  float weights[SONAR_PYRAMID_MAX_LEVELS];
  const float* tables[SONAR_MIX_MAX_TABLES];
} SonarMixTerm;
typedef struct SonarNoiseMixParams {
  float* out;
  int64_t n;           /* local elements (whole planes) */
  int64_t begin;       /* element offset of the local slice in the global tensor (a multiple of H * W) */
  int64_t numel_total; /* < 2^32; H * W <= 2^24, fewer than 2^24 planes, W and C < 2^16 */
  int32_t C;
  int32_t H;
  int32_t W;
  int32_t blend_mode; /* SONAR_BLEND_* */
  float blend_t;
  uint32_t grid_blocks;
  uint64_t seed;
  SonarMixTerm a;
  SonarMixTerm b;
  double* sums;       /* optional double[2], zero on entry: += {sum, sum of squares} of `out` */
  double* sums_clear; /* optional double[2] the launch zeroes */
} SonarNoiseMixParams;
int sonar_noise_mix_f32(const SonarNoiseMixParams* params_host, void* stream);
int sonar_perlin_tables_f32(float* const* tables_host, const uint64_t* offsets_host, const uint32_t* grid_blocks_host,
                            int32_t n_tables, uint64_t seed, int32_t C, int32_t H, int32_t W, int32_t blend_mode, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Element-wise combinators.
 * replaces: BLENDING_MODES                                 py/utils.py:17-21
 *           BlendedNoise noise_sampler                     py/noise.py:1391-1405
 *           CompositeNoise noise_sampler                   py/noise.py:524-531
 *           MixedNoiseGenerator.generate accumulate        py/noise_generation.py:240-249
 *           PowerLawNoiseGenerator.generate                py/noise_generation.py:775-786
 *           WaveletNoiseGenerator.generate (octave noise)  py/noise_generation.py:2204-2327
 *           normalize_to_scale                             py/utils.py:452-470
 * blend: out = mode(a, b, t) with t = t_tensor[i] if t_tensor else t_scalar. `out` may alias a or b.
 * axpby: out = a*alpha + b*beta (b may be NULL).
 * composite: out = dst*(1-mask) + src*mask, mask (batch,1,H,W) broadcast over channels.
 * `sums` (optional double[2], zero on entry) receives {sum, sum of squares} of `out`, reduced in the same
 * launch: the producer of a tensor hands scale_noise its statistics, saving the separate read pass.
 * `sums_clear` (optional) is another double[2] the launch zeroes, so a caller can recycle a ring of
 * slots (slot i's producer clears slot i+1) without a memset per launch.
 * item_*: reductions over everything but the leading dim; scratch from sonar_item_range_scratch_bytes.
 * ---------------------------------------------------------------------------------------------- */
int sonar_blend_f32(const float* a, const float* b, const float* t_tensor, float t_scalar, float* out, int64_t n,
                    int mode, double* sums, double* sums_clear, void* stream);
int sonar_axpby_f32(const float* a, float alpha, const float* b, float beta, float* out, int64_t n, double* sums,
                    double* sums_clear, void* stream);
/* out = ((x + pre_add) * mul) + post_add, each step rounded (UniformNoiseGenerator.generate,
 * py/noise_generation.py:508-514) */
int sonar_affine_f32(const float* x, float* out, int64_t n, float pre_add, float mul, float post_add, void* stream);
/* out = x / divisor, IEEE division (Tensor.div_(scalar); WaveletNoiseGenerator.generate, py/noise_generation.py:2325) */
int sonar_div_scalar_f32(const float* x, float* out, int64_t n, float divisor, void* stream);
int sonar_composite_f32(const float* dst, const float* src, const float* mask, float* out, int64_t batch,
                        int64_t channels, int64_t hw, void* stream);
int sonar_powerlaw_f32(const float* x, float* out, int64_t n, float alpha, int use_sign, void* stream);
int sonar_item_range_scratch_bytes(int64_t items);
int sonar_item_div_max_f32(const float* x, float* out, int64_t items, int64_t per_item, int use_abs, void* scratch,
                           void* stream);
int sonar_item_minmax_rescale_f32(const float* x, float* out, int64_t items, int64_t per_item, float target_min,
                                  float target_max, float eps, void* scratch, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Spectral shaping: [rfft2 ->] gain mask -> irfft2, one CTA per (H, W) plane, spectrum resident in
 * shared memory (or in `scratch` when H*(W/2+1) complex64 exceeds it).
 * replaces: PowerNoiseItem sampler (irfft2 of a shaped half spectrum, optional rfft2 front end)
 *                                                          py/nodes/powernoise.py:355-366
 *           OneFNoiseGenerator.generate (fftn -> gain -> ifftn.real)
 *                                                          py/noise_generation.py:737-759
 *           GreenTestNoiseGenerator.generate               py/noise_generation.py:694-704
 *           FreeU-Extreme ffilter (same op, "next" caller) py/nodes/freeu_extreme.py:10-29
 * Exactly one of in_real (planes,H,W) / in_spec (planes,H,W/2+1 complex64, interleaved) is set.
 * mask is a real (H, W/2+1) gain shared by all planes (NULL = 1). out = out_scale * result:
 * 1/sqrt(H*W) for an "ortho" irfft2 alone, 1/(H*W) for a forward+inverse round trip.
 * H and W may be any positive sizes (mixed radix 4/2 + direct prime radices).
 * ---------------------------------------------------------------------------------------------- */
#define SONAR_FFT_MAX_FACTORS 24
#define SONAR_SPECTRAL_MAX_SEGMENTS 32
typedef struct SonarSpectralParams {
  float* out;
  const float* in_real;
  const float* in_spec; /* complex64 as interleaved (re, im) */
  const float* mask;
  float* scratch;       /* sonar_spectral_scratch_bytes(H, W) bytes, may be NULL when that is 0 */
  int64_t planes;
  int32_t H;
  int32_t W;
  float out_scale;
  double* sums;       /* optional double[2], zero on entry: += {sum, sum of squares} of `out` */
  double* sums_clear; /* optional double[2] the launch zeroes */
  /* > 0: `sums` is double[2 * ceil(planes / sums_segment_planes)], one {sum, sum of squares} pair per run of
   * sums_segment_planes consecutive planes (at most SONAR_SPECTRAL_MAX_SEGMENTS runs) -- several noise samples (each normalised on its own) in ONE launch. */
  int64_t sums_segment_planes;
  /* in_real == NULL and in_spec == NULL: the half spectrum is torch.randn(complex64) REGENERATED from the Philox
   * stream inside the first column stage (never stored). Plane 0 starts at float `philox_begin` of a draw of
   * `philox_numel_total` floats (2 per complex element; real and imaginary part ~ N(0, philox_std^2), 1/sqrt(2) for
   * torch's complex normal). Costs one Philox call per float: meant for draws of a single ATen row
   * (numel_total <= 256 * grid_blocks), where torch's own kernel uses one lane per call as well. */
  uint64_t philox_seed;
  uint64_t philox_offset;
  uint32_t philox_grid_blocks;
  float philox_std;
  int64_t philox_begin;
  int64_t philox_numel_total;
} SonarSpectralParams;

int64_t sonar_spectral_scratch_bytes(int H, int W);
int sonar_spectral_filter_f32(const SonarSpectralParams* params_host, void* stream);

/* Channel-correlation mixer of the power-noise family: out[b, c, p] = sum_k mixer[c, k] * in[b, k, p] for a dense
 * (batch, channels, hw) tensor; `mixer` is the (channels x channels) row-major matrix ON THE DEVICE.
 * replaces: ChannelMixer.apply                             py/nodes/powernoise.py:94-101
 * channels <= SONAR_MIXER_SMALL_MAX and mixer_host != NULL (the same matrix in host memory): per-pixel mat-vec
 * with the matrix passed by value. Otherwise a tiled fp32 GEMM; with `mixer_packed` (device, 16-byte aligned; hw a
 * multiple of 4) its tiles are streamed by bulk-async copies behind mbarriers. mixer_packed is the matrix pre-tiled
 * as [ceil(C/64)][ceil(C/16)][64][20] floats, zero padded (entry [tm][kb][m][k] = mixer[64 tm + m][16 kb + k], k < 16),
 * sonar_channel_mix_packed_floats(C) floats in all. `in` must not alias `out`.
 * sums / sums_clear: as SonarSpectralParams (moments of the output accumulated into sums; may be NULL). */
#define SONAR_MIXER_SMALL_MAX 8
int64_t sonar_channel_mix_packed_floats(int32_t channels);
int sonar_channel_mix_f32(const float* in, float* out, const float* mixer, const float* mixer_host, const float* mixer_packed,
                          int64_t batch, int32_t channels, int64_t hw, double* sums, double* sums_clear, void* stream);

/* Host-only query (no GPU work): how sonar_spectral_filter_f32 would run `planes` planes of (H, W).
 * batched = 1: the in-place shared-memory kernel with `group` planes per CTA pass, `threads` threads per CTA,
 * `ctas_per_sm` resident CTAs, `grid` CTAs, `smem_bytes` of dynamic shared memory and the listed radices for the
 * column axis (length H) and the row axis (length W / 2), in stage order. batched = 0: the generic
 * warp-per-transform kernel (odd W, other prime factors, planes too large for shared memory). */
#define SONAR_SPECTRAL_PLAN_MAX_STAGES 16
typedef struct SonarSpectralPlanInfo {
  int32_t batched;
  int32_t group;
  int32_t threads;
  int32_t ctas_per_sm;
  int64_t grid;
  int64_t smem_bytes;
  int32_t n_col_stages;
  int32_t n_row_stages;
  int32_t col_radix[SONAR_SPECTRAL_PLAN_MAX_STAGES];
  int32_t row_radix[SONAR_SPECTRAL_PLAN_MAX_STAGES];
  int32_t cluster;  /* 1: the plane lives in the distributed shared memory of a 2-CTA cluster (e.g. 256x256) */
  int32_t pad_;
} SonarSpectralPlanInfo;
int sonar_spectral_plan(int H, int W, int64_t planes, int real_input, SonarSpectralPlanInfo* info_host);

/* ------------------------------------------------------------------------------------------------
 * 2-D DWT levels + wavelet-CFG combine.
 * replaces: Wavelet.forward / Wavelet.inverse               py/wavelet_functions.py:81-105
 *             (-> pytorch_wavelets.DWTForward / DWTInverse, [upstream], restated)
 *           wavelet_scaling / wavelet_blend                 py/wavelet_functions.py:193-238
 *           WaveletCFG.wavelet_cfg / process_output         py/wavelet_cfg.py:729-791
 * One analysis call = one decomposition level: in (planes,H,W) -> ll (planes,h,w) and
 * hi (planes,3,h,w) with h = (H+L-1)/2. in_b (optional) is subtracted on load, so level 1 can
 * transform (cond - uncond) straight from the fp32 inputs (in_is_f32). Coefficients are fp64
 * (use_f64, the reference's high_precision_mode default) or fp32.
 * One synthesis call = one reconstruction level over 1 or 2 coefficient sets, each band multiplied
 * by scales[set][ll, hi0, hi1, hi2] on load. Intermediate levels write `out` (planes, 2h-L+2,
 * 2w-L+2) in the coefficient type; the final level writes fp32 `out_f32` cropped to
 * (crop_h, crop_w):  out_f32 = x_scale*x + (float)(recon_sign*(recon + addend_scale*addend)).
 * ---------------------------------------------------------------------------------------------- */
#define SONAR_DWT_MAX_TAPS 40
enum {
  SONAR_DWT_MODE_SYMMETRIC = 0,
  SONAR_DWT_MODE_ZERO = 1,
  SONAR_DWT_MODE_REFLECT = 2,
  SONAR_DWT_MODE_PERIODIC = 3,
  SONAR_DWT_MODE_PERIODIZATION = 4 /* non-expansive: h = ceil(H / 2); analysis only, synthesis = sonar_dwt2_synthesis_per */
};

typedef struct SonarWaveletFilters {
  int32_t length;
  double dec_lo[SONAR_DWT_MAX_TAPS];
  double dec_hi[SONAR_DWT_MAX_TAPS];
  double rec_lo[SONAR_DWT_MAX_TAPS];
  double rec_hi[SONAR_DWT_MAX_TAPS];
} SonarWaveletFilters;

typedef struct SonarDwtAnalysisParams {
  const void* in_a;
  const void* in_b; /* optional, subtracted */
  void* ll;
  void* hi;
  int64_t planes;
  int32_t H;
  int32_t W;
  int32_t in_stride_h; /* rows per plane of the input buffer (>= H; 0 = H); row length is W */
  int32_t h;
  int32_t w;
  int32_t mode;      /* SONAR_DWT_MODE_* */
  int32_t in_is_f32; /* inputs are fp32 tensors (level 1) */
  int32_t use_f64;
  SonarWaveletFilters filters;
} SonarDwtAnalysisParams;

typedef struct SonarDwtSynthesisParams {
  const void* ll[2];
  const void* hi[2];
  int32_t ll_rows[2]; /* actual rows / cols of the ll buffers (>= h, w) */
  int32_t ll_cols[2];
  double scales[2][4];
  int32_t n_sets;
  int64_t planes;
  int32_t h;
  int32_t w;
  void* out;      /* intermediate level output, coefficient type */
  float* out_f32; /* final level output */
  int32_t crop_h;
  int32_t crop_w;
  const float* addend;
  float addend_scale;
  const float* x;
  float x_scale;
  float recon_sign;
  int32_t use_f64;
  SonarWaveletFilters filters;
} SonarDwtSynthesisParams;

int sonar_dwt_coeff_len(int n, int filter_len);
int sonar_dwt2_analysis(const SonarDwtAnalysisParams* params_host, void* stream);
int sonar_dwt2_synthesis(const SonarDwtSynthesisParams* params_host, void* stream);
/* One reconstruction level of the non-expansive "periodization" transform (WaveletFilteredNoiseGenerator's default
 * mode, py/noise_generation.py:1908-2032): one coefficient set (n_sets = 1), out (planes, 2h, 2w) in the
 * coefficient type; ll may be one row / column larger than h, w (ll_rows / ll_cols). */
int sonar_dwt2_synthesis_per(const SonarDwtSynthesisParams* params_host, void* stream);

/* The whole wavelet-CFG combine in ONE launch (one CTA per plane, every coefficient of every level in
 * shared memory):  value = in_a - in_b (in_b optional), coefficients scaled per band (scale_ll for the
 * coarsest approximation, scale_hi[level][orientation], levels fine -> coarse), reconstruction cropped
 * to (H, W):   out = x_scale*x + (float)(recon_sign*(recon + addend_scale*addend)).
 * replaces: WaveletCFG.wavelet_cfg + process_output (2 forward + 1 inverse transforms and ~8 coefficient
 *           passes upstream)                               py/wavelet_cfg.py:729-791
 * sonar_wcfg_fused_smem_bytes returns the shared memory the call needs, or 0 when the coefficient
 * pyramid does not fit one SM (the caller then uses the per-level entry points above). */
#define SONAR_WCFG_MAX_LEVELS 8
typedef struct SonarWcfgFusedParams {
  const float* in_a;
  const float* in_b;
  float* out;
  const float* addend;
  float addend_scale;
  const float* x;
  float x_scale;
  float recon_sign;
  int64_t planes;
  int32_t H;
  int32_t W;
  int32_t levels;
  int32_t mode; /* SONAR_DWT_MODE_* */
  int32_t use_f64;
  double scale_ll;
  double scale_hi[SONAR_WCFG_MAX_LEVELS][3];
  SonarWaveletFilters filters;
} SonarWcfgFusedParams;

int64_t sonar_wcfg_fused_smem_bytes(int H, int W, int filter_len, int levels, int use_f64);
int sonar_wcfg_fused(const SonarWcfgFusedParams* params_host, void* stream);

/* ------------------------------------------------------------------------------------------------
 * FreeU-Extreme epilogue around the spectral filter ("next" row).
 * replaces: FreeUExtremeConfig.get_scale (hidden mean: channel mean of the activation, rescaled to
 *           [0, 1] per batch item, -> 1 + (scale - 1) * mean)   py/nodes/freeu_extreme.py:183-194
 *           FreeUExtremeConfig.apply (filtered slice * scale written back over the channel slice,
 *           optionally through BLENDING_MODES)                  py/nodes/freeu_extreme.py:203-227
 * The filter itself (ffilter, :10-29) is sonar_spectral_filter_f32 with in_real set.
 * sonar_freeu_hidden_mean_f32: hidden (batch, hw) = mean over channels of h (batch, channels, hw);
 * range_partial (sonar_freeu_range_bytes(batch) bytes) receives per-item (min, max) partials.
 * sonar_freeu_apply_f32: x[b][slice_offset + c][p] = blend(x, src * sc, blend) in place, src =
 * filtered[b][c][p] (dense slice) or x itself when filtered is NULL, sc = scale, or with `hidden`
 * 1 + scale_minus_one * (hidden[b][p] - min_b) / (max_b - min_b). use_blend = 0 stores src * sc.
 * ---------------------------------------------------------------------------------------------- */
typedef struct SonarFreeuParams {
  float* x;
  const float* filtered;
  const float* hidden;       /* NULL: plain scalar scale */
  const void* hidden_range;  /* from sonar_freeu_hidden_mean_f32, required with hidden */
  int64_t batch;
  int64_t channels;
  int64_t hw;
  int64_t slice_offset;
  int64_t slice_channels;
  float scale;
  float scale_minus_one;
  float blend;
  int32_t blend_mode; /* SONAR_BLEND_* */
  int32_t use_blend;
} SonarFreeuParams;

int64_t sonar_freeu_range_bytes(int64_t batch);
int sonar_freeu_hidden_mean_f32(const float* h, float* hidden, void* range_partial, int64_t batch, int64_t channels,
                                int64_t hw, void* stream);
int sonar_freeu_apply_f32(const SonarFreeuParams* params_host, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Peer-memory exchange of the global scale_noise statistics (one node, NVLink 5 / NVSwitch).
 * replaces: the whole-batch reduction inside scale_noise when the batch is sharded over GPUs
 *                                                          py/utils.py:100-106
 * Each rank allocates one mailbox, exports its CUDA IPC handle (64 bytes, exchanged by the host
 * over torch.distributed), opens every peer's handle, then per exchange calls
 * sonar_peer_publish_sums (stores its two partial sums into every rank's mailbox over NVLink) and
 * launches the step with peer_world / peer_mailbox / peer_epoch set. Epochs must increase by 1 per
 * exchange on every rank; two parity slots make the scheme race free without a barrier.
 * ---------------------------------------------------------------------------------------------- */
#define SONAR_PEER_MAX_RANKS 8
int sonar_peer_alloc(void** mailbox_out_host);
int sonar_peer_free(void* mailbox);
int sonar_peer_get_handle(const void* mailbox, unsigned char* handle64_host);
int sonar_peer_open_handle(const unsigned char* handle64_host, void** mapped_out_host);
int sonar_peer_close_handle(void* mapped);
/* scale_noise whose statistics are the total of the `world` partial sums of `epoch` in `mailbox` */
int sonar_scale_noise_peers_f32(const float* x, float* out, int64_t n, const double* mailbox, int world, double epoch,
                                int64_t count, float factor, float threshold_std_devs, void* stream);
int sonar_peer_publish_sums(void* const* mailboxes_host, int rank, int world, const double* local_sums, double epoch,
                            void* stream);
/* In-place sum over ranks of a small table of doubles (n <= SONAR_PEER_TABLE_MAX), e.g. the (K, 2)
 * look-ahead statistics of all noise draws of a sampler run: one launch per rank, stores over NVLink,
 * device-side wait, ranks added in rank order (identical bits everywhere). Same epoch rule as above. */
#define SONAR_PEER_TABLE_MAX 512
int sonar_peer_allreduce_table(void* const* mailboxes_host, int rank, int world, double* table, int n, double epoch,
                               void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SONAR_B200_H_ */
