// Streaming noise-synthesis kernels: fused multi-level resample-and-accumulate (pyramid family),
// Perlin 2x2 gradient stencil, blends / composites, power-law shaping and per-item range reductions.
//
// Reference (py/noise_generation.py): PyramidNoiseGenerator.generate :621-649,
// HighresPyramidNoiseGenerator.generate :539-564, PyramidOldNoiseGenerator.generate :579-606,
// PerlinOldNoiseGenerator :289-493, PowerLawNoiseGenerator.generate :775-786;
// (py/noise.py) CompositeNoise :524-531, BlendedNoise :1391-1405; (py/utils.py) normalize_to_scale
// :452-470. Resampling follows ATen's area_pixel_compute_source_index (ATen/native/UpSample.h:289-312).
#include "common.cuh"
#include "../../include/sonar_b200.h"

namespace sonar {

// ---------------------------------------------------------------------------------------------
// resampling taps (align_corners = False), identical index math to ATen
// ---------------------------------------------------------------------------------------------
struct LinTap {
  int i0, i1;
  float w0, w1;
};

__device__ __forceinline__ LinTap linear_tap(int dst, int in_size, float scale) {
  float src = scale * ((float)dst + 0.5f) - 0.5f;
  if (src < 0.0f) src = 0.0f;
  int i0 = (int)src;
  if (i0 > in_size - 1) i0 = in_size - 1;
  float l1 = src - (float)i0;
  l1 = fminf(fmaxf(l1, 0.0f), 1.0f);
  LinTap t;
  t.i0 = i0;
  t.i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  t.w1 = l1;
  t.w0 = 1.0f - l1;
  return t;
}

__device__ __forceinline__ int nearest_exact_idx(int dst, int in_size, float scale) {
  const int i = (int)floorf(((float)dst + 0.5f) * scale);
  return i < in_size - 1 ? i : in_size - 1;
}

__device__ __forceinline__ int pool_start(int dst, int out_size, int in_size) {
  return (dst / out_size) * in_size + ((dst % out_size) * in_size) / out_size;
}
__device__ __forceinline__ int pool_end(int dst, int out_size, int in_size) {
  return 1 + ((dst + 1) * in_size - 1) / out_size;
}

__device__ __forceinline__ float sample_level(const float* __restrict__ src, int lh, int lw, int H, int W, int y, int x,
                                              int mode) {
  if (lh == H && lw == W) return src[(int64_t)y * lw + x];  // identity for every mode
  if (mode == SONAR_RESAMPLE_BILINEAR) {
    const LinTap ty = linear_tap(y, lh, (float)lh / (float)H);
    const LinTap tx = linear_tap(x, lw, (float)lw / (float)W);
    const float* r0 = src + (int64_t)ty.i0 * lw;
    const float* r1 = src + (int64_t)ty.i1 * lw;
    return ty.w0 * (tx.w0 * r0[tx.i0] + tx.w1 * r0[tx.i1]) + ty.w1 * (tx.w0 * r1[tx.i0] + tx.w1 * r1[tx.i1]);
  }
  if (mode == SONAR_RESAMPLE_NEAREST_EXACT) {
    const int sy = nearest_exact_idx(y, lh, (float)lh / (float)H);
    const int sx = nearest_exact_idx(x, lw, (float)lw / (float)W);
    return src[(int64_t)sy * lw + sx];
  }
  // area == adaptive average pooling
  const int y0 = pool_start(y, H, lh), y1 = pool_end(y, H, lh);
  const int x0 = pool_start(x, W, lw), x1 = pool_end(x, W, lw);
  float acc = 0.0f;
  for (int yy = y0; yy < y1; ++yy)
    for (int xx = x0; xx < x1; ++xx) acc += src[(int64_t)yy * lw + xx];
  return acc / (float)(y1 - y0) / (float)(x1 - x0);
}

// out[p, y, x] = base_scale * base[p, y, x] + sum_i weight_i * resample(level_i[p])[y, x]
// One thread produces 4 consecutive x (float4 store) when W % 4 == 0, else one pixel.
template <int VEC>
__global__ void __launch_bounds__(kBlock)
pyramid_accum_kernel(SonarPyramidParams p) {
  const int Wv = p.W / VEC;
  const int64_t total = p.planes * (int64_t)p.H * Wv;
  float ms = 0.0f, mss = 0.0f;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int xv = (int)(idx % Wv);
    const int y = (int)((idx / Wv) % p.H);
    const int64_t plane = idx / ((int64_t)Wv * p.H);
    const int64_t o = (plane * p.H + y) * (int64_t)p.W + (int64_t)xv * VEC;
    float acc[VEC];
    if (p.base != nullptr) {
      if (VEC == 4) {
        const float4 b = ld4_stream(p.base + o);
        acc[0] = b.x;
        acc[1 % VEC] = b.y;
        acc[2 % VEC] = b.z;
        acc[3 % VEC] = b.w;
      } else {
        acc[0] = p.base[o];
      }
      if (p.base_scale != 1.0f) {
#pragma unroll
        for (int v = 0; v < VEC; ++v) acc[v] *= p.base_scale;
      }
    } else {
#pragma unroll
      for (int v = 0; v < VEC; ++v) acc[v] = 0.0f;
    }
    for (int l = 0; l < p.n_levels; ++l) {
      const int lh = p.level_h[l], lw = p.level_w[l];
      const float* src = p.levels[l] + plane * (int64_t)lh * lw;
      const float wgt = p.weights[l];
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        const float s = sample_level(src, lh, lw, p.H, p.W, y, xv * VEC + v, p.mode);
        // reference: noise += upsampled.mul_(discount ** i)  -> separate mul then add
        acc[v] = acc[v] + __fmul_rn(s, wgt);
      }
    }
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      ms += acc[v];
      mss += acc[v] * acc[v];
    }
    if (VEC == 4) {
      st4_stream(p.out + o, make_float4(acc[0], acc[1 % VEC], acc[2 % VEC], acc[3 % VEC]));
    } else {
      p.out[o] = acc[0];
    }
  }
  commit_moments(p.sums, p.sums_clear, ms, mss);
}

// Table-driven variant (bilinear / nearest-exact): the x taps of every level depend only on x, so
// each CTA computes them once into shared memory; a warp then walks whole rows, computing the y taps
// once per row. Per pixel and level that leaves 3 table reads, 4 cached loads and 6 FMAs instead of
// the float divisions and tap arithmetic of the generic path.
// VEC = 4: each lane owns 4 consecutive pixels (W % 4 == 0, 16-byte aligned rows): base, full-size
// level and output move as float4; VEC = 1: any width / alignment.
//
// Bilinear levels are evaluated rows-first: for an output row the warp blends the two source rows of a
// coarse level ONCE into a per-warp shared-memory row v[i] = wy0*r0[i] + wy1*r1[i] (lw values), and every
// pixel then needs two shared-memory reads and one lerp per level instead of four gathers and three
// lerps. ATen evaluates columns-first; the two orders agree to ~1 ulp (parity bar for interpolation:
// 1e-5). Nearest-exact levels copy the single source row.
struct PyrTap {  // one x (or y) tap: source indices and the weight of the second one
  unsigned short i0, i1;
  float w1;
};

template <int VEC>
__global__ void __launch_bounds__(kBlock)
pyramid_rows_kernel(SonarPyramidParams p, int row_pitch) {
  extern __shared__ __align__(16) unsigned char pyr_smem[];
  const int W = p.W, H = p.H, NL = p.n_levels;
  // x taps [n_levels][W] (VEC = 4: stored [v][W/4] so the 32 lanes of a warp read 32 consecutive
  // entries), y taps [n_levels][H], then one blended-row scratch [n_levels][row_pitch] per warp
  PyrTap* xtab = reinterpret_cast<PyrTap*>(pyr_smem);
  PyrTap* ytab = xtab + NL * W;
  float* vrows = reinterpret_cast<float*>(ytab + NL * H) + (size_t)(threadIdx.x >> 5) * NL * row_pitch;
  const bool bilinear = p.mode == SONAR_RESAMPLE_BILINEAR;
  const int Wq = W / VEC;
  for (int i = threadIdx.x; i < NL * (W + H); i += blockDim.x) {
    const bool is_x = i < NL * W;
    const int j = is_x ? i : i - NL * W, n = is_x ? W : H;
    const int l = j / n, pos = j - l * n;
    const int ln = is_x ? p.level_w[l] : p.level_h[l];
    PyrTap t;
    if (bilinear) {
      const LinTap lt = linear_tap(pos, ln, (float)ln / (float)n);
      t.i0 = (unsigned short)lt.i0;
      t.i1 = (unsigned short)lt.i1;
      t.w1 = lt.w1;
    } else {
      t.i0 = t.i1 = (unsigned short)nearest_exact_idx(pos, ln, (float)ln / (float)n);
      t.w1 = 0.0f;
    }
    if (is_x)
      xtab[l * W + (VEC == 4 ? (pos & 3) * Wq + (pos >> 2) : pos)] = t;
    else
      ytab[j] = t;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const unsigned n_rows = (unsigned)(p.planes * (int64_t)H);  // < 2^31 (checked by the launcher)
  float ms = 0.0f, mss = 0.0f;
  for (unsigned row = blockIdx.x * warps_per_block + (threadIdx.x >> 5); row < n_rows; row += gridDim.x * warps_per_block) {
    const unsigned plane_u = row / (unsigned)H;
    const int y = (int)(row - plane_u * (unsigned)H);
    const int64_t plane = plane_u;
    const int64_t obase = (int64_t)row * W;
    // ---- blend the two source rows of every resampled level for this output row ----
    for (int l = 0; l < NL; ++l) {
      const int lh = p.level_h[l], lw = p.level_w[l];
      if (lh == H && lw == W) continue;  // identity level: read straight from global below
      const float* src = p.levels[l] + plane * (int64_t)lh * lw;
      float* v = vrows + l * row_pitch;
      const PyrTap ty = ytab[l * H + y];
      const float* r0 = src + (int)ty.i0 * lw;
      const float* r1 = src + (int)ty.i1 * lw;
      const float wy0 = 1.0f - ty.w1;
#pragma unroll 1
      for (int i = lane; i < lw; i += 32) v[i] = bilinear ? wy0 * __ldg(r0 + i) + ty.w1 * __ldg(r1 + i) : __ldg(r0 + i);
    }
    __syncwarp();
    for (int x0 = lane * VEC; x0 < W; x0 += 32 * VEC) {
      float acc[VEC];
      if (p.base != nullptr) {
        if (VEC == 4) {
          const float4 b4 = ld4_stream(p.base + obase + x0);
          acc[0] = b4.x; acc[1 % VEC] = b4.y; acc[2 % VEC] = b4.z; acc[3 % VEC] = b4.w;
        } else {
          acc[0] = __ldg(p.base + obase + x0);
        }
        if (p.base_scale != 1.0f) {
#pragma unroll
          for (int v = 0; v < VEC; ++v) acc[v] *= p.base_scale;
        }
      } else {
#pragma unroll
        for (int v = 0; v < VEC; ++v) acc[v] = 0.0f;
      }
      const PyrTap* tx = xtab + (VEC == 4 ? (x0 >> 2) : x0);
#pragma unroll 1
      for (int l = 0; l < NL; ++l, tx += W) {
        const int lh = p.level_h[l], lw = p.level_w[l];
        const float wgt = p.weights[l];
        if (lh == H && lw == W) {  // identity level: straight copy-accumulate
          const float* r = p.levels[l] + plane * (int64_t)lh * lw + (int64_t)y * W + x0;
          if (VEC == 4) {
            const float4 s4 = ld4_stream(r);
            acc[0] = acc[0] + __fmul_rn(s4.x, wgt);
            acc[1 % VEC] = acc[1 % VEC] + __fmul_rn(s4.y, wgt);
            acc[2 % VEC] = acc[2 % VEC] + __fmul_rn(s4.z, wgt);
            acc[3 % VEC] = acc[3 % VEC] + __fmul_rn(s4.w, wgt);
          } else {
            acc[0] = acc[0] + __fmul_rn(__ldg(r), wgt);
          }
          continue;
        }
        const float* v = vrows + l * row_pitch;
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
          const PyrTap t = tx[VEC == 4 ? k * Wq : 0];
          float sv = v[t.i0];
          if (bilinear) sv = (1.0f - t.w1) * sv + t.w1 * v[t.i1];
          acc[k] = acc[k] + __fmul_rn(sv, wgt);
        }
      }
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        ms += acc[v];
        mss += acc[v] * acc[v];
      }
      if (VEC == 4)
        st4_stream(p.out + obase + x0, make_float4(acc[0], acc[1 % VEC], acc[2 % VEC], acc[3 % VEC]));
      else
        p.out[obase + x0] = acc[0];
    }
    __syncwarp();  // the next row overwrites the blended rows
  }
  commit_moments(p.sums, p.sums_clear, ms, mss);
}

// ---------------------------------------------------------------------------------------------
// Perlin: grid cell == 1 pixel, so every output is the blend of four corner gradients dotted with
// (+-0.5, +-0.5) (closed form verified against perlin_noise).
// The (C,H,W) stencil result is shared by the whole batch: compute once, add to every item.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float perlin_value(const float* __restrict__ ang, int H, int W, int y, int x, int mode) {
  const int64_t gw = W + 1;
  const float* a0 = ang + (int64_t)y * gw + x;
  const float* a1 = a0 + gw;
  float s00, c00, s01, c01, s10, c10, s11, c11;
  sincosf(a0[0], &s00, &c00);
  sincosf(a0[1], &s01, &c01);
  sincosf(a1[0], &s10, &c10);
  sincosf(a1[1], &s11, &c11);
  // positions = (0.5, 0.5); offsets (1,0), (0,1), (1,1) subtracted for the other corners
  const float d00 = __fadd_rn(__fmul_rn(c00, 0.5f), __fmul_rn(s00, 0.5f));
  const float d01 = __fadd_rn(__fmul_rn(c01, -0.5f), __fmul_rn(s01, 0.5f));
  const float d10 = __fadd_rn(__fmul_rn(c10, 0.5f), __fmul_rn(s10, -0.5f));
  const float d11 = __fadd_rn(__fmul_rn(c11, -0.5f), __fmul_rn(s11, -0.5f));
  const float row0 = blend<float>(mode, d00, d01, 0.5f);
  const float row1 = blend<float>(mode, d10, d11, 0.5f);
  return blend<float>(mode, row0, row1, 0.5f);
}

// corner gradient . offset for the four corners of pixel (y, x), from precomputed sin / cos
__device__ __forceinline__ float perlin_from_corners(float s00, float c00, float s01, float c01, float s10, float c10,
                                                     float s11, float c11, int mode) {
  const float d00 = __fadd_rn(__fmul_rn(c00, 0.5f), __fmul_rn(s00, 0.5f));
  const float d01 = __fadd_rn(__fmul_rn(c01, -0.5f), __fmul_rn(s01, 0.5f));
  const float d10 = __fadd_rn(__fmul_rn(c10, 0.5f), __fmul_rn(s10, -0.5f));
  const float d11 = __fadd_rn(__fmul_rn(c11, -0.5f), __fmul_rn(s11, -0.5f));
  const float row0 = blend<float>(mode, d00, d01, 0.5f);
  const float row1 = blend<float>(mode, d10, d11, 0.5f);
  return blend<float>(mode, row0, row1, 0.5f);
}

// One thread per (c, y, VEC consecutive x): the 2 x (VEC+1) corner angles are turned into sin/cos once
// (10 sincosf for 4 pixels instead of 16), the stencil value is kept in registers and added to every
// batch item with float4 traffic. ITERS is a template parameter so pv[][] stays in registers.
template <int VEC, int ITERS>
__global__ void __launch_bounds__(kBlock)
perlin_accum_kernel(SonarPerlinParams p) {
  const int Wv = p.W / VEC;
  const int64_t chw = (int64_t)p.C * p.H * p.W;
  const int64_t total = (int64_t)p.C * p.H * Wv;
  const float inv_div = 1.0f / p.div_fac;
  float ms = 0.0f, mss = 0.0f;
  for (int64_t idx64 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx64 < total;
       idx64 += (int64_t)gridDim.x * blockDim.x) {
    const unsigned idx = (unsigned)idx64;  // C*H*W/VEC < 2^31 (checked by the launcher): 32-bit divisions
    const unsigned row = idx / (unsigned)Wv;
    const int x0 = (int)(idx - row * (unsigned)Wv) * VEC;
    const int c = (int)(row / (unsigned)p.H);
    const int y = (int)(row - (unsigned)c * (unsigned)p.H);
    float pv[ITERS][VEC];
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
      const int64_t gw = p.W + 1;
      const float* a0 = p.angles[it] + (int64_t)c * (p.H + 1) * gw + (int64_t)y * gw + x0;
      const float* a1 = a0 + gw;
      float s0[VEC + 1], c0[VEC + 1], s1[VEC + 1], c1[VEC + 1];
#pragma unroll
      for (int v = 0; v <= VEC; ++v) {
        sincosf(__ldg(a0 + v), &s0[v], &c0[v]);
        sincosf(__ldg(a1 + v), &s1[v], &c1[v]);
      }
#pragma unroll
      for (int v = 0; v < VEC; ++v)
        pv[it][v] = perlin_from_corners(s0[v], c0[v], s0[v + 1], c0[v + 1], s1[v], c1[v], s1[v + 1], c1[v + 1], p.blend_mode);
    }
    const int64_t o0 = (int64_t)c * p.H * p.W + (int64_t)y * p.W + x0;
#pragma unroll 4
    for (int b = blockIdx.y; b < p.B; b += gridDim.y) {  // grid.y splits the batch when C*H*W alone is too few CTAs
      const int64_t o = (int64_t)b * chw + o0;
      float v[VEC];
      if (p.base != nullptr) {
        if (VEC == 4) {
          const float4 b4 = ld4_stream(p.base + o);
          v[0] = b4.x; v[1 % VEC] = b4.y; v[2 % VEC] = b4.z; v[3 % VEC] = b4.w;
        } else {
          v[0] = __ldg(p.base + o);
        }
#pragma unroll
        for (int k = 0; k < VEC; ++k) v[k] = div_by(v[k], p.div_fac, inv_div);
      } else {
#pragma unroll
        for (int k = 0; k < VEC; ++k) v[k] = 0.0f;
      }
#pragma unroll
      for (int it = 0; it < ITERS; ++it)
#pragma unroll
        for (int k = 0; k < VEC; ++k) v[k] += pv[it][k];
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        ms += v[k];
        mss += v[k] * v[k];
      }
      if (VEC == 4)
        st4_stream(p.out + o, make_float4(v[0], v[1 % VEC], v[2 % VEC], v[3 % VEC]));
      else
        p.out[o] = v[0];
    }
  }
  commit_moments(p.sums, p.sums_clear, ms, mss);
}

// any iteration count (up to SONAR_PERLIN_MAX_ITERS): one pixel per thread
__global__ void __launch_bounds__(kBlock)
perlin_accum_generic_kernel(SonarPerlinParams p) {
  const int64_t chw = (int64_t)p.C * p.H * p.W;
  float ms = 0.0f, mss = 0.0f;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < chw;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(idx % p.W);
    const int y = (int)((idx / p.W) % p.H);
    const int c = (int)(idx / ((int64_t)p.W * p.H));
    float pv[SONAR_PERLIN_MAX_ITERS];
    for (int it = 0; it < p.iterations; ++it)
      pv[it] = perlin_value(p.angles[it] + (int64_t)c * (p.H + 1) * (p.W + 1), p.H, p.W, y, x, p.blend_mode);
    for (int b = 0; b < p.B; ++b) {
      const int64_t o = (int64_t)b * chw + idx;
      float v = p.base != nullptr ? p.base[o] / p.div_fac : 0.0f;
      for (int it = 0; it < p.iterations; ++it) v += pv[it];
      p.out[o] = v;
      ms += v;
      mss += v * v;
    }
  }
  commit_moments(p.sums, p.sums_clear, ms, mss);
}

// ---------------------------------------------------------------------------------------------
// element-wise: blend, axpby, composite, power law
// ---------------------------------------------------------------------------------------------
// VEC = 4: float4 accesses (all pointers 16-byte aligned, n % 4 handled by a scalar tail)
template <int VEC>
__global__ void __launch_bounds__(kBlock)
blend_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ t_tensor, float t_scalar,
             float* __restrict__ out, int64_t n, int mode, double* __restrict__ sums, double* __restrict__ sums_clear) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  float ms = 0.0f, mss = 0.0f;
  int64_t done = 0;
  if (VEC == 4) {
    const int64_t n4 = n >> 2;
    for (int64_t i = tid; i < n4; i += stride) {
      const float4 va = ld4(a + 4 * i), vb = ld4(b + 4 * i);
      const float4 vt = t_tensor != nullptr ? ld4(t_tensor + 4 * i) : make_float4(t_scalar, t_scalar, t_scalar, t_scalar);
      float4 r;
      r.x = blend<float>(mode, va.x, vb.x, vt.x);
      r.y = blend<float>(mode, va.y, vb.y, vt.y);
      r.z = blend<float>(mode, va.z, vb.z, vt.z);
      r.w = blend<float>(mode, va.w, vb.w, vt.w);
      st4(out + 4 * i, r);
      ms += (r.x + r.y) + (r.z + r.w);
      mss += (r.x * r.x + r.y * r.y) + (r.z * r.z + r.w * r.w);
    }
    done = n4 << 2;
  }
  for (int64_t i = done + tid; i < n; i += stride) {
    const float t = t_tensor != nullptr ? t_tensor[i] : t_scalar;
    const float r = blend<float>(mode, a[i], b[i], t);
    out[i] = r;
    ms += r;
    mss += r * r;
  }
  commit_moments(sums, sums_clear, ms, mss);
}

// out = a * alpha + b * beta (b may be NULL)
__device__ __forceinline__ float axpby_one(float a, float alpha, float b, float beta, bool has_b) {
  float v = alpha == 1.0f ? a : a * alpha;
  if (has_b) v = v + __fmul_rn(b, beta);
  return v;
}

template <int VEC>
__global__ void __launch_bounds__(kBlock)
axpby_kernel(const float* __restrict__ a, float alpha, const float* __restrict__ b, float beta, float* __restrict__ out,
             int64_t n, double* __restrict__ sums, double* __restrict__ sums_clear) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const bool has_b = b != nullptr;
  float ms = 0.0f, mss = 0.0f;
  int64_t done = 0;
  if (VEC == 4) {
    const int64_t n4 = n >> 2;
    for (int64_t i = tid; i < n4; i += stride) {
      const float4 va = ld4(a + 4 * i);
      const float4 vb = has_b ? ld4(b + 4 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
      float4 r;
      r.x = axpby_one(va.x, alpha, vb.x, beta, has_b);
      r.y = axpby_one(va.y, alpha, vb.y, beta, has_b);
      r.z = axpby_one(va.z, alpha, vb.z, beta, has_b);
      r.w = axpby_one(va.w, alpha, vb.w, beta, has_b);
      st4(out + 4 * i, r);
      ms += (r.x + r.y) + (r.z + r.w);
      mss += (r.x * r.x + r.y * r.y) + (r.z * r.z + r.w * r.w);
    }
    done = n4 << 2;
  }
  for (int64_t i = done + tid; i < n; i += stride) {
    const float r = axpby_one(a[i], alpha, has_b ? b[i] : 0.0f, beta, has_b);
    out[i] = r;
    ms += r;
    mss += r * r;
  }
  commit_moments(sums, sums_clear, ms, mss);
}

// out = x / d with a correctly rounded quotient (tensor.div_(python scalar), e.g. the octave-noise
// "result /= total_amplitude", py/noise_generation.py:2325)
__global__ void __launch_bounds__(kBlock)
div_scalar_kernel(const float* __restrict__ x, float* __restrict__ out, int64_t n, float d) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = x[i] / d;
}

// out = ((x + pre) * mul) + post with three separately rounded steps (sub_/mul_/add_ chains such as
// UniformNoiseGenerator.generate, py/noise_generation.py:508-514, stay bit-exact)
__global__ void __launch_bounds__(kBlock)
affine_kernel(const float* __restrict__ x, float* __restrict__ out, int64_t n, float pre, float mul, float post) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = __fadd_rn(__fmul_rn(__fadd_rn(x[i], pre), mul), post);
}

// out = dst * (1 - mask) + src * mask, mask is (B, 1, H, W) broadcast over channels.
// grid.y = batch item; a thread owns VEC consecutive positions of the (H, W) plane, loads their mask
// values once and walks the channels: no integer division, the mask is read once instead of C times.
template <int VEC>
__global__ void __launch_bounds__(kBlock)
composite_kernel(const float* __restrict__ dst, const float* __restrict__ src, const float* __restrict__ mask,
                 float* __restrict__ out, int64_t channels, int64_t hw) {
  const int64_t b = blockIdx.y;
  const float* mb = mask + b * hw;
  const int64_t item = b * channels * hw;
  for (int64_t pos = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * VEC; pos < hw;
       pos += (int64_t)gridDim.x * blockDim.x * VEC) {
    float m[VEC], im[VEC];
    if (VEC == 4) {
      const float4 m4 = ld4(mb + pos);
      m[0] = m4.x; m[1 % VEC] = m4.y; m[2 % VEC] = m4.z; m[3 % VEC] = m4.w;
    } else {
      m[0] = mb[pos];
    }
#pragma unroll
    for (int v = 0; v < VEC; ++v) im[v] = 1.0f - m[v];
    for (int64_t c = 0; c < channels; ++c) {
      const int64_t o = item + c * hw + pos;
      if (VEC == 4) {
        const float4 d = ld4(dst + o), sv = ld4(src + o);
        st4(out + o, make_float4(__fadd_rn(__fmul_rn(d.x, im[0]), __fmul_rn(sv.x, m[0])),
                                 __fadd_rn(__fmul_rn(d.y, im[1 % VEC]), __fmul_rn(sv.y, m[1 % VEC])),
                                 __fadd_rn(__fmul_rn(d.z, im[2 % VEC]), __fmul_rn(sv.z, m[2 % VEC])),
                                 __fadd_rn(__fmul_rn(d.w, im[3 % VEC]), __fmul_rn(sv.w, m[3 % VEC]))));
      } else {
        out[o] = __fadd_rn(__fmul_rn(dst[o], im[0]), __fmul_rn(src[o], m[0]));
      }
    }
  }
}

__global__ void __launch_bounds__(kBlock)
powerlaw_kernel(const float* __restrict__ x, float* __restrict__ out, int64_t n, float alpha, int use_sign) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = x[i];
    const float mod = powf(fabsf(v), alpha);
    const float lead = use_sign ? (float)((v > 0.0f) - (v < 0.0f)) : v;
    out[i] = lead * mod;
  }
}

// ---------------------------------------------------------------------------------------------
// per-item (leading dim) range reductions, two stage, no atomics
// ---------------------------------------------------------------------------------------------
constexpr int kRangeChunks = 64;

// partial[item][chunk] = {min, max} over that chunk (optionally of |x|)
__global__ void __launch_bounds__(kBlock)
item_range_kernel(const float* __restrict__ x, int64_t per_item, int use_abs, float2* __restrict__ partial) {
  __shared__ float smin[32], smax[32];
  const int item = blockIdx.y, chunk = blockIdx.x;
  const int64_t chunk_len = (per_item + kRangeChunks - 1) / kRangeChunks;
  const int64_t lo = (int64_t)chunk * chunk_len;
  const int64_t hi = lo + chunk_len < per_item ? lo + chunk_len : per_item;
  const float* base = x + (int64_t)item * per_item;
  float mn = INFINITY, mx = -INFINITY;
  for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    float v = base[i];
    if (use_abs) v = fabsf(v);
    mn = fminf(mn, v);
    mx = fmaxf(mx, v);
  }
  mn = warp_min(mn);
  mx = warp_max(mx);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) {
    smin[warp] = mn;
    smax[warp] = mx;
  }
  __syncthreads();
  if (warp == 0) {
    const int nw = blockDim.x >> 5;
    mn = lane < nw ? smin[lane] : INFINITY;
    mx = lane < nw ? smax[lane] : -INFINITY;
    mn = warp_min(mn);
    mx = warp_max(mx);
    if (lane == 0) partial[item * kRangeChunks + chunk] = make_float2(mn, mx);
  }
}

__device__ __forceinline__ float2 item_range(const float2* __restrict__ partial, int item) {
  float mn = INFINITY, mx = -INFINITY;
  for (int c = 0; c < kRangeChunks; ++c) {
    const float2 v = partial[item * kRangeChunks + c];
    mn = fminf(mn, v.x);
    mx = fmaxf(mx, v.y);
  }
  return make_float2(mn, mx);
}

// op 0: out = x / max ; op 1: normalize_to_scale
__global__ void __launch_bounds__(kBlock)
item_range_apply_kernel(const float* __restrict__ x, float* __restrict__ out, int64_t per_item,
                        const float2* __restrict__ partial, int op, float tmin, float tmax, float eps) {
  __shared__ float2 rng;
  const int item = blockIdx.y;
  if (threadIdx.x == 0) rng = item_range(partial, item);
  __syncthreads();
  const float mn = rng.x, mx = rng.y;
  const float* src = x + (int64_t)item * per_item;
  float* dst = out + (int64_t)item * per_item;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < per_item; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = src[i];
    if (op == 0) {
      dst[i] = v / mx;
    } else {
      float r = (v - mn) / ((mx - mn) + eps);
      r = __fadd_rn(__fmul_rn(r, tmax - tmin), tmin);
      dst[i] = fminf(fmaxf(r, tmin), tmax);
    }
  }
}

}  // namespace sonar

extern "C" {

int sonar_pyramid_accum_f32(const SonarPyramidParams* params, void* stream) {
  using namespace sonar;
  if (params == nullptr) return (int)cudaErrorInvalidValue;
  SonarPyramidParams p = *params;
  if (p.planes <= 0 || p.H <= 0 || p.W <= 0) return 0;
  if (p.n_levels < 0 || p.n_levels > SONAR_PYRAMID_MAX_LEVELS || p.out == nullptr) return (int)cudaErrorInvalidValue;
  for (int l = 0; l < p.n_levels; ++l)
    if (p.levels[l] == nullptr || p.level_h[l] <= 0 || p.level_w[l] <= 0) return (int)cudaErrorInvalidValue;
  bool vec = (p.W % 4 == 0) && aligned16(p.out) && (p.base == nullptr || aligned16(p.base));
  for (int l = 0; l < p.n_levels; ++l)
    if (p.level_h[l] == p.H && p.level_w[l] == p.W && !aligned16(p.levels[l])) vec = false;
  // table-driven row kernel: tap tables + one blended source row per level and warp in shared memory
  int max_lw = 1;
  bool taps_fit_u16 = true;  // tap indices are stored as unsigned short
  for (int l = 0; l < p.n_levels; ++l) {
    if (!(p.level_h[l] == p.H && p.level_w[l] == p.W) && p.level_w[l] > max_lw) max_lw = p.level_w[l];
    if (p.level_h[l] > 65535 || p.level_w[l] > 65535) taps_fit_u16 = false;
  }
  const int row_pitch = max_lw;
  const size_t tab_bytes = (size_t)p.n_levels * (p.W + p.H) * 8 + (size_t)(kBlock / 32) * p.n_levels * row_pitch * sizeof(float);
  if (p.mode != SONAR_RESAMPLE_AREA && p.n_levels > 0 && taps_fit_u16 && tab_bytes <= 96 * 1024 && p.planes * (int64_t)p.H < (1ll << 31) &&
      (int64_t)p.H * p.W < (1ll << 30) && p.H < 65536 && p.W < 65536) {
    // one warp per row; PERSISTENT CTAs (4 per SM): every CTA builds the tap tables once and then walks
    // many rows (with one CTA per 8 rows the table build was ~40 % of the issued instructions)
    const int64_t n_rows = p.planes * (int64_t)p.H;
    int64_t grid = (n_rows + kBlock / 32 - 1) / (kBlock / 32);
    const int64_t cap = (int64_t)device_info().sm_count * 4;
    if (grid > cap) grid = cap;
    if (vec) {
      SONAR_CUDA_TRY(cudaFuncSetAttribute(pyramid_rows_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tab_bytes));
      pyramid_rows_kernel<4><<<(unsigned)grid, kBlock, tab_bytes, (cudaStream_t)stream>>>(p, row_pitch);
    } else {
      SONAR_CUDA_TRY(cudaFuncSetAttribute(pyramid_rows_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tab_bytes));
      pyramid_rows_kernel<1><<<(unsigned)grid, kBlock, tab_bytes, (cudaStream_t)stream>>>(p, row_pitch);
    }
    SONAR_LAUNCH_CHECK();
    return 0;
  }
  const int64_t total = p.planes * (int64_t)p.H * (vec ? p.W / 4 : p.W);
  const int grid = streaming_grid(total, kBlock, 4);
  if (vec)
    pyramid_accum_kernel<4><<<grid, kBlock, 0, (cudaStream_t)stream>>>(p);
  else
    pyramid_accum_kernel<1><<<grid, kBlock, 0, (cudaStream_t)stream>>>(p);
  SONAR_LAUNCH_CHECK();
  return 0;
}

int sonar_perlin_accum_f32(const SonarPerlinParams* params, void* stream) {
  using namespace sonar;
  if (params == nullptr) return (int)cudaErrorInvalidValue;
  SonarPerlinParams p = *params;
  if (p.B <= 0 || p.C <= 0 || p.H <= 0 || p.W <= 0) return 0;
  if (p.iterations < 0 || p.iterations > SONAR_PERLIN_MAX_ITERS || p.out == nullptr) return (int)cudaErrorInvalidValue;
  for (int i = 0; i < p.iterations; ++i)
    if (p.angles[i] == nullptr) return (int)cudaErrorInvalidValue;
  const bool vec = (p.W % 4 == 0) && aligned16(p.out) && (p.base == nullptr || aligned16(p.base));
  const int64_t threads = (int64_t)p.C * p.H * (vec ? p.W / 4 : p.W);
  const int grid = streaming_grid(threads, kBlock, 4);
  cudaStream_t st = (cudaStream_t)stream;
  // the stencil (20 sincosf per thread at VEC = 4) is shared by the batch and costs about as much as 16
  // batch items of traffic, so the batch is split over grid.y only when (C, H, W) alone cannot give every
  // SM a CTA (measured at 16x16x128x128: 15 us unsplit, 21 us split 3-way)
  const int want = device_info().sm_count;
  int gy = grid >= want ? 1 : (want + grid - 1) / grid;
  if (gy > p.B) gy = p.B;
  const dim3 grid2((unsigned)grid, (unsigned)gy);
#define PERLIN_LAUNCH(ITERS)                                        \
  do {                                                              \
    if (vec)                                                        \
      perlin_accum_kernel<4, ITERS><<<grid2, kBlock, 0, st>>>(p);   \
    else                                                            \
      perlin_accum_kernel<1, ITERS><<<grid2, kBlock, 0, st>>>(p);   \
  } while (0)
  switch ((int64_t)p.C * p.H * p.W < (1ll << 31) ? p.iterations : 0) {
    case 1: PERLIN_LAUNCH(1); break;
    case 2: PERLIN_LAUNCH(2); break;
    case 3: PERLIN_LAUNCH(3); break;
    case 4: PERLIN_LAUNCH(4); break;
    default:
      perlin_accum_generic_kernel<<<streaming_grid((int64_t)p.C * p.H * p.W, kBlock, 4), kBlock, 0, st>>>(p);
      break;
  }
#undef PERLIN_LAUNCH
  SONAR_LAUNCH_CHECK();
  return 0;
}

int sonar_blend_f32(const float* a, const float* b, const float* t_tensor, float t_scalar, float* out, int64_t n,
                    int mode, double* sums, double* sums_clear, void* stream) {
  using namespace sonar;
  if (n <= 0) return 0;
  const bool vec = aligned16(a) && aligned16(b) && aligned16(out) && (t_tensor == nullptr || aligned16(t_tensor));
  const int grid = streaming_grid(vec ? (n + 3) / 4 : n, kBlock, 2);
  if (vec)
    blend_kernel<4><<<grid, kBlock, 0, (cudaStream_t)stream>>>(a, b, t_tensor, t_scalar, out, n, mode, sums, sums_clear);
  else
    blend_kernel<1><<<grid, kBlock, 0, (cudaStream_t)stream>>>(a, b, t_tensor, t_scalar, out, n, mode, sums, sums_clear);
  SONAR_LAUNCH_CHECK();
  return 0;
}

int sonar_axpby_f32(const float* a, float alpha, const float* b, float beta, float* out, int64_t n, double* sums,
                    double* sums_clear, void* stream) {
  using namespace sonar;
  if (n <= 0) return 0;
  const bool vec = aligned16(a) && aligned16(out) && (b == nullptr || aligned16(b));
  const int grid = streaming_grid(vec ? (n + 3) / 4 : n, kBlock, 2);
  if (vec)
    axpby_kernel<4><<<grid, kBlock, 0, (cudaStream_t)stream>>>(a, alpha, b, beta, out, n, sums, sums_clear);
  else
    axpby_kernel<1><<<grid, kBlock, 0, (cudaStream_t)stream>>>(a, alpha, b, beta, out, n, sums, sums_clear);
  SONAR_LAUNCH_CHECK();
  return 0;
}

int sonar_div_scalar_f32(const float* x, float* out, int64_t n, float divisor, void* stream) {
  if (n <= 0) return 0;
  const int grid = sonar::streaming_grid(n, sonar::kBlock, 4);
  sonar::div_scalar_kernel<<<grid, sonar::kBlock, 0, (cudaStream_t)stream>>>(x, out, n, divisor);
  SONAR_LAUNCH_CHECK();
  return 0;
}

int sonar_affine_f32(const float* x, float* out, int64_t n, float pre_add, float mul, float post_add, void* stream) {
  if (n <= 0) return 0;
  const int grid = sonar::streaming_grid(n, sonar::kBlock, 4);
  sonar::affine_kernel<<<grid, sonar::kBlock, 0, (cudaStream_t)stream>>>(x, out, n, pre_add, mul, post_add);
  SONAR_LAUNCH_CHECK();
  return 0;
}

int sonar_composite_f32(const float* dst, const float* src, const float* mask, float* out, int64_t batch,
                        int64_t channels, int64_t hw, void* stream) {
  using namespace sonar;
  if (batch <= 0 || channels <= 0 || hw <= 0) return 0;
  if (batch > 65535) return (int)cudaErrorInvalidValue;
  const bool vec = hw % 4 == 0 && aligned16(dst) && aligned16(src) && aligned16(mask) && aligned16(out);
  const int64_t threads = vec ? hw / 4 : hw;
  int64_t gx = (threads + kBlock - 1) / kBlock;
  const int64_t cap = (int64_t)device_info().sm_count * 16 / batch + 1;
  if (gx > cap) gx = cap;
  const dim3 grid((unsigned)gx, (unsigned)batch);
  if (vec)
    composite_kernel<4><<<grid, kBlock, 0, (cudaStream_t)stream>>>(dst, src, mask, out, channels, hw);
  else
    composite_kernel<1><<<grid, kBlock, 0, (cudaStream_t)stream>>>(dst, src, mask, out, channels, hw);
  SONAR_LAUNCH_CHECK();
  return 0;
}

int sonar_powerlaw_f32(const float* x, float* out, int64_t n, float alpha, int use_sign, void* stream) {
  if (n <= 0) return 0;
  const int grid = sonar::streaming_grid(n, sonar::kBlock, 4);
  sonar::powerlaw_kernel<<<grid, sonar::kBlock, 0, (cudaStream_t)stream>>>(x, out, n, alpha, use_sign);
  SONAR_LAUNCH_CHECK();
  return 0;
}

int sonar_item_range_scratch_bytes(int64_t items) { return (int)(items * sonar::kRangeChunks * sizeof(float2)); }

int sonar_item_div_max_f32(const float* x, float* out, int64_t items, int64_t per_item, int use_abs, void* scratch,
                           void* stream) {
  using namespace sonar;
  if (items <= 0 || per_item <= 0) return 0;
  if (items > 65535) return (int)cudaErrorInvalidValue;
  float2* partial = (float2*)scratch;
  item_range_kernel<<<dim3(kRangeChunks, (unsigned)items), kBlock, 0, (cudaStream_t)stream>>>(x, per_item, use_abs, partial);
  SONAR_LAUNCH_CHECK();
  int gx = (int)((per_item + kBlock * 4 - 1) / (kBlock * 4));
  if (gx < 1) gx = 1;
  if (gx > 1024) gx = 1024;
  item_range_apply_kernel<<<dim3(gx, (unsigned)items), kBlock, 0, (cudaStream_t)stream>>>(x, out, per_item, partial, 0,
                                                                                          0.f, 0.f, 0.f);
  SONAR_LAUNCH_CHECK();
  return 0;
}

int sonar_item_minmax_rescale_f32(const float* x, float* out, int64_t items, int64_t per_item, float target_min,
                                  float target_max, float eps, void* scratch, void* stream) {
  using namespace sonar;
  if (items <= 0 || per_item <= 0) return 0;
  if (items > 65535) return (int)cudaErrorInvalidValue;
  float2* partial = (float2*)scratch;
  item_range_kernel<<<dim3(kRangeChunks, (unsigned)items), kBlock, 0, (cudaStream_t)stream>>>(x, per_item, 0, partial);
  SONAR_LAUNCH_CHECK();
  int gx = (int)((per_item + kBlock * 4 - 1) / (kBlock * 4));
  if (gx < 1) gx = 1;
  if (gx > 1024) gx = 1024;
  item_range_apply_kernel<<<dim3(gx, (unsigned)items), kBlock, 0, (cudaStream_t)stream>>>(
      x, out, per_item, partial, 1, target_min, target_max, eps);
  SONAR_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
