// 2-D discrete wavelet transform levels (analysis / synthesis) with the band-split CFG combine
// folded into the synthesis reads and the final crop + "x - result" epilogue.
//
// Reference: Wavelet.forward/inverse py/wavelet_functions.py:81-105 (pytorch_wavelets DWTForward /
// DWTInverse, an [upstream] dependency absent from the reference tree: no pinned version, restated
// from its published algorithm -- lowlevel.afb1d / sfb1d: pad-correlate-decimate with reversed
// dec_lo/dec_hi, transposed conv with rec_lo/rec_hi trimmed by L-2), wavelet_scaling :193-216,
// wavelet_blend :219-238, WaveletCFG.wavelet_cfg py/wavelet_cfg.py:749-791, process_output :729-747.
//
// NOT a lifting scheme: the reference default mode is "symmetric", an expansive transform with
// floor((N+L-1)/2) coefficients per axis, so this is the direct 2-D separable form of
// pad -> correlate -> decimate; row and column passes are fused (no row-filtered intermediate).
// Band order of the detail tensor (planes, 3, h, w): [0] high along H / low along W,
// [1] low along H / high along W, [2] high / high  (pytorch_wavelets AFB2D channel order).
#include <cooperative_groups.h>
#include <cstdlib>

#include "common.cuh"
#include "../../include/sonar_b200.h"

namespace sonar {

namespace cg = cooperative_groups;

__device__ __forceinline__ int extend_index_wrapped(int i, int n, int mode) {
  // general case: the index lies more than one signal length outside [0, n)
  switch (mode) {
    case SONAR_DWT_MODE_PERIODIC: {
      int m = i % n;
      return m < 0 ? m + n : m;
    }
    case SONAR_DWT_MODE_PERIODIZATION: {  // period n rounded up to even, the extra sample repeats the last one
      const int period = n + (n & 1);
      int m = i % period;
      if (m < 0) m += period;
      return m < n ? m : n - 1;
    }
    case SONAR_DWT_MODE_REFLECT: {  // whole-sample symmetry about 0 and n-1
      if (n == 1) return 0;
      const int period = 2 * (n - 1);
      int m = i % period;
      if (m < 0) m += period;
      return m < n ? m : period - m;
    }
    default: {  // symmetric: half-sample symmetry (edge value repeated)
      const int period = 2 * n;
      int m = i % period;
      if (m < 0) m += period;
      return m < n ? m : period - 1 - m;
    }
  }
}

__device__ __forceinline__ int extend_index(int i, int n, int mode) {
  // maps an index of the padded signal (already shifted by the left pad) into [0, n), or -1 (zero).
  // One reflection / wrap covers every tap unless the filter is longer than the signal.
  if ((unsigned)i < (unsigned)n) return i;
  if (mode == SONAR_DWT_MODE_ZERO) return -1;
  if (mode == SONAR_DWT_MODE_PERIODIZATION) return extend_index_wrapped(i, n, mode);
  int m;
  if (mode == SONAR_DWT_MODE_PERIODIC)
    m = i < 0 ? i + n : i - n;
  else if (mode == SONAR_DWT_MODE_REFLECT)
    m = i < 0 ? -i : 2 * (n - 1) - i;
  else
    m = i < 0 ? -1 - i : 2 * n - 1 - i;
  return (unsigned)m < (unsigned)n ? m : extend_index_wrapped(i, n, mode);
}

template <typename T>
struct Filters {
  T a_lo[SONAR_DWT_MAX_TAPS];  // analysis, already reversed (correlation form)
  T a_hi[SONAR_DWT_MAX_TAPS];
  T s_lo[SONAR_DWT_MAX_TAPS];  // synthesis rec_lo / rec_hi
  T s_hi[SONAR_DWT_MAX_TAPS];
  int L;
};

// ---------------------------------------------------------------------------------------------
// analysis level: in (planes, H, W) -> ll (planes, h, w), hi (planes, 3, h, w)
// value(y, x) = in_a[y, x] - in_b[y, x] (in_b optional) so that level 1 can transform cond - uncond
// straight from the fp32 inputs.
// ---------------------------------------------------------------------------------------------
// One output position (ky, kx) of an analysis level: the four sub-band coefficients of the
// L x L window at (2ky - pad_t, 2kx - pad_l) of value = pa - pb (pb optional). Rows of the input are
// `row_stride` elements apart (global planes: W; shared-memory buffers: their own pitch).
// LT > 0: filter length known at compile time (loops fully unrolled, taps addressed with immediates);
// LT == 0: any even length up to SONAR_DWT_MAX_TAPS at run time.
template <typename T, typename Tin, int LT>
__device__ __forceinline__ void analysis_point(const Tin* __restrict__ pa, const Tin* __restrict__ pb, int row_stride,
                                               int H, int W, int ky, int kx, int pad_t, int pad_l, int mode,
                                               const Filters<T>& f, T& acc_ll, T& acc_lh, T& acc_hl, T& acc_hh) {
  const int L = LT > 0 ? LT : f.L;
  acc_ll = 0, acc_lh = 0, acc_hl = 0, acc_hh = 0;
  const int x_first = 2 * kx - pad_l, y_first = 2 * ky - pad_t;
  if (LT > 0) {
    // compile-time length: ONE code path for interior and border outputs (a warp almost always holds
    // both, so a separate slow border path would serialise in front of every interior output).
    // Tap offsets are resolved once per output: 2L index computations instead of L*L.
    int ox[LT > 0 ? LT : 1], oy[LT > 0 ? LT : 1];
#pragma unroll
    for (int j = 0; j < LT; ++j) {
      ox[j] = extend_index(x_first + j, W, mode);
      const int sy = extend_index(y_first + j, H, mode);
      oy[j] = sy < 0 ? -1 : sy * row_stride;
    }
#pragma unroll
    for (int jy = 0; jy < LT; ++jy) {
      T row_lo = 0, row_hi = 0;
#pragma unroll
      for (int jx = 0; jx < LT; ++jx) {
        const bool in = (oy[jy] | ox[jx]) >= 0;  // zero padding: either index negative
        const int o = in ? oy[jy] + ox[jx] : 0;
        T v = (T)pa[o];
        if (pb != nullptr) v -= (T)pb[o];
        if (!in) v = 0;
        row_lo += f.a_lo[jx] * v;
        row_hi += f.a_hi[jx] * v;
      }
      acc_ll += f.a_lo[jy] * row_lo;
      acc_lh += f.a_hi[jy] * row_lo;  // high along H, low along W
      acc_hl += f.a_lo[jy] * row_hi;  // low along H, high along W
      acc_hh += f.a_hi[jy] * row_hi;
    }
    return;
  }
  const bool interior = x_first >= 0 && x_first + L <= W && y_first >= 0 && y_first + L <= H;
  if (interior) {  // the whole L x L window lies inside the plane: no boundary extension
    const Tin* ra = pa + (y_first * row_stride + x_first);
    const Tin* rb = pb != nullptr ? pb + (y_first * row_stride + x_first) : nullptr;
#pragma unroll
    for (int jy = 0; jy < L; ++jy, ra += row_stride, rb += (pb != nullptr ? row_stride : 0)) {
      T row_lo = 0, row_hi = 0;
#pragma unroll
      for (int jx = 0; jx < L; ++jx) {
        T v = (T)ra[jx];
        if (rb != nullptr) v -= (T)rb[jx];
        row_lo += f.a_lo[jx] * v;
        row_hi += f.a_hi[jx] * v;
      }
      acc_ll += f.a_lo[jy] * row_lo;
      acc_lh += f.a_hi[jy] * row_lo;
      acc_hl += f.a_lo[jy] * row_hi;
      acc_hh += f.a_hi[jy] * row_hi;
    }
  } else {
    for (int jy = 0; jy < L; ++jy) {
      const int sy = extend_index(y_first + jy, H, mode);
      if (sy < 0) continue;
      T row_lo = 0, row_hi = 0;
      for (int jx = 0; jx < L; ++jx) {
        const int sx = extend_index(x_first + jx, W, mode);
        if (sx < 0) continue;
        const int o = sy * row_stride + sx;
        T v = (T)pa[o];
        if (pb != nullptr) v -= (T)pb[o];
        row_lo += f.a_lo[jx] * v;
        row_hi += f.a_hi[jx] * v;
      }
      acc_ll += f.a_lo[jy] * row_lo;
      acc_lh += f.a_hi[jy] * row_lo;  // high along H, low along W
      acc_hl += f.a_lo[jy] * row_hi;  // low along H, high along W
      acc_hh += f.a_hi[jy] * row_hi;
    }
  }
}

// Interior outputs only (the whole LT x LT window lies inside the plane): no index resolution, no
// padding selects, row pointers advanced by a constant. VEC2: fp32 input rows read as float2 pairs
// (x_first even, even row stride, 8-byte aligned planes).
template <typename T, typename Tin, int LT, bool VEC2>
__device__ __forceinline__ void analysis_interior(const Tin* __restrict__ pa, const Tin* __restrict__ pb, int row_stride,
                                                  int y_first, int x_first, const Filters<T>& f, T& acc_ll, T& acc_lh,
                                                  T& acc_hl, T& acc_hh) {
  acc_ll = 0, acc_lh = 0, acc_hl = 0, acc_hh = 0;
  const Tin* ra = pa + (y_first * row_stride + x_first);
  const Tin* rb = pb != nullptr ? pb + (y_first * row_stride + x_first) : nullptr;
#pragma unroll
  for (int jy = 0; jy < LT; ++jy) {
    T v[LT > 0 ? LT : 1];
    if (VEC2) {
#pragma unroll
      for (int jx = 0; jx < LT; jx += 2) {
        const float2 a2 = *reinterpret_cast<const float2*>(ra + jx);
        v[jx] = (T)a2.x;
        v[jx + 1] = (T)a2.y;
        if (rb != nullptr) {
          const float2 b2 = *reinterpret_cast<const float2*>(rb + jx);
          v[jx] -= (T)b2.x;
          v[jx + 1] -= (T)b2.y;
        }
      }
    } else {
#pragma unroll
      for (int jx = 0; jx < LT; ++jx) {
        v[jx] = (T)ra[jx];
        if (rb != nullptr) v[jx] -= (T)rb[jx];
      }
    }
    T row_lo = 0, row_hi = 0;
#pragma unroll
    for (int jx = 0; jx < LT; ++jx) {
      row_lo += f.a_lo[jx] * v[jx];
      row_hi += f.a_hi[jx] * v[jx];
    }
    acc_ll += f.a_lo[jy] * row_lo;
    acc_lh += f.a_hi[jy] * row_lo;
    acc_hl += f.a_lo[jy] * row_hi;
    acc_hh += f.a_hi[jy] * row_hi;
    ra += row_stride;
    if (rb != nullptr) rb += row_stride;
  }
}

template <typename T, typename Tin, int LT>
__global__ void __launch_bounds__(kBlock)
dwt2_analysis_kernel(const Tin* __restrict__ in_a, const Tin* __restrict__ in_b, T* __restrict__ ll,
                     T* __restrict__ hi, int64_t planes, int H, int W, int in_stride_h, int h, int w, int mode,
                     Filters<T> f) {
  const int L = LT > 0 ? LT : f.L;
  const int pad_t = (2 * (h - 1) - H + L) / 2;
  const int pad_l = (2 * (w - 1) - W + L) / 2;
  const int64_t total = planes * (int64_t)h * w;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    int kx, ky;
    int64_t plane;
    if (total < (1ll << 31)) {  // 32-bit division (the 64-bit one is emulated: ~100 instructions each)
      const unsigned i32 = (unsigned)idx, row = i32 / (unsigned)w;
      kx = (int)(i32 - row * (unsigned)w);
      plane = row / (unsigned)h;
      ky = (int)(row - (unsigned)plane * (unsigned)h);
    } else {
      kx = (int)(idx % w);
      ky = (int)((idx / w) % h);
      plane = idx / ((int64_t)w * h);
    }
    const Tin* pa = in_a + plane * (int64_t)in_stride_h * W;
    const Tin* pb = in_b != nullptr ? in_b + plane * (int64_t)in_stride_h * W : nullptr;
    T acc_ll, acc_lh, acc_hl, acc_hh;
    analysis_point<T, Tin, LT>(pa, pb, W, H, W, ky, kx, pad_t, pad_l, mode, f, acc_ll, acc_lh, acc_hl, acc_hh);
    const int64_t hw = (int64_t)h * w;
    const int64_t o = (int64_t)ky * w + kx;
    ll[plane * hw + o] = acc_ll;
    T* ph = hi + plane * 3 * hw;
    ph[o] = acc_lh;
    ph[hw + o] = acc_hl;
    ph[2 * hw + o] = acc_hh;
  }
}

// ---------------------------------------------------------------------------------------------
// synthesis level: up to two coefficient sets with per-band scales (the CFG combine), output
// (planes, 2h-L+2, 2w-L+2) in T, or -- final level -- cropped fp32 with the epilogue
//   out = sign * (recon + addend_scale * addend) + x_scale * x
// ---------------------------------------------------------------------------------------------
template <typename T>
struct SynthSet {
  const T* ll;
  const T* hi;
  int ll_stride_h;  // rows/cols of the ll buffer (may be one larger than h/w: "unpad")
  int ll_stride_w;
  T s_ll, s_lh, s_hl, s_hh;
};

// Accumulates the contribution of one coefficient set to the 2x2 output quad (2qy+py, 2qx+px).
// pll: approximation band with row pitch ll_pitch; phi: three detail bands, `band_stride` apart, row
// pitch hi_pitch; (h, w) = valid coefficient extent. The per-band scales are folded into the x taps
// (one product per tap and band instead of one per coefficient) and every accumulation is a single FMA.
template <typename T, int LT>
__device__ __forceinline__ void synthesis_quad(const T* __restrict__ pll, int ll_pitch, const T* __restrict__ phi,
                                               int64_t band_stride, int hi_pitch, int h, int w, int qy, int qx, T s_ll,
                                               T s_lh, T s_hl, T s_hh, const Filters<T>& f, T& o00, T& o01, T& o10,
                                               T& o11) {
  const int L = LT > 0 ? LT : f.L, half = L >> 1;
#pragma unroll
  for (int ia = 0; ia < half; ++ia) {
    const int ky = qy + ia;
    if (ky >= h) break;  // only reachable for the cropped-away overhang
    T rl0 = 0, rl1 = 0, rh0 = 0, rh1 = 0;
    const T* rll = pll + ky * ll_pitch + qx;
    const T* rhi = phi + ky * hi_pitch + qx;
#pragma unroll
    for (int ib = 0; ib < half; ++ib) {
      if (qx + ib >= w) break;
      const int tx = L - 2 - 2 * ib;
      const T lo0 = f.s_lo[tx], lo1 = f.s_lo[tx + 1], hi0 = f.s_hi[tx], hi1 = f.s_hi[tx + 1];
      const T v_ll = rll[ib];
      const T v_lh = rhi[ib];                    // high along H, low along W
      const T v_hl = rhi[band_stride + ib];      // low along H, high along W
      const T v_hh = rhi[2 * band_stride + ib];
      rl0 = fma(lo0 * s_ll, v_ll, rl0);
      rl0 = fma(hi0 * s_hl, v_hl, rl0);
      rl1 = fma(lo1 * s_ll, v_ll, rl1);
      rl1 = fma(hi1 * s_hl, v_hl, rl1);
      rh0 = fma(lo0 * s_lh, v_lh, rh0);
      rh0 = fma(hi0 * s_hh, v_hh, rh0);
      rh1 = fma(lo1 * s_lh, v_lh, rh1);
      rh1 = fma(hi1 * s_hh, v_hh, rh1);
    }
    const int ty = L - 2 - 2 * ia;
    const T gly0 = f.s_lo[ty], gly1 = f.s_lo[ty + 1], ghy0 = f.s_hi[ty], ghy1 = f.s_hi[ty + 1];
    o00 = fma(gly0, rl0, o00);
    o00 = fma(ghy0, rh0, o00);
    o01 = fma(gly0, rl1, o01);
    o01 = fma(ghy0, rh1, o01);
    o10 = fma(gly1, rl0, o10);
    o10 = fma(ghy1, rh0, o10);
    o11 = fma(gly1, rl1, o11);
    o11 = fma(ghy1, rh1, o11);
  }
}

// Polyphase form: the 2x2 output quad (2qy+py, 2qx+px) reads the SAME (L/2) x (L/2) window of
// coefficients k = (qy + a, qx + b); tap index t = p + L - 2 - 2a. Every k of a quad is in range
// (q <= n - L/2), so the loops carry no boundary tests; rows are combined along W first.
template <typename T, int LT>
__global__ void __launch_bounds__(kBlock)
dwt2_synthesis_kernel(SynthSet<T> a, SynthSet<T> b, int n_sets, int64_t planes, int h, int w, int out_h, int out_w,
                      T* __restrict__ out_t, float* __restrict__ out_f32, int crop_h, int crop_w,
                      const float* __restrict__ addend, float addend_scale, const float* __restrict__ x,
                      float x_scale, float recon_sign, Filters<T> f) {
  const int oh = out_f32 != nullptr ? crop_h : out_h;
  const int ow = out_f32 != nullptr ? crop_w : out_w;
  const int qh = (oh + 1) >> 1, qw = (ow + 1) >> 1;
  const int64_t total = planes * (int64_t)qh * qw;
  const int64_t hw = (int64_t)h * w;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    int qx, qy;
    int64_t plane;
    if (total < (1ll << 31)) {
      const unsigned i32 = (unsigned)idx, row = i32 / (unsigned)qw;
      qx = (int)(i32 - row * (unsigned)qw);
      plane = row / (unsigned)qh;
      qy = (int)(row - (unsigned)plane * (unsigned)qh);
    } else {
      qx = (int)(idx % qw);
      qy = (int)((idx / qw) % qh);
      plane = idx / ((int64_t)qw * qh);
    }
    T o00 = 0, o01 = 0, o10 = 0, o11 = 0;
    for (int s = 0; s < n_sets; ++s) {
      const SynthSet<T>& c = s == 0 ? a : b;
      synthesis_quad<T, LT>(c.ll + plane * (int64_t)c.ll_stride_h * c.ll_stride_w, c.ll_stride_w, c.hi + plane * 3 * hw, hw, w,
                        h, w, qy, qx, c.s_ll, c.s_lh, c.s_hl, c.s_hh, f, o00, o01, o10, o11);
    }
    const T vals[4] = {o00, o01, o10, o11};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int iy = 2 * qy + (q >> 1), ix = 2 * qx + (q & 1);
      if (iy >= oh || ix >= ow) continue;
      if (out_f32 != nullptr) {
        const int64_t o = (plane * crop_h + iy) * (int64_t)crop_w + ix;
        T r = vals[q];
        if (addend != nullptr) r += (T)addend_scale * (T)addend[o];
        r = (T)recon_sign * r;
        // reference casts the reconstruction to x.dtype first, then forms x - result in that dtype
        float rf = (float)r;
        if (x != nullptr) rf = x_scale * x[o] + rf;
        out_f32[o] = rf;
      } else {
        out_t[(plane * out_h + iy) * (int64_t)out_w + ix] = vals[q];
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Periodization (non-expansive) synthesis level: N = 2h outputs per axis,
//   out[i] = sum over taps t == u (mod 2) of c[((u - t) / 2) mod h] * g[t],  u = (i + L/2 - 1) mod N
// (pytorch_wavelets sfb1d, mode "periodization": transposed convolution, the L-2 tail wrapped onto the
// head, rolled by 1 - L/2). One thread per output pixel; used by the wavelet-filtered noise type
// (py/noise_generation.py:1908-2032), not by the wavelet-CFG hot path.
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kBlock)
dwt2_per_synthesis_kernel(SynthSet<T> c, int64_t planes, int h, int w, T* __restrict__ out, Filters<T> f) {
  const int L = f.L, oh = 2 * h, ow = 2 * w;
  const int64_t total = planes * (int64_t)oh * ow, hw = (int64_t)h * w;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int ix = (int)(idx % ow), iy = (int)((idx / ow) % oh);
    const int64_t plane = idx / ((int64_t)ow * oh);
    const int uy = (iy + L / 2 - 1) % oh, ux = (ix + L / 2 - 1) % ow;
    const T* pll = c.ll + plane * (int64_t)c.ll_stride_h * c.ll_stride_w;
    const T* phi = c.hi + plane * 3 * hw;
    T acc = 0;
    for (int ty = uy & 1; ty < L; ty += 2) {
      int ky = ((uy - ty) / 2) % h;
      if (ky < 0) ky += h;
      T row_lo = 0, row_hi = 0;  // combined along W for the low-along-H and the high-along-H bands
      for (int tx = ux & 1; tx < L; tx += 2) {
        int kx = ((ux - tx) / 2) % w;
        if (kx < 0) kx += w;
        const int64_t o = (int64_t)ky * w + kx;
        const T v_ll = pll[(int64_t)ky * c.ll_stride_w + kx] * c.s_ll;
        const T v_lh = phi[o] * c.s_lh;           // high along H, low along W
        const T v_hl = phi[hw + o] * c.s_hl;      // low along H, high along W
        const T v_hh = phi[2 * hw + o] * c.s_hh;
        row_lo += f.s_lo[tx] * v_ll + f.s_hi[tx] * v_hl;
        row_hi += f.s_lo[tx] * v_lh + f.s_hi[tx] * v_hh;
      }
      acc += f.s_lo[ty] * row_lo + f.s_hi[ty] * row_hi;
    }
    out[idx] = acc;
  }
}

// ---------------------------------------------------------------------------------------------
// Whole wavelet-CFG call in ONE launch: one CTA per plane, all J analysis levels, the per-band
// scaling and all J synthesis levels with every coefficient resident in shared memory (db2 / 3 levels
// on a 128x128 plane: 184 KB of fp64 coefficients). HBM traffic is the algorithmic minimum: the fp32
// inputs are read once (level-1 windows overlap in L1), the fp32 result is written once.
//   value = in_a - in_b ;  out = x_scale*x + (float)(recon_sign*(IDWT(S (.) DWT(value)) + addend_scale*addend))
// Buffers: LL[j] holds ll_j during analysis and is reused for the reconstruction rec_{j+1} on the way
// back (which may be one row/column larger); HI[j] holds the three detail bands of level j.
// ---------------------------------------------------------------------------------------------
struct WcfgGeom {
  int levels;
  int pair;        // 1: a plane is shared by a cluster of two CTAs (see wcfg_fused_kernel)
  int hi0_rows;    // rows of the level-1 detail bands a CTA holds (pair: its half plus the synthesis halo)
  int h[SONAR_WCFG_MAX_LEVELS], w[SONAR_WCFG_MAX_LEVELS];            // coefficient extents, fine -> coarse
  int ll_off[SONAR_WCFG_MAX_LEVELS], hi_off[SONAR_WCFG_MAX_LEVELS];  // element offsets into shared memory
  int total;                                                         // elements
};

// Level-1 coefficient rows [lo, hi) that rank `rank` of a CTA pair computes and keeps: the rows its half
// of the final synthesis quads reads (quad qy reads rows qy .. qy + L/2 - 1).
__host__ __device__ inline void wcfg_pair_rows(int H, int h0, int L, int rank, int* q_lo, int* q_hi, int* r_lo, int* r_hi) {
  const int qh = (H + 1) >> 1, q_mid = (qh + 1) >> 1;
  *q_lo = rank ? q_mid : 0;
  *q_hi = rank ? qh : q_mid;
  *r_lo = *q_lo;
  const int top = *q_hi + L / 2 - 1;
  *r_hi = top < h0 ? top : h0;
}

static bool wcfg_geometry(int H, int W, int L, int levels, WcfgGeom* g, bool pair = false) {
  if (levels < 1 || levels > SONAR_WCFG_MAX_LEVELS || H <= 0 || W <= 0) return false;
  g->levels = levels;
  g->pair = pair ? 1 : 0;
  int hh = H, ww = W;
  for (int j = 0; j < levels; ++j) {
    hh = (hh + L - 1) / 2;
    ww = (ww + L - 1) / 2;
    g->h[j] = hh;
    g->w[j] = ww;
  }
  int64_t off = 0;
  for (int j = 0; j < levels; ++j) {
    int64_t ll = (int64_t)g->h[j] * g->w[j];
    if (j + 1 < levels) {  // rec_{j+1} lands here on the way back
      const int64_t rec = (int64_t)(2 * g->h[j + 1] - L + 2) * (2 * g->w[j + 1] - L + 2);
      if (rec > ll) ll = rec;
    }
    g->ll_off[j] = (int)off;
    off += ll + (ll & 1);
    g->hi_off[j] = (int)off;
    int hi_rows = g->h[j];
    if (j == 0) {
      if (pair) {
        int q_lo, q_hi, r_lo, r_hi;
        wcfg_pair_rows(H, g->h[0], L, 0, &q_lo, &q_hi, &r_lo, &r_hi);
        hi_rows = r_hi - r_lo;
        wcfg_pair_rows(H, g->h[0], L, 1, &q_lo, &q_hi, &r_lo, &r_hi);
        if (r_hi - r_lo > hi_rows) hi_rows = r_hi - r_lo;
      }
      g->hi0_rows = hi_rows;
    }
    off += 3 * (int64_t)hi_rows * g->w[j];
    off += off & 1;
    if (off > (1 << 28)) return false;
  }
  g->total = (int)off;
  return true;
}

constexpr int kWcfgThreads = 1024;

template <typename T>
struct WcfgScales {
  T v[SONAR_WCFG_MAX_LEVELS * 3];  // [level][orientation], fine -> coarse
};

// PAIR: a cluster of two CTAs shares a plane, for calls with fewer planes than half the SMs (SDXL batch
// 16 x 4 channels = 64 planes on 148 SMs). The two heavy phases split by rows with no exchange of detail
// coefficients: each CTA runs the level-1 analysis only for the coefficient rows its half of the final
// synthesis reads (half the plane plus an L/2 - 1 row halo) and keeps those detail rows to itself; the
// level-1 approximation rows are stored into BOTH CTAs' shared memory (distributed shared memory), after
// which each CTA runs the small coarser levels redundantly and synthesises its half of the output rows.
// ~63 % of the single-CTA work per CTA, one cluster barrier per plane.
template <typename T, int LT, bool PAIR>
__global__ void __launch_bounds__(kWcfgThreads, 1)
wcfg_fused_kernel(const float* __restrict__ in_a, const float* __restrict__ in_b, float* __restrict__ out,
                  const float* __restrict__ addend, float addend_scale, const float* __restrict__ x, float x_scale,
                  float recon_sign, int64_t planes, int H, int W, int mode, WcfgGeom g, T scale_ll, WcfgScales<T> scales,
                  Filters<T> f) {
  const T* scale_hi = scales.v;
  extern __shared__ __align__(16) unsigned char wcfg_smem[];
  T* sm = reinterpret_cast<T*>(wcfg_smem);
  const int L = LT > 0 ? LT : f.L, J = g.levels;
  const int tid = threadIdx.x, nthr = blockDim.x;
  const bool vec2_ok = (W & 1) == 0 && (((uintptr_t)out | (uintptr_t)addend | (uintptr_t)x) & 7u) == 0;
  // level-1 rows / final quads of this CTA (everything without PAIR)
  int q_lo = 0, q_hi = (H + 1) >> 1, r_lo = 0, r_hi = g.h[0], own_lo = 0, own_hi = g.h[0];
  T* peer_ll0 = nullptr;
  if (PAIR) {
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    wcfg_pair_rows(H, g.h[0], L, rank, &q_lo, &q_hi, &r_lo, &r_hi);
    int a, b, c, split;
    wcfg_pair_rows(H, g.h[0], L, 0, &a, &b, &c, &split);  // rank 0 owns approximation rows [0, split), rank 1 the rest
    own_lo = rank ? split : 0;
    own_hi = rank ? g.h[0] : split;
    peer_ll0 = cluster.map_shared_rank(sm, rank ^ 1) + g.ll_off[0];
    cluster.sync();  // no store into the partner's shared memory before the partner CTA has started executing
  }
  const int64_t plane_first = PAIR ? blockIdx.x >> 1 : blockIdx.x, plane_step = PAIR ? gridDim.x >> 1 : gridDim.x;
  for (int64_t plane = plane_first; plane < planes; plane += plane_step) {
    const float* pa = in_a + plane * (int64_t)H * W;
    const float* pb = in_b != nullptr ? in_b + plane * (int64_t)H * W : nullptr;
    // ---------------- analysis, fine -> coarse ----------------
    for (int j = 0; j < J; ++j) {
      const int Hin = j == 0 ? H : g.h[j - 1], Win = j == 0 ? W : g.w[j - 1];
      const int h = g.h[j], w = g.w[j];
      const int pad_t = (2 * (h - 1) - Hin + L) / 2, pad_l = (2 * (w - 1) - Win + L) / 2;
      T* ll = sm + g.ll_off[j];
      // level 1 keeps only this CTA's detail rows [ra, rb): pointer biased so that ky * w + kx still indexes it
      const int ra = j == 0 ? r_lo : 0, rb = j == 0 ? r_hi : h;
      T* hi = sm + g.hi_off[j] - ra * w;
      const int hw = (j == 0 ? g.hi0_rows : h) * w;
      // interior outputs first (fast path, all lanes alike), then the border ring: no warp ever runs
      // both code paths back to back
      int kx_lo = 0, kx_hi = -1, ky_lo = 0, ky_hi = -1;
      if (LT > 0) {
        kx_lo = (pad_l + 1) >> 1;
        ky_lo = (pad_t + 1) >> 1;
        kx_hi = Win - L + pad_l >= 0 ? min(w - 1, (Win - L + pad_l) >> 1) : -1;
        ky_hi = Hin - L + pad_t >= 0 ? min(h - 1, (Hin - L + pad_t) >> 1) : -1;
        const int iy_lo = max(ky_lo, ra), iy_hi = min(ky_hi, rb - 1);
        const int iw = max(0, kx_hi - kx_lo + 1), ih = max(0, iy_hi - iy_lo + 1);
        const bool vec2 = j == 0 && ((pad_l | W) & 1) == 0 && (((uintptr_t)in_a | (uintptr_t)in_b) & 7u) == 0 &&
                          ((H * (int64_t)W) & 1) == 0;
        for (int i = tid; i < iw * ih; i += nthr) {
          const int iy = i / iw, ix = i - iy * iw;
          const int ky = iy_lo + iy, kx = kx_lo + ix, idx = ky * w + kx;
          T c_ll, c_lh, c_hl, c_hh;
          if (j == 0) {
            if (vec2)
              analysis_interior<T, float, LT, true>(pa, pb, W, 2 * ky - pad_t, 2 * kx - pad_l, f, c_ll, c_lh, c_hl, c_hh);
            else
              analysis_interior<T, float, LT, false>(pa, pb, W, 2 * ky - pad_t, 2 * kx - pad_l, f, c_ll, c_lh, c_hl, c_hh);
          } else {
            analysis_interior<T, T, LT, false>(sm + g.ll_off[j - 1], nullptr, Win, 2 * ky - pad_t, 2 * kx - pad_l, f, c_ll,
                                               c_lh, c_hl, c_hh);
          }
          if (!PAIR || j > 0) {
            ll[idx] = c_ll;
          } else if (ky >= own_lo && ky < own_hi) {
            ll[idx] = c_ll;
            peer_ll0[idx] = c_ll;
          }
          hi[idx] = c_lh;
          hi[hw + idx] = c_hl;
          hi[2 * hw + idx] = c_hh;
        }
      }
      // border ring, enumerated densely (rows above the interior, rows below it, then the left / right
      // columns of the interior rows): a warp of the ring holds 32 border outputs, not 1 or 2
      {
        const bool has_interior = kx_hi >= kx_lo && ky_hi >= ky_lo;
        // rows [ra, rb) of this CTA: those above the interior (or all of them), those below it, and the
        // left / right columns of its interior rows
        const int top_hi = has_interior ? min(rb, ky_lo) : rb;
        const int bot_lo = has_interior ? max(ra, ky_hi + 1) : rb;
        const int mid_lo = max(ky_lo, ra), mid_rows = has_interior ? max(0, min(ky_hi, rb - 1) - mid_lo + 1) : 0;
        const int top = max(0, top_hi - ra) * w;
        const int bottom = max(0, rb - bot_lo) * w;
        const int right0 = kx_hi + 1;                              // first column right of the interior
        const int side = has_interior ? kx_lo + (w - right0) : 0;  // border outputs per interior row
        const int ring = top + bottom + mid_rows * side;
        for (int b = tid; b < ring; b += nthr) {
          int ky, kx;
          if (b < top) {
            const int r = b / w;
            ky = ra + r;
            kx = b - r * w;
          } else if (b < top + bottom) {
            const int r = (b - top) / w;
            ky = bot_lo + r;
            kx = (b - top) - r * w;
          } else {
            const int r = (b - top - bottom) / side, c = (b - top - bottom) - r * side;
            ky = mid_lo + r;
            kx = c < kx_lo ? c : right0 + (c - kx_lo);
          }
          const int idx = ky * w + kx;
          T c_ll, c_lh, c_hl, c_hh;
          if (j == 0)
            analysis_point<T, float, LT>(pa, pb, W, Hin, Win, ky, kx, pad_t, pad_l, mode, f, c_ll, c_lh, c_hl, c_hh);
          else
            analysis_point<T, T, LT>(sm + g.ll_off[j - 1], nullptr, Win, Hin, Win, ky, kx, pad_t, pad_l, mode, f, c_ll,
                                     c_lh, c_hl, c_hh);
          if (!PAIR || j > 0) {
            ll[idx] = c_ll;
          } else if (ky >= own_lo && ky < own_hi) {
            ll[idx] = c_ll;
            peer_ll0[idx] = c_ll;
          }
          hi[idx] = c_lh;
          hi[hw + idx] = c_hl;
          hi[2 * hw + idx] = c_hh;
        }
      }
      if (PAIR && j == 0)
        cg::this_cluster().sync();  // both halves of the level-1 approximation have landed in both CTAs
      else
        __syncthreads();
    }
    // ---------------- synthesis, coarse -> fine ----------------
    for (int j = J - 1; j >= 0; --j) {
      const int h = g.h[j], w = g.w[j];
      const bool coarsest = j == J - 1;
      const T* ll = sm + g.ll_off[j];
      const int ll_pitch = coarsest ? w : 2 * g.w[j + 1] - L + 2;
      const T s_ll = coarsest ? scale_ll : (T)1;
      const T s_lh = scale_hi[3 * j], s_hl = scale_hi[3 * j + 1], s_hh = scale_hi[3 * j + 2];
      const T* hi = sm + g.hi_off[j] - (j == 0 ? r_lo : 0) * w;  // level 1: this CTA's rows only (biased pointer)
      const int64_t hi_band = (int64_t)(j == 0 ? g.hi0_rows : h) * w;
      const int out_h = 2 * h - L + 2, out_w = 2 * w - L + 2;
      const int oh = j == 0 ? H : out_h, ow = j == 0 ? W : out_w;  // the final level is cropped to the input size
      const int qh = (oh + 1) >> 1, qw = (ow + 1) >> 1;
      T* rec = j == 0 ? nullptr : sm + g.ll_off[j - 1];
      const int qy_first = j == 0 ? q_lo : 0, qy_count = (j == 0 ? min(q_hi, qh) : qh) - qy_first;
      for (int idx = tid; idx < qy_count * qw; idx += nthr) {
        const int qr = idx / qw, qy = qy_first + qr, qx = idx - qr * qw;
        T o00 = 0, o01 = 0, o10 = 0, o11 = 0;
        synthesis_quad<T, LT>(ll, ll_pitch, hi, hi_band, w, h, w, qy, qx, s_ll, s_lh, s_hl, s_hh, f, o00, o01, o10, o11);
        const T vals[4] = {o00, o01, o10, o11};
        if (j == 0 && vec2_ok) {
          // final level, even width: the quad's two rows are 8-byte aligned float2 accesses
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            const int iy = 2 * qy + r;
            if (iy >= oh) continue;
            const int64_t o = (plane * H + iy) * (int64_t)W + 2 * qx;
            T r0 = vals[2 * r], r1 = vals[2 * r + 1];
            if (addend != nullptr) {
              const float2 ad = *reinterpret_cast<const float2*>(addend + o);
              r0 += (T)addend_scale * (T)ad.x;
              r1 += (T)addend_scale * (T)ad.y;
            }
            float2 res = make_float2((float)((T)recon_sign * r0), (float)((T)recon_sign * r1));
            if (x != nullptr) {
              const float2 xv = *reinterpret_cast<const float2*>(x + o);
              res.x = x_scale * xv.x + res.x;
              res.y = x_scale * xv.y + res.y;
            }
            *reinterpret_cast<float2*>(out + o) = res;
          }
          continue;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int iy = 2 * qy + (q >> 1), ix = 2 * qx + (q & 1);
          if (iy >= oh || ix >= ow) continue;
          if (j == 0) {
            const int64_t o = (plane * H + iy) * (int64_t)W + ix;
            T r = vals[q];
            if (addend != nullptr) r += (T)addend_scale * (T)addend[o];
            r = (T)recon_sign * r;
            float rf = (float)r;  // cast to the latent dtype first, then x - result (py/wavelet_cfg.py:729-747)
            if (x != nullptr) rf = x_scale * x[o] + rf;
            out[o] = rf;
          } else {
            rec[iy * out_w + ix] = vals[q];
          }
        }
      }
      __syncthreads();
    }
    // the partner stores the next plane's level-1 approximation into this CTA's LL[0], which held rec_1 until now
    if (PAIR) cg::this_cluster().sync();
  }
}

// ---------------------------------------------------------------------------------------------
// Strip kernel (the default for compile-time filter lengths): same plan as wcfg_fused_kernel -- one CTA or one
// 2-CTA cluster per plane, every coefficient in shared memory, one launch -- re-organised around what the ncu
// capture of that kernel showed (profiles/r02_c4_wcfg_fused_*): 2,640 instructions per warp of which 13 % were
// DFMA; the rest was per-output index arithmetic, boundary selects, constant-bank loads of run-time indexed
// geometry and the break tests of the synthesis taps. Here
//   * every approximation plane that feeds another analysis level is stored WITH its boundary extension
//     (a halo ring filled once per level from the interior), so inner loops never test a boundary;
//   * a thread owns a column of R consecutive outputs ("strip"): an input row is loaded and row-filtered once
//     and contributes to the LT/2 outputs whose windows contain it (32 + 16/R-ish DFMA per output instead of 48,
//     2 + 2/R row loads instead of 4), synthesis likewise shares the row combination between stacked quads;
//   * shared-memory windows are read as aligned pairs (LDS.128 for fp64): the window of output kx starts at the
//     even padded column 2 kx;
//   * the per-band CFG scales are applied once, where the coefficients are stored, not per synthesis tap;
//   * R is picked per phase so that the strips just cover the CTA's threads (one round where possible).
// ---------------------------------------------------------------------------------------------
struct WcfgStripGeom {
  int levels, pair, hi0_rows;
  int h[SONAR_WCFG_MAX_LEVELS], w[SONAR_WCFG_MAX_LEVELS];            // coefficient extents, fine -> coarse
  int ll_off[SONAR_WCFG_MAX_LEVELS], hi_off[SONAR_WCFG_MAX_LEVELS];  // element offsets into shared memory
  int ll_pitch[SONAR_WCFG_MAX_LEVELS], ll_rows[SONAR_WCFG_MAX_LEVELS];  // LL[j] as the padded input of level j+1
  int pad_t[SONAR_WCFG_MAX_LEVELS], pad_l[SONAR_WCFG_MAX_LEVELS];    // pads of level j's OWN input
  int stage_off, stage_rows, stage_pitch;  // level-1 input (a - b, extended) staged in shared memory; rows == 0: read from global
  int total;
};

static bool wcfg_strip_geometry(int H, int W, int L, int levels, WcfgStripGeom* g, bool pair, size_t elem_bytes = 0,
                                size_t smem_limit = 0) {
  if (levels < 1 || levels > SONAR_WCFG_MAX_LEVELS || H <= 0 || W <= 0) return false;
  g->levels = levels;
  g->pair = pair ? 1 : 0;
  int hh = H, ww = W;
  for (int j = 0; j < levels; ++j) {
    if (hh < L || ww < L) return false;  // halos are filled with ONE reflection / wrap (extend_index_near)
    const int h = (hh + L - 1) / 2, w = (ww + L - 1) / 2;
    g->h[j] = h;
    g->w[j] = w;
    g->pad_t[j] = (2 * (h - 1) - hh + L) / 2;
    g->pad_l[j] = (2 * (w - 1) - ww + L) / 2;
    hh = h;
    ww = w;
  }
  int64_t off = 0;
  for (int j = 0; j < levels; ++j) {
    int rows = g->h[j], pitch = g->w[j] + (g->w[j] & 1);
    if (j + 1 < levels) {  // padded input of level j+1; rec_{j+1} (2h' - L + 2 <= 2(h' - 1) + L) lands here on the way back
      rows = 2 * (g->h[j + 1] - 1) + L;
      pitch = 2 * (g->w[j + 1] - 1) + L;
    }
    g->ll_rows[j] = rows;
    g->ll_pitch[j] = pitch;
    g->ll_off[j] = (int)off;
    off += (int64_t)rows * pitch;
    off += off & 1;
    g->hi_off[j] = (int)off;
    int hi_rows = g->h[j];
    if (j == 0) {
      if (pair) {
        int q_lo, q_hi, r_lo, r_hi;
        wcfg_pair_rows(H, g->h[0], L, 0, &q_lo, &q_hi, &r_lo, &r_hi);
        hi_rows = r_hi - r_lo;
        wcfg_pair_rows(H, g->h[0], L, 1, &q_lo, &q_hi, &r_lo, &r_hi);
        if (r_hi - r_lo > hi_rows) hi_rows = r_hi - r_lo;
      }
      g->hi0_rows = hi_rows;
    }
    off += 3 * (int64_t)hi_rows * g->w[j];
    off += off & 1;
    if (off > (1 << 28)) return false;
  }
  // the level-1 input of this CTA's output rows (all of them, or the wider half of a pair), when it still fits
  g->stage_off = (int)off;
  g->stage_rows = 0;
  g->stage_pitch = 2 * (g->w[0] - 1) + L;
  int out_rows = g->h[0];
  if (pair) {
    int q_lo, q_hi, r_lo, r_hi;
    wcfg_pair_rows(H, g->h[0], L, 0, &q_lo, &q_hi, &r_lo, &r_hi);
    out_rows = r_hi - r_lo;
    wcfg_pair_rows(H, g->h[0], L, 1, &q_lo, &q_hi, &r_lo, &r_hi);
    if (r_hi - r_lo > out_rows) out_rows = r_hi - r_lo;
  }
  const int64_t stage = (int64_t)(2 * out_rows + L - 2) * g->stage_pitch;
  if (elem_bytes > 0 && (size_t)(off + stage) * elem_bytes <= smem_limit) {
    g->stage_rows = 2 * out_rows + L - 2;
    off += stage;
  }
  g->total = (int)off;
  return true;
}

template <typename T> struct PairOf;
template <> struct PairOf<double> { using type = double2; };
template <> struct PairOf<float> { using type = float2; };

// strip length that covers rows x cols outputs with the fewest (rounds x input rows per strip)
// (the accumulators of a strip live in registers: 4 R values, at 64 registers per thread)
template <typename T, int LT>
struct StripMax {
  static constexpr int value = sizeof(T) == 8 ? (LT <= 4 ? 2 : 1) : 2;
};

template <typename T, int LT>
__device__ __forceinline__ int pick_strip(int rows, int cols, int nthr, bool analysis) {
  constexpr int kMaxR = StripMax<T, LT>::value;
  int best = 1, best_cost = 1 << 30;
#pragma unroll
  for (int r = 1; r <= kMaxR; ++r) {
    const int items = ((rows + r - 1) / r) * cols;
    const int rounds = (items + nthr - 1) / nthr;
    const int cost = rounds * (analysis ? 2 * r + LT - 2 : r + LT / 2 - 1 + r);
    if (cost < best_cost) {
      best_cost = cost;
      best = r;
    }
  }
  return best;
}

// R stacked outputs (ky0 .. ky0+R-1, kx) of an analysis level whose input is a padded shared-memory plane: output
// (ky, kx) reads padded rows 2ky .. 2ky+LT-1, padded columns 2kx .. 2kx+LT-1. acc[o] = {ll, lh, hl, hh}.
template <typename T, int LT, int R>
__device__ __forceinline__ void analysis_strip_smem(const T* __restrict__ src, int pitch, int last_row, int ky0, int kx,
                                                    const Filters<T>& f, T (&acc)[R][4]) {
  using P = typename PairOf<T>::type;
#pragma unroll
  for (int o = 0; o < R; ++o) acc[o][0] = acc[o][1] = acc[o][2] = acc[o][3] = 0;
  const T* col = src + 2 * kx;
#pragma unroll
  for (int r = 0; r < 2 * R + LT - 2; ++r) {
    const int pr = min(2 * ky0 + r, last_row);  // rows past the plane belong to outputs the caller drops
    const T* p = col + pr * pitch;
    T lo = 0, hi = 0;
#pragma unroll
    for (int jx = 0; jx < LT; jx += 2) {
      const P v = *reinterpret_cast<const P*>(p + jx);
      lo = fma(f.a_lo[jx], (T)v.x, lo);
      hi = fma(f.a_hi[jx], (T)v.x, hi);
      lo = fma(f.a_lo[jx + 1], (T)v.y, lo);
      hi = fma(f.a_hi[jx + 1], (T)v.y, hi);
    }
#pragma unroll
    for (int o = 0; o < R; ++o) {
      const int jy = r - 2 * o;
      if (jy >= 0 && jy < LT) {
        acc[o][0] = fma(f.a_lo[jy], lo, acc[o][0]);
        acc[o][1] = fma(f.a_hi[jy], lo, acc[o][1]);  // high along H, low along W
        acc[o][2] = fma(f.a_lo[jy], hi, acc[o][2]);  // low along H, high along W
        acc[o][3] = fma(f.a_hi[jy], hi, acc[o][3]);
      }
    }
  }
}

// The same for level 1, whose input is value = a - b in global memory (fp32): the LT column offsets of the
// window are resolved once per strip (ox[jx] < 0: zero padding), the row once per input row.
template <typename T, int LT, int R>
__device__ __forceinline__ void analysis_strip_global(const float* __restrict__ pa, const float* __restrict__ pb, int H,
                                                      int W, int y_first, const int (&ox)[LT], int mode,
                                                      const Filters<T>& f, T (&acc)[R][4]) {
#pragma unroll
  for (int o = 0; o < R; ++o) acc[o][0] = acc[o][1] = acc[o][2] = acc[o][3] = 0;
#pragma unroll
  for (int r = 0; r < 2 * R + LT - 2; ++r) {
    const int sy = extend_index(min(y_first + r, 2 * H), H, mode);
    const int base = max(sy, 0) * W;
    T lo = 0, hi = 0;
#pragma unroll
    for (int jx = 0; jx < LT; ++jx) {
      const int o = base + max(ox[jx], 0);
      T v = (T)pa[o];
      if (pb != nullptr) v -= (T)pb[o];
      if ((sy | ox[jx]) < 0) v = 0;
      lo = fma(f.a_lo[jx], v, lo);
      hi = fma(f.a_hi[jx], v, hi);
    }
#pragma unroll
    for (int o = 0; o < R; ++o) {
      const int jy = r - 2 * o;
      if (jy >= 0 && jy < LT) {
        acc[o][0] = fma(f.a_lo[jy], lo, acc[o][0]);
        acc[o][1] = fma(f.a_hi[jy], lo, acc[o][1]);
        acc[o][2] = fma(f.a_lo[jy], hi, acc[o][2]);
        acc[o][3] = fma(f.a_hi[jy], hi, acc[o][3]);
      }
    }
  }
}

// R stacked 2x2 output quads (qy0 .. qy0+R-1, qx) of a synthesis level: quad q reads coefficient rows q .. q+LT/2-1,
// columns qx .. qx+LT/2-1 (always in range, see dwt2_synthesis_kernel); the row combination along W is shared by the
// quads that contain the row. Coefficients arrive already scaled. out[o] = {o00, o01, o10, o11}.
template <typename T, int LT, int R>
__device__ __forceinline__ void synthesis_strip(const T* __restrict__ pll, int ll_pitch, const T* __restrict__ phi,
                                                int band_stride, int hi_pitch, int last_row, int qy0, int qx,
                                                const Filters<T>& f, T (&out)[R][4]) {
  constexpr int half = LT / 2;
#pragma unroll
  for (int o = 0; o < R; ++o) out[o][0] = out[o][1] = out[o][2] = out[o][3] = 0;
#pragma unroll
  for (int rr = 0; rr < R + half - 1; ++rr) {
    const int ky = min(qy0 + rr, last_row);
    const T* rll = pll + ky * ll_pitch + qx;
    const T* rhi = phi + ky * hi_pitch + qx;
    T rl0 = 0, rl1 = 0, rh0 = 0, rh1 = 0;
#pragma unroll
    for (int ib = 0; ib < half; ++ib) {
      const int tx = LT - 2 - 2 * ib;
      const T v_ll = rll[ib], v_lh = rhi[ib], v_hl = rhi[band_stride + ib], v_hh = rhi[2 * band_stride + ib];
      rl0 = fma(f.s_lo[tx], v_ll, rl0);
      rl0 = fma(f.s_hi[tx], v_hl, rl0);
      rl1 = fma(f.s_lo[tx + 1], v_ll, rl1);
      rl1 = fma(f.s_hi[tx + 1], v_hl, rl1);
      rh0 = fma(f.s_lo[tx], v_lh, rh0);
      rh0 = fma(f.s_hi[tx], v_hh, rh0);
      rh1 = fma(f.s_lo[tx + 1], v_lh, rh1);
      rh1 = fma(f.s_hi[tx + 1], v_hh, rh1);
    }
#pragma unroll
    for (int o = 0; o < R; ++o) {
      const int ia = rr - o;
      if (ia >= 0 && ia < half) {
        const int ty = LT - 2 - 2 * ia;
        out[o][0] = fma(f.s_lo[ty], rl0, out[o][0]);
        out[o][0] = fma(f.s_hi[ty], rh0, out[o][0]);
        out[o][1] = fma(f.s_lo[ty], rl1, out[o][1]);
        out[o][1] = fma(f.s_hi[ty], rh1, out[o][1]);
        out[o][2] = fma(f.s_lo[ty + 1], rl0, out[o][2]);
        out[o][2] = fma(f.s_hi[ty + 1], rh0, out[o][2]);
        out[o][3] = fma(f.s_lo[ty + 1], rl1, out[o][3]);
        out[o][3] = fma(f.s_hi[ty + 1], rh1, out[o][3]);
      }
    }
  }
}

template <typename T>
struct WcfgStripCtx {
  T* sm;
  T* peer_ll0;     // PAIR: the partner CTA's LL[0]
  int own_lo, own_hi;  // PAIR: level-1 approximation rows this CTA publishes
};

// One analysis level over output rows [ra, rb): strips of R rows. j == 0 reads the fp32 inputs.
template <typename T, int LT, int R, bool PAIR>
__device__ __forceinline__ void wcfg_analysis_level(const WcfgStripGeom& g, int j, const float* pa, const float* pb, int H,
                                                    int W, int mode, int ra, int rb, const WcfgStripCtx<T>& c, T s_ll,
                                                    T s_lh, T s_hl, T s_hh, const Filters<T>& f) {
  const int h = g.h[j], w = g.w[j], J = g.levels;
  const int tid = threadIdx.x, nthr = blockDim.x;
  // where ll_j goes: the padded input plane of level j+1 (interior at (pad_t, pad_l) of that level), or the plain
  // coarsest plane
  const bool last = j == J - 1;
  const int dst_pitch = g.ll_pitch[j];
  const int dst_org = last ? 0 : g.pad_t[j + 1] * dst_pitch + g.pad_l[j + 1];
  T* ll = c.sm + g.ll_off[j] + dst_org;
  T* hi = c.sm + g.hi_off[j] - ra * w;  // biased: row ky of the CTA's band lives at (ky - ra)
  const int hw = (j == 0 ? g.hi0_rows : h) * w;
  const int strips = (rb - ra + R - 1) / R;
  const int items = strips * w;
  for (int i = tid; i < items; i += nthr) {
    const int s = i / w, kx = i - s * w;
    const int ky0 = ra + s * R;
    T acc[R][4];
    if (j == 0 && g.stage_rows > 0) {
      // staged rows start at padded row 2 ra
      analysis_strip_smem<T, LT, R>(c.sm + g.stage_off - 2 * ra * g.stage_pitch, g.stage_pitch, 2 * ra + g.stage_rows - 1, ky0,
                                    kx, f, acc);
    } else if (j == 0) {
      int ox[LT];
#pragma unroll
      for (int jx = 0; jx < LT; ++jx) ox[jx] = extend_index(2 * kx - g.pad_l[0] + jx, W, mode);
      analysis_strip_global<T, LT, R>(pa, pb, H, W, 2 * ky0 - g.pad_t[0], ox, mode, f, acc);
    } else {
      analysis_strip_smem<T, LT, R>(c.sm + g.ll_off[j - 1], g.ll_pitch[j - 1], g.ll_rows[j - 1] - 1, ky0, kx, f, acc);
    }
#pragma unroll
    for (int o = 0; o < R; ++o) {
      const int ky = ky0 + o;
      if (ky >= rb) break;
      const T v_ll = last ? acc[o][0] * s_ll : acc[o][0];
      const int at = ky * dst_pitch + kx;
      if (!PAIR || j > 0) {
        ll[at] = v_ll;
      } else if (ky >= c.own_lo && ky < c.own_hi) {
        ll[at] = v_ll;
        c.peer_ll0[dst_org + at] = v_ll;
      }
      const int idx = ky * w + kx;
      hi[idx] = acc[o][1] * s_lh;
      hi[hw + idx] = acc[o][2] * s_hl;
      hi[2 * hw + idx] = acc[o][3] * s_hh;
    }
  }
}

// One reflection / wrap only (callers guarantee the index is less than one signal length outside [0, n), which holds
// for every halo when the filter is not longer than the signal): no integer division anywhere.
__device__ __forceinline__ int extend_index_near(int i, int n, int mode) {
  if ((unsigned)i < (unsigned)n) return i;
  if (mode == SONAR_DWT_MODE_ZERO) return -1;
  int m;
  if (mode == SONAR_DWT_MODE_PERIODIC)
    m = i < 0 ? i + n : i - n;
  else if (mode == SONAR_DWT_MODE_REFLECT)
    m = i < 0 ? -i : 2 * (n - 1) - i;
  else
    m = i < 0 ? -1 - i : 2 * n - 1 - i;
  return min(max(m, 0), n - 1);
}

// Level-1 input of output rows [ra, ...): value = a - b converted once, stored with its boundary extension as padded
// rows 2 ra .. 2 ra + stage_rows - 1. Interior: one item = 4 consecutive pixels of a row (16-byte loads of a and b,
// two 16-byte shared stores); the pad_l / right halo columns are a second, tiny pass over scalar loads.
template <typename T>
__device__ __forceinline__ void wcfg_stage_input(const WcfgStripGeom& g, const float* __restrict__ pa,
                                                 const float* __restrict__ pb, int H, int W, int mode, int ra, T* sm) {
  using P = typename PairOf<T>::type;
  const int pitch = g.stage_pitch, pl = g.pad_l[0], pt = g.pad_t[0], rows = g.stage_rows;
  T* dst = sm + g.stage_off;
  const bool vec4 = (W & 3) == 0 && (pl & 1) == 0 && (((uintptr_t)pa | (uintptr_t)pb) & 15u) == 0;
  const int groups = vec4 ? W >> 2 : 0;
  for (int i = threadIdx.x; i < rows * groups; i += blockDim.x) {
    const int r = i / groups, c4 = (i - r * groups) << 2;
    const int sy = extend_index_near(min(2 * ra + r - pt, 2 * H - 1), H, mode);
    P v0, v1;
    v0.x = v0.y = v1.x = v1.y = 0;
    if (sy >= 0) {
      const float4 a4 = ld4(pa + sy * W + c4);
      v0.x = (T)a4.x; v0.y = (T)a4.y; v1.x = (T)a4.z; v1.y = (T)a4.w;
      if (pb != nullptr) {
        const float4 b4 = ld4(pb + sy * W + c4);
        v0.x -= (T)b4.x; v0.y -= (T)b4.y; v1.x -= (T)b4.z; v1.y -= (T)b4.w;
      }
    }
    T* d = dst + r * pitch + pl + c4;
    *reinterpret_cast<P*>(d) = v0;
    *reinterpret_cast<P*>(d + 2) = v1;
  }
  // what the vector pass did not cover: the halo columns, or every column when the plane is not 16-byte friendly
  const int first = vec4 ? 0 : 0, side = vec4 ? pitch - W : pitch;
  for (int i = threadIdx.x; i < rows * side; i += blockDim.x) {
    const int r = i / side, cc = i - r * side;
    const int pc = !vec4 ? cc : (cc < pl ? cc : W + cc);
    const int sy = extend_index_near(min(2 * ra + r - pt, 2 * H - 1), H, mode);
    const int sx = extend_index_near(pc - pl, W, mode);
    T v = 0;
    if ((sy | sx) >= 0) v = (T)pa[sy * W + sx] - (pb != nullptr ? (T)pb[sy * W + sx] : (T)0);
    dst[r * pitch + pc] = v;
  }
  (void)first;
}

// Boundary extension of LL[j] (the input plane of level j+1): every cell outside the interior copies the interior
// cell the extension mode maps it to (or zero).
template <typename T>
__device__ __forceinline__ void wcfg_fill_halo(const WcfgStripGeom& g, int j, int mode, T* sm) {
  const int rows = g.ll_rows[j], pitch = g.ll_pitch[j], h = g.h[j], w = g.w[j];
  const int pt = g.pad_t[j + 1], pl = g.pad_l[j + 1];
  T* ll = sm + g.ll_off[j];
  // ring enumerated densely: full rows above / below the interior, then the side columns of the interior rows
  const int top = pt * pitch, bottom = max(0, rows - pt - h) * pitch;
  const int side = pitch - w;
  const int ring = top + bottom + h * side;
  for (int b = threadIdx.x; b < ring; b += blockDim.x) {
    int pr, pc;
    if (b < top) {
      pr = b / pitch;
      pc = b - pr * pitch;
    } else if (b < top + bottom) {
      const int r = (b - top) / pitch;
      pr = pt + h + r;
      pc = (b - top) - r * pitch;
    } else {
      const int r = (b - top - bottom) / side, cc = (b - top - bottom) - r * side;
      pr = pt + r;
      pc = cc < pl ? cc : w + cc;
    }
    const int sy = extend_index_near(pr - pt, h, mode), sx = extend_index_near(pc - pl, w, mode);
    ll[pr * pitch + pc] = (sy | sx) < 0 ? (T)0 : ll[(sy + pt) * pitch + (sx + pl)];
  }
}

template <typename T, int LT, int R, bool FINAL>
__device__ __forceinline__ void wcfg_synthesis_level(const WcfgStripGeom& g, int j, int qy_first, int qy_end, int qw,
                                                     const T* sm_ll, int ll_pitch, T* rec, int rec_pitch, float* out,
                                                     const float* addend, float addend_scale, const float* x, float x_scale,
                                                     float recon_sign, int64_t plane, int H, int W, bool vec2_ok, int hi_bias,
                                                     T* sm, const Filters<T>& f) {
  using P = typename PairOf<T>::type;
  const int h = g.h[j], w = g.w[j];
  const T* hi = sm + g.hi_off[j] - hi_bias * w;
  const int band = (j == 0 ? g.hi0_rows : h) * w;
  const int strips = (qy_end - qy_first + R - 1) / R;
  const int items = strips * qw;
  const int oh = H, ow = W;  // FINAL only
  if (FINAL) {  // plane bases once, 32-bit offsets inside the plane
    const int64_t pbase = plane * (int64_t)H * W;
    out += pbase;
    if (addend != nullptr) addend += pbase;
    if (x != nullptr) x += pbase;
  }
  for (int i = threadIdx.x; i < items; i += blockDim.x) {
    const int s = i / qw, qx = i - s * qw;
    const int qy0 = qy_first + s * R;
    T o[R][4];
    synthesis_strip<T, LT, R>(sm_ll, ll_pitch, hi, band, w, h - 1, qy0, qx, f, o);
#pragma unroll
    for (int k = 0; k < R; ++k) {
      const int qy = qy0 + k;
      if (qy >= qy_end) break;
      if (!FINAL) {
        T* r0 = rec + (2 * qy) * rec_pitch + 2 * qx;
        P a, b;
        a.x = o[k][0]; a.y = o[k][1]; b.x = o[k][2]; b.y = o[k][3];
        *reinterpret_cast<P*>(r0) = a;
        *reinterpret_cast<P*>(r0 + rec_pitch) = b;
        continue;
      }
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int iy = 2 * qy + r;
        if (iy >= oh) continue;
        const int at = iy * W + 2 * qx;  // (a plane has fewer than 2^31 elements)
        T r0 = o[k][2 * r], r1 = o[k][2 * r + 1];
        if (vec2_ok) {
          if (addend != nullptr) {
            const float2 ad = *reinterpret_cast<const float2*>(addend + at);
            r0 += (T)addend_scale * (T)ad.x;
            r1 += (T)addend_scale * (T)ad.y;
          }
          // the reference casts the reconstruction to the latent dtype first, then forms x - result (py/wavelet_cfg.py:729-747)
          float2 res = make_float2((float)((T)recon_sign * r0), (float)((T)recon_sign * r1));
          if (x != nullptr) {
            const float2 xv = *reinterpret_cast<const float2*>(x + at);
            res.x = x_scale * xv.x + res.x;
            res.y = x_scale * xv.y + res.y;
          }
          *reinterpret_cast<float2*>(out + at) = res;
        } else {
          const T vals[2] = {r0, r1};
#pragma unroll
          for (int cx = 0; cx < 2; ++cx) {
            if (2 * qx + cx >= ow) continue;
            T v = vals[cx];
            if (addend != nullptr) v += (T)addend_scale * (T)addend[at + cx];
            float rf = (float)((T)recon_sign * v);
            if (x != nullptr) rf = x_scale * x[at + cx] + rf;
            out[at + cx] = rf;
          }
        }
      }
    }
  }
}

#define SONAR_STRIP_DISPATCH(R_, CALL)                                  \
  if ((R_) >= 2 && StripMax<T, LT>::value >= 2) {                       \
    constexpr int R = StripMax<T, LT>::value >= 2 ? 2 : 1;              \
    CALL;                                                               \
  } else {                                                              \
    constexpr int R = 1;                                                \
    CALL;                                                               \
  }

template <typename T, int LT, bool PAIR>
__global__ void __launch_bounds__(kWcfgThreads, 1)
wcfg_strip_kernel(const float* __restrict__ in_a, const float* __restrict__ in_b, float* __restrict__ out,
                  const float* __restrict__ addend, float addend_scale, const float* __restrict__ x, float x_scale,
                  float recon_sign, int64_t planes, int H, int W, int mode, WcfgStripGeom g, T scale_ll, WcfgScales<T> scales,
                  Filters<T> f) {
  extern __shared__ __align__(16) unsigned char wcfg_smem[];
  T* sm = reinterpret_cast<T*>(wcfg_smem);
  const int J = g.levels, nthr = blockDim.x;
  const bool vec2_ok = (W & 1) == 0 && (((uintptr_t)out | (uintptr_t)addend | (uintptr_t)x) & 7u) == 0;
  int q_lo = 0, q_hi = (H + 1) >> 1, r_lo = 0, r_hi = g.h[0];
  WcfgStripCtx<T> c{sm, nullptr, 0, g.h[0]};
  if (PAIR) {
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    wcfg_pair_rows(H, g.h[0], LT, rank, &q_lo, &q_hi, &r_lo, &r_hi);
    int a, b, cc, split;
    wcfg_pair_rows(H, g.h[0], LT, 0, &a, &b, &cc, &split);  // rank 0 publishes approximation rows [0, split), rank 1 the rest
    c.own_lo = rank ? split : 0;
    c.own_hi = rank ? g.h[0] : split;
    c.peer_ll0 = cluster.map_shared_rank(sm, rank ^ 1) + g.ll_off[0];
    cluster.sync();  // no store into the partner's shared memory before the partner CTA has started executing
  }
  const int64_t plane_first = PAIR ? blockIdx.x >> 1 : blockIdx.x, plane_step = PAIR ? gridDim.x >> 1 : gridDim.x;
  for (int64_t plane = plane_first; plane < planes; plane += plane_step) {
    const float* pa = in_a + plane * (int64_t)H * W;
    const float* pb = in_b != nullptr ? in_b + plane * (int64_t)H * W : nullptr;
    if (g.stage_rows > 0) {
      wcfg_stage_input<T>(g, pa, pb, H, W, mode, r_lo, sm);
      __syncthreads();
    }
    // ---------------- analysis, fine -> coarse ----------------
    for (int j = 0; j < J; ++j) {
      const int ra = j == 0 ? r_lo : 0, rb = j == 0 ? r_hi : g.h[j];
      const T s_ll = j == J - 1 ? scale_ll : (T)1;
      const T s_lh = scales.v[3 * j], s_hl = scales.v[3 * j + 1], s_hh = scales.v[3 * j + 2];
      const int strip = pick_strip<T, LT>(rb - ra, g.w[j], nthr, true);
      SONAR_STRIP_DISPATCH(strip, (wcfg_analysis_level<T, LT, R, PAIR>(g, j, pa, pb, H, W, mode, ra, rb, c, s_ll, s_lh, s_hl, s_hh, f)))
      if (PAIR && j == 0)
        cg::this_cluster().sync();  // both parts of the level-1 approximation have landed in both CTAs
      else
        __syncthreads();
      if (j + 1 < J) {
        wcfg_fill_halo<T>(g, j, mode, sm);
        __syncthreads();
      }
    }
    // ---------------- synthesis, coarse -> fine ----------------
    for (int j = J - 1; j >= 0; --j) {
      const int h = g.h[j], w = g.w[j];
      const int out_h = 2 * h - LT + 2, out_w = 2 * w - LT + 2;
      const int oh = j == 0 ? H : out_h, ow = j == 0 ? W : out_w;  // the final level is cropped to the input size
      const int qh = (oh + 1) >> 1, qw = (ow + 1) >> 1;
      const int qy_first = j == 0 ? q_lo : 0, qy_end = j == 0 ? min(q_hi, qh) : qh;
      const T* ll = sm + g.ll_off[j];  // the coarsest approximation, or rec_{j+1} stored at the origin of LL[j]
      const int strip = pick_strip<T, LT>(qy_end - qy_first, qw, nthr, false);
      if (j == 0) {
        SONAR_STRIP_DISPATCH(strip, (wcfg_synthesis_level<T, LT, R, true>(g, j, qy_first, qy_end, qw, ll, g.ll_pitch[0], nullptr, 0, out, addend, addend_scale, x, x_scale, recon_sign, plane, H, W, vec2_ok, r_lo, sm, f)))
      } else {
        SONAR_STRIP_DISPATCH(strip, (wcfg_synthesis_level<T, LT, R, false>(g, j, qy_first, qy_end, qw, ll, g.ll_pitch[j], sm + g.ll_off[j - 1], g.ll_pitch[j - 1], nullptr, nullptr, 0.f, nullptr, 0.f, 0.f, plane, H, W, false, 0, sm, f)))
      }
      __syncthreads();
    }
    // the partner stores the next plane's level-1 approximation into this CTA's LL[0], which held rec_1 until now
    if (PAIR) cg::this_cluster().sync();
  }
}

template <typename T>
Filters<T> make_filters(const SonarWaveletFilters* wf) {
  Filters<T> f;
  f.L = wf->length;
  for (int i = 0; i < SONAR_DWT_MAX_TAPS; ++i) {
    const bool in = i < wf->length;
    // analysis filters are stored in correlation form: reversed decomposition filters
    f.a_lo[i] = in ? (T)wf->dec_lo[wf->length - 1 - i] : (T)0;
    f.a_hi[i] = in ? (T)wf->dec_hi[wf->length - 1 - i] : (T)0;
    f.s_lo[i] = in ? (T)wf->rec_lo[i] : (T)0;
    f.s_hi[i] = in ? (T)wf->rec_hi[i] : (T)0;
  }
  return f;
}

// compile-time filter lengths with their own instantiation (haar, db2, db3, db4); others run the generic loops
#define SONAR_DISPATCH_TAPS(len, CALL) \
  switch (len) {                       \
    case 2: CALL(2); break;            \
    case 4: CALL(4); break;            \
    case 6: CALL(6); break;            \
    case 8: CALL(8); break;            \
    default: CALL(0); break;           \
  }

template <typename T>
int launch_analysis(const SonarDwtAnalysisParams& p, cudaStream_t stream) {
  const Filters<T> f = make_filters<T>(&p.filters);
  const int64_t total = p.planes * (int64_t)p.h * p.w;
  const int grid = streaming_grid(total, kBlock, 4);
  if ((int64_t)(p.in_stride_h > p.H ? p.in_stride_h : p.H) * p.W >= (1ll << 31)) return (int)cudaErrorInvalidValue;
#define ANALYSIS_F32(LT)                                                                                            \
  dwt2_analysis_kernel<T, float, LT><<<grid, kBlock, 0, stream>>>((const float*)p.in_a, (const float*)p.in_b, (T*)p.ll, \
                                                                  (T*)p.hi, p.planes, p.H, p.W, p.H, p.h, p.w, p.mode, f)
#define ANALYSIS_T(LT)                                                                                          \
  dwt2_analysis_kernel<T, T, LT><<<grid, kBlock, 0, stream>>>((const T*)p.in_a, (const T*)p.in_b, (T*)p.ll, (T*)p.hi, \
                                                              p.planes, p.H, p.W, p.in_stride_h > 0 ? p.in_stride_h : p.H, \
                                                              p.h, p.w, p.mode, f)
  if (p.in_is_f32) {
    SONAR_DISPATCH_TAPS(p.filters.length, ANALYSIS_F32)
  } else {
    SONAR_DISPATCH_TAPS(p.filters.length, ANALYSIS_T)
  }
#undef ANALYSIS_F32
#undef ANALYSIS_T
  SONAR_LAUNCH_CHECK();
  return 0;
}

template <typename T>
int launch_synthesis(const SonarDwtSynthesisParams& p, cudaStream_t stream) {
  const Filters<T> f = make_filters<T>(&p.filters);
  SynthSet<T> sets[2];
  for (int s = 0; s < 2; ++s) {
    sets[s].ll = (const T*)p.ll[s];
    sets[s].hi = (const T*)p.hi[s];
    sets[s].ll_stride_h = p.ll_rows[s];
    sets[s].ll_stride_w = p.ll_cols[s];
    sets[s].s_ll = (T)p.scales[s][0];
    sets[s].s_lh = (T)p.scales[s][1];
    sets[s].s_hl = (T)p.scales[s][2];
    sets[s].s_hh = (T)p.scales[s][3];
  }
  const int out_h = 2 * p.h - p.filters.length + 2, out_w = 2 * p.w - p.filters.length + 2;
  const bool final_level = p.out_f32 != nullptr;
  if (final_level && (p.crop_h > out_h || p.crop_w > out_w)) return (int)cudaErrorInvalidValue;
  if ((int64_t)p.ll_rows[0] * p.ll_cols[0] >= (1ll << 31)) return (int)cudaErrorInvalidValue;
  const int oh = final_level ? p.crop_h : out_h, ow = final_level ? p.crop_w : out_w;
  const int64_t total = p.planes * (int64_t)((oh + 1) / 2) * ((ow + 1) / 2);  // one thread per 2x2 output quad
  const int grid = streaming_grid(total, kBlock, 4);
#define SYNTHESIS(LT)                                                                                               \
  dwt2_synthesis_kernel<T, LT><<<grid, kBlock, 0, stream>>>(sets[0], sets[1], p.n_sets, p.planes, p.h, p.w, out_h, out_w, \
                                                            (T*)p.out, p.out_f32, p.crop_h, p.crop_w, p.addend,         \
                                                            p.addend_scale, p.x, p.x_scale, p.recon_sign, f)
  SONAR_DISPATCH_TAPS(p.filters.length, SYNTHESIS)
#undef SYNTHESIS
  SONAR_LAUNCH_CHECK();
  return 0;
}

template <typename T>
int launch_wcfg_fused(const SonarWcfgFusedParams& p, const WcfgGeom& g, cudaStream_t stream) {
  const Filters<T> f = make_filters<T>(&p.filters);
  const size_t smem = (size_t)g.total * sizeof(T);
  WcfgScales<T> sc;
  for (int j = 0; j < SONAR_WCFG_MAX_LEVELS; ++j)
    for (int o = 0; o < 3; ++o) sc.v[3 * j + o] = j < p.levels ? (T)p.scale_hi[j][o] : (T)0;
  const DeviceInfo& di = device_info();
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[1];
  cfg.blockDim = dim3(kWcfgThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  if (g.pair) {  // one cluster of two CTAs per plane
    const int64_t clusters = p.planes < di.sm_count / 2 ? p.planes : di.sm_count / 2;
    cfg.gridDim = dim3((unsigned)(2 * clusters));
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
  } else {
    cfg.gridDim = dim3((unsigned)(p.planes < di.sm_count ? p.planes : di.sm_count));
  }
  const T scale_ll = (T)p.scale_ll;
  // compile-time filter lengths whose padded planes fit: the strip kernel (SONAR_B200_WCFG_STRIP=0: the first version)
  WcfgStripGeom sg;
  static const bool strip_enabled = [] {
    const char* e = getenv("SONAR_B200_WCFG_STRIP");
    return e == nullptr || e[0] != '0';
  }();
  const int L = p.filters.length;
  if (strip_enabled && (L == 2 || L == 4 || L == 6 || L == 8) && wcfg_strip_geometry(p.H, p.W, L, p.levels, &sg, g.pair != 0, sizeof(T), (size_t)di.max_smem_optin) &&
      (size_t)sg.total * sizeof(T) <= (size_t)di.max_smem_optin) {
    cfg.dynamicSmemBytes = (size_t)sg.total * sizeof(T);
#define WCFG_STRIP_KERNEL(LT, PAIR)                                                                                       \
  do {                                                                                                                   \
    SONAR_CUDA_TRY(cudaFuncSetAttribute(wcfg_strip_kernel<T, LT, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize,       \
                                        (int)cfg.dynamicSmemBytes));                                                      \
    SONAR_CUDA_TRY(cudaLaunchKernelEx(&cfg, wcfg_strip_kernel<T, LT, PAIR>, p.in_a, p.in_b, p.out, p.addend, p.addend_scale, \
                                      p.x, p.x_scale, p.recon_sign, p.planes, p.H, p.W, p.mode, sg, scale_ll, sc, f));     \
  } while (0)
#define WCFG_STRIP(LT)                 \
  do {                                 \
    if (g.pair)                        \
      WCFG_STRIP_KERNEL(LT, true);     \
    else                               \
      WCFG_STRIP_KERNEL(LT, false);    \
  } while (0)
    switch (L) {
      case 2: WCFG_STRIP(2); break;
      case 4: WCFG_STRIP(4); break;
      case 6: WCFG_STRIP(6); break;
      default: WCFG_STRIP(8); break;
    }
#undef WCFG_STRIP
#undef WCFG_STRIP_KERNEL
    SONAR_LAUNCH_CHECK();
    return 0;
  }
#define WCFG_FUSED_KERNEL(LT, PAIR)                                                                                      \
  do {                                                                                                                   \
    SONAR_CUDA_TRY(cudaFuncSetAttribute(wcfg_fused_kernel<T, LT, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    SONAR_CUDA_TRY(cudaLaunchKernelEx(&cfg, wcfg_fused_kernel<T, LT, PAIR>, p.in_a, p.in_b, p.out, p.addend, p.addend_scale, \
                                      p.x, p.x_scale, p.recon_sign, p.planes, p.H, p.W, p.mode, g, scale_ll, sc, f));      \
  } while (0)
#define WCFG_FUSED(LT)                 \
  do {                                 \
    if (g.pair)                        \
      WCFG_FUSED_KERNEL(LT, true);     \
    else                               \
      WCFG_FUSED_KERNEL(LT, false);    \
  } while (0)
  SONAR_DISPATCH_TAPS(p.filters.length, WCFG_FUSED)
#undef WCFG_FUSED
#undef WCFG_FUSED_KERNEL
  SONAR_LAUNCH_CHECK();
  return 0;
}

}  // namespace sonar

extern "C" {

int sonar_dwt_coeff_len(int n, int filter_len) { return (n + filter_len - 1) / 2; }

int sonar_dwt2_analysis(const SonarDwtAnalysisParams* params, void* stream) {
  using namespace sonar;
  if (params == nullptr) return (int)cudaErrorInvalidValue;
  const SonarDwtAnalysisParams& p = *params;
  if (p.planes <= 0) return 0;
  if (p.filters.length < 2 || p.filters.length > SONAR_DWT_MAX_TAPS || (p.filters.length & 1)) return (int)cudaErrorInvalidValue;
  if (p.in_a == nullptr || p.ll == nullptr || p.hi == nullptr || p.H <= 0 || p.W <= 0) return (int)cudaErrorInvalidValue;
  const bool per = p.mode == SONAR_DWT_MODE_PERIODIZATION;  // non-expansive: ceil(n / 2) coefficients
  if (p.h != (per ? (p.H + 1) / 2 : sonar_dwt_coeff_len(p.H, p.filters.length)) ||
      p.w != (per ? (p.W + 1) / 2 : sonar_dwt_coeff_len(p.W, p.filters.length)))
    return (int)cudaErrorInvalidValue;
  return p.use_f64 ? launch_analysis<double>(p, (cudaStream_t)stream) : launch_analysis<float>(p, (cudaStream_t)stream);
}

int64_t sonar_wcfg_fused_smem_bytes(int H, int W, int filter_len, int levels, int use_f64) {
  sonar::WcfgGeom g;
  if (filter_len < 2 || filter_len > SONAR_DWT_MAX_TAPS || (filter_len & 1)) return 0;
  if (!sonar::wcfg_geometry(H, W, filter_len, levels, &g)) return 0;
  const int64_t bytes = (int64_t)g.total * (use_f64 ? 8 : 4);
  return bytes <= (int64_t)sonar::device_info().max_smem_optin ? bytes : 0;
}

int sonar_wcfg_fused(const SonarWcfgFusedParams* params, void* stream) {
  using namespace sonar;
  if (params == nullptr) return (int)cudaErrorInvalidValue;
  const SonarWcfgFusedParams& p = *params;
  if (p.planes <= 0) return 0;
  if (p.in_a == nullptr || p.out == nullptr || p.mode == SONAR_DWT_MODE_PERIODIZATION) return (int)cudaErrorInvalidValue;
  if (sonar_wcfg_fused_smem_bytes(p.H, p.W, p.filters.length, p.levels, p.use_f64) <= 0) return (int)cudaErrorInvalidValue;
  WcfgGeom g;
  // fewer planes than half the SMs: two CTAs (a cluster) per plane
  const bool pair = 2 * p.planes <= device_info().sm_count && p.H >= 4 * p.filters.length;
  wcfg_geometry(p.H, p.W, p.filters.length, p.levels, &g, pair);
  // the final level must reconstruct at least the input extent (true for every orthogonal bank here)
  if (2 * g.h[0] - p.filters.length + 2 < p.H || 2 * g.w[0] - p.filters.length + 2 < p.W) return (int)cudaErrorInvalidValue;
  return p.use_f64 ? launch_wcfg_fused<double>(p, g, (cudaStream_t)stream) : launch_wcfg_fused<float>(p, g, (cudaStream_t)stream);
}

int sonar_dwt2_synthesis(const SonarDwtSynthesisParams* params, void* stream) {
  using namespace sonar;
  if (params == nullptr) return (int)cudaErrorInvalidValue;
  const SonarDwtSynthesisParams& p = *params;
  if (p.planes <= 0) return 0;
  if (p.filters.length < 2 || p.filters.length > SONAR_DWT_MAX_TAPS || (p.filters.length & 1)) return (int)cudaErrorInvalidValue;
  if (p.n_sets < 1 || p.n_sets > 2 || p.h <= 0 || p.w <= 0) return (int)cudaErrorInvalidValue;
  if (p.out == nullptr && p.out_f32 == nullptr) return (int)cudaErrorInvalidValue;
  for (int s = 0; s < p.n_sets; ++s)
    if (p.ll[s] == nullptr || p.hi[s] == nullptr || p.ll_rows[s] < p.h || p.ll_cols[s] < p.w) return (int)cudaErrorInvalidValue;
  return p.use_f64 ? launch_synthesis<double>(p, (cudaStream_t)stream)
                   : launch_synthesis<float>(p, (cudaStream_t)stream);
}

int sonar_dwt2_synthesis_per(const SonarDwtSynthesisParams* params, void* stream) {
  using namespace sonar;
  if (params == nullptr) return (int)cudaErrorInvalidValue;
  const SonarDwtSynthesisParams& p = *params;
  if (p.planes <= 0 || p.h <= 0 || p.w <= 0) return 0;
  if (p.n_sets != 1 || p.out == nullptr || p.out_f32 != nullptr || p.ll[0] == nullptr || p.hi[0] == nullptr)
    return (int)cudaErrorInvalidValue;
  const int L = p.filters.length;
  if (L < 2 || (L & 1) || L > SONAR_DWT_MAX_TAPS || p.ll_rows[0] < p.h || p.ll_cols[0] < p.w) return (int)cudaErrorInvalidValue;
  const int64_t total = p.planes * (int64_t)(2 * p.h) * (2 * p.w);
  const int grid = streaming_grid(total, kBlock, 4);
  if (p.use_f64) {
    SynthSet<double> c{(const double*)p.ll[0], (const double*)p.hi[0], p.ll_rows[0], p.ll_cols[0],
                       p.scales[0][0], p.scales[0][1], p.scales[0][2], p.scales[0][3]};
    dwt2_per_synthesis_kernel<double><<<grid, kBlock, 0, (cudaStream_t)stream>>>(c, p.planes, p.h, p.w, (double*)p.out,
                                                                                make_filters<double>(&p.filters));
  } else {
    SynthSet<float> c{(const float*)p.ll[0], (const float*)p.hi[0], p.ll_rows[0], p.ll_cols[0],
                      (float)p.scales[0][0], (float)p.scales[0][1], (float)p.scales[0][2], (float)p.scales[0][3]};
    dwt2_per_synthesis_kernel<float><<<grid, kBlock, 0, (cudaStream_t)stream>>>(c, p.planes, p.h, p.w, (float*)p.out,
                                                                               make_filters<float>(&p.filters));
  }
  SONAR_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
