// Shared-memory-resident 2-D spectral shaping: [rfft2 ->] gain mask -> irfft2, one CTA per latent
// plane, one HBM read and one HBM write per plane.
//
// Reference: PowerNoiseItem sampler py/nodes/powernoise.py:355-366 (irfft2(noise_rfft * filter,
// s=(H,W), norm="ortho"), optional rfft2 front end :357-359), OneFNoiseGenerator.generate
// py/noise_generation.py:737-759 and GreenTestNoiseGenerator.generate :694-704 (fft -> gain -> ifft,
// real part; per-plane 2-D because the gain is constant over batch and channel).
//
// irfft2 semantics for the NON-Hermitian half spectrum the reference feeds it (SURVEY.md section 7,
// hard part 6): complex inverse DFT along H for each of the W/2+1 columns, then a c2r transform
// along W that ignores the imaginary parts of the k=0 (and, for even W, k=W/2) bins.
//
// Layout: the half spectrum S[H][Wh_pad] (complex64, Wh_pad odd so column walks are conflict-free)
// stays in shared memory -- or in an L2-resident global scratch when a plane is too large (256^2:
// 258 KB > 227 KB) -- and every 1-D FFT is a warp-level mixed-radix Stockham transform in a per-warp
// ping-pong scratch. Rows are processed two at a time (real pair <-> one complex FFT).
#include <cooperative_groups.h>

#include <algorithm>
#include <cstdlib>
#include <cstdint>
#include <type_traits>

#include "common.cuh"
#include "../../include/sonar_b200.h"

namespace sonar {

constexpr int kFftThreads = 512;
constexpr int kFftWarps = kFftThreads / 32;

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cconj(float2 a) { return make_float2(a.x, -a.y); }

struct FftPlan {
  int n;
  int n_factors;
  int factors[SONAR_FFT_MAX_FACTORS];
  int ns[SONAR_FFT_MAX_FACTORS];       // product of the radices before stage f
  int tab_off[SONAR_FFT_MAX_FACTORS];  // offset of stage f in the butterfly table
  int tab_size;
};

// tw[k] = exp(-2*pi*i*k/n); inverse transforms conjugate on the fly.
__device__ __forceinline__ float2 twiddle(const float2* __restrict__ tw, int idx, bool inverse) {
  const float2 w = tw[idx];
  return inverse ? make_float2(w.x, -w.y) : w;
}

// multiply by -i (forward) / +i (inverse)
__device__ __forceinline__ float2 rot90(float2 d, bool inverse) {
  return inverse ? make_float2(-d.y, d.x) : make_float2(d.y, -d.x);
}

// Per-stage butterfly table, built once per CTA: for butterfly j of stage f
//   .x = output base  (j / Ns) * Ns * R + (j % Ns)      .y = twiddle base (j % Ns) * N / (Ns * R)
// so the transform loops contain no integer division.
__device__ void build_fft_table(const FftPlan& plan, ushort2* __restrict__ tab) {
  for (int f = 0; f < plan.n_factors; ++f) {
    const int R = plan.factors[f], Ns = plan.ns[f], stride = plan.n / R, tw_step = plan.n / (Ns * R);
    for (int j = threadIdx.x; j < stride; j += blockDim.x) {
      const int k = j % Ns;
      tab[plan.tab_off[f] + j] = make_ushort2((unsigned short)((j / Ns) * Ns * R + k), (unsigned short)(k * tw_step));
    }
  }
}

// Warp-cooperative Stockham autosort FFT of length plan.n. Input in `a`; returns the buffer that
// holds the result (a or b). Unnormalised. All 32 lanes must call.
template <bool INVERSE>
__device__ float2* warp_fft(float2* a, float2* b, const FftPlan& plan, const float2* __restrict__ tw,
                            const ushort2* __restrict__ tab, int lane) {
  const int N = plan.n;
  float2* src = a;
  float2* dst = b;
  for (int f = 0; f < plan.n_factors; ++f) {
    const int R = plan.factors[f];
    const int Ns = plan.ns[f];
    const int stride = N / R;  // distance between the R inputs of one butterfly
    const ushort2* t = tab + plan.tab_off[f];
    if (R == 4) {
      for (int j = lane; j < stride; j += 32) {
        const ushort2 ix = t[j];
        const int tb = ix.y;
        const float2 v0 = src[j];
        const float2 v1 = cmul(src[j + stride], twiddle(tw, tb, INVERSE));
        const float2 v2 = cmul(src[j + 2 * stride], twiddle(tw, 2 * tb, INVERSE));
        const float2 v3 = cmul(src[j + 3 * stride], twiddle(tw, 3 * tb, INVERSE));
        const float2 a0 = cadd(v0, v2), a1 = csub(v0, v2), a2 = cadd(v1, v3);
        const float2 a3 = rot90(csub(v1, v3), INVERSE);
        float2* o = dst + ix.x;
        o[0] = cadd(a0, a2);
        o[Ns] = cadd(a1, a3);
        o[2 * Ns] = csub(a0, a2);
        o[3 * Ns] = csub(a1, a3);
      }
    } else if (R == 2) {
      for (int j = lane; j < stride; j += 32) {
        const ushort2 ix = t[j];
        const float2 v0 = src[j];
        const float2 v1 = cmul(src[j + stride], twiddle(tw, ix.y, INVERSE));
        dst[ix.x] = cadd(v0, v1);
        dst[ix.x + Ns] = csub(v0, v1);
      }
    } else if (R == 3) {
      for (int j = lane; j < stride; j += 32) {
        const ushort2 ix = t[j];
        const int tb = ix.y;
        const float2 v0 = src[j];
        const float2 v1 = cmul(src[j + stride], twiddle(tw, tb, INVERSE));
        const float2 v2 = cmul(src[j + 2 * stride], twiddle(tw, 2 * tb, INVERSE));
        const float2 t1 = cadd(v1, v2);
        const float2 t2 = make_float2(v0.x - 0.5f * t1.x, v0.y - 0.5f * t1.y);
        const float2 d = csub(v1, v2);
        const float2 t3 = rot90(make_float2(0.86602540378443865f * d.x, 0.86602540378443865f * d.y), INVERSE);
        float2* o = dst + ix.x;
        o[0] = cadd(v0, t1);
        o[Ns] = cadd(t2, t3);
        o[2 * Ns] = csub(t2, t3);
      }
    } else if (R == 5) {
      constexpr float c1 = 0.30901699437494742f, c2 = -0.80901699437494742f;
      constexpr float s1 = 0.95105651629515357f, s2 = 0.58778525229247313f;
      for (int j = lane; j < stride; j += 32) {
        const ushort2 ix = t[j];
        const int tb = ix.y;
        const float2 v0 = src[j];
        const float2 v1 = cmul(src[j + stride], twiddle(tw, tb, INVERSE));
        const float2 v2 = cmul(src[j + 2 * stride], twiddle(tw, 2 * tb, INVERSE));
        const float2 v3 = cmul(src[j + 3 * stride], twiddle(tw, 3 * tb, INVERSE));
        const float2 v4 = cmul(src[j + 4 * stride], twiddle(tw, 4 * tb, INVERSE));
        const float2 a1 = cadd(v1, v4), a2 = cadd(v2, v3), b1 = csub(v1, v4), b2 = csub(v2, v3);
        const float2 m1 = make_float2(v0.x + c1 * a1.x + c2 * a2.x, v0.y + c1 * a1.y + c2 * a2.y);
        const float2 m2 = make_float2(v0.x + c2 * a1.x + c1 * a2.x, v0.y + c2 * a1.y + c1 * a2.y);
        const float2 n1 = rot90(make_float2(s1 * b1.x + s2 * b2.x, s1 * b1.y + s2 * b2.y), INVERSE);
        const float2 n2 = rot90(make_float2(s2 * b1.x - s1 * b2.x, s2 * b1.y - s1 * b2.y), INVERSE);
        float2* o = dst + ix.x;
        o[0] = make_float2(v0.x + a1.x + a2.x, v0.y + a1.y + a2.y);
        o[Ns] = cadd(m1, n1);
        o[2 * Ns] = cadd(m2, n2);
        o[3 * Ns] = csub(m2, n2);
        o[4 * Ns] = csub(m1, n1);
      }
    } else {
      // generic (prime) radix: each output is a direct R-term sum; one table lookup per term
      const int rot = N / R;
      const int tw_step = N / (Ns * R);
      for (int o = lane; o < N; o += 32) {
        const int j = o % stride;   // butterfly
        const int tt = o / stride;  // which output of it
        const int k = j % Ns;
        const int step = (k * tw_step + tt * rot) % N;
        float2 acc = src[j];
        int idx = 0;
        for (int s2 = 1; s2 < R; ++s2) {
          idx += step;
          if (idx >= N) idx -= N;
          acc = cadd(acc, cmul(src[j + s2 * stride], twiddle(tw, idx, INVERSE)));
        }
        dst[(j / Ns) * Ns * R + k + tt * Ns] = acc;
      }
    }
    __syncwarp();
    float2* tmp = src;
    src = dst;
    dst = tmp;
  }
  return src;
}

struct SpectralLaunch {
  SonarSpectralParams p;
  FftPlan plan_h;
  FftPlan plan_w;
  int wh;       // W/2 + 1
  int wh_pad;   // odd
  int nmax;     // max(H, W)
  int spectrum_in_smem;
};

__global__ void __launch_bounds__(kFftThreads, 2)
spectral_plane_kernel(SpectralLaunch L) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const SonarSpectralParams& p = L.p;
  const int H = p.H, W = p.W, Wh = L.wh, WhP = L.wh_pad;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  float2* tw_h = reinterpret_cast<float2*>(smem_raw);
  float2* tw_w = tw_h + H;
  float2* warp_buf = tw_w + W + (size_t)warp * 2 * L.nmax;
  float2* bufA = warp_buf;
  float2* bufB = warp_buf + L.nmax;
  float2* after_bufs = tw_w + W + (size_t)kFftWarps * 2 * L.nmax;
  ushort2* tab_h = reinterpret_cast<ushort2*>(after_bufs);
  ushort2* tab_w = tab_h + L.plan_h.tab_size;
  float2* S = L.spectrum_in_smem
                  ? reinterpret_cast<float2*>(tab_w + L.plan_w.tab_size + ((L.plan_h.tab_size + L.plan_w.tab_size) & 1))
                  : (reinterpret_cast<float2*>(p.scratch) + (size_t)blockIdx.x * H * WhP);
  build_fft_table(L.plan_h, tab_h);
  build_fft_table(L.plan_w, tab_w);

  for (int k = threadIdx.x; k < H; k += blockDim.x) {
    double s, c;
    sincospi(-2.0 * (double)k / (double)H, &s, &c);
    tw_h[k] = make_float2((float)c, (float)s);
  }
  for (int k = threadIdx.x; k < W; k += blockDim.x) {
    double s, c;
    sincospi(-2.0 * (double)k / (double)W, &s, &c);
    tw_w[k] = make_float2((float)c, (float)s);
  }
  __syncthreads();

  float ms = 0.0f, mss = 0.0f;  // moments of everything this thread writes
  __shared__ double seg_sums[2 * SONAR_SPECTRAL_MAX_SEGMENTS];  // per-segment statistics, see spectral_batched_kernel
  if (threadIdx.x < 2 * SONAR_SPECTRAL_MAX_SEGMENTS) seg_sums[threadIdx.x] = 0.0;
  for (int64_t plane = blockIdx.x; plane < p.planes; plane += gridDim.x) {
    // ---------------- phase A: fill S ----------------
    if (p.in_real != nullptr) {
      const float* src = p.in_real + plane * (int64_t)H * W;
      for (int y = 2 * warp; y < H; y += 2 * kFftWarps) {
        const bool pair = (y + 1) < H;
        for (int x = lane; x < W; x += 32)
          bufA[x] = make_float2(src[(int64_t)y * W + x], pair ? src[(int64_t)(y + 1) * W + x] : 0.0f);
        __syncwarp();
        const float2* Z = warp_fft<false>(bufA, bufB, L.plan_w, tw_w, tab_w, lane);
        // X1[k] = (Z[k] + conj(Z[-k])) / 2 ; X2[k] = (Z[k] - conj(Z[-k])) / (2i)
        for (int k = lane; k < Wh; k += 32) {
          const float2 zk = Z[k];
          const float2 zm = cconj(Z[k == 0 ? 0 : W - k]);
          S[(size_t)y * WhP + k] = make_float2(0.5f * (zk.x + zm.x), 0.5f * (zk.y + zm.y));
          if (pair) {
            const float2 d = csub(zk, zm);  // (d) / (2i) = (d.y - i d.x) / 2
            S[(size_t)(y + 1) * WhP + k] = make_float2(0.5f * d.y, -0.5f * d.x);
          }
        }
        __syncwarp();
      }
    } else {
      const float2* src = reinterpret_cast<const float2*>(p.in_spec) + plane * (int64_t)H * Wh;
      for (int i = threadIdx.x; i < H * Wh; i += blockDim.x) {
        const int y = i / Wh, k = i - y * Wh;
        S[(size_t)y * WhP + k] = src[i];
      }
    }
    __syncthreads();

    // ---------------- phase B: columns ([forward FFT] -> gain -> inverse FFT) ----------------
    for (int k = warp; k < Wh; k += kFftWarps) {
      for (int y = lane; y < H; y += 32) bufA[y] = S[(size_t)y * WhP + k];
      __syncwarp();
      float2* cur = bufA;
      float2* other = bufB;
      if (p.in_real != nullptr) {
        cur = warp_fft<false>(bufA, bufB, L.plan_h, tw_h, tab_h, lane);
        other = cur == bufA ? bufB : bufA;
      }
      if (p.mask != nullptr) {
        for (int y = lane; y < H; y += 32) {
          const float g = p.mask[(size_t)y * Wh + k];
          cur[y].x *= g;
          cur[y].y *= g;
        }
        __syncwarp();
      }
      const float2* res = warp_fft<true>(cur, other, L.plan_h, tw_h, tab_h, lane);
      for (int y = lane; y < H; y += 32) S[(size_t)y * WhP + k] = res[y];
      __syncwarp();
    }
    __syncthreads();

    // ---------------- phase C: rows, Hermitian c2r, two rows per complex FFT ----------------
    float* dst = p.out + plane * (int64_t)H * W;
    const int half = W / 2;
    const bool even = (W & 1) == 0;
    for (int y = 2 * warp; y < H; y += 2 * kFftWarps) {
      const bool pair = (y + 1) < H;
      const float2* r1 = S + (size_t)y * WhP;
      const float2* r2 = S + (size_t)(y + 1) * WhP;
      // Z = Z1 + i*Z2 with Z1, Z2 the Hermitian extensions of the two half rows
      for (int k = lane; k < W; k += 32) {
        float2 z1, z2 = make_float2(0.f, 0.f);
        if (k <= half) {
          z1 = r1[k];
          if (pair) z2 = r2[k];
          if (k == 0 || (even && k == half)) {
            z1.y = 0.0f;
            z2.y = 0.0f;
          }
        } else {
          z1 = cconj(r1[W - k]);
          if (pair) z2 = cconj(r2[W - k]);
        }
        bufA[k] = make_float2(z1.x - z2.y, z1.y + z2.x);
      }
      __syncwarp();
      const float2* z = warp_fft<true>(bufA, bufB, L.plan_w, tw_w, tab_w, lane);
      for (int x = lane; x < W; x += 32) {
        const float2 v = z[x];
        const float o0 = v.x * p.out_scale;
        dst[(int64_t)y * W + x] = o0;
        ms += o0;
        mss += o0 * o0;
        if (pair) {
          const float o1 = v.y * p.out_scale;
          dst[(int64_t)(y + 1) * W + x] = o1;
          ms += o1;
          mss += o1 * o1;
        }
      }
      __syncwarp();
    }
    if (p.sums != nullptr) {
      const float ws = warp_sum(ms), wss = warp_sum(mss);
      if (lane == 0) {
        const int seg = p.sums_segment_planes > 0 ? (int)(plane / p.sums_segment_planes) : 0;
        atomicAdd(&seg_sums[2 * seg], (double)ws);
        atomicAdd(&seg_sums[2 * seg + 1], (double)wss);
      }
      ms = mss = 0.0f;
    }
    __syncthreads();
  }
  if (p.sums != nullptr) {
    if (threadIdx.x < 2 * SONAR_SPECTRAL_MAX_SEGMENTS && seg_sums[threadIdx.x] != 0.0) atomicAdd(&p.sums[threadIdx.x], seg_sums[threadIdx.x]);
    if (p.sums_clear != nullptr && blockIdx.x == 0 && threadIdx.x < 2) p.sums_clear[threadIdx.x] = 0.0;
  }
}

// =============================================================================================
// Batched in-place variant for the hot case (even W, lengths that factor into radices 2/3/4/5/8/9/10/16).
//  * Every FFT stage is ONE pass of the CTA over a group of planes, one register-resident radix-R butterfly
//    per work item (butterfly j, transform b), b fastest: adjacent lanes run the same butterfly of adjacent
//    transforms (columns: adjacent column slots, contiguous in shared memory; rows: adjacent rows,
//    conflict-free thanks to the odd pitch) whatever the batch size. Small planes (UNet activations,
//    32x32) are processed several per CTA.
//  * Large radices keep the pass count low: 90 = 10 x 9 and 80 = 10 x 8 are two passes each (radices 8..16
//    are Cooley-Tukey compositions of the 2/3/4/5 butterflies inside registers, inner twiddles compile-time
//    constants); the block-length-R stage of every axis has no twiddle multiplications.
//  * The transforms are IN PLACE (decimation in time: digit-reversed slots in, natural order out;
//    decimation in frequency: natural in, digit-reversed out): each butterfly owns its R slots, a plane
//    needs ONE shared-memory copy (58 KB at 90x160, 67 KB at 128x128), so 2-4 CTAs share an SM and one
//    CTA's barrier / global-load phases overlap the others' arithmetic. The ping-pong (Stockham) form of
//    this kernel fit one CTA per SM and sat at 49 % issue-slot utilisation with no eligible warp half of
//    the time (ncu r01h); on 64x64 planes, where both forms fit, 4 CTAs of 256 threads ran 1.5x faster
//    than 1 CTA of 1024.
//  * Digit reversal costs nothing, and there is no separate load or store pass: the block-length-R stage at
//    the global-memory end of an axis does the permutation with its addresses. Spectrum input: the first
//    inverse column stage (DIT) gathers whole rows of the global half spectrum into digit-reversed row slots
//    (coalesced along k, gain applied on the fly) and the column transform ends in natural order; the
//    Hermitian fold runs in natural order; the row transform is DIF and its last stage, run
//    butterfly-fastest so that adjacent lanes touch adjacent float2 of the row, scales and stores straight
//    to the global plane in natural order while reducing the output moments.
// Rows use the half-length c2r trick: for Hermitian X of even length W, with M = W/2,
//   Z[k] = (X[k] + conj X[M-k]) + i (X[k] - conj X[M-k]) e^{+2 pi i k / W},   k = 0..M-1
//   z = sum_k Z[k] e^{+2 pi i k n / M}  ==>  x[2n] = Re z[n], x[2n+1] = Im z[n]
// which also reproduces irfft2's treatment of the reference's NON-Hermitian input (the imaginary
// parts of the k = 0 and k = M bins are dropped, nothing else of the upper half is read).
// Real input (rfft2 front end: PowerFilterNoiseItem, OneF / GreenTest, FreeU-Extreme ffilter) runs the
// same inverse-sign stages on conjugated data (FFT(z) = conj(IFFT(conj z))): row DIT whose first stage gathers
// the packed real rows from global memory -> r2c unfold -> column DIF -> gain (.) conj on the loads of the
// column DIT -> the fold and row DIF above. The plane still makes exactly one trip from and one trip to HBM.
// History: warp-per-transform kernel 237 issued instructions per output element at 90x160, radix 2..5
// batched-lane kernel 158 (ncu r01g), ping-pong large-radix kernel 92 (ncu r01h).
// =============================================================================================
constexpr int kBatchedThreads = 512;  // per CTA, at most; >= 2 CTAs per SM
constexpr int kBatchedMaxStages = 16;

struct AxisPlan {
  int n;
  int n_stages;
  int radix[kBatchedMaxStages];
  int m[kBatchedMaxStages];        // distance between a butterfly's slots: n / (radix[0] * ... * radix[f])
  int nb[kBatchedMaxStages];       // butterflies per transform in the stage (n / radix)
  int tab_off[kBatchedMaxStages];  // offset of the stage's butterfly table
  int tab_size;
};

struct SpectralBatchedLaunch {
  SonarSpectralParams p;
  AxisPlan col;     // length H
  AxisPlan row;     // length M = W / 2
  int wh;           // M + 1
  int pitch;        // odd, >= wh: row pitch of a plane in shared memory (complex elements)
  int plane_elems;  // H * pitch
  int group;        // planes per CTA pass
  int fuse_fold;    // 1: the c2r fold runs inside the first row stage (run_stage_fold) instead of as a pass of its own
  float scale;      // out_scale (x 1/2 for real input: the r2c unfold leaves a factor 2)
  unsigned magic_wh, magic_cols, magic_rows, magic_row_nb;  // ceil(2^32 / d): d = wh, group * wh, group * H, M / last row radix
  unsigned long long magic_T;  // floor(2^64 / T) + 1, T = threads of the emulated ATen launch (Philox input)
};

__device__ __forceinline__ float2 cmul_conj(float2 a, float2 w) {  // a * conj(w)
  return make_float2(a.x * w.x + a.y * w.y, a.y * w.x - a.x * w.y);
}

// ---- compile-time roots of unity for the composite radices ----
constexpr double kPiD = 3.14159265358979323846264338327950288;
constexpr double cx_cos(double x) {
  double term = 1.0, sum = 1.0;
  for (int n = 1; n < 24; ++n) {
    term *= -x * x / (double)((2 * n - 1) * (2 * n));
    sum += term;
  }
  return sum;
}
constexpr double cx_sin(double x) {
  double term = x, sum = x;
  for (int n = 1; n < 24; ++n) {
    term *= -x * x / (double)((2 * n) * (2 * n + 1));
    sum += term;
  }
  return sum;
}
template <int R, int K>
struct Root {  // e^{+2 pi i K / R}
  static constexpr int k = ((K % R) + R) % R;
  static constexpr double angle = 2.0 * kPiD * (double)(2 * k <= R ? k : k - R) / (double)R;
  static constexpr float c = (float)cx_cos(angle);
  static constexpr float s = (float)cx_sin(angle);
};
template <int R, int K>
__device__ __forceinline__ float2 mul_root(float2 a) {  // a * e^{+2 pi i K / R}
  constexpr int k = ((K % R) + R) % R;
  if constexpr (k == 0) {
    return a;
  } else if constexpr (4 * k == R) {
    return make_float2(-a.y, a.x);
  } else if constexpr (2 * k == R) {
    return make_float2(-a.x, -a.y);
  } else if constexpr (4 * k == 3 * R) {
    return make_float2(a.y, -a.x);
  } else {
    constexpr float c = Root<R, K>::c, s = Root<R, K>::s;
    return make_float2(a.x * c - a.y * s, a.x * s + a.y * c);
  }
}

template <int I, int N, typename F>
__device__ __forceinline__ void static_for(F&& f) {
  if constexpr (I < N) {
    f(std::integral_constant<int, I>{});
    static_for<I + 1, N>(f);
  }
}

// Prime-ish radix inverse butterfly on registers, natural order in and out.
template <int R>
__device__ __forceinline__ void butterfly_inverse(float2 (&v)[R]) {
  if constexpr (R == 2) {
    const float2 a = v[0], b = v[1];
    v[0] = cadd(a, b);
    v[1] = csub(a, b);
  } else if constexpr (R == 4) {
    const float2 a0 = cadd(v[0], v[2]), a1 = csub(v[0], v[2]), a2 = cadd(v[1], v[3]);
    const float2 d = csub(v[1], v[3]);
    const float2 a3 = make_float2(-d.y, d.x);  // * (+i)
    v[0] = cadd(a0, a2);
    v[1] = cadd(a1, a3);
    v[2] = csub(a0, a2);
    v[3] = csub(a1, a3);
  } else if constexpr (R == 3) {
    const float2 t1 = cadd(v[1], v[2]);
    const float2 t2 = make_float2(v[0].x - 0.5f * t1.x, v[0].y - 0.5f * t1.y);
    const float2 d = csub(v[1], v[2]);
    const float2 t3 = make_float2(-0.86602540378443865f * d.y, 0.86602540378443865f * d.x);
    v[0] = cadd(v[0], t1);
    v[1] = cadd(t2, t3);
    v[2] = csub(t2, t3);
  } else {
    static_assert(R == 5, "butterfly radix");
    constexpr float c1 = 0.30901699437494742f, c2 = -0.80901699437494742f;
    constexpr float s1 = 0.95105651629515357f, s2 = 0.58778525229247313f;
    const float2 v0 = v[0];
    const float2 a1 = cadd(v[1], v[4]), a2 = cadd(v[2], v[3]), b1 = csub(v[1], v[4]), b2 = csub(v[2], v[3]);
    const float2 m1 = make_float2(v0.x + c1 * a1.x + c2 * a2.x, v0.y + c1 * a1.y + c2 * a2.y);
    const float2 m2 = make_float2(v0.x + c2 * a1.x + c1 * a2.x, v0.y + c2 * a1.y + c1 * a2.y);
    const float2 e1 = make_float2(s1 * b1.x + s2 * b2.x, s1 * b1.y + s2 * b2.y);
    const float2 e2 = make_float2(s2 * b1.x - s1 * b2.x, s2 * b1.y - s1 * b2.y);
    const float2 n1 = make_float2(-e1.y, e1.x), n2 = make_float2(-e2.y, e2.x);  // * (+i)
    v[0] = make_float2(v0.x + a1.x + a2.x, v0.y + a1.y + a2.y);
    v[1] = cadd(m1, n1);
    v[2] = cadd(m2, n2);
    v[3] = csub(m2, n2);
    v[4] = csub(m1, n1);
  }
}

// Composite radices R = R1 * R2 (Cooley-Tukey inside registers); R1 == 1: a plain butterfly.
template <int R>
struct RadixSplit {
  static constexpr int r1 = R == 6 ? 2 : R == 8 ? 2 : R == 9 ? 3 : R == 10 ? 2 : R == 16 ? 4 : 1;
  static constexpr int r2 = R / r1;
};
// Output index held by register slot `slot` after dft_inverse<R> (the digit reversal is applied by the
// store addresses, not by moving registers).
template <int R>
__device__ __forceinline__ constexpr int dft_out_index(int slot) {
  return RadixSplit<R>::r1 == 1 ? slot : (slot / RadixSplit<R>::r2) + RadixSplit<R>::r1 * (slot % RadixSplit<R>::r2);
}
// X[u1 + R1 u2] = sum_t2 W_R2^{t2 u2} W_R^{t2 u1} sum_t1 W_R1^{t1 u1} v[t1 R2 + t2], W_n = e^{+2 pi i / n};
// slot u1 R2 + u2 receives X[u1 + R1 u2].
template <int R>
__device__ __forceinline__ void dft_inverse(float2 (&v)[R]) {
  constexpr int R1 = RadixSplit<R>::r1, R2 = RadixSplit<R>::r2;
  if constexpr (R1 == 1) {
    butterfly_inverse<R>(v);
  } else {
    static_for<0, R2>([&](auto t2c) {
      constexpr int t2 = decltype(t2c)::value;
      float2 a[R1];
#pragma unroll
      for (int t1 = 0; t1 < R1; ++t1) a[t1] = v[t1 * R2 + t2];
      butterfly_inverse<R1>(a);
      static_for<0, R1>([&](auto u1c) {
        constexpr int u1 = decltype(u1c)::value;
        v[u1 * R2 + decltype(t2c)::value] = mul_root<R, u1 * decltype(t2c)::value>(a[u1]);
      });
    });
#pragma unroll
    for (int u1 = 0; u1 < R1; ++u1) {
      float2 b[R2];
#pragma unroll
      for (int t2 = 0; t2 < R2; ++t2) b[t2] = v[u1 * R2 + t2];
      butterfly_inverse<R2>(b);
#pragma unroll
      for (int u2 = 0; u2 < R2; ++u2) v[u1 * R2 + u2] = b[u2];
    }
  }
}

// Where the first executed stage of an axis gets its inputs from.
enum StageSource : int {
  SRC_PLAIN = 0,      // the butterfly's own slots
  SRC_CONJ_IN = 1,    // own slots, conjugated
  SRC_SPECTRUM = 2,   // first inverse column stage, spectrum input: global half spectrum (.) gain mask
  SRC_CONJ_MASK = 3,  // first inverse column stage, real input: conj(forward result) (.) gain mask
  SRC_REAL_ROWS = 4,  // first forward row stage: packed pairs (x[2n], -x[2n+1]) gathered from the global real plane
  SRC_PHILOX = 5      // first inverse column stage: torch.randn(complex64) regenerated from the Philox stream (.) gain mask
};
enum StageSink : int {
  SINK_SLOTS = 0,  // back into the butterfly's own slots
  SINK_GLOBAL = 1  // last inverse row stage: x[y][2n], x[y][2n+1] = z[n] * scale straight to the global plane
};
enum StageMode : int { MODE_DIT = 0, MODE_DIF = 1 };

struct StageArgs {
  float2* buf;                  // the CTA's planes, [plane][row][pitch]
  const float2* spec;           // SRC_SPECTRUM: the group's first global plane
  const float2* tw;             // e^{-2 pi i k / n} of the axis
  const ushort2* tab;           // butterfly table of the stage: .x first slot, .y twiddle step
  const float* mask;            // (H, wh) gain or nullptr
  const unsigned short* pos_m;  // row axis: index n -> slot after a DIF transform (= slot a DIT transform reads it from)
  const unsigned short* idx_h;  // column axis: slot -> index ky
  const float2* real_rows;      // SRC_REAL_ROWS: the group's first global plane as (W/2) float2 per row
  float2* out_rows;             // SINK_GLOBAL: the group's first output plane as (W/2) float2 per row
  float scale;                  // SINK_GLOBAL
  float ms, mss;                // SINK_GLOBAL: moments of what this thread stored
  unsigned magic_nb;            // ceil(2^32 / nb) of the block-length-R row stage
  int nb, m;                    // butterflies per transform, distance between a butterfly's slots
  int nbatch;                   // transforms in this pass (valid planes x columns or rows)
  unsigned magic_batch;         // ceil(2^32 / nbatch)
  unsigned magic_wh;            // ceil(2^32 / wh)
  int group;                    // planes per pass the kernel was launched with (1: no plane split of b)
  int wh, pitch, plane_elems, M;
  int gstride;                  // row stride of the global spectrum / mask (wh; the cluster kernel's wh is its share of the columns)
  int spec_plane_elems;         // H * wh (global spectrum plane)
  // SRC_PHILOX: float index (in the global draw) of the group's first plane; the draw itself is read from the
  // launch block in the constant bank (no registers in the kernels that do not use it)
  int64_t philox_first;
  const SpectralBatchedLaunch* launch;
};

// One in-place inverse-sign stage over a batch of transforms: every (butterfly j, transform b) item owns
// its R slots, so a thread may run any number of items and the CTA needs one buffer only.
//   DIT (digit-reversed in -> natural out, stages last..first): twiddle the inputs, butterfly, store;
//   DIF (natural in -> digit-reversed out, stages first..last): butterfly, twiddle the outputs, store.
// NOTW: the stage with block length R (all twiddles are 1). Slot s of transform b lives at base(b) + s *
// stride: columns base = plane * plane_elems + column slot, stride = pitch; rows base = b * pitch, stride 1.
template <int MODE, int SOURCE, int SINK, int R, bool COLS, bool NOTW>
__device__ __forceinline__ void run_stage(StageArgs& a) {
  // The block-length-R row stage that talks to global memory runs its items butterfly-fastest: butterfly
  // `lo` of a row owns indices lo + t * nb, so adjacent lanes touch adjacent float2 of the global row.
  constexpr bool GLOBAL_ROWS = SOURCE == SRC_REAL_ROWS || SINK == SINK_GLOBAL;
  static_assert(!GLOBAL_ROWS || (!COLS && NOTW), "global row stages are the block-length-R row stage");
  const int total = a.nb * a.nbatch;
  const int stride = COLS ? a.pitch : 1;
  const int estep = a.m * stride;
  for (int w = threadIdx.x; w < total; w += blockDim.x) {
    int j, b, g = 0, c = 0, base, first_slot, tstep = 0;
    if (GLOBAL_ROWS) {
      b = a.nb == 1 ? w : (int)__umulhi((unsigned)w, a.magic_nb);  // (ceil(2^32 / 1) does not fit 32 bits)
      j = w - b * a.nb;                 // lowest index of the butterfly
      first_slot = (int)a.pos_m[j];     // its slots are first_slot .. first_slot + R - 1
      base = b * a.pitch;
    } else {
      j = (int)__umulhi((unsigned)w, a.magic_batch);
      b = w - j * a.nbatch;
      c = b;
      if (COLS) {
        if (a.group > 1) {
          g = (int)__umulhi((unsigned)b, a.magic_wh);
          c = b - g * a.wh;
        }
        base = g * a.plane_elems + c;
      } else {
        base = b * a.pitch;
      }
      const ushort2 e = a.tab[j];
      first_slot = e.x;
      tstep = e.y;
    }
    float2* p = a.buf + base + first_slot * stride;
    float2 v[R];
    if (SOURCE == SRC_SPECTRUM) {  // block length R: the slots are rows first_slot .. first_slot + R - 1
      const float2* gp = a.spec + ((int64_t)g * a.spec_plane_elems + c);
      const float* mp = a.mask + c;
      const bool has_mask = a.mask != nullptr;
#pragma unroll
      for (int t = 0; t < R; ++t) {
        const int off = (int)a.idx_h[first_slot + t] * a.gstride;
        v[t] = __ldg(gp + off);
        if (has_mask) {
          const float gain = __ldg(mp + off);
          v[t].x *= gain;
          v[t].y *= gain;
        }
      }
    } else if (SOURCE == SRC_PHILOX) {
      // element (g, row, c) of the local tensor is complex number e = (g * H + row) * wh + c, floats 2e and 2e + 1 of
      // the draw; float li is lane (li / T) % 4 of call (li / T) / 4 of Philox subsequence li % T. T is even, so
      // both floats of a complex number share the call index and the lane.
      const float* mp = a.mask + c;
      const bool has_mask = a.mask != nullptr;
      const int64_t e0 = a.philox_first + 2 * ((int64_t)g * a.spec_plane_elems + c);
      const SonarSpectralParams& pp = a.launch->p;
      const PhiloxStream st{pp.philox_seed, pp.philox_offset, pp.philox_grid_blocks * (uint32_t)kBlock};
      const float std = pp.philox_std;
#pragma unroll
      for (int t = 0; t < R; ++t) {
        const int off = (int)a.idx_h[first_slot + t] * a.gstride;
        const uint64_t li = (uint64_t)(e0 + 2 * off);
        const uint64_t q = __umul64hi(li, a.launch->magic_T);
        const uint32_t vt = (uint32_t)(li - q * st.threads);
        const int lane = (int)(q & 3u);
        const uint4 r0 = philox_raw(st, vt, q >> 2), r1 = philox_raw(st, vt + 1u, q >> 2);
        v[t] = make_float2(philox_normal_lane(r0, lane) * std, philox_normal_lane(r1, lane) * std);
        if (has_mask) {
          const float gain = __ldg(mp + off);
          v[t].x *= gain;
          v[t].y *= gain;
        }
      }
    } else if (SOURCE == SRC_REAL_ROWS) {
      const float2* gp = a.real_rows + ((int64_t)b * a.M + j);
#pragma unroll
      for (int t = 0; t < R; ++t) {
        v[t] = __ldg(gp + t * a.nb);
        v[t].y = -v[t].y;
      }
    } else {
#pragma unroll
      for (int t = 0; t < R; ++t) v[t] = p[t * estep];
      if (SOURCE == SRC_CONJ_IN) {
#pragma unroll
        for (int t = 0; t < R; ++t) v[t].y = -v[t].y;
      }
      if (SOURCE == SRC_CONJ_MASK) {
        if (a.mask != nullptr) {
          const float* mp = a.mask + c;
#pragma unroll
          for (int t = 0; t < R; ++t) {
            const float gain = __ldg(mp + (int)a.idx_h[first_slot + t] * a.gstride);
            v[t] = make_float2(v[t].x * gain, -v[t].y * gain);
          }
        } else {
#pragma unroll
          for (int t = 0; t < R; ++t) v[t].y = -v[t].y;
        }
      }
    }
    if (MODE == MODE_DIT && !NOTW) {
#pragma unroll
      for (int t = 1; t < R; ++t) v[t] = cmul_conj(v[t], a.tw[t * tstep]);
    }
    dft_inverse<R>(v);
    if (MODE == MODE_DIF && !NOTW) {
#pragma unroll
      for (int s = 0; s < R; ++s)
        if (dft_out_index<R>(s) != 0) v[s] = cmul_conj(v[s], a.tw[dft_out_index<R>(s) * tstep]);
    }
    if (SINK == SINK_GLOBAL) {
      float2* gp = a.out_rows + ((int64_t)b * a.M + j);
#pragma unroll
      for (int s = 0; s < R; ++s) {
        const float2 o = make_float2(v[s].x * a.scale, v[s].y * a.scale);
        gp[dft_out_index<R>(s) * a.nb] = o;
        a.ms += o.x + o.y;
        a.mss += o.x * o.x + o.y * o.y;
      }
    } else {
#pragma unroll
      for (int s = 0; s < R; ++s) p[dft_out_index<R>(s) * estep] = v[s];
    }
  }
}

template <int MODE, int SOURCE, int SINK, bool COLS, bool NOTW>
__device__ __forceinline__ void dispatch_stage(int R, StageArgs& a) {
  switch (R) {  // block-uniform
    case 16: run_stage<MODE, SOURCE, SINK, 16, COLS, NOTW>(a); break;
    case 10: run_stage<MODE, SOURCE, SINK, 10, COLS, NOTW>(a); break;
    case 9: run_stage<MODE, SOURCE, SINK, 9, COLS, NOTW>(a); break;
    case 8: run_stage<MODE, SOURCE, SINK, 8, COLS, NOTW>(a); break;
    case 5: run_stage<MODE, SOURCE, SINK, 5, COLS, NOTW>(a); break;
    case 4: run_stage<MODE, SOURCE, SINK, 4, COLS, NOTW>(a); break;
    case 3: run_stage<MODE, SOURCE, SINK, 3, COLS, NOTW>(a); break;
    default: run_stage<MODE, SOURCE, SINK, 2, COLS, NOTW>(a); break;
  }
}

// ---- c2r fold fused into the first row stage -------------------------------------------------------------------
// The fold turns the pair (A[k], A[M-k]) into (B[k], B[M-k]) (see pair_pass). The first DIF row stage (block length M,
// m = M / R) gives butterfly j the slots j + t m: their mirrors M - j - t m = (m - j) + (R - 1 - t) m are the slots of
// butterfly m - j. One item therefore takes BOTH butterflies j and m - j (2R values), folds them in registers and
// runs the two radix-R butterflies: the fold's own trip through shared memory (a load and a store per element, its
// index arithmetic and its barrier) disappears. j = 0 (mirror: itself plus the Nyquist slot M) and j = m / 2 (its own
// mirror) are single-butterfly items. Needs m even and R <= 8 (2R complex values in registers).
__device__ __forceinline__ void fold_pair(float2& ak, float2& am, const float2 w) {
  const float2 s = make_float2(ak.x + am.x, ak.y - am.y);
  const float2 d = make_float2(ak.x - am.x, ak.y + am.y);
  const float2 wd = cmul(d, w);
  ak = make_float2(s.x - wd.y, s.y + wd.x);
  am = make_float2(s.x + wd.y, wd.x - s.y);
}

// DIF butterfly of slots p[0], p[estep], ...: butterfly, twiddle the outputs, store (digit reversal by the addresses)
template <int R>
__device__ __forceinline__ void dif_finish(float2* p, float2 (&v)[R], int tstep, int estep, const float2* __restrict__ tw) {
  dft_inverse<R>(v);
#pragma unroll
  for (int s = 0; s < R; ++s)
    if (dft_out_index<R>(s) != 0) v[s] = cmul_conj(v[s], tw[dft_out_index<R>(s) * tstep]);
#pragma unroll
  for (int s = 0; s < R; ++s) p[dft_out_index<R>(s) * estep] = v[s];
}

template <int R>
__device__ __forceinline__ void run_stage_fold(StageArgs& a, const float2* __restrict__ tw_w) {
  const int m = a.m, M = a.M, half = m >> 1;
  const int total = (half + 1) * a.nbatch;
  for (int w = threadIdx.x; w < total; w += blockDim.x) {
    const int jj = (int)__umulhi((unsigned)w, a.magic_batch);
    const int b = w - jj * a.nbatch;
    float2* row = a.buf + b * a.pitch;
    if (jj == 0 || jj == half) {
      float2 v[R];
#pragma unroll
      for (int t = 0; t < R; ++t) v[t] = row[jj + t * m];
      if (jj == 0) {
        const float x0 = v[0].x, xm = row[M].x;  // the fold drops the imaginary parts of the DC and Nyquist bins
        v[0] = make_float2(x0 + xm, x0 - xm);
#pragma unroll
        for (int t = 1; 2 * t < R; ++t) fold_pair(v[t], v[R - t], tw_w[t * m]);
        if ((R & 1) == 0) {  // k = M / 2 pairs with itself
          float2 twin = v[R / 2];
          fold_pair(v[R / 2], twin, tw_w[(R / 2) * m]);
        }
      } else {
#pragma unroll
        for (int t = 0; 2 * t < R - 1; ++t) fold_pair(v[t], v[R - 1 - t], tw_w[jj + t * m]);
        if (R & 1) {
          float2 twin = v[(R - 1) / 2];
          fold_pair(v[(R - 1) / 2], twin, tw_w[jj + ((R - 1) / 2) * m]);
        }
      }
      dif_finish<R>(row + jj, v, jj, m, a.tw);
    } else {
      const int j2 = m - jj;
      float2 va[R], vb[R];
#pragma unroll
      for (int t = 0; t < R; ++t) {
        va[t] = row[jj + t * m];
        vb[t] = row[j2 + t * m];
      }
#pragma unroll
      for (int t = 0; t < R; ++t) fold_pair(va[t], vb[R - 1 - t], tw_w[jj + t * m]);
      dif_finish<R>(row + jj, va, jj, m, a.tw);
      dif_finish<R>(row + j2, vb, j2, m, a.tw);
    }
  }
}

__device__ __forceinline__ void dispatch_stage_fold(int R, StageArgs& a, const float2* __restrict__ tw_w) {
  switch (R) {  // block-uniform
    case 8: run_stage_fold<8>(a, tw_w); break;
    case 5: run_stage_fold<5>(a, tw_w); break;
    case 4: run_stage_fold<4>(a, tw_w); break;
    case 3: run_stage_fold<3>(a, tw_w); break;
    default: run_stage_fold<2>(a, tw_w); break;
  }
}

// Inverse-sign transform of one axis by decimation in time, stages last..first: the block-length-R stage
// reads FIRST_SOURCE (which supplies index k at slot pos[k]), the result is in natural order.
template <int FIRST_SOURCE, bool COLS>
__device__ __forceinline__ void run_axis_dit(const AxisPlan& plan, StageArgs& a, const ushort2* tab) {
  for (int f = plan.n_stages - 1; f >= 0; --f) {
    a.nb = plan.nb[f];
    a.m = plan.m[f];
    a.tab = tab + plan.tab_off[f];
    if (f == plan.n_stages - 1)
      dispatch_stage<MODE_DIT, FIRST_SOURCE, SINK_SLOTS, COLS, true>(plan.radix[f], a);
    else
      dispatch_stage<MODE_DIT, SRC_PLAIN, SINK_SLOTS, COLS, false>(plan.radix[f], a);
    __syncthreads();
  }
}

// Inverse-sign transform of one axis by decimation in frequency, stages first..last: natural order in,
// index k ends at slot pos[k] -- or, with LAST_SINK = SINK_GLOBAL, in natural order in the global plane.
template <int LAST_SINK, bool COLS>
__device__ __forceinline__ void run_axis_dif(const AxisPlan& plan, StageArgs& a, const ushort2* tab, int first = 0) {
  const int last = plan.n_stages - 1;
  for (int f = first; f <= last; ++f) {
    a.nb = plan.nb[f];
    a.m = plan.m[f];
    a.tab = tab + plan.tab_off[f];
    if (f == last)
      dispatch_stage<MODE_DIF, SRC_PLAIN, LAST_SINK, COLS, true>(plan.radix[f], a);
    else
      dispatch_stage<MODE_DIF, SRC_PLAIN, SINK_SLOTS, COLS, false>(plan.radix[f], a);
    __syncthreads();
  }
}

// Hermitian pair pass over every row, in place and in natural order (column k at slot k, Nyquist at M):
// each item turns the pair (A[k], A[M-k]) into (B[k], B[M-k]),
//   B[k] = (A[k] + conj A[M-k]) + i e^{+2 pi i k / W} (A[k] - conj A[M-k]),
// which is both the c2r fold (A = X, B = Z feeding the half-length inverse row transform) and the r2c
// unfold (A = Y = conj FFT_M(packed row), B = 2 conj X). Only k = 0 differs: the fold drops the imaginary
// parts of the DC and Nyquist bins and frees slot M, the unfold fills slot M.
template <bool UNFOLD>
__device__ __forceinline__ void pair_pass(float2* buf, int rows_total, unsigned magic_rows, int M, int pitch,
                                          const float2* __restrict__ tw_w) {
  const int total = ((M >> 1) + 1) * rows_total;
  for (int w = threadIdx.x; w < total; w += blockDim.x) {
    const int k = (int)__umulhi((unsigned)w, magic_rows);
    float2* row = buf + (w - k * rows_total) * pitch;
    if (k == 0) {
      if (UNFOLD) {
        const float2 y0 = row[0];
        row[0] = make_float2(2.0f * (y0.x - y0.y), 0.0f);
        row[M] = make_float2(2.0f * (y0.x + y0.y), 0.0f);
      } else {
        const float x0 = row[0].x, xm = row[M].x;
        row[0] = make_float2(x0 + xm, x0 - xm);
      }
      continue;
    }
    const float2 ak = row[k], am = row[M - k];
    const float2 s = make_float2(ak.x + am.x, ak.y - am.y);
    const float2 d = make_float2(ak.x - am.x, ak.y + am.y);
    const float2 wd = cmul(d, tw_w[k]);
    row[k] = make_float2(s.x - wd.y, s.y + wd.x);
    if (2 * k != M) row[M - k] = make_float2(s.x + wd.y, wd.x - s.y);
  }
}

// Per-CTA tables of one axis. tab: for butterfly j of stage f (block length L = n / (R_0 .. R_{f-1}),
// sub-length m = L / R_f): .x = first slot (j / m) * L + (j % m), .y = twiddle step (j % m) * (n / L).
// pos[k] = slot of index k after a DIF transform (= slot a DIT transform wants it in), idx = its inverse.
__device__ void build_axis_tables(const AxisPlan& plan, ushort2* __restrict__ tab, unsigned short* __restrict__ pos,
                                  unsigned short* __restrict__ idx) {
  for (int f = 0; f < plan.n_stages; ++f) {
    const int R = plan.radix[f], m = plan.m[f], L = m * R, nb = plan.n / R, unit = plan.n / L;
    for (int j = threadIdx.x; j < nb; j += blockDim.x) {
      const int i = j % m;
      tab[plan.tab_off[f] + j] = make_ushort2((unsigned short)((j / m) * L + i), (unsigned short)(i * unit));
    }
  }
  for (int k = threadIdx.x; k < plan.n; k += blockDim.x) {
    int rem = k, slot = 0;
    for (int f = 0; f < plan.n_stages; ++f) {
      slot += (rem % plan.radix[f]) * plan.m[f];
      rem /= plan.radix[f];
    }
    if (pos != nullptr) pos[k] = (unsigned short)slot;
    if (idx != nullptr) idx[slot] = (unsigned short)k;
  }
}

__host__ __device__ __forceinline__ unsigned magic_of(int d) { return (unsigned)((0x100000000ull + (unsigned)d - 1u) / (unsigned)d); }

enum SpectralInput : int { INPUT_SPECTRUM = 0, INPUT_REAL = 1, INPUT_PHILOX = 2 };

template <int INPUT>
__device__ __forceinline__ void spectral_batched_body(const SpectralBatchedLaunch& L) {
  constexpr bool REAL = INPUT == INPUT_REAL;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const SonarSpectralParams& p = L.p;
  const int H = p.H, W = p.W, M = W >> 1, Wh = L.wh, P = L.pitch, G = L.group;
  float2* tw_h = reinterpret_cast<float2*>(smem_raw);  // e^{-2 pi i k / H}
  float2* tw_m = tw_h + H;                             // e^{-2 pi i k / M}
  float2* tw_w = tw_m + M;                             // e^{+2 pi i k / W}, k < M
  float2* buf = tw_w + M;
  ushort2* tab_col = reinterpret_cast<ushort2*>(buf + (size_t)G * L.plane_elems);
  ushort2* tab_row = tab_col + L.col.tab_size;
  unsigned short* pos_m = reinterpret_cast<unsigned short*>(tab_row + L.row.tab_size);
  unsigned short* idx_h = pos_m + M;
  build_axis_tables(L.col, tab_col, nullptr, idx_h);
  build_axis_tables(L.row, tab_row, pos_m, nullptr);
  for (int k = threadIdx.x; k < H; k += blockDim.x) {
    double sn, cs;
    sincospi(-2.0 * (double)k / (double)H, &sn, &cs);
    tw_h[k] = make_float2((float)cs, (float)sn);
  }
  for (int k = threadIdx.x; k < M; k += blockDim.x) {
    double sn, cs;
    sincospi(-2.0 * (double)k / (double)M, &sn, &cs);
    tw_m[k] = make_float2((float)cs, (float)sn);
    sincospi(2.0 * (double)k / (double)W, &sn, &cs);
    tw_w[k] = make_float2((float)cs, (float)sn);
  }
  __syncthreads();
  StageArgs a;
  a.buf = buf;
  a.mask = p.mask;
  a.pos_m = pos_m;
  a.idx_h = idx_h;
  a.group = G;
  a.wh = Wh;
  a.gstride = Wh;
  a.pitch = P;
  a.plane_elems = L.plane_elems;
  a.M = M;
  a.spec_plane_elems = H * Wh;
  a.magic_wh = L.magic_wh;
  a.magic_nb = L.magic_row_nb;
  a.scale = L.scale;
  a.ms = 0.0f;
  a.mss = 0.0f;
  a.launch = &L;
  // Statistics of what the CTA stores: after every group each warp adds its partials to a shared fp64 pair of the
  // group's segment (a run of sums_segment_planes planes = one noise sample; the planner lets no group straddle two),
  // and the CTA flushes the pairs it touched with one global atomic each at the end.
  __shared__ double seg_sums[2 * SONAR_SPECTRAL_MAX_SEGMENTS];
  if (threadIdx.x < 2 * SONAR_SPECTRAL_MAX_SEGMENTS) seg_sums[threadIdx.x] = 0.0;  // (ordered by the barriers of the first stage)
  int magic_for = G;
  unsigned magic_cols = L.magic_cols, magic_rows = L.magic_rows;
  for (int64_t plane0 = (int64_t)blockIdx.x * G; plane0 < p.planes; plane0 += (int64_t)gridDim.x * G) {
    const int valid = (int)(p.planes - plane0 < G ? p.planes - plane0 : G);
    if (valid != magic_for) {  // ragged last group only
      magic_cols = magic_of(valid * Wh);
      magic_rows = magic_of(valid * H);
      magic_for = valid;
    }
    const int rows_total = valid * H;
    if constexpr (REAL) {
      // forward rows (half length, conj trick; the block-length-R stage gathers the packed real rows from
      // global memory), r2c unfold, forward columns: conj rfft2 with ky at slot pos_h[ky]
      a.real_rows = reinterpret_cast<const float2*>(p.in_real + plane0 * (int64_t)H * W);
      a.tw = tw_m;
      a.nbatch = rows_total;
      a.magic_batch = magic_rows;
      run_axis_dit<SRC_REAL_ROWS, false>(L.row, a, tab_row);
      pair_pass<true>(buf, rows_total, magic_rows, M, P, tw_w);
      __syncthreads();
      a.tw = tw_h;
      a.nbatch = valid * Wh;
      a.magic_batch = magic_cols;
      run_axis_dif<SINK_SLOTS, true>(L.col, a, tab_col);
      // gain, then inverse columns back to natural row order
      run_axis_dit<SRC_CONJ_MASK, true>(L.col, a, tab_col);
    } else if constexpr (INPUT == INPUT_PHILOX) {
      // inverse columns whose block-length-R stage draws the half spectrum from the Philox stream in registers
      a.philox_first = p.philox_begin + 2 * plane0 * (int64_t)H * Wh;
      a.tw = tw_h;
      a.nbatch = valid * Wh;
      a.magic_batch = magic_cols;
      run_axis_dit<SRC_PHILOX, true>(L.col, a, tab_col);
    } else {
      // inverse columns: the block-length-R stage gathers the rows of the global half spectrum into
      // digit-reversed row slots (coalesced along k) and applies the gain on the fly
      a.spec = reinterpret_cast<const float2*>(p.in_spec) + plane0 * (int64_t)H * Wh;
      a.tw = tw_h;
      a.nbatch = valid * Wh;
      a.magic_batch = magic_cols;
      run_axis_dit<SRC_SPECTRUM, true>(L.col, a, tab_col);
    }
    // c2r fold, then inverse rows of length M; the last stage scales and stores x[y][2n], x[y][2n+1] =
    // z[y][n] to the global plane (its trailing barrier also protects the buffer from the next group)
    a.out_rows = reinterpret_cast<float2*>(p.out + plane0 * (int64_t)H * W);
    a.tw = tw_m;
    a.nbatch = rows_total;
    a.magic_batch = magic_rows;
    if (L.fuse_fold) {
      a.nb = L.row.nb[0];
      a.m = L.row.m[0];
      dispatch_stage_fold(L.row.radix[0], a, tw_w);
      __syncthreads();
      run_axis_dif<SINK_GLOBAL, false>(L.row, a, tab_row, 1);
    } else {
      pair_pass<false>(buf, rows_total, magic_rows, M, P, tw_w);
      __syncthreads();
      run_axis_dif<SINK_GLOBAL, false>(L.row, a, tab_row);
    }
    if (p.sums != nullptr) {
      const float ws = warp_sum(a.ms), wss = warp_sum(a.mss);
      if ((threadIdx.x & 31) == 0) {
        const int seg = p.sums_segment_planes > 0 ? (int)(plane0 / p.sums_segment_planes) : 0;
        atomicAdd(&seg_sums[2 * seg], (double)ws);
        atomicAdd(&seg_sums[2 * seg + 1], (double)wss);
      }
      a.ms = a.mss = 0.0f;
    }
  }
  if (p.sums != nullptr) {
    __syncthreads();
    if (threadIdx.x < 2 * SONAR_SPECTRAL_MAX_SEGMENTS && seg_sums[threadIdx.x] != 0.0) atomicAdd(&p.sums[threadIdx.x], seg_sums[threadIdx.x]);
    if (p.sums_clear != nullptr && blockIdx.x == 0 && threadIdx.x < 2) p.sums_clear[threadIdx.x] = 0.0;
  }
}

template <int INPUT>
__global__ void __launch_bounds__(kBatchedThreads, 2)
spectral_batched_kernel(const __grid_constant__ SpectralBatchedLaunch L) {
  spectral_batched_body<INPUT>(L);
}

// Fewest stages over the radix set, ties broken by the smaller radix sum (8 x 8 before 16 x 4).
static void search_plan(int rem, int depth, int sum, int* cur, int* best, int* best_depth, int* best_sum) {
  static const int kRadices[] = {16, 10, 9, 8, 5, 4, 3, 2};
  if (rem == 1) {
    if (depth < *best_depth || (depth == *best_depth && sum < *best_sum)) {
      *best_depth = depth;
      *best_sum = sum;
      for (int i = 0; i < depth; ++i) best[i] = cur[i];
    }
    return;
  }
  if (depth + 1 > *best_depth || depth >= kBatchedMaxStages) return;
  for (int r : kRadices) {
    if (rem % r != 0 || (depth > 0 && r > cur[depth - 1])) continue;  // non-increasing: each multiset once
    cur[depth] = r;
    search_plan(rem / r, depth + 1, sum + r, cur, best, best_depth, best_sum);
  }
}

// `ascending` puts the largest radix last. The row axis wants that: its block-length-R stage runs
// butterfly-fastest against global memory, where butterfly j owns slots j * R .. j * R + R - 1, and a large
// (ideally not power-of-two) R spreads adjacent lanes over the shared-memory banks (90x160: 37.9 -> 35.8 us).
static bool make_axis_plan(int n, AxisPlan* plan, bool ascending = false) {
  plan->n = n;
  plan->n_stages = 0;
  plan->tab_size = 0;
  if (n < 2 || n > 65535) return false;
  int cur[kBatchedMaxStages], best[kBatchedMaxStages], best_depth = kBatchedMaxStages + 1, best_sum = 1 << 30;
  search_plan(n, 0, 0, cur, best, &best_depth, &best_sum);
  if (best_depth > kBatchedMaxStages) return false;  // another prime factor: the generic kernel handles it
  int block = n;
  for (int f = 0; f < best_depth; ++f) {
    const int r = ascending ? best[best_depth - 1 - f] : best[f];
    plan->radix[f] = r;
    plan->m[f] = block / r;
    plan->nb[f] = n / r;
    plan->tab_off[f] = plan->tab_size;
    plan->tab_size += n / r;
    block /= r;
  }
  plan->n_stages = best_depth;
  return true;
}

static size_t batched_smem_bytes(int H, int W, int group, const AxisPlan& col, const AxisPlan& row) {
  const int M = W / 2, wh = M + 1, pitch = wh | 1;
  size_t bytes = ((size_t)H + 2 * (size_t)M + (size_t)group * H * pitch) * sizeof(float2) +
                 (size_t)(col.tab_size + row.tab_size) * sizeof(ushort2) + (size_t)(M + H) * sizeof(unsigned short);
  return (bytes + 15) & ~(size_t)15;
}

// Host-side planning of the batched kernel (no CUDA calls besides the cached device attributes): fills the
// launch block and the launch geometry; false = not applicable (the generic kernel handles the call).
static bool plan_spectral_batched(const SonarSpectralParams& p, SpectralBatchedLaunch* out, int* threads_out,
                                  int* ctas_per_sm_out, int64_t* grid_out, size_t* smem_out) {
  if ((p.W & 1) || p.W < 4 || p.H < 2) return false;
  if ((reinterpret_cast<uintptr_t>(p.out) & 7u) != 0) return false;
  if (p.in_real != nullptr && (reinterpret_cast<uintptr_t>(p.in_real) & 7u) != 0) return false;
  SpectralBatchedLaunch& L = *out;
  L.p = p;
  if (!make_axis_plan(p.H, &L.col) || !make_axis_plan(p.W / 2, &L.row, /*ascending=*/true)) return false;
  const int M = p.W / 2;
  L.wh = M + 1;
  L.pitch = L.wh | 1;
  L.plane_elems = p.H * L.pitch;
  L.scale = p.in_real != nullptr ? 0.5f * p.out_scale : p.out_scale;
  const DeviceInfo& di = device_info();
  const size_t budget = (size_t)di.max_smem_optin;
  if (batched_smem_bytes(p.H, p.W, 1, L.col, L.row) > budget) return false;
  auto ctas_that_fit = [&](int group) {  // by shared memory (1 KB per CTA is reserved by the runtime), at most 4
    const size_t need = batched_smem_bytes(p.H, p.W, group, L.col, L.row) + 1024;
    const size_t n = (budget + 1024) / need;
    return (int)(n > 4 ? 4 : n);
  };
  // planes per pass: tiny planes are grouped until the leanest stage has ~256 butterflies, without starving SMs
  int64_t items_min = INT64_MAX;
  for (int f = 0; f < L.col.n_stages; ++f) items_min = std::min<int64_t>(items_min, (int64_t)L.col.nb[f] * L.wh);
  for (int f = 0; f < L.row.n_stages; ++f) items_min = std::min<int64_t>(items_min, (int64_t)L.row.nb[f] * p.H);
  int64_t group = (256 + items_min - 1) / items_min;
  while (group > 1 && (ctas_that_fit((int)group) < 1 ||
                       (p.planes + group - 1) / group < (int64_t)di.sm_count * ctas_that_fit((int)group)))
    --group;
  if (group < 1) group = 1;
  while (group > 1 && p.sums_segment_planes > 0 && p.sums_segment_planes % group != 0) --group;  // no group straddles a segment
  L.group = (int)group;
  // the fold rides on the first row stage when that stage is not the last one, pairs up (m even) and is narrow
  // enough for two butterflies in registers (SONAR_B200_FFT_FUSE_FOLD=0: the separate pass)
  static const bool fuse_enabled = [] {
    const char* e = getenv("SONAR_B200_FFT_FUSE_FOLD");
    return e == nullptr || e[0] != '0';
  }();
  L.fuse_fold = fuse_enabled && L.row.n_stages >= 2 && L.row.radix[0] <= 8 && (L.row.m[0] & 1) == 0 ? 1 : 0;
  L.magic_T = 0;
  if (p.philox_grid_blocks != 0)
    L.magic_T = ~0ull / ((unsigned long long)p.philox_grid_blocks * kBlock) + 1ull;  // floor(2^64 / T) + 1 (T is no power of two > 2^63)
  // index arithmetic: 16-bit slot tables, 32-bit magic division exact for w * d < 2^32
  const int64_t nbatch_max = group * (L.wh > p.H ? L.wh : p.H);
  const int64_t items_max = nbatch_max * ((p.H > M ? p.H : M) / 2);
  if (items_max * nbatch_max >= (1ll << 32) || group * L.plane_elems >= (1 << 24)) return false;
  if (group * p.H * M * (int64_t)M >= (1ll << 32)) return false;
  L.magic_row_nb = magic_of(L.row.nb[L.row.n_stages - 1]);
  L.magic_wh = magic_of(L.wh);
  L.magic_cols = magic_of(L.group * L.wh);
  L.magic_rows = magic_of(L.group * p.H);
  *smem_out = batched_smem_bytes(p.H, p.W, L.group, L.col, L.row);
  // the register file holds 1024 threads at the kernel's 64 registers: split them over the CTAs that fit
  int ctas_per_sm = ctas_that_fit(L.group);
  if (ctas_per_sm < 1) ctas_per_sm = 1;
  *ctas_per_sm_out = ctas_per_sm;
  *threads_out = ctas_per_sm >= 4 ? 256 : ctas_per_sm == 3 ? 320 : kBatchedThreads;
  int64_t grid = (p.planes + L.group - 1) / L.group;
  // long launches (many passes per persistent CTA) run slightly faster on 256-thread CTAs: 9504 planes of 90x160
  // 383 vs 390 us; at one or two passes per CTA 320 threads are equal or better (1056 planes: 59.4 vs 57.4 us)
  if (ctas_per_sm == 3 && grid > 8 * (int64_t)di.sm_count * ctas_per_sm) *threads_out = 256;
  // persistent CTAs: the per-CTA tables (twiddles, slot maps) are built once and reused for every group
  // co-scheduling hint (sonar_set_grid_limit): fewer resident CTAs per SM, same CTA shape -- the registers and thread
  // slots left over go to a kernel on another stream
  const int limit = grid_limit_ctas_per_sm();
  const int resident = limit > 0 && limit < ctas_per_sm ? limit : ctas_per_sm;
  if (grid > (int64_t)di.sm_count * resident) grid = (int64_t)di.sm_count * resident;
  *grid_out = grid;
  return true;
}

// 0 = launched; -1 = not applicable (caller falls back to the generic kernel); > 0 = CUDA error
static int launch_spectral_batched(const SonarSpectralParams& p, cudaStream_t stream) {
  SpectralBatchedLaunch L;
  int threads = 0, ctas_per_sm = 0;
  int64_t grid = 0;
  size_t smem = 0;
  if (!plan_spectral_batched(p, &L, &threads, &ctas_per_sm, &grid, &smem)) return -1;
  auto kernel = p.in_real != nullptr   ? spectral_batched_kernel<INPUT_REAL>
                : p.in_spec != nullptr ? spectral_batched_kernel<INPUT_SPECTRUM>
                                       : spectral_batched_kernel<INPUT_PHILOX>;
  // Co-scheduling hint >= 3 CTAs per SM (sonar_set_grid_limit): 256-thread CTAs. 3 x 256 threads x 64 registers = 48 K
  // registers per SM leave 16 K = two 256-thread CTAs of the fused step resident on the same SM, so the issue-bound FFT
  // runs under the HBM-bound second part of the step launch. (Alone, 256 threads are as fast as 320: 179 vs 182 us for
  // 4224 planes of 90x160; a 48-register instantiation at 320 threads was slower, 193 us.)
  if (grid_limit_ctas_per_sm() >= 3 && ctas_per_sm >= 3 && threads > 256) threads = 256;
  cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (err != cudaSuccess) return (int)err;
  kernel<<<(unsigned)grid, threads, smem, stream>>>(L);
  err = cudaGetLastError();
  return err == cudaSuccess ? 0 : (int)err;
}

// =============================================================================================
// Cluster variant: planes whose half spectrum does not fit one SM (256x256: 264 KB) live in the DISTRIBUTED shared
// memory of a cluster of two CTAs -- still one HBM read and one HBM write per plane, no L2 scratch.
//  * column phase: CTA r holds columns [r * cols0, ...) of all H rows ([H][pitch_c], 133 KB at 256x256) and runs the
//    inverse column transforms of its share exactly like the batched kernel (gather from the global spectrum with
//    the gain applied, in place, natural row order out);
//  * exchange: row y belongs to CTA y / (H/2) in the row phase. Every thread takes its elements into registers,
//    the cluster synchronises (both buffers are now free to be overwritten), and each element is stored into its
//    row-phase slot [y % (H/2)][col] of the owning CTA -- a local store or a store into the peer's shared memory;
//  * row phase: fold, half-length inverse row transforms of the CTA's H/2 rows ([H/2][pitch_r]), last stage straight
//    to the global plane with the moments, as in the batched kernel.
// Real input (rfft2 front end) starts in the row layout -- forward rows of the CTA's half gathered from the global
// real plane, unfold -- crosses to the column layout for the forward and inverse column transforms around the gain,
// and crosses back: two exchanges.
// =============================================================================================
constexpr int kClusterThreads = 1024;
constexpr int kClusterMaxRegsElems = 20;  // complex elements a thread carries through the exchange (thread-local staging)

struct SpectralClusterLaunch {
  SonarSpectralParams p;
  AxisPlan col;   // length H
  AxisPlan row;   // length M = W / 2
  int wh;         // M + 1
  int cols0;      // columns of CTA 0 in the column phase (CTA 1: wh - cols0)
  int pitch_c;    // odd, >= cols0
  int rows_half;  // H / 2
  int pitch_r;    // odd, >= wh
  int buf_elems;  // max(H * pitch_c, rows_half * pitch_r)
  float scale;
  unsigned magic_cols[2], magic_rows, magic_row_nb, magic_wh;
};

// Moves the cluster's plane between its two shared-memory layouts through thread-local staging: every thread takes
// its elements out, the cluster synchronises (both buffers may now be overwritten), every element is stored into
// its slot of the owning CTA (a local store or a store into the peer's shared memory), the cluster synchronises.
//   TO_ROWS: [H][pitch_c] (my columns, all rows)  ->  [H/2][pitch_r] (my rows, all columns)
//  !TO_ROWS: the way back.
template <bool TO_ROWS>
__device__ __forceinline__ void cluster_exchange(cooperative_groups::cluster_group& cluster, const SpectralClusterLaunch& L,
                                                 float2* __restrict__ buf, float2* __restrict__ peer_buf, int rank, int H) {
  const int col0 = rank * L.cols0, my_cols = rank == 0 ? L.cols0 : L.wh - L.cols0, row0 = rank * L.rows_half;
  const int inner = TO_ROWS ? my_cols : L.wh;  // fastest index of the source walk
  const unsigned magic = TO_ROWS ? L.magic_cols[rank] : L.magic_wh;
  const int n = (TO_ROWS ? H : L.rows_half) * inner;
  // (the staging array lives in the thread's local memory -- L1 -- on purpose: 17 complex values per thread would
  // not fit the 64-register budget of a 1024-thread CTA next to the butterfly code)
  float2 v[kClusterMaxRegsElems];
#pragma unroll 1
  for (int i = 0; i < kClusterMaxRegsElems; ++i) {
    const int idx = threadIdx.x + i * kClusterThreads;
    if (idx < n) {
      const int r = (int)__umulhi((unsigned)idx, magic);
      v[i] = buf[r * (TO_ROWS ? L.pitch_c : L.pitch_r) + (idx - r * inner)];
    }
  }
  cluster.sync();  // every element of both buffers has been taken out: both may be overwritten
#pragma unroll 1
  for (int i = 0; i < kClusterMaxRegsElems; ++i) {
    const int idx = threadIdx.x + i * kClusterThreads;
    if (idx < n) {
      const int r = (int)__umulhi((unsigned)idx, magic);
      const int c = idx - r * inner;
      if (TO_ROWS) {  // element (row r, column col0 + c) goes to the CTA that owns row r
        const int owner = r >= L.rows_half ? 1 : 0;
        (owner == rank ? buf : peer_buf)[(r - owner * L.rows_half) * L.pitch_r + col0 + c] = v[i];
      } else {  // element (row row0 + r, column c) goes to the CTA that owns column c
        const int owner = c >= L.cols0 ? 1 : 0;
        (owner == rank ? buf : peer_buf)[(row0 + r) * L.pitch_c + (c - owner * L.cols0)] = v[i];
      }
    }
  }
  cluster.sync();  // my share is complete (the peer's stores included)
}

template <bool REAL>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kClusterThreads, 1)
spectral_cluster_kernel(const __grid_constant__ SpectralClusterLaunch L) {
  namespace cg = cooperative_groups;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const SonarSpectralParams& p = L.p;
  const int H = p.H, W = p.W, M = W >> 1, Wh = L.wh;
  float2* tw_h = reinterpret_cast<float2*>(smem_raw);
  float2* tw_m = tw_h + H;
  float2* tw_w = tw_m + M;
  float2* buf = tw_w + M;
  ushort2* tab_col = reinterpret_cast<ushort2*>(buf + L.buf_elems);
  ushort2* tab_row = tab_col + L.col.tab_size;
  unsigned short* pos_m = reinterpret_cast<unsigned short*>(tab_row + L.row.tab_size);
  unsigned short* idx_h = pos_m + M;
  float2* peer_buf = cluster.map_shared_rank(buf, rank ^ 1);
  build_axis_tables(L.col, tab_col, nullptr, idx_h);
  build_axis_tables(L.row, tab_row, pos_m, nullptr);
  for (int k = threadIdx.x; k < H; k += blockDim.x) {
    double sn, cs;
    sincospi(-2.0 * (double)k / (double)H, &sn, &cs);
    tw_h[k] = make_float2((float)cs, (float)sn);
  }
  for (int k = threadIdx.x; k < M; k += blockDim.x) {
    double sn, cs;
    sincospi(-2.0 * (double)k / (double)M, &sn, &cs);
    tw_m[k] = make_float2((float)cs, (float)sn);
    sincospi(2.0 * (double)k / (double)W, &sn, &cs);
    tw_w[k] = make_float2((float)cs, (float)sn);
  }
  __shared__ double seg_sums[2 * SONAR_SPECTRAL_MAX_SEGMENTS];
  if (threadIdx.x < 2 * SONAR_SPECTRAL_MAX_SEGMENTS) seg_sums[threadIdx.x] = 0.0;
  __syncthreads();
  const int col0 = rank * L.cols0, my_cols = rank == 0 ? L.cols0 : Wh - L.cols0;
  const int row0 = rank * L.rows_half;
  StageArgs a;
  a.buf = buf;
  a.pos_m = pos_m;
  a.idx_h = idx_h;
  a.group = 1;
  a.M = M;
  a.spec_plane_elems = H * Wh;
  a.gstride = Wh;
  a.magic_wh = 0;
  a.magic_nb = L.magic_row_nb;
  a.scale = L.scale;
  a.ms = 0.0f;
  a.mss = 0.0f;
  a.launch = nullptr;
  a.mask = p.mask == nullptr ? nullptr : p.mask + col0;
  auto column_phase_args = [&]() {
    a.tw = tw_h;
    a.wh = my_cols;
    a.pitch = L.pitch_c;
    a.plane_elems = H * L.pitch_c;
    a.nbatch = my_cols;
    a.magic_batch = L.magic_cols[rank];
  };
  auto row_phase_args = [&]() {
    a.tw = tw_m;
    a.pitch = L.pitch_r;
    a.nbatch = L.rows_half;
    a.magic_batch = L.magic_rows;
  };
  for (int64_t plane = blockIdx.x >> 1; plane < p.planes; plane += gridDim.x >> 1) {
    if constexpr (REAL) {
      // forward rows of my half of the plane (half length, conj trick; gathered from the global real rows), r2c
      // unfold, to the column layout, forward columns, gain (.) conj, inverse columns
      a.real_rows = reinterpret_cast<const float2*>(p.in_real + plane * (int64_t)H * W + (int64_t)row0 * W);
      row_phase_args();
      run_axis_dit<SRC_REAL_ROWS, false>(L.row, a, tab_row);
      pair_pass<true>(buf, L.rows_half, L.magic_rows, M, L.pitch_r, tw_w);
      __syncthreads();
      cluster_exchange<false>(cluster, L, buf, peer_buf, rank, H);
      column_phase_args();
      run_axis_dif<SINK_SLOTS, true>(L.col, a, tab_col);
      run_axis_dit<SRC_CONJ_MASK, true>(L.col, a, tab_col);
    } else {
      // inverse columns of my share of the columns, all rows (gather from the global spectrum with the gain applied)
      a.spec = reinterpret_cast<const float2*>(p.in_spec) + plane * (int64_t)H * Wh + col0;
      column_phase_args();
      run_axis_dit<SRC_SPECTRUM, true>(L.col, a, tab_col);
    }
    cluster_exchange<true>(cluster, L, buf, peer_buf, rank, H);
    // ---- row phase: my half of the rows, all columns ----
    pair_pass<false>(buf, L.rows_half, L.magic_rows, M, L.pitch_r, tw_w);
    __syncthreads();
    a.out_rows = reinterpret_cast<float2*>(p.out + plane * (int64_t)H * W + (int64_t)row0 * W);
    row_phase_args();
    run_axis_dif<SINK_GLOBAL, false>(L.row, a, tab_row);
    if (p.sums != nullptr) {
      const float ws = warp_sum(a.ms), wss = warp_sum(a.mss);
      if ((threadIdx.x & 31) == 0) {
        const int seg = p.sums_segment_planes > 0 ? (int)(plane / p.sums_segment_planes) : 0;
        atomicAdd(&seg_sums[2 * seg], (double)ws);
        atomicAdd(&seg_sums[2 * seg + 1], (double)wss);
      }
      a.ms = a.mss = 0.0f;
    }
    // (the next plane's first phase writes only my own buffer; the peer stores into it again only after the next
    // exchange's first barrier, which I reach after this row phase: no barrier needed here)
  }
  if (p.sums != nullptr) {
    __syncthreads();
    if (threadIdx.x < 2 * SONAR_SPECTRAL_MAX_SEGMENTS && seg_sums[threadIdx.x] != 0.0) atomicAdd(&p.sums[threadIdx.x], seg_sums[threadIdx.x]);
    if (p.sums_clear != nullptr && blockIdx.x == 0 && threadIdx.x < 2) p.sums_clear[threadIdx.x] = 0.0;
  }
  cluster.sync();  // no CTA exits while its peer may still store into its shared memory
}

static size_t cluster_smem_bytes(int H, int W, int buf_elems, const AxisPlan& col, const AxisPlan& row) {
  const int M = W / 2;
  size_t bytes = ((size_t)H + 2 * (size_t)M + (size_t)buf_elems) * sizeof(float2) +
                 (size_t)(col.tab_size + row.tab_size) * sizeof(ushort2) + (size_t)(M + H) * sizeof(unsigned short);
  return (bytes + 15) & ~(size_t)15;
}

// 0 = launched; -1 = not applicable; > 0 = CUDA error.
static bool plan_spectral_cluster(const SonarSpectralParams& p, SpectralClusterLaunch* out, size_t* smem_out) {
  if ((p.W & 1) || (p.H & 1) || p.W < 4 || p.H < 4) return false;
  if ((reinterpret_cast<uintptr_t>(p.out) & 7u) != 0) return false;
  if (p.in_real != nullptr && (reinterpret_cast<uintptr_t>(p.in_real) & 7u) != 0) return false;
  SpectralClusterLaunch& L = *out;
  L.p = p;
  if (!make_axis_plan(p.H, &L.col) || !make_axis_plan(p.W / 2, &L.row, /*ascending=*/true)) return false;
  const int M = p.W / 2;
  L.wh = M + 1;
  L.cols0 = (L.wh + 1) / 2;
  L.pitch_c = L.cols0 | 1;
  L.rows_half = p.H / 2;
  L.pitch_r = L.wh | 1;
  L.buf_elems = std::max(p.H * L.pitch_c, L.rows_half * L.pitch_r);
  L.scale = p.in_real != nullptr ? 0.5f * p.out_scale : p.out_scale;
  if ((int64_t)p.H * L.cols0 > (int64_t)kClusterMaxRegsElems * kClusterThreads) return false;
  if ((int64_t)L.rows_half * L.wh > (int64_t)kClusterMaxRegsElems * kClusterThreads) return false;
  L.magic_wh = magic_of(L.wh);
  if (L.buf_elems >= (1 << 24) || (int64_t)p.H * M * (int64_t)M >= (1ll << 32)) return false;
  L.magic_cols[0] = magic_of(L.cols0);
  L.magic_cols[1] = magic_of(L.wh - L.cols0);
  L.magic_rows = magic_of(L.rows_half);
  L.magic_row_nb = magic_of(L.row.nb[L.row.n_stages - 1]);
  *smem_out = cluster_smem_bytes(p.H, p.W, L.buf_elems, L.col, L.row);
  return *smem_out <= (size_t)device_info().max_smem_optin;
}

static int launch_spectral_cluster(const SonarSpectralParams& p, cudaStream_t stream) {
  SpectralClusterLaunch L;
  size_t smem = 0;
  if (!plan_spectral_cluster(p, &L, &smem)) return -1;
  auto kernel = p.in_real != nullptr ? spectral_cluster_kernel<true> : spectral_cluster_kernel<false>;
  cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (err != cudaSuccess) return (int)err;
  const DeviceInfo& di = device_info();
  const int64_t clusters = p.planes < di.sm_count / 2 ? p.planes : di.sm_count / 2;
  kernel<<<(unsigned)(2 * clusters), kClusterThreads, smem, stream>>>(L);
  err = cudaGetLastError();
  return err == cudaSuccess ? 0 : (int)err;
}

static bool make_plan(int n, FftPlan* plan) {
  plan->n = n;
  plan->n_factors = 0;
  plan->tab_size = 0;
  if (n > 65535) return false;
  int rem = n;
  auto push = [&](int f) {
    if (plan->n_factors >= SONAR_FFT_MAX_FACTORS) return false;
    plan->factors[plan->n_factors++] = f;
    return true;
  };
  while (rem % 4 == 0) {
    if (!push(4)) return false;
    rem /= 4;
  }
  for (int f = 2; (int64_t)f * f <= rem; ++f)
    while (rem % f == 0) {
      if (!push(f)) return false;
      rem /= f;
    }
  if (rem > 1 && !push(rem)) return false;
  int ns = 1;
  for (int f = 0; f < plan->n_factors; ++f) {
    plan->ns[f] = ns;
    plan->tab_off[f] = plan->tab_size;
    plan->tab_size += n / plan->factors[f];
    ns *= plan->factors[f];
  }
  return true;
}

static size_t fixed_smem_bytes(int H, int W, const FftPlan& ph, const FftPlan& pw) {
  const int nmax = H > W ? H : W;
  size_t tab = (size_t)(ph.tab_size + pw.tab_size);
  tab += tab & 1;  // keep the spectrum 8-byte aligned
  return ((size_t)H + W + (size_t)kFftWarps * 2 * nmax) * sizeof(float2) + tab * sizeof(ushort2);
}

}  // namespace sonar

extern "C" {

int64_t sonar_spectral_scratch_bytes(int H, int W) {
  using namespace sonar;
  if (H <= 0 || W <= 0) return 0;
  FftPlan ph, pw;
  if (!make_plan(H, &ph) || !make_plan(W, &pw)) return 0;
  const int wh = W / 2 + 1, wh_pad = wh | 1;
  const size_t fixed = fixed_smem_bytes(H, W, ph, pw);
  const size_t spec = (size_t)H * wh_pad * sizeof(float2);
  const DeviceInfo& di = device_info();
  if (fixed + spec <= (size_t)di.max_smem_optin) return 0;
  {  // planes the 2-CTA cluster kernel keeps in distributed shared memory need no scratch either
    SonarSpectralParams probe = {};
    probe.out = reinterpret_cast<float*>(uintptr_t{64});
    probe.in_spec = probe.out;
    probe.planes = 1;
    probe.H = H;
    probe.W = W;
    SpectralClusterLaunch C;
    size_t smem = 0;
    if (plan_spectral_cluster(probe, &C, &smem)) return 0;
  }
  return (int64_t)spec * di.sm_count * 2;
}

int sonar_spectral_plan(int H, int W, int64_t planes, int real_input, SonarSpectralPlanInfo* info) {
  using namespace sonar;
  if (info == nullptr || H <= 0 || W <= 0 || planes < 0) return (int)cudaErrorInvalidValue;
  *info = SonarSpectralPlanInfo{};
  SonarSpectralParams p = {};
  float* const aligned = reinterpret_cast<float*>(uintptr_t{64});  // alignment matters to the plan, the address does not
  p.out = aligned;
  p.in_real = real_input ? aligned : nullptr;
  p.in_spec = real_input ? nullptr : aligned;
  p.planes = planes;
  p.H = H;
  p.W = W;
  p.out_scale = 1.0f;
  SpectralBatchedLaunch L;
  int threads = 0, ctas_per_sm = 0;
  int64_t grid = 0;
  size_t smem = 0;
  if (!plan_spectral_batched(p, &L, &threads, &ctas_per_sm, &grid, &smem)) {
    SpectralClusterLaunch C;
    if (!plan_spectral_cluster(p, &C, &smem)) return 0;  // generic kernel
    const int64_t clusters = planes < device_info().sm_count / 2 ? planes : device_info().sm_count / 2;
    info->cluster = 1;
    info->group = 1;
    info->threads = kClusterThreads;
    info->ctas_per_sm = 1;
    info->grid = 2 * clusters;
    info->smem_bytes = (int64_t)smem;
    info->n_col_stages = C.col.n_stages;
    info->n_row_stages = C.row.n_stages;
    for (int f = 0; f < C.col.n_stages && f < SONAR_SPECTRAL_PLAN_MAX_STAGES; ++f) info->col_radix[f] = C.col.radix[f];
    for (int f = 0; f < C.row.n_stages && f < SONAR_SPECTRAL_PLAN_MAX_STAGES; ++f) info->row_radix[f] = C.row.radix[f];
    return 0;
  }
  info->batched = 1;
  info->group = L.group;
  info->threads = threads;
  info->ctas_per_sm = ctas_per_sm;
  info->grid = grid;
  info->smem_bytes = (int64_t)smem;
  info->n_col_stages = L.col.n_stages;
  info->n_row_stages = L.row.n_stages;
  for (int f = 0; f < L.col.n_stages && f < SONAR_SPECTRAL_PLAN_MAX_STAGES; ++f) info->col_radix[f] = L.col.radix[f];
  for (int f = 0; f < L.row.n_stages && f < SONAR_SPECTRAL_PLAN_MAX_STAGES; ++f) info->row_radix[f] = L.row.radix[f];
  return 0;
}

int sonar_spectral_filter_f32(const SonarSpectralParams* params, void* stream_) {
  using namespace sonar;
  if (params == nullptr) return (int)cudaErrorInvalidValue;
  SpectralLaunch L;
  L.p = *params;
  const SonarSpectralParams& p = L.p;
  if (p.planes <= 0) return 0;
  if (p.H <= 0 || p.W <= 0 || p.out == nullptr || p.sums_segment_planes < 0) return (int)cudaErrorInvalidValue;
  if (p.sums_segment_planes > 0 && (p.planes + p.sums_segment_planes - 1) / p.sums_segment_planes > SONAR_SPECTRAL_MAX_SEGMENTS)
    return (int)cudaErrorInvalidValue;
  if (p.in_real != nullptr && p.in_spec != nullptr) return (int)cudaErrorInvalidValue;
  const bool philox = p.in_real == nullptr && p.in_spec == nullptr;
  if (philox) {
    const int64_t floats = 2 * p.planes * (int64_t)p.H * (p.W / 2 + 1);
    if (p.philox_grid_blocks == 0 || p.philox_begin < 0 || (p.philox_begin & 1) || p.philox_begin + floats > p.philox_numel_total)
      return (int)cudaErrorInvalidValue;
  }
  {
    const int rc = launch_spectral_batched(p, (cudaStream_t)stream_);
    if (rc >= 0) return rc;
  }
  if (!philox) {  // too large for one SM's shared memory: a cluster of two
    const int rc = launch_spectral_cluster(p, (cudaStream_t)stream_);
    if (rc >= 0) return rc;
  }
  if (philox) return (int)cudaErrorInvalidValue;  // only the batched kernel regenerates its input (even W, radices 2..16)
  if (!make_plan(p.H, &L.plan_h) || !make_plan(p.W, &L.plan_w)) return (int)cudaErrorInvalidValue;
  L.wh = p.W / 2 + 1;
  L.wh_pad = L.wh | 1;
  L.nmax = p.H > p.W ? p.H : p.W;
  const DeviceInfo& di = device_info();
  const size_t fixed = fixed_smem_bytes(p.H, p.W, L.plan_h, L.plan_w);
  const size_t spec = (size_t)p.H * L.wh_pad * sizeof(float2);
  if (fixed > (size_t)di.max_smem_optin) return (int)cudaErrorInvalidValue;  // 1-D length too large
  L.spectrum_in_smem = (fixed + spec <= (size_t)di.max_smem_optin) ? 1 : 0;
  const size_t smem = fixed + (L.spectrum_in_smem ? spec : 0);
  // two CTAs per SM when shared memory allows (the kernel is built for <= 64 registers / thread)
  const int per_sm = (int)((size_t)(di.max_smem_optin + 1024) / (smem + 1024)) >= 2 ? 2 : 1;
  int grid = di.sm_count * per_sm;
  if (!L.spectrum_in_smem && p.scratch == nullptr) return (int)cudaErrorInvalidValue;
  if ((int64_t)grid > p.planes) grid = (int)p.planes;
  SONAR_CUDA_TRY(cudaFuncSetAttribute(spectral_plane_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  spectral_plane_kernel<<<grid, kFftThreads, smem, (cudaStream_t)stream_>>>(L);
  SONAR_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
