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
#include <cstdlib>
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
    __syncthreads();
  }
  commit_moments(p.sums, p.sums_clear, ms, mss);
}

// =============================================================================================
// Batched variant for the hot case (even W, lengths that factor into radices 2/3/4/5/8/9/10/16): every
// FFT stage is ONE pass of the whole CTA over a group of planes, one register-resident radix-R butterfly
// per thread. Work items (butterfly j, transform b) are flattened with b fastest, so adjacent lanes run
// the same butterfly of adjacent transforms (columns: adjacent k, contiguous in shared memory; rows:
// adjacent rows, conflict-free thanks to the odd pitch) whatever the batch size; small planes (UNet
// activations, 32x32) are processed several per CTA so the 1024 threads stay busy.
// Large radices keep the pass count low: 90 = 10 x 9 and 80 = 10 x 8 are two passes each (radices 6..16
// are Cooley-Tukey compositions of the 2/3/4/5 butterflies inside registers, inner twiddles compile-time
// constants), and the first pass of every axis has no twiddle multiplications at all.
// Rows use the half-length c2r trick: for Hermitian X of even length W, with M = W/2,
//   Z[k] = (X[k] + conj X[M-k]) + i (X[k] - conj X[M-k]) e^{+2 pi i k / W},   k = 0..M-1
//   z = sum_k Z[k] e^{+2 pi i k n / M}  ==>  x[2n] = Re z[n], x[2n+1] = Im z[n]
// which also reproduces irfft2's treatment of the reference's NON-Hermitian input (the imaginary
// parts of the k = 0 and k = M bins are dropped, nothing else of the upper half is read).
// Real input (rfft2 front end: PowerFilterNoiseItem, OneF / GreenTest, FreeU-Extreme ffilter) runs the
// same inverse stages on conjugated data (FFT(z) = conj(IFFT(conj z))): rows r2c by the half-length
// trick in reverse (the "unfold" rides on the first forward column stage), then columns, gain, and the
// inverse path above -- the plane still makes exactly one trip from and one trip to HBM.
// History: warp-per-transform kernel 237 issued instructions per output element at 90x160, radix 2..5
// batched-lane kernel 158 (ncu r01g: 62 % issue-slot busy, 43 % of it integer / address arithmetic).
// =============================================================================================
constexpr int kBatchedThreads = 1024;
constexpr int kBatchedMaxStages = 16;

struct AxisPlan {
  int n;
  int n_stages;
  int radix[kBatchedMaxStages];
  int ns[kBatchedMaxStages];       // product of the radices before the stage
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
  float scale;      // out_scale (x 1/2 for real input: the r2c unfold leaves a factor 2)
  unsigned magic_m, magic_wh, magic_cols, magic_rows;  // ceil(2^32 / d) for d = M, wh, group * wh, group * H
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

// Where the first stage of an axis reads its inputs from.
enum StageSource : int {
  SRC_SMEM = 0,       // the other ping-pong buffer
  SRC_SMEM_CONJ = 1,  // the other buffer, conjugated (first forward row stage over the packed real rows)
  SRC_SPECTRUM = 2,   // first inverse column stage, spectrum input: global half spectrum (.) gain mask
  SRC_CONJ_MASK = 3,  // first inverse column stage, real input: conj(forward result) (.) gain mask
  SRC_FOLD = 4,       // first inverse row stage: Z[k] = (X[k] + conj X[M-k]) + i (X[k] - conj X[M-k]) e^{+2 pi i k / W}
  SRC_UNFOLD = 5      // first forward column stage: conj X[k] from the half-length row transforms
};

struct StageArgs {
  const float2* src;     // shared-memory source (SRC_SPECTRUM: the group's first global plane)
  float2* dst;
  const float2* tw;      // e^{-2 pi i k / n} of the axis
  const ushort2* tab;    // butterfly table of the stage (unused by first stages)
  const float* mask;     // (H, wh) gain or nullptr
  const float2* tw_w;    // e^{+2 pi i k / W}, k < M
  int nb, ns;
  int nbatch;            // transforms in this pass (valid planes x columns or rows)
  unsigned magic_batch;  // ceil(2^32 / nbatch)
  unsigned magic_wh;     // ceil(2^32 / wh)
  int group;             // planes per pass the kernel was launched with (1: no plane split of b)
  int wh, pitch, plane_elems, M;
  int spec_plane_elems;  // H * wh (global spectrum plane)
};

// One inverse Stockham stage over a batch of transforms, one radix-R butterfly per thread and iteration.
// Element t of transform b lives at base(b) + t * stride: columns base = plane * plane_elems + k, stride =
// pitch; rows base = b * pitch (b = plane * H + y), stride = 1.
template <int SOURCE, int R, bool COLS, bool FIRST>
__device__ __forceinline__ void run_stage(const StageArgs& a) {
  const int total = a.nb * a.nbatch;
  const int stride = COLS ? a.pitch : 1;
  for (int w = threadIdx.x; w < total; w += blockDim.x) {
    const int j = (int)__umulhi((unsigned)w, a.magic_batch);
    const int b = w - j * a.nbatch;
    int g = 0, k = b, base;
    if (COLS) {
      if (a.group > 1) {
        g = (int)__umulhi((unsigned)b, a.magic_wh);
        k = b - g * a.wh;
      }
      base = g * a.plane_elems + k;
    } else {
      base = b * a.pitch;
    }
    float2 v[R];
    if (SOURCE == SRC_SPECTRUM) {
      const int e0 = j * a.wh + k, estep = a.nb * a.wh;
      const float2* gp = a.src + ((int64_t)g * a.spec_plane_elems + e0);
#pragma unroll
      for (int t = 0; t < R; ++t) v[t] = __ldg(gp + t * estep);
      if (a.mask != nullptr) {
        const float* mp = a.mask + e0;
#pragma unroll
        for (int t = 0; t < R; ++t) {
          const float gain = __ldg(mp + t * estep);
          v[t].x *= gain;
          v[t].y *= gain;
        }
      }
    } else if (SOURCE == SRC_FOLD) {
      const float2* p0 = a.src + base;
#pragma unroll
      for (int t = 0; t < R; ++t) {
        const int idx = j + t * a.nb;
        float2 xk = p0[idx], xm = p0[a.M - idx];
        if (t == 0 && j == 0) {  // c2r ignores the imaginary parts of the DC and Nyquist bins
          xk.y = 0.0f;
          xm.y = 0.0f;
        }
        const float2 s = make_float2(xk.x + xm.x, xk.y - xm.y);  // X[k] + conj X[M-k]
        const float2 d = make_float2(xk.x - xm.x, xk.y + xm.y);  // X[k] - conj X[M-k]
        const float2 wd = cmul(d, a.tw_w[idx]);
        v[t] = make_float2(s.x - wd.y, s.y + wd.x);  // s + i w d
      }
    } else if (SOURCE == SRC_UNFOLD) {
      // conj X[k] = 1/2 [(Y[k] + conj Y[M-k]) + i e^{+2 pi i k / W} (Y[k] - conj Y[M-k])], Y = conj FFT_M(packed
      // row), indices mod M (the 1/2 is folded into the output scale)
      const bool nyq = k == a.M;
      const int kk = nyq ? 0 : k, mm = (k == 0 || nyq) ? 0 : a.M - k;
      const float2 wk = nyq ? make_float2(-1.0f, 0.0f) : a.tw_w[k];
      const float2* yp = a.src + g * a.plane_elems + j * a.pitch;
      const int ystep = a.nb * a.pitch;
#pragma unroll
      for (int t = 0; t < R; ++t) {
        const float2 yk = yp[t * ystep + kk], ym = yp[t * ystep + mm];
        const float2 s = make_float2(yk.x + ym.x, yk.y - ym.y);
        const float2 d = make_float2(yk.x - ym.x, yk.y + ym.y);
        const float2 wd = cmul(d, wk);
        v[t] = make_float2(s.x - wd.y, s.y + wd.x);
      }
    } else {
      const float2* p = a.src + base + j * stride;
      const int step = a.nb * stride;
#pragma unroll
      for (int t = 0; t < R; ++t) v[t] = p[t * step];
      if (SOURCE == SRC_SMEM_CONJ) {
#pragma unroll
        for (int t = 0; t < R; ++t) v[t].y = -v[t].y;
      }
      if (SOURCE == SRC_CONJ_MASK) {
        const float* mp = a.mask + (j * a.wh + k);  // COLS only
        const int mstep = a.nb * a.wh;
#pragma unroll
        for (int t = 0; t < R; ++t) {
          const float gain = a.mask != nullptr ? __ldg(mp + t * mstep) : 1.0f;
          v[t] = make_float2(v[t].x * gain, -v[t].y * gain);
        }
      }
    }
    int out_base = j * R;  // first stage: Ns = 1
    if (!FIRST) {
      const ushort2 e = a.tab[j];
      out_base = e.x;
      const float2* twp = a.tw;
      const int tstep = e.y;
#pragma unroll
      for (int t = 1; t < R; ++t) v[t] = cmul_conj(v[t], twp[t * tstep]);
    }
    dft_inverse<R>(v);
    float2* out = a.dst + base + out_base * stride;
    const int ostep = (FIRST ? 1 : a.ns) * stride;
#pragma unroll
    for (int t = 0; t < R; ++t) out[dft_out_index<R>(t) * ostep] = v[t];
  }
}

template <int SOURCE, bool COLS, bool FIRST>
__device__ __forceinline__ void dispatch_stage(int R, const StageArgs& a) {
  switch (R) {  // block-uniform
    case 16: run_stage<SOURCE, 16, COLS, FIRST>(a); break;
    case 10: run_stage<SOURCE, 10, COLS, FIRST>(a); break;
    case 9: run_stage<SOURCE, 9, COLS, FIRST>(a); break;
    case 8: run_stage<SOURCE, 8, COLS, FIRST>(a); break;
    case 5: run_stage<SOURCE, 5, COLS, FIRST>(a); break;
    case 4: run_stage<SOURCE, 4, COLS, FIRST>(a); break;
    case 3: run_stage<SOURCE, 3, COLS, FIRST>(a); break;
    default: run_stage<SOURCE, 2, COLS, FIRST>(a); break;
  }
}

// All stages of one axis. `first_src` feeds stage 0 (SOURCE kind FIRST_SOURCE); later stages ping-pong.
// On return `cur` holds the result.
template <int FIRST_SOURCE, bool COLS>
__device__ __forceinline__ void run_axis(const AxisPlan& plan, StageArgs& a, const float2* first_src, const ushort2* tab,
                                         float2*& cur, float2*& oth) {
  for (int f = 0; f < plan.n_stages; ++f) {
    a.dst = oth;
    a.nb = plan.nb[f];
    a.ns = plan.ns[f];
    a.tab = tab + plan.tab_off[f];
    if (f == 0) {
      a.src = first_src;
      dispatch_stage<FIRST_SOURCE, COLS, true>(plan.radix[0], a);
    } else {
      a.src = cur;
      dispatch_stage<SRC_SMEM, COLS, false>(plan.radix[f], a);
    }
    __syncthreads();
    float2* t = cur;
    cur = oth;
    oth = t;
  }
}

// butterfly table of one axis: for butterfly j of stage f, .x = output base (j / Ns) * Ns * R + (j % Ns),
// .y = twiddle base (j % Ns) * N / (Ns * R). Built once per CTA so the stage loops hold no division.
__device__ void build_axis_table(const AxisPlan& plan, ushort2* __restrict__ tab) {
  for (int f = 1; f < plan.n_stages; ++f) {
    const int R = plan.radix[f], Ns = plan.ns[f], nb = plan.n / R, unit = plan.n / (Ns * R);
    for (int j = threadIdx.x; j < nb; j += blockDim.x) {
      const int k0 = j % Ns;
      tab[plan.tab_off[f] + j] = make_ushort2((unsigned short)((j / Ns) * Ns * R + k0), (unsigned short)(k0 * unit));
    }
  }
}

__host__ __device__ __forceinline__ unsigned magic_of(int d) { return (unsigned)((0x100000000ull + (unsigned)d - 1u) / (unsigned)d); }

template <bool REAL>
__global__ void __launch_bounds__(kBatchedThreads, 1)
spectral_batched_kernel(SpectralBatchedLaunch L) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const SonarSpectralParams& p = L.p;
  const int H = p.H, W = p.W, M = W >> 1, Wh = L.wh, P = L.pitch, G = L.group;
  float2* tw_h = reinterpret_cast<float2*>(smem_raw);  // e^{-2 pi i k / H}
  float2* tw_m = tw_h + H;                             // e^{-2 pi i k / M}
  float2* tw_w = tw_m + M;                             // e^{+2 pi i k / W}, k < M
  float2* buf0 = tw_w + M;
  float2* buf1 = buf0 + (size_t)G * L.plane_elems;
  const unsigned magic_m = L.magic_m;
  ushort2* tab_col = reinterpret_cast<ushort2*>(buf1 + (size_t)G * L.plane_elems);
  ushort2* tab_row = tab_col + L.col.tab_size;
  build_axis_table(L.col, tab_col);
  build_axis_table(L.row, tab_row);
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
  a.mask = p.mask;
  a.tw_w = tw_w;
  a.group = G;
  a.wh = Wh;
  a.pitch = P;
  a.plane_elems = L.plane_elems;
  a.M = M;
  a.spec_plane_elems = H * Wh;
  a.magic_wh = L.magic_wh;
  int magic_for = G;
  unsigned magic_cols = L.magic_cols, magic_rows = L.magic_rows;
  float ms = 0.0f, mss = 0.0f;
  for (int64_t plane0 = (int64_t)blockIdx.x * G; plane0 < p.planes; plane0 += (int64_t)gridDim.x * G) {
    const int valid = (int)(p.planes - plane0 < G ? p.planes - plane0 : G);
    if (valid != magic_for) {  // ragged last group only
      magic_cols = magic_of(valid * Wh);
      magic_rows = magic_of(valid * H);
      magic_for = valid;
    }
    const int rows_total = valid * H;
    float2* cur = buf0;
    float2* oth = buf1;
    if (REAL) {
      // packed rows: (x[2n], x[2n+1]) as one complex value, coalesced float2 loads, odd pitch in smem
      const float2* src = reinterpret_cast<const float2*>(p.in_real + plane0 * (int64_t)H * W);
      for (int i = threadIdx.x; i < rows_total * M; i += blockDim.x) {
        const int r = (int)__umulhi((unsigned)i, magic_m);
        cur[r * P + (i - r * M)] = __ldg(src + i);
      }
      __syncthreads();
      a.tw = tw_m;
      a.nbatch = rows_total;
      a.magic_batch = magic_rows;
      run_axis<SRC_SMEM_CONJ, false>(L.row, a, cur, tab_row, cur, oth);
      a.tw = tw_h;
      a.nbatch = valid * Wh;
      a.magic_batch = magic_cols;
      run_axis<SRC_UNFOLD, true>(L.col, a, cur, tab_col, cur, oth);
      run_axis<SRC_CONJ_MASK, true>(L.col, a, cur, tab_col, cur, oth);
    } else {
      // columns: inverse complex FFT of length H, batch = columns of the group's planes; the first stage
      // reads the global half spectrum (coalesced along k) and applies the gain on the fly
      a.tw = tw_h;
      a.nbatch = valid * Wh;
      a.magic_batch = magic_cols;
      run_axis<SRC_SPECTRUM, true>(L.col, a, reinterpret_cast<const float2*>(p.in_spec) + plane0 * (int64_t)H * Wh, tab_col, cur,
                                   oth);
    }
    // rows: inverse complex FFT of length M, batch = rows of the group's planes; the first stage folds the
    // Hermitian half row into M complex points while loading
    a.tw = tw_m;
    a.nbatch = rows_total;
    a.magic_batch = magic_rows;
    run_axis<SRC_FOLD, false>(L.row, a, cur, tab_row, cur, oth);
    // store: x[y][2n], x[y][2n+1] = z[y][n] * scale, coalesced float2, moments on the fly
    float2* dst = reinterpret_cast<float2*>(p.out + plane0 * (int64_t)H * W);
    for (int i = threadIdx.x; i < rows_total * M; i += blockDim.x) {
      const int r = (int)__umulhi((unsigned)i, magic_m);
      const float2 z = cur[r * P + (i - r * M)];
      const float2 o = make_float2(z.x * L.scale, z.y * L.scale);
      dst[i] = o;
      ms += o.x + o.y;
      mss += o.x * o.x + o.y * o.y;
    }
    __syncthreads();  // the next group's first stage overwrites a buffer this pass reads
  }
  commit_moments(p.sums, p.sums_clear, ms, mss);
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

static bool make_axis_plan(int n, AxisPlan* plan) {
  plan->n = n;
  plan->n_stages = 0;
  plan->tab_size = 0;
  if (n < 2 || n > 65535) return false;
  int cur[kBatchedMaxStages], best[kBatchedMaxStages], best_depth = kBatchedMaxStages + 1, best_sum = 1 << 30;
  search_plan(n, 0, 0, cur, best, &best_depth, &best_sum);
  if (best_depth > kBatchedMaxStages) return false;  // another prime factor: the generic kernel handles it
  int ns = 1;
  for (int f = 0; f < best_depth; ++f) {
    const int r = best[f];
    plan->radix[f] = r;
    plan->ns[f] = ns;
    plan->nb[f] = n / r;
    plan->tab_off[f] = plan->tab_size;
    plan->tab_size += n / r;
    ns *= r;
  }
  plan->n_stages = best_depth;
  return true;
}

static size_t batched_smem_bytes(int H, int W, int group, const AxisPlan& col, const AxisPlan& row) {
  const int M = W / 2, wh = M + 1, pitch = wh | 1;
  return ((size_t)H + 2 * (size_t)M + 2 * (size_t)group * H * pitch) * sizeof(float2) +
         (size_t)(col.tab_size + row.tab_size) * sizeof(ushort2);
}

// 0 = launched; -1 = not applicable (caller falls back to the generic kernel); > 0 = CUDA error
static int launch_spectral_batched(const SonarSpectralParams& p, cudaStream_t stream) {
  if ((p.W & 1) || p.W < 4 || p.H < 2) return -1;
  if ((reinterpret_cast<uintptr_t>(p.out) & 7u) != 0) return -1;
  if (p.in_real != nullptr && (reinterpret_cast<uintptr_t>(p.in_real) & 7u) != 0) return -1;
  SpectralBatchedLaunch L;
  L.p = p;
  if (!make_axis_plan(p.H, &L.col) || !make_axis_plan(p.W / 2, &L.row)) return -1;
  const int M = p.W / 2;
  L.wh = M + 1;
  L.pitch = L.wh | 1;
  L.plane_elems = p.H * L.pitch;
  L.scale = p.in_real != nullptr ? 0.5f * p.out_scale : p.out_scale;
  const DeviceInfo& di = device_info();
  if (batched_smem_bytes(p.H, p.W, 1, L.col, L.row) > (size_t)di.max_smem_optin) return -1;
  // planes per pass: enough butterflies to occupy the CTA in the leanest stage, without starving SMs
  const int64_t items_col = (int64_t)(p.H / L.col.radix[0]) * L.wh, items_row = (int64_t)(M / L.row.radix[0]) * p.H;
  const int64_t items_min = items_col < items_row ? items_col : items_row;
  int64_t group = (kBatchedThreads + items_min - 1) / items_min;
  const int64_t per_sm = (p.planes + di.sm_count - 1) / di.sm_count;
  if (group > per_sm) group = per_sm;
  if (const char* e = getenv("SONAR_SPECTRAL_GROUP")) group = atoi(e);  // experiment
  while (group > 1 && batched_smem_bytes(p.H, p.W, (int)group, L.col, L.row) > (size_t)di.max_smem_optin) --group;
  if (group < 1) group = 1;
  L.group = (int)group;
  // index arithmetic: 16-bit butterfly tables, 32-bit magic division exact for w * nbatch < 2^32
  const int64_t nbatch_max = group * (L.wh > p.H ? L.wh : p.H);
  const int64_t items_max = nbatch_max * ((p.H > M ? p.H : M) / 2);
  if (items_max * nbatch_max >= (1ll << 32) || group * L.plane_elems >= (1 << 24)) return -1;
  if (group * p.H * M * (int64_t)M >= (1ll << 32)) return -1;
  L.magic_m = magic_of(M);
  L.magic_wh = magic_of(L.wh);
  L.magic_cols = magic_of(L.group * L.wh);
  L.magic_rows = magic_of(L.group * p.H);
  const size_t smem = batched_smem_bytes(p.H, p.W, L.group, L.col, L.row);
  auto kernel = p.in_real != nullptr ? spectral_batched_kernel<true> : spectral_batched_kernel<false>;
  cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (err != cudaSuccess) return (int)err;
  int64_t grid = (p.planes + L.group - 1) / L.group;
  int threads = kBatchedThreads, ctas_per_sm = 1;
  if (const char* e = getenv("SONAR_SPECTRAL_THREADS")) {  // experiment
    threads = atoi(e);
    ctas_per_sm = kBatchedThreads / threads;
    if ((smem + 1024) * ctas_per_sm > (size_t)di.max_smem_optin + 1024) ctas_per_sm = 1;
  }
  if (grid > di.sm_count * ctas_per_sm) grid = di.sm_count * ctas_per_sm;
  kernel<<<(unsigned)grid, threads, smem, stream>>>(L);
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
  return (int64_t)spec * di.sm_count * 2;
}

int sonar_spectral_filter_f32(const SonarSpectralParams* params, void* stream_) {
  using namespace sonar;
  if (params == nullptr) return (int)cudaErrorInvalidValue;
  SpectralLaunch L;
  L.p = *params;
  const SonarSpectralParams& p = L.p;
  if (p.planes <= 0) return 0;
  if (p.H <= 0 || p.W <= 0 || p.out == nullptr) return (int)cudaErrorInvalidValue;
  if ((p.in_real == nullptr) == (p.in_spec == nullptr)) return (int)cudaErrorInvalidValue;
  {
    const int rc = launch_spectral_batched(p, (cudaStream_t)stream_);
    if (rc >= 0) return rc;
  }
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
