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
// Batched-lane variant for the hot case (half spectrum in, even W, radices 2/3/4/5): every FFT stage
// is ONE pass of the whole CTA over the whole plane, with the lanes of a warp running the SAME
// butterfly of 32 different transforms (columns: 32 adjacent k, contiguous in shared memory; rows: 32
// rows, conflict-free thanks to the odd pitch). Twiddles and output offsets are warp-uniform, there
// is no per-lane index arithmetic, no idle lanes for short transforms and no per-warp scratch.
// Rows use the half-length c2r trick: for Hermitian X of even length W, with M = W/2,
//   Z[k] = (X[k] + conj X[M-k]) + i (X[k] - conj X[M-k]) e^{+2 pi i k / W},   k = 0..M-1
//   z = sum_k Z[k] e^{+2 pi i k n / M}  ==>  x[2n] = Re z[n], x[2n+1] = Im z[n]
// which also reproduces irfft2's treatment of the reference's NON-Hermitian input (the imaginary
// parts of the k = 0 and k = M bins are dropped, nothing else of the upper half is read).
// The round-1 warp-per-transform kernel needed ~237 issued instructions per output element at
// 90x160 (ncu: 62 % issue-slot busy, instruction bound); this form needs ~50.
// =============================================================================================
constexpr int kBatchedThreads = 1024;
constexpr int kBatchedMaxStages = 16;

struct AxisPlan {
  int n;
  int n_stages;
  int radix[kBatchedMaxStages];
  int ns[kBatchedMaxStages];
  int nb[kBatchedMaxStages];       // butterflies per transform in the stage (n / radix)
  int tab_off[kBatchedMaxStages];  // offset of the stage's butterfly table
  int tab_size;
};

struct SpectralBatchedLaunch {
  SonarSpectralParams p;
  AxisPlan col;  // length H
  AxisPlan row;  // length M = W / 2
  int wh;        // M + 1
  int pitch;     // odd, >= wh
};

__device__ __forceinline__ float2 cmul_conj(float2 a, float2 w) {  // a * conj(w)
  return make_float2(a.x * w.x + a.y * w.y, a.y * w.x - a.x * w.y);
}

// butterfly table of one axis: for butterfly j of stage f, .x = output base (j / Ns) * Ns * R + (j % Ns),
// .y = twiddle base (j % Ns) * N / (Ns * R). Built once per CTA so the stage loops hold no division.
__device__ void build_axis_table(const AxisPlan& plan, ushort2* __restrict__ tab) {
  for (int f = 0; f < plan.n_stages; ++f) {
    const int R = plan.radix[f], Ns = plan.ns[f], nb = plan.n / R, unit = plan.n / (Ns * R);
    for (int j = threadIdx.x; j < nb; j += blockDim.x) {
      const int k0 = j % Ns;
      tab[plan.tab_off[f] + j] = make_ushort2((unsigned short)((j / Ns) * Ns * R + k0), (unsigned short)(k0 * unit));
    }
  }
}

// Where a stage reads its inputs from.
enum StageSource : int {
  SRC_SMEM = 0,      // the other ping-pong buffer
  SRC_SPECTRUM = 1,  // first column stage: global half spectrum (.) gain mask
  SRC_FOLD = 2       // first row stage: Z[k] = (X[k] + conj X[M-k]) + i (X[k] - conj X[M-k]) e^{+2 pi i k / W}
};

struct StageIo {
  const float2* src;   // SRC_SMEM / SRC_FOLD: shared-memory buffer; SRC_SPECTRUM: global plane
  const float* mask;   // SRC_SPECTRUM
  const float2* tw_w;  // SRC_FOLD
  int M;               // SRC_FOLD
  int src_pitch;       // SRC_SPECTRUM: row length of the global plane (Wh)
};

// Input t' (0..R-1) of butterfly j for transform b: one base offset per (j, b), then a fixed step.
template <int SOURCE>
struct StageInputs {
  const float2* p0;   // SRC_SMEM / SRC_SPECTRUM: &src[first element]; SRC_FOLD: row base
  const float* m0;    // SRC_SPECTRUM: &mask[first element] (or nullptr)
  const float2* tw_w; // SRC_FOLD
  int step;           // distance between consecutive inputs (elements)
  int t0, M;          // SRC_FOLD: first k, fold length

  __device__ __forceinline__ float2 get(int i) const {
    if (SOURCE == SRC_SPECTRUM) {
      float2 v = p0[i * step];
      if (m0 != nullptr) {
        const float g = __ldg(m0 + i * step);
        v.x *= g;
        v.y *= g;
      }
      return v;
    }
    if (SOURCE == SRC_FOLD) {
      const int t = t0 + i * step;
      float2 xk = p0[t], xm = p0[M - t];
      if (t == 0) {  // c2r ignores the imaginary parts of the DC and Nyquist bins
        xk.y = 0.0f;
        xm.y = 0.0f;
      }
      const float2 a = make_float2(xk.x + xm.x, xk.y - xm.y);  // X[k] + conj X[M-k]
      const float2 d = make_float2(xk.x - xm.x, xk.y + xm.y);  // X[k] - conj X[M-k]
      const float2 bt = cmul(d, tw_w[t]);
      return make_float2(a.x - bt.y, a.y + bt.x);  // A + i B
    }
    return p0[i * step];
  }
};

template <int SOURCE>
__device__ __forceinline__ StageInputs<SOURCE> stage_inputs(const StageIo& io, int j, int nb, int b, int stride_t,
                                                            int stride_b) {
  StageInputs<SOURCE> in;
  in.tw_w = io.tw_w;
  in.M = io.M;
  in.m0 = nullptr;
  in.t0 = j;
  if (SOURCE == SRC_SPECTRUM) {  // element (row t, column b) of the global plane
    in.p0 = io.src + (j * io.src_pitch + b);
    in.m0 = io.mask != nullptr ? io.mask + (j * io.src_pitch + b) : nullptr;
    in.step = nb * io.src_pitch;
  } else if (SOURCE == SRC_FOLD) {  // transform index = k along the row of transform b
    in.p0 = io.src + b * stride_b;
    in.step = nb;
  } else {
    in.p0 = io.src + (j * stride_t + b * stride_b);
    in.step = nb * stride_t;
  }
  return in;
}

// Radix-R inverse butterfly on registers: v[t] already multiplied by its twiddle.
template <int R>
__device__ __forceinline__ void butterfly_inverse(float2 (&v)[R]) {
  if (R == 2) {
    const float2 a = v[0], b = v[1];
    v[0] = cadd(a, b);
    v[1] = csub(a, b);
  } else if (R == 4) {
    const float2 a0 = cadd(v[0], v[2]), a1 = csub(v[0], v[2]), a2 = cadd(v[1], v[3]);
    const float2 d = csub(v[1], v[3]);
    const float2 a3 = make_float2(-d.y, d.x);  // * (+i)
    v[0] = cadd(a0, a2);
    v[1] = cadd(a1, a3);
    v[2] = csub(a0, a2);
    v[3] = csub(a1, a3);
  } else if (R == 3) {
    const float2 t1 = cadd(v[1], v[2]);
    const float2 t2 = make_float2(v[0].x - 0.5f * t1.x, v[0].y - 0.5f * t1.y);
    const float2 d = csub(v[1], v[2]);
    const float2 t3 = make_float2(-0.86602540378443865f * d.y, 0.86602540378443865f * d.x);
    v[0] = cadd(v[0], t1);
    v[1] = cadd(t2, t3);
    v[2] = csub(t2, t3);
  } else {  // R == 5
    constexpr float c1 = 0.30901699437494742f, c2 = -0.80901699437494742f;
    constexpr float s1 = 0.95105651629515357f, s2 = 0.58778525229247313f;
    const float2 v0 = v[0];
    const float2 a1 = cadd(v[1], v[4 % R]), a2 = cadd(v[2], v[3]), b1 = csub(v[1], v[4 % R]), b2 = csub(v[2], v[3]);
    const float2 m1 = make_float2(v0.x + c1 * a1.x + c2 * a2.x, v0.y + c1 * a1.y + c2 * a2.y);
    const float2 m2 = make_float2(v0.x + c2 * a1.x + c1 * a2.x, v0.y + c2 * a1.y + c1 * a2.y);
    const float2 e1 = make_float2(s1 * b1.x + s2 * b2.x, s1 * b1.y + s2 * b2.y);
    const float2 e2 = make_float2(s2 * b1.x - s1 * b2.x, s2 * b1.y - s1 * b2.y);
    const float2 n1 = make_float2(-e1.y, e1.x), n2 = make_float2(-e2.y, e2.x);  // * (+i)
    v[0] = make_float2(v0.x + a1.x + a2.x, v0.y + a1.y + a2.y);
    v[1] = cadd(m1, n1);
    v[2] = cadd(m2, n2);
    v[3] = csub(m2, n2);
    v[4 % R] = csub(m1, n1);
  }
}

// One inverse Stockham stage over a batch of transforms, radix R a template parameter (no per-butterfly
// radix dispatch). Element t of transform b lives at t * stride_t + b * stride_b. The CTA's warps form a
// (butterfly, chunk) grid: a warp owns butterflies j, j + jw, ... and, for each, the 32 transforms of its
// chunk(s); table entry and twiddles are fetched once per butterfly and reused across chunks.
// How the CTA's warps tile (butterfly, chunk-of-32-transforms) for one axis: computed once per kernel
// (it only depends on the batch size), not once per stage call -- the integer divisions of this setup
// were ~20 % of the issued instructions when every stage recomputed them.
struct WarpGrid {
  int nchunks;  // 32-transform chunks of the batch
  int cw;       // warps along the chunk axis
  int jw;       // warps along the butterfly axis
  int wj, wc;   // this warp's coordinates (wj >= jw: idle)
};

__device__ __forceinline__ WarpGrid make_warp_grid(int nbatch) {
  const int warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  WarpGrid g;
  g.nchunks = (nbatch + 31) >> 5;
  g.cw = g.nchunks < nwarps ? g.nchunks : nwarps;
  g.jw = nwarps / g.cw;
  g.wj = warp / g.cw;
  g.wc = warp - g.wj * g.cw;
  return g;
}

template <int SOURCE, int R>
__device__ __forceinline__ void batched_stage_radix(const StageIo& io, float2* __restrict__ dst, int nb, int Ns,
                                                    const float2* __restrict__ tw, const ushort2* __restrict__ tab,
                                                    int nbatch, int stride_t, int stride_b, const WarpGrid& g) {
  const int lane = threadIdx.x & 31;
  const int nchunks = g.nchunks, cw = g.cw, jw = g.jw, wj = g.wj, wc = g.wc;
  if (wj >= jw) return;  // nwarps % cw leftover warps idle for this stage
  const int out_step = Ns * stride_t;
  for (int j = wj; j < nb; j += jw) {
    const ushort2 e = tab[j];
    float2 w[R];
#pragma unroll
    for (int t = 1; t < R; ++t) w[t] = tw[t * (int)e.y];
    for (int c = wc; c < nchunks; c += cw) {
      const int b = (c << 5) + lane;
      if (b >= nbatch) continue;
      const StageInputs<SOURCE> in = stage_inputs<SOURCE>(io, j, nb, b, stride_t, stride_b);
      float2 v[R];
      v[0] = in.get(0);
#pragma unroll
      for (int t = 1; t < R; ++t) v[t] = cmul_conj(in.get(t), w[t]);
      butterfly_inverse<R>(v);
      float2* out = dst + (int)e.x * stride_t + b * stride_b;
#pragma unroll
      for (int t = 0; t < R; ++t) out[t * out_step] = v[t];
    }
  }
}

template <int SOURCE>
__device__ __forceinline__ void batched_stage_inverse(const StageIo& io, float2* __restrict__ dst, int nb, int R, int Ns,
                                                      const float2* __restrict__ tw, const ushort2* __restrict__ tab,
                                                      int nbatch, int stride_t, int stride_b, const WarpGrid& g) {
  switch (R) {  // block-uniform
    case 4: batched_stage_radix<SOURCE, 4>(io, dst, nb, Ns, tw, tab, nbatch, stride_t, stride_b, g); break;
    case 2: batched_stage_radix<SOURCE, 2>(io, dst, nb, Ns, tw, tab, nbatch, stride_t, stride_b, g); break;
    case 3: batched_stage_radix<SOURCE, 3>(io, dst, nb, Ns, tw, tab, nbatch, stride_t, stride_b, g); break;
    default: batched_stage_radix<SOURCE, 5>(io, dst, nb, Ns, tw, tab, nbatch, stride_t, stride_b, g); break;
  }
}

__global__ void __launch_bounds__(kBatchedThreads, 1)
spectral_batched_kernel(SpectralBatchedLaunch L) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const SonarSpectralParams& p = L.p;
  const int H = p.H, W = p.W, M = W >> 1, Wh = L.wh, P = L.pitch;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  float2* tw_h = reinterpret_cast<float2*>(smem_raw);  // e^{-2 pi i k / H}
  float2* tw_m = tw_h + H;                             // e^{-2 pi i k / M}
  float2* tw_w = tw_m + M;                             // e^{+2 pi i k / W}, k < M
  float2* buf0 = tw_w + M;
  float2* buf1 = buf0 + (size_t)H * P;
  ushort2* tab_col = reinterpret_cast<ushort2*>(buf1 + (size_t)H * P);
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
  const WarpGrid grid_col = make_warp_grid(Wh), grid_row = make_warp_grid(H);
  float ms = 0.0f, mss = 0.0f;
  for (int64_t plane = blockIdx.x; plane < p.planes; plane += gridDim.x) {
    // ---- columns: inverse complex FFT of length H, batch = Wh columns (lanes = adjacent k). The first
    // stage reads the global half spectrum (coalesced along k) and applies the gain on the fly. ----
    float2* cur = buf1;  // "previous" buffer; the first stage ignores it
    float2* oth = buf0;
    StageIo io;
    io.mask = p.mask;
    io.tw_w = tw_w;
    io.M = M;
    io.src_pitch = Wh;
    for (int f = 0; f < L.col.n_stages; ++f) {
      if (f == 0) {
        io.src = reinterpret_cast<const float2*>(p.in_spec) + plane * (int64_t)H * Wh;
        batched_stage_inverse<SRC_SPECTRUM>(io, oth, L.col.nb[0], L.col.radix[0], L.col.ns[0], tw_h, tab_col, Wh, P, 1, grid_col);
      } else {
        io.src = cur;
        batched_stage_inverse<SRC_SMEM>(io, oth, L.col.nb[f], L.col.radix[f], L.col.ns[f], tw_h, tab_col + L.col.tab_off[f], Wh, P, 1, grid_col);
      }
      __syncthreads();
      float2* t = cur;
      cur = oth;
      oth = t;
    }
    // ---- rows: inverse complex FFT of length M, batch = H rows (lanes = 32 rows, odd pitch). The first
    // stage folds the Hermitian half row into M complex points while loading. ----
    for (int f = 0; f < L.row.n_stages; ++f) {
      io.src = cur;
      if (f == 0)
        batched_stage_inverse<SRC_FOLD>(io, oth, L.row.nb[0], L.row.radix[0], L.row.ns[0], tw_m, tab_row, H, 1, P, grid_row);
      else
        batched_stage_inverse<SRC_SMEM>(io, oth, L.row.nb[f], L.row.radix[f], L.row.ns[f], tw_m, tab_row + L.row.tab_off[f], H, 1, P, grid_row);
      __syncthreads();
      float2* t = cur;
      cur = oth;
      oth = t;
    }
    // ---- store: x[y][2n], x[y][2n+1] = z[y][n] * scale, coalesced float2, moments on the fly ----
    float* dst = p.out + plane * (int64_t)H * W;
    for (int y = warp; y < H; y += nwarps) {
      for (int n = lane; n < M; n += 32) {
        const float2 z = cur[y * P + n];
        const float2 o = make_float2(z.x * p.out_scale, z.y * p.out_scale);
        *reinterpret_cast<float2*>(dst + y * W + 2 * n) = o;
        ms += o.x + o.y;
        mss += o.x * o.x + o.y * o.y;
      }
    }
    __syncthreads();  // the next plane's first stage overwrites a buffer this pass reads
  }
  commit_moments(p.sums, p.sums_clear, ms, mss);
}

static bool make_axis_plan(int n, AxisPlan* plan) {
  plan->n = n;
  plan->n_stages = 0;
  plan->tab_size = 0;
  if (n > 65535) return false;
  int rem = n, ns = 1;
  auto push = [&](int r) {
    if (plan->n_stages >= kBatchedMaxStages) return false;
    plan->radix[plan->n_stages] = r;
    plan->ns[plan->n_stages] = ns;
    plan->nb[plan->n_stages] = n / r;
    plan->tab_off[plan->n_stages] = plan->tab_size;
    plan->tab_size += n / r;
    ++plan->n_stages;
    ns *= r;
    return true;
  };
  while (rem % 4 == 0) {
    if (!push(4)) return false;
    rem /= 4;
  }
  for (int r : {2, 3, 5})
    while (rem % r == 0) {
      if (!push(r)) return false;
      rem /= r;
    }
  return rem == 1;  // any other prime factor: the generic kernel handles it
}

static size_t batched_smem_bytes(int H, int W, const AxisPlan& col, const AxisPlan& row) {
  const int M = W / 2, wh = M + 1, pitch = wh | 1;
  return ((size_t)H + 2 * (size_t)M + 2 * (size_t)H * pitch) * sizeof(float2) +
         (size_t)(col.tab_size + row.tab_size) * sizeof(ushort2);
}

// 0 = launched; -1 = not applicable (caller falls back to the generic kernel); > 0 = CUDA error
static int launch_spectral_batched(const SonarSpectralParams& p, cudaStream_t stream) {
  if (p.in_spec == nullptr || (p.W & 1) || p.W < 2 || p.H < 1) return -1;
  if ((reinterpret_cast<uintptr_t>(p.out) & 7u) != 0) return -1;
  SpectralBatchedLaunch L;
  L.p = p;
  if (!make_axis_plan(p.H, &L.col) || !make_axis_plan(p.W / 2, &L.row)) return -1;
  // the spectrum load and the Hermitian fold ride on the first stage of each axis: both must exist
  if (L.col.n_stages == 0 || L.row.n_stages == 0) return -1;
  L.wh = p.W / 2 + 1;
  L.pitch = L.wh | 1;
  const DeviceInfo& di = device_info();
  const size_t smem = batched_smem_bytes(p.H, p.W, L.col, L.row);
  if (smem > (size_t)di.max_smem_optin || (int64_t)p.H * L.pitch >= (1 << 24)) return -1;
  cudaError_t err = cudaFuncSetAttribute(spectral_batched_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (err != cudaSuccess) return (int)err;
  const int per_sm = (int)((size_t)(di.max_smem_optin + 1024) / (smem + 1024)) >= 2 ? 2 : 1;
  int64_t grid = (int64_t)di.sm_count * per_sm;
  if (grid > p.planes) grid = p.planes;
  spectral_batched_kernel<<<(unsigned)grid, kBatchedThreads, smem, stream>>>(L);
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
