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
};

// tw[k] = exp(-2*pi*i*k/n); inverse transforms conjugate on the fly.
__device__ __forceinline__ float2 twiddle(const float2* __restrict__ tw, int idx, bool inverse) {
  const float2 w = tw[idx];
  return inverse ? make_float2(w.x, -w.y) : w;
}

// Warp-cooperative Stockham autosort FFT of length plan.n. Input in `a`; returns the buffer that
// holds the result (a or b). Unnormalised. All 32 lanes must call.
__device__ float2* warp_fft(float2* a, float2* b, const FftPlan& plan, const float2* __restrict__ tw, bool inverse,
                            int lane) {
  const int N = plan.n;
  int Ns = 1;
  float2* src = a;
  float2* dst = b;
  for (int f = 0; f < plan.n_factors; ++f) {
    const int R = plan.factors[f];
    const int stride = N / R;          // distance between the R inputs of one butterfly
    const int tw_step = N / (Ns * R);  // w_N^(tw_step * k * s) == exp(-2 pi i k s / (Ns R))
    if (R == 2) {
      for (int j = lane; j < stride; j += 32) {
        const int k = j % Ns;
        const float2 v0 = src[j];
        const float2 v1 = cmul(src[j + stride], twiddle(tw, k * tw_step, inverse));
        const int j0 = (j / Ns) * Ns * 2 + k;
        dst[j0] = cadd(v0, v1);
        dst[j0 + Ns] = csub(v0, v1);
      }
    } else if (R == 4) {
      for (int j = lane; j < stride; j += 32) {
        const int k = j % Ns;
        const int base = k * tw_step;
        const float2 v0 = src[j];
        const float2 v1 = cmul(src[j + stride], twiddle(tw, base, inverse));
        const float2 v2 = cmul(src[j + 2 * stride], twiddle(tw, 2 * base, inverse));
        const float2 v3 = cmul(src[j + 3 * stride], twiddle(tw, 3 * base, inverse));
        const float2 a0 = cadd(v0, v2), a1 = csub(v0, v2), a2 = cadd(v1, v3);
        const float2 d = csub(v1, v3);
        // multiply by -i (forward) or +i (inverse)
        const float2 a3 = inverse ? make_float2(-d.y, d.x) : make_float2(d.y, -d.x);
        const int j0 = (j / Ns) * Ns * 4 + k;
        dst[j0] = cadd(a0, a2);
        dst[j0 + Ns] = cadd(a1, a3);
        dst[j0 + 2 * Ns] = csub(a0, a2);
        dst[j0 + 3 * Ns] = csub(a1, a3);
      }
    } else {
      // generic (prime) radix: each output is a direct R-term sum; one table lookup per term
      const int rot = N / R;
      for (int o = lane; o < N; o += 32) {
        const int j = o % stride;   // butterfly
        const int t = o / stride;   // which output of it
        const int k = j % Ns;
        const int step = (k * tw_step + t * rot) % N;
        float2 acc = src[j];
        int idx = 0;
        for (int s = 1; s < R; ++s) {
          idx += step;
          if (idx >= N) idx -= N;
          acc = cadd(acc, cmul(src[j + s * stride], twiddle(tw, idx, inverse)));
        }
        dst[(j / Ns) * Ns * R + k + t * Ns] = acc;
      }
    }
    __syncwarp();
    Ns *= R;
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

__global__ void __launch_bounds__(kFftThreads, 1)
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
  float2* S = L.spectrum_in_smem ? (tw_w + W + (size_t)kFftWarps * 2 * L.nmax)
                                 : (reinterpret_cast<float2*>(p.scratch) + (size_t)blockIdx.x * H * WhP);

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

  for (int64_t plane = blockIdx.x; plane < p.planes; plane += gridDim.x) {
    // ---------------- phase A: fill S ----------------
    if (p.in_real != nullptr) {
      const float* src = p.in_real + plane * (int64_t)H * W;
      for (int y = 2 * warp; y < H; y += 2 * kFftWarps) {
        const bool pair = (y + 1) < H;
        for (int x = lane; x < W; x += 32)
          bufA[x] = make_float2(src[(int64_t)y * W + x], pair ? src[(int64_t)(y + 1) * W + x] : 0.0f);
        __syncwarp();
        const float2* Z = warp_fft(bufA, bufB, L.plan_w, tw_w, false, lane);
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
        cur = warp_fft(bufA, bufB, L.plan_h, tw_h, false, lane);
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
      const float2* res = warp_fft(cur, other, L.plan_h, tw_h, true, lane);
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
      const float2* z = warp_fft(bufA, bufB, L.plan_w, tw_w, true, lane);
      for (int x = lane; x < W; x += 32) {
        const float2 v = z[x];
        dst[(int64_t)y * W + x] = v.x * p.out_scale;
        if (pair) dst[(int64_t)(y + 1) * W + x] = v.y * p.out_scale;
      }
      __syncwarp();
    }
    __syncthreads();
  }
}

static bool make_plan(int n, FftPlan* plan) {
  plan->n = n;
  plan->n_factors = 0;
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
  if (n == 1) plan->n_factors = 0;
  return true;
}

}  // namespace sonar

extern "C" {

int64_t sonar_spectral_scratch_bytes(int H, int W) {
  using namespace sonar;
  if (H <= 0 || W <= 0) return 0;
  const int wh = W / 2 + 1, wh_pad = wh | 1, nmax = H > W ? H : W;
  const size_t fixed = ((size_t)H + W + (size_t)kFftWarps * 2 * nmax) * sizeof(float2);
  const size_t spec = (size_t)H * wh_pad * sizeof(float2);
  const DeviceInfo& di = device_info();
  if (fixed + spec <= (size_t)di.max_smem_optin) return 0;
  return (int64_t)spec * di.sm_count;
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
  if (!make_plan(p.H, &L.plan_h) || !make_plan(p.W, &L.plan_w)) return (int)cudaErrorInvalidValue;
  L.wh = p.W / 2 + 1;
  L.wh_pad = L.wh | 1;
  L.nmax = p.H > p.W ? p.H : p.W;
  const DeviceInfo& di = device_info();
  const size_t fixed = ((size_t)p.H + p.W + (size_t)kFftWarps * 2 * L.nmax) * sizeof(float2);
  const size_t spec = (size_t)p.H * L.wh_pad * sizeof(float2);
  if (fixed > (size_t)di.max_smem_optin) return (int)cudaErrorInvalidValue;  // 1-D length too large
  L.spectrum_in_smem = (fixed + spec <= (size_t)di.max_smem_optin) ? 1 : 0;
  const size_t smem = fixed + (L.spectrum_in_smem ? spec : 0);
  int grid = di.sm_count;
  if (L.spectrum_in_smem) {
    const int per_sm = (int)((size_t)di.max_smem_optin / (smem + 1024));
    grid = di.sm_count * (per_sm < 1 ? 1 : (per_sm > 2 ? 2 : per_sm));
  } else if (p.scratch == nullptr) {
    return (int)cudaErrorInvalidValue;
  }
  if ((int64_t)grid > p.planes) grid = (int)p.planes;
  SONAR_CUDA_TRY(cudaFuncSetAttribute(spectral_plane_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  spectral_plane_kernel<<<grid, kFftThreads, smem, (cudaStream_t)stream_>>>(L);
  SONAR_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
