// RNG-fused noise synthesis: pyramid and Perlin noise (and their blend) computed element-wise straight from the
// Philox stream, without materialising the full-size base draws.
//
// Reference: PyramidNoiseGenerator.generate py/noise_generation.py:621-649 (base randn + sum_i bilinear-upsampled
// randn level * discount^i), PerlinOldNoiseGenerator.generate :478-493 (U(0,1) / div_fac + 2 x 2 gradient stencils
// shared by the batch), BlendedNoise py/noise.py:1391-1405 (blend_function(noise_1, noise_2, t)).
//
// The unfused path writes every base draw to HBM and reads it back (C3: ~240 MB of traffic for a 16.8 MB result,
// five launches). Here one thread owns one Philox CALL of the emulated ATen launch -- virtual thread vt, call k --
// exactly like ATen's own distribution kernel, so all four lanes of every call are used: elements li = vt + T (4k +
// lane), T apart, i.e. four different planes. Every full-size draw of the graph (pyramid base, the full-size
// pyramid level 0, the Perlin base) has the same element count, hence the same (vt, k, lane) -> element map, and the
// thread produces its four output elements from registers. What is NOT element-wise is small and comes from L2:
//  * coarse pyramid levels (36^2, 7^2, 1^2 per plane at C3): materialised by the Philox fill kernel, sampled with
//    per-CTA tap tables (same rows-first bilinear arithmetic as pyramid_rows_kernel);
//  * the Perlin stencil: a (C, H, W) table per iteration built once by perlin_tables_kernel (the reference
//    broadcasts it over the batch), its corner angles regenerated from their own Philox draws.
// Consecutive threads touch consecutive addresses, so the scalar stores coalesce into 128-byte transactions.
#include "common.cuh"
#include "../../include/sonar_b200.h"

namespace sonar {

struct MixTap {  // one x (or y) tap of a resampled level: source indices and the weight of the second one
  unsigned short i0, i1;
  float w1;
};

__device__ __forceinline__ MixTap make_tap(int dst, int in_size, int out_size, bool bilinear) {
  MixTap t;
  const float scale = (float)in_size / (float)out_size;
  if (bilinear) {  // ATen area_pixel_compute_source_index, align_corners = false (ATen/native/UpSample.h:289-312)
    float src = scale * ((float)dst + 0.5f) - 0.5f;
    if (src < 0.0f) src = 0.0f;
    int i0 = (int)src;
    if (i0 > in_size - 1) i0 = in_size - 1;
    const float l1 = fminf(fmaxf(src - (float)i0, 0.0f), 1.0f);
    t.i0 = (unsigned short)i0;
    t.i1 = (unsigned short)(i0 + (i0 < in_size - 1 ? 1 : 0));
    t.w1 = l1;
  } else {  // nearest-exact
    const int i = (int)floorf(((float)dst + 0.5f) * scale);
    t.i0 = t.i1 = (unsigned short)(i < in_size - 1 ? i : in_size - 1);
    t.w1 = 0.0f;
  }
  return t;
}

// one Philox lane as curand_uniform4 would return it, then at::uniform_'s transform
__device__ __forceinline__ float philox_uniform_lane(const uint4& r, int lane, float from, float to) {
  const unsigned x = lane == 0 ? r.x : lane == 1 ? r.y : lane == 2 ? r.z : r.w;
  return uniform_transform(_curand_uniform(x), from, to);
}

// ---------------------------------------------------------------------------------------------
// Perlin stencil tables: table[it][c][y][x] = blend of the four corner gradients of pixel (y, x) dotted with
// (+-0.5, +-0.5); corner angle (c, gy, gx) = element (c (H+1) + gy) (W+1) + gx of a uniform [0, 2 pi) draw.
// ---------------------------------------------------------------------------------------------
struct PerlinTablesLaunch {
  float* table[SONAR_MIX_MAX_TABLES];
  uint64_t offset[SONAR_MIX_MAX_TABLES];
  uint32_t grid_blocks[SONAR_MIX_MAX_TABLES];
  uint64_t seed;
  int C, H, W, blend_mode;
};

__device__ __forceinline__ float grid_angle(const PhiloxStream& st, unsigned long long magic_T, int64_t li) {
  const uint64_t q = __umul64hi((uint64_t)li, magic_T);
  const uint32_t vt = (uint32_t)((uint64_t)li - q * st.threads);
  return philox_uniform_lane(philox_raw(st, vt, q >> 2), (int)(q & 3u), 0.0f, 6.283185307179586f);
}

// One CTA = one 16 x 16 pixel tile of one channel of one iteration: its 17 x 17 corner angles are drawn and turned
// into (sin, cos) ONCE in shared memory (1.13 evaluations per pixel instead of 4), then every pixel blends its four.
constexpr int kPerlinTile = 16;

__global__ void __launch_bounds__(kBlock)
perlin_tables_kernel(PerlinTablesLaunch L) {
  __shared__ float2 corner[kPerlinTile + 1][kPerlinTile + 2];
  const int it = blockIdx.z;
  const PhiloxStream st{L.seed, L.offset[it], L.grid_blocks[it] * (uint32_t)kBlock};
  const unsigned long long magic_T = ~0ull / (unsigned long long)st.threads + 1ull;
  const int tiles_x = (L.W + kPerlinTile - 1) / kPerlinTile;
  const int c = blockIdx.y;
  const int y0 = (blockIdx.x / tiles_x) * kPerlinTile, x0 = (blockIdx.x % tiles_x) * kPerlinTile;
  const int64_t gw = L.W + 1;
  for (int i = threadIdx.x; i < (kPerlinTile + 1) * (kPerlinTile + 1); i += blockDim.x) {
    const int cy = i / (kPerlinTile + 1), cx = i - cy * (kPerlinTile + 1);
    const int gy = y0 + cy, gx = x0 + cx;
    float sn = 0.0f, cs = 0.0f;
    if (gy <= L.H && gx <= L.W) sincosf(grid_angle(st, magic_T, ((int64_t)c * (L.H + 1) + gy) * gw + gx), &sn, &cs);
    corner[cy][cx] = make_float2(sn, cs);
  }
  __syncthreads();
  const int ty = threadIdx.x / kPerlinTile, tx = threadIdx.x % kPerlinTile;
  const int y = y0 + ty, x = x0 + tx;
  if (y >= L.H || x >= L.W) return;
  const float2 k00 = corner[ty][tx], k01 = corner[ty][tx + 1], k10 = corner[ty + 1][tx], k11 = corner[ty + 1][tx + 1];
  // identical arithmetic to perlin_from_corners (noise_ops.cu): gradient . offset for the four corners
  const float d00 = __fadd_rn(__fmul_rn(k00.y, 0.5f), __fmul_rn(k00.x, 0.5f));
  const float d01 = __fadd_rn(__fmul_rn(k01.y, -0.5f), __fmul_rn(k01.x, 0.5f));
  const float d10 = __fadd_rn(__fmul_rn(k10.y, 0.5f), __fmul_rn(k10.x, -0.5f));
  const float d11 = __fadd_rn(__fmul_rn(k11.y, -0.5f), __fmul_rn(k11.x, -0.5f));
  const float row0 = blend<float>(L.blend_mode, d00, d01, 0.5f);
  const float row1 = blend<float>(L.blend_mode, d10, d11, 0.5f);
  L.table[it][((int64_t)c * L.H + y) * L.W + x] = blend<float>(L.blend_mode, row0, row1, 0.5f);
}

// ---------------------------------------------------------------------------------------------
// the fused element-wise kernel
// ---------------------------------------------------------------------------------------------
struct MixLaunch {
  SonarNoiseMixParams p;
  // (n * magic) >> 40 == n / d for n < 2^24, d < 2^16 (d = W, C); T = q_T * (H * W) + r_T
  unsigned long long magic_w, magic_c;
  unsigned q_T, r_T;
  uint32_t k_lo, k_hi;
};

// Per-CTA copy of what the element loop needs of a pyramid level, in shared memory: one 16-byte load per level
// (the parameter block in the constant bank costs a dependent LDC per field when indexed by the level loop).
struct LevelDesc {
  const float* src;  // materialised level, or nullptr: THE full-size level (regenerated from the stream)
  int lw;
  int plane_stride;  // lh * lw
  float weight;
  int pad_;
};
struct TermShared {
  const MixTap* xtab;  // [n_resampled][W]
  const MixTap* ytab;  // [n_resampled][H]
  const LevelDesc* lev;
  int n_levels;
};

// The four elements of a Philox call at once: (local plane, y, x, offset in the (C, H, W) tables) per lane. Every
// level issues the loads of all four lanes before any arithmetic (16 independent L2 / L1 requests in flight per
// thread: dependent tap -> load -> blend chains at 37-50 % occupancy are latency bound otherwise).
struct Lane4 {
  int lp[4], y[4], x[4], chw[4];
};

template <int KIND>
__device__ __forceinline__ void term_values4(const SonarMixTerm& t, const TermShared& sh, const float (&base)[4],
                                             const float (&level_full)[4], const Lane4& e, int H, int W, float inv_div,
                                             float (&out)[4]) {
  if (KIND == SONAR_TERM_PYRAMID) {
#pragma unroll
    for (int i = 0; i < 4; ++i) out[i] = t.base_scale != 1.0f ? base[i] * t.base_scale : base[i];
    const bool bilinear = t.mode == SONAR_RESAMPLE_BILINEAR;
    int res = 0;
    for (int l = 0; l < sh.n_levels; ++l) {
      const LevelDesc d = sh.lev[l];
      float sv[4];
      if (d.src == nullptr) {
#pragma unroll
        for (int i = 0; i < 4; ++i) sv[i] = level_full[i];
      } else {
        const MixTap* ty = sh.ytab + res * H;
        const MixTap* tx = sh.xtab + res * W;
        ++res;
        MixTap a[4], b[4];
        float r00[4], r01[4], r10[4], r11[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          a[i] = ty[e.y[i]];
          b[i] = tx[e.x[i]];
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float* src = d.src + e.lp[i] * d.plane_stride;  // (planes * lh * lw < 2^31: checked by the launcher)
          const float* r0 = src + (int)a[i].i0 * d.lw;
          const float* r1 = src + (int)a[i].i1 * d.lw;
          r00[i] = __ldg(r0 + b[i].i0);
          if (bilinear) {
            r01[i] = __ldg(r0 + b[i].i1);
            r10[i] = __ldg(r1 + b[i].i0);
            r11[i] = __ldg(r1 + b[i].i1);
          }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (bilinear) {
            // rows first, like pyramid_rows_kernel: blend the two source rows, then the two columns
            const float wy0 = 1.0f - a[i].w1;
            const float v0 = wy0 * r00[i] + a[i].w1 * r10[i];
            const float v1 = wy0 * r01[i] + a[i].w1 * r11[i];
            sv[i] = (1.0f - b[i].w1) * v0 + b[i].w1 * v1;
          } else {
            sv[i] = r00[i];
          }
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) out[i] = out[i] + __fmul_rn(sv[i], d.weight);  // noise += upsampled.mul_(discount ** i)
    }
    return;
  }
  // SONAR_TERM_PERLIN
  float tv[SONAR_MIX_MAX_TABLES][4];
#pragma unroll
  for (int it = 0; it < SONAR_MIX_MAX_TABLES; ++it)
    if (it < t.iterations) {
#pragma unroll
      for (int i = 0; i < 4; ++i) tv[it][i] = __ldg(t.tables[it] + e.chw[i]);
    }
#pragma unroll
  for (int i = 0; i < 4; ++i) out[i] = div_by(base[i], t.div_fac, inv_div);
#pragma unroll
  for (int it = 0; it < SONAR_MIX_MAX_TABLES; ++it)
    if (it < t.iterations) {
#pragma unroll
      for (int i = 0; i < 4; ++i) out[i] += tv[it][i];
    }
}

// the four lanes of the full-size draws of a term for call (vt, k)
template <int KIND>
__device__ __forceinline__ void term_draws(const SonarMixTerm& t, uint64_t seed, uint32_t threads, uint32_t vt, uint32_t k,
                                           float (&base4)[4], float (&level4)[4]) {
  const PhiloxStream st{seed, t.base_offset, threads};
  if (KIND == SONAR_TERM_PERLIN) {
    const float4 u = philox_uniform4(st, vt, k);
    base4[0] = uniform_transform(u.x, t.uniform_from, t.uniform_to);
    base4[1] = uniform_transform(u.y, t.uniform_from, t.uniform_to);
    base4[2] = uniform_transform(u.z, t.uniform_from, t.uniform_to);
    base4[3] = uniform_transform(u.w, t.uniform_from, t.uniform_to);
    return;
  }
  const float4 z = philox_normal4(st, vt, k);
  base4[0] = z.x; base4[1] = z.y; base4[2] = z.z; base4[3] = z.w;
  if (KIND != SONAR_TERM_PYRAMID || t.full_level < 0) return;
  const PhiloxStream ls{seed, t.level_offset[t.full_level], threads};
  const float4 w = philox_normal4(ls, vt, k);
  level4[0] = w.x; level4[1] = w.y; level4[2] = w.z; level4[3] = w.w;
}

// One thread = one virtual ATen thread vt (the grid is the emulated ATen grid); per Philox call k it owns the
// elements li = vt + T (4k + lane). The (plane, offset in plane) pair of lane 0 is found once and stepped by T from
// lane to lane (a compare instead of divisions). FULL: the launch covers the whole draw (no slice bounds per lane).
template <int KIND_A, int KIND_B, bool FULL>
__global__ void __launch_bounds__(kBlock, 3)
noise_mix_kernel(const __grid_constant__ MixLaunch L) {
  extern __shared__ __align__(16) unsigned char mix_smem[];
  const SonarNoiseMixParams& p = L.p;
  const int H = p.H, W = p.W;
  // per-CTA level descriptors and tap tables of the resampled pyramid levels of both terms
  LevelDesc* descs = reinterpret_cast<LevelDesc*>(mix_smem);
  MixTap* cur = reinterpret_cast<MixTap*>(descs + 2 * SONAR_PYRAMID_MAX_LEVELS);
  TermShared sa{nullptr, nullptr, descs, 0}, sb{nullptr, nullptr, descs + SONAR_PYRAMID_MAX_LEVELS, 0};
  for (int side = 0; side < 2; ++side) {
    const SonarMixTerm& t = side == 0 ? p.a : p.b;
    if (t.kind != SONAR_TERM_PYRAMID) continue;
    LevelDesc* lev = descs + side * SONAR_PYRAMID_MAX_LEVELS;
    int n_res = 0;
    for (int l = 0; l < t.n_levels; ++l) n_res += t.levels[l] != nullptr;
    MixTap* xt = cur;
    MixTap* yt = cur + n_res * W;
    cur += n_res * (W + H);
    int res = 0;
    const bool bil = t.mode == SONAR_RESAMPLE_BILINEAR;
    for (int l = 0; l < t.n_levels; ++l) {
      if (threadIdx.x == 0) lev[l] = LevelDesc{t.levels[l], t.level_w[l], t.level_h[l] * t.level_w[l], t.weights[l], 0};
      if (t.levels[l] == nullptr) continue;
      for (int i = threadIdx.x; i < W; i += blockDim.x) xt[res * W + i] = make_tap(i, t.level_w[l], W, bil);
      for (int i = threadIdx.x; i < H; i += blockDim.x) yt[res * H + i] = make_tap(i, t.level_h[l], H, bil);
      ++res;
    }
    (side == 0 ? sa : sb) = TermShared{xt, yt, lev, t.n_levels};
  }
  __syncthreads();
  const uint32_t T = p.grid_blocks * (uint32_t)kBlock;
  const uint32_t vt = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned hw = (unsigned)(H * W);
  const int64_t begin = p.begin, end = p.begin + p.n;
  const unsigned plane0 = (unsigned)(begin / hw);  // global index of the first local plane (slices are plane aligned)
  const float inv_div_a = 1.0f / p.a.div_fac, inv_div_b = 1.0f / p.b.div_fac;
  float* __restrict__ out = p.out - begin;  // indexed by the GLOBAL element index
  const unsigned long long magic_w = L.magic_w, magic_c = L.magic_c;
  const unsigned q_T = L.q_T, r_T = L.r_T, C = (unsigned)p.C;
  float ms = 0.0f, mss = 0.0f;
  // (plane, rem) of element vt + T * 4 * k_lo, stepped lane by lane
  uint64_t li = (uint64_t)vt + (uint64_t)T * 4u * L.k_lo;
  unsigned plane = (unsigned)(li / hw), rem = (unsigned)(li - (uint64_t)plane * hw);
  for (uint32_t k = L.k_lo; k <= L.k_hi; ++k) {
    if ((int64_t)li >= end) break;
    const bool any = FULL || (int64_t)(li + 3ull * T) >= begin;
    float a4[4], b4[4], la4[4] = {0.f, 0.f, 0.f, 0.f}, lb4[4] = {0.f, 0.f, 0.f, 0.f};
    if (any) {
      term_draws<KIND_A>(p.a, p.seed, T, vt, k, a4, la4);
      if (KIND_B != SONAR_TERM_NONE) term_draws<KIND_B>(p.b, p.seed, T, vt, k, b4, lb4);
    }
    Lane4 e;
    bool ok[4];
    const uint64_t li_call = li;
#pragma unroll
    for (int lane = 0; lane < 4; ++lane) {
      ok[lane] = any && (FULL ? (int64_t)li < end : ((int64_t)li >= begin && (int64_t)li < end));
      // lanes outside the slice / past the end read plane0's element 0 (a valid address) and store nothing
      const unsigned r = ok[lane] ? rem : 0u, pl = ok[lane] ? plane : plane0;
      const unsigned y = (unsigned)(((unsigned long long)r * magic_w) >> 40);
      e.y[lane] = (int)y;
      e.x[lane] = (int)(r - y * (unsigned)W);
      e.lp[lane] = (int)(pl - plane0);
      e.chw[lane] = (int)((pl - (unsigned)(((unsigned long long)pl * magic_c) >> 40) * C) * hw + r);
      li += T;
      plane += q_T;
      rem += r_T;
      if (rem >= hw) {
        rem -= hw;
        ++plane;
      }
    }
    if (!any) continue;
    float va[4], vb[4];
    term_values4<KIND_A>(p.a, sa, a4, la4, e, H, W, inv_div_a, va);
    if (KIND_B != SONAR_TERM_NONE) term_values4<KIND_B>(p.b, sb, b4, lb4, e, H, W, inv_div_b, vb);
#pragma unroll
    for (int lane = 0; lane < 4; ++lane) {
      const float v = KIND_B != SONAR_TERM_NONE ? blend<float>(p.blend_mode, va[lane], vb[lane], p.blend_t) : va[lane];
      if (ok[lane]) {
        out[li_call + (uint64_t)T * lane] = v;
        ms += v;
        mss += v * v;
      }
    }
  }
  commit_moments(p.sums, p.sums_clear, ms, mss);
}

static bool term_ok(const SonarMixTerm& t) {
  switch (t.kind) {
    case SONAR_TERM_NONE: return true;
    case SONAR_TERM_PYRAMID: {
      if (t.n_levels < 0 || t.n_levels > SONAR_PYRAMID_MAX_LEVELS) return false;
      if (t.mode != SONAR_RESAMPLE_BILINEAR && t.mode != SONAR_RESAMPLE_NEAREST_EXACT) return false;
      int full = 0;
      for (int l = 0; l < t.n_levels; ++l) {
        if (t.levels[l] == nullptr) {
          ++full;
          if (t.full_level != l) return false;
        } else if (t.level_h[l] < 1 || t.level_w[l] < 1 || t.level_h[l] > 65535 || t.level_w[l] > 65535) {
          return false;
        }
      }
      return full <= 1 && (full == 1 || t.full_level < 0);
    }
    case SONAR_TERM_PERLIN:
      if (t.iterations < 0 || t.iterations > SONAR_MIX_MAX_TABLES || t.div_fac == 0.0f) return false;
      for (int i = 0; i < t.iterations; ++i)
        if (t.tables[i] == nullptr) return false;
      return true;
    default: return false;
  }
}

template <int KIND_A, bool FULL>
static void launch_mix_b(const MixLaunch& L, int grid, size_t smem, cudaStream_t stream) {
  switch (L.p.b.kind) {
    case SONAR_TERM_PYRAMID: noise_mix_kernel<KIND_A, SONAR_TERM_PYRAMID, FULL><<<grid, kBlock, smem, stream>>>(L); break;
    case SONAR_TERM_PERLIN: noise_mix_kernel<KIND_A, SONAR_TERM_PERLIN, FULL><<<grid, kBlock, smem, stream>>>(L); break;
    default: noise_mix_kernel<KIND_A, SONAR_TERM_NONE, FULL><<<grid, kBlock, smem, stream>>>(L); break;
  }
}

template <bool FULL>
static void launch_mix_a(const MixLaunch& L, int grid, size_t smem, cudaStream_t stream) {
  if (L.p.a.kind == SONAR_TERM_PYRAMID)
    launch_mix_b<SONAR_TERM_PYRAMID, FULL>(L, grid, smem, stream);
  else
    launch_mix_b<SONAR_TERM_PERLIN, FULL>(L, grid, smem, stream);
}

}  // namespace sonar

extern "C" {

int sonar_perlin_tables_f32(float* const* tables_host, const uint64_t* offsets_host, const uint32_t* grid_blocks_host,
                            int32_t n_tables, uint64_t seed, int32_t C, int32_t H, int32_t W, int32_t blend_mode, void* stream) {
  using namespace sonar;
  if (n_tables <= 0) return 0;
  if (n_tables > SONAR_MIX_MAX_TABLES || C <= 0 || H <= 0 || W <= 0 || tables_host == nullptr || offsets_host == nullptr ||
      grid_blocks_host == nullptr)
    return (int)cudaErrorInvalidValue;
  PerlinTablesLaunch L;
  for (int i = 0; i < SONAR_MIX_MAX_TABLES; ++i) {
    const bool on = i < n_tables;
    L.table[i] = on ? tables_host[i] : nullptr;
    L.offset[i] = on ? offsets_host[i] : 0;
    L.grid_blocks[i] = on ? grid_blocks_host[i] : 1;
    if (on && (L.table[i] == nullptr || L.grid_blocks[i] == 0)) return (int)cudaErrorInvalidValue;
  }
  L.seed = seed;
  L.C = C;
  L.H = H;
  L.W = W;
  L.blend_mode = blend_mode;
  const int64_t tiles = (int64_t)((H + kPerlinTile - 1) / kPerlinTile) * ((W + kPerlinTile - 1) / kPerlinTile);
  if (tiles > 0x7fffffffll || C > 65535) return (int)cudaErrorInvalidValue;
  static_assert(kPerlinTile * kPerlinTile == kBlock, "one thread per pixel of the tile");
  perlin_tables_kernel<<<dim3((unsigned)tiles, (unsigned)C, (unsigned)n_tables), kBlock, 0, (cudaStream_t)stream>>>(L);
  SONAR_LAUNCH_CHECK();
  return 0;
}

int sonar_noise_mix_f32(const SonarNoiseMixParams* params, void* stream_) {
  using namespace sonar;
  if (params == nullptr) return (int)cudaErrorInvalidValue;
  MixLaunch L;
  L.p = *params;
  const SonarNoiseMixParams& p = L.p;
  if (p.n <= 0) return 0;
  const int64_t hw = (int64_t)p.H * p.W;
  if (p.out == nullptr || p.H <= 0 || p.W <= 0 || p.C <= 0 || p.grid_blocks == 0 || p.begin < 0 ||
      p.begin + p.n > p.numel_total || p.numel_total >= (1ll << 32) || hw >= (1ll << 31) || p.W < 2 || p.C < 1 || p.begin % hw != 0 || p.n % hw != 0 ||
      (p.a.kind != SONAR_TERM_PYRAMID && p.a.kind != SONAR_TERM_PERLIN) || !term_ok(p.a) || !term_ok(p.b))
    return (int)cudaErrorInvalidValue;
  // the element loop finds (y, x) and the channel with 24-bit magic divisions
  const int64_t planes_total = p.numel_total / hw;
  if (hw > (1 << 24) || planes_total >= (1 << 24) || p.W >= 65536 || p.C >= 65536) return (int)cudaErrorInvalidValue;
  L.magic_w = ((1ull << 40) + (unsigned)p.W - 1) / (unsigned)p.W;
  L.magic_c = ((1ull << 40) + (unsigned)p.C - 1) / (unsigned)p.C;
  const int64_t T = (int64_t)p.grid_blocks * kBlock, end = p.begin + p.n;
  L.k_lo = (uint32_t)((p.begin / T) / 4);
  L.k_hi = (uint32_t)(((end - 1) / T) / 4);
  size_t smem = 2 * SONAR_PYRAMID_MAX_LEVELS * sizeof(LevelDesc);
  const int64_t local_planes = p.n / hw;
  if ((int64_t)p.C * hw >= (1ll << 31)) return (int)cudaErrorInvalidValue;
  for (const SonarMixTerm* t : {&p.a, &p.b}) {
    if (t->kind != SONAR_TERM_PYRAMID) continue;
    int n_res = 0;
    for (int l = 0; l < t->n_levels; ++l) {
      if (t->levels[l] == nullptr) continue;
      ++n_res;
      if (local_planes * t->level_h[l] * (int64_t)t->level_w[l] >= (1ll << 31)) return (int)cudaErrorInvalidValue;
    }
    smem += (size_t)n_res * (p.W + p.H) * sizeof(MixTap);
  }
  if (smem > 48 * 1024) return (int)cudaErrorInvalidValue;  // (4 resampled levels of a 1024 x 1024 plane: 64 KB -- use the unfused path)
  L.q_T = (unsigned)(T / hw);
  L.r_T = (unsigned)(T % hw);
  const int grid = (int)p.grid_blocks;  // one CUDA thread per virtual ATen thread
  cudaStream_t stream = (cudaStream_t)stream_;
  if (p.begin == 0 && p.n == p.numel_total)
    launch_mix_a<true>(L, grid, smem, stream);
  else
    launch_mix_a<false>(L, grid, smem, stream);
  SONAR_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
