// Channel-correlation mixer of the power-noise family: out[b, c, p] = sum_k M[c, k] * in[b, k, p].
//
// Reference: ChannelMixer.apply py/nodes/powernoise.py:94-101 --
//   noise = mixer @ noise.swapaxes(0, 1).reshape(c, -1); noise.reshape(c, b, h, w).swapaxes(1, 0)
// i.e. one (C x C) @ (C x B*H*W) fp32 product per sample, bracketed by two layout changes. Here the layout
// never changes: every batch item is a (C x HW) row-major matrix already, M is shared by all of them, and
// the kernel also reduces {sum, sum^2} of what it writes (the scale_noise that always follows needs no
// separate read pass).
//
// Two shapes of work:
//  * image latents (C <= 8; SD/SDXL 4, 16-channel models take the tiled path): a per-pixel C x C
//    mat-vec. One thread owns 4 consecutive pixels of one batch item, holds the C input float4 in
//    registers and writes C output float4: exactly one read and one write of the tensor, float4 both ways.
//  * video latents folded by frames_to_channels (C5: C = 16 x 33 = 528): a real GEMM, 64 x 128 output
//    tiles, K in steps of 16 through a ring of shared-memory stages filled by bulk-async copies (cp.async.bulk +
//    mbarrier, a dedicated producer warp), 4 x 8 outputs per thread. It stays on the fp32 FMA
//    pipe on purpose: the north-star tolerance is 1e-5 and the tensor cores have no fp32 mode (TF32
//    keeps 10 mantissa bits); the mixer is the identity (skipped) unless common_mode != 0.
#include "common.cuh"
#include "../../include/sonar_b200.h"

namespace sonar {

struct SmallMixer {
  float m[SONAR_MIXER_SMALL_MAX * SONAR_MIXER_SMALL_MAX];
};

template <int C, int VEC>
__global__ void __launch_bounds__(kBlock)
channel_mix_small_kernel(const float* __restrict__ in, float* __restrict__ out, SmallMixer mx, int64_t batch, int64_t hw,
                         double* sums, double* sums_clear) {
  const int64_t groups = VEC == 4 ? hw >> 2 : hw;  // pixel groups per plane
  const int64_t total = batch * groups;
  float s = 0.0f, ss = 0.0f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / groups, g = i - b * groups;
    const float* src = in + b * C * hw + g * VEC;
    float* dst = out + b * C * hw + g * VEC;
    float v[C][VEC];
#pragma unroll
    for (int k = 0; k < C; ++k) {
      if constexpr (VEC == 4) {
        const float4 t = ld4_stream(src + k * hw);
        v[k][0] = t.x; v[k][1] = t.y; v[k][2] = t.z; v[k][3] = t.w;
      } else {
        v[k][0] = __ldg(src + k * hw);
      }
    }
#pragma unroll
    for (int c = 0; c < C; ++c) {
      float acc[VEC];
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        acc[e] = mx.m[c * C] * v[0][e];
#pragma unroll
        for (int k = 1; k < C; ++k) acc[e] = fmaf(mx.m[c * C + k], v[k][e], acc[e]);
        s += acc[e];
        ss += acc[e] * acc[e];
      }
      if constexpr (VEC == 4)
        st4(dst + c * hw, make_float4(acc[0], acc[1], acc[2], acc[3]));
      else
        dst[c * hw] = acc[0];
    }
  }
  commit_moments(sums, sums_clear, s, ss);
}

constexpr int kMixBM = 64, kMixBN = 128, kMixBK = 16;

// grid: x = n tile (fastest: neighbouring CTAs share the A tile in L2), then m tile, then batch item
__global__ void __launch_bounds__(kBlock)
channel_mix_tiled_kernel(const float* __restrict__ in, float* __restrict__ out, const float* __restrict__ mixer, int C,
                         int64_t hw, int tiles_n, int tiles_m, int vec_ok, double* sums, double* sums_clear) {
  __shared__ float As[kMixBK][kMixBM + 4];  // [k][m]: M[m0 + m][k0 + k]
  __shared__ float Bs[kMixBK][kMixBN];      // [k][n]: in[b][k0 + k][n0 + n]
  const int tile = blockIdx.x;
  const int tn = tile % tiles_n, tm = (tile / tiles_n) % tiles_m;
  const int64_t b = tile / (tiles_n * tiles_m);
  const int m0 = tm * kMixBM;
  const int64_t n0 = (int64_t)tn * kMixBN;
  const float* inb = in + b * C * hw;
  float* outb = out + b * C * hw;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // 16 x 16 threads: 8 columns x 4 rows each
  float acc[4][8];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;
  const bool n_vec = vec_ok && (hw & 3) == 0 && n0 + kMixBN <= hw;  // float4 rows: 16-byte aligned tensors only
  for (int k0 = 0; k0 < C; k0 += kMixBK) {
    // A tile: 64 x 16 = 1024 elements, 4 per thread, read along k (contiguous in M's rows)
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int e = threadIdx.x + r * kBlock;
      const int m = e >> 4, k = e & 15;
      As[k][m] = (m0 + m < C && k0 + k < C) ? __ldg(mixer + (int64_t)(m0 + m) * C + k0 + k) : 0.0f;
    }
    // B tile: 16 x 128 = 2048 elements, 8 per thread as two float4 along n
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int e = threadIdx.x + r * kBlock;  // float4 index
      const int k = e >> 5, n4 = (e & 31) << 2;
      float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k0 + k < C) {
        const float* p = inb + (int64_t)(k0 + k) * hw + n0 + n4;
        if (n_vec) {
          t = ld4(p);
        } else {
          if (n0 + n4 + 0 < hw) t.x = p[0];
          if (n0 + n4 + 1 < hw) t.y = p[1];
          if (n0 + n4 + 2 < hw) t.z = p[2];
          if (n0 + n4 + 3 < hw) t.w = p[3];
        }
      }
      *reinterpret_cast<float4*>(&Bs[k][n4]) = t;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kMixBK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[k][64 + tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  float s = 0.0f, ss = 0.0f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= C) continue;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int64_t n = n0 + half * 64 + tx * 4;
      float* p = outb + (int64_t)m * hw + n;
      const float* a = &acc[i][half * 4];
      if (n_vec) {
        st4(p, make_float4(a[0], a[1], a[2], a[3]));
        s += (a[0] + a[1]) + (a[2] + a[3]);
        ss += (a[0] * a[0] + a[1] * a[1]) + (a[2] * a[2] + a[3] * a[3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (n + j < hw) {
            p[j] = a[j];
            s += a[j];
            ss += a[j] * a[j];
          }
      }
    }
  }
  commit_moments(sums, sums_clear, s, ss);
}

// ---------------------------------------------------------------------------------------------
// Bulk-async pipelined variant of the tiled GEMM (the default for 16-byte-aligned operands): a producer warp streams
// the A (64 x 16) and B (16 x 128) tiles of the next k-blocks into a ring of shared-memory stages with
// cp.async.bulk (the TMA engine's 1-D bulk copy: one instruction per tile row, no registers, no address math in
// the consumers), completion is counted in bytes on an mbarrier per stage (expect-tx), and eight consumer warps
// do nothing but shared-memory loads and FMAs; a second mbarrier per stage hands the slot back to the producer.
// ---------------------------------------------------------------------------------------------
constexpr int kMixStages = 3;
constexpr int kMixAPitch = kMixBK + 4;  // floats per A row in shared memory: 80 bytes (16-byte multiple, spreads the banks)
constexpr int kMixConsumers = 256;
constexpr int kMixBulkThreads = kMixConsumers + 32;

struct MixStageBuf {
  float A[kMixBM][kMixAPitch];  // [m][k]: M[m0 + m][k0 + k]
  float B[kMixBK][kMixBN];      // [k][n]: in[b][k0 + k][n0 + n]
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// 1-D bulk copy global -> shared, completion signalled in bytes on `bar` (size and both addresses: multiples of 16)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// `packed` = the mixer pre-tiled by sonar_channel_mix_pack_bytes' layout: [tiles_m][k-blocks][64][kMixAPitch], zero
// padded, so an A tile is ONE 5 KB bulk copy (64 separate 64-byte row copies made the kernel producer-bound: the
// copy engine retires a descriptor every ~15 cycles) and ragged edges need no special case.
__global__ void __launch_bounds__(kMixBulkThreads)
channel_mix_bulk_kernel(const float* __restrict__ in, float* __restrict__ out, const float* __restrict__ packed, int C,
                        int64_t hw, int tiles_n, int tiles_m, double* sums, double* sums_clear) {
  __shared__ __align__(128) MixStageBuf stage[kMixStages];
  __shared__ __align__(8) uint64_t full_bar[kMixStages], empty_bar[kMixStages];
  const int tile = blockIdx.x;
  const int tn = tile % tiles_n, tm = (tile / tiles_n) % tiles_m;
  const int64_t b = tile / (tiles_n * tiles_m);
  const int m0 = tm * kMixBM;
  const int64_t n0 = (int64_t)tn * kMixBN;
  const float* inb = in + b * C * hw;
  float* outb = out + b * C * hw;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kMixStages; ++s) {
      mbar_init(&full_bar[s], 1);                   // the producer's arrive.expect_tx; the copies complete the bytes
      mbar_init(&empty_bar[s], kMixConsumers / 32);  // one arrival per consumer warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int n_kblocks = (C + kMixBK - 1) / kMixBK;
  constexpr uint32_t kBytesA = kMixBM * kMixAPitch * sizeof(float);
  const float* a_tiles = packed + (int64_t)tm * n_kblocks * (kMixBM * kMixAPitch);
  const uint32_t bytes_b_row = (uint32_t)((hw - n0 < kMixBN ? hw - n0 : kMixBN) * sizeof(float));
  float s = 0.0f, ss = 0.0f;
  if (warp == kMixConsumers / 32) {
    // ---------------- producer warp ----------------
    for (int kb = 0; kb < n_kblocks; ++kb) {
      const int st = kb % kMixStages;
      const uint32_t phase = (uint32_t)(kb / kMixStages) & 1u;
      mbar_wait(&empty_bar[st], phase ^ 1u);  // (a fresh barrier passes the first round)
      const int k0 = kb * kMixBK;
      const int kk = C - k0 < kMixBK ? C - k0 : kMixBK;  // rows of B that exist (the A tile is zero beyond them)
      if (lane == 0) {
        mbar_arrive_expect_tx(&full_bar[st], kBytesA + (uint32_t)kk * bytes_b_row);
        bulk_g2s(&stage[st].A[0][0], a_tiles + (int64_t)kb * (kMixBM * kMixAPitch), kBytesA, &full_bar[st]);
      }
      __syncwarp();
      if (lane < kk) bulk_g2s(&stage[st].B[lane][0], inb + (int64_t)(k0 + lane) * hw + n0, bytes_b_row, &full_bar[st]);
    }
  } else {
    // ---------------- consumer warps: 16 x 16 threads, 4 rows x 8 columns each ----------------
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    float acc[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;
    for (int kb = 0; kb < n_kblocks; ++kb) {
      const int st = kb % kMixStages;
      const uint32_t phase = (uint32_t)(kb / kMixStages) & 1u;
      mbar_wait(&full_bar[st], phase);
      const int k0 = kb * kMixBK;
      const MixStageBuf& sb = stage[st];
      if (C - k0 >= kMixBK) {
#pragma unroll
        for (int k4 = 0; k4 < kMixBK; k4 += 4) {
          float4 a4[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) a4[i] = *reinterpret_cast<const float4*>(&sb.A[ty * 4 + i][k4]);
#pragma unroll
          for (int kq = 0; kq < 4; ++kq) {
            const float4 b0 = *reinterpret_cast<const float4*>(&sb.B[k4 + kq][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4*>(&sb.B[k4 + kq][64 + tx * 4]);
            const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float av = kq == 0 ? a4[i].x : kq == 1 ? a4[i].y : kq == 2 ? a4[i].z : a4[i].w;
#pragma unroll
              for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av, bv[j], acc[i][j]);
            }
          }
        }
      } else {  // ragged last k-block: only the first C - k0 columns / rows of the tiles were copied
        for (int k = 0; k < C - k0; ++k) {
          const float4 b0 = *reinterpret_cast<const float4*>(&sb.B[k][tx * 4]);
          const float4 b1 = *reinterpret_cast<const float4*>(&sb.B[k][64 + tx * 4]);
          const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float av = sb.A[ty * 4 + i][k];
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av, bv[j], acc[i][j]);
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[st]);  // this warp is done with the slot
    }
    const bool n_vec = n0 + kMixBN <= hw;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int m = m0 + ty * 4 + i;
      if (m >= C) continue;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int64_t n = n0 + half * 64 + tx * 4;
        float* p = outb + (int64_t)m * hw + n;
        const float* a = &acc[i][half * 4];
        if (n_vec) {
          st4(p, make_float4(a[0], a[1], a[2], a[3]));
          s += (a[0] + a[1]) + (a[2] + a[3]);
          ss += (a[0] * a[0] + a[1] * a[1]) + (a[2] * a[2] + a[3] * a[3]);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (n + j < hw) {
              p[j] = a[j];
              s += a[j];
              ss += a[j] * a[j];
            }
        }
      }
    }
  }
  commit_moments(sums, sums_clear, s, ss);
}

template <int C>
static void launch_small(const float* in, float* out, const float* mixer_host, int64_t batch, int64_t hw, double* sums,
                         double* sums_clear, cudaStream_t stream) {
  SmallMixer mx;
  for (int i = 0; i < C * C; ++i) mx.m[i] = mixer_host[i];
  const bool vec = (hw & 3) == 0 && aligned16(in) && aligned16(out);
  const int64_t items = batch * (vec ? hw >> 2 : hw);
  const int grid = streaming_grid(items, kBlock, 2);
  if (vec)
    channel_mix_small_kernel<C, 4><<<grid, kBlock, 0, stream>>>(in, out, mx, batch, hw, sums, sums_clear);
  else
    channel_mix_small_kernel<C, 1><<<grid, kBlock, 0, stream>>>(in, out, mx, batch, hw, sums, sums_clear);
}

}  // namespace sonar

extern "C" int64_t sonar_channel_mix_packed_floats(int32_t channels) {
  using namespace sonar;
  if (channels <= 0) return 0;
  const int64_t tiles_m = (channels + kMixBM - 1) / kMixBM, kblocks = (channels + kMixBK - 1) / kMixBK;
  return tiles_m * kblocks * kMixBM * kMixAPitch;
}

extern "C" int sonar_channel_mix_f32(const float* in, float* out, const float* mixer, const float* mixer_host,
                                     const float* mixer_packed, int64_t batch, int32_t channels, int64_t hw, double* sums,
                                     double* sums_clear, void* stream_) {
  using namespace sonar;
  if (batch <= 0 || hw <= 0 || channels <= 0) return 0;
  if (in == nullptr || out == nullptr || in == out || mixer == nullptr) return (int)cudaErrorInvalidValue;
  cudaStream_t stream = (cudaStream_t)stream_;
  if (channels <= SONAR_MIXER_SMALL_MAX && mixer_host != nullptr) {
    switch (channels) {
      case 1: launch_small<1>(in, out, mixer_host, batch, hw, sums, sums_clear, stream); break;
      case 2: launch_small<2>(in, out, mixer_host, batch, hw, sums, sums_clear, stream); break;
      case 3: launch_small<3>(in, out, mixer_host, batch, hw, sums, sums_clear, stream); break;
      case 4: launch_small<4>(in, out, mixer_host, batch, hw, sums, sums_clear, stream); break;
      case 5: launch_small<5>(in, out, mixer_host, batch, hw, sums, sums_clear, stream); break;
      case 6: launch_small<6>(in, out, mixer_host, batch, hw, sums, sums_clear, stream); break;
      case 7: launch_small<7>(in, out, mixer_host, batch, hw, sums, sums_clear, stream); break;
      default: launch_small<8>(in, out, mixer_host, batch, hw, sums, sums_clear, stream); break;
    }
    SONAR_LAUNCH_CHECK();
    return 0;
  }
  const int tiles_m = (channels + kMixBM - 1) / kMixBM;
  const int64_t tiles_n = (hw + kMixBN - 1) / kMixBN;
  const int64_t grid = tiles_n * tiles_m * batch;
  if (grid > 0x7fffffffll || tiles_n > 0x7fffffffll) return (int)cudaErrorInvalidValue;
  // rows of A and B start on 16-byte boundaries: the bulk-async pipeline; anything else: the plain tiled kernel
  // (below 32 channels most of a 64-row tile is padding and the plain kernel is as fast)
  const bool bulk_ok = mixer_packed != nullptr && channels >= 32 && (hw & 3) == 0 && aligned16(in) && aligned16(mixer_packed) && aligned16(out);
  if (bulk_ok)
    channel_mix_bulk_kernel<<<(unsigned)grid, kMixBulkThreads, 0, stream>>>(in, out, mixer_packed, channels, hw, (int)tiles_n,
                                                                           tiles_m, sums, sums_clear);
  else
    channel_mix_tiled_kernel<<<(unsigned)grid, kBlock, 0, stream>>>(in, out, mixer, channels, hw, (int)tiles_n, tiles_m,
                                                                   aligned16(in) && aligned16(out) ? 1 : 0, sums, sums_clear);
  SONAR_LAUNCH_CHECK();
  return 0;
}
