// LayerNorm backward for H = 128 as a bulk-copy-staged streaming kernel.
//
// The math is that of ln_bwd128_kernel (rowwise.cu): dy = dy_a + dy_b (+ dy_f32), z -> xhat, dz, d_o = dropout mask
// re-applied, d_gamma / d_beta column sums.  What changes is how the bytes move: the register-only version keeps two
// rows per half-warp in flight, which with 115 registers / thread caps an SM at ~49 KB of outstanding loads -- ncu
// showed it latency-bound at 46 % of DRAM peak with 24 % of the warp slots active.  Here ONE producer thread streams
// 64-row chunks of every input (contiguous 16 KB / 32 KB pieces, cp.async.bulk + mbarrier complete_tx) through a
// 4-stage shared-memory ring, so ~150-190 KB per SM are in flight regardless of what the compute warps are doing;
// 16 compute warps (a half-warp per row, 16-byte conflict-free smem reads) drain the ring and release a stage as soon
// as its rows are in registers.  Outputs are 256-byte-contiguous row stores.
#include "umma.cuh"

namespace pmgt {

namespace {

constexpr int kRows = 64;                 // rows per chunk
constexpr int kStages = 4;
constexpr int kComputeWarps = 16;
constexpr int kThreadsLn = 32 + 32 * kComputeWarps;
constexpr int kRowBytes = 256;            // 128 bf16

struct LnBars {
  uint64_t full[kStages], empty[kStages];
};

__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ float hsum(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ void unpack8s(const uint4& u, float* f) {
  unpack_bf16x2(u.x, f[0], f[1]); unpack_bf16x2(u.y, f[2], f[3]);
  unpack_bf16x2(u.z, f[4], f[5]); unpack_bf16x2(u.w, f[6], f[7]);
}
__device__ __forceinline__ uint4 pack8s(const float* f) {
  uint4 u;
  u.x = pack_bf16x2(f[0], f[1]); u.y = pack_bf16x2(f[2], f[3]);
  u.z = pack_bf16x2(f[4], f[5]); u.w = pack_bf16x2(f[6], f[7]);
  return u;
}

// NB = number of bf16 gradient inputs (1: dy_a, 2: dy_a + dy_b); F32 = an fp32 gradient input instead (last layer)
template <int NB, bool F32>
__global__ void __launch_bounds__(kThreadsLn, 1) ln_bwd_stream_kernel(const pmgt_lnbwd_args a, const uint16_t* g0,
                                                                      const uint16_t* g1, const int reverse) {
  constexpr int kZ = kRows * kRowBytes;                                     // 16 KB
  constexpr int kStageBytes = kZ + (F32 ? 2 * kZ : NB * kZ);
  extern __shared__ __align__(128) unsigned char smem[];
  LnBars* bars = reinterpret_cast<LnBars*>(smem + kStages * kStageBytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_launch_dependents();
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&bars->full[s], 1);
      mbar_init(&bars->empty[s], kComputeWarps);
    }
    fence_barrier_init();
  }
  __syncthreads();
  pdl_wait();
  const long long T = a.T;
  const long long n_chunks = (T + kRows - 1) / kRows;

  float dgam[8], dbet[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) dgam[j] = dbet[j] = 0.f;
  const int c = (lane & 15) * 8;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t i = 0;
      for (long long lc = blockIdx.x; lc < n_chunks; lc += gridDim.x, ++i) {
        const long long ch = reverse ? n_chunks - 1 - lc : lc;  // alternating traversal order, see next_tile_order
        const int s = i % kStages;
        mbar_wait(&bars->empty[s], ((i / kStages) & 1u) ^ 1u);
        const long long row0 = ch * kRows;
        const uint32_t rows = (uint32_t)((T - row0) < kRows ? (T - row0) : kRows);
        const uint32_t zb = rows * kRowBytes;
        mbar_arrive_expect_tx(&bars->full[s], F32 ? 3u * zb : (uint32_t)(1 + NB) * zb);
        const uint32_t dst = smem_u32(smem + s * kStageBytes);
        bulk_load(dst, a.z + row0 * 128, zb, &bars->full[s]);
        if (F32) {
          bulk_load(dst + kZ, a.dy_f32 + row0 * 128, 2u * zb, &bars->full[s]);
        } else {
          bulk_load(dst + kZ, g0 + row0 * 128, zb, &bars->full[s]);
          if (NB == 2) bulk_load(dst + 2 * kZ, g1 + row0 * 128, zb, &bars->full[s]);
        }
      }
    }
  } else {
    const int hw = (threadIdx.x - 32) >> 4;  // 0..31: rows hw and hw + 32 of every chunk
    const float ks = a.dropout_p > 0.f ? 1.f / (1.f - a.dropout_p) : 1.f;
    const bool sep_do = a.d_o != nullptr && a.d_o != a.dz;
    float gam[8];
    {
      const float4 q0 = *reinterpret_cast<const float4*>(a.ln_g + c), q1 = *reinterpret_cast<const float4*>(a.ln_g + c + 4);
      gam[0] = q0.x; gam[1] = q0.y; gam[2] = q0.z; gam[3] = q0.w; gam[4] = q1.x; gam[5] = q1.y; gam[6] = q1.z; gam[7] = q1.w;
    }
    uint32_t i = 0;
    for (long long lc = blockIdx.x; lc < n_chunks; lc += gridDim.x, ++i) {
      const long long ch = reverse ? n_chunks - 1 - lc : lc;
      const int s = i % kStages;
      mbar_wait(&bars->full[s], (i / kStages) & 1u);
      const unsigned char* st = smem + s * kStageBytes;
      float z[2][8], dy[2][8];
      bool ok[2];
      long long tok[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int r = hw + 32 * u;
        tok[u] = ch * kRows + r;
        ok[u] = tok[u] < T;
        if (ok[u]) {  // rows past T were not copied: their smem bytes are stale
          unpack8s(*reinterpret_cast<const uint4*>(st + r * kRowBytes + c * 2), z[u]);
          if (F32) {
            const float4 f0 = *reinterpret_cast<const float4*>(st + kZ + r * 512 + c * 4);
            const float4 f1 = *reinterpret_cast<const float4*>(st + kZ + r * 512 + c * 4 + 16);
            dy[u][0] = f0.x; dy[u][1] = f0.y; dy[u][2] = f0.z; dy[u][3] = f0.w;
            dy[u][4] = f1.x; dy[u][5] = f1.y; dy[u][6] = f1.z; dy[u][7] = f1.w;
          } else {
            unpack8s(*reinterpret_cast<const uint4*>(st + kZ + r * kRowBytes + c * 2), dy[u]);
            if (NB == 2) {
              float t[8];
              unpack8s(*reinterpret_cast<const uint4*>(st + 2 * kZ + r * kRowBytes + c * 2), t);
#pragma unroll
              for (int j = 0; j < 8; ++j) dy[u][j] += t[j];
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) { z[u][j] = 0.f; dy[u][j] = 0.f; }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->empty[s]);  // this warp's rows are in registers
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) sum += z[u][j];
        const float mean = hsum(sum) * (1.f / 128.f);
        float q = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) { z[u][j] -= mean; q = fmaf(z[u][j], z[u][j], q); }
        const float rstd = rsqrtf(hsum(q) * (1.f / 128.f) + a.ln_eps);
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          z[u][j] *= rstd;  // xhat
          dgam[j] = fmaf(dy[u][j], z[u][j], dgam[j]);  // rows past T carry dy = 0
          dbet[j] += dy[u][j];
          dy[u][j] *= gam[j];
          s1 += dy[u][j];
          s2 = fmaf(dy[u][j], z[u][j], s2);
        }
        s1 = hsum(s1) * (1.f / 128.f);
        s2 = hsum(s2) * (1.f / 128.f);
        float dz[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) dz[j] = rstd * (dy[u][j] - s1 - z[u][j] * s2);
        if (ok[u]) {
          *reinterpret_cast<uint4*>(a.dz + tok[u] * 128 + c) = pack8s(dz);
          if (sep_do) {
            if (a.dropout_p > 0.f) {
              const uint32_t k8 = dropout_keep8(a.dropout_seed, a.dropout_site, (uint64_t)tok[u] * 128u + (uint64_t)c, a.dropout_p);
#pragma unroll
              for (int j = 0; j < 8; ++j) dz[j] = (k8 >> j) & 1u ? dz[j] * ks : 0.f;
            }
            *reinterpret_cast<uint4*>(a.d_o + tok[u] * 128 + c) = pack8s(dz);
          }
        }
      }
    }
  }
  // column sums: fold the 32 half-warps through the (now idle) ring, one atomic per column and CTA
  __syncthreads();
  float* red = reinterpret_cast<float*>(smem);  // [32][2][128]
  if (warp > 0) {
    const int hw = (threadIdx.x - 32) >> 4;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      red[(hw * 2 + 0) * 128 + c + j] = dgam[j];
      red[(hw * 2 + 1) * 128 + c + j] = dbet[j];
    }
  }
  __syncthreads();
  if (threadIdx.x < 256) {
    const int which = threadIdx.x >> 7, col = threadIdx.x & 127;
    float v = 0.f;
#pragma unroll 8
    for (int g = 0; g < 32; ++g) v += red[(g * 2 + which) * 128 + col];
    float* dst = which == 0 ? a.d_g : a.d_b;
    if (dst && v != 0.f) atomicAdd(dst + col, v);
  }
}

template <int NB, bool F32>
int launch(const pmgt_lnbwd_args* a, const uint16_t* g0, const uint16_t* g1, cudaStream_t st) {
  constexpr int kStageBytes = kRows * kRowBytes * (F32 ? 3 : 1 + NB);
  constexpr int smem = kStages * kStageBytes + (int)sizeof(LnBars);
  static_assert(smem >= 32 * 2 * 128 * 4, "the ring doubles as the column-sum scratch");
  auto kern = ln_bwd_stream_kernel<NB, F32>;
  static unsigned long long cfg = 0;
  if (first_use_on_device(cfg)) {
    PMGT_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  }
  long long chunks = (a->T + kRows - 1) / kRows;
  int grid = num_sms();
  if (grid > chunks) grid = (int)chunks;
  PMGT_CHECK_CUDA(launch_kernel(true, kern, dim3(grid), dim3(kThreadsLn), smem, st, *a, g0, g1, next_tile_order()));
  return PMGT_OK;
}

}  // namespace

// Called by pmgt_ln_bwd (rowwise.cu) after validation.  Returns 1 when this input combination is not covered
// (the register kernel handles it), else the launch status.
int ln_bwd_stream(const pmgt_lnbwd_args* a, cudaStream_t st) {
  const uintptr_t al = (uintptr_t)a->z | (uintptr_t)a->dy_a | (uintptr_t)a->dy_b | (uintptr_t)a->dy_f32 | (uintptr_t)a->dz |
                       (uintptr_t)a->d_o | (uintptr_t)a->ln_g;
  if (al & 15) return 1;
  if (a->dy_f32) {
    if (a->dy_a || a->dy_b) return 1;
    return launch<0, true>(a, nullptr, nullptr, st);
  }
  const uint16_t* g0 = a->dy_a ? a->dy_a : a->dy_b;
  const uint16_t* g1 = a->dy_a ? a->dy_b : nullptr;
  if (g1) return launch<2, false>(a, g0, g1, st);
  return launch<1, false>(a, g0, nullptr, st);
}

}  // namespace pmgt
