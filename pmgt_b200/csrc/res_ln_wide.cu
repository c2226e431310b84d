// y = LayerNorm(dropout(o) + res) and its backward for hidden sizes that are multiples of 256 (BASELINE config 5:
// H = 768) -- BertSelfOutput / BertOutput after their dense layer (called at pmgt/pmgt/modeling_pmgt.py:371,324).
// Same contract as res_ln_fwd_kernel / res_ln_bwd_kernel of rowwise.cu (which keep serving every other width); those
// ran at ~1.4 TB/s on H = 768: 8-byte accesses, a dropout hash evaluated per element, and three shared-memory
// read-modify-writes per four columns and row for the parameter gradients.  Here
//   * a warp owns a row, a lane owns G8 groups of EIGHT consecutive columns: 16-byte loads / stores, and one block of
//     the dropout stream (common.cuh) is exactly one lane's group;
//   * the raw loads of the NEXT row are issued before the current row is processed (a warp keeps ~9 KB in flight);
//   * d_gamma, d_beta and d_bias accumulate in REGISTERS over all rows of the lane (72 accumulators at H = 768), are
//     combined across the CTA's warps in shared memory at the end, and leave as one atomic per column and CTA.
#include "common.cuh"

namespace pmgt {

namespace {

constexpr int kWideThreads = 256;

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  unpack_bf16x2(u.x, f[0], f[1]); unpack_bf16x2(u.y, f[2], f[3]);
  unpack_bf16x2(u.z, f[4], f[5]); unpack_bf16x2(u.w, f[6], f[7]);
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 u;
  u.x = pack_bf16x2(f[0], f[1]); u.y = pack_bf16x2(f[2], f[3]);
  u.z = pack_bf16x2(f[4], f[5]); u.w = pack_bf16x2(f[6], f[7]);
  return u;
}
__device__ __forceinline__ uint4 ldg16(const uint16_t* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void ld8f32(const float* __restrict__ p, float (&f)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  const float4 b = __ldg(reinterpret_cast<const float4*>(p + 4));
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

// dropout multipliers (1 / (1 - p) or 0) of the 8 elements idx8 .. idx8 + 7 and their keep bits
__device__ __forceinline__ uint32_t drop8(uint64_t seed, uint32_t site, uint64_t idx8, uint32_t thr, float ks, float (&f)[8]) {
  uint32_t w[4];
  dropout_words8(seed, site, idx8, w);
  uint32_t bits = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const bool lo = (w[i] & 0xffffu) >= thr, hi = (w[i] >> 16) >= thr;
    f[2 * i] = lo ? ks : 0.f;
    f[2 * i + 1] = hi ? ks : 0.f;
    bits |= (lo ? 1u : 0u) << (2 * i) | (hi ? 1u : 0u) << (2 * i + 1);
  }
  return bits;
}

template <int G8>
__global__ void __launch_bounds__(kWideThreads) res_ln_fwd_wide_kernel(const pmgt_resln_args a) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * (kWideThreads / 32) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (kWideThreads / 32);
  const int H = a.H;
  const bool drop = a.dropout_p > 0.f;
  const float ks = drop ? 1.f / (1.f - a.dropout_p) : 1.f;
  const uint32_t thr = dropout_threshold(a.dropout_p);
  const float inv_h = 1.f / (float)H;
  uint4 no[G8], nr[G8];
  if (warp0 < a.T) {
#pragma unroll
    for (int i = 0; i < G8; ++i) {
      no[i] = ldg16(a.o + warp0 * H + (lane + 32 * i) * 8);
      nr[i] = ldg16(a.res + warp0 * H + (lane + 32 * i) * 8);
    }
  }
  for (long long tok = warp0; tok < a.T; tok += nwarps) {
    float z[G8][8];
#pragma unroll
    for (int i = 0; i < G8; ++i) {
      float r[8];
      unpack8(no[i], z[i]);
      unpack8(nr[i], r);
      if (drop) {
        float f[8];
        drop8(a.dropout_seed, a.dropout_site, (uint64_t)tok * H + (uint64_t)((lane + 32 * i) * 8), thr, ks, f);
#pragma unroll
        for (int j = 0; j < 8; ++j) z[i][j] = fmaf(z[i][j], f[j], r[j]);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) z[i][j] += r[j];
      }
    }
    const long long nxt = tok + nwarps;
    if (nxt < a.T) {
#pragma unroll
      for (int i = 0; i < G8; ++i) {
        no[i] = ldg16(a.o + nxt * H + (lane + 32 * i) * 8);
        nr[i] = ldg16(a.res + nxt * H + (lane + 32 * i) * 8);
      }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < G8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) s += z[i][j];
    const float mean = warp_sum(s) * inv_h;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < G8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) { const float d = z[i][j] - mean; q = fmaf(d, d, q); }
    const float rstd = rsqrtf(warp_sum(q) * inv_h + a.ln_eps);
    const float nm = -mean * rstd;
#pragma unroll
    for (int i = 0; i < G8; ++i) {
      const int h = (lane + 32 * i) * 8;
      float g[8], b[8];
      ld8f32(a.ln_g + h, g);
      ld8f32(a.ln_b + h, b);
#pragma unroll
      for (int j = 0; j < 8; ++j) z[i][j] = fmaf(fmaf(z[i][j], rstd, nm), g[j], b[j]);
      *reinterpret_cast<uint4*>(a.y + tok * H + h) = pack8(z[i]);
      if (a.y_f32) {
        *reinterpret_cast<float4*>(a.y_f32 + tok * H + h) = make_float4(z[i][0], z[i][1], z[i][2], z[i][3]);
        *reinterpret_cast<float4*>(a.y_f32 + tok * H + h + 4) = make_float4(z[i][4], z[i][5], z[i][6], z[i][7]);
      }
    }
  }
}

template <int G8>
__global__ void __launch_bounds__(kWideThreads, 1) res_ln_bwd_wide_kernel(const pmgt_resln_args a) {
  extern __shared__ float red[];  // [3][H]: d_g | d_b | d_bias of the CTA
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * (kWideThreads / 32) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (kWideThreads / 32);
  const int H = a.H;
  const bool drop = a.dropout_p > 0.f;
  const float ks = drop ? 1.f / (1.f - a.dropout_p) : 1.f;
  const uint32_t thr = dropout_threshold(a.dropout_p);
  const float inv_h = 1.f / (float)H;
  const bool sep_do = a.d_o != nullptr && a.d_o != a.dz;
  for (int i = threadIdx.x; i < 3 * H; i += kWideThreads) red[i] = 0.f;
  __syncthreads();

  float acc_g[G8][8], acc_b[G8][8], acc_o[G8][8];
#pragma unroll
  for (int i = 0; i < G8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) { acc_g[i][j] = 0.f; acc_b[i][j] = 0.f; acc_o[i][j] = 0.f; }

  uint4 no[G8], nr[G8], nd[G8];
  auto fetch = [&](long long tok) {
#pragma unroll
    for (int i = 0; i < G8; ++i) {
      const long long off = tok * H + (lane + 32 * i) * 8;
      no[i] = ldg16(a.o + off);
      nr[i] = ldg16(a.res + off);
      nd[i] = a.dy ? ldg16(a.dy + off) : make_uint4(0u, 0u, 0u, 0u);
    }
  };
  if (warp0 < a.T) fetch(warp0);
  for (long long tok = warp0; tok < a.T; tok += nwarps) {
    float z[G8][8], dy[G8][8];
    uint32_t keep[G8];
#pragma unroll
    for (int i = 0; i < G8; ++i) {
      float r[8];
      unpack8(no[i], z[i]);
      unpack8(nr[i], r);
      unpack8(nd[i], dy[i]);
      keep[i] = 0xffu;
      if (drop) {
        float f[8];
        keep[i] = drop8(a.dropout_seed, a.dropout_site, (uint64_t)tok * H + (uint64_t)((lane + 32 * i) * 8), thr, ks, f);
#pragma unroll
        for (int j = 0; j < 8; ++j) z[i][j] = fmaf(z[i][j], f[j], r[j]);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) z[i][j] += r[j];
      }
      if (a.dy_f32) {
        float e[8];
        ld8f32(a.dy_f32 + tok * H + (lane + 32 * i) * 8, e);
#pragma unroll
        for (int j = 0; j < 8; ++j) dy[i][j] += e[j];
      }
    }
    const long long nxt = tok + nwarps;
    if (nxt < a.T) fetch(nxt);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < G8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) s += z[i][j];
    const float mean = warp_sum(s) * inv_h;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < G8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) { const float d = z[i][j] - mean; q = fmaf(d, d, q); }
    const float rstd = rsqrtf(warp_sum(q) * inv_h + a.ln_eps);
    const float nm = -mean * rstd;
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < G8; ++i) {
      float g[8];
      ld8f32(a.ln_g + (lane + 32 * i) * 8, g);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = dy[i][j];
        const float xh = fmaf(z[i][j], rstd, nm);
        acc_g[i][j] = fmaf(d, xh, acc_g[i][j]);
        acc_b[i][j] += d;
        const float dg = d * g[j];
        s1 += dg;
        s2 = fmaf(dg, xh, s2);
        dy[i][j] = dg;
        z[i][j] = xh;
      }
    }
    s1 = warp_sum(s1) * inv_h;
    s2 = warp_sum(s2) * inv_h;
#pragma unroll
    for (int i = 0; i < G8; ++i) {
      const int h = (lane + 32 * i) * 8;
      float dz[8], dov[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        dz[j] = rstd * (dy[i][j] - s1 - z[i][j] * s2);
        dov[j] = (keep[i] >> j) & 1u ? dz[j] * ks : 0.f;
      }
      const uint4 pdz = pack8(dz);
      *reinterpret_cast<uint4*>(a.dz + tok * H + h) = pdz;
      uint4 pdo = pdz;
      if (sep_do) {
        pdo = pack8(dov);
        *reinterpret_cast<uint4*>(a.d_o + tok * H + h) = pdo;
      }
      float rb[8];  // d_bias = column sums of the STORED (bf16-rounded) d_o
      unpack8(pdo, rb);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc_o[i][j] += rb[j];
    }
  }
  // CTA-level reduction in shared memory, then one atomic per column and CTA
#pragma unroll
  for (int i = 0; i < G8; ++i) {
    const int h = (lane + 32 * i) * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      atomicAdd(red + h + j, acc_g[i][j]);
      atomicAdd(red + H + h + j, acc_b[i][j]);
      atomicAdd(red + 2 * H + h + j, acc_o[i][j]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < H; i += kWideThreads) {
    if (a.d_g && red[i] != 0.f) atomicAdd(a.d_g + i, red[i]);
    if (a.d_b && red[H + i] != 0.f) atomicAdd(a.d_b + i, red[H + i]);
    if (a.d_bias && red[2 * H + i] != 0.f) atomicAdd(a.d_bias + i, red[2 * H + i]);
  }
}

bool wide_ok(const pmgt_resln_args* a) {
  if (a->H % 256 != 0 || a->H > 1024) return false;
  uintptr_t al = (uintptr_t)a->o | (uintptr_t)a->res;
  if (a->y) al |= (uintptr_t)a->y;
  if (a->y_f32) al |= (uintptr_t)a->y_f32;
  if (a->dy) al |= (uintptr_t)a->dy;
  if (a->dy_f32) al |= (uintptr_t)a->dy_f32;
  if (a->dz) al |= (uintptr_t)a->dz;
  if (a->d_o) al |= (uintptr_t)a->d_o;
  return (al & 15) == 0;
}

int wide_grid(long long T, int ctas_per_sm) {
  long long need = (T + kWideThreads / 32 - 1) / (kWideThreads / 32);
  const long long cap = (long long)num_sms() * ctas_per_sm;
  if (need > cap) need = cap;
  return (int)(need < 1 ? 1 : need);
}

}  // namespace

// return 1 when handled here, 0 when the caller must use the generic kernels, < 0 on error
int res_ln_fwd_wide(const pmgt_resln_args* a, cudaStream_t st) {
  if (!wide_ok(a)) return 0;
  const int grid = wide_grid(a->T, 4);
  switch (a->H / 256) {
    case 1: res_ln_fwd_wide_kernel<1><<<grid, kWideThreads, 0, st>>>(*a); break;
    case 2: res_ln_fwd_wide_kernel<2><<<grid, kWideThreads, 0, st>>>(*a); break;
    case 3: res_ln_fwd_wide_kernel<3><<<grid, kWideThreads, 0, st>>>(*a); break;
    default: res_ln_fwd_wide_kernel<4><<<grid, kWideThreads, 0, st>>>(*a); break;
  }
  PMGT_LAUNCH_CHECK();
  return 1;
}

int res_ln_bwd_wide(const pmgt_resln_args* a, cudaStream_t st) {
  if (!wide_ok(a)) return 0;
  const int grid = wide_grid(a->T, 1);
  const size_t smem = (size_t)3 * a->H * sizeof(float);
  switch (a->H / 256) {
    case 1: res_ln_bwd_wide_kernel<1><<<grid, kWideThreads, smem, st>>>(*a); break;
    case 2: res_ln_bwd_wide_kernel<2><<<grid, kWideThreads, smem, st>>>(*a); break;
    case 3: res_ln_bwd_wide_kernel<3><<<grid, kWideThreads, smem, st>>>(*a); break;
    default: res_ln_bwd_wide_kernel<4><<<grid, kWideThreads, smem, st>>>(*a); break;
  }
  PMGT_LAUNCH_CHECK();
  return 1;
}

}  // namespace pmgt
