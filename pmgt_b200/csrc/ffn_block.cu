// Fused feed-forward block of one PMGT encoder layer (BertIntermediate + BertOutput as the reference composes them in
// PMGTLayer.feed_forward_chunk, pmgt/pmgt/modeling_pmgt.py:296-325), default widths H = I = 128:
//
//   forward   out = LayerNorm(dropout(gelu(a W1^T + b1) W2^T + b2) + a)          one kernel, a -> out
//   backward  d_a, dW1, db1, dW2, db2, d_gamma, d_beta from (a, d_out)           one kernel, (a, dy) -> d_a
//
// Both are persistent tcgen05 token-tile kernels (one CTA per SM, 128-token tiles, W1 and W2 resident in shared memory
// as swizzled images that serve as K-major operands of the forward products and as MN-major operands of the dX
// products).  Forward saves h = gelu(h_pre) and gelu'(h_pre) (both cost nothing extra to form: the erf and the
// Gaussian density share one exponential); the backward kernel re-forms the pre-LayerNorm sum z from h with one more
// 128^3 product on an otherwise idle tensor pipe, so neither h_pre nor z is ever stored and the backward epilogues do
// no transcendental work at all -- both kernels are bound by instruction issue, not by HBM.
//
// HBM rows (256-byte bf16 token rows) per token: forward 4 (was 7 as GELU + RES_LN token-tile kernels), backward 5-6
// (was 12 as LayerNorm backward + two fused dX+dW kernels).
#include "tile.cuh"

namespace pmgt {

constexpr int kFfnEpiWarps = 16;
constexpr int kFfnThreads = 96 + 32 * kFfnEpiWarps;   // producer | MMA | store | 16 epilogue warps

struct FfnParams {
  int T, num_tiles, reverse;
  const float* b1;
  const float* b2;
  const float* ln_g;
  const float* ln_b;
  float ln_eps, dropout_p;
  uint64_t seed;
  uint32_t site;
  float* out_f32;
  const uint16_t* a;      // the block input again: the residual rows are re-read in the accumulator layout
  long long ld_a;
  int save_act;           // forward: write h and gelu'(h_pre) for the backward kernel
  // backward
  const uint16_t* dy;     // gradient terms, read straight from global memory in the accumulator layout
  long long ld_dy;
  const uint16_t* dy_b;   // optional second term
  long long ld_dy_b;
  float *dw1, *dw2, *db1, *db2, *dg, *dbeta;
};

__device__ __forceinline__ int ffn_tile(const FfnParams& p, int lt) { return p.reverse ? p.num_tiles - 1 - lt : lt; }

__device__ __forceinline__ void tmem_alloc_512(uint32_t* slot) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(512u));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
}
__device__ __forceinline__ void tmem_dealloc_512(uint32_t base) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(512u));
}
__device__ __forceinline__ void load_image(uint32_t dst, const CUtensorMap* tm, uint64_t* bar, int col0, int row0) {
  tma_load_2d(dst, tm, bar, col0, row0);
  tma_load_2d(dst + kSlabBytes, tm, bar, col0 + 64, row0);
}
__device__ __forceinline__ void store_image(const CUtensorMap* tm, uint32_t src, int col0, int row0) {
  tma_store_2d(tm, src, col0, row0);
  tma_store_2d(tm, src + kSlabBytes, col0 + 64, row0);
}

// ---------------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------------
struct FfnFwdLayout {
  static constexpr int kW1 = 0, kW2 = kImgBytes;
  static constexpr int kA = 2 * kImgBytes;          // a(n): operand of the first product only (the residual is re-read from L2)
  static constexpr int kHH = 3 * kImgBytes;         // 2 stages: gelu output = operand of the second product + TMA store source
  static constexpr int kGP = 5 * kImgBytes;         // gelu'(h_pre): TMA store source
  static constexpr int kYS = 6 * kImgBytes;         // LayerNorm output: TMA store source
  static constexpr int kBar = 7 * kImgBytes;
  static constexpr int kTotal = kBar + 256 + 1024;
};

struct FfnFwdBars {
  uint64_t w_full;
  uint64_t a_full, a_empty;
  uint64_t sh_full[2], sh_empty[2];
  uint64_t hh_full[2], hh_empty[2];
  uint64_t gp_empty;
  uint64_t sz_full, sz_empty;
  uint64_t y_full, y_empty;
  uint32_t tmem_base;
};
static_assert(sizeof(FfnFwdBars) <= 256, "barrier block");

__device__ __forceinline__ void tmem_st_x2(uint32_t taddr, float a, float b) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(taddr), "r"(__float_as_uint(a)),
               "r"(__float_as_uint(b))
               : "memory");
}
__device__ __forceinline__ void tmem_ld_x8(uint32_t taddr, float (&r)[8]) {
  uint32_t u[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7])
               : "r"(taddr)
               : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) r[i] = __uint_as_float(u[i]);
}

// Row statistics of a 128-column row whose four 32-column quarters live in four different warps (the TMEM accumulator
// layout): Chan's combination of (mean, M2) partials of equal weight.
__device__ __forceinline__ void combine_stats(const float (&q)[8], float eps, float& mean, float& rstd) {
  mean = 0.25f * (q[0] + q[2] + q[4] + q[6]);
  const float d0 = q[0] - mean, d1 = q[2] - mean, d2 = q[4] - mean, d3 = q[6] - mean;
  const float m2 = q[1] + q[3] + q[5] + q[7] + 32.f * (d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3);
  rstd = rsqrtf(m2 * (1.f / 128.f) + eps);
}

// warp 0: TMA producer | warp 1: MMA issuer | warp 2: store warp | warps 3..18: epilogue (TMEM lane quarter x 32-column
// quarter per warp: one row x 32 columns per thread).
// Epilogue order is software-pipelined: gelu(n + 1) runs BEFORE layernorm(n), so the second product of tile n has the
// whole gelu epilogue of tile n + 1 to complete and neither epilogue waits for the tensor pipe.  LayerNorm stays in the
// accumulator layout: the four warps that share a row exchange (mean, M2) partials through 8 spare TMEM columns.
__global__ void __launch_bounds__(kFfnThreads, 1)
ffn_fwd_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_w1,
               const __grid_constant__ CUtensorMap tm_w2, const __grid_constant__ CUtensorMap tm_out,
               const __grid_constant__ CUtensorMap tm_h, const __grid_constant__ CUtensorMap tm_gp, const FfnParams p) {
  using Lay = FfnFwdLayout;
  extern __shared__ __align__(1024) unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  FfnFwdBars* bars = reinterpret_cast<FfnFwdBars*>(smem + Lay::kBar);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool save = p.save_act != 0;   // h and gelu' are written out for the backward kernel
  pdl_launch_dependents();

  if (threadIdx.x == 0) {
    mbar_init(&bars->w_full, 1);
    mbar_init(&bars->a_full, 1);
    mbar_init(&bars->a_empty, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bars->sh_full[s], 1);
      mbar_init(&bars->sh_empty[s], kFfnEpiWarps);
      mbar_init(&bars->hh_full[s], kFfnEpiWarps);
      mbar_init(&bars->hh_empty[s], save ? 2 : 1);   // second product done (+ the store of h has read it)
    }
    mbar_init(&bars->gp_empty, 1);
    mbar_init(&bars->sz_full, 1);
    mbar_init(&bars->sz_empty, kFfnEpiWarps);
    mbar_init(&bars->y_full, kFfnEpiWarps);
    mbar_init(&bars->y_empty, 1);
    fence_barrier_init();
    prefetch_tmap(&tm_a);
    prefetch_tmap(&tm_w1);
    prefetch_tmap(&tm_w2);
    prefetch_tmap(&tm_out);
    if (save) {
      prefetch_tmap(&tm_h);
      prefetch_tmap(&tm_gp);
    }
    // weights are never written by the preceding kernels of the chain: request them before the dependency wait
    mbar_arrive_expect_tx(&bars->w_full, 2u * kImgBytes);
    load_image(smem_u32(smem + Lay::kW1), &tm_w1, &bars->w_full, 0, 0);
    load_image(smem_u32(smem + Lay::kW2), &tm_w2, &bars->w_full, 0, 0);
  }
  if (warp == 1) tmem_alloc_512(&bars->tmem_base);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  pdl_wait();
  const int n_local = (p.num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (warp == 0) {
    if (lane == 0) {
      for (int n = 0; n < n_local; ++n) {
        const int tile = ffn_tile(p, blockIdx.x + n * gridDim.x);
        mbar_wait(&bars->a_empty, (n & 1u) ^ 1u);
        mbar_arrive_expect_tx(&bars->a_full, (uint32_t)kImgBytes);
        load_image(smem_u32(smem + Lay::kA), &tm_a, &bars->a_full, 0, tile * 128);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(false, false);
      mbar_wait(&bars->w_full, 0u);
      tcgen05_fence_after();
      const uint32_t w1 = smem_u32(smem + Lay::kW1), w2 = smem_u32(smem + Lay::kW2);
      auto mma1 = [&](int n) {  // S_h[n % 2] = a(n) W1^T
        const int sl = n & 1;
        mbar_wait(&bars->a_full, n & 1u);
        mbar_wait(&bars->sh_empty[sl], ((n >> 1) & 1u) ^ 1u);
        tcgen05_fence_after();
        mma_128x128x128(tmem_base + sl * 128u, smem_u32(smem + Lay::kA), false, w1, false, idesc, false);
        umma_commit(&bars->sh_full[sl]);
        umma_commit(&bars->a_empty);
      };
      mma1(0);
      for (int n = 0; n < n_local; ++n) {
        if (n + 1 < n_local) mma1(n + 1);
        const int hb = n & 1;
        mbar_wait(&bars->hh_full[hb], (n >> 1) & 1u);
        mbar_wait(&bars->sz_empty, (n & 1u) ^ 1u);
        tcgen05_fence_after();
        mma_128x128x128(tmem_base + 256u, smem_u32(smem + Lay::kHH + hb * kImgBytes), false, w2, false, idesc, false);
        umma_commit(&bars->sz_full);
        umma_commit(&bars->hh_empty[hb]);
      }
    }
  } else if (warp == 2) {
    if (lane == 0) {
      auto store_act = [&](int n) {
        if (!save) return;
        const int tile = ffn_tile(p, blockIdx.x + n * gridDim.x);
        const int hb = n & 1;
        mbar_wait(&bars->hh_full[hb], (n >> 1) & 1u);
        store_image(&tm_h, smem_u32(smem + Lay::kHH + hb * kImgBytes), 0, tile * 128);
        store_image(&tm_gp, smem_u32(smem + Lay::kGP), 0, tile * 128);
        tma_store_commit();
        tma_store_wait_read0();
        mbar_arrive(&bars->hh_empty[hb]);
        mbar_arrive(&bars->gp_empty);
      };
      store_act(0);
      for (int n = 0; n < n_local; ++n) {
        if (n + 1 < n_local) store_act(n + 1);
        const int tile = ffn_tile(p, blockIdx.x + n * gridDim.x);
        mbar_wait(&bars->y_full, n & 1u);
        store_image(&tm_out, smem_u32(smem + Lay::kYS), 0, tile * 128);
        tma_store_commit();
        tma_store_wait_read0();
        mbar_arrive(&bars->y_empty);
      }
      tma_store_wait_all0();
    }
  } else {
    const int ew = warp - 3;
    const int quarter = warp & 3;   // TMEM lane quarter this warp may access
    const int cq = ew >> 2;         // column quarter
    const int r = quarter * 32 + lane;
    const int c0 = cq * 32;
    const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const uint32_t lane_addr = lane_base + (uint32_t)c0;
    const float ks = p.dropout_p > 0.f ? 1.f / (1.f - p.dropout_p) : 1.f;

    auto epi_gelu = [&](int n) {
      const int sl = n & 1;
      mbar_wait(&bars->sh_full[sl], (n >> 1) & 1u);
      tcgen05_fence_after();
      float v[32];
      {
        uint32_t acc[32];
        tmem_ld_x32(lane_addr + sl * 128u, acc);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(acc[j]);
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->sh_empty[sl]);
      mbar_wait(&bars->hh_empty[sl], ((n >> 1) & 1u) ^ 1u);
      if (save) mbar_wait(&bars->gp_empty, (n & 1u) ^ 1u);
      unsigned char* hh = smem + Lay::kHH + sl * kImgBytes;
      unsigned char* gpi = smem + Lay::kGP;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        float b[8], h[8], gp[8];
        ld8f(p.b1 + c0 + g * 8, b);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float x = v[g * 8 + j] + b[j];
          float cdf, pdf_x;
          gelu_parts(x, cdf, pdf_x);
          h[j] = x * cdf;
          gp[j] = cdf + pdf_x;
        }
        *reinterpret_cast<uint4*>(hh + img_off(r, cq * 4 + g)) = pack8f(h);
        if (save) *reinterpret_cast<uint4*>(gpi + img_off(r, cq * 4 + g)) = pack8f(gp);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->hh_full[sl]);
    };

    auto epi_ln = [&](int n) {
      const int tile = ffn_tile(p, blockIdx.x + n * gridDim.x);
      const long long tok = (long long)tile * 128 + r;
      // the residual row comes straight from L2 (the tile was loaded by TMA a moment ago); issue before the wait
      uint4 res[4];
      if (tok < p.T) {
        const uint4* src = reinterpret_cast<const uint4*>(p.a + tok * p.ld_a + c0);
#pragma unroll
        for (int g = 0; g < 4; ++g) res[g] = __ldg(src + g);
      } else {
#pragma unroll
        for (int g = 0; g < 4; ++g) res[g] = make_uint4(0, 0, 0, 0);
      }
      mbar_wait(&bars->sz_full, n & 1u);
      tcgen05_fence_after();
      float v[32];
      {
        uint32_t acc[32];
        tmem_ld_x32(lane_addr + 256u, acc);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(acc[j]);
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->sz_empty);
      // z = bf16(dropout(acc + b2) + a): the backward kernel forms the same z the same way
      float lsum = 0.f;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        float b[8], x[8];
        ld8f(p.b2 + c0 + g * 8, b);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[g * 8 + j] += b[j];
        if (p.dropout_p > 0.f) {
          const uint64_t idx = (uint64_t)tok * 128u + (uint64_t)(c0 + g * 8);
          const uint32_t k8 = dropout_keep8(p.seed, p.site, idx, p.dropout_p);
#pragma unroll
          for (int j = 0; j < 8; ++j) v[g * 8 + j] = (k8 >> j) & 1u ? v[g * 8 + j] * ks : 0.f;
        }
        unpack8f(res[g], x);
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
          float lo, hi;
          unpack_bf16x2(pack_bf16x2(x[j] + v[g * 8 + j], x[j + 1] + v[g * 8 + j + 1]), lo, hi);
          v[g * 8 + j] = lo;
          v[g * 8 + j + 1] = hi;
          lsum += lo + hi;
        }
      }
      const float lmean = lsum * (1.f / 32.f);
      float m2 = 0.f;
#pragma unroll
      for (int j = 0; j < 32; ++j) { const float d = v[j] - lmean; m2 = fmaf(d, d, m2); }
      // exchange through 8 spare TMEM columns (two sets, alternating per tile: a fast warp's next write never lands on a
      // set a slow warp is still reading)
      const uint32_t xcol = lane_base + 384u + (uint32_t)((n & 1) * 8);
      tmem_st_x2(xcol + (uint32_t)(cq * 2), lmean, m2);
      tmem_wait_st();
      tcgen05_fence_before();
      named_bar_sync(1, 32 * kFfnEpiWarps);
      tcgen05_fence_after();
      float q[8], mean, rstd;
      tmem_ld_x8(xcol, q);
      combine_stats(q, p.ln_eps, mean, rstd);
      mbar_wait(&bars->y_empty, (n & 1u) ^ 1u);
      unsigned char* ys = smem + Lay::kYS;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        float gm[8], bt[8];
        ld8f(p.ln_g + c0 + g * 8, gm);
        ld8f(p.ln_b + c0 + g * 8, bt);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[g * 8 + j] = (v[g * 8 + j] - mean) * rstd * gm[j] + bt[j];
        *reinterpret_cast<uint4*>(ys + img_off(r, cq * 4 + g)) = pack8f(v + g * 8);
      }
      if (p.out_f32 != nullptr && tok < p.T) {
        float4* o32 = reinterpret_cast<float4*>(p.out_f32 + tok * 128 + c0);
#pragma unroll
        for (int j = 0; j < 8; ++j) o32[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->y_full);
    };

    epi_gelu(0);
    for (int n = 0; n < n_local; ++n) {
      if (n + 1 < n_local) epi_gelu(n + 1);
      epi_ln(n);
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc_512(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------------
// Column sums over the 32 rows a warp holds (lane = row, v[j] = column j): transpose-reduce butterfly, 31 shuffles.
// Returns the sum of column `lane`.
__device__ __forceinline__ float warp_colsum32(float (&v)[32], int lane) {
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int j = 0; j < s; ++j) {
      const float keep = up ? v[j + s] : v[j];
      const float send = up ? v[j] : v[j + s];
      v[j] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  return v[0];
}

struct FfnBwdLayout {
  static constexpr int kW1 = 0, kW2 = kImgBytes;
  static constexpr int kX = 2 * kImgBytes;         // a(n): MN-major operand of the dW1 product (last product of the tile)
  static constexpr int kY = 3 * kImgBytes;         // h(n): operand of the z product and of the dW2 product
  static constexpr int kDO = 4 * kImgBytes;        // d_o: dropout-masked LayerNorm input gradient (operand of d_h, dW2)
  static constexpr int kGP = 5 * kImgBytes;        // gelu'(h_pre) -> (in place) d h_pre -> (in place) staging of d_a
  static constexpr int kExch = 6 * kImgBytes;      // 2 x float2 [128 rows][4 column quarters]: row-statistic partials
  static constexpr int kBar = kExch + 8192;        // (the exchange area doubles as the end-of-kernel reduction scratch)
  static constexpr int kTotal = kBar + 256 + 1024;
};
static_assert(FfnBwdLayout::kTotal <= 232448, "shared memory budget (227 KiB)");
static_assert(FfnFwdLayout::kTotal <= 232448, "shared memory budget (227 KiB)");

struct FfnBwdBars {
  uint64_t w_full;
  uint64_t x_full, y_full, gp_full;
  uint64_t y_free;           // h(n) no longer read by the tensor pipe (after the dW2 product)
  uint64_t xg_free;          // a(n) and d h_pre(n) no longer read by the tensor pipe (after the dW1 product)
  uint64_t st_free;          // the store of d_a(n) has read its staging (the GP buffer)
  uint64_t c2, c3, c4;       // tensor-pipe commits: z | d_h + dW2 | d_a + dW1
  uint64_t e2, e3, e4;       // epilogue hand-offs (one arrival per epilogue warp)
  uint64_t dw_done;
  uint32_t tmem_base;
};
static_assert(sizeof(FfnBwdBars) <= 256, "barrier block");

// Per tile n (Sa = TMEM columns 0..127, Sb = 128..255, dW1 = 256..383, dW2 = 384..511):
//   TMA   h(n) -> Y, gelu'(n) -> GP, a(n) -> X          (h one tile ahead: Y is free after the dW2 product)
//   MMA2  Sb  = h W2^T                                   (issued as soon as E3 of tile n - 1 has drained Sb)
//   E2    z = bf16(dropout(Sb + b2) + a)   [a, dy, dy_b read from global memory in the accumulator layout];
//         LayerNorm backward: dz -> Sa (fp32, tcgen05.st: the accumulator of the d_a product starts from the
//         residual-branch gradient), d_o = dropout-masked dz -> DO; d_gamma, d_beta, d_b2 column sums
//   MMA3  Sb  = d_o W2        MMA4  dW2 += d_o^T h
//   E3    d h_pre = Sb * gelu' -> GP in place; d_b1 column sums
//   MMA5  Sa += d h_pre W1    MMA6  dW1 += d h_pre^T a
//   E4    d_a = Sa -> bf16 -> GP (staging) -> TMA store
// dW1 / dW2 live in TMEM for the whole kernel and are flushed once with vector reductions.
__global__ void __launch_bounds__(kFfnThreads, 1)
ffn_bwd_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_w1,
               const __grid_constant__ CUtensorMap tm_w2, const __grid_constant__ CUtensorMap tm_h,
               const __grid_constant__ CUtensorMap tm_gp, const __grid_constant__ CUtensorMap tm_da, const FfnParams p) {
  using Lay = FfnBwdLayout;
  extern __shared__ __align__(1024) unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  FfnBwdBars* bars = reinterpret_cast<FfnBwdBars*>(smem + Lay::kBar);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_launch_dependents();

  if (threadIdx.x == 0) {
    mbar_init(&bars->w_full, 1);
    mbar_init(&bars->x_full, 1);
    mbar_init(&bars->y_full, 1);
    mbar_init(&bars->gp_full, 1);
    mbar_init(&bars->y_free, 1);
    mbar_init(&bars->xg_free, 1);
    mbar_init(&bars->st_free, 1);
    mbar_init(&bars->c2, 1);
    mbar_init(&bars->c3, 1);
    mbar_init(&bars->c4, 1);
    mbar_init(&bars->e2, kFfnEpiWarps);
    mbar_init(&bars->e3, kFfnEpiWarps);
    mbar_init(&bars->e4, kFfnEpiWarps);
    mbar_init(&bars->dw_done, 1);
    fence_barrier_init();
    prefetch_tmap(&tm_a);
    prefetch_tmap(&tm_w1);
    prefetch_tmap(&tm_w2);
    prefetch_tmap(&tm_h);
    prefetch_tmap(&tm_gp);
    prefetch_tmap(&tm_da);
    mbar_arrive_expect_tx(&bars->w_full, 2u * kImgBytes);
    load_image(smem_u32(smem + Lay::kW1), &tm_w1, &bars->w_full, 0, 0);
    load_image(smem_u32(smem + Lay::kW2), &tm_w2, &bars->w_full, 0, 0);
  }
  if (warp == 1) tmem_alloc_512(&bars->tmem_base);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  pdl_wait();
  const int n_local = (p.num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      for (int n = 0; n < n_local; ++n) {
        const int row0 = ffn_tile(p, blockIdx.x + n * gridDim.x) * 128;
        if (n > 0) mbar_wait(&bars->y_free, (n - 1) & 1u);
        mbar_arrive_expect_tx(&bars->y_full, (uint32_t)kImgBytes);
        load_image(smem_u32(smem + Lay::kY), &tm_h, &bars->y_full, 0, row0);
        if (n > 0) mbar_wait(&bars->xg_free, (n - 1) & 1u);
        mbar_arrive_expect_tx(&bars->x_full, (uint32_t)kImgBytes);
        load_image(smem_u32(smem + Lay::kX), &tm_a, &bars->x_full, 0, row0);
        if (n > 0) mbar_wait(&bars->st_free, (n - 1) & 1u);
        mbar_arrive_expect_tx(&bars->gp_full, (uint32_t)kImgBytes);
        load_image(smem_u32(smem + Lay::kGP), &tm_gp, &bars->gp_full, 0, row0);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t id_kk = make_idesc(false, false);   // activations K-major, weights K-major (y = x W^T)
      constexpr uint32_t id_kn = make_idesc(false, true);    // weights MN-major (dx = dy W)
      constexpr uint32_t id_nn = make_idesc(true, true);     // dW += dy^T x: both operands MN-major
      mbar_wait(&bars->w_full, 0u);
      tcgen05_fence_after();
      const uint32_t w1 = smem_u32(smem + Lay::kW1), w2 = smem_u32(smem + Lay::kW2);
      const uint32_t xi = smem_u32(smem + Lay::kX), yi = smem_u32(smem + Lay::kY);
      const uint32_t doi = smem_u32(smem + Lay::kDO), gpi = smem_u32(smem + Lay::kGP);
      const uint32_t sa = tmem_base, sb = tmem_base + 128u, t_dw1 = tmem_base + 256u, t_dw2 = tmem_base + 384u;
      auto mma2 = [&](int n) {
        mbar_wait(&bars->y_full, n & 1u);
        tcgen05_fence_after();
        mma_128x128x128(sb, yi, false, w2, false, id_kk, false);
        umma_commit(&bars->c2);
      };
      mma2(0);
      for (int n = 0; n < n_local; ++n) {
        const uint32_t ph = n & 1u;
        mbar_wait(&bars->e2, ph);
        tcgen05_fence_after();
        mma_128x128x128(sb, doi, false, w2, true, id_kn, false);
        mma_128x128x128(t_dw2, doi, true, yi, true, id_nn, n > 0);
        umma_commit(&bars->c3);
        umma_commit(&bars->y_free);
        mbar_wait(&bars->e3, ph);       // d h_pre sits in GP, Sb has been drained
        mbar_wait(&bars->x_full, ph);
        tcgen05_fence_after();
        mma_128x128x128(sa, gpi, false, w1, true, id_kn, true);   // accumulates onto the dz the epilogue stored
        mma_128x128x128(t_dw1, gpi, true, xi, true, id_nn, n > 0);
        umma_commit(&bars->c4);
        umma_commit(&bars->xg_free);
        if (n + 1 < n_local) mma2(n + 1);
      }
      umma_commit(&bars->dw_done);
    }
  } else if (warp == 2) {
    // ===================== store warp =====================
    if (lane == 0) {
      for (int n = 0; n < n_local; ++n) {
        const int tile = ffn_tile(p, blockIdx.x + n * gridDim.x);
        mbar_wait(&bars->e4, n & 1u);
        store_image(&tm_da, smem_u32(smem + Lay::kGP), 0, tile * 128);
        tma_store_commit();
        tma_store_wait_read0();
        mbar_arrive(&bars->st_free);
      }
      tma_store_wait_all0();
    }
  } else {
    // ===================== epilogue warps =====================
    const int ew = warp - 3;
    const int quarter = warp & 3;
    const int cq = ew >> 2;
    const int r = quarter * 32 + lane;
    const int c0 = cq * 32;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0;
    const uint32_t t_sa = lane_addr, t_sb = lane_addr + 128u;
    const float ks = p.dropout_p > 0.f ? 1.f / (1.f - p.dropout_p) : 1.f;
    float2* exch = reinterpret_cast<float2*>(smem + Lay::kExch);
    unsigned char* doimg = smem + Lay::kDO;
    unsigned char* gpimg = smem + Lay::kGP;
    float acc_dg = 0.f, acc_dbeta = 0.f, acc_db2 = 0.f, acc_db1 = 0.f;   // column c0 + lane, rows of this warp, all tiles

    for (int n = 0; n < n_local; ++n) {
      const int tile = ffn_tile(p, blockIdx.x + n * gridDim.x);
      const uint32_t ph = n & 1u;
      const long long tok = (long long)tile * 128 + r;
      float v[32];
      uint32_t dyp[16];   // dy of this thread's 32 columns, packed bf16x2 (dy + dy_b summed)

      // ---- E2: recompute z, LayerNorm backward.  a / dy / dy_b rows come from global memory in the accumulator
      // layout (64 contiguous bytes per thread); the loads are issued before the wait for the z product.
      uint4 res[4];
      if (tok < p.T) {
        const uint4* sa_ = reinterpret_cast<const uint4*>(p.a + tok * p.ld_a + c0);
        const uint4* sd_ = reinterpret_cast<const uint4*>(p.dy + tok * p.ld_dy + c0);
        uint4 d[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) { res[g] = __ldg(sa_ + g); d[g] = __ldg(sd_ + g); }
        if (p.dy_b != nullptr) {
          const uint4* sb_ = reinterpret_cast<const uint4*>(p.dy_b + tok * p.ld_dy_b + c0);
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const uint4 e = __ldg(sb_ + g);
            float x[8], y[8];
            unpack8f(d[g], x);
            unpack8f(e, y);
#pragma unroll
            for (int j = 0; j < 8; ++j) x[j] += y[j];
            d[g] = pack8f(x);
          }
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) { dyp[4 * g] = d[g].x; dyp[4 * g + 1] = d[g].y; dyp[4 * g + 2] = d[g].z; dyp[4 * g + 3] = d[g].w; }
      } else {
#pragma unroll
        for (int g = 0; g < 4; ++g) res[g] = make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int j = 0; j < 16; ++j) dyp[j] = 0u;
      }
      mbar_wait(&bars->c2, ph);
      tcgen05_fence_after();
      {
        uint32_t acc[32];
        tmem_ld_x32(t_sb, acc);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(acc[j]);
      }
      uint32_t kbits = 0xffffffffu;
      float lsum = 0.f;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        float b[8], x[8];
        ld8f(p.b2 + c0 + g * 8, b);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[g * 8 + j] += b[j];
        if (p.dropout_p > 0.f) {
          const uint64_t idx = (uint64_t)tok * 128u + (uint64_t)(c0 + g * 8);
          const uint32_t k8 = dropout_keep8(p.seed, p.site, idx, p.dropout_p);
          kbits = (kbits & ~(0xffu << (8 * g))) | (k8 << (8 * g));
#pragma unroll
          for (int j = 0; j < 8; ++j) v[g * 8 + j] = (k8 >> j) & 1u ? v[g * 8 + j] * ks : 0.f;
        }
        unpack8f(res[g], x);
        // z rounded to bf16 exactly as the forward kernel rounds it before its LayerNorm
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
          float lo, hi;
          unpack_bf16x2(pack_bf16x2(x[j] + v[g * 8 + j], x[j + 1] + v[g * 8 + j + 1]), lo, hi);
          v[g * 8 + j] = lo;
          v[g * 8 + j + 1] = hi;
          lsum += lo + hi;
        }
      }
      {
        const float lmean = lsum * (1.f / 32.f);
        float m2 = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) { const float d = v[j] - lmean; m2 = fmaf(d, d, m2); }
        exch[r * 4 + cq] = make_float2(lmean, m2);
      }
      named_bar_sync(1, 32 * kFfnEpiWarps);
      float mean, rstd;
      {
        float q[8];
        const float4 p01 = *reinterpret_cast<const float4*>(&exch[r * 4]);
        const float4 p23 = *reinterpret_cast<const float4*>(&exch[r * 4 + 2]);
        q[0] = p01.x; q[1] = p01.y; q[2] = p01.z; q[3] = p01.w; q[4] = p23.x; q[5] = p23.y; q[6] = p23.z; q[7] = p23.w;
        combine_stats(q, p.ln_eps, mean, rstd);
      }
      float s1 = 0.f, s2 = 0.f;
      {
        float pg[32];   // dy * xhat: the d_gamma terms
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float gm[8];
          ld8f(p.ln_g + c0 + g * 8, gm);
#pragma unroll
          for (int j = 0; j < 8; j += 2) {
            float d0, d1;
            unpack_bf16x2(dyp[g * 4 + (j >> 1)], d0, d1);
            const int c = g * 8 + j;
            const float x0 = (v[c] - mean) * rstd, x1 = (v[c + 1] - mean) * rstd;
            v[c] = x0;
            v[c + 1] = x1;
            const float g0 = d0 * gm[j], g1 = d1 * gm[j + 1];
            s1 += g0 + g1;
            s2 = fmaf(g0, x0, fmaf(g1, x1, s2));
            pg[c] = d0 * x0;
            pg[c + 1] = d1 * x1;
          }
        }
        // second exchange in the other half of the exchange area: slow readers of the first one are not disturbed
        exch[512 + r * 4 + cq] = make_float2(s1, s2);
        acc_dg += warp_colsum32(pg, lane);
      }
      {
        float dyf[32];
#pragma unroll
        for (int j = 0; j < 16; ++j) unpack_bf16x2(dyp[j], dyf[2 * j], dyf[2 * j + 1]);
        acc_dbeta += warp_colsum32(dyf, lane);
      }
      named_bar_sync(1, 32 * kFfnEpiWarps);
      {
        const float4 p01 = *reinterpret_cast<const float4*>(&exch[512 + r * 4]);
        const float4 p23 = *reinterpret_cast<const float4*>(&exch[512 + r * 4 + 2]);
        const float m1 = (p01.x + p01.z + p23.x + p23.z) * (1.f / 128.f);
        const float m2 = (p01.y + p01.w + p23.y + p23.w) * (1.f / 128.f);
        uint32_t dzb[32];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float gm[8];
          ld8f(p.ln_g + c0 + g * 8, gm);
#pragma unroll
          for (int j = 0; j < 8; j += 2) {
            float d0, d1;
            unpack_bf16x2(dyp[g * 4 + (j >> 1)], d0, d1);
            const int c = g * 8 + j;
            const float z0 = rstd * (d0 * gm[j] - m1 - v[c] * m2);
            const float z1 = rstd * (d1 * gm[j + 1] - m1 - v[c + 1] * m2);
            dzb[c] = __float_as_uint(z0);
            dzb[c + 1] = __float_as_uint(z1);
            v[c] = (kbits >> c) & 1u ? z0 * ks : 0.f;         // d_o: gradient wrt the dense output
            v[c + 1] = (kbits >> (c + 1)) & 1u ? z1 * ks : 0.f;
          }
          *reinterpret_cast<uint4*>(doimg + img_off(r, cq * 4 + g)) = pack8f(v + g * 8);
        }
        tmem_st_x32(t_sa, dzb);   // residual-branch gradient: the d_a accumulator starts from it
        tmem_wait_st();
      }
      tcgen05_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->e2);
      acc_db2 += warp_colsum32(v, lane);

      // ---- E3: d h_pre = d_h * gelu'(h_pre)
      mbar_wait(&bars->gp_full, ph);
      mbar_wait(&bars->c3, ph);
      tcgen05_fence_after();
      {
        uint32_t acc[32];
        tmem_ld_x32(t_sb, acc);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(acc[j]);
      }
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint4* gpp = reinterpret_cast<uint4*>(gpimg + img_off(r, cq * 4 + g));
        float gp[8];
        unpack8f(*gpp, gp);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[g * 8 + j] *= gp[j];
        *gpp = pack8f(v + g * 8);
      }
      tcgen05_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->e3);
      acc_db1 += warp_colsum32(v, lane);

      // ---- E4: d_a
      mbar_wait(&bars->c4, ph);
      tcgen05_fence_after();
      {
        uint32_t acc[32];
        tmem_ld_x32(t_sa, acc);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(acc[j]);
      }
#pragma unroll
      for (int g = 0; g < 4; ++g) *reinterpret_cast<uint4*>(gpimg + img_off(r, cq * 4 + g)) = pack8f(v + g * 8);
      tcgen05_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->e4);
    }

    // ---- bias / LayerNorm parameter gradients: reduce the four lane quarters in shared memory, one atomic per column
    {
      named_bar_sync(1, 32 * kFfnEpiWarps);   // every reader of the exchange area is done: reuse it
      float* red = reinterpret_cast<float*>(smem + Lay::kExch);   // [4 quantities][4 quarters][128 columns]
      red[(0 * 4 + quarter) * 128 + c0 + lane] = acc_dg;
      red[(1 * 4 + quarter) * 128 + c0 + lane] = acc_dbeta;
      red[(2 * 4 + quarter) * 128 + c0 + lane] = acc_db2;
      red[(3 * 4 + quarter) * 128 + c0 + lane] = acc_db1;
      named_bar_sync(1, 32 * kFfnEpiWarps);
      const int t = threadIdx.x - 96;   // 0..511: quantity t / 128, column t % 128
      const int qn = t >> 7, col = t & 127;
      const float sum = red[(qn * 4 + 0) * 128 + col] + red[(qn * 4 + 1) * 128 + col] + red[(qn * 4 + 2) * 128 + col] +
                        red[(qn * 4 + 3) * 128 + col];
      float* dst = qn == 0 ? p.dg : (qn == 1 ? p.dbeta : (qn == 2 ? p.db2 : p.db1));
      if (dst != nullptr && sum != 0.f) atomicAdd(dst + col, sum);
    }
    // ---- dW1 / dW2 flush (rotated order per CTA: 148 CTAs add into the same 2 x 64 KB)
    mbar_wait(&bars->dw_done, 0u);
    tcgen05_fence_after();
#pragma unroll 1
    for (int m = 0; m < 2; ++m) {
      const int mm = (m + (int)blockIdx.x) & 1;
      float* dst = (mm == 0 ? p.dw1 : p.dw2) + (long long)r * 128 + c0;
      uint32_t acc[32];
      tmem_ld_x32(lane_addr + 256u + (uint32_t)mm * 128u, acc);
      tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        red_add_v4(dst + j, __uint_as_float(acc[j]), __uint_as_float(acc[j + 1]), __uint_as_float(acc[j + 2]),
                   __uint_as_float(acc[j + 3]));
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc_512(tmem_base);
  }
}

static int check_ffn(const pmgt_ffn_args* a, bool bwd) {
  PMGT_REQUIRE(a && a->a && a->w1 && a->w2 && a->b1 && a->b2 && a->ln_g && a->ln_b, "pmgt_ffn: null argument");
  PMGT_REQUIRE(a->T >= 0 && a->T < (1ll << 31) - 128, "pmgt_ffn: bad T");
  PMGT_REQUIRE(a->ld_a % 8 == 0 && ((uintptr_t)a->a & 15) == 0, "pmgt_ffn: a must be 16-byte aligned with a pitch multiple of 8");
  PMGT_REQUIRE((((uintptr_t)a->w1 | (uintptr_t)a->w2 | (uintptr_t)a->b1 | (uintptr_t)a->b2 | (uintptr_t)a->ln_g |
                 (uintptr_t)a->ln_b) & 15) == 0, "pmgt_ffn: parameter alignment");
  PMGT_REQUIRE(a->dropout_p >= 0.f && a->dropout_p < 1.f, "pmgt_ffn: bad dropout_p");
  PMGT_REQUIRE((((uintptr_t)a->h | (uintptr_t)a->gp) & 15) == 0 && a->ld_h % 8 == 0, "pmgt_ffn: h / gp alignment");
  if (!bwd) {
    PMGT_REQUIRE(a->out && a->ld_out % 8 == 0 && ((uintptr_t)a->out & 15) == 0, "pmgt_ffn_fwd: out alignment");
    PMGT_REQUIRE(((uintptr_t)a->out_f32 & 15) == 0, "pmgt_ffn_fwd: out_f32 alignment");
    PMGT_REQUIRE((a->h == nullptr) == (a->gp == nullptr), "pmgt_ffn_fwd: pass both h and gp, or neither");
  } else {
    PMGT_REQUIRE(a->dy && a->da && a->dw1 && a->dw2 && a->h && a->gp, "pmgt_ffn_bwd: h, gp, dy, da, dw1, dw2 required");
    PMGT_REQUIRE(a->ld_dy % 8 == 0 && a->ld_da % 8 == 0 && a->ld_dy_b % 8 == 0 &&
                 (((uintptr_t)a->dy | (uintptr_t)a->da | (uintptr_t)a->dy_b | (uintptr_t)a->dw1 | (uintptr_t)a->dw2) & 15) == 0,
                 "pmgt_ffn_bwd: gradient buffer alignment");
  }
  return PMGT_OK;
}

static void fill_params(const pmgt_ffn_args* a, FfnParams& p) {
  memset(&p, 0, sizeof(p));
  p.T = (int)a->T;
  p.num_tiles = (int)((a->T + 127) / 128);
  p.reverse = next_tile_order();
  p.b1 = a->b1; p.b2 = a->b2; p.ln_g = a->ln_g; p.ln_b = a->ln_b; p.ln_eps = a->ln_eps;
  p.dropout_p = a->dropout_p; p.seed = a->dropout_seed; p.site = a->dropout_site;
  p.out_f32 = a->out_f32;
  p.a = a->a; p.ld_a = a->ld_a;
  p.save_act = a->h != nullptr;
  p.dy = a->dy; p.ld_dy = a->ld_dy;
  p.dy_b = a->dy_b; p.ld_dy_b = a->ld_dy_b;
  p.dw1 = a->dw1; p.dw2 = a->dw2; p.db1 = a->db1; p.db2 = a->db2; p.dg = a->d_ln_g; p.dbeta = a->d_ln_b;
}

}  // namespace pmgt

using namespace pmgt;

extern "C" {

int pmgt_ffn_fwd(const pmgt_ffn_args* a, void* stream) {
  int rc = check_ffn(a, false);
  if (rc) return rc;
  if (a->T == 0) return PMGT_OK;
  static unsigned long long configured = 0;
  if (first_use_on_device(configured))
    PMGT_CHECK_CUDA(cudaFuncSetAttribute(ffn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FfnFwdLayout::kTotal));
  CUtensorMap ta, tw1, tw2, to, th, tg;
  if ((rc = make_tmap(&ta, a->a, 128, a->T, a->ld_a, 64, 128))) return rc;
  if ((rc = make_tmap(&tw1, a->w1, 128, 128, 128, 64, 128))) return rc;
  if ((rc = make_tmap(&tw2, a->w2, 128, 128, 128, 64, 128))) return rc;
  if ((rc = make_tmap(&to, a->out, 128, a->T, a->ld_out, 64, 128))) return rc;
  th = to;
  tg = to;
  if (a->h != nullptr) {
    if ((rc = make_tmap(&th, a->h, 128, a->T, a->ld_h, 64, 128))) return rc;
    if ((rc = make_tmap(&tg, a->gp, 128, a->T, a->ld_h, 64, 128))) return rc;
  }
  FfnParams p;
  fill_params(a, p);
  int grid = num_sms();
  if (grid > p.num_tiles) grid = p.num_tiles;
  PMGT_CHECK_CUDA(launch_kernel(true, ffn_fwd_kernel, dim3(grid), dim3(kFfnThreads), FfnFwdLayout::kTotal,
                                (cudaStream_t)stream, ta, tw1, tw2, to, th, tg, p));
  return PMGT_OK;
}

int pmgt_ffn_bwd(const pmgt_ffn_args* a, void* stream) {
  int rc = check_ffn(a, true);
  if (rc) return rc;
  if (a->T == 0) return PMGT_OK;
  static unsigned long long configured = 0;
  if (first_use_on_device(configured))
    PMGT_CHECK_CUDA(cudaFuncSetAttribute(ffn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FfnBwdLayout::kTotal));
  CUtensorMap ta, tw1, tw2, th, tg, tda;
  if ((rc = make_tmap(&ta, a->a, 128, a->T, a->ld_a, 64, 128))) return rc;
  if ((rc = make_tmap(&tw1, a->w1, 128, 128, 128, 64, 128))) return rc;
  if ((rc = make_tmap(&tw2, a->w2, 128, 128, 128, 64, 128))) return rc;
  if ((rc = make_tmap(&th, a->h, 128, a->T, a->ld_h, 64, 128))) return rc;
  if ((rc = make_tmap(&tg, a->gp, 128, a->T, a->ld_h, 64, 128))) return rc;
  if ((rc = make_tmap(&tda, a->da, 128, a->T, a->ld_da, 64, 128))) return rc;
  FfnParams p;
  fill_params(a, p);
  int grid = num_sms();
  if (grid > p.num_tiles) grid = p.num_tiles;
  PMGT_CHECK_CUDA(launch_kernel(true, ffn_bwd_kernel, dim3(grid), dim3(kFfnThreads), FfnBwdLayout::kTotal,
                                (cudaStream_t)stream, ta, tw1, tw2, th, tg, tda, p));
  return PMGT_OK;
}

}  // extern "C"
