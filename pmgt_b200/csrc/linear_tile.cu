// Persistent tcgen05 "token-tile" kernels for the small hidden sizes of the default PMGT encoder
// (H = I = 128): every nn.Linear of PMGTSelfAttention / BertSelfOutput / BertIntermediate / BertOutput
// (pmgt/pmgt/modeling_pmgt.py:429-433,371,322-325), forward and backward, with the element-wise work
// that follows it in the reference fused into the epilogue.
//
// Unit of work: a tile of 128 consecutive tokens.  One CTA per SM loops over tiles.  All operands are
// "images": a [128 rows][128 cols] bf16 block stored as two 64-column slabs of 128 rows x 128 bytes with
// the 128-byte TMA/UMMA swizzle (32 KiB).  The same image serves as a K-major operand (K = its columns)
// and as an MN-major operand (K = its rows), so activations are loaded ONCE by TMA and weights stay
// resident in shared memory for the whole kernel.
//
//   linear_tile_kernel   out[T][N] = epi(sum_kc X_kc * W_kc):  y = x W^T (+ bias, GELU, residual + dropout +
//                        LayerNorm) or dx = dy W (* gelu').  Warp roles: TMA producer | MMA issuer |
//                        1-2 epilogue groups of 4 warps (TMEM -> registers -> swizzled staging -> TMA store).
//                        Four 128-column TMEM accumulator slots are handed between MMA and epilogue through
//                        mbarriers, so the MMAs of the next items overlap the epilogue of the current one.
//   dw_tile_kernel       dW[N][K] += dY^T X and dbias += colsum(dY): both operands MN-major images, the
//                        accumulators stay in TMEM across ALL tiles of the CTA and are flushed once with
//                        vector reductions; four extra warps form the column sums from the staged images.
//
// Everything here is HBM-bound by construction (<= 128 FLOP/B); the roofline is the measured copy bandwidth.
#include "tile.cuh"

namespace pmgt {

constexpr int kAccSlots = 4;
constexpr int kEpiWarps = 16;                     // 4 TMEM lane quarters x 4 column quarters (32 columns per thread)
constexpr int kLtThreads = 96 + 32 * kEpiWarps;   // producer | MMA | store | 16 epilogue warps
constexpr int kDwWarps = 4;                       // fused dX + dW variant: column sums of dY, then the dW flush

enum { LT_BIAS = PMGT_LT_BIAS, LT_GELU = PMGT_LT_GELU, LT_RES_LN = PMGT_LT_RES_LN, LT_PLAIN = PMGT_LT_PLAIN,
       LT_GELU_BWD = PMGT_LT_GELU_BWD };

struct LtParams {
  int T, num_tiles, N;
  const float* bias;
  const float* ln_g;
  const float* ln_b;
  float ln_eps;
  float dropout_p;
  uint64_t seed;
  uint32_t site;
  float* out_f32;
  float* dw;          // fused dW (DW variants): [N][ld_dw] fp32, accumulated
  long long ld_dw;
  float* dbias;       // [N] fp32, accumulated, may be NULL
  int reverse;        // walk the token tiles in descending order (see next_tile_order)
};

template <int NC, int KC, int EPI, int SA, int SE, int NSTG, int SX = 0>
struct LtLayout {
  static constexpr bool kHasE = (EPI == LT_RES_LN || EPI == LT_GELU_BWD);
  static constexpr int kNOut = (EPI == LT_GELU) ? 2 : 1;
  static constexpr int kW = 0;
  static constexpr int kA = kW + NC * KC * kImgBytes;
  static constexpr int kX = kA + SA * kImgBytes;          // SX > 0: second activation stream (the dW operand)
  static constexpr int kE = kX + SX * kImgBytes;
  static constexpr int kStg = kE + (kHasE ? SE : 0) * kImgBytes;
  static constexpr int kBar = kStg + NSTG * kNOut * kImgBytes;
  static constexpr int kTotal = kBar + 256 + 1024;        // barriers + alignment slack
};

struct LtBars {
  uint64_t w_full;
  uint64_t a_full[3], a_empty[3];
  uint64_t e_full[3], e_empty[3];
  uint64_t acc_full[kAccSlots], acc_empty[kAccSlots];
  uint64_t stg_full[2], stg_empty[2];
  uint64_t x_full[2], x_empty[2];
  uint64_t dw_done;
  uint32_t tmem_base;
};
static_assert(sizeof(LtBars) <= 256, "barrier block");

// Work item = one 128-token x 128-column output block.  Roles:
//   warp 0      TMA producer: weights once, then activation (A) and epilogue-input (E) images, SA / SE deep
//   warp 1      MMA issuer, accumulators in 4 TMEM slots of 128 columns
//   warp 2      store warp: waits for a staged output block, issues the TMA stores, waits until they have READ the
//               staging images and hands the buffers (and, for RES_LN, the E slot that doubles as the z image) back
//   warps 3-18  epilogue: warp w reads TMEM lanes 32*(w%4).. (its hardware quarter) and columns 32*((w-3)/4)..,
//               i.e. one row x 32 columns per thread.  Sixteen warps (instead of four per block) are what keeps the
//               issue slots busy: the epilogue math (GELU, LayerNorm, Philox dropout) was the limiter, not HBM.
template <int NC, int KC, bool B_MN, int EPI, int SA, int SE, int NSTG, int SX>
__global__ void __launch_bounds__(kLtThreads + (SX > 0 ? 32 * kDwWarps : 0), 1)
linear_tile_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w,
                   const __grid_constant__ CUtensorMap tm_e, const __grid_constant__ CUtensorMap tm_out,
                   const __grid_constant__ CUtensorMap tm_aux, const LtParams p) {
  using Lay = LtLayout<NC, KC, EPI, SA, SE, NSTG, SX>;
  constexpr bool kHasE = Lay::kHasE;
  constexpr int kNOut = Lay::kNOut;
  // DW: the dX kernels of a Linear also form its weight gradient.  The dY tile already in shared memory is the K-major
  // operand of dX = dY W and, read MN-major, the operand of dW += dY^T X; X (the Linear's input) arrives as a second
  // TMA stream (tensor map passed in the tm_aux slot), dW accumulates in the last 128 TMEM columns over all tiles of
  // the CTA and is flushed once, four extra warps form the bias gradient from the staged dY images.
  constexpr bool DW = SX > 0;
  constexpr int kSlots = DW ? kAccSlots - 1 : kAccSlots;
  static_assert(!kHasE || NC == 1, "epilogue-input variants are single-chunk");
  static_assert(NSTG == 1 || NSTG == 2, "one or two staging buffers");
  static_assert(SA <= 3 && SE <= 3 && SX <= 2, "barrier arrays");
  static_assert(!DW || (B_MN && NC == 1 && KC == 1 && (EPI == LT_PLAIN || EPI == LT_GELU_BWD)), "dW fusion: dX kernels only");
  extern __shared__ __align__(1024) unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  LtBars* bars = reinterpret_cast<LtBars*>(smem + Lay::kBar);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_launch_dependents();  // the next kernel's CTAs may take an SM as soon as this grid's CTAs leave it

  if (threadIdx.x == 0) {
    mbar_init(&bars->w_full, 1);
    for (int s = 0; s < 3; ++s) {
      mbar_init(&bars->a_full[s], 1);
      mbar_init(&bars->a_empty[s], DW ? 1 + kDwWarps : 1);  // MMA commit (+ one arrival per column-sum warp)
    }
    for (int s = 0; s < 3; ++s) {
      mbar_init(&bars->e_full[s], 1);
      mbar_init(&bars->e_empty[s], EPI == LT_RES_LN ? 1 : kEpiWarps);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bars->stg_full[s], kEpiWarps);
      mbar_init(&bars->stg_empty[s], 1);
      mbar_init(&bars->x_full[s], 1);
      mbar_init(&bars->x_empty[s], 1);
    }
    for (int s = 0; s < kAccSlots; ++s) {
      mbar_init(&bars->acc_full[s], 1);
      mbar_init(&bars->acc_empty[s], kEpiWarps);
    }
    mbar_init(&bars->dw_done, 1);
    fence_barrier_init();
    prefetch_tmap(&tm_x);
    prefetch_tmap(&tm_w);
    prefetch_tmap(&tm_out);
    if (kHasE) prefetch_tmap(&tm_e);
    if (EPI == LT_GELU || EPI == LT_RES_LN || DW) prefetch_tmap(&tm_aux);
    // The resident weight images are requested BEFORE the dependency wait: weights (the bf16 shadow written once per
    // step by the cast kernel) are never produced by the kernels that immediately precede this one in the chain, so
    // their load latency overlaps the predecessor's tail.  Callers that rewrite `w` right before this call must turn
    // programmatic dependent launch off (pmgt_set_pdl).
    mbar_arrive_expect_tx(&bars->w_full, (uint32_t)(NC * KC * kImgBytes));
    for (int kc = 0; kc < KC; ++kc)
      for (int n = 0; n < NC; ++n) {
        const uint32_t dst = smem_u32(smem + Lay::kW + (kc * NC + n) * kImgBytes);
        const int row0 = B_MN ? kc * 128 : n * 128, col0 = B_MN ? n * 128 : kc * 128;
        tma_load_2d(dst, &tm_w, &bars->w_full, col0, row0);
        tma_load_2d(dst + kSlabBytes, &tm_w, &bars->w_full, col0 + 64, row0);
      }
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)),
                 "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  pdl_wait();  // everything above overlapped the predecessor's tail; its outputs are visible from here on

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t ia = 0, ie = 0, ix = 0;
      (void)ix;
      for (int lt = blockIdx.x; lt < p.num_tiles; lt += gridDim.x) {
        const int tile = p.reverse ? p.num_tiles - 1 - lt : lt;
        for (int kc = 0; kc < KC; ++kc, ++ia) {
          const int s = ia % SA;
          mbar_wait(&bars->a_empty[s], ((ia / SA) & 1u) ^ 1u);
          mbar_arrive_expect_tx(&bars->a_full[s], (uint32_t)kImgBytes);
          const uint32_t dst = smem_u32(smem + Lay::kA + s * kImgBytes);
          tma_load_2d(dst, &tm_x, &bars->a_full[s], kc * 128, tile * 128);
          tma_load_2d(dst + kSlabBytes, &tm_x, &bars->a_full[s], kc * 128 + 64, tile * 128);
        }
        if (DW) {
          const int s = ix % (SX > 0 ? SX : 1);
          mbar_wait(&bars->x_empty[s], ((ix / (SX > 0 ? SX : 1)) & 1u) ^ 1u);
          mbar_arrive_expect_tx(&bars->x_full[s], (uint32_t)kImgBytes);
          const uint32_t dst = smem_u32(smem + Lay::kX + s * kImgBytes);
          tma_load_2d(dst, &tm_aux, &bars->x_full[s], 0, tile * 128);
          tma_load_2d(dst + kSlabBytes, &tm_aux, &bars->x_full[s], 64, tile * 128);
          ++ix;
        }
        if (kHasE) {
          const int s = ie % SE;
          mbar_wait(&bars->e_empty[s], ((ie / SE) & 1u) ^ 1u);
          mbar_arrive_expect_tx(&bars->e_full[s], (uint32_t)kImgBytes);
          const uint32_t dst = smem_u32(smem + Lay::kE + s * kImgBytes);
          tma_load_2d(dst, &tm_e, &bars->e_full[s], 0, tile * 128);
          tma_load_2d(dst + kSlabBytes, &tm_e, &bars->e_full[s], 64, tile * 128);
          ++ie;
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(false, B_MN);
      mbar_wait(&bars->w_full, 0u);
      tcgen05_fence_after();
      uint32_t ia = 0, item = 0, ix = 0;
      (void)ix;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, item += NC) {
        for (int kc = 0; kc < KC; ++kc, ++ia) {
          const int s = ia % SA;
          mbar_wait(&bars->a_full[s], (ia / SA) & 1u);
          tcgen05_fence_after();
          const uint32_t a_img = smem_u32(smem + Lay::kA + s * kImgBytes);
          for (int n = 0; n < NC; ++n) {
            const uint32_t it = item + n, slot = it % kSlots;
            if (kc == 0) {
              mbar_wait(&bars->acc_empty[slot], ((it / kSlots) & 1u) ^ 1u);
              tcgen05_fence_after();
            }
            const uint32_t b_img = smem_u32(smem + Lay::kW + (kc * NC + n) * kImgBytes);
            mma_128x128x128(tmem_base + slot * 128u, a_img, false, b_img, B_MN, idesc, kc > 0);
            if (kc == KC - 1) umma_commit(&bars->acc_full[slot]);
          }
          if (DW) {
            // dW[128 x 128] += dY^T (M = dY columns, K = tokens) * X (K = tokens, N = X columns): both MN-major
            constexpr uint32_t idesc_dw = make_idesc(true, true);
            const int sx = ix % (SX > 0 ? SX : 1);
            mbar_wait(&bars->x_full[sx], (ix / (SX > 0 ? SX : 1)) & 1u);
            tcgen05_fence_after();
            const uint32_t x_img = smem_u32(smem + Lay::kX + sx * kImgBytes);
            mma_128x128x128(tmem_base + (uint32_t)(kSlots * 128), a_img, true, x_img, true, idesc_dw, ix > 0);
            umma_commit(&bars->x_empty[sx]);
            ++ix;
          }
          umma_commit(&bars->a_empty[s]);
        }
      }
      if (DW) umma_commit(&bars->dw_done);
    }
  } else if (warp == 2) {
    // ===================== store warp =====================
    if (lane == 0) {
      uint32_t it = 0;
      for (int lt = blockIdx.x; lt < p.num_tiles; lt += gridDim.x) {
        const int tile = p.reverse ? p.num_tiles - 1 - lt : lt;
        for (int n = 0; n < NC; ++n, ++it) {
          const uint32_t buf = it % NSTG;
          mbar_wait(&bars->stg_full[buf], (it / NSTG) & 1u);
          const uint32_t stg0 = smem_u32(smem + Lay::kStg + buf * kNOut * kImgBytes);
          const int col = n * 128, row = tile * 128;
          if (EPI == LT_GELU) {
            tma_store_2d(&tm_aux, stg0, col, row);
            tma_store_2d(&tm_aux, stg0 + kSlabBytes, col + 64, row);
            tma_store_2d(&tm_out, stg0 + kImgBytes, col, row);
            tma_store_2d(&tm_out, stg0 + kImgBytes + kSlabBytes, col + 64, row);
          } else {
            tma_store_2d(&tm_out, stg0, col, row);
            tma_store_2d(&tm_out, stg0 + kSlabBytes, col + 64, row);
            if (EPI == LT_RES_LN) {
              const uint32_t eimg = smem_u32(smem + Lay::kE + (it % SE) * kImgBytes);
              tma_store_2d(&tm_aux, eimg, col, row);
              tma_store_2d(&tm_aux, eimg + kSlabBytes, col + 64, row);
            }
          }
          tma_store_commit();
          tma_store_wait_read0();
          mbar_arrive(&bars->stg_empty[buf]);
          if (EPI == LT_RES_LN) mbar_arrive(&bars->e_empty[it % SE]);
        }
      }
      tma_store_wait_all0();
    }
  } else if (warp < 3 + kEpiWarps) {
    // ===================== epilogue warps =====================
    const int quarter = warp & 3;               // TMEM lane quarter this warp may access
    const int cq = (warp - 3) >> 2;             // column quarter
    const int r = quarter * 32 + lane;          // tile row == TMEM lane
    const int c0 = cq * 32;
    const float ks = p.dropout_p > 0.f ? 1.f / (1.f - p.dropout_p) : 1.f;
    float ln_gm[8], ln_bt[8];  // LayerNorm affine of the 8 columns this lane owns in phase B of the RES_LN epilogue
    if (EPI == LT_RES_LN) {
      ld8f(p.ln_g + (lane & 15) * 8, ln_gm);
      ld8f(p.ln_b + (lane & 15) * 8, ln_bt);
    }
    uint32_t it = 0;
    for (int lt = blockIdx.x; lt < p.num_tiles; lt += gridDim.x) {
      const int tile = p.reverse ? p.num_tiles - 1 - lt : lt;
      for (int n = 0; n < NC; ++n, ++it) {
        const uint32_t slot = it % kSlots;
        mbar_wait(&bars->acc_full[slot], (it / kSlots) & 1u);
        tcgen05_fence_after();
        float v[32];
        {
          uint32_t acc[32];
          tmem_ld_x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + slot * 128u + (uint32_t)c0, acc);
          tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(acc[j]);
        }
        // accumulator slot drained: the MMAs of a later block may overwrite it while this block is post-processed
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->acc_empty[slot]);

        unsigned char* eimg = nullptr;
        int se = 0;
        if (kHasE) {
          se = it % SE;
          mbar_wait(&bars->e_full[se], (it / SE) & 1u);
          eimg = smem + Lay::kE + se * kImgBytes;
        }
        const uint32_t buf = it % NSTG;
        // RES_LN writes its staging image only in phase B: the wait moves there, so a single staging buffer's drain
        // (the previous tile's TMA store reading it) overlaps phase A
        if (EPI != LT_RES_LN) mbar_wait(&bars->stg_empty[buf], ((it / NSTG) & 1u) ^ 1u);
        unsigned char* stg0 = smem + Lay::kStg + buf * kNOut * kImgBytes;
        unsigned char* stg1 = stg0 + kImgBytes;  // only when kNOut == 2
        const long long tok = (long long)tile * 128 + r;

        if (EPI == LT_BIAS || EPI == LT_GELU || EPI == LT_RES_LN) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            float b[8];
            ld8f(p.bias + n * 128 + c0 + g * 8, b);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[g * 8 + j] += b[j];
          }
        }
        if (EPI == LT_GELU) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            *reinterpret_cast<uint4*>(stg0 + img_off(r, cq * 4 + g)) = pack8f(v + g * 8);
            float h[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) h[j] = gelu_erf(v[g * 8 + j]);
            *reinterpret_cast<uint4*>(stg1 + img_off(r, cq * 4 + g)) = pack8f(h);
          }
        } else if (EPI == LT_GELU_BWD) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const uint4 pre = *reinterpret_cast<const uint4*>(eimg + img_off(r, cq * 4 + g));
            float x[8];
            unpack8f(pre, x);
#pragma unroll
            for (int j = 0; j < 8; ++j) x[j] = v[g * 8 + j] * gelu_erf_grad(x[j]);
            *reinterpret_cast<uint4*>(stg0 + img_off(r, cq * 4 + g)) = pack8f(x);
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars->e_empty[se]);  // pre-activation image consumed
        } else if (EPI == LT_RES_LN) {
          // phase A (row x 32 columns per thread, the TMEM layout): z = dropout(acc + bias) + residual, rounded to
          // bf16 and written over the residual image in place
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            if (p.dropout_p > 0.f) {
              const uint64_t idx = (uint64_t)tok * 128u + (uint64_t)(c0 + g * 8);
              const uint32_t k8 = dropout_keep8(p.seed, p.site, idx, p.dropout_p);
#pragma unroll
              for (int j = 0; j < 8; ++j) v[g * 8 + j] = (k8 >> j) & 1u ? v[g * 8 + j] * ks : 0.f;
            }
            uint4* zp = reinterpret_cast<uint4*>(eimg + img_off(r, cq * 4 + g));
            float x[8];
            unpack8f(*zp, x);
#pragma unroll
            for (int j = 0; j < 8; ++j) x[j] += v[g * 8 + j];
            *zp = pack8f(x);
          }
          named_bar_sync(1, 32 * kEpiWarps);
          mbar_wait(&bars->stg_empty[buf], ((it / NSTG) & 1u) ^ 1u);
          // phase B (half-warp per row, 8 rows per warp): LayerNorm over the bf16-rounded z (exactly what the backward
          // pass re-reads) with the row statistics in four shuffles -- one block barrier per tile instead of an
          // all-to-all exchange of partial sums, and the optional fp32 copy leaves as 512-byte-contiguous rows
          const int hl = lane & 15, sub = lane >> 4;
#pragma unroll
          for (int i2 = 0; i2 < 4; ++i2) {
            const int row = (warp - 3) * 8 + i2 * 2 + sub;
            float zf[8];
            unpack8f(*reinterpret_cast<const uint4*>(eimg + img_off(row, hl)), zf);
            float sum = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) sum += zf[j];
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            const float mean = sum * (1.f / 128.f);
            float q = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) { zf[j] -= mean; q = fmaf(zf[j], zf[j], q); }
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
            const float rstd = rsqrtf(q * (1.f / 128.f) + p.ln_eps);
            float y[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) y[j] = zf[j] * rstd * ln_gm[j] + ln_bt[j];
            *reinterpret_cast<uint4*>(stg0 + img_off(row, hl)) = pack8f(y);
            const long long trow = (long long)tile * 128 + row;
            if (p.out_f32 != nullptr && trow < p.T) {
              float* o32 = p.out_f32 + trow * 128 + hl * 8;
              *reinterpret_cast<float4*>(o32) = make_float4(y[0], y[1], y[2], y[3]);
              *reinterpret_cast<float4*>(o32 + 4) = make_float4(y[4], y[5], y[6], y[7]);
            }
          }
        } else {  // LT_BIAS, LT_PLAIN
#pragma unroll
          for (int g = 0; g < 4; ++g) *reinterpret_cast<uint4*>(stg0 + img_off(r, cq * 4 + g)) = pack8f(v + g * 8);
        }
        // hand the staged block to the store warp
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->stg_full[buf]);
      }
    }
  } else if (DW) {
    // ===================== column sums of dY, then the dW flush =====================
    const int t = threadIdx.x - kLtThreads;   // 0..127
    const int cg = t & 15, rg = t >> 4;       // 8 columns x 16 rows per thread
    float cs[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) cs[j] = 0.f;
    uint32_t ia = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++ia) {
      const int s = ia % SA;
      mbar_wait(&bars->a_full[s], (ia / SA) & 1u);
      if (p.dbias) {
        const unsigned char* img = smem + Lay::kA + s * kImgBytes;
#pragma unroll 4
        for (int i = 0; i < 16; ++i) {
          float f[8];
          unpack8f(*reinterpret_cast<const uint4*>(img + img_off(rg * 16 + i, cg)), f);
#pragma unroll
          for (int j = 0; j < 8; ++j) cs[j] += f[j];
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->a_empty[s]);
    }
    mbar_wait(&bars->dw_done, 0u);  // every MMA of this CTA has completed: the X slots are idle, dW is final
    tcgen05_fence_after();
    if (p.dbias) {
      float* red = reinterpret_cast<float*>(smem + Lay::kX);  // [8][128]
#pragma unroll
      for (int j = 0; j < 8; ++j) red[rg * 128 + cg * 8 + j] = cs[j];
      named_bar_sync(2, 32 * kDwWarps);
      float v = 0.f;
#pragma unroll
      for (int g = 0; g < 8; ++g) v += red[g * 128 + t];
      if (v != 0.f) atomicAdd(p.dbias + t, v);
    }
    {
      const int quarter = warp & 3;
      const int r = quarter * 32 + lane;
      float* dst = p.dw + (long long)r * p.ld_dw;
      // every CTA adds its partial dW into the same 64 KB: the L2 atomic units serialise per address, so CTAs walk
      // the column chunks in rotated order and rarely meet on one
#pragma unroll 1
      for (int ci = 0; ci < 4; ++ci) {
        const int c0 = ((ci + (int)blockIdx.x) & 3) * 32;
        uint32_t acc[32];
        tmem_ld_x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(kSlots * 128 + c0), acc);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + c0 + j), "f"(__uint_as_float(acc[j])),
                       "f"(__uint_as_float(acc[j + 1])), "f"(__uint_as_float(acc[j + 2])), "f"(__uint_as_float(acc[j + 3]))
                       : "memory");
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
  }
}

// ---------------------------------------------------------------------------------------------------
// dW / dbias
// ---------------------------------------------------------------------------------------------------
struct DwBars {
  uint64_t x_full[4], x_empty[4];
  uint64_t dy_full[8], dy_empty[8];
  uint64_t done;
  uint32_t tmem_base;
};

struct DwParams {
  int T, num_tiles;
  float* dw;
  long long ld_dw;
  float* dbias;
};

constexpr int kDwBatchMax = 8;

// Up to kDwBatchMax independent dW problems of one shape in ONE launch (e.g. the Q/K/V/C weight gradients of every
// layer, deferred to the end of the backward pass): CTA b works on problem b % n with the CTAs b, b + n, b + 2n ...
// Each dW then receives ~gridDim / n partial sums instead of gridDim, and the TMEM flush -- 148 CTAs adding 256 KB
// each into the same 256 KB through the L2 atomic units, ~24 us per launch -- is paid once instead of once per layer.
struct DwBatch {
  CUtensorMap dy[kDwBatchMax];
  CUtensorMap x[kDwBatchMax];
  DwParams p[kDwBatchMax];
  int n;
};

template <int NC, int SX, int SD>
__global__ void __launch_bounds__(192, 1) dw_tile_kernel(const __grid_constant__ DwBatch batch) {
  const int prob = (int)blockIdx.x % batch.n;
  const int cta = (int)blockIdx.x / batch.n;                                   // index among this problem's CTAs
  const int ncta = ((int)gridDim.x - prob + batch.n - 1) / batch.n;           // how many CTAs share this problem
  const CUtensorMap& tm_dy = batch.dy[prob];
  const CUtensorMap& tm_x = batch.x[prob];
  const DwParams& p = batch.p[prob];
  constexpr int kX = 0;
  constexpr int kDY = SX * kImgBytes;
  constexpr int kRed = kDY + SD * kImgBytes;       // [8][128] floats for the column-sum reduction
  constexpr int kBar = kRed + 8 * 128 * 4;
  constexpr uint32_t kCols = NC * 128 <= 128 ? 128u : (NC * 128 <= 256 ? 256u : 512u);
  extern __shared__ __align__(1024) unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  DwBars* bars = reinterpret_cast<DwBars*>(smem + kBar);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_launch_dependents();
  if (threadIdx.x == 0) {
    for (int s = 0; s < 4; ++s) {
      mbar_init(&bars->x_full[s], 1);
      mbar_init(&bars->x_empty[s], 1);
    }
    for (int s = 0; s < 8; ++s) {
      mbar_init(&bars->dy_full[s], 1);
      mbar_init(&bars->dy_empty[s], 1 + 4);  // MMA commit + one arrival per column-sum warp
    }
    mbar_init(&bars->done, 1);
    fence_barrier_init();
    prefetch_tmap(&tm_dy);
    prefetch_tmap(&tm_x);
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)),
                 "r"(kCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  const bool has_tiles = cta < p.num_tiles;
  pdl_wait();

  if (warp == 0) {
    if (lane == 0) {
      uint32_t ix = 0, id = 0;
      for (int tile = cta; tile < p.num_tiles; tile += ncta, ++ix) {
        {
          const int s = ix % SX;
          mbar_wait(&bars->x_empty[s], ((ix / SX) & 1u) ^ 1u);
          mbar_arrive_expect_tx(&bars->x_full[s], (uint32_t)kImgBytes);
          const uint32_t dst = smem_u32(smem + kX + s * kImgBytes);
          tma_load_2d(dst, &tm_x, &bars->x_full[s], 0, tile * 128);
          tma_load_2d(dst + kSlabBytes, &tm_x, &bars->x_full[s], 64, tile * 128);
        }
        for (int n = 0; n < NC; ++n, ++id) {
          const int s = id % SD;
          mbar_wait(&bars->dy_empty[s], ((id / SD) & 1u) ^ 1u);
          mbar_arrive_expect_tx(&bars->dy_full[s], (uint32_t)kImgBytes);
          const uint32_t dst = smem_u32(smem + kDY + s * kImgBytes);
          tma_load_2d(dst, &tm_dy, &bars->dy_full[s], n * 128, tile * 128);
          tma_load_2d(dst + kSlabBytes, &tm_dy, &bars->dy_full[s], n * 128 + 64, tile * 128);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(true, true);
      uint32_t ix = 0, id = 0;
      bool first = true;
      for (int tile = cta; tile < p.num_tiles; tile += ncta, ++ix) {
        const int sx = ix % SX;
        mbar_wait(&bars->x_full[sx], (ix / SX) & 1u);
        const uint32_t x_img = smem_u32(smem + kX + sx * kImgBytes);
        for (int n = 0; n < NC; ++n, ++id) {
          const int s = id % SD;
          mbar_wait(&bars->dy_full[s], (id / SD) & 1u);
          tcgen05_fence_after();
          const uint32_t dy_img = smem_u32(smem + kDY + s * kImgBytes);
          // dW_n[128 x 128] += dY_n^T (M = dY columns, K = tokens) * X (K = tokens, N = X columns)
          mma_128x128x128(tmem_base + (uint32_t)n * 128u, dy_img, true, x_img, true, idesc, !first);
          umma_commit(&bars->dy_empty[s]);
        }
        umma_commit(&bars->x_empty[sx]);
        first = false;
      }
      umma_commit(&bars->done);
    }
  } else {
    // ===================== column sums of dY, then the TMEM flush =====================
    const int t = threadIdx.x - 64;        // 0..127
    const int cg = t & 15, rg = t >> 4;    // 8 columns x 16 rows per thread
    float cs[NC][8];
#pragma unroll
    for (int n = 0; n < NC; ++n)
#pragma unroll
      for (int j = 0; j < 8; ++j) cs[n][j] = 0.f;
    uint32_t id = 0;
    for (int tile = cta; tile < p.num_tiles; tile += ncta) {
#pragma unroll
      for (int n = 0; n < NC; ++n, ++id) {
        const int s = id % SD;
        mbar_wait(&bars->dy_full[s], (id / SD) & 1u);
        if (p.dbias) {
          const unsigned char* img = smem + kDY + s * kImgBytes;
#pragma unroll 4
          for (int i = 0; i < 16; ++i) {
            const uint4 u = *reinterpret_cast<const uint4*>(img + img_off(rg * 16 + i, cg));
            float f[8];
            unpack_bf16x2(u.x, f[0], f[1]); unpack_bf16x2(u.y, f[2], f[3]);
            unpack_bf16x2(u.z, f[4], f[5]); unpack_bf16x2(u.w, f[6], f[7]);
#pragma unroll
            for (int j = 0; j < 8; ++j) cs[n][j] += f[j];
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->dy_empty[s]);
      }
    }
    if (p.dbias && has_tiles) {
      float* red = reinterpret_cast<float*>(smem + kRed);
#pragma unroll
      for (int n = 0; n < NC; ++n) {
#pragma unroll
        for (int j = 0; j < 8; ++j) red[rg * 128 + cg * 8 + j] = cs[n][j];
        named_bar_sync(1, 128);
        float v = 0.f;
#pragma unroll
        for (int g = 0; g < 8; ++g) v += red[g * 128 + t];
        if (v != 0.f) atomicAdd(p.dbias + n * 128 + t, v);
        named_bar_sync(1, 128);
      }
    }
    if (has_tiles) {
      mbar_wait(&bars->done, 0u);
      tcgen05_fence_after();
      const int quarter = warp & 3;
      const int r = quarter * 32 + lane;
      // rotated chunk order per CTA (see the fused dX + dW flush): 148 CTAs add into the same NC * 64 KB
#pragma unroll 1
      for (int ci = 0; ci < NC * 4; ++ci) {
        const int ch = (ci + cta) % (NC * 4);
        const int n = ch >> 2, c0 = (ch & 3) * 32;
        float* dst = p.dw + (long long)(n * 128 + r) * p.ld_dw;
        uint32_t acc[32];
        tmem_ld_x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(n * 128 + c0), acc);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          red_add_v4(dst + c0 + j, __uint_as_float(acc[j]), __uint_as_float(acc[j + 1]), __uint_as_float(acc[j + 2]),
                     __uint_as_float(acc[j + 3]));
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kCols));
  }
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
template <int NC, int KC, bool B_MN, int EPI, int SA, int SE, int NSTG, int SX = 0>
static int launch_lt(const pmgt_linear_tile_args* a, cudaStream_t st) {
  using Lay = LtLayout<NC, KC, EPI, SA, SE, NSTG, SX>;
  static_assert(Lay::kTotal <= 232448, "shared memory budget (227 KiB)");
  auto kern = linear_tile_kernel<NC, KC, B_MN, EPI, SA, SE, NSTG, SX>;
  static unsigned long long configured = 0;
  if (first_use_on_device(configured)) {
    PMGT_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Lay::kTotal));
  }
  CUtensorMap tx, tw, te, to, ta;
  memset(&te, 0, sizeof(te));
  memset(&ta, 0, sizeof(ta));
  int rc;
  if ((rc = make_tmap(&tx, a->x, a->K, a->T, a->ldx, 64, 128))) return rc;
  // weights: w_mn = 0 -> [N][K]; w_mn = 1 -> [K][N]  (rows x inner)
  if ((rc = B_MN ? make_tmap(&tw, a->w, a->N, a->K, a->ldw, 64, 128) : make_tmap(&tw, a->w, a->K, a->N, a->ldw, 64, 128)))
    return rc;
  if ((rc = make_tmap(&to, a->out, a->N, a->T, a->ldo, 64, 128))) return rc;
  if (Lay::kHasE && (rc = make_tmap(&te, a->e_in, a->N, a->T, a->ld_e, 64, 128))) return rc;
  if ((EPI == LT_GELU || EPI == LT_RES_LN) && (rc = make_tmap(&ta, a->aux_out, a->N, a->T, a->ld_aux_out, 64, 128)))
    return rc;
  if (SX > 0 && (rc = make_tmap(&ta, a->dw_x, 128, a->T, a->ld_dw_x, 64, 128))) return rc;  // the dW operand stream
  LtParams p;
  p.T = (int)a->T;
  p.num_tiles = (int)((a->T + 127) / 128);
  p.N = a->N;
  p.bias = a->bias; p.ln_g = a->ln_g; p.ln_b = a->ln_b; p.ln_eps = a->ln_eps;
  p.dropout_p = a->dropout_p; p.seed = a->dropout_seed; p.site = a->dropout_site;
  p.out_f32 = a->out_f32;
  p.dw = a->dw; p.ld_dw = a->ld_dw; p.dbias = a->dbias;
  p.reverse = next_tile_order();
  int grid = num_sms();
  if (grid > p.num_tiles) grid = p.num_tiles;
  PMGT_CHECK_CUDA(launch_kernel(true, kern, dim3(grid), dim3(kLtThreads + (SX > 0 ? 32 * kDwWarps : 0)), Lay::kTotal, st, tx,
                                tw, te, to, ta, p));
  return PMGT_OK;
}

template <int NC, int SX, int SD>
static int launch_dw(const pmgt_dw_tile_args* a, int n, cudaStream_t st) {
  constexpr int smem = (SX + SD) * kImgBytes + 8 * 128 * 4 + 256 + 1024;
  static_assert(smem <= 232448, "shared memory budget (227 KiB)");
  static_assert(sizeof(DwBatch) <= 4000, "kernel parameter space");
  auto kern = dw_tile_kernel<NC, SX, SD>;
  static unsigned long long configured = 0;
  if (first_use_on_device(configured)) {
    PMGT_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  }
  DwBatch b;
  memset(&b, 0, sizeof(b));
  b.n = n;
  long long total_tiles = 0;
  for (int i = 0; i < n; ++i) {
    int rc;
    if ((rc = make_tmap(&b.dy[i], a[i].dy, a[i].N, a[i].T, a[i].ld_dy, 64, 128))) return rc;
    if ((rc = make_tmap(&b.x[i], a[i].x, a[i].K, a[i].T, a[i].ldx, 64, 128))) return rc;
    b.p[i].T = (int)a[i].T;
    b.p[i].num_tiles = (int)((a[i].T + 127) / 128);
    b.p[i].dw = a[i].dw; b.p[i].ld_dw = a[i].ld_dw; b.p[i].dbias = a[i].dbias;
    total_tiles += b.p[i].num_tiles;
  }
  int grid = num_sms();
  if (grid > total_tiles) grid = (int)total_tiles;
  if (grid < n) grid = n;
  PMGT_CHECK_CUDA(launch_kernel(true, kern, dim3(grid), dim3(192), smem, st, b));
  return PMGT_OK;
}

}  // namespace pmgt

using namespace pmgt;

extern "C" {

int pmgt_linear_tile_supported(int64_t K, int64_t N, int w_mn, int epi) {
  if (K % 128 != 0 || N % 128 != 0 || K <= 0 || N <= 0) return 0;
  const int kc = (int)(K / 128), nc = (int)(N / 128);
  if (epi == PMGT_LT_BIAS && !w_mn) return (kc == 1 && (nc == 1 || nc == 4)) ? 1 : 0;
  if (epi == PMGT_LT_GELU && !w_mn) return (kc == 1 && nc == 1) ? 1 : 0;
  if (epi == PMGT_LT_RES_LN && !w_mn) return (kc == 1 && nc == 1) ? 1 : 0;
  if (epi == PMGT_LT_PLAIN && w_mn) return (nc == 1 && (kc == 1 || kc == 4)) ? 1 : 0;
  if (epi == PMGT_LT_GELU_BWD && w_mn) return (kc == 1 && nc == 1) ? 1 : 0;
  return 0;
}

int pmgt_linear_tile(const pmgt_linear_tile_args* a, void* stream) {
  PMGT_REQUIRE(a && a->x && a->w && a->out, "pmgt_linear_tile: null argument");
  PMGT_REQUIRE(a->T >= 0 && a->T < (1ll << 31) - 128, "pmgt_linear_tile: bad T");
  if (a->T == 0) return PMGT_OK;
  PMGT_REQUIRE(pmgt_linear_tile_supported(a->K, a->N, a->w_mn, a->epi),
               "pmgt_linear_tile: unsupported shape K=%d N=%d w_mn=%d epi=%d (use pmgt_gemm_bf16)", a->K, a->N, a->w_mn,
               a->epi);
  PMGT_REQUIRE(a->ldx % 8 == 0 && a->ldw % 8 == 0 && a->ldo % 8 == 0, "pmgt_linear_tile: row pitches must be multiples of 8");
  PMGT_REQUIRE((((uintptr_t)a->x | (uintptr_t)a->w | (uintptr_t)a->out) & 15) == 0, "pmgt_linear_tile: 16-byte alignment");
  const int kc = a->K / 128, nc = a->N / 128;
  cudaStream_t st = (cudaStream_t)stream;
  if (a->dw) {
    PMGT_REQUIRE((a->epi == PMGT_LT_PLAIN || a->epi == PMGT_LT_GELU_BWD) && a->w_mn && kc == 1 && nc == 1,
                 "pmgt_linear_tile: the fused weight gradient needs a dX call (w_mn = 1) with K = N = 128");
    PMGT_REQUIRE(a->dw_x && a->ld_dw_x % 8 == 0 && a->ld_dw % 4 == 0 && (((uintptr_t)a->dw_x | (uintptr_t)a->dw) & 15) == 0,
                 "pmgt_linear_tile: dw_x / dw alignment");
  }
  switch (a->epi) {
    case PMGT_LT_BIAS:
      PMGT_REQUIRE(a->bias, "pmgt_linear_tile: bias required");
      if (nc == 4) return launch_lt<4, 1, false, LT_BIAS, 1, 1, 2>(a, st);
      return launch_lt<1, 1, false, LT_BIAS, 3, 1, 2>(a, st);
    case PMGT_LT_GELU:
      PMGT_REQUIRE(a->bias && a->aux_out && a->ld_aux_out % 8 == 0, "pmgt_linear_tile: GELU needs bias and aux_out");
      return launch_lt<1, 1, false, LT_GELU, 2, 1, 2>(a, st);
    case PMGT_LT_RES_LN:
      PMGT_REQUIRE(a->bias && a->aux_out && a->e_in && a->ln_g && a->ln_b && a->ld_aux_out % 8 == 0 && a->ld_e % 8 == 0,
                   "pmgt_linear_tile: RES_LN needs bias, residual (e_in), z out (aux_out), ln_g, ln_b");
      PMGT_REQUIRE(a->dropout_p >= 0.f && a->dropout_p < 1.f, "pmgt_linear_tile: bad dropout_p");
      return launch_lt<1, 1, false, LT_RES_LN, 2, 3, 1>(a, st);
    case PMGT_LT_PLAIN:
      if (kc == 4) return launch_lt<1, 4, true, LT_PLAIN, 2, 1, 1>(a, st);
      if (a->dw) return launch_lt<1, 1, true, LT_PLAIN, 2, 1, 2, 2>(a, st);
      return launch_lt<1, 1, true, LT_PLAIN, 3, 1, 2>(a, st);
    case PMGT_LT_GELU_BWD:
      PMGT_REQUIRE(a->e_in && a->ld_e % 8 == 0, "pmgt_linear_tile: GELU_BWD needs the pre-activation (e_in)");
      if (a->dw) return launch_lt<1, 1, true, LT_GELU_BWD, 2, 2, 1, 1>(a, st);
      return launch_lt<1, 1, true, LT_GELU_BWD, 2, 2, 2>(a, st);
  }
  set_error("pmgt_linear_tile: unknown epilogue %d", a->epi);
  return PMGT_ERR_INVALID;
}

int pmgt_dw_tile_supported(int64_t N, int64_t K) { return (K == 128 && (N == 128 || N == 512)) ? 1 : 0; }

static int check_dw_args(const pmgt_dw_tile_args* a) {
  PMGT_REQUIRE(a && a->dy && a->x && a->dw, "pmgt_dw_tile: null argument");
  PMGT_REQUIRE(a->T > 0 && a->T < (1ll << 31) - 128, "pmgt_dw_tile: bad T");
  PMGT_REQUIRE(pmgt_dw_tile_supported(a->N, a->K), "pmgt_dw_tile: unsupported shape N=%d K=%d (use pmgt_gemm_bf16)", a->N,
               a->K);
  PMGT_REQUIRE(a->ld_dy % 8 == 0 && a->ldx % 8 == 0 && a->ld_dw % 4 == 0, "pmgt_dw_tile: row pitch alignment");
  PMGT_REQUIRE((((uintptr_t)a->dy | (uintptr_t)a->x | (uintptr_t)a->dw) & 15) == 0, "pmgt_dw_tile: 16-byte alignment");
  return PMGT_OK;
}

int pmgt_dw_tile(const pmgt_dw_tile_args* a, void* stream) {
  PMGT_REQUIRE(a, "pmgt_dw_tile: null argument");
  if (a->T == 0) return PMGT_OK;
  int rc = check_dw_args(a);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (a->N == 512) return launch_dw<4, 2, 4>(a, 1, st);
  return launch_dw<1, 3, 3>(a, 1, st);
}

int pmgt_dw_tile_batch(const pmgt_dw_tile_args* a, int n, void* stream) {
  PMGT_REQUIRE(a && n >= 1, "pmgt_dw_tile_batch: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  for (int lo = 0; lo < n; lo += kDwBatchMax) {  // more than kDwBatchMax problems: several launches
    pmgt_dw_tile_args chunk[kDwBatchMax];
    int m = 0;
    for (int i = lo; i < n && i < lo + kDwBatchMax; ++i) {
      if (a[i].T == 0) continue;
      int rc = check_dw_args(&a[i]);
      if (rc) return rc;
      PMGT_REQUIRE(m == 0 || (a[i].N == chunk[0].N && a[i].K == chunk[0].K),
                   "pmgt_dw_tile_batch: all problems of a batch must share N and K");
      chunk[m++] = a[i];
    }
    if (m == 0) continue;
    int rc = chunk[0].N == 512 ? launch_dw<4, 2, 4>(chunk, m, st) : launch_dw<1, 3, 3>(chunk, m, st);
    if (rc) return rc;
  }
  return PMGT_OK;
}

}  // extern "C"
