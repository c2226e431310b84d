// Fused post-attention blocks of one PMGT encoder layer, default width H = I = 128.  Two shapes share the kernels:
//
//   dense block (ffn = 0)  BertSelfOutput as PMGTAttention composes it (pmgt/pmgt/modeling_pmgt.py:358-375):
//       out = LayerNorm(dropout(in W2^T + b2) + res)
//   FFN block   (ffn = 1)  BertIntermediate + BertOutput in PMGTLayer.feed_forward_chunk (modeling_pmgt.py:296-325):
//       out = LayerNorm(dropout(gelu(in W1^T + b1) W2^T + b2) + in)
//
// forward   one persistent tcgen05 kernel per block (in -> out); besides `out` it saves what the backward pass needs
//           in its cheapest form: xhat (the normalised pre-affine LayerNorm value, bf16), rstd (fp32 per row), the
//           dropout keep bits (128 bits per row) and -- FFN -- h = gelu(h_pre) and gelu'(h_pre).
// backward  one persistent tcgen05 kernel per block: LayerNorm backward from (xhat, rstd, dy [+ dy_b]) straight into
//           the shared-memory operand of the dX and dW products; dW accumulators stay in tensor memory for the whole
//           kernel; d(in) leaves as one bf16 row.
//
// These kernels are bound by instruction issue in the epilogues and by HBM, never by the tensor pipe, so the design
// goal is many independent warps in different phases rather than one wide lock-step epilogue: every epilogue thread
// owns ONE token row (the TMEM 32x32b layout: lane = row) and walks its 128 columns in chunks, so LayerNorm needs no
// cross-thread exchange at all, and the stages of the block (GELU | LayerNorm; LayerNorm-backward | GELU' | d_in) run
// in different warp groups on different tiles at the same time, handing tiles over through mbarriers.  Saved rows are
// written and read as 32-byte vectors per thread (full sectors) directly from / to registers.
#include "tile.cuh"

namespace pmgt {

constexpr int kBlkFwdThreads = 640;   // warps 0-3: TMA producer | MMA issuer | h store | idle; 4-19: epilogue groups
constexpr int kBlkBwdThreads = 512;   // warps 0-3: TMA producer | MMA issuer | idle | idle; 4-11: E2; 12-15: E3 + E4

struct BlkParams {
  int T, num_tiles, reverse, save;
  const float* b1;
  const float* b2;
  const float* ln_g;
  const float* ln_b;
  float ln_eps, dropout_p;
  uint64_t seed;
  uint32_t site;
  const uint16_t* res;
  long long ld_res;
  uint16_t* out;
  long long ld_out;
  float* out_f32;
  uint16_t* gp;
  uint16_t* xhat;
  long long ld_save;
  float* rstd;
  // backward
  const uint16_t* dy;
  long long ld_dy;
  const uint16_t* dy_b;
  long long ld_dy_b;
  uint16_t* dx;
  long long ld_dx;
  uint16_t* dz;
  long long ld_dz;
  float *dw1, *dw2, *db1, *db2, *dg, *dbeta;
  unsigned long long* trace;   // diagnostics: per-stage clock64() stamps of CTA 0 (pmgt_block_set_trace), else nullptr
};

// trace record: [role 0..7][tile 0..31][event 0..3]
__device__ __forceinline__ void blk_trace(const BlkParams& p, int role, int n, int ev) {
  if (p.trace != nullptr && blockIdx.x == 0 && n < 32) p.trace[(role * 32 + n) * 4 + ev] = (unsigned long long)clock64();
}

__device__ __forceinline__ int blk_tile(const BlkParams& p, int lt) { return p.reverse ? p.num_tiles - 1 - lt : lt; }

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

__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st_x8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}

__device__ __forceinline__ void tmem_st_x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(
          taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
// tcgen05.wait::ld that names the registers of an earlier (software-pipelined) tcgen05.ld as in/out operands: the
// compiler then keeps every use of them behind the wait
__device__ __forceinline__ void tmem_wait_ld16(uint32_t (&r)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :
               : "memory");
}

// dropout multipliers (1 / (1 - p) or 0) of the 8 elements idx8 .. idx8 + 7 as four packed pairs; thr16 =
// dropout_threshold(p) << 16.  Same stream as dropout_keep8 (common.cuh).
__device__ __forceinline__ void dropout_factors8(uint64_t seed, uint32_t site, uint64_t idx8, uint32_t thr16, float ks, f32x2 (&f)[4]) {
  uint32_t w[4];
  dropout_words8(seed, site, idx8, w);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float lo = (w[i] << 16) >= thr16 ? ks : 0.f;
    const float hi = w[i] >= thr16 ? ks : 0.f;
    f[i] = pk2(lo, hi);
  }
}

// 32-byte global accesses: one full sector per thread
__device__ __forceinline__ void ldg256(const void* p, uint32_t* r) {
  asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "l"(p));
}
// rows written earlier by another warp of this CTA (the dz scratch): read at L2
__device__ __forceinline__ void ldg256_cg(const void* p, uint32_t* r) {
  asm volatile("ld.global.cg.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "l"(p)
               : "memory");
}
__device__ __forceinline__ void stg256(void* p, const uint32_t* r) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(r[0]), "r"(r[1]), "r"(r[2]),
               "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void zero8(uint32_t* r) {
#pragma unroll
  for (int i = 0; i < 8; ++i) r[i] = 0u;
}

// Column sums over the 32 rows a warp holds (lane = row, v[j] = column j): transpose-reduce butterfly, 31 shuffles.
// Returns the sum of column `lane`.  Destroys v.
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

// ---------------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------------
template <bool FFN>
struct BlkFwdLayout {
  static constexpr int kSlots = FFN ? 2 : 4;                       // tiles in flight per stage
  static constexpr int kW1 = 0;                                     // FFN only
  static constexpr int kW2 = FFN ? kImgBytes : 0;
  static constexpr int kA = kW2 + kImgBytes;                        // kSlots input images
  static constexpr int kHH = kA + kSlots * kImgBytes;               // FFN: 2 gelu-output images
  static constexpr int kPar = kHH + (FFN ? 2 * kImgBytes : 0);      // b1 | b2 | gamma | beta  (fp32 [128] each)
  static constexpr int kBar = kPar + 2048;
  static constexpr int kTotal = kBar + 256 + 1024;
};

struct BlkFwdBars {
  uint64_t w_full;
  uint64_t a_full[4], a_empty[4];
  uint64_t sh_full[2], sh_empty[2];
  uint64_t hh_full[2], hh_empty[2];
  uint64_t sz_full[4], sz_empty[4];
  uint32_t tmem_base;
};
static_assert(sizeof(BlkFwdBars) <= 256, "barrier block");

// LayerNorm stage of one tile row: z = bf16(dropout(acc + b2) + res) is formed 32 columns at a time and parked back
// in the accumulator's own TMEM columns as packed bf16 (columns [0, 64) of the slot once the pass is over) while the
// shifted sums for the row statistics build up; the second pass reads the packed row back, normalises and stores.
// The chunk loops are NOT unrolled: four stages of this kernel run at the same time on every scheduler, and their
// combined code has to stay inside the instruction cache (the fully unrolled version was 119 KB and stalled on
// instruction fetch).
__device__ __forceinline__ void blk_ln_row(const BlkParams& p, const float* sp, uint32_t taddr, long long tok, bool valid,
                                           uint64_t* full, uint32_t full_par, uint64_t* empty, int lane, int trole, int tn) {
  const float* sb2 = sp + 128;
  const float* sg = sp + 256;
  const float* sbt = sp + 384;
  const bool drop = p.dropout_p > 0.f;
  const bool save = p.save != 0;
  const float ks = drop ? 1.f / (1.f - p.dropout_p) : 1.f;
  const uint32_t thr16 = dropout_threshold(p.dropout_p) << 16;
  const uint16_t* resrow = p.res + tok * p.ld_res;
  uint32_t rn[16];   // residual chunk requested one chunk ahead of its use
  if (valid) { ldg256(resrow, rn); ldg256(resrow + 16, rn + 8); } else { zero8(rn); zero8(rn + 8); }
  if (trole >= 0) blk_trace(p, trole, tn, 0);
  mbar_wait(full, full_par);
  tcgen05_fence_after();
  if (trole >= 0) blk_trace(p, trole, tn, 1);
  f32x2 shift = 0ull, sum = 0ull, sq = 0ull;
  // One 16-column half of a chunk: z = dropout(acc + b2) + res, statistics, z (fp32) back into the accumulator's columns.
  // The tensor-memory load of the NEXT half is in flight meanwhile (a tcgen05.ld takes ~1000 cycles under load).
  auto half1 = [&](int col0, uint32_t (&acc)[16], const uint32_t* rq, bool first) {
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      const int col = col0 + g * 8;
      const ulonglong2 ba = *reinterpret_cast<const ulonglong2*>(sb2 + col);
      const ulonglong2 bb = *reinterpret_cast<const ulonglong2*>(sb2 + col + 4);
      f32x2 v[4] = {add2(pk2u(acc[g * 8 + 0], acc[g * 8 + 1]), ba.x), add2(pk2u(acc[g * 8 + 2], acc[g * 8 + 3]), ba.y),
                    add2(pk2u(acc[g * 8 + 4], acc[g * 8 + 5]), bb.x), add2(pk2u(acc[g * 8 + 6], acc[g * 8 + 7]), bb.y)};
      if (drop) {
        f32x2 f[4];
        dropout_factors8(p.seed, p.site, (uint64_t)tok * 128u + (uint64_t)col, thr16, ks, f);
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = mul2(v[j], f[j]);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const f32x2 z = add2(v[j], unpack2_bf16(rq[g * 4 + j]));
        if (first && g == 0 && j == 0) { float z0, z1; up2(z, z0, z1); shift = pk1(-z0); }
        const f32x2 d = add2(z, shift);
        sum = add2(sum, d);
        sq = fma2(d, d, sq);
        up2u(z, acc[g * 8 + 2 * j], acc[g * 8 + 2 * j + 1]);
      }
    }
    tmem_st_x16(taddr + col0, acc);
  };
  uint32_t accA[16], accB[16];
  tmem_ld_x16(taddr, accA);
#pragma unroll 1
  for (int c = 0; c < 4; ++c) {
    uint32_t rq[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) rq[i] = rn[i];
    if (c < 3) {
      if (valid) { ldg256(resrow + (c + 1) * 32, rn); ldg256(resrow + (c + 1) * 32 + 16, rn + 8); }
    }
    tmem_wait_ld16(accA);
    tmem_ld_x16(taddr + c * 32 + 16, accB);
    half1(c * 32, accA, rq, c == 0);
    tmem_wait_ld16(accB);
    if (c < 3) tmem_ld_x16(taddr + (c + 1) * 32, accA);
    half1(c * 32 + 16, accB, rq + 8, false);
  }
  tmem_wait_st();
  if (trole >= 0) blk_trace(p, trole, tn, 2);
  float sum0, sum1, sq0, sq1, sh0, sh1;
  up2(sum, sum0, sum1);
  up2(sq, sq0, sq1);
  up2(shift, sh0, sh1);
  const float ms = (sum0 + sum1) * (1.f / 128.f);
  const float var = fmaxf((sq0 + sq1) * (1.f / 128.f) - ms * ms, 0.f);
  const float rstd = rsqrtf(var + p.ln_eps);
  const f32x2 rs2 = pk1(rstd), nm2 = pk1(-(ms - sh0) * rstd);
  uint16_t* xrow = p.xhat + tok * p.ld_save;
  uint16_t* orow = p.out + tok * p.ld_out;
  // second pass, 16 columns per step, the next step's z already on its way from tensor memory
  auto half2 = [&](int col0, const uint32_t (&zz)[16]) {
    uint32_t xo[8], oo[8];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int col = col0 + q * 4;
      const ulonglong2 g4 = *reinterpret_cast<const ulonglong2*>(sg + col);
      const ulonglong2 b4 = *reinterpret_cast<const ulonglong2*>(sbt + col);
      const f32x2 x01 = fma2(pk2u(zz[q * 4], zz[q * 4 + 1]), rs2, nm2);
      const f32x2 x23 = fma2(pk2u(zz[q * 4 + 2], zz[q * 4 + 3]), rs2, nm2);
      const f32x2 o01 = fma2(x01, g4.x, b4.x), o23 = fma2(x23, g4.y, b4.y);
      xo[q * 2] = pack2_bf16(x01);
      xo[q * 2 + 1] = pack2_bf16(x23);
      oo[q * 2] = pack2_bf16(o01);
      oo[q * 2 + 1] = pack2_bf16(o23);
      if (p.out_f32 != nullptr && valid) {
        ulonglong2 o4; o4.x = o01; o4.y = o23;
        *reinterpret_cast<ulonglong2*>(p.out_f32 + tok * 128 + col) = o4;
      }
    }
    if (valid) {
      if (save) stg256(xrow + col0, xo);
      stg256(orow + col0, oo);
    }
  };
  tmem_ld_x16(taddr, accA);
#pragma unroll 1
  for (int c = 0; c < 4; ++c) {
    tmem_wait_ld16(accA);
    tmem_ld_x16(taddr + c * 32 + 16, accB);
    half2(c * 32, accA);
    tmem_wait_ld16(accB);
    if (c < 3) {
      tmem_ld_x16(taddr + (c + 1) * 32, accA);
    } else {   // the slot may be overwritten by the product of the tile after next
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(empty);
    }
    half2(c * 32 + 16, accB);
  }
  if (trole >= 0) blk_trace(p, trole, tn, 3);
  if (valid && save) p.rstd[tok] = rstd;
}

template <bool FFN>
__global__ void __launch_bounds__(kBlkFwdThreads, 1)
blk_fwd_kernel(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_w1,
               const __grid_constant__ CUtensorMap tm_w2, const __grid_constant__ CUtensorMap tm_h, const BlkParams p) {
  using Lay = BlkFwdLayout<FFN>;
  constexpr int NS = Lay::kSlots;
  extern __shared__ __align__(1024) unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  BlkFwdBars* bars = reinterpret_cast<BlkFwdBars*>(smem + Lay::kBar);
  float* sp = reinterpret_cast<float*>(smem + Lay::kPar);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool save = p.save != 0;
  pdl_launch_dependents();

  if (threadIdx.x == 0) {
    mbar_init(&bars->w_full, 1);
    for (int s = 0; s < 4; ++s) {
      mbar_init(&bars->a_full[s], 1);
      mbar_init(&bars->a_empty[s], 1);
      mbar_init(&bars->sz_full[s], 1);
      mbar_init(&bars->sz_empty[s], 4);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bars->sh_full[s], 1);
      mbar_init(&bars->sh_empty[s], 4);
      mbar_init(&bars->hh_full[s], 4);
      mbar_init(&bars->hh_empty[s], save ? 2 : 1);   // second product done (+ the store of h has read the image)
    }
    fence_barrier_init();
    prefetch_tmap(&tm_in);
    prefetch_tmap(&tm_w2);
    if (FFN) {
      prefetch_tmap(&tm_w1);
      if (save) prefetch_tmap(&tm_h);
    }
    // parameters are never written by the preceding kernels of the chain: request them before the dependency wait
    mbar_arrive_expect_tx(&bars->w_full, (FFN ? 2u : 1u) * kImgBytes);
    if (FFN) load_image(smem_u32(smem + Lay::kW1), &tm_w1, &bars->w_full, 0, 0);
    load_image(smem_u32(smem + Lay::kW2), &tm_w2, &bars->w_full, 0, 0);
  }
  if (threadIdx.x >= 128 && threadIdx.x < 256) {
    const int c = threadIdx.x - 128;
    sp[c] = FFN ? p.b1[c] : 0.f;
    sp[128 + c] = p.b2[c];
    sp[256 + c] = p.ln_g[c];
    sp[384 + c] = p.ln_b[c];
  }
  if (warp == 1) tmem_alloc_512(&bars->tmem_base);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  pdl_wait();
  const int n_local = (p.num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  constexpr uint32_t kSz0 = FFN ? 256u : 0u;   // first TMEM column of the LayerNorm-stage accumulators

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      for (int n = 0; n < n_local; ++n) {
        const int s = n % NS, k = n / NS;
        const int tile = blk_tile(p, blockIdx.x + n * gridDim.x);
        mbar_wait_idle(&bars->a_empty[s], (k & 1u) ^ 1u);
        mbar_arrive_expect_tx(&bars->a_full[s], (uint32_t)kImgBytes);
        load_image(smem_u32(smem + Lay::kA + s * kImgBytes), &tm_in, &bars->a_full[s], 0, tile * 128);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(false, false);
      mbar_wait_idle(&bars->w_full, 0u);
      tcgen05_fence_after();
      const uint32_t w1 = smem_u32(smem + Lay::kW1), w2 = smem_u32(smem + Lay::kW2);
      if (FFN) {
        auto mma1 = [&](int n) {   // S_h[s] = in(n) W1^T
          const int s = n & 1, k = n >> 1;
          mbar_wait_idle(&bars->a_full[s], k & 1u);
          mbar_wait_idle(&bars->sh_empty[s], (k & 1u) ^ 1u);
          tcgen05_fence_after();
          blk_trace(p, 0, n, 0);
          mma_128x128x128(tmem_base + s * 128u, smem_u32(smem + Lay::kA + s * kImgBytes), false, w1, false, idesc, false);
          umma_commit(&bars->sh_full[s]);
          umma_commit(&bars->a_empty[s]);
        };
        for (int it = 0; it < n_local + 2; ++it) {   // first products run two tiles ahead of the second ones
          if (it < n_local) mma1(it);
          if (it < 2) continue;
          const int n = it - 2;
          const int s = n & 1, k = n >> 1;
          mbar_wait_idle(&bars->hh_full[s], k & 1u);
          mbar_wait_idle(&bars->sz_empty[s], (k & 1u) ^ 1u);
          tcgen05_fence_after();
          blk_trace(p, 0, n, 1);
          mma_128x128x128(tmem_base + kSz0 + s * 128u, smem_u32(smem + Lay::kHH + s * kImgBytes), false, w2, false, idesc,
                          false);
          umma_commit(&bars->sz_full[s]);
          umma_commit(&bars->hh_empty[s]);
        }
      } else {
        for (int n = 0; n < n_local; ++n) {
          const int s = n % NS, k = n / NS;
          mbar_wait_idle(&bars->a_full[s], k & 1u);
          mbar_wait_idle(&bars->sz_empty[s], (k & 1u) ^ 1u);
          tcgen05_fence_after();
          mma_128x128x128(tmem_base + kSz0 + s * 128u, smem_u32(smem + Lay::kA + s * kImgBytes), false, w2, false, idesc,
                          false);
          umma_commit(&bars->sz_full[s]);
          umma_commit(&bars->a_empty[s]);
        }
      }
    }
  } else if (warp == 2) {
    // ===================== store warp: h = gelu(h_pre) leaves through TMA from the operand image =====================
    if (FFN && save && lane == 0) {
      for (int n = 0; n < n_local; ++n) {
        const int s = n & 1, k = n >> 1;
        const int tile = blk_tile(p, blockIdx.x + n * gridDim.x);
        mbar_wait_idle(&bars->hh_full[s], k & 1u);
        store_image(&tm_h, smem_u32(smem + Lay::kHH + s * kImgBytes), 0, tile * 128);
        tma_store_commit();
        tma_store_wait_read0();
        mbar_arrive(&bars->hh_empty[s]);
      }
      tma_store_wait_all0();
    }
  } else if (warp >= 4) {
    const int quarter = warp & 3;   // TMEM lane quarter this warp may access
    const int r = quarter * 32 + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const int ew = warp - 4;
    if (FFN && ew < 8) {
      // ===================== GELU stage: group g handles tiles n = g, g + 2, ... =====================
      const int grp = ew >> 2;
      const float* sb1 = sp;
      for (int n = grp; n < n_local; n += 2) {
        const int k = n >> 1;
        const int tile = blk_tile(p, blockIdx.x + n * gridDim.x);
        const long long tok = (long long)tile * 128 + r;
        const bool valid = tok < p.T;
        unsigned char* hh = smem + Lay::kHH + grp * kImgBytes;
        uint16_t* gprow = p.gp + tok * p.ld_save;
        if (quarter == 0 && lane == 0) blk_trace(p, 1 + grp, n, 0);
        mbar_wait(&bars->sh_full[grp], k & 1u);
        tcgen05_fence_after();
        if (quarter == 0 && lane == 0) blk_trace(p, 1 + grp, n, 1);
        // one 16-column step: erf-GELU and its derivative (gelu_parts of common.cuh, packed two elements per
        // instruction; Abramowitz & Stegun 7.1.26 with the 1/2 of Phi folded into the coefficients)
        auto gelu16 = [&](int sc, const uint32_t (&acc)[16]) {
          uint32_t gpk[8];
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            const int col = sc * 16 + g * 8;
            const ulonglong2 ba = *reinterpret_cast<const ulonglong2*>(sb1 + col);
            const ulonglong2 bb = *reinterpret_cast<const ulonglong2*>(sb1 + col + 4);
            const f32x2 b[4] = {ba.x, ba.y, bb.x, bb.y};
            uint32_t hk[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const f32x2 x = add2(pk2u(acc[g * 8 + 2 * j], acc[g * 8 + 2 * j + 1]), b[j]);
              float x0, x1;
              up2(x, x0, x1);
              const float t0 = __fdividef(1.0f, fmaf(0.2316418882f, fabsf(x0), 1.0f));   // 0.3275911 / sqrt 2
              const float t1 = __fdividef(1.0f, fmaf(0.2316418882f, fabsf(x1), 1.0f));
              float a0, a1;
              up2(mul2(mul2(x, x), pk1(-0.72134752044448170368f)), a0, a1);
              const f32x2 e = pk2(exp2f(a0), exp2f(a1));                                  // e^{-x^2/2}
              const f32x2 t = pk2(t0, t1);
              f32x2 poly = fma2(pk1(0.5f * 1.061405429f), t, pk1(0.5f * -1.453152027f));
              poly = fma2(poly, t, pk1(0.5f * 1.421413741f));
              poly = fma2(poly, t, pk1(0.5f * -0.284496736f));
              poly = fma2(poly, t, pk1(0.5f * 0.254829592f));
              const f32x2 half = mul2(mul2(poly, t), e);            // (1 - erf(|x| / sqrt 2)) / 2 = Phi(-|x|)
              uint32_t q0, q1;
              up2u(fma2(half, pk1(-1.f), pk1(0.5f)), q0, q1);       // 1/2 - Phi(-|x|) >= 0
              q0 |= __float_as_uint(x0) & 0x80000000u;              // ... with the sign of x
              q1 |= __float_as_uint(x1) & 0x80000000u;
              const f32x2 cdf = add2(pk2u(q0, q1), pk1(0.5f));      // Phi(x)
              const f32x2 pdfx = mul2(mul2(x, e), pk1(0.39894228040143267794f));   // x phi(x)
              hk[j] = pack2_bf16(mul2(x, cdf));
              gpk[g * 4 + j] = pack2_bf16(add2(cdf, pdfx));
            }
            *reinterpret_cast<uint4*>(hh + img_off(r, sc * 2 + g)) = make_uint4(hk[0], hk[1], hk[2], hk[3]);
          }
          if (save && valid) stg256(gprow + sc * 16, gpk);
        };
        const uint32_t t_h = lane_base + grp * 128u;
        uint32_t accA[16], accB[16];
        tmem_ld_x16(t_h, accA);
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {   // the next 16 columns are on their way from tensor memory while these are processed
          tmem_wait_ld16(accA);
          tmem_ld_x16(t_h + c * 32 + 16, accB);
          if (c == 0) mbar_wait(&bars->hh_empty[grp], (k & 1u) ^ 1u);
          gelu16(2 * c, accA);
          tmem_wait_ld16(accB);
          if (c < 3) {
            tmem_ld_x16(t_h + (c + 1) * 32, accA);
          } else {
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars->sh_empty[grp]);
          }
          gelu16(2 * c + 1, accB);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->hh_full[grp]);
        if (quarter == 0 && lane == 0) blk_trace(p, 1 + grp, n, 2);
      }
    } else {
      // ===================== LayerNorm stage =====================
      const int grp = FFN ? (ew - 8) >> 2 : ew >> 2;
      for (int n = grp; n < n_local; n += NS) {
        const int k = n / NS;
        const int tile = blk_tile(p, blockIdx.x + n * gridDim.x);
        const long long tok = (long long)tile * 128 + r;
        blk_ln_row(p, sp, lane_base + kSz0 + grp * 128u, tok, tok < p.T, &bars->sz_full[grp], k & 1u, &bars->sz_empty[grp],
                   lane, (quarter == 0 && lane == 0) ? 3 + grp : -1, n);
      }
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
template <bool FFN>
struct BlkBwdLayout {
  static constexpr int kW1 = 0;                                  // FFN only
  static constexpr int kW2 = FFN ? kImgBytes : 0;
  static constexpr int kH = kW2 + kImgBytes;                     // FFN: h(n); dense: the block input (ctx): MN-major operand of dW2
  static constexpr int kA = kH + kImgBytes;                      // FFN: in(n) = a: MN-major operand of dW1
  static constexpr int kDO = kA + (FFN ? kImgBytes : 0);         // 2 x d_o: dropout-masked LayerNorm input gradient
  static constexpr int kDHP = kDO + 2 * kImgBytes;               // FFN: d h_pre
  static constexpr int kPar = kDHP + (FFN ? kImgBytes : 0);      // gamma fp32 [128]
  static constexpr int kBar = kPar + 512;
  static constexpr int kTotal = kBar + 256 + 1024;
  static constexpr int kRed = kDO;                               // end-of-kernel reduction scratch (14 KiB)
};
static_assert(BlkBwdLayout<true>::kTotal <= 232448, "shared memory budget (227 KiB)");
static_assert(BlkFwdLayout<true>::kTotal <= 232448, "shared memory budget (227 KiB)");

struct BlkBwdBars {
  uint64_t w_full;
  uint64_t h_full, h_free, a_full, a_free;
  uint64_t e2_full[2], do_free[2];
  uint64_t c3[2], c5[2], slot_free[2];
  uint64_t e3_full, dhp_free;
  uint64_t dw_done;
  uint32_t tmem_base;
};
static_assert(sizeof(BlkBwdBars) <= 256, "barrier block");

// TMEM columns: S[0] = 0, S[1] = 128 (per tile parity: d_h, then d_in), dW1 = 256, dW2 = 384.
// Per tile n (parity e):
//   E2 (group e)  LayerNorm backward from (xhat, rstd, dy [+ dy_b]): dz -> global (bf16; FFN: into the dx rows, as
//                 scratch), d_o = dropout-masked dz -> DO[e]; d_gamma, d_beta, d_b2 column sums
//   MMA3  S[e] = d_o W2          MMA4  dW2 += d_o^T h            (dense block: h = the block input)
//   FFN: E3  d h_pre = S[e] * gelu' -> DHP; d_b1 column sums
//        MMA5  S[e] = d h_pre W1     MMA6  dW1 += d h_pre^T in
//   E4    d_in = S[e] (+ dz, FFN) -> global
template <bool FFN>
__global__ void __launch_bounds__(kBlkBwdThreads, 1)
blk_bwd_kernel(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_w1,
               const __grid_constant__ CUtensorMap tm_w2, const __grid_constant__ CUtensorMap tm_h, const BlkParams p) {
  using Lay = BlkBwdLayout<FFN>;
  extern __shared__ __align__(1024) unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  BlkBwdBars* bars = reinterpret_cast<BlkBwdBars*>(smem + Lay::kBar);
  float* sg = reinterpret_cast<float*>(smem + Lay::kPar);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_launch_dependents();

  if (threadIdx.x == 0) {
    mbar_init(&bars->w_full, 1);
    mbar_init(&bars->h_full, 1);
    mbar_init(&bars->h_free, 1);
    mbar_init(&bars->a_full, 1);
    mbar_init(&bars->a_free, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bars->e2_full[s], 4);
      mbar_init(&bars->do_free[s], 1);
      mbar_init(&bars->c3[s], 1);
      mbar_init(&bars->c5[s], 1);
      mbar_init(&bars->slot_free[s], 4);
    }
    mbar_init(&bars->e3_full, 4);
    mbar_init(&bars->dhp_free, 1);
    mbar_init(&bars->dw_done, 1);
    fence_barrier_init();
    prefetch_tmap(&tm_w2);
    prefetch_tmap(&tm_h);
    if (FFN) {
      prefetch_tmap(&tm_in);
      prefetch_tmap(&tm_w1);
    }
    mbar_arrive_expect_tx(&bars->w_full, (FFN ? 2u : 1u) * kImgBytes);
    if (FFN) load_image(smem_u32(smem + Lay::kW1), &tm_w1, &bars->w_full, 0, 0);
    load_image(smem_u32(smem + Lay::kW2), &tm_w2, &bars->w_full, 0, 0);
  }
  if (threadIdx.x >= 128 && threadIdx.x < 256) sg[threadIdx.x - 128] = p.ln_g[threadIdx.x - 128];
  if (warp == 1) tmem_alloc_512(&bars->tmem_base);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  pdl_wait();
  const int n_local = (p.num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const bool drop = p.dropout_p > 0.f;
  const float ks = drop ? 1.f / (1.f - p.dropout_p) : 1.f;
  const uint32_t thr16 = dropout_threshold(p.dropout_p) << 16;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      for (int n = 0; n < n_local; ++n) {
        const int row0 = blk_tile(p, blockIdx.x + n * gridDim.x) * 128;
        if (n > 0) mbar_wait_idle(&bars->h_free, (n - 1) & 1u);
        mbar_arrive_expect_tx(&bars->h_full, (uint32_t)kImgBytes);
        load_image(smem_u32(smem + Lay::kH), &tm_h, &bars->h_full, 0, row0);
        if (FFN) {
          if (n > 0) mbar_wait_idle(&bars->a_free, (n - 1) & 1u);
          mbar_arrive_expect_tx(&bars->a_full, (uint32_t)kImgBytes);
          load_image(smem_u32(smem + Lay::kA), &tm_in, &bars->a_full, 0, row0);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t id_kn = make_idesc(false, true);    // dx = dy W: activations K-major, weights MN-major
      constexpr uint32_t id_nn = make_idesc(true, true);     // dW += dy^T x: both operands MN-major
      mbar_wait_idle(&bars->w_full, 0u);
      tcgen05_fence_after();
      const uint32_t w1 = smem_u32(smem + Lay::kW1), w2 = smem_u32(smem + Lay::kW2);
      const uint32_t hi = smem_u32(smem + Lay::kH), ai = smem_u32(smem + Lay::kA), dhp = smem_u32(smem + Lay::kDHP);
      const uint32_t t_dw1 = tmem_base + 256u, t_dw2 = tmem_base + 384u;
      auto do34 = [&](int n) {
        const int e = n & 1, k = n >> 1;
        const uint32_t doi = smem_u32(smem + Lay::kDO + e * kImgBytes);
        mbar_wait_idle(&bars->e2_full[e], k & 1u);
        mbar_wait_idle(&bars->h_full, n & 1u);
        mbar_wait_idle(&bars->slot_free[e], (k & 1u) ^ 1u);
        tcgen05_fence_after();
        blk_trace(p, 0, n, 0);
        mma_128x128x128(tmem_base + e * 128u, doi, false, w2, true, id_kn, false);
        umma_commit(&bars->c3[e]);
        mma_128x128x128(t_dw2, doi, true, hi, true, id_nn, n > 0);
        umma_commit(&bars->h_free);
        umma_commit(&bars->do_free[e]);
      };
      auto do56 = [&](int n) {
        const int e = n & 1;
        mbar_wait_idle(&bars->e3_full, n & 1u);
        mbar_wait_idle(&bars->a_full, n & 1u);
        tcgen05_fence_after();
        blk_trace(p, 0, n, 1);
        mma_128x128x128(tmem_base + e * 128u, dhp, false, w1, true, id_kn, false);
        umma_commit(&bars->c5[e]);
        mma_128x128x128(t_dw1, dhp, true, ai, true, id_nn, n > 0);
        umma_commit(&bars->a_free);
        umma_commit(&bars->dhp_free);
      };
      for (int it = 0; it <= n_local; ++it) {   // products 3/4 of tile `it`, then 5/6 of tile `it - 1`... in issue order
        if (it > 0 && FFN) do56(it - 1);
        if (it < n_local) do34(it);
      }
      umma_commit(&bars->dw_done);
    }
  } else if (warp >= 4) {
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16);
    float* red = reinterpret_cast<float*>(smem + Lay::kRed);   // [3][8 warps][128] E2 sums | [4 warps][128] d_b1
    if (warp < 12) {
      // ===================== E2: LayerNorm backward, group e handles tiles n = e, e + 2, ... =====================
      const int e = (warp - 4) >> 2;
      unsigned char* doimg = smem + Lay::kDO + e * kImgBytes;
      float acc_dg[4] = {0.f, 0.f, 0.f, 0.f}, acc_dbeta[4] = {0.f, 0.f, 0.f, 0.f}, acc_db2[4] = {0.f, 0.f, 0.f, 0.f};
      for (int n = e; n < n_local; n += 2) {
        const int k = n >> 1;
        const int tile = blk_tile(p, blockIdx.x + n * gridDim.x);
        const long long tok = (long long)tile * 128 + r;
        const bool valid = tok < p.T;
        const uint16_t* xr = p.xhat + tok * p.ld_save;
        const uint16_t* dyr = p.dy + tok * p.ld_dy;
        const uint16_t* dybr = p.dy_b != nullptr ? p.dy_b + tok * p.ld_dy_b : nullptr;
        uint16_t* dzr = FFN ? p.dx + tok * p.ld_dx : p.dz + tok * p.ld_dz;
        const float rstd = valid ? __ldg(p.rstd + tok) : 0.f;

        // xhat / dy chunks (32 columns = 2 x 32 bytes each) are requested one chunk ahead of their use, packed; the
        // chunk loops stay rolled (instruction-cache footprint, see blk_ln_row)
        uint32_t xn[16], dn[16];
        auto request = [&](int c) {
          if (valid) {
            ldg256(xr + c * 32, xn);
            ldg256(xr + c * 32 + 16, xn + 8);
            ldg256(dyr + c * 32, dn);
            ldg256(dyr + c * 32 + 16, dn + 8);
          }
        };
        // second gradient term: folded into the packed dy chunk (fp32 sum, one more bf16 rounding)
        auto fold_second = [&](int c, uint32_t (&dq)[16]) {
          if (dybr != nullptr && valid) {
            uint32_t eq[16];
            ldg256(dybr + c * 32, eq);
            ldg256(dybr + c * 32 + 16, eq + 8);
#pragma unroll
            for (int i = 0; i < 16; ++i) dq[i] = pack2_bf16(add2(unpack2_bf16(dq[i]), unpack2_bf16(eq[i])));
          }
        };
        const bool tr = quarter == 0 && lane == 0;
        if (tr) blk_trace(p, 1 + e, n, 0);
        zero8(xn); zero8(xn + 8); zero8(dn); zero8(dn + 8);
        request(0);

        // ---- pass 1: row sums of dy*gamma and dy*gamma*xhat; d_gamma / d_beta column sums
        f32x2 s1 = 0ull, s2 = 0ull;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t xc[16], dc[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) { xc[i] = xn[i]; dc[i] = dn[i]; }
          request((c + 1) & 3);   // c = 3: chunk 0 again, for the second pass
          fold_second(c, dc);
          float w[32];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const ulonglong2 g4 = *reinterpret_cast<const ulonglong2*>(sg + c * 32 + q * 4);
            const f32x2 x01 = unpack2_bf16(xc[q * 2]), x23 = unpack2_bf16(xc[q * 2 + 1]);
            const f32x2 d01 = unpack2_bf16(dc[q * 2]), d23 = unpack2_bf16(dc[q * 2 + 1]);
            const f32x2 t01 = mul2(d01, g4.x), t23 = mul2(d23, g4.y);
            s1 = add2(s1, add2(t01, t23));
            s2 = fma2(t01, x01, fma2(t23, x23, s2));
            up2(mul2(d01, x01), w[q * 4], w[q * 4 + 1]);   // d_gamma terms
            up2(mul2(d23, x23), w[q * 4 + 2], w[q * 4 + 3]);
          }
          const float cg = warp_colsum32(w, lane);
#pragma unroll
          for (int i = 0; i < 16; ++i) unpack_bf16x2(dc[i], w[2 * i], w[2 * i + 1]);
          const float cb = warp_colsum32(w, lane);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            acc_dg[i] += (c == i) ? cg : 0.f;
            acc_dbeta[i] += (c == i) ? cb : 0.f;
          }
        }
        float s1a, s1b, s2a, s2b;
        up2(s1, s1a, s1b);
        up2(s2, s2a, s2b);
        const f32x2 rs2 = pk1(rstd), nm1 = pk1(-(s1a + s1b) * (1.f / 128.f) * rstd), nm2 = pk1(-(s2a + s2b) * (1.f / 128.f) * rstd);
        // ---- pass 2: dz, d_o
        if (tr) blk_trace(p, 1 + e, n, 1);
        mbar_wait(&bars->do_free[e], (k & 1u) ^ 1u);
        if (tr) blk_trace(p, 1 + e, n, 2);
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t xc[16], dc[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) { xc[i] = xn[i]; dc[i] = dn[i]; }
          if (c < 3) request(c + 1);
          fold_second(c, dc);
          float w[32];
          uint32_t dzp[16];
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            f32x2 dz[4];
#pragma unroll
            for (int h2 = 0; h2 < 2; ++h2) {
              const int q = g * 2 + h2;
              const ulonglong2 g4 = *reinterpret_cast<const ulonglong2*>(sg + c * 32 + q * 4);
              const f32x2 x01 = unpack2_bf16(xc[q * 2]), x23 = unpack2_bf16(xc[q * 2 + 1]);
              const f32x2 d01 = unpack2_bf16(dc[q * 2]), d23 = unpack2_bf16(dc[q * 2 + 1]);
              dz[h2 * 2] = fma2(x01, nm2, fma2(mul2(d01, g4.x), rs2, nm1));
              dz[h2 * 2 + 1] = fma2(x23, nm2, fma2(mul2(d23, g4.y), rs2, nm1));
              dzp[q * 2] = pack2_bf16(dz[h2 * 2]);
              dzp[q * 2 + 1] = pack2_bf16(dz[h2 * 2 + 1]);
            }
            if (drop) {
              f32x2 f[4];
              dropout_factors8(p.seed, p.site, (uint64_t)tok * 128u + (uint64_t)(c * 32 + g * 8), thr16, ks, f);
              uint32_t dop[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const f32x2 v = mul2(dz[j], f[j]);
                up2(v, w[g * 8 + 2 * j], w[g * 8 + 2 * j + 1]);
                dop[j] = pack2_bf16(v);
              }
              *reinterpret_cast<uint4*>(doimg + img_off(r, c * 4 + g)) = make_uint4(dop[0], dop[1], dop[2], dop[3]);
            } else {
#pragma unroll
              for (int j = 0; j < 4; ++j) up2(dz[j], w[g * 8 + 2 * j], w[g * 8 + 2 * j + 1]);
              *reinterpret_cast<uint4*>(doimg + img_off(r, c * 4 + g)) =
                  make_uint4(dzp[g * 4], dzp[g * 4 + 1], dzp[g * 4 + 2], dzp[g * 4 + 3]);
            }
          }
          if (valid) {
            stg256(dzr + c * 32, dzp);
            stg256(dzr + c * 32 + 16, dzp + 8);
          }
          const float cs = warp_colsum32(w, lane);
#pragma unroll
          for (int i = 0; i < 4; ++i) acc_db2[i] += (c == i) ? cs : 0.f;
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->e2_full[e]);
        if (tr) blk_trace(p, 1 + e, n, 3);
      }
      // ---- hand the column sums over (after every product has completed: the scratch aliases the d_o images)
      mbar_wait(&bars->dw_done, 0u);
      tcgen05_fence_after();
      const int w8 = warp - 4;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        red[(0 * 8 + w8) * 128 + c * 32 + lane] = acc_dg[c];
        red[(1 * 8 + w8) * 128 + c * 32 + lane] = acc_dbeta[c];
        red[(2 * 8 + w8) * 128 + c * 32 + lane] = acc_db2[c];
      }
    } else {
      // ===================== E3 + E4 (FFN) / E4 (dense block): one group, every tile =====================
      float acc_db1[4] = {0.f, 0.f, 0.f, 0.f};
      unsigned char* dhpimg = smem + Lay::kDHP;
      auto E3 = [&](int n) {
        const int e = n & 1, k = n >> 1;
        const int tile = blk_tile(p, blockIdx.x + n * gridDim.x);
        const long long tok = (long long)tile * 128 + r;
        const bool valid = tok < p.T;
        const uint16_t* gpr = p.gp + tok * p.ld_save;
        uint32_t gq[16];
        if (valid) { ldg256(gpr, gq); ldg256(gpr + 16, gq + 8); } else { zero8(gq); zero8(gq + 8); }
        if (quarter == 0 && lane == 0) blk_trace(p, 3, n, 0);
        mbar_wait(&bars->c3[e], k & 1u);
        if (n > 0) mbar_wait(&bars->dhp_free, (n - 1) & 1u);
        tcgen05_fence_after();
        if (quarter == 0 && lane == 0) blk_trace(p, 3, n, 1);
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t gn[16];
          if (c < 3) {
            if (valid) { ldg256(gpr + (c + 1) * 32, gn); ldg256(gpr + (c + 1) * 32 + 16, gn + 8); } else { zero8(gn); zero8(gn + 8); }
          }
          uint32_t acc[32];
          tmem_ld_x32(lane_base + e * 128u + c * 32, acc);
          tmem_wait_ld();
          float v[32];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float g0, g1;
            unpack_bf16x2(gq[i], g0, g1);
            v[2 * i] = __uint_as_float(acc[2 * i]) * g0;
            v[2 * i + 1] = __uint_as_float(acc[2 * i + 1]) * g1;
          }
#pragma unroll
          for (int g = 0; g < 4; ++g) *reinterpret_cast<uint4*>(dhpimg + img_off(r, c * 4 + g)) = pack8f(v + g * 8);
          const float cs = warp_colsum32(v, lane);
#pragma unroll
          for (int i = 0; i < 4; ++i) acc_db1[i] += (c == i) ? cs : 0.f;
          if (c < 3) {
#pragma unroll
            for (int i = 0; i < 16; ++i) gq[i] = gn[i];
          }
        }
        tcgen05_fence_before();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->e3_full);
        if (quarter == 0 && lane == 0) blk_trace(p, 3, n, 2);
      };
      auto E4 = [&](int n) {
        const int e = n & 1, k = n >> 1;
        const int tile = blk_tile(p, blockIdx.x + n * gridDim.x);
        const long long tok = (long long)tile * 128 + r;
        const bool valid = tok < p.T;
        uint16_t* dxr = p.dx + tok * p.ld_dx;
        // dz rows (written by the E2 group a whole pipeline stage ago) are requested one chunk ahead
        uint32_t zn[16];
        auto request = [&](int c) {
          if (FFN && valid) { ldg256_cg(dxr + c * 32, zn); ldg256_cg(dxr + c * 32 + 16, zn + 8); }
        };
        zero8(zn); zero8(zn + 8);
        request(0);
        if (quarter == 0 && lane == 0) blk_trace(p, 4, n, 0);
        mbar_wait(FFN ? &bars->c5[e] : &bars->c3[e], k & 1u);
        tcgen05_fence_after();
        if (quarter == 0 && lane == 0) blk_trace(p, 4, n, 1);
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t zq[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) zq[i] = zn[i];
          if (c < 3) request(c + 1);
          uint32_t acc[32];
          tmem_ld_x32(lane_base + e * 128u + c * 32, acc);
          tmem_wait_ld();
          if (c == 3) {
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars->slot_free[e]);
          }
          uint32_t o[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float a0 = __uint_as_float(acc[2 * i]), a1 = __uint_as_float(acc[2 * i + 1]);
            if (FFN) {
              float z0, z1;
              unpack_bf16x2(zq[i], z0, z1);
              a0 += z0;
              a1 += z1;
            }
            o[i] = pack_bf16x2(a0, a1);
          }
          if (valid) {
            stg256(dxr + c * 32, o);
            stg256(dxr + c * 32 + 16, o + 8);
          }
        }
        if (quarter == 0 && lane == 0) blk_trace(p, 4, n, 2);
      };
      if (FFN) {
        for (int it = 0; it <= n_local; ++it) {   // E3 runs one tile ahead of E4 (single call sites: code size)
          if (it < n_local) E3(it);
          if (it > 0) E4(it - 1);
        }
      } else {
        for (int n = 0; n < n_local; ++n) E4(n);
      }
      mbar_wait(&bars->dw_done, 0u);
      tcgen05_fence_after();
      if (FFN) {
#pragma unroll
        for (int c = 0; c < 4; ++c) red[(24 + quarter) * 128 + c * 32 + lane] = acc_db1[c];
      }
    }
    // ---- bias / LayerNorm parameter gradients: one atomic per column and quantity
    named_bar_sync(1, 384);
    {
      const int t = threadIdx.x - 128;   // 0..383: quantity t / 128, column t % 128
      const int qn = t >> 7, col = t & 127;
      float sum = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) sum += red[(qn * 8 + w) * 128 + col];
      float* dst = qn == 0 ? p.dg : (qn == 1 ? p.dbeta : p.db2);
      if (dst != nullptr) atomicAdd(dst + col, sum);
      if (FFN && t < 128 && p.db1 != nullptr)
        atomicAdd(p.db1 + col, red[24 * 128 + col] + red[25 * 128 + col] + red[26 * 128 + col] + red[27 * 128 + col]);
    }
    // ---- dW flush (E2 warps; rotated start per CTA: 148 CTAs add into the same 64 KB matrices)
    if (warp < 12) {
      const int e = (warp - 4) >> 2;
      if (FFN) {
        float* dst = (e == 0 ? p.dw1 : p.dw2) + (long long)r * 128;
        const uint32_t t0 = lane_base + 256u + (uint32_t)e * 128u;
#pragma unroll 1
        for (int cc = 0; cc < 4; ++cc) {
          const int c = (cc + (int)blockIdx.x) & 3;
          uint32_t acc[32];
          tmem_ld_x32(t0 + c * 32, acc);
          tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            red_add_v4(dst + c * 32 + j, __uint_as_float(acc[j]), __uint_as_float(acc[j + 1]), __uint_as_float(acc[j + 2]),
                       __uint_as_float(acc[j + 3]));
        }
      } else {
        float* dst = p.dw2 + (long long)r * 128 + e * 64;
        const uint32_t t0 = lane_base + 384u + (uint32_t)e * 64u;
#pragma unroll 1
        for (int cc = 0; cc < 2; ++cc) {
          const int c = (cc + (int)blockIdx.x) & 1;
          uint32_t acc[32];
          tmem_ld_x32(t0 + c * 32, acc);
          tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            red_add_v4(dst + c * 32 + j, __uint_as_float(acc[j]), __uint_as_float(acc[j + 1]), __uint_as_float(acc[j + 2]),
                       __uint_as_float(acc[j + 3]));
        }
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc_512(tmem_base);
  }
}

static unsigned long long* g_trace = nullptr;

static bool al32(const void* p) { return ((uintptr_t)p & 31) == 0; }

static int check_blk(const pmgt_block_args* a, bool bwd) {
  PMGT_REQUIRE(a && a->in && a->w2 && a->b2 && a->ln_g && a->ln_b, "pmgt_block: null argument");
  PMGT_REQUIRE(!a->ffn || (a->w1 && a->b1), "pmgt_block: the FFN block needs w1 / b1");
  PMGT_REQUIRE(a->T >= 0 && a->T < (1ll << 31) - 128, "pmgt_block: bad T");
  PMGT_REQUIRE(a->ld_in % 16 == 0 && al32(a->in), "pmgt_block: `in` must be 32-byte aligned with a pitch multiple of 16");
  PMGT_REQUIRE(a->ffn || (a->res && a->ld_res % 16 == 0 && al32(a->res)), "pmgt_block: the dense block needs `res` (32-byte aligned)");
  PMGT_REQUIRE((((uintptr_t)a->w1 | (uintptr_t)a->w2 | (uintptr_t)a->b1 | (uintptr_t)a->b2 | (uintptr_t)a->ln_g |
                 (uintptr_t)a->ln_b) & 15) == 0, "pmgt_block: parameter alignment");
  PMGT_REQUIRE(a->dropout_p >= 0.f && a->dropout_p < 1.f, "pmgt_block: bad dropout_p");
  const bool saved = a->xhat != nullptr;
  if (saved || bwd) {
    PMGT_REQUIRE(a->xhat && a->rstd && (!a->ffn || (a->h && a->gp)),
                 "pmgt_block: saved activations: xhat, rstd (+ h, gp for the FFN block)");
    PMGT_REQUIRE(a->ld_save % 16 == 0 && al32(a->xhat) && al32(a->h) && al32(a->gp) &&
                 ((uintptr_t)a->rstd & 3) == 0, "pmgt_block: saved activation alignment");
  }
  if (!bwd) {
    PMGT_REQUIRE(a->out && a->ld_out % 16 == 0 && al32(a->out), "pmgt_block_fwd: out alignment");
    PMGT_REQUIRE(((uintptr_t)a->out_f32 & 15) == 0, "pmgt_block_fwd: out_f32 alignment");
  } else {
    PMGT_REQUIRE(a->dy && a->dx && a->dw2 && (!a->ffn || a->dw1) && (a->ffn || a->dz), "pmgt_block_bwd: dy, dx, dw2 (+ dw1 | dz) required");
    PMGT_REQUIRE(a->ld_dy % 16 == 0 && a->ld_dx % 16 == 0 && a->ld_dy_b % 16 == 0 && a->ld_dz % 16 == 0 && al32(a->dy) &&
                 al32(a->dx) && al32(a->dy_b) && al32(a->dz) && (((uintptr_t)a->dw1 | (uintptr_t)a->dw2) & 15) == 0,
                 "pmgt_block_bwd: gradient buffer alignment");
  }
  return PMGT_OK;
}

static void fill_params(const pmgt_block_args* a, BlkParams& p) {
  memset(&p, 0, sizeof(p));
  p.T = (int)a->T;
  p.num_tiles = (int)((a->T + 127) / 128);
  p.reverse = next_tile_order();
  p.save = a->xhat != nullptr;
  p.b1 = a->b1; p.b2 = a->b2; p.ln_g = a->ln_g; p.ln_b = a->ln_b; p.ln_eps = a->ln_eps;
  p.dropout_p = a->dropout_p; p.seed = a->dropout_seed; p.site = a->dropout_site;
  p.res = a->ffn ? a->in : a->res;
  p.ld_res = a->ffn ? a->ld_in : a->ld_res;
  p.out = a->out; p.ld_out = a->ld_out; p.out_f32 = a->out_f32;
  p.gp = a->gp; p.xhat = a->xhat; p.ld_save = a->ld_save; p.rstd = a->rstd;
  p.dy = a->dy; p.ld_dy = a->ld_dy; p.dy_b = a->dy_b; p.ld_dy_b = a->ld_dy_b;
  p.dx = a->dx; p.ld_dx = a->ld_dx; p.dz = a->dz; p.ld_dz = a->ld_dz;
  p.trace = g_trace;
  p.dw1 = a->dw1; p.dw2 = a->dw2; p.db1 = a->db1; p.db2 = a->db2; p.dg = a->d_ln_g; p.dbeta = a->d_ln_b;
}

template <bool FFN>
static int launch_fwd(const pmgt_block_args* a, void* stream) {
  static unsigned long long configured = 0;
  if (first_use_on_device(configured))
    PMGT_CHECK_CUDA(cudaFuncSetAttribute(blk_fwd_kernel<FFN>, cudaFuncAttributeMaxDynamicSharedMemorySize, BlkFwdLayout<FFN>::kTotal));
  int rc;
  CUtensorMap ti, tw1, tw2, th;
  if ((rc = make_tmap(&ti, a->in, 128, a->T, a->ld_in, 64, 128))) return rc;
  if ((rc = make_tmap(&tw2, a->w2, 128, 128, 128, 64, 128))) return rc;
  tw1 = tw2;
  th = ti;
  if (FFN) {
    if ((rc = make_tmap(&tw1, a->w1, 128, 128, 128, 64, 128))) return rc;
    if (a->h != nullptr && (rc = make_tmap(&th, a->h, 128, a->T, a->ld_save, 64, 128))) return rc;
  }
  BlkParams p;
  fill_params(a, p);
  int grid = num_sms();
  if (grid > p.num_tiles) grid = p.num_tiles;
  PMGT_CHECK_CUDA(launch_kernel(true, blk_fwd_kernel<FFN>, dim3(grid), dim3(kBlkFwdThreads), BlkFwdLayout<FFN>::kTotal,
                                (cudaStream_t)stream, ti, tw1, tw2, th, p));
  return PMGT_OK;
}

template <bool FFN>
static int launch_bwd(const pmgt_block_args* a, void* stream) {
  static unsigned long long configured = 0;
  if (first_use_on_device(configured))
    PMGT_CHECK_CUDA(cudaFuncSetAttribute(blk_bwd_kernel<FFN>, cudaFuncAttributeMaxDynamicSharedMemorySize, BlkBwdLayout<FFN>::kTotal));
  int rc;
  CUtensorMap ti, tw1, tw2, th;
  if ((rc = make_tmap(&ti, a->in, 128, a->T, a->ld_in, 64, 128))) return rc;
  if ((rc = make_tmap(&tw2, a->w2, 128, 128, 128, 64, 128))) return rc;
  tw1 = tw2;
  th = ti;   // dense block: the dW2 operand is the block input itself
  if (FFN) {
    if ((rc = make_tmap(&tw1, a->w1, 128, 128, 128, 64, 128))) return rc;
    if ((rc = make_tmap(&th, a->h, 128, a->T, a->ld_save, 64, 128))) return rc;
  }
  BlkParams p;
  fill_params(a, p);
  int grid = num_sms();
  if (grid > p.num_tiles) grid = p.num_tiles;
  PMGT_CHECK_CUDA(launch_kernel(true, blk_bwd_kernel<FFN>, dim3(grid), dim3(kBlkBwdThreads), BlkBwdLayout<FFN>::kTotal,
                                (cudaStream_t)stream, ti, tw1, tw2, th, p));
  return PMGT_OK;
}

}  // namespace pmgt

using namespace pmgt;

extern "C" {

int pmgt_block_set_trace(void* buf) {
  g_trace = static_cast<unsigned long long*>(buf);
  return PMGT_OK;
}

int pmgt_block_fwd(const pmgt_block_args* a, void* stream) {
  int rc = check_blk(a, false);
  if (rc) return rc;
  if (a->T == 0) return PMGT_OK;
  return a->ffn ? launch_fwd<true>(a, stream) : launch_fwd<false>(a, stream);
}

int pmgt_block_bwd(const pmgt_block_args* a, void* stream) {
  int rc = check_blk(a, true);
  if (rc) return rc;
  if (a->T == 0) return PMGT_OK;
  return a->ffn ? launch_bwd<true>(a, stream) : launch_bwd<false>(a, stream);
}

}  // extern "C"
