// tcgen05 (UMMA) GEMM for sm_100a: bf16 operands, fp32 accumulators in TMEM.
//
//   D[M,N] (+)= op(A)[M,K] * op(B)[K,N]
//
// CTA tile 128 x 128, K block 64, 128B-swizzled shared-memory slabs.  Operands
// are staged either by TMA (cp.async.bulk.tensor.2d, hardware swizzle) or, when
// rows are fetched through an index vector (the multimodal feature gather of
// PMGT, pmgt/pmgt/utils.py:43-50), by 16-byte cp.async copies that write the
// same swizzle pattern by hand, so the gathered [tokens, 2304] feature matrix
// is never materialised in HBM.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator +
// single-thread MMA issuer, warps 2..5 = cp.async gather producers (if any
// operand is gathered) and then the epilogue (one accumulator row per thread,
// tcgen05.ld 32x32b).
//
// Shared-memory slab layouts (one "slab" = rows of 128 bytes = 64 bf16):
//   K-major operand  : 1 slab of 128 rows (row = M/N index, 64 K elements).
//                      UMMA desc: SWIZZLE_128B, SBO = 1024 B; K advance = +32 B.
//   MN-major operand : 2 slabs of 64 rows (row = K index, 64 M/N elements each).
//                      UMMA desc: SWIZZLE_128B, LBO = 8192 B (next 64 M/N),
//                      SBO = 1024 B (next 8 K rows); K advance = +2048 B.
#include <stdlib.h>

#include "umma.cuh"

namespace pmgt {

constexpr int BM = 128;
constexpr int BN = 128;
constexpr int BK = 64;
constexpr int kGemmThreads = 192;
constexpr int kOperandStageBytes = 128 * BK * 2;  // 16 KiB for either layout
constexpr int kTmemCols = 128;

struct GemmKernelArgs {
  int M, N, K;
  // gather sources (used when the corresponding operand is gathered)
  const uint16_t* a_src; long long lda; const long long* a_rows; long long a_src_rows;
  const uint16_t* b_src; long long ldb; const long long* b_rows; long long b_src_rows;
  void* out; long long ldo;
  const float* bias;
  const uint16_t* addend; long long ld_addend;
  uint16_t* aux; long long ld_aux;
  float alpha;
  uint32_t epi;
  int kb_per_split;  // k-blocks per blockIdx.z
  int vec32;         // bf16 out (and GELU aux) rows are 32-byte aligned: 32-byte epilogue stores
};

struct GemmSmem {
  uint64_t full[4];
  uint64_t empty[4];
  uint64_t accum_full;
  uint32_t tmem_base;
};

// Fill one 128B-swizzled slab with `nrows` gathered row segments of 64 bf16.
//   slab row r <- src[row_index(r)] [col0 .. col0+64)
template <int ROWS_PER_THREAD>
__device__ __forceinline__ void gather_slab(uint32_t slab, int t, const uint16_t* __restrict__ src, long long ld,
                                            const long long* rows /*ROWS_PER_THREAD resolved row ids, <0 = zero*/,
                                            int col0, int col_limit) {
  const int c = t & 7;
  const int col = col0 + c * 8;
#pragma unroll
  for (int i = 0; i < ROWS_PER_THREAD; ++i) {
    const int r = (t >> 3) + 16 * i;
    const long long row = rows[i];
    const bool ok = row >= 0 && col < col_limit;
    const void* g = ok ? (const void*)(src + row * ld + col) : (const void*)src;
    cp_async_16(slab + r * 128 + ((c ^ (r & 7)) << 4), g, ok ? 16u : 0u);
  }
}

__device__ __forceinline__ void st_global_256(void* ptr, const uint4& lo, const uint4& hi) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(ptr), "r"(lo.x), "r"(lo.y), "r"(lo.z), "r"(lo.w),
               "r"(hi.x), "r"(hi.y), "r"(hi.z), "r"(hi.w)
               : "memory");
}

// Epilogue of 8 accumulator columns [n, n + 8) of output row m; returns the bf16 result (the caller stores it) unless
// the output is fp32 / atomic, which is written here
__device__ __forceinline__ uint4 epilogue_group(const GemmKernelArgs& p, uint32_t epi, int m, int n, const uint32_t* r, uint4& pre_out) {
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[j]) * p.alpha;
  if (epi & PMGT_EPI_BIAS) {
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + n));
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + n + 4));
    v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
    v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
  }
  if (epi & PMGT_EPI_GELU) {   // two elements per instruction (gelu_pair): this epilogue is bound by instruction issue
    pre_out.x = pack_bf16x2(v[0], v[1]); pre_out.y = pack_bf16x2(v[2], v[3]);
    pre_out.z = pack_bf16x2(v[4], v[5]); pre_out.w = pack_bf16x2(v[6], v[7]);
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
      f32x2 h;
      gelu_pair(pk2(v[j], v[j + 1]), h, nullptr);
      up2(h, v[j], v[j + 1]);
    }
  }
  if (epi & PMGT_EPI_GELU_BWD) {
    const uint4 pre = *reinterpret_cast<const uint4*>(p.aux + (long long)m * p.ld_aux + n);
    const uint32_t pw[4] = {pre.x, pre.y, pre.z, pre.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      f32x2 h, gr;
      gelu_pair(unpack2_bf16(pw[j]), h, &gr);
      up2(mul2(pk2(v[2 * j], v[2 * j + 1]), gr), v[2 * j], v[2 * j + 1]);
    }
  }
  if (epi & PMGT_EPI_ADDEND) {
    const uint4 ad = *reinterpret_cast<const uint4*>(p.addend + (long long)m * p.ld_addend + n);
    float x[8];
    unpack_bf16x2(ad.x, x[0], x[1]); unpack_bf16x2(ad.y, x[2], x[3]);
    unpack_bf16x2(ad.z, x[4], x[5]); unpack_bf16x2(ad.w, x[6], x[7]);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] += x[j];
  }
  uint4 o = make_uint4(0u, 0u, 0u, 0u);
  if (epi & PMGT_EPI_ATOMIC) {
    float* d = reinterpret_cast<float*>(p.out) + (long long)m * p.ldo + n;   // 16-byte aligned: ldo % 4 == 0, n % 8 == 0
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(d), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(d + 4), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
  } else if (epi & PMGT_EPI_OUT_F32) {
    float* d = reinterpret_cast<float*>(p.out) + (long long)m * p.ldo + n;
    *reinterpret_cast<float4*>(d) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(d + 4) = make_float4(v[4], v[5], v[6], v[7]);
  } else {
    o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]);
    o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
  }
  return o;
}

// Epilogue of 32 accumulator columns [n_base, n_base + 32) of output row m (one row per thread).  bf16 results (and
// the GELU pre-activations) leave as 32-byte stores -- one full sector per thread and instruction -- when the rows
// are 32-byte aligned (p.vec32), else as 16-byte stores.
__device__ __forceinline__ void epilogue_chunk(const GemmKernelArgs& p, uint32_t epi, int m, int n_base, const uint32_t (&r)[32]) {
  const bool bf16_out = !(epi & (PMGT_EPI_ATOMIC | PMGT_EPI_OUT_F32));
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int n = n_base + h * 16;
    if (n >= p.N) break;
    const bool two = n + 8 < p.N;
    uint4 pre0, pre1, o0, o1 = make_uint4(0u, 0u, 0u, 0u);
    o0 = epilogue_group(p, epi, m, n, r + h * 16, pre0);
    if (two) o1 = epilogue_group(p, epi, m, n + 8, r + h * 16 + 8, pre1);
    if (epi & PMGT_EPI_GELU) {
      uint16_t* a = p.aux + (long long)m * p.ld_aux + n;
      if (two && p.vec32) st_global_256(a, pre0, pre1);
      else {
        *reinterpret_cast<uint4*>(a) = pre0;
        if (two) *reinterpret_cast<uint4*>(a + 8) = pre1;
      }
    }
    if (bf16_out) {
      uint16_t* d = reinterpret_cast<uint16_t*>(p.out) + (long long)m * p.ldo + n;
      if (two && p.vec32) st_global_256(d, o0, o1);
      else {
        *reinterpret_cast<uint4*>(d) = o0;
        if (two) *reinterpret_cast<uint4*>(d + 8) = o1;
      }
    }
  }
}

// Staged variant of epilogue_chunk for the persistent kernel: the 32 columns are written as bf16 into a 128-byte-swizzled
// shared-memory image of the output tile (and of the GELU pre-activation tile), which one thread then hands to TMA as
// [128 rows x 64 columns] boxes -- whole 128-byte lines per row instead of one 32-byte piece per thread and
// instruction (the row-per-thread stores bound the epilogue: 505 vs ~420 us for the [101376, 768] x [768, 3072] Linear).
__device__ __forceinline__ void epilogue_chunk_staged(const GemmKernelArgs& p, uint32_t epi, int m, bool row_ok, int r, int n_base,
                                                      int c_tile, const uint32_t (&acc)[32], unsigned char* stg_out,
                                                      unsigned char* stg_aux) {
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const int n = n_base + g * 8;
    uint4 pre = make_uint4(0u, 0u, 0u, 0u), o = pre;
    if (row_ok && n < p.N) o = epilogue_group(p, epi, m, n, acc + g * 8, pre);
    const int c = c_tile + g * 8;   // column inside the tile
    const uint32_t off = (uint32_t)((c >> 6) * 16384 + r * 128 + ((((c & 63) >> 3) ^ (r & 7)) << 4));
    *reinterpret_cast<uint4*>(stg_out + off) = o;
    if (epi & PMGT_EPI_GELU) *reinterpret_cast<uint4*>(stg_aux + off) = pre;
  }
}

template <bool A_MN, bool B_MN, bool GATHER_A, bool GATHER_B, int STAGES>
__global__ void __launch_bounds__(kGemmThreads)
umma_gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const GemmKernelArgs p) {
  static_assert(!(GATHER_A && A_MN), "gathered A must be K-major");
  static_assert(!(GATHER_B && !B_MN), "gathered B must be MN-major");
  extern __shared__ __align__(1024) unsigned char smem_dyn[];
  // 1024-byte alignment is required by SWIZZLE_128B
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  unsigned char* smem_a = smem;
  unsigned char* smem_b = smem + STAGES * kOperandStageBytes;
  GemmSmem* sh = reinterpret_cast<GemmSmem*>(smem + 2 * STAGES * kOperandStageBytes);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN;
  const int m0 = blockIdx.y * BM;
  const int num_kb_total = (p.K + BK - 1) / BK;
  const int kb_begin = blockIdx.z * p.kb_per_split;
  int kb_end = kb_begin + p.kb_per_split;
  if (kb_end > num_kb_total) kb_end = num_kb_total;
  const int num_kb = kb_end - kb_begin;
  if (num_kb <= 0) return;  // uniform for the whole CTA

  constexpr bool kAnyGather = GATHER_A || GATHER_B;
  constexpr uint32_t kTmaBytes = (GATHER_A ? 0u : (uint32_t)kOperandStageBytes) + (GATHER_B ? 0u : (uint32_t)kOperandStageBytes);

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&sh->full[s], 1u + (kAnyGather ? 128u : 0u));
      mbar_init(&sh->empty[s], 1u);
    }
    mbar_init(&sh->accum_full, 1u);
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sh->tmem_base)),
                 "r"((uint32_t)kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = sh->tmem_base;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      for (int i = 0; i < num_kb; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (uint32_t)(i / STAGES) & 1u;
        mbar_wait(&sh->empty[s], ph ^ 1u);
        mbar_arrive_expect_tx(&sh->full[s], kTmaBytes);
        const int k0 = (kb_begin + i) * BK;
        if (!GATHER_A) {
          const uint32_t dst = smem_u32(smem_a + s * kOperandStageBytes);
          if (!A_MN) {
            tma_load_2d(dst, &tmap_a, &sh->full[s], k0, m0);
          } else {
            tma_load_2d(dst, &tmap_a, &sh->full[s], m0, k0);
            tma_load_2d(dst + 8192, &tmap_a, &sh->full[s], m0 + 64, k0);
          }
        }
        if (!GATHER_B) {
          const uint32_t dst = smem_u32(smem_b + s * kOperandStageBytes);
          if (!B_MN) {
            tma_load_2d(dst, &tmap_b, &sh->full[s], k0, n0);
          } else {
            tma_load_2d(dst, &tmap_b, &sh->full[s], n0, k0);
            tma_load_2d(dst + 8192, &tmap_b, &sh->full[s], n0 + 64, k0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((A_MN ? 1u : 0u) << 15) |
                             ((B_MN ? 1u : 0u) << 16) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      for (int i = 0; i < num_kb; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (uint32_t)(i / STAGES) & 1u;
        mbar_wait(&sh->full[s], ph);
        tcgen05_fence_after();
        const uint32_t a_base = smem_u32(smem_a + s * kOperandStageBytes);
        const uint32_t b_base = smem_u32(smem_b + s * kOperandStageBytes);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          const uint64_t da = A_MN ? umma_desc(a_base + k * 2048, 8192, 1024) : umma_desc(a_base + k * 32, 16, 1024);
          const uint64_t db = B_MN ? umma_desc(b_base + k * 2048, 8192, 1024) : umma_desc(b_base + k * 32, 16, 1024);
          umma_bf16(tmem_base, da, db, idesc, (i > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&sh->empty[s]);  // frees the smem stage once these MMAs have read it
      }
      umma_commit(&sh->accum_full);
    }
  } else {
    // ===================== gather producers, then epilogue =====================
    const int t = threadIdx.x - 64;  // 0..127
    if (kAnyGather) {
      long long a_rows_reg[8];
      if (GATHER_A) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int m = m0 + (t >> 3) + 16 * i;
          long long row = -1;
          if (m < p.M) {
            row = p.a_rows[m];
            if (row >= p.a_src_rows) row = -1;
          }
          a_rows_reg[i] = row;
        }
      }
      constexpr int LAG = STAGES - 1;
      for (int i = 0; i < num_kb; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (uint32_t)(i / STAGES) & 1u;
        // the stage that has landed is handed to the MMA issuer BEFORE waiting for a free slot (see gather_proj.cu)
        if (i >= LAG) {
          cp_async_wait<LAG - 1>();
          fence_proxy_async_smem();
          mbar_arrive(&sh->full[(i - LAG) % STAGES]);
        }
        mbar_wait(&sh->empty[s], ph ^ 1u);
        const int k0 = (kb_begin + i) * BK;
        if (GATHER_A) {
          gather_slab<8>(smem_u32(smem_a + s * kOperandStageBytes), t, p.a_src, p.lda, a_rows_reg, k0, p.K);
        }
        if (GATHER_B) {
          long long b_rows_reg[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int kk = k0 + (t >> 3) + 16 * j;
            long long row = -1;
            if (kk < p.K) {
              row = p.b_rows[kk];
              if (row >= p.b_src_rows) row = -1;
            }
            b_rows_reg[j] = row;
          }
          const uint32_t bs = smem_u32(smem_b + s * kOperandStageBytes);
          gather_slab<4>(bs, t, p.b_src, p.ldb, b_rows_reg, n0, p.N);
          gather_slab<4>(bs + 8192, t, p.b_src, p.ldb, b_rows_reg, n0 + 64, p.N);
        }
        cp_async_commit();
      }
      // drain
      cp_async_wait<0>();
      fence_proxy_async_smem();
      for (int i = (num_kb > LAG ? num_kb - LAG : 0); i < num_kb; ++i) mbar_arrive(&sh->full[i % STAGES]);
    }

    // ---- epilogue: thread owns accumulator row (quarter*32 + lane) ----
    mbar_wait(&sh->accum_full, 0u);
    tcgen05_fence_after();
    const int quarter = warp & 3;
    const int m = m0 + quarter * 32 + lane;
    const bool row_ok = m < p.M;
    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const uint32_t epi = p.epi;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      if (n0 + c0 >= p.N) break;  // warp-uniform
      uint32_t r[32];
      tmem_ld_x32(taddr + (uint32_t)c0, r);
      tmem_wait_ld();
      if (!row_ok) continue;
      epilogue_chunk(p, epi, m, n0 + c0, r);
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)kTmemCols));
  }
}


// ---------------------------------------------------------------------------
// Persistent variant for dense operands (both by TMA) without split-K: one CTA per SM walks the output tiles
// (n fastest, so that concurrently running CTAs share A rows and W in L2), a 6-stage operand ring, two TMEM
// accumulators and EIGHT dedicated epilogue warps (two per TMEM lane quarter, 64 columns each), so that the epilogue
// of tile i runs under the main loop of tile i + 1.  With K = 768 a tile's main loop is ~3000 cycles -- about as long
// as its prologue + one-row-per-thread epilogue, which the one-tile-per-CTA kernel above can only hide behind the
// second resident CTA (0.5 PFLOP/s on the H = 768 Linears where the long-K weight-gradient GEMM reaches 1.2).
// ---------------------------------------------------------------------------
constexpr int kPersistThreads = 320;  // warp 0: TMA producer | 1: MMA issuer | 2-9: epilogue
constexpr int kPersistStages = 6;

struct PersistSmem {
  uint64_t full[kPersistStages];
  uint64_t empty[kPersistStages];
  uint64_t acc_full[2], acc_empty[2];
  uint64_t stg_empty;   // staged epilogue: the TMA store of the previous tile has finished reading the staging images
  uint32_t tmem_base;
};

// operand ring depth and staging bytes of a persistent-kernel configuration
template <int BN_, bool STG>
struct PersistCfg {
  static constexpr int kStages = !STG ? (BN_ == 128 ? kPersistStages : 4) : (BN_ == 128 ? 4 : 3);
  static constexpr int kStgOut = STG ? 128 * BN_ * 2 : 0;              // bf16 image of the output tile
  static constexpr int kStgAux = (STG && BN_ == 128) ? 128 * BN_ * 2 : 0;   // GELU pre-activation image (narrow tiles only)
  static constexpr int kSmem = kStages * (kOperandStageBytes + BN_ * BK * 2) + kStgOut + kStgAux + (int)sizeof(PersistSmem) + 1024;
};

template <bool A_MN, bool B_MN, int BN_, bool STG>  // BN_ = 128 or 256 output columns per tile; STG: TMA-store epilogue
__global__ void __launch_bounds__(kPersistThreads, 1)
umma_gemm_persist_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                         const __grid_constant__ CUtensorMap tmap_out, const __grid_constant__ CUtensorMap tmap_aux,
                         const GemmKernelArgs p) {
  using Cfg = PersistCfg<BN_, STG>;
  constexpr int STAGES = Cfg::kStages;
  constexpr int kAStage = kOperandStageBytes;         // 128 rows x 64 k
  constexpr int kBStage = BN_ * BK * 2;               // BN_ rows x 64 k (K-major) or BN_ / 64 slabs of 64 k x 64 n
  extern __shared__ __align__(1024) unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  unsigned char* smem_a = smem;
  unsigned char* smem_b = smem + STAGES * kAStage;
  unsigned char* stg_out = smem + STAGES * (kAStage + kBStage);
  unsigned char* stg_aux = stg_out + Cfg::kStgOut;
  PersistSmem* sh = reinterpret_cast<PersistSmem*>(stg_aux + Cfg::kStgAux);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ntn = (p.N + BN_ - 1) / BN_, ntm = (p.M + BM - 1) / BM;
  const int num_tiles = ntn * ntm;
  const int num_kb = (p.K + BK - 1) / BK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&sh->full[s], 1u); mbar_init(&sh->empty[s], 1u); }
    for (int s = 0; s < 2; ++s) { mbar_init(&sh->acc_full[s], 1u); mbar_init(&sh->acc_empty[s], 8u); }
    mbar_init(&sh->stg_empty, 1u);
    fence_barrier_init();
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
    if (STG) prefetch_tmap(&tmap_out);
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sh->tmem_base)), "r"((uint32_t)(2 * BN_)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = sh->tmem_base;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / ntn) * BM, n0 = (tile % ntn) * BN_;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const uint32_t s = it % STAGES;
          mbar_wait(&sh->empty[s], ((it / STAGES) & 1u) ^ 1u);
          mbar_arrive_expect_tx(&sh->full[s], (uint32_t)(kAStage + kBStage));
          const int k0 = kb * BK;
          const uint32_t da = smem_u32(smem_a + s * kAStage), db = smem_u32(smem_b + s * kBStage);
          if (!A_MN) {
            tma_load_2d(da, &tmap_a, &sh->full[s], k0, m0);
          } else {
            tma_load_2d(da, &tmap_a, &sh->full[s], m0, k0);
            tma_load_2d(da + 8192, &tmap_a, &sh->full[s], m0 + 64, k0);
          }
          if (!B_MN) {
#pragma unroll
            for (int j = 0; j < BN_ / 128; ++j) tma_load_2d(db + j * 16384, &tmap_b, &sh->full[s], k0, n0 + 128 * j);
          } else {
#pragma unroll
            for (int j = 0; j < BN_ / 64; ++j) tma_load_2d(db + j * 8192, &tmap_b, &sh->full[s], n0 + 64 * j, k0);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) |
                             ((uint32_t)(BN_ >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      uint32_t it = 0, tl = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tl) {
        const uint32_t slot = tl & 1u;
        mbar_wait(&sh->acc_empty[slot], ((tl >> 1) & 1u) ^ 1u);
        tcgen05_fence_after();
        const uint32_t tacc = tmem_base + slot * (uint32_t)BN_;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const uint32_t s = it % STAGES;
          mbar_wait(&sh->full[s], (it / STAGES) & 1u);
          tcgen05_fence_after();
          const uint32_t a_base = smem_u32(smem_a + s * kAStage);
          const uint32_t b_base = smem_u32(smem_b + s * kBStage);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t da = A_MN ? umma_desc(a_base + k * 2048, 8192, 1024) : umma_desc(a_base + k * 32, 16, 1024);
            const uint64_t db = B_MN ? umma_desc(b_base + k * 2048, 8192, 1024) : umma_desc(b_base + k * 32, 16, 1024);
            umma_bf16(tacc, da, db, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&sh->empty[s]);
        }
        umma_commit(&sh->acc_full[slot]);
      }
    }
  } else {
    const int quarter = warp & 3;            // TMEM lane quarter this warp may read
    const int half = (warp - 2) >> 2;        // columns [BN_/2 half, BN_/2 half + BN_/2)
    const uint32_t epi = p.epi;
    uint32_t tl = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tl) {
      const int m0 = (tile / ntn) * BM, n0 = (tile % ntn) * BN_;
      const uint32_t slot = tl & 1u;
      mbar_wait(&sh->acc_full[slot], (tl >> 1) & 1u);
      tcgen05_fence_after();
      const int m = m0 + quarter * 32 + lane;
      const bool row_ok = m < p.M;
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + slot * (uint32_t)BN_;
      // (loading the next 32 columns from tensor memory while the current ones are converted was measured slower:
      // 644 vs 568 us on the [101376, 768] x [768, 3072] forward -- the epilogue is bound by its stores)
      if (STG) mbar_wait(&sh->stg_empty, (tl & 1u) ^ 1u);   // the previous tile's TMA store has read the staging images
#pragma unroll 1
      for (int c0 = half * (BN_ / 2); c0 < (half + 1) * (BN_ / 2); c0 += 32) {
        if (!STG && n0 + c0 >= p.N) break;  // warp-uniform
        uint32_t r[32];
        tmem_ld_x32(taddr + (uint32_t)c0, r);
        tmem_wait_ld();
        if (STG) epilogue_chunk_staged(p, epi, m, row_ok, quarter * 32 + lane, n0 + c0, c0, r, stg_out, stg_aux);
        else if (row_ok) epilogue_chunk(p, epi, m, n0 + c0, r);
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sh->acc_empty[slot]);
      if (STG) {
        fence_proxy_async_smem();          // the generic-proxy writes of the staging images, before the TMA reads them
        named_bar_sync(2, 256);            // all eight epilogue warps
        if (threadIdx.x == 64) {           // first epilogue thread: one box per 64-column slab (TMA clips at M and N)
#pragma unroll
          for (int sl = 0; sl < BN_ / 64; ++sl) {
            if (n0 + sl * 64 < p.N) {
              tma_store_2d(&tmap_out, smem_u32(stg_out + sl * 16384), n0 + sl * 64, m0);
              if (Cfg::kStgAux > 0 && (epi & PMGT_EPI_GELU)) tma_store_2d(&tmap_aux, smem_u32(stg_aux + sl * 16384), n0 + sl * 64, m0);
            }
          }
          tma_store_commit();
          tma_store_wait_read0();
          mbar_arrive(&sh->stg_empty);
        }
      }
    }
    if (STG && threadIdx.x == 64) tma_store_wait_all0();   // the last tile's stores have reached global memory
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(2 * BN_)));
  }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2D bf16 tensor map over a row-major [rows][inner] matrix with row pitch ld (elements)
int make_tmap(CUtensorMap* map, const void* base, long long inner, long long rows, long long ld,
                     int box_inner, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled entry point not available"); return PMGT_ERR_CUDA; }
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): base=%p inner=%lld rows=%lld ld=%lld", (int)r, base, inner, rows, ld);
    return PMGT_ERR_CUDA;
  }
  return PMGT_OK;
}

template <bool A_MN, bool B_MN, bool GA, bool GB, int STAGES>
static int launch(const CUtensorMap& ta, const CUtensorMap& tb, const GemmKernelArgs& ka, dim3 grid, cudaStream_t st) {
  auto kern = umma_gemm_kernel<A_MN, B_MN, GA, GB, STAGES>;
  const int smem = 2 * STAGES * kOperandStageBytes + (int)sizeof(GemmSmem) + 1024;
  static unsigned long long configured = 0;  // one flag per instantiation
  if (first_use_on_device(configured)) {
    PMGT_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  }
  kern<<<grid, kGemmThreads, smem, st>>>(ta, tb, ka);
  PMGT_LAUNCH_CHECK();
  return PMGT_OK;
}

// PMGT_GEMM_NARROW=1 keeps the persistent kernel on 128-column tiles (measurement switch)
static bool persist_narrow() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("PMGT_GEMM_NARROW");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v != 0;
}

template <bool A_MN, bool B_MN, int BN_, bool STG>
static int launch_persist_cfg(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to, const CUtensorMap& tx,
                              const GemmKernelArgs& ka, cudaStream_t st) {
  auto kern = umma_gemm_persist_kernel<A_MN, B_MN, BN_, STG>;
  constexpr int smem = PersistCfg<BN_, STG>::kSmem;
  static unsigned long long configured = 0;
  if (first_use_on_device(configured)) {
    PMGT_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  }
  const long long tiles = (long long)((ka.N + BN_ - 1) / BN_) * ((ka.M + BM - 1) / BM);
  const int grid = (int)(tiles < num_sms() ? tiles : num_sms());
  kern<<<grid, kPersistThreads, smem, st>>>(ta, tb, to, tx, ka);
  PMGT_LAUNCH_CHECK();
  return PMGT_OK;
}

// bf16 outputs leave through the staged TMA-store epilogue (PMGT_GEMM_DIRECT_STORE=1: row-per-thread stores)
template <bool A_MN, bool B_MN, int BN_>
static int launch_persist(const CUtensorMap& ta, const CUtensorMap& tb, const GemmKernelArgs& ka, cudaStream_t st) {
  static const bool direct = [] { const char* e = getenv("PMGT_GEMM_DIRECT_STORE"); return e && e[0] == '1'; }();
  // staged stores pay when the epilogue is long next to the main loop (K <= 1024: 428 vs 499 us on the qkvc Linear); with
  // a long K loop the shallower operand ring they leave costs more than the stores save (K = 3072: 383 vs 354 us)
  const bool bf16_out = !(ka.epi & (PMGT_EPI_OUT_F32 | PMGT_EPI_ATOMIC)) && ka.K <= 1024;
  CUtensorMap to, tx;
  memset(&to, 0, sizeof(to));
  memset(&tx, 0, sizeof(tx));
  if (bf16_out && !direct) {
    int rc = make_tmap(&to, ka.out, ka.N, ka.M, ka.ldo, 64, 128);
    if (rc) return rc;
    if (ka.epi & PMGT_EPI_GELU) {
      rc = make_tmap(&tx, ka.aux, ka.N, ka.M, ka.ld_aux, 64, 128);
      if (rc) return rc;
    }
    return launch_persist_cfg<A_MN, B_MN, BN_, true>(ta, tb, to, tx, ka, st);
  }
  return launch_persist_cfg<A_MN, B_MN, BN_, false>(ta, tb, to, tx, ka, st);
}

}  // namespace pmgt

using namespace pmgt;

extern "C" int pmgt_gemm_bf16(const pmgt_gemm_args* a, void* stream) {
  PMGT_REQUIRE(a, "pmgt_gemm_bf16: null args");
  PMGT_REQUIRE(a->M >= 0 && a->N >= 0 && a->K >= 0 && a->M < (1ll << 31) && a->N < (1ll << 31) && a->K < (1ll << 31),
               "pmgt_gemm_bf16: bad shape M=%lld N=%lld K=%lld", (long long)a->M, (long long)a->N, (long long)a->K);
  if (a->M == 0 || a->N == 0) return PMGT_OK;
  PMGT_REQUIRE(a->K > 0, "pmgt_gemm_bf16: K must be > 0");
  PMGT_REQUIRE(a->a && a->b && a->out, "pmgt_gemm_bf16: null operand");
  PMGT_REQUIRE(a->N % 8 == 0, "pmgt_gemm_bf16: N (%lld) must be a multiple of 8", (long long)a->N);
  PMGT_REQUIRE(a->lda % 8 == 0 && a->ldb % 8 == 0, "pmgt_gemm_bf16: lda/ldb must be multiples of 8 elements");
  PMGT_REQUIRE(((uintptr_t)a->a & 15) == 0 && ((uintptr_t)a->b & 15) == 0 && ((uintptr_t)a->out & 15) == 0,
               "pmgt_gemm_bf16: operands must be 16-byte aligned");
  const bool ga = a->a_rows != nullptr, gb = a->b_rows != nullptr;
  PMGT_REQUIRE(!(ga && a->a_mn), "pmgt_gemm_bf16: gathered A must be K-major (a_mn=0)");
  PMGT_REQUIRE(!(gb && !a->b_mn), "pmgt_gemm_bf16: gathered B must be MN-major (b_mn=1)");
  PMGT_REQUIRE(!(ga && gb), "pmgt_gemm_bf16: at most one gathered operand");
  PMGT_REQUIRE(!ga || a->K % 8 == 0, "pmgt_gemm_bf16: gathered A needs K %% 8 == 0");
  uint32_t epi = a->epi;
  if (epi & PMGT_EPI_ATOMIC) epi |= PMGT_EPI_OUT_F32;
  PMGT_REQUIRE(!(epi & PMGT_EPI_BIAS) || a->bias, "pmgt_gemm_bf16: EPI_BIAS without bias");
  PMGT_REQUIRE(!(epi & (PMGT_EPI_GELU | PMGT_EPI_GELU_BWD)) || (a->aux && a->ld_aux % 8 == 0),
               "pmgt_gemm_bf16: GELU epilogue needs aux with ld_aux %% 8 == 0");
  PMGT_REQUIRE(!(epi & PMGT_EPI_ADDEND) || (a->addend && a->ld_addend % 8 == 0),
               "pmgt_gemm_bf16: EPI_ADDEND needs addend with ld %% 8 == 0");
  PMGT_REQUIRE(a->ldo % ((epi & PMGT_EPI_OUT_F32) ? 4 : 8) == 0, "pmgt_gemm_bf16: ldo alignment");
  int split = a->split_k < 1 ? 1 : a->split_k;
  PMGT_REQUIRE(split == 1 || (epi & PMGT_EPI_ATOMIC), "pmgt_gemm_bf16: split_k > 1 requires PMGT_EPI_ATOMIC");
  PMGT_REQUIRE(split == 1 || !(epi & (PMGT_EPI_BIAS | PMGT_EPI_GELU | PMGT_EPI_GELU_BWD | PMGT_EPI_ADDEND)),
               "pmgt_gemm_bf16: split_k > 1 supports only the plain atomic epilogue");

  const int num_kb = (int)((a->K + BK - 1) / BK);
  if (split > num_kb) split = num_kb;
  const int kb_per = (num_kb + split - 1) / split;
  split = (num_kb + kb_per - 1) / kb_per;

  CUtensorMap ta, tb;
  memset(&ta, 0, sizeof(ta));
  memset(&tb, 0, sizeof(tb));
  int rc;
  if (!ga) {
    rc = a->a_mn ? make_tmap(&ta, a->a, a->M, a->K, a->lda, 64, 64) : make_tmap(&ta, a->a, a->K, a->M, a->lda, 64, 128);
    if (rc) return rc;
  }
  if (!gb) {
    rc = a->b_mn ? make_tmap(&tb, a->b, a->N, a->K, a->ldb, 64, 64) : make_tmap(&tb, a->b, a->K, a->N, a->ldb, 64, 128);
    if (rc) return rc;
  }
  GemmKernelArgs ka;
  ka.M = (int)a->M; ka.N = (int)a->N; ka.K = (int)a->K;
  ka.a_src = a->a; ka.lda = a->lda; ka.a_rows = (const long long*)a->a_rows; ka.a_src_rows = a->a_src_rows;
  ka.b_src = a->b; ka.ldb = a->ldb; ka.b_rows = (const long long*)a->b_rows; ka.b_src_rows = a->b_src_rows;
  ka.out = a->out; ka.ldo = a->ldo; ka.bias = a->bias;
  ka.addend = a->addend; ka.ld_addend = a->ld_addend; ka.aux = a->aux; ka.ld_aux = a->ld_aux;
  ka.alpha = a->alpha; ka.epi = epi; ka.kb_per_split = kb_per;
  ka.vec32 = (((uintptr_t)a->out & 31) == 0 && a->ldo % 16 == 0 &&
              (!(epi & PMGT_EPI_GELU) || (((uintptr_t)a->aux & 31) == 0 && a->ld_aux % 16 == 0))) ? 1 : 0;
  dim3 grid((unsigned)((a->N + BN - 1) / BN), (unsigned)((a->M + BM - 1) / BM), (unsigned)split);
  PMGT_REQUIRE(grid.y <= 65535u && grid.z <= 65535u, "pmgt_gemm_bf16: grid too large (M tiles %u, split %u)", grid.y, grid.z);
  cudaStream_t st = (cudaStream_t)stream;
  // dense operands, no split-K, more tiles than one wave of the two-CTA-per-SM kernel: the persistent kernel; with 256
  // output columns per tile when N allows (85 instead of 64 FLOP per operand byte through L2 -> SM)
  if (!ga && !gb && split == 1 && !(epi & PMGT_EPI_ATOMIC) && (long long)grid.x * grid.y > 2ll * num_sms() && num_kb >= 4) {
    // (the GELU epilogues are bound by their own instruction issue, and measured slower on the wide tiles: 1353 vs 919 us)
    const bool wide_n = a->N >= 512 && !(epi & (PMGT_EPI_GELU | PMGT_EPI_GELU_BWD)) && !persist_narrow();
    if (!a->a_mn && !a->b_mn) return wide_n ? launch_persist<false, false, 256>(ta, tb, ka, st) : launch_persist<false, false, 128>(ta, tb, ka, st);
    if (!a->a_mn && a->b_mn) return wide_n ? launch_persist<false, true, 256>(ta, tb, ka, st) : launch_persist<false, true, 128>(ta, tb, ka, st);
    if (a->a_mn && a->b_mn) return wide_n ? launch_persist<true, true, 256>(ta, tb, ka, st) : launch_persist<true, true, 128>(ta, tb, ka, st);
  }
  // 3 stages = 96 KiB of operand ring: TWO CTAs fit per SM, so one tile's prologue / epilogue overlaps the other's
  // main loop (with 4 stages a single resident CTA left the SM idle during every tile's head and tail)
  const bool deep = kb_per > 2;

#define PMGT_DISPATCH(AM, BMN, GA_, GB_)                                              \
  return deep ? launch<AM, BMN, GA_, GB_, 3>(ta, tb, ka, grid, st) : launch<AM, BMN, GA_, GB_, 2>(ta, tb, ka, grid, st)
  if (!a->a_mn && !a->b_mn && !ga) { PMGT_DISPATCH(false, false, false, false); }
  if (!a->a_mn && !a->b_mn && ga) { PMGT_DISPATCH(false, false, true, false); }
  if (!a->a_mn && a->b_mn && !ga && !gb) { PMGT_DISPATCH(false, true, false, false); }
  if (a->a_mn && a->b_mn && !gb) { PMGT_DISPATCH(true, true, false, false); }
  if (a->a_mn && a->b_mn && gb) { PMGT_DISPATCH(true, true, false, true); }
#undef PMGT_DISPATCH
  set_error("pmgt_gemm_bf16: unsupported operand layout a_mn=%d b_mn=%d gather_a=%d gather_b=%d", a->a_mn, a->b_mn,
            (int)ga, (int)gb);
  return PMGT_ERR_UNSUPPORTED;
}
