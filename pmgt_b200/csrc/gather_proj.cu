// Gather-fused feature projections of PMGTEmbeddings on large graphs (H = 128): the rows of a frozen bf16 feature
// table are fetched through the token -> node-id vector INSIDE the GEMM (replaces get_input_feat_embeds,
// pmgt/pmgt/utils.py:43-50, + feat_linear, pmgt/pmgt/modeling_pmgt.py:195-198, and the weight gradient of that Linear).
//
//   forward   out[T][128]  = table[rows[t]][0:K] . W[128][K]^T + bias                       (gather_proj_fwd_kernel)
//   backward  dW[128][K]  += dY[T][128]^T . table[rows[t]][0:K]                              (gather_proj_dw_kernel)
//
// Both are bound by the random row gather from HBM (3072-byte / 1536-byte rows out of a multi-GB table).  What the
// memory system rewards (tools/probes/gather_probe.cu, B200): requests that cover >= 256 contiguous bytes of a row at
// a time (128-byte pieces of a row spread over time cap at ~4.6 TB/s whatever the concurrency; 256-byte pieces reach
// 6.5+) and ~150-250 KB of requests in flight per SM.  So, unlike the general GEMM (gemm_umma.cu: 64-column stages,
// two CTAs per SM, 96 KB ring shared by both operands):
//   * one persistent CTA per SM whose shared memory is almost entirely the ring of the GATHERED operand
//     (forward 5 x 32 KB: 128 rows x 256 B per stage; backward 5 x 32 KB: 32 rows x 1024 B per stage);
//   * the rows are fetched by TMA (cp.async.bulk.tensor.2d tile::gather4, box {64 columns, 1 row}, 128-byte swizzle):
//     one instruction lands 4 arbitrary table rows x 128 B as 4 consecutive rows of a UMMA slab and reports its bytes
//     to the stage's mbarrier, so NO thread hands a stage over -- the MMA issuer wakes when the last byte lands.  An
//     instruction occupies its issuing thread for ~120 clocks and a warp serialises its lanes, so 8 warps issue with one
//     lane each (tools/probes/gather4_ring.cu: one warp 2.1 TB/s, 8 warps 5.8, 16 warps 6.4);
//   * the fallback / comparison path (PMGT_GATHER_TMA=0, bit 0 forward, bit 1 backward): 128 threads x 16-byte cp.async
//     with hand-written swizzle, a warp covering 512 contiguous bytes of ONE row, and a gather thread handing over the
//     stage that has landed BEFORE it waits for the next free slot (the other order chains consume -> issue -> arrive
//     into one serial loop of ~1 us per stage: 4.4 instead of 5.4 TB/s);
//   * the dense operand (W from L2, dY) streams through a small TMA ring of its own;
//   * forward: two TMEM accumulators, the epilogue of tile n runs under the gather of tile n + 1;
//     backward: every CTA owns a [128 x 512] (or 384) slice of dW in TMEM for its token range and flushes it once.
// Measured on B200 (tools/bench_gather.py, 294,912 tokens, 1M-row tables, 75 % unique rows; gathered + written bytes):
//              TMA gather4                        cp.async
//   forward    187 us visual / 96 us text         185 / 105 us   (5.3 / 5.5 TB/s  vs  5.3 / 5.1)
//   dW         164 us visual / 90 us text         224 / 120 us   (6.0 / 5.9 TB/s  vs  4.4 / 4.4)
// (pmgt_gemm_bf16 with a_rows / b_rows: 237 / 160 and 245 / 144 us.)  Tried and dropped: cp.async.bulk.prefetch.L2 of the
// rows of later stages (forward 270 us, dW 378 us: the prefetches compete with the demand fetches).  The forward
// kernel does not depend on its ring depths (gathered operand / W stages 5 + 3, 5 + 4, 4 + 6, 4 + 5: 184-186 us visual,
// 94.6-95.0 us text), so neither the W stream from L2 nor the bytes in flight bound it.
#include "umma.cuh"

namespace pmgt {

constexpr int kGpRing = 5;  // stages of the gathered operand in flight
constexpr int kGpFwdWarpsDefault = 8;   // 8 warps: 184 / 95 us, 16 warps: 199 / 98 us
constexpr int kGpDwWarpsDefault = 16;  // 8 warps: 161 / 89 us, 16 warps: 148 / 85 us (visual / text)
constexpr int kGpTmaDefault = 3;  // PMGT_GATHER_TMA when the variable is not set: TMA row gather in both kernels

__device__ __forceinline__ unsigned char* gp_align1024(unsigned char* p) {
  return reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(p) + 1023) & ~(uintptr_t)1023);
}
__device__ __forceinline__ void gp_cp_async_8(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
// TMA row gather: 4 arbitrary rows x one 64-column (128-byte) box of a row-major bf16 matrix land as 4 consecutive
// 128-byte rows of a 128B-swizzled slab (dst = slab + 512 * row group; tools/probes/gather4_probe.cu).  Row indices
// outside the tensor map read as zeros.  The instruction occupies its issuing thread for ~120 clocks and a warp
// serialises its lanes, so the issuing lanes are spread over 8 warps (tools/probes/gather4_ring.cu).
__device__ __forceinline__ void gp_tma_gather4(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int col, int r0, int r1,
                                               int r2, int r3) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(col),
      "r"(r0), "r"(r1), "r"(r2), "r"(r3)
      : "memory");
}
__device__ __forceinline__ void gp_stg256(void* p, const uint32_t* r) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(r[0]), "r"(r[1]), "r"(r[2]),
               "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}

// ---------------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------------
constexpr int kGpFwdThreads = 320;  // warp 0: W TMA producer | 1: MMA issuer | 2-5: gather | 6-9: epilogue
constexpr int gp_fwd_threads(bool tmag, int gw) { return tmag ? 64 + 128 + gw * 32 : kGpFwdThreads; }  // TMA gather: warps 2-5 epilogue | 6..: one issuing lane each
constexpr int kGpFwdAStage = 32768; // 128 rows x 128 columns: two 64-column slabs
constexpr int kGpFwdWStage = 16384; // W[128][64 columns]

struct GpFwdParams {
  int T, K, num_tiles;
  const uint16_t* table;
  long long ld;
  const long long* rows;
  long long src_rows;
  uint16_t* out;
  long long ldo;
  const float* bias;
};

template <bool TMAG, int ARING, int WRING>
struct GpFwdShared {
  uint64_t a_full[ARING], a_empty[ARING];
  uint64_t w_full[WRING], w_empty[WRING];
  uint64_t acc_full[2], acc_empty[2];
  uint32_t tmem_base;
  int ids[TMAG ? 1 : 4][TMAG ? 1 : 128];  // node ids of the tiles in flight (the TMA path keeps them in registers)
};

template <bool TMAG, int ARING, int WRING>
constexpr int gp_fwd_smem() {
  return ARING * kGpFwdAStage + WRING * kGpFwdWStage + (int)sizeof(GpFwdShared<TMAG, ARING, WRING>) + 1024;
}

__device__ __forceinline__ int gp_row_id(const long long* rows, long long src_rows, int m, int T) {
  if (m >= T) return -1;
  const long long r = rows[m];
  return (r < 0 || r >= src_rows) ? -1 : (int)r;
}

// Gathered operand of the forward kernel by TMA: gather warp gw of GW owns RPW = 128 / GW rows of every tile; per stage
// its lane 0 issues RPW / 4 row groups x 2 slabs (512 B each) against the stage's barrier, which needs no thread to hand it over.
template <int ARING, int GW, class Shared>
__device__ __forceinline__ void gp_fwd_gather_tma(const CUtensorMap& tmap_tab, const GpFwdParams& p, Shared* sh, unsigned char* sA,
                                                  int gw, int lane, int nstage) {
  constexpr int RPW = 128 / GW;
  auto load_id = [&](int tile) -> int {
    if (lane >= RPW || tile >= p.num_tiles) return (int)p.src_rows;
    const int r = gp_row_id(p.rows, p.src_rows, tile * 128 + gw * RPW + lane, p.T);
    return r < 0 ? (int)p.src_rows : r;  // beyond the tensor map: zero row
  };
  int myid = load_id(blockIdx.x);
  uint32_t g = 0;
  for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
    const int nid = load_id(tile + gridDim.x);
    int rid[RPW];
#pragma unroll
    for (int i = 0; i < RPW; ++i) rid[i] = __shfl_sync(0xffffffffu, myid, i);
    if (lane == 0) {
      for (int st = 0; st < nstage; ++st, ++g) {
        const uint32_t s = g % ARING;
        mbar_wait(&sh->a_empty[s], ((g / ARING) & 1u) ^ 1u);
        mbar_arrive_expect_tx(&sh->a_full[s], (uint32_t)(kGpFwdAStage / GW));
        const uint32_t base = smem_u32(sA + s * kGpFwdAStage) + (uint32_t)(gw * RPW * 128);
#pragma unroll
        for (int q = 0; q < RPW / 4; ++q) {
#pragma unroll
          for (int h = 0; h < 2; ++h)
            gp_tma_gather4(base + h * 16384 + q * 512, &tmap_tab, &sh->a_full[s], st * 128 + h * 64, rid[4 * q],
                           rid[4 * q + 1], rid[4 * q + 2], rid[4 * q + 3]);
        }
      }
    }
    __syncwarp();
    myid = nid;
  }
}

// Gathered operand of the forward kernel by cp.async (PMGT_GATHER_TMA bit 0 clear): 128 threads, 16 lanes x 16 B = 256
// contiguous bytes of one table row.
template <int ARING, class Shared>
__device__ __forceinline__ void gp_fwd_gather_cp_async(const GpFwdParams& p, Shared* sh, unsigned char* sA, int t, int nstage) {
  const int c16 = t & 15, rb = t >> 4;
  const uint32_t dst_off = (uint32_t)((c16 >> 3) * 16384 + rb * 128 + (((c16 & 7) ^ rb) << 4));
  constexpr int LAG = ARING - 1;
  uint32_t g = 0;
  sh->ids[0][t] = gp_row_id(p.rows, p.src_rows, blockIdx.x * 128 + t, p.T);
  named_bar_sync(1, 128);
  uint32_t tl = 0;
  for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++tl) {
    const int buf = tl & 3, nbuf = (tl + 1) & 3;
    const int next = tile + gridDim.x;
    // the next tile's node ids are requested now and parked in the next buffer after the first stage is on its way
    const int nid = next < p.num_tiles ? gp_row_id(p.rows, p.src_rows, next * 128 + t, p.T) : -1;
    int rid[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) rid[i] = sh->ids[buf][rb + 8 * i];
    for (int st = 0; st < nstage; ++st, ++g) {
      const uint32_t s = g % ARING;
      // hand over the stage issued LAG iterations ago BEFORE waiting for a free slot: the MMA on it then overlaps the
      // address arithmetic and issue of this stage (arriving after the issue chained consume -> issue -> arrive ->
      // consume into one serial loop of ~1 us per stage, whatever the memory latency)
      if (g >= (uint32_t)LAG) {
        cp_async_wait<LAG - 1>();
        fence_proxy_async_smem();
        mbar_arrive(&sh->a_full[(g - LAG) % ARING]);
      }
      mbar_wait(&sh->a_empty[s], ((g / ARING) & 1u) ^ 1u);
      const uint32_t base = smem_u32(sA + s * kGpFwdAStage) + dst_off;
      const int col = st * 128 + c16 * 8;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const bool ok = rid[i] >= 0;
        const uint16_t* src = ok ? p.table + (long long)rid[i] * p.ld + col : p.table;
        cp_async_16(base + i * 1024, src, ok ? 16u : 0u);
      }
      cp_async_commit();
      if (st == 0) {
        sh->ids[nbuf][t] = nid;
        named_bar_sync(1, 128);
      }
    }
  }
  cp_async_wait<0>();
  fence_proxy_async_smem();
  for (uint32_t i = (g > (uint32_t)LAG ? g - LAG : 0u); i < g; ++i) mbar_arrive(&sh->a_full[i % ARING]);  // the last LAG stages
}

// TMAG: gathered operand by TMA tile::gather4 (true) or by 16-byte cp.async (false); ARING / WRING: stages of the gathered
// operand (32 KB each) and of W (16 KB each) in flight
template <bool TMAG, int ARING, int WRING, int GW>
__global__ void __launch_bounds__(gp_fwd_threads(TMAG, GW), 1)
gather_proj_fwd_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_tab,
                       const GpFwdParams p) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = gp_align1024(smem_dyn);
  unsigned char* sA = smem;
  unsigned char* sW = smem + ARING * kGpFwdAStage;
  using Shared = GpFwdShared<TMAG, ARING, WRING>;
  Shared* sh = reinterpret_cast<Shared*>(sW + WRING * kGpFwdWStage);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nstage = p.K >> 7;  // 128-column stages per tile

  if (threadIdx.x == 0) {
    for (int s = 0; s < ARING; ++s) { mbar_init(&sh->a_full[s], TMAG ? (uint32_t)GW : 128u); mbar_init(&sh->a_empty[s], 1u); }
    for (int s = 0; s < WRING; ++s) { mbar_init(&sh->w_full[s], 1u); mbar_init(&sh->w_empty[s], 1u); }
    for (int s = 0; s < 2; ++s) { mbar_init(&sh->acc_full[s], 1u); mbar_init(&sh->acc_empty[s], 4u); }
    fence_barrier_init();
    prefetch_tmap(&tmap_w);
    if (TMAG) prefetch_tmap(&tmap_tab);
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sh->tmem_base)), "r"(256u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = sh->tmem_base;

  if (warp == 0) {
    // ===================== W producer: one [128 x 64] k-block per TMA, straight from L2 =====================
    if (lane == 0) {
      uint32_t wi = 0;
      const int nkb = p.K >> 6;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        for (int kb = 0; kb < nkb; ++kb, ++wi) {
          const uint32_t s = wi % WRING;
          mbar_wait_idle(&sh->w_empty[s], ((wi / WRING) & 1u) ^ 1u);
          mbar_arrive_expect_tx(&sh->w_full[s], (uint32_t)kGpFwdWStage);
          tma_load_2d(smem_u32(sW + s * kGpFwdWStage), &tmap_w, &sh->w_full[s], kb * 64, 0);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      uint32_t ai = 0, wi = 0, tl = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++tl) {
        const uint32_t slot = tl & 1u;
        mbar_wait(&sh->acc_empty[slot], ((tl >> 1) & 1u) ^ 1u);
        tcgen05_fence_after();
        const uint32_t tacc = tmem_base + slot * 128u;
        for (int st = 0; st < nstage; ++st, ++ai) {
          const uint32_t sa = ai % ARING;
          mbar_wait(&sh->a_full[sa], (ai / ARING) & 1u);
          const uint32_t a_base = smem_u32(sA + sa * kGpFwdAStage);
#pragma unroll
          for (int h = 0; h < 2; ++h, ++wi) {
            const uint32_t sw = wi % WRING;
            mbar_wait(&sh->w_full[sw], (wi / WRING) & 1u);
            tcgen05_fence_after();
            const uint32_t w_base = smem_u32(sW + sw * kGpFwdWStage);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              umma_bf16(tacc, umma_desc(a_base + h * 16384 + k * 32, 16, 1024), umma_desc(w_base + k * 32, 16, 1024), idesc,
                        (st > 0 || h > 0 || k > 0) ? 1u : 0u);
            }
            umma_commit(&sh->w_empty[sw]);
          }
          umma_commit(&sh->a_empty[sa]);
        }
        umma_commit(&sh->acc_full[slot]);
      }
    }
  } else if (TMAG ? (warp >= 6) : (warp < 6)) {
    // ===================== gather =====================
    if constexpr (TMAG) gp_fwd_gather_tma<ARING, GW>(tmap_tab, p, sh, sA, warp - 6, lane, nstage);
    else gp_fwd_gather_cp_async<ARING>(p, sh, sA, (int)threadIdx.x - 64, nstage);
  } else {
    // ===================== epilogue: + bias, bf16, one output row per thread =====================
    const int quarter = warp & 3;
    uint32_t tl = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++tl) {
      const uint32_t slot = tl & 1u;
      mbar_wait(&sh->acc_full[slot], (tl >> 1) & 1u);
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + slot * 128u;
      const int m = tile * 128 + quarter * 32 + lane;
      uint16_t* orow = p.out + (long long)m * p.ldo;
#pragma unroll 1
      for (int c0 = 0; c0 < 128; c0 += 32) {
        uint32_t r[32];
        tmem_ld_x32(taddr + (uint32_t)c0, r);
        tmem_wait_ld();
        if (m < p.T) {
#pragma unroll
          for (int g8 = 0; g8 < 2; ++g8) {
            uint32_t o[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int c = c0 + g8 * 16 + 2 * j;
              const float2 b = p.bias != nullptr ? __ldg(reinterpret_cast<const float2*>(p.bias + c)) : make_float2(0.f, 0.f);
              o[j] = pack_bf16x2(__uint_as_float(r[g8 * 16 + 2 * j]) + b.x, __uint_as_float(r[g8 * 16 + 2 * j + 1]) + b.y);
            }
            gp_stg256(orow + c0 + g8 * 16, o);
          }
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sh->acc_empty[slot]);
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u));
  }
}

// ---------------------------------------------------------------------------------------------------
// backward (weight gradient)
// ---------------------------------------------------------------------------------------------------
constexpr int kGpDwThreads = 192;  // warp 0: dY TMA producer | 1: MMA issuer | 2-5: gather, then the flush
constexpr int gp_dw_threads(bool tmag, int gw) { return tmag ? 64 + gw * 32 : kGpDwThreads; }  // TMA gather: warps 2.. one issuing lane each, warps 2-5 flush
constexpr int kGpDwTok = 32;       // tokens per stage
constexpr int kGpDwAStage = 8192;  // dY[32 tokens][128]: two slabs of 32 rows x 128 B

struct GpDwParams {
  int T, K, tok_per_cta;
  const uint16_t* table;
  long long ld;
  const long long* rows;
  long long src_rows;
  float* dw;
  long long ld_dw;
};

constexpr int kGpDwIdRing = 16;   // stages of node ids held in shared memory
constexpr int kGpDwIdAhead = 12;  // ... requested this many stages ahead (8-byte cp.async in the gather's own groups)

struct GpDwShared {
  uint64_t full[kGpRing], empty[kGpRing];
  uint64_t acc_full;
  uint32_t tmem_base;
  long long ids[kGpDwIdRing][kGpDwTok];
};

template <int NC>
constexpr int gp_dw_smem() { return kGpRing * (kGpDwAStage + NC * 4096) + (int)sizeof(GpDwShared) + 1024; }

// Gathered operand of the weight-gradient kernel by TMA: issuing warp gw = (warp - 2) % 8 owns tokens 4 gw .. 4 gw + 3 of every
// 32-token stage (GW = 16: and one half of the slabs); its lane 0 issues one gather4 per 64-column slab.  Node ids: one
// load per 8 stages (lane -> stage lane / 4, token lane % 4), the next block requested before the current one is issued.
template <int NC, int GW>
__device__ __forceinline__ void gp_dw_gather_tma(const CUtensorMap& tmap_tab, const GpDwParams& p, GpDwShared* sh, unsigned char* sB,
                                                 int warp, int lane, int n0, int tok0, int tok1, int nstage) {
  constexpr int kSlabs = NC / (GW / 8);  // slabs per issuing warp
  const int gw = (warp - 2) & 7, c_first = ((warp - 2) >> 3) * kSlabs;
  auto load_ids = [&](int blk) -> int {
    const int st = blk * 8 + (lane >> 2);
    const long long tok = (long long)tok0 + (long long)st * kGpDwTok + gw * 4 + (lane & 3);
    if (st >= nstage || tok >= tok1) return (int)p.src_rows;
    const long long r = p.rows[tok];
    return (r < 0 || r >= p.src_rows) ? (int)p.src_rows : (int)r;  // beyond the tensor map: zero row
  };
  int cur = load_ids(0);
  for (int blk = 0; blk * 8 < nstage; ++blk) {
    const int nxt = load_ids(blk + 1);
    const int jn = nstage - blk * 8 < 8 ? nstage - blk * 8 : 8;
    for (int j = 0; j < jn; ++j) {
      const int i = blk * 8 + j;
      const int r0 = __shfl_sync(0xffffffffu, cur, 4 * j), r1 = __shfl_sync(0xffffffffu, cur, 4 * j + 1);
      const int r2 = __shfl_sync(0xffffffffu, cur, 4 * j + 2), r3 = __shfl_sync(0xffffffffu, cur, 4 * j + 3);
      if (lane == 0) {
        const uint32_t s = (uint32_t)i % kGpRing;
        mbar_wait(&sh->empty[s], (((uint32_t)i / kGpRing) & 1u) ^ 1u);
        mbar_arrive_expect_tx(&sh->full[s], (uint32_t)(kSlabs * 512));
        const uint32_t base = smem_u32(sB + s * (NC * 4096)) + (uint32_t)gw * 512u;
#pragma unroll
        for (int c = c_first; c < c_first + kSlabs; ++c)
          gp_tma_gather4(base + c * 4096, &tmap_tab, &sh->full[s], n0 + c * 64, r0, r1, r2, r3);
      }
      __syncwarp();
    }
    cur = nxt;
  }
}

// Gathered operand of the weight-gradient kernel by cp.async (PMGT_GATHER_TMA bit 1 clear): 128 threads, a warp's 32 lanes
// cover 512 contiguous bytes of one table row.
template <int NC>
__device__ __forceinline__ void gp_dw_gather_cp_async(const GpDwParams& p, GpDwShared* sh, unsigned char* sB, int t, int n0, int tok0,
                                                      int tok1, int nstage) {
  constexpr int kPerThread = 2 * NC;  // 16-byte copies per gather thread and stage
  constexpr int LAG = kGpRing - 1;
  constexpr int kChunks = NC * 8;  // 16-byte chunks per row and stage
  int row_of[kPerThread];
  uint32_t dst_of[kPerThread];
  int col_of[kPerThread];
#pragma unroll
  for (int j = 0; j < kPerThread; ++j) {
    const int flat = j * 128 + t;
    const int r = flat / kChunks, ch = flat % kChunks;
    row_of[j] = r;
    col_of[j] = n0 + ch * 8;
    dst_of[j] = (uint32_t)((ch >> 3) * 4096 + r * 128 + (((ch & 7) ^ (r & 7)) << 4));
  }
  // node ids: thread t < 32 owns token t of every stage.  Stages [0, kGpDwIdAhead) are loaded here; inside the loop
  // the ids of stage i + kGpDwIdAhead ride in stage i's cp.async group (no register ever waits on an id load), and
  // become visible to the other threads through the full -> MMA -> empty barrier chain long before they are read.
  auto id_src = [&](int stage) { return p.rows + tok0 + stage * kGpDwTok + t; };
  auto id_ok = [&](int stage) { return tok0 + stage * kGpDwTok + t < tok1; };
  if (t < kGpDwTok) {
    for (int st = 0; st < kGpDwIdAhead && st < nstage; ++st) sh->ids[st % kGpDwIdRing][t] = id_ok(st) ? *id_src(st) : -1ll;
  }
  named_bar_sync(1, 128);
  for (int i = 0; i < nstage; ++i) {
    const uint32_t s = (uint32_t)i % kGpRing;
    if (i >= LAG) {  // see the forward kernel: arrive first, then wait for the free slot
      cp_async_wait<LAG - 1>();
      fence_proxy_async_smem();
      mbar_arrive(&sh->full[(i - LAG) % kGpRing]);
    }
    mbar_wait(&sh->empty[s], (((uint32_t)i / kGpRing) & 1u) ^ 1u);
    const uint32_t base = smem_u32(sB + s * (NC * 4096));
    const long long* ids = sh->ids[i % kGpDwIdRing];
#pragma unroll
    for (int j = 0; j < kPerThread; ++j) {
      const long long r = ids[row_of[j]];
      const bool ok = r >= 0 && r < p.src_rows;
      const uint16_t* src = ok ? p.table + r * p.ld + col_of[j] : p.table;
      cp_async_16(base + dst_of[j], src, ok ? 16u : 0u);
    }
    if (t < kGpDwTok && i + kGpDwIdAhead < nstage) {
      long long* dst = &sh->ids[(i + kGpDwIdAhead) % kGpDwIdRing][t];
      if (id_ok(i + kGpDwIdAhead)) gp_cp_async_8(smem_u32(dst), id_src(i + kGpDwIdAhead));
      else *dst = -1ll;
    }
    cp_async_commit();
  }
  cp_async_wait<0>();
  fence_proxy_async_smem();
  for (int i = (nstage > LAG ? nstage - LAG : 0); i < nstage; ++i) mbar_arrive(&sh->full[i % kGpRing]);
}

template <int NC, bool TMAG, int GW>  // NC: 64-column slabs of the table row one CTA covers (8: 512 columns, 6: 384, 4: 256)
__global__ void __launch_bounds__(gp_dw_threads(TMAG, GW), 1)
gather_proj_dw_kernel(const __grid_constant__ CUtensorMap tmap_dy, const __grid_constant__ CUtensorMap tmap_tab,
                      const GpDwParams p) {
  constexpr int kBStage = NC * 4096;
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = gp_align1024(smem_dyn);
  unsigned char* sB = smem;
  unsigned char* sAm = smem + kGpRing * kBStage;
  GpDwShared* sh = reinterpret_cast<GpDwShared*>(sAm + kGpRing * kGpDwAStage);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * NC * 64;
  const int tok0 = blockIdx.y * p.tok_per_cta;
  int tok1 = tok0 + p.tok_per_cta;
  if (tok1 > p.T) tok1 = p.T;
  const int nstage = tok1 > tok0 ? (tok1 - tok0 + kGpDwTok - 1) / kGpDwTok : 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kGpRing; ++s) { mbar_init(&sh->full[s], 1u + (TMAG ? (uint32_t)GW : 128u)); mbar_init(&sh->empty[s], 1u); }
    mbar_init(&sh->acc_full, 1u);
    fence_barrier_init();
    prefetch_tmap(&tmap_dy);
    if (TMAG) prefetch_tmap(&tmap_tab);
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sh->tmem_base)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = sh->tmem_base;

  if (warp == 0) {
    if (lane == 0) {
      for (int i = 0; i < nstage; ++i) {
        const uint32_t s = (uint32_t)i % kGpRing;
        mbar_wait_idle(&sh->empty[s], (((uint32_t)i / kGpRing) & 1u) ^ 1u);
        mbar_arrive_expect_tx(&sh->full[s], (uint32_t)kGpDwAStage);
        const uint32_t dst = smem_u32(sAm + s * kGpDwAStage);
        tma_load_2d(dst, &tmap_dy, &sh->full[s], 0, tok0 + i * kGpDwTok);
        tma_load_2d(dst + 4096, &tmap_dy, &sh->full[s], 64, tok0 + i * kGpDwTok);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // M = 128 (output features, MN-major dY), N = NC * 32 per instruction (MN-major table rows), K = 16 tokens
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)((NC * 32) >> 3) << 17) |
                             ((uint32_t)(128 >> 4) << 24);
      for (int i = 0; i < nstage; ++i) {
        const uint32_t s = (uint32_t)i % kGpRing;
        mbar_wait(&sh->full[s], ((uint32_t)i / kGpRing) & 1u);
        tcgen05_fence_after();
        const uint32_t a_base = smem_u32(sAm + s * kGpDwAStage);
        const uint32_t b_base = smem_u32(sB + s * kBStage);
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const uint64_t da = umma_desc(a_base + k * 2048, 4096, 1024);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const uint64_t db = umma_desc(b_base + h * (NC / 2) * 4096 + k * 2048, 4096, 1024);
            umma_bf16(tmem_base + (uint32_t)(h * NC * 32), da, db, idesc, (i > 0 || k > 0) ? 1u : 0u);
          }
        }
        umma_commit(&sh->empty[s]);
      }
      umma_commit(&sh->acc_full);
    }
  } else {
    // ===================== gather (warps 2..), then the flush (warps 2-5) =====================
    if constexpr (TMAG) gp_dw_gather_tma<NC, GW>(tmap_tab, p, sh, sB, warp, lane, n0, tok0, tok1, nstage);
    else gp_dw_gather_cp_async<NC>(p, sh, sB, (int)threadIdx.x - 64, n0, tok0, tok1, nstage);

    // ---- flush: dW[m][n0 ..] += accumulator row m, 32 columns at a time, the start column rotated per CTA so that
    //      the CTAs of one column group do not hit the same L2 lines in lock-step
    if (nstage > 0 && warp < 6) {
      mbar_wait(&sh->acc_full, 0u);
      tcgen05_fence_after();
      const int quarter = warp & 3;
      const int m = quarter * 32 + lane;
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
      float* drow = p.dw + (long long)m * p.ld_dw + n0;
      constexpr int kGroups = NC * 2;  // 32-column groups
#pragma unroll 1
      for (int gq = 0; gq < kGroups; ++gq) {
        const int c0 = ((gq + (int)blockIdx.y) % kGroups) * 32;
        uint32_t r[32];
        tmem_ld_x32(taddr + (uint32_t)c0, r);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(drow + c0 + 4 * j), "f"(__uint_as_float(r[4 * j])),
                       "f"(__uint_as_float(r[4 * j + 1])), "f"(__uint_as_float(r[4 * j + 2])), "f"(__uint_as_float(r[4 * j + 3]))
                       : "memory");
        }
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

template <int NC, bool TMAG, int GW>
static int launch_dw(const CUtensorMap& tm, const CUtensorMap& ttab, const GpDwParams& kp, dim3 grid, cudaStream_t st) {
  auto kern = gather_proj_dw_kernel<NC, TMAG, GW>;
  constexpr int smem = gp_dw_smem<NC>();
  static unsigned long long configured = 0;
  if (first_use_on_device(configured)) PMGT_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  kern<<<grid, gp_dw_threads(TMAG, GW), smem, st>>>(tm, ttab, kp);
  PMGT_LAUNCH_CHECK();
  return PMGT_OK;
}

template <bool TMAG, int ARING, int WRING, int GW>
static int launch_fwd(const CUtensorMap& tw, const CUtensorMap& ttab, const GpFwdParams& kp, int grid, cudaStream_t st) {
  auto kern = gather_proj_fwd_kernel<TMAG, ARING, WRING, GW>;
  constexpr int smem = gp_fwd_smem<TMAG, ARING, WRING>();
  static_assert(smem <= 232448, "shared memory of the forward gather kernel");
  static unsigned long long configured = 0;
  if (first_use_on_device(configured)) PMGT_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  kern<<<grid, gp_fwd_threads(TMAG, GW), smem, st>>>(tw, ttab, kp);
  PMGT_LAUNCH_CHECK();
  return PMGT_OK;
}

// PMGT_GATHER_TMA (read at every call, so a test can compare both paths in one process): bit 0 = forward, bit 1 = weight
// gradient fetch the table rows by TMA tile::gather4 instead of 16-byte cp.async
static int gather_tma_mode() {
  const char* e = getenv("PMGT_GATHER_TMA");
  return e && *e ? (atoi(e) & 3) : kGpTmaDefault;
}

static int dw_slabs(long long K) { return K % 512 == 0 ? 8 : (K % 384 == 0 ? 6 : (K % 256 == 0 ? 4 : 0)); }

}  // namespace pmgt

using namespace pmgt;

extern "C" int pmgt_gather_proj_supported(int64_t N, int64_t K) {
  return N == 128 && K > 0 && K % 128 == 0 && dw_slabs(K) != 0 ? 1 : 0;
}

static int gp_check(const pmgt_gather_proj_args* a, const char* who) {
  PMGT_REQUIRE(a, "%s: null args", who);
  PMGT_REQUIRE(a->T >= 0 && a->T < (1ll << 31) - 4096, "%s: bad token count %lld", who, (long long)a->T);
  PMGT_REQUIRE(pmgt_gather_proj_supported(128, a->K), "%s: K = %lld is not supported (multiple of 128 and of 256/384/512)", who,
               (long long)a->K);
  PMGT_REQUIRE(a->table && a->rows, "%s: null table / rows", who);
  PMGT_REQUIRE(a->ld % 8 == 0 && ((uintptr_t)a->table & 15) == 0, "%s: table rows must be 16-byte aligned", who);
  PMGT_REQUIRE(a->table_rows > 0 && a->table_rows < (1ll << 31), "%s: bad table_rows %lld", who, (long long)a->table_rows);
  return PMGT_OK;
}

extern "C" int pmgt_gather_proj_fwd(const pmgt_gather_proj_args* a, void* stream) {
  int rc = gp_check(a, "pmgt_gather_proj_fwd");
  if (rc) return rc;
  if (a->T == 0) return PMGT_OK;
  PMGT_REQUIRE(a->w && a->out, "pmgt_gather_proj_fwd: null w / out");
  PMGT_REQUIRE(a->ldw % 8 == 0 && ((uintptr_t)a->w & 15) == 0, "pmgt_gather_proj_fwd: w must be 16-byte aligned");
  PMGT_REQUIRE(a->ldo % 16 == 0 && ((uintptr_t)a->out & 31) == 0, "pmgt_gather_proj_fwd: out rows must be 32-byte aligned");
  CUtensorMap tw;
  memset(&tw, 0, sizeof(tw));
  rc = make_tmap(&tw, a->w, a->K, 128, a->ldw, 64, 128);
  if (rc) return rc;
  GpFwdParams kp;
  kp.T = (int)a->T; kp.K = (int)a->K; kp.num_tiles = (int)((a->T + 127) / 128);
  kp.table = a->table; kp.ld = a->ld; kp.rows = (const long long*)a->rows; kp.src_rows = a->table_rows;
  kp.out = a->out; kp.ldo = a->ldo; kp.bias = a->bias;
  const int grid = kp.num_tiles < num_sms() ? kp.num_tiles : num_sms();
  CUtensorMap ttab;
  memset(&ttab, 0, sizeof(ttab));
  cudaStream_t st = (cudaStream_t)stream;
  if (gather_tma_mode() & 1) {
    rc = make_tmap(&ttab, a->table, a->K, a->table_rows, a->ld, 64, 1);
    if (rc) return rc;
    const char* e = getenv("PMGT_GATHER_FWD_WARPS");  // measurement switch: issuing warps of the forward kernel
    if ((e && *e ? atoi(e) : kGpFwdWarpsDefault) == 16) return launch_fwd<true, 5, 3, 16>(tw, ttab, kp, grid, st);
    return launch_fwd<true, 5, 3, 8>(tw, ttab, kp, grid, st);
  }
  return launch_fwd<false, 5, 3, 4>(tw, ttab, kp, grid, st);
}

extern "C" int pmgt_gather_proj_dw(const pmgt_gather_proj_args* a, void* stream) {
  int rc = gp_check(a, "pmgt_gather_proj_dw");
  if (rc) return rc;
  if (a->T == 0) return PMGT_OK;
  PMGT_REQUIRE(a->dy && a->dw, "pmgt_gather_proj_dw: null dy / dw");
  PMGT_REQUIRE(a->ld_dy % 8 == 0 && ((uintptr_t)a->dy & 15) == 0, "pmgt_gather_proj_dw: dy must be 16-byte aligned");
  PMGT_REQUIRE(a->ld_dw % 4 == 0 && ((uintptr_t)a->dw & 15) == 0, "pmgt_gather_proj_dw: dw must be 16-byte aligned");
  CUtensorMap tdy;
  memset(&tdy, 0, sizeof(tdy));
  rc = make_tmap(&tdy, a->dy, 128, a->T, a->ld_dy, 64, kGpDwTok);
  if (rc) return rc;
  const int nc = dw_slabs(a->K);
  const int col_groups = (int)(a->K / (nc * 64));
  int ranges = num_sms() / col_groups;
  if (ranges < 1) ranges = 1;
  long long tpc = (a->T + ranges - 1) / ranges;
  tpc = ((tpc + kGpDwTok - 1) / kGpDwTok) * kGpDwTok;
  ranges = (int)((a->T + tpc - 1) / tpc);
  GpDwParams kp;
  kp.T = (int)a->T; kp.K = (int)a->K; kp.tok_per_cta = (int)tpc;
  kp.table = a->table; kp.ld = a->ld; kp.rows = (const long long*)a->rows; kp.src_rows = a->table_rows;
  kp.dw = a->dw; kp.ld_dw = a->ld_dw;
  dim3 grid((unsigned)col_groups, (unsigned)ranges, 1);
  cudaStream_t st = (cudaStream_t)stream;
  CUtensorMap ttab;
  memset(&ttab, 0, sizeof(ttab));
  if (gather_tma_mode() & 2) {
    rc = make_tmap(&ttab, a->table, a->K, a->table_rows, a->ld, 64, 1);
    if (rc) return rc;
    const char* e = getenv("PMGT_GATHER_DW_WARPS");  // measurement switch: issuing warps of the weight-gradient kernel
    if ((e && *e ? atoi(e) : kGpDwWarpsDefault) == 16) {
      if (nc == 8) return launch_dw<8, true, 16>(tdy, ttab, kp, grid, st);
      if (nc == 6) return launch_dw<6, true, 16>(tdy, ttab, kp, grid, st);
      return launch_dw<4, true, 16>(tdy, ttab, kp, grid, st);
    }
    if (nc == 8) return launch_dw<8, true, 8>(tdy, ttab, kp, grid, st);
    if (nc == 6) return launch_dw<6, true, 8>(tdy, ttab, kp, grid, st);
    return launch_dw<4, true, 8>(tdy, ttab, kp, grid, st);
  }
  if (nc == 8) return launch_dw<8, false, 4>(tdy, ttab, kp, grid, st);
  if (nc == 6) return launch_dw<6, false, 4>(tdy, ttab, kp, grid, st);
  return launch_dw<4, false, 4>(tdy, ttab, kp, grid, st);
}
