// K1 -- MCNSampling on the device: CSR item graph, per-row softmax CDF,
// counter-based Philox4x32-10.  Replaces pmgt/pmgt/datasets.py:14-79 (context
// sampler) and :125-183 (positive / negative pair selection) of the reference.
//
// One CTA (128 threads) builds one context:
//   hop k: n_k = s_1*...*s_k draws; draw d has parent d / s_k in the previous
//          hop's list; u -> upper_bound over the parent's CDF slice -> neighbour.
//   scoring: smem open-addressing table  node -> (score += depth-k+1,
//            first = min(global draw index)); the root itself is not scored.
//   top-k: max_ctx rounds of a block-wide arg-max over
//          key = (score << 32) | ~first  (score desc, first appearance asc --
//          Python's stable sorted(..., reverse=True) over dict insertion order).
// The CPU replay of the same stream is oracle/philox_sampler.c.
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "common.cuh"

namespace pmgt {

struct pmgt_graph_impl {
  int device;
  int64_t num_nodes;
  int64_t num_edges;
  int64_t* indptr;
  int32_t* indices;
  float* cdf;
  // Inverse-CDF acceleration, built once in pmgt_graph_create (see find_neighbour):
  uint2* ec;         // [E] (cdf bits, neighbour id) interleaved: the probe that ends the search also delivers the id
  uint16_t* guide;   // [E] guide[rs + j] = #{i : cdf[rs + i] <= j / deg}; NULL when some row has more than 65535 entries
};

constexpr int kSamplerThreads = 128;
constexpr int kMaxDepth = 8;

struct SamplerParams {
  const int64_t* indptr;
  const int32_t* indices;
  const float* cdf;
  const uint2* ec;
  const uint16_t* guide;
  const int64_t* roots;
  const int64_t* keys;
  int64_t n_ctx;
  int64_t n_node_ids;  // num_nodes + 2
  int depth;
  int hops[kMaxDepth];
  int max_ctx;
  int list_cap;   // capacity of each hop list (largest stored level)
  int table_cap;  // power of two
  int table_shift;
  uint32_t seed_lo, seed_hi;
  int64_t* out_ids;
  float* out_mask;
  int64_t* out_visited_deg;
};

__device__ __forceinline__ int upper_bound_cdf(const float* __restrict__ cdf, int n, float u) {
  // number of entries <= u  (numpy searchsorted(side="right")), clamped to n-1
  int lo = 0, hi = n;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (__ldg(cdf + mid) <= u) lo = mid + 1; else hi = mid;
  }
  return lo < n ? lo : n - 1;
}

// Inverse-CDF draw with a guide table (Chen & Asau 1974): numpy's legacy choice() returns
// searchsorted(cdf, u, side="right") -- the number of entries <= u -- clamped to deg - 1.  With u = u24 / 2^24 and
// j = floor(u24 * deg / 2^24) (exact integer arithmetic) we have j / deg <= u, so guide[j] = #{cdf <= j / deg} is a lower
// bound of the answer and a forward scan from there ends after ~1 entry on average: the dependent-load chain of a
// draw shrinks from 2 + ceil(log2 deg) (row pointers, binary search, neighbour id) to ~4 (row pointers, guide entry,
// one or two (cdf, id) pairs) -- on graphs that spill out of L2 every link of that chain is an HBM round trip, and the
// kernel is bound by exactly that latency.  The result is identical to the binary search, entry for entry.
__device__ __forceinline__ int32_t find_neighbour(const SamplerParams& p, int64_t rs, int deg, uint32_t word) {
  const uint32_t u24 = word >> 8;
  const float u = (float)u24 * (1.0f / 16777216.0f);
  if (p.guide == nullptr) return __ldg(p.indices + rs + upper_bound_cdf(p.cdf + rs, deg, u));
  const uint32_t j = (uint32_t)(((uint64_t)u24 * (uint64_t)(uint32_t)deg) >> 24);
  int pos = (int)__ldg(p.guide + rs + j);
  pos = pos < deg - 1 ? pos : deg - 1;
  uint2 e = __ldg(p.ec + rs + pos);
  while (pos < deg - 1 && __uint_as_float(e.x) <= u) {
    ++pos;
    e = __ldg(p.ec + rs + pos);
  }
  return (int32_t)e.y;
}

// SPT > 0: table_cap == SPT * 128 and max_ctx <= 8 -- the top-k keeps every thread's slots in registers, each warp
// extracts its own max_ctx best without block barriers, and warp 0 merges the 4 * max_ctx finalists (one barrier
// instead of two table scans and two barriers per output position).  SPT == 0: any table size / max_ctx.
// ILP: the late hops advance the four draws of a Philox block in lock-step (memory-level parallelism for graphs that
// spill out of L2: every step of a draw's pointer chase is then an HBM round trip).  On L2-resident graphs the kernel
// is issue-bound and the sequential form, with its smaller register footprint and higher occupancy, is faster.
template <int SPT, bool ILP>
__global__ void __launch_bounds__(kSamplerThreads)
sample_contexts_kernel(const SamplerParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  int32_t* list_a = reinterpret_cast<int32_t*>(smem_raw);
  int32_t* list_b = list_a + p.list_cap;
  int32_t* tkeys = list_a + ((2 * p.list_cap + 3) & ~3);  // 16-byte aligned (vector clear)
  uint32_t* tscore = reinterpret_cast<uint32_t*>(tkeys + p.table_cap);
  uint32_t* tfirst = tscore + p.table_cap;  // holds 0xffffffff - (first appearance), so an empty table is all zero
  __shared__ unsigned long long red[kSamplerThreads / 32];
  __shared__ unsigned long long best_sh;
  __shared__ unsigned long long deg_sh;
  __shared__ unsigned long long cand_k[32];
  __shared__ int32_t cand_n[32];

  const int tid = threadIdx.x;
  const int L = p.max_ctx + 1;

  for (int64_t ctx = blockIdx.x; ctx < p.n_ctx; ctx += gridDim.x) {
    const int64_t root64 = p.roots[ctx];
    const int32_t root = (int32_t)root64;
    const uint64_t key = (uint64_t)p.keys[ctx];
    const uint32_t key_lo = (uint32_t)key, key_hi = (uint32_t)(key >> 32);

    {  // keys, scores and first-appearance words are contiguous: one 16-byte-vector clear
      uint4* t4 = reinterpret_cast<uint4*>(tkeys);
      for (int i = tid; i < p.table_cap * 3 / 4; i += kSamplerThreads) t4[i] = make_uint4(0, 0, 0, 0);
    }
    if (tid == 0) deg_sh = 0ull;
    __syncthreads();

    int32_t* prev = list_a;
    int32_t* cur = list_b;
    uint32_t base = 0;       // global draw index of the first draw of this hop
    uint32_t n_prev = 1;     // number of parents
    unsigned long long my_deg = 0;
    const bool root_ok = root64 >= 2 && root64 < p.n_node_ids;

    for (int k = 1; k <= p.depth; ++k) {
      const uint32_t s = (uint32_t)p.hops[k - 1];
      const uint32_t n_k = n_prev * s;
      const uint32_t hop_w = (uint32_t)(p.depth - k + 1);
      const bool store = k < p.depth;
      // one draw: parent row -> inverse-CDF position -> neighbour -> score table
      auto draw = [&](uint32_t gd, uint32_t word) {
        const uint32_t d = gd - base;
        const uint32_t ppos = d / s;
        const int32_t parent = (k == 1) ? (root_ok ? root : 0) : prev[ppos];
        int32_t nb = 0;
        if (parent != 0) {
          const int64_t rs = __ldg(p.indptr + parent);
          const int deg = (int)(__ldg(p.indptr + parent + 1) - rs);
          if (d - ppos * s == 0) my_deg += (unsigned long long)deg;
          if (deg > 0) nb = find_neighbour(p, rs, deg, word);
        }
        if (store) cur[d] = nb;
        if (nb != 0 && nb != root) {
          uint32_t slot = ((uint32_t)nb * 2654435761u) >> p.table_shift;
          while (true) {
            int32_t old = atomicCAS(&tkeys[slot], 0, nb);
            if (old == 0 || old == nb) break;
            slot = (slot + 1) & (uint32_t)(p.table_cap - 1);
          }
          atomicAdd(&tscore[slot], hop_w);
          atomicMax(&tfirst[slot], 0xffffffffu - gd);
        }
      };
      if (n_k <= (uint32_t)kSamplerThreads) {
        // early hops (16 and 128 draws with the default sizes): one THREAD per draw.  Up to four threads recompute the
        // same Philox block, but the dependent-load chain of every draw (row pointer -> CDF search -> neighbour ->
        // table) runs in parallel instead of four deep per thread; these hops were barrier-stall bound.
        if ((uint32_t)tid < n_k) {
          const uint32_t gd = base + (uint32_t)tid;
          const Philox4 r = philox4x32_10(gd >> 2, PMGT_STREAM_CTX, key_lo, key_hi, p.seed_lo, p.seed_hi);
          draw(gd, philox_word(r, (int)(gd & 3u)));
        }
      } else if constexpr (ILP) {
        // late hops: one Philox block = four draws per thread, advanced in LOCK-STEP so that their dependent-load
        // chains (row pointers -> CDF binary search -> neighbour id) overlap four deep instead of running one after
        // the other; on graphs that do not fit in L2 every step of such a chain is an HBM round trip
        const uint32_t q_lo = base >> 2, q_hi = (base + n_k - 1) >> 2;
        for (uint32_t q = q_lo + tid; q <= q_hi; q += kSamplerThreads) {
          const Philox4 r = philox4x32_10(q, PMGT_STREAM_CTX, key_lo, key_hi, p.seed_lo, p.seed_hi);
          bool act[4];
          uint32_t dd[4];
          int32_t par[4];
          int64_t rs[4];
          int dg[4], lo[4], hi[4];
          float uu[4];
#pragma unroll
          for (int w = 0; w < 4; ++w) {
            const uint32_t gd = (q << 2) + w;
            act[w] = gd >= base && gd < base + n_k;
            dd[w] = gd - base;
            par[w] = 0;
            if (act[w]) par[w] = (k == 1) ? (root_ok ? root : 0) : prev[dd[w] / s];
          }
#pragma unroll
          for (int w = 0; w < 4; ++w) {
            rs[w] = 0;
            dg[w] = 0;
            if (par[w] != 0) {
              rs[w] = __ldg(p.indptr + par[w]);
              dg[w] = (int)(__ldg(p.indptr + par[w] + 1) - rs[w]);
            }
          }
#pragma unroll
          for (int w = 0; w < 4; ++w) {
            if (par[w] != 0 && dd[w] % s == 0) my_deg += (unsigned long long)dg[w];
            uu[w] = (float)(philox_word(r, w) >> 8) * (1.0f / 16777216.0f);
            lo[w] = 0;
            hi[w] = dg[w];
          }
          int32_t nb[4];
          if (p.guide != nullptr) {
            // guide entry -> first (cdf, id) pair -> forward scan, the four draws in lock-step
            uint2 e[4];
            int pos[4];
#pragma unroll
            for (int w = 0; w < 4; ++w) {
              pos[w] = 0;
              if (dg[w] > 0) {
                const uint32_t u24 = philox_word(r, w) >> 8;
                const uint32_t j = (uint32_t)(((uint64_t)u24 * (uint64_t)(uint32_t)dg[w]) >> 24);
                pos[w] = (int)__ldg(p.guide + rs[w] + j);
              }
            }
#pragma unroll
            for (int w = 0; w < 4; ++w) {
              e[w] = make_uint2(0x7f800000u, 0u);  // +inf: an inactive draw never advances
              if (dg[w] > 0) {
                pos[w] = pos[w] < dg[w] - 1 ? pos[w] : dg[w] - 1;
                e[w] = __ldg(p.ec + rs[w] + pos[w]);
              }
            }
            while (true) {
              bool adv[4];
              bool any = false;
#pragma unroll
              for (int w = 0; w < 4; ++w) {
                adv[w] = dg[w] > 0 && pos[w] < dg[w] - 1 && __uint_as_float(e[w].x) <= uu[w];
                any |= adv[w];
              }
              if (!any) break;
#pragma unroll
              for (int w = 0; w < 4; ++w)
                if (adv[w]) {
                  ++pos[w];
                  e[w] = __ldg(p.ec + rs[w] + pos[w]);
                }
            }
#pragma unroll
            for (int w = 0; w < 4; ++w) nb[w] = dg[w] > 0 ? (int32_t)e[w].y : 0;
          } else {
          while ((lo[0] < hi[0]) | (lo[1] < hi[1]) | (lo[2] < hi[2]) | (lo[3] < hi[3])) {
            float cv[4];
            int mid[4];
#pragma unroll
            for (int w = 0; w < 4; ++w) {
              mid[w] = (lo[w] + hi[w]) >> 1;
              cv[w] = lo[w] < hi[w] ? __ldg(p.cdf + rs[w] + mid[w]) : 0.f;
            }
#pragma unroll
            for (int w = 0; w < 4; ++w)
              if (lo[w] < hi[w]) {
                if (cv[w] <= uu[w]) lo[w] = mid[w] + 1; else hi[w] = mid[w];
              }
          }
#pragma unroll
          for (int w = 0; w < 4; ++w)
            nb[w] = dg[w] > 0 ? __ldg(p.indices + rs[w] + (lo[w] < dg[w] ? lo[w] : dg[w] - 1)) : 0;
          }
#pragma unroll
          for (int w = 0; w < 4; ++w) {
            if (!act[w]) continue;
            if (store) cur[dd[w]] = nb[w];
            if (nb[w] != 0 && nb[w] != root) {
              const uint32_t gd = (q << 2) + w;
              uint32_t slot = ((uint32_t)nb[w] * 2654435761u) >> p.table_shift;
              while (true) {
                int32_t old = atomicCAS(&tkeys[slot], 0, nb[w]);
                if (old == 0 || old == nb[w]) break;
                slot = (slot + 1) & (uint32_t)(p.table_cap - 1);
              }
              atomicAdd(&tscore[slot], hop_w);
              atomicMax(&tfirst[slot], 0xffffffffu - gd);
            }
          }
        }
      } else {
        const uint32_t q_lo = base >> 2, q_hi = (base + n_k - 1) >> 2;
        for (uint32_t q = q_lo + tid; q <= q_hi; q += kSamplerThreads) {
          const Philox4 r = philox4x32_10(q, PMGT_STREAM_CTX, key_lo, key_hi, p.seed_lo, p.seed_hi);
#pragma unroll
          for (int w = 0; w < 4; ++w) {
            const uint32_t gd = (q << 2) + w;
            if (gd < base || gd >= base + n_k) continue;
            draw(gd, philox_word(r, w));
          }
        }
      }
      __syncthreads();
      int32_t* t = prev; prev = cur; cur = t;
      base += n_k;
      n_prev = n_k;
    }

    if (p.out_visited_deg) {
      for (int o = 16; o > 0; o >>= 1) my_deg += __shfl_xor_sync(0xffffffffu, my_deg, o);
      if ((tid & 31) == 0 && my_deg) atomicAdd(&deg_sh, my_deg);
    }

    // top-k by (score desc, first asc)
    unsigned long long bound = ~0ull;
    if (tid == 0) {
      p.out_ids[ctx * L] = root64;
      p.out_mask[ctx * L] = 1.0f;
    }
    if constexpr (SPT > 0) {
      unsigned long long mk[SPT];
      int32_t mn[SPT];
#pragma unroll
      for (int s2 = 0; s2 < SPT; ++s2) {
        const int i = tid + s2 * kSamplerThreads;
        mn[s2] = tkeys[i];
        mk[s2] = mn[s2] != 0 ? (((unsigned long long)tscore[i] << 32) | (unsigned long long)tfirst[i]) : 0ull;
      }
      const int lane = tid & 31, wrp = tid >> 5;
      for (int r = 0; r < p.max_ctx; ++r) {
        unsigned long long lb = 0ull;
#pragma unroll
        for (int s2 = 0; s2 < SPT; ++s2) lb = mk[s2] > lb ? mk[s2] : lb;
        unsigned long long wb = lb;
        for (int o = 16; o > 0; o >>= 1) {
          const unsigned long long other = __shfl_xor_sync(0xffffffffu, wb, o);
          wb = other > wb ? other : wb;
        }
        if (wb != 0ull && lb == wb) {  // keys are unique: exactly one lane of the warp
#pragma unroll
          for (int s2 = 0; s2 < SPT; ++s2)
            if (mk[s2] == wb) { cand_k[wrp * 8 + r] = wb; cand_n[wrp * 8 + r] = mn[s2]; mk[s2] = 0ull; }
        }
        if (wb == 0ull && lane == 0) { cand_k[wrp * 8 + r] = 0ull; cand_n[wrp * 8 + r] = 0; }
      }
      __syncthreads();
      if (wrp == 0) {
        const int cw = lane >> 3, cr = lane & 7;
        unsigned long long ck = cr < p.max_ctx ? cand_k[cw * 8 + cr] : 0ull;
        const int32_t cn = cr < p.max_ctx ? cand_n[cw * 8 + cr] : 0;
        for (int r = 0; r < p.max_ctx; ++r) {
          unsigned long long wb = ck;
          for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long other = __shfl_xor_sync(0xffffffffu, wb, o);
            wb = other > wb ? other : wb;
          }
          if (wb != 0ull && ck == wb) {
            p.out_ids[ctx * L + 1 + r] = (int64_t)cn;
            p.out_mask[ctx * L + 1 + r] = 1.0f;
            ck = 0ull;
          }
          if (wb == 0ull && lane == 0) {
            p.out_ids[ctx * L + 1 + r] = 0;
            p.out_mask[ctx * L + 1 + r] = 0.0f;
          }
        }
      }
    } else
    for (int r = 0; r < p.max_ctx; ++r) {
      unsigned long long best = 0ull;
      for (int i = tid; i < p.table_cap; i += kSamplerThreads) {
        if (tkeys[i] != 0) {
          unsigned long long kk = ((unsigned long long)tscore[i] << 32) | (unsigned long long)tfirst[i];
          if (kk < bound && kk > best) best = kk;
        }
      }
      for (int o = 16; o > 0; o >>= 1) {
        unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
        best = other > best ? other : best;
      }
      if ((tid & 31) == 0) red[tid >> 5] = best;
      __syncthreads();
      if (tid == 0) {
        unsigned long long b = red[0];
        for (int i = 1; i < kSamplerThreads / 32; ++i) b = red[i] > b ? red[i] : b;
        best_sh = b;
      }
      __syncthreads();
      best = best_sh;
      // the winner's slot writes the node id (keys are unique: `first` is unique)
      if (best != 0ull) {
        for (int i = tid; i < p.table_cap; i += kSamplerThreads) {
          if (tkeys[i] != 0) {
            unsigned long long kk = ((unsigned long long)tscore[i] << 32) | (unsigned long long)tfirst[i];
            if (kk == best) {
              p.out_ids[ctx * L + 1 + r] = (int64_t)tkeys[i];
              p.out_mask[ctx * L + 1 + r] = 1.0f;
            }
          }
        }
        bound = best;
      } else {
        if (tid == 0) {
          p.out_ids[ctx * L + 1 + r] = 0;
          p.out_mask[ctx * L + 1 + r] = 0.0f;
        }
        bound = 0ull;
      }
    }
    __syncthreads();
    if (p.out_visited_deg && tid == 0) p.out_visited_deg[ctx] = (int64_t)deg_sh;
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------
// pair selection: one warp per target
// ---------------------------------------------------------------------------
constexpr int kPairMaxPos = 32;

struct PairParams {
  const int64_t* indptr;
  const int32_t* indices;
  const int64_t* targets;
  const int64_t* keys;
  int64_t n_tgt;
  int64_t num_nodes;
  int max_pos, min_neg, max_total, stride;
  uint32_t seed_lo, seed_hi;
  int64_t* out_pairs;
  float* out_labels;
  int64_t* out_num;
};

__device__ __forceinline__ uint32_t philox_draw(uint32_t draw, uint32_t stream, uint32_t key_lo,
                                                uint32_t key_hi, uint32_t s0, uint32_t s1) {
  Philox4 r = philox4x32_10(draw >> 2, stream, key_lo, key_hi, s0, s1);
  return philox_word(r, (int)(draw & 3));
}

__global__ void __launch_bounds__(128) sample_pairs_kernel(const PairParams p) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t n_warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t t = warp; t < p.n_tgt; t += n_warps) {
    const int64_t tgt = p.targets[t];
    const uint64_t key = (uint64_t)p.keys[t];
    const uint32_t key_lo = (uint32_t)key, key_hi = (uint32_t)(key >> 32);
    int64_t rs = 0; int deg = 0;
    if (tgt >= 2 && tgt < p.num_nodes + 2) {
      rs = __ldg(p.indptr + tgt);
      deg = (int)(__ldg(p.indptr + tgt + 1) - rs);
    }
    int64_t* row = p.out_pairs + t * p.stride;
    float* lab = p.out_labels + t * p.stride;
    for (int i = lane; i < p.stride; i += 32) { row[i] = 0; lab[i] = 0.0f; }
    __syncwarp();
    const int n_pos = p.max_pos < deg ? p.max_pos : deg;
    int n_neg = p.max_total - n_pos;
    if (n_neg < p.min_neg) n_neg = p.min_neg;
    if (lane == 0) {
      // partial Fisher-Yates over positions 0..deg-1 with a sparse swap list
      int sw_from[kPairMaxPos], sw_to[kPairMaxPos];
      int n_sw = 0;
      for (int i = 0; i < n_pos; ++i) {
        uint32_t w = philox_draw((uint32_t)i, PMGT_STREAM_POS, key_lo, key_hi, p.seed_lo, p.seed_hi);
        int j = i + (int)(((uint64_t)w * (uint64_t)(deg - i)) >> 32);
        // value currently at position j and at position i
        int vj = j, vi = i;
        for (int s = 0; s < n_sw; ++s) { if (sw_from[s] == j) vj = sw_to[s]; if (sw_from[s] == i) vi = sw_to[s]; }
        // perm[i] = vj ; perm[j] = vi
        bool found = false;
        for (int s = 0; s < n_sw; ++s) if (sw_from[s] == j) { sw_to[s] = vi; found = true; }
        if (!found && n_sw < kPairMaxPos) { sw_from[n_sw] = j; sw_to[n_sw] = vi; ++n_sw; }
        row[i] = (int64_t)__ldg(p.indices + rs + vj);
        lab[i] = 1.0f;
      }
      p.out_num[t] = (int64_t)(n_pos + n_neg);
    }
    for (int n = 0; n < n_neg; ++n) {
      int64_t cand = 0;
      for (int a = 0; a < PMGT_MAX_NEG_ATTEMPTS; ++a) {
        uint32_t w = philox_draw((uint32_t)(n * PMGT_MAX_NEG_ATTEMPTS + a), PMGT_STREAM_NEG, key_lo,
                                 key_hi, p.seed_lo, p.seed_hi);
        cand = 2 + (int64_t)(((uint64_t)w * (uint64_t)p.num_nodes) >> 32);
        bool hit = false;
        for (int j = lane; j < deg; j += 32) hit |= ((int64_t)__ldg(p.indices + rs + j) == cand);
        if (!__any_sync(0xffffffffu, hit)) break;
      }
      if (lane == 0) { row[n_pos + n] = cand; lab[n_pos + n] = 0.0f; }
    }
    __syncwarp();
  }
}

}  // namespace pmgt

using namespace pmgt;

extern "C" {

int pmgt_graph_create(pmgt_graph** out, int device, int64_t num_nodes, int64_t num_edges,
                      const int64_t* indptr_host, const int32_t* indices_host,
                      const float* cdf_host) {
  PMGT_REQUIRE(out && indptr_host && (num_edges == 0 || (indices_host && cdf_host)),
               "pmgt_graph_create: null argument");
  PMGT_REQUIRE(num_nodes > 0 && num_edges >= 0 && num_nodes < (int64_t)0x7fffffff - 2,
               "pmgt_graph_create: bad sizes (num_nodes=%lld num_edges=%lld)", (long long)num_nodes,
               (long long)num_edges);
  PMGT_REQUIRE(indptr_host[0] == 0 && indptr_host[num_nodes + 2] == num_edges,
               "pmgt_graph_create: indptr must have num_nodes+3 entries spanning [0, num_edges]");
  PMGT_CHECK_CUDA(cudaSetDevice(device));
  pmgt_graph_impl* g = new pmgt_graph_impl();
  g->device = device; g->num_nodes = num_nodes; g->num_edges = num_edges;
  g->indptr = nullptr; g->indices = nullptr; g->cdf = nullptr; g->ec = nullptr; g->guide = nullptr;
  // host-side build of the interleaved (cdf, id) array and the guide table (one pass over the edges)
  std::vector<uint2> ec((size_t)(num_edges > 0 ? num_edges : 1));
  std::vector<uint16_t> guide((size_t)(num_edges > 0 ? num_edges : 1));
  bool guide_ok = true;
  for (int64_t row = 0; row < num_nodes + 2; ++row) {
    const int64_t rs = indptr_host[row], m = indptr_host[row + 1] - rs;
    if (m > 65535) guide_ok = false;
    int64_t i = 0;
    for (int64_t j = 0; j < m; ++j) {
      uint32_t bits;
      memcpy(&bits, &cdf_host[rs + j], 4);
      ec[(size_t)(rs + j)] = make_uint2(bits, (uint32_t)indices_host[rs + j]);
      const double thr = (double)j / (double)m;
      while (i < m && (double)cdf_host[rs + i] <= thr) ++i;
      guide[(size_t)(rs + j)] = (uint16_t)(i > 65535 ? 65535 : i);
    }
  }
  cudaError_t e = cudaMalloc(&g->indptr, sizeof(int64_t) * (num_nodes + 3));
  if (e == cudaSuccess) e = cudaMalloc(&g->ec, sizeof(uint2) * ec.size());
  if (e == cudaSuccess && num_edges) e = cudaMemcpy(g->ec, ec.data(), sizeof(uint2) * (size_t)num_edges, cudaMemcpyHostToDevice);
  if (e == cudaSuccess && guide_ok) {
    e = cudaMalloc(&g->guide, sizeof(uint16_t) * guide.size());
    if (e == cudaSuccess && num_edges)
      e = cudaMemcpy(g->guide, guide.data(), sizeof(uint16_t) * (size_t)num_edges, cudaMemcpyHostToDevice);
  }
  if (e == cudaSuccess) e = cudaMalloc(&g->indices, sizeof(int32_t) * (num_edges > 0 ? num_edges : 1));
  if (e == cudaSuccess) e = cudaMalloc(&g->cdf, sizeof(float) * (num_edges > 0 ? num_edges : 1));
  if (e == cudaSuccess) e = cudaMemcpy(g->indptr, indptr_host, sizeof(int64_t) * (num_nodes + 3), cudaMemcpyHostToDevice);
  if (e == cudaSuccess && num_edges) e = cudaMemcpy(g->indices, indices_host, sizeof(int32_t) * num_edges, cudaMemcpyHostToDevice);
  if (e == cudaSuccess && num_edges) e = cudaMemcpy(g->cdf, cdf_host, sizeof(float) * num_edges, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    set_error("pmgt_graph_create: %s", cudaGetErrorString(e));
    cudaFree(g->indptr); cudaFree(g->indices); cudaFree(g->cdf); cudaFree(g->ec); cudaFree(g->guide);
    delete g;
    return PMGT_ERR_CUDA;
  }
  *out = reinterpret_cast<pmgt_graph*>(g);
  return PMGT_OK;
}

int pmgt_graph_destroy(pmgt_graph* gh) {
  if (!gh) return PMGT_OK;
  pmgt_graph_impl* g = reinterpret_cast<pmgt_graph_impl*>(gh);
  cudaFree(g->indptr); cudaFree(g->indices); cudaFree(g->cdf); cudaFree(g->ec); cudaFree(g->guide);
  delete g;
  return PMGT_OK;
}

int64_t pmgt_graph_num_nodes(const pmgt_graph* g) { return g ? reinterpret_cast<const pmgt_graph_impl*>(g)->num_nodes : -1; }
int64_t pmgt_graph_num_edges(const pmgt_graph* g) { return g ? reinterpret_cast<const pmgt_graph_impl*>(g)->num_edges : -1; }
const int64_t* pmgt_graph_indptr(const pmgt_graph* g) { return g ? reinterpret_cast<const pmgt_graph_impl*>(g)->indptr : nullptr; }
const int32_t* pmgt_graph_indices(const pmgt_graph* g) { return g ? reinterpret_cast<const pmgt_graph_impl*>(g)->indices : nullptr; }
const float* pmgt_graph_cdf(const pmgt_graph* g) { return g ? reinterpret_cast<const pmgt_graph_impl*>(g)->cdf : nullptr; }

int pmgt_sample_contexts(const pmgt_graph* gh, const int64_t* roots, const int64_t* ctx_keys,
                         int64_t n_ctx, const int32_t* hops_host, int depth, int max_ctx,
                         uint64_t seed, int64_t* out_ids, float* out_mask,
                         int64_t* out_visited_deg, void* stream) {
  PMGT_REQUIRE(gh && roots && ctx_keys && hops_host && out_ids && out_mask,
               "pmgt_sample_contexts: null argument");
  PMGT_REQUIRE(depth >= 1 && depth <= kMaxDepth, "pmgt_sample_contexts: depth must be in [1,%d]", kMaxDepth);
  PMGT_REQUIRE(max_ctx >= 1 && max_ctx <= 1024, "pmgt_sample_contexts: max_ctx out of range");
  if (n_ctx == 0) return PMGT_OK;
  PMGT_REQUIRE(n_ctx > 0, "pmgt_sample_contexts: negative n_ctx");
  const pmgt_graph_impl* g = reinterpret_cast<const pmgt_graph_impl*>(gh);
  SamplerParams p{};
  p.indptr = g->indptr; p.indices = g->indices; p.cdf = g->cdf;
  p.ec = g->ec;
  {
    static int use_guide = -1;  // PMGT_SAMPLER_GUIDE=0: plain binary search (A/B measurements)
    if (use_guide < 0) {
      const char* ev = getenv("PMGT_SAMPLER_GUIDE");
      use_guide = (ev && ev[0] == '0') ? 0 : 1;
    }
    p.guide = use_guide ? g->guide : nullptr;
  }
  p.roots = roots; p.keys = ctx_keys; p.n_ctx = n_ctx; p.n_node_ids = g->num_nodes + 2;
  p.depth = depth; p.max_ctx = max_ctx;
  int64_t level = 1, total = 0, stored = 1;
  for (int k = 0; k < depth; ++k) {
    PMGT_REQUIRE(hops_host[k] >= 1, "pmgt_sample_contexts: hop size must be >= 1");
    p.hops[k] = hops_host[k];
    level *= hops_host[k];
    total += level;
    PMGT_REQUIRE(total <= 16384, "pmgt_sample_contexts: more than 16384 draws per context");
    if (k < depth - 1) stored = level;
  }
  p.list_cap = (int)stored;
  int cap = 64;
  while (cap < (total * 3 + 1) / 2) cap <<= 1;
  p.table_cap = cap;
  int lg = 0; while ((1 << lg) < cap) ++lg;
  p.table_shift = 32 - lg;
  p.seed_lo = (uint32_t)seed; p.seed_hi = (uint32_t)(seed >> 32);
  p.out_ids = out_ids; p.out_mask = out_mask; p.out_visited_deg = out_visited_deg;
  size_t smem = sizeof(int32_t) * (size_t)((2 * p.list_cap + 3) & ~3) + 12 * (size_t)cap;
  PMGT_CHECK_CUDA(cudaSetDevice(g->device));
  // CSR footprint (row pointers + neighbour ids + CDF) against the L2 capacity decides the late-hop strategy
  const double csr_mb = ((double)g->num_edges * 8.0 + (double)g->num_nodes * 8.0) / 1.0e6;
  const bool ilp = csr_mb > 48.0;
  void (*kern)(const SamplerParams) = ilp ? sample_contexts_kernel<0, true> : sample_contexts_kernel<0, false>;
  if (max_ctx <= 8) {
    if (cap == 4 * kSamplerThreads) kern = ilp ? sample_contexts_kernel<4, true> : sample_contexts_kernel<4, false>;
    else if (cap == 8 * kSamplerThreads) kern = ilp ? sample_contexts_kernel<8, true> : sample_contexts_kernel<8, false>;
    else if (cap == 16 * kSamplerThreads) kern = ilp ? sample_contexts_kernel<16, true> : sample_contexts_kernel<16, false>;
  }
  if (smem > 48 * 1024)
    PMGT_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 1;
  PMGT_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kSamplerThreads, smem));
  if (occ < 1) occ = 1;
  // One context per CTA: contexts differ a lot in cost (degree of the visited rows, hash-table collisions), and the
  // hardware CTA scheduler balances them better than a persistent grid's static striding -- measured 0.77 -> 0.71 ms on
  // the 1M graph, 0.50 -> 0.44 ms on TG, and the short-lived CTAs of the side-stream launch interleave better with the
  // main stream's persistent kernels (1M step 6.16 -> 6.01 ms).  PMGT_SAMPLER_CTAS_PER_SM=n restores a persistent grid.
  int64_t grid = n_ctx;
  static const int persistent_ctas = [] { const char* e = getenv("PMGT_SAMPLER_CTAS_PER_SM"); return e ? atoi(e) : 0; }();
  if (persistent_ctas > 0) grid = (int64_t)num_sms() * (persistent_ctas < occ ? persistent_ctas : occ);
  if (grid > n_ctx) grid = n_ctx;
  kern<<<(unsigned)grid, kSamplerThreads, smem, (cudaStream_t)stream>>>(p);
  PMGT_LAUNCH_CHECK();
  return PMGT_OK;
}

int pmgt_sample_pairs(const pmgt_graph* gh, const int64_t* targets, const int64_t* tgt_keys,
                      int64_t n_tgt, int max_pos, int min_neg, int max_total, int pair_stride,
                      uint64_t seed, int64_t* out_pairs, float* out_labels,
                      int64_t* out_num_pairs, void* stream) {
  PMGT_REQUIRE(gh && targets && tgt_keys && out_pairs && out_labels && out_num_pairs,
               "pmgt_sample_pairs: null argument");
  PMGT_REQUIRE(max_pos >= 0 && max_pos <= kPairMaxPos, "pmgt_sample_pairs: max_pos must be in [0,%d]", kPairMaxPos);
  PMGT_REQUIRE(min_neg >= 0 && max_total >= 0, "pmgt_sample_pairs: negative sizes");
  int need = max_pos + (min_neg > max_total ? min_neg : max_total);
  PMGT_REQUIRE(pair_stride >= 1 && pair_stride <= 4096, "pmgt_sample_pairs: bad pair_stride");
  {
    // worst case row length: n_pos + max(min_neg, max_total - n_pos) <= max(max_pos + min_neg, max_total)
    int worst = max_pos + min_neg > max_total ? max_pos + min_neg : max_total;
    PMGT_REQUIRE(pair_stride >= worst, "pmgt_sample_pairs: pair_stride %d < %d", pair_stride, worst);
    (void)need;
  }
  if (n_tgt == 0) return PMGT_OK;
  const pmgt_graph_impl* g = reinterpret_cast<const pmgt_graph_impl*>(gh);
  PairParams p{};
  p.indptr = g->indptr; p.indices = g->indices; p.targets = targets; p.keys = tgt_keys;
  p.n_tgt = n_tgt; p.num_nodes = g->num_nodes;
  p.max_pos = max_pos; p.min_neg = min_neg; p.max_total = max_total; p.stride = pair_stride;
  p.seed_lo = (uint32_t)seed; p.seed_hi = (uint32_t)(seed >> 32);
  p.out_pairs = out_pairs; p.out_labels = out_labels; p.out_num = out_num_pairs;
  PMGT_CHECK_CUDA(cudaSetDevice(g->device));
  int64_t blocks = (n_tgt + 3) / 4;
  int64_t cap = (int64_t)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  sample_pairs_kernel<<<(unsigned)blocks, 128, 0, (cudaStream_t)stream>>>(p);
  PMGT_LAUNCH_CHECK();
  return PMGT_OK;
}

}  // extern "C"
