/*
 * TEST INFRASTRUCTURE ONLY -- CPU replay of the Philox-driven MCNSampling stream.
 *
 * Plain-C restatement of what pmgt_sample_contexts / pmgt_sample_pairs
 * (include/pmgt_b200.h) must produce, written independently of the CUDA
 * kernels (linear candidate list instead of a hash table, selection sort
 * instead of block arg-max) so that agreement is a real cross-check.
 * The algorithm follows the reference sampler:
 *   - multi-hop weighted draws with replacement through the per-row softmax
 *     CDF, numpy legacy choice = cdf.searchsorted(u, side="right")
 *     (pmgt/pmgt/datasets.py:24-33);
 *   - per hop, score[node] += freq * (depth - k + 1), skipping the target
 *     (datasets.py:35-40);
 *   - stable sort by score descending, i.e. ties keep first-appearance order
 *     (datasets.py:42); truncate / right-pad with 0 (datasets.py:46-51);
 *   - positives: distinct neighbours, uniform without replacement
 *     (datasets.py:167-171); negatives: uniform node ids in [2, N+2) rejected
 *     while adjacent to the target (datasets.py:173-180).
 * Only the random stream differs from the reference (np.random Mersenne
 * Twister there, counter-based Philox4x32-10 here), which is why parity with
 * the reference itself is statistical (tests/test_sampler_stats.py) while
 * parity with this replay is bit-exact (tests/test_sampler_gpu.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
 * load this file's shared object.  Build: see oracle/build.py.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define STREAM_CTX 0u
#define STREAM_POS 1u
#define STREAM_NEG 2u
#define MAX_NEG_ATTEMPTS 64

static void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}

static uint32_t draw_word(uint32_t draw, uint32_t stream, uint64_t key, uint64_t seed) {
  uint32_t c[4] = {draw >> 2, stream, (uint32_t)key, (uint32_t)(key >> 32)};
  philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  return c[draw & 3u];
}

/* exported for known-answer tests of the generator itself */
void pmgt_oracle_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                        uint32_t out[4]) {
  uint32_t c[4] = {c0, c1, c2, c3};
  philox4x32_10(c, k0, k1);
  memcpy(out, c, sizeof(c));
}

static int searchsorted_right(const float* cdf, int n, float u) {
  int lo = 0, hi = n;
  while (lo < hi) {
    int mid = (lo + hi) / 2;
    if (cdf[mid] <= u) lo = mid + 1; else hi = mid;
  }
  return lo < n ? lo : n - 1;
}

int pmgt_oracle_sample_contexts(const int64_t* indptr, const int32_t* indices, const float* cdf,
                                int64_t num_nodes, const int64_t* roots, const int64_t* keys,
                                int64_t n_ctx, const int32_t* hops, int depth, int max_ctx,
                                uint64_t seed, int64_t* out_ids, float* out_mask,
                                int64_t* out_visited_deg) {
  int64_t total = 0, level = 1;
  for (int k = 0; k < depth; ++k) { level *= hops[k]; total += level; }
  int32_t* prev = (int32_t*)malloc(sizeof(int32_t) * (size_t)(level > 0 ? level : 1));
  int32_t* cur = (int32_t*)malloc(sizeof(int32_t) * (size_t)(level > 0 ? level : 1));
  int32_t* cand = (int32_t*)malloc(sizeof(int32_t) * (size_t)total);
  int64_t* score = (int64_t*)malloc(sizeof(int64_t) * (size_t)total);
  int64_t* first = (int64_t*)malloc(sizeof(int64_t) * (size_t)total);
  char* taken = (char*)malloc((size_t)total);
  if (!prev || !cur || !cand || !score || !first || !taken) return -1;
  const int L = max_ctx + 1;

  for (int64_t c = 0; c < n_ctx; ++c) {
    const int64_t root = roots[c];
    const uint64_t key = (uint64_t)keys[c];
    int64_t n_cand = 0, visited = 0;
    uint32_t gd = 0;
    int64_t n_prev = 1;
    prev[0] = (root >= 2 && root < num_nodes + 2) ? (int32_t)root : 0;
    for (int k = 1; k <= depth; ++k) {
      const int s = hops[k - 1];
      for (int64_t pp = 0; pp < n_prev; ++pp) {
        const int32_t parent = prev[pp];
        int64_t rs = 0; int deg = 0;
        if (parent != 0) { rs = indptr[parent]; deg = (int)(indptr[parent + 1] - rs); visited += deg; }
        for (int i = 0; i < s; ++i, ++gd) {
          int32_t nb = 0;
          if (deg > 0) {
            uint32_t w = draw_word(gd, STREAM_CTX, key, seed);
            float u = (float)(w >> 8) * (1.0f / 16777216.0f);
            nb = indices[rs + searchsorted_right(cdf + rs, deg, u)];
          }
          cur[pp * s + i] = nb;
          if (nb != 0 && (int64_t)nb != root) {
            int64_t j = 0;
            while (j < n_cand && cand[j] != nb) ++j;
            if (j == n_cand) { cand[j] = nb; score[j] = 0; first[j] = gd; ++n_cand; }
            score[j] += depth - k + 1;
          }
        }
      }
      n_prev *= s;
      int32_t* t = prev; prev = cur; cur = t;
    }
    /* candidates are already in first-appearance order; stable selection of the
       max_ctx best scores == stable descending sort + truncate */
    memset(taken, 0, (size_t)n_cand);
    out_ids[c * L] = root; out_mask[c * L] = 1.0f;
    for (int r = 0; r < max_ctx; ++r) {
      int64_t best = -1;
      for (int64_t j = 0; j < n_cand; ++j)
        if (!taken[j] && (best < 0 || score[j] > score[best])) best = j;
      if (best >= 0) { taken[best] = 1; out_ids[c * L + 1 + r] = cand[best]; out_mask[c * L + 1 + r] = 1.0f; }
      else { out_ids[c * L + 1 + r] = 0; out_mask[c * L + 1 + r] = 0.0f; }
    }
    if (out_visited_deg) out_visited_deg[c] = visited;
  }
  free(prev); free(cur); free(cand); free(score); free(first); free(taken);
  return 0;
}

int pmgt_oracle_sample_pairs(const int64_t* indptr, const int32_t* indices, int64_t num_nodes,
                             const int64_t* targets, const int64_t* keys, int64_t n_tgt, int max_pos,
                             int min_neg, int max_total, int stride, uint64_t seed,
                             int64_t* out_pairs, float* out_labels, int64_t* out_num) {
  for (int64_t t = 0; t < n_tgt; ++t) {
    const int64_t tgt = targets[t];
    const uint64_t key = (uint64_t)keys[t];
    int64_t rs = 0; int deg = 0;
    if (tgt >= 2 && tgt < num_nodes + 2) { rs = indptr[tgt]; deg = (int)(indptr[tgt + 1] - rs); }
    int64_t* row = out_pairs + t * stride;
    float* lab = out_labels + t * stride;
    for (int i = 0; i < stride; ++i) { row[i] = 0; lab[i] = 0.0f; }
    const int n_pos = max_pos < deg ? max_pos : deg;
    int n_neg = max_total - n_pos;
    if (n_neg < min_neg) n_neg = min_neg;
    /* explicit Fisher-Yates on a materialised permutation */
    int* perm = (int*)malloc(sizeof(int) * (size_t)(deg > 0 ? deg : 1));
    if (!perm) return -1;
    for (int i = 0; i < deg; ++i) perm[i] = i;
    for (int i = 0; i < n_pos; ++i) {
      uint32_t w = draw_word((uint32_t)i, STREAM_POS, key, seed);
      int j = i + (int)(((uint64_t)w * (uint64_t)(deg - i)) >> 32);
      int tmp = perm[i]; perm[i] = perm[j]; perm[j] = tmp;
      row[i] = indices[rs + perm[i]];
      lab[i] = 1.0f;
    }
    free(perm);
    for (int n = 0; n < n_neg; ++n) {
      int64_t cand = 0;
      for (int a = 0; a < MAX_NEG_ATTEMPTS; ++a) {
        uint32_t w = draw_word((uint32_t)(n * MAX_NEG_ATTEMPTS + a), STREAM_NEG, key, seed);
        cand = 2 + (int64_t)(((uint64_t)w * (uint64_t)num_nodes) >> 32);
        int hit = 0;
        for (int j = 0; j < deg && !hit; ++j) hit = ((int64_t)indices[rs + j] == cand);
        if (!hit) break;
      }
      row[n_pos + n] = cand;
      lab[n_pos + n] = 0.0f;
    }
    out_num[t] = n_pos + n_neg;
  }
  return 0;
}
