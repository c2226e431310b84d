// K3 attention core, register-resident variant for short sequences (L <= 8) and
// head sizes 32 / 64 / 128: the default PMGT shape is L = 6, one head of 128
// (pmgt/pmgt/modeling_pmgt.py:435-526; formulas in attention.cu).
//
// G lanes of a warp own one (sequence, head): each lane holds DH/G feature
// dimensions of every row of Q/K/V/C (16-byte coalesced loads, a whole 256-byte
// row per group of lanes), the L x L score matrices are reduced over the group
// with xor-shuffles and then live in registers of every lane, so the softmax /
// cosine algebra needs no shared memory at all.  Backward recomputes the
// probabilities and re-reads Q/K/C for the output products (L1/L2 hits).
//
// HBM-bound by design: forward 10*H bytes per token, backward 18*H bytes per
// token (bf16), no intermediate ever written.
#include <stdlib.h>

#include "common.cuh"

namespace pmgt {

__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  unpack_bf16x2(u.x, f[0], f[1]);
  unpack_bf16x2(u.y, f[2], f[3]);
  unpack_bf16x2(u.z, f[4], f[5]);
  unpack_bf16x2(u.w, f[6], f[7]);
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 u;
  u.x = pack_bf16x2(f[0], f[1]);
  u.y = pack_bf16x2(f[2], f[3]);
  u.z = pack_bf16x2(f[4], f[5]);
  u.w = pack_bf16x2(f[6], f[7]);
  return u;
}

template <int G>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// rows [L][DPL] of one tensor of one item: lane t of the group owns the 8-element chunks (m * G + t), m < NCH
template <int L, int DPL, int G>
__device__ __forceinline__ void load_rows(const uint16_t* __restrict__ src, long long ld, int t, float (&f)[L][DPL]) {
  constexpr int NCH = DPL / 8;
  uint4 u[L][NCH];
#pragma unroll
  for (int i = 0; i < L; ++i)
#pragma unroll
    for (int m = 0; m < NCH; ++m) u[i][m] = *reinterpret_cast<const uint4*>(src + (long long)i * ld + (m * G + t) * 8);
#pragma unroll
  for (int i = 0; i < L; ++i)
#pragma unroll
    for (int m = 0; m < NCH; ++m) unpack8(u[i][m], &f[i][m * 8]);
}

template <int DPL, int G>
__device__ __forceinline__ void load_row(const uint16_t* __restrict__ src, int t, float (&f)[DPL]) {
  constexpr int NCH = DPL / 8;
  uint4 u[NCH];
#pragma unroll
  for (int m = 0; m < NCH; ++m) u[m] = *reinterpret_cast<const uint4*>(src + (m * G + t) * 8);
#pragma unroll
  for (int m = 0; m < NCH; ++m) unpack8(u[m], &f[m * 8]);
}

template <int DPL, int G>
__device__ __forceinline__ void store_row(uint16_t* __restrict__ dst, int t, const float (&f)[DPL]) {
  constexpr int NCH = DPL / 8;
#pragma unroll
  for (int m = 0; m < NCH; ++m) *reinterpret_cast<uint4*>(dst + (m * G + t) * 8) = pack8(&f[m * 8]);
}

// keep-bit e of the result = dropout_keep(seed, site, base_idx + e, p) for e < LL (<= 64); the Philox blocks
// (8 elements each) are spread over the G lanes of the group and OR-reduced (stream of common.cuh).
template <int LL, int G>
__device__ __forceinline__ uint64_t dropout_bits(uint64_t seed, uint32_t site, uint64_t base_idx, float p, int t) {
  constexpr int NB = (LL + 7) / 8 + 1;
  const uint64_t b0 = base_idx & ~(uint64_t)7;
  uint64_t bits = 0;
  for (int bb = t; bb < NB; bb += G) {
    const uint64_t i8 = b0 + 8ull * bb;
    if (i8 < base_idx + LL) {
      const uint64_t m = dropout_keep8(seed, site, i8, p);
      const long long sh = (long long)i8 - (long long)base_idx;
      bits |= sh >= 0 ? (m << sh) : (m >> (-sh));
    }
  }
  uint32_t lo = (uint32_t)bits, hi = (uint32_t)(bits >> 32);
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) {
    lo |= __shfl_xor_sync(0xffffffffu, lo, o);
    hi |= __shfl_xor_sync(0xffffffffu, hi, o);
  }
  bits = ((uint64_t)hi << 32) | lo;
  return LL >= 64 ? bits : (bits & ((1ull << (LL & 63)) - 1ull));
}

// From the reduced raw products: probabilities of both branches (in place: s2 -> P2, p1 out),
// and the norms.  g holds C C^T on i <= j.
template <int L>
__device__ __forceinline__ void dual_softmax(float (&s2)[L][L], const float (&g)[L][L], const float* __restrict__ mask_row,
                                             float inv_sqrt_dh, float (&p1)[L][L], float (&nrm)[L]) {
  float madd[L];
#pragma unroll
  for (int j = 0; j < L; ++j) {
    madd[j] = (1.f - mask_row[j]) * -10000.f;
    nrm[j] = sqrtf(g[j][j]);
  }
#pragma unroll
  for (int i = 0; i < L; ++i) {
    float m1 = -INFINITY, m2 = -INFINITY;
#pragma unroll
    for (int j = 0; j < L; ++j) {
      const float gij = i <= j ? g[i][j] : g[j][i];
      const float cs = gij / (nrm[i] * nrm[j]);
      p1[i][j] = 1.f - cs + (i == j ? 1.f : 0.f) + madd[j];
      s2[i][j] = s2[i][j] * inv_sqrt_dh + madd[j];
      m1 = fmaxf(m1, p1[i][j]);
      m2 = fmaxf(m2, s2[i][j]);
    }
    float z1 = 0.f, z2 = 0.f;
#pragma unroll
    for (int j = 0; j < L; ++j) {
      p1[i][j] = __expf(p1[i][j] - m1);
      s2[i][j] = __expf(s2[i][j] - m2);
      z1 += p1[i][j];
      z2 += s2[i][j];
    }
    const float r1 = 1.f / z1, r2 = 1.f / z2;
#pragma unroll
    for (int j = 0; j < L; ++j) {
      p1[i][j] *= r1;
      s2[i][j] *= r2;
    }
  }
}

template <int L, int DH, int G>
__global__ void __launch_bounds__(128) attn_small_fwd_kernel(const pmgt_attn_args a) {
  constexpr int DPL = DH / G, IPW = 32 / G;
  const int lane = threadIdx.x & 31;
  const int t = lane % G, slot = lane / G;
  const int H = a.H, heads = a.heads;
  const long long ld = 4ll * H;
  const long long n_items = a.rows * heads;
  const long long w0 = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nw = (long long)gridDim.x * (blockDim.x >> 5);
  const float inv_sqrt_dh = rsqrtf((float)DH);
  const float ks = a.dropout_p > 0.f ? 1.f / (1.f - a.dropout_p) : 1.f;
  for (long long it0 = w0 * IPW; it0 < n_items; it0 += nw * IPW) {
    long long item = it0 + slot;
    const bool valid = item < n_items;
    if (!valid) item = n_items - 1;  // lanes stay converged for the shuffles; stores are predicated
    const long long row = item / heads;
    const int head = (int)(item - row * heads);
    const uint16_t* src = a.qkvc + row * L * ld + head * DH;

    float s2[L][L], g[L][L];
    {
      float kf[L][DPL];
      load_rows<L, DPL, G>(src + H, ld, t, kf);
#pragma unroll
      for (int i = 0; i < L; ++i) {
        float qf[DPL];
        load_row<DPL, G>(src + (long long)i * ld, t, qf);
#pragma unroll
        for (int j = 0; j < L; ++j) {
          float acc = 0.f;
#pragma unroll
          for (int d = 0; d < DPL; ++d) acc = fmaf(qf[d], kf[j][d], acc);
          s2[i][j] = acc;
        }
      }
    }
    {
      float cf[L][DPL];
      load_rows<L, DPL, G>(src + 3 * H, ld, t, cf);
#pragma unroll
      for (int i = 0; i < L; ++i)
#pragma unroll
        for (int j = i; j < L; ++j) {
          float acc = 0.f;
#pragma unroll
          for (int d = 0; d < DPL; ++d) acc = fmaf(cf[i][d], cf[j][d], acc);
          g[i][j] = acc;
        }
    }
#pragma unroll
    for (int i = 0; i < L; ++i)
#pragma unroll
      for (int j = 0; j < L; ++j) {
        s2[i][j] = group_sum<G>(s2[i][j]);
        if (j >= i) g[i][j] = group_sum<G>(g[i][j]);
      }
    float p1[L][L], nrm[L];
    dual_softmax<L>(s2, g, a.mask + row * L, inv_sqrt_dh, p1, nrm);
    // A = beta * drop(P1) + (1 - beta) * drop(P2), kept in p1
    if (a.dropout_p > 0.f) {
      const uint64_t base = (uint64_t)item * (L * L);
      const uint64_t k1 = dropout_bits<L * L, G>(a.dropout_seed, a.dropout_site, base, a.dropout_p, t);
      const uint64_t k2 = dropout_bits<L * L, G>(a.dropout_seed, a.dropout_site + 1, base, a.dropout_p, t);
#pragma unroll
      for (int i = 0; i < L; ++i)
#pragma unroll
        for (int j = 0; j < L; ++j) {
          const int e = i * L + j;
          const float d1 = (k1 >> e) & 1ull ? p1[i][j] * ks : 0.f;
          const float d2 = (k2 >> e) & 1ull ? s2[i][j] * ks : 0.f;
          p1[i][j] = a.beta * d1 + (1.f - a.beta) * d2;
        }
    } else {
#pragma unroll
      for (int i = 0; i < L; ++i)
#pragma unroll
        for (int j = 0; j < L; ++j) p1[i][j] = a.beta * p1[i][j] + (1.f - a.beta) * s2[i][j];
    }
    float vf[L][DPL];
    load_rows<L, DPL, G>(src + 2 * H, ld, t, vf);
    uint16_t* dst = a.ctx + row * L * (long long)H + head * DH;
#pragma unroll
    for (int i = 0; i < L; ++i) {
      float o[DPL];
#pragma unroll
      for (int d = 0; d < DPL; ++d) o[d] = 0.f;
#pragma unroll
      for (int j = 0; j < L; ++j)
#pragma unroll
        for (int d = 0; d < DPL; ++d) o[d] = fmaf(p1[i][j], vf[j][d], o[d]);
      if (valid) store_row<DPL, G>(dst + (long long)i * H, t, o);
    }
  }
}

template <int L, int DH, int G, int MINB>
__global__ void __launch_bounds__(128, MINB) attn_small_bwd_kernel(const pmgt_attn_args a) {
  constexpr int DPL = DH / G, IPW = 32 / G;
  const int lane = threadIdx.x & 31;
  const int t = lane % G, slot = lane / G;
  const int H = a.H, heads = a.heads;
  const long long ld = 4ll * H;
  const long long n_items = a.rows * heads;
  const long long w0 = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nw = (long long)gridDim.x * (blockDim.x >> 5);
  const float inv_sqrt_dh = rsqrtf((float)DH);
  const float ks = a.dropout_p > 0.f ? 1.f / (1.f - a.dropout_p) : 1.f;
  for (long long it0 = w0 * IPW; it0 < n_items; it0 += nw * IPW) {
    long long item = it0 + slot;
    const bool valid = item < n_items;
    if (!valid) item = n_items - 1;
    const long long row = item / heads;
    const int head = (int)(item - row * heads);
    const uint16_t* src = a.qkvc + row * L * ld + head * DH;
    const uint16_t* dsrc = a.dctx + row * L * (long long)H + head * DH;
    uint16_t* dst = a.dqkvc + row * L * ld + head * DH;

    float s2[L][L], g[L][L], dA[L][L];
    {
      float kf[L][DPL];
      load_rows<L, DPL, G>(src + H, ld, t, kf);
#pragma unroll
      for (int i = 0; i < L; ++i) {
        float qf[DPL];
        load_row<DPL, G>(src + (long long)i * ld, t, qf);
#pragma unroll
        for (int j = 0; j < L; ++j) {
          float acc = 0.f;
#pragma unroll
          for (int d = 0; d < DPL; ++d) acc = fmaf(qf[d], kf[j][d], acc);
          s2[i][j] = acc;
        }
      }
    }
    {
      float cf[L][DPL];
      load_rows<L, DPL, G>(src + 3 * H, ld, t, cf);
#pragma unroll
      for (int i = 0; i < L; ++i)
#pragma unroll
        for (int j = i; j < L; ++j) {
          float acc = 0.f;
#pragma unroll
          for (int d = 0; d < DPL; ++d) acc = fmaf(cf[i][d], cf[j][d], acc);
          g[i][j] = acc;
        }
    }
    {
      float df[L][DPL], vf[L][DPL];
      load_rows<L, DPL, G>(dsrc, H, t, df);
      load_rows<L, DPL, G>(src + 2 * H, ld, t, vf);
#pragma unroll
      for (int i = 0; i < L; ++i)
#pragma unroll
        for (int j = 0; j < L; ++j) {
          float acc = 0.f;
#pragma unroll
          for (int d = 0; d < DPL; ++d) acc = fmaf(df[i][d], vf[j][d], acc);
          dA[i][j] = acc;
        }
    }
#pragma unroll
    for (int i = 0; i < L; ++i)
#pragma unroll
      for (int j = 0; j < L; ++j) {
        s2[i][j] = group_sum<G>(s2[i][j]);
        dA[i][j] = group_sum<G>(dA[i][j]);
        if (j >= i) g[i][j] = group_sum<G>(g[i][j]);
      }
    float p1[L][L], nrm[L];
    dual_softmax<L>(s2, g, a.mask + row * L, inv_sqrt_dh, p1, nrm);  // s2 = P2, p1 = P1
    uint64_t k1 = ~0ull, k2 = ~0ull;
    if (a.dropout_p > 0.f) {
      const uint64_t base = (uint64_t)item * (L * L);
      k1 = dropout_bits<L * L, G>(a.dropout_seed, a.dropout_site, base, a.dropout_p, t);
      k2 = dropout_bits<L * L, G>(a.dropout_seed, a.dropout_site + 1, base, a.dropout_p, t);
    }
    // dV_j = sum_i A_ij dctx_i   (dctx re-read: an L1 hit, cheaper than 6 rows of live registers)
    float df[L][DPL];
    load_rows<L, DPL, G>(dsrc, H, t, df);
#pragma unroll
    for (int j = 0; j < L; ++j) {
      float o[DPL];
#pragma unroll
      for (int d = 0; d < DPL; ++d) o[d] = 0.f;
#pragma unroll
      for (int i = 0; i < L; ++i) {
        const int e = i * L + j;
        const float d1 = (k1 >> e) & 1ull ? p1[i][j] * ks : 0.f;
        const float d2 = (k2 >> e) & 1ull ? s2[i][j] * ks : 0.f;
        const float aij = a.beta * d1 + (1.f - a.beta) * d2;
#pragma unroll
        for (int d = 0; d < DPL; ++d) o[d] = fmaf(aij, df[i][d], o[d]);
      }
      if (valid) store_row<DPL, G>(dst + 2 * H + (long long)j * ld, t, o);
    }
    // softmax backward of both branches: dS = P * (dP - sum_j dP P); dS1 -> p1, dS2 -> s2
#pragma unroll
    for (int i = 0; i < L; ++i) {
      float g1[L], g2[L], r1 = 0.f, r2 = 0.f;
#pragma unroll
      for (int j = 0; j < L; ++j) {
        const int e = i * L + j;
        g1[j] = (k1 >> e) & 1ull ? a.beta * dA[i][j] * ks : 0.f;
        g2[j] = (k2 >> e) & 1ull ? (1.f - a.beta) * dA[i][j] * ks : 0.f;
        r1 = fmaf(g1[j], p1[i][j], r1);
        r2 = fmaf(g2[j], s2[i][j], r2);
      }
#pragma unroll
      for (int j = 0; j < L; ++j) {
        p1[i][j] = p1[i][j] * (g1[j] - r1);
        s2[i][j] = s2[i][j] * (g2[j] - r2) * inv_sqrt_dh;
      }
    }
    // dQ_i = sum_j dS2_ij K_j ;  dK_j = sum_i dS2_ij Q_i   (1/sqrt(dh) folded into dS2)
    {
      float kf[L][DPL];
      load_rows<L, DPL, G>(src + H, ld, t, kf);
#pragma unroll
      for (int i = 0; i < L; ++i) {
        float o[DPL];
#pragma unroll
        for (int d = 0; d < DPL; ++d) o[d] = 0.f;
#pragma unroll
        for (int j = 0; j < L; ++j)
#pragma unroll
          for (int d = 0; d < DPL; ++d) o[d] = fmaf(s2[i][j], kf[j][d], o[d]);
        if (valid) store_row<DPL, G>(dst + (long long)i * ld, t, o);
      }
    }
    {
      float qf[L][DPL];
      load_rows<L, DPL, G>(src, ld, t, qf);
#pragma unroll
      for (int j = 0; j < L; ++j) {
        float o[DPL];
#pragma unroll
        for (int d = 0; d < DPL; ++d) o[d] = 0.f;
#pragma unroll
        for (int i = 0; i < L; ++i)
#pragma unroll
          for (int d = 0; d < DPL; ++d) o[d] = fmaf(s2[i][j], qf[i][d], o[d]);
        if (valid) store_row<DPL, G>(dst + H + (long long)j * ld, t, o);
      }
    }
    // cosine branch: D = -(dS1 + dS1^T);  dC_i = sum_j E_ij C_j with
    // E_ij = D_ij / (n_i n_j) - [i == j] * (sum_j' D_ij' cos_ij') / n_i^2
    {
      float E[L][L];
#pragma unroll
      for (int i = 0; i < L; ++i) {
        float sdc = 0.f;
#pragma unroll
        for (int j = 0; j < L; ++j) {
          const float Dij = -(p1[i][j] + p1[j][i]);
          const float inv = 1.f / (nrm[i] * nrm[j]);
          E[i][j] = Dij * inv;
          sdc = fmaf(Dij, (i <= j ? g[i][j] : g[j][i]) * inv, sdc);
        }
        E[i][i] -= sdc / (nrm[i] * nrm[i]);
      }
      float cf[L][DPL];
      load_rows<L, DPL, G>(src + 3 * H, ld, t, cf);
#pragma unroll
      for (int i = 0; i < L; ++i) {
        float o[DPL];
#pragma unroll
        for (int d = 0; d < DPL; ++d) o[d] = 0.f;
#pragma unroll
        for (int j = 0; j < L; ++j)
#pragma unroll
          for (int d = 0; d < DPL; ++d) o[d] = fmaf(E[i][j], cf[j][d], o[d]);
        if (valid) store_row<DPL, G>(dst + 3 * H + (long long)i * ld, t, o);
      }
    }
  }
}

template <int L, int DH, int G, bool BWD>
static int launch_small(const pmgt_attn_args* a, cudaStream_t st) {
  constexpr int IPW = 32 / G;
  const long long items = a->rows * a->heads;
  long long warps = (items + IPW - 1) / IPW;
  long long ctas = (warps + 3) / 4;
  static int minb = -1;  // tuning knob: resident CTAs per SM the backward kernel is compiled for (2 = 255 regs, 3 = 168)
  if (minb < 0) {
    const char* e = getenv("PMGT_ATTN_BWD_MINB");
    minb = (e && e[0] == '3') ? 3 : 2;
  }
  const long long cap = (long long)num_sms() * (BWD ? 2 * minb : 8);
  if (ctas > cap) ctas = cap;
  if (BWD) {
    if (minb == 3) attn_small_bwd_kernel<L, DH, G, 3><<<(unsigned)ctas, 128, 0, st>>>(*a);
    else attn_small_bwd_kernel<L, DH, G, 2><<<(unsigned)ctas, 128, 0, st>>>(*a);
  } else {
    attn_small_fwd_kernel<L, DH, G><<<(unsigned)ctas, 128, 0, st>>>(*a);
  }
  PMGT_LAUNCH_CHECK();
  return PMGT_OK;
}

// returns 1 if the shape was handled here, 0 if the caller must use the generic kernel, < 0 on error
template <bool BWD>
static int dispatch_small(const pmgt_attn_args* a, cudaStream_t st) {
  const int dh = a->H / a->heads;
  if ((a->H % 8) != 0 || (((uintptr_t)a->qkvc) & 15) != 0) return 0;
  int rc = 1;
#define PMGT_ATTN_CASE(L_, DH_, G_)                                    \
  if (a->L == L_ && dh == DH_) {                                       \
    const int r = launch_small<L_, DH_, G_, BWD>(a, st);               \
    return r ? r : rc;                                                 \
  }
  PMGT_ATTN_CASE(6, 128, 16)
  PMGT_ATTN_CASE(6, 64, 8)
  PMGT_ATTN_CASE(6, 32, 4)
#undef PMGT_ATTN_CASE
  return 0;
}

int attn_small_fwd(const pmgt_attn_args* a, cudaStream_t st) { return dispatch_small<false>(a, st); }
int attn_small_bwd(const pmgt_attn_args* a, cudaStream_t st) { return dispatch_small<true>(a, st); }

}  // namespace pmgt
