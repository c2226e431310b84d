// PMGTEmbeddings fusion (modeling_pmgt.py:199-208) specialised for hidden size 128, the default encoder.
//
// A token row is 128 bf16 = 256 B = sixteen 16-byte chunks, so HALF a warp owns one token: lane hl = lane & 15
// holds columns 8*hl .. 8*hl+7 in registers (one 16-byte load per operand, one Philox block per lane for the
// dropout mask), the two halves of a warp work on two tokens at once and every row reduction is four shuffles.
// The next token's operands are fetched before the current one is processed (register double buffer), because the
// kernel is a gather: its speed is set by the bytes in flight, not by arithmetic.
//
// Backward keeps ALL parameter-gradient accumulators in registers (d LayerNorm gamma/beta, d position row, the four
// rows of the modality-attention weight, the two projection-bias rows: 72 floats per lane), folds them across the
// CTA through shared memory once at the end and issues one global atomic per value and CTA.
//
// The generic row-per-warp kernels in rowwise.cu remain the path for every other hidden size; both produce the same
// dropout stream (element index tok * H + h), so forward and backward may mix them.
#include "common.cuh"

namespace pmgt {

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ float half_sum(float v) {  // reduction over the 16 lanes of a half-warp
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ void ld8(const float* __restrict__ p, float (&v)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  const float4 b = __ldg(reinterpret_cast<const float4*>(p + 4));
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

__device__ __forceinline__ void unpack8(const uint4& u, float (&v)[8]) {
  unpack_bf16x2(u.x, v[0], v[1]); unpack_bf16x2(u.y, v[2], v[3]);
  unpack_bf16x2(u.z, v[4], v[5]); unpack_bf16x2(u.w, v[6], v[7]);
}

__device__ __forceinline__ uint4 pack8(const float (&v)[8]) {
  uint4 u;
  u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]);
  u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
  return u;
}

__device__ __forceinline__ uint4 ldg16(const uint16_t* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }

// Forward math of one token shared by both passes.  In: projected rows ev/et (fp32), w = modality-attention
// weight rows [logit0: visual, textual | logit1: visual, textual], pr = position + role.  Out: tanh values,
// softmax weights a0/a1, z (pre-LayerNorm), mean, rstd.
__device__ __forceinline__ void fuse_row(const float (&ev)[8], const float (&et)[8], const float (&w)[4][8], float b0,
                                         float b1, const float (&pr)[8], float eps, float (&tv)[8], float (&tt)[8],
                                         float (&z)[8], float& a0, float& a1, float& mean, float& rstd) {
  float s0 = 0.f, s1 = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    tv[j] = tanhf(ev[j]);
    tt[j] = tanhf(et[j]);
    s0 += w[0][j] * tv[j] + w[1][j] * tt[j];
    s1 += w[2][j] * tv[j] + w[3][j] * tt[j];
  }
  s0 = half_sum(s0) + b0;
  s1 = half_sum(s1) + b1;
  const float mx = fmaxf(s0, s1);
  const float e0 = __expf(s0 - mx), e1 = __expf(s1 - mx);
  const float inv = 1.f / (e0 + e1);
  a0 = e0 * inv;
  a1 = e1 * inv;
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    z[j] = a0 * ev[j] + a1 * et[j] + pr[j];
    s += z[j];
  }
  mean = half_sum(s) * (1.f / 128.f);
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) { const float d = z[j] - mean; q = fmaf(d, d, q); }
  rstd = rsqrtf(half_sum(q) * (1.f / 128.f) + eps);
}

__global__ void __launch_bounds__(kThreads) embed_fwd128_kernel(const pmgt_embed_args a) {
  const int lane = threadIdx.x & 31, hl = lane & 15, sub = lane >> 4;
  const int h0 = hl * 8;
  float w[4][8], g[8], b[8];
#pragma unroll
  for (int k = 0; k < 4; ++k) ld8(a.w_att + k * 128 + h0, w[k]);
  ld8(a.ln_g + h0, g);
  ld8(a.ln_b + h0, b);
  const float b0 = __ldg(a.b_att), b1 = __ldg(a.b_att + 1);
  const float ks = a.dropout_p > 0.f ? 1.f / (1.f - a.dropout_p) : 1.f;
  const long long T = a.rows * a.L;
  const long long stride = (long long)gridDim.x * (kThreads / 32) * 2;
  long long base = ((long long)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5)) * 2;
  if (base >= T) return;
  // register double buffer: operands of the token after this one
  uint4 nv, nt;
  {
    const long long t0 = min(base + sub, T - 1);
    const long long er = a.row_idx ? __ldg(a.row_idx + t0) : t0;
    nv = ldg16(a.ev + er * 128 + h0);
    nt = ldg16(a.et + er * 128 + h0);
  }
  // position of this half-warp's token within its sequence, advanced incrementally (a 64-bit modulo per token costs
  // more instructions than the LayerNorm)
  const int l_step = (int)(stride % a.L);
  int l = (int)((base + sub) % a.L);
  for (; base < T; base += stride) {
    const long long tok = base + sub;
    const bool valid = tok < T;
    const uint4 cv = nv, ct = nt;
    if (base + stride < T) {
      const long long t1 = min(base + stride + sub, T - 1);
      const long long er = a.row_idx ? __ldg(a.row_idx + t1) : t1;
      nv = ldg16(a.ev + er * 128 + h0);
      nt = ldg16(a.et + er * 128 + h0);
    }
    const long long tc = valid ? tok : T - 1;
    float pr[8], rr[8];
    ld8(a.pos + (long long)l * 128 + h0, pr);
    ld8(a.role + (l > 0 ? 128 : 0) + h0, rr);
#pragma unroll
    for (int j = 0; j < 8; ++j) pr[j] += rr[j];
    float ev[8], et[8], tv[8], tt[8], z[8], a0, a1, mean, rstd;
    unpack8(cv, ev);
    unpack8(ct, et);
    fuse_row(ev, et, w, b0, b1, pr, a.ln_eps, tv, tt, z, a0, a1, mean, rstd);
    uint32_t keep = 0xffu;
    if (a.dropout_p > 0.f) keep = dropout_keep8(a.dropout_seed, a.dropout_site, (uint64_t)tc * 128u + (uint64_t)h0, a.dropout_p);
    float y[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float v = (z[j] - mean) * rstd * g[j] + b[j];
      y[j] = (keep >> j) & 1u ? v * ks : 0.f;
    }
    if (valid) *reinterpret_cast<uint4*>(a.x_out + tok * 128 + h0) = pack8(y);
    l += l_step;
    if (l >= a.L) l -= a.L;
  }
}

__device__ __forceinline__ void red_v4(float* addr, float x, float y, float z, float w) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}

// accumulator rows folded through shared memory: 0 d_ln_g, 1 d_ln_b, 2..5 d_w_att, 6 d_bias_v, 7 d_bias_t, then L rows of d_pos
template <bool TABLE>
__global__ void __launch_bounds__(kThreads) embed_bwd128_kernel(const pmgt_embed_args a) {
  extern __shared__ float red[];  // [(8 + L)][128] + 2
  const int lane = threadIdx.x & 31, hl = lane & 15, sub = lane >> 4;
  const int h0 = hl * 8;
  const int L = a.L;
  for (int i = threadIdx.x; i < (8 + L) * 128 + 2; i += kThreads) red[i] = 0.f;
  __syncthreads();
  float w[4][8], g[8];
#pragma unroll
  for (int k = 0; k < 4; ++k) ld8(a.w_att + k * 128 + h0, w[k]);
  ld8(a.ln_g + h0, g);
  const float b0 = __ldg(a.b_att), b1 = __ldg(a.b_att + 1);
  const float ks = a.dropout_p > 0.f ? 1.f / (1.f - a.dropout_p) : 1.f;
  float acc[8][8];
#pragma unroll
  for (int k = 0; k < 8; ++k)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[k][j] = 0.f;
  float db0 = 0.f, db1 = 0.f;
  const long long stride = (long long)gridDim.x * (kThreads / 32) * 2;
  const long long first = ((long long)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5)) * 2;
  // position-major order: all rows of position l, then l + 1, so the position gradient accumulates in registers
  for (int l = 0; l < L; ++l) {
    float pr[8], rr[8], dpos[8];
    ld8(a.pos + (long long)l * 128 + h0, pr);
    ld8(a.role + (l > 0 ? 128 : 0) + h0, rr);
#pragma unroll
    for (int j = 0; j < 8; ++j) { pr[j] += rr[j]; dpos[j] = 0.f; }
    long long base = first;
    uint4 nv = make_uint4(0, 0, 0, 0), nt = nv, nd = nv, nb = nv;
    long long ner = 0;
    if (base < a.rows) {
      const long long t0 = min(base + sub, (long long)a.rows - 1) * L + l;
      ner = a.row_idx ? __ldg(a.row_idx + t0) : t0;
      nv = ldg16(a.ev + ner * 128 + h0);
      nt = ldg16(a.et + ner * 128 + h0);
      nd = ldg16(a.dx + t0 * 128 + h0);
      if (a.dx_b) nb = ldg16(a.dx_b + t0 * 128 + h0);
    }
    for (; base < a.rows; base += stride) {
      const long long row = base + sub;
      const bool valid = row < a.rows;
      const long long tok = (valid ? row : a.rows - 1) * L + l;
      const uint4 cv = nv, ct = nt, cd = nd, cb = nb;
      const long long er = ner;
      if (base + stride < a.rows) {
        const long long t1 = min(base + stride + sub, (long long)a.rows - 1) * L + l;
        ner = a.row_idx ? __ldg(a.row_idx + t1) : t1;
        nv = ldg16(a.ev + ner * 128 + h0);
        nt = ldg16(a.et + ner * 128 + h0);
        nd = ldg16(a.dx + t1 * 128 + h0);
        if (a.dx_b) nb = ldg16(a.dx_b + t1 * 128 + h0);
      }
      float ev[8], et[8], tv[8], tt[8], z[8], a0, a1, mean, rstd;
      unpack8(cv, ev);
      unpack8(ct, et);
      fuse_row(ev, et, w, b0, b1, pr, a.ln_eps, tv, tt, z, a0, a1, mean, rstd);
      float d[8];
      unpack8(cd, d);
      if (a.dx_b) {
        float e[8];
        unpack8(cb, e);
#pragma unroll
        for (int j = 0; j < 8; ++j) d[j] += e[j];
      }
      uint32_t keep = valid ? 0xffu : 0u;  // an out-of-range half contributes exact zeros everywhere below
      if (a.dropout_p > 0.f) keep &= dropout_keep8(a.dropout_seed, a.dropout_site, (uint64_t)tok * 128u + (uint64_t)h0, a.dropout_p);
      // LayerNorm backward
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float dj = (keep >> j) & 1u ? d[j] * ks : 0.f;
        const float xh = (z[j] - mean) * rstd;
        acc[0][j] = fmaf(dj, xh, acc[0][j]);
        acc[1][j] += dj;
        const float dg = dj * g[j];
        s1 += dg;
        s2 = fmaf(dg, xh, s2);
        d[j] = dg;   // d now holds dy * gamma
        z[j] = xh;   // z now holds xhat
      }
      s1 = half_sum(s1) * (1.f / 128.f);
      s2 = half_sum(s2) * (1.f / 128.f);
      float da0 = 0.f, da1 = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float dz = rstd * (d[j] - s1 - z[j] * s2);
        dpos[j] += dz;
        da0 = fmaf(dz, ev[j], da0);
        da1 = fmaf(dz, et[j], da1);
        d[j] = dz;   // d now holds dz
      }
      da0 = half_sum(da0);
      da1 = half_sum(da1);
      // softmax over the two modality logits
      const float dot = da0 * a0 + da1 * a1;
      const float ds0 = a0 * (da0 - dot), ds1 = a1 * (da1 - dot);
      db0 += ds0;
      db1 += ds1;
      float dv[8], dt[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        acc[2][j] = fmaf(ds0, tv[j], acc[2][j]);
        acc[3][j] = fmaf(ds0, tt[j], acc[3][j]);
        acc[4][j] = fmaf(ds1, tv[j], acc[4][j]);
        acc[5][j] = fmaf(ds1, tt[j], acc[5][j]);
        dv[j] = a0 * d[j] + (ds0 * w[0][j] + ds1 * w[2][j]) * (1.f - tv[j] * tv[j]);
        dt[j] = a1 * d[j] + (ds0 * w[1][j] + ds1 * w[3][j]) * (1.f - tt[j] * tt[j]);
        // bias gradient of the projections = column sums of the bf16-rounded dev / det
        acc[6][j] += bf16_bits_to_float(float_to_bf16_bits(dv[j]));
        acc[7][j] += bf16_bits_to_float(float_to_bf16_bits(dt[j]));
      }
      if (valid) {
        if (TABLE) {
          if (!(a.skip_row0 && er == 0)) {
            float* pv = a.dev_acc + er * 128 + h0;
            float* pt = a.det_acc + er * 128 + h0;
            red_v4(pv, dv[0], dv[1], dv[2], dv[3]);
            red_v4(pv + 4, dv[4], dv[5], dv[6], dv[7]);
            red_v4(pt, dt[0], dt[1], dt[2], dt[3]);
            red_v4(pt + 4, dt[4], dt[5], dt[6], dt[7]);
          }
        } else {
          *reinterpret_cast<uint4*>(a.dev + tok * 128 + h0) = pack8(dv);
          *reinterpret_cast<uint4*>(a.det + tok * 128 + h0) = pack8(dt);
        }
      }
    }
    // position gradient of this l: fold the two halves, then into the CTA's shared row
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float v = dpos[j] + __shfl_xor_sync(0xffffffffu, dpos[j], 16);
      if (sub == 0 && v != 0.f) atomicAdd(&red[(8 + l) * 128 + h0 + j], v);
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float v = acc[k][j] + __shfl_xor_sync(0xffffffffu, acc[k][j], 16);
      if (sub == 0 && v != 0.f) atomicAdd(&red[k * 128 + h0 + j], v);
    }
  if (hl == 0) {  // every lane of a half-warp holds the same per-token scalars
    if (db0 != 0.f) atomicAdd(&red[(8 + L) * 128 + 0], db0);
    if (db1 != 0.f) atomicAdd(&red[(8 + L) * 128 + 1], db1);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 8 * 128; i += kThreads) {
    const float v = red[i];
    if (v == 0.f) continue;
    const int k = i >> 7, h = i & 127;
    float* dst = k == 0 ? a.d_ln_g : k == 1 ? a.d_ln_b : k < 6 ? a.d_w_att + (k - 2) * 128 : k == 6 ? a.d_bias_v : a.d_bias_t;
    atomicAdd(dst + h, v);
  }
  for (int i = threadIdx.x; i < L * 128; i += kThreads) {
    const float v = red[8 * 128 + i];
    if (v == 0.f) continue;
    const int l = i >> 7, h = i & 127;
    atomicAdd(a.d_pos + (long long)l * 128 + h, v);
    atomicAdd(a.d_role + (l > 0 ? 128 : 0) + h, v);
  }
  if (threadIdx.x < 2) {
    const float v = red[(8 + L) * 128 + threadIdx.x];
    if (v != 0.f) atomicAdd(a.d_b_att + threadIdx.x, v);
  }
}

int grid_for(long long units_of_two, int ctas_per_sm) {
  long long need = (units_of_two + (kThreads / 32) * 2 - 1) / ((kThreads / 32) * 2);
  const long long cap = (long long)num_sms() * ctas_per_sm;
  if (need > cap) need = cap;
  if (need < 1) need = 1;
  return (int)need;
}

}  // namespace

// called from pmgt_embed_fuse_fwd / pmgt_embed_fuse_bwd (rowwise.cu) after argument validation, when H == 128
int embed128_supported(const pmgt_embed_args* a) {
  if (a->H != 128 || a->L > 256) return 0;
  const uintptr_t p = (uintptr_t)a->ev | (uintptr_t)a->et | (uintptr_t)a->x_out | (uintptr_t)a->dx | (uintptr_t)a->dx_b |
                      (uintptr_t)a->dev | (uintptr_t)a->det | (uintptr_t)a->w_att | (uintptr_t)a->pos | (uintptr_t)a->role |
                      (uintptr_t)a->ln_g | (uintptr_t)a->ln_b | (uintptr_t)a->dev_acc | (uintptr_t)a->det_acc;
  return (p & 15) == 0;
}

int embed128_fwd(const pmgt_embed_args* a, cudaStream_t st) {
  embed_fwd128_kernel<<<grid_for(a->rows * a->L, 2), kThreads, 0, st>>>(*a);
  PMGT_LAUNCH_CHECK();
  return PMGT_OK;
}

int embed128_bwd(const pmgt_embed_args* a, cudaStream_t st) {
  const size_t smem = ((size_t)(8 + a->L) * 128 + 2) * sizeof(float);
  static unsigned long long cfg = 0;
  if (first_use_on_device(cfg)) {
    PMGT_CHECK_CUDA(cudaFuncSetAttribute(embed_bwd128_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    PMGT_CHECK_CUDA(cudaFuncSetAttribute(embed_bwd128_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  }
  const int grid = grid_for(a->rows, 1);
  if (a->row_idx)
    embed_bwd128_kernel<true><<<grid, kThreads, smem, st>>>(*a);
  else
    embed_bwd128_kernel<false><<<grid, kThreads, smem, st>>>(*a);
  PMGT_LAUNCH_CHECK();
  return PMGT_OK;
}

}  // namespace pmgt
