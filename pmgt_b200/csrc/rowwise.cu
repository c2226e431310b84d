// Row-wise (one warp per token) kernels of the PMGT encoder:
//   * PMGTEmbeddings fusion + LayerNorm (modeling_pmgt.py:199-208) fwd/bwd
//   * BertSelfOutput / BertOutput "dropout + residual + LayerNorm" fwd/bwd
//   * column sums (bias gradients), fp32->bf16 cast, row gather, sum of squares,
//     fused AdamW (pmgt/optimizers.py:256-270)
// All are HBM-bandwidth-bound streaming kernels: 8-byte coalesced bf16 loads,
// fp32 math in registers, persistent grids of (SM count x resident CTAs).
#include "common.cuh"

namespace pmgt {

constexpr int kRowThreads = 256;  // 8 warps per CTA

// A token row of H values distributed over a warp: lane owns groups of 4
// consecutive values, group index g = lane + 32*i, i < G.
template <int G>
struct RowRegs {
  float v[G][4];
};

template <int G>
__device__ __forceinline__ void load_row_bf16(const uint16_t* __restrict__ p, int H, int lane, RowRegs<G>& r) {
#pragma unroll
  for (int i = 0; i < G; ++i) {
    const int h = (lane + 32 * i) * 4;
    if (h < H) {
      const uint2 u = *reinterpret_cast<const uint2*>(p + h);
      unpack_bf16x2(u.x, r.v[i][0], r.v[i][1]);
      unpack_bf16x2(u.y, r.v[i][2], r.v[i][3]);
    } else {
      r.v[i][0] = r.v[i][1] = r.v[i][2] = r.v[i][3] = 0.f;
    }
  }
}
template <int G>
__device__ __forceinline__ void load_row_f32(const float* __restrict__ p, int H, int lane, RowRegs<G>& r) {
#pragma unroll
  for (int i = 0; i < G; ++i) {
    const int h = (lane + 32 * i) * 4;
    if (h < H) {
      const float4 u = *reinterpret_cast<const float4*>(p + h);
      r.v[i][0] = u.x; r.v[i][1] = u.y; r.v[i][2] = u.z; r.v[i][3] = u.w;
    } else {
      r.v[i][0] = r.v[i][1] = r.v[i][2] = r.v[i][3] = 0.f;
    }
  }
}
template <int G>
__device__ __forceinline__ void store_row_bf16(uint16_t* __restrict__ p, int H, int lane, const RowRegs<G>& r) {
#pragma unroll
  for (int i = 0; i < G; ++i) {
    const int h = (lane + 32 * i) * 4;
    if (h < H) {
      uint2 u;
      u.x = pack_bf16x2(r.v[i][0], r.v[i][1]);
      u.y = pack_bf16x2(r.v[i][2], r.v[i][3]);
      *reinterpret_cast<uint2*>(p + h) = u;
    }
  }
}
template <int G>
__device__ __forceinline__ void store_row_f32(float* __restrict__ p, int H, int lane, const RowRegs<G>& r) {
#pragma unroll
  for (int i = 0; i < G; ++i) {
    const int h = (lane + 32 * i) * 4;
    if (h < H) *reinterpret_cast<float4*>(p + h) = make_float4(r.v[i][0], r.v[i][1], r.v[i][2], r.v[i][3]);
  }
}

// mean / rstd of a row (two-pass, values already in registers)
template <int G>
__device__ __forceinline__ void row_stats(const RowRegs<G>& z, int H, int lane, float eps, float& mean, float& rstd) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < G; ++i)
    if ((lane + 32 * i) * 4 < H) s += z.v[i][0] + z.v[i][1] + z.v[i][2] + z.v[i][3];
  mean = warp_sum(s) / (float)H;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < G; ++i)
    if ((lane + 32 * i) * 4 < H) {
#pragma unroll
      for (int j = 0; j < 4; ++j) { const float d = z.v[i][j] - mean; q += d * d; }
    }
  rstd = rsqrtf(warp_sum(q) / (float)H + eps);
}

// per-warp shared accumulators: acc[k][h] += val, flushed with global atomics
__device__ __forceinline__ void flush_acc(float* __restrict__ dst, float* acc, int H, int lane) {
  if (!dst) return;
  for (int h = lane; h < H; h += 32) {
    const float v = acc[h];
    if (v != 0.f) atomicAdd(dst + h, v);
    acc[h] = 0.f;
  }
}

// ---------------------------------------------------------------------------
// embeddings: modality attention fusion + position/role + LayerNorm + dropout
// ---------------------------------------------------------------------------
template <int G>
__device__ __forceinline__ void embed_forward_row(const pmgt_embed_args& a, long long tok, int l, int lane,
                                                  RowRegs<G>& ev, RowRegs<G>& et, RowRegs<G>& tv, RowRegs<G>& tt,
                                                  RowRegs<G>& z, float& a0, float& a1, float& mean, float& rstd) {
  const int H = a.H;
  const long long er = a.row_idx ? a.row_idx[tok] : tok;
  load_row_bf16<G>(a.ev + er * H, H, lane, ev);
  load_row_bf16<G>(a.et + er * H, H, lane, et);
  float s0 = 0.f, s1 = 0.f;
#pragma unroll
  for (int i = 0; i < G; ++i) {
    const int h = (lane + 32 * i) * 4;
    if (h < H) {
      const float4 w0v = *reinterpret_cast<const float4*>(a.w_att + h);
      const float4 w0t = *reinterpret_cast<const float4*>(a.w_att + H + h);
      const float4 w1v = *reinterpret_cast<const float4*>(a.w_att + 2 * H + h);
      const float4 w1t = *reinterpret_cast<const float4*>(a.w_att + 3 * H + h);
      const float w0va[4] = {w0v.x, w0v.y, w0v.z, w0v.w}, w0ta[4] = {w0t.x, w0t.y, w0t.z, w0t.w};
      const float w1va[4] = {w1v.x, w1v.y, w1v.z, w1v.w}, w1ta[4] = {w1t.x, w1t.y, w1t.z, w1t.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        tv.v[i][j] = tanhf(ev.v[i][j]);
        tt.v[i][j] = tanhf(et.v[i][j]);
        s0 += w0va[j] * tv.v[i][j] + w0ta[j] * tt.v[i][j];
        s1 += w1va[j] * tv.v[i][j] + w1ta[j] * tt.v[i][j];
      }
    }
  }
  s0 = warp_sum(s0) + a.b_att[0];
  s1 = warp_sum(s1) + a.b_att[1];
  const float mx = fmaxf(s0, s1);
  const float e0 = __expf(s0 - mx), e1 = __expf(s1 - mx);
  a0 = e0 / (e0 + e1);
  a1 = e1 / (e0 + e1);
  const float* pos = a.pos + (long long)l * H;
  const float* role = a.role + (l > 0 ? H : 0);
#pragma unroll
  for (int i = 0; i < G; ++i) {
    const int h = (lane + 32 * i) * 4;
    if (h < H) {
      const float4 pp = *reinterpret_cast<const float4*>(pos + h);
      const float4 rr = *reinterpret_cast<const float4*>(role + h);
      z.v[i][0] = a0 * ev.v[i][0] + a1 * et.v[i][0] + pp.x + rr.x;
      z.v[i][1] = a0 * ev.v[i][1] + a1 * et.v[i][1] + pp.y + rr.y;
      z.v[i][2] = a0 * ev.v[i][2] + a1 * et.v[i][2] + pp.z + rr.z;
      z.v[i][3] = a0 * ev.v[i][3] + a1 * et.v[i][3] + pp.w + rr.w;
    } else {
      z.v[i][0] = z.v[i][1] = z.v[i][2] = z.v[i][3] = 0.f;
    }
  }
  row_stats<G>(z, H, lane, a.ln_eps, mean, rstd);
}

template <int G>
__global__ void __launch_bounds__(kRowThreads) embed_fuse_fwd_kernel(const pmgt_embed_args a) {
  const int lane = threadIdx.x & 31;
  const long long T = a.rows * a.L;
  const long long warp0 = (long long)blockIdx.x * (kRowThreads / 32) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (kRowThreads / 32);
  const int H = a.H;
  const float keep_scale = a.dropout_p > 0.f ? 1.f / (1.f - a.dropout_p) : 1.f;
  for (long long tok = warp0; tok < T; tok += nwarps) {
    const int l = (int)(tok % a.L);
    RowRegs<G> ev, et, tv, tt, z;
    float a0, a1, mean, rstd;
    embed_forward_row<G>(a, tok, l, lane, ev, et, tv, tt, z, a0, a1, mean, rstd);
#pragma unroll
    for (int i = 0; i < G; ++i) {
      const int h = (lane + 32 * i) * 4;
      if (h < H) {
        const float4 g = *reinterpret_cast<const float4*>(a.ln_g + h);
        const float4 b = *reinterpret_cast<const float4*>(a.ln_b + h);
        const float ga[4] = {g.x, g.y, g.z, g.w}, ba[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float y = (z.v[i][j] - mean) * rstd * ga[j] + ba[j];
          if (a.dropout_p > 0.f)
            y = dropout_keep(a.dropout_seed, a.dropout_site, (uint64_t)tok * H + h + j, a.dropout_p) ? y * keep_scale : 0.f;
          z.v[i][j] = y;
        }
      }
    }
    store_row_bf16<G>(a.x_out + tok * H, H, lane, z);
  }
}

// Backward.  Work is ordered position-major (all rows of position l, then l+1)
// so the position/role gradient of one l accumulates locally before a flush.
// per-warp smem accumulators: [0]=d_ln_g [1]=d_ln_b [2]=d_pos(l) [3..6]=d_w_att (4 x H) [7]=dbias_v [8]=dbias_t
template <int G>
__global__ void __launch_bounds__(kRowThreads) embed_fuse_bwd_kernel(const pmgt_embed_args a) {
  extern __shared__ float acc_all[];
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const int H = a.H;
  float* acc = acc_all + (size_t)wib * 9 * H;
  for (int i = lane; i < 9 * H; i += 32) acc[i] = 0.f;
  float db_att0 = 0.f, db_att1 = 0.f;
  __syncwarp();
  const long long warp0 = (long long)blockIdx.x * (kRowThreads / 32) + wib;
  const long long nwarps = (long long)gridDim.x * (kRowThreads / 32);
  const float keep_scale = a.dropout_p > 0.f ? 1.f / (1.f - a.dropout_p) : 1.f;
  for (int l = 0; l < a.L; ++l) {
    for (long long row = warp0; row < a.rows; row += nwarps) {
      const long long tok = row * a.L + l;
      RowRegs<G> ev, et, tv, tt, z, dy;
      float a0, a1, mean, rstd;
      embed_forward_row<G>(a, tok, l, lane, ev, et, tv, tt, z, a0, a1, mean, rstd);
      load_row_bf16<G>(a.dx + tok * H, H, lane, dy);
      if (a.dx_b) {
        RowRegs<G> e;
        load_row_bf16<G>(a.dx_b + tok * H, H, lane, e);
#pragma unroll
        for (int i = 0; i < G; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) dy.v[i][j] += e.v[i][j];
      }
      // LayerNorm backward
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int i = 0; i < G; ++i) {
        const int h = (lane + 32 * i) * 4;
        if (h < H) {
          const float4 g = *reinterpret_cast<const float4*>(a.ln_g + h);
          const float ga[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float d = dy.v[i][j];
            if (a.dropout_p > 0.f)
              d = dropout_keep(a.dropout_seed, a.dropout_site, (uint64_t)tok * H + h + j, a.dropout_p) ? d * keep_scale : 0.f;
            const float xh = (z.v[i][j] - mean) * rstd;
            acc[0 * H + h + j] += d * xh;
            acc[1 * H + h + j] += d;
            const float dg = d * ga[j];
            s1 += dg;
            s2 += dg * xh;
            dy.v[i][j] = dg;   // reuse: dy now holds dy*gamma
            z.v[i][j] = xh;    // reuse: z now holds xhat
          }
        }
      }
      s1 = warp_sum(s1) / (float)H;
      s2 = warp_sum(s2) / (float)H;
      // dz -> d_pos / d_role ; d a0/a1 ; d ev / d et through the weighted sum
      float da0 = 0.f, da1 = 0.f;
#pragma unroll
      for (int i = 0; i < G; ++i) {
        const int h = (lane + 32 * i) * 4;
        if (h < H) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float dz = rstd * (dy.v[i][j] - s1 - z.v[i][j] * s2);
            acc[2 * H + h + j] += dz;
            da0 += dz * ev.v[i][j];
            da1 += dz * et.v[i][j];
            dy.v[i][j] = dz;  // reuse: dy now holds dz
          }
        }
      }
      da0 = warp_sum(da0);
      da1 = warp_sum(da1);
      // softmax over the two modality logits
      const float dot = da0 * a0 + da1 * a1;
      const float ds0 = a0 * (da0 - dot), ds1 = a1 * (da1 - dot);
      db_att0 += ds0;
      db_att1 += ds1;
      RowRegs<G> dev, det;
#pragma unroll
      for (int i = 0; i < G; ++i) {
        const int h = (lane + 32 * i) * 4;
        if (h < H) {
          const float4 w0v = *reinterpret_cast<const float4*>(a.w_att + h);
          const float4 w0t = *reinterpret_cast<const float4*>(a.w_att + H + h);
          const float4 w1v = *reinterpret_cast<const float4*>(a.w_att + 2 * H + h);
          const float4 w1t = *reinterpret_cast<const float4*>(a.w_att + 3 * H + h);
          const float w0va[4] = {w0v.x, w0v.y, w0v.z, w0v.w}, w0ta[4] = {w0t.x, w0t.y, w0t.z, w0t.w};
          const float w1va[4] = {w1v.x, w1v.y, w1v.z, w1v.w}, w1ta[4] = {w1t.x, w1t.y, w1t.z, w1t.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float thv = tv.v[i][j], tht = tt.v[i][j];
            acc[3 * H + h + j] += ds0 * thv;
            acc[4 * H + h + j] += ds0 * tht;
            acc[5 * H + h + j] += ds1 * thv;
            acc[6 * H + h + j] += ds1 * tht;
            const float dv = a0 * dy.v[i][j] + (ds0 * w0va[j] + ds1 * w1va[j]) * (1.f - thv * thv);
            const float dt = a1 * dy.v[i][j] + (ds0 * w0ta[j] + ds1 * w1ta[j]) * (1.f - tht * tht);
            dev.v[i][j] = dv;
            det.v[i][j] = dt;
            // bias gradient of the projections = column sums of the bf16-rounded dev/det
            acc[7 * H + h + j] += bf16_bits_to_float(float_to_bf16_bits(dv));
            acc[8 * H + h + j] += bf16_bits_to_float(float_to_bf16_bits(dt));
          }
        } else {
          dev.v[i][0] = dev.v[i][1] = dev.v[i][2] = dev.v[i][3] = 0.f;
          det.v[i][0] = det.v[i][1] = det.v[i][2] = det.v[i][3] = 0.f;
        }
      }
      if (a.row_idx) {
        const long long er = a.row_idx[tok];
        if (!(a.skip_row0 && er == 0)) {
#pragma unroll
          for (int i = 0; i < G; ++i) {
            const int h = (lane + 32 * i) * 4;
            if (h < H) {
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(a.dev_acc + er * H + h), "f"(dev.v[i][0]),
                           "f"(dev.v[i][1]), "f"(dev.v[i][2]), "f"(dev.v[i][3]) : "memory");
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(a.det_acc + er * H + h), "f"(det.v[i][0]),
                           "f"(det.v[i][1]), "f"(det.v[i][2]), "f"(det.v[i][3]) : "memory");
            }
          }
        }
      } else {
        store_row_bf16<G>(a.dev + tok * H, H, lane, dev);
        store_row_bf16<G>(a.det + tok * H, H, lane, det);
      }
    }
    // flush position / role gradient of this l
    __syncwarp();
    for (int h = lane; h < H; h += 32) {
      const float v = acc[2 * H + h];
      if (v != 0.f) {
        atomicAdd(a.d_pos + (long long)l * H + h, v);
        atomicAdd(a.d_role + (l > 0 ? H : 0) + h, v);
      }
      acc[2 * H + h] = 0.f;
    }
    __syncwarp();
  }
  __syncwarp();
  flush_acc(a.d_ln_g, acc + 0 * H, H, lane);
  flush_acc(a.d_ln_b, acc + 1 * H, H, lane);
  flush_acc(a.d_w_att + 0 * H, acc + 3 * H, H, lane);
  flush_acc(a.d_w_att + 1 * H, acc + 4 * H, H, lane);
  flush_acc(a.d_w_att + 2 * H, acc + 5 * H, H, lane);
  flush_acc(a.d_w_att + 3 * H, acc + 6 * H, H, lane);
  flush_acc(a.d_bias_v, acc + 7 * H, H, lane);
  flush_acc(a.d_bias_t, acc + 8 * H, H, lane);
  if (lane == 0) {
    if (db_att0 != 0.f) atomicAdd(a.d_b_att + 0, db_att0);
    if (db_att1 != 0.f) atomicAdd(a.d_b_att + 1, db_att1);
  }
}

// ---------------------------------------------------------------------------
// y = LayerNorm(dropout(o) + res)
// ---------------------------------------------------------------------------
template <int G>
__global__ void __launch_bounds__(kRowThreads) res_ln_fwd_kernel(const pmgt_resln_args a) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * (kRowThreads / 32) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (kRowThreads / 32);
  const int H = a.H;
  const float keep_scale = a.dropout_p > 0.f ? 1.f / (1.f - a.dropout_p) : 1.f;
  for (long long tok = warp0; tok < a.T; tok += nwarps) {
    RowRegs<G> o, r;
    load_row_bf16<G>(a.o + tok * H, H, lane, o);
    load_row_bf16<G>(a.res + tok * H, H, lane, r);
#pragma unroll
    for (int i = 0; i < G; ++i) {
      const int h = (lane + 32 * i) * 4;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float v = o.v[i][j];
        if (a.dropout_p > 0.f && h < H)
          v = dropout_keep(a.dropout_seed, a.dropout_site, (uint64_t)tok * H + h + j, a.dropout_p) ? v * keep_scale : 0.f;
        o.v[i][j] = v + r.v[i][j];
      }
    }
    float mean, rstd;
    row_stats<G>(o, H, lane, a.ln_eps, mean, rstd);
#pragma unroll
    for (int i = 0; i < G; ++i) {
      const int h = (lane + 32 * i) * 4;
      if (h < H) {
        const float4 g = *reinterpret_cast<const float4*>(a.ln_g + h);
        const float4 b = *reinterpret_cast<const float4*>(a.ln_b + h);
        o.v[i][0] = (o.v[i][0] - mean) * rstd * g.x + b.x;
        o.v[i][1] = (o.v[i][1] - mean) * rstd * g.y + b.y;
        o.v[i][2] = (o.v[i][2] - mean) * rstd * g.z + b.z;
        o.v[i][3] = (o.v[i][3] - mean) * rstd * g.w + b.w;
      }
    }
    store_row_bf16<G>(a.y + tok * H, H, lane, o);
    if (a.y_f32) store_row_f32<G>(a.y_f32 + tok * H, H, lane, o);
  }
}

// per-warp smem accumulators: [0]=d_g [1]=d_b [2]=d_bias
template <int G>
__global__ void __launch_bounds__(kRowThreads) res_ln_bwd_kernel(const pmgt_resln_args a) {
  extern __shared__ float acc_all[];
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const int H = a.H;
  float* acc = acc_all + (size_t)wib * 3 * H;
  for (int i = lane; i < 3 * H; i += 32) acc[i] = 0.f;
  __syncwarp();
  const long long warp0 = (long long)blockIdx.x * (kRowThreads / 32) + wib;
  const long long nwarps = (long long)gridDim.x * (kRowThreads / 32);
  const float keep_scale = a.dropout_p > 0.f ? 1.f / (1.f - a.dropout_p) : 1.f;
  const bool sep_do = a.d_o != nullptr && a.d_o != a.dz;
  for (long long tok = warp0; tok < a.T; tok += nwarps) {
    RowRegs<G> z, r, dy;
    load_row_bf16<G>(a.o + tok * H, H, lane, z);
    load_row_bf16<G>(a.res + tok * H, H, lane, r);
    if (a.dy) load_row_bf16<G>(a.dy + tok * H, H, lane, dy);
    else {
#pragma unroll
      for (int i = 0; i < G; ++i) dy.v[i][0] = dy.v[i][1] = dy.v[i][2] = dy.v[i][3] = 0.f;
    }
    if (a.dy_f32) {
      RowRegs<G> e;
      load_row_f32<G>(a.dy_f32 + tok * H, H, lane, e);
#pragma unroll
      for (int i = 0; i < G; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dy.v[i][j] += e.v[i][j];
    }
    uint32_t keep_bits = 0xffffffffu;  // bit (i*4+j), G*4 <= 32
#pragma unroll
    for (int i = 0; i < G; ++i) {
      const int h = (lane + 32 * i) * 4;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float v = z.v[i][j];
        if (a.dropout_p > 0.f && h < H) {
          const bool k = dropout_keep(a.dropout_seed, a.dropout_site, (uint64_t)tok * H + h + j, a.dropout_p);
          if (!k) keep_bits &= ~(1u << (i * 4 + j));
          v = k ? v * keep_scale : 0.f;
        }
        z.v[i][j] = v + r.v[i][j];
      }
    }
    float mean, rstd;
    row_stats<G>(z, H, lane, a.ln_eps, mean, rstd);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < G; ++i) {
      const int h = (lane + 32 * i) * 4;
      if (h < H) {
        const float4 g = *reinterpret_cast<const float4*>(a.ln_g + h);
        const float ga[4] = {g.x, g.y, g.z, g.w};
        // the lane owns these four columns for every token: 16-byte read-modify-write of its accumulators
        // (conflict-free; the scalar form was a 4-way bank conflict and four times the instructions)
        float4* pg = reinterpret_cast<float4*>(acc + 0 * H + h);
        float4* pb = reinterpret_cast<float4*>(acc + 1 * H + h);
        float4 ag = *pg, ab = *pb;
        float* agf = reinterpret_cast<float*>(&ag);
        float* abf = reinterpret_cast<float*>(&ab);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float d = dy.v[i][j];
          const float xh = (z.v[i][j] - mean) * rstd;
          agf[j] += d * xh;
          abf[j] += d;
          const float dg = d * ga[j];
          s1 += dg;
          s2 += dg * xh;
          dy.v[i][j] = dg;
          z.v[i][j] = xh;
        }
        *pg = ag;
        *pb = ab;
      }
    }
    s1 = warp_sum(s1) / (float)H;
    s2 = warp_sum(s2) / (float)H;
    RowRegs<G> dout;
#pragma unroll
    for (int i = 0; i < G; ++i) {
      const int h = (lane + 32 * i) * 4;
      float4 ad = make_float4(0.f, 0.f, 0.f, 0.f);
      float* adf = reinterpret_cast<float*>(&ad);
      if (h < H) ad = *reinterpret_cast<const float4*>(acc + 2 * H + h);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float dz = (h < H) ? rstd * (dy.v[i][j] - s1 - z.v[i][j] * s2) : 0.f;
        dy.v[i][j] = dz;
        const float dov = (keep_bits >> (i * 4 + j)) & 1u ? dz * keep_scale : 0.f;
        dout.v[i][j] = dov;
        adf[j] += bf16_bits_to_float(float_to_bf16_bits(dov));
      }
      if (h < H) *reinterpret_cast<float4*>(acc + 2 * H + h) = ad;
    }
    store_row_bf16<G>(a.dz + tok * H, H, lane, dy);
    if (sep_do) store_row_bf16<G>(a.d_o + tok * H, H, lane, dout);
  }
  __syncwarp();
  flush_acc(a.d_g, acc + 0 * H, H, lane);
  flush_acc(a.d_b, acc + 1 * H, H, lane);
  flush_acc(a.d_bias, acc + 2 * H, H, lane);
}

// ---------------------------------------------------------------------------
// LayerNorm backward from the saved pre-LayerNorm input z, H = 128.
// A half-warp owns a row (8 columns = one 16-byte load per lane), two rows per half-warp in flight;
// the column sums d_g / d_b stay in registers over all rows of the thread and are reduced once per CTA.
// ---------------------------------------------------------------------------
__device__ __forceinline__ float half_warp_sum(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ void unpack8_row(const uint4& u, float* f) {
  unpack_bf16x2(u.x, f[0], f[1]); unpack_bf16x2(u.y, f[2], f[3]);
  unpack_bf16x2(u.z, f[4], f[5]); unpack_bf16x2(u.w, f[6], f[7]);
}
__device__ __forceinline__ uint4 pack8_row(const float* f) {
  uint4 u;
  u.x = pack_bf16x2(f[0], f[1]); u.y = pack_bf16x2(f[2], f[3]);
  u.z = pack_bf16x2(f[4], f[5]); u.w = pack_bf16x2(f[6], f[7]);
  return u;
}

__global__ void __launch_bounds__(256) ln_bwd128_kernel(const pmgt_lnbwd_args a) {
  constexpr int H = 128, U = 2;
  __shared__ float red[16][2][H];
  const int lane = threadIdx.x & 31, hw = threadIdx.x >> 4;  // 16 half-warps per CTA
  const int c = (lane & 15) * 8;
  const long long hw0 = (long long)blockIdx.x * 16 + hw;
  const long long nhw = (long long)gridDim.x * 16;
  const float ks = a.dropout_p > 0.f ? 1.f / (1.f - a.dropout_p) : 1.f;
  const bool sep_do = a.d_o != nullptr && a.d_o != a.dz;
  float gam[8];
  {
    const float4 g0 = *reinterpret_cast<const float4*>(a.ln_g + c), g1 = *reinterpret_cast<const float4*>(a.ln_g + c + 4);
    gam[0] = g0.x; gam[1] = g0.y; gam[2] = g0.z; gam[3] = g0.w; gam[4] = g1.x; gam[5] = g1.y; gam[6] = g1.z; gam[7] = g1.w;
  }
  float dgam[8], dbet[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) dgam[j] = dbet[j] = 0.f;
  // all lanes of a warp run the same number of iterations (shuffles need full participation)
  const long long iters = (a.T + nhw * U - 1) / (nhw * U);
  for (long long itn = 0; itn < iters; ++itn) {
    uint4 zu[U], da[U], db[U];
    float4 df0[U], df1[U];
    bool ok[U];
    long long tok[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      tok[u] = (itn * U + u) * nhw + hw0;
      ok[u] = tok[u] < a.T;
      const long long o = (ok[u] ? tok[u] : 0) * H + c;
      zu[u] = *reinterpret_cast<const uint4*>(a.z + o);
      da[u] = a.dy_a ? *reinterpret_cast<const uint4*>(a.dy_a + o) : make_uint4(0, 0, 0, 0);
      db[u] = a.dy_b ? *reinterpret_cast<const uint4*>(a.dy_b + o) : make_uint4(0, 0, 0, 0);
      if (a.dy_f32) {
        df0[u] = *reinterpret_cast<const float4*>(a.dy_f32 + o);
        df1[u] = *reinterpret_cast<const float4*>(a.dy_f32 + o + 4);
      } else {
        df0[u] = df1[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float z[8], dy[8], t[8];
      unpack8_row(zu[u], z);
      unpack8_row(da[u], dy);
      unpack8_row(db[u], t);
      dy[0] += t[0] + df0[u].x; dy[1] += t[1] + df0[u].y; dy[2] += t[2] + df0[u].z; dy[3] += t[3] + df0[u].w;
      dy[4] += t[4] + df1[u].x; dy[5] += t[5] + df1[u].y; dy[6] += t[6] + df1[u].z; dy[7] += t[7] + df1[u].w;
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) s += z[j];
      const float mean = half_warp_sum(s) * (1.f / H);
      float q = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) { z[j] -= mean; q = fmaf(z[j], z[j], q); }
      const float rstd = rsqrtf(half_warp_sum(q) * (1.f / H) + a.ln_eps);
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        z[j] *= rstd;  // xhat
        if (ok[u]) { dgam[j] = fmaf(dy[j], z[j], dgam[j]); dbet[j] += dy[j]; }
        dy[j] *= gam[j];
        s1 += dy[j];
        s2 = fmaf(dy[j], z[j], s2);
      }
      s1 = half_warp_sum(s1) * (1.f / H);
      s2 = half_warp_sum(s2) * (1.f / H);
      float dz[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) dz[j] = rstd * (dy[j] - s1 - z[j] * s2);
      if (ok[u]) {
        *reinterpret_cast<uint4*>(a.dz + tok[u] * H + c) = pack8_row(dz);
        if (sep_do) {
          if (a.dropout_p > 0.f) {
            const uint64_t idx = (uint64_t)tok[u] * H + c;
            const uint32_t k8 = dropout_keep8(a.dropout_seed, a.dropout_site, idx, a.dropout_p);
#pragma unroll
            for (int j = 0; j < 8; ++j) dz[j] = (k8 >> j) & 1u ? dz[j] * ks : 0.f;
          }
          *reinterpret_cast<uint4*>(a.d_o + tok[u] * H + c) = pack8_row(dz);
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    red[hw][0][c + j] = dgam[j];
    red[hw][1][c + j] = dbet[j];
  }
  __syncthreads();
  {
    const int which = threadIdx.x >> 7, col = threadIdx.x & 127;
    float v = 0.f;
#pragma unroll
    for (int g = 0; g < 16; ++g) v += red[g][which][col];
    float* dst = which == 0 ? a.d_g : a.d_b;
    if (dst && v != 0.f) atomicAdd(dst + col, v);
  }
}

// ---------------------------------------------------------------------------
// column sums of a bf16 matrix: out[n] += sum_t x[t][n]
// ---------------------------------------------------------------------------
// 256 threads = (N / 8) column threads x `groups` row groups; every thread keeps 8 column sums of the rows
// it streams (4 independent 16-byte loads in flight), the row groups are then reduced through shared memory
// and ONE atomic per column per CTA is issued (contended same-address atomics were the old bottleneck).
__global__ void __launch_bounds__(256) colsum_kernel(const uint16_t* __restrict__ x, long long T, int N, long long ldx,
                                                     float* __restrict__ out, int rows_per_cta) {
  __shared__ float red[256 * 8];
  const int tpr = (N + 7) / 8;               // threads per row (<= 256)
  const int groups = blockDim.x / tpr;       // rows processed concurrently
  const int tg = threadIdx.x / tpr;
  const int tc = threadIdx.x % tpr;
  const int n = tc * 8;
  float s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const long long r0 = (long long)blockIdx.x * rows_per_cta;
  long long r1 = r0 + rows_per_cta;
  if (r1 > T) r1 = T;
  if (tg < groups) {
    long long r = r0 + tg;
    for (; r + 3ll * groups < r1; r += 4ll * groups) {
      uint4 u[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) u[q] = *reinterpret_cast<const uint4*>(x + (r + (long long)q * groups) * ldx + n);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float f[8];
        unpack_bf16x2(u[q].x, f[0], f[1]); unpack_bf16x2(u[q].y, f[2], f[3]);
        unpack_bf16x2(u[q].z, f[4], f[5]); unpack_bf16x2(u[q].w, f[6], f[7]);
#pragma unroll
        for (int j = 0; j < 8; ++j) s[j] += f[j];
      }
    }
    for (; r < r1; r += groups) {
      const uint4 u = *reinterpret_cast<const uint4*>(x + r * ldx + n);
      float f[8];
      unpack_bf16x2(u.x, f[0], f[1]); unpack_bf16x2(u.y, f[2], f[3]);
      unpack_bf16x2(u.z, f[4], f[5]); unpack_bf16x2(u.w, f[6], f[7]);
#pragma unroll
      for (int j = 0; j < 8; ++j) s[j] += f[j];
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[j * 256 + threadIdx.x] = (tg < groups) ? s[j] : 0.f;
  __syncthreads();
  // thread c < N sums column c over the row groups
  for (int c = threadIdx.x; c < N; c += blockDim.x) {
    const int tcc = c >> 3, j = c & 7;
    float v = 0.f;
    for (int gI = 0; gI < groups; ++gI) v += red[j * 256 + gI * tpr + tcc];
    if (v != 0.f) atomicAdd(out + c, v);
  }
}

__global__ void __launch_bounds__(256) cast_f32_bf16_kernel(const float* __restrict__ src, uint16_t* __restrict__ dst,
                                                            long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x * 4;
  for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
    if (i + 3 < n) {
      const float4 v = *reinterpret_cast<const float4*>(src + i);
      uint2 u;
      u.x = pack_bf16x2(v.x, v.y);
      u.y = pack_bf16x2(v.z, v.w);
      *reinterpret_cast<uint2*>(dst + i) = u;
    } else {
      for (long long j = i; j < n; ++j) dst[j] = float_to_bf16_bits(src[j]);
    }
  }
}

__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ x, long long n, float* __restrict__ out) {
  float s = 0.f;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float v = x[i];
    s += v * v;
  }
  s = warp_sum(s);
  __shared__ float red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += red[i];
    atomicAdd(out, t);
  }
}

// clip_grad_norm_ coefficient folded with the gradient scale: out = scale * min(1, max_norm / (sqrt(sumsq) * scale + 1e-6));
// sumsq is reset for the next step (one launch instead of seven tiny eager kernels between backward and AdamW)
__global__ void clip_coef_kernel(float* __restrict__ sumsq, float scale, float max_norm, float* __restrict__ out) {
  const float norm = sqrtf(sumsq[0]) * scale;
  out[0] = scale * fminf(1.f, max_norm / (norm + 1e-6f));
  sumsq[0] = 0.f;
}

__global__ void __launch_bounds__(256) gather_rows_kernel(const uint16_t* __restrict__ src, long long ld_src,
                                                          const long long* __restrict__ idx, long long n_rows, int D,
                                                          uint16_t* __restrict__ out, long long ld_out) {
  const int chunks = D / 8;
  const long long total = n_rows * chunks;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const long long r = i / chunks;
    const int c = (int)(i % chunks) * 8;
    const long long s = idx[r];
    uint4 v = make_uint4(0, 0, 0, 0);
    if (s >= 0) v = *reinterpret_cast<const uint4*>(src + s * ld_src + c);
    *reinterpret_cast<uint4*>(out + r * ld_out + c) = v;
  }
}

// dst[idx[r]][:] = src[r][:]   (idx unique; rows of dst not named by idx are left as they are)
__global__ void __launch_bounds__(256) scatter_rows_kernel(const uint16_t* __restrict__ src, long long ld_src,
                                                           const long long* __restrict__ idx, long long n_rows, int D,
                                                           uint16_t* __restrict__ dst, long long ld_dst) {
  const int chunks = D / 8;
  const long long total = n_rows * chunks;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const long long r = i / chunks;
    const int c = (int)(i % chunks) * 8;
    const long long d = idx[r];
    if (d >= 0) *reinterpret_cast<uint4*>(dst + d * ld_dst + c) = *reinterpret_cast<const uint4*>(src + r * ld_src + c);
  }
}

// AdamW, dense branch of DenseSparseAdamW (pmgt/optimizers.py:256-270)
__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                    float* __restrict__ m, float* __restrict__ v,
                                                    const uint8_t* __restrict__ decay_mask, long long n, float lr,
                                                    float beta1, float beta2, float eps, float wd, float inv_sqrt_bc2,
                                                    float step_size, float grad_scale,
                                                    const float* __restrict__ grad_scale_dev) {
  const float gs = grad_scale_dev ? *grad_scale_dev : grad_scale;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float gi = g[i] * gs;
    float pi = p[i];
    const float decay = (decay_mask == nullptr || decay_mask[i]) ? wd : 0.f;
    pi *= 1.f - lr * decay;
    const float mi = beta1 * m[i] + (1.f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) * inv_sqrt_bc2 + eps;
    p[i] = pi - step_size * (mi / denom);
  }
}

// hidden size 128 (the default encoder): half-warp-per-token kernels in embed128.cu
int embed128_supported(const pmgt_embed_args* a);
int embed128_fwd(const pmgt_embed_args* a, cudaStream_t st);
int embed128_bwd(const pmgt_embed_args* a, cudaStream_t st);
int ln_bwd_stream(const pmgt_lnbwd_args* a, cudaStream_t st);  // ln_bwd_stream.cu
int res_ln_fwd_wide(const pmgt_resln_args* a, cudaStream_t st);  // res_ln_wide.cu (H a multiple of 256)
int res_ln_bwd_wide(const pmgt_resln_args* a, cudaStream_t st);

static int persistent_grid(long long work_warps, int warps_per_cta, int ctas_per_sm) {
  long long need = (work_warps + warps_per_cta - 1) / warps_per_cta;
  long long cap = (long long)num_sms() * ctas_per_sm;
  if (need > cap) need = cap;
  if (need < 1) need = 1;
  return (int)need;
}

}  // namespace pmgt

using namespace pmgt;

// G = groups of 128 columns a lane walks (4 columns per lane and group): 1 for H <= 128, 6 for H <= 768 (BERT-base
// width: no predicated-off groups, smaller register footprint than the 1024-column variant), 8 up to H = 1024
#define PMGT_DISPATCH_G(H, CALL1, CALL6, CALL8)                                  \
  do {                                                                           \
    if ((H) <= 128) { CALL1; } else if ((H) <= 768) { CALL6; } else { CALL8; }   \
  } while (0)

extern "C" {

int pmgt_embed_fuse_fwd(const pmgt_embed_args* a, void* stream) {
  PMGT_REQUIRE(a && a->ev && a->et && a->w_att && a->b_att && a->pos && a->role && a->ln_g && a->ln_b && a->x_out,
               "pmgt_embed_fuse_fwd: null argument");
  PMGT_REQUIRE(a->H % 4 == 0 && a->H >= 4 && a->H <= 1024, "pmgt_embed_fuse_fwd: H must be a multiple of 4 in [4,1024]");
  PMGT_REQUIRE(a->L >= 1 && a->rows >= 0, "pmgt_embed_fuse_fwd: bad sizes");
  if (a->rows == 0) return PMGT_OK;
  if (embed128_supported(a)) return embed128_fwd(a, (cudaStream_t)stream);
  const int grid = persistent_grid(a->rows * a->L, kRowThreads / 32, 8);
  PMGT_DISPATCH_G(a->H, (embed_fuse_fwd_kernel<1><<<grid, kRowThreads, 0, (cudaStream_t)stream>>>(*a)),
                  (embed_fuse_fwd_kernel<6><<<grid, kRowThreads, 0, (cudaStream_t)stream>>>(*a)),
                  (embed_fuse_fwd_kernel<8><<<grid, kRowThreads, 0, (cudaStream_t)stream>>>(*a)));
  PMGT_LAUNCH_CHECK();
  return PMGT_OK;
}

int pmgt_embed_fuse_bwd(const pmgt_embed_args* a, void* stream) {
  PMGT_REQUIRE(a && a->ev && a->et && a->w_att && a->b_att && a->pos && a->role && a->ln_g && a->ln_b && a->dx &&
                   (a->row_idx ? (a->dev_acc && a->det_acc) : (a->dev && a->det)) && a->d_w_att && a->d_b_att && a->d_pos && a->d_role && a->d_ln_g && a->d_ln_b &&
                   a->d_bias_v && a->d_bias_t,
               "pmgt_embed_fuse_bwd: null argument");
  PMGT_REQUIRE(a->H % 4 == 0 && a->H >= 4 && a->H <= 1024, "pmgt_embed_fuse_bwd: H must be a multiple of 4 in [4,1024]");
  if (a->rows == 0) return PMGT_OK;
  if (embed128_supported(a)) return embed128_bwd(a, (cudaStream_t)stream);
  const size_t smem = (size_t)(kRowThreads / 32) * 9 * a->H * sizeof(float);
  const int grid = persistent_grid(a->rows, kRowThreads / 32, 2);
  if (a->H <= 128) {
    embed_fuse_bwd_kernel<1><<<grid, kRowThreads, smem, (cudaStream_t)stream>>>(*a);
  } else {
    static unsigned long long cfg = 0;
    if (first_use_on_device(cfg)) {
      PMGT_CHECK_CUDA(cudaFuncSetAttribute(embed_fuse_bwd_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    }
    PMGT_REQUIRE(smem <= 227 * 1024, "pmgt_embed_fuse_bwd: H too large for shared-memory accumulators");
    embed_fuse_bwd_kernel<8><<<grid, kRowThreads, smem, (cudaStream_t)stream>>>(*a);
  }
  PMGT_LAUNCH_CHECK();
  return PMGT_OK;
}

int pmgt_res_ln_fwd(const pmgt_resln_args* a, void* stream) {
  PMGT_REQUIRE(a && a->o && a->res && a->ln_g && a->ln_b && a->y, "pmgt_res_ln_fwd: null argument");
  PMGT_REQUIRE(a->H % 4 == 0 && a->H >= 4 && a->H <= 1024, "pmgt_res_ln_fwd: H must be a multiple of 4 in [4,1024]");
  if (a->T == 0) return PMGT_OK;
  {
    const int r = res_ln_fwd_wide(a, (cudaStream_t)stream);
    if (r != 0) return r < 0 ? r : PMGT_OK;
  }
  const int grid = persistent_grid(a->T, kRowThreads / 32, 8);
  PMGT_DISPATCH_G(a->H, (res_ln_fwd_kernel<1><<<grid, kRowThreads, 0, (cudaStream_t)stream>>>(*a)),
                  (res_ln_fwd_kernel<6><<<grid, kRowThreads, 0, (cudaStream_t)stream>>>(*a)),
                  (res_ln_fwd_kernel<8><<<grid, kRowThreads, 0, (cudaStream_t)stream>>>(*a)));
  PMGT_LAUNCH_CHECK();
  return PMGT_OK;
}

int pmgt_res_ln_bwd(const pmgt_resln_args* a, void* stream) {
  PMGT_REQUIRE(a && a->o && a->res && a->ln_g && (a->dy || a->dy_f32) && a->dz, "pmgt_res_ln_bwd: null argument");
  PMGT_REQUIRE(a->H % 4 == 0 && a->H >= 4 && a->H <= 1024, "pmgt_res_ln_bwd: H must be a multiple of 4 in [4,1024]");
  PMGT_REQUIRE(a->dropout_p == 0.f || (a->d_o && a->d_o != a->dz), "pmgt_res_ln_bwd: dropout needs a separate d_o buffer");
  if (a->T == 0) return PMGT_OK;
  {
    const int r = res_ln_bwd_wide(a, (cudaStream_t)stream);
    if (r != 0) return r < 0 ? r : PMGT_OK;
  }
  const size_t smem = (size_t)(kRowThreads / 32) * 3 * a->H * sizeof(float);
  const int grid = persistent_grid(a->T, kRowThreads / 32, 4);
  if (a->H <= 128) {
    res_ln_bwd_kernel<1><<<grid, kRowThreads, smem, (cudaStream_t)stream>>>(*a);
  } else if (a->H <= 768) {
    static unsigned long long cfg6 = 0;
    if (first_use_on_device(cfg6)) {
      PMGT_CHECK_CUDA(cudaFuncSetAttribute(res_ln_bwd_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    }
    res_ln_bwd_kernel<6><<<grid, kRowThreads, smem, (cudaStream_t)stream>>>(*a);
  } else {
    static unsigned long long cfg = 0;
    if (first_use_on_device(cfg)) {
      PMGT_CHECK_CUDA(cudaFuncSetAttribute(res_ln_bwd_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    }
    res_ln_bwd_kernel<8><<<grid, kRowThreads, smem, (cudaStream_t)stream>>>(*a);
  }
  PMGT_LAUNCH_CHECK();
  return PMGT_OK;
}

int pmgt_ln_bwd(const pmgt_lnbwd_args* a, void* stream) {
  PMGT_REQUIRE(a && a->z && a->ln_g && a->dz && (a->dy_a || a->dy_b || a->dy_f32), "pmgt_ln_bwd: null argument");
  PMGT_REQUIRE(a->H == 128, "pmgt_ln_bwd: H must be 128 (other sizes: pmgt_res_ln_bwd)");
  PMGT_REQUIRE(a->dropout_p == 0.f || (a->d_o && a->d_o != a->dz), "pmgt_ln_bwd: dropout needs a separate d_o buffer");
  PMGT_REQUIRE(a->dropout_p >= 0.f && a->dropout_p < 1.f, "pmgt_ln_bwd: bad dropout_p");
  if (a->T == 0) return PMGT_OK;
  {
    const int rc = ln_bwd_stream(a, (cudaStream_t)stream);  // bulk-copy-staged version; 1 = combination not covered
    if (rc != 1) return rc;
  }
  long long ctas = (a->T + 31) / 32;
  const long long cap = (long long)num_sms() * 8;
  if (ctas > cap) ctas = cap;
  ln_bwd128_kernel<<<(unsigned)ctas, 256, 0, (cudaStream_t)stream>>>(*a);
  PMGT_LAUNCH_CHECK();
  return PMGT_OK;
}

int pmgt_colsum_bf16(const uint16_t* x, int64_t T, int64_t N, int64_t ldx, float* out, void* stream) {
  PMGT_REQUIRE(x && out, "pmgt_colsum_bf16: null argument");
  PMGT_REQUIRE(N % 8 == 0 && N > 0 && ldx % 8 == 0, "pmgt_colsum_bf16: N and ldx must be multiples of 8");
  if (T == 0) return PMGT_OK;
  // wide matrices are processed in column panels of 2048
  for (int64_t nb = 0; nb < N; nb += 2048) {
    const int n = (int)((N - nb) < 2048 ? (N - nb) : 2048);
    const int tpr = n / 8;
    const int groups = 256 / tpr > 0 ? 256 / tpr : 1;
    const int threads = tpr * groups;
    long long ctas = (long long)num_sms() * 8;
    long long rows_per = (T + ctas - 1) / ctas;
    if (rows_per < groups) rows_per = groups;
    ctas = (T + rows_per - 1) / rows_per;
    colsum_kernel<<<(unsigned)ctas, threads, 0, (cudaStream_t)stream>>>(x + nb, T, n, ldx, out + nb, (int)rows_per);
    PMGT_LAUNCH_CHECK();
  }
  return PMGT_OK;
}

int pmgt_cast_f32_bf16(const float* src, uint16_t* dst, int64_t n, void* stream) {
  PMGT_REQUIRE(src && dst && n >= 0, "pmgt_cast_f32_bf16: bad argument");
  PMGT_REQUIRE(((uintptr_t)src & 15) == 0 && ((uintptr_t)dst & 7) == 0, "pmgt_cast_f32_bf16: alignment");
  if (n == 0) return PMGT_OK;
  long long blocks = (n / 4 + 255) / 256;
  long long cap = (long long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  cast_f32_bf16_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(src, dst, n);
  PMGT_LAUNCH_CHECK();
  return PMGT_OK;
}

int pmgt_clip_coef(float* sumsq, float scale, float max_norm, float* out, void* stream) {
  PMGT_REQUIRE(sumsq && out, "pmgt_clip_coef: null argument");
  clip_coef_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(sumsq, scale, max_norm, out);
  PMGT_LAUNCH_CHECK();
  return PMGT_OK;
}

int pmgt_sumsq_f32(const float* x, int64_t n, float* out, void* stream) {
  PMGT_REQUIRE(x && out && n >= 0, "pmgt_sumsq_f32: bad argument");
  if (n == 0) return PMGT_OK;
  long long blocks = (n + 255) / 256;
  long long cap = (long long)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  sumsq_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, n, out);
  PMGT_LAUNCH_CHECK();
  return PMGT_OK;
}

int pmgt_gather_rows_bf16(const uint16_t* src, int64_t ld_src, const int64_t* idx, int64_t n_rows, int64_t D,
                          uint16_t* out, int64_t ld_out, void* stream) {
  PMGT_REQUIRE(src && idx && out, "pmgt_gather_rows_bf16: null argument");
  PMGT_REQUIRE(D % 8 == 0 && ld_src % 8 == 0 && ld_out % 8 == 0, "pmgt_gather_rows_bf16: D/ld must be multiples of 8");
  if (n_rows == 0 || D == 0) return PMGT_OK;
  long long total = n_rows * (D / 8);
  long long blocks = (total + 255) / 256;
  long long cap = (long long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  gather_rows_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(src, ld_src, (const long long*)idx, n_rows, (int)D,
                                                                        out, ld_out);
  PMGT_LAUNCH_CHECK();
  return PMGT_OK;
}

int pmgt_scatter_rows_bf16(const uint16_t* src, int64_t ld_src, const int64_t* idx, int64_t n_rows, int64_t D,
                           uint16_t* dst, int64_t ld_dst, void* stream) {
  PMGT_REQUIRE(src && idx && dst, "pmgt_scatter_rows_bf16: null argument");
  PMGT_REQUIRE(D % 8 == 0 && ld_src % 8 == 0 && ld_dst % 8 == 0, "pmgt_scatter_rows_bf16: D/ld must be multiples of 8");
  if (n_rows == 0 || D == 0) return PMGT_OK;
  long long total = n_rows * (D / 8);
  long long blocks = (total + 255) / 256;
  long long cap = (long long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  scatter_rows_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(src, ld_src, (const long long*)idx, n_rows, (int)D,
                                                                         dst, ld_dst);
  PMGT_LAUNCH_CHECK();
  return PMGT_OK;
}

int pmgt_adamw_step(float* p, const float* g, float* m, float* v, const uint8_t* decay_mask, int64_t n, float lr,
                    float beta1, float beta2, float eps, float weight_decay, int64_t step, float grad_scale,
                    const float* grad_scale_dev, void* stream) {
  PMGT_REQUIRE(p && g && m && v && n >= 0 && step >= 1, "pmgt_adamw_step: bad argument");
  if (n == 0) return PMGT_OK;
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  long long blocks = (n + 255) / 256;
  long long cap = (long long)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  adamw_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(p, g, m, v, decay_mask, n, lr, beta1, beta2, eps,
                                                                  weight_decay, (float)(1.0 / sqrt(bc2)),
                                                                  (float)(lr / bc1), grad_scale, grad_scale_dev);
  PMGT_LAUNCH_CHECK();
  return PMGT_OK;
}

}  // extern "C"
