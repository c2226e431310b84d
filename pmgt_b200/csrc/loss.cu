// K4 / K5: fused reduction kernels for the two pre-training losses.
//   GSR  PMGTGraphConstructLoss (modeling_pmgt.py:537-546) applied per target
//        and averaged over targets (models.py:104-127): cosine logits against
//        the target's position-0 state + BCEWithLogits.
//   NFR  the MSE half of PMGTNodeConstructLoss (modeling_pmgt.py:566-569)
//        against gathered rows of the frozen feature table.
#include "common.cuh"

namespace pmgt {

constexpr float kNormEps = 1e-12f;  // F.normalize default eps

__device__ __forceinline__ float block_sum_128(float v, float* red) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  return red[0] + red[1] + red[2] + red[3];
}

// one CTA (128 threads = 4 warps) per target; warps stride over the target's pairs
__global__ void __launch_bounds__(128) gsr_fwd_kernel(const pmgt_gsr_args a) {
  __shared__ float red[4];
  __shared__ float loss_w[4];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int H = a.H;
  for (long long b = blockIdx.x; b < a.B; b += gridDim.x) {
    const float* t = a.tgt_h + b * a.ld_t;
    float tt = 0.f;
    for (int h = threadIdx.x; h < H; h += 128) tt += t[h] * t[h];
    tt = block_sum_128(tt, red);
    const float tn = fmaxf(sqrtf(tt), kNormEps);
    const long long p0 = a.pair_off[b], p1 = a.pair_off[b + 1];
    float lsum = 0.f;
    for (long long p = p0 + w; p < p1; p += 4) {
      const float* x = a.pair_h + p * a.ld_p;
      float xx = 0.f, xt = 0.f;
      for (int h = lane; h < H; h += 32) { const float xv = x[h]; xx += xv * xv; xt += xv * t[h]; }
      xx = warp_sum(xx);
      xt = warp_sum(xt);
      const float logit = xt / (fmaxf(sqrtf(xx), kNormEps) * tn);
      const float y = a.labels[p];
      const float l = fmaxf(logit, 0.f) - logit * y + log1pf(__expf(-fabsf(logit)));
      if (lane == 0) a.logits[p] = logit;
      lsum += l;
    }
    if (lane == 0) loss_w[w] = lsum;
    __syncthreads();
    if (threadIdx.x == 0) {
      const float s = loss_w[0] + loss_w[1] + loss_w[2] + loss_w[3];
      // mean over this target's pairs, then mean over targets (0/0 -> NaN like the reference)
      atomicAdd(a.loss_out, s / (float)(p1 - p0) / (float)a.B);
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(128) gsr_bwd_kernel(const pmgt_gsr_args a) {
  extern __shared__ float dt_acc[];  // [4][H] per-warp partial of d target
  __shared__ float red[4];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int H = a.H;
  const float g = *a.grad_out;
  for (long long b = blockIdx.x; b < a.B; b += gridDim.x) {
    const float* t = a.tgt_h + b * a.ld_t;
    float tt = 0.f;
    for (int h = threadIdx.x; h < H; h += 128) tt += t[h] * t[h];
    tt = block_sum_128(tt, red);
    const float tn = fmaxf(sqrtf(tt), kNormEps);
    for (int h = lane; h < H; h += 32) dt_acc[w * H + h] = 0.f;
    const long long p0 = a.pair_off[b], p1 = a.pair_off[b + 1];
    const float scale = g / ((float)(p1 - p0) * (float)a.B);
    for (long long p = p0 + w; p < p1; p += 4) {
      const float* x = a.pair_h + p * a.ld_p;
      float xx = 0.f, xt = 0.f;
      for (int h = lane; h < H; h += 32) { const float xv = x[h]; xx += xv * xv; xt += xv * t[h]; }
      xx = warp_sum(xx);
      xt = warp_sum(xt);
      const float xn = fmaxf(sqrtf(xx), kNormEps);
      const float logit = xt / (xn * tn);
      const float dlogit = scale * (1.f / (1.f + __expf(-logit)) - a.labels[p]);
      float* dx = a.d_pair + p * a.ld_p;
      for (int h = lane; h < H; h += 32) {
        const float xh = x[h] / xn, th = t[h] / tn;
        dx[h] = dlogit * (th - logit * xh) / xn;
        dt_acc[w * H + h] += dlogit * (xh - logit * th) / tn;
      }
    }
    __syncthreads();
    float* dt = a.d_tgt + b * a.ld_t;
    for (int h = threadIdx.x; h < H; h += 128)
      dt[h] = dt_acc[h] + dt_acc[H + h] + dt_acc[2 * H + h] + dt_acc[3 * H + h];
    __syncthreads();
  }
}

// NFR: sum over Mm x D of (proj - table[target])^2
__global__ void __launch_bounds__(256) nfr_mse_kernel(const pmgt_nfr_args a, int write_grad) {
  const int chunks = (int)(a.D / 8);
  const long long total = a.Mm * chunks;
  const long long stride = (long long)gridDim.x * blockDim.x;
  float s = 0.f;
  const float coef = write_grad ? (*a.grad_out) * a.weight * 2.f / ((float)a.Mm * (float)a.D) : 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const long long r = i / chunks;
    const int c = (int)(i - r * chunks) * 8;
    const uint4 pu = *reinterpret_cast<const uint4*>(a.proj + r * a.ld_proj + c);
    const uint4 tu = *reinterpret_cast<const uint4*>(a.table + a.target_ids[r] * a.ld_table + c);
    float pf[8], tf[8], d[8];
    unpack_bf16x2(pu.x, pf[0], pf[1]); unpack_bf16x2(pu.y, pf[2], pf[3]);
    unpack_bf16x2(pu.z, pf[4], pf[5]); unpack_bf16x2(pu.w, pf[6], pf[7]);
    unpack_bf16x2(tu.x, tf[0], tf[1]); unpack_bf16x2(tu.y, tf[2], tf[3]);
    unpack_bf16x2(tu.z, tf[4], tf[5]); unpack_bf16x2(tu.w, tf[6], tf[7]);
#pragma unroll
    for (int j = 0; j < 8; ++j) { d[j] = pf[j] - tf[j]; s += d[j] * d[j]; }
    if (write_grad) {
      uint4 o;
      o.x = pack_bf16x2(coef * d[0], coef * d[1]); o.y = pack_bf16x2(coef * d[2], coef * d[3]);
      o.z = pack_bf16x2(coef * d[4], coef * d[5]); o.w = pack_bf16x2(coef * d[6], coef * d[7]);
      *reinterpret_cast<uint4*>(a.dproj + r * a.ld_proj + c) = o;
    }
  }
  if (!write_grad) {
    s = warp_sum(s);
    __shared__ float red[8];
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
      for (int i = 0; i < 8; ++i) t += red[i];
      atomicAdd(a.loss_out, a.weight * t / ((float)a.Mm * (float)a.D));
    }
  }
}

__global__ void add_nan_kernel(float* out) { atomicAdd(out, __int_as_float(0x7fc00000)); }

}  // namespace pmgt

using namespace pmgt;

extern "C" {

int pmgt_gsr_fwd(const pmgt_gsr_args* a, void* stream) {
  PMGT_REQUIRE(a && a->tgt_h && a->pair_off && a->labels && a->logits && a->loss_out && (a->SP == 0 || a->pair_h),
               "pmgt_gsr_fwd: null argument");
  PMGT_REQUIRE(a->H >= 1 && a->B >= 0, "pmgt_gsr_fwd: bad sizes");
  if (a->B == 0) return PMGT_OK;
  long long grid = a->B;
  long long cap = (long long)num_sms() * 16;
  if (grid > cap) grid = cap;
  gsr_fwd_kernel<<<(unsigned)grid, 128, 0, (cudaStream_t)stream>>>(*a);
  PMGT_LAUNCH_CHECK();
  return PMGT_OK;
}

int pmgt_gsr_bwd(const pmgt_gsr_args* a, void* stream) {
  PMGT_REQUIRE(a && a->tgt_h && a->pair_off && a->labels && a->grad_out && a->d_tgt && (a->SP == 0 || (a->pair_h && a->d_pair)),
               "pmgt_gsr_bwd: null argument");
  if (a->B == 0) return PMGT_OK;
  const size_t smem = 4 * (size_t)a->H * sizeof(float);
  PMGT_REQUIRE(smem <= 48 * 1024, "pmgt_gsr_bwd: H too large");
  long long grid = a->B;
  long long cap = (long long)num_sms() * 16;
  if (grid > cap) grid = cap;
  gsr_bwd_kernel<<<(unsigned)grid, 128, smem, (cudaStream_t)stream>>>(*a);
  PMGT_LAUNCH_CHECK();
  return PMGT_OK;
}

static int nfr_launch(const pmgt_nfr_args* a, int write_grad, void* stream) {
  PMGT_REQUIRE(a->D % 8 == 0 && a->ld_proj % 8 == 0 && a->ld_table % 8 == 0, "pmgt_nfr_mse: D/ld must be multiples of 8");
  if (a->Mm == 0) {
    if (!write_grad) { add_nan_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(a->loss_out); PMGT_LAUNCH_CHECK(); }
    return PMGT_OK;
  }
  long long total = a->Mm * (a->D / 8);
  long long blocks = (total + 255) / 256;
  long long cap = (long long)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  nfr_mse_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(*a, write_grad);
  PMGT_LAUNCH_CHECK();
  return PMGT_OK;
}

int pmgt_nfr_mse_fwd(const pmgt_nfr_args* a, void* stream) {
  PMGT_REQUIRE(a && a->loss_out && (a->Mm == 0 || (a->proj && a->table && a->target_ids)), "pmgt_nfr_mse_fwd: null argument");
  return nfr_launch(a, 0, stream);
}

int pmgt_nfr_mse_bwd(const pmgt_nfr_args* a, void* stream) {
  PMGT_REQUIRE(a && a->grad_out && (a->Mm == 0 || (a->proj && a->table && a->target_ids && a->dproj)),
               "pmgt_nfr_mse_bwd: null argument");
  return nfr_launch(a, 1, stream);
}

}  // extern "C"
