// Shared helpers for libpmgt_b200 (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/pmgt_b200.h"

namespace pmgt {

void set_error(const char* fmt, ...);

#define PMGT_CHECK_CUDA(expr)                                                              \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      pmgt::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,    \
                      __LINE__);                                                           \
      return PMGT_ERR_CUDA;                                                                \
    }                                                                                      \
  } while (0)

#define PMGT_REQUIRE(cond, ...)                                                            \
  do {                                                                                     \
    if (!(cond)) {                                                                         \
      pmgt::set_error(__VA_ARGS__);                                                        \
      return PMGT_ERR_INVALID;                                                             \
    }                                                                                      \
  } while (0)

#define PMGT_LAUNCH_CHECK()                                                                \
  do {                                                                                     \
    cudaError_t _e = cudaGetLastError();                                                   \
    if (_e != cudaSuccess) {                                                               \
      pmgt::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, \
                      __LINE__);                                                           \
      return PMGT_ERR_CUDA;                                                                \
    }                                                                                      \
  } while (0)

int num_sms();  // SM count of the current device (cached)

// Per-kernel one-time setup (cudaFuncSetAttribute for > 48 KiB of dynamic shared memory) is per DEVICE: `mask` is the
// call site's static bit set of devices already configured.  Returns true when the current device still needs it.
inline bool first_use_on_device(unsigned long long& mask) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;
  if ((mask >> dev) & 1ull) return false;
  mask |= 1ull << dev;
  return true;
}

// ---------------------------------------------------------------------------
// Programmatic dependent launch.  The encoder is a chain of ~170 short persistent kernels; launched the ordinary
// way, each one pays launch latency + its own prologue (barrier init, TMEM allocation, tensor-map fetch) + the idle
// tail of its predecessor (CTAs that ran one tile fewer).  With the stream-serialization attribute a kernel's CTAs
// are scheduled as soon as the predecessor's CTAs have all started and SM resources free up, run their prologue,
// and block in pdl_wait() until the predecessor has completed and its writes are visible.  Rules: every kernel
// launched through launch_kernel(pdl = true) executes pdl_wait() in EVERY thread before its first global-memory
// access, and calls pdl_launch_dependents() at its start.  PMGT_PDL=0 turns the attribute off.
// ---------------------------------------------------------------------------
bool pdl_enabled();
// Traversal order of the next persistent kernel of the encoder chain: the kernels alternate between ascending and
// descending tile order, so each one starts on the tiles its predecessor wrote LAST -- the part of a 75-300 MB
// activation that is still in the 126 MB L2.  Returns 1 for "descending" and flips the library-wide toggle.
int next_tile_order();
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(bool pdl, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                 Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (pdl && pdl_enabled()) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// ---------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11).  The same function is restated in
// plain C in oracle/philox_sampler.c; the two must agree bit for bit.
// ---------------------------------------------------------------------------
struct Philox4 {
  uint32_t x, y, z, w;
};

__host__ __device__ __forceinline__ uint32_t mulhi32(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
  return __umulhi(a, b);
#else
  return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}

__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2,
                                                         uint32_t c3, uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
  const uint32_t W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = mulhi32(M0, c0), lo0 = M0 * c0;
    uint32_t hi1 = mulhi32(M1, c2), lo1 = M1 * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0;
    uint32_t n1 = lo1;
    uint32_t n2 = hi0 ^ c3 ^ k1;
    uint32_t n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  return Philox4{c0, c1, c2, c3};
}

__host__ __device__ __forceinline__ uint32_t philox_word(const Philox4& p, int i) {
  return i == 0 ? p.x : (i == 1 ? p.y : (i == 2 ? p.z : p.w));
}

// Dropout stream: a counter-based hash, not Philox.  The element index is split into a block of 8 consecutive elements
// (idx >> 3) and a lane j = idx & 7; block b under key (seed, site) expands to four 32-bit words
//     w_i = dmix32((4 b + i) ^ k1) ^ k2,   i = 0..3      (dmix32 = two multiply / xor-shift rounds, a bijection of 2^32)
// and element j uses the 16-bit field j of the 128 bits (word j >> 1, half j & 1); it is KEPT when field >=
// round(p * 65536); the caller scales kept values by 1 / (1 - p).  The LayerNorm epilogues are issue-bound and spent
// ~45 % of their instructions in Philox4x32-10 (ten rounds for 8 elements); a dropout mask needs decorrelated bits,
// not a Crush-resistant generator -- the sampler, whose stream must be replayable bit for bit, stays on Philox.
// Every kernel (forward and the backward pass that regenerates the mask) goes through these helpers, so the masks
// agree by construction.
__device__ __forceinline__ uint32_t dropout_threshold(float p) { return (uint32_t)(p * 65536.0f + 0.5f); }

__device__ __forceinline__ uint32_t fmix32(uint32_t x) {
  x ^= x >> 16;
  x *= 0x85EBCA6Bu;
  x ^= x >> 13;
  x *= 0xC2B2AE35u;
  x ^= x >> 16;
  return x;
}
// the per-word mixer of the dropout stream: the MurmurHash3 finaliser without its first xor-shift (the input is a
// counter xor a key: its low bits already differ from word to word, and the first multiply carries them upward) --
// 6 instructions instead of 8 in epilogues that are bound by instruction issue
__device__ __forceinline__ uint32_t dmix32(uint32_t x) {
  x *= 0x85EBCA6Bu;
  x ^= x >> 15;
  x *= 0xC2B2AE35u;
  x ^= x >> 16;
  return x;
}

// keep bits of the 8 elements idx8 .. idx8 + 7 (idx8 a multiple of 8): bit j = keep element idx8 + j
__device__ __forceinline__ uint32_t dropout_keep8(uint64_t seed, uint32_t site, uint64_t idx8, float p) {
  const uint64_t blk = idx8 >> 3;
  // per-(seed, site) keys: loop-invariant, hoisted by the compiler
  const uint32_t k1 = fmix32((uint32_t)seed ^ (site * 0x9E3779B9u) ^ 0x5eedu);
  const uint32_t k2 = fmix32((uint32_t)(seed >> 32) + site) ^ ((uint32_t)(blk >> 30) * 0x9E3779B9u);
  const uint32_t c = (uint32_t)blk << 2;
  const uint32_t w0 = dmix32((c + 0u) ^ k1) ^ k2, w1 = dmix32((c + 1u) ^ k1) ^ k2;
  const uint32_t w2 = dmix32((c + 2u) ^ k1) ^ k2, w3 = dmix32((c + 3u) ^ k1) ^ k2;
  const uint32_t thr = dropout_threshold(p);
  uint32_t m = 0;
  m |= ((w0 & 0xffffu) >= thr) ? 1u : 0u;
  m |= ((w0 >> 16) >= thr) ? 2u : 0u;
  m |= ((w1 & 0xffffu) >= thr) ? 4u : 0u;
  m |= ((w1 >> 16) >= thr) ? 8u : 0u;
  m |= ((w2 & 0xffffu) >= thr) ? 16u : 0u;
  m |= ((w2 >> 16) >= thr) ? 32u : 0u;
  m |= ((w3 & 0xffffu) >= thr) ? 64u : 0u;
  m |= ((w3 >> 16) >= thr) ? 128u : 0u;
  return m;
}

// the four 32-bit words of block idx8 >> 3 (16-bit field j of the 128 bits belongs to element idx8 + j)
__device__ __forceinline__ void dropout_words8(uint64_t seed, uint32_t site, uint64_t idx8, uint32_t (&w)[4]) {
  const uint64_t blk = idx8 >> 3;
  const uint32_t k1 = fmix32((uint32_t)seed ^ (site * 0x9E3779B9u) ^ 0x5eedu);
  const uint32_t k2 = fmix32((uint32_t)(seed >> 32) + site) ^ ((uint32_t)(blk >> 30) * 0x9E3779B9u);
  const uint32_t c = (uint32_t)blk << 2;
#pragma unroll
  for (int i = 0; i < 4; ++i) w[i] = dmix32((c + (uint32_t)i) ^ k1) ^ k2;
}

// single element (row-wise kernels call it for 4 consecutive elements of one block; the hash is shared)
__device__ __forceinline__ bool dropout_keep(uint64_t seed, uint32_t site, uint64_t idx, float p) {
  return (dropout_keep8(seed, site, idx & ~(uint64_t)7, p) >> (uint32_t)(idx & 7)) & 1u;
}

// ---------------------------------------------------------------------------
// small device utilities
// ---------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ float bf16_bits_to_float(uint16_t b) {
  return __uint_as_float(((uint32_t)b) << 16);
}
__device__ __forceinline__ uint16_t float_to_bf16_bits(float f) {
  __nv_bfloat16 h = __float2bfloat16_rn(f);
  return *reinterpret_cast<uint16_t*>(&h);
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void unpack_bf16x2(uint32_t v, float& lo, float& hi) {
  lo = __uint_as_float(v << 16);
  hi = __uint_as_float(v & 0xffff0000u);
}

// erf-GELU (transformers ACT2FN["gelu"]) and its derivative.  erf by Abramowitz & Stegun 7.1.26 (|error| <= 1.5e-7,
// far below the bf16 rounding of the stored activations): one reciprocal, one exp2, five FMAs -- the libdevice
// erff costs ~3x the instructions and these epilogues are issue-bound.  The exponential e^{-x^2/2} is shared
// between erf(x / sqrt 2) and the Gaussian density of the derivative.
__device__ __forceinline__ void gelu_parts(float x, float& cdf, float& pdf_x) {
  const float z = fabsf(x) * 0.70710678118654752440f;           // |x| / sqrt(2)
  const float t = __fdividef(1.0f, fmaf(0.3275911f, z, 1.0f));
  const float e = exp2f(-0.72134752044448170368f * x * x);      // e^{-x^2/2} = 2^{-x^2 / (2 ln 2)}
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float erfc_abs = poly * t * e;                          // 1 - erf(|x| / sqrt 2)
  const float half = 0.5f * erfc_abs;
  cdf = x >= 0.f ? 1.0f - half : half;                          // Phi(x)
  pdf_x = x * 0.39894228040143267794f * e;                      // x * phi(x)
}
__device__ __forceinline__ float gelu_erf(float x) {
  float cdf, pdf_x;
  gelu_parts(x, cdf, pdf_x);
  return x * cdf;
}
__device__ __forceinline__ float gelu_erf_grad(float x) {
  float cdf, pdf_x;
  gelu_parts(x, cdf, pdf_x);
  return cdf + pdf_x;
}

// Packed fp32 pairs (sm_100 FFMA2 / FMUL2 / FADD2): the epilogues are bound by instruction issue, and one packed
// instruction does the work of two
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float a, float b) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ f32x2 pk2u(uint32_t a, uint32_t b) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ f32x2 pk1(float a) { return pk2(a, a); }
__device__ __forceinline__ void up2(f32x2 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ void up2u(f32x2 v, uint32_t& a, uint32_t& b) { asm("mov.b64 {%0, %1}, %2;" : "=r"(a), "=r"(b) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ uint32_t pack2_bf16(f32x2 v) { float a, b; up2(v, a, b); return pack_bf16x2(a, b); }
__device__ __forceinline__ f32x2 unpack2_bf16(uint32_t v) { return pk2u(v << 16, v & 0xffff0000u); }

// erf-GELU of two elements at once (gelu_parts with packed arithmetic; the 1/2 of Phi folded into the coefficients):
// h = x Phi(x) and, if asked, gelu'(x) = Phi(x) + x phi(x).  ~11 instructions per element instead of ~17: the GELU
// epilogues are bound by instruction issue.
__device__ __forceinline__ void gelu_pair(f32x2 x, f32x2& h, f32x2* grad) {
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
  h = mul2(x, cdf);
  if (grad) *grad = add2(cdf, mul2(mul2(x, e), pk1(0.39894228040143267794f)));   // + x phi(x)
}

}  // namespace pmgt
