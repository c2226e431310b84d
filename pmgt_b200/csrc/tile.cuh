// Token-tile "images" shared by the persistent tcgen05 kernels (linear_tile.cu, ffn_block.cu, attn_block.cu).
//
// An image is a [128 rows][128 cols] bf16 block stored as two 64-column slabs of 128 rows x 128 bytes in the
// 128-byte TMA/UMMA swizzle (32 KiB).  One image serves as a K-major operand (K = its columns) and as an MN-major
// operand (K = its rows) of a 128 x 128 x 128 tcgen05 product.  sm_100a only.
#pragma once
#include "umma.cuh"

namespace pmgt {

constexpr int kImgBytes = 32768;
constexpr int kSlabBytes = 16384;

// 16-byte chunk c8 (8 columns) of row r of an image
__device__ __forceinline__ uint32_t img_off(int r, int c8) {
  return (uint32_t)((c8 >> 3) * kSlabBytes + r * 128 + (((c8 & 7) ^ (r & 7)) << 4));
}

__device__ __forceinline__ void mma_128x128x128(uint32_t tmem_d, uint32_t a_img, bool a_mn, uint32_t b_img, bool b_mn,
                                                uint32_t idesc, bool accumulate_first) {
#pragma unroll
  for (int ks = 0; ks < 8; ++ks) {
    const uint64_t da = a_mn ? umma_desc(a_img + ks * 2048, kSlabBytes, 1024)
                             : umma_desc(a_img + (ks >> 2) * kSlabBytes + (ks & 3) * 32, 16, 1024);
    const uint64_t db = b_mn ? umma_desc(b_img + ks * 2048, kSlabBytes, 1024)
                             : umma_desc(b_img + (ks >> 2) * kSlabBytes + (ks & 3) * 32, 16, 1024);
    umma_bf16(tmem_d, da, db, idesc, (accumulate_first || ks > 0) ? 1u : 0u);
  }
}

__host__ __device__ constexpr uint32_t make_idesc(bool a_mn, bool b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
         ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

__device__ __forceinline__ void ld8f(const float* __restrict__ p, float (&v)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  const float4 b = __ldg(reinterpret_cast<const float4*>(p + 4));
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ uint4 pack8f(const float* v) {
  uint4 o;
  o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]);
  o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
  return o;
}
__device__ __forceinline__ void unpack8f(const uint4& u, float* v) {
  unpack_bf16x2(u.x, v[0], v[1]); unpack_bf16x2(u.y, v[2], v[3]);
  unpack_bf16x2(u.z, v[4], v[5]); unpack_bf16x2(u.w, v[6], v[7]);
}

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}


__device__ __forceinline__ void tmem_st_x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

}  // namespace pmgt
