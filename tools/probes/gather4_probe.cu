// Does cp.async.bulk.tensor.2d ... tile::gather4 fill a 128B-swizzled UMMA slab the way the hand-swizzled cp.async gather
// of gather_proj.cu does?  Loads 4 arbitrary rows of a [rows][64] bf16 table with each candidate tensor-map box shape
// and compares the shared-memory image with the expected swizzled layout.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gather4_probe gather4_probe.cu -lcuda
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>

__global__ void k(const __grid_constant__ CUtensorMap tm, int r0, int r1, int r2, int r3, int col, uint16_t* out) {
  __shared__ __align__(1024) unsigned char slab[4096];
  __shared__ uint64_t bar;
  const uint32_t sb = (uint32_t)__cvta_generic_to_shared(slab), bb = (uint32_t)__cvta_generic_to_shared(&bar);
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) reinterpret_cast<uint16_t*>(slab)[i] = 0xdead;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bb));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bb), "r"(512u) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                 ::"r"(sb + 512u), "l"(reinterpret_cast<uint64_t>(&tm)), "r"(bb), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3) : "memory");
  }
  uint32_t done = 0;
  int spins = 0;
  while (!done && ++spins < (1 << 22)) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(bb) : "memory");
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) out[i] = reinterpret_cast<uint16_t*>(slab)[i];
  if (threadIdx.x == 0) out[2048] = (uint16_t)done;
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  const int only = argc > 1 ? atoi(argv[1]) : 0;
  const int rows = 1000, cols = 192;   // row pitch 384 B; the box takes 64 columns at column offset 64
  std::vector<uint16_t> h((size_t)rows * cols);
  for (int r = 0; r < rows; ++r) for (int c = 0; c < cols; ++c) h[(size_t)r * cols + c] = (uint16_t)((r * 7 + c * 3) & 0x7fff);
  uint16_t *d, *out;
  cudaMalloc(&d, h.size() * 2); cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
  cudaMalloc(&out, 4200 * 2);
  void* fp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  EncodeFn enc = (EncodeFn)fp;
  const int rr[4] = {5, 917, 33, 2};
  for (int box_rows : {4, 1}) {
    if (only && box_rows != only) continue;
    CUtensorMap tm; memset(&tm, 0, sizeof(tm));
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows}; cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)box_rows}; cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("box {64, %d}: encode rc=%d\n", box_rows, (int)r);
    if (r != CUDA_SUCCESS) continue;
    cudaMemset(out, 0, 4200 * 2);
    k<<<1, 128>>>(tm, rr[0], rr[1], rr[2], rr[3], 64, out);
    cudaError_t e = cudaDeviceSynchronize();
    printf("  kernel: %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) { cudaGetLastError(); continue; }
    std::vector<uint16_t> o(2049); cudaMemcpy(o.data(), out, 2049 * 2, cudaMemcpyDeviceToHost);
    printf("  barrier completed: %d\n", (int)o[2048]);
    // expected: slab row (4 + i) <- table row rr[i], columns 64..127, 16-byte chunk c at ((c ^ (row & 7)) << 4)
    int bad = 0, untouched_ok = 1;
    for (int i = 0; i < 4; ++i) {
      const int srow = 4 + i;
      for (int c = 0; c < 64; ++c) {
        const int chunk = c >> 3, within = c & 7;
        const int pos = srow * 64 + ((chunk ^ (srow & 7)) << 3) + within;
        const uint16_t want = h[(size_t)rr[i] * cols + 64 + c];
        if (o[pos] != want) ++bad;
      }
    }
    for (int i = 0; i < 4 * 64; ++i) if (o[i] != 0xdead) untouched_ok = 0;
    for (int i = 8 * 64; i < 2048; ++i) if (o[i] != 0xdead) untouched_ok = 0;
    printf("  swizzled rows 4..7 mismatches: %d of 256; other rows untouched: %d\n", bad, untouched_ok);
  }
  return 0;
}
