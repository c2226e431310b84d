// TMA row gather (cp.async.bulk.tensor.2d ... tile::gather4) into a shared-memory ring of 128B-swizzled UMMA slabs, no
// consumer work: how fast can ONE producer warp per CTA fetch random table rows with the stage geometry of
// gather_proj.cu?  The counterpart of ring_probe.cu (same table, ids, token count), which does it with 128 threads x
// 16-byte cp.async.  A stage = ROWS rows x SLABS slabs of 64 bf16 columns (128 B); one gather4 instruction brings 4 rows
// x 128 B, so a stage is ROWS/4 x SLABS instructions spread over the warp's lanes.
// `./gather4_ring 3072` varies the stage geometry / ring depth / CTAs per SM with one issuing warp (k_g4); `./gather4_ring 3072
// sweep` varies the number of issuing warps and lanes (k_g4w) -- the result that shaped gather_proj.cu: an instruction holds
// its issuing thread for ~120 clocks and a warp serialises its lanes, so the gather needs 8-16 warps with one lane each.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gather4_ring gather4_ring.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>

__device__ int g_timeout;

__device__ __forceinline__ bool wait_bar(uint32_t bar, uint32_t parity) {
  for (int i = 0; i < (1 << 22); ++i) {
    uint32_t done;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (done) return true;
  }
  g_timeout = 1;
  return false;
}

template <int ROWS, int SLABS, int RING>
__global__ void __launch_bounds__(64) k_g4(const __grid_constant__ CUtensorMap tm, const int* __restrict__ ids, int T, int cols) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  __shared__ uint64_t full[RING], empty[RING];
  __shared__ int rid[2][128];
  constexpr int STAGE = ROWS * SLABS * 128;
  constexpr int NI = ROWS / 4 * SLABS;  // gather4 instructions per stage
  const uint32_t sbase = ((uint32_t)__cvta_generic_to_shared(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < RING; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&full[s])));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&empty[s])));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int stages_per_tile = (128 / ROWS) * (cols / (SLABS * 64));
  int g = 0;
  if (warp == 0) {  // producer
    int it = 0;
    for (int tile = blockIdx.x; tile * 128 < T; tile += gridDim.x, ++it) {
      int* r = rid[it & 1];
      for (int i = lane; i < 128; i += 32) r[i] = ids[min(tile * 128 + i, T - 1)];
      __syncwarp();
      for (int r0 = 0; r0 < 128; r0 += ROWS) {
        for (int c0 = 0; c0 < cols; c0 += SLABS * 64, ++g) {
          const int s = g % RING;
          const uint32_t fb = (uint32_t)__cvta_generic_to_shared(&full[s]), eb = (uint32_t)__cvta_generic_to_shared(&empty[s]);
          if (!wait_bar(eb, ((g / RING) & 1) ^ 1)) return;
          if (lane == 0)
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fb), "r"((uint32_t)STAGE) : "memory");
          __syncwarp();
          const uint32_t dst = sbase + s * STAGE;
#pragma unroll
          for (int i = lane; i < NI; i += 32) {
            const int grp = i / SLABS, j = i % SLABS;  // row group of 4, slab
            const int* q = r + r0 + grp * 4;
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes "
                         "[%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                         ::"r"(dst + j * (ROWS * 128) + grp * 512), "l"(reinterpret_cast<uint64_t>(&tm)), "r"(fb),
                           "r"(c0 + j * 64), "r"(q[0]), "r"(q[1]), "r"(q[2]), "r"(q[3]) : "memory");
          }
        }
      }
    }
  } else if (lane == 0) {  // consumer: release every landed stage at once
    for (int tile = blockIdx.x; tile * 128 < T; tile += gridDim.x) {
      for (int k = 0; k < stages_per_tile; ++k, ++g) {
        const int s = g % RING;
        if (!wait_bar((uint32_t)__cvta_generic_to_shared(&full[s]), (g / RING) & 1)) return;
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((uint32_t)__cvta_generic_to_shared(&empty[s])) : "memory");
      }
    }
  }
}

// PW producer warps per CTA, LPW issuing lanes in each (a warp serialises the TMA instructions of its lanes)
template <int ROWS, int SLABS, int RING, int PW, int LPW>
__global__ void __launch_bounds__((PW + 1) * 32) k_g4w(const __grid_constant__ CUtensorMap tm, const int* __restrict__ ids, int T, int cols) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  __shared__ uint64_t full[RING], empty[RING];
  constexpr int STAGE = ROWS * SLABS * 128;
  constexpr int NI = ROWS / 4 * SLABS;
  const uint32_t sbase = ((uint32_t)__cvta_generic_to_shared(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < RING; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&full[s])));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&empty[s])));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int stages_per_tile = (128 / ROWS) * (cols / (SLABS * 64));
  int g = 0;
  if (warp < PW) {
    for (int tile = blockIdx.x; tile * 128 < T; tile += gridDim.x) {
      for (int r0 = 0; r0 < 128; r0 += ROWS) {
        for (int c0 = 0; c0 < cols; c0 += SLABS * 64, ++g) {
          const int s = g % RING;
          const uint32_t fb = (uint32_t)__cvta_generic_to_shared(&full[s]), eb = (uint32_t)__cvta_generic_to_shared(&empty[s]);
          if (!wait_bar(eb, ((g / RING) & 1) ^ 1)) return;
          if (warp == 0 && lane == 0)
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fb), "r"((uint32_t)STAGE) : "memory");
          const uint32_t dst = sbase + s * STAGE;
          if (lane < LPW) {
            for (int i = warp * LPW + lane; i < NI; i += PW * LPW) {
              const int grp = i / SLABS, j = i % SLABS;
              const int4 q = *reinterpret_cast<const int4*>(ids + min(tile * 128 + r0 + grp * 4, T - 4));
              asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes "
                           "[%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                           ::"r"(dst + j * (ROWS * 128) + grp * 512), "l"(reinterpret_cast<uint64_t>(&tm)), "r"(fb),
                             "r"(c0 + j * 64), "r"(q.x), "r"(q.y), "r"(q.z), "r"(q.w) : "memory");
            }
          }
        }
      }
    }
  } else if (lane == 0) {
    for (int tile = blockIdx.x; tile * 128 < T; tile += gridDim.x) {
      for (int k = 0; k < stages_per_tile; ++k, ++g) {
        const int s = g % RING;
        if (!wait_bar((uint32_t)__cvta_generic_to_shared(&full[s]), (g / RING) & 1)) return;
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((uint32_t)__cvta_generic_to_shared(&empty[s])) : "memory");
      }
    }
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  const int N = 1000002, T = 294912;
  const int rowb = argc > 1 ? atoi(argv[1]) : 3072;
  const int cols = rowb / 2;
  unsigned char* tab; int* ids;
  cudaMalloc(&tab, (size_t)N * rowb); cudaMemset(tab, 1, (size_t)N * rowb);
  cudaMalloc(&ids, T * 4);
  std::vector<int> h(T); srand(1); for (auto& x : h) x = (int)(((unsigned)rand() * 32768u + (unsigned)rand()) % N);
  cudaMemcpy(ids, h.data(), T * 4, cudaMemcpyHostToDevice);
  void* fp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  EncodeFn enc = (EncodeFn)fp;
  CUtensorMap tms[2];
  const CUtensorMapL2promotion promo[2] = {CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B};
  for (int p = 0; p < 2; ++p) {
    memset(&tms[p], 0, sizeof(CUtensorMap));
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)N}; cuuint64_t strides[1] = {(cuuint64_t)rowb};
    cuuint32_t box[2] = {64, 1}; cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&tms[p], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, tab, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, promo[p], CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
  }
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto run = [&](const char* name, auto kern, int ctas_per_sm, int smem, int p, int slabs, int threads = 64) {
    if (cols % (slabs * 64)) { printf("%-58s skipped (row width)\n", name); return; }
    smem += 1024;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    int zero = 0; cudaMemcpyToSymbol(g_timeout, &zero, 4);
    for (int i = 0; i < 2; ++i) kern<<<148 * ctas_per_sm, threads, smem>>>(tms[p], ids, T, cols);
    cudaError_t e = cudaDeviceSynchronize();
    int to = 0; cudaMemcpyFromSymbol(&to, g_timeout, 4);
    if (e != cudaSuccess || to) { printf("%-58s FAILED (%s, timeout %d)\n", name, cudaGetErrorString(e), to); return; }
    cudaEventRecord(e0); for (int i = 0; i < 10; ++i) kern<<<148 * ctas_per_sm, threads, smem>>>(tms[p], ids, T, cols);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 10;
    printf("%-58s %8.1f us  %7.1f GB/s  (%s)\n", name, ms * 1e3, (double)T * rowb / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
  };
  if (argc <= 2) {
  run("g4 128r x 256B, ring 5 (160K), 1 CTA, L2 256B", k_g4<128, 2, 5>, 1, 5 * 32768, 0, 2);
  run("g4 128r x 256B, ring 5 (160K), 1 CTA, L2 128B", k_g4<128, 2, 5>, 1, 5 * 32768, 1, 2);
  run("g4 128r x 256B, ring 6 (192K), 1 CTA, L2 256B", k_g4<128, 2, 6>, 1, 6 * 32768, 0, 2);
  run("g4 128r x 128B, ring 10 (160K), 1 CTA, L2 256B", k_g4<128, 1, 10>, 1, 10 * 16384, 0, 1);
  run("g4 128r x 512B, ring 3 (192K), 1 CTA, L2 256B", k_g4<128, 4, 3>, 1, 3 * 65536, 0, 4);
  run("g4 64r x 512B, ring 5 (160K), 1 CTA, L2 256B", k_g4<64, 4, 5>, 1, 5 * 32768, 0, 4);
  run("g4 32r x 1024B, ring 5 (160K), 1 CTA, L2 256B", k_g4<32, 8, 5>, 1, 5 * 32768, 0, 8);
  run("g4 16r x 1536B, ring 6 (144K), 1 CTA, L2 256B", k_g4<16, 12, 6>, 1, 6 * 24576, 0, 12);
  run("g4 128r x 256B, ring 3 (96K), 2 CTA, L2 256B", k_g4<128, 2, 3>, 2, 3 * 32768, 0, 2);
  run("g4 128r x 256B, ring 2 (64K), 3 CTA, L2 256B", k_g4<128, 2, 2>, 3, 2 * 32768, 0, 2);
  run("g4 32r x 1024B, ring 3 (96K), 2 CTA, L2 256B", k_g4<32, 8, 3>, 2, 3 * 32768, 0, 8);
  }
  if (argc > 2) {  // producer-warp sweep
    run("g4w 128r x 256B ring 5, 1 warp x 1 lane", k_g4w<128, 2, 5, 1, 1>, 1, 5 * 32768, 0, 2, 64);
    run("g4w 128r x 256B ring 5, 1 warp x 32 lanes", k_g4w<128, 2, 5, 1, 32>, 1, 5 * 32768, 0, 2, 64);
    run("g4w 128r x 256B ring 5, 2 warps x 1 lane", k_g4w<128, 2, 5, 2, 1>, 1, 5 * 32768, 0, 2, 96);
    run("g4w 128r x 256B ring 5, 4 warps x 1 lane", k_g4w<128, 2, 5, 4, 1>, 1, 5 * 32768, 0, 2, 160);
    run("g4w 128r x 256B ring 5, 4 warps x 4 lanes", k_g4w<128, 2, 5, 4, 4>, 1, 5 * 32768, 0, 2, 160);
    run("g4w 128r x 256B ring 5, 8 warps x 1 lane", k_g4w<128, 2, 5, 8, 1>, 1, 5 * 32768, 0, 2, 288);
    run("g4w 128r x 256B ring 5, 8 warps x 8 lanes", k_g4w<128, 2, 5, 8, 8>, 1, 5 * 32768, 0, 2, 288);
    run("g4w 128r x 256B ring 5, 16 warps x 1 lane", k_g4w<128, 2, 5, 16, 1>, 1, 5 * 32768, 0, 2, 544);
    run("g4w 128r x 256B ring 5, 16 warps x 4 lanes", k_g4w<128, 2, 5, 16, 4>, 1, 5 * 32768, 0, 2, 544);
    run("g4w 128r x 512B ring 3, 8 warps x 1 lane", k_g4w<128, 4, 3, 8, 1>, 1, 3 * 65536, 0, 4, 288);
    run("g4w 128r x 512B ring 3, 16 warps x 1 lane", k_g4w<128, 4, 3, 16, 1>, 1, 3 * 65536, 0, 4, 544);
    run("g4w 32r x 1024B ring 5, 8 warps x 1 lane", k_g4w<32, 8, 5, 8, 1>, 1, 5 * 32768, 0, 8, 288);
    run("g4w 128r x 256B ring 6, 8 warps x 1 lane", k_g4w<128, 2, 6, 8, 1>, 1, 6 * 32768, 0, 2, 288);
  }
  return 0;
}
