// cp.async gather into a shared-memory ring, no consumer: how fast can the ring structure of gather_proj.cu fetch
// random table rows?  Variants: stage geometry (rows x bytes per row), ring depth, threads, CTAs per SM.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o ring_probe ring_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ void cp16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// tile = 128 rows; a stage = ROWS rows x SEG bytes; thread mapping: SEG/16 consecutive lanes on one row
template <int ROWS, int SEG, int RING, int THREADS, bool SYNC>
__global__ void __launch_bounds__(THREADS) k_ring(const unsigned char* __restrict__ tab, const int* __restrict__ ids, int T, int rowb) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ int rid[128];
  constexpr int STAGE = ROWS * SEG;
  constexpr int PER = STAGE / 16 / THREADS;  // copies per thread and stage
  constexpr int LPR = SEG / 16;              // lanes per row
  const int t = threadIdx.x;
  const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem);
  int g = 0;
  for (int tile = blockIdx.x; tile * 128 < T; tile += gridDim.x) {
    __syncthreads();
    if (t < 128) rid[t] = ids[min(tile * 128 + t, T - 1)];
    __syncthreads();
    for (int r0 = 0; r0 < 128; r0 += ROWS) {
      for (int c0 = 0; c0 < rowb; c0 += SEG, ++g) {
        const uint32_t dst = sbase + (g % RING) * STAGE;
#pragma unroll
        for (int i = 0; i < PER; ++i) {
          const int flat = i * THREADS + t;
          const int r = flat / LPR, c = flat % LPR;
          cp16(dst + flat * 16, tab + (size_t)rid[r0 + r] * rowb + c0 + c * 16);
        }
        commit();
        wait<RING - 1>();
        if (SYNC) __syncthreads();   // a stage is usable only when every thread's copies have landed
      }
    }
  }
  wait<0>();
}

int main(int argc, char** argv) {
  const int N = 1000002, T = 294912;
  const int rowb = argc > 1 ? atoi(argv[1]) : 3072;
  unsigned char* tab; int* ids;
  cudaMalloc(&tab, (size_t)N * rowb); cudaMemset(tab, 1, (size_t)N * rowb);
  cudaMalloc(&ids, T * 4);
  std::vector<int> h(T); srand(1); for (auto& x : h) x = (int)(((unsigned)rand() * 32768u + (unsigned)rand()) % N);
  cudaMemcpy(ids, h.data(), T * 4, cudaMemcpyHostToDevice);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto run = [&](const char* name, auto kern, int ctas_per_sm, int threads, int smem) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int i = 0; i < 2; ++i) kern<<<148 * ctas_per_sm, threads, smem>>>(tab, ids, T, rowb);
    cudaEventRecord(e0); for (int i = 0; i < 10; ++i) kern<<<148 * ctas_per_sm, threads, smem>>>(tab, ids, T, rowb);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 10;
    printf("%-52s %8.1f us  %7.1f GB/s  (%s)\n", name, ms * 1e3, (double)T * rowb / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
  };
  run("128r x 256B, ring 5 (160K), 128 thr, 1 CTA, sync", k_ring<128, 256, 5, 128, true>, 1, 128, 5 * 32768);
  run("128r x 256B, ring 5 (160K), 128 thr, 1 CTA, nosync", k_ring<128, 256, 5, 128, false>, 1, 128, 5 * 32768);
  run("128r x 256B, ring 6 (192K), 128 thr, 1 CTA, sync", k_ring<128, 256, 6, 128, true>, 1, 128, 6 * 32768);
  run("128r x 256B, ring 5 (160K), 256 thr, 1 CTA, sync", k_ring<128, 256, 5, 256, true>, 1, 256, 5 * 32768);
  run("128r x 256B, ring 3 (96K), 128 thr, 2 CTA, sync", k_ring<128, 256, 3, 128, true>, 2, 128, 3 * 32768);
  run("128r x 256B, ring 2 (64K), 128 thr, 3 CTA, sync", k_ring<128, 256, 2, 128, true>, 3, 128, 2 * 32768);
  run("128r x 128B, ring 3 (48K), 128 thr, 4 CTA, sync", k_ring<128, 128, 3, 128, true>, 4, 128, 3 * 16384);
  run("128r x 128B, ring 6 (96K), 128 thr, 2 CTA, sync", k_ring<128, 128, 6, 128, true>, 2, 128, 6 * 16384);
  run("64r x 512B, ring 5 (160K), 128 thr, 1 CTA, sync", k_ring<64, 512, 5, 128, true>, 1, 128, 5 * 32768);
  run("32r x 1024B, ring 5 (160K), 128 thr, 1 CTA, sync", k_ring<32, 1024, 5, 128, true>, 1, 128, 5 * 32768);
  run("32r x 1024B, ring 3 (96K), 128 thr, 2 CTA, sync", k_ring<32, 1024, 3, 128, true>, 2, 128, 3 * 32768);
  run("16r x 1024B, ring 10 (160K), 128 thr, 1 CTA, sync", k_ring<16, 1024, 10, 128, true>, 1, 128, 10 * 16384);
  run("16r x 1024B, ring 6 (96K), 128 thr, 2 CTA, sync", k_ring<16, 1024, 6, 128, true>, 2, 128, 6 * 16384);
  run("128r x 256B, ring 7 (224K), 128 thr, 1 CTA, sync", k_ring<128, 256, 7, 128, true>, 1, 128, 7 * 32768 - 2048);
  return 0;
}
