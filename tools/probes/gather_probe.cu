// DRAM efficiency of a random row gather by request pattern (rows of ROWB bytes from a table much larger than L2).
//   mode 0: a warp fetches SEG contiguous bytes of one row per instruction group, all segments of a row back to back
//   mode 1: GEMM-like: a 128-row tile is walked k-block by k-block (128 B of every row per k-block, barrier between)
//   mode 2: like 1 with SEGB bytes of every row per step (SEGB = 256 / 512)
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gather_probe gather_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

__global__ void k_rows(const uint4* __restrict__ tab, const int* __restrict__ ids, int T, int row16, unsigned* sink) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int nw = (gridDim.x * blockDim.x) >> 5;
  unsigned acc = 0;
  for (int t = warp; t < T; t += nw) {
    const uint4* r = tab + (size_t)ids[t] * row16;
    for (int c = lane; c < row16; c += 32) { uint4 v = __ldg(r + c); acc ^= v.x ^ v.y ^ v.z ^ v.w; }
  }
  if (acc == 0x12345u) *sink = acc;
}

// tile walk: 128 rows per tile, per step every row contributes seg16 16-byte chunks; 128 threads
template <int SEG16>
__global__ void k_tile(const uint4* __restrict__ tab, const int* __restrict__ ids, int T, int row16, unsigned* sink) {
  __shared__ int rid[128];
  unsigned acc = 0;
  const int t = threadIdx.x;
  for (int tile = blockIdx.x; tile * 128 < T; tile += gridDim.x) {
    __syncthreads();
    rid[t] = ids[min(tile * 128 + t, T - 1)];
    __syncthreads();
    for (int c0 = 0; c0 < row16; c0 += SEG16) {
      // 128 rows x SEG16 chunks per step: thread -> (row, chunk) with SEG16 consecutive lanes on one row
      uint4 v[SEG16];
#pragma unroll
      for (int i = 0; i < SEG16; ++i) {
        const int flat = i * 128 + t;
        const int r = flat / SEG16, c = flat % SEG16;
        v[i] = __ldg(tab + (size_t)rid[r] * row16 + c0 + c);
      }
#pragma unroll
      for (int i = 0; i < SEG16; ++i) acc ^= v[i].x ^ v[i].y ^ v[i].z ^ v[i].w;
    }
  }
  if (acc == 0x12345u) *sink = acc;
}

// random 32-byte sectors out of a table much larger than L2: the ceiling of a dependent-gather kernel such as the
// neighbour sampler (every access = one sector); ILP independent loads per thread, ids from a cheap hash
template <int ILP>
__global__ void k_sector(const uint2* __restrict__ tab, size_t n_sectors, int iters, unsigned* sink) {
  unsigned x = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
  unsigned acc = 0;
  for (int it = 0; it < iters; ++it) {
    uint2 v[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      x = x * 1664525u + 1013904223u;
      unsigned h = x ^ (x >> 15); h *= 0x2c1b3c6du; h ^= h >> 12;
      v[i] = __ldg(tab + ((size_t)h % n_sectors) * 4);   // first 8 bytes of sector h
    }
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc ^= v[i].x ^ v[i].y;
  }
  if (acc == 0x12345u) *sink = acc;
}

// row-visit probe for the sampler's data layout: every "visit" reads three 32-byte sectors that depend on each other
// (row pointer -> guide entry -> (cdf, id) pair).  SPLIT: the three sectors lie in three different 256 MB arrays (the
// CSR layout: indptr | guide | ec).  RECORD: all three lie inside one 512-byte record of a single array.
template <bool RECORD>
__global__ void k_visit(const uint2* __restrict__ tab, int iters, unsigned* sink) {
  const size_t region = (size_t)1 << 23;   // sectors per 256 MB region
  unsigned x = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 777u;
  unsigned acc = 0;
  for (int it = 0; it < iters; ++it) {
    uint2 v[4];
    size_t s0[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {   // four independent visits in flight per thread
      x = x * 1664525u + 1013904223u;
      unsigned h = x ^ (x >> 15); h *= 0x2c1b3c6du; h ^= h >> 12;
      s0[i] = RECORD ? ((size_t)h % (region * 3 / 16)) * 16 : (size_t)h % region;   // record = 16 sectors
      v[i] = __ldg(tab + s0[i] * 4);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {   // dependent second access
      const unsigned d = (v[i].x ^ s0[i]) & 7u;
      const size_t s1 = RECORD ? s0[i] + 1 + d : region + (s0[i] * 2654435761ull + d) % region;
      v[i] = __ldg(tab + s1 * 4);
      s0[i] = s1;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {   // dependent third access
      const unsigned d = (v[i].y ^ s0[i]) & 3u;
      const size_t s2 = RECORD ? (s0[i] & ~(size_t)15) + 9 + d : 2 * region + (s0[i] * 40503ull + d) % region;
      v[i] = __ldg(tab + s2 * 4);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) acc ^= v[i].x ^ v[i].y;
  }
  if (acc == 0x12345u) *sink = acc;
}

int main(int argc, char** argv) {
  const int N = 1000002, T = 294912;
  const int rowb = argc > 1 ? atoi(argv[1]) : 3072;
  const int row16 = rowb / 16;
  uint4* tab; int* ids; unsigned* sink;
  cudaMalloc(&tab, (size_t)N * rowb); cudaMemset(tab, 1, (size_t)N * rowb);
  cudaMalloc(&ids, T * 4); cudaMalloc(&sink, 4);
  std::vector<int> h(T); srand(1); for (auto& x : h) x = (int)(((unsigned)rand() * 32768u + (unsigned)rand()) % N);
  cudaMemcpy(ids, h.data(), T * 4, cudaMemcpyHostToDevice);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto run = [&](const char* name, auto fn) {
    for (int i = 0; i < 2; ++i) fn();
    cudaEventRecord(e0); for (int i = 0; i < 10; ++i) fn(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 10;
    printf("%-28s rowb=%d  %8.1f us  %7.1f GB/s  (%s)\n", name, rowb, ms * 1e3, (double)T * rowb / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
  };
  run("warp per row, 8 CTA x 256/SM", [&] { k_rows<<<148 * 8, 256>>>(tab, ids, T, row16, sink); });
  run("warp per row, 4 CTA x 256/SM", [&] { k_rows<<<148 * 4, 256>>>(tab, ids, T, row16, sink); });
  run("tile walk 128B/step 4 CTA/SM", [&] { k_tile<8><<<148 * 4, 128>>>(tab, ids, T, row16, sink); });
  run("tile walk 128B/step 8 CTA/SM", [&] { k_tile<8><<<148 * 8, 128>>>(tab, ids, T, row16, sink); });
  run("tile walk 128B/step 16 CTA/SM", [&] { k_tile<8><<<148 * 16, 128>>>(tab, ids, T, row16, sink); });
  run("tile walk 256B/step 4 CTA/SM", [&] { k_tile<16><<<148 * 4, 128>>>(tab, ids, T, row16, sink); });
  run("tile walk 256B/step 8 CTA/SM", [&] { k_tile<16><<<148 * 8, 128>>>(tab, ids, T, row16, sink); });
  run("tile walk 512B/step 2 CTA/SM", [&] { k_tile<32><<<148 * 2, 128>>>(tab, ids, T, row16, sink); });
  run("tile walk 512B/step 4 CTA/SM", [&] { k_tile<32><<<148 * 4, 128>>>(tab, ids, T, row16, sink); });
  if (argc > 2) {   // row-visit probe (768 MB window)
    for (int rec = 0; rec < 2; ++rec) for (int cps : {8, 16}) {
      const int iters = 64, threads = 128;
      const double visits = (double)148 * cps * threads * iters * 4;
      auto go = [&] { if (rec) k_visit<true><<<148 * cps, threads>>>((const uint2*)tab, iters, sink); else k_visit<false><<<148 * cps, threads>>>((const uint2*)tab, iters, sink); };
      go(); go();
      cudaEventRecord(e0); for (int i = 0; i < 5; ++i) go();
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
      printf("row visits of 3 dependent sectors, %s layout, %2d CTA x 128 thr / SM: %6.1f G visits/s (%s)\n", rec ? "RECORD" : "SPLIT ", cps,
             visits / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
    }
  }
  if (argc > 2) {   // sector probe: a 1 GB window of the table
    const size_t n_sectors = (size_t)1 << 25;
    for (int cps : {8, 16, 32}) {
      const int iters = 64, threads = 128, ilp = 8;
      const double n_acc = (double)148 * cps * threads * iters * ilp;
      for (int i = 0; i < 2; ++i) k_sector<8><<<148 * cps, threads>>>((const uint2*)tab, n_sectors, iters, sink);
      cudaEventRecord(e0); for (int i = 0; i < 5; ++i) k_sector<8><<<148 * cps, threads>>>((const uint2*)tab, n_sectors, iters, sink);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
      printf("random 32B sectors, %2d CTA x 128 thr / SM, ILP 8: %6.1f G sectors/s = %7.1f GB/s of 32-byte sectors (%s)\n", cps,
             n_acc / ms / 1e6, n_acc * 32 / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
    }
  }
  return 0;
}
