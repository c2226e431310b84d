// Item-graph ingestion on the device ("next" row (f)4 of SURVEY section 8: pmgt/pmgt/trainer.py:34-41 builds the
// adjacency through networkx; notebooks/PMGT.ipynb cell 20 defines the edge weights).
//
// pmgt_graph_create_device takes a CSR whose rows are already in adjacency-insertion order (the host-side caller sorts
// the doubled edge list by (row, sequence number) with a stable device sort) plus the fp64 edge weights and builds,
// entirely on the GPU, what pmgt_graph_create builds on the host:
//   cdf    per-row running softmax CDF, fp64 math like scipy.special.softmax + numpy's legacy choice()
//          (pmgt/pmgt/datasets.py:27-32: p = softmax(w); cdf = p.cumsum(); cdf /= cdf[-1]), stored fp32, last entry 1
//   ec     (cdf bits, neighbour id) interleaved   }  the guide-table inverse-CDF structures of sampler.cu
//   guide  #{i : cdf[i] <= j / deg} per entry     }
// One warp per row, rows dealt round-robin to the warps of a persistent grid.
#include <vector>

#include "common.cuh"

namespace pmgt {

struct pmgt_graph_impl {  // keep in sync with sampler.cu
  int device;
  int64_t num_nodes;
  int64_t num_edges;
  int64_t* indptr;
  int32_t* indices;
  float* cdf;
  uint2* ec;
  uint16_t* guide;
};

__device__ __forceinline__ double warp_max_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(256) row_cdf_kernel(const int64_t* __restrict__ indptr, const double* __restrict__ w,
                                                      int64_t n_rows, float* __restrict__ cdf, int* __restrict__ max_deg) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t n_warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  int local_max = 0;
  for (int64_t row = warp; row < n_rows; row += n_warps) {
    const int64_t rs = indptr[row];
    const int m = (int)(indptr[row + 1] - rs);
    if (m <= 0) continue;
    local_max = m > local_max ? m : local_max;
    double mx = -INFINITY;
    for (int i = lane; i < m; i += 32) mx = fmax(mx, w[rs + i]);
    mx = warp_max_d(mx);
    double tot = 0.0;
    for (int i = lane; i < m; i += 32) tot += exp(w[rs + i] - mx);
    tot = warp_sum_d(tot);
    double carry = 0.0;
    for (int base = 0; base < m; base += 32) {
      const int i = base + lane;
      double e = i < m ? exp(w[rs + i] - mx) : 0.0;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {  // inclusive scan of the 32 entries of this chunk
        const double up = __shfl_up_sync(0xffffffffu, e, o);
        if (lane >= o) e += up;
      }
      e += carry;
      carry = __shfl_sync(0xffffffffu, e, 31);
      if (i < m) cdf[rs + i] = i == m - 1 ? 1.0f : (float)(e / tot);
    }
  }
  if (local_max > 0 && lane == 0) atomicMax(max_deg, local_max);
}

__global__ void __launch_bounds__(256) ec_guide_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                                                       const float* __restrict__ cdf, int64_t n_rows, uint2* __restrict__ ec,
                                                       uint16_t* __restrict__ guide) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t n_warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t row = warp; row < n_rows; row += n_warps) {
    const int64_t rs = indptr[row];
    const int m = (int)(indptr[row + 1] - rs);
    for (int j = lane; j < m; j += 32) {
      ec[rs + j] = make_uint2(__float_as_uint(cdf[rs + j]), (uint32_t)indices[rs + j]);
      if (guide != nullptr) {
        const double thr = (double)j / (double)m;
        int lo = 0, hi = m;  // number of entries <= thr
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if ((double)cdf[rs + mid] <= thr) lo = mid + 1; else hi = mid;
        }
        guide[rs + j] = (uint16_t)(lo > 65535 ? 65535 : lo);
      }
    }
  }
}

}  // namespace pmgt

using namespace pmgt;

extern "C" {

int pmgt_graph_create_device(pmgt_graph** out, int device, int64_t num_nodes, int64_t num_edges,
                             const int64_t* indptr_dev, const int32_t* indices_dev, const double* weights_dev,
                             void* stream) {
  PMGT_REQUIRE(out && indptr_dev && (num_edges == 0 || (indices_dev && weights_dev)), "pmgt_graph_create_device: null argument");
  PMGT_REQUIRE(num_nodes > 0 && num_edges >= 0 && num_nodes < (int64_t)0x7fffffff - 2,
               "pmgt_graph_create_device: bad sizes (num_nodes=%lld num_edges=%lld)", (long long)num_nodes, (long long)num_edges);
  PMGT_CHECK_CUDA(cudaSetDevice(device));
  cudaStream_t st = (cudaStream_t)stream;
  pmgt_graph_impl* g = new pmgt_graph_impl();
  g->device = device; g->num_nodes = num_nodes; g->num_edges = num_edges;
  g->indptr = nullptr; g->indices = nullptr; g->cdf = nullptr; g->ec = nullptr; g->guide = nullptr;
  const size_t ne = (size_t)(num_edges > 0 ? num_edges : 1);
  int* max_deg = nullptr;
  int max_deg_host = 0;
  cudaError_t e = cudaMalloc(&g->indptr, sizeof(int64_t) * (num_nodes + 3));
  if (e == cudaSuccess) e = cudaMalloc(&g->indices, sizeof(int32_t) * ne);
  if (e == cudaSuccess) e = cudaMalloc(&g->cdf, sizeof(float) * ne);
  if (e == cudaSuccess) e = cudaMalloc(&g->ec, sizeof(uint2) * ne);
  if (e == cudaSuccess) e = cudaMalloc(&g->guide, sizeof(uint16_t) * ne);
  if (e == cudaSuccess) e = cudaMalloc(&max_deg, sizeof(int));
  if (e == cudaSuccess) e = cudaMemsetAsync(max_deg, 0, sizeof(int), st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(g->indptr, indptr_dev, sizeof(int64_t) * (num_nodes + 3), cudaMemcpyDeviceToDevice, st);
  if (e == cudaSuccess && num_edges)
    e = cudaMemcpyAsync(g->indices, indices_dev, sizeof(int32_t) * (size_t)num_edges, cudaMemcpyDeviceToDevice, st);
  if (e == cudaSuccess && num_edges) {
    const int grid = num_sms() * 8;
    row_cdf_kernel<<<grid, 256, 0, st>>>(g->indptr, weights_dev, num_nodes + 2, g->cdf, max_deg);
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(&max_deg_host, max_deg, sizeof(int), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e == cudaSuccess && max_deg_host > 65535) {  // 16-bit guide entries cannot address this row: binary search
      cudaFree(g->guide);
      g->guide = nullptr;
    }
    if (e == cudaSuccess) {
      ec_guide_kernel<<<grid, 256, 0, st>>>(g->indptr, g->indices, g->cdf, num_nodes + 2, g->ec, g->guide);
      e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  }
  cudaFree(max_deg);
  if (e != cudaSuccess) {
    set_error("pmgt_graph_create_device: %s", cudaGetErrorString(e));
    cudaFree(g->indptr); cudaFree(g->indices); cudaFree(g->cdf); cudaFree(g->ec); cudaFree(g->guide);
    delete g;
    return PMGT_ERR_CUDA;
  }
  *out = reinterpret_cast<pmgt_graph*>(g);
  return PMGT_OK;
}

}  // extern "C"
