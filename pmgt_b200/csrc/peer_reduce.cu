// Data-parallel gradient exchange without a collective library call: the flat fp32 gradient of every rank lives in
// symmetric memory (torch.distributed._symmetric_memory: the same virtual layout on every GPU of the NVSwitch domain,
// peers mapped into this process), and ONE small kernel per rank sums it in place over NVLink --
//   rank r owns slice r of the vector: it loads that slice from every rank's copy (peer loads), adds, and stores the
//   sum back into every rank's copy (peer stores);
// a two-shot all-reduce (reduce-scatter + all-gather) in which no element is read by one rank while another writes
// it, so it runs in place between two device-side barriers of the symmetric-memory handle.  This replaces the two
// ncclAllReduce calls of the implicit DDP of the reference (pmgt/base_trainer.py:309-322): for the 4.75 MB gradient of the
// default model NCCL's channels (one CTA each, with their own shared memory) displaced the persistent main-stream
// kernels for longer than the exchange itself takes (8.3 MB per rank over NVLink).
#include "common.cuh"

namespace pmgt {

namespace {

struct PeerPtrs {
  float* p[8];
};

__device__ __forceinline__ float4 ld_sys(const float4* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_sys(float4* p, const float4& v) {
  asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__global__ void __launch_bounds__(256) peer_reduce_kernel(const PeerPtrs ptrs, int ws, int rank, long long n4, long long n) {
  const long long lo = n4 * rank / ws, hi = n4 * (rank + 1) / ws;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = lo + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += stride) {
    float4 acc = ld_sys(reinterpret_cast<const float4*>(ptrs.p[0]) + i);
    for (int r = 1; r < ws; ++r) {
      const float4 v = ld_sys(reinterpret_cast<const float4*>(ptrs.p[r]) + i);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    for (int r = 0; r < ws; ++r) st_sys(reinterpret_cast<float4*>(ptrs.p[r]) + i, acc);
  }
  // the (at most three) elements beyond the last whole float4: rank 0
  if (rank == 0 && blockIdx.x == 0 && threadIdx.x < (int)(n - 4 * n4)) {
    const long long i = 4 * n4 + threadIdx.x;
    float acc = 0.f;
    for (int r = 0; r < ws; ++r) acc += *reinterpret_cast<volatile float*>(ptrs.p[r] + i);
    for (int r = 0; r < ws; ++r) *reinterpret_cast<volatile float*>(ptrs.p[r] + i) = acc;
  }
}

}  // namespace

}  // namespace pmgt

using namespace pmgt;

extern "C" int pmgt_peer_reduce_f32(const uint64_t* peer_ptrs, int world_size, int rank, int64_t n, void* stream) {
  PMGT_REQUIRE(peer_ptrs && world_size >= 1 && world_size <= 8 && rank >= 0 && rank < world_size && n >= 0,
               "pmgt_peer_reduce_f32: bad argument (world_size %d, rank %d)", world_size, rank);
  if (n == 0 || world_size == 1) return PMGT_OK;
  PeerPtrs pp;
  for (int r = 0; r < 8; ++r) pp.p[r] = r < world_size ? reinterpret_cast<float*>((uintptr_t)peer_ptrs[r]) : nullptr;
  for (int r = 0; r < world_size; ++r)
    PMGT_REQUIRE(pp.p[r] != nullptr && ((uintptr_t)pp.p[r] & 15) == 0, "pmgt_peer_reduce_f32: peer buffer %d is null or not 16-byte aligned", r);
  const long long n4 = n / 4;
  const long long mine = n4 / world_size + 1;
  long long blocks = (mine + 255) / 256;
  if (blocks > 4ll * num_sms()) blocks = 4ll * num_sms();
  if (blocks < 1) blocks = 1;
  peer_reduce_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(pp, world_size, rank, n4, (long long)n);
  PMGT_LAUNCH_CHECK();
  return PMGT_OK;
}
