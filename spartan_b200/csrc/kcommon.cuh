// spartan_b200 — device helpers shared by the kernel translation units (kernels.cu, kernels_sc.cu).
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <stdexcept>
#include <string>
#include <vector>
#include "dev.hpp"

namespace sp {
namespace dev {

extern std::atomic<unsigned long long> g_launches;
#define SP_LAUNCHED() (g_launches.fetch_add(1, std::memory_order_relaxed))

// ---- per-kernel-family CUDA-event profiler (bench.py roofline leg).  Off by default: one branch per wrapper.
struct ProfRec { const char* name; cudaEvent_t a, b; double bytes; };
extern bool g_prof;
extern std::vector<ProfRec> g_recs;
cudaEvent_t prof_event();
struct ProfScope {
  bool on; size_t idx; cudaStream_t s;
  ProfScope(const char* name, double bytes, cudaStream_t st) : on(g_prof), idx(0), s(st) {
    if (!on) return;
    ProfRec r{name, prof_event(), prof_event(), bytes};
    cudaEventRecord(r.a, s);
    idx = g_recs.size();
    g_recs.push_back(r);
  }
  ~ProfScope() { if (on) cudaEventRecord(g_recs[idx].b, s); }
};

unsigned int grid_for(size_t work, int threads, int per_sm);

__device__ __forceinline__ u256 ld256(const u256* p) {  // two 128-bit loads
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 a = q[0], b = q[1];
  u256 r;
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w; r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}
__device__ __forceinline__ u256 ld256_ro(const u256* p) {  // read-only path for data never written by the kernel
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 a = __ldg(q), b = __ldg(q + 1);
  u256 r;
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w; r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}
__device__ __forceinline__ u256 ld256_cg(const u256* p) {  // L2-coherent loads for cross-block partial sums
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 a = __ldcg(q), b = __ldcg(q + 1);
  u256 r;
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w; r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}
__device__ __forceinline__ void st256(u256* p, const u256& x) {
  uint4* q = reinterpret_cast<uint4*>(p);
  q[0] = make_uint4(x.v[0], x.v[1], x.v[2], x.v[3]);
  q[1] = make_uint4(x.v[4], x.v[5], x.v[6], x.v[7]);
}
__device__ __forceinline__ u256 shfl_down_256(const u256& x, int delta) {
  u256 r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = __shfl_down_sync(0xffffffffu, x.v[i], delta);
  return r;
}
__device__ __forceinline__ u256 warp_sum_fq(u256 x) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) x = fq_add(x, shfl_down_256(x, d));
  return x;
}

// ---- system-scope flag accesses and the cross-rank completion of a reduction (dev.hpp: XRank)
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// spin until *flag >= seq (sequence numbers only grow); a peer that never arrives traps the kernel after 30 s instead of hanging the GPU
__device__ __forceinline__ void wait_flag_sys(const unsigned int* flag, unsigned int seq) {
  const unsigned long long t0 = global_timer_ns();
  unsigned int spins = 0;
  while ((int)(ld_acquire_sys(flag) - seq) < 0) {
    if ((++spins & 0x3ff) == 0 && global_timer_ns() - t0 > 30000000000ull) __trap();
  }
}
// One warp (all 32 lanes): out[0..nvals) holds this rank's totals; on return it holds the sum over all ranks (also stored to host_out).
__device__ __forceinline__ void xrank_exchange(const XRank& xr, u256* out, int nvals, u256* host_out) {
  const int lane = threadIdx.x & 31;
  const unsigned int slot = xr.seq % SP_XR_SLOTS;
  for (int v = lane; v < nvals; v += 32) {
    const u256 x = ld256_cg(out + v);
    for (int p = 0; p < xr.world; p++)
      if (p != xr.rank) st256(&xr.win[p]->mbox[slot][xr.rank][v], x);          // NVLink store into the peer's mailbox
  }
  __threadfence_system();
  __syncwarp();
  if (lane < xr.world && lane != xr.rank) {
    st_release_sys(&xr.win[lane]->xflag[xr.rank], xr.seq);                     // tell peer `lane` that message `seq` of this rank has landed
    wait_flag_sys(&xr.win[xr.rank]->xflag[lane], xr.seq);                      // and wait for its message
  }
  __syncwarp();
  for (int v = lane; v < nvals; v += 32) {
    u256 s = fq_zero();
    for (int p = 0; p < xr.world; p++)                                          // rank order; exact arithmetic: identical bytes on every rank
      s = fq_add(s, p == xr.rank ? ld256_cg(out + v) : ld256_cg(&xr.win[xr.rank]->mbox[slot][p][v]));
    st256(out + v, s);
    if (host_out) st256(host_out + v, s);
  }
}

// Block-wide sum of NV field values per thread, then cross-block finalisation by the last block to arrive.
// partials: [gridDim.y][gridDim.x][NV]; counters: [gridDim.y] zero-initialised, self-resetting.
// `sig` (optional): the finishing block also stores the results into mapped pinned host memory and then publishes `seq` in a host flag
// word, so the host can pick a round's evaluations up by polling instead of a memcpy + stream synchronise.
template <int NV>
__device__ __forceinline__ void block_reduce_finish(u256 (&acc)[NV], u256* partials, unsigned int* counters, u256* out, int out_stride,
                                                    HostSig sig = HostSig(), const XRank& xr = XRank()) {
  __shared__ u256 sm[32][NV];
  __shared__ bool is_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
#pragma unroll
  for (int k = 0; k < NV; k++) {
    u256 s = warp_sum_fq(acc[k]);
    if (lane == 0) sm[warp][k] = s;
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int k = 0; k < NV; k++) {
      u256 s = lane < nwarps ? sm[lane][k] : fq_zero();
      s = warp_sum_fq(s);
      if (lane == 0) st256(&partials[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * NV + k], s);
    }
  }
  if (threadIdx.x == 0) {
    __threadfence();
    unsigned int ticket = atomicAdd(&counters[blockIdx.y], 1u);
    is_last = (ticket == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    if (warp == 0) {
      const bool xr_on = xr.world > 1;       // sharded: the host gets the sum over all ranks, not this rank's totals
#pragma unroll
      for (int k = 0; k < NV; k++) {
        u256 s = fq_zero();
        for (unsigned int b = lane; b < gridDim.x; b += 32) s = fq_add(s, ld256_cg(&partials[((size_t)blockIdx.y * gridDim.x + b) * NV + k]));
        s = warp_sum_fq(s);
        if (lane == 0) {
          st256(&out[(size_t)blockIdx.y * out_stride + k], s);
          if (sig.host_out && !xr_on) st256(&sig.host_out[(size_t)blockIdx.y * out_stride + k], s);
        }
      }
      unsigned int last_inst = 0;
      if (lane == 0) {
        counters[blockIdx.y] = 0;
        if (sig.done) {                                          // set whenever somebody waits for ALL instances: the host (flag) or the peers (xr)
          __threadfence_system();
          unsigned int done = atomicAdd(sig.done, 1u) + 1;       // instances (blockIdx.y) finish independently
          last_inst = done == gridDim.y;
          if (last_inst) *sig.done = 0;
        }
      }
      last_inst = __shfl_sync(0xffffffffu, last_inst, 0);
      if (last_inst) {
        if (xr_on) { __threadfence(); xrank_exchange(xr, out, (int)(gridDim.y * out_stride), sig.host_out); }
        if (lane == 0 && sig.flag) { __threadfence_system(); *((volatile unsigned int*)sig.flag) = sig.seq; }
      }
    }
  }
}

}  // namespace dev
}  // namespace sp
