// spartan_b200 — hand-written sm_100a kernels for the Spartan prover hot path.
//
// Kernel families:
//   K1/K2  sc_eval / sc_fold_eval / fold_top : dense-multilinear sumcheck rounds
//          (reference loops: /root/reference/src/sumcheck.rs:460-469, :204-228, :296-355, :625-652 and
//           src/dense_mlpoly.rs:215-223).  Streaming scans: 128-bit coalesced loads, per-thread field accumulators,
//           warp-shuffle + shared-memory segmented reduction, last-block finalisation (no second launch).
//   K3/K4  msm_rows : Pedersen commitments sum_j s_j*G_j (+ blind*h) for L rows sharing one generator set
//          (reference: dalek vartime_multiscalar_mul behind src/group.rs:98-117, called from src/commitments.rs:80-92 and
//           src/dense_mlpoly.rs:165-177).  Generators are fixed per `Gens`, so the kernel is a fixed-base comb: 8-bit signed
//           windows, 128-entry affine-niels tables per (generator, window) resident in HBM, 7M mixed additions.
//   K6     compress_batch / decompress_batch / gens_from_uniform (RFC 9496).
//   K7     eq_evals, dot, bound_rows, lincomb3, hadamard, spmv, SPARK hash layer, IPA helpers.
// All arithmetic is exact 256-bit integer work on the INT32 pipe; no tensor-core formulation exists.
#include <cuda_runtime.h>
#include <cstdio>
#include <cuda.h>
#include <stdexcept>
#include <string>
#include <map>
#include <mutex>
#include <atomic>
#include <vector>
#include "dev.hpp"
#include "kcommon.cuh"

namespace sp {
namespace dev {

std::atomic<unsigned long long> g_launches{0};
unsigned long long launch_count() { return g_launches.load(); }
static std::atomic<unsigned long long> g_h2d{0}, g_d2h{0};
void io_bytes(unsigned long long* h2d_b, unsigned long long* d2h_b) { *h2d_b = g_h2d.load(); *d2h_b = g_d2h.load(); }

// ---- per-kernel-family CUDA-event profiler (bench.py roofline leg).  Off by default: one branch per wrapper.
bool g_prof = false;
std::vector<ProfRec> g_recs;
static std::vector<cudaEvent_t> g_event_pool;
cudaEvent_t prof_event() {
  if (!g_event_pool.empty()) { cudaEvent_t e = g_event_pool.back(); g_event_pool.pop_back(); return e; }
  cudaEvent_t e; cudaEventCreate(&e); return e;
}
void prof_enable(bool on) {
  g_prof = on;
  for (auto& r : g_recs) { g_event_pool.push_back(r.a); g_event_pool.push_back(r.b); }
  g_recs.clear();
}
// "name:launches:total_ms:total_algorithmic_bytes:largest_launch_bytes:largest_launch_ms;" per family, after a device synchronise
std::string prof_report() {
  cudaDeviceSynchronize();
  struct Agg { double ms = 0, bytes = 0, big_bytes = 0, big_ms = 0; unsigned long long n = 0; };
  std::vector<std::pair<std::string, Agg>> aggs;
  for (auto& r : g_recs) {
    float ms = 0;
    cudaEventElapsedTime(&ms, r.a, r.b);
    size_t k = 0;
    for (; k < aggs.size(); k++) if (aggs[k].first == r.name) break;
    if (k == aggs.size()) aggs.push_back({r.name, Agg()});
    aggs[k].second.ms += ms; aggs[k].second.bytes += r.bytes; aggs[k].second.n++;
    if (r.bytes > aggs[k].second.big_bytes) { aggs[k].second.big_bytes = r.bytes; aggs[k].second.big_ms = ms; }
  }
  std::string out;
  for (auto& a : aggs)
    out += a.first + ":" + std::to_string(a.second.n) + ":" + std::to_string(a.second.ms) + ":" + std::to_string(a.second.bytes) + ":" +
           std::to_string(a.second.big_bytes) + ":" + std::to_string(a.second.big_ms) + ";";
  return out;
}

void check(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) throw std::runtime_error(std::string("spartan_b200 CUDA error in ") + what + ": " + cudaGetErrorString(e));
}
static void ck(cudaError_t e, const char* what) {
  if (e != cudaSuccess) throw std::runtime_error(std::string("spartan_b200 CUDA error in ") + what + ": " + cudaGetErrorString(e));
}
int device_count() { int n = 0; if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; } return n; }
void set_device(int d) { ck(cudaSetDevice(d), "cudaSetDevice"); }
cudaStream_t stream_create() { cudaStream_t s; ck(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking), "cudaStreamCreate"); return s; }
// level > 0: the device's greatest priority (the transcript-serialised stream of small kernels), level < 0: its least (background MSMs)
cudaStream_t stream_create_prio(int level) {
  int least = 0, greatest = 0;
  ck(cudaDeviceGetStreamPriorityRange(&least, &greatest), "cudaDeviceGetStreamPriorityRange");
  cudaStream_t s;
  ck(cudaStreamCreateWithPriority(&s, cudaStreamNonBlocking, level > 0 ? greatest : (level < 0 ? least : (least + greatest) / 2)), "cudaStreamCreateWithPriority");
  return s;
}
// A stream whose kernels run on a fixed subset of `sms` SMs (a CUDA green context carved out of the device's SM resource; driver entry points are
// resolved at run time so that the library has no link-time dependency on libcuda).  The remaining SMs are never touched by work on this stream, so
// the short kernels of the prover's stream find idle SMs however long the background MSMs run.  Returns nullptr when the driver cannot do it.
cudaStream_t stream_create_partition(int sms, int level, int* sms_granted) {
  typedef CUresult (*GetRes)(CUdevice, CUdevResource*, CUdevResourceType);
  typedef CUresult (*Split)(CUdevResource*, unsigned int*, const CUdevResource*, CUdevResource*, unsigned int, unsigned int);
  typedef CUresult (*GenDesc)(CUdevResourceDesc*, CUdevResource*, unsigned int);
  typedef CUresult (*GreenCreate)(CUgreenCtx*, CUdevResourceDesc, CUdevice, unsigned int);
  typedef CUresult (*GreenStream)(CUstream*, CUgreenCtx, unsigned int, int);
  typedef CUresult (*DevGet)(CUdevice*, int);
  auto ep = [](const char* name) -> void* {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult st;
    if (cudaGetDriverEntryPoint(name, &f, cudaEnableDefault, &st) != cudaSuccess || st != cudaDriverEntryPointSuccess) { cudaGetLastError(); return nullptr; }
    return f;
  };
  GetRes get_res = (GetRes)ep("cuDeviceGetDevResource");
  Split split = (Split)ep("cuDevSmResourceSplitByCount");
  GenDesc gen = (GenDesc)ep("cuDevResourceGenerateDesc");
  GreenCreate gcreate = (GreenCreate)ep("cuGreenCtxCreate");
  GreenStream gstream = (GreenStream)ep("cuGreenCtxStreamCreate");
  DevGet devget = (DevGet)ep("cuDeviceGet");
  if (!get_res || !split || !gen || !gcreate || !gstream || !devget) return nullptr;
  int ord = 0;
  cudaGetDevice(&ord);
  cudaFree(0);   // the primary context must exist
  CUdevice dev;
  if (devget(&dev, ord) != CUDA_SUCCESS) return nullptr;
  CUdevResource all, grp, rem;
  if (get_res(dev, &all, CU_DEV_RESOURCE_TYPE_SM) != CUDA_SUCCESS) return nullptr;
  if (sms < 8 || (unsigned)sms >= all.sm.smCount) return nullptr;
  unsigned int nb = 1;
  if (split(&grp, &nb, &all, &rem, 0, (unsigned)sms) != CUDA_SUCCESS || nb < 1) return nullptr;
  CUdevResourceDesc desc;
  if (gen(&desc, &grp, 1) != CUDA_SUCCESS) return nullptr;
  CUgreenCtx g;
  if (gcreate(&g, desc, dev, CU_GREEN_CTX_DEFAULT_STREAM) != CUDA_SUCCESS) return nullptr;   // lives as long as the process (one per prover context)
  int least = 0, greatest = 0;
  cudaDeviceGetStreamPriorityRange(&least, &greatest);
  CUstream st;
  if (gstream(&st, g, CU_STREAM_NON_BLOCKING, level > 0 ? greatest : least) != CUDA_SUCCESS) return nullptr;
  if (sms_granted) *sms_granted = (int)grp.sm.smCount;
  return (cudaStream_t)st;
}
void stream_wait_event(cudaStream_t s, void* ev) { ck(cudaStreamWaitEvent(s, (cudaEvent_t)ev, 0), "cudaStreamWaitEvent"); }
void stream_destroy(cudaStream_t s) { cudaStreamDestroy(s); }
void stream_sync(cudaStream_t s) { ck(cudaStreamSynchronize(s), "cudaStreamSynchronize"); }
void* dmalloc(size_t b) { void* p = nullptr; ck(cudaMalloc(&p, b ? b : 16), "cudaMalloc"); return p; }
void dfree(void* p) { if (p) cudaFree(p); }
namespace {
std::mutex g_pool_mu;
std::multimap<std::pair<int, size_t>, void*> g_pool_free;  // (device, rounded size) -> cached block
std::map<void*, std::pair<int, size_t>> g_pool_live;       // block -> (device, rounded size)
size_t pool_round(size_t b) {
  if (b < 256) b = 256;
  if (b <= (1u << 20)) { size_t r = 256; while (r < b) r <<= 1; return r; }
  const size_t g = 2u << 20;                   // 2 MiB granularity above 1 MiB
  return (b + g - 1) / g * g;
}
}  // namespace
void* pool_alloc(size_t bytes) {
  size_t r = pool_round(bytes);
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lk(g_pool_mu);
  auto it = g_pool_free.find({dev, r});
  void* p;
  if (it != g_pool_free.end()) { p = it->second; g_pool_free.erase(it); }
  else {
    cudaError_t e = cudaMalloc(&p, r);
    if (e != cudaSuccess) {  // out of memory: drop the cache and retry once
      cudaGetLastError();
      for (auto& kv : g_pool_free) cudaFree(kv.second);
      g_pool_free.clear();
      ck(cudaMalloc(&p, r), "cudaMalloc");
    }
  }
  g_pool_live[p] = {dev, r};
  return p;
}
void pool_free(void* p) {
  if (!p) return;
  std::lock_guard<std::mutex> lk(g_pool_mu);
  auto it = g_pool_live.find(p);
  if (it == g_pool_live.end()) { cudaFree(p); return; }
  // blocks of tens of GB (generator tables) are long-lived: do not hoard them.  Per-proof buffers stay cached even when they pass 1 GiB (2^22 constraints: the
  // 1.2 GB dot-product circuit buffer) — a cudaFree synchronises the device, i.e. waits for the background stream's MSM, and the cudaMalloc of the next proof
  // made SNARK::prove at 2^22 bimodal (101 / 200 ms, GPU call 20)
  if (it->second.second > ((size_t)8 << 30)) cudaFree(p);
  else g_pool_free.insert({it->second, p});
  g_pool_live.erase(it);
}
void pool_trim() {
  std::lock_guard<std::mutex> lk(g_pool_mu);
  cudaDeviceSynchronize();
  for (auto& kv : g_pool_free) cudaFree(kv.second);
  g_pool_free.clear();
}
void* hmalloc_pinned(size_t b) { void* p = nullptr; ck(cudaMallocHost(&p, b ? b : 16), "cudaMallocHost"); return p; }
void hfree_pinned(void* p) { if (p) cudaFreeHost(p); }
void h2d(void* d, const void* h, size_t b, cudaStream_t s) { g_h2d += b; if (b) ck(cudaMemcpyAsync(d, h, b, cudaMemcpyHostToDevice, s), "h2d"); }
void d2h(void* h, const void* d, size_t b, cudaStream_t s) { g_d2h += b; if (b) ck(cudaMemcpyAsync(h, d, b, cudaMemcpyDeviceToHost, s), "d2h"); }
void d2d(void* dst, const void* src, size_t b, cudaStream_t s) { if (b) ck(cudaMemcpyAsync(dst, src, b, cudaMemcpyDeviceToDevice, s), "d2d"); }
void dzero(void* d, size_t b, cudaStream_t s) { if (b) ck(cudaMemsetAsync(d, 0, b, s), "memset"); }
int sm_count() {
  static int n = 0;
  if (!n) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev); if (n <= 0) n = 148; }
  return n;
}
void mem_info(size_t* free_bytes, size_t* total_bytes) { ck(cudaMemGetInfo(free_bytes, total_bytes), "cudaMemGetInfo"); }
void* event_create() { cudaEvent_t e; ck(cudaEventCreate(&e), "cudaEventCreate"); return (void*)e; }
void event_record(void* ev, cudaStream_t s) { ck(cudaEventRecord((cudaEvent_t)ev, s), "cudaEventRecord"); }
float event_elapsed_ms(void* a, void* b) {
  float ms = 0;
  ck(cudaEventSynchronize((cudaEvent_t)b), "cudaEventSynchronize");
  ck(cudaEventElapsedTime(&ms, (cudaEvent_t)a, (cudaEvent_t)b), "cudaEventElapsedTime");
  return ms;
}
void event_destroy(void* ev) { cudaEventDestroy((cudaEvent_t)ev); }

unsigned int grid_for(size_t work, int threads, int per_sm) {
  size_t blocks = (work + threads - 1) / threads;
  size_t cap = (size_t)sm_count() * per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (unsigned int)blocks;
}


// =============================================================================================== dense helpers
// eq(r, .) table (dense_mlpoly.rs:68-84): out[i] = prod_j (bit_j(i) ? r_j : 1 - r_j), bit 0 of r = most significant bit of i.
// Two-level: small tables for the high / low halves of the variables, then one product per output element.
__global__ void k_eq_small(u256* out, const u256* __restrict__ r, int nv) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ((size_t)1 << nv)) return;
  u256 acc = fq_one();
  for (int j = 0; j < nv; j++) {
    u256 rj = ld256_ro(r + j);
    bool bit = (i >> (nv - 1 - j)) & 1;
    acc = fq_mul(acc, bit ? rj : fq_sub(fq_one(), rj));
  }
  st256(out + i, acc);
}
__global__ void k_eq_combine(u256* out, const u256* __restrict__ hi, const u256* __restrict__ lo, int nlo, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    st256(out + i, fq_mul(ld256_ro(hi + (i >> nlo)), ld256_ro(lo + (i & (((size_t)1 << nlo) - 1)))));
}
void eq_evals(u256* out, const u256* d_r, int ell, u256* small, cudaStream_t s) {
  ProfScope ps("eq_evals", 32.0 * (double)((size_t)1 << ell), s);
  if (ell <= 10) {
    size_t n = (size_t)1 << ell;
    k_eq_small<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(out, d_r, ell);
    SP_LAUNCHED(); check("eq_small");
    return;
  }
  int nhi = ell / 2, nlo = ell - nhi;
  u256* thi = small; u256* tlo = small + ((size_t)1 << nhi);
  k_eq_small<<<(unsigned)((((size_t)1 << nhi) + 127) / 128), 128, 0, s>>>(thi, d_r, nhi);
  k_eq_small<<<(unsigned)((((size_t)1 << nlo) + 127) / 128), 128, 0, s>>>(tlo, d_r + nhi, nlo);
  size_t n = (size_t)1 << ell;
  k_eq_combine<<<grid_for(n, 256, 8), 256, 0, s>>>(out, thi, tlo, nlo, n);
  SP_LAUNCHED(); SP_LAUNCHED(); SP_LAUNCHED(); check("eq_evals");
}

__global__ void __launch_bounds__(256) k_dot(const u256* __restrict__ a, const u256* __restrict__ b, const u256* __restrict__ c, size_t n,
                                            u256* partials, unsigned int* counters, u256* out) {
  u256 acc[1] = {fq_zero()};
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    u256 m = fq_mul(ld256_ro(a + i), ld256_ro(b + i));
    if (c) m = fq_mul(m, ld256_ro(c + i));
    acc[0] = fq_add(acc[0], m);
  }
  block_reduce_finish<1>(acc, partials, counters, out, 1);
}
void dot3(u256* out, const u256* a, const u256* b, const u256* c, size_t n, void* scratch, cudaStream_t s) {
  ProfScope ps("dot", (c ? 96.0 : 64.0) * (double)n, s);
  unsigned int* counters = (unsigned int*)scratch;
  u256* partials = (u256*)((char*)scratch + 256);
  dim3 grid(grid_for(n, 256, 2), 1);
  k_dot<<<grid, 256, 0, s>>>(a, b, c, n, partials, counters, out);
  SP_LAUNCHED(); check("dot");
}
void dot(u256* out, const u256* a, const u256* b, size_t n, void* scratch, cudaStream_t s) { dot3(out, a, b, nullptr, n, scratch, s); }

// out[k] = <a_k, b> for up to 32 tables sharing b (the 21 + 2 evaluations of HashLayerProof::prove, sparse_mlpoly.rs:696-764, all at one point)
struct DotBatch { const u256* a[32]; };
__global__ void __launch_bounds__(256) k_dot_many(DotBatch batch, const u256* __restrict__ b, size_t n, u256* partials, unsigned int* counters, u256* out) {
  const u256* a = batch.a[blockIdx.y];
  u256 acc[1] = {fq_zero()};
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    acc[0] = fq_add(acc[0], fq_mul(ld256_ro(a + i), ld256_ro(b + i)));
  block_reduce_finish<1>(acc, partials, counters, out, 1);
}
void dot_many(u256* out, const u256* const* a_list, int count, const u256* b, size_t n, void* scratch, cudaStream_t s) {
  ProfScope ps("dot_many", 32.0 * (double)n * (count + 1), s);
  if (count > 32) throw std::runtime_error("spartan_b200: dot_many supports at most 32 tables");
  DotBatch db;
  for (int i = 0; i < count; i++) db.a[i] = a_list[i];
  unsigned int* counters = (unsigned int*)scratch;
  u256* partials = (u256*)((char*)scratch + 256);
  dim3 grid(grid_for(n, 256, 1), count);
  k_dot_many<<<grid, 256, 0, s>>>(db, b, n, partials, counters, out);
  SP_LAUNCHED(); check("dot_many");
}

// out[k] = <a_k, b_k> for up to 8 independent pairs in one launch; the results can also be published to the host (HostSig)
struct DotPairs { const u256* a[8]; const u256* b[8]; };
__global__ void __launch_bounds__(256) k_dot_pairs(DotPairs dp, size_t n, u256* partials, unsigned int* counters, u256* out, HostSig sig) {
  const u256* a = dp.a[blockIdx.y];
  const u256* b = dp.b[blockIdx.y];
  u256 acc[1] = {fq_zero()};
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    acc[0] = fq_add(acc[0], fq_mul(ld256_ro(a + i), ld256_ro(b + i)));
  block_reduce_finish<1>(acc, partials, counters, out, 1, sig);
}
void dot_pairs(u256* out, const u256* const* a_list, const u256* const* b_list, int count, size_t n, void* scratch, cudaStream_t s, HostSig sig) {
  ProfScope ps("dot", 64.0 * (double)n * count, s);
  if (count > 8) throw std::runtime_error("spartan_b200: dot_pairs supports at most 8 pairs");
  DotPairs dp;
  for (int i = 0; i < count; i++) { dp.a[i] = a_list[i]; dp.b[i] = b_list[i]; }
  unsigned int* counters = (unsigned int*)scratch;
  u256* partials = (u256*)((char*)scratch + 256);
  dim3 grid(grid_for(n, 256, 1), count);
  k_dot_pairs<<<grid, 256, 0, s>>>(dp, n, partials, counters, out, sig);
  SP_LAUNCHED(); check("dot_pairs");
}

// first elements of up to 64 tables -> one contiguous array (+ the host, through HostSig): the layer claims of the batched product proofs
struct HeadBatch { const u256* p[64]; };
__global__ void __launch_bounds__(64) k_heads(HeadBatch hb, int count, u256* out, HostSig sig) {
  if ((int)threadIdx.x < count) {
    u256 v = ld256(hb.p[threadIdx.x]);
    st256(out + threadIdx.x, v);
    if (sig.host_out) st256(sig.host_out + threadIdx.x, v);
  }
  if (sig.flag) {
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) { __threadfence_system(); *((volatile unsigned int*)sig.flag) = sig.seq; }
  }
}
void heads(u256* out, const u256* const* tables, int count, cudaStream_t s, HostSig sig) {
  if (count > 64) throw std::runtime_error("spartan_b200: heads supports at most 64 tables");
  HeadBatch hb;
  for (int i = 0; i < count; i++) hb.p[i] = tables[i];
  k_heads<<<1, 64, 0, s>>>(hb, count, out, sig);
  SP_LAUNCHED(); check("heads");
}

// DensePolynomial::bound (dense_mlpoly.rs:206-213): out[i] = sum_j L[j]*Z[j*R+i].  Column-per-thread (coalesced over i),
// rows split into gridDim.y slabs whose partial sums land in scratch and are combined by a second tiny kernel.
__global__ void __launch_bounds__(128) k_bound_rows_partial(u256* part, const u256* __restrict__ Z, const u256* __restrict__ L, size_t L_size,
                                                           size_t R_size, size_t rows_per_slab) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R_size) return;
  size_t j0 = (size_t)blockIdx.y * rows_per_slab, j1 = j0 + rows_per_slab;
  if (j1 > L_size) j1 = L_size;
  u256 acc = fq_zero();
  for (size_t j = j0; j < j1; j++) acc = fq_add(acc, fq_mul(ld256_ro(L + j), ld256_ro(Z + j * R_size + i)));
  st256(part + (size_t)blockIdx.y * R_size + i, acc);
}
__global__ void k_sum_slabs(u256* out, const u256* __restrict__ part, size_t R_size, int nslabs) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R_size) return;
  u256 acc = fq_zero();
  for (int k = 0; k < nslabs; k++) acc = fq_add(acc, ld256_ro(part + (size_t)k * R_size + i));
  st256(out + i, acc);
}
void bound_rows(u256* out, const u256* Z, const u256* L, size_t L_size, size_t R_size, u256* scratch, cudaStream_t s) {
  ProfScope ps("bound_rows", 32.0 * (double)L_size * (double)R_size, s);
  int nslabs = 64;
  while (nslabs > 1 && (size_t)nslabs > L_size) nslabs >>= 1;
  size_t rows_per_slab = (L_size + nslabs - 1) / nslabs;
  dim3 grid((unsigned)((R_size + 127) / 128), nslabs);
  k_bound_rows_partial<<<grid, 128, 0, s>>>(scratch, Z, L, L_size, R_size, rows_per_slab);
  k_sum_slabs<<<(unsigned)((R_size + 127) / 128), 128, 0, s>>>(out, scratch, R_size, nslabs);
  SP_LAUNCHED(); SP_LAUNCHED(); check("bound_rows");
}

__global__ void k_lincomb3(u256* out, const u256* __restrict__ A, const u256* __restrict__ B, const u256* __restrict__ C,
                           const u256* __restrict__ rabc, size_t n) {
  u256 ra = ld256_ro(rabc), rb = ld256_ro(rabc + 1), rc = ld256_ro(rabc + 2);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    st256(out + i, fq_add(fq_add(fq_mul(ra, ld256_ro(A + i)), fq_mul(rb, ld256_ro(B + i))), fq_mul(rc, ld256_ro(C + i))));
}
void lincomb3(u256* out, const u256* A, const u256* B, const u256* C, const u256* d_rabc, size_t n, cudaStream_t s) {
  ProfScope ps("lincomb3", 128.0 * (double)n, s);
  k_lincomb3<<<grid_for(n, 256, 4), 256, 0, s>>>(out, A, B, C, d_rabc, n);
  SP_LAUNCHED(); check("lincomb3");
}
__global__ void k_hadamard(u256* out, const u256* a, const u256* b, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    st256(out + i, fq_mul(ld256(a + i), ld256(b + i)));
}
void hadamard(u256* out, const u256* a, const u256* b, size_t n, cudaStream_t s) {
  ProfScope ps("hadamard", 96.0 * (double)n, s);
  k_hadamard<<<grid_for(n, 256, 8), 256, 0, s>>>(out, a, b, n);
  SP_LAUNCHED(); check("hadamard");
}
// one product-tree layer of up to 16 circuits of equal size in one launch (blockIdx.y = circuit)
struct HadBatch { u256* out[16]; const u256* a[16]; const u256* b[16]; };
__global__ void k_hadamard_many(HadBatch hb, size_t n) {
  u256* out = hb.out[blockIdx.y]; const u256* a = hb.a[blockIdx.y]; const u256* b = hb.b[blockIdx.y];
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    st256(out + i, fq_mul(ld256(a + i), ld256(b + i)));
}
void hadamard_many(u256* const* outs, const u256* const* as, const u256* const* bs, int count, size_t n, cudaStream_t s) {
  ProfScope ps("hadamard", 96.0 * (double)n * count, s);
  if (count > 16) throw std::runtime_error("spartan_b200: hadamard_many supports at most 16 tables");
  HadBatch hb;
  for (int i = 0; i < count; i++) { hb.out[i] = outs[i]; hb.a[i] = as[i]; hb.b[i] = bs[i]; }
  dim3 grid(grid_for(n, 256, 4), count);
  k_hadamard_many<<<grid, 256, 0, s>>>(hb, n);
  SP_LAUNCHED(); check("hadamard_many");
}
__global__ void k_from_u64(u256* out, const uint64_t* __restrict__ v, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) st256(out + i, fq_from_u64(v[i]));
}
void from_u64(u256* out, const uint64_t* v, size_t n, cudaStream_t s) {
  k_from_u64<<<grid_for(n, 256, 8), 256, 0, s>>>(out, v, n);
  SP_LAUNCHED(); check("from_u64");
}
__global__ void k_from_wide(u256* out, const uint8_t* __restrict__ in, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const u256* p = reinterpret_cast<const u256*>(in + 64 * i);
    st256(out + i, fq_from_wide(ld256_ro(p), ld256_ro(p + 1)));
  }
}
void from_bytes_wide(u256* out, const uint8_t* in64, size_t n, cudaStream_t s) {
  k_from_wide<<<grid_for(n, 256, 8), 256, 0, s>>>(out, in64, n);
  SP_LAUNCHED(); check("from_bytes_wide");
}
__global__ void k_to_canonical(u256* out, const u256* __restrict__ in, size_t n) {   // Scalar::to_bytes (ristretto255.rs:419) for a whole vector
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) st256(out + i, fq_from_mont(ld256_ro(in + i)));
}
void to_canonical(u256* out, const u256* in, size_t n, cudaStream_t s) {
  k_to_canonical<<<grid_for(n, 128, 8), 128, 0, s>>>(out, in, n);
  SP_LAUNCHED(); check("to_canonical");
}
__global__ void k_invert(u256* x, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) st256(x + i, fq_inv(ld256(x + i)));
}
void batch_invert_elems(u256* x, size_t n, cudaStream_t s) {
  k_invert<<<grid_for(n, 128, 16), 128, 0, s>>>(x, n);
  SP_LAUNCHED(); check("invert");
}
__global__ void k_gather(u256* out, const u256* __restrict__ mem, const uint32_t* __restrict__ idx, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) st256(out + i, ld256_ro(mem + idx[i]));
}
void gather(u256* out, const u256* mem, const uint32_t* idx, size_t n, cudaStream_t s) {
  ProfScope ps("gather", 68.0 * (double)n, s);
  k_gather<<<grid_for(n, 256, 8), 256, 0, s>>>(out, mem, idx, n);
  SP_LAUNCHED(); check("gather");
}
__global__ void k_spark_hash(u256* out, size_t n, const u256* __restrict__ addr, const u256* __restrict__ val, const u256* __restrict__ ts,
                             int ts_plus_one, const u256* __restrict__ rg) {
  u256 r = ld256_ro(rg), g = ld256_ro(rg + 1), r2 = fq_sqr(r);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    u256 a = addr ? ld256_ro(addr + i) : fq_from_u64((uint64_t)i);
    u256 t = ts ? ld256_ro(ts + i) : fq_zero();
    if (ts_plus_one) t = fq_add(t, fq_one());
    u256 h = fq_add(fq_add(fq_mul(t, r2), fq_mul(ld256_ro(val + i), r)), a);
    st256(out + i, fq_sub(h, g));
  }
}
void spark_hash(u256* out, size_t n, const u256* addr, const u256* val, const u256* ts, int ts_plus_one, const u256* d_rg, cudaStream_t s) {
  ProfScope ps("spark_hash", 128.0 * (double)n, s);
  k_spark_hash<<<grid_for(n, 256, 4), 256, 0, s>>>(out, n, addr, val, ts, ts_plus_one, d_rg);
  SP_LAUNCHED(); check("spark_hash");
}
__global__ void k_fill_one(u256* out, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) st256(out + i, fq_one());
}
void fill_one(u256* out, size_t n, cudaStream_t s) {
  k_fill_one<<<grid_for(n, 256, 8), 256, 0, s>>>(out, n);
  SP_LAUNCHED(); check("fill_one");
}

// ---- IPA helpers (nizk/bullet.rs:72-119)
__global__ void k_ipa_fold_ab(u256* a, u256* b, size_t n, const u256 u, const u256 ui) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    u256 aL = ld256(a + i), aR = ld256(a + i + n), bL = ld256(b + i), bR = ld256(b + i + n);
    st256(a + i, fq_add(fq_mul(aL, u), fq_mul(ui, aR)));   // bullet.rs:106
    st256(b + i, fq_add(fq_mul(bL, ui), fq_mul(u, bR)));   // bullet.rs:107
  }
}
void ipa_fold_ab(u256* a, u256* b, size_t n, const u256& u, const u256& uinv, cudaStream_t s) {
  ProfScope ps("ipa_fold_ab", 192.0 * (double)n, s);
  k_ipa_fold_ab<<<grid_for(n, 128, 8), 128, 0, s>>>(a, b, n, u, uinv);
  SP_LAUNCHED(); check("ipa_fold_ab");
}
// G is never folded on the device: after k rounds G_k[i] = sum_{j = i mod n_k} s[j]*G[j], so the round's
// L = <a_L, G_R> and R = <a_R, G_L> are full-length fixed-base MSMs with scalars a[.]*s[j] (zero on the other half).
__global__ void k_ipa_lr(u256* outL, u256* outR, const u256* __restrict__ a, const u256* __restrict__ sv, size_t n_cur, size_t n_full) {
  size_t half = n_cur >> 1;
  for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < n_full; j += (size_t)gridDim.x * blockDim.x) {
    size_t ip = j & (n_cur - 1);
    u256 sj = ld256_ro(sv + j);
    if (ip >= half) { st256(outL + j, fq_mul(ld256_ro(a + ip - half), sj)); st256(outR + j, fq_zero()); }
    else { st256(outL + j, fq_zero()); st256(outR + j, fq_mul(ld256_ro(a + ip + half), sj)); }
  }
}
void ipa_lr_scalars(u256* outL, u256* outR, const u256* a, const u256* svec, size_t n_cur, size_t n_full, cudaStream_t s) {
  k_ipa_lr<<<grid_for(n_full, 128, 8), 128, 0, s>>>(outL, outR, a, svec, n_cur, n_full);
  SP_LAUNCHED(); check("ipa_lr");
}
__global__ void k_ipa_update_s(u256* sv, size_t half, size_t n_full, const u256 u, const u256 ui) {
  for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < n_full; j += (size_t)gridDim.x * blockDim.x) {
    bool right = (j & (2 * half - 1)) >= half;
    st256(sv + j, fq_mul(ld256(sv + j), right ? u : ui));   // G_L[i] <- u^-1 G_L[i] + u G_R[i], bullet.rs:108
  }
}
// both updates of an inner-product round in one launch: rows of blocks y = 0: s[j] *= u^(+-1); y = 1: fold a; y = 2: fold b (one or two products per thread)
__global__ void k_ipa_fold_update(u256* a, u256* b, u256* sv, size_t half, size_t n_full, const u256 u, const u256 ui) {
  const size_t t0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x, step = (size_t)gridDim.x * blockDim.x;
  if (blockIdx.y == 0) {
    for (size_t j = t0; j < n_full; j += step) {
      const bool right = (j & (2 * half - 1)) >= half;
      st256(sv + j, fq_mul(ld256(sv + j), right ? u : ui));   // bullet.rs:108 on the unfolded generators
    }
  } else if (blockIdx.y == 1) {
    for (size_t i = t0; i < half; i += step) st256(a + i, fq_add(fq_mul(ld256(a + i), u), fq_mul(ui, ld256(a + i + half))));   // bullet.rs:106
  } else {
    for (size_t i = t0; i < half; i += step) st256(b + i, fq_add(fq_mul(ld256(b + i), ui), fq_mul(u, ld256(b + i + half))));   // bullet.rs:107
  }
}
void ipa_fold_update(u256* a, u256* b, u256* svec, size_t half, size_t n_full, const u256& u, const u256& uinv, cudaStream_t s) {
  ProfScope ps("ipa_fold_ab", 192.0 * (double)half + 64.0 * (double)n_full, s);
  dim3 grid(grid_for(n_full, 128, 8), 3);
  k_ipa_fold_update<<<grid, 128, 0, s>>>(a, b, svec, half, n_full, u, uinv);
  SP_LAUNCHED(); check("ipa_fold_update");
}
void ipa_update_s(u256* svec, size_t half, size_t n_full, const u256& u, const u256& uinv, cudaStream_t s) {
  k_ipa_update_s<<<grid_for(n_full, 128, 8), 128, 0, s>>>(svec, half, n_full, u, uinv);
  SP_LAUNCHED(); check("ipa_update_s");
}

// ---- sparse
__global__ void k_spmv(u256* out, size_t nrows, const uint32_t* __restrict__ ptr, const uint32_t* __restrict__ idx, const u256* __restrict__ val,
                       const u256* __restrict__ x) {
  for (size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x; r < nrows; r += (size_t)gridDim.x * blockDim.x) {
    u256 acc = fq_zero();
    for (uint32_t k = ptr[r]; k < ptr[r + 1]; k++) acc = fq_add(acc, fq_mul(ld256_ro(val + k), ld256_ro(x + idx[k])));
    st256(out + r, acc);
  }
}
void spmv(u256* out, size_t nrows, const uint32_t* ptr, const uint32_t* idx, const u256* val, const u256* x, cudaStream_t s) {
  ProfScope ps("spmv", 100.0 * (double)nrows, s);
  k_spmv<<<grid_for(nrows, 128, 8), 128, 0, s>>>(out, nrows, ptr, idx, val, x);
  SP_LAUNCHED(); check("spmv");
}
__global__ void __launch_bounds__(256) k_sparse_eval3(const uint32_t* __restrict__ row, const uint32_t* __restrict__ col, const u256* __restrict__ val,
                                                     size_t nnz, const u256* __restrict__ trx, const u256* __restrict__ try_, u256* partials,
                                                     unsigned int* counters, u256* out) {
  u256 acc[1] = {fq_zero()};
  for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < nnz; k += (size_t)gridDim.x * blockDim.x)
    acc[0] = fq_add(acc[0], fq_mul(fq_mul(ld256_ro(trx + row[k]), ld256_ro(try_ + col[k])), ld256_ro(val + k)));
  block_reduce_finish<1>(acc, partials, counters, out, 1);
}
void sparse_eval3(u256* out, const uint32_t* row, const uint32_t* col, const u256* val, size_t nnz, const u256* trx, const u256* try_,
                  void* scratch, cudaStream_t s) {
  unsigned int* counters = (unsigned int*)scratch;
  u256* partials = (u256*)((char*)scratch + 256);
  dim3 grid(grid_for(nnz, 256, 2), 1);
  k_sparse_eval3<<<grid, 256, 0, s>>>(row, col, val, nnz, trx, try_, partials, counters, out);
  SP_LAUNCHED(); check("sparse_eval3");
}

// =============================================================================================== group kernels
__device__ __forceinline__ ge ld_ge(const ge* p) {
  ge r;
  r.X = ld256(&p->X); r.Y = ld256(&p->Y); r.Z = ld256(&p->Z); r.T = ld256(&p->T);
  return r;
}
__device__ __forceinline__ void st_ge(ge* p, const ge& g) { st256(&p->X, g.X); st256(&p->Y, g.Y); st256(&p->Z, g.Z); st256(&p->T, g.T); }

__global__ void k_gens_from_uniform(ge* out, const uint8_t* __restrict__ uni, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const u256* p = reinterpret_cast<const u256*>(uni + 64 * i);
  st_ge(out + i, ristretto_from_uniform(ld256_ro(p), ld256_ro(p + 1)));
}
void gens_from_uniform(ge* out, const uint8_t* d_uniform64, size_t n, cudaStream_t s) {
  k_gens_from_uniform<<<(unsigned)((n + 63) / 64), 64, 0, s>>>(out, d_uniform64, n);
  SP_LAUNCHED(); check("gens_from_uniform");
}
__global__ void k_decompress(ge* out, int* ok, const uint8_t* __restrict__ in, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  ge g = ge_identity();
  bool good = ristretto_decode(g, ld256_ro(reinterpret_cast<const u256*>(in + 32 * i)));
  st_ge(out + i, g);
  if (ok) ok[i] = good ? 1 : 0;
}
void decompress_batch(ge* out, int* ok, const uint8_t* in32, size_t n, cudaStream_t s) {
  k_decompress<<<(unsigned)((n + 63) / 64), 64, 0, s>>>(out, ok, in32, n);
  SP_LAUNCHED(); check("decompress");
}
__global__ void k_compress(uint8_t* out, const ge* __restrict__ in, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  st256(reinterpret_cast<u256*>(out + 32 * i), ristretto_encode(ld_ge(in + i)));
}
void compress_batch(uint8_t* out32, const ge* in, size_t n, cudaStream_t s) {
  ProfScope ps("compress_batch", 160.0 * (double)n, s);
  k_compress<<<(unsigned)((n + 63) / 64), 64, 0, s>>>(out32, in, n);
  SP_LAUNCHED(); check("compress");
}

// ---- fixed-base window tables: entry (j, w, d-1) = d * 2^(W*w) * G_j, affine niels.  W = 8 (32 windows x 128 entries, 393 KB per
// generator) for small generator sets and the host-side copies; W = 13 (20 windows x 4096 entries, 7.9 MB per generator) for the large sets:
// 180 GB of HBM buys 37% fewer point additions per term (20 instead of 32).
static inline int msm_nwin(int wbits) { return (253 + wbits - 1) / wbits; }
static inline size_t msm_depth(int wbits) { return (size_t)1 << (wbits - 1); }
size_t table_entries(size_t nbases, int wbits) { return nbases * (size_t)msm_nwin(wbits) * msm_depth(wbits); }
__global__ void __launch_bounds__(64) k_build_tables(ge_niels* table, const ge* __restrict__ G, size_t nbases, int wbits, int nwin, int depth) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nbases * (size_t)nwin) return;
  size_t j = t / nwin;
  int w = (int)(t % nwin);
  ge P = ld_ge(G + j);
  for (int k = 0; k < wbits * w; k++) P = ge_dbl(P);
  ge acc = P;
  ge_niels* dst = table + t * (size_t)depth;
  for (int d = 0; d < depth; d++) {
    ge_niels nl = ge_to_niels(acc);  // the a=-1 addition law is complete: Z never vanishes
    st256(&dst[d].ypx, nl.ypx); st256(&dst[d].ymx, nl.ymx); st256(&dst[d].t2d, nl.t2d);
    acc = ge_add(acc, P);
  }
}
void build_tables(ge_niels* table, const ge* G, size_t nbases, int wbits, cudaStream_t s) {
  size_t threads = nbases * (size_t)msm_nwin(wbits);
  k_build_tables<<<(unsigned)((threads + 63) / 64), 64, 0, s>>>(table, G, nbases, wbits, msm_nwin(wbits), (int)msm_depth(wbits));
  SP_LAUNCHED(); check("build_tables");
}

// ---- multi-row fixed-base MSM
// grid = (chunks, L); block = 128 threads = COLS columns x GROUPS window-groups of one row.  A thread owns one scalar and the windows
// [g*WPT, (g+1)*WPT) of it; signed W-bit digits, zero digits are skipped (vartime, like the reference's vartime_multiscalar_mul).
// Partial sums are tree-reduced through shared memory.
#ifndef SP_MSM_LB
#define SP_MSM_LB 1
#endif
template <int WBITS>
__device__ __forceinline__ uint32_t msm_window(const u256& k, int w) {
  const int b = w * WBITS, limb = b >> 5, sh = b & 31;
  uint32_t v = k.v[limb] >> sh;
  if (sh + WBITS > 32 && limb + 1 < 8) v |= k.v[limb + 1] << (32 - sh);
  return v & ((1u << WBITS) - 1u);
}
__device__ __forceinline__ ge shfl_down_ge(const ge& p, int delta) {
  ge r;
  r.X = shfl_down_256(p.X, delta); r.Y = shfl_down_256(p.Y, delta); r.Z = shfl_down_256(p.Z, delta); r.T = shfl_down_256(p.T, delta);
  return r;
}
// block-wide point sum: shuffle tree inside each warp (no barriers, so a warp that runs out of non-zero digits early reduces at once),
// then the four warp leaders through shared memory
__device__ __forceinline__ ge block_sum_ge_128(ge acc) {
#pragma unroll 1
  for (int d = 16; d > 0; d >>= 1) acc = ge_add(acc, shfl_down_ge(acc, d));
  __shared__ ge sm[4];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) { acc = ge_add(acc, sm[1]); acc = ge_add(acc, ge_add(sm[2], sm[3])); }
  return acc;
}
template <int WBITS, int GROUPS, int CPT>
__global__ void __launch_bounds__(128, SP_MSM_LB) k_msm_rows(ge* partial, const ge_niels* __restrict__ table, const u256* __restrict__ scalars, size_t stride,
                                                             size_t R, const u256* __restrict__ blinds, size_t blind_base) {
  constexpr int NWIN = (253 + WBITS - 1) / WBITS;
  constexpr int WPT = (NWIN + GROUPS - 1) / GROUPS;   // windows per thread
  constexpr int COLS = 128 / GROUPS;                  // columns per block per pass; CPT passes amortise the reduction tree
  constexpr uint32_t HALF = 1u << (WBITS - 1);
  constexpr size_t DEPTH = (size_t)1 << (WBITS - 1);
  const size_t row = blockIdx.y;
  const size_t ncols = R + (blinds ? 1 : 0);
  const int g = threadIdx.x % GROUPS;
  ge acc = ge_identity();
#pragma unroll 1
  for (int pass = 0; pass < CPT; pass++) {
    const size_t col = ((size_t)blockIdx.x * CPT + pass) * COLS + threadIdx.x / GROUPS;
    if (col >= ncols) break;
    u256 k;
    size_t base;
    if (col < R) { k = ld256_ro(scalars + row * stride + col); base = col; }
    else { k = ld256_ro(blinds + row); base = blind_base; }
    if (fq_is_zero(k)) continue;
    k = fq_from_mont(k);  // group.rs:110-113: scalars leave Montgomery form before the MSM
    // carry into this thread's first window from the signed recoding of the lower windows
    uint32_t carry = 0;
    for (int w = 0; w < g * WPT; w++) carry = (msm_window<WBITS>(k, w) + carry) > HALF ? 1u : 0u;
    const int w_end = (g + 1) * WPT < NWIN ? (g + 1) * WPT : NWIN;
    const ge_niels* tb = table + (base * NWIN + (size_t)g * WPT) * DEPTH;
#pragma unroll 1
    for (int w = g * WPT; w < w_end; w++, tb += DEPTH) {
      uint32_t v = msm_window<WBITS>(k, w) + carry;
      int d;
      if (v > HALF) { d = (int)v - (int)(2 * HALF); carry = 1; } else { d = (int)v; carry = 0; }
      if (d == 0) continue;
      int ad = d < 0 ? -d : d;
      const ge_niels* e = tb + (ad - 1);
      ge_niels nl;
      nl.ypx = ld256_ro(&e->ypx); nl.ymx = ld256_ro(&e->ymx); nl.t2d = ld256_ro(&e->t2d);
      acc = ge_madd(acc, nl, d < 0);
    }
  }
  acc = block_sum_ge_128(acc);
  if (threadIdx.x == 0) st_ge(partial + row * gridDim.x + blockIdx.x, acc);
}
__global__ void __launch_bounds__(128) k_msm_reduce(ge* out, const ge* __restrict__ partial, int chunks) {
  const size_t row = blockIdx.x;
  ge acc = ge_identity();
  for (int c = threadIdx.x; c < chunks; c += 128) acc = ge_add(acc, ld_ge(partial + row * chunks + c));
  acc = block_sum_ge_128(acc);
  if (threadIdx.x == 0) st_ge(out + row, acc);
}
__global__ void __launch_bounds__(32) k_msm_reduce_small(ge* out, const ge* __restrict__ partial, int chunks) {   // chunks <= 32: one warp per row
  const size_t row = blockIdx.x;
  ge acc = (int)threadIdx.x < chunks ? ld_ge(partial + row * chunks + threadIdx.x) : ge_identity();
  for (int d = 16; d > 0; d >>= 1) if (d < 2 * chunks) acc = ge_add(acc, shfl_down_ge(acc, d));
  if (threadIdx.x == 0) st_ge(out + row, acc);
}
__global__ void __launch_bounds__(128) k_ge_add_arrays(ge* a, const ge* __restrict__ b, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) st_ge(a + i, ge_add(ld_ge(a + i), ld_ge(b + i)));
}
// a[i] += b[i] (the blind term of a row commitment, added after the row's MSM: the blinds are drawn on the host while the MSM runs)
void add_points(ge* a, const ge* b, size_t n, cudaStream_t s) {
  if (!n) return;
  k_ge_add_arrays<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(a, b, n);
  SP_LAUNCHED(); check("add_points");
}
// out[0] = sum of n <= 32 points (the "point-add allreduce" of a sharded MSM, after the all-gather of the partial results)
void sum_points(ge* out, const ge* in, int n, cudaStream_t s) {
  if (n < 1 || n > 32) throw std::runtime_error("spartan_b200: sum_points supports 1..32 points");
  k_msm_reduce_small<<<1, 32, 0, s>>>(out, in, n);
  SP_LAUNCHED(); check("sum_points");
}
// ---- the two MSMs of one inner-product round in a single launch (bullet.rs:83-97 with unfolded generators, see k_ipa_lr):
// every generator j carries exactly one non-zero scalar, a[.]*s[j], for L (j in the right half of its n_cur-block) or for R (left half).
// grid = (chunks, 2 sides); the last block to finish sums the partial points of both sides and publishes L, R to the host.
// (latency-bound: 256 CTAs of which each thread makes <= 3 mixed additions and then two point-addition trees; no register cap, so that ptxas can
//  keep the four independent field products of a point addition in flight together)
#ifndef SP_IPA_LB
#define SP_IPA_LB 1
#endif
template <int WBITS, int GROUPS>
__global__ void __launch_bounds__(128, SP_IPA_LB) k_ipa_msm(ge* partial, const ge_niels* __restrict__ table, const u256* __restrict__ a, const u256* __restrict__ sv,
                                                           size_t n_cur, size_t n_full, unsigned int* ticket, ge* out, HostSig sig) {
  constexpr int NWIN = (253 + WBITS - 1) / WBITS;
  constexpr int WPT = (NWIN + GROUPS - 1) / GROUPS;
  constexpr int COLS = 128 / GROUPS;
  constexpr uint32_t HALF = 1u << (WBITS - 1);
  constexpr size_t DEPTH = (size_t)1 << (WBITS - 1);
  const int side = blockIdx.y;
  const size_t half = n_cur >> 1, total = n_full >> 1;
  const int g = threadIdx.x % GROUPS;
  const size_t t = (size_t)blockIdx.x * COLS + threadIdx.x / GROUPS;
  ge acc = ge_identity();
  if (t < total) {
    const size_t blk = t / half, off = t - blk * half;
    const size_t j = blk * n_cur + off + (side == 0 ? half : 0);
    u256 k = fq_mul(ld256_ro(a + (side == 0 ? off : off + half)), ld256_ro(sv + j));
    if (!fq_is_zero(k)) {
      k = fq_from_mont(k);
      uint32_t carry = 0;
      for (int w = 0; w < g * WPT; w++) carry = (msm_window<WBITS>(k, w) + carry) > HALF ? 1u : 0u;
      const int w_end = (g + 1) * WPT < NWIN ? (g + 1) * WPT : NWIN;
      const ge_niels* tb = table + (j * NWIN + (size_t)g * WPT) * DEPTH;
#pragma unroll 1
      for (int w = g * WPT; w < w_end; w++, tb += DEPTH) {
        uint32_t v = msm_window<WBITS>(k, w) + carry;
        int d;
        if (v > HALF) { d = (int)v - (int)(2 * HALF); carry = 1; } else { d = (int)v; carry = 0; }
        if (d == 0) continue;
        int ad = d < 0 ? -d : d;
        const ge_niels* e = tb + (ad - 1);
        ge_niels nl;
        nl.ypx = ld256_ro(&e->ypx); nl.ymx = ld256_ro(&e->ymx); nl.t2d = ld256_ro(&e->t2d);
        acc = ge_madd(acc, nl, d < 0);
      }
    }
  }
  acc = block_sum_ge_128(acc);
  __shared__ bool is_last;
  if (threadIdx.x == 0) {
    st_ge(partial + (size_t)side * gridDim.x + blockIdx.x, acc);
    __threadfence();
    is_last = atomicAdd(ticket, 1u) == 2 * gridDim.x - 1;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  const int s2 = threadIdx.x >> 6, tt = threadIdx.x & 63;   // threads 0..63 finish L, 64..127 finish R
  ge r = ge_identity();
  for (unsigned int c = tt; c < gridDim.x; c += 64) {
    const ge* p = partial + (size_t)s2 * gridDim.x + c;
    ge x;
    x.X = ld256_cg(&p->X); x.Y = ld256_cg(&p->Y); x.Z = ld256_cg(&p->Z); x.T = ld256_cg(&p->T);
    r = ge_add(r, x);
  }
#pragma unroll 1
  for (int d = 16; d > 0; d >>= 1) r = ge_add(r, shfl_down_ge(r, d));
  __shared__ ge fin[4];
  if ((threadIdx.x & 31) == 0) fin[threadIdx.x >> 5] = r;
  __syncthreads();
  if (tt == 0) {
    r = ge_add(fin[2 * s2], fin[2 * s2 + 1]);
    st_ge(out + s2, r);
    if (sig.host_out) {
      st256(sig.host_out + 4 * s2, r.X); st256(sig.host_out + 4 * s2 + 1, r.Y); st256(sig.host_out + 4 * s2 + 2, r.Z); st256(sig.host_out + 4 * s2 + 3, r.T);
      __threadfence_system();
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    *ticket = 0;
    if (sig.flag) { __threadfence_system(); *((volatile unsigned int*)sig.flag) = sig.seq; }
  }
}
// ---- quad-lane point arithmetic for latency-bound reductions.
// A lone warp gains nothing from the four independent field products of a point addition (profiles/r02_probe_latency.txt: 4473 cycles = 9 products
// back to back), so a reduction tree of point additions costs 2.3 us per level.  Here FOUR LANES share one point — lane c of a quad holds coordinate
// c (0: X, 1: Y, 2: Z, 3: T) — and run the products of add-2008-hwcd-3 side by side: three product latencies per addition (A | B | Z1 Z2 | T1 T2, then
// 2d on the T lane, then X3 | Y3 | Z3 | T3) and a few 8-word shuffles instead of nine products in a row.  Same formulas, same field values as
// ge_add / ge_madd; every lane of the warp must take part in every call (full-mask shuffles).
__device__ __forceinline__ u256 shfl_idx_256(const u256& x, int src) {
  u256 r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = __shfl_sync(0xffffffffu, x.v[i], src);
  return r;
}
__device__ __forceinline__ u256 shfl_xor_256(const u256& x, int m) {
  u256 r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = __shfl_xor_sync(0xffffffffu, x.v[i], m);
  return r;
}
__device__ __forceinline__ u256 quad_identity(int c) { return (c == 1 || c == 2) ? fp_one() : fp_zero(); }
// lane-dependent choice without a branch: the sums / differences below are computed by every lane and selected, because a divergent
// `if (c == k)` makes the warp run each arm one after the other (measured: the four arms of quad_finish cost as much as a field product)
__device__ __forceinline__ u256 sel256(bool p, const u256& a, const u256& b) {
  u256 r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = p ? a.v[i] : b.v[i];
  return r;
}
// coordinates -> operands of the first products: lane 0: Y - X, lane 1: Y + X, lane 2: Z, lane 3: T
__device__ __forceinline__ u256 quad_prep(const u256& v, int c) {
  const u256 o = shfl_xor_256(v, 1);
  const u256 d = fp_sub(o, v), sm = fp_add(v, o);          // lane 0: Y - X = o - v; lane 1: Y + X
  return c >= 2 ? v : sel256(c == 0, d, sm);
}
// m = (A, B, D, C) on lanes 0..3  ->  the sum's coordinates (X3, Y3, Z3, T3) on lanes 0..3
static __device__ __noinline__ u256 quad_finish(u256 m, int c, int lane) {
  const u256 o = shfl_xor_256(m, 1);                       // lane 0 <- B, lane 1 <- A, lane 2 <- C, lane 3 <- D
  // lane 0: E = B - A = o - m, lane 1: H = B + A, lane 2: F = D - C = m - o, lane 3: G = D + C
  const u256 df = fp_sub(sel256(c == 0, o, m), sel256(c == 0, m, o)), sm = fp_add(m, o);
  const u256 w = sel256((c & 1) != 0, sm, df);
  const int base = lane & ~3;
  const u256 x = shfl_idx_256(w, base + ((0x2031 >> (4 * c)) & 3));   // lane 0 (E) <- H, lane 1 (H) <- G, lane 2 (F) <- E, lane 3 (G) <- F
  const u256 r = fp_mul(w, x);                             // lane 0: T3 = E H, lane 1: Y3 = H G, lane 2: X3 = F E, lane 3: Z3 = G F
  return shfl_idx_256(r, base + ((0x0312 >> (4 * c)) & 3));            // lane 0 <- X3 (lane 2), lane 1: Y3, lane 2 <- Z3 (lane 3), lane 3 <- T3 (lane 0)
}
static __device__ __noinline__ u256 quad_add(u256 p, u256 q, int c, int lane) {   // P + Q (ge_add)
  const u256 up = quad_prep(p, c), uq = quad_prep(q, c);
  const u256 m = fp_mul(up, uq);                           // A | B | Z1 Z2 | T1 T2
  const u256 m2 = fp_add(m, m);                            // D on lane 2
  const u256 mc = fp_mul(m, fp_2D());                      // C on lane 3 (every lane multiplies: one more product latency either way, no divergence)
  return quad_finish(c == 3 ? mc : (c == 2 ? m2 : m), c, lane);
}
// P + q for an affine niels operand given per lane as (y - x, y + x, -, 2d x y) of +/-q (ge_madd)
__device__ __forceinline__ u256 quad_madd(const u256& p, const u256& nq, int c, int lane) {
  const u256 up = quad_prep(p, c);
  const u256 m = fp_mul(up, nq);                           // A | B | (unused) | C
  const u256 m2 = fp_add(up, up);                          // D = 2 Z1 on lane 2
  return quad_finish(c == 2 ? m2 : m, c, lane);
}
// sum over the 8 quads of a warp -> quad 0
__device__ __forceinline__ u256 quad_warp_sum(u256 P, int c, int lane) {
#pragma unroll 1
  for (int d = 16; d >= 4; d >>= 1) P = quad_add(P, shfl_down_256(P, d), c, lane);
  return P;
}
// Both MSMs of an inner-product round, quad-lane formulation of k_ipa_msm (same inputs, same outputs).  grid = (chunks, 2 sides), IPAQ_THREADS
// threads = IPAQ_THREADS/4 quads; a quad owns WPQ adjacent windows of one scalar and chains their table entries by mixed additions (two product
// latencies each), then 3 tree levels inside the warp, one gather + 3 levels across the warps, and the last block to finish sums the blocks' partial
// points of both sides.  WPQ balances the two costs measured on the B200 (profiles/r02_tuning.md section 7): a lone warp already keeps its
// sub-partition's FMA pipe 64 % busy, so the first quad formulation (two windows per quad, 4608 warps for 4096 generators) was throughput-bound at the
// old kernel's time; six or seven windows per quad need a quarter of the warps and add only ~3 us to the dependent chain.
#ifndef SP_IPAQ_THREADS
#define SP_IPAQ_THREADS 256
#endif
#ifndef SP_IPAQ_SPLIT
#define SP_IPAQ_SPLIT 4   // window groups per scalar (17 windows -> 5 per quad; measured best of 2 / 3 / 4 on the B200, profiles/r02_tuning.md section 7)
#endif
#define IPAQ_THREADS SP_IPAQ_THREADS
template <int WBITS, int WPQ>
__global__ void __launch_bounds__(IPAQ_THREADS, 2) k_ipa_msm_quad(ge* partial, const ge_niels* __restrict__ table, const u256* __restrict__ a, const u256* __restrict__ sv,
                                                                  size_t n_cur, size_t n_full, unsigned int* ticket, ge* out, HostSig sig,
                                                                  const u256* __restrict__ b, u256* c_out, HostSig sigc) {
  if (blockIdx.y == 2) {
    // fused dot products of the round (bullet.rs:78-79), two blocks beside the MSM's: c_L = <a_L, b_R> (block 0), c_R = <a_R, b_L> (block 1); published
    // through their own flag word (sigc), long before the MSM finishes: the host turns them into c*Q while the MSM runs
    if (blockIdx.x >= 2) return;
    const size_t hn = n_cur >> 1;
    const u256* pa = a + (blockIdx.x == 0 ? 0 : hn);
    const u256* pb = b + (blockIdx.x == 0 ? hn : 0);
    u256 acc = fq_zero();
    for (size_t i = threadIdx.x; i < hn; i += IPAQ_THREADS) acc = fq_add(acc, fq_mul(ld256_ro(pa + i), ld256_ro(pb + i)));
    acc = warp_sum_fq(acc);
    __shared__ u256 ds[IPAQ_THREADS / 32];
    if ((threadIdx.x & 31) == 0) ds[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
      u256 v = threadIdx.x < IPAQ_THREADS / 32 ? ds[threadIdx.x] : fq_zero();
      v = warp_sum_fq(v);
      if (threadIdx.x == 0) {
        st256(c_out + blockIdx.x, v);
        if (sigc.host_out) st256(sigc.host_out + blockIdx.x, v);
        __threadfence_system();
        if (atomicAdd(sigc.done, 1u) == 1u) {                 // the second of the two blocks publishes
          *sigc.done = 0;
          if (sigc.flag) { __threadfence_system(); *((volatile unsigned int*)sigc.flag) = sigc.seq; }
        }
      }
    }
    return;
  }
  constexpr int NWIN = (253 + WBITS - 1) / WBITS;
  constexpr int QPS = (NWIN + WPQ - 1) / WPQ;              // quads per scalar
  constexpr int QPB = IPAQ_THREADS / 4, NWARP = IPAQ_THREADS / 32;   // quads per block, warps per block
  constexpr uint32_t HALF = 1u << (WBITS - 1);
  constexpr size_t DEPTH = (size_t)1 << (WBITS - 1);
  const int side = blockIdx.y, tid = threadIdx.x, lane = tid & 31, c = lane & 3, warp = tid >> 5;
  const size_t half = n_cur >> 1, total = n_full >> 1;
  const size_t total_quads = total * QPS, stride = (size_t)gridDim.x * QPB;
#ifdef SP_IPA_TIMELINE
  unsigned long long tl[10];
  int tli = 0;
#define SP_TL() do { if (tid == 0) tl[tli++] = global_timer_ns(); } while (0)
#else
#define SP_TL() do { } while (0)
#endif
  SP_TL();
  u256 P = quad_identity(c);
  // grid-stride over the (scalar, window group) work items with the same trip count in every quad (the shuffles inside quad_madd need the whole
  // warp): a launch may be capped to fewer blocks than work items (the prover does that while a background MSM occupies most SMs)
  for (size_t g0 = (size_t)blockIdx.x * QPB; g0 < total_quads; g0 += stride) {
    const size_t gq = g0 + (tid >> 2);
    const size_t t = gq / QPS;
    const int w0 = WPQ * (int)(gq - t * QPS);
    const bool live = t < total;                           // uniform inside a quad
    u256 k = fq_zero();
    size_t j = 0;
    if (live) {
      const size_t blk = t / half, off = t - blk * half;
      j = blk * n_cur + off + (side == 0 ? half : 0);
      k = fq_mul(ld256_ro(a + (side == 0 ? off : off + half)), ld256_ro(sv + j));
      if (!fq_is_zero(k)) k = fq_from_mont(k);
    }
    SP_TL();
    uint32_t carry = 0;
    for (int w = 0; w < w0; w++) carry = (msm_window<WBITS>(k, w) + carry) > HALF ? 1u : 0u;
    const ge_niels* tb = table + (j * NWIN + (size_t)w0) * DEPTH;
    // all WPQ table entries are fetched up front (independent gathers from a table of up to 110 GB: DRAM + TLB latency paid once, not once per
    // link of the chain); a zero digit or a window past the top contributes the identity as a niels operand: (y - x, y + x, -, 2dxy) = (1, 1, -, 0)
    u256 x[WPQ];
#pragma unroll
    for (int h = 0; h < WPQ; h++) {
      x[h] = c == 3 ? fp_zero() : fp_one();
      if (live && w0 + h < NWIN) {
        uint32_t v = msm_window<WBITS>(k, w0 + h) + carry;
        int d;
        if (v > HALF) { d = (int)v - (int)(2 * HALF); carry = 1; } else { d = (int)v; carry = 0; }
        if (d != 0) {
          const ge_niels* e = tb + (size_t)h * DEPTH + ((d < 0 ? -d : d) - 1);
          if (c == 0) x[h] = ld256_ro(d < 0 ? &e->ypx : &e->ymx);          // y - x of +/-q
          else if (c == 1) x[h] = ld256_ro(d < 0 ? &e->ymx : &e->ypx);     // y + x of +/-q
          else if (c == 3) { x[h] = ld256_ro(&e->t2d); if (d < 0) x[h] = fp_neg(x[h]); }
        }
      }
    }
    SP_TL();
#pragma unroll
    for (int h = 0; h < WPQ; h++) P = quad_madd(P, x[h], c, lane);
  }
  SP_TL();
  P = quad_warp_sum(P, c, lane);
  SP_TL();
  __shared__ u256 sm[NWARP][4];
  __shared__ bool is_last;
  if (lane < 4) sm[warp][c] = P;
  __syncthreads();
  if (warp == 0) {
    const int q = lane >> 2;                               // NWARP = 8: one warp's sum per quad
    P = q < NWARP ? sm[q][c] : quad_identity(c);
    P = quad_warp_sum(P, c, lane);
    if (lane < 4) { st256(reinterpret_cast<u256*>(partial + (size_t)side * gridDim.x + blockIdx.x) + c, P); __threadfence(); }
    __syncwarp();
    if (lane == 0) is_last = atomicAdd(ticket, 1u) == 2 * gridDim.x - 1;
  }
  __syncthreads();
  SP_TL();
#ifdef SP_IPA_TIMELINE
  if (tid == 0 && !is_last && blockIdx.x == 0 && blockIdx.y == 0 && n_cur == n_full)
    printf("ipa_tl %s blk(%d,%d) n=%llu: prep %llu loads+digits %llu madds %llu warp_tree %llu cross+ticket %llu ns (start %llu)\n", is_last ? "LAST" : "first", blockIdx.x, blockIdx.y,
           (unsigned long long)n_full, tl[1] - tl[0], tl[2] - tl[1], tl[3] - tl[2], tl[4] - tl[3], tl[5] - tl[4], tl[0]);
#endif
  if (!is_last) return;
  __threadfence();
  // last block: the first half of its quads finishes L, the second half R
  constexpr int QH = QPB / 2;
  const int s2 = (tid >> 2) / QH, qq = (tid >> 2) % QH;
  const unsigned int nparts = gridDim.x;
  P = quad_identity(c);
  for (unsigned int base = 0; base < nparts; base += QH) {   // same trip count for every quad; out-of-range slots contribute the identity
    const unsigned int idx = base + qq;
    u256 Q = quad_identity(c);
    if (idx < nparts) Q = ld256_cg(reinterpret_cast<const u256*>(partial + (size_t)s2 * nparts + idx) + c);
    P = base == 0 ? Q : quad_add(P, Q, c, lane);
  }
  P = quad_warp_sum(P, c, lane);
  if (lane < 4) sm[warp][c] = P;
  __syncthreads();
  if (warp == 0 || warp == NWARP / 2) {                    // NWARP/2 = 4 warp sums per side
    const int q = lane >> 2;
    P = q < NWARP / 2 ? sm[warp + q][c] : quad_identity(c);
    P = quad_warp_sum(P, c, lane);
    if (lane < 4) {
      st256(reinterpret_cast<u256*>(out + s2) + c, P);
      if (sig.host_out) { st256(sig.host_out + 4 * s2 + c, P); __threadfence_system(); }
    }
  }
  __syncthreads();
#ifdef SP_IPA_TIMELINE
  if (tid == 0 && n_cur == n_full) {
    unsigned long long te = global_timer_ns();
    printf("ipa_tl LAST blk(%d,%d) n=%llu: prep %llu loads+digits %llu madds %llu warp_tree %llu cross+ticket %llu final %llu ns, span %llu ns\n", blockIdx.x, blockIdx.y, (unsigned long long)n_full,
           tl[1] - tl[0], tl[2] - tl[1], tl[3] - tl[2], tl[4] - tl[3], tl[5] - tl[4], te - tl[5], te - tl[0]);
  }
#endif
  if (tid == 0) {
    *ticket = 0;
    if (sig.flag) { __threadfence_system(); *((volatile unsigned int*)sig.flag) = sig.seq; }
  }
}
#undef SP_TL
// windows per quad: the window count split into three groups (17 -> 6, 20 -> 7, 32 -> 11)
template <int WBITS> struct IpaQ { static constexpr int NWIN = (253 + WBITS - 1) / WBITS, WPQ = (NWIN + SP_IPAQ_SPLIT - 1) / SP_IPAQ_SPLIT, QPS = (NWIN + WPQ - 1) / WPQ; };
template <int WBITS>
static void ipa_msm_quad_launch(ge* out, const ge_niels* table, const u256* a, const u256* svec, size_t n_cur, size_t n_full, void* scratch, unsigned int* ticket,
                                cudaStream_t s, HostSig sig, int max_ctas, const u256* b, u256* c_out, HostSig sigc) {
  const size_t quads = (n_full / 2) * IpaQ<WBITS>::QPS, qpb = IPAQ_THREADS / 4;
  dim3 grid((unsigned)((quads + qpb - 1) / qpb), 2);
  if (max_ctas >= 2 && grid.x > (unsigned)max_ctas / 2) grid.x = (unsigned)max_ctas / 2;   // the kernel strides over the work items
  if (b) { grid.y = 3; if (grid.x < 2) grid.x = 2; }        // a third row of blocks, of which two compute the round's dot products
  k_ipa_msm_quad<WBITS, IpaQ<WBITS>::WPQ><<<grid, IPAQ_THREADS, 0, s>>>((ge*)scratch, table, a, svec, n_cur, n_full, ticket, out, sig, b, c_out, sigc);
}
#ifndef SP_IPA_QUAD_DEFAULT
#define SP_IPA_QUAD_DEFAULT 1
#endif
bool ipa_msm_fuses_dots() {   // true when ipa_msm can compute the round's two dot products in the same launch (quad-lane kernel)
  static const bool quad = [] { const char* e = getenv("SP_IPA_QUAD"); return e ? atoi(e) != 0 : SP_IPA_QUAD_DEFAULT != 0; }();
  static const bool fuse = getenv("SP_IPA_NO_FUSED_DOT") == nullptr;
  return quad && fuse;
}
size_t ipa_msm_scratch_points(size_t n_full, int wbits) {   // partial points of either formulation
  const size_t nwin = (size_t)msm_nwin(wbits), total = n_full / 2, wpq = (nwin + SP_IPAQ_SPLIT - 1) / SP_IPAQ_SPLIT, qps = (nwin + wpq - 1) / wpq;
  const size_t quad_chunks = (total * qps + IPAQ_THREADS / 4 - 1) / (IPAQ_THREADS / 4), plain_chunks = (total + 15) / 16;
  return 2 * std::max(quad_chunks, plain_chunks);
}
void ipa_msm(ge* out, const ge_niels* table, int wbits, const u256* a, const u256* svec, size_t n_cur, size_t n_full, void* scratch, unsigned int* ticket,
             cudaStream_t s, HostSig sig, int max_ctas, const u256* b, u256* c_out, HostSig sigc) {
  ProfScope ps("ipa_msm", 64.0 * (double)n_full, s);
  static const bool quad = [] { const char* e = getenv("SP_IPA_QUAD"); return e ? atoi(e) != 0 : SP_IPA_QUAD_DEFAULT != 0; }();
  if (quad) {
    if (wbits == 8) ipa_msm_quad_launch<8>(out, table, a, svec, n_cur, n_full, scratch, ticket, s, sig, max_ctas, b, c_out, sigc);
    else if (wbits == 13) ipa_msm_quad_launch<13>(out, table, a, svec, n_cur, n_full, scratch, ticket, s, sig, max_ctas, b, c_out, sigc);
    else if (wbits == 15) ipa_msm_quad_launch<15>(out, table, a, svec, n_cur, n_full, scratch, ticket, s, sig, max_ctas, b, c_out, sigc);
    else throw std::runtime_error("spartan_b200: unsupported MSM window width");
    SP_LAUNCHED(); check("ipa_msm_quad");
    return;
  }
  if (b) throw std::runtime_error("spartan_b200: the fused dot products need the quad-lane inner-product kernel");
  constexpr int GROUPS = 8;
  const size_t cols = 128 / GROUPS, total = n_full / 2;
  dim3 grid((unsigned)((total + cols - 1) / cols), 2);
  if (wbits == 8) k_ipa_msm<8, GROUPS><<<grid, 128, 0, s>>>((ge*)scratch, table, a, svec, n_cur, n_full, ticket, out, sig);
  else if (wbits == 13) k_ipa_msm<13, GROUPS><<<grid, 128, 0, s>>>((ge*)scratch, table, a, svec, n_cur, n_full, ticket, out, sig);
  else if (wbits == 15) k_ipa_msm<15, GROUPS><<<grid, 128, 0, s>>>((ge*)scratch, table, a, svec, n_cur, n_full, ticket, out, sig);
  else throw std::runtime_error("spartan_b200: unsupported MSM window width");
  SP_LAUNCHED(); check("ipa_msm");
}

static int msm_pick_groups(size_t L, size_t R) {
  // enough threads to fill the chip: one thread per scalar once L*R is large (least reduction work), else split the windows
  size_t cols = L * (R + 1);
  size_t target = (size_t)sm_count() * 1024;
  if (cols >= target) return 1;
  if (cols * 4 >= target) return 4;
  return 8;
}
static int msm_pick_cpt(size_t L, size_t ncols, int groups) {
  // several column passes per thread amortise the block reduction, as long as the grid still oversubscribes the chip
  size_t cols = 128 / groups;
  size_t blocks4 = L * ((ncols + 4 * cols - 1) / (4 * cols));
  return (ncols >= 4 * cols && blocks4 >= (size_t)sm_count() * 4) ? 4 : 1;
}
static size_t msm_chunks(size_t R1, int groups, int cpt) { size_t cols = (128 / groups) * (size_t)cpt; return (R1 + cols - 1) / cols; }
size_t msm_scratch_bytes(size_t L, size_t R) { return L * msm_chunks(R + 1, 8, 1) * sizeof(ge); }
template <int WBITS, int GROUPS>
static void msm_launch2(int cpt, dim3 grid, cudaStream_t s, ge* pp, const ge_niels* table, const u256* sc, size_t stride, size_t R, const u256* bl,
                        size_t blind_base, size_t smem_pad) {
  // smem_pad: unused dynamic shared memory that only lowers the number of resident CTAs per SM (a background MSM leaves room for the
  // latency-bound kernels of the main stream)
  if (cpt == 4) {
    if (smem_pad > (48u << 10)) cudaFuncSetAttribute(k_msm_rows<WBITS, GROUPS, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_pad);
    k_msm_rows<WBITS, GROUPS, 4><<<grid, 128, smem_pad, s>>>(pp, table, sc, stride, R, bl, blind_base);
  } else {
    if (smem_pad > (48u << 10)) cudaFuncSetAttribute(k_msm_rows<WBITS, GROUPS, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_pad);
    k_msm_rows<WBITS, GROUPS, 1><<<grid, 128, smem_pad, s>>>(pp, table, sc, stride, R, bl, blind_base);
  }
}
template <int WBITS>
static void msm_launch(int groups, int cpt, dim3 grid, cudaStream_t s, ge* pp, const ge_niels* table, const u256* sc, size_t stride, size_t R, const u256* bl,
                       size_t blind_base, size_t smem_pad) {
  if (groups == 1) msm_launch2<WBITS, 1>(cpt, grid, s, pp, table, sc, stride, R, bl, blind_base, smem_pad);
  else if (groups == 4) msm_launch2<WBITS, 4>(cpt, grid, s, pp, table, sc, stride, R, bl, blind_base, smem_pad);
  else msm_launch2<WBITS, 8>(cpt, grid, s, pp, table, sc, stride, R, bl, blind_base, smem_pad);
}
void msm_rows(ge* out, const ge_niels* table, int wbits, const u256* scalars, size_t stride, size_t L, size_t R, const u256* blinds, size_t blind_base,
              void* scratch, cudaStream_t s, const MsmTune& tune) {
  ProfScope ps("msm_rows", 32.0 * (double)L * (double)R + 32.0 * (double)R, s);
  const size_t ncols = R + (blinds ? 1 : 0);
  int groups = msm_pick_groups(L, R);
  int cpt = tune.cpt ? tune.cpt : msm_pick_cpt(L, ncols, groups);
  if (cpt != 4) cpt = 1;
  size_t chunks = msm_chunks(ncols, groups, cpt);
  const size_t smem_pad = tune.smem_pad;
  ge* partial = (ge*)scratch;
  for (size_t row0 = 0; row0 < L; row0 += 32768) {   // gridDim.y limit 65535
    size_t rows = L - row0 < 32768 ? L - row0 : 32768;
    dim3 grid((unsigned)chunks, (unsigned)rows);
    const u256* sc = scalars + row0 * stride;
    const u256* bl = blinds ? blinds + row0 : nullptr;
    ge* pp = partial + row0 * chunks;
    if (wbits == 8) msm_launch<8>(groups, cpt, grid, s, pp, table, sc, stride, R, bl, blind_base, smem_pad);
    else if (wbits == 13) msm_launch<13>(groups, cpt, grid, s, pp, table, sc, stride, R, bl, blind_base, smem_pad);
    else if (wbits == 15) msm_launch<15>(groups, cpt, grid, s, pp, table, sc, stride, R, bl, blind_base, smem_pad);
    else throw std::runtime_error("spartan_b200: unsupported MSM window width");
    SP_LAUNCHED();
  }
  if (chunks <= 32) k_msm_reduce_small<<<(unsigned)L, 32, 0, s>>>(out, partial, (int)chunks);
  else k_msm_reduce<<<(unsigned)L, 128, 0, s>>>(out, partial, (int)chunks);
  SP_LAUNCHED(); check("msm_rows");
}

}  // namespace dev
}  // namespace sp
