// spartan_b200 — device context, buffers and generator sets shared by the host prover.
#pragma once
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <chrono>
#include <cstdlib>
#include <map>
#include <memory>
#include <string>
#include <vector>
#include "dev.hpp"
#include "host.hpp"

namespace sp {

void shake256(uint8_t* out, size_t outlen, const uint8_t* in, size_t inlen);

struct SpError : std::runtime_error {  // carries a C-ABI status code (include/spartan_b200.h)
  int code;
  SpError(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

template <class T>
struct DevBuf {  // owning device allocation
  T* p = nullptr;
  size_t n = 0;
  DevBuf() {}
  explicit DevBuf(size_t count) { alloc(count); }
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
  DevBuf& operator=(DevBuf&& o) noexcept { if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; } return *this; }
  ~DevBuf() { release(); }
  void alloc(size_t count) { release(); p = (T*)dev::pool_alloc(count * sizeof(T)); n = count; }
  void release() { if (p) dev::pool_free(p); p = nullptr; n = 0; }
};

// Intra-proof sharding: this rank's window, the peers' windows mapped through CUDA IPC, and the sequence counters every rank
// advances in lock step (all ranks run the same prover code on the same transcript, so they issue the same collectives in the same order).
struct Comm {
  int world = 1, rank = 0;
  uint8_t* win[SP_MAX_RANKS] = {};     // win[rank]: own allocation; others: cudaIpcOpenMemHandle mappings
  bool connected = false;
  unsigned int xseq = 0, bseq = 0;     // per-round partial-sum exchanges / bulk all-gathers issued so far
  DevBuf<unsigned int> ticket;         // block counter of the push kernels (zeroed once, self-resetting)
  Comm();
  ~Comm();
  Comm(const Comm&) = delete;
  Comm& operator=(const Comm&) = delete;
  static size_t window_bytes() { return SP_WIN_CTRL_BYTES + 2 * SP_WIN_HALF_BYTES; }
  void export_handle(uint8_t* out) const { dev::ipc_export(win[rank], out); }   // valid before connect(): rank is 0 then, win[0] is the own window
  void connect(int rank_, int world_, const uint8_t* handles);                   // handles: world x ipc_handle_bytes(), in rank order
  dev::XRank next_xr() {
    dev::XRank x; x.world = world; x.rank = rank; x.seq = ++xseq;
    for (int p = 0; p < world; p++) x.win[p] = reinterpret_cast<dev::WinCtrl*>(win[p]);
    return x;
  }
  dev::CommDev devview() const { dev::CommDev c; c.world = world; c.rank = rank; for (int p = 0; p < world; p++) c.win[p] = win[p]; return c; }
  // data half of the next bulk all-gather (alternating: a half is reused only after every rank has passed the collective in between)
  size_t next_half_off() { ++bseq; return SP_WIN_CTRL_BYTES + (size_t)(bseq & 1) * SP_WIN_HALF_BYTES; }
};

struct Ctx {
  int device = 0;
  std::unique_ptr<Comm> comm;          // null: single-GPU context
  int world() const { return comm && comm->connected ? comm->world : 1; }
  int rank() const { return comm && comm->connected ? comm->rank : 0; }
  // all-gather helpers (no-ops' worth of work when world() == 1 is the caller's business: they require a connected communicator)
  // every rank contributes `bytes` from device memory `src`; returns a device pointer (inside the own window) to world x bytes, rank-major
  const uint8_t* allgather_block(const void* src, size_t bytes);
  // every rank contributes `ntables` tables of n_local elements (cyclic shards); returns the window copy: table t = ptr + t*n_local*world, global order
  u256* allgather_cyclic(const u256* const* tables, int ntables, size_t n_local);
  void comm_create();                  // allocate the window (idempotent); connect through comm->connect
  // Sharding is switched on per prove call (ShardScope): operator-level entry points stay single-GPU even on a connected context, because a
  // collective needs every rank to make the same call.
  bool sharding = false;
  bool shard_enabled = true;           // sp_comm_set_enabled: a connected context can also prove on its own GPU only (every rank must agree)
  int shard_world() const { return sharding ? world() : 1; }
  // a table of `global_len` elements is worth sharding when every rank keeps at least one streaming-size fused round (see sc_fold_eval)
  static constexpr size_t SHARD_MIN_LOCAL = 8192;
  bool shard_table(size_t global_len) const { return shard_world() > 1 && global_len >= 2 * SHARD_MIN_LOCAL * (size_t)shard_world(); }
  cudaStream_t stream = nullptr;       // the prover's stream (greatest priority: its kernels are short and the transcript waits for each of them)
  // background stream (least priority) for throughput work whose inputs are known early and whose result the transcript needs late: the commitment
  // to the dereferenced SPARK values runs there under the second sumcheck phase and the witness evaluation proof (snark.cpp)
  cudaStream_t stream2 = nullptr;
  bool overlap = true;                 // sp_ctx_set_overlap: 0 keeps every kernel on the prover's stream (profiling: per-kernel event times free of concurrent work)
  bool bg_busy = false;                // between fork and join of background work (snark.cpp): latency kernels shape their grids for the free SMs
  int stream2_sms = 0;                 // > 0: stream2 is confined to that many SMs (green context); 0: it shares every SM with the prover's stream
  void* ev_fork = nullptr; void* ev_join = nullptr;
  DevBuf<uint8_t> scratch2;        // MSM partial sums of the background stream
  uint8_t* pinned = nullptr;       // staging for small host<->device exchanges
  size_t pinned_bytes = 0;
  DevBuf<uint8_t> scratch;         // MSM partial sums (growable)
  DevBuf<uint8_t> red;             // reduction tickets + per-block partials (fixed size, tickets zeroed once and self-resetting)
  DevBuf<u256> small;              // challenges, results (device side); [4000, 4002) = two Montgomery ones
  const u256* ones() const { return small.p + 4000; }
  std::string last_error;
  // per-phase timers (profile feature of the reference, src/timer.rs): label -> milliseconds of the last prove
  std::vector<std::pair<std::string, double>> timings;
  std::map<std::string, double> fine;   // accumulated sub-phase timers (SP_FINE_TIMERS=1), flushed into `timings` by the prove entry points

  void* ev_a = nullptr; void* ev_b = nullptr; bool timer_running = false;   // sp_timer_start / sp_timer_stop_ms
  explicit Ctx(int dev);
  ~Ctx();
  // host-polled kernel results (mapped pinned memory): see dev::HostSig
  u256* host_res = nullptr;
  unsigned int* host_flag = nullptr;
  DevBuf<unsigned int> sig_done;
  unsigned int sig_seq = 0;
  unsigned int sigc_seq = 0;           // sequence of the second flag word (host_flag + 16)
  dev::HostSig next_sig() { dev::HostSig s; s.host_out = host_res; s.flag = host_flag; s.done = sig_done.p; s.seq = ++sig_seq; return s; }
  void wait_sig(const dev::HostSig& s);   // spins on the flag; falls back to a stream synchronise to surface CUDA errors
  // host -> persistent-kernel mailbox (dev::sc_persist): the next challenge and its sequence number, in mapped pinned memory
  dev::PersistMail* mail = nullptr;
  DevBuf<dev::PersistMail> dmail;      // device-side copy the polling CTA forwards the challenge through
  unsigned int mail_seq = 0;
  void post_challenge(const Fq& r) {
    memcpy((void*)&mail->r, &r.m, sizeof(u256));
    std::atomic_thread_fence(std::memory_order_release);
    *((volatile unsigned int*)&mail->seq) = ++mail_seq;
  }
  void sync() { dev::stream_sync(stream); }
  void ensure_scratch(size_t bytes) { if (scratch.n < bytes) { sync(); scratch.alloc(bytes); } }
  // upload k scalars to small[slot..]
  void put_small(size_t slot, const Fq* v, size_t k);
  void get_small(size_t slot, Fq* v, size_t k);  // synchronises
  void upload(u256* d, const Fq* h, size_t n);
  std::vector<Fq> download(const u256* d, size_t n);
};

struct ShardScope {   // RAII: sharded proving for the duration of one prove call on a connected context
  Ctx& ctx; bool prev;
  explicit ShardScope(Ctx& c) : ctx(c), prev(c.sharding) { c.sharding = c.world() > 1 && c.shard_enabled; }
  ~ShardScope() { ctx.sharding = prev; }
};
// eq(r, j*W + rank) = c * eq(r[0 .. ell-log2 W), j): the factor contributed by the last log2(W) variables, which the rank fixes
inline Fq shard_eq_scale(const std::vector<Fq>& r, int W, int rank) {
  int logW = 0;
  while ((1 << logW) < W) logW++;
  Fq c = Fq::one();
  for (int k = 0; k < logW; k++) {
    const Fq& rj = r[r.size() - logW + k];
    c *= ((rank >> (logW - 1 - k)) & 1) ? rj : Fq::one() - rj;
  }
  return c;
}

// One SHAKE256 generator stream (a `label` of MultiCommitGens::new, commitments.rs:15-33) expanded to `nbases` points,
// with the fixed-base window table on the device and host copies of the tables of a few named bases.
struct FineTimer {  // host wall-clock accumulator for latency hunting; a no-op unless SP_FINE_TIMERS is set
  Ctx& ctx; const char* name; std::chrono::steady_clock::time_point t0; bool on;
  static bool enabled() { static const bool e = getenv("SP_FINE_TIMERS") != nullptr; return e; }
  FineTimer(Ctx& c, const char* n) : ctx(c), name(n), on(enabled()) { if (on) t0 = std::chrono::steady_clock::now(); }
  void stop() { if (on) { ctx.fine[name] += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); on = false; } }
  ~FineTimer() { stop(); }
};

struct GenSet {
  Ctx* ctx;
  std::string label;
  size_t nbases;
  DevBuf<ge> G;
  DevBuf<ge_niels> table;
  int wbits = 8;               // window width of `table` (host_tab is always 8-bit)
  std::map<size_t, HostBaseTable> host_tab;
  std::vector<Cp> compressed;  // lazily filled export
  GenSet(Ctx* c, const std::string& label, size_t nbases, const std::vector<size_t>& host_bases);
  GenSet(Ctx* c, const ge* d_points, size_t nbases, const std::vector<size_t>& host_bases);   // caller-supplied generators (device array)
  void finish(const std::vector<size_t>& host_bases);   // window tables + host copies, once G is filled
  const HostBaseTable& tab(size_t base) const {
    auto it = host_tab.find(base);
    if (it == host_tab.end()) throw std::runtime_error("spartan_b200: no host table for generator " + std::to_string(base));
    return it->second;
  }
  hge host_point(size_t base) const;  // 1 * G_base from the host table
};

// A MultiCommitGens view (commitments.rs:8-12): G = set.G[off .. off+n), h = set.G[h]
struct CommitKey {
  const GenSet* set = nullptr;
  size_t off = 0, n = 0, h = 0;
};

// A few helper threads for the host's single-point commitments (process-wide, created on first use).  The ZK sumcheck rounds are host-bound: between a
// round's evaluations and its challenge the prover thread commits to the four coefficients of the round polynomial — four independent fixed-base scalar
// multiplications of ~4 us each.  run(n, fn) executes fn(0) .. fn(n-1), fn(0) on the caller and the rest on helpers that spin for work while proofs are in
// flight (they fall asleep on a condition variable after ~200 us without any).  SP_HOST_THREADS=<n> switches them on (default 0: measured, no gain — the
// hand-over costs what the parallel scalar multiplications save); results do not depend on it (group addition is commutative, encodings are canonical).
class HostPool {
 public:
  static HostPool& get();
  int helpers() const { return (int)th_.size(); }
  void run(int njobs, const std::function<void(int)>& fn);
  ~HostPool();
 private:
  HostPool();
  void worker();
  std::vector<std::thread> th_;
  std::atomic<uint64_t> epoch_{0};
  // next_ = (run tag << 32) | next job index, desc_ = (run tag << 32) | job count: a helper that drew an index from an earlier run's counter sees the
  // tag mismatch and drops it (index and job count are never read from two different runs)
  std::atomic<uint64_t> next_{0}, desc_{0};
  std::atomic<int> pending_{0}, sleepers_{0};
  std::atomic<bool> stop_{false};
  std::atomic<const std::function<void(int)>*> fn_{nullptr};
  uint64_t tag_ = 0;                   // caller side only
  std::mutex mu_;
  std::condition_variable cv_;
};

struct Term { size_t base; Fq k; };
// sum of k_i * G_{base_i} over host tables
hge host_commit(const GenSet& gs, const Term* terms, size_t nterms);
inline Cp host_commit_c(const GenSet& gs, const std::vector<Term>& t) { return compress(host_commit(gs, t.data(), t.size())); }

}  // namespace sp
