// spartan_b200 — host prover for the R1CS satisfiability proof (NIZK path), driving the sm_100a kernels.
// Follows the exact Fiat-Shamir schedule of /root/reference/src/r1csproof.rs:144-349 so that the
// proof bytes equal the reference's for identical instance, assignment, transcript label and RandomTape seed.
#include "prover.hpp"
#include "../../include/spartan_b200.h"
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <functional>

namespace sp {

// ================================================================================================ plumbing
void shake256(uint8_t* out, size_t outlen, const uint8_t* in, size_t inlen) {
  uint64_t st[25];
  memset(st, 0, sizeof st);
  uint8_t* sb = reinterpret_cast<uint8_t*>(st);
  const size_t rate = 136;
  while (inlen >= rate) { for (size_t i = 0; i < rate; i++) sb[i] ^= in[i]; Keccak::f1600(st); in += rate; inlen -= rate; }
  for (size_t i = 0; i < inlen; i++) sb[i] ^= in[i];
  sb[inlen] ^= 0x1f; sb[rate - 1] ^= 0x80;
  Keccak::f1600(st);
  while (outlen) {
    size_t n = outlen < rate ? outlen : rate;
    memcpy(out, sb, n); out += n; outlen -= n;
    if (outlen) Keccak::f1600(st);
  }
}

#ifndef SP_BG_SMS_DEFAULT
#define SP_BG_SMS_DEFAULT 0
#endif
Ctx::Ctx(int dev_) : device(dev_) {
  if (dev::device_count() <= dev_) throw std::runtime_error("spartan_b200: no CUDA device " + std::to_string(dev_) + " (the prover has no CPU fallback)");
  dev::set_device(dev_);
  stream = dev::stream_create_prio(+1);
  {
    // SP_BG_SMS: SMs of the background partition (0: no partition, the background stream shares all SMs at the least priority)
    const char* e = getenv("SP_BG_SMS");
    const int want = e ? atoi(e) : SP_BG_SMS_DEFAULT;
    if (want > 0) stream2 = dev::stream_create_partition(want, -1, &stream2_sms);
    if (!stream2) { stream2_sms = 0; stream2 = dev::stream_create_prio(-1); }
  }
  ev_fork = dev::event_create(); ev_join = dev::event_create();
  pinned_bytes = 1 << 20;
  pinned = (uint8_t*)dev::hmalloc_pinned(pinned_bytes);
  host_res = reinterpret_cast<u256*>(pinned + (512 << 10));
  host_flag = reinterpret_cast<unsigned int*>(pinned + (768 << 10));
  *host_flag = 0;
  host_flag[16] = 0;   // second flag word (its own cache line): the dot products fused into the inner-product MSM launch
  mail = reinterpret_cast<dev::PersistMail*>(pinned + (900 << 10));
  memset(mail, 0, sizeof(dev::PersistMail));
  dmail.alloc(1);
  dev::dzero(dmail.p, sizeof(dev::PersistMail), stream);
  sig_done.alloc(4);
  dev::dzero(sig_done.p, 16, stream);
  small.alloc(4096);
  dev::fill_one(small.p + 4000, 2, stream);   // ones(): two Montgomery ones
  scratch.alloc(1 << 20);
  red.alloc(dev::sc_scratch_bytes(24));
  dev::dzero(red.p, red.n, stream);
  sync();
}
void Ctx::comm_create() {
  if (comm) return;
  dev::set_device(device);
  comm.reset(new Comm());
  dev::dzero(comm->ticket.p, 16, stream);
  sync();
}
Ctx::~Ctx() {
  try { sync(); } catch (...) {}
  try { dev::stream_sync(stream2); } catch (...) {}
  scratch.release(); scratch2.release(); red.release(); small.release(); dmail.release();
  dev::event_destroy(ev_fork); dev::event_destroy(ev_join);
  dev::stream_destroy(stream2);
  if (ev_a) { dev::event_destroy(ev_a); dev::event_destroy(ev_b); }
  dev::hfree_pinned(pinned);
  dev::stream_destroy(stream);
}
void Ctx::wait_sig(const dev::HostSig& s) {
  volatile unsigned int* f = s.flag;
  auto t0 = std::chrono::steady_clock::now();
  unsigned long spins = 0;
  while ((int)(*f - s.seq) < 0) {   // sequence numbers only grow (one stream): a later signal also satisfies an earlier wait
    if ((++spins & 0xffff) == 0 && std::chrono::steady_clock::now() - t0 > std::chrono::seconds(5)) {
      sync();   // throws on a CUDA error; otherwise the kernel has finished and the flag must be visible
      if ((int)(*f - s.seq) < 0) throw std::runtime_error("spartan_b200: kernel completed without publishing its result flag");
    }
  }
  std::atomic_thread_fence(std::memory_order_acquire);
}
// ---- communicator
Comm::Comm() {
  win[0] = (uint8_t*)dev::win_alloc(window_bytes());
  ticket.alloc(4);
}
Comm::~Comm() {
  for (int p = 0; p < SP_MAX_RANKS; p++) {
    if (!win[p]) continue;
    if (p == rank) dev::win_free(win[p]); else dev::ipc_close(win[p]);
  }
}
void Comm::connect(int rank_, int world_, const uint8_t* handles) {
  if (connected) throw SpError(SP_ERR_INVALID_ARG, "communicator already connected");
  if (world_ < 1 || world_ > SP_MAX_RANKS || (world_ & (world_ - 1)) || rank_ < 0 || rank_ >= world_) throw SpError(SP_ERR_INVALID_ARG, "sharding needs a power-of-two world of at most 8 ranks");
  uint8_t* own = win[0];
  win[0] = nullptr;
  rank = rank_; world = world_;
  const size_t hb = dev::ipc_handle_bytes();
  for (int p = 0; p < world; p++) win[p] = p == rank ? own : (uint8_t*)dev::ipc_open(handles + (size_t)p * hb);
  connected = true;
}
const uint8_t* Ctx::allgather_block(const void* src, size_t bytes) {
  Comm& c = *comm;
  if ((size_t)c.world * bytes > SP_WIN_HALF_BYTES) throw std::runtime_error("spartan_b200: all-gather larger than the window");
  const size_t off = c.next_half_off();
  dev::push_block(c.devview(), src, bytes, off, c.bseq, c.ticket.p, stream);
  dev::wait_peers(c.devview(), c.bseq, stream);
  return c.win[c.rank] + off;
}
u256* Ctx::allgather_cyclic(const u256* const* tables, int ntables, size_t n_local) {
  Comm& c = *comm;
  if ((size_t)ntables * n_local * c.world * sizeof(u256) > SP_WIN_HALF_BYTES) throw std::runtime_error("spartan_b200: all-gather larger than the window");
  const size_t off = c.next_half_off();
  dev::push_cyclic(c.devview(), tables, ntables, n_local, off, c.bseq, c.ticket.p, stream);
  dev::wait_peers(c.devview(), c.bseq, stream);
  return reinterpret_cast<u256*>(c.win[c.rank] + off);
}

void Ctx::put_small(size_t slot, const Fq* v, size_t k) {
  // staged through pageable memory on purpose: cudaMemcpyAsync from pageable memory snapshots the source before returning
  dev::h2d(small.p + slot, v, k * sizeof(u256), stream);
}
void Ctx::get_small(size_t slot, Fq* v, size_t k) {
  dev::d2h(pinned, small.p + slot, k * sizeof(u256), stream);
  sync();
  memcpy(v, pinned, k * sizeof(u256));
}
void Ctx::upload(u256* d, const Fq* h, size_t n) { dev::h2d(d, h, n * sizeof(u256), stream); sync(); }
std::vector<Fq> Ctx::download(const u256* d, size_t n) {
  std::vector<Fq> v(n);
  dev::d2h(v.data(), d, n * sizeof(u256), stream);
  sync();
  return v;
}

GenSet::GenSet(Ctx* c, const std::string& label_, size_t nbases_, const std::vector<size_t>& host_bases) : ctx(c), label(label_), nbases(nbases_) {
  // MultiCommitGens::new (commitments.rs:15-33): SHAKE256(label || basepoint), 64 bytes per generator
  static const uint8_t basepoint[32] = {0xe2, 0xf2, 0xae, 0x0a, 0x6a, 0xbc, 0x4e, 0x71, 0xa8, 0x84, 0xa9, 0x61, 0xc5, 0x00, 0x51, 0x5f,
                                        0x58, 0xe3, 0x0b, 0x6a, 0xa5, 0x82, 0xdd, 0x8d, 0xb6, 0xa6, 0x59, 0x45, 0xe0, 0x8d, 0x2d, 0x76};
  std::vector<uint8_t> seed(label.begin(), label.end());
  seed.insert(seed.end(), basepoint, basepoint + 32);
  std::vector<uint8_t> uni(64 * nbases);
  shake256(uni.data(), uni.size(), seed.data(), seed.size());
  DevBuf<uint8_t> d_uni(uni.size());
  dev::h2d(d_uni.p, uni.data(), uni.size(), ctx->stream);
  G.alloc(nbases);
  dev::gens_from_uniform(G.p, d_uni.p, nbases, ctx->stream);
  finish(host_bases);
}
GenSet::GenSet(Ctx* c, const ge* d_points, size_t nbases_, const std::vector<size_t>& host_bases) : ctx(c), label("<uploaded>"), nbases(nbases_) {
  G.alloc(nbases);
  dev::d2d(G.p, d_points, nbases * sizeof(ge), ctx->stream);
  finish(host_bases);
}
void GenSet::finish(const std::vector<size_t>& host_bases) {
  // large generator sets get 13-bit windows (20 additions per term instead of 32; 7.9 MB of table per generator); SP_MSM_WINDOW=15 trades
  // 26.7 MB per generator for 17 additions per term (110 GB for the 4098 generators of a 2^20 SNARK: fits one B200, not a 2^22 one)
  const char* wenv = getenv("SP_MSM_WINDOW");
  if (nbases < 512) wbits = 8;
  else if (wenv) wbits = atoi(wenv);
  else {
    // 15-bit windows for the big SPARK generator set when the table fits comfortably in what is free now (measured: msm_rows 14.3 -> 13.1 ms per
    // 2^20 proof); the 1026-generator witness set stays at 13 bits (its tables would triple for 0.3 ms)
    wbits = 13;
    if (nbases >= 2048) {
      size_t free_b = 0, total_b = 0;
      dev::mem_info(&free_b, &total_b);
      if ((double)dev::table_entries(nbases, 15) * sizeof(ge_niels) <= 0.65 * (double)free_b) wbits = 15;
    }
  }
  if (wbits != 8 && wbits != 13 && wbits != 15) throw std::runtime_error("spartan_b200: SP_MSM_WINDOW must be 8, 13 or 15");
  table.alloc(dev::table_entries(nbases, wbits));
  dev::build_tables(table.p, G.p, nbases, wbits, ctx->stream);
  // host copies (8-bit windows) of the few generators the sigma protocols commit against
  std::vector<size_t> hb;
  for (size_t b : host_bases) if (b < nbases && std::find(hb.begin(), hb.end(), b) == hb.end()) hb.push_back(b);
  if (!hb.empty()) {
    DevBuf<ge> sel(hb.size());
    for (size_t i = 0; i < hb.size(); i++) dev::d2d(sel.p + i, G.p + hb[i], sizeof(ge), ctx->stream);
    DevBuf<ge_niels> small(dev::table_entries(hb.size(), 8));
    dev::build_tables(small.p, sel.p, hb.size(), 8, ctx->stream);
    for (size_t i = 0; i < hb.size(); i++) {
      std::vector<ge_niels> raw(32 * 128);
      dev::d2h(raw.data(), small.p + i * 32 * 128, sizeof(ge_niels) * 32 * 128, ctx->stream);
      ctx->sync();
      HostBaseTable t;
      t.e.resize(32 * 128);
      for (size_t e = 0; e < raw.size(); e++) t.e[e] = to_hniels(raw[e]);
      host_tab[hb[i]] = std::move(t);
    }
  }
  ctx->sync();
}
hge GenSet::host_point(size_t base) const {
  hge acc = hge_identity();
  return hge_madd(acc, tab(base).e[0], false);
}
// ---- HostPool
HostPool& HostPool::get() { static HostPool p; return p; }
HostPool::HostPool() {
  int n = 0;   // off by default: measured on the B200 box (profiles/r02_tuning.md section 10), three helpers cost the ZK rounds as much as they save
  if (const char* e = getenv("SP_HOST_THREADS")) n = atoi(e);
  if (std::thread::hardware_concurrency() < 8) n = 0;
  if (n < 0) n = 0;
  if (n > 7) n = 7;
  for (int i = 0; i < n; i++) th_.emplace_back([this] { worker(); });
}
HostPool::~HostPool() {
  stop_.store(true);
  { std::lock_guard<std::mutex> lk(mu_); cv_.notify_all(); }
  for (auto& t : th_) t.join();
}
void HostPool::worker() {
  uint64_t seen = 0;
  auto last_work = std::chrono::steady_clock::now();
  unsigned spins = 0;
  while (!stop_.load(std::memory_order_relaxed)) {
    const uint64_t e = epoch_.load(std::memory_order_acquire);
    if (e != seen) {
      seen = e;
      for (;;) {
        const uint64_t v = next_.fetch_add(1, std::memory_order_acq_rel);
        const uint64_t d = desc_.load(std::memory_order_acquire);
        if ((v >> 32) != (d >> 32) || (uint32_t)v >= (uint32_t)d) break;
        (*fn_.load(std::memory_order_acquire))((int)(uint32_t)v);
        pending_.fetch_sub(1, std::memory_order_acq_rel);
      }
      last_work = std::chrono::steady_clock::now();
      spins = 0;
      continue;
    }
#if defined(__x86_64__)
    __builtin_ia32_pause();
#endif
    if ((++spins & 0x3ff) == 0 && std::chrono::steady_clock::now() - last_work > std::chrono::microseconds(200)) {
      std::unique_lock<std::mutex> lk(mu_);
      sleepers_.fetch_add(1);
      cv_.wait_for(lk, std::chrono::milliseconds(50), [&] { return stop_.load() || epoch_.load(std::memory_order_acquire) != seen; });
      sleepers_.fetch_sub(1);
      last_work = std::chrono::steady_clock::now();
    }
  }
}
void HostPool::run(int njobs, const std::function<void(int)>& fn) {
  if (njobs <= 1 || th_.empty()) { for (int i = 0; i < njobs; i++) fn(i); return; }
  const uint64_t tag = ++tag_ & 0xffffffu;
  fn_.store(&fn, std::memory_order_relaxed);
  pending_.store(njobs - 1, std::memory_order_relaxed);
  desc_.store(tag << 32 | (uint32_t)njobs, std::memory_order_release);
  next_.store(tag << 32 | 1u, std::memory_order_release);
  epoch_.fetch_add(1, std::memory_order_release);
  if (sleepers_.load(std::memory_order_acquire) > 0) { std::lock_guard<std::mutex> lk(mu_); cv_.notify_all(); }
  fn(0);
  for (;;) {   // the caller takes whatever the helpers have not started
    const uint64_t v = next_.fetch_add(1, std::memory_order_acq_rel);
    if ((uint32_t)v >= (uint32_t)njobs) break;
    fn((int)(uint32_t)v);
    pending_.fetch_sub(1, std::memory_order_acq_rel);
  }
  while (pending_.load(std::memory_order_acquire) > 0) {
#if defined(__x86_64__)
    __builtin_ia32_pause();
#endif
  }
}

hge host_commit(const GenSet& gs, const Term* terms, size_t nterms) {
  if (nterms >= 3 && HostPool::get().helpers() > 0) {   // one fixed-base scalar multiplication per job, partial points added here
    hge part[8];
    const int n = (int)std::min<size_t>(nterms, 8);
    const HostBaseTable* tb[8];
    for (int i = 0; i < n; i++) tb[i] = &gs.tab(terms[i].base);   // look-ups (and their exceptions) stay on this thread
    HostPool::get().run(n, [&](int i) { part[i] = hge_identity(); host_fixed_mul_acc(part[i], *tb[i], terms[i].k); });
    hge acc = part[0];
    for (int i = 1; i < n; i++) acc = hge_add(acc, part[i]);
    for (size_t i = (size_t)n; i < nterms; i++) host_fixed_mul_acc(acc, gs.tab(terms[i].base), terms[i].k);
    return acc;
  }
  hge acc = hge_identity();
  for (size_t i = 0; i < nterms; i++) host_fixed_mul_acc(acc, gs.tab(terms[i].base), terms[i].k);
  return acc;
}

R1CSGens::R1CSGens(Ctx* ctx, const std::string& label, size_t num_vars) {
  // R1CSGens::new (r1csproof.rs:68-73) -> PolyCommitmentGens::new(log2 num_vars) (dense_mlpoly.rs:31-35) ->
  // DotProductProofGens::new(2^ceil(ell/2)) = MultiCommitGens::new(n+1).split_at(n) (nizk/mod.rs:414-418)
  size_t ell = 0;
  while (((size_t)1 << ell) < num_vars) ell++;
  size_t n = (size_t)1 << (ell - ell / 2);
  set.reset(new GenSet(ctx, label, std::max<size_t>(n + 2, 5), {0, 1, 2, 3, 4, n, n + 1}));  // gens_4 draws 5 points of the same stream
  gens_pc.n = n;
  gens_pc.gens_n = CommitKey{set.get(), 0, n, n + 1};
  gens_pc.gens_1 = CommitKey{set.get(), n, 1, n + 1};
  gens_1 = gens_pc.gens_1;                       // R1CSSumcheckGens::new clones gens_pc.gens.gens_1 (r1csproof.rs:48-58)
  gens_3 = CommitKey{set.get(), 0, 3, 3};        // MultiCommitGens::new(3, label): same SHAKE prefix, h = 4th point
  gens_4 = CommitKey{set.get(), 0, 4, 4};
}

// ================================================================================================ small host protocol pieces
static Cp commit1(const CommitKey& k, const Fq& x, const Fq& blind) {  // impl Commitments for Scalar (commitments.rs:73-78)
  Term t[2] = {{k.off, x}, {k.h, blind}};
  return compress(host_commit(*k.set, t, 2));
}
static Cp commitv(const CommitKey& k, const std::vector<Fq>& x, const Fq& blind) {  // impl Commitments for [Scalar] (:80-92), small n only
  std::vector<Term> t;
  for (size_t i = 0; i < x.size(); i++) t.push_back({k.off + i, x[i]});
  t.push_back({k.h, blind});
  return compress(host_commit(*k.set, t.data(), t.size()));
}

static KnowledgeProof knowledge_prove(const CommitKey& g, Transcript& T, RandomTape& tape, const Fq& x, const Fq& r, Cp& C_out) {  // nizk/mod.rs:27-52
  T.append_protocol_name("knowledge proof");
  Fq t1 = tape.random_scalar("t1"), t2 = tape.random_scalar("t2");
  C_out = commit1(g, x, r);
  T.append_point("C", C_out.b);
  KnowledgeProof p;
  p.alpha = commit1(g, t1, t2);
  T.append_point("alpha", p.alpha.b);
  Fq c = T.challenge_scalar("c");
  p.z1 = x * c + t1;
  p.z2 = r * c + t2;
  return p;
}
static EqualityProof equality_prove(const CommitKey& g, Transcript& T, RandomTape& tape, const Fq& v1, const Fq& s1, const Fq& v2, const Fq& s2) {  // :88-116
  T.append_protocol_name("equality proof");
  Fq r = tape.random_scalar("r");
  Cp C1 = commit1(g, v1, s1); T.append_point("C1", C1.b);
  Cp C2 = commit1(g, v2, s2); T.append_point("C2", C2.b);
  EqualityProof p;
  Term t[1] = {{g.h, r}};
  p.alpha = compress(host_commit(*g.set, t, 1));  // r * h
  T.append_point("alpha", p.alpha.b);
  Fq c = T.challenge_scalar("c");
  p.z = c * (s1 - s2) + r;
  return p;
}
static ProductProof product_prove(const CommitKey& g, Transcript& T, RandomTape& tape, const Fq& x, const Fq& rX, const Fq& y, const Fq& rY, const Fq& z,
                                  const Fq& rZ, Cp& X, Cp& Y, Cp& Z) {  // nizk/mod.rs:159-229
  T.append_protocol_name("product proof");
  Fq b1 = tape.random_scalar("b1"), b2 = tape.random_scalar("b2"), b3 = tape.random_scalar("b3"), b4 = tape.random_scalar("b4"), b5 = tape.random_scalar("b5");
  X = commit1(g, x, rX); T.append_point("X", X.b);
  Y = commit1(g, y, rY); T.append_point("Y", Y.b);
  Z = commit1(g, z, rZ); T.append_point("Z", Z.b);
  ProductProof p;
  p.alpha = commit1(g, b1, b2); T.append_point("alpha", p.alpha.b);
  p.beta = commit1(g, b3, b4); T.append_point("beta", p.beta.b);
  // delta = b3*X + b5*h with X = x*G + rX*h (nizk/mod.rs:199-206): same group element as (b3*x)*G + (b3*rX + b5)*h
  p.delta = commit1(g, b3 * x, b3 * rX + b5);
  T.append_point("delta", p.delta.b);
  Fq c = T.challenge_scalar("c");
  p.z = {b1 + c * x, b2 + c * rX, b3 + c * y, b4 + c * rY, b5 + c * (rZ - rX * y)};
  return p;
}
// DotProductProof::prove (nizk/mod.rs:311-370); Cx is passed in when the caller already holds commit(x_vec; blind_x)
// `pre` (optional): tape draws made ahead of time in the reference's order, with the commitments that depend on the tape only
struct DotPre { std::vector<Fq> d_vec; Fq r_delta, r_beta; Cp delta; hge r_beta_h; };
static DotProductProof dotproduct_prove(const CommitKey& g1, const CommitKey& gn, Transcript& T, RandomTape& tape, const std::vector<Fq>& x_vec,
                                        const Fq& blind_x, const std::vector<Fq>& a_vec, const Fq& y, const Fq& blind_y, const Cp* Cx_known,
                                        const DotPre* pre = nullptr) {
  T.append_protocol_name("dot product proof");
  size_t n = x_vec.size();
  std::vector<Fq> d_vec = pre ? pre->d_vec : tape.random_vector("d_vec", n);
  Fq r_delta = pre ? pre->r_delta : tape.random_scalar("r_delta"), r_beta = pre ? pre->r_beta : tape.random_scalar("r_beta");
  Cp Cx = Cx_known ? *Cx_known : commitv(gn, x_vec, blind_x);
  T.append_point("Cx", Cx.b);
  Cp Cy = commit1(g1, y, blind_y);
  T.append_point("Cy", Cy.b);
  T.append_scalars("a", a_vec);
  DotProductProof p;
  p.delta = pre ? pre->delta : commitv(gn, d_vec, r_delta);
  T.append_point("delta", p.delta.b);
  Fq dot = Fq::zero();
  for (size_t i = 0; i < n; i++) dot += a_vec[i] * d_vec[i];
  if (pre) { Term t1[1] = {{g1.off, dot}}; p.beta = compress(hge_add(host_commit(*g1.set, t1, 1), pre->r_beta_h)); }
  else p.beta = commit1(g1, dot, r_beta);
  T.append_point("beta", p.beta.b);
  Fq c = T.challenge_scalar("c");
  p.z.resize(n);
  for (size_t i = 0; i < n; i++) p.z[i] = c * x_vec[i] + d_vec[i];
  p.z_delta = c * blind_x + r_delta;
  p.z_beta = c * blind_y + r_beta;
  return p;
}

std::vector<Fq> host_eq_evals(const std::vector<Fq>& r) {  // EqPolynomial::evals (dense_mlpoly.rs:68-84)
  size_t ell = r.size();
  std::vector<Fq> ev((size_t)1 << ell, Fq::one());
  size_t size = 1;
  for (size_t j = 0; j < ell; j++) {
    size *= 2;
    for (size_t i = size - 1;; i -= 2) {
      Fq s = ev[i / 2];
      ev[i] = s * r[j];
      ev[i - 1] = s - ev[i];
      if (i == 1) break;
    }
  }
  return ev;
}

// ================================================================================================ ZK sumcheck on the device
// prove_quad (sumcheck.rs:428-586) and prove_cubic_with_additive_term (sumcheck.rs:588-776).
// tables: NT device arrays of length 2^num_rounds, folded in place (their first element holds the final evaluation).
// Everything of a ZK sumcheck that depends on the random tape only (RandomTape is independent of the transcript): the blinds of the round polynomials and
// evaluations, the per-round randomness of DotProductProof::prove (nizk/mod.rs:329-331) in the reference's order, and the commitments to them —
//   delta_j = commit(d_vec_j; r_delta_j) and the blind halves blinds_poly[j]*h, blinds_evals[j]*h, r_beta_j*h — in four batched device MSMs.
// enqueue() draws and launches (no synchronisation: the results travel to pinned host memory), finish() converts them once the stream has been synchronised;
// r1cs_prove enqueues phase one's behind the witness commitment's MSM, so that neither the draws nor the launches sit on the critical path.
struct ZkPre {
  size_t num_rounds = 0, nco = 0;
  std::vector<Fq> blinds_poly, blinds_evals;
  std::vector<DotPre> pre;
  std::vector<hge> bp_h, be_h;
  DevBuf<u256> d_s, d_b;
  DevBuf<ge> d_pts;
  DevBuf<uint8_t> d_c;
  uint8_t* stage = nullptr;   // pinned: 32*num_rounds encodings, then 3*num_rounds points
  bool enqueued = false, finished = false;
  void enqueue(Ctx& ctx, int degree, size_t rounds, const CommitKey& g1, const CommitKey& gn, RandomTape& tape) {
    num_rounds = rounds; nco = (size_t)degree + 1;
    blinds_poly = tape.random_vector("blinds_poly", num_rounds);
    blinds_evals = tape.random_vector("blinds_evals", num_rounds);
    pre.resize(num_rounds);
    std::vector<Fq> dmat(num_rounds * nco), rdel(num_rounds), rbet(num_rounds);
    for (size_t j = 0; j < num_rounds; j++) {
      pre[j].d_vec = tape.random_vector("d_vec", nco);
      pre[j].r_delta = tape.random_scalar("r_delta");
      pre[j].r_beta = tape.random_scalar("r_beta");
      for (size_t i = 0; i < nco; i++) dmat[j * nco + i] = pre[j].d_vec[i];
      rdel[j] = pre[j].r_delta; rbet[j] = pre[j].r_beta;
    }
    const size_t stage_bytes = num_rounds * (32 + 3 * sizeof(ge));
    if (stage_bytes > (96u << 10)) throw std::runtime_error("spartan_b200: ZK sumcheck with too many rounds for the staging area");
    stage = ctx.pinned + (384 << 10);
    d_s.alloc(num_rounds * nco); d_b.alloc(4 * num_rounds); d_pts.alloc(4 * num_rounds); d_c.alloc(32 * num_rounds);
    // pageable sources: the copies snapshot them before returning
    dev::h2d(d_s.p, dmat.data(), dmat.size() * sizeof(u256), ctx.stream);
    dev::h2d(d_b.p, rdel.data(), num_rounds * sizeof(u256), ctx.stream);
    dev::h2d(d_b.p + num_rounds, rbet.data(), num_rounds * sizeof(u256), ctx.stream);
    dev::h2d(d_b.p + 2 * num_rounds, blinds_poly.data(), num_rounds * sizeof(u256), ctx.stream);
    dev::h2d(d_b.p + 3 * num_rounds, blinds_evals.data(), num_rounds * sizeof(u256), ctx.stream);
    if (ctx.scratch.n < dev::msm_scratch_bytes(num_rounds, nco) + 64) ctx.ensure_scratch(dev::msm_scratch_bytes(num_rounds, nco) + 64);
    const GenSet& gs = *gn.set;
    dev::msm_rows(d_pts.p, gs.table.p, gs.wbits, d_s.p, nco, num_rounds, nco, d_b.p, gn.h, ctx.scratch.p, ctx.stream);                    // delta_j
    dev::compress_batch(d_c.p, d_pts.p, num_rounds, ctx.stream);
    dev::msm_rows(d_pts.p + num_rounds, gs.table.p, gs.wbits, d_s.p, 0, num_rounds, 0, d_b.p + num_rounds, g1.h, ctx.scratch.p, ctx.stream);      // r_beta_j * h
    dev::msm_rows(d_pts.p + 2 * num_rounds, gs.table.p, gs.wbits, d_s.p, 0, num_rounds, 0, d_b.p + 2 * num_rounds, gn.h, ctx.scratch.p, ctx.stream);  // blinds_poly[j] * h
    dev::msm_rows(d_pts.p + 3 * num_rounds, gs.table.p, gs.wbits, d_s.p, 0, num_rounds, 0, d_b.p + 3 * num_rounds, g1.h, ctx.scratch.p, ctx.stream);  // blinds_evals[j] * h
    dev::d2h(stage, d_c.p, 32 * num_rounds, ctx.stream);
    dev::d2h(stage + 32 * num_rounds, d_pts.p + num_rounds, 3 * num_rounds * sizeof(ge), ctx.stream);
    enqueued = true;
  }
  void finish(Ctx& ctx) {   // the stream must have been synchronised after enqueue()
    if (finished) return;
    bp_h.resize(num_rounds); be_h.resize(num_rounds);
    const ge* pts = reinterpret_cast<const ge*>(stage + 32 * num_rounds);
    for (size_t j = 0; j < num_rounds; j++) {
      memcpy(pre[j].delta.b, stage + 32 * j, 32);
      pre[j].r_beta_h = to_hge(pts[j]); bp_h[j] = to_hge(pts[num_rounds + j]); be_h[j] = to_hge(pts[2 * num_rounds + j]);
    }
    d_s.release(); d_b.release(); d_pts.release(); d_c.release();
    finished = true;
  }
};

// sharded: `tables_in` are this rank's cyclic shards (length 2^num_rounds / world).  The first rounds then run on the shards — the fused
// kernels exchange their partial sums over NVLink before handing the round's evaluations to the host — until the local tables drop below
// streaming size; the shards are then all-gathered into replicated tables (in the window) and the latency-bound tail runs on every rank alike.
static void zk_sumcheck_prove(Ctx& ctx, dev::ScKind kind, const Fq& claim, const Fq& blind_claim, size_t num_rounds, u256* const* tables_in, int nt,
                              const CommitKey& g1, const CommitKey& gn, Transcript& T, RandomTape& tape, ZKSumcheckInstanceProof& proof,
                              std::vector<Fq>& r, std::vector<Fq>& finals, Fq& blind_post, bool sharded = false, ZkPre* pre_in = nullptr) {
  std::vector<u256*> tabs(tables_in, tables_in + nt);
  u256* const* tables = tabs.data();
  bool sh = sharded && ctx.shard_world() > 1;
  const int W = sh ? ctx.shard_world() : 1;
  const int degree = kind == dev::SC_QUAD ? 2 : 3;
  ZkPre pre_local;
  ZkPre& zp = pre_in ? *pre_in : pre_local;
  if (!zp.enqueued) zp.enqueue(ctx, degree, num_rounds, g1, gn, tape);
  if (zp.num_rounds != num_rounds || zp.nco != (size_t)degree + 1) throw std::runtime_error("spartan_b200: ZK sumcheck pre-computation of the wrong shape");
  if (!zp.finished) { ctx.sync(); zp.finish(ctx); }
  const std::vector<Fq>& blinds_poly = zp.blinds_poly;
  const std::vector<Fq>& blinds_evals = zp.blinds_evals;
  const std::vector<DotPre>& pre = zp.pre;
  const std::vector<hge>& bp_h = zp.bp_h;
  const std::vector<hge>& be_h = zp.be_h;
  Fq claim_per_round = claim;
  Cp comm_claim_per_round = commit1(g1, claim_per_round, blind_claim);
  auto commit_poly_pre = [&](const std::vector<Fq>& coeffs, size_t j) {   // commit(coeffs; blinds_poly[j]; gens_n) = sum coeff_i*G_i + (blinds_poly[j]*h)
    std::vector<Term> t;
    for (size_t i = 0; i < coeffs.size(); i++) t.push_back({gn.off + i, coeffs[i]});
    return compress(hge_add(host_commit(*gn.set, t.data(), t.size()), bp_h[j]));
  };
  dev::ScInst inst;
  for (int t = 0; t < 4; t++) inst.t[t] = t < nt ? tables[t] : nullptr;
  inst.c_out = inst.t[2];
  inst.write_c = 1;
  u256* d_out = ctx.small.p + 0;     // 3 result scalars
  size_t len = ((size_t)1 << num_rounds) / W;   // current length of the tables this rank holds
  dev::HostSig sig = ctx.next_sig();
  dev::sc_eval(kind, &inst, 1, len, d_out, ctx.red.p, ctx.stream, sig, sh ? ctx.comm->next_xr() : dev::XRank());
  for (size_t j = 0; j < num_rounds; j++) {
    Fq e[3];
    FineTimer fw(ctx, "zk wait evals");
    ctx.wait_sig(sig);
    fw.stop();
    FineTimer fh(ctx, "zk host commit_poly+challenge");
    memcpy(e, ctx.host_res, sizeof e);
    std::vector<Fq> evals = {e[0], claim_per_round - e[0], e[1]};
    if (degree == 3) evals.push_back(e[2]);
    UniPoly poly = UniPoly::from_evals(evals);
    Cp comm_poly = commit_poly_pre(poly.coeffs, j);
    T.append_point("comm_poly", comm_poly.b);
    proof.comm_polys.push_back(comm_poly);
    Fq r_j = T.challenge_scalar("challenge_nextround");
    // bind the tables to r_j on the device right away (fused with the next round's evaluation); the host continues with the
    // sigma protocol of this round while the kernel runs
    if (j + 1 < num_rounds && sh && len < Ctx::SHARD_MIN_LOCAL) {
      // leave the sharded stage: bind locally, all-gather the shards into replicated tables, evaluate the next round there
      dev::fold_top(tables, nt, len, r_j.m, ctx.stream);
      const size_t glen = (len / 2) * W;
      u256* g = ctx.allgather_cyclic(tables, nt, len / 2);
      for (int t = 0; t < nt; t++) { tabs[t] = g + (size_t)t * glen; inst.t[t] = tabs[t]; }
      inst.c_out = inst.t[2];
      sh = false;
      len = glen;
      sig = ctx.next_sig();
      dev::sc_eval(kind, &inst, 1, len, d_out, ctx.red.p, ctx.stream, sig);
    } else {
      if (j + 1 < num_rounds) { sig = ctx.next_sig(); dev::sc_fold_eval(kind, &inst, 1, len, r_j.m, d_out, ctx.red.p, ctx.stream, sig, sh ? ctx.comm->next_xr() : dev::XRank()); }
      else dev::fold_top(tables, nt, len, r_j.m, ctx.stream);
      len >>= 1;
    }
    fh.stop();
    FineTimer fs(ctx, "zk host sigma (overlaps kernel)");

    Fq eval = poly.evaluate(r_j);
    Term te[1] = {{g1.off, eval}};
    Cp comm_eval = compress(hge_add(host_commit(*g1.set, te, 1), be_h[j]));   // eval*G + blinds_evals[j]*h
    T.append_point("comm_claim_per_round", comm_claim_per_round.b);
    T.append_point("comm_eval", comm_eval.b);
    std::vector<Fq> w = T.challenge_vector("combine_two_claims_to_one", 2);
    Fq target = w[0] * claim_per_round + w[1] * eval;
    const Fq& blind_sc = j == 0 ? blind_claim : blinds_evals[j - 1];
    Fq blind = w[0] * blind_sc + w[1] * blinds_evals[j];
    // (the reference also recomputes comm_target from the two decompressed commitments and asserts equality, sumcheck.rs:531/:722;
    //  it is a self-check with no effect on the transcript)
    std::vector<Fq> a(degree + 1);
    Fq rpow = Fq::one();
    for (int i = 0; i <= degree; i++) {
      Fq a_sc = i == 0 ? Fq::from_u64(2) : Fq::one();
      a[i] = w[0] * a_sc + w[1] * rpow;
      rpow *= r_j;
    }
    proof.proofs.push_back(dotproduct_prove(g1, gn, T, tape, poly.coeffs, blinds_poly[j], a, target, blind, &comm_poly, &pre[j]));
    claim_per_round = eval;
    comm_claim_per_round = comm_eval;
    r.push_back(r_j);
    proof.comm_evals.push_back(comm_claim_per_round);
  }
  finals.resize(nt);
  for (int t = 0; t < nt; t++) dev::d2h(ctx.pinned + 32 * t, tables[t], 32, ctx.stream);
  ctx.sync();
  memcpy(finals.data(), ctx.pinned, 32 * nt);
  blind_post = blinds_evals[num_rounds - 1];
}

// ================================================================================================ polynomial commitment / evaluation proof
Cp commit_rows_and_compress(Ctx& ctx, const CommitKey& key, const u256* d_scalars, size_t stride, size_t L, size_t R, const Fq* blinds,
                            std::vector<Cp>& out, const std::function<const Fq*()>& blinds_late) {
  // DensePolynomial::commit_inner (dense_mlpoly.rs:148-177): C_i = (MSM(Z[iR..(i+1)R], G) + blinds[i]*h).compress()
  // blinds_late (instead of blinds): called AFTER the rows' MSM is in flight, so that the host draws the blinds from the random tape (0.8 us per
  // scalar of Keccak) while the device works; the blind terms blinds[i]*h are then a second, tiny launch added onto the rows.
  if (key.off != 0 || R > key.n) throw std::runtime_error("spartan_b200: commit_rows key mismatch");
  const bool blind_on_host = !blinds || blinds[0].is_zero() || key.set->host_tab.count(key.h) != 0;   // the blind term needs the host copy of h's table
  if (L == 1 && R >= 2 && R % 2 == 0 && !blinds_late && ctx.shard_world() == 1 && blind_on_host) {
    // one row (the Cx commitment of DotProductProofLog, nizk/mod.rs:466): a latency problem.  The inner-product round kernel with a = (1, 1) and
    // blocks of two generators returns sum_{j odd} x_j G_j and sum_{j even} x_j G_j straight to the host; the blind term, the two additions and the
    // encoding happen there (5 us instead of a 265-product chain on one GPU thread): ~60 us instead of ~200 us for msm_rows + reduce + k_compress + copy
    ctx.ensure_scratch(std::max(dev::msm_scratch_bytes(1, R), dev::ipa_msm_scratch_points(R, key.set->wbits) * sizeof(ge)) + 64);
    dev::HostSig sg = ctx.next_sig();
    sg.host_out = ctx.host_res + 8;
    DevBuf<ge> pts(2);   // the kernel's device-side copy of the two sums (the host reads its own copy); released after the wait below
    dev::ipa_msm(pts.p, key.set->table.p, key.set->wbits, ctx.ones(), d_scalars, 2, R, ctx.scratch.p, ctx.sig_done.p + 1, ctx.stream, sg);
    hge acc = hge_identity();
    if (blinds && !blinds[0].is_zero()) { Term t[1] = {{key.h, blinds[0]}}; acc = host_commit(*key.set, t, 1); }
    ctx.wait_sig(sg);
    ge odd, even;
    memcpy(&odd, ctx.host_res + 8, sizeof(ge)); memcpy(&even, ctx.host_res + 12, sizeof(ge));
    out.resize(1);
    out[0] = compress(hge_add(hge_add(to_hge(odd), to_hge(even)), acc));
    return out[0];
  }
  ctx.ensure_scratch(dev::msm_scratch_bytes(L, R) + 64);
  DevBuf<ge> rows(L), bh;
  DevBuf<u256> d_bl;
  if (blinds) { d_bl.alloc(L); dev::h2d(d_bl.p, blinds, L * sizeof(u256), ctx.stream); }
  DevBuf<uint8_t> comp(32 * L);
  out.resize(L);
  const size_t W = (size_t)ctx.shard_world();
  // rows are independent MSMs over the same generators: when sharded, rank r commits rows [r*L/W, (r+1)*L/W) and the 32-byte encodings are all-gathered
  const bool split = W > 1 && L >= 2 * W && L % W == 0;
  const size_t Lr = split ? L / W : L, row0 = split ? (size_t)ctx.rank() * Lr : 0;
  dev::msm_rows(rows.p, key.set->table.p, key.set->wbits, d_scalars + row0 * stride, stride, Lr, R, blinds ? d_bl.p + row0 : nullptr, key.h, ctx.scratch.p, ctx.stream);
  if (!blinds && blinds_late) {
    const Fq* bl = blinds_late();
    d_bl.alloc(L); bh.alloc(Lr);
    dev::h2d(d_bl.p, bl, L * sizeof(u256), ctx.stream);
    dev::msm_rows(bh.p, key.set->table.p, key.set->wbits, d_scalars, 0, Lr, 0, d_bl.p + row0, key.h, ctx.scratch.p, ctx.stream);   // blinds[i]*h (rows without generator terms)
    dev::add_points(rows.p, bh.p, Lr, ctx.stream);
  }
  dev::compress_batch(comp.p, rows.p, Lr, ctx.stream);
  if (split) dev::d2h(out.data(), ctx.allgather_block(comp.p, 32 * Lr), 32 * L, ctx.stream);
  else dev::d2h(out.data(), comp.p, 32 * L, ctx.stream);
  ctx.sync();
  return out.empty() ? Cp() : out[0];
}

void append_poly_commitment(Transcript& T, const char* label, const PolyCommitment& c) {  // dense_mlpoly.rs:292-300
  T.append_message(label, "poly_commitment_begin");
  for (auto& p : c.C) T.append_point("poly_commitment_share", p.b);
  T.append_message(label, "poly_commitment_end");
}

// BulletReductionProof::prove (nizk/bullet.rs:32-132) with Q = r*G1 and H = h folded into host fixed-base terms, and the
// generator vector left unfolded (see k_ipa_lr in kernels.cu).  d_a / d_b are consumed (folded in place).
static void bullet_prove(Ctx& ctx, Transcript& T, const PolyCommitmentGens& gens, const std::function<Fq()>& get_r_scale, u256* d_a, u256* d_b, size_t n,
                         const Fq& blind, const std::vector<std::pair<Fq, Fq>>& blinds_vec, BulletReductionProof& proof, Fq& a_hat, Fq& b_hat, hge& g_hat,
                         Fq& blind_final) {
  const GenSet& gs = *gens.gens_n.set;
  const size_t rounds = blinds_vec.size();
  DevBuf<u256> svec(n), lr(2 * n), d_bl(2 * rounds + 1);
  DevBuf<ge> pts(2), d_bh(2 * rounds + 1);
  dev::fill_one(svec.p, n, ctx.stream);
  ctx.ensure_scratch(dev::msm_scratch_bytes(2 * rounds + 2, n) + 64);
  // blind_L[k]*H and blind_R[k]*H depend on the tape only: one batched launch up front (rows with no generator terms, just the blind)
  std::vector<ge> bh(2 * rounds);
  if (rounds) {
    std::vector<Fq> bl(2 * rounds);
    for (size_t k = 0; k < rounds; k++) { bl[2 * k] = blinds_vec[k].first; bl[2 * k + 1] = blinds_vec[k].second; }
    dev::h2d(d_bl.p, bl.data(), bl.size() * sizeof(u256), ctx.stream);
    dev::msm_rows(d_bh.p, gs.table.p, gs.wbits, lr.p, 0, 2 * rounds, 0, d_bl.p, gens.gens_n.h, ctx.scratch.p, ctx.stream);
    dev::d2h(bh.data(), d_bh.p, bh.size() * sizeof(ge), ctx.stream);   // pageable target: completes before the call returns
  }
  u256* d_c = ctx.small.p + 16;  // c_L, c_R
  blind_final = blind;
  Fq r_scale = Fq::zero();
  size_t cur = n, k = 0;
  while (cur != 1) {
    size_t half = cur / 2;
    FineTimer f1(ctx, "ipa launch");
    const u256* da[2] = {d_a, d_a + half};
    const u256* db[2] = {d_b + half, d_b};
    // c_L = <a_L, b_R>, c_R = <a_R, b_L> (bullet.rs:78-79) and L = <a_L, G_R>, R = <a_R, G_L> over the unfolded generators (scalars a[.]*s[j] formed
    // inside the kernel), all published to the host.  With the quad-lane kernel the dot products ride in the MSM's launch (two extra blocks with
    // their own flag word); otherwise a dot_pairs launch precedes the MSM.
    dev::HostSig sig, sig2;
    // while a background MSM owns most SMs (partitioned stream2), two blocks per free SM is all that can start at once: cap the grid, the kernel strides
    const int ipa_cap = ctx.bg_busy && ctx.stream2_sms > 0 ? 2 * std::max(4, dev::sm_count() - ctx.stream2_sms) : 0;
    if (dev::ipa_msm_fuses_dots()) {
      sig.host_out = ctx.host_res; sig.flag = ctx.host_flag + 16; sig.done = ctx.sig_done.p + 2; sig.seq = ++ctx.sigc_seq;
      sig2 = ctx.next_sig();
      sig2.host_out = ctx.host_res + 8;
      dev::ipa_msm(pts.p, gs.table.p, gs.wbits, d_a, svec.p, cur, n, ctx.scratch.p, ctx.sig_done.p + 1, ctx.stream, sig2, ipa_cap, d_b, d_c, sig);
    } else {
      sig = ctx.next_sig();
      dev::dot_pairs(d_c, da, db, 2, half, ctx.red.p, ctx.stream, sig);
      sig2 = ctx.next_sig();
      sig2.host_out = ctx.host_res + 8;
      dev::ipa_msm(pts.p, gs.table.p, gs.wbits, d_a, svec.p, cur, n, ctx.scratch.p, ctx.sig_done.p + 1, ctx.stream, sig2, ipa_cap);
    }
    f1.stop();
    // the transcript work that precedes the first round (absorbing a_vec, deriving r) runs while the device computes round 0
    if (k == 0) r_scale = get_r_scale();
    FineTimer f2(ctx, "ipa wait c");
    ctx.wait_sig(sig);
    f2.stop();
    FineTimer f3(ctx, "ipa host c*Q (overlaps MSM)");
    Fq c_L, c_R;
    memcpy(&c_L, ctx.host_res, 32); memcpy(&c_R, ctx.host_res + 1, 32);
    // + c_L*Q + blind_L*H with Q = r*G1 (nizk/mod.rs:479-480), H = gens_n.h   (bullet.rs:83-97)
    Term tl[1] = {{gens.gens_1.off, c_L * r_scale}};
    Term tr[1] = {{gens.gens_1.off, c_R * r_scale}};
    hge hl = hge_add(host_commit(gs, tl, 1), to_hge(bh[2 * k]));
    hge hr = hge_add(host_commit(gs, tr, 1), to_hge(bh[2 * k + 1]));
    f3.stop();
    FineTimer f4(ctx, "ipa wait MSM");
    ctx.wait_sig(sig2);
    f4.stop();
    FineTimer f5(ctx, "ipa host compress+transcript");
    ge Lg, Rg;
    memcpy(&Lg, ctx.host_res + 8, sizeof(ge)); memcpy(&Rg, ctx.host_res + 12, sizeof(ge));
    Cp Lc, Rc;
    compress2(hge_add(to_hge(Lg), hl), hge_add(to_hge(Rg), hr), Lc, Rc);
    T.append_point("L", Lc.b);
    T.append_point("R", Rc.b);
    Fq u = T.challenge_scalar("u");
    Fq u_inv = u.inv();
    f5.stop();
    dev::ipa_fold_update(d_a, d_b, svec.p, half, n, u.m, u_inv.m, ctx.stream);
    blind_final = blind_final + blinds_vec[k].first * u * u + blinds_vec[k].second * u_inv * u_inv;  // bullet.rs:111
    proof.L_vec.push_back(Lc);
    proof.R_vec.push_back(Rc);
    cur = half;
    k++;
  }
  if (k == 0) r_scale = get_r_scale();
  // g_hat = G_final[0] = <s, G>
  FineTimer f6(ctx, "ipa final");
  dev::d2h(ctx.pinned, d_a, 32, ctx.stream);
  dev::d2h(ctx.pinned + 32, d_b, 32, ctx.stream);
  if (n >= 2) {
    // the round kernel with a = (1, 1) and blocks of two generators: "L" = sum over the odd j of s[j] G_j, "R" = over the even j; both reach the host
    // through the kernel's own publication and are added there (one launch instead of msm_rows + reduce + copy)
    dev::HostSig sg = ctx.next_sig();
    sg.host_out = ctx.host_res + 8;
    dev::ipa_msm(pts.p, gs.table.p, gs.wbits, ctx.ones(), svec.p, 2, n, ctx.scratch.p, ctx.sig_done.p + 1, ctx.stream, sg);
    ctx.sync();
    ge odd, even;
    memcpy(&odd, ctx.host_res + 8, sizeof(ge)); memcpy(&even, ctx.host_res + 12, sizeof(ge));
    g_hat = hge_add(to_hge(odd), to_hge(even));
  } else {
    dev::msm_rows(pts.p, gs.table.p, gs.wbits, svec.p, n, 1, n, nullptr, 0, ctx.scratch.p, ctx.stream);
    dev::d2h(ctx.pinned + 64, pts.p, sizeof(ge), ctx.stream);
    ctx.sync();
    ge g;
    memcpy(&g, ctx.pinned + 64, sizeof(ge));
    g_hat = to_hge(g);
  }
  memcpy(&a_hat, ctx.pinned, 32); memcpy(&b_hat, ctx.pinned + 32, 32);
}

// DotProductProofLog::prove (nizk/mod.rs:440-525).  d_x: device x_vec (n, consumed); a_vec on the host (it is absorbed by the transcript).
// d_x: device x_vec (n, consumed); d_avec: device a_vec (n, consumed); a_canon: the canonical bytes of a_vec, which the transcript absorbs
static void dotproduct_log_prove(Ctx& ctx, const PolyCommitmentGens& gens, Transcript& T, RandomTape& tape, u256* d_x, const Fq& blind_x,
                                 u256* d_avec, const std::vector<uint8_t>& a_canon, const Fq& y, const Fq& blind_y, DotProductProofLog& proof, Cp& Cy_out) {
  T.append_protocol_name("dot product proof (log)");
  FineTimer f0(ctx, "dotlog pre-ipa");
  size_t n = a_canon.size() / 32;
  if (gens.n != n) throw std::runtime_error("spartan_b200: DotProductProofLog size mismatch");
  const GenSet& gs = *gens.gens_n.set;
  Fq d = tape.random_scalar("d");
  Fq r_delta = tape.random_scalar("r_delta");
  Fq r_beta = tape.random_scalar("r_delta");  // sic: the reference reuses the label (nizk/mod.rs:459)
  size_t lg_n = 0;
  while (((size_t)1 << lg_n) < n) lg_n++;
  std::vector<Fq> v1 = tape.random_vector("blinds_vec_1", lg_n), v2 = tape.random_vector("blinds_vec_2", lg_n);
  std::vector<std::pair<Fq, Fq>> blinds_vec;
  for (size_t i = 0; i < lg_n; i++) blinds_vec.push_back({v1[i], v2[i]});
  std::vector<Cp> cx;
  commit_rows_and_compress(ctx, gens.gens_n, d_x, n, 1, n, &blind_x, cx);
  T.append_point("Cx", cx[0].b);
  Cy_out = commit1(gens.gens_1, y, blind_y);
  T.append_point("Cy", Cy_out.b);
  Fq r = Fq::zero();
  auto absorb_a = [&]() {   // called by bullet_prove once round 0 is in flight on the device (same transcript order as nizk/mod.rs:475-477)
    T.append_scalar_bytes("a", a_canon.data(), n);
    r = T.challenge_scalar("r");
    return r;
  };
  Fq x_hat, a_hat, rhat_Gamma;
  hge g_hat;
  f0.stop();
  // blind_Gamma = blind_x + r*blind_y enters the reduction only through the final blind, which is linear in it: pass blind_x and add r*blind_y after
  bullet_prove(ctx, T, gens, absorb_a, d_x, d_avec, n, blind_x, blinds_vec, proof.bullet_reduction_proof, x_hat, a_hat, g_hat, rhat_Gamma);
  rhat_Gamma = rhat_Gamma + r * blind_y;
  FineTimer f6(ctx, "dotlog post-ipa");
  Fq y_hat = x_hat * a_hat;
  // delta = d*g_hat + r_delta*h (gens_hat, nizk/mod.rs:497-505)
  Term th[1] = {{gens.gens_1.h, r_delta}};
  proof.delta = compress(hge_add(hge_scalarmul(d.canonical(), g_hat), host_commit(gs, th, 1)));
  T.append_point("delta", proof.delta.b);
  // beta = d*(r*G1) + r_beta*h (gens_1_scaled, nizk/mod.rs:507)
  proof.beta = commit1(gens.gens_1, d * r, r_beta);
  T.append_point("beta", proof.beta.b);
  Fq c = T.challenge_scalar("c");
  proof.z1 = d + c * y_hat;
  proof.z2 = a_hat * (c * rhat_Gamma + r_beta) + r_delta;
}

// PolyEvalProof::prove (dense_mlpoly.rs:312-365).  d_Z: device table of 2^|r| scalars (read-only).
void polyeval_prove(Ctx& ctx, const u256* d_Z, const std::vector<Fq>* blinds_opt, const std::vector<Fq>& r, const Fq& Zr, const Fq* blind_Zr_opt,
                           const PolyCommitmentGens& gens, Transcript& T, RandomTape& tape, PolyEvalProof& proof, Cp& C_Zr) {
  T.append_protocol_name("polynomial evaluation proof");
  FineTimer fp(ctx, "polyeval pre (eq, bound_rows)");
  size_t ell = r.size(), lv = ell / 2;
  size_t L_size = (size_t)1 << lv, R_size = (size_t)1 << (ell - lv);
  // compute_factored_evals (dense_mlpoly.rs:90-98) on the device; the R half is also needed as canonical bytes (the transcript absorbs a_vec)
  DevBuf<u256> d_r(r.size() + 1), d_L(L_size), d_R(R_size), d_Rc(R_size), d_LZ(R_size), tmp(64 * R_size), eqs(2 * ((size_t)1 << ((ell - lv + 1) / 2)) + 8);
  dev::h2d(d_r.p, r.data(), r.size() * sizeof(u256), ctx.stream);
  dev::eq_evals(d_L.p, d_r.p, (int)lv, eqs.p, ctx.stream);
  dev::eq_evals(d_R.p, d_r.p + lv, (int)(ell - lv), eqs.p, ctx.stream);
  dev::to_canonical(d_Rc.p, d_R.p, R_size, ctx.stream);
  std::vector<uint8_t> a_canon(32 * R_size);
  dev::d2h(a_canon.data(), d_Rc.p, 32 * R_size, ctx.stream);
  dev::bound_rows(d_LZ.p, d_Z, d_L.p, L_size, R_size, tmp.p, ctx.stream);  // DensePolynomial::bound (dense_mlpoly.rs:206-213)
  Fq LZ_blind = Fq::zero();
  if (blinds_opt) {
    std::vector<Fq> Lev = ctx.download(d_L.p, L_size);
    for (size_t i = 0; i < L_size; i++) LZ_blind += (*blinds_opt)[i] * Lev[i];
  }
  ctx.sync();
  Fq blind_Zr = blind_Zr_opt ? *blind_Zr_opt : Fq::zero();
  fp.stop();
  dotproduct_log_prove(ctx, gens, T, tape, d_LZ.p, LZ_blind, d_R.p, a_canon, Zr, blind_Zr, proof.proof, C_Zr);
}

// ================================================================================================ instance
static void build_compressed(size_t nmajor, const std::vector<uint32_t>& major, const std::vector<uint32_t>& minor, const std::vector<Fq>& val,
                             std::vector<uint32_t>& ptr, std::vector<uint32_t>& idx, std::vector<Fq>& v) {
  size_t nnz = major.size();
  ptr.assign(nmajor + 1, 0);
  for (size_t k = 0; k < nnz; k++) ptr[major[k] + 1]++;
  for (size_t i = 0; i < nmajor; i++) ptr[i + 1] += ptr[i];
  std::vector<uint32_t> fill(ptr.begin(), ptr.end() - 1);
  idx.resize(nnz); v.resize(nnz);
  for (size_t k = 0; k < nnz; k++) { uint32_t p = fill[major[k]]++; idx[p] = minor[k]; v[p] = val[k]; }
}
template <class T>
static void up(Ctx* ctx, DevBuf<T>& d, const std::vector<T>& h) { d.alloc(h.size()); dev::h2d(d.p, h.data(), h.size() * sizeof(T), ctx->stream); }
static void up(Ctx* ctx, DevBuf<u256>& d, const std::vector<Fq>& h) { d.alloc(h.size()); dev::h2d(d.p, h.data(), h.size() * sizeof(u256), ctx->stream); }

std::vector<uint8_t> Instance::shape_bincode() const {
  // bincode(R1CSShape{num_cons,num_vars,num_inputs,A,B,C}), SparseMatPolynomial{num_vars_x,num_vars_y,M:Vec<{row,col,val}>} (r1cs.rs:19-26, sparse_mlpoly.rs:19-37)
  Writer w;
  w.u64(num_cons); w.u64(num_vars); w.u64(num_inputs);
  size_t nx = 0, ny = 0;
  while (((size_t)1 << nx) < num_cons) nx++;
  while (((size_t)1 << ny) < 2 * num_vars) ny++;
  for (int m = 0; m < 3; m++) {
    w.out.reserve(w.out.size() + 24 + 48 * M[m].row.size());
    w.u64(nx); w.u64(ny); w.u64(M[m].row.size());
    for (size_t k = 0; k < M[m].row.size(); k++) { w.u64(M[m].row[k]); w.u64(M[m].col[k]); w.scalar(M[m].val[k]); }
  }
  return std::move(w.out);
}
const std::vector<uint8_t>& Instance::shape_digest() const {
  if (digest.empty()) { std::vector<uint8_t> raw = shape_bincode(); digest = miniz_zlib_level6(raw.data(), raw.size()); }
  return digest;
}

void Instance::finalize(Ctx* ctx) {
  size_t ncols = 2 * num_vars;
  for (int m = 0; m < 3; m++) {
    SparseMatDev& M_ = M[m];
    std::vector<uint32_t> ptr, idx;
    std::vector<Fq> v;
    build_compressed(num_cons, M_.row, M_.col, M_.val, ptr, idx, v);
    up(ctx, M_.csr_ptr, ptr); up(ctx, M_.csr_idx, idx); up(ctx, M_.csr_val, v);
    build_compressed(ncols, M_.col, M_.row, M_.val, ptr, idx, v);
    up(ctx, M_.csc_ptr, ptr); up(ctx, M_.csc_idx, idx); up(ctx, M_.csc_val, v);
    up(ctx, M_.coo_row, M_.row); up(ctx, M_.coo_col, M_.col); up(ctx, M_.coo_val, M_.val);
    ctx->sync();
  }
}

// ================================================================================================ R1CSProof::prove
struct PhaseTimer {
  Ctx& ctx; const char* name; std::chrono::steady_clock::time_point t0;
  PhaseTimer(Ctx& c, const char* n) : ctx(c), name(n), t0(std::chrono::steady_clock::now()) {}
  ~PhaseTimer() { ctx.timings.push_back({name, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count()}); }
};

void r1cs_prove(Ctx& ctx, const Instance& inst, const u256* d_vars, const std::vector<Fq>& input, const R1CSGens& gens, Transcript& T,
                RandomTape& tape, R1CSProof& proof, std::vector<Fq>& rx, std::vector<Fq>& ry, const R1csHooks* hooks) {
  PhaseTimer t_all(ctx, "R1CSProof::prove");
  T.append_protocol_name("R1CS proof");
  const size_t num_vars = inst.num_vars, num_cons = inst.num_cons;
  if (!(input.size() < num_vars)) throw std::runtime_error("spartan_b200: |input| + 1 must be at most the number of variables");  // r1csproof.rs:156
  T.append_scalars("input", input);

  size_t ell = 0;
  while (((size_t)1 << ell) < num_vars) ell++;
  const size_t L_size = (size_t)1 << (ell / 2), R_size = (size_t)1 << (ell - ell / 2);
  std::vector<Fq> blinds_vars;
  {
    PhaseTimer t(ctx, "polycommit");
    // dense_mlpoly.rs:193-196; the blinds are the tape's first draw and are made while the rows' MSM already runs
    // (queueing phase one's tape-only commitments here as well — ZkPre::enqueue behind the rows' MSM — was measured and lost 0.3 ms: their small kernels
    //  lengthen this commitment's chain by what they save phase one, whose rounds are host-bound anyway; profiles/r02_tuning.md section 10)
    commit_rows_and_compress(ctx, gens.gens_pc.gens_n, d_vars, R_size, L_size, R_size, nullptr, proof.comm_vars.C,
                             [&]() { blinds_vars = tape.random_vector("poly_blinds", L_size); return blinds_vars.data(); });
    append_poly_commitment(T, "poly_commitment", proof.comm_vars);
  }

  // z = vars || 1 || input || 0...                                          (r1csproof.rs:177-185)
  const size_t zlen = 2 * num_vars;
  DevBuf<u256> d_z(zlen);
  size_t num_rounds_x = 0, num_rounds_y = 0;
  while (((size_t)1 << num_rounds_x) < num_cons) num_rounds_x++;
  while (((size_t)1 << num_rounds_y) < zlen) num_rounds_y++;
  const int W = ctx.shard_world(), rk = ctx.rank();
  int logW = 0;
  while ((1 << logW) < W) logW++;
  const bool sh1 = ctx.shard_table(num_cons), sh2 = ctx.shard_table(zlen);   // sumcheck tables as cyclic shards (rank r holds the indices = r mod W)
  const size_t n1 = sh1 ? num_cons / W : num_cons, n2 = sh2 ? zlen / W : zlen;
  DevBuf<u256> d_tau(num_cons), d_Az(n1), d_Bz(n1), d_Cz(n1), d_chal(64), eq_small(2 * ((size_t)1 << ((std::max(num_rounds_x, num_rounds_y) + 1) / 2)) + 8);
  Fq blind_claim_postsc1;
  std::vector<Fq> claims1;
  {
    PhaseTimer t(ctx, "prove_sc_phase_one");
    dev::d2d(d_z.p, d_vars, num_vars * sizeof(u256), ctx.stream);
    std::vector<Fq> tail(1 + input.size());
    tail[0] = Fq::one();
    for (size_t i = 0; i < input.size(); i++) tail[1 + i] = input[i];
    dev::h2d(d_z.p + num_vars, tail.data(), tail.size() * sizeof(u256), ctx.stream);
    dev::dzero(d_z.p + num_vars + tail.size(), (num_vars - tail.size()) * sizeof(u256), ctx.stream);
    std::vector<Fq> tau = T.challenge_vector("challenge_tau", num_rounds_x);
    dev::h2d(d_chal.p, tau.data(), tau.size() * sizeof(u256), ctx.stream);
    // inst.multiply_vec (r1cs.rs:268-282)
    u256* outs[3] = {d_Az.p, d_Bz.p, d_Cz.p};
    if (sh1) {
      // this rank's slice of eq(tau, .): eq over the leading variables times the factor its low index bits fix; its rows of A z, B z, C z
      dev::eq_evals(d_tau.p, d_chal.p, (int)num_rounds_x - logW, eq_small.p, ctx.stream);
      dev::scale(d_tau.p, shard_eq_scale(tau, W, rk).m, n1, ctx.stream);
      for (int m = 0; m < 3; m++) dev::spmv_cyclic(outs[m], n1, rk, W, inst.M[m].csr_ptr.p, inst.M[m].csr_idx.p, inst.M[m].csr_val.p, d_z.p, ctx.stream);
    } else {
      dev::eq_evals(d_tau.p, d_chal.p, (int)num_rounds_x, eq_small.p, ctx.stream);
      for (int m = 0; m < 3; m++) dev::spmv(outs[m], num_cons, inst.M[m].csr_ptr.p, inst.M[m].csr_idx.p, inst.M[m].csr_val.p, d_z.p, ctx.stream);
    }
    u256* tabs[4] = {d_tau.p, d_Az.p, d_Bz.p, d_Cz.p};
    zk_sumcheck_prove(ctx, dev::SC_CUBIC4, Fq::zero(), Fq::zero(), num_rounds_x, tabs, 4, gens.gens_1, gens.gens_4, T, tape, proof.sc_proof_phase1, rx, claims1,
                      blind_claim_postsc1, sh1);
  }
  if (hooks && hooks->on_rx) hooks->on_rx(rx);
  const Fq tau_claim = claims1[0], Az_claim = claims1[1], Bz_claim = claims1[2], Cz_claim = claims1[3];
  Fq Az_blind = tape.random_scalar("Az_blind"), Bz_blind = tape.random_scalar("Bz_blind"), Cz_blind = tape.random_scalar("Cz_blind"),
     prod_Az_Bz_blind = tape.random_scalar("prod_Az_Bz_blind");
  Cp comm_Cz_claim, comm_Az_claim, comm_Bz_claim, comm_prod;
  proof.pok_Cz = knowledge_prove(gens.gens_1, T, tape, Cz_claim, Cz_blind, comm_Cz_claim);
  Fq prod = Az_claim * Bz_claim;
  proof.proof_prod = product_prove(gens.gens_1, T, tape, Az_claim, Az_blind, Bz_claim, Bz_blind, prod, prod_Az_Bz_blind, comm_Az_claim, comm_Bz_claim, comm_prod);
  T.append_point("comm_Az_claim", comm_Az_claim.b);
  T.append_point("comm_Bz_claim", comm_Bz_claim.b);
  T.append_point("comm_Cz_claim", comm_Cz_claim.b);
  T.append_point("comm_prod_Az_Bz_claims", comm_prod.b);
  proof.claims_phase2 = {comm_Az_claim, comm_Bz_claim, comm_Cz_claim, comm_prod};
  Fq blind_expected_claim_postsc1 = tau_claim * (prod_Az_Bz_blind - Cz_blind);
  Fq claim_post_phase1 = (Az_claim * Bz_claim - Cz_claim) * tau_claim;
  proof.proof_eq_sc_phase1 = equality_prove(gens.gens_1, T, tape, claim_post_phase1, blind_expected_claim_postsc1, claim_post_phase1, blind_claim_postsc1);

  Fq blind_claim_postsc2;
  std::vector<Fq> claims2;
  {
    PhaseTimer t(ctx, "prove_sc_phase_two");
    Fq rabc[3] = {T.challenge_scalar("challenge_Az"), T.challenge_scalar("challenge_Bz"), T.challenge_scalar("challenge_Cz")};
    Fq claim_phase2 = rabc[0] * Az_claim + rabc[1] * Bz_claim + rabc[2] * Cz_claim;
    Fq blind_claim_phase2 = rabc[0] * Az_blind + rabc[1] * Bz_blind + rabc[2] * Cz_blind;
    // evals_rx = eq(rx, .), then the three transposed SpMVs of compute_eval_table_sparse (r1cs.rs:284-298), then r_A*A + r_B*B + r_C*C
    dev::h2d(d_chal.p, rx.data(), rx.size() * sizeof(u256), ctx.stream);
    dev::eq_evals(d_tau.p, d_chal.p, (int)num_rounds_x, eq_small.p, ctx.stream);
    DevBuf<u256> eA(n2), eB(n2), eC(n2), d_ABC(n2), d_zloc;
    u256* outs[3] = {eA.p, eB.p, eC.p};
    for (int m = 0; m < 3; m++) {
      if (sh2) dev::spmv_cyclic(outs[m], n2, rk, W, inst.M[m].csc_ptr.p, inst.M[m].csc_idx.p, inst.M[m].csc_val.p, d_tau.p, ctx.stream);   // this rank's columns
      else dev::spmv(outs[m], zlen, inst.M[m].csc_ptr.p, inst.M[m].csc_idx.p, inst.M[m].csc_val.p, d_tau.p, ctx.stream);
    }
    dev::h2d(d_chal.p + 32, rabc, 3 * sizeof(u256), ctx.stream);
    dev::lincomb3(d_ABC.p, eA.p, eB.p, eC.p, d_chal.p + 32, n2, ctx.stream);
    if (sh2) { d_zloc.alloc(n2); dev::take_cyclic(d_zloc.p, d_z.p, n2, rk, W, ctx.stream); }
    u256* tabs[2] = {sh2 ? d_zloc.p : d_z.p, d_ABC.p};
    zk_sumcheck_prove(ctx, dev::SC_QUAD, claim_phase2, blind_claim_phase2, num_rounds_y, tabs, 2, gens.gens_1, gens.gens_3, T, tape, proof.sc_proof_phase2, ry, claims2,
                      blind_claim_postsc2, sh2);
  }
  if (hooks && hooks->on_ry) hooks->on_ry(ry);
  {
    PhaseTimer t(ctx, "polyeval");
    // eval_vars_at_ry = poly_vars.evaluate(ry[1..]) (dense_mlpoly.rs:236-242)
    std::vector<Fq> ry1(ry.begin() + 1, ry.end());
    DevBuf<u256> d_eq(num_vars);
    dev::h2d(d_chal.p, ry1.data(), ry1.size() * sizeof(u256), ctx.stream);
    dev::eq_evals(d_eq.p, d_chal.p, (int)ry1.size(), eq_small.p, ctx.stream);
    dev::dot(ctx.small.p + 32, d_vars, d_eq.p, num_vars, ctx.red.p, ctx.stream);
    Fq eval_vars_at_ry;
    ctx.get_small(32, &eval_vars_at_ry, 1);
    Fq blind_eval = tape.random_scalar("blind_eval");
    polyeval_prove(ctx, d_vars, &blinds_vars, ry1, eval_vars_at_ry, &blind_eval, gens.gens_pc, T, tape, proof.proof_eval_vars_at_ry, proof.comm_vars_at_ry);
    Fq blind_eval_Z_at_ry = (Fq::one() - ry[0]) * blind_eval;
    Fq blind_expected_claim_postsc2 = claims2[1] * blind_eval_Z_at_ry;
    Fq claim_post_phase2 = claims2[0] * claims2[1];
    proof.proof_eq_sc_phase2 = equality_prove(gens.gens_pc.gens_1, T, tape, claim_post_phase2, blind_expected_claim_postsc2, claim_post_phase2, blind_claim_postsc2);
  }
}

void nizk_prove(Ctx& ctx, const Instance& inst, const u256* d_vars, const std::vector<Fq>& input, const R1CSGens& gens, Transcript& T, const Fq& tape_seed,
                NizkProof& out) {
  ctx.timings.clear();
  ShardScope shard(ctx);
  PhaseTimer t(ctx, "NIZK::prove");
  RandomTape tape("proof", tape_seed);                                   // lib.rs:511
  T.append_protocol_name("Spartan NIZK proof");                          // lib.rs:513
  const std::vector<uint8_t>& digest = inst.shape_digest();                // r1cs.rs:154-158; never empty: the transcript always binds the shape
  T.append_message("R1CSShapeDigest", digest.data(), digest.size());       // lib.rs:514
  r1cs_prove(ctx, inst, d_vars, input, gens, T, tape, out.r1cs_sat_proof, out.rx, out.ry);
}

}  // namespace sp
