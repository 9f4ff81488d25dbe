// spartan_b200 — extern "C" boundary (include/spartan_b200.h).  Catches every exception and maps it to a status code.
#include "../../include/spartan_b200.h"
#include <stdlib.h>
#include <algorithm>
#include <sstream>
#include "prover.hpp"
#include "snark.hpp"

using namespace sp;

struct sp_ctx { Ctx c; explicit sp_ctx(int d) : c(d) {} };
struct sp_poly { Ctx* ctx; DevBuf<u256> d; size_t len; };
struct sp_gens { std::unique_ptr<GenSet> set; size_t n; };
struct sp_points { Ctx* ctx; DevBuf<ge_niels> pts; size_t n; mutable DevBuf<uint8_t> scratch; mutable size_t scratch_bytes = 0; };
struct sp_instance { Instance inst; };
struct sp_nizk_gens { std::unique_ptr<R1CSGens> g; };
struct sp_snark_gens { std::unique_ptr<SnarkGens> g; };
struct sp_snark_encoding { std::unique_ptr<SnarkEncoding> e; };

static std::string g_create_error;

#define SP_TRY(ctxp) try {
#define SP_CATCH(ctxp)                                                                           \
  }                                                                                              \
  catch (const SpError& e) { if (ctxp) (ctxp)->c.last_error = e.what(); return e.code; }        \
  catch (const std::exception& e) {                                                              \
    if (ctxp) (ctxp)->c.last_error = e.what();                                                   \
    return std::string(e.what()).find("CUDA") != std::string::npos ? SP_ERR_CUDA : SP_ERR_INTERNAL; \
  }                                                                                              \
  catch (...) { if (ctxp) (ctxp)->c.last_error = "unknown error"; return SP_ERR_INTERNAL; }      \
  return SP_OK;

static Fq fq_in(const uint64_t l[4]) { Fq f; memcpy(&f.m, l, 32); return f; }
static void fq_out(uint64_t l[4], const Fq& f) { memcpy(l, &f.m, 32); }
static std::vector<Fq> fq_vec(const uint64_t* l, size_t n) { std::vector<Fq> v(n); if (n) memcpy(v.data(), l, 32 * n); return v; }
// scalars that enter host arithmetic (public inputs, tape seed) must be reduced residues: limbs >= q are rejected at the boundary
static void require_reduced(const uint64_t* l, size_t n, const char* what) {
  for (size_t i = 0; i < n; i++)
    if (!fq_bytes_canonical(reinterpret_cast<const uint8_t*>(l + 4 * i))) throw SpError(SP_ERR_INVALID_SCALAR, std::string(what) + ": scalar limbs are not reduced modulo q");
}
static uint8_t* dup_bytes(const std::vector<uint8_t>& v) { uint8_t* p = (uint8_t*)malloc(v.size() ? v.size() : 1); memcpy(p, v.data(), v.size()); return p; }

extern "C" {

int sp_device_count(void) { return dev::device_count(); }
int sp_ctx_create(int device, sp_ctx** out) {
  try {
    if (dev::device_count() <= device) { g_create_error = "no CUDA device (spartan_b200 has no CPU fallback)"; return SP_ERR_NO_DEVICE; }
    *out = new sp_ctx(device);
    return SP_OK;
  } catch (const std::exception& e) { g_create_error = e.what(); return SP_ERR_CUDA; }
}
void sp_ctx_destroy(sp_ctx* ctx) { delete ctx; }
const char* sp_last_error(const sp_ctx* ctx) { return ctx ? ctx->c.last_error.c_str() : g_create_error.c_str(); }
unsigned long long sp_kernel_launches(void) { return dev::launch_count(); }
int sp_timings(sp_ctx* ctx, char* buf, size_t buflen) {
  std::ostringstream os;
  for (auto& t : ctx->c.timings) os << t.first << "=" << t.second << ";";
  std::string s = os.str();
  if (s.size() + 1 > buflen) return SP_ERR_INVALID_ARG;
  memcpy(buf, s.c_str(), s.size() + 1);
  return SP_OK;
}

// ---- intra-proof sharding: one process per GPU, windows exchanged as CUDA IPC handles by whatever transport the host has
size_t sp_comm_handle_bytes(void) { return dev::ipc_handle_bytes(); }
int sp_comm_export(sp_ctx* ctx, uint8_t* handle_out) {
  SP_TRY(ctx)
  ctx->c.comm_create();
  if (ctx->c.comm->connected) throw SpError(SP_ERR_INVALID_ARG, "communicator already connected");
  ctx->c.comm->export_handle(handle_out);
  SP_CATCH(ctx)
}
int sp_comm_connect(sp_ctx* ctx, int rank, int world, const uint8_t* handles) {
  SP_TRY(ctx)
  if (!ctx->c.comm) throw SpError(SP_ERR_INVALID_ARG, "sp_comm_export must be called first (it creates the window whose handle the peers map)");
  ctx->c.comm->connect(rank, world, handles);
  SP_CATCH(ctx)
}
int sp_comm_set_enabled(sp_ctx* ctx, int enabled) { ctx->c.shard_enabled = enabled != 0; return SP_OK; }
int sp_host_pool_selftest(int iterations, int* helpers) {   // runs on the CPU: every job of every run must execute exactly once
  HostPool& p = HostPool::get();
  if (helpers) *helpers = p.helpers();
  int bad = 0;
  for (int it = 0; it < iterations; it++) {
    const int n = 1 + it % 6;
    std::atomic<int> sum{0};
    int hit[8] = {0};
    p.run(n, [&](int i) { hit[i]++; sum.fetch_add(i + 1); });
    if (sum.load() != n * (n + 1) / 2) bad++;
    for (int i = 0; i < n; i++) if (hit[i] != 1) bad++;
    if (it % 20000 == 19999) std::this_thread::sleep_for(std::chrono::milliseconds(1));   // lets the helpers fall asleep once in a while
  }
  return bad;
}
int sp_ctx_set_overlap(sp_ctx* ctx, int enabled) { ctx->c.overlap = enabled != 0; return SP_OK; }
int sp_comm_info(const sp_ctx* ctx, int* rank, int* world) { *rank = ctx->c.rank(); *world = ctx->c.world(); return SP_OK; }

void sp_io_bytes(unsigned long long* h, unsigned long long* d) { dev::io_bytes(h, d); }
void sp_prof_enable(int on) { dev::prof_enable(on != 0); }
int sp_prof_report(char* buf, size_t buflen) {
  std::string s = dev::prof_report();
  if (s.size() + 1 > buflen) return SP_ERR_INVALID_ARG;
  memcpy(buf, s.c_str(), s.size() + 1);
  return SP_OK;
}
int sp_timer_start(sp_ctx* ctx) {
  SP_TRY(ctx)
  if (!ctx->c.ev_a) { ctx->c.ev_a = dev::event_create(); ctx->c.ev_b = dev::event_create(); }   // events belong to the context (device)
  dev::event_record(ctx->c.ev_a, ctx->c.stream);
  ctx->c.timer_running = true;
  SP_CATCH(ctx)
}
int sp_timer_stop_ms(sp_ctx* ctx, float* ms) {
  SP_TRY(ctx)
  if (!ctx->c.timer_running) throw SpError(SP_ERR_INVALID_ARG, "sp_timer_stop_ms without sp_timer_start");
  dev::event_record(ctx->c.ev_b, ctx->c.stream);
  *ms = dev::event_elapsed_ms(ctx->c.ev_a, ctx->c.ev_b);
  ctx->c.timer_running = false;
  SP_CATCH(ctx)
}

// ---- scalars
int sp_scalar_from_bytes(const uint8_t b[32], uint64_t out[4]) {
  Fq f(fq_to_mont(bytes_to_u256(b)));
  fq_out(out, f);
  return fq_bytes_canonical(b) ? SP_OK : SP_ERR_INVALID_SCALAR;
}
void sp_scalar_to_bytes(const uint64_t m[4], uint8_t out[32]) { fq_in(m).to_bytes(out); }
void sp_scalar_from_bytes_wide(const uint8_t w[64], uint64_t out[4]) { fq_out(out, Fq::from_bytes_wide(w)); }
void sp_scalar_mul(const uint64_t a[4], const uint64_t b[4], uint64_t out[4]) { fq_out(out, fq_in(a) * fq_in(b)); }
void sp_scalar_add(const uint64_t a[4], const uint64_t b[4], uint64_t out[4]) { fq_out(out, fq_in(a) + fq_in(b)); }
void sp_scalar_sub(const uint64_t a[4], const uint64_t b[4], uint64_t out[4]) { fq_out(out, fq_in(a) - fq_in(b)); }
int sp_scalar_invert(const uint64_t a[4], uint64_t out[4]) {
  Fq x = fq_in(a);
  fq_out(out, x.inv());
  return x.is_zero() ? SP_ERR_INVALID_SCALAR : SP_OK;
}

// ---- polys
int sp_poly_upload(sp_ctx* ctx, const uint64_t* limbs, size_t len, sp_poly** out) {
  SP_TRY(ctx)
  std::unique_ptr<sp_poly> p(new sp_poly{&ctx->c, DevBuf<u256>(len), len});
  dev::h2d(p->d.p, limbs, len * 32, ctx->c.stream);
  ctx->c.sync();
  *out = p.release();
  SP_CATCH(ctx)
}
int sp_poly_download(sp_ctx* ctx, const sp_poly* p, uint64_t* limbs, size_t len) {
  SP_TRY(ctx)
  if (len > p->d.n) throw SpError(SP_ERR_INVALID_ARG, "download length exceeds allocation");
  dev::d2h(limbs, p->d.p, len * 32, ctx->c.stream);
  ctx->c.sync();
  SP_CATCH(ctx)
}
size_t sp_poly_len(const sp_poly* p) { return p->len; }
void sp_poly_free(sp_poly* p) { delete p; }

static void check_same_len(sp_poly* const* polys, int k) {
  for (int i = 1; i < k; i++) if (polys[i]->len != polys[0]->len) throw SpError(SP_ERR_INVALID_ARG, "polynomials differ in length");
  if (polys[0]->len < 2 || (polys[0]->len & (polys[0]->len - 1))) throw SpError(SP_ERR_INVALID_ARG, "length must be a power of two >= 2");
}
int sp_fold_top(sp_ctx* ctx, sp_poly* const* polys, int k, const uint64_t r[4]) {
  SP_TRY(ctx)
  if (k < 1) throw SpError(SP_ERR_INVALID_ARG, "fold_top: no polynomials");
  check_same_len(polys, k);
  for (int i = 0; i < k; i++) for (int j = 0; j < i; j++) if (polys[i] == polys[j]) throw SpError(SP_ERR_INVALID_ARG, "fold_top: the same polynomial passed twice");
  Fq rr = fq_in(r);
  std::vector<u256*> t(k);
  for (int i = 0; i < k; i++) t[i] = polys[i]->d.p;
  dev::fold_top(t.data(), k, polys[0]->len, rr.m, ctx->c.stream);
  ctx->c.sync();
  for (int i = 0; i < k; i++) polys[i]->len /= 2;
  SP_CATCH(ctx)
}
static int kind_tables(int kind) { return kind == 0 ? 2 : kind == 1 ? 3 : 4; }
static dev::ScInst make_inst(sp_poly* const* polys, int nt) {
  dev::ScInst in;
  for (int t = 0; t < 4; t++) in.t[t] = t < nt ? polys[t]->d.p : nullptr;
  in.c_out = in.t[2];
  in.write_c = 1;
  return in;
}
int sp_sumcheck_eval(sp_ctx* ctx, int kind, sp_poly* const* polys, uint64_t out[3][4]) {
  SP_TRY(ctx)
  if (kind < 0 || kind > 2) throw SpError(SP_ERR_INVALID_ARG, "bad sumcheck kind");
  int nt = kind_tables(kind);
  check_same_len(polys, nt);
  dev::ScInst in = make_inst(polys, nt);
  dev::sc_eval((dev::ScKind)kind, &in, 1, polys[0]->len, ctx->c.small.p, ctx->c.red.p, ctx->c.stream);
  Fq e[3];
  ctx->c.get_small(0, e, 3);
  memcpy(out, e, 96);
  SP_CATCH(ctx)
}
int sp_sumcheck_fold_eval(sp_ctx* ctx, int kind, sp_poly* const* polys, const uint64_t r[4], uint64_t out[3][4]) {
  SP_TRY(ctx)
  if (kind < 0 || kind > 2) throw SpError(SP_ERR_INVALID_ARG, "bad sumcheck kind");
  int nt = kind_tables(kind);
  check_same_len(polys, nt);
  if (polys[0]->len < 4) throw SpError(SP_ERR_INVALID_ARG, "fold_eval needs length >= 4");
  Fq rr = fq_in(r);
  dev::ScInst in = make_inst(polys, nt);
  dev::sc_fold_eval((dev::ScKind)kind, &in, 1, polys[0]->len, rr.m, ctx->c.small.p, ctx->c.red.p, ctx->c.stream);
  Fq e[3];
  ctx->c.get_small(0, e, 3);
  memcpy(out, e, 96);
  for (int i = 0; i < nt; i++) polys[i]->len /= 2;
  SP_CATCH(ctx)
}
// The same two calls on a connected context (sp_comm_connect), every rank passing its CYCLIC SHARDS of the tables (rank r holds the entries
// r, r+W, r+2W, ...): the kernels exchange their partial sums over NVLink and every rank receives the evaluations of the whole tables.
// Binding the top variable stays local as long as the local length is >= 2; shards must be of streaming size (>= 8192 entries) for fold_eval.
static dev::HostSig xr_sig(Ctx& c) { dev::HostSig s; s.done = c.sig_done.p; return s; }
int sp_sumcheck_eval_sharded(sp_ctx* ctx, int kind, sp_poly* const* polys, uint64_t out[3][4]) {
  SP_TRY(ctx)
  if (ctx->c.world() < 2) throw SpError(SP_ERR_INVALID_ARG, "not a connected multi-GPU context");
  if (kind < 0 || kind > 2) throw SpError(SP_ERR_INVALID_ARG, "bad sumcheck kind");
  int nt = kind_tables(kind);
  check_same_len(polys, nt);
  dev::ScInst in = make_inst(polys, nt);
  dev::sc_eval((dev::ScKind)kind, &in, 1, polys[0]->len, ctx->c.small.p, ctx->c.red.p, ctx->c.stream, xr_sig(ctx->c), ctx->c.comm->next_xr());
  Fq e[3];
  ctx->c.get_small(0, e, 3);
  memcpy(out, e, 96);
  SP_CATCH(ctx)
}
int sp_sumcheck_fold_eval_sharded(sp_ctx* ctx, int kind, sp_poly* const* polys, const uint64_t r[4], uint64_t out[3][4]) {
  SP_TRY(ctx)
  if (ctx->c.world() < 2) throw SpError(SP_ERR_INVALID_ARG, "not a connected multi-GPU context");
  if (kind < 0 || kind > 2) throw SpError(SP_ERR_INVALID_ARG, "bad sumcheck kind");
  int nt = kind_tables(kind);
  check_same_len(polys, nt);
  if (polys[0]->len < Ctx::SHARD_MIN_LOCAL) throw SpError(SP_ERR_INVALID_ARG, "sharded fold_eval needs shards of at least 8192 entries (gather the shards and finish on one GPU)");
  Fq rr = fq_in(r);
  dev::ScInst in = make_inst(polys, nt);
  dev::sc_fold_eval((dev::ScKind)kind, &in, 1, polys[0]->len, rr.m, ctx->c.small.p, ctx->c.red.p, ctx->c.stream, xr_sig(ctx->c), ctx->c.comm->next_xr());
  Fq e[3];
  ctx->c.get_small(0, e, 3);
  memcpy(out, e, 96);
  for (int i = 0; i < nt; i++) polys[i]->len /= 2;
  SP_CATCH(ctx)
}
// prove_cubic_batched's evaluation loops (sumcheck.rs:290-357): ninst instances of A*B*C; an sp_poly may serve as C of several instances
// (the shared eq table poly_C_par) — it is then bound once, through a temporary, and copied back
static void batched_insts(sp_ctx* ctx, int ninst, sp_poly* const* A, sp_poly* const* B, sp_poly* const* Cc, std::vector<dev::ScInst>& insts,
                          std::vector<std::pair<sp_poly*, std::unique_ptr<DevBuf<u256>>>>& shared, bool fold) {
  if (ninst < 1 || ninst > 24) throw SpError(SP_ERR_INVALID_ARG, "batched sumcheck: 1..24 instances");
  const size_t len = A[0]->len;
  if (len < 2 || (len & (len - 1))) throw SpError(SP_ERR_INVALID_ARG, "length must be a power of two >= 2");
  for (int i = 0; i < ninst; i++) {
    if (A[i]->len != len || B[i]->len != len || Cc[i]->len != len) throw SpError(SP_ERR_INVALID_ARG, "polynomials differ in length");
    for (int j = 0; j < ninst; j++)
      if (A[i] == B[j] || A[i] == Cc[j] || B[i] == Cc[j] || (i != j && (A[i] == A[j] || B[i] == B[j])))
        throw SpError(SP_ERR_INVALID_ARG, "batched sumcheck: only the C table may be shared between instances");
  }
  insts.resize(ninst);
  for (int i = 0; i < ninst; i++) {
    dev::ScInst& in = insts[i];
    in.t[0] = A[i]->d.p; in.t[1] = B[i]->d.p; in.t[2] = Cc[i]->d.p; in.t[3] = nullptr;
    in.c_out = in.t[2]; in.write_c = 1;
    if (!fold) continue;
    int users = 0, first = -1;
    for (int j = 0; j < ninst; j++) if (Cc[j] == Cc[i]) { users++; if (first < 0) first = j; }
    if (users > 1) {
      DevBuf<u256>* tmp = nullptr;
      for (auto& sh : shared) if (sh.first == Cc[i]) tmp = sh.second.get();
      if (!tmp) { shared.push_back({Cc[i], std::unique_ptr<DevBuf<u256>>(new DevBuf<u256>(len / 2))}); tmp = shared.back().second.get(); }
      in.c_out = tmp->p;
      in.write_c = first == i;
    }
  }
  (void)ctx;
}
int sp_sumcheck_batched_eval(sp_ctx* ctx, int ninst, sp_poly* const* A, sp_poly* const* B, sp_poly* const* Cc, uint64_t* out) {
  SP_TRY(ctx)
  std::vector<dev::ScInst> insts;
  std::vector<std::pair<sp_poly*, std::unique_ptr<DevBuf<u256>>>> shared;
  batched_insts(ctx, ninst, A, B, Cc, insts, shared, false);
  DevBuf<u256> d_out(3 * ninst);
  dev::sc_eval(dev::SC_CUBIC3, insts.data(), ninst, A[0]->len, d_out.p, ctx->c.red.p, ctx->c.stream);
  dev::d2h(out, d_out.p, 96 * (size_t)ninst, ctx->c.stream);
  ctx->c.sync();
  SP_CATCH(ctx)
}
int sp_sumcheck_batched_fold_eval(sp_ctx* ctx, int ninst, sp_poly* const* A, sp_poly* const* B, sp_poly* const* Cc, const uint64_t r[4], uint64_t* out) {
  SP_TRY(ctx)
  if (ninst < 1 || ninst > 24) throw SpError(SP_ERR_INVALID_ARG, "batched sumcheck: 1..24 instances");
  if (A[0]->len < 4) throw SpError(SP_ERR_INVALID_ARG, "fold_eval needs length >= 4");
  std::vector<dev::ScInst> insts;
  std::vector<std::pair<sp_poly*, std::unique_ptr<DevBuf<u256>>>> shared;
  batched_insts(ctx, ninst, A, B, Cc, insts, shared, true);
  Fq rr = fq_in(r);
  DevBuf<u256> d_out(3 * ninst);
  const size_t len = A[0]->len;
  dev::sc_fold_eval(dev::SC_CUBIC3, insts.data(), ninst, len, rr.m, d_out.p, ctx->c.red.p, ctx->c.stream);
  for (auto& sh : shared) dev::d2d(sh.first->d.p, sh.second->p, (len / 2) * sizeof(u256), ctx->c.stream);
  dev::d2h(out, d_out.p, 96 * (size_t)ninst, ctx->c.stream);
  ctx->c.sync();
  std::vector<sp_poly*> seen;
  for (int i = 0; i < ninst; i++)
    for (sp_poly* p : {A[i], B[i], Cc[i]})
      if (std::find(seen.begin(), seen.end(), p) == seen.end()) { seen.push_back(p); p->len /= 2; }
  SP_CATCH(ctx)
}
int sp_eq_evals(sp_ctx* ctx, const uint64_t* r, size_t ell, sp_poly** out) {
  SP_TRY(ctx)
  size_t n = (size_t)1 << ell;
  std::unique_ptr<sp_poly> p(new sp_poly{&ctx->c, DevBuf<u256>(n), n});
  DevBuf<u256> d_r(ell + 1), small(2 * ((size_t)1 << ((ell + 1) / 2)) + 8);
  dev::h2d(d_r.p, r, ell * 32, ctx->c.stream);
  dev::eq_evals(p->d.p, d_r.p, (int)ell, small.p, ctx->c.stream);
  ctx->c.sync();
  *out = p.release();
  SP_CATCH(ctx)
}
int sp_poly_evaluate(sp_ctx* ctx, const sp_poly* p, const uint64_t* r, size_t ell, uint64_t out[4]) {
  SP_TRY(ctx)
  if (p->len != (size_t)1 << ell) throw SpError(SP_ERR_INVALID_ARG, "evaluate: |r| != num_vars");
  DevBuf<u256> d_r(ell + 1), small(2 * ((size_t)1 << ((ell + 1) / 2)) + 8), eq(p->len);
  dev::h2d(d_r.p, r, ell * 32, ctx->c.stream);
  dev::eq_evals(eq.p, d_r.p, (int)ell, small.p, ctx->c.stream);
  dev::dot(ctx->c.small.p + 32, p->d.p, eq.p, p->len, ctx->c.red.p, ctx->c.stream);
  Fq v;
  ctx->c.get_small(32, &v, 1);
  fq_out(out, v);
  SP_CATCH(ctx)
}
int sp_poly_bound_rows(sp_ctx* ctx, const sp_poly* p, const uint64_t* Lm, size_t L_size, sp_poly** out) {
  SP_TRY(ctx)
  if (L_size == 0 || p->len % L_size) throw SpError(SP_ERR_INVALID_ARG, "bound: L does not divide the length");
  size_t R_size = p->len / L_size;
  std::unique_ptr<sp_poly> o(new sp_poly{&ctx->c, DevBuf<u256>(R_size), R_size});
  DevBuf<u256> d_L(L_size), tmp(64 * R_size);
  dev::h2d(d_L.p, Lm, L_size * 32, ctx->c.stream);
  dev::bound_rows(o->d.p, p->d.p, d_L.p, L_size, R_size, tmp.p, ctx->c.stream);
  ctx->c.sync();
  *out = o.release();
  SP_CATCH(ctx)
}
int sp_dot(sp_ctx* ctx, const sp_poly* a, const sp_poly* b, uint64_t out[4]) {
  SP_TRY(ctx)
  if (a->len != b->len) throw SpError(SP_ERR_INVALID_ARG, "dot: length mismatch");
  dev::dot(ctx->c.small.p + 32, a->d.p, b->d.p, a->len, ctx->c.red.p, ctx->c.stream);
  Fq v;
  ctx->c.get_small(32, &v, 1);
  fq_out(out, v);
  SP_CATCH(ctx)
}

// ---- gens / commitments
int sp_gens_create(sp_ctx* ctx, const uint8_t* label, size_t label_len, size_t n, sp_gens** out) {
  SP_TRY(ctx)
  std::unique_ptr<sp_gens> g(new sp_gens);
  g->n = n;
  g->set.reset(new GenSet(&ctx->c, std::string((const char*)label, label_len), n + 1, {}));
  *out = g.release();
  SP_CATCH(ctx)
}
int sp_gens_upload(sp_ctx* ctx, const uint8_t* compressed32, size_t n, sp_gens** out) {
  SP_TRY(ctx)
  DevBuf<uint8_t> d_in(32 * (n + 1));
  DevBuf<ge> g(n + 1);
  DevBuf<int> ok(n + 1);
  dev::h2d(d_in.p, compressed32, 32 * (n + 1), ctx->c.stream);
  dev::decompress_batch(g.p, ok.p, d_in.p, n + 1, ctx->c.stream);
  std::vector<int> h_ok(n + 1);
  dev::d2h(h_ok.data(), ok.p, sizeof(int) * (n + 1), ctx->c.stream);
  ctx->c.sync();
  for (size_t i = 0; i <= n; i++)
    if (!h_ok[i]) throw SpError(SP_ERR_INVALID_POINT, "gens_upload: encoding " + std::to_string(i) + " is not a ristretto255 point");
  std::unique_ptr<sp_gens> G(new sp_gens);
  G->n = n;
  G->set.reset(new GenSet(&ctx->c, g.p, n + 1, {}));
  *out = G.release();
  SP_CATCH(ctx)
}
void sp_gens_free(sp_gens* g) { delete g; }
int sp_gens_export(sp_ctx* ctx, const sp_gens* g, uint8_t* out32) {
  SP_TRY(ctx)
  DevBuf<uint8_t> comp(32 * (g->n + 1));
  dev::compress_batch(comp.p, g->set->G.p, g->n + 1, ctx->c.stream);
  dev::d2h(out32, comp.p, 32 * (g->n + 1), ctx->c.stream);
  ctx->c.sync();
  SP_CATCH(ctx)
}
int sp_msm(sp_ctx* ctx, const sp_gens* g, const uint64_t* scalars, size_t n, uint8_t out32[32]) {
  SP_TRY(ctx)
  if (n > g->n) throw SpError(SP_ERR_INVALID_ARG, "msm: more scalars than generators");
  DevBuf<u256> d(n ? n : 1);
  dev::h2d(d.p, scalars, n * 32, ctx->c.stream);
  CommitKey key{g->set.get(), 0, g->n, g->n};
  std::vector<Cp> out;
  commit_rows_and_compress(ctx->c, key, d.p, n, 1, n, nullptr, out);
  memcpy(out32, out[0].b, 32);
  SP_CATCH(ctx)
}
int sp_commit_rows(sp_ctx* ctx, const sp_gens* g, const sp_poly* p, size_t L, size_t R, const uint64_t* blinds, uint8_t* out32) {
  SP_TRY(ctx)
  if (L * R != p->len || R > g->n) throw SpError(SP_ERR_INVALID_ARG, "commit_rows: shape mismatch");
  CommitKey key{g->set.get(), 0, g->n, g->n};
  std::vector<Fq> bl;
  if (blinds) bl = fq_vec(blinds, L);
  std::vector<Cp> out;
  commit_rows_and_compress(ctx->c, key, p->d.p, R, L, R, blinds ? bl.data() : nullptr, out);
  memcpy(out32, out.data(), 32 * L);
  SP_CATCH(ctx)
}
int sp_point_decompress_check(sp_ctx* ctx, const uint8_t* in32, size_t n, int* ok) {
  SP_TRY(ctx)
  DevBuf<uint8_t> d_in(32 * n);
  DevBuf<ge> pts(n);
  DevBuf<int> d_ok(n);
  dev::h2d(d_in.p, in32, 32 * n, ctx->c.stream);
  dev::decompress_batch(pts.p, d_ok.p, d_in.p, n, ctx->c.stream);
  dev::d2h(ok, d_ok.p, sizeof(int) * n, ctx->c.stream);
  ctx->c.sync();
  SP_CATCH(ctx)
}
int sp_point_roundtrip(sp_ctx* ctx, const uint8_t* in32, size_t n, uint8_t* out32) {
  SP_TRY(ctx)
  DevBuf<uint8_t> d_in(32 * n), d_out(32 * n);
  DevBuf<ge> pts(n);
  dev::h2d(d_in.p, in32, 32 * n, ctx->c.stream);
  dev::decompress_batch(pts.p, nullptr, d_in.p, n, ctx->c.stream);
  dev::compress_batch(d_out.p, pts.p, n, ctx->c.stream);
  dev::d2h(out32, d_out.p, 32 * n, ctx->c.stream);
  ctx->c.sync();
  SP_CATCH(ctx)
}

// ---- inner-product argument, operator level: BulletReductionProof::prove (nizk/bullet.rs:32-132) with the host keeping the transcript.
// One handle per reduction: a_vec / b_vec (consumed), the generator set G (with its fixed-base tables) and the running scalars s of the
// unfolded-G formulation (DESIGN.md "IPA without folding G").  Per round: sp_ipa_round_LR (bullet.rs:78-97) -> transcript -> sp_ipa_fold (:105-111).
struct sp_ipa {
  Ctx* ctx; const GenSet* gs; size_t n, cur;
  DevBuf<u256> a, b, svec;
  DevBuf<ge> pts;
};
static hge point_in(const uint8_t p32[32], const char* what) {
  ge g;
  if (!ristretto_decode(g, bytes_to_u256(p32))) throw SpError(SP_ERR_INVALID_POINT, std::string(what) + " is not a ristretto255 point");
  return to_hge(g);
}
int sp_ipa_begin(sp_ctx* ctx, const sp_gens* gens, const sp_poly* a_vec, const sp_poly* b_vec, sp_ipa** out) {
  SP_TRY(ctx)
  const size_t n = a_vec->len;
  if (n < 1 || (n & (n - 1)) || b_vec->len != n || n > gens->n) throw SpError(SP_ERR_INVALID_ARG, "ipa: |a| = |b| = a power of two, at most the generator count");
  std::unique_ptr<sp_ipa> h(new sp_ipa);
  h->ctx = &ctx->c; h->gs = gens->set.get(); h->n = n; h->cur = n;
  h->a.alloc(n); h->b.alloc(n); h->svec.alloc(n); h->pts.alloc(2);
  dev::d2d(h->a.p, a_vec->d.p, n * sizeof(u256), ctx->c.stream);
  dev::d2d(h->b.p, b_vec->d.p, n * sizeof(u256), ctx->c.stream);
  dev::fill_one(h->svec.p, n, ctx->c.stream);
  ctx->c.ensure_scratch(std::max(dev::msm_scratch_bytes(4, n), dev::ipa_msm_scratch_points(n, gens->set->wbits) * sizeof(sp::ge)) + 64);
  ctx->c.sync();
  *out = h.release();
  SP_CATCH(ctx)
}
void sp_ipa_free(sp_ipa* h) { delete h; }
// L = <a_L, G_R> + c_L*Q + blind_L*H,  R = <a_R, G_L> + c_R*Q + blind_R*H  with c_L = <a_L, b_R>, c_R = <a_R, b_L>   (bullet.rs:74-97)
int sp_ipa_round_LR(sp_ctx* ctx, sp_ipa* h, const uint8_t Q32[32], const uint8_t H32[32], const uint64_t blind_L[4], const uint64_t blind_R[4], uint8_t L32[32],
                    uint8_t R32[32]) {
  SP_TRY(ctx)
  Ctx& c = ctx->c;
  if (h->cur < 2) throw SpError(SP_ERR_INVALID_ARG, "ipa: the reduction is complete");
  const size_t half = h->cur / 2;
  const u256* da[2] = {h->a.p, h->a.p + half};
  const u256* db[2] = {h->b.p + half, h->b.p};
  dev::dot_pairs(c.small.p + 16, da, db, 2, half, c.red.p, c.stream);
  dev::ipa_msm(h->pts.p, h->gs->table.p, h->gs->wbits, h->a.p, h->svec.p, h->cur, h->n, c.scratch.p, c.sig_done.p + 1, c.stream);
  Fq cc[2];
  c.get_small(16, cc, 2);
  ge lr[2];
  dev::d2h(lr, h->pts.p, 2 * sizeof(ge), c.stream);
  c.sync();
  const hge Q = point_in(Q32, "Q"), H = point_in(H32, "H");
  const Fq bl[2] = {fq_in(blind_L), fq_in(blind_R)};
  Cp outp[2];
  hge full[2];
  for (int k = 0; k < 2; k++) full[k] = hge_add(to_hge(lr[k]), hge_add(hge_scalarmul(cc[k].canonical(), Q), hge_scalarmul(bl[k].canonical(), H)));
  compress2(full[0], full[1], outp[0], outp[1]);
  memcpy(L32, outp[0].b, 32); memcpy(R32, outp[1].b, 32);
  SP_CATCH(ctx)
}
// a_L <- a_L*u + u^-1*a_R,  b_L <- b_L*u^-1 + u*b_R,  G_L <- u^-1*G_L + u*G_R (kept implicit in s)      (bullet.rs:105-108)
int sp_ipa_fold(sp_ctx* ctx, sp_ipa* h, const uint64_t u[4], const uint64_t u_inv[4]) {
  SP_TRY(ctx)
  if (h->cur < 2) throw SpError(SP_ERR_INVALID_ARG, "ipa: the reduction is complete");
  const Fq uu = fq_in(u), ui = fq_in(u_inv);
  if (!(uu * ui == Fq::one())) throw SpError(SP_ERR_INVALID_SCALAR, "ipa: u_inv is not the inverse of u");
  const size_t half = h->cur / 2;
  dev::ipa_fold_ab(h->a.p, h->b.p, half, uu.m, ui.m, ctx->c.stream);
  dev::ipa_update_s(h->svec.p, half, h->n, uu.m, ui.m, ctx->c.stream);
  h->cur = half;
  ctx->c.sync();
  SP_CATCH(ctx)
}
// a[0], b[0] and G[0] of the fully folded vectors (bullet.rs:113-122: the caller forms Gamma_hat from them)
int sp_ipa_finish(sp_ctx* ctx, sp_ipa* h, uint64_t a_hat[4], uint64_t b_hat[4], uint8_t G_hat32[32]) {
  SP_TRY(ctx)
  Ctx& c = ctx->c;
  if (h->cur != 1) throw SpError(SP_ERR_INVALID_ARG, "ipa: rounds remain");
  dev::msm_rows(h->pts.p, h->gs->table.p, h->gs->wbits, h->svec.p, h->n, 1, h->n, nullptr, 0, c.scratch.p, c.stream);
  DevBuf<uint8_t> comp(32);
  dev::compress_batch(comp.p, h->pts.p, 1, c.stream);
  dev::d2h(a_hat, h->a.p, 32, c.stream);
  dev::d2h(b_hat, h->b.p, 32, c.stream);
  dev::d2h(G_hat32, comp.p, 32, c.stream);
  c.sync();
  SP_CATCH(ctx)
}

// ---- variable-base MSM (kernels_pip.cu)
static void points_from_ge(sp_ctx* ctx, sp_points* P, const ge* d_ge, size_t n) {
  P->ctx = &ctx->c; P->n = n;
  P->pts.alloc(n ? n : 1);
  dev::points_to_niels(P->pts.p, d_ge, n, ctx->c.stream);
}
int sp_points_upload(sp_ctx* ctx, const uint8_t* in32, size_t n, sp_points** out) {
  SP_TRY(ctx)
  DevBuf<uint8_t> d_in(32 * n + 1);
  DevBuf<ge> g(n + 1);
  DevBuf<int> ok(n + 1);
  dev::h2d(d_in.p, in32, 32 * n, ctx->c.stream);
  dev::decompress_batch(g.p, ok.p, d_in.p, n, ctx->c.stream);
  std::vector<int> h_ok(n);
  dev::d2h(h_ok.data(), ok.p, sizeof(int) * n, ctx->c.stream);
  ctx->c.sync();
  for (size_t i = 0; i < n; i++)
    if (!h_ok[i]) throw SpError(SP_ERR_INVALID_POINT, "points_upload: encoding " + std::to_string(i) + " is not a ristretto255 point");
  std::unique_ptr<sp_points> P(new sp_points);
  points_from_ge(ctx, P.get(), g.p, n);
  ctx->c.sync();
  *out = P.release();
  SP_CATCH(ctx)
}
int sp_points_derive(sp_ctx* ctx, const uint8_t* label, size_t label_len, size_t n, sp_points** out) {
  SP_TRY(ctx)
  // the same SHAKE256(label || basepoint) stream as MultiCommitGens::new, squeezed and mapped in slabs so 2^24 points need no 1 GiB staging buffer
  static const uint8_t basepoint[32] = {0xe2, 0xf2, 0xae, 0x0a, 0x6a, 0xbc, 0x4e, 0x71, 0xa8, 0x84, 0xa9, 0x61, 0xc5, 0x00, 0x51, 0x5f,
                                        0x58, 0xe3, 0x0b, 0x6a, 0xa5, 0x82, 0xdd, 0x8d, 0xb6, 0xa6, 0x59, 0x45, 0xe0, 0x8d, 0x2d, 0x76};
  std::vector<uint8_t> seed(label, label + label_len);
  seed.insert(seed.end(), basepoint, basepoint + 32);
  Shake256 xof(seed.data(), seed.size());
  std::unique_ptr<sp_points> P(new sp_points);
  P->ctx = &ctx->c; P->n = n;
  P->pts.alloc(n ? n : 1);
  const size_t slab = (size_t)1 << 18;
  std::vector<uint8_t> uni(64 * std::min(slab, n ? n : 1));
  DevBuf<uint8_t> d_uni(uni.size());
  DevBuf<ge> g(std::min(slab, n ? n : 1));
  for (size_t i0 = 0; i0 < n; i0 += slab) {
    size_t m = std::min(slab, n - i0);
    xof.squeeze(uni.data(), 64 * m);
    dev::h2d(d_uni.p, uni.data(), 64 * m, ctx->c.stream);
    dev::gens_from_uniform(g.p, d_uni.p, m, ctx->c.stream);
    dev::points_to_niels(P->pts.p + i0, g.p, m, ctx->c.stream);
    ctx->c.sync();
  }
  *out = P.release();
  SP_CATCH(ctx)
}
size_t sp_points_len(const sp_points* p) { return p->n; }
void sp_points_free(sp_points* p) { delete p; }
int sp_points_export(sp_ctx* ctx, const sp_points* p, size_t offset, size_t n, uint8_t* out32) {
  SP_TRY(ctx)
  if (offset + n > p->n) throw SpError(SP_ERR_INVALID_ARG, "points_export: range outside the point set");
  // 1 * P_i through the MSM path's own mixed addition, then compress
  DevBuf<ge> g(n + 1);
  DevBuf<uint8_t> comp(32 * n + 1);
  dev::niels_to_ge(g.p, p->pts.p + offset, n, ctx->c.stream);
  dev::compress_batch(comp.p, g.p, n, ctx->c.stream);
  dev::d2h(out32, comp.p, 32 * n, ctx->c.stream);
  ctx->c.sync();
  SP_CATCH(ctx)
}
static void msm_var_run(sp_ctx* ctx, const sp_points* p, size_t offset, const u256* d_scalars, size_t n, uint8_t out32[32]) {
  if (offset + n > p->n) throw SpError(SP_ERR_INVALID_ARG, "msm_var: range outside the point set");
  DevBuf<ge> out(1);
  DevBuf<uint8_t> comp(32);
  if (n == 0) {
    ge id = ge_identity();
    dev::h2d(out.p, &id, sizeof(ge), ctx->c.stream);
  } else {
    const char* cenv = getenv("SP_PIP_WINDOW");
    dev::PipPlan plan = dev::pip_plan(n, cenv ? atoi(cenv) : 0);
    const size_t need = dev::pip_scratch_bytes(plan);
    if (need > p->scratch_bytes) { p->scratch.alloc(need); p->scratch_bytes = need; }   // the sort workspace stays with the point set
    dev::msm_var(out.p, p->pts.p + offset, d_scalars, plan, p->scratch.p, ctx->c.stream);
  }
  dev::compress_batch(comp.p, out.p, 1, ctx->c.stream);
  dev::d2h(out32, comp.p, 32, ctx->c.stream);
  ctx->c.sync();
}
// MSM split over the ranks of a connected context: every rank passes ITS slice of the points (offset, scalars->len) and of the scalars; the
// partial sums meet in a point-add all-reduce (all-gather of the W extended points into every window, W-term sum) and every rank returns the
// same encoding of sum over all ranks.
int sp_msm_var_sharded(sp_ctx* ctx, const sp_points* p, size_t offset, const sp_poly* scalars, uint8_t out32[32]) {
  SP_TRY(ctx)
  Ctx& c = ctx->c;
  if (c.world() < 2) throw SpError(SP_ERR_INVALID_ARG, "not a connected multi-GPU context");
  const size_t n = scalars->len;
  if (offset + n > p->n) throw SpError(SP_ERR_INVALID_ARG, "msm_var: range outside the point set");
  DevBuf<ge> part(1), total(1);
  DevBuf<uint8_t> comp(32);
  if (n == 0) { ge id = ge_identity(); dev::h2d(part.p, &id, sizeof(ge), c.stream); }
  else {
    const char* cenv = getenv("SP_PIP_WINDOW");
    dev::PipPlan plan = dev::pip_plan(n, cenv ? atoi(cenv) : 0);
    const size_t need = dev::pip_scratch_bytes(plan);
    if (need > p->scratch_bytes) { p->scratch.alloc(need); p->scratch_bytes = need; }
    dev::msm_var(part.p, p->pts.p + offset, scalars->d.p, plan, p->scratch.p, c.stream);
  }
  const ge* all = reinterpret_cast<const ge*>(c.allgather_block(part.p, sizeof(ge)));
  dev::sum_points(total.p, all, c.world(), c.stream);
  dev::compress_batch(comp.p, total.p, 1, c.stream);
  dev::d2h(out32, comp.p, 32, c.stream);
  c.sync();
  SP_CATCH(ctx)
}
int sp_msm_var(sp_ctx* ctx, const sp_points* p, size_t offset, const uint64_t* scalars, size_t n, uint8_t out32[32]) {
  SP_TRY(ctx)
  DevBuf<u256> d(n ? n : 1);
  dev::h2d(d.p, scalars, n * 32, ctx->c.stream);
  msm_var_run(ctx, p, offset, d.p, n, out32);
  SP_CATCH(ctx)
}
int sp_msm_var_resident(sp_ctx* ctx, const sp_points* p, size_t offset, const sp_poly* scalars, uint8_t out32[32]) {
  SP_TRY(ctx)
  msm_var_run(ctx, p, offset, scalars->d.p, scalars->len, out32);
  SP_CATCH(ctx)
}

// ---- instances
static size_t next_pow2(size_t x) { size_t p = 1; while (p < x) p <<= 1; return p; }

int sp_instance_create(sp_ctx* ctx, size_t num_cons, size_t num_vars, size_t num_inputs, const uint64_t* A_row, const uint64_t* A_col, const uint8_t* A_val,
                       size_t nA, const uint64_t* B_row, const uint64_t* B_col, const uint8_t* B_val, size_t nB, const uint64_t* C_row, const uint64_t* C_col,
                       const uint8_t* C_val, size_t nC, sp_instance** out) {
  SP_TRY(ctx)
  // Instance::new padding rules (lib.rs:129-198)
  size_t num_vars_padded = next_pow2(std::max(num_vars, num_inputs + 1));
  size_t num_cons_padded = num_cons;
  if (num_cons_padded == 0 || num_cons_padded == 1) num_cons_padded = 2;
  if (next_pow2(num_cons) != num_cons) num_cons_padded = next_pow2(num_cons);
  std::unique_ptr<sp_instance> I(new sp_instance);
  I->inst.num_cons = num_cons_padded; I->inst.num_vars = num_vars_padded; I->inst.num_inputs = num_inputs;
  const uint64_t* rows[3] = {A_row, B_row, C_row};
  const uint64_t* cols[3] = {A_col, B_col, C_col};
  const uint8_t* vals[3] = {A_val, B_val, C_val};
  size_t nn[3] = {nA, nB, nC};
  for (int m = 0; m < 3; m++) {
    SparseMatDev& M = I->inst.M[m];
    for (size_t k = 0; k < nn[m]; k++) {
      size_t row = rows[m][k], col = cols[m][k];
      if (row >= num_cons) throw SpError(SP_ERR_INVALID_INDEX, "R1CSError::InvalidIndex (row)");
      if (col >= num_vars + 1 + num_inputs) throw SpError(SP_ERR_INVALID_INDEX, "R1CSError::InvalidIndex (col)");
      const uint8_t* vb = vals[m] + 32 * k;
      if (!fq_bytes_canonical(vb)) throw SpError(SP_ERR_INVALID_SCALAR, "R1CSError::InvalidScalar");
      M.row.push_back((uint32_t)row);
      M.col.push_back((uint32_t)(col >= num_vars ? col + num_vars_padded - num_vars : col));
      M.val.push_back(Fq(fq_to_mont(bytes_to_u256(vb))));
    }
    if (num_cons == 0 || num_cons == 1)
      for (size_t i = nn[m]; i < num_cons_padded; i++) { M.row.push_back((uint32_t)i); M.col.push_back((uint32_t)num_vars); M.val.push_back(Fq::zero()); }
  }
  I->inst.finalize(&ctx->c);
  *out = I.release();
  SP_CATCH(ctx)
}

static std::vector<Fq> prg_scalars(const std::string& tag, size_t n, uint64_t seed) {
  // DESIGN.md "deterministic inputs": SHAKE256("spartan-b200/v1/" || tag || LE64(seed)), 64 bytes per scalar -> from_bytes_wide
  std::string s = "spartan-b200/v1/" + tag;
  std::vector<uint8_t> in(s.begin(), s.end());
  for (int i = 0; i < 8; i++) in.push_back((uint8_t)(seed >> (8 * i)));
  std::vector<uint8_t> raw(64 * n);
  shake256(raw.data(), raw.size(), in.data(), in.size());
  std::vector<Fq> out(n);
  for (size_t i = 0; i < n; i++) out[i] = Fq::from_bytes_wide(raw.data() + 64 * i);
  return out;
}

int sp_instance_synthetic(sp_ctx* ctx, size_t num_cons, size_t num_vars, size_t num_inputs, uint64_t seed, sp_instance** out, uint64_t* vars_out,
                          uint64_t* inputs_out) {
  SP_TRY(ctx)
  // R1CSShape::produce_synthetic_r1cs (r1cs.rs:160-238)
  if (next_pow2(num_cons) != num_cons || next_pow2(num_vars) != num_vars || !(num_inputs < num_vars)) throw SpError(SP_ERR_INVALID_ARG, "synthetic: sizes");
  size_t size_z = num_vars + num_inputs + 1;
  std::vector<Fq> Z = prg_scalars("Z", size_z, seed);
  Z[num_vars] = Fq::one();
  std::unique_ptr<sp_instance> I(new sp_instance);
  I->inst.num_cons = num_cons; I->inst.num_vars = num_vars; I->inst.num_inputs = num_inputs;
  SparseMatDev &A = I->inst.M[0], &B = I->inst.M[1], &Cm = I->inst.M[2];
  // batch inversion of the Z values used as C_val denominators
  std::vector<Fq> zinv(size_z), pref(size_z);
  Fq acc = Fq::one();
  for (size_t i = 0; i < size_z; i++) { pref[i] = acc; if (!Z[i].is_zero()) acc *= Z[i]; }
  Fq ainv = acc.inv();
  for (size_t i = size_z; i-- > 0;) { if (Z[i].is_zero()) { zinv[i] = Fq::zero(); continue; } zinv[i] = ainv * pref[i]; ainv *= Z[i]; }
  for (size_t i = 0; i < num_cons; i++) {
    size_t a = i % size_z, b = (i + 2) % size_z, c = (i + 3) % size_z;
    A.row.push_back((uint32_t)i); A.col.push_back((uint32_t)a); A.val.push_back(Fq::one());
    B.row.push_back((uint32_t)i); B.col.push_back((uint32_t)b); B.val.push_back(Fq::one());
    Fq ab = Z[a] * Z[b];
    Cm.row.push_back((uint32_t)i);
    if (Z[c].is_zero()) { Cm.col.push_back((uint32_t)num_vars); Cm.val.push_back(ab); }
    else { Cm.col.push_back((uint32_t)c); Cm.val.push_back(ab * zinv[c]); }
  }
  // the COO column index space is [vars | 1 | inputs] padded to 2*num_vars: columns >= num_vars need no shift here because
  // produce_synthetic_r1cs requires num_vars to be a power of two already (r1cs.rs:172-176)
  I->inst.finalize(&ctx->c);
  memcpy(vars_out, Z.data(), 32 * num_vars);
  memcpy(inputs_out, Z.data() + num_vars + 1, 32 * num_inputs);
  *out = I.release();
  SP_CATCH(ctx)
}
void sp_instance_free(sp_instance* inst) { delete inst; }
int sp_instance_dims(const sp_instance* inst, size_t* nc, size_t* nv, size_t* ni) {
  *nc = inst->inst.num_cons; *nv = inst->inst.num_vars; *ni = inst->inst.num_inputs;
  return SP_OK;
}
int sp_instance_set_digest(sp_instance* inst, const uint8_t* digest, size_t len) { inst->inst.digest.assign(digest, digest + len); return SP_OK; }
int sp_instance_bincode(const sp_instance* inst, uint8_t** out, size_t* len) {
  std::vector<uint8_t> raw = inst->inst.shape_bincode();
  *out = dup_bytes(raw); *len = raw.size();
  return SP_OK;
}
// R1CSShape::get_digest (r1cs.rs:154-158): the bytes NIZK::prove / verify absorb (lib.rs:514) — the caller's (sp_instance_set_digest) or, by
// default, this library's miniz-level-6 zlib stream of bincode(shape), computed on first use
int sp_instance_digest(const sp_instance* inst, uint8_t** out, size_t* len) {
  try {
    const std::vector<uint8_t>& d = inst->inst.shape_digest();
    *out = dup_bytes(d); *len = d.size();
    return SP_OK;
  } catch (...) { return SP_ERR_INTERNAL; }
}
int sp_instance_nnz(const sp_instance* inst, int m, size_t* nnz) { if (m < 0 || m > 2) return SP_ERR_INVALID_ARG; *nnz = inst->inst.M[m].row.size(); return SP_OK; }
int sp_instance_export(const sp_instance* inst, int m, uint64_t* row, uint64_t* col, uint64_t* val) {
  if (m < 0 || m > 2) return SP_ERR_INVALID_ARG;
  const SparseMatDev& M = inst->inst.M[m];
  for (size_t k = 0; k < M.row.size(); k++) { row[k] = M.row[k]; col[k] = M.col[k]; }
  memcpy(val, M.val.data(), 32 * M.val.size());
  return SP_OK;
}

static DevBuf<u256> upload_padded_vars(Ctx& c, const Instance& I, const uint64_t* vars, size_t nvars) {
  if (nvars > I.num_vars) throw SpError(SP_ERR_INVALID_INPUTS, "R1CSError::InvalidNumberOfInputs (vars)");
  DevBuf<u256> d(I.num_vars);
  dev::h2d(d.p, vars, nvars * 32, c.stream);
  if (nvars < I.num_vars) dev::dzero(d.p + nvars, (I.num_vars - nvars) * 32, c.stream);  // Assignment::pad (lib.rs:91-104)
  return d;
}

int sp_instance_is_sat(sp_ctx* ctx, const sp_instance* inst, const uint64_t* vars, size_t nvars, const uint64_t* inputs, size_t ninputs, int* sat) {
  SP_TRY(ctx)
  const Instance& I = inst->inst;
  if (ninputs != I.num_inputs) throw SpError(SP_ERR_INVALID_INPUTS, "R1CSError::InvalidNumberOfInputs");
  Ctx& c = ctx->c;
  DevBuf<u256> d_vars = upload_padded_vars(c, I, vars, nvars);
  DevBuf<u256> z(2 * I.num_vars), Az(I.num_cons), Bz(I.num_cons), Cz(I.num_cons);
  dev::d2d(z.p, d_vars.p, I.num_vars * 32, c.stream);
  std::vector<Fq> tail(1 + ninputs);
  tail[0] = Fq::one();
  if (ninputs) memcpy(&tail[1], inputs, 32 * ninputs);
  dev::h2d(z.p + I.num_vars, tail.data(), tail.size() * 32, c.stream);
  dev::dzero(z.p + I.num_vars + tail.size(), (I.num_vars - tail.size()) * 32, c.stream);
  u256* outs[3] = {Az.p, Bz.p, Cz.p};
  for (int m = 0; m < 3; m++) dev::spmv(outs[m], I.num_cons, I.M[m].csr_ptr.p, I.M[m].csr_idx.p, I.M[m].csr_val.p, z.p, c.stream);
  DevBuf<u256> prod(I.num_cons);
  dev::hadamard(prod.p, Az.p, Bz.p, I.num_cons, c.stream);
  std::vector<Fq> p = c.download(prod.p, I.num_cons), cz = c.download(Cz.p, I.num_cons);
  *sat = 1;
  for (size_t i = 0; i < I.num_cons; i++) if (!(p[i] == cz[i])) { *sat = 0; break; }   // r1cs.rs:265
  SP_CATCH(ctx)
}

// ---- NIZK
int sp_nizk_gens_create(sp_ctx* ctx, size_t num_cons, size_t num_vars, size_t num_inputs, sp_nizk_gens** out) {
  SP_TRY(ctx)
  (void)num_cons;
  size_t num_vars_padded = next_pow2(std::max(num_vars, num_inputs + 1));  // lib.rs:475-481
  std::unique_ptr<sp_nizk_gens> g(new sp_nizk_gens);
  g->g.reset(new R1CSGens(&ctx->c, "gens_r1cs_sat", num_vars_padded));
  *out = g.release();
  SP_CATCH(ctx)
}
void sp_nizk_gens_free(sp_nizk_gens* g) { delete g; }

// the caller's transcript: either a fresh Transcript::new(label) (state == NULL) or the caller-owned STROBE state, updated in place
struct TranscriptArg {
  const uint8_t* label; size_t label_len; uint8_t* state;
  Transcript open() const { return state ? Transcript(Transcript::FromState(), state) : Transcript(std::string((const char*)label, label_len)); }
  void close(const Transcript& T) const { if (state) T.export_state(state); }
};
static int nizk_prove_common(sp_ctx* ctx, const sp_instance* inst, const u256* d_vars, const uint64_t* inputs, size_t ninputs, const sp_nizk_gens* gens,
                             const TranscriptArg& ta, const uint64_t seed[4], uint8_t** proof, size_t* proof_len) {
  SP_TRY(ctx)
  if (ninputs != inst->inst.num_inputs) throw SpError(SP_ERR_INVALID_INPUTS, "R1CSError::InvalidNumberOfInputs");
  if (!seed) throw SpError(SP_ERR_INVALID_ARG, "tape seed is NULL: draw it from the OS RNG (random.rs:13-15); a fixed seed makes every blind public");
  require_reduced(inputs, ninputs, "inputs"); require_reduced(seed, 1, "tape seed");
  Transcript T = ta.open();
  NizkProof P;
  nizk_prove(ctx->c, inst->inst, d_vars, fq_vec(inputs, ninputs), *gens->g, T, fq_in(seed), P);
  ta.close(T);
  Writer w;
  P.ser(w);
  *proof = dup_bytes(w.out); *proof_len = w.out.size();
  SP_CATCH(ctx)
}
int sp_nizk_prove(sp_ctx* ctx, const sp_instance* inst, const uint64_t* vars, size_t nvars, const uint64_t* inputs, size_t ninputs, const sp_nizk_gens* gens,
                  const uint8_t* label, size_t label_len, const uint64_t seed[4], uint8_t** proof, size_t* proof_len) {
  DevBuf<u256> d_vars;
  try { d_vars = upload_padded_vars(ctx->c, inst->inst, vars, nvars); }
  catch (const SpError& e) { ctx->c.last_error = e.what(); return e.code; }
  catch (const std::exception& e) { ctx->c.last_error = e.what(); return SP_ERR_CUDA; }
  return nizk_prove_common(ctx, inst, d_vars.p, inputs, ninputs, gens, TranscriptArg{label, label_len, nullptr}, seed, proof, proof_len);
}
// NIZK::prove(&inst, vars, &inputs, &gens, transcript: &mut Transcript) with the caller's transcript state in / out (lib.rs:501-508)
int sp_nizk_prove_t(sp_ctx* ctx, const sp_instance* inst, const uint64_t* vars, size_t nvars, const uint64_t* inputs, size_t ninputs, const sp_nizk_gens* gens,
                    uint8_t* strobe_state, const uint64_t seed[4], uint8_t** proof, size_t* proof_len) {
  if (!strobe_state) { ctx->c.last_error = "transcript state is NULL"; return SP_ERR_INVALID_ARG; }
  DevBuf<u256> d_vars;
  try { d_vars = upload_padded_vars(ctx->c, inst->inst, vars, nvars); }
  catch (const SpError& e) { ctx->c.last_error = e.what(); return e.code; }
  catch (const std::exception& e) { ctx->c.last_error = e.what(); return SP_ERR_CUDA; }
  return nizk_prove_common(ctx, inst, d_vars.p, inputs, ninputs, gens, TranscriptArg{nullptr, 0, strobe_state}, seed, proof, proof_len);
}
int sp_nizk_prove_resident(sp_ctx* ctx, const sp_instance* inst, const sp_poly* vars, const uint64_t* inputs, size_t ninputs, const sp_nizk_gens* gens,
                           const uint8_t* label, size_t label_len, const uint64_t seed[4], uint8_t** proof, size_t* proof_len) {
  if (vars->len != inst->inst.num_vars) { ctx->c.last_error = "resident vars must already be padded to num_vars"; return SP_ERR_INVALID_INPUTS; }
  return nizk_prove_common(ctx, inst, vars->d.p, inputs, ninputs, gens, TranscriptArg{label, label_len, nullptr}, seed, proof, proof_len);
}

// ---- SNARK
int sp_snark_gens_create(sp_ctx* ctx, size_t num_cons, size_t num_vars, size_t num_inputs, size_t num_nz_entries, sp_snark_gens** out) {
  SP_TRY(ctx)
  std::unique_ptr<sp_snark_gens> g(new sp_snark_gens);
  g->g.reset(new SnarkGens(&ctx->c, num_cons, num_vars, num_inputs, num_nz_entries));
  *out = g.release();
  SP_CATCH(ctx)
}
void sp_snark_gens_free(sp_snark_gens* g) { delete g; }
int sp_snark_encode(sp_ctx* ctx, const sp_instance* inst, const sp_snark_gens* gens, sp_snark_encoding** out) {
  SP_TRY(ctx)
  std::unique_ptr<sp_snark_encoding> e(new sp_snark_encoding);
  e->e.reset(new SnarkEncoding());
  snark_encode(ctx->c, inst->inst, *gens->g, *e->e);
  *out = e.release();
  SP_CATCH(ctx)
}
void sp_snark_encoding_free(sp_snark_encoding* e) { delete e; }
int sp_snark_commitment_bytes(const sp_snark_encoding* e, uint8_t** out, size_t* len) {
  Writer w;
  e->e->ser_commitment(w);
  *out = dup_bytes(w.out); *len = w.out.size();
  return SP_OK;
}
// a verifier holds only the ComputationCommitment: bincode(ComputationCommitment) -> a handle usable with sp_snark_verify (not with sp_snark_prove)
int sp_snark_commitment_load(sp_ctx* ctx, const uint8_t* bytes, size_t len, sp_snark_encoding** out) {
  SP_TRY(ctx)
  size_t pos = 0;
  auto u64 = [&]() { if (len - pos < 8) throw SpError(SP_ERR_INVALID_ARG, "commitment truncated"); uint64_t x = 0; for (int i = 0; i < 8; i++) x |= (uint64_t)bytes[pos + i] << (8 * i); pos += 8; return x; };
  auto pts = [&](PolyCommitment& c) {
    uint64_t k = u64();
    if (k > (len - pos) / 32) throw SpError(SP_ERR_INVALID_ARG, "commitment truncated");
    c.C.resize(k);
    for (auto& p : c.C) { memcpy(p.b, bytes + pos, 32); pos += 32; }
  };
  std::unique_ptr<sp_snark_encoding> E(new sp_snark_encoding);
  E->e.reset(new SnarkEncoding);
  SnarkEncoding& e = *E->e;
  e.num_cons = u64(); e.num_vars = u64(); e.num_inputs = u64(); e.batch_size = u64(); e.num_ops = u64(); e.num_mem_cells = u64();
  pts(e.comm_comb_ops); pts(e.comm_comb_mem);
  if (pos != len) throw SpError(SP_ERR_INVALID_ARG, "trailing bytes after the commitment");
  *out = E.release();
  SP_CATCH(ctx)
}
static int snark_prove_common(sp_ctx* ctx, const sp_instance* inst, const sp_snark_encoding* enc, const u256* d_vars, const uint64_t* inputs, size_t ninputs,
                              const sp_snark_gens* gens, const TranscriptArg& ta, const uint64_t seed[4], uint8_t** proof, size_t* proof_len) {
  SP_TRY(ctx)
  if (ninputs != inst->inst.num_inputs) throw SpError(SP_ERR_INVALID_INPUTS, "R1CSError::InvalidNumberOfInputs");
  if (!enc->e->comb_ops.p) throw SpError(SP_ERR_INVALID_ARG, "this handle holds a commitment only (sp_snark_commitment_load): proving needs sp_snark_encode");
  if (!seed) throw SpError(SP_ERR_INVALID_ARG, "tape seed is NULL: draw it from the OS RNG (random.rs:13-15); a fixed seed makes every blind public");
  require_reduced(inputs, ninputs, "inputs"); require_reduced(seed, 1, "tape seed");
  Transcript T = ta.open();
  Writer w;
  snark_prove(ctx->c, inst->inst, *enc->e, d_vars, fq_vec(inputs, ninputs), *gens->g, T, fq_in(seed), w);
  ta.close(T);
  *proof = dup_bytes(w.out); *proof_len = w.out.size();
  SP_CATCH(ctx)
}
int sp_snark_prove(sp_ctx* ctx, const sp_instance* inst, const sp_snark_encoding* enc, const uint64_t* vars, size_t nvars, const uint64_t* inputs, size_t ninputs,
                   const sp_snark_gens* gens, const uint8_t* label, size_t label_len, const uint64_t seed[4], uint8_t** proof, size_t* proof_len) {
  DevBuf<u256> d_vars;
  try { d_vars = upload_padded_vars(ctx->c, inst->inst, vars, nvars); }
  catch (const SpError& e) { ctx->c.last_error = e.what(); return e.code; }
  catch (const std::exception& e) { ctx->c.last_error = e.what(); return SP_ERR_CUDA; }
  return snark_prove_common(ctx, inst, enc, d_vars.p, inputs, ninputs, gens, TranscriptArg{label, label_len, nullptr}, seed, proof, proof_len);
}
// SNARK::prove(..., transcript: &mut Transcript) with the caller's transcript state in / out (lib.rs:339-347)
int sp_snark_prove_t(sp_ctx* ctx, const sp_instance* inst, const sp_snark_encoding* enc, const uint64_t* vars, size_t nvars, const uint64_t* inputs, size_t ninputs,
                     const sp_snark_gens* gens, uint8_t* strobe_state, const uint64_t seed[4], uint8_t** proof, size_t* proof_len) {
  if (!strobe_state) { ctx->c.last_error = "transcript state is NULL"; return SP_ERR_INVALID_ARG; }
  DevBuf<u256> d_vars;
  try { d_vars = upload_padded_vars(ctx->c, inst->inst, vars, nvars); }
  catch (const SpError& e) { ctx->c.last_error = e.what(); return e.code; }
  catch (const std::exception& e) { ctx->c.last_error = e.what(); return SP_ERR_CUDA; }
  return snark_prove_common(ctx, inst, enc, d_vars.p, inputs, ninputs, gens, TranscriptArg{nullptr, 0, strobe_state}, seed, proof, proof_len);
}
int sp_snark_prove_resident(sp_ctx* ctx, const sp_instance* inst, const sp_snark_encoding* enc, const sp_poly* vars, const uint64_t* inputs, size_t ninputs,
                            const sp_snark_gens* gens, const uint8_t* label, size_t label_len, const uint64_t seed[4], uint8_t** proof, size_t* proof_len) {
  if (vars->len != inst->inst.num_vars) { ctx->c.last_error = "resident vars must already be padded to num_vars"; return SP_ERR_INVALID_INPUTS; }
  return snark_prove_common(ctx, inst, enc, vars->d.p, inputs, ninputs, gens, TranscriptArg{label, label_len, nullptr}, seed, proof, proof_len);
}

// ---- verifiers
int sp_nizk_verify(sp_ctx* ctx, const sp_instance* inst, const uint64_t* inputs, size_t ninputs, const sp_nizk_gens* gens, const uint8_t* label, size_t label_len,
                   const uint8_t* proof, size_t proof_len) {
  SP_TRY(ctx)
  require_reduced(inputs, ninputs, "inputs");
  Transcript T(std::string((const char*)label, label_len));
  nizk_verify(ctx->c, inst->inst, fq_vec(inputs, ninputs), *gens->g, T, proof, proof_len);
  SP_CATCH(ctx)
}
int sp_snark_verify(sp_ctx* ctx, const sp_snark_encoding* comm, const uint64_t* inputs, size_t ninputs, const sp_snark_gens* gens, const uint8_t* label,
                    size_t label_len, const uint8_t* proof, size_t proof_len) {
  SP_TRY(ctx)
  require_reduced(inputs, ninputs, "inputs");
  Transcript T(std::string((const char*)label, label_len));
  snark_verify(ctx->c, *comm->e, fq_vec(inputs, ninputs), *gens->g, T, proof, proof_len);
  SP_CATCH(ctx)
}

// verify(..., transcript: &mut Transcript) on the caller's transcript state (lib.rs:423-429, :549-555); the state is updated in place either way
int sp_nizk_verify_t(sp_ctx* ctx, const sp_instance* inst, const uint64_t* inputs, size_t ninputs, const sp_nizk_gens* gens, uint8_t* strobe_state,
                     const uint8_t* proof, size_t proof_len) {
  SP_TRY(ctx)
  if (!strobe_state) throw SpError(SP_ERR_INVALID_ARG, "transcript state is NULL");
  require_reduced(inputs, ninputs, "inputs");
  Transcript T(Transcript::FromState(), strobe_state);
  struct Out { Transcript& T; uint8_t* s; ~Out() { T.export_state(s); } } out{T, strobe_state};
  nizk_verify(ctx->c, inst->inst, fq_vec(inputs, ninputs), *gens->g, T, proof, proof_len);
  SP_CATCH(ctx)
}
int sp_snark_verify_t(sp_ctx* ctx, const sp_snark_encoding* comm, const uint64_t* inputs, size_t ninputs, const sp_snark_gens* gens, uint8_t* strobe_state,
                      const uint8_t* proof, size_t proof_len) {
  SP_TRY(ctx)
  if (!strobe_state) throw SpError(SP_ERR_INVALID_ARG, "transcript state is NULL");
  require_reduced(inputs, ninputs, "inputs");
  Transcript T(Transcript::FromState(), strobe_state);
  struct Out { Transcript& T; uint8_t* s; ~Out() { T.export_state(s); } } out{T, strobe_state};
  snark_verify(ctx->c, *comm->e, fq_vec(inputs, ninputs), *gens->g, T, proof, proof_len);
  SP_CATCH(ctx)
}
// merlin::Transcript for hosts without merlin (tests, C callers): the state buffer is the whole object
int sp_transcript_new(const uint8_t* label, size_t label_len, uint8_t* strobe_state) {
  Transcript T(std::string((const char*)label, label_len));
  T.export_state(strobe_state);
  return SP_OK;
}
int sp_transcript_append_message(uint8_t* strobe_state, const uint8_t* label, size_t label_len, const uint8_t* msg, size_t msg_len) {
  try {
    Transcript T(Transcript::FromState(), strobe_state);
    T.append_message(std::string((const char*)label, label_len).c_str(), msg, msg_len);
    T.export_state(strobe_state);
    return SP_OK;
  } catch (...) { return SP_ERR_INVALID_ARG; }
}
int sp_transcript_challenge_bytes(uint8_t* strobe_state, const uint8_t* label, size_t label_len, uint8_t* out, size_t n) {
  try {
    Transcript T(Transcript::FromState(), strobe_state);
    T.challenge_bytes(std::string((const char*)label, label_len).c_str(), out, n);
    T.export_state(strobe_state);
    return SP_OK;
  } catch (...) { return SP_ERR_INVALID_ARG; }
}

void sp_free(void* p) { free(p); }

}  // extern "C"
