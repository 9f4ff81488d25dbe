// spartan_b200 — SNARK path: SNARK::encode / SNARK::prove (src/lib.rs:325-420) with the SPARK sparse-polynomial evaluation proof
// (src/sparse_mlpoly.rs:1447-1514, src/product_tree.rs:259-383) driven on the device.  Every table of the memory-checking network
// (hash layers, 16 product trees with all their layers, dereferenced values) lives in HBM; the host only runs the transcript.
#include "snark.hpp"
#include <cstdlib>
#include <algorithm>
#include "../../include/spartan_b200.h"

namespace sp {

static size_t log2_ceil(size_t x) { size_t l = 0; while (((size_t)1 << l) < x) l++; return l; }
static size_t next_pow2(size_t x) { return (size_t)1 << log2_ceil(x); }

static PolyCommitmentGens pc_view(const GenSet* set, size_t num_vars) {  // PolyCommitmentGens::new (dense_mlpoly.rs:31-35)
  PolyCommitmentGens g;
  g.n = (size_t)1 << (num_vars - num_vars / 2);
  g.gens_n = CommitKey{set, 0, g.n, g.n + 1};
  g.gens_1 = CommitKey{set, g.n, 1, g.n + 1};
  return g;
}

SnarkGens::SnarkGens(Ctx* ctx, size_t num_cons, size_t num_vars, size_t num_inputs, size_t num_nz_entries) {
  size_t nvp = next_pow2(std::max(num_vars, num_inputs + 1));  // lib.rs:288-294
  gens_r1cs_sat.reset(new R1CSGens(ctx, "gens_r1cs_sat", nvp));
  // R1CSCommitmentGens::new (r1cs.rs:34-47) -> SparseMatPolyCommitmentGens::new(label, x, y, nz, 3) (sparse_mlpoly.rs:292-317)
  size_t x = log2_ceil(num_cons), y = log2_ceil(2 * nvp), lgnz = log2_ceil(next_pow2(num_nz_entries));
  size_t v_ops = lgnz + log2_ceil(next_pow2(3 * 5)), v_mem = std::max(x, y) + 1, v_derefs = lgnz + log2_ceil(next_pow2(3 * 2));
  size_t n_ops = (size_t)1 << (v_ops - v_ops / 2), n_mem = (size_t)1 << (v_mem - v_mem / 2), n_der = (size_t)1 << (v_derefs - v_derefs / 2);
  size_t nmax = std::max(n_ops, std::max(n_mem, n_der));
  eval_set.reset(new GenSet(ctx, "gens_r1cs_eval", nmax + 2, {n_ops, n_ops + 1, n_mem, n_mem + 1, n_der, n_der + 1}));
  gens_ops = pc_view(eval_set.get(), v_ops);
  gens_mem = pc_view(eval_set.get(), v_mem);
  gens_derefs = pc_view(eval_set.get(), v_derefs);
}

void SnarkEncoding::ser_commitment(Writer& w) const {  // ComputationCommitment { R1CSCommitment { .., SparseMatPolyCommitment } } (r1cs.rs:49-55, sparse_mlpoly.rs:319-327)
  w.u64(num_cons); w.u64(num_vars); w.u64(num_inputs);
  w.u64(batch_size); w.u64(num_ops); w.u64(num_mem_cells);
  comm_comb_ops.ser(w); comm_comb_mem.ser(w);
}

static void commit_poly(Ctx& ctx, const u256* d_Z, size_t len, const PolyCommitmentGens& gens, PolyCommitment& out) {  // DensePolynomial::commit, no blinds
  size_t ell = log2_ceil(len);
  size_t L = (size_t)1 << (ell / 2), R = (size_t)1 << (ell - ell / 2);
  if (R != gens.n) throw SpError(SP_ERR_INVALID_ARG, "polynomial size does not match the commitment generators (num_nz_entries too small?)");
  commit_rows_and_compress(ctx, gens.gens_n, d_Z, R, L, R, nullptr, out.C);
}

// ================================================================================================ SNARK::encode
void snark_encode(Ctx& ctx, const Instance& inst, const SnarkGens& gens, SnarkEncoding& e) {
  // SparseMatPolynomial::multi_commit -> multi_sparse_to_dense_rep (sparse_mlpoly.rs:366-420, :483-503)
  e.num_cons = inst.num_cons; e.num_vars = inst.num_vars; e.num_inputs = inst.num_inputs;
  e.batch_size = 3;
  size_t N = 1;
  for (int m = 0; m < 3; m++) N = std::max(N, next_pow2(inst.M[m].row.size()));
  size_t x = log2_ceil(inst.num_cons), y = log2_ceil(2 * inst.num_vars);
  size_t cells = (size_t)1 << std::max(x, y);
  e.num_ops = N; e.num_mem_cells = cells;
  e.comb_ops.alloc(16 * N);   // row.ops_addr[3] | row.read_ts[3] | col.ops_addr[3] | col.read_ts[3] | val[3] | zero pad  (DensePolynomial::merge)
  e.comb_mem.alloc(2 * cells);
  dev::dzero(e.comb_ops.p, 16 * N * sizeof(u256), ctx.stream);
  DevBuf<uint64_t> tmp64(std::max(N, cells));
  std::vector<uint64_t> audit(cells), rts(N), addr64(N);
  for (int side = 0; side < 2; side++) {  // AddrTimestamps::new (sparse_mlpoly.rs:220-254), row then col
    std::fill(audit.begin(), audit.end(), 0);
    AddrTimestampsDev& at = side == 0 ? e.row : e.col;
    at.ops_addr_idx.resize(3);
    for (int m = 0; m < 3; m++) {
      const std::vector<uint32_t>& src = side == 0 ? inst.M[m].row : inst.M[m].col;
      std::vector<uint32_t> a32(N, 0);
      std::copy(src.begin(), src.end(), a32.begin());
      for (size_t i = 0; i < N; i++) {
        size_t a = a32[i];
        if (a >= cells) throw SpError(SP_ERR_INVALID_INDEX, "address out of range");
        rts[i] = audit[a];
        audit[a] += 1;
        addr64[i] = a;
      }
      at.ops_addr_idx[m].alloc(N);
      dev::h2d(at.ops_addr_idx[m].p, a32.data(), N * sizeof(uint32_t), ctx.stream);
      u256* d_addr = e.comb_ops.p + (size_t)(side * 6 + m) * N;
      u256* d_rts = e.comb_ops.p + (size_t)(side * 6 + 3 + m) * N;
      dev::h2d(tmp64.p, addr64.data(), N * 8, ctx.stream);
      dev::from_u64(d_addr, tmp64.p, N, ctx.stream);
      dev::h2d(tmp64.p, rts.data(), N * 8, ctx.stream);
      dev::from_u64(d_rts, tmp64.p, N, ctx.stream);
      ctx.sync();
    }
    dev::h2d(tmp64.p, audit.data(), cells * 8, ctx.stream);
    dev::from_u64(e.comb_mem.p + (size_t)side * cells, tmp64.p, cells, ctx.stream);
    ctx.sync();
  }
  for (int m = 0; m < 3; m++)
    dev::h2d(e.comb_ops.p + (size_t)(12 + m) * N, inst.M[m].val.data(), inst.M[m].val.size() * sizeof(u256), ctx.stream);
  ctx.sync();
  commit_poly(ctx, e.comb_ops.p, 16 * N, gens.gens_ops, e.comm_comb_ops);
  commit_poly(ctx, e.comb_mem.p, 2 * cells, gens.gens_mem, e.comm_comb_mem);
}

// views into the dense representation
struct DenseView {
  size_t N, cells;
  const u256 *row_addr[3], *row_ts[3], *col_addr[3], *col_ts[3], *val[3], *row_audit, *col_audit;
  explicit DenseView(const SnarkEncoding& e) : N(e.num_ops), cells(e.num_mem_cells) {
    for (int m = 0; m < 3; m++) {
      row_addr[m] = e.comb_ops.p + (size_t)m * N; row_ts[m] = e.comb_ops.p + (size_t)(3 + m) * N;
      col_addr[m] = e.comb_ops.p + (size_t)(6 + m) * N; col_ts[m] = e.comb_ops.p + (size_t)(9 + m) * N;
      val[m] = e.comb_ops.p + (size_t)(12 + m) * N;
    }
    row_audit = e.comb_mem.p; col_audit = e.comb_mem.p + cells;
  }
};

// ProductCircuit (product_tree.rs:11-63): every layer kept; layer k has n/2^k entries (left half | right half).
// Single GPU: one 2n-scalar buffer, layer k at offset 2n - 2n/2^k.  Sharded over W ranks (cyclic partition: entry i of a layer on rank i mod W):
// the big layers [0, Ks) hold only this rank's n/(2^k W) entries (`loc`, same offset rule on the local sizes) — the Hadamard product that builds
// layer k+1 pairs i with i + len/2, which stay on one rank — and the layers from Ks on are replicated in `rep` (a circuit over n/2^Ks inputs).
struct ProdCircuit {
  DevBuf<u256> loc, rep;
  size_t n = 0, num_layers = 0, Ks = 0;
  int W = 1;
  void alloc(Ctx& ctx, size_t n_) {
    n = n_; num_layers = log2_ceil(n_); W = ctx.shard_world(); Ks = 0;
    // layer k's sumcheck runs on tables of n/2^(k+1) entries: shard it while those are worth sharding
    while (Ks + 1 < num_layers && ctx.shard_table((n >> Ks) / 2)) Ks++;
    if (Ks) loc.alloc(2 * (n / W));
    rep.alloc(2 * (n >> Ks));
  }
  bool sharded(size_t k) const { return k < Ks; }
  size_t layer_len(size_t k) const { return n >> k; }                                  // global entries
  size_t local_len(size_t k) const { return k < Ks ? (n >> k) / W : n >> k; }          // entries this rank holds
  u256* layer(size_t k) {
    if (k < Ks) { const size_t nl = n / W; return loc.p + (2 * nl - 2 * (nl >> k)); }
    const size_t nr = n >> Ks, kk = k - Ks;
    return rep.p + (2 * nr - 2 * (nr >> kk));
  }
};
// every layer of a group of equally sized circuits: one launch per layer (compute_layer, product_tree.rs:18-34); at the boundary between the
// sharded and the replicated layers the local products are all-gathered into every rank's copy
static void build_circuits(Ctx& ctx, std::vector<ProdCircuit*> cs) {
  ProdCircuit& c0 = *cs[0];
  for (size_t k = 0; k + 1 < c0.num_layers; k++) {
    const size_t h = c0.local_len(k) / 2;
    std::vector<u256*> outs; std::vector<const u256*> as, bs;
    if (k + 1 == c0.Ks) {
      DevBuf<u256> tmp(cs.size() * h);
      for (size_t i = 0; i < cs.size(); i++) { outs.push_back(tmp.p + i * h); as.push_back(cs[i]->layer(k)); bs.push_back(cs[i]->layer(k) + h); }
      dev::hadamard_many(outs.data(), as.data(), bs.data(), (int)cs.size(), h, ctx.stream);
      std::vector<const u256*> src(outs.begin(), outs.end());
      u256* g = ctx.allgather_cyclic(src.data(), (int)cs.size(), h);
      const size_t glen = h * (size_t)c0.W;
      for (size_t i = 0; i < cs.size(); i++) dev::d2d(cs[i]->layer(k + 1), g + i * glen, glen * sizeof(u256), ctx.stream);
      ctx.sync();   // tmp is released here
      continue;
    }
    for (auto* c : cs) { outs.push_back(c->layer(k + 1)); as.push_back(c->layer(k)); bs.push_back(c->layer(k) + h); }
    dev::hadamard_many(outs.data(), as.data(), bs.data(), (int)cs.size(), h, ctx.stream);
  }
}

// DotProductCircuit (product_tree.rs:66-108) of `len` entries; the tables the sumcheck binds are this rank's cyclic shards when `sharded`
struct DotpCircuit { u256 *left, *right, *weight; size_t len; bool sharded; Fq claim; };

// ProductCircuitEvalProofBatched::prove (product_tree.rs:259-383) with SumcheckInstanceProof::prove_cubic_batched (sumcheck.rs:254-424) inlined
static void batched_prove(Ctx& ctx, std::vector<ProdCircuit*>& prods, std::vector<DotpCircuit>& dotps, Transcript& T, ProductCircuitEvalProofBatched& out,
                          std::vector<Fq>& rand_out) {
  const size_t np = prods.size(), nd = dotps.size();
  const size_t num_layers = prods[0]->num_layers;
  std::vector<Fq> claims_to_verify(np);
  {  // all evaluate()s with one sync
    for (size_t i = 0; i < np; i++) dev::d2h(ctx.pinned + 64 * i, prods[i]->layer(num_layers - 1), 64, ctx.stream);
    ctx.sync();
    for (size_t i = 0; i < np; i++) { Fq v[2]; memcpy(v, ctx.pinned + 64 * i, 64); claims_to_verify[i] = v[0] * v[1]; }
  }
  const int W = prods[0]->W, rk = ctx.rank();
  const size_t Ks = prods[0]->Ks;
  // the shared eq table and its two ping-pong halves: local sizes while a layer is sharded, full sizes for the replicated layers
  const size_t max_half = std::max(prods[0]->local_len(0) / 2, Ks < num_layers ? prods[0]->local_len(Ks) / 2 : 1);
  DevBuf<u256> cpar_a(std::max<size_t>(max_half / 2, 1)), cpar_b(std::max<size_t>(max_half / 2, 1)), d_rand(64), eq_small(2 * ((size_t)1 << ((num_layers + 1) / 2)) + 8);
  DevBuf<u256> cpar0(std::max<size_t>(max_half, 1));
  std::vector<Fq> rand;
  // persistent tail (dev::sc_persist): once a layer's tables are small, one launch runs all its remaining rounds; A/B switch
#ifndef SP_SC_PERSIST_DEFAULT
#define SP_SC_PERSIST_DEFAULT 0
#endif
#ifndef SP_SC_EQFACTOR_DEFAULT
#define SP_SC_EQFACTOR_DEFAULT 1   // measured on the B200 (profiles/r02_tuning.md section 8): -1.3 ms per 2^20 proof, parity suite green
#endif
  static const bool eqfactor_on = [] { const char* e = getenv("SP_SC_EQFACTOR"); return e ? atoi(e) != 0 : SP_SC_EQFACTOR_DEFAULT != 0; }();
  const size_t G_MIN = 2 * Ctx::SHARD_MIN_LOCAL;   // 16384 entries per table: the fused G kernel then always streams >= 32768
  static const bool persist_on = [] { const char* e = getenv("SP_SC_PERSIST"); return e ? atoi(e) != 0 : SP_SC_PERSIST_DEFAULT != 0; }();
  DevBuf<u256> persist_c;
  if (persist_on) persist_c.alloc(np * (size_t)SC_PERSIST_MAX_LEN);
  u256* d_out = ctx.small.p + 64;   // ninst * 3 scalars
  DevBuf<u256> d_heads(64);
  for (size_t layer_id = num_layers; layer_id-- > 0;) {
    const size_t len = prods[0]->layer_len(layer_id);  // left + right
    const size_t half = len / 2;                       // table length of this layer's sumcheck (global)
    const size_t num_rounds = log2_ceil(half);
    bool sh = prods[0]->sharded(layer_id);             // this layer's tables are cyclic shards: the first rounds run sharded
    const size_t lhalf = sh ? half / W : half;         // ... of this many entries per rank
    // poly_C_par = eq(rand)                                                  (product_tree.rs:279-280)
    if (!rand.empty()) dev::h2d(d_rand.p, rand.data(), rand.size() * sizeof(u256), ctx.stream);
    int logW = 0;
    if (sh) while ((1 << logW) < W) logW++;
    const Fq c_rank = sh ? shard_eq_scale(rand, W, rk) : Fq::one();   // factor of eq(rand, .) that this rank's low index bits fix (cyclic shards)
    // eq-factored streaming rounds (kernels_sc.cu: k_sc_eval_g): rounds 0 .. jG-1, i.e. those whose evaluation streams >= G_MIN entries per table;
    // the suffix tables E_j = eq(rand[j+1..], .) live in cpar0 (E_0, then the levels 1 .. jG-1); from round jG on the standard kernels run on
    // C = prefix * eq(rand[jG..], .), built at the switch.  Off / not applicable: the standard method from round 0.
    size_t jG = 0;
    if (eqfactor_on && num_rounds >= 2) {
      while (jG + 1 < num_rounds && (lhalf >> jG) >= G_MIN) jG++;
      for (size_t j = 0; j < jG; j++) if (rand[j].is_zero()) jG = 0;   // q(1) is recovered by a division by tau_j
    }
    std::vector<const u256*> Elev(jG);
    if (jG) {
      const size_t n0 = lhalf / 2;
      dev::eq_evals(cpar0.p, d_rand.p + 1, (int)rand.size() - 1 - logW, eq_small.p, ctx.stream);
      if (sh) dev::scale(cpar0.p, c_rank.m, n0, ctx.stream);
      if (jG > 1) dev::eq_suffix(cpar0.p + n0, cpar0.p, n0, (int)jG - 1, ctx.stream);
      Elev[0] = cpar0.p;
      size_t off = n0;
      for (size_t k = 1; k < jG; k++) { Elev[k] = cpar0.p + off; off += n0 >> k; }
    } else if (sh) {
      dev::eq_evals(cpar0.p, d_rand.p, (int)rand.size() - logW, eq_small.p, ctx.stream);
      dev::scale(cpar0.p, c_rank.m, lhalf, ctx.stream);
    } else dev::eq_evals(cpar0.p, d_rand.p, (int)rand.size(), eq_small.p, ctx.stream);
    std::vector<dev::ScInst> insts;
    for (size_t i = 0; i < np; i++) {
      dev::ScInst in;
      in.t[0] = prods[i]->layer(layer_id); in.t[1] = prods[i]->layer(layer_id) + lhalf; in.t[2] = cpar0.p; in.t[3] = nullptr;
      in.c_out = cpar_a.p; in.write_c = i == 0;
      insts.push_back(in);
    }
    const bool with_dotp = layer_id == 0 && nd > 0;
    if (with_dotp) {
      for (size_t i = 0; i < nd; i++) {
        if (dotps[i].sharded != sh) throw SpError(SP_ERR_INTERNAL, "dot-product circuits and product circuits disagree on sharding");
        dev::ScInst in;
        in.t[0] = dotps[i].left; in.t[1] = dotps[i].right; in.t[2] = dotps[i].weight; in.t[3] = nullptr;
        in.c_out = dotps[i].weight; in.write_c = 1;
        insts.push_back(in);
        claims_to_verify.push_back(dotps[i].claim);   // DotProductCircuit::evaluate, taken on the unsharded tables by the caller
      }
    }
    const size_t ninst = insts.size();
    std::vector<Fq> coeff_vec = T.challenge_vector("rand_coeffs_next_layer", claims_to_verify.size());
    Fq claim = Fq::zero();
    for (size_t i = 0; i < claims_to_verify.size(); i++) claim += claims_to_verify[i] * coeff_vec[i];

    LayerProofBatched lp;
    std::vector<Fq> rand_prod;
    Fq e = claim;
    size_t cur = lhalf;                                // current length of the tables this rank holds
    dev::HostSig sig;
    auto xr_next = [&]() { return sh ? ctx.comm->next_xr() : dev::XRank(); };
    if (jG) {
      // ---- eq-factored rounds: the np product instances through the G kernels (2 values each), the dot-product circuits (their third table is not an
      // eq table) through the standard kernels in a second launch whose results land behind the first's
      const size_t ndi = ninst - np;
      auto launch_dotp = [&](bool first, const Fq& r) {
        if (!ndi) return;
        sig = ctx.next_sig();
        sig.host_out = ctx.host_res + 3 * np;
        if (first) dev::sc_eval(dev::SC_CUBIC3, insts.data() + np, (int)ndi, cur, d_out + 3 * np, ctx.red.p, ctx.stream, sig, xr_next());
        else dev::sc_fold_eval(dev::SC_CUBIC3, insts.data() + np, (int)ndi, cur, r.m, d_out + 3 * np, ctx.red.p, ctx.stream, sig, xr_next());
      };
      std::vector<Fq> tau_inv(jG);   // batch inversion of tau_0 .. tau_{jG-1}
      {
        std::vector<Fq> pre(jG);
        Fq acc = Fq::one();
        for (size_t j = 0; j < jG; j++) { pre[j] = acc; acc *= rand[j]; }
        Fq inv = acc.inv();
        for (size_t j = jG; j-- > 0;) { tau_inv[j] = inv * pre[j]; inv *= rand[j]; }
      }
      Fq cP = Fq::zero(), eD = Fq::zero(), prefix = Fq::one();
      for (size_t i = 0; i < np; i++) cP += claims_to_verify[i] * coeff_vec[i];
      for (size_t i = np; i < ninst; i++) eD += claims_to_verify[i] * coeff_vec[i];
      sig = ctx.next_sig();
      dev::sc_eval_g(insts.data(), (int)np, cur, Elev[0], d_out, ctx.red.p, ctx.stream, sig, xr_next());
      launch_dotp(true, Fq::zero());
      const Fq one = Fq::one();
      for (size_t j = 0; j < jG; j++) {
        std::vector<Fq> ev(3 * ninst);
        FineTimer fw(ctx, "batched wait evals");
        ctx.wait_sig(sig);
        fw.stop();
        FineTimer fh(ctx, "batched host round");
        memcpy(ev.data(), ctx.host_res, 3 * ninst * sizeof(u256));
        Fq Q0 = Fq::zero(), Qinf = Fq::zero(), D0 = Fq::zero(), D2 = Fq::zero(), D3 = Fq::zero();
        for (size_t i = 0; i < np; i++) { Q0 += ev[3 * i] * coeff_vec[i]; Qinf += ev[3 * i + 1] * coeff_vec[i]; }
        for (size_t i = np; i < ninst; i++) { D0 += ev[3 * i] * coeff_vec[i]; D2 += ev[3 * i + 1] * coeff_vec[i]; D3 += ev[3 * i + 2] * coeff_vec[i]; }
        const Fq& tau = rand[j];
        const Fq omt = one - tau;
        const Fq Q1 = (cP - omt * Q0) * tau_inv[j];            // c_j = (1 - tau_j) q(0) + tau_j q(1)
        const Fq qb = Q1 - Q0 - Qinf;
        auto q_at = [&](const Fq& t) { return (Qinf * t + qb) * t + Q0; };
        const Fq slope = tau - omt;                              // eq(tau_j, t) = (1 - tau_j) + t (2 tau_j - 1)
        std::vector<Fq> sD = {D0, eD - D0, D2, D3}, evals(4);
        Fq t = Fq::zero(), lin = omt;
        for (int k = 0; k < 4; k++) { evals[k] = prefix * lin * q_at(t) + (ndi ? sD[k] : Fq::zero()); t += one; lin += slope; }
        UniPoly poly = UniPoly::from_evals(evals);             // the same cubic the reference interpolates from its evaluations at 0, 1, 2, 3
        poly.append_to_transcript("poly", T);
        Fq r_j = T.challenge_scalar("challenge_nextround");
        rand_prod.push_back(r_j);
        cP = q_at(r_j);
        prefix *= omt + r_j * slope;
        if (ndi) eD = UniPoly::from_evals(sD).evaluate(r_j);
        e = poly.evaluate(r_j);
        lp.proof.compressed_polys.push_back(poly.compress());
        if (j + 1 < jG) {
          sig = ctx.next_sig();
          dev::sc_fold_eval_g(insts.data(), (int)np, cur, r_j.m, Elev[j + 1], d_out, ctx.red.p, ctx.stream, sig, xr_next());
          launch_dotp(false, r_j);
          cur >>= 1;
        } else {
          // switch: bind with r_j, build C = prefix * eq(rand[jG..], .) (this rank's slice when sharded) and evaluate round jG the standard way
          std::vector<u256*> tabs;
          for (size_t i = 0; i < ninst; i++) { tabs.push_back(insts[i].t[0]); tabs.push_back(insts[i].t[1]); if (i >= np) tabs.push_back(insts[i].t[2]); }
          dev::fold_top(tabs.data(), (int)tabs.size(), cur, r_j.m, ctx.stream);
          cur >>= 1;
          dev::eq_evals(cpar0.p, d_rand.p + jG, (int)rand.size() - (int)jG - logW, eq_small.p, ctx.stream);
          dev::scale(cpar0.p, (prefix * c_rank).m, cur, ctx.stream);
          sig = ctx.next_sig();
          dev::sc_eval(dev::SC_CUBIC3, insts.data(), (int)ninst, cur, d_out, ctx.red.p, ctx.stream, sig, xr_next());
        }
      }
    } else if (num_rounds > 0) { sig = ctx.next_sig(); dev::sc_eval(dev::SC_CUBIC3, insts.data(), (int)ninst, cur, d_out, ctx.red.p, ctx.stream, sig, xr_next()); }
    u256* cin = cpar0.p;
    u256* cpp[2] = {cpar_a.p, cpar_b.p};
    int flip = 0;
    bool persist = false;
    unsigned int persist_seq0 = 0;
    size_t persist_j0 = 0;
    for (size_t j = jG; j < num_rounds; j++) {
      std::vector<Fq> ev(3 * ninst);
      FineTimer fw(ctx, "batched wait evals");
      ctx.wait_sig(sig);
      fw.stop();
      FineTimer fh(ctx, "batched host round");
      memcpy(ev.data(), ctx.host_res, 3 * ninst * sizeof(u256));
      Fq c0 = Fq::zero(), c2 = Fq::zero(), c3 = Fq::zero();
      for (size_t i = 0; i < ninst; i++) { c0 += ev[3 * i] * coeff_vec[i]; c2 += ev[3 * i + 1] * coeff_vec[i]; c3 += ev[3 * i + 2] * coeff_vec[i]; }  // sumcheck.rs:359-361
      UniPoly poly = UniPoly::from_evals({c0, e - c0, c2, c3});
      poly.append_to_transcript("poly", T);
      Fq r_j = T.challenge_scalar("challenge_nextround");
      rand_prod.push_back(r_j);
      if (persist) {   // the persistent kernel binds (and evaluates the next round) as soon as it sees the challenge
        ctx.post_challenge(r_j);
        sig.seq = persist_seq0 + (unsigned int)(j - persist_j0);
        cur >>= 1;
        e = poly.evaluate(r_j);
        lp.proof.compressed_polys.push_back(poly.compress());
        continue;
      }
      // bind every table (shared C written once, through a ping-pong buffer)
      for (size_t i = 0; i < np; i++) { insts[i].t[2] = cin; insts[i].c_out = cpp[flip]; }
      if (persist_on && !sh && j + 1 < num_rounds && cur >= 4 && cur <= SC_PERSIST_MAX_LEN) {
        // small tables from here on: one launch for all the remaining rounds of this layer, its first bind with r_j
        const unsigned int nfold = (unsigned int)log2_ceil(cur);
        sig = ctx.next_sig();
        ctx.sig_seq += nfold - 1;        // one sequence number per publication: nfold-1 round evaluations, then the bound heads
        dev::sc_persist(insts.data(), (int)ninst, (int)np, persist_c.p, cur, r_j.m, ctx.mail, ctx.dmail.p, ctx.mail_seq, d_out, ctx.stream, sig);
        persist = true; persist_seq0 = sig.seq; persist_j0 = j;
        cur >>= 1;
        e = poly.evaluate(r_j);
        lp.proof.compressed_polys.push_back(poly.compress());
        continue;
      }
      if (j + 1 < num_rounds && sh && cur < Ctx::SHARD_MIN_LOCAL) {
        // leave the sharded stage: bind locally, all-gather every table into replicated copies (in the window), evaluate the next round there
        std::vector<u256*> tabs;
        for (size_t i = 0; i < ninst; i++) { tabs.push_back(insts[i].t[0]); tabs.push_back(insts[i].t[1]); if (i >= np) tabs.push_back(insts[i].t[2]); }
        tabs.push_back(cin);
        dev::fold_top(tabs.data(), (int)tabs.size(), cur, r_j.m, ctx.stream);
        const size_t glen = (cur / 2) * (size_t)W;
        std::vector<const u256*> src(tabs.begin(), tabs.end());
        u256* g = ctx.allgather_cyclic(src.data(), (int)src.size(), cur / 2);
        size_t k = 0;
        for (size_t i = 0; i < ninst; i++) {
          insts[i].t[0] = g + (k++) * glen; insts[i].t[1] = g + (k++) * glen;
          if (i >= np) { insts[i].t[2] = g + (k++) * glen; insts[i].c_out = insts[i].t[2]; }
        }
        cin = g + k * glen;
        for (size_t i = 0; i < np; i++) insts[i].t[2] = cin;
        sh = false;
        cur = glen;
        sig = ctx.next_sig();
        dev::sc_eval(dev::SC_CUBIC3, insts.data(), (int)ninst, cur, d_out, ctx.red.p, ctx.stream, sig);
        e = poly.evaluate(r_j);
        lp.proof.compressed_polys.push_back(poly.compress());
        continue;
      }
      if (j + 1 < num_rounds) { sig = ctx.next_sig(); dev::sc_fold_eval(dev::SC_CUBIC3, insts.data(), (int)ninst, cur, r_j.m, d_out, ctx.red.p, ctx.stream, sig, sh ? ctx.comm->next_xr() : dev::XRank()); }
      else {
        std::vector<u256*> tabs;
        for (size_t i = 0; i < ninst; i++) { tabs.push_back(insts[i].t[0]); tabs.push_back(insts[i].t[1]); if (i >= np) tabs.push_back(insts[i].t[2]); }
        dev::fold_top(tabs.data(), (int)tabs.size(), cur, r_j.m, ctx.stream);
        // the shared C's final value is never used by the prover (claims_prod.2 is dropped, product_tree.rs:336)
      }
      cin = cpp[flip];
      flip ^= 1;
      cur >>= 1;
      e = poly.evaluate(r_j);
      lp.proof.compressed_polys.push_back(poly.compress());
    }
    // final claims: first element of every A / B (and C for the dot-product circuits)
    FineTimer fl(ctx, "batched layer tail (claims d2h + transcript)");
    if (persist) ctx.wait_sig(sig);   // the heads arrive with the persistent kernel's last publication, in the layout of the gather below
    else {  // one gather kernel that also publishes the values to the host (instead of 2-3 tiny copies per instance and a stream synchronise)
      std::vector<const u256*> hp(3 * ninst);
      for (size_t i = 0; i < ninst; i++) { hp[3 * i] = insts[i].t[0]; hp[3 * i + 1] = insts[i].t[1]; hp[3 * i + 2] = i >= np ? insts[i].t[2] : insts[i].t[0]; }
      dev::HostSig hs = ctx.next_sig();
      dev::heads(d_heads.p, hp.data(), (int)hp.size(), ctx.stream, hs);
      ctx.wait_sig(hs);
    }
    const uint8_t* heads_h = reinterpret_cast<const uint8_t*>(ctx.host_res);
    lp.claims_prod_left.resize(np); lp.claims_prod_right.resize(np);
    for (size_t i = 0; i < np; i++) { memcpy(&lp.claims_prod_left[i], heads_h + 96 * i, 32); memcpy(&lp.claims_prod_right[i], heads_h + 96 * i + 32, 32); }
    for (size_t i = 0; i < np; i++) {
      T.append_scalar("claim_prod_left", lp.claims_prod_left[i]);
      T.append_scalar("claim_prod_right", lp.claims_prod_right[i]);
    }
    if (with_dotp) {
      out.dotp_left.resize(nd); out.dotp_right.resize(nd); out.dotp_weight.resize(nd);
      for (size_t i = 0; i < nd; i++) {
        memcpy(&out.dotp_left[i], heads_h + 96 * (np + i), 32);
        memcpy(&out.dotp_right[i], heads_h + 96 * (np + i) + 32, 32);
        memcpy(&out.dotp_weight[i], heads_h + 96 * (np + i) + 64, 32);
        T.append_scalar("claim_dotp_left", out.dotp_left[i]);
        T.append_scalar("claim_dotp_right", out.dotp_right[i]);
        T.append_scalar("claim_dotp_weight", out.dotp_weight[i]);
      }
    }
    Fq r_layer = T.challenge_scalar("challenge_r_layer");
    claims_to_verify.resize(np);
    for (size_t i = 0; i < np; i++) claims_to_verify[i] = lp.claims_prod_left[i] + r_layer * (lp.claims_prod_right[i] - lp.claims_prod_left[i]);
    std::vector<Fq> ext = {r_layer};
    ext.insert(ext.end(), rand_prod.begin(), rand_prod.end());
    rand = ext;
    out.proof.push_back(std::move(lp));
  }
  rand_out = rand;
}

static void append_scalar_vec(Transcript& T, const char* label, const std::vector<Fq>& v) { T.append_scalars(label, v); }

static std::vector<Fq> bound_bot_all(std::vector<Fq> v, const std::vector<Fq>& ch) {  // repeated bound_poly_var_bot (dense_mlpoly.rs:225-233), last challenge first
  for (size_t k = ch.size(); k-- > 0;) {
    size_t n = v.size() / 2;
    std::vector<Fq> o(n);
    for (size_t i = 0; i < n; i++) o[i] = v[2 * i] + ch[k] * (v[2 * i + 1] - v[2 * i]);
    v = o;
  }
  return v;
}

// ================================================================================================ SNARK::prove
void snark_prove(Ctx& ctx, const Instance& inst, const SnarkEncoding& enc, const u256* d_vars, const std::vector<Fq>& input, const SnarkGens& gens,
                 Transcript& T, const Fq& tape_seed, Writer& w) {
  ctx.timings.clear();
  ctx.timings.push_back({"info:background_stream_sms", (double)ctx.stream2_sms});
  ShardScope shard(ctx);   // on a connected multi-GPU context every rank runs this same function on the same inputs (see DESIGN.md "Multi-GPU")
  auto t_start = std::chrono::steady_clock::now();
  auto mark = [&](const char* name, std::chrono::steady_clock::time_point t0) {
    ctx.timings.push_back({name, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count()});
  };
  RandomTape tape("proof", tape_seed);                    // lib.rs:352
  T.append_protocol_name("Spartan SNARK proof");          // lib.rs:354
  // comm.comm.append_to_transcript (r1cs.rs:58-65, sparse_mlpoly.rs:329-341)
  T.append_u64("num_cons", enc.num_cons); T.append_u64("num_vars", enc.num_vars); T.append_u64("num_inputs", enc.num_inputs);
  T.append_u64("batch_size", enc.batch_size); T.append_u64("num_ops", enc.num_ops); T.append_u64("num_mem_cells", enc.num_mem_cells);
  append_poly_commitment(T, "comm_comb_ops", enc.comm_comb_ops);
  append_poly_commitment(T, "comm_comb_mem", enc.comm_comb_mem);

  const DenseView dv(enc);
  const size_t N = dv.N, cells = dv.cells;
  const size_t lgN = log2_ceil(N), lgC = log2_ceil(cells);
  DevBuf<u256> d_chal(64), eq_small(2 * ((size_t)1 << ((std::max(lgN + 4, lgC + 1) + 1) / 2)) + 8);

  // ---- dereferenced values and their commitment (sparse_mlpoly.rs:507-512, :213-219), started EARLY on the background stream.
  // comb = row_ops_val[3] | col_ops_val[3] | zero pad, committed as an L x R matrix (dense_mlpoly.rs:165-177).  row_ops_val depends only on rx
  // (known after the first sumcheck phase), col_ops_val only on ry (after the second), and the transcript needs the commitment only after the
  // whole R1CS proof: the two 3N-term MSMs (ALU-bound, the longest kernels of the proof) run on ctx.stream2 underneath the latency-bound rounds
  // of phase two and of the witness evaluation proof, which keep the greatest stream priority.  Same values, same bytes, different timing.
  DevBuf<u256> mem_rx(cells), mem_ry(cells), derefs(8 * N);
  dev::dzero(derefs.p + 6 * N, 2 * N * sizeof(u256), ctx.stream);
  u256 *row_val[3], *col_val[3];
  for (int m = 0; m < 3; m++) { row_val[m] = derefs.p + (size_t)m * N; col_val[m] = derefs.p + (size_t)(3 + m) * N; }
  const size_t ell_d = log2_ceil(8 * N), L_d = (size_t)1 << (ell_d / 2), R_d = (size_t)1 << (ell_d - ell_d / 2);
  if (R_d != gens.gens_derefs.n) throw SpError(SP_ERR_INVALID_ARG, "polynomial size does not match the commitment generators (num_nz_entries too small?)");
  static const bool early_ok = getenv("SP_NO_EARLY_DEREFS") == nullptr;
  const bool early = early_ok && ctx.overlap && N % R_d == 0;                    // each of the two parts is a whole number of matrix rows
  const size_t n_part = early ? 3 * N / R_d : 0;                  // rows per part; rows [2 n_part, L_d) are zero: the identity, 32 zero bytes
  const size_t Wd = (size_t)ctx.shard_world();
  const bool shard_rows = early && Wd > 1 && n_part % Wd == 0;    // rank r commits rows [r n_part/W, (r+1) n_part/W) of either part
  const size_t cnt = shard_rows ? n_part / Wd : n_part;
  const size_t nx = log2_ceil(inst.num_cons), ny = log2_ceil(2 * inst.num_vars), nxy = std::max(nx, ny);
  DevBuf<u256> d_chal2, eq_small2;
  DevBuf<ge> early_rows;
  DevBuf<uint8_t> early_comp;
  struct ForkGuard {   // an exception between fork and join must not release buffers the background stream still works on
    Ctx& c; bool forked = false;
    ~ForkGuard() { c.bg_busy = false; if (forked) { try { dev::stream_sync(c.stream2); } catch (...) {} } }
  } fork_guard{ctx};
  dev::MsmTune early_tune;
  {
    // measured on the B200 (profiles/r02_tuning.md section 6).  Shared SMs: three resident MSM CTAs per SM instead of four leave a quarter of every
    // register file to the prover's stream.  Partitioned (ctx.stream2_sms > 0): the background stream owns its SMs, full occupancy there.
    static const char* e_cpt = getenv("SP_EARLY_MSM_CPT");
    static const char* e_smem = getenv("SP_EARLY_MSM_SMEM");
    early_tune.smem_pad = ctx.stream2_sms > 0 ? 0 : (60 << 10);
    if (e_cpt) early_tune.cpt = atoi(e_cpt);
    if (e_smem) early_tune.smem_pad = (size_t)atol(e_smem);
  }
  if (early) {
    d_chal2.alloc(64); eq_small2.alloc(eq_small.n); early_rows.alloc(2 * cnt); early_comp.alloc(64 * cnt);
    const size_t need = dev::msm_scratch_bytes(cnt, R_d) + 64;
    if (ctx.scratch2.n < need) { dev::stream_sync(ctx.stream2); ctx.scratch2.alloc(need); }
  }
  auto early_part = [&](int part, const std::vector<Fq>& pt) {   // part 0: rows from rx, part 1: columns from ry
    // equalize (sparse_mlpoly.rs:1429-1445): left-pad the shorter point with zeros; staged in pinned memory so that the copy never waits for the stream
    Fq* ext = reinterpret_cast<Fq*>(ctx.pinned + (960 << 10) + (size_t)part * 4096);
    for (size_t i = 0; i < nxy - pt.size(); i++) ext[i] = Fq::zero();
    for (size_t i = 0; i < pt.size(); i++) ext[nxy - pt.size() + i] = pt[i];
    dev::event_record(ctx.ev_fork, ctx.stream);
    dev::stream_wait_event(ctx.stream2, ctx.ev_fork);
    fork_guard.forked = true;
    ctx.bg_busy = true;
    cudaStream_t s2 = ctx.stream2;
    u256* chal = d_chal2.p + 32 * part;
    u256* mem = part ? mem_ry.p : mem_rx.p;
    dev::h2d(chal, ext, nxy * sizeof(u256), s2);
    dev::eq_evals(mem, chal, (int)nxy, eq_small2.p, s2);
    for (int m = 0; m < 3; m++) dev::gather(part ? col_val[m] : row_val[m], mem, (part ? enc.col : enc.row).ops_addr_idx[m].p, N, s2);
    const size_t first = (size_t)part * n_part + (shard_rows ? (size_t)ctx.rank() * cnt : 0);
    const CommitKey& key = gens.gens_derefs.gens_n;
    dev::msm_rows(early_rows.p + (size_t)part * cnt, key.set->table.p, key.set->wbits, derefs.p + first * R_d, R_d, cnt, R_d, nullptr, key.h, ctx.scratch2.p, s2, early_tune);
    dev::compress_batch(early_comp.p + 32 * (size_t)part * cnt, early_rows.p + (size_t)part * cnt, cnt, s2);
  };
  R1csHooks hooks;
  if (early) {
    hooks.on_rx = [&](const std::vector<Fq>& p) { early_part(0, p); };
    hooks.on_ry = [&](const std::vector<Fq>& p) { early_part(1, p); };
  }

  R1CSProof sat;
  std::vector<Fq> rx, ry;
  r1cs_prove(ctx, inst, d_vars, input, *gens.gens_r1cs_sat, T, tape, sat, rx, ry, early ? &hooks : nullptr);

  // ---- inst.evaluate(rx, ry) (r1cs.rs:300-303 -> multi_evaluate sparse_mlpoly.rs:440-452)
  auto t0 = std::chrono::steady_clock::now();
  Fq inst_evals[3];
  {
    DevBuf<u256> trx((size_t)1 << rx.size()), try_((size_t)1 << ry.size());
    dev::h2d(d_chal.p, rx.data(), rx.size() * sizeof(u256), ctx.stream);
    dev::eq_evals(trx.p, d_chal.p, (int)rx.size(), eq_small.p, ctx.stream);
    dev::h2d(d_chal.p + 32, ry.data(), ry.size() * sizeof(u256), ctx.stream);
    dev::eq_evals(try_.p, d_chal.p + 32, (int)ry.size(), eq_small.p, ctx.stream);
    for (int m = 0; m < 3; m++)
      dev::sparse_eval3(ctx.small.p + 40 + m, inst.M[m].coo_row.p, inst.M[m].coo_col.p, inst.M[m].coo_val.p, inst.M[m].row.size(), trx.p, try_.p, ctx.red.p, ctx.stream);
    ctx.get_small(40, inst_evals, 3);
  }
  T.append_scalar("Ar_claim", inst_evals[0]);
  T.append_scalar("Br_claim", inst_evals[1]);
  T.append_scalar("Cr_claim", inst_evals[2]);
  mark("eval_sparse_polys", t0);

  // ---- R1CSEvalProof::prove -> SparseMatPolyEvalProof::prove (sparse_mlpoly.rs:1447-1514)
  auto t_eval = std::chrono::steady_clock::now();
  SparseMatPolyEvalProof ep;
  T.append_protocol_name("Sparse polynomial evaluation proof");
  if (nx != rx.size() || ny != ry.size()) throw std::runtime_error("spartan_b200: unexpected sumcheck point lengths");
  {
    auto tc = std::chrono::steady_clock::now();
    if (early) {
      // join: the prover's stream continues after the background commitments; gather the rank shares when the rows were split
      dev::event_record(ctx.ev_join, ctx.stream2);
      dev::stream_wait_event(ctx.stream, ctx.ev_join);
      std::vector<uint8_t> hb(64 * cnt * (shard_rows ? Wd : 1));
      if (shard_rows) dev::d2h(hb.data(), ctx.allgather_block(early_comp.p, 64 * cnt), hb.size(), ctx.stream);
      else dev::d2h(hb.data(), early_comp.p, hb.size(), ctx.stream);
      ctx.sync();
      fork_guard.forked = false;
      ctx.bg_busy = false;
      ep.comm_derefs.C.assign(L_d, Cp{});                                  // value-initialised: 32 zero bytes = the identity's encoding
      for (size_t r = 0; r < (shard_rows ? Wd : 1); r++)
        for (int part = 0; part < 2; part++)
          memcpy(ep.comm_derefs.C[(size_t)part * n_part + r * cnt].b, hb.data() + (r * 2 + part) * 32 * cnt, 32 * cnt);
    } else {
      // equalize (sparse_mlpoly.rs:1429-1445): left-pad the shorter point with zeros
      std::vector<Fq> rx_ext = rx, ry_ext = ry;
      if (rx.size() < ry.size()) rx_ext.insert(rx_ext.begin(), ry.size() - rx.size(), Fq::zero());
      if (ry.size() < rx.size()) ry_ext.insert(ry_ext.begin(), rx.size() - ry.size(), Fq::zero());
      dev::h2d(d_chal.p, rx_ext.data(), rx_ext.size() * sizeof(u256), ctx.stream);
      dev::eq_evals(mem_rx.p, d_chal.p, (int)rx_ext.size(), eq_small.p, ctx.stream);
      dev::h2d(d_chal.p + 32, ry_ext.data(), ry_ext.size() * sizeof(u256), ctx.stream);
      dev::eq_evals(mem_ry.p, d_chal.p + 32, (int)ry_ext.size(), eq_small.p, ctx.stream);
      for (int m = 0; m < 3; m++) {
        dev::gather(row_val[m], mem_rx.p, enc.row.ops_addr_idx[m].p, N, ctx.stream);
        dev::gather(col_val[m], mem_ry.p, enc.col.ops_addr_idx[m].p, N, ctx.stream);
      }
      commit_poly(ctx, derefs.p, 8 * N, gens.gens_derefs, ep.comm_derefs);
    }
    T.append_message("derefs_commitment", "begin_derefs_commitment");  // sparse_mlpoly.rs:213-219
    append_poly_commitment(T, "comm_poly_row_col_ops_val", ep.comm_derefs);
    T.append_message("derefs_commitment", "end_derefs_commitment");
    mark("commit_nondet_witness", tc);
  }
  std::vector<Fq> r_mem_check = T.challenge_vector("challenge_r_hash", 2);

  // ---- build_layered_network: hash layers (sparse_mlpoly.rs:529-604) written straight into layer 0 of the 16 product circuits
  auto tb = std::chrono::steady_clock::now();
  dev::h2d(d_chal.p, r_mem_check.data(), 2 * sizeof(u256), ctx.stream);
  struct Side { ProdCircuit init, audit, read[3], write[3]; } S[2];
  const int W = ctx.shard_world(), rk = ctx.rank();
  auto hash_into = [&](ProdCircuit& c, size_t n, const u256* addr, const u256* val, const u256* ts, int plus_one) {
    c.alloc(ctx, n);
    if (c.sharded(0)) dev::spark_hash_cyclic(c.layer(0), n / W, rk, W, addr, val, ts, plus_one, d_chal.p, ctx.stream);   // this rank's entries of the hash layer
    else dev::spark_hash(c.layer(0), n, addr, val, ts, plus_one, d_chal.p, ctx.stream);
  };
  for (int side = 0; side < 2; side++) {
    Side& s = S[side];
    const u256* mem = side == 0 ? mem_rx.p : mem_ry.p;
    hash_into(s.init, cells, nullptr, mem, nullptr, 0);
    hash_into(s.audit, cells, nullptr, mem, side == 0 ? dv.row_audit : dv.col_audit, 0);
    for (int m = 0; m < 3; m++) {
      const u256* addr = side == 0 ? dv.row_addr[m] : dv.col_addr[m];
      const u256* ts = side == 0 ? dv.row_ts[m] : dv.col_ts[m];
      const u256* val = side == 0 ? row_val[m] : col_val[m];
      hash_into(s.read[m], N, addr, val, ts, 0);
      hash_into(s.write[m], N, addr, val, ts, 1);
    }
  }
  // every layer of the 4 memory-sized and the 12 ops-sized product trees: one launch per layer (product_tree.rs:36-56)
  build_circuits(ctx, {&S[0].init, &S[0].audit, &S[1].init, &S[1].audit});
  build_circuits(ctx, {&S[0].read[0], &S[0].read[1], &S[0].read[2], &S[0].write[0], &S[0].write[1], &S[0].write[2],
                       &S[1].read[0], &S[1].read[1], &S[1].read[2], &S[1].write[0], &S[1].write[1], &S[1].write[2]});
  ctx.sync();
  mark("build_layered_network", tb);

  // ---- PolyEvalNetworkProof::prove (sparse_mlpoly.rs:1318-1354)
  auto tn = std::chrono::steady_clock::now();
  T.append_protocol_name("Sparse polynomial evaluation proof");
  ProductLayerProof& pl = ep.proof_prod_layer;
  T.append_protocol_name("Sparse polynomial product layer proof");  // ProductLayerProof::prove (sparse_mlpoly.rs:1035-1226)
  // ProductCircuit::evaluate of all 16 circuits (product_tree.rs:58-63): the two entries of every last layer through one gather kernel that
  // publishes them to the host, instead of 16 copies each followed by a stream synchronise
  Fq roots[2][8];
  {
    std::vector<const u256*> hp;
    for (int side = 0; side < 2; side++) {
      Side& s = S[side];
      ProdCircuit* cs[8] = {&s.init, &s.audit, &s.read[0], &s.read[1], &s.read[2], &s.write[0], &s.write[1], &s.write[2]};
      for (auto* c : cs) { hp.push_back(c->layer(c->num_layers - 1)); hp.push_back(c->layer(c->num_layers - 1) + 1); }
    }
    DevBuf<u256> d_roots(hp.size());
    dev::HostSig hs = ctx.next_sig();
    dev::heads(d_roots.p, hp.data(), (int)hp.size(), ctx.stream, hs);
    ctx.wait_sig(hs);
    Fq v[32];
    memcpy(v, ctx.host_res, sizeof v);
    for (int side = 0; side < 2; side++) for (int c = 0; c < 8; c++) roots[side][c] = v[16 * side + 2 * c] * v[16 * side + 2 * c + 1];
  }
  for (int side = 0; side < 2; side++) {
    Fq init = roots[side][0], audit = roots[side][1];
    std::vector<Fq> read = {roots[side][2], roots[side][3], roots[side][4]}, write = {roots[side][5], roots[side][6], roots[side][7]};
    Fq ws = write[0] * write[1] * write[2], rs = read[0] * read[1] * read[2];
    if (!(init * ws == rs * audit)) throw SpError(SP_ERR_INTERNAL, "memory-check subset test failed (sparse_mlpoly.rs:1060)");
    const char* li = side == 0 ? "claim_row_eval_init" : "claim_col_eval_init";
    const char* lr = side == 0 ? "claim_row_eval_read" : "claim_col_eval_read";
    const char* lw = side == 0 ? "claim_row_eval_write" : "claim_col_eval_write";
    const char* la = side == 0 ? "claim_row_eval_audit" : "claim_col_eval_audit";
    T.append_scalar(li, init); append_scalar_vec(T, lr, read); append_scalar_vec(T, lw, write); T.append_scalar(la, audit);
    if (side == 0) { pl.row_init = init; pl.row_read = read; pl.row_write = write; pl.row_audit = audit; }
    else { pl.col_init = init; pl.col_read = read; pl.col_write = write; pl.col_audit = audit; }
  }
  // dot-product circuits: clones of (row_ops_val, col_ops_val, val), split in halves (sparse_mlpoly.rs:1090-1117).  Their claims are taken on the
  // full tables; the copies the sumcheck binds are this rank's cyclic shards when the ops circuits are sharded.
  const bool sh_dotp = S[0].read[0].sharded(0);
  const size_t hN = N / 2, hloc = sh_dotp ? hN / W : hN;
  DevBuf<u256> dotp_buf(18 * hloc);
  std::vector<DotpCircuit> dotps;
  for (int m = 0; m < 3; m++) {
    const u256* srcs[3] = {row_val[m], col_val[m], dv.val[m]};
    dev::dot3(ctx.small.p + 48, srcs[0], srcs[1], srcs[2], hN, ctx.red.p, ctx.stream);
    dev::dot3(ctx.small.p + 49, srcs[0] + hN, srcs[1] + hN, srcs[2] + hN, hN, ctx.red.p, ctx.stream);
    Fq lr2[2];
    ctx.get_small(48, lr2, 2);
    T.append_scalar("claim_eval_dotp_left", lr2[0]);
    T.append_scalar("claim_eval_dotp_right", lr2[1]);
    if (!(lr2[0] + lr2[1] == inst_evals[m])) throw SpError(SP_ERR_INTERNAL, "dot-product circuit does not evaluate to the claimed matrix evaluation");
    pl.eval_dotp_left.push_back(lr2[0]); pl.eval_dotp_right.push_back(lr2[1]);
    for (int part = 0; part < 2; part++) {
      u256* base = dotp_buf.p + (size_t)(6 * m + 3 * part) * hloc;
      for (int t = 0; t < 3; t++) {
        if (sh_dotp) dev::take_cyclic(base + t * hloc, srcs[t] + part * hN, hloc, rk, W, ctx.stream);
        else dev::d2d(base + t * hloc, srcs[t] + part * hN, hN * sizeof(u256), ctx.stream);
      }
      dotps.push_back(DotpCircuit{base, base + hloc, base + 2 * hloc, hloc, sh_dotp, lr2[part]});
    }
  }
  std::vector<Fq> rand_ops, rand_mem;
  {
    std::vector<ProdCircuit*> prods = {&S[0].read[0], &S[0].read[1], &S[0].read[2], &S[0].write[0], &S[0].write[1], &S[0].write[2],
                                       &S[1].read[0], &S[1].read[1], &S[1].read[2], &S[1].write[0], &S[1].write[1], &S[1].write[2]};
    auto tq = std::chrono::steady_clock::now();
    batched_prove(ctx, prods, dotps, T, pl.proof_ops, rand_ops);
    mark("  product_layer_ops(12 trees + 6 dotp)", tq);
    tq = std::chrono::steady_clock::now();
    std::vector<ProdCircuit*> mems = {&S[0].init, &S[0].audit, &S[1].init, &S[1].audit};
    std::vector<DotpCircuit> none;
    batched_prove(ctx, mems, none, T, pl.proof_mem, rand_mem);
    mark("  product_layer_mem(4 trees)", tq);
  }
  auto th = std::chrono::steady_clock::now();

  // ---- HashLayerProof::prove (sparse_mlpoly.rs:722-835)
  HashLayerProof& hl = ep.proof_hash_layer;
  T.append_protocol_name("Sparse polynomial hash layer proof");
  {
    DevBuf<u256> eq_ops(N), eq_mem(cells);
    dev::h2d(d_chal.p, rand_ops.data(), rand_ops.size() * sizeof(u256), ctx.stream);
    dev::eq_evals(eq_ops.p, d_chal.p, (int)rand_ops.size(), eq_small.p, ctx.stream);
    dev::h2d(d_chal.p + 32, rand_mem.data(), rand_mem.size() * sizeof(u256), ctx.stream);
    dev::eq_evals(eq_mem.p, d_chal.p + 32, (int)rand_mem.size(), eq_small.p, ctx.stream);
    // 21 evaluations at rand_ops, 2 at rand_mem
    std::vector<const u256*> tabs;
    for (int m = 0; m < 3; m++) tabs.push_back(row_val[m]);
    for (int m = 0; m < 3; m++) tabs.push_back(col_val[m]);
    for (int m = 0; m < 3; m++) tabs.push_back(dv.row_addr[m]);
    for (int m = 0; m < 3; m++) tabs.push_back(dv.row_ts[m]);
    for (int m = 0; m < 3; m++) tabs.push_back(dv.col_addr[m]);
    for (int m = 0; m < 3; m++) tabs.push_back(dv.col_ts[m]);
    for (int m = 0; m < 3; m++) tabs.push_back(dv.val[m]);
    dev::dot_many(ctx.small.p + 64, tabs.data(), (int)tabs.size(), eq_ops.p, N, ctx.red.p, ctx.stream);
    const u256* mt[2] = {dv.row_audit, dv.col_audit};
    dev::dot_many(ctx.small.p + 64 + 21, mt, 2, eq_mem.p, cells, ctx.red.p, ctx.stream);
    std::vector<Fq> ev(23);
    ctx.get_small(64, ev.data(), 23);
    hl.derefs_row.assign(ev.begin(), ev.begin() + 3);
    hl.derefs_col.assign(ev.begin() + 3, ev.begin() + 6);
    hl.row_addr.assign(ev.begin() + 6, ev.begin() + 9);
    hl.row_read_ts.assign(ev.begin() + 9, ev.begin() + 12);
    hl.col_addr.assign(ev.begin() + 12, ev.begin() + 15);
    hl.col_read_ts.assign(ev.begin() + 15, ev.begin() + 18);
    hl.eval_val.assign(ev.begin() + 18, ev.begin() + 21);
    hl.row_audit_ts = ev[21]; hl.col_audit_ts = ev[22];
  }
  mark("  hash_layer_evaluations", th);
  th = std::chrono::steady_clock::now();
  Cp dummy;
  {  // DerefsEvalProof::prove (sparse_mlpoly.rs:125-149, prove_single :80-123)
    T.append_protocol_name("Derefs evaluation proof");
    std::vector<Fq> evals = hl.derefs_row;
    evals.insert(evals.end(), hl.derefs_col.begin(), hl.derefs_col.end());
    evals.resize(next_pow2(evals.size()), Fq::zero());
    T.append_scalars("evals_ops_val", evals);
    std::vector<Fq> ch = T.challenge_vector("challenge_combine_n_to_one", log2_ceil(evals.size()));
    Fq joint = bound_bot_all(evals, ch)[0];
    std::vector<Fq> r_joint = ch;
    r_joint.insert(r_joint.end(), rand_ops.begin(), rand_ops.end());
    T.append_scalar("joint_claim_eval", joint);
    polyeval_prove(ctx, derefs.p, nullptr, r_joint, joint, nullptr, gens.gens_derefs, T, tape, hl.proof_derefs, dummy);
  }
  mark("  polyeval_derefs(2^23)", th);
  th = std::chrono::steady_clock::now();
  {  // ops decommitment (sparse_mlpoly.rs:766-797)
    std::vector<Fq> evals;
    for (auto* v : {&hl.row_addr, &hl.row_read_ts, &hl.col_addr, &hl.col_read_ts, &hl.eval_val}) evals.insert(evals.end(), v->begin(), v->end());
    evals.resize(next_pow2(evals.size()), Fq::zero());
    T.append_scalars("claim_evals_ops", evals);
    std::vector<Fq> ch = T.challenge_vector("challenge_combine_n_to_one", log2_ceil(evals.size()));
    Fq joint = bound_bot_all(evals, ch)[0];
    std::vector<Fq> r_joint = ch;
    r_joint.insert(r_joint.end(), rand_ops.begin(), rand_ops.end());
    T.append_scalar("joint_claim_eval_ops", joint);
    polyeval_prove(ctx, enc.comb_ops.p, nullptr, r_joint, joint, nullptr, gens.gens_ops, T, tape, hl.proof_ops, dummy);
  }
  mark("  polyeval_ops(2^24)", th);
  th = std::chrono::steady_clock::now();
  {  // mem decommitment (sparse_mlpoly.rs:799-824)
    std::vector<Fq> evals = {hl.row_audit_ts, hl.col_audit_ts};
    T.append_scalars("claim_evals_mem", evals);
    std::vector<Fq> ch = T.challenge_vector("challenge_combine_two_to_one", 1);
    Fq joint = bound_bot_all(evals, ch)[0];
    std::vector<Fq> r_joint = ch;
    r_joint.insert(r_joint.end(), rand_mem.begin(), rand_mem.end());
    T.append_scalar("joint_claim_eval_mem", joint);
    polyeval_prove(ctx, enc.comb_mem.p, nullptr, r_joint, joint, nullptr, gens.gens_mem, T, tape, hl.proof_mem, dummy);
  }
  mark("  polyeval_mem(2^22)", th);
  mark("evalproof_layered_network", tn);
  mark("R1CSEvalProof::prove", t_eval);
  mark("SNARK::prove", t_start);
  for (auto& f : ctx.fine) ctx.timings.push_back({"fine:" + f.first, f.second});
  ctx.fine.clear();

  // bincode(SNARK { r1cs_sat_proof, inst_evals, r1cs_eval_proof }) (lib.rs:313-317)
  sat.ser(w);
  for (int m = 0; m < 3; m++) w.scalar(inst_evals[m]);
  ep.ser(w);
}

}  // namespace sp
