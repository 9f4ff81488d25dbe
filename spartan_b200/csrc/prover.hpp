// spartan_b200 — host prover: the reference's proof structs (wire order) and the GPU-driven provers.
// Struct and field names follow /root/reference/src so the bincode layout can be checked line by line.
#pragma once
#include <functional>
#include <array>
#include "engine.hpp"

namespace sp {

struct Writer {  // bincode 1.x default config
  std::vector<uint8_t> out;
  void u64(uint64_t x) { for (int i = 0; i < 8; i++) out.push_back((uint8_t)(x >> (8 * i))); }
  void scalar(const Fq& s) { uint8_t b[32]; u256_to_bytes(b, s.m); out.insert(out.end(), b, b + 32); }  // Montgomery limbs, not canonical bytes
  void point(const Cp& c) { out.insert(out.end(), c.b, c.b + 32); }
  void scalars(const std::vector<Fq>& v) { u64(v.size()); for (auto& s : v) scalar(s); }
  void points(const std::vector<Cp>& v) { u64(v.size()); for (auto& p : v) point(p); }
};

// ---- nizk/mod.rs
struct KnowledgeProof { Cp alpha; Fq z1, z2; void ser(Writer& w) const { w.point(alpha); w.scalar(z1); w.scalar(z2); } };
struct EqualityProof { Cp alpha; Fq z; void ser(Writer& w) const { w.point(alpha); w.scalar(z); } };
struct ProductProof { Cp alpha, beta, delta; std::array<Fq, 5> z; void ser(Writer& w) const { w.point(alpha); w.point(beta); w.point(delta); for (auto& s : z) w.scalar(s); } };
struct DotProductProof { Cp delta, beta; std::vector<Fq> z; Fq z_delta, z_beta; void ser(Writer& w) const { w.point(delta); w.point(beta); w.scalars(z); w.scalar(z_delta); w.scalar(z_beta); } };
struct BulletReductionProof { std::vector<Cp> L_vec, R_vec; void ser(Writer& w) const { w.points(L_vec); w.points(R_vec); } };
struct DotProductProofLog { BulletReductionProof bullet_reduction_proof; Cp delta, beta; Fq z1, z2;
  void ser(Writer& w) const { bullet_reduction_proof.ser(w); w.point(delta); w.point(beta); w.scalar(z1); w.scalar(z2); } };
// ---- dense_mlpoly.rs
struct PolyCommitment { std::vector<Cp> C; void ser(Writer& w) const { w.points(C); } };
struct PolyEvalProof { DotProductProofLog proof; void ser(Writer& w) const { proof.ser(w); } };
// ---- sumcheck.rs
struct ZKSumcheckInstanceProof { std::vector<Cp> comm_polys, comm_evals; std::vector<DotProductProof> proofs;
  void ser(Writer& w) const { w.points(comm_polys); w.points(comm_evals); w.u64(proofs.size()); for (auto& p : proofs) p.ser(w); } };
struct CompressedUniPoly { std::vector<Fq> coeffs_except_linear_term; void ser(Writer& w) const { w.scalars(coeffs_except_linear_term); } };
struct SumcheckInstanceProof { std::vector<CompressedUniPoly> compressed_polys; void ser(Writer& w) const { w.u64(compressed_polys.size()); for (auto& p : compressed_polys) p.ser(w); } };
// ---- r1csproof.rs:21-37
struct R1CSProof {
  PolyCommitment comm_vars;
  ZKSumcheckInstanceProof sc_proof_phase1;
  std::array<Cp, 4> claims_phase2;
  KnowledgeProof pok_Cz; ProductProof proof_prod;  // pok_claims_phase2
  EqualityProof proof_eq_sc_phase1;
  ZKSumcheckInstanceProof sc_proof_phase2;
  Cp comm_vars_at_ry;
  PolyEvalProof proof_eval_vars_at_ry;
  EqualityProof proof_eq_sc_phase2;
  void ser(Writer& w) const {
    comm_vars.ser(w); sc_proof_phase1.ser(w); for (auto& c : claims_phase2) w.point(c); pok_Cz.ser(w); proof_prod.ser(w);
    proof_eq_sc_phase1.ser(w); sc_proof_phase2.ser(w); w.point(comm_vars_at_ry); proof_eval_vars_at_ry.ser(w); proof_eq_sc_phase2.ser(w);
  }
};
// ---- product_tree.rs / sparse_mlpoly.rs (SNARK only)
struct LayerProofBatched { SumcheckInstanceProof proof; std::vector<Fq> claims_prod_left, claims_prod_right;
  void ser(Writer& w) const { proof.ser(w); w.scalars(claims_prod_left); w.scalars(claims_prod_right); } };
struct ProductCircuitEvalProofBatched { std::vector<LayerProofBatched> proof; std::vector<Fq> dotp_left, dotp_right, dotp_weight;
  void ser(Writer& w) const { w.u64(proof.size()); for (auto& l : proof) l.ser(w); w.scalars(dotp_left); w.scalars(dotp_right); w.scalars(dotp_weight); } };
struct ProductLayerProof {
  Fq row_init; std::vector<Fq> row_read, row_write; Fq row_audit;
  Fq col_init; std::vector<Fq> col_read, col_write; Fq col_audit;
  std::vector<Fq> eval_dotp_left, eval_dotp_right;
  ProductCircuitEvalProofBatched proof_mem, proof_ops;
  void ser(Writer& w) const {
    w.scalar(row_init); w.scalars(row_read); w.scalars(row_write); w.scalar(row_audit);
    w.scalar(col_init); w.scalars(col_read); w.scalars(col_write); w.scalar(col_audit);
    w.scalars(eval_dotp_left); w.scalars(eval_dotp_right); proof_mem.ser(w); proof_ops.ser(w);
  }
};
struct HashLayerProof {
  std::vector<Fq> row_addr, row_read_ts; Fq row_audit_ts;
  std::vector<Fq> col_addr, col_read_ts; Fq col_audit_ts;
  std::vector<Fq> eval_val, derefs_row, derefs_col;
  PolyEvalProof proof_ops, proof_mem, proof_derefs;
  void ser(Writer& w) const {
    w.scalars(row_addr); w.scalars(row_read_ts); w.scalar(row_audit_ts); w.scalars(col_addr); w.scalars(col_read_ts); w.scalar(col_audit_ts);
    w.scalars(eval_val); w.scalars(derefs_row); w.scalars(derefs_col); proof_ops.ser(w); proof_mem.ser(w); proof_derefs.ser(w);
  }
};
struct SparseMatPolyEvalProof {  // = R1CSEvalProof.proof
  PolyCommitment comm_derefs;    // DerefsCommitment.comm_ops_val
  ProductLayerProof proof_prod_layer;
  HashLayerProof proof_hash_layer;
  void ser(Writer& w) const { comm_derefs.ser(w); proof_prod_layer.ser(w); proof_hash_layer.ser(w); }
};

// ---- generator bundles
struct PolyCommitmentGens {  // dense_mlpoly.rs:24-36 -> DotProductProofGens (nizk/mod.rs:407-419)
  size_t n = 0;              // gens_n.n = 2^ceil(ell/2)
  CommitKey gens_n, gens_1;
};
struct R1CSGens {  // r1csproof.rs:39-74
  std::unique_ptr<GenSet> set;
  PolyCommitmentGens gens_pc;
  CommitKey gens_1, gens_3, gens_4;
  R1CSGens(Ctx* ctx, const std::string& label, size_t num_vars);
};

// ---- instance
struct SparseMatDev {  // one of A, B, C: COO on host (reference order), CSR + CSC on device
  std::vector<uint32_t> row, col;
  std::vector<Fq> val;
  DevBuf<uint32_t> csr_ptr, csr_idx, csc_ptr, csc_idx, coo_row, coo_col;
  DevBuf<u256> csr_val, csc_val, coo_val;
};
struct Instance {  // lib.rs:111-114 (R1CSShape + digest)
  size_t num_cons = 0, num_vars = 0, num_inputs = 0;
  SparseMatDev M[3];
  // R1CSShape::get_digest (r1cs.rs:154-158): zlib(level 6, miniz) of bincode(shape), computed on first use (deflate.cpp) unless the caller
  // supplied the bytes of its own compressor through sp_instance_set_digest
  mutable std::vector<uint8_t> digest;
  std::vector<uint8_t> shape_bincode() const;
  const std::vector<uint8_t>& shape_digest() const;
  void finalize(Ctx* ctx);  // build the device copies
};

std::vector<uint8_t> miniz_zlib_level6(const uint8_t* data, size_t len);   // deflate.cpp

struct NizkProof { R1CSProof r1cs_sat_proof; std::vector<Fq> rx, ry;
  void ser(Writer& w) const { r1cs_sat_proof.ser(w); w.scalars(rx); w.scalars(ry); } };

// R1CSProof::prove (r1csproof.rs:144-349).  d_vars: device array of num_vars Montgomery scalars (consumed read-only).
// hooks: called (when set) as soon as the first / the second sumcheck phase has produced its point, so that the caller can start work that
// depends only on rx / ry on another stream while the rest of the proof runs (SNARK::prove: the dereferenced SPARK values and their commitment)
struct R1csHooks { std::function<void(const std::vector<Fq>&)> on_rx, on_ry; };
void r1cs_prove(Ctx& ctx, const Instance& inst, const u256* d_vars, const std::vector<Fq>& input, const R1CSGens& gens, Transcript& T,
                RandomTape& tape, R1CSProof& proof, std::vector<Fq>& rx, std::vector<Fq>& ry, const R1csHooks* hooks = nullptr);
// NIZK::prove (lib.rs:501-546)
void nizk_prove(Ctx& ctx, const Instance& inst, const u256* d_vars, const std::vector<Fq>& input, const R1CSGens& gens, Transcript& T,
                const Fq& tape_seed, NizkProof& out);

struct UniPoly {  // unipoly.rs
  std::vector<Fq> coeffs;
  static UniPoly from_evals(const std::vector<Fq>& e) {  // unipoly.rs:23-54
    static const Fq two_inv = Fq::from_u64(2).inv(), six_inv = Fq::from_u64(6).inv();
    UniPoly p;
    if (e.size() == 3) {
      Fq c = e[0];
      Fq a = two_inv * (e[2] - e[1] - e[1] + c);
      Fq b = e[1] - c - a;
      p.coeffs = {c, b, a};
    } else {
      Fq d = e[0];
      Fq a = six_inv * (e[3] - e[2] - e[2] - e[2] + e[1] + e[1] + e[1] - e[0]);
      Fq b = two_inv * (e[0] + e[0] - e[1] - e[1] - e[1] - e[1] - e[1] + e[2] + e[2] + e[2] + e[2] - e[3]);
      Fq c = e[1] - d - a - b;
      p.coeffs = {d, c, b, a};
    }
    return p;
  }
  Fq evaluate(const Fq& r) const {  // unipoly.rs:72-80
    Fq ev = coeffs[0], power = r;
    for (size_t i = 1; i < coeffs.size(); i++) { ev += power * coeffs[i]; power *= r; }
    return ev;
  }
  CompressedUniPoly compress() const {  // unipoly.rs:82-88
    CompressedUniPoly c;
    c.coeffs_except_linear_term.push_back(coeffs[0]);
    for (size_t i = 2; i < coeffs.size(); i++) c.coeffs_except_linear_term.push_back(coeffs[i]);
    return c;
  }
  void append_to_transcript(const char* label, Transcript& T) const {  // unipoly.rs:112-120
    T.append_message(label, "UniPoly_begin");
    for (auto& c : coeffs) T.append_scalar("coeff", c);
    T.append_message(label, "UniPoly_end");
  }
};

void append_poly_commitment(Transcript& T, const char* label, const PolyCommitment& c);
// PolyEvalProof::prove (dense_mlpoly.rs:312-365) on a device-resident table
void polyeval_prove(Ctx& ctx, const u256* d_Z, const std::vector<Fq>* blinds_opt, const std::vector<Fq>& r, const Fq& Zr, const Fq* blind_Zr_opt,
                    const PolyCommitmentGens& gens, Transcript& T, RandomTape& tape, PolyEvalProof& proof, Cp& C_Zr);

// exposed pieces (C-ABI operator level and tests)
Cp commit_rows_and_compress(Ctx& ctx, const CommitKey& key, const u256* d_scalars, size_t stride, size_t L, size_t R, const Fq* blinds,
                            std::vector<Cp>& out, const std::function<const Fq*()>& blinds_late = nullptr);
std::vector<Fq> host_eq_evals(const std::vector<Fq>& r);

}  // namespace sp
