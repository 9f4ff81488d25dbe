// spartan_b200 — SNARK path: SPARK sparse-polynomial commitment (src/sparse_mlpoly.rs, src/product_tree.rs) driven on the device.
#pragma once
#include "prover.hpp"
#include <chrono>

namespace sp {

struct SnarkGens {  // lib.rs:277-309
  std::unique_ptr<R1CSGens> gens_r1cs_sat;
  std::unique_ptr<GenSet> eval_set;  // label "gens_r1cs_eval": ops / mem / derefs generators share one SHAKE stream (sparse_mlpoly.rs:292-317)
  PolyCommitmentGens gens_ops, gens_mem, gens_derefs;
  SnarkGens(Ctx* ctx, size_t num_cons, size_t num_vars, size_t num_inputs, size_t num_nz_entries);
};

struct AddrTimestampsDev {  // sparse_mlpoly.rs:213-272
  std::vector<DevBuf<uint32_t>> ops_addr_idx;  // per instance, N entries
  std::vector<DevBuf<u256>> ops_addr, read_ts; // as field elements
  DevBuf<u256> audit_ts;
};
struct SnarkEncoding {  // ComputationCommitment + ComputationDecommitment (lib.rs:44-55; sparse_mlpoly.rs:274-327)
  size_t num_cons = 0, num_vars = 0, num_inputs = 0;
  size_t batch_size = 0, num_ops = 0, num_mem_cells = 0;
  PolyCommitment comm_comb_ops, comm_comb_mem;
  std::vector<DevBuf<u256>> val;  // dense.val
  AddrTimestampsDev row, col;
  DevBuf<u256> comb_ops, comb_mem;
  void ser_commitment(Writer& w) const;
};

void snark_encode(Ctx& ctx, const Instance& inst, const SnarkGens& gens, SnarkEncoding& out);
void snark_prove(Ctx& ctx, const Instance& inst, const SnarkEncoding& enc, const u256* d_vars, const std::vector<Fq>& input, const SnarkGens& gens,
                 Transcript& T, const Fq& tape_seed, Writer& out);

// ---- verifiers (verifier.cpp): throw SpError with code SP_ERR_VERIFY / SP_ERR_DECOMPRESS when the proof is rejected
void instance_evaluate(Ctx& ctx, const Instance& inst, const std::vector<Fq>& rx, const std::vector<Fq>& ry, Fq out[3]);
void nizk_verify(Ctx& ctx, const Instance& inst, const std::vector<Fq>& input, const R1CSGens& gens, Transcript& T, const uint8_t* proof, size_t len);
void snark_verify(Ctx& ctx, const SnarkEncoding& comm, const std::vector<Fq>& input, const SnarkGens& gens, Transcript& T, const uint8_t* proof, size_t len);

}  // namespace sp
