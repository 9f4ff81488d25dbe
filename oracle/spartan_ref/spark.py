"""
ORACLE — TEST INFRASTRUCTURE ONLY (see core.py header).

Restatement of the SPARK sparse-polynomial commitment and the SNARK wrapper:
  src/sparse_mlpoly.rs, src/product_tree.rs, src/r1cs.rs (R1CSCommitment / R1CSEvalProof), src/lib.rs:277-465.
"""
import ctypes as C
from dataclasses import dataclass
from typing import List, Tuple

import numpy as np

from . import core as oc
from . import protocol as pr
from . import r1cs as r1
from .core import Q


# ----------------------------------------------------------------------------- product_tree.rs
class ProductCircuit:
    """product_tree.rs:11-63: left_vec / right_vec hold every layer"""

    def __init__(self, poly):
        n = len(poly)
        num_layers = oc.log_2(n)
        self.left_vec = [poly[: n // 2].copy()]
        self.right_vec = [poly[n // 2:].copy()]
        for i in range(num_layers - 1):
            prod = oc.hadamard(self.left_vec[i], self.right_vec[i])  # compute_layer, :18-34
            h = len(prod) // 2
            self.left_vec.append(prod[:h].copy())
            self.right_vec.append(prod[h:].copy())

    def evaluate(self):
        assert len(self.left_vec[-1]) == 1 and len(self.right_vec[-1]) == 1
        return oc.arr_get(self.left_vec[-1], 0) * oc.arr_get(self.right_vec[-1], 0) % Q


class DotProductCircuit:
    """product_tree.rs:66-108"""

    def __init__(self, left, right, weight):
        assert len(left) == len(right) == len(weight)
        self.left, self.right, self.weight = left, right, weight

    def evaluate(self):
        return oc.dot3(self.left, self.right, self.weight)

    def split(self):
        idx = len(self.left) // 2
        return (DotProductCircuit(self.left[:idx].copy(), self.right[:idx].copy(), self.weight[:idx].copy()),
                DotProductCircuit(self.left[idx:2 * idx].copy(), self.right[idx:2 * idx].copy(), self.weight[idx:2 * idx].copy()))


@dataclass
class LayerProofBatched:
    proof: pr.SumcheckInstanceProof
    claims_prod_left: List[int]
    claims_prod_right: List[int]


@dataclass
class ProductCircuitEvalProofBatched:
    proof: List[LayerProofBatched]
    claims_dotp: Tuple[List[int], List[int], List[int]]

    @staticmethod
    def prove(prod_circuit_vec, dotp_circuit_vec, T):
        """product_tree.rs:259-383"""
        assert prod_circuit_vec
        claims_dotp_final = ([], [], [])
        proof_layers = []
        num_layers = len(prod_circuit_vec[0].left_vec)
        claims_to_verify = [c.evaluate() for c in prod_circuit_vec]
        rand = []
        for layer_id in reversed(range(num_layers)):
            ln = len(prod_circuit_vec[0].left_vec[layer_id]) + len(prod_circuit_vec[0].right_vec[layer_id])
            poly_C_par = oc.eq_evals(rand)
            assert len(poly_C_par) == ln // 2
            num_rounds_prod = oc.log_2(len(poly_C_par))
            A_par = [c.left_vec[layer_id] for c in prod_circuit_vec]
            B_par = [c.right_vec[layer_id] for c in prod_circuit_vec]
            A_seq, B_seq, C_seq = [], [], []
            if layer_id == 0 and dotp_circuit_vec:
                for item in dotp_circuit_vec:
                    claims_to_verify.append(item.evaluate())
                    assert ln // 2 == len(item.left) == len(item.right) == len(item.weight)
                for d in dotp_circuit_vec:
                    A_seq.append(d.left); B_seq.append(d.right); C_seq.append(d.weight)
            coeff_vec = T.challenge_vector(b"rand_coeffs_next_layer", len(claims_to_verify))
            claim = sum(a * b for a, b in zip(claims_to_verify, coeff_vec)) % Q
            proof, rand_prod, claims_prod, claims_dotp = pr.prove_cubic_batched(claim, num_rounds_prod, (A_par, B_par, poly_C_par), (A_seq, B_seq, C_seq), coeff_vec, T)
            claims_prod_left, claims_prod_right, _claims_eq = claims_prod
            for i in range(len(prod_circuit_vec)):
                T.append_scalar(b"claim_prod_left", claims_prod_left[i])
                T.append_scalar(b"claim_prod_right", claims_prod_right[i])
            if layer_id == 0 and dotp_circuit_vec:
                cl, cr, cw = claims_dotp
                for i in range(len(dotp_circuit_vec)):
                    T.append_scalar(b"claim_dotp_left", cl[i])
                    T.append_scalar(b"claim_dotp_right", cr[i])
                    T.append_scalar(b"claim_dotp_weight", cw[i])
                claims_dotp_final = (cl, cr, cw)
            r_layer = T.challenge_scalar(b"challenge_r_layer")
            claims_to_verify = [(claims_prod_left[i] + r_layer * (claims_prod_right[i] - claims_prod_left[i])) % Q for i in range(len(prod_circuit_vec))]
            rand = [r_layer] + rand_prod
            proof_layers.append(LayerProofBatched(proof, claims_prod_left, claims_prod_right))
        return ProductCircuitEvalProofBatched(proof_layers, claims_dotp_final), rand

    def verify(self, claims_prod_vec, claims_dotp_vec, ln, T):
        """product_tree.rs:385-485"""
        num_layers = oc.log_2(ln)
        rand = []
        assert len(self.proof) == num_layers
        claims_to_verify = list(claims_prod_vec)
        claims_to_verify_dotp = []
        for num_rounds, i in enumerate(range(num_layers)):
            if i == num_layers - 1:
                claims_to_verify = claims_to_verify + list(claims_dotp_vec)
            coeff_vec = T.challenge_vector(b"rand_coeffs_next_layer", len(claims_to_verify))
            claim = sum(a * b for a, b in zip(claims_to_verify, coeff_vec)) % Q
            claim_last, rand_prod = self.proof[i].proof.verify(claim, num_rounds, 3, T)
            cpl, cpr = self.proof[i].claims_prod_left, self.proof[i].claims_prod_right
            assert len(cpl) == len(claims_prod_vec) == len(cpr)
            for k in range(len(claims_prod_vec)):
                T.append_scalar(b"claim_prod_left", cpl[k])
                T.append_scalar(b"claim_prod_right", cpr[k])
            assert len(rand) == len(rand_prod)
            eq = 1
            for a, b in zip(rand, rand_prod):
                eq = eq * (a * b + (1 - a) * (1 - b)) % Q
            claim_expected = sum(coeff_vec[k] * (cpl[k] * cpr[k] % Q * eq) for k in range(len(claims_prod_vec))) % Q
            if i == num_layers - 1:
                npi = len(claims_prod_vec)
                cl, cr, cw = self.claims_dotp
                for k in range(len(cl)):
                    T.append_scalar(b"claim_dotp_left", cl[k])
                    T.append_scalar(b"claim_dotp_right", cr[k])
                    T.append_scalar(b"claim_dotp_weight", cw[k])
                    claim_expected = (claim_expected + coeff_vec[k + npi] * cl[k] % Q * cr[k] % Q * cw[k]) % Q
            if claim_expected != claim_last % Q:
                raise pr.ProofVerifyError("product circuit layer %d" % i)
            r_layer = T.challenge_scalar(b"challenge_r_layer")
            claims_to_verify = [(cpl[k] + r_layer * (cpr[k] - cpl[k])) % Q for k in range(len(cpl))]
            if i == num_layers - 1:
                cl, cr, cw = self.claims_dotp
                for k in range(len(claims_dotp_vec) // 2):
                    claims_to_verify_dotp.append((cl[2 * k] + r_layer * (cl[2 * k + 1] - cl[2 * k])) % Q)
                    claims_to_verify_dotp.append((cr[2 * k] + r_layer * (cr[2 * k + 1] - cr[2 * k])) % Q)
                    claims_to_verify_dotp.append((cw[2 * k] + r_layer * (cw[2 * k + 1] - cw[2 * k])) % Q)
            rand = [r_layer] + rand_prod
        return claims_to_verify, claims_to_verify_dotp, rand


# ----------------------------------------------------------------------------- sparse_mlpoly.rs: dense representation
def merge(polys):
    """DensePolynomial::merge (dense_mlpoly.rs:259-272): concatenate and zero-pad to a power of two"""
    Z = np.concatenate(polys)
    n = oc.next_pow2(len(Z))
    if n > len(Z):
        Z = np.concatenate([Z, oc.zeros(n - len(Z))])
    return np.ascontiguousarray(Z)


class AddrTimestamps:
    """sparse_mlpoly.rs:213-272"""

    def __init__(self, num_cells, num_ops, ops_addr):
        audit = np.zeros(num_cells, dtype=np.uint64)
        self.ops_addr_usize = ops_addr
        self.ops_addr, self.read_ts = [], []
        for addr in ops_addr:
            assert len(addr) == num_ops
            addr = np.ascontiguousarray(addr, dtype=np.uint64)
            assert int(addr.max(initial=0)) < num_cells
            rts = np.zeros(num_ops, dtype=np.uint64)
            oc.lib.spark_timestamps(oc._ptr(rts), oc._ptr(audit), oc._ptr(addr), C.c_size_t(num_ops))
            self.ops_addr.append(oc.from_u64(addr))
            self.read_ts.append(oc.from_u64(rts))
        self.audit_ts = oc.from_u64(audit)

    def deref(self, mem_val):
        """sparse_mlpoly.rs:256-271"""
        return [np.ascontiguousarray(mem_val[a.astype(np.int64)]) for a in self.ops_addr_usize]


class MultiSparseMatPolynomialAsDense:
    """sparse_mlpoly.rs:274-282, built by multi_sparse_to_dense_rep (:366-420)"""

    def __init__(self, sparse_polys):
        for p in sparse_polys[1:]:
            assert p.num_vars_x == sparse_polys[0].num_vars_x and p.num_vars_y == sparse_polys[0].num_vars_y
        N = max(p.get_num_nz_entries() for p in sparse_polys)
        rows, cols, self.val = [], [], []
        for p in sparse_polys:
            r = np.zeros(N, dtype=np.uint64); c = np.zeros(N, dtype=np.uint64); v = oc.zeros(N)
            r[: p.nnz()] = p.row; c[: p.nnz()] = p.col; v[: p.nnz()] = p.val
            rows.append(r); cols.append(c); self.val.append(v)
        any_poly = sparse_polys[0]
        num_mem_cells = 1 << max(any_poly.num_vars_x, any_poly.num_vars_y)
        self.batch_size = len(sparse_polys)
        self.row = AddrTimestamps(num_mem_cells, N, rows)
        self.col = AddrTimestamps(num_mem_cells, N, cols)
        self.comb_ops = merge(self.row.ops_addr + self.row.read_ts + self.col.ops_addr + self.col.read_ts + self.val)
        self.comb_mem = np.ascontiguousarray(np.concatenate([self.row.audit_ts, self.col.audit_ts]))  # clone + extend (:410-411)

    def deref(self, row_mem_val, col_mem_val):
        return Derefs(self.row.deref(row_mem_val), self.col.deref(col_mem_val))


class Derefs:
    """sparse_mlpoly.rs:39-68"""

    def __init__(self, row_ops_val, col_ops_val):
        assert len(row_ops_val) == len(col_ops_val)
        self.row_ops_val, self.col_ops_val = row_ops_val, col_ops_val
        self.comb = merge(row_ops_val + col_ops_val)


class SparseMatPolyCommitmentGens:
    """sparse_mlpoly.rs:285-317"""

    def __init__(self, label, num_vars_x, num_vars_y, num_nz_entries, batch_size):
        num_vars_ops = oc.log_2(oc.next_pow2(num_nz_entries)) + oc.log_2(oc.next_pow2(batch_size * 5))
        num_vars_mem = max(num_vars_x, num_vars_y) + 1
        num_vars_derefs = oc.log_2(oc.next_pow2(num_nz_entries)) + oc.log_2(oc.next_pow2(batch_size * 2))
        self.gens_ops = pr.PolyCommitmentGens(num_vars_ops, label)
        self.gens_mem = pr.PolyCommitmentGens(num_vars_mem, label)
        self.gens_derefs = pr.PolyCommitmentGens(num_vars_derefs, label)


@dataclass
class SparseMatPolyCommitment:
    """sparse_mlpoly.rs:319-341 (usize fields are written as u64 by bincode)"""
    batch_size: int
    num_ops: int
    num_mem_cells: int
    comm_comb_ops: pr.PolyCommitment
    comm_comb_mem: pr.PolyCommitment

    def append_to_transcript(self, T):
        T.append_u64(b"batch_size", self.batch_size)
        T.append_u64(b"num_ops", self.num_ops)
        T.append_u64(b"num_mem_cells", self.num_mem_cells)
        self.comm_comb_ops.append_to_transcript(b"comm_comb_ops", T)
        self.comm_comb_mem.append_to_transcript(b"comm_comb_mem", T)

    def ser(self):
        return (self.batch_size.to_bytes(8, "little") + self.num_ops.to_bytes(8, "little") + self.num_mem_cells.to_bytes(8, "little")
                + pr.ser(self.comm_comb_ops) + pr.ser(self.comm_comb_mem))


def multi_commit(sparse_polys, gens):
    """SparseMatPolynomial::multi_commit (sparse_mlpoly.rs:483-503)"""
    dense = MultiSparseMatPolynomialAsDense(sparse_polys)
    comm_ops, _ = pr.poly_commit(dense.comb_ops, gens.gens_ops, None)
    comm_mem, _ = pr.poly_commit(dense.comb_mem, gens.gens_mem, None)
    return SparseMatPolyCommitment(len(sparse_polys), len(dense.row.read_ts[0]), len(dense.row.audit_ts), comm_ops, comm_mem), dense


# ----------------------------------------------------------------------------- layers
def build_hash_layer(eval_table, addrs_vec, derefs_vec, read_ts_vec, audit_ts, r_mem_check):
    """Layers::build_hash_layer (sparse_mlpoly.rs:529-604)"""
    r_hash, r_multiset = r_mem_check
    rh, rm = oc._fqp(r_hash), oc._fqp(r_multiset)
    n = len(eval_table)

    def hashed(count, addr, val, ts, plus_one):
        out = oc.zeros(count)
        oc.lib.spark_hash_layer(oc._ptr(out), C.c_size_t(count), oc._ptr(addr) if addr is not None else None, None, oc._ptr(np.ascontiguousarray(val)),
                                oc._ptr(ts) if ts is not None else None, C.c_int(plus_one), oc._ptr(rh), oc._ptr(rm))
        return out
    init = hashed(n, None, eval_table, None, 0)
    audit = hashed(n, None, eval_table, np.ascontiguousarray(audit_ts), 0)
    reads, writes = [], []
    for addrs, derefs, rts in zip(addrs_vec, derefs_vec, read_ts_vec):
        assert len(addrs) == len(derefs) == len(rts)
        reads.append(hashed(len(addrs), np.ascontiguousarray(addrs), derefs, np.ascontiguousarray(rts), 0))
        writes.append(hashed(len(addrs), np.ascontiguousarray(addrs), derefs, np.ascontiguousarray(rts), 1))
    return init, reads, writes, audit


class ProductLayer:
    def __init__(self, init, read_vec, write_vec, audit):
        self.init, self.read_vec, self.write_vec, self.audit = init, read_vec, write_vec, audit


def build_layers(eval_table, addr_timestamps, poly_ops_val, r_mem_check):
    """Layers::new (sparse_mlpoly.rs:606-653)"""
    init, reads, writes, audit = build_hash_layer(eval_table, addr_timestamps.ops_addr, poly_ops_val, addr_timestamps.read_ts, addr_timestamps.audit_ts, r_mem_check)
    layer = ProductLayer(ProductCircuit(init), [ProductCircuit(p) for p in reads], [ProductCircuit(p) for p in writes], ProductCircuit(audit))
    hw = layer.init.evaluate()
    for c in layer.write_vec:
        hw = hw * c.evaluate() % Q
    hr = layer.audit.evaluate()
    for c in layer.read_vec:
        hr = hr * c.evaluate() % Q
    assert hw == hr  # debug_assert_eq!(hashed_read_set, hashed_write_set)
    return layer


# ----------------------------------------------------------------------------- proofs
@dataclass
class DerefsEvalProof:
    proof_derefs: pr.PolyEvalProof

    @staticmethod
    def prove(derefs, eval_row, eval_col, r, gens, T, tape):
        """sparse_mlpoly.rs:125-149 + prove_single :80-123"""
        T.append_protocol_name(b"Derefs evaluation proof")
        evals = list(eval_row) + list(eval_col)
        evals += [0] * (oc.next_pow2(len(evals)) - len(evals))
        T.append_scalars(b"evals_ops_val", evals)
        challenges = T.challenge_vector(b"challenge_combine_n_to_one", oc.log_2(len(evals)))
        pe = list(evals)
        for i in reversed(range(len(challenges))):
            pe = oc.bound_bot_ints(pe, challenges[i])
        assert len(pe) == 1
        joint = pe[0]
        r_joint = challenges + list(r)
        T.append_scalar(b"joint_claim_eval", joint)
        proof, _ = pr.PolyEvalProof.prove(derefs.comb, None, r_joint, joint, None, gens, T, tape)
        return DerefsEvalProof(proof)

    def verify(self, r, eval_row, eval_col, gens, comm, T):
        """sparse_mlpoly.rs:151-211"""
        T.append_protocol_name(b"Derefs evaluation proof")
        evals = list(eval_row) + list(eval_col)
        evals += [0] * (oc.next_pow2(len(evals)) - len(evals))
        T.append_scalars(b"evals_ops_val", evals)
        challenges = T.challenge_vector(b"challenge_combine_n_to_one", oc.log_2(len(evals)))
        pe = list(evals)
        for i in reversed(range(len(challenges))):
            pe = oc.bound_bot_ints(pe, challenges[i])
        joint = pe[0]
        r_joint = challenges + list(r)
        T.append_scalar(b"joint_claim_eval", joint)
        self.proof_derefs.verify_plain(gens, T, r_joint, joint, comm)


@dataclass
class HashLayerProof:
    eval_row: Tuple[List[int], List[int], int]
    eval_col: Tuple[List[int], List[int], int]
    eval_val: List[int]
    eval_derefs: Tuple[List[int], List[int]]
    proof_ops: pr.PolyEvalProof
    proof_mem: pr.PolyEvalProof
    proof_derefs: DerefsEvalProof

    @staticmethod
    def _helper(rand, at):
        rand_mem, rand_ops = rand
        return ([oc.evaluate(a, rand_ops) for a in at.ops_addr], [oc.evaluate(a, rand_ops) for a in at.read_ts], oc.evaluate(at.audit_ts, rand_mem))

    @staticmethod
    def prove(rand, dense, derefs, gens, T, tape):
        """sparse_mlpoly.rs:722-835"""
        T.append_protocol_name(b"Sparse polynomial hash layer proof")
        rand_mem, rand_ops = rand
        eval_row_ops_val = [oc.evaluate(p, rand_ops) for p in derefs.row_ops_val]
        eval_col_ops_val = [oc.evaluate(p, rand_ops) for p in derefs.col_ops_val]
        proof_derefs = DerefsEvalProof.prove(derefs, eval_row_ops_val, eval_col_ops_val, rand_ops, gens.gens_derefs, T, tape)
        er = HashLayerProof._helper((rand_mem, rand_ops), dense.row)
        ec = HashLayerProof._helper((rand_mem, rand_ops), dense.col)
        eval_val_vec = [oc.evaluate(v, rand_ops) for v in dense.val]
        evals_ops = er[0] + er[1] + ec[0] + ec[1] + eval_val_vec
        evals_ops += [0] * (oc.next_pow2(len(evals_ops)) - len(evals_ops))
        T.append_scalars(b"claim_evals_ops", evals_ops)
        ch_ops = T.challenge_vector(b"challenge_combine_n_to_one", oc.log_2(len(evals_ops)))
        pe = list(evals_ops)
        for i in reversed(range(len(ch_ops))):
            pe = oc.bound_bot_ints(pe, ch_ops[i])
        joint_ops = pe[0]
        r_joint_ops = ch_ops + list(rand_ops)
        T.append_scalar(b"joint_claim_eval_ops", joint_ops)
        proof_ops, _ = pr.PolyEvalProof.prove(dense.comb_ops, None, r_joint_ops, joint_ops, None, gens.gens_ops, T, tape)
        evals_mem = [er[2], ec[2]]
        T.append_scalars(b"claim_evals_mem", evals_mem)
        ch_mem = T.challenge_vector(b"challenge_combine_two_to_one", 1)
        pe = oc.bound_bot_ints(evals_mem, ch_mem[0])
        joint_mem = pe[0]
        r_joint_mem = ch_mem + list(rand_mem)
        T.append_scalar(b"joint_claim_eval_mem", joint_mem)
        proof_mem, _ = pr.PolyEvalProof.prove(dense.comb_mem, None, r_joint_mem, joint_mem, None, gens.gens_mem, T, tape)
        return HashLayerProof(er, ec, eval_val_vec, (eval_row_ops_val, eval_col_ops_val), proof_ops, proof_mem, proof_derefs)

    @staticmethod
    def _verify_helper(rand, claims, eval_ops_val, eval_ops_addr, eval_read_ts, eval_audit_ts, r, r_hash, r_multiset):
        """sparse_mlpoly.rs:837-890"""
        r2 = r_hash * r_hash % Q

        def h(addr, val, ts):
            return (ts * r2 + val * r_hash + addr) % Q
        rand_mem, _ = rand
        claim_init, claim_read, claim_write, claim_audit = claims
        ln = len(rand_mem)
        eval_init_addr = sum((1 << (ln - i - 1)) * rand_mem[i] for i in range(ln)) % Q  # IdentityPolynomial::evaluate (dense_mlpoly.rs:105-115)
        eval_init_val = 1
        for a, b in zip(r, rand_mem):
            eval_init_val = eval_init_val * (a * b + (1 - a) * (1 - b)) % Q
        if (h(eval_init_addr, eval_init_val, 0) - r_multiset) % Q != claim_init % Q:
            raise pr.ProofVerifyError("hash init")
        for i in range(len(eval_ops_addr)):
            if (h(eval_ops_addr[i], eval_ops_val[i], eval_read_ts[i]) - r_multiset) % Q != claim_read[i] % Q:
                raise pr.ProofVerifyError("hash read")
            if (h(eval_ops_addr[i], eval_ops_val[i], eval_read_ts[i] + 1) - r_multiset) % Q != claim_write[i] % Q:
                raise pr.ProofVerifyError("hash write")
        if (h(eval_init_addr, eval_init_val, eval_audit_ts) - r_multiset) % Q != claim_audit % Q:
            raise pr.ProofVerifyError("hash audit")

    def verify(self, rand, claims_row, claims_col, claims_dotp, comm, gens, comm_derefs, rx, ry, r_hash, r_multiset, T):
        """sparse_mlpoly.rs:892-1019"""
        T.append_protocol_name(b"Sparse polynomial hash layer proof")
        rand_mem, rand_ops = rand
        eval_row_ops_val, eval_col_ops_val = self.eval_derefs
        assert len(eval_row_ops_val) == len(eval_col_ops_val)
        self.proof_derefs.verify(rand_ops, eval_row_ops_val, eval_col_ops_val, gens.gens_derefs, comm_derefs, T)
        eval_val_vec = self.eval_val
        assert len(claims_dotp) == 3 * len(eval_row_ops_val)
        for i in range(len(claims_dotp) // 3):
            if claims_dotp[3 * i] % Q != eval_row_ops_val[i] or claims_dotp[3 * i + 1] % Q != eval_col_ops_val[i] or claims_dotp[3 * i + 2] % Q != eval_val_vec[i]:
                raise pr.ProofVerifyError("dotp claims")
        er, ec = self.eval_row, self.eval_col
        evals_ops = er[0] + er[1] + ec[0] + ec[1] + eval_val_vec
        evals_ops = evals_ops + [0] * (oc.next_pow2(len(evals_ops)) - len(evals_ops))
        T.append_scalars(b"claim_evals_ops", evals_ops)
        ch_ops = T.challenge_vector(b"challenge_combine_n_to_one", oc.log_2(len(evals_ops)))
        pe = list(evals_ops)
        for i in reversed(range(len(ch_ops))):
            pe = oc.bound_bot_ints(pe, ch_ops[i])
        joint_ops = pe[0]
        T.append_scalar(b"joint_claim_eval_ops", joint_ops)
        self.proof_ops.verify_plain(gens.gens_ops, T, ch_ops + list(rand_ops), joint_ops, comm.comm_comb_ops)
        evals_mem = [er[2], ec[2]]
        T.append_scalars(b"claim_evals_mem", evals_mem)
        ch_mem = T.challenge_vector(b"challenge_combine_two_to_one", 1)
        joint_mem = oc.bound_bot_ints(evals_mem, ch_mem[0])[0]
        T.append_scalar(b"joint_claim_eval_mem", joint_mem)
        self.proof_mem.verify_plain(gens.gens_mem, T, ch_mem + list(rand_mem), joint_mem, comm.comm_comb_mem)
        HashLayerProof._verify_helper((rand_mem, rand_ops), claims_row, eval_row_ops_val, er[0], er[1], er[2], rx, r_hash, r_multiset)
        HashLayerProof._verify_helper((rand_mem, rand_ops), claims_col, eval_col_ops_val, ec[0], ec[1], ec[2], ry, r_hash, r_multiset)


@dataclass
class ProductLayerProof:
    eval_row: Tuple[int, List[int], List[int], int]
    eval_col: Tuple[int, List[int], List[int], int]
    eval_val: Tuple[List[int], List[int]]
    proof_mem: ProductCircuitEvalProofBatched
    proof_ops: ProductCircuitEvalProofBatched

    @staticmethod
    def prove(row_layer, col_layer, dense, derefs, evals, T):
        """sparse_mlpoly.rs:1035-1226"""
        T.append_protocol_name(b"Sparse polynomial product layer proof")

        def side(layer, name):
            init, audit = layer.init.evaluate(), layer.audit.evaluate()
            read = [c.evaluate() for c in layer.read_vec]
            write = [c.evaluate() for c in layer.write_vec]
            ws, rs = 1, 1
            for w in write:
                ws = ws * w % Q
            for r_ in read:
                rs = rs * r_ % Q
            assert init * ws % Q == rs * audit % Q
            T.append_scalar(b"claim_%s_eval_init" % name, init)
            T.append_scalars(b"claim_%s_eval_read" % name, read)
            T.append_scalars(b"claim_%s_eval_write" % name, write)
            T.append_scalar(b"claim_%s_eval_audit" % name, audit)
            return init, read, write, audit
        er = side(row_layer, b"row")
        ec = side(col_layer, b"col")
        assert len(evals) == len(derefs.row_ops_val) == len(derefs.col_ops_val) == len(dense.val)
        dl, dr, el, erv = [], [], [], []
        for i in range(len(derefs.row_ops_val)):
            circ = DotProductCircuit(derefs.row_ops_val[i].copy(), derefs.col_ops_val[i].copy(), dense.val[i].copy())
            left, right = circ.split()
            e_l, e_r = left.evaluate(), right.evaluate()
            T.append_scalar(b"claim_eval_dotp_left", e_l)
            T.append_scalar(b"claim_eval_dotp_right", e_r)
            assert (e_l + e_r) % Q == evals[i] % Q
            el.append(e_l); erv.append(e_r); dl.append(left); dr.append(right)
        assert len(row_layer.read_vec) == 3
        prods = (row_layer.read_vec + row_layer.write_vec + col_layer.read_vec + col_layer.write_vec)
        dotps = [dl[0], dr[0], dl[1], dr[1], dl[2], dr[2]]
        proof_ops, rand_ops = ProductCircuitEvalProofBatched.prove(prods, dotps, T)
        proof_mem, rand_mem = ProductCircuitEvalProofBatched.prove([row_layer.init, row_layer.audit, col_layer.init, col_layer.audit], [], T)
        return ProductLayerProof(er, ec, (el, erv), proof_mem, proof_ops), rand_mem, rand_ops

    def verify(self, num_ops, num_cells, evals, T):
        """sparse_mlpoly.rs:1228-1305"""
        T.append_protocol_name(b"Sparse polynomial product layer proof")
        num_instances = len(evals)

        def side(e, name):
            init, read, write, audit = e
            assert len(write) == num_instances == len(read)
            ws, rs = 1, 1
            for w in write:
                ws = ws * w % Q
            for r_ in read:
                rs = rs * r_ % Q
            if init * ws % Q != rs * audit % Q:
                raise pr.ProofVerifyError("subset check")
            T.append_scalar(b"claim_%s_eval_init" % name, init)
            T.append_scalars(b"claim_%s_eval_read" % name, read)
            T.append_scalars(b"claim_%s_eval_write" % name, write)
            T.append_scalar(b"claim_%s_eval_audit" % name, audit)
        side(self.eval_row, b"row")
        side(self.eval_col, b"col")
        el, er = self.eval_val
        assert len(el) == len(er) == num_instances
        claims_dotp_circuit = []
        for i in range(num_instances):
            if (el[i] + er[i]) % Q != evals[i] % Q:
                raise pr.ProofVerifyError("dotp split")
            T.append_scalar(b"claim_eval_dotp_left", el[i])
            T.append_scalar(b"claim_eval_dotp_right", er[i])
            claims_dotp_circuit += [el[i], er[i]]
        claims_prod = self.eval_row[1] + self.eval_row[2] + self.eval_col[1] + self.eval_col[2]
        claims_ops, claims_dotp, rand_ops = self.proof_ops.verify(claims_prod, claims_dotp_circuit, num_ops, T)
        claims_mem, _, rand_mem = self.proof_mem.verify([self.eval_row[0], self.eval_row[3], self.eval_col[0], self.eval_col[3]], [], num_cells, T)
        return claims_mem, rand_mem, claims_ops, claims_dotp, rand_ops


@dataclass
class PolyEvalNetworkProof:
    proof_prod_layer: ProductLayerProof
    proof_hash_layer: HashLayerProof

    @staticmethod
    def prove(row_layer, col_layer, dense, derefs, evals, gens, T, tape):
        """sparse_mlpoly.rs:1318-1354"""
        T.append_protocol_name(b"Sparse polynomial evaluation proof")
        ppl, rand_mem, rand_ops = ProductLayerProof.prove(row_layer, col_layer, dense, derefs, evals, T)
        phl = HashLayerProof.prove((rand_mem, rand_ops), dense, derefs, gens, T, tape)
        return PolyEvalNetworkProof(ppl, phl)

    def verify(self, comm, comm_derefs, evals, gens, rx, ry, r_mem_check, nz, T):
        """sparse_mlpoly.rs:1356-1416"""
        T.append_protocol_name(b"Sparse polynomial evaluation proof")
        ni = len(evals)
        r_hash, r_multiset = r_mem_check
        num_ops = oc.next_pow2(nz)
        num_cells = 1 << len(rx)
        assert len(rx) == len(ry)
        claims_mem, rand_mem, claims_ops, claims_dotp, rand_ops = self.proof_prod_layer.verify(num_ops, num_cells, evals, T)
        assert len(claims_mem) == 4 and len(claims_ops) == 4 * ni and len(claims_dotp) == 3 * ni
        row_read, row_write = claims_ops[:ni], claims_ops[ni:2 * ni]
        col_read, col_write = claims_ops[2 * ni:3 * ni], claims_ops[3 * ni:]
        self.proof_hash_layer.verify((rand_mem, rand_ops), (claims_mem[0], row_read, row_write, claims_mem[1]), (claims_mem[2], col_read, col_write, claims_mem[3]),
                                     claims_dotp, comm, gens, comm_derefs, rx, ry, r_hash, r_multiset, T)


def equalize(rx, ry):
    """sparse_mlpoly.rs:1429-1445"""
    if len(rx) < len(ry):
        return [0] * (len(ry) - len(rx)) + list(rx), list(ry)
    if len(rx) > len(ry):
        return list(rx), [0] * (len(rx) - len(ry)) + list(ry)
    return list(rx), list(ry)


@dataclass
class SparseMatPolyEvalProof:
    comm_derefs: pr.PolyCommitment   # DerefsCommitment { comm_ops_val }
    poly_eval_network_proof: PolyEvalNetworkProof

    @staticmethod
    def _append_derefs_comm(comm, T):
        """sparse_mlpoly.rs:213-219 (label passed down is b"comm_poly_row_col_ops_val")"""
        T.append_message(b"derefs_commitment", b"begin_derefs_commitment")
        comm.append_to_transcript(b"comm_poly_row_col_ops_val", T)
        T.append_message(b"derefs_commitment", b"end_derefs_commitment")

    @staticmethod
    def prove(dense, rx, ry, evals, gens, T, tape):
        """sparse_mlpoly.rs:1447-1514"""
        T.append_protocol_name(b"Sparse polynomial evaluation proof")
        assert len(evals) == dense.batch_size
        rx_ext, ry_ext = equalize(rx, ry)
        mem_rx, mem_ry = oc.eq_evals(rx_ext), oc.eq_evals(ry_ext)
        derefs = dense.deref(mem_rx, mem_ry)
        comm_derefs, _ = pr.poly_commit(derefs.comb, gens.gens_derefs, None)
        SparseMatPolyEvalProof._append_derefs_comm(comm_derefs, T)
        r_mem_check = T.challenge_vector(b"challenge_r_hash", 2)
        row_layer = build_layers(mem_rx, dense.row, derefs.row_ops_val, (r_mem_check[0], r_mem_check[1]))
        col_layer = build_layers(mem_ry, dense.col, derefs.col_ops_val, (r_mem_check[0], r_mem_check[1]))
        net = PolyEvalNetworkProof.prove(row_layer, col_layer, dense, derefs, evals, gens, T, tape)
        return SparseMatPolyEvalProof(comm_derefs, net)

    def verify(self, comm, rx, ry, evals, gens, T):
        """sparse_mlpoly.rs:1516-1557"""
        T.append_protocol_name(b"Sparse polynomial evaluation proof")
        rx_ext, ry_ext = equalize(rx, ry)
        nz, num_mem_cells = comm.num_ops, comm.num_mem_cells
        assert 1 << len(rx_ext) == num_mem_cells
        SparseMatPolyEvalProof._append_derefs_comm(self.comm_derefs, T)
        r_mem_check = T.challenge_vector(b"challenge_r_hash", 2)
        self.poly_eval_network_proof.verify(comm, self.comm_derefs, evals, gens, rx_ext, ry_ext, (r_mem_check[0], r_mem_check[1]), nz, T)


# ----------------------------------------------------------------------------- r1cs.rs commitment + lib.rs SNARK
class R1CSCommitmentGens:
    """r1cs.rs:28-47"""

    def __init__(self, label, num_cons, num_vars, num_inputs, num_nz_entries):
        assert num_inputs < num_vars
        self.gens = SparseMatPolyCommitmentGens(label, oc.log_2(num_cons), oc.log_2(2 * num_vars), num_nz_entries, 3)


@dataclass
class R1CSCommitment:
    num_cons: int
    num_vars: int
    num_inputs: int
    comm: SparseMatPolyCommitment

    def append_to_transcript(self, T):
        """r1cs.rs:58-65"""
        T.append_u64(b"num_cons", self.num_cons)
        T.append_u64(b"num_vars", self.num_vars)
        T.append_u64(b"num_inputs", self.num_inputs)
        self.comm.append_to_transcript(T)

    def ser(self):
        return self.num_cons.to_bytes(8, "little") + self.num_vars.to_bytes(8, "little") + self.num_inputs.to_bytes(8, "little") + self.comm.ser()


class SNARKGens:
    """lib.rs:277-309"""

    def __init__(self, num_cons, num_vars, num_inputs, num_nz_entries):
        nvp = oc.next_pow2(max(num_vars, num_inputs + 1))
        self.gens_r1cs_sat = pr.R1CSGens(b"gens_r1cs_sat", num_cons, nvp)
        self.gens_r1cs_eval = R1CSCommitmentGens(b"gens_r1cs_eval", num_cons, nvp, num_inputs, num_nz_entries)


@dataclass
class SNARK:
    r1cs_sat_proof: pr.R1CSProof
    inst_evals: Tuple[int, int, int]
    r1cs_eval_proof: SparseMatPolyEvalProof   # R1CSEvalProof { proof }

    @staticmethod
    def encode(inst, gens):
        """SNARK::encode (lib.rs:325-336) -> R1CSShape::commit (r1cs.rs:305-318)"""
        comm, dense = multi_commit([inst.inst.A, inst.inst.B, inst.inst.C], gens.gens_r1cs_eval.gens)
        return R1CSCommitment(inst.inst.num_cons, inst.inst.num_vars, inst.inst.num_inputs, comm), dense

    @staticmethod
    def prove(inst, comm, decomm, vars_arr, inputs, gens, T, seed_scalar):
        """lib.rs:339-420"""
        tape = oc.RandomTape(b"proof", seed_scalar)
        T.append_protocol_name(b"Spartan SNARK proof")
        comm.append_to_transcript(T)
        padded = r1.pad_vars(vars_arr, inst.inst.num_vars)
        sat, rx, ry = pr.R1CSProof.prove(inst.inst, padded, inputs, gens.gens_r1cs_sat, T, tape)
        Ar, Br, Cr = inst.inst.evaluate(rx, ry)
        T.append_scalar(b"Ar_claim", Ar)
        T.append_scalar(b"Br_claim", Br)
        T.append_scalar(b"Cr_claim", Cr)
        ev = SparseMatPolyEvalProof.prove(decomm, rx, ry, [Ar, Br, Cr], gens.gens_r1cs_eval.gens, T, tape)
        return SNARK(sat, (Ar, Br, Cr), ev)

    def verify(self, comm, inputs, T, gens):
        """lib.rs:423-465"""
        T.append_protocol_name(b"Spartan SNARK proof")
        comm.append_to_transcript(T)
        assert len(inputs) == comm.num_inputs
        rx, ry = self.r1cs_sat_proof.verify(comm.num_vars, comm.num_cons, inputs, self.inst_evals, T, gens.gens_r1cs_sat)
        Ar, Br, Cr = self.inst_evals
        T.append_scalar(b"Ar_claim", Ar)
        T.append_scalar(b"Br_claim", Br)
        T.append_scalar(b"Cr_claim", Cr)
        self.r1cs_eval_proof.verify(comm.comm, rx, ry, [Ar, Br, Cr], gens.gens_r1cs_eval.gens, T)

    def ser(self):
        return pr.ser(self)
