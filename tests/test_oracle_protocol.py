"""Oracle protocol layer (oracle/spartan_ref): prove -> verify round trips exactly like the reference's own integration tests
(SURVEY.md §4), the polynomial known answers, and the structural pins (bincode lengths reproduced from README.md:362,371,374)."""
import numpy as np
import pytest

from oracle.spartan_ref import core as oc
from oracle.spartan_ref import protocol as pr
from oracle.spartan_ref import r1cs, spark

Q = oc.Q


def test_dense_evaluate_known_answer():
    """dense_mlpoly.rs:434-452: Z = [1,2,1,4], r = [4,3] -> 28, and the L/R factorisation gives the same"""
    Z = oc.to_arr([1, 2, 1, 4])
    r = [4, 3]
    assert oc.evaluate(Z, r) == 28
    L, R = oc.eq_evals(r[:1]), oc.eq_evals(r[1:])
    LZ = oc.bound_rows(Z, L, 2, 2)
    assert oc.dot(LZ, R) == 28


def test_eq_tables_vs_naive():
    """dense_mlpoly.rs:523-565: memoised eq tables equal the naive bit-product"""
    for ell in range(0, 8):
        r = oc.to_ints(oc.prg_scalars("eqt", ell, ell)) if ell else []
        table = oc.to_ints(oc.eq_evals(r))
        for i in range(1 << ell):
            want = 1
            for j in range(ell):
                bit = (i >> (ell - 1 - j)) & 1
                want = want * (r[j] if bit else (1 - r[j])) % Q
            assert table[i] == want
        lv = ell // 2
        L, R = oc.to_ints(oc.eq_evals(r[:lv])), oc.to_ints(oc.eq_evals(r[lv:]))
        assert table == [l * rr % Q for l in L for rr in R]


def test_unipoly_known_answers():
    """unipoly.rs:128-181"""
    p = pr.UniPoly.from_evals([1, 6, 15])
    assert p.coeffs == [1, 3, 2] and p.evaluate(3) == 28
    assert p.compress().decompress(1 + 6).coeffs == p.coeffs
    p = pr.UniPoly.from_evals([1, 7, 23, 55])
    assert p.coeffs == [1, 3, 2, 1] and p.evaluate(4) == 109
    assert p.compress().decompress(1 + 7).coeffs == p.coeffs


def test_sigma_protocols_roundtrip():
    """nizk/mod.rs:585-735 style: knowledge / equality / product / dot-product / dot-product-log proofs verify"""
    tape = oc.RandomTape(b"proof", 12345)
    gens_1 = oc.MultiCommitGens.new(1, b"test-two")
    x, r = 11111, 22222
    proof, Cc = pr.KnowledgeProof.prove(gens_1, oc.Transcript(b"example"), tape, x, r)
    proof.verify(gens_1, oc.Transcript(b"example"), Cc)
    proof, C1, C2 = pr.EqualityProof.prove(gens_1, oc.Transcript(b"example"), tape, x, 5, x, 7)
    proof.verify(gens_1, oc.Transcript(b"example"), C1, C2)
    y = 987654321
    proof, X, Y, Z = pr.ProductProof.prove(gens_1, oc.Transcript(b"example"), tape, x, 3, y, 4, x * y % Q, 5)
    proof.verify(gens_1, oc.Transcript(b"example"), X, Y, Z)
    with pytest.raises(pr.ProofVerifyError):
        proof.verify(gens_1, oc.Transcript(b"example"), X, Y, X)
    n = 1024
    gens_n = oc.MultiCommitGens.new(n, b"test-1024")
    xs = oc.to_ints(oc.prg_scalars("x", n))
    a = oc.to_ints(oc.prg_scalars("a", n))
    yv = sum(p * q for p, q in zip(xs, a)) % Q
    proof, Cx, Cy = pr.DotProductProof.prove(gens_1, gens_n, oc.Transcript(b"example"), tape, xs, 9, a, yv, 8)
    proof.verify(gens_1, gens_n, oc.Transcript(b"example"), a, Cx, Cy)
    gens = pr.DotProductProofGens(n, b"test-1024")
    proof, Cx, Cy = pr.DotProductProofLog.prove(gens, oc.Transcript(b"example"), tape, xs, 9, a, yv, 8)
    proof.verify(n, gens, oc.Transcript(b"example"), a, Cx, Cy)
    with pytest.raises(pr.ProofVerifyError):
        proof.verify(n, gens, oc.Transcript(b"other"), a, Cx, Cy)


def test_polycommit_roundtrip():
    """dense_mlpoly.rs:568-602"""
    Z = oc.to_arr([1, 2, 1, 4])
    r = [4, 3]
    gens = pr.PolyCommitmentGens(2, b"test-two")
    tape = oc.RandomTape(b"proof", 77)
    comm, blinds = pr.poly_commit(Z, gens, tape)
    proof, C_Zr = pr.PolyEvalProof.prove(Z, blinds, r, 28, 5, gens, oc.Transcript(b"example"), tape)
    proof.verify(gens, oc.Transcript(b"example"), r, C_Zr, comm)


def _len_r1cs_sat(ell_vars, rounds_x, rounds_y):
    """SURVEY Appendix B"""
    L = 1 << (ell_vars // 2)
    lgR = ell_vars - ell_vars // 2
    comm_vars = 8 + L * 32
    ph1 = 2 * (8 + rounds_x * 32) + 8 + rounds_x * (32 + 32 + 8 + 4 * 32 + 64)
    ph2 = 2 * (8 + rounds_y * 32) + 8 + rounds_y * (32 + 32 + 8 + 3 * 32 + 64)
    return comm_vars + ph1 + 128 + (96 + 256) + 64 + ph2 + 32 + (2 * (8 + lgR * 32) + 128) + 64


def test_published_proof_lengths_follow_from_layout():
    """README.md:362 len_r1cs_sat_proof 47024 at 2^20; SURVEY §8c item 6: 8720 at 2^10"""
    assert _len_r1cs_sat(20, 20, 21) == 47024
    assert _len_r1cs_sat(10, 10, 11) == 8720


@pytest.mark.parametrize("num_cons,num_vars,num_inputs", [(1024, 1024, 10), (16, 16, 3), (64, 256, 7), (512, 32, 5), (2, 2, 1)])
def test_nizk_roundtrip_and_layout(num_cons, num_vars, num_inputs):
    """r1csproof.rs:570-602 / lib.rs NIZK: synthetic instance proves and verifies; bincode length matches the layout formula"""
    inst, vars_arr, inputs = r1cs.Instance.produce_synthetic_r1cs(num_cons, num_vars, num_inputs, 5)
    assert inst.is_sat(vars_arr, inputs)
    gens = r1cs.NIZKGens(num_cons, num_vars, num_inputs)
    proof = r1cs.NIZK.prove(inst, vars_arr, inputs, gens, oc.Transcript(b"nizk_example"), r1cs.tape_seed(5))
    proof.verify(inst, inputs, oc.Transcript(b"nizk_example"), gens)
    rx, ry = oc.log_2(num_cons), oc.log_2(2 * num_vars)
    assert len(pr.ser(proof.r1cs_sat_proof)) == _len_r1cs_sat(oc.log_2(num_vars), rx, ry)
    assert len(proof.ser()) == _len_r1cs_sat(oc.log_2(num_vars), rx, ry) + 16 + 32 * (rx + ry)
    # a different transcript label or a tampered proof must be rejected
    with pytest.raises((pr.ProofVerifyError, AssertionError)):
        proof.verify(inst, inputs, oc.Transcript(b"other"), gens)
    bad_inputs = list(inputs)
    bad_inputs[0] = (bad_inputs[0] + 1) % Q
    with pytest.raises((pr.ProofVerifyError, AssertionError)):
        proof.verify(inst, bad_inputs, oc.Transcript(b"nizk_example"), gens)


def test_nizk_is_deterministic_in_its_explicit_inputs():
    inst, vars_arr, inputs = r1cs.Instance.produce_synthetic_r1cs(64, 64, 4, 1)
    gens = r1cs.NIZKGens(64, 64, 4)
    a = r1cs.NIZK.prove(inst, vars_arr, inputs, gens, oc.Transcript(b"t"), r1cs.tape_seed(1)).ser()
    b = r1cs.NIZK.prove(inst, vars_arr, inputs, gens, oc.Transcript(b"t"), r1cs.tape_seed(1)).ser()
    c = r1cs.NIZK.prove(inst, vars_arr, inputs, gens, oc.Transcript(b"t"), r1cs.tape_seed(2)).ser()
    assert a == b and a != c


def test_padding_edge_case_one_constraint_no_vars():
    """lib.rs:672-752 test_padded_constraints: num_cons = 1, num_vars = 0, three inputs, a^2 + b + 13 = z"""
    def sb(v):
        return (v % Q).to_bytes(32, "little")
    num_cons, num_vars, num_inputs, nz = 1, 0, 3, 3
    A = [(0, num_vars + 2, sb(1))]
    B = [(0, num_vars + 2, sb(1))]
    Cm = [(0, num_vars + 1, sb(1)), (0, num_vars, sb(-13)), (0, num_vars + 3, sb(-1))]
    inst = r1cs.Instance.new(num_cons, num_vars, num_inputs, A, B, Cm)
    assert (inst.inst.num_cons, inst.inst.num_vars) == (2, 4)
    vars_arr = oc.zeros(0)
    inputs = [16, 1, 2]
    assert inst.is_sat(vars_arr, inputs)
    sgens = spark.SNARKGens(num_cons, num_vars, num_inputs, nz)
    comm, decomm = spark.SNARK.encode(inst, sgens)
    sp = spark.SNARK.prove(inst, comm, decomm, vars_arr, inputs, sgens, oc.Transcript(b"snark_example"), r1cs.tape_seed(0))
    sp.verify(comm, inputs, oc.Transcript(b"snark_example"), sgens)
    gens = r1cs.NIZKGens(num_cons, num_vars, num_inputs)
    proof = r1cs.NIZK.prove(inst, vars_arr, inputs, gens, oc.Transcript(b"nizk_example"), r1cs.tape_seed(0))
    proof.verify(inst, inputs, oc.Transcript(b"nizk_example"), gens)


def test_instance_new_errors():
    """lib.rs:627-670"""
    one = (1).to_bytes(32, "little")
    with pytest.raises(r1cs.R1CSError):
        r1cs.Instance.new(4, 4, 1, [(4, 0, one)], [], [])
    with pytest.raises(r1cs.R1CSError):
        r1cs.Instance.new(4, 4, 1, [(0, 6, one)], [], [])
    with pytest.raises(r1cs.R1CSError):
        r1cs.Instance.new(4, 4, 1, [(0, 0, b"\xff" * 32)], [], [])


def _len_product_layer(log_ops, log_cells, ni=3):
    def batched(nprod, ndotp, layers):
        total = 8
        for r in range(layers):
            total += 8 + r * (8 + 3 * 32) + 2 * (8 + nprod * 32)
        return total + 3 * (8 + ndotp * 32)
    head = 2 * (32 + 2 * (8 + ni * 32) + 32) + 2 * (8 + ni * 32)
    return head + batched(4, 0, log_cells) + batched(4 * ni, 2 * ni, log_ops)


def test_published_spark_lengths_follow_from_layout():
    """README.md:371 len_product_layer_proof 64712 at 2^20 (num_ops 2^20, num_mem_cells 2^21)"""
    assert _len_product_layer(20, 21) == 64712


@pytest.mark.parametrize("logn", [4, 8, 10])
def test_snark_roundtrip_and_layout(logn):
    """lib.rs:594-625 (SNARK) / sparse_mlpoly.rs:1602-1666: encode, prove, verify; eval-proof length matches the layout"""
    n = 1 << logn
    ni = min(10, n - 1)
    inst, vars_arr, inputs = r1cs.Instance.produce_synthetic_r1cs(n, n, ni, 2)
    gens = spark.SNARKGens(n, n, ni, n)
    comm, decomm = spark.SNARK.encode(inst, gens)
    proof = spark.SNARK.prove(inst, comm, decomm, vars_arr, inputs, gens, oc.Transcript(b"snark_example"), r1cs.tape_seed(2))
    proof.verify(comm, inputs, oc.Transcript(b"snark_example"), gens)
    pl = proof.r1cs_eval_proof.poly_eval_network_proof.proof_prod_layer
    assert len(pr.ser(pl)) == _len_product_layer(logn, logn + 1)
    with pytest.raises((pr.ProofVerifyError, AssertionError)):
        proof.verify(comm, inputs, oc.Transcript(b"other"), gens)
