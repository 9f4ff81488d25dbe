"""GPU parity for the verifiers (sp_nizk_verify / sp_snark_verify, csrc/verifier.cpp): proofs from this library and proofs produced by the
oracle (same bincode bytes as the reference) are accepted; every single-byte corruption tried is rejected with the reference's error kinds
(ProofVerifyError::InternalError / DecompressionError, errors.rs:5-12), as are wrong inputs, a wrong transcript label and truncated proofs.
Run on the B200 box: pytest -m gpu."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle.spartan_ref import core as oc  # noqa: E402
from oracle.spartan_ref import r1cs, spark  # noqa: E402


@pytest.fixture(scope="module")
def sb():
    import spartan_b200 as m
    m.default_context()
    return m


def corruptions(data, n, seed):
    """n single-byte flips spread over the proof (every region: commitments, sumcheck messages, sigma proofs, IPA, claims)"""
    rng = np.random.default_rng(seed)
    pos = sorted(set(int(x) for x in rng.integers(0, len(data), size=n)) | {0, len(data) - 1, len(data) // 2})
    for p in pos:
        b = bytearray(data)
        b[p] ^= 1 << int(rng.integers(0, 8))
        yield p, bytes(b)


@pytest.mark.parametrize("num_cons,num_vars,num_inputs,seed", [(16, 16, 3, 0), (256, 256, 10, 1), (64, 256, 7, 3), (512, 32, 5, 4), (2, 2, 1, 5), (1024, 1024, 10, 2)])
def test_nizk_verify(sb, num_cons, num_vars, num_inputs, seed):
    inst, vars_, inputs = sb.Instance.produce_synthetic_r1cs(num_cons, num_vars, num_inputs, seed=seed)
    gens = sb.NIZKGens(num_cons, num_vars, num_inputs)
    proof = sb.NIZK.prove(inst, vars_, inputs, gens, b"nizk_example", sb.tape_seed(seed))
    proof.verify(inst, inputs, b"nizk_example", gens)
    # a proof made by the oracle (= the reference's algorithm on the CPU) for the same statement verifies too
    oi, ovars, oinputs = r1cs.Instance.produce_synthetic_r1cs(num_cons, num_vars, num_inputs, seed)
    oi.digest = inst.digest
    oproof = r1cs.NIZK.prove(oi, ovars, oinputs, r1cs.NIZKGens(num_cons, num_vars, num_inputs), oc.Transcript(b"nizk_example"), r1cs.tape_seed(seed + 11))
    sb.NIZK(oproof.ser()).verify(inst, inputs, b"nizk_example", gens)
    with pytest.raises(sb.ProofVerifyError):
        proof.verify(inst, inputs, b"another label", gens)
    bad_inputs = sb.Assignment(inputs.limbs.copy())
    bad_inputs.limbs[0, 0] ^= np.uint64(1)
    with pytest.raises(sb.ProofVerifyError):
        proof.verify(inst, bad_inputs, b"nizk_example", gens)
    with pytest.raises(sb.ProofVerifyError):
        sb.NIZK(proof.bytes[:-1]).verify(inst, inputs, b"nizk_example", gens)
    with pytest.raises(sb.ProofVerifyError):
        sb.NIZK(proof.bytes + b"\x00").verify(inst, inputs, b"nizk_example", gens)
    for pos, bad in corruptions(proof.bytes, 40 if num_cons <= 256 else 12, seed):
        with pytest.raises(sb.ProofVerifyError):
            sb.NIZK(bad).verify(inst, inputs, b"nizk_example", gens)


@pytest.mark.parametrize("num_cons,num_vars,num_inputs,seed", [(16, 16, 3, 0), (256, 256, 10, 1), (64, 256, 7, 3), (512, 32, 5, 4), (2, 2, 1, 5)])
def test_snark_verify(sb, num_cons, num_vars, num_inputs, seed):
    nz = num_cons
    inst, vars_, inputs = sb.Instance.produce_synthetic_r1cs(num_cons, num_vars, num_inputs, seed=seed)
    gens = sb.SNARKGens(num_cons, num_vars, num_inputs, nz)
    comm = sb.SNARK.encode(inst, gens)
    proof = sb.SNARK.prove(inst, comm, vars_, inputs, gens, b"snark_example", sb.tape_seed(seed))
    proof.verify(comm, inputs, b"snark_example", gens)
    # a verifier that only holds the commitment bytes
    from spartan_b200 import api
    comm_only = api.ComputationCommitment.from_bytes(comm.commitment_bytes())
    assert comm_only.commitment_bytes() == comm.commitment_bytes()
    proof.verify(comm_only, inputs, b"snark_example", gens)
    with pytest.raises(sb.SpartanB200Error):
        sb.SNARK.prove(inst, comm_only, vars_, inputs, gens, b"snark_example", sb.tape_seed(seed))
    oi, ovars, oinputs = r1cs.Instance.produce_synthetic_r1cs(num_cons, num_vars, num_inputs, seed)
    ogens = spark.SNARKGens(num_cons, num_vars, num_inputs, nz)
    ocomm, odecomm = spark.SNARK.encode(oi, ogens)
    oproof = spark.SNARK.prove(oi, ocomm, odecomm, ovars, oinputs, ogens, oc.Transcript(b"snark_example"), r1cs.tape_seed(seed + 11))
    sb.SNARK(oproof.ser()).verify(comm, inputs, b"snark_example", gens)
    with pytest.raises(sb.ProofVerifyError):
        proof.verify(comm, inputs, b"another label", gens)
    bad_inputs = sb.Assignment(inputs.limbs.copy())
    bad_inputs.limbs[0, 0] ^= np.uint64(1)
    with pytest.raises(sb.ProofVerifyError):
        proof.verify(comm, bad_inputs, b"snark_example", gens)
    with pytest.raises(sb.ProofVerifyError):
        sb.SNARK(proof.bytes[:-32]).verify(comm, inputs, b"snark_example", gens)
    # a proof for a different instance of the same shape does not verify against this commitment
    inst2, vars2, inputs2 = sb.Instance.produce_synthetic_r1cs(num_cons, num_vars, num_inputs, seed=seed + 100)
    comm2 = sb.SNARK.encode(inst2, gens)
    proof2 = sb.SNARK.prove(inst2, comm2, vars2, inputs2, gens, b"snark_example", sb.tape_seed(seed))
    proof2.verify(comm2, inputs2, b"snark_example", gens)
    with pytest.raises(sb.ProofVerifyError):
        proof2.verify(comm, inputs2, b"snark_example", gens)
    for pos, bad in corruptions(proof.bytes, 60 if num_cons <= 64 else 25, seed):
        with pytest.raises(sb.ProofVerifyError):
            sb.SNARK(bad).verify(comm, inputs, b"snark_example", gens)


def test_verify_2p16(sb):
    """both verifiers at 2^16 (the sizes where the device MSMs inside the verifier matter)"""
    n = 1 << 16
    inst, vars_, inputs = sb.Instance.produce_synthetic_r1cs(n, n, 10, seed=0)
    g1 = sb.NIZKGens(n, n, 10)
    p1 = sb.NIZK.prove(inst, vars_, inputs, g1, b"v", sb.tape_seed(0))
    p1.verify(inst, inputs, b"v", g1)
    gens = sb.SNARKGens(n, n, 10, n)
    comm = sb.SNARK.encode(inst, gens)
    p2 = sb.SNARK.prove(inst, comm, vars_, inputs, gens, b"v", sb.tape_seed(0))
    p2.verify(comm, inputs, b"v", gens)
    for pos, bad in corruptions(p2.bytes, 6, 1):
        with pytest.raises(sb.ProofVerifyError):
            sb.SNARK(bad).verify(comm, inputs, b"v", gens)
