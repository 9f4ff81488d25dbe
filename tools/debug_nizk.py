import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["SP_DEBUG"] = "1"
import numpy as np
import spartan_b200 as sb
from oracle.spartan_ref import core as oc, protocol as pr, r1cs
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
oi, ovars, oinputs = r1cs.Instance.produce_synthetic_r1cs(n, n, 3, 1)
ogens = r1cs.NIZKGens(n, n, 3)
# instrument the oracle
orig = pr.UniPoly.from_evals
first = [True]
def patched(e):
    p = orig(e)
    if first[0]:
        first[0] = False
        print("ORACLE evals", [x.to_bytes(32, "little").hex() for x in e])
        print("ORACLE coeffs", [x.to_bytes(32, "little").hex() for x in p.coeffs])
    return p
pr.UniPoly.from_evals = staticmethod(patched)
oproof = r1cs.NIZK.prove(oi, ovars, oinputs, ogens, oc.Transcript(b"example"), r1cs.tape_seed(1))
print("ORACLE comm_poly", oproof.r1cs_sat_proof.sc_proof_phase1.comm_polys[0].hex())
inst, vars_, inputs = sb.Instance.produce_synthetic_r1cs(n, n, 3, seed=1)
inst.set_digest(oi.digest)
gens = sb.NIZKGens(n, n, 3)
proof = sb.NIZK.prove(inst, vars_, inputs, gens, b"example", sb.tape_seed(1))
w = oproof.ser()
print("equal", proof.bytes == w, "first diff", next((i for i in range(len(w)) if w[i] != proof.bytes[i]), None))
