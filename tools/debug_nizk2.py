import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spartan_b200 as sb
from oracle.spartan_ref import core as oc, protocol as pr, r1cs
for (nc, nv, ni, seed) in [(1024, 1024, 10, 0), (1024, 1024, 10, 1), (1024, 1024, 3, 0), (16, 16, 3, 0), (16, 16, 10, 1), (64, 256, 7, 2), (512, 32, 5, 3), (2, 2, 1, 4)]:
    oi, ovars, oinputs = r1cs.Instance.produce_synthetic_r1cs(nc, nv, ni, seed)
    ogens = r1cs.NIZKGens(nc, nv, ni)
    oproof = r1cs.NIZK.prove(oi, ovars, oinputs, ogens, oc.Transcript(b"example"), r1cs.tape_seed(seed))
    inst, vars_, inputs = sb.Instance.produce_synthetic_r1cs(nc, nv, ni, seed=seed)
    inst.set_digest(oi.digest)
    gens = sb.NIZKGens(nc, nv, ni)
    proof = sb.NIZK.prove(inst, vars_, inputs, gens, b"example", sb.tape_seed(seed))
    w = oproof.ser()
    print((nc, nv, ni, seed), "equal", proof.bytes == w, "len", len(w), len(proof.bytes), "first diff", next((i for i in range(min(len(w), len(proof.bytes))) if w[i] != proof.bytes[i]), None), flush=True)
