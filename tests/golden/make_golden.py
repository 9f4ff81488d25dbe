"""Regenerates tests/golden/*.json from the oracle (oracle/ = CPU restatement of the reference; /root/reference is Rust and cannot run here, so these
are the oracle's bytes — "parity unpinned" against the Rust binary, see DESIGN.md section 4 — frozen so that neither the oracle nor the CUDA path can
drift silently).  Usage: python tests/golden/make_golden.py        (about two minutes of CPU)"""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import numpy as np  # noqa: E402
from oracle.spartan_ref import core as oc, r1cs, spark  # noqa: E402

NIZK_CASES = [(16, 16, 3, 1), (64, 256, 7, 2), (512, 32, 5, 3), (2, 2, 1, 4), (1024, 1024, 10, 0)]
SNARK_CASES = [(16, 16, 3, 0), (64, 256, 7, 3), (512, 32, 5, 4), (2, 2, 1, 5), (256, 256, 10, 1)]


def digest(b):
    return {"len": len(b), "sha256": hashlib.sha256(b).hexdigest(), "head": b[:48].hex(), "tail": b[-48:].hex()}


def build():
    out = {"generator": "tests/golden/make_golden.py", "transcript_labels": {"nizk": "example", "snark": "snark_example"}, "nizk": [], "snark": [], "msm": [], "sumcheck": []}
    for (nc, nv, ni, seed) in NIZK_CASES:
        inst, vars_arr, inputs = r1cs.Instance.produce_synthetic_r1cs(nc, nv, ni, seed)
        gens = r1cs.NIZKGens(nc, nv, ni)
        inst.digest = b"golden R1CSShapeDigest %d" % seed     # the digest is an opaque input (zlib output in the reference): fixed bytes here
        proof = r1cs.NIZK.prove(inst, vars_arr, inputs, gens, oc.Transcript(b"example"), r1cs.tape_seed(seed))
        proof.verify(inst, inputs, oc.Transcript(b"example"), gens)
        out["nizk"].append({"num_cons": nc, "num_vars": nv, "num_inputs": ni, "seed": seed, "digest": inst.digest.decode(), "proof": digest(proof.ser())})
    for (nc, nv, ni, seed) in SNARK_CASES:
        inst, vars_arr, inputs = r1cs.Instance.produce_synthetic_r1cs(nc, nv, ni, seed)
        gens = spark.SNARKGens(nc, nv, ni, nc)
        comm, decomm = spark.SNARK.encode(inst, gens)
        proof = spark.SNARK.prove(inst, comm, decomm, vars_arr, inputs, gens, oc.Transcript(b"snark_example"), r1cs.tape_seed(seed))
        proof.verify(comm, inputs, oc.Transcript(b"snark_example"), gens)
        out["snark"].append({"num_cons": nc, "num_vars": nv, "num_inputs": ni, "seed": seed, "commitment": digest(comm.ser()), "proof": digest(proof.ser())})
    for n in (1, 2, 33, 190, 1024, 4096):
        g = oc.MultiCommitGens.new(n, b"msm-test")
        sc = oc.prg_scalars("msm", n, n)
        out["msm"].append({"n": n, "label": "msm-test", "scalars": "prg_scalars('msm', n, seed=n)", "result": oc.msm(sc, g.G).compress().hex(),
                           "gens_head": [g.g(i).compress().hex() for i in range(min(n, 3))], "h": g.h.compress().hex()})
    for logn in (3, 10):
        n = 1 << logn
        t = [oc.prg_scalars("sc%d" % k, n, logn) for k in range(4)]
        e0, e2 = oc.sc_eval_quad(t[0], t[1])
        out["sumcheck"].append({"logn": logn, "tables": "prg_scalars('sc<k>', n, seed=logn), k = 0..3",
                                "quad": [hex(e0), hex(e2)], "cubic3": [hex(x) for x in oc.sc_eval_cubic(t[0], t[1], t[2], None)],
                                "cubic4": [hex(x) for x in oc.sc_eval_cubic(t[0], t[1], t[2], t[3])]})
    return out


def main():
    with open(os.path.join(HERE, "oracle_golden.json"), "w") as f:
        json.dump(build(), f, indent=1)
    print("wrote", os.path.join(HERE, "oracle_golden.json"))


if __name__ == "__main__":
    main()
