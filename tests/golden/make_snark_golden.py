"""Golden SNARK proofs at sizes the oracle needs minutes for (BASELINE.json configs[1] and configs[4]): generated ONCE here, on the CPU, by the
oracle (the restatement of the reference — the Rust crate cannot be built in this image), so that GPU boxes can diff the proof bytes of the
2^20 / 2^22 runs (including the 8-GPU sharded run of configs[4]) without spending GPU-box minutes on a CPU prover.

    python tests/golden/make_snark_golden.py 16 18 20 22

Instance seed 0, tape seed 0, transcript label b"example", 10 inputs, num_nz_entries = n (what bench.py and tools/run_sharded.py prove).
Writes tests/golden/snark_proof_sha256.json  {"<logn>": sha256 hex of bincode(SNARK)}  and the full proof of the largest size."""
import hashlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
os.environ.setdefault("OMP_WAIT_POLICY", "passive")
from oracle.spartan_ref import core as oc, r1cs, spark  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "snark_proof_sha256.json")


def main():
    oc.lib.oracle_set_threads(os.cpu_count() or 1)
    have = json.load(open(OUT)) if os.path.exists(OUT) else {}
    for logn in [int(a) for a in sys.argv[1:]]:
        n = 1 << logn
        t0 = time.time()
        inst, v, i = r1cs.Instance.produce_synthetic_r1cs(n, n, 10, 0)
        gens = spark.SNARKGens(n, n, 10, n)
        comm, decomm = spark.SNARK.encode(inst, gens)
        proof = spark.SNARK.prove(inst, comm, decomm, v.copy(), i, gens, oc.Transcript(b"example"), r1cs.tape_seed(0)).ser()
        have[str(logn)] = hashlib.sha256(proof).hexdigest()
        have["%d_commitment" % logn] = hashlib.sha256(comm.ser()).hexdigest()
        have["%d_len" % logn] = len(proof)
        json.dump(have, open(OUT, "w"), indent=1, sort_keys=True)
        if logn >= 20:   # the two BASELINE sizes are kept in full, so a mismatch on the GPU box can be localised to a byte offset
            open(os.path.join(ROOT, "tests", "golden", "snark_2p%d_proof.bin" % logn), "wb").write(proof)
        print("2^%d: %d bytes, sha256 %s (%.0f s)" % (logn, len(proof), have[str(logn)], time.time() - t0), flush=True)


if __name__ == "__main__":
    main()
