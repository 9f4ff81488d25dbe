"""Golden result of BASELINE.json configs[2] as bench.py runs it: MSM over the first 2^24 points of MultiCommitGens::new(2^24, b"msm-bench") with the
scalars numpy.random.default_rng(0) draws (Montgomery limbs, top limb masked below 2^60 so every value is a residue < q) — computed by the
oracle (C Pippenger under OpenMP) on the CPU, a few minutes.  Writes tests/golden/msm_2p24.json {"encoding": hex, "points": n, ...}.
    python tests/golden/make_msm_golden.py [logn=24]"""
import hashlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
from oracle.spartan_ref import core as oc


def main():
    logn = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    n = 1 << logn
    oc.lib.oracle_set_threads(os.cpu_count() or 1)
    t0 = time.time()
    gens = oc.MultiCommitGens.new(n, b"msm-bench")
    t1 = time.time()
    rng = np.random.default_rng(0)
    t = rng.integers(0, 2 ** 63, size=(n, 4), dtype=np.uint64)
    t[:, 3] &= np.uint64(0x0FFFFFFFFFFFFFFF)
    out = {"points": n, "label": "msm-bench", "scalars": "numpy default_rng(0).integers(0, 2**63, (n,4), uint64), limb 3 & 0x0FFFFFFFFFFFFFFF, as Montgomery limbs"}
    out["encoding"] = oc.msm(t, gens.G).compress().hex()
    # the eight index-range slices of the 8-GPU run
    per = n // 8
    out["slices8"] = [oc.msm(np.ascontiguousarray(t[k * per:(k + 1) * per]), gens.G[k * per:(k + 1) * per]).compress().hex() for k in range(8)]
    out["generator_digest"] = hashlib.sha256(b"".join(gens.g(i).compress() for i in (0, 1, n // 2, n - 1))).hexdigest()
    json.dump(out, open(os.path.join(ROOT, "tests", "golden", "msm_2p%d.json" % logn), "w"), indent=1, sort_keys=True)
    print(out["encoding"], "gens %.0f s, msm %.0f s" % (t1 - t0, time.time() - t1))


if __name__ == "__main__":
    main()
