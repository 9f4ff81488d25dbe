"""Standalone variable-base MSM (BASELINE.json configs[2]: N = 2^24 ristretto255 points, 253-bit scalars) on one B200.
Reports ms per MSM, reference-equivalent Mpoint-adds/s := 33*N/t and raw Mpoints/s, per window width.
Usage: bench_msm_var.py [logn ...] [--windows 13,15,16] [--small]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import spartan_b200 as sb
from spartan_b200 import api

logns = [int(a) for a in sys.argv[1:] if a.isdigit()] or [20]
windows = [0]
for i, a in enumerate(sys.argv):
    if a == "--windows":
        windows = [int(x) for x in sys.argv[i + 1].split(",")]
ctx = sb.default_context()
rng = np.random.default_rng(0)


def rand_table(n):
    t = rng.integers(0, 2**63, size=(n, 4), dtype=np.uint64)
    t[:, 3] &= np.uint64(0x0FFFFFFFFFFFFFFF)  # < q: valid Montgomery residues
    return t


for logn in logns:
    n = 1 << logn
    P = api.Points.derive(n, b"msm-bench")
    dists = {"uniform": sb.DensePolynomial(rand_table(n))}
    if "--small" in sys.argv:
        dists["u64"] = sb.DensePolynomial(sb.from_u64(rng.integers(0, 1 << 63, size=n, dtype=np.uint64))) if hasattr(sb, "from_u64") else None
    for name, S in dists.items():
        if S is None:
            continue
        for c in windows:
            if c:
                os.environ["SP_PIP_WINDOW"] = str(c)
            elif "SP_PIP_WINDOW" in os.environ:
                del os.environ["SP_PIP_WINDOW"]
            out = P.msm(S)
            ts = []
            for _ in range(3):
                api.timer_start(); out2 = P.msm(S); ts.append(api.timer_stop_ms())
            assert out == out2
            api.prof_enable(True); P.msm(S); rep = api.prof_report(); api.prof_enable(False)
            ms = min(ts)
            print(json.dumps({"logn": logn, "dist": name, "window": c, "ms": round(ms, 3), "Mpoint_adds_per_s_ref_equiv": round(33 * n / ms / 1e3, 1),
                              "Mpoints_per_s": round(n / ms / 1e3, 1), "out": out.hex()[:16],
                              "stages_ms": {k: round(v["ms"], 3) for k, v in rep.items() if k.startswith("pip") or k == "msm_var"}}), flush=True)
    del P, dists
