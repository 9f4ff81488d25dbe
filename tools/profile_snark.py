import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spartan_b200 as sb
from spartan_b200 import api
for logn in [int(a) for a in sys.argv[1:]] or [10, 16, 20]:
    n = 1 << logn
    t0 = time.time()
    inst, vars_, inputs = sb.Instance.produce_synthetic_r1cs(n, n, 10, seed=0)
    gens = sb.SNARKGens(n, n, 10, n)
    t1 = time.time()
    comm = sb.SNARK.encode(inst, gens)
    t1b = time.time()
    dvars = sb.DensePolynomial(vars_.limbs)
    for it in range(2):
        p = sb.SNARK.prove(inst, comm, dvars, inputs, gens, b"example", sb.tape_seed(0))
    api.timer_start(); t2 = time.time()
    p = sb.SNARK.prove(inst, comm, dvars, inputs, gens, b"example", sb.tape_seed(0))
    ms = api.timer_stop_ms(); t3 = time.time()
    print("SNARK 2^%d setup %.2fs encode %.2fs prove(resident) wall %.2f ms events %.2f ms -> %.3g constraints/s  proof %d B" % (logn, t1 - t0, t1b - t1, (t3 - t2) * 1e3, ms, n / (t3 - t2), len(p.bytes)))
    print("  phases", {k: round(v, 2) for k, v in inst.ctx.timings().items()})
    t2 = time.time(); p = sb.SNARK.prove(inst, comm, vars_, inputs, gens, b"example", sb.tape_seed(0)); t3 = time.time()
    print("  prove(host vars) wall %.2f ms" % ((t3 - t2) * 1e3))
    api.prof_enable(True)
    p = sb.SNARK.prove(inst, comm, dvars, inputs, gens, b"example", sb.tape_seed(0))
    rep = api.prof_report(); api.prof_enable(False)
    tot = sum(v["ms"] for v in rep.values())
    for k, v in sorted(rep.items(), key=lambda kv: -kv[1]["ms"]):
        print("  %-14s launches %4d  %8.3f ms (%4.1f%%)  %8.1f GB/s algorithmic" % (k, v["launches"], v["ms"], 100 * v["ms"] / tot, v["bytes"] / 1e9 / (v["ms"] / 1e3) if v["ms"] else 0))
    print("  sum of kernel time %.2f ms, launches %d" % (tot, sum(v["launches"] for v in rep.values())), flush=True)
