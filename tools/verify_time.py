"""wall time of the library's verifiers on one B200: NIZK::verify / SNARK::verify at 2^LOGN (reference README.md at 2^20: 414.5 ms / 103.05 ms on its CPU)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spartan_b200 as sb
logn = int(sys.argv[1]) if len(sys.argv) > 1 else 20
n = 1 << logn
inst, vars_, inputs = sb.Instance.produce_synthetic_r1cs(n, n, 10, seed=0)
g1 = sb.NIZKGens(n, n, 10)
p1 = sb.NIZK.prove(inst, vars_, inputs, g1, b"v", sb.tape_seed(0))
p1.verify(inst, inputs, b"v", g1)
t0 = time.perf_counter(); p1.verify(inst, inputs, b"v", g1); t1 = time.perf_counter()
print("NIZK::verify 2^%d: %.1f ms (proof %d B)" % (logn, (t1 - t0) * 1e3, len(p1.bytes)))
del g1
gens = sb.SNARKGens(n, n, 10, n)
comm = sb.SNARK.encode(inst, gens)
p2 = sb.SNARK.prove(inst, comm, vars_, inputs, gens, b"v", sb.tape_seed(0))
p2.verify(comm, inputs, b"v", gens)
t0 = time.perf_counter(); p2.verify(comm, inputs, b"v", gens); t1 = time.perf_counter()
print("SNARK::verify 2^%d: %.1f ms (proof %d B)" % (logn, (t1 - t0) * 1e3, len(p2.bytes)))
