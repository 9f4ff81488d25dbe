"""small SNARK + NIZK prove for compute-sanitizer (memcheck / racecheck / initcheck)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spartan_b200 as sb
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
inst, vars_, inputs = sb.Instance.produce_synthetic_r1cs(n, n, 5, seed=0)
gens = sb.SNARKGens(n, n, 5, n)
comm = sb.SNARK.encode(inst, gens)
p = sb.SNARK.prove(inst, comm, vars_, inputs, gens, b"example", sb.tape_seed(0))
g2 = sb.NIZKGens(n, n, 5)
q = sb.NIZK.prove(inst, vars_, inputs, g2, b"example", sb.tape_seed(0))
print("ok", len(p.bytes), len(q.bytes))
