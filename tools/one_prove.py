"""Setup + N SNARK::prove calls at 2^logn (for ncu launch lists: `ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 python tools/one_prove.py 20 1`)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spartan_b200 as sb
logn = int(sys.argv[1]) if len(sys.argv) > 1 else 20
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
n = 1 << logn
inst, vars_, inputs = sb.Instance.produce_synthetic_r1cs(n, n, 10, seed=0)
gens = sb.SNARKGens(n, n, 10, n)
comm = sb.SNARK.encode(inst, gens)
dv = sb.DensePolynomial(vars_.limbs)
for _ in range(reps):
    p = sb.SNARK.prove(inst, comm, dv, inputs, gens, b"example", sb.tape_seed(0))
print(len(p.bytes), sb.kernel_launches())
