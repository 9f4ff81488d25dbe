"""small SNARK + NIZK prove / verify, variable-base MSM and batched operators for compute-sanitizer (memcheck / racecheck / initcheck)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import spartan_b200 as sb
from spartan_b200 import api
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
inst, vars_, inputs = sb.Instance.produce_synthetic_r1cs(n, n, 5, seed=0)
gens = sb.SNARKGens(n, n, 5, n)
comm = sb.SNARK.encode(inst, gens)
p = sb.SNARK.prove(inst, comm, vars_, inputs, gens, b"example", sb.tape_seed(0))
p.verify(comm, inputs, b"example", gens)
g2 = sb.NIZKGens(n, n, 5)
q = sb.NIZK.prove(inst, vars_, inputs, g2, b"example", sb.tape_seed(0))
q.verify(inst, inputs, b"example", g2)
# bucket MSM at a few sizes / window widths, incl. an uneven size and small scalars
P = api.Points.derive(3000, b"sanitize")
outs = []
for m, c in [(1, 0), (33, 0), (1000, 6), (3000, 0), (3000, 11)]:
    if c:
        os.environ["SP_PIP_WINDOW"] = str(c)
    elif "SP_PIP_WINDOW" in os.environ:
        del os.environ["SP_PIP_WINDOW"]
    outs.append(P.msm(sb.prg_scalars("s", m, m)))
    small = np.zeros((m, 4), dtype=np.uint64)
    small[:, 0] = np.arange(m) % 3
    outs.append(P.msm(np.array([api.scalar_from_bytes(int(v).to_bytes(32, "little")) for v in small[:, 0]], dtype=np.uint64)))
# batched sumcheck operators with a shared C, through the small-table kernel and the streaming kernel
for logn in (4, 13):
    k = 1 << logn
    A = [sb.DensePolynomial(sb.prg_scalars("a%d" % i, k)) for i in range(3)]
    B = [sb.DensePolynomial(sb.prg_scalars("b%d" % i, k)) for i in range(3)]
    Cc = sb.DensePolynomial(sb.prg_scalars("c", k))
    api.sumcheck_batched_eval(A, B, [Cc, Cc, Cc])
    for j in range(logn - 1):
        api.sumcheck_batched_fold_eval(A, B, [Cc, Cc, Cc], sb.prg_scalars("r", 1, j)[0])
print("ok", len(p.bytes), len(q.bytes), len(outs))
