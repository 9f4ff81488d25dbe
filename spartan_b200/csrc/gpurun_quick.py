import time, sys
sys.path.insert(0, '.')
import spartan_b200 as sb
for logn in (10, 16, 20):
    n = 1 << logn
    t0 = time.time()
    inst, vars_, inputs = sb.Instance.produce_synthetic_r1cs(n, n, 10, seed=0)
    gens = sb.NIZKGens(n, n, 10)
    t1 = time.time()
    for it in range(3):
        t2 = time.time()
        p = sb.NIZK.prove(inst, vars_, inputs, gens, b"example", sb.tape_seed(0))
        t3 = time.time()
    print(logn, "setup %.2fs prove %.1f ms" % (t1 - t0, (t3 - t2) * 1e3), inst.ctx.timings())
