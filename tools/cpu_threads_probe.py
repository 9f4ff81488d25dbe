import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.spartan_ref import core as oc, r1cs, spark
print("cpu_count", os.cpu_count(), "omp max", oc.lib.oracle_max_threads(), "OMP_NUM_THREADS", os.environ.get("OMP_NUM_THREADS"))
n = 1 << 14
inst, vars_arr, inputs = r1cs.Instance.produce_synthetic_r1cs(n, n, 10, 0)
gens = spark.SNARKGens(n, n, 10, n)
comm, decomm = spark.SNARK.encode(inst, gens)
for t in [int(x) for x in sys.argv[1:]]:
    oc.lib.oracle_set_threads(t)
    t0 = time.perf_counter()
    spark.SNARK.prove(inst, comm, decomm, vars_arr.copy(), inputs, gens, oc.Transcript(b"example"), r1cs.tape_seed(0))
    print("threads", t, "2^14 prove %.2f s" % (time.perf_counter() - t0), flush=True)
