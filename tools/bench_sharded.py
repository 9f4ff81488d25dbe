"""torchrun --nproc-per-node W tools/bench_sharded.py [logn]: BASELINE.json configs[3]-shaped dense sumcheck (2^logn evaluations, cubic-4 and quad)
with the tables cyclically sharded over W GPUs, row commitments with rows sharded, and the configs[2] standalone 2^24-point variable-base MSM with
points sharded by index range; results checked against the W=1 path on rank 0's GPU.  Prints one JSON line from rank 0."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from spartan_b200 import dist as sd, sharded
import spartan_b200 as sb

logn = int(sys.argv[1]) if len(sys.argv) > 1 else 22
rank, world, local = sd.init()
ctx = sb.Context(local if world > 1 else 0)
be = sharded.GpuBackend(ctx)
coll = sharded.Collective()
rng = np.random.default_rng(0)   # same seed on every rank -> identical full tables
n = 1 << logn


def rand_table(n):
    t = rng.integers(0, 2**63, size=(n, 4), dtype=np.uint64)
    t[:, 3] &= np.uint64(0x0FFFFFFFFFFFFFFF)
    return t


res = {"world": world, "logn": logn}
for kind, nt, name in [(2, 4, "cubic4"), (0, 2, "quad")]:
    tabs = [rand_table(n) for _ in range(nt)]
    ch = list(sb.prg_scalars("r", logn))
    sharded.sharded_sumcheck_rounds(be, coll, kind, tabs, ch)   # warm-up
    sd.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
    evals, finals = sharded.sharded_sumcheck_rounds(be, coll, kind, tabs, ch)
    torch.cuda.synchronize(); sd.barrier(); dt = time.perf_counter() - t0
    if rank == 0:
        class One:   # unsharded check on this GPU
            world, rank = 1, 0
            def all_gather_bytes(self, p): return [bytes(p)]
        e1, f1 = sharded.sharded_sumcheck_rounds(be, One(), kind, tabs, ch)
        ok = all(np.array_equal(np.asarray(a).reshape(3, 4), np.asarray(b).reshape(3, 4)) for a, b in zip(evals, e1)) and np.array_equal(finals, f1)
        res[name] = {"ms_incl_upload": round(dt * 1e3, 2), "matches_single_gpu": bool(ok)}
L, R = 1024, 1024
Z = rand_table(L * R)
gens = sb.MultiCommitGens(R, b"shard-bench", ctx=ctx)
sharded.sharded_commit_rows(be, coll, Z, gens, L, R)
sd.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
got = sharded.sharded_commit_rows(be, coll, Z, gens, L, R)
torch.cuda.synchronize(); sd.barrier(); dt = time.perf_counter() - t0
if rank == 0:
    want = sb.DensePolynomial(Z, ctx=ctx).commit(gens, L, R)
    res["commit_1024x1024"] = {"ms_incl_upload": round(dt * 1e3, 2), "matches_single_gpu": got == want}
del Z, gens
# BASELINE.json configs[2]: standalone variable-base MSM, points and scalars split by index range, partial sums met by a point-add allreduce
mlog = int(os.environ.get("SP_SHARD_MSM_LOGN", "24"))
m = 1 << mlog
lo, hi = sharded.index_range(m, rank, world)
pts = be.points_derive(b"msm-bench", lo, hi)
sc_full = rand_table(m)
sc = sb.DensePolynomial(np.ascontiguousarray(sc_full[lo:hi]), ctx=ctx)
sharded.sharded_msm_var(be, coll, pts, sc)
ts = []
for _ in range(3):
    sd.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
    got = sharded.sharded_msm_var(be, coll, pts, sc)
    torch.cuda.synchronize(); sd.barrier(); ts.append(time.perf_counter() - t0)
if rank == 0:
    from spartan_b200 import api
    full = pts[0] if hi == m else api.Points.derive(m, b"msm-bench", ctx=ctx)
    want = full.msm(sb.DensePolynomial(sc_full, ctx=ctx))
    dt = min(ts)
    res["msm_var_2p%d" % mlog] = {"ms": round(dt * 1e3, 2), "Mpoints_per_s": round(m / dt / 1e6, 1), "Mpoint_adds_per_s_reference_equivalent": round(33 * m / dt / 1e6, 1),
                                  "matches_single_gpu": got == want}
    print(json.dumps(res), flush=True)
sd.finalize()
