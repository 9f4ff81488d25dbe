"""N>1 path on CPU: the cyclic partition + all-gather logic of spartan_b200/sharded.py under gloo, world_size 2, with an oracle-backed
stand-in for the local device operations (the oracle is the checker: the sharded result must equal the unsharded oracle result)."""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, %r)
    import numpy as np
    import torch.distributed as dist
    from spartan_b200 import dist as sd, sharded
    from oracle.spartan_ref import core as oc

    class OracleBackend:   # test stand-in for the GPU: same interface as sharded.GpuBackend
        def poly(self, t): return {"z": np.ascontiguousarray(t).copy()}
        def length(self, p): return len(p["z"])
        def to_numpy(self, p): return p["z"]
        def _ev(self, kind, ps):
            z = [p["z"] for p in ps]
            if kind == 0:
                e0, e2 = oc.sc_eval_quad(z[0], z[1]); e = [e0, e2, 0]
            else:
                e = oc.sc_eval_cubic(z[0], z[1], z[2], z[3] if kind == 2 else None)
            return oc.to_arr(e)
        def sc_eval(self, kind, ps): return self._ev(kind, ps)
        def fold(self, ps, r):
            rr = oc.from_mont_bytes(np.ascontiguousarray(r).tobytes())
            for p in ps: p["z"] = oc.bound_top(p["z"], rr).copy()
        def sc_fold_eval(self, kind, ps, r):
            self.fold(ps, r); return self._ev(kind, ps)
        def commit_rows(self, table, gens, L, R, blinds):
            bl = [0] * L if blinds is None else oc.to_ints(blinds)
            return oc.commit_rows(np.ascontiguousarray(table), L, R, bl, gens)
        def msm_var(self, points, scalars):
            return oc.msm(np.ascontiguousarray(scalars), points).compress()
        def sum_points(self, encodings):
            acc = oc.Point.identity()
            for e in encodings: acc = acc + oc.Point.decompress(e)
            return acc.compress()
        def add(self, a, b):
            return oc.to_arr([(oc.from_mont_bytes(a.tobytes()) + oc.from_mont_bytes(b.tobytes())) %% oc.Q])[0]

    rank, world, _ = sd.init("gloo")
    coll = sharded.Collective()
    be = OracleBackend()
    for kind, nt in [(0, 2), (1, 3), (2, 4)]:
        for logn in (1, 2, 5, 9):
            n = 1 << logn
            tabs = [oc.prg_scalars("t%%d" %% k, n, logn) for k in range(nt)]
            ch = oc.prg_scalars("r", logn, kind)
            evals, finals = sharded.sharded_sumcheck_rounds(be, coll, kind, tabs, list(ch))
            # unsharded reference
            cur = [t.copy() for t in tabs]
            for j in range(logn):
                want = be._ev(kind, [{"z": c} for c in cur])
                assert np.array_equal(np.asarray(evals[j]).reshape(3, 4), want), (kind, logn, j)
                cur = [oc.bound_top(c, oc.arr_get(ch, j)).copy() for c in cur]
            assert np.array_equal(finals, np.stack([c[0] for c in cur]))
    L, R = 8, 16
    gens = oc.MultiCommitGens.new(R, b"shard-test")
    Z = oc.prg_scalars("Z", L * R)
    bl = oc.prg_scalars("b", L)
    got = sharded.sharded_commit_rows(be, coll, Z, gens, L, R, bl)
    assert got == oc.commit_rows(Z, L, R, oc.to_ints(bl), gens)
    # variable-base MSM split by index range (uneven: 1001 points over 2 ranks), partial results met by a point-add allreduce
    n = 1001
    pts = oc.MultiCommitGens.new(n, b"shard-msm").G
    sc = oc.prg_scalars("ms", n)
    lo, hi = sharded.index_range(n, rank, world)
    assert sharded.index_range(n, 0, world)[0] == 0 and sharded.index_range(n, world - 1, world)[1] == n
    got = sharded.sharded_msm_var(be, coll, pts[lo:hi], sc[lo:hi])
    assert got == oc.msm(sc, pts).compress()
    sd.finalize()
    print("rank", rank, "ok")
""")


def test_sharded_sumcheck_and_commit_under_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    procs = []
    for r in range(2):
        e = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29544", RANK=str(r), WORLD_SIZE="2", LOCAL_RANK=str(r))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=e, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=300)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, o
        assert "rank %d ok" % r in o
