"""One proof over N GPUs (intra-proof sharding, DESIGN.md "Multi-GPU"): launched under torchrun, one rank per GPU.
Every rank proves the same instance twice — sharded over all ranks, then alone on its own GPU — and the two proofs must be the same bytes on
every rank; rank 0 can also diff them against the oracle (--oracle) or a golden fixture (--golden FILE with the sha256 of the proof).
Prints one JSON line per size from rank 0: sharded / single-GPU milliseconds (CUDA events, max over ranks) and the speed-up."""
import argparse, hashlib, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("OMP_WAIT_POLICY", "passive")
import numpy as np
import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--logn", type=int, nargs="+", default=[16])
    ap.add_argument("--oracle", action="store_true", help="rank 0 also diffs against the oracle's proof (CPU, seconds at 2^16, ~20 s at 2^20)")
    ap.add_argument("--golden", default=None, help="JSON fixture {logn: sha256 of the SNARK proof} (tests/golden/snark_proof_sha256.json)")
    ap.add_argument("--nizk", action="store_true")
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    from spartan_b200 import dist as sd
    rank, world, local = sd.init("nccl" if int(os.environ.get("WORLD_SIZE", "1")) > 1 else None)
    import spartan_b200 as sb
    from spartan_b200 import api
    ctx = sb.Context(local if world > 1 else 0)
    sd.connect(ctx)
    golden = json.load(open(args.golden)) if args.golden else {}
    ok = True
    for logn in args.logn:
        n = 1 << logn
        inst, vars_, inputs = sb.Instance.produce_synthetic_r1cs(n, n, 10, seed=0, ctx=ctx)
        seed = sb.tape_seed(0)
        if args.nizk:
            gens = sb.NIZKGens(n, n, 10, ctx=ctx)
            prove = lambda v: sb.NIZK.prove(inst, v, inputs, gens, b"example", seed)
        else:
            gens = sb.SNARKGens(n, n, 10, n, ctx=ctx)
            ctx.set_sharding(False)
            comm = sb.SNARK.encode(inst, gens)
            ctx.set_sharding(True)
            prove = lambda v: sb.SNARK.prove(inst, comm, v, inputs, gens, b"example", seed)
        dvars = sb.DensePolynomial(vars_.limbs, ctx=ctx)
        res = {}
        for mode in ("sharded", "single"):
            ctx.set_sharding(mode == "sharded")
            sd.barrier(); torch.cuda.synchronize()
            proof = prove(dvars)
            ts = []
            for _ in range(args.reps):
                sd.barrier(); torch.cuda.synchronize()
                api.timer_start(ctx)
                p2 = prove(dvars)
                ts.append(api.timer_stop_ms(ctx))
                assert p2.bytes == proof.bytes
            res[mode] = (proof.bytes, sd.max_over_ranks([min(ts)])[0], dict(ctx.timings()))
        same = res["sharded"][0] == res["single"][0]
        digest = hashlib.sha256(res["sharded"][0]).hexdigest()
        all_same = sd.max_over_ranks([0.0 if same else 1.0])[0] == 0.0
        line = {"what": ("NIZK" if args.nizk else "SNARK") + "::prove 2^%d, one proof over %d GPUs" % (logn, world), "world": world, "logn": logn,
                "sharded_ms": round(res["sharded"][1], 3), "single_gpu_ms": round(res["single"][1], 3), "speedup": round(res["single"][1] / res["sharded"][1], 3),
                "bytes_identical_to_single_gpu_on_every_rank": all_same, "proof_bytes": len(res["sharded"][0]), "sha256": digest}
        if rank == 0 and str(logn) in golden and not args.nizk:
            line["matches_golden_fixture"] = golden[str(logn)] == digest
            ok = ok and line["matches_golden_fixture"]
        if rank == 0 and args.oracle:
            from oracle.spartan_ref import core as oc, r1cs, spark
            oc.lib.oracle_set_threads(max(1, (os.cpu_count() or 2) // 2))
            oi, ovars, oinputs = r1cs.Instance.produce_synthetic_r1cs(n, n, 10, 0)
            if args.nizk:
                oi.digest = inst.digest
                want = r1cs.NIZK.prove(oi, ovars, oinputs, r1cs.NIZKGens(n, n, 10), oc.Transcript(b"example"), r1cs.tape_seed(0)).ser()
            else:
                og = spark.SNARKGens(n, n, 10, n)
                ocomm, odecomm = spark.SNARK.encode(oi, og)
                want = spark.SNARK.prove(oi, ocomm, odecomm, ovars.copy(), oinputs, og, oc.Transcript(b"example"), r1cs.tape_seed(0)).ser()
            line["bytes_identical_to_oracle"] = want == res["sharded"][0]
            ok = ok and line["bytes_identical_to_oracle"]
        ok = ok and all_same
        if rank == 0:
            line["phases_sharded_ms"] = {k: round(v, 2) for k, v in res["sharded"][2].items() if not k.startswith("fine:")}
            print(json.dumps(line), flush=True)
        del dvars, gens, inst
    sd.finalize()
    if not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
