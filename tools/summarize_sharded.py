"""gpurun_out/c2_sharded_*.txt (JSON lines of tools/run_sharded.py) -> markdown table for profiles/r02_sharded.md.  Usage: python tools/summarize_sharded.py gpurun_out/c2_sharded_*.txt"""
import json, sys
rows = []
for path in sys.argv[1:]:
    for line in open(path, errors="replace"):
        if line.startswith("{"):
            try:
                rows.append(json.loads(line))
            except ValueError:
                pass
print("| proof | GPUs | one GPU (ms) | sharded (ms) | speed-up | bytes = single-GPU proof on every rank | = oracle | = golden fixture | sha256 |")
print("|---|---|---|---|---|---|---|---|---|")
for r in sorted(rows, key=lambda r: (r["what"][:5], r["logn"], r["world"])):
    print("| %s 2^%d | %d | %.2f | %.2f | %.2fx | %s | %s | %s | %s... |" % (r["what"].split("::")[0], r["logn"], r["world"], r["single_gpu_ms"], r["sharded_ms"], r["speedup"],
          r["bytes_identical_to_single_gpu_on_every_rank"], r.get("bytes_identical_to_oracle", "-"), r.get("matches_golden_fixture", "-"), r["sha256"][:16]))
for r in rows:
    if "phases_sharded_ms" in r:
        print("\nphases of the sharded %s 2^%d on %d GPUs (ms): %s" % (r["what"].split("::")[0], r["logn"], r["world"], json.dumps(r["phases_sharded_ms"])))
