#!/bin/bash
# round-2 GPU call 23 (one GPU): last full check of the committed tree — GPU test-suite, smoke, bench line
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/c23_pytest.txt 2>&1 ); tail -4 gpurun_out/c23_pytest.txt
( timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c23_smoke.txt 2>&1 ); tail -1 gpurun_out/c23_smoke.txt
( timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/c23_bench_n1.json 2> gpurun_out/c23_bench_n1.err )
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c23_bench_n1.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['roofline']['largest_launch']['frac'], d['clocks'])
PY
tail -2 gpurun_out/c23_bench_n1.err
