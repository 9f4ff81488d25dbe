#!/bin/bash
# round-2 GPU call 24 (one GPU): _dev = witness upload pipelined with the witness commitment (end-to-end path); bench e2e of both builds + GPU tests on _dev
mkdir -p gpurun_out
for tag in "" _dev "" _dev; do
  ( SP_LIB_TAG=$tag timeout 600 python bench.py --steps 10 --warmup 3 --no-msm-var > gpurun_out/c24_bench$tag.json 2> gpurun_out/c24_bench$tag.err )
  python - "$tag" <<'PY'
import json, sys
d=json.loads(open('gpurun_out/c24_bench%s.json' % sys.argv[1]).read().strip().splitlines()[-1])
print(repr(sys.argv[1]).ljust(8), 'device', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3))
PY
done
( SP_LIB_TAG=_dev timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/c24_pytest_dev.txt 2>&1 ); tail -4 gpurun_out/c24_pytest_dev.txt
