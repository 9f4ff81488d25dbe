#!/bin/bash
# round-2 GPU call 17 (one GPU): _dev = phase one's tape-only commitments queued behind the witness commitment's MSM; A/B + GPU tests
mkdir -p gpurun_out
AB=gpurun_out/c17_ab.txt; : > $AB
run_ab() { ( env "$@" timeout 300 python tools/ab_prove.py "$*" 20 9 >> $AB 2>> gpurun_out/c17_ab.err ); }
run_ab SP_LIB_TAG=
run_ab SP_LIB_TAG=_dev
run_ab SP_LIB_TAG=
run_ab SP_LIB_TAG=_dev
python - <<'PY'
import json
for l in open('gpurun_out/c17_ab.txt'):
    d=json.loads(l); ph=d['phases']
    print(d['label'].ljust(22), d['median_ms'], d['best_ms'], [ph.get(k) for k in ('polycommit','prove_sc_phase_one','prove_sc_phase_two','polyeval')], d['sha256'])
PY
tail -3 gpurun_out/c17_ab.err
( SP_LIB_TAG=_dev timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/c17_pytest_dev.txt 2>&1 ); tail -4 gpurun_out/c17_pytest_dev.txt
