#!/bin/bash
# round-2 GPU call 18 (one GPU): helper threads for the host's multi-term commitments (SP_HOST_THREADS = 0 / 2 / 3 (default) / 5), GPU tests
mkdir -p gpurun_out
nproc > gpurun_out/c18_nproc.txt
AB=gpurun_out/c18_ab.txt; : > $AB
run_ab() { ( env "$@" timeout 300 python tools/ab_prove.py "$*" 20 9 >> $AB 2>> gpurun_out/c18_ab.err ); }
run_ab SP_HOST_THREADS=0
run_ab SP_HOST_THREADS=3
run_ab SP_HOST_THREADS=0
run_ab SP_HOST_THREADS=3
run_ab SP_HOST_THREADS=2
run_ab SP_HOST_THREADS=5
python - <<'PY'
import json
for l in open('gpurun_out/c18_ab.txt'):
    d=json.loads(l); ph=d['phases']
    print(d['label'].ljust(22), d['median_ms'], d['best_ms'], [ph.get(k) for k in ('polycommit','prove_sc_phase_one','prove_sc_phase_two','polyeval')], d['sha256'])
PY
tail -3 gpurun_out/c18_ab.err
( timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/c18_pytest.txt 2>&1 ); tail -4 gpurun_out/c18_pytest.txt
