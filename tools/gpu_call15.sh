#!/bin/bash
# round-2 GPU call 15 (one GPU): main build (+ single-row commitments through the inner-product round kernel) against the previous main (_prev), tests
mkdir -p gpurun_out
AB=gpurun_out/c15_ab.txt; : > $AB
run_ab() { ( env "$@" timeout 300 python tools/ab_prove.py "$*" 20 9 >> $AB 2>> gpurun_out/c15_ab.err ); }
run_ab SP_LIB_TAG=
run_ab SP_LIB_TAG=_prev
run_ab SP_LIB_TAG=
run_ab SP_LIB_TAG=_prev
python - <<'PY'
import json
for l in open('gpurun_out/c15_ab.txt'):
    d=json.loads(l); ph=d['phases']
    print(d['label'].ljust(22), d['median_ms'], d['best_ms'], [ph.get(k) for k in ('polyeval','  polyeval_derefs(2^23)','  polyeval_ops(2^24)','  polyeval_mem(2^22)')], d['sha256'])
PY
tail -3 gpurun_out/c15_ab.err
( SP_FINE_TIMERS=1 timeout 300 python tools/profile_snark.py 20 > gpurun_out/c15_profile.txt 2>&1 ); tail -24 gpurun_out/c15_profile.txt | cut -c1-1800
( timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/c15_pytest.txt 2>&1 ); tail -4 gpurun_out/c15_pytest.txt
