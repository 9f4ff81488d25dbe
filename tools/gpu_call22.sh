#!/bin/bash
# round-2 GPU call 22 (one GPU): k_msm_rows with some of the seven products of the mixed addition out of line (instruction-fetch stalls are its top stall)
mkdir -p gpurun_out
AB=gpurun_out/c22_ab.txt; : > $AB
run_ab() { ( env "$@" timeout 300 python tools/ab_prove.py "$*" 20 7 >> $AB 2>> gpurun_out/c22_ab.err ); }
run_ab SP_NO_EARLY_DEREFS=1 SP_LIB_TAG=
run_ab SP_NO_EARLY_DEREFS=1 SP_LIB_TAG=_ni7
run_ab SP_NO_EARLY_DEREFS=1 SP_LIB_TAG=_ni78
run_ab SP_NO_EARLY_DEREFS=1 SP_LIB_TAG=_ni7f
run_ab SP_NO_EARLY_DEREFS=1 SP_LIB_TAG=_ni5
run_ab SP_LIB_TAG=
run_ab SP_LIB_TAG=_ni78
python - <<'PY'
import json
for l in open('gpurun_out/c22_ab.txt'):
    d=json.loads(l); ph=d['phases']
    print(d['label'].ljust(44), d['median_ms'], d['best_ms'], [ph.get(k) for k in ('polycommit','commit_nondet_witness')], d['sha256'])
PY
tail -3 gpurun_out/c22_ab.err
