#!/bin/bash
# round-2 GPU call 13 (one GPU): branch-free quad-lane point arithmetic (main build, 4 window groups) against the divergent version (_s4), timeline, GPU tests
mkdir -p gpurun_out
( SP_LIB_TAG=_tl timeout 300 python tools/one_prove.py 20 2 2>&1 | grep ipa_tl | tail -16 > gpurun_out/c13_ipa_timeline.txt ); cat gpurun_out/c13_ipa_timeline.txt | cut -c1-220
AB=gpurun_out/c13_ab.txt; : > $AB
run_ab() { ( env "$@" timeout 300 python tools/ab_prove.py "$*" 20 9 >> $AB 2>> gpurun_out/c13_ab.err ); }
run_ab SP_LIB_TAG=
run_ab SP_LIB_TAG=_s4
run_ab SP_LIB_TAG=
run_ab SP_LIB_TAG=_s4
python - <<'PY'
import json
for l in open('gpurun_out/c13_ab.txt'):
    d=json.loads(l); ph=d['phases']
    print(d['label'].ljust(22), d['median_ms'], d['best_ms'], [ph.get(k) for k in ('polyeval','  polyeval_derefs(2^23)','  polyeval_ops(2^24)','  polyeval_mem(2^22)')], d['sha256'])
PY
tail -3 gpurun_out/c13_ab.err
( timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/c13_pytest.txt 2>&1 ); tail -4 gpurun_out/c13_pytest.txt
