#!/bin/bash
# round-2 GPU call 10 (one GPU): _dev build = eq-factored streaming rounds of the batched product-circuit sumchecks (SP_SC_EQFACTOR=1), A/B + parity
mkdir -p gpurun_out
AB=gpurun_out/c10_ab.txt; : > $AB
run_ab() { ( env "$@" timeout 300 python tools/ab_prove.py "$*" 20 9 >> $AB 2>> gpurun_out/c10_ab.err ); }
run_ab SP_LIB_TAG=_dev
run_ab SP_LIB_TAG=_dev SP_SC_EQFACTOR=1
run_ab SP_LIB_TAG=_dev
run_ab SP_LIB_TAG=_dev SP_SC_EQFACTOR=1
cut -c1-200 $AB; tail -3 gpurun_out/c10_ab.err
( SP_LIB_TAG=_dev SP_SC_EQFACTOR=1 SP_FINE_TIMERS=1 timeout 300 python tools/profile_snark.py 20 > gpurun_out/c10_profile_eqf.txt 2>&1 ); tail -24 gpurun_out/c10_profile_eqf.txt | cut -c1-1200
( SP_LIB_TAG=_dev SP_SC_EQFACTOR=1 timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/c10_pytest_eqf.txt 2>&1 )
tail -8 gpurun_out/c10_pytest_eqf.txt
du -sh gpurun_out
