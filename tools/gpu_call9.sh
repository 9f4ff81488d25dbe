#!/bin/bash
# round-2 GPU call 9 (one GPU): _dev build = quad-lane IPA MSM with table entries fetched before the chain + the round's dot products fused into the launch
mkdir -p gpurun_out
AB=gpurun_out/c9_ab.txt; : > $AB
run_ab() { ( env "$@" timeout 300 python tools/ab_prove.py "$*" 20 9 >> $AB 2>> gpurun_out/c9_ab.err ); }
run_ab SP_LIB_TAG=
run_ab SP_LIB_TAG=_dev
run_ab SP_LIB_TAG=_dev SP_IPA_NO_FUSED_DOT=1
run_ab SP_LIB_TAG=_dev SP_IPA_QUAD=0
run_ab SP_LIB_TAG=
run_ab SP_LIB_TAG=_dev
cut -c1-200 $AB; tail -3 gpurun_out/c9_ab.err
( SP_LIB_TAG=_dev SP_FINE_TIMERS=1 timeout 300 python tools/profile_snark.py 20 > gpurun_out/c9_profile_dev.txt 2>&1 ); tail -24 gpurun_out/c9_profile_dev.txt | cut -c1-2500
( SP_LIB_TAG=_dev timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/c9_pytest_dev.txt 2>&1 )
tail -5 gpurun_out/c9_pytest_dev.txt
du -sh gpurun_out
