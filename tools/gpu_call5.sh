#!/bin/bash
# round-2 GPU call 5 (one GPU): A/B of the early (background-stream) derefs commitment and of the quad-lane inner-product MSM on the _dev build, then the evidence of the main build:
# ncu launch list + full captures, bench line, CPU arm, GPU test-suite (on the _dev build), smoke
mkdir -p gpurun_out
AB=gpurun_out/c5_ab.txt; : > $AB
run_ab() { ( env "$@" timeout 300 python tools/ab_prove.py "$*" 20 7 >> $AB 2>> gpurun_out/c5_ab.err ); }
run_ab SP_LIB_TAG=
run_ab SP_LIB_TAG=_dev SP_NO_EARLY_DEREFS=1 SP_IPA_QUAD=0
run_ab SP_LIB_TAG=_dev SP_IPA_QUAD=0
run_ab SP_LIB_TAG=_dev SP_NO_EARLY_DEREFS=1
run_ab SP_LIB_TAG=_dev
run_ab SP_LIB_TAG=_dev SP_EARLY_MSM_CPT=1
run_ab SP_LIB_TAG=_dev SP_EARLY_MSM_SMEM=61440
run_ab SP_LIB_TAG=_dev SP_EARLY_MSM_SMEM=81920
cut -c1-260 $AB; tail -3 gpurun_out/c5_ab.err
bash tools/gpu_call_ncu.sh > gpurun_out/c5_ncu.log 2>&1
( timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/c5_bench_n1.json 2> gpurun_out/c5_bench_n1.err )
tail -c 400 gpurun_out/c5_bench_n1.json; tail -2 gpurun_out/c5_bench_n1.err
( timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/c5_bench_ref.json 2> gpurun_out/c5_bench_ref.err )
cut -c1-600 gpurun_out/c5_bench_ref.json; tail -2 gpurun_out/c5_bench_ref.err
( SP_LIB_TAG=_dev timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/c5_pytest_dev.txt 2>&1 )
tail -5 gpurun_out/c5_pytest_dev.txt
( timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c5_smoke.txt 2>&1 ); tail -1 gpurun_out/c5_smoke.txt
