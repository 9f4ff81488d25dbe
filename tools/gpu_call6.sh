#!/bin/bash
# round-2 GPU call 6 (one GPU): the build with the background derefs commitment (3 MSM CTAs/SM), the quad-lane inner-product MSM and the overlapped
# tape draw — A/B of its switches, per-family profile, ncu evidence (summarised on the box), bench line, GPU test-suite, smoke
mkdir -p gpurun_out
AB=gpurun_out/c6_ab.txt; : > $AB
run_ab() { ( env "$@" timeout 300 python tools/ab_prove.py "$*" 20 7 >> $AB 2>> gpurun_out/c6_ab.err ); }
run_ab SP_X=default
run_ab SP_NO_EARLY_DEREFS=1 SP_IPA_QUAD=0
run_ab SP_IPA_QUAD=0
run_ab SP_EARLY_MSM_SMEM=0
run_ab SP_EARLY_MSM_SMEM=58368
run_ab SP_EARLY_MSM_SMEM=70000
cut -c1-200 $AB; tail -3 gpurun_out/c6_ab.err
( SP_FINE_TIMERS=1 timeout 300 python tools/profile_snark.py 20 > gpurun_out/c6_profile_snark_2p20.txt 2>&1 ); tail -32 gpurun_out/c6_profile_snark_2p20.txt | cut -c1-1500
( SP_FINE_TIMERS=1 SP_NO_EARLY_DEREFS=1 timeout 300 python tools/profile_snark.py 20 > gpurun_out/c6_profile_snark_2p20_noearly.txt 2>&1 )
bash tools/gpu_call_ncu.sh > gpurun_out/c6_ncu.log 2>&1; tail -3 gpurun_out/c6_ncu.log
( timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/c6_bench_n1.json 2> gpurun_out/c6_bench_n1.err )
tail -c 300 gpurun_out/c6_bench_n1.json; tail -2 gpurun_out/c6_bench_n1.err
( timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/c6_pytest.txt 2>&1 )
tail -5 gpurun_out/c6_pytest.txt
( timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c6_smoke.txt 2>&1 ); tail -1 gpurun_out/c6_smoke.txt
du -sh gpurun_out
