#!/bin/bash
# round-2 GPU call 1: parity at the bench size, latency probes, A/B of the sumcheck kernel variants
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/c1_smi.txt 2>&1
( timeout 300 tools/probe/probe > gpurun_out/c1_probe.txt 2>&1 )
for v in nocf persist cf v2 tma tma32; do
  case $v in
    nocf) export SP_SC_CONSTFOLD=0; unset SP_SC_V2 SP_SC_TMA SP_LIB_TAG SP_SC_PERSIST;;
    persist) export SP_SC_CONSTFOLD=0 SP_SC_PERSIST=1; unset SP_SC_V2 SP_SC_TMA SP_LIB_TAG;;
    cf) export SP_SC_CONSTFOLD=1; unset SP_SC_V2 SP_SC_TMA SP_LIB_TAG SP_SC_PERSIST;;
    v2) export SP_SC_CONSTFOLD=1 SP_SC_V2=1; unset SP_SC_TMA SP_LIB_TAG;;
    v2b6) export SP_SC_CONSTFOLD=1 SP_SC_V2=1 SP_LIB_TAG=_v2b6; unset SP_SC_TMA;;
    tma) export SP_SC_CONSTFOLD=1 SP_SC_TMA=1; unset SP_SC_V2 SP_LIB_TAG;;
    tma32) export SP_SC_CONSTFOLD=1 SP_SC_TMA=1 SP_LIB_TAG=_v2b6; unset SP_SC_V2;;
  esac
  ( timeout 300 python tools/bench_kernels.py 22 --snark > gpurun_out/c1_kern22_$v.json 2> gpurun_out/c1_kern22_$v.err )
done
unset SP_LIB_TAG SP_SC_V2 SP_SC_CONSTFOLD SP_SC_TMA
export SP_SC_PERSIST=1
( timeout 900 python -m pytest tests/test_gpu_snark.py tests/test_gpu_golden.py -m gpu -x -q -k "not 20" > gpurun_out/c1_pytest_persist.txt 2>&1 )
tail -3 gpurun_out/c1_pytest_persist.txt
( SP_FINE_TIMERS=1 timeout 300 python tools/profile_snark.py 20 > gpurun_out/c1_profile20_persist.txt 2>&1 )
unset SP_SC_PERSIST
export SP_SC_CONSTFOLD=1 SP_SC_TMA=1
( timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_golden.py -m gpu -x -q > gpurun_out/c1_pytest_tma.txt 2>&1 )
( timeout 300 python -m pytest tests/test_gpu_snark.py -m gpu -x -q -k "bytes_match_oracle and not bench and 1024" >> gpurun_out/c1_pytest_tma.txt 2>&1 )
tail -3 gpurun_out/c1_pytest_tma.txt
unset SP_SC_TMA
export SP_SC_CONSTFOLD=1 SP_SC_V2=1
( timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_golden.py -m gpu -x -q > gpurun_out/c1_pytest_v2.txt 2>&1 )
( timeout 300 python -m pytest tests/test_gpu_snark.py -m gpu -x -q -k "bytes_match_oracle and not bench and 1024" >> gpurun_out/c1_pytest_v2.txt 2>&1 )
unset SP_SC_V2
( timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_golden.py tests/test_gpu_prover.py -m gpu -x -q -k "not large" > gpurun_out/c1_pytest_cf.txt 2>&1 )
tail -3 gpurun_out/c1_pytest_cf.txt
( SP_FINE_TIMERS=1 timeout 300 python tools/profile_snark.py 20 > gpurun_out/c1_profile20_cf.txt 2>&1 )
unset SP_SC_CONSTFOLD
( SP_FINE_TIMERS=1 timeout 300 python tools/profile_snark.py 20 > gpurun_out/c1_profile20.txt 2>&1 )
( SP_MSM_WINDOW=15 timeout 300 python tools/profile_snark.py 20 > gpurun_out/c1_profile20_w15.txt 2>&1 )
( SP_MSM_WINDOW=15 timeout 300 python -m pytest tests/test_gpu_snark.py -m gpu -x -q -k "bench_configuration and 16" > gpurun_out/c1_pytest_w15.txt 2>&1 )
tail -2 gpurun_out/c1_pytest_w15.txt
( timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/c1_pytest.txt 2>&1 )
tail -3 gpurun_out/c1_pytest.txt
tail -3 gpurun_out/c1_pytest_v2.txt
