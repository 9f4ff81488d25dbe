#!/bin/bash
# round-2 GPU call 16 (one GPU): evidence of the final build — GPU test-suite, smoke, ncu launch list + full captures (summarised on the box), bench line, CPU arm
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/c16_pytest.txt 2>&1 ); tail -4 gpurun_out/c16_pytest.txt
( timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c16_smoke.txt 2>&1 ); tail -1 gpurun_out/c16_smoke.txt
bash tools/gpu_call_ncu.sh > gpurun_out/c16_ncu.log 2>&1; tail -2 gpurun_out/c16_ncu.log
( timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/c16_bench_n1.json 2> gpurun_out/c16_bench_n1.err )
tail -c 300 gpurun_out/c16_bench_n1.json; tail -2 gpurun_out/c16_bench_n1.err
( timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/c16_bench_ref.json 2> gpurun_out/c16_bench_ref.err )
cut -c1-400 gpurun_out/c16_bench_ref.json; tail -2 gpurun_out/c16_bench_ref.err
( SP_FINE_TIMERS=1 timeout 300 python tools/profile_snark.py 20 > gpurun_out/c16_profile.txt 2>&1 )
( timeout 300 python tools/profile_snark.py 10 16 18 > gpurun_out/c16_profile_other_sizes.txt 2>&1 )
du -sh gpurun_out
