#!/bin/bash
# round-2 GPU call 4 (one GPU): ncu evidence of the final build, the bench line, the CPU arm, the whole GPU test-suite
mkdir -p gpurun_out
bash tools/gpu_call_ncu.sh > gpurun_out/c4_ncu.log 2>&1
( timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/c4_bench_n1.json 2> gpurun_out/c4_bench_n1.err )
tail -c 400 gpurun_out/c4_bench_n1.json; tail -2 gpurun_out/c4_bench_n1.err
( timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/c4_bench_ref.json 2> gpurun_out/c4_bench_ref.err )
cut -c1-600 gpurun_out/c4_bench_ref.json; tail -2 gpurun_out/c4_bench_ref.err
( timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/c4_pytest.txt 2>&1 )
tail -3 gpurun_out/c4_pytest.txt
( timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c4_smoke.txt 2>&1 ); tail -1 gpurun_out/c4_smoke.txt
