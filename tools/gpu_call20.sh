#!/bin/bash
# round-2 GPU call 20 (one GPU): final refresh — stand-alone 2^22 proof, bench line, GPU tests, smoke
mkdir -p gpurun_out
( timeout 400 python tools/ab_prove.py final22 22 3 > gpurun_out/c20_ab22.txt 2> gpurun_out/c20_ab22.err ); cut -c1-900 gpurun_out/c20_ab22.txt; tail -2 gpurun_out/c20_ab22.err
( timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/c20_bench_n1.json 2> gpurun_out/c20_bench_n1.err )
tail -c 300 gpurun_out/c20_bench_n1.json; tail -2 gpurun_out/c20_bench_n1.err
( timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/c20_pytest.txt 2>&1 ); tail -4 gpurun_out/c20_pytest.txt
( timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c20_smoke.txt 2>&1 ); tail -1 gpurun_out/c20_smoke.txt
