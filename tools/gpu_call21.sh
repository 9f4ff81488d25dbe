#!/bin/bash
# round-2 GPU call 21 (one GPU): stand-alone 2^22 proofs after the allocator change (per-proof buffers above 1 GiB stay cached), 2^20 sanity, other sizes
mkdir -p gpurun_out
( timeout 400 python tools/ab_prove.py alloc22 22 7 > gpurun_out/c21_ab22.txt 2> gpurun_out/c21_ab22.err ); cut -c1-700 gpurun_out/c21_ab22.txt; tail -2 gpurun_out/c21_ab22.err
( timeout 300 python tools/ab_prove.py alloc20 20 7 > gpurun_out/c21_ab20.txt 2> gpurun_out/c21_ab20.err ); cut -c1-200 gpurun_out/c21_ab20.txt; tail -2 gpurun_out/c21_ab20.err
( timeout 300 python tools/profile_snark.py 10 16 18 > gpurun_out/c21_profile_other_sizes.txt 2>&1 ); grep "^SNARK" gpurun_out/c21_profile_other_sizes.txt
( timeout 600 python -m pytest tests/test_gpu_snark.py tests/test_gpu_prover.py -m gpu -q > gpurun_out/c21_pytest.txt 2>&1 ); tail -3 gpurun_out/c21_pytest.txt
