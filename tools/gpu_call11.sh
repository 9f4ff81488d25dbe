#!/bin/bash
# round-2 GPU call 11 (2 GPUs): the final build (eq-factored rounds, fused dot products, background commitment) through the sharded prover
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_gpu_sharded.py -m gpu -q > gpurun_out/c11_pytest_sharded.txt 2>&1 ); tail -4 gpurun_out/c11_pytest_sharded.txt
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
( timeout 600 $TR --nproc-per-node 2 --master-port 29621 tools/run_sharded.py --logn 20 22 --golden tests/golden/snark_proof_sha256.json --reps 3 > gpurun_out/c11_sharded_n2_20_22.txt 2>&1 )
grep "^{" gpurun_out/c11_sharded_n2_20_22.txt | cut -c1-400; tail -2 gpurun_out/c11_sharded_n2_20_22.txt | cut -c1-300
( timeout 600 $TR --nproc-per-node 2 --master-port 29623 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/c11_bench_n2.json 2> gpurun_out/c11_bench_n2.err )
tail -c 600 gpurun_out/c11_bench_n2.json; tail -3 gpurun_out/c11_bench_n2.err | cut -c1-300
du -sh gpurun_out
