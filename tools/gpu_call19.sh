#!/bin/bash
# round-2 GPU call 19 (8 GPUs): the final build through the sharded prover at N = 8 — configs[4] again (2^22, golden bytes), the 2^20 strong point, the bench line
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
( timeout 500 $TR --nproc-per-node 8 --master-port 29631 tools/run_sharded.py --logn 20 22 --golden tests/golden/snark_proof_sha256.json --reps 2 > gpurun_out/c19_sharded_n8_20_22.txt 2>&1 )
grep "^{" gpurun_out/c19_sharded_n8_20_22.txt | cut -c1-400; tail -2 gpurun_out/c19_sharded_n8_20_22.txt | cut -c1-300
( timeout 500 $TR --nproc-per-node 8 --master-port 29633 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/c19_bench_n8.json 2> gpurun_out/c19_bench_n8.err )
tail -c 500 gpurun_out/c19_bench_n8.json; tail -3 gpurun_out/c19_bench_n8.err | cut -c1-300
