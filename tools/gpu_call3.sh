#!/bin/bash
# round-2 GPU call 3 (8 GPUs): BASELINE.json configs[4] — SNARK 2^22 over 8 GPUs, bytes diffed against the oracle's golden proof — plus the
# 2^20 strong-scaling point and the bench line at N = 8
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi topo -m > gpurun_out/c3_topo.txt 2>&1
( timeout 900 $TR --master-port 29611 tools/run_sharded.py --logn 20 22 --golden tests/golden/snark_proof_sha256.json --reps 3 > gpurun_out/c3_sharded_n${N}_20_22.txt 2>&1 )
grep "^{" gpurun_out/c3_sharded_n${N}_20_22.txt | cut -c1-420
( timeout 300 $TR --master-port 29612 tools/run_sharded.py --logn 16 --oracle --nizk --reps 2 > gpurun_out/c3_sharded_n${N}_nizk16.txt 2>&1 )
grep "^{" gpurun_out/c3_sharded_n${N}_nizk16.txt | cut -c1-300
( timeout 900 $TR --master-port 29613 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/c3_bench_n${N}.json 2> gpurun_out/c3_bench_n${N}.err )
tail -c 600 gpurun_out/c3_bench_n${N}.json; tail -3 gpurun_out/c3_bench_n${N}.err
