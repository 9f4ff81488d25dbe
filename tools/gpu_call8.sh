#!/bin/bash
# round-2 GPU call 8 (8 GPUs of one box): BASELINE.json configs[4] — SNARK 2^22 sharded over 8 GPUs, bytes diffed against the oracle's golden proof —,
# the 2^20 strong-scaling points at N = 8 and N = 4, and the bench line at N = 8 and N = 4 with the sharded legs forced on
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/c8_topo.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
( timeout 600 $TR --nproc-per-node 8 --master-port 29611 tools/run_sharded.py --logn 20 22 --golden tests/golden/snark_proof_sha256.json --reps 3 > gpurun_out/c8_sharded_n8_20_22.txt 2>&1 )
grep "^{" gpurun_out/c8_sharded_n8_20_22.txt | cut -c1-420; tail -2 gpurun_out/c8_sharded_n8_20_22.txt | cut -c1-300
( SP_BENCH_SHARDED=1 timeout 600 $TR --nproc-per-node 8 --master-port 29613 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/c8_bench_n8.json 2> gpurun_out/c8_bench_n8.err )
tail -c 700 gpurun_out/c8_bench_n8.json; tail -3 gpurun_out/c8_bench_n8.err | cut -c1-300
( timeout 400 $TR --nproc-per-node 4 --master-port 29614 tools/run_sharded.py --logn 20 --golden tests/golden/snark_proof_sha256.json --reps 3 > gpurun_out/c8_sharded_n4_20.txt 2>&1 )
grep "^{" gpurun_out/c8_sharded_n4_20.txt | cut -c1-420; tail -2 gpurun_out/c8_sharded_n4_20.txt | cut -c1-300
( SP_BENCH_SHARDED=1 timeout 500 $TR --nproc-per-node 4 --master-port 29615 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/c8_bench_n4.json 2> gpurun_out/c8_bench_n4.err )
tail -c 500 gpurun_out/c8_bench_n4.json; tail -3 gpurun_out/c8_bench_n4.err | cut -c1-300
du -sh gpurun_out
