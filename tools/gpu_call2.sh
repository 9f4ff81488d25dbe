#!/bin/bash
# round-2 GPU call 2 (N GPUs, default 2): the sharded prover — byte parity against the single-GPU proof, the oracle and the golden fixtures
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi topo -m > gpurun_out/c2_topo.txt 2>&1
( timeout 600 $TR --master-port 29601 tools/run_sharded.py --logn 16 --oracle --golden tests/golden/snark_proof_sha256.json --reps 2 > gpurun_out/c2_sharded_n${N}_16.txt 2>&1 )
tail -5 gpurun_out/c2_sharded_n${N}_16.txt
( timeout 600 $TR --master-port 29602 tools/run_sharded.py --logn 16 --oracle --nizk --reps 2 > gpurun_out/c2_sharded_n${N}_nizk16.txt 2>&1 )
tail -3 gpurun_out/c2_sharded_n${N}_nizk16.txt
( timeout 900 $TR --master-port 29603 tools/run_sharded.py --logn 18 20 --golden tests/golden/snark_proof_sha256.json --reps 3 > gpurun_out/c2_sharded_n${N}_18_20.txt 2>&1 )
tail -4 gpurun_out/c2_sharded_n${N}_18_20.txt
( SP_BENCH_SHARDED=1 timeout 900 $TR --master-port 29604 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/c2_bench_n${N}.json 2> gpurun_out/c2_bench_n${N}.err )
tail -c 1500 gpurun_out/c2_bench_n${N}.json; tail -3 gpurun_out/c2_bench_n${N}.err
# piggy-back (one process, GPU 0): persistent-tail v2 of the small batched rounds
( SP_SC_PERSIST=1 timeout 300 python -m pytest tests/test_gpu_snark.py tests/test_gpu_golden.py -m gpu -x -q -k "not 20" > gpurun_out/c2_pytest_persist.txt 2>&1 )
tail -2 gpurun_out/c2_pytest_persist.txt
( SP_SC_PERSIST=1 SP_FINE_TIMERS=1 timeout 300 python tools/profile_snark.py 20 > gpurun_out/c2_profile20_persist.txt 2>&1 )
( SP_FINE_TIMERS=1 timeout 300 python tools/profile_snark.py 20 > gpurun_out/c2_profile20.txt 2>&1 )
grep -E "SNARK 2|sc_persist|sc_fold_eval" gpurun_out/c2_profile20_persist.txt gpurun_out/c2_profile20.txt
