#!/bin/bash
# round-2 GPU call 26 (4 GPUs): bench line of the final tree at N = 4 (replicas + one proof over 4 GPUs)
mkdir -p gpurun_out
( timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29641 bench.py --gpus 4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c26_bench_n4.json 2> gpurun_out/c26_bench_n4.err )
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c26_bench_n4.json').read().strip().splitlines()[-1])
s=d['strong']
print(d['n_gpus'], round(d['value']), round(d['ms_per_step'],2), 'strong', round(s['ms_per_proof'],2), s['proof_bytes_identical_to_single_gpu_on_every_rank'], 'msm', d['msm_var_2p24']['ms'], d['msm_var_2p24']['result_equals_oracle_golden'])
PY
tail -2 gpurun_out/c26_bench_n4.err | cut -c1-300
