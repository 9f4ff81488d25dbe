#!/bin/bash
# round-2 GPU call 25 (one GPU): ncu --set full captures of the remaining round-2 kernels (summaries made on the box)
mkdir -p gpurun_out
cap() {  # name regex skip count title
  ( timeout 400 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 -o gpurun_out/r02_$1 -f python tools/profile_snark.py 20 > gpurun_out/r02_ncu_$1.log 2>&1 )
  python tools/ncu_summary.py full gpurun_out/r02_$1.ncu-rep "$5" > gpurun_out/r02_ncu_full_$1.txt 2>&1
  ncu -i gpurun_out/r02_$1.ncu-rep --page details > gpurun_out/r02_ncu_details_$1.txt 2>/dev/null
  rm -f gpurun_out/r02_$1.ncu-rep
  head -8 gpurun_out/r02_ncu_full_$1.txt | cut -c1-160
}
cap sc_eval_g k_sc_eval_g 4 1 "ncu --set full: k_sc_eval_g, first (eq-factored) evaluation of the bottom layer of the ops product circuits (12 instances, tables of 2^19 entries), SNARK::prove 2^20"
cap sc_fold_eval_small k_sc_fold_eval_small 40 2 "ncu --set full: two k_sc_fold_eval_small launches (small-table rounds of the batched product-circuit sumchecks), SNARK::prove 2^20"
cap spark_hash k_spark_hash 0 1 "ncu --set full: k_spark_hash (first hash layer of the SPARK memory check), SNARK::prove 2^20"
du -sh gpurun_out
