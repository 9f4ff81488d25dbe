#!/bin/bash
# ncu evidence for the round: launch list of one bench step + full captures of the dominant kernels (single GPU; never a bench number)
mkdir -p gpurun_out
export SP_BENCH_NO_EXTRA=1
( timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02_launches.csv python tools/one_prove.py 20 1 > gpurun_out/r02_ncu_launch.log 2>&1 )
( timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sc_fold_eval -s 0 -c 2 -o gpurun_out/r02_fold -f python tools/profile_snark.py 20 > gpurun_out/r02_ncu_fold.log 2>&1 )
( timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sc_fold_eval_g -s 10 -c 1 -o gpurun_out/r02_foldg -f python tools/profile_snark.py 20 > gpurun_out/r02_ncu_foldg.log 2>&1 )
( timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_msm_rows -s 0 -c 3 -o gpurun_out/r02_msm -f python tools/profile_snark.py 20 > gpurun_out/r02_ncu_msm.log 2>&1 )
( timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ipa_msm -s 2 -c 1 -o gpurun_out/r02_ipa -f python tools/profile_snark.py 20 > gpurun_out/r02_ncu_ipa.log 2>&1 )
ls -la gpurun_out/*.ncu-rep
# summarise on the box, then drop the reports that would push gpurun_out over the 64 MiB that travel back
bash tools/make_r02_profiles.sh gpurun_out/prof > gpurun_out/r02_make_profiles.log 2>&1
for f in gpurun_out/*.ncu-rep; do
  ncu -i $f --page details > ${f%.ncu-rep}_details.txt 2>/dev/null
  if [ $(stat -c %s $f) -gt 12000000 ]; then rm -f $f; fi
done
du -sh gpurun_out
