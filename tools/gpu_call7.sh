#!/bin/bash
# round-2 GPU call 7 (one GPU): _dev build = multi-window quad-lane inner-product MSM + green-context partition for the background stream (SP_BG_SMS)
mkdir -p gpurun_out
AB=gpurun_out/c7_ab.txt; : > $AB
run_ab() { ( env "$@" timeout 300 python tools/ab_prove.py "$*" 20 7 >> $AB 2>> gpurun_out/c7_ab.err ); }
run_ab SP_LIB_TAG=
run_ab SP_LIB_TAG=_dev
run_ab SP_LIB_TAG=_dev SP_NO_EARLY_DEREFS=1
run_ab SP_LIB_TAG=_dev SP_NO_EARLY_DEREFS=1 SP_IPA_QUAD=0
run_ab SP_LIB_TAG=_dev SP_BG_SMS=128
run_ab SP_LIB_TAG=_dev SP_BG_SMS=136
run_ab SP_LIB_TAG=_dev SP_BG_SMS=120
run_ab SP_LIB_TAG=_dev SP_BG_SMS=112
run_ab SP_LIB_TAG=_dev SP_BG_SMS=128 SP_IPA_QUAD=0
run_ab SP_LIB_TAG=_dev SP_BG_SMS=128 SP_EARLY_MSM_SMEM=61440
cut -c1-200 $AB; tail -3 gpurun_out/c7_ab.err
( SP_LIB_TAG=_dev SP_BG_SMS=128 SP_FINE_TIMERS=1 timeout 300 python tools/profile_snark.py 20 > gpurun_out/c7_profile_bg128.txt 2>&1 ); tail -24 gpurun_out/c7_profile_bg128.txt | cut -c1-1800
( SP_LIB_TAG=_dev SP_NO_EARLY_DEREFS=1 SP_FINE_TIMERS=1 timeout 300 python tools/profile_snark.py 20 > gpurun_out/c7_profile_noearly.txt 2>&1 ); tail -24 gpurun_out/c7_profile_noearly.txt | cut -c1-1800
( SP_LIB_TAG=_dev timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ipa_msm -s 30 -c 1 -o gpurun_out/c7_ipa -f python tools/profile_snark.py 20 > gpurun_out/c7_ncu_ipa.log 2>&1 )
python tools/ncu_summary.py full gpurun_out/c7_ipa.ncu-rep "k_ipa_msm_quad<15,6>, 4096 generators" > gpurun_out/c7_ncu_ipa.txt 2>&1; rm -f gpurun_out/c7_ipa.ncu-rep; head -30 gpurun_out/c7_ncu_ipa.txt
( SP_LIB_TAG=_dev SP_BG_SMS=128 timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/c7_pytest_bg128.txt 2>&1 )
tail -5 gpurun_out/c7_pytest_bg128.txt
du -sh gpurun_out
