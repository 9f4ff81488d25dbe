#!/bin/bash
# OUT: where the summaries go (default profiles/; on the GPU box: gpurun_out/prof so that the large .ncu-rep files need not travel back)
OUT=${1:-profiles}
# after tools/gpu_call_ncu.sh has run on the box: turn gpurun_out/r02_* into the committed summaries under profiles/
cd "$(dirname "$0")/.."
mkdir -p $OUT
python tools/ncu_summary.py launches gpurun_out/r02_launches.csv "ncu launch list, round 2 — \`ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 python tools/one_prove.py 20 1\` (setup + ONE SNARK::prove at 2^20; per-launch times are cold-cache and serialised: shares, not absolutes)" > $OUT/r02_ncu_launches_snark_2p20.md
python tools/ncu_summary.py full gpurun_out/r02_fold.ncu-rep "ncu --set full --clock-control none --import-source on: first two k_sc_fold_eval launches of a SNARK::prove at 2^20 (ZK cubic sumcheck, 4 tables of 2^20 then 2^19), round-2 build" > $OUT/r02_ncu_full_sc_fold_eval.txt
python tools/ncu_summary.py full gpurun_out/r02_foldg.ncu-rep "ncu --set full: k_sc_fold_eval_g, the eq-factored fused round of the 12 product circuits of the ops proof (bottom layer, tables of 2^19 -> 2^18 entries, 12 instances x 2 tables + the shared suffix eq table)" > $OUT/r02_ncu_full_sc_fold_eval_g.txt
python tools/ncu_summary.py full gpurun_out/r02_msm.ncu-rep "ncu --set full: first k_msm_rows launches of a SNARK::prove at 2^20, round-2 build" > $OUT/r02_ncu_full_msm_rows.txt
python tools/ncu_summary.py full gpurun_out/r02_ipa.ncu-rep "ncu --set full: one k_ipa_msm_quad launch (inner-product round of the witness evaluation proof), round-2 build" > $OUT/r02_ncu_full_ipa_msm.txt
OUT_DIR=$OUT python - <<'PY'
import json, subprocess, sys
sys.argv = ["x", "%s" % __import__("os").environ.get("OUT_DIR", "profiles")]
sys.path.insert(0, "tools")
import ncu_summary as n
res = {}
for rep in ("gpurun_out/r02_fold.ncu-rep", "gpurun_out/r02_msm.ncu-rep", "gpurun_out/r02_ipa.ncu-rep"):
    try:
        n.traffic_json(rep, "/tmp/_t.json")
        res.update(json.load(open("/tmp/_t.json")))
    except Exception as e:
        print("skip", rep, e)
if "sc_fold_eval" in res:
    res["sc_fold_eval"]["algorithmic_bytes"] = 4 * (1 << 20) * 48.0
json.dump(res, open(sys.argv[1] + "/r02_ncu_traffic.json", "w"), indent=1, sort_keys=True)
PY
ls -la $OUT/r02_*
