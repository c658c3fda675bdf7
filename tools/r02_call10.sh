#!/bin/bash
# Round 2, GPU call 10 (one GPU): c2 re-tuned after the deferred-scale change, then the contract bench.
O=gpurun_out/r02_call10
mkdir -p $O
ROOT=$(pwd)
cd $O
PYTHONPATH=$ROOT timeout 900 python -m drstencil_b200.tuner.tune ../../stc/baseline/c2_2d9pt_box.stc --step 4 --budget-s 200 --top 3 --ncu --min-seconds 0.5 \
    --out tune_c2.json --first fu4d0bx128sn256u4bmx2mf5st2 > tune_c2.log 2>&1; echo "tune c2 rc=$?"; grep WINNER tune_c2.log
cd $ROOT
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_call10/bench.json"))
print("c5 value %.1f frac %.3f e2e %.1f parity %s" % (d["value"], d["roofline"]["frac"], d["e2e"]["value"], d["parity"]["ok"]))
for e in d["per_config"]:
    print("   %-8s %.1f GStencil/s frac %.3f clocks %s" % (e["workload"][:8], e["value"], e["roofline"]["frac"], e["clocks"]["sm_mhz"]))
PY
