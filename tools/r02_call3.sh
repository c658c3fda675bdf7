#!/bin/bash
# Round 2, GPU call 3 (one GPU): full GPU test suite, the reference's own search space timed on B200,
# the engine's tuner for every BASELINE workload (sustained timing, ncu for the winners).
O=gpurun_out/r02_call3
mkdir -p $O
t0=$(date +%s)
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$? seconds=$(( $(date +%s) - t0 ))" >> $O/pytest_gpu.log
tail -4 $O/pytest_gpu.log
if [ -n "$REF_TUNE" ]; then
t1=$(date +%s)
timeout 900 python oracle/tune_ref.py time $O/ref_tune.json > $O/ref_tune.log 2>&1; echo "ref tune rc=$? seconds=$(( $(date +%s) - t1 ))" >> $O/ref_tune.log
grep -E "^ref_tune|rc=" $O/ref_tune.log
fi
ROOT=$(pwd); cd $O
tune() {  # name stc budget extra...
  local wl=$1 stc=$2 budget=$3; shift 3
  local t=$(date +%s)
  PYTHONPATH=$ROOT timeout $(( budget * 2 + 400 )) python -m drstencil_b200.tuner.tune ../../stc/baseline/$stc --budget-s $budget --top 3 --ncu --min-seconds 0.5 \
      --out tune_$wl.json "$@" > tune_$wl.log 2>&1
  echo "tune $wl rc=$? seconds=$(( $(date +%s) - t ))"; grep WINNER tune_$wl.log
}
tune c5 c5_3d7pt_star.stc 330 --first fu1d0bx32y4sn64u4bmx1bmy1mf5ry8sx2sy2 fu1d0bx32y4sn64u4bmx1bmy1mf5ry6 --ncu-size 256 1536 1536
tune c4 c4_3d7pt_star.stc 150 --first fu1d0bx32y2sn16u4bmx1bmy1mf5ry4
tune c1 c1_2d5pt_star.stc 150 --first fu1d0bx64sn128u4bmx2mf5st2
tune c2 c2_2d9pt_box.stc 150 --step 4 --first fu4d0bx64sn256u4bmx2mf5st2
tune c3 c3_2d25pt_box.stc 150 --dtype f32 --first fu1d0bx64sn32u8bmx1mf5st2mb4f32
cd ../..
echo "total seconds $(( $(date +%s) - t0 ))"
