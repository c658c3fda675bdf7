#!/bin/bash
# Round 2: why is a deeper input ring slower in the fused 3D temporal kernel?  ncu --set full of 2 and 4 stages on the
# c4 grid (768^3, depth 2); each ncu pass only after the same command exited 0 without ncu.
O=gpurun_out/r02_call18
mkdir -p $O
STC=stc/baseline/c4_3d7pt_star.stc
for st in 2 4; do
  cmd="python -m drstencil_b200.tuner.run_one $STC --launches 6 -- --step 2 --stages $st"
  timeout 120 $cmd > $O/run_st$st.log 2>&1 || { echo "plain run st$st failed"; tail -3 $O/run_st$st.log; continue; }
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:^dr_ -s 3 -c 2 -f -o $O/prof_t3_st$st $cmd > $O/ncu_st$st.log 2>&1
  echo "ncu st$st rc=$?"; tail -2 $O/ncu_st$st.log
done
ls -la $O
