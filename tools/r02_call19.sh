#!/bin/bash
# Round 2: L2 policy experiments on the 3D kernels (c5 single step, c5 depth 2), sustained:
# plane loads evict_last / evict_first, output stores st.global.cs.
O=gpurun_out/r02_call19
mkdir -p $O
export MIN_SECONDS=0.6
for x in "" "DRS_ST_HINT=1" "DRS_LD_HINT=1" "DRS_LD_HINT=1;DRS_ST_HINT=1" "DRS_LD_HINT=2" ""; do
  echo "== ${x:-default}"
  DRS_EXTRA_DEFINES="$x" timeout 100 python tools/time_presets.py c5 c5t2
done 2>&1 | tee $O/l2_policy.txt
