#!/bin/bash
# Round 2: launch order of the CTA tiles of the shared-ring 3D kernel (x-fastest / y-fastest / 2 x 2 blocks) on c5, sustained.
O=gpurun_out/r02_call15
mkdir -p $O
export MIN_SECONDS=1.0
for x in "" "DRS_S3C_YFAST" "DRS_S3C_BLOCKED" ""; do
  echo "== order: ${x:-x-fastest (default)}"
  DRS_EXTRA_DEFINES="$x" python tools/time_presets.py c5 c4
done 2>&1 | tee $O/cta_order.txt
