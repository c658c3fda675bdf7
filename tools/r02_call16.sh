#!/bin/bash
# Round 2: (1) the one-GPU emulated-ranks peer-store tests, (2) the fused 3D temporal kernel's store path A/B
# (DRS_T3_LEANSTORE 0 / 1, sustained), (3) parity of the lean store path (3D temporal + peer tests).
O=gpurun_out/r02_call16
mkdir -p $O
timeout 200 python -m pytest tests/test_peer_store_gpu.py -q > $O/peer_default.log 2>&1; echo "rc=$?" >> $O/peer_default.log
tail -3 $O/peer_default.log
export MIN_SECONDS=1.0
for v in 0 1 0 1; do
  echo "== DRS_T3_LEANSTORE=$v"
  DRS_EXTRA_DEFINES="DRS_T3_LEANSTORE=$v" timeout 120 python tools/time_presets.py c5t2 c4t2
done 2>&1 | tee $O/leanstore_ab.txt
DRS_EXTRA_DEFINES="DRS_T3_LEANSTORE=1" timeout 300 python -m pytest tests/test_peer_store_gpu.py tests/test_parity_gpu.py -q \
  -k "temporal or peer or whole_3d" > $O/lean_parity.log 2>&1; echo "rc=$?" >> $O/lean_parity.log
tail -3 $O/lean_parity.log
