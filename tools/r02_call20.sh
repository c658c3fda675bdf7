#!/bin/bash
# Round 2: the evict_last plane loads as the default for large 3D arrays: parity of the hinted kernels on the test grids
# (forced with DRS_EXTRA_DEFINES), then c4 / c4t2 / c5 with and without (DRS_NO_LD_HINT), sustained.
O=gpurun_out/r02_call20
mkdir -p $O
DRS_EXTRA_DEFINES="DRS_LD_HINT=1" timeout 300 python -m pytest tests/test_parity_gpu.py tests/test_peer_store_gpu.py tests/test_shared_ring.py -q \
  -k "3d or peer or shared or whole" > $O/hint_parity.log 2>&1; echo "rc=$?" >> $O/hint_parity.log
tail -3 $O/hint_parity.log
export MIN_SECONDS=0.6
{
echo "== default (hint on)";  timeout 100 python tools/time_presets.py c4 c4t2 c5
echo "== DRS_NO_LD_HINT=1";   DRS_NO_LD_HINT=1 timeout 100 python tools/time_presets.py c4 c4t2 c5
echo "== default (hint on)";  timeout 100 python tools/time_presets.py c4 c4t2 c5
} 2>&1 | tee $O/hint_ab.txt
