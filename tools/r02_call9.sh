#!/bin/bash
# Round 2, GPU call 9 (one GPU): deferred-scale scatter forms -- parity of the temporal kernels, A/B timing.
O=gpurun_out/r02_call9
mkdir -p $O
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_ref_gold_gpu.py tests/test_run_host_streamed_gpu.py -m gpu -q -k "temporal or whole or knob or fp32 or ref_gold or fusion or streamed or unaligned" > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest.log
for rep in 1 2; do
  DRS_NO_DEFERRED_SCALE=1 python tools/time_presets.py c2 c3t2 c1t2 c4t2 c5t2 2>&1 | sed 's/^/without deferred scale: /'
  python tools/time_presets.py c2 c3t2 c1t2 c4t2 c5t2 2>&1 | sed 's/^/with    deferred scale: /'
done | tee $O/timing.txt
