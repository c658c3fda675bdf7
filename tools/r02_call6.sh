#!/bin/bash
# Round 2, GPU call 6 (one GPU): the unaligned-pitch variants (per-row TMA and cp.async): parity and the cliff test.
O=gpurun_out/r02_call6
mkdir -p $O
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -s -k "unaligned" > $O/pytest_unaligned.log 2>&1; echo "rc=$?" >> $O/pytest_unaligned.log
grep -E "unaligned / aligned|passed|failed|rc=|Error" $O/pytest_unaligned.log | head -30
